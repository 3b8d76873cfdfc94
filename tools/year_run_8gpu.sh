#!/bin/bash
# The file-driven NetCDF-3 year (tools/gpu_year_files.py) on the 8 GPUs of one box, full run + resumed run + comparison.
set -x
mkdir -p gpurun_out
D=/tmp/nmp_year
timeout 600 python tools/gpu_year_files.py generate --dir $D --procs $(nproc) 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4; do
  rm -rf $D/out $D/restart $D/resumed
  timeout 600 $TR --nproc-per-node $N --master-port $((29570+N)) tools/gpu_year_files.py run --dir $D --tag full > gpurun_out/r02_year_full_${N}gpu.log 2>&1; echo full rc=$?; tail -1 gpurun_out/r02_year_full_${N}gpu.log
  timeout 600 $TR --nproc-per-node $N --master-port $((29580+N)) tools/gpu_year_files.py run --dir $D --tag resumed --resume > gpurun_out/r02_year_resumed_${N}gpu.log 2>&1; echo resumed rc=$?; tail -1 gpurun_out/r02_year_resumed_${N}gpu.log
  timeout 300 python tools/gpu_year_files.py compare --dir $D --out gpurun_out/r02_year_files_${N}gpu.json
done
rm -rf $D
