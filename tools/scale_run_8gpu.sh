#!/bin/bash
# One 8-GPU box: the file-driven NetCDF-3 year on 8 ranks (tools/gpu_year_files.py) and the bench at N = 8, 4, 2
# (kernel + e2e + host-link probe; CPU baseline, other variants and other workloads skipped).  Logs -> gpurun_out/.
set -x
mkdir -p gpurun_out
nproc; free -g | head -2; nvidia-smi topo -m > gpurun_out/r02_topo_8gpu.txt 2>&1; lscpu | grep -i "numa\|socket\|model name" > gpurun_out/r02_lscpu_8gpu.txt
D=/tmp/nmp_year
timeout 600 python tools/gpu_year_files.py generate --dir $D --procs $(nproc) 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29577 tools/gpu_year_files.py run --dir $D --tag full > gpurun_out/r02_year_full_8gpu.log 2>&1; echo full rc=$?; tail -1 gpurun_out/r02_year_full_8gpu.log
timeout 600 $TR --nproc-per-node 8 --master-port 29578 tools/gpu_year_files.py run --dir $D --tag resumed --resume > gpurun_out/r02_year_resumed_8gpu.log 2>&1; echo resumed rc=$?; tail -1 gpurun_out/r02_year_resumed_8gpu.log
timeout 300 python tools/gpu_year_files.py compare --dir $D --out gpurun_out/r02_year_files_8gpu.json
rm -rf $D
for N in 8 4 2; do
  timeout 600 $TR --nproc-per-node $N --master-port $((29600+N)) bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-other-variants --no-other-workloads > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo bench $N rc=$?
  python -c "import json;d=json.load(open('gpurun_out/r02_bench_${N}gpu.json'));print(d['n_gpus'],d['value'],d['e2e']['value'],d['e2e'].get('host_link'))"
done
