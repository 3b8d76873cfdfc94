set -x
mkdir -p gpurun_out
nproc; free -g | head -2; df -h /tmp | tail -1
(timeout 600 python -m pytest tests/test_driver_gpu.py -m gpu -q -k "two_rank" > gpurun_out/r02_two_rank_driver_test.log 2>&1; echo "rc=$?" >> gpurun_out/r02_two_rank_driver_test.log); tail -3 gpurun_out/r02_two_rank_driver_test.log
D=/tmp/nmp_year
timeout 600 python tools/gpu_year_files.py generate --dir $D --procs $(nproc) 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577"
timeout 900 $TR tools/gpu_year_files.py run --dir $D --tag full > gpurun_out/r02_year_full_2gpu.log 2>&1; echo full rc=$?; tail -2 gpurun_out/r02_year_full_2gpu.log
timeout 600 $TR tools/gpu_year_files.py run --dir $D --tag resumed --resume > gpurun_out/r02_year_resumed_2gpu.log 2>&1; echo resumed rc=$?; tail -2 gpurun_out/r02_year_resumed_2gpu.log
timeout 300 python tools/gpu_year_files.py compare --dir $D --out gpurun_out/r02_year_files_2gpu.json
