set -x
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29708 tools/h2d_topology_probe.py 2>/dev/null | tee -a gpurun_out/r02_h2d_topology.jsonl
F="--no-cpu-baseline --no-other-variants --no-other-workloads"
for N in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 5 --warmup 3 $F > gpurun_out/r02_bench_${N}gpu_spread.json 2> gpurun_out/r02_bench_${N}gpu_spread.err; echo bench $N rc=$?
  python -c "import json;d=json.load(open('gpurun_out/r02_bench_${N}gpu_spread.json'));print(d['n_gpus'],d['config']['gpus'],d['value'],d['e2e']['value'],d['e2e']['host_link']['h2d_gbs_all_gpus'])"
done
