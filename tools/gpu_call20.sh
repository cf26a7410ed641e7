set +e
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(time timeout 900 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3) > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
echo "bench8 rc=$?"; tail -4 gpurun_out/r2_bench_8gpu.err
(time timeout 600 $TR --nproc-per-node 4 --master-port 29542 bench.py --gpus 4 --steps 10 --warmup 3 --no-configs) > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err
echo "bench4 rc=$?"; tail -4 gpurun_out/r2_bench_4gpu.err
python - <<'PY'
import json
for n in (8,4):
    b=json.loads(open(f"gpurun_out/r2_bench_{n}gpu.json").read().strip().splitlines()[-1])
    print(n, f"value {b['value']:.3e} e2e {b['e2e']['value']:.3e} ceil {b['e2e']['ceiling']['value']:.3e} frac {b['e2e']['frac_of_ceiling']:.3f}")
    g=b.get('gather') or {}
    print(" nccl", g.get('nccl',{}).get('value'), g.get('nccl',{}).get('gather_gbs_per_rank'), "peer", g.get('peer'))
    for row in b.get('configs',[]):
        print(" ", row.get('config'), row.get('key'), row.get('error') or f"{row['value']:.3e} {row['roofline']['frac']:.3f}")
PY
