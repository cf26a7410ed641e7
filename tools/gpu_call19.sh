set +e
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(time timeout 600 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-configs) > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
echo "bench2 rc=$?"; tail -3 gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r2_bench_2gpu.json").read().strip().splitlines()[-1])
print("N2", f"value {b['value']:.3e} e2e {b['e2e']['value']:.3e} ceil {b['e2e']['ceiling']['value']:.3e} frac {b['e2e']['frac_of_ceiling']:.3f}")
print(json.dumps(b.get('gather'), indent=1))
PY
