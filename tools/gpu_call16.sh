set +e
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(time $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3) > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
echo "bench2 rc=$?"
(time $TR --nproc-per-node 2 --master-port 29542 bench.py --impl reference --gpus 2 --steps 3 --warmup 1) > gpurun_out/r2_bench_ref_2gpu.json 2> gpurun_out/r2_bench_ref_2gpu.err
echo "ref2 rc=$?"
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r2_bench_2gpu.json").read().strip().splitlines()[-1])
print("N2", f"value {b['value']:.3e} e2e {b['e2e']['value']:.3e} frac {b['e2e']['frac_of_ceiling']:.3f}")
for row in b.get('configs',[]):
    print(row.get('config'), row.get('key'), row.get('error') or f"{row['value']:.3e} {row['roofline']['frac']:.3f}")
r=open("gpurun_out/r2_bench_ref_2gpu.json").read().strip().splitlines()
print("ref lines:", len(r), r[-1][:200] if r else None)
PY
timeout 300 compute-sanitizer --tool memcheck python tools/ncu_target.py spline_value 200000 > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "sanitizer rc=$?"; tail -5 gpurun_out/r2_sanitizer_memcheck.log
