set +e
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_fuzz.py -q -m gpu -p no:cacheprovider) > gpurun_out/r2_fuzz.log 2>&1
echo "fuzz rc=$?"; tail -40 gpurun_out/r2_fuzz.log | grep -E "passed|failed|FAILED|Error|error" | head -40
(time python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_fuzz.py) > gpurun_out/r2_gputests.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2_gputests.log
(time python bench.py --steps 20 --warmup 5) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench rc=$?"
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r2_bench_default.json").read().strip().splitlines()[-1])
print(f"value {b['value']:.4e} frac {b['roofline']['frac']:.3f} e2e {b['e2e']['value']:.4e} ceil {b['e2e']['ceiling']['value']:.4e} frac {b['e2e']['frac_of_ceiling']:.3f}")
print("cpu", b['cpu_baseline'])
for row in b.get('configs',[]):
    print(row.get('config'), row.get('key'), row.get('error') or f"{row['value']:.3e} {row['roofline']['frac']:.3f} {row.get('kernel')}")
PY
