set +e
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py -m gpu -q -k "spline or slider or dmma_2d" ) > gpurun_out/r2_dmma2d_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_dmma2d_tests.log
for w in spline2d spline2d_greeks slider10d; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_dmma2d_$w.json 2> gpurun_out/r2_dmma2d_$w.err
  PCB_NO_DMMA2D=1 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_bank_$w.json 2>> gpurun_out/r2_dmma2d_$w.err
done
python tools/parity_report.py > gpurun_out/r2_parity.log 2>&1
tail -5 gpurun_out/r2_dmma2d_tests.log
python - <<'PY'
import json
for w in ("spline2d","spline2d_greeks","slider10d"):
    for k in ("dmma2d","bank"):
        try:
            d=json.loads(open(f"gpurun_out/r2_{k}_{w}.json").read().strip().splitlines()[-1])
            print(w,k,f"{d['value']:.3e}", d['roofline']['frac'], d['e2e'] and f"{d['e2e']['value']:.3e}")
        except Exception as e: print(w,k,'ERR',e)
PY
tail -3 gpurun_out/r2_parity.log
