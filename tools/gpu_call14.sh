set +e
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py -m gpu -q -k "spline or slider or dmma") > gpurun_out/r2_tiles_tests.log 2>&1
grep -n "FAILED\|passed\|failed\|Error" gpurun_out/r2_tiles_tests.log | tail -6
for w in spline2d spline2d_greeks spline3d slider10d; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_ab_$w.json 2> gpurun_out/r2_ab_$w.err
done
python - <<'PY'
import json
for w in ("spline2d","spline2d_greeks","spline3d","slider10d"):
    try:
        d=json.loads(open(f"gpurun_out/r2_ab_{w}.json").read().strip().splitlines()[-1])
        print(w,f"{d['value']:.3e}", round(d['roofline']['frac'],4), d['roofline']['kernel'])
    except Exception as e: print(w,'ERR',e)
PY
