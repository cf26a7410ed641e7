set +e
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py -m gpu -q -k "spline or slider or dmma") > gpurun_out/r2_tiles4_tests.log 2>&1
grep -n "FAILED\|passed\|failed\|Error" gpurun_out/r2_tiles4_tests.log | tail -6
cd oracle/_ref/ref_tests && PYTHONPATH=$GRAFT_REPO_ROOT/tests/plugins:$GRAFT_REPO_ROOT PYTHONDONTWRITEBYTECODE=1 timeout 600 python -m pytest -q -p no:cacheprovider -p ref_suite_plugin --rootdir . test_from_values.py > $GRAFT_REPO_ROOT/gpurun_out/r2_ref_from_values.log 2>&1; cd $GRAFT_REPO_ROOT
tail -8 gpurun_out/r2_ref_from_values.log | cut -c1-200
for w in slider10d spline3d spline3d_greeks spline2d; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_ab_$w.json 2> gpurun_out/r2_ab_$w.err
done
python - <<'PY'
import json
for w in ("slider10d","spline3d","spline3d_greeks","spline2d"):
    try:
        d=json.loads(open(f"gpurun_out/r2_ab_{w}.json").read().strip().splitlines()[-1])
        print(w,f"{d['value']:.3e}", round(d['roofline']['frac'],4), d['roofline']['kernel'])
    except Exception as e: print(w,'ERR',e)
PY
