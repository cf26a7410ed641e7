set +e
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -rs) > gpurun_out/r2_gputests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
for w in spline2d spline2d_greeks slider10d full_bs5d; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_ab_$w.json 2> gpurun_out/r2_ab_$w.err
done
PCB_NO_DMMA2=1 python bench.py --workload full_bs5d --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_ab_full_bs5d_old.json 2>> gpurun_out/r2_ab_full_bs5d.err
python - <<'PY'
import json
for w in ("spline2d","spline2d_greeks","slider10d","full_bs5d","full_bs5d_old"):
    try:
        d=json.loads(open(f"gpurun_out/r2_ab_{w}.json").read().strip().splitlines()[-1])
        print(w,f"{d['value']:.3e}", round(d['roofline']['frac'],4), d['roofline']['kernel'])
    except Exception as e: print(w,'ERR',e)
PY
grep -n "FAILED\|passed\|failed" gpurun_out/r2_gputests.log | tail -8
