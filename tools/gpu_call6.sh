set +e
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -rs) > gpurun_out/r2_gputests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
python bench.py --workload spline2d_greeks --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_ab_spline2d_greeks.json 2> gpurun_out/r2_ab_spline2d_greeks.err
tail -5 gpurun_out/r2_ab_spline2d_greeks.err
for w in full_bs5d; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_ab_$w.json 2> gpurun_out/r2_ab_$w.err
  PCB_NO_DMMA2=1 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_ab_${w}_old.json 2>> gpurun_out/r2_ab_$w.err
done
python - <<'PY'
import json
for w in ("spline2d_greeks","full_bs5d","full_bs5d_old"):
    try:
        d=json.loads(open(f"gpurun_out/r2_ab_{w}.json").read().strip().splitlines()[-1])
        print(w,f"{d['value']:.3e}", round(d['roofline']['frac'],4))
    except Exception as e: print(w,'ERR',e)
PY
cap() {  # cap <name> <regex> <target args...>
  name=$1; shift; pat=$1; shift
  python tools/ncu_target.py "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 1 -f -o /tmp/r2_$name python tools/ncu_target.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
  python tools/ncu_summary.py kernel /tmp/r2_$name.ncu-rep gpurun_out/r2_$name.md > /dev/null 2>&1
}
cap spline2d_dmma spline2d_dmma spline_value 20000000
cap slider2d_dmma slider2d_dmma slider 10000000
grep -n "FAILED\|passed\|failed" gpurun_out/r2_gputests.log | tail -8
