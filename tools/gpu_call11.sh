set +e
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py tests/test_reference_suite.py -m gpu -q -k "spline or dmma or special_points or test_spline" ) > gpurun_out/r2_dmma3d_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_dmma3d_tests.log
grep -n "FAILED\|passed\|failed\|Error" gpurun_out/r2_dmma3d_tests.log | tail -8
for w in spline3d spline3d_greeks; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_ab_$w.json 2> gpurun_out/r2_ab_$w.err
  PCB_NO_DMMA3D=1 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_ab_${w}_bank.json 2>> gpurun_out/r2_ab_$w.err
done
python - <<'PY'
import json
for w in ("spline3d","spline3d_bank","spline3d_greeks","spline3d_greeks_bank"):
    try:
        d=json.loads(open(f"gpurun_out/r2_ab_{w}.json").read().strip().splitlines()[-1])
        print(w,f"{d['value']:.3e}", round(d['roofline']['frac'],4), d['roofline']['kernel'])
    except Exception as e: print(w,'ERR',e)
PY
tail -3 gpurun_out/r2_ab_spline3d.err
cap() {  # cap <name> <regex> <target args...>
  name=$1; shift; pat=$1; shift
  python tools/ncu_target.py "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 1 -f -o /tmp/r2_$name python tools/ncu_target.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
  python tools/ncu_summary.py kernel /tmp/r2_$name.ncu-rep gpurun_out/r2_$name.md > /dev/null 2>&1
}
cap full_dmma2 full_dmma2 full_dmma2
