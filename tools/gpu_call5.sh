set +e
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -rs) > gpurun_out/r2_gputests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
for w in spline2d spline2d_greeks slider10d tt_rank20; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_ab_$w.json 2> gpurun_out/r2_ab_$w.err
done
python - <<'PY'
import json
for w in ("spline2d","spline2d_greeks","slider10d","tt_rank20"):
    try:
        d=json.loads(open(f"gpurun_out/r2_ab_{w}.json").read().strip().splitlines()[-1])
        print(w,f"{d['value']:.3e}", round(d['roofline']['frac'],4), d['roofline']['kernel'])
    except Exception as e: print(w,'ERR',e)
PY
cap() {  # cap <name> <regex> <target args...>
  name=$1; shift; pat=$1; shift
  python tools/ncu_target.py "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 2 -f -o gpurun_out/r2_$name python tools/ncu_target.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
}
cap spline2d_dmma spline2d_dmma spline_value 20000000
cap slider2d_dmma slider2d_dmma slider 10000000
cap ttc_fd2 ttc_fd_shared tt_fd2 8000000
cap ttc_value ttc_value tt_value 8000000
python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
tail -4 gpurun_out/r2_gputests.log
