# Full validation of a build on one B200: GPU test suite, smoke, both bench arms, one ncu capture,
# the launch list of the bench command.   gpurun --timeout 1800 -- 'bash tools/gpu/validate.sh'
set +e
mkdir -p gpurun_out
(time python -m pytest tests -q -m gpu -p no:cacheprovider) > gpurun_out/r2_gputests.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_gputests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
(time python bench.py --steps 20 --warmup 5) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench rc=$?"
(time python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
echo "ref rc=$?"
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r2_bench_default.json").read().strip().splitlines()[-1])
print(f"value {b['value']:.4e} frac {b['roofline']['frac']:.3f} e2e {b['e2e']['value']:.4e} ceil {b['e2e']['ceiling']['value']:.4e} frac {b['e2e']['frac_of_ceiling']:.3f} launches {b['gpu_launches']}")
print("cpu", b['cpu_baseline']['value'], b['cpu_baseline']['kind'])
for row in b.get('configs',[]):
    print(row.get('config'), row.get('key'), row.get('error') or f"{row['value']:.3e} {row['roofline']['frac']:.3f}")
r=json.loads(open("gpurun_out/r2_bench_reference.json").read().strip().splitlines()[-1]); print("ref", r['value'], r['cpu_baseline']['cores'])
PY
cap() {  # cap <name> <regex> <target args...>
  name=$1; shift; pat=$1; shift
  python tools/ncu_target.py "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 1 -f -o /tmp/r2_$name python tools/ncu_target.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
  python tools/ncu_summary.py kernel /tmp/r2_$name.ncu-rep gpurun_out/r2_$name.md > gpurun_out/sum_$name.log 2>&1
}
cap spline3d_dmma spline3d_dmma spline3d_value 8000000
python bench.py --steps 2 --warmup 3 --no-configs --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/ncu_summary.py launches gpurun_out/r2_bench_launches.csv gpurun_out/r2_bench_launches.md > gpurun_out/sum_launches.log 2>&1; tail -2 gpurun_out/sum_launches.log
