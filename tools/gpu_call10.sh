set +e
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -rs) > gpurun_out/r2_gputests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
grep -n "FAILED\|passed\|failed" gpurun_out/r2_gputests.log | tail -5
(time python bench.py --steps 20 --warmup 5) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench rc=$?"
(time python bench.py --impl reference --steps 20 --warmup 5) > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
cap() {  # cap <name> <regex> <target args...>
  name=$1; shift; pat=$1; shift
  python tools/ncu_target.py "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 1 -f -o /tmp/r2_$name python tools/ncu_target.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
  python tools/ncu_summary.py kernel /tmp/r2_$name.ncu-rep gpurun_out/r2_$name.md > /dev/null 2>&1
}
cap full_dmma2 full_dmma2 full_dmma
cap ttc_fd2 ttc_fd_shared tt_fd2 8000000
cap ttc_value ttc_value tt_value 8000000
python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
E2E_QUERIES=100000000 python tools/e2e_sweep.py > gpurun_out/r2_e2e_sweep_N1_100M.log 2>&1
E2E_QUERIES=30000000 python tools/e2e_sweep.py > gpurun_out/r2_e2e_sweep_N1_30M.log 2>&1
grep "chunk\|ceiling" gpurun_out/r2_e2e_sweep_N1_100M.log gpurun_out/r2_e2e_sweep_N1_30M.log | cut -c1-220
