set +e
mkdir -p gpurun_out
cap() {  # cap <name> <regex> <target args...>
  name=$1; shift; pat=$1; shift
  python tools/ncu_target.py "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 1 -f -o /tmp/r2_$name python tools/ncu_target.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
  python tools/ncu_summary.py kernel /tmp/r2_$name.ncu-rep gpurun_out/r2_$name.md > gpurun_out/sum_$name.log 2>&1
}
cap spline2d_dmma spline2d_dmma spline_value 20000000
cap ttc_fd2 ttc_fd_shared tt_fd2 8000000
