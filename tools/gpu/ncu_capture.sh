# One `ncu --set full` capture of one kernel, summarised to gpurun_out/r2_<name>.md (the .ncu-rep
# stays in /tmp on the box: four of them exceed gpurun_out's 64 MiB).  The target runs plain first.
#   gpurun -- 'bash tools/gpu/ncu_capture.sh spline3d_dmma spline3d_dmma spline3d_value 8000000'
#   args: <name> <kernel regex> <tools/ncu_target.py arguments...>
set +e
mkdir -p gpurun_out
name=$1; shift; pat=$1; shift
python tools/ncu_target.py "$@" > gpurun_out/plain_$name.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 1 -f -o /tmp/r2_$name \
    python tools/ncu_target.py "$@" > gpurun_out/ncu_$name.log 2>&1
echo "$name rc=$?"
python tools/ncu_summary.py kernel /tmp/r2_$name.ncu-rep gpurun_out/r2_$name.md > gpurun_out/sum_$name.log 2>&1
