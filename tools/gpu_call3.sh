set +e
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -rs --durations=15) > gpurun_out/r2_gputests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
rm -f gpurun_out/r2_ref_suite_hot.log
(time PCB_FULL_REF_SUITE=hot PCB_REF_SUITE_LOG=$PWD/gpurun_out/r2_ref_suite_hot.log timeout 1200 python -m pytest tests/test_reference_suite.py -m gpu -q) > gpurun_out/r2_ref_suite_hot.out 2>&1
python tools/parity_report.py > gpurun_out/r2_parity.log 2>&1
TT_N=32000000 TT_CONFIGS=0x0,2x512,2x256,1x512 python tools/tt_sweep.py > gpurun_out/r2_tt_sweep3.log 2>&1
(time python bench.py --steps 20 --warmup 5) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -6 gpurun_out/r2_gputests.log; tail -4 gpurun_out/r2_ref_suite_hot.out; tail -3 gpurun_out/r2_parity.log; cat gpurun_out/r2_tt_sweep3.log
