set +e
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -rs --durations=25) > gpurun_out/r2_gputests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
TT_N=32000000 TT_CONFIGS=0x0,2x512,2x256,2x128 python tools/tt_sweep.py > gpurun_out/r2_tt_sweep2.log 2>&1
TT_CASE=tt_basket10d TT_N=16000000 TT_CONFIGS=0x0,2x512,2x256,2x128 python tools/tt_sweep.py >> gpurun_out/r2_tt_sweep2.log 2>&1
tail -8 gpurun_out/r2_gputests.log; cat gpurun_out/r2_tt_sweep2.log
