set +e
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_guard_bands.py tests/test_gpu_fuzz.py -q -m gpu -p no:cacheprovider) > gpurun_out/r2_guard.log 2>&1
echo "rc=$?"; grep -E "passed|failed|FAILED|Error|assert" gpurun_out/r2_guard.log | head -40
