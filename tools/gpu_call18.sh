set +e
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_peer_gather.py tests/test_gpu_fuzz.py -q -m gpu -p no:cacheprovider) > gpurun_out/r2_fuzz.log 2>&1
echo "rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_fuzz.log | head -40
