set +e
mkdir -p gpurun_out
rm -f gpurun_out/r2_ref_suite_all.log
(time PCB_FULL_REF_SUITE=all PCB_REF_SUITE_LOG=$PWD/gpurun_out/r2_ref_suite_all.log timeout 2000 python -m pytest tests/test_reference_suite.py -m gpu -q) > gpurun_out/r2_ref_suite_all.out 2>&1
tail -15 gpurun_out/r2_ref_suite_all.out | cut -c1-300
grep "^==\|passed\|failed" gpurun_out/r2_ref_suite_all.log | cut -c1-160
