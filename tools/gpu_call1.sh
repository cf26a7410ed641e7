set +e
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_box.txt 2>&1
nproc >> gpurun_out/r2_box.txt; free -g | head -2 >> gpurun_out/r2_box.txt
(time timeout 900 python -m pytest tests -m gpu -x -q -rs --durations=15) > gpurun_out/r2_gputests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
python tools/copy_ceiling.py > gpurun_out/r2_copy_N1.log 2>&1
(time python bench.py --steps 10 --warmup 3) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench rc=$?" >> gpurun_out/r2_bench_default.err
(time python bench.py --impl reference --steps 5 --warmup 2) > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
TT_N=32000000 python tools/tt_sweep.py > gpurun_out/r2_tt_sweep.log 2>&1
tail -5 gpurun_out/r2_gputests.log; tail -3 gpurun_out/r2_tt_sweep.log
