set +e
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for q in 25000000 100000000; do
  COPY_QUICK=1 COPY_QUERIES=$q timeout 200 $TR --nproc-per-node 8 --master-port 29530 tools/copy_ceiling.py > gpurun_out/r2_copyq_$q.log 2>&1
  tail -1 gpurun_out/r2_copyq_$q.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('probe', d['queries_per_rank'], {k:round(v,1) for k,v in d['total'].items()}, [round(x,1) for x in list(d['per_rank'].values())[0]])"
done
E2E_QUERIES=100000000 timeout 300 $TR --nproc-per-node 8 --master-port 29531 tools/e2e_sweep.py > gpurun_out/r2_e2e_sweep_100M.log 2>&1
grep "chunk\|ceiling" gpurun_out/r2_e2e_sweep_100M.log | cut -c1-200
mv gpurun_out/e2e_sweep_N8.json gpurun_out/e2e_sweep_N8_100M.json
E2E_QUERIES=30000000 timeout 300 $TR --nproc-per-node 8 --master-port 29532 tools/e2e_sweep.py > gpurun_out/r2_e2e_sweep_30M.log 2>&1
grep "chunk\|ceiling" gpurun_out/r2_e2e_sweep_30M.log | cut -c1-200
mv gpurun_out/e2e_sweep_N8.json gpurun_out/e2e_sweep_N8_30M.json
