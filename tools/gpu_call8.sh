set +e
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo8.txt 2>&1
lscpu | head -30 >> gpurun_out/r2_topo8.txt; nproc >> gpurun_out/r2_topo8.txt; free -g | head -2 >> gpurun_out/r2_topo8.txt
for d in /sys/bus/pci/devices/*; do if grep -qi 0x10de $d/vendor 2>/dev/null && grep -q 0x0302 $d/class 2>/dev/null; then echo $(basename $d) numa=$(cat $d/numa_node); fi; done >> gpurun_out/r2_topo8.txt
cat /sys/devices/system/node/node*/cpulist >> gpurun_out/r2_topo8.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port $((29500+n)) tools/copy_ceiling.py > gpurun_out/r2_copy_N$n.log 2>&1
done
$TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
echo "bench8 rc=$?"
$TR --nproc-per-node 2 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 --no-configs > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
echo "bench2 rc=$?"
$TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 --no-configs > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err
echo "bench4 rc=$?"
python - <<'PY'
import json
for n in (2,4,8):
    try:
        d=json.load(open(f"gpurun_out/copy_ceiling_N{n}.json"))
        print("copy N",n,"total both GB/s:", {k:round(v,1) for k,v in d['total'].items() if k.endswith('both')}, "h2d/d2h:", {k:round(v,1) for k,v in d['total'].items() if not k.endswith('both')}, d['gpu_numa_node_per_rank'])
    except Exception as e: print(n,'ERR',e)
    try:
        b=json.loads(open(f"gpurun_out/r2_bench_{n}gpu.json").read().strip().splitlines()[-1])
        print("bench N",n,f"value {b['value']:.3e} e2e {b['e2e']['value']:.3e} ceiling {b['e2e']['ceiling']['value']:.3e} frac {b['e2e']['frac_of_ceiling']:.3f}")
    except Exception as e: print(n,'bench ERR',e)
PY
tail -3 gpurun_out/r2_bench_8gpu.err
