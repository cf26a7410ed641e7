# bench.py under torchrun at N GPUs (default 8):  gpurun --gpus N -- 'bash tools/gpu/multi_gpu.sh N [extra bench flags]'
set +e
N=${1:-8}; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(time timeout 900 $TR --nproc-per-node $N --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 "$@") \
  > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "bench$N rc=$?"; tail -4 gpurun_out/r2_bench_${N}gpu.err
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
b = json.loads(open(f"gpurun_out/r2_bench_{n}gpu.json").read().strip().splitlines()[-1])
print(n, f"value {b['value']:.3e} e2e {b['e2e']['value']:.3e} ceiling {b['e2e']['ceiling']['value']:.3e} "
         f"frac {b['e2e']['frac_of_ceiling']:.3f}")
g = b.get("gather") or {}
print(" gather: nccl", g.get("nccl", {}).get("value"), "peer", g.get("peer", {}).get("value"),
      g.get("peer", {}).get("of_device_resident_value"))
for row in b.get("configs", []):
    print(" ", row.get("config"), row.get("key"),
          row.get("error") or f"{row['value']:.3e} {row['roofline']['frac']:.3f}")
PY
