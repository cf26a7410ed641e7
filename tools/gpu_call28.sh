set +e
for w in spline3d spline3d_greeks; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-configs 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('aligned $w', f\"{b['value']:.4e}\", f\"{b['roofline']['frac']:.3f}\")"
  PCB_BL3_UNALIGNED=1 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-configs 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('unaligned $w', f\"{b['value']:.4e}\", f\"{b['roofline']['frac']:.3f}\")"
done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py tests/test_gpu_fuzz.py tests/test_gpu_guard_bands.py -q -m gpu -p no:cacheprovider -k "spline or 3d" 2>&1 | tail -2
