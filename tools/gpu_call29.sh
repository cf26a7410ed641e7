set +e
for w in spline2d spline2d_greeks spline3d spline3d_greeks slider10d; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-configs 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', f\"{b['value']:.4e}\", f\"{b['roofline']['frac']:.3f}\")"
done
PCB_NO_DMMA2D=1 PCB_NO_DMMA3D=1 python bench.py --workload spline2d --steps 10 --warmup 3 --no-cpu --no-configs 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('spline2d bank', f\"{b['value']:.4e}\")"
python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py tests/test_gpu_fuzz.py tests/test_gpu_guard_bands.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
