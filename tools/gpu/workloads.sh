# Device-resident rate and roofline fraction of single bench workloads (A/B runs of one kernel):
#   gpurun -- 'bash tools/gpu/workloads.sh spline3d spline3d_greeks'      (env vars pass through)
set +e
for w in "$@"; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-configs 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', f\"{b['value']:.4e}\", f\"{b['roofline']['frac']:.3f}\", b['roofline']['kernel'])"
done
