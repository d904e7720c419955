# A/B of an environment switch on ONE box: tools/gpu_env_ab.sh VAR valA valB
mkdir -p gpurun_out
for rep in 1 2; do for v in "$2" "$3"; do
  env $1=$v python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$1=$v rep$rep: value %.1f e2e %.1f trace %.0f ms (%.0f Mq/s) shadow %.0f ms dev %.0f ms' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['mqueries_per_s'], r['k_shadow']['kernel_ms'], d['device_ms']))"
done; done
