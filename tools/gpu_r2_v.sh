# r02-v (1 GPU): triangle records in flight in the LEAF block (2 vs 4) x resident blocks, C1
mkdir -p gpurun_out
run() { python bench.py --config $2 --steps $3 --warmup 3 --spp-per-step $4 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; s=d['roofline_streaming']
print('$1 $2: value %.1f e2e %.1f trace %.0f ms (%.0f Mq/s) shadow %.0f ms shade %.0f ms accum %.0f ms dev %.0f ms' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['mqueries_per_s'], r['k_shadow']['kernel_ms'], s['k_shade']['kernel_ms'], s['k_accum']['kernel_ms'], d['device_ms']))"; }
for v in base quad6 quad5 pair5 pair7 base; do
  if [ "$v" = base ]; then unset TGB200_LIB; else export TGB200_LIB=$PWD/tungsten_b200/libtgb200_$v.so; fi
  run $v c1 4 64
done
