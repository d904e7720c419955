# r02-n (1 GPU): analytic primitives as BVH leaves (tests), per-refill atomics restored, triangle prefetch variants on C1/C3
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | grep -v "^$" | tail -12
run() { python bench.py --config $2 --steps $3 --warmup 3 --spp-per-step $4 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$1 $2: value %.1f e2e %.1f trace %.0f ms (%.0f Mq/s) shadow %.0f ms (%.0f Mq/s) dev %.0f ms' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['mqueries_per_s'], r['k_shadow']['kernel_ms'], r['k_shadow'].get('mqueries_per_s', 0), d['device_ms']))"; }
for v in base pf pf_r12 pf_minb5 base; do
  if [ "$v" = base ]; then unset TGB200_LIB; else export TGB200_LIB=$PWD/tungsten_b200/libtgb200_$v.so; fi
  run $v c1 4 64
done
for v in base pf; do
  if [ "$v" = base ]; then unset TGB200_LIB; else export TGB200_LIB=$PWD/tungsten_b200/libtgb200_$v.so; fi
  run $v c3 3 8; run $v c2 3 8
done
unset TGB200_LIB
TGB_L2_PERSIST=1 run l2persist c1 4 64
