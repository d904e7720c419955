# r02-m (1 GPU): machine refill tuning on C1 (ray chunks per atomic, idle threshold) + ncu of the machine k_trace / k_shadow_bvh
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_scenes.py -m gpu -q -x --tb=short 2>&1 | grep -v "^$" | tail -5
run() { python bench.py --config $2 --steps $3 --warmup 3 --spp-per-step $4 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$1 $2: value %.1f e2e %.1f trace %.0f ms (%.0f Mq/s) shadow %.0f ms (%.0f Mq/s) dev %.0f ms' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['mqueries_per_s'], r['k_shadow']['kernel_ms'], r['k_shadow'].get('mqueries_per_s', 0), d['device_ms']))"; }
for v in base c32 c128 r20 r24 r12c128 base; do
  if [ "$v" = base ]; then unset TGB200_LIB; else export TGB200_LIB=$PWD/tungsten_b200/libtgb200_$v.so; fi
  run $v c1 4 64
done
unset TGB200_LIB
run base c4 3 8
B="python bench.py --steps 1 --warmup 1 --spp-per-step 8 --no-cpu-baseline --no-other-configs"
ncu --set full --clock-control none --import-source on -k 'regex:k_trace|k_shadow_bvh' -s 12 -c 2 -f -o /tmp/m $B > gpurun_out/ncu_m.log 2>&1
ncu -i /tmp/m.ncu-rep --page raw --csv > gpurun_out/r02m_c1_machine.raw.csv 2>/dev/null
ncu -i /tmp/m.ncu-rep --page source --csv > gpurun_out/r02m_c1_machine.source.csv 2>/dev/null
