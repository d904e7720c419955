# r02-j (1 GPU): curve traversal as a warp state machine: parity, then C4 with tuning variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_scenes.py tests/test_gpu_dropin.py -m gpu -q -k "hair or curve or c4" 2>&1 | tail -4
for v in base cm_refill4 cm_refill16 cm_w1 cm_w2 cm_minb5 cm_minb3; do
  if [ "$v" = base ]; then unset TGB200_LIB; else export TGB200_LIB=$PWD/tungsten_b200/libtgb200_$v.so; fi
  python bench.py --config c4 --steps 3 --warmup 2 --spp-per-step 8 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$v c4: value %.1f e2e %.1f trace %.0f ms (%.0f Mq/s) shadow %.0f ms dev %.0f ms' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['mqueries_per_s'], r['k_shadow']['kernel_ms'], d['device_ms']))"
done
unset TGB200_LIB
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 1 -f -o /tmp/c4 python bench.py --steps 1 --warmup 1 --spp-per-step 8 --no-cpu-baseline --no-other-configs --config c4 > gpurun_out/ncu_c4.log 2>&1
ncu -i /tmp/c4.ncu-rep --page raw --csv > gpurun_out/r02j_c4_k_trace_sm.raw.csv 2>/dev/null
ncu -i /tmp/c4.ncu-rep --page source --csv > gpurun_out/r02j_c4_k_trace_sm.source.csv 2>/dev/null
