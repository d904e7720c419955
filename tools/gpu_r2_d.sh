# r02-d: traversal ALU fixes (shared-space stack, PRMT immediates, no empty-slot test), prep un-fused
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_scenes.py -m gpu -x -q 2>&1 | tail -6
for v in base f32nodes; do for t in 0 320; do
  if [ "$v" = base ]; then unset TGB200_LIB; else export TGB200_LIB=$PWD/tungsten_b200/libtgb200_$v.so; fi
  TGB_TREELET=$t python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$v treelet $t: value %.1f e2e %.1f trace %.0f ms (%.0f Mq/s) shadow %.0f ms dev %.0f ms' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['mqueries_per_s'], r['k_shadow']['kernel_ms'], d['device_ms']))"
done; done
unset TGB200_LIB
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02d_launches_q4.csv python bench.py --steps 1 --warmup 1 --spp-per-step 8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
TGB_TREELET=0 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 1 -f -o /tmp/kt python bench.py --steps 1 --warmup 1 --spp-per-step 8 --no-cpu-baseline > gpurun_out/ncu_kt.log 2>&1
ncu -i /tmp/kt.ncu-rep --page raw --csv > gpurun_out/r02d_k_trace_q4_t0.raw.csv 2>/dev/null
ncu -i /tmp/kt.ncu-rep --page source --csv > gpurun_out/r02d_k_trace_q4_t0.source.csv 2>/dev/null
