# r02-i (1 GPU): C4 curve parity + bench, ncu evidence: C4 k_trace (full set + source), C1 k_trace with the launch query count -> traffic JSON, C1 streaming kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_scenes.py -m gpu -q -k "hair or curve or c4" 2>&1 | tail -4
python bench.py --config c4 --steps 3 --warmup 2 --spp-per-step 8 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('c4: value %.1f e2e %.1f trace %.0f ms (%.0f Mq/s) shadow %.0f ms dev %.0f ms' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['mqueries_per_s'], r['k_shadow']['kernel_ms'], d['device_ms']))"
# r02-i (1 GPU): ncu evidence.  (a) C4 curve traversal, full set + source; (b) C1 k_trace with the launch's query count -> traffic JSON;
# (c) C1 streaming kernels (k_regen, k_shade, k_shadow_prep, k_shadow_bvh, k_accum), one mid-step launch each
B="python bench.py --steps 1 --warmup 1 --spp-per-step 8 --no-cpu-baseline --no-other-configs"
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 1 -f -o /tmp/c4 $B --config c4 > gpurun_out/ncu_c4.log 2>&1
ncu -i /tmp/c4.ncu-rep --page raw --csv > gpurun_out/r02i_c4_k_trace.raw.csv 2>/dev/null
ncu -i /tmp/c4.ncu-rep --page source --csv > gpurun_out/r02i_c4_k_trace.source.csv 2>/dev/null
TGB_TRACE_BOUNCES=1 ncu --set full --clock-control none -k regex:k_trace -s 6 -c 1 -f -o /tmp/c1 $B > gpurun_out/ncu_c1.log 2> gpurun_out/r02i_c1_k_trace.bounces.log
ncu -i /tmp/c1.ncu-rep --page raw --csv > gpurun_out/r02i_c1_k_trace.raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k 'regex:k_regen|k_shade|k_shadow_prep|k_shadow_bvh|k_accum' -s 30 -c 5 -f -o /tmp/st $B > gpurun_out/ncu_st.log 2>&1
ncu -i /tmp/st.ncu-rep --page raw --csv > gpurun_out/r02i_c1_streaming.raw.csv 2>/dev/null
ncu -i /tmp/st.ncu-rep --page source --csv > gpurun_out/r02i_c1_streaming.source.csv 2>/dev/null
ls -la gpurun_out | tail -12; tail -3 gpurun_out/ncu_c4.log gpurun_out/ncu_st.log
