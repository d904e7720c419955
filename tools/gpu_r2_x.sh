# r02-x (1 GPU): final validation of the round: full GPU suite, smoke, default bench (both arms), launch list, k_trace traffic capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02x_bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02x_bench_reference.json 2> gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.err
python -c "
import json
d=json.loads(open('gpurun_out/r02x_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value %.1f e2e %.1f ms/step %.1f launches %s clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['clocks']))
print('k_trace: %.1f GB/s frac %.4f, %.0f Mq/s, kernel_ms %.0f' % (r['achieved'], r['frac'], r['mqueries_per_s'], r['kernel_ms']))
print('others', {k: (round(v['value'],1), round(v['e2e'],1)) for k, v in (d.get('other_configs') or {}).items()})
print('streaming', {k: (round(v['frac'],3), round(v['kernel_ms'])) for k, v in d['roofline_streaming'].items() if isinstance(v, dict)})
print('loop', d['loop'])
r=json.loads(open('gpurun_out/r02x_bench_reference.json').read().strip().splitlines()[-1]); print('reference arm', r.get('value'))
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02x_launches.csv python bench.py --steps 1 --warmup 1 --spp-per-step 8 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_bench.log 2>&1
B="python bench.py --steps 1 --warmup 1 --spp-per-step 8 --no-cpu-baseline --no-other-configs"
TGB_TRACE_BOUNCES=1 ncu --set full --clock-control none -k regex:k_trace -s 6 -c 1 -f -o /tmp/c1 $B > gpurun_out/ncu_c1.log 2> gpurun_out/r02x_c1_k_trace.bounces.log
ncu -i /tmp/c1.ncu-rep --page raw --csv > gpurun_out/r02x_c1_k_trace.raw.csv 2>/dev/null
ncu --set full --clock-control none -k 'regex:k_shade|k_shadow_bvh' -s 12 -c 2 -f -o /tmp/st $B > gpurun_out/ncu_st.log 2>&1
ncu -i /tmp/st.ncu-rep --page raw --csv > gpurun_out/r02x_c1_shade_shadow.raw.csv 2>/dev/null
