# r02-o (1 GPU): full GPU test suite + smoke + the default bench line (both arms) + launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02o_bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02o_bench_reference.json 2> gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.err
python -c "
import json
d=json.loads(open('gpurun_out/r02o_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value %.1f e2e %.1f ms/step %.1f launches %s clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['clocks']))
print('roofline', {k: r[k] for k in r if k not in ('k_shadow',)})
print('cpu', d.get('cpu_baseline'))
print('others', {k: (round(v['value'],1), round(v['e2e'],1)) for k, v in (d.get('other_configs') or {}).items()})
print('streaming', {k: round(v['frac'],3) for k, v in d['roofline_streaming'].items() if isinstance(v, dict)})
r=json.loads(open('gpurun_out/r02o_bench_reference.json').read().strip().splitlines()[-1]); print('reference arm', r.get('value'), r.get('cpu_baseline'))
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02o_launches.csv python bench.py --steps 1 --warmup 1 --spp-per-step 8 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_bench.log 2>&1
