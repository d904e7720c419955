# r02-e: SoA float4 state records, leaf-order shading records, drop-in adaptive/resume tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 8 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/bench.err; cat gpurun_out/r02e_bench.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('value %.1f e2e %.1f trace %.0f ms (%.0f Mq/s) shadow %.0f ms dev %.0f ms cpu %s' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['mqueries_per_s'], r['k_shadow']['kernel_ms'], d['device_ms'], d['cpu_baseline']['value'] if d['cpu_baseline'] else None))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 1 --warmup 1 --spp-per-step 8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
for c in c2 c3 c4; do python bench.py --config $c --steps 4 --warmup 3 --spp-per-step 8 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$c value %.1f e2e %.1f trace %.0f Mq/s' % (d['value'], d['e2e']['value'], r['mqueries_per_s']))"; done
