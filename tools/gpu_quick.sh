mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/perf_probe.py 7 1024 8 2>&1 | tail -3
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); r=d['roofline']
print('value %.1f Msamples/s  e2e %.1f  rays %.0f M/s hits %.0f M/s  launches %d  trace %.0f ms (%.0f Mq/s)  shadow %.0f ms  dev %.0f ms' % (d['value'], d['e2e']['value'], d['mrays_per_s'], d['mray_hits_per_s'], d['gpu_launches'], r['kernel_ms'], r['mqueries_per_s'], r['k_shadow']['kernel_ms'], d['device_ms']))"
tail -3 gpurun_out/bench_quick.err
