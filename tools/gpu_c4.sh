python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "curve or hair or golden" 2>&1 | grep -E "exact|passed|failed|identical|assert" | tail -14
python bench.py --config c4 --steps 2 --warmup 3 --spp-per-step 8 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']
print('C4 value %.1f Msamples/s e2e %.1f rays %.0f M/s trace %.0f ms (%.0f Mq/s) shadow %.0f ms dev %.0f ms' % (d['value'], d['e2e']['value'], d['mrays_per_s'], r['kernel_ms'], r['mqueries_per_s'], r['k_shadow']['kernel_ms'], d['device_ms']))"
