# r02-g: f3/f4 additions on the GPU (Dirac lobes, cap + skydome, resume state, adaptive), cut-box sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12
for cb in 8 12 16; do TGB_CUT_BOXES=$cb python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; s=d['roofline_streaming']
print('cut $cb: value %.1f trace %.0f ms (%.0f Mq/s, %d q) shadow %.0f ms shade %.0f accum %.0f prep %.0f regen %.0f' % (d['value'], r['kernel_ms'], r['mqueries_per_s'], r['queries'], r['k_shadow']['kernel_ms'], s['k_shade']['kernel_ms'], s['k_accum']['kernel_ms'], d['loop']['k_shadow_prep_ms'], s['k_regen']['kernel_ms']))"; done
