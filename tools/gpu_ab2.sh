bash tools/gpu_ab.sh base shade10 shade12 2>&1 | grep rep1
TGB_PERSIST_MULT=2 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('mult2: value %.1f trace %.0f ms shadow %.0f ms dev %.0f ms' % (d['value'], r['kernel_ms'], r['k_shadow']['kernel_ms'], d['device_ms']))"
