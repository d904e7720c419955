# r02-q (1 GPU): all-Lambert instantiation of k_shade: parity + C1 A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_scenes.py -m gpu -q --tb=short -k "cornell or golden or c1 or many_lights or analytic or adaptive or resume" 2>&1 | grep -v "^$" | tail -6
run() { python bench.py --config $2 --steps $3 --warmup 3 --spp-per-step $4 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; s=d['roofline_streaming']
print('$1 $2: value %.1f e2e %.1f trace %.0f ms shadow %.0f ms shade %.0f ms accum %.0f ms dev %.0f ms' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['k_shadow']['kernel_ms'], s['k_shade']['kernel_ms'], s['k_accum']['kernel_ms'], d['device_ms']))"; }
run diffuse c1 4 64
TGB_LOBE_SET=0 run generic c1 4 64
run diffuse c1 4 64
