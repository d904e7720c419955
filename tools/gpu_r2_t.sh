# r02-t (1 GPU): window of the material-sorted k_shade (slots sorted together = block x items per thread) on C2 / C3
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_scenes.py -m gpu -q --tb=short -k "material or c2 or c3 or golden or coat or dirac or forest" 2>&1 | grep -v "^$" | tail -4
run() { python bench.py --config $2 --steps $3 --warmup 3 --spp-per-step $4 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; s=d['roofline_streaming']
print('$1 $2: value %.1f e2e %.1f trace %.0f ms shadow %.0f ms shade %.0f ms accum %.0f ms dev %.0f ms' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['k_shadow']['kernel_ms'], s['k_shade']['kernel_ms'], s['k_accum']['kernel_ms'], d['device_ms']))"; }
for v in base w1024 w4096 w512x4 w512x8; do
  if [ "$v" = base ]; then unset TGB200_LIB; else export TGB200_LIB=$PWD/tungsten_b200/libtgb200_$v.so; fi
  run $v c2 3 8; run $v c3 3 8
done
