# r02-s (1 GPU): block size of the material-sorted k_shade on C2 / C3
mkdir -p gpurun_out
run() { python bench.py --config $2 --steps $3 --warmup 3 --spp-per-step $4 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; s=d['roofline_streaming']
print('$1 $2: value %.1f e2e %.1f trace %.0f ms shadow %.0f ms shade %.0f ms accum %.0f ms dev %.0f ms' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['k_shadow']['kernel_ms'], s['k_shade']['kernel_ms'], s['k_accum']['kernel_ms'], d['device_ms']))"; }
for v in base sb256 sb128 sb1024; do
  if [ "$v" = base ]; then unset TGB200_LIB; else export TGB200_LIB=$PWD/tungsten_b200/libtgb200_$v.so; fi
  run $v c2 3 8; run $v c3 3 8
done
unset TGB200_LIB
run base c1 4 64
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_scenes.py -m gpu -q --tb=short -k "material or c2 or golden or coat" 2>&1 | grep -v "^$" | tail -4
