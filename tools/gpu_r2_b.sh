# r02-b: first run of the rewritten wavefront (16-byte records, device-resident loop control, QNode4 + TMA-staged treelet)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_bench_scenes.py -m gpu -x -q -s 2>&1 | tail -25
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/gpu_ab.sh base f32nodes cta768 2>&1 | tail -8
echo "--- treelet sweep (base = 128x6 CTAs)"
for t in 0 64 320; do TGB_TREELET=$t python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('treelet $t: value %.1f trace %.0f ms (%.0f Mq/s) shadow %.0f ms' % (d['value'], r['kernel_ms'], r['mqueries_per_s'], r['k_shadow']['kernel_ms']))"; done
echo "--- treelet sweep (cta768)"
for t in 0 512 1536 2300; do TGB_TREELET=$t TGB200_LIB=$PWD/tungsten_b200/libtgb200_cta768.so python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('cta768 treelet $t: value %.1f trace %.0f ms (%.0f Mq/s) shadow %.0f ms' % (d['value'], r['kernel_ms'], r['mqueries_per_s'], r['k_shadow']['kernel_ms']))"; done
