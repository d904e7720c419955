for cfg in "4 1" "2 4" "1 8"; do set -- $cfg
  TGB_CURVE_LEAF=$1 TGB_CURVE_COST=$2 python bench.py --config c4 --steps 2 --warmup 3 --spp-per-step 8 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']
print('leaf $1 cost $2: C4 value %.1f Msamples/s rays %.0f M/s trace %.0f ms shadow %.0f ms dev %.0f ms nodes %d' % (d['value'], d['mrays_per_s'], r['kernel_ms'], r['k_shadow']['kernel_ms'], d['device_ms'], d['config']['bvh_nodes']))"
done
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "curve or hair" 2>&1 | tail -1
