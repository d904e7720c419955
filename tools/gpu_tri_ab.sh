for cfg in "4 1" "4 0.5" "4 2" "2 1"; do set -- $cfg
  TGB_TRI_LEAF=$1 TGB_TRI_COST=$2 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']
print('leaf $1 cost $2: value %.1f trace %.0f ms shadow %.0f ms dev %.0f ms nodes %d' % (d['value'], r['kernel_ms'], r['k_shadow']['kernel_ms'], d['device_ms'], d['config']['bvh_nodes']))"
done
