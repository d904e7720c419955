python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
run() { python bench.py --config $2 --steps 4 --warmup 3 $3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1 $2: value %.1f e2e %.1f dev %.0f ms trace %.0f shadow %.0f' % (d['value'], d['e2e']['value'], d['device_ms'], r['kernel_ms'], r['k_shadow']['kernel_ms']))"; }
run sorted c2 "--spp-per-step 8"
TGB_SORT_MATERIALS=0 run unsorted c2 "--spp-per-step 8"
run default c1 ""
