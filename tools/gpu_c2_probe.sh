run() { python bench.py --config c2 --steps 2 --warmup 3 --spp-per-step 4 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1: value %.1f dev %.0f ms trace %.0f shadow %.0f' % (d['value'], d['device_ms'], r['kernel_ms'], r['k_shadow']['kernel_ms']))"; }
run base
TGB200_LIB=$PWD/tungsten_b200/libtgb200_shade1.so run shade1
python - <<'PY'
import json
p='/tmp/tgb200_bench_scene/room.json'
js=json.load(open(p))
for q in js['primitives']:
    if q.get('type')=='infinite_sphere': q['emission']=[0.4,0.5,0.7]
json.dump(js, open(p,'w'))
PY
run const_env
python - <<'PY'
import json
p='/tmp/tgb200_bench_scene/room.json'
js=json.load(open(p))
js['primitives']=[q for q in js['primitives'] if q.get('name')!='lamp' or q.get('type')!='mesh']
json.dump(js, open(p,'w'))
PY
run const_env_no_mesh_light
