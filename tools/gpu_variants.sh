TGB_TRACE_BOUNCES=1 python - <<'PY' 2>&1 | tail -60
import sys, time, os
sys.path.insert(0, os.getcwd())
from tungsten_b200 import scene, synth, lib
import bench
p = bench.make_scene(1024)
fs = scene.load_scene(p)
ctx = lib.Context(fs)
ctx.render_resident(2)
ctx.clear(); ctx.reset_stats(); ctx.set_profiling(True)
t=time.time(); ctx.render_resident(2); print("wall ms", (time.time()-t)*1e3, "dev", ctx.stats().total_ms)
PY
