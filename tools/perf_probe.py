"""Quick device-resident throughput probe (not the bench): python tools/perf_probe.py [subdiv] [res] [spp]"""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from tungsten_b200 import scene, synth, lib

subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 7
res = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 8
d = "/tmp/tgb_probe"
t = time.time()
p = synth.cornell_mesh(d, "probe%d" % subdiv, subdiv=subdiv, res=(res, res), spp=spp)
fs = scene.load_scene(p)
print("scene gen+load %.2fs, %d tris" % (time.time() - t, fs.n_triangles))
t = time.time(); ctx = lib.Context(fs); print("create %.2fs" % (time.time() - t), ctx.scene_info())
ctx.render_resident(1)  # warm-up
for prof in (False, True):
    ctx.clear(); ctx.reset_stats(); ctx.set_profiling(prof)
    t = time.time(); ctx.render_resident(spp); wall = time.time() - t
    st = ctx.stats()
    print("profiling=%s: %.1f ms device, wall %.1f ms, %.1f Msamples/s, %.1f Mrays/s, %.1f Mhits/s, launches %d, trace %.1f ms over %d launches" % (
        prof, st.total_ms, wall*1e3, st.samples/st.total_ms/1e3, st.rays/st.total_ms/1e3, st.hits/st.total_ms/1e3,
        st.kernel_launches, st.trace_ms, st.trace_launches))
img, cnt = ctx.read_framebuffer()
print("mean radiance", img.mean(), "count", cnt.min(), cnt.max())
