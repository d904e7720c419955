#!/usr/bin/env python
"""Small renders of every kernel family for `compute-sanitizer --tool memcheck python tools/sanitize_small.py`:
triangles + analytic loop, analytic primitives as BVH leaves, curves (state machine with BISECT/CYL), material sort, adaptive."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tungsten_b200 import scene, synth, lib

d = tempfile.mkdtemp(prefix="tgb_san_")
cases = [("cornell_mesh", synth.cornell_mesh(d, res=(40, 40), spp=2, subdiv=2)),
         ("cube_city", synth.cube_city(res=(40, 40), spp=2, n=8)),
         ("materials", synth.material_room(d, res=(40, 40), spp=2, subdiv=2)),
         ("hair", synth.hair_scene(d, n_curves=60, res=(40, 40), spp=2)),
         ("many_lights", synth.many_lights(d, res=(40, 40), spp=2, subdiv=2))]
for name, sc in cases:
    fs = scene.load_scene(sc)
    ctx = lib.Context(fs, device=0)
    img, cnt = ctx.render_tiles(2)
    st = ctx.stats()
    print("%-12s mean %.5f rays %d launches %d" % (name, float(img.mean()), st.rays, st.kernel_launches), flush=True)
    ctx.close()
print("done")
