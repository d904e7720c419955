#!/usr/bin/env python
"""In-library multi-GPU path (tgb_settings::devices): ONE process, the scene replicated on N GPUs, tiles dealt in Morton order,
shares gathered on devices[0] by peer copies.  This is what the C++ drop-in adapter uses; bench.py (one process per GPU +
NCCL all-gather, the driver's contract) measures the same kernels.  Prints one JSON line.

    python tools/bench_group.py --gpus N [--config c1] [--steps 4] [--spp-per-step 64]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--config", default="c1")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--spp-per-step", type=int, default=64)
    a = ap.parse_args()
    import numpy as np
    from tungsten_b200 import scene, lib
    W, H = bench.CONFIGS[a.config]["res"]
    bench.W, bench.H = W, H
    fs = scene.load_scene(bench.make_scene(1024, a.config))
    out = {}
    for n in sorted({1, a.gpus}):
        t0 = time.perf_counter()
        ctx = lib.Context(fs, devices=list(range(n))) if n > 1 else lib.Context(fs, device=0)
        create_s = time.perf_counter() - t0
        spp = a.spp_per_step*n                       # weak scaling: fixed work per GPU
        ctx.clear()
        for i in range(a.warmup):
            ctx.render_resident(spp, spp_begin=i*spp)
        ctx.clear(); ctx.reset_stats(); ctx.set_profiling(True)
        t0 = time.perf_counter()
        for i in range(a.steps):
            ctx.render_resident(spp, spp_begin=i*spp)
        secs = time.perf_counter() - t0
        st = ctx.stats()
        img, cnt = ctx.read_framebuffer()
        out[n] = {"msamples_per_s": st.samples/secs/1e6, "ms_per_step": 1e3*secs/a.steps, "device_ms_slowest_gpu": st.total_ms,
                  "gather_and_sort_ms": st.sort_ms, "create_s": create_s, "checksum": float(np.float64(img).sum()), "count": int(cnt.min())}
        ctx.close()
    n = a.gpus
    line = {"what": "in-library multi-GPU (single process, tgb_settings::devices), " + bench.CONFIGS[a.config]["label"],
            "n_gpus": n, "steps": a.steps, "spp_per_step_per_gpu": a.spp_per_step, "results": out,
            "efficiency_vs_1gpu": out[n]["msamples_per_s"]/(n*out[1]["msamples_per_s"]) if n > 1 else 1.0}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
