#!/usr/bin/env python
"""Headline benchmark of the B200 path_tracer hot path (BASELINE.json metric: Msamples/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--spp-per-step S]

Workload (config C1 of BASELINE.json): Cornell box + 868,480-triangle Lambert mesh (procedural stand-in for the
Stanford dragon, see tungsten_b200/synth.py), 1920x1080, path_tracer, max_bounces 64, Sobol sampler,
target 1024 spp.  One STEP = one pass of the hot path over one batch: 64 samples per pixel of the whole frame
(132.7 M camera paths); the default 16 timed steps are the complete 1024-spp frame.  With N GPUs the image's
16x16 tiles are dealt round-robin to the ranks and a step renders 64*N spp (fixed work per GPU: weak scaling),
followed by the path's single collective, an all-gather of the tile-major float3 framebuffer.

`value`   : whole-job Msamples/s with everything resident in HBM (framebuffer stays on the device).
`e2e`     : the same metric through the reference-facing C-ABI call tgb200_render_tiles with HOST buffers
            (running-mean framebuffer + counts copied host->device and back inside every step).
`roofline`: traversal kernel (k_trace): algorithmic bytes = 48 B per closest-hit query (32 B ray read + 16 B hit
            write) / CUDA-event duration of the launches in the timed region / measured HBM peak.
`cpu_baseline`: the reference's own CPU renderer (oracle/_ref/tungsten, built from /root/reference by
            oracle/ref/Makefile) on this box's host cores, same scene, bounded sample (fewer spp).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
ALG_BYTES_PER_QUERY = 48
CONFIGS = {
    "c1": {"res": (1920, 1080), "label": "C1: Cornell box + 868,480-tri Lambert mesh (procedural dragon stand-in), 1920x1080, path_tracer, max_bounces 64, Sobol"},
    "c2": {"res": (1920, 1080), "label": "C2: room with 327,680-triangle furniture stand-ins in rough conductor / rough dielectric / plastic / rough plastic, checker floor, quad + mesh lights + importance-sampled HDR environment, 1920x1080, max_bounces 16, Sobol"},
    "c3": {"res": (3840, 2160), "label": "C3: 12,544,000-triangle instanced forest (490 trees x 2 masters, flattened), Lambert + rough plastic + HDR sky, 3840x2160, max_bounces 16, Sobol"},
    "c4": {"res": (1920, 1080), "label": "C4: 650,000 quadratic B-spline curve segments (10,000 curly strands x 67 nodes, bcsdf_cylinder) with the hair BCSDF over a Lambert floor, quad light + constant sky, 1920x1080, max_bounces 16, Sobol"},
}


def _clock_sampler(stop, out, idx):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(idx), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 6:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.2)


def _summarise_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm)//2] if sm else None, "sm_max_mhz": int(samples[0][1]) if samples[0][1].isdigit() else None,
            "reasons": reasons, "samples": len(samples)}


def _scene_dir():
    d = os.path.join(tempfile.gettempdir(), "tgb200_bench_scene")
    os.makedirs(d, exist_ok=True)
    return d


def make_scene(spp, config="c1"):
    from tungsten_b200 import synth
    d = _scene_dir()
    if config == "c3":
        path = os.path.join(d, "forest10m.json")
        if not os.path.exists(path):
            synth.instanced_forest(d, "forest10m", n_instances=490, tree_subdiv=5, res=CONFIGS["c3"]["res"], spp=spp, extent=40.0)
        return path
    if config == "c2":
        path = os.path.join(d, "room.json")
        if not os.path.exists(path):
            synth.save_rgbe(os.path.join(d, "room_env.hdr"), synth.sky_envmap(512, 256))
            synth.material_room(d, "room", res=CONFIGS["c2"]["res"], spp=spp, max_bounces=16, subdiv=6, env="room_env.hdr")
        return path
    if config == "c4":
        path = os.path.join(d, "hair650k.json")
        if not os.path.exists(path):
            synth.hair_scene(d, "hair650k", n_curves=10000, nodes_per_curve=67, res=CONFIGS["c4"]["res"], spp=spp, width=0.004)
        return path
    path = os.path.join(d, "cornell_dragon.json")
    marker = os.path.join(d, "cornell_dragon_body.wo3")
    if not (os.path.exists(path) and os.path.exists(marker)):
        synth.cornell_dragon_standin(d, res=(W, H), spp=spp)
    return path


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def parse_duration(text):
    """Seconds from the reference's `durationToString` (src/core/io/StringUtils.cpp:51-68):
    "[Dd ][Hh ][Mm ]Ss MSms" when >= 1 s, else the plain double followed by "s"."""
    m = re.search(r"Render time ((?:\d+d )?(?:\d+h )?(?:\d+m )?)(\d+)s (\d+)ms", text)
    if m:
        secs = float(m.group(2)) + float(m.group(3))/1e3
        for tok in m.group(1).split():
            secs += float(tok[:-1])*{"d": 86400.0, "h": 3600.0, "m": 60.0}[tok[-1]]
        return secs
    m = re.search(r"Render time ([0-9.eE+-]+)s", text)
    return float(m.group(1)) if m else None


def host_cpu_info():
    """What the CPU arm can actually use: affinity mask, cgroup quota, model string."""
    info = {"os_cpu_count": os.cpu_count() or 1}
    try:
        info["affinity"] = len(os.sched_getaffinity(0))
    except Exception:
        info["affinity"] = info["os_cpu_count"]
    quota = None
    for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(p).read().split()
            if p.endswith("cpu.max"):
                if txt[0] != "max":
                    quota = float(txt[0])/float(txt[1])
            else:
                q = float(txt[0])
                if q > 0:
                    quota = q/float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            break
        except Exception:
            continue
    info["cgroup_cpus"] = quota
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                info["model"] = ln.split(":", 1)[1].strip(); break
    except Exception:
        pass
    usable = info["affinity"]
    if quota:
        usable = max(1, min(usable, int(quota + 0.5)))
    info["usable"] = usable
    return info


def run_reference_binary(scene_path, spp, threads):
    """Times oracle/_ref/tungsten (the unmodified reference) on `spp` samples per pixel of the bench scene.
    Returns (Msamples/s, render seconds): the reference's own "Render time", which excludes scene load and its
    Embree BVH build (src/tungsten/Shared.hpp:255-319) -- as our arm's timed region excludes tgb200_create."""
    exe = os.path.join(ROOT, "oracle", "_ref", "tungsten")
    if not os.path.exists(exe):
        return None
    js = json.load(open(scene_path))
    js["renderer"].update(spp=spp, spp_step=spp, adaptive_sampling=False, stratified_sampler=True,
                          hdr_output_file="ref.pfm", output_file="ref.png")
    d = os.path.dirname(scene_path)
    rp = os.path.join(d, "ref_run.json")
    json.dump(js, open(rp, "w"))
    out = subprocess.run([exe, "-t", str(threads), "-d", os.path.join(d, "ref_out"), rp], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True)
    secs = parse_duration(out.stdout)
    if not secs:
        return None
    return W*H*spp/secs/1e6, secs


def bench_reference(args, rank, world):
    """--impl reference: the reference's own CPU path on the host cores (rank 0 only).  One step = one run of the
    unmodified binary on a bounded sample (spp chosen from a calibration run so that a step RENDERS for >= ~6 s)."""
    if rank != 0:
        return
    cpu = host_cpu_info()
    threads = cpu["usable"]
    scene_path = make_scene(1024, args.config)
    cal = run_reference_binary(scene_path, 2, threads)          # calibration, untimed
    if cal is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/tungsten is missing (run make -C oracle/ref)"}))
        return
    spp = args.ref_spp if args.ref_spp > 0 else int(min(256, max(4, round(cal[0]*1e6*args.ref_seconds/(W*H)))))
    runs = []
    for i in range(args.warmup + args.steps):
        r = run_reference_binary(scene_path, spp if i >= args.warmup else 1, threads)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "the reference binary printed no render time"}))
            return
        if i >= args.warmup:
            runs.append(r)
    secs = sum(v[1] for v in runs)
    value = W*H*spp*len(runs)/secs/1e6
    rates = sorted(v[0] for v in runs)
    line = {"impl": "reference", "metric": "Msamples/sec (paths x spp)", "value": value, "unit": "Msamples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3*secs/len(runs),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIGS[args.config]["label"]},
            "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": threads, "kind": "reference",
                             "sample": "%d spp of the whole %dx%d frame per step (bounded sample of the job; non-adaptive, so the per-sample rate is "
                                       "spp independent), tungsten -t %d, SSE4.2 Embree build without AVX (the reference's own ISA policy); time = the "
                                       "binary's 'Render time' (excludes scene load + BVH build, as our arm excludes tgb200_create)" % (spp, W, H, threads),
                             "min": rates[0], "median": rates[len(rates)//2], "max": rates[-1], "runs": len(rates),
                             "host": cpu},
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--spp-per-step", type=int, default=64)
    ap.add_argument("--ref-spp", type=int, default=0, help="samples per pixel of one reference step (0 = calibrate to --ref-seconds)")
    ap.add_argument("--ref-seconds", type=float, default=6.0, help="target render time of one reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c1", choices=sorted(CONFIGS))
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    global W, H
    W, H = CONFIGS[args.config]["res"]
    if args.impl == "reference":
        if args.config == "c3":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the stock reference cannot load mesh-instanced scenes from JSON (Instance::loadResources never loads its masters)"}))
            return
        return bench_reference(args, rank, world)

    import numpy as np
    import torch
    from tungsten_b200 import scene, lib, abi, integrator
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the CUDA library is the only implementation")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene_path = make_scene(1024, args.config) if rank == 0 else None
    if world > 1:
        dist.barrier()
        scene_path = make_scene(1024, args.config)
    fs = scene.load_scene(scene_path)
    ctx = lib.Context(fs, device=local_rank)
    info = ctx.scene_info()
    seed = 0xBA5EBA11
    all_tiles = integrator.dice_tiles(W, H, seed)
    my_tiles = integrator.shard_tiles(all_tiles, rank, world)
    spp_step = args.spp_per_step*world
    n_my_pix = sum(t.w*t.h for t in my_tiles)

    # multi-GPU: one all-gather of the tile-major framebuffer (the only collective on the path)
    if world > 1:
        max_pix = max(sum(t.w*t.h for t in integrator.shard_tiles(all_tiles, r, world)) for r in range(world))
        send = torch.zeros(max_pix*3, dtype=torch.float32, device="cuda")
        recv = torch.zeros(world*max_pix*3, dtype=torch.float32, device="cuda")

    def step(i, resident=True, mean=None, count=None):
        if resident:
            ctx.render_resident(spp_step, seed=seed, spp_begin=i*spp_step, tiles=my_tiles)
        else:
            ctx.render_tiles(spp_step, seed=seed, spp_begin=i*spp_step, tiles=my_tiles, mean=mean, count=count)
        if world > 1:
            ctx.pack_tiles(my_tiles, send.data_ptr())
            dist.all_gather_into_tensor(recv, send)

    def timed(n_steps, first, resident, mean=None, count=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n_steps):
            step(first + i, resident, mean, count)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        return t

    # ---- device-resident measurement -----------------------------------------------------------
    ctx.clear()
    for i in range(args.warmup):
        step(i)
    ctx.clear(); ctx.reset_stats(); ctx.set_profiling(True)
    stop = threading.Event(); clk = []
    th = threading.Thread(target=_clock_sampler, args=(stop, clk, local_rank), daemon=True); th.start()
    wall = timed(args.steps, 0, True)
    stop.set(); th.join()
    st = ctx.stats()
    dev_ms = st.total_ms
    if world > 1:
        tt = torch.tensor([float(st.samples), float(st.rays), float(st.hits), float(st.kernel_launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt)
        tot_samples, tot_rays, tot_hits, tot_launches = [float(x) for x in tt.tolist()]
        de = torch.tensor([dev_ms], dtype=torch.float64, device="cuda"); dist.all_reduce(de, op=dist.ReduceOp.MAX)
        dev_ms_max = float(de.item())
    else:
        tot_samples, tot_rays, tot_hits, tot_launches = float(st.samples), float(st.rays), float(st.hits), float(st.kernel_launches)
        dev_ms_max = dev_ms
    value = tot_samples/wall/1e6
    peak, peak_src = hbm_peak()
    trace_gbs = (st.path_rays_traversed*ALG_BYTES_PER_QUERY/1e9)/(st.trace_ms/1e3) if st.trace_ms > 0 else 0.0
    shadow_gbs = (st.shadow_rays_traversed*ALG_BYTES_PER_QUERY/1e9)/(st.shadow_ms/1e3) if st.shadow_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k_trace_dram_bytes_per_launch.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    # ---- end to end through the C ABI with host buffers ------------------------------------------
    ctx.set_profiling(False)
    mean = np.zeros((H, W, 3), dtype=np.float32); count = np.zeros((H, W), dtype=np.uint32)
    e2e_steps = max(2, min(args.steps, 4))
    timed(1, 0, False, mean, count)
    mean[:] = 0; count[:] = 0
    ctx.reset_stats()
    e2e_wall = timed(e2e_steps, 0, False, mean, count)
    e2e_samples = float(ctx.stats().samples)
    if world > 1:
        tt = torch.tensor([e2e_samples], dtype=torch.float64, device="cuda"); dist.all_reduce(tt); e2e_samples = float(tt.item())
    e2e_value = e2e_samples/e2e_wall/1e6
    fb_bytes = W*H*(12 + 4)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline and args.config != "c3":
            host = host_cpu_info()
            cal = run_reference_binary(scene_path, 2, host["usable"])
            if cal is not None:
                spp_c = args.ref_spp if args.ref_spp > 0 else int(min(256, max(4, round(cal[0]*1e6*10.0/(W*H)))))
                r = run_reference_binary(scene_path, spp_c, host["usable"])
                if r is not None:
                    cpu = {"value": r[0], "unit": "Msamples/s", "cores": host["usable"], "kind": "reference",
                           "sample": "%d spp x %dx%d of the same scene rendered in %.2f s by oracle/_ref/tungsten -t %d ('Render time': excludes "
                                     "scene load + BVH build, as the GPU arm excludes tgb200_create)" % (spp_c, W, H, r[1], host["usable"]),
                           "host": host}
        line = {
            "metric": "Msamples/sec (paths x spp)", "value": value, "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3*wall/args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIGS[args.config]["label"] + ", %d spp per step (x%d steps = %d spp)" % (spp_step, args.steps, spp_step*args.steps),
                       "tiles": "16x16, round-robin over %d rank(s)" % world, "paths_in_flight": info["capacity"],
                       "triangles": info["n_tris"], "bvh_nodes": info["n_nodes"], "geom_bytes": info["geom_bytes"],
                       "l2": "per-batch path state (%.0f MB) and geometry exceed the 126 MB L2; no explicit flush" % (info["capacity"]*230/1e6)},
            "mrays_per_s": tot_rays/wall/1e6, "mray_hits_per_s": tot_hits/wall/1e6,
            "device_ms": dev_ms_max, "gpu_launches": int(tot_launches),
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": fb_bytes, "d2h_bytes_per_step": fb_bytes,
                    "steps": e2e_steps},
            "roofline": {"bound": "hbm", "kernel": "k_trace (closest-hit 4-ary BVH traversal of path rays)",
                         "achieved": trace_gbs, "peak": peak, "unit": "GB/s", "frac": trace_gbs/peak, "traffic": traffic,
                         "peak_source": peak_src, "alg_bytes_per_query": ALG_BYTES_PER_QUERY,
                         "queries": int(st.path_rays_traversed), "path_rays_total": int(st.path_rays), "kernel_ms": st.trace_ms, "launches": int(st.trace_launches),
                         "mqueries_per_s": st.path_rays_traversed/st.trace_ms/1e3 if st.trace_ms > 0 else 0.0,
                         "k_shadow": {"achieved": shadow_gbs, "frac": shadow_gbs/peak, "queries": int(st.shadow_rays_traversed), "shadow_rays_total": int(st.shadow_rays), "kernel_ms": st.shadow_ms},
                         "note": "traversal is latency/divergence bound with an L2-resident BVH; see DESIGN.md section 6"},
            "cpu_baseline": cpu,
            "clocks": _summarise_clocks(clk),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
