"""N>1 host path on CPU: world_size-2 gloo.  Each rank owns the round-robin share of the 16x16 tiles, produces
its tile-major float3 buffer, the ranks all-gather once, and every rank de-tiles the same full image.
The renderer stand-in here is the CPU oracle (tests may use it); the property checked is the sharding +
gather + de-tile plumbing that bench.py / the C++ adapter use on GPUs: result identical to a 1-rank render."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pack(img, tiles):
    return np.concatenate([img[t.y:t.y + t.h, t.x:t.x + t.w].reshape(-1, 3) for t in tiles], axis=0)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tungsten_b200 import scene, synth, integrator
    from oracle import pyoracle
    fs = scene.load_scene(synth.cornell_box(res=(72, 40), spp=2))
    w, h = fs.resolution
    tiles = integrator.dice_tiles(w, h, 0xBA5EBA11)
    mine = integrator.shard_tiles(tiles, rank, world)
    o = pyoracle.Oracle(fs)
    img, _ = o.render(2, tiles=mine, threads=1)
    npix = [sum(t.w*t.h for t in integrator.shard_tiles(tiles, r, world)) for r in range(world)]
    send = torch.zeros(max(npix)*3)
    send[:npix[rank]*3] = torch.from_numpy(_pack(img, mine).reshape(-1))
    recv = torch.zeros(world*max(npix)*3)
    dist.all_gather_into_tensor(recv, send)
    full = np.zeros((h, w, 3), dtype=np.float32)
    for r in range(world):
        buf = recv[r*max(npix)*3:(r*max(npix) + npix[r])*3].numpy().reshape(-1, 3)
        k = 0
        for t in integrator.shard_tiles(tiles, r, world):
            full[t.y:t.y + t.h, t.x:t.x + t.w] = buf[k:k + t.w*t.h].reshape(t.h, t.w, 3); k += t.w*t.h
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_tile_shard_allgather(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path/"rank0.npy"); b = np.load(tmp_path/"rank1.npy")
    assert np.array_equal(a, b)
    sys.path.insert(0, ROOT)
    from tungsten_b200 import scene, synth
    from oracle import pyoracle
    fs = scene.load_scene(synth.cornell_box(res=(72, 40), spp=2))
    ref, _ = pyoracle.Oracle(fs).render(2)
    assert np.array_equal(a, ref)          # partition-independent: same image as a single-rank render
