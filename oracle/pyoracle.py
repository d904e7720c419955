"""ctypes wrapper around oracle/_build/liboracle.so (CPU restatement; TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

from tungsten_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "_build", "liboracle.so")
    src = [os.path.join(_HERE, "pt_oracle.c"), os.path.join(_HERE, "pt_oracle.h"),
           os.path.join(_HERE, "..", "include", "tgb200.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_build/liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(abi.SceneDesc), C.c_void_p, C.c_char_p, C.c_int]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_render_tiles.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                          C.c_void_p, C.c_void_p, C.c_int, C.POINTER(abi.Stats)]
        L.oracle_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.oracle_dice_tiles.restype = C.c_uint32
        L.oracle_dice_tiles.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.oracle_hash32.restype = C.c_uint32; L.oracle_hash32.argtypes = [C.c_uint32]
        L.oracle_pcg_next.restype = C.c_uint32; L.oracle_pcg_next.argtypes = [C.POINTER(C.c_uint64)]
        L.oracle_normalized_uint.restype = C.c_float; L.oracle_normalized_uint.argtypes = [C.c_uint32]
        L.oracle_sobol_sample.restype = C.c_uint32
        L.oracle_sobol_sample.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.oracle_filter_cdf.argtypes = [C.c_uint32, C.c_void_p, C.POINTER(C.c_float)]
        L.oracle_diffuse_fresnel.restype = C.c_float; L.oracle_diffuse_fresnel.argtypes = [C.c_float, C.c_int]
        L.oracle_kat_eval.restype = C.c_int; L.oracle_kat_eval.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def sobol_table():
    p = os.path.join(_HERE, "..", "tungsten_b200", "data", "sobol_1024x32.u32")
    return np.fromfile(p, dtype="<u4").reshape(1024, 32)


class Oracle:
    def __init__(self, flat_scene):
        self.fs = flat_scene
        self._sobol = np.ascontiguousarray(sobol_table())
        self._desc = flat_scene.desc()
        err = C.create_string_buffer(512)
        self.h = lib().oracle_create(C.byref(self._desc), self._sobol.ctypes.data, err, 512)
        if not self.h:
            raise RuntimeError("oracle_create failed: %s" % err.value.decode())
        self.stats = abi.Stats()

    def close(self):
        if self.h:
            lib().oracle_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dice_tiles(self, seed):
        w, h = self.fs.resolution
        n = lib().oracle_dice_tiles(w, h, seed, None)
        tiles = (abi.Tile*n)()
        lib().oracle_dice_tiles(w, h, seed, tiles)
        return tiles

    def render(self, spp, seed=0xBA5EBA11, spp_begin=0, tiles=None, threads=0, mean=None, count=None):
        w, h = self.fs.resolution
        if mean is None:
            mean = np.zeros((h, w, 3), dtype=np.float32)
        if count is None:
            count = np.zeros((h, w), dtype=np.uint32)
        n = 0 if tiles is None else len(tiles)
        rc = lib().oracle_render_tiles(self.h, tiles, n, seed, spp_begin, spp, mean.ctypes.data, count.ctypes.data,
                                       threads, C.byref(self.stats))
        if rc != 0:
            raise RuntimeError("oracle_render_tiles -> %d" % rc)
        return mean, count

    def trace(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        hits = (abi.Hit*len(rays))()
        lib().oracle_trace_closest(self.h, rays.ctypes.data, hits, len(rays))
        return np.ctypeslib.as_array(hits).copy() if len(rays) else np.zeros(0)


def kat_eval(which, args):
    a = np.array(args, dtype=np.float32); out = np.zeros(8, dtype=np.float32)
    n = lib().oracle_kat_eval(which, a.ctypes.data, out.ctypes.data)
    return out[:n]
