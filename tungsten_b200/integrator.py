"""Host-side mirror of the reference's Integrator interface for the path_tracer hot path.

`B200PathTraceIntegrator` has the reference's method names, argument meaning and error behaviour
(reference: src/core/integrators/Integrator.hpp:16-63, Integrator.cpp:51;
integrators/path_tracer/PathTraceIntegrator.cpp:27-42,110-134,184-254) and drives the CUDA library through the
C ABI.  It is the Python twin of the C++ adapter in integration/B200PathTraceIntegrator.{hpp,cpp};
the parity tests use it the way `StandaloneRenderer::renderScene` uses an Integrator (src/tungsten/Shared.hpp:255-319).

Host logic only (tile dicing, tile seeds, spp stepping, async start/wait/abort, tile sharding across ranks);
no rendering arithmetic lives here.
"""
import threading

import numpy as np

from . import abi, lib

TILE_SIZE = 16            # PathTraceIntegrator::TileSize (PathTraceIntegrator.hpp:27)
MASK32 = 0xFFFFFFFF


def hash32(x):
    """MathUtil::hash32 (math/MathUtil.hpp:120-128)."""
    x &= MASK32
    x = (~x + (x << 15)) & MASK32
    x ^= x >> 12
    x = (x + (x << 2)) & MASK32
    x ^= x >> 4
    x = (x*2057) & MASK32
    x ^= x >> 16
    return x


class UniformSampler:
    """PCG32 (sampling/UniformSampler.hpp:40-52), sequence 0."""

    def __init__(self, seed):
        self.state = seed & 0xFFFFFFFFFFFFFFFF

    def next_i(self):
        old = self.state
        self.state = (old*6364136223846793005 + 1) & 0xFFFFFFFFFFFFFFFF
        xs = (((old >> 18) ^ old) >> 27) & MASK32
        rot = old >> 59
        return ((xs >> rot) | (xs << ((-rot) & 31))) & MASK32


def dice_tiles(w, h, seed, sampler=None):
    """PathTraceIntegrator::diceTiles with the sampler seeded as in prepareForRender
    (PathTraceIntegrator.cpp:27-42,187): row-major 16x16 tiles, tile sampler seed = hash32(_sampler.nextI()).
    `sampler`: the integrator's own UniformSampler (its stream continues into distributeAdaptiveSamples)."""
    s = sampler if sampler is not None else UniformSampler(hash32(seed))
    n = ((w + TILE_SIZE - 1)//TILE_SIZE)*((h + TILE_SIZE - 1)//TILE_SIZE)
    tiles = (abi.Tile*n)()
    i = 0
    for y in range(0, h, TILE_SIZE):
        for x in range(0, w, TILE_SIZE):
            tiles[i] = abi.Tile(x, y, min(TILE_SIZE, w - x), min(TILE_SIZE, h - y), hash32(s.next_i()))
            i += 1
    return tiles


def _morton2(x, y):
    """Interleave the bits of two 16-bit integers (x in the even bits)."""
    def part(v):
        v &= 0xFFFF
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        v = (v | (v << 1)) & 0x55555555
        return v
    return part(x) | (part(y) << 1)


def shard_order(tiles, deal="morton"):
    """Order in which tiles are dealt to ranks.  "round_robin": the reference's row-major tile ids.  "morton": Z-order of
    the tile grid, so any `world` consecutive tiles form a compact block and every rank's share samples the whole image
    evenly (row-major `id mod N` degenerates to column stripes whenever tiles-per-row is a multiple of N)."""
    idx = list(range(len(tiles)))
    if deal == "morton":
        idx.sort(key=lambda i: _morton2(tiles[i].x//TILE_SIZE, tiles[i].y//TILE_SIZE))
    elif deal != "round_robin":
        raise ValueError(deal)
    return idx


def shard_tiles(tiles, rank, world, deal="morton"):
    """Deal tiles to ranks (position in shard_order mod world): scenes shard by image tile only.  Must match
    tgb200_shard_tiles (csrc/tgb200_api.cu), which the C++ adapter and the library-side multi-GPU path use."""
    order = shard_order(tiles, deal)
    mine = [tiles[order[i]] for i in range(rank, len(order), world)]
    return (abi.Tile*len(mine))(*mine)


class B200PathTraceIntegrator:
    """prepareForRender / startRender / waitForCompletion / abortRender / teardownAfterRender / done /
    currentSpp / nextSpp -- same contract as the reference's PathTraceIntegrator, GPU underneath."""

    def __init__(self, rank=0, world=1, device=-1, max_paths_in_flight=0):
        self._scene = None
        self._ctx = None
        self._current_spp = 0
        self._next_spp = 0
        self._thread = None
        self._error = None
        self.rank, self.world, self.device = rank, world, device
        self.max_paths_in_flight = max_paths_in_flight
        self.tiles = None
        self.all_tiles = None
        self.seed = 0
        self.records = None         # one SampleRecord per 4x4 block (PathTraceIntegrator::_samples)
        self._sampler = None        # PathTraceIntegrator::_sampler

    # -- Integrator.cpp:51
    def _advance_spp(self):
        self._next_spp = min(self._current_spp + self._scene.spp_step, self._scene.spp)

    def prepareForRender(self, flat_scene, seed):
        """PathTraceIntegrator::prepareForRender (PathTraceIntegrator.cpp:184-201)."""
        self._scene = flat_scene
        self.seed = seed & MASK32
        self._current_spp = 0
        self._advance_spp()
        self._ctx = lib.Context(flat_scene, device=self.device, max_paths_in_flight=self.max_paths_in_flight)
        w, h = flat_scene.resolution
        self._sampler = UniformSampler(hash32(self.seed))
        self.all_tiles = dice_tiles(w, h, self.seed, self._sampler)
        self.tiles = shard_tiles(self.all_tiles, self.rank, self.world)
        self.records = (abi.SampleRecord*(((w + 3)//4)*((h + 3)//4)))()
        self._ctx.clear()

    def _generate_work(self):
        """PathTraceIntegrator::generateWork (PathTraceIntegrator.cpp:110-134) through the library's host-side helper."""
        import ctypes as C
        w, h = self._scene.resolution
        st = C.c_uint64(self._sampler.state)
        rc = self._ctx.L.tgb200_generate_work(self.records, w, h, self._current_spp, self._next_spp,
                                              1 if self._scene.adaptive else 0, C.byref(st))
        self._sampler.state = st.value
        if rc < 0:
            raise lib.TgbError(rc, "tgb200_generate_work failed")
        return rc == 1

    def teardownAfterRender(self):
        self.waitForCompletion_noraise()
        if self._ctx is not None:
            self._ctx.close()
        self._ctx = None
        self.tiles = self.all_tiles = None

    def done(self):
        return self._current_spp >= self._next_spp

    def currentSpp(self):
        return self._current_spp

    def nextSpp(self):
        return self._next_spp

    def startRender(self, completionCallback=lambda: None):
        """Asynchronous like the reference (PathTraceIntegrator.cpp:220-239): returns after enqueueing; the
        callback fires when the step's samples are in the framebuffer (or immediately if there is no work)."""
        if self.done() or not self._generate_work():
            self._current_spp = self._next_spp
            self._advance_spp()
            completionCallback()
            return
        begin, count = self._current_spp, self._next_spp - self._current_spp
        self._ctx.clear_abort()

        def work():
            try:
                if self._scene.adaptive:        # per-block sample counts + Welford statistics (SampleRecord) on the device
                    self._ctx.render_adaptive(self.records, seed=self.seed, tiles=self.tiles)
                else:
                    self._ctx.render_resident(count, seed=self.seed, spp_begin=begin, tiles=self.tiles)
                self._current_spp = self._next_spp
                self._advance_spp()
            except Exception as e:      # captured and rethrown from waitForCompletion (thread/TaskGroup.hpp:57-74)
                self._error = e
            finally:
                completionCallback()
        self._thread = threading.Thread(target=work, daemon=True)
        self._thread.start()

    def waitForCompletion_noraise(self):
        if self._thread is not None:
            self._thread.join()
            self._thread = None

    def waitForCompletion(self):
        self.waitForCompletion_noraise()
        if self._error is not None:
            e, self._error = self._error, None
            if isinstance(e, lib.TgbError) and e.code == abi.TGB_ERR_ABORTED:
                return
            raise e

    # -- Integrator::saveRenderResumeData / resumeRender (Integrator.cpp:108-162) + PathTraceIntegrator::saveState/loadState
    #    (PathTraceIntegrator.cpp:158-172): current spp, the colour buffer (running mean + counts), the block records and the
    #    integrator's sampler state.  (Per-tile samplers carry no running state under the per-path reseed contract.)
    def save_state(self):
        mean, count = self._ctx.read_framebuffer()
        return {"current_spp": self._current_spp, "mean": mean, "count": count,
                "records": bytes(self.records), "sampler_state": self._sampler.state}

    def load_state(self, state):
        """After prepareForRender, before the next startRender: continue a render from a saved state."""
        import ctypes as C
        C.memmove(self.records, state["records"], len(state["records"]))
        self._sampler.state = state["sampler_state"]
        self._ctx.write_framebuffer(state["mean"], state["count"])
        self._current_spp = state["current_spp"]
        self._advance_spp()

    def abortRender(self):
        """PathTraceIntegrator::abortRender (PathTraceIntegrator.cpp:249-254): abort + wait."""
        if self._ctx is not None and self._thread is not None:
            self._ctx.abort()
        self.waitForCompletion()

    def framebuffer(self):
        """Camera::getLinear for every pixel (cameras/Camera.hpp:163-172): float32 (h, w, 3), top row first."""
        return self._ctx.read_framebuffer()[0]

    def render(self, flat_scene, seed=0xBA5EBA11):
        """The loop of StandaloneRenderer::renderScene (src/tungsten/Shared.hpp:283-296)."""
        self.prepareForRender(flat_scene, seed)
        try:
            while not self.done():
                self.startRender()
                self.waitForCompletion()
            return self.framebuffer()
        finally:
            self.teardownAfterRender()

    @property
    def context(self):
        return self._ctx
