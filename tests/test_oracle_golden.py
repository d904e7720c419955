"""The CPU oracle against framebuffers rendered by the reference itself (tests/golden/*/ref_*.pfm).

  * ref_stock.pfm    = the UNMODIFIED reference binary; reproduced with supplemental_mode=1 (per-tile serial PCG)
  * ref_pathseed.pfm = reference + the per-path reseed of the supplemental PCG (the parity contract the CUDA path
                       implements); reproduced with supplemental_mode=0

Scenes made only of analytic primitives must match BIT FOR BIT.  Scenes with triangle meshes differ where Embree's
t,u,v (rcpps + one Newton step instead of a division) flips a discrete decision: >= 80% of pixels bit-exact, >= 99%
within 1e-5*(1+L), no systematic error (mean within 0.5%).
"""
import os

import numpy as np
import pytest

from oracle import pyoracle
from tungsten_b200 import scene

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")


def _render(name, mode):
    fs = scene.load_scene(os.path.join(G, name, "scene.json"))
    fs.settings.supplemental_mode = mode
    o = pyoracle.Oracle(fs)
    img, cnt = o.render(fs.spp)
    o.close()
    assert int(cnt.min()) == fs.spp and int(cnt.max()) == fs.spp
    return img


@pytest.mark.parametrize("name", ["cornell", "cornell_short", "cube_city"])
@pytest.mark.parametrize("mode,ref", [(0, "ref_pathseed.pfm"), (1, "ref_stock.pfm")])
def test_analytic_scenes_bit_exact(name, mode, ref):
    img = _render(name, mode)
    want = scene.load_pfm(os.path.join(G, name, ref))
    assert np.array_equal(img, want)


@pytest.mark.parametrize("name", ["cornell_mesh", "materials", "materials_env", "coat_env", "dirac", "many_lights", "coats"])
@pytest.mark.parametrize("mode,ref", [(0, "ref_pathseed.pfm"), (1, "ref_stock.pfm")])
def test_mesh_scenes_close(name, mode, ref):
    img = _render(name, mode)
    want = scene.load_pfm(os.path.join(G, name, ref))
    d = np.abs(img - want).max(axis=2)
    exact = float((d == 0).mean())
    close = float((d <= 1e-5*(1.0 + np.abs(want).max(axis=2))).mean())
    print(name, mode, "exact %.4f close %.4f" % (exact, close))
    assert exact >= 0.80
    assert close >= 0.99
    assert abs(float(img.mean()) - float(want.mean())) <= 5e-3*float(want.mean())


@pytest.mark.parametrize("name", ["hair", "hair_dark", "curves_lambert", "curves_plastic", "hair_sky"])
@pytest.mark.parametrize("mode,ref", [(0, "ref_pathseed.pfm"), (1, "ref_stock.pfm")])
def test_curve_scenes_close(name, mode, ref):
    """Curves (Curves.cpp) + hair BCSDF (HairBcsdf.cpp) against the reference's own renders.  The oracle's segment BVH
    is not the reference's, and the reference's per-segment bisection prunes with a bound that is not strictly
    conservative, so a few grazing rays resolve differently: >= 98.5% of pixels bit-exact under the per-path reseed contract;
    with the stock serial PCG stream one such ray also shifts the Russian-roulette draws of the rest of its tile: >= 95%
    (hair_sky -- the shipped hair scene's emitters, infinite_sphere_cap + skydome, min_bounces 1, deep hair paths under a bright
    sky -- draws more roulette decisions per tile: >= 90 % there)."""
    img = _render(name, mode)
    want = scene.load_pfm(os.path.join(G, name, ref))
    d = np.abs(img - want).max(axis=2)
    exact = float((d == 0).mean())
    close = float((d <= 1e-5*(1.0 + np.abs(want).max(axis=2))).mean())
    print(name, mode, "exact %.4f close %.4f" % (exact, close))
    stock_bar = 0.90 if name == "hair_sky" else 0.95
    assert exact >= (0.985 if mode == 0 else stock_bar)
    assert close >= (0.985 if mode == 0 else stock_bar)
    assert abs(float(img.mean()) - float(want.mean())) <= 5e-3*float(want.mean())


def test_pathseed_and_stock_references_differ():
    a = scene.load_pfm(os.path.join(G, "cornell", "ref_pathseed.pfm"))
    b = scene.load_pfm(os.path.join(G, "cornell", "ref_stock.pfm"))
    assert not np.array_equal(a, b)        # Russian roulette draws differ -> the contract matters


@pytest.mark.skipif(not os.path.exists("/root/reference/data/materialtest/materialtest.json") or
                    not os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "tungsten_pathseed")),
                    reason="needs the mounted reference (build container only)")
def test_real_materialtest_scene_against_reference_binary(tmp_path):
    """BASELINE.json config C0, the reference's own shipped scene (80,768 triangles, smooth_coat over rough_conductor,
    checker floor, envmap.hdr importance sampling): oracle vs the reference binary at 96x96, 8 spp."""
    import json, shutil, subprocess
    src = "/root/reference/data/materialtest"
    for f in os.listdir(src):
        shutil.copy(os.path.join(src, f), tmp_path)
    js = json.load(open(tmp_path/"materialtest.json"))
    js["camera"]["resolution"] = [96, 96]
    js["renderer"].update(spp=8, spp_step=8, adaptive_sampling=False, stratified_sampler=True, hdr_output_file="out.pfm",
                          output_file="out.png", overwrite_output_files=True)
    json.dump(js, open(tmp_path/"mt.json", "w"))
    exe = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "tungsten_pathseed")
    subprocess.check_call([exe, "-t", "4", "-d", str(tmp_path/"ref"), str(tmp_path/"mt.json")], stdout=subprocess.DEVNULL)
    want = scene.load_pfm(str(tmp_path/"ref"/"out.pfm"))
    fs = scene.load_scene(str(tmp_path/"mt.json"))
    o = pyoracle.Oracle(fs); img, _ = o.render(8); o.close()
    d = np.abs(img - want).max(axis=2)
    assert float((d == 0).mean()) >= 0.7
    assert float((d <= 1e-5*(1.0 + np.abs(want).max(axis=2))).mean()) >= 0.99


@pytest.mark.skipif(not os.path.exists("/root/reference/data/example-scenes/hair/scene.json") or
                    not os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "tungsten_pathseed")),
                    reason="needs the mounted reference (build container only)")
def test_real_hair_scene_against_reference_binary(tmp_path):
    """BASELINE.json config C4, the reference's own shipped hair scene (curl.fiber: 670,000 nodes / 10,000 curves, `subsample`
    0.5, hair BCSDF, bcsdf_cylinder) with its two out-of-scope emitters (infinite_sphere_cap + skydome) replaced by an
    environment sphere and a quad light: oracle vs the reference binary at 90x300, 2 spp."""
    import json, subprocess
    src = "/root/reference/data/example-scenes/hair"
    os.symlink(os.path.join(src, "curl.fiber"), tmp_path/"curl.fiber")
    js = json.load(open(os.path.join(src, "scene.json")))
    js["primitives"] = [p for p in js["primitives"] if p["type"] == "curves"] + [
        {"name": "env", "type": "infinite_sphere", "emission": [0.5, 0.6, 0.8], "sample": True},
        {"name": "key", "type": "quad", "emission": [60, 55, 50], "bsdf": {"type": "null"},
         "transform": {"position": [3.0, 9.0, 6.0], "scale": [3, 1, 3], "rotation": [0, 0, 140]}}]
    js["camera"]["resolution"] = [90, 300]
    js["integrator"]["min_bounces"] = 0            # the shipped value 1 hides directly visible emitters: show the sky
    js["renderer"].update(spp=2, spp_step=2, adaptive_sampling=False, stratified_sampler=True, hdr_output_file="out.pfm",
                          output_file="out.png", overwrite_output_files=True)
    json.dump(js, open(tmp_path/"hair.json", "w"))
    exe = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "tungsten_pathseed")
    subprocess.check_call([exe, "-t", "8", "-d", str(tmp_path/"ref"), str(tmp_path/"hair.json")], stdout=subprocess.DEVNULL)
    want = scene.load_pfm(str(tmp_path/"ref"/"out.pfm"))
    fs = scene.load_scene(str(tmp_path/"hair.json"))
    assert 300000 < fs.n_curve_segments < 350000                 # `subsample` 0.5 drops about half of the 650,000 segments
    o = pyoracle.Oracle(fs); img, _ = o.render(2); o.close()
    d = np.abs(img - want).max(axis=2)
    exact = float((d == 0).mean()); close = float((d <= 1e-5*(1.0 + np.abs(want).max(axis=2))).mean())
    print("real hair scene: exact %.4f close %.4f mean %.4f/%.4f" % (exact, close, img.mean(), want.mean()))
    assert float((want.max(axis=2) > 0).mean()) > 0.9          # not a black frame
    assert close >= 0.97
    assert abs(float(img.mean()) - float(want.mean())) <= 1e-2*float(want.mean())


@pytest.mark.skipif(not os.path.exists("/root/reference/data/example-scenes/hair/scene.json") or
                    not os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "tungsten_pathseed")),
                    reason="needs the mounted reference (build container only)")
def test_shipped_hair_scene_unmodified_against_reference_binary(tmp_path):
    """BASELINE.json config C4 AS SHIPPED: data/example-scenes/hair/scene.json with its own emitters (infinite_sphere_cap sun,
    sampled; skydome, unsampled), its own integrator settings (min_bounces 1, max_bounces 64) -- only resolution / spp are
    reduced and the skydome gets the "sky_image" the flattener needs (Skydome::prepareForRender's image, dumped through the
    reference's own classes; the reference ignores the key).  Oracle vs the reference binary."""
    import json, subprocess, sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    src = "/root/reference/data/example-scenes/hair"
    os.symlink(os.path.join(src, "curl.fiber"), tmp_path/"curl.fiber")
    js = json.load(open(os.path.join(src, "scene.json")))
    assert sorted(p["type"] for p in js["primitives"]) == ["curves", "infinite_sphere_cap", "skydome"]
    js["camera"]["resolution"] = [90, 300]
    js["renderer"].update(spp=2, spp_step=2, adaptive_sampling=False, stratified_sampler=True, hdr_output_file="out.pfm",
                          output_file="out.png", overwrite_output_files=True)
    for p in js["primitives"]:
        if p["type"] == "skydome":
            p["sky_image"] = "sky.pfm"
    json.dump(js, open(tmp_path/"hair.json", "w"))
    make_golden.dump_sky_image(str(tmp_path/"hair.json"), str(tmp_path/"sky.pfm"))
    exe = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "tungsten_pathseed")
    subprocess.check_call([exe, "-t", "8", "-d", str(tmp_path/"ref"), str(tmp_path/"hair.json")], stdout=subprocess.DEVNULL)
    want = scene.load_pfm(str(tmp_path/"ref"/"out.pfm"))
    fs = scene.load_scene(str(tmp_path/"hair.json"))
    o = pyoracle.Oracle(fs); img, _ = o.render(2); o.close()
    d = np.abs(img - want).max(axis=2)
    exact = float((d == 0).mean()); close = float((d <= 1e-5*(1.0 + np.abs(want).max(axis=2))).mean())
    print("shipped hair scene, unmodified lights: exact %.4f close %.4f mean %.4f/%.4f" % (exact, close, img.mean(), want.mean()))
    assert float((want.max(axis=2) > 0).mean()) > 0.05         # (min_bounces 1 hides the directly visible sky: only the strands are lit)
    assert close >= 0.97
    assert abs(float(img.mean()) - float(want.mean())) <= 1e-2*float(want.mean())
