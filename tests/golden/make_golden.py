"""Regenerates the golden fixtures in this directory FROM THE REFERENCE (run in the build container only:
needs /root/reference and the binaries built by `make -C oracle/ref`).

  kat.json            known-answer values printed by a ~60-line host program compiled against the reference's
                      own headers (hash32, PCG32, normalizedUint, sobol::sample, cosineHemisphere, Fresnel terms,
                      computeDiffuseFresnel, the tent filter CDF, TangentFrame)
  <scene>/            scene JSON + .wo3 written by tungsten_b200.synth, plus
  <scene>/ref_pathseed.pfm   framebuffer of oracle/_ref/tungsten_pathseed (per-path reseed contract)
  <scene>/ref_stock.pfm      framebuffer of the UNMODIFIED reference binary

The fixtures are small (64x64) so that the CPU test-suite stays fast; they travel to the GPU box with the
repository, /root/reference does not.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

KAT_CPP = r'''
#include <cstdio>
#include <cmath>
#include "math/MathUtil.hpp"
#include "math/BitManip.hpp"
#include "math/TangentFrame.hpp"
#include "sampling/UniformSampler.hpp"
#include "sampling/SampleWarp.hpp"
#include "bsdfs/Fresnel.hpp"
#include "bsdfs/Microfacet.hpp"
#include <sobol/sobol.h>
using namespace Tungsten;
static void pf(const char *k, float v, bool last = false) { union { float f; unsigned u; } c; c.f = v; printf("\"%s\": %u%s\n", k, c.u, last ? "" : ","); }
int main() {
    printf("{\n");
    printf("\"hash32\": [");
    unsigned xs[] = {0u, 1u, 2u, 0xBA5EBA11u, 0xFFFFFFFFu, 123456789u};
    for (int i = 0; i < 6; ++i) printf("[%u, %u]%s", xs[i], MathUtil::hash32(xs[i]), i < 5 ? ", " : "");
    printf("],\n\"pcg\": [");
    UniformSampler s(MathUtil::hash32(0xBA5EBA11u));
    for (int i = 0; i < 8; ++i) printf("%u%s", s.nextI(), i < 7 ? ", " : "");
    printf("],\n\"pcg_seed\": %u,\n\"sobol\": [", MathUtil::hash32(0xBA5EBA11u));
    unsigned idx[] = {0u, 1u, 2u, 3u, 255u, 256u, 1023u, 65535u, 0x12345678u, 0xFFFFFFFFu};
    unsigned dims[] = {0u, 1u, 2u, 7u, 100u, 1023u};
    bool first = true;
    for (unsigned i : idx) for (unsigned d : dims) { printf("%s[%u, %u, %u, %u]", first ? "" : ", ", i, d, 0x9E3779B9u, sobol::sample(i, d, 0x9E3779B9u)); first = false; }
    printf("],\n\"normalized_uint\": [");
    for (int i = 0; i < 6; ++i) { union { float f; unsigned u; } c; c.f = BitManip::normalizedUint(xs[i]); printf("[%u, %u]%s", xs[i], c.u, i < 5 ? ", " : ""); }
    printf("],\n");
    Vec3f ch = SampleWarp::cosineHemisphere(Vec2f(0.3f, 0.7f));
    pf("cos_hemi_x", ch.x()); pf("cos_hemi_y", ch.y()); pf("cos_hemi_z", ch.z());
    pf("dielectric_1.5_0.3", Fresnel::dielectricReflectance(1.0f/1.5f, 0.3f));
    pf("dielectric_1.5_m0.3", Fresnel::dielectricReflectance(1.0f/1.5f, -0.3f));
    pf("conductor_cu_r_0.4", Fresnel::conductorReflectance(0.200438f, 3.91295f, 0.4f));
    pf("diffuse_fresnel_1.5", Fresnel::computeDiffuseFresnel(1.5f, 1000000));
    pf("diffuse_fresnel_1.4", Fresnel::computeDiffuseFresnel(1.4f, 1000000));
    pf("power_heuristic", SampleWarp::powerHeuristic(0.3f, 1.7f));
    TangentFrame tf(Vec3f(0.36f, -0.48f, 0.8f));
    pf("frame_tx", tf.tangent.x()); pf("frame_ty", tf.tangent.y()); pf("frame_tz", tf.tangent.z());
    pf("frame_bx", tf.bitangent.x()); pf("frame_by", tf.bitangent.y()); pf("frame_bz", tf.bitangent.z());
    pf("ggx_D", Microfacet::D(Microfacet::Distribution("ggx"), 0.25f, Vec3f(0.1f, 0.2f, 0.9746794f)));
    pf("beckmann_G1", Microfacet::G1(Microfacet::Distribution("beckmann"), 0.3f, Vec3f(0.6f, 0.0f, 0.8f), Vec3f(0.0f, 0.0f, 1.0f)), true);
    printf("}\n");
}
'''


def make_kat():
    d = tempfile.mkdtemp()
    src = os.path.join(d, "kat.cpp")
    open(src, "w").write(KAT_CPP)
    exe = os.path.join(d, "kat")
    objs = [os.path.join(ROOT, "oracle", "_ref", "obj", "tp", "sobol", "sobol.o")]
    core = os.path.join(ROOT, "oracle", "_ref", "obj", "core")
    # StringableEnum tables for Microfacet::Distribution live in bsdfs/Microfacet.cpp
    objs += [os.path.join(core, "bsdfs", "Microfacet.o")]
    cmd = ["g++", "-std=c++11", "-O2", "-march=core2", "-mno-fma", "-DCONSTEXPR=constexpr", "-I" + REF + "/src/core",
           "-I" + REF + "/src/thirdparty", "-I" + REF + "/src", src] + objs + ["-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        # fall back to linking the whole core archive-free object set
        print(r.stdout[-3000:])
        raise SystemExit("KAT program failed to build")
    out = subprocess.check_output([exe], text=True)
    json.loads(out)
    open(os.path.join(HERE, "kat.json"), "w").write(out)
    shutil.rmtree(d)


def make_scenes():
    from tungsten_b200 import synth
    scenes = {}
    res, spp = (64, 64), 8
    d = os.path.join(HERE, "cornell"); os.makedirs(d, exist_ok=True)
    scenes["cornell"] = synth.write_scene(d, "scene", synth.cornell_box(res=res, spp=spp))
    d = os.path.join(HERE, "cornell_short"); os.makedirs(d, exist_ok=True)
    scenes["cornell_short"] = synth.write_scene(d, "scene", synth.cornell_box(res=res, spp=spp, max_bounces=3, filter_name="gaussian"))
    scenes["cornell_mesh"] = synth.cornell_mesh(os.path.join(HERE, "cornell_mesh"), "scene", subdiv=2, res=res, spp=spp)
    scenes["materials"] = synth.material_room(os.path.join(HERE, "materials"), "scene", res=res, spp=spp, subdiv=2)
    scenes["materials_env"] = synth.material_room(os.path.join(HERE, "materials_env"), "scene", res=res, spp=spp, subdiv=2, env=[0.4, 0.5, 0.7])
    scenes["coat_env"] = synth.materialtest_standin(os.path.join(HERE, "coat_env"), "scene", res=res, spp=spp, subdiv=2, env_res=(64, 32))
    # C4 stand-ins: curves primitive + hair BCSDF (bcsdf_cylinder), and the two other curve modes with ordinary lobes
    scenes["hair"] = synth.hair_scene(os.path.join(HERE, "hair"), "scene", n_curves=400, res=res, spp=spp)
    scenes["hair_dark"] = synth.hair_scene(os.path.join(HERE, "hair_dark"), "scene", n_curves=300, res=res, spp=spp, thickness=0.008,
                                           subsample=0.5, env=None, bsdf={"type": "hair", "roughness": 0.1, "scale_angle": 3, "sigma_a": [0.1, 0.2, 0.5]})
    scenes["curves_lambert"] = synth.hair_scene(os.path.join(HERE, "curves_lambert"), "scene", n_curves=300, res=res, spp=spp, mode="half_cylinder",
                                                bsdf={"type": "lambert", "albedo": [0.6, 0.4, 0.2]}, width=0.02)
    scenes["curves_plastic"] = synth.hair_scene(os.path.join(HERE, "curves_plastic"), "scene", n_curves=300, res=res, spp=spp, mode="cylinder",
                                                bsdf={"type": "rough_plastic", "albedo": [0.6, 0.4, 0.2], "roughness": 0.2}, thickness=0.015,
                                                taper=True, subsample=0.3)
    for name, path in scenes.items():
        d = os.path.dirname(path)
        for exe, tag in (("tungsten_pathseed", "ref_pathseed"), ("tungsten", "ref_stock")):
            out = tempfile.mkdtemp()
            subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", exe), "-t", "4", "-d", out, path], stdout=subprocess.DEVNULL)
            shutil.copy(os.path.join(out, "out.pfm"), os.path.join(d, tag + ".pfm"))
            shutil.rmtree(out)
        print("golden:", name)


if __name__ == "__main__":
    make_kat()
    make_scenes()
