#include "B200PathTraceIntegrator.hpp"

#include "tgb200.h"

#include "renderer/TraceableScene.hpp"
#include "cameras/PinholeCamera.hpp"
#include "cameras/OutputBufferSettings.hpp"
#include "primitives/InfiniteSphere.hpp"
#include "primitives/InfiniteSphereCap.hpp"
#include "primitives/Skydome.hpp"
#include "math/Angle.hpp"
#include "primitives/TriangleMesh.hpp"
#include "primitives/Curves.hpp"
#include "bsdfs/HairBcsdf.hpp"
#include "io/CurveIO.hpp"
#include "sampling/UniformSampler.hpp"
#include "primitives/Quad.hpp"
#include "primitives/Cube.hpp"
#include "bsdfs/RoughDielectricBsdf.hpp"
#include "bsdfs/DielectricBsdf.hpp"
#include "bsdfs/ConductorBsdf.hpp"
#include "bsdfs/MirrorBsdf.hpp"
#include "bsdfs/RoughConductorBsdf.hpp"
#include "bsdfs/RoughPlasticBsdf.hpp"
#include "bsdfs/SmoothCoatBsdf.hpp"
#include "bsdfs/RoughCoatBsdf.hpp"
#include "bsdfs/PlasticBsdf.hpp"
#include "bsdfs/LambertBsdf.hpp"
#include "bsdfs/NullBsdf.hpp"
#include "textures/ConstantTexture.hpp"
#include "textures/CheckerTexture.hpp"
#include "textures/BitmapTexture.hpp"
#include "io/JsonObject.hpp"
#include "Debug.hpp"

#include <cstring>
#include <map>
#include <sstream>
#include <vector>

namespace Tungsten {

namespace {

// Owns every buffer a tgb_scene_desc points to for the duration of tgb200_create().
struct Flattener
{
    std::vector<tgb_texture> textures;
    std::vector<tgb_bsdf> bsdfs;
    std::vector<tgb_primitive> prims;
    std::vector<uint32_t> slots;
    std::vector<std::vector<float>> texelStore;
    std::vector<std::vector<tgb_vertex>> vertStore;
    std::vector<std::vector<tgb_triangle>> triStore;
    std::vector<std::vector<Vec4f>> nodeStore;
    std::vector<std::vector<uint32_t>> segStore;
    rapidjson::Document jsonDoc;    // allocator for the toJson() calls that read back private parameters
    std::map<const Bsdf *, int> bsdfIds;
    std::map<const Texture *, int> texIds;
    const std::vector<std::shared_ptr<Primitive>> *scenePrims = nullptr;

    static void put(float *dst, const Vec3f &v) { dst[0] = v.x(); dst[1] = v.y(); dst[2] = v.z(); }

    int texture(const Texture *t)
    {
        auto iter = texIds.find(t);
        if (iter != texIds.end())
            return iter->second;
        tgb_texture o;
        std::memset(&o, 0, sizeof(o));
        if (const CheckerTexture *c = dynamic_cast<const CheckerTexture *>(t)) {
            o.type = TGB_TEX_CHECKER;
            put(o.value, c->onColor());
            put(o.value2, c->offColor());
            o.res_u = c->resU();
            o.res_v = c->resV();
        } else if (const BitmapTexture *b = dynamic_cast<const BitmapTexture *>(t)) {
            // texel centres through the public operator[] of a NEAREST-filtered clone: (x+0.5)/w and 1-(y+0.5)/h land inside texel
            // (x, y) for every size, so the stored texels come back exactly (a bilinear read-back would mix in a neighbour with
            // a weight of ~1e-7..1e-4 whenever w or h is not a power of two)
            std::unique_ptr<Texture> nearestOwner(b->clone());
            BitmapTexture *nearest = static_cast<BitmapTexture *>(nearestOwner.get());
            nearest->setLinear(false);
            o.type = TGB_TEX_BITMAP;
            o.res_u = b->w();
            o.res_v = b->h();
            o.flags = (b->linear() ? 1u : 0u) | (b->clamp() ? 2u : 0u);
            texelStore.emplace_back(size_t(b->w())*b->h()*3);
            std::vector<float> &tx = texelStore.back();
            for (int y = 0; y < b->h(); ++y) {
                for (int x = 0; x < b->w(); ++x) {
                    Vec3f c = (*nearest)[Vec2f((x + 0.5f)/b->w(), 1.0f - (y + 0.5f)/b->h())];
                    put(&tx[3*(size_t(y)*b->w() + x)], c);
                }
            }
            o.texels = tx.data();
        } else if (t->isConstant()) {
            o.type = TGB_TEX_CONSTANT;
            put(o.value, t->average());
        } else {
            FAIL("b200_path_tracer: texture type outside the hot path");
        }
        textures.push_back(o);
        texIds[t] = int(textures.size()) - 1;
        return texIds[t];
    }

    static uint32_t distribution(const char *name)
    {
        if (std::strcmp(name, "beckmann") == 0) return TGB_DIST_BECKMANN;
        if (std::strcmp(name, "phong") == 0) return TGB_DIST_PHONG;
        return TGB_DIST_GGX;
    }

    int bsdf(const Bsdf *b)
    {
        auto iter = bsdfIds.find(b);
        if (iter != bsdfIds.end())
            return iter->second;
        if (b->bump() && !b->bump()->isConstant())
            FAIL("b200_path_tracer: bump maps are outside the hot path");
        tgb_bsdf o;
        std::memset(&o, 0, sizeof(o));
        o.albedo_tex = texture(b->albedo().get());
        o.roughness_tex = -1;
        o.substrate = -1;
        o.ior = 1.5f;
        o.thickness = 1.0f;
        o.enable_refraction = 1;
        if (dynamic_cast<const NullBsdf *>(b)) {
            o.type = TGB_BSDF_NULL;
        } else if (dynamic_cast<const LambertBsdf *>(b)) {
            o.type = TGB_BSDF_LAMBERT;
        } else if (const RoughConductorBsdf *c = dynamic_cast<const RoughConductorBsdf *>(b)) {
            o.type = TGB_BSDF_ROUGH_CONDUCTOR;
            o.distribution = distribution(c->distributionName());
            o.roughness_tex = texture(c->roughness().get());
            put(o.eta, c->eta());
            put(o.k, c->k());
        } else if (dynamic_cast<const MirrorBsdf *>(b)) {
            o.type = TGB_BSDF_MIRROR;
        } else if (const ConductorBsdf *cd = dynamic_cast<const ConductorBsdf *>(b)) {
            o.type = TGB_BSDF_CONDUCTOR;
            put(o.eta, cd->eta());
            put(o.k, cd->k());
        } else if (const DielectricBsdf *di = dynamic_cast<const DielectricBsdf *>(b)) {
            o.type = TGB_BSDF_DIELECTRIC;
            o.ior = di->ior();
            o.enable_refraction = di->enableTransmission() ? 1 : 0;
        } else if (const RoughDielectricBsdf *d = dynamic_cast<const RoughDielectricBsdf *>(b)) {
            o.type = TGB_BSDF_ROUGH_DIELECTRIC;
            o.distribution = distribution(d->distributionName());
            o.roughness_tex = texture(d->roughness().get());
            o.ior = d->ior();
            o.enable_refraction = d->enableTransmission() ? 1 : 0;
        } else if (const RoughPlasticBsdf *rp = dynamic_cast<const RoughPlasticBsdf *>(b)) {
            o.type = TGB_BSDF_ROUGH_PLASTIC;
            o.distribution = distribution(rp->distributionName());
            o.roughness_tex = texture(rp->roughness().get());
            o.ior = rp->ior();
            o.thickness = rp->thickness();
            put(o.sigma_a, rp->sigmaA());
        } else if (const PlasticBsdf *p = dynamic_cast<const PlasticBsdf *>(b)) {
            o.type = TGB_BSDF_PLASTIC;
            o.ior = p->ior();
            o.thickness = p->thickness();
            put(o.sigma_a, p->sigmaA());
        } else if (const SmoothCoatBsdf *sc = dynamic_cast<const SmoothCoatBsdf *>(b)) {
            o.type = TGB_BSDF_SMOOTH_COAT;
            o.ior = sc->ior();
            o.thickness = sc->thickness();
            put(o.sigma_a, sc->sigmaA());
            o.substrate = bsdf(sc->substrate().get());
        } else if (const RoughCoatBsdf *rc = dynamic_cast<const RoughCoatBsdf *>(b)) {
            o.type = TGB_BSDF_ROUGH_COAT;
            o.distribution = distribution(rc->distributionName());
            o.roughness_tex = texture(rc->roughness().get());
            o.ior = rc->ior();
            o.thickness = rc->thickness();
            put(o.sigma_a, rc->sigmaA());
            o.substrate = bsdf(rc->substrate().get());
        } else if (const HairBcsdf *hb = dynamic_cast<const HairBcsdf *>(b)) {
            // HairBcsdf keeps its parameters private; its own toJson() returns them (HairBcsdf.cpp:173-181),
            // the melanin mix is HairBcsdf::prepareForRender's (:435-443)
            o.type = TGB_BSDF_HAIR;
            rapidjson::Value v = hb->toJson(jsonDoc.GetAllocator());
            o.hair_scale_angle_deg = float(v["scale_angle"].GetDouble());
            o.hair_roughness = float(v["roughness"].GetDouble());
            Vec3f sigmaA;
            if (v.HasMember("sigma_a")) {
                const rapidjson::Value &s = v["sigma_a"];
                sigmaA = s.IsArray() ? Vec3f(float(s[0u].GetDouble()), float(s[1u].GetDouble()), float(s[2u].GetDouble()))
                                     : Vec3f(float(s.GetDouble()));
            } else {
                const Vec3f eumelaninSigmaA = Vec3f(0.419f, 0.697f, 1.37f);
                const Vec3f pheomelaninSigmaA = Vec3f(0.187f, 0.4f, 1.05f);
                sigmaA = float(v["melanin_concentration"].GetDouble())
                        *lerp(eumelaninSigmaA, pheomelaninSigmaA, float(v["melanin_ratio"].GetDouble()));
            }
            put(o.sigma_a, sigmaA);
        } else {
            FAIL("b200_path_tracer: BSDF type outside the hot path");
        }
        bsdfs.push_back(o);
        bsdfIds[b] = int(bsdfs.size()) - 1;
        return bsdfIds[b];
    }

    void primitive(Primitive &p)
    {
        tgb_primitive o;
        std::memset(&o, 0, sizeof(o));
        o.emission_tex = p.emission() ? texture(p.emission().get()) : -1;
        o.bsdf_first = uint32_t(slots.size());
        o.bsdf_count = uint32_t(p.numBsdfs());
        for (int i = 0; i < p.numBsdfs(); ++i)
            slots.push_back(uint32_t(bsdf(p.bsdf(i).get())));
        const Mat4f &tform = p.transform();
        if (TriangleMesh *m = dynamic_cast<TriangleMesh *>(&p)) {
            // TriangleMesh::prepareForRender (TriangleMesh.cpp:539-552) with the reference's own Mat4f arithmetic
            o.type = TGB_PRIM_MESH;
            o.smooth = m->smoothed() ? 1 : 0;
            Mat4f normalTform(tform.toNormalMatrix());
            vertStore.emplace_back(m->verts().size());
            triStore.emplace_back(m->tris().size());
            std::vector<tgb_vertex> &vs = vertStore.back();
            std::vector<tgb_triangle> &ts = triStore.back();
            for (size_t i = 0; i < vs.size(); ++i) {
                const Vertex &v = m->verts()[i];
                put(vs[i].pos, tform*v.pos());
                put(vs[i].normal, normalTform.transformVector(v.normal()));
                vs[i].uv[0] = v.uv().x();
                vs[i].uv[1] = v.uv().y();
            }
            for (size_t i = 0; i < ts.size(); ++i) {
                const TriangleI &t = m->tris()[i];
                ts[i].v0 = t.v0; ts[i].v1 = t.v1; ts[i].v2 = t.v2; ts[i].material = t.material;
            }
            o.verts = vs.data(); o.n_verts = uint32_t(vs.size());
            o.tris = ts.data(); o.n_tris = uint32_t(ts.size());
        } else if (Curves *cv = dynamic_cast<Curves *>(&p)) {
            // Curves exposes neither its nodes nor its settings; the scene's own Curves::prepareForRender has already
            // moved the (private) nodes to world space.  Re-read the .fiber file through the reference's CurveIO and redo
            // Curves::loadCurves + prepareForRender (Curves.cpp:268-296,572-611) with the parameters toJson() reports.
            o.type = TGB_PRIM_CURVES;
            rapidjson::Value v = cv->toJson(jsonDoc.GetAllocator());
            std::string mode = v["mode"].GetString();
            if (mode == "cylinder") o.curve_mode = TGB_CURVE_CYLINDER;
            else if (mode == "half_cylinder") o.curve_mode = TGB_CURVE_HALF_CYLINDER;
            else if (mode == "bcsdf_cylinder") o.curve_mode = TGB_CURVE_BCSDF_CYLINDER;
            else FAIL("b200_path_tracer: curve mode '%s' is outside the hot path", mode);
            bool taper = v["curve_taper"].GetBool();
            float subsample = float(v["subsample"].GetDouble());
            bool overrideThickness = v.HasMember("curve_thickness");
            float thickness = overrideThickness ? float(v["curve_thickness"].GetDouble()) : 0.0f;
            std::vector<uint32> curveEnds;
            nodeStore.emplace_back();
            std::vector<Vec4f> &nodes = nodeStore.back();
            CurveIO::CurveData data;
            data.curveEnds = &curveEnds;
            data.nodeData = &nodes;
            if (!cv->path() || !CurveIO::load(*cv->path(), data))
                FAIL("b200_path_tracer: cannot read the curve file of '%s'", p.name());
            if (overrideThickness || taper) {
                for (size_t i = 0; i < curveEnds.size(); ++i) {
                    uint32 start = i ? curveEnds[i - 1] : 0;
                    for (uint32 t = start; t < curveEnds[i]; ++t) {
                        float w = overrideThickness ? thickness : nodes[t].w();
                        if (taper)
                            w *= 1.0f - (t - start - 0.5f)/(curveEnds[i] - start - 1);
                        nodes[t].w() = w;
                    }
                }
            }
            float widthScale = tform.extractScaleVec().avg();
            for (Vec4f &n : nodes) {
                Vec3f q = tform*n.xyz();
                n = Vec4f(q.x(), q.y(), q.z(), n.w()*widthScale);
            }
            segStore.emplace_back();
            std::vector<uint32_t> &segs = segStore.back();
            UniformSampler rand;
            for (size_t i = 0; i < curveEnds.size(); ++i) {
                uint32 start = i ? curveEnds[i - 1] : 0;
                if (subsample > 0.0f && rand.next1D() < subsample)
                    continue;
                for (uint32 t = start + 2; t < curveEnds[i]; ++t)
                    segs.push_back(t);
            }
            o.curve_nodes = reinterpret_cast<const float *>(nodes.data());
            o.n_curve_nodes = uint32_t(nodes.size());
            o.curve_segments = segs.data();
            o.n_curve_segments = uint32_t(segs.size());
        } else if (dynamic_cast<Quad *>(&p)) {
            // Quad::prepareForRender (Quad.cpp:298-305)
            o.type = TGB_PRIM_QUAD;
            Vec3f base = tform*Vec3f(0.0f);
            Vec3f edge0 = tform.transformVector(Vec3f(1.0f, 0.0f, 0.0f));
            Vec3f edge1 = tform.transformVector(Vec3f(0.0f, 0.0f, 1.0f));
            base -= edge0*0.5f;
            base -= edge1*0.5f;
            put(o.base, base); put(o.edge0, edge0); put(o.edge1, edge1);
        } else if (dynamic_cast<Cube *>(&p)) {
            // Cube::prepareForRender (Cube.cpp:351-355)
            o.type = TGB_PRIM_CUBE;
            put(o.pos, tform*Vec3f(0.0f));
            put(o.scale, tform.extractScale()*Vec3f(0.5f));
            Mat4f rot = tform.extractRotation();
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                    o.rot[r*3 + c] = rot[r*4 + c];
        } else if (InfiniteSphereCap *cap = dynamic_cast<InfiniteSphereCap *>(&p)) {
            // InfiniteSphereCap::prepareForRender (InfiniteSphereCap.cpp:231-247); the pivot object's name is private: read it back
            o.type = TGB_PRIM_INFINITE_SPHERE_CAP;
            Mat4f capTform = tform;
            rapidjson::Value js = cap->toJson(jsonDoc.GetAllocator());
            if (js.HasMember("skydome") && js["skydome"].IsString() && scenePrims) {
                for (const std::shared_ptr<Primitive> &q : *scenePrims)
                    if (q->name() == js["skydome"].GetString())
                        capTform = q->transform();
            }
            put(o.cap_dir, capTform.transformVector(Vec3f(0.0f, 1.0f, 0.0f)).normalized());
            o.cap_cos = std::cos(Angle::degToRad(cap->capAngleDeg()));
            o.do_sample = p.isSamplable() ? 1 : 0;
            o.bsdf_count = 0;
        } else if (dynamic_cast<Skydome *>(&p)) {
            // Skydome::prepareForRender has already built the 512x256 sky image: it IS the primitive's emission texture
            // (Skydome.cpp:317-320); the library treats it as an environment sphere without rotation
            o.type = TGB_PRIM_SKYDOME;
            if (o.emission_tex < 0 || textures[size_t(o.emission_tex)].type != TGB_TEX_BITMAP)
                FAIL("b200_path_tracer: skydome without a prepared sky image");
            const float id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            std::memcpy(o.rot, id, sizeof(id));
            o.do_sample = p.isSamplable() ? 1 : 0;
            o.bsdf_count = 0;
        } else if (dynamic_cast<InfiniteSphere *>(&p)) {
            o.type = TGB_PRIM_INFINITE_SPHERE;
            Mat4f rot = tform.extractRotation();
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                    o.rot[r*3 + c] = rot[r*4 + c];
            o.do_sample = p.isSamplable() ? 1 : 0;
            o.bsdf_count = 0;
        } else {
            FAIL("b200_path_tracer: primitive type outside the hot path");
        }
        prims.push_back(o);
    }
};

}

B200PathTraceIntegrator::B200PathTraceIntegrator()
: Integrator(),
  _ctx(nullptr),
  _seed(0xBA5EBA11),
  _sampler(0xBA5EBA11),
  _needsFramebufferPush(false)
{
}

B200PathTraceIntegrator::~B200PathTraceIntegrator()
{
    if (_ctx)
        tgb200_destroy(_ctx);
}

void B200PathTraceIntegrator::fromJson(JsonPtr value, const Scene &/*scene*/)
{
    _settings.fromJson(value);
    // "devices": [0, 1, ...] (or a single number N = devices 0..N-1): tgb_settings::devices, the in-library multi-GPU path
    _devices.clear();
    if (auto d = value["devices"]) {
        if (d.isArray()) {
            for (unsigned i = 0; i < d.size(); ++i)
                _devices.push_back(d[i].cast<int>());
        } else {
            int n = d.cast<int>();
            for (int i = 0; i < n; ++i)
                _devices.push_back(i);
        }
        if (_devices.size() > 8)
            value.parseError("b200_path_tracer: at most 8 devices");
    }
}

rapidjson::Value B200PathTraceIntegrator::toJson(Allocator &allocator) const
{
    rapidjson::Value v = _settings.toJson(allocator);
    v["type"].SetString("b200_path_tracer");
    if (!_devices.empty()) {
        rapidjson::Value a(rapidjson::kArrayType);
        for (int d : _devices)
            a.PushBack(d, allocator);
        v.AddMember("devices", a, allocator);
    }
    return v;
}

// Resume files (Integrator::saveRenderResumeData / resumeRender, Integrator.cpp:108-162): the block records field by field
// as SampleRecord::saveState streams them (SampleRecord.hpp:24-42), then the integrator's sampler state.  (The reference also
// streams one sampler per tile: its supplemental PCG runs on across a tile's pixels; under the per-path reseed contract a
// tile's sampler is its seed, which diceTiles re-derives from the render seed.)
void B200PathTraceIntegrator::saveState(OutputStreamHandle &out)
{
    for (const tgb_sample_record &r : _samples) {
        FileUtils::streamWrite(out, r.sample_count);
        FileUtils::streamWrite(out, r.next_sample_count);
        FileUtils::streamWrite(out, r.sample_index);
        FileUtils::streamWrite(out, r.adaptive_weight);
        FileUtils::streamWrite(out, r.mean);
        FileUtils::streamWrite(out, r.running_variance);
    }
    _sampler.saveState(out);
}

void B200PathTraceIntegrator::loadState(InputStreamHandle &in)
{
    for (tgb_sample_record &r : _samples) {
        FileUtils::streamRead(in, r.sample_count);
        FileUtils::streamRead(in, r.next_sample_count);
        FileUtils::streamRead(in, r.sample_index);
        FileUtils::streamRead(in, r.adaptive_weight);
        FileUtils::streamRead(in, r.mean);
        FileUtils::streamRead(in, r.running_variance);
    }
    _sampler.loadState(in);
    _needsFramebufferPush = true;      // resumeRender() has just deserialised the camera's output buffers (Integrator.cpp:152)
}

bool B200PathTraceIntegrator::supportsResumeRender() const
{
    return true;
}

// PathTraceIntegrator::generateWork (PathTraceIntegrator.cpp:110-134) through the library's host-side implementation
bool B200PathTraceIntegrator::generateWork()
{
    Vec2u res = _scene->cam().resolution();
    uint64_t state = _sampler.state();
    int rc = tgb200_generate_work(_samples.data(), res.x(), res.y(), _currentSpp, _nextSpp,
            _scene->rendererSettings().useAdaptiveSampling() ? 1 : 0, &state);
    _sampler = UniformSampler(state);
    if (rc < 0)
        FAIL("b200_path_tracer: tgb200_generate_work failed");
    return rc == 1;
}

void B200PathTraceIntegrator::prepareForRender(TraceableScene &scene, uint32 seed)
{
    _currentSpp = 0;
    _seed = seed;
    _scene = &scene;
    advanceSpp();
    scene.cam().requestColorBuffer();

    if (!scene.rendererSettings().useSobol())
        FAIL("b200_path_tracer needs renderer.stratified_sampler=true");
    if (!scene.media().empty() || scene.cam().medium())
        FAIL("b200_path_tracer: participating media are outside the hot path");
    // The framebuffer hand-off (uploadFramebuffer) fills ONE plain colour buffer: running mean + sample counts.  Feature
    // buffers (depth / normal / albedo / visibility) and the colour buffer's two_buffer_variance / sample_variance planes
    // (cameras/OutputBuffer.hpp:95-132, Camera.cpp:148-160) are not produced by the GPU path: fail loudly, never write garbage.
    for (const OutputBufferSettings &b : scene.rendererSettings().renderOutputs()) {
        if (b.type() != OutputColor)
            FAIL("b200_path_tracer: renderer.output_buffers entry '%s' is outside the hot path (only 'color' is produced)", b.typeString());
        if (b.twoBufferVariance() || b.sampleVariance())
            FAIL("b200_path_tracer: two_buffer_variance / sample_variance on the colour buffer are outside the hot path");
    }
    PinholeCamera *cam = dynamic_cast<PinholeCamera *>(&scene.cam());
    if (!cam)
        FAIL("b200_path_tracer: only the pinhole camera is on the hot path");

    Flattener f;
    f.scenePrims = &scene.primitives();
    for (const std::shared_ptr<Primitive> &p : scene.primitives())
        f.primitive(*p);

    tgb_scene_desc desc;
    std::memset(&desc, 0, sizeof(desc));
    desc.abi_version = TGB200_ABI_VERSION;
    Flattener::put(desc.camera.pos, cam->pos());
    const Mat4f &t = cam->transform();                // already has setRight(-right) applied (Camera.cpp:62)
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            desc.camera.xform[r*3 + c] = t[r*4 + c];
    desc.camera.fov_deg = cam->fovDeg();
    desc.camera.res_x = cam->resolution().x();
    desc.camera.res_y = cam->resolution().y();
    const std::string filter = cam->reconstructionFilter().name();
    const char *names[] = {"dirac", "box", "tent", "gaussian", "mitchell_netravali", "catmull_rom", "lanczos"};
    for (uint32_t i = 0; i < 7; ++i)
        if (filter == names[i])
            desc.camera.filter = i;
    desc.settings.min_bounces = _settings.minBounces;
    desc.settings.max_bounces = _settings.maxBounces;
    desc.settings.enable_light_sampling = _settings.enableLightSampling;
    desc.settings.enable_two_sided_shading = _settings.enableTwoSidedShading;
    desc.settings.enable_consistency_checks = _settings.enableConsistencyChecks;
    desc.settings.use_sobol = 1;
    desc.settings.supplemental_mode = 0;
    desc.settings.device = -1;
    if (_devices.size() == 1)
        desc.settings.device = _devices[0];
    desc.settings.n_devices = uint32_t(_devices.size());
    for (size_t i = 0; i < _devices.size(); ++i)
        desc.settings.devices[i] = _devices[i];
    desc.primitives = f.prims.data(); desc.n_primitives = uint32_t(f.prims.size());
    desc.bsdfs = f.bsdfs.data(); desc.n_bsdfs = uint32_t(f.bsdfs.size());
    desc.bsdf_slots = f.slots.data(); desc.n_bsdf_slots = uint32_t(f.slots.size());
    desc.textures = f.textures.data(); desc.n_textures = uint32_t(f.textures.size());

    if (tgb200_create(&desc, &_ctx) != TGB_OK)
        FAIL("b200_path_tracer: %s", tgb200_last_error(nullptr));
    tgb200_clear_framebuffer(_ctx);

    // PathTraceIntegrator::prepareForRender (PathTraceIntegrator.cpp:184-201): _sampler is seeded with hash32(seed) and
    // diceTiles draws one value per 16x16 tile (the library derives the same tile seeds from `seed`); what follows in the
    // stream feeds distributeAdaptiveSamples.
    Vec2u res = cam->resolution();
    _sampler = UniformSampler(MathUtil::hash32(seed));
    for (uint32 y = 0; y < res.y(); y += 16)
        for (uint32 x = 0; x < res.x(); x += 16)
            _sampler.nextI();
    _samples.assign(size_t((res.x() + 3)/4)*((res.y() + 3)/4), tgb_sample_record());
    std::memset(_samples.data(), 0, _samples.size()*sizeof(tgb_sample_record));
    _needsFramebufferPush = false;
}

void B200PathTraceIntegrator::teardownAfterRender()
{
    waitForCompletion();
    if (_ctx)
        tgb200_destroy(_ctx);
    _ctx = nullptr;
}

// Camera::colorBuffer() <- GPU running mean + counts, through OutputBuffer's public deserialize()
void B200PathTraceIntegrator::uploadFramebuffer()
{
    Vec2u res = _scene->cam().resolution();
    size_t n = size_t(res.x())*res.y();
    std::string blob(n*(sizeof(Vec3f) + sizeof(uint32)), '\0');
    float *rgb = reinterpret_cast<float *>(&blob[0]);
    uint32 *count = reinterpret_cast<uint32 *>(&blob[n*sizeof(Vec3f)]);
    if (tgb200_read_framebuffer(_ctx, rgb, count) != TGB_OK) {
        _error = tgb200_last_error(_ctx);
        return;
    }
    InputStreamHandle in(new std::istringstream(blob, std::ios_base::in | std::ios_base::binary));
    _scene->cam().colorBuffer()->deserialize(in);
}

// Camera::colorBuffer() -> GPU (after a resume): OutputBuffer::serialize is the mirror of deserialize
void B200PathTraceIntegrator::pushFramebuffer()
{
    Vec2u res = _scene->cam().resolution();
    size_t n = size_t(res.x())*res.y();
    std::ostringstream *os = new std::ostringstream(std::ios_base::out | std::ios_base::binary);
    OutputStreamHandle out(os);
    _scene->cam().colorBuffer()->serialize(out);
    std::string blob = os->str();
    if (blob.size() != n*(sizeof(Vec3f) + sizeof(uint32)))
        FAIL("b200_path_tracer: unexpected colour buffer layout in the resume data");
    if (tgb200_write_framebuffer(_ctx, reinterpret_cast<const float *>(blob.data()),
            reinterpret_cast<const uint32_t *>(blob.data() + n*sizeof(Vec3f))) != TGB_OK)
        FAIL("b200_path_tracer: %s", tgb200_last_error(_ctx));
}

void B200PathTraceIntegrator::startRender(std::function<void()> completionCallback)
{
    if (done() || !generateWork()) {
        _currentSpp = _nextSpp;
        advanceSpp();
        completionCallback();
        return;
    }
    if (_needsFramebufferPush) {
        pushFramebuffer();
        _needsFramebufferPush = false;
    }
    uint32 begin = _currentSpp, count = _nextSpp - _currentSpp;
    const bool adaptive = _scene->rendererSettings().useAdaptiveSampling();
    tgb200_clear_abort(_ctx);          // an abortRender() from now on cancels THIS step, even before the worker reaches the library
    _worker.reset(new std::thread([this, begin, count, adaptive, completionCallback]() {
        // adaptive scenes: per-block sample counts from generateWork, SampleRecord statistics updated on the device
        int rc = adaptive ? tgb200_render_adaptive(_ctx, nullptr, 0, _seed, _samples.data())
                          : tgb200_render_resident(_ctx, nullptr, 0, _seed, begin, count);
        if (rc == TGB_OK) {
            uploadFramebuffer();
            _currentSpp = _nextSpp;
            advanceSpp();
        } else if (rc != TGB_ERR_ABORTED) {
            _error = tgb200_last_error(_ctx);
        }
        completionCallback();
    }));
}

void B200PathTraceIntegrator::waitForCompletion()
{
    if (_worker) {
        _worker->join();
        _worker.reset();
    }
    if (!_error.empty()) {
        std::string e;
        e.swap(_error);
        FAIL("b200_path_tracer: %s", e);
    }
}

void B200PathTraceIntegrator::abortRender()
{
    if (_worker && _ctx)
        tgb200_abort(_ctx);
    waitForCompletion();
}

}
