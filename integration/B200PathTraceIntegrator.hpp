// In-tree adapter: a Tungsten `Integrator` whose render loop runs on a B200 through libtgb200.so.
//
// Drop-in for `PathTraceIntegrator` (reference: src/core/integrators/path_tracer/PathTraceIntegrator.{hpp,cpp});
// registered in IntegratorFactory as "b200_path_tracer" (INTEGRATION.md).  It reads the scene ONLY through
// TraceableScene / Primitive / Bsdf / Texture public accessors, hands the prepared world-space geometry to
// tgb200_create(), renders spp steps with tgb200_render_resident() and loads the resulting running mean into
// the camera's OutputBuffer through its public deserialize() (the same path resume files use).
#ifndef B200PATHTRACEINTEGRATOR_HPP_
#define B200PATHTRACEINTEGRATOR_HPP_

#include "integrators/Integrator.hpp"
#include "integrators/path_tracer/PathTracerSettings.hpp"
#include "sampling/UniformSampler.hpp"

#include "tgb200.h"

#include <atomic>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace Tungsten {

class B200PathTraceIntegrator : public Integrator
{
    PathTracerSettings _settings;
    tgb_ctx *_ctx;
    uint32 _seed;
    std::unique_ptr<std::thread> _worker;
    std::string _error;
    // PathTraceIntegrator::_sampler / _samples (PathTraceIntegrator.hpp): the integrator's own PCG stream (tile seeds, adaptive
    // sample distribution) and one SampleRecord per 4x4 pixel block
    UniformSampler _sampler;
    std::vector<tgb_sample_record> _samples;
    bool _needsFramebufferPush;      // after loadState(): the camera's colour buffer has to go to the device first
    std::vector<int> _devices;       // "devices": [0, 1, ...] in the integrator's JSON block: GPUs to spread the tiles over (empty = the current one)

    void uploadFramebuffer();
    void pushFramebuffer();
    bool generateWork();

protected:
    virtual void saveState(OutputStreamHandle &out) override;
    virtual void loadState(InputStreamHandle &in) override;

public:
    B200PathTraceIntegrator();
    virtual ~B200PathTraceIntegrator();

    virtual void fromJson(JsonPtr value, const Scene &scene) override;
    virtual rapidjson::Value toJson(Allocator &allocator) const override;

    virtual bool supportsResumeRender() const override;

    virtual void prepareForRender(TraceableScene &scene, uint32 seed) override;
    virtual void teardownAfterRender() override;

    virtual void startRender(std::function<void()> completionCallback) override;
    virtual void waitForCompletion() override;
    virtual void abortRender() override;
};

}

#endif /* B200PATHTRACEINTEGRATOR_HPP_ */
