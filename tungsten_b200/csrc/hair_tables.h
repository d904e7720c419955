// Host-side precomputation for the hair BCSDF (reference: HairBcsdf::prepareForRender + precomputeAzimuthalDistributions,
// src/core/bsdfs/HairBcsdf.cpp:318-446; PrecomputedAzimuthalLobe.cpp:7-33; sampling/InterpolatedDistribution1D.hpp:42-72;
// math/GaussLegendre.hpp).  Runs once per hair material in tgb200_create; the device code only reads the tables.
#pragma once
#include <vector>

namespace tgb {

struct HairLobeTables {
    std::vector<float> table;   // 64 x 64 RGB: N_p(phi, cos(theta_d)), row = cos(theta_d)
    std::vector<float> pdfs;    // 64 x 64: dilated max-channel weights, normalised per row
    std::vector<float> cdfs;    // 64 x 65: running sums of pdfs per row
    std::vector<float> sums;    // 64: row sums before normalisation
};
struct HairTables {
    float v[3];                 // longitudinal variances beta_R^2, beta_TT^2, beta_TRT^2
    float scale_angle_rad;
    HairLobeTables lobe[3];     // R, TT, TRT
};

void hair_precompute(float roughness, float scale_angle_deg, const float sigma_a[3], HairTables &out);

}  // namespace tgb
