#include "hair_tables.h"

#include <algorithm>
#include <array>
#include <cmath>

namespace tgb {
namespace {

constexpr float kPi = 3.1415926536f, kTwoPi = kPi*2.0f, kInvTwoPi = 0.5f*(1.0f/kPi), kPiHalf = kPi*0.5f;   // math/Angle.hpp:8-16
constexpr float kEta = 1.55f;                                                                                  // HairBcsdf.hpp:19
constexpr int kRes = 64, kQuadrature = 140, kDetectorSamples = 2048;

struct Rgb { float r, g, b; };
inline float tmax(float a, float b) { return a > b ? a : b; }
inline float tmin(float a, float b) { return a < b ? a : b; }
inline float clampf(float v, float lo, float hi) { return tmin(tmax(v, lo), hi); }

// Fresnel::dielectricReflectance (bsdfs/Fresnel.hpp:75-98)
float dielectric_reflectance(float eta, float cosThetaI) {
    if (cosThetaI < 0.0f) { eta = 1.0f/eta; cosThetaI = -cosThetaI; }
    float sinThetaTSq = eta*eta*(1.0f - cosThetaI*cosThetaI);
    if (sinThetaTSq > 1.0f) return 1.0f;
    float cosThetaT = std::sqrt(tmax(1.0f - sinThetaTSq, 0.0f));
    float Rs = (eta*cosThetaI - cosThetaT)/(eta*cosThetaI + cosThetaT);
    float Rp = (eta*cosThetaT - cosThetaI)/(eta*cosThetaT + cosThetaI);
    return (Rs*Rs + Rp*Rp)*0.5f;
}

// Gauss-Legendre nodes/weights on [-1, 1]: Newton iteration in double from Tricomi's initial guess, rounded to float
struct Quadrature {
    std::array<float, kQuadrature> x, w;
    static double P(double t, int n) {
        if (n == 0) return 1.0;
        if (n == 1) return t;
        double p0 = 1.0, p1 = t;
        for (int i = 2; i <= n; ++i) { double pi = ((2.0*i - 1.0)*t*p1 - (i - 1.0)*p0)/i; p0 = p1; p1 = pi; }
        return p1;
    }
    static double dP(double t, int n) { return n/(t*t - 1.0)*(t*P(t, n) - P(t, n - 1)); }
    Quadrature() {
        const int N = kQuadrature;
        for (int k = 1; k <= N; ++k) {
            double t = std::cos(double(kPi)*(4.0*k - 1.0)/(4.0*N + 2.0))*(1.0 - 1.0/(8.0*N*N) + 1.0/(8.0*N*N*N));
            for (int it = 0; it < 100; ++it) {
                double f = P(t, N);
                t -= f/dP(t, N);
                if (std::abs(f) < 1e-6) break;
            }
            x[k - 1] = float(t);
            float sq = x[k - 1]*x[k - 1];
            double d = dP(double(x[k - 1]), N);
            w[k - 1] = float(2.0/((1.0 - double(sq))*(d*d)));
        }
    }
};

// normalised Gaussian and its 2*pi-periodic wrap (HairBcsdf.cpp:49-72)
float gauss(float beta, float theta) { return std::exp(-theta*theta/(2.0f*beta*beta))/(std::sqrt(2.0f*kPi)*beta); }
float detector(float beta, float phi) {
    float result = 0.0f, delta, shift = 0.0f;
    do {
        delta = gauss(beta, phi + shift) + gauss(beta, phi - shift - kTwoPi);
        result += delta;
        shift += kTwoPi;
    } while (delta > 1e-4f);
    return result;
}
inline float exit_azimuth(float gammaI, float gammaT, int p) { return 2.0f*p*gammaT - 2.0f*gammaI + p*kPi; }   // Phi(), :77-80

// conservative sampling weights of one lobe + per-row distributions
void finish_lobe(HairLobeTables &l) {
    const int S = kRes;
    l.pdfs.resize(size_t(S)*S); l.cdfs.resize(size_t(S + 1)*S); l.sums.resize(S);
    float *w = l.pdfs.data();
    for (int i = 0; i < S*S; ++i) w[i] = tmax(tmax(l.table[3*i], l.table[3*i + 1]), l.table[3*i + 2]);
    for (int y = 0; y < S; ++y) {           // dilate by one texel in x, then in y
        for (int x = 0; x < S - 1; ++x) w[x + y*S] = tmax(w[x + y*S], w[x + 1 + y*S]);
        for (int x = S - 1; x > 0; --x) w[x + y*S] = tmax(w[x + y*S], w[x - 1 + y*S]);
    }
    for (int x = 0; x < S; ++x) {
        for (int y = 0; y < S - 1; ++y) w[x + y*S] = tmax(w[x + y*S], w[x + (y + 1)*S]);
        for (int y = S - 1; y > 0; --y) w[x + y*S] = tmax(w[x + y*S], w[x + (y - 1)*S]);
    }
    for (int row = 0; row < S; ++row) {
        float *pdf = l.pdfs.data() + size_t(row)*S, *cdf = l.cdfs.data() + size_t(row)*(S + 1);
        cdf[0] = 0.0f;
        for (int x = 0; x < S; ++x) cdf[x + 1] = pdf[x] + cdf[x];
        l.sums[row] = cdf[S];
        if (l.sums[row] < 1e-4f) {
            float ratio = 1.0f/S;
            for (int x = 0; x < S; ++x) { pdf[x] = ratio; cdf[x] = x*ratio; }
        } else {
            float scale = 1.0f/l.sums[row];
            for (int x = 0; x < S; ++x) { pdf[x] *= scale; cdf[x] *= scale; }
        }
        cdf[S] = 1.0f;
    }
}

}  // namespace

void hair_precompute(float roughness, float scale_angle_deg, const float sigma_a[3], HairTables &out) {
    const float betaR = tmax(kPiHalf*roughness, 0.04f);
    const float beta[3] = {betaR, betaR*0.5f, betaR*2.0f};
    for (int p = 0; p < 3; ++p) out.v[p] = beta[p]*beta[p];
    out.scale_angle_rad = scale_angle_deg*(kPi/180.0f);
    for (int p = 0; p < 3; ++p) out.lobe[p].table.assign(size_t(3)*kRes*kRes, 0.0f);

    static const Quadrature gl;
    std::array<float, kQuadrature> gammaI;
    for (int i = 0; i < kQuadrature; ++i) gammaI[i] = std::asin(gl.x[i]);

    // the detector is tabulated once (the reference uses the R-lobe width for all three tables, HairBcsdf.cpp:343-347)
    std::vector<float> D(kDetectorSamples);
    for (int i = 0; i < kDetectorSamples; ++i) D[i] = detector(betaR, i/(kDetectorSamples - 1.0f)*kTwoPi);
    auto approxD = [&](float phi) {
        float u = std::abs(phi*(kInvTwoPi*(kDetectorSamples - 1)));
        int x0 = int(u), x1 = x0 + 1;
        u -= x0;
        return D[x0 % kDetectorSamples]*(1.0f - u) + D[x1 % kDetectorSamples]*u;
    };

    for (int y = 0; y < kRes; ++y) {
        float cosHalfAngle = y/(kRes - 1.0f);
        float iorPrime = std::sqrt(kEta*kEta - (1.0f - cosHalfAngle*cosHalfAngle))/cosHalfAngle;
        float invEta = 1.0f/kEta;
        float cosThetaT = std::sqrt(1.0f - (1.0f - cosHalfAngle*cosHalfAngle)*(invEta*invEta));
        Rgb sigmaPrime = {sigma_a[0]/cosThetaT, sigma_a[1]/cosThetaT, sigma_a[2]/cosThetaT};

        std::array<float, kQuadrature> fresnel, gammaT; std::array<Rgb, kQuadrature> absorb;
        for (int i = 0; i < kQuadrature; ++i) {
            gammaT[i] = std::asin(clampf(gl.x[i]/iorPrime, -1.0f, 1.0f));
            fresnel[i] = dielectric_reflectance(1.0f/kEta, cosHalfAngle*std::cos(gammaI[i]));
            float cg = std::cos(gammaT[i]);
            absorb[i] = {std::exp(-sigmaPrime.r*2.0f*cg), std::exp(-sigmaPrime.g*2.0f*cg), std::exp(-sigmaPrime.b*2.0f*cg)};
        }
        for (int xi = 0; xi < kRes; ++xi) {
            float phi = kTwoPi*xi/(kRes - 1.0f);
            float sumR = 0.0f; Rgb sumTT = {0, 0, 0}, sumTRT = {0, 0, 0};
            for (int i = 0; i < kQuadrature; ++i) {
                float fR = fresnel[i]; Rgb T = absorb[i];
                float tt = (1.0f - fR)*(1.0f - fR);
                Rgb ATT = {tt*T.r, tt*T.g, tt*T.b};
                Rgb ATRT = {ATT.r*fR*T.r, ATT.g*fR*T.g, ATT.b*fR*T.b};
                float wR = gl.w[i]*approxD(phi - exit_azimuth(gammaI[i], gammaT[i], 0));
                float wTT = gl.w[i]*approxD(phi - exit_azimuth(gammaI[i], gammaT[i], 1));
                float wTRT = gl.w[i]*approxD(phi - exit_azimuth(gammaI[i], gammaT[i], 2));
                sumR += wR*fR;
                sumTT.r += wTT*ATT.r; sumTT.g += wTT*ATT.g; sumTT.b += wTT*ATT.b;
                sumTRT.r += wTRT*ATRT.r; sumTRT.g += wTRT*ATRT.g; sumTRT.b += wTRT*ATRT.b;
            }
            size_t at = 3*(size_t(xi) + size_t(y)*kRes);
            out.lobe[0].table[at] = out.lobe[0].table[at + 1] = out.lobe[0].table[at + 2] = 0.5f*sumR;
            out.lobe[1].table[at] = 0.5f*sumTT.r; out.lobe[1].table[at + 1] = 0.5f*sumTT.g; out.lobe[1].table[at + 2] = 0.5f*sumTT.b;
            out.lobe[2].table[at] = 0.5f*sumTRT.r; out.lobe[2].table[at + 1] = 0.5f*sumTRT.g; out.lobe[2].table[at + 2] = 0.5f*sumTRT.b;
        }
    }
    for (int p = 0; p < 3; ++p) finish_lobe(out.lobe[p]);
}

}  // namespace tgb
