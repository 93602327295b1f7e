// ok_spectral.h — ORACLE (test infrastructure, NOT product code).
// RGB->spectrum lookup, uplifts, D65, CIE XYZ, piecewise-linear spectra, filter + camera sampling.
#pragma once
#include "ok_core.h"
#include "../include/hikari_cuda.h"

namespace ok {

// ---------------------------------------------------------------------------------------------
// RGBSigmoidPolynomial  src/spectral/rgb2spec.jl:17-53
// ---------------------------------------------------------------------------------------------
struct Poly { float c0, c1, c2; };
inline float sigmoid(float x) {
    if (std::isinf(x)) return x > 0 ? 1.0f : 0.0f;
    return 0.5f + x / (2.0f * std::sqrt(1.0f + x * x));
}
inline float poly_eval(const Poly& p, float lambda) {
    float x = p.c0 * lambda * lambda + p.c1 * lambda + p.c2;
    return sigmoid(x);
}
inline float poly_max_value(const Poly& p) {
    float result = std::max(poly_eval(p, 360.0f), poly_eval(p, 830.0f));
    if (p.c0 != 0) {
        float lc = -p.c1 / (2.0f * p.c0);
        if (360.0f <= lc && lc <= 830.0f) result = std::max(result, poly_eval(p, lc));
    }
    return result;
}

struct Tables {
    const uint32_t* sobol;
    const float *cie_x, *cie_y, *cie_z, *d65;
    int32_t res;
    const float* scale;
    const float* coeffs;  // Julia column-major Array(3,res,res,res,3) indexed [maxc,z,y,x,coef]
    inline float coef(int maxc, int zi, int yi, int xi, int c) const {  // all 1-based like the reference
        size_t r = (size_t)res;
        return coeffs[(size_t)(maxc - 1) + 3 * ((size_t)(zi - 1) + r * ((size_t)(yi - 1) + r * ((size_t)(xi - 1) + r * (size_t)(c - 1))))];
    }
};

// src/spectral/rgb2spec.jl:83-167
inline Poly rgb_to_spectrum(const Tables& T, float r, float g, float b) {
    r = clampf(r, 0.0f, 1.0f); g = clampf(g, 0.0f, 1.0f); b = clampf(b, 0.0f, 1.0f);
    if (r == g && g == b) {
        float c2;
        if (r > 0.0f && r < 1.0f) c2 = (r - 0.5f) / std::sqrt(r * (1.0f - r));
        else if (r <= 0.0f) c2 = -1.0e10f;
        else c2 = 1.0e10f;
        return Poly{0.0f, 0.0f, c2};
    }
    int maxc = r > g ? (r > b ? 1 : 3) : (g > b ? 2 : 3);
    float z = maxc == 1 ? r : (maxc == 2 ? g : b);
    float xc = maxc == 1 ? g : (maxc == 2 ? b : r);
    float yc = maxc == 1 ? b : (maxc == 2 ? r : g);
    int res = T.res;
    float x = xc * (float)(res - 1) / z;
    float y = yc * (float)(res - 1) / z;
    int zi = 1;
    for (int i = 1; i <= res - 1; i++) if (T.scale[i - 1] < z) zi = i;
    zi = std::min(zi, res - 1);
    int xi = std::min(u_int32(x) + 1, res - 1);
    int yi = std::min(u_int32(y) + 1, res - 1);
    float dx = x - (float)(xi - 1);
    float dy = y - (float)(yi - 1);
    float dz = (z - T.scale[zi - 1]) / (T.scale[zi] - T.scale[zi - 1]);
    float c[3];
    for (int k = 1; k <= 3; k++) {
        c[k - 1] = (1.0f - dz) * ((1.0f - dy) * ((1.0f - dx) * T.coef(maxc, zi, yi, xi, k) + dx * T.coef(maxc, zi, yi, xi + 1, k)) +
                                  dy * ((1.0f - dx) * T.coef(maxc, zi, yi + 1, xi, k) + dx * T.coef(maxc, zi, yi + 1, xi + 1, k))) +
                   dz * ((1.0f - dy) * ((1.0f - dx) * T.coef(maxc, zi + 1, yi, xi, k) + dx * T.coef(maxc, zi + 1, yi, xi + 1, k)) +
                         dy * ((1.0f - dx) * T.coef(maxc, zi + 1, yi + 1, xi, k) + dx * T.coef(maxc, zi + 1, yi + 1, xi + 1, k)));
    }
    return Poly{c[0], c[1], c[2]};
}

// src/spectral/uplift.jl:255-266  (uplift_rgb, sigmoid method)
inline Spec uplift_rgb(const Tables& T, const float* rgb, const Wavelengths& l) {
    Poly p = rgb_to_spectrum(T, rgb[0], rgb[1], rgb[2]);
    return Spec(poly_eval(p, l.lambda[0]), poly_eval(p, l.lambda[1]), poly_eval(p, l.lambda[2]), poly_eval(p, l.lambda[3]));
}
// src/spectral/uplift.jl:286-308
inline Spec uplift_rgb_unbounded(const Tables& T, const float* rgb, const Wavelengths& l) {
    float m = std::max(std::max(rgb[0], rgb[1]), rgb[2]);
    if (m <= 0.0f) return Spec(0.0f);
    Poly p = rgb_to_spectrum(T, rgb[0] / m, rgb[1] / m, rgb[2] / m);
    float scale = m / poly_max_value(p);
    return Spec(scale * poly_eval(p, l.lambda[0]), scale * poly_eval(p, l.lambda[1]), scale * poly_eval(p, l.lambda[2]), scale * poly_eval(p, l.lambda[3]));
}
// src/spectral/uplift.jl:437-457
inline float sample_d65(const Tables& T, float lambda) {
    if (lambda <= 300.0f) return T.d65[0];
    else if (lambda >= 830.0f) return T.d65[106];
    float t = (lambda - 300.0f) / 5.0f;
    int idx = floor_int32(t) + 1;
    idx = clampi(idx, 1, 106);
    float frac = t - (float)floor_int32(t);
    float v0 = T.d65[idx - 1], v1 = T.d65[idx];
    return v0 * (1.0f - frac) + v1 * frac;
}
// src/spectral/uplift.jl:514-538
inline Spec uplift_rgb_illuminant(const Tables& T, const float* rgb, const Wavelengths& l) {
    float m = std::max(std::max(rgb[0], rgb[1]), rgb[2]);
    if (m <= 0.0f) return Spec(0.0f);
    float scale = 2.0f * m;
    Poly p = rgb_to_spectrum(T, rgb[0] / scale, rgb[1] / scale, rgb[2] / scale);
    Spec r;
    for (int i = 0; i < 4; i++) r.v[i] = scale * poly_eval(p, l.lambda[i]) * sample_d65(T, l.lambda[i]);
    return r;
}
// src/spectral/uplift.jl:496-505  Sample(s::RGBIlluminantSpectrum, lambda)
inline Spec sample_illuminant_baked(const Tables& T, const float* poly, float scale, const Wavelengths& l) {
    Poly p{poly[0], poly[1], poly[2]};
    Spec r;
    for (int i = 0; i < 4; i++) r.v[i] = scale * poly_eval(p, l.lambda[i]) * sample_d65(T, l.lambda[i]);
    return r;
}
// Sample(table, light.i, lambda): dispatch on the two spectrum kinds (uplift.jl:556-571)
inline Spec sample_light_spectrum(const Tables& T, const HkLight& L, const Wavelengths& l) {
    if (L.spectrum_kind == HK_SPECTRUM_ILLUMINANT) return sample_illuminant_baked(T, L.poly, L.illum_scale, l);
    return uplift_rgb_illuminant(T, L.rgb, l);
}

// src/spectral/piecewise-linear.jl:11-31
inline float pls_sample(const float* lam, const float* val, int N, float l) {
    if (l <= lam[0]) return val[0];
    if (l >= lam[N - 1]) return val[N - 1];
    int lo = 1, hi = N;
    while (lo + 1 < hi) {
        int mid = (lo + hi) >> 1;
        if (lam[mid - 1] <= l) lo = mid; else hi = mid;
    }
    float t = (l - lam[lo - 1]) / (lam[hi - 1] - lam[lo - 1]);
    return val[lo - 1] * (1.0f - t) + val[hi - 1] * t;
}

// src/spectral/color.jl:364-440, 572-579
inline float sample_cie(const float* tab, float lambda) {
    int off = round_int32(lambda) - 360;
    if (off < 0 || off >= 471) return 0.0f;
    return tab[off];
}
inline V3 spectral_to_xyz(const Tables& T, const Spec& L, const Wavelengths& l) {
    V3 sum(0.0f);
    for (int i = 0; i < 4; i++) {
        float pdf = l.pdf[i];
        if (pdf != 0.0f) {
            V3 cmf(sample_cie(T.cie_x, l.lambda[i]), sample_cie(T.cie_y, l.lambda[i]), sample_cie(T.cie_z, l.lambda[i]));
            sum = sum + cmf * L.v[i] / pdf;
        }
    }
    return sum * 0.25f;
}
inline V3 xyz_to_linear_srgb(V3 c) {
    return V3(3.2404542f * c.x - 1.5371385f * c.y - 0.4985314f * c.z,
              -0.9692660f * c.x + 1.8760108f * c.y + 0.0415560f * c.z,
              0.0556434f * c.x - 0.2040259f * c.y + 1.0572252f * c.z);
}

// ---------------------------------------------------------------------------------------------
// Filter sampling   src/filter.jl:101-118, 727-953
// ---------------------------------------------------------------------------------------------
struct FilterSample { V2 p; float weight; };
inline float sample_tent(float u, float r) {
    if (u < 0.5f) { float ur = 2.0f * u; return -r + r * std::sqrt(ur); }
    float ur = 2.0f * (1.0f - u);
    return r * (1.0f - std::sqrt(ur));
}
// filter.jl:727-741 find_interval; arrays are 0-based here, returned index is 1-based like the reference
inline int filter_find_interval(const float* cdf, float u, int n, int stride = 1) {
    int lo = 1, hi = n + 1;
    for (int it = 0; it < 20; it++) {
        int mid = (lo + hi) >> 1;
        bool c = cdf[(mid - 1) * stride] <= u;
        lo = c ? mid : lo;
        hi = c ? hi : mid;
    }
    return lo;
}
// filter.jl:834-870
inline FilterSample filter_sample_tabulated(const HkFilter& F, V2 u) {
    // marginal (y)
    int ny = F.ny, nx = F.nx;
    int o = clampi(filter_find_interval(F.marginal_cdf, u.y, ny), 1, ny);
    float du = u.y - F.marginal_cdf[o - 1];
    float diff = F.marginal_cdf[o] - F.marginal_cdf[o - 1];
    if (diff > 0.0f) du /= diff; else du = 0.0f;
    float pdf_y = F.func_integral > 0.0f ? F.marginal_func[o - 1] / F.func_integral : 0.0f;
    float t = ((float)(o - 1) + du) / (float)ny;
    float py = lerpf(F.domain_min[1], F.domain_max[1], t);
    int iy = o;
    // conditional (x | y)
    float row_integral = F.marginal_func[iy - 1];
    const float* ccdf = F.conditional_cdf + (size_t)(iy - 1) * (nx + 1);
    int ox = clampi(filter_find_interval(ccdf, u.x, nx), 1, nx);
    float dux = u.x - ccdf[ox - 1];
    float diffx = ccdf[ox] - ccdf[ox - 1];
    if (diffx > 0.0f) dux /= diffx; else dux = 0.0f;
    float fval = F.func[(size_t)(iy - 1) * nx + (ox - 1)];
    float pdf_x = row_integral > 0.0f ? fval / row_integral : 0.0f;
    float tx = ((float)(ox - 1) + dux) / (float)nx;
    float px = lerpf(F.domain_min[0], F.domain_max[0], tx);
    float pdf = pdf_x * pdf_y;
    float w = pdf > 0.0f ? fval / pdf : 0.0f;
    return FilterSample{V2(px, py), w};
}
// filter.jl:926-953
inline FilterSample filter_sample(const HkFilter& F, V2 u) {
    if (F.type == 1) return FilterSample{V2(lerpf(-F.radius[0], F.radius[0], u.x), lerpf(-F.radius[1], F.radius[1], u.y)), 1.0f};
    if (F.type == 2) return FilterSample{V2(sample_tent(u.x, F.radius[0]), sample_tent(u.y, F.radius[1])), 1.0f};
    return filter_sample_tabulated(F, u);
}

// ---------------------------------------------------------------------------------------------
// PerspectiveCamera.generate_ray   src/camera/perspective.jl:95-128
// ---------------------------------------------------------------------------------------------
struct Ray { V3 o, d; float t_max, time; };
inline Ray camera_generate_ray(const HkCamera& C, V2 p_film, V2 lens, float time_u) {
    V3 p_camera = xform_point(C.raster_to_camera, V3(p_film.x, p_film.y, 0.0f));
    V3 o(0.0f);
    V3 d = normalize(p_camera);
    if (C.lens_radius > 0) {
        V2 dl = concentric_sample_disk(lens);
        V2 p_lens(C.lens_radius * dl.x, C.lens_radius * dl.y);
        float t = -C.focal_distance / d.z;
        V3 p_focus = o + d * t;
        o = V3(p_lens.x, p_lens.y, 0.0f);
        d = normalize(p_focus - o);
    }
    float time = lerpf(C.shutter_open, C.shutter_close, time_u);
    Ray r;
    r.o = xform_point(C.camera_to_world, o);
    r.d = normalize(xform_vec(C.camera_to_world, d));
    r.t_max = INF_F;
    r.time = time;
    return r;
}

}  // namespace ok
