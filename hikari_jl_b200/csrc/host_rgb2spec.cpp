// host_rgb2spec.cpp — host-side (CPU, one-time) generator of the sRGB -> sigmoid-polynomial table.
//
// The reference ships this table as src/spectral/srgb_spectrum_table.dat, which is absent from the
// checkout (.MISSING_LARGE_BLOBS), and regenerates it with src/spectral/rgb2spec_gen.jl:332-409
// (a Julia port of pbrt-v4's rgb2spec_opt).  This is the same Gauss-Newton fit in Float64, run once at
// build time; the result is an INPUT of the render path (hk_upload_tables), not part of it.
// Output layout = the reference's file format (rgb2spec.jl:403-412): scale[res], coeffs as the raw memory
// of Array{Float32,5}(3,res,res,res,3) indexed [maxc, z, y, x, coef] column-major.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace {

constexpr int kCie = 95;
constexpr double kLamMin = 360.0, kLamMax = 830.0;
constexpr int kFine = (kCie - 1) * 3 + 1;
const double XYZ_TO_SRGB[3][3] = {{3.240479, -1.537150, -0.498535}, {-0.969256, 1.875991, 0.041556}, {0.055648, -0.204043, 1.057311}};
const double SRGB_TO_XYZ[3][3] = {{0.412453, 0.357580, 0.180423}, {0.212671, 0.715160, 0.072169}, {0.019334, 0.119193, 0.950227}};

struct Tab { double lambda[kFine]; double w[3][kFine]; double white[3]; };

double interp(const double* d, double lambda) {   // rgb2spec_gen.jl:126-133
    double x = (lambda - kLamMin) * ((kCie - 1) / (kLamMax - kLamMin));
    int o = (int)std::floor(x);
    o = std::min(std::max(o, 0), kCie - 2);
    double wt = x - o;
    return (1.0 - wt) * d[o] + wt * d[o + 1];
}
void init_tables(Tab& T, const double* cx, const double* cy, const double* cz, const double* d65) {   // :171-212
    double h = (kLamMax - kLamMin) / (kFine - 1);
    std::memset(&T, 0, sizeof(T));
    for (int i = 0; i < kFine; i++) {
        double lam = kLamMin + i * h;
        T.lambda[i] = lam;
        double xyz[3] = {interp(cx, lam), interp(cy, lam), interp(cz, lam)};
        double I = interp(d65, lam);
        double weight = 3.0 / 8.0 * h;
        if (i == 0 || i == kFine - 1) {}
        else if ((i - 1) % 3 == 2) weight *= 2.0;
        else weight *= 3.0;
        for (int k = 0; k < 3; k++) for (int j = 0; j < 3; j++) T.w[k][i] += XYZ_TO_SRGB[k][j] * xyz[j] * I * weight;
        for (int k = 0; k < 3; k++) T.white[k] += xyz[k] * I * weight;
    }
}
double lab_f(double t) { const double d = 6.0 / 29.0; return t > d * d * d ? std::cbrt(t) : t / (3.0 * d * d) + 4.0 / 29.0; }
void rgb_to_lab(const double* rgb, const Tab& T, double* lab) {   // :142-156
    double xyz[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) xyz[i] += SRGB_TO_XYZ[i][j] * rgb[j];
    double fx = lab_f(xyz[0] / T.white[0]), fy = lab_f(xyz[1] / T.white[1]), fz = lab_f(xyz[2] / T.white[2]);
    lab[0] = 116.0 * fy - 16.0; lab[1] = 500.0 * (fx - fy); lab[2] = 200.0 * (fy - fz);
}
void residual(const double* c, const double* target, const Tab& T, double* r) {   // :221-248
    double out[3] = {0, 0, 0};
    for (int i = 0; i < kFine; i++) {
        double ln = (T.lambda[i] - kLamMin) / (kLamMax - kLamMin);
        double x = c[0] * ln * ln + c[1] * ln + c[2];
        double s = 0.5 * x / std::sqrt(1.0 + x * x) + 0.5;
        for (int j = 0; j < 3; j++) out[j] += T.w[j][i] * s;
    }
    double a[3], b[3];
    rgb_to_lab(out, T, a); rgb_to_lab(target, T, b);
    for (int j = 0; j < 3; j++) r[j] = b[j] - a[j];
}
bool solve3(double A[3][3], const double* b, double* x) {   // LU with partial pivoting (Julia's `\`)
    double M[3][4];
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) M[i][j] = A[i][j]; M[i][3] = b[i]; }
    for (int k = 0; k < 3; k++) {
        int p = k;
        for (int i = k + 1; i < 3; i++) if (std::fabs(M[i][k]) > std::fabs(M[p][k])) p = i;
        if (M[p][k] == 0.0 || !std::isfinite(M[p][k])) return false;
        if (p != k) for (int j = 0; j < 4; j++) std::swap(M[p][j], M[k][j]);
        for (int i = k + 1; i < 3; i++) { double f = M[i][k] / M[k][k]; for (int j = k; j < 4; j++) M[i][j] -= f * M[k][j]; }
    }
    for (int i = 2; i >= 0; i--) { double s = M[i][3]; for (int j = i + 1; j < 3; j++) s -= M[i][j] * x[j]; x[i] = s / M[i][i]; }
    return std::isfinite(x[0]) && std::isfinite(x[1]) && std::isfinite(x[2]);
}
void gauss_newton(double* c, const double* target, const Tab& T) {   // :274-305
    const double eps = 1e-4;
    for (int it = 0; it < 15; it++) {
        double r[3]; residual(c, target, T, r);
        double J[3][3];
        for (int i = 0; i < 3; i++) {
            double tmp[3] = {c[0], c[1], c[2]}, r0[3], r1[3];
            tmp[i] = c[i] - eps; residual(tmp, target, T, r0);
            tmp[i] = c[i] + eps; residual(tmp, target, T, r1);
            for (int j = 0; j < 3; j++) J[j][i] = (r1[j] - r0[j]) / (2 * eps);
        }
        double x[3];
        if (!solve3(J, r, x)) break;
        for (int i = 0; i < 3; i++) c[i] -= x[i];
        double mx = std::max(std::max(std::fabs(c[0]), std::fabs(c[1])), std::fabs(c[2]));
        if (mx > 200.0) for (int i = 0; i < 3; i++) c[i] *= 200.0 / mx;
        double rr = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        if (rr < 1e-6) break;
    }
}
double smoothstep(double x) { return x * x * (3.0 - 2.0 * x); }

}  // namespace

extern "C" int32_t hk_host_generate_rgb2spec(int32_t res, const double* cie_x, const double* cie_y, const double* cie_z,
                                              const double* d65_normalised, float* out_scale, float* out_coeffs) {
    if (res < 2) return -1;
    static Tab T;
    init_tables(T, cie_x, cie_y, cie_z, d65_normalised);
    for (int k = 0; k < res; k++) out_scale[k] = (float)smoothstep(smoothstep((double)k / (res - 1)));
    const size_t R = (size_t)res;
    auto store = [&](int l, int k, int j, int i, const double* oc) {   // 0-based l,k,j,i
        const double c0 = 360.0, c1 = 1.0 / (830.0 - 360.0);
        double A = oc[0], B = oc[1], C = oc[2];
        float v[3] = {(float)(A * c1 * c1), (float)(B * c1 - 2 * A * c0 * c1 * c1), (float)(C - B * c0 * c1 + A * (c0 * c1) * (c0 * c1))};
        for (int c = 0; c < 3; c++) out_coeffs[(size_t)l + 3 * ((size_t)k + R * ((size_t)j + R * ((size_t)i + R * (size_t)c)))] = v[c];
    };
    for (int l = 0; l < 3; l++) {
        #pragma omp parallel for schedule(dynamic, 1)
        for (int j = 0; j < res; j++) {
            double y = (double)j / (res - 1);
            for (int i = 0; i < res; i++) {
                double x = (double)i / (res - 1);
                int start_k = res / 5;
                double oc[3] = {0, 0, 0}, rgb[3];
                for (int k = start_k; k < res; k++) {       // Julia (start_k+1):res, 1-based
                    double b = (double)out_scale[k];
                    rgb[l] = b; rgb[(l + 1) % 3] = x * b; rgb[(l + 2) % 3] = y * b;
                    gauss_newton(oc, rgb, T);
                    store(l, k, j, i, oc);
                }
                oc[0] = oc[1] = oc[2] = 0;
                for (int k = start_k; k >= 0; k--) {        // Julia (start_k+1):-1:1
                    double b = (double)out_scale[k];
                    rgb[l] = b; rgb[(l + 1) % 3] = x * b; rgb[(l + 2) % 3] = y * b;
                    gauss_newton(oc, rgb, T);
                    store(l, k, j, i, oc);
                }
            }
        }
    }
    return 0;
}
