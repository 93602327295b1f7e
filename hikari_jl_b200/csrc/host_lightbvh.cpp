// host_lightbvh.cpp — host-side (CPU, scene-build time) construction of the BVH light sampler arrays.
//
// Mirrors what the Julia host does in BVHLightSampler(lights; scene_radius)
// (src/lights/bvh-light-sampler.jl:283-466) with light_bounds() from src/lights/light-bounds.jl:231-295:
// the reference builds this tree on the CPU and uploads it (volpath-state.jl:148-149); the device only
// samples it.  A Julia host would pass its own arrays to hk_upload_lights; this builder exists so the
// C++/Python host mirror can construct the same inputs.
#include "../../include/hikari_cuda.h"
#include <cmath>
#include <vector>
#include <algorithm>
#include <limits>

namespace {

const float PI_F = 3.14159265358979323846f;
struct V3 { float x, y, z; float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); } };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) { float i = 1.0f / norm(a); return {a.x * i, a.y * i, a.z * i}; }
inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct Bounds { V3 lo, hi; };
struct LB { Bounds b; V3 w; float phi, cos_o, cos_e; bool two_sided; };
const float INF = std::numeric_limits<float>::infinity();
LB lb_empty() { return LB{{{INF, INF, INF}, {-INF, -INF, -INF}}, {0, 0, 1}, 0.0f, 1.0f, 1.0f, false}; }
Bounds b_union(Bounds a, Bounds b) {
    return {{std::min(a.lo.x, b.lo.x), std::min(a.lo.y, b.lo.y), std::min(a.lo.z, b.lo.z)}, {std::max(a.hi.x, b.hi.x), std::max(a.hi.y, b.hi.y), std::max(a.hi.z, b.hi.z)}};
}
V3 centroid(const LB& l) { return (l.b.lo + l.b.hi) * 0.5f; }

struct Cone { V3 w; float c; };
float angle_between(V3 a, V3 b) {   // light-bounds.jl:44-50
    if (dot(a, b) < 0.0f) return PI_F - 2.0f * std::asin(clampf(norm(a + b) * 0.5f, -1.0f, 1.0f));
    return 2.0f * std::asin(clampf(norm(b - a) * 0.5f, -1.0f, 1.0f));
}
Cone cone_union(Cone a, Cone b) {   // light-bounds.jl:58-87
    if (a.c == INF) return b;
    if (b.c == INF) return a;
    float ta = std::acos(clampf(a.c, -1.0f, 1.0f)), tb = std::acos(clampf(b.c, -1.0f, 1.0f));
    float td = angle_between(a.w, b.w);
    if (std::min(td + tb, PI_F) <= ta) return a;
    if (std::min(td + ta, PI_F) <= tb) return b;
    float to = (ta + td + tb) * 0.5f;
    if (to >= PI_F) return Cone{{0, 0, 1}, -1.0f};
    float tr = to - ta;
    V3 wr = cross(a.w, b.w);
    if (dot(wr, wr) == 0.0f) return Cone{{0, 0, 1}, -1.0f};
    V3 ax = normalize(wr);
    float s = std::sin(tr), c = std::cos(tr);
    V3 w = a.w * c + cross(ax, a.w) * s + ax * dot(ax, a.w) * (1.0f - c);
    return Cone{normalize(w), std::cos(to)};
}
LB lb_union(const LB& a, const LB& b) {   // light-bounds.jl:142-158
    if (a.phi == 0.0f) return b;
    if (b.phi == 0.0f) return a;
    Cone c = cone_union(Cone{a.w, a.cos_o}, Cone{b.w, b.cos_o});
    return LB{b_union(a.b, b.b), c.w, a.phi + b.phi, c.c, std::min(a.cos_e, b.cos_e), a.two_sided || b.two_sided};
}
float sigmoidf(float x) { if (std::isinf(x)) return x > 0 ? 1.0f : 0.0f; return 0.5f + x / (2.0f * std::sqrt(1.0f + x * x)); }
float poly_eval(const float* p, float l) { return sigmoidf(p[0] * l * l + p[1] * l + p[2]); }
float poly_max(const float* p) {   // rgb2spec.jl:39-53
    float r = std::max(poly_eval(p, 360.0f), poly_eval(p, 830.0f));
    if (p[0] != 0) { float lc = -p[1] / (2.0f * p[0]); if (360.0f <= lc && lc <= 830.0f) r = std::max(r, poly_eval(p, lc)); }
    return r;
}
float light_luminance(const HkLight& L) {   // light-sampler.jl:444-456
    if (L.spectrum_kind == HK_SPECTRUM_ILLUMINANT) return L.illum_scale * poly_max(L.poly) * 100.0f;
    return 0.212671f * L.rgb[0] + 0.715160f * L.rgb[1] + 0.072169f * L.rgb[2];
}
// light_bounds(), light-bounds.jl:231-295; returns false for infinite lights
bool light_bounds(const HkLight& L, LB& out) {
    const float cos_pi = (float)std::cos(M_PI), cos_half_pi = (float)std::cos(M_PI / 2);
    if (L.type == HK_LIGHT_POINT) {
        V3 p{L.position[0], L.position[1], L.position[2]};
        out = LB{{p, p}, {0, 0, 1}, 4.0f * PI_F * L.scale * light_luminance(L), cos_pi, cos_half_pi, false};
        return true;
    }
    if (L.type == HK_LIGHT_SPOT) {
        V3 p{L.position[0], L.position[1], L.position[2]};
        // light_to_world * (0,0,1): third column of inverse(world_to_light); for a rigid transform = third row of W2L's 3x3
        V3 w = normalize(V3{L.world_to_light[8], L.world_to_light[9], L.world_to_light[10]});
        float ce = (float)std::cos(std::acos(L.cos_total_width) - std::acos(L.cos_falloff_start));
        if (ce == 1.0f && L.cos_total_width != L.cos_falloff_start) ce = 0.999f;
        out = LB{{p, p}, w, 4.0f * PI_F * L.scale * light_luminance(L), L.cos_falloff_start, ce, false};
        return true;
    }
    if (L.type == HK_LIGHT_DIFFUSE_AREA) {
        V3 v0{L.v[0], L.v[1], L.v[2]}, v1{L.v[3], L.v[4], L.v[5]}, v2{L.v[6], L.v[7], L.v[8]};
        Bounds b{v0, v0}; b = b_union(b, {v1, v1}); b = b_union(b, {v2, v2});
        float sided = L.two_sided ? 2.0f : 1.0f;
        float lum = 0.212671f * L.rgb[0] + 0.715160f * L.rgb[1] + 0.072169f * L.rgb[2];
        float phi = PI_F * sided * L.area * L.scale * lum;
        out = LB{b, {L.normal[0], L.normal[1], L.normal[2]}, phi, 1.0f, cos_half_pi, L.two_sided != 0};
        return true;
    }
    return false;
}
float evaluate_cost(const LB& lb, const Bounds& b, int dim) {   // bvh-light-sampler.jl:242-258
    float to = std::acos(clampf(lb.cos_o, -1.0f, 1.0f)), te = std::acos(clampf(lb.cos_e, -1.0f, 1.0f));
    float tw = std::min(to + te, PI_F);
    float so = std::sqrt(std::max(0.0f, 1.0f - lb.cos_o * lb.cos_o));
    float M = 2.0f * PI_F * (1.0f - lb.cos_o) + PI_F / 2.0f * (2.0f * tw * so - std::cos(to - 2.0f * tw) - 2.0f * to * so + lb.cos_o);
    V3 d = b.hi - b.lo;
    float md = std::max(std::max(d.x, d.y), d.z), dd = d[dim];
    float Kr = dd > 1.0e-10f ? md / dd : md / 1.0e-10f;
    float sa = 2.0f * (d.x * d.y + d.x * d.z + d.y * d.z);
    return lb.phi * M * Kr * sa;
}
HkLightBVHNode make_node(const LB& lb, uint32_t child_or_light, bool leaf) {
    HkLightBVHNode n{};
    n.bounds_min[0] = lb.b.lo.x; n.bounds_min[1] = lb.b.lo.y; n.bounds_min[2] = lb.b.lo.z;
    n.bounds_max[0] = lb.b.hi.x; n.bounds_max[1] = lb.b.hi.y; n.bounds_max[2] = lb.b.hi.z;
    n.w[0] = lb.w.x; n.w[1] = lb.w.y; n.w[2] = lb.w.z;
    n.phi = lb.phi; n.cos_theta_o = lb.cos_o; n.cos_theta_e = lb.cos_e; n.two_sided = lb.two_sided ? 1 : 0;
    n.child1_or_light_idx = child_or_light; n.is_leaf = leaf ? 1 : 0;
    return n;
}
struct Item { int32_t flat; LB lb; };
float offset_dim(const Bounds& cb, V3 c, int dim) {   // Raycore.offset(bounds, p)[dim]
    float o = c[dim] - cb.lo[dim];
    if (cb.hi[dim] > cb.lo[dim]) o /= cb.hi[dim] - cb.lo[dim];
    return o;
}
const int NB = 12;
LB build(std::vector<HkLightBVHNode>& nodes, std::vector<uint32_t>& trail, std::vector<Item>& items, int start, int stop, uint32_t bits, int depth) {
    int count = stop - start + 1;
    if (count == 1) {
        nodes.push_back(make_node(items[start].lb, (uint32_t)items[start].flat, true));
        trail[items[start].flat - 1] = bits;
        return items[start].lb;
    }
    LB overall = items[start].lb;
    V3 c0 = centroid(items[start].lb);
    Bounds cb{c0, c0};
    for (int i = start + 1; i <= stop; i++) { overall = lb_union(overall, items[i].lb); V3 c = centroid(items[i].lb); cb = b_union(cb, {c, c}); }
    float best_cost = INF; int best_dim = -1, best_bucket = 0;
    for (int dim = 0; dim < 3; dim++) {
        if (cb.hi[dim] - cb.lo[dim] <= 0.0f) continue;
        LB bb[NB]; int bc[NB];
        for (int b = 0; b < NB; b++) { bb[b] = lb_empty(); bc[b] = 0; }
        for (int i = start; i <= stop; i++) {
            int b = (int)std::floor(NB * offset_dim(cb, centroid(items[i].lb), dim));
            b = std::min(std::max(b, 0), NB - 1);
            bb[b] = lb_union(bb[b], items[i].lb); bc[b]++;
        }
        for (int split = 1; split < NB; split++) {
            LB below = lb_empty(), above = lb_empty(); int nb = 0, na = 0;
            for (int b = 0; b < split; b++) { below = lb_union(below, bb[b]); nb += bc[b]; }
            for (int b = split; b < NB; b++) { above = lb_union(above, bb[b]); na += bc[b]; }
            if (nb == 0 || na == 0) continue;
            float cost = evaluate_cost(below, overall.b, dim) + evaluate_cost(above, overall.b, dim);
            if (cost < best_cost) { best_cost = cost; best_dim = dim; best_bucket = split; }
        }
    }
    int mid;
    if (best_dim >= 0) {
        int pivot = start;
        for (int i = start; i <= stop; i++) {
            int b = (int)std::floor(NB * offset_dim(cb, centroid(items[i].lb), best_dim));
            b = std::min(std::max(b, 0), NB - 1) + 1;
            if (b <= best_bucket) { if (i != pivot) std::swap(items[pivot], items[i]); pivot++; }
        }
        mid = (pivot == start || pivot > stop) ? start + count / 2 : pivot - 1;
    } else mid = start + count / 2 - 1;
    mid = std::min(std::max(mid, start), stop - 1);
    size_t me = nodes.size();
    nodes.push_back(make_node(overall, 0, false));
    LB l0 = build(nodes, trail, items, start, mid, bits, depth + 1);
    uint32_t child1 = (uint32_t)nodes.size() + 1;   // 1-based
    LB l1 = build(nodes, trail, items, mid + 1, stop, bits | (1u << depth), depth + 1);
    LB merged = lb_union(l0, l1);
    nodes[me] = make_node(merged, child1, false);
    return merged;
}

}  // namespace

// Builds the sampler arrays. out_nodes must hold 2*n_lights entries, out_trails n_lights, out_infinite n_lights.
extern "C" int32_t hk_host_build_light_sampler(const HkLight* lights, uint32_t n_lights, HkLightBVHNode* out_nodes, uint32_t* out_n_nodes,
                                               uint32_t* out_trails, int32_t* out_infinite, uint32_t* out_n_infinite, uint32_t* out_n_bvh) {
    std::vector<Item> items; std::vector<int32_t> inf;
    for (uint32_t i = 0; i < n_lights; i++) {
        LB lb;
        if (!light_bounds(lights[i], lb)) inf.push_back((int32_t)i + 1);
        else if (lb.phi > 0.0f) items.push_back(Item{(int32_t)i + 1, lb});
    }
    std::vector<uint32_t> trail(n_lights, 0xFFFFFFFFu);
    std::vector<HkLightBVHNode> nodes;
    if (!items.empty()) build(nodes, trail, items, 0, (int)items.size() - 1, 0u, 0);
    for (size_t i = 0; i < nodes.size(); i++) out_nodes[i] = nodes[i];
    for (uint32_t i = 0; i < n_lights; i++) out_trails[i] = trail[i];
    for (size_t i = 0; i < inf.size(); i++) out_infinite[i] = inf[i];
    *out_n_nodes = (uint32_t)nodes.size(); *out_n_infinite = (uint32_t)inf.size(); *out_n_bvh = (uint32_t)items.size();
    return 0;
}
