// ok_accel.h — ORACLE (test infrastructure, NOT product code).
// Closest-hit over a triangle soup: brute force and a plain BVH2, both implementing the STATED
// tie-break rule (SURVEY §8c / DESIGN.md): argmin t over triangles with 0 < t < t_max using the fixed
// Möller–Trumbore sequence below evaluated in f32 without FMA contraction; equal t -> smallest
// global primitive id.  Raycore.jl (the reference's real traversal) is not on disk: parity of hit ids
// against upstream is UNPINNED; this file is the definition both implementations are held to.
#pragma once
#include "ok_core.h"
#include <vector>
#include <numeric>

namespace ok {

struct Tri { V3 v0, e1, e2; };
struct Hit { bool hit; uint32_t prim; float t, b1, b2; };   // prim 0-based here

// The normative ray/triangle test.  Operation order is part of the contract.
inline bool tri_intersect(V3 o, V3 d, const Tri& T, float t_max, float& t, float& u, float& v) {
    V3 pvec = cross(d, T.e2);
    float det = dot(T.e1, pvec);
    if (det == 0.0f) return false;
    float inv_det = 1.0f / det;
    V3 tvec = o - T.v0;
    u = dot(tvec, pvec) * inv_det;
    if (u < 0.0f || u > 1.0f) return false;
    V3 qvec = cross(tvec, T.e1);
    v = dot(d, qvec) * inv_det;
    if (v < 0.0f || u + v > 1.0f) return false;
    t = dot(T.e2, qvec) * inv_det;
    return t > 0.0f && t < t_max;
}
inline bool hit_better(float t, uint32_t prim, const Hit& best) {
    return !best.hit || t < best.t || (t == best.t && prim < best.prim);
}

struct Accel {
    std::vector<Tri> tris;
    struct Node { float bmin[3], bmax[3]; uint32_t left, right, first, count; };   // count>0 => leaf
    std::vector<Node> nodes;
    std::vector<uint32_t> order;

    void build(const float* pos, const uint32_t* idx, uint32_t n_tris) {
        tris.resize(n_tris);
        for (uint32_t i = 0; i < n_tris; i++) {
            const float* a = pos + 3 * (size_t)idx[3 * i], *b = pos + 3 * (size_t)idx[3 * i + 1], *c = pos + 3 * (size_t)idx[3 * i + 2];
            V3 v0(a[0], a[1], a[2]), v1(b[0], b[1], b[2]), v2(c[0], c[1], c[2]);
            tris[i] = Tri{v0, v1 - v0, v2 - v0};
        }
        order.resize(n_tris);
        std::iota(order.begin(), order.end(), 0u);
        nodes.clear();
        if (n_tris == 0) return;
        std::vector<float> cb(6 * (size_t)n_tris), cen(3 * (size_t)n_tris);
        for (uint32_t i = 0; i < n_tris; i++) {
            V3 p[3] = {tris[i].v0, tris[i].v0 + tris[i].e1, tris[i].v0 + tris[i].e2};
            for (int k = 0; k < 3; k++) {
                float lo = std::min(std::min(p[0][k], p[1][k]), p[2][k]), hi = std::max(std::max(p[0][k], p[1][k]), p[2][k]);
                cb[6 * (size_t)i + k] = lo; cb[6 * (size_t)i + 3 + k] = hi; cen[3 * (size_t)i + k] = 0.5f * (lo + hi);
            }
        }
        nodes.reserve(2 * (size_t)n_tris);
        build_rec(0, n_tris, cb, cen);
    }
    uint32_t build_rec(uint32_t first, uint32_t count, const std::vector<float>& cb, const std::vector<float>& cen) {
        uint32_t me = (uint32_t)nodes.size();
        nodes.push_back(Node());
        float bmin[3] = {INF_F, INF_F, INF_F}, bmax[3] = {-INF_F, -INF_F, -INF_F}, cmin[3] = {INF_F, INF_F, INF_F}, cmax[3] = {-INF_F, -INF_F, -INF_F};
        for (uint32_t i = first; i < first + count; i++) {
            uint32_t t = order[i];
            for (int k = 0; k < 3; k++) {
                bmin[k] = std::min(bmin[k], cb[6 * (size_t)t + k]); bmax[k] = std::max(bmax[k], cb[6 * (size_t)t + 3 + k]);
                cmin[k] = std::min(cmin[k], cen[3 * (size_t)t + k]); cmax[k] = std::max(cmax[k], cen[3 * (size_t)t + k]);
            }
        }
        for (int k = 0; k < 3; k++) {   // conservative pad: the triangle test is not exact w.r.t. the box
            float pad = 1.0e-5f * std::max(std::fabs(bmin[k]), std::fabs(bmax[k])) + 1.0e-6f;
            nodes[me].bmin[k] = bmin[k] - pad; nodes[me].bmax[k] = bmax[k] + pad;
        }
        int axis = 0; float ext = cmax[0] - cmin[0];
        for (int k = 1; k < 3; k++) if (cmax[k] - cmin[k] > ext) { ext = cmax[k] - cmin[k]; axis = k; }
        if (count <= 4 || ext <= 0.0f) { nodes[me].first = first; nodes[me].count = count; nodes[me].left = nodes[me].right = 0; return me; }
        uint32_t mid = first + count / 2;
        std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + first + count,
                         [&](uint32_t a, uint32_t b) { float ca = cen[3 * (size_t)a + axis], cbv = cen[3 * (size_t)b + axis]; return ca < cbv || (ca == cbv && a < b); });
        nodes[me].count = 0; nodes[me].first = 0;
        uint32_t l = build_rec(first, mid - first, cb, cen);
        uint32_t r = build_rec(mid, first + count - mid, cb, cen);
        nodes[me].left = l; nodes[me].right = r;
        return me;
    }
    static inline bool slab(const Node& N, V3 o, V3 inv, float tbest, float& tnear) {
        float t0 = 0.0f, t1 = tbest;
        for (int k = 0; k < 3; k++) {
            float ta = (N.bmin[k] - o[k]) * inv[k], tb = (N.bmax[k] - o[k]) * inv[k];
            if (ta > tb) std::swap(ta, tb);
            tb *= 1.0000004f;   // 1 + 2*gamma(3)
            if (ta != ta || tb != tb) {   // NaN from 0*inf: origin on a slab plane of a zero-direction axis
                if (o[k] < N.bmin[k] || o[k] > N.bmax[k]) return false;
                continue;
            }
            t0 = ta > t0 ? ta : t0;
            t1 = tb < t1 ? tb : t1;
            if (t0 > t1) return false;
        }
        tnear = t0;
        return true;
    }
    Hit closest_hit_bvh(V3 o, V3 d, float t_max) const {
        Hit best{false, 0, t_max, 0, 0};
        if (nodes.empty()) return best;
        V3 inv(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        uint32_t stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const Node& N = nodes[stack[--sp]];
            float tn;
            if (!slab(N, o, inv, best.hit ? best.t : t_max, tn)) continue;
            if (N.count) {
                for (uint32_t i = N.first; i < N.first + N.count; i++) {
                    uint32_t p = order[i]; float t, u, v;
                    if (tri_intersect(o, d, tris[p], t_max, t, u, v) && hit_better(t, p, best)) best = Hit{true, p, t, u, v};
                }
            } else { stack[sp++] = N.left; stack[sp++] = N.right; }
        }
        return best;
    }
    Hit closest_hit_brute(V3 o, V3 d, float t_max) const {
        Hit best{false, 0, t_max, 0, 0};
        for (uint32_t p = 0; p < tris.size(); p++) {
            float t, u, v;
            if (tri_intersect(o, d, tris[p], t_max, t, u, v) && hit_better(t, p, best)) best = Hit{true, p, t, u, v};
        }
        return best;
    }
};

}  // namespace ok
