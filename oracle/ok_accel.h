// ok_accel.h — ORACLE (test infrastructure, NOT product code).
// Closest-hit over a triangle soup: brute force and a plain BVH2, both implementing the STATED
// tie-break rule (SURVEY §8c / DESIGN.md): argmin t over triangles with 0 < t < t_max using the fixed
// Möller–Trumbore sequence below evaluated in f32 without FMA contraction; equal t -> smallest
// global primitive id.  Raycore.jl (the reference's real traversal) is not on disk: parity of hit ids
// against upstream is UNPINNED; this file is the definition both implementations are held to.
// Instanced scenes (InstancedAccel below, HkGeometry.instances): the ray goes to object space as o' = W o, d' = W d (W =
// world_to_object, d' not renormalised: t means the same on both sides), is tested against the mesh's OBJECT-space triangles with the
// same test, and the closest hit is argmin t over all (instance, face) with equal t -> smallest GLOBAL primitive id
// (instance-major: id = faces of the earlier instances + face index).
#pragma once
#include "ok_core.h"
#include <vector>
#include <numeric>
#include <array>
#include "../include/hikari_cuda.h"

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

// ---- instances: one Accel per mesh + a linear, conservatively culled loop over the instances ------------------------------------
inline V3 inst_point(const float* m, V3 p) { return V3(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7], m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]); }
inline V3 inst_vector(const float* m, V3 v) { return V3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z); }
inline V3 inst_normal(const float* w2o, V3 n) {      // normalize(W^T n): the inverse-transpose of the object-to-world linear part
    return normalize(V3(w2o[0] * n.x + w2o[4] * n.y + w2o[8] * n.z, w2o[1] * n.x + w2o[5] * n.y + w2o[9] * n.z, w2o[2] * n.x + w2o[6] * n.y + w2o[10] * n.z));
}
struct Instance { uint32_t mesh, iface, prim_base, n_tris, first_tri; float o2w[12], w2o[12]; float wmin[3], wmax[3]; };
struct InstancedAccel {
    std::vector<Accel> meshes;
    std::vector<Instance> inst;
    std::vector<uint32_t> prim_base;          // per instance, ascending: global id of its first face
    bool enabled() const { return !inst.empty(); }
    void build(const float* pos, const uint32_t* idx, const HkMesh* ms, uint32_t n_meshes, const HkInstance* is, uint32_t n_inst) {
        meshes.assign(n_meshes, Accel()); inst.clear(); prim_base.clear();
        std::vector<std::array<float, 6>> obox(n_meshes);
        for (uint32_t m = 0; m < n_meshes; m++) {
            meshes[m].build(pos, idx + 3 * (size_t)ms[m].first_tri, ms[m].n_tris);
            std::array<float, 6> b = {INF_F, INF_F, INF_F, -INF_F, -INF_F, -INF_F};
            for (uint32_t t = 0; t < 3 * ms[m].n_tris; t++) for (int k = 0; k < 3; k++) {
                const float v = pos[3 * (size_t)idx[3 * (size_t)ms[m].first_tri + t] + k];
                b[k] = std::min(b[k], v); b[3 + k] = std::max(b[3 + k], v);
            }
            obox[m] = b;
        }
        uint32_t base = 0;
        for (uint32_t i = 0; i < n_inst; i++) {
            Instance I; I.mesh = is[i].mesh; I.iface = is[i].medium_interface_idx; I.prim_base = base; I.n_tris = ms[I.mesh].n_tris; I.first_tri = ms[I.mesh].first_tri;
            std::memcpy(I.o2w, is[i].object_to_world, 48); std::memcpy(I.w2o, is[i].world_to_object, 48);
            for (int k = 0; k < 3; k++) { I.wmin[k] = INF_F; I.wmax[k] = -INF_F; }
            const auto& b = obox[I.mesh];
            for (int c = 0; c < 8; c++) {           // world box of the 8 transformed corners, padded: culling must only ever be conservative
                V3 p = inst_point(I.o2w, V3(b[(c & 1) ? 3 : 0], b[(c & 2) ? 4 : 1], b[(c & 4) ? 5 : 2]));
                for (int k = 0; k < 3; k++) { I.wmin[k] = std::min(I.wmin[k], p[k]); I.wmax[k] = std::max(I.wmax[k], p[k]); }
            }
            for (int k = 0; k < 3; k++) { float pad = 1.0e-4f * std::max(std::fabs(I.wmin[k]), std::fabs(I.wmax[k])) + 1.0e-5f; I.wmin[k] -= pad; I.wmax[k] += pad; }
            inst.push_back(I); prim_base.push_back(base);
            base += I.n_tris;
        }
    }
    uint32_t n_prims() const { return inst.empty() ? 0u : inst.back().prim_base + inst.back().n_tris; }
    uint32_t instance_of(uint32_t prim) const { return (uint32_t)(std::upper_bound(prim_base.begin(), prim_base.end(), prim) - prim_base.begin()) - 1u; }
    Hit closest_hit(V3 o, V3 d, float t_max, bool brute) const {
        Hit best{false, 0, t_max, 0, 0};
        const V3 inv(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        for (const Instance& I : inst) {
            if (!brute) {
                Accel::Node N; for (int k = 0; k < 3; k++) { N.bmin[k] = I.wmin[k]; N.bmax[k] = I.wmax[k]; }
                float tn;
                if (!Accel::slab(N, o, inv, t_max, tn)) continue;
            }
            const V3 oo = inst_point(I.w2o, o), od = inst_vector(I.w2o, d);
            const Hit h = brute ? meshes[I.mesh].closest_hit_brute(oo, od, t_max) : meshes[I.mesh].closest_hit_bvh(oo, od, t_max);
            if (h.hit && hit_better(h.t, I.prim_base + h.prim, best)) best = Hit{true, I.prim_base + h.prim, h.t, h.b1, h.b2};
        }
        return best;
    }
};

}  // namespace ok
