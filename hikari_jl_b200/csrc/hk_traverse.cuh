// hk_traverse.cuh — closest-hit / any-hit traversal of the 80-byte BVH8 (hk_bvh.h).
// Replaces Raycore.closest_hit(accel, ray) (call sites src/integrators/volpath/intersection.jl:200,225,323,703).
//
// Contract (DESIGN.md "closest hit"): the hit is argmin t over triangles with 0 < t < t_max under the fixed
// Moller-Trumbore sequence in tri_test() (same operation order as oracle/ok_accel.h, no FMA contraction);
// equal t -> smallest global primitive id.  Nodes are culled only when t_near > t_best, so ties survive.
//
// Per-thread traversal keeps a short stack of (node-group, triangle-group) pairs: first HK_SM_STACK entries in
// shared memory (one column per thread, bank-conflict free), the rest in local memory.
#pragma once
#include "hk_math.cuh"
#include "hk_bvh.h"

#define HK_SM_STACK 8
#define HK_LM_STACK 24
#define HK_TRACE_THREADS 128

struct DevBvh { const float4* __restrict__ nodes; const float4* __restrict__ tris; };
struct HitRec { float t; uint32_t prim1; float b1, b2; };   // prim1: 1-based global id, 0 = miss

HK_DEV bool tri_test(float3 o, float3 d, float3 v0, float3 e1, float3 e2, float t_max, float& t, float& u, float& v) {
    float3 pvec = cross3(d, e2);
    float det = dot3(e1, pvec);
    if (det == 0.0f) return false;
    float inv_det = 1.0f / det;
    float3 tvec = o - v0;
    u = dot3(tvec, pvec) * inv_det;
    if (u < 0.0f || u > 1.0f) return false;
    float3 qvec = cross3(tvec, e1);
    v = dot3(d, qvec) * inv_det;
    if (v < 0.0f || u + v > 1.0f) return false;
    t = dot3(e2, qvec) * inv_det;
    return t > 0.0f && t < t_max;
}

struct TravStack {
    uint2* sm;          // shared column base for this thread (stride HK_TRACE_THREADS)
    uint2 lm[HK_LM_STACK];
    int n;
    HK_DEV void push(uint2 v) { if (n < HK_SM_STACK) sm[n * HK_TRACE_THREADS] = v; else lm[n - HK_SM_STACK] = v; n++; }
    HK_DEV uint2 pop() { n--; return n < HK_SM_STACK ? sm[n * HK_TRACE_THREADS] : lm[n - HK_SM_STACK]; }
};

// COUNT: accumulate node visits / triangle tests (roofline accounting).  ANY: stop at the first accepted hit.
template <bool ANY, bool COUNT>
HK_DEV HitRec bvh8_trace(const DevBvh& B, uint2* sm_stack, float3 o, float3 d, float t_max, uint32_t* n_nodes = nullptr, uint32_t* n_tris = nullptr) {
    HitRec best; best.t = t_max; best.prim1 = 0; best.b1 = 0.0f; best.b2 = 0.0f;
    const float3 inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    // octant: bit set <=> direction component is non-negative (near children then sit at the - side)
    const uint32_t oct_inv = (d.x >= 0.0f ? 4u : 0u) | (d.y >= 0.0f ? 2u : 0u) | (d.z >= 0.0f ? 1u : 0u);
    TravStack st; st.sm = sm_stack; st.n = 0;
    uint2 ngroup = make_uint2(0u, 0x80000000u);   // root: node base 0, "child bit 31" set, imask irrelevant (relative index 0)
    uint2 tgroup = make_uint2(0u, 0u);
    bool root = true;
    for (;;) {
        if (ngroup.y > 0x00FFFFFFu) {
            // ---- pop the nearest pending internal child of this group -------------------------------------
            uint32_t hits = ngroup.y;
            uint32_t bit = 31u - (uint32_t)__clz(hits);
            ngroup.y &= ~(1u << bit);
            if (ngroup.y > 0x00FFFFFFu) st.push(ngroup);
            uint32_t node_idx;
            if (root) { node_idx = 0; root = false; }
            else {
                uint32_t slot = (bit - 24u) ^ oct_inv;
                uint32_t imask = hits & 0xFFu;
                node_idx = ngroup.x + (uint32_t)__popc(imask & ((1u << slot) - 1u));
            }
            // ---- fetch the 80-byte node as five 16-byte loads --------------------------------------------
            const float4* np = B.nodes + (size_t)node_idx * 5;
            float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            if (COUNT) (*n_nodes)++;
            uint32_t ex = __float_as_uint(n0.w);
            float sx = __uint_as_float((ex & 0xFFu) << 23), sy = __uint_as_float(((ex >> 8) & 0xFFu) << 23), sz = __uint_as_float(((ex >> 16) & 0xFFu) << 23);
            uint32_t imask = ex >> 24;
            uint32_t child_base = __float_as_uint(n1.x), tri_base = __float_as_uint(n1.y);
            uint32_t meta_lo = __float_as_uint(n1.z), meta_hi = __float_as_uint(n1.w);
            // quantised planes: qlo x/y/z = n2.xy, n2.zw, n3.xy ; qhi x/y/z = n3.zw, n4.xy, n4.zw
            uint32_t q[12] = {__float_as_uint(n2.x), __float_as_uint(n2.y), __float_as_uint(n2.z), __float_as_uint(n2.w),
                              __float_as_uint(n3.x), __float_as_uint(n3.y), __float_as_uint(n3.z), __float_as_uint(n3.w),
                              __float_as_uint(n4.x), __float_as_uint(n4.y), __float_as_uint(n4.z), __float_as_uint(n4.w)};
            const float ax = sx * inv.x, ay = sy * inv.y, az = sz * inv.z;
            const float bx = (n0.x - o.x) * inv.x, by = (n0.y - o.y) * inv.y, bz = (n0.z - o.z) * inv.z;
            const float tlim = best.t;
            uint32_t hitmask = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t meta = ((i < 4 ? meta_lo : meta_hi) >> (8 * (i & 3))) & 0xFFu;
                if (meta == 0) continue;
                const int w = i >> 2, sh = 8 * (i & 3);
                float lox = (float)((q[0 + w] >> sh) & 0xFFu), loy = (float)((q[2 + w] >> sh) & 0xFFu), loz = (float)((q[4 + w] >> sh) & 0xFFu);
                float hix = (float)((q[6 + w] >> sh) & 0xFFu), hiy = (float)((q[8 + w] >> sh) & 0xFFu), hiz = (float)((q[10 + w] >> sh) & 0xFFu);
                float tx0 = (d.x >= 0.0f ? lox : hix) * ax + bx, tx1 = (d.x >= 0.0f ? hix : lox) * ax + bx;
                float ty0 = (d.y >= 0.0f ? loy : hiy) * ay + by, ty1 = (d.y >= 0.0f ? hiy : loy) * ay + by;
                float tz0 = (d.z >= 0.0f ? loz : hiz) * az + bz, tz1 = (d.z >= 0.0f ? hiz : loz) * az + bz;
                // fmaxf/fminf drop NaNs (0*inf when the origin lies in a slab plane of a zero direction): conservative
                float tn = fmaxf(fmaxf(tx0, ty0), fmaxf(tz0, 0.0f));
                float tf = fminf(fminf(tx1, ty1), fminf(tz1, tlim)) * 1.0000004f;
                if (tn <= tf) {
                    uint32_t inner = (meta & (meta << 1)) & 0x10u;             // bits 3 and 4 both set <=> internal child
                    uint32_t bit_index = (meta ^ (inner ? oct_inv : 0u)) & 0x1Fu;
                    hitmask |= (meta >> 5) << bit_index;
                }
            }
            ngroup = make_uint2(child_base, (hitmask & 0xFF000000u) | imask);
            tgroup = make_uint2(tri_base, hitmask & 0x00FFFFFFu);
        } else {
            tgroup = ngroup;
            ngroup = make_uint2(0u, 0u);
        }
        // ---- triangles of this node ----------------------------------------------------------------------
        while (tgroup.y != 0u) {
            uint32_t ti = (uint32_t)__ffs(tgroup.y) - 1u;
            tgroup.y &= tgroup.y - 1u;
            const float4* tp = B.tris + (size_t)(tgroup.x + ti) * 3;
            float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
            if (COUNT) (*n_tris)++;
            float t, u, v;
            if (tri_test(o, d, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), t_max, t, u, v)) {
                uint32_t prim1 = __float_as_uint(a.w) + 1u;
                if (ANY) { best.t = t; best.prim1 = prim1; best.b1 = u; best.b2 = v; return best; }
                if (best.prim1 == 0u || t < best.t || (t == best.t && prim1 < best.prim1)) { best.t = t; best.prim1 = prim1; best.b1 = u; best.b2 = v; }
            }
        }
        if (ngroup.y <= 0x00FFFFFFu) {
            if (st.n == 0) break;
            ngroup = st.pop();
        }
    }
    return best;
}
