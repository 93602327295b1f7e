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

// Traversal state machine.  One call of step() either expands the nearest pending internal node of the current node
// group (one 80-byte node fetch + 8 slab tests) or hands the group to the triangle loop, tests all pending triangles and
// pops the next group.  Persistent kernels interleave step() with per-lane refill so that a finished lane takes a new ray
// instead of idling until the slowest ray of its warp is done (hk_wavefront.cuh::trace_queue).
// COUNT: accumulate node visits / triangle tests (roofline accounting).  ANY: stop at the first accepted hit.
template <bool ANY, bool COUNT>
struct Bvh8Walker {
    float3 o, d, inv;
    float t_max;
    uint32_t oct_inv;
    uint2 ngroup, tgroup;
    TravStack st;
    HitRec best;
    bool root;

    HK_DEV void begin(uint2* sm_stack, float3 o_, float3 d_, float t_max_) {
        o = o_; d = d_; t_max = t_max_;
        inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        // octant: bit set <=> direction component is non-negative (near children then sit at the - side)
        oct_inv = (d.x >= 0.0f ? 4u : 0u) | (d.y >= 0.0f ? 2u : 0u) | (d.z >= 0.0f ? 1u : 0u);
        st.sm = sm_stack; st.n = 0;
        ngroup = make_uint2(0u, 0x80000000u);   // root: node base 0, "child bit 31" set
        tgroup = make_uint2(0u, 0u);
        best.t = t_max; best.prim1 = 0; best.b1 = 0.0f; best.b2 = 0.0f;
        root = true;
    }
    // returns true when the traversal is finished
    HK_DEV bool step(const DevBvh& B, uint32_t* n_nodes, uint32_t* n_tris) {
        if (ngroup.y > 0x00FFFFFFu) {
            // ---- pop the nearest pending internal child of this group -------------------------------------
            const uint32_t hits = ngroup.y;
            const uint32_t bit = 31u - (uint32_t)__clz(hits);
            ngroup.y &= ~(1u << bit);
            if (ngroup.y > 0x00FFFFFFu) st.push(ngroup);
            uint32_t node_idx;
            if (root) { node_idx = 0; root = false; }
            else {
                const uint32_t slot = (bit - 24u) ^ oct_inv;
                node_idx = ngroup.x + (uint32_t)__popc((hits & 0xFFu) & ((1u << slot) - 1u));
            }
            // ---- fetch the 80-byte node as five 16-byte loads --------------------------------------------
            const float4* np = B.nodes + (size_t)node_idx * 5;
            const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            if (COUNT) (*n_nodes)++;
            const uint32_t ex = __float_as_uint(n0.w);
            // slab coefficients: t = q * (2^e / d) + (p - o) / d.  Node culling only has to be conservative, so FMA is fine
            // here (the exact, contraction-free arithmetic is reserved for the triangle test).
            const float ax = __uint_as_float((ex & 0xFFu) << 23) * inv.x, ay = __uint_as_float(((ex >> 8) & 0xFFu) << 23) * inv.y, az = __uint_as_float(((ex >> 16) & 0xFFu) << 23) * inv.z;
            const float bx = (n0.x - o.x) * inv.x, by = (n0.y - o.y) * inv.y, bz = (n0.z - o.z) * inv.z;
            const uint32_t meta_lo = __float_as_uint(n1.z), meta_hi = __float_as_uint(n1.w);
            // near/far plane words per axis, selected once per node by the ray octant:
            // qlo x/y/z = n2.xy, n2.zw, n3.xy ; qhi x/y/z = n3.zw, n4.xy, n4.zw
            const bool px = d.x >= 0.0f, py = d.y >= 0.0f, pz = d.z >= 0.0f;
            const uint32_t nx0 = __float_as_uint(px ? n2.x : n3.z), nx1 = __float_as_uint(px ? n2.y : n3.w), fx0 = __float_as_uint(px ? n3.z : n2.x), fx1 = __float_as_uint(px ? n3.w : n2.y);
            const uint32_t ny0 = __float_as_uint(py ? n2.z : n4.x), ny1 = __float_as_uint(py ? n2.w : n4.y), fy0 = __float_as_uint(py ? n4.x : n2.z), fy1 = __float_as_uint(py ? n4.y : n2.w);
            const uint32_t nz0 = __float_as_uint(pz ? n3.x : n4.z), nz1 = __float_as_uint(pz ? n3.y : n4.w), fz0 = __float_as_uint(pz ? n4.z : n3.x), fz1 = __float_as_uint(pz ? n4.w : n3.y);
            const float tlim = best.t;
            uint32_t hitmask = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int sh = 8 * (i & 3);
                const uint32_t meta = ((i < 4 ? meta_lo : meta_hi) >> sh) & 0xFFu;
                const float tx0 = __fmaf_rn((float)(((i < 4 ? nx0 : nx1) >> sh) & 0xFFu), ax, bx), tx1 = __fmaf_rn((float)(((i < 4 ? fx0 : fx1) >> sh) & 0xFFu), ax, bx);
                const float ty0 = __fmaf_rn((float)(((i < 4 ? ny0 : ny1) >> sh) & 0xFFu), ay, by), ty1 = __fmaf_rn((float)(((i < 4 ? fy0 : fy1) >> sh) & 0xFFu), ay, by);
                const float tz0 = __fmaf_rn((float)(((i < 4 ? nz0 : nz1) >> sh) & 0xFFu), az, bz), tz1 = __fmaf_rn((float)(((i < 4 ? fz0 : fz1) >> sh) & 0xFFu), az, bz);
                // fmaxf/fminf drop NaNs (0*inf when the origin lies in a slab plane of a zero direction): conservative
                const float tn = fmaxf(fmaxf(tx0, ty0), fmaxf(tz0, 0.0f));
                const float tf = fminf(fminf(tx1, ty1), fminf(tz1, tlim)) * 1.0000004f;
                if (meta != 0u && tn <= tf) {
                    const uint32_t inner = (meta & (meta << 1)) & 0x10u;             // bits 3 and 4 both set <=> internal child
                    hitmask |= (meta >> 5) << ((meta ^ (inner ? oct_inv : 0u)) & 0x1Fu);
                }
            }
            ngroup = make_uint2(__float_as_uint(n1.x), (hitmask & 0xFF000000u) | (ex >> 24));
            tgroup = make_uint2(__float_as_uint(n1.y), hitmask & 0x00FFFFFFu);
        } else {
            tgroup = ngroup;
            ngroup = make_uint2(0u, 0u);
        }
        // ---- triangles of this node ----------------------------------------------------------------------
        while (tgroup.y != 0u) {
            const uint32_t ti = (uint32_t)__ffs(tgroup.y) - 1u;
            tgroup.y &= tgroup.y - 1u;
            const float4* tp = B.tris + (size_t)(tgroup.x + ti) * 3;
            const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
            if (COUNT) (*n_tris)++;
            float t, u, v;
            if (tri_test(o, d, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), t_max, t, u, v)) {
                const uint32_t prim1 = __float_as_uint(a.w) + 1u;
                if (ANY) { best.t = t; best.prim1 = prim1; best.b1 = u; best.b2 = v; return true; }
                if (best.prim1 == 0u || t < best.t || (t == best.t && prim1 < best.prim1)) { best.t = t; best.prim1 = prim1; best.b1 = u; best.b2 = v; }
            }
        }
        if (ngroup.y <= 0x00FFFFFFu) {
            if (st.n == 0) return true;
            ngroup = st.pop();
        }
        return false;
    }
};

// blocking form (one ray, run to completion)
template <bool ANY, bool COUNT>
HK_DEV HitRec bvh8_trace(const DevBvh& B, uint2* sm_stack, float3 o, float3 d, float t_max, uint32_t* n_nodes = nullptr, uint32_t* n_tris = nullptr) {
    Bvh8Walker<ANY, COUNT> w;
    w.begin(sm_stack, o, d, t_max);
    while (!w.step(B, n_nodes, n_tris)) {}
    return w.best;
}

// Persistent per-lane refill loop: every lane owns one in-flight ray; a lane whose ray finishes immediately claims the
// next queue entry (claims of the lanes that are idle at the same time are aggregated into one atomicAdd).
// IO supplies  bool load(idx, o, d, t_max)  and  void store(idx, hit).
template <bool ANY, bool COUNT, class IO>
HK_DEV void trace_queue(const DevBvh& B, uint2* sm_stack, uint32_t n, uint32_t* cursor, IO& io, uint32_t& traced, uint32_t& wn, uint32_t& wt) {
    Bvh8Walker<ANY, COUNT> w;
    bool busy = false;
    uint32_t idx = 0;
    const unsigned lane = threadIdx.x & 31u;
    for (;;) {
        if (!busy) {
            const unsigned m = __activemask();
            const unsigned leader = (unsigned)__ffs(m) - 1u;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(m));
            base = __shfl_sync(m, base, leader);
            idx = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
            if (idx >= n) break;
            float3 o, d; float tm;
            io.load(idx, o, d, tm);
            w.begin(sm_stack, o, d, tm);
            busy = true; traced++;
        }
        if (w.step(B, &wn, &wt)) { io.store(idx, w.best); busy = false; }
    }
}
