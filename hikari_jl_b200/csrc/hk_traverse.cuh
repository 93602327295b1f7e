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

#ifndef HK_SM_STACK
#define HK_SM_STACK 8
#endif
#define HK_LM_STACK 24
#ifndef HK_TRACE_THREADS
#define HK_TRACE_THREADS 128
#endif
// tuning switches (tools/variants.sh builds and times the alternatives on the GPU; defaults = the measured best)
#ifndef HK_QF_PRMT
#define HK_QF_PRMT 1          // 1: byte -> float by PRMT (ALU pipe), 0: by I2F (XU pipe)
#endif
#ifndef HK_REFILL_MIN
#define HK_REFILL_MIN 8       // idle lanes of a warp needed before the warp fetches new rays
#endif
#ifndef HK_NODE_PREFETCH
#define HK_NODE_PREFETCH 0    // 1: L1-prefetch the node that will be visited next while this node's triangles are tested
#endif
// 1: MUFU.RCP (rcp.approx.ftz, <= 1 ulp) for the inverse direction instead of the IEEE division (~8 instructions and a slow-path call per
// component, paid at every instance entry of the two-level walk: 7 % of k_trace's samples on C5 at 4-8 active lanes).  The inverse
// direction feeds the slab tests only -- both the slope A = inv * scale and the offset b = -o * inv of an axis use the SAME value, so an
// approximate reciprocal scales every slab distance of that axis by (1 +- 1 ulp) and nothing else; the far-side inflation below grows
// from 6 to 10 ulp to cover it.  Hits are decided by the exact triangle test, so primitive ids / t / barycentrics stay bit-exact
// (brute-force tests, image parity).  C5 trace 12.63 -> 12.17 ms/step, C3 2.33 -> 2.29, C2 1.055 -> 1.037.
#ifndef HK_RCP_APPROX
#define HK_RCP_APPROX 1
#endif
// far slab distance inflation: 6 ulp cover the rounding of the IEEE reciprocal and the two FMAs of a slab test; the approximate
// reciprocal (<= 1 ulp off, both on the near and the far side) gets 10
#define HK_SLAB_INFLATE (HK_RCP_APPROX ? 1.0000012f : 1.0000007f)
#ifndef HK_TRACE_BLOCKS_PER_SM
#define HK_TRACE_BLOCKS_PER_SM 8
#endif
#ifndef HK_TRI_HYBRID
#define HK_TRI_HYBRID 0
#endif
#ifndef HK_TRI_VOTE
#define HK_TRI_VOTE 0         // > 0: triangle tests wait until that many lanes of the warp have one pending
#endif
#ifndef HK_TRACE_BLOCKS_PER_SM_INST
#define HK_TRACE_BLOCKS_PER_SM_INST 8      // (the instanced walker's world-space ray lives in shared memory; measured on C5: 6 blocks 14.1 ms, 7 13.5, 8 13.0)
#endif

// one_bits = 0x3F800000, supplied by the host so that it reaches the kernels as a run-time value: see HK_QF in node_step()
// inst != nullptr: two-level BVH (HkGeometry.instances).  The top level occupies nodes[0..) with node 0 as its root and its leaf
// records tris[k] = {instance slot k, ...}; inst[4 * k ..] is the 64-byte record of the instance in leaf slot k: rows 0-2 of
// world_to_object and (root node of its mesh, first global primitive id, shading class << 28, instance index).
struct DevBvh { const float4* __restrict__ nodes; const float4* __restrict__ tris; uint32_t one_bits; const float4* __restrict__ inst; };
// prim1: 1-based global primitive id in the low 28 bits (0 = miss) | material type of the triangle's interface << 28.
// The type bits ride along for free (they sit in the spare word of the 48-byte triangle record, written by
// k_patch_tri_types once geometry and materials are both uploaded) so that the routing kernel needs no
// TriangleMeta -> interface -> material gather chain per ray.
struct HitRec { float t; uint32_t prim1; float b1, b2; };
#define HK_PRIM_MASK 0x0FFFFFFFu
#define HK_HIT_PRIM1(bits) ((bits) & HK_PRIM_MASK)
#define HK_HIT_MTYPE(bits) ((bits) >> 28)

HK_DEV float __frcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
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
    uint2* lm;          // overflow entries: a local-memory array owned by the caller (kept OUT of the walker struct --
                        // a dynamically indexed member array pins the whole struct, n and the node groups included,
                        // in local memory, which showed up as ~20 M local-store sectors per launch in ncu)
    int n;
    HK_DEV void push(uint2 v) { if (n < HK_SM_STACK) sm[n * HK_TRACE_THREADS] = v; else lm[n - HK_SM_STACK] = v; n++; }
    HK_DEV uint2 pop() { n--; return n < HK_SM_STACK ? sm[n * HK_TRACE_THREADS] : lm[n - HK_SM_STACK]; }
};

// Traversal state machine, split into two kinds of unit work so that a warp can interleave them lane by lane:
//   node_step(): pop the nearest pending internal child (from the current node group or the stack), fetch that 80-byte
//                node, slab-test its 8 quantised children -> new node group + the triangles of the leaf children hit;
//   tri_step():  test ONE pending triangle.
// The persistent loop (trace_queue) runs "node_step for lanes without pending triangles, then tri_step for lanes with
// pending triangles" per iteration: a lane whose leaf has several triangles keeps testing them in the following iterations
// while its neighbours already expand their next node, instead of the whole warp waiting for the longest triangle list.
// COUNT: accumulate node visits / triangle tests (roofline accounting).  ANY: stop at the first accepted hit.
// return marker pushed when the walk enters an instance: y has its top byte clear (a pending node group never has), bits 0-7 =
// the other instances of the top-level leaf group still to visit, bits 8-15 = that group's valid mask (one instance per slot)
HK_DEV uint32_t spread3(uint32_t b8) { uint32_t v = b8; v = (v | (v << 8)) & 0x00F00Fu; v = (v | (v << 4)) & 0x0C30C3u; v = (v | (v << 2)) & 0x249249u; return v; }
HK_DEV uint32_t gather3(uint32_t b24) { uint32_t v = b24 & 0x249249u; v = (v | (v >> 2)) & 0x0C30C3u; v = (v | (v >> 4)) & 0x00F00Fu; v = (v | (v >> 8)) & 0xFFu; return v; }

template <bool ANY, bool COUNT, bool INST = false>
struct Bvh8Walker {
    float3 o, d, inv;
    float t_max;
    uint32_t oct_inv;
    // INST only: the world-space ray (o, d above are the ray in the space being traversed) and its inverse direction live in this
    // lane's column of shared memory (wr[k * HK_TRACE_THREADS]: o 0-2, d 3-5, 1/d 6-8) -- they are only touched when the walk enters
    // or leaves an instance, so they neither occupy nine registers for the whole walk nor cost three divisions per return to the
    // top level (ncu on C5: the recomputed 1/d was 14 % of the kernel's samples, executed by 3.7 lanes) -- and the instance being
    // traversed
    float* wr;
    uint32_t prim_base, mtype_bits;
    bool in_blas;
    uint2 ngroup, tgroup;      // ngroup: (first internal child, octant-ordered hit bits << 24 | imask); tgroup: (first triangle, pending triangle bits)
    uint32_t tvalid;           // HkBvhNode::trivalid of the node tgroup came from
    TravStack st;
    HitRec best;

    HK_DEV void begin(uint2* sm_stack, uint2* lm_stack, float3 o_, float3 d_, float t_max_, float* wray = nullptr) {
        o = o_; d = d_; t_max = t_max_;
#if HK_RCP_APPROX
        inv = f3(__frcp_approx(d.x), __frcp_approx(d.y), __frcp_approx(d.z));
#else
        inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);   // feeds the (conservative) slab tests only
#endif
        // octant: bit set <=> direction component is non-negative (near children then sit at the - side)
        oct_inv = (d.x >= 0.0f ? 4u : 0u) | (d.y >= 0.0f ? 2u : 0u) | (d.z >= 0.0f ? 1u : 0u);
        st.sm = sm_stack; st.lm = lm_stack; st.n = 0;
        ngroup = make_uint2(0u, 0x80000000u);   // root: base 0 and an empty internal mask => child index 0 whatever the octant
        // A ray with a NaN / infinite component, a zero direction or a NaN t_max cannot pass the exact triangle test (every
        // comparison fails), but its slab distances are NaN, which fminf/fmaxf drop: it would "hit" every box and walk the
        // whole tree (measured: a handful of such rays from degenerate BSDF samples cost 0.8 s per pass at 10 M triangles).
        // It is a miss by definition, so it never starts.
        const bool finite = fabsf(o.x) <= 3.4e38f && fabsf(o.y) <= 3.4e38f && fabsf(o.z) <= 3.4e38f &&
                            fabsf(d.x) <= 3.4e38f && fabsf(d.y) <= 3.4e38f && fabsf(d.z) <= 3.4e38f && t_max == t_max;
        if (!finite || (d.x == 0.0f && d.y == 0.0f && d.z == 0.0f)) ngroup.y = 0u;
        tgroup = make_uint2(0u, 0u); tvalid = 0u;
        best.t = t_max; best.prim1 = 0; best.b1 = 0.0f; best.b2 = 0.0f;
        if (INST) {
            wr = wray; in_blas = false; prim_base = 0u; mtype_bits = 0u;
            wr[0] = o.x; wr[HK_TRACE_THREADS] = o.y; wr[2 * HK_TRACE_THREADS] = o.z; wr[3 * HK_TRACE_THREADS] = d.x; wr[4 * HK_TRACE_THREADS] = d.y; wr[5 * HK_TRACE_THREADS] = d.z;
            wr[6 * HK_TRACE_THREADS] = inv.x; wr[7 * HK_TRACE_THREADS] = inv.y; wr[8 * HK_TRACE_THREADS] = inv.z;
        }
    }
    HK_DEV void set_ray(float3 o_, float3 d_) {
        o = o_; d = d_;
#if HK_RCP_APPROX
        inv = f3(__frcp_approx(d.x), __frcp_approx(d.y), __frcp_approx(d.z));
#else
        inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
#endif
        oct_inv = (d.x >= 0.0f ? 4u : 0u) | (d.y >= 0.0f ? 2u : 0u) | (d.z >= 0.0f ? 1u : 0u);
    }
    // precondition: no pending triangles.  Returns true when the traversal is finished (nothing left to visit).
    HK_DEV bool node_step(const DevBvh& B, uint32_t* n_nodes) {
        if (ngroup.y <= 0x00FFFFFFu) {
            if (st.n == 0) return true;
            ngroup = st.pop();
            if (INST && ngroup.y <= 0x00FFFFFFu) {      // return marker: the instance is done, back to the top level in world space
                in_blas = false;
                o = f3(wr[0], wr[HK_TRACE_THREADS], wr[2 * HK_TRACE_THREADS]); d = f3(wr[3 * HK_TRACE_THREADS], wr[4 * HK_TRACE_THREADS], wr[5 * HK_TRACE_THREADS]);
                inv = f3(wr[6 * HK_TRACE_THREADS], wr[7 * HK_TRACE_THREADS], wr[8 * HK_TRACE_THREADS]);
                oct_inv = (d.x >= 0.0f ? 4u : 0u) | (d.y >= 0.0f ? 2u : 0u) | (d.z >= 0.0f ? 1u : 0u);
                tgroup = make_uint2(ngroup.x, spread3(ngroup.y & 0xFFu)); tvalid = spread3((ngroup.y >> 8) & 0xFFu);
                ngroup.y = 0u;
                return false;
            }
        }
        // ---- pop the nearest pending internal child of this group -------------------------------------
        const uint32_t hits = ngroup.y;
        const uint32_t bit = 31u - (uint32_t)__clz(hits);
        ngroup.y &= ~(1u << bit);
        if (ngroup.y > 0x00FFFFFFu) st.push(ngroup);
        const uint32_t slot = (bit - 24u) ^ oct_inv;
        const uint32_t node_idx = ngroup.x + (uint32_t)__popc((hits & 0xFFu) & ((1u << slot) - 1u));
        // ---- fetch the 80-byte node as five 16-byte loads --------------------------------------------
        const float4* np = B.nodes + (size_t)node_idx * 5;
        const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
        if (COUNT) (*n_nodes)++;
        const uint32_t ex = __float_as_uint(n0.w);
        // Slab test of the 8 quantised child boxes: t = q * (2^e / d) + (p - o) / d.  Node culling only has to be
        // conservative (the exact, contraction-free arithmetic is reserved for the triangle test), so:
        //  * a quantised byte q becomes a float with one PRMT instead of an I2F (quarter-rate XU pipe, which ncu showed
        //    as the busiest pipe of this kernel): bytes [3F 80 q 00] = 1 + q * 2^-15, and
        //    t = (1 + q 2^-15) * A + (b - A) with A = 2^15 * 2^e / d;
        //  * b - A is rounded down for the entry planes and up for the exit planes, so the extra rounding can only
        //    widen a box; the remaining FMA rounding is covered by the relative slack on the exit distance.
        const float Ax = __uint_as_float(((ex & 0xFFu) + 15u) << 23) * inv.x, Ay = __uint_as_float((((ex >> 8) & 0xFFu) + 15u) << 23) * inv.y,
                    Az = __uint_as_float((((ex >> 16) & 0xFFu) + 15u) << 23) * inv.z;
        const float bx = (n0.x - o.x) * inv.x, by = (n0.y - o.y) * inv.y, bz = (n0.z - o.z) * inv.z;
        const float bxn = __fsub_rd(bx, Ax), byn = __fsub_rd(by, Ay), bzn = __fsub_rd(bz, Az);
        const float bxf = __fsub_ru(bx, Ax), byf = __fsub_ru(by, Ay), bzf = __fsub_ru(bz, Az);
        // near/far plane words per axis, selected once per node by the ray octant:
        // qlo x/y/z = n2.xy, n2.zw, n3.xy ; qhi x/y/z = n3.zw, n4.xy, n4.zw
        const bool px = d.x >= 0.0f, py = d.y >= 0.0f, pz = d.z >= 0.0f;
        const uint32_t nx0 = __float_as_uint(px ? n2.x : n3.z), nx1 = __float_as_uint(px ? n2.y : n3.w), fx0 = __float_as_uint(px ? n3.z : n2.x), fx1 = __float_as_uint(px ? n3.w : n2.y);
        const uint32_t ny0 = __float_as_uint(py ? n2.z : n4.x), ny1 = __float_as_uint(py ? n2.w : n4.y), fy0 = __float_as_uint(py ? n4.x : n2.z), fy1 = __float_as_uint(py ? n4.y : n2.w);
        const uint32_t nz0 = __float_as_uint(pz ? n3.x : n4.z), nz1 = __float_as_uint(pz ? n3.y : n4.w), fz0 = __float_as_uint(pz ? n4.z : n3.x), fz1 = __float_as_uint(pz ? n4.w : n3.y);
        const float tlim = best.t;
        uint32_t hits8 = 0;                       // bit s: the ray overlaps the box of child slot s (empty slots may set bits; they are masked below)
#if HK_QF_PRMT
        // 0x3F800000 arrives as a run-time value (DevBvh::one_bits) so that it sits in ONE register and PRMT takes the byte
        // selector as its immediate; with both compile-time constants ptxas put the float constant in the immediate slot and
        // re-materialised the four selectors with ~40 extra moves per node
        const uint32_t k_one = B.one_bits;
#define HK_QF(word, k) __uint_as_float(__byte_perm((word), k_one, 0x7604u | ((k) << 4)))
#else
#define HK_QF(word, k) __fmaf_rn((float)(((word) >> (8 * (k))) & 0xFFu), 3.0517578125e-05f, 1.0f)
#endif
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int k = i & 3;
            const float tx0 = __fmaf_rn(HK_QF(i < 4 ? nx0 : nx1, k), Ax, bxn), tx1 = __fmaf_rn(HK_QF(i < 4 ? fx0 : fx1, k), Ax, bxf);
            const float ty0 = __fmaf_rn(HK_QF(i < 4 ? ny0 : ny1, k), Ay, byn), ty1 = __fmaf_rn(HK_QF(i < 4 ? fy0 : fy1, k), Ay, byf);
            const float tz0 = __fmaf_rn(HK_QF(i < 4 ? nz0 : nz1, k), Az, bzn), tz1 = __fmaf_rn(HK_QF(i < 4 ? fz0 : fz1, k), Az, bzf);
            // fmaxf/fminf drop NaNs (inf - inf when a direction component is zero): that axis is left unconstrained
            const float tn = fmaxf(fmaxf(tx0, ty0), fmaxf(tz0, 0.0f));
            const float tf = fminf(fminf(tx1, ty1), fminf(tz1, tlim)) * HK_SLAB_INFLATE;
            if (tn <= tf) hits8 |= 1u << i;
        }
#undef HK_QF
        // internal children: reorder the slot-indexed hit bits into octant traversal order (bit j <- bit j ^ oct_inv) with
        // three conditional bit-group swaps, instead of one variable shift per child
        const uint32_t imask = ex >> 24;
        uint32_t ih = hits8 & imask;
        { const uint32_t s1 = oct_inv & 1u, s2 = oct_inv & 2u, s4 = oct_inv & 4u;
          ih = ((ih >> s1) & 0x55u) | ((ih << s1) & 0xAAu);
          ih = ((ih >> s2) & 0x33u) | ((ih << s2) & 0xCCu);
          ih = ((ih >> s4) & 0x0Fu) | ((ih << s4) & 0xF0u); }
        // leaf children: slot s owns triangle bits 3s..3s+2; spread the hit bits to those positions and keep the valid ones
        uint32_t th = hits8;
        th = (th | (th << 8)) & 0x00F00Fu; th = (th | (th << 4)) & 0x0C30C3u; th = (th | (th << 2)) & 0x249249u;
        tvalid = __float_as_uint(n1.z);
        ngroup = make_uint2(__float_as_uint(n1.x), (ih << 24) | imask);
        tgroup = make_uint2(__float_as_uint(n1.y), (th * 7u) & tvalid);
        return false;
    }
    // precondition: tgroup.y != 0.  Tests one pending triangle; returns true only for ANY when a hit was accepted.
    // INST, at the top level: the pending leaf entry is an instance -- enter it (object-space ray, the mesh's root as the node group).
    HK_DEV bool tri_step(const DevBvh& B, uint32_t* n_tris) {
        const uint32_t tbit = (uint32_t)__ffs(tgroup.y) - 1u;
        tgroup.y &= tgroup.y - 1u;
        if (INST && !in_blas) {
            const uint32_t slot = tgroup.x + (uint32_t)__popc(tvalid & ((1u << tbit) - 1u));
            const float4* ip = B.inst + (size_t)slot * 4;
            const float4 r0 = __ldg(ip), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2), r3 = __ldg(ip + 3);
            if (ngroup.y > 0x00FFFFFFu) st.push(ngroup);
            st.push(make_uint2(tgroup.x, gather3(tgroup.y) | (gather3(tvalid) << 8)));
            const float3 wo = f3(wr[0], wr[HK_TRACE_THREADS], wr[2 * HK_TRACE_THREADS]), wd = f3(wr[3 * HK_TRACE_THREADS], wr[4 * HK_TRACE_THREADS], wr[5 * HK_TRACE_THREADS]);
            const float3 oo = f3(r0.x * wo.x + r0.y * wo.y + r0.z * wo.z + r0.w, r1.x * wo.x + r1.y * wo.y + r1.z * wo.z + r1.w, r2.x * wo.x + r2.y * wo.y + r2.z * wo.z + r2.w);
            const float3 od = f3(r0.x * wd.x + r0.y * wd.y + r0.z * wd.z, r1.x * wd.x + r1.y * wd.y + r1.z * wd.z, r2.x * wd.x + r2.y * wd.y + r2.z * wd.z);
            set_ray(oo, od);
            in_blas = true; prim_base = __float_as_uint(r3.y); mtype_bits = __float_as_uint(r3.z);
            ngroup = make_uint2(__float_as_uint(r3.x), 0x80000000u);
            tgroup.y = 0u;
            return false;
        }
        const float4* tp = B.tris + (size_t)(tgroup.x + (uint32_t)__popc(tvalid & ((1u << tbit) - 1u))) * 3;
        const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
        if (COUNT) (*n_tris)++;
        float t, u, v;
        if (tri_test(o, d, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), t_max, t, u, v)) {
            const uint32_t prim1 = INST ? ((prim_base + __float_as_uint(a.w) + 1u) | mtype_bits) : ((__float_as_uint(a.w) + 1u) | (__float_as_uint(b.w) << 28));
            if (ANY) { best.t = t; best.prim1 = prim1; best.b1 = u; best.b2 = v; return true; }
            if (best.prim1 == 0u || t < best.t || (t == best.t && HK_HIT_PRIM1(prim1) < HK_HIT_PRIM1(best.prim1))) { best.t = t; best.prim1 = prim1; best.b1 = u; best.b2 = v; }
        }
        return false;
    }
};

// shared memory of a traversal kernel: the per-lane stack columns and, for the instanced walker, the per-lane world-space ray
#define HK_TRACE_SMEM(INST) __shared__ uint2 sm_stack[HK_SM_STACK * HK_TRACE_THREADS]; __shared__ float sm_wray[(INST) ? 9 * HK_TRACE_THREADS : 1]
// blocking form (one ray, run to completion)
template <bool ANY, bool COUNT, bool INST = false>
HK_DEV HitRec bvh8_trace(const DevBvh& B, uint2* sm_stack, float* sm_wray, float3 o, float3 d, float t_max, uint32_t* n_nodes = nullptr, uint32_t* n_tris = nullptr) {
    Bvh8Walker<ANY, COUNT, INST> w;
    uint2 lm_stack[HK_LM_STACK];
    w.begin(sm_stack, lm_stack, o, d, t_max, sm_wray);
    for (;;) {
        if (w.node_step(B, n_nodes)) break;
        bool hit = false;
        while (w.tgroup.y != 0u && !hit) hit = w.tri_step(B, n_tris);
        if (hit) break;
    }
    return w.best;
}

// Persistent per-lane refill loop: every lane owns one in-flight ray.  Finished lanes claim new queue entries together
// (one atomicAdd per refill) once at least HK_REFILL_MIN lanes of the warp are idle -- the ray set-up is a divergent
// region of its own, so it is batched instead of being run for single lanes in almost every iteration.
// IO supplies  uint32_t load(idx, o, d, t_max) -> token  and  void store(token, hit).
template <bool ANY, bool COUNT, bool INST, class IO>
HK_DEV void trace_queue(const DevBvh& B, uint2* sm_stack, float* sm_wray, uint32_t n, uint32_t* cursor, IO& io, uint32_t& traced, uint32_t& wn, uint32_t& wt) {
    Bvh8Walker<ANY, COUNT, INST> w;
    uint2 lm_stack[HK_LM_STACK];
    bool busy = false, exhausted = false;
    uint32_t token = 0;
    const unsigned lane = threadIdx.x & 31u;
    for (;;) {
        unsigned idle = __ballot_sync(0xFFFFFFFFu, !busy);
        if (!exhausted && (uint32_t)__popc(idle) >= (uint32_t)HK_REFILL_MIN) {
            if (!busy) {
                const unsigned leader = (unsigned)__ffs(idle) - 1u;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(idle));
                base = __shfl_sync(idle, base, leader);
                const uint32_t idx = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
                if (idx < n) {
                    float3 o, d; float tm;
                    token = io.load(idx, o, d, tm);
                    w.begin(sm_stack, lm_stack, o, d, tm, sm_wray);
                    busy = true; traced++;
                }
            }
            idle = __ballot_sync(0xFFFFFFFFu, !busy);
            exhausted = idle != 0u;          // a lane that found the queue empty
        }
        if (idle == 0xFFFFFFFFu) break;      // only reachable once the queue is exhausted
#if HK_TRI_VOTE > 0 && HK_TRI_HYBRID
        // node steps every iteration; the triangle tests of the warp wait until HK_TRI_VOTE lanes have one pending (or no lane took a node step)
        bool fin = false, did_node = false;
        if (busy && w.tgroup.y == 0u) { fin = w.node_step(B, &wn); did_node = true; }
        const unsigned want_tri = __ballot_sync(0xFFFFFFFFu, busy && !fin && w.tgroup.y != 0u);
        const unsigned nodes = __ballot_sync(0xFFFFFFFFu, did_node);
        if ((uint32_t)__popc(want_tri) >= (uint32_t)HK_TRI_VOTE || nodes == 0u) { if (busy && !fin && w.tgroup.y != 0u) fin = w.tri_step(B, &wt); }
        if (busy && fin) { io.store(token, w.best); busy = false; }
#elif HK_TRI_VOTE > 0
        // per-warp vote between the two kinds of unit work: triangle tests only run once HK_TRI_VOTE lanes have one pending (or no lane can
        // take a node step), so their ~80 instructions are issued for many lanes at a time instead of ~7 of 32 in every iteration
        const unsigned want_tri = __ballot_sync(0xFFFFFFFFu, busy && w.tgroup.y != 0u);
        const unsigned want_node = __ballot_sync(0xFFFFFFFFu, busy && w.tgroup.y == 0u);
        if (busy) {
            bool fin = false;
            if ((uint32_t)__popc(want_tri) >= (uint32_t)HK_TRI_VOTE || want_node == 0u) { if (w.tgroup.y != 0u) fin = w.tri_step(B, &wt); }
            else if (w.tgroup.y == 0u) fin = w.node_step(B, &wn);
            if (fin) { io.store(token, w.best); busy = false; }
        }
#else
        if (busy) {
            bool fin = false;
            if (w.tgroup.y == 0u) fin = w.node_step(B, &wn);
            if (!fin && w.tgroup.y != 0u) fin = w.tri_step(B, &wt);
            if (fin) { io.store(token, w.best); busy = false; }
        }
#endif
    }
}
