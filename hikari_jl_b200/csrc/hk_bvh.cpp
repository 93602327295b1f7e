// hk_bvh.cpp — host builder: binned-SAH BVH2 down to single triangles -> SAH-optimal collapse to BVH8 by dynamic
// programming (Ylitie, Karras, Laine 2017, section 4: per BVH2 node the cheapest forest of <= i wide subtrees, i = 1..7;
// leaves of <= 3 triangles are formed by the same optimisation) -> octant slot assignment -> 8-bit quantisation
// (conservative) -> BFS layout with contiguous internal children / leaf triangles.
// (The first version expanded the largest child greedily top-down: 3.9 children per 8-wide node on a tessellated mesh,
// half of all nodes with only 2 -- every visit pays for 8 slab tests.)
// Replaces Raycore's BVH/TLAS construction behind scene.accel (reference: Raycore.jl, not vendored;
// call sites src/scene.jl:146-151 sync!).
#include "hk_bvh.h"
#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>

namespace {

const float INF = std::numeric_limits<float>::infinity();

struct Box {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; k++) { lo[k] = INF; hi[k] = -INF; } }
    void grow(const Box& b) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
    void grow(const float* p) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
    float half_area() const {
        float d[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
        if (d[0] < 0) return 0.0f;
        return d[0] * d[1] + d[1] * d[2] + d[2] * d[0];
    }
};
struct Node2 { Box box; uint32_t left, right, first, count; };   // [first, first+count) of `order` = the subtree's triangles; left == 0 => BVH2 leaf

struct Builder {
    const std::vector<Box>& pb;      // per-primitive (padded) boxes
    const std::vector<float>& cen;   // centroids [n][3]
    std::vector<uint32_t> order;
    std::vector<Node2> nodes;
    std::atomic<uint32_t> n_nodes{0};
#ifndef HK_BVH_NBINS
#define HK_BVH_NBINS 32
#endif
#ifndef HK_BVH_CPRIM
#define HK_BVH_CPRIM 0.4f     // cost of one triangle test relative to one 8-wide node visit (~90 vs ~210 instructions)
#endif
#ifndef HK_BVH_CT
#define HK_BVH_CT 0.125f      // traversal-step cost relative to one leaf (<= 3 triangles) test in the SAH
#endif
    static constexpr int NBINS = HK_BVH_NBINS;
    uint32_t MAX_LEAF = 3;                        // primitives per BVH8 leaf child (unary count in 3 bits of trivalid); 1 for a top-level BVH over instances

    Builder(const std::vector<Box>& b, const std::vector<float>& c) : pb(b), cen(c) {}

    uint32_t alloc() { return n_nodes.fetch_add(1); }

    void build(uint32_t me, uint32_t first, uint32_t count, int depth) {
        Box box, cbox; box.reset(); cbox.reset();
        for (uint32_t i = first; i < first + count; i++) { uint32_t p = order[i]; box.grow(pb[p]); cbox.grow(&cen[3 * (size_t)p]); }
        nodes[me].box = box; nodes[me].first = first; nodes[me].count = count;
        if (count == 1 || (count <= MAX_LEAF && depth > 60)) { make_leaf(me, first, count); return; }
        // binned SAH over the three axes
        float best_cost = INF; int best_axis = -1, best_bin = 0;
        for (int ax = 0; ax < 3; ax++) {
            float ext = cbox.hi[ax] - cbox.lo[ax];
            if (!(ext > 0.0f)) continue;
            Box bb[NBINS]; uint32_t bc[NBINS];
            for (int b = 0; b < NBINS; b++) { bb[b].reset(); bc[b] = 0; }
            float k1 = NBINS * (1.0f - 1e-6f) / ext;
            for (uint32_t i = first; i < first + count; i++) {
                uint32_t p = order[i];
                int b = (int)(k1 * (cen[3 * (size_t)p + ax] - cbox.lo[ax]));
                b = b < 0 ? 0 : (b >= NBINS ? NBINS - 1 : b);
                bb[b].grow(pb[p]); bc[b]++;
            }
            float ra[NBINS]; uint32_t rc[NBINS];
            Box acc; acc.reset(); uint32_t c = 0;
            for (int b = NBINS - 1; b > 0; b--) { acc.grow(bb[b]); c += bc[b]; ra[b] = acc.half_area(); rc[b] = c; }
            acc.reset(); c = 0;
            for (int b = 0; b < NBINS - 1; b++) {
                acc.grow(bb[b]); c += bc[b];
                if (c == 0 || rc[b + 1] == 0) continue;
                float cost = acc.half_area() * (float)c + ra[b + 1] * (float)rc[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = ax; best_bin = b; }
            }
        }
        uint32_t mid;
        if (best_axis >= 0) {
            float ext = cbox.hi[best_axis] - cbox.lo[best_axis];
            float k1 = NBINS * (1.0f - 1e-6f) / ext;
            auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](uint32_t p) {
                int b = (int)(k1 * (cen[3 * (size_t)p + best_axis] - cbox.lo[best_axis]));
                b = b < 0 ? 0 : (b >= NBINS ? NBINS - 1 : b);
                return b <= best_bin;
            });
            mid = (uint32_t)(it - order.begin());
            if (mid == first || mid == first + count) mid = first + count / 2;
        } else {
            if (count <= MAX_LEAF) { make_leaf(me, first, count); return; }
            mid = first + count / 2;   // coincident centroids: split the list
        }
        uint32_t l = alloc(), r = alloc();
        nodes[me].left = l; nodes[me].right = r;
        if (count > 200000) {
            #pragma omp task shared(nodes)
            build(l, first, mid - first, depth + 1);
            #pragma omp task shared(nodes)
            build(r, mid, first + count - mid, depth + 1);
            #pragma omp taskwait
        } else {
            build(l, first, mid - first, depth + 1);
            build(r, mid, first + count - mid, depth + 1);
        }
    }
    void make_leaf(uint32_t me, uint32_t first, uint32_t count) {
        nodes[me].left = nodes[me].right = 0;
    }
};

inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

}  // namespace

// Generic core: a BVH8 over n primitives given by their (padded) boxes and centroids; emit(prim, record) fills the 48-byte leaf
// record of one primitive (a triangle for a mesh, an instance reference for the top level).  Returns the depth of the tree
// (number of 8-wide levels), which bounds the traversal stack.
template <class Emit>
static int build_bvh8_core(const std::vector<Box>& pb, const std::vector<float>& cen, uint32_t n_tris, uint32_t max_leaf, Emit emit, HkBvh& out) {
    out.nodes.clear(); out.tris.clear();
    for (int k = 0; k < 3; k++) { out.bounds_min[k] = 0; out.bounds_max[k] = 0; }
    if (n_tris == 0) {
        HkBvhNode root; std::memset(&root, 0, sizeof(root)); out.nodes.push_back(root); return 1;
    }
    Builder B(pb, cen);
    B.MAX_LEAF = max_leaf;
    B.order.resize(n_tris); std::iota(B.order.begin(), B.order.end(), 0u);
    B.nodes.resize(2 * (size_t)n_tris);
    uint32_t root2 = B.alloc();
    #pragma omp parallel
    {
        #pragma omp single
        B.build(root2, 0, n_tris, 0);
    }
    const std::vector<Node2>& N2 = B.nodes;
    for (int k = 0; k < 3; k++) { out.bounds_min[k] = N2[root2].box.lo[k]; out.bounds_max[k] = N2[root2].box.hi[k]; }

    // ---- SAH-optimal collapse: cost[n][i-1] = cheapest representation of BVH2 subtree n as a forest of <= i wide-BVH roots
    // (a root = a leaf of <= MAX_LEAF triangles or an internal 8-wide node), i = 1..7.  Children carry larger indices than
    // their parent (allocated later), so one sweep from the last node to the first visits children before parents. ---------
    const uint32_t n2 = B.n_nodes.load();
    const float C_NODE = 1.0f, C_PRIM = HK_BVH_CPRIM;
    std::vector<std::array<float, 7>> cost(n2);
    auto distribute = [&](uint32_t l, uint32_t r, int j, int* best_k) {       // min over 0 < k < j of cost[l][k] + cost[r][j-k]
        float best = INF; int bk = 1;
        for (int k = 1; k < j; k++) {
            const int kl = std::min(k, 7), kr = std::min(j - k, 7);
            const float c = cost[l][kl - 1] + cost[r][kr - 1];
            if (c < best) { best = c; bk = k; }
        }
        if (best_k) *best_k = bk;
        return best;
    };
    auto leaf_cost = [&](uint32_t n) { return N2[n].count <= B.MAX_LEAF ? N2[n].box.half_area() * (float)N2[n].count * C_PRIM : INF; };
    for (int64_t n = (int64_t)n2 - 1; n >= 0; n--) {
        const Node2& nd = N2[n];
        if (nd.left == 0) { cost[n].fill(leaf_cost((uint32_t)n)); continue; }
        const float c_internal = distribute(nd.left, nd.right, 8, nullptr) + nd.box.half_area() * C_NODE;
        cost[n][0] = std::min(leaf_cost((uint32_t)n), c_internal);
        for (int i = 2; i <= 7; i++) cost[n][i - 1] = std::min(distribute(nd.left, nd.right, i, nullptr), cost[n][i - 2]);
    }
    // the roots of the cheapest forest of <= i subtrees below BVH2 node n (decisions re-derived from the cost table)
    struct Root { uint32_t n2; bool leaf; };
    auto gather = [&](auto&& self, uint32_t n, int i, std::vector<Root>& out_roots) -> void {
        const Node2& nd = N2[n];
        if (nd.left == 0) { out_roots.push_back(Root{n, true}); return; }
        while (i > 1 && cost[n][i - 1] == cost[n][i - 2]) i--;       // the extra roots bought nothing
        if (i == 1) {
            const float lc = leaf_cost(n);
            out_roots.push_back(Root{n, lc <= cost[n][0] && lc < INF});
            return;
        }
        int k; distribute(nd.left, nd.right, i, &k);
        self(self, nd.left, std::min(k, 7), out_roots); self(self, nd.right, std::min(i - k, 7), out_roots);
    };

    // ---- layout (BFS) -----------------------------------------------------------------------------
    out.nodes.reserve((size_t)n_tris / 6 + 16); out.tris.reserve(n_tris);
    struct Pending { uint32_t n2; uint32_t out_idx; int depth; };
    int max_depth = 1;
    std::vector<Pending> queue; queue.reserve((size_t)n_tris / 6 + 16);
    out.nodes.emplace_back(); std::memset(&out.nodes[0], 0, sizeof(HkBvhNode));
    // a root that is itself a leaf gets wrapped: treat it as a wide node with one leaf child
    queue.push_back(Pending{root2, 0, 1});
    std::vector<Root> roots;
    for (size_t qi = 0; qi < queue.size(); qi++) {
        Pending cur = queue[qi];
        uint32_t ch[8]; bool ch_leaf[8]; int nch = 0;
        roots.clear();
        if (N2[cur.n2].left == 0) roots.push_back(Root{cur.n2, true});
        else {
            int k; distribute(N2[cur.n2].left, N2[cur.n2].right, 8, &k);
            gather(gather, N2[cur.n2].left, std::min(k, 7), roots); gather(gather, N2[cur.n2].right, std::min(8 - k, 7), roots);
        }
        for (const Root& r : roots) { ch[nch] = r.n2; ch_leaf[nch] = r.leaf; nch++; }
        // node frame
        Box nb; nb.reset();
        for (int i = 0; i < nch; i++) nb.grow(N2[ch[i]].box);
        float ctr[3] = {0.5f * (nb.lo[0] + nb.hi[0]), 0.5f * (nb.lo[1] + nb.hi[1]), 0.5f * (nb.lo[2] + nb.hi[2])};
        // greedy octant slot assignment: slot bit set <=> child lies towards + along that axis
        int slot_of[8]; bool slot_used[8] = {false}; bool assigned[8] = {false};
        float scost[8][8];
        for (int i = 0; i < nch; i++) {
            const Box& b = N2[ch[i]].box;
            float c[3] = {0.5f * (b.lo[0] + b.hi[0]) - ctr[0], 0.5f * (b.lo[1] + b.hi[1]) - ctr[1], 0.5f * (b.lo[2] + b.hi[2]) - ctr[2]};
            for (int s = 0; s < 8; s++) scost[i][s] = ((s & 4) ? c[0] : -c[0]) + ((s & 2) ? c[1] : -c[1]) + ((s & 1) ? c[2] : -c[2]);
        }
        for (int it = 0; it < nch; it++) {
            int bi = -1, bs = -1; float bc = -INF;
            for (int i = 0; i < nch; i++) if (!assigned[i]) for (int s = 0; s < 8; s++) if (!slot_used[s] && scost[i][s] > bc) { bc = scost[i][s]; bi = i; bs = s; }
            assigned[bi] = true; slot_used[bs] = true; slot_of[bi] = bs;
        }
        int child_at[8]; for (int s = 0; s < 8; s++) child_at[s] = -1;
        for (int i = 0; i < nch; i++) child_at[slot_of[i]] = i;

        HkBvhNode node; std::memset(&node, 0, sizeof(node));
        float scale[3];
        for (int k = 0; k < 3; k++) {
            node.p[k] = nb.lo[k];
            float ext = nb.hi[k] - nb.lo[k];
            int e;
            if (!(ext > 0.0f)) e = -100;
            else {
                int ex; std::frexp(ext / 255.0f, &ex);   // ext/255 = m * 2^ex, m in [0.5,1)
                e = ex;                                   // 2^ex >= ext/255
                if (e < -120) e = -120;
            }
            for (;;) {   // make sure the largest child max fits in 8 bits
                scale[k] = std::ldexp(1.0f, e);
                if (std::ceil((nb.hi[k] - nb.lo[k]) / scale[k]) <= 255.0f) break;
                e++;
            }
            node.e[k] = (uint8_t)(e + 127);
            scale[k] = u2f((uint32_t)node.e[k] << 23);
        }
        node.child_base = (uint32_t)out.nodes.size();
        node.tri_base = (uint32_t)out.tris.size();
        for (int s = 0; s < 8; s++) {
            int i = child_at[s];
            if (i < 0) continue;
            const Node2& c = N2[ch[i]];
            for (int k = 0; k < 3; k++) {
                float lo = std::floor((c.box.lo[k] - node.p[k]) / scale[k]);
                float hi = std::ceil((c.box.hi[k] - node.p[k]) / scale[k]);
                lo = std::min(std::max(lo, 0.0f), 255.0f); hi = std::min(std::max(hi, 0.0f), 255.0f);
                while (lo > 0.0f && node.p[k] + lo * scale[k] > c.box.lo[k]) lo -= 1.0f;
                while (hi < 255.0f && node.p[k] + hi * scale[k] < c.box.hi[k]) hi += 1.0f;
                node.qlo[k][s] = (uint8_t)lo; node.qhi[k][s] = (uint8_t)hi;
            }
            if (!ch_leaf[i]) {
                node.imask |= (uint8_t)(1u << s);
            } else {
                uint32_t unary = c.count == 1 ? 1u : (c.count == 2 ? 3u : 7u);
                node.trivalid |= unary << (3 * s);
                uint32_t prims[3];
                for (uint32_t t = 0; t < c.count; t++) prims[t] = B.order[c.first + t];
                std::sort(prims, prims + c.count);                  // deterministic leaf order
                for (uint32_t t = 0; t < c.count; t++) {
                    uint32_t prim = prims[t];
                    HkBvhTri T; std::memset(&T, 0, sizeof(T));
                    emit(prim, T);
                    out.tris.push_back(T);
                }
            }
        }
        // internal children in slot order, contiguous
        for (int s = 0; s < 8; s++) {
            int i = child_at[s];
            if (i < 0 || ch_leaf[i]) continue;
            uint32_t idx = (uint32_t)out.nodes.size();
            out.nodes.emplace_back(); std::memset(&out.nodes.back(), 0, sizeof(HkBvhNode));
            queue.push_back(Pending{ch[i], idx, cur.depth + 1}); max_depth = std::max(max_depth, cur.depth + 1);
        }
        out.nodes[cur.out_idx] = node;
    }
    return max_depth;
}

static void triangle_boxes(const float* positions, const uint32_t* indices, uint32_t n_tris, std::vector<Box>& pb, std::vector<float>& cen) {
    pb.resize(n_tris); cen.resize(3 * (size_t)n_tris);
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n_tris; i++) {
        Box b; b.reset();
        for (int v = 0; v < 3; v++) b.grow(positions + 3 * (size_t)indices[3 * i + v]);
        for (int k = 0; k < 3; k++) {
            cen[3 * i + k] = 0.5f * (b.lo[k] + b.hi[k]);
            float pad = 1.0e-5f * std::max(std::fabs(b.lo[k]), std::fabs(b.hi[k])) + 1.0e-6f;
            b.lo[k] -= pad; b.hi[k] += pad;
        }
        pb[i] = b;
    }
}

int hk_build_bvh8(const float* positions, const uint32_t* indices, uint32_t n_tris, HkBvh& out) {
    std::vector<Box> pb; std::vector<float> cen;
    triangle_boxes(positions, indices, n_tris, pb, cen);
    return build_bvh8_core(pb, cen, n_tris, 3, [&](uint32_t prim, HkBvhTri& T) {
        const float* a = positions + 3 * (size_t)indices[3 * (size_t)prim];
        const float* b = positions + 3 * (size_t)indices[3 * (size_t)prim + 1];
        const float* cc = positions + 3 * (size_t)indices[3 * (size_t)prim + 2];
        for (int k = 0; k < 3; k++) { T.v0[k] = a[k]; T.e1[k] = b[k] - a[k]; T.e2[k] = cc[k] - a[k]; }
        T.prim = prim;
    }, out);
}

// Two-level scene BVH (HkGeometry.instances): one bottom-level BVH8 per mesh over its OBJECT-space triangles (leaf records carry the
// face index within the mesh) and a top-level BVH8 over the instances' world boxes (one instance per leaf slot; its 48-byte leaf
// record is {instance index, 0...}).  Everything lands in ONE node array -- top level first, root = node 0, then the meshes with their
// child / triangle bases rebased -- and ONE leaf-record array (top-level records first).  mesh_root[m] = node index of mesh m's root.
int hk_build_scene_bvh(const float* positions, const uint32_t* indices, const HkMeshRange* meshes, uint32_t n_meshes,
                       const HkInstanceXf* inst, uint32_t n_inst, HkBvh& out, std::vector<uint32_t>& mesh_root, int* depth_top, int* depth_bottom) {
    std::vector<HkBvh> blas(n_meshes);
    std::vector<Box> obox(n_meshes);
    int dbot = 1;
    for (uint32_t m = 0; m < n_meshes; m++) {
        dbot = std::max(dbot, hk_build_bvh8(positions, indices + 3 * (size_t)meshes[m].first_tri, meshes[m].n_tris, blas[m]));
        Box b; b.reset();
        for (uint32_t t = 0; t < 3 * meshes[m].n_tris; t++) b.grow(positions + 3 * (size_t)indices[3 * (size_t)meshes[m].first_tri + t]);
        obox[m] = b;
    }
    std::vector<Box> pb(n_inst); std::vector<float> cen(3 * (size_t)n_inst);
    for (uint32_t i = 0; i < n_inst; i++) {
        const Box& ob = obox[inst[i].mesh]; const float* M = inst[i].object_to_world;
        Box b; b.reset();
        for (int c = 0; c < 8; c++) {
            const float x = (c & 1) ? ob.hi[0] : ob.lo[0], y = (c & 2) ? ob.hi[1] : ob.lo[1], z = (c & 4) ? ob.hi[2] : ob.lo[2];
            const float p[3] = {M[0] * x + M[1] * y + M[2] * z + M[3], M[4] * x + M[5] * y + M[6] * z + M[7], M[8] * x + M[9] * y + M[10] * z + M[11]};
            b.grow(p);
        }
        for (int k = 0; k < 3; k++) {
            cen[3 * (size_t)i + k] = 0.5f * (b.lo[k] + b.hi[k]);
            const float pad = 1.0e-4f * std::max(std::fabs(b.lo[k]), std::fabs(b.hi[k])) + 1.0e-5f;      // culling must only ever be conservative
            b.lo[k] -= pad; b.hi[k] += pad;
        }
        pb[i] = b;
    }
    const int dtop = build_bvh8_core(pb, cen, n_inst, 1, [&](uint32_t prim, HkBvhTri& T) { T.prim = prim; }, out);
    mesh_root.assign(n_meshes, 0);
    for (uint32_t m = 0; m < n_meshes; m++) {
        const uint32_t node_off = (uint32_t)out.nodes.size(), tri_off = (uint32_t)out.tris.size();
        mesh_root[m] = node_off;
        for (HkBvhNode nd : blas[m].nodes) { nd.child_base += node_off; nd.tri_base += tri_off; out.nodes.push_back(nd); }
        out.tris.insert(out.tris.end(), blas[m].tris.begin(), blas[m].tris.end());
        blas[m] = HkBvh();
    }
    if (depth_top) *depth_top = dtop;
    if (depth_bottom) *depth_bottom = dbot;
    return dtop + dbot;
}
// Host-only entry point (no GPU needed): build the BVH8 of a triangle soup and hand back its node / triangle arrays, so the
// builder's invariants can be tested on the CPU (tests/test_host_logic.py).  Returns the sizes; copies only when they fit.
extern "C" int32_t hk_host_build_bvh8(const float* positions, const uint32_t* indices, uint32_t n_tris, void* out_nodes, uint64_t nodes_cap,
                                      void* out_tris, uint64_t tris_cap, uint64_t* n_nodes, uint64_t* n_out_tris) {
    if (!positions || !indices || !n_nodes || !n_out_tris) return -1;
    HkBvh bvh;
    hk_build_bvh8(positions, indices, n_tris, bvh);
    *n_nodes = bvh.nodes.size(); *n_out_tris = bvh.tris.size();
    if (out_nodes && nodes_cap >= bvh.nodes.size()) std::memcpy(out_nodes, bvh.nodes.data(), bvh.nodes.size() * sizeof(HkBvhNode));
    if (out_tris && tris_cap >= bvh.tris.size()) std::memcpy(out_tris, bvh.tris.data(), bvh.tris.size() * sizeof(HkBvhTri));
    return 0;
}
