// hk_bvh.h — 8-wide BVH with quantised child bounds (80-byte nodes) + 48-byte triangles.
// Layout follows the compressed wide BVH idea (Ylitie, Karras, Laine 2017): children share one
// origin/exponent frame, child boxes are 8-bit quantised, internal children and leaf triangles of a
// node are stored contiguously so one base index each suffices.  Sized to stay L2-resident for the
// small/medium configs (126 MB L2) and to stream from HBM3e for the 50 M-triangle config.
#pragma once
#include <cstdint>
#include <vector>

struct alignas(16) HkBvhNode {        // 80 bytes = 5 x 16-byte vector loads
    float    p[3];                     // quantisation origin
    uint8_t  e[3];                     // per-axis exponent (biased by 127 like IEEE)
    uint8_t  imask;                    // bit i set <=> child slot i is an internal node
    uint32_t child_base;               // index of the first internal child
    uint32_t tri_base;                 // index of the first triangle referenced by leaf children
    uint32_t trivalid;                 // bits 3s..3s+2 = unary triangle count of the leaf child in slot s (0 for internal / empty
                                       // slots); triangle k of slot s lives at tri_base + popc(trivalid & ((1 << (3s+k)) - 1))
    uint32_t pad;
    uint8_t  qlo[3][8];                // quantised child mins  [axis][slot]
    uint8_t  qhi[3][8];                // quantised child maxs
};
static_assert(sizeof(HkBvhNode) == 80, "node must be 80 bytes");

struct alignas(16) HkBvhTri {          // 48 bytes = 3 x 16-byte vector loads
    float v0[3]; uint32_t prim;        // global primitive id (0-based)
    float e1[3]; uint32_t pad1;
    float e2[3]; uint32_t pad2;
};
static_assert(sizeof(HkBvhTri) == 48, "triangle must be 48 bytes");

struct HkBvh {
    std::vector<HkBvhNode> nodes;      // nodes[0] = root
    std::vector<HkBvhTri>  tris;       // in leaf order
    float bounds_min[3], bounds_max[3];
};

// positions [n_verts][3], indices [n_tris][3]; returns the depth of the tree in 8-wide levels (bounds the traversal stack)
int hk_build_bvh8(const float* positions, const uint32_t* indices, uint32_t n_tris, HkBvh& out);
// two-level BVH of an instanced scene (hk_bvh.cpp); returns top-level depth + bottom-level depth
struct HkMeshRange { uint32_t first_tri, n_tris; };
struct HkInstanceXf { uint32_t mesh; const float* object_to_world; };
int hk_build_scene_bvh(const float* positions, const uint32_t* indices, const HkMeshRange* meshes, uint32_t n_meshes,
                       const HkInstanceXf* inst, uint32_t n_inst, HkBvh& out, std::vector<uint32_t>& mesh_root, int* depth_top, int* depth_bottom);
