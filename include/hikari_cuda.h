/*
 * hikari_cuda.h — C ABI of libhikari_cuda.so, the sm_100a back end of Hikari.jl's VolPath
 * wavefront integrator.
 *
 * The reference has no FFI seam for integrators: the seam is Julia multiple dispatch on
 *   render!(vp::VolPath, scene, film, camera)          src/integrators/volpath/volpath.jl:445-636
 *   (vp::VolPath)(scene, film, camera)                 src/integrators/volpath/volpath.jl:655-670
 *   clear!(vp::VolPath)                                src/integrators/volpath/volpath.jl:108-113
 * A Julia shim overloads those methods and ccalls the entry points below.
 * Every entry point cites the reference interface it replaces.
 *
 * Conventions
 *  - all functions return 0 on success, a negative HK_ERR_* code otherwise; hk_last_error()
 *    gives a human-readable message for the last failure on that context.
 *  - every pointer argument is a HOST pointer borrowed for the duration of the call unless the
 *    name ends in _dev.  The library owns all device memory behind the opaque HkContext.
 *  - all multi-dimensional arrays are dense C (row-major) arrays with the documented shape; all
 *    indices stored INSIDE arrays keep the reference's 1-based convention (0 = "none").
 *  - matrices are 4x4 row-major (m[r*4+c]); points transform as M*(x,y,z,1) with a divide by w
 *    when w != 1, vectors by the upper 3x3 (Raycore.Transformation semantics).
 *  - a context is bound to one CUDA device and is not thread-safe; drive one context per GPU.
 */
#ifndef HIKARI_CUDA_H
#define HIKARI_CUDA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HK_ABI_VERSION 4

/* ---- status codes ------------------------------------------------------------------------ */
#define HK_OK                 0
#define HK_ERR_INVALID       -1   /* bad argument / missing prerequisite upload            */
#define HK_ERR_CUDA          -2   /* CUDA runtime error, see hk_last_error                */
#define HK_ERR_NO_DEVICE     -3   /* no usable CUDA device: there is NO CPU fallback      */
#define HK_ERR_UNSUPPORTED   -4   /* feature outside the implemented hot-path scope       */
#define HK_ERR_OOM           -5

typedef struct HkContext HkContext;

/* ---- constant tables ----------------------------------------------------------------------
 * replaces: SobolRNG upload (src/sampler/sobol.jl:370-379), CIEXYZTable (src/spectral/color.jl:28-32),
 * D65 table (src/spectral/uplift.jl:412-429) and RGBToSpectrumTable (src/spectral/rgb2spec.jl:71-75)
 * allocation in VolPathState (src/integrators/volpath/volpath-state.jl:98-181).               */
typedef struct HkTables {
    const uint32_t* sobol_matrices;   /* [1024*52]                                            */
    const float*    cie_x;            /* [471] 360..830 nm, 1 nm                              */
    const float*    cie_y;            /* [471]                                                */
    const float*    cie_z;            /* [471]                                                */
    const float*    d65;              /* [107] 300..830 nm, 5 nm                              */
    int32_t         rgb2spec_res;     /* 64                                                   */
    const float*    rgb2spec_scale;   /* [res]                                                */
    const float*    rgb2spec_coeffs;  /* [3][res][res][res][3] = [coef][x][y][z][maxc] i.e. the
                                         raw memory of the Julia Array{Float32,5}(3,res,res,res,3)
                                         indexed [maxc, z, y, x, coef] (column-major)          */
} HkTables;

/* ---- geometry -----------------------------------------------------------------------------
 * replaces: scene.accel (Raycore TLAS) as consumed by Raycore.closest_hit at
 * src/integrators/volpath/intersection.jl:200,225,323,703 and the primitive fields read at
 * intersection.jl:14-16,30-32,86-88,130-132 (vertices/normals/tangents/uv/metadata).
 * World-space triangle soup; the global primitive id is the triangle's index here (instance-major,
 * then face order) and is what hk_trace_closest reports.                                      */
/* Instancing (replaces: the instances of the Raycore TLAS behind scene.accel, src/scene.jl:21-28, 146-151; one per
 * push!(scene, mesh, material; transform), src/scene-mesh.jl:9-16).  With n_instances > 0 the vertex arrays are in OBJECT space,
 * `meshes` cuts `indices` into meshes, and every instance places one mesh with an affine transform and one medium interface.
 * The library builds one bottom-level BVH per mesh and a top-level BVH over the instances; nothing is flattened.
 * Global primitive id of (instance i, face f of its mesh) = sum of the face counts of instances 0..i-1, + f (instance-major, then
 * face order: the order the reference's flattened TLAS enumerates), and TriangleMeta of that primitive is
 * (instances[i].medium_interface_idx, f + 1, 0): instanced meshes cannot be area lights (pass emissive meshes un-instanced).
 * Closest-hit contract for instances (the reference's Raycore is not on disk; stated here, restated in oracle/ok_accel.h): the ray is
 * taken to object space as o' = W o, d' = W d (W = world_to_object, d' NOT renormalised, so t means the same on both sides),
 * tested against the mesh's object-space triangles with the fixed Moller-Trumbore sequence; argmin t over all (instance, face)
 * with 0 < t < t_max, equal t -> smallest global primitive id.  World-space shading data (vertices, normals) are O v and
 * normalize(W^T n) evaluated in f32 in the operation order of csrc/hk_wavefront.cuh::prim_vertices.                      */
typedef struct HkMesh { uint32_t first_tri, n_tris; } HkMesh;            /* a range of `indices` */
typedef struct HkInstance {
    uint32_t mesh;                    /* 0-based index into meshes                                          */
    uint32_t medium_interface_idx;    /* 1-based, as TriangleMeta.medium_interface_idx                      */
    float    object_to_world[12];     /* 3x4 row-major affine O                                            */
    float    world_to_object[12];     /* 3x4 row-major affine W = O^-1                                     */
} HkInstance;

typedef struct HkGeometry {
    const float*    positions;   /* [n_verts][3]                                              */
    const float*    normals;     /* [n_verts][3] or NULL; NaN x-component = "no normal"        */
    const float*    tangents;    /* [n_verts][3] or NULL; NaN x-component = "no tangent"       */
    const float*    uvs;         /* [n_verts][2] or NULL (treated as 0)                        */
    const uint32_t* indices;     /* [n_tris][3]  0-based vertex indices                        */
    const uint32_t* tri_meta;    /* [n_tris][3]  TriangleMeta (src/scene.jl:11-15):
                                    medium_interface_idx (1-based), primitive_index (1-based face
                                    index within its mesh), arealight_flat_idx (0 = none);
                                    ignored (may be NULL) when n_instances > 0                  */
    uint32_t        n_verts;
    uint32_t        n_tris;
    const HkMesh*     meshes;      /* [n_meshes]     (n_instances > 0 only)                    */
    uint32_t          n_meshes;
    const HkInstance* instances;   /* [n_instances]; 0 = plain world-space triangle soup        */
    uint32_t          n_instances;
} HkGeometry;

/* ---- textures -----------------------------------------------------------------------------
 * replaces: the texture arrays a Raycore.TextureRef points at (src/textures/texture-ref.jl:50-76).  The reference's RGBSpectrum
 * texels are r, g, b, alpha (spectrum.jl:62-70); the caller splits them into a packed RGB image and an optional alpha plane, both
 * in the reference's memory order, an (h, w) column-major matrix: texel (row y, column x), 0-based, at rgb[3 * (x * h + y)] and
 * alpha[x * h + y].
 * Colours are sampled bilinearly at the hit's interpolated uv exactly as _sample_texture_bilinear (:160-190): px = u (w-1) + 1,
 * py = (1 - v)(h-1) + 1, indices clamped, no wrap.  alpha (null = opaque everywhere) makes the surfaces of a MatteMaterial whose Kd is
 * this texture alpha-tested: get_surface_alpha (spectral-eval.jl:3882-3888) point-samples it (_sample_texture_data,
 * textures/basic.jl:19-25) and the trace / shadow stages skip the surface when pcg32(hash(o), hash(d)) > alpha, without consuming
 * depth (intersection.jl:221-266, 349-372).  A VertexColorTexture (HK_MATFLAG_VERTEX_COLORS) is passed as h = 3, w = n_faces.
 * Upload before the materials that reference them.                                                              */
typedef struct HkTexture {
    const float* rgb;
    int32_t      h, w;
    const float* alpha;
} HkTexture;

/* ---- materials ----------------------------------------------------------------------------
 * replaces: scene.materials (MultiTypeSet) + scene.media_interfaces (src/scene.jl:21-28,
 * src/materials/medium-interface.jl:78-82).  Every RGB and scalar parameter is a constant or a texture (tex[] / ftex[]);
 * MatteMaterial.Kd may also carry alpha or be a VertexColorTexture; MixMaterial.amount = f0 / ftex[0].  Not texturable: the
 * integer parameters, piecewise-linear eta / k spectra.                                                                */
#define HK_MAT_MATTE                1   /* src/materials/spectral-eval.jl:42-101, 371-397      */
#define HK_MAT_MIRROR               2   /* :108-132                                            */
#define HK_MAT_GLASS                3   /* :140-198, 407-413                                   */
#define HK_MAT_CONDUCTOR            4   /* :223-318, 421-488                                   */
#define HK_MAT_COATED_DIFFUSE       5   /* :1232-1937                                          */
#define HK_MAT_THIN_DIELECTRIC      6   /* :1975-2051                                          */
#define HK_MAT_DIFFUSE_TRANSMISSION 7   /* :2083-2218                                          */
#define HK_MAT_MIX                  8   /* src/materials/mix-material.jl: resolved to one of its two sub-materials at
                                           intersection time (resolve_mix_material :253-268), never shaded itself   */
#define HK_MAT_COATED_CONDUCTOR     9   /* src/materials/spectral-eval.jl:2877-3237 (sample), 3243-3418 (eval);
                                           src/materials/coated-conductor.jl:48-105                                 */
#define HK_MAT_COATED_DIFFUSE_TRANSMISSION 10 /* :2340-2494 (sample), 2498-2763 (eval), 2767-2840 (pdf);
                                           src/materials/coated-diffuse-transmission.jl                             */

#define HK_MATFLAG_REMAP_ROUGHNESS  1u
#define HK_MATFLAG_SPECTRAL_ETA_K   2u  /* conductor eta/k are piecewise-linear spectra (ids in spec[]) */
#define HK_MATFLAG_VERTEX_COLORS    8u  /* tex[0] is a VertexColorTexture (src/textures/basic.jl:43-46, texture-ref.jl:240-245): a (3, n_faces)
                                           table of per-face corner colours = an HkTexture with h = 3, w = n_faces, evaluated as
                                           sum_k data[k, face] * bary[k] with face = TriangleMeta.primitive_index (MatteMaterial.Kd) */
#define HK_MATFLAG_USE_ETA_K        4u  /* CoatedConductor: rgb0 / rgb1 (or spec[]) are eta / k; clear = rgb0 is the
                                           artist reflectance (use_eta_k, coated-conductor.jl:98)                   */

typedef struct HkMaterial {
    int32_t  type;
    uint32_t flags;
    float    rgb0[3];   /* Matte Kd | Mirror Kr | Glass Kr | Conductor eta | Coated reflectance | DiffTrans reflectance */
    float    rgb1[3];   /* Glass Kt | Conductor k | Coated albedo | DiffTrans transmittance                            */
    float    rgb2[4];   /* CoatedConductor: rgb0 = conductor eta (or reflectance), rgb1 = conductor k, rgb2 = albedo.
                           CoatedDiffuseTransmission: rgb0 = reflectance, rgb1 = albedo, rgb2 = transmittance; f[] and
                           ival[] as CoatedDiffuse                                                                     */
    float    f[8];      /* Matte: f0=sigma. Glass: f0=index. Conductor: f0=roughness.
                           ThinDielectric: f0=eta. DiffuseTransmission: f0=scale.
                           CoatedDiffuse: f0=u_roughness f1=v_roughness f2=thickness f3=eta f4=g
                           CoatedConductor: f0/f1=interface u/v roughness f2=thickness f3=interface_eta f4=g
                                            f5/f6=conductor u/v roughness                              */
    int32_t  spec[2];   /* 1-based ids into the uploaded piecewise-linear spectra (eta, k); 0 = unused  */
    int32_t  ival[2];   /* CoatedDiffuse / CoatedConductor: ival0=max_depth ival1=n_samples             */
                        /* Mix: f0 / ftex[0] = amount (constant or texture, mix-material.jl:183); ival0 / ival1 = 1-based material1 / material2;
                           the SetKeys hashed by mix_hash_float (mix-material.jl:114-158): spec0 / spec1 = vec_idx of
                           material1 / material2, flags = type_idx1 | type_idx2 << 8                             */
    int32_t  tex[4];    /* 1-based ids into the uploaded textures replacing rgb0 / rgb1 / rgb2 (0 = the constant; tex[3] unused):
                           eval_tex(textures, mat.<param>, tfc) of the reference (spectral-eval.jl, texture-ref.jl:72-80) =
                           the bilinear texel at the hit's uv, substituted for the constant before anything else is done
                           with it.  HK_MATFLAG_VERTEX_COLORS: tex[0] of a MatteMaterial is a VertexColorTexture          */
    int32_t  ftex[8];   /* the same for the scalar parameters f[0..7] (sigma, roughness, thickness, g, index, ...): the first
                           channel of the texture (a Texture{Float32} is uploaded as r = g = b)                            */
} HkMaterial;

typedef struct HkMediumInterface {   /* MediumInterfaceIdx, src/materials/medium-interface.jl:78-82 */
    uint32_t material;  /* 1-based index into materials                                          */
    uint32_t inside;    /* 1-based medium index, 0 = vacuum                                      */
    uint32_t outside;
} HkMediumInterface;

/* piecewise-linear spectra (src/spectral/piecewise-linear.jl:4-7): spectrum i (1-based) occupies
 * lambdas/values[offsets[i-1] .. offsets[i])                                                     */
typedef struct HkSpectra {
    const float*    lambdas;
    const float*    values;
    const uint32_t* offsets;   /* [n_spectra+1] */
    uint32_t        n_spectra;
} HkSpectra;

/* ---- lights -------------------------------------------------------------------------------
 * replaces: scene.lights (MultiTypeSet) in FLAT index order (src/lights/light-sampler.jl:289-329)
 * and the BVHLightSampler arrays (src/lights/bvh-light-sampler.jl:269-275).                     */
#define HK_LIGHT_POINT        1   /* src/integrators/physical-wavefront/lights.jl:39-59   */
#define HK_LIGHT_SPOT         2   /* :66-105                                              */
#define HK_LIGHT_DIRECTIONAL  3   /* :112-128                                             */
#define HK_LIGHT_SUN          4   /* :135-150                                             */
#define HK_LIGHT_ENVIRONMENT  5   /* :158-189, 408-419                                    */
#define HK_LIGHT_AMBIENT      6   /* :199-221, 427-433                                    */
#define HK_LIGHT_DIFFUSE_AREA 7   /* :235-290, src/lights/diffuse-area.jl:54-81           */

#define HK_SPECTRUM_RGB        0  /* RGBSpectrum: uplifted at evaluation time                     */
#define HK_SPECTRUM_ILLUMINANT 1  /* RGBIlluminantSpectrum: poly + scale baked (rgb2spec.jl:331-334) */

typedef struct HkLight {
    int32_t type;
    int32_t spectrum_kind;
    float   scale;              /* light.scale (photometric normalisation); area light: scale   */
    float   rgb[3];             /* light.i.c (HK_SPECTRUM_RGB) or area light Le                  */
    float   poly[3];            /* HK_SPECTRUM_ILLUMINANT: c0,c1,c2                              */
    float   illum_scale;        /* HK_SPECTRUM_ILLUMINANT: s.scale                               */
    float   position[3];        /* point / spot                                                  */
    float   direction[3];       /* directional / sun: normalised travel direction                */
    float   cos_total_width;    /* spot                                                          */
    float   cos_falloff_start;  /* spot                                                          */
    float   world_to_light[16]; /* spot                                                          */
    float   v[9];               /* area light triangle vertices                                  */
    float   normal[3];          /* area light geometric normal                                   */
    float   area;
    float   uv[6];
    int32_t two_sided;
    int32_t env_map;            /* environment: 1-based id of the uploaded env map               */
} HkLight;

typedef struct HkEnvMap {   /* src/textures/environment_map.jl:9-45 + Distribution2D sampling.jl:179-193 */
    const float* rgb;                   /* [h][w][3]                                             */
    int32_t      w, h;
    float        rotation[9];           /* Mat3f column-major as stored by Julia                 */
    float        scale_rgb[3];          /* EnvironmentLight.scale (src/lights/environment.jl:10) */
    const float* conditional_func;      /* [nv][nu]                                              */
    const float* conditional_cdf;       /* [nv][nu+1]                                            */
    const float* conditional_func_int;  /* [nv]                                                  */
    const float* marginal_func;         /* [nv]                                                  */
    const float* marginal_cdf;          /* [nv+1]                                                */
    float        marginal_func_int;
    int32_t      nu, nv;
} HkEnvMap;

typedef struct HkLightBVHNode {  /* LightBVHNode, src/lights/bvh-light-sampler.jl:26-38 */
    float    bounds_min[3];
    float    bounds_max[3];
    float    w[3];
    float    phi;
    float    cos_theta_o;
    float    cos_theta_e;
    uint32_t two_sided;
    uint32_t child1_or_light_idx;   /* interior: 1-based index of child1 (child0 = self+1); leaf: flat light idx */
    uint32_t is_leaf;
    uint32_t _pad;
} HkLightBVHNode;

typedef struct HkLightSampler {
    const HkLightBVHNode* nodes;            uint32_t n_nodes;
    const uint32_t* light_to_bit_trail;     /* [n_lights], 0xFFFFFFFF = infinite-light sentinel   */
    const int32_t*  infinite_light_indices; uint32_t n_infinite;
    uint32_t        n_bvh_lights;
} HkLightSampler;

/* ---- media (src/integrators/volpath/media.jl, nanovdb.jl) ----------------------------------- */
#define HK_MEDIUM_HOMOGENEOUS 1   /* media.jl:735-793   */
#define HK_MEDIUM_GRID        2   /* media.jl:800-1000, 1544-1623 */
#define HK_MEDIUM_NANOVDB     3   /* nanovdb.jl:153-191, 315-543   */
#define HK_MEDIUM_RGBGRID     4   /* media.jl:1002-1456: per-voxel RGB sigma_a / sigma_s / Le     */

typedef struct HkMedium {
    int32_t  type;
    float    sigma_a_rgb[3];
    float    sigma_s_rgb[3];
    float    Le_rgb[3];
    float    scale;              /* density scale                                           */
    float    g;                  /* Henyey-Greenstein asymmetry                              */
    float    bounds_min[3];
    float    bounds_max[3];
    float    render_from_medium[16];
    float    medium_from_render[16];
    int32_t  density_res[3];     /* Grid: nx,ny,nz                                          */
    const float* density;        /* Grid: [nz][ny][nx]                                      */
    int32_t  majorant_res[3];
    const float* majorant;       /* [rz][ry][rx] max density per coarse voxel               */
    const uint8_t* nanovdb_buf;  /* NanoVDB: raw grid buffer                                */
    uint64_t nanovdb_bytes;
    float    nanovdb_inv_mat[9]; /* index-from-world 3x3, row-major                         */
    float    nanovdb_vec[3];     /* translation                                             */
    uint64_t nanovdb_root_offset, nanovdb_upper_offset, nanovdb_lower_offset, nanovdb_leaf_offset;
    int32_t  nanovdb_root_tiles, nanovdb_upper_count, nanovdb_lower_count, nanovdb_leaf_count;
    /* RGBGrid: density_res = grid_res; grids [nz][ny][nx][3] (RGBSpectrum per voxel), NULL = absent (sigma_a / sigma_s
       default to 1, Le to 0, media.jl:1252-1273); scale = sigma_scale; majorant = build_rgb_majorant_grid (:1122-1183),
       i.e. sigma_scale * (max sigma_a + max sigma_s) per coarse voxel; bounds and transforms as Grid             */
    const float* rgb_sigma_a;
    const float* rgb_sigma_s;
    const float* rgb_Le;
    float    Le_scale;
    /* ABI v4.  majorant == NULL (Grid / RGBGrid / NanoVDB): the library builds the majorant grid on the device from the
       uploaded voxels -- build_majorant_grid (media.jl:1459-1496), build_rgb_majorant_grid (:1123-1183),
       build_nanovdb_majorant_grid (nanovdb.jl:1174-1235); the NanoVDB build clips its voxel boxes to
       [nanovdb_index_min, nanovdb_index_max] = metadata.index_min / index_max (nanovdb.jl:1137-1150, :853).        */
    int32_t  nanovdb_index_min[3];
    int32_t  nanovdb_index_max[3];
    /* NanoVDB with nanovdb_buf == NULL: the tree is built on the device from the dense volume in density / density_res
       ([nz][ny][nx], background 0) -- build_nanovdb_from_dense (nanovdb.jl:602-858): same bytes as the host builder's buffer;
       nanovdb_inv_mat / nanovdb_vec are still the caller's, offsets / counts / index range are the library's                */
} HkMedium;

/* ---- camera, filter, params ---------------------------------------------------------------- */
typedef struct HkCamera {       /* PerspectiveCamera, src/camera/perspective.jl:1-80 */
    float raster_to_camera[16];
    float camera_to_world[16];
    float lens_radius;
    float focal_distance;
    float shutter_open;
    float shutter_close;
    float dx_camera[3];
    float dy_camera[3];
} HkCamera;

typedef struct HkFilter {       /* GPUFilterParams + GPUFilterSamplerData, src/filter.jl:574-725 */
    int32_t      type;          /* 1 Box, 2 Triangle, 3 Gaussian, 4 Mitchell, 5 Lanczos          */
    float        radius[2];
    int32_t      nx, ny;        /* tabulated sampler (types 3..5); 0 for Box/Triangle            */
    const float* func;              /* [ny][nx]                                                  */
    const float* marginal_cdf;      /* [ny+1]                                                    */
    const float* marginal_func;     /* [ny]                                                      */
    const float* conditional_cdf;   /* [ny][nx+1]                                                */
    float        domain_min[2];
    float        domain_max[2];
    float        func_integral;
} HkFilter;

typedef struct HkRenderParams {  /* VolPath fields, src/integrators/volpath/volpath.jl:29-42,75-101 */
    int32_t  width, height;
    int32_t  max_depth;
    int32_t  samples_per_pixel;      /* vp.samples_per_pixel (only used for texture footprints)   */
    int32_t  regularize;
    float    max_component_value;
    uint32_t sampler_seed;           /* 0 in the reference (volpath.jl:480)                       */
    int32_t  sobol_log2_spp;         /* compute_zsobol_params(max(spp,4096),W,H) sobol.jl:317-323 */
    int32_t  sobol_n_base4_digits;
    int32_t  material_coherence;     /* 0 :none, 1 :sorted, 2 :per_type — all map to per-type queues */
    int32_t  sample_batch;           /* library extension: samples in flight per pass; <= 0 = auto
                                      * (~32 M path states: 16 at 1080p, 4 at 4K).  Images are bitwise
                                      * independent of it.                                          */
} HkRenderParams;

typedef struct HkStats {
    uint64_t rays_traced;        /* closest-hit queries: primary + continuation + every shadow segment */
    uint64_t samples_rendered;   /* pixel samples accumulated                                          */
    uint64_t queue_overflows;    /* always 0: queues are sized to the slot count                       */
    uint64_t bvh_nodes;          /* number of 80-byte wide-BVH nodes                                   */
    uint64_t bvh_bytes;          /* node + triangle bytes resident                                     */
    uint64_t kernel_launches;    /* kernels launched by this context so far                           */
    float    last_render_ms;     /* device time of the last hk_render_samples (CUDA events)            */
    float    last_trace_ms;      /* device time of the last hk_trace_closest kernel                    */
    uint64_t path_vertices;      /* surface hits + medium scattering events shaded so far (the unit of the
                                  * reference's per-vertex queue traffic)                   */
} HkStats;

/* ---- lifecycle ---------------------------------------------------------------------------- */
int32_t hk_abi_version(void);
/* replaces: VolPathState(backend, ...) allocation, volpath-state.jl:98-181 */
int32_t hk_create(int32_t device, HkContext** out_ctx);
/* replaces: cleanup!(state) volpath-state.jl:238-273 / Base.close(::Integrator) Hikari.jl:47 */
int32_t hk_destroy(HkContext* ctx);
const char* hk_last_error(HkContext* ctx);

/* ---- scene upload (replaces Adapt.adapt(backend, scene), volpath.jl:455-461) ---------------- */
int32_t hk_upload_tables(HkContext* ctx, const HkTables* tables);
int32_t hk_upload_geometry(HkContext* ctx, const HkGeometry* geom);
int32_t hk_upload_spectra(HkContext* ctx, const HkSpectra* spectra);
int32_t hk_upload_textures(HkContext* ctx, const HkTexture* textures, uint32_t n_textures);
int32_t hk_upload_materials(HkContext* ctx, const HkMaterial* materials, uint32_t n_materials,
                            const HkMediumInterface* interfaces, uint32_t n_interfaces);
/* update_material!(scene, idx, new_material), src/scene.jl:109-112: replace one uploaded material in place (1-based
 * index into the material array); the scene is not re-uploaded.  A change of material type re-tags the BVH triangles. */
int32_t hk_update_material(HkContext* ctx, uint32_t index, const HkMaterial* material);
int32_t hk_upload_envmaps(HkContext* ctx, const HkEnvMap* maps, uint32_t n_maps);
int32_t hk_upload_lights(HkContext* ctx, const HkLight* lights, uint32_t n_lights,
                         const HkLightSampler* sampler);
int32_t hk_upload_media(HkContext* ctx, const HkMedium* media, uint32_t n_media);
/* build_majorant_grid! / build_rgb_majorant_grid! (media.jl:1498-1530, 1185-1240): the host swapped the voxel data of ONE
 * medium (1-based index).  The record replaces that medium wholesale, its majorant grid (majorant == NULL) and empty-cell
 * mask are rebuilt on the device; geometry, materials, lights and the other media stay.                                */
int32_t hk_update_medium(HkContext* ctx, uint32_t index, const HkMedium* medium);
/* the NanoVDB buffer of medium `index` as the device holds it (uploaded, or built on the device from a dense volume): *bytes = its
 * size; copied to `out` when out != NULL and capacity >= *bytes (e.g. to write a .nvdb file)                                */
int32_t hk_read_nanovdb(HkContext* ctx, uint32_t index, uint8_t* out, uint64_t capacity, uint64_t* bytes);
/* the majorant grid of medium `index` as the device holds it ([rz][ry][rx], n_cells = rx*ry*rz): uploaded or device-built */
int32_t hk_read_majorant(HkContext* ctx, uint32_t index, float* out, uint64_t n_cells);
int32_t hk_set_camera(HkContext* ctx, const HkCamera* camera);
int32_t hk_set_filter(HkContext* ctx, const HkFilter* filter);
/* replaces: VolPath(...) fields + state (re)allocation on resize, volpath.jl:466-482 */
int32_t hk_set_params(HkContext* ctx, const HkRenderParams* params);

/* ---- rendering ---------------------------------------------------------------------------- */
/* replaces: clear!(vp) volpath.jl:108-113 */
int32_t hk_clear(HkContext* ctx);
/* replaces: `count` consecutive render!(vp, scene, film, camera) calls (volpath.jl:445-636) for
 * sample indices first_sample_idx .. first_sample_idx+count-1 (1-based, = film.iteration_index+1). */
int32_t hk_render_samples(HkContext* ctx, int32_t first_sample_idx, int32_t count);
/* same, strided: sample indices first, first+stride, ... (count of them) — the multi-GPU partition */
int32_t hk_render_samples_strided(HkContext* ctx, int32_t first_sample_idx, int32_t stride, int32_t count);
/* replaces: vp_finalize_film_kernel! + host read of film.framebuffer (volpath.jl:384-417).
 * Writes RGB{Float32} in the reference's (H, W) column-major layout: out[((px-1)*H + (py-1))*3 + c]. */
int32_t hk_read_film(HkContext* ctx, float* out_rgb_hw_colmajor);
/* Pipelined read-out for progressive display (library extension): finalize + device->host copy of the current film are
 * enqueued (copy on its own stream) and the call returns a ticket at once, so the next hk_render_samples overlaps the
 * DMA; hk_read_film_wait blocks until that frame is in `out` (page-locked memory: hk_pinned_alloc).  Four frames
 * at most in flight (a fifth call reuses the oldest ticket's staging, after its copy); one host buffer per frame in flight. */
int32_t hk_read_film_async(HkContext* ctx, float* out_rgb_hw_colmajor_pinned, int32_t* out_ticket);
int32_t hk_read_film_wait(HkContext* ctx, int32_t ticket);
/* Zero-copy read-out: the reference takes its GPU path when film.framebuffer is a device array
 * (KA.get_backend(film.framebuffer), volpath.jl:453) and vp_finalize_film_kernel! writes straight into it (:384-417).
 * out_dev is a DEVICE pointer to H*W RGB{Float32} in the same (H, W) column-major layout (a CuArray's pointer); the
 * finalize kernel writes it on the render stream and the call returns without a host synchronisation
 * (hk_synchronize, or stream order on a caller-supplied stream, orders later reads).                              */
int32_t hk_read_film_dev(HkContext* ctx, float* out_dev_rgb_hw_colmajor);
/* All rendering, read-out and upload-side kernels of `ctx` are enqueued on ONE stream.  By default the library owns it;
 * hk_set_stream makes it the caller's (a cudaStream_t passed as void*; NULL = back to the library's own stream), which is
 * how the Julia host orders hk_render_samples / hk_read_film_dev against its own CUDA.jl work without device-wide
 * synchronisation.  The previous stream is drained first.  The stream must outlive the context or be reset first.  */
int32_t hk_set_stream(HkContext* ctx, void* cuda_stream);

/* postprocess!(film; exposure, tonemap, gamma, white_point, sensor), src/postprocess.jl:187-357 -- fused into the film
 * read-out: framebuffer = sum / weight, then exposure, Bradford white balance, sensor imaging ratio, tone map, gamma.
 * Same (H, W) column-major RGB layout as hk_read_film.  mask_escaped != 0 blends `background` over pixels whose
 * film.depth is infinite (postprocess.jl:220-245, 3x3 anti-aliased); it needs hk_fill_aux_buffers first.            */
#define HK_TONEMAP_NONE         0   /* linear clamp                          postprocess.jl:160-182 */
#define HK_TONEMAP_REINHARD     1
#define HK_TONEMAP_REINHARD_EXT 2
#define HK_TONEMAP_ACES         3
#define HK_TONEMAP_UNCHARTED2   4
#define HK_TONEMAP_FILMIC       5
typedef struct HkPostprocess {
    float   exposure;
    int32_t tonemap_mode;
    float   inv_gamma;      /* 1 / gamma                                                             */
    int32_t apply_gamma;    /* 0: gamma = nothing                                                    */
    float   white_point;    /* extended Reinhard                                                     */
    float   imaging_ratio;  /* sensor.exposure_time * sensor.iso / 100                               */
    int32_t apply_wb;       /* sensor.white_balance > 0                                              */
    float   wb[9];          /* compute_white_balance_matrix(T), row-major (spectral/color.jl:522-546) */
    int32_t mask_escaped;   /* background !== nothing (postprocess.jl:339): blend towards `background` by the fraction
                               of escaped (depth = Inf) pixels in the 3x3 neighbourhood of the depth buffer
                               (:220-245); needs hk_fill_aux_buffers first                            */
    float   background[3];
} HkPostprocess;
int32_t hk_postprocess(HkContext* ctx, const HkPostprocess* params, float* out_rgb_hw_colmajor);
/* same into a DEVICE buffer (film.postprocess as a device array), no host synchronisation */
int32_t hk_postprocess_dev(HkContext* ctx, const HkPostprocess* params, float* out_dev_rgb_hw_colmajor);
/* ---- auxiliary buffers ---------------------------------------------------------------------
 * replaces: fill_aux_buffers!(film, scene, camera; has_infinite_lights) (src/film.jl:410-431, kernel :433-488):
 * one primary ray per pixel through the pixel centre (lens sample (0.5, 0.5)); writes film.albedo (0.8 on a hit,
 * 0 otherwise), film.normal (geometric normal on the shading-normal side, as vp_compute_surface_geometry
 * intersection.jl:13-21,176; 0 on a miss) and film.depth (|hit - ray.o|; on a miss Inf, or 1e30 when
 * has_infinite_lights).  All three are (H, W) column-major like the framebuffer; hk_clear zeroes them (film.jl:343-346).
 * hk_read_aux_buffers: any pointer may be NULL.                                                                  */
int32_t hk_fill_aux_buffers(HkContext* ctx, int32_t has_infinite_lights);
int32_t hk_read_aux_buffers(HkContext* ctx, float* albedo_hw3, float* normal_hw3, float* depth_hw);
/* ---- denoise -------------------------------------------------------------------------------
 * replaces: denoise!(film; config) (src/denoise.jl:301-372): 3x3 luminance variance of film.framebuffer when use_variance,
 * then `iterations` edge-avoiding a-trous passes (step 1, 2, 4, ...) guided by film.normal / film.depth
 * (hk_fill_aux_buffers first), result into film.postprocess.  The reference ping-pongs between film.framebuffer itself
 * and a scratch copy, so from two iterations on it leaves the last even pass in film.framebuffer: pass
 * out_framebuffer_hw3 (may be NULL) to receive that side effect (written only when iterations >= 2).           */
typedef struct HkDenoiseConfig {     /* DenoiseConfig, denoise.jl:29-57; defaults 5, 4.0, 128.0, 1.0, true */
    int32_t iterations;
    float   sigma_color, sigma_normal, sigma_depth;
    int32_t use_variance;
} HkDenoiseConfig;
int32_t hk_denoise(HkContext* ctx, const HkDenoiseConfig* config, float* out_postprocess_hw3, float* out_framebuffer_hw3);
/* raw accumulators for the multi-GPU film reduce (pixel_rgb ‖ pixel_weight_sum, volpath-state.jl):
 * device pointer to [n*3] rgb sums followed by [n] weight sums, valid until the next hk_set_params. */
int32_t hk_film_accum_dev(HkContext* ctx, float** out_accum_dev, uint64_t* out_count);
/* host copies of the accumulators (rgb [n][3], weight [n]) */
int32_t hk_read_accum(HkContext* ctx, float* out_rgb_sum, float* out_weight_sum);
int32_t hk_write_accum(HkContext* ctx, const float* rgb_sum, const float* weight_sum);

/* ---- stand-alone traversal (parity + roofline measurement) --------------------------------
 * replaces: Raycore.closest_hit(accel, ray) (call sites intersection.jl:200,225,323,703).
 * rays: [n][8] = o.xyz, d.xyz, t_max, time.  hits: [n][4] = t (f32), prim (u32 bits, 1-based global
 * primitive id, 0 = miss), b1, b2 (barycentrics of v1, v2; b0 = 1-b1-b2).
 * Tie-break: smallest t wins; equal t -> smallest primitive id.                */
int32_t hk_trace_closest(HkContext* ctx, const float* rays, uint64_t n, float* hits);
/* same on device-resident buffers (timed path of bench.py); repeat>1 re-runs the kernel */
int32_t hk_trace_closest_dev(HkContext* ctx, const float* rays_dev, uint64_t n, float* hits_dev, int32_t repeat);
/* any-hit: occluded[i] = 1 if any triangle has 0 < t < t_max */
int32_t hk_trace_any(HkContext* ctx, const float* rays, uint64_t n, uint8_t* occluded);

int32_t hk_stats(HkContext* ctx, HkStats* out);

/* per-stage device timing (CUDA events on the launching stream) and traversal work counters, for the roofline
 * line of bench.py.  mode bit 0: time every stage launch; bit 1: count node visits / triangle tests.
 * hk_stage_times: out_ms[HK_N_STAGES], out_launches[HK_N_STAGES], out_work[6] = closest-hit {rays, node visits,
 * triangle tests}, shadow {rays, node visits, triangle tests} accumulated since hk_set_profiling.          */
#define HK_STAGE_CAMERA  0
#define HK_STAGE_TRACE   1
#define HK_STAGE_MEDIUM  2
#define HK_STAGE_ESCAPED 3
#define HK_STAGE_SHADE   4
#define HK_STAGE_SHADOW  5
#define HK_STAGE_FILM    6
#define HK_STAGE_ROUTE   7
#define HK_N_STAGES      8
int32_t hk_set_profiling(HkContext* ctx, int32_t mode);
int32_t hk_stage_times(HkContext* ctx, double* out_ms, uint64_t* out_launches, uint64_t* out_work);
/* mode 5 additionally records, per bounce of the most recent sample pass, the 16 queue counters and the stage times
 * (one host sync per bounce: diagnostics only).  counts[max_depth][16], ms[max_depth][HK_N_STAGES]. */
int32_t hk_bounce_profile(HkContext* ctx, int32_t max_depth, uint32_t* counts, double* ms);
int32_t hk_synchronize(HkContext* ctx);

/* device memory helpers so a host without a CUDA binding (the Python mirror, tests) can keep
 * inputs resident in HBM across timed calls */
int32_t hk_pinned_alloc(uint64_t bytes, void** out_host);   /* page-locked host memory for film.framebuffer */
int32_t hk_pinned_free(void* host);
int32_t hk_dev_alloc(HkContext* ctx, uint64_t bytes, void** out_dev);
int32_t hk_dev_free(HkContext* ctx, void* dev);
int32_t hk_dev_upload(HkContext* ctx, void* dst_dev, const void* src_host, uint64_t bytes);
int32_t hk_dev_download(HkContext* ctx, void* dst_host, const void* src_dev, uint64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* HIKARI_CUDA_H */
