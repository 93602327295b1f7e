# HikariCUDA.jl — the Julia side of the drop-in boundary: methods on Hikari's own generic functions (render!, clear!, the
# callable VolPath, update_material!, postprocess!, fill_aux_buffers!, denoise!) that ccall libhikari_cuda.so
# (include/hikari_cuda.h) when the Film lives on a CUDA device.  Scene / material / light / medium types, push!, sync!,
# Film, PerspectiveCamera and VolPath(samples=…, max_depth=…) are Hikari's, untouched.
#
# STATUS: written against the reference sources at the file:line cited on every function; NOT executed — the build image has
# no Julia (and Raycore.jl, which owns the TLAS, is an un-vendored dependency pinned to a branch, Project.toml:37-38).  The
# executed mirror of exactly this flattening is hikari_jl_b200/host.py (every test drives the C ABI through it) and
# tests/c/abi_client.c (plain C).  The one place that needs Raycore internals is `tlas_arrays` below; everything else only
# touches fields defined in /root/reference/src.
module HikariCUDA

using Hikari, CUDA, StaticArrays
import Hikari: render!, clear!, VolPath, Film, Scene, AbstractScene, Camera, PerspectiveCamera, Material, Light, Medium
import KernelAbstractions as KA

const lib = "libhikari_cuda"                       # on the loader path (hikari_jl_b200/csrc/libhikari_cuda.so)

# The library's frame pipeline needs more hardware work queues than the driver's default of 8 (INTEGRATION.md, "Frame pipelining");
# the variable is read when the CUDA context is created, so this only helps when no context exists yet -- otherwise set it in
# the shell / startup.jl.
__init__() = (get!(ENV, "CUDA_DEVICE_MAX_CONNECTIONS", "32"); nothing)

# ---- struct mirrors of include/hikari_cuda.h (isbits, C layout) ---------------------------------------------------------
const z3_ = (0f0, 0f0, 0f0); const I16 = ntuple(i -> (i - 1) % 5 == 0 ? 1f0 : 0f0, 16)      # zero vector / row-major 4x4 identity
struct HkTables
    sobol_matrices::Ptr{UInt32}; cie_x::Ptr{Float32}; cie_y::Ptr{Float32}; cie_z::Ptr{Float32}; d65::Ptr{Float32}
    rgb2spec_res::Int32; rgb2spec_scale::Ptr{Float32}; rgb2spec_coeffs::Ptr{Float32}
end
struct HkMesh; first_tri::UInt32; n_tris::UInt32; end
struct HkInstance; mesh::UInt32; medium_interface_idx::UInt32; object_to_world::NTuple{12,Float32}; world_to_object::NTuple{12,Float32}; end
struct HkGeometry
    positions::Ptr{Float32}; normals::Ptr{Float32}; tangents::Ptr{Float32}; uvs::Ptr{Float32}; indices::Ptr{UInt32}; tri_meta::Ptr{UInt32}
    n_verts::UInt32; n_tris::UInt32; meshes::Ptr{HkMesh}; n_meshes::UInt32; instances::Ptr{HkInstance}; n_instances::UInt32
end
struct HkTexture; rgb::Ptr{Float32}; h::Int32; w::Int32; alpha::Ptr{Float32}; end
struct HkMaterial
    type::Int32; flags::UInt32; rgb0::NTuple{3,Float32}; rgb1::NTuple{3,Float32}; rgb2::NTuple{4,Float32}
    f::NTuple{8,Float32}; spec::NTuple{2,Int32}; ival::NTuple{2,Int32}; tex::NTuple{4,Int32}; ftex::NTuple{8,Int32}
end
struct HkMediumInterface; material::UInt32; inside::UInt32; outside::UInt32; end
struct HkSpectra; lambdas::Ptr{Float32}; values::Ptr{Float32}; offsets::Ptr{UInt32}; n_spectra::UInt32; end
struct HkLight
    type::Int32; spectrum_kind::Int32; scale::Float32; rgb::NTuple{3,Float32}; poly::NTuple{3,Float32}; illum_scale::Float32
    position::NTuple{3,Float32}; direction::NTuple{3,Float32}; cos_total_width::Float32; cos_falloff_start::Float32
    world_to_light::NTuple{16,Float32}; v::NTuple{9,Float32}; normal::NTuple{3,Float32}; area::Float32; uv::NTuple{6,Float32}
    two_sided::Int32; env_map::Int32
end
struct HkEnvMap
    rgb::Ptr{Float32}; w::Int32; h::Int32; rotation::NTuple{9,Float32}; scale_rgb::NTuple{3,Float32}
    conditional_func::Ptr{Float32}; conditional_cdf::Ptr{Float32}; conditional_func_int::Ptr{Float32}
    marginal_func::Ptr{Float32}; marginal_cdf::Ptr{Float32}; marginal_func_int::Float32; nu::Int32; nv::Int32
end
struct HkLightBVHNode            # 64 bytes; written by bvh_nodes() below from Hikari.LightBVHNode (bvh-light-sampler.jl:21-56)
    bounds_min::NTuple{3,Float32}; bounds_max::NTuple{3,Float32}; w::NTuple{3,Float32}; phi::Float32; cos_theta_o::Float32; cos_theta_e::Float32
    two_sided::UInt32; child1_or_light_idx::UInt32; is_leaf::UInt32; _pad::UInt32
end
struct HkLightSampler
    nodes::Ptr{HkLightBVHNode}; n_nodes::UInt32; light_to_bit_trail::Ptr{UInt32}; infinite_light_indices::Ptr{Int32}
    n_infinite::UInt32; n_bvh_lights::UInt32
end
Base.@kwdef struct HkMedium          # field for field include/hikari_cuda.h (ABI v4, 416 bytes)
    type::Int32; sigma_a_rgb::NTuple{3,Float32} = z3_; sigma_s_rgb::NTuple{3,Float32} = z3_; Le_rgb::NTuple{3,Float32} = z3_
    scale::Float32 = 1f0; g::Float32 = 0f0; bounds_min::NTuple{3,Float32} = z3_; bounds_max::NTuple{3,Float32} = z3_
    render_from_medium::NTuple{16,Float32} = I16; medium_from_render::NTuple{16,Float32} = I16
    density_res::NTuple{3,Int32} = (0, 0, 0); density::Ptr{Float32} = C_NULL
    majorant_res::NTuple{3,Int32} = (0, 0, 0); majorant::Ptr{Float32} = C_NULL      # NULL: built on the device (k_build_majorant)
    nanovdb_buf::Ptr{UInt8} = C_NULL; nanovdb_bytes::UInt64 = 0
    nanovdb_inv_mat::NTuple{9,Float32} = ntuple(_ -> 0f0, 9); nanovdb_vec::NTuple{3,Float32} = z3_
    nanovdb_root_offset::UInt64 = 0; nanovdb_upper_offset::UInt64 = 0; nanovdb_lower_offset::UInt64 = 0; nanovdb_leaf_offset::UInt64 = 0
    nanovdb_root_tiles::Int32 = 0; nanovdb_upper_count::Int32 = 0; nanovdb_lower_count::Int32 = 0; nanovdb_leaf_count::Int32 = 0
    rgb_sigma_a::Ptr{Float32} = C_NULL; rgb_sigma_s::Ptr{Float32} = C_NULL; rgb_Le::Ptr{Float32} = C_NULL; Le_scale::Float32 = 0f0
    nanovdb_index_min::NTuple{3,Int32} = (0, 0, 0); nanovdb_index_max::NTuple{3,Int32} = (0, 0, 0)
end
struct HkCamera
    raster_to_camera::NTuple{16,Float32}; camera_to_world::NTuple{16,Float32}; lens_radius::Float32; focal_distance::Float32
    shutter_open::Float32; shutter_close::Float32; dx_camera::NTuple{3,Float32}; dy_camera::NTuple{3,Float32}
end
struct HkFilter
    type::Int32; radius::NTuple{2,Float32}; nx::Int32; ny::Int32; func::Ptr{Float32}; marginal_cdf::Ptr{Float32}
    marginal_func::Ptr{Float32}; conditional_cdf::Ptr{Float32}; domain_min::NTuple{2,Float32}; domain_max::NTuple{2,Float32}; func_integral::Float32
end
struct HkRenderParams
    width::Int32; height::Int32; max_depth::Int32; samples_per_pixel::Int32; regularize::Int32
    max_component_value::Float32; sampler_seed::UInt32; sobol_log2_spp::Int32; sobol_n_base4_digits::Int32
    material_coherence::Int32; sample_batch::Int32
end
struct HkPostprocess
    exposure::Float32; tonemap_mode::Int32; inv_gamma::Float32; apply_gamma::Int32; white_point::Float32; imaging_ratio::Float32
    apply_wb::Int32; wb::NTuple{9,Float32}; mask_escaped::Int32; background::NTuple{3,Float32}
end
struct HkDenoiseConfig; iterations::Int32; sigma_color::Float32; sigma_normal::Float32; sigma_depth::Float32; use_variance::Int32; end

const HK_MAT = (matte=1, mirror=2, glass=3, conductor=4, coated_diffuse=5, thin_dielectric=6, diffuse_transmission=7, mix=8,
                coated_conductor=9, coated_diffuse_transmission=10)
const MATFLAG_REMAP, MATFLAG_SPECTRAL_ETA_K, MATFLAG_USE_ETA_K, MATFLAG_VERTEX_COLORS = UInt32(1), UInt32(2), UInt32(4), UInt32(8)

# ---- context -----------------------------------------------------------------------------------------------------------
mutable struct Ctx
    ptr::Ptr{Cvoid}
    scene_id::UInt64            # objectid of the synced scene the device copy was built from
    size::Tuple{Int,Int}
    params_key::Any
    keep::Vector{Any}           # host arrays borrowed by the last upload (the library copies during the call; kept for clarity)
end
const CTX = IdDict{VolPath,Ctx}()

errmsg(c::Ptr{Cvoid}) = unsafe_string(ccall((:hk_last_error, lib), Cstring, (Ptr{Cvoid},), c))
check(rc::Int32, c::Ptr{Cvoid}) = rc == 0 || error("hikari_cuda ($rc): " * errmsg(c))

function context(vp::VolPath)
    get!(CTX, vp) do
        p = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:hk_create, lib), Int32, (Int32, Ref{Ptr{Cvoid}}), CUDA.deviceid(CUDA.device()), p)
        rc == 0 || error("hikari_cuda: no usable CUDA device (status $rc); there is no CPU fallback")
        c = Ctx(p[], 0, (0, 0), nothing, Any[])
        # all library work is ordered on the task-local CUDA.jl stream: no device-wide synchronisation between Hikari's own
        # CuArray operations and the renderer (hk_set_stream, include/hikari_cuda.h)
        check(ccall((:hk_set_stream, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), c.ptr, Base.unsafe_convert(Ptr{Cvoid}, CUDA.stream().handle)), c.ptr)
        c
    end
end

t3(v) = (Float32(v[1]), Float32(v[2]), Float32(v[3]))
const z3 = (0f0, 0f0, 0f0)
rgb3(s::Hikari.RGBSpectrum) = (s.c[1], s.c[2], s.c[3])                                  # spectrum.jl:61-70 (c[4] = alpha)
rowmajor16(m) = ntuple(i -> Float32(m[(i - 1) ÷ 4 + 1, (i - 1) % 4 + 1]), 16)            # Mat4f is column-major; the ABI wants rows
rowmajor12(m) = ntuple(i -> Float32(m[(i - 1) ÷ 4 + 1, (i - 1) % 4 + 1]), 12)

# ---- geometry: scene.accel (Raycore.TLAS) -> HkGeometry ------------------------------------------------------------------
# Contract needed from Raycore (un-vendored; these are the only TLAS internals the shim reads — scene.jl:196-199 names
# `accel.instances` and `accel.blas_array`): per BLAS the object-space vertex / normal / uv arrays, the face index triples and the
# per-face TriangleMeta (scene.jl:11-15, attached by push! as `face_meta`, scene-mesh.jl:12-17); per instance its BLAS index and
# the 4x4 object-to-world transform.  Instances are passed through (HkGeometry.instances): the library builds one bottom-level
# BVH per BLAS and a top level over the instances, the global primitive id is instance-major (= the reference's TLAS order).
function tlas_arrays(accel)
    pos = Float32[]; nrm = Float32[]; uvs = Float32[]; idx = UInt32[]; meshes = HkMesh[]; metas = Vector{Vector{Hikari.TriangleMeta}}()
    has_n = true; has_uv = true
    for blas in accel.blas_array
        voff = UInt32(length(pos) ÷ 3)
        for p in Raycore.vertices(blas); append!(pos, t3(p)); end
        ns = Raycore.normals(blas); has_n &= ns !== nothing
        ns === nothing ? append!(nrm, fill(NaN32, 3 * length(Raycore.vertices(blas)))) : foreach(n -> append!(nrm, t3(n)), ns)
        us = Raycore.uvs(blas); has_uv &= us !== nothing
        us === nothing ? append!(uvs, zeros(Float32, 2 * length(Raycore.vertices(blas)))) : foreach(u -> append!(uvs, (Float32(u[1]), Float32(u[2]))), us)
        first_tri = UInt32(length(idx) ÷ 3)
        for f in Raycore.faces(blas); append!(idx, (voff + UInt32(f[1]) - 1, voff + UInt32(f[2]) - 1, voff + UInt32(f[3]) - 1)); end   # 0-based vertices
        push!(meshes, HkMesh(first_tri, UInt32(length(Raycore.faces(blas)))))
        push!(metas, collect(Raycore.face_meta(blas)))
    end
    inst = HkInstance[]
    for I in accel.instances
        o2w = Raycore.transform(I); w2o = inv(o2w)
        # one medium interface per instance: push!(scene, mesh, material) gives every face of a mesh the same interface
        # (build_face_meta, scene-mesh.jl:56-72); per-face materials / emissive meshes take the flattened path below
        push!(inst, HkInstance(UInt32(Raycore.blas_index(I) - 1), metas[Raycore.blas_index(I)][1].medium_interface_idx, rowmajor12(o2w), rowmajor12(w2o)))
    end
    (; pos, nrm = has_n ? nrm : Float32[], uvs = has_uv ? uvs : Float32[], idx, meshes, metas, inst)
end

uniform_interfaces(metas) = all(m -> all(x -> x.medium_interface_idx == m[1].medium_interface_idx && x.arealight_flat_idx == 0, m), metas)

function upload_geometry!(c::Ctx, accel)
    A = tlas_arrays(accel)
    if uniform_interfaces(A.metas) && !isempty(A.inst)
        GC.@preserve A begin
            g = HkGeometry(pointer(A.pos), isempty(A.nrm) ? C_NULL : pointer(A.nrm), C_NULL, isempty(A.uvs) ? C_NULL : pointer(A.uvs), pointer(A.idx), C_NULL,
                           length(A.pos) ÷ 3, length(A.idx) ÷ 3, pointer(A.meshes), length(A.meshes), pointer(A.inst), length(A.inst))
            check(ccall((:hk_upload_geometry, lib), Int32, (Ptr{Cvoid}, Ref{HkGeometry}), c.ptr, g), c.ptr)
        end
        return
    end
    # per-face materials or emissive faces: world-space triangle soup with one TriangleMeta (3 x u32) per triangle
    pos = Float32[]; nrm = Float32[]; uvs = Float32[]; idx = UInt32[]; meta = UInt32[]
    for I in A.inst
        m = A.meshes[I.mesh + 1]; O = reshape(collect(I.object_to_world), 4, 3)'; W = reshape(collect(I.world_to_object), 4, 3)'
        voff = UInt32(length(pos) ÷ 3)
        vs = unique(sort(A.idx[3 * m.first_tri + 1:3 * (m.first_tri + m.n_tris)])); remap = Dict(v => voff + UInt32(k - 1) for (k, v) in enumerate(vs))
        for v in vs
            p = A.pos[3v + 1:3v + 3]; append!(pos, O[:, 1:3] * p + O[:, 4])
            if !isempty(A.nrm); n = W[:, 1:3]' * A.nrm[3v + 1:3v + 3]; append!(nrm, n / sqrt(sum(abs2, n))); end     # normals by the inverse transpose
            isempty(A.uvs) || append!(uvs, A.uvs[2v + 1:2v + 2])
        end
        for t in 0:m.n_tris - 1
            tri = m.first_tri + t
            append!(idx, (remap[A.idx[3tri + 1]], remap[A.idx[3tri + 2]], remap[A.idx[3tri + 3]]))
            tm = A.metas[I.mesh + 1][t + 1]; append!(meta, (tm.medium_interface_idx, tm.primitive_index, tm.arealight_flat_idx))
        end
    end
    GC.@preserve pos nrm uvs idx meta begin
        g = HkGeometry(pointer(pos), isempty(nrm) ? C_NULL : pointer(nrm), C_NULL, isempty(uvs) ? C_NULL : pointer(uvs), pointer(idx), pointer(meta),
                       length(pos) ÷ 3, length(idx) ÷ 3, C_NULL, 0, C_NULL, 0)
        check(ccall((:hk_upload_geometry, lib), Int32, (Ptr{Cvoid}, Ref{HkGeometry}), c.ptr, g), c.ptr)
    end
end

# ---- materials: scene.materials (MultiTypeSet) + scene.media_interfaces -> HkMaterial[] / HkMediumInterface[] ---------------
# Flat 1-based ids: type groups in MultiTypeSet order, elements within a group in push order — the numbering resolve_mix_material
# and MediumInterfaceIdx use through SetKey(type_idx, vec_idx) (multi-material-eval.jl:128,181; mix-material.jl:214-268).
struct Flat; offsets::Vector{Int}; end                     # offsets[type_idx] = number of elements in earlier groups
Flat(set) = Flat(cumsum([0; [length(group) for group in set][1:end-1]]))
flat_id(F::Flat, k) = k.type_idx == 0 ? UInt32(0) : UInt32(F.offsets[k.type_idx] + k.vec_idx)      # SetKey() = (0, 0) = none

struct TexPool; list::Vector{HkTexture}; keep::Vector{Any}; ids::IdDict{Any,Int32}; end
TexPool() = TexPool(HkTexture[], Any[], IdDict{Any,Int32}())
# a Texture over a Matrix{RGBSpectrum} (textures/basic.jl:5-10): r, g, b packed to 3 floats and, when some alpha != 1, the alpha
# plane; a Texture over a Matrix{Float32}: r = g = b (HkMaterial.ftex reads the first channel); both keep the Matrix' (h, w)
# column-major order.  VertexColorTexture (basic.jl:43-46): its (3, n_faces) matrix as h = 3.
texel(s::Hikari.RGBSpectrum) = (s.c[1], s.c[2], s.c[3], s.c[4]); texel(v::Real) = (Float32(v), Float32(v), Float32(v), 1f0)
function texture_id!(P::TexPool, t)
    get!(P.ids, t) do
        data = t.data; h, w = size(data)
        rgb = Vector{Float32}(undef, 3 * h * w); alpha = Vector{Float32}(undef, h * w)
        for (k, s) in enumerate(data); r, g, b, a = texel(s); rgb[3k - 2], rgb[3k - 1], rgb[3k] = r, g, b; alpha[k] = a; end
        has_alpha = any(!=(1f0), alpha)
        push!(P.keep, rgb); has_alpha && push!(P.keep, alpha)
        push!(P.list, HkTexture(pointer(rgb), h, w, has_alpha ? pointer(alpha) : C_NULL))
        Int32(length(P.list))
    end
end
isconst(t) = !(t isa Hikari.Texture) || t.isconst
constval(t) = t isa Hikari.Texture ? t.constval : t                                        # raw value, or ConstTexture (basic.jl:11-14)

struct SpecPool; lambdas::Vector{Float32}; values::Vector{Float32}; offsets::Vector{UInt32}; end
SpecPool() = SpecPool(Float32[], Float32[], UInt32[0])
function spectrum_id!(S::SpecPool, s::Hikari.PiecewiseLinearSpectrum)                      # piecewise-linear.jl:4-7
    append!(S.lambdas, s.lambdas); append!(S.values, s.values); push!(S.offsets, length(S.lambdas)); Int32(length(S.offsets) - 1)
end

# HkMaterial under construction.  Every RGB / scalar parameter of the reference is `Texture, TextureRef or a raw value`: a constant goes
# into rgb<slot> / f[k], a texture into tex[slot] / ftex[k] (eval_tex at the hit on the device).  Slots per type: include/hikari_cuda.h.
mutable struct MatB
    type::Int32; flags::UInt32; rgb::Vector{NTuple{3,Float32}}; f::Vector{Float32}; spec::Vector{Int32}; ival::Vector{Int32}; tex::Vector{Int32}; ftex::Vector{Int32}
end
MatB(type; flags=UInt32(0)) = MatB(type, flags, fill((0f0, 0f0, 0f0), 3), zeros(Float32, 8), zeros(Int32, 2), zeros(Int32, 2), zeros(Int32, 4), zeros(Int32, 8))
rgb!(b::MatB, P, slot, v) = (isconst(v) ? (b.rgb[slot] = rgb3(constval(v))) : (b.tex[slot] = texture_id!(P, v)); b)
f!(b::MatB, P, k, v) = (isconst(v) ? (b.f[k] = Float32(constval(v))) : (b.ftex[k] = texture_id!(P, v)); b)
HkMaterial(b::MatB) = HkMaterial(b.type, b.flags, b.rgb[1], b.rgb[2], (b.rgb[3]..., 0f0), Tuple(b.f), Tuple(b.spec), Tuple(b.ival), Tuple(b.tex), Tuple(b.ftex))
remap(m) = m.remap_roughness ? MATFLAG_REMAP : UInt32(0)

function hk(m::Hikari.MatteMaterial, F, P, S)                                                 # uber-material.jl:180-183; f0 = sigma
    b = MatB(HK_MAT.matte); rgb!(b, P, 1, m.Kd); f!(b, P, 1, m.σ)
    m.Kd isa Hikari.VertexColorTexture && (b.flags |= MATFLAG_VERTEX_COLORS)
    HkMaterial(b)
end
hk(m::Hikari.MirrorMaterial, F, P, S) = HkMaterial(rgb!(MatB(HK_MAT.mirror), P, 1, m.Kr))                                       # :193-194
function hk(m::Hikari.GlassMaterial, F, P, S)                                                 # :209-215; f0 = index (the roughness fields are ignored by the spectral path)
    b = MatB(HK_MAT.glass; flags=remap(m)); rgb!(b, P, 1, m.Kr); rgb!(b, P, 2, m.Kt); f!(b, P, 1, m.index); HkMaterial(b)
end
function hk(m::Hikari.ConductorMaterial, F, P, S)                                             # :378-383; f0 = roughness
    b = MatB(HK_MAT.conductor; flags=remap(m)); f!(b, P, 1, m.roughness)
    if m.eta isa Hikari.PiecewiseLinearSpectrum
        b.flags |= MATFLAG_SPECTRAL_ETA_K; b.spec[1] = spectrum_id!(S, m.eta); b.spec[2] = spectrum_id!(S, m.k)
    else
        rgb!(b, P, 1, m.eta); rgb!(b, P, 2, m.k)
    end
    HkMaterial(b)
end
function coat!(b, P, m)                                                                       # f0 / f1 = u / v roughness, f2 = thickness, f3 = eta, f4 = g
    f!(b, P, 1, m.u_roughness); f!(b, P, 2, m.v_roughness); f!(b, P, 3, m.thickness); b.f[4] = m.eta; f!(b, P, 5, m.g)
    b.ival[1] = m.max_depth; b.ival[2] = m.n_samples; b
end
function hk(m::Hikari.CoatedDiffuseMaterial, F, P, S)                                         # coated-diffuse.jl:32-42
    b = MatB(HK_MAT.coated_diffuse; flags=remap(m)); rgb!(b, P, 1, m.reflectance); rgb!(b, P, 2, m.albedo); HkMaterial(coat!(b, P, m))
end
function hk(m::Hikari.CoatedDiffuseTransmissionMaterial, F, P, S)                             # coated-diffuse-transmission.jl:12-23
    b = MatB(HK_MAT.coated_diffuse_transmission; flags=remap(m)); rgb!(b, P, 1, m.reflectance); rgb!(b, P, 2, m.albedo); rgb!(b, P, 3, m.transmittance); HkMaterial(coat!(b, P, m))
end
hk(m::Hikari.ThinDielectricMaterial, F, P, S) = (b = MatB(HK_MAT.thin_dielectric); b.f[1] = m.eta; HkMaterial(b))               # thin-dielectric.jl:45-46; f0 = eta
function hk(m::Hikari.DiffuseTransmissionMaterial, F, P, S)                                   # diffuse-transmission.jl:39-42; f0 = scale
    b = MatB(HK_MAT.diffuse_transmission); rgb!(b, P, 1, m.reflectance); rgb!(b, P, 2, m.transmittance); b.f[1] = m.scale; HkMaterial(b)
end
function hk(m::Hikari.CoatedConductorMaterial, F, P, S)                                       # coated-conductor.jl:48-76
    b = MatB(HK_MAT.coated_conductor; flags=remap(m) | (m.use_eta_k ? MATFLAG_USE_ETA_K : UInt32(0)))
    f!(b, P, 1, m.interface_u_roughness); f!(b, P, 2, m.interface_v_roughness); f!(b, P, 3, m.thickness); b.f[4] = m.interface_eta; f!(b, P, 5, m.g)
    f!(b, P, 6, m.conductor_u_roughness); f!(b, P, 7, m.conductor_v_roughness)
    rgb!(b, P, 3, m.albedo); b.ival[1] = m.max_depth; b.ival[2] = m.n_samples
    if m.use_eta_k
        e, k = m.conductor_eta, m.conductor_k
        if constval(e) isa Hikari.PiecewiseLinearSpectrum
            b.flags |= MATFLAG_SPECTRAL_ETA_K; b.spec[1] = spectrum_id!(S, constval(e)); b.spec[2] = spectrum_id!(S, constval(k))
        else
            rgb!(b, P, 1, e); rgb!(b, P, 2, k)
        end
    else
        rgb!(b, P, 1, m.reflectance)
    end
    HkMaterial(b)
end
# MixMaterial (mix-material.jl:39-99): the two sub-materials' SetKeys are stored in the material (material_indices); the mix hash
# consumes their type_idx / vec_idx (mix_hash_float :114-158), the library needs their flat ids as well.  f0 / ftex[0] = amount (a constant, or a texture evaluated at the hit's uv, :183).
function hk(m::Hikari.MixMaterial, F, P, S)
    k1, k2 = m.material_indices
    b = MatB(HK_MAT.mix; flags=UInt32(k1.type_idx) | (UInt32(k2.type_idx) << 8)); f!(b, P, 1, m.amount)
    b.spec[1] = k1.vec_idx; b.spec[2] = k2.vec_idx; b.ival[1] = flat_id(F, k1); b.ival[2] = flat_id(F, k2)
    HkMaterial(b)
end
hk(m::Hikari.MediumInterface, F, P, S) = hk(m.material, F, P, S)                              # medium-interface.jl:39-42
hk(m::Material, F, P, S) = error("hikari_cuda: material type $(typeof(m)) is not on the VolPath spectral path")

function upload_materials!(c::Ctx, scene)
    F = Flat(scene.materials); P = TexPool(); S = SpecPool()
    mats = HkMaterial[hk(m, F, P, S) for group in scene.materials for m in group]
    FM = Flat(scene.media)
    ifaces = HkMediumInterface[HkMediumInterface(flat_id(F, mi.material), flat_id(FM, mi.inside), flat_id(FM, mi.outside)) for mi in scene.media_interfaces]
    GC.@preserve P S begin
        check(ccall((:hk_upload_textures, lib), Int32, (Ptr{Cvoid}, Ptr{HkTexture}, UInt32), c.ptr, P.list, length(P.list)), c.ptr)
        sp = HkSpectra(pointer(S.lambdas), pointer(S.values), pointer(S.offsets), length(S.offsets) - 1)
        check(ccall((:hk_upload_spectra, lib), Int32, (Ptr{Cvoid}, Ref{HkSpectra}), c.ptr, sp), c.ptr)
        check(ccall((:hk_upload_materials, lib), Int32, (Ptr{Cvoid}, Ptr{HkMaterial}, UInt32, Ptr{HkMediumInterface}, UInt32), c.ptr, mats, length(mats), ifaces, length(ifaces)), c.ptr)
    end
    F
end

# ---- lights: FLAT index order = flat_to_light_index (light-sampler.jl:289-329) ----------------------------------------------
spectrum_fields(i::Hikari.RGBSpectrum) = (Int32(0), rgb3(i), z3, 0f0)
spectrum_fields(i::Hikari.RGBIlluminantSpectrum) = (Int32(1), z3, (i.poly.c0, i.poly.c1, i.poly.c2), i.scale)              # rgb2spec.jl:331-334
const Z16 = ntuple(_ -> 0f0, 16); const Z9 = ntuple(_ -> 0f0, 9); const Z6 = ntuple(_ -> 0f0, 6)
function hk(l::Hikari.PointLight, envs)                                                                                    # point.jl:1-13
    k, rgb, poly, s = spectrum_fields(l.i); HkLight(1, k, l.scale, rgb, poly, s, t3(l.position), z3, 0, 0, Z16, Z9, z3, 0, Z6, 0, 0)
end
function hk(l::Hikari.SpotLight, envs)                                                                                     # spot.jl:1-12
    k, rgb, poly, s = spectrum_fields(l.i)
    HkLight(2, k, l.scale, rgb, poly, s, t3(l.position), z3, l.cos_total_width, l.cos_falloff_start, rowmajor16(l.world_to_light.m), Z9, z3, 0, Z6, 0, 0)
end
function hk(l::Union{Hikari.DirectionalLight,Hikari.SunLight}, envs)                                                       # directional.jl:6-17, sun.jl:7-17
    k, rgb, poly, s = spectrum_fields(l.i); HkLight(l isa Hikari.SunLight ? 4 : 3, k, l.scale, rgb, poly, s, z3, t3(l.direction), 0, 0, Z16, Z9, z3, 0, Z6, 0, 0)
end
function hk(l::Hikari.AmbientLight, envs)                                                                                  # ambient.jl:1-3
    k, rgb, poly, s = spectrum_fields(l.i); HkLight(6, k, l.scale, rgb, poly, s, z3, z3, 0, 0, Z16, Z9, z3, 0, Z6, 0, 0)
end
function hk(l::Hikari.EnvironmentLight, envs)                                                                              # environment.jl:5-10
    push!(envs, l); HkLight(5, 0, 1f0, z3, z3, 0f0, z3, z3, 0, 0, Z16, Z9, z3, 0, Z6, 0, Int32(length(envs)))
end
function hk(l::Hikari.DiffuseAreaLight, envs)                                                                              # diffuse-area.jl:25-32
    v = ntuple(i -> Float32(l.vertices[(i - 1) ÷ 3 + 1][(i - 1) % 3 + 1]), 9); uv = ntuple(i -> Float32(l.uv[(i - 1) ÷ 2 + 1][(i - 1) % 2 + 1]), 6)
    HkLight(7, 0, l.scale, (isconst(l.Le) ? rgb3(constval(l.Le)) : error("hikari_cuda: a textured DiffuseAreaLight.Le is not supported")), z3, 0f0, z3, z3, 0, 0, Z16, v, t3(l.normal), l.area, uv, l.two_sided, 0)
end

function upload_lights!(c::Ctx, scene)
    envs = Any[]
    lights = HkLight[hk(l, envs) for group in scene.lights for l in group]
    keep = Any[]
    maps = map(envs) do l                                                                  # environment_map.jl:9-45, sampling.jl:179-193
        em, d = l.env_map, l.env_map.distribution
        h, w = size(em.data)
        rgb = Float32[s.c[k] for s in permutedims(em.data) for k in 1:3]                   # [h][w][3] row-major
        cf, cc = collect(d.conditional_func), collect(d.conditional_cdf)                   # (nu, nv) column-major == the ABI's [nv][nu]
        cfi, mf, mc = collect(d.conditional_func_int), collect(d.marginal_func), collect(d.marginal_cdf)
        append!(keep, (rgb, cf, cc, cfi, mf, mc))
        HkEnvMap(pointer(rgb), w, h, ntuple(i -> em.rotation[i], 9), rgb3(l.scale), pointer(cf), pointer(cc), pointer(cfi), pointer(mf), pointer(mc),
                 d.marginal_func_int, size(cf, 1), size(cf, 2))
    end
    smp = Hikari.BVHLightSampler(scene.lights; scene_radius=Hikari.world_radius(scene))    # built on the CPU as today (bvh-light-sampler.jl:283-466)
    nodes = HkLightBVHNode[HkLightBVHNode(t3(n.bounds.bounds.p_min), t3(n.bounds.bounds.p_max), t3(n.bounds.w), n.bounds.phi, n.bounds.cos_theta_o, n.bounds.cos_theta_e,
                                          n.bounds.two_sided, n.child_or_light_index, n.is_leaf, 0) for n in smp.nodes]
    trails = collect(UInt32, smp.light_to_bit_trail); inf = collect(Int32, smp.infinite_light_indices)
    GC.@preserve keep maps nodes trails inf begin
        isempty(maps) || check(ccall((:hk_upload_envmaps, lib), Int32, (Ptr{Cvoid}, Ptr{HkEnvMap}, UInt32), c.ptr, maps, length(maps)), c.ptr)
        s = HkLightSampler(pointer(nodes), length(nodes), pointer(trails), pointer(inf), length(inf), length(lights) - length(inf))
        check(ccall((:hk_upload_lights, lib), Int32, (Ptr{Cvoid}, Ptr{HkLight}, UInt32, Ref{HkLightSampler}), c.ptr, lights, length(lights), s), c.ptr)
    end
end

# ---- media ----------------------------------------------------------------------------------------------------------------
const NULLF = Ptr{Float32}(C_NULL)
bmin(b) = t3(b.p_min); bmax(b) = t3(b.p_max)
# Majorant grids are left to the library (majorant = NULL: built on the device from the voxels uploaded here, bit for bit the grid
# build_majorant_grid / build_rgb_majorant_grid / build_nanovdb_majorant_grid produce); only their resolution is passed.
hk(m::Hikari.HomogeneousMedium, keep) = HkMedium(type=1, sigma_a_rgb=rgb3(m.σ_a), sigma_s_rgb=rgb3(m.σ_s), Le_rgb=rgb3(m.Le), g=m.g)       # media.jl:762-766
function hk(m::Hikari.GridMedium, keep)                                                    # media.jl:873-895 (density[x, y, z] is already x-fastest)
    d = collect(Float32, m.density); push!(keep, d)
    HkMedium(type=2, sigma_a_rgb=rgb3(m.σ_a), sigma_s_rgb=rgb3(m.σ_s), g=m.g, bounds_min=bmin(m.bounds), bounds_max=bmax(m.bounds),
             render_from_medium=rowmajor16(m.medium_to_render), medium_from_render=rowmajor16(m.render_to_medium),
             density_res=Tuple(Int32.(m.density_res)), density=pointer(d), majorant_res=Tuple(Int32.(m.majorant_grid.res)))
end
function hk(m::Hikari.NanoVDBMedium, keep)                                                 # nanovdb.jl:153-182; byte offsets become 0-based
    buf = collect(UInt8, m.buffer); push!(keep, buf)
    HkMedium(type=3, sigma_a_rgb=rgb3(m.σ_a), sigma_s_rgb=rgb3(m.σ_s), g=m.g, bounds_min=bmin(m.bounds), bounds_max=bmax(m.bounds),
             majorant_res=Tuple(Int32.(m.majorant_grid.res)), nanovdb_buf=pointer(buf), nanovdb_bytes=length(buf), nanovdb_inv_mat=m.inv_mat, nanovdb_vec=m.vec,
             nanovdb_root_offset=m.root_offset - 1, nanovdb_upper_offset=m.upper_offset - 1, nanovdb_lower_offset=m.lower_offset - 1, nanovdb_leaf_offset=m.leaf_offset - 1,
             nanovdb_root_tiles=m.root_table_size, nanovdb_upper_count=m.upper_count, nanovdb_lower_count=m.lower_count, nanovdb_leaf_count=m.leaf_count,
             nanovdb_index_min=m.index_bbox_min, nanovdb_index_max=m.index_bbox_max)
end
function hk(m::Hikari.RGBGridMedium, keep)                                                 # media.jl:1002-1075
    pack(g) = g === nothing ? Float32[] : Float32[s.c[k] for s in g for k in 1:3]
    a, s, e = pack(m.σ_a_grid), pack(m.σ_s_grid), pack(m.Le_grid); append!(keep, (a, s, e))
    res = size(something(m.σ_a_grid, m.σ_s_grid))
    HkMedium(type=4, g=m.g, scale=m.sigma_scale, Le_scale=m.Le_scale, bounds_min=bmin(m.bounds), bounds_max=bmax(m.bounds),
             render_from_medium=rowmajor16(m.medium_to_render), medium_from_render=rowmajor16(m.render_to_medium),
             density_res=Tuple(Int32.(res)), majorant_res=Tuple(Int32.(m.majorant_grid.res)),
             rgb_sigma_a=isempty(a) ? NULLF : pointer(a), rgb_sigma_s=isempty(s) ? NULLF : pointer(s), rgb_Le=isempty(e) ? NULLF : pointer(e))
end
function upload_media!(c::Ctx, scene)
    keep = Any[]
    media = HkMedium[hk(m, keep) for group in scene.media for m in group]
    GC.@preserve keep check(ccall((:hk_upload_media, lib), Int32, (Ptr{Cvoid}, Ptr{HkMedium}, UInt32), c.ptr, media, length(media)), c.ptr)
end

# ---- camera / filter / params -------------------------------------------------------------------------------------------
HkCamera(cam::PerspectiveCamera) = HkCamera(rowmajor16(cam.core.raster_to_camera.m), rowmajor16(cam.core.core.camera_to_world.m), cam.core.lens_radius, cam.core.focal_distance,
                                            cam.core.core.shutter_open, cam.core.core.shutter_close, t3(cam.dx_camera), t3(cam.dy_camera))        # perspective.jl:1-10, 41-51
function upload_filter!(c::Ctx, vp)                                                       # filter.jl:574-633 (GPUFilterParams + GPUFilterSamplerData)
    p, d = vp.filter_params, vp.filter_sampler_data
    if p.filter_type <= 2
        f = HkFilter(p.filter_type, (p.radius[1], p.radius[2]), 0, 0, NULLF, NULLF, NULLF, NULLF, (0f0, 0f0), (0f0, 0f0), 0f0)
        return check(ccall((:hk_set_filter, lib), Int32, (Ptr{Cvoid}, Ref{HkFilter}), c.ptr, f), c.ptr)
    end
    func = collect(permutedims(d.func)); ccdf = collect(permutedims(d.conditional_cdf)); mcdf = collect(d.marginal_cdf); mfunc = collect(d.marginal_func)   # (ny, nx) -> [ny][nx]
    GC.@preserve func ccdf mcdf mfunc begin
        f = HkFilter(p.filter_type, (p.radius[1], p.radius[2]), d.nx, d.ny, pointer(func), pointer(mcdf), pointer(mfunc), pointer(ccdf),
                     (d.domain_min[1], d.domain_min[2]), (d.domain_max[1], d.domain_max[2]), d.func_integral)
        check(ccall((:hk_set_filter, lib), Int32, (Ptr{Cvoid}, Ref{HkFilter}), c.ptr, f), c.ptr)
    end
end

function upload_tables!(c::Ctx)                                                           # volpath-state.jl:140-160, sobol.jl:370-379
    tab = Hikari.get_srgb_table(); d65 = collect(Float32, Hikari.D65_ILLUMINANT_VALUES)
    sob = collect(UInt32, Hikari.SobolMatrices32); cx, cy, cz = collect(Float32, Hikari.CIE_X), collect(Float32, Hikari.CIE_Y), collect(Float32, Hikari.CIE_Z)
    GC.@preserve tab d65 sob cx cy cz begin
        t = HkTables(pointer(sob), pointer(cx), pointer(cy), pointer(cz), pointer(d65), tab.res, pointer(tab.scale), pointer(tab.coeffs))
        check(ccall((:hk_upload_tables, lib), Int32, (Ptr{Cvoid}, Ref{HkTables}), c.ptr, t), c.ptr)
    end
end

function prepare!(c::Ctx, vp::VolPath, scene, film, camera)
    h, w = size(film.framebuffer)
    if c.scene_id != objectid(scene.accel) || c.size != (h, w)                             # the reference re-adapts the scene every sample (volpath.jl:455-461)
        upload_tables!(c); upload_geometry!(c, scene.accel); upload_materials!(c, scene); upload_media!(c, scene); upload_lights!(c, scene)
        c.scene_id = objectid(scene.accel); c.size = (h, w); c.params_key = nothing
    end
    key = (w, h, vp.max_depth, vp.samples_per_pixel, vp.regularize, vp.max_component_value)
    if c.params_key != key                                                                 # VolPath fields are read on every render! (volpath.jl:445-520)
        upload_filter!(c, vp)
        l2, nb4 = Hikari.compute_zsobol_params(max(Int(vp.samples_per_pixel), 4096), w, h)     # volpath.jl:475, sobol.jl:317-323
        p = HkRenderParams(w, h, vp.max_depth, vp.samples_per_pixel, vp.regularize, vp.max_component_value, 0, l2, nb4, 0, 0)
        check(ccall((:hk_set_params, lib), Int32, (Ptr{Cvoid}, Ref{HkRenderParams}), c.ptr, p), c.ptr)
        c.params_key = key
    end
    check(ccall((:hk_set_camera, lib), Int32, (Ptr{Cvoid}, Ref{HkCamera}), c.ptr, HkCamera(camera)), c.ptr)
end

# film.framebuffer is a CuMatrix{RGB{Float32}} of size (H, W): exactly the layout hk_read_film_dev writes (volpath.jl:384-417)
fb_ptr(a::CuArray) = reinterpret(Ptr{Float32}, pointer(a))

# ---- the overloaded entry points -------------------------------------------------------------------------------------------
const CuFilm = Film{<:Any,<:Any,<:CuArray}                                                  # same test as KA.get_backend(film.framebuffer) (volpath.jl:453)

function render!(vp::VolPath, scene::AbstractScene, film::CuFilm, camera::Camera)          # volpath.jl:445-636: ONE sample per call
    c = context(vp); prepare!(c, vp, scene, film, camera)
    idx = film.iteration_index[] + Int32(1); film.iteration_index[] = idx
    check(ccall((:hk_render_samples, lib), Int32, (Ptr{Cvoid}, Int32, Int32), c.ptr, idx, 1), c.ptr)
    check(ccall((:hk_read_film_dev, lib), Int32, (Ptr{Cvoid}, Ptr{Float32}), c.ptr, fb_ptr(film.framebuffer)), c.ptr)      # stream-ordered, no host sync
    nothing
end

function (vp::VolPath)(scene::AbstractScene, film::CuFilm, camera::Camera)                 # volpath.jl:655-670: all samples in ONE call (batched passes)
    film.iteration_index[] = 0
    c = context(vp); prepare!(c, vp, scene, film, camera)
    check(ccall((:hk_clear, lib), Int32, (Ptr{Cvoid},), c.ptr), c.ptr)
    check(ccall((:hk_render_samples, lib), Int32, (Ptr{Cvoid}, Int32, Int32), c.ptr, 1, vp.samples_per_pixel), c.ptr)
    film.iteration_index[] = vp.samples_per_pixel
    check(ccall((:hk_read_film_dev, lib), Int32, (Ptr{Cvoid}, Ptr{Float32}), c.ptr, fb_ptr(film.framebuffer)), c.ptr)
    film.framebuffer
end

clear!(vp::VolPath) = haskey(CTX, vp) ? (check(ccall((:hk_clear, lib), Int32, (Ptr{Cvoid},), CTX[vp].ptr), CTX[vp].ptr); nothing) : nothing     # volpath.jl:108-113
function Base.close(vp::VolPath)                                                            # Hikari.jl:47, volpath-state.jl:238-273
    haskey(CTX, vp) && (ccall((:hk_destroy, lib), Int32, (Ptr{Cvoid},), CTX[vp].ptr); delete!(CTX, vp)); nothing
end

# scene.jl:104-112 — interactive edits go straight to the device copy: one struct, no scene re-upload
function Hikari.update_material!(scene::Scene, idx::UInt32, new_material::Material, vp::VolPath)
    Hikari.update_material!(scene, idx, new_material)
    haskey(CTX, vp) || return
    c = CTX[vp]; F = Flat(scene.materials); P = TexPool(); S = SpecPool()
    m = hk(new_material, F, P, S)
    isempty(P.list) && isempty(S.lambdas) || error("hikari_cuda: update_material! with a new texture / spectrum needs a scene re-upload")
    check(ccall((:hk_update_material, lib), Int32, (Ptr{Cvoid}, UInt32, Ref{HkMaterial}), c.ptr, flat_id(F, scene.media_interfaces[idx].material), m), c.ptr)
end

const TONEMAP = Dict(nothing => 0, :none => 0, :reinhard => 1, :reinhard_extended => 2, :aces => 3, :uncharted2 => 4, :filmic => 5)
function Hikari.postprocess!(film::CuFilm, vp::VolPath; exposure=1f0, tonemap=:aces, gamma=2.2f0, white_point=4f0, sensor=nothing, background=nothing)   # postprocess.jl:281-357
    wb = sensor === nothing || sensor.white_balance <= 0 ? nothing : Hikari.compute_white_balance_matrix(sensor.white_balance)      # spectral/color.jl:522-546 (host)
    ratio = sensor === nothing ? 1f0 : Float32(sensor.exposure_time * sensor.iso / 100)
    bg = background === nothing ? z3 : (Float32(background.r), Float32(background.g), Float32(background.b))
    p = HkPostprocess(exposure, TONEMAP[tonemap], gamma === nothing ? 1f0 : 1f0 / Float32(gamma), gamma === nothing ? 0 : 1, white_point, ratio,
                      wb === nothing ? 0 : 1, wb === nothing ? Z9 : ntuple(i -> Float32(wb[(i - 1) ÷ 3 + 1, (i - 1) % 3 + 1]), 9), background === nothing ? 0 : 1, bg)
    c = CTX[vp]
    check(ccall((:hk_postprocess_dev, lib), Int32, (Ptr{Cvoid}, Ref{HkPostprocess}, Ptr{Float32}), c.ptr, p, fb_ptr(film.postprocess)), c.ptr)
    film.postprocess
end

function Hikari.fill_aux_buffers!(film::CuFilm, vp::VolPath; has_infinite_lights::Bool=false)          # film.jl:410-431
    c = CTX[vp]
    check(ccall((:hk_fill_aux_buffers, lib), Int32, (Ptr{Cvoid}, Int32), c.ptr, has_infinite_lights), c.ptr)
    a, n, d = Array(film.albedo), Array(film.normal), Array(film.depth)                    # (the aux read-out has a host variant only)
    check(ccall((:hk_read_aux_buffers, lib), Int32, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}), c.ptr, pointer(reinterpret(Float32, a)), pointer(reinterpret(Float32, n)), pointer(d)), c.ptr)
    copyto!(film.albedo, a); copyto!(film.normal, n); copyto!(film.depth, d)
    film
end

function Hikari.denoise!(film::CuFilm, vp::VolPath; config::Hikari.DenoiseConfig=Hikari.DenoiseConfig())   # denoise.jl:301-372
    c = CTX[vp]; cfg = HkDenoiseConfig(config.iterations, config.sigma_color, config.sigma_normal, config.sigma_depth, config.use_variance)
    pp = Array(film.postprocess); fb = Array(film.framebuffer)
    check(ccall((:hk_denoise, lib), Int32, (Ptr{Cvoid}, Ref{HkDenoiseConfig}, Ptr{Float32}, Ptr{Float32}), c.ptr, cfg, pointer(reinterpret(Float32, pp)), pointer(reinterpret(Float32, fb))), c.ptr)
    copyto!(film.postprocess, pp); config.iterations >= 2 && copyto!(film.framebuffer, fb)
    nothing
end

end # module
