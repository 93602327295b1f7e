/* A plain C client of include/hikari_cuda.h (compiled with gcc by tests/test_abi.py): proves the header is valid C, that the
 * library links from C without any C++ / CUDA type in the signatures, and exercises the calling sequence a foreign host makes:
 * create -> upload_* -> set_* -> render -> read-out -> destroy, checking every status code.
 * Without a CUDA device hk_create must return HK_ERR_NO_DEVICE (there is no CPU fallback): the program then prints "no-device"
 * and exits 0.  With a device it renders one lit triangle and checks the image is finite, lit in the middle and
 * black in a corner, that the device-pointer read-out equals the host read-out, and that errors come back as status codes.
 * The tables arrive in a binary file written by the test (argv[1]); everything else is built here.                         */
#include "hikari_cuda.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(call) do { int32_t rc__ = (call); if (rc__ != HK_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc__, hk_last_error(ctx)); return 10; } } while (0)

static float* read_floats(FILE* f, size_t n) { float* p = (float*)malloc(4 * n); if (fread(p, 4, n, f) != n) { fprintf(stderr, "short table file\n"); exit(11); } return p; }

int main(int argc, char** argv) {
    HkContext* ctx = NULL;
    int32_t rc = hk_create(0, &ctx);
    if (rc == HK_ERR_NO_DEVICE) { printf("no-device abi=%d\n", hk_abi_version()); return ctx == NULL ? 0 : 12; }
    if (rc != HK_OK || !ctx) { fprintf(stderr, "hk_create -> %d\n", rc); return 13; }
    if (argc < 2) { fprintf(stderr, "usage: abi_client tables.bin\n"); return 14; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 15; }
    int32_t res = 0;
    if (fread(&res, 4, 1, f) != 1) return 16;
    HkTables T; memset(&T, 0, sizeof(T));
    uint32_t* sobol = (uint32_t*)read_floats(f, 52 * 1024);
    T.sobol_matrices = sobol; T.cie_x = read_floats(f, 471); T.cie_y = read_floats(f, 471); T.cie_z = read_floats(f, 471); T.d65 = read_floats(f, 107);
    T.rgb2spec_res = res; T.rgb2spec_scale = read_floats(f, (size_t)res); T.rgb2spec_coeffs = read_floats(f, 9 * (size_t)res * res * res);
    fclose(f);
    /* rendering before anything is uploaded is an error code, not a crash */
    if (hk_render_samples(ctx, 1, 1) != HK_ERR_INVALID) { fprintf(stderr, "render before upload must fail\n"); return 17; }
    CHECK(hk_upload_tables(ctx, &T));
    /* examples/single_triangle_test.jl:12-90 */
    const float pos[9] = {-1, -0.5f, 0, 1, -0.5f, 0, 0, 1, 0};
    const float nrm[9] = {0, 0, 1, 0.7f, 0, 0.714f, 0, 0.7f, 0.714f};
    const uint32_t idx[3] = {0, 1, 2};
    const uint32_t meta[3] = {1, 1, 0};      /* medium_interface_idx, primitive_index, arealight_flat_idx */
    HkGeometry G; memset(&G, 0, sizeof(G));
    G.positions = pos; G.normals = nrm; G.indices = idx; G.tri_meta = meta; G.n_verts = 3; G.n_tris = 1;
    CHECK(hk_upload_geometry(ctx, &G));
    HkMaterial M; memset(&M, 0, sizeof(M));
    M.type = HK_MAT_MATTE; M.rgb0[0] = M.rgb0[1] = M.rgb0[2] = 0.8f;
    HkMediumInterface MI = {1, 0, 0};
    CHECK(hk_upload_materials(ctx, &M, 1, &MI, 1));
    CHECK(hk_upload_media(ctx, NULL, 0));
    HkLight L; memset(&L, 0, sizeof(L));
    L.type = HK_LIGHT_DIRECTIONAL; L.rgb[0] = L.rgb[1] = L.rgb[2] = 2.0f; L.scale = 1.0f / 10567.0f; L.direction[0] = 0; L.direction[1] = 0; L.direction[2] = -1.0f; /* travel direction */
    HkLightBVHNode nodes[2]; uint32_t trails[1]; int32_t inf[1]; uint32_t nn = 0, ni = 0, nb = 0;
    hk_host_build_light_sampler(&L, 1, nodes, &nn, trails, inf, &ni, &nb);
    HkLightSampler S = {nodes, nn, trails, inf, ni, nb};
    CHECK(hk_upload_lights(ctx, &L, 1, &S));
    const int W = 64, H = 48;
    HkCamera C; memset(&C, 0, sizeof(C));
    /* camera at (0, 0, 3) looking at the origin, fov 50: built the way perspective.jl:38-78 does, with plain loops */
    {
        const float fov = 50.0f * 3.14159265f / 180.0f, aspect = (float)W / (float)H;
        const float th = tanf(0.5f * fov);
        /* raster -> camera: x in [-aspect, aspect] * th at z = 1, y flipped */
        float r2c[16] = {2 * aspect * th / W, 0, 0, -aspect * th,   0, -2 * th / H, 0, th,   0, 0, 0, 1,   0, 0, 0, 1};
        float c2w[16] = {-1, 0, 0, 0,   0, 1, 0, 0,   0, 0, -1, 3,   0, 0, 0, 1};
        memcpy(C.raster_to_camera, r2c, sizeof(r2c)); memcpy(C.camera_to_world, c2w, sizeof(c2w));
        C.lens_radius = 0.0f; C.focal_distance = 1.0e6f;
    }
    CHECK(hk_set_camera(ctx, &C));
    HkFilter F; memset(&F, 0, sizeof(F)); F.type = 1; /* Box */ F.radius[0] = F.radius[1] = 0.5f;
    CHECK(hk_set_filter(ctx, &F));
    HkRenderParams P; memset(&P, 0, sizeof(P));
    P.width = W; P.height = H; P.max_depth = 3; P.samples_per_pixel = 4; P.regularize = 1; P.max_component_value = 10.0f;
    P.sampler_seed = 0; P.sobol_log2_spp = 12; P.sobol_n_base4_digits = 6 + 6; P.sample_batch = 0;
    CHECK(hk_set_params(ctx, &P));
    CHECK(hk_clear(ctx));
    CHECK(hk_render_samples(ctx, 1, 4));
    float* img = (float*)malloc(12 * (size_t)W * H);
    CHECK(hk_read_film(ctx, img));
    double sum = 0; int finite = 1;
    for (int i = 0; i < 3 * W * H; i++) { if (!isfinite(img[i])) finite = 0; sum += img[i]; }
    const float centre = img[3 * ((W / 2) * H + H / 2)], corner = img[0];
    /* device-pointer read-out (film.framebuffer as a device array) must give the same bytes */
    void* dev = NULL; float* img2 = (float*)malloc(12 * (size_t)W * H);
    CHECK(hk_dev_alloc(ctx, 12 * (uint64_t)W * H, &dev));
    CHECK(hk_read_film_dev(ctx, (float*)dev));
    CHECK(hk_dev_download(ctx, img2, dev, 12 * (uint64_t)W * H));
    const int same = memcmp(img, img2, 12 * (size_t)W * H) == 0;
    if (hk_read_film_dev(ctx, img) != HK_ERR_INVALID) { fprintf(stderr, "a host pointer must be refused by hk_read_film_dev\n"); return 18; }
    CHECK(hk_dev_free(ctx, dev));
    HkStats st; CHECK(hk_stats(ctx, &st));
    printf("rendered abi=%d finite=%d mean=%.6f centre=%.6f corner=%.6f dev_same=%d rays=%llu launches=%llu\n", hk_abi_version(), finite, sum / (3.0 * W * H), centre, corner,
           same, (unsigned long long)st.rays_traced, (unsigned long long)st.kernel_launches);
    CHECK(hk_destroy(ctx));
    return (finite && centre > 0.05f && corner == 0.0f && same && st.rays_traced >= (unsigned long long)W * H * 4) ? 0 : 20;
}
