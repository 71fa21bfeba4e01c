// host/volpath_host.cpp -- headless C++ host for the render hot path.
//
// The reference's host side is C++ (src/volumeRender.cpp); this file restates the part of it that drives the
// path -- main() (:1284-1403), cuda_volpath() (:613-653), update_sunsky() (:276-345) and capture() (:585-610) --
// without the GLUT/GL shell, and calls NOTHING but the reference's own 14 extern "C" entry points, declared
// below exactly as the reference declares them (:117-128, :347-356).  The same object file therefore links
// against either implementation of that boundary:
//     make -C host            -> volpath_host      (libvolpath_b200.so, this repo's CUDA path)
//     make -C host ref        -> volpath_host_ref  (oracle/_ref/libvolpath_ref_cuda.so, the reference kernel)
// tests/test_gpu_host_driver.py runs both on the same inputs and compares the dumped float4 sums.
//
//   volpath_host [--volume file.bin | --blob N] [--quantized] [--size W H] [--spp N] [--env file.f32 W H]
//                [--density D] [--albedo A] [--g G] [--point] [--fast] [--dump out.f32] [--ppm out.ppm]
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

// src/param.h:4-12
struct Param
{
    unsigned int width, height;
    float        density, brightness;
    float3       albedo;
    float        g;
    float3       sigma_t;
};

// the boundary (src/volumeRender.cpp:117-128, 347-356)
extern "C" void render_kernel(dim3 gridSize, dim3 blockSize, float4* d_output, int spp, const Param& p);
extern "C" void copy_inv_view_matrix(float* invViewMatrix, size_t sizeofMatrix);
extern "C" void copy_inv_model_matrix(float* invModelMatrix, size_t sizeofMatrix);
extern "C" void init_rng(dim3 gridSize, dim3 blockSize, int width, int height);
extern "C" void free_rng();
extern "C" void scale(float4* dst, float4* src, int size, float scale);
extern "C" void gamma_correct(float4* dst, float4* src, int size, float scale, float gamma);
extern "C" void init_envmap(const float4* data, int width, int height);
extern "C" void free_envmap();
extern "C" void set_sun(float* sun_dir, float* sun_power);
extern "C" void precompute_opacity(const float* light_dir);
extern "C" void init_cuda(void* h_volume, cudaExtent volumeSize, bool quantized, const float3* boxmin, const float3* boxmax);
extern "C" void set_texture_filter_mode(bool bLinearFilter);
extern "C" void free_cuda_buffers();
#ifdef VOLPATH_B200
extern "C" void        vp_shim_set_mode(int mode);  // 0 = parity (drop-in identical), 1 = fast
extern "C" const char* vp_last_error(void);
// the handle-based core behind the shims (include/volpath.h): frames [first, first + n) of every pixel in ONE launch
struct vp_context;
extern "C" vp_context* vp_shim_context(void);
extern "C" int vp_render(vp_context* ctx, void* d_sum, int first_frame, int n_frames, int frame_stride, const Param* p, int mode,
                         void* stream);
#endif

static void die(const char* what)
{
    fprintf(stderr, "volpath_host: %s\n", what);
    exit(1);
}
#define CK(x)                                                                       \
    do {                                                                            \
        cudaError_t e_ = (x);                                                       \
        if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } \
    } while (0)

// loadBinaryFile (src/volumeRender.cpp:915-965): int nx, ny, nz + nx*ny*nz floats; optional uchar quantisation
static void* load_bin(const char* fn, int& nx, int& ny, int& nz, bool quantized)
{
    FILE* fp = fopen(fn, "rb");
    if (!fp) die("cannot open volume file");
    if (fread(&nx, 4, 1, fp) != 1 || fread(&ny, 4, 1, fp) != 1 || fread(&nz, 4, 1, fp) != 1 || nx <= 0 || ny <= 0 || nz <= 0)
        die("bad volume header");
    size_t total = (size_t)nx * ny * nz;
    float* f     = (float*)malloc(total * sizeof(float));
    if (fread(f, sizeof(float), total, fp) != total) die("short volume file");
    fclose(fp);
    if (!quantized) return f;
    unsigned char* q = (unsigned char*)malloc(total);
    for (size_t i = 0; i < total; i++) q[i] = (unsigned char)(std::max(0.0f, std::min(f[i], 1.0f)) * 255.0f);
    free(f);
    return q;
}

// a smooth analytic test volume when no file is given
static void* make_blob(int n, int& nx, int& ny, int& nz, bool quantized)
{
    nx = n; ny = (n * 2) / 3; nz = (n * 5) / 4;
    size_t total = (size_t)nx * ny * nz;
    float* f     = (float*)malloc(total * sizeof(float));
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++)
            {
                float x = (2 * i + 1.0f) / nx - 1, y = (2 * j + 1.0f) / ny - 1, z = (2 * k + 1.0f) / nz - 1;
                float r2 = x * x + y * y + z * z;
                float w  = 0.5f + 0.5f * sinf(9 * x) * sinf(7 * y + 1) * sinf(8 * z + 2);
                float v  = 1.3f - 1.9f * r2 + 0.35f * w - 0.25f;
                f[((size_t)k * ny + j) * nx + i] = std::max(0.0f, std::min(1.0f, v));
            }
    if (!quantized) return f;
    unsigned char* q = (unsigned char*)malloc(total);
    for (size_t i = 0; i < total; i++) q[i] = (unsigned char)(f[i] * 255.0f);
    free(f);
    return q;
}

// glm::lookAt (RH) -> inverse -> transpose, first three rows (src/volumeRender.cpp:617-623): columns s, u, -f, eye
static void inv_view_rows(const float eye[3], const float fwd[3], const float up[3], float focus, float m[12])
{
    float c[3] = {eye[0] + fwd[0] * focus, eye[1] + fwd[1] * focus, eye[2] + fwd[2] * focus};
    float f[3] = {c[0] - eye[0], c[1] - eye[1], c[2] - eye[2]};
    float fl   = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    for (float& v : f) v /= fl;
    float s[3] = {f[1] * up[2] - f[2] * up[1], f[2] * up[0] - f[0] * up[2], f[0] * up[1] - f[1] * up[0]};
    float sl   = sqrtf(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    for (float& v : s) v /= sl;
    float u[3] = {s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]};
    for (int r = 0; r < 3; r++)
    {
        m[4 * r + 0] = s[r]; m[4 * r + 1] = u[r]; m[4 * r + 2] = -f[r]; m[4 * r + 3] = eye[r];
    }
}

int main(int argc, char** argv)
{
    // main() defaults (src/volumeRender.cpp:1286-1292); the Mat() table ends on the white preset (:1308)
    Param P;
    P.width = 960; P.height = 512; P.density = 800; P.brightness = 1.0f;
    P.albedo = make_float3(1, 1, 1); P.g = 0.877f; P.sigma_t = make_float3(1, 1, 1);
    std::string volume, dump, ppm, envfile;
    int  blob = 64, spp = 16, envw = 0, envh = 0, batch = 1;
    bool quantized = false, linear = true, fast = false;
    for (int i = 1; i < argc; i++)
    {
        std::string a = argv[i];
        auto next = [&]() { if (i + 1 >= argc) die("missing argument value"); return argv[++i]; };
        if (a == "--volume") volume = next();
        else if (a == "--blob") blob = atoi(next());
        else if (a == "--quantized") quantized = true;
        else if (a == "--size") { P.width = atoi(next()); P.height = atoi(next()); }
        else if (a == "--spp") spp = atoi(next());
        else if (a == "--env") { envfile = next(); envw = atoi(next()); envh = atoi(next()); }
        else if (a == "--density") P.density = (float)atof(next());
        else if (a == "--albedo") { float v = (float)atof(next()); P.albedo = make_float3(v, v, v); }
        else if (a == "--g") P.g = (float)atof(next());
        else if (a == "--point") linear = false;
        else if (a == "--fast") fast = true;
        else if (a == "--batch") batch = atoi(next());  // frames per launch (volpath-b200 build): what a host does for throughput
        else if (a == "--dump") dump = next();
        else if (a == "--ppm") ppm = next();
        else die("unknown option");
    }
#ifdef VOLPATH_B200
    vp_shim_set_mode(fast ? 1 : 0);
#else
    if (fast) die("--fast exists only in the volpath-b200 build");
    if (batch != 1) die("--batch exists only in the volpath-b200 build");
#endif
    if (batch < 1) die("--batch must be >= 1");

    // volume -> init_cuda (H.cpp:1332-1344); the caller frees the host copy after the call
    int   nx, ny, nz;
    void* h_volume = volume.empty() ? make_blob(blob, nx, ny, nz, quantized) : load_bin(volume.c_str(), nx, ny, nz, quantized);
    init_cuda(h_volume, make_cudaExtent(nx, ny, nz), quantized, nullptr, nullptr);
    free(h_volume);
    set_texture_filter_mode(linear);
    float identity[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    copy_inv_model_matrix(identity, sizeof(float4) * 3);  // H.cpp:1350-1353

    // environment + sun: a raw float4 lat-long map (e.g. the baked Hosek sky) or the two-colour test map (H.cpp:1372-1385)
    std::vector<float4> env;
    if (!envfile.empty())
    {
        env.resize((size_t)envw * envh);
        FILE* fp = fopen(envfile.c_str(), "rb");
        if (!fp || fread(env.data(), sizeof(float4), env.size(), fp) != env.size()) die("cannot read env map");
        fclose(fp);
    }
    else
    {
        envw = 16; envh = 8;
        env.resize((size_t)envw * envh);
        for (int j = 0; j < envh; j++)
            for (int i = 0; i < envw; i++) env[i + j * envw] = j < 5 ? make_float4(0.03f, 0.07f, 0.23f, 1) : make_float4(0.03f, 0.03f, 0.03f, 1);
    }
    init_envmap(env.data(), envw, envh);
    float sun_dir[3]   = {-2.7e-8f, 0.951057f, -0.309017f};     // setup_sunsky(0.5, 0.2) (H.cpp:1388-1390)
    float sun_power[3] = {51797.3f, 42480.1f, 32578.5f};        // sunColor * sunsky_scale
    set_sun(sun_dir, sun_power);

    // camera (H.cpp:108-112, 617-623)
    const float eye[3] = {3.922986f, -0.782739f, 0.030000f}, fwd[3] = {-0.978148f, 0.207912f, 0.0f}, up[3] = {0.207912f, 0.978148f, -0.0f};
    float       m[12];
    inv_view_rows(eye, fwd, up, 4.0f, m);
    copy_inv_view_matrix(m, sizeof(float4) * 3);

    // CudaFrameBuffer (H.cpp:358-389): the caller owns the float4 sum
    const int n = (int)(P.width * P.height);
    float4 *  d_sum = nullptr, *d_out = nullptr;
    CK(cudaMalloc(&d_sum, n * sizeof(float4)));
    CK(cudaMalloc(&d_out, n * sizeof(float4)));
    CK(cudaMemset(d_sum, 0, n * sizeof(float4)));
    dim3 blockSize(8, 8), gridSize((P.width + 7) / 8, (P.height + 7) / 8);  // H.cpp:100, 862
    init_rng(gridSize, blockSize, P.width, P.height);

    bool opacity_dirty = true;
    auto t0            = std::chrono::high_resolution_clock::now();
    for (int s = 0; s < spp; s++)  // display() / cuda_volpath() once per sample
    {
        if (s > 10 && opacity_dirty)  // update_sunsky (H.cpp:336-343)
        {
            precompute_opacity(sun_dir);
            opacity_dirty = false;
        }
#ifdef VOLPATH_B200
        if (batch > 1)
        {
            // one launch for frames s .. s + nb - 1; never across the frame-11 boundary, where the opacity table comes in
            int nb = std::min(batch, spp - s);
            if (s <= 10) nb = std::min(nb, 11 - s);
            if (vp_render(vp_shim_context(), d_sum, s, nb, 1, &P, fast ? 1 : 0, nullptr) != 0) die(vp_last_error());
            s += nb - 1;
            continue;
        }
#endif
        render_kernel(gridSize, blockSize, d_sum, s, P);
    }
    CK(cudaDeviceSynchronize());
    double us = std::chrono::duration<double, std::micro>(std::chrono::high_resolution_clock::now() - t0).count();
    printf("%f M samples / s, %d x %d, %d spp\n", (double)n * spp / us, P.width, P.height, spp);  // H.cpp:634-638

    std::vector<float4> h(n);
    if (!dump.empty())
    {
        CK(cudaMemcpy(h.data(), d_sum, n * sizeof(float4), cudaMemcpyDeviceToHost));
        FILE* fp = fopen(dump.c_str(), "wb");
        if (!fp) die("cannot write dump");
        fwrite(h.data(), sizeof(float4), n, fp);
        fclose(fp);
    }
    if (!ppm.empty())
    {
        gamma_correct(d_out, d_sum, n, 1.0f / spp, 2.2f);  // finalize_gamma (H.cpp:477-495)
        CK(cudaMemcpy(h.data(), d_out, n * sizeof(float4), cudaMemcpyDeviceToHost));
        FILE* fp = fopen(ppm.c_str(), "wb");  // Image::dump_ppm (src/image.cpp): P6, flipped like the GL view
        if (!fp) die("cannot write ppm");
        fprintf(fp, "P6\n%d %d\n255\n", P.width, P.height);
        for (int y = (int)P.height - 1; y >= 0; y--)
            for (unsigned x = 0; x < P.width; x++)
            {
                float4        c = h[x + (size_t)y * P.width];
                unsigned char b[3] = {(unsigned char)(std::min(1.0f, std::max(0.0f, c.x)) * 255.0f),
                                      (unsigned char)(std::min(1.0f, std::max(0.0f, c.y)) * 255.0f),
                                      (unsigned char)(std::min(1.0f, std::max(0.0f, c.z)) * 255.0f)};
                fwrite(b, 1, 3, fp);
            }
        fclose(fp);
    }
    free_rng();
    free_envmap();
    free_cuda_buffers();
    cudaFree(d_sum);
    cudaFree(d_out);
#ifdef VOLPATH_B200
    if (vp_last_error()[0]) { fprintf(stderr, "volpath_host: %s\n", vp_last_error()); return 2; }
#endif
    return 0;
}
