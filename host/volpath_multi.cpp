// host/volpath_multi.cpp -- single-process multi-GPU C++ host over the handle-based C ABI (include/volpath.h).
//
// The reference is single-GPU (one process, device 0, file-scope statics: src/volumeRender_kernel.cu:337-352); its
// host loop (src/volumeRender.cpp:613-653) adds one frame per launch into ONE float4 sum.  A path-sample is addressed
// by (x, y, frame) (src/sampler.h:35-43), so this host gives GPU r of G the frames r, r + G, r + 2G, ... of every pixel
// (vp_render with frame_stride = G), each into its own float4[W*H] sum on its own device, all devices in flight at
// once, and combines the sums on device 0 with vp_reduce (NCCL over NVLink; peer copies when libnccl is absent).
// No CUDA runtime calls of its own: device memory comes from the library's plain-C helpers.
//
//   volpath_multi [--devices 0,1,..] [--blob N | --cloud NX NY NZ] [--size W H] [--spp N] [--density D] [--dump out.f32]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/volpath.h"

static void die(const char* what)
{
    fprintf(stderr, "volpath_multi: %s\n", what);
    exit(1);
}
#define VP(x)                                                                   \
    do {                                                                        \
        if ((x) != VP_OK) { fprintf(stderr, "%s: %s\n", #x, vp_last_error()); exit(1); } \
    } while (0)

int main(int argc, char** argv)
{
    vp_param P;
    P.width = 960; P.height = 512; P.density = 800; P.brightness = 1.0f;  // main() defaults (volumeRender.cpp:1286-1292)
    P.albedo = {1, 1, 1}; P.g = 0.877f; P.sigma_t = {1, 1, 1};
    std::vector<int> devices = {0};
    int         blob = 64, spp = 16, cloud[3] = {0, 0, 0};
    std::string dump;
    for (int i = 1; i < argc; i++)
    {
        std::string a = argv[i];
        auto next = [&]() { if (i + 1 >= argc) die("missing argument value"); return argv[++i]; };
        if (a == "--devices")
        {
            devices.clear();
            for (char* t = strtok(next(), ","); t; t = strtok(nullptr, ",")) devices.push_back(atoi(t));
        }
        else if (a == "--blob") blob = atoi(next());
        else if (a == "--cloud") { cloud[0] = atoi(next()); cloud[1] = atoi(next()); cloud[2] = atoi(next()); }
        else if (a == "--size") { P.width = atoi(next()); P.height = atoi(next()); }
        else if (a == "--spp") spp = atoi(next());
        else if (a == "--density") P.density = (float)atof(next());
        else if (a == "--dump") dump = next();
        else die("unknown option");
    }
    const int G = (int)devices.size();
    if (G < 1) die("no devices");

    // a smooth analytic test volume (the same one host/volpath_host.cpp makes) unless a synthetic cloud is asked for
    int nx = blob, ny = (blob * 2) / 3, nz = (blob * 5) / 4;
    std::vector<float> vol;
    if (!cloud[0])
    {
        vol.resize((size_t)nx * ny * nz);
        for (int k = 0; k < nz; k++)
            for (int j = 0; j < ny; j++)
                for (int i = 0; i < nx; i++)
                {
                    float x = (2 * i + 1.0f) / nx - 1, y = (2 * j + 1.0f) / ny - 1, z = (2 * k + 1.0f) / nz - 1;
                    float r2 = x * x + y * y + z * z;
                    float w  = 0.5f + 0.5f * sinf(9 * x) * sinf(7 * y + 1) * sinf(8 * z + 2);
                    float v  = 1.3f - 1.9f * r2 + 0.35f * w - 0.25f;
                    vol[((size_t)k * ny + j) * nx + i] = v < 0 ? 0 : (v > 1 ? 1 : v);
                }
    }
    // two-colour test environment (volumeRender.cpp:1372-1385), default sun (setup_sunsky(0.5, 0.2)), default camera
    const int          envw = 16, envh = 8;
    std::vector<float> env((size_t)envw * envh * 4);
    for (int j = 0; j < envh; j++)
        for (int i = 0; i < envw; i++)
        {
            float* e = &env[(size_t)(i + j * envw) * 4];
            e[0] = 0.03f; e[1] = j < 5 ? 0.07f : 0.03f; e[2] = j < 5 ? 0.23f : 0.03f; e[3] = 1.0f;
        }
    const float sun_dir[3] = {-2.7e-8f, 0.951057f, -0.309017f}, sun_power[3] = {51797.3f, 42480.1f, 32578.5f};
    // rows of inverse(lookAt(eye, eye + 4 fwd, up)) (volumeRender.cpp:108-112, 617-623): s, u, -f, eye
    const float view[12] = {0.0f, 0.207912f, 0.978148f, 3.922986f, 0.0f, 0.978148f, -0.207912f, -0.782739f, -1.0f, 0.0f, 0.0f, 0.03f};

    std::vector<vp_context*> ctx(G, nullptr);
    std::vector<void*>       sums(G, nullptr);
    const size_t bytes = (size_t)P.width * P.height * 16;
    for (int r = 0; r < G; r++)
    {
        VP(vp_create(devices[r], &ctx[r]));  // makes devices[r] current
        if (cloud[0])
            VP(vp_generate_cloud(ctx[r], cloud[0], cloud[1], cloud[2], 0, VP_VOXEL_F32, nullptr, nullptr, VP_BOUNDS_CELL, 0));
        else
            VP(vp_upload_volume(ctx[r], vol.data(), nx, ny, nz, VP_VOXEL_F32, VP_VOXEL_F32, VP_MEM_HOST, nullptr, nullptr, VP_BOUNDS_CELL));
        VP(vp_set_filter(ctx[r], 1));
        VP(vp_set_envmap(ctx[r], env.data(), envw, envh));
        VP(vp_set_sun(ctx[r], sun_dir, sun_power));
        VP(vp_set_inv_view(ctx[r], view));
        if (spp > 11) VP(vp_precompute_opacity(ctx[r], sun_dir));
        sums[r] = vp_dev_alloc(bytes);  // on devices[r] (current), zero-filled
        if (!sums[r]) die("out of device memory");
    }
    auto t0 = std::chrono::high_resolution_clock::now();
    // frames f = r, r + G, ... < spp on GPU r; every launch is asynchronous, so all GPUs render at once
    for (int r = 0; r < G; r++)
    {
        const int count = spp > r ? (spp - r + G - 1) / G : 0;
        VP(vp_render(ctx[r], sums[r], r, count, G, &P, VP_MODE_FAST, nullptr));
    }
    VP(vp_reduce(ctx.data(), sums.data(), G, (int)(P.width * P.height), 0));  // synchronises every device
    double us = std::chrono::duration<double, std::micro>(std::chrono::high_resolution_clock::now() - t0).count();
    printf("%f M samples / s, %d x %d, %d spp, %d GPU(s), nccl %d\n", (double)P.width * P.height * spp / us, P.width, P.height, spp, G,
           vp_nccl_available());
    if (!dump.empty())
    {
        std::vector<float> h((size_t)P.width * P.height * 4);
        if (vp_dev_to_host(h.data(), sums[0], bytes) != 0) die("device to host copy failed");
        FILE* fp = fopen(dump.c_str(), "wb");
        if (!fp) die("cannot write dump");
        fwrite(h.data(), 16, (size_t)P.width * P.height, fp);
        fclose(fp);
    }
    for (int r = 0; r < G; r++)
    {
        vp_dev_free(sums[r]);
        vp_destroy(ctx[r]);
    }
    return 0;
}
