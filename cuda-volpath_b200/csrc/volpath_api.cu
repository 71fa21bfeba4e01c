// volpath_api.cu -- the C ABI of libvolpath_b200.so (include/volpath.h): a handle-based core (vp_*) and
// the 14 reference-named extern "C" shims the reference's host code binds (src/volumeRender.cpp:117-128,
// 347-356; defined in src/volumeRender_kernel.cu).  No torch, no C++ types in any signature.
#include <cub/device/device_scan.cuh>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <chrono>
#include <mutex>
#include <new>
#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "volpath_common.cuh"
#include "volpath_kernels.h"

using namespace vp;

#ifndef VP_DEFAULT_CELL_WINDOWS
#define VP_DEFAULT_CELL_WINDOWS 1  // coarse bound cells: 0 = union of the voxels' windows, 1 = the centre voxel's window, 2 = max from the union, min from the centre
#endif

namespace
{
thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define VP_CUDA(x)                                                                                     \
    do {                                                                                               \
        cudaError_t e_ = (x);                                                                          \
        if (e_ != cudaSuccess) return fail((int)e_, "%s: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define VP_TRY(x)            \
    do {                     \
        int r_ = (x);        \
        if (r_ != 0) return r_; \
    } while (0)

template <class T>
void dev_free(T*& p)
{
    if (p) cudaFree((void*)p);
    p = nullptr;
}
// a device temporary that is freed on every exit path (device out-of-memory is the EXPECTED failure of a large upload)
struct DevTmp
{
    void* p = nullptr;
    DevTmp() = default;
    DevTmp(const DevTmp&) = delete;
    DevTmp& operator=(const DevTmp&) = delete;
    ~DevTmp() { release(); }
    cudaError_t alloc(size_t bytes)
    {
        release();
        return cudaMalloc(&p, bytes ? bytes : 1);
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
    }
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};
bool env_flag(const char* name)
{
    const char* v = getenv(name);
    return v && atoi(v) != 0;
}
// cudaMalloc of the large pools, with the host time it took on stderr when VOLPATH_TIMING=1 (multi-process setups)
template <class T>
cudaError_t big_malloc(const char* what, T** p, size_t bytes)
{
    static const bool   timing = env_flag("VOLPATH_TIMING");
    const auto          t0     = std::chrono::steady_clock::now();
    const cudaError_t   e      = cudaMalloc((void**)p, bytes ? bytes : 32);
    if (timing)
        fprintf(stderr, "volpath: cudaMalloc %-12s %8.2f GB  %7.1f ms\n", what, bytes * 1e-9,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    return e;
}
struct StageTimer
{
    const char* what;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit StageTimer(const char* w) : what(w) {}
    ~StageTimer()
    {
        static const bool timing = env_flag("VOLPATH_TIMING");
        if (timing)
        {
            cudaDeviceSynchronize();
            fprintf(stderr, "volpath: stage %-24s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        }
    }
};
}  // namespace

struct vp_context
{
    int   device   = 0;
    int   num_sms  = 0;
    Scene S;
    // volume
    float*    dense       = nullptr;  // dense fp32 value copy (kept on request)
    uint2*    words       = nullptr;  // rank directory of the octet store
    uint32_t* table       = nullptr;  // flat slot table (small volumes only)
    uint32_t* slot_brick  = nullptr;
    void*     octets      = nullptr;
    float2*   bounds_voxel = nullptr;
    float2*   bounds_cell = nullptr;
    float*    opacity     = nullptr;  // bit-faithful apron table (parity)
    void*     opacity_oct = nullptr;  // fp16 octets (production renderers)
    float     opacity_ms  = 0.0f;     // device time of the last vp_precompute_opacity
    float*    sun_clear   = nullptr;
    uint16_t* top_jump    = nullptr;     // top level of the bound grid (staged in shared memory by the production renderers)
    uint32_t* bounds_half = nullptr;     // half-precision copies for the production renderers (coarse cells only)
    uint16_t* sun_clear_half = nullptr;
    float4*   env         = nullptr;
    std::vector<float> env_host;      // host copy of the env map (the CDF tables are built on the host, like init_envmap)
    bool               env_host_stale = false;  // the device map was baked in place (vp_bake_sunsky) since the last copy
    float*    env_cdf_y   = nullptr;
    float*    env_cdf_x   = nullptr;
    bool      env_sampling = false;   // the reference's PASSIVE_ENVMAP 0 variant
    float4*   host_acc    = nullptr;  // device accumulator of vp_render_to_host
    size_t    host_acc_bytes = 0;
    uint32_t  n_slots     = 0;
    size_t    n_bricks    = 0;
    size_t    octet_bytes = 0;
    int       bound_D     = 0;
    int       vac_margin  = 0;
    bool      have_volume = false;
    // instrumentation
    unsigned long long* d_stats = nullptr;
    // work-pool counters of the production renderers, one per launch in flight.  Slots rotate; each remembers the stream
    // of its last launch and an event recorded behind it, and a launch on ANOTHER stream first waits for that event
    // (cudaStreamWaitEvent): any number of caller streams is safe, launch n + kWorkSlots merely queues behind launch n.
    static constexpr int kWorkSlots = 16;
    unsigned long long* d_work  = nullptr;
    unsigned            work_rr = 0;
    cudaEvent_t         work_done[kWorkSlots] = {};
    cudaStream_t        work_stream[kWorkSlots] = {};
    bool                work_used[kWorkSlots] = {};
    // render_kernel shim, VP_MODE_FAST: consecutive one-frame launches go round-robin to four internal BLOCKING streams, so
    // the long tail of frame n (its last few paths) overlaps the bulk of frame n + 1 whenever the host does not
    // synchronise in between; blocking streams keep the legacy default-stream ordering the reference host relies on
    // (its copies, scale / gamma_correct and synchronisations still wait for every frame).  VOLPATH_SHIM_OVERLAP=0: off
    cudaStream_t        ring[8] = {};  // VOLPATH_SHIM_STREAMS of them in use (default 4)
    unsigned            ring_rr = 0;
    bool                stats_on = false;
    unsigned long long  launches = 0;
    cudaEvent_t         ev0 = nullptr, ev1 = nullptr;
    bool                timed = false;
    float               inv_model[12];
    // peer accumulators opened through CUDA IPC (vp_reduce_ipc): handle bytes -> mapped pointer
    std::vector<std::pair<std::string, void*>> ipc_open;
    void*               nccl_comm = nullptr;  // ncclComm_t of vp_nccl_init (one per context = per GPU)
    int                 nccl_rank = -1, nccl_ranks = 0;
};

static void scene_defaults(Scene& S)
{
    memset(&S, 0, sizeof(S));
    S.bmin = make_float3(-1, -1, -1);
    S.bmax = make_float3(1, 1, 1);
    S.l_inv = make_float3(0.5f, 0.5f, 0.5f);
    S.sun_dir = make_float3(0, 1, 0);
    S.sun_inv = make_float3(1.0f / 0.0f, 1.0f, 1.0f / 0.0f);
    const float id[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    memcpy(S.inv_view, id, sizeof(id));
    float fovx = 54.43;                                       // K.cu:1981
    S.cam_z    = (float)(-1.0f / tan(fovx * 0.00872664626));  // K.cu:1985, same double expression
    S.linear   = 0;                                           // K.cu:351 (the host flips it, H.cpp:1344)
}

static void free_volume(vp_context* c)
{
    dev_free(c->dense);
    dev_free(c->words);
    dev_free(c->table);
    dev_free(c->slot_brick);
    dev_free(c->octets);
    dev_free(c->bounds_voxel);
    dev_free(c->bounds_cell);
    dev_free(c->opacity);
    dev_free(c->opacity_oct);
    c->S.opacity_oct = nullptr;
    dev_free(c->sun_clear);
    c->S.sun_clear = nullptr;
    dev_free(c->bounds_half);
    dev_free(c->sun_clear_half);
    dev_free(c->top_jump);
    c->S.top_jump       = nullptr;
    c->S.bounds_half    = nullptr;
    c->S.sun_clear_half = nullptr;
    c->n_slots = 0;
    c->have_volume = false;
    c->S.brick_words = nullptr;
    c->S.brick_table = nullptr;
    c->S.octets = nullptr;
    c->S.bounds_voxel = nullptr;
    c->S.bounds_cell = nullptr;
    c->S.opacity = nullptr;
    c->S.have_opacity = 0;
    c->S.julia = 0;
}

static void set_box(vp_context* c, int nx, int ny, int nz, const float* bmin, const float* bmax)
{
    Scene& S = c->S;
    if (bmin && bmax)
    {
        S.bmin = make_float3(bmin[0], bmin[1], bmin[2]);
        S.bmax = make_float3(bmax[0], bmax[1], bmax[2]);
    }
    else
    {
        // K.cu:373-378
        S.bmin = make_float3(-1.0f, -(float)ny / (float)nx, -(float)nz / (float)nx);
        S.bmax = make_float3(1.0f, (float)ny / (float)nx, (float)nz / (float)nx);
    }
    S.l_inv    = make_float3(1.0f / (S.bmax.x - S.bmin.x), 1.0f / (S.bmax.y - S.bmin.y), 1.0f / (S.bmax.z - S.bmin.z));  // K.cu:313
    S.vs_scale = make_float3(S.l_inv.x * nx, S.l_inv.y * ny, S.l_inv.z * nz);
    S.vs_off   = make_float3(-S.bmin.x * S.vs_scale.x, -S.bmin.y * S.vs_scale.y, -S.bmin.z * S.vs_scale.z);
    S.vs_off_lin = make_float3(S.vs_off.x + 0.5f, S.vs_off.y + 0.5f, S.vs_off.z + 0.5f);
}

// host part of init_envmap for the PASSIVE_ENVMAP 0 variant (K.cu:1036-1070, 1144-1210; PRE_WARP 1): luminance * sin(phi),
// row CDFs + the CDF of the row sums, HDRpdfnormAlt.  Same float sequence as the reference's host code.
static float build_cdf_1d(const float* f, float* cdf, int size)
{
    float sum = 0.0f;
    for (int i = 0; i < size; i++) sum += f[i];
    float norm = 1.0f / sum;
    float I    = 0.0f;
    for (int i = 0; i < size; i++)
    {
        float p = f[i] * norm;
        I += p;
        cdf[i] = I;
    }
    cdf[size - 1] = 1.0f;
    return sum;
}
static int update_env_sampling(vp_context* c)
{
    Scene& S  = c->S;
    S.env_mis = 0;
    if (!c->env_sampling) return VP_OK;
    const int w = S.env_w, h = S.env_h;
    if (c->env_host.size() != (size_t)w * h * 4) return fail(VP_ERR_INVALID, "env sampling needs an environment map (init_envmap)");
    if (c->env_host_stale)
    {
        VP_CUDA(cudaMemcpy(c->env_host.data(), c->env, c->env_host.size() * sizeof(float), cudaMemcpyDeviceToHost));
        c->env_host_stale = false;
    }
    const size_t       total = (size_t)w * h;
    std::vector<float> lum(total), cdf_x(total), cdf_y(h), row_sum(h);
    for (size_t i = 0; i < total; i++)
    {
        const float* px = &c->env_host[i * 4];
        lum[i]          = (float)(px[0] * 0.2126 + px[1] * 0.7152 + px[2] * 0.0722);  // luminance(float4), K.cu:946
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
        {
            float phi = kPi * (y + 0.5f) / h;  // K.cu:1158
            lum[x + (size_t)y * w] *= sinf(phi);
        }
    float lumsum = 0.0f;
    for (size_t i = 0; i < total; i++) lumsum += lum[i];
    const float k1TwoPiPi = 1.0f / kPi / kTwoPi;  // vecmath.h:16
    S.env_pdfnorm_alt     = (float)w * (float)h * k1TwoPiPi / lumsum;
    for (int y = 0; y < h; y++) row_sum[y] = build_cdf_1d(lum.data() + (size_t)y * w, cdf_x.data() + (size_t)y * w, w);
    build_cdf_1d(row_sum.data(), cdf_y.data(), h);
    DevTmp nx_, ny_;
    VP_CUDA(nx_.alloc(total * sizeof(float)));
    VP_CUDA(ny_.alloc((size_t)h * sizeof(float)));
    VP_CUDA(cudaMemcpy(nx_.p, cdf_x.data(), total * sizeof(float), cudaMemcpyHostToDevice));
    VP_CUDA(cudaMemcpy(ny_.p, cdf_y.data(), (size_t)h * sizeof(float), cudaMemcpyHostToDevice));
    VP_CUDA(cudaDeviceSynchronize());
    dev_free(c->env_cdf_x);
    dev_free(c->env_cdf_y);
    c->env_cdf_x = nx_.as<float>(); nx_.p = nullptr;
    c->env_cdf_y = ny_.as<float>(); ny_.p = nullptr;
    S.env_cdf_x = c->env_cdf_x;
    S.env_cdf_y = c->env_cdf_y;
    S.env_mis   = 1;
    return VP_OK;
}

// (re)build the sun-clear distances of the fast renderer: needs the bound grid and the sun direction
static int update_sun_clear(vp_context* c)
{
    Scene& S = c->S;
    if (!c->have_volume || S.julia || !c->bounds_cell) return VP_OK;
    const size_t cells = (size_t)S.ncx * S.ncy * S.ncz;
    VP_CUDA(cudaDeviceSynchronize());  // renders on caller streams may still be reading the tables rewritten below
    if (!c->sun_clear) VP_CUDA(cudaMalloc(&c->sun_clear, cells * sizeof(float)));
    const float cell = (float)(1 << S.cell_log2);
    const float wx = cell / S.vs_scale.x, wy = cell / S.vs_scale.y, wz = cell / S.vs_scale.z;  // world cell extents
    const float step = 0.5f * fminf(wx, fminf(wy, wz));
    S.sun_clear    = nullptr;
    // neighbour rings a vacuum sample needs on top of its own +-D window (see k_sun_clear): 0 except on tiny grids
    const float need_vox = 0.75f * cell + 1.0f;
    const int   ring     = (float)c->bound_D >= need_vox ? 0 : (int)ceilf((need_vox - (float)c->bound_D) / cell);
    VP_CUDA(launch_sun_clear(S, S.sun_dir, step, ring, c->sun_clear, 0));
    c->launches++;
    VP_CUDA(cudaDeviceSynchronize());
    S.sun_clear    = c->sun_clear;
    S.clear_margin = 0.25f * step;
    S.sun_clear_half = nullptr;
    if (c->bounds_half)
    {
        if (!c->sun_clear_half && cudaMalloc(&c->sun_clear_half, cells * sizeof(uint16_t)) != cudaSuccess)
        {
            cudaGetLastError();  // no room for the copy: both tables stay float (the renderers take the pair together)
            dev_free(c->bounds_half);
            S.bounds_half = nullptr;
            return VP_OK;
        }
        VP_CUDA(launch_pack_clear_half(c->sun_clear, c->sun_clear_half, cells, 0));
        c->launches++;
        VP_CUDA(cudaDeviceSynchronize());
        S.sun_clear_half = c->sun_clear_half;
    }
    return VP_OK;
}

// dense fp32 value volume on the device -> bound grids + octet store.  Order matters for the peak footprint: the
// per-voxel bound sweeps need the dense copy plus two float2 volumes (C2: 26 + 2 x 53 GB), the octet pool (54 GB) is
// allocated only after their temporary is gone, so that even the parity-capable context of the full C2 grid fits 180 GB.
static int build_from_dense_impl(vp_context* c, int nx, int ny, int nz, int store_voxel, int bounds_flags, int keep_dense)
{
    Scene& S = c->S;
    S.nx = nx; S.ny = ny; S.nz = nz;
    S.nbx = (nx + 1 + kBrick - 1) / kBrick; S.nby = (ny + 1 + kBrick - 1) / kBrick; S.nbz = (nz + 1 + kBrick - 1) / kBrick;
    S.voxel_type = store_voxel;
    const size_t nb = (size_t)S.nbx * S.nby * S.nbz;
    c->n_bricks     = nb;

    // 1. bounds.  D as the reference computes it (H.cpp:1098-1101)
    float cell_size = 2.0f / (float)nx;
    int   D         = (int)ceil(kSearchRadius / cell_size);
    c->bound_D      = D;
    const size_t N  = (size_t)nx * ny * nz;
    // fast-renderer bound grid: cell edge = largest power of two <= max(1, D/6) voxels (<= 8)
    int cl = 0;
    while (cl < 3 && (2 << cl) * 6 <= D) cl++;
    // up to 128 Mi voxels (1 GiB of bounds) the fast renderer simply uses the reference's own per-voxel windows: the
    // reference estimator is biased by construction and its expectation moves with the window
    if ((bounds_flags & VP_BOUNDS_EXACT) || N <= ((size_t)128 << 20)) cl = 0;
    if (const char* f = getenv("VOLPATH_FORCE_CELL_LOG2")) cl = atoi(f) < 0 ? 0 : (atoi(f) > 3 ? 3 : atoi(f));  // tests / experiments
    const int cell = 1 << cl;
    S.cell_log2    = cl;
    S.ncx = (nx + cell - 1) >> cl; S.ncy = (ny + cell - 1) >> cl; S.ncz = (nz + cell - 1) >> cl;
    S.cs_scale = make_float3(S.vs_scale.x / cell, S.vs_scale.y / cell, S.vs_scale.z / cell);
    S.cs_off   = make_float3(S.vs_off.x / cell, S.vs_off.y / cell, S.vs_off.z / cell);
    if (bounds_flags & VP_BOUNDS_VOXEL)
    {
        DevTmp t0;
        VP_CUDA(cudaMalloc(&c->bounds_voxel, N * sizeof(float2)));
        VP_CUDA(t0.alloc(N * sizeof(float2)));
        VP_CUDA(launch_bounds_axis_f32(c->dense, c->bounds_voxel, nx, ny, nz, 0, D, 1, 0));
        VP_CUDA(launch_bounds_axis(c->bounds_voxel, t0.as<float2>(), nx, ny, nz, 1, D, 1, 0));
        VP_CUDA(launch_bounds_axis(t0.as<float2>(), c->bounds_voxel, nx, ny, nz, 2, D, 1, 0));
        VP_CUDA(cudaDeviceSynchronize());
    }
    if ((bounds_flags & VP_BOUNDS_CELL) && cell == 1)
    {
        // always a separate array: the fast grid carries vacuum jump distances in its (negative) max field
        VP_CUDA(cudaMalloc(&c->bounds_cell, N * sizeof(float2)));
        if (c->bounds_voxel)
            VP_CUDA(cudaMemcpy(c->bounds_cell, c->bounds_voxel, N * sizeof(float2), cudaMemcpyDeviceToDevice));
        else
        {
            DevTmp t0;
            VP_CUDA(t0.alloc(N * sizeof(float2)));
            VP_CUDA(launch_bounds_axis_f32(c->dense, c->bounds_cell, nx, ny, nz, 0, D, 1, 0));
            VP_CUDA(launch_bounds_axis(c->bounds_cell, t0.as<float2>(), nx, ny, nz, 1, D, 1, 0));
            VP_CUDA(launch_bounds_axis(t0.as<float2>(), c->bounds_cell, nx, ny, nz, 2, D, 1, 0));
            VP_CUDA(cudaDeviceSynchronize());
        }
    }
    else if (bounds_flags & VP_BOUNDS_CELL)
    {
        DevTmp t0, t1;
        VP_CUDA(t0.alloc((size_t)S.ncx * ny * nz * sizeof(float2)));
        VP_CUDA(t1.alloc((size_t)S.ncx * S.ncy * nz * sizeof(float2)));
        VP_CUDA(cudaMalloc(&c->bounds_cell, (size_t)S.ncx * S.ncy * S.ncz * sizeof(float2)));
        VP_CUDA(launch_bounds_axis_f32(c->dense, t0.as<float2>(), nx, ny, nz, 0, D, cell, 0));
        VP_CUDA(launch_bounds_axis(t0.as<float2>(), t1.as<float2>(), S.ncx, ny, nz, 1, D, cell, 0));
        VP_CUDA(launch_bounds_axis(t1.as<float2>(), c->bounds_cell, S.ncx, S.ncy, nz, 2, D, cell, 0));
        // values: the reference's window of the cell's centre voxel; vacuum classification: the union window above
        // (k_merge_cell_bounds; VOLPATH_UNION_WINDOWS=1 keeps the union values, for the bias measurement in DESIGN.md)
        const char* wm = getenv("VOLPATH_CELL_WINDOWS");  // union | centre | mixed (measurements: DESIGN.md section 2)
        const int   window_mode = !wm ? VP_DEFAULT_CELL_WINDOWS : (!strcmp(wm, "union") ? 0 : (!strcmp(wm, "centre") ? 1 : 2));
        if (window_mode != 0)
        {
            DevTmp cb;
            const size_t cells = (size_t)S.ncx * S.ncy * S.ncz;
            VP_CUDA(cb.alloc(cells * sizeof(float2)));
            VP_CUDA(launch_bounds_axis_f32(c->dense, t0.as<float2>(), nx, ny, nz, 0, D, cell, 0, 1));
            VP_CUDA(launch_bounds_axis(t0.as<float2>(), t1.as<float2>(), S.ncx, ny, nz, 1, D, cell, 0, 1));
            VP_CUDA(launch_bounds_axis(t1.as<float2>(), cb.as<float2>(), S.ncx, S.ncy, nz, 2, D, cell, 0, 1));
            VP_CUDA(launch_merge_cell_bounds(c->bounds_cell, cb.as<float2>(), cells, window_mode == 2, 0));
            VP_CUDA(cudaDeviceSynchronize());
        }
        VP_CUDA(cudaDeviceSynchronize());
    }
    VP_CUDA(cudaDeviceSynchronize());

    // 2. brick classification + exclusive scan -> slots
    {
        DevTmp flags, scan, tmp;
        size_t tmp_bytes = 0;
        VP_CUDA(flags.alloc(nb * 4));
        VP_CUDA(scan.alloc(nb * 4));
        VP_CUDA(launch_classify_bricks(c->dense, nx, ny, nz, S.nbx, S.nby, S.nbz, flags.as<uint32_t>(), 0));
        VP_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flags.as<uint32_t>(), scan.as<uint32_t>(), (int)nb, 0));
        VP_CUDA(tmp.alloc(tmp_bytes));
        VP_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, flags.as<uint32_t>(), scan.as<uint32_t>(), (int)nb, 0));
        uint32_t last_scan = 0, last_flag = 0;
        VP_CUDA(cudaMemcpy(&last_scan, scan.as<uint32_t>() + nb - 1, 4, cudaMemcpyDeviceToHost));
        VP_CUDA(cudaMemcpy(&last_flag, flags.as<uint32_t>() + nb - 1, 4, cudaMemcpyDeviceToHost));
        c->n_slots = last_scan + last_flag;
        VP_CUDA(cudaMalloc(&c->words, ((nb + 31) / 32) * sizeof(uint2)));
        VP_CUDA(cudaMalloc(&c->slot_brick, (size_t)(c->n_slots ? c->n_slots : 1) * 4));
        // <= 4 MB of flat table stays cache-resident; larger grids (and VOLPATH_FORCE_RANK_DIR=1, which lets the tests
        // put small volumes on the large-volume path) look slots up in the rank directory
        if (nb * 4 <= (size_t)4 << 20 && !env_flag("VOLPATH_FORCE_RANK_DIR")) VP_CUDA(cudaMalloc(&c->table, nb * 4));
        VP_CUDA(launch_make_words(flags.as<uint32_t>(), scan.as<uint32_t>(), nb, c->words, c->slot_brick, c->table, 0));
        VP_CUDA(cudaDeviceSynchronize());
    }

    // 3. octet pool
    const size_t ob = store_voxel == kF32 ? 32 : (store_voxel == kF16 ? 16 : 8);
    c->octet_bytes  = (size_t)c->n_slots * kBrickCells * ob;
    VP_CUDA(big_malloc("octets", &c->octets, c->octet_bytes));
    StageTimer st_oct("fill octets");
    VP_CUDA(launch_fill_octets(c->dense, nx, ny, nz, S.nbx, S.nby, c->slot_brick, c->n_slots, c->octets, store_voxel, 0));
    VP_CUDA(cudaDeviceSynchronize());
    if (!keep_dense)
    {
        StageTimer st_free("free dense");
        dev_free(c->dense);
    }
    if (c->bounds_cell)
    {
        // vacuum jump distances (breadth-first dilation over the bound cells, up to 63 cells)
        const size_t cells = (size_t)S.ncx * S.ncy * S.ncz;
        DevTmp       tmp;
        VP_CUDA(tmp.alloc(cells));
        // world sizes: a voxel per axis, the smallest cell edge, what the +-D window covers at least, and what a 0.05
        // segment plus the trilinear footprint (and a voxel of slack) needs; the deficit becomes the fringe margin
        const float vx = 1.0f / S.vs_scale.x, vy = 1.0f / S.vs_scale.y, vz = 1.0f / S.vs_scale.z;
        const float cw = (float)cell * fminf(vx, fminf(vy, vz));
        const float cover  = (float)D * fminf(vx, fminf(vy, vz));
        const float need   = kSearchRadius + 2.0f * fmaxf(vx, fmaxf(vy, vz));
        int         margin = need > cover ? (int)ceilf((need - cover) / cw) : 0;
        if (margin > 32) margin = 32;
        c->vac_margin = margin;
        // breadth-first depth: 63 cells, fewer on very fine grids (each level is a pass over all cells)
        const int kmax = cells > ((size_t)1 << 30) ? 3 : (cells > ((size_t)160 << 20) ? 15 : 63);
        if (margin > kmax - 1) margin = kmax - 1;
        VP_CUDA(launch_vacuum_jumps(c->bounds_cell, tmp.as<uint8_t>(), S.ncx, S.ncy, S.ncz, kmax, margin, cw, 0));
        VP_CUDA(cudaDeviceSynchronize());
        // top level for shared-memory staging: blocks of t^3 cells, t the smallest power of two that fits kTopCellsMax
        int tl = 0;
        while ((size_t)((S.ncx + (1 << tl) - 1) >> tl) * ((S.ncy + (1 << tl) - 1) >> tl) * ((S.ncz + (1 << tl) - 1) >> tl) > (size_t)kTopCellsMax) tl++;
        S.top_log2 = tl;
        S.ntx = (S.ncx + (1 << tl) - 1) >> tl; S.nty = (S.ncy + (1 << tl) - 1) >> tl; S.ntz = (S.ncz + (1 << tl) - 1) >> tl;
        VP_CUDA(cudaMalloc(&c->top_jump, (size_t)S.ntx * S.nty * S.ntz * sizeof(uint16_t)));
        VP_CUDA(launch_build_top(c->bounds_cell, S.ncx, S.ncy, S.ncz, tl, S.ntx, S.nty, S.ntz, c->top_jump, 0));
        VP_CUDA(cudaDeviceSynchronize());
        S.top_jump = c->top_jump;
    }

    // coarse cells = a volume too large for per-voxel windows: keep the per-cell tables L2-resident in half precision
    // (VOLPATH_HALF_TABLES=0/1 overrides the choice, for measurements and tests)
    S.bounds_half = nullptr;
    {
        const char* ov   = getenv("VOLPATH_HALF_TABLES");
        const bool  want = ov ? atoi(ov) != 0 : S.cell_log2 > 0;
        if (c->bounds_cell && want)
        {
            const size_t cells = (size_t)S.ncx * S.ncy * S.ncz;
            // the copies are an optimisation: out of memory or values beyond the half range -> stay with the float tables
            DevTmp d_over;
            int    h_over = 1;
            if (d_over.alloc(sizeof(int)) == cudaSuccess && cudaMalloc(&c->bounds_half, cells * sizeof(uint32_t)) == cudaSuccess &&
                cudaMemset(d_over.p, 0, sizeof(int)) == cudaSuccess &&
                launch_pack_bounds_half(c->bounds_cell, c->bounds_half, cells, d_over.as<int>(), 0) == cudaSuccess)
            {
                c->launches++;
                if (cudaMemcpy(&h_over, d_over.p, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) h_over = 1;
            }
            cudaGetLastError();
            if (h_over) dev_free(c->bounds_half);
            S.bounds_half = c->bounds_half;
        }
    }
    S.brick_words  = c->words;
    S.stream_octets = c->octet_bytes > ((size_t)8 << 30) ? 1 : 0;  // pool >> 126 MB of L2: no reuse to protect
    S.brick_table  = c->table;
    S.octets       = c->octets;
    S.bounds_voxel = c->bounds_voxel;
    S.bounds_cell  = c->bounds_cell;
    S.opacity      = nullptr;
    S.have_opacity = 0;
    S.julia        = 0;
    c->have_volume = true;
    if (getenv("VOLPATH_PARITY_USES_CELL_BOUNDS") && c->bounds_voxel && c->bounds_cell)  // experiments only
    {
        VP_CUDA(launch_expand_cell_bounds(S, c->bounds_voxel, 0));
        VP_CUDA(cudaDeviceSynchronize());
    }
    return update_sun_clear(c);
}
static int build_from_dense(vp_context* c, int nx, int ny, int nz, int store_voxel, int bounds_flags, int keep_dense)
{
    const int rc = build_from_dense_impl(c, nx, ny, nz, store_voxel, bounds_flags, keep_dense);
    if (rc != VP_OK)
    {
        cudaGetLastError();
        free_volume(c);  // a failed build leaves no half-initialised volume behind
    }
    return rc;
}

extern "C" {

int vp_destroy(vp_context* c);
const char* vp_last_error(void) { return g_err; }
const char* vp_version(void) { return "volpath-b200 0.1 (sm_100a)"; }

static int create_impl(vp_context* c, int device)
{
    c->device = device;
    cudaDeviceProp prop;
    VP_CUDA(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    set_build_sm_count(c->num_sms);
    // measurement hook (device-wide hint): how much DRAM a random 32-byte sector read drags in (tools/sector_probe.cu)
    if (const char* g = getenv("VOLPATH_L2_FETCH_GRANULARITY"))
        if (atoi(g) > 0) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
    scene_defaults(c->S);
    const float id[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    memcpy(c->inv_model, id, sizeof(id));
    VP_CUDA(cudaMalloc(&c->d_stats, 16 * sizeof(unsigned long long)));
    VP_CUDA(cudaMemset(c->d_stats, 0, 16 * sizeof(unsigned long long)));
    VP_CUDA(cudaMalloc(&c->d_work, vp_context::kWorkSlots * sizeof(unsigned long long)));
    for (cudaEvent_t& e : c->work_done) VP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    VP_CUDA(cudaEventCreate(&c->ev0));
    VP_CUDA(cudaEventCreate(&c->ev1));
    // a 1x1 black environment until init_envmap is called
    VP_CUDA(cudaMalloc(&c->env, sizeof(float4)));
    VP_CUDA(cudaMemset(c->env, 0, sizeof(float4)));
    c->S.env = c->env; c->S.env_w = 1; c->S.env_h = 1;
    return VP_OK;
}

int vp_create(int device, vp_context** out)
{
    if (!out) return fail(VP_ERR_INVALID, "vp_create: null out");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return fail(VP_ERR_NO_DEVICE, "vp_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= n) return fail(VP_ERR_INVALID, "vp_create: device %d out of range (%d devices)", device, n);
    VP_CUDA(cudaSetDevice(device));
    vp_context* c = new (std::nothrow) vp_context();
    if (!c) return fail(VP_ERR_INVALID, "vp_create: out of host memory");
    const int rc = create_impl(c, device);
    if (rc != VP_OK)
    {
        vp_destroy(c);  // frees whatever was allocated; g_err keeps the first failure
        return rc;
    }
    *out = c;
    return VP_OK;
}

int vp_nccl_destroy(vp_context* c);
int vp_ipc_close(vp_context* c);
int vp_destroy(vp_context* c)
{
    if (!c) return VP_OK;
    cudaSetDevice(c->device);
    vp_nccl_destroy(c);
    vp_ipc_close(c);
    free_volume(c);
    dev_free(c->env);
    dev_free(c->env_cdf_x);
    dev_free(c->env_cdf_y);
    dev_free(c->d_stats);
    dev_free(c->d_work);
    for (cudaStream_t& r : c->ring)
        if (r) cudaStreamDestroy(r);
    for (cudaEvent_t& e : c->work_done)
        if (e) cudaEventDestroy(e);
    dev_free(c->host_acc);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    delete c;
    return VP_OK;
}

int vp_free_volume(vp_context* c)
{
    if (!c) return fail(VP_ERR_INVALID, "null context");
    VP_CUDA(cudaSetDevice(c->device));
    free_volume(c);
    return VP_OK;
}

int vp_upload_volume(vp_context* c, const void* volume, int nx, int ny, int nz, int src_voxel, int store_voxel, int memspace,
                     const float* boxmin3, const float* boxmax3, int bounds_flags)
{
    if (!c) return fail(VP_ERR_INVALID, "null context");
    if (!volume) return fail(VP_ERR_NO_VOLUME, "cannot init without host volume");  // K.cu:360-364 (the reference exits)
    if (nx < 1 || ny < 1 || nz < 1 || nx > 8184 || ny > 8184 || nz > 8184) return fail(VP_ERR_INVALID, "bad volume dims %d %d %d", nx, ny, nz);
    if (src_voxel != VP_VOXEL_U8 && src_voxel != VP_VOXEL_F32) return fail(VP_ERR_UNSUPPORTED, "source voxels must be u8 or f32");
    if (store_voxel < VP_VOXEL_U8 || store_voxel > VP_VOXEL_F32) return fail(VP_ERR_INVALID, "bad store voxel type");
    if (!(bounds_flags & (VP_BOUNDS_VOXEL | VP_BOUNDS_CELL | VP_BOUNDS_EXACT))) return fail(VP_ERR_INVALID, "bounds_flags selects no bound grid");
    if (bounds_flags & VP_BOUNDS_EXACT) bounds_flags |= VP_BOUNDS_CELL;
    VP_CUDA(cudaSetDevice(c->device));
    free_volume(c);
    const size_t N = (size_t)nx * ny * nz;
    VP_CUDA(cudaMalloc(&c->dense, N * sizeof(float)));
    const cudaMemcpyKind kind = memspace == VP_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (src_voxel == VP_VOXEL_F32)
    {
        VP_CUDA(cudaMemcpy(c->dense, volume, N * sizeof(float), kind));
    }
    else
    {
        DevTmp d8;
        VP_CUDA(d8.alloc(N));
        VP_CUDA(cudaMemcpy(d8.p, volume, N, kind));
        VP_CUDA(launch_u8_to_f32(d8.as<uint8_t>(), c->dense, N, 0));
        VP_CUDA(cudaDeviceSynchronize());
    }
    set_box(c, nx, ny, nz, boxmin3, boxmax3);
    return build_from_dense(c, nx, ny, nz, store_voxel, bounds_flags, 0);
}

int vp_generate_cloud(vp_context* c, int nx, int ny, int nz, unsigned int seed, int store_voxel, const float* boxmin3,
                      const float* boxmax3, int bounds_flags, int keep_dense)
{
    if (!c) return fail(VP_ERR_INVALID, "null context");
    if (nx < 1 || ny < 1 || nz < 1 || nx > 8184 || ny > 8184 || nz > 8184) return fail(VP_ERR_INVALID, "bad volume dims %d %d %d", nx, ny, nz);
    if (store_voxel < VP_VOXEL_U8 || store_voxel > VP_VOXEL_F32) return fail(VP_ERR_INVALID, "bad store voxel type");
    if (!(bounds_flags & (VP_BOUNDS_VOXEL | VP_BOUNDS_CELL | VP_BOUNDS_EXACT))) return fail(VP_ERR_INVALID, "bounds_flags selects no bound grid");
    if (bounds_flags & VP_BOUNDS_EXACT) bounds_flags |= VP_BOUNDS_CELL;
    VP_CUDA(cudaSetDevice(c->device));
    free_volume(c);
    const size_t N = (size_t)nx * ny * nz;
    VP_CUDA(big_malloc("dense", &c->dense, N * sizeof(float)));
    {
        StageTimer st_cloud("fbm cloud");
        VP_CUDA(launch_fbm_cloud(c->dense, nx, ny, nz, seed, 0));
    }
    set_box(c, nx, ny, nz, boxmin3, boxmax3);
    return build_from_dense(c, nx, ny, nz, store_voxel, bounds_flags, keep_dense);
}

const void* vp_dense_volume(vp_context* c) { return c ? c->dense : nullptr; }
int         vp_release_dense(vp_context* c)
{
    if (!c) return fail(VP_ERR_INVALID, "null context");
    dev_free(c->dense);
    return VP_OK;
}

int vp_set_julia(vp_context* c)
{
    if (!c) return fail(VP_ERR_INVALID, "null context");
    VP_CUDA(cudaSetDevice(c->device));
    free_volume(c);
    Scene& S = c->S;
    S.nx = S.ny = S.nz = 32;  // H.cpp:1346 (extent only; the density is procedural)
    set_box(c, 1, 1, 1, nullptr, nullptr);
    S.julia        = 1;
    c->have_volume = true;
    return VP_OK;
}

int vp_set_filter(vp_context* c, int linear)
{
    if (!c) return fail(VP_ERR_INVALID, "null context");
    c->S.linear = linear ? 1 : 0;
    return VP_OK;
}

int vp_set_envmap(vp_context* c, const float* rgba, int width, int height)
{
    if (!c || !rgba || width < 1 || height < 1) return fail(VP_ERR_INVALID, "vp_set_envmap: bad arguments");
    VP_CUDA(cudaSetDevice(c->device));
    // allocate the new map first and swap it in on success: a failed upload leaves the old map in use
    DevTmp fresh;
    VP_CUDA(fresh.alloc((size_t)width * height * sizeof(float4)));
    VP_CUDA(cudaMemcpy(fresh.p, rgba, (size_t)width * height * sizeof(float4), cudaMemcpyHostToDevice));
    VP_CUDA(cudaDeviceSynchronize());  // no render may still be reading the old map when it is freed
    dev_free(c->env);
    c->env  = fresh.as<float4>();
    fresh.p = nullptr;
    c->S.env = c->env; c->S.env_w = width; c->S.env_h = height;
    c->env_host.assign(rgba, rgba + (size_t)width * height * 4);
    c->env_host_stale = false;
    return update_env_sampling(c);
}

int vp_bake_sunsky(vp_context* c, const vp_sky_state* st, int width, int height)
{
    if (!c || !st || width < 1 || height < 2) return fail(VP_ERR_INVALID, "vp_bake_sunsky: bad arguments");
    VP_CUDA(cudaSetDevice(c->device));
    if (!c->env || c->S.env_w != width || c->S.env_h != height)
    {
        DevTmp fresh;
        VP_CUDA(fresh.alloc((size_t)width * height * sizeof(float4)));
        VP_CUDA(cudaDeviceSynchronize());
        dev_free(c->env);
        c->env  = fresh.as<float4>();
        fresh.p = nullptr;
        c->S.env = c->env; c->S.env_w = width; c->S.env_h = height;
    }
    VP_CUDA(launch_bake_sunsky(*st, c->env, width, height, 0));
    c->launches++;
    c->S.env = c->env; c->S.env_w = width; c->S.env_h = height;
    // the host copy is only needed by the CDF build of the env-sampling variant
    c->env_host.resize((size_t)width * height * 4);
    if (c->env_sampling)
        VP_CUDA(cudaMemcpy(c->env_host.data(), c->env, c->env_host.size() * sizeof(float), cudaMemcpyDeviceToHost));
    else
        c->env_host_stale = true;
    VP_CUDA(cudaDeviceSynchronize());
    return update_env_sampling(c);
}

int vp_get_envmap(vp_context* c, float* h_out, int* wh2)
{
    if (!c || !wh2) return fail(VP_ERR_INVALID, "vp_get_envmap: bad arguments");
    wh2[0] = c->S.env_w; wh2[1] = c->S.env_h;
    if (!h_out || !c->env) return VP_OK;
    VP_CUDA(cudaSetDevice(c->device));
    VP_CUDA(cudaMemcpy(h_out, c->env, (size_t)c->S.env_w * c->S.env_h * sizeof(float4), cudaMemcpyDeviceToHost));
    return VP_OK;
}

int vp_set_env_sampling(vp_context* c, int enable)
{
    if (!c) return fail(VP_ERR_INVALID, "null context");
    VP_CUDA(cudaSetDevice(c->device));
    c->env_sampling = enable != 0;
    if (!c->env_sampling)
    {
        c->S.env_mis = 0;
        return VP_OK;
    }
    return update_env_sampling(c);
}

int vp_set_sun(vp_context* c, const float* dir3, const float* power3)
{
    if (!c || !dir3 || !power3) return fail(VP_ERR_INVALID, "vp_set_sun: bad arguments");
    // K.cu:1269-1283: disk radiance -> directional power, r = 0.45/94 (float)(double/float)
    Scene& S = c->S;
    S.sun_power_original = make_float3(power3[0], power3[1], power3[2]);
    float r     = (float)(0.45 / 94.0f);
    float scale = kPi * (r * r);
    S.sun_power = make_float3(S.sun_power_original.x * scale, S.sun_power_original.y * scale, S.sun_power_original.z * scale);
    S.sun_dir   = make_float3(dir3[0], dir3[1], dir3[2]);
    S.sun_inv   = make_float3(1.0f / S.sun_dir.x, 1.0f / S.sun_dir.y, 1.0f / S.sun_dir.z);  // +-inf for an axis-parallel sun, as the slab test expects
    VP_CUDA(cudaSetDevice(c->device));
    return update_sun_clear(c);
}

int vp_set_inv_view(vp_context* c, const float* m12)
{
    if (!c || !m12) return fail(VP_ERR_INVALID, "vp_set_inv_view: bad arguments");
    memcpy(c->S.inv_view, m12, 12 * sizeof(float));
    return VP_OK;
}

static int all_gather_bytes(vp_context* c, void* buf, size_t bytes_per_rank);
static int precompute_opacity_impl(vp_context* c, const float* dir3, bool sharded)
{
    if (!c || !dir3) return fail(VP_ERR_INVALID, "vp_precompute_opacity: bad arguments");
    if (!c->have_volume) return fail(VP_ERR_NO_VOLUME, "vp_precompute_opacity: no volume");
    VP_CUDA(cudaSetDevice(c->device));
    if (c->S.julia) return VP_OK;  // the table of the no-OpenVDB build is built from a zero density
    VP_CUDA(cudaDeviceSynchronize());
    dev_free(c->opacity);
    dev_free(c->opacity_oct);
    c->S.have_opacity = 0;
    c->S.opacity      = nullptr;
    c->S.opacity_oct  = nullptr;
    const float3 dir  = make_float3(dir3[0], dir3[1], dir3[2]);
    const size_t slots = c->n_slots ? c->n_slots : 1;
    struct Ev
    {
        cudaEvent_t e = nullptr;
        ~Ev() { if (e) cudaEventDestroy(e); }
    } ev0, ev1;
    VP_CUDA(cudaEventCreate(&ev0.e));
    VP_CUDA(cudaEventCreate(&ev1.e));
    cudaEvent_t t0 = ev0.e, t1 = ev1.e;
    cudaEventRecord(t0, 0);
    int rc = VP_OK;
    // (1) the bit-faithful per-voxel march (K.cu:483-524) for VP_MODE_PARITY: only contexts that can run that mode
    //     (VP_BOUNDS_VOXEL) pay for it; VOLPATH_OPACITY_FAITHFUL=1 forces it (tests compare the two tables)
    if (c->bounds_voxel || env_flag("VOLPATH_OPACITY_FAITHFUL"))
    {
        cudaError_t e = cudaMalloc(&c->opacity, slots * kOpBrickPad * sizeof(float));
        if (e == cudaSuccess) e = launch_precompute_opacity(c->S, c->slot_brick, c->n_slots, c->opacity, dir, 0);
        if (e != cudaSuccess) rc = fail((int)e, "vp_precompute_opacity (per-voxel march): %s", cudaGetErrorString(e));
        c->launches++;
    }
    // (2) the production table: swept build, fp16 octets (volpath_build.cu)
    if (rc == VP_OK && c->bounds_cell)
    {
        int K = 64;
        if (const char* k = getenv("VOLPATH_OPACITY_K")) K = atoi(k) > 0 ? atoi(k) : K;
        // sharded build (multi-GPU hosts, volume replicated): every rank sweeps the slots [r, r + 1) * per of the table and
        // an in-place ncclAllGather over NVLink hands everybody the rest (the checkpoint slabs are cheap and replicated)
        const int    G   = sharded && c->nccl_comm ? c->nccl_ranks : 1;
        const size_t per = (slots + G - 1) / G;
        cudaError_t  e   = big_malloc("opacity_oct", &c->opacity_oct, per * G * kBrickCells * 16);
        StageTimer   st_op("opacity sweep");
        const uint32_t b0 = G > 1 ? (uint32_t)std::min<size_t>(per * c->nccl_rank, c->n_slots) : 0u;
        const uint32_t b1 = G > 1 ? (uint32_t)std::min<size_t>(per * (c->nccl_rank + 1), c->n_slots) : c->n_slots;
        if (e == cudaSuccess) e = launch_opacity_octets(c->S, c->slot_brick, c->n_slots, c->opacity_oct, dir, K, 0, b0, b1);
        if (e != cudaSuccess) rc = fail((int)e, "vp_precompute_opacity (swept octets): %s", cudaGetErrorString(e));
        c->launches += 2;
        if (rc == VP_OK && G > 1) rc = all_gather_bytes(c, c->opacity_oct, per * kBrickCells * 16);
    }
    cudaEventRecord(t1, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc == VP_OK && e != cudaSuccess) rc = fail((int)e, "vp_precompute_opacity: %s", cudaGetErrorString(e));
    if (rc == VP_OK) cudaEventElapsedTime(&c->opacity_ms, t0, t1);
    if (rc != VP_OK)
    {
        cudaGetLastError();
        dev_free(c->opacity);
        dev_free(c->opacity_oct);
        return rc;
    }
    c->S.opacity      = c->opacity;
    c->S.have_opacity = c->opacity ? 1 : 0;
    c->S.opacity_oct  = c->opacity_oct;
    return VP_OK;
}

int vp_precompute_opacity(vp_context* c, const float* dir3) { return precompute_opacity_impl(c, dir3, false); }
int vp_precompute_opacity_sharded(vp_context* c, const float* dir3) { return precompute_opacity_impl(c, dir3, true); }

static int render_on(vp_context* c, void* d_sum, int first_frame, int n_frames, int frame_stride, const vp_param* p, int mode,
                     cudaStream_t st);

int vp_render(vp_context* c, void* d_sum, int first_frame, int n_frames, int frame_stride, const vp_param* p, int mode,
              vp_stream stream)
{
    if (!c) return fail(VP_ERR_INVALID, "vp_render: bad arguments");
    return render_on(c, d_sum, first_frame, n_frames, frame_stride, p, mode, (cudaStream_t)stream);
}

// the render_kernel shim's launch: see vp_context::ring
static int render_shim(vp_context* c, void* d_sum, int frame, const vp_param* p, int mode)
{
    static const bool overlap = !(getenv("VOLPATH_SHIM_OVERLAP") && atoi(getenv("VOLPATH_SHIM_OVERLAP")) == 0);
    if (mode != VP_MODE_FAST || !overlap) return vp_render(c, d_sum, frame, 1, 1, p, mode, nullptr);
    VP_CUDA(cudaSetDevice(c->device));
    static const unsigned n_ring = getenv("VOLPATH_SHIM_STREAMS") ? (unsigned)std::min(8, std::max(1, atoi(getenv("VOLPATH_SHIM_STREAMS")))) : 4u;
    const unsigned slot = c->ring_rr++ % n_ring;
    if (!c->ring[slot]) VP_CUDA(cudaStreamCreate(&c->ring[slot]));
    return render_on(c, d_sum, frame, 1, 1, p, mode, c->ring[slot]);
}

static int render_on(vp_context* c, void* d_sum, int first_frame, int n_frames, int frame_stride, const vp_param* p, int mode,
                     cudaStream_t st)
{
    if (!c || !d_sum || !p) return fail(VP_ERR_INVALID, "vp_render: bad arguments");
    if (!c->have_volume) return fail(VP_ERR_NO_VOLUME, "vp_render: no volume uploaded");
    if (n_frames <= 0) return VP_OK;
    if (p->width == 0 || p->height == 0 || p->width > 65535 || p->height > 65535) return fail(VP_ERR_INVALID, "vp_render: bad image size");
    VP_CUDA(cudaSetDevice(c->device));
    VP_CUDA(cudaEventRecord(c->ev0, st));
    if (mode == VP_MODE_PARITY)
    {
        if (!c->S.julia && !c->S.bounds_voxel) return fail(VP_ERR_INVALID, "vp_render: parity mode needs VP_BOUNDS_VOXEL");
        VP_CUDA(launch_render_parity(c->S, (float4*)d_sum, first_frame, n_frames, frame_stride, *p, st));
        c->launches++;
    }
    else if (mode == VP_MODE_FAST || mode == VP_MODE_WAVE)
    {
        if (!c->S.julia && !c->S.bounds_cell) return fail(VP_ERR_INVALID, "vp_render: fast mode needs VP_BOUNDS_CELL");
        // this launch's work-pool counter (see vp_context::d_work)
        const unsigned ws = c->work_rr++ % vp_context::kWorkSlots;
        if (c->work_used[ws] && c->work_stream[ws] != st) VP_CUDA(cudaStreamWaitEvent(st, c->work_done[ws], 0));
        unsigned long long* d_work = c->d_work + ws;
        // all frames of the call are ONE launch (one work pool), as long as tiles * frames fits 31 bits
        const long long tiles = (long long)((p->width + 7) / 8) * ((p->height + 3) / 4);
        long long       cap   = ((1ll << 31) - 1) / tiles;
        if (cap < 1) cap = 1;
        for (long long f = 0; f < n_frames; f += cap)
        {
            int nf = (int)(n_frames - f < cap ? n_frames - f : cap);
            if (mode == VP_MODE_WAVE)
                VP_CUDA(launch_render_wave(c->S, (float4*)d_sum, first_frame + (int)f * frame_stride, nf, frame_stride, *p, d_work,
                                           c->stats_on ? c->d_stats : nullptr, c->num_sms, st));
            else
                VP_CUDA(launch_render_fast(c->S, (float4*)d_sum, first_frame + (int)f * frame_stride, nf, frame_stride, *p, d_work,
                                           c->stats_on ? c->d_stats : nullptr, c->num_sms, st));
            c->launches++;
        }
        VP_CUDA(cudaEventRecord(c->work_done[ws], st));
        c->work_stream[ws] = st;
        c->work_used[ws]   = true;
    }
    else
        return fail(VP_ERR_INVALID, "vp_render: unknown mode %d", mode);
    VP_CUDA(cudaEventRecord(c->ev1, st));
    c->timed = true;
    return VP_OK;
}

int vp_resolve(vp_context* c, void* dst, const void* src, int size, float scale, float gamma, vp_stream stream)
{
    if (!c || !dst || !src) return fail(VP_ERR_INVALID, "vp_resolve: bad arguments");
    VP_CUDA(cudaSetDevice(c->device));
    VP_CUDA(launch_resolve((float4*)dst, (const float4*)src, size, scale, gamma, (cudaStream_t)stream));
    c->launches++;
    return VP_OK;
}

int vp_accumulate(vp_context* c, void* d_dst, const void* d_src, int size, vp_stream stream)
{
    if (!c || !d_dst || !d_src || size < 0) return fail(VP_ERR_INVALID, "vp_accumulate: bad arguments");
    VP_CUDA(cudaSetDevice(c->device));
    VP_CUDA(launch_accumulate((float4*)d_dst, (const float4*)d_src, size, (cudaStream_t)stream));
    c->launches++;
    return VP_OK;
}

int vp_get_half_tables(vp_context* c, unsigned short* h_maxmin, unsigned short* h_clear, float* h_clear_f32, int* present)
{
    if (!c || !present) return fail(VP_ERR_INVALID, "vp_get_half_tables: bad arguments");
    *present = c->bounds_half && c->sun_clear_half ? 1 : 0;
    VP_CUDA(cudaSetDevice(c->device));
    const size_t cells = (size_t)c->S.ncx * c->S.ncy * c->S.ncz;
    if (*present && h_maxmin) VP_CUDA(cudaMemcpy(h_maxmin, c->bounds_half, cells * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (*present && h_clear) VP_CUDA(cudaMemcpy(h_clear, c->sun_clear_half, cells * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    if (c->sun_clear && h_clear_f32) VP_CUDA(cudaMemcpy(h_clear_f32, c->sun_clear, cells * sizeof(float), cudaMemcpyDeviceToHost));
    return VP_OK;
}

int vp_render_to_host(vp_context* c, void* h_sum, int first_frame, int n_frames, int frame_stride, const vp_param* p, int mode)
{
    if (!c || !h_sum || !p) return fail(VP_ERR_INVALID, "vp_render_to_host: bad arguments");
    VP_CUDA(cudaSetDevice(c->device));
    const size_t bytes = (size_t)p->width * p->height * sizeof(float4);
    if (bytes > c->host_acc_bytes)
    {
        dev_free(c->host_acc);
        c->host_acc_bytes = 0;
        VP_CUDA(cudaMalloc(&c->host_acc, bytes));
        c->host_acc_bytes = bytes;
    }
    VP_CUDA(cudaMemcpyAsync(c->host_acc, h_sum, bytes, cudaMemcpyHostToDevice, 0));
    VP_TRY(vp_render(c, c->host_acc, first_frame, n_frames, frame_stride, p, mode, nullptr));
    VP_CUDA(cudaMemcpyAsync(h_sum, c->host_acc, bytes, cudaMemcpyDeviceToHost, 0));
    VP_CUDA(cudaStreamSynchronize(0));
    return VP_OK;
}

int vp_sync(vp_context* c)
{
    if (!c) return fail(VP_ERR_INVALID, "null context");
    VP_CUDA(cudaSetDevice(c->device));
    VP_CUDA(cudaDeviceSynchronize());
    return VP_OK;
}

// ---- introspection ------------------------------------------------------------------------------------
int vp_get_bounds_voxel(vp_context* c, float* h_out)
{
    if (!c || !c->bounds_voxel) return fail(VP_ERR_INVALID, "no per-voxel bounds");
    VP_CUDA(cudaMemcpy(h_out, c->bounds_voxel, (size_t)c->S.nx * c->S.ny * c->S.nz * sizeof(float2), cudaMemcpyDeviceToHost));
    return VP_OK;
}
int vp_get_bounds_cell(vp_context* c, float* h_out, int* dims3)
{
    const bool raw_jumps = dims3 && dims3[0] == -1;  // test hook: keep the encoded vacuum jumps
    if (!c || !c->bounds_cell) return fail(VP_ERR_INVALID, "no per-cell bounds");
    if (dims3) { dims3[0] = c->S.ncx; dims3[1] = c->S.ncy; dims3[2] = c->S.ncz; }
    if (h_out)
    {
        const size_t cells = (size_t)c->S.ncx * c->S.ncy * c->S.ncz;
        VP_CUDA(cudaMemcpy(h_out, c->bounds_cell, cells * sizeof(float2), cudaMemcpyDeviceToHost));
        if (!raw_jumps)
            for (size_t i = 0; i < cells; i++)
                if (h_out[2 * i] < 1e-20f) h_out[2 * i] = 0.0f;  // vacuum cells store -jump (fringe: 1e-30) in the max field
    }
    return VP_OK;
}
int vp_get_opacity(vp_context* c, float* h_out)
{
    if (!c || !c->opacity) return fail(VP_ERR_INVALID, "no opacity table");
    size_t N = (size_t)c->S.nx * c->S.ny * c->S.nz;
    float* d = nullptr;
    VP_CUDA(cudaMalloc(&d, N * sizeof(float)));
    VP_CUDA(launch_gather_opacity(c->S, d, 0));
    cudaError_t e = cudaMemcpy(h_out, d, N * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d);
    VP_CUDA(e);
    return VP_OK;
}
int vp_get_opacity_fast(vp_context* c, float* h_out)
{
    if (!c || !c->opacity_oct) return fail(VP_ERR_INVALID, "no production opacity table");
    size_t N = (size_t)c->S.nx * c->S.ny * c->S.nz;
    DevTmp d;
    VP_CUDA(d.alloc(N * sizeof(float)));
    VP_CUDA(launch_gather_opacity_oct(c->S, d.as<float>(), 0));
    VP_CUDA(cudaMemcpy(h_out, d.p, N * sizeof(float), cudaMemcpyDeviceToHost));
    return VP_OK;
}
int vp_opacity_build_ms(vp_context* c, float* ms)
{
    if (!c || !ms) return fail(VP_ERR_INVALID, "bad arguments");
    *ms = c->opacity_ms;
    return VP_OK;
}
int vp_fetch_density(vp_context* c, const float* h_pos3, int n, int parity_filter, float* h_out)
{
    if (!c || !c->have_volume || c->S.julia) return fail(VP_ERR_NO_VOLUME, "vp_fetch_density: no stored volume");
    float3* dp = nullptr;
    float*  dv = nullptr;
    VP_CUDA(cudaMalloc(&dp, (size_t)n * sizeof(float3)));
    VP_CUDA(cudaMalloc(&dv, (size_t)n * sizeof(float)));
    VP_CUDA(cudaMemcpy(dp, h_pos3, (size_t)n * sizeof(float3), cudaMemcpyHostToDevice));
    VP_CUDA(launch_fetch_density(c->S, dp, n, parity_filter, dv, 0));
    cudaError_t e = cudaMemcpy(h_out, dv, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(dp);
    cudaFree(dv);
    VP_CUDA(e);
    return VP_OK;
}
int vp_volume_stats(vp_context* c, unsigned long long* out8)
{
    if (!c || !out8) return fail(VP_ERR_INVALID, "bad arguments");
    out8[0] = c->n_bricks;
    out8[1] = c->n_slots;
    out8[2] = c->octet_bytes;
    out8[3] = (unsigned long long)c->bound_D;
    out8[4] = c->bounds_cell ? (unsigned long long)c->S.ncx * c->S.ncy * c->S.ncz * sizeof(float2) : 0;
    out8[5] = c->bounds_voxel ? (unsigned long long)c->S.nx * c->S.ny * c->S.nz * sizeof(float2) : 0;
    out8[6] = (c->opacity ? (unsigned long long)c->n_slots * kOpBrickPad * sizeof(float) : 0) +
              (c->opacity_oct ? (unsigned long long)c->n_slots * kBrickCells * 16 : 0);
    out8[7] = (unsigned long long)(1u << c->S.cell_log2);
    return VP_OK;
}
int vp_rng_sequence(vp_context* c, unsigned int x, unsigned int y, unsigned int frame, int n, float* h_out_f, unsigned int* h_out_u)
{
    if (!c || n < 1) return fail(VP_ERR_INVALID, "bad arguments");
    float*    df = nullptr;
    uint32_t* du = nullptr;
    VP_CUDA(cudaMalloc(&df, (size_t)n * 4));
    VP_CUDA(cudaMalloc(&du, (size_t)n * 4));
    VP_CUDA(launch_rng_sequence(x, y, frame, n, df, du, 0));
    cudaError_t e = cudaSuccess;
    if (h_out_f) e = cudaMemcpy(h_out_f, df, (size_t)n * 4, cudaMemcpyDeviceToHost);
    if (h_out_u && e == cudaSuccess) e = cudaMemcpy(h_out_u, du, (size_t)n * 4, cudaMemcpyDeviceToHost);
    cudaFree(df);
    cudaFree(du);
    VP_CUDA(e);
    return VP_OK;
}
int vp_philox2x32(vp_context* c, unsigned int c0, unsigned int c1, unsigned int key, unsigned int* h_out2)
{
    if (!c || !h_out2) return fail(VP_ERR_INVALID, "bad arguments");
    uint32_t* d = nullptr;
    VP_CUDA(cudaMalloc(&d, 8));
    VP_CUDA(launch_philox(c0, c1, key, d, 0));
    cudaError_t e = cudaMemcpy(h_out2, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    VP_CUDA(e);
    return VP_OK;
}
int vp_set_stats(vp_context* c, int enable)
{
    if (!c) return fail(VP_ERR_INVALID, "null context");
    c->stats_on = enable != 0;
    return VP_OK;
}
int vp_render_counters(vp_context* c, unsigned long long* out8, int reset)
{
    if (!c || !out8) return fail(VP_ERR_INVALID, "bad arguments");
    VP_CUDA(cudaDeviceSynchronize());
    VP_CUDA(cudaMemcpy(out8, c->d_stats, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (reset) VP_CUDA(cudaMemset(c->d_stats, 0, 16 * sizeof(unsigned long long)));
    return VP_OK;
}
int vp_last_kernel_ms(vp_context* c, float* ms)
{
    if (!c || !ms || !c->timed) return fail(VP_ERR_INVALID, "no timed render");
    VP_CUDA(cudaEventSynchronize(c->ev1));
    VP_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return VP_OK;
}
int vp_launch_count(vp_context* c, unsigned long long* n)
{
    if (!c || !n) return fail(VP_ERR_INVALID, "bad arguments");
    *n = c->launches;
    return VP_OK;
}

// plain device-memory helpers for C / ctypes callers that have no CUDA runtime of their own
void* vp_dev_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, bytes);
    return p;
}
int vp_dev_free(void* p) { return (int)cudaFree(p); }
int vp_dev_zero(void* p, size_t bytes) { return (int)cudaMemset(p, 0, bytes); }
int vp_dev_to_host(void* h, const void* d, size_t bytes) { return (int)cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost); }
int vp_host_to_dev(void* d, const void* h, size_t bytes) { return (int)cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice); }

// =======================================================================================================
// (3) combining the per-GPU float4 sums of a sample-sharded render: NCCL over NVLink / NVSwitch.
// libnccl.so.2 is bound at run time (dlopen), so the library itself has no link-time dependency on it and a process
// that already carries an NCCL (e.g. torch's) shares that copy.  The few NCCL types used are ABI-stable across 2.x
// (nccl.h: ncclUniqueId = 128 opaque bytes, ncclComm_t = opaque pointer, ncclFloat = 7, ncclSum = 0, ncclSuccess = 0).
// =======================================================================================================
namespace
{
struct NcclId { char internal[128]; };
struct NcclApi
{
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*)                                                               = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int)                                             = nullptr;
    int (*CommInitAll)(void**, int, const int*)                                               = nullptr;
    int (*CommDestroy)(void*)                                                                 = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t)             = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t)               = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t)                    = nullptr;
    int (*GroupStart)()                                                                       = nullptr;
    int (*GroupEnd)()                                                                         = nullptr;
    const char* (*GetErrorString)(int)                                                        = nullptr;
    int (*GetVersion)(int*)                                                                   = nullptr;
    bool ok = false;
};
NcclApi& nccl()
{
    static NcclApi    api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("VOLPATH_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names)
        {
            if (!n || !*n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
#define VP_SYM(field, name) *(void**)(&api.field) = dlsym(api.lib, name)
        VP_SYM(GetUniqueId, "ncclGetUniqueId");
        VP_SYM(CommInitRank, "ncclCommInitRank");
        VP_SYM(CommInitAll, "ncclCommInitAll");
        VP_SYM(CommDestroy, "ncclCommDestroy");
        VP_SYM(Reduce, "ncclReduce");
        VP_SYM(AllReduce, "ncclAllReduce");
        VP_SYM(AllGather, "ncclAllGather");
        VP_SYM(GroupStart, "ncclGroupStart");
        VP_SYM(GroupEnd, "ncclGroupEnd");
        VP_SYM(GetErrorString, "ncclGetErrorString");
        VP_SYM(GetVersion, "ncclGetVersion");
#undef VP_SYM
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.Reduce && api.GroupStart &&
                 api.GroupEnd && api.GetErrorString;
    });
    return api;
}
constexpr int kNcclFloat = 7, kNcclSum = 0;
#define VP_NCCL(x)                                                                                              \
    do {                                                                                                        \
        int r_ = (x);                                                                                           \
        if (r_ != 0) return fail(20000 + r_, "%s: %s", #x, nccl().GetErrorString ? nccl().GetErrorString(r_) : "?"); \
    } while (0)
}  // namespace

// in-place all-gather of `bytes_per_rank` bytes per rank into buf (rank r's share already sits at offset r * bytes_per_rank)
static int all_gather_bytes(vp_context* c, void* buf, size_t bytes_per_rank)
{
    if (!c->nccl_comm || !nccl().AllGather) return fail(VP_ERR_UNSUPPORTED, "all-gather needs vp_nccl_init (and an NCCL with ncclAllGather)");
    VP_NCCL(nccl().AllGather((const char*)buf + (size_t)c->nccl_rank * bytes_per_rank, buf, bytes_per_rank, /*ncclChar*/ 0, c->nccl_comm, 0));
    c->launches++;
    return VP_OK;
}

int vp_nccl_available(void)
{
    if (!nccl().ok) return 0;
    int v = 0;
    if (nccl().GetVersion) nccl().GetVersion(&v);
    return v > 0 ? v : 1;
}

int vp_nccl_unique_id(char* out128)
{
    if (!out128) return fail(VP_ERR_INVALID, "vp_nccl_unique_id: null out");
    if (!nccl().ok) return fail(VP_ERR_UNSUPPORTED, "vp_nccl_unique_id: libnccl.so.2 not found (set VOLPATH_NCCL_LIB)");
    NcclId id;
    VP_NCCL(nccl().GetUniqueId(&id));
    memcpy(out128, id.internal, sizeof(id.internal));
    return VP_OK;
}

int vp_nccl_init(vp_context* c, int n_ranks, int rank, const char* id128)
{
    if (!c || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(VP_ERR_INVALID, "vp_nccl_init: bad arguments");
    if (!nccl().ok) return fail(VP_ERR_UNSUPPORTED, "vp_nccl_init: libnccl.so.2 not found (set VOLPATH_NCCL_LIB)");
    VP_CUDA(cudaSetDevice(c->device));
    vp_nccl_destroy(c);
    NcclId id;
    memcpy(id.internal, id128, sizeof(id.internal));
    VP_NCCL(nccl().CommInitRank(&c->nccl_comm, n_ranks, id, rank));
    c->nccl_rank  = rank;
    c->nccl_ranks = n_ranks;
    return VP_OK;
}

int vp_nccl_destroy(vp_context* c)
{
    if (!c) return fail(VP_ERR_INVALID, "null context");
    if (c->nccl_comm && nccl().ok)
    {
        cudaSetDevice(c->device);
        nccl().CommDestroy(c->nccl_comm);
    }
    c->nccl_comm  = nullptr;
    c->nccl_rank  = -1;
    c->nccl_ranks = 0;
    return VP_OK;
}

int vp_reduce_nccl(vp_context* c, const void* d_send, void* d_recv, int size, int root, vp_stream stream)
{
    if (!c || !d_send || size < 0) return fail(VP_ERR_INVALID, "vp_reduce_nccl: bad arguments");
    if (!c->nccl_comm) return fail(VP_ERR_INVALID, "vp_reduce_nccl: vp_nccl_init has not been called on this context");
    if (root < 0 || root >= c->nccl_ranks) return fail(VP_ERR_INVALID, "vp_reduce_nccl: root %d out of range", root);
    if (c->nccl_rank == root && !d_recv) return fail(VP_ERR_INVALID, "vp_reduce_nccl: the root needs a receive buffer");
    VP_CUDA(cudaSetDevice(c->device));
    VP_NCCL(nccl().Reduce(d_send, d_recv, (size_t)size * 4, kNcclFloat, kNcclSum, root, c->nccl_comm, (cudaStream_t)stream));
    c->launches++;
    return VP_OK;
}

// single-process hosts: one context per GPU, all driven from this process; the sums are added into d_sums[root].
// NCCL (one communicator per device, grouped ncclReduce) when it can be loaded, else peer copies + the add kernel.
int vp_reduce(vp_context** ctxs, void** d_sums, int n, int size, int root)
{
    if (!ctxs || !d_sums || n < 1 || size < 0 || root < 0 || root >= n) return fail(VP_ERR_INVALID, "vp_reduce: bad arguments");
    for (int i = 0; i < n; i++)
        if (!ctxs[i] || !d_sums[i]) return fail(VP_ERR_INVALID, "vp_reduce: null context or buffer %d", i);
    if (n == 1) return VP_OK;
    for (int i = 0; i < n; i++)
    {
        VP_CUDA(cudaSetDevice(ctxs[i]->device));
        VP_CUDA(cudaDeviceSynchronize());
    }
    // one communicator set per device list, shared by all callers of this process
    static std::mutex           cache_lock;
    std::lock_guard<std::mutex> guard(cache_lock);
    static std::vector<int>     cached_devs;
    static std::vector<void*>   cached_comms;
    std::vector<int>          devs(n);
    for (int i = 0; i < n; i++) devs[i] = ctxs[i]->device;
    const bool use_nccl = nccl().ok && !env_flag("VOLPATH_REDUCE_P2P");
    if (use_nccl)
    {
        if (cached_devs != devs)
        {
            for (void* cm : cached_comms) nccl().CommDestroy(cm);
            cached_comms.assign(n, nullptr);
            cached_devs.clear();
            VP_NCCL(nccl().CommInitAll(cached_comms.data(), n, devs.data()));
            cached_devs = devs;
        }
        VP_NCCL(nccl().GroupStart());
        for (int i = 0; i < n; i++)
        {
            cudaSetDevice(devs[i]);
            int r = nccl().Reduce(d_sums[i], d_sums[root], (size_t)size * 4, kNcclFloat, kNcclSum, root, cached_comms[i], 0);
            if (r != 0)
            {
                nccl().GroupEnd();
                return fail(20000 + r, "ncclReduce: %s", nccl().GetErrorString(r));
            }
        }
        VP_NCCL(nccl().GroupEnd());
        for (int i = 0; i < n; i++)
        {
            VP_CUDA(cudaSetDevice(devs[i]));
            VP_CUDA(cudaDeviceSynchronize());
            ctxs[i]->launches++;
        }
        return VP_OK;
    }
    // no NCCL: stage each peer sum next to the root's and add it there
    VP_CUDA(cudaSetDevice(devs[root]));
    DevTmp stage;
    VP_CUDA(stage.alloc((size_t)size * sizeof(float4)));
    for (int i = 0; i < n; i++)
    {
        if (i == root) continue;
        VP_CUDA(cudaMemcpyPeer(stage.p, devs[root], d_sums[i], devs[i], (size_t)size * sizeof(float4)));
        VP_CUDA(launch_accumulate((float4*)d_sums[root], stage.as<float4>(), size, 0));
        ctxs[root]->launches++;
        VP_CUDA(cudaDeviceSynchronize());
    }
    return VP_OK;
}

// -------------------------------------------------------------------------------------------------------
// The same reduce over PEER MEMORY, without any communicator: on one node every GPU's accumulator can be mapped into
// the root's address space through CUDA IPC (64-byte handle, exchanged by the host's own means) and summed by ONE kernel
// that reads the peers over NVLink / NVSwitch (k_sum_peers: fixed rank order, so the result is bitwise reproducible).
// No bootstrap: ncclCommInitRank plus NCCL's ring set-up cost 4-10 s at 8 ranks (profiles/README.md), more than the
// whole 4K render takes on 8 GPUs; opening 7 handles costs milliseconds.
// -------------------------------------------------------------------------------------------------------
int vp_ipc_export(vp_context* c, const void* d_ptr, char* out64)
{
    if (!c || !d_ptr || !out64) return fail(VP_ERR_INVALID, "vp_ipc_export: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "VP_IPC_HANDLE_BYTES");
    VP_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    VP_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
    memcpy(out64, &h, sizeof(h));
    return VP_OK;
}

int vp_reduce_ipc(vp_context* c, void* d_sum, const char* peer_handles, int n_peers, int size, vp_stream stream)
{
    if (!c || !d_sum || (n_peers > 0 && !peer_handles) || n_peers < 0 || size < 0) return fail(VP_ERR_INVALID, "vp_reduce_ipc: bad arguments");
    VP_CUDA(cudaSetDevice(c->device));
    std::vector<const float4*> peers;
    for (int i = 0; i < n_peers; i++)
    {
        const std::string key(peer_handles + (size_t)i * 64, 64);
        void*             p = nullptr;
        for (auto& kv : c->ipc_open)
            if (kv.first == key) p = kv.second;
        if (!p)
        {
            cudaIpcMemHandle_t h;
            memcpy(&h, key.data(), sizeof(h));
            VP_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            c->ipc_open.emplace_back(key, p);
        }
        peers.push_back(static_cast<const float4*>(p));
    }
    if (n_peers > 0) VP_CUDA(launch_sum_peers((float4*)d_sum, peers.data(), n_peers, size, (cudaStream_t)stream));
    c->launches += (n_peers + 7) / 8;
    return VP_OK;
}

int vp_ipc_close(vp_context* c)
{
    if (!c) return fail(VP_ERR_INVALID, "null context");
    if (!c->ipc_open.empty())
    {
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();
        for (auto& kv : c->ipc_open) cudaIpcCloseMemHandle(kv.second);
        c->ipc_open.clear();
    }
    return VP_OK;
}

// =======================================================================================================
// (1) reference-named shims over one implicit context on the current device
// =======================================================================================================
static vp_context* g_shim      = nullptr;
// Default renderer of the render_kernel shim: the production megakernel (the same estimator in distribution; 3.5x the
// reference kernel when frames are batched or not synchronised one by one, profiles/README.md).  VP_MODE_PARITY -- fixed-seed
// traces equal to the reference kernel's, at 0.83x its speed (software texture emulation) -- by vp_shim_set_mode(0) or
// VOLPATH_SHIM_MODE=parity.
static int shim_default_mode()
{
    const char* m = getenv("VOLPATH_SHIM_MODE");
    if (m && (!strcmp(m, "parity") || !strcmp(m, "0"))) return VP_MODE_PARITY;
    if (m && (!strcmp(m, "wave") || !strcmp(m, "2"))) return VP_MODE_WAVE;
    return VP_MODE_FAST;
}
static int g_shim_mode = shim_default_mode();

vp_context* vp_shim_context(void)
{
    if (!g_shim)
    {
        int dev = 0;
        cudaGetDevice(&dev);
        if (vp_create(dev, &g_shim) != VP_OK)
        {
            fprintf(stderr, "volpath: %s\n", g_err);
            return nullptr;
        }
    }
    return g_shim;
}
void vp_shim_set_mode(int mode) { g_shim_mode = mode; }
// wait for every frame the render_kernel shim has in flight on its internal streams (see vp_context::ring)
int vp_shim_sync(void)
{
    vp_context* c = vp_shim_context();
    if (!c) return VP_ERR_NO_DEVICE;
    VP_CUDA(cudaSetDevice(c->device));
    for (cudaStream_t r : c->ring)
        if (r) VP_CUDA(cudaStreamSynchronize(r));
    return VP_OK;
}

#define SHIM_CTX()                         \
    vp_context* c = vp_shim_context();     \
    if (!c) return
#define SHIM_REPORT(call)                                              \
    do {                                                               \
        if ((call) != VP_OK) fprintf(stderr, "volpath: %s\n", g_err);  \
    } while (0)

void init_cuda(void* h_volume, vp_extent volumeSize, bool quantized, const vp_float3* boxmin, const vp_float3* boxmax)
{
    SHIM_CTX();
    const float* lo = (boxmin && boxmax) ? &boxmin->x : nullptr;
    const float* hi = (boxmin && boxmax) ? &boxmax->x : nullptr;
    // the filter mode survives re-initialisation like the reference's file-scope `linear_interp` (K.cu:351)
    SHIM_REPORT(vp_upload_volume(c, h_volume, (int)volumeSize.width, (int)volumeSize.height, (int)volumeSize.depth,
                                 quantized ? VP_VOXEL_U8 : VP_VOXEL_F32, quantized ? VP_VOXEL_U8 : VP_VOXEL_F32, VP_MEM_HOST, lo, hi,
                                 VP_BOUNDS_VOXEL | VP_BOUNDS_CELL));
}
void set_texture_filter_mode(bool bLinearFilter)
{
    SHIM_CTX();
    vp_set_filter(c, bLinearFilter ? 1 : 0);
}
void free_cuda_buffers(void)
{
    SHIM_CTX();
    vp_free_volume(c);
}
void precompute_opacity(const float* light_dir)
{
    SHIM_CTX();
    SHIM_REPORT(vp_precompute_opacity(c, light_dir));
}
void init_envmap(const vp_float4* HDRmap, int width, int height)
{
    SHIM_CTX();
    SHIM_REPORT(vp_set_envmap(c, &HDRmap->x, width, height));
}
void free_envmap(void) {}
void set_sun(float* sun_dir, float* sun_power)
{
    SHIM_CTX();
    SHIM_REPORT(vp_set_sun(c, sun_dir, sun_power));
}
void copy_inv_view_matrix(float* invViewMatrix, size_t sizeofMatrix)
{
    SHIM_CTX();
    float m[12];
    memcpy(m, c->S.inv_view, sizeof(m));
    memcpy(m, invViewMatrix, sizeofMatrix < sizeof(m) ? sizeofMatrix : sizeof(m));
    vp_set_inv_view(c, m);
}
void copy_inv_model_matrix(float* invModelMatrix, size_t sizeofMatrix)
{
    SHIM_CTX();
    memcpy(c->inv_model, invModelMatrix, sizeofMatrix < sizeof(c->inv_model) ? sizeofMatrix : sizeof(c->inv_model));
}
void init_rng(vp_dim3, vp_dim3, int, int) {}
void free_rng(void) {}
void scale(vp_float4* dst, vp_float4* src, int size, float scale_)
{
    SHIM_CTX();
    SHIM_REPORT(vp_resolve(c, dst, src, size, scale_, 0.0f, nullptr));
}
void gamma_correct(vp_float4* dst, vp_float4* src, int size, float scale_, float gamma)
{
    SHIM_CTX();
    SHIM_REPORT(vp_resolve(c, dst, src, size, scale_, gamma, nullptr));
}
void render_kernel(vp_dim3, vp_dim3, vp_float4* d_output, int spp, const vp_param* p)
{
    SHIM_CTX();
    SHIM_REPORT(render_shim(c, d_output, spp, p, g_shim_mode));
}

}  // extern "C"
