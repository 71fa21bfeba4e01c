// oracle/volpath_oracle.cpp -- TEST INFRASTRUCTURE.  NOT PRODUCT CODE.
//
// CPU restatement of the CUDA-volpath render hot path, written from an understanding of the
// reference (not copied): every function cites the reference file:line it follows
// (K.cu = src/volumeRender_kernel.cu, H.cpp = src/volumeRender.cpp of RNG65536/CUDA-volpath).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
// product (cuda-volpath_b200/) never does.
//
// PARITY PIN: the reference ships no tests, golden vectors or fixtures for this path (SURVEY.md
// section 4) -> "parity unpinned" by the reference's own tests.  The pin used instead is the
// reference itself, compiled from its sources by oracle/build_ref.py:
//   * libvolpath_ref_host.so  (the reference kernel source compiled by g++) -- tests/test_oracle_vs_ref.py
//     checks this restatement against it BIT FOR BIT on CPU (same libm, no FMA contraction on
//     either side, same texture emulation oracle/tex_emul.h);
//   * libvolpath_ref_cuda.so  (the reference kernel rebuilt for sm_100) -- the trace oracle on the
//     GPU box;
//   * golden vectors generated from those two are committed under tests/golden/.
//
// Build: g++ -O2 -fopenmp -ffp-contract=off -shared -fPIC (oracle/Makefile).
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tex_emul.h"

namespace
{
// ------------------------------------------------------------------------------------------
// small vector algebra with the reference's operation order (src/cuda/helper_math.h:1248-1312,
// 1420-1423: dot = x*x' + y*y' + z*z' left to right; normalize = v * rsqrtf(dot); host rsqrtf =
// 1/sqrtf, helper_math.h:62-65)
// ------------------------------------------------------------------------------------------
struct V3
{
    float x, y, z;
};
inline V3    v3(float a, float b, float c) { return V3{a, b, c}; }
inline V3    v3(float a) { return V3{a, a, a}; }
inline V3    operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3    operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3    operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline V3    operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3    operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3    operator/(V3 a, V3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline V3    operator/(V3 a, float b) { return v3(a.x / b, a.y / b, a.z / b); }  // helper_math.h:997-1000
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3    cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float host_rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline V3    normalize(V3 v) { return v * host_rsqrtf(dot(v, v)); }
inline float length(V3 v) { return sqrtf(dot(v, v)); }
// helper_math.h:42-50 host fminf/fmaxf are plain comparisons
inline float fmin_h(float a, float b) { return a < b ? a : b; }
inline float fmax_h(float a, float b) { return a > b ? a : b; }
inline V3    vmin(V3 a, V3 b) { return v3(fmin_h(a.x, b.x), fmin_h(a.y, b.y), fmin_h(a.z, b.z)); }
inline V3    vmax(V3 a, V3 b) { return v3(fmax_h(a.x, b.x), fmax_h(a.y, b.y), fmax_h(a.z, b.z)); }
inline float max_of(V3 v) { return fmax_h(fmax_h(v.x, v.y), v.z); }  // K.cu:67
inline float min_of(V3 v) { return fmin_h(fmin_h(v.x, v.y), v.z); }  // K.cu:71

// src/vecmath.h:9-16 (float constexprs, evaluated in float)
constexpr float kPi       = 3.1415926535897932384626422832795028841971f;
constexpr float kTwoPi    = kPi * 2.0f;
constexpr float kPi2      = kPi / 2.0f;
constexpr float k1Pi      = 1.0f / kPi;
constexpr float k1TwoPi   = 1.0f / kTwoPi;

// ------------------------------------------------------------------------------------------
// RNG (src/sampler.h:3-46)
// ------------------------------------------------------------------------------------------
inline uint32_t wang_hash(uint32_t seed)  // sampler.h:3-11
{
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed *= 9u;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    return seed;
}
struct RefRng
{
    uint32_t sx, sy;
    inline uint32_t next_u32()  // sampler.h:13-22 (xoroshiro64*)
    {
        uint32_t result = sx * 0x9e3779bbu;
        sy ^= sx;
        sx = ((sx << 26) | (sx >> 6)) ^ sy ^ (sy << 9);
        sy = (sx << 13) | (sx >> 19);
        return result;
    }
    inline void init(uint32_t px, uint32_t py, uint32_t frame)  // sampler.h:35-43
    {
        sx = wang_hash((px << 16) | py);
        sy = wang_hash(frame);
        next_u32();
    }
    inline float next()  // sampler.h:24-28: top 23 bits into the mantissa of [1,2), minus 1
    {
        uint32_t u = 0x3f800000u | (next_u32() >> 9);
        float    f;
        memcpy(&f, &u, 4);
        return (float)((double)f - 1.0);
    }
};

// Philox4x32-10 (Salmon et al., "Parallel Random Numbers: As Easy as 1, 2, 3", SC'11) -- the
// counter-based generator of the product's non-parity mode; restated here for known-answer tests.
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++)
    {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ------------------------------------------------------------------------------------------
// Procedural density of the no-OpenVDB build (K.cu:84-140)
// ------------------------------------------------------------------------------------------
inline float julia_density(V3 pos)
{
    const float radius = 1.4f;                       // K.cu:122
    const float cx = -0.2f, cy = 0.8f, cz = 0.0f, cw = 0.0f;  // K.cu:129
    const int   maxIter = 30;                        // K.cu:138
    float       qx = pos.x * radius, qy = pos.y * radius, qz = pos.z * radius, qw = 0.0f;  // K.cu:102
    int         iter = 0;
    float       d;
    do
    {
        // quaternion square, K.cu:90-98: r0 = x*x - dot(yzw,yzw); r_yzw = yzw * (x*2)
        float r0 = qx * qx - (qy * qy + qz * qz + qw * qw);
        float t  = qx * 2;
        float ry = qy * t, rz = qz * t, rw = qw * t;
        qx = r0 + cx; qy = ry + cy; qz = rz + cz; qw = rw + cw;  // K.cu:108
        d  = qx * qx + qy * qy + qz * qz + qw * qw;              // helper_math.h:1252 dot(float4)
    } while (d < 10.0f && iter++ < maxIter);                      // K.cu:109
    return (float)((double)iter > (double)maxIter * 0.9);         // K.cu:114
}

// ------------------------------------------------------------------------------------------
// Local (max,min) density bounds (H.cpp:1089-1267).  The reference computes, per voxel, the max
// and min over the cube of +-D voxels clamped to the grid, D = ceil(search_radius / (2/nx)), as
// three separable 1D sliding windows.  Max/min are exact (comparisons only) so any evaluation order
// gives identical bits; we use the separable form too (O(N*D)), and tests compare it with a
// brute-force cube on small grids and with the reference's own routine.
// ------------------------------------------------------------------------------------------
template <class T>
void window_minmax_axis(const T* src_max, const T* src_min, T* dst_max, T* dst_min, int nx, int ny, int nz, int axis,
                        int D)
{
    const int64_t sx = 1, sy = nx, sz = (int64_t)nx * ny;
    int           n  = axis == 0 ? nx : (axis == 1 ? ny : nz);
    int64_t       st = axis == 0 ? sx : (axis == 1 ? sy : sz);
    int           na = axis == 0 ? ny : nx;          // first orthogonal extent
    int           nb = axis == 2 ? ny : nz;          // second orthogonal extent
    int64_t       sa = axis == 0 ? sy : sx;
    int64_t       sb = axis == 2 ? sy : sz;
#pragma omp parallel for collapse(2)
    for (int b = 0; b < nb; b++)
        for (int a = 0; a < na; a++)
        {
            int64_t base = a * sa + b * sb;
            for (int i = 0; i < n; i++)
            {
                int lo = std::max(0, i - D), hi = std::min(n - 1, i + D);
                T   mx = src_max[base + lo * st], mn = src_min[base + lo * st];
                for (int j = lo + 1; j <= hi; j++)
                {
                    T a_ = src_max[base + j * st], b_ = src_min[base + j * st];
                    if (a_ > mx) mx = a_;
                    if (b_ < mn) mn = b_;
                }
                dst_max[base + i * st] = mx;
                dst_min[base + i * st] = mn;
            }
        }
}

inline int bound_radius_voxels(int nx, float search_radius)
{
    float cell_size = 2.0f / nx;                      // H.cpp:1098 (float / size_t -> float)
    return (int)ceil(search_radius / cell_size);      // H.cpp:1101
}

template <class T>
void compute_bounds(const T* vol, int nx, int ny, int nz, float search_radius, T* out_maxmin /* interleaved */)
{
    int            D = bound_radius_voxels(nx, search_radius);
    size_t         N = (size_t)nx * ny * nz;
    std::vector<T> amax(vol, vol + N), amin(vol, vol + N), bmax(N), bmin(N);
    window_minmax_axis(amax.data(), amin.data(), bmax.data(), bmin.data(), nx, ny, nz, 0, D);
    window_minmax_axis(bmax.data(), bmin.data(), amax.data(), amin.data(), nx, ny, nz, 1, D);
    window_minmax_axis(amax.data(), amin.data(), bmax.data(), bmin.data(), nx, ny, nz, 2, D);
    for (size_t i = 0; i < N; i++)
    {
        out_maxmin[2 * i]     = bmax[i];  // .x = max  (H.cpp:1141-1144)
        out_maxmin[2 * i + 1] = bmin[i];  // .y = min
    }
}

template <class T>
void compute_bounds_brute(const T* vol, int nx, int ny, int nz, int D, T* out_maxmin)
{
#pragma omp parallel for collapse(2)
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++)
            {
                T mx = vol[((size_t)k * ny + j) * nx + i], mn = mx;
                for (int kk = std::max(0, k - D); kk <= std::min(nz - 1, k + D); kk++)
                    for (int jj = std::max(0, j - D); jj <= std::min(ny - 1, j + D); jj++)
                        for (int ii = std::max(0, i - D); ii <= std::min(nx - 1, i + D); ii++)
                        {
                            T v = vol[((size_t)kk * ny + jj) * nx + ii];
                            if (v > mx) mx = v;
                            if (v < mn) mn = v;
                        }
                size_t o = ((size_t)k * ny + j) * nx + i;
                out_maxmin[2 * o]     = mx;
                out_maxmin[2 * o + 1] = mn;
            }
}

// ------------------------------------------------------------------------------------------
// Param (src/param.h:4-12): 44-byte POD, passed by value to the kernel
// ------------------------------------------------------------------------------------------
struct Param
{
    uint32_t width, height;
    float    density, brightness;
    float    albedo[3];
    float    g;
    float    sigma_t[3];
};
static_assert(sizeof(Param) == 44, "Param layout");

// ------------------------------------------------------------------------------------------
// Scene state = the reference's file-scope statics / __constant__ symbols (K.cu:337-352, 626,
// 858-880, 1254-1256) gathered in one context
// ------------------------------------------------------------------------------------------
struct Ctx
{
    bool           julia = false;       // no-OpenVDB build: procedural density, bounds (1,0) (K.cu:705-706,1622)
    int            nx = 0, ny = 0, nz = 0;
    bool           quantized = false;
    bool           linear    = false;   // K.cu:351 (host flips it to true, H.cpp:39,1344)
    V3             bmin{-1, -1, -1}, bmax{1, 1, 1}, l_inv{0.5f, 0.5f, 0.5f};
    texemu::Array  density;             // K.cu:384-404
    texemu::Array  bounds;              // K.cu:392-412  (.x max, .y min)
    texemu::Array  opacity;             // K.cu:540
    bool           have_opacity = false;
    texemu::Array  env;                 // K.cu:1098-1141
    // env-map importance sampling (PASSIVE_ENVMAP 0 variant, K.cu:21, 904-1034, 1144-1210)
    bool               passive_envmap = true;
    std::vector<float> env_cdf_y, env_cdf_x;   // EnvmapCdfY / EnvmapCdfX textures (point, unnormalised coords)
    float              env_pdfnorm_alt = 0.0f; // HDRpdfnormAlt
    float          inv_view[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};  // K.cu:626
    V3             sun_dir{0, 1, 0}, sun_power{0, 0, 0}, sun_power_original{0, 0, 0};  // K.cu:1254-1256

    texemu::Texture tex_density() const { return texemu::Texture{&density, linear, true}; }
    texemu::Texture tex_bounds() const { return texemu::Texture{&bounds, false, true}; }   // K.cu:395,412
    texemu::Texture tex_opacity() const { return texemu::Texture{&opacity, true, true}; }  // K.cu:541-542
    texemu::Texture tex_env() const { return texemu::Texture{&env, false, true}; }         // K.cu:1099-1100
};

// CudaTexture::sample_w (K.cu:173-178): p = (pos - min) * l_inv, normalized coordinates
inline float sample_density_raw(const Ctx& c, V3 pos)
{
    V3 p = (pos - c.bmin) * c.l_inv;
    return texemu::fetch3(c.tex_density(), p.x, p.y, p.z, 0);
}

// vol_sigma_t (K.cu:682-708)
inline float vol_sigma_t(const Ctx& c, V3 pos, float density)
{
    if (c.julia) return julia_density(pos * 1.0f) * density;  // c_world_to_normalized := 1
    float t = sample_density_raw(c, pos);
    t *= density;
    return t;
}

// vol_bound_minmax (K.cu:1610-1624)
inline void vol_bound_minmax(const Ctx& c, V3 pos, float& bmax_, float& bmin_)
{
    if (c.julia)
    {
        bmax_ = 1.0f;
        bmin_ = 0.0f;
        return;
    }
    V3 p  = (pos - c.bmin) * c.l_inv;
    bmax_ = texemu::fetch3(c.tex_bounds(), p.x, p.y, p.z, 0);
    bmin_ = texemu::fetch3(c.tex_bounds(), p.x, p.y, p.z, 1);
}

// intersectBox (K.cu:654-680); `clamp_near` selects the intersect_box variant (K.cu:453-481)
inline bool intersect_box(V3 o, V3 d, V3 bmin, V3 bmax, float& tnear, float& tfar, bool clamp_near)
{
    V3 invR = v3(1.0f) / d;
    V3 tbot = invR * (bmin - o);
    V3 ttop = invR * (bmax - o);
    V3 tmn  = vmin(ttop, tbot);
    V3 tmx  = vmax(ttop, tbot);
    float largest_tmin  = max_of(tmn);
    float smallest_tmax = min_of(tmx);
    tnear = largest_tmin;
    tfar  = smallest_tmax;
    if (clamp_near && tnear <= 0) tnear = 0;
    return smallest_tmax > largest_tmin && smallest_tmax >= 1e-3f;
}

constexpr float kSearchRadius = 0.05f;  // K.cu:151
constexpr int   kMaxDepth     = 800;    // K.cu:34

// intersectSuperVolume (K.cu:1626-1661)
inline bool intersect_super_volume(const Ctx& c, V3 o, V3 d, float& tnear, float& tfar, float& dmin, float& dmax)
{
    V3 invR = v3(1.0f) / d;
    V3 tbot = invR * (c.bmin - o);
    V3 ttop = invR * (c.bmax - o);
    V3 tmn  = vmin(ttop, tbot);
    V3 tmx  = vmax(ttop, tbot);
    float largest_tmin  = max_of(tmn);
    float smallest_tmax = min_of(tmx);
    tnear = fmax_h(largest_tmin, 0.0f);
    tfar  = fmin_h(smallest_tmax, kSearchRadius);
    float bx, by;
    vol_bound_minmax(c, o + d * tnear, bx, by);
    dmin = by;
    dmax = fmax_h(0.0001f, bx);
    return smallest_tmax > largest_tmin && smallest_tmax >= 1e-3f;
}

// Frame (K.cu:557-573)
struct Frame
{
    V3 n, t, b;
    explicit Frame(V3 normal)
    {
        n    = normal;
        V3 a = (double)fabsf(n.x) > 0.1 ? v3(0, 1, 0) : v3(1, 0, 0);
        t    = normalize(cross(a, n));
        b    = cross(n, t);
    }
    V3 to_world(V3 c) const { return t * c.x + b * c.y + n * c.z; }
};

// HGPhaseFunction (K.cu:575-619)
struct HG
{
    float g;
    V3    sample_local(float rnd0, float rnd1) const  // K.cu:580-598
    {
        float cos_theta;
        if (fabsf(g) > 1e-6f)
        {
            float s   = 2.0f * rnd0 - 1.0f;
            float f   = (1.0f - g * g) / (1.0f + g * s);
            cos_theta = (0.5f / g) * (1.0f + g * g - f * f);
            cos_theta = fmax_h(0.0f, fmin_h(1.0f, cos_theta));  // Q3: clamp to [0,1]
        }
        else
        {
            cos_theta = 2.0f * rnd0 - 1.0f;
        }
        float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
        float phi       = 2.0f * kPi * rnd1;
        return v3(cosf(phi) * sin_theta, sinf(phi) * sin_theta, cos_theta);
    }
    float evaluate(float cos_theta) const  // K.cu:600-603
    {
        return (1.0f - g * g) / (4.0f * kPi * powf(1.0f + g * g - 2 * g * cos_theta, 1.5f));
    }
};

// Envmap::dir_to_uv / eval_envmap (K.cu:882-895, 956-973)
inline V3 eval_envmap(const Ctx& c, V3 dir)
{
    float phi   = acosf(dir.y);
    float theta = atanf(dir.z / dir.x) + kPi2;
    if (dir.x < 0) theta += kPi;
    float u = theta * k1TwoPi;
    float v = phi * k1Pi;
    texemu::Texture t = c.tex_env();
    return v3(texemu::fetch2(t, u, v, 0), texemu::fetch2(t, u, v, 1), texemu::fetch2(t, u, v, 2));
}

// background (K.cu:1258-1267)
inline V3 background(const Ctx& c, V3 dir, int depth)
{
    if (depth == 0 && (dot(dir, c.sun_dir) > 94.0f / sqrtf(94.0f * 94.0f + 0.45f * 0.45f))) return c.sun_power_original;
    return eval_envmap(c, dir);
}

// luminance (K.cu:946-954): float * double literals, summed in double, returned as float
inline float luminance(V3 c) { return (float)(c.x * 0.2126 + c.y * 0.7152 + c.z * 0.0722); }

// build_cdf_1d / build_cdf_2d + the host part of init_envmap (K.cu:1036-1070, 1144-1210), PRE_WARP 1, MULT_PDF 0
inline float build_cdf_1d(const float* f, float* pdf, float* cdf, int size)
{
    if (size < 1) return 0;
    float sum = 0.0f;
    for (int i = 0; i < size; i++) sum += f[i];
    float norm = 1.0f / sum;
    float I    = 0.0f;
    for (int i = 0; i < size; i++)
    {
        float p = f[i] * norm;
        I += p;
        pdf[i] = p;
        cdf[i] = I;
    }
    cdf[size - 1] = 1.0f;
    return sum;
}
inline void build_env_cdf(Ctx& c)
{
    const int w = c.env.w, h = c.env.h;
    if (w < 1 || h < 1) return;
    size_t             total = (size_t)w * h;
    std::vector<float> lum(total);
    for (size_t i = 0; i < total; i++)
    {
        float px[4];
        memcpy(px, &c.env.bytes[i * 16], 16);
        lum[i] = (float)(px[0] * 0.2126 + px[1] * 0.7152 + px[2] * 0.0722);  // luminance(float4)
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
    c.env_pdfnorm_alt     = (float)w * (float)h * k1TwoPiPi / lumsum;
    std::vector<float> pdfY(h), pdfX(total), row_sum(h);
    c.env_cdf_y.assign(h, 0.0f);
    c.env_cdf_x.assign(total, 0.0f);
    for (int y = 0; y < h; y++)
        row_sum[y] = build_cdf_1d(lum.data() + (size_t)y * w, pdfX.data() + (size_t)y * w, c.env_cdf_x.data() + (size_t)y * w, w);
    build_cdf_1d(row_sum.data(), pdfY.data(), c.env_cdf_y.data(), h);
}

// sample_y / sample_x (K.cu:904-944): lower-bound binary searches in the CDF rows
inline int env_sample_y(const Ctx& c, float r)
{
    int begin = 0, end = c.env.h - 1;
    while (end > begin)
    {
        int mid = begin + (end - begin) / 2;
        if (c.env_cdf_y[mid] >= r) end = mid; else begin = mid + 1;
    }
    return begin;
}
inline int env_sample_x(const Ctx& c, int y, float r)
{
    int begin = 0, end = c.env.w - 1;
    while (end > begin)
    {
        int mid = begin + (end - begin) / 2;
        if (c.env_cdf_x[(size_t)y * c.env.w + mid] >= r) end = mid; else begin = mid + 1;
    }
    return begin;
}
// sample_envmap (K.cu:979-1009): u, v in: random numbers; out: texel-centre coordinates; returns the pdf
inline float sample_envmap(const Ctx& c, float& u, float& v, V3& col)
{
    int iy = env_sample_y(c, v);
    int ix = env_sample_x(c, iy, u);
    u      = ((float)ix + 0.5f) / (float)c.env.w;
    v      = ((float)iy + 0.5f) / (float)c.env.h;
    texemu::Texture t = c.tex_env();
    col    = v3(texemu::fetch2(t, u, v, 0), texemu::fetch2(t, u, v, 1), texemu::fetch2(t, u, v, 2));
    return luminance(col) * c.env_pdfnorm_alt;
}
inline float pdf_envmap(const Ctx& c, V3 col) { return luminance(col) * c.env_pdfnorm_alt; }  // K.cu:1011-1034
inline V3 uv_to_dir(float u, float v)  // K.cu:897-902
{
    float theta = u * kTwoPi;
    float phi   = v * kPi;
    return v3(sinf(phi) * sinf(theta), cosf(phi), sinf(phi) * -cosf(theta));
}
inline float mis_balance(float a, float b) { return a / (a + b); }  // K.cu:54

// Tr_spectral (K.cu:754-808)
inline V3 tr_spectral(const Ctx& c, V3 start, V3 end, float inv_sigma, float density, V3 sigma_t, RefRng& rng,
                      uint64_t* n_fetch)
{
    V3    o = start;
    V3    d = normalize(end - start);
    float t_near, t_far;
    if (!intersect_box(o, d, c.bmin, c.bmax, t_near, t_far, false)) return v3(1.0f);
    if (t_near < 0.0f) t_near = 0.0f;
    float max_t = fmin_h(t_far, length(start - end));
    float dist  = t_near;
    int   xterm = 0, yterm = 0, zterm = 0;
    for (;;)
    {
        dist += -logf(rng.next()) * inv_sigma;
        if (dist >= max_t || (xterm && yterm && zterm)) break;  // Q8: the test comes after the draw
        V3    pos = o + d * dist;
        float e   = rng.next();
        float den = vol_sigma_t(c, pos, density);
        if (n_fetch) ++*n_fetch;
        if (!xterm && e < sigma_t.x * den * inv_sigma) xterm = 1;
        if (!yterm && e < sigma_t.y * den * inv_sigma) yterm = 1;
        if (!zterm && e < sigma_t.z * den * inv_sigma) zterm = 1;
    }
    return v3((float)(1 - xterm), (float)(1 - yterm), (float)(1 - zterm));
}

struct PathStats
{
    uint64_t track_fetch = 0, shadow_fetch = 0, segments = 0, opacity_fetch = 0, env_eval = 0, scatters = 0;
};

// One path-sample = one thread of __d_render_bounded_decomp (K.cu:1958-2318)
inline void trace_path(const Ctx& c, const Param& P, uint32_t x, uint32_t y, int spp, float out4[4], PathStats* st)
{
    const float density    = P.density;
    const float brightness = P.brightness;
    RefRng      rng;
    rng.init(x, y, (uint32_t)spp);  // K.cu:1972-1973

    float u = (x * 2.0f - P.width) / P.width;    // K.cu:1977
    float v = (y * 2.0f - P.height) / P.width;   // K.cu:1978 (Q7: divided by width)
    float fovx = 54.43;                          // K.cu:1981
    const float* M = c.inv_view;
    // mul(M, float4(0,0,0,1)) (K.cu:641-649): dot(v, row) in x,y,z,w order
    V3 o = v3(0.0f * M[0] + 0.0f * M[1] + 0.0f * M[2] + 1.0f * M[3], 0.0f * M[4] + 0.0f * M[5] + 0.0f * M[6] + 1.0f * M[7],
              0.0f * M[8] + 0.0f * M[9] + 0.0f * M[10] + 1.0f * M[11]);
    V3 dc = v3(u, v, (float)(-1.0f / tan((double)fovx * 0.00872664626)));  // K.cu:1985 (double tan)
    V3 d  = normalize(v3(dot(dc, v3(M[0], M[1], M[2])), dot(dc, v3(M[4], M[5], M[6])), dot(dc, v3(M[8], M[9], M[10]))));

    V3 radiance   = v3(0.0f);
    V3 throughput = v3(1.0f);

    V3    sigma_t_spectral = v3(P.sigma_t[0], P.sigma_t[1], P.sigma_t[2]);                       // K.cu:1996
    V3    sigma_s_spectral = sigma_t_spectral * v3(P.albedo[0], P.albedo[1], P.albedo[2]);      // K.cu:1997
    float max_sigma_t      = max_of(sigma_t_spectral);
    float min_sigma_t      = min_of(sigma_t_spectral);

    float sigma_c_prime = 0, distc = 0, sigma_r_prime = 0, inv_sigma = 0, inv_sigma_t = 0;
    V3    sigma_c_spectral = v3(0.0f);
    int   num_scatters     = 0;

    while (num_scatters < kMaxDepth)
    {
        float t_near, t_far, d_min, d_max;
        bool  hit = intersect_super_volume(c, o, d, t_near, t_far, d_min, d_max);  // K.cu:2020
        if (st) st->segments++;
        bool use_decomposition = d_min > 0.0f;                                      // K.cu:2021
        if (!hit)
        {
            // K.cu:2026-2030: passive env map adds the environment at any depth, the MIS variant only at depth 0
            if (c.passive_envmap || num_scatters == 0) radiance = radiance + background(c, d, num_scatters) * throughput;
            if (st) st->env_eval++;
            break;
        }
        V3    pos  = o + d * t_near;
        float dist = t_near;

        // "hyperion trick" (K.cu:2039-2044): reduced scattering coefficients after 5 bounces
        float s = fmax_h(0.0f, fmin_h(1.0f, (num_scatters - 5) * 0.066666666666666666667f));
        float g = (1 - s) * P.g;
        float reduction_factor = (1 - s) + s * (1 - P.g);
        float density_prime    = reduction_factor * density;
        float sigma_t_prime    = max_sigma_t * density_prime * d_max;

        if (use_decomposition)  // K.cu:2048-2059
        {
            sigma_c_prime    = min_sigma_t * density_prime * d_min;
            distc            = dist - logf(rng.next()) / fmax_h(sigma_c_prime, 1e-20f);
            sigma_r_prime    = fmax_h(sigma_t_prime - sigma_c_prime, 1e-20f);
            sigma_c_spectral = v3(sigma_c_prime);
        }
        else
        {
            distc            = 1e20f;
            sigma_c_spectral = v3(0);
        }
        HG phase{g};  // Q4: g of the pre-increment scatter count
        inv_sigma_t = 1.0f / sigma_t_prime;                                         // K.cu:2067
        inv_sigma   = use_decomposition ? 1.0f / sigma_r_prime : inv_sigma_t;       // K.cu:2068-2075

        for (;;)  // K.cu:2082-2142
        {
            dist += -logf(rng.next()) * inv_sigma;
            if (dist >= distc || dist >= t_far)
            {
                pos = o + d * distc;
                break;
            }
            else
            {
                pos = o + d * dist;
            }
            float den = vol_sigma_t(c, pos, density_prime);
            if (st) st->track_fetch++;
            V3 sigma_t_den    = sigma_t_spectral * den - sigma_c_spectral;
            V3 sigma_s_den    = sigma_s_spectral * den - sigma_c_spectral;
            V3 sigma_null_den = v3(sigma_t_prime) - sigma_t_den;
            float Ps = fabsf(sigma_t_den.x * throughput.x) + fabsf(sigma_t_den.y * throughput.y) +
                       fabsf(sigma_t_den.z * throughput.z);
            float Pn = fabsf(sigma_null_den.x * throughput.x) + fabsf(sigma_null_den.y * throughput.y) +
                       fabsf(sigma_null_den.z * throughput.z);
            float cc = (Ps + Pn);
            float e  = rng.next() * cc;
            if (e < Ps)
            {
                throughput = throughput * (sigma_s_den * (inv_sigma_t * cc / (Ps)));
                break;
            }
            else
            {
                throughput = throughput * (sigma_null_den * (inv_sigma_t * cc / Pn));
            }
        }

        bool through = fmin_h(distc, dist) >= t_far;  // K.cu:2145
        num_scatters += (!through);
        if (through)
        {
            o = o + d * t_far;  // K.cu:2153 tracking restart
            continue;
        }
        if (st) st->scatters++;

        Frame frame(d);
        {
            // K.cu:2168-2178 (post-increment scatter count; the inner g is dead, Q4)
            float s2 = fmax_h(0.0f, fmin_h(1.0f, (num_scatters - 5) * 0.066666666666666666667f));
            float reduction_factor2 = (1 - s2) + s2 * (1 - P.g);
            float density_prime2    = reduction_factor2 * density;
            float sigma_t_prime2    = max_sigma_t * density_prime2 * d_max;
            float inv_sigma2        = 1.0f / sigma_t_prime2;
            float ph                = phase.evaluate(dot(frame.n, c.sun_dir));
            V3    a;
            if (spp > 10 && num_scatters > 20)  // K.cu:2183
            {
                float tau;
                if (c.have_opacity)
                {
                    V3 p = (pos - c.bmin) * c.l_inv;
                    tau  = texemu::fetch3(c.tex_opacity(), p.x, p.y, p.z, 0);
                }
                else
                {
                    tau = 0.0f;  // table built from an absent (zero) density: see DESIGN.md, config C1
                }
                if (st) st->opacity_fetch++;
                V3 e3 = (-sigma_t_spectral) * density_prime2 * tau;
                a     = v3(expf(e3.x), expf(e3.y), expf(e3.z));
            }
            else
            {
                a = tr_spectral(c, pos, c.sun_dir * 1e10f, inv_sigma2, density_prime2, sigma_t_spectral, rng,
                                st ? &st->shadow_fetch : nullptr);
            }
            radiance = radiance + c.sun_power * (throughput * ph * a);  // K.cu:2188-2189, 2209-2210

            if (!c.passive_envmap)
            {
                // one-sample MIS between phase-function and env-map sampling (K.cu:2220-2297)
                const float P_phase = 0.5f, P_envmap = 1.0f - P_phase;
                if (rng.next() < P_phase)
                {
                    float u = rng.next();
                    float v = rng.next();
                    V3    brdf_dir = frame.to_world(phase.sample_local(u, v));
                    V3    envc     = eval_envmap(c, brdf_dir);
                    float pdf_brdf = phase.evaluate(dot(frame.n, brdf_dir));
                    float pdf_env_virtual = pdf_envmap(c, envc);
                    float weight   = mis_balance(pdf_brdf * P_phase, pdf_env_virtual * P_envmap) / P_phase;
                    V3    a2 = tr_spectral(c, pos, brdf_dir * 1e10f, inv_sigma2, density_prime2, sigma_t_spectral, rng,
                                           st ? &st->shadow_fetch : nullptr);
                    radiance = radiance + envc * (throughput * weight * a2);
                }
                else
                {
                    float u = rng.next();
                    float v = rng.next();
                    V3    envc;
                    float pdf_env = sample_envmap(c, u, v, envc);
                    if (pdf_env <= 0.0f) continue;  // K.cu:2266: back to the while loop WITHOUT a new direction (quirk)
                    V3    envmap_dir = uv_to_dir(u, v);
                    float pdf_brdf_virtual = phase.evaluate(dot(frame.n, envmap_dir));
                    float weight = mis_balance(pdf_env * P_envmap, pdf_brdf_virtual * P_phase) / P_envmap;
                    V3    a2 = tr_spectral(c, pos, envmap_dir * 1e10f, inv_sigma2, density_prime2, sigma_t_spectral, rng,
                                           st ? &st->shadow_fetch : nullptr);
                    radiance = radiance + envc * (throughput * phase.evaluate(dot(frame.n, envmap_dir)) / pdf_env * weight * a2);
                }
            }
        }
        // K.cu:2301 -- device order: first draw -> rnd0 (cos theta), second -> rnd1 (phi)  (Q6)
        float r0 = rng.next();
        float r1 = rng.next();
        V3    nd = normalize(frame.to_world(phase.sample_local(r0, r1)));
        o        = pos;
        d        = nd;
    }
    radiance = radiance * brightness;
    out4[0]  = fmax_h(radiance.x, 0.0f);  // K.cu:2315-2316 (Q9)
    out4[1]  = fmax_h(radiance.y, 0.0f);
    out4[2]  = fmax_h(radiance.z, 0.0f);
    out4[3]  = (float)num_scatters;       // K.cu:2309 (Q10)
}

// ------------------------------------------------------------------------------------------
// Synthetic fBm cloud (this repo's own deterministic input generator, SURVEY.md 8d "C2"): value
// noise on an integer-hashed lattice, all lattice arithmetic in integers, the few float operations
// are single IEEE operations (no contraction), so the CUDA generator reproduces it bit for bit.
// ------------------------------------------------------------------------------------------
inline uint32_t lattice_hash(uint32_t x, uint32_t y, uint32_t z, uint32_t seed)
{
    uint32_t h = seed;
    h ^= x * 0x8da6b343u; h = (h << 13) | (h >> 19); h *= 0x9e3779b1u;
    h ^= y * 0xd8163841u; h = (h << 13) | (h >> 19); h *= 0x9e3779b1u;
    h ^= z * 0xcb1ab31fu; h = (h << 13) | (h >> 19); h *= 0x9e3779b1u;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

// 16.16 fixed-point value noise, returns [0, 65535]
inline uint32_t value_noise_fx(uint32_t px, uint32_t py, uint32_t pz, uint32_t seed)
{
    uint32_t ix = px >> 16, iy = py >> 16, iz = pz >> 16;
    uint32_t fx = px & 0xffffu, fy = py & 0xffffu, fz = pz & 0xffffu;
    // smoothstep 3t^2 - 2t^3 in 16-bit fixed point
    auto sm = [](uint32_t t) -> uint32_t {
        uint64_t t2 = ((uint64_t)t * t) >> 16;
        uint64_t r  = (t2 * (3u * 65536u - 2u * t)) >> 16;
        return (uint32_t)(r > 65535u ? 65535u : r);
    };
    uint32_t wx = sm(fx), wy = sm(fy), wz = sm(fz);
    auto     L  = [&](uint32_t dx, uint32_t dy, uint32_t dz) -> uint64_t {
        return lattice_hash(ix + dx, iy + dy, iz + dz, seed) & 0xffffu;
    };
    auto lerp = [](uint64_t a, uint64_t b, uint32_t w) -> uint64_t { return (a * (65536u - w) + b * w) >> 16; };
    uint64_t x00 = lerp(L(0, 0, 0), L(1, 0, 0), wx), x10 = lerp(L(0, 1, 0), L(1, 1, 0), wx);
    uint64_t x01 = lerp(L(0, 0, 1), L(1, 0, 1), wx), x11 = lerp(L(0, 1, 1), L(1, 1, 1), wx);
    uint64_t y0 = lerp(x00, x10, wy), y1 = lerp(x01, x11, wy);
    return (uint32_t)lerp(y0, y1, wz);
}

// density in [0,1] at voxel (i,j,k) of an nx*ny*nz grid
inline float fbm_cloud_voxel(int i, int j, int k, int nx, int ny, int nz, uint32_t seed)
{
    // position in 16.16 lattice units: base frequency 3 cells across the LONGEST axis
    int      nmax = std::max(nx, std::max(ny, nz));
    uint64_t sum  = 0;
    for (int o = 0; o < 5; o++)
    {
        uint64_t f  = (uint64_t)3 << o;
        uint32_t px = (uint32_t)((((uint64_t)(2 * i + 1) * f) << 15) / (uint64_t)nmax);
        uint32_t py = (uint32_t)((((uint64_t)(2 * j + 1) * f) << 15) / (uint64_t)nmax);
        uint32_t pz = (uint32_t)((((uint64_t)(2 * k + 1) * f) << 15) / (uint64_t)nmax);
        sum += (uint64_t)value_noise_fx(px, py, pz, seed + 1234u + (uint32_t)o) << (4 - o);  // gain 0.5
    }
    // sum in [0, 65535*31]; normalise to [0,1] with one float division
    float n = (float)sum / (float)(65535u * 31u);
    // ellipsoidal falloff + flat base, in normalised box coordinates [-1,1]^3
    float ux = ((float)(2 * i + 1) / (float)nx) - 1.0f;
    float uy = ((float)(2 * j + 1) / (float)ny) - 1.0f;
    float uz = ((float)(2 * k + 1) / (float)nz) - 1.0f;
    float r2 = ux * ux;
    r2       = r2 + uy * uy;
    r2       = r2 + uz * uz;
    float fall = 1.0f - r2;                    // 1 at the centre, 0 on the unit sphere
    float base = (uy + 0.75f) * 4.0f;          // flat cloud base near y = -0.75
    base       = base < 0.0f ? 0.0f : (base > 1.0f ? 1.0f : base);
    float v    = n + fall * 0.7f;
    v          = v - 0.76f;
    v          = v * 3.0f;
    v          = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    return v * base;
}
}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

uint32_t vo_hash(uint32_t s) { return wang_hash(s); }

// the first n floats (and raw u32 if wanted) of the stream of pixel (x,y), frame f
void vo_rng_sequence(uint32_t x, uint32_t y, uint32_t frame, int n, float* out_f, uint32_t* out_u)
{
    RefRng r;
    r.init(x, y, frame);
    for (int i = 0; i < n; i++)
    {
        RefRng   c = r;
        uint32_t u = c.next_u32();
        if (out_u) out_u[i] = u;
        float f = r.next();
        if (out_f) out_f[i] = f;
    }
}

void vo_philox4x32_10(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) { philox4x32_10(ctr4, key2, out4); }

int vo_bound_radius_voxels(int nx, float radius) { return bound_radius_voxels(nx, radius); }

void vo_bounds_u8(const uint8_t* vol, int nx, int ny, int nz, float radius, uint8_t* out2)
{
    compute_bounds<uint8_t>(vol, nx, ny, nz, radius, out2);
}
void vo_bounds_f32(const float* vol, int nx, int ny, int nz, float radius, float* out2)
{
    compute_bounds<float>(vol, nx, ny, nz, radius, out2);
}
void vo_bounds_brute_u8(const uint8_t* vol, int nx, int ny, int nz, int D, uint8_t* out2)
{
    compute_bounds_brute<uint8_t>(vol, nx, ny, nz, D, out2);
}
void vo_bounds_brute_f32(const float* vol, int nx, int ny, int nz, int D, float* out2)
{
    compute_bounds_brute<float>(vol, nx, ny, nz, D, out2);
}

float vo_julia_density(float x, float y, float z) { return julia_density(v3(x, y, z)); }

void vo_fbm_cloud_f32(int nx, int ny, int nz, uint32_t seed, float* out)
{
#pragma omp parallel for collapse(2)
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) out[((size_t)k * ny + j) * nx + i] = fbm_cloud_voxel(i, j, k, nx, ny, nz, seed);
}

void* vo_create() { return new Ctx(); }
void  vo_destroy(void* h) { delete (Ctx*)h; }

// init_cuda (K.cu:354-420): density array + CPU bound volume; null box -> +-(1, ny/nx, nz/nx)
int vo_set_volume(void* h, const void* vol, int nx, int ny, int nz, int quantized, const float* bmin, const float* bmax)
{
    Ctx& c = *(Ctx*)h;
    if (!vol) return 1;  // K.cu:360-364 (the reference exits; the oracle reports)
    c.julia     = false;
    c.nx        = nx; c.ny = ny; c.nz = nz;
    c.quantized = quantized != 0;
    if (bmin && bmax)
    {
        c.bmin = v3(bmin[0], bmin[1], bmin[2]);
        c.bmax = v3(bmax[0], bmax[1], bmax[2]);
    }
    else
    {
        c.bmin = v3(-1.0f, -(float)ny / (float)nx, -(float)nz / (float)nx);  // K.cu:373-378
        c.bmax = v3(1.0f, (float)ny / (float)nx, (float)nz / (float)nx);
    }
    c.l_inv = v3(1.0f) / (c.bmax - c.bmin);  // K.cu:313
    size_t N = (size_t)nx * ny * nz;
    if (c.quantized)
    {
        c.density.alloc(nx, ny, nz, texemu::FMT_U8);
        memcpy(c.density.bytes.data(), vol, N);
        c.bounds.alloc(nx, ny, nz, texemu::FMT_U8x2);
        compute_bounds<uint8_t>((const uint8_t*)vol, nx, ny, nz, kSearchRadius, c.bounds.bytes.data());
    }
    else
    {
        c.density.alloc(nx, ny, nz, texemu::FMT_F32);
        memcpy(c.density.bytes.data(), vol, N * 4);
        c.bounds.alloc(nx, ny, nz, texemu::FMT_F32x2);
        compute_bounds<float>((const float*)vol, nx, ny, nz, kSearchRadius, (float*)c.bounds.bytes.data());
    }
    c.have_opacity = false;
    return 0;
}

// config C1: the reference's no-OpenVDB build with the assumptions of SURVEY.md 8c
int vo_set_julia(void* h)
{
    Ctx& c  = *(Ctx*)h;
    c.julia = true;
    c.bmin  = v3(-1.0f, -1.0f, -1.0f);
    c.bmax  = v3(1.0f, 1.0f, 1.0f);
    c.l_inv = v3(1.0f) / (c.bmax - c.bmin);
    c.have_opacity = false;
    return 0;
}

int vo_set_filter(void* h, int linear)
{
    ((Ctx*)h)->linear = linear != 0;
    return 0;
}

int vo_set_envmap(void* h, const float* rgba, int w, int hh)  // init_envmap (K.cu:1072-1141)
{
    Ctx& c = *(Ctx*)h;
    c.env.alloc(w, hh, 1, texemu::FMT_F32x4);
    memcpy(c.env.bytes.data(), rgba, (size_t)w * hh * 16);
    if (!c.passive_envmap) build_env_cdf(c);
    return 0;
}

// PASSIVE_ENVMAP switch (K.cu:21): 0 builds the CDF tables like init_envmap does and enables the MIS block
int vo_set_env_sampling(void* h, int enable)
{
    Ctx& c           = *(Ctx*)h;
    c.passive_envmap = enable == 0;
    if (!c.passive_envmap) build_env_cdf(c);
    return 0;
}

int vo_set_sun(void* h, const float* dir3, const float* power3)  // set_sun (K.cu:1269-1283)
{
    Ctx& c               = *(Ctx*)h;
    c.sun_power_original = v3(power3[0], power3[1], power3[2]);
    float r              = (float)(0.45 / 94.0f);
    c.sun_power          = c.sun_power_original * (kPi * (r * r));
    c.sun_dir            = v3(dir3[0], dir3[1], dir3[2]);
    return 0;
}

int vo_set_inv_view(void* h, const float* m12)  // copy_inv_view_matrix (K.cu:2320-2323)
{
    memcpy(((Ctx*)h)->inv_view, m12, 48);
    return 0;
}

// precompute_opacity + _precompute_opacity (K.cu:483-553)
int vo_precompute_opacity(void* h, const float* dir3)
{
    Ctx& c = *(Ctx*)h;
    if (c.julia) { c.have_opacity = false; return 0; }
    c.opacity.alloc(c.nx, c.ny, c.nz, texemu::FMT_F32);
    float* out = (float*)c.opacity.bytes.data();
    V3     ld  = v3(dir3[0], dir3[1], dir3[2]);
    const float dt = 0.001f;
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int k = 0; k < c.nz; k++)
        for (int j = 0; j < c.ny; j++)
            for (int i = 0; i < c.nx; i++)
            {
                V3 start0 = v3((i + 0.5f) / c.nx, (j + 0.5f) / c.ny, (k + 0.5f) / c.nz);  // K.cu:164-167
                V3 start  = start0 * (c.bmax - c.bmin) + c.bmin;                           // K.cu:171
                float tn, tf;
                bool  hit     = intersect_box(start, ld, c.bmin, c.bmax, tn, tf, true);
                float opacity = 0.0f;
                if (hit)
                {
                    for (float t = tn; t < tf; t += dt) opacity += sample_density_raw(c, start + ld * t);
                    opacity *= dt;
                }
                out[((size_t)k * c.ny + j) * c.nx + i] = opacity;
            }
    c.have_opacity = true;
    return 0;
}

int vo_get_bounds(void* h, void* out)
{
    Ctx& c = *(Ctx*)h;
    memcpy(out, c.bounds.bytes.data(), c.bounds.bytes.size());
    return 0;
}
int vo_get_opacity(void* h, float* out)
{
    Ctx& c = *(Ctx*)h;
    memcpy(out, c.opacity.bytes.data(), c.opacity.bytes.size());
    return 0;
}

// raw filtered density at world positions (CudaTexture::sample_w, K.cu:173-178) -- fetch-level tests
int vo_sample_density(void* h, const float* pos3, int n, float* out)
{
    const Ctx& c = *(Ctx*)h;
    for (int i = 0; i < n; i++) out[i] = sample_density_raw(c, v3(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]));
    return 0;
}

// render_kernel x n_frames (K.cu:2364-2370 + the host loop H.cpp:627-641): sum[x + y*W] += sample
// stats8 (optional): totals of {track fetches, shadow fetches, segments, opacity fetches, env
// evaluations, scatters} -- the L, S, O, E of SURVEY.md 8d.
int vo_render(void* h, float* sum, int first_frame, int n_frames, const void* param44, unsigned long long* stats8)
{
    const Ctx& c = *(Ctx*)h;
    Param      P;
    memcpy(&P, param44, sizeof(P));
    unsigned long long tot[6] = {0, 0, 0, 0, 0, 0};
    for (int f = 0; f < n_frames; f++)
    {
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : tot[:6])
        for (int64_t idx = 0; idx < (int64_t)P.width * P.height; idx++)
        {
            uint32_t  x = (uint32_t)(idx % P.width), y = (uint32_t)(idx / P.width);
            float     o4[4];
            PathStats st;
            trace_path(c, P, x, y, first_frame + f, o4, stats8 ? &st : nullptr);
            float* dst = sum + 4 * (x + (size_t)y * P.width);  // K.cu:2315
            dst[0] += o4[0]; dst[1] += o4[1]; dst[2] += o4[2]; dst[3] += o4[3];
            tot[0] += st.track_fetch; tot[1] += st.shadow_fetch; tot[2] += st.segments;
            tot[3] += st.opacity_fetch; tot[4] += st.env_eval; tot[5] += st.scatters;
        }
    }
    if (stats8) { for (int i = 0; i < 6; i++) stats8[i] = tot[i]; stats8[6] = stats8[7] = 0; }
    return 0;
}

// one path, for trace-level tests
int vo_trace_path(void* h, unsigned x, unsigned y, int frame, const void* param44, float* out4)
{
    Param P;
    memcpy(&P, param44, sizeof(P));
    trace_path(*(Ctx*)h, P, x, y, frame, out4, nullptr);
    return 0;
}

// __scale / __gamma_correct (K.cu:2333-2362)
void vo_scale(float* dst4, const float* src4, int size, float scale)
{
    for (int i = 0; i < size * 4; i++) dst4[i] = src4[i] * scale;
}
void vo_gamma_correct(float* dst4, const float* src4, int size, float scale, float gamma)
{
    float ig = 1.0f / gamma;  // K.cu:2361
    for (int i = 0; i < size; i++)
    {
        dst4[4 * i + 0] = powf(src4[4 * i + 0] * scale, ig);
        dst4[4 * i + 1] = powf(src4[4 * i + 1] * scale, ig);
        dst4[4 * i + 2] = powf(src4[4 * i + 2] * scale, ig);
        dst4[4 * i + 3] = 1.0f;
    }
}

int vo_num_threads() { return omp_get_max_threads(); }
void vo_set_num_threads(int n) { omp_set_num_threads(n); }

}  // extern "C"
