// volpath_common.cuh -- device-side building blocks shared by the volpath kernels (sm_100a).
//
// Reference citations: K.cu = src/volumeRender_kernel.cu, H.cpp = src/volumeRender.cpp of
// RNG65536/CUDA-volpath.  Nothing here is copied from the reference; the functions restate what the
// reference computes on our own data layout:
//
//   * density lives in HBM as a BRICKED OCTET STORE: the grid of trilinear cells (i,j,k),
//     i in [-1, N-1], is cut into 8^3-cell bricks; an L1/L2-resident rank directory maps a brick to a slot
//     in the octet pool or to EMPTY (all 8 corners of all its cells are zero); a slot holds, for
//     each of its 512 cells, the cell's 8 corner voxels contiguously (clamped at the grid border, so
//     clamp addressing costs nothing at fetch time).  A trilinear fetch is ONE table load and ONE
//     32-byte (fp32) / 16-byte (fp16) / 8-byte (u8) load -- exactly one DRAM sector -- instead of
//     eight scattered texel reads.  The 8x redundancy is paid in HBM capacity (180 GB on B200) and
//     won back by skipping empty bricks.
//   * the reference's per-voxel (max,min) bound texture (K.cu:392-412, H.cpp:1089-1267) is kept per
//     voxel for the parity renderer and per 8^3-voxel cell for the fast renderer.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vp
{
constexpr int      kBrickLog2   = 3;
constexpr int      kBrick       = 1 << kBrickLog2;           // cells per brick edge
constexpr int      kBrickCells  = kBrick * kBrick * kBrick;  // 512
constexpr int      kOpBrick     = 9 * 9 * 9;                  // opacity brick: 9^3 voxels (with apron)
constexpr int      kOpBrickPad  = 736;                        // padded to a multiple of 32 bytes
constexpr uint32_t kEmptyBrick  = 0xFFFFFFFFu;
constexpr float    kSearchRadius = 0.05f;                     // K.cu:151
constexpr int      kMaxDepth     = 800;                       // K.cu:34

enum VoxelType { kU8 = 0, kF16 = 1, kF32 = 2 };

// Everything a render kernel reads, passed as one __grid_constant__ struct.
struct Scene
{
    int    nx, ny, nz;       // voxels
    int    nbx, nby, nbz;    // bricks (cells 0..N per axis, cell' = i + 1)
    int    ncx, ncy, ncz;    // bound cells
    int    voxel_type, linear, julia, have_opacity;
    float3 bmin, bmax, l_inv;  // K.cu:155-159 (min, max, 1/(max-min))
    const uint2*    brick_words;   // per 32 bricks (x-fastest order): {occupancy bits, slot of the first set bit}
    const uint32_t* brick_table;   // flat slot table, kept only while it is small enough to live in L1/L2 (else null)
    int             stream_octets; // octet pool >> L2: load octets with L2 evict_first (see ldg256_stream)
    const void*     octets;
    const float2*   bounds_voxel;  // [nz][ny][nx] (max,min)   -- parity
    const float2*   bounds_cell;   // [ncz][ncy][ncx] (max,min) -- fast; cell = (1 << cell_log2)^3 voxels
    int             cell_log2;
    const float*    opacity;       // per brick slot: 9^3 floats (+pad) -- the bit-faithful table VP_MODE_PARITY reads
    const void*     opacity_oct;   // per cell: the 8 corner values as fp16 (16 B) -- what the production renderers read
    const float4*   env;           // [env_h][env_w]
    int             env_w, env_h;
    // env-map importance sampling (the reference's PASSIVE_ENVMAP 0 variant, K.cu:21): CDF rows + normalisation
    int             env_mis;
    const float*    env_cdf_y;     // [env_h]          (EnvmapCdfY, K.cu:1174)
    const float*    env_cdf_x;     // [env_h][env_w]   (EnvmapCdfX, K.cu:1192)
    float           env_pdfnorm_alt;  // HDRpdfnormAlt (K.cu:1166)
    float3          sun_dir, sun_power, sun_power_original;  // K.cu:1254-1256
    float           inv_view[12];                            // K.cu:626
    // fast renderer: world -> voxel space (p * N) as one FMA per axis
    float3          vs_scale, vs_off;
    float           cam_z;           // (float)(-1.0f / tan(54.43f * 0.00872664626)) evaluated on the host (K.cu:1981-1985)
    // per bound cell: distance along the sun direction beyond which no medium can be met (exact vacuum clip of
    // the shadow walk); clear_margin covers the offset between a point and its cell centre
    const float*    sun_clear;
    float           clear_margin;
    float3          vs_off_lin;      // vs_off - 0.5 + 1: position -> cell' coordinate of the linear filter (floor = cell', fraction = weight)
    float3          sun_inv;         // 1 / sun_dir (slab test of the sun shadow ray without divisions)
    float3          cs_scale, cs_off;  // world -> bound-cell space
    // large volumes (coarse bound cells): the two per-cell tables of the production renderers again in half precision,
    // rounded to the safe side (max and sun-clear up, min down, vacuum jumps = negative max toward zero): 6 B per cell
    // instead of 12, so that both stay resident in L2 next to the streaming octets (C2: 78 MB instead of 155 MB)
    const uint32_t* bounds_half;     // half2 {max, min} per cell, or null
    const uint16_t* sun_clear_half;  // half per cell, or null
    // top level of the bound grid: one half per block of (1 << top_log2)^3 bound cells = the distance a ray anywhere in
    // the block may advance in any direction without meeting medium (0: the block is not all vacuum).  At most
    // kTopCellsMax cells, so that every CTA of the production renderers can stage the whole level in shared memory.
    const uint16_t* top_jump;
    int             top_log2, ntx, nty, ntz;
};
#ifndef VP_TOP_CELLS_MAX
#define VP_TOP_CELLS_MAX 6144
#endif
constexpr int kTopCellsMax = VP_TOP_CELLS_MAX;  // 12 KB of halves per CTA

// ---- float3 helpers (operation order of src/cuda/helper_math.h) --------------------------------
__device__ __forceinline__ float3 f3(float a, float b, float c) { return make_float3(a, b, c); }
__device__ __forceinline__ float3 f3(float a) { return make_float3(a, a, a); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ float  dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross3(float3 a, float3 b)
{
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float3 normalize3(float3 v) { return v * rsqrtf(dot3(v, v)); }  // helper_math.h:1309
__device__ __forceinline__ float  max_of(float3 v) { return fmaxf(fmaxf(v.x, v.y), v.z); }
__device__ __forceinline__ float  min_of(float3 v) { return fminf(fminf(v.x, v.y), v.z); }
__device__ __forceinline__ float3 fmin3(float3 a, float3 b) { return f3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
__device__ __forceinline__ float3 fmax3(float3 a, float3 b) { return f3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }

// src/vecmath.h:9-16
constexpr float kPi     = 3.1415926535897932384626422832795028841971f;
constexpr float kTwoPi  = kPi * 2.0f;
constexpr float kPi2    = kPi / 2.0f;
constexpr float k1Pi    = 1.0f / kPi;
constexpr float k1TwoPi = 1.0f / kTwoPi;

// ---- RNG ---------------------------------------------------------------------------------------
// Reference stream (src/sampler.h:3-46): Wang hash seeding of ((x<<16)|y, frame), xoroshiro64*,
// float from the top 23 bits.  Integer work: bit-exact.
__host__ __device__ __forceinline__ uint32_t wang_hash(uint32_t seed)
{
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed *= 9u;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    return seed;
}
__device__ __forceinline__ float u32_to_unit_float(uint32_t r)  // sampler.h:24-28
{
    return __uint_as_float(0x3f800000u | (r >> 9)) - 1.0f;
}
struct RefRng
{
    uint32_t sx, sy;
    __device__ __forceinline__ uint32_t next_u32()
    {
        uint32_t result = sx * 0x9e3779bbu;
        sy ^= sx;
        sx = __funnelshift_l(sx, sx, 26) ^ sy ^ (sy << 9);
        sy = __funnelshift_l(sx, sx, 13);
        return result;
    }
    __device__ __forceinline__ void init(uint32_t px, uint32_t py, uint32_t frame)
    {
        sx = wang_hash((px << 16) | py);
        sy = wang_hash(frame);
        next_u32();
    }
    __device__ __forceinline__ float next() { return u32_to_unit_float(next_u32()); }
};

// Philox2x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel Random Numbers: As Easy as 1, 2, 3", SC'11):
// counter (c0, c1), key k; 10 rounds of {hi,lo = M*c0; (c0,c1) = (hi^k^c1, lo); k += W}.
// The fast renderer uses counter = (draw index, frame), key = pixel: any (pixel, frame, draw) triple
// is addressable without carried state, so a regenerated lane needs no RNG hand-over.
__host__ __device__ __forceinline__ void philox2x32_10(uint32_t c0, uint32_t c1, uint32_t k, uint32_t& o0, uint32_t& o1)
{
#pragma unroll
    for (int r = 0; r < 10; r++)
    {
        uint64_t p  = (uint64_t)0xD256D193u * c0;
        uint32_t hi = (uint32_t)(p >> 32), lo = (uint32_t)p;
        c0 = hi ^ k ^ c1;
        c1 = lo;
        k += 0x9E3779B9u;
    }
    o0 = c0;
    o1 = c1;
}

// ---- L2 residency control (sm_100a) -------------------------------------------------------------
// At full C2 the octet pool (54 GB) streams through the 126 MB L2 at > 3 TB/s and would evict the small tables every
// path reads again and again (bound grid 103 MB, sun-clear 52 MB, rank directory 3 MB).  The production kernels
// therefore load the tables with an L2 evict_last policy (+0.7 %).  Loading the octets with ONE 256-bit evict_first load
// (LDG.E.EFL2.256, Blackwell's 32-byte vector load) instead of two 128-bit ones was measured too and LOSES 4 %: the
// walk re-reads neighbouring octets of the same 128-byte line within a few steps, so the pool does have short-term L2
// reuse worth keeping.  The code path stays behind VP_L2_STREAM for the record.
__device__ __forceinline__ uint64_t l2_policy_keep()
{
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
#ifndef VP_L2_KEEP
#define VP_L2_KEEP 1
#endif
#ifndef VP_L2_STREAM
#define VP_L2_STREAM 0  // measured on the full C2 volume: no hints 939, keep only 946, stream only 900, both 907 M path-samples/s
#endif
__device__ __forceinline__ float2 ldg_keep(const float2* a)
{
    if (!VP_L2_KEEP) return __ldg(a);
    float2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(a), "l"(l2_policy_keep()));
    return v;
}
__device__ __forceinline__ float ldg_keep(const float* a)
{
    if (!VP_L2_KEEP) return __ldg(a);
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(l2_policy_keep()));
    return v;
}
#ifndef VP_TABLE_L1_KEEP
#define VP_TABLE_L1_KEEP 0  // experiment: 1 = the small tables are also loaded with L1::evict_last
#endif
__device__ __forceinline__ uint32_t ldg_keep(const uint32_t* a)
{
    if (!VP_L2_KEEP) return __ldg(a);
    uint32_t v;
    if (VP_TABLE_L1_KEEP)
        asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(l2_policy_keep()));
    else
        asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(l2_policy_keep()));
    return v;
}
__device__ __forceinline__ uint16_t ldg_keep(const uint16_t* a)
{
    if (!VP_L2_KEEP) return __ldg(a);
    uint16_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(a), "l"(l2_policy_keep()));
    return v;
}
__device__ __forceinline__ uint2 ldg_keep(const uint2* a)
{
    if (!VP_L2_KEEP) return __ldg(a);
    uint2 v;
    if (VP_TABLE_L1_KEEP)
        asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(a), "l"(l2_policy_keep()));
    else
        asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(a), "l"(l2_policy_keep()));
    return v;
}
__device__ __forceinline__ void ldg256_stream(const void* a, float v[8])
{
    uint32_t w[8];
    asm volatile("ld.global.nc.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(a));
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(w[i]);
}

// ---- octet store -------------------------------------------------------------------------------
template <int VT> struct OctetBytes;
template <> struct OctetBytes<kU8> { static constexpr int value = 8; };
template <> struct OctetBytes<kF16> { static constexpr int value = 16; };
template <> struct OctetBytes<kF32> { static constexpr int value = 32; };

// the 8 corners of one cell as floats, v[dz*4 + dy*2 + dx]; u8 is returned UN-normalised (0..255)
template <int VT>
__device__ __forceinline__ void load_octet(const void* pool, size_t cell_index, float v[8])
{
#ifndef VP_OCTET_256
#define VP_OCTET_256 0  // one 256-bit load per fp32 octet (Blackwell) instead of two 128-bit ones: measured, see profiles/README.md
#endif
    if (VT == kF32 && VP_OCTET_256)
    {
        const float4* p = reinterpret_cast<const float4*>(pool) + cell_index * 2;
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                     : "l"(p));
    }
#ifndef VP_OCTET_L1
#define VP_OCTET_L1 0  // experiment: 1 = octet loads do not allocate in L1 (leave it to the directory and the bound tables)
#endif
    else if (VT == kF32 && VP_OCTET_L1 == 1)
    {
        const float4* p = reinterpret_cast<const float4*>(pool) + cell_index * 2;
        float4        a, b;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 1));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    else if (VT == kF32)
    {
        const float4* p = reinterpret_cast<const float4*>(pool) + cell_index * 2;
        float4        a = __ldg(p), b = __ldg(p + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    else if (VT == kF16)
    {
        uint4   a = __ldg(reinterpret_cast<const uint4*>(pool) + cell_index);
        __half2 h0 = *reinterpret_cast<__half2*>(&a.x), h1 = *reinterpret_cast<__half2*>(&a.y);
        __half2 h2 = *reinterpret_cast<__half2*>(&a.z), h3 = *reinterpret_cast<__half2*>(&a.w);
        float2  f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3_ = __half22float2(h3);
        v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y; v[4] = f2.x; v[5] = f2.y; v[6] = f3_.x; v[7] = f3_.y;
    }
    else
    {
        uint2 a = __ldg(reinterpret_cast<const uint2*>(pool) + cell_index);
        v[0] = (float)(a.x & 0xff); v[1] = (float)((a.x >> 8) & 0xff); v[2] = (float)((a.x >> 16) & 0xff); v[3] = (float)(a.x >> 24);
        v[4] = (float)(a.y & 0xff); v[5] = (float)((a.y >> 8) & 0xff); v[6] = (float)((a.y >> 16) & 0xff); v[7] = (float)(a.y >> 24);
    }
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return max(lo, min(v, hi)); }

// slot lookup for cell' = (cx, cy, cz), each in [0, N]
// LY ("layout") lets a kernel variant fix at compile time what the scene decides at run time: 0 = generic (test the scene),
// 1 = small volume (flat brick table, float per-cell tables), 2 = large volume (rank directory, half-precision tables);
// 1 and 2 also imply the linear filter and a sun-clear table.  The tests are warp-uniform branches, but they sit in the
// fetch and in the segment loop: three instructions each, every time.
template <int LY = 0>
__device__ __forceinline__ uint32_t brick_slot(const Scene& S, int cx, int cy, int cz)
{
    int bx = cx >> kBrickLog2, by = cy >> kBrickLog2, bz = cz >> kBrickLog2;
    // rank directory instead of a 4-byte table entry per brick: slots are numbered in brick order, so
    // slot = prefix(word) + popc(bits below).  8 bytes per 32 bricks (C2: 3.2 MB instead of 52 MB) -- the first of
    // the two dependent loads of a density fetch now hits L1/L2 instead of competing with the octets for L2.
    const uint32_t b   = (uint32_t)((bz * S.nby + by) * S.nbx + bx);  // < 2^31 bricks (dims <= 8184)
    if (LY == 1 || (LY == 0 && S.brick_table)) return __ldg(S.brick_table + b);  // small volumes: 4 MB of flat table is cache-resident anyway
    const uint2    w   = ldg_keep(S.brick_words + (b >> 5));
    const uint32_t bit = 1u << (b & 31u);
    return (w.x & bit) ? w.y + __popc(w.x & (bit - 1u)) : kEmptyBrick;
}
// Order of the 512 cells inside a brick slot.  A miss fills a whole 128-byte line whatever the load asks for
// (tools/sector_probe.cu), so VP_CELL_ORDER 1 makes a line a compact block instead of a run along x: the low three index
// bits are (z0, y0, x0), i.e. 2x2x1 cells per line for 32-byte fp32 octets, 2x2x2 for 16-byte fp16 octets.
#ifndef VP_CELL_ORDER
#define VP_CELL_ORDER 0
#endif
__host__ __device__ __forceinline__ uint32_t cell_local(int cx, int cy, int cz)
{
    cx &= kBrick - 1; cy &= kBrick - 1; cz &= kBrick - 1;
#if VP_CELL_ORDER
    return (uint32_t)(((cz >> 1) << 7) | ((cy >> 1) << 5) | ((cx >> 1) << 3) | ((cz & 1) << 2) | ((cy & 1) << 1) | (cx & 1));
#else
    return (uint32_t)((cz << (2 * kBrickLog2)) | (cy << kBrickLog2) | cx);
#endif
}
__device__ __forceinline__ size_t cell_in_slot(uint32_t slot, int cx, int cy, int cz)
{
    return (size_t)slot * kBrickCells + cell_local(cx, cy, cz);
}

// PARITY density fetch: restates the CUDA texture unit exactly as oracle/tex_emul.h does
// (K.cu:173-178 sample_w: p = (pos-min)*l_inv; clamp addressing; point: floor(p*N); linear:
// xB = p*N - 0.5, 1.8 fixed-point weights, (1-a)*p + a*q without contraction; u8 -> /255).
template <int VT>
__device__ __forceinline__ float fetch_density_parity(const Scene& S, float3 pos)
{
    float3 p = (pos - S.bmin) * S.l_inv;
    float  x = __fmul_rn(p.x, (float)S.nx), y = __fmul_rn(p.y, (float)S.ny), z = __fmul_rn(p.z, (float)S.nz);
    float  v[8];
    if (!S.linear)
    {
        int      ix = clampi((int)floorf(x), 0, S.nx - 1) + 1, iy = clampi((int)floorf(y), 0, S.ny - 1) + 1,
                 iz = clampi((int)floorf(z), 0, S.nz - 1) + 1;
        uint32_t slot = brick_slot(S, ix, iy, iz);
        if (slot == kEmptyBrick) return 0.0f;
        load_octet<VT>(S.octets, cell_in_slot(slot, ix, iy, iz), v);
        return VT == kU8 ? __fdiv_rn(v[0], 255.0f) : v[0];
    }
    float xb = __fsub_rn(x, 0.5f), yb = __fsub_rn(y, 0.5f), zb = __fsub_rn(z, 0.5f);
    float fx = floorf(xb), fy = floorf(yb), fz = floorf(zb);
    float a = __fsub_rn(xb, fx), b = __fsub_rn(yb, fy), g = __fsub_rn(zb, fz);
    a = __fmul_rn(floorf(__fadd_rn(__fmul_rn(a, 256.0f), 0.5f)), 1.0f / 256.0f);
    b = __fmul_rn(floorf(__fadd_rn(__fmul_rn(b, 256.0f), 0.5f)), 1.0f / 256.0f);
    g = __fmul_rn(floorf(__fadd_rn(__fmul_rn(g, 256.0f), 0.5f)), 1.0f / 256.0f);
    // cell' = floor(xB) + 1 clamped to [0, N]; the stored corners are already border-clamped
    int ix = clampi((int)fx + 1, 0, S.nx), iy = clampi((int)fy + 1, 0, S.ny), iz = clampi((int)fz + 1, 0, S.nz);
    uint32_t slot = brick_slot(S, ix, iy, iz);
    if (slot == kEmptyBrick) return 0.0f;
    load_octet<VT>(S.octets, cell_in_slot(slot, ix, iy, iz), v);
    if (VT == kU8)
    {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __fdiv_rn(v[i], 255.0f);
    }
    float ia = __fsub_rn(1.0f, a), ib = __fsub_rn(1.0f, b), ig = __fsub_rn(1.0f, g);
    float c00 = __fadd_rn(__fmul_rn(ia, v[0]), __fmul_rn(a, v[1]));
    float c10 = __fadd_rn(__fmul_rn(ia, v[2]), __fmul_rn(a, v[3]));
    float c01 = __fadd_rn(__fmul_rn(ia, v[4]), __fmul_rn(a, v[5]));
    float c11 = __fadd_rn(__fmul_rn(ia, v[6]), __fmul_rn(a, v[7]));
    float c0  = __fadd_rn(__fmul_rn(ib, c00), __fmul_rn(b, c10));
    float c1  = __fadd_rn(__fmul_rn(ib, c01), __fmul_rn(b, c11));
    return __fadd_rn(__fmul_rn(ig, c0), __fmul_rn(g, c1));
}

// Procedural density of the reference's no-OpenVDB build (K.cu:84-140): quaternion Julia set,
// q <- q^2 + c until |q|^2 >= 10 or 31 iterations; density = (iterations > 27).
__device__ __forceinline__ float julia_density(float3 pos)
{
    const float radius = 1.4f;
    float       qx = pos.x * radius, qy = pos.y * radius, qz = pos.z * radius, qw = 0.0f;
    int         iter = 0;
    float       d;
    do
    {
        float r0 = qx * qx - (qy * qy + qz * qz + qw * qw);
        float t  = qx * 2;
        float ry = qy * t, rz = qz * t, rw = qw * t;
        qx = r0 + -0.2f; qy = ry + 0.8f; qz = rz + 0.0f; qw = rw + 0.0f;
        d  = qx * qx + qy * qy + qz * qz + qw * qw;
    } while (d < 10.0f && iter++ < 30);
    return (float)(iter > 27);  // iter > 30 * 0.9
}

// opacity table: trilinear (1.8 fixed-point weights, always linear: K.cu:541-542) from 9^3 apron bricks
__device__ __forceinline__ float fetch_opacity(const Scene& S, float3 pos, bool parity)
{
    float3 p = (pos - S.bmin) * S.l_inv;
    float  x = __fmul_rn(p.x, (float)S.nx), y = __fmul_rn(p.y, (float)S.ny), z = __fmul_rn(p.z, (float)S.nz);
    float  xb = __fsub_rn(x, 0.5f), yb = __fsub_rn(y, 0.5f), zb = __fsub_rn(z, 0.5f);
    float  fx = floorf(xb), fy = floorf(yb), fz = floorf(zb);
    float  a = __fsub_rn(xb, fx), b = __fsub_rn(yb, fy), g = __fsub_rn(zb, fz);
    if (parity)
    {
        a = __fmul_rn(floorf(__fadd_rn(__fmul_rn(a, 256.0f), 0.5f)), 1.0f / 256.0f);
        b = __fmul_rn(floorf(__fadd_rn(__fmul_rn(b, 256.0f), 0.5f)), 1.0f / 256.0f);
        g = __fmul_rn(floorf(__fadd_rn(__fmul_rn(g, 256.0f), 0.5f)), 1.0f / 256.0f);
    }
    int ix = clampi((int)fx + 1, 0, S.nx), iy = clampi((int)fy + 1, 0, S.ny), iz = clampi((int)fz + 1, 0, S.nz);
    uint32_t slot = brick_slot(S, ix, iy, iz);
    if (slot == kEmptyBrick) return 0.0f;  // only reachable where the density is zero around pos
    const float* B  = S.opacity + (size_t)slot * kOpBrickPad;
    int          lx = ix & (kBrick - 1), ly = iy & (kBrick - 1), lz = iz & (kBrick - 1);
    const float* q  = B + (lz * 9 + ly) * 9 + lx;
    float v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 9), v3 = __ldg(q + 10);
    float v4 = __ldg(q + 81), v5 = __ldg(q + 82), v6 = __ldg(q + 90), v7 = __ldg(q + 91);
    float ia = __fsub_rn(1.0f, a), ib = __fsub_rn(1.0f, b), ig = __fsub_rn(1.0f, g);
    float c00 = __fadd_rn(__fmul_rn(ia, v0), __fmul_rn(a, v1));
    float c10 = __fadd_rn(__fmul_rn(ia, v2), __fmul_rn(a, v3));
    float c01 = __fadd_rn(__fmul_rn(ia, v4), __fmul_rn(a, v5));
    float c11 = __fadd_rn(__fmul_rn(ia, v6), __fmul_rn(a, v7));
    float c0  = __fadd_rn(__fmul_rn(ib, c00), __fmul_rn(b, c10));
    float c1  = __fadd_rn(__fmul_rn(ib, c01), __fmul_rn(b, c11));
    return __fadd_rn(__fmul_rn(ig, c0), __fmul_rn(g, c1));
}

// ---- environment / sun (K.cu:882-895, 956-973, 1258-1267) -----------------------------------------
__device__ __forceinline__ float3 eval_envmap(const Scene& S, float3 dir)
{
    float phi   = acosf(dir.y);
    float theta = atanf(dir.z / dir.x) + kPi2;
    if (dir.x < 0) theta += kPi;
    float u = theta * k1TwoPi;
    float v = phi * k1Pi;
    // point filter, normalized coordinates, clamp (legacy texture reference defaults, K.cu:1099-1100)
    int    i = clampi((int)floorf(__fmul_rn(u, (float)S.env_w)), 0, S.env_w - 1);
    int    j = clampi((int)floorf(__fmul_rn(v, (float)S.env_h)), 0, S.env_h - 1);
    float4 c = __ldg(S.env + (size_t)j * S.env_w + i);
    return f3(c.x, c.y, c.z);
}
// env-map importance sampling (K.cu:904-1034, MULT_PDF 0, PRE_WARP 1)
__device__ __forceinline__ float env_luminance(float3 c) { return (float)(c.x * 0.2126 + c.y * 0.7152 + c.z * 0.0722); }  // K.cu:951
__device__ __forceinline__ int env_sample_row(const float* __restrict__ cdf, int n, float r)  // sample_y / sample_x
{
    int begin = 0, end = n - 1;
    while (end > begin)
    {
        int mid = begin + (end - begin) / 2;
        if (__ldg(cdf + mid) >= r) end = mid; else begin = mid + 1;
    }
    return begin;
}
__device__ __forceinline__ float sample_envmap(const Scene& S, float& u, float& v, float3& col)  // K.cu:979-1009
{
    int iy = env_sample_row(S.env_cdf_y, S.env_h, v);
    int ix = env_sample_row(S.env_cdf_x + (size_t)iy * S.env_w, S.env_w, u);
    u      = ((float)ix + 0.5f) / (float)S.env_w;
    v      = ((float)iy + 0.5f) / (float)S.env_h;
    int    i = clampi((int)floorf(__fmul_rn(u, (float)S.env_w)), 0, S.env_w - 1);
    int    j = clampi((int)floorf(__fmul_rn(v, (float)S.env_h)), 0, S.env_h - 1);
    float4 c = __ldg(S.env + (size_t)j * S.env_w + i);
    col      = f3(c.x, c.y, c.z);
    return env_luminance(col) * S.env_pdfnorm_alt;
}
__device__ __forceinline__ float pdf_envmap(const Scene& S, float3 col) { return env_luminance(col) * S.env_pdfnorm_alt; }  // K.cu:1011
__device__ __forceinline__ float3 uv_to_dir(float u, float v)  // K.cu:897-902
{
    float theta = u * kTwoPi;
    float phi   = v * kPi;
    return f3(sinf(phi) * sinf(theta), cosf(phi), sinf(phi) * -cosf(theta));
}

__device__ __forceinline__ float3 background(const Scene& S, float3 dir, int depth)
{
    if (depth == 0 && (dot3(dir, S.sun_dir) > 94.0f / sqrtf(94.0f * 94.0f + 0.45f * 0.45f))) return S.sun_power_original;
    return eval_envmap(S, dir);
}

// Henyey-Greenstein (K.cu:575-619); sampling clamps cos(theta) to [0,1] like the reference (Q3)
__device__ __forceinline__ float hg_evaluate(float g, float cos_theta)
{
    return (1.0f - g * g) / (4.0f * kPi * powf(1.0f + g * g - 2 * g * cos_theta, 1.5f));
}
__device__ __forceinline__ float3 hg_sample_local(float g, float rnd0, float rnd1)
{
    float cos_theta;
    if (fabsf(g) > 1e-6f)
    {
        float s   = 2.0f * rnd0 - 1.0f;
        float f   = (1.0f - g * g) / (1.0f + g * s);
        cos_theta = (0.5f / g) * (1.0f + g * g - f * f);
        cos_theta = fmaxf(0.0f, fminf(1.0f, cos_theta));
    }
    else
    {
        cos_theta = 2.0f * rnd0 - 1.0f;
    }
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    float phi       = 2.0f * kPi * rnd1;
    return f3(cosf(phi) * sin_theta, sinf(phi) * sin_theta, cos_theta);
}
// Frame (K.cu:557-573): orthonormal basis around n
__device__ __forceinline__ void make_frame(float3 n, float3& t, float3& b)
{
    float3 a = (double)fabsf(n.x) > 0.1 ? f3(0, 1, 0) : f3(1, 0, 0);
    t        = normalize3(cross3(a, n));
    b        = cross3(n, t);
}

// slab test shared by intersectBox / intersect_box / intersectSuperVolume (K.cu:453-481, 654-680, 1626-1661)
__device__ __forceinline__ void box_slabs(const Scene& S, float3 o, float3 d, float& largest_tmin, float& smallest_tmax)
{
    float3 invR = f3(1.0f) / d;
    float3 tbot = invR * (S.bmin - o);
    float3 ttop = invR * (S.bmax - o);
    largest_tmin  = max_of(fmin3(ttop, tbot));
    smallest_tmax = min_of(fmax3(ttop, tbot));
}

// camera ray of pixel (x, y) (K.cu:1977-1987).  tan() is evaluated in double like the reference (Q7).
__device__ __forceinline__ void camera_ray(const Scene& S, uint32_t x, uint32_t y, uint32_t W, uint32_t H, float3& o, float3& d)
{
    float u    = (x * 2.0f - W) / W;
    float v    = (y * 2.0f - H) / W;
    float fovx = 54.43;
    const float* M = S.inv_view;
    o = f3(M[3], M[7], M[11]);
    float3 dc = f3(u, v, (float)(-1.0f / tan(fovx * 0.00872664626)));
    d = normalize3(f3(dot3(dc, f3(M[0], M[1], M[2])), dot3(dc, f3(M[4], M[5], M[6])), dot3(dc, f3(M[8], M[9], M[10]))));
}
// fast-math variants for the production renderer (equal in distribution, not bit-for-bit)
__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float3 hg_sample_local_fast(float g, float rnd0, float rnd1)
{
    float cos_theta;
    if (fabsf(g) > 1e-6f)
    {
        float s   = 2.0f * rnd0 - 1.0f;
        float f   = __fdividef(1.0f - g * g, 1.0f + g * s);
        cos_theta = __fdividef(0.5f, g) * (1.0f + g * g - f * f);
        cos_theta = fmaxf(0.0f, fminf(1.0f, cos_theta));  // Q3
    }
    else
    {
        cos_theta = 2.0f * rnd0 - 1.0f;
    }
    float sin_theta = sqrt_approx(fmaxf(0.0f, 1.0f - cos_theta * cos_theta));
    float sp, cp;
    __sincosf(2.0f * kPi * rnd1, &sp, &cp);
    return f3(cp * sin_theta, sp * sin_theta, cos_theta);
}
// the reference's frame (K.cu:557-573) with the two cases of its helper axis written out
__device__ __forceinline__ void make_frame_fast(float3 n, float3& t, float3& b)
{
    const bool big = fabsf(n.x) > 0.1f;
    float3     c   = big ? f3(n.z, 0.0f, -n.x) : f3(0.0f, -n.z, n.y);
    t              = c * rsqrtf(n.z * n.z + (big ? n.x * n.x : n.y * n.y));
    b              = cross3(n, t);
}
// slab test with a precomputed reciprocal direction
__device__ __forceinline__ void box_slabs_inv(const Scene& S, float3 o, float3 invR, float& largest_tmin, float& smallest_tmax)
{
    float3 tbot = invR * (S.bmin - o);
    float3 ttop = invR * (S.bmax - o);
    largest_tmin  = max_of(fmin3(ttop, tbot));
    smallest_tmax = min_of(fmax3(ttop, tbot));
}
__device__ __forceinline__ void box_slabs_fast(const Scene& S, float3 o, float3 d, float& largest_tmin, float& smallest_tmax)
{
    float3 invR = f3(__fdividef(1.0f, d.x), __fdividef(1.0f, d.y), __fdividef(1.0f, d.z));
    float3 tbot = invR * (S.bmin - o);
    float3 ttop = invR * (S.bmax - o);
    largest_tmin  = max_of(fmin3(ttop, tbot));
    smallest_tmax = min_of(fmax3(ttop, tbot));
}

// same ray with the per-launch constant hoisted to the host (fast renderer)
__device__ __forceinline__ void camera_ray_fast(const Scene& S, uint32_t x, uint32_t y, uint32_t W, uint32_t H, float3& o, float3& d)
{
    float u = (x * 2.0f - W) / W;
    float v = (y * 2.0f - H) / W;
    const float* M = S.inv_view;
    o = f3(M[3], M[7], M[11]);
    float3 dc = f3(u, v, S.cam_z);
    d = normalize3(f3(dot3(dc, f3(M[0], M[1], M[2])), dot3(dc, f3(M[4], M[5], M[6])), dot3(dc, f3(M[8], M[9], M[10]))));
}
}  // namespace vp
