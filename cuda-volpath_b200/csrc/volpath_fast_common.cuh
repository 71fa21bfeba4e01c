// volpath_fast_common.cuh -- device helpers shared by the two production renderers (megakernel: volpath_render_fast.cu,
// wavefront: volpath_render_wave.cu): octet fetch, bound-grid lookup, Philox draws, work-item mapping, accumulation.
#pragma once
#include "volpath_common.cuh"

namespace vp
{
// two uniforms in [0, 1) from Philox2x32-10: key = pixel, counter = (draw index, frame); the counter advances
__device__ __forceinline__ void philox_draw(uint32_t key, uint32_t frame, uint32_t& ctr, float& u0, float& u1)
{
    uint32_t a, b;
    philox2x32_10(ctr++, frame, key, a, b);
    u0 = u32_to_unit_float(a);
    u1 = u32_to_unit_float(b);
}

// density at a world position inside the box: one brick-table load + one octet load.  Positions come from
// o + s * t with t inside the box interval, so cell' = floor(p * N - 0.5) + 1 is within [0, N] up to rounding;
// a single unsigned range test replaces the six clamps of the texture unit's clamp addressing.
template <int VT, bool JULIA, int LY = 0>
__device__ __forceinline__ float density_at(const Scene& S, float3 pos)
{
    if (JULIA) return julia_density(pos);
    float v[8];
    if (LY == 0 && !S.linear)
    {
        int ix = __float2int_rd(fmaf(pos.x, S.vs_scale.x, S.vs_off.x)) + 1, iy = __float2int_rd(fmaf(pos.y, S.vs_scale.y, S.vs_off.y)) + 1,
            iz = __float2int_rd(fmaf(pos.z, S.vs_scale.z, S.vs_off.z)) + 1;
        ix = clampi(ix, 1, S.nx); iy = clampi(iy, 1, S.ny); iz = clampi(iz, 1, S.nz);
        uint32_t slot = brick_slot(S, ix, iy, iz);
        if (slot == kEmptyBrick) return 0.0f;
        load_octet<VT>(S.octets, cell_in_slot(slot, ix, iy, iz), v);
        return VT == kU8 ? v[0] * (1.0f / 255.0f) : v[0];
    }
    float xb = fmaf(pos.x, S.vs_scale.x, S.vs_off_lin.x), yb = fmaf(pos.y, S.vs_scale.y, S.vs_off_lin.y),
          zb = fmaf(pos.z, S.vs_scale.z, S.vs_off_lin.z);
    float fx = floorf(xb), fy = floorf(yb), fz = floorf(zb);
    int   ix = (int)fx, iy = (int)fy, iz = (int)fz;  // vs_off_lin carries the + 1 of cell' = floor(p * N - 0.5) + 1
    if (((unsigned)ix > (unsigned)S.nx) | ((unsigned)iy > (unsigned)S.ny) | ((unsigned)iz > (unsigned)S.nz)) return 0.0f;
    uint32_t slot = brick_slot<LY>(S, ix, iy, iz);
    if (slot == kEmptyBrick) return 0.0f;
    if (VP_L2_STREAM && VT == kF32 && S.stream_octets)
        ldg256_stream(reinterpret_cast<const float4*>(S.octets) + cell_in_slot(slot, ix, iy, iz) * 2, v);
    else
        load_octet<VT>(S.octets, cell_in_slot(slot, ix, iy, iz), v);
    float a = xb - fx, b = yb - fy, g = zb - fz;
    float c00 = fmaf(a, v[1] - v[0], v[0]);
    float c10 = fmaf(a, v[3] - v[2], v[2]);
    float c01 = fmaf(a, v[5] - v[4], v[4]);
    float c11 = fmaf(a, v[7] - v[6], v[6]);
    float c0  = fmaf(b, c10 - c00, c00);
    float c1  = fmaf(b, c11 - c01, c01);
    float r   = fmaf(g, c1 - c0, c0);
    return VT == kU8 ? r * (1.0f / 255.0f) : r;
}

// The same fetch for a walk step, plus EMPTY-BRICK SKIPPING: when the position lies in a brick that stores nothing (all
// 9^3 voxels it touches are zero, so the trilinear density is exactly 0 in all of it), `skip` returns the distance along
// s to the brick's exit.  Every tentative collision up to there would be a null collision with weight exactly 1 (tracking
// without a control component) or a shadow step that cannot kill, so the walk may move to the exit without drawing:
// an exponential walk through vacuum is memoryless -- same distribution, fewer steps.  skip = 0 otherwise.
template <int VT, bool JULIA, int LY = 0>
__device__ __forceinline__ float density_at_skip(const Scene& S, float3 pos, float3 s, float& skip)
{
    skip = 0.0f;
    if (JULIA) return julia_density(pos);
    if (LY == 0 && !S.linear) return density_at<VT, JULIA, LY>(S, pos);
    float v[8];
    float xb = fmaf(pos.x, S.vs_scale.x, S.vs_off_lin.x), yb = fmaf(pos.y, S.vs_scale.y, S.vs_off_lin.y),
          zb = fmaf(pos.z, S.vs_scale.z, S.vs_off_lin.z);
    float fx = floorf(xb), fy = floorf(yb), fz = floorf(zb);
    int   ix = (int)fx, iy = (int)fy, iz = (int)fz;
    if (((unsigned)ix > (unsigned)S.nx) | ((unsigned)iy > (unsigned)S.ny) | ((unsigned)iz > (unsigned)S.nz)) return 0.0f;
    uint32_t slot = brick_slot<LY>(S, ix, iy, iz);
    if (slot == kEmptyBrick)
    {
        // exit of the brick [8b, 8b + 8) in cell' coordinates along d' = s * vs_scale
        const float dx = s.x * S.vs_scale.x, dy = s.y * S.vs_scale.y, dz = s.z * S.vs_scale.z;
        const float bx = (float)(ix & ~(kBrick - 1)), by = (float)(iy & ~(kBrick - 1)), bz = (float)(iz & ~(kBrick - 1));
        const float tx = __fdividef((dx > 0.0f ? bx + (float)kBrick : bx) - xb, dx);
        const float ty = __fdividef((dy > 0.0f ? by + (float)kBrick : by) - yb, dy);
        const float tz = __fdividef((dz > 0.0f ? bz + (float)kBrick : bz) - zb, dz);
        // 0 / 0 (on a face, moving parallel to it) is NaN and x / 0 is +inf: fminf drops the NaN, keeps the finite exits
        skip = fmaxf(fminf(fminf(tx, ty), tz), 0.0f);
        return 0.0f;
    }
    load_octet<VT>(S.octets, cell_in_slot(slot, ix, iy, iz), v);
    float a = xb - fx, b = yb - fy, g = zb - fz;
    float c00 = fmaf(a, v[1] - v[0], v[0]);
    float c10 = fmaf(a, v[3] - v[2], v[2]);
    float c01 = fmaf(a, v[5] - v[4], v[4]);
    float c11 = fmaf(a, v[7] - v[6], v[6]);
    float c0  = fmaf(b, c10 - c00, c00);
    float c1  = fmaf(b, c11 - c01, c01);
    float r   = fmaf(g, c1 - c0, c0);
    return VT == kU8 ? r * (1.0f / 255.0f) : r;
}

// When the skip pays: it costs ~25 instructions in a divergent branch of the step block, and saves the further no-op steps
// inside the same empty brick -- about (brick edge) x (majorant) of them.  Measured at full C2 dims (profiles/README.md):
// density 800 (6 expected no-op steps per brick crossing at d_max = 1): +-0 %; density 3000 (24): +12.7 %; chromatic media
// lose (the 3-channel step spills with it).  Rule, the same for every production kernel so that they keep rendering the
// same samples: gray media with density x sigma_t x brick edge >= 12.
__host__ __device__ inline bool use_brick_skip(const Scene& S, float density, float max_sig_t, bool gray)
{
    const float vmax = fmaxf(S.vs_scale.x, fmaxf(S.vs_scale.y, S.vs_scale.z));
    return gray && !S.julia && S.linear && density * max_sig_t * ((float)kBrick / vmax) >= 12.0f;
}

// opacity table of the production renderers: the cell's 8 corner values sit in one 16-byte fp16 octet (same slot and
// cell addressing as the density octets), so the lookup of K.cu:2183-2195 is one directory load + ONE vector load;
// trilinear weights as the texture unit defines them (K.cu:541-542: always linear), FMA lerps
template <int LY = 0>
__device__ __forceinline__ float opacity_at(const Scene& S, float3 pos)
{
    float xb = fmaf(pos.x, S.vs_scale.x, S.vs_off_lin.x), yb = fmaf(pos.y, S.vs_scale.y, S.vs_off_lin.y),
          zb = fmaf(pos.z, S.vs_scale.z, S.vs_off_lin.z);
    float fx = floorf(xb), fy = floorf(yb), fz = floorf(zb);
    int   ix = clampi((int)fx, 0, S.nx), iy = clampi((int)fy, 0, S.ny), iz = clampi((int)fz, 0, S.nz);
    uint32_t slot = brick_slot<LY>(S, ix, iy, iz);
    if (slot == kEmptyBrick) return 0.0f;  // only reachable where the density is zero around pos
    float v[8];
    load_octet<kF16>(S.opacity_oct, cell_in_slot(slot, ix, iy, iz), v);
    float a = xb - fx, b = yb - fy, g = zb - fz;
    float c00 = fmaf(a, v[1] - v[0], v[0]), c10 = fmaf(a, v[3] - v[2], v[2]), c01 = fmaf(a, v[5] - v[4], v[4]), c11 = fmaf(a, v[7] - v[6], v[6]);
    float c0 = fmaf(b, c10 - c00, c00), c1 = fmaf(b, c11 - c01, c01);
    return fmaf(g, c1 - c0, c0);
}

__device__ __forceinline__ uint32_t bound_cell_index(const Scene& S, float3 pos)
{
    // cell-space coordinates in one FMA per axis; the grid has < 2^31 cells
    int i = clampi(__float2int_rd(fmaf(pos.x, S.cs_scale.x, S.cs_off.x)), 0, S.ncx - 1);
    int j = clampi(__float2int_rd(fmaf(pos.y, S.cs_scale.y, S.cs_off.y)), 0, S.ncy - 1);
    int k = clampi(__float2int_rd(fmaf(pos.z, S.cs_scale.z, S.cs_off.z)), 0, S.ncz - 1);
    return (uint32_t)((k * S.ncy + j) * S.ncx + i);
}

// local (max, min) at pos from the bound grid of the fast renderer: cells of (1 << cell_log2)^3 voxels, each
// holding the (max, min) over the cell +-D voxels.  The cell edge is <= D/6, so the window is at most ~7 % wider
// than the reference's per-voxel window (and identical to it when cell_log2 == 0).
template <int LY = 0>
__device__ __forceinline__ float2 bounds_at(const Scene& S, float3 pos)
{
    const uint32_t i = bound_cell_index(S, pos);
    if (LY == 2 || (LY == 0 && S.bounds_half))
    {
        const uint32_t w = ldg_keep(S.bounds_half + i);
        return __half22float2(*reinterpret_cast<const __half2*>(&w));
    }
    return ldg_keep(S.bounds_cell + i);
}
// distance from pos toward the sun after which only vacuum follows (plus the point-to-cell-centre margin)
template <int LY = 0>
__device__ __forceinline__ float sun_clear_at(const Scene& S, float3 pos)
{
    const uint32_t i = bound_cell_index(S, pos);
    if (LY == 2 || (LY == 0 && S.sun_clear_half)) return __half2float(__ushort_as_half(ldg_keep(S.sun_clear_half + i))) + S.clear_margin;
    return ldg_keep(S.sun_clear + i) + S.clear_margin;
}

__device__ __forceinline__ float hg_eval_fast(float g, float c)
{
    float d = 1.0f + g * g - 2.0f * g * c;
    return __fdividef(1.0f - g * g, 4.0f * kPi * d * sqrt_approx(d));
}

// item -> pixel: 8x4-pixel tiles, frames innermost per tile, so the 32 lanes of a fresh claim start on one tile
__device__ __forceinline__ void item_to_sample(unsigned long long item, uint32_t n_frames, uint32_t tiles_x, uint32_t& x,
                                               uint32_t& y, uint32_t& f)
{
    uint32_t p    = (uint32_t)(item & 31u);
    uint32_t q    = (uint32_t)(item >> 5);  // the launcher keeps tiles * frames below 2^31
    uint32_t tile = q / n_frames;
    f             = q - tile * n_frames;
    x = (tile % tiles_x) * 8 + (p & 7);
    y = (tile / tiles_x) * 4 + (p >> 3);
}

__device__ __forceinline__ void accumulate(float4* __restrict__ d_sum, uint32_t pix, float3 L, int n, float brightness)
{
    // Q9: per-sample clamp (K.cu:2315-2316); one red.global.add.v4.f32.  The reference's fmaxf drops a NaN sample (its
    // weights go 0/0 for zero albedo) but lets +inf through, which then owns the pixel for the rest of the render: the
    // weighted tracker's throughput is unbounded where a local bound is exceeded and overflows about once in 10^9..10^10
    // path-samples on chromatic media (bench.py `image.nonfinite_pixels`).  The production renderers drop such a sample
    // like a NaN -- the one deliberate deviation from the reference's arithmetic.
    float4 v = make_float4(fmaxf(L.x * brightness, 0.0f), fmaxf(L.y * brightness, 0.0f), fmaxf(L.z * brightness, 0.0f), (float)n);
    if (!(v.x + v.y + v.z < 3.0e38f)) v.x = v.y = v.z = 0.0f;
    atomicAdd(d_sum + pix, v);
}

}  // namespace vp
