// volpath_render_fast.cu -- the B200 production renderer (VP_MODE_FAST), megakernel form.
//
// Same estimator as the reference's __d_render_bounded_decomp (K.cu:1958-2318) in distribution -- weighted
// spectral delta tracking against a local majorant, analog decomposition where the local minimum is
// positive (with the reference's quirks Q1-Q4), reduced-scattering switch after 5
// bounces, sun NEE by shadow walk or opacity table, HG phase sampling -- re-organised for the machine:
//
//   * a CTA owns a 16x8 pixel tile for ALL frames of the launch and keeps its float4 sums in shared memory;
//     lanes pull (pixel, frame) items from a shared counter, so a lane whose path ended regenerates at
//     once (no warp waits for its longest path: the reference's one-thread-one-pixel launch idles ~70 % of
//     its lanes on clouds).  One global RMW per pixel per launch instead of one per frame.
//   * one loop body serves both walks (primary tracking and the sun shadow walk): draw, advance, fetch,
//     decide -- the density fetches of all 32 lanes are issued together whatever each lane is doing.
//   * the ray is intersected with the box once; the reference's 0.05-step approach march outside the
//     medium (Q5: ~60 of ~65 segments per path) and segments whose local max is exactly zero are skipped
//     without random draws (an exponential walk through vacuum is memoryless: same distribution).
//   * counter-based Philox2x32-10: key = pixel, counter = (draw index, frame); no carried RNG state.
//   * density comes from the octet store: one table load + one 32/16/8-byte load per trilinear sample.
#include "volpath_common.cuh"
#include "volpath_kernels.h"

namespace vp
{
constexpr int kTileW = 16, kTileH = 8, kTilePix = kTileW * kTileH;  // = threads per CTA

enum : uint32_t
{
    kModePath  = 0,  // needs a new (pixel, frame) item
    kModeRay   = 1,  // (o, s) set: intersect the box
    kModeSeg   = 2,  // find the next segment with medium and set its majorants
    kModeStep  = 3,  // walking
    kModeIdle  = 4,  // the tile's items are exhausted: wait for the rest of the warp
    kModeMask  = 7,
    kShadow    = 8,    // walking toward the sun (else: primary tracking)
    kLimIsCtrl = 16,   // `lim` is the control-component collision distance (else: segment end)
    kKillX = 32, kKillY = 64, kKillZ = 128,
};

struct Philox
{
    uint32_t key, frame, ctr;
    __device__ __forceinline__ void draw(float& u0, float& u1)
    {
        uint32_t a, b;
        philox2x32_10(ctr++, frame, key, a, b);
        u0 = u32_to_unit_float(a);
        u1 = u32_to_unit_float(b);
    }
};

template <int VT, bool JULIA>
__device__ __forceinline__ float density_at(const Scene& S, float3 pos)
{
    if (JULIA) return julia_density(pos);
    return fetch_density_fast<VT>(S, fmaf(pos.x, S.vs_scale.x, S.vs_off.x), fmaf(pos.y, S.vs_scale.y, S.vs_off.y),
                                  fmaf(pos.z, S.vs_scale.z, S.vs_off.z));
}

// local (max, min) at pos from the bound grid of the fast renderer: cells of (1 << cell_log2)^3 voxels, each
// holding the (max, min) over the cell +-D voxels.  The cell edge is <= D/6, so the window is at most ~7 % wider
// than the reference's per-voxel window (and identical to it when cell_log2 == 0).
__device__ __forceinline__ float2 bounds_at(const Scene& S, float3 pos)
{
    int i = clampi(__float2int_rd(fmaf(pos.x, S.vs_scale.x, S.vs_off.x)), 0, S.nx - 1) >> S.cell_log2;
    int j = clampi(__float2int_rd(fmaf(pos.y, S.vs_scale.y, S.vs_off.y)), 0, S.ny - 1) >> S.cell_log2;
    int k = clampi(__float2int_rd(fmaf(pos.z, S.vs_scale.z, S.vs_off.z)), 0, S.nz - 1) >> S.cell_log2;
    return __ldg(S.bounds_cell + ((size_t)k * S.ncy + j) * S.ncx + i);
}

__device__ __forceinline__ float hg_eval_fast(float g, float c)
{
    float d = 1.0f + g * g - 2.0f * g * c;
    return __fdividef(1.0f - g * g, 4.0f * kPi * d * sqrtf(d));
}

template <int VT, bool JULIA, bool STATS>
__global__ void __launch_bounds__(kTilePix) k_render_fast(const __grid_constant__ Scene S, float4* __restrict__ d_sum,
                                                           int first_frame, int n_frames, int frame_stride,
                                                           const __grid_constant__ vp_param P, int tiles_x,
                                                           unsigned long long* __restrict__ d_stats)
{
    __shared__ float    acc[kTilePix * 4];
    __shared__ uint32_t next_item;
    const int           tid = threadIdx.x;
    acc[tid * 4 + 0] = acc[tid * 4 + 1] = acc[tid * 4 + 2] = acc[tid * 4 + 3] = 0.0f;
    if (tid == 0) next_item = kTilePix;
    __syncthreads();

    const uint32_t tile_x0 = (blockIdx.x % tiles_x) * kTileW, tile_y0 = (blockIdx.x / tiles_x) * kTileH;
    const uint32_t n_items = (uint32_t)kTilePix * (uint32_t)n_frames;

    const float3 sig_t = f3(P.sigma_t.x, P.sigma_t.y, P.sigma_t.z);
    const float3 sig_s = sig_t * f3(P.albedo.x, P.albedo.y, P.albedo.z);
    const float  max_sig_t = max_of(sig_t), min_sig_t = min_of(sig_t);

    // lane state
    float3   o = f3(0.f), s = f3(0.f), pend = f3(0.f), T = f3(1.f), L = f3(0.f);
    float    dist = 0.f, lim = 0.f, inv = 0.f, dens = 0.f, maj = 0.f, sigc = 0.f, t_exit = 0.f, ph = 0.f, dmax = 0.f;
    int      n = 0;
    uint32_t st = kModePath, item = tid, pslot = 0;
    Philox   rng{0, 0, 0};
    unsigned long long c_track = 0, c_shadow = 0, c_seg = 0, c_op = 0, c_env = 0, c_scat = 0;

    for (;;)
    {
        // warp-converged loop head: every lane (idle ones included) votes here, so a lane that ran out of work
        // never leaves its warp-mates waiting at a warp-level barrier
        if (!__any_sync(0xffffffffu, (st & kModeMask) != kModeIdle)) break;
        if ((st & kModeMask) == kModePath)
        {
            if (item >= n_items)
            {
                st = kModeIdle;
                continue;
            }
            pslot = item & (kTilePix - 1);
            uint32_t f  = item >> 7;
            uint32_t lane = pslot & 31, wrp = pslot >> 5;
            uint32_t x = tile_x0 + (wrp & 1) * 8 + (lane & 7), y = tile_y0 + (wrp >> 1) * 4 + (lane >> 3);
            item = atomicAdd(&next_item, 1u);
            if (x >= P.width || y >= P.height) continue;
            rng.key   = y * P.width + x;
            rng.frame = (uint32_t)(first_frame + (int)f * frame_stride);
            rng.ctr   = 0;
            camera_ray_fast(S, x, y, P.width, P.height, o, s);
            T = f3(1.f);
            L = f3(0.f);
            n  = 0;
            st = kModeRay;
        }
        if ((st & kModeMask) == kModeRay)
        {
            // one slab test per ray (the reference repeats it every 0.05 step, K.cu:1626-1661)
            float tn, tf;
            box_slabs(S, o, s, tn, tf);
            bool hit = tf > tn && tf >= 1e-3f;
            dist   = fmaxf(tn, 0.0f);
            t_exit = hit ? tf : -1.0f;
            st     = kModeSeg;
        }
        if ((st & kModeMask) == kModeSeg)
        {
            bool found = false;
            while (dist < t_exit)
            {
                if (STATS) c_seg++;
                float seg_end = JULIA ? t_exit : fminf(dist + kSearchRadius, t_exit);
                float2 bnd    = JULIA ? make_float2(1.0f, 0.0f) : bounds_at(S, o + s * dist);
                if (bnd.x <= 0.0f)
                {
                    dist = seg_end;  // no medium within reach: the walk passes with probability 1
                    continue;
                }
                dmax = fmaxf(1e-4f, bnd.x);
                // reduced scattering after 5 bounces (K.cu:2039-2044)
                float sr = fmaxf(0.0f, fminf(1.0f, (n - 5) * 0.066666666666666666667f));
                dens     = ((1 - sr) + sr * (1 - P.g)) * P.density;
                maj      = max_sig_t * dens * dmax;
                lim      = seg_end;
                st       = kModeStep;
                if (bnd.y > 0.0f)  // analog decomposition (K.cu:2048-2054, Q2)
                {
                    float u0, u1;
                    rng.draw(u0, u1);
                    sigc        = min_sig_t * dens * bnd.y;
                    float distc = dist - __fdividef(__logf(u0), fmaxf(sigc, 1e-20f));
                    inv         = __fdividef(1.0f, fmaxf(maj - sigc, 1e-20f));
                    if (distc < seg_end)
                    {
                        lim = distc;
                        st |= kLimIsCtrl;
                    }
                }
                else
                {
                    sigc = 0.0f;
                    inv  = __fdividef(1.0f, maj);
                }
                found = true;
                break;
            }
            if (!found)
            {
                // escaped (or never hit): environment / sun disk, then the path is complete
                if (STATS) c_env++;
                L = L + background(S, s, n) * T;
                goto path_done;
            }
        }
        if ((st & kModeMask) != kModeStep) continue;
        {
            // ---- one step of whichever walk this lane is on ----
            float u0, u1;
            rng.draw(u0, u1);
            dist += -__logf(u0) * inv;
            const bool past = dist >= lim;
            float3     pos  = o + s * (past ? lim : dist);
            float      den  = 0.0f;
            if (!past)
            {
                den = density_at<VT, JULIA>(S, pos) * dens;
                if (STATS) { if (st & kShadow) c_shadow++; else c_track++; }
            }
            if (st & kShadow)
            {
                if (!past)
                {
                    // Tr_spectral (K.cu:782-806): per-channel kill flags on one shared walk
                    float q = den * inv;
                    if (u1 < sig_t.x * q) st |= kKillX;
                    if (u1 < sig_t.y * q) st |= kKillY;
                    if (u1 < sig_t.z * q) st |= kKillZ;
                }
                if (past || (st & (kKillX | kKillY | kKillZ)) == (kKillX | kKillY | kKillZ))
                {
                    float3 a = f3((st & kKillX) ? 0.f : 1.f, (st & kKillY) ? 0.f : 1.f, (st & kKillZ) ? 0.f : 1.f);
                    L        = L + S.sun_power * (T * ph * a);
                    s        = pend;
                    st       = kModeRay;
                    if (n >= kMaxDepth) goto path_done;
                }
                continue;
            }
            bool scatter;
            if (past)
            {
                scatter = (st & kLimIsCtrl) != 0;  // control-component collision: no weight (Q2)
                if (!scatter)
                {
                    dist = lim;  // crossed the segment: tracking restart
                    st   = kModeSeg;
                    continue;
                }
            }
            else
            {
                float3 t_den = sig_t * den - f3(sigc);
                float3 s_den = sig_s * den - f3(sigc);
                float3 n_den = f3(maj) - t_den;
                float  Ps = fabsf(t_den.x * T.x) + fabsf(t_den.y * T.y) + fabsf(t_den.z * T.z);
                float  Pn = fabsf(n_den.x * T.x) + fabsf(n_den.y * T.y) + fabsf(n_den.z * T.z);
                float  c  = Ps + Pn;
                float  e  = u1 * c;
                scatter   = e < Ps;
                float k   = __fdividef(c, maj * (scatter ? Ps : Pn));
                T         = T * ((scatter ? s_den : n_den) * k);
                if (!scatter) continue;
            }
            // ---- scattering event at pos ----
            {
                if (STATS) c_scat++;
                float sr_pre = fmaxf(0.0f, fminf(1.0f, (n - 5) * 0.066666666666666666667f));
                float g      = (1 - sr_pre) * P.g;  // Q4: g of the pre-increment count
                n++;
                float3 ft, fb;
                make_frame(s, ft, fb);
                ph = hg_eval_fast(g, dot3(s, S.sun_dir));
                float r0, r1;
                rng.draw(r0, r1);
                float3 l = hg_sample_local(g, r0, r1);
                pend     = normalize3(ft * l.x + fb * l.y + s * l.z);
                o        = pos;
                float sr = fmaxf(0.0f, fminf(1.0f, (n - 5) * 0.066666666666666666667f));
                dens     = ((1 - sr) + sr * (1 - P.g)) * P.density;
                if ((int)rng.frame > 10 && n > 20)  // K.cu:2183: precomputed sun opacity
                {
                    if (STATS) c_op++;
                    float  tau = (!JULIA && S.have_opacity) ? fetch_opacity(S, pos, false) : 0.0f;
                    float3 a   = f3(__expf(-sig_t.x * dens * tau), __expf(-sig_t.y * dens * tau), __expf(-sig_t.z * dens * tau));
                    L          = L + S.sun_power * (T * ph * a);
                    s          = pend;
                    st         = kModeRay;
                    if (n >= kMaxDepth) goto path_done;
                    continue;
                }
                // shadow walk toward the sun with the LOCAL majorant (Q1), K.cu:2173-2208
                inv = __fdividef(1.0f, max_sig_t * dens * dmax);
                s   = S.sun_dir;  // normalize(sun_dir * 1e10 - pos) up to rounding
                float tn, tf;
                box_slabs(S, o, s, tn, tf);
                dist = 0.0f;
                lim  = (tf > tn && tf >= 1e-3f) ? tf : 0.0f;
                st   = kModeStep | kShadow;
                continue;
            }
        }
    path_done:
        {
            float* a = acc + pslot * 4;
            atomicAdd(a + 0, fmaxf(L.x * P.brightness, 0.0f));  // Q9 clamp per sample (K.cu:2315-2316)
            atomicAdd(a + 1, fmaxf(L.y * P.brightness, 0.0f));
            atomicAdd(a + 2, fmaxf(L.z * P.brightness, 0.0f));
            atomicAdd(a + 3, (float)n);
            st = kModePath;
        }
    }
    if (STATS)
    {
        atomicAdd(d_stats + 0, c_track); atomicAdd(d_stats + 1, c_shadow); atomicAdd(d_stats + 2, c_seg);
        atomicAdd(d_stats + 3, c_op);    atomicAdd(d_stats + 4, c_env);    atomicAdd(d_stats + 5, c_scat);
    }
    __syncthreads();
    {
        uint32_t lane = tid & 31, wrp = tid >> 5;
        uint32_t x = tile_x0 + (wrp & 1) * 8 + (lane & 7), y = tile_y0 + (wrp >> 1) * 4 + (lane >> 3);
        if (x < P.width && y < P.height)
        {
            float4* p = d_sum + (x + (size_t)y * P.width);
            float4  v = *p;
            v.x += acc[tid * 4 + 0]; v.y += acc[tid * 4 + 1]; v.z += acc[tid * 4 + 2]; v.w += acc[tid * 4 + 3];
            *p = v;
        }
    }
}

template <int VT, bool JULIA>
static cudaError_t launch_fast_t(const Scene& S, float4* d_sum, int first_frame, int n_frames, int frame_stride, const vp_param& P,
                                 unsigned long long* d_stats, cudaStream_t stream)
{
    int tiles_x = (P.width + kTileW - 1) / kTileW, tiles_y = (P.height + kTileH - 1) / kTileH;
    if (d_stats)
        k_render_fast<VT, JULIA, true><<<tiles_x * tiles_y, kTilePix, 0, stream>>>(S, d_sum, first_frame, n_frames, frame_stride, P, tiles_x, d_stats);
    else
        k_render_fast<VT, JULIA, false><<<tiles_x * tiles_y, kTilePix, 0, stream>>>(S, d_sum, first_frame, n_frames, frame_stride, P, tiles_x, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_render_fast(const Scene& S, float4* d_sum, int first_frame, int n_frames, int frame_stride, const vp_param& P,
                               unsigned long long* d_stats, cudaStream_t stream)
{
    if (S.julia) return launch_fast_t<kF32, true>(S, d_sum, first_frame, n_frames, frame_stride, P, d_stats, stream);
    if (S.voxel_type == kU8) return launch_fast_t<kU8, false>(S, d_sum, first_frame, n_frames, frame_stride, P, d_stats, stream);
    if (S.voxel_type == kF16) return launch_fast_t<kF16, false>(S, d_sum, first_frame, n_frames, frame_stride, P, d_stats, stream);
    return launch_fast_t<kF32, false>(S, d_sum, first_frame, n_frames, frame_stride, P, d_stats, stream);
}
}  // namespace vp
