// volpath_render_fast.cu -- the B200 production renderer (VP_MODE_FAST), megakernel form.
//
// Same estimator as the reference's __d_render_bounded_decomp (K.cu:1958-2318) in distribution -- weighted
// spectral delta tracking against a local majorant, analog decomposition where the local minimum is
// positive (with the reference's quirks Q1-Q4), reduced-scattering switch after 5
// bounces, sun NEE by shadow walk or opacity table, HG phase sampling -- re-organised for the machine.
// ncu on the reference kernel rebuilt for sm_100 (profiles/r1a_ref_ncu_details.txt): 7.6 of 32 threads active
// per warp, 20 % achieved occupancy: one thread = one pixel = one path per launch leaves the machine waiting
// for the longest path of every warp and of every launch.  Here:
//
//   * PERSISTENT WARPS, GLOBAL WORK POOL.  The grid is sized to the SMs; a path-sample is an item
//     (pixel, frame) of one global index space; each warp claims ranges of 32..256 items (by launch size) with ONE atomic and hands
//     them to lanes as they finish (ballot/popc ranking).  No lane idles before the pool is empty; nothing
//     waits for a launch boundary (all frames of a call are one launch).
//   * WARP-LEVEL EVENT BINNING.  A lane is in one of four states -- new path, segment setup, walk step,
//     scatter event.  Every iteration the warp votes (match_any + redux.max) and runs ONLY the block the most
//     lanes wait for; the others keep their state in registers.  Divergent one-lane excursions through the
//     expensive rare blocks (HG sampling, camera setup) are replaced by batched ones.
//   * one walk-step body serves both walks (primary tracking and the sun shadow walk): draw, advance, fetch,
//     decide -- the density fetches of all stepping lanes are issued together.
//   * the ray is intersected with the box once; the reference's 0.05-step approach march outside the medium
//     (Q5: ~40 of ~59 segments per path, profiles/ref_counters.json) and segments whose local max is exactly
//     zero are skipped without random draws (an exponential walk through vacuum is memoryless).
//   * counter-based Philox2x32-10: key = pixel, counter = (draw index, frame); no carried RNG state.
//   * density comes from the octet store: one table load + one 32/16/8-byte load per trilinear sample.
//   * one vector atomic (red.global.add.v4.f32) per finished path into the float4 sum.
//   * rare work runs where it is batched best, not where it arises: the environment lookup of an escaped path waits
//     for the path block, the sun contribution of a finished shadow walk for the head of the segment block.
//   * the storage layout (flat brick table / rank directory, float / half per-cell tables) is a template parameter
//     of the production variants (LY), so the fetch and the segment loop carry no layout tests.
#include "volpath_fast_common.cuh"
#include "volpath_kernels.h"

namespace vp
{
#ifndef VP_FAST_THREADS
#define VP_FAST_THREADS 128
#endif
constexpr int      kFastThreads  = VP_FAST_THREADS;  // CTA size; the CTA counts below are stated for 128 and scale with it
// Occupancy: the kernel waits on dependent loads (ncu: 7 of 15 stall cycles per issue are long-scoreboard), so warps
// per SM matter.  8 / 9 / 10 CTAs of 128 threads: 820 / 878 / 903 M path-samples/s (1/4-dims cloud); with the cold
// lane state in shared memory (VP_COLD_SMEM) the step block fits 40 registers and 12 CTAs (48 warps) are resident:
// full C2 volume 894 -> 935 M/s (no change on the small cloud).  Two walk steps per vote: +3 %.
#ifndef VP_COLD_SMEM
#define VP_COLD_SMEM 1
#endif
#ifndef VP_CTAS_PER_SM
#define VP_CTAS_PER_SM (VP_COLD_SMEM ? 12 : 10)
#endif
#ifndef VP_STEP_REPS
#define VP_STEP_REPS 2
#endif
#ifndef VP_STEP_MAXREPS
#define VP_STEP_MAXREPS VP_STEP_REPS
#endif
#ifndef VP_STEP_MINLANES
#define VP_STEP_MINLANES 16
#endif
#ifndef VP_INLINE_SEG
#define VP_INLINE_SEG 0
#endif
#ifndef VP_STEP_FLAT
#define VP_STEP_FLAT 1  // gray media: branch-free step decision (same arithmetic, same samples): 1088 -> 1100 M/s at full C2; with the
                        // brick-skip variants it loses 0.9 % (C4), so those keep the branches
#endif
#ifndef VP_STEP_FLAT_CHROMA
#define VP_STEP_FLAT_CHROMA 0
#endif
#ifndef VP_SUN_NO_SLAB
#define VP_SUN_NO_SLAB 0  // experiment: sun shadow walks ended by the sun-clear distance alone (+0.7 % at full C2) -- NOT exact when medium
                          // touches the box wall: clamp addressing extends the border voxels half a voxel beyond the box
                          // (caught by test_tiny_full_box_grid_sun_clear_clip_is_conservative); off
#endif
#ifndef VP_SEG_ITERS
#define VP_SEG_ITERS 0  // 0: the segment block loops over vacuum jumps until medium or exit
#endif
#ifndef VP_SMEM_TOP
#define VP_SMEM_TOP 0  // stage the top level of the bound grid (Scene::top_jump, <= 12 KB) in shared memory for the vacuum jumps of the segment block
#endif
#ifndef VP_CHROMA_CTAS
#define VP_CHROMA_CTAS 12  // chromatic media: ptxas fits the 3-channel step into 40 registers without spills; 10 / 11 / 12 CTAs: 894 / 917 / 935 M/s (preset 8, full C2)
#endif

constexpr int      kFastCtasPerSm = VP_CTAS_PER_SM;
// chromatic media carry a 3-channel throughput: the same 12 CTAs since the launch bound makes ptxas fit 40 registers (VP_CHROMA_CTAS); the MIS variant 8
__host__ __device__ constexpr int fast_ctas_per_sm(bool gray, bool mis)
{
    return (mis ? 8 : (gray ? kFastCtasPerSm : (kFastCtasPerSm > VP_CHROMA_CTAS ? VP_CHROMA_CTAS : kFastCtasPerSm))) * (128 / kFastThreads);
}
constexpr uint32_t kFull         = 0xffffffffu;
#ifndef VP_CLAIM
#define VP_CLAIM 64  // 512 / 256 / 128 / 96 / 64 / 32 items per claim: 1058 / 1093 / 1116 / 1120 / 1120 / 1115 M/s (full C2, 64-frame launches)
#endif
constexpr uint32_t kClaim        = VP_CLAIM;  // items per warp-level claim (large launches); small launches claim less, see launch_fast_t
// vote weights of the four blocks {-, path, scatter, segment, step}: a block runs when lanes x weight is largest, so a
// lower weight makes an expensive block wait for more lanes (tuned on B200, profiles/)
#ifndef VP_W_PATH
#define VP_W_PATH 6
#define VP_W_SCAT 6
#define VP_W_SEG 5
#define VP_W_STEP 4
#endif
constexpr uint32_t kPickWeights = (VP_W_PATH << 4) | (VP_W_SCAT << 8) | (VP_W_SEG << 12) | (VP_W_STEP << 16);  // 4 bits per mode

enum : uint32_t
{
    kModeIdle  = 0,  // pool exhausted
    kModePath  = 1,  // needs a new (pixel, frame) item
    kModeScat  = 2,  // scattering event pending at o (direction s = incoming)
    kModeSeg   = 3,  // find the next segment with medium and set its majorants
    kModeStep  = 4,  // walking
    kModeMask  = 7,
    kShadow    = 8,    // walking toward the sun (else: primary tracking)
    kLimIsCtrl = 16,   // `lim` is the control-component collision distance (else: segment end)
    kNeedRay   = 32,   // (o, s) changed: intersect the box before the next segment
    kKillX = 64, kKillY = 128, kKillZ = 256,
    kEscaped   = 512,  // path left the medium: environment lookup + accumulate pending (done in the path block)
    // env-map importance sampling variant (MIS = true) only: the scatter block runs in three stages
    kMisWalk   = 1024,  // the shadow walk in progress is the one toward the MIS direction (else: toward the sun)
    kScatB     = 2048,  // sun NEE done: choose the MIS direction and start its shadow walk
    kScatC     = 4096,  // both NEE walks done: sample the scattered direction
    kShadowDone = 8192,  // a sun shadow walk just ended: its contribution is added at the head of the segment block
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

// "cold" lane state -- touched at scatter / segment / path events only, never in a walk step -- can live in shared
// memory (one word per field per thread, field-major: conflict-free), which takes it out of the register budget of the
// step block.  Cold3 / ColdF read and write like a float3 / float.
struct Cold3
{
    float* p;
    __device__ __forceinline__ operator float3() const { return f3(p[0], p[kFastThreads], p[2 * kFastThreads]); }
    __device__ __forceinline__ void operator=(float3 v) const
    {
        p[0] = v.x; p[kFastThreads] = v.y; p[2 * kFastThreads] = v.z;
    }
};
struct ColdF
{
    float* p;
    __device__ __forceinline__ operator float() const { return *p; }
    __device__ __forceinline__ void operator=(float v) const { *p = v; }
};

template <int VT, bool JULIA, bool GRAY, bool MIS, bool STATS, int LYX>
__global__ void __launch_bounds__(kFastThreads, fast_ctas_per_sm(GRAY, MIS)) k_render_fast(const __grid_constant__ Scene S, float4* __restrict__ d_sum,
                                                                  int first_frame, int n_frames, int frame_stride,
                                                                  const __grid_constant__ vp_param P,
                                                                  unsigned long long* __restrict__ d_work,
                                                                  unsigned long long* __restrict__ d_stats, uint32_t claim, int skip_rt)
{
    // LYX = layout (0 generic, 1 small volume, 2 large volume) + 2 when empty-brick skipping is compiled in (3, 4);
    // the generic kernel takes the same decision from its argument
    constexpr int  LY    = LYX >= 3 ? LYX - 2 : LYX;
    constexpr bool kSkip = LYX >= 3;
    const bool     skip_on = kSkip || (LYX == 0 && skip_rt != 0);
    const uint32_t lane    = threadIdx.x & 31;
    const uint32_t tiles_x = (P.width + 7) >> 3, tiles_y = (P.height + 3) >> 2;
    const unsigned long long n_items = (unsigned long long)tiles_x * tiles_y * 32ull * (unsigned long long)n_frames;

    const float3 sig_t = f3(P.sigma_t.x, P.sigma_t.y, P.sigma_t.z);
    const float3 sig_s = sig_t * f3(P.albedo.x, P.albedo.y, P.albedo.z);
    const float  max_sig_t = max_of(sig_t), min_sig_t = min_of(sig_t);

    // warp-uniform claim window
    unsigned long long w_next = 0, w_end = 0;
    // lane state
#if VP_COLD_SMEM
    __shared__ float cold[12 * kFastThreads];
    float*           cb = cold + threadIdx.x;
    const Cold3      L{cb}, pend{cb + 3 * kFastThreads}, C{cb + 6 * kFastThreads};
    const ColdF      t_exit{cb + 9 * kFastThreads}, ph{cb + 10 * kFastThreads}, dmax{cb + 11 * kFastThreads};
    float3           o = f3(0.f), s = f3(0.f), T = f3(1.f);
    float            dist = 0.f, lim = 0.f, inv = 0.f, dens = 0.f, maj = 0.f, sigc = 0.f;
#else
    float3   o = f3(0.f), s = f3(0.f), pend = f3(0.f), T = f3(1.f), L = f3(0.f);
    float3   C = f3(0.f);  // MIS only: radiance the pending env-direction walk will add if it survives
    float    dist = 0.f, lim = 0.f, inv = 0.f, dens = 0.f, maj = 0.f, sigc = 0.f, t_exit = 0.f, ph = 0.f, dmax = 0.f;
#endif
#if VP_SMEM_TOP
    __shared__ uint16_t top_sm[kTopCellsMax];
    const bool use_top = !JULIA && S.top_jump != nullptr;
    if (use_top)
    {
        const int nt = S.ntx * S.nty * S.ntz;
        for (int i = threadIdx.x; i < nt; i += kFastThreads) top_sm[i] = __ldg(S.top_jump + i);
    }
    __syncthreads();
#endif
    int      n = 0;
    uint32_t st = kModePath, pix = 0;
    Philox   rng{0, 0, 0};
    unsigned long long c_track = 0, c_shadow = 0, c_seg = 0, c_op = 0, c_env = 0, c_scat = 0;
    unsigned long long c_blk[4] = {0, 0, 0, 0}, c_act[4] = {0, 0, 0, 0}, c_zero_s = 0, c_zero_t = 0;

    for (;;)
    {
        // ---- vote: run the block the most lanes wait for (ties: step > segment > scatter > path) ----
        const uint32_t mode = st & kModeMask;
        const uint32_t same = __match_any_sync(kFull, mode);
        const uint32_t key  = ((__popc(same) * ((kPickWeights >> (mode * 4)) & 15u)) << 3) | mode;  // idle: 0
        const uint32_t pick = __reduce_max_sync(kFull, key) & 7u;
        if (pick == kModeIdle) break;
        if (STATS)
        {
            // binning efficiency: per block type, warp-level executions (lane 0) and active lanes
            if (lane == 0) c_blk[pick - 1]++;
            if (mode == pick) c_act[pick - 1]++;
        }
        if (pick == kModePath)
        {
            const uint32_t needy = __ballot_sync(kFull, mode == kModePath);
            const uint32_t cnt = __popc(needy), rank = __popc(needy & ((1u << lane) - 1u));
            const uint32_t avail = (uint32_t)(w_end - w_next);
            unsigned long long item;
            if (cnt > avail)
            {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(d_work, (unsigned long long)claim);
                base   = __shfl_sync(kFull, base, 0);
                item   = rank < avail ? w_next + rank : base + (rank - avail);
                w_next = base + (cnt - avail);
                w_end  = base + claim;
            }
            else
            {
                item = w_next + rank;
                w_next += cnt;
            }
            if (mode == kModePath)
            {
                if (st & kEscaped)
                {
                    // finish the previous path here, where many lanes are in the same situation: environment /
                    // sun disk (acosf + atanf + one texel), then one vector atomic
                    if (STATS) c_env++;
                    // K.cu:2026-2030: with env-map sampling the environment is picked up by escaping paths at depth 0 only
                    if (!MIS || n == 0) L = float3(L) + background(S, s, n) * (GRAY ? f3(T.x) : T);
                    accumulate(d_sum, pix, float3(L), n, P.brightness);
                    st = kModePath;
                }
                if (item >= n_items)
                    st = kModeIdle;
                else
                {
                    uint32_t x, y, f;
                    item_to_sample(item, (uint32_t)n_frames, tiles_x, x, y, f);
                    if (x < P.width && y < P.height)
                    {
                        pix       = y * P.width + x;
                        rng.key   = pix;
                        rng.frame = (uint32_t)(first_frame + (int)f * frame_stride);
                        rng.ctr   = 0;
                        camera_ray_fast(S, x, y, P.width, P.height, o, s);
                        T  = f3(1.f);
                        L  = f3(0.f);
                        n  = 0;
                        st = kModeSeg | kNeedRay;
                    }
                }
            }
        }
        else if (pick == kModeSeg)
        {
            if (!MIS && (st & kShadowDone))
            {
                // a sun shadow walk ended (kill flags in st): add the sun's contribution, turn into the sampled direction
                if (GRAY)
                {
                    if (!(st & kKillX)) L = float3(L) + S.sun_power * (T.x * float(ph));
                }
                else
                {
                    float3 a = f3((st & kKillX) ? 0.f : 1.f, (st & kKillY) ? 0.f : 1.f, (st & kKillZ) ? 0.f : 1.f);
                    L        = float3(L) + S.sun_power * (T * float(ph) * a);
                }
                s  = float3(pend);
                st = kModeSeg | kNeedRay;
                if (n >= kMaxDepth)
                {
                    accumulate(d_sum, pix, float3(L), n, P.brightness);
                    st = kModePath;
                }
            }
            if ((st & kModeMask) == kModeSeg)
            {
                if (st & kNeedRay)
                {
                    // one slab test per ray (the reference repeats it every 0.05 step, K.cu:1626-1661)
                    float tn, tf;
                    box_slabs_fast(S, o, s, tn, tf);
                    dist   = fmaxf(tn, 0.0f);
                    t_exit = (tf > tn && tf >= 1e-3f) ? tf : -1.0f;
                }
                bool found = false;
                const float t_ex = t_exit;
#if VP_SEG_ITERS
                int iters = 0;  // at most VP_SEG_ITERS vacuum jumps per vote: lanes that need more stay in the segment state
#endif
                while (dist < t_ex)
                {
#if VP_SEG_ITERS
                    if (iters++ >= VP_SEG_ITERS)
                    {
                        found = true;
                        st    = kModeSeg;
                        break;
                    }
#endif
                    if (STATS) c_seg++;
#if VP_SMEM_TOP
                    if (use_top)
                    {
                        // shared-memory top level: a block of bound cells that is all vacuum answers without a global load
                        const float3 q  = o + s * dist;
                        const int    sh = S.top_log2;
                        const int    ti = clampi(__float2int_rd(fmaf(q.x, S.cs_scale.x, S.cs_off.x)), 0, S.ncx - 1) >> sh;
                        const int    tj = clampi(__float2int_rd(fmaf(q.y, S.cs_scale.y, S.cs_off.y)), 0, S.ncy - 1) >> sh;
                        const int    tk = clampi(__float2int_rd(fmaf(q.z, S.cs_scale.z, S.cs_off.z)), 0, S.ncz - 1) >> sh;
                        const float  J  = __half2float(__ushort_as_half(top_sm[(tk * S.nty + tj) * S.ntx + ti]));
                        if (J > 0.0f)
                        {
                            dist = fminf(dist + fmaxf(kSearchRadius, J), t_ex);
                            continue;
                        }
                    }
#endif
                    float  seg_end = JULIA ? t_ex : fminf(dist + kSearchRadius, t_ex);
                    float2 bnd     = JULIA ? make_float2(1.0f, 0.0f) : bounds_at<LY>(S, o + s * dist);
                    if (bnd.x <= 0.0f)
                    {
                        // no medium within reach: the walk passes with probability 1; -bnd.x is how far it may go
                        // in any direction without leaving vacuum (breadth-first distance over the bound cells)
                        dist = fminf(dist + fmaxf(kSearchRadius, -bnd.x), t_ex);
                        continue;
                    }
                    const float dmx = fmaxf(1e-4f, bnd.x);
                    dmax            = dmx;
                    // reduced scattering after 5 bounces (K.cu:2039-2044)
                    float sr = fmaxf(0.0f, fminf(1.0f, (n - 5) * 0.066666666666666666667f));
                    dens     = ((1 - sr) + sr * (1 - P.g)) * P.density;
                    maj      = max_sig_t * dens * dmx;
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
                    st = kModePath | kEscaped;  // escaped (or never hit the box)
                }
            }
        }
        else if (pick == kModeStep)
        {
            // ---- one step of whichever walk this lane is on ----
            auto step = [&](float u0, float u1)
            {
                dist += -__logf(u0) * inv;
                const bool past = dist >= lim;
                float3     pos  = o + s * (past ? lim : dist);
                float      den  = 0.0f;
                if (!past)
                {
                    if (!JULIA && GRAY && (kSkip || LYX == 0) && skip_on)
                    {
                        float skip;
                        den = density_at_skip<VT, JULIA, LY>(S, pos, s, skip) * dens;
                        // empty brick: move to its exit without drawing (exact: see density_at_skip); not in a decomposition
                        // segment, whose control component makes a zero-density step a real event
                        if (sigc == 0.0f || (st & kShadow)) dist += skip;
                    }
                    else
                        den = density_at<VT, JULIA, LY>(S, pos) * dens;
                    if (STATS)
                    {
                        if (st & kShadow) c_shadow++; else c_track++;
                        if (den == 0.0f) { if (st & kShadow) c_zero_s++; else c_zero_t++; }
                    }
                }
#if VP_STEP_FLAT
                if (GRAY && !MIS && (LYX == 1 || LYX == 2))  // the two production layouts without brick skipping: +1.0 % at full C2
                {
                    // branch-free decision for gray media: the three outcomes (shadow step / segment end / tracking event)
                    // are evaluated by every stepping lane and selected, instead of three divergent branches in a row
                    const bool  shadow = (st & kShadow) != 0, ctrl = (st & kLimIsCtrl) != 0;
                    const float t_den = sig_t.x * den - sigc, s_den = sig_s.x * den - sigc, n_den = maj - t_den;
                    const float at = fabsf(t_den), an = fabsf(n_den), c = at + an;
                    const bool  hit = u1 * c < at;
                    const float w   = (hit ? s_den : n_den) * __fdividef(c, maj * (hit ? at : an));
                    const bool  track = !shadow && !past;
                    const bool  kill  = shadow && !past && (u1 < sig_t.x * (den * inv));
                    const bool  to_scat = (track && hit) || (!shadow && past && ctrl);
                    const bool  to_seg  = !shadow && past && !ctrl;
                    const bool  sh_done = shadow && (past || kill);
                    T.x  = track ? T.x * w : T.x;
                    o    = f3(to_scat ? pos.x : o.x, to_scat ? pos.y : o.y, to_scat ? pos.z : o.z);
                    dist = to_seg ? lim : dist;
                    uint32_t ns = to_scat ? (uint32_t)kModeScat : st;
                    ns = to_seg ? (uint32_t)kModeSeg : ns;
                    ns = sh_done ? (kModeSeg | kNeedRay | kShadowDone | (kill ? (kKillX | kKillY | kKillZ) : 0u)) : ns;
                    st = ns;
                }
                else
#endif
#if VP_STEP_FLAT_CHROMA
                if (!GRAY && !MIS && (LYX == 1 || LYX == 2))
                {
                    // the same for chromatic media: per-channel kill flags accumulate over the steps of a shadow walk
                    const bool   shadow = (st & kShadow) != 0, ctrl = (st & kLimIsCtrl) != 0;
                    const float3 t_den = sig_t * den - f3(sigc), s_den = sig_s * den - f3(sigc), n_den = f3(maj) - t_den;
                    const float  Ps = fabsf(t_den.x * T.x) + fabsf(t_den.y * T.y) + fabsf(t_den.z * T.z);
                    const float  Pn = fabsf(n_den.x * T.x) + fabsf(n_den.y * T.y) + fabsf(n_den.z * T.z);
                    const float  c  = Ps + Pn;
                    const bool   hit = u1 * c < Ps;
                    const float  k   = __fdividef(c, maj * (hit ? Ps : Pn));
                    const bool   track = !shadow && !past;
                    const float  q = den * inv;
                    uint32_t     kills = st & (kKillX | kKillY | kKillZ);
                    if (shadow && !past)
                        kills |= (u1 < sig_t.x * q ? kKillX : 0u) | (u1 < sig_t.y * q ? kKillY : 0u) | (u1 < sig_t.z * q ? kKillZ : 0u);
                    const bool to_scat = (track && hit) || (!shadow && past && ctrl);
                    const bool to_seg  = !shadow && past && !ctrl;
                    const bool sh_done = shadow && (past || kills == (kKillX | kKillY | kKillZ));
                    if (track) T = T * ((hit ? s_den : n_den) * k);
                    o    = f3(to_scat ? pos.x : o.x, to_scat ? pos.y : o.y, to_scat ? pos.z : o.z);
                    dist = to_seg ? lim : dist;
                    uint32_t ns = shadow ? (st | kills) : st;
                    ns = to_scat ? (uint32_t)kModeScat : ns;
                    ns = to_seg ? (uint32_t)kModeSeg : ns;
                    ns = sh_done ? (kModeSeg | kNeedRay | kShadowDone | kills) : ns;
                    st = ns;
                }
                else
#endif
                if (st & kShadow)
                {
                    if (!past)
                    {
                        // Tr_spectral (K.cu:782-806): per-channel kill flags on one shared walk
                        float q = den * inv;
                        if (GRAY)
                        {
                            if (u1 < sig_t.x * q) st |= kKillX | kKillY | kKillZ;
                        }
                        else
                        {
                            if (u1 < sig_t.x * q) st |= kKillX;
                            if (u1 < sig_t.y * q) st |= kKillY;
                            if (u1 < sig_t.z * q) st |= kKillZ;
                        }
                    }
                    if (past || (st & (kKillX | kKillY | kKillZ)) == (kKillX | kKillY | kKillZ))
                    {
                        if (MIS)
                        {
                            float3 a = f3((st & kKillX) ? 0.f : 1.f, (st & kKillY) ? 0.f : 1.f, (st & kKillZ) ? 0.f : 1.f);
                            // sun walk done -> MIS stage; MIS walk done -> direction sampling stage
                            if (st & kMisWalk)
                            {
                                L  = float3(L) + float3(C) * a;
                                st = kModeScat | kScatC;
                            }
                            else
                            {
                                L  = float3(L) + S.sun_power * ((GRAY ? f3(T.x) : T) * float(ph) * a);
                                st = kModeScat | kScatB;
                            }
                        }
                        else
                        {
                            // the sun contribution and the pending direction are taken up at the head of the segment
                            // block, which this lane needs next anyway: there the few lanes that finish a shadow
                            // walk in any one step (2 of 32, ncu) are batched with the other lanes that start a ray
                            st = kModeSeg | kNeedRay | kShadowDone | (st & (kKillX | kKillY | kKillZ));
                        }
                    }
                }
                else if (past)
                {
                    if (st & kLimIsCtrl)
                    {
                        o  = pos;  // control-component collision: no weight (Q2)
                        st = kModeScat;
                    }
                    else
                    {
                        dist = lim;  // crossed the segment: tracking restart
                        st   = kModeSeg;
#if VP_INLINE_SEG
                        // The common continuation -- the next 0.05 segment has medium and no control component -- is set
                        // up right here, so the lane keeps walking instead of waiting for a segment-block vote; exits,
                        // vacuum (jump distances) and decomposition segments still go through the segment block.
                        if (!JULIA && !STATS)
                        {
                            const float t_ex = t_exit;
                            if (dist < t_ex)
                            {
                                const float2 bnd = bounds_at<LY>(S, pos);
                                if (bnd.x > 0.0f && !(bnd.y > 0.0f))
                                {
                                    const float dmx = fmaxf(1e-4f, bnd.x);
                                    dmax = dmx;
                                    maj  = max_sig_t * dens * dmx;  // dens already belongs to the current scatter count
                                    lim  = fminf(dist + kSearchRadius, t_ex);
                                    sigc = 0.0f;
                                    inv  = __fdividef(1.0f, maj);
                                    st   = kModeStep;
                                }
                            }
                        }
#endif
                    }
                }
                else if (GRAY)
                {
                    // gray medium (sigma_t and albedo equal in r, g, b): the throughput is a scalar and sum|T|
                    // cancels out of Ps / (Ps + Pn) -- same probabilities and weights as the spectral form
                    float t_den = sig_t.x * den - sigc;
                    float s_den = sig_s.x * den - sigc;
                    float n_den = maj - t_den;
                    float at = fabsf(t_den), an = fabsf(n_den), c = at + an;
                    bool  hit = u1 * c < at;
                    float k   = __fdividef(c, maj * (hit ? at : an));
                    T.x *= (hit ? s_den : n_den) * k;
                    if (hit)
                    {
                        o  = pos;
                        st = kModeScat;
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
                    bool   hit = e < Ps;
                    float  k   = __fdividef(c, maj * (hit ? Ps : Pn));
                    T          = T * ((hit ? s_den : n_den) * k);
                    if (hit)
                    {
                        o  = pos;
                        st = kModeScat;
                    }
                }
            };
#pragma unroll 1
            for (int rep = 0; rep < VP_STEP_MAXREPS; rep++)
            {
                // after the guaranteed VP_STEP_REPS steps, keep stepping without a new vote while enough lanes still walk
                if (rep >= VP_STEP_REPS && __popc(__ballot_sync(kFull, (st & kModeMask) == kModeStep)) < VP_STEP_MINLANES) break;
                if ((st & kModeMask) == kModeStep)
                {
                    float u0, u1;
                    rng.draw(u0, u1);
                    step(u0, u1);
                }
            }
        }
        else  // kModeScat
        {
            if (mode == kModeScat && MIS && (st & (kScatB | kScatC)))
            {
                // ---- env-map sampling variant, stages B and C of a scattering event (incoming direction in pend) ----
                float  sr_pre = fmaxf(0.0f, fminf(1.0f, (n - 1 - 5) * 0.066666666666666666667f));
                float  g      = (1 - sr_pre) * P.g;
                float3 ft, fb;
                const float3 din = pend;
                make_frame_fast(din, ft, fb);
                if (st & kScatB)
                {
                    // one-sample MIS between phase-function and env-map sampling (K.cu:2220-2297)
                    float rsel, u, v, unused;
                    rng.draw(rsel, u);
                    rng.draw(v, unused);
                    float3 dir, envc;
                    bool   ok = true;
                    if (rsel < 0.5f)
                    {
                        float3 ls = hg_sample_local_fast(g, u, v);
                        dir       = ft * ls.x + fb * ls.y + din * ls.z;
                        envc      = eval_envmap(S, dir);
                        float pdf_brdf = hg_eval_fast(g, dot3(din, dir));
                        float pdf_env  = pdf_envmap(S, envc);
                        float weight   = __fdividef(pdf_brdf * 0.5f, pdf_brdf * 0.5f + pdf_env * 0.5f) * 2.0f;
                        C              = envc * ((GRAY ? f3(T.x) : T) * weight);
                    }
                    else
                    {
                        float pdf_env = sample_envmap(S, u, v, envc);
                        ok            = pdf_env > 0.0f;  // (the reference `continue`s here, K.cu:2266: probability ~2^-23)
                        dir           = uv_to_dir(u, v);
                        float pb      = hg_eval_fast(g, dot3(din, dir));
                        float weight  = __fdividef(pdf_env * 0.5f, pdf_env * 0.5f + pb * 0.5f) * 2.0f;
                        C             = envc * ((GRAY ? f3(T.x) : T) * (__fdividef(pb, pdf_env) * weight));
                    }
                    if (ok)
                    {
                        s = normalize3(dir);
                        float tn, tf;
                        box_slabs_fast(S, o, s, tn, tf);
                        dist = 0.0f;
                        lim  = (tf > tn && tf >= 1e-3f) ? tf : 0.0f;
                        inv  = __fdividef(1.0f, max_sig_t * dens * float(dmax));
                        st   = kModeStep | kShadow | kMisWalk;
                    }
                    else
                        st = kModeScat | kScatC;
                }
                else
                {
                    float r0, r1;
                    rng.draw(r0, r1);
                    float3 l = hg_sample_local_fast(g, r0, r1);
                    s        = normalize3(ft * l.x + fb * l.y + din * l.z);
                    st       = kModeSeg | kNeedRay;
                    if (n >= kMaxDepth)
                    {
                        accumulate(d_sum, pix, float3(L), n, P.brightness);
                        st = kModePath;
                    }
                }
            }
            else if (mode == kModeScat)
            {
                // ---- scattering event at o, incoming direction s ----
                if (STATS) c_scat++;
                float sr_pre = fmaxf(0.0f, fminf(1.0f, (n - 5) * 0.066666666666666666667f));
                float g      = (1 - sr_pre) * P.g;  // Q4: g of the pre-increment count
                n++;
                ph = hg_eval_fast(g, dot3(s, S.sun_dir));
                if (MIS)
                    pend = s;  // keep the incoming direction: the new one is sampled after both NEE walks (stage C)
                else
                {
                    float3 ft, fb;
                    make_frame_fast(s, ft, fb);
                    float r0, r1;
                    rng.draw(r0, r1);
                    float3 l = hg_sample_local_fast(g, r0, r1);
                    pend     = normalize3(ft * l.x + fb * l.y + s * l.z);
                }
                float sr = fmaxf(0.0f, fminf(1.0f, (n - 5) * 0.066666666666666666667f));
                dens     = ((1 - sr) + sr * (1 - P.g)) * P.density;
                if ((int)rng.frame > 10 && n > 20)  // K.cu:2183: precomputed sun opacity
                {
                    if (STATS) c_op++;
                    float  tau = (!JULIA && S.opacity_oct) ? opacity_at<LY>(S, o) : 0.0f;
                    float3 a   = GRAY ? f3(__expf(-sig_t.x * dens * tau))
                                      : f3(__expf(-sig_t.x * dens * tau), __expf(-sig_t.y * dens * tau), __expf(-sig_t.z * dens * tau));
                    L          = float3(L) + S.sun_power * ((GRAY ? f3(T.x) : T) * float(ph) * a);
                    if (MIS)
                        st = kModeScat | kScatB;
                    else
                    {
                        s  = float3(pend);
                        st = kModeSeg | kNeedRay;
                        if (n >= kMaxDepth)
                        {
                            accumulate(d_sum, pix, float3(L), n, P.brightness);
                            st = kModePath;
                        }
                    }
                }
                else
                {
                    // shadow walk toward the sun with the LOCAL majorant (Q1), K.cu:2173-2208
                    inv = __fdividef(1.0f, max_sig_t * dens * float(dmax));
                    s   = S.sun_dir;  // normalize(sun_dir * 1e10 - pos) up to rounding
                    dist = 0.0f;
#if VP_SUN_NO_SLAB
                    // With a sun-clear table the slab test is redundant: the table ends the walk where the last medium
                    // toward the sun ends, and a step beyond the box reads density 0 (range test of the fetch) -- a no-op.
                    // (The rule must not depend on the layout variant: the number of draws a walk consumes follows from it.)
                    if (!JULIA && (LY != 0 || S.sun_clear))
                        lim = sun_clear_at<LY>(S, o);
                    else
#endif
                    {
                        float tn, tf;
                        box_slabs_inv(S, o, S.sun_inv, tn, tf);
                        lim = (tf > tn && tf >= 1e-3f) ? tf : 0.0f;
                        // exact vacuum clip: beyond sun_clear no medium is left on the way to the sun
                        if (!JULIA && (LY != 0 || S.sun_clear)) lim = fminf(lim, sun_clear_at<LY>(S, o));
                    }
                    st   = kModeStep | kShadow;
                }
            }
        }
    }
    if (STATS)
    {
        atomicAdd(d_stats + 0, c_track); atomicAdd(d_stats + 1, c_shadow); atomicAdd(d_stats + 2, c_seg);
        atomicAdd(d_stats + 3, c_op);    atomicAdd(d_stats + 4, c_env);    atomicAdd(d_stats + 5, c_scat);
        atomicAdd(d_stats + 6, c_zero_t);
        atomicAdd(d_stats + 7, c_zero_s);
        for (int i = 0; i < 4; i++)
        {
            if (lane == 0) atomicAdd(d_stats + 8 + i, c_blk[i]);
            atomicAdd(d_stats + 12 + i, c_act[i]);
        }
    }
}

template <int VT, bool JULIA, bool GRAY, bool MIS>
static cudaError_t launch_fast_t(const Scene& S, float4* d_sum, int first_frame, int n_frames, int frame_stride, const vp_param& P,
                                 unsigned long long* d_work, unsigned long long* d_stats, int num_sms, cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(d_work, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    // persistent grid: kFastCtasPerSm CTAs of 128 threads per SM, never more warps than claims
    unsigned long long items = (unsigned long long)((P.width + 7) / 8) * ((P.height + 3) / 4) * 32ull * n_frames;
    // claim size: 256 items per atomic when every warp will come back many times, down to 32 (one per lane) for small
    // launches (e.g. one frame per launch, the reference host's own pattern) so that the pool keeps balancing the load
    const unsigned long long warps_total = (unsigned long long)num_sms * fast_ctas_per_sm(GRAY, MIS) * (kFastThreads / 32);
    unsigned long long       per_warp    = items / (warps_total * 16);
    uint32_t                 claim       = (uint32_t)(per_warp < 32 ? 32 : (per_warp > kClaim ? kClaim : (per_warp / 32) * 32));
    unsigned long long want  = (items + claim - 1) / claim;                      // warps that can get a claim
    unsigned long long ctas  = (want + kFastThreads / 32 - 1) / (kFastThreads / 32);
    unsigned int       grid  = (unsigned int)(ctas < (unsigned long long)num_sms * fast_ctas_per_sm(GRAY, MIS) ? ctas : (unsigned long long)num_sms * fast_ctas_per_sm(GRAY, MIS));
    if (grid < 1) grid = 1;
    // layout known at compile time for the two production cases (see brick_slot); everything else tests the scene
    int ly = 0;
    if (!JULIA && !MIS && !d_stats && S.linear && S.sun_clear)
    {
        if (S.brick_table && !S.bounds_half && !S.sun_clear_half) ly = 1;
        if (!S.brick_table && S.bounds_half && S.sun_clear_half) ly = 2;
    }
    // empty-brick skipping: one rule for every production kernel (use_brick_skip), compiled in for the two layouts
    const int skip = use_brick_skip(S, P.density, fmaxf(P.sigma_t.x, fmaxf(P.sigma_t.y, P.sigma_t.z)), GRAY) ? 1 : 0;
#define VP_LAUNCH(JJ, MM, SS, LL) \
    k_render_fast<VT, JJ, GRAY, MM, SS, LL><<<grid, kFastThreads, 0, stream>>>(S, d_sum, first_frame, n_frames, frame_stride, P, d_work, SS ? d_stats : nullptr, claim, skip)
    if (d_stats)
        VP_LAUNCH(JULIA, MIS, true, 0);
    else if (!JULIA && !MIS && ly == 1)
    {
        if (GRAY && skip) VP_LAUNCH(false, false, false, 3); else VP_LAUNCH(false, false, false, 1);
    }
    else if (!JULIA && !MIS && ly == 2)
    {
        if (GRAY && skip) VP_LAUNCH(false, false, false, 4); else VP_LAUNCH(false, false, false, 2);
    }
    else
        VP_LAUNCH(JULIA, MIS, false, 0);
#undef VP_LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_render_fast(const Scene& S, float4* d_sum, int first_frame, int n_frames, int frame_stride, const vp_param& P,
                               unsigned long long* d_work, unsigned long long* d_stats, int num_sms, cudaStream_t stream)
{
    const bool gray = P.sigma_t.x == P.sigma_t.y && P.sigma_t.y == P.sigma_t.z && P.albedo.x == P.albedo.y && P.albedo.y == P.albedo.z;
#define VP_FAST(VT, J)                                                                                                                  \
    do {                                                                                                                                \
        if (S.env_mis)                                                                                                                  \
            return gray ? launch_fast_t<VT, J, true, true>(S, d_sum, first_frame, n_frames, frame_stride, P, d_work, d_stats, num_sms, stream) \
                        : launch_fast_t<VT, J, false, true>(S, d_sum, first_frame, n_frames, frame_stride, P, d_work, d_stats, num_sms, stream); \
        return gray ? launch_fast_t<VT, J, true, false>(S, d_sum, first_frame, n_frames, frame_stride, P, d_work, d_stats, num_sms, stream)     \
                    : launch_fast_t<VT, J, false, false>(S, d_sum, first_frame, n_frames, frame_stride, P, d_work, d_stats, num_sms, stream);   \
    } while (0)
    if (S.julia) VP_FAST(kF32, true);
    if (S.voxel_type == kU8) VP_FAST(kU8, false);
    if (S.voxel_type == kF16) VP_FAST(kF16, false);
    VP_FAST(kF32, false);
#undef VP_FAST
}
}  // namespace vp
