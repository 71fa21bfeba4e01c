// volpath_render_wave.cu -- VP_MODE_WAVE: the wavefront form of the production renderer, at warp scope.
//
// Same estimator, same Philox streams and therefore THE SAME SAMPLES as the megakernel form
// (volpath_render_fast.cu): a (pixel, frame) item produces the same path in both.  What differs is where the ray
// states live and how lanes are filled:
//
//   * every warp owns a pool of kPool = 64 ray states as structure-of-arrays in SHARED memory (29 words per
//     state, field-major, so the 32 lanes of a batch read one field of 32 states in one LDS);
//   * each iteration the warp picks the event type with the most waiting states (counts kept warp-uniform),
//     COMPACTS up to 32 of them into a batch with ballot/popc ranks, loads only the fields that event needs,
//     runs the event for the whole batch, stores only the fields it changed, and re-files the states by their
//     new event type.  States are sorted by event type, not by lane, so the batch is full as long as the pool
//     holds 32 states of one type: measured lanes per executed block rise from ~16 (megakernel binning, 32
//     states per warp) towards 28+ (profiles/), at the price of ~25 shared-memory operations per event.
//   * the work pool (one atomic per 256 items), vacuum jumps, sun-clear clip, octet fetch and the vector-atomic
//     accumulation are those of the megakernel.
#include "volpath_fast_common.cuh"
#include "volpath_kernels.h"

namespace vp
{
namespace wave
{
constexpr int      kThreads = 128;
constexpr int      kWarps   = kThreads / 32;
constexpr int      kPool    = 64;  // ray states per warp
constexpr int      kRows    = kPool / 32;
constexpr uint32_t kFull    = 0xffffffffu;
constexpr uint32_t kClaim   = 256;

enum : uint32_t
{
    kModeIdle = 0, kModePath = 1, kModeScat = 2, kModeSeg = 3, kModeStep = 4, kModeMask = 7,
    kShadow = 8, kLimIsCtrl = 16, kNeedRay = 32, kKillX = 64, kKillY = 128, kKillZ = 256, kEscaped = 512,
    // env-map importance sampling variant (MIS): the same three-stage scatter event as the megakernel
    kMisWalk = 1024, kScatB = 2048, kScatC = 4096,
};

// state fields (one float/uint word each), field-major in the warp's pool
enum
{
    F_OX, F_OY, F_OZ, F_SX, F_SY, F_SZ, F_PX, F_PY, F_PZ, F_TX, F_TY, F_TZ, F_LX, F_LY, F_LZ,
    F_DIST, F_LIM, F_INV, F_DENS, F_MAJ, F_SIGC, F_TEXIT, F_PH, F_DMAX, F_N, F_ST, F_PIX, F_FRAME, F_CTR,
    F_COUNT,
    F_CX = F_COUNT, F_CY, F_CZ,  // MIS only: radiance the pending env-direction walk adds if it survives
    F_COUNT_MIS
};
__host__ __device__ constexpr int fields(bool mis) { return mis ? (int)F_COUNT_MIS : (int)F_COUNT; }
__host__ __device__ constexpr int warp_smem_words(bool mis) { return fields(mis) * kPool + kPool / 4 + 32 / 4; }  // pool + mode bytes + batch bytes

#define FLD(f) pool[(f) * kPool + slot]
#define LDF(f) FLD(f)
#define LDU(f) __float_as_uint(FLD(f))
#define STF(f, v) FLD(f) = (v)
#define STU(f, v) FLD(f) = __uint_as_float(v)
#define LD3(f) f3(FLD(f), FLD((f) + 1), FLD((f) + 2))
#define ST3(f, v)              \
    do {                       \
        FLD(f)       = (v).x;  \
        FLD((f) + 1) = (v).y;  \
        FLD((f) + 2) = (v).z;  \
    } while (0)

template <int VT, bool JULIA, bool GRAY, bool STATS, bool MIS>
__global__ void __launch_bounds__(kThreads) k_render_wave(const __grid_constant__ Scene S, float4* __restrict__ d_sum, int first_frame,
                                                          int n_frames, int frame_stride, const __grid_constant__ vp_param P,
                                                          unsigned long long* __restrict__ d_work,
                                                          unsigned long long* __restrict__ d_stats, int skip_rt)
{
    __shared__ float smem[kWarps * warp_smem_words(MIS)];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float*         pool  = smem + warp * warp_smem_words(MIS);
    uint8_t*       modes = reinterpret_cast<uint8_t*>(pool + fields(MIS) * kPool);
    uint8_t*       batch = modes + kPool;

    const uint32_t tiles_x = (P.width + 7) >> 3, tiles_y = (P.height + 3) >> 2;
    const unsigned long long n_items = (unsigned long long)tiles_x * tiles_y * 32ull * (unsigned long long)n_frames;
    const float3 sig_t = f3(P.sigma_t.x, P.sigma_t.y, P.sigma_t.z);
    const float3 sig_s = sig_t * f3(P.albedo.x, P.albedo.y, P.albedo.z);
    const float  max_sig_t = max_of(sig_t), min_sig_t = min_of(sig_t);

#pragma unroll
    for (int r = 0; r < kRows; r++)
    {
        modes[r * 32 + lane]                    = kModePath;
        pool[F_ST * kPool + r * 32 + lane]      = __uint_as_float(kModePath);
    }
    __syncwarp();
    // warp-uniform bookkeeping
    unsigned long long w_next = 0, w_end = 0;
    int cnt[5] = {0, kPool, 0, 0, 0};  // states per mode (index = mode)
    unsigned long long c_blk[4] = {0, 0, 0, 0}, c_act[4] = {0, 0, 0, 0};

    for (;;)
    {
        // ---- pick the event type with the most waiting states (ties: step > segment > scatter > path) ----
        int pick = kModeIdle, best = 0;
#pragma unroll
        for (int m = 1; m <= 4; m++)
            if (cnt[m] >= best && cnt[m] > 0)
            {
                best = cnt[m];
                pick = m;
            }
        if (pick == kModeIdle) break;

        // ---- compact up to 32 states of that type into a batch (ballot / popc ranks) ----
        int base = 0;
#pragma unroll
        for (int r = 0; r < kRows; r++)
        {
            const bool     mine = modes[r * 32 + lane] == pick;
            const uint32_t b    = __ballot_sync(kFull, mine);
            const int      rank = base + __popc(b & ((1u << lane) - 1u));
            if (mine && rank < 32) batch[rank] = (uint8_t)(r * 32 + lane);
            base += __popc(b);
        }
        __syncwarp();
        const int  nb     = base < 32 ? base : 32;
        const bool active = (int)lane < nb;
        const int  slot   = active ? batch[lane] : 0;
        uint32_t   st     = active ? LDU(F_ST) : kModeIdle;
        if (STATS)
        {
#pragma unroll
            for (int m = 1; m <= 4; m++)
                if (m == pick)
                {
                    if (lane == 0) c_blk[m - 1]++;
                    if (active) c_act[m - 1]++;
                }
        }

        if (pick == kModePath)
        {
            // finish escaped paths, then hand out new items
            if (active && (st & kEscaped))
            {
                float3 s = LD3(F_SX), T = LD3(F_TX), L = LD3(F_LX);
                int    n = (int)LDU(F_N);
                if (!MIS || n == 0) L = L + background(S, s, n) * (GRAY ? f3(T.x) : T);  // K.cu:2026-2030
                accumulate(d_sum, LDU(F_PIX), L, n, P.brightness);
            }
            const uint32_t avail = (uint32_t)(w_end - w_next);
            unsigned long long item;
            if ((uint32_t)nb > avail)
            {
                unsigned long long b0 = 0;
                if (lane == 0) b0 = atomicAdd(d_work, (unsigned long long)kClaim);
                b0     = __shfl_sync(kFull, b0, 0);
                item   = lane < avail ? w_next + lane : b0 + (lane - avail);
                w_next = b0 + ((uint32_t)nb - avail);
                w_end  = b0 + kClaim;
            }
            else
            {
                item = w_next + lane;
                w_next += (uint32_t)nb;
            }
            if (active)
            {
                st = kModePath;
                if (item >= n_items)
                    st = kModeIdle;
                else
                {
                    uint32_t x, y, f;
                    item_to_sample(item, (uint32_t)n_frames, tiles_x, x, y, f);
                    if (x < P.width && y < P.height)
                    {
                        float3 o, s;
                        camera_ray_fast(S, x, y, P.width, P.height, o, s);
                        ST3(F_OX, o);
                        ST3(F_SX, s);
                        ST3(F_TX, f3(1.f));
                        ST3(F_LX, f3(0.f));
                        STU(F_N, 0u);
                        STU(F_PIX, y * P.width + x);
                        STU(F_FRAME, (uint32_t)(first_frame + (int)f * frame_stride));
                        STU(F_CTR, 0u);
                        st = kModeSeg | kNeedRay;
                    }
                }
            }
        }
        else if (pick == kModeSeg)
        {
            if (active)
            {
                float3 o = LD3(F_OX), s = LD3(F_SX);
                float  dist = LDF(F_DIST), t_exit = LDF(F_TEXIT);
                if (st & kNeedRay)
                {
                    float tn, tf;
                    box_slabs_fast(S, o, s, tn, tf);
                    dist   = fmaxf(tn, 0.0f);
                    t_exit = (tf > tn && tf >= 1e-3f) ? tf : -1.0f;
                    STF(F_TEXIT, t_exit);
                }
                bool found = false;
                while (dist < t_exit)
                {
                    float  seg_end = JULIA ? t_exit : fminf(dist + kSearchRadius, t_exit);
                    float2 bnd     = JULIA ? make_float2(1.0f, 0.0f) : bounds_at(S, o + s * dist);
                    if (bnd.x <= 0.0f)
                    {
                        dist = fminf(dist + fmaxf(kSearchRadius, -bnd.x), t_exit);
                        continue;
                    }
                    int   n    = (int)LDU(F_N);
                    float dmax = fmaxf(1e-4f, bnd.x);
                    float sr   = fmaxf(0.0f, fminf(1.0f, (n - 5) * 0.066666666666666666667f));
                    float dens = ((1 - sr) + sr * (1 - P.g)) * P.density;
                    float maj  = max_sig_t * dens * dmax;
                    float lim  = seg_end, sigc, inv;
                    st         = kModeStep;
                    if (bnd.y > 0.0f)
                    {
                        uint32_t ctr = LDU(F_CTR);
                        float    u0, u1;
                        philox_draw(LDU(F_PIX), LDU(F_FRAME), ctr, u0, u1);
                        STU(F_CTR, ctr);
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
                    STF(F_DMAX, dmax); STF(F_DENS, dens); STF(F_MAJ, maj); STF(F_LIM, lim); STF(F_SIGC, sigc); STF(F_INV, inv);
                    found = true;
                    break;
                }
                STF(F_DIST, dist);
                if (!found) st = kModePath | kEscaped;
            }
        }
        else if (pick == kModeStep)
        {
            if (active)
            {
                float3   o = LD3(F_OX), s = LD3(F_SX);
                float3   T = GRAY ? f3(LDF(F_TX)) : LD3(F_TX);
                float    dist = LDF(F_DIST), lim = LDF(F_LIM), inv = LDF(F_INV), dens = LDF(F_DENS), maj = LDF(F_MAJ), sigc = LDF(F_SIGC);
                uint32_t ctr = LDU(F_CTR);
                const uint32_t key = LDU(F_PIX), frame = LDU(F_FRAME);
#pragma unroll 1
                for (int rep = 0; rep < 2 && (st & kModeMask) == kModeStep; rep++)
                {
                    float u0, u1;
                    philox_draw(key, frame, ctr, u0, u1);
                    dist += -__logf(u0) * inv;
                    const bool past = dist >= lim;
                    float3     pos  = o + s * (past ? lim : dist);
                    float      den  = 0.0f;
                    if (!past)
                    {
                        if (!JULIA && GRAY && skip_rt)
                        {
                            // empty-brick skipping, the megakernel's rule and arithmetic (use_brick_skip, density_at_skip)
                            float skip;
                            den = density_at_skip<VT, JULIA>(S, pos, s, skip) * dens;
                            if (sigc == 0.0f || (st & kShadow)) dist += skip;
                        }
                        else
                            den = density_at<VT, JULIA>(S, pos) * dens;
                    }
                    if (st & kShadow)
                    {
                        if (!past)
                        {
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
                            float3 a = f3((st & kKillX) ? 0.f : 1.f, (st & kKillY) ? 0.f : 1.f, (st & kKillZ) ? 0.f : 1.f);
                            float3 L = LD3(F_LX);
                            if (MIS)
                            {
                                // sun walk done -> MIS stage; MIS walk done -> direction sampling stage
                                if (st & kMisWalk)
                                {
                                    L  = L + LD3(F_CX) * a;
                                    st = kModeScat | kScatC;
                                }
                                else
                                {
                                    L  = L + S.sun_power * ((GRAY ? f3(T.x) : T) * LDF(F_PH) * a);
                                    st = kModeScat | kScatB;
                                }
                                ST3(F_LX, L);
                            }
                            else
                            {
                                L = L + S.sun_power * ((GRAY ? f3(T.x) : T) * LDF(F_PH) * a);
                                ST3(F_LX, L);
                                s  = LD3(F_PX);
                                ST3(F_SX, s);
                                st = kModeSeg | kNeedRay;
                                int n = (int)LDU(F_N);
                                if (n >= kMaxDepth)
                                {
                                    accumulate(d_sum, key, L, n, P.brightness);
                                    st = kModePath;
                                }
                            }
                        }
                    }
                    else if (past)
                    {
                        if (st & kLimIsCtrl)
                        {
                            o = pos;
                            ST3(F_OX, o);
                            st = kModeScat;
                        }
                        else
                        {
                            dist = lim;
                            st   = kModeSeg;
                        }
                    }
                    else if (GRAY)
                    {
                        float t_den = sig_t.x * den - sigc;
                        float s_den = sig_s.x * den - sigc;
                        float n_den = maj - t_den;
                        float at = fabsf(t_den), an = fabsf(n_den), c = at + an;
                        bool  hit = u1 * c < at;
                        float k   = __fdividef(c, maj * (hit ? at : an));
                        T.x *= (hit ? s_den : n_den) * k;
                        if (hit)
                        {
                            o = pos;
                            ST3(F_OX, o);
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
                        bool   hit = u1 * c < Ps;
                        float  k   = __fdividef(c, maj * (hit ? Ps : Pn));
                        T          = T * ((hit ? s_den : n_den) * k);
                        if (hit)
                        {
                            o = pos;
                            ST3(F_OX, o);
                            st = kModeScat;
                        }
                    }
                }
                STF(F_DIST, dist);
                STU(F_CTR, ctr);
                if (GRAY)
                    STF(F_TX, T.x);
                else
                    ST3(F_TX, T);
            }
        }
        else  // kModeScat
        {
            if (active && MIS && (st & (kScatB | kScatC)))
            {
                // ---- env-map sampling variant, stages B and C of a scattering event (incoming direction in F_P*) ----
                float3   o = LD3(F_OX);
                float3   T = GRAY ? f3(LDF(F_TX)) : LD3(F_TX);
                int      n = (int)LDU(F_N);
                uint32_t ctr = LDU(F_CTR);
                const uint32_t key = LDU(F_PIX), frame = LDU(F_FRAME);
                float  sr_pre = fmaxf(0.0f, fminf(1.0f, (n - 1 - 5) * 0.066666666666666666667f));
                float  g      = (1 - sr_pre) * P.g;
                float3 ft, fb;
                const float3 din = LD3(F_PX);
                make_frame_fast(din, ft, fb);
                if (st & kScatB)
                {
                    // one-sample MIS between phase-function and env-map sampling (K.cu:2220-2297)
                    float rsel, u, v, unused;
                    philox_draw(key, frame, ctr, rsel, u);
                    philox_draw(key, frame, ctr, v, unused);
                    float3 dir, envc, C;
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
                    ST3(F_CX, C);
                    if (ok)
                    {
                        float3 s = normalize3(dir);
                        float  tn, tf;
                        box_slabs_fast(S, o, s, tn, tf);
                        ST3(F_SX, s);
                        STF(F_DIST, 0.0f);
                        STF(F_LIM, (tf > tn && tf >= 1e-3f) ? tf : 0.0f);
                        STF(F_INV, __fdividef(1.0f, max_sig_t * LDF(F_DENS) * LDF(F_DMAX)));
                        st = kModeStep | kShadow | kMisWalk;
                    }
                    else
                        st = kModeScat | kScatC;
                }
                else
                {
                    float r0, r1;
                    philox_draw(key, frame, ctr, r0, r1);
                    float3 l = hg_sample_local_fast(g, r0, r1);
                    ST3(F_SX, normalize3(ft * l.x + fb * l.y + din * l.z));
                    st = kModeSeg | kNeedRay;
                    if (n >= kMaxDepth)
                    {
                        accumulate(d_sum, key, LD3(F_LX), n, P.brightness);
                        st = kModePath;
                    }
                }
                STU(F_CTR, ctr);
            }
            else if (active)
            {
                float3   o = LD3(F_OX), s = LD3(F_SX);
                float3   T = GRAY ? f3(LDF(F_TX)) : LD3(F_TX);
                int      n = (int)LDU(F_N);
                uint32_t ctr = LDU(F_CTR);
                const uint32_t key = LDU(F_PIX), frame = LDU(F_FRAME);
                float sr_pre = fmaxf(0.0f, fminf(1.0f, (n - 5) * 0.066666666666666666667f));
                float g      = (1 - sr_pre) * P.g;
                n++;
                STU(F_N, (uint32_t)n);
                float  ph = hg_eval_fast(g, dot3(s, S.sun_dir));
                float3 pend;
                if (MIS)
                    pend = s;  // keep the incoming direction: the new one is sampled after both NEE walks (stage C)
                else
                {
                    float3 ft, fb;
                    make_frame_fast(s, ft, fb);
                    float r0, r1;
                    philox_draw(key, frame, ctr, r0, r1);
                    STU(F_CTR, ctr);
                    float3 l = hg_sample_local_fast(g, r0, r1);
                    pend     = normalize3(ft * l.x + fb * l.y + s * l.z);
                }
                float  sr   = fmaxf(0.0f, fminf(1.0f, (n - 5) * 0.066666666666666666667f));
                float  dens = ((1 - sr) + sr * (1 - P.g)) * P.density;
                STF(F_DENS, dens);
                if ((int)frame > 10 && n > 20)
                {
                    float  tau = (!JULIA && S.opacity_oct) ? opacity_at(S, o) : 0.0f;
                    float3 a   = f3(__expf(-sig_t.x * dens * tau), __expf(-sig_t.y * dens * tau), __expf(-sig_t.z * dens * tau));
                    float3 L   = LD3(F_LX);
                    L          = L + S.sun_power * ((GRAY ? f3(T.x) : T) * ph * a);
                    ST3(F_LX, L);
                    if (MIS)
                    {
                        ST3(F_PX, pend);
                        st = kModeScat | kScatB;
                    }
                    else
                    {
                        ST3(F_SX, pend);
                        st = kModeSeg | kNeedRay;
                        if (n >= kMaxDepth)
                        {
                            accumulate(d_sum, key, L, n, P.brightness);
                            st = kModePath;
                        }
                    }
                }
                else
                {
                    float inv = __fdividef(1.0f, max_sig_t * dens * LDF(F_DMAX));
                    s         = S.sun_dir;
                    float tn, tf;
                    box_slabs_inv(S, o, S.sun_inv, tn, tf);  // the megakernel's expression, bit for bit
                    float lim = (tf > tn && tf >= 1e-3f) ? tf : 0.0f;
                    if (!JULIA && S.sun_clear) lim = fminf(lim, sun_clear_at(S, o));
                    STF(F_INV, inv); STF(F_LIM, lim); STF(F_DIST, 0.0f); STF(F_PH, ph);
                    ST3(F_SX, s);
                    ST3(F_PX, pend);
                    st = kModeStep | kShadow;
                }
            }
        }

        // ---- re-file the batch by its new event types ----
        const uint32_t nm = st & kModeMask;
        if (active)
        {
            STU(F_ST, st);
            modes[slot] = (uint8_t)nm;
        }
#pragma unroll
        for (int m = 1; m <= 4; m++) cnt[m] += __popc(__ballot_sync(kFull, active && nm == (uint32_t)m)) - (m == pick ? nb : 0);
        __syncwarp();
    }
    if (STATS)
    {
        for (int i = 0; i < 4; i++)
        {
            if (lane == 0) atomicAdd(d_stats + 8 + i, c_blk[i]);
            atomicAdd(d_stats + 12 + i, c_act[i]);
        }
    }
}
}  // namespace wave

template <int VT, bool JULIA, bool GRAY, bool MIS>
static cudaError_t launch_wave_t(const Scene& S, float4* d_sum, int first_frame, int n_frames, int frame_stride, const vp_param& P,
                                 unsigned long long* d_work, unsigned long long* d_stats, int num_sms, cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(d_work, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    if (d_stats)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wave::k_render_wave<VT, JULIA, GRAY, true, MIS>, wave::kThreads, 0);
    else
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wave::k_render_wave<VT, JULIA, GRAY, false, MIS>, wave::kThreads, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    unsigned long long items = (unsigned long long)((P.width + 7) / 8) * ((P.height + 3) / 4) * 32ull * n_frames;
    unsigned long long want  = (items + wave::kClaim - 1) / wave::kClaim;
    unsigned long long ctas  = (want + wave::kWarps - 1) / wave::kWarps;
    unsigned long long cap   = (unsigned long long)num_sms * per_sm;
    unsigned int       grid  = (unsigned int)(ctas < cap ? ctas : cap);
    if (grid < 1) grid = 1;
    const int skip = use_brick_skip(S, P.density, fmaxf(P.sigma_t.x, fmaxf(P.sigma_t.y, P.sigma_t.z)), GRAY) ? 1 : 0;
    if (d_stats)
        wave::k_render_wave<VT, JULIA, GRAY, true, MIS><<<grid, wave::kThreads, 0, stream>>>(S, d_sum, first_frame, n_frames, frame_stride, P, d_work, d_stats, skip);
    else
        wave::k_render_wave<VT, JULIA, GRAY, false, MIS><<<grid, wave::kThreads, 0, stream>>>(S, d_sum, first_frame, n_frames, frame_stride, P, d_work, nullptr, skip);
    return cudaGetLastError();
}

cudaError_t launch_render_wave(const Scene& S, float4* d_sum, int first_frame, int n_frames, int frame_stride, const vp_param& P,
                               unsigned long long* d_work, unsigned long long* d_stats, int num_sms, cudaStream_t stream)
{
    const bool gray = P.sigma_t.x == P.sigma_t.y && P.sigma_t.y == P.sigma_t.z && P.albedo.x == P.albedo.y && P.albedo.y == P.albedo.z;
#define VP_WAVE(VT, J)                                                                                                                     \
    do {                                                                                                                                   \
        if (S.env_mis)                                                                                                                     \
            return gray ? launch_wave_t<VT, J, true, true>(S, d_sum, first_frame, n_frames, frame_stride, P, d_work, d_stats, num_sms, stream)   \
                        : launch_wave_t<VT, J, false, true>(S, d_sum, first_frame, n_frames, frame_stride, P, d_work, d_stats, num_sms, stream); \
        return gray ? launch_wave_t<VT, J, true, false>(S, d_sum, first_frame, n_frames, frame_stride, P, d_work, d_stats, num_sms, stream)      \
                    : launch_wave_t<VT, J, false, false>(S, d_sum, first_frame, n_frames, frame_stride, P, d_work, d_stats, num_sms, stream);    \
    } while (0)
    if (S.julia) VP_WAVE(kF32, true);
    if (S.voxel_type == kU8) VP_WAVE(kU8, false);
    if (S.voxel_type == kF16) VP_WAVE(kF16, false);
    VP_WAVE(kF32, false);
#undef VP_WAVE
}
}  // namespace vp
