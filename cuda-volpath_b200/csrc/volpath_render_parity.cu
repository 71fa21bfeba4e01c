// volpath_render_parity.cu -- the reference-faithful renderer (VP_MODE_PARITY).
//
// One thread = one pixel, like the reference's __d_render_bounded_decomp (K.cu:1958-2318); the
// kernel restates that estimator on the octet store with the reference RNG, the reference's draw
// order and all its quirks (Q1-Q12), so that a fixed (pixel, frame) produces the
// same path as the reference kernel.  Frames [first, first + n) are looped INSIDE the thread and the
// running pixel sum is kept in registers: v = sum[pix]; v += s_0; v += s_1; ... ; sum[pix] = v --
// the same float additions in the same order as n reference launches, with one RMW instead of n.
#include "volpath_common.cuh"
#include "volpath_kernels.h"

namespace vp
{
template <int VT, bool JULIA>
struct ParityMedium
{
    static constexpr bool kJulia = JULIA;
    // vol_sigma_t (K.cu:682-708)
    static __device__ __forceinline__ float sigma(const Scene& S, float3 pos, float density)
    {
        float t = JULIA ? julia_density(pos * 1.0f) : fetch_density_parity<VT>(S, pos);
        t *= density;
        return t;
    }
    // vol_bound_minmax (K.cu:1610-1624): point lookup of (max, min)
    static __device__ __forceinline__ float2 bound(const Scene& S, float3 pos)
    {
        if (JULIA) return make_float2(1.0f, 0.0f);
        float3 p = (pos - S.bmin) * S.l_inv;
        int    i = clampi((int)floorf(__fmul_rn(p.x, (float)S.nx)), 0, S.nx - 1);
        int    j = clampi((int)floorf(__fmul_rn(p.y, (float)S.ny)), 0, S.ny - 1);
        int    k = clampi((int)floorf(__fmul_rn(p.z, (float)S.nz)), 0, S.nz - 1);
        return __ldg(S.bounds_voxel + ((size_t)k * S.ny + j) * S.nx + i);
    }
};

// Tr_spectral (K.cu:754-808): one shadow walk, per-channel kill flags, result in {0,1}^3
template <class M>
__device__ float3 tr_spectral(const Scene& S, float3 start, float3 end, float inv_sigma, float density, float3 sigma_t,
                              RefRng& rng)
{
    float3 o = start;
    float3 d = normalize3(end - start);
    float  t_near, t_far;
    box_slabs(S, o, d, t_near, t_far);
    if (!(t_far > t_near && t_far >= 1e-3f)) return f3(1.0f);
    if (t_near < 0.0f) t_near = 0.0f;
    float3 se    = start - end;
    float  max_t = fminf(t_far, sqrtf(dot3(se, se)));
    float  dist  = t_near;
    int    xterm = 0, yterm = 0, zterm = 0;
    for (;;)
    {
        dist += -logf(rng.next()) * inv_sigma;
        if (dist >= max_t || (xterm && yterm && zterm)) break;
        float3 pos = o + d * dist;
        float  e   = rng.next();
        float  den = M::sigma(S, pos, density);
        if (!xterm && e < sigma_t.x * den * inv_sigma) xterm = 1;
        if (!yterm && e < sigma_t.y * den * inv_sigma) yterm = 1;
        if (!zterm && e < sigma_t.z * den * inv_sigma) zterm = 1;
    }
    return f3((float)(1 - xterm), (float)(1 - yterm), (float)(1 - zterm));
}

template <class M, bool MIS>
__device__ float4 trace_path_parity(const Scene& S, const vp_param& P, uint32_t x, uint32_t y, int spp)
{
    const float density = P.density;
    RefRng      rng;
    rng.init(x, y, (uint32_t)spp);

    float3 o, d;
    camera_ray(S, x, y, P.width, P.height, o, d);

    float3 radiance   = f3(0.0f);
    float3 throughput = f3(1.0f);

    const float3 sigma_t_spectral = f3(P.sigma_t.x, P.sigma_t.y, P.sigma_t.z);
    const float3 sigma_s_spectral = sigma_t_spectral * f3(P.albedo.x, P.albedo.y, P.albedo.z);
    const float  max_sigma_t      = max_of(sigma_t_spectral);
    const float  min_sigma_t      = min_of(sigma_t_spectral);

    int num_scatters = 0;
    while (num_scatters < kMaxDepth)
    {
        // intersectSuperVolume (K.cu:1626-1661): slab test, segment clipped to 0.05, bound at entry
        float largest_tmin, smallest_tmax;
        box_slabs(S, o, d, largest_tmin, smallest_tmax);
        float  t_near = fmaxf(largest_tmin, 0.0f);
        float  t_far  = fminf(smallest_tmax, kSearchRadius);
        float2 bnd    = M::bound(S, o + d * t_near);
        float  d_min  = bnd.y;
        float  d_max  = fmaxf(0.0001f, bnd.x);
        bool   hit    = smallest_tmax > largest_tmin && smallest_tmax >= 1e-3f;
        bool   use_decomposition = d_min > 0.0f;
        if (!hit)
        {
            // K.cu:2026-2030: the MIS variant adds the environment only at depth 0
            if (!MIS || num_scatters == 0) radiance = radiance + background(S, d, num_scatters) * throughput;
            break;
        }
        float3 pos  = o + d * t_near;
        float  dist = t_near;

        // reduced scattering after 5 bounces (K.cu:2039-2044)
        float s = fmaxf(0.0f, fminf(1.0f, (num_scatters - 5) * 0.066666666666666666667f));
        float g = (1 - s) * P.g;
        float reduction_factor = (1 - s) + s * (1 - P.g);
        float density_prime    = reduction_factor * density;
        float sigma_t_prime    = max_sigma_t * density_prime * d_max;

        float  distc, sigma_r_prime = 0.f;
        float3 sigma_c_spectral;
        if (use_decomposition)  // K.cu:2048-2054
        {
            float sigma_c_prime = min_sigma_t * density_prime * d_min;
            distc               = dist - logf(rng.next()) / fmaxf(sigma_c_prime, 1e-20f);
            sigma_r_prime       = fmaxf(sigma_t_prime - sigma_c_prime, 1e-20f);
            sigma_c_spectral    = f3(sigma_c_prime);
        }
        else
        {
            distc            = 1e20f;
            sigma_c_spectral = f3(0.0f);
        }
        float inv_sigma_t = 1.0f / sigma_t_prime;
        float inv_sigma   = use_decomposition ? 1.0f / sigma_r_prime : inv_sigma_t;

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
            float  den            = M::sigma(S, pos, density_prime);
            float3 sigma_t_den    = sigma_t_spectral * den - sigma_c_spectral;
            float3 sigma_s_den    = sigma_s_spectral * den - sigma_c_spectral;
            float3 sigma_null_den = f3(sigma_t_prime) - sigma_t_den;
            float  Ps = fabsf(sigma_t_den.x * throughput.x) + fabsf(sigma_t_den.y * throughput.y) +
                       fabsf(sigma_t_den.z * throughput.z);
            float Pn = fabsf(sigma_null_den.x * throughput.x) + fabsf(sigma_null_den.y * throughput.y) +
                       fabsf(sigma_null_den.z * throughput.z);
            float c = (Ps + Pn);
            float e = rng.next() * c;
            if (e < Ps)
            {
                throughput = throughput * (sigma_s_den * (inv_sigma_t * c / (Ps)));
                break;
            }
            else
            {
                throughput = throughput * (sigma_null_den * (inv_sigma_t * c / Pn));
            }
        }

        bool through = fminf(distc, dist) >= t_far;
        num_scatters += (!through);
        if (through)
        {
            o = o + d * t_far;
            continue;
        }

        float3 ft, fb;
        make_frame(d, ft, fb);
        {
            float s2 = fmaxf(0.0f, fminf(1.0f, (num_scatters - 5) * 0.066666666666666666667f));
            float reduction_factor2 = (1 - s2) + s2 * (1 - P.g);
            float density_prime2    = reduction_factor2 * density;
            float sigma_t_prime2    = max_sigma_t * density_prime2 * d_max;
            float inv_sigma2        = 1.0f / sigma_t_prime2;
            float ph                = hg_evaluate(g, dot3(d, S.sun_dir));
            float3 a;
            if (spp > 10 && num_scatters > 20)  // K.cu:2183
            {
                // Julia / no table: the reference's table would be built from a zero density -> tau = 0
                float  tau = (!M::kJulia && S.have_opacity) ? fetch_opacity(S, pos, true) : 0.0f;
                float3 e3  = (-sigma_t_spectral) * density_prime2 * tau;
                a          = f3(expf(e3.x), expf(e3.y), expf(e3.z));
            }
            else
            {
                a = tr_spectral<M>(S, pos, S.sun_dir * 1e10f, inv_sigma2, density_prime2, sigma_t_spectral, rng);
            }
            radiance = radiance + S.sun_power * (throughput * ph * a);

            if (MIS)
            {
                // one-sample MIS between phase-function and env-map sampling (K.cu:2220-2297)
                const float P_phase = 0.5f, P_envmap = 1.0f - P_phase;
                if (rng.next() < P_phase)
                {
                    float  u = rng.next();
                    float  v = rng.next();
                    float3 ls = hg_sample_local(g, u, v);
                    float3 brdf_dir = ft * ls.x + fb * ls.y + d * ls.z;
                    float3 envc     = eval_envmap(S, brdf_dir);
                    float  pdf_brdf = hg_evaluate(g, dot3(d, brdf_dir));
                    float  pdf_env_virtual = pdf_envmap(S, envc);
                    float  weight = (pdf_brdf * P_phase) / (pdf_brdf * P_phase + pdf_env_virtual * P_envmap) / P_phase;
                    float3 a2 = tr_spectral<M>(S, pos, brdf_dir * 1e10f, inv_sigma2, density_prime2, sigma_t_spectral, rng);
                    radiance  = radiance + envc * (throughput * weight * a2);
                }
                else
                {
                    float  u = rng.next();
                    float  v = rng.next();
                    float3 envc;
                    float  pdf_env = sample_envmap(S, u, v, envc);
                    if (pdf_env <= 0.0f) continue;  // K.cu:2266: no new direction, the ray restarts from its old origin
                    float3 envmap_dir = uv_to_dir(u, v);
                    float  pdf_brdf_virtual = hg_evaluate(g, dot3(d, envmap_dir));
                    float  weight = (pdf_env * P_envmap) / (pdf_env * P_envmap + pdf_brdf_virtual * P_phase) / P_envmap;
                    float3 a2 = tr_spectral<M>(S, pos, envmap_dir * 1e10f, inv_sigma2, density_prime2, sigma_t_spectral, rng);
                    float3 tw = throughput * hg_evaluate(g, dot3(d, envmap_dir));
                    tw        = f3(tw.x / pdf_env, tw.y / pdf_env, tw.z / pdf_env);
                    radiance  = radiance + envc * (tw * weight * a2);
                }
            }
        }
        float  r0 = rng.next();  // device order: first draw -> cos(theta), second -> phi (Q6)
        float  r1 = rng.next();
        float3 l  = hg_sample_local(g, r0, r1);
        float3 nd = normalize3(ft * l.x + fb * l.y + d * l.z);
        o         = pos;
        d         = nd;
    }
    radiance = radiance * P.brightness;
    return make_float4(fmaxf(radiance.x, 0.0f), fmaxf(radiance.y, 0.0f), fmaxf(radiance.z, 0.0f), (float)num_scatters);
}

template <int VT, bool JULIA, bool MIS>
#ifndef VP_PARITY_MIN_BLOCKS
#define VP_PARITY_MIN_BLOCKS 16  // 61 registers, no spill: 16 CTAs of 64 threads per SM; measured against the reference kernel: 1 / 16 / 20 -> 0.71 / 0.83 / 0.83 x
#endif
__global__ void __launch_bounds__(64, VP_PARITY_MIN_BLOCKS) k_render_parity(const __grid_constant__ Scene S, float4* __restrict__ d_sum,
                                                       int first_frame, int n_frames, int frame_stride,
                                                       const __grid_constant__ vp_param P)
{
    uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.width || y >= P.height) return;
    float4 v = d_sum[x + (size_t)y * P.width];
    for (int f = 0; f < n_frames; f++)
    {
        float4 c = trace_path_parity<ParityMedium<VT, JULIA>, MIS>(S, P, x, y, first_frame + f * frame_stride);
        v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;  // K.cu:2315
    }
    d_sum[x + (size_t)y * P.width] = v;
}

cudaError_t launch_render_parity(const Scene& S, float4* d_sum, int first_frame, int n_frames, int frame_stride,
                                 const vp_param& P, cudaStream_t stream)
{
    dim3 block(8, 8);  // H.cpp:100
    dim3 grid((P.width + block.x - 1) / block.x, (P.height + block.y - 1) / block.y);
#define VP_PARITY(VT, J)                                                                                        \
    do {                                                                                                        \
        if (S.env_mis)                                                                                          \
            k_render_parity<VT, J, true><<<grid, block, 0, stream>>>(S, d_sum, first_frame, n_frames, frame_stride, P);  \
        else                                                                                                    \
            k_render_parity<VT, J, false><<<grid, block, 0, stream>>>(S, d_sum, first_frame, n_frames, frame_stride, P); \
    } while (0)
    if (S.julia)
        VP_PARITY(kF32, true);
    else if (S.voxel_type == kU8)
        VP_PARITY(kU8, false);
    else if (S.voxel_type == kF16)
        VP_PARITY(kF16, false);
    else
        VP_PARITY(kF32, false);
#undef VP_PARITY
    return cudaGetLastError();
}
}  // namespace vp
