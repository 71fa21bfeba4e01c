// oracle/host_shim.h -- TEST INFRASTRUCTURE.
//
// Lets plain g++ compile the reference's CUDA kernel source (a patched TEMP copy, see
// build_ref.py) as host code, so the reference's own tracking loop can run on CPU cores:
//   * threadIdx / blockIdx / blockDim become thread-local variables driven by VPREF_LAUNCH
//     (OpenMP over blocks, schedule(dynamic)), replacing the <<<grid, block>>> launches
//   * cudaArray / texture / surface objects and tex1D/2D/3D, surf3Dread/write are emulated in
//     software by oracle/tex_emul.h (clamp addressing, point / 1.8-fixed-point linear filter)
//   * cudaMemcpyToSymbol[Async] becomes memcpy into the (now ordinary) global
// One reference statement needs care on the host: nvcc device code evaluates
// `phase.sample(frame, rng.next(), rng.next())` left to right, g++ right to left; build_ref.py
// rewrites that call to VPREF_SAMPLE_LR so the host build draws in the device's order.
#pragma once
#define __DEVICE_LAUNCH_PARAMETERS_H__ 1  // we provide threadIdx & co. ourselves
#define CURAND_KERNEL_H_ 1                // the kernel includes curand_kernel.h but uses nothing of it
#include <cuda_runtime.h>
#include <omp.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "tex_emul.h"

namespace vps
{
static thread_local uint3 t_threadIdx;
static thread_local uint3 t_blockIdx;
static thread_local dim3  t_blockDim;
static thread_local dim3  t_gridDim;

struct HostTexture
{
    texemu::Texture tex;
};

inline texemu::Format format_of(const cudaChannelFormatDesc& d)
{
    int n = (d.x > 0) + (d.y > 0) + (d.z > 0) + (d.w > 0);
    if (d.f == cudaChannelFormatKindFloat)
        return n == 1 ? texemu::FMT_F32 : (n == 2 ? texemu::FMT_F32x2 : texemu::FMT_F32x4);
    return n == 1 ? texemu::FMT_U8 : texemu::FMT_U8x2;
}

inline cudaError_t malloc3DArray(cudaArray_t* a, const cudaChannelFormatDesc* d, cudaExtent e)
{
    auto* A = new texemu::Array();
    A->alloc((int)e.width, (int)e.height, (int)e.depth, format_of(*d));
    *a = reinterpret_cast<cudaArray_t>(A);
    return cudaSuccess;
}
inline cudaError_t mallocArray(cudaArray_t* a, const cudaChannelFormatDesc* d, size_t w, size_t h = 0)
{
    return malloc3DArray(a, d, make_cudaExtent(w, h, 0));
}
inline cudaError_t freeArray(cudaArray_t a)
{
    delete reinterpret_cast<texemu::Array*>(a);
    return cudaSuccess;
}
inline cudaError_t memcpy3D(const cudaMemcpy3DParms* p)
{
    auto*  A  = reinterpret_cast<texemu::Array*>(p->dstArray);
    size_t ts = A->texel_size();
    for (size_t k = 0; k < p->extent.depth; k++)
        for (size_t j = 0; j < p->extent.height; j++)
        {
            const char* src = (const char*)p->srcPtr.ptr + (k * p->srcPtr.ysize + j) * p->srcPtr.pitch;
            memcpy(&A->bytes[((k * A->h + j) * A->w) * ts], src, p->extent.width * ts);
        }
    return cudaSuccess;
}
inline cudaError_t memcpy2DToArray(cudaArray_t dst, size_t, size_t, const void* src, size_t spitch,
                                   size_t width_bytes, size_t height, cudaMemcpyKind)
{
    auto*  A  = reinterpret_cast<texemu::Array*>(dst);
    size_t ts = A->texel_size();
    for (size_t j = 0; j < height; j++)
        memcpy(&A->bytes[(j * A->w) * ts], (const char*)src + j * spitch, width_bytes);
    return cudaSuccess;
}
inline cudaError_t memcpyToArray(cudaArray_t dst, size_t, size_t, const void* src, size_t bytes, cudaMemcpyKind)
{
    auto* A = reinterpret_cast<texemu::Array*>(dst);
    memcpy(A->bytes.data(), src, bytes);
    return cudaSuccess;
}
inline cudaError_t createTextureObject(cudaTextureObject_t* t, const cudaResourceDesc* r,
                                       const cudaTextureDesc* d, const void*)
{
    auto* T          = new HostTexture();
    T->tex.arr       = reinterpret_cast<texemu::Array*>(r->res.array.array);
    T->tex.linear    = d->filterMode == cudaFilterModeLinear;
    T->tex.normalized = d->normalizedCoords != 0;
    *t               = (cudaTextureObject_t) reinterpret_cast<uintptr_t>(T);
    return cudaSuccess;
}
inline cudaError_t createSurfaceObject(cudaSurfaceObject_t* s, const cudaResourceDesc* r)
{
    *s = (cudaSurfaceObject_t) reinterpret_cast<uintptr_t>(r->res.array.array);
    return cudaSuccess;
}
inline cudaError_t destroyObject(unsigned long long) { return cudaSuccess; }  // leaked, as in the reference

inline const texemu::Texture& T(cudaTextureObject_t t)
{
    return reinterpret_cast<HostTexture*>((uintptr_t)t)->tex;
}

template <class K>
struct Launcher
{
    K    k;
    dim3 g, b;
    template <class... A>
    void operator()(A... a) const
    {
        long nb = (long)g.x * g.y * g.z;
#pragma omp parallel for schedule(dynamic)
        for (long bi = 0; bi < nb; bi++)
        {
            t_gridDim    = g;
            t_blockDim   = b;
            t_blockIdx.x = (unsigned)(bi % g.x);
            t_blockIdx.y = (unsigned)((bi / g.x) % g.y);
            t_blockIdx.z = (unsigned)(bi / ((long)g.x * g.y));
            for (unsigned tz = 0; tz < b.z; tz++)
                for (unsigned ty = 0; ty < b.y; ty++)
                    for (unsigned tx = 0; tx < b.x; tx++)
                    {
                        t_threadIdx = make_uint3(tx, ty, tz);
                        k(a...);
                    }
        }
    }
};
template <class K>
Launcher<K> make_launcher(K k, dim3 g, dim3 b)
{
    return Launcher<K>{k, g, b};
}
}  // namespace vps

#define threadIdx vps::t_threadIdx
#define blockIdx vps::t_blockIdx
#define blockDim vps::t_blockDim
#define gridDim vps::t_gridDim
#define VPREF_LAUNCH(k, g, b) vps::make_launcher(k, dim3(g), dim3(b))

#define cudaMalloc3DArray vps::malloc3DArray
#define cudaMallocArray vps::mallocArray
#define cudaFreeArray vps::freeArray
#define cudaMemcpy3D vps::memcpy3D
#define cudaMemcpy2DToArrayAsync vps::memcpy2DToArray
#define cudaMemcpy2DToArray vps::memcpy2DToArray
#define cudaMemcpyToArray vps::memcpyToArray
#define cudaCreateTextureObject vps::createTextureObject
#define cudaCreateSurfaceObject vps::createSurfaceObject
#define cudaDestroyTextureObject vps::destroyObject
#define cudaDestroySurfaceObject vps::destroyObject
#define cudaMemcpyToSymbolAsync(sym, src, n) (memcpy((void*)&(sym), (src), (n)), cudaSuccess)
#define cudaMemcpyToSymbol(sym, src, n) (memcpy((void*)&(sym), (src), (n)), cudaSuccess)

// host stand-in for the legacy-texture-reference bind (device build: see vpref_prelude.h)
#define VPREF_BIND(sym, array__, norm)                               \
    do {                                                             \
        auto* T_          = new vps::HostTexture();                  \
        T_->tex.arr       = reinterpret_cast<texemu::Array*>(array__); \
        T_->tex.linear    = false;                                   \
        T_->tex.normalized = (norm) != 0;                            \
        sym               = (cudaTextureObject_t)(uintptr_t)T_;      \
    } while (0)

// the device's left-to-right draw order for phase.sample(frame, rng.next(), rng.next()) (Q6)
#define VPREF_SAMPLE_LR(phase, frame, rng) \
    ([&]() { float r0_ = (rng).next(); float r1_ = (rng).next(); return (phase).sample((frame), r0_, r1_); }())

static inline float __uint_as_float(unsigned int u)
{
    float f;
    memcpy(&f, &u, 4);
    return f;
}

template <class R>
struct vps_fetch;
template <>
struct vps_fetch<float>
{
    static float f3(const texemu::Texture& t, float x, float y, float z) { return texemu::fetch3(t, x, y, z, 0); }
    static float f2(const texemu::Texture& t, float x, float y) { return texemu::fetch2(t, x, y, 0); }
    static float f1(const texemu::Texture& t, float x) { return texemu::fetch1(t, x, 0); }
};
template <>
struct vps_fetch<float2>
{
    static float2 f3(const texemu::Texture& t, float x, float y, float z)
    {
        return make_float2(texemu::fetch3(t, x, y, z, 0), texemu::fetch3(t, x, y, z, 1));
    }
};
template <>
struct vps_fetch<float4>
{
    static float4 f2(const texemu::Texture& t, float x, float y)
    {
        return make_float4(texemu::fetch2(t, x, y, 0), texemu::fetch2(t, x, y, 1), texemu::fetch2(t, x, y, 2),
                           texemu::fetch2(t, x, y, 3));
    }
};

template <class R>
static inline R tex3D(cudaTextureObject_t t, float x, float y, float z)
{
    return vps_fetch<R>::f3(vps::T(t), x, y, z);
}
template <class R>
static inline R tex2D(cudaTextureObject_t t, float x, float y)
{
    return vps_fetch<R>::f2(vps::T(t), x, y);
}
template <class R>
static inline R tex1D(cudaTextureObject_t t, float x)
{
    return vps_fetch<R>::f1(vps::T(t), x);
}

// surface access: byte x offset, as in CUDA (K.cu:186-196)
template <class V>
static inline void surf3Dwrite(const V& v, cudaSurfaceObject_t s, int xbytes, int y, int z)
{
    auto* A = reinterpret_cast<texemu::Array*>((uintptr_t)s);
    memcpy(&A->bytes[(((size_t)z * A->h + y) * A->w) * A->texel_size() + xbytes], &v, sizeof(V));
}
template <class V>
static inline V surf3Dread(cudaSurfaceObject_t s, int xbytes, int y, int z)
{
    auto* A = reinterpret_cast<texemu::Array*>((uintptr_t)s);
    V     v;
    memcpy(&v, &A->bytes[(((size_t)z * A->h + y) * A->w) * A->texel_size() + xbytes], sizeof(V));
    return v;
}
