// oracle/tex_emul.h -- TEST INFRASTRUCTURE (CPU emulation of the CUDA texture unit).
//
// Shared by oracle/host_shim.h (which lets g++ compile the reference's kernel source) and by
// oracle/volpath_oracle.cpp (the restatement).  It restates the addressing / filtering rules the
// reference relies on (CUDA C Programming Guide, appendix "Texture Fetching"):
//   * normalized coordinates: x = xn * N
//   * clamp addressing on every axis (K.cu:215-217 and the other get_texture_desc<> variants)
//   * point filter  : texel floor(x), clamped to [0, N-1]
//   * linear filter : xB = x - 0.5, i = floor(xB), alpha = frac(xB) kept in 1.8 fixed point
//                     (rounded to the nearest 1/256), neighbours i, i+1 clamped
//   * cudaReadModeNormalizedFloat for 8-bit texels: value / 255 (K.cu:247,261)
// The arithmetic of the lerp itself is not documented by NVIDIA; we use fp32
// (1-a)*p + a*q along x, then y, then z with no FMA contraction.  The CUDA parity kernels of the
// product restate exactly this sequence, so CPU/GPU comparisons are meaningful; agreement with the
// real texture unit is measured on the GPU box against the rebuilt reference kernel.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace texemu
{
enum Format
{
    FMT_U8   = 0,  // 1 x uint8, normalized float read
    FMT_U8x2 = 1,  // 2 x uint8, normalized float read
    FMT_F32  = 2,
    FMT_F32x2 = 3,
    FMT_F32x4 = 4,
};

struct Array
{
    int                  w = 0, h = 1, d = 1;
    Format               fmt = FMT_F32;
    std::vector<uint8_t> bytes;

    int comps() const { return fmt == FMT_U8 || fmt == FMT_F32 ? 1 : (fmt == FMT_F32x4 ? 4 : 2); }
    int comp_size() const { return (fmt == FMT_U8 || fmt == FMT_U8x2) ? 1 : 4; }
    size_t texel_size() const { return (size_t)comps() * comp_size(); }
    void   alloc(int w_, int h_, int d_, Format f)
    {
        w = w_; h = h_ > 0 ? h_ : 1; d = d_ > 0 ? d_ : 1; fmt = f;
        bytes.assign((size_t)w * h * d * texel_size(), 0);
    }
    // component c of texel (i,j,k), already clamped by the caller
    inline float get(int i, int j, int k, int c) const
    {
        size_t idx = ((size_t)k * h + j) * w + i;
        if (comp_size() == 1) return (float)bytes[idx * comps() + c] / 255.0f;
        float v;
        memcpy(&v, &bytes[(idx * comps() + c) * 4], 4);
        return v;
    }
};

inline int clampi(int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }

inline int point_texel(float x_unnorm, int n)
{
    return clampi((int)floorf(x_unnorm), n);
}

// i0 and the 1.8 fixed-point weight of the upper neighbour
inline void linear_texel(float x_unnorm, int n, int& i0, int& i1, float& a)
{
    float xb = x_unnorm - 0.5f;
    float fl = floorf(xb);
    a        = xb - fl;
    a        = floorf(a * 256.0f + 0.5f) * (1.0f / 256.0f);
    int i    = (int)fl;
    i0       = clampi(i, n);
    i1       = clampi(i + 1, n);
}

inline float lerp1(float p, float q, float a) { return (1.0f - a) * p + a * q; }

struct Texture
{
    const Array* arr        = nullptr;
    bool         linear     = false;
    bool         normalized = true;
};

// one component of a 3D fetch
inline float fetch3(const Texture& t, float x, float y, float z, int c)
{
    const Array& A = *t.arr;
    if (t.normalized) { x *= (float)A.w; y *= (float)A.h; z *= (float)A.d; }
    if (!t.linear) return A.get(point_texel(x, A.w), point_texel(y, A.h), point_texel(z, A.d), c);
    int   i0, i1, j0, j1, k0, k1;
    float a, b, g;
    linear_texel(x, A.w, i0, i1, a);
    linear_texel(y, A.h, j0, j1, b);
    linear_texel(z, A.d, k0, k1, g);
    float c00 = lerp1(A.get(i0, j0, k0, c), A.get(i1, j0, k0, c), a);
    float c10 = lerp1(A.get(i0, j1, k0, c), A.get(i1, j1, k0, c), a);
    float c01 = lerp1(A.get(i0, j0, k1, c), A.get(i1, j0, k1, c), a);
    float c11 = lerp1(A.get(i0, j1, k1, c), A.get(i1, j1, k1, c), a);
    float c0  = lerp1(c00, c10, b);
    float c1  = lerp1(c01, c11, b);
    return lerp1(c0, c1, g);
}

inline float fetch2(const Texture& t, float x, float y, int c)
{
    const Array& A = *t.arr;
    if (t.normalized) { x *= (float)A.w; y *= (float)A.h; }
    if (!t.linear) return A.get(point_texel(x, A.w), point_texel(y, A.h), 0, c);
    int   i0, i1, j0, j1;
    float a, b;
    linear_texel(x, A.w, i0, i1, a);
    linear_texel(y, A.h, j0, j1, b);
    float c0 = lerp1(A.get(i0, j0, 0, c), A.get(i1, j0, 0, c), a);
    float c1 = lerp1(A.get(i0, j1, 0, c), A.get(i1, j1, 0, c), a);
    return lerp1(c0, c1, b);
}

inline float fetch1(const Texture& t, float x, int c)
{
    const Array& A = *t.arr;
    if (t.normalized) x *= (float)A.w;
    if (!t.linear) return A.get(point_texel(x, A.w), 0, 0, c);
    int   i0, i1;
    float a;
    linear_texel(x, A.w, i0, i1, a);
    return lerp1(A.get(i0, 0, 0, c), A.get(i1, 0, 0, c), a);
}
}  // namespace texemu
