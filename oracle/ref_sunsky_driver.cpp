// oracle/ref_sunsky_driver.cpp -- TEST INFRASTRUCTURE.
// Thin driver over the reference's own sun/sky model (src/sunsky/*, compiled where it lies by
// build_ref.py).  It follows update_sunsky() (volumeRender.cpp:276-333) to bake the lat-long env
// map and the sun direction / disk radiance the reference hands to init_envmap / set_sun.
// Used once by tests/golden/make_sunsky.py to produce the committed default-sky fixture.
#include <cmath>
#include <vector>

#include "vecmath.h"
#include "sunsky/sunsky.h"
#include "sunsky/hosek/ArHosekSkyModel.h"

namespace Tungsten { namespace Spectral { void spectralXyzWeights(int samples, float lambdas[], Vec3f weights[]); } }

extern "C" int ref_bake_sunsky(float x, float y, int hdrwidth, int hdrheight, float* rgba_out, float* sun_dir3,
                               float* sun_power3)
{
    SkyModel<Tungsten::Skydome> s;
    y *= 0.5f;
    y = fminf(fmaxf(y, 0.0f), 0.49999f);             // H.cpp:282-283
    s.setSunPhi(x * M_PI * 2);                       // H.cpp:288
    s.setSunTheta(y * M_PI);                         // H.cpp:289
    const bool      bake_sun     = false;            // H.cpp:291
    constexpr float sunsky_scale = 0.02;             // H.cpp:292
    float3          sun_dir      = s.getSunDir();
    float3          sun_power    = s.sunColor() * sunsky_scale;
    float4*         img          = reinterpret_cast<float4*>(rgba_out);
#pragma omp parallel for
    for (int i = 0; i < hdrwidth; i++)
    {
        for (int j = 0; j < hdrheight; j++)
        {
            if (j < hdrheight / 2)
            {
                float  phi   = float(i) / hdrwidth * 2 * M_PI;
                float  theta = (float(j) / hdrheight) * M_PI;
                float3 d     = make_float3(sinf(theta) * sinf(phi), cosf(theta), sinf(theta) * -cosf(phi));
                float3 c     = s.skyColor(d, bake_sun);
                img[i + j * hdrwidth] = make_float4(c, 1.0f) * sunsky_scale;
            }
            else
            {
                float3 ground_albedo      = make_float3(0.01f);
                float3 reflected_radiance = ground_albedo * sun_dir.y * sun_power * (M_PI * (0.45 / 94.0f * 0.45 / 94.0f));
                img[i + j * hdrwidth]     = make_float4(reflected_radiance, 1.0f);
            }
        }
    }
    sun_dir3[0] = sun_dir.x; sun_dir3[1] = sun_dir.y; sun_dir3[2] = sun_dir.z;
    sun_power3[0] = sun_power.x; sun_power3[1] = sun_power.y; sun_power3[2] = sun_power.z;
    return 0;
}

// The state the reference's host holds after Skydome::prepareForRender() (sky_tungsten.cpp:377-398) for the sun
// position (x, y) of setup_sunsky: the cooked Hosek configurations of the 11 spectral bands, their radiances and
// emission corrections, and the spectral XYZ weights / wavelengths.  This is the INPUT of the GPU sky bake
// (vp_bake_sunsky); the per-texel evaluation above is what that kernel replaces.
extern "C" int ref_sky_state(float x, float y, double* configs99, double* radiances11, double* ecf_sky11, float* lambdas10,
                             float* weights30, float* sun_dir3, float* sun_power3)
{
    SkyModel<Tungsten::Skydome> s;
    y *= 0.5f;
    y = fminf(fmaxf(y, 0.0f), 0.49999f);
    s.setSunPhi(x * M_PI * 2);
    s.setSunTheta(y * M_PI);
    Tungsten::Skydome dome;  // constructor defaults: temperature 5777, gamma scale 1, turbidity 2, intensity 100
    float3 sun          = s.getSunDir();
    float  sunElevation = std::asin(fminf(fmaxf(sun.y, -1.0f), 1.0f));
    ArHosekSkyModelState* st =
        arhosekskymodelstate_alienworld_alloc_init(sunElevation, dome.intensity(), 5777.0f, dome.turbidity(), 0.2f);
    for (int w = 0; w < 11; w++)
    {
        for (int k = 0; k < 9; k++) configs99[w * 9 + k] = st->configs[w][k];
        radiances11[w] = st->radiances[w];
        ecf_sky11[w]   = st->emission_correction_factor_sky[w];
    }
    arhosekskymodelstate_free(st);
    float3 wts[10];
    Tungsten::Spectral::spectralXyzWeights(10, lambdas10, wts);
    for (int i = 0; i < 10; i++) { weights30[3 * i] = wts[i].x; weights30[3 * i + 1] = wts[i].y; weights30[3 * i + 2] = wts[i].z; }
    float3 sp = s.sunColor() * 0.02f;
    sun_dir3[0] = sun.x; sun_dir3[1] = sun.y; sun_dir3[2] = sun.z;
    sun_power3[0] = sp.x; sun_power3[1] = sp.y; sun_power3[2] = sp.z;
    return 0;
}
