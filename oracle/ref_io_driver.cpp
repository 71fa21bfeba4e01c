// oracle/ref_io_driver.cpp -- TEST INFRASTRUCTURE ONLY (never imported by the product).
// C entry points around the reference's OWN file-format code, compiled where it lies by oracle/build_ref.py:
//   * src/image.cpp (Image::dump_ppm / dump_hdr / scale / tonemap_gamma, image.cpp:20-209), against a stand-in for the
//     un-vendored GLM dependency that declares only what image.cpp uses of glm::vec4 (x, y, z, w, vec4(float), *=);
//   * loadBinaryFile and the uchar quantisation rule of loadVdbFile, extracted from src/volumeRender.cpp:915-965 and
//     :1003-1009 into ref_loader.inc.
// tests/golden/make_io_golden.py calls these to produce tests/golden/io_golden.npz.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "image.h"

typedef unsigned char VolumeType;  // src/volumeRender.cpp:98
#include "ref_loader.inc"

extern "C" {
// rgba: float[h][w][4], row 0 first (the layout of the float4 sum after scale())
int ref_dump_ppm(const float* rgba, int w, int h, const char* path)
{
    Image img(w, h);
    memcpy(img.buffer(), rgba, sizeof(float) * 4 * (size_t)w * h);
    img.dump_ppm(path);
    return 0;
}
int ref_dump_hdr(const float* rgba, int w, int h, const char* path)
{
    Image img(w, h);
    memcpy(img.buffer(), rgba, sizeof(float) * 4 * (size_t)w * h);
    img.dump_hdr(path);
    return 0;
}
int ref_tonemap_gamma(float* rgba, int w, int h, float scale, float gamma)
{
    Image img(w, h);
    memcpy(img.buffer(), rgba, sizeof(float) * 4 * (size_t)w * h);
    img.scale(scale);
    img.tonemap_gamma(gamma);
    memcpy(rgba, img.buffer(), sizeof(float) * 4 * (size_t)w * h);
    return 0;
}
// loadBinaryFile (volumeRender.cpp:915-965): returns the number of voxels, fills dims3 and out (uchar if quantized, else float)
long long ref_load_bin(const char* path, int* dims3, int quantized, void* out, long long out_bytes)
{
    int   w = 0, h = 0, d = 0;
    void* p = loadBinaryFile(const_cast<char*>(path), w, h, d, quantized != 0);
    if (!p) return -1;
    dims3[0] = w; dims3[1] = h; dims3[2] = d;
    long long total = (long long)w * h * d, bytes = total * (quantized ? 1 : 4);
    if (out && bytes <= out_bytes) memcpy(out, p, (size_t)bytes);
    free(p);
    return total;
}
// the uchar rule of loadVdbFile (volumeRender.cpp:1003-1009)
void ref_quantize_by_max(const float* dataf, long long total, float max_value, unsigned char* data)
{
    max_value = std::max(max_value, 0.0001f);  // volumeRender.cpp:977
    ref_quantize_by_max_loop(dataf, (size_t)total, max_value, data);
}
}
