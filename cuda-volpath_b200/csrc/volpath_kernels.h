// volpath_kernels.h -- host-callable launchers of the volpath CUDA kernels (internal header).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/volpath.h"

namespace vp
{
struct Scene;

// --- rendering (volpath_render_parity.cu, volpath_render_fast.cu, volpath_render_wave.cu) --------------
cudaError_t launch_render_parity(const Scene& S, float4* d_sum, int first_frame, int n_frames, int frame_stride,
                                 const vp_param& P, cudaStream_t stream);
// d_stats: 8 device counters {track fetches, shadow fetches, segments, opacity fetches, env evaluations,
// scatters, -, -} or nullptr (the uninstrumented kernel)
// d_work: one device counter (the work pool of the launch); num_sms sizes the persistent grid
cudaError_t launch_render_fast(const Scene& S, float4* d_sum, int first_frame, int n_frames, int frame_stride,
                               const vp_param& P, unsigned long long* d_work, unsigned long long* d_stats, int num_sms,
                               cudaStream_t stream);
// the wavefront form (warp-private ray-state pools in shared memory); same samples as launch_render_fast
cudaError_t launch_render_wave(const Scene& S, float4* d_sum, int first_frame, int n_frames, int frame_stride,
                               const vp_param& P, unsigned long long* d_work, unsigned long long* d_stats, int num_sms,
                               cudaStream_t stream);
cudaError_t launch_accumulate(float4* dst, const float4* src, int size, cudaStream_t stream);
cudaError_t launch_sum_peers(float4* dst, const float4* const* peers, int n, int size, cudaStream_t stream);
cudaError_t launch_resolve(float4* dst, const float4* src, int size, float scale, float gamma, cudaStream_t stream);

// --- scene build (volpath_build.cu) ---------------------------------------------------------------------
void        set_build_sm_count(int num_sms);  // grid sizing of the build kernels
cudaError_t launch_fbm_cloud(float* d_dense, int nx, int ny, int nz, uint32_t seed, cudaStream_t stream);
cudaError_t launch_u8_to_f32(const uint8_t* src, float* dst, size_t n, cudaStream_t stream);
// one axis of the separable clamped-window (max,min): in [n2][n1][n0] -> out with the swept axis reduced by
// `cell` (out index c covers inputs [c*cell - D, c*cell + cell - 1 + D])
// centred != 0: the window of the cell's centre voxel alone (c*cell + cell/2 +- D) instead of the union of its voxels' windows
cudaError_t launch_bounds_axis_f32(const float* in, float2* out, int n0, int n1, int n2, int axis, int D, int cell,
                                   cudaStream_t stream, int centred = 0);
cudaError_t launch_bounds_axis(const float2* in, float2* out, int n0, int n1, int n2, int axis, int D, int cell,
                               cudaStream_t stream, int centred = 0);
cudaError_t launch_merge_cell_bounds(float2* union_bounds, const float2* centre_bounds, size_t total, int max_from_union, cudaStream_t stream);
cudaError_t launch_classify_bricks(const float* dense, int nx, int ny, int nz, int nbx, int nby, int nbz, uint32_t* flags,
                                   cudaStream_t stream);
cudaError_t launch_make_words(const uint32_t* flags, const uint32_t* scan, size_t nb, uint2* words, uint32_t* slot_brick,
                              uint32_t* table, cudaStream_t stream);
cudaError_t launch_fill_octets(const float* dense, int nx, int ny, int nz, int nbx, int nby, const uint32_t* slot_brick,
                               uint32_t n_slots, void* pool, int voxel_type, cudaStream_t stream);
cudaError_t launch_precompute_opacity(const Scene& S, const uint32_t* slot_brick, uint32_t n_slots, float* opacity_bricks,
                                      float3 light_dir, cudaStream_t stream);
// production table: swept build (checkpoint slabs every K voxels along the sun's dominant axis), fp16 octets per cell
cudaError_t launch_opacity_octets(const Scene& S, const uint32_t* slot_brick, uint32_t n_slots, void* octets_f16, float3 light_dir,
                                  int K, cudaStream_t stream, uint32_t slot_begin = 0, uint32_t slot_end = 0xffffffffu);
cudaError_t launch_gather_opacity_oct(const Scene& S, float* dense_out, cudaStream_t stream);
cudaError_t launch_vacuum_jumps(float2* bounds_cell, uint8_t* tmp, int ncx, int ncy, int ncz, int kmax, int margin,
                                float cell_world, cudaStream_t stream);
cudaError_t launch_build_top(const float2* bounds_cell, int ncx, int ncy, int ncz, int top_log2, int ntx, int nty, int ntz, uint16_t* top,
                             cudaStream_t stream);
cudaError_t launch_sun_clear(const Scene& S, float3 sun, float step, int ring, float* out, cudaStream_t stream);
cudaError_t launch_bake_sunsky(const vp_sky_state& st, float4* env, int width, int height, cudaStream_t stream);
cudaError_t launch_pack_bounds_half(const float2* bounds_cell, uint32_t* out, size_t total, int* d_overflow, cudaStream_t stream);
cudaError_t launch_pack_clear_half(const float* sun_clear, uint16_t* out, size_t total, cudaStream_t stream);
cudaError_t launch_expand_cell_bounds(const Scene& S, float2* bounds_voxel, cudaStream_t stream);
cudaError_t launch_gather_opacity(const Scene& S, float* dense_out, cudaStream_t stream);

// --- probes for tests -------------------------------------------------------------------------------------
cudaError_t launch_fetch_density(const Scene& S, const float3* pos, int n, int parity, float* out, cudaStream_t stream);
cudaError_t launch_rng_sequence(uint32_t x, uint32_t y, uint32_t frame, int n, float* out_f, uint32_t* out_u,
                                cudaStream_t stream);
cudaError_t launch_philox(uint32_t c0, uint32_t c1, uint32_t key, uint32_t* out2, cudaStream_t stream);
}  // namespace vp
