/* include/volpath.h -- C ABI of libvolpath_b200.so, the B200-native replacement of CUDA-volpath's
 * render hot path.
 *
 * Two layers, both plain C (pointers, ints, floats; no torch / C++ types):
 *
 *  (1) reference-named shims: the 14 `extern "C"` entry points the reference's host code binds
 *      (declared in src/volumeRender.cpp:117-128, 347-356; defined in src/volumeRender_kernel.cu),
 *      same names, same argument meaning, same ABI.  A maintainer links this library instead of
 *      compiling volumeRender_kernel.cu and nothing above the boundary changes.
 *      They operate on one implicit default context on the current CUDA device, like the reference's
 *      file-scope statics, but report failures through vp_last_error() instead of exit(1).
 *
 *  (2) the handle-based core the shims are built on (vp_*): one context per GPU, explicit stream,
 *      int status returns (0 = ok, otherwise a cudaError_t value or a VP_ERR_* code).
 *
 * Layout conventions are the reference's: volumes are dense, x fastest (index i + j*nx + k*nx*ny,
 * vdbloader/load_vdb.cpp:47-50), the accumulator is a device float4[W*H] SUM owned by the caller
 * (CudaFrameBuffer, src/volumeRender.cpp:358-389), .w accumulates the scatter count, frame index
 * `spp` is the RNG stream id (src/sampler.h:35-43).
 */
#ifndef VOLPATH_B200_H
#define VOLPATH_B200_H

#include <stdbool.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden: only this ABI is exported */
#endif

/* ---- plain-C mirrors of the CUDA vector types that cross the reference boundary ------------- */
typedef struct vp_float3 { float x, y, z; } vp_float3;
typedef struct vp_float4 { float x, y, z, w; } vp_float4;              /* 16-byte aligned in CUDA; device data */
typedef struct vp_dim3 { unsigned int x, y, z; } vp_dim3;              /* = dim3 */
typedef struct vp_extent { size_t width, height, depth; } vp_extent;  /* = cudaExtent */

/* src/param.h:4-12 -- 44-byte POD, passed by value to the reference's kernels */
typedef struct vp_param {
    unsigned int width, height;
    float        density, brightness;
    vp_float3    albedo;
    float        g;
    vp_float3    sigma_t;
} vp_param;

typedef struct vp_context vp_context;
typedef void*             vp_stream; /* cudaStream_t */

enum { VP_OK = 0, VP_ERR_INVALID = 10001, VP_ERR_NO_VOLUME = 10002, VP_ERR_NO_DEVICE = 10003, VP_ERR_UNSUPPORTED = 10004 };

/* voxel types: what the caller hands in (src) and what is stored in HBM (store) */
enum { VP_VOXEL_U8 = 0, VP_VOXEL_F16 = 1, VP_VOXEL_F32 = 2 };
enum { VP_MEM_HOST = 0, VP_MEM_DEVICE = 1 };

/* local-bound ("majorant") grids built at upload time
 *   VP_BOUNDS_VOXEL : per-voxel (max,min) over the clamped +-D voxel cube, D = ceil(0.05/(2/nx)) --
 *                     bit-identical to the reference's compute_volume_value_bound_
 *                     (src/volumeRender.cpp:1089-1267); needed by the parity renderer.
 *   VP_BOUNDS_CELL  : the grid the production renderers read: cells of c^3 voxels, c = 1 (the reference's own per-voxel
 *                     windows) up to 128 Mi voxels; above that the largest power of two <= max(1, D/6), at most 8, so
 *                     that the grid stays ~100 MB at any resolution.  A coarse cell holds the reference window of its
 *                     CENTRE voxel (centre +-D: the window size of the reference, shifted by at most c/2 voxels); whether
 *                     it is vacuum -- skippable without random draws -- is decided by the conservative union of its
 *                     voxels' windows.  Measured against the reference-faithful renderer on the full 1987x1351x2449 grid
 *                     (c = 8, D = 50): scatter count -0.12 %, image mean 9e-5 (union windows: -0.96 % / 5e-4; the
 *                     reference estimator is biased by construction and its expectation moves with the window size,
 *                     DESIGN.md section 2).
 *   VP_BOUNDS_EXACT : VP_BOUNDS_CELL with c forced to 1: the fast renderer sees exactly the reference's windows
 *                     (8 B/voxel).
 * The flags can be or-ed. */
enum { VP_BOUNDS_VOXEL = 1, VP_BOUNDS_CELL = 2, VP_BOUNDS_EXACT = 4 };

/* render modes of vp_render
 *   VP_MODE_PARITY : one thread per pixel, reference RNG (Wang hash + xoroshiro64*), reference draw
 *                    order and segmenting; observationally identical to render_kernel.
 *   VP_MODE_FAST   : megakernel: persistent warps on one global work pool, warp-level event binning, Philox2x32-10
 *                    counter RNG, exact skipping of the empty-space march and of vacuum; same estimator in
 *                    distribution.
 *   VP_MODE_WAVE   : the wavefront form of VP_MODE_FAST: ray states as SoA pools in shared memory, batches of 32
 *                    same-event states compacted with ballot/popc; the SAME samples as VP_MODE_FAST (env-map
 *                    importance sampling included). */
enum { VP_MODE_PARITY = 0, VP_MODE_FAST = 1, VP_MODE_WAVE = 2 };

/* ---- (2) handle-based core -------------------------------------------------------------------- */
const char* vp_last_error(void);                 /* message of the last failing call on this thread */
const char* vp_version(void);
int vp_create(int device, vp_context** out);     /* one context per GPU */
int vp_destroy(vp_context* ctx);

/* replaces init_cuda (K.cu:354-420): dense volume in, bricked "octet" store + bound grids out.
 * boxmin/boxmax may be NULL -> +-(1, ny/nx, nz/nx) like the reference (K.cu:373-378). */
int vp_upload_volume(vp_context* ctx, const void* volume, int nx, int ny, int nz, int src_voxel, int store_voxel,
                     int memspace, const float* boxmin3, const float* boxmax3, int bounds_flags);
/* synthetic fBm cloud generated on the device, bit-identical to
 * oracle vo_fbm_cloud_f32; the dense fp32 copy is kept until vp_release_dense() when keep_dense != 0 */
int vp_generate_cloud(vp_context* ctx, int nx, int ny, int nz, unsigned int seed, int store_voxel,
                      const float* boxmin3, const float* boxmax3, int bounds_flags, int keep_dense);
const void* vp_dense_volume(vp_context* ctx);    /* device fp32 dense copy, or NULL */
int vp_release_dense(vp_context* ctx);
/* config C1: the reference's no-OpenVDB build -- procedural Julia set, box [-1,1]^3, bounds (1,0) */
int vp_set_julia(vp_context* ctx);
int vp_set_filter(vp_context* ctx, int linear);                                   /* set_texture_filter_mode */
int vp_set_envmap(vp_context* ctx, const float* rgba, int width, int height);     /* init_envmap, host ptr */
int vp_set_sun(vp_context* ctx, const float* dir3, const float* power3);          /* set_sun */
/* Sun/sky bake on the device: the per-texel loop of update_sunsky(baked = true) (volumeRender.cpp:296-322) --
 * Skydome::skyColor (sunsky/sky_tungsten.cpp:400-419) over arhosekskymodel_radiance (sunsky/hosek/ArHosekSkyModel.cpp:
 * 519-561) -- evaluated by one kernel straight into the context's environment map (replaces the host loop + init_envmap;
 * no 8 MB upload per sun change).  The caller passes the small host-side state the reference holds after
 * Skydome::prepareForRender(): the cooked Hosek configurations stay host code. */
typedef struct vp_sky_state
{
    double configs[11][9];                      /* ArHosekSkyModelState::configs */
    double radiances[11];                       /* ::radiances */
    double emission_correction_factor_sky[11];  /* ::emission_correction_factor_sky */
    float  lambdas[7];                          /* Spectral::spectralXyzWeights, the NumSamplesValid = 7 used ones */
    float  weights[7][3];
    float  gamma_scale;                         /* Skydome::_gammaScale (1) */
    float  sun_dir[3];                          /* Skydome::sunDirection() */
    float  ground_rgb[3];                       /* lower half: ground_albedo * sun_dir.y * sun_power * pi (0.45/94)^2 */
    float  sunsky_scale;                        /* 0.02 (volumeRender.cpp:292) */
} vp_sky_state;
int vp_bake_sunsky(vp_context* ctx, const vp_sky_state* state, int width, int height);
int vp_get_envmap(vp_context* ctx, float* h_out_rgba, int* wh2);  /* introspection: the environment map in use */

/* the reference's PASSIVE_ENVMAP switch (K.cu:21), a compile-time macro there: 0 (default, as shipped) = the environment
 * is picked up by escaping paths; enable != 0 = env-map importance sampling + one-sample MIS with phase sampling at
 * every scatter event (K.cu:904-1034, 2220-2297); the CDF tables are built like init_envmap builds them */
int vp_set_env_sampling(vp_context* ctx, int enable);
int vp_set_inv_view(vp_context* ctx, const float* m12);                           /* copy_inv_view_matrix */
int vp_precompute_opacity(vp_context* ctx, const float* light_dir3);              /* precompute_opacity */
int vp_free_volume(vp_context* ctx);                                              /* free_cuda_buffers */

/* replaces render_kernel + the host frame loop: frames first_frame, first_frame+frame_stride, ...
 * (n_frames of them) are added into d_sum.  n_frames = 1, stride 1, VP_MODE_PARITY is
 * observationally identical to one render_kernel launch.  Asynchronous on `stream`; any number of caller streams may
 * have launches of one context in flight (each launch owns a work-pool counter; beyond 16 launches in flight a launch
 * queues behind the one whose counter it takes over).  Scene setters (vp_set_sun, vp_set_envmap, vp_precompute_opacity,
 * uploads) synchronise the device before they replace a table a render may still be reading. */
int vp_render(vp_context* ctx, void* d_sum_float4, int first_frame, int n_frames, int frame_stride,
              const vp_param* p, int mode, vp_stream stream);
/* __scale / __gamma_correct (K.cu:2333-2362); gamma <= 0 -> plain scale */
int vp_resolve(vp_context* ctx, void* d_dst_float4, const void* d_src_float4, int size, float scale, float gamma,
               vp_stream stream);
/* the host-buffer form of the same call (what the reference's capture() does around its kernel,
 * src/volumeRender.cpp:585-610): h_sum is a HOST float4[W*H] sum; it is copied to the device, frames are
 * added, and the result is copied back.  Synchronous. */
int vp_render_to_host(vp_context* ctx, void* h_sum_float4, int first_frame, int n_frames, int frame_stride,
                      const vp_param* p, int mode);
/* d_dst[i] += d_src[i] (float4, i < size), both on ctx's device: combines the per-GPU sums of a sample-sharded render in
 * single-process hosts (cudaMemcpyPeerAsync the peer sum next to it, then add); multi-process hosts reduce with NCCL */
int vp_accumulate(vp_context* ctx, void* d_dst_float4, const void* d_src_float4, int size, vp_stream stream);
int vp_sync(vp_context* ctx);

/* ---- combining the per-GPU sums of a sample-sharded render.  The reference has no multi-GPU path; a
 * path-sample is addressed by (x, y, frame) (src/sampler.h:35-43), so GPU r of G renders frames r, r + G, ... with
 * vp_render(first + r, count, stride = G) into its own float4[W*H] sum and the sums are added on one root.
 * NCCL (libnccl.so.2, bound at run time) over NVLink / NVSwitch; no torch involved.
 *   multi-process hosts (one process per GPU): rank 0 calls vp_nccl_unique_id, the host distributes the 128 bytes by
 *   its own means (MPI_Bcast, a file, a socket, torch's store), every rank calls vp_nccl_init on its context, then
 *   vp_reduce_nccl per image: ncclReduce(sum, float, 4 * size) on `stream` (d_recv may equal d_send; NULL off-root).
 *   single-process hosts (several contexts in one process): vp_reduce(ctxs, d_sums, n, size, root).  Synchronous. */
#define VP_NCCL_ID_BYTES 128
int vp_nccl_available(void);                       /* 0, or the NCCL version code of the library that was loaded */
int vp_nccl_unique_id(char* out128);
int vp_nccl_init(vp_context* ctx, int n_ranks, int rank, const char* id128);
int vp_reduce_nccl(vp_context* ctx, const void* d_send_float4, void* d_recv_float4, int size, int root, vp_stream stream);
int vp_nccl_destroy(vp_context* ctx);
int vp_reduce(vp_context** ctxs, void** d_sums_float4, int n, int size, int root);
/* The same reduce over PEER MEMORY, single node, no communicator and no bootstrap (NCCL's costs 4-10 s at 8 ranks): every
 * rank exports its accumulator -- a cudaMalloc / vp_dev_alloc BASE pointer -- with vp_ipc_export (64 bytes, carried to the
 * root by the host's own means); the root maps the peers through CUDA IPC (cached per context) and ONE kernel adds them
 * into its own sum in rank order, reading the peers over NVLink / NVSwitch.  The caller orders the processes: the peers'
 * renders are complete (their streams synchronised, then a host-level barrier) before the root calls vp_reduce_ipc, and
 * the peers leave their buffers alone until the root's stream has finished (a second barrier).  Mappings are cached per
 * context (keyed by the handle bytes; size 0 just maps) until vp_ipc_close or vp_destroy. */
#define VP_IPC_HANDLE_BYTES 64
int vp_ipc_export(vp_context* ctx, const void* d_base_ptr, char* out64);
int vp_reduce_ipc(vp_context* ctx, void* d_sum_float4, const char* peer_handles64, int n_peers, int size, vp_stream stream);
int vp_ipc_close(vp_context* ctx);
/* vp_precompute_opacity for a multi-process host whose ranks hold the same volume (after vp_nccl_init): rank r sweeps 1/G
 * of the production table and an in-place ncclAllGather completes it on every rank -- the one setup step that is worth
 * sharding (0.55 s on one B200 at the full C2 grid).  Collective: every rank must call it.  Without a communicator it is
 * vp_precompute_opacity. */
int vp_precompute_opacity_sharded(vp_context* ctx, const float* light_dir3);

/* introspection for tests / benchmarks */
int vp_get_bounds_voxel(vp_context* ctx, float* h_out_maxmin);   /* [nz][ny][nx][2], (max,min) */
int vp_get_bounds_cell(vp_context* ctx, float* h_out_maxmin, int* dims3); /* [cz][cy][cx][2] */
/* half-precision copies of the per-cell tables the production renderers read on large volumes: (max,min) as two
 * IEEE halves per cell and the sun-clear distance as one; *present = 0 when the float tables are in use */
int vp_get_half_tables(vp_context* ctx, unsigned short* h_out_maxmin, unsigned short* h_out_clear, float* h_out_clear_f32,
                       int* present);
int vp_get_opacity(vp_context* ctx, float* h_out);               /* [nz][ny][nx], 0 where not stored: the bit-faithful table */
int vp_get_opacity_fast(vp_context* ctx, float* h_out);          /* the production table (swept build, fp16 octets), same shape */
int vp_opacity_build_ms(vp_context* ctx, float* ms);             /* device time of the last vp_precompute_opacity */
int vp_fetch_density(vp_context* ctx, const float* h_pos3, int n, int parity_filter, float* h_out); /* world pos */
int vp_volume_stats(vp_context* ctx, unsigned long long* out8);  /* bricks, nonempty bricks, bytes ... */
int vp_rng_sequence(vp_context* ctx, unsigned int x, unsigned int y, unsigned int frame, int n, float* h_out_f,
                    unsigned int* h_out_u);                      /* reference RNG stream, from the device */
int vp_philox2x32(vp_context* ctx, unsigned int c0, unsigned int c1, unsigned int key, unsigned int* h_out2);
int vp_set_stats(vp_context* ctx, int enable);                   /* fast mode: run the counting kernel variant */
int vp_render_counters(vp_context* ctx, unsigned long long* out16, int reset); /* [0..5] {track fetches, shadow
                                  fetches, segments, opacity fetches, env evals, scatters}; [8..11] warp-level block
                                  executions {path, scatter, segment, step}; [12..15] active lanes in them */
int vp_last_kernel_ms(vp_context* ctx, float* ms);               /* CUDA-event time of the last vp_render */
int vp_launch_count(vp_context* ctx, unsigned long long* n);     /* kernels launched by this context */

/* plain device-memory helpers for C / ctypes callers without a CUDA runtime of their own */
void* vp_dev_alloc(size_t bytes);                                /* zero-filled */
int vp_dev_free(void* p);
int vp_dev_zero(void* p, size_t bytes);
int vp_dev_to_host(void* h, const void* d, size_t bytes);
int vp_host_to_dev(void* d, const void* h, size_t bytes);

/* ---- (1) reference-named shims (signatures: see the file:line next to each) ------------------- */
void init_cuda(void* h_volume, vp_extent volumeSize, bool quantized, const vp_float3* boxmin,
               const vp_float3* boxmax);                                 /* K.cu:354 */
void set_texture_filter_mode(bool bLinearFilter);                 /* K.cu:422 */
void free_cuda_buffers(void);                                             /* K.cu:441 */
void precompute_opacity(const float* light_dir);                          /* K.cu:526 */
void init_envmap(const vp_float4* HDRmap, int width, int height);         /* K.cu:1072 */
void free_envmap(void);                                                   /* K.cu:1231 */
void set_sun(float* sun_dir, float* sun_power);                           /* K.cu:1269 */
void copy_inv_view_matrix(float* invViewMatrix, size_t sizeofMatrix);     /* K.cu:2320 */
void copy_inv_model_matrix(float* invModelMatrix, size_t sizeofMatrix);   /* K.cu:2325 (USE_MODEL_TRANSFORM 0: stored, unused) */
void init_rng(vp_dim3 gridSize, vp_dim3 blockSize, int width, int height);/* K.cu:2330 (empty in the reference) */
void free_rng(void);                                                      /* K.cu:2331 */
void scale(vp_float4* dst, vp_float4* src, int size, float scale);        /* K.cu:2343, device pointers */
void gamma_correct(vp_float4* dst, vp_float4* src, int size, float scale, float gamma); /* K.cu:2359 */
/* K.cu:2364: `const Param& p` in the reference; a C++ reference is a pointer at the ABI level */
void render_kernel(vp_dim3 gridSize, vp_dim3 blockSize, vp_float4* d_output, int spp, const vp_param* p);
/* which renderer the render_kernel shim uses.  Default VP_MODE_FAST (the same estimator in distribution, verified against
 * the reference kernel; environment VOLPATH_SHIM_MODE=parity|fast|wave overrides the default).  VP_MODE_PARITY reproduces
 * the reference kernel's fixed-seed traces (1e-5 relative) at 0.83x its speed (profiles/r2_parity_mode_vs_ref.json).
 * VP_MODE_FAST: consecutive one-frame launches rotate over four internal BLOCKING streams, so the
 * long tail of frame n overlaps frame n + 1 while the host does not synchronise.  Blocking streams order themselves against
 * the legacy default stream exactly like the reference's own launches, so a host that uses the default stream,
 * cudaMemcpy or cudaDeviceSynchronize (the reference host does, volumeRender.cpp:627-641) needs no change.  A host
 * compiled with --default-stream per-thread, or one that reads d_output from a cudaStreamNonBlocking stream, must call
 * vp_shim_sync() (or cudaDeviceSynchronize) before touching d_output -- or switch the ring off: VOLPATH_SHIM_OVERLAP=0. */
void vp_shim_set_mode(int mode);
int  vp_shim_sync(void);
vp_context* vp_shim_context(void);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* VOLPATH_B200_H */
