/*
 * dsnerf.h -- C ABI of libdsnerf.so, the B200 (sm_100a) volume-rendering path
 * for Dual-Space NeRF.
 *
 * The reference (zyhbili/Dual-Space-NeRF) is pure Python: its "operator
 * interface" for this path is the class can_render.Renderer
 * (can_render.py:14-406).  It has no FFI, so these entry points are what a
 * binding for that class has to call; each one names the reference code it
 * replaces.  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success or a negative dsnerf_status; the
 *    message for the last failure on a context is dsnerf_last_error(ctx);
 *    nothing throws or aborts;
 *  - all arrays are fp32, contiguous, row-major; "dev" pointers are CUDA device
 *    memory on the context's device, "host" pointers are ordinary host memory;
 *  - the caller owns every I/O buffer; the context owns workspace and the
 *    staged weights/meshes;
 *  - calls are asynchronous on `stream` (a cudaStream_t passed as void*, 0 =
 *    default stream) unless the name ends in _host; a context is bound to one
 *    device and is not thread-safe: use one context per rank/GPU;
 *  - near/far are never modified (the reference's in-place overwrite,
 *    utils/pts_utils.py:52-53, is an accident of its implementation).
 */
#ifndef DSNERF_H_
#define DSNERF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSNERF_ABI_VERSION 1

typedef struct dsnerf_ctx dsnerf_ctx;

typedef enum {
  DSNERF_OK = 0,
  DSNERF_ERR_INVALID = -1,  /* bad argument / call order */
  DSNERF_ERR_CUDA = -2,     /* CUDA runtime error (message has the details) */
  DSNERF_ERR_NO_DEVICE = -3,
  DSNERF_ERR_STATE = -4     /* weights / mesh / frame not set yet */
} dsnerf_status;

/* dsnerf_render flags */
#define DSNERF_SAMPLE_UNIFORM 0u   /* utils/pts_utils.py:3  uniform_sampling */
#define DSNERF_SAMPLE_GG 1u        /* utils/pts_utils.py:18 geometry_guided_ray_marching */
#define DSNERF_MLP_FP32_SIMT 2u    /* debug: evaluate the MLP with the fp32 SIMT kernel instead of tcgen05 */
#define DSNERF_EARLY_STOP 4u       /* optional early ray termination: the samples of a ray are evaluated front to back in four
                                    * waves and a ray whose transmittance has dropped to <= 1e-6 is not evaluated further (the
                                    * reference evaluates every sample; the remaining ones can change acc by < 1e-6, colour and
                                    * depth by < 4e-6).  OFF by default: the default path evaluates every non-transparent sample. */

/* number of tensors in DualSpaceNeRF.state_dict() (model/spacenet.py), order in SURVEY.md 8b */
#define DSNERF_NUM_WEIGHT_TENSORS 33

typedef struct {
  int64_t rays;               /* R of the last render */
  int64_t samples;            /* R*N nominal samples */
  int64_t evaluated_samples;  /* samples that were not transparent => went through the MLP */
  int64_t nn_candidates;      /* centroid distance evaluations of the last render (0 unless profiling is on) */
  double algorithmic_flop;    /* evaluated_samples * 1 804 544 */
  int32_t kernel_launches;    /* kernels launched by the last render call */
  int32_t reserved;           /* with profile bit 2: samples that needed an exact nearest-centroid search */
} dsnerf_stats_t;

int dsnerf_abi_version(void);

/* Renderer.__init__ (can_render.py:15-23): one context per GPU. */
int dsnerf_create(dsnerf_ctx** out, int device);
void dsnerf_destroy(dsnerf_ctx* ctx);
const char* dsnerf_last_error(const dsnerf_ctx* ctx);

/* render.net.load_state_dict(ckpt["model"]) (validate.py:27): 33 HOST pointers in
 * state_dict order (nerf.embedding.weight, nerf.stage1.0.weight, ... pose_mlp.4.bias);
 * nn.Linear layout [out,in] row-major.  Splits/pads/transposes once into the
 * device layouts the kernels use. */
int dsnerf_set_weights(dsnerf_ctx* ctx, const float* const* host_tensors, int n_tensors);

/* Renderer.load_body_model (can_render.py:382-406): faces (F,3) int32 and the
 * canonical (X-pose) vertices (V,3), HOST pointers.  Builds the canonical-space
 * nearest-centroid grid used by normal_local2world (model/spacenet.py:278-298). */
int dsnerf_set_mesh(dsnerf_ctx* ctx, const int32_t* faces, int n_faces, const float* canonical_verts, int n_verts);

/* Per-frame state read from the reference's batch dict + model switches:
 *   posed_verts (V,3) host   batch["xyz"]                          (can_render.py:353)
 *   poses (24,3) host        batch["poses"], joint 0 ignored       (model/spacenet.py:223)
 *   frame                    batch["frame"], row of the 500x8 code table
 *   zero_code != 0           render.net.nerf.w = 0                 (model/spacenet.py:126-129)
 *   light_shift (3) or NULL  light_center - mean(batch["Th"])      (model/spacenet.py:260-263)
 *   rot (2,2)+rot_center (2) or NULL   set_rot / set_rot_center    (model/spacenet.py:254-258)
 * Computes the pose feature (pose_mlp), folds code+pose into the first layer's
 * bias, uploads the posed mesh and rebuilds its nearest-centroid grid. */
int dsnerf_set_frame(dsnerf_ctx* ctx, const float* posed_verts, const float* poses, int frame, int zero_code,
                     const float* light_shift, const float* rot, const float* rot_center, void* stream);

/* Renderer.render / batchify_rays_view (can_render.py:137-168, 172-245), eval mode:
 * sample -> warp -> SpaceNet + density-gradient normal + lighting -> raw2outputs.
 * DEVICE pointers: ray_o, ray_d (R,3); near, far (R); outputs rgb (R,3), depth,
 * acc, disp (R); optional weights, z_vals (R,N) (NULL to skip). */
int dsnerf_render(dsnerf_ctx* ctx, const float* ray_o, const float* ray_d, const float* near, const float* far,
                  int64_t n_rays, int n_samples, unsigned flags, float* rgb, float* depth, float* acc, float* disp,
                  float* weights, float* z_vals, void* stream);

/* Renderer.render with net.training == True (can_render.py:26-31, 105-108), forward only: the stratified jitter of
 * uniform_sampling (utils/pts_utils.py:6-13, cfg.MODEL.perturb > 0) and the density noise of raw2outputs
 * (utils/nerf_net_utils.py:29-33, cfg.MODEL.raw_noise_std > 0).  The reference draws both from torch's global generator;
 * here the draws are inputs, DEVICE (R,N): jitter = torch.rand (NULL: perturb off), raw_noise = torch.randn * raw_noise_std
 * (NULL: noise off).  With noise a transparent sample has weight relu(noise)-dependent > 0 and its colour counts, so the
 * network runs on every sample as in the reference.  z_vals is required when jitter is given; near/far are not modified. */
int dsnerf_render_train(dsnerf_ctx* ctx, const float* ray_o, const float* ray_d, const float* near, const float* far,
                        int64_t n_rays, int n_samples, unsigned flags, const float* jitter, const float* raw_noise,
                        float* rgb, float* depth, float* acc, float* disp, float* weights, float* z_vals, void* stream);

/* Fused render + all-gather over NVLink / NVSwitch (the one exchange step of the path; configs 4 and 5).
 * Same as dsnerf_render, but the compositor kernel itself stores the per-ray outputs into this rank's block
 *   [rgb (n_rays,3) | depth (n_rays) | acc (n_rays) | disp (n_rays)]   (6 * n_rays floats)
 * of the frame buffer of EVERY GPU of the group: `own_block` is the block inside this GPU's buffer, `peer_blocks[i]`
 * (i < n_peers <= 7) the same block inside peer i's buffer as a peer-mapped DEVICE pointer valid on this GPU (CUDA IPC /
 * symmetric memory), written with coalesced 128-byte stores while the kernel runs -- no separate collective, no staging
 * copy.  If `multicast_block` is not NULL it is the NVSwitch multicast (multimem) address of the block and one
 * multimem.st per value replaces the per-peer stores (own_block is still written locally; it may alias the multicast
 * target's local copy).  The caller synchronises the group afterwards (e.g. a symmetric-memory barrier on `stream`)
 * before any GPU reads another GPU's block.  dual_space_nerf_b200.dist.FrameExchange sets this up with
 * torch.distributed._symmetric_memory.  Does not combine with DSNERF_EARLY_STOP. */
int dsnerf_render_gather(dsnerf_ctx* ctx, const float* ray_o, const float* ray_d, const float* near, const float* far,
                         int64_t n_rays, int n_samples, unsigned flags, float* own_block, float* const* peer_blocks, int n_peers,
                         float* multicast_block, void* stream);

/* Same call with HOST buffers (pinned or pageable): copies inputs to the device,
 * renders, copies the outputs back and synchronises the stream.  This is the
 * entry point Renderer.render_view (can_render.py:248-278) maps to, and what
 * bench.py times as "e2e". */
int dsnerf_render_host(dsnerf_ctx* ctx, const float* ray_o, const float* ray_d, const float* near, const float* far,
                       int64_t n_rays, int n_samples, unsigned flags, float* rgb, float* depth, float* acc,
                       float* disp, float* weights, float* z_vals, void* stream);

/* Pipelined form of dsnerf_render_host for a stream of frames: returns as soon as the work is enqueued -- upload on the
 * context's copy stream, kernels on `stream`, read-back on a download stream behind them -- and writes a ticket;
 * dsnerf_wait(ticket) blocks until that frame's outputs are in the caller's HOST buffers.  Two frames may be in flight
 * (device staging is double-buffered; a third submission first waits for the oldest), so the read-back of frame k and
 * the upload of frame k + 2 overlap the kernels of frame k + 1.  Input host buffers must stay untouched until the frame's
 * kernels have consumed the upload, i.e. until dsnerf_wait of that ticket; output buffers until dsnerf_wait returns.
 * dsnerf_render_host = dsnerf_render_host_async + dsnerf_wait. */
int dsnerf_render_host_async(dsnerf_ctx* ctx, const float* ray_o, const float* ray_d, const float* near, const float* far,
                             int64_t n_rays, int n_samples, unsigned flags, float* rgb, float* depth, float* acc, float* disp,
                             float* weights, float* z_vals, void* stream, int* ticket);
int dsnerf_wait(dsnerf_ctx* ctx, int ticket);

/* Second pass of a hierarchical render on caller-supplied, sorted z (R,N) (DEVICE).
 * The reference's Renderer.resampling is undefined (can_render.py:213); see
 * DESIGN.md "Config 3". */
int dsnerf_render_z(dsnerf_ctx* ctx, const float* ray_o, const float* ray_d, const float* z_vals, int64_t n_rays,
                    int n_samples, unsigned flags, float* rgb, float* depth, float* acc, float* disp, float* weights,
                    void* stream);

/* Deterministic inverse-CDF resampling + merge (own spec, oracle/oracle.py:sample_pdf):
 * z_in, weights (R,N) -> z_out (R, N + n_importance) sorted.  DEVICE pointers. */
int dsnerf_resample(dsnerf_ctx* ctx, const float* z_in, const float* weights, int64_t n_rays, int n_samples,
                    int n_importance, float* z_out, void* stream);

/* utils/nerf_net_utils.py:5-56 raw2outputs (noise 0, white_bkgd False) as a
 * stand-alone op.  raw (R,N,4) = rgb + density, z_vals (R,N), ray_d (R,3); DEVICE. */
int dsnerf_composite(dsnerf_ctx* ctx, const float* raw, const float* z_vals, const float* ray_d, int64_t n_rays,
                     int n_samples, float* rgb, float* depth, float* acc, float* disp, float* weights, void* stream);

/* raw2outputs with raw_noise_std > 0 (utils/nerf_net_utils.py:29-33): raw_noise (R,N) = randn * raw_noise_std (DEVICE,
 * NULL = none) is added to the density before the ReLU. */
int dsnerf_composite_noise(dsnerf_ctx* ctx, const float* raw, const float* z_vals, const float* ray_d, const float* raw_noise,
                           int64_t n_rays, int n_samples, float* rgb, float* depth, float* acc, float* disp, float* weights,
                           void* stream);

/* Renderer.w2l_without_lbs (can_render.py:333-379) as a stand-alone op on the
 * current frame: pts (P,3) -> xyz_cano (P,3), transparent (P) uint8, idx (P) int32
 * (nearest posed-triangle index; NULL to skip).  DEVICE pointers. */
int dsnerf_warp_points(dsnerf_ctx* ctx, const float* pts, int64_t n_pts, float* xyz_cano, uint8_t* transparent,
                       int32_t* idx, void* stream);

/* Renderer.query_volume (can_render.py:280-296) / net(..., density_only=True)
 * (model/spacenet.py:238-241): canonical points (P,3) -> density (P); entries with
 * transparent[p] != 0 (may be NULL) are set to 0.  DEVICE pointers. */
int dsnerf_query_density(dsnerf_ctx* ctx, const float* xyz_cano, const uint8_t* transparent, int64_t n_pts,
                         float* density, unsigned flags, void* stream);

/* DualSpaceNeRF.forward (model/spacenet.py:210-266) on explicit points:
 * xyz_world, xyz_cano, view_dir (P,3) -> color (P,3), density (P).  DEVICE pointers. */
int dsnerf_eval_points(dsnerf_ctx* ctx, const float* xyz_world, const float* xyz_cano, const float* view_dir,
                       int64_t n_pts, float* color, float* density, unsigned flags, void* stream);

/* utils/blend_utils.py:72-81 ppts_to_pts (inverse linear-blend skinning; no caller inside the reference):
 * ppts (P,3) posed points, bw (24,P) blend weights, A (24,4,4) joint transforms -> out (P,3) = R^-1 (p - t) with
 * [R|t] = sum_j bw[j,p] A[j].  DEVICE pointers, one batch element per call.  A singular blend gives inf/NaN (torch.inverse
 * raises there).  Needs no weights / mesh / frame. */
int dsnerf_ppts_to_pts(dsnerf_ctx* ctx, const float* ppts, const float* bw, const float* A, int64_t n_pts, float* out, void* stream);

/* Camera rays on the device: utils/rays_utils.py:16-30 get_rays + :63-97 get_near_far (the inference branch of
 * my_sample_ray, :173-189).  K, R (3x3 row-major) and T (3) are HOST doubles, bounds HOST floats (min xyz, max xyz; the
 * 0.01 pad of get_near_far is applied here).  Outputs are DEVICE buffers over all H*W pixels in row-major pixel order:
 * ray_o, ray_d (H*W,3), near, far (H*W; 0 where the ray misses) and mask_at_box (H*W bytes).  The reference then keeps the
 * masked rays only (rays_utils.py:181-183): compact with the mask before calling dsnerf_render. */
int dsnerf_camera_rays(dsnerf_ctx* ctx, int H, int W, const double* K, const double* R, const double* T, const float* bounds,
                       float* ray_o, float* ray_d, float* near, float* far, uint8_t* mask_at_box, void* stream);

/* batch["transparent_mask"] of Renderer.render (can_render.py:156, get_transparent_mask utils/render_utils.py:103-109): the
 * per-sample mask of the LAST dsnerf_render* call on this context, which must have had exactly n_rays x n_samples samples:
 * transparent[r * n_samples + i] = 1 where the sample is transparent (density forced to 0), else 0.  DEVICE buffer of
 * n_rays * n_samples bytes; enqueued on `stream` (use the stream of the render call). */
int dsnerf_last_transparent_mask(dsnerf_ctx* ctx, int64_t n_rays, int n_samples, uint8_t* transparent, void* stream);

/* Precision mode dsnerf_set_weights chose for the staged weights: 1 = tcgen05 kernel, rgb head single-pass fp16 (enough for
 * weights of default-init scale); 3 = tcgen05 kernel with the 3-pass rgb head (a host-side probe of 64 points found the
 * single pass outside the colour budget: checkpoints whose activations reach O(1..10); +4 % kernel time); 0 = a weight is
 * outside fp16 range (|w| >= 60000): every evaluation runs on the fp32 CUDA kernel (slower); < 0 without weights. */
int dsnerf_tensor_path_active(const dsnerf_ctx* ctx);

/* Which tcgen05 kernel evaluates SpaceNet on this context: 2 = two 128-point tiles in flight per CTA (csrc/mlp_tc2.cuh: the
 * operands of both tiles share one FIFO ring of shared-memory slots, one TMEM accumulator per tile; the default), 1 = one tile
 * per CTA (csrc/mlp_tc.cuh; also serves density-only calls and is selected for everything with DSNERF_MLP_VARIANT=1 in the
 * environment at dsnerf_create time).  Both compute the same arithmetic: density and essence bit-identical, the gradient's
 * chain rule differs in the last bits. */
int dsnerf_mlp_kernel_variant(const dsnerf_ctx* ctx);

/* Counters of the last render on this context (synchronises the stream it ran on). */
int dsnerf_get_stats(dsnerf_ctx* ctx, dsnerf_stats_t* out);

/* Profiling switches (bit mask): 1 = CUDA-event timing of the MLP kernel inside render calls
 * (read back with dsnerf_profile_read: accumulated milliseconds / launches since the last
 * reset); 2 = count nearest-centroid distance evaluations (dsnerf_stats_t.nn_candidates; slows
 * the warp kernel, never enable it in a timed run); 4 = clock stamps of the tensor-core kernel; 8, 16 = debug switches of its
 * weight stream (garbage results); 64 / 128 = measurement switch: issue only 1 (64) or 2 (128) of the three MMAs of every forward
 * k-step (garbage results; bounds what a cheaper operand split could gain); 32 = test switch: shrink the candidate-list pool of meshes built from now on to 4096
 * entries so that most lookup cells take the ball-scan fallback (results must not change). */
int dsnerf_profile(dsnerf_ctx* ctx, int enable);
/* debug (profile bit 4): clock64 stamps of the tensor-core kernel's first tile; writes 128 values (the caller's buffer must
 * hold 128 long long: [0..63] epilogue stamps, [64..127] MMA-warp stamps) */
int dsnerf_debug_tc_timing(dsnerf_ctx* ctx, long long* out128);
int dsnerf_profile_read(dsnerf_ctx* ctx, double* mlp_ms, int64_t* mlp_launches, int reset);
/* debug: counters of the lazily built lookup table of the posed (which = 0) or canonical (1) mesh, 16 ints:
 * [0] candidate-list entries in use, [1] table cells / [2] enumeration cells requested by the last call, [3] table cells left
 * to the second build level, [11] table cells, [12] enumeration cells; with profile bit 2 also [4] far / settled through the
 * parent, [5] certified transparent, [6] cells with a list, [7] scan fallbacks, [8] list entries, [9]/[10] centroids visited
 * by the two build levels. */
int dsnerf_debug_table(dsnerf_ctx* ctx, int which, int* out16);
/* Test aid: the active list of the last dsnerf_render* call on this context, copied to HOST buffers: per evaluated
 * sample (x_cano, y_cano, z_cano, bits(sample id = ray * n_samples + i)) and the posed-space nearest triangle.
 * Order is not deterministic.  n_out receives the list length (may exceed capacity). */
int dsnerf_debug_active(dsnerf_ctx* ctx, int64_t capacity, float* active_xyz_id, int32_t* active_tri, int64_t* n_out);
/* Measurement aid: enqueue a one-warp kernel on `stream` that spins ~30 us and writes (SM MHz, microseconds) =
 * clock64 cycles per global-timer time to two DEVICE floats.  Lets a benchmark record the SM clock inside its timed
 * region without NVML queries (which stall kernel launches for tens of milliseconds on this driver). */
int dsnerf_debug_sm_clock(dsnerf_ctx* ctx, float* d_out2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DSNERF_H_ */
