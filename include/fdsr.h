/* libfdsr — B200 (sm_100a) implementation of the FastDiffSR T-step conditional sampling path.
 *
 * C ABI, no C++ / torch types.  The reference is pure Python and has no native interface to
 * mirror, so every entry point below cites the reference code it replaces (paths relative to
 * FastDiffSR/ in the reference tree).
 *
 * Conventions
 *   - every function returns 0 on success or a negative FDSR_E_* code; nothing throws;
 *     fdsr_last_error() returns a human-readable message for the last failure on that context
 *   - "dev" pointers are CUDA device pointers owned by the caller, valid until the stream work
 *     enqueued by the call has completed; "host" pointers are ordinary host memory
 *   - all image tensors are fp32 NCHW (B,3,H,W), the reference's layout; H and W must be
 *     multiples of 2^(n_levels-1): 8 for the FastDiffSR UNet (three stride-2 levels,
 *     model/fastdiffsr_modules/unet.py:77-83), 32 for the SR3 baseline (five)
 *   - calls are asynchronous with respect to the host (enqueue on `stream` and return) unless
 *     stated; a context is single-owner, not thread-safe, bound to the device that was current
 *     at fdsr_create
 */
#ifndef FDSR_H_
#define FDSR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDSR_OK 0
#define FDSR_E_INVALID -1   /* bad argument / unsupported configuration */
#define FDSR_E_CUDA -2      /* a CUDA runtime call failed */
#define FDSR_E_STATE -3     /* call order violated (weights / schedule not loaded) */
#define FDSR_E_NOTFOUND -4  /* unknown tensor / layer name */
#define FDSR_E_OVERFLOW -5  /* fp16 mode: an activation left the fp16 range (stored saturated); use FDSR_DTYPE_BF16 */

#define FDSR_DTYPE_FP16 0   /* fp16 operands + activations, fp32 accumulate (tcgen05 kind::f16).  Activation stores
                               saturate at +-65504 and raise an overflow flag (fdsr_check_overflow) */
#define FDSR_DTYPE_BF16 1   /* bf16 activations (fp32 exponent range: safe for a trained network's un-normalised residual
                               stream), fp32 accumulate.  Operands: swish(GroupNorm(x)) -- bounded -- is rounded to fp16
                               and multiplied with fp16 weights; raw inputs (1x1 residual / up / down convs) bf16 x bf16 */
#define FDSR_DTYPE_FP32 2   /* parity mode: fp32 activations, weights and accumulation on the CUDA cores (same fused
                               layer plan; eps relative L2 <= 1e-4 against the reference; ~60x slower) */

#define FDSR_MAX_LEVELS 8

#define FDSR_MODEL_FASTDIFFSR 0 /* which_model_G = "fastdiffsr": model/fastdiffsr_modules (noise-level FiLM, CLAM/SLAM) */
#define FDSR_MODEL_SR3 1        /* which_model_G = "ddpm": the SR3 comparison baseline, model/ddpm_modules (time
                                   embedding, SelfAttention); it predicts the image itself (no res2img) */

typedef struct fdsr_ctx fdsr_ctx;

/* UNet hyper-parameters: opt['model']['unet'] consumed by define_G (model/networks.py:82-119)
 * and UNet.__init__ (model/fastdiffsr_modules/unet.py:224-297). */
typedef struct fdsr_config {
  int32_t in_channel;     /* 6: cat[cond, x_t] */
  int32_t out_channel;    /* 3 */
  int32_t inner_channel;  /* 64 */
  int32_t norm_groups;    /* 32 */
  int32_t n_levels;       /* len(channel_multiplier) */
  int32_t channel_mults[FDSR_MAX_LEVELS];
  int32_t res_blocks;     /* 2 */
  int32_t dtype;          /* FDSR_DTYPE_* */
  int32_t model;          /* FDSR_MODEL_* (networks.py:84-91) */
  int32_t attn_levels;    /* FDSR_MODEL_SR3: bit l set = the ResnetBlocks of resolution level l carry SelfAttention,
                             i.e. (image_size >> l) is in attn_res (ddpm_modules/unet.py:184, 211); mid[0] always does */
} fdsr_config;

/* Replaces UNet.__init__ + GaussianDiffusion.__init__ (diffusion.py:80-99): builds the layer plan. */
int fdsr_create(const fdsr_config* cfg, fdsr_ctx** out);
int fdsr_destroy(fdsr_ctx* ctx);
const char* fdsr_last_error(const fdsr_ctx* ctx);
/* Library-level message for failures that have no context (fdsr_create). */
const char* fdsr_global_error(void);

/* Replaces netG.load_state_dict (model/model.py:148-160).  `names[i]` is the reference state_dict
 * key ("denoise_fn.downs.1.res_block.block1.block.3.weight", ...), `host_ptrs[i]` a contiguous
 * fp32 HOST array in the reference layout (conv OIHW, linear [out,in]), `numels[i]` its element
 * count (checked).  Keys of never-executed tensors (the dead 1x1 `.conv`, unet.py:212, and the 12
 * schedule buffers) are accepted and ignored.  Synchronous.  Repacks into the kernels' operand
 * layout; must be called before fdsr_set_schedule. */
int fdsr_load_weights(fdsr_ctx* ctx, const char* const* names, const float* const* host_ptrs,
                      const int64_t* numels, int32_t n);

/* The same from DEVICE memory (netG.to('cuda') weights need not bounce through the caller's host
 * memory): `dev_ptrs[i]` is a contiguous fp32 device array of numels[i] elements, borrowed for the duration of the
 * call; copies are ordered after prior work on `stream`.  Synchronous. */
int fdsr_load_weights_dev(fdsr_ctx* ctx, const char* const* names, const float* const* dev_ptrs,
                          const int64_t* numels, int32_t n, void* stream);

/* Replaces GaussianDiffusion.set_new_noise_schedule (diffusion.py:109-155): derives every table
 * from the T betas in float64 exactly as the reference does, and precomputes the per-step FiLM
 * bias vectors (noise level depends on t only, diffusion.py:169-170; unet.py:22-54, 242-248).
 * Synchronous.  Re-entrant (the reference calls it for 'train' and again for 'val'). */
int fdsr_set_schedule(fdsr_ctx* ctx, const double* betas, int32_t T);
/* Copies one derived table (name as registered by the reference: "betas", "alphas_cumprod",
 * ..., "posterior_mean_coef2", or "sqrt_alphas_cumprod_prev" which has T+1 entries) to host. */
int fdsr_get_table(fdsr_ctx* ctx, const char* name, double* out_host, int32_t capacity);

/* Sizes / (re)allocates the activation workspace for this shape.  Implicit in the calls below.  The allocation only
 * ever grows, so the captured graphs of previously used shapes stay valid.
 * Deviation from SURVEY 8(b), which sketched `size_t fdsr_workspace_bytes(ctx, B, H, W)` with a caller-provided
 * workspace: the context owns its workspace (captured CUDA graphs hold pointers into it), so sizing is split into
 * fdsr_reserve (allocate for a shape) + fdsr_workspace_bytes (bytes the current shape uses). */
int fdsr_reserve(fdsr_ctx* ctx, int32_t B, int32_t H, int32_t W);
size_t fdsr_workspace_bytes(const fdsr_ctx* ctx);

/* Replaces UNet.forward as called from p_mean_variance (diffusion.py:169-173): eps for step t,
 * noise_level = sqrt_alphas_cumprod_prev[t+1].  cond, x_t: (B,3,H,W) dev; eps_out: (B,3,H,W) dev. */
int fdsr_unet_forward(fdsr_ctx* ctx, const float* cond_dev, const float* xt_dev, int32_t t,
                      float* eps_out_dev, int32_t B, int32_t H, int32_t W, void* stream);

/* Replaces p_sample (diffusion.py:157-190) given eps: x0 = clamp(a_t x - b_t eps), mean, + sigma_t z.
 * z_dev may be NULL for t == 0.  In-place allowed (x_prev_dev == xt_dev). */
int fdsr_posterior_step(fdsr_ctx* ctx, const float* xt_dev, const float* eps_dev, const float* z_dev,
                        int32_t t, float* x_prev_dev, int64_t numel, void* stream);

/* Replaces GaussianDiffusion.super_resolution / p_sample_loop (diffusion.py:192-231), batch-safe.
 *   cond_dev  (B,3,H,W) bicubic conditioning in [-1,1]
 *   noise_dev (T,B,3,H,W) injected Gaussian noise in the reference's draw order (x_T, then z for
 *             t=T-1..1), or NULL to draw from the built-in counter-based generator with `seed`
 *   sr_out_dev (B,3,H,W) = clamp(x_0,-1,1)/2 + cond   (res2img, diffusion.py:275-281)
 *   trace_out_dev NULL, or (B, 1+n_frames, 3, H, W): per sample, res2img(cond) followed by
 *             res2img of x after every step with t % (1 | T/10) == 0 — the continous=True output. */
int fdsr_sample(fdsr_ctx* ctx, const float* cond_dev, const float* noise_dev, uint64_t seed,
                float* sr_out_dev, float* trace_out_dev, int32_t B, int32_t H, int32_t W,
                void* stream);
int32_t fdsr_trace_frames(const fdsr_ctx* ctx); /* 1 + n_frames for the current schedule */

/* fp16 mode only (always FDSR_OK otherwise): synchronises `stream`, then returns FDSR_E_OVERFLOW if any activation
 * stored since the last check had left the fp16 range (it was stored saturated, so nothing downstream is inf / NaN,
 * but the image is not the network's output), and clears the flag.  fdsr_super_resolve_u8 performs this check itself;
 * callers of the asynchronous fdsr_sample call it when they synchronise. */
int fdsr_check_overflow(fdsr_ctx* ctx, void* stream);

/* Position of subsequent batches in the job: `first_image` is the global index of image 0 of the next fdsr_sample /
 * fdsr_super_resolve_u8 batches.  The built-in generator's counter is (position in the image, GLOBAL image index,
 * step), so image k of a job receives the same noise whether it is sampled alone, inside a batch, or by another rank:
 * a batch sharded over N ranks (rank r passes the index of its first image) equals the single-rank result bit for bit.
 * Sticky; default 0.  Read from device memory by the captured graph (no re-capture). */
int fdsr_set_image_offset(fdsr_ctx* ctx, uint64_t first_image);

/* Host-buffer convenience for callers without device memory management (the end-to-end path that
 * bench.py times): uint8 LR (B,h,w,3) host -> PIL-exact bicubic -> sampling -> fp32 SR (B,3,H,W)
 * host.  Copies go through context-owned pinned staging buffers.  Synchronous. */
int fdsr_super_resolve_u8(fdsr_ctx* ctx, const uint8_t* lr_host, int32_t B, int32_t h, int32_t w,
                          int32_t H, int32_t W, const float* noise_dev, uint64_t seed,
                          float* sr_out_host, void* stream);

/* The same, pipelined for throughput: two slots.  submit enqueues H2D, bicubic, sampling and — on an internal copy stream —
 * the D2H of one batch into `slot` (0 or 1) and returns; wait blocks until that slot's result is in its pinned buffer,
 * copies it to sr_out_host and reports FDSR_E_OVERFLOW if the fp16 guard fired for that batch.  With one batch submitted
 * ahead, the copies and the host-side memcpy of batch i overlap the sampling of batch i+1. */
int fdsr_super_resolve_u8_submit(fdsr_ctx* ctx, int32_t slot, const uint8_t* lr_host, int32_t B, int32_t h, int32_t w,
                                 int32_t H, int32_t W, const float* noise_dev, uint64_t seed, void* stream);
int fdsr_super_resolve_u8_wait(fdsr_ctx* ctx, int32_t slot, float* sr_out_host);

/* New at the boundary (the reference precomputes it offline with PIL, data/prepare_data_mfe_dm.py
 * :30-40): bit-exact Pillow BICUBIC resize of uint8 HWC images (horizontal pass, uint8 rounding,
 * vertical pass; 22-bit fixed-point taps), then optional /255*2-1 to fp32 NCHW (data/util.py:66-75).
 *   lr_dev (B,h,w,3) u8;  out_u8_dev (B,H,W,3) u8 or NULL;  cond_out_dev (B,3,H,W) fp32 or NULL */
int fdsr_bicubic_u8(fdsr_ctx* ctx, const uint8_t* lr_dev, int32_t B, int32_t h, int32_t w,
                    int32_t H, int32_t W, uint8_t* out_u8_dev, float* cond_out_dev, void* stream);

/* PSNR accumulators for the sharded evaluation loop (sr_mfe.py:315-356, core/metrics.py:16-42,
 * 94-101): per image, quantise both tensors like tensor2img (clamp [-1,1] -> uint8) and add the
 * squared error sum to sse_out_dev[b] (double, B entries).  PSNR = 10 log10(255^2 * 3HW / sse). */
int fdsr_sse_u8(fdsr_ctx* ctx, const float* a_dev, const float* b_dev, int32_t B, int32_t H,
                int32_t W, double* sse_out_dev, void* stream);

/* Device-side evaluation metrics replacing the host block of sr_mfe.py:315-356 (skimage compare_mse /
 * compare_psnr / compare_ssim(multichannel=True) and core/metrics.py:88-93 calculate_ergas on
 * Metrics.tensor2img uint8 images, core/metrics.py:16-42).  a = the evaluated image (SR or bicubic),
 * b = HR, both (B,3,H,W) fp32 in [-1,1] on the device.  out_dev: B x 4 doubles (mse, psnr, ssim, ergas).
 * SSIM: 7x7 uniform window, sample covariance, K1 = .01, K2 = .03, data_range 255, 3-pixel crop. */
int fdsr_metrics_u8(fdsr_ctx* ctx, const float* a_dev, const float* b_dev, int32_t B, int32_t H, int32_t W,
                    double ergas_scale, double* out_dev, void* stream);

/* ---- test / profiling hooks (not part of the drop-in surface) ---- */
/* Number of activation tensors of the current plan and their names ("downs.0", "downs.1.h", ...). */
int32_t fdsr_debug_num_tensors(const fdsr_ctx* ctx);
const char* fdsr_debug_tensor_name(const fdsr_ctx* ctx, int32_t i);
/* Converts activation tensor `name` of the last fdsr_unet_forward (NHWC 16-bit) to fp32 NCHW. */
int fdsr_debug_read_tensor(fdsr_ctx* ctx, const char* name, float* out_dev, int64_t capacity,
                           int32_t* C, int32_t* H, int32_t* W, void* stream);
/* Per-op profiling of one UNet evaluation on the context's current buffers (call after a forward
 * or sample at the shape of interest): runs the ops `reps` times with a CUDA-event pair around
 * each op's launch(es) on `stream`, writes the mean milliseconds per op to ms_out_host (for reps >= 4 the
 * mean of the last reps/2 repetitions: the first half warms the GPU up to its sustained clocks) and returns
 * the number of ops.  Synchronous.  Op names / algorithmic conv FLOPs (at the reserved shape) for
 * turning the times into TFLOP/s: */
int32_t fdsr_debug_num_ops(const fdsr_ctx* ctx);
const char* fdsr_debug_op_name(const fdsr_ctx* ctx, int32_t i);
double fdsr_debug_op_flops(const fdsr_ctx* ctx, int32_t i);
/* The multiply-adds the op really executes: equal to the algorithmic figure except for the three nearest-upsample convs,
 * which run as four 2x2 phase convs on the low-resolution input (4/9 of the nine-tap MACs). */
double fdsr_debug_op_flops_executed(const fdsr_ctx* ctx, int32_t i);
int fdsr_debug_profile_unet(fdsr_ctx* ctx, int32_t t, int32_t reps, float* ms_out_host, int32_t cap,
                            void* stream);
/* Role-level cycle counters of one conv op (only populated by -DFDSR_PROFILE builds of the library,
 * used by tools/role_profile.py): out_host[(cta*5 + role)*8 + slot], role 0 = MMA issuer, 1 = epilogue,
 * 2 = producer, 3 = timeline (cycles since kernel entry of: prologue done, previous launch complete, GroupNorm table
 * built, first patch landed, first MMA, last MMA issued, epilogue done, exit), 4 = anchor {SM id, the SM's clock at
 * kernel entry, %globaltimer at entry}.  Re-runs op `op` on the current buffers; returns the number of entries.
 * Synchronous. */
int fdsr_debug_role_cycles(fdsr_ctx* ctx, int32_t op, int32_t t, int64_t* out_host, int32_t cap, void* stream);
/* The same counters of EVERY conv op inside a running UNet evaluation (tools/timeline.py): `reps` evaluations back to
 * back, launched as the sampler launches them (programmatic dependent launch); out_host[op][sm][role][slot] of the
 * last one (cap >= num_ops * num_sms * 40 entries).  Consecutive launches on one SM share its clock, so
 * first-MMA(op k+1) - last-MMA(op k) is the tensor pipe's idle time between two layers.  Returns the number of ops. */
int fdsr_debug_timeline(fdsr_ctx* ctx, int32_t t, int32_t reps, int64_t* out_host, int64_t cap, void* stream);
/* Fills out_dev (B,3,H,W) with the N(0,1) values the built-in generator hands the sampler for step stream
 * `stream_id` (T for x_T, t for the z of step t) of images first_image .. first_image+B-1 under `seed`: lets tests
 * check the distribution (moments, Kolmogorov-Smirnov) and reproduce a built-in-noise sample through the
 * injected-noise path bit for bit. */
int fdsr_debug_noise(fdsr_ctx* ctx, float* out_dev, int32_t B, int32_t H, int32_t W, uint64_t seed,
                     uint64_t first_image, int32_t stream_id, void* stream);
/* Number of CUDA-graph captures of the sampling loop so far (cache misses of the per-shape graph cache). */
int64_t fdsr_graph_captures(const fdsr_ctx* ctx);
/* Kernel launches enqueued by this context so far (library kernels only). */
int64_t fdsr_launch_count(const fdsr_ctx* ctx);
/* Algorithmic conv FLOPs (2*MAC, padding counted) of one UNet forward at the reserved shape. */
double fdsr_unet_flops(const fdsr_ctx* ctx);
/* 1 = capture the T-step loop into a CUDA graph per shape (default), 0 = plain stream launches. */
int fdsr_set_use_graph(fdsr_ctx* ctx, int32_t enable);

#ifdef __cplusplus
}
#endif
#endif /* FDSR_H_ */
