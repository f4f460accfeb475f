/*
 * fcwdm.h -- C-ABI of libfcwdm.so: hand-written sm_100a (B200) kernels for the fast-cwdm hot path.
 *
 * The reference (tsereda/fast-cwdm) has no native code and no FFI: every operator below replaces a chain
 * of PyTorch calls in the reference's Python (cited per function as path:line relative to the reference
 * root).  The binding a maintainer adds is the ctypes stub shown in INTEGRATION.md
 * (fast-cwdm_b200/fcwdm/native.py is that stub).
 *
 * Conventions
 *   - plain C types only: device pointers (void* / typed pointers), int64_t sizes and strides (in ELEMENTS),
 *     float scalars, `stream` = a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every function returns 0 on success or a negative FCWDM_ERR_* code and never throws;
 *     fcwdm_last_error() returns a thread-local description of the last failure;
 *   - nothing is allocated and nothing synchronises: the caller owns all memory, outputs are preallocated,
 *     work is enqueued on `stream`; functions are re-entrant per stream;
 *   - "planar" = the reference's NCDHW layout (W fastest); "cl" = channels-last NDHWC (C fastest), the
 *     layout the denoiser keeps internally in bf16;
 *   - Haar band order everywhere: LLL, LLH, LHL, LHH, HLL, HLH, HHL, HHH; letters are the filters on
 *     (D, H, W) (DWT_IDWT/DWT_IDWT_Functions.py:128-136).
 */
#ifndef FCWDM_H_
#define FCWDM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCWDM_VERSION 100 /* 0.1.0 */
#define FCWDM_GN_STAT_REPLICAS 16

enum fcwdm_dtype { FCWDM_F32 = 0, FCWDM_BF16 = 1 };

enum fcwdm_status {
    FCWDM_OK = 0,
    FCWDM_ERR_INVALID = -1,     /* bad argument (null pointer, odd size, negative dim ...) */
    FCWDM_ERR_UNSUPPORTED = -2, /* valid in the reference, not implemented by this kernel */
    FCWDM_ERR_CUDA = -3,        /* a CUDA runtime / driver call failed; see fcwdm_last_error() */
    FCWDM_ERR_ARCH = -4         /* device is not sm_100 */
};

int fcwdm_version(void);
const char* fcwdm_last_error(void);
/* Idempotent per device: checks compute capability 10.x, resolves the driver entry point used to encode
 * TMA descriptors and raises the dynamic shared-memory limit of the conv kernels. */
int fcwdm_init(int device);

/* ------------------------------------------------------------------------------------------------------
 * K1 / K2: 3-D Haar DWT / IDWT, planar (NCDHW) tensors.
 * Replaces DWT_3D.forward + DWTFunction_3D (DWT_IDWT/DWT_IDWT_layer.py:520-531,
 * DWT_IDWT_Functions.py:115-156) and IDWT_3D.forward + IDWTFunction_3D (layer.py:624-646,
 * Functions.py:159-208): 14 band-matrix matmuls become one 2x2x2 butterfly pass.  Each is the other's
 * autograd backward.
 *
 * x: (N, C, D, H, W) with spatial dims contiguous; element (n,c,d,h,w) at x[n*x_sn + c*x_sc + (d*H+h)*W+w].
 * bands: band b of (n,c) starts at out[n*o_sn + c*o_sc + b*o_sb], spatial (D/2,H/2,W/2) contiguous.
 *   8 separate tensors stacked as (8,N,C,d,h,w): o_sb = N*C*d*h*w, o_sn = C*d*h*w, o_sc = d*h*w.
 *   channel-concatenated (N, 8*C, d,h,w) as th.cat([LLL/3, ...], dim=1) with C == 1
 *   (scripts/sample.py:92-97, gaussian_diffusion.py:1131-1140): o_sb = d*h*w, o_sn = 8*d*h*w.
 * lll_scale multiplies the LLL band on output (DWT: 1 or 1/3) / on input (IDWT: 1 or 3).
 * D, H, W must be even (the reference's matrices floor odd sizes, layer.py:466-494; odd sizes are rejected
 * here with FCWDM_ERR_UNSUPPORTED).
 * ---------------------------------------------------------------------------------------------------- */
int fcwdm_dwt3d_fwd(const void* x, void* bands, int dtype, int64_t N, int64_t C, int64_t D, int64_t H,
                    int64_t W, int64_t x_sn, int64_t x_sc, int64_t o_sn, int64_t o_sc, int64_t o_sb,
                    float lll_scale, void* stream);
int fcwdm_idwt3d_fwd(const void* bands, void* y, int dtype, int64_t N, int64_t C, int64_t D, int64_t H,
                     int64_t W, int64_t b_sn, int64_t b_sc, int64_t b_sb, int64_t y_sn, int64_t y_sc,
                     float lll_scale, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Channels-last bf16 DWT / IDWT used inside the denoiser (Downsample / Upsample / WaveletDownsample,
 * guided_diffusion/wunet.py:40-145).
 * x: (N, D, H, W, C) bf16, voxel stride x_ld (>= C).  C % 8 == 0.
 * dwt: LLL -> lll[voxel*lll_ld + c] * lll_scale (+ lll_bias[n*bias_ld + c] if non-null: the timestep-embedding
 *      add that follows the down-sampling, wunet.py:262; bias_ld = row stride of the per-sample bias matrix);
 *      band b=1..7 -> hi[(b-1)*hi_sb + voxel*hi_ld + c] * hi_scale; hi == NULL skips them (x_upd branch,
 *      wunet.py:241, which discards the 7 bands).
 *      WaveletDownsample's cat(...)/3 (wunet.py:143-144): lll = buf, hi = buf + C, hi_sb = C,
 *      lll_ld = hi_ld = 8*C, both scales 1/3.
 * idwt: y[voxel*y_ld + c] = IDWT(lll*lll_scale, hi...) (+ bias[n*bias_ld + c]).
 * ---------------------------------------------------------------------------------------------------- */
int fcwdm_dwt3d_cl(const void* x, int64_t x_ld, void* lll, int64_t lll_ld, void* hi, int64_t hi_ld,
                   int64_t hi_sb, const float* lll_bias, int64_t bias_ld, int64_t N, int64_t D, int64_t H,
                   int64_t W, int64_t C, float lll_scale, float hi_scale, void* stream);
int fcwdm_idwt3d_cl(const void* lll, int64_t lll_ld, const void* hi, int64_t hi_ld, int64_t hi_sb,
                    void* y, int64_t y_ld, const float* bias, int64_t bias_ld, int64_t N, int64_t D, int64_t H,
                    int64_t W, int64_t C, float lll_scale, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * K3: fused reverse-diffusion step.  Replaces process_xstart + q_posterior_mean_variance + the sampling
 * line of p_sample (guided_diffusion/gaussian_diffusion.py:335-355, 244-267, 565-573) and the six
 * _extract_into_tensor gathers (:1246-1263) with one elementwise pass:
 *   x0   = model_out (START_X) or coef[t][3]*x_t - coef[t][4]*model_out (EPSILON, :390-395)
 *   x0   = DWT(clamp(IDWT(x0; LLL*3), 0, 1)); LLL/3           if clip_denoised
 *   mean = coef[t][0]*x0 + coef[t][1]*x_t
 *   x_prev = mean + coef[t][2]*noise        (coef[t][2] = exp(0.5*logvar[t]) * (t != 0), host-prepared)
 * coef: device float [T][5]; t: device int64 [N] (per-sample timestep, no host sync).
 * model_out: planar f32 (N,8,d,h,w) when mo_cl_ld == 0, else channels-last bf16 with voxel stride mo_cl_ld.
 * x_t, noise, x_prev, pred_xstart (optional): planar f32 (N,8,d,h,w).
 * x_prev_cl (optional): bf16 channels-last copy of x_prev written to x_prev_cl[voxel*xp_cl_ld + band], the
 * first 8 channels of the persistent denoiser input (replaces th.cat([x, cond]), :297).
 * x_prev may alias x_t (the update is local to one latent voxel, so the chain state is updated in place);
 * pred_xstart may be NULL.
 * ---------------------------------------------------------------------------------------------------- */
int fcwdm_p_sample_step(const void* model_out, int64_t mo_cl_ld, const float* x_t, const float* noise,
                        float* x_prev, float* pred_xstart, void* x_prev_cl, int64_t xp_cl_ld,
                        const float* coef, const int64_t* t, int64_t T, int64_t N, int64_t d, int64_t h,
                        int64_t w, int clip_denoised, int predict_xstart, void* stream);

/* q_sample (gaussian_diffusion.py:224-242): out = coef[t][0]*x_start + coef[t][1]*noise, any shape with
 * `per_sample` elements per batch entry; coef: device float [T][2]. */
int fcwdm_q_sample(const float* x_start, const float* noise, float* out, const float* coef,
                   const int64_t* t, int64_t T, int64_t N, int64_t per_sample, void* stream);

/* Final image-space step of scripts/sample.py:113-131: IDWT(LLL*3) of the (N,8,d,h,w) planar sample,
 * clamp to [0,1] (the two masked writes), zero where cond_1 == 0 (cond_1 may be NULL), written as planar
 * f32 (N,1,2d,2h,2w).  The [..., :155] crop is a view taken by the caller. */
int fcwdm_sample_to_image(const float* sample, const float* cond_1, float* image, int64_t N, int64_t d,
                          int64_t h, int64_t w, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Layout / precision converters at the denoiser boundary (the reference is NCDHW fp32 end to end).
 * planar f32 (N,C,D,H,W) -> cl bf16 written at dst[voxel*dst_ld + c] for c < C, and back.
 * ---------------------------------------------------------------------------------------------------- */
int fcwdm_planar_to_cl(const float* src, void* dst, int64_t dst_ld, int64_t N, int64_t C, int64_t S,
                       void* stream);
int fcwdm_cl_to_planar(const void* src, int64_t src_ld, float* dst, int64_t N, int64_t C, int64_t S,
                       void* stream);

/* ------------------------------------------------------------------------------------------------------
 * K4: GroupNorm32 (+SiLU).  Replaces nn.GroupNorm in fp32 + nn.SiLU (guided_diffusion/nn.py:17-19,
 * wunet.py:186-187,210-211,702-703).  x, y: cl bf16 (N, S voxels, C), voxel stride ld; statistics in fp32
 * per thread, fp64 across blocks.  stats: device double [N][FCWDM_GN_STAT_REPLICAS][G][2] (sum, sum of squares;
 * blocks spread their atomics over the replicas, `apply` adds them up), zeroed by fcwdm_groupnorm_stats itself.  C % 8 == 0, C % G == 0, (C/G) in {1,2,4,8,...}.
 * ---------------------------------------------------------------------------------------------------- */
int fcwdm_groupnorm_stats(const void* x, int64_t ld, double* stats, int64_t N, int64_t S, int64_t C,
                          int64_t G, void* stream);
int fcwdm_groupnorm_apply(const void* x, int64_t x_ld, void* y, int64_t y_ld, const double* stats,
                          const float* gamma, const float* beta, int64_t N, int64_t S, int64_t C, int64_t G,
                          float eps, int silu, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Timestep path: sinusoidal embedding (nn.py:103-121) and small dense layers (time_embed, emb_layers;
 * wunet.py:472-475, 203-206).  y[n][m] = act_out( b[m] + sum_k act_in(x[n][k]) * W[m][k] ), fp32,
 * act: 0 = identity, 1 = SiLU.
 * ---------------------------------------------------------------------------------------------------- */
int fcwdm_timestep_embedding(const int64_t* t, float* out, int64_t N, int64_t dim, float max_period,
                             void* stream);
/* Same for fractional timesteps: what the model sees with rescale_timesteps=True (t * 1000 / T as float,
 * gaussian_diffusion.py:417-420, respace.py:128-132). */
int fcwdm_timestep_embedding_f32(const float* t, float* out, int64_t N, int64_t dim, float max_period,
                                 void* stream);
int fcwdm_linear(const float* x, const float* W, const float* b, float* y, int64_t N, int64_t K, int64_t M,
                 int act_in, int act_out, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * K5: 3-D convolution, stride 1, "same" zero padding, kernel 3x3x3 or 1x1x1, as a tcgen05 implicit GEMM
 * (bf16 operands staged by TMA, fp32 accumulation in TMEM).  Replaces nn.Conv3d -> cuDNN
 * (guided_diffusion/nn.py:22-32; call sites wunet.py:139,188,213,220,483,704).
 *
 * fcwdm_conv3d_pack_weights: w (Cout, Cin, k, k, k) f32 device -> wp [k^3][Cout_p][Cin_p] bf16 with
 *   Cin_p = round_up(Cin, 64), Cout_p = round_up(Cout, 16), zero padded (query sizes with
 *   fcwdm_conv3d_packed_elems).
 * fcwdm_conv3d_fwd: x cl bf16 (N,D,H,W, x_ld >= Cin_p channels readable; channels Cin..Cin_p-1 must be
 *   finite, they meet zero weights), y cl bf16 (N,D,H,W,Cout) with voxel stride y_ld;
 *   y = conv(x, w) + bias[c] (+ chan_bias[n*cb_ld + c]: the timestep-embedding add, wunet.py:262)
 *       (+ residual[voxel*res_ld + c]: the ResBlock skip add, wunet.py:266, or `input_pyramid + h`, :759).
 *   gn_stats (optional): the epilogue also accumulates the GroupNorm statistics of y (sum, sum of squares of the
 *   stored bf16 values per (n, group), gn_groups groups of Cout/gn_groups channels) into a PRE-ZEROED
 *   double [N][FCWDM_GN_STAT_REPLICAS][gn_groups][2] that fcwdm_groupnorm_apply consumes directly, so the
 *   next GroupNorm needs no statistics pass over y.
 * ---------------------------------------------------------------------------------------------------- */
/* fcwdm_conv3d_gn_fwd: the same 3x3x3 convolution of SiLU(GroupNorm(x)) with the normalisation + activation applied to
 * the halo planes in shared memory by four extra warps (x is the RAW tensor; gn_in_stats = its statistics in the
 * fcwdm_groupnorm_stats layout): the GroupNorm-apply pass and its intermediate tensor (nn.py:17-19 + nn.SiLU,
 * wunet.py:186-187,210-211; unet.py:225-229,249-256) disappear.  C_in a multiple of 64 (<= 256), C_out >= 64. */
int64_t fcwdm_conv3d_packed_elems(int64_t Cout, int64_t Cin, int ksize);
int fcwdm_conv3d_pack_weights(const float* w, void* wp, int64_t Cout, int64_t Cin, int ksize, void* stream);
int fcwdm_conv3d_fwd(const void* x, int64_t x_ld, const void* wp, const float* bias, const float* chan_bias,
                     int64_t cb_ld, const void* residual, int64_t res_ld, void* y, int64_t y_ld, double* gn_stats,
                     int64_t gn_groups, int64_t N, int64_t D, int64_t H, int64_t W, int64_t Cin, int64_t Cout,
                     int ksize, void* stream);
int fcwdm_conv3d_gn_fwd(const void* x, int64_t x_ld, const void* wp, const float* bias, const float* chan_bias,
                        int64_t cb_ld, const void* residual, int64_t res_ld, void* y, int64_t y_ld, double* gn_stats,
                        int64_t gn_groups, const double* gn_in_stats, const float* gn_in_gamma, const float* gn_in_beta,
                        int64_t gn_in_groups, float gn_in_eps, int64_t N, int64_t D, int64_t H, int64_t W, int64_t Cin,
                        int64_t Cout, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * K5c: a RUN of consecutive low-resolution 3x3x3 convolutions in ONE persistent launch (csrc/conv3d_chain.cu): the
 * layers of WavUNetModel at 28x28x20 voxels and below (wunet.py:533-609,615-675), where one layer has fewer output
 * tiles than the GPU has SMs and launch / fill / drain latency, not arithmetic, sets its time.  Layer l+1 may read
 * (as input, residual or GroupNorm statistics) anything layers <= l wrote: layers are separated by grid barriers.
 * Every layer is an fcwdm_conv3d_fwd / fcwdm_conv3d_gn_fwd with the same operand meaning (gn_in_stats == NULL: plain
 * input); C_in a multiple of 64, C_out a multiple of 128, weights packed by fcwdm_conv3d_pack_weights.
 * sync_counter: 4 bytes of device memory, ZERO at launch (the grid-barrier counter; left non-zero).
 * The launch needs every CTA co-resident (one per SM, clusters of 4): call it on a stream whose GPU it does not share
 * with another concurrently running fcwdm_conv3d_chain.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct fcwdm_chain_layer {
    const void* x;
    int64_t x_ld;
    const void* wp;
    const float* bias;
    const float* chan_bias;
    int64_t cb_ld;
    const void* residual;
    int64_t res_ld;
    void* y;
    int64_t y_ld;
    double* gn_stats;
    int64_t gn_groups;
    const double* gn_in_stats;
    const float* gn_in_gamma;
    const float* gn_in_beta;
    int64_t gn_in_groups;
    float gn_in_eps;
    int64_t N, D, H, W, Cin, Cout;
    /* kind != FCWDM_CHAIN_CONV: a wavelet re-sampling op between the convs of the run, executed by the same launch
     * (same operand meaning as fcwdm_dwt3d_cl / fcwdm_idwt3d_cl; wp, bias, residual, gn_in_* and Cout are ignored):
     *   FCWDM_CHAIN_DWT : x (N,D,H,W,Cin) -> y = LLL * lll_scale + chan_bias[n], aux = the 7 high bands * hi_scale (or NULL)
     *   FCWDM_CHAIN_IDWT: x = LLL (N,D/2,H/2,W/2,Cin) * lll_scale, aux = the 7 high bands -> y (N,D,H,W,Cin) + chan_bias[n]
     * band b of aux starts at aux + b * aux_sb (elements), voxel stride aux_ld; gn_stats (optional, pre-zeroed) receives
     * the GroupNorm statistics of y.  Cin must be 64, 128 or 256. */
    int64_t kind;
    void* aux;
    int64_t aux_ld, aux_sb;
    float lll_scale, hi_scale;
} fcwdm_chain_layer;
#define FCWDM_CHAIN_CONV 0
#define FCWDM_CHAIN_DWT 1
#define FCWDM_CHAIN_IDWT 2
int fcwdm_conv3d_chain_supported(int64_t Cin, int64_t Cout, int ksize);
int fcwdm_conv3d_chain_max_layers(void);
int fcwdm_conv3d_chain(const fcwdm_chain_layer* layers, int64_t n_layers, void* sync_counter, void* stream);
int fcwdm_debug_set_chain_trace(void* device_buffer);   /* development builds only (FCWDM_CONV_TRACE=1) */

/* ------------------------------------------------------------------------------------------------------
 * K5b: the same 3x3x3 convolution for C_in <= 64 and C_out <= 64 (the full-resolution layers) as a kd-fused,
 * two-CTA (cta_group::2) implicit GEMM with the weights resident in the CTA pair's shared memory
 * (csrc/conv3d_pair.cu).  Same semantics and epilogue options as fcwdm_conv3d_fwd; its own weight packing
 * [kh*3+kw][kd][C_out_p][64] (C_out_p = 16 or 64).  x must expose 64 readable channels per voxel (x_ld >= 64).
 * gn_in_stats (optional): x is the RAW tensor and the kernel convolves SiLU(GroupNorm(x)) instead, normalising and
 * activating in its operand producers from the statistics [N][FCWDM_GN_STAT_REPLICAS][gn_in_groups][2] of x
 * (as written by fcwdm_groupnorm_stats or a producing conv's gn_stats) and gamma/beta[C_in] -- the GroupNorm-apply
 * pass and its intermediate tensor (nn.py:17-19 + nn.SiLU, wunet.py:186-187,210-211,702-703) disappear.
 * ---------------------------------------------------------------------------------------------------- */
int fcwdm_conv3d_pair_supported(int64_t Cin, int64_t Cout, int ksize);
int64_t fcwdm_conv3d_pair_packed_elems(int64_t Cout, int64_t Cin);
int fcwdm_conv3d_pair_pack_weights(const float* w, void* wp, int64_t Cout, int64_t Cin, void* stream);
int fcwdm_conv3d_pair_fwd(const void* x, int64_t x_ld, const void* wp, const float* bias, const float* chan_bias,
                          int64_t cb_ld, const void* residual, int64_t res_ld, void* y, int64_t y_ld,
                          double* gn_stats, int64_t gn_groups, const double* gn_in_stats, const float* gn_in_gamma,
                          const float* gn_in_beta, int64_t gn_in_groups, float gn_in_eps, int64_t N, int64_t D,
                          int64_t H, int64_t W, int64_t Cin, int64_t Cout, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Training path (scripts/train.py -> TrainLoop.forward_backward, guided_diffusion/train_util.py:396-460): the
 * reference differentiates WavUNetModel with autograd (cuDNN dgrad/wgrad, native GroupNorm/SiLU backward, the
 * matmul DWT/IDWT backward of DWT_IDWT_Functions.py:139-156,184-208).  Here the backward is the explicit kernels
 * below on channels-last bf16 gradients; parameter gradients are fp32 in the reference's layouts and ACCUMULATE
 * into their destinations (zeroed once per step by the caller), which weight-tied ResBlocks (wunet.py:647-673) need.
 * ---------------------------------------------------------------------------------------------------- */

/* K5w: conv3d weight gradient, dW[co][ci][kd][kh][kw] (=, or += when accumulate) sum_v dY[v][co] * X[v + off][ci],
 * a tcgen05 GEMM with both operands MN-major (csrc/conv3d_wgrad.cu).  x: the conv's INPUT, cl bf16 with x_ld >=
 * round_up(Cin, 64) readable finite channels; dy: gradient of the conv's output, cl bf16 with dy_ld >= round_up(Cout, 64).
 * workspace: device scratch of fcwdm_conv3d_wgrad_workspace_bytes(...) bytes (per-split fp32 partial sums). */
int64_t fcwdm_conv3d_wgrad_workspace_bytes(int64_t N, int64_t D, int64_t H, int64_t W, int64_t Cin, int64_t Cout,
                                           int ksize);
int fcwdm_conv3d_wgrad(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw, void* workspace,
                       int64_t workspace_bytes, int accumulate, int64_t N, int64_t D, int64_t H, int64_t W, int64_t Cin,
                       int64_t Cout, int ksize, void* stream);
/* conv3d data gradient = fcwdm_conv3d_fwd / fcwdm_conv3d_pair_fwd of dY with these weights:
 * wt[ci][co][tap] = w[co][ci][taps-1-tap] (f32, standard layout; pack it with the forward packers). */
int fcwdm_conv3d_transpose_flip_weights(const float* w, float* wt, int64_t Cout, int64_t Cin, int ksize, void* stream);

/* GroupNorm(+SiLU) backward (nn.py:17-19): given x, the forward statistics (fcwdm_groupnorm_stats layout) and
 * dy = dL/d SiLU(GN(x)), writes dx (+ acc[voxel*acc_ld + c] if acc != NULL: gradient fan-in of x) and ADDS into
 * dgamma[C], dbeta[C] (either may be NULL).  sums: scratch double [N][FCWDM_GN_STAT_REPLICAS][C][2], zeroed here. */
int fcwdm_groupnorm_bwd(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, const double* stats,
                        const float* gamma, const float* beta, double* sums, const void* acc, int64_t acc_ld, void* dx,
                        int64_t dx_ld, float* dgamma, float* dbeta, int64_t N, int64_t S, int64_t C, int64_t G, float eps,
                        int silu, void* stream);
/* the same, and colsum[n*colsum_ld + c] += sum_v dx[n,v,c] over the values it stores (colsum may be NULL): the bias /
 * timestep-embedding gradient of the conv that produced x without another pass over dx.  fcwdm_colsum_scatter then adds
 * such a [N][C] partial into its destinations: out_sample[n*os_ld + c] += part[n][c], out_total[c] += sum_n part[n][c]. */
int fcwdm_groupnorm_bwd_colsum(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, const double* stats,
                               const float* gamma, const float* beta, double* sums, const void* acc, int64_t acc_ld,
                               void* dx, int64_t dx_ld, float* dgamma, float* dbeta, float* colsum, int64_t colsum_ld,
                               int64_t N, int64_t S, int64_t C, int64_t G, float eps, int silu, void* stream);
int fcwdm_colsum_scatter(const float* part, int64_t part_ld, float* out_sample, int64_t os_ld, float* out_total, int64_t N,
                         int64_t C, void* stream);
/* column sums of a cl bf16 tensor (conv bias and timestep-embedding gradients): out_sample[n*os_ld + c] += sum_v,
 * out_total[c] += sum_{n,v}; either may be NULL. */
int fcwdm_colsum_cl(const void* x, int64_t ld, float* out_sample, int64_t os_ld, float* out_total, int64_t N, int64_t S,
                    int64_t C, void* stream);
/* adjoint of fcwdm_dwt3d_cl: dx = IDWT(lll_scale*dlll, hi_scale*dhi (zero when dhi == NULL)) (+ acc). (D,H,W) = dx dims. */
int fcwdm_dwt3d_cl_bwd(const void* dlll, int64_t lll_ld, const void* dhi, int64_t hi_ld, int64_t hi_sb, const void* acc,
                       int64_t acc_ld, void* dx, int64_t dx_ld, int64_t N, int64_t D, int64_t H, int64_t W, int64_t C,
                       float lll_scale, float hi_scale, void* stream);
/* adjoint of fcwdm_idwt3d_cl: dlll = lll_scale*LLL(dy) (+ lll_acc), dhi (=, += when hi_accumulate) the 7 high bands of
 * dy; dlll or dhi may be NULL.  (D,H,W) = dy dims. */
int fcwdm_idwt3d_cl_bwd(const void* dy, int64_t dy_ld, const void* lll_acc, int64_t acc_ld, void* dlll, int64_t lll_ld,
                        void* dhi, int64_t hi_ld, int64_t hi_sb, int hi_accumulate, int64_t N, int64_t D, int64_t H,
                        int64_t W, int64_t C, float lll_scale, void* stream);
/* y = a + b on cl bf16 buffers (rows x C). */
int fcwdm_add_cl(const void* a, int64_t a_ld, const void* b, int64_t b_ld, void* y, int64_t y_ld, int64_t rows, int64_t C,
                 void* stream);
/* backward of fcwdm_linear with act_out = 0: dW[m][k] += sum_n dy[n][m]*act(x[n][k]), db[m] += sum_n dy[n][m],
 * dx[n][k] (=, += when accumulate_dx) act'(x[n][k]) * sum_m dy[n][m]*W[m][k]; dW/db/dx may be NULL. */
int fcwdm_linear_bwd(const float* x, const float* W, const float* dy, int64_t dy_ld, float* dx, float* dW, float* db,
                     int64_t N, int64_t K, int64_t M, int act_in, int accumulate_dx, void* stream);
/* Fused AdamW step over a flat fp32 range (torch.optim.AdamW semantics; train_util.py:75-82,391); the gradient is
 * multiplied by grad_scale first (1/world_size after a sum all-reduce). step counts from 1. */
int fcwdm_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                float weight_decay, int64_t step, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Plain UNetModel resampling (guided_diffusion/unet.py:40-100), channels-last bf16 (N, D, H, W, C), C % 8 == 0.
 * avgpool: avg_pool_nd(kernel = stride = 2) -- or (1,2,2) when pool_depth == 0 (resample_2d) -- of x (N,D,H,W,C);
 * upsample: F.interpolate(mode="nearest") by 2 in H and W, and in D when up_depth != 0; (D,H,W) = INPUT dims.
 * ---------------------------------------------------------------------------------------------------- */
int fcwdm_avgpool2_cl(const void* x, int64_t x_ld, void* y, int64_t y_ld, int64_t N, int64_t D, int64_t H, int64_t W,
                      int64_t C, int pool_depth, void* stream);
int fcwdm_upsample2_cl(const void* x, int64_t x_ld, void* y, int64_t y_ld, int64_t N, int64_t D, int64_t H, int64_t W,
                       int64_t C, int up_depth, void* stream);
/* their adjoints (training): dx[brick] = dy/(4 fd) (+ acc), (D,H,W) = dims of dy (pooled); dx = brick sums of dy (+ acc),
 * (D,H,W) = dims of dy (up-sampled). */
int fcwdm_avgpool2_cl_bwd(const void* dy, int64_t dy_ld, const void* acc, int64_t acc_ld, void* dx, int64_t dx_ld, int64_t N,
                          int64_t D, int64_t H, int64_t W, int64_t C, int pool_depth, void* stream);
int fcwdm_upsample2_cl_bwd(const void* dy, int64_t dy_ld, const void* acc, int64_t acc_ld, void* dx, int64_t dx_ld, int64_t N,
                           int64_t D, int64_t H, int64_t W, int64_t C, int up_depth, void* stream);

/* One-launch re-packing of all conv weights of a model.  jobs: device array of n_jobs records of 8 int64:
 * {src f32 master weight ptr, dst bf16 packed ptr, O, I, taps, pair, transposed, total}: O/I = output/input channels of
 * the conv being packed; pair = fcwdm_conv3d_pair_pack_weights layout, else fcwdm_conv3d_pack_weights layout;
 * transposed = the data-gradient form (fcwdm_conv3d_transpose_flip_weights folded in); total = packed element count.
 * max_total = the largest `total` (sizes the grid). */
int fcwdm_conv3d_pack_all(const void* jobs, int64_t n_jobs, int64_t max_total, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * BraTS volume preprocessing (guided_diffusion/bratsloader.py:44-50,107-111: clip_and_normalize + pad + crop), V raw
 * volumes (V, X, Y, Z) f32 -> (V, 1, X - 2*crop_x, Y - 2*crop_y, pad_z_to) f32 in [0, 1]:
 *   lo, hi = np.quantile(x, q_lo), np.quantile(x, q_hi) ('linear' interpolation, exact order statistics by radix select)
 *   out = (clip(x, lo, hi) - lo) / (hi - lo), zero-padded along Z, cropped in X and Y.
 * quantiles: device float [V][2] output (lo, hi).  workspace: fcwdm_clip_normalize_workspace_bytes(V) device bytes.
 * ---------------------------------------------------------------------------------------------------- */
int64_t fcwdm_clip_normalize_workspace_bytes(int64_t V);
int fcwdm_clip_normalize(const float* x, float* out, float* quantiles, void* workspace, int64_t workspace_bytes, int64_t V,
                         int64_t X, int64_t Y, int64_t Z, int64_t crop_x, int64_t crop_y, int64_t pad_z_to, double q_lo,
                         double q_hi, void* stream);

/* Development aid: per-CTA clock64 stamps of the next fcwdm_conv3d_fwd launches into a device buffer [grid][16]
 * (see csrc/conv3d.cu); NULL switches it off.  Not used by the product path. */
int fcwdm_debug_set_conv_trace(void* device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* FCWDM_H_ */
