/* dsengine.h — C ABI of libdsengine.so, the sm_100a kernel library behind diffsensei_b200.
 *
 * The reference (jianzongwu/DiffSensei) has no FFI: every kernel on its UNet sampling path is a
 * PyTorch library call.  Each entry point below therefore names the reference call site whose
 * arithmetic it replaces (paths relative to the reference repo root); the Python adapters in
 * diffsensei_b200/ bind them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless a comment says "host";
 *   - activations are bf16, channels-last: images are NHWC [B][H][W][C], token tensors [B][N][C]
 *     (the same memory — a transformer block needs no transpose);
 *   - norm/bias parameters are fp32; GEMM/conv weights are bf16, K-major ([out][in]);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - no entry point allocates device memory: scratch is passed in by the caller;
 *   - return value 0 = ok, otherwise a DS_ERR_* code; ds_last_error() gives the message
 *     (thread-local).  There is no CPU fallback: on a box without an sm_100 device every compute
 *     entry point returns DS_ERR_CUDA.
 */
#ifndef DSENGINE_H_
#define DSENGINE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DS_OK 0
#define DS_ERR_INVALID 1 /* bad argument / unsupported shape */
#define DS_ERR_CUDA 2    /* CUDA runtime / driver error      */

/* library version (major*10000 + minor*100 + patch) and last error message of the calling thread */
int ds_version(void);
const char* ds_last_error(void);
/* number of kernels this library has launched in the calling process (for bench.py's gpu_launches) */
uint64_t ds_launch_count(void);
/* cudaMemsetAsync(ptr, 0, bytes) on `stream`: a memset node (not a fill kernel) for the statistics pools the step clears */
int ds_zero_async(void* ptr, int64_t bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * GroupNorm (+ optional SiLU), NHWC bf16.                     [HBM-bound]
 * Replaces diffusers ResnetBlock2D.norm1/norm2 + nonlinearity (GroupNorm(32, eps=1e-5) -> SiLU),
 * Transformer2DModel.norm (eps=1e-6, no SiLU) and conv_norm_out+conv_act reached from
 * src/models/unet.py:251-261,281-290,316-338 — and the torch.cat([hidden, skip], 1) in front of every up-block
 * ResnetBlock2D (:316-332), which is never materialised.
 *
 * Statistics are per (sample, CHANNEL): fp64 [B][C][2] {sum, sum of squares} of the bf16 tensor.  They normally
 * come from the epilogue of the ds_gemm_bf16 / ds_conv3x3_nhwc call that produced the tensor (`chan_stats`), so the
 * GroupNorm itself is ONE pass: read x, write y (the algorithmic 4 B/element).  Per-channel sums compose: the
 * statistics of a channel concatenation are the two tensors' statistics side by side.
 *   ds_channel_stats   : statistics of a tensor that has no such producer; ACCUMULATES into `stats` (caller zeroes).
 *   ds_groupnorm_apply : y[B][HW][C1+C2] = act(GroupNorm([x1 | x2])) from stats1 [B][C1][2] / stats2 [B][C2][2];
 *                        x2 / stats2 NULL and C2 = 0 for a single source.  The channel sums are folded into the
 *                        `groups` group statistics in shared memory in a fixed order; normalisation, affine and SiLU
 *                        in fp32, one rounding to bf16.  y must not alias x1 / x2 when C2 > 0.
 *   ds_groupnorm_silu  : stand-alone form = zero scratch + ds_channel_stats + ds_groupnorm_apply (two passes over x).
 *                        `stats`: scratch of ds_groupnorm_scratch_floats(B, C) floats, 16-byte aligned.
 * All pointers 16-byte aligned; C, C1, C2 multiples of 8; (C1 + C2) % groups == 0; groups <= 64.
 * --------------------------------------------------------------------------------------------- */
int ds_channel_stats(const void* x, double* stats, int B, int HW, int C, void* stream);
int ds_groupnorm_apply(const void* x1, const double* stats1, int C1, const void* x2, const double* stats2, int C2,
                       void* y, const float* gamma, const float* beta, int B, int HW, int groups, float eps,
                       int apply_silu, void* stream);
int64_t ds_groupnorm_scratch_floats(int B, int C);
int ds_groupnorm_silu(const void* x, void* y, const float* gamma, const float* beta, float* stats, int B, int HW,
                      int C, int groups, float eps, int apply_silu, void* stream);

/* LayerNorm over the last dim, bf16 in/out, fp32 affine.     [HBM-bound]
 * Replaces BasicTransformerBlock.norm1/2/3 (eps 1e-5) and Resampler LayerNorms
 * (src/models/resampler.py:14,40-41,104). rows x C, C % 8 == 0, C <= 5120. */
int ds_layernorm(const void* x, void* y, const float* gamma, const float* beta, int rows, int C, float eps,
                 void* stream);

/* Dialog-bbox embedding add, in place on the NHWC conv_in output.   [HBM-bound]
 * Replaces UNetMangaModel.encode_dialog_bbox (src/models/unet.py:88-114):
 *   sample[b, y, x, :] += emb[:]  iff (x, y) lies in the union of the half-open pixel boxes
 *   [int(x1*W), int(x2*W)) x [int(y1*H), int(y2*H)), clamped to the image.
 * `dialog_bbox` is [B][num_dialogs][4] fp32 holding values ALREADY rounded to the unet dtype
 * (the reference multiplies in the unet dtype before int(): src/models/unet.py:102-105); set
 * `round_bf16` = 1 to reproduce the bf16 product rounding (int(bf16(0.9)*152) = 137), 0 for fp32. */
int ds_dialog_embed_add(void* sample, const float* emb, const float* dialog_bbox, int B, int H, int W, int C,
                        int num_dialogs, int round_bf16, void* stream);

/* Stand-alone IP attention mask (parity aid only — the fused cross-attention kernel computes the
 * same predicate in registers and never materialises it).
 * Replaces MaskedIPAttnProcessor2_0.prepare_attention_mask_ip (src/models/attention_processor.py:115-169).
 *   bbox : [B][num_ips][4] fp32;  mask out: [B][N][num_dummy + num_ips*tokens_per_ip] fp32 in {0,-10000}
 *   (identical across heads, so the head dim is not materialised). (H', W') are re-derived from
 *   (N, aspect_ratio) exactly as the reference does (:131-139). */
int ds_ip_mask(const float* bbox, float* mask, int B, int N, double aspect_ratio, int num_ips, int tokens_per_ip,
               int num_dummy, void* stream);

/* ---------------------------------------------------------------------------------------------
 * bf16 GEMM on tcgen05 tensor cores, TMA-fed, fp32 accumulation in TMEM.      [tensor-bound]
 *   out[M][Nout] = epilogue( A[M][K] * W[N][K]^T )
 * Replaces every nn.Linear on the path: attn.to_q/to_k/to_v/to_out (src/models/attention_processor.py:
 * 56,63-64,84,207,225-226,245-246,261), diffusers FeedForward/GEGLU, Transformer2D proj_in/proj_out,
 * ResnetBlock2D.time_emb_proj, TimestepEmbedding, Resampler linears (src/models/resampler.py:15-17,43-45,
 * 100-103), and 1x1 shortcut convs.
 * Epilogue, applied in fp32 before one rounding to bf16:
 *   v = acc + bias[n] + rowbias[row / rows_per_batch][n]
 *   DS_EPI_GEGLU: W/bias rows are packed in blocks of 128 "value" rows followed by their 128 "gate"
 *                 rows (see diffsensei_b200.weights.pack_geglu); out[:, j] = v_val * gelu_erf(v_gate),
 *                 Nout = N/2.  gelu_erf(x) = x * Phi(x) is evaluated as x * sigmoid(2k(x + a x^3 + b x^5)) with
 *                 (k, a, b) fitted to Phi: |abs err| <= 2.6e-5 for all x (the bf16 output rounding is >= 10x larger).
 *   DS_EPI_GELU / DS_EPI_SILU / DS_EPI_QUICKGELU : v = act(v)
 *   then v += residual[row][n] (bf16) and v *= out_scale (if != 0).
 * LayerNorm fusion (diffusers BasicTransformerBlock.norm1/2/3 -> the linears on either side of them):
 *   consumer: with ln_stats != NULL, A holds the UN-normalised rows and the caller passes pre-folded weights
 *             W' = W * gamma (per input feature), bias' = bias + W beta, ln_colsum[n] = sum_k W'[n][k]; then
 *             acc is replaced by rstd[row] * (acc - mean[row] * ln_colsum[n]) before anything else, where
 *             mean = sum/K, rstd = rsqrt(sumsq/K - mean^2 + ln_eps) from ln_stats[row] = {sum, sumsq}.
 *             Algebraically identical to LayerNorm(A) W^T + bias.
 *   producer: with row_stats_out != NULL the call zeroes it (unless row_stats_zeroed), then accumulates {sum, sum
 *             of squares} of every output row (fp32 values, before the bf16 rounding; per-tile fp32 partials are
 *             combined with fp64 atomics, which is exact, so the statistics do not depend on the arrival order) —
 *             bf16 outputs with 16-byte rows only.  zero_rows != NULL: the call also resets that [M][2] buffer (the first n-tile of
 *             every m-tile does it), which lets a chain of GEMMs rotate three statistics buffers without any
 *             memset node: the consumer of buffer k clears buffer k+2.
 * Split-K tail: the kernel is persistent (one CTA pair per two SMs, static round-robin over 256 x BN output tiles);
 *   when the tile count is not a multiple of the resident pairs, the tiles of the partial last round are cut along
 *   K into slices that run on different pairs, reduced in fp32 through `splitk_ws` (vector red.global.add) and
 *   finished by whichever slice arrives last.  The caller provides a 16-byte aligned workspace that is ALL ZERO on
 *   entry (ds_gemm_splitk_ws_bytes() bytes cover every shape; the kernel leaves it all zero again) and must not be
 *   shared by GEMMs running concurrently on different streams.  NULL simply disables the feature.
 * Mixed-width schedule (default on, DS_GEMM_TAIL=0 disables): when the last round of the persistent schedule would
 *   fill less than ~45 % of the CTA pairs, the m-rows that fall into it run as 128-column units (twice as many, half
 *   as long) appended to the 256-column units of the same launch, instead of paying a whole round for a handful of
 *   tiles.  Results are bit-identical to the plain schedule (same K order per output element).
 * Constraints: K % 8 == 0, lda % 8 == 0 (16-byte TMA strides). M, N, K tails are handled by TMA
 * zero-fill and masked stores.
 * --------------------------------------------------------------------------------------------- */
#define DS_EPI_NONE 0
#define DS_EPI_GEGLU 1
#define DS_EPI_GELU 2
#define DS_EPI_SILU 3
#define DS_EPI_QUICKGELU 4 /* v * sigmoid(1.702 v): CLIP-L text encoder MLP (hidden_act "quick_gelu") */

typedef struct {
  const void* a;        /* bf16 [M][lda]                                  */
  const void* w;        /* bf16 [N][ldw]  (row = output feature, K-major) */
  void* out;            /* bf16 [M][ldo]  (or fp32 when out_fp32 != 0)    */
  const float* bias;    /* [N] or NULL                                    */
  const float* rowbias; /* [ceil(M/rows_per_batch)][N] or NULL            */
  const void* residual; /* bf16 [M][ldres] or NULL                        */
  int32_t M, N, K;
  int32_t lda, ldw, ldo, ldres;
  int32_t rows_per_batch; /* only read when rowbias != NULL                */
  int32_t rowbias_ld;     /* row stride of rowbias in floats; 0 means N     */
  int32_t epilogue;       /* DS_EPI_*                                      */
  int32_t out_fp32;       /* 1: `out` is fp32                              */
  float out_scale;        /* 0 or 1: no scaling                            */
  const double* ln_stats; /* [M][2] fp64 (sum, sumsq) of A's rows, or NULL  */
  const float* ln_colsum; /* [N] fp32; required with ln_stats               */
  float ln_eps;
  double* row_stats_out;  /* [M][2] fp64 or NULL (see "producer" above)     */
  double* zero_rows;      /* [M][2] fp64 or NULL: rows reset to 0 by this call */
  int32_t row_stats_zeroed; /* 1: row_stats_out is already 0, skip the memset  */
  void* splitk_ws;        /* split-K workspace (see below) or NULL             */
  int64_t splitk_ws_bytes;
  /* second A operand (optional): the GEMM reads [a | a2] along K without the concatenation ever being written —
   * the 1x1 shortcut of an up-block ResnetBlock2D on torch.cat([hidden, skip], 1) (src/models/unet.py:316-332).
   * a2: bf16 [M][lda2], its K2 = K - K1 columns follow a's K1 columns; K1 % 64 == 0.  NULL: off (K1 ignored). */
  const void* a2;
  int32_t K1, lda2;
  /* producer-side GroupNorm statistics (optional): fp64 [M/stats_rows_per_sample][N][2] (sum, sum of squares) per
   * (sample, output channel) of the bf16-rounded outputs, ACCUMULATED with atomics (the caller zeroes it) — read by
   * ds_groupnorm_apply, so the GroupNorm that follows needs no statistics pass over the tensor.  Needs the bf16
   * TMA epilogue (16-byte rows), no GEGLU, stats_rows_per_sample % 128 == 0.  NULL: off. */
  double* chan_stats;
  int32_t stats_rows_per_sample;
  /* 1: `w` is constant data (model weights: never written by a kernel that can still be in flight) — the kernel may
   * then request its first weight tiles before it waits for the preceding kernel (programmatic dependent launch).
   * 0: `w` may have been produced by a preceding kernel (an activation used as the B operand): fetched after the wait. */
  int32_t w_is_constant;
} ds_gemm_args;

int ds_gemm_bf16(const ds_gemm_args* args, void* stream);
/* bytes of split-K workspace that suffice for any ds_gemm_bf16 / ds_conv3x3_nhwc call on the current device */
int64_t ds_gemm_splitk_ws_bytes(void);

/* A CHAIN of n (<= ds_gemm_chain_max()) dependent GEMMs as ONE persistent launch: args[q+1].a must be args[q].out
 * (the BasicTransformerBlock sequences attn.to_out -> attn2.to_q and attn2.to_out -> ff.net.0 -> ff.net.2 -> next
 * attn1.to_qkv of diffusers' BasicTransformerBlock.forward, each a torch.nn.Linear call in the reference; the
 * LayerNorms between them are folded as in ds_gemm_bf16).  Results are bit-identical to n ds_gemm_bf16 calls: same
 * tiles, same K order; only the schedule differs — the CTA pairs walk the problems back to back and a tile of problem
 * q+1 on row block m starts as soon as all column tiles of problem q for those 128 rows are written (a counter per
 * (problem, row block) in `dep`), so launch / prologue / drain are paid once and the partial last round of one
 * problem is filled with the first tiles of the next.  Every problem: same M > 128, bf16 output with 16-byte
 * addressable rows, no chan_stats, no split-K; residual / ln_stats / row_stats_out / zero_rows may refer to buffers
 * written by EARLIER problems of the chain for the same rows.
 * dep: dep_len >= ds_gemm_chain_max() * 2 * ceil(M / 256) + 1 ints, ALL ZERO on entry; the kernel leaves them all
 *      zero again (no memset between launches or graph replays); not to be shared by chains running concurrently. */
int ds_gemm_chain(const ds_gemm_args* args, int n, int* dep, int dep_len, void* stream);
int ds_gemm_chain_max(void);

/* ---------------------------------------------------------------------------------------------
 * 3x3 convolution, padding 1, stride 1 or 2, NHWC bf16, as an implicit GEMM on tcgen05:
 * the A operand is gathered by 4-D TMA tiles (one 8x16-pixel patch x 64 channels per filter tap;
 * the halo / zero padding is TMA out-of-bounds fill), never materialised.       [tensor-bound]
 * Replaces nn.Conv2d 3x3 in diffusers ResnetBlock2D.conv1/conv2, Downsample2D.conv (stride 2),
 * Upsample2D.conv and conv_out, reached from src/models/unet.py:251-261,281-290,316-338.
 *   x   : [B][H][W][Cin]          w: [Cout][3][3][Cin] bf16 (tap-major K; see weights.pack_conv3x3)
 *   out : [B][Ho][Wo][Cout],  Ho = (H-1)/stride+1 (same for Wo)
 *   rowbias : [B][Cout] fp32 or NULL — the ResnetBlock2D time-embedding projection, broadcast over pixels
 *   residual: bf16 [B][Ho][Wo][Cout] or NULL — the block's skip / shortcut branch
 * Constraints: Cin % 64 == 0 (conv_in with Cin=4 has its own entry point below).
 * --------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x;
  const void* w;
  void* out;
  const float* bias;
  const float* rowbias;
  const void* residual;
  int32_t B, H, W, Cin, Cout;
  int32_t stride;
  int32_t rowbias_ld; /* row stride of rowbias in floats; 0 means Cout */
  int32_t out_fp32;
  float out_scale;
  void* splitk_ws;    /* as in ds_gemm_args */
  int64_t splitk_ws_bytes;
  double* chan_stats; /* fp64 [B][Cout][2] producer-side GroupNorm statistics, as in ds_gemm_args; NULL: off */
  /* 1: out = conv3x3(nearest_x2(x)) — diffusers Upsample2D (F.interpolate(scale_factor=2, mode="nearest") + conv,
   * reached from src/models/unet.py:335-338) — WITHOUT materialising the upsampled tensor: four 2x2 convolutions of x,
   * one per output-pixel parity, whose taps are the pre-summed 3x3 taps that read the same input pixel (16 instead of
   * 36 MACs per input pixel).  x: [B][H][W][Cin], out: [B][2H][2W][Cout] bf16, w: [4][Cout][2][2][Cin] bf16 from
   * weights.pack_conv3x3_up2 (phase = 2*row_parity + col_parity).  stride 1, no residual / rowbias. */
  int32_t upsample2;
} ds_conv3x3_args;

int ds_conv3x3_nhwc(const ds_conv3x3_args* args, void* stream);

/* conv_in: 3x3, Cin = 4 (latent channels), direct CUDA-core kernel (K = 36 is not GEMM-shaped; the op is
 * bound by writing the [B][H][W][Cout] output).  x: NHWC bf16 [B][H][W][4]; w: fp32 [Cout][3][3][4].
 * Replaces UNet2DConditionModel.conv_in (src/models/unet.py:206).                       [HBM-bound] */
int ds_conv_in_3x3(const void* x, const float* w, const float* bias, void* out, int B, int H, int W, int Cout,
                   void* stream);
/* The same conv on the tensor cores: ds_im2col_latent writes A[B*H*W][64] bf16 (column tap*4 + c, taps row-major over
 * the 3x3 window, zero padding outside the image, columns 36..63 zero) and ds_gemm_bf16 with the weights packed as
 * [Cout][64] (weights.pack_conv_in) does the rest — bias, GroupNorm statistics of the output (chan_stats) and the
 * coalesced TMA store in its epilogue.  x: NHWC bf16 [B][H][W][4].                        [HBM-bound] */
int ds_im2col_latent(const void* x, void* a, int B, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused self-attention, head_dim 64, no mask: softmax(Q K^T / 8) V.              [tensor-bound]
 * Replaces F.scaled_dot_product_attention in AttnProcessor2_0.__call__
 * (src/models/attention_processor.py:69-81).
 *   qkv : bf16 [B][N][3*C] — the fused to_q|to_k|to_v projection output (C = heads*64); head h reads
 *         columns h*64.. of each third, via TMA boxes, so no head transpose is ever materialised
 *   out : bf16 [B][N][C]
 * --------------------------------------------------------------------------------------------- */
int ds_attention_self(const void* qkv, void* out, int B, int N, int heads, void* stream);

/* Fused text + masked-IP cross-attention, head_dim 64:                              [HBM-bound]
 *   out = softmax(Q Kt^T/8) Vt + scale * softmax(Q Kip^T/8 + M(bbox)) Vip
 * Replaces both SDPA calls, prepare_attention_mask_ip and the blend in
 * MaskedIPAttnProcessor2_0.__call__ (src/models/attention_processor.py:231-258).
 *   q    : bf16 [B][N][C]
 *   kv_t : bf16 [B][n_text][2*C]  (to_k | to_v of the text tokens; timestep-invariant)
 *   kv_ip: bf16 [B][n_ip][2*C]    (to_k_ip | to_v_ip of the image tokens, n_ip = num_dummy + num_ips*tokens_per_ip)
 *   bbox : fp32 [B][num_ips][4];  the additive mask is evaluated in registers with the reference's
 *          closed-interval linspace membership and derived (H', W') (:131-163); masked keys get -10000.
 * n_text, n_ip <= 128. */
typedef struct {
  const void* q;
  const void* kv_text;
  const void* kv_ip;
  const float* bbox;
  void* out;
  int32_t B, N, heads;
  int32_t n_text, n_ip;
  int32_t num_ips, tokens_per_ip, num_dummy;
  double aspect_ratio; /* latent H / W as a Python float (double): pipeline_diffsensei.py:272 */
  float ip_scale;
} ds_cross_ip_args;

int ds_attention_cross_ip(const ds_cross_ip_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Layout / glue kernels (all HBM-bound, vectorised)
 * --------------------------------------------------------------------------------------------- */
/* NCHW (fp32 or bf16) <-> NHWC bf16 at the diffusers-facing boundary of UNetMangaModel.forward */
int ds_nchw_to_nhwc(const void* src, int src_is_fp32, void* dst_bf16, int B, int C, int H, int W, void* stream);
int ds_nhwc_to_nchw(const void* src_bf16, void* dst, int dst_is_fp32, int B, int C, int H, int W, void* stream);
/* nearest-neighbour resize to (Ho, Wo) (Upsample2D's F.interpolate; src index = floor(dst * in/out)) */
int ds_upsample_nearest(const void* x, void* y, int B, int H, int W, int C, int Ho, int Wo, void* stream);
/* channel concat of two NHWC tensors: y[..., :C1] = a, y[..., C1:] = b  (torch.cat([hidden, skip], 1)) */
int ds_concat_channels(const void* a, const void* b, void* y, int pixels, int C1, int C2, void* stream);
/* elementwise y = silu(x), n bf16 elements */
int ds_silu(const void* x, void* y, int64_t n, void* stream);
/* Timesteps(num_channels, flip_sin_to_cos=True, downscale_freq_shift=0): out[r][:] = [cos | sin](t[r]*w) bf16 */
int ds_timestep_embedding(const float* t, void* out, int rows, int dim, void* stream);

/* ---------------------------------------------------------------------------------------------
 * CFG blend + DDIM (eta = 0, epsilon-prediction) update, fused.                    [HBM-bound]
 * Replaces src/pipelines/pipeline_diffsensei.py:315,332-337 (chunk, u + g*(t-u), scheduler.step,
 * and the torch.cat([latents]*2) for the next step):
 *   eps   = e_uncond + guidance * (e_text - e_uncond)
 *   x0    = (x - sqrt(1-a_t) * eps) / sqrt(a_t)
 *   x_new = sqrt(a_prev) * x0 + sqrt(1-a_prev) * eps
 *   noise_pred : bf16 NHWC [2*bs][H][W][4] (uncond half first)
 *   latents    : fp32 [bs][H][W][4] updated in place (fp32 master copy)
 *   model_in   : bf16 NHWC [2*bs][H][W][4] — x_new duplicated for both CFG halves (next step's UNet input)
 *   coef       : device pointer to 2 floats {alpha_prod_t, alpha_prod_t_prev} for this step
 * --------------------------------------------------------------------------------------------- */
int ds_cfg_ddim_step(const void* noise_pred, float* latents, void* model_in, const float* coef, float guidance,
                     int bs, int HW, int C, void* stream);

/* Perceiver attention of the character Resampler: 16 latent queries x (n_kv) keys per (character, head),
 * q and k each pre-scaled by dim_head^-0.25, fp32 softmax (src/models/resampler.py:64-74).
 *   q: bf16 [Bc][nq][C], kv: bf16 [Bc][n_kv][2*C] (k | v), out: bf16 [Bc][nq][C]; C = heads*64 */
int ds_resampler_attn(const void* q, const void* kv, void* out, int Bc, int nq, int n_kv, int heads, void* stream);

/* ---------------------------------------------------------------------------------------------
 * AutoencoderKL decoder helpers (the step right after the denoise loop,
 * src/pipelines/pipeline_diffsensei.py:339-367).  The decoder's convs / GroupNorms / linears / upsampling run on
 * the entry points above; these cover what is specific to it.                               [HBM-bound]
 *   ds_latent_pointwise : out[b][p][0..4) = W (latents[b][:][p] * inv_scale) + bias — `latents / scaling_factor`
 *                         followed by AutoencoderKL.post_quant_conv (1x1, 4 -> 4).  latents fp32 NCHW [B][4][HW],
 *                         w fp32 [4][4] (out, in), bias fp32 [4] or NULL, out bf16 NHWC [B][HW][4].
 *   ds_softmax_rows     : P[r][:] = softmax(scale * S[r][:]), S fp32 [rows][lds], P bf16 [rows][ldp], n <= 32768
 *                         columns — between the QK^T and PV GEMMs of the decoder's single 512-wide attention head.
 *   ds_image_postprocess: out = clamp(x / 2 + 0.5, 0, 1), x bf16 NHWC [B][HW][C] -> out fp32 NCHW [B][C][HW]
 *                         (VaeImageProcessor.postprocess with do_denormalize, output_type "pt").
 * --------------------------------------------------------------------------------------------- */
int ds_latent_pointwise(const float* latents, const float* w, const float* bias, void* out, float inv_scale, int B,
                        int HW, void* stream);
int ds_softmax_rows(const float* S, void* P, int rows, int n, int64_t lds, int64_t ldp, float scale, void* stream);
int ds_image_postprocess(const void* x, float* out, int B, int HW, int C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Conditioning-encoder helpers (CLIP ViT-H / Magi ViT-MAE image encoders,
 * src/pipelines/pipeline_diffsensei.py:125-128, and the two SDXL CLIP text encoders of encode_prompt, :232-245).
 * Their linears run on ds_gemm_bf16 and their LayerNorms on ds_layernorm.                    [latency-bound]
 *   ds_attention_small : out = softmax(scale * Q K^T [+ causal mask]) V for short sequences (Nk <= 320) and any
 *                        head_dim that is a multiple of 8 up to 256 (80 for ViT-H, 160 for the MLLM input resampler: outside the flash
 *                        kernel's 64).
 *                        q/k/v/out: bf16 [B][N][ld*] with head h at columns [h*head_dim, (h+1)*head_dim); the row
 *                        strides ld* are in elements, so q/k/v may point into one fused [B][N][3C] projection.
 *                        causal != 0: key j visible from query i iff j <= i (CLIP text), needs Nq == Nk.
 *   ds_embed_tokens    : out[b][t][:] = tok_emb[ids[b][t]][:] + pos_emb[t][:], bf16, ids int32 [B][L] (clamped to
 *                        the vocabulary) — CLIPTextEmbeddings.
 * --------------------------------------------------------------------------------------------- */
int ds_attention_small(const void* q, const void* k, const void* v, void* out, int B, int Nq, int Nk, int heads,
                       int head_dim, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float scale, int causal,
                       void* stream);
int ds_embed_tokens(const int* ids, const void* tok_emb, const void* pos_emb, void* out, int B, int L, int C, int vocab,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DSENGINE_H_ */
