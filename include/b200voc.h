/* b200voc -- C ABI of the B200-native (sm_100a) vocoder7 waveform-synthesis hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry point
 * cites the reference interface it replaces (paths relative to the reference repository
 * ChiefTriston/TTS-Core-Remastered-1).  The reference has no FFI of its own (it is pure Python
 * calling PyTorch library ops), so the "binding a maintainer would add" is the ctypes stub in
 * INTEGRATION.md / tts-core-remastered-1_b200/b200voc/_lib.py.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name says `host`;
 *   - the library never allocates or frees caller-visible device memory in the forward calls:
 *     outputs and workspaces are caller-allocated, sizes come from the *_workspace_bytes queries
 *     (handles own their packed weights);
 *   - every launch goes on the `stream` argument (a cudaStream_t passed as void*); calls are
 *     CUDA-graph capturable; handles are immutable after finalize, so concurrent forwards on
 *     different streams with different workspaces are safe;
 *   - return value 0 = ok, negative = error, text via b200voc_last_error_string() (thread local);
 *   - there is NO CPU fallback: without an sm_100 device the compute entry points fail.
 */
#ifndef B200VOC_H_
#define B200VOC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200VOC_OK 0
#define B200VOC_ERR_BAD_ARG (-1)
#define B200VOC_ERR_UNSUPPORTED (-2)
#define B200VOC_ERR_CUDA (-3)
#define B200VOC_ERR_STATE (-4)
#define B200VOC_ERR_OVERFLOW (-5)   /* overflow check enabled and a 16-bit activation left the storage format's range */

/* 16-bit tensor-core operand / activation storage formats (tcgen05 kind::f16 runs both at the
 * same rate). */
#define B200VOC_FMT_FP16 0
#define B200VOC_FMT_BF16 1

/* precision plans (DESIGN.md "Numerics plan"): per-stage operand format. */
#define B200VOC_PLAN_FP16 0  /* fp16 operands everywhere: meets max-abs<=1e-3, SNR>=40 dB (default) */
#define B200VOC_PLAN_BF16 1  /* bf16 operands everywhere: SNR gate only (max-abs ~5e-3)          */
#define B200VOC_PLAN_MIXED 2 /* bf16 stage 0, fp16 stages 1..3                                   */

int b200voc_version(void);
const char* b200voc_last_error_string(void);
/* 0 if device `dev` is sm_100 (B200), B200VOC_ERR_UNSUPPORTED otherwise. */
int b200voc_device_supported(int dev);

/* ----------------------------------------------------------------------------------------
 * Generator (replaces vocoder7/generator.py:9-98 `class Generator`).
 * -------------------------------------------------------------------------------------- */
typedef struct {
  /* vocoder7/config.py:11-20 */
  int32_t channels;      /* 80  */
  int32_t cond_dim;      /* 128 */
  int32_t style_dim;     /* 128 */
  int32_t num_bands;     /* 4   */
  int32_t n_stages;      /* len(upsample_factors) = 4 */
  int32_t upsample_factors[8];
  int32_t n_dilations;   /* len(res_dilations) = 3 */
  int32_t res_dilations[8];
  /* repairs R1/R3 (the reference leaves them undefined) */
  int32_t hidden_dim;    /* 512 */
  int32_t use_attention; /* generator.py:43-44 */
  int32_t attn_window;   /* 0 = global attention over all L positions, else block-local window */
  /* B200 knobs */
  int32_t precision_plan; /* B200VOC_PLAN_* */
  int32_t reserved[7];
} b200voc_gen_config;

typedef struct b200voc_gen b200voc_gen;

/* Generator.__init__ (generator.py:13-48): builds the layer plan, allocates packed-weight storage. */
int b200voc_gen_create(const b200voc_gen_config* cfg, b200voc_gen** out);
/* load_state_dict (train/blocks/vocoder.py:20-24 path): `name` is the reference state_dict key
 * (e.g. "upsample_blocks.0.0.weight"), `w` an fp32 device tensor in the reference layout
 * (Conv1d [out,in,k]; ConvTranspose1d [in,out,k]; Linear [out,in]).  Packs to the kernel layout. */
int b200voc_gen_set_weight(b200voc_gen* g, const char* name, const float* w, int64_t numel, void* stream);
/* number of state_dict entries the plan expects / i-th key (for host-side validation). */
int b200voc_gen_num_weights(const b200voc_gen* g);
const char* b200voc_gen_weight_name(const b200voc_gen* g, int i);
int64_t b200voc_gen_weight_numel(const b200voc_gen* g, int i);
/* fails with B200VOC_ERR_STATE if a key was never set. */
int b200voc_gen_finalize(b200voc_gen* g);
int64_t b200voc_gen_workspace_bytes(const b200voc_gen* g, int B, int T);
/* Generator.forward (generator.py:50-98).
 *   mel[B,channels,T] prosody[B,T,18] style[B,style_dim] emotion[B,6]  (fp32, contiguous)
 *   wav_out[B,1,hop*T] fp32.
 * Stream-ordered on `stream` for the caller (inputs may be released / outputs read by later work on that stream) and
 * capturable in a CUDA graph; internally the conditioning chain (style / emotion projections, prosody MLP, FiLM GEMM)
 * is forked onto a stream owned by the handle and joined before the first residual block, so two forwards on the SAME
 * handle must themselves be ordered (same stream, or an event between them); use one handle per concurrent stream.
 * tap_name/tap_out (both NULL normally): copy the named intermediate ("split","up0","res0.0",..,
 * "attn", oracle tap names) as fp32 [num_bands*B? no: B*num_bands, C, L] into tap_out. */
int b200voc_gen_forward(b200voc_gen* g, const float* mel, const float* prosody, const float* style,
                        const float* emotion, int B, int T, int style_drop, int emo_drop, float w_style,
                        float w_emo, float* wav_out, void* workspace, int64_t workspace_bytes,
                        const char* tap_name, float* tap_out, void* stream);
/* Generator.forward with the wire formats either side of the path:
 *   io->mel_time_major  1: `mel` is [B,T,channels], the layout the refiner / acoustic model emit
 *                          (sde_refiner5/model.py:304-306); vocoder7/trainer.py:77 transposes it on
 *                          the host, here the transpose is folded into the band-split load;
 *   io->out_format      B200VOC_OUT_F32: wav_out is float[B,1,hop*T] in (-1,1);
 *                       B200VOC_OUT_PCM16: wav_out is int16[B,1,hop*T] = round(clamp(wav)*32767);
 *   io->valid_samples   optional int32[B] (device): samples at or past valid_samples[b] are written
 *                       as 0 (the collator's wav_length = hop*frame_length of a padded batch,
 *                       batching2/colate.py:140-146,184-191).
 * io == NULL is b200voc_gen_forward. */
#define B200VOC_OUT_F32 0
#define B200VOC_OUT_PCM16 1
typedef struct {
  int32_t mel_time_major;
  int32_t out_format;
  const int32_t* valid_samples;
  int32_t reserved[4];
} b200voc_gen_io;
int b200voc_gen_forward_ex(b200voc_gen* g, const float* mel, const float* prosody, const float* style,
                           const float* emotion, int B, int T, int style_drop, int emo_drop, float w_style,
                           float w_emo, const b200voc_gen_io* io, void* wav_out, void* workspace,
                           int64_t workspace_bytes, const char* tap_name, float* tap_out, void* stream);

/* GlobalStyleTokens.forward (vocoder7/gst.py:24-35; the step before the Generator in its only caller,
 * vocoder7/trainer.py:73): style[B,style_dim] from mel[B,channels,T] (or [B,T,channels]).
 * conv0_w[style_dim,channels,3] conv0_b[style_dim] (attn_conv.0), conv2_w[num_tokens,style_dim]
 * conv2_b[num_tokens] (attn_conv.2), tokens[num_tokens,style_dim]; all fp32 device pointers. */
int64_t b200voc_gst_scratch_bytes(int B, int T, int num_tokens);
int b200voc_gst_forward(const float* mel, int mel_time_major, int B, int T, int channels, int style_dim,
                        int num_tokens, const float* conv0_w, const float* conv0_b, const float* conv2_w,
                        const float* conv2_b, const float* tokens, void* scratch, int64_t scratch_bytes,
                        float* style_out, void* stream);

/* ---- Discriminator forwards (forward half; vocoder7/discriminators.py:8-157) ----------
 * The three critics are stacks of strided 1-D convolutions + LeakyReLU(0.2) whose every intermediate map is
 * returned as a feature (discriminators.py:52-59, 101-107, 148-156).  The host module (b200voc/discriminators.py)
 * walks the layer list; these entry points are the layers.
 *
 * b200voc_disc_conv: y[b,co,lo,c] = bias[co] + sum_{ci,k} w[co,ci,k] * x[b,ci,lo*stride+k-pad,c] over fp32 maps
 * laid out [B, C, L, P] with the P columns innermost and untouched by the convolution:
 *   - MultiPeriodDiscriminator Conv2d(k=(5,1), stride=(3,1), padding=(2,0)) (discriminators.py:22-31): P = period;
 *   - MultiScale / MultiBand Conv1d (discriminators.py:77-89, 126-138): P = 1.
 * y_pre receives conv+bias, y_act LeakyReLU(slope) of it (either may be NULL).  in_batch_stride (elements, 0 =
 * contiguous) lets a batch of time-chunks (torch.chunk, discriminators.py:147) be read in place; in_valid
 * (elements per (b,ci) row, 0 = Lin*P) makes reads past the end of the waveform return zero, which is F.pad of
 * discriminators.py:46-48 without materialising the padded copy.  w is the spectral-normalised weight. */
int b200voc_disc_conv_out_len(int Lin, int K, int stride, int pad);
int b200voc_disc_conv(const float* x, const float* w, const float* bias, int B, int Cin, int Cout, int Lin, int P,
                      int K, int stride, int pad, int64_t in_batch_stride, int64_t in_valid, float slope,
                      float* y_pre, float* y_act, void* stream);
/* The same convolution on the tensor cores for the GEMM-shaped layers -- stride 1, P = 1, Cin a multiple of 64, Cout a
 * multiple of 128, odd K <= 41: MultiScaleDiscriminator's Conv1d(64 -> 256, k) and Conv1d(256 -> 1024, k)
 * (discriminators.py:71-92, 97 % of the critics' FLOPs).  tcgen05 implicit GEMM with split-bf16 operands
 * (x = hi + lo, three MMAs per k-step, fp32 accumulation: fp32-level accuracy at bf16 range).  w_split comes from
 * b200voc_disc_pack_weight_split (bf16 [2][Cout][K*Cin], b200voc_disc_split_weight_elems elements) applied to the
 * spectral-normalised weight; workspace (b200voc_disc_conv_tc_workspace_bytes, 16-byte aligned) receives the
 * channels-last split copy of x.  x is fp32 [B, Cin, L]; y_pre / y_act are fp32 [B, Cout, L + 2 pad - K + 1]. */
int b200voc_disc_conv_tc_supported(int Cin, int Cout, int K, int stride, int P);
int64_t b200voc_disc_conv_tc_workspace_bytes(int B, int Cin, int L);
int64_t b200voc_disc_split_weight_elems(int Cout, int Cin, int K);
int b200voc_disc_pack_weight_split(const float* w, int Cout, int Cin, int K, void* out, void* stream);
int b200voc_disc_conv_tc(const float* x, const void* w_split, const float* bias, int B, int Cin, int Cout, int L, int K,
                         int pad, float slope, float* y_pre, float* y_act, void* workspace, int64_t workspace_bytes,
                         void* stream);
/* torch.nn.utils.spectral_norm as the reference wraps every critic conv (discriminators.py:21-31, 76-89,
 * 125-138), evaluation mode (no power iteration): w_out = w_orig / sigma, sigma = u . (W v) with W = w_orig viewed
 * as [rows = Cout][cols = Cin*K(*1)].  sigma_out: one device float. */
int b200voc_spectral_norm_weight(const float* w_orig, const float* u, const float* v, int rows, int cols,
                                 float* w_out, float* sigma_out, void* stream);
/* The same in TRAINING mode -- what every critic forward of the reference trainer runs (vocoder7/trainer.py:86-115 on
 * modules in .train()): one power iteration first, v <- normalize(W^T u), u <- normalize(W v) with x / max(||x||, eps)
 * (torch's eps = 1e-12), then sigma = u . (W v) and w_out = w_orig / sigma.  u[rows] and v[cols] are UPDATED IN PLACE
 * (the module's weight_u / weight_v buffers).  scratch: rows + cols floats. */
int b200voc_spectral_norm_train(const float* w_orig, float* u, float* v, int rows, int cols, float eps, float* w_out,
                                float* sigma_out, float* scratch, void* stream);
/* F.avg_pool1d(x, 4, 2, 1) of discriminators.py:99 over `rows` rows of Lin samples -> (Lin-2)/2+1 samples. */
int b200voc_avg_pool1d_k4s2p1(const float* x, int64_t rows, int Lin, float* y, void* stream);

/* ---- Discriminator backward (the critic half of the training step) --------------------
 * Replaces what autograd runs for `d_loss.backward()` / `g_loss.backward()` through the critics
 * (vocoder7/trainer.py:86-115 over vocoder7/discriminators.py:8-157).  The host module chains these per layer, last
 * layer first (b200voc/discriminators.py, _CriticStackFn.backward).  Shapes and the in_batch_stride / in_valid
 * conventions are those of b200voc_disc_conv; every reduction runs in a fixed order (deterministic).
 *
 * b200voc_disc_lrelu_bwd: g = gy_pre + (gy_act + g_next) * (y_pre > 0 ? 1 : slope) -- the gradient of a conv map that
 *   is itself a returned feature (gy_pre), whose LeakyReLU is a returned feature (gy_act) and the next layer's input
 *   (g_next = that layer's dgrad).  Any of the three may be NULL; y_pre may be NULL only if gy_act and g_next are.
 * b200voc_disc_bias_grad: db[co] = sum over batch and the per_channel = Lout*P positions of g[B, Cout, per_channel].
 * b200voc_disc_conv_dgrad: dx[b,ci,li,c] = sum_{co,k} w[co,ci,k] g[b,co,lo,c] with lo*stride - pad + k = li, written
 *   in the layer input's layout (positions past in_valid are F.pad's zeros and receive nothing); accumulate != 0 adds
 *   to dx.  w is the spectral-normalised weight the forward used.
 * b200voc_disc_conv_wgrad: dw[co,ci,k] = sum_{b,lo,c} g[b,co,lo,c] x[b,ci,lo*stride - pad + k,c]; scratch
 *   (b200voc_disc_conv_wgrad_scratch_bytes, may be 0 -> NULL) holds per-slice partial sums.
 * b200voc_avg_pool1d_k4s2p1_bwd: gradient of F.avg_pool1d(x, 4, 2, 1) (discriminators.py:99): dx[rows, Lin].
 * b200voc_spectral_norm_bwd: dw_orig = (dw - <dw, w> u v^T) / sigma for w = w_orig / sigma, sigma = u . (W_orig v) with
 *   u, v constants (torch computes the power iteration under no_grad); u, v, sigma are the ones the forward used;
 *   scratch: b200voc_spectral_norm_bwd_scratch_bytes bytes, 8-byte aligned. */
int b200voc_disc_lrelu_bwd(const float* y_pre, const float* gy_pre, const float* gy_act, const float* g_next, float slope,
                           int64_t n, float* g_out, void* stream);
int b200voc_disc_bias_grad(const float* g, int B, int Cout, int64_t per_channel, float* db, void* stream);
int b200voc_disc_conv_dgrad(const float* g, const float* w, int B, int Cin, int Cout, int Lin, int P, int K, int stride,
                            int pad, int64_t in_batch_stride, int64_t in_valid, int accumulate, float* dx, void* stream);
int64_t b200voc_disc_conv_wgrad_scratch_bytes(int B, int Cin, int Cout, int Lin, int P, int K, int stride, int pad);
int b200voc_disc_conv_wgrad(const float* x, const float* g, int B, int Cin, int Cout, int Lin, int P, int K, int stride,
                            int pad, int64_t in_batch_stride, int64_t in_valid, float* dw, float* scratch, void* stream);
int b200voc_avg_pool1d_k4s2p1_bwd(const float* gy, int64_t rows, int Lin, float* dx, void* stream);
int64_t b200voc_spectral_norm_bwd_scratch_bytes(void);
int b200voc_spectral_norm_bwd(const float* dw, const float* w, const float* u, const float* v, const float* sigma, int rows,
                              int cols, float* dw_orig, void* scratch, void* stream);
/* Tensor-core forms for the GEMM-shaped layers (stride 1, P = 1: MSD's Conv1d(64 -> 256, k) and Conv1d(256 -> 1024, k),
 * discriminators.py:71-92).
 *   dgrad: a stride-1 convolution of g with the transposed, tap-flipped weight -- b200voc_disc_flip_weight
 *     (wt[ci][co][k] = w[co][ci][K-1-k]) -> b200voc_disc_pack_weight_split(wt, Cin, Cout, K) -> b200voc_disc_conv_tc(g,
 *     ..., B, Cout, Cin, Lout, K, pad' = K-1-pad, ...) with a zero bias; b200voc_disc_conv_dgrad_tc_supported says
 *     whether the shape qualifies (Cout % 64 == 0, Cin % 128 == 0).
 *   wgrad: one split-bf16 tcgen05 GEMM over positions, M = Cout, N = Cin*K, operands [g_hi|g_lo|g_hi] and the im2col
 *     [x_hi|x_hi|x_lo] packed into `workspace` (b200voc_disc_conv_wgrad_tc_workspace_bytes, 1024-byte aligned; the batch
 *     is processed in chunks that keep the packed operands under 1 GiB).  x [B, Cin, Lin], g [B, Cout, Lin+2pad-K+1]. */
int b200voc_disc_conv_dgrad_tc_supported(int Cin, int Cout, int K, int stride, int P, int pad);
int b200voc_disc_flip_weight(const float* w, int Cout, int Cin, int K, float* wt, void* stream);
int b200voc_disc_conv_wgrad_tc_supported(int B, int Cin, int Cout, int Lin, int K, int stride, int P, int pad);
int64_t b200voc_disc_conv_wgrad_tc_workspace_bytes(int B, int Cin, int Cout, int Lin, int K, int pad);
int b200voc_disc_conv_wgrad_tc(const float* x, const float* g, int B, int Cin, int Cout, int Lin, int K, int pad, float* dw,
                               void* workspace, int64_t workspace_bytes, void* stream);

/* Per-launch CUDA-event timing of the LAST forward (bench.py's roofline numbers).  Enable, run a
 * forward, synchronise the stream, then read entry i: layer name (oracle tap names), elapsed ms,
 * algorithmic FLOPs and algorithmic HBM bytes (DESIGN.md states the per-unit figures). */
int b200voc_gen_profile_enable(b200voc_gen* g, int enable);
int b200voc_gen_profile_count(const b200voc_gen* g);
const char* b200voc_gen_profile_name(const b200voc_gen* g, int i);
float b200voc_gen_profile_ms(const b200voc_gen* g, int i);
double b200voc_gen_profile_flops(const b200voc_gen* g, int i);
double b200voc_gen_profile_bytes(const b200voc_gen* g, int i);
/* Debug aid for the 16-bit storage plan: when enabled, every forward counts the Inf / NaN values of each layer's stored
 * activations (fp16 saturates at 65504) and fails with B200VOC_ERR_OVERFLOW naming the first layer that overflowed
 * (costs one pass over every activation buffer and a stream synchronisation; the last stage then runs without the fused
 * band_merge so that its output can be inspected).  Off by default. */
int b200voc_gen_set_overflow_check(b200voc_gen* g, int enable);
/* how many kernels one forward(B,T) launches (bench.py's gpu_launches). */
int b200voc_gen_launch_count(const b200voc_gen* g);
int b200voc_gen_destroy(b200voc_gen* g);

/* ----------------------------------------------------------------------------------------
 * Layer-level entry points (used by the Generator internally and by the layer-wise parity tests).
 * Activations are channels-last 16-bit: x16[N, L, C].
 * -------------------------------------------------------------------------------------- */
/* nn.ConvTranspose1d(Cin, Cout, 2*s, stride=s, padding=s/2)  (generator.py:35-38,87).
 * w_packed from b200voc_pack_convt_weight ([s*Cout rows][2*Cin] 16-bit).  out16[N, s*Lin, Cout];
 * store_lrelu != 0 stores leaky_relu(y, 0.1) instead of y. */
int64_t b200voc_convt_packed_elems(int Cin, int Cout, int s);
int b200voc_pack_convt_weight(const float* w_ref, int Cin, int Cout, int s, int fmt, void* w_packed, void* stream);
int b200voc_convt1d(const void* x16, const void* w_packed, const float* bias, int N, int Lin, int Cin, int Cout,
                    int s, int fmt, int store_lrelu, void* out16, void* stream);

/* ResidualBlock(C, dilation, cond_dim).forward(x, cond) (generator.py:40-41,89-90; body = repair
 * R2).  a16 is the activation in the form its producer stores it: leaky_relu(x) for the wide
 * stages (C >= 128), raw x for the narrow ones (C <= 64, where the kernel applies leaky_relu on
 * chip and adds the residual on the tensor core) -- b200voc_resblock_input_is_lrelu(C) tells.
 * film[B,T,2C] fp32 holds (1+scale | shift) at frame rate; N = num_bands*B sequences, sequence n
 * uses film row n/num_bands.  w1_packed [2C rows (GLU-interleaved)][3C]; w2_packed [C][C], or
 * [C][2C] = [W_proj | I] with w1 pre-scaled by 1/2 for C <= 64.  store_lrelu != 0 stores
 * leaky_relu(y). */
int64_t b200voc_resblock_packed_elems(int C);   /* elements of w1_packed + w2_packed */
int b200voc_resblock_input_is_lrelu(int C);
int b200voc_pack_resblock_weights(const float* w_conv, const float* w_proj, int C, int fmt, void* w_packed,
                                  void* stream);
int b200voc_resblock(const void* a16, const void* w_packed, const float* b_conv, const float* b_proj,
                     const float* film, int N, int L, int C, int dilation, int T, int num_bands, int fmt,
                     int store_lrelu, void* out16, void* stream);

/* One NARROW stage of Generator.forward (generator.py:85-98: `for layer in block` over ConvTranspose1d + the three
 * ResidualBlocks of stages 2 and 3, C = 64 / 32 output channels, stride 2) as one or two kernels with the intermediate
 * activations on chip, and -- C = 32, num_bands = 4, wav_out != NULL -- generator.py:96-98 (cat + band_merge + tanh)
 * folded in.  x16[N, Lin, 2C] raw; convt_w_packed / res_w_packed[3] from the pack functions above; film[B, T,
 * film_stride] fp32 with block j's (1+scale | shift) at column film_cols[j]; out16[N, 2*Lin, C] raw (or NULL when
 * wav_out[N/4, 2*Lin] fp32 is given); scratch16[N, 2*Lin, C] is needed for C = 64 only (the stage runs as two
 * launches).  merge_w_packed: b200voc_pack_merge_weight of band_merge.weight [1, 4*32, 7]. */
int64_t b200voc_merge_packed_elems(int num_bands);
int b200voc_pack_merge_weight(const float* w_ref, int num_bands, int fmt, void* w_packed, void* stream);
int b200voc_stage_fused(const void* x16, const void* convt_w_packed, const float* convt_bias,
                        const void* const* res_w_packed, const float* const* b_conv, const float* const* b_proj,
                        const int* dilations, const float* film, const int* film_cols, int film_stride, int N, int Lin,
                        int C, int T, int num_bands, int fmt, void* out16, void* scratch16, const void* merge_w_packed,
                        const float* merge_bias, float* wav_out, void* stream);

/* ----------------------------------------------------------------------------------------
 * STFT family (replaces vocoder7/stft.py:9-54 and the torchaudio MelSpectrogram call sites
 * reference_encoder/utils.py:31-36).  fp32 throughout.  frames = 1 + N / hop, bins = n_fft/2+1.
 * -------------------------------------------------------------------------------------- */
/* Builds and caches (per device) the window / twiddle tables of n_fft and, for n_mels > 0, the sparse HTK mel
 * filterbank, so that the calls below never allocate or copy synchronously (required before CUDA-graph capture; without
 * it the first call of a configuration builds its tables lazily). */
int b200voc_stft_prepare(int n_fft, int n_mels, int sample_rate);
/* LearnableSTFT.forward (stft.py:22-34): out[B,bins,frames] = |STFT(wav[B,N])| * gain[bins]
 * (gain may be NULL = ones). */
int b200voc_stft_mag(const float* wav, int B, int N, int n_fft, int hop, const float* gain, float* out, void* stream);
/* complex STFT, out_ri[B,bins,frames,2] (interleaved re,im = torch complex64 layout). */
int b200voc_stft_complex(const float* wav, int B, int N, int n_fft, int hop, float* out_ri, void* stream);
/* fused STFT -> |X|^2 -> 80-bin HTK mel -> log(clamp(.,1e-5)); out[B,n_mels,frames].
 * log_compress=0 returns the linear power mel. */
int b200voc_stft_logmel(const float* wav, int B, int N, int n_fft, int hop, int n_mels, int sample_rate,
                        int log_compress, float* out, void* stream);
/* torch.istft semantics (center, hann, length=N): spec_ri[B,bins,frames,2] -> wav[B,N]. */
int b200voc_istft(const float* spec_ri, int B, int frames, int n_fft, int hop, int N, float* wav, void* stream);

/* Backward of one resolution of STFTLoss.forward (vocoder7/stft.py:48-54), the first training-side
 * consumer of the STFT kernels:  L = scale * mean_{b,k,m} | |STFT(wav_fake)| g_k - |STFT(wav_real)| g_k |.
 * grad_wav[B,N] += dL/dwav_fake and grad_gain[n_fft/2+1] += dL/dg (grad_gain may be NULL); both are
 * ACCUMULATED so the three resolutions sum in place.  `gain` is the signed per-bin gain (LearnableSTFT.filterbank),
 * `scale` = lambda_stft * upstream gradient.  Workspace from the query below, 16-byte aligned. */
int64_t b200voc_stft_l1_backward_workspace_bytes(int B, int N, int n_fft, int hop);
int b200voc_stft_l1_backward(const float* wav_fake, const float* wav_real, int B, int N, int n_fft, int hop,
                             const float* gain, float scale, float* grad_wav, float* grad_gain, void* workspace,
                             int64_t workspace_bytes, void* stream);
/* Backward of LearnableSTFT.forward (vocoder7/stft.py:22-34: out = |STFT(wav)| * filterbank[:, None], differentiable in
 * the reference through torch.stft's autograd) for an ARBITRARY upstream gradient grad_out[B, n_fft/2+1, frames]:
 * grad_wav[B, N] += d/dwav (accumulates: zero it first), grad_gain[n_fft/2+1] += sum_{b,t} grad_out * |X| (may be NULL;
 * gain may be NULL = all ones).  Same construction as the L1 form above (complex STFT -> spectral gradient in place ->
 * adjoint overlap-add -> reflect fold); hop must divide n_fft/2.  workspace: b200voc_stft_mag_backward_workspace_bytes
 * bytes, 16-byte aligned. */
int64_t b200voc_stft_mag_backward_workspace_bytes(int B, int N, int n_fft, int hop);
int b200voc_stft_mag_backward(const float* wav, int B, int N, int n_fft, int hop, const float* gain, const float* grad_out,
                              float* grad_wav, float* grad_gain, void* workspace, int64_t workspace_bytes, void* stream);
/* STFTLoss.forward (stft.py:48-54) partial sums: out_sum[0] += sum |mag(fake)-mag(real)|*|gain|
 * for one resolution (host divides by numel and multiplies lambda). */
int b200voc_stft_l1(const float* wav_fake, const float* wav_real, int B, int N, int n_fft, int hop,
                    const float* gain, double* out_sum, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200VOC_H_ */
