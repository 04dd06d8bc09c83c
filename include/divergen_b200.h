/* divergen_b200 -- C ABI of the B200-native Stable-Diffusion denoising hot path.
 *
 * The reference (aim-uofa/DiverGen) has no FFI of its own: its boundary for this path is the diffusers Python class
 * surface used at DiverGen/generation/txt2img_diffusers_stages_from_txt.py:139-143, :242, :255-259.
 * This header is the C ABI that sits directly beneath that surface; each entry point names the diffusers call it
 * replaces.  Plain C: raw device pointers, sizes, a cudaStream_t passed as void*.  No torch types.
 *
 * Conventions
 *   - every function returns 0 on success or a negative DG_E_* code; dg_last_error() gives the message
 *     (thread-local).  No C++ exception crosses the ABI.
 *   - all tensors are fp16, contiguous, 16-byte aligned, owned by the caller; NCHW at this surface
 *     (kernel-internal NHWC is private).  The library owns packed weights, workspaces, TMA descriptors, CUDA graphs.
 *   - work is enqueued asynchronously on the supplied stream.  A handle is not thread-safe (one per process per GPU,
 *     the reference's process model: txt2img_diffusers_stages_from_txt.py:14-23,122).
 *   - creating a context on a device that is not sm_100 fails with DG_E_ARCH: there is no fallback path.
 */
#ifndef DIVERGEN_B200_H
#define DIVERGEN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DG_OK 0
#define DG_E_ARG (-1)
#define DG_E_SHAPE (-2)
#define DG_E_UNSUPPORTED (-3)
#define DG_E_CUDA (-4)
#define DG_E_NOMEM (-5)
#define DG_E_ARCH (-6)
#define DG_E_STATE (-7)

typedef struct dg_ctx dg_ctx;
typedef struct dg_unet dg_unet;

/* Mirrors the fields of diffusers' unet/config.json that SD-1.x / SD-2.x use. */
typedef struct dg_unet_config {
  int32_t in_channels;            /* 4 */
  int32_t out_channels;           /* 4 */
  int32_t sample_size;            /* 64 (SD-1.5) / 96 (SD-2.1-768) */
  int32_t block_out_channels[4];  /* 320, 640, 1280, 1280 */
  int32_t layers_per_block;       /* 2 */
  int32_t num_heads[4];           /* diffusers `attention_head_dim` (a head COUNT): 8,8,8,8 / 5,10,20,20 */
  int32_t cross_attention_dim;    /* 768 / 1024 */
  int32_t norm_num_groups;        /* 32 */
  float norm_eps;                 /* 1e-5 */
  int32_t use_linear_projection;  /* 0 / 1 : proj_in/proj_out are Conv1x1 or Linear (same arithmetic) */
  int32_t upcast_attention;       /* softmax is always fp32 here; accepted for config parity */
  int32_t down_has_attn[4];       /* 1,1,1,0 */
  int32_t flip_sin_to_cos;        /* 1 */
  float freq_shift;               /* 0 */
} dg_unet_config;

int32_t dg_version(void);
const char* dg_last_error(void);

/* ---- context --------------------------------------------------------------------------------------------------- */
int32_t dg_ctx_create(int32_t device_ordinal, dg_ctx** out);
void dg_ctx_destroy(dg_ctx* ctx);

/* ---- UNet2DConditionModel ------------------------------------------------------------------------------------------
 * dg_unet_create      <- UNet2DConditionModel.__init__ / from_pretrained (txt2img_...py:139)
 * dg_unet_set_weight  <- load_state_dict: one call per diffusers state-dict key (686 for SD-1.5/2.1).  `src` is a device
 *                        pointer to the fp16 tensor in PyTorch layout; it is repacked into kernel layout and may be freed
 *                        by the caller afterwards.
 * dg_unet_prepare     <- .to(device): allocates the activation workspace for up to max_batch samples of latent h x w.
 * dg_unet_forward     <- UNet2DConditionModel.forward(sample, timestep, encoder_hidden_states).sample
 *                        sample/out: [batch, C, h, w]; ehs: [batch, tokens, cross_attention_dim];
 *                        timesteps: host array of n_timesteps values (1 = broadcast, or one per sample).
 */
int32_t dg_unet_create(dg_ctx* ctx, const dg_unet_config* cfg, dg_unet** out);
void dg_unet_destroy(dg_unet* unet);
int32_t dg_unet_num_weights(dg_unet* unet);
const char* dg_unet_weight_name(dg_unet* unet, int32_t index);
int32_t dg_unet_weight_shape(dg_unet* unet, int32_t index, int64_t* shape4, int32_t* ndim);
int32_t dg_unet_set_weight(dg_unet* unet, const char* key, const void* src, int32_t ndim, const int64_t* shape);
int32_t dg_unet_missing_weights(dg_unet* unet); /* number of keys not yet set */
int32_t dg_unet_prepare(dg_unet* unet, int32_t max_batch, int32_t h, int32_t w, int32_t ctx_tokens);
int32_t dg_unet_forward(dg_unet* unet, const void* sample, const float* timesteps_host, int32_t n_timesteps,
                        const void* ehs, int32_t ctx_tokens, void* out, int32_t batch, int32_t h, int32_t w,
                        void* stream);
/* 1 (default): the forward is captured into a CUDA graph per (batch, pointers) and replayed; 0: eager launches. */
int32_t dg_unet_set_graphs(dg_unet* unet, int32_t enabled);
/* kernels launched by the last dg_unet_forward / dg_denoise_loop call (graph replays count their kernel nodes). */
int64_t dg_unet_last_launch_count(dg_unet* unet);
/* Measurement only: bit mask of the kernel families a forward enqueues (1 GEMM/conv, 2 attention, 4 normalisation, 8 other;
 * default 15).  With a family switched off the output is meaningless but the others' timing is not: bench.py uses it to time
 * one family as a replayed CUDA graph.  Never set it in production code. */
int32_t dg_unet_set_family_mask(dg_unet* unet, int32_t mask);

/* One eager (non-graph) forward with a CUDA-event pair around every launch: per kernel family
 * {0: tcgen05 GEMM/conv, 1: tcgen05 attention, 2: Group/LayerNorm, 3: other} accumulated device ms, algorithmic FLOPs,
 * algorithmic bytes and launch count (arrays of 4), plus the whole-forward device ms.  bench.py's roofline source. */
int32_t dg_unet_profile_forward(dg_unet* unet, const void* sample, const float* timesteps_host, int32_t n_timesteps,
                                const void* ehs, int32_t ctx_tokens, void* out, int32_t batch, int32_t h, int32_t w,
                                void* stream, double* ms_by_family, double* flops_by_family, double* bytes_by_family,
                                int64_t* launches_by_family, double* total_ms);

/* ---- StableDiffusionPipeline.__call__ steps 4-7 -----------------------------------------------------
 * dg_cfg_ddim_step <- `u + g*(c-u)` + DDIMScheduler.step(...).prev_sample (eta = 0), fused, in place on `latents`.
 *    noise_pred: [2*n_images (cfg) or n_images, elems_per_image]; prediction_type 0 = epsilon, 1 = v_prediction.
 * dg_denoise_loop  <- the whole loop: for t in timesteps: unet(cat([x,x]), t, cat([neg,pos])) -> CFG -> step.
 *    latents [n_images,4,h,w] are updated in place; ehs is [2*n_images, tokens, D] (uncond rows first) when
 *    guidance_scale > 1 else [n_images, tokens, D].  alphas: host arrays a_t[i], a_prev[i] per step.
 */
int32_t dg_cfg_ddim_step(dg_ctx* ctx, const void* noise_pred, void* latents, int32_t n_images, int64_t elems_per_image,
                         float alpha_t, float alpha_prev, float guidance_scale, int32_t prediction_type, void* stream);
int32_t dg_denoise_loop(dg_unet* unet, void* latents, const void* ehs, int32_t ctx_tokens, int32_t n_images, int32_t h,
                        int32_t w, const float* timesteps_host, const float* alpha_t_host, const float* alpha_prev_host,
                        int32_t n_steps, float guidance_scale, int32_t prediction_type, void* stream);

/* ---- AutoencoderKL.decode (StableDiffusionPipeline.__call__ step 8: `vae.decode(latents / scaling_factor).sample`) ----
 * The step between the denoising loop and `pt_to_pil(image)[j].save(...)` (txt2img_...py:267).  Weights by diffusers
 * state-dict key (`post_quant_conv.*`, `decoder.*`; 140 tensors for the SD VAE).  block_out_channels: 4 ints or NULL for
 * the SD config (128, 256, 512, 512); layers_per_block <= 0 -> 2.
 * dg_vae_decode: latents [B, 4, h, w] fp16 NCHW are multiplied by `scale` (pass 1 / scaling_factor, or 1 if the caller has
 * already divided); out [B, 3, 8h, 8w] fp16 NCHW in roughly [-1, 1]. */
typedef struct dg_vae dg_vae;
int32_t dg_vae_create(dg_ctx* ctx, const int32_t* block_out_channels, int32_t layers_per_block, dg_vae** out);
void dg_vae_destroy(dg_vae* vae);
int32_t dg_vae_num_weights(dg_vae* vae);
const char* dg_vae_weight_name(dg_vae* vae, int32_t index);
int32_t dg_vae_weight_shape(dg_vae* vae, int32_t index, int64_t* shape4, int32_t* ndim);
int32_t dg_vae_set_weight(dg_vae* vae, const char* key, const void* src, int32_t ndim, const int64_t* shape);
int32_t dg_vae_prepare(dg_vae* vae, int32_t max_batch, int32_t h, int32_t w);
int32_t dg_vae_decode(dg_vae* vae, const void* latents, float scale, void* out, int32_t batch, int32_t h, int32_t w,
                      void* stream);

/* ---- CLIPTextModel.forward(input_ids).last_hidden_state (the `prompt_embeds` of `stage_1.encode_prompt(prompt)`,
 * txt2img_diffusers_stages_from_txt.py:242; transformers `CLIPTextModel`, 123 060 480 parameters in 196 tensors for SD-1.x) ----
 * Weights by transformers state-dict key (`text_model.embeddings.*`, `text_model.encoder.layers.N.*`,
 * `text_model.final_layer_norm.*`).  Head dim must be 64 (hidden = 64 * heads).
 * dg_clip_encode: input_ids = HOST int32 [batch, seq] (tokenizer output, seq <= max_positions); out = device fp16
 * [batch, seq, hidden]. */
typedef struct dg_clip dg_clip;
int32_t dg_clip_create(dg_ctx* ctx, int32_t vocab, int32_t hidden, int32_t intermediate, int32_t layers, int32_t heads,
                       int32_t max_positions, dg_clip** out);
void dg_clip_destroy(dg_clip* clip);
int32_t dg_clip_num_weights(dg_clip* clip);
const char* dg_clip_weight_name(dg_clip* clip, int32_t index);
int32_t dg_clip_weight_shape(dg_clip* clip, int32_t index, int64_t* shape4, int32_t* ndim);
int32_t dg_clip_set_weight(dg_clip* clip, const char* key, const void* src, int32_t ndim, const int64_t* shape);
/* MLP activation: 0 = quick_gelu (OpenAI CLIP ViT-L/14, the SD-1.x text encoder; default), 1 = erf GELU (OpenCLIP ViT-H/14,
 * the SD-2.x text encoder; transformers `hidden_act: "gelu"`). */
int32_t dg_clip_set_activation(dg_clip* clip, int32_t act);
int32_t dg_clip_prepare(dg_clip* clip, int32_t max_batch);
int32_t dg_clip_encode(dg_clip* clip, const int32_t* input_ids, int32_t batch, int32_t seq, void* out, void* stream);

/* ---- CLIP similarity scores: `_, logits_per_text = clip_model(images, text)` of OpenAI CLIP ViT-L/14
 * (DiverGen/filteration/get_clip_score.py:176-180); weights by transformers `CLIPModel` state-dict key (text_model.*,
 * vision_model.*, visual_projection.weight, text_projection.weight; `logit_scale` through dg_clipscore_set_logit_scale) ----
 * text_cfg6 = {vocab, hidden, intermediate, layers, heads, max_positions}, vision_cfg6 = {image_size, patch_size, hidden,
 * intermediate, layers, heads}; NULL = ViT-L/14.  Head dim must be 64 in both towers.
 * dg_clipscore_score: pixel_values = device fp16 [n_images, 3, image, image] (already resized / cropped / normalised as
 * clip's `preprocess` does), input_ids = HOST int32 [n_texts, seq], eos_index = HOST int32 [n_texts] (position of the
 * end-of-text token: `input_ids.argmax(-1)`), logits_per_text = device fp32 [n_texts, n_images]. */
typedef struct dg_clipscore dg_clipscore;
int32_t dg_clipscore_create(dg_ctx* ctx, const int32_t* text_cfg6, const int32_t* vision_cfg6, int32_t projection_dim, dg_clipscore** out);
void dg_clipscore_destroy(dg_clipscore* cs);
int32_t dg_clipscore_num_weights(dg_clipscore* cs);
const char* dg_clipscore_weight_name(dg_clipscore* cs, int32_t index);
int32_t dg_clipscore_weight_shape(dg_clipscore* cs, int32_t index, int64_t* shape4, int32_t* ndim);
int32_t dg_clipscore_set_weight(dg_clipscore* cs, const char* key, const void* src, int32_t ndim, const int64_t* shape);
int32_t dg_clipscore_set_logit_scale(dg_clipscore* cs, float logit_scale);
int32_t dg_clipscore_prepare(dg_clipscore* cs, int32_t max_images, int32_t max_texts);
int32_t dg_clipscore_score(dg_clipscore* cs, const void* pixel_values, int32_t n_images, const int32_t* input_ids, const int32_t* eos_index,
                           int32_t n_texts, int32_t seq, float* logits_per_text, void* stream);

/* ---- output path: `pt_to_pil(image)` arithmetic on the device (txt2img_diffusers_stages_from_txt.py:267) ----
 * img [B, C<=4, H, W] fp16 in [-1, 1]  ->  out_u8 [B, H, W, C] uint8 = round(clamp(img / 2 + 0.5, 0, 1) * 255). */
int32_t dg_op_image_to_uint8(dg_ctx* ctx, const void* img, void* out_u8, int32_t B, int32_t C, int32_t H, int32_t W, void* stream);

/* ---- clip `preprocess` on the device (filteration/get_clip_score.py:128-150: Resize(224, BICUBIC) / CenterCrop / ToTensor /
 * Normalize on the PIL image): Pillow's 8-bit fixed-point antialiased resampling, one axis per pass.
 * dg_op_resample_u8: in [B, Hin, Win, C] uint8 -> out [B, Hout, Wout, C]; axis 0 resamples x (Hout == Hin), axis 1 y;
 * bounds = device int32 [n_out][2] (first input index, tap count), coeffs = device int32 [n_out][ksize] (22-bit fixed point,
 * computed by the host exactly as Pillow's precompute_coeffs / normalize_coeffs_8bpc do).
 * dg_op_clip_normalize: crop [top, top+n) x [left, left+n), (v / 255 - mean) / std, -> fp16 [B, 3, n, n]. */
/* Mask compositing of filteration/get_clip_score.py:133-146 (--use_mask): out = mask > 128 ? img : 1 on [B, H, W, 3] uint8 images
 * with [B, H, W] uint8 masks; count[b] (device uint32, zeroed here) = number of mask pixels > 128 (area = count / (H*W)). */
int32_t dg_op_mask_composite_u8(dg_ctx* ctx, const void* img_u8, const void* mask_u8, void* out_u8, uint32_t* count, int32_t B,
                                int32_t H, int32_t W, void* stream);
int32_t dg_op_resample_u8(dg_ctx* ctx, const void* in_u8, void* out_u8, int32_t B, int32_t Hin, int32_t Win, int32_t C, int32_t Hout,
                          int32_t Wout, const int32_t* bounds, const int32_t* coeffs, int32_t ksize, int32_t axis, void* stream);
int32_t dg_op_clip_normalize(dg_ctx* ctx, const void* in_u8, void* out, int32_t B, int32_t H, int32_t W, int32_t top, int32_t left, int32_t n,
                             const float* mean3, const float* std3, void* stream);

/* ---- single operators (exported for the parity tests; the same launchers the UNet uses) ---------------------------
 * dg_op_gemm: out[M, n_out] = epi(A[M, K] * W[n_w, K]^T)      <- torch.nn.Linear / Conv2d 1x1
 *    bias [n_w] / residual [M, n_out] optional; geglu: W is GEGLU-packed (see dg_op_pack_geglu), n_out = inner dim.
 * dg_op_conv3x3: NHWC x[B,H,W,C0] (++ x1[B,H,W,C1]) * Wp[N, 9*(C0+C1)] <- Conv2d 3x3 stride 1 pad 1
 *    rowvec [B, ld_rowvec] optional per-sample additive vector (time embedding).  ldo: row pitch of out / residual in
 *    elements (0 = N); must be a multiple of 8 (TMA store needs 16-byte pitches), so N = 4 (conv_out) uses ldo = 8.
 * dg_op_attention: out[B,Sq,heads*d] = softmax(Q K^T / sqrt(d)) V  <- diffusers Attention core
 *    q/k/v are base pointers of [B, S, ld] fp16 matrices (head h at columns [h*d, (h+1)*d)).
 */
int32_t dg_op_gemm(dg_ctx* ctx, const void* A, const void* W, const void* bias, const void* residual, void* out,
                   int32_t M, int32_t K, int32_t n_w, int32_t n_out, int32_t geglu, void* stream);
int32_t dg_op_pack_geglu(dg_ctx* ctx, const void* w, const void* b, void* w_out, void* b_out, int32_t inner, int32_t K,
                         void* stream);
int32_t dg_op_geglu_packed_rows(int32_t inner);
/* Fused-epilogue variants (the forms the UNet actually launches):
 * dg_op_gemm_fused: dg_op_gemm plus
 *    - LayerNorm fold   <- torch.nn.LayerNorm + Linear: W/colsum/bias32 from dg_op_fold_layernorm, ln_stats = per-row
 *      (sum, sumsq) partials [M][dg_op_gemm_row_parts(ln_c)][2] of A (from a producing GEMM's row_stats_out or
 *      dg_op_row_stats); out = rstd*(A W^T - mean*colsum) + bias32.
 *    - row_stats_out    per-row (sum, sumsq) partials of the fp16 result, [M][dg_op_gemm_row_parts(n_out)][2]
 *    - gn_stats_out     <- torch.nn.GroupNorm statistics of the result: fp32 sums over gn_blk-channel blocks,
 *      [M/hw][n_out/gn_blk][2], ACCUMULATED (caller zeroes); hw = rows per sample.
 * dg_op_groupnorm_fused: GroupNorm apply (+SiLU) from such block sums (one array per concatenated source). */
int32_t dg_op_gemm_row_parts(int32_t n_out);
int32_t dg_op_gemm_fused(dg_ctx* ctx, const void* A, const void* W, const void* bias, const float* bias32,
                         const float* colsum, const float* ln_stats, int32_t ln_c, float ln_eps, const void* residual,
                         void* out, int32_t M, int32_t K, int32_t n_w, int32_t n_out, int32_t geglu, float* row_stats_out,
                         float* gn_stats_out, int32_t gn_blk, int32_t hw, void* stream);
int32_t dg_op_fold_layernorm(dg_ctx* ctx, const void* W, const void* bias, const void* gamma, const void* beta, void* Wf,
                             float* colsum, float* b32, int32_t N, int32_t K, void* stream);
int32_t dg_op_row_stats(dg_ctx* ctx, const void* x, float* stats, int32_t rows, int32_t C, int32_t parts, void* stream);
int32_t dg_op_conv3x3_stats(dg_ctx* ctx, const void* x0, int32_t C0, const void* Wp, const void* bias, void* out, int32_t B,
                            int32_t H, int32_t Wd, int32_t N, float* gn_stats_out, int32_t gn_blk, void* stream);
int32_t dg_op_groupnorm_fused(dg_ctx* ctx, const void* x0, int32_t C0, const float* stats0, const void* x1, int32_t C1,
                              const float* stats1, int32_t blk, const void* gamma, const void* beta, void* out, int32_t B,
                              int32_t HW, int32_t groups, float eps, int32_t silu, void* stream);
int32_t dg_op_pack_conv3x3(dg_ctx* ctx, const void* w_oihw, void* w_out, int32_t O, int32_t I, void* stream);
int32_t dg_op_conv3x3(dg_ctx* ctx, const void* x0, int32_t C0, const void* x1, int32_t C1, const void* Wp,
                      const void* bias, const void* rowvec, int32_t ld_rowvec, const void* residual, void* out,
                      int32_t B, int32_t H, int32_t Wd, int32_t N, int32_t ldo, void* stream);
/* conv3x3 (taps = 9) or 1x1 / Linear (taps = 1) over GroupNorm(x) [+ SiLU], the normalisation applied INSIDE the GEMM's operand
 * path (diffusers ResnetBlock2D: conv(act(norm(x)))): x0 / x1 are the RAW NHWC sources (x1 optional: channel concat), stats0 /
 * stats1 their fused block sums [B][H*W/32][C/blk] float2 as written by dg_op_conv3x3_stats / the GEMM epilogues, Wp the packed
 * weight, out [B, H, W, ldo] (row pitch ldo >= N, a multiple of 8; 0 = N).  Equals dg_op_groupnorm_fused followed by
 * dg_op_conv3x3 up to fp16 rounding. */
int32_t dg_op_conv3x3_gn(dg_ctx* ctx, const void* x0, int32_t C0, const float* stats0, const void* x1, int32_t C1, const float* stats1,
                         int32_t blk, const void* gamma, const void* beta, int32_t groups, float eps, int32_t silu, const void* Wp,
                         const void* bias, const void* residual, void* out, int32_t B, int32_t H, int32_t Wd, int32_t N, int32_t ldo,
                         int32_t taps, void* stream);
/* ResnetBlock2D tail: out = conv3x3(x) + conv1x1(cat[xs0, xs1]) + bias in ONE K loop (the conv_shortcut is two more operand
 * sources after the nine taps).  Wcat [N, 9*C + Cs0 + Cs1] = [dg_op_pack_conv3x3 layout | shortcut weight [N, Cs0 + Cs1]] per row;
 * bias = conv bias + shortcut bias; x [B, H, W, C], xs0 / xs1 [B, H, W, Cs0 / Cs1] (xs1 optional); out [B, H, W, N]. */
int32_t dg_op_conv3x3_shortcut(dg_ctx* ctx, const void* x, int32_t C, const void* Wcat, const void* bias, const void* xs0, int32_t Cs0,
                               const void* xs1, int32_t Cs1, void* out, int32_t B, int32_t H, int32_t Wd, int32_t N, void* stream);
/* diffusers Downsample2D: conv3x3, stride 2, pad 1 on NHWC x [B, 2*Hout, 2*Wout, C] with the packed weight of dg_op_pack_conv3x3;
 * out [B, Hout, Wout, N].  The operand boxes take every second input pixel through a strided tensor map (no im2col tensor). */
int32_t dg_op_conv3x3_stride2(dg_ctx* ctx, const void* x, int32_t C, const void* Wp, const void* bias, void* out, int32_t B, int32_t Hout,
                              int32_t Wout, int32_t N, float* gn_stats_out, int32_t gn_blk, void* stream);
/* diffusers Upsample2D: nearest x2 followed by conv3x3 (pad 1), computed as four 2x2 "phase" convolutions on the LOW-resolution
 * NHWC input x [B, H, W, C] (weights summed per phase: 2.25x fewer multiply-adds, no upsampled tensor).  w_oihw [N, C, 3, 3] is
 * the ordinary conv weight; out [B, 2H, 2W, N].  gn_stats_out (optional): block sums [B][4*H*W/32][N/gn_blk] float2 of out. */
int32_t dg_op_upsample_conv3x3(dg_ctx* ctx, const void* x, int32_t C, const void* w_oihw, const void* bias, void* out, int32_t B, int32_t H,
                               int32_t Wd, int32_t N, float* gn_stats_out, int32_t gn_blk, void* stream);
int32_t dg_op_attention(dg_ctx* ctx, const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                        void* out, int32_t B, int32_t heads, int32_t Sq, int32_t Sk, int32_t d, void* stream);
/* Host-only (no GPU needed): the wave-balanced grid dg_op_attention / the UNet use for n_bh = B * heads (batch, head) pairs of
 * q_tiles 128-row query tiles each, with kernels that take tiles_per_cta query tiles per CTA, on `sms` one-CTA SMs.
 * plan[0] = 1 if a balanced grid beats the uniform one (else 0 and the rest is 0); plan[1] = pairs in group 1;
 * plan[2], plan[3] = CTAs of tiles_per_cta / tiles_per_cta - 1 tiles per pair of group 1; plan[4], plan[5] = the same for the
 * remaining pairs; plan[6] = CTAs in the grid (full-size CTAs first); makespan[0] / makespan[1] = simulated / uniform run
 * time in query-tile units.  (Exported for tests/test_attn_plan.py.) */
int32_t dg_plan_attention_grid(int32_t n_bh, int32_t q_tiles, int32_t tiles_per_cta, int32_t sms, int32_t* plan, double* makespan);
int32_t dg_op_groupnorm(dg_ctx* ctx, const void* x0, int32_t C0, const void* x1, int32_t C1, const void* gamma,
                        const void* beta, void* out, int32_t B, int32_t HW, int32_t groups, float eps, int32_t silu,
                        void* stream);
int32_t dg_op_layernorm(dg_ctx* ctx, const void* x, const void* gamma, const void* beta, void* out, int32_t rows,
                        int32_t C, float eps, void* stream);
/* Timesteps(dim) sinusoid -> Linear -> SiLU -> Linear  <- diffusers Timesteps + TimestepEmbedding */
int32_t dg_op_time_embedding(dg_ctx* ctx, const float* timesteps_host, int32_t B, int32_t dim, int32_t temb_dim,
                             const void* w1, const void* b1, const void* w2, const void* b2, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIVERGEN_B200_H */
