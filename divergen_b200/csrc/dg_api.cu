// C-ABI translation unit (include/divergen_b200.h): context, UNet2DConditionModel graph, denoise loop, operator
// entry points.  Host C++ only enqueues the hand-written sm_100a kernels; there is no CPU or library compute path.
#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "host_common.cuh"

using namespace dg;

struct dg_ctx {
  int device = 0;
  int num_sms = 0;
  float* gn_stats = nullptr;  // scratch for the stand-alone groupnorm op (tests)
  GemmRes gemm;               // split-K workspace, tickets, persistent grid size
  int fuse_ln = 1;            // DG_FUSE_LN=0: stand-alone LayerNorm kernels instead of the folded GEMM epilogue
  int fuse_gn = 1;            // DG_FUSE_GN=0: stand-alone GroupNorm statistics kernels instead of epilogue sums
  int fuse_xf = 0;            // DG_FUSE_XF=0: stand-alone GroupNorm-apply (+SiLU) pass instead of the transform inside the consuming conv;
                              // 2: also the transformer's GroupNorm inside proj_in
  int cfg_dedup = 1;          // DG_CFG_DEDUP=0: the guidance loop runs the identical prefix of the two CFG halves twice (round 1)
  int fuse_sc = 1;            // DG_FUSE_SC=0: conv_shortcut as its own 1x1 GEMM whose output conv2 adds as a residual (round 1)
  int up_phases = 1;          // DG_UPCONV_PHASES=0: materialise the nearest-x2 tensor and run the 9-tap conv on it (round 1)
  __half* xf_tab = nullptr;   // scratch table for the stand-alone fused-GroupNorm operators (tests)
  size_t xf_tab_cap = 0;
};

// ================================================================== arena allocator (deterministic, graph friendly)
struct Arena {
  uint8_t* base = nullptr;
  size_t size = 0;
  struct Blk { size_t off, len; bool used; };
  std::vector<Blk> blks;
  size_t high_water = 0;
  void reset() { blks.clear(); blks.push_back({0, size, false}); }
  void* alloc(size_t bytes) {
    bytes = (bytes + 1023) & ~size_t(1023);
    for (size_t i = 0; i < blks.size(); ++i) {
      if (!blks[i].used && blks[i].len >= bytes) {
        if (blks[i].len > bytes) {
          Blk rest{blks[i].off + bytes, blks[i].len - bytes, false};
          blks[i].len = bytes;
          blks.insert(blks.begin() + i + 1, rest);
        }
        blks[i].used = true;
        high_water = std::max(high_water, blks[i].off + bytes);
        return base + blks[i].off;
      }
    }
    return nullptr;
  }
  void release(void* p) {
    if (!p) return;
    size_t off = (uint8_t*)p - base;
    for (size_t i = 0; i < blks.size(); ++i) {
      if (blks[i].off == off && blks[i].used) {
        blks[i].used = false;
        if (i + 1 < blks.size() && !blks[i + 1].used) { blks[i].len += blks[i + 1].len; blks.erase(blks.begin() + i + 1); }
        if (i > 0 && !blks[i - 1].used) { blks[i - 1].len += blks[i].len; blks.erase(blks.begin() + i); }
        return;
      }
    }
  }
};

// ================================================================== weights
enum PackKind { PK_COPY, PK_CONV3, PK_CONV_IN, PK_GEGLU_W, PK_GEGLU_B, PK_ROWS, PK_PAD_ROWS };

struct Slot {
  std::string key;
  std::vector<int64_t> shape;  // expected PyTorch shape
  PackKind kind;
  __half* dst = nullptr;       // destination base
  __half* dst2 = nullptr;      // PK_CONV3 of an upsampler: also the four phase matrices [4][O][4*I] (pack_upconv_kernel)
  int64_t row_off = 0;         // PK_ROWS: destination row offset (fused QKV / KV / time_emb_proj tables)
  int a = 0, b = 0;            // kind-specific dims
  bool set = false;
};

struct Norm { __half* g = nullptr; __half* b = nullptr; int c = 0; };
struct Lin {
  __half* w = nullptr; __half* b = nullptr; int in = 0, out = 0; int rows = 0;
  __half* wph = nullptr;       // upsampler convs: phase-decomposed weights [4][out][4*in] (see pack_upconv_kernel)
  // LayerNorm-folded copy (finalize_weights): wf = w * gamma, cs = row sums of wf, b32 = b + w . beta
  __half* wf = nullptr; float* cs = nullptr; float* b32 = nullptr;
};
struct Res {
  Norm n1, n2; Lin c1, c2, sc; bool has_sc = false; int cin = 0, cout = 0; int temb_off = 0;
  Lin c2x;       // has_sc: [conv2 | conv_shortcut] weights per row + summed bias (finalize_weights): the shortcut runs inside conv2's K loop
};
struct Xf {
  Norm gn, ln1, ln2, ln3;
  Lin proj_in, qkv, o1, q2, kv2, o2, ff1, ff2, proj_out;
  int c = 0, heads = 0, ff_inner = 0;
};
struct DownBlk { std::vector<Res> res; std::vector<Xf> xf; bool has_down = false; Lin down; };
struct UpBlk { std::vector<Res> res; std::vector<Xf> xf; bool has_up = false; Lin up; };

struct T4 {
  __half* p = nullptr; int B = 0, H = 0, W = 0, C = 0;
  float* gst = nullptr;   // fused GroupNorm block sums [B][C/blk][2] written by the producing GEMM (nullptr: none)
  size_t bytes() const { return (size_t)B * H * W * C * 2; }
};

struct GraphKey {
  int batch, h, w, tokens; const void* sample; const void* ehs; void* out; int hoisted; int fam_mask;
  bool operator==(const GraphKey& o) const {
    return batch == o.batch && h == o.h && w == o.w && tokens == o.tokens && sample == o.sample && ehs == o.ehs && out == o.out &&
           hoisted == o.hoisted && fam_mask == o.fam_mask;
  }
};

// Packed-weight registry shared by the model handles (UNet, VAE decoder): state-dict key -> destination + packing rule.
struct WeightStore {
  dg_ctx* ctx = nullptr;
  std::vector<Slot> slots;
  std::map<std::string, int> slot_index;
  std::vector<void*> owned;  // cudaMalloc'd weight buffers
};

struct dg_unet : WeightStore {
  dg_unet_config cfg{};
  // modules
  Lin conv_in, conv_out, time1, time2;
  Norm norm_out;
  __half* temb_proj_w = nullptr; __half* temb_proj_b = nullptr; int temb_total = 0;
  std::vector<DownBlk> down;
  Res mid_r0, mid_r1; Xf mid_xf;
  std::vector<UpBlk> up;
  int temb_dim = 0;
  // workspace
  Arena arena;
  int max_batch = 0, ws_h = 0, ws_w = 0, ws_tokens = 0;
  float* d_t = nullptr;       // [max_batch] timesteps (device)
  float* gn_stats = nullptr;  // [max_batch, groups, 2]
  float4* d_coef = nullptr;   // DDIM coefficient table (per step)
  int* d_step = nullptr;      // current step index (device)
  __half* loop_in = nullptr;  // [2B,4,h,w] CFG-duplicated UNet input
  __half* loop_out = nullptr; // [2B,4,h,w] noise prediction
  int coef_cap = 0;
  // fused-statistics scratch: bump-allocated per forward in launch order (graph-stable addresses)
  float* gn_arena = nullptr; size_t gn_cap = 0, gn_off = 0;   // GroupNorm slab/block sums (fully overwritten, never zeroed)
  float* ln_arena = nullptr; size_t ln_cap = 0, ln_off = 0;   // LayerNorm row partials (fully overwritten, never zeroed)
  __half* xf_arena = nullptr; size_t xf_cap = 0, xf_off = 0;  // (mean_h, scale', shift') tables of the GroupNorms applied inside convs
  int gn_blk = 0;             // channel-block width of the fused GroupNorm sums (block_out_channels[0] / groups), 0 = off
  // step-invariant work hoisted out of the denoising loop (dg_denoise_loop): cross-attention K/V of the text embedding
  // (one buffer per transformer block, in forward order) and the per-step time-embedding projections
  std::vector<__half*> kv_cache; size_t kv_cache_elems = 0;
  bool kv_ready = false;      // the forward reads kv_cache instead of projecting encoder_hidden_states
  __half* temb_table = nullptr; int temb_table_rows = 0;   // [n_steps, temb_total]
  __half* temb_cur = nullptr;                               // [temb_total]: row of the current step (set_step_kernel)
  bool temb_ready = false;    // the forward reads temb_cur (same row for every sample) instead of running the MLP
  // classifier-free guidance inside dg_denoise_loop: rows [0, B/2) and [B/2, B) of the UNet input are the SAME latents at the same
  // timestep, and nothing before the first cross-attention sees the text embeddings -- conv_in, the first ResnetBlock2D and the
  // first transformer block up to its self-attention out-projection run on B/2 rows and are then duplicated (exact: identical
  // inputs give identical outputs; DG_CFG_DEDUP=0 switches it off)
  bool cfg_pairs = false;
  bool temb_cache = true;     // DG_TEMB_CACHE=0: rebuild the time-embedding table on every call (round-2 start)
  bool finalized = false;     // LayerNorm folds are up to date with the loaded weights
  std::vector<float> temb_ts; // timesteps the rows of temb_table were computed for (empty: table invalid); every call of a sweep
                              // uses the same 50 timesteps, so the table (a function of timesteps and weights only) is built once
  // graphs
  bool use_graphs = true;
  // measurement only (dg_unet_set_family_mask): kernel families that run_forward enqueues (bit = 1 << Family).  With a family
  // switched off the result is garbage, the timing of the others is not: bench.py times the GEMM family alone as a replayed
  // graph (programmatic launch chain intact) instead of eager launches separated by events.
  int fam_mask = 15;
  cudaStream_t cap_stream = nullptr;
  struct CachedGraph { GraphKey key; cudaGraphExec_t exec; long long launches; };
  std::vector<CachedGraph> graphs;
  long long last_launches = 0;
};

namespace {

int dev_alloc(WeightStore* u, void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) return fail(DG_E_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  cudaMemset(*p, 0, bytes);
  u->owned.push_back(*p);
  return DG_OK;
}

int add_slot(WeightStore* u, const std::string& key, std::vector<int64_t> shape, PackKind kind, __half* dst,
             int64_t row_off = 0, int a = 0, int b = 0) {
  Slot s; s.key = key; s.shape = std::move(shape); s.kind = kind; s.dst = dst; s.row_off = row_off; s.a = a; s.b = b;
  u->slot_index[key] = (int)u->slots.size();
  u->slots.push_back(std::move(s));
  return DG_OK;
}

int make_norm(WeightStore* u, const std::string& pfx, int c, Norm* n) {
  n->c = c;
  DG_TRY(dev_alloc(u, (void**)&n->g, c * 2));
  DG_TRY(dev_alloc(u, (void**)&n->b, c * 2));
  add_slot(u, pfx + ".weight", {c}, PK_COPY, n->g);
  add_slot(u, pfx + ".bias", {c}, PK_COPY, n->b);
  return DG_OK;
}
int make_linear(WeightStore* u, const std::string& pfx, int in, int out, bool bias, Lin* l, bool conv1x1 = false) {
  l->in = in; l->out = out; l->rows = out;
  DG_TRY(dev_alloc(u, (void**)&l->w, (size_t)in * out * 2));
  if (conv1x1) add_slot(u, pfx + ".weight", {out, in, 1, 1}, PK_COPY, l->w);
  else add_slot(u, pfx + ".weight", {out, in}, PK_COPY, l->w);
  if (bias) {
    DG_TRY(dev_alloc(u, (void**)&l->b, out * 2));
    add_slot(u, pfx + ".bias", {out}, PK_COPY, l->b);
  }
  return DG_OK;
}
int make_conv3(WeightStore* u, const std::string& pfx, int in, int out, Lin* l, bool upsampler = false) {
  l->in = in; l->out = out; l->rows = out;
  DG_TRY(dev_alloc(u, (void**)&l->w, (size_t)in * 9 * out * 2));
  DG_TRY(dev_alloc(u, (void**)&l->b, out * 2));
  add_slot(u, pfx + ".weight", {out, in, 3, 3}, PK_CONV3, l->w, 0, out, in);
  if (upsampler) {
    DG_TRY(dev_alloc(u, (void**)&l->wph, (size_t)16 * in * out * 2));
    u->slots.back().dst2 = l->wph;
  }
  add_slot(u, pfx + ".bias", {out}, PK_COPY, l->b);
  return DG_OK;
}
int make_res(dg_unet* u, const std::string& pfx, int cin, int cout, Res* r, int* temb_off) {
  r->cin = cin; r->cout = cout;
  DG_TRY(make_norm(u, pfx + ".norm1", cin, &r->n1));
  DG_TRY(make_conv3(u, pfx + ".conv1", cin, cout, &r->c1));
  r->temb_off = *temb_off; *temb_off += cout;
  DG_TRY(make_norm(u, pfx + ".norm2", cout, &r->n2));
  DG_TRY(make_conv3(u, pfx + ".conv2", cout, cout, &r->c2));
  r->has_sc = cin != cout;
  if (r->has_sc) DG_TRY(make_linear(u, pfx + ".conv_shortcut", cin, cout, true, &r->sc, true));
  return DG_OK;
}
int make_fused_rows(WeightStore* u, const std::vector<std::string>& keys, int in, int out_each, Lin* l) {
  l->in = in; l->out = out_each * (int)keys.size(); l->rows = l->out;
  DG_TRY(dev_alloc(u, (void**)&l->w, (size_t)in * l->out * 2));
  for (size_t i = 0; i < keys.size(); ++i) add_slot(u, keys[i], {out_each, in}, PK_ROWS, l->w, (int64_t)i * out_each, out_each, in);
  return DG_OK;
}
int geglu_rows(int inner) { return ((inner + kGegluHalf - 1) / kGegluHalf) * kGegluTile; }
int make_xf(dg_unet* u, const std::string& pfx, int c, int heads, Xf* x) {
  const dg_unet_config& cf = u->cfg;
  x->c = c; x->heads = heads; x->ff_inner = 4 * c;
  const bool lin = cf.use_linear_projection != 0;
  DG_TRY(make_norm(u, pfx + ".norm", c, &x->gn));
  DG_TRY(make_linear(u, pfx + ".proj_in", c, c, true, &x->proj_in, !lin));
  const std::string tb = pfx + ".transformer_blocks.0";
  DG_TRY(make_norm(u, tb + ".norm1", c, &x->ln1));
  DG_TRY(make_fused_rows(u, {tb + ".attn1.to_q.weight", tb + ".attn1.to_k.weight", tb + ".attn1.to_v.weight"}, c, c, &x->qkv));
  DG_TRY(make_linear(u, tb + ".attn1.to_out.0", c, c, true, &x->o1));
  DG_TRY(make_norm(u, tb + ".norm2", c, &x->ln2));
  DG_TRY(make_fused_rows(u, {tb + ".attn2.to_q.weight"}, c, c, &x->q2));
  DG_TRY(make_fused_rows(u, {tb + ".attn2.to_k.weight", tb + ".attn2.to_v.weight"}, cf.cross_attention_dim, c, &x->kv2));
  DG_TRY(make_linear(u, tb + ".attn2.to_out.0", c, c, true, &x->o2));
  DG_TRY(make_norm(u, tb + ".norm3", c, &x->ln3));
  // GEGLU projection, packed in tiles of [value | gate]
  x->ff1.in = c; x->ff1.out = x->ff_inner; x->ff1.rows = geglu_rows(x->ff_inner);
  DG_TRY(dev_alloc(u, (void**)&x->ff1.w, (size_t)x->ff1.rows * c * 2));
  DG_TRY(dev_alloc(u, (void**)&x->ff1.b, (size_t)x->ff1.rows * 2));
  add_slot(u, tb + ".ff.net.0.proj.weight", {2 * x->ff_inner, c}, PK_GEGLU_W, x->ff1.w, 0, x->ff_inner, c);
  add_slot(u, tb + ".ff.net.0.proj.bias", {2 * x->ff_inner}, PK_GEGLU_B, x->ff1.b, 0, x->ff_inner, 1);
  DG_TRY(make_linear(u, tb + ".ff.net.2", x->ff_inner, c, true, &x->ff2));
  DG_TRY(make_linear(u, pfx + ".proj_out", c, c, true, &x->proj_out, !lin));
  return DG_OK;
}

int build_modules(dg_unet* u) {
  const dg_unet_config& cf = u->cfg;
  const int* ch = cf.block_out_channels;
  u->temb_dim = ch[0] * 4;
  // conv_in: im2col K padded to 64
  u->conv_in.in = 64; u->conv_in.out = ch[0]; u->conv_in.rows = ch[0];
  if (cf.in_channels * 9 > 64) return fail(DG_E_UNSUPPORTED, "in_channels %d too large for the conv_in gather", cf.in_channels);
  DG_TRY(dev_alloc(u, (void**)&u->conv_in.w, (size_t)64 * ch[0] * 2));
  DG_TRY(dev_alloc(u, (void**)&u->conv_in.b, ch[0] * 2));
  add_slot(u, "conv_in.weight", {ch[0], cf.in_channels, 3, 3}, PK_CONV_IN, u->conv_in.w, 0, ch[0], cf.in_channels);
  add_slot(u, "conv_in.bias", {ch[0]}, PK_COPY, u->conv_in.b);
  DG_TRY(make_linear(u, "time_embedding.linear_1", ch[0], u->temb_dim, true, &u->time1));
  DG_TRY(make_linear(u, "time_embedding.linear_2", u->temb_dim, u->temb_dim, true, &u->time2));

  int temb_off = 0;
  const int L = cf.layers_per_block;
  int cout = ch[0];
  u->down.resize(4);
  for (int i = 0; i < 4; ++i) {
    const int cin = cout; cout = ch[i];
    DownBlk& d = u->down[i];
    d.res.resize(L);
    if (cf.down_has_attn[i]) d.xf.resize(L);
    const std::string pfx = "down_blocks." + std::to_string(i);
    for (int j = 0; j < L; ++j) {
      DG_TRY(make_res(u, pfx + ".resnets." + std::to_string(j), j == 0 ? cin : cout, cout, &d.res[j], &temb_off));
      if (cf.down_has_attn[i]) DG_TRY(make_xf(u, pfx + ".attentions." + std::to_string(j), cout, cf.num_heads[i], &d.xf[j]));
    }
    d.has_down = i != 3;
    if (d.has_down) DG_TRY(make_conv3(u, pfx + ".downsamplers.0.conv", cout, cout, &d.down));
  }
  DG_TRY(make_res(u, "mid_block.resnets.0", ch[3], ch[3], &u->mid_r0, &temb_off));
  DG_TRY(make_xf(u, "mid_block.attentions.0", ch[3], cf.num_heads[3], &u->mid_xf));
  DG_TRY(make_res(u, "mid_block.resnets.1", ch[3], ch[3], &u->mid_r1, &temb_off));
  u->up.resize(4);
  cout = ch[3];
  for (int i = 0; i < 4; ++i) {
    const int prev = cout; cout = ch[3 - i];
    const int cin = ch[std::max(3 - i - 1, 0)];
    const bool attn = cf.down_has_attn[3 - i] != 0;
    UpBlk& b = u->up[i];
    b.res.resize(L + 1);
    if (attn) b.xf.resize(L + 1);
    const std::string pfx = "up_blocks." + std::to_string(i);
    for (int j = 0; j < L + 1; ++j) {
      const int skip = (j == L) ? cin : cout;
      const int rin = (j == 0) ? prev : cout;
      DG_TRY(make_res(u, pfx + ".resnets." + std::to_string(j), rin + skip, cout, &b.res[j], &temb_off));
      if (attn) DG_TRY(make_xf(u, pfx + ".attentions." + std::to_string(j), cout, cf.num_heads[3 - i], &b.xf[j]));
    }
    b.has_up = i != 3;
    if (b.has_up) DG_TRY(make_conv3(u, pfx + ".upsamplers.0.conv", cout, cout, &b.up, true));
  }
  DG_TRY(make_norm(u, "conv_norm_out", ch[0], &u->norm_out));
  DG_TRY(make_conv3(u, "conv_out", ch[0], cf.out_channels, &u->conv_out));

  // all 22 time_emb_proj layers live in one [sum(Cout), temb_dim] table so a single GEMV serves a forward
  u->temb_total = temb_off;
  DG_TRY(dev_alloc(u, (void**)&u->temb_proj_w, (size_t)temb_off * u->temb_dim * 2));
  DG_TRY(dev_alloc(u, (void**)&u->temb_proj_b, (size_t)temb_off * 2));
  auto reg = [&](const std::string& pfx, const Res& r) {
    add_slot(u, pfx + ".time_emb_proj.weight", {r.cout, u->temb_dim}, PK_ROWS, u->temb_proj_w, r.temb_off, r.cout, u->temb_dim);
    add_slot(u, pfx + ".time_emb_proj.bias", {r.cout}, PK_ROWS, u->temb_proj_b, r.temb_off, r.cout, 1);
  };
  for (int i = 0; i < 4; ++i)
    for (size_t j = 0; j < u->down[i].res.size(); ++j) reg("down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), u->down[i].res[j]);
  reg("mid_block.resnets.0", u->mid_r0);
  reg("mid_block.resnets.1", u->mid_r1);
  for (int i = 0; i < 4; ++i)
    for (size_t j = 0; j < u->up[i].res.size(); ++j) reg("up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), u->up[i].res[j]);
  return DG_OK;
}

// ================================================================== forward
__global__ void dup_latents_kernel(const __half* __restrict__ lat, __half* __restrict__ dst, size_t n_vec8, int copies);

struct Fwd {
  dg_unet* u; cudaStream_t s; int sms; int B; int tokens; const __half* ehs; __half* temb_all;
  int err = DG_OK;
  int temb_ld = 0;      // row pitch of temb_all (0: one row shared by every sample)
  int xf_index = 0;     // transformer blocks visited so far (indexes kv_cache)

  __half* alloc(size_t bytes) {
    void* p = u->arena.alloc(bytes);
    if (!p && err == DG_OK) err = fail(DG_E_NOMEM, "activation arena exhausted (%zu bytes requested); call dg_unet_prepare with a larger batch", bytes);
    return (__half*)p;
  }
  T4 talloc(int B_, int H, int W, int C) { T4 t; t.B = B_; t.H = H; t.W = W; t.C = C; t.p = alloc(t.bytes()); return t; }
  void free_(T4& t) { u->arena.release(t.p); t.p = nullptr; }
  void free_(__half* p) { u->arena.release(p); }

#define FW(expr) do { if (err == DG_OK) err = (expr); } while (0)

  bool on(int fam) const { return (u->fam_mask >> fam) & 1; }
  bool fuse_gn() const { return u->gn_blk > 0; }
  bool fuse_ln() const { return u->ctx->fuse_ln != 0; }
  // fused GroupNorm block sums [B_][H*W/32][C/blk] float2 for a tensor about to be produced by a 3x3 conv (conv = true:
  // pixel-box tiles) or a plain GEMM over its B_*H*W rows; nullptr when the shape cannot carry them (the consumer then
  // runs the stand-alone statistics kernel)
  float* gn_alloc(int B_, int H, int W, int C, bool conv) {
    if (!fuse_gn() || !gn_stats_supported(H * W, conv ? W : 0, conv ? H : 0, u->gn_blk, C)) return nullptr;
    const size_t n = (size_t)B_ * (H * W / 32) * (C / u->gn_blk) * 2;
    if (u->gn_off + n > u->gn_cap) { if (err == DG_OK) err = fail(DG_E_NOMEM, "GroupNorm statistics arena exhausted"); return nullptr; }
    float* p = u->gn_arena + u->gn_off;
    u->gn_off += (n + 3) & ~size_t(3);
    return p;
  }
  // LayerNorm row partials [rows][gemm_row_parts(C)][2] for a token matrix about to be produced
  float* ln_alloc(int rows, int C) {
    if (!fuse_ln()) return nullptr;
    const size_t n = (size_t)rows * gemm_row_parts(C) * 2;
    if (u->ln_off + n > u->ln_cap) { if (err == DG_OK) err = fail(DG_E_NOMEM, "LayerNorm statistics arena exhausted"); return nullptr; }
    float* p = u->ln_arena + u->ln_off;
    u->ln_off += (n + 3) & ~size_t(3);
    return p;
  }

  // GroupNorm (+ SiLU) applied inside the consuming conv (gemm2_kernel<..., kXf>): possible when both sources carry fused
  // block sums.  Returns the (mean_h, scale', shift') table the conv reads, or nullptr (the caller then runs the stand-alone pass).
  const __half* gn_fold(const T4& x0, const T4* x1, const Norm& n, float eps, int silu) {
    if (!u->ctx->fuse_xf || u->ctx->gemm.cta_mode == 1 || !fuse_gn() || !x0.gst || (x1 && !x1->gst)) return nullptr;
    const int C = x0.C + (x1 ? x1->C : 0);
    if (C / u->gn_blk > 256 || (x0.H * x0.W) % 32) return nullptr;
    const size_t nh = (size_t)x0.B * 3 * C;
    if (u->xf_off + nh > u->xf_cap) { if (err == DG_OK) err = fail(DG_E_NOMEM, "GroupNorm table arena exhausted"); return nullptr; }
    __half* tab = u->xf_arena + u->xf_off;
    u->xf_off += (nh + 7) & ~size_t(7);
    if (on(FAM_NORM))
      FW(launch_gn_fold(s, x0.C, x0.gst, x1 ? x1->C : 0, x1 ? x1->gst : nullptr, u->gn_blk, n.g, n.b, tab, x0.B, x0.H * x0.W,
                        u->cfg.norm_num_groups, eps, silu));
    return tab;
  }
  // per-(sample, group) totals of one GroupNorm (two-step apply): a slice of the table arena
  float* gt_alloc(int B_) {
    const size_t nh = (size_t)B_ * u->cfg.norm_num_groups * 2 * 2;     // floats, counted in halves
    if (u->xf_off + nh > u->xf_cap) return nullptr;
    float* p = reinterpret_cast<float*>(u->xf_arena + u->xf_off);
    u->xf_off += (nh + 7) & ~size_t(7);
    return p;
  }
  void gn(const T4& x0, const T4* x1, const Norm& n, float eps, int silu, T4& out) {
    if (!on(FAM_NORM)) return;
    if (fuse_gn() && x0.gst && (!x1 || x1->gst) && (x0.C + (x1 ? x1->C : 0)) / u->gn_blk <= 256)
      FW(launch_groupnorm_fused(s, sms, x0.p, x0.C, x0.gst, x1 ? x1->p : nullptr, x1 ? x1->C : 0, x1 ? x1->gst : nullptr,
                                u->gn_blk, n.g, n.b, out.p, x0.B, x0.H * x0.W, u->cfg.norm_num_groups, eps, silu, gt_alloc(x0.B)));
    else
      FW(launch_groupnorm(s, sms, x0.p, x0.C, x1 ? x1->p : nullptr, x1 ? x1->C : 0, n.g, n.b, out.p, u->gn_stats, x0.B,
                          x0.H * x0.W, u->cfg.norm_num_groups, eps, silu));
  }
  // 3x3 conv; `out.gst` (if the caller allocated it) receives the fused GroupNorm sums of the result
  void conv3(const T4& x, const Lin& w, const __half* rowvec, const __half* residual, T4& out, const T4* x1 = nullptr,
             const __half* xf_tab = nullptr) {
    GemmArgs a; a.a0 = x.p; a.c0 = x.C; a.B = x.B; a.H = x.H; a.W = x.W; a.taps = 9; a.w = w.w; a.n_w = w.rows;
    if (x1) { a.a1 = x1->p; a.c1 = x1->C; }
    a.xf_tab = xf_tab; a.xf_silu = 1;
    a.n_out = w.out; a.bias = w.b; a.rowvec = rowvec; a.ld_rowvec = temb_ld; a.residual = residual; a.ld_res = w.out;
    a.out = out.p; a.ldo = w.out; a.gn_stats_out = out.gst; a.gn_blk = u->gn_blk;
    if (on(FAM_GEMM)) FW(launch_gemm(s, u->ctx->gemm, a));
  }
  // Upsample2D: nearest x2 + conv3x3 as four 2x2 phase convolutions on the low-resolution input (pack_upconv_kernel): 2.25x
  // fewer multiply-adds than the 9-tap conv on the upsampled tensor, which is never materialised.  `out` is [B, 2H, 2W, N]; its
  // fused GroupNorm sums (if allocated) are filled phase by phase (a quarter of each sample's slabs per launch).
  void upconv(const T4& x, const Lin& w, T4& out) {
    const bool stats = out.gst && gn_stats_supported(x.H * x.W, x.W, x.H, u->gn_blk, w.out);
    if (!stats) out.gst = nullptr;              // the consumer falls back to the stand-alone statistics kernel
    for (int ph = 0; ph < 4; ++ph) {
      GemmArgs a; a.a0 = x.p; a.c0 = x.C; a.B = x.B; a.H = x.H; a.W = x.W; a.taps = 4; a.phase_y = ph >> 1; a.phase_x = ph & 1;
      a.w = w.wph + (size_t)ph * w.out * 4 * w.in; a.n_w = w.rows; a.n_out = w.out; a.bias = w.b; a.out = out.p; a.ldo = w.out;
      if (stats) { a.gn_stats_out = out.gst; a.gn_blk = u->gn_blk; a.gn_slot0 = ph * (x.H * x.W / 32); a.gn_slots = 4 * (x.H * x.W / 32); }
      if (on(FAM_GEMM)) FW(launch_gemm(s, u->ctx->gemm, a));
    }
  }
  struct LinOpt {
    const __half* residual = nullptr; int geglu = 0;
    float* gn_out = nullptr; int hw = 0;         // fused GroupNorm sums of the result (rows per sample = hw)
    float* rows_out = nullptr;                   // LayerNorm row partials of the result
    const float* ln_in = nullptr; int ln_c = 0;  // LayerNorm-folded GEMM: row partials of the input (w must carry wf/cs/b32)
    const __half* xf_tab = nullptr;              // GroupNorm (no activation) applied to the input inside the GEMM (needs hw)
  };
  // plain GEMM over rows = B*H*W of x (optionally 2-source concat along channels)
  void linear(const __half* x0, int c0, const __half* x1, int c1, int rows, const Lin& w, __half* out, const LinOpt& o) {
    GemmArgs a; a.a0 = x0; a.c0 = c0; a.a1 = x1; a.c1 = c1; a.B = 1; a.H = 1; a.W = rows; a.taps = 1; a.w = w.w; a.n_w = w.rows;
    a.n_out = w.out; a.bias = w.b; a.residual = o.residual; a.ld_res = w.out; a.geglu = o.geglu; a.out = out; a.ldo = w.out;
    a.gn_stats_out = o.gn_out; a.gn_blk = u->gn_blk; a.hw = o.hw; a.row_stats_out = o.rows_out;
    a.xf_tab = o.xf_tab; a.xf_silu = 0;
    if (o.ln_in) {
      a.w = w.wf; a.bias = nullptr; a.bias32 = w.b32; a.colsum = w.cs; a.ln_stats = o.ln_in; a.ln_parts = gemm_row_parts(o.ln_c);
      a.ln_c = o.ln_c; a.ln_eps = 1e-5f;
    }
    if (on(FAM_GEMM)) FW(launch_gemm(s, u->ctx->gemm, a));
  }
  void linear(const __half* x0, int c0, const __half* x1, int c1, int rows, const Lin& w, const __half* residual, int geglu,
              __half* out) {
    LinOpt o; o.residual = residual; o.geglu = geglu;
    linear(x0, c0, x1, c1, rows, w, out, o);
  }

  // [Bh, ...] -> [2 Bh, ...]: both halves = the source (fp16 tensors; fp32 statistics are copied as raw 16-byte vectors)
  void dup_rows(const void* src, void* dst, size_t bytes) {
    if (!src || !dst || err != DG_OK || !on(FAM_OTHER)) return;
    dup_latents_kernel<<<grid_for(bytes / 16, 256, sms), 256, 0, s>>>((const __half*)src, (__half*)dst, bytes / 16, 2);
    ++g_launch_counter;
    if (cudaGetLastError() != cudaSuccess) err = fail(DG_E_CUDA, "dup_rows launch failed");
  }
  T4 expand2(const T4& t, bool conv_stats) {
    T4 o = talloc(2 * t.B, t.H, t.W, t.C);
    dup_rows(t.p, o.p, t.bytes());
    if (t.gst) {
      o.gst = gn_alloc(2 * t.B, t.H, t.W, t.C, conv_stats);
      if (o.gst) dup_rows(t.gst, o.gst, (size_t)t.B * (t.H * t.W / 32) * (t.C / u->gn_blk) * 2 * sizeof(float));
    }
    return o;
  }

  T4 resnet(const Res& r, const T4& x0, const T4* x1) {
    const int B_ = x0.B, H = x0.H, W = x0.W;
    // conv1(silu(norm1(x))): normalisation + activation inside the conv's operand path when the statistics are fused
    T4 h1 = talloc(B_, H, W, r.cout);
    h1.gst = gn_alloc(B_, H, W, r.cout, true);
    if (const __half* tab1 = gn_fold(x0, x1, r.n1, u->cfg.norm_eps, 1)) {
      conv3(x0, r.c1, temb_all + r.temb_off, nullptr, h1, x1, tab1);
    } else {
      T4 hn = talloc(B_, H, W, r.cin);
      gn(x0, x1, r.n1, u->cfg.norm_eps, 1, hn);
      conv3(hn, r.c1, temb_all + r.temb_off, nullptr, h1);
      free_(hn);
    }
    T4 out = talloc(B_, H, W, r.cout);
    out.gst = gn_alloc(B_, H, W, r.cout, true);
    const __half* resid = x0.p;
    T4 sc{};
    const bool sc_in_conv = r.has_sc && u->ctx->fuse_sc && r.c2x.w;
    if (r.has_sc && !sc_in_conv) {
      sc = talloc(B_, H, W, r.cout);
      linear(x0.p, x0.C, x1 ? x1->p : nullptr, x1 ? x1->C : 0, B_ * H * W, r.sc, nullptr, 0, sc.p);
      resid = sc.p;
    }
    auto conv2 = [&](const T4& in, const __half* tab) {
      if (sc_in_conv) {
        // out = conv2(in) + conv_shortcut(cat[x0, x1]) + (b2 + bs): the 1x1 shortcut is two more sources of conv2's K loop
        GemmArgs a; a.a0 = in.p; a.c0 = in.C; a.B = B_; a.H = H; a.W = W; a.taps = 9; a.w = r.c2x.w; a.n_w = r.c2x.rows; a.n_out = r.c2x.out;
        a.bias = r.c2x.b; a.x0 = x0.p; a.cx0 = x0.C; a.x1 = x1 ? x1->p : nullptr; a.cx1 = x1 ? x1->C : 0;
        a.out = out.p; a.ldo = r.c2x.out; a.gn_stats_out = out.gst; a.gn_blk = u->gn_blk; a.xf_tab = nullptr;
        (void)tab;
        if (on(FAM_GEMM)) FW(launch_gemm(s, u->ctx->gemm, a));
      } else {
        conv3(in, r.c2, nullptr, resid, out, nullptr, tab);
      }
    };
    const __half* tab2 = sc_in_conv ? nullptr : gn_fold(h1, nullptr, r.n2, u->cfg.norm_eps, 1);
    if (tab2) {
      conv2(h1, tab2);
    } else {
      T4 h2n = talloc(B_, H, W, r.cout);
      gn(h1, nullptr, r.n2, u->cfg.norm_eps, 1, h2n);
      conv2(h2n, nullptr);
      free_(h2n);
    }
    free_(h1);
    if (sc.p) free_(sc);
    return out;
  }

  // `expand`: `in_` holds B/2 samples whose duplicates form the real batch (cfg_pairs): everything up to the self-attention
  // out-projection runs on the half, then h (+ its LayerNorm partials) and the block input are duplicated
  T4 transformer(const Xf& x, const T4& in_, bool expand = false) {
    T4 in = in_;
    int B_ = in.B, rows = B_ * in.H * in.W;
    const int S = in.H * in.W, C = x.c;
    const int d = C / x.heads;
    const bool fl = fuse_ln();
    T4 xn = talloc(B_, in.H, in.W, C);      // (also the attention outputs' buffer below)
    T4 h = talloc(B_, in.H, in.W, C);
    float* rs = ln_alloc(rows, C);
    // proj_in(norm(x)): the GroupNorm (eps 1e-6, no activation) inside the GEMM's operand path when the statistics are fused
    if (const __half* tabn = u->ctx->fuse_xf >= 2 ? gn_fold(in, nullptr, x.gn, 1e-6f, 0) : nullptr) {
      LinOpt o; o.rows_out = rs; o.xf_tab = tabn; o.hw = S; linear(in.p, C, nullptr, 0, rows, x.proj_in, h.p, o);
    } else {
      gn(in, nullptr, x.gn, 1e-6f, 0, xn);
      LinOpt o; o.rows_out = rs; linear(xn.p, C, nullptr, 0, rows, x.proj_in, h.p, o);
    }
    // self-attention: LayerNorm folded into the QKV projection (or a stand-alone LayerNorm pass when fusion is off)
    __half* qkv = alloc((size_t)rows * 3 * C * 2);
    if (fl) { LinOpt o; o.ln_in = rs; o.ln_c = C; linear(h.p, C, nullptr, 0, rows, x.qkv, qkv, o); }
    else {
      if (on(FAM_NORM)) FW(launch_layernorm(s, h.p, x.ln1.g, x.ln1.b, xn.p, rows, C, 1e-5f));
      linear(xn.p, C, nullptr, 0, rows, x.qkv, nullptr, 0, qkv);
    }
    if (err == DG_OK && on(FAM_ATTN)) FW(launch_attention(s, qkv, 3 * C, qkv + C, 3 * C, qkv + 2 * C, 3 * C, xn.p, B_, x.heads, S, S, d));
    free_(qkv);
    rs = ln_alloc(rows, C);
    { LinOpt o; o.residual = h.p; o.rows_out = rs; linear(xn.p, C, nullptr, 0, rows, x.o1, h.p, o); }
    T4 in_full{};
    if (expand) {
      T4 h2 = talloc(2 * B_, in.H, in.W, C);
      dup_rows(h.p, h2.p, h.bytes());
      float* rs2 = ln_alloc(2 * rows, C);
      if (rs && rs2) dup_rows(rs, rs2, (size_t)rows * gemm_row_parts(C) * 2 * sizeof(float));
      in_full = talloc(2 * B_, in.H, in.W, C);
      dup_rows(in.p, in_full.p, in.bytes());
      free_(h); free_(xn);
      h = h2; rs = rs2; in = in_full;
      B_ *= 2; rows *= 2;
      xn = talloc(B_, in.H, in.W, C);
    }
    // cross-attention
    __half* q = alloc((size_t)rows * C * 2);
    if (fl) { LinOpt o; o.ln_in = rs; o.ln_c = C; linear(h.p, C, nullptr, 0, rows, x.q2, q, o); }
    else {
      if (on(FAM_NORM)) FW(launch_layernorm(s, h.p, x.ln2.g, x.ln2.b, xn.p, rows, C, 1e-5f));
      linear(xn.p, C, nullptr, 0, rows, x.q2, nullptr, 0, q);
    }
    __half* kv;
    const int xi = xf_index++;
    if (u->kv_ready) {
      kv = u->kv_cache[xi];      // projected once per prompt batch by dg_denoise_loop
    } else {
      kv = alloc((size_t)B_ * tokens * 2 * C * 2);
      linear(ehs, u->cfg.cross_attention_dim, nullptr, 0, B_ * tokens, x.kv2, nullptr, 0, kv);
    }
    if (err == DG_OK && on(FAM_ATTN)) FW(launch_attention(s, q, C, kv, 2 * C, kv + C, 2 * C, xn.p, B_, x.heads, S, tokens, d));
    free_(q);
    if (!u->kv_ready) free_(kv);
    rs = ln_alloc(rows, C);
    { LinOpt o; o.residual = h.p; o.rows_out = rs; linear(xn.p, C, nullptr, 0, rows, x.o2, h.p, o); }
    // feed-forward (GEGLU)
    __half* g = alloc((size_t)rows * x.ff_inner * 2);
    if (fl) { LinOpt o; o.ln_in = rs; o.ln_c = C; o.geglu = 1; linear(h.p, C, nullptr, 0, rows, x.ff1, g, o); }
    else {
      if (on(FAM_NORM)) FW(launch_layernorm(s, h.p, x.ln3.g, x.ln3.b, xn.p, rows, C, 1e-5f));
      linear(xn.p, C, nullptr, 0, rows, x.ff1, nullptr, 1, g);
    }
    linear(g, x.ff_inner, nullptr, 0, rows, x.ff2, h.p, 0, h.p);
    free_(g);
    // proj_out + residual with the block input
    T4 out = talloc(B_, in.H, in.W, C);
    out.gst = gn_alloc(B_, in.H, in.W, C, false);
    { LinOpt o; o.residual = in.p; o.gn_out = out.gst; o.hw = S; linear(h.p, C, nullptr, 0, rows, x.proj_out, out.p, o); }
    free_(xn); free_(h);
    if (in_full.p) free_(in_full);
    return out;
  }
};

int run_forward(dg_unet* u, cudaStream_t s, const __half* sample, const __half* ehs, int tokens, __half* out, int B, int h, int w) {
  const dg_unet_config& cf = u->cfg;
  const int sms = u->ctx->num_sms;
  u->arena.reset();
  u->gn_off = 0; u->ln_off = 0; u->xf_off = 0;
  Fwd f{u, s, sms, B, tokens, ehs, nullptr};

  // ---- time embedding: sinusoid -> MLP -> all time_emb_proj(SiLU(emb)) in one GEMV
  const int c0 = cf.block_out_channels[0];
  if (u->temb_ready) {
    f.temb_all = u->temb_cur; f.temb_ld = 0;     // hoisted: every sample shares the current step's row
  } else {
  __half* tsin = f.alloc((size_t)B * c0 * 2);
  __half* t1 = f.alloc((size_t)B * u->temb_dim * 2);
  __half* emb = f.alloc((size_t)B * u->temb_dim * 2);
  f.temb_all = f.alloc((size_t)B * u->temb_total * 2);
  f.temb_ld = u->temb_total;
  if (f.err) return f.err;
  if (f.on(FAM_OTHER)) {
  timestep_sinusoid_kernel<<<(B * c0 / 2 + 127) / 128, 128, 0, s>>>(u->d_t, tsin, B, c0, cf.freq_shift, cf.flip_sin_to_cos);
  DG_LAUNCH_CHECK();
  }
  for (int b0 = 0; b0 < B && f.on(FAM_OTHER); b0 += 8) {
    const int nb = std::min(8, B - b0);
    DG_TRY(launch_gemv(s, tsin + (size_t)b0 * c0, c0, u->time1.w, u->time1.b, t1 + (size_t)b0 * u->temb_dim, u->temb_dim, nb, u->temb_dim, c0, 0, 1));
    DG_TRY(launch_gemv(s, t1 + (size_t)b0 * u->temb_dim, u->temb_dim, u->time2.w, u->time2.b, emb + (size_t)b0 * u->temb_dim, u->temb_dim, nb, u->temb_dim, u->temb_dim, 0, 0));
    DG_TRY(launch_gemv(s, emb + (size_t)b0 * u->temb_dim, u->temb_dim, u->temb_proj_w, u->temb_proj_b, f.temb_all + (size_t)b0 * u->temb_total, u->temb_total, nb, u->temb_total, u->temb_dim, 1, 0));
  }
  }

  // ---- conv_in (4-channel NCHW gather -> K=64 GEMM)
  // cfg_pairs: the two halves of the batch are identical until the first cross-attention -- the prefix runs on the first half
  const bool dedup = u->cfg_pairs && u->ctx->cfg_dedup && B % 2 == 0 && !u->down[0].xf.empty() && u->fam_mask == 15;
  const int Bp = dedup ? B / 2 : B;
  __half* col = f.alloc((size_t)Bp * h * w * 64 * 2);
  T4 x = f.talloc(Bp, h, w, c0);
  if (f.err) return f.err;
  {
    const size_t items = (size_t)Bp * h * w * 64;
    if (f.on(FAM_OTHER)) {
      im2col_conv_in_kernel<<<grid_for(items, 256, sms), 256, 0, s>>>(sample, col, Bp, cf.in_channels, h, w, 64);
      DG_LAUNCH_CHECK();
    }
    x.gst = f.gn_alloc(Bp, h, w, c0, false);
    { Fwd::LinOpt o; o.gn_out = x.gst; o.hw = h * w; f.linear(col, 64, nullptr, 0, Bp * h * w, u->conv_in, x.p, o); }
    f.free_(col);
  }
  std::vector<T4> skips;
  if (!dedup) skips.push_back(x);
  // ---- down
  for (int i = 0; i < 4 && !f.err; ++i) {
    DownBlk& d = u->down[i];
    for (size_t j = 0; j < d.res.size(); ++j) {
      const bool half = dedup && i == 0 && j == 0;           // x still holds B/2 samples
      T4 y = f.resnet(d.res[j], x, nullptr);
      if (half) {                                             // the skip connection of conv_in's output needs the whole batch
        T4 xf_ = f.expand2(x, false);
        f.free_(x);
        skips.push_back(xf_);
      }
      if (!d.xf.empty()) { T4 z = f.transformer(d.xf[j], y, half); f.free_(y); y = z; }
      x = y; skips.push_back(x);
    }
    if (d.has_down) {
      const int Ho = (x.H - 1) / 2 + 1, Wo = (x.W - 1) / 2 + 1;
      T4 y = f.talloc(x.B, Ho, Wo, x.C);
      if (f.err) break;
      if (u->ctx->up_phases && u->ctx->gemm.cta_mode != 1 && x.H % 2 == 0 && x.W % 2 == 0) {
        // Downsample2D (conv3x3, stride 2, pad 1): the 9-tap implicit GEMM over the OUTPUT grid, its operand boxes taking every
        // second input pixel through a strided tensor map (no materialised im2col tensor)
        y.gst = f.gn_alloc(x.B, Ho, Wo, x.C, true);
        GemmArgs a; a.a0 = x.p; a.c0 = x.C; a.B = x.B; a.H = Ho; a.W = Wo; a.taps = 9; a.in_stride = 2; a.w = d.down.w; a.n_w = d.down.rows;
        a.n_out = d.down.out; a.bias = d.down.b; a.out = y.p; a.ldo = d.down.out; a.gn_stats_out = y.gst; a.gn_blk = u->gn_blk;
        if (f.on(FAM_GEMM)) f.err = launch_gemm(s, u->ctx->gemm, a);
      } else {
        __half* c2 = f.alloc((size_t)x.B * Ho * Wo * 9 * x.C * 2);
        if (f.err) break;
        const size_t items = (size_t)x.B * Ho * Wo * 9 * (x.C / 8);
        if (f.on(FAM_OTHER)) {
          im2col3x3_nhwc_kernel<<<grid_for(items, 256, sms), 256, 0, s>>>(x.p, c2, x.B, x.H, x.W, x.C, 2, Ho, Wo);
          DG_LAUNCH_CHECK();
        }
        y.gst = f.gn_alloc(x.B, Ho, Wo, x.C, false);
        { Fwd::LinOpt o; o.gn_out = y.gst; o.hw = Ho * Wo; f.linear(c2, 9 * x.C, nullptr, 0, x.B * Ho * Wo, d.down, y.p, o); }
        f.free_(c2);
      }
      x = y; skips.push_back(x);
    }
  }
  // ---- mid
  if (!f.err) {
    T4 a = f.resnet(u->mid_r0, x, nullptr);
    T4 b = f.transformer(u->mid_xf, a); f.free_(a);
    T4 c = f.resnet(u->mid_r1, b, nullptr); f.free_(b);
    x = c;  // the last skip (== previous x) stays alive on the stack
  }
  // ---- up
  for (int i = 0; i < 4 && !f.err; ++i) {
    UpBlk& b = u->up[i];
    for (size_t j = 0; j < b.res.size(); ++j) {
      T4 skip = skips.back(); skips.pop_back();
      T4 y = f.resnet(b.res[j], x, &skip);
      f.free_(x); f.free_(skip);
      if (!b.xf.empty()) { T4 z = f.transformer(b.xf[j], y); f.free_(y); y = z; }
      x = y;
    }
    if (b.has_up) {
      T4 y = f.talloc(x.B, x.H * 2, x.W * 2, x.C);
      if (f.err) break;
      y.gst = f.gn_alloc(x.B, x.H * 2, x.W * 2, x.C, true);
      if (u->ctx->up_phases && u->ctx->gemm.cta_mode != 1) {
        f.upconv(x, b.up, y);
      } else {
        T4 upx = f.talloc(x.B, x.H * 2, x.W * 2, x.C);
        if (f.err) break;
        const size_t items = (size_t)x.B * 4 * x.H * x.W * (x.C / 8);
        if (f.on(FAM_OTHER)) {
          upsample2x_nhwc_kernel<<<grid_for(items, 256, sms), 256, 0, s>>>(x.p, upx.p, x.B, x.H, x.W, x.C);
          DG_LAUNCH_CHECK();
        }
        f.conv3(upx, b.up, nullptr, nullptr, y);
        f.free_(upx);
      }
      f.free_(x);
      x = y;
    }
  }
  if (f.err) return f.err;
  // ---- out: conv_out(silu(conv_norm_out(x))), the normalisation inside the conv when the statistics are fused
  // conv_out writes rows padded to 8 channels (TMA store needs a 16-byte row pitch); the NCHW exit kernel reads that pitch
  const int opitch = (cf.out_channels + 7) / 8 * 8;
  T4 o = f.talloc(B, h, w, opitch);
  {
    const __half* tabo = f.gn_fold(x, nullptr, u->norm_out, cf.norm_eps, 1);
    T4 xn{};
    if (!tabo) {
      xn = f.talloc(B, h, w, c0);
      f.gn(x, nullptr, u->norm_out, cf.norm_eps, 1, xn);
    }
    GemmArgs a; a.a0 = tabo ? x.p : xn.p; a.c0 = c0; a.B = B; a.H = h; a.W = w; a.taps = 9; a.w = u->conv_out.w; a.n_w = u->conv_out.rows;
    a.n_out = cf.out_channels; a.bias = u->conv_out.b; a.out = o.p; a.ldo = opitch; a.xf_tab = tabo; a.xf_silu = 1;
    if (f.err == DG_OK && f.on(FAM_GEMM)) f.err = launch_gemm(s, u->ctx->gemm, a);
  }
  if (f.err) return f.err;
  if (f.on(FAM_OTHER)) {
    const size_t n = (size_t)B * cf.out_channels * h * w;
    nhwc_to_nchw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(o.p, out, B, cf.out_channels, h * w, opitch);
    DG_LAUNCH_CHECK();
  }
  return DG_OK;
}

__global__ void set_timesteps_kernel(float* dst, int n, float t0, float t1, float t2, float t3, float t4, float t5, float t6, float t7, int bcast) {
  const float v[8] = {t0, t1, t2, t3, t4, t5, t6, t7};
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = bcast ? t0 : v[i & 7];
}
__global__ void set_step_kernel(int* step, float* d_t, const float* t_table, int idx, int n, const __half* __restrict__ temb_table,
                                __half* __restrict__ temb_cur, int temb_total) {
  if (threadIdx.x == 0) *step = idx;
  for (int i = threadIdx.x; i < n; i += blockDim.x) d_t[i] = t_table[idx];
  if (temb_table)   // this step's time-embedding projections become the row the forward graph reads
    for (int i = threadIdx.x; i < temb_total / 8; i += blockDim.x)
      reinterpret_cast<Half8*>(temb_cur)[i] = reinterpret_cast<const Half8*>(temb_table + (size_t)idx * temb_total)[i];
}
__global__ void dup_latents_kernel(const __half* __restrict__ lat, __half* __restrict__ dst, size_t n_vec8, int copies) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec8; i += (size_t)gridDim.x * blockDim.x) {
    const Half8 v = *reinterpret_cast<const Half8*>(lat + i * 8);
    for (int c = 0; c < copies; ++c) *reinterpret_cast<Half8*>(dst + (c * n_vec8 + i) * 8) = v;
  }
}

int forward_maybe_graph(dg_unet* u, cudaStream_t s, const __half* sample, const __half* ehs, int tokens, __half* out, int B, int h, int w) {
  if (!u->use_graphs) {
    const long long c0 = g_launch_counter;
    DG_TRY(run_forward(u, s, sample, ehs, tokens, out, B, h, w));
    u->last_launches += g_launch_counter - c0;
    return DG_OK;
  }
  GraphKey key{B, h, w, tokens, sample, ehs, out, (u->kv_ready ? 1 : 0) | (u->temb_ready ? 2 : 0) | (u->cfg_pairs ? 4 : 0), u->fam_mask};
  for (auto& g : u->graphs) {
    if (g.key == key) {
      DG_CUDA(cudaGraphLaunch(g.exec, s));
      u->last_launches += g.launches;
      return DG_OK;
    }
  }
  // capture
  // Capture on a private stream (the caller's stream may be the legacy default stream, which cannot capture); the
  // instantiated graph is then launched into the caller's stream.
  cudaGraph_t graph = nullptr;
  const long long c0 = g_launch_counter;
  if (!u->cap_stream) DG_CUDA(cudaStreamCreateWithFlags(&u->cap_stream, cudaStreamNonBlocking));
  cudaStream_t cs = u->cap_stream;
  DG_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
  int r = run_forward(u, cs, sample, ehs, tokens, out, B, h, w);
  cudaError_t e = cudaStreamEndCapture(cs, &graph);
  if (r != DG_OK) { if (graph) cudaGraphDestroy(graph); return r; }
  if (e != cudaSuccess) return fail(DG_E_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
  cudaGraphExec_t exec = nullptr;
  DG_CUDA(cudaGraphInstantiate(&exec, graph, 0));
  cudaGraphDestroy(graph);
  if (u->graphs.size() >= 8) { cudaGraphExecDestroy(u->graphs.front().exec); u->graphs.erase(u->graphs.begin()); }
  u->graphs.push_back({key, exec, g_launch_counter - c0});
  DG_CUDA(cudaGraphLaunch(exec, s));
  u->last_launches += g_launch_counter - c0;
  return DG_OK;
}

// LayerNorm folds of the three GEMMs that consume a LayerNorm output in every transformer block (qkv, to_q of the
// cross-attention, GEGLU projection).  Runs once after load_state_dict, outside any graph capture.
int fold_lin(dg_unet* u, Lin* l, const Norm& n) {
  if (!l->wf) {
    DG_TRY(dev_alloc(u, (void**)&l->wf, (size_t)l->rows * l->in * 2));
    DG_TRY(dev_alloc(u, (void**)&l->cs, (size_t)l->rows * 4));
    DG_TRY(dev_alloc(u, (void**)&l->b32, (size_t)l->rows * 4));
  }
  fold_layernorm_kernel<<<(l->rows + 7) / 8, 256>>>(l->w, l->b, n.g, n.b, l->wf, l->cs, l->b32, l->rows, l->in);
  DG_LAUNCH_CHECK();
  return DG_OK;
}
int finalize_weights(dg_unet* u) {
  if (u->finalized) return DG_OK;
  auto fold_xf = [&](Xf& x) -> int {
    DG_TRY(fold_lin(u, &x.qkv, x.ln1));
    DG_TRY(fold_lin(u, &x.q2, x.ln2));
    DG_TRY(fold_lin(u, &x.ff1, x.ln3));
    return DG_OK;
  };
  for (auto& d : u->down) for (auto& x : d.xf) DG_TRY(fold_xf(x));
  DG_TRY(fold_xf(u->mid_xf));
  for (auto& b : u->up) for (auto& x : b.xf) DG_TRY(fold_xf(x));
  // conv_shortcut inside conv2: [conv2 | shortcut] weight rows and the summed bias
  auto fuse_sc = [&](Res& r) -> int {
    if (!r.has_sc) return DG_OK;
    const int k0 = 9 * r.cout, k1 = r.cin;
    if (!r.c2x.w) {
      DG_TRY(dev_alloc(u, (void**)&r.c2x.w, (size_t)r.cout * (k0 + k1) * 2));
      DG_TRY(dev_alloc(u, (void**)&r.c2x.b, (size_t)r.cout * 2));
    }
    r.c2x.in = r.cout; r.c2x.out = r.cout; r.c2x.rows = r.cout;
    concat_weight_rows_kernel<<<grid_for((size_t)r.cout * (k0 + k1), 256, u->ctx->num_sms), 256>>>(r.c2.w, k0, r.sc.w, k1, r.c2x.w, r.cout);
    DG_LAUNCH_CHECK();
    add_bias_kernel<<<(r.cout + 255) / 256, 256>>>(r.c2.b, r.sc.b, r.c2x.b, r.cout);
    DG_LAUNCH_CHECK();
    return DG_OK;
  };
  for (auto& d : u->down) for (auto& r : d.res) DG_TRY(fuse_sc(r));
  DG_TRY(fuse_sc(u->mid_r0)); DG_TRY(fuse_sc(u->mid_r1));
  for (auto& b : u->up) for (auto& r : b.res) DG_TRY(fuse_sc(r));
  DG_CUDA(cudaDeviceSynchronize());
  for (auto& g : u->graphs) cudaGraphExecDestroy(g.exec);   // captured graphs point at stale folds
  u->graphs.clear();
  u->finalized = true;
  u->temb_ts.clear();         // the time-embedding table was computed from the previous weights
  return DG_OK;
}

// Every transformer block in the order run_forward visits them.
std::vector<Xf*> all_xf(dg_unet* u) {
  std::vector<Xf*> v;
  for (auto& d : u->down) for (auto& x : d.xf) v.push_back(&x);
  v.push_back(&u->mid_xf);
  for (auto& b : u->up) for (auto& x : b.xf) v.push_back(&x);
  return v;
}

// Step-invariant work of the denoising loop, done once per dg_denoise_loop call (eager launches on `s`):
//  - cross-attention K/V = to_k / to_v of the text embedding for all 16 transformer blocks (32 of the 210 GEMMs of a forward);
//  - Timesteps -> TimestepEmbedding -> all 22 time_emb_proj(SiLU(emb)) rows for every step of the schedule.
// Exact: the same kernels on the same inputs, only not repeated 50 times.
int hoist_loop_invariants(dg_unet* u, cudaStream_t s, const __half* ehs, int tokens, int B, const float* d_ttab, const float* t_host,
                          int n_steps) {
  const dg_unet_config& cf = u->cfg;
  u->arena.reset();
  Fwd f{u, s, u->ctx->num_sms, B, tokens, ehs, nullptr};
  auto xfs = all_xf(u);
  for (size_t i = 0; i < xfs.size(); ++i) {
    const Xf& x = *xfs[i];
    if ((size_t)B * tokens * 2 * x.c > u->kv_cache_elems) return fail(DG_E_SHAPE, "context exceeds the prepared K/V cache");
    f.linear(ehs, cf.cross_attention_dim, nullptr, 0, B * tokens, x.kv2, nullptr, 0, u->kv_cache[i]);
  }
  if (f.err) return f.err;
  if (u->temb_cache && (int)u->temb_ts.size() == n_steps && memcmp(u->temb_ts.data(), t_host, sizeof(float) * n_steps) == 0)
    return DG_OK;             // same timesteps, same weights: the table of the previous call stands (stream order keeps it valid)
  u->temb_ts.clear();
  if (u->temb_table_rows < n_steps) {
    cudaFree(u->temb_table);
    u->temb_table = nullptr; u->temb_table_rows = 0;
    DG_CUDA(cudaMalloc((void**)&u->temb_table, (size_t)n_steps * u->temb_total * 2));
    u->temb_table_rows = n_steps;
  }
  const int c0 = cf.block_out_channels[0];
  __half* tsin = f.alloc((size_t)n_steps * c0 * 2);
  __half* t1 = f.alloc((size_t)n_steps * u->temb_dim * 2);
  __half* emb = f.alloc((size_t)n_steps * u->temb_dim * 2);
  if (f.err) return f.err;
  timestep_sinusoid_kernel<<<(n_steps * c0 / 2 + 127) / 128, 128, 0, s>>>(d_ttab, tsin, n_steps, c0, cf.freq_shift, cf.flip_sin_to_cos);
  DG_LAUNCH_CHECK();
  for (int b0 = 0; b0 < n_steps; b0 += 8) {
    const int nb = std::min(8, n_steps - b0);
    DG_TRY(launch_gemv(s, tsin + (size_t)b0 * c0, c0, u->time1.w, u->time1.b, t1 + (size_t)b0 * u->temb_dim, u->temb_dim, nb, u->temb_dim, c0, 0, 1));
    DG_TRY(launch_gemv(s, t1 + (size_t)b0 * u->temb_dim, u->temb_dim, u->time2.w, u->time2.b, emb + (size_t)b0 * u->temb_dim, u->temb_dim, nb, u->temb_dim, u->temb_dim, 0, 0));
    DG_TRY(launch_gemv(s, emb + (size_t)b0 * u->temb_dim, u->temb_dim, u->temb_proj_w, u->temb_proj_b, u->temb_table + (size_t)b0 * u->temb_total, u->temb_total, nb, u->temb_total, u->temb_dim, 1, 0));
  }
  u->temb_ts.assign(t_host, t_host + n_steps);
  return DG_OK;
}

int check_prepared(dg_unet* u, int batch, int h, int w, int tokens) {
  if (!u->arena.base) return fail(DG_E_STATE, "dg_unet_prepare has not been called");
  if (batch > u->max_batch || h * w > u->ws_h * u->ws_w || tokens > u->ws_tokens)
    return fail(DG_E_SHAPE, "forward (batch %d, %dx%d, %d tokens) exceeds prepared workspace (batch %d, %dx%d, %d tokens)",
                batch, h, w, tokens, u->max_batch, u->ws_h, u->ws_w, u->ws_tokens);
  if (h % 8 || w % 8) return fail(DG_E_SHAPE, "latent size %dx%d must be a multiple of 8", h, w);
  int missing = 0;
  for (auto& sl : u->slots) missing += sl.set ? 0 : 1;
  if (missing) return fail(DG_E_STATE, "%d weights have not been set", missing);
  return finalize_weights(u);
}

// ================================================================== VAE decoder
// AutoencoderKL.decode of StableDiffusionPipeline.__call__ step 8: post_quant_conv -> conv_in -> mid (resnet, single-head
// attention, resnet) -> 4 up blocks of 3 resnets (+ nearest-x2 upsample conv) -> GroupNorm + SiLU -> conv_out.  Same kernels as
// the UNet: tcgen05 implicit-GEMM convs / linears (gemm2_kernel), GroupNorm (+SiLU) kernels; the 512-wide single-head
// attention runs as two tcgen05 GEMMs (Q K^T, P V) around a row-softmax kernel.  Eager launches, no graph (3 % of the image).
struct VRes { Norm n1, n2; Lin c1, c2, sc; bool has_sc = false; int cin = 0, cout = 0; };
struct VUp { std::vector<VRes> res; bool has_up = false; Lin up; };
}  // namespace

struct dg_vae : WeightStore {
  int ch[4] = {128, 256, 512, 512};
  int layers = 2, groups = 32, latent_ch = 4, out_ch = 3;
  float eps = 1e-6f;
  __half* pq_w = nullptr; __half* pq_b = nullptr;   // post_quant_conv
  Lin conv_in, conv_out, aq, ak, av, ao;
  Norm attn_gn, norm_out;
  VRes mid0, mid1;
  std::vector<VUp> up;
  Arena arena;
  float* gn_stats = nullptr;
  int max_batch = 0, ws_h = 0, ws_w = 0;
};

namespace {

int make_vres(dg_vae* v, const std::string& pfx, int cin, int cout, VRes* r) {
  r->cin = cin; r->cout = cout;
  DG_TRY(make_norm(v, pfx + ".norm1", cin, &r->n1));
  DG_TRY(make_conv3(v, pfx + ".conv1", cin, cout, &r->c1));
  DG_TRY(make_norm(v, pfx + ".norm2", cout, &r->n2));
  DG_TRY(make_conv3(v, pfx + ".conv2", cout, cout, &r->c2));
  r->has_sc = cin != cout;
  if (r->has_sc) DG_TRY(make_linear(v, pfx + ".conv_shortcut", cin, cout, true, &r->sc, true));
  return DG_OK;
}

int build_vae(dg_vae* v) {
  const int* ch = v->ch;
  const int cm = ch[3];
  DG_TRY(dev_alloc(v, (void**)&v->pq_w, (size_t)v->latent_ch * v->latent_ch * 2));
  DG_TRY(dev_alloc(v, (void**)&v->pq_b, (size_t)v->latent_ch * 2));
  add_slot(v, "post_quant_conv.weight", {v->latent_ch, v->latent_ch, 1, 1}, PK_COPY, v->pq_w);
  add_slot(v, "post_quant_conv.bias", {v->latent_ch}, PK_COPY, v->pq_b);
  v->conv_in.in = 64; v->conv_in.out = cm; v->conv_in.rows = cm;
  DG_TRY(dev_alloc(v, (void**)&v->conv_in.w, (size_t)64 * cm * 2));
  DG_TRY(dev_alloc(v, (void**)&v->conv_in.b, cm * 2));
  add_slot(v, "decoder.conv_in.weight", {cm, v->latent_ch, 3, 3}, PK_CONV_IN, v->conv_in.w, 0, cm, v->latent_ch);
  add_slot(v, "decoder.conv_in.bias", {cm}, PK_COPY, v->conv_in.b);
  DG_TRY(make_vres(v, "decoder.mid_block.resnets.0", cm, cm, &v->mid0));
  DG_TRY(make_norm(v, "decoder.mid_block.attentions.0.group_norm", cm, &v->attn_gn));
  DG_TRY(make_linear(v, "decoder.mid_block.attentions.0.to_q", cm, cm, true, &v->aq));
  DG_TRY(make_linear(v, "decoder.mid_block.attentions.0.to_k", cm, cm, true, &v->ak));
  DG_TRY(make_linear(v, "decoder.mid_block.attentions.0.to_v", cm, cm, true, &v->av));
  DG_TRY(make_linear(v, "decoder.mid_block.attentions.0.to_out.0", cm, cm, true, &v->ao));
  DG_TRY(make_vres(v, "decoder.mid_block.resnets.1", cm, cm, &v->mid1));
  v->up.resize(4);
  int cin = cm;
  for (int i = 0; i < 4; ++i) {
    const int cout = ch[3 - i];
    VUp& u = v->up[i];
    u.res.resize(v->layers + 1);
    const std::string pfx = "decoder.up_blocks." + std::to_string(i);
    for (int j = 0; j <= v->layers; ++j) DG_TRY(make_vres(v, pfx + ".resnets." + std::to_string(j), j == 0 ? cin : cout, cout, &u.res[j]));
    u.has_up = i != 3;
    if (u.has_up) DG_TRY(make_conv3(v, pfx + ".upsamplers.0.conv", cout, cout, &u.up, true));
    cin = cout;
  }
  DG_TRY(make_norm(v, "decoder.conv_norm_out", ch[0], &v->norm_out));
  DG_TRY(make_conv3(v, "decoder.conv_out", ch[0], v->out_ch, &v->conv_out));
  return DG_OK;
}

struct VFwd {
  dg_vae* v; cudaStream_t s; int err = DG_OK;
  __half* alloc(size_t bytes) {
    void* p = v->arena.alloc(bytes);
    if (!p && err == DG_OK) err = fail(DG_E_NOMEM, "VAE arena exhausted (%zu bytes requested); call dg_vae_prepare with a larger batch", bytes);
    return (__half*)p;
  }
  T4 talloc(int B, int H, int W, int C) { T4 t; t.B = B; t.H = H; t.W = W; t.C = C; t.p = alloc(t.bytes()); return t; }
  void free_(T4& t) { v->arena.release(t.p); t.p = nullptr; }
  void gn(const T4& x, const Norm& n, int silu, T4& out) {
    FW(launch_groupnorm(s, v->ctx->num_sms, x.p, x.C, nullptr, 0, n.g, n.b, out.p, v->gn_stats, x.B, x.H * x.W, v->groups, v->eps, silu));
  }
  void conv3(const T4& x, const Lin& w, const __half* residual, T4& out, int ldo = 0) {
    GemmArgs a; a.a0 = x.p; a.c0 = x.C; a.B = x.B; a.H = x.H; a.W = x.W; a.taps = 9; a.w = w.w; a.n_w = w.rows; a.n_out = w.out;
    a.bias = w.b; a.residual = residual; a.ld_res = w.out; a.out = out.p; a.ldo = ldo ? ldo : w.out;
    FW(launch_gemm(s, v->ctx->gemm, a));
  }
  void linear(const __half* x, int K, int rows, const __half* w, int n_w, const __half* bias, const __half* residual, __half* out, int ldo) {
    GemmArgs a; a.a0 = x; a.c0 = K; a.B = 1; a.H = 1; a.W = rows; a.taps = 1; a.w = w; a.n_w = n_w; a.n_out = n_w; a.bias = bias;
    a.residual = residual; a.ld_res = ldo; a.out = out; a.ldo = ldo;
    FW(launch_gemm(s, v->ctx->gemm, a));
  }
  T4 resnet(const VRes& r, T4& x) {       // consumes x
    T4 hn = talloc(x.B, x.H, x.W, r.cin);
    gn(x, r.n1, 1, hn);
    T4 h1 = talloc(x.B, x.H, x.W, r.cout);
    conv3(hn, r.c1, nullptr, h1);
    free_(hn);
    T4 h2n = talloc(x.B, x.H, x.W, r.cout);
    gn(h1, r.n2, 1, h2n);
    free_(h1);
    T4 out = talloc(x.B, x.H, x.W, r.cout);
    const __half* resid = x.p;
    T4 sc{};
    if (r.has_sc) {
      sc = talloc(x.B, x.H, x.W, r.cout);
      linear(x.p, x.C, x.B * x.H * x.W, r.sc.w, r.sc.rows, r.sc.b, nullptr, sc.p, r.cout);
      resid = sc.p;
    }
    conv3(h2n, r.c2, resid, out);
    free_(h2n);
    if (r.has_sc) free_(sc);
    free_(x);
    return out;
  }
  T4 attention(T4& x) {                    // consumes x
    const int C = x.C, S = x.H * x.W, rows = x.B * S;
    T4 xn = talloc(x.B, x.H, x.W, C);
    gn(x, v->attn_gn, 0, xn);
    __half* q = alloc((size_t)rows * C * 2);
    __half* k = alloc((size_t)rows * C * 2);
    __half* vv = alloc((size_t)rows * C * 2);
    linear(xn.p, C, rows, v->aq.w, C, v->aq.b, nullptr, q, C);
    linear(xn.p, C, rows, v->ak.w, C, v->ak.b, nullptr, k, C);
    linear(xn.p, C, rows, v->av.w, C, v->av.b, nullptr, vv, C);
    // S = h*w is a multiple of 64 (latent sides are multiples of 8): K of the P*V GEMM is whole 64-wide k-blocks
    __half* sc = alloc((size_t)S * S * 2);
    __half* vt = alloc((size_t)C * S * 2);
    if (err) return x;
    const float scale_log2 = (1.0f / sqrtf((float)C)) * 1.4426950408889634f;
    for (int b = 0; b < x.B && !err; ++b) {
      // S = Q_b K_b^T   (K_b as the K-major "weight": [S keys, C])
      linear(q + (size_t)b * S * C, C, S, k + (size_t)b * S * C, S, nullptr, nullptr, sc, S);
      softmax_rows_kernel<<<S, 256, 0, s>>>(sc, S, scale_log2);
      ++g_launch_counter;
      dim3 tb(32, 8), tg((C + 31) / 32, (S + 31) / 32);
      transpose_rows_kernel<<<tg, tb, 0, s>>>(vv + (size_t)b * S * C, vt, S, C);
      ++g_launch_counter;
      // O_b = P V_b   (V_b^T as the K-major weight: [C, S])
      linear(sc, S, S, vt, C, nullptr, nullptr, xn.p + (size_t)b * S * C, C);
    }
    v->arena.release(sc); v->arena.release(vt); v->arena.release(q); v->arena.release(k); v->arena.release(vv);
    T4 out = talloc(x.B, x.H, x.W, C);
    linear(xn.p, C, rows, v->ao.w, C, v->ao.b, x.p, out.p, C);
    free_(xn); free_(x);
    return out;
  }
};

int vae_missing(dg_vae* v) { int m = 0; for (auto& s : v->slots) m += s.set ? 0 : 1; return m; }

}  // namespace

// ================================================================== C ABI
// ====================================================================================================================
// CLIP text encoder: transformers `CLIPTextModel.forward(input_ids).last_hidden_state`, the
// `prompt_embeds` the reference obtains through `stage_1.encode_prompt(prompt)` (txt2img_diffusers_stages_from_txt.py:242).
// Pre-LN transformer, causal self-attention, quick_gelu MLP, final LayerNorm; weights by transformers state-dict key.
struct ClipLayer { Norm ln1, ln2; Lin qkv, out, fc1, fc2; };
struct dg_clip : WeightStore {
  int vocab = 49408, hidden = 768, inter = 3072, layers = 12, heads = 12, max_pos = 77;
  float eps = 1e-5f;
  int act = 0;               // 0 quick_gelu, 1 erf gelu (dg_clip_set_activation)
  __half* tok = nullptr; __half* pos = nullptr;
  std::vector<ClipLayer> L;
  Norm final_ln;
  __half* ws = nullptr; int* ids_dev = nullptr;
  int max_batch = 0;
};

namespace {

int build_clip_layers(WeightStore* c, const std::string& tower, int C, int inter, int layers, std::vector<ClipLayer>* L) {
  L->resize(layers);
  for (int i = 0; i < layers; ++i) {
    ClipLayer& l = (*L)[i];
    const std::string pfx = tower + ".encoder.layers." + std::to_string(i);
    DG_TRY(make_norm(c, pfx + ".layer_norm1", C, &l.ln1));
    DG_TRY(make_fused_rows(c, {pfx + ".self_attn.q_proj.weight", pfx + ".self_attn.k_proj.weight", pfx + ".self_attn.v_proj.weight"}, C, C, &l.qkv));
    DG_TRY(dev_alloc(c, (void**)&l.qkv.b, (size_t)3 * C * 2));
    add_slot(c, pfx + ".self_attn.q_proj.bias", {C}, PK_ROWS, l.qkv.b, 0, C, 1);
    add_slot(c, pfx + ".self_attn.k_proj.bias", {C}, PK_ROWS, l.qkv.b, C, C, 1);
    add_slot(c, pfx + ".self_attn.v_proj.bias", {C}, PK_ROWS, l.qkv.b, 2 * C, C, 1);
    DG_TRY(make_linear(c, pfx + ".self_attn.out_proj", C, C, true, &l.out));
    DG_TRY(make_norm(c, pfx + ".layer_norm2", C, &l.ln2));
    DG_TRY(make_linear(c, pfx + ".mlp.fc1", C, inter, true, &l.fc1));
    DG_TRY(make_linear(c, pfx + ".mlp.fc2", inter, C, true, &l.fc2));
  }
  return DG_OK;
}

int build_clip(dg_clip* c) {
  const int C = c->hidden;
  DG_TRY(dev_alloc(c, (void**)&c->tok, (size_t)c->vocab * C * 2));
  DG_TRY(dev_alloc(c, (void**)&c->pos, (size_t)c->max_pos * C * 2));
  add_slot(c, "text_model.embeddings.token_embedding.weight", {c->vocab, C}, PK_COPY, c->tok);
  add_slot(c, "text_model.embeddings.position_embedding.weight", {c->max_pos, C}, PK_COPY, c->pos);
  DG_TRY(build_clip_layers(c, "text_model", C, c->inter, c->layers, &c->L));
  DG_TRY(make_norm(c, "text_model.final_layer_norm", C, &c->final_ln));
  return DG_OK;
}

int clip_linear(dg_ctx* ctx, cudaStream_t s, const __half* x, int K, int rows, const Lin& w, const __half* residual, __half* out) {
  GemmArgs a; a.a0 = x; a.c0 = K; a.B = 1; a.H = 1; a.W = rows; a.taps = 1; a.w = w.w; a.n_w = w.rows; a.n_out = w.out; a.bias = w.b;
  a.residual = residual; a.ld_res = w.out; a.out = out; a.ldo = w.out;
  return launch_gemm(s, ctx->gemm, a);
}

// Pre-LN transformer layers (CLIPEncoderLayer x N) over x [batch * seq, C]; buffers: x2, h, att [rows, C], qkv [rows, 3C],
// f [rows, inter].  The result is back in x.
int run_clip_layers(dg_ctx* ctx, cudaStream_t s, const std::vector<ClipLayer>& L, __half* x, __half* x2, __half* h, __half* att,
                    __half* qkv, __half* f, int batch, int seq, int C, int inter, int heads, int causal, float eps, int act = 0) {
  const int rows = batch * seq;
  for (const ClipLayer& l : L) {
    DG_TRY(launch_layernorm(s, x, l.ln1.g, l.ln1.b, h, rows, C, eps));
    DG_TRY(clip_linear(ctx, s, h, C, rows, l.qkv, nullptr, qkv));
    DG_TRY(launch_attention(s, qkv, 3 * C, qkv + C, 3 * C, qkv + 2 * C, 3 * C, att, batch, heads, seq, seq, 64, causal));
    DG_TRY(clip_linear(ctx, s, att, C, rows, l.out, x, x2));
    DG_TRY(launch_layernorm(s, x2, l.ln2.g, l.ln2.b, h, rows, C, eps));
    DG_TRY(clip_linear(ctx, s, h, C, rows, l.fc1, nullptr, f));
    quick_gelu_kernel<<<grid_for((size_t)rows * inter / 8, 256, ctx->num_sms), 256, 0, s>>>(f, (size_t)rows * inter / 8, act);
    DG_LAUNCH_CHECK();
    DG_TRY(clip_linear(ctx, s, f, inter, rows, l.fc2, x2, x));
  }
  return DG_OK;
}

}  // namespace

// CLIP similarity scorer: `clip_model(images, text)` -> logits_per_text of OpenAI CLIP ViT-L/14 as used
// by DiverGen/filteration/get_clip_score.py:176-180 (weights by transformers `CLIPModel` state-dict key).
struct dg_clipscore : WeightStore {
  int vocab = 49408, t_hidden = 768, t_inter = 3072, t_layers = 12, t_heads = 12, max_pos = 77;
  int image = 224, patch = 14, v_hidden = 1024, v_inter = 4096, v_layers = 24, v_heads = 16, proj = 768;
  int kpad = 640;
  float eps = 1e-5f, logit_scale = 4.6052f;
  __half* tok = nullptr; __half* tpos = nullptr; std::vector<ClipLayer> TL; Norm t_final; Lin t_proj;
  __half* patch_w = nullptr; __half* cls = nullptr; __half* vpos = nullptr; Norm pre_ln, post_ln; std::vector<ClipLayer> VL; Lin v_proj;
  __half* ws = nullptr; int* ids_dev = nullptr; int* eos_dev = nullptr;
  int max_images = 0, max_texts = 0;
  size_t ws_elems = 0;
};

namespace {

int build_clipscore(dg_clipscore* c) {
  const int Ct = c->t_hidden, Cv = c->v_hidden, n = c->image / c->patch, S = n * n + 1, kreal = 3 * c->patch * c->patch;
  c->kpad = (kreal + 63) / 64 * 64;
  DG_TRY(dev_alloc(c, (void**)&c->tok, (size_t)c->vocab * Ct * 2));
  DG_TRY(dev_alloc(c, (void**)&c->tpos, (size_t)c->max_pos * Ct * 2));
  add_slot(c, "text_model.embeddings.token_embedding.weight", {c->vocab, Ct}, PK_COPY, c->tok);
  add_slot(c, "text_model.embeddings.position_embedding.weight", {c->max_pos, Ct}, PK_COPY, c->tpos);
  DG_TRY(build_clip_layers(c, "text_model", Ct, c->t_inter, c->t_layers, &c->TL));
  DG_TRY(make_norm(c, "text_model.final_layer_norm", Ct, &c->t_final));
  DG_TRY(make_linear(c, "text_projection", Ct, c->proj, false, &c->t_proj));
  DG_TRY(dev_alloc(c, (void**)&c->patch_w, (size_t)Cv * c->kpad * 2));
  DG_TRY(dev_alloc(c, (void**)&c->cls, (size_t)Cv * 2));
  DG_TRY(dev_alloc(c, (void**)&c->vpos, (size_t)S * Cv * 2));
  add_slot(c, "vision_model.embeddings.class_embedding", {Cv}, PK_COPY, c->cls);
  add_slot(c, "vision_model.embeddings.patch_embedding.weight", {Cv, 3, c->patch, c->patch}, PK_PAD_ROWS, c->patch_w, c->kpad, Cv, kreal);
  add_slot(c, "vision_model.embeddings.position_embedding.weight", {S, Cv}, PK_COPY, c->vpos);
  DG_TRY(make_norm(c, "vision_model.pre_layrnorm", Cv, &c->pre_ln));      // (sic: the upstream key is misspelt)
  DG_TRY(build_clip_layers(c, "vision_model", Cv, c->v_inter, c->v_layers, &c->VL));
  DG_TRY(make_norm(c, "vision_model.post_layernorm", Cv, &c->post_ln));
  DG_TRY(make_linear(c, "visual_projection", Cv, c->proj, false, &c->v_proj));
  return DG_OK;
}

}  // namespace

extern "C" {

int32_t dg_version(void) { return 100; }
int32_t dg_plan_attention_grid(int32_t n_bh, int32_t q_tiles, int32_t tiles_per_cta, int32_t sms, int32_t* plan, double* makespan) {
  if (!plan || !makespan || n_bh <= 0 || q_tiles <= 0 || tiles_per_cta <= 0 || sms <= 0) return fail(DG_E_ARG, "bad argument");
  const AttnPlan pl = plan_attn_grid(n_bh, q_tiles, tiles_per_cta, sms);
  plan[0] = pl.on; plan[1] = pl.on ? pl.n_g1 : 0; plan[2] = pl.on ? pl.a1 : 0; plan[3] = pl.on ? pl.b1 : 0;
  plan[4] = pl.on ? pl.a2 : 0; plan[5] = pl.on ? pl.b2 : 0; plan[6] = pl.on ? pl.ctas : 0;
  makespan[0] = pl.makespan; makespan[1] = pl.uniform;
  return DG_OK;
}
const char* dg_last_error(void) { return g_last_error.c_str(); }

int32_t dg_ctx_create(int32_t device, dg_ctx** out) {
  if (!out) return fail(DG_E_ARG, "out is null");
  int n = 0;
  DG_CUDA(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) return fail(DG_E_ARG, "device %d out of range (%d devices)", device, n);
  cudaDeviceProp prop;
  DG_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(DG_E_ARCH, "device %d is sm_%d%d; this library contains sm_100a code only (no fallback)", device, prop.major, prop.minor);
  DG_CUDA(cudaSetDevice(device));
  if (!get_encode_fn()) return fail(DG_E_CUDA, "driver entry point cuTensorMapEncodeTiled not found");
  DG_TRY(init_kernel_attributes());
  dg_ctx* c = new dg_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  DG_CUDA(cudaMalloc(&c->gn_stats, sizeof(float) * 2 * 64 * 64));
  c->gemm.num_sms = c->num_sms;
  DG_TRY(query_max_pairs(&c->gemm.max_pairs));
  DG_CUDA(cudaMalloc(&c->gemm.ws, kSplitWsFloats * sizeof(float)));
  DG_CUDA(cudaMemset(c->gemm.ws, 0, kSplitWsFloats * sizeof(float)));
  DG_CUDA(cudaMalloc(&c->gemm.tickets, kSplitTickets * sizeof(int)));
  DG_CUDA(cudaMemset(c->gemm.tickets, 0, kSplitTickets * sizeof(int)));
  auto env_int = [](const char* name, int dflt) { const char* e = getenv(name); return e && e[0] ? atoi(e) : dflt; };
  c->gemm.cta_mode = env_int("DG_GEMM_CTA", 2) == 1 ? 1 : 2;
  if (env_int("DG_SPLITK", 1) == 0) { cudaFree(c->gemm.ws); c->gemm.ws = nullptr; }
  c->fuse_ln = env_int("DG_FUSE_LN", 1);
  c->fuse_gn = env_int("DG_FUSE_GN", 1);
  // off by default: measured SLOWER (13.2 vs 9.7 ms per forward, profiles/r02_ab.md) -- the activation is re-evaluated for
  // each of the 9 taps and one MUFU.TANH per element (512 cycles per k-block per SM) does not fit under the MMA time
  c->fuse_xf = env_int("DG_FUSE_XF", 0);
  c->up_phases = env_int("DG_UPCONV_PHASES", 1);
  c->fuse_sc = env_int("DG_FUSE_SC", 1);
  c->cfg_dedup = env_int("DG_CFG_DEDUP", 1);
  *out = c;
  return DG_OK;
}
void dg_ctx_destroy(dg_ctx* ctx) {
  if (!ctx) return;
  cudaFree(ctx->gn_stats); cudaFree(ctx->gemm.ws); cudaFree(ctx->gemm.tickets); cudaFree(ctx->xf_tab);
  delete ctx;
}

int32_t dg_unet_create(dg_ctx* ctx, const dg_unet_config* cfg, dg_unet** out) {
  if (!ctx || !cfg || !out) return fail(DG_E_ARG, "null argument");
  for (int i = 0; i < 4; ++i) {
    if (cfg->block_out_channels[i] % 64) return fail(DG_E_UNSUPPORTED, "block_out_channels[%d]=%d must be a multiple of 64", i, cfg->block_out_channels[i]);
    if (cfg->num_heads[i] <= 0 || cfg->block_out_channels[i] % cfg->num_heads[i]) return fail(DG_E_SHAPE, "heads[%d]=%d", i, cfg->num_heads[i]);
  }
  if (cfg->cross_attention_dim % 64) return fail(DG_E_UNSUPPORTED, "cross_attention_dim must be a multiple of 64");
  if (cfg->norm_num_groups > 64) return fail(DG_E_UNSUPPORTED, "norm_num_groups > 64");
  DG_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<dg_unet> u(new dg_unet());
  u->ctx = ctx; u->cfg = *cfg;
  { const char* e = getenv("DG_TEMB_CACHE"); u->temb_cache = !(e && e[0] == '0'); }
  {
    // fused GroupNorm sums are kept per (sample, blk-channel block); every channel count of the UNet is a multiple of
    // block_out_channels[0], so blk = block_out_channels[0] / groups tiles every (concatenated) group exactly.
    const int c0 = cfg->block_out_channels[0], g = cfg->norm_num_groups;
    bool ok = ctx->fuse_gn && g > 0 && c0 % g == 0;
    const int blk = ok ? c0 / g : 0;
    ok = ok && blk % 2 == 0;
    for (int i = 0; ok && i < 4; ++i) ok = cfg->block_out_channels[i] % c0 == 0;
    u->gn_blk = ok ? blk : 0;
  }
  int r = build_modules(u.get());
  if (r != DG_OK) { for (void* p : u->owned) cudaFree(p); return r; }
  *out = u.release();
  return DG_OK;
}
void dg_unet_destroy(dg_unet* u) {
  if (!u) return;
  for (auto& g : u->graphs) cudaGraphExecDestroy(g.exec);
  if (u->cap_stream) cudaStreamDestroy(u->cap_stream);
  for (void* p : u->owned) cudaFree(p);
  cudaFree(u->arena.base); cudaFree(u->d_t); cudaFree(u->gn_stats); cudaFree(u->d_coef); cudaFree(u->d_step);
  cudaFree(u->loop_in); cudaFree(u->loop_out); cudaFree(u->gn_arena); cudaFree(u->ln_arena); cudaFree(u->xf_arena);
  cudaFree(u->temb_cur); cudaFree(u->temb_table);
  for (__half* p : u->kv_cache) cudaFree(p);
  delete u;
}
int32_t dg_unet_num_weights(dg_unet* u) { return u ? (int32_t)u->slots.size() : 0; }
const char* dg_unet_weight_name(dg_unet* u, int32_t i) {
  if (!u || i < 0 || i >= (int)u->slots.size()) return nullptr;
  return u->slots[i].key.c_str();
}
int32_t dg_unet_weight_shape(dg_unet* u, int32_t i, int64_t* shape4, int32_t* ndim) {
  if (!u || i < 0 || i >= (int)u->slots.size() || !shape4 || !ndim) return fail(DG_E_ARG, "bad argument");
  *ndim = (int32_t)u->slots[i].shape.size();
  for (int k = 0; k < *ndim; ++k) shape4[k] = u->slots[i].shape[k];
  return DG_OK;
}
int32_t dg_unet_missing_weights(dg_unet* u) {
  int m = 0;
  if (u) for (auto& s : u->slots) m += s.set ? 0 : 1;
  return m;
}

static int store_set_weight(WeightStore* u, const char* key, const void* src, int32_t ndim, const int64_t* shape) {
  if (!u || !key || !src || !shape) return fail(DG_E_ARG, "null argument");
  auto it = u->slot_index.find(key);
  if (it == u->slot_index.end()) return fail(DG_E_ARG, "unexpected state-dict key '%s'", key);
  Slot& s = u->slots[it->second];
  bool ok = (int)s.shape.size() == ndim;
  for (int i = 0; ok && i < ndim; ++i) ok = s.shape[i] == shape[i];
  if (!ok) return fail(DG_E_SHAPE, "shape mismatch for '%s'", key);
  DG_CUDA(cudaSetDevice(u->ctx->device));
  const __half* w = (const __half*)src;
  size_t n = 1;
  for (auto d : s.shape) n *= (size_t)d;
  const int sms = u->ctx->num_sms;
  switch (s.kind) {
    case PK_COPY: DG_CUDA(cudaMemcpy(s.dst, w, n * 2, cudaMemcpyDeviceToDevice)); break;
    case PK_ROWS: DG_CUDA(cudaMemcpy(s.dst + (size_t)s.row_off * s.b, w, n * 2, cudaMemcpyDeviceToDevice)); break;
    case PK_PAD_ROWS:      // a rows of b elements -> rows of row_off elements, zero padded (patch embedding: K 588 -> 640)
      DG_CUDA(cudaMemset(s.dst, 0, (size_t)s.a * s.row_off * 2));
      DG_CUDA(cudaMemcpy2D(s.dst, (size_t)s.row_off * 2, w, (size_t)s.b * 2, (size_t)s.b * 2, s.a, cudaMemcpyDeviceToDevice));
      break;
    case PK_CONV3:
      pack_conv3x3_kernel<<<grid_for(n, 256, sms), 256>>>(w, s.dst, s.a, s.b, s.b);
      DG_LAUNCH_CHECK();
      if (s.dst2) {
        pack_upconv_kernel<<<grid_for((size_t)16 * s.a * s.b, 256, sms), 256>>>(w, s.dst2, s.a, s.b);
        DG_LAUNCH_CHECK();
      }
      break;
    case PK_CONV_IN:
      pack_conv_in_kernel<<<grid_for((size_t)s.a * 64, 256, sms), 256>>>(w, s.dst, s.a, s.b, 64);
      DG_LAUNCH_CHECK();
      break;
    case PK_GEGLU_W:
    case PK_GEGLU_B: {
      const int tiles = geglu_rows(s.a) / kGegluTile;
      pack_geglu_kernel<<<grid_for((size_t)tiles * kGegluTile * s.b, 256, sms), 256>>>(w, s.dst, s.a, s.b, kGegluTile, kGegluHalf, tiles);
      DG_LAUNCH_CHECK();
      break;
    }
  }
  DG_CUDA(cudaDeviceSynchronize());
  s.set = true;
  return DG_OK;
}
int32_t dg_unet_set_weight(dg_unet* u, const char* key, const void* src, int32_t ndim, const int64_t* shape) {
  DG_TRY(store_set_weight(u, key, src, ndim, shape));
  u->finalized = false;
  return DG_OK;
}

int32_t dg_unet_prepare(dg_unet* u, int32_t max_batch, int32_t h, int32_t w, int32_t ctx_tokens) {
  if (!u || max_batch <= 0 || h <= 0 || w <= 0 || ctx_tokens <= 0) return fail(DG_E_ARG, "bad argument");
  DG_CUDA(cudaSetDevice(u->ctx->device));
  for (auto& g : u->graphs) cudaGraphExecDestroy(g.exec);
  u->graphs.clear();
  cudaFree(u->arena.base); cudaFree(u->d_t); cudaFree(u->gn_stats); cudaFree(u->loop_in); cudaFree(u->loop_out);
  u->arena.base = nullptr; u->d_t = nullptr; u->gn_stats = nullptr; u->loop_in = nullptr; u->loop_out = nullptr;
  // Peak live activation set: generous closed-form bound (skips + the widest transformer intermediates).
  const size_t c0 = u->cfg.block_out_channels[0];
  const size_t pix = (size_t)max_batch * h * w;
  size_t bytes = pix * c0 * 2 * 40 + ((size_t)64 << 20);
  u->arena.size = bytes;
  cudaError_t e = cudaMalloc((void**)&u->arena.base, bytes);
  if (e != cudaSuccess) return fail(DG_E_NOMEM, "workspace cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  DG_CUDA(cudaMalloc((void**)&u->d_t, sizeof(float) * max_batch));
  DG_CUDA(cudaMalloc((void**)&u->gn_stats, sizeof(float) * 2 * u->cfg.norm_num_groups * max_batch));
  const size_t lat = (size_t)max_batch * u->cfg.in_channels * h * w * 2;
  DG_CUDA(cudaMalloc((void**)&u->loop_in, lat));
  DG_CUDA(cudaMalloc((void**)&u->loop_out, lat));
  if (!u->d_step) DG_CUDA(cudaMalloc((void**)&u->d_step, sizeof(int)));
  cudaFree(u->gn_arena); cudaFree(u->ln_arena); cudaFree(u->temb_cur);
  u->gn_arena = nullptr; u->ln_arena = nullptr; u->temb_cur = nullptr;
  for (__half* p : u->kv_cache) cudaFree(p);
  u->kv_cache.clear();
  {
    size_t cmax = 0;
    for (int i = 0; i < 4; ++i) cmax = std::max(cmax, (size_t)u->cfg.block_out_channels[i]);
    u->kv_cache_elems = (size_t)max_batch * ctx_tokens * 2 * cmax;
    const size_t nxf = all_xf(u).size();
    for (size_t i = 0; i < nxf; ++i) {
      __half* p = nullptr;
      DG_CUDA(cudaMalloc((void**)&p, u->kv_cache_elems * 2));
      u->kv_cache.push_back(p);
    }
    DG_CUDA(cudaMalloc((void**)&u->temb_cur, (size_t)u->temb_total * 2));
  }
  {
    // ~70 GroupNorm inputs of at most 4*c0 channels; 48 LayerNorm inputs of at most pix rows x 8 partials
    const int blk = u->gn_blk > 0 ? u->gn_blk : 2;
    u->gn_cap = (size_t)96 * max_batch * ((size_t)h * w / 32 + 1) * (c0 / blk + 1) * 2;   // level 0 dominates: slabs shrink 4x per level, channels grow <= 2x
    u->ln_cap = (size_t)pix * 64 * (2 + c0 / 80) + ((size_t)1 << 20);   // >= 3x the sum over blocks of rows*parts*2
    DG_CUDA(cudaMalloc((void**)&u->gn_arena, u->gn_cap * sizeof(float)));
    DG_CUDA(cudaMalloc((void**)&u->ln_arena, u->ln_cap * sizeof(float)));
    // one (mean_h, scale', shift') table per GroupNorm applied inside a conv: <= 48 of them, each B x 3 x (<= 8 c0) halves
    cudaFree(u->xf_arena); u->xf_arena = nullptr;
    u->xf_cap = (size_t)64 * max_batch * 3 * 8 * c0;
    DG_CUDA(cudaMalloc((void**)&u->xf_arena, u->xf_cap * sizeof(__half)));
  }
  u->max_batch = max_batch; u->ws_h = h; u->ws_w = w; u->ws_tokens = ctx_tokens;
  return DG_OK;
}

int32_t dg_unet_set_graphs(dg_unet* u, int32_t enabled) {
  if (!u) return fail(DG_E_ARG, "null");
  u->use_graphs = enabled != 0;
  return DG_OK;
}
int64_t dg_unet_last_launch_count(dg_unet* u) { return u ? u->last_launches : 0; }
int32_t dg_unet_set_family_mask(dg_unet* u, int32_t mask) {
  if (!u) return fail(DG_E_ARG, "null");
  u->fam_mask = mask & 15;
  return DG_OK;
}

int32_t dg_unet_forward(dg_unet* u, const void* sample, const float* t_host, int32_t n_t, const void* ehs,
                        int32_t tokens, void* out, int32_t batch, int32_t h, int32_t w, void* stream) {
  if (!u || !sample || !t_host || !ehs || !out) return fail(DG_E_ARG, "null argument");
  if (n_t != 1 && n_t != batch) return fail(DG_E_SHAPE, "n_timesteps must be 1 or batch");
  if (n_t > 8) {
    bool same = true;
    for (int i = 1; i < n_t; ++i) same = same && t_host[i] == t_host[0];
    if (!same) return fail(DG_E_UNSUPPORTED, "more than 8 distinct per-sample timesteps");
    n_t = 1;
  }
  DG_TRY(check_prepared(u, batch, h, w, tokens));
  DG_CUDA(cudaSetDevice(u->ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  u->last_launches = 0;
  float t[8] = {0};
  for (int i = 0; i < n_t; ++i) t[i] = t_host[i];
  set_timesteps_kernel<<<1, 32, 0, s>>>(u->d_t, batch, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], n_t == 1);
  DG_LAUNCH_CHECK();
  u->last_launches += 1;
  return forward_maybe_graph(u, s, (const __half*)sample, (const __half*)ehs, tokens, (__half*)out, batch, h, w);
}

int32_t dg_unet_profile_forward(dg_unet* u, const void* sample, const float* t_host, int32_t n_t, const void* ehs,
                                int32_t tokens, void* out, int32_t batch, int32_t h, int32_t w, void* stream,
                                double* ms_by_family, double* flops_by_family, double* bytes_by_family,
                                int64_t* launches_by_family, double* total_ms) {
  if (!u || !ms_by_family || !flops_by_family || !bytes_by_family || !launches_by_family || !total_ms) return fail(DG_E_ARG, "null argument");
  const bool graphs = u->use_graphs;
  u->use_graphs = false;
  Profiler prof;
  cudaStream_t s = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, s);
  g_prof = &prof;
  int r = dg_unet_forward(u, sample, t_host, n_t, ehs, tokens, out, batch, h, w, stream);
  g_prof = nullptr;
  u->use_graphs = graphs;
  cudaEventRecord(e1, s);
  cudaError_t e = cudaStreamSynchronize(s);
  for (int i = 0; i < FAM_COUNT; ++i) { ms_by_family[i] = 0; flops_by_family[i] = 0; bytes_by_family[i] = 0; launches_by_family[i] = 0; }
  float ms = 0.f;
  if (e == cudaSuccess) { cudaEventElapsedTime(&ms, e0, e1); *total_ms = ms; }
  for (auto& rec : prof.recs) {
    if (e == cudaSuccess && rec.a && rec.b) {
      cudaEventElapsedTime(&ms, rec.a, rec.b);
      ms_by_family[rec.fam] += ms; flops_by_family[rec.fam] += rec.flops; bytes_by_family[rec.fam] += rec.bytes;
      launches_by_family[rec.fam] += 1;
    }
    cudaEventDestroy(rec.a); cudaEventDestroy(rec.b);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (r != DG_OK) return r;
  if (e != cudaSuccess) return fail(DG_E_CUDA, "profile forward failed: %s", cudaGetErrorString(e));
  return DG_OK;
}

int32_t dg_cfg_ddim_step(dg_ctx* ctx, const void* noise, void* latents, int32_t n_images, int64_t elems, float a_t,
                         float a_prev, float guidance, int32_t pred_type, void* stream) {
  if (!ctx || !noise || !latents) return fail(DG_E_ARG, "null argument");
  const size_t n = (size_t)n_images * elems;
  if (n % 8) return fail(DG_E_SHAPE, "element count must be a multiple of 8");
  if (pred_type != 0 && pred_type != 1) return fail(DG_E_ARG, "prediction_type must be 0 (epsilon) or 1 (v_prediction)");
  cudaStream_t s = (cudaStream_t)stream;
  // coefficient passed through a small device table owned by the context scratch (first 4 floats of gn_stats)
  const float4 cf = make_float4(sqrtf(a_t), sqrtf(1.f - a_t), sqrtf(a_prev), sqrtf(1.f - a_prev));
  DG_CUDA(cudaMemcpyAsync(ctx->gn_stats, &cf, sizeof(cf), cudaMemcpyHostToDevice, s));
  cfg_ddim_step_kernel<<<grid_for(n / 8, 256, ctx->num_sms), 256, 0, s>>>((const __half*)noise, (__half*)latents, n / 8, n,
                                                                         guidance, guidance > 1.0f, pred_type,
                                                                         (const float4*)ctx->gn_stats, nullptr);
  DG_LAUNCH_CHECK();
  return DG_OK;
}

int32_t dg_denoise_loop(dg_unet* u, void* latents, const void* ehs, int32_t tokens, int32_t n_images, int32_t h, int32_t w,
                        const float* t_host, const float* a_t, const float* a_prev, int32_t n_steps, float guidance,
                        int32_t pred_type, void* stream) {
  if (!u || !latents || !ehs || !t_host || !a_t || !a_prev || n_steps <= 0) return fail(DG_E_ARG, "bad argument");
  const bool cfg_on = guidance > 1.0f;
  const int B = cfg_on ? 2 * n_images : n_images;
  DG_TRY(check_prepared(u, B, h, w, tokens));
  DG_CUDA(cudaSetDevice(u->ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t elems = (size_t)n_images * u->cfg.in_channels * h * w;
  if (elems % 8) return fail(DG_E_SHAPE, "latent element count must be a multiple of 8");
  // per-step tables (device): coefficients + timestep values
  if (u->coef_cap < n_steps) {
    cudaFree(u->d_coef);
    DG_CUDA(cudaMalloc((void**)&u->d_coef, (sizeof(float4) + sizeof(float)) * n_steps));
    u->coef_cap = n_steps;
  }
  std::vector<float4> cf(n_steps);
  for (int i = 0; i < n_steps; ++i) cf[i] = make_float4(sqrtf(a_t[i]), sqrtf(1.f - a_t[i]), sqrtf(a_prev[i]), sqrtf(1.f - a_prev[i]));
  float* d_ttab = (float*)(u->d_coef + n_steps);
  // (pageable sources: cudaMemcpyAsync returns once they have been copied to the driver's staging memory, so the host buffers
  // may go out of scope -- and the host may run a whole call ahead of the GPU: no synchronisation here)
  DG_CUDA(cudaMemcpyAsync(u->d_coef, cf.data(), sizeof(float4) * n_steps, cudaMemcpyHostToDevice, s));
  DG_CUDA(cudaMemcpyAsync(d_ttab, t_host, sizeof(float) * n_steps, cudaMemcpyHostToDevice, s));
  u->last_launches = 0;
  const long long c_h = g_launch_counter;
  DG_TRY(hoist_loop_invariants(u, s, (const __half*)ehs, tokens, B, d_ttab, t_host, n_steps));
  u->last_launches += g_launch_counter - c_h;
  u->kv_ready = true; u->temb_ready = true;
  u->cfg_pairs = cfg_on;      // dup_latents_kernel below makes rows [0, n) and [n, 2n) of the UNet input identical
  struct Unhoist { dg_unet* u; ~Unhoist() { u->kv_ready = false; u->temb_ready = false; u->cfg_pairs = false; } } unhoist{u};
  for (int i = 0; i < n_steps; ++i) {
    set_step_kernel<<<1, 256, 0, s>>>(u->d_step, u->d_t, d_ttab, i, B, u->temb_table, u->temb_cur, u->temb_total);
    DG_LAUNCH_CHECK();
    dup_latents_kernel<<<grid_for(elems / 8, 256, u->ctx->num_sms), 256, 0, s>>>((const __half*)latents, u->loop_in, elems / 8, cfg_on ? 2 : 1);
    DG_LAUNCH_CHECK();
    u->last_launches += 2;
    DG_TRY(forward_maybe_graph(u, s, u->loop_in, (const __half*)ehs, tokens, u->loop_out, B, h, w));
    cfg_ddim_step_kernel<<<grid_for(elems / 8, 256, u->ctx->num_sms), 256, 0, s>>>(u->loop_out, (__half*)latents, elems / 8, elems,
                                                                                 guidance, cfg_on, pred_type, u->d_coef, u->d_step);
    DG_LAUNCH_CHECK();
    u->last_launches += 1;
  }
  return DG_OK;
}

// ------------------------------------------------------------------ single operators
int32_t dg_op_gemm(dg_ctx* ctx, const void* A, const void* W, const void* bias, const void* residual, void* out, int32_t M,
                   int32_t K, int32_t n_w, int32_t n_out, int32_t geglu, void* stream) {
  if (!ctx || !A || !W || !out) return fail(DG_E_ARG, "null argument");
  GemmArgs a; a.a0 = (const __half*)A; a.c0 = K; a.B = 1; a.H = 1; a.W = M; a.taps = 1; a.w = (const __half*)W; a.n_w = n_w;
  a.n_out = n_out; a.bias = (const __half*)bias; a.residual = (const __half*)residual; a.ld_res = n_out; a.geglu = geglu;
  a.out = (__half*)out; a.ldo = n_out;
  return launch_gemm((cudaStream_t)stream, ctx->gemm, a);
}
int32_t dg_op_geglu_packed_rows(int32_t inner) { return geglu_rows(inner); }
int32_t dg_op_gemm_row_parts(int32_t n_out) { return gemm_row_parts(n_out); }
int32_t dg_op_gemm_fused(dg_ctx* ctx, const void* A, const void* W, const void* bias, const float* bias32, const float* colsum,
                         const float* ln_stats, int32_t ln_c, float ln_eps, const void* residual, void* out, int32_t M, int32_t K,
                         int32_t n_w, int32_t n_out, int32_t geglu, float* row_stats_out, float* gn_stats_out, int32_t gn_blk,
                         int32_t hw, void* stream) {
  if (!ctx || !A || !W || !out) return fail(DG_E_ARG, "null argument");
  GemmArgs a; a.a0 = (const __half*)A; a.c0 = K; a.B = 1; a.H = 1; a.W = M; a.taps = 1; a.w = (const __half*)W; a.n_w = n_w;
  a.n_out = n_out; a.bias = (const __half*)bias; a.bias32 = bias32; a.colsum = colsum; a.ln_stats = ln_stats;
  a.ln_parts = gemm_row_parts(ln_c); a.ln_c = ln_c; a.ln_eps = ln_eps;
  a.residual = (const __half*)residual; a.ld_res = n_out; a.geglu = geglu; a.out = (__half*)out; a.ldo = n_out;
  a.row_stats_out = row_stats_out; a.gn_stats_out = gn_stats_out; a.gn_blk = gn_blk; a.hw = hw;
  return launch_gemm((cudaStream_t)stream, ctx->gemm, a);
}
int32_t dg_op_fold_layernorm(dg_ctx* ctx, const void* W, const void* bias, const void* gamma, const void* beta, void* Wf,
                             float* colsum, float* b32, int32_t N, int32_t K, void* stream) {
  if (!ctx || !W || !gamma || !beta || !Wf || !colsum || !b32) return fail(DG_E_ARG, "null argument");
  fold_layernorm_kernel<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream>>>((const __half*)W, (const __half*)bias, (const __half*)gamma,
                                                                      (const __half*)beta, (__half*)Wf, colsum, b32, N, K);
  DG_LAUNCH_CHECK();
  return DG_OK;
}
int32_t dg_op_row_stats(dg_ctx* ctx, const void* x, float* stats, int32_t rows, int32_t C, int32_t parts, void* stream) {
  if (!ctx || !x || !stats || parts <= 0) return fail(DG_E_ARG, "bad argument");
  row_stats_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>((const __half*)x, stats, rows, C, parts);
  DG_LAUNCH_CHECK();
  return DG_OK;
}
int32_t dg_op_conv3x3_stats(dg_ctx* ctx, const void* x0, int32_t C0, const void* Wp, const void* bias, void* out, int32_t B,
                            int32_t H, int32_t Wd, int32_t N, float* gn_stats_out, int32_t gn_blk, void* stream) {
  if (!ctx || !x0 || !Wp || !out || !gn_stats_out) return fail(DG_E_ARG, "null argument");
  GemmArgs a; a.a0 = (const __half*)x0; a.c0 = C0; a.B = B; a.H = H; a.W = Wd; a.taps = 9;
  a.w = (const __half*)Wp; a.n_w = N; a.n_out = N; a.bias = (const __half*)bias; a.out = (__half*)out; a.ldo = N;
  a.gn_stats_out = gn_stats_out; a.gn_blk = gn_blk;
  return launch_gemm((cudaStream_t)stream, ctx->gemm, a);
}
int32_t dg_op_groupnorm_fused(dg_ctx* ctx, const void* x0, int32_t C0, const float* stats0, const void* x1, int32_t C1,
                              const float* stats1, int32_t blk, const void* gamma, const void* beta, void* out, int32_t B,
                              int32_t HW, int32_t groups, float eps, int32_t silu, void* stream) {
  if (!ctx || !x0 || !stats0 || !gamma || !beta || !out || (x1 && !stats1)) return fail(DG_E_ARG, "null argument");
  return launch_groupnorm_fused((cudaStream_t)stream, ctx->num_sms, (const __half*)x0, C0, stats0, (const __half*)x1, C1, stats1,
                                blk, (const __half*)gamma, (const __half*)beta, (__half*)out, B, HW, groups, eps, silu);
}

int32_t dg_op_pack_geglu(dg_ctx* ctx, const void* w, const void* b, void* w_out, void* b_out, int32_t inner, int32_t K, void* stream) {
  if (!ctx || !w || !b || !w_out || !b_out) return fail(DG_E_ARG, "null argument");
  const int tiles = geglu_rows(inner) / kGegluTile;
  cudaStream_t s = (cudaStream_t)stream;
  pack_geglu_kernel<<<grid_for((size_t)tiles * kGegluTile * K, 256, ctx->num_sms), 256, 0, s>>>((const __half*)w, (__half*)w_out, inner, K, kGegluTile, kGegluHalf, tiles);
  DG_LAUNCH_CHECK();
  pack_geglu_kernel<<<grid_for((size_t)tiles * kGegluTile, 256, ctx->num_sms), 256, 0, s>>>((const __half*)b, (__half*)b_out, inner, 1, kGegluTile, kGegluHalf, tiles);
  DG_LAUNCH_CHECK();
  return DG_OK;
}
int32_t dg_op_pack_conv3x3(dg_ctx* ctx, const void* w, void* w_out, int32_t O, int32_t I, void* stream) {
  if (!ctx || !w || !w_out) return fail(DG_E_ARG, "null argument");
  pack_conv3x3_kernel<<<grid_for((size_t)O * 9 * I, 256, ctx->num_sms), 256, 0, (cudaStream_t)stream>>>((const __half*)w, (__half*)w_out, O, I, I);
  DG_LAUNCH_CHECK();
  return DG_OK;
}
int32_t dg_op_conv3x3(dg_ctx* ctx, const void* x0, int32_t C0, const void* x1, int32_t C1, const void* Wp, const void* bias,
                      const void* rowvec, int32_t ld_rowvec, const void* residual, void* out, int32_t B, int32_t H, int32_t Wd,
                      int32_t N, int32_t ldo, void* stream) {
  if (!ctx || !x0 || !Wp || !out) return fail(DG_E_ARG, "null argument");
  if (ldo <= 0) ldo = N;
  GemmArgs a; a.a0 = (const __half*)x0; a.c0 = C0; a.a1 = (const __half*)x1; a.c1 = C1; a.B = B; a.H = H; a.W = Wd; a.taps = 9;
  a.w = (const __half*)Wp; a.n_w = N; a.n_out = N; a.bias = (const __half*)bias; a.rowvec = (const __half*)rowvec; a.ld_rowvec = ld_rowvec;
  a.residual = (const __half*)residual; a.ld_res = ldo; a.out = (__half*)out; a.ldo = ldo;
  return launch_gemm((cudaStream_t)stream, ctx->gemm, a);
}
int32_t dg_op_conv3x3_gn(dg_ctx* ctx, const void* x0, int32_t C0, const float* stats0, const void* x1, int32_t C1, const float* stats1,
                         int32_t blk, const void* gamma, const void* beta, int32_t groups, float eps, int32_t silu, const void* Wp,
                         const void* bias, const void* residual, void* out, int32_t B, int32_t H, int32_t Wd, int32_t N, int32_t ldo,
                         int32_t taps, void* stream) {
  if (!ctx || !x0 || !stats0 || !gamma || !beta || !Wp || !out || (x1 && !stats1)) return fail(DG_E_ARG, "null argument");
  if (taps != 1 && taps != 9) return fail(DG_E_ARG, "taps must be 1 or 9");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t need = (size_t)B * 3 * (C0 + C1);
  if (need > ctx->xf_tab_cap) {
    DG_CUDA(cudaStreamSynchronize(s));
    cudaFree(ctx->xf_tab); ctx->xf_tab = nullptr; ctx->xf_tab_cap = 0;
    DG_CUDA(cudaMalloc((void**)&ctx->xf_tab, need * sizeof(__half)));
    ctx->xf_tab_cap = need;
  }
  DG_TRY(launch_gn_fold(s, C0, stats0, x1 ? C1 : 0, stats1, blk, (const __half*)gamma, (const __half*)beta, ctx->xf_tab, B, H * Wd, groups, eps, silu));
  GemmArgs a; a.a0 = (const __half*)x0; a.c0 = C0; a.a1 = (const __half*)x1; a.c1 = x1 ? C1 : 0; a.B = B; a.H = H; a.W = Wd; a.taps = taps;
  a.hw = H * Wd;
  if (ldo <= 0) ldo = N;
  a.w = (const __half*)Wp; a.n_w = N; a.n_out = N; a.bias = (const __half*)bias; a.residual = (const __half*)residual; a.ld_res = ldo;
  a.out = (__half*)out; a.ldo = ldo; a.xf_tab = ctx->xf_tab; a.xf_silu = silu;
  return launch_gemm(s, ctx->gemm, a);
}
int32_t dg_op_conv3x3_shortcut(dg_ctx* ctx, const void* x, int32_t C, const void* Wcat, const void* bias, const void* xs0, int32_t Cs0,
                               const void* xs1, int32_t Cs1, void* out, int32_t B, int32_t H, int32_t Wd, int32_t N, void* stream) {
  if (!ctx || !x || !Wcat || !xs0 || !out) return fail(DG_E_ARG, "null argument");
  GemmArgs a; a.a0 = (const __half*)x; a.c0 = C; a.B = B; a.H = H; a.W = Wd; a.taps = 9; a.w = (const __half*)Wcat; a.n_w = N; a.n_out = N;
  a.bias = (const __half*)bias; a.x0 = (const __half*)xs0; a.cx0 = Cs0; a.x1 = (const __half*)xs1; a.cx1 = xs1 ? Cs1 : 0;
  a.out = (__half*)out; a.ldo = N;
  return launch_gemm((cudaStream_t)stream, ctx->gemm, a);
}
int32_t dg_op_conv3x3_stride2(dg_ctx* ctx, const void* x, int32_t C, const void* Wp, const void* bias, void* out, int32_t B, int32_t Hout,
                              int32_t Wout, int32_t N, float* gn_stats_out, int32_t gn_blk, void* stream) {
  if (!ctx || !x || !Wp || !out) return fail(DG_E_ARG, "null argument");
  GemmArgs a; a.a0 = (const __half*)x; a.c0 = C; a.B = B; a.H = Hout; a.W = Wout; a.taps = 9; a.in_stride = 2;
  a.w = (const __half*)Wp; a.n_w = N; a.n_out = N; a.bias = (const __half*)bias; a.out = (__half*)out; a.ldo = N;
  a.gn_stats_out = gn_stats_out; a.gn_blk = gn_blk;
  return launch_gemm((cudaStream_t)stream, ctx->gemm, a);
}
int32_t dg_op_upsample_conv3x3(dg_ctx* ctx, const void* x, int32_t C, const void* w_oihw, const void* bias, void* out, int32_t B, int32_t H,
                               int32_t Wd, int32_t N, float* gn_stats_out, int32_t gn_blk, void* stream) {
  if (!ctx || !x || !w_oihw || !out) return fail(DG_E_ARG, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  __half* wph = nullptr;
  DG_CUDA(cudaMalloc((void**)&wph, (size_t)16 * C * N * 2));
  pack_upconv_kernel<<<grid_for((size_t)16 * C * N, 256, ctx->num_sms), 256, 0, s>>>((const __half*)w_oihw, wph, N, C);
  int r = DG_OK;
  for (int ph = 0; ph < 4 && r == DG_OK; ++ph) {
    GemmArgs a; a.a0 = (const __half*)x; a.c0 = C; a.B = B; a.H = H; a.W = Wd; a.taps = 4; a.phase_y = ph >> 1; a.phase_x = ph & 1;
    a.w = wph + (size_t)ph * N * 4 * C; a.n_w = N; a.n_out = N; a.bias = (const __half*)bias; a.out = (__half*)out; a.ldo = N;
    if (gn_stats_out) { a.gn_stats_out = gn_stats_out; a.gn_blk = gn_blk; a.gn_slot0 = ph * (H * Wd / 32); a.gn_slots = 4 * (H * Wd / 32); }
    r = launch_gemm(s, ctx->gemm, a);
  }
  cudaStreamSynchronize(s);
  cudaFree(wph);
  return r;
}
int32_t dg_op_attention(dg_ctx* ctx, const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv, void* out,
                        int32_t B, int32_t heads, int32_t Sq, int32_t Sk, int32_t d, void* stream) {
  if (!ctx || !q || !k || !v || !out) return fail(DG_E_ARG, "null argument");
  return launch_attention((cudaStream_t)stream, (const __half*)q, ldq, (const __half*)k, ldk, (const __half*)v, ldv, (__half*)out, B, heads, Sq, Sk, d);
}
int32_t dg_op_groupnorm(dg_ctx* ctx, const void* x0, int32_t C0, const void* x1, int32_t C1, const void* gamma, const void* beta,
                        void* out, int32_t B, int32_t HW, int32_t groups, float eps, int32_t silu, void* stream) {
  if (!ctx || !x0 || !gamma || !beta || !out) return fail(DG_E_ARG, "null argument");
  if (B * groups > 64 * 64) return fail(DG_E_SHAPE, "groupnorm op: B*groups too large for the context scratch");
  return launch_groupnorm((cudaStream_t)stream, ctx->num_sms, (const __half*)x0, C0, (const __half*)x1, C1, (const __half*)gamma,
                          (const __half*)beta, (__half*)out, ctx->gn_stats, B, HW, groups, eps, silu);
}
int32_t dg_op_layernorm(dg_ctx* ctx, const void* x, const void* gamma, const void* beta, void* out, int32_t rows, int32_t C, float eps, void* stream) {
  if (!ctx || !x || !gamma || !beta || !out) return fail(DG_E_ARG, "null argument");
  return launch_layernorm((cudaStream_t)stream, (const __half*)x, (const __half*)gamma, (const __half*)beta, (__half*)out, rows, C, eps);
}
int32_t dg_op_time_embedding(dg_ctx* ctx, const float* t_host, int32_t B, int32_t dim, int32_t temb_dim, const void* w1, const void* b1,
                             const void* w2, const void* b2, void* out, void* stream) {
  if (!ctx || !t_host || !w1 || !w2 || !out) return fail(DG_E_ARG, "null argument");
  if (B > 8) return fail(DG_E_SHAPE, "time embedding op: B <= 8");
  cudaStream_t s = (cudaStream_t)stream;
  float* d_t = nullptr; __half* tsin = nullptr; __half* t1 = nullptr;
  DG_CUDA(cudaMalloc((void**)&d_t, sizeof(float) * B));
  DG_CUDA(cudaMalloc((void**)&tsin, (size_t)B * dim * 2));
  DG_CUDA(cudaMalloc((void**)&t1, (size_t)B * temb_dim * 2));
  DG_CUDA(cudaMemcpyAsync(d_t, t_host, sizeof(float) * B, cudaMemcpyHostToDevice, s));
  timestep_sinusoid_kernel<<<(B * dim / 2 + 127) / 128, 128, 0, s>>>(d_t, tsin, B, dim, 0.f, 1);
  DG_LAUNCH_CHECK();
  int r = launch_gemv(s, tsin, dim, (const __half*)w1, (const __half*)b1, t1, temb_dim, B, temb_dim, dim, 0, 1);
  if (r == DG_OK) r = launch_gemv(s, t1, temb_dim, (const __half*)w2, (const __half*)b2, (__half*)out, temb_dim, B, temb_dim, temb_dim, 0, 0);
  cudaStreamSynchronize(s);
  cudaFree(d_t); cudaFree(tsin); cudaFree(t1);
  return r;
}


// ---- AutoencoderKL decoder -------------------------------------------------------------------------------------------
int32_t dg_vae_create(dg_ctx* ctx, const int32_t* block_out_channels, int32_t layers_per_block, dg_vae** out) {
  if (!ctx || !out) return fail(DG_E_ARG, "null argument");
  std::unique_ptr<dg_vae> v(new dg_vae());
  v->ctx = ctx;
  if (block_out_channels) for (int i = 0; i < 4; ++i) v->ch[i] = block_out_channels[i];
  if (layers_per_block > 0) v->layers = layers_per_block;
  for (int i = 0; i < 4; ++i)
    if (v->ch[i] % 64 || v->ch[i] % v->groups) return fail(DG_E_UNSUPPORTED, "vae block_out_channels[%d]=%d must be a multiple of 64", i, v->ch[i]);
  DG_CUDA(cudaSetDevice(ctx->device));
  int r = build_vae(v.get());
  if (r != DG_OK) { for (void* p : v->owned) cudaFree(p); return r; }
  *out = v.release();
  return DG_OK;
}
void dg_vae_destroy(dg_vae* v) {
  if (!v) return;
  for (void* p : v->owned) cudaFree(p);
  cudaFree(v->arena.base); cudaFree(v->gn_stats);
  delete v;
}
int32_t dg_vae_num_weights(dg_vae* v) { return v ? (int32_t)v->slots.size() : 0; }
const char* dg_vae_weight_name(dg_vae* v, int32_t i) {
  if (!v || i < 0 || i >= (int)v->slots.size()) return nullptr;
  return v->slots[i].key.c_str();
}
int32_t dg_vae_weight_shape(dg_vae* v, int32_t i, int64_t* shape4, int32_t* ndim) {
  if (!v || i < 0 || i >= (int)v->slots.size() || !shape4 || !ndim) return fail(DG_E_ARG, "bad argument");
  *ndim = (int32_t)v->slots[i].shape.size();
  for (int k = 0; k < *ndim; ++k) shape4[k] = v->slots[i].shape[k];
  return DG_OK;
}
int32_t dg_vae_set_weight(dg_vae* v, const char* key, const void* src, int32_t ndim, const int64_t* shape) {
  return store_set_weight(v, key, src, ndim, shape);
}
int32_t dg_vae_prepare(dg_vae* v, int32_t max_batch, int32_t h, int32_t w) {
  if (!v || max_batch <= 0 || h <= 0 || w <= 0) return fail(DG_E_ARG, "bad argument");
  DG_CUDA(cudaSetDevice(v->ctx->device));
  cudaFree(v->arena.base); cudaFree(v->gn_stats);
  v->arena.base = nullptr; v->gn_stats = nullptr;
  // peak live set: ~6 full-resolution tensors of ch[0] channels at 8h x 8w (+ the S x S score matrix of one image)
  const size_t pix = (size_t)max_batch * (8 * h) * (8 * w);
  const size_t S = (size_t)h * w;
  size_t bytes = pix * v->ch[0] * 2 * 8 + S * S * 2 + ((size_t)64 << 20);
  v->arena.size = bytes;
  cudaError_t e = cudaMalloc((void**)&v->arena.base, bytes);
  if (e != cudaSuccess) return fail(DG_E_NOMEM, "VAE workspace cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  DG_CUDA(cudaMalloc((void**)&v->gn_stats, sizeof(float) * 2 * v->groups * max_batch));
  v->max_batch = max_batch; v->ws_h = h; v->ws_w = w;
  return DG_OK;
}
int32_t dg_vae_decode(dg_vae* v, const void* latents, float scale, void* out, int32_t B, int32_t h, int32_t w, void* stream) {
  if (!v || !latents || !out) return fail(DG_E_ARG, "null argument");
  if (!v->arena.base) return fail(DG_E_STATE, "dg_vae_prepare has not been called");
  if (B > v->max_batch || h * w > v->ws_h * v->ws_w) return fail(DG_E_SHAPE, "decode (batch %d, %dx%d) exceeds prepared workspace", B, h, w);
  if (h % 8 || w % 8) return fail(DG_E_SHAPE, "latent size %dx%d must be a multiple of 8", h, w);
  if (vae_missing(v)) return fail(DG_E_STATE, "%d VAE weights have not been set", vae_missing(v));
  DG_CUDA(cudaSetDevice(v->ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int sms = v->ctx->num_sms;
  v->arena.reset();
  VFwd f{v, s};
  // post_quant_conv (+ 1/scaling_factor) -> conv_in (4-channel NCHW gather -> K = 64 GEMM)
  __half* z = f.alloc((size_t)B * v->latent_ch * h * w * 2);
  __half* col = f.alloc((size_t)B * h * w * 64 * 2);
  T4 x = f.talloc(B, h, w, v->ch[3]);
  if (f.err) return f.err;
  latent_pointwise_kernel<<<(unsigned)(((size_t)B * h * w + 255) / 256), 256, 0, s>>>((const __half*)latents, v->pq_w, v->pq_b, z, B, v->latent_ch, h * w, scale);
  DG_LAUNCH_CHECK();
  im2col_conv_in_kernel<<<grid_for((size_t)B * h * w * 64, 256, sms), 256, 0, s>>>(z, col, B, v->latent_ch, h, w, 64);
  DG_LAUNCH_CHECK();
  f.linear(col, 64, B * h * w, v->conv_in.w, v->conv_in.rows, v->conv_in.b, nullptr, x.p, v->ch[3]);
  v->arena.release(z); v->arena.release(col);
  // mid block
  x = f.resnet(v->mid0, x);
  if (!f.err) x = f.attention(x);
  if (!f.err) x = f.resnet(v->mid1, x);
  // up blocks
  for (int i = 0; i < 4 && !f.err; ++i) {
    VUp& u = v->up[i];
    for (size_t j = 0; j < u.res.size() && !f.err; ++j) x = f.resnet(u.res[j], x);
    if (u.has_up && !f.err) {
      T4 y = f.talloc(x.B, x.H * 2, x.W * 2, x.C);
      if (f.err) break;
      if (v->ctx->up_phases && v->ctx->gemm.cta_mode != 1) {
        // nearest x2 + conv3x3 as four 2x2 phase convolutions on the low-resolution tensor (see Fwd::upconv)
        for (int ph = 0; ph < 4 && !f.err; ++ph) {
          GemmArgs a; a.a0 = x.p; a.c0 = x.C; a.B = x.B; a.H = x.H; a.W = x.W; a.taps = 4; a.phase_y = ph >> 1; a.phase_x = ph & 1;
          a.w = u.up.wph + (size_t)ph * u.up.out * 4 * u.up.in; a.n_w = u.up.rows; a.n_out = u.up.out; a.bias = u.up.b; a.out = y.p; a.ldo = u.up.out;
          f.err = launch_gemm(s, v->ctx->gemm, a);
        }
      } else {
        T4 upx = f.talloc(x.B, x.H * 2, x.W * 2, x.C);
        if (f.err) break;
        upsample2x_nhwc_kernel<<<grid_for((size_t)x.B * 4 * x.H * x.W * (x.C / 8), 256, sms), 256, 0, s>>>(x.p, upx.p, x.B, x.H, x.W, x.C);
        DG_LAUNCH_CHECK();
        f.conv3(upx, u.up, nullptr, y);
        f.free_(upx);
      }
      f.free_(x);
      x = y;
    }
  }
  if (f.err) return f.err;
  T4 xn = f.talloc(x.B, x.H, x.W, x.C);
  f.gn(x, v->norm_out, 1, xn);
  const int opitch = (v->out_ch + 7) / 8 * 8;
  T4 o = f.talloc(x.B, x.H, x.W, opitch);
  if (f.err) return f.err;
  f.conv3(xn, v->conv_out, nullptr, o, opitch);
  if (f.err) return f.err;
  const size_t n = (size_t)B * v->out_ch * x.H * x.W;
  nhwc_to_nchw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(o.p, (__half*)out, B, v->out_ch, x.H * x.W, opitch);
  DG_LAUNCH_CHECK();
  return DG_OK;
}


// ---- CLIP text encoder ------------------------------------------------------------------------------------------------
int32_t dg_clip_create(dg_ctx* ctx, int32_t vocab, int32_t hidden, int32_t intermediate, int32_t layers, int32_t heads,
                       int32_t max_positions, dg_clip** out) {
  if (!ctx || !out) return fail(DG_E_ARG, "null argument");
  if (hidden <= 0 || heads <= 0 || hidden != heads * 64) return fail(DG_E_UNSUPPORTED, "clip: head dim must be 64 (hidden %d, heads %d)", hidden, heads);
  if (hidden % 64 || hidden > 1280 || intermediate % 64 || vocab <= 0 || layers <= 0 || max_positions <= 0)
    return fail(DG_E_UNSUPPORTED, "clip: unsupported configuration");
  std::unique_ptr<dg_clip> c(new dg_clip());
  c->ctx = ctx; c->vocab = vocab; c->hidden = hidden; c->inter = intermediate; c->layers = layers; c->heads = heads; c->max_pos = max_positions;
  DG_CUDA(cudaSetDevice(ctx->device));
  int r = build_clip(c.get());
  if (r != DG_OK) { for (void* p : c->owned) cudaFree(p); return r; }
  *out = c.release();
  return DG_OK;
}
void dg_clip_destroy(dg_clip* c) {
  if (!c) return;
  for (void* p : c->owned) cudaFree(p);
  cudaFree(c->ws); cudaFree(c->ids_dev);
  delete c;
}
int32_t dg_clip_num_weights(dg_clip* c) { return c ? (int32_t)c->slots.size() : 0; }
const char* dg_clip_weight_name(dg_clip* c, int32_t i) {
  if (!c || i < 0 || i >= (int)c->slots.size()) return nullptr;
  return c->slots[i].key.c_str();
}
int32_t dg_clip_weight_shape(dg_clip* c, int32_t i, int64_t* shape4, int32_t* ndim) {
  if (!c || i < 0 || i >= (int)c->slots.size() || !shape4 || !ndim) return fail(DG_E_ARG, "bad argument");
  *ndim = (int32_t)c->slots[i].shape.size();
  for (int k = 0; k < *ndim; ++k) shape4[k] = c->slots[i].shape[k];
  return DG_OK;
}
int32_t dg_clip_set_weight(dg_clip* c, const char* key, const void* src, int32_t ndim, const int64_t* shape) {
  return store_set_weight(c, key, src, ndim, shape);
}
int32_t dg_clip_set_activation(dg_clip* c, int32_t act) {
  if (!c || (act != 0 && act != 1)) return fail(DG_E_ARG, "clip: activation must be 0 (quick_gelu) or 1 (gelu)");
  c->act = act;
  return DG_OK;
}
int32_t dg_clip_prepare(dg_clip* c, int32_t max_batch) {
  if (!c || max_batch <= 0) return fail(DG_E_ARG, "bad argument");
  DG_CUDA(cudaSetDevice(c->ctx->device));
  cudaFree(c->ws); cudaFree(c->ids_dev);
  c->ws = nullptr; c->ids_dev = nullptr;
  const size_t rows = (size_t)max_batch * c->max_pos;
  // x, x2, h, att [rows, C]; qkv [rows, 3C]; f [rows, inter]
  const size_t elems = rows * ((size_t)7 * c->hidden + c->inter);
  cudaError_t e = cudaMalloc((void**)&c->ws, elems * 2);
  if (e != cudaSuccess) return fail(DG_E_NOMEM, "clip workspace cudaMalloc(%zu) failed: %s", elems * 2, cudaGetErrorString(e));
  DG_CUDA(cudaMalloc((void**)&c->ids_dev, rows * sizeof(int)));
  c->max_batch = max_batch;
  return DG_OK;
}
int32_t dg_clip_encode(dg_clip* c, const int32_t* input_ids, int32_t batch, int32_t seq, void* out, void* stream) {
  if (!c || !input_ids || !out) return fail(DG_E_ARG, "null argument");
  if (!c->ws) return fail(DG_E_STATE, "dg_clip_prepare has not been called");
  if (batch <= 0 || batch > c->max_batch || seq <= 0 || seq > c->max_pos) return fail(DG_E_SHAPE, "clip: batch %d x %d tokens exceeds the prepared workspace", batch, seq);
  for (const Slot& sl : c->slots) if (!sl.set) return fail(DG_E_STATE, "clip weight %s has not been set", sl.key.c_str());
  DG_CUDA(cudaSetDevice(c->ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int C = c->hidden, rows = batch * seq;
  const size_t R = (size_t)c->max_batch * c->max_pos;
  __half* x = c->ws; __half* x2 = x + R * C; __half* h = x2 + R * C; __half* att = h + R * C;
  __half* qkv = att + R * C; __half* f = qkv + R * 3 * C;
  DG_CUDA(cudaMemcpyAsync(c->ids_dev, input_ids, (size_t)rows * sizeof(int), cudaMemcpyHostToDevice, s));
  clip_embed_kernel<<<(unsigned)(((size_t)rows * (C / 8) + 255) / 256), 256, 0, s>>>(c->ids_dev, c->tok, c->pos, x, rows, seq, C, c->vocab);
  DG_LAUNCH_CHECK();
  DG_TRY(run_clip_layers(c->ctx, s, c->L, x, x2, h, att, qkv, f, batch, seq, C, c->inter, c->heads, /*causal=*/1, c->eps, c->act));
  DG_TRY(launch_layernorm(s, x, c->final_ln.g, c->final_ln.b, (__half*)out, rows, C, c->eps));
  return DG_OK;
}


int32_t dg_op_image_to_uint8(dg_ctx* ctx, const void* img, void* out_u8, int32_t B, int32_t C, int32_t H, int32_t W, void* stream) {
  if (!ctx || !img || !out_u8 || B <= 0 || C <= 0 || C > 4 || H <= 0 || W <= 0) return fail(DG_E_ARG, "bad argument");
  DG_CUDA(cudaSetDevice(ctx->device));
  const size_t n = (size_t)B * H * W;
  image_to_uint8_hwc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __half*)img, (unsigned char*)out_u8, B, C, H * W);
  DG_LAUNCH_CHECK();
  return DG_OK;
}


// ---- CLIP similarity scorer -------------------------------------------------------------------------------------------
int32_t dg_clipscore_create(dg_ctx* ctx, const int32_t* text_cfg6, const int32_t* vision_cfg6, int32_t projection_dim, dg_clipscore** out) {
  if (!ctx || !out) return fail(DG_E_ARG, "null argument");
  std::unique_ptr<dg_clipscore> c(new dg_clipscore());
  c->ctx = ctx;
  if (text_cfg6) { c->vocab = text_cfg6[0]; c->t_hidden = text_cfg6[1]; c->t_inter = text_cfg6[2]; c->t_layers = text_cfg6[3]; c->t_heads = text_cfg6[4]; c->max_pos = text_cfg6[5]; }
  if (vision_cfg6) { c->image = vision_cfg6[0]; c->patch = vision_cfg6[1]; c->v_hidden = vision_cfg6[2]; c->v_inter = vision_cfg6[3]; c->v_layers = vision_cfg6[4]; c->v_heads = vision_cfg6[5]; }
  if (projection_dim > 0) c->proj = projection_dim;
  if (c->t_hidden != 64 * c->t_heads || c->v_hidden != 64 * c->v_heads) return fail(DG_E_UNSUPPORTED, "clipscore: head dim must be 64 in both towers");
  if (c->t_hidden % 64 || c->v_hidden % 64 || c->t_hidden > 1280 || c->v_hidden > 1280 || c->t_inter % 64 || c->v_inter % 64 || c->proj % 8 ||
      c->patch <= 0 || c->image % c->patch || c->vocab <= 0 || c->max_pos <= 0)
    return fail(DG_E_UNSUPPORTED, "clipscore: unsupported configuration");
  DG_CUDA(cudaSetDevice(ctx->device));
  int r = build_clipscore(c.get());
  if (r != DG_OK) { for (void* p : c->owned) cudaFree(p); return r; }
  *out = c.release();
  return DG_OK;
}
void dg_clipscore_destroy(dg_clipscore* c) {
  if (!c) return;
  for (void* p : c->owned) cudaFree(p);
  cudaFree(c->ws); cudaFree(c->ids_dev); cudaFree(c->eos_dev);
  delete c;
}
int32_t dg_clipscore_num_weights(dg_clipscore* c) { return c ? (int32_t)c->slots.size() : 0; }
const char* dg_clipscore_weight_name(dg_clipscore* c, int32_t i) {
  if (!c || i < 0 || i >= (int)c->slots.size()) return nullptr;
  return c->slots[i].key.c_str();
}
int32_t dg_clipscore_weight_shape(dg_clipscore* c, int32_t i, int64_t* shape4, int32_t* ndim) {
  if (!c || i < 0 || i >= (int)c->slots.size() || !shape4 || !ndim) return fail(DG_E_ARG, "bad argument");
  *ndim = (int32_t)c->slots[i].shape.size();
  for (int k = 0; k < *ndim; ++k) shape4[k] = c->slots[i].shape[k];
  return DG_OK;
}
int32_t dg_clipscore_set_weight(dg_clipscore* c, const char* key, const void* src, int32_t ndim, const int64_t* shape) {
  return store_set_weight(c, key, src, ndim, shape);
}
int32_t dg_clipscore_set_logit_scale(dg_clipscore* c, float logit_scale) {
  if (!c) return fail(DG_E_ARG, "null argument");
  c->logit_scale = logit_scale;
  return DG_OK;
}
int32_t dg_clipscore_prepare(dg_clipscore* c, int32_t max_images, int32_t max_texts) {
  if (!c || max_images <= 0 || max_texts <= 0) return fail(DG_E_ARG, "bad argument");
  DG_CUDA(cudaSetDevice(c->ctx->device));
  cudaFree(c->ws); cudaFree(c->ids_dev); cudaFree(c->eos_dev);
  c->ws = nullptr; c->ids_dev = nullptr; c->eos_dev = nullptr;
  const int n = c->image / c->patch, S = n * n + 1;
  const size_t vrows = (size_t)max_images * S, trows = (size_t)max_texts * c->max_pos;
  const size_t v_elems = vrows * ((size_t)7 * c->v_hidden + c->v_inter) + (size_t)max_images * n * n * (c->kpad + c->v_hidden);
  const size_t t_elems = trows * ((size_t)7 * c->t_hidden + c->t_inter);
  const size_t feat = (size_t)(max_images + max_texts) * (c->proj + 2 * (size_t)std::max(c->v_hidden, c->t_hidden));
  c->ws_elems = std::max(v_elems, t_elems) + feat + 1024;
  cudaError_t e = cudaMalloc((void**)&c->ws, c->ws_elems * 2);
  if (e != cudaSuccess) return fail(DG_E_NOMEM, "clipscore workspace cudaMalloc(%zu) failed: %s", c->ws_elems * 2, cudaGetErrorString(e));
  DG_CUDA(cudaMalloc((void**)&c->ids_dev, trows * sizeof(int)));
  DG_CUDA(cudaMalloc((void**)&c->eos_dev, (size_t)max_texts * sizeof(int)));
  c->max_images = max_images; c->max_texts = max_texts;
  return DG_OK;
}
int32_t dg_clipscore_score(dg_clipscore* c, const void* pixel_values, int32_t n_images, const int32_t* input_ids, const int32_t* eos_index,
                           int32_t n_texts, int32_t seq, float* logits_per_text, void* stream) {
  if (!c || !pixel_values || !input_ids || !eos_index || !logits_per_text) return fail(DG_E_ARG, "null argument");
  if (!c->ws) return fail(DG_E_STATE, "dg_clipscore_prepare has not been called");
  if (n_images <= 0 || n_images > c->max_images || n_texts <= 0 || n_texts > c->max_texts || seq <= 0 || seq > c->max_pos)
    return fail(DG_E_SHAPE, "clipscore: %d images / %d texts x %d tokens exceed the prepared workspace", n_images, n_texts, seq);
  for (const Slot& sl : c->slots) if (!sl.set) return fail(DG_E_STATE, "clipscore weight %s has not been set", sl.key.c_str());
  DG_CUDA(cudaSetDevice(c->ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int sms = c->ctx->num_sms;
  const int n = c->image / c->patch, P = n * n, S = P + 1, Cv = c->v_hidden, Ct = c->t_hidden, D = c->proj;
  // features live at the top of the workspace, the towers share the rest
  const int Cmax = std::max(Cv, Ct);
  __half* feat_img = c->ws; __half* feat_txt = feat_img + (size_t)c->max_images * D;
  __half* pooled = feat_txt + (size_t)c->max_texts * D; __half* pooled_n = pooled + (size_t)(c->max_images + c->max_texts) * Cmax;
  __half* base = pooled_n + (size_t)(c->max_images + c->max_texts) * Cmax;
  base = (__half*)(((uintptr_t)base + 255) & ~(uintptr_t)255);
  {   // ---- image tower
    const size_t R = (size_t)c->max_images * S;
    __half* x = base; __half* x2 = x + R * Cv; __half* h = x2 + R * Cv; __half* att = h + R * Cv; __half* qkv = att + R * Cv;
    __half* f = qkv + R * 3 * Cv; __half* col = f + R * c->v_inter; __half* pe = col + (size_t)c->max_images * P * c->kpad;
    const int rows = n_images * S;
    clip_patch_im2col_kernel<<<grid_for((size_t)n_images * P * c->kpad, 256, sms), 256, 0, s>>>((const __half*)pixel_values, col, n_images, 3, c->image, c->patch, c->kpad);
    DG_LAUNCH_CHECK();
    Lin pw; pw.w = c->patch_w; pw.b = nullptr; pw.in = c->kpad; pw.out = Cv; pw.rows = Cv;
    DG_TRY(clip_linear(c->ctx, s, col, c->kpad, n_images * P, pw, nullptr, pe));
    clip_vision_embed_kernel<<<(unsigned)(((size_t)rows * (Cv / 8) + 255) / 256), 256, 0, s>>>(pe, c->cls, c->vpos, x2, n_images, S, Cv);
    DG_LAUNCH_CHECK();
    DG_TRY(launch_layernorm(s, x2, c->pre_ln.g, c->pre_ln.b, x, rows, Cv, c->eps));
    DG_TRY(run_clip_layers(c->ctx, s, c->VL, x, x2, h, att, qkv, f, n_images, S, Cv, c->v_inter, c->v_heads, /*causal=*/0, c->eps));
    gather_rows_kernel<<<(unsigned)(((size_t)n_images * (Cv / 8) + 255) / 256), 256, 0, s>>>(x, nullptr, pooled, n_images, S, Cv);
    DG_LAUNCH_CHECK();
    DG_TRY(launch_layernorm(s, pooled, c->post_ln.g, c->post_ln.b, pooled_n, n_images, Cv, c->eps));
    DG_TRY(clip_linear(c->ctx, s, pooled_n, Cv, n_images, c->v_proj, nullptr, feat_img));
  }
  {   // ---- text tower
    const size_t R = (size_t)c->max_texts * c->max_pos;
    __half* x = base; __half* x2 = x + R * Ct; __half* h = x2 + R * Ct; __half* att = h + R * Ct; __half* qkv = att + R * Ct;
    __half* f = qkv + R * 3 * Ct;
    const int rows = n_texts * seq;
    DG_CUDA(cudaMemcpyAsync(c->ids_dev, input_ids, (size_t)rows * sizeof(int), cudaMemcpyHostToDevice, s));
    DG_CUDA(cudaMemcpyAsync(c->eos_dev, eos_index, (size_t)n_texts * sizeof(int), cudaMemcpyHostToDevice, s));
    clip_embed_kernel<<<(unsigned)(((size_t)rows * (Ct / 8) + 255) / 256), 256, 0, s>>>(c->ids_dev, c->tok, c->tpos, x, rows, seq, Ct, c->vocab);
    DG_LAUNCH_CHECK();
    DG_TRY(run_clip_layers(c->ctx, s, c->TL, x, x2, h, att, qkv, f, n_texts, seq, Ct, c->t_inter, c->t_heads, /*causal=*/1, c->eps));
    DG_TRY(launch_layernorm(s, x, c->t_final.g, c->t_final.b, x2, rows, Ct, c->eps));
    __half* tp = pooled + (size_t)c->max_images * Cmax;
    gather_rows_kernel<<<(unsigned)(((size_t)n_texts * (Ct / 8) + 255) / 256), 256, 0, s>>>(x2, c->eos_dev, tp, n_texts, seq, Ct);
    DG_LAUNCH_CHECK();
    DG_TRY(clip_linear(c->ctx, s, tp, Ct, n_texts, c->t_proj, nullptr, feat_txt));
  }
  const int pairs = n_texts * n_images;
  clip_logits_kernel<<<(pairs + 7) / 8, 256, 0, s>>>(feat_txt, feat_img, logits_per_text, n_texts, n_images, D, expf(c->logit_scale));
  DG_LAUNCH_CHECK();
  return DG_OK;
}


int32_t dg_op_mask_composite_u8(dg_ctx* ctx, const void* img_u8, const void* mask_u8, void* out_u8, uint32_t* count, int32_t B, int32_t H,
                                int32_t W, void* stream) {
  if (!ctx || !img_u8 || !mask_u8 || !out_u8 || !count || B <= 0 || H <= 0 || W <= 0) return fail(DG_E_ARG, "bad argument");
  DG_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  DG_CUDA(cudaMemsetAsync(count, 0, sizeof(uint32_t) * B, s));
  const int hw = H * W;
  dim3 grid((unsigned)std::min((hw + 255) / 256, 4 * ctx->num_sms), (unsigned)B);
  mask_composite_u8_kernel<<<grid, 256, 0, s>>>((const unsigned char*)img_u8, (const unsigned char*)mask_u8, (unsigned char*)out_u8, count, B, hw);
  DG_LAUNCH_CHECK();
  return DG_OK;
}
int32_t dg_op_resample_u8(dg_ctx* ctx, const void* in_u8, void* out_u8, int32_t B, int32_t Hin, int32_t Win, int32_t C, int32_t Hout,
                          int32_t Wout, const int32_t* bounds, const int32_t* coeffs, int32_t ksize, int32_t axis, void* stream) {
  if (!ctx || !in_u8 || !out_u8 || !bounds || !coeffs || B <= 0 || C <= 0 || ksize <= 0 || (axis != 0 && axis != 1)) return fail(DG_E_ARG, "bad argument");
  if ((axis == 0 && Hout != Hin) || (axis == 1 && Wout != Win)) return fail(DG_E_SHAPE, "resample: one axis per pass");
  DG_CUDA(cudaSetDevice(ctx->device));
  const size_t n = (size_t)B * Hout * Wout * C;
  resample_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const unsigned char*)in_u8, (unsigned char*)out_u8, B, Hin, Win, C,
                                                                                   Hout, Wout, bounds, coeffs, ksize, axis);
  DG_LAUNCH_CHECK();
  return DG_OK;
}
int32_t dg_op_clip_normalize(dg_ctx* ctx, const void* in_u8, void* out, int32_t B, int32_t H, int32_t W, int32_t top, int32_t left, int32_t n,
                             const float* mean3, const float* std3, void* stream) {
  if (!ctx || !in_u8 || !out || !mean3 || !std3 || B <= 0 || n <= 0 || top < 0 || left < 0 || top + n > H || left + n > W) return fail(DG_E_ARG, "bad argument");
  DG_CUDA(cudaSetDevice(ctx->device));
  const size_t cnt = (size_t)B * 3 * n * n;
  clip_normalize_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const unsigned char*)in_u8, (__half*)out, B, H, W, top, left, n,
                                                                                        mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
  DG_LAUNCH_CHECK();
  return DG_OK;
}

}  // extern "C"
