// The C ABI (include/mickey_b200.h): handle, packed-weight registry, workspace carving and the kernel
// sequence of the three stages (extract -> match -> solve).
#include "../../include/mickey_b200.h"
#include "gemm.h"
#include "ops.h"

#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

using namespace mk;

struct Tensor { const void* ptr; int dtype; long long numel; };

struct mk_handle {
  int device;
  mk_config cfg;
  std::unordered_map<std::string, Tensor> tensors;
  int geo_h = 0, geo_w = 0;
  bool finalized = false;
  unsigned long long* seed_dev = nullptr;   // RNG state of the solver (device memory, advanced after every solve)
  long long launches = 0;
  // optional per-kernel-class timing with CUDA events on the launch stream (mk_profile_*)
  bool profiling = false;
  struct ProfRec { std::string tag; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof;
};

namespace {

struct Geo {
  int H, W, gh, gw, N, T, h2, w2, per_img, n_img;
  long long M, Mp, R;
};

Geo make_geo(int n_pairs, int H, int W) {
  Geo g;
  g.H = H; g.W = W; g.gh = H / 14; g.gw = W / 14;
  g.N = g.gh * g.gw; g.T = g.N + 1; g.h2 = g.gh + 2; g.w2 = g.gw + 2; g.per_img = g.h2 * g.w2;
  g.n_img = 2 * n_pairs;
  g.M = (long long)g.n_img * g.T; g.Mp = (long long)g.n_img * g.N; g.R = (long long)g.n_img * g.per_img;
  return g;
}

constexpr int KPAD = 640;       // 3*14*14 = 588 padded to a multiple of 64
constexpr int G = 4;            // heads: depth_head, det_offset, det_head, dsc_head

// Bump allocator over the caller's workspace; with base == nullptr it only measures.
struct Carver {
  uint8_t* base; size_t off = 0;
  explicit Carver(void* b) : base(reinterpret_cast<uint8_t*>(b)) {}
  template <typename T> T* take(size_t count) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

struct Workspace {
  // backbone
  __half* P; float* X; __half* XN; __half* QKV; __half* ATT; __half* H1;
  // heads
  __half* F; __half *T1, *S1, *O1, *T2, *S2, *O2, *T3, *S3, *CAT, *MSG, *HM, *T4k, *S4k, *T4d;
  float *X32, *QKV32, *KV, *KVP, *Y4k, *Y4d;
  // head outputs kept for the matcher
  float* score_raw; __half* DSCX; float* nrm2; float* scr_copy;
  // matcher
  float *part_row, *part_col, *lse_r, *lse_c;
  // solver
  void* samp_ws; int* idx; float* hyp_scores; float* hyp_Rt; int* status; int* best_hyp;
  size_t bytes;
};

Workspace carve(void* base, const mk_config& c, const Geo& g, int n_pairs) {
  Workspace w;
  Carver cv(base);
  const size_t D = c.embed_dim, M = g.M, R = g.R;
  const int* bd = c.block_dims;
  w.P = cv.take<__half>((size_t)g.Mp * KPAD);
  w.X = cv.take<float>(M * D);
  w.XN = cv.take<__half>(M * D);
  w.QKV = cv.take<__half>(M * 3 * D);
  w.ATT = cv.take<__half>(M * D);
  w.H1 = cv.take<__half>(M * 4 * D);
  w.F = cv.take<__half>(R * D);
  w.T1 = cv.take<__half>(R * G * bd[0]); w.S1 = cv.take<__half>(R * G * bd[0]); w.O1 = cv.take<__half>(R * G * bd[0]);
  w.T2 = cv.take<__half>(R * G * bd[1]); w.S2 = cv.take<__half>(R * G * bd[1]); w.O2 = cv.take<__half>(R * G * bd[1]);
  w.T3 = cv.take<__half>(R * G * bd[2]); w.S3 = cv.take<__half>(R * G * bd[2]);
  w.CAT = cv.take<__half>(R * G * 256); w.MSG = cv.take<__half>(R * G * 128); w.HM = cv.take<__half>(R * G * 256);
  w.T4k = cv.take<__half>(R * 3 * bd[3]); w.S4k = cv.take<__half>(R * 3 * bd[3]); w.T4d = cv.take<__half>(R * c.desc_dim);
  w.X32 = cv.take<float>(R * G * 128); w.QKV32 = cv.take<float>(R * G * 384);
  w.KV = cv.take<float>((size_t)g.n_img * G * 8 * 272);
  w.KVP = cv.take<float>((size_t)g.n_img * G * linattn_kv_chunks(g.h2, g.w2) * 8 * 272);
  w.Y4k = cv.take<float>(R * 3 * bd[3]); w.Y4d = cv.take<float>(R * c.desc_dim);
  w.score_raw = cv.take<float>((size_t)g.n_img * g.N);
  w.DSCX = cv.take<__half>((size_t)g.n_img * g.N * 384);
  w.nrm2 = cv.take<float>((size_t)g.n_img * g.N);
  w.scr_copy = cv.take<float>((size_t)g.n_img * g.N);
  const size_t npad = (size_t)ceil_div(g.N, 128) * 128;          // matcher: float2 partials [pair][slot][npad], lse vectors [pair][npad]
  w.part_row = cv.take<float>((size_t)n_pairs * (npad / 64) * npad * 2); w.part_col = cv.take<float>((size_t)n_pairs * (npad / 32) * npad * 2);
  w.lse_r = cv.take<float>((size_t)n_pairs * npad); w.lse_c = cv.take<float>((size_t)n_pairs * npad);
  const size_t streams = (size_t)n_pairs * c.it_matches;
  w.samp_ws = cv.take<uint8_t>(sampler_workspace_bytes(n_pairs, c.it_matches));
  w.idx = cv.take<int>(streams * c.num_sampled);
  w.hyp_scores = cv.take<float>(streams * c.it_ransac);
  w.hyp_Rt = cv.take<float>(streams * c.it_ransac * 12);
  w.status = cv.take<int>(SOLVER_COUNTER_BASE + n_pairs);     // status bits + the fused solver's completion counters
  w.best_hyp = cv.take<int>(n_pairs);
  w.bytes = cv.off + 256;
  return w;
}

// ---- weight lookup ---------------------------------------------------------------------------------------
struct Lookup {
  mk_handle* h; bool ok = true;
  const void* get(const std::string& name, int dtype, long long numel) {
    auto it = h->tensors.find(name);
    if (it == h->tensors.end()) { set_last_error("missing tensor '%s'", name.c_str()); ok = false; return nullptr; }
    if (it->second.dtype != dtype || it->second.numel != numel) {
      set_last_error("tensor '%s': expected dtype %d numel %lld, got dtype %d numel %lld", name.c_str(), dtype, numel,
                     it->second.dtype, it->second.numel);
      ok = false; return nullptr;
    }
    return it->second.ptr;
  }
  const float* f(const std::string& n, long long numel) { return reinterpret_cast<const float*>(get(n, 0, numel)); }
  const __half* hh(const std::string& n, long long numel) { return reinterpret_cast<const __half*>(get(n, 1, numel)); }
};

GemmParams base_params(long long M, int N, int K) {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)M; p.N = N; p.k_chunks = K / 64; p.chunks_per_tap = p.k_chunks; p.num_taps = 1; p.groups = 1;
  return p;
}

void set_conv_taps(GemmParams& p, int cin, int w2, bool three) {
  p.chunks_per_tap = cin / 64;
  if (three) {
    p.num_taps = 9;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) p.tap_shift[ky * 3 + kx] = (ky - 1) * w2 + (kx - 1);
  } else {
    p.num_taps = 1; p.tap_shift[0] = 0;
  }
  p.k_chunks = p.num_taps * p.chunks_per_tap;
}

#define MK_TRY(x) do { int rc_ = (x); if (rc_ != MK_OK) return rc_; } while (0)

struct ProfScope {
  mk_handle* h; cudaStream_t st; int slot = -1;
  ProfScope(mk_handle* h_, const char* tag, cudaStream_t st_) : h(h_), st(st_) {
    if (!h->profiling) return;
    mk_handle::ProfRec r;
    r.tag = tag;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, st);
    h->prof.push_back(r);
    slot = (int)h->prof.size() - 1;
  }
  ~ProfScope() { if (slot >= 0) cudaEventRecord(h->prof[slot].e1, st); }
};

int gemm(mk_handle* h, const char* tag, int epi, const void* a, long long a_rows, long long a_cols, const void* b,
         long long b_rows, long long b_cols, const GemmParams& p, cudaStream_t st) {
  GemmOperand A{a, a_rows, a_cols, a_cols}, B{b, b_rows, b_cols, b_cols};
  h->launches++;
  ProfScope ps(h, tag, st);
  return launch_gemm(epi, A, B, p, st);
}
#define MK_KERNEL(tag, call) do { ProfScope ps_(h, tag, st); h->launches++; MK_TRY(call); } while (0)

// ---- stage 1 ----------------------------------------------------------------------------------------------
// img_fmt 0: fp32 NCHW in [0,1] (the reference's tensors); 1: uint8 NHWC RGB as cv2 delivers it (mk_*_u8)
int run_extract(mk_handle* h, const void* images, int img_fmt, int n_pairs, int H, int W, float* kps, float* depth, float* scr,
                float* dsc, Workspace& w, cudaStream_t st) {
  const mk_config& c = h->cfg;
  const Geo g = make_geo(n_pairs, H, W);
  const int D = c.embed_dim;
  Lookup L{h};
  // -- tokens: patch embedding + cls + position embedding (dinov2.py:191-200)
  if (img_fmt == 1)
    MK_KERNEL("vit.ingest_u8", ingest_u8(reinterpret_cast<const uint8_t*>(images), w.P, g.n_img, H, W, KPAD, w.X, L.f("patch.clspos", D), D, st));
  else
    MK_KERNEL("vit.patch_gather", patch_gather(reinterpret_cast<const float*>(images), w.P, g.n_img, H, W, KPAD, w.X, L.f("patch.clspos", D), D, st));
  {
    GemmParams p = base_params(g.Mp, D, KPAD);
    p.aux = L.f("patch.posb", (long long)g.N * D); p.tok_per_img = g.N; p.out_f = w.X; p.out_f_ld = D;
    const __half* wt = L.hh("patch.w", (long long)D * KPAD);
    if (!L.ok) return MK_ERR_MISSING_TENSOR;
    MK_TRY(gemm(h, "vit.patch_embed", EPI_PATCH, w.P, g.Mp, KPAD, wt, D, KPAD, p, st));
  }
  // -- transformer blocks (layers/block.py:105-106)
  const bool fuse_ln = gemm_resid_ln_supported((int)g.M, D, D / 64) && gemm_resid_ln_supported((int)g.M, D, 4 * D / 64);
  for (int i = 0; i < c.depth; ++i) {
    const std::string b = "blk" + std::to_string(i) + ".";
    const float *ln1w = L.f(b + "ln1.w", D), *ln1b = L.f(b + "ln1.b", D), *ln2w = L.f(b + "ln2.w", D), *ln2b = L.f(b + "ln2.b", D);
    const __half *wqkv = L.hh(b + "qkv.w", 3LL * D * D), *wproj = L.hh(b + "proj.w", (long long)D * D);
    const __half *wfc1 = L.hh(b + "fc1.w", 4LL * D * D), *wfc2 = L.hh(b + "fc2.w", 4LL * D * D);
    const float *bqkv = L.f(b + "qkv.b", 3 * D), *bproj = L.f(b + "proj.b", D), *bfc1 = L.f(b + "fc1.b", 4 * D), *bfc2 = L.f(b + "fc2.b", D);
    const float *ls1 = L.f(b + "ls1", D), *ls2 = L.f(b + "ls2", D);
    if (!L.ok) return MK_ERR_MISSING_TENSOR;
    // LayerNorm rides in the epilogue of the GEMM that produces its input whenever that GEMM runs on a one-tile
    // kernel (EPI_RESID_LN: cluster of N/128 CTAs, row statistics through DSMEM): norm2 in attn.proj, the next
    // block's norm1 in mlp.fc2.  Only the first norm1 (after the patch embedding) and the final norm stay kernels.
    if (i == 0 || !fuse_ln) MK_KERNEL("vit.layernorm", layernorm(w.X, ln1w, ln1b, w.XN, (int)g.M, D, 1e-6f, 0, 0, 0, st));
    { GemmParams p = base_params(g.M, 3 * D, D); p.bias = bqkv; p.out_h = w.QKV; p.out_h_ld = 3 * D;
      MK_TRY(gemm(h, "vit.qkv", EPI_STORE_H, w.XN, g.M, D, wqkv, 3 * D, D, p, st)); }
    MK_KERNEL("vit.attention", attention_dispatch(w.QKV, w.ATT, g.n_img, g.T, D, c.heads, 0, st));
    { GemmParams p = base_params(g.M, D, D); p.bias = bproj; p.gamma = ls1; p.out_f = w.X; p.out_f_ld = D;
      if (fuse_ln) { p.aux = ln2w; p.beta = ln2b; p.out_h = w.XN; p.out_h_ld = D; p.eps = 1e-6f; }
      MK_TRY(gemm(h, "vit.proj", fuse_ln ? EPI_RESID_LN : EPI_RESID_F, w.ATT, g.M, D, wproj, D, D, p, st)); }
    if (!fuse_ln) MK_KERNEL("vit.layernorm", layernorm(w.X, ln2w, ln2b, w.XN, (int)g.M, D, 1e-6f, 0, 0, 0, st));
    { GemmParams p = base_params(g.M, 4 * D, D); p.bias = bfc1; p.act = ACT_GELU; p.out_h = w.H1; p.out_h_ld = 4 * D;
      MK_TRY(gemm(h, "vit.fc1", EPI_STORE_H, w.XN, g.M, D, wfc1, 4 * D, D, p, st)); }
    { GemmParams p = base_params(g.M, D, 4 * D); p.bias = bfc2; p.gamma = ls2; p.out_f = w.X; p.out_f_ld = D;
      const bool fuse_next = fuse_ln && i + 1 < c.depth;
      if (fuse_next) {
        const std::string nb = "blk" + std::to_string(i + 1) + ".";
        p.aux = L.f(nb + "ln1.w", D); p.beta = L.f(nb + "ln1.b", D); p.out_h = w.XN; p.out_h_ld = D; p.eps = 1e-6f;
        if (!L.ok) return MK_ERR_MISSING_TENSOR;
      }
      MK_TRY(gemm(h, "vit.fc2", fuse_next ? EPI_RESID_LN : EPI_RESID_F, w.H1, g.M, 4 * D, wfc2, D, 4 * D, p, st)); }
  }
  // -- final norm, drop cls, scatter into the zero-padded NHWC feature image (dinov2.py:230-233, mickey_extractor.py:49-51)
  MK_CUDA_CHECK(cudaMemsetAsync(w.F, 0, (size_t)g.R * D * sizeof(__half), st));
  MK_KERNEL("vit.layernorm", layernorm(w.X, L.f("norm.w", D), L.f("norm.b", D), w.F, (int)g.M, D, 1e-6f, 1, g.gh, g.gw, st));
  if (!L.ok) return MK_ERR_MISSING_TENSOR;

  // -- heads: three grouped residual blocks (extractor_utils.py:28-35; G = 4 heads side by side in channels)
  const int* bd = c.block_dims;
  struct Rb { const char* name; const __half* in; int cin; int in_goff; __half *T, *S, *O; int cout; };
  const Rb rbs[3] = {{"rb1", w.F, D, 0, w.T1, w.S1, w.O1, bd[0]},
                     {"rb2", w.O1, bd[0], bd[0], w.T2, w.S2, w.O2, bd[1]},
                     {"rb3", w.O2, bd[1], bd[1], w.T3, w.S3, nullptr, bd[2]}};
  if (bd[2] != 128) { set_last_error("KP_HEADS.BLOCKS_DIM[2] must be 128 (transformer width)"); return MK_ERR_UNSUPPORTED; }
  for (int r = 0; r < 3; ++r) {
    const Rb& rb = rbs[r];
    const std::string n = std::string(rb.name) + ".";
    const long long in_cols = (r == 0) ? D : (long long)G * rb.cin;
    const __half *wc1 = L.hh(n + "c1.w", (long long)G * rb.cout * 9 * rb.cin), *wsc = L.hh(n + "sc.w", (long long)G * rb.cout * rb.cin);
    const __half* wc2 = L.hh(n + "c2.w", (long long)G * rb.cout * 9 * rb.cout);
    const float *b1 = L.f(n + "c1.b", G * rb.cout), *b2 = L.f(n + "c2.b", G * rb.cout);
    if (!L.ok) return MK_ERR_MISSING_TENSOR;
    {  // conv1 + bn1 + relu
      GemmParams p = base_params(g.R, rb.cout, 64); set_conv_taps(p, rb.cin, g.w2, true);
      p.groups = G; p.a_col_group_off = rb.in_goff; p.b_row_group_off = rb.cout; p.bias = b1; p.bias_group_off = rb.cout;
      p.act = ACT_RELU; p.pad_h2 = g.h2; p.pad_w2 = g.w2; p.out_h = rb.T; p.out_h_ld = (long long)G * rb.cout; p.out_h_group_off = rb.cout;
      MK_TRY(gemm(h, "head.conv3x3", EPI_CONV, rb.in, g.R, in_cols, wc1, (long long)G * rb.cout, 9LL * rb.cin, p, st));
    }
    {  // 1x1 shortcut
      GemmParams p = base_params(g.R, rb.cout, 64); set_conv_taps(p, rb.cin, g.w2, false);
      p.groups = G; p.a_col_group_off = rb.in_goff; p.b_row_group_off = rb.cout;
      p.out_h = rb.S; p.out_h_ld = (long long)G * rb.cout; p.out_h_group_off = rb.cout;
      MK_TRY(gemm(h, "head.conv1x1", EPI_CONV, rb.in, g.R, in_cols, wsc, (long long)G * rb.cout, rb.cin, p, st));
    }
    {  // conv2 + bn2 + shortcut + relu (+ sine position encoding and fp32 copy after block 3)
      GemmParams p = base_params(g.R, rb.cout, 64); set_conv_taps(p, rb.cout, g.w2, true);
      p.groups = G; p.a_col_group_off = rb.cout; p.b_row_group_off = rb.cout; p.bias = b2; p.bias_group_off = rb.cout;
      p.res_h = rb.S; p.res_h_ld = (long long)G * rb.cout; p.res_h_group_off = rb.cout;
      p.act = ACT_RELU; p.pad_h2 = g.h2; p.pad_w2 = g.w2;
      if (r < 2) { p.out_h = rb.O; p.out_h_ld = (long long)G * rb.cout; p.out_h_group_off = rb.cout; }
      else {
        p.out_h = w.CAT; p.out_h_ld = G * 256; p.out_h_group_off = 256;
        p.out_f = w.X32; p.out_f_ld = G * 128; p.out_f_group_off = 128;
        p.aux = L.f("head.pe", (long long)g.per_img * 128);
        p.aux_group_mask = (c.kp_pos_enc ? 0x7 : 0) | (c.dsc_pos_enc ? 0x8 : 0);
        if (!L.ok) return MK_ERR_MISSING_TENSOR;
      }
      MK_TRY(gemm(h, "head.conv3x3", EPI_CONV, rb.T, g.R, (long long)G * rb.cout, wc2, (long long)G * rb.cout, 9LL * rb.cout, p, st));
    }
  }
  // -- linear-attention transformer, 3 layers (att_layers/transformer_utils.py:40-66)
  for (int l = 0; l < 3; ++l) {
    const std::string n = "att" + std::to_string(l) + ".";
    const __half *wqkv = L.hh(n + "qkv.w", (long long)G * 384 * 128), *wmerge = L.hh(n + "merge.w", (long long)G * 128 * 128);
    const __half *wm0 = L.hh(n + "mlp0.w", (long long)G * 256 * 256), *wm2 = L.hh(n + "mlp2.w", (long long)G * 128 * 256);
    const float *n1w = L.f(n + "n1.w", G * 128), *n1b = L.f(n + "n1.b", G * 128), *n2w = L.f(n + "n2.w", G * 128), *n2b = L.f(n + "n2.b", G * 128);
    if (!L.ok) return MK_ERR_MISSING_TENSOR;
    { GemmParams p = base_params(g.R, 384, 128); p.groups = G; p.a_col_group_off = 256; p.b_row_group_off = 384;
      p.out_f = w.QKV32; p.out_f_ld = G * 384; p.out_f_group_off = 384;
      MK_TRY(gemm(h, "head.att.qkv", EPI_STORE_F, w.CAT, g.R, G * 256, wqkv, G * 384, 128, p, st)); }
    { ProfScope ps_(h, "head.att.kv", st); h->launches += 2; MK_TRY(linattn_kv(w.QKV32, w.KVP, w.KV, g.n_img, G, g.h2, g.w2, st)); }
    MK_KERNEL("head.att.msg", linattn_msg(w.QKV32, w.KV, w.MSG, g.n_img, G, g.h2, g.w2, 1e-6f, st));
    { GemmParams p = base_params(g.R, 128, 128); p.groups = G; p.a_col_group_off = 128; p.b_row_group_off = 128;
      p.gamma = n1w; p.beta = n1b; p.ln_group_off = 128; p.eps = 1e-5f;
      p.out_h = w.CAT + 128; p.out_h_ld = G * 256; p.out_h_group_off = 256;
      MK_TRY(gemm(h, "head.att.merge_ln", EPI_LN, w.MSG, g.R, G * 128, wmerge, G * 128, 128, p, st)); }
    { GemmParams p = base_params(g.R, 256, 256); p.groups = G; p.a_col_group_off = 256; p.b_row_group_off = 256; p.act = ACT_RELU;
      p.out_h = w.HM; p.out_h_ld = G * 256; p.out_h_group_off = 256;
      MK_TRY(gemm(h, "head.att.mlp0", EPI_STORE_H, w.CAT, g.R, G * 256, wm0, G * 256, 256, p, st)); }
    { GemmParams p = base_params(g.R, 128, 256); p.groups = G; p.a_col_group_off = 256; p.b_row_group_off = 128;
      p.gamma = n2w; p.beta = n2b; p.ln_group_off = 128; p.eps = 1e-5f;
      p.out_f = w.X32; p.out_f_ld = G * 128; p.out_f_group_off = 128;
      p.out_h = w.CAT; p.out_h_ld = G * 256; p.out_h_group_off = 256;
      if (l == 2) { p.pad_h2 = g.h2; p.pad_w2 = g.w2; }       // zero the pad rows again before the next 3x3 conv
      MK_TRY(gemm(h, "head.att.mlp2_ln", EPI_LN, w.HM, g.R, G * 256, wm2, G * 128, 256, p, st)); }
  }
  // -- residual block 4: three keypoint heads (128 -> 64, with shortcut conv) and the descriptor head (128 -> desc_dim)
  {
    const int co = bd[3];
    const __half *wc1 = L.hh("rb4k.c1.w", 3LL * co * 9 * 128), *wsc = L.hh("rb4k.sc.w", 3LL * co * 128), *wc2 = L.hh("rb4k.c2.w", 3LL * co * 9 * co);
    const float *b1 = L.f("rb4k.c1.b", 3 * co), *b2 = L.f("rb4k.c2.b", 3 * co);
    if (!L.ok) return MK_ERR_MISSING_TENSOR;
    { GemmParams p = base_params(g.R, co, 64); set_conv_taps(p, 128, g.w2, true);
      p.groups = 3; p.a_col_group_off = 256; p.b_row_group_off = co; p.bias = b1; p.bias_group_off = co; p.act = ACT_RELU;
      p.pad_h2 = g.h2; p.pad_w2 = g.w2; p.out_h = w.T4k; p.out_h_ld = 3 * co; p.out_h_group_off = co;
      MK_TRY(gemm(h, "head.conv3x3", EPI_CONV, w.CAT, g.R, G * 256, wc1, 3 * co, 9 * 128, p, st)); }
    { GemmParams p = base_params(g.R, co, 64); set_conv_taps(p, 128, g.w2, false);
      p.groups = 3; p.a_col_group_off = 256; p.b_row_group_off = co; p.out_h = w.S4k; p.out_h_ld = 3 * co; p.out_h_group_off = co;
      MK_TRY(gemm(h, "head.conv1x1", EPI_CONV, w.CAT, g.R, G * 256, wsc, 3 * co, 128, p, st)); }
    { GemmParams p = base_params(g.R, co, 64); set_conv_taps(p, co, g.w2, true);
      p.groups = 3; p.a_col_group_off = co; p.b_row_group_off = co; p.bias = b2; p.bias_group_off = co; p.act = ACT_RELU;
      p.res_h = w.S4k; p.res_h_ld = 3 * co; p.res_h_group_off = co; p.pad_h2 = g.h2; p.pad_w2 = g.w2;
      p.out_f = w.Y4k; p.out_f_ld = 3 * co; p.out_f_group_off = co;
      MK_TRY(gemm(h, "head.conv3x3", EPI_CONV, w.T4k, g.R, 3 * co, wc2, 3 * co, 9 * co, p, st)); }
  }
  {
    const int co = c.desc_dim;
    const __half *wc1 = L.hh("rb4d.c1.w", (long long)co * 9 * 128), *wc2 = L.hh("rb4d.c2.w", (long long)co * 9 * co);
    const float *b1 = L.f("rb4d.c1.b", co), *b2 = L.f("rb4d.c2.b", co);
    if (co != 128) { set_last_error("DSC_HEAD.LAST_DIM must be 128"); return MK_ERR_UNSUPPORTED; }
    if (!L.ok) return MK_ERR_MISSING_TENSOR;
    { GemmParams p = base_params(g.R, co, 64); set_conv_taps(p, 128, g.w2, true);
      p.a_col_base = 3 * 256; p.bias = b1; p.act = ACT_RELU; p.pad_h2 = g.h2; p.pad_w2 = g.w2; p.out_h = w.T4d; p.out_h_ld = co;
      MK_TRY(gemm(h, "head.conv3x3", EPI_CONV, w.CAT, g.R, G * 256, wc1, co, 9 * 128, p, st)); }
    { GemmParams p = base_params(g.R, co, 64); set_conv_taps(p, co, g.w2, true);
      p.bias = b2; p.act = ACT_NONE; p.res_h = w.CAT + 3 * 256; p.res_h_ld = G * 256;   // identity shortcut (in == out planes)
      p.pad_h2 = g.h2; p.pad_w2 = g.w2; p.out_f = w.Y4d; p.out_f_ld = co;
      MK_TRY(gemm(h, "head.conv3x3", EPI_CONV, w.T4d, g.R, co, wc2, co, 9 * co, p, st)); }
  }
  // -- output layers and activations
  { ProfScope ps_(h, "head.kp_out", st); h->launches += 2;
    MK_TRY(kp_head_out(w.Y4k, L.f("out.depth.w", bd[3]), L.f("out.xy.w", 2 * bd[3]), L.f("out.score.w", bd[3]), depth, kps,
                       w.score_raw, scr, g.n_img, g.gh, g.gw, c.depth_sigmoid, c.max_depth, (float)c.down_factor, c.use_softmax, st)); }
  if (!L.ok) return MK_ERR_MISSING_TENSOR;
  if (bd[3] != 64) { set_last_error("KP_HEADS.BLOCKS_DIM[3] must be 64"); return MK_ERR_UNSUPPORTED; }
  MK_KERNEL("head.desc_out", desc_out(w.Y4d, dsc, w.DSCX, w.nrm2, g.n_img, g.gh, g.gw, c.norm_dsc, st));
  MK_CUDA_CHECK(cudaMemcpyAsync(w.scr_copy, scr, (size_t)g.n_img * g.N * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return MK_OK;
}

// ---- stage 2 ----------------------------------------------------------------------------------------------
// nn_pitch: row pitch (floats) of the three N x N outputs; N = the reference's contiguous layout.  With a pitch that is a
// multiple of 4 (16-byte rows) the outputs leave through TMA tensor stores; N = 1938 itself cannot (7752-byte rows).
int run_match(mk_handle* h, int n_pairs, int N, float* scores, float* kp_scores, float* final_scores, long long nn_pitch,
              Workspace& w, cudaStream_t st) {
  const mk_config& c = h->cfg;
  Lookup L{h};
  const float* dust = c.use_dustbin ? L.f("dustbin", 1) : nullptr;
  if (!L.ok) return MK_ERR_MISSING_TENSOR;
  if (!final_scores || ((scores == nullptr) != (kp_scores == nullptr))) {
    set_last_error("mk_match: final_scores is required; scores and kp_scores are given together or both NULL (lean mode)");
    return MK_ERR_INVALID;
  }
  if (nn_pitch <= 0) nn_pitch = N;
  if (nn_pitch < N) { set_last_error("mk_match: nn_pitch %lld < N %d", nn_pitch, N); return MK_ERR_INVALID; }
  const bool tma_ok = nn_pitch % 4 == 0 && reinterpret_cast<uintptr_t>(final_scores) % 16 == 0 &&
                      (!scores || (reinterpret_cast<uintptr_t>(scores) % 16 == 0 && reinterpret_cast<uintptr_t>(kp_scores) % 16 == 0));
  const float inv_t = 1.0f / c.temperature;
  const __half* A0 = w.DSCX;                                   // role-0 descriptors [n_pairs*N, 384]
  const __half* A1 = w.DSCX + (size_t)n_pairs * N * 384;       // role-1 descriptors
  const long long rows = (long long)n_pairs * N;
  const int npad = ceil_div(N, 128) * 128;
  auto mp = [&]() {
    GemmParams p = base_params(N, N, 384);
    p.groups = n_pairs; p.a_row_group_off = N; p.b_row_group_off = N; p.n_valid = N; p.inv_temp = inv_t; p.part_ld = npad;
    return p;
  };
  // pass 1: S once, row and column partials from the same tile; then the tiny fold (+ dustbin); pass 2: outputs
  { GemmParams p = mp(); p.part_row = reinterpret_cast<float2*>(w.part_row); p.part_col = reinterpret_cast<float2*>(w.part_col);
    // L2-normalised descriptors: |S| <= 1 (+ rounding), so 1.001 / T bounds every logit; used while 2 * 1.001 / T * log2(e)
    // stays far inside the fp32 exponent range (T >= 0.04); otherwise, and for un-normalised descriptors, true maxima
    p.lse_bound = (c.norm_dsc && inv_t <= 25.0f) ? 1.001f : 0.0f;
    MK_TRY(gemm(h, "match.lse", EPI_LSE, A0, rows, 384, A1, rows, 384, p, st)); }
  MK_KERNEL("match.reduce", matcher_lse_reduce(w.part_row, w.part_col, dust, n_pairs, N, npad, w.lse_r, w.lse_c, st));
  { GemmParams p = mp(); p.lse_r = w.lse_r; p.lse_c = w.lse_c; p.scr0 = w.scr_copy; p.scr1 = w.scr_copy + (size_t)n_pairs * N;
    p.scores = scores; p.kp_scores = kp_scores; p.final_scores = final_scores; p.out_pitch = nn_pitch; p.out_tma = tma_ok ? 1 : 0;
    MK_TRY(gemm(h, "match.dual_softmax", EPI_DUAL, A0, rows, 384, A1, rows, 384, p, st)); }
  return MK_OK;
}

// ---- stage 3 ----------------------------------------------------------------------------------------------
int run_solve(mk_handle* h, const float* final_scores, long long nn_pitch, const float* kps, const float* depth, const float* K0,
              const float* K1, int n_pairs, int N, unsigned long long seed, const int* outer_idx, const int* inner_idx,
              float* pose, int* best_set, float* inl_mask, int* sampled_out, float* hyp_scores_out, int* status_out,
              Workspace& w, cudaStream_t st) {
  const mk_config& c = h->cfg;
  RansacParams rp{c.it_matches, c.it_ransac, c.num_sampled, c.num_corr, c.num_refine, c.th_inlier, c.th_soft_inlier, h->seed_dev};
  if (nn_pitch <= 0) nn_pitch = N;
  if (seed != 0) MK_TRY(seed_set(h->seed_dev, seed, st));      // seed == 0: continue the device-side sequence
  MK_CUDA_CHECK(cudaMemsetAsync(w.status, 0, (size_t)(SOLVER_COUNTER_BASE + n_pairs) * sizeof(int), st));
  const size_t n_idx = (size_t)n_pairs * c.it_matches * c.num_sampled;
  const int* idx = outer_idx;
  if (!idx) {
    { ProfScope ps_(h, "solve.sample_outer", st); h->launches += 4;
      MK_TRY(sample_outer(final_scores, n_pairs, N, nn_pitch, c.it_matches, c.num_sampled, h->seed_dev, w.samp_ws, w.idx, w.status, st)); }
    idx = w.idx;
  }
  const float* kps0 = kps;
  const float* kps1 = kps + (size_t)n_pairs * 2 * N;
  const float* d0 = depth;
  const float* d1 = depth + (size_t)n_pairs * N;
  int* bs = best_set ? best_set : w.best_hyp;     // scratch when the caller does not want it
  { ProfScope ps_(h, "solve.ransac", st); h->launches += 1;
    MK_TRY(ransac_solve(final_scores, nn_pitch, kps0, d0, kps1, d1, K0, K1, n_pairs, N, rp, idx, inner_idx, w.hyp_scores,
                        w.hyp_Rt, w.status, pose, bs, inl_mask, best_set ? w.best_hyp : nullptr, st)); }
  MK_TRY(seed_advance(h->seed_dev, st));
  if (sampled_out) MK_CUDA_CHECK(cudaMemcpyAsync(sampled_out, idx, n_idx * sizeof(int), cudaMemcpyDeviceToDevice, st));
  if (hyp_scores_out)
    MK_CUDA_CHECK(cudaMemcpyAsync(hyp_scores_out, w.hyp_scores, (size_t)n_pairs * c.it_matches * c.it_ransac * sizeof(float),
                                  cudaMemcpyDeviceToDevice, st));
  if (status_out) MK_CUDA_CHECK(cudaMemcpyAsync(status_out, w.status, sizeof(int), cudaMemcpyDeviceToDevice, st));
  return MK_OK;
}

int check_ws(mk_handle* h, int n_pairs, int H, int W, void* ws, long long ws_bytes, Workspace& out) {
  if (!h || !h->finalized) { set_last_error("handle not finalized"); return MK_ERR_INVALID; }
  MK_CUDA_CHECK(cudaSetDevice(h->device));        // one handle per device: every entry point runs on the handle's device
  if (H != h->geo_h || W != h->geo_w) {
    set_last_error("geometry %dx%d does not match the finalized geometry %dx%d", H, W, h->geo_h, h->geo_w);
    return MK_ERR_INVALID;
  }
  const Geo g = make_geo(n_pairs, H, W);
  out = carve(ws, h->cfg, g, n_pairs);
  if (!ws || (long long)out.bytes > ws_bytes) {
    set_last_error("workspace too small: need %zu bytes, got %lld", out.bytes, ws_bytes);
    return MK_ERR_INVALID;
  }
  return MK_OK;
}

}  // namespace

// ============================================================================================================
extern "C" {

const char* mk_last_error(void) { return mk::last_error(); }
const char* mk_version(void) { return "mickey_b200 0.1.0 (sm_100a)"; }
int mk_sizeof_config(void) { return (int)sizeof(mk_config); }
int mk_sizeof_gemm_args(void) { return (int)sizeof(mk_gemm_args); }

int mk_create(int device, const mk_config* cfg, mk_handle** out) {
  if (!cfg || !out) { set_last_error("null argument"); return MK_ERR_INVALID; }
  if (cfg->embed_dim != cfg->heads * 64 || cfg->embed_dim % 128) {
    set_last_error("unsupported backbone: embed_dim %d heads %d", cfg->embed_dim, cfg->heads);
    return MK_ERR_UNSUPPORTED;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    set_last_error("CUDA device %d not available (a B200 is required; there is no CPU path)", device);
    return MK_ERR_CUDA;
  }
  cudaDeviceProp prop;
  MK_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_last_error("device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
    return MK_ERR_UNSUPPORTED;
  }
  mk_handle* h = new mk_handle();
  h->device = device;
  h->cfg = *cfg;
  MK_CUDA_CHECK(cudaSetDevice(device));
  MK_CUDA_CHECK(cudaMalloc(&h->seed_dev, sizeof(unsigned long long)));
  const unsigned long long s0 = 0x243F6A8885A308D3ull;
  MK_CUDA_CHECK(cudaMemcpy(h->seed_dev, &s0, sizeof(s0), cudaMemcpyHostToDevice));
  *out = h;
  return MK_OK;
}

int mk_destroy(mk_handle* h) {
  if (h) {
    if (h->seed_dev) cudaFree(h->seed_dev);
    for (auto& r : h->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    delete h;
  }
  return MK_OK;
}

int mk_set_tensor(mk_handle* h, const char* name, const void* ptr, int dtype, long long numel) {
  if (!h || !name || !ptr) { set_last_error("null argument"); return MK_ERR_INVALID; }
  h->tensors[name] = Tensor{ptr, dtype, numel};
  return MK_OK;
}

int mk_finalize(mk_handle* h, int H, int W) {
  if (!h) return MK_ERR_INVALID;
  if (H < 14 * 7 || W < 14 * 7) { set_last_error("image %dx%d too small (3-px border mask needs >= 7 cells)", H, W); return MK_ERR_INVALID; }
  h->geo_h = H; h->geo_w = W; h->finalized = true;
  return MK_OK;
}

long long mk_workspace_bytes(mk_handle* h, int n_pairs, int H, int W) {
  if (!h) return -1;
  const Geo g = make_geo(n_pairs, H, W);
  return (long long)carve(nullptr, h->cfg, g, n_pairs).bytes;
}

int mk_extract(mk_handle* h, const float* images, int n_pairs, int H, int W, float* kps, float* depth, float* scr,
               float* dsc, void* ws, long long ws_bytes, void* stream) {
  Workspace w;
  MK_TRY(check_ws(h, n_pairs, H, W, ws, ws_bytes, w));
  return run_extract(h, images, 0, n_pairs, H, W, kps, depth, scr, dsc, w, (cudaStream_t)stream);
}

int mk_extract_u8(mk_handle* h, const unsigned char* images, int n_pairs, int H, int W, float* kps, float* depth, float* scr,
                  float* dsc, void* ws, long long ws_bytes, void* stream) {
  Workspace w;
  MK_TRY(check_ws(h, n_pairs, H, W, ws, ws_bytes, w));
  return run_extract(h, images, 1, n_pairs, H, W, kps, depth, scr, dsc, w, (cudaStream_t)stream);
}

int mk_match(mk_handle* h, int n_pairs, float* scores, float* kp_scores, float* final_scores, long long nn_pitch, void* ws,
             long long ws_bytes, void* stream) {
  Workspace w;
  MK_TRY(check_ws(h, n_pairs, h ? h->geo_h : 0, h ? h->geo_w : 0, ws, ws_bytes, w));
  const Geo g = make_geo(n_pairs, h->geo_h, h->geo_w);
  return run_match(h, n_pairs, g.N, scores, kp_scores, final_scores, nn_pitch, w, (cudaStream_t)stream);
}

int mk_solve_pose(mk_handle* h, const float* final_scores, long long nn_pitch, const float* kps, const float* depth, const float* K0,
                  const float* K1, int n_pairs, int n_kpts, unsigned long long seed, const int* outer_idx,
                  const int* inner_idx, float* pose, int* best_set, float* inl_mask, int* sampled_out,
                  float* hyp_scores_out, int* status, void* ws, long long ws_bytes, void* stream) {
  Workspace w;
  MK_TRY(check_ws(h, n_pairs, h ? h->geo_h : 0, h ? h->geo_w : 0, ws, ws_bytes, w));
  const Geo g = make_geo(n_pairs, h->geo_h, h->geo_w);
  if (n_kpts != g.N) { set_last_error("n_kpts %d does not match the geometry (%d)", n_kpts, g.N); return MK_ERR_INVALID; }
  return run_solve(h, final_scores, nn_pitch, kps, depth, K0, K1, n_pairs, n_kpts, seed, outer_idx, inner_idx, pose, best_set,
                   inl_mask, sampled_out, hyp_scores_out, status, w, (cudaStream_t)stream);
}

static int forward_any(mk_handle* h, const void* images, int img_fmt, const float* K0, const float* K1, int n_pairs, int H, int W,
                       unsigned long long seed, float* kps, float* depth, float* scr, float* dsc, float* scores,
                       float* kp_scores, float* final_scores, long long nn_pitch, float* pose, int* best_set, float* inl_mask,
                       int* sampled_out, int* status, void* ws, long long ws_bytes, void* stream) {
  Workspace w;
  MK_TRY(check_ws(h, n_pairs, H, W, ws, ws_bytes, w));
  const Geo g = make_geo(n_pairs, H, W);
  cudaStream_t st = (cudaStream_t)stream;
  MK_TRY(run_extract(h, images, img_fmt, n_pairs, H, W, kps, depth, scr, dsc, w, st));
  MK_TRY(run_match(h, n_pairs, g.N, scores, kp_scores, final_scores, nn_pitch, w, st));
  return run_solve(h, final_scores, nn_pitch, kps, depth, K0, K1, n_pairs, g.N, seed, nullptr, nullptr, pose, best_set, inl_mask,
                   sampled_out, nullptr, status, w, st);
}

int mk_forward(mk_handle* h, const float* images, const float* K0, const float* K1, int n_pairs, int H, int W,
               unsigned long long seed, float* kps, float* depth, float* scr, float* dsc, float* scores,
               float* kp_scores, float* final_scores, long long nn_pitch, float* pose, int* best_set, float* inl_mask,
               int* sampled_out, int* status, void* ws, long long ws_bytes, void* stream) {
  return forward_any(h, images, 0, K0, K1, n_pairs, H, W, seed, kps, depth, scr, dsc, scores, kp_scores, final_scores, nn_pitch,
                     pose, best_set, inl_mask, sampled_out, status, ws, ws_bytes, stream);
}

int mk_forward_u8(mk_handle* h, const unsigned char* images, const float* K0, const float* K1, int n_pairs, int H, int W,
                  unsigned long long seed, float* kps, float* depth, float* scr, float* dsc, float* scores,
                  float* kp_scores, float* final_scores, long long nn_pitch, float* pose, int* best_set, float* inl_mask,
                  int* sampled_out, int* status, void* ws, long long ws_bytes, void* stream) {
  return forward_any(h, images, 1, K0, K1, n_pairs, H, W, seed, kps, depth, scr, dsc, scores, kp_scores, final_scores, nn_pitch,
                     pose, best_set, inl_mask, sampled_out, status, ws, ws_bytes, stream);
}

int mk_pose_to_submission(const float* pose, int n_pairs, double* out, void* stream) {
  if (!pose || !out || n_pairs < 0) { set_last_error("null argument"); return MK_ERR_INVALID; }
  return pose_to_submission(pose, n_pairs, out, (cudaStream_t)stream);
}

long long mk_launch_count(mk_handle* h) { return h ? h->launches : -1; }

int mk_set_seed(mk_handle* h, unsigned long long seed, void* stream) {
  if (!h) return MK_ERR_INVALID;
  return seed_set(h->seed_dev, seed, (cudaStream_t)stream);
}

long long mk_workspace_offset(mk_handle* h, const char* name, int n_pairs, int H, int W) {
  if (!h || !name) return -1;
  const Geo g = make_geo(n_pairs, H, W);
  uint8_t* base = reinterpret_cast<uint8_t*>(0x1000);     // fake base: only differences are used
  Workspace w = carve(base, h->cfg, g, n_pairs);
  const std::unordered_map<std::string, const void*> m = {
      {"P", w.P}, {"X", w.X}, {"XN", w.XN}, {"QKV", w.QKV}, {"ATT", w.ATT}, {"H1", w.H1}, {"F", w.F}, {"T1", w.T1},
      {"S1", w.S1}, {"O1", w.O1}, {"T2", w.T2}, {"S2", w.S2}, {"O2", w.O2}, {"T3", w.T3}, {"S3", w.S3}, {"CAT", w.CAT},
      {"MSG", w.MSG}, {"HM", w.HM}, {"T4k", w.T4k}, {"S4k", w.S4k}, {"T4d", w.T4d}, {"X32", w.X32}, {"QKV32", w.QKV32},
      {"KV", w.KV}, {"Y4k", w.Y4k}, {"Y4d", w.Y4d}, {"score_raw", w.score_raw}, {"DSCX", w.DSCX}, {"nrm2", w.nrm2},
      {"lse_r", w.lse_r}, {"lse_c", w.lse_c}, {"idx", w.idx}, {"hyp_scores", w.hyp_scores},
      {"hyp_Rt", w.hyp_Rt}};
  auto it = m.find(name);
  if (it == m.end()) { set_last_error("unknown workspace buffer '%s'", name); return -1; }
  return (long long)(reinterpret_cast<const uint8_t*>(it->second) - base);
}

int mk_profile_enable(mk_handle* h, int enable) {
  if (!h) return MK_ERR_INVALID;
  for (auto& r : h->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  h->prof.clear();
  h->profiling = enable != 0;
  return MK_OK;
}

// Synchronises the device and writes one line per kernel class: "<tag> <launch scopes> <total ms>\n".
int mk_profile_read(mk_handle* h, char* buf, int buf_bytes) {
  if (!h || !buf || buf_bytes <= 0) return MK_ERR_INVALID;
  MK_CUDA_CHECK(cudaDeviceSynchronize());
  std::vector<std::string> order;
  std::unordered_map<std::string, std::pair<int, double>> acc;
  for (auto& r : h->prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
    if (!acc.count(r.tag)) order.push_back(r.tag);
    acc[r.tag].first += 1;
    acc[r.tag].second += ms;
  }
  std::string out;
  char line[256];
  for (auto& t : order) {
    snprintf(line, sizeof(line), "%s %d %.6f\n", t.c_str(), acc[t].first, acc[t].second);
    out += line;
  }
  if ((int)out.size() + 1 > buf_bytes) { set_last_error("profile buffer too small"); return MK_ERR_INVALID; }
  memcpy(buf, out.c_str(), out.size() + 1);
  return MK_OK;
}

// ---- operator-level entry points ---------------------------------------------------------------------------------
int mk_op_gemm(const mk_gemm_args* a, void* stream) {
  if (!a) return MK_ERR_INVALID;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = a->M; p.N = a->N; p.k_chunks = a->k_chunks; p.chunks_per_tap = a->chunks_per_tap; p.num_taps = a->num_taps;
  for (int i = 0; i < 9; ++i) p.tap_shift[i] = a->tap_shift[i];
  p.groups = a->groups; p.a_row_group_off = a->a_row_group_off; p.a_col_group_off = a->a_col_group_off;
  p.a_col_base = a->a_col_base; p.b_row_group_off = a->b_row_group_off; p.act = a->act;
  p.bias = a->bias; p.bias_group_off = a->bias_group_off; p.gamma = a->gamma; p.beta = a->beta; p.ln_group_off = a->ln_group_off;
  p.out_f = a->out_f; p.out_f_ld = a->out_f_ld; p.out_f_group_off = a->out_f_group_off;
  p.out_h = (__half*)a->out_h; p.out_h_ld = a->out_h_ld; p.out_h_group_off = a->out_h_group_off;
  p.res_h = (const __half*)a->res_h; p.res_h_ld = a->res_h_ld; p.res_h_group_off = a->res_h_group_off;
  p.aux = a->aux; p.aux_group_mask = a->aux_group_mask; p.pad_h2 = a->pad_h2; p.pad_w2 = a->pad_w2; p.tok_per_img = a->tok_per_img;
  p.eps = a->eps; p.n_valid = a->n_valid; p.inv_temp = a->inv_temp; p.dustbin = a->dustbin;
  p.part_row = reinterpret_cast<float2*>(a->part_row); p.part_col = reinterpret_cast<float2*>(a->part_col); p.part_ld = a->part_ld;
  p.lse_r = a->lse_r; p.lse_c = a->lse_c; p.scr0 = a->scr0; p.scr1 = a->scr1; p.lse_bound = a->lse_bound;
  p.out_pitch = a->out_pitch > 0 ? a->out_pitch : a->n_valid;
  p.out_tma = (a->final_scores && p.out_pitch % 4 == 0 && reinterpret_cast<uintptr_t>(a->final_scores) % 16 == 0 &&
               (!a->scores || (reinterpret_cast<uintptr_t>(a->scores) % 16 == 0 && reinterpret_cast<uintptr_t>(a->kp_scores) % 16 == 0))) ? 1 : 0;
  p.scores = a->scores; p.kp_scores = a->kp_scores; p.final_scores = a->final_scores;
  GemmOperand A{a->a, a->a_rows, a->a_cols, a->a_ld}, B{a->b, a->b_rows, a->b_cols, a->b_ld};
  return launch_gemm(a->epi, A, B, p, (cudaStream_t)stream, a->impl);
}

int mk_op_patch_gather(const float* img, void* P, int n_img, int H, int W, int kpad, float* X, const float* cls_pos,
                       int D, void* stream) {
  return patch_gather(img, P, n_img, H, W, kpad, X, cls_pos, D, (cudaStream_t)stream);
}
int mk_op_ingest_u8(const unsigned char* img, void* P, int n_img, int H, int W, int kpad, float* X, const float* cls_pos,
                    int D, void* stream) {
  return ingest_u8(img, P, n_img, H, W, kpad, X, cls_pos, D, (cudaStream_t)stream);
}
int mk_op_layernorm(const float* x, const float* w, const float* b, void* out, int rows, int D, float eps, int mode,
                    int gh, int gw, void* stream) {
  return layernorm(x, w, b, out, rows, D, eps, mode, gh, gw, (cudaStream_t)stream);
}
int mk_op_attention(const void* qkv, void* out, int n_img, int T, int D, int heads, int impl, void* stream) {
  return attention_dispatch(qkv, out, n_img, T, D, heads, impl, (cudaStream_t)stream);
}
int mk_op_linattn(const float* qkv, float* kv_part, float* kv, void* msg, int n_img, int Gn, int h2, int w2, float eps,
                  void* stream) {
  MK_TRY(linattn_kv(qkv, kv_part, kv, n_img, Gn, h2, w2, (cudaStream_t)stream));
  return linattn_msg(qkv, kv, msg, n_img, Gn, h2, w2, eps, (cudaStream_t)stream);
}
int mk_op_matcher_reduce(const float* part_row, const float* part_col, const float* dustbin, int B, int N, int part_ld,
                         float* lse_r, float* lse_c, void* stream) {
  return matcher_lse_reduce(part_row, part_col, dustbin, B, N, part_ld, lse_r, lse_c, (cudaStream_t)stream);
}
long long mk_op_sample_workspace_bytes(int B, int IM) { return (long long)sampler_workspace_bytes(B, IM) + 512; }
int mk_op_sample(const float* fs, int B, int N, long long pitch, int IM, int n_sample, unsigned long long seed, void* ws,
                 long long ws_bytes, int* idx_out, int* status, void* stream) {
  if (pitch <= 0) pitch = N;
  if ((long long)sampler_workspace_bytes(B, IM) + 256 > ws_bytes) { set_last_error("sampler workspace too small"); return MK_ERR_INVALID; }
  MK_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int), (cudaStream_t)stream));
  // the seed word lives at the (256-byte aligned) end of the caller's workspace
  const size_t off = ((size_t)sampler_workspace_bytes(B, IM) + 255) & ~(size_t)255;
  unsigned long long* sd = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(ws) + off);
  MK_TRY(seed_set(sd, seed, (cudaStream_t)stream));
  return sample_outer(fs, B, N, pitch, IM, n_sample, sd, ws, idx_out, status, (cudaStream_t)stream);
}

}  // extern "C"
