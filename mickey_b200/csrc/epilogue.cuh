// GEMM epilogues shared by the tcgen05 kernel and the SIMT debug kernel.
// Contract: the calling thread owns output row `m` (local to its group `g`) and receives the fp32
// accumulators of 32 consecutive columns n0..n0+31 in v[32] (this is the natural tcgen05.ld 32x32b
// register layout: one TMEM lane == one row per thread).
#pragma once
#include "common.cuh"

namespace mk {

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == ACT_GELU) return gelu_erf(x);
  if (act == ACT_RELU) return fmaxf(x, 0.0f);
  return x;
}

__device__ __forceinline__ void store_h32(__half* dst, const float* v) {
  // dst is 16-byte aligned by construction (ld and column offsets are multiples of 8)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    __half2 h0 = __floats2half2_rn(v[q * 8 + 0], v[q * 8 + 1]);
    __half2 h1 = __floats2half2_rn(v[q * 8 + 2], v[q * 8 + 3]);
    __half2 h2 = __floats2half2_rn(v[q * 8 + 4], v[q * 8 + 5]);
    __half2 h3 = __floats2half2_rn(v[q * 8 + 6], v[q * 8 + 7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2);
    u.w = *reinterpret_cast<uint32_t*>(&h3);
    reinterpret_cast<uint4*>(dst)[q] = u;
  }
}

__device__ __forceinline__ bool pad_valid(const GemmParams& p, int m, int& pos) {
  const int per_img = p.pad_h2 * p.pad_w2;
  pos = m % per_img;
  const int y = pos / p.pad_w2, x = pos % p.pad_w2;
  return y >= 1 && y <= p.pad_h2 - 2 && x >= 1 && x <= p.pad_w2 - 2;
}

// ---- per-chunk epilogues ---------------------------------------------------------------------------
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int g, int m, int n0, float (&v)[32]) {
  if constexpr (EPI == EPI_STORE_H) {
    if (m >= p.M) return;
    if (p.bias) {
      const float* b = p.bias + (size_t)g * p.bias_group_off + n0;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += __ldg(b + j);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
    store_h32(p.out_h + (size_t)g * p.out_h_group_off + (size_t)m * p.out_h_ld + n0, v);
  } else if constexpr (EPI == EPI_RESID_F) {
    if (m >= p.M) return;
    float* o = p.out_f + (size_t)m * p.out_f_ld + n0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 x = reinterpret_cast<float4*>(o)[q];
      const float4 gm = __ldg(reinterpret_cast<const float4*>(p.gamma + n0) + q);
      const float4 bs = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + q);
      x.x += gm.x * (v[q * 4 + 0] + bs.x);
      x.y += gm.y * (v[q * 4 + 1] + bs.y);
      x.z += gm.z * (v[q * 4 + 2] + bs.z);
      x.w += gm.w * (v[q * 4 + 3] + bs.w);
      reinterpret_cast<float4*>(o)[q] = x;
    }
  } else if constexpr (EPI == EPI_PATCH) {
    if (m >= p.M) return;
    const int img = m / p.tok_per_img, tk = m % p.tok_per_img;
    float* o = p.out_f + ((size_t)img * (p.tok_per_img + 1) + 1 + tk) * p.out_f_ld + n0;
    const float* a = p.aux + (size_t)tk * p.N + n0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(a) + q);
      float4 x;
      x.x = v[q * 4 + 0] + t.x; x.y = v[q * 4 + 1] + t.y; x.z = v[q * 4 + 2] + t.z; x.w = v[q * 4 + 3] + t.w;
      reinterpret_cast<float4*>(o)[q] = x;
    }
  } else if constexpr (EPI == EPI_CONV) {
    if (m >= p.M) return;
    int pos = 0;
    const bool valid = p.pad_h2 ? pad_valid(p, m, pos) : true;
    if (p.bias) {
      const float* b = p.bias + (size_t)g * p.bias_group_off + n0;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += __ldg(b + j);
    }
    if (p.res_h) {
      const uint4* r = reinterpret_cast<const uint4*>(p.res_h + (size_t)g * p.res_h_group_off + (size_t)m * p.res_h_ld + n0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u = r[q];
        const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f = __half22float2(h[e]);
          v[q * 8 + e * 2] += f.x;
          v[q * 8 + e * 2 + 1] += f.y;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
    if (p.aux && ((p.aux_group_mask >> g) & 1)) {
      const float* a = p.aux + (size_t)pos * p.N + n0;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += __ldg(a + j);
    }
    if (!valid) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.0f;
    }
    if (p.out_f) {
      float* o = p.out_f + (size_t)g * p.out_f_group_off + (size_t)m * p.out_f_ld + n0;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        reinterpret_cast<float4*>(o)[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
    }
    if (p.out_h) store_h32(p.out_h + (size_t)g * p.out_h_group_off + (size_t)m * p.out_h_ld + n0, v);
  } else if constexpr (EPI == EPI_STORE_F) {
    if (m >= p.M) return;
    float* o = p.out_f + (size_t)g * p.out_f_group_off + (size_t)m * p.out_f_ld + n0;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      reinterpret_cast<float4*>(o)[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
  } else if constexpr (EPI == EPI_DUAL) {
    if (m >= p.n_valid) return;
    const float L2E = 1.4426950408889634f;
    const float sh = __ldg(p.shift + g) * L2E;
    const float dust = p.dustbin ? exp2f(__ldg(p.dustbin) * L2E - sh) : 0.0f;
    const size_t gv = (size_t)g * p.n_valid;
    const float inv_r = 1.0f / (__ldg(p.rs + gv + m) + dust);
    const float s0 = __ldg(p.scr0 + gv + m);
    const size_t base = (gv + m) * (size_t)p.n_valid + n0;
    const float k2 = p.inv_temp * L2E;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = n0 + j;
      if (n < p.n_valid) {
        const float e = exp2f(fmaf(v[j], k2, -sh));
        const float sc = (e * inv_r) * (e / (__ldg(p.cs + gv + n) + dust));
        const float kp = s0 * __ldg(p.scr1 + gv + n);
        p.scores[base + j] = sc;
        p.kp_scores[base + j] = kp;
        p.final_scores[base + j] = sc * kp;
      }
    }
  }
}

// ---- whole-row epilogues (accumulated across the chunks of one tile) ------------------------------
// EPI_LSE: partial sum over this tile's valid columns of exp(S/T - shift)
__device__ __forceinline__ float lse_partial(const GemmParams& p, int g, int n0, const float (&v)[32]) {
  const float L2E = 1.4426950408889634f;
  const float sh = __ldg(p.shift + g) * L2E;
  const float k2 = p.inv_temp * L2E;
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (n0 + j < p.n_valid) s += exp2f(fmaf(v[j], k2, -sh));
  return s;
}

// EPI_LN: normalise the 128-wide row (N == 128 == tile width), then optional residual add.
//   out_f (if set): x32[m] = x32[m] + LN(acc)  (in place), fp16 copy of the sum -> out_h
//   else          : out_h = LN(acc)
__device__ __forceinline__ void ln_store_chunk(const GemmParams& p, int g, int m, int n0, float (&v)[32],
                                               float mean, float rstd) {
  if (m >= p.M) return;
  const float* gm = p.gamma + (size_t)g * p.ln_group_off + n0;
  const float* bt = p.beta + (size_t)g * p.ln_group_off + n0;
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = (v[j] - mean) * rstd * __ldg(gm + j) + __ldg(bt + j);
  if (p.out_f) {
    float* o = p.out_f + (size_t)g * p.out_f_group_off + (size_t)m * p.out_f_ld + n0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 x = reinterpret_cast<float4*>(o)[q];
      v[q * 4 + 0] += x.x; v[q * 4 + 1] += x.y; v[q * 4 + 2] += x.z; v[q * 4 + 3] += x.w;
    }
  }
  if (p.pad_h2) {
    int pos;
    if (!pad_valid(p, m, pos)) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.0f;
    }
  }
  if (p.out_f) {
    float* o = p.out_f + (size_t)g * p.out_f_group_off + (size_t)m * p.out_f_ld + n0;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      reinterpret_cast<float4*>(o)[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
  }
  store_h32(p.out_h + (size_t)g * p.out_h_group_off + (size_t)m * p.out_h_ld + n0, v);
}

}  // namespace mk
