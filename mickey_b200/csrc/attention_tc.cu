// Fused multi-head attention on tcgen05 / TMEM / TMA (reference layers/attention.py:49-62).
//
//   out[img, q, head] = softmax(q k^T / 8) v          head_dim 64, T tokens per image (1939 at 720x540)
//
// CTA = 128 queries of one (image, head); KV is streamed in 128-key tiles.  320 threads:
//   warps 0-7 : softmax + correction + epilogue.  Warp w owns TMEM lanes 32(w&3)..+31 (thread == query row) and the
//               key half w>>2 of every tile (64 of the 128 S columns), so a row is shared by two threads
//   warp 8    : TMA producer (Q once; K and V rings, 2 stages each, 16 KB tiles, 128-byte swizzle)
//   warp 9    : TMEM allocator + single-thread MMA issuer
// TMEM (256 columns, two CTAs per SM):
//   [0,128)   S = Q K^T, fp32             (tcgen05.mma SS, both operands K-major)
//   [128,192) P = exp2(S - m), fp16 x2    (written by the softmax threads with tcgen05.st, read as the A operand)
//   [192,256) O += P V, fp32              (tcgen05.mma TS, B = V tile as an MN-major operand)
// Softmax is "online" with lazy rescaling: the running reference max m_ref only moves when the tile max exceeds
// it by more than 8 (log2 units), so O in TMEM is rescaled rarely and P stays below 2^8 in fp16.  QK^T of tile
// j+1 is issued as soon as the softmax warps have pulled S_j into registers, so it overlaps their exp work.
//
// Why two threads per row: with head_dim 64 the kernel is bound by MUFU.EX2 (16/clk/SM: 1024 clocks per 128x128
// tile against 512 for its two MMAs).  Per tile a softmax warp alternates that MUFU phase with ~900 clocks of TMEM
// traffic, max and barrier latency, so the sub-partition's MUFU idles unless other warps are in their exp phase.
// With one thread per row (2 warps per sub-partition at 2 CTAs/SM) clock stamps showed 3200 clocks per tile pair
// (MUFU 67 % busy, the older CTA winning arbitration); splitting the columns gives 4 half-size warps per
// sub-partition, enough independent phases to keep the MUFU fed (profiles/r01_notes.md).
#include "gemm_tc.cuh"
#include "ops.h"
#include "gemm.h"
#include <cstdlib>
#include <cstring>

namespace mk {

constexpr int FA_BQ = 128, FA_BK = 128, FA_D = 64, FA_THREADS = 320, FA_KV_STAGES = 2;
constexpr int FA_SOFTMAX_WARPS = 8, FA_WARP_TMA = 8, FA_WARP_MMA = 9;
constexpr int FA_TILE_BYTES = 128 * 128;                                   // 128 rows x 64 fp16
constexpr int FA_XCHG_BYTES = 2 * 2 * 128 * 4;                              // [tile parity][key half][row] floats
constexpr int FA_SMEM = FA_TILE_BYTES * (1 + 2 * FA_KV_STAGES) + 1024 + 256 + FA_XCHG_BYTES;
constexpr uint32_t FA_COL_S = 0, FA_COL_P = 128, FA_COL_O = 192, FA_TMEM_COLS = 256;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// Two 32-column TMEM loads issued back to back with ONE wait (the single-load helper waits after each: ~250 clocks of
// exposed latency per tile and softmax warp).
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
  uint32_t r[64];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%64];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%65];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr), "r"(taddr + 32) : "memory");
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA / ALU pipes (no MUFU): x = n + f with n = round(x), f in [-0.5, 0.5]; a degree-4 polynomial for 2^f
// (|rel err| < 5e-5, far below the fp16 rounding of P) and n added into the exponent field.  Used for a fraction of
// the logits so that the MUFU, which bounds this kernel, sees fewer of them.
__device__ __forceinline__ float exp2_fma(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;                 // 1.5 * 2^23: the low mantissa bits of t hold round(x)
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.0096181291f, 0.0555041087f);
  p = fmaf(p, f, 0.2402265070f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// Packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2): one issue slot for two logits.  The softmax loop is bound by
// instruction issue (ncu: 71 % issue-active, MUFU 52 %, FMA 39 %), not by any single pipe.
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %5};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1) {
  asm("{\n\t.reg .b64 ra, rd;\n\tmov.b64 rd, {%0, %1};\n\tmov.b64 ra, {%2, %3};\n\tadd.rn.f32x2 rd, rd, ra;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1));
}
// pair * pair + broadcast, pair * broadcast + pair, pair + broadcast: the FMA-pipe exp2 on two logits per instruction
__device__ __forceinline__ void fma2_ppb(float& d0, float& d1, float a0, float a1, float b0, float b1, float c) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c));
}
__device__ __forceinline__ void fma2_pbp(float& d0, float& d1, float a0, float a1, float b, float c0, float c1) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void add2_pb(float& d0, float& d1, float a0, float a1, float b) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tadd.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b));
}
// exp2_fma on a pair of logits: the same operations in the same order (fma.rn.f32x2 / add.rn.f32x2 round each half like
// the scalar instruction), 11 issue slots per pair instead of 18.
__device__ __forceinline__ void exp2_fma_pair(float& p0, float& p1, float x0, float x1) {
  x0 = fmaxf(x0, -125.0f); x1 = fmaxf(x1, -125.0f);
  float t0, t1, u0, u1, f0, f1;
  add2_pb(t0, t1, x0, x1, 12582912.0f);
  add2_pb(u0, u1, t0, t1, -12582912.0f);
  fma2_pbp(f0, f1, u0, u1, -1.0f, x0, x1);         // x - u, exact product: identical to the subtraction
  fma2_pbp(p0, p1, f0, f1, 0.0096181291f, 0.0555041087f, 0.0555041087f);
  fma2_ppb(p0, p1, p0, p1, f0, f1, 0.2402265070f);
  fma2_ppb(p0, p1, p0, p1, f0, f1, 0.6931471806f);
  fma2_ppb(p0, p1, p0, p1, f0, f1, 1.0f);
  p0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// MN-major operand (V tile: rows = keys (K dim), 64 contiguous fp16 of head_dim (N dim) per row, 128-byte swizzle):
// canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -> 8-key groups 1024 bytes apart (SBO); n == 1.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;            // LBO: stride between 64-element N blocks (single block here)
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO: stride between groups of 8 K rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}

template <int POLY, int PK>     // every POLY-th pair of logits takes the FMA-pipe exp2 (0: none); PK >= 1: packed fp32x2 scale / row sum, 2: and polynomial
__global__ void __launch_bounds__(FA_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, __half* __restrict__ out, int T, int D, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sQ = base;
  const uint32_t sK = base + FA_TILE_BYTES;
  const uint32_t sV = sK + FA_KV_STAGES * FA_TILE_BYTES;
  const uint32_t bars = sV + FA_KV_STAGES * FA_TILE_BYTES;
  const uint32_t q_full = bars, k_full = bars + 8, k_empty = bars + 24, v_full = bars + 40, v_empty = bars + 56;
  const uint32_t s_full = bars + 72, s_empty = bars + 80, p_full = bars + 88, pv_done = bars + 96, tmem_slot = bars + 104;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* xchg = reinterpret_cast<float*>(smem_raw + (bars + 256 - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * FA_BQ, head = blockIdx.y, im = blockIdx.z;
  const int n_tiles = (T + FA_BK - 1) / FA_BK;
  const int row_base = im * T;

  if (warp == FA_WARP_TMA && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQKV) : "memory");
    mbar_init(q_full, 1);
    for (int s = 0; s < FA_KV_STAGES; ++s) {
      mbar_init(k_full + 8 * s, 1); mbar_init(k_empty + 8 * s, 1);
      mbar_init(v_full + 8 * s, 1); mbar_init(v_empty + 8 * s, 1);
    }
    mbar_init(s_full, 1); mbar_init(s_empty, FA_SOFTMAX_WARPS); mbar_init(p_full, FA_SOFTMAX_WARPS); mbar_init(pv_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == FA_WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(FA_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  pdl_wait();

  if (warp == FA_WARP_TMA) {
    // ===== TMA producer =====
    if (elect_one()) {
      mbar_expect_tx(q_full, FA_TILE_BYTES);
      tma_load_2d(sQ, &tmQKV, q_full, head * FA_D, row_base + q0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % FA_KV_STAGES;
        const uint32_t par = ((j / FA_KV_STAGES) & 1) ^ 1;
        mbar_wait(k_empty + 8 * st, par);
        mbar_expect_tx(k_full + 8 * st, FA_TILE_BYTES);
        tma_load_2d(sK + st * FA_TILE_BYTES, &tmQKV, k_full + 8 * st, D + head * FA_D, row_base + j * FA_BK);
        mbar_wait(v_empty + 8 * st, par);
        mbar_expect_tx(v_full + 8 * st, FA_TILE_BYTES);
        tma_load_2d(sV + st * FA_TILE_BYTES, &tmQKV, v_full + 8 * st, 2 * D + head * FA_D, row_base + j * FA_BK);
      }
    }
  } else if (warp == FA_WARP_MMA) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc_qk = (1u << 4) | ((uint32_t)(FA_BK >> 3) << 17) | ((uint32_t)(FA_BQ >> 4) << 24);
    constexpr uint32_t idesc_pv = (1u << 4) | (1u << 16) | ((uint32_t)(FA_D >> 3) << 17) | ((uint32_t)(FA_BQ >> 4) << 24);
    auto issue_qk = [&](int j) {
      const int st = j % FA_KV_STAGES;
      const uint64_t da = umma_desc_sw128(sQ), db = umma_desc_sw128(sK + st * FA_TILE_BYTES);
#pragma unroll
      for (int k = 0; k < FA_D / UMMA_K; ++k)
        umma_f16(tmem_base + FA_COL_S, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_qk, k > 0 ? 1u : 0u);
      umma_commit(s_full);
      umma_commit(k_empty + 8 * st);
    };
    mbar_wait(q_full, 0);
    mbar_wait(k_full, 0);
    tc_fence_after();
    if (elect_one()) issue_qk(0);
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      if (j + 1 < n_tiles) {
        const int st1 = (j + 1) % FA_KV_STAGES;
        mbar_wait(k_full + 8 * st1, ((j + 1) / FA_KV_STAGES) & 1);
        mbar_wait(s_empty, j & 1);                 // softmax warps hold S_j in registers
        tc_fence_after();
        if (elect_one()) issue_qk(j + 1);
        __syncwarp();
      }
      const int st = j % FA_KV_STAGES;
      mbar_wait(v_full + 8 * st, (j / FA_KV_STAGES) & 1);
      mbar_wait(p_full, j & 1);                    // P_j written (and O rescaled if needed)
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dv = umma_desc_mn_sw128(sV + st * FA_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < FA_BK / UMMA_K; ++k)   // 16 keys per MMA: 8 TMEM columns of P, 16 rows (2048 B) of V
          umma_f16_ts(tmem_base + FA_COL_O, tmem_base + FA_COL_P + k * 8, dv + (uint64_t)(k * 128), idesc_pv,
                      (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(pv_done);
        umma_commit(v_empty + 8 * st);
      }
      __syncwarp();
    }
  } else {
    // ===== softmax / correction / epilogue warps =====
    const int quad = warp & 3, half = warp >> 2;            // TMEM lane quadrant, key half of every tile
    const int row = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + FA_COL_S + 64 * half;
    const uint32_t tP = tmem_base + lane_off + FA_COL_P + 32 * half;
    const uint32_t tO = tmem_base + lane_off + FA_COL_O + 32 * half;
    const uint32_t pair_bar = 1 + quad;                     // named barrier of the two warps that share these rows
    float m_ref = -INFINITY, l_run = 0.f;
    const bool rows_dead = (q0 + quad * 32 >= T);           // every row of this warp lies beyond the sequence (last query tile)
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      float s[64];
      tmem_ld64(tS, s);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);         // S buffer may be overwritten by QK^T of the next tile
      const int valid = T - j * FA_BK - 64 * half; // keys of this half tile that exist (may be <= 0 in the last tile)
      if (valid < 64) {                            // warp-uniform branch
#pragma unroll
        for (int i = 0; i < 64; ++i) s[i] = (i < valid) ? s[i] : -INFINITY;
      }
      // max of the RAW logits (scale_log2 > 0 commutes with max); 8 independent chains instead of one 64-long one
      float mxa[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) mxa[k] = s[k];
#pragma unroll
      for (int i = 8; i < 64; ++i) mxa[i & 7] = fmaxf(mxa[i & 7], s[i]);
      float mx = fmaxf(fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])), fmaxf(fmaxf(mxa[4], mxa[5]), fmaxf(mxa[6], mxa[7])));
      // row max over both key halves: exchange through shared memory (double-buffered by tile parity, so the next
      // tile's write cannot overtake the partner's read), one 64-thread named barrier per tile
      float* xm = xchg + (j & 1) * 256;
      xm[half * 128 + row] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      mx = fmaxf(mx, xm[(half ^ 1) * 128 + row]) * scale_log2;   // finite: the tile holds at least one key
      float alpha = 1.0f;
      const bool move = (mx > m_ref + 8.0f);       // always true for j == 0 (m_ref = -inf); identical in both halves
      if (move) { alpha = ex2_approx(m_ref - mx); m_ref = mx; }
      // p = 2^(s*scale - m_ref): one FFMA + one MUFU.EX2 per element, 4 partial sums for ILP.
      // T = 1939 = 15 * 128 + 19: the last key tile holds 19 keys and the last query tile 19 rows.  Groups of 16 logits
      // that are entirely masked (or whose rows are all beyond T) skip the exponentials: their P is exactly 0 either way
      // (warp-uniform branches; static register indexing is kept).
      const int live_groups = rows_dead ? 0 : (valid >= 64 ? 4 : (valid <= 0 ? 0 : (valid + 15) >> 4));
      float sa[4] = {0.f, 0.f, 0.f, 0.f};
      const float neg_m = -m_ref;
      uint32_t pk[32];
      auto exp_pair = [&](int i) {                 // i is a compile-time constant after unrolling
        float x0, x1;
        if (PK) fma2(x0, x1, s[2 * i], s[2 * i + 1], scale_log2, neg_m);
        else { x0 = fmaf(s[2 * i], scale_log2, neg_m); x1 = fmaf(s[2 * i + 1], scale_log2, neg_m); }
        const bool poly = (POLY > 0) && (i % (POLY > 0 ? POLY : 1) == (POLY > 0 ? POLY : 1) - 1);
        float p0, p1;
        if (poly && PK == 2) exp2_fma_pair(p0, p1, x0, x1);
        else { p0 = poly ? exp2_fma(x0) : ex2_approx(x0); p1 = poly ? exp2_fma(x1) : ex2_approx(x1); }
        if (PK) { if (i & 1) add2(sa[2], sa[3], p0, p1); else add2(sa[0], sa[1], p0, p1); }
        else { sa[(2 * i) & 3] += p0; sa[(2 * i + 1) & 3] += p1; }
        __half2 h = __floats2half2_rn(p0, p1);
        pk[i] = *reinterpret_cast<uint32_t*>(&h);
      };
      if (live_groups == 4) {                      // every tile but the last one: one straight-line block of 32 pairs
#pragma unroll
        for (int i = 0; i < 32; ++i) exp_pair(i);
      } else {
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          if (g8 < live_groups) {
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) exp_pair(g8 * 8 + ii);
          } else {
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) pk[g8 * 8 + ii] = 0u;
          }
        }
      }
      l_run = l_run * alpha + ((sa[0] + sa[1]) + (sa[2] + sa[3]));     // this thread's key half only
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);           // P buffer free, O quiescent
        tc_fence_after();
        if (__any_sync(0xffffffffu, move)) {       // rescale this warp's 32 rows x 32 columns of O (alpha == 1 where unchanged)
          float o[32];
          tmem_ld32(tO, o);
          uint32_t ob[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) ob[i] = __float_as_uint(o[i] * alpha);
          tmem_st32(tO, ob);
        }
      }
      tmem_st32(tP, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue: O / (l of both key halves) -> fp16; this warp stores 32 of the head's 64 columns
    float* xl = xchg + (n_tiles & 1) * 256;        // the parity the last tile did not use
    xl[half * 128 + row] = l_run;
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    const float inv = 1.0f / (l_run + xl[(half ^ 1) * 128 + row]);
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    pdl_trigger();
    const int q = q0 + row;
    __half* dst = out + ((long long)(row_base + q)) * D + head * FA_D + 32 * half;
    {
      float o[32];
      tmem_ld32(tO, o);
      if (q < T) {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] *= inv;
        store_h32(dst, o);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == FA_WARP_MMA) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(FA_TMEM_COLS) : "memory");
  }
}

int attention_tc(const void* qkv, void* out, int n_img, int T, int D, int heads, cudaStream_t s) {
  if (D != heads * FA_D) { set_last_error("attention: head_dim must be 64 (D=%d heads=%d)", D, heads); return MK_ERR_UNSUPPORTED; }
  static unsigned long long attr_mask = 0;
  static int poly = 0, pack = 0;
  if (first_use_on_device(attr_mask)) {
    const void* all[] = {(const void*)attention_tc_kernel<0, 0>, (const void*)attention_tc_kernel<8, 0>, (const void*)attention_tc_kernel<4, 0>,
                         (const void*)attention_tc_kernel<0, 1>, (const void*)attention_tc_kernel<8, 1>, (const void*)attention_tc_kernel<4, 1>,
                         (const void*)attention_tc_kernel<8, 2>, (const void*)attention_tc_kernel<4, 2>, (const void*)attention_tc_kernel<3, 2>};
    for (const void* f : all) MK_CUDA_CHECK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    // packed fp32x2 scale / row sum (FFMA2 / FADD2): on by default (64 images, ViT-B: 1113 -> 1081 us with every 4th pair on
    // the FMA-pipe exp2, 1058 us with every 8th).  Every 4th stays the default: the `scores` parity metric, which amplifies
    // the features' error through logits of +-10, was consistently better with it (ViT-L 720x540: 7.9e-4 vs 1.03e-3 for
    // every 8th and 8.6e-4 for none; profiles/r02_notes.md).  MICKEY_ATTN_PACK2=0|1 / MICKEY_ATTN_POLY=0|8|4|3 override.
    // MICKEY_ATTN_PACK2=2 (default) also evaluates the polynomial on packed pairs: 61 fewer instructions per 64 logits, the
    // same results bit for bit, and the same time (64 images: 1073 vs 1074 us, session 14) -- with every 3rd pair on the
    // polynomial 1052 us, every 2nd 1178 us, every 8th 1117 us: neither the issue slots nor the FMA pipe alone bound the loop.
    // Every 3rd pair (MICKEY_ATTN_POLY=3, opt-in): 1075 -> 1053 us and C3 630 -> 641 pairs/s on one box (session 22), all parity
    // tests green, but `scores` of the ViT-L 720x540 fixture moves from 7.9e-4 to 9.9e-4 of the 1e-3 north-star bound: not taken.
    { const char* e = getenv("MICKEY_ATTN_PACK2"); pack = e ? atoi(e) : 2; if (pack < 0 || pack > 2) pack = 2; }
    // default: every 4th pair (25 %) on the FMA pipe -- measured 32.0 -> 29.9 us (one 720x540 pair) and 1156 -> 1100 us
    // (64 images, ViT-B); 12.5 % gives half of that, 50 % is slower than none (issue-bound).  MICKEY_ATTN_POLY=0 disables.
    const char* e = getenv("MICKEY_ATTN_POLY"); poly = e ? atoi(e) : 4;
  }
  CUtensorMap tm;
  int rc = make_tensor_map_f16(&tm, qkv, (long long)n_img * T, 3LL * D, 3LL * D, FA_BQ);
  if (rc) return rc;
  dim3 grid(ceil_div(T, FA_BQ), heads, n_img);
  const float scale_log2 = 0.125f * 1.4426950408889634f;
  auto kern = pack == 2 ? (poly == 8 ? attention_tc_kernel<8, 2> : poly == 4 ? attention_tc_kernel<4, 2> : poly == 3 ? attention_tc_kernel<3, 2> : attention_tc_kernel<0, 1>)
            : pack == 1 ? (poly == 8 ? attention_tc_kernel<8, 1> : poly == 4 ? attention_tc_kernel<4, 1> : attention_tc_kernel<0, 1>)
                        : (poly == 8 ? attention_tc_kernel<8, 0> : poly == 4 ? attention_tc_kernel<4, 0> : attention_tc_kernel<0, 0>);
  MK_CUDA_CHECK(launch_k(kern, grid, dim3(FA_THREADS), (size_t)FA_SMEM, s, tm, (__half*)out, T, D, scale_log2));
  return MK_OK;
}

// A ping-pong variant (one CTA per SM, two Q tiles, the two softmax warps of a sub-partition alternating their exp loops
// through named barriers) was written at the end of round 1 and measured in round 2: correct, but SLOWER than the kernel
// above on B200 (one pair: 36.3 vs 31.8 us; 64 images: 1391 vs 1100 us, 7.5 vs 9.5 exp2/clk/SM) -- the strict
// alternation serialises the two groups' non-exp phases instead of hiding them.  It was removed (profiles/r02_notes.md).

// impl: 0 = default (tcgen05 unless MICKEY_ATTN_IMPL=mma), 1 = tcgen05, 2 = mma.sync
int attention_dispatch(const void* qkv, void* out, int n_img, int T, int D, int heads, int impl, cudaStream_t s) {
  if (impl == 0) {
    static int def = 0;
    if (!def) { const char* e = getenv("MICKEY_ATTN_IMPL"); def = (e && strcmp(e, "mma") == 0) ? 2 : 1; }
    impl = def;
  }
  if (impl != 1 && impl != 2) { set_last_error("attention: unknown impl %d", impl); return MK_ERR_INVALID; }
  return impl == 2 ? attention(qkv, out, n_img, T, D, heads, s) : attention_tc(qkv, out, n_img, T, D, heads, s);
}

}  // namespace mk
