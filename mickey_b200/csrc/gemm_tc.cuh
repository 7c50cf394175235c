// tcgen05 / TMEM / TMA GEMM for sm_100a.
//
//   D[M,N] (fp32 in TMEM) = A[M,K] * B[N,K]^T,   A and B fp16, K-major, 128-byte swizzled tiles
//
// CTA = one 128 x BN output tile.  Warp roles (256 threads):
//   warp 0  : TMA producer   (one elected lane issues cp.async.bulk.tensor.2d into the smem ring)
//   warp 1  : MMA issuer     (one elected lane issues tcgen05.mma, commits to the ring's empty barriers)
//   warp 2  : TMEM allocator (tcgen05.alloc / dealloc of BN fp32 columns)
//   warp 3  : idle
//   warps 4-7: epilogue      (tcgen05.ld 32x32b: thread == output row, 32 columns per load)
// Two CTAs are resident per SM (3-stage ring, <= 128 TMEM columns each) so one CTA's epilogue
// overlaps the other's main loop.
//
// The K loop walks `k_chunks` 64-element chunks.  For convolutions a chunk also selects a filter tap:
// the A tile of tap t is the same 2-D tensor read at row offset tap_shift[t] (negative / overflowing
// rows are zero-filled by TMA), which turns a 3x3 convolution over a zero-padded NHWC image into 9
// accumulated GEMMs without materialising im2col.
#pragma once
#include "common.cuh"
#include "epilogue.cuh"

namespace mk {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;        // 64 fp16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 256;

// STAGES = 3: two CTAs resident per SM (large grids: one CTA's epilogue overlaps the other's main loop)
// STAGES = 6: one CTA per SM with twice the bytes in flight (grids of <= ~1 CTA per SM, where the 3-deep ring was
//             latency-bound: ViT proj / fc2 at B=1 moved 64 GB/s per SM)
template <int BN, int STAGES>
constexpr int gemm_smem_bytes() {
  // the ring is reused as the epilogue's staging area (8 warps x 32 rows x (BN/2 + 4) floats), which a 2-stage ring
  // does not cover
  // (EPI_DUAL's TMA path: the eight warps' 3 x 4 KB output boxes are exactly the 96 KB of a 3-stage 128 x 128 ring; their column
  // operands (8 x 64 floats) use the 2 KB statistics block behind the barriers)
  constexpr int ring = STAGES * (BLOCK_M * BLOCK_K * 2 + BN * BLOCK_K * 2), staging = 8 * 32 * (BN / 2 + 4) * 4;
  // + row statistics of EPI_RESID_LN (never run on the 2-stage ring, whose three CTAs per SM have no room to spare)
  return (ring > staging ? ring : staging) + 1024 /*align slack*/ + 256 /*barriers*/ + (STAGES == 2 ? 0 : 2048);
}

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xFFFFFFFF;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}

// UMMA shared-memory descriptor: K-major operand, 128-byte swizzle, rows of 128 bytes, 8-row groups
// 1024 bytes apart (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout SWIZZLE_128B=2 [61,64)).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;            // LBO (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO = 1024 bytes
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16: D fp32, A/B fp16, both K-major, M=128, N=BN
// (cute::UMMA::InstrDescriptor: c_format [4,6)=1, a/b_format=0, n_dim=N>>3 [17,23), m_dim=M>>4 [24,29)).
template <int BN>
__device__ __forceinline__ constexpr uint32_t umma_idesc_f16() {
  return (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

// ---- thread-block cluster helpers (EPI_RESID_LN) -------------------------------------------------------
__device__ __forceinline__ void cluster_sync_all() {     // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(cta_rank));
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
  return v;
}

// ---- tile epilogue -----------------------------------------------------------------------------------
// One warp's share of a 128 x BN accumulator tile: TMEM lanes 32q..32q+31 (its quadrant q) and half of the tile's
// 32-column chunks (warps q and q+4 split the columns).  `stage` is the warp's private 32 x (BN/2 + 4) fp32
// staging block; `release()` is called as soon as the accumulator has been read for the last time (the persistent
// kernel hands the TMEM buffer back to the MMA warp there, before the global stores).
// Per-warp staging of EPI_DUAL's TMA path: three 32 x 32 fp32 boxes (1024-byte aligned) + 64 floats of column operands
constexpr int DUAL_STAGE_BYTES = 3 * 4096;
constexpr int DUAL_AUX_BYTES = 8 * 64 * 4;

template <int BN, int EPI, int PWMAX = 64, int NPARTS = 2, typename Release>
__device__ __forceinline__ void tile_epilogue(const GemmParams& p, const OutMaps& om, int g, int m0, int n0, int n_tile, int q, int half,
                                              int lane, uint32_t tmem_acc, float* stage, Release release,
                                              float* red = nullptr, uint32_t red_saddr = 0, float* dual_stage = nullptr,
                                              float* dual_aux = nullptr) {
  // NPARTS warps share a TMEM lane quadrant and split the tile's columns (`half` = this warp's part, 0..NPARTS-1)
  static_assert(NPARTS == 2 || (NPARTS == 4 && BN == 256 && (EPI == EPI_STORE_H || EPI == EPI_RESID_F || EPI == EPI_STORE_F)),
                "4 column parts: 256-wide tiles, plain store / residual epilogues");
  constexpr int CHUNKS = BN / 32;
  constexpr int CPH = (CHUNKS + NPARTS - 1) / NPARTS;   // chunks per part
  constexpr int W = BN / NPARTS;                        // columns owned by this warp
  const int c_begin = half * CPH;
  const int c_end = (c_begin + CPH < CHUNKS) ? c_begin + CPH : CHUNKS;
  const int m = m0 + q * 32 + lane;
  const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
  float v[32];
  if constexpr (EPI == EPI_LN) {
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      tmem_ld32(taddr + c * 32, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) sum += v[j];
    }
    const float mean = sum * (1.0f / BN);
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      tmem_ld32(taddr + c * 32, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) { const float d = v[j] - mean; sq += d * d; }
    }
    const float rstd = rsqrtf(sq * (1.0f / BN) + p.eps);
    for (int c = c_begin; c < c_end; ++c) {
      tmem_ld32(taddr + c * 32, v);
      float4* dst = reinterpret_cast<float4*>(stage + lane * (W + 4) + (c - c_begin) * 32);
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) dst[k4] = make_float4(v[k4 * 4], v[k4 * 4 + 1], v[k4 * 4 + 2], v[k4 * 4 + 3]);
    }
    release();
    __syncwarp();
    epilogue_rows<EPI_LN, W>(p, g, m0 + q * 32, lane, n0 + half * W, stage, mean, rstd);
  } else if constexpr (EPI == EPI_RESID_LN) {
    // `red` = this CTA's float[2 statistics][2 column halves][128 rows]; the CTAs of the cluster hold the other
    // 128-column tiles of the same rows
    static_assert(BN == 128, "EPI_RESID_LN: 128-wide tiles");
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      tmem_ld32(taddr + (c_begin + cc) * 32, v);
      float4* dst = reinterpret_cast<float4*>(stage + lane * (64 + 4) + cc * 32);
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) dst[k4] = make_float4(v[k4 * 4], v[k4 * 4 + 1], v[k4 * 4 + 2], v[k4 * 4 + 3]);
    }
    release();
    __syncwarp();
    const int row0 = m0 + q * 32, col_base = n0 + half * 64, slot = half * 128 + q * 32 + lane;
    const uint32_t n_cta = cluster_size();
    const float inv_n = 1.0f / (float)p.N;
    red[slot] = resid_ln_pass1(p, row0, lane, col_base, stage);
    cluster_sync_all();
    float tot = 0.f;
    for (uint32_t c = 0; c < n_cta; ++c)
      tot += ld_dsmem_f32(red_saddr + (q * 32 + lane) * 4, c) + ld_dsmem_f32(red_saddr + (128 + q * 32 + lane) * 4, c);
    const float mean = tot * inv_n;
    red[256 + slot] = resid_ln_pass2(lane, stage, mean);
    cluster_sync_all();
    float tot2 = 0.f;
    for (uint32_t c = 0; c < n_cta; ++c)
      tot2 += ld_dsmem_f32(red_saddr + (256 + q * 32 + lane) * 4, c) + ld_dsmem_f32(red_saddr + (384 + q * 32 + lane) * 4, c);
    const float rstd = rsqrtf(tot2 * inv_n + p.eps);
    resid_ln_pass3(p, row0, lane, col_base, stage, mean, rstd);
    cluster_sync_all();                  // no CTA may exit while a peer can still read its statistics
  } else if constexpr (EPI == EPI_LSE) {
    static_assert(BN == 128, "matcher epilogues: 128-wide tiles");
    const bool row_ok = m < p.n_valid;
    float rmax = MK_NEG_INF, rsum = 0.f;
    if (p.lse_bound > 0.f) {
      rmax = p.lse_bound * p.inv_temp * 1.4426950408889634f;
      for (int c = c_begin; c < c_end; ++c) {
        if (n0 + c * 32 < p.n_valid) {
          tmem_ld32(taddr + c * 32, v);
          lse_chunk_bounded(p, g, n0 + c * 32, lane, (m0 / BLOCK_M) * 4 + q, row_ok, v, rsum);
        }
      }
    } else {
      for (int c = c_begin; c < c_end; ++c) {
        if (n0 + c * 32 < p.n_valid) {
          tmem_ld32(taddr + c * 32, v);
          lse_chunk(p, g, n0 + c * 32, lane, (m0 / BLOCK_M) * 4 + q, row_ok, v, rmax, rsum);
        }
      }
    }
    release();
    if (row_ok) p.part_row[((size_t)g * (p.part_ld / 64) + n_tile * 2 + half) * p.part_ld + m] = make_float2(rmax, rsum);
  } else if constexpr (EPI == EPI_DUAL) {
    static_assert(BN == 128, "matcher epilogues: 128-wide tiles");
    // per-row operands: lse of the row (+inf for rows beyond the valid range -> score 0) and its keypoint score
    const bool row_ok = m < p.n_valid;
    const float lr = row_ok ? __ldg(p.lse_r + (size_t)g * p.part_ld + m) : -MK_NEG_INF;
    const float s0 = row_ok ? __ldg(p.scr0 + (size_t)g * p.n_valid + m) : 0.0f;
    if (p.out_tma) {
      float lc_pre[CPH], s1_pre[CPH];
#pragma unroll
      for (int ci = 0; ci < CPH; ++ci) {
        const int col = n0 + (c_begin + ci) * 32 + lane;
        const bool ok = col < p.n_valid;
        lc_pre[ci] = ok ? __ldg(p.lse_c + (size_t)g * p.part_ld + col) : -MK_NEG_INF;
        s1_pre[ci] = ok ? __ldg(p.scr1 + (size_t)g * p.n_valid + col) : 0.0f;
      }
#pragma unroll
      for (int ci = 0; ci < CPH; ++ci) {
        const int c = c_begin + ci;
        if (c < c_end && n0 + c * 32 < p.n_valid && m0 + q * 32 < p.n_valid) {
          tmem_ld32(taddr + c * 32, v);
          dual_store_chunk_tma(p, om, g, m0 + q * 32, lane, n0 + c * 32, v, dual_stage, dual_aux, lr, s0, lc_pre[ci], s1_pre[ci]);
        }
      }
    } else {
      for (int c = c_begin; c < c_end; ++c) {
        if (n0 + c * 32 < p.n_valid) {
          tmem_ld32(taddr + c * 32, v);
          dual_store_chunk(p, g, m0 + q * 32, lane, n0 + c * 32, v, stage, lr, s0);
        }
      }
    }
    release();
  } else {
    // the warp's columns are processed in passes of at most PWMAX (64: two 32-column TMEM loads) through the staging block
    constexpr int PW = (W > PWMAX) ? PWMAX : W;           // staged columns per pass
    constexpr int PCH = PW / 32;                          // chunks per pass
    for (int c0 = c_begin; c0 < c_end; c0 += PCH) {
#pragma unroll
      for (int cc = 0; cc < PCH; ++cc) {
        tmem_ld32(taddr + (c0 + cc) * 32, v);
        float4* dst = reinterpret_cast<float4*>(stage + lane * (PW + 4) + cc * 32);
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) dst[k4] = make_float4(v[k4 * 4], v[k4 * 4 + 1], v[k4 * 4 + 2], v[k4 * 4 + 3]);
      }
      if (c0 + PCH >= c_end) release();                   // accumulator fully read
      __syncwarp();
      epilogue_rows<EPI, PW>(p, g, m0 + q * 32, lane, n0 + c0 * 32, stage, 0.f, 0.f);
      __syncwarp();
    }
  }
}

// ---- kernel ------------------------------------------------------------------------------------------
template <int BN, int EPI, int GEMM_STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, (GEMM_STAGES <= 2) ? 3 : (GEMM_STAGES <= 3) ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p,
               const __grid_constant__ OutMaps om) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  constexpr int B_BYTES = BN * BLOCK_K * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;

  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // swizzle-128B tiles need 1024-byte alignment
  const uint32_t bar_base = base + gemm_smem_bytes<BN, GEMM_STAGES>() - 1024 - 256 - (GEMM_STAGES == 2 ? 0 : 2048);   // full[S], empty[S], tmem_full, tmem_ptr (behind ring / staging)
  const uint32_t full_bar0 = bar_base;
  const uint32_t empty_bar0 = bar_base + 8 * GEMM_STAGES;
  const uint32_t tmem_full_bar = bar_base + 16 * GEMM_STAGES;
  const uint32_t tmem_ptr_addr = bar_base + 16 * GEMM_STAGES + 8;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.z;
  const int m0 = blockIdx.x * BLOCK_M;
  const int n0 = blockIdx.y * BN;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(full_bar0 + 8 * s, 1);
      mbar_init(empty_bar0 + 8 * s, 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  pdl_wait();                    // everything above touched no global memory; operands of the previous kernel are now visible

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int a_col0 = p.a_col_base + g * p.a_col_group_off;
      const int a_row0 = m0 + g * p.a_row_group_off;
      const int b_row0 = n0 + g * p.b_row_group_off;
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        const int tap = kc / p.chunks_per_tap;
        const int kin = kc - tap * p.chunks_per_tap;
        mbar_wait(empty_bar0 + 8 * stage, phase ^ 1);
        const uint32_t sa = base + stage * STAGE_BYTES;
        const uint32_t sb = sa + A_BYTES;
        const uint32_t fb = full_bar0 + 8 * stage;
        mbar_expect_tx(fb, STAGE_BYTES);
        tma_load_2d(sa, &tmA, fb, a_col0 + kin * BLOCK_K, a_row0 + p.tap_shift[tap]);
        tma_load_2d(sb, &tmB, fb, kc * BLOCK_K, b_row0);
        if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    int stage = 0;
    uint32_t phase = 0;
    constexpr uint32_t idesc = umma_idesc_f16<BN>();
    for (int kc = 0; kc < p.k_chunks; ++kc) {
      mbar_wait(full_bar0 + 8 * stage, phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = base + stage * STAGE_BYTES;
        const uint32_t sb = sa + A_BYTES;
        const uint64_t da = umma_desc_sw128(sa);
        const uint64_t db = umma_desc_sw128(sb);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // advance 32 bytes (16 fp16) along K inside the 128-byte swizzle row: +2 in the (addr>>4) field
          umma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kc > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar0 + 8 * stage);                   // frees this smem stage when the MMAs retire
        if (kc == p.k_chunks - 1) umma_commit(tmem_full_bar);  // accumulator complete
      }
      __syncwarp();
      if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
    }
  }
  __syncwarp();

  // ===== epilogue: TMEM -> registers -> smem -> global, by ALL 8 warps =====
  // The producer / MMA warps join once their loops have drained.
  mbar_wait(tmem_full_bar, 0);
  tc_fence_after();
  pdl_trigger();                 // main loop done: the next kernel's CTAs may start their prologue under our epilogue
  // EPI_DUAL / TMA: the (idle) ring holds the warps' output boxes (1024-byte aligned), the column operands follow them
  tile_epilogue<BN, EPI>(p, om, g, m0, n0, (int)blockIdx.y, warp & 3, warp >> 2, lane, tmem_base,
                         reinterpret_cast<float*>(smem_raw + (base - raw)) + warp * (32 * (BN / 2 + 4)), [] {},
                         reinterpret_cast<float*>(smem_raw + (bar_base + 256 - raw)), bar_base + 256,
                         reinterpret_cast<float*>(smem_raw + (base - raw) + warp * DUAL_STAGE_BYTES),
                         reinterpret_cast<float*>(smem_raw + (bar_base + 256 - raw)) + warp * 64);
  if constexpr (EPI == EPI_DUAL) { if (p.out_tma && lane == 0) tma_store_wait_all(); }   // smem must outlive the bulk reads

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

// ---- persistent kernel ---------------------------------------------------------------------------------
// For grids with more tiles than SMs: one CTA per SM walks tiles t = blockIdx.x, +gridDim.x, ... (n fastest, so
// neighbouring CTAs share the A tile in L2).  The smem ring never drains between tiles, and the accumulator is
// double-buffered in TMEM (2 x BN columns) so that the MMA of tile i+1 runs under the epilogue of tile i:
//   warp 0: TMA producer | warp 1: MMA issuer | warp 2: TMEM allocator | warp 3: idle | warps 4-11: epilogue
constexpr int PERSIST_THREADS = 384;
// 3 x 48 KB or 4 x 32 KB; EPI_DUAL: 3 x 32 KB (K = 384 is six chunks) to make room for the TMA staging boxes
template <int BN, int EPI = EPI_STORE_H> constexpr int persist_stages() { return (BN >= 256 || EPI == EPI_DUAL) ? 3 : 4; }
template <int BN> constexpr int staged_cols() { return (BN / 2 > 64) ? 64 : BN / 2; }
template <int BN, int EPI = EPI_STORE_H> constexpr int persist_staging_bytes() {
  return (EPI == EPI_DUAL) ? 8 * DUAL_STAGE_BYTES + DUAL_AUX_BYTES : 8 * 32 * (staged_cols<BN>() + 4) * 4;
}

template <int BN, int EPI = EPI_STORE_H>
constexpr int gemm_persistent_smem_bytes() {
  return persist_stages<BN, EPI>() * (BLOCK_M * BLOCK_K * 2 + BN * BLOCK_K * 2) + persist_staging_bytes<BN, EPI>() + 1024 + 256;
}

template <int BN, int EPI>
__global__ void __launch_bounds__(PERSIST_THREADS, 1)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const GemmParams p, const int tiles_m, const int tiles_n, const __grid_constant__ OutMaps om) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  constexpr int B_BYTES = BN * BLOCK_K * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int PERSIST_STAGES = persist_stages<BN, EPI>();
  constexpr int STAGING_BYTES = persist_staging_bytes<BN, EPI>();
  constexpr uint32_t TMEM_COLS = 2 * BN;

  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t staging = base + PERSIST_STAGES * STAGE_BYTES;        // 1024-byte aligned (stages are 32 / 48 KB)
  const uint32_t bar_base = staging + STAGING_BYTES;
  const uint32_t full_bar0 = bar_base;
  const uint32_t empty_bar0 = bar_base + 8 * PERSIST_STAGES;
  const uint32_t tfull_bar0 = bar_base + 16 * PERSIST_STAGES;        // [2]
  const uint32_t tempty_bar0 = tfull_bar0 + 16;                      // [2]
  const uint32_t tmem_ptr_addr = tempty_bar0 + 16;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_per_group = tiles_m * tiles_n;
  const int total = tiles_per_group * p.groups;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < PERSIST_STAGES; ++s) { mbar_init(full_bar0 + 8 * s, 1); mbar_init(empty_bar0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar0 + 8 * a, 1); mbar_init(tempty_bar0 + 8 * a, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer: the ring runs across tile boundaries =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int g = t / tiles_per_group, r = t - g * tiles_per_group;
        const int m0 = (r / tiles_n) * BLOCK_M, n0 = (r % tiles_n) * BN;
        const int a_col0 = p.a_col_base + g * p.a_col_group_off;
        const int a_row0 = m0 + g * p.a_row_group_off;
        const int b_row0 = n0 + g * p.b_row_group_off;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          const int tap = kc / p.chunks_per_tap;
          const int kin = kc - tap * p.chunks_per_tap;
          mbar_wait(empty_bar0 + 8 * stage, phase ^ 1);
          const uint32_t sa = base + stage * STAGE_BYTES;
          const uint32_t fb = full_bar0 + 8 * stage;
          mbar_expect_tx(fb, STAGE_BYTES);
          tma_load_2d(sa, &tmA, fb, a_col0 + kin * BLOCK_K, a_row0 + p.tap_shift[tap]);
          tma_load_2d(sa + A_BYTES, &tmB, fb, kc * BLOCK_K, b_row0);
          if (++stage == PERSIST_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: accumulator (it & 1) =====
    int stage = 0;
    uint32_t phase = 0;
    constexpr uint32_t idesc = umma_idesc_f16<BN>();
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(tempty_bar0 + 8 * acc, ((it >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + acc * BN;
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        mbar_wait(full_bar0 + 8 * stage, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = base + stage * STAGE_BYTES;
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_f16(tmem_acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kc > 0 || k > 0) ? 1u : 0u);
          umma_commit(empty_bar0 + 8 * stage);
          if (kc == p.k_chunks - 1) umma_commit(tfull_bar0 + 8 * acc);
        }
        __syncwarp();
        if (++stage == PERSIST_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue warps =====
    const int ew = warp - 4;
    const int q = warp & 3, half = ew >> 2;
    float* stage_buf = reinterpret_cast<float*>(smem_raw + (staging - raw)) + ew * (32 * (staged_cols<BN>() + 4));
    float* dual_stage = reinterpret_cast<float*>(smem_raw + (staging - raw) + ew * DUAL_STAGE_BYTES);
    float* dual_aux = reinterpret_cast<float*>(smem_raw + (staging - raw) + 8 * DUAL_STAGE_BYTES) + ew * 64;
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int g = t / tiles_per_group, r = t - g * tiles_per_group;
      const int n_tile = r % tiles_n;
      const int m0 = (r / tiles_n) * BLOCK_M, n0 = n_tile * BN;
      const int acc = it & 1;
      mbar_wait(tfull_bar0 + 8 * acc, (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tb = tempty_bar0 + 8 * acc;
      tile_epilogue<BN, EPI>(p, om, g, m0, n0, n_tile, q, half, lane, tmem_base + acc * BN, stage_buf, [&] {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tb) : "memory");
      }, nullptr, 0, dual_stage, dual_aux);
      __syncwarp();                       // staging block is reused by the next tile
    }
    if constexpr (EPI == EPI_DUAL) { if (p.out_tma && lane == 0) tma_store_wait_all(); }   // smem must outlive the bulk reads
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- 2-SM persistent kernel (cta_group::2) ----------------------------------------------------------------
// STATUS: written at the end of round 1 after the GPU budget was spent -- compiles for sm_100a, NOT yet executed on
// hardware, opt-in (MICKEY_GEMM_2SM=1), not covered by the default test run.
// Rationale (profiles/r01_notes.md): the big GEMMs sit on the chip's L2->SM operand throughput, not on the tensor
// pipe: a 128 x 256 tile loads 48 KB per 64-deep K step (85 flop per byte).  Here a CTA PAIR (cluster of 2 on one
// TPC) computes a 256 x 256 tile with tcgen05.mma.cta_group::2: each CTA loads its 128 rows of A and its 128 rows of
// B (32 KB per K step for 128 x 256 outputs: 128 flop per byte), the leader CTA issues one M=256, N=256 UMMA per
// 16-deep K slice that reads both CTAs' shared memory, and each CTA keeps its 128 accumulator rows in its own TMEM
// (2 x 256 columns, double-buffered across tiles).  Barriers: `full` lives in the leader (both CTAs' TMA loads
// complete_tx on it: peer bit of the mbarrier address cleared), `empty` / `tmem_full` are signalled in both CTAs by a
// multicast tcgen05.commit, `tmem_empty` lives in the leader and collects the 8 epilogue warps of both CTAs (the peer
// arrives remotely).  Warp roles as in the 1-SM persistent kernel; the peer's MMA warp idles.
// Ring depth vs. epilogue staging (227 KB per CTA): 4 stages of 32 KB + 64-column staging passes (8 warps x 32 x 68 floats),
// or 5 stages + 32-column passes (8 x 32 x 36 floats).  Measured (round 2, sessions 14/15, 64 images of ViT-B): the deeper
// ring pays where K is long and the A operand streams from HBM (mlp.fc2, K = 3072: 464 -> 452 us; 16384 x 4096 x 4096:
// 1526 -> 1546 TFLOP/s) and the shorter passes cost where K = 768 (qkv 338 -> 350, proj 224 -> 239, fc1 544 -> 560 us), so
// the dispatcher takes 5 stages from 32 K chunks on.  A staging-free variant (thread == row, 256-bit stores straight from the
// tcgen05.ld layout, 6-7 stages) was correct but slower everywhere K is short (proj 222 -> 263 us, qkv 343 -> 360 us): 32
// different 128-byte lines per store instruction cost more than the shared-memory round trip they saved; it was removed.
constexpr int TWO_SM_LONG_K_CHUNKS = 32;
// EPI_WARPS = 16 (four warps per TMEM lane quadrant, 64 columns each, 640 threads): twice the issue capacity for the
// epilogue; 32-column passes so that 16 staging blocks fit next to a 4-stage ring.
template <int TWO_SM_STAGES, int EPI_WARPS> constexpr int two_sm_pass_cols() { return (TWO_SM_STAGES >= 5 || EPI_WARPS > 8) ? 32 : 64; }
template <int TWO_SM_STAGES, int EPI_WARPS> constexpr int two_sm_staging_bytes() {
  return EPI_WARPS * 32 * (two_sm_pass_cols<TWO_SM_STAGES, EPI_WARPS>() + 4) * 4;
}
template <int TWO_SM_STAGES, int EPI_WARPS = 8> constexpr int gemm_2sm_smem_bytes() {
  return TWO_SM_STAGES * (BLOCK_M * BLOCK_K * 2 + 128 * BLOCK_K * 2) + two_sm_staging_bytes<TWO_SM_STAGES, EPI_WARPS>() + 1024 + 256;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (same offset, peer bit cleared)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once all previously issued MMAs have retired
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// arrive on the mbarrier at `local_bar`'s offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(rank));
  // default semantics (release at CTA scope: no fence instruction).  `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR
  // in front of the arrive, i.e. every epilogue warp waited for its in-flight global stores to be acknowledged before it
  // handed each accumulator back (ncu: `stall membar` among the top stalls of the K = 768 GEMMs).  Nothing in generic memory
  // is published through this barrier: it orders tcgen05.ld (completed by tcgen05.wait::ld + fence::before_thread_sync)
  // against the leader's next tcgen05.mma into the same TMEM columns.
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

template <int EPI, int TWO_SM_STAGES, int EPI_WARPS = 8>
__global__ void __launch_bounds__(128 + 32 * EPI_WARPS, 1)
gemm_tc_2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const GemmParams p, const int tiles_m, const int tiles_n, const __grid_constant__ OutMaps om) {
  // tiles_m = ceil(M / 256), tiles_n = N / 256; cluster c = blockIdx.x / 2 walks tiles c, c + gridDim.x / 2, ...
  extern __shared__ uint8_t smem_raw[];
  constexpr int BN = 256;
  constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;                    // this CTA's 128 rows of A
  constexpr int B_BYTES = 128 * BLOCK_K * 2;                        // this CTA's 128 rows (N half) of B
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int PASS_COLS = two_sm_pass_cols<TWO_SM_STAGES, EPI_WARPS>();
  constexpr int STAGING_BYTES = two_sm_staging_bytes<TWO_SM_STAGES, EPI_WARPS>();
  constexpr int NPARTS = EPI_WARPS / 4;
  constexpr uint32_t TMEM_COLS = 2 * BN;

  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t staging = base + TWO_SM_STAGES * STAGE_BYTES;
  const uint32_t bar_base = staging + STAGING_BYTES;
  const uint32_t full_bar0 = bar_base;
  const uint32_t empty_bar0 = bar_base + 8 * TWO_SM_STAGES;
  const uint32_t tfull_bar0 = bar_base + 16 * TWO_SM_STAGES;        // [2]
  const uint32_t tempty_bar0 = tfull_bar0 + 16;                     // [2], used in the leader only
  const uint32_t tmem_ptr_addr = tempty_bar0 + 16;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                          // 0 = leader
  const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  const int tiles_per_group = tiles_m * tiles_n;
  const int total = tiles_per_group * p.groups;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TWO_SM_STAGES; ++s) { mbar_init(full_bar0 + 8 * s, 1); mbar_init(empty_bar0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar0 + 8 * a, 1); mbar_init(tempty_bar0 + 8 * a, 2 * EPI_WARPS); }   // epilogue warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {     // both CTAs, same warp id, same destination offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                                               // barriers of both CTAs initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own 128 rows of A, own 128 rows of B; bytes credited to the leader's barrier =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < total; t += n_clusters) {
        const int g = t / tiles_per_group, r = t - g * tiles_per_group;
        const int m0 = (r / tiles_n) * 256 + (int)rank * 128, n0 = (r % tiles_n) * BN + (int)rank * 128;
        const int a_col0 = p.a_col_base + g * p.a_col_group_off;
        const int a_row0 = m0 + g * p.a_row_group_off;
        const int b_row0 = n0 + g * p.b_row_group_off;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          const int tap = kc / p.chunks_per_tap;
          const int kin = kc - tap * p.chunks_per_tap;
          mbar_wait(empty_bar0 + 8 * stage, phase ^ 1);             // local: the multicast commit frees the stage in both CTAs
          const uint32_t sa = base + stage * STAGE_BYTES;
          const uint32_t fb = full_bar0 + 8 * stage;
          if (rank == 0) mbar_expect_tx(fb, 2 * STAGE_BYTES);       // the pair's four loads of this stage
          tma_load_2d_2sm(sa, &tmA, fb, a_col0 + kin * BLOCK_K, a_row0 + p.tap_shift[tap]);
          tma_load_2d_2sm(sa + A_BYTES, &tmB, fb, kc * BLOCK_K, b_row0);
          if (++stage == TWO_SM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===== MMA issuer (leader CTA only): M = 256 over the pair, N = 256 =====
    int stage = 0;
    uint32_t phase = 0;
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    int it = 0;
    for (int t = cluster_id; t < total; t += n_clusters, ++it) {
      const int acc = it & 1;
      mbar_wait(tempty_bar0 + 8 * acc, ((it >> 1) & 1) ^ 1);       // both CTAs' epilogues have drained this accumulator
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + acc * BN;
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        mbar_wait(full_bar0 + 8 * stage, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = base + stage * STAGE_BYTES;
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_f16_2sm(tmem_acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kc > 0 || k > 0) ? 1u : 0u);
          umma_commit_2sm(empty_bar0 + 8 * stage);
          if (kc == p.k_chunks - 1) umma_commit_2sm(tfull_bar0 + 8 * acc);
        }
        __syncwarp();
        if (++stage == TWO_SM_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue warps (both CTAs): this CTA's 128 rows x 256 columns =====
    const int ew = warp - 4;
    const int q = warp & 3, half = ew >> 2;
    float* stage_buf = reinterpret_cast<float*>(smem_raw + (staging - raw)) + ew * (32 * (PASS_COLS + 4));
    int it = 0;
    for (int t = cluster_id; t < total; t += n_clusters, ++it) {
      const int g = t / tiles_per_group, r = t - g * tiles_per_group;
      const int n_tile = r % tiles_n;
      const int m0 = (r / tiles_n) * 256 + (int)rank * 128, n0 = n_tile * BN;
      const int acc = it & 1;
      mbar_wait(tfull_bar0 + 8 * acc, (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tb = tempty_bar0 + 8 * acc;
      tile_epilogue<BN, EPI, PASS_COLS, NPARTS>(p, om, g, m0, n0, n_tile, q, half, lane, tmem_base + acc * BN, stage_buf, [&] {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(tb, 0);                   // the leader's barrier (also from the leader itself)
      });
      __syncwarp();
    }
  }

  tc_fence_before();
  cluster_sync_all();                                               // nobody leaves while the pair still signals / reads
  tc_fence_after();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace mk
