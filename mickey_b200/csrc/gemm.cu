// Host side of the GEMM family: TMA descriptor encoding, launch dispatch, and the SIMT debug kernel.
#include "gemm.h"
#include "gemm_tc.cuh"

#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace mk {

// ---- cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda needed) ---------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// Descriptors are pure functions of (pointer, shape, box); the engine reuses the same workspace and
// weight buffers every call, so they are cached.
struct MapKey {
  const void* ptr; long long rows, cols, ld; int box;
  bool operator==(const MapKey& o) const { return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box == o.box; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h = h * 1000003u ^ (size_t)k.rows; h = h * 1000003u ^ (size_t)k.cols; h = h * 1000003u ^ (size_t)k.ld;
    return h * 1000003u ^ (size_t)k.box;
  }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;
static std::mutex g_map_mutex;

static int encode_tensor_map_f16(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld_elems,
                                 int box_rows);

int make_tensor_map_f16(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld_elems,
                        int box_rows) {
  MapKey key{ptr, rows, cols, ld_elems, box_rows};
  std::lock_guard<std::mutex> lock(g_map_mutex);
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) { *map = it->second; return MK_OK; }
  int rc = encode_tensor_map_f16(map, ptr, rows, cols, ld_elems, box_rows);
  if (rc == MK_OK) {
    if (g_map_cache.size() > 4096) g_map_cache.clear();
    g_map_cache.emplace(key, *map);
  }
  return rc;
}

static int encode_tensor_map_f16(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld_elems,
                                 int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_last_error("cuTensorMapEncodeTiled entry point not available"); return MK_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld_elems % 8)) {
    set_last_error("tensor map operand must be 16-byte aligned with ld %% 8 == 0 (ptr=%p ld=%lld)", ptr, ld_elems);
    return MK_ERR_INVALID;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return MK_ERR_CUDA; }
  return MK_OK;
}

// fp32 [groups][n][n] tensor with row pitch `pitch` floats: box = 32 columns (128 bytes) x 32 rows x 1, 128-byte swizzle
// on the shared-memory side (the layout dual_store_chunk_tma writes), out-of-range rows / columns clipped by the hardware.
static int encode_tensor_map_out_f32(CUtensorMap* map, const void* ptr, long long n, long long pitch, long long groups) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_last_error("cuTensorMapEncodeTiled entry point not available"); return MK_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (pitch % 4) || pitch < n) {
    set_last_error("TMA output needs a 16-byte aligned base and a row pitch that is a multiple of 4 floats (ptr=%p pitch=%lld)", ptr, pitch);
    return MK_ERR_INVALID;
  }
  cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)groups};
  cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * 4 * (cuuint64_t)n};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled (fp32 output) failed with CUresult %d", (int)r); return MK_ERR_CUDA; }
  return MK_OK;
}

static const OutMaps& no_out_maps() { static OutMaps z = {}; return z; }

// ---- SIMT debug kernel --------------------------------------------------------------------------------
// Same operand addressing and the same epilogues as the tcgen05 kernel, computed with plain FFMA.  It is
// NOT a product path: it exists so that a GPU test can tell a tcgen05/TMA descriptor bug from an epilogue
// bug (tests/test_gpu_ops.py runs both and compares), selectable with MICKEY_GEMM_IMPL=simt.
template <int BN, int EPI>
__global__ void __launch_bounds__(128)
gemm_simt_kernel(const __half* __restrict__ A, long long a_rows, long long lda, const __half* __restrict__ B,
                 long long b_rows, long long ldb, const GemmParams p) {
  __shared__ __align__(16) __half As[BLOCK_M][BLOCK_K + 8];
  __shared__ __half Bs[BN][BLOCK_K + 8];
  pdl_trigger();
  pdl_wait();
  const int g = blockIdx.z, m0 = blockIdx.x * BLOCK_M, n0 = blockIdx.y * BN;
  const int t = threadIdx.x;
  float acc[BN];
#pragma unroll
  for (int j = 0; j < BN; ++j) acc[j] = 0.f;
  const int a_col0 = p.a_col_base + g * p.a_col_group_off;
  const long long a_row0 = (long long)m0 + (long long)g * p.a_row_group_off;
  const long long b_row0 = (long long)n0 + (long long)g * p.b_row_group_off;
  for (int kc = 0; kc < p.k_chunks; ++kc) {
    const int tap = kc / p.chunks_per_tap, kin = kc - tap * p.chunks_per_tap;
    for (int idx = t; idx < BLOCK_M * BLOCK_K; idx += 128) {
      const int r = idx / BLOCK_K, c = idx % BLOCK_K;
      const long long row = a_row0 + r + p.tap_shift[tap];
      const long long col = a_col0 + kin * BLOCK_K + c;
      As[r][c] = (row >= 0 && row < a_rows && col < lda) ? A[row * lda + col] : __float2half(0.f);
    }
    for (int idx = t; idx < BN * BLOCK_K; idx += 128) {
      const int r = idx / BLOCK_K, c = idx % BLOCK_K;
      const long long row = b_row0 + r;
      Bs[r][c] = (row < b_rows) ? B[row * ldb + (long long)kc * BLOCK_K + c] : __float2half(0.f);
    }
    __syncthreads();
    for (int k = 0; k < BLOCK_K; ++k) {
      const float a = __half2float(As[t][k]);
#pragma unroll
      for (int j = 0; j < BN; ++j) acc[j] = fmaf(a, __half2float(Bs[j][k]), acc[j]);
    }
    __syncthreads();
  }
  const int m = m0 + t;
  float v[32];
  if constexpr (EPI == EPI_LN) {
    float sum = 0.f;
    for (int j = 0; j < BN; ++j) sum += acc[j];
    const float mean = sum / BN;
    float sq = 0.f;
    for (int j = 0; j < BN; ++j) { const float d = acc[j] - mean; sq += d * d; }
    const float rstd = rsqrtf(sq / BN + p.eps);
    for (int c = 0; c < BN / 32; ++c) {
      for (int j = 0; j < 32; ++j) v[j] = acc[c * 32 + j];
      ln_store_chunk(p, g, m, n0 + c * 32, v, mean, rstd);
    }
  } else if constexpr (EPI == EPI_LSE || EPI == EPI_DUAL) {
    // the matcher epilogues exist on the tcgen05 kernels only (launch_gemm rejects them for this kernel)
  } else {
    for (int c = 0; c < BN / 32; ++c) {
      if (n0 + c * 32 < p.N) {
        for (int j = 0; j < 32; ++j) v[j] = acc[c * 32 + j];
        epilogue_chunk<EPI>(p, g, m, n0 + c * 32, v);
      }
    }
  }
}

bool pdl_enabled() {
  static int v = -1;
  // on by default (MICKEY_PDL=0 disables): with three steps in flight it measured +5.7 % pairs/s on the C2 workload
  // (775 vs 734, profiles/r01_notes.md).  An earlier build, with thread-per-row epilogues and no pipelining, was 2-3 %
  // slower with it -- re-measure when the kernels change.
  if (v < 0) { const char* e = getenv("MICKEY_PDL"); v = (e && strcmp(e, "0") == 0) ? 0 : 1; }
  return v == 1;
}

// ---- dispatch -----------------------------------------------------------------------------------------
static bool use_simt() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MICKEY_GEMM_IMPL");
    v = (e && strcmp(e, "simt") == 0) ? 1 : 0;
  }
  return v == 1;
}

template <int BN, int EPI, int STAGES>
static int launch_tc(dim3 grid, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream,
                     const OutMaps& om = no_out_maps()) {
  static_assert(EPI != EPI_DUAL || (BN == 128 && STAGES >= 3), "EPI_DUAL / TMA stages its output boxes in a >= 96 KB ring");
  static unsigned long long attr_mask = 0;
  constexpr int smem = gemm_smem_bytes<BN, STAGES>();
  if (first_use_on_device(attr_mask)) {
    MK_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  MK_CUDA_CHECK(launch_k(gemm_tc_kernel<BN, EPI, STAGES>, grid, dim3(GEMM_THREADS), (size_t)smem, stream, tmA, tmB, p, om));
  return MK_OK;
}

// EPI_RESID_LN: the grid.y CTAs of a row of tiles form one thread-block cluster (row statistics through DSMEM)
template <int BN, int EPI, int STAGES>
static int launch_tc_cluster(dim3 grid, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  static unsigned long long attr_mask = 0;
  constexpr int smem = gemm_smem_bytes<BN, STAGES>();
  if (first_use_on_device(attr_mask)) {
    MK_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = grid.y; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  MK_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, EPI, STAGES>, tmA, tmB, p, no_out_maps()));
  return MK_OK;
}

static bool wide_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MICKEY_GEMM_WIDE"); v = (e && strcmp(e, "0") == 0) ? 0 : 1; }
  return v == 1;
}

static int wide_min_pct() {   // 128 x 256 tiles when at least this many (in % of the SM count) remain; MICKEY_GEMM_WIDE_MIN_PCT
  static int v = -1;
  if (v < 0) { const char* e = getenv("MICKEY_GEMM_WIDE_MIN_PCT"); v = e ? atoi(e) : 200; if (v <= 0) v = 200; }
  return v;
}

static int sm_count() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& c = n[dev & 63];
  if (!c) { cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev); if (c <= 0) c = 148; }
  return c;
}

template <int BN, int EPI>
static int launch_persistent(dim3 tiles, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream,
                             const OutMaps& om = no_out_maps()) {
  static unsigned long long attr_mask = 0;
  constexpr int smem = gemm_persistent_smem_bytes<BN, EPI>();
  if (first_use_on_device(attr_mask)) {
    MK_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_persistent_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  const int sms = sm_count();
  const long long total = (long long)tiles.x * tiles.y * tiles.z;
  const unsigned grid = (unsigned)(total < sms ? total : sms);
  MK_CUDA_CHECK(launch_k(gemm_tc_persistent_kernel<BN, EPI>, dim3(grid), dim3(PERSIST_THREADS), (size_t)smem, stream, tmA, tmB, p,
                         (int)tiles.x, (int)tiles.y, om));
  return MK_OK;
}

// cta_group::2 kernel (256 x 256 tiles over a CTA pair): on by default where every CTA pair gets a tile
// (MICKEY_GEMM_2SM=0 disables).  First run on B200 in round 2: 16384x4096x4096 1365 -> 1526 TFLOP/s, the 4 x (512->512)
// 3x3 head convolution 75.8 -> 64.7 us (1055 -> 1238 TFLOP/s) against the 1-SM persistent kernel with 128 x 256 tiles.
static bool two_sm_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MICKEY_GEMM_2SM"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

// ring depth of the cta_group::2 kernel: 4 stages + 64-column epilogue passes, or 5 stages + 32-column passes for long K
// (gemm_tc.cuh; MICKEY_GEMM_2SM_STAGES=4|5 forces one)
static int two_sm_stages_forced() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MICKEY_GEMM_2SM_STAGES"); const int n = e ? atoi(e) : 0; v = (n == 4 || n == 5) ? n : 0; }
  return v;
}

static int two_sm_epi_warps_forced() {      // MICKEY_GEMM_2SM_EPIWARPS=8|16 forces one; default: by epilogue (launch_2sm)
  static int v = -1;
  if (v < 0) { const char* e = getenv("MICKEY_GEMM_2SM_EPIWARPS"); const int n = e ? atoi(e) : 0; v = (n == 8 || n == 16) ? n : 0; }
  return v;
}

template <int EPI, int STAGES, int EPI_WARPS>
static int launch_2sm_s(dim3 tiles256, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  static unsigned long long attr_mask = 0;
  constexpr int smem = gemm_2sm_smem_bytes<STAGES, EPI_WARPS>();
  static_assert(smem <= 227 * 1024, "cta_group::2 kernel: shared memory");
  if (first_use_on_device(attr_mask)) {
    MK_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_2sm_kernel<EPI, STAGES, EPI_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  const long long total = (long long)tiles256.x * tiles256.y * tiles256.z;
  const long long pairs = sm_count() / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * (total < pairs ? total : pairs))); cfg.blockDim = dim3(128 + 32 * EPI_WARPS);
  cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  MK_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_tc_2sm_kernel<EPI, STAGES, EPI_WARPS>, tmA, tmB, p, (int)tiles256.x, (int)tiles256.y, no_out_maps()));
  return MK_OK;
}
template <int EPI>
static int launch_2sm(dim3 tiles256, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  const int forced = two_sm_stages_forced();
  // the plain fp16 store (attn.qkv) also takes the deeper ring: 338 -> 330 us on 64 images of ViT-B (session 17)
  const bool plain_h = (EPI == EPI_STORE_H) && p.act == ACT_NONE && p.k_chunks >= 8;
  const int st = forced ? forced : ((p.k_chunks >= TWO_SM_LONG_K_CHUNKS || plain_h) ? 5 : 4);
  if constexpr (EPI == EPI_STORE_H || EPI == EPI_RESID_F || EPI == EPI_STORE_F) {
    // Short K: a heavy epilogue needs more issue slots than two warps per sub-partition give.  16 epilogue warps (four per
    // TMEM lane quadrant, 32-column passes) measured against 8 on 64 images of ViT-B: mlp.fc1 + GELU 546 -> 516 us alone,
    // 7.29 -> 6.42 ms in the C3 step; attn.proj (fp32 read-modify-write) 218 -> 204 us, 2.96 -> 2.63 ms; the plain fp16
    // store of attn.qkv loses 1 % to the shorter passes and keeps 8.
    const int forced_w = two_sm_epi_warps_forced();
    const bool heavy = (EPI == EPI_RESID_F) || (EPI == EPI_STORE_H && p.act != ACT_NONE);
    if (st == 4 && (forced_w ? forced_w == 16 : heavy)) return launch_2sm_s<EPI, 4, 16>(tiles256, tmA, tmB, p, stream);
  }
  return st == 5 ? launch_2sm_s<EPI, 5, 8>(tiles256, tmA, tmB, p, stream) : launch_2sm_s<EPI, 4, 8>(tiles256, tmA, tmB, p, stream);
}

static bool three_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MICKEY_GEMM_THREE"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

static bool persistent_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MICKEY_GEMM_PERSISTENT"); v = (e && strcmp(e, "0") == 0) ? 0 : 1; }
  return v == 1;
}

template <int BN, int EPI>
static int launch_one(const GemmOperand& A, const GemmOperand& B, const GemmParams& p, cudaStream_t stream, int impl) {
  dim3 grid(ceil_div(p.M, BLOCK_M), ceil_div(p.N, BN), p.groups);
  if constexpr (EPI == EPI_RESID_LN) {
    if (impl == GEMM_IMPL_SIMT || BN != 128 || grid.y > 8 || grid.z != 1) { set_last_error("EPI_RESID_LN: tcgen05 path, N <= 1024, one group"); return MK_ERR_UNSUPPORTED; }
  }
  if (impl == GEMM_IMPL_SIMT) {
    if constexpr (EPI != EPI_RESID_LN)
    MK_CUDA_CHECK(launch_k(gemm_simt_kernel<BN, EPI>, grid, dim3(128), 0, stream, reinterpret_cast<const __half*>(A.ptr),
                           (long long)A.rows, (long long)A.ld, reinterpret_cast<const __half*>(B.ptr), (long long)B.rows,
                           (long long)B.ld, p));
  } else {
    CUtensorMap tmA, tmB;
    int rc = make_tensor_map_f16(&tmA, A.ptr, A.rows, A.cols, A.ld, BLOCK_M);
    if (rc) return rc;
    rc = make_tensor_map_f16(&tmB, B.ptr, B.rows, B.cols, B.ld, BN);
    if (rc) return rc;
    if constexpr (EPI == EPI_DUAL) {
      if (p.out_tma) {
        // the three N x N outputs leave through TMA tensor stores.  One pair (256 tiles): one-tile CTAs, two per SM, all
        // resident at once (the kernel is latency-bound there); batches: the persistent kernel, one CTA per SM.
        OutMaps om = {};
        float* outs[3] = {p.scores, p.kp_scores, p.final_scores};
        for (int i = 0; i < 3; ++i)
          if (outs[i]) { rc = encode_tensor_map_out_f32(&om.m[i], outs[i], p.n_valid, p.out_pitch, p.groups); if (rc) return rc; }
        if constexpr (BN == 128) {
          if ((long long)grid.x * grid.y * grid.z <= 2LL * sm_count()) return launch_tc<BN, EPI, 3>(grid, tmA, tmB, p, stream, om);
        }
        return launch_persistent<BN, EPI>(grid, tmA, tmB, p, stream, om);
      }
    }
    // grids that give every SM at most ~one CTA run the deep ring; bigger grids keep two CTAs per SM
    const long long ctas = (long long)grid.x * grid.y * grid.z;
    const bool deep = ctas <= (long long)sm_count() * 5 / 4 && p.k_chunks > 3;
    if constexpr (EPI == EPI_RESID_LN) {       // one-tile kernels only (every warp of the CTA takes part in the cluster barriers)
      if constexpr (BN == 128) {
        if (deep) return launch_tc_cluster<BN, EPI, 6>(grid, tmA, tmB, p, stream);
        return launch_tc_cluster<BN, EPI, 3>(grid, tmA, tmB, p, stream);
      }
      return MK_ERR_UNSUPPORTED;
    }
    if (deep) return launch_tc<BN, EPI, 6>(grid, tmA, tmB, p, stream);
    // short-K grids of 2..3 CTAs per SM (ViT-S mlp.fc1: 372 tiles on 148 SMs) would run a second, quarter-full wave
    // with two resident CTAs; a 2-stage ring fits three per SM and keeps the GEMM in one wave
    if constexpr (BN == 128 && EPI == EPI_STORE_H) {
      if (three_enabled() && p.k_chunks <= 8 && ctas > 2LL * sm_count() && ctas <= 3LL * sm_count())
        return launch_tc<BN, EPI, 2>(grid, tmA, tmB, p, stream);
    }
    // (Short-K head GEMMs, K = 128 / 256 with LayerNorm or fp32-store epilogues, were also tried on one-tile CTAs, two per
    // SM: att.qkv 0.78 -> 1.22 ms, mlp.0 0.34 -> 0.69 ms per C3 step -- the persistent kernel stays; session 20.)
    // Persistent tile loop (accumulator double-buffered in TMEM, ring never drains) when the main loop dominates a
    // tile (K >= 768) or there are many tiles per SM.  Measured on B200: +29 % on the ViT-B GEMMs of the B=32
    // workload (523 -> 674 TFLOP/s), 1.13 PFLOP/s on a 16384x4096x4096 GEMM; but for the K=384, ~2-tiles-per-SM GEMMs
    // of the B=1 workload the two co-resident one-tile CTAs (16 epilogue warps per SM instead of 8) are faster.
    if (persistent_enabled() && (p.k_chunks >= 12 || ctas >= 4LL * sm_count())) {
      // 128 x 256 tiles (one N=256 UMMA per K step, A tile re-read half as often: 85 instead of 64 flop per byte of
      // L2 traffic, which is what bounds the 128 x 128 tiling) when N allows and enough tiles remain
      if constexpr (BN == 128 && (EPI == EPI_STORE_H || EPI == EPI_RESID_F || EPI == EPI_CONV || EPI == EPI_STORE_F)) {
        // tmB's 128-row box is exactly one CTA's half of the 256 B rows
        if (two_sm_enabled() && p.N % 256 == 0 && (long long)ceil_div(p.M, 256) * (p.N / 256) * grid.z >= sm_count() / 2)
          return launch_2sm<EPI>(dim3(ceil_div(p.M, 256), p.N / 256, grid.z), tmA, tmB, p, stream);
        if (wide_enabled() && p.N % 256 == 0 && (ctas / 2) * 100 >= (long long)wide_min_pct() * sm_count()) {
          CUtensorMap tmB2;
          rc = make_tensor_map_f16(&tmB2, B.ptr, B.rows, B.cols, B.ld, 256);
          if (rc) return rc;
          return launch_persistent<256, EPI>(dim3(grid.x, grid.y / 2, grid.z), tmA, tmB2, p, stream);
        }
      }
      return launch_persistent<BN, EPI>(grid, tmA, tmB, p, stream);
    }
    return launch_tc<BN, EPI, 3>(grid, tmA, tmB, p, stream);
  }
  MK_CUDA_CHECK(cudaGetLastError());
  return MK_OK;
}

template <int EPI>
static int launch_bn(int bn, const GemmOperand& A, const GemmOperand& B, const GemmParams& p, cudaStream_t s, int impl) {
  if (bn == 128) return launch_one<128, EPI>(A, B, p, s, impl);
  if (bn == 64) return launch_one<64, EPI>(A, B, p, s, impl);
  set_last_error("unsupported BLOCK_N %d", bn);
  return MK_ERR_INVALID;
}

// Would launch_gemm run a [M, N] x K RESID_F GEMM on a one-tile kernel (so that LayerNorm can be fused into it)?
bool gemm_resid_ln_supported(int M, int N, int k_chunks) {
  // opt-in (MICKEY_FUSE_LN=1).  Measured on the C2 workload: the fused epilogue removes 23 LayerNorm launches (4.8 us each)
  // but its three cluster barriers and two extra passes over the staged tile cost attn.proj / mlp.fc2 +5.5 us each:
  // 1.852 vs 1.808 ms per step, 754 vs 798 pairs/s (profiles/r01_notes.md).  Kept for the next round (single-exchange
  // statistics, push instead of pull).
  static int on = -1;
  if (on < 0) { const char* e = getenv("MICKEY_FUSE_LN"); on = (e && e[0] == '1') ? 1 : 0; }
  if (!on || use_simt() || N % 128 || N > 1024) return false;
  const long long ctas = (long long)ceil_div(M, BLOCK_M) * (N / 128);
  const bool deep = ctas <= (long long)sm_count() * 5 / 4 && k_chunks > 3;
  if (deep) return true;
  return !(persistent_enabled() && (k_chunks >= 12 || ctas >= 4LL * sm_count()));
}

int launch_gemm(int epi, const GemmOperand& A, const GemmOperand& B, const GemmParams& p, cudaStream_t stream, int impl) {
  if (impl == GEMM_IMPL_DEFAULT) impl = use_simt() ? GEMM_IMPL_SIMT : GEMM_IMPL_TC;
  if (p.k_chunks <= 0 || p.M <= 0 || p.N <= 0 || p.groups <= 0) { set_last_error("bad GEMM shape"); return MK_ERR_INVALID; }
  const bool matcher = (epi == EPI_LSE || epi == EPI_DUAL);
  if (matcher && impl == GEMM_IMPL_SIMT) { set_last_error("the matcher epilogues run on the tcgen05 kernels only"); return MK_ERR_UNSUPPORTED; }
  if (matcher && (p.part_ld % 128 || p.part_ld < p.n_valid)) { set_last_error("matcher: part_ld must be n_valid rounded up to 128"); return MK_ERR_INVALID; }
  if (epi == EPI_DUAL && p.out_pitch < p.n_valid) { set_last_error("matcher: out_pitch %lld < n_valid %d", p.out_pitch, p.n_valid); return MK_ERR_INVALID; }
  if (!matcher && (p.N % 32)) { set_last_error("GEMM N=%d must be a multiple of 32", p.N); return MK_ERR_INVALID; }
  int bn = (matcher || p.N % 128 == 0) ? 128 : 64;
  if (!matcher && p.N % bn) { set_last_error("GEMM N=%d not tileable", p.N); return MK_ERR_INVALID; }
  if (epi == EPI_LN && p.N != 128) { set_last_error("EPI_LN needs N == 128"); return MK_ERR_INVALID; }
  if (epi == EPI_RESID_LN && (p.N % 128 || p.N > 1024 || p.groups != 1 || !p.aux || !p.beta || !p.out_h || !p.out_f)) {
    set_last_error("EPI_RESID_LN needs N % 128 == 0, N <= 1024, one group, LN weight (aux), LN bias (beta), out_f and out_h");
    return MK_ERR_INVALID;
  }
  switch (epi) {
    case EPI_STORE_H: return launch_bn<EPI_STORE_H>(bn, A, B, p, stream, impl);
    case EPI_RESID_F: return launch_bn<EPI_RESID_F>(bn, A, B, p, stream, impl);
    case EPI_RESID_LN: return launch_one<128, EPI_RESID_LN>(A, B, p, stream, impl);
    case EPI_PATCH:   return launch_bn<EPI_PATCH>(bn, A, B, p, stream, impl);
    case EPI_CONV:    return launch_bn<EPI_CONV>(bn, A, B, p, stream, impl);
    case EPI_STORE_F: return launch_bn<EPI_STORE_F>(bn, A, B, p, stream, impl);
    case EPI_LN:      return launch_one<128, EPI_LN>(A, B, p, stream, impl);
    case EPI_LSE:     return launch_one<128, EPI_LSE>(A, B, p, stream, impl);
    case EPI_DUAL:    return launch_one<128, EPI_DUAL>(A, B, p, stream, impl);
  }
  set_last_error("unknown epilogue %d", epi);
  return MK_ERR_INVALID;
}

// ---- error string -------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

}  // namespace mk
