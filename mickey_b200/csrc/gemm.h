// Host-callable interface of the GEMM family (see gemm_tc.cuh for the kernel).
#pragma once
#include "common.cuh"

namespace mk {

struct GemmOperand {       // a row-major fp16 matrix [rows, cols] with leading dimension ld (elements)
  const void* ptr;
  long long rows, cols, ld;
};

enum GemmImpl : int { GEMM_IMPL_DEFAULT = 0, GEMM_IMPL_TC = 1, GEMM_IMPL_SIMT = 2 };

int make_tensor_map_f16(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld_elems, int box_rows);
int launch_gemm(int epi, const GemmOperand& A, const GemmOperand& B, const GemmParams& p, cudaStream_t stream,
                int impl = GEMM_IMPL_DEFAULT);
// true when launch_gemm would put a [M, N] x (64 k_chunks) residual GEMM on a one-tile kernel, i.e. EPI_RESID_LN may be used
bool gemm_resid_ln_supported(int M, int N, int k_chunks);
const char* last_error();

}  // namespace mk
