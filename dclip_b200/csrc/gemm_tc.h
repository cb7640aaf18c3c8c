// Grouped tcgen05 GEMM (gemm_tc.cu): host-side problem description.
#pragma once
#include <cuda_runtime.h>

namespace dclip {

constexpr int kTcGemmMax = 8;

// C[M,N] (fp32, ldc) = A'(M,K) . B'(K,N) + bias[n] + residual[m,n];  A, B are 16-bit (bf16 or fp16).
//   a_mn = 0: A stored [M, K] (lda = row stride in elements);  a_mn = 1: A stored [K, M]
//   b_mn = 0: B stored [N, K];                                  b_mn = 1: B stored [K, N]
// bias / residual are fp32 (kind 0) or of the operands' 16-bit type (kind 1); outputs: fp32 C and / or 16-bit C16.
struct TcGemmProblem {
    const void* A; const void* B; float* C; const void* bias; const void* residual; void* C16;
    int M, N, K, lda, ldb, ldc, ldr;
    int a_mn, b_mn, accumulate;
    int ldc16;      // row stride of C16 (0 = ldc); C may be null when only the 16-bit output is wanted
    int bias_kind, res_kind;
    float alpha;    // C = alpha * A'B' + ... (0 = 1)
};

int launch_gemm_tc(const TcGemmProblem* probs, int count, int dtype16, cudaStream_t st);

}  // namespace dclip
