// gemm_tf32_sm100.h -- TF32 tcgen05 contraction C[M,N] = A[M,K] * B[N,K]^T for tall-skinny outputs
// (N = rank k <= 256, M*K = the data matrix), sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace rri {
struct Tf32Gemm;
Tf32Gemm* tf32_gemm_create(int sm_count, int nmax, std::string& err);
void tf32_gemm_destroy(Tf32Gemm* g);
// returns the number of kernels launched, or -1 (err set).
// Optional auxiliary output in the same launch: C2[M2, N] (leading dimension ldc2) = A2[M2, K] * B[N, K]^T -- with
// A2 = B this is the k x k Gram matrix of the factor, computed as one more row tile of the streaming contraction.
int tf32_gemm_run(Tf32Gemm* g, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                  int64_t ldc, int64_t M, int N, int64_t K, cudaStream_t st, std::string& err,
                  const float* A2 = nullptr, int64_t lda2 = 0, int M2 = 0, float* C2 = nullptr, int64_t ldc2 = 0);
}  // namespace rri
