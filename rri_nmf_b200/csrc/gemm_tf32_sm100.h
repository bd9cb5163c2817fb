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
// returns the number of kernels launched, or -1 (err set)
int tf32_gemm_run(Tf32Gemm* g, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                  int64_t ldc, int64_t M, int N, int64_t K, cudaStream_t st, std::string& err);
}  // namespace rri
