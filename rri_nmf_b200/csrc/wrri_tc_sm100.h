// wrri_tc_sm100.h -- tensor-core (tcgen05 TF32) statistics of the masked WRRI half-steps, sm_100a, fp32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace rri {
struct WrriTc;
WrriTc* wrri_tc_create(int sm_count, int64_t n, int64_t d, int k, std::string& err);
void wrri_tc_destroy(WrriTc* g);
float* wrri_tc_Wp(WrriTc* g);      // zero-padded operand copies [n,KP], [d,KP] (kept in step with W, T)
float* wrri_tc_Tp(WrriTc* g);
int wrri_tc_KP(WrriTc* g);
int wrri_tc_groups(WrriTc* g, int mode);           // partial slices written by mode 0 (T-step) / 1 (W-step)
void wrri_tc_load_factors(WrriTc* g, const float* W, const float* T, cudaStream_t st);
// mode 0: numer_part/denom_part[groups][d] (nmf.py:700-701); mode 1: [groups][n] (nmf.py:745-746).
// returns the number of kernels launched or -1
// Trow = row t of the caller's T (k x d, contiguous row of d floats).
int wrri_tc_stats(WrriTc* g, int mode, const float* X, int64_t ldx, const void* M, int mk, int64_t ldm, int t,
                  const float* Trow, float* numer_part, float* denom_part, int groups, cudaStream_t st, std::string& err);
}  // namespace rri
