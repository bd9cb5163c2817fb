// placeholder until the tcgen05 kernel lands (next commit)
#include "gemm_tf32_sm100.h"
namespace rri {
struct Tf32Gemm { int dummy; };
Tf32Gemm* tf32_gemm_create(int, int, std::string& err) { err = "tcgen05 contraction not built yet"; return nullptr; }
void tf32_gemm_destroy(Tf32Gemm* g) { delete g; }
int tf32_gemm_run(Tf32Gemm*, const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int, int64_t,
                  cudaStream_t, std::string& err) { err = "tcgen05 contraction not built yet"; return -1; }
}
