// simplex_kernels.cu -- Euclidean projection of matrix rows onto the simplex {x >= 0, sum x = s}.
//
// The reference (matrixops.py:5-69, Duchi et al.) sorts the vector and finds
//     rho = max{ j : u_j * j > cumsum(u)_j - s },  theta = (cumsum(u)_rho - s) / rho,  w = max(v - theta, 0).
// The same theta is the fixed point of Michelot's iteration
//     theta <- (sum_{v_i > theta} v_i - s) / |{v_i > theta}|,   starting from the full set,
// which needs only reductions (no sort) and reaches the identical active set in a few rounds.
// One block per row; sums are accumulated in double for both element types.
#include "common.cuh"
#include "kernels.h"

namespace rri {

template <typename T>
__global__ void project_rows_simplex_kernel(T* __restrict__ A, int64_t rows, int64_t cols, double s)
{
    __shared__ double rs[32];
    __shared__ double rc[32];
    __shared__ double bc[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        T* v = A + r * cols;
        double theta = -1.0e300;        // "everything active" on the first round
        double prev_cnt = -1.0;
        for (int iter = 0; iter < 10000; ++iter) {
            double sum = 0.0, cnt = 0.0;
            for (int64_t i = tid; i < cols; i += blockDim.x) {
                const double x = (double)v[i];
                if (x > theta) { sum += x; cnt += 1.0; }
            }
            sum = warp_sum(sum); cnt = warp_sum(cnt);
            if (lane == 0) { rs[warp] = sum; rc[warp] = cnt; }
            __syncthreads();
            if (tid == 0) {
                double a = 0.0, b = 0.0;
                for (int w = 0; w < nw; ++w) { a += rs[w]; b += rc[w]; }
                bc[0] = a; bc[1] = b;
            }
            __syncthreads();
            const double tot = bc[0], c = bc[1];
            __syncthreads();
            if (c == prev_cnt || c == 0.0) break;       // active set unchanged: theta is final
            prev_cnt = c;
            theta = (tot - s) / c;
        }
        for (int64_t i = tid; i < cols; i += blockDim.x) {
            const double x = (double)v[i] - theta;
            v[i] = (T)(x > 0.0 ? x : 0.0);
        }
        __syncthreads();
    }
}

template <typename T>
void launch_project_rows_simplex(T* A, int64_t rows, int64_t cols, double s, cudaStream_t st)
{
    int threads = cols <= 64 ? 32 : (cols <= 4096 ? 256 : 1024);
    int64_t blocks = rows < 65535LL * 16 ? rows : 65535LL * 16;
    project_rows_simplex_kernel<T><<<(unsigned)blocks, threads, 0, st>>>(A, rows, cols, s);
}

template void launch_project_rows_simplex<float>(float*, int64_t, int64_t, double, cudaStream_t);
template void launch_project_rows_simplex<double>(double*, int64_t, int64_t, double, cudaStream_t);

__global__ void flag_from_sums_kernel(const double* __restrict__ sums, int off, int k, int zero_flag,
                                      int* __restrict__ flags)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < k) {
        const double v = sums[off + t];
        if (!(v > 1e-10)) atomicOr(flags, zero_flag);
        if (!isfinite(v)) atomicOr(flags, 8);
    }
}

void launch_flag_from_sums(const double* sums, int off, int k, int zero_flag, int* flags, cudaStream_t st)
{
    flag_from_sums_kernel<<<(k + 127) / 128, 128, 0, st>>>(sums, off, k, zero_flag, flags);
}

}  // namespace rri
