// rri_kernels.cu -- interleaved-order (reference-exact) RRI sweep, unmasked.
//
// The reference's loop body for topic t (src/rri_nmf/nmf.py:415-476) is
//     T-step:  T[t,:] <- [ w_t'X - (w_t'W)_{t->0} T - reg ]_+ / (|w_t|^2 + reg + eps)     (:670-676, :437-447)
//     W-step:  W[:,t] <- [ X T_t' - W (T T_t')_{t->0} - reg ]_+ / (|T_t|^2 + reg + eps)    (:728-734, :464-469)
// and each half-step streams X once (2k passes per sweep).  Here one pass per topic streams X ONCE
// and feeds both contractions that are independent of each other:
//     y   = X T_t'        (needed by the W-step of topic t)
//     p   = w_{t+1}' X    (needed by the T-step of topic t+1; w_{t+1} is not touched by W-step t)
// so an exact interleaved sweep costs k (+1 per call) passes of X instead of 2k.  The small
// Gram vectors  g = w_{t+1}'W  and  h = T T_t'  are produced by the W-step / T-step kernels as
// per-block partials that the consumer reduces in a fixed order (no float atomics: N sweeps in one
// call are bit-identical to N calls of one sweep).
#include "common.cuh"
#include "kernels.h"

namespace rri {

constexpr int PASS_THREADS = 256;
constexpr int PASS_RU = 4;        // rows in flight per thread
constexpr int MAX_NCH = 5;

// ------------------------------------------------------------------------------------------------
// streaming pass
// ------------------------------------------------------------------------------------------------
// Every WARP owns a column slice of width 32*VEC*NCH (its T_t slice and p accumulators live in
// registers) and walks the rows of its row group PASS_RU at a time; the row dot products are finished
// with warp shuffles and written as per-slice partials ypart[slice][row].  No shared memory and no
// block barrier: the 8 warps of a CTA (and the 2 CTAs of an SM) drift freely, which keeps ~160 KB of
// 16-byte streaming loads in flight per SM.
template <typename T, int VEC, int NCH>
__global__ void __launch_bounds__(PASS_THREADS, 2)
rri_pass_kernel(const T* __restrict__ X, int64_t ldx, int64_t n, int64_t d,
                const T* __restrict__ tvec, const T* __restrict__ W, int k, int tn,
                T* __restrict__ ypart, T* __restrict__ ppart, int do_y, int do_p, int n_slices)
{
    using V = typename Vec<T>::type;
    constexpr int NW = PASS_THREADS / WARP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slice = blockIdx.x * NW + warp;
    if (slice >= n_slices) return;
    const int64_t c0 = (int64_t)slice * (WARP * VEC * NCH);
    int64_t r0, r1;
    part_range(n, gridDim.y, blockIdx.y, r0, r1);

    int64_t col[NCH];
    bool cok[NCH];
    T tv[NCH][VEC], pacc[NCH][VEC];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        col[j] = c0 + ((int64_t)j * WARP + lane) * VEC;
        cok[j] = col[j] < d;                       // d % VEC == 0 on the vector path
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            tv[j][v] = (do_y && cok[j]) ? tvec[col[j] + v] : T(0);
            pacc[j][v] = T(0);
        }
    }

    for (int64_t i = r0; i < r1; i += PASS_RU) {
        T x[PASS_RU][NCH][VEC];
        T wv[PASS_RU];
#pragma unroll
        for (int r = 0; r < PASS_RU; ++r) {
            const bool rok = (i + r) < r1;
            const T* xr = X + (i + r) * ldx;
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                if (rok && cok[j]) {
                    if (VEC > 1) {
                        V v = ld_stream(reinterpret_cast<const V*>(xr + col[j]));
                        unpack(v, x[r][j]);
                    } else {
                        x[r][j][0] = ld_stream(xr + col[j]);
                    }
                } else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) x[r][j][v] = T(0);
                }
            }
            wv[r] = (do_p && rok) ? W[(i + r) * k + tn] : T(0);
        }
        T ys[PASS_RU];
#pragma unroll
        for (int r = 0; r < PASS_RU; ++r) {
            T s = T(0);
#pragma unroll
            for (int j = 0; j < NCH; ++j)
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    s = fma(x[r][j][v], tv[j][v], s);
                    pacc[j][v] = fma(wv[r], x[r][j][v], pacc[j][v]);
                }
            ys[r] = s;
        }
        if (do_y) {
            T mine = T(0);
#pragma unroll
            for (int r = 0; r < PASS_RU; ++r) {
                const T tot = warp_sum(ys[r]);
                if (lane == r) mine = tot;
            }
            if (lane < PASS_RU && (i + lane) < r1) ypart[(int64_t)slice * n + i + lane] = mine;
        }
    }
    if (do_p) {
        T* pp = ppart + (int64_t)blockIdx.y * d;
#pragma unroll
        for (int j = 0; j < NCH; ++j)
            if (cok[j]) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) pp[col[j] + v] = pacc[j][v];
            }
    }
}

PassPlan plan_pass(int64_t n, int64_t d, int64_t ldx, const void* X, int elem_size, int sm_count)
{
    PassPlan pl;
    const int vmax = 16 / elem_size;
    const bool aligned = (d % vmax == 0) && (ldx % vmax == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
    pl.vec = aligned ? vmax : 1;
    const int64_t chunk = (int64_t)WARP * pl.vec;                    // columns one warp covers per chunk
    const int NW = PASS_THREADS / WARP;
    // number of CTAs along the columns with the widest slices, then the narrowest slice that still covers d
    const int64_t cta_cols = chunk * MAX_NCH * NW;
    const int ctas_x = (int)((d + cta_cols - 1) / cta_cols);
    pl.nch = (int)((d + chunk * NW * ctas_x - 1) / (chunk * NW * ctas_x));
    if (pl.nch < 1) pl.nch = 1;
    pl.cw = chunk * pl.nch;                                           // slice width
    pl.ct = (int)((d + pl.cw - 1) / pl.cw);                           // number of warp slices = y partials per row
    int target = 2 * sm_count;                                        // two resident CTAs per SM
    const int gx = (pl.ct + NW - 1) / NW;
    pl.rg = target / gx;
    if (pl.rg < 1) pl.rg = 1;
    int64_t max_rg = (n + PASS_RU - 1) / PASS_RU;
    if (pl.rg > max_rg) pl.rg = (int)(max_rg < 1 ? 1 : max_rg);
    return pl;
}

template <typename T, int VEC>
static void launch_pass_vec(const T* X, int64_t ldx, int64_t n, int64_t d, const T* tvec, const T* W,
                            int k, int tn, T* ypart, T* ppart, bool do_y, bool do_p,
                            const PassPlan& pl, cudaStream_t st)
{
    dim3 grid((pl.ct + PASS_THREADS / WARP - 1) / (PASS_THREADS / WARP), pl.rg);
#define RRI_PASS_CASE(N)                                                                      \
    case N:                                                                                   \
        rri_pass_kernel<T, VEC, N><<<grid, PASS_THREADS, 0, st>>>(X, ldx, n, d, tvec, W, k, tn, \
                                                                   ypart, ppart, do_y, do_p, pl.ct); \
        break;
    switch (pl.nch) {
        RRI_PASS_CASE(1) RRI_PASS_CASE(2) RRI_PASS_CASE(3) RRI_PASS_CASE(4) RRI_PASS_CASE(5)
    }
#undef RRI_PASS_CASE
}

template <typename T>
void launch_rri_pass(const T* X, int64_t ldx, int64_t n, int64_t d, const T* tvec, const T* W, int k,
                     int tn, T* ypart, T* ppart, bool do_y, bool do_p, const PassPlan& pl,
                     cudaStream_t st)
{
    if (pl.vec > 1) launch_pass_vec<T, Vec<T>::N>(X, ldx, n, d, tvec, W, k, tn, ypart, ppart, do_y, do_p, pl, st);
    else            launch_pass_vec<T, 1>(X, ldx, n, d, tvec, W, k, tn, ypart, ppart, do_y, do_p, pl, st);
}

// ------------------------------------------------------------------------------------------------
// T-step
// ------------------------------------------------------------------------------------------------
constexpr int TS_THREADS = 256;
int tstep_blocks(int64_t d) { return (int)((d + TS_THREADS - 1) / TS_THREADS); }

template <typename T>
__global__ void __launch_bounds__(TS_THREADS)
rri_tstep_kernel(T* __restrict__ Tm, int64_t d, int k, int t,
                 const T* __restrict__ ppart, int rg, int64_t pstride,
                 const T* __restrict__ gpart, int gb,
                 T reg_l1, T reg_l2, T eps, T ub, int has_ub,
                 T* __restrict__ hpart, double* __restrict__ sums, int t_prev,
                 int* __restrict__ flags, int do_update)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* gs = reinterpret_cast<T*>(smem_raw);          // [k+1]
    T* xs = gs + (k + 1);                            // [TS_THREADS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ks = k + 1;

    for (int j = tid; j < ks; j += TS_THREADS) {
        T s = T(0);
#pragma unroll 8
        for (int b = 0; b < gb; ++b) s += gpart[(int64_t)b * ks + j];
        gs[j] = s;
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid == 0 && t_prev >= 0) {
        // sum of the W column updated by the preceding W-step (nmf.py:793-794)
        double sw = (double)gs[k];
        sums[k + t_prev] = sw;
        if (!(sw > 1e-10)) atomicOr(flags, 2);
        if (!isfinite(sw)) atomicOr(flags, 8);
    }

    const int64_t c = (int64_t)blockIdx.x * TS_THREADS + tid;
    T x = T(0);
    if (c < d) {
        if (do_update) {
            T p = T(0);
#pragma unroll 8
            for (int r = 0; r < rg; ++r) p += ppart[(int64_t)r * pstride + c];
            T dot = T(0);
#pragma unroll 8
            for (int j = 0; j < k; ++j) {
                const T tj = Tm[(int64_t)j * d + c];
                dot = fma(j != t ? gs[j] : T(0), tj, dot);
            }
            bool unb = false;
            x = solve_scalar_c<T>(p - dot - reg_l1, gs[t] + reg_l2, eps, ub, has_ub != 0, unb);
            if (unb) atomicOr(flags, 4);
            Tm[(int64_t)t * d + c] = x;
        } else {
            x = Tm[(int64_t)t * d + c];
        }
    }
    xs[tid] = x;
    __syncthreads();

    // h[j] = sum_c T[j,c] x[c] over this block's columns; slot k = sum_c x[c]
    const int64_t cb = (int64_t)blockIdx.x * TS_THREADS;
    for (int j = warp; j < ks; j += TS_THREADS / WARP) {
        T s = T(0);
#pragma unroll
        for (int i = 0; i < TS_THREADS / WARP; ++i) {
            const int cc = lane + 32 * i;
            if (cb + cc < d) {
                T other = (j == k) ? T(1) : ((j == t) ? xs[cc] : Tm[(int64_t)j * d + cb + cc]);
                s = fma(other, xs[cc], s);
            }
        }
        s = warp_sum(s);
        if (lane == 0) hpart[(int64_t)blockIdx.x * ks + j] = s;
    }
}

template <typename T>
void launch_rri_tstep(T* Tm, int64_t d, int k, int t, const T* ppart, int rg, int64_t pstride,
                      const T* gpart, int gb, const SolveArgs& a, T* hpart, double* sums, int t_prev,
                      int* flags, bool do_update, cudaStream_t st)
{
    const int blocks = tstep_blocks(d);
    const size_t smem = sizeof(T) * (size_t)(k + 1 + TS_THREADS);
    rri_tstep_kernel<T><<<blocks, TS_THREADS, smem, st>>>(Tm, d, k, t, ppart, rg, pstride, gpart, gb,
                                                         (T)a.reg_l1, (T)a.reg_l2, (T)a.eps, (T)a.ub,
                                                         a.has_ub, hpart, sums, t_prev, flags,
                                                         do_update ? 1 : 0);
}

// ------------------------------------------------------------------------------------------------
// W-step
// ------------------------------------------------------------------------------------------------
constexpr int WS_THREADS = 256;
int wstep_blocks(int64_t n, int sm_count)
{
    int64_t b = (n + 63) / 64;                  // at least 64 rows (8 per warp) per block
    if (b > 2 * sm_count) b = 2 * sm_count;
    return (int)(b < 1 ? 1 : b);
}

template <typename T, int KL>
__global__ void __launch_bounds__(WS_THREADS)
rri_wstep_kernel(T* __restrict__ W, int64_t n, int k, int t, int tn,
                 const T* __restrict__ ypart, int ct, int64_t ystride,
                 const T* __restrict__ hpart, int hb,
                 T reg_l1, T reg_l2, T eps, T ub, int has_ub,
                 T* __restrict__ gpart, double* __restrict__ sums, int* __restrict__ flags, int do_update)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = WS_THREADS / WARP;
    T* hs = reinterpret_cast<T*>(smem_raw);            // [k+1]
    T* gsm = hs + (k + 1);                             // [NW][k+1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ks = k + 1;

    if (do_update) {
        for (int j = tid; j < ks; j += WS_THREADS) {
            T s = T(0);
#pragma unroll 8
            for (int b = 0; b < hb; ++b) s += hpart[(int64_t)b * ks + j];
            hs[j] = s;
        }
    }
    __syncthreads();
    T nt = T(0);
    if (do_update) {
        nt = hs[t];
        if (blockIdx.x == 0 && tid == 0) {
            double stt = (double)hs[k];                // sum of the T row just updated (nmf.py:757)
            sums[t] = stt;
            if (!(stt > 1e-10)) atomicOr(flags, 1);
            if (!isfinite(stt)) atomicOr(flags, 8);
        }
    }
    T hreg[KL];
#pragma unroll
    for (int l = 0; l < KL; ++l) {
        const int j = lane + 32 * l;
        hreg[l] = (do_update && j < k && j != t) ? hs[j] : T(0);
    }

    int64_t r0, r1;
    part_range(n, gridDim.x, blockIdx.x, r0, r1);
    T gacc[KL];
#pragma unroll
    for (int l = 0; l < KL; ++l) gacc[l] = T(0);
    T xsum = T(0);
    bool unb = false;
    const T denom = nt + reg_l2;

    // RW rows per warp iteration: their loads are independent, so the L2 round trips overlap
    constexpr int RW = 4;
    for (int64_t ib = r0 + (int64_t)warp * RW; ib < r1; ib += (int64_t)NW * RW) {
        T wrow[RW][KL], yv[RW], sv[RW];
#pragma unroll
        for (int q = 0; q < RW; ++q) {
            const int64_t i = ib + q;
            const bool rok = i < r1;
#pragma unroll
            for (int l = 0; l < KL; ++l) {
                const int j = lane + 32 * l;
                wrow[q][l] = (rok && j < k) ? W[i * k + j] : T(0);
            }
            T y = T(0);
            if (do_update && rok)
                for (int c = lane; c < ct; c += WARP) y += ypart[(int64_t)c * ystride + i];
            yv[q] = y;
        }
        if (do_update) {
#pragma unroll
            for (int q = 0; q < RW; ++q) {
                T s = T(0);
#pragma unroll
                for (int l = 0; l < KL; ++l) s = fma(wrow[q][l], hreg[l], s);
                sv[q] = s;
            }
#pragma unroll
            for (int q = 0; q < RW; ++q) { sv[q] = warp_sum(sv[q]); yv[q] = warp_sum(yv[q]); }
#pragma unroll
            for (int q = 0; q < RW; ++q) {
                const int64_t i = ib + q;
                if (i < r1) {
                    const T x = solve_scalar_c<T>(yv[q] - sv[q] - reg_l1, denom, eps, ub, has_ub != 0, unb);
                    xsum += x;
#pragma unroll
                    for (int l = 0; l < KL; ++l)
                        if (lane + 32 * l == t) { wrow[q][l] = x; W[i * k + t] = x; }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < RW; ++q) {
            T wsel = T(0);
#pragma unroll
            for (int l = 0; l < KL; ++l)
                if (l == (tn >> 5)) wsel = wrow[q][l];
            const T wtn = __shfl_sync(0xffffffffu, wsel, tn & 31);
#pragma unroll
            for (int l = 0; l < KL; ++l) gacc[l] = fma(wtn, wrow[q][l], gacc[l]);
        }
    }
    if (unb) atomicOr(flags, 4);

#pragma unroll
    for (int l = 0; l < KL; ++l) {
        const int j = lane + 32 * l;
        if (j < k) gsm[warp * ks + j] = gacc[l];
    }
    if (lane == 0) gsm[warp * ks + k] = xsum;
    __syncthreads();
    for (int j = tid; j < ks; j += WS_THREADS) {
        T s = T(0);
#pragma unroll
        for (int w = 0; w < NW; ++w) s += gsm[w * ks + j];
        gpart[(int64_t)blockIdx.x * ks + j] = s;
    }
}

template <typename T>
void launch_rri_wstep(T* W, int64_t n, int k, int t, int tn, const T* ypart, int ct, int64_t ystride,
                      const T* hpart, int hb, const SolveArgs& a, T* gpart, double* sums, int* flags,
                      bool do_update, int blocks, cudaStream_t st)
{
    const size_t smem = sizeof(T) * (size_t)(k + 1) * (1 + WS_THREADS / WARP);
    const int kl = (k + 31) / 32;
#define RRI_WS_CASE(N)                                                                             \
    rri_wstep_kernel<T, N><<<blocks, WS_THREADS, smem, st>>>(W, n, k, t, tn, ypart, ct, ystride,   \
        hpart, hb, (T)a.reg_l1, (T)a.reg_l2, (T)a.eps, (T)a.ub, a.has_ub, gpart, sums, flags,      \
        do_update ? 1 : 0)
    if (kl <= 1) RRI_WS_CASE(1);
    else if (kl <= 2) RRI_WS_CASE(2);
    else if (kl <= 4) RRI_WS_CASE(4);
    else RRI_WS_CASE(8);
#undef RRI_WS_CASE
}

// ------------------------------------------------------------------------------------------------
// small reductions
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void reduce_stat_kernel(const T* __restrict__ ppart, int rg, int64_t d,
                                   const T* __restrict__ gpart, int gb, int k, T* __restrict__ stat)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ks = k + 1;
    if (c < d) {
        T s = T(0);
        for (int r = 0; r < rg; ++r) s += ppart[(int64_t)r * d + c];
        stat[c] = s;
    } else if (c < d + ks) {
        const int j = (int)(c - d);
        T s = T(0);
        for (int b = 0; b < gb; ++b) s += gpart[(int64_t)b * ks + j];
        stat[c] = s;
    }
}

template <typename T>
void launch_reduce_stat(const T* ppart, int rg, int64_t d, const T* gpart, int gb, int k, T* stat,
                        cudaStream_t st)
{
    const int64_t tot = d + k + 1;
    reduce_stat_kernel<T><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(ppart, rg, d, gpart, gb, k, stat);
}

template <typename T>
__global__ void finalize_sums_kernel(const T* __restrict__ gpart, int gb, int k, int t_prev,
                                     double* __restrict__ sums, int* __restrict__ flags)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        T s = T(0);
        for (int b = 0; b < gb; ++b) s += gpart[(int64_t)b * (k + 1) + k];
        const double sw = (double)s;
        sums[k + t_prev] = sw;
        if (!(sw > 1e-10)) atomicOr(flags, 2);
        if (!isfinite(sw)) atomicOr(flags, 8);
    }
}

template <typename T>
void launch_finalize_sums(const T* gpart, int gb, int k, int t_prev, double* sums, int* flags,
                          cudaStream_t st)
{
    finalize_sums_kernel<T><<<1, 32, 0, st>>>(gpart, gb, k, t_prev, sums, flags);
}

#define RRI_INST(T)                                                                                   \
    template void launch_rri_pass<T>(const T*, int64_t, int64_t, int64_t, const T*, const T*, int, int, \
                                     T*, T*, bool, bool, const PassPlan&, cudaStream_t);              \
    template void launch_rri_tstep<T>(T*, int64_t, int, int, const T*, int, int64_t, const T*, int,    \
                                      const SolveArgs&, T*, double*, int, int*, bool, cudaStream_t);  \
    template void launch_rri_wstep<T>(T*, int64_t, int, int, int, const T*, int, int64_t, const T*,   \
                                      int, const SolveArgs&, T*, double*, int*, bool, int,            \
                                      cudaStream_t);                                                  \
    template void launch_reduce_stat<T>(const T*, int, int64_t, const T*, int, int, T*, cudaStream_t); \
    template void launch_finalize_sums<T>(const T*, int, int, int, double*, int*, cudaStream_t);
RRI_INST(float)
RRI_INST(double)

}  // namespace rri
