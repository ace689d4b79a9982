// Shared helpers for the apyib_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/apyib_b200.h"

namespace apyib {

void set_error(const char *fmt, ...);

#define APYIB_CUDA_CHECK(expr)                                                              \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            apyib::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                  \
                             cudaGetErrorString(_e));                                       \
            return APYIB_ERR_CUDA;                                                          \
        }                                                                                   \
    } while (0)

#define APYIB_LAUNCH_CHECK() APYIB_CUDA_CHECK(cudaGetLastError())

#define APYIB_REQUIRE(cond, msg)                                                            \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            apyib::set_error("%s:%d: argument check failed: %s (%s)", __FILE__, __LINE__,   \
                             #cond, msg);                                                   \
            return APYIB_ERR_ARG;                                                           \
        }                                                                                   \
    } while (0)

// interleaved complex128, layout-compatible with numpy / torch complex128
struct __align__(16) cplx {
    double x, y;
};

__host__ __device__ __forceinline__ cplx make_cplx(double x, double y) {
    cplx r;
    r.x = x;
    r.y = y;
    return r;
}
__host__ __device__ __forceinline__ cplx operator+(cplx a, cplx b) { return make_cplx(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx operator-(cplx a, cplx b) { return make_cplx(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx operator*(cplx a, cplx b) {
    return make_cplx(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx operator*(double s, cplx a) { return make_cplx(s * a.x, s * a.y); }
__host__ __device__ __forceinline__ cplx conj(cplx a) { return make_cplx(a.x, -a.y); }
__host__ __device__ __forceinline__ double abs2(cplx a) { return a.x * a.x + a.y * a.y; }
__host__ __device__ __forceinline__ cplx operator/(cplx a, double s) { return make_cplx(a.x / s, a.y / s); }
// Smith-free complex division (inputs are O(1) magnitudes; used for pivots / denominators).
__host__ __device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    double d = b.x * b.x + b.y * b.y;
    return make_cplx((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}

__device__ __forceinline__ cplx ldg(const cplx *p) {
    const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
    return make_cplx(v.x, v.y);
}

// scalar traits so that kernels can be templated on double / cplx
template <typename T> struct scalar;
template <> struct scalar<double> {
    static constexpr int is_complex = 0;
    __host__ __device__ static double zero() { return 0.0; }
    __host__ __device__ static double make(double re, double) { return re; }
    __host__ __device__ static double re(double a) { return a; }
    __host__ __device__ static double im(double) { return 0.0; }
    __host__ __device__ static double cj(double a) { return a; }
    __host__ __device__ static double div_real(double a, double d) { return a / d; }
};
template <> struct scalar<cplx> {
    static constexpr int is_complex = 1;
    __host__ __device__ static cplx zero() { return make_cplx(0.0, 0.0); }
    __host__ __device__ static cplx make(double re, double im) { return make_cplx(re, im); }
    __host__ __device__ static double re(cplx a) { return a.x; }
    __host__ __device__ static double im(cplx a) { return a.y; }
    __host__ __device__ static cplx cj(cplx a) { return conj(a); }
    __host__ __device__ static cplx div_real(cplx a, double d) { return make_cplx(a.x / d, a.y / d); }
};

// Deterministic grid-wide sum: every block writes its partial (NV values) to
// partials[block*NV + v]; the last block to finish (ticket counter) adds them in block
// order and hands the totals to `fin`.  The counter resets itself for the next launch.
constexpr int kReduceMaxBlocks = 1184;   // 8 x 148 SMs
constexpr int kReduceMaxVals = 16;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of NV doubles per thread; result valid in thread 0 (vals[] overwritten).
template <int NV, int THREADS> __device__ __forceinline__ void block_sum(double (&vals)[NV]) {
    __shared__ double sm[NV][THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double s = warp_sum(vals[v]);
        if (lane == 0) sm[v][w] = s;
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double s = (lane < THREADS / 32) ? sm[v][lane] : 0.0;
            s = warp_sum(s);
            vals[v] = s;
        }
    }
    __syncthreads();
}

// partials layout: [0 .. kReduceMaxBlocks*kReduceMaxVals) doubles, then one uint32 ticket.
template <int NV, int THREADS, typename Fin>
__device__ __forceinline__ void grid_sum_finish(double (&vals)[NV], double *partials, Fin fin) {
    block_sum<NV, THREADS>(vals);
    unsigned int *ticket = reinterpret_cast<unsigned int *>(partials + kReduceMaxBlocks * kReduceMaxVals);
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) partials[(size_t)blockIdx.x * NV + v] = vals[v];
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = 0.0;
        // fixed order: thread t sums blocks t, t+THREADS, ... then tree in block_sum
        for (int b = threadIdx.x; b < (int)gridDim.x; b += THREADS) {
#pragma unroll
            for (int v = 0; v < NV; ++v) acc[v] += __ldcg(&partials[(size_t)b * NV + v]);
        }
        block_sum<NV, THREADS>(acc);
        if (threadIdx.x == 0) {
            fin(acc);
            *ticket = 0u;
        }
    }
}

inline int reduce_grid(int64_t work_items, int threads, int sm_count_hint = 148) {
    int64_t b = (work_items + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count_hint * 8;
    if (cap > kReduceMaxBlocks) cap = kReduceMaxBlocks;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace apyib
