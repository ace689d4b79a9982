// Determinants of substituted occupied-overlap matrices (aats.py:120-130, 558-642) and the
// fused det-table x amplitude-vector products of the AAT assembly (aats.py:718-737 ...).
//
// One sub-warp ("group" of G = 4/8/16/32 lanes, n <= G) per n x n complex matrix: lane l owns
// matrix row l in registers, Gaussian elimination with partial pivoting (LAPACK zgetrf pivot
// rule: max |re|+|im|) runs with warp shuffles, rows are never physically swapped -- the
// permutation parity is tracked with a ballot.  The matrix is formed on the fly from the
// (L1/L2-resident) MO overlap S and two index lists, so neither the substituted matrices nor,
// in the fused variant, the determinant table (the reference's 8-index tensor, aats.py:575)
// ever exist in memory.
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace apyib {

constexpr int kDetThreads = 256;

// LU with partial pivoting of an N x N complex matrix spread over a group of N lanes (lane l of
// the group = matrix row l, in registers, fully unrolled -> static register indexing).
//   pivot search : one REDUX.MAX on the high word of |re|+|im| (LAPACK izamax metric, 20-bit
//                  mantissa resolution is plenty for choosing a pivot) + one ballot;
//   pivot row    : broadcast with shuffles from the owning lane, rows are never moved -- the
//                  permutation parity comes from a ballot of the not-yet-eliminated rows;
//   multipliers  : one reciprocal per step instead of a complex division per row.
// Lanes that belong to no group (32 % N leftovers) run the same instruction stream on a dummy
// identity row with a single-lane member mask.
template <int N>
__device__ __forceinline__ cplx group_lu_det(cplx (&a)[N], const int l, const int g, const int gbase) {
    constexpr int GPW = 32 / N;
    constexpr unsigned gm = (N == 32) ? 0xffffffffu : ((1u << N) - 1u);
    bool done = (g >= GPW);                          // leftover lanes never take part
    double detx = 1.0, dety = 0.0;
    unsigned parity = 0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double mag = fabs(a[k].x) + fabs(a[k].y);
        const unsigned key = done ? 0u : ((unsigned)__double2hiint(mag) + 1u);
        // per-group maximum with GPW full-warp REDUX ops (a REDUX over sub-warp member masks
        // would be serialised by the compiler into a loop over the distinct masks)
        unsigned kmax = 0u;
#pragma unroll
        for (int G = 0; G < GPW; ++G) {
            const unsigned m = __reduce_max_sync(0xffffffffu, (g == G) ? key : 0u);
            kmax = (g == G) ? m : kmax;
        }
        const unsigned undone = (__ballot_sync(0xffffffffu, !done) >> gbase) & gm;
        const unsigned cand = (__ballot_sync(0xffffffffu, (!done) && key == kmax) >> gbase) & gm;
        const int who = cand ? (__ffs(cand) - 1) : 0;          // lowest row among the maxima
        parity ^= (unsigned)__popc(undone & ((1u << who) - 1u));
        const int src = gbase + who;
        const double pvx = __shfl_sync(0xffffffffu, a[k].x, src);
        const double pvy = __shfl_sync(0xffffffffu, a[k].y, src);
        const double ndx = detx * pvx - dety * pvy;
        dety = detx * pvy + dety * pvx;
        detx = ndx;
        done = done || (l == who);
        const double d2 = fma(pvx, pvx, pvy * pvy);
        const double rinv = (d2 > 0.0) ? __drcp_rn(d2) : 0.0;  // exactly singular column -> det = 0
        const double ix = pvx * rinv, iy = -pvy * rinv;
        double fx = fma(a[k].x, ix, -a[k].y * iy), fy = fma(a[k].x, iy, a[k].y * ix);
        fx = done ? 0.0 : fx;
        fy = done ? 0.0 : fy;
#pragma unroll
        for (int j = k + 1; j < N; ++j) {
            const double px = __shfl_sync(0xffffffffu, a[j].x, src);
            const double py = __shfl_sync(0xffffffffu, a[j].y, src);
            a[j].x = fma(fy, py, fma(-fx, px, a[j].x));
            a[j].y = fma(-fy, px, fma(-fx, py, a[j].y));
        }
    }
    const double sgn = (parity & 1u) ? -1.0 : 1.0;
    return make_cplx(sgn * detx, sgn * dety);
}

// n <= N: the matrix is embedded as blockdiag(M, 1).  Each warp carries 32/N groups; a group owns
// one row list r and loops over a chunk of column lists.
// grid: x = row blocks, y = column chunks.
// OUTER : out[r*ncol + c] = det
// !OUTER: Zp[(chunk*ny + iy)*nrow + r] = sum_{c in chunk} det(r,c) * Y[iy*ncol + c]
template <int N, bool OUTER>
__global__ void __launch_bounds__(kDetThreads, (N <= 10) ? 3 : 2)
det_kernel(const cplx *__restrict__ S, int ns, int n, const int32_t *__restrict__ rows, int64_t nrow,
           const int32_t *__restrict__ cols, int64_t ncol, int64_t chunk_len, const cplx *__restrict__ Y, int ny,
           cplx *__restrict__ out) {
    constexpr int GPW = 32 / N;                       // groups per warp
    constexpr int GPB = (kDetThreads / 32) * GPW;     // groups per block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / N;
    const bool in_group = g < GPW;
    const int l = in_group ? lane - g * N : 0;
    const int gbase = in_group ? g * N : lane;
    const int64_t r = (int64_t)blockIdx.x * GPB + warp * GPW + g;
    const bool rvalid = in_group && r < nrow;
    const bool real_row = rvalid && l < n;
    const int64_t c0 = (int64_t)blockIdx.y * chunk_len;
    int64_t c1 = c0 + chunk_len;
    if (c1 > ncol) c1 = ncol;

    const cplx *Srow = S;
    if (real_row) Srow = S + (int64_t)rows[r * n + l] * ns;

    constexpr int NYMAX = 4;
    cplx z[NYMAX];
#pragma unroll
    for (int q = 0; q < NYMAX; ++q) z[q] = make_cplx(0.0, 0.0);

    for (int64_t c = c0; c < c1; ++c) {
        cplx a[N];
        const int32_t *cl = cols + c * n;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            if (real_row && j < n) a[j] = ldg(&Srow[__ldg(&cl[j])]);
            else a[j] = make_cplx((j == l && !(real_row)) || (j == l && j >= n) ? 1.0 : 0.0, 0.0);
        }
        const cplx d = group_lu_det<N>(a, l, g, gbase);
        if (l == 0 && rvalid) {
            if (OUTER) {
                out[r * ncol + c] = d;
            } else {
#pragma unroll
                for (int q = 0; q < NYMAX; ++q)
                    if (q < ny) z[q] = z[q] + d * ldg(&Y[(int64_t)q * ncol + c]);
            }
        }
    }
    if (!OUTER && l == 0 && rvalid) {
#pragma unroll
        for (int q = 0; q < NYMAX; ++q)
            if (q < ny) out[((int64_t)blockIdx.y * ny + q) * nrow + r] = z[q];
    }
}

// Z[q*nrow + r] = sum_chunk Zp[(chunk*ny + q)*nrow + r]   (fixed order: lane-strided partial sums, then a
// shuffle tree).  One warp per output element: the sum over up to 4096 chunks is spread over 32 lanes
// instead of one thread walking it (32 us per call for the 117-row shapes before).
__global__ void __launch_bounds__(256) chunk_reduce_kernel(const cplx *Zp, int nchunk, int64_t len, cplx *Z) {
    Zp += (size_t)blockIdx.y * nchunk * len;                    // blockIdx.y = overlap of a stack
    Z += (size_t)blockIdx.y * len;
    const int lane = threadIdx.x & 31;
    const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < len; i += nwarp) {
        double sx = 0.0, sy = 0.0;
        for (int ch = lane; ch < nchunk; ch += 32) {
            const cplx v = Zp[(int64_t)ch * len + i];
            sx += v.x;
            sy += v.y;
        }
        sx = warp_sum(sx);
        sy = warp_sum(sy);
        if (lane == 0) Z[i] = make_cplx(sx, sy);
    }
}

// Antisymmetric completion folded into the amplitudes (aats.py:620-630 applied to the tensor
// is equivalent to applying it to the amplitude that multiplies it):
//   out[q*P + r] = sum over the 4 (i<->j, a<->b) images of (x - x.swapaxes(2,3))
//               = 2 ( x[i,j,a,b] - x[i,j,b,a] - x[j,i,a,b] + x[j,i,b,a] ),  r = (i,a,j,b), i<j, a<b
__global__ void __launch_bounds__(256)
pack_doubles_kernel(const cplx *__restrict__ x, int64_t xstride, int nq, int o, int v, int nf,
                    const int32_t *__restrict__ tab, int64_t P, cplx *__restrict__ out) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < P * nq;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx % P, q = idx / P;
        const int i = tab[4 * r] - nf, a = tab[4 * r + 1], j = tab[4 * r + 2] - nf, b = tab[4 * r + 3];
        const cplx *t = x + q * xstride;
        auto at = [&](int p0, int p1, int p2, int p3) { return t[(((int64_t)p0 * o + p1) * v + p2) * v + p3]; };
        const cplx s = at(i, j, a, b) - at(i, j, b, a) - at(j, i, a, b) + at(j, i, b, a);
        out[idx] = make_cplx(2.0 * s.x, 2.0 * s.y);
    }
}

// instantiated matrix sizes; n is padded up to the next one
static int padded_size(int n) {
    static const int sizes[] = {2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16, 20, 24, 28, 32};
    for (int s : sizes)
        if (n <= s) return s;
    return 32;
}
static int64_t groups_per_block(int n) { return (kDetThreads / 32) * (32 / padded_size(n)); }

template <bool OUTER>
static int launch_det(int n, dim3 grid, cudaStream_t st, const cplx *S, int ns, const int32_t *rows, int64_t nrow,
                      const int32_t *cols, int64_t ncol, int64_t chunk_len, const cplx *Y, int ny, cplx *out) {
#define APYIB_DET_CASE(NN)                                                                                         \
    case NN:                                                                                                       \
        det_kernel<NN, OUTER><<<grid, kDetThreads, 0, st>>>(S, ns, n, rows, nrow, cols, ncol, chunk_len, Y, ny, out); \
        break;
    switch (padded_size(n)) {
        APYIB_DET_CASE(2) APYIB_DET_CASE(3) APYIB_DET_CASE(4) APYIB_DET_CASE(5) APYIB_DET_CASE(6) APYIB_DET_CASE(7)
        APYIB_DET_CASE(8) APYIB_DET_CASE(9) APYIB_DET_CASE(10) APYIB_DET_CASE(12) APYIB_DET_CASE(14)
        APYIB_DET_CASE(16) APYIB_DET_CASE(20) APYIB_DET_CASE(24) APYIB_DET_CASE(28)
        default:
            det_kernel<32, OUTER><<<grid, kDetThreads, 0, st>>>(S, ns, n, rows, nrow, cols, ncol, chunk_len, Y, ny, out);
            break;
    }
#undef APYIB_DET_CASE
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

// thread-per-matrix LU (dets_tpm.cu), n <= 12
int launch_det_tpm(int n, cudaStream_t st, const cplx *S, int ns, const int32_t *rows, int64_t nrow,
                   const int32_t *cols, int64_t ncol, int64_t chunk_len, int64_t nchunk, const double *csign,
                   const int32_t *cindex, const cplx *Y, int ny, cplx *out, int outer, int nS, int64_t y_stride,
                   int64_t out_stride);
int tpm_total_warps(int n);
// prefix-shared LU (dets_pairs.cu), n <= 12, k = 1 or 2 trailing substituted columns
int launch_det_pairs(int n, int k, cudaStream_t st, const cplx *S, int ns, const int32_t *rows, int64_t nrow,
                     const int32_t *cols, int64_t ngroup, int64_t npair, const int32_t *cand, int nc, int64_t gchunk,
                     int64_t nchunk, const double *csign, const int32_t *cindex, const cplx *Y, int ny, int64_t ncol,
                     cplx *out, int nS, int64_t y_stride, int64_t out_stride);
int pairs_total_warps(int n, int k, int ns, int nc);
extern int g_pairs_variant;
constexpr int kTpmMaxN = 12;
static int g_det_kernel = 0;   // 0 = thread-per-matrix for 2 <= n <= 12, sub-warp above; 1 = sub-warp always
static bool use_tpm(int n) { return g_det_kernel == 0 && n >= 2 && n <= kTpmMaxN; }

// column chunks for a (row blocks) x (chunks) grid of about `target` CTAs, never more than one wave over
static int64_t chunks_for(int64_t rb, int64_t ncol, int64_t target, int64_t cap) {
    int64_t nchunk = target / rb;
    if (nchunk > ncol) nchunk = ncol;
    if (nchunk > cap) nchunk = cap;
    if (nchunk < 1) nchunk = 1;
    const int64_t chunk_len = (ncol + nchunk - 1) / nchunk;
    return (ncol + chunk_len - 1) / chunk_len;
}

}  // namespace apyib

using namespace apyib;

extern "C" int apyib_det_set_kernel(int which) {
    APYIB_REQUIRE(which == 0 || which == 1, "0 = auto (thread-per-matrix for n <= 12), 1 = sub-warp LU");
    g_det_kernel = which;
    return APYIB_OK;
}

static int det_outer_impl(const void *d_S, int ns, int n, const int32_t *d_rows, int64_t nrow, const int32_t *d_cols,
                          int64_t ncol, const double *d_csign, const int32_t *d_cindex, void *d_out, void *stream,
                          int nS = 1) {
    APYIB_REQUIRE(d_S && d_rows && d_cols && d_out, "null pointer");
    APYIB_REQUIRE(n >= 1 && n <= 32 && ns >= n, "1 <= n <= 32 supported by the sub-warp LU");
    APYIB_REQUIRE(nS >= 1 && nS <= 65535, "stack size");
    if (nrow == 0 || ncol == 0) return APYIB_OK;
    if (use_tpm(n)) {
        // a stack shares the device: size the chunks so that ALL overlaps together fill about one wave
        const int64_t target = tpm_total_warps(n) / nS > 0 ? tpm_total_warps(n) / nS : 1;
        const int64_t nchunk = chunks_for((nrow + 31) / 32, ncol, target, 1 << 20);
        const int64_t chunk_len = (ncol + nchunk - 1) / nchunk;
        return launch_det_tpm(n, (cudaStream_t)stream, (const cplx *)d_S, ns, d_rows, nrow, d_cols, ncol, chunk_len,
                              nchunk, d_csign, d_cindex, nullptr, 0, (cplx *)d_out, 1, nS, 0, nrow * ncol);
    }
    if (nS > 1) {                                   // sub-warp kernel: one launch per overlap of the stack
        for (int s = 0; s < nS; ++s) {
            const int rc = det_outer_impl((const cplx *)d_S + (size_t)s * ns * ns, ns, n, d_rows, nrow, d_cols, ncol, d_csign,
                                          d_cindex, (cplx *)d_out + (size_t)s * nrow * ncol, stream, 1);
            if (rc != APYIB_OK) return rc;
        }
        return APYIB_OK;
    }
    if (d_csign || d_cindex) {
        set_error("sorted column lists need the thread-per-matrix LU kernel (2 <= n <= 12)");
        return APYIB_ERR_UNSUPPORTED;
    }
    const int64_t gpb = groups_per_block(n);
    const int64_t rb = (nrow + gpb - 1) / gpb;
    int64_t nchunk = (148 * 8 + rb - 1) / rb;
    if (nchunk > ncol) nchunk = ncol;
    if (nchunk > 65535) nchunk = 65535;
    if (nchunk < 1) nchunk = 1;
    const int64_t chunk_len = (ncol + nchunk - 1) / nchunk;
    nchunk = (ncol + chunk_len - 1) / chunk_len;
    APYIB_REQUIRE(rb <= 2147483647LL, "too many rows");
    dim3 grid((unsigned)rb, (unsigned)nchunk);
    return launch_det<true>(n, grid, (cudaStream_t)stream, (const cplx *)d_S, ns, d_rows, nrow, d_cols, ncol,
                            chunk_len, nullptr, 0, (cplx *)d_out);
}

// Z[iy*nrow + r] = sum_c det(S[rows[r], cols[c]]) * Y[iy*ncol + c]
// d_work: scratch of apyib_det_matvec_work_len(nrow, ncol, ny, n) complex128 elements.
extern "C" int64_t apyib_det_matvec_nchunk(int64_t nrow, int64_t ncol, int n) {
    if (use_tpm(n)) return chunks_for((nrow + 31) / 32, ncol, tpm_total_warps(n), 4096);
    const int64_t gpb = groups_per_block(n);
    const int64_t rb = (nrow + gpb - 1) / gpb;
    int64_t nchunk = (148 * 8 + rb - 1) / rb;
    if (nchunk > ncol) nchunk = ncol;
    if (nchunk > 4096) nchunk = 4096;
    if (nchunk < 1) nchunk = 1;
    const int64_t chunk_len = (ncol + nchunk - 1) / nchunk;
    return (ncol + chunk_len - 1) / chunk_len;
}

extern "C" int64_t apyib_det_matvec_work_len(int64_t nrow, int64_t ncol, int ny, int n) {
    return apyib_det_matvec_nchunk(nrow, ncol, n) * ny * nrow;
}

static int det_matvec_impl(const void *d_S, int ns, int n, const int32_t *d_rows, int64_t nrow, const int32_t *d_cols,
                           int64_t ncol, const double *d_csign, const int32_t *d_cindex, const void *d_Y, int ny,
                           void *d_Z, void *d_work, void *stream, int nS = 1, int64_t y_stride = 0) {
    APYIB_REQUIRE(d_S && d_rows && d_cols && d_Y && d_Z && d_work, "null pointer");
    if ((d_csign || d_cindex) && !use_tpm(n)) {
        set_error("sorted column lists need the thread-per-matrix LU kernel (2 <= n <= 12)");
        return APYIB_ERR_UNSUPPORTED;
    }
    APYIB_REQUIRE(n >= 1 && n <= 32 && ns >= n, "1 <= n <= 32 supported by the sub-warp LU");
    APYIB_REQUIRE(ny >= 1 && ny <= 4, "1 <= ny <= 4");
    APYIB_REQUIRE(nS >= 1 && nS <= 65535, "stack size");
    if (nrow == 0) return APYIB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (ncol == 0) {
        APYIB_CUDA_CHECK(cudaMemsetAsync(d_Z, 0, sizeof(cplx) * ny * nrow * nS, st));
        return APYIB_OK;
    }
    const int64_t gpb = groups_per_block(n);
    const int64_t rb = (nrow + gpb - 1) / gpb;
    const int64_t nchunk = apyib_det_matvec_nchunk(nrow, ncol, n);
    const int64_t chunk_len = (ncol + nchunk - 1) / nchunk;
    const int64_t len = (int64_t)ny * nrow;
    if (!use_tpm(n) && nS > 1) {                    // sub-warp kernel: one pass per overlap of the stack
        for (int s = 0; s < nS; ++s) {
            const int rc = det_matvec_impl((const cplx *)d_S + (size_t)s * ns * ns, ns, n, d_rows, nrow, d_cols, ncol, d_csign,
                                           d_cindex, (const cplx *)d_Y + (size_t)s * y_stride, ny,
                                           (cplx *)d_Z + (size_t)s * len, d_work, stream, 1, 0);
            if (rc != APYIB_OK) return rc;
        }
        return APYIB_OK;
    }
    dim3 grid((unsigned)rb, (unsigned)nchunk);
    int rc = use_tpm(n) ? launch_det_tpm(n, st, (const cplx *)d_S, ns, d_rows, nrow, d_cols, ncol, chunk_len, nchunk,
                                         d_csign, d_cindex, (const cplx *)d_Y, ny, (cplx *)d_work, 0, nS, y_stride,
                                         nchunk * len)
                        : launch_det<false>(n, grid, st, (const cplx *)d_S, ns, d_rows, nrow, d_cols, ncol, chunk_len,
                               (const cplx *)d_Y, ny, (cplx *)d_work);
    if (rc != APYIB_OK) return rc;
    int64_t b = (len + 7) / 8;                       // one warp per output element, 8 warps per block
    if (b > 148 * 8) b = 148 * 8;
    chunk_reduce_kernel<<<dim3((unsigned)b, (unsigned)nS), 256, 0, st>>>((const cplx *)d_work, (int)nchunk, len, (cplx *)d_Z);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_det_outer(const void *d_S, int ns, int n, const int32_t *d_rows, int64_t nrow,
                               const int32_t *d_cols, int64_t ncol, void *d_out, void *stream) {
    return det_outer_impl(d_S, ns, n, d_rows, nrow, d_cols, ncol, nullptr, nullptr, d_out, stream);
}

extern "C" int apyib_det_matvec(const void *d_S, int ns, int n, const int32_t *d_rows, int64_t nrow,
                                const int32_t *d_cols, int64_t ncol, const void *d_Y, int ny, void *d_Z,
                                void *d_work, void *stream) {
    return det_matvec_impl(d_S, ns, n, d_rows, nrow, d_cols, ncol, nullptr, nullptr, d_Y, ny, d_Z, d_work, stream);
}

extern "C" int apyib_det_outer_sorted(const void *d_S, int ns, int n, const int32_t *d_rows, int64_t nrow,
                                      const int32_t *d_cols, const double *d_col_sign, const int32_t *d_col_index,
                                      int64_t ncol, void *d_out, void *stream) {
    APYIB_REQUIRE(d_col_sign && d_col_index, "null pointer");
    return det_outer_impl(d_S, ns, n, d_rows, nrow, d_cols, ncol, d_col_sign, d_col_index, d_out, stream);
}

extern "C" int apyib_det_matvec_sorted(const void *d_S, int ns, int n, const int32_t *d_rows, int64_t nrow,
                                       const int32_t *d_cols, const double *d_col_sign, const int32_t *d_col_index,
                                       int64_t ncol, const void *d_Y, int ny, void *d_Z, void *d_work, void *stream) {
    APYIB_REQUIRE(d_col_sign && d_col_index, "null pointer");
    return det_matvec_impl(d_S, ns, n, d_rows, nrow, d_cols, ncol, d_col_sign, d_col_index, d_Y, ny, d_Z, d_work, stream);
}

static int64_t pairs_nchunk(int64_t nrow, int64_t ngroup, int n, int k, int ns, int nc) {
    const int w = pairs_total_warps(n, k, ns, nc);
    if (w <= 0) return 0;
    return chunks_for((nrow + 31) / 32, ngroup, w, 4096);
}

// per overlap: nchunk x ny x nrow partial sums + ny x ncol sorted, sign-folded amplitude entries (a stack of nS
// overlaps needs nS times this; the partial sums of all overlaps come first, then the amplitude copies)
extern "C" int64_t apyib_det_matvec_pairs_work_len(int64_t nrow, int64_t ngroup, int ny, int n, int k, int ns, int nc) {
    const int64_t c = pairs_nchunk(nrow, ngroup, n, k, ns, nc);
    const int64_t group_len = (k == 1) ? (int64_t)nc : (int64_t)nc * (nc - 1) / 2;
    return (c > 0 ? c : 1) * ny * nrow + (int64_t)ny * ngroup * group_len;
}

// Ys[s][q][c] = sign[c] * Y[s][q][index[c]]: the amplitude vector in the order (and with the signs) of the sorted
// column lists, so that the pair loop of the prefix-shared LU kernel reads ONE warp-uniform, contiguous entry per
// determinant instead of sign + index + a gathered Y entry.
__global__ void __launch_bounds__(256) permute_y_kernel(const cplx *__restrict__ Y, int64_t y_stride, const double *__restrict__ sign,
                                                        const int32_t *__restrict__ index, int64_t ncol, int ny,
                                                        cplx *__restrict__ Ys) {
    Y += (size_t)blockIdx.y * y_stride;
    Ys += (size_t)blockIdx.y * ny * ncol;
    const int64_t total = (int64_t)ny * ncol;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = e / ncol, c = e - q * ncol;
        const cplx v = Y[q * ncol + index[c]];
        const double sg = sign[c];
        Ys[e] = make_cplx(sg * v.x, sg * v.y);
    }
}

extern "C" int apyib_det_set_pairs_variant(int which) {
    APYIB_REQUIRE(which == 0 || which == 1, "0 = generic (1..4 vectors, predicated), 1 = + single-vector specialisation");
    g_pairs_variant = which;
    return APYIB_OK;
}

static int det_matvec_pairs_impl(const void *d_S, int nS, int ns, int n, int k, const int32_t *d_rows, int64_t nrow,
                                 const int32_t *d_cols_sorted, const double *d_col_sign, const int32_t *d_col_index,
                                 int64_t ncol, int64_t group_len, const int32_t *d_cand, int nc, const void *d_Y,
                                 int64_t y_stride, int ny, void *d_Z, void *d_work, void *stream) {
    APYIB_REQUIRE(d_S && d_rows && d_cols_sorted && d_col_sign && d_col_index && d_cand && d_Y && d_Z && d_work, "null pointer");
    APYIB_REQUIRE(ny >= 1 && ny <= 4, "1 <= ny <= 4");
    APYIB_REQUIRE(nS >= 1 && nS <= 65535, "stack size");
    APYIB_REQUIRE((k == 1 || k == 2) && n > k && ns >= n && nc >= k, "sizes");
    APYIB_REQUIRE(group_len == (k == 1 ? (int64_t)nc : (int64_t)nc * (nc - 1) / 2), "group_len must be nc (k=1) or C(nc,2) (k=2)");
    APYIB_REQUIRE(nrow >= 1 && ncol >= group_len && ncol % group_len == 0, "ncol must be a whole number of groups");
    if (n > kTpmMaxN || g_det_kernel != 0) {
        set_error("prefix-shared LU needs 2 <= n <= 12 and the thread-per-matrix kernel family");
        return APYIB_ERR_UNSUPPORTED;
    }
    const int64_t ngroup = ncol / group_len;
    const int64_t nchunk = pairs_nchunk(nrow, ngroup, n, k, ns, nc);
    if (nchunk <= 0) {
        set_error("prefix-shared LU: (n, k) = (%d, %d) with %d candidate columns does not fit", n, k, nc);
        return APYIB_ERR_UNSUPPORTED;
    }
    const int64_t gchunk = (ngroup + nchunk - 1) / nchunk;
    const int64_t len = (int64_t)ny * nrow;
    cudaStream_t st = (cudaStream_t)stream;
    // amplitude vectors in sorted-list order with the list signs folded in (one copy per overlap, or one for all)
    cplx *Ys = (cplx *)d_work + (int64_t)nS * nchunk * len;
    const int nYs = y_stride == 0 ? 1 : nS;
    {
        int64_t pb = ((int64_t)ny * ncol + 255) / 256;
        if (pb > 148 * 4) pb = 148 * 4;
        permute_y_kernel<<<dim3((unsigned)pb, (unsigned)nYs), 256, 0, st>>>((const cplx *)d_Y, y_stride, d_col_sign, d_col_index,
                                                                           ncol, ny, Ys);
        APYIB_LAUNCH_CHECK();
    }
    int rc = launch_det_pairs(n, k, st, (const cplx *)d_S, ns, d_rows, nrow, d_cols_sorted, ngroup, group_len, d_cand, nc,
                              gchunk, nchunk, nullptr, nullptr, Ys, ny, ncol, (cplx *)d_work, nS,
                              y_stride == 0 ? 0 : (int64_t)ny * ncol, nchunk * len);
    if (rc != APYIB_OK) return rc;
    int64_t b = (len + 7) / 8;
    if (b > 148 * 8) b = 148 * 8;
    chunk_reduce_kernel<<<dim3((unsigned)b, (unsigned)nS), 256, 0, st>>>((const cplx *)d_work, (int)nchunk, len, (cplx *)d_Z);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_det_matvec_pairs(const void *d_S, int ns, int n, int k, const int32_t *d_rows, int64_t nrow,
                                      const int32_t *d_cols_sorted, const double *d_col_sign,
                                      const int32_t *d_col_index, int64_t ncol, int64_t group_len,
                                      const int32_t *d_cand, int nc, const void *d_Y, int ny, void *d_Z, void *d_work,
                                      void *stream) {
    return det_matvec_pairs_impl(d_S, 1, ns, n, k, d_rows, nrow, d_cols_sorted, d_col_sign, d_col_index, ncol, group_len,
                                 d_cand, nc, d_Y, 0, ny, d_Z, d_work, stream);
}

// ---- stacks of overlaps: the same index lists applied to nS overlap matrices in ONE launch (grid.y = overlap) ----
// d_S: [nS][ns][ns]; outputs [nS][...] contiguous; d_Y: overlap s reads d_Y + s * y_stride (0 = one Y for all);
// d_work: nS x the single-overlap work length.  d_col_sign / d_col_index may be NULL (plain lists).
extern "C" int apyib_det_outer_stack(const void *d_S, int nS, int ns, int n, const int32_t *d_rows, int64_t nrow,
                                     const int32_t *d_cols, const double *d_col_sign, const int32_t *d_col_index,
                                     int64_t ncol, void *d_out, void *stream) {
    APYIB_REQUIRE((d_col_sign == nullptr) == (d_col_index == nullptr), "sign and index come together");
    return det_outer_impl(d_S, ns, n, d_rows, nrow, d_cols, ncol, d_col_sign, d_col_index, d_out, stream, nS);
}

extern "C" int apyib_det_matvec_stack(const void *d_S, int nS, int ns, int n, const int32_t *d_rows, int64_t nrow,
                                      const int32_t *d_cols, const double *d_col_sign, const int32_t *d_col_index,
                                      int64_t ncol, const void *d_Y, int64_t y_stride, int ny, void *d_Z, void *d_work,
                                      void *stream) {
    APYIB_REQUIRE((d_col_sign == nullptr) == (d_col_index == nullptr), "sign and index come together");
    return det_matvec_impl(d_S, ns, n, d_rows, nrow, d_cols, ncol, d_col_sign, d_col_index, d_Y, ny, d_Z, d_work, stream, nS,
                           y_stride);
}

extern "C" int apyib_det_matvec_pairs_stack(const void *d_S, int nS, int ns, int n, int k, const int32_t *d_rows,
                                            int64_t nrow, const int32_t *d_cols_sorted, const double *d_col_sign,
                                            const int32_t *d_col_index, int64_t ncol, int64_t group_len,
                                            const int32_t *d_cand, int nc, const void *d_Y, int64_t y_stride, int ny,
                                            void *d_Z, void *d_work, void *stream) {
    return det_matvec_pairs_impl(d_S, nS, ns, n, k, d_rows, nrow, d_cols_sorted, d_col_sign, d_col_index, ncol, group_len,
                                 d_cand, nc, d_Y, y_stride, ny, d_Z, d_work, stream);
}

// Re-orders column index lists for factorisation reuse (host): inside every list the substituted entries
// (value >= n) move to the end, order preserved, and the lists are sorted lexicographically, so consecutive
// lists share the longest possible leading part.  sign[c] = parity of the in-list move (det of the original
// list = sign * det of the re-ordered one); index[c] = position of sorted list c in the input enumeration.
extern "C" int apyib_det_sort_lists(int n, const int32_t *h_lists, int64_t count, int32_t *h_sorted, double *h_sign,
                                    int32_t *h_index) {
    APYIB_REQUIRE(n >= 1 && count >= 0 && count < 2147483647LL && (count == 0 || (h_lists && h_sorted && h_sign && h_index)),
                  "arguments");
    std::vector<int32_t> tmp((size_t)count * n);
    std::vector<double> sg((size_t)count);
    for (int64_t q = 0; q < count; ++q) {
        const int32_t *in = h_lists + q * n;
        int32_t *o = tmp.data() + q * n;
        int w = 0, inversions = 0, nsub = 0;
        for (int j = 0; j < n; ++j) {
            if (in[j] < n) { o[w++] = in[j]; inversions += nsub; } else { ++nsub; }
        }
        for (int j = 0; j < n; ++j)
            if (in[j] >= n) o[w++] = in[j];
        sg[q] = (inversions & 1) ? -1.0 : 1.0;
    }
    std::vector<int32_t> order((size_t)count);
    for (int64_t q = 0; q < count; ++q) order[q] = (int32_t)q;
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
        return std::lexicographical_compare(tmp.begin() + (size_t)a * n, tmp.begin() + (size_t)(a + 1) * n,
                                            tmp.begin() + (size_t)b * n, tmp.begin() + (size_t)(b + 1) * n);
    });
    for (int64_t c = 0; c < count; ++c) {
        const int32_t q = order[c];
        for (int j = 0; j < n; ++j) h_sorted[c * n + j] = tmp[(size_t)q * n + j];
        h_sign[c] = sg[q];
        h_index[c] = q;
    }
    return APYIB_OK;
}

extern "C" int apyib_pack_doubles(const void *d_x, int64_t x_stride, int nq, int o, int v, int nf,
                                  const int32_t *d_doubles, int64_t P, void *d_out, void *stream) {
    APYIB_REQUIRE(d_x && d_doubles && d_out, "null pointer");
    APYIB_REQUIRE(nq >= 1 && o >= 0 && v >= 0 && nf >= 0, "sizes");
    if (P == 0) return APYIB_OK;
    int64_t b = (P * nq + 255) / 256;
    if (b > 148 * 8) b = 148 * 8;
    pack_doubles_kernel<<<(unsigned)b, 256, 0, (cudaStream_t)stream>>>((const cplx *)d_x, x_stride, nq, o, v, nf,
                                                                     d_doubles, P, (cplx *)d_out);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

// ---- host-side bit-exact index tables -------------------------------------------------------
extern "C" int apyib_get_slices(int nbf, int ndocc, int nfzc, int spin_orbital, int32_t b[16]) {
    APYIB_REQUIRE(b != nullptr && nbf >= ndocc && ndocc >= nfzc && nfzc >= 0, "sizes");
    // C_list (utils.py:192-197)
    b[0] = 0; b[1] = nfzc; b[2] = nfzc; b[3] = ndocc; b[4] = ndocc; b[5] = nbf; b[6] = nfzc; b[7] = nbf;
    const int s = spin_orbital ? 2 : 1;   // utils.py:200-209
    b[8] = 0; b[9] = s * nfzc;
    b[10] = 0; b[11] = s * ndocc - s * nfzc;
    b[12] = s * ndocc - s * nfzc; b[13] = s * nbf - s * nfzc;
    b[14] = 0; b[15] = s * nbf - s * nfzc;
    return APYIB_OK;
}

extern "C" int apyib_det_enumeration(int ndocc, int nfzc, int nvirt, int32_t *h_singles, int64_t *n_singles,
                                     int32_t *h_doubles, int64_t *n_doubles) {
    APYIB_REQUIRE(ndocc >= nfzc && nfzc >= 0 && nvirt >= 0, "sizes");
    int64_t ns = 0, nd = 0;
    for (int i = nfzc; i < ndocc; ++i)
        for (int a = 0; a < nvirt; ++a) {
            if (h_singles) { h_singles[2 * ns] = i; h_singles[2 * ns + 1] = a; }
            ++ns;
            for (int j = i + 1; j < ndocc; ++j)
                for (int bb = a + 1; bb < nvirt; ++bb) {
                    if (h_doubles) {
                        h_doubles[4 * nd] = i; h_doubles[4 * nd + 1] = a;
                        h_doubles[4 * nd + 2] = j; h_doubles[4 * nd + 3] = bb;
                    }
                    ++nd;
                }
        }
    if (n_singles) *n_singles = ns;
    if (n_doubles) *n_doubles = nd;
    return APYIB_OK;
}

extern "C" int apyib_det_index_lists(int n, const int32_t *h_sub, int64_t count, int nsub, int32_t *h_out) {
    APYIB_REQUIRE(h_out && (h_sub || nsub == 0) && n >= 1 && nsub >= 0, "arguments");
    for (int64_t q = 0; q < count; ++q) {
        int32_t *o = h_out + q * n;
        for (int j = 0; j < n; ++j) o[j] = j;
        for (int t = 0; t < nsub; ++t) {
            const int32_t pos = h_sub[(q * nsub + t) * 2], vir = h_sub[(q * nsub + t) * 2 + 1];
            APYIB_REQUIRE(pos >= 0 && pos < n && vir >= 0, "substitution out of range");
            o[pos] = vir + n;
        }
    }
    return APYIB_OK;
}

extern "C" int apyib_so_index_lists(int nso, int nocc, const int32_t *h_sub, int64_t count, int nsub,
                                    int32_t *h_out) {
    APYIB_REQUIRE(h_out && (h_sub || nsub == 0) && nso >= nocc && nocc >= 1 && nso <= 4096, "arguments");
    int32_t perm[4096];
    for (int64_t q = 0; q < count; ++q) {
        for (int j = 0; j < nso; ++j) perm[j] = j;
        for (int t = 0; t < nsub; ++t) {   // sequential pair swaps, aats.py:123-126
            const int32_t x = h_sub[(q * nsub + t) * 2], y = h_sub[(q * nsub + t) * 2 + 1];
            APYIB_REQUIRE(x >= 0 && x < nso && y >= 0 && y < nso, "swap index out of range");
            const int32_t tmp = perm[x]; perm[x] = perm[y]; perm[y] = tmp;
        }
        for (int j = 0; j < nocc; ++j) h_out[q * nocc + j] = perm[j];
    }
    return APYIB_OK;
}
