// Determinant-lemma path for the substituted occupied-overlap determinants (SURVEY.md 8(f).1,
// App. B.5).  With A = S_oo, P = S_vo A^-1, Q = A^-1 S_ov, R = S_vv - S_vo A^-1 S_ov every
// determinant of aats.py:581-618 with r <= 2 substituted rows {(i_m -> a_m)} and c <= 2 substituted
// columns {(k_n -> c_n)} is
//
//      det = det(A) * (-1)^c * det [ P[a_m, i_m']   R[a_m, c_n'] ]        (r+c) x (r+c) <= 4 x 4
//                                  [ A^-1[k_n, i_m']  -Q[k_n, c_n'] ]
//
// (verified against the batched-LU kernel in tests/test_gpu_aat.py).  A = S_oo is close to the unit
// matrix for finite-difference overlaps, so the inverse is benign; the small determinants are
// evaluated division-free by cofactor expansion, one thread per (row-substitution,
// column-substitution) pair: no shuffles, no pivot search, ~7x fewer flops than the 9x9 LU and
// ~20x fewer than the 16x16 one.  The fused variant accumulates Z[q,r] = sum_c det(r,c) Y[q,c]
// exactly like apyib_det_matvec, so the determinant tables still never exist in memory.
#include "common.cuh"

namespace apyib {

constexpr int kLemmaThreads = 128;

// ---- prepare: det(A), A^-1, P, Q, R for a stack of overlaps (one CTA per overlap) -------------
// out (per overlap, complex): [0] det(A) | Ainv[no*no] | P[nv*no] | Q[no*nv] | R[nv*nv]
__global__ void __launch_bounds__(256) lemma_prepare_kernel(const cplx *__restrict__ S_all, int ns, int no,
                                                            cplx *__restrict__ out_all, int64_t out_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *W = reinterpret_cast<cplx *>(smem_raw);         // [no][2*no] augmented (A | 1)
    __shared__ int piv_row;
    __shared__ double det_re, det_im;
    const cplx *S = S_all + (int64_t)blockIdx.x * ns * ns;
    cplx *out = out_all + (int64_t)blockIdx.x * out_stride;
    const int nv = ns - no, w = 2 * no, tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < no * w; e += nt) {
        const int i = e / w, j = e % w;
        W[e] = (j < no) ? S[(int64_t)i * ns + j] : make_cplx(j - no == i ? 1.0 : 0.0, 0.0);
    }
    if (tid == 0) { det_re = 1.0; det_im = 0.0; }
    __syncthreads();
    for (int k = 0; k < no; ++k) {
        if (tid == 0) {   // partial pivoting (|re|+|im|), tiny serial scan
            int best = k;
            double bm = fabs(W[k * w + k].x) + fabs(W[k * w + k].y);
            for (int i = k + 1; i < no; ++i) {
                const double m = fabs(W[i * w + k].x) + fabs(W[i * w + k].y);
                if (m > bm) { bm = m; best = i; }
            }
            piv_row = best;
        }
        __syncthreads();
        const int pr = piv_row;
        if (pr != k) {
            for (int j = tid; j < w; j += nt) {
                const cplx t = W[k * w + j];
                W[k * w + j] = W[pr * w + j];
                W[pr * w + j] = t;
            }
        }
        __syncthreads();
        const cplx pv = W[k * w + k];
        if (tid == 0) {
            const cplx d = make_cplx(det_re, det_im) * pv;
            const double sgn = (pr != k) ? -1.0 : 1.0;
            det_re = sgn * d.x;
            det_im = sgn * d.y;
        }
        __syncthreads();
        const double d2 = pv.x * pv.x + pv.y * pv.y;
        const cplx inv = (d2 > 0.0) ? make_cplx(pv.x / d2, -pv.y / d2) : make_cplx(0.0, 0.0);
        for (int j = tid; j < w; j += nt) W[k * w + j] = W[k * w + j] * inv;
        __syncthreads();
        // eliminate column k from every other row; factors are read before anyone overwrites column k
        for (int e = tid; e < no * w; e += nt) {
            const int i = e / w, j = e % w;
            if (i == k || j == k) continue;
            W[e] = W[e] - W[i * w + k] * W[k * w + j];
        }
        __syncthreads();
        for (int i = tid; i < no; i += nt)
            if (i != k) W[i * w + k] = make_cplx(0.0, 0.0);
        __syncthreads();
    }
    cplx *Ainv = out + 1, *P = Ainv + no * no, *Q = P + nv * no, *R = Q + no * nv;
    if (tid == 0) out[0] = make_cplx(det_re, det_im);
    for (int e = tid; e < no * no; e += nt) Ainv[e] = W[(e / no) * w + no + (e % no)];
    for (int e = tid; e < nv * no; e += nt) {          // P[a,i] = sum_j S[no+a, j] Ainv[j,i]
        const int a = e / no, i = e % no;
        cplx s = make_cplx(0.0, 0.0);
        for (int j = 0; j < no; ++j) s = s + S[(int64_t)(no + a) * ns + j] * W[j * w + no + i];
        P[e] = s;
    }
    for (int e = tid; e < no * nv; e += nt) {          // Q[k,c] = sum_j Ainv[k,j] S[j, no+c]
        const int k = e / nv, c = e % nv;
        cplx s = make_cplx(0.0, 0.0);
        for (int j = 0; j < no; ++j) s = s + W[k * w + no + j] * S[(int64_t)j * ns + no + c];
        Q[e] = s;
    }
    __syncthreads();
    __threadfence_block();
    for (int e = tid; e < nv * nv; e += nt) {          // R[a,c] = S[no+a,no+c] - sum_j S[no+a,j] Q[j,c]
        const int a = e / nv, c = e % nv;
        cplx s = S[(int64_t)(no + a) * ns + no + c];
        for (int j = 0; j < no; ++j) s = s - S[(int64_t)(no + a) * ns + j] * Q[j * nv + c];
        R[e] = s;
    }
}

// ---- small determinants, division free ----------------------------------------------------------
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_cplx(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ cplx minor2(cplx a, cplx b, cplx c, cplx d) {   // a d - b c
    return make_cplx(fma(a.x, d.x, -a.y * d.y) - fma(b.x, c.x, -b.y * c.y),
                     fma(a.x, d.y, a.y * d.x) - fma(b.x, c.y, b.y * c.x));
}
template <int K> __device__ __forceinline__ cplx det_small(const cplx (&t)[4][4]) {
    if constexpr (K == 0) return make_cplx(1.0, 0.0);
    if constexpr (K == 1) return t[0][0];
    if constexpr (K == 2) return minor2(t[0][0], t[0][1], t[1][0], t[1][1]);
    if constexpr (K == 3) {
        const cplx m0 = minor2(t[1][1], t[1][2], t[2][1], t[2][2]);
        const cplx m1 = minor2(t[1][0], t[1][2], t[2][0], t[2][2]);
        const cplx m2 = minor2(t[1][0], t[1][1], t[2][0], t[2][1]);
        return cmul(t[0][0], m0) - cmul(t[0][1], m1) + cmul(t[0][2], m2);
    }
    if constexpr (K == 4) {   // Laplace expansion along rows (0,1) x (2,3)
        const cplx a01 = minor2(t[0][0], t[0][1], t[1][0], t[1][1]), b23 = minor2(t[2][2], t[2][3], t[3][2], t[3][3]);
        const cplx a02 = minor2(t[0][0], t[0][2], t[1][0], t[1][2]), b13 = minor2(t[2][1], t[2][3], t[3][1], t[3][3]);
        const cplx a03 = minor2(t[0][0], t[0][3], t[1][0], t[1][3]), b12 = minor2(t[2][1], t[2][2], t[3][1], t[3][2]);
        const cplx a12 = minor2(t[0][1], t[0][2], t[1][1], t[1][2]), b03 = minor2(t[2][0], t[2][3], t[3][0], t[3][3]);
        const cplx a13 = minor2(t[0][1], t[0][3], t[1][1], t[1][3]), b02 = minor2(t[2][0], t[2][2], t[3][0], t[3][2]);
        const cplx a23 = minor2(t[0][2], t[0][3], t[1][2], t[1][3]), b01 = minor2(t[2][0], t[2][1], t[3][0], t[3][1]);
        return cmul(a01, b23) - cmul(a02, b13) + cmul(a03, b12) + cmul(a12, b03) - cmul(a13, b02) + cmul(a23, b01);
    }
    return make_cplx(0.0, 0.0);
}

// rsub: (nrow x 2*RK) int32 (i, a[, j, b]); csub: (ncol x 2*CK) (k, c[, l, d]); occupied indices are
// full-space (0..no), virtual indices zero-based.  grid: x = row blocks, y = column chunks, z = overlap.
// OUTER : out[(s*nrow + r)*ncol + c] = det
// !OUTER: out[((s*nchunk + chunk)*ny + q)*nrow + r] = sum_{c in chunk} det(r,c) * Y[q*ncol + c]
template <int RK, int CK, bool OUTER>
__global__ void __launch_bounds__(kLemmaThreads)
lemma_kernel(const cplx *__restrict__ prep_all, int64_t prep_stride, int no, int nv,
             const int32_t *__restrict__ rsub, int64_t nrow, const int32_t *__restrict__ csub, int64_t ncol,
             int64_t chunk_len, const cplx *__restrict__ Y_all, int64_t y_sstride, int ny, cplx *__restrict__ out) {
    constexpr int K = RK + CK;
    const cplx *Y = OUTER ? nullptr : Y_all + (int64_t)blockIdx.z * y_sstride;
    const cplx *prep = prep_all + (int64_t)blockIdx.z * prep_stride;
    const cplx detA = prep[0];
    const cplx *Ainv = prep + 1, *P = Ainv + no * no, *Q = P + nv * no, *R = Q + no * nv;
    const int64_t r = (int64_t)blockIdx.x * kLemmaThreads + threadIdx.x;
    const bool rvalid = r < nrow;
    const int64_t c0 = (int64_t)blockIdx.y * chunk_len;
    int64_t c1 = c0 + chunk_len;
    if (c1 > ncol) c1 = ncol;
    int ri[RK > 0 ? RK : 1], ra[RK > 0 ? RK : 1];
    cplx t[4][4];
#pragma unroll
    for (int m = 0; m < RK; ++m) {
        ri[m] = rvalid ? rsub[r * 2 * RK + 2 * m] : 0;
        ra[m] = rvalid ? rsub[r * 2 * RK + 2 * m + 1] : 0;
    }
#pragma unroll
    for (int m = 0; m < RK; ++m)
#pragma unroll
        for (int mp = 0; mp < RK; ++mp) t[m][mp] = ldg(&P[ra[m] * no + ri[mp]]);
    constexpr int NYMAX = 4;
    cplx z[NYMAX];
#pragma unroll
    for (int q = 0; q < NYMAX; ++q) z[q] = make_cplx(0.0, 0.0);
    const double sgn = (CK & 1) ? -1.0 : 1.0;
    for (int64_t c = c0; c < c1; ++c) {
        int ck[CK > 0 ? CK : 1], cc[CK > 0 ? CK : 1];
#pragma unroll
        for (int n = 0; n < CK; ++n) {
            ck[n] = __ldg(&csub[c * 2 * CK + 2 * n]);
            cc[n] = __ldg(&csub[c * 2 * CK + 2 * n + 1]);
        }
#pragma unroll
        for (int n = 0; n < CK; ++n) {
#pragma unroll
            for (int m = 0; m < RK; ++m) {
                t[m][RK + n] = ldg(&R[ra[m] * nv + cc[n]]);
                t[RK + n][m] = ldg(&Ainv[ck[n] * no + ri[m]]);
            }
#pragma unroll
            for (int n2 = 0; n2 < CK; ++n2) {
                const cplx q = ldg(&Q[ck[n] * nv + cc[n2]]);
                t[RK + n][RK + n2] = make_cplx(-q.x, -q.y);
            }
        }
        cplx d = cmul(detA, det_small<K>(t));
        d = make_cplx(sgn * d.x, sgn * d.y);
        if (rvalid) {
            if (OUTER) {
                out[((int64_t)blockIdx.z * nrow + r) * ncol + c] = d;
            } else {
#pragma unroll
                for (int q = 0; q < NYMAX; ++q)
                    if (q < ny) z[q] = z[q] + cmul(d, ldg(&Y[(int64_t)q * ncol + c]));
            }
        }
    }
    if (!OUTER && rvalid) {
#pragma unroll
        for (int q = 0; q < NYMAX; ++q)
            if (q < ny) out[(((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * ny + q) * nrow + r] = z[q];
    }
}

// Z[(s*ny + q)*nrow + r] = sum_chunk Zp[((s*nchunk + chunk)*ny + q)*nrow + r]
__global__ void __launch_bounds__(256)
lemma_chunk_reduce_kernel(const cplx *Zp, int nchunk, int64_t len, int64_t total, cplx *Z) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = i / len, e = i % len;
        cplx acc = make_cplx(0.0, 0.0);
        for (int ch = 0; ch < nchunk; ++ch) acc = acc + Zp[((int64_t)s * nchunk + ch) * len + e];
        Z[i] = acc;
    }
}

static int64_t lemma_nchunk(int64_t nrow, int64_t ncol, int nS) {
    const int64_t rb = (nrow + kLemmaThreads - 1) / kLemmaThreads;
    int64_t nchunk = (148 * 8 + rb * nS - 1) / (rb * nS);
    if (nchunk > ncol) nchunk = ncol;
    if (nchunk > 1024) nchunk = 1024;
    if (nchunk < 1) nchunk = 1;
    const int64_t chunk_len = (ncol + nchunk - 1) / nchunk;
    return (ncol + chunk_len - 1) / chunk_len;
}

template <bool OUTER>
static int launch_lemma(int rk, int ck, dim3 grid, cudaStream_t st, const cplx *prep, int64_t ps, int no, int nv,
                        const int32_t *rsub, int64_t nrow, const int32_t *csub, int64_t ncol, int64_t chunk_len,
                        const cplx *Y, int64_t y_sstride, int ny, cplx *out) {
#define APYIB_LEMMA_CASE(R_, C_)                                                                                   \
    if (rk == R_ && ck == C_)                                                                                      \
        lemma_kernel<R_, C_, OUTER><<<grid, kLemmaThreads, 0, st>>>(prep, ps, no, nv, rsub, nrow, csub, ncol,      \
                                                                    chunk_len, Y, y_sstride, ny, out);
    APYIB_LEMMA_CASE(0, 0) APYIB_LEMMA_CASE(0, 1) APYIB_LEMMA_CASE(0, 2) APYIB_LEMMA_CASE(1, 0) APYIB_LEMMA_CASE(1, 1)
    APYIB_LEMMA_CASE(1, 2) APYIB_LEMMA_CASE(2, 0) APYIB_LEMMA_CASE(2, 1) APYIB_LEMMA_CASE(2, 2)
#undef APYIB_LEMMA_CASE
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

}  // namespace apyib

using namespace apyib;

extern "C" int64_t apyib_lemma_prep_len(int ns, int no) {
    const int64_t nv = ns - no;
    return 1 + (int64_t)no * no + 2 * nv * no + nv * nv;
}

extern "C" int apyib_lemma_prepare(const void *d_S, int nS, int ns, int no, void *d_prep, void *stream) {
    APYIB_REQUIRE(d_S && d_prep, "null pointer");
    APYIB_REQUIRE(nS >= 1 && no >= 1 && ns >= no && no <= 64, "1 <= no <= 64");
    const size_t smem = (size_t)no * 2 * no * sizeof(cplx);
    static bool attr_set = false;
    if (!attr_set) {
        APYIB_CUDA_CHECK(cudaFuncSetAttribute(lemma_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 128 * 16));
        attr_set = true;
    }
    lemma_prepare_kernel<<<nS, 256, smem, (cudaStream_t)stream>>>((const cplx *)d_S, ns, no, (cplx *)d_prep,
                                                                 apyib_lemma_prep_len(ns, no));
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_lemma_outer(const void *d_prep, int nS, int ns, int no, int rk, const int32_t *d_rsub, int64_t nrow,
                                 int ck, const int32_t *d_csub, int64_t ncol, void *d_out, void *stream) {
    APYIB_REQUIRE(d_prep && d_out, "null pointer");
    APYIB_REQUIRE(rk >= 0 && rk <= 2 && ck >= 0 && ck <= 2 && nS >= 1, "0 <= rk, ck <= 2");
    APYIB_REQUIRE((rk == 0 || d_rsub) && (ck == 0 || d_csub), "substitution lists");
    if (nrow == 0 || ncol == 0) return APYIB_OK;
    const int64_t rb = (nrow + kLemmaThreads - 1) / kLemmaThreads;
    const int64_t nchunk = lemma_nchunk(nrow, ncol, nS);
    const int64_t chunk_len = (ncol + nchunk - 1) / nchunk;
    dim3 grid((unsigned)rb, (unsigned)nchunk, (unsigned)nS);
    return launch_lemma<true>(rk, ck, grid, (cudaStream_t)stream, (const cplx *)d_prep, apyib_lemma_prep_len(ns, no), no,
                              ns - no, d_rsub, nrow, d_csub, ncol, chunk_len, nullptr, 0, 0, (cplx *)d_out);
}

extern "C" int64_t apyib_lemma_matvec_work_len(int64_t nrow, int64_t ncol, int ny, int nS) {
    return lemma_nchunk(nrow, ncol, nS) * ny * nrow * nS;
}

// Z[(s*ny + q)*nrow + r] = sum_c det_s(r,c) * Y[s*y_sstride + q*ncol + c]   (y_sstride = 0: shared Y)
extern "C" int apyib_lemma_matvec(const void *d_prep, int nS, int ns, int no, int rk, const int32_t *d_rsub,
                                  int64_t nrow, int ck, const int32_t *d_csub, int64_t ncol, const void *d_Y,
                                  int64_t y_sstride, int ny, void *d_Z, void *d_work, void *stream) {
    APYIB_REQUIRE(d_prep && d_Y && d_Z && d_work, "null pointer");
    APYIB_REQUIRE(rk >= 0 && rk <= 2 && ck >= 0 && ck <= 2 && nS >= 1, "0 <= rk, ck <= 2");
    APYIB_REQUIRE((rk == 0 || d_rsub) && (ck == 0 || d_csub), "substitution lists");
    APYIB_REQUIRE(ny >= 1 && ny <= 4, "1 <= ny <= 4");
    if (nrow == 0) return APYIB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (ncol == 0) {
        APYIB_CUDA_CHECK(cudaMemsetAsync(d_Z, 0, sizeof(cplx) * ny * nrow * nS, st));
        return APYIB_OK;
    }
    const int64_t rb = (nrow + kLemmaThreads - 1) / kLemmaThreads;
    const int64_t nchunk = lemma_nchunk(nrow, ncol, nS);
    const int64_t chunk_len = (ncol + nchunk - 1) / nchunk;
    dim3 grid((unsigned)rb, (unsigned)nchunk, (unsigned)nS);
    int rc = launch_lemma<false>(rk, ck, grid, st, (const cplx *)d_prep, apyib_lemma_prep_len(ns, no), no, ns - no,
                                 d_rsub, nrow, d_csub, ncol, chunk_len, (const cplx *)d_Y, y_sstride, ny, (cplx *)d_work);
    if (rc != APYIB_OK) return rc;
    const int64_t len = (int64_t)ny * nrow, total = len * nS;
    int64_t b = (total + 255) / 256;
    if (b > 148 * 8) b = 148 * 8;
    lemma_chunk_reduce_kernel<<<(unsigned)b, 256, 0, st>>>((const cplx *)d_work, (int)nchunk, len, total, (cplx *)d_Z);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}
