// Prefix-shared LU for the doubly (K = 2) or singly (K = 1) column-substituted determinants of the
// finite-difference AAT tables (aats.py:581-618), n <= 12.  Third generation of the LU kernel.
//
// The column lists of one table come in GROUPS: all lists that substitute the same occupied columns
// share their N-K unsubstituted columns (the PREFIX, kept in front by apyib_det_sort_lists) and differ
// only in the K virtual columns that follow -- for doubles every pair (c < d) of the nc virtual columns,
// C(13,2) = 78 lists per group at H2O2/6-31G.  dets_tpm.cu already re-uses the leading panels between
// consecutive lists; here the sharing is taken to its end.  With partial pivoting confined to the prefix
// (the trailing K x K block needs none: its determinant is written out),
//
//      P A = [ L11  0 ] [ U11  U12 ]        det A = sign(P) prod(diag U11) det(S22),
//            [ L21  I ] [  0   S22 ]        S22 = rows N-K.. of  L^-1 P [s_c s_d],
//
// and rows N-K.. of L^-1 are [X | I] with X L11 = -L21, so every candidate column v contributes ONE
// K-vector w_v = X (P s_v)[:N-K] + (P s_v)[N-K:] per (row list, group) -- N*K complex MACs -- and a
// determinant of the group is a K x K determinant of two of those vectors.  Per (row list, group) at
// n = 9, nc = 13: prefix LU 133 + X 49 + candidates 234 + pairs 78*3 complex MACs for 78 determinants,
// 9.4 per determinant instead of ~110 with last-panel reuse (243 for a full LU).  It is still a
// partial-pivoting LU of the actual substituted matrices (columns permuted), valid for any overlap.
//
// Mapping as in dets_tpm.cu: one thread = one row list, a warp takes 32 consecutive row lists x one chunk
// of groups (the group is warp-uniform); L of the prefix lives in shared memory, thread-interleaved, and
// is overlaid by the w vectors once X is in registers.  Fused mode only: z[q] += det * Y[q, c].
#pragma once
#include <type_traits>
#include "common.cuh"

namespace apyib {

__host__ __device__ constexpr int pfx_threads(int n) {
    return n <= 4 ? 512 : n <= 6 ? 256 : n <= 9 ? 352 : n == 10 ? 288 : n == 11 ? 224 : 192;
}
constexpr int kPfxSmemMax = 227 * 1024;

template <int N, int K> struct pfx_cfg {
    static constexpr int NPRE = N - K;                                   // prefix columns of a group
    static constexpr int B = NPRE >= 3 ? 3 : (NPRE >= 1 ? NPRE : 1);     // panel width of the prefix LU
    __host__ __device__ static constexpr int loff(int k) { return k * (N - 1) - k * (k - 1) / 2; }
    static constexpr int LCOUNT = NPRE * (N - 1) - NPRE * (NPRE - 1) / 2;   // sum_{k<NPRE} (N-1-k)
};

// NYT = 0: up to 4 amplitude vectors, count given at run time (predicated); NYT = 1: exactly one vector, no
// predicated slots in the pair loop (the doubles x doubles table of every (beta x pp/pn/np/nn) stack has one).
template <int N, int K, bool SSM, int NYT = 0>
__global__ void __launch_bounds__(pfx_threads(N), 1)
det_pairs_kernel(const cplx *__restrict__ S, int ns, const int32_t *__restrict__ rows, int64_t nrow,
                 const int32_t *__restrict__ cols, int64_t ngroup, int64_t npair, const int32_t *__restrict__ cand,
                 int nc, int slots, int64_t gchunk, int64_t nchunk, const double *__restrict__ csign,
                 const int32_t *__restrict__ cindex, const cplx *__restrict__ Y, int ny, int64_t ncol,
                 cplx *__restrict__ out, int64_t y_stride, int64_t out_stride) {
    using cfg = pfx_cfg<N, K>;
    // blockIdx.y = overlap of a stack (same index lists, own S / Y / output slab)
    S += (size_t)blockIdx.y * ns * ns;
    Y += (size_t)blockIdx.y * y_stride;
    out += (size_t)blockIdx.y * out_stride;
    constexpr int NPRE = cfg::NPRE, B = cfg::B;
    constexpr int NPX = NPRE > 0 ? NPRE : 1;
    const int T = blockDim.x;
    extern __shared__ __align__(16) unsigned char pfx_smem[];
    cplx *Ssm = reinterpret_cast<cplx *>(pfx_smem);
    const int ssz = SSM ? ns * ns : 0;
    cplx *Lsm = Ssm + ssz + threadIdx.x;                                 // thread-interleaved: [slot * T]
    int *rpsm = reinterpret_cast<int *>(Ssm + ssz + (size_t)slots * T) + threadIdx.x;
    if (SSM) {
        for (int e = threadIdx.x; e < ssz; e += T) Ssm[e] = ldg(&S[e]);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int64_t nrg = (nrow + 31) >> 5;
    const int64_t ntask = nrg * nchunk;
    constexpr int NYMAX = NYT ? NYT : 4;
    auto ld = [&](int off) -> cplx { return SSM ? Ssm[off] : ldg(&S[off]); };

    for (int64_t task = (int64_t)blockIdx.x * (T / 32) + (threadIdx.x >> 5); task < ntask;
         task += (int64_t)gridDim.x * (T / 32)) {
        const int64_t ch = task / nrg, rg = task - ch * nrg;
        const int64_t r = rg * 32 + lane;
        const bool rvalid = r < nrow;
        const int64_t rr = rvalid ? r : nrow - 1;
        int rowoff[N];
#pragma unroll
        for (int i = 0; i < N; ++i) rowoff[i] = __ldg(&rows[rr * N + i]) * ns;
        cplx z[NYMAX];
#pragma unroll
        for (int q = 0; q < NYMAX; ++q) z[q] = make_cplx(0.0, 0.0);
        const int64_t g0 = ch * gchunk;
        int64_t g1 = g0 + gchunk;
        if (g1 > ngroup) g1 = ngroup;

        for (int64_t g = g0; g < g1; ++g) {
            const int32_t *cl = cols + g * npair * N;                    // prefix = first NPRE entries of the group's lists
#pragma unroll
            for (int i = 0; i < N; ++i) rpsm[i * T] = rowoff[i];
            double detx = 1.0, dety = 0.0;
            bool neg = false;
            // ---------------- prefix LU: left-looking panels of B columns, partial pivoting ----------------
#pragma unroll
            for (int jb = 0; jb < NPRE; jb += B) {
                const int bw = (NPRE - jb < B) ? (NPRE - jb) : B;
                cplx a[B][N];
                {
                    int ro[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) ro[i] = (jb == 0) ? rowoff[i] : rpsm[i * T];
#pragma unroll
                    for (int jj = 0; jj < B; ++jj) {
                        if (jj < bw) {
                            const int col = __ldg(&cl[jb + jj]);
#pragma unroll
                            for (int i = 0; i < N; ++i) a[jj][i] = ld(ro[i] + col);
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < jb; ++k) {
#pragma unroll
                    for (int i = k + 1; i < N; ++i) {
                        const cplx l = Lsm[(cfg::loff(k) + i - k - 1) * T];
#pragma unroll
                        for (int jj = 0; jj < B; ++jj) {
                            if (jj < bw) {
                                a[jj][i].x = fma(l.y, a[jj][k].y, fma(-l.x, a[jj][k].x, a[jj][i].x));
                                a[jj][i].y = fma(-l.y, a[jj][k].x, fma(-l.x, a[jj][k].y, a[jj][i].y));
                            }
                        }
                    }
                }
#pragma unroll
                for (int jj = 0; jj < B; ++jj) {
                    if (jj < bw) {
                        const int j = jb + jj;                           // j <= NPRE-1 <= N-2: there is always a row below
                        unsigned best = 0u;
#pragma unroll
                        for (int i = j; i < N; ++i) {
                            const double mag = fabs(a[jj][i].x) + fabs(a[jj][i].y);
                            const unsigned key = (((unsigned)__double2hiint(mag)) & 0xffffffe0u) | (unsigned)(31 - i);
                            best = (key > best) ? key : best;
                        }
                        const int p = 31 - (int)(best & 31u);
                        const bool sw = (p != j);
                        auto step = [&](auto swap_tag) {
                            constexpr bool SWAP = decltype(swap_tag)::value;
                            cplx u[B];
#pragma unroll
                            for (int j2 = jj; j2 < B; ++j2) {
                                if (j2 < bw) {
                                    u[j2] = a[j2][j];
                                    if (SWAP) {
#pragma unroll
                                        for (int i = j + 1; i < N; ++i) {
                                            const bool m = (i == p);
                                            u[j2].x = m ? a[j2][i].x : u[j2].x;
                                            u[j2].y = m ? a[j2][i].y : u[j2].y;
                                        }
                                    }
                                }
                            }
                            const double pvx = u[jj].x, pvy = u[jj].y;
                            const double ndx = detx * pvx - dety * pvy;
                            dety = detx * pvy + dety * pvx;
                            detx = ndx;
                            const double d2 = fma(pvx, pvx, pvy * pvy);
                            const double rinv = (d2 > 0.0) ? __drcp_rn(d2) : 0.0;   // singular prefix -> det = 0
                            const double ix = pvx * rinv, iy = -pvy * rinv;
#pragma unroll
                            for (int i = j + 1; i < N; ++i) {
                                const bool m = SWAP && (i == p);
                                const double xr = m ? a[jj][j].x : a[jj][i].x, xi = m ? a[jj][j].y : a[jj][i].y;
                                const double lx = fma(xr, ix, -xi * iy);
                                const double ly = fma(xr, iy, xi * ix);
                                Lsm[(cfg::loff(j) + i - j - 1) * T] = make_cplx(lx, ly);
#pragma unroll
                                for (int j2 = jj + 1; j2 < B; ++j2) {
                                    if (j2 < bw) {
                                        const double yr = m ? a[j2][j].x : a[j2][i].x, yi = m ? a[j2][j].y : a[j2][i].y;
                                        a[j2][i].x = fma(ly, u[j2].y, fma(-lx, u[j2].x, yr));
                                        a[j2][i].y = fma(-ly, u[j2].x, fma(-lx, u[j2].y, yi));
                                    }
                                }
                            }
#pragma unroll
                            for (int j2 = jj + 1; j2 < B; ++j2)
                                if (j2 < bw) a[j2][j] = u[j2];
                        };
                        if (__any_sync(0xffffffffu, sw)) {
                            if (sw) {
                                {
                                    const int t0 = rpsm[j * T], t1 = rpsm[p * T];
                                    rpsm[j * T] = t1;
                                    rpsm[p * T] = t0;
                                }
                                cplx lj[NPX], lp[NPX];
#pragma unroll
                                for (int k = 0; k < j; ++k) {
                                    lj[k] = Lsm[(cfg::loff(k) + j - k - 1) * T];
                                    lp[k] = Lsm[(cfg::loff(k) + p - k - 1) * T];
                                }
#pragma unroll
                                for (int k = 0; k < j; ++k) {
                                    Lsm[(cfg::loff(k) + j - k - 1) * T] = lp[k];
                                    Lsm[(cfg::loff(k) + p - k - 1) * T] = lj[k];
                                }
                            }
                            neg = neg != sw;
                            step(std::true_type{});
                        } else {
                            step(std::false_type{});
                        }
                    }
                }
            }
            // ---------------- X = -L21 L11^-1 (rows N-K.. of L^-1, prefix columns) ----------------
            cplx X[K][NPX];
#pragma unroll
            for (int k = NPRE - 1; k >= 0; --k) {
                cplx acc[K];
#pragma unroll
                for (int t = 0; t < K; ++t) {
                    const cplx l = Lsm[(cfg::loff(k) + (NPRE + t) - k - 1) * T];
                    acc[t] = make_cplx(-l.x, -l.y);
                }
#pragma unroll
                for (int m = k + 1; m < NPRE; ++m) {
                    const cplx l = Lsm[(cfg::loff(k) + m - k - 1) * T];
#pragma unroll
                    for (int t = 0; t < K; ++t) {
                        acc[t].x = fma(X[t][m].y, l.y, fma(-X[t][m].x, l.x, acc[t].x));
                        acc[t].y = fma(-X[t][m].y, l.x, fma(-X[t][m].x, l.y, acc[t].y));
                    }
                }
#pragma unroll
                for (int t = 0; t < K; ++t) X[t][k] = acc[t];
            }
            // ---------------- candidate columns: w_v = X (P s_v)[:NPRE] + (P s_v)[NPRE:] ----------------
            {
                int ro[N];
#pragma unroll
                for (int i = 0; i < N; ++i) ro[i] = rpsm[i * T];
                // (L is dead from here on: the w vectors overlay it, K slots per candidate)
                for (int v = 0; v < nc; ++v) {
                    const int col = __ldg(&cand[v]);
                    cplx w[K];
#pragma unroll
                    for (int t = 0; t < K; ++t) w[t] = ld(ro[NPRE + t] + col);
#pragma unroll
                    for (int k = 0; k < NPRE; ++k) {
                        const cplx s = ld(ro[k] + col);
#pragma unroll
                        for (int t = 0; t < K; ++t) {
                            w[t].x = fma(-X[t][k].y, s.y, fma(X[t][k].x, s.x, w[t].x));
                            w[t].y = fma(X[t][k].y, s.x, fma(X[t][k].x, s.y, w[t].y));
                        }
                    }
#pragma unroll
                    for (int t = 0; t < K; ++t) Lsm[(v * K + t) * T] = w[t];
                }
            }
            // ---------------- the group's determinants, fused with the table x vector product ----------------
            const double sgp = neg ? -1.0 : 1.0;
            const cplx pd = make_cplx(sgp * detx, sgp * dety);
            const int64_t cbase = g * npair;
            if (K == 2) {
                // determinants of the group: d(x, y) = prefix * (w_x[0] w_y[1] - w_x[1] w_y[0]) for x < y, list index
                // t(x, y) = x nc - x (x + 1) / 2 + (y - x - 1) (lexicographic).  Two x per pass over y: every w_y read
                // from shared memory feeds two determinants (the loop is bound by LDS issue otherwise).
                auto det2 = [](const cplx &a0, const cplx &a1, const cplx &b0, const cplx &b1) {
                    cplx d;
                    d.x = fma(-a1.x, b0.x, fma(a1.y, b0.y, fma(a0.x, b1.x, -a0.y * b1.y)));
                    d.y = fma(-a1.x, b0.y, fma(-a1.y, b0.x, fma(a0.x, b1.y, a0.y * b1.x)));
                    return d;
                };
                auto accum = [&](const cplx &d, int64_t c) {
#pragma unroll
                    for (int q = 0; q < NYMAX; ++q)
                        if (NYT || q < ny) z[q] = z[q] + d * ldg(&Y[(int64_t)q * ncol + c]);
                };
                for (int x = 0; x + 1 < nc; x += 2) {
                    const int64_t t0 = cbase + (int64_t)x * nc - (int64_t)x * (x + 1) / 2;           // t(x, x + 1)
                    const cplx a0 = pd * Lsm[(2 * x) * T], a1 = pd * Lsm[(2 * x + 1) * T];
                    const cplx e0r = Lsm[(2 * x + 2) * T], e1r = Lsm[(2 * x + 3) * T];               // w of x + 1
                    accum(det2(a0, a1, e0r, e1r), t0);                                              // pair (x, x + 1)
                    if (x + 2 < nc) {
                        const cplx e0 = pd * e0r, e1 = pd * e1r;
                        const int64_t t1 = cbase + (int64_t)(x + 1) * nc - (int64_t)(x + 1) * (x + 2) / 2;   // t(x + 1, x + 2)
                        for (int y = x + 2; y < nc; ++y) {
                            const cplx b0 = Lsm[(2 * y) * T], b1 = Lsm[(2 * y + 1) * T];
                            accum(det2(a0, a1, b0, b1), t0 + (y - x - 1));
                            accum(det2(e0, e1, b0, b1), t1 + (y - x - 2));
                        }
                    }
                }
            } else {
                for (int x = 0; x < nc; ++x) {
                    const cplx d = pd * Lsm[x * T];
                    const int64_t c = cbase + x;
#pragma unroll
                    for (int q = 0; q < NYMAX; ++q)
                        if (NYT || q < ny) z[q] = z[q] + d * ldg(&Y[(int64_t)q * ncol + c]);
                }
            }
        }
        if (rvalid) {
#pragma unroll
            for (int q = 0; q < NYMAX; ++q)
                if (NYT || q < ny) out[(ch * ny + q) * nrow + r] = z[q];
        }
    }
}

extern int g_pairs_variant;   // 0 = one kernel for 1..4 vectors; 1 = + single-vector specialisation for K = 2 (dets_pairs.cu)

// block size the shared-memory footprint allows (0 = does not fit: caller falls back to dets_tpm.cu)
template <int N, int K> static int pfx_block(int ns, int nc, bool *ssm, int *slots, size_t *smem) {
    using cfg = pfx_cfg<N, K>;
    *slots = cfg::LCOUNT > nc * K ? cfg::LCOUNT : nc * K;
    if (*slots < 1) *slots = 1;
    const size_t per_thread = (size_t)*slots * sizeof(cplx) + N * sizeof(int);
    const size_t s_bytes = (size_t)ns * ns * sizeof(cplx);
    const size_t budget = (size_t)kPfxSmemMax - 1024;
    *ssm = s_bytes + 64 * per_thread <= budget;
    const size_t avail = budget - (*ssm ? s_bytes : 0);
    int T = (int)(avail / per_thread) / 32 * 32;
    if (T > pfx_threads(N)) T = pfx_threads(N);
    *smem = per_thread * T + (*ssm ? s_bytes : 0);
    return T >= 64 ? T : 0;
}

template <int N, int K>
static int launch_pairs_nk(cudaStream_t st, const cplx *S, int ns, const int32_t *rows, int64_t nrow,
                           const int32_t *cols, int64_t ngroup, int64_t npair, const int32_t *cand, int nc,
                           int64_t gchunk, int64_t nchunk, const double *csign, const int32_t *cindex, const cplx *Y,
                           int ny, int64_t ncol, cplx *out, int nS, int64_t y_stride, int64_t out_stride) {
    bool ssm;
    int slots;
    size_t smem;
    const int T = pfx_block<N, K>(ns, nc, &ssm, &slots, &smem);
    if (T == 0) {
        set_error("prefix-shared LU: shared-memory footprint too large (n = %d, %d candidate columns)", N, nc);
        return APYIB_ERR_UNSUPPORTED;
    }
    const int64_t ntask = ((nrow + 31) / 32) * nchunk;
    int64_t blocks = (ntask + T / 32 - 1) / (T / 32);
    if (blocks > 148) blocks = 148;
    auto kern = ssm ? det_pairs_kernel<N, K, true> : det_pairs_kernel<N, K, false>;
    int slot = ssm ? 1 : 0;
    if constexpr (K == 2) {
        if (g_pairs_variant == 1 && ny == 1) {      // single-vector specialisation (apyib_det_set_pairs_variant)
            kern = ssm ? det_pairs_kernel<N, K, true, 1> : det_pairs_kernel<N, K, false, 1>;
            slot += 2;
        }
    }
    static bool attr_done[4] = {false, false, false, false};
    if (!attr_done[slot]) {
        APYIB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPfxSmemMax));
        attr_done[slot] = true;
    }
    kern<<<dim3((unsigned)blocks, (unsigned)nS), T, smem, st>>>(S, ns, rows, nrow, cols, ngroup, npair, cand, nc, slots,
                                                                gchunk, nchunk, csign, cindex, Y, ny, ncol, out, y_stride,
                                                                out_stride);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

#define APYIB_PAIRS_ARGS_DECL                                                                                          \
    cudaStream_t st, const cplx *S, int ns, const int32_t *rows, int64_t nrow, const int32_t *cols, int64_t ngroup,    \
        int64_t npair, const int32_t *cand, int nc, int64_t gchunk, int64_t nchunk, const double *csign,                \
        const int32_t *cindex, const cplx *Y, int ny, int64_t ncol, cplx *out, int nS, int64_t y_stride,                \
        int64_t out_stride
#define APYIB_PAIRS_ARGS                                                                                               \
    st, S, ns, rows, nrow, cols, ngroup, npair, cand, nc, gchunk, nchunk, csign, cindex, Y, ny, ncol, out, nS, y_stride, \
        out_stride
#define APYIB_PFX_CASE(NN, KK) \
    case NN: return launch_pairs_nk<NN, KK>(APYIB_PAIRS_ARGS);

// the instantiations are spread over three translation units (compile time): K = 1, K = 2 with n <= 9, K = 2 above
int launch_det_pairs_k1(int n, APYIB_PAIRS_ARGS_DECL);
int launch_det_pairs_k2_small(int n, APYIB_PAIRS_ARGS_DECL);
int launch_det_pairs_k2_large(int n, APYIB_PAIRS_ARGS_DECL);

}  // namespace apyib
