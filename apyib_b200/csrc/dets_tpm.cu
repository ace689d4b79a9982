// Thread-per-matrix LU determinants of substituted occupied-overlap matrices (aats.py:120-130,
// 558-642), n <= 12.  Second generation of the LU kernel of dets.cu (which stays for n > 12).
//
// The sub-warp kernel of dets.cu keeps one matrix row per lane: half of the lanes idle on average
// (rows above the pivot), 32 % N lanes idle always, and every pivot-row element is a 4-SHFL
// broadcast -- 1250 issue slots per 3 matrices at n = 9, 9 % of the FP64 pipe.  Here ONE THREAD
// owns ONE MATRIX, so a warp instruction does useful work in all 32 lanes and nothing is shuffled:
//
//   * left-looking LU by column panels of B columns held in registers (B*N complex numbers,
//     fully unrolled -> static register indices); the already-factorised L columns live in shared
//     memory, thread-interleaved (slot*T + tid -> conflict-free LDS.128), and each L element
//     loaded feeds B complex MACs, which keeps the kernel off the shared-memory roofline
//     (B = 1 would need one 512-byte LDS per 4 DFMA: 2x over the 128 B/clk/SM crossbar);
//   * partial pivoting with LAPACK's izamax metric |re|+|im| (first maximum): one integer max
//     per element on (high word of the magnitude | reversed row index);
//   * a row interchange is register selects on the panel columns plus a swap of the thread's
//     L rows and of its row-offset list in shared memory; it is skipped warp-uniformly when no
//     lane needs it (finite-difference overlaps are I + O(h): only the substituted rows/columns
//     ever pivot);
//   * the matrix is formed on the fly, a[i][j] = S[rows[i]][cols[j]], from a shared-memory copy
//     of S (or through L1 when S is too large for it): no substituted matrix, and in the fused
//     mode no determinant table (the reference's 8-index tensor, aats.py:575), exists in memory.
//
// Mapping: thread = one row list r (kept for the whole kernel), the block walks a chunk of column
// lists c (uniform per warp -> broadcast loads); fused mode accumulates
// z[q] += det(r,c) * Y[q,c] in registers and writes ny numbers per (chunk, r).
#include "common.cuh"

namespace apyib {

constexpr int kTpmThreads = 128;

template <int N, int B> struct tpm_cfg {
    static constexpr int NP = (N + B - 1) / B;                  // panels
    static constexpr int NL = (NP - 1) * B;                     // columns whose L is read again later
    __host__ __device__ static constexpr int loff(int k) { return k * (N - 1) - k * (k - 1) / 2; }
    static constexpr int LCOUNT = NL * (N - 1) - NL * (NL - 1) / 2;   // sum_{k<NL} (N-1-k)
    static constexpr int per_thread_bytes = LCOUNT * 16 + N * 4;
};

// panel width by size: B*N complex = 4*B*N registers for the panel
__host__ __device__ constexpr int tpm_panel(int n) { return n <= 6 ? n : (n <= 8 ? 4 : (n <= 12 ? 3 : 2)); }

template <int N, int B, bool SSM>
__global__ void __launch_bounds__(kTpmThreads, (N <= 10) ? 3 : 2)
det_tpm_kernel(const cplx *__restrict__ S, int ns, const int32_t *__restrict__ rows, int64_t nrow,
               const int32_t *__restrict__ cols, int64_t ncol, int64_t chunk_len, const cplx *__restrict__ Y, int ny,
               cplx *__restrict__ out, int outer) {
    using cfg = tpm_cfg<N, B>;
    constexpr int T = kTpmThreads;
    extern __shared__ __align__(16) unsigned char tpm_smem[];
    cplx *Ssm = reinterpret_cast<cplx *>(tpm_smem);
    const int ssz = SSM ? ns * ns : 0;
    cplx *Lsm = Ssm + ssz + threadIdx.x;                         // thread-interleaved: [slot*T]
    int *rpsm = reinterpret_cast<int *>(Ssm + ssz + (size_t)cfg::LCOUNT * T) + threadIdx.x;
    if (SSM) {
        for (int e = threadIdx.x; e < ssz; e += T) Ssm[e] = ldg(&S[e]);
        __syncthreads();
    }
    const int64_t r = (int64_t)blockIdx.x * T + threadIdx.x;
    const bool rvalid = r < nrow;
    const int64_t rr = rvalid ? r : nrow - 1;
    int rowoff[N];
#pragma unroll
    for (int i = 0; i < N; ++i) rowoff[i] = __ldg(&rows[rr * N + i]) * ns;

    const int64_t c0 = (int64_t)blockIdx.y * chunk_len;
    int64_t c1 = c0 + chunk_len;
    if (c1 > ncol) c1 = ncol;

    constexpr int NYMAX = 4;
    cplx z[NYMAX];
#pragma unroll
    for (int q = 0; q < NYMAX; ++q) z[q] = make_cplx(0.0, 0.0);

    for (int64_t c = c0; c < c1; ++c) {
        const int32_t *cl = cols + c * N;
#pragma unroll
        for (int i = 0; i < N; ++i) rpsm[i * T] = rowoff[i];
        double detx = 1.0, dety = 0.0;
        bool neg = false;
#pragma unroll
        for (int jb = 0; jb < N; jb += B) {
            constexpr int dummy = 0;
            (void)dummy;
            const int bw = (N - jb < B) ? (N - jb) : B;
            cplx a[B][N];
            // ---- form the panel columns from S (row order = current permutation) ----
            {
                int ro[N];
#pragma unroll
                for (int i = 0; i < N; ++i) ro[i] = (jb == 0) ? rowoff[i] : rpsm[i * T];
#pragma unroll
                for (int jj = 0; jj < B; ++jj) {
                    if (jj < bw) {
                        const int col = __ldg(&cl[jb + jj]);
#pragma unroll
                        for (int i = 0; i < N; ++i) a[jj][i] = SSM ? Ssm[ro[i] + col] : ldg(&S[ro[i] + col]);
                    }
                }
            }
            // ---- left-looking update with the L columns of earlier panels ----
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
            // ---- factorise the panel (right-looking inside it) ----
#pragma unroll
            for (int jj = 0; jj < B; ++jj) {
                if (jj < bw) {
                    const int j = jb + jj;
                    // pivot: first maximum of |re|+|im| over rows j..N-1
                    unsigned best = 0u;
#pragma unroll
                    for (int i = j; i < N; ++i) {
                        const double mag = fabs(a[jj][i].x) + fabs(a[jj][i].y);
                        const unsigned key = (((unsigned)__double2hiint(mag)) & 0xffffffe0u) | (unsigned)(31 - i);
                        best = (key > best) ? key : best;
                    }
                    const int p = 31 - (int)(best & 31u);
                    const bool sw = (p != j);
                    if (j + 1 < N && __any_sync(0xffffffffu, sw)) {
#pragma unroll
                        for (int j2 = jj; j2 < B; ++j2) {
                            if (j2 < bw) {
                                const cplx t = a[j2][j];
#pragma unroll
                                for (int i = j + 1; i < N; ++i) {
                                    const bool m = (i == p);
                                    const cplx ci = a[j2][i];
                                    a[j2][j].x = m ? ci.x : a[j2][j].x;
                                    a[j2][j].y = m ? ci.y : a[j2][j].y;
                                    a[j2][i].x = m ? t.x : ci.x;
                                    a[j2][i].y = m ? t.y : ci.y;
                                }
                            }
                        }
                        if (sw) {
                            if (cfg::NL > 0 && j < N - 1) {
                                const int t0 = rpsm[j * T], t1 = rpsm[p * T];
                                rpsm[j * T] = t1;
                                rpsm[p * T] = t0;
                            }
#pragma unroll
                            for (int k = 0; k < j; ++k) {
                                if (k < cfg::NL) {
                                    cplx *pj = &Lsm[(cfg::loff(k) + j - k - 1) * T];
                                    cplx *pp = &Lsm[(cfg::loff(k) + p - k - 1) * T];
                                    const cplx u = *pj, w = *pp;
                                    *pj = w;
                                    *pp = u;
                                }
                            }
                            neg = !neg;
                        }
                    }
                    const double pvx = a[jj][j].x, pvy = a[jj][j].y;
                    const double ndx = detx * pvx - dety * pvy;
                    dety = detx * pvy + dety * pvx;
                    detx = ndx;
                    if (j + 1 < N) {
                        const double d2 = fma(pvx, pvx, pvy * pvy);
                        const double rinv = (d2 > 0.0) ? __drcp_rn(d2) : 0.0;   // singular column -> det = 0
                        const double ix = pvx * rinv, iy = -pvy * rinv;
#pragma unroll
                        for (int i = j + 1; i < N; ++i) {
                            const double lx = fma(a[jj][i].x, ix, -a[jj][i].y * iy);
                            const double ly = fma(a[jj][i].x, iy, a[jj][i].y * ix);
                            a[jj][i].x = lx;
                            a[jj][i].y = ly;
                            if (j < cfg::NL) Lsm[(cfg::loff(j) + i - j - 1) * T] = make_cplx(lx, ly);
#pragma unroll
                            for (int j2 = jj + 1; j2 < B; ++j2) {
                                if (j2 < bw) {
                                    a[j2][i].x = fma(ly, a[j2][j].y, fma(-lx, a[j2][j].x, a[j2][i].x));
                                    a[j2][i].y = fma(-ly, a[j2][j].x, fma(-lx, a[j2][j].y, a[j2][i].y));
                                }
                            }
                        }
                    }
                }
            }
        }
        const cplx d = make_cplx(neg ? -detx : detx, neg ? -dety : dety);
        if (outer) {
            if (rvalid) out[r * ncol + c] = d;
        } else {
#pragma unroll
            for (int q = 0; q < NYMAX; ++q)
                if (q < ny) z[q] = z[q] + d * ldg(&Y[(int64_t)q * ncol + c]);
        }
    }
    if (!outer && rvalid) {
#pragma unroll
        for (int q = 0; q < NYMAX; ++q)
            if (q < ny) out[((int64_t)blockIdx.y * ny + q) * nrow + r] = z[q];
    }
}

template <int N>
static int launch_tpm_n(dim3 grid, cudaStream_t st, const cplx *S, int ns, const int32_t *rows, int64_t nrow,
                        const int32_t *cols, int64_t ncol, int64_t chunk_len, const cplx *Y, int ny, cplx *out,
                        int outer) {
    constexpr int B = tpm_panel(N);
    using cfg = tpm_cfg<N, B>;
    const size_t base = (size_t)cfg::per_thread_bytes * kTpmThreads;
    const size_t s_bytes = (size_t)ns * ns * sizeof(cplx);
    // S in shared memory only while three blocks per SM still fit (227 KB per SM)
    const bool ssm = (base + s_bytes) * 3 <= 220 * 1024;
    const size_t smem = base + (ssm ? s_bytes : 0);
    if (ssm) {
        static bool attr_done = false;
        if (!attr_done) {
            APYIB_CUDA_CHECK(cudaFuncSetAttribute(det_tpm_kernel<N, B, true>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr_done = true;
        }
        det_tpm_kernel<N, B, true><<<grid, kTpmThreads, smem, st>>>(S, ns, rows, nrow, cols, ncol, chunk_len, Y, ny,
                                                                    out, outer);
    } else {
        static bool attr_done = false;
        if (!attr_done) {
            APYIB_CUDA_CHECK(cudaFuncSetAttribute(det_tpm_kernel<N, B, false>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr_done = true;
        }
        det_tpm_kernel<N, B, false><<<grid, kTpmThreads, smem, st>>>(S, ns, rows, nrow, cols, ncol, chunk_len, Y, ny,
                                                                     out, outer);
    }
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

// resident blocks per SM (shared-memory bound), used by the host to size the grid in whole waves
int tpm_blocks_per_sm(int n, int ns) {
    int per_thread = 0;
    switch (n) {
#define APYIB_TPM_PT(NN) case NN: per_thread = tpm_cfg<NN, tpm_panel(NN)>::per_thread_bytes; break;
        APYIB_TPM_PT(2) APYIB_TPM_PT(3) APYIB_TPM_PT(4) APYIB_TPM_PT(5) APYIB_TPM_PT(6) APYIB_TPM_PT(7)
        APYIB_TPM_PT(8) APYIB_TPM_PT(9) APYIB_TPM_PT(10) APYIB_TPM_PT(11) APYIB_TPM_PT(12)
#undef APYIB_TPM_PT
        default: return 0;
    }
    const size_t base = (size_t)per_thread * kTpmThreads;
    const size_t s_bytes = (size_t)ns * ns * sizeof(cplx);
    const size_t smem = ((base + s_bytes) * 3 <= 220 * 1024) ? base + s_bytes : base;
    int b = (int)((227 * 1024) / (smem + 1024));
    if (b > (n <= 10 ? 3 : 2)) b = (n <= 10 ? 3 : 2);     // register bound (launch bounds of the kernel)
    if (b < 1) b = 1;
    return b;
}

int launch_det_tpm(int n, dim3 grid, cudaStream_t st, const cplx *S, int ns, const int32_t *rows, int64_t nrow,
                   const int32_t *cols, int64_t ncol, int64_t chunk_len, const cplx *Y, int ny, cplx *out, int outer) {
    switch (n) {
#define APYIB_TPM_CASE(NN) \
    case NN: return launch_tpm_n<NN>(grid, st, S, ns, rows, nrow, cols, ncol, chunk_len, Y, ny, out, outer);
        APYIB_TPM_CASE(2) APYIB_TPM_CASE(3) APYIB_TPM_CASE(4) APYIB_TPM_CASE(5) APYIB_TPM_CASE(6) APYIB_TPM_CASE(7)
        APYIB_TPM_CASE(8) APYIB_TPM_CASE(9) APYIB_TPM_CASE(10) APYIB_TPM_CASE(11) APYIB_TPM_CASE(12)
#undef APYIB_TPM_CASE
    }
    set_error("thread-per-matrix LU: n = %d not instantiated", n);
    return APYIB_ERR_ARG;
}

}  // namespace apyib
