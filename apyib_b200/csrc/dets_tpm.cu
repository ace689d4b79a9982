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
// Mapping: a warp takes one task = (32 consecutive row lists r) x (one chunk of column lists c);
// lane = row list, the column list is warp-uniform (broadcast loads).  One block per SM, sized by
// registers and shared memory; tasks are dealt out per warp.  Fused mode accumulates
// z[q] += det(r,c) * Y[q,c] in registers and writes ny numbers per (chunk, r).
#include <type_traits>
#include "common.cuh"

namespace apyib {

// panel width: the whole matrix in registers up to n = 6, three columns above
__host__ __device__ constexpr int tpm_panel(int n) { return n <= 6 ? n : 3; }

template <int N, int B> struct tpm_cfg {
    static constexpr int NP = (N + B - 1) / B;                  // panels
    static constexpr int NL = (NP - 1) * B;                     // columns whose L is read again later
    __host__ __device__ static constexpr int loff(int k) { return k * (N - 1) - k * (k - 1) / 2; }
    static constexpr int LCOUNT = NL * (N - 1) - NL * (NL - 1) / 2;   // sum_{k<NL} (N-1-k)
    static constexpr int per_thread_bytes = LCOUNT * 16 + N * 4;
};

// ONE block per SM; its size is what registers (launch bounds -> <= 65536/T per thread) and the
// per-thread shared-memory footprint allow.  Work is handed out per warp, so the block size does
// not quantise the problem.
__host__ __device__ constexpr int tpm_threads(int n) {
    return n <= 4 ? 512 : n <= 6 ? 256 : n <= 9 ? 384 : n == 10 ? 288 : n == 11 ? 224 : 192;
}
constexpr int kTpmSmemMax = 227 * 1024;

template <int N, int B, bool SSM>
__global__ void __launch_bounds__(tpm_threads(N), 1)
det_tpm_kernel(const cplx *__restrict__ S, int ns, const int32_t *__restrict__ rows, int64_t nrow,
               const int32_t *__restrict__ cols, int64_t ncol, int64_t chunk_len, int64_t nchunk,
               const double *__restrict__ csign, const int32_t *__restrict__ cindex,
               const cplx *__restrict__ Y, int ny, cplx *__restrict__ out, int outer, int64_t y_stride,
               int64_t out_stride) {
    using cfg = tpm_cfg<N, B>;
    // blockIdx.y = overlap of a stack (same index lists, own S / Y / output slab)
    S += (size_t)blockIdx.y * ns * ns;
    if (Y != nullptr) Y += (size_t)blockIdx.y * y_stride;
    out += (size_t)blockIdx.y * out_stride;
    constexpr int T = tpm_threads(N);
    static_assert((size_t)cfg::per_thread_bytes * T <= kTpmSmemMax - 1024, "shared-memory footprint");
    extern __shared__ __align__(16) unsigned char tpm_smem[];
    cplx *Ssm = reinterpret_cast<cplx *>(tpm_smem);
    const int ssz = SSM ? ns * ns : 0;
    cplx *Lsm = Ssm + ssz + threadIdx.x;                         // thread-interleaved: [slot*T]
    int *rpsm = reinterpret_cast<int *>(Ssm + ssz + (size_t)cfg::LCOUNT * T) + threadIdx.x;
    if (SSM) {
        for (int e = threadIdx.x; e < ssz; e += T) Ssm[e] = ldg(&S[e]);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int64_t nrg = (nrow + 31) >> 5;                        // groups of 32 row lists
    const int64_t ntask = nrg * nchunk;
    constexpr int NYMAX = 4;

    for (int64_t task = (int64_t)blockIdx.x * (T / 32) + (threadIdx.x >> 5); task < ntask;
         task += (int64_t)gridDim.x * (T / 32)) {
        const int64_t ch = task / nrg, rg = task - ch * nrg;
        const int64_t r = rg * 32 + lane;
        const bool rvalid = r < nrow;
        const int64_t rr = rvalid ? r : nrow - 1;
        int rowoff[N];
#pragma unroll
        for (int i = 0; i < N; ++i) rowoff[i] = __ldg(&rows[rr * N + i]) * ns;
        const int64_t c0 = ch * chunk_len;
        int64_t c1 = c0 + chunk_len;
        if (c1 > ncol) c1 = ncol;
        cplx z[NYMAX];
#pragma unroll
        for (int q = 0; q < NYMAX; ++q) z[q] = make_cplx(0.0, 0.0);

        double pdx = 1.0, pdy = 0.0;          // determinant prefix after the leading panels (factorisation reuse)
        bool pneg = false;
        for (int64_t c = c0; c < c1; ++c) {
            const int32_t *cl = cols + c * N;
            // Left-looking LU touches column j only after columns < j: if this column list starts with the
            // same NL columns as the previous one (same thread -> same rows), the leading panels, their L
            // columns in shared memory, the row permutation and the partial determinant are all still valid
            // and only the last panel is factorised.  Callers order the column lists accordingly
            // (substituted columns last, lists sorted; apyib_det_sort_lists).
            bool reuse = false;
            if (cfg::NL > 0 && c > c0) {
                reuse = true;
#pragma unroll
                for (int j = 0; j < cfg::NL; ++j) reuse = reuse && (__ldg(&cl[j]) == __ldg(&cl[j - N]));
            }
            if (cfg::NP > 1 && !reuse) {
#pragma unroll
                for (int i = 0; i < N; ++i) rpsm[i * T] = rowoff[i];
            }
            double detx = reuse ? pdx : 1.0, dety = reuse ? pdy : 0.0;
            bool neg = reuse ? pneg : false;
#pragma unroll
            for (int jb = 0; jb < N; jb += B) {
                const bool last_panel = (jb + B >= N);
                if (!last_panel && reuse) continue;
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
                        // One elimination step.  SWAP = true first gathers the pivot row (dynamic p) and
                        // drops the old row j into slot p with selects; the selected values are only
                        // temporaries of the FMAs that follow, so the two code versions merge without
                        // register shuffling (an in-place swap under a branch costs ~2 MOVs per select).
                        auto step = [&](auto swap_tag) {
                            constexpr bool SWAP = decltype(swap_tag)::value;
                            cplx u[B];                       // pivot row in the panel columns
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
                            const double rinv = (d2 > 0.0) ? __drcp_rn(d2) : 0.0;   // singular column -> det = 0
                            const double ix = pvx * rinv, iy = -pvy * rinv;
#pragma unroll
                            for (int i = j + 1; i < N; ++i) {
                                const bool m = SWAP && (i == p);
                                const double xr = m ? a[jj][j].x : a[jj][i].x, xi = m ? a[jj][j].y : a[jj][i].y;
                                const double lx = fma(xr, ix, -xi * iy);
                                const double ly = fma(xr, iy, xi * ix);
                                if (j < cfg::NL) Lsm[(cfg::loff(j) + i - j - 1) * T] = make_cplx(lx, ly);
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
                        if (j + 1 < N) {
                            if (__any_sync(0xffffffffu, sw)) {
                                if (sw && !last_panel) {       // (nobody reads L or the row list after the last panel)
                                    {
                                        const int t0 = rpsm[j * T], t1 = rpsm[p * T];
                                        rpsm[j * T] = t1;
                                        rpsm[p * T] = t0;
                                    }
                                    // L rows j <-> p: all loads first, then all stores (latency overlapped)
                                    constexpr int KS = (cfg::NL < N ? cfg::NL : N);
                                    cplx lj[KS > 0 ? KS : 1], lp[KS > 0 ? KS : 1];
#pragma unroll
                                    for (int k = 0; k < j; ++k) {
                                        if (k < cfg::NL) {
                                            lj[k] = Lsm[(cfg::loff(k) + j - k - 1) * T];
                                            lp[k] = Lsm[(cfg::loff(k) + p - k - 1) * T];
                                        }
                                    }
#pragma unroll
                                    for (int k = 0; k < j; ++k) {
                                        if (k < cfg::NL) {
                                            Lsm[(cfg::loff(k) + j - k - 1) * T] = lp[k];
                                            Lsm[(cfg::loff(k) + p - k - 1) * T] = lj[k];
                                        }
                                    }
                                }
                                neg = neg != sw;
                                step(std::true_type{});
                            } else {
                                step(std::false_type{});
                            }
                        } else {
                            const double pvx = a[jj][j].x, pvy = a[jj][j].y;
                            const double ndx = detx * pvx - dety * pvy;
                            dety = detx * pvy + dety * pvx;
                            detx = ndx;
                        }
                    }
                }
                if (cfg::NL > 0 && jb + B == cfg::NL) { pdx = detx; pdy = dety; pneg = neg; }
            }
            const double sg = (csign ? __ldg(&csign[c]) : 1.0) * (neg ? -1.0 : 1.0);
            const cplx d = make_cplx(sg * detx, sg * dety);
            const int64_t cc = cindex ? (int64_t)__ldg(&cindex[c]) : c;
            if (outer) {
                if (rvalid) out[r * ncol + cc] = d;
            } else {
#pragma unroll
                for (int q = 0; q < NYMAX; ++q)
                    if (q < ny) z[q] = z[q] + d * ldg(&Y[(int64_t)q * ncol + cc]);
            }
        }
        if (!outer && rvalid) {
#pragma unroll
            for (int q = 0; q < NYMAX; ++q)
                if (q < ny) out[(ch * ny + q) * nrow + r] = z[q];
        }
    }
}

template <int N>
static int launch_tpm_n(cudaStream_t st, const cplx *S, int ns, const int32_t *rows, int64_t nrow,
                        const int32_t *cols, int64_t ncol, int64_t chunk_len, int64_t nchunk, const double *csign,
                        const int32_t *cindex, const cplx *Y, int ny, cplx *out, int outer, int nS, int64_t y_stride,
                        int64_t out_stride) {
    constexpr int B = tpm_panel(N);
    constexpr int T = tpm_threads(N);
    using cfg = tpm_cfg<N, B>;
    const size_t base = (size_t)cfg::per_thread_bytes * T;
    const size_t s_bytes = (size_t)ns * ns * sizeof(cplx);
    const bool ssm = base + s_bytes <= (size_t)kTpmSmemMax - 1024;   // S in shared memory when it fits
    const size_t smem = base + (ssm ? s_bytes : 0);
    const int64_t ntask = ((nrow + 31) / 32) * nchunk;
    int64_t blocks = (ntask + T / 32 - 1) / (T / 32);
    if (blocks > 148) blocks = 148;
    auto kern = ssm ? det_tpm_kernel<N, B, true> : det_tpm_kernel<N, B, false>;
    static bool attr_done[2] = {false, false};
    if (!attr_done[ssm]) {
        APYIB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTpmSmemMax));
        attr_done[ssm] = true;
    }
    kern<<<dim3((unsigned)blocks, (unsigned)nS), T, smem, st>>>(S, ns, rows, nrow, cols, ncol, chunk_len, nchunk, csign,
                                                                cindex, Y, ny, out, outer, y_stride, out_stride);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

// warps resident on the device for size n (one block per SM): the host sizes the column chunks so
// that (row groups) x (chunks) fills them in whole waves
int tpm_total_warps(int n) { return 148 * (tpm_threads(n) / 32); }

int launch_det_tpm(int n, cudaStream_t st, const cplx *S, int ns, const int32_t *rows, int64_t nrow,
                   const int32_t *cols, int64_t ncol, int64_t chunk_len, int64_t nchunk, const double *csign,
                   const int32_t *cindex, const cplx *Y, int ny, cplx *out, int outer, int nS, int64_t y_stride,
                   int64_t out_stride) {
    switch (n) {
#define APYIB_TPM_CASE(NN) \
    case NN: return launch_tpm_n<NN>(st, S, ns, rows, nrow, cols, ncol, chunk_len, nchunk, csign, cindex, Y, ny, out, outer, nS, y_stride, out_stride);
        APYIB_TPM_CASE(2) APYIB_TPM_CASE(3) APYIB_TPM_CASE(4) APYIB_TPM_CASE(5) APYIB_TPM_CASE(6) APYIB_TPM_CASE(7)
        APYIB_TPM_CASE(8) APYIB_TPM_CASE(9) APYIB_TPM_CASE(10) APYIB_TPM_CASE(11) APYIB_TPM_CASE(12)
#undef APYIB_TPM_CASE
    }
    set_error("thread-per-matrix LU: n = %d not instantiated", n);
    return APYIB_ERR_ARG;
}

}  // namespace apyib
