// Tensor contraction on the FP64 tensor path (DMMA, mma.sync.m8n8k4.f64) for sm_100a.
//
//   C[c_m[m] + c_n[n]] = alpha * sum_k opA(A[a_m[m] + a_k[k]]) * opB(B[b_k[k] + b_n[n]]) + beta * C
//
// One launch == one `opt_einsum.contract` of the reference (ci_wfn.py:84-90, 211-218,
// 310-333, 458-482; utils.py:240-252, 274-277, 381).  The reference lets opt_einsum
// transpose-copy the operands into GEMM layout; here the regrouping of tensor
// indices into (m | k | n) lives in six offset tables, and the operand tiles are gathered
// straight into shared memory with per-element cp.async (a complex128 element is exactly one
// 16-byte LDGSTS), so no transposed copy of an amplitude or integral block is ever made.
//
// Complex arithmetic: one 16-byte LDS gives a thread the (re, im) pair of its A (or B)
// fragment element; the complex product is four real DMMAs on the same fragments
// (Cr += Ar*Br, Cr += (-Ai)*Bi, Ci += Ar*Bi, Ci += Ai*Br), so the kernel does the
// "complex GEMM as 4 real GEMMs" split of the north star entirely in registers.
//
// FP64 has no tcgen05/TMEM path on sm_100a (`tcgen05.mma.kind::f64` does not exist);
// the FP64 tensor instruction is the warp-level DMMA.8x8x4.
#include "common.cuh"

namespace apyib {

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int BYTES> __device__ __forceinline__ void cp_async(void *smem, const void *gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int src = valid ? BYTES : 0;   // src-size 0 -> zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;\n" ::"r"(s), "l"(gmem), "n"(BYTES), "r"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

struct ContractArgs {
    const void *A, *B;
    void *C;
    const int64_t *a_m, *a_k, *b_k, *b_n, *c_m, *c_n;
    int64_t M, N, K;
    int64_t a_bs, b_bs, c_bs;
    const int32_t *active;
    double alpha_re, alpha_im, beta_re, beta_im;
    int a_kfast, b_kfast, conj_a, conj_b;
    int ksplit;          // > 1: K is split over `ksplit` CTAs; raw partial tiles go to `work`
    void *work;          // [batch][ksplit][M][N] elements
};

template <bool CPLX> struct elem_t { using type = double; };
template <> struct elem_t<true> { using type = cplx; };

// Shared-memory tile layouts.  An operand whose k index is the contiguous one in global memory
// (KF = true) is staged as [row][k] (k contiguous), otherwise as [k][row]; either way consecutive
// threads of the gather write consecutive shared-memory words (no store conflicts) and read
// consecutive global elements (coalesced).  Pitches make the DMMA fragment reads
// (lane -> row = lane/4, k = lane%4) conflict-free:
//   [k][row]: pitch = 4 (mod 16) doubles / 2 (mod 8) complex
//   [row][k]: pitch = 4 (mod 16) doubles / 4 (mod 8) complex
template <bool CPLX, bool KF, int ROWS, int BK> struct TileLayout {
    static constexpr int pitch = KF ? (BK + 4) : (ROWS + (CPLX ? 2 : 4));
    static constexpr int size = KF ? ROWS * pitch : BK * pitch;
    __device__ static __forceinline__ int at(int row, int k) { return KF ? row * pitch + k : k * pitch + row; }
};

// BM x BN CTA tile, BK slab, warps arranged (BM/WM) x (BN/WN), each warp WM x WN.
template <bool CPLX, int BM, int BN, int BK, int WM, int WN, int STAGES, bool AKF, bool BKF>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
contract_kernel(const ContractArgs p) {
    using T = typename elem_t<CPLX>::type;
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    constexpr int EB = CPLX ? 16 : 8;
    using LA = TileLayout<CPLX, AKF, BM, BK>;
    using LB = TileLayout<CPLX, BKF, BN, BK>;
    constexpr int TM = WM / 8, TN = WN / 8;
    constexpr int ITA = (BM * BK) / NT, ITB = (BN * BK) / NT;   // elements per thread per slab
    constexpr int RG = NT / BK;                                 // row groups: threads sharing one k
    static_assert((BM * BK) % NT == 0 && (BN * BK) % NT == 0 && NT % BK == 0, "tile/threads mismatch");
    static_assert(BK % 4 == 0, "BK must be a multiple of the DMMA k");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *As = reinterpret_cast<T *>(smem_raw);                       // [STAGES][LA::size]
    T *Bs = As + (size_t)STAGES * LA::size;                        // [STAGES][LB::size]

    const int z = blockIdx.z / p.ksplit, ks = blockIdx.z % p.ksplit;
    if (p.active != nullptr && p.active[z] == 0) return;
    const T *A = reinterpret_cast<const T *>(p.A) + (size_t)z * p.a_bs;
    const T *B = reinterpret_cast<const T *>(p.B) + (size_t)z * p.b_bs;
    T *C = reinterpret_cast<T *>(p.C) + (size_t)z * p.c_bs;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp / (BN / WN)) * WM, wn0 = (warp % (BN / WN)) * WN;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;

    // Gather mapping: every thread owns ONE k of the slab and ITA (ITB) rows, fixed for the whole
    // kernel -> row offsets live in registers, one k-offset load per operand per slab.
    const int a_k = AKF ? tid % BK : tid / RG, a_r0 = AKF ? tid / BK : tid % RG;
    const int b_k = BKF ? tid % BK : tid / RG, b_r0 = BKF ? tid / BK : tid % RG;
    int64_t a_off[ITA], b_off[ITB];
    unsigned a_ok = 0, b_ok = 0;
#pragma unroll
    for (int i = 0; i < ITA; ++i) {
        const int64_t m = m0 + a_r0 + i * RG;
        const bool ok = m < p.M;
        a_ok |= (ok ? 1u : 0u) << i;
        a_off[i] = ok ? p.a_m[m] : 0;
    }
#pragma unroll
    for (int i = 0; i < ITB; ++i) {
        const int64_t n = n0 + b_r0 + i * RG;
        const bool ok = n < p.N;
        b_ok |= (ok ? 1u : 0u) << i;
        b_off[i] = ok ? p.b_n[n] : 0;
    }

    auto k_offsets = [&](int64_t kbase, int64_t &ak, int64_t &bk) {
        const int64_t ka = kbase + a_k, kb = kbase + b_k;
        ak = (ka < p.K) ? p.a_k[ka] : -1;       // -1: beyond K -> zero fill
        bk = (kb < p.K) ? p.b_k[kb] : -1;
    };
    auto load_slab = [&](int stage, int64_t ak, int64_t bk) {
        T *as = As + (size_t)stage * LA::size;
        T *bs = Bs + (size_t)stage * LB::size;
#pragma unroll
        for (int i = 0; i < ITA; ++i) {
            const bool ok = ((a_ok >> i) & 1u) && ak >= 0;
            cp_async<EB>(as + LA::at(a_r0 + i * RG, a_k), A + (ok ? a_off[i] + ak : 0), ok);
        }
#pragma unroll
        for (int i = 0; i < ITB; ++i) {
            const bool ok = ((b_ok >> i) & 1u) && bk >= 0;
            cp_async<EB>(bs + LB::at(b_r0 + i * RG, b_k), B + (ok ? b_off[i] + bk : 0), ok);
        }
    };

    double cr[TM][TN][2];
    double ci[CPLX ? TM : 1][CPLX ? TN : 1][2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            cr[i][j][0] = cr[i][j][1] = 0.0;
            if (CPLX) ci[i][j][0] = ci[i][j][1] = 0.0;
        }

    // slab range of this CTA (split-K: contiguous chunk of the K slabs)
    const int64_t nslab_all = (p.K + BK - 1) / BK;
    const int64_t per = (nslab_all + p.ksplit - 1) / p.ksplit;
    const int64_t slab0 = (int64_t)ks * per;
    int64_t nslab = nslab_all - slab0;
    if (nslab > per) nslab = per;
    if (nslab < 0) nslab = 0;
    int64_t ak_next, bk_next;                  // k offsets of the next slab to be issued (prefetched)
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nslab) {
            k_offsets((slab0 + s) * BK, ak_next, bk_next);
            load_slab(s, ak_next, bk_next);
        }
        cp_async_commit();
    }
    k_offsets((slab0 + STAGES - 1) * BK, ak_next, bk_next);

    const int fr = lane >> 2, fk = lane & 3;   // fragment row (or col) / k within the 8x4 atom
    const double sa = p.conj_a ? -1.0 : 1.0, sb = p.conj_b ? -1.0 : 1.0;

    for (int64_t kt = 0; kt < nslab; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {   // prefetch slab kt + STAGES - 1 into the stage freed in the previous iteration
            const int64_t nk = kt + STAGES - 1;
            if (nk < nslab) load_slab((int)(nk % STAGES), ak_next, bk_next);
            cp_async_commit();
            k_offsets((slab0 + nk + 1) * BK, ak_next, bk_next);
        }
        const T *as = As + (size_t)(kt % STAGES) * LA::size;
        const T *bs = Bs + (size_t)(kt % STAGES) * LB::size;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            T af[TM], bf[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) af[i] = as[LA::at(wm0 + i * 8 + fr, kk + fk)];
#pragma unroll
            for (int j = 0; j < TN; ++j) bf[j] = bs[LB::at(wn0 + j * 8 + fr, kk + fk)];
            if constexpr (CPLX) {
                // four passes so that DMMAs hitting the same accumulator are TM*TN instructions apart
                double ai[TM], bi[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) ai[i] = sa * af[i].y;
#pragma unroll
                for (int j = 0; j < TN; ++j) bi[j] = sb * bf[j].y;
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) dmma884(cr[i][j][0], cr[i][j][1], af[i].x, bf[j].x);
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) dmma884(ci[i][j][0], ci[i][j][1], af[i].x, bi[j]);
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) dmma884(cr[i][j][0], cr[i][j][1], -ai[i], bi[j]);
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) dmma884(ci[i][j][0], ci[i][j][1], ai[i], bf[j].x);
            } else {
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) dmma884(cr[i][j][0], cr[i][j][1], af[i], bf[j]);
            }
        }
    }
    cp_async_wait<0>();

    if (p.ksplit > 1) {   // split-K: raw partial tile -> work[z][ks][m][n]; alpha/beta applied by the reduce kernel
        T *W = reinterpret_cast<T *>(p.work) + ((size_t)z * p.ksplit + ks) * (size_t)p.M * p.N;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int64_t m = m0 + wm0 + i * 8 + fr;
            if (m >= p.M) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int64_t n = n0 + wn0 + j * 8 + fk * 2 + q;
                    if (n >= p.N) continue;
                    if constexpr (CPLX) W[m * p.N + n] = make_cplx(cr[i][j][q], ci[i][j][q]);
                    else W[m * p.N + n] = cr[i][j][q];
                }
        }
        return;
    }
    // epilogue: thread owns rows fr (+8i), column pairs 2*fk (+8j) of its warp tile.  With beta != 0 the old values
    // of C are read for a whole row of the thread's fragment (2 TN independent loads in flight) BEFORE anything of
    // that row is written: interleaving load / store element by element serialises the L2 round trips (the
    // compiler cannot hoist a load of C above a store to C), which cost the skinny HBM-bound contractions 40 % of
    // their time and the 840^3 ring terms about a fifth.
    const bool has_beta = (p.beta_re != 0.0) || (p.beta_im != 0.0);
    int64_t on[TN][2];
    bool nok[TN][2];
#pragma unroll
    for (int j = 0; j < TN; ++j)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int64_t n = n0 + wn0 + j * 8 + fk * 2 + q;
            nok[j][q] = n < p.N;
            on[j][q] = nok[j][q] ? p.c_n[n] : 0;
        }
    int64_t om[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t m = m0 + wm0 + i * 8 + fr;
        om[i] = (m < p.M) ? p.c_m[m] : -1;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        if (om[i] < 0) continue;
        T *row = C + om[i];
        T oldv[TN][2];
        if (has_beta) {
#pragma unroll
            for (int j = 0; j < TN; ++j)
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (nok[j][q]) oldv[j][q] = row[on[j][q]];
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (!nok[j][q]) continue;
                T *dst = row + on[j][q];
                if constexpr (CPLX) {
                    double xr = cr[i][j][q], xi = ci[i][j][q];
                    double vr = p.alpha_re * xr - p.alpha_im * xi, vi = p.alpha_re * xi + p.alpha_im * xr;
                    if (has_beta) {
                        const cplx o = oldv[j][q];
                        vr += p.beta_re * o.x - p.beta_im * o.y;
                        vi += p.beta_re * o.y + p.beta_im * o.x;
                    }
                    *dst = make_cplx(vr, vi);
                } else {
                    double v = p.alpha_re * cr[i][j][q];
                    if (has_beta) v += p.beta_re * oldv[j][q];
                    *dst = v;
                }
            }
        }
    }
}

// Dot-product-like contractions (M*N <= 16 outputs: norms <t|t>, energy-like sums, the scalar and
// 3-vector intermediates of the AAT assembly): a DMMA tile would be 99 % padding and one CTA would
// walk K slab by slab at L2 latency (60 us for K = 840).  Here every thread strides over k, keeps
// the M*N partial sums in registers and the block reduces them in fixed order.
//   grid = (ksplit, batch); ksplit > 1: raw partials -> work[z][ks][m][n] (then splitk_reduce_kernel).
constexpr int kDotThreads = 256;
constexpr int kDotMaxOut = 16;
template <bool CPLX>
__global__ void __launch_bounds__(kDotThreads) contract_dot_kernel(const ContractArgs p) {
    using T = typename elem_t<CPLX>::type;
    const int z = blockIdx.y, ks = blockIdx.x;
    if (p.active != nullptr && p.active[z] == 0) return;
    const T *A = reinterpret_cast<const T *>(p.A) + (size_t)z * p.a_bs;
    const T *B = reinterpret_cast<const T *>(p.B) + (size_t)z * p.b_bs;
    const int M = (int)p.M, N = (int)p.N, MN = M * N;
    const int64_t per = (p.K + p.ksplit - 1) / p.ksplit;
    const int64_t k0 = (int64_t)ks * per;
    int64_t k1 = k0 + per;
    if (k1 > p.K) k1 = p.K;
    double accr[kDotMaxOut], acci[CPLX ? kDotMaxOut : 1];
#pragma unroll
    for (int e = 0; e < kDotMaxOut; ++e) { accr[e] = 0.0; if (CPLX) acci[e] = 0.0; }
    const double sa = p.conj_a ? -1.0 : 1.0, sb = p.conj_b ? -1.0 : 1.0;
    int64_t am[kDotMaxOut], bn[kDotMaxOut];            // per-output operand offsets, hoisted out of the k loop
#pragma unroll
    for (int e = 0; e < kDotMaxOut; ++e) {
        am[e] = (e < MN) ? p.a_m[e / N] : 0;
        bn[e] = (e < MN) ? p.b_n[e % N] : 0;
    }
    for (int64_t k = k0 + threadIdx.x; k < k1; k += kDotThreads) {
        const int64_t ak = p.a_k[k], bk = p.b_k[k];
#pragma unroll
        for (int e = 0; e < kDotMaxOut; ++e) {
            if (e >= MN) break;
            const T av = A[am[e] + ak], bv = B[bk + bn[e]];     // repeats hit L1
            if constexpr (CPLX) {
                const double ar = av.x, ai = sa * av.y, br = bv.x, bi = sb * bv.y;
                accr[e] = fma(ar, br, fma(-ai, bi, accr[e]));
                acci[e] = fma(ar, bi, fma(ai, br, acci[e]));
            } else {
                accr[e] = fma(av, bv, accr[e]);
            }
        }
    }
    // fixed-order block reduction, one output at a time (MN <= 16)
    __shared__ double red[2][kDotThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool has_beta = (p.beta_re != 0.0) || (p.beta_im != 0.0);
    T *C = reinterpret_cast<T *>(p.C) + (size_t)z * p.c_bs;
    T *W = reinterpret_cast<T *>(p.work) + ((size_t)z * p.ksplit + ks) * (size_t)MN;
#pragma unroll
    for (int e = 0; e < kDotMaxOut; ++e) {
        if (e >= MN) break;
        double xr = warp_sum(accr[e]), xi = CPLX ? warp_sum(acci[e]) : 0.0;
        if (lane == 0) { red[0][w] = xr; red[1][w] = xi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            xr = 0.0; xi = 0.0;
            for (int q = 0; q < kDotThreads / 32; ++q) { xr += red[0][q]; xi += red[1][q]; }
            if (p.ksplit > 1) {
                if constexpr (CPLX) W[e] = make_cplx(xr, xi); else W[e] = xr;
            } else {
                T *dst = C + p.c_m[e / N] + p.c_n[e % N];
                if constexpr (CPLX) {
                    double vr = p.alpha_re * xr - p.alpha_im * xi, vi = p.alpha_re * xi + p.alpha_im * xr;
                    if (has_beta) { const cplx o = *dst; vr += p.beta_re * o.x - p.beta_im * o.y; vi += p.beta_re * o.y + p.beta_im * o.x; }
                    *dst = make_cplx(vr, vi);
                } else {
                    double v = p.alpha_re * xr;
                    if (has_beta) v += p.beta_re * (*dst);
                    *dst = v;
                }
            }
        }
        __syncthreads();
    }
}

// C[c_m[m] + c_n[n]] = alpha * sum_ks work[z][ks][m][n] + beta * C   (fixed summation order)
template <bool CPLX>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const ContractArgs p, int batch) {
    using T = typename elem_t<CPLX>::type;
    const int64_t MN = p.M * p.N, total = MN * batch;
    const bool has_beta = (p.beta_re != 0.0) || (p.beta_im != 0.0);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t z = e / MN, mn = e % MN, m = mn / p.N, n = mn % p.N;
        if (p.active != nullptr && p.active[z] == 0) continue;
        const T *W = reinterpret_cast<const T *>(p.work) + (size_t)z * p.ksplit * MN + mn;
        T *dst = reinterpret_cast<T *>(p.C) + (size_t)z * p.c_bs + p.c_m[m] + p.c_n[n];
        if constexpr (CPLX) {
            double xr = 0.0, xi = 0.0;
            for (int ks = 0; ks < p.ksplit; ++ks) { const cplx w = W[(size_t)ks * MN]; xr += w.x; xi += w.y; }
            double vr = p.alpha_re * xr - p.alpha_im * xi, vi = p.alpha_re * xi + p.alpha_im * xr;
            if (has_beta) { const cplx o = *dst; vr += p.beta_re * o.x - p.beta_im * o.y; vi += p.beta_re * o.y + p.beta_im * o.x; }
            *dst = make_cplx(vr, vi);
        } else {
            double x = 0.0;
            for (int ks = 0; ks < p.ksplit; ++ks) x += W[(size_t)ks * MN];
            double v = p.alpha_re * x;
            if (has_beta) v += p.beta_re * (*dst);
            *dst = v;
        }
    }
}

template <bool CPLX, int BM, int BN, int BK, int WM, int WN, int STAGES, bool AKF, bool BKF>
static int launch_contract2(const ContractArgs &a, int batch, cudaStream_t st) {
    using T = typename elem_t<CPLX>::type;
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    constexpr size_t smem = (size_t)STAGES * (TileLayout<CPLX, AKF, BM, BK>::size + TileLayout<CPLX, BKF, BN, BK>::size) * sizeof(T);
    auto kern = contract_kernel<CPLX, BM, BN, BK, WM, WN, STAGES, AKF, BKF>;
    static bool attr_set = false;
    if (!attr_set) {
        APYIB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    dim3 grid((unsigned)((a.N + BN - 1) / BN), (unsigned)((a.M + BM - 1) / BM), (unsigned)(batch * a.ksplit));
    kern<<<grid, NT, smem, st>>>(a);
    APYIB_LAUNCH_CHECK();
    if (a.ksplit > 1) {
        int64_t b = (a.M * a.N * batch + 255) / 256;
        if (b > 148 * 8) b = 148 * 8;
        splitk_reduce_kernel<CPLX><<<(unsigned)b, 256, 0, st>>>(a, batch);
        APYIB_LAUNCH_CHECK();
    }
    return APYIB_OK;
}

template <bool CPLX, int BM, int BN, int BK, int WM, int WN, int STAGES>
static int launch_contract(const ContractArgs &a, int batch, cudaStream_t st) {
    if (a.a_kfast) {
        return a.b_kfast ? launch_contract2<CPLX, BM, BN, BK, WM, WN, STAGES, true, true>(a, batch, st)
                         : launch_contract2<CPLX, BM, BN, BK, WM, WN, STAGES, true, false>(a, batch, st);
    }
    return a.b_kfast ? launch_contract2<CPLX, BM, BN, BK, WM, WN, STAGES, false, true>(a, batch, st)
                     : launch_contract2<CPLX, BM, BN, BK, WM, WN, STAGES, false, false>(a, batch, st);
}

}  // namespace apyib

using namespace apyib;

extern "C" int apyib_contract(int dtype, const void *d_A, const void *d_B, void *d_C, int64_t M, int64_t N,
                              int64_t K, const int64_t *d_a_m, const int64_t *d_a_k, const int64_t *d_b_k,
                              const int64_t *d_b_n, const int64_t *d_c_m, const int64_t *d_c_n, int a_kfast,
                              int b_kfast, int conj_a, int conj_b, double alpha_re, double alpha_im,
                              double beta_re, double beta_im, int batch, int64_t a_bstride,
                              int64_t b_bstride, int64_t c_bstride, const int32_t *d_active, int ksplit,
                              void *d_work, void *stream) {
    APYIB_REQUIRE(ksplit >= 1 && (ksplit == 1 || d_work != nullptr) && (int64_t)batch * ksplit <= 65535, "split-K");
    APYIB_REQUIRE(dtype == APYIB_F64 || dtype == APYIB_C128, "dtype");
    APYIB_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batch >= 1 && batch <= 65535, "sizes");
    APYIB_REQUIRE(d_A && d_B && d_C && d_a_m && d_a_k && d_b_k && d_b_n && d_c_m && d_c_n, "null pointer");
    if (M == 0 || N == 0) return APYIB_OK;
    APYIB_REQUIRE((M + 31) / 32 <= 65535, "M too large for grid.y");
    ContractArgs a;
    a.A = d_A; a.B = d_B; a.C = d_C;
    a.a_m = d_a_m; a.a_k = d_a_k; a.b_k = d_b_k; a.b_n = d_b_n; a.c_m = d_c_m; a.c_n = d_c_n;
    a.M = M; a.N = N; a.K = K;
    a.a_bs = a_bstride; a.b_bs = b_bstride; a.c_bs = c_bstride;
    a.active = d_active;
    a.alpha_re = alpha_re; a.alpha_im = alpha_im; a.beta_re = beta_re; a.beta_im = beta_im;
    a.a_kfast = a_kfast; a.b_kfast = b_kfast; a.conj_a = conj_a; a.conj_b = conj_b;
    a.ksplit = ksplit; a.work = d_work;
    cudaStream_t st = (cudaStream_t)stream;
    if (M * N <= kDotMaxOut && K >= 64) {      // dot-product-like: strided-k reduction, no tiles
        dim3 grid((unsigned)ksplit, (unsigned)batch);
        if (dtype == APYIB_C128) contract_dot_kernel<true><<<grid, kDotThreads, 0, st>>>(a);
        else contract_dot_kernel<false><<<grid, kDotThreads, 0, st>>>(a);
        APYIB_LAUNCH_CHECK();
        if (ksplit > 1) {
            if (dtype == APYIB_C128) splitk_reduce_kernel<true><<<1, 256, 0, st>>>(a, batch);
            else splitk_reduce_kernel<false><<<1, 256, 0, st>>>(a, batch);
            APYIB_LAUNCH_CHECK();
        }
        return APYIB_OK;
    }
    // Skinny shapes (T1 <-> T2 couplings of ci_wfn.py:457-470, Fock-like terms :471-474, J/K builds): one
    // side is <= 16 wide, the other operand is streamed once -> HBM-bound.  A 64-wide tile would spend
    // >= 75 % of its DMMAs and of its B-tile loads on padding, so these get N (or M) = 16 tiles.
    if (N <= 16 && M > 16) {
        return dtype == APYIB_C128 ? launch_contract<true, 64, 16, 16, 16, 16, 3>(a, batch, st)
                                   : launch_contract<false, 128, 16, 16, 32, 16, 3>(a, batch, st);
    }
    if (M <= 16 && N > 16) {
        return dtype == APYIB_C128 ? launch_contract<true, 16, 64, 16, 16, 16, 3>(a, batch, st)
                                   : launch_contract<false, 16, 128, 16, 16, 32, 3>(a, batch, st);
    }
    // tile choice: 64x64 when that already yields >= 1 wave of CTAs on 148 SMs, else 32x32
    const int64_t big_tiles = ((M + 63) / 64) * ((N + 63) / 64) * batch;
    const bool big = big_tiles >= 148 && ksplit == 1;
    if (dtype == APYIB_C128) {
        return big ? launch_contract<true, 64, 64, 8, 32, 32, 3>(a, batch, st)
                   : launch_contract<true, 32, 32, 16, 16, 16, 4>(a, batch, st);
    }
    // small problems are latency bound (few CTAs, each walking K alone): deeper pipeline, longer slabs
    return big ? launch_contract<false, 64, 64, 16, 32, 32, 3>(a, batch, st)
               : launch_contract<false, 32, 32, 32, 16, 16, 4>(a, batch, st);
}
