// TMA-fed variant of the DMMA contraction kernel for operands that are plain strided matrices
// with a contiguous k index (the ladder term  r_ijab += <ab|cd> t_ijcd,  the first AO->MO
// quarter transform, ...):  A[m, k] at A + m*lda + k,  B[n, k] at B + n*ldb + k.
//
// Tiles arrive by `cp.async.bulk.tensor` (TMA; SASS UTMALDG) into a 4-stage shared-memory ring,
// completion is signalled on mbarriers, the tensor maps carry the batch (finite-difference point)
// as third dimension and zero-fill the M/N/K tails in hardware.  Tiles are stored densely with the
// 128-byte hardware swizzle (16-byte chunk index XOR row%8), so the DMMA fragment reads need no
// padding; the FP64 arithmetic (four real DMMAs per complex product) and the offset-table epilogue
// are those of contract.cu.  One elected thread issues the loads; the address generation that
// costs the gather kernel ~70 integer instructions per thread and slab disappears.
#include <cuda.h>
#include <type_traits>
#include "common.cuh"

namespace apyib {

struct TmaArgs {
    void *C;
    const int64_t *c_m, *c_n;
    int64_t M, N, K, c_bs;
    const int32_t *active;
    double alpha_re, alpha_im, beta_re, beta_im;
    int conj_a, conj_b;
};

__device__ __forceinline__ void dmma884_t(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *map, unsigned bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// 64 x BN CTA tile (BN = 48, 64 or 80), 4 warps of 32 x BN/2, slab = 128 bytes of k (8 complex / 16 real),
// STAGES-deep ring.  BN = 48 exists for the ladder term: N = o^2 = 144 is 3 x 48 but 2.25 x 64, i.e. a
// quarter of the DMMAs of a 64-wide tiling would multiply padding.
template <bool CPLX, int STAGES, int BN>
__global__ void __launch_bounds__(128)
contract_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TmaArgs p) {
    constexpr int BM = 64, WM = 32, WN = BN / 2, TM = 4, TN = WN / 8;
    constexpr int BK = CPLX ? 8 : 16;                     // elements per 128-byte row
    constexpr unsigned TILE_BYTES = 64 * 128;             // A tile
    constexpr unsigned TILE_BYTES_B = BN * 128;           // B tile (multiple of 1024: swizzle atom)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1024-byte aligned tiles (required by the 128B swizzle), barriers after them
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *As = base, *Bs = base + STAGES * TILE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + STAGES * (TILE_BYTES + TILE_BYTES_B));

    const int z = blockIdx.z;
    if (p.active != nullptr && p.active[z] == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp >> 1) * WM, wn0 = (warp & 1) * WN;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int64_t nslab = (p.K + BK - 1) / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int64_t slab) {
        const int s = (int)(slab % STAGES);
        const unsigned bar = smem_u32(&bars[s]);
        mbar_expect_tx(bar, TILE_BYTES + TILE_BYTES_B);
        const int kcoord = (int)(slab * BK) * (CPLX ? 2 : 1);       // inner coordinate in doubles
        tma_load_3d(smem_u32(As + s * TILE_BYTES), &mapA, bar, kcoord, (int)m0, z);
        tma_load_3d(smem_u32(Bs + s * TILE_BYTES_B), &mapB, bar, kcoord, (int)n0, z);
    };
    if (tid == 0)
        for (int s = 0; s < STAGES && s < nslab; ++s) issue(s);

    double cr[TM][TN][2];
    double ci[CPLX ? TM : 1][CPLX ? TN : 1][2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            cr[i][j][0] = cr[i][j][1] = 0.0;
            if (CPLX) ci[i][j][0] = ci[i][j][1] = 0.0;
        }
    const int fr = lane >> 2, fk = lane & 3;
    const double sa = p.conj_a ? -1.0 : 1.0, sb = p.conj_b ? -1.0 : 1.0;

    for (int64_t kt = 0; kt < nslab; ++kt) {
        const int s = (int)(kt % STAGES);
        mbar_wait(smem_u32(&bars[s]), (unsigned)((kt / STAGES) & 1));
        const unsigned char *as = As + s * TILE_BYTES, *bs = Bs + s * TILE_BYTES_B;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            if constexpr (CPLX) {
                cplx af[TM], bf[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    const int r = wm0 + i * 8 + fr;
                    af[i] = *reinterpret_cast<const cplx *>(as + r * 128 + (((kk + fk) ^ (r & 7)) << 4));
                }
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    const int r = wn0 + j * 8 + fr;
                    bf[j] = *reinterpret_cast<const cplx *>(bs + r * 128 + (((kk + fk) ^ (r & 7)) << 4));
                }
                double ai[TM], bi[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) ai[i] = sa * af[i].y;
#pragma unroll
                for (int j = 0; j < TN; ++j) bi[j] = sb * bf[j].y;
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) dmma884_t(cr[i][j][0], cr[i][j][1], af[i].x, bf[j].x);
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) dmma884_t(ci[i][j][0], ci[i][j][1], af[i].x, bi[j]);
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) dmma884_t(cr[i][j][0], cr[i][j][1], -ai[i], bi[j]);
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) dmma884_t(ci[i][j][0], ci[i][j][1], ai[i], bf[j].x);
            } else {
                double af[TM], bf[TN];
                const int k = kk + fk;
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    const int r = wm0 + i * 8 + fr;
                    af[i] = *reinterpret_cast<const double *>(as + r * 128 + (((k >> 1) ^ (r & 7)) << 4) + ((k & 1) << 3));
                }
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    const int r = wn0 + j * 8 + fr;
                    bf[j] = *reinterpret_cast<const double *>(bs + r * 128 + (((k >> 1) ^ (r & 7)) << 4) + ((k & 1) << 3));
                }
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) dmma884_t(cr[i][j][0], cr[i][j][1], af[i], bf[j]);
            }
        }
        __syncthreads();                                   // everyone is done with stage s
        if (tid == 0 && kt + STAGES < nslab) issue(kt + STAGES);
    }

    // epilogue (see contract.cu): all old values of a fragment row are loaded before the row is written
    const bool has_beta = (p.beta_re != 0.0) || (p.beta_im != 0.0);
    using T = typename std::conditional<CPLX, cplx, double>::type;
    T *C = reinterpret_cast<T *>(p.C) + (size_t)z * p.c_bs;
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
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t m = m0 + wm0 + i * 8 + fr;
        if (m >= p.M) continue;
        T *row = C + p.c_m[m];
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// rows x kdoubles matrix of doubles (complex = 2 doubles), `batch` of them
static bool make_map(CUtensorMap *map, const void *base, int64_t kdoubles, int64_t rows, int64_t ld_bytes, int batch,
                     int64_t bs_bytes, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)kdoubles, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld_bytes, (cuuint64_t)(batch > 1 ? bs_bytes : ld_bytes * rows)};
    cuuint32_t box[3] = {16, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace apyib

using namespace apyib;

// Same contract as apyib_contract for operands that are k-contiguous strided matrices:
//   opA(A[z*a_bs + m*lda + k]),  opB(B[z*b_bs + n*ldb + k]);  C through offset tables as usual.
// Returns APYIB_ERR_UNSUPPORTED (nothing launched) if the shapes / alignments do not fit TMA.
extern "C" int apyib_contract_tma(int dtype, const void *d_A, const void *d_B, void *d_C, int64_t M, int64_t N,
                                  int64_t K, int64_t lda, int64_t ldb, const int64_t *d_c_m, const int64_t *d_c_n,
                                  int conj_a, int conj_b, double alpha_re, double alpha_im, double beta_re,
                                  double beta_im, int batch, int64_t a_bstride, int64_t b_bstride, int64_t c_bstride,
                                  const int32_t *d_active, void *stream) {
    APYIB_REQUIRE(dtype == APYIB_F64 || dtype == APYIB_C128, "dtype");
    APYIB_REQUIRE(d_A && d_B && d_C && d_c_m && d_c_n, "null pointer");
    APYIB_REQUIRE(M > 0 && N > 0 && K > 0 && batch >= 1 && batch <= 65535, "sizes");
    const int64_t es = dtype == APYIB_C128 ? 16 : 8;
    const bool ok = ((uintptr_t)d_A % 16 == 0) && ((uintptr_t)d_B % 16 == 0) && (lda * es) % 16 == 0 && (ldb * es) % 16 == 0 &&
                    (a_bstride * es) % 16 == 0 && (b_bstride * es) % 16 == 0 && lda >= K && ldb >= K &&
                    (batch == 1 || (a_bstride > 0 && b_bstride > 0)) && M < (1LL << 31) && N < (1LL << 31) &&
                    K * 2 < (1LL << 31) && (M + 63) / 64 <= 65535;
    if (!ok) {
        set_error("apyib_contract_tma: operands do not meet the TMA alignment / shape rules");
        return APYIB_ERR_UNSUPPORTED;
    }
    CUtensorMap mapA, mapB;
    const int64_t kd = K * (dtype == APYIB_C128 ? 2 : 1);
    // B-tile width: 48, 64 or 80 columns, whichever pads N least (ties: the wider tile).  N = o^2 = 144 is 3 x 48;
    // the pair-packed ladder has N = o(o+1)/2 = 78 -> one 80-wide tile instead of 2 x 48 (19 % padding).
    int bn = 64;
    {
        int64_t best = ((N + 63) / 64) * 64;
        const int cand[2] = {80, 48};
        for (int c : cand) {
            const int64_t padded = ((N + c - 1) / c) * c;
            if (padded < best || (padded == best && c > bn)) { best = padded; bn = c; }
        }
    }
    if (!make_map(&mapA, d_A, kd, M, lda * es, batch, a_bstride * es, 64) ||
        !make_map(&mapB, d_B, kd, N, ldb * es, batch, b_bstride * es, bn)) {
        set_error("apyib_contract_tma: cuTensorMapEncodeTiled failed or is unavailable");
        return APYIB_ERR_UNSUPPORTED;
    }
    TmaArgs a;
    a.C = d_C; a.c_m = d_c_m; a.c_n = d_c_n; a.M = M; a.N = N; a.K = K; a.c_bs = c_bstride; a.active = d_active;
    a.alpha_re = alpha_re; a.alpha_im = alpha_im; a.beta_re = beta_re; a.beta_im = beta_im;
    a.conj_a = conj_a; a.conj_b = conj_b;
    constexpr int STAGES = 4;
    const size_t smem = (size_t)STAGES * (64 + bn) * 128 + STAGES * 8 + 1024;
    dim3 grid((unsigned)((N + bn - 1) / bn), (unsigned)((M + 63) / 64), (unsigned)batch);
    cudaStream_t st = (cudaStream_t)stream;
#define APYIB_TMA_LAUNCH(CP, BNN)                                                                                  \
    do {                                                                                                           \
        static bool attr_done = false;                                                                             \
        if (!attr_done) {                                                                                          \
            APYIB_CUDA_CHECK(cudaFuncSetAttribute(contract_tma_kernel<CP, STAGES, BNN>,                            \
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize,                      \
                                                  (int)(STAGES * (64 + BNN) * 128 + STAGES * 8 + 1024)));           \
            attr_done = true;                                                                                      \
        }                                                                                                          \
        contract_tma_kernel<CP, STAGES, BNN><<<grid, 128, smem, st>>>(mapA, mapB, a);                               \
    } while (0)
    if (dtype == APYIB_C128) {
        if (bn == 48) APYIB_TMA_LAUNCH(true, 48); else if (bn == 80) APYIB_TMA_LAUNCH(true, 80); else APYIB_TMA_LAUNCH(true, 64);
    } else {
        if (bn == 48) APYIB_TMA_LAUNCH(false, 48); else if (bn == 80) APYIB_TMA_LAUNCH(false, 80); else APYIB_TMA_LAUNCH(false, 64);
    }
#undef APYIB_TMA_LAUNCH
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}
