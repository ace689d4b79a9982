// HBM-streaming kernels of the hot path: block gather / spin blocking / antisymmetrisation,
// MP2 amplitudes + energy, CI Jacobi update, DIIS dots / extrapolation with fused energy and
// rms reductions.  All reductions are deterministic (fixed summation order, no float atomics).
#include "common.cuh"

namespace apyib {

constexpr int kThreads = 256;

// --------------------------------------------------------------------------------------
// gather4 / gather2
// --------------------------------------------------------------------------------------
struct Gather4Args {
    int64_t sd[4];      // source dims (spatial)
    int64_t od[4];      // output dims
    int perm1[4], perm2[4];
    int64_t st1[4], st2[4];
    double c1, c2;
    int spin;
};

template <typename T>
__device__ __forceinline__ T fetch4(const T *src, const int64_t (&sd)[4], int spin, int64_t p0, int64_t p1,
                                    int64_t p2, int64_t p3) {
    if (spin) {
        if (((p0 ^ p1) & 1) || ((p2 ^ p3) & 1)) return scalar<T>::zero();
        p0 >>= 1; p1 >>= 1; p2 >>= 1; p3 >>= 1;
    }
    return src[((p0 * sd[1] + p1) * sd[2] + p2) * sd[3] + p3];
}

// One CTA per leading-index group (x0, x1[, x2]); the trailing `inner` output elements are walked with
// 32-bit index arithmetic and four independent elements in flight per thread.  (The first version
// decomposed a flat int64 index per element: three 64-bit div/mod pairs per output made the one-off block
// extraction run at 31 % of HBM peak.)
template <typename T> __global__ void __launch_bounds__(kThreads) gather4_kernel(const T *__restrict__ src, T *__restrict__ out, Gather4Args g, int lead, int64_t src_bstride, int64_t out_bstride) {
    src += (size_t)blockIdx.y * src_bstride;                  // blockIdx.y = finite-difference point of a stack
    out += (size_t)blockIdx.y * out_bstride;
    const int64_t nouter = (lead == 3) ? g.od[0] * g.od[1] * g.od[2] : g.od[0] * g.od[1];
    const unsigned od3 = (unsigned)g.od[3];
    const unsigned inner = (lead == 3) ? od3 : (unsigned)(g.od[2] * g.od[3]);
    const bool two = g.c2 != 0.0;
    for (int64_t outer = blockIdx.x; outer < nouter; outer += gridDim.x) {
        int64_t x[4];
        int64_t r = outer;
        if (lead == 3) { x[2] = r % g.od[2]; r /= g.od[2]; } else { x[2] = 0; }
        x[1] = r % g.od[1];
        x[0] = r / g.od[1];
        T *dst = out + outer * inner;
        for (unsigned e0 = threadIdx.x; e0 < inner; e0 += 4 * kThreads) {
            T v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned e = e0 + u * kThreads;
                if (e >= inner) break;
                int64_t y[4] = {x[0], x[1], x[2], 0};
                if (lead == 3) { y[3] = e; } else { const unsigned q = e / od3; y[2] = q; y[3] = e - q * od3; }
                T a = fetch4<T>(src, g.sd, g.spin, g.st1[0] + y[g.perm1[0]], g.st1[1] + y[g.perm1[1]],
                                g.st1[2] + y[g.perm1[2]], g.st1[3] + y[g.perm1[3]]);
                a = g.c1 * a;
                if (two) {
                    const T w = fetch4<T>(src, g.sd, g.spin, g.st2[0] + y[g.perm2[0]], g.st2[1] + y[g.perm2[1]],
                                          g.st2[2] + y[g.perm2[2]], g.st2[3] + y[g.perm2[3]]);
                    a = a + g.c2 * w;
                }
                v[u] = a;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned e = e0 + u * kThreads;
                if (e < inner) dst[e] = v[u];
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
gather2_kernel(const T *src, T *out, int64_t s0, int64_t s1, int64_t o0, int64_t o1, int pa, int pb, int64_t st0,
               int64_t st1, int spin) {
    const int64_t total = o0 * o1;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t x[2] = {idx / o1, idx % o1};
        int64_t p = st0 + x[pa], q = st1 + x[pb];
        T v = scalar<T>::zero();
        if (spin) {
            if (((p ^ q) & 1) == 0) v = src[(p >> 1) * s1 + (q >> 1)];
        } else {
            v = src[p * s1 + q];
        }
        out[idx] = v;
    }
}

// --------------------------------------------------------------------------------------
// MP2  (mp2_wfn.py:42-59, 64-87)
//
// t2[i,j,a,b] = (ai|bj)/D is a transpose of the MO integrals: a one-thread-per-output kernel
// reads 16-byte elements at stride n (half of every 32-byte sector wasted, 29 % of HBM peak
// measured).  The tiled kernels below read only runs that are contiguous in the source
// (>= o elements = 192 B at o = 12), transpose through shared memory and write t2 in runs of v
// elements.  The exchange part of the energy is relabelled (dummy indices a <-> b, exact):
//     E = sum (2<ij|ab> - <ij|ba>) t_ijab = sum_ijab (ia|jb) (2 t_ijab - t_ijba)
// so that every integral read is contiguous as well.
// --------------------------------------------------------------------------------------
constexpr int kMp2Threads = 256;

// spatial: one CTA per (i, a); (j, b) tiles staged in shared memory.
//   s1[b][j] = (a i | b j)   -> t_ijab        s2[b][j] = (b i | a j) -> t_ijba
// SO_ENERGY: accumulate the spin-orbital energy 1/4 sum <IJ||AB> t_IJAB instead, spin-summed analytically
// and relabelled so that all reads stay contiguous:
//     E_SO = sum_ijab [ t_ijab ((ia|jb) - 1/2 (ja|ib)) + t_ijba ((ja|ib) - 1/2 (ia|jb)) ]
// HERM: the caller guarantees (pq|rs) = conj((qp|sr)) -- true for every MO tensor transformed from real
// AO integrals (utils.py:274-277) -- so (ia|jb) = conj((ai|bj)) and (ja|ib) = conj((bi|aj)) are already in
// the shared-memory tiles and the energy needs no second pass over the integrals: 2 reads + 1 write per
// amplitude, the algorithmic minimum of SURVEY 8(d) U2.
template <typename T, bool SO_ENERGY, bool HERM>
__global__ void __launch_bounds__(kMp2Threads)
mp2_spatial_kernel(const T *__restrict__ eri, int64_t n, int64_t o, const double *__restrict__ eps, T *__restrict__ t2,
                   double *E_out, double *partials, int bt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int64_t v = n - o;
    const int oi = (int)o;
    const int ld = oi | 1;                               // odd pitch: conflict-free transposed reads
    T *s1 = reinterpret_cast<T *>(smem_raw), *s2 = s1 + (size_t)bt * ld;
    double acc[2] = {0.0, 0.0};
    // element e of a tile is (e / o, e % o) in the load phase and (e / nb, e % nb) in the compute phase;
    // both are advanced incrementally (no integer divisions in the loops)
    const int l_db = kMp2Threads / oi, l_dj = kMp2Threads % oi;
    for (int64_t blk = blockIdx.x; blk < o * v; blk += gridDim.x) {
        const int64_t i = blk / v, a = blk % v, A = o + a;
        const T *row1 = eri + ((A * n + i) * n + o) * n;          // + b*n + j : (a i | b j)
        const T *row2 = eri + (o * n + i) * n * n + A * n;         // + b*n^3 + j : (b i | a j)
        const T *rowg = eri + ((i * n + A) * n) * n + o;           // + j*n + b : (i a | j b)
        const T *rowh = eri + A * n * n + i * n + o;               // + j*n^3 + b : (j a | i b)
        const double Dia = eps[i] - eps[A];
        for (int64_t b0 = 0; b0 < v; b0 += bt) {
            const int nb = (int)((v - b0 < bt) ? (v - b0) : bt);
            const int total = nb * oi;
            {
                int b = threadIdx.x / oi, j = threadIdx.x % oi;
                for (int e0 = threadIdx.x; e0 < total; e0 += 4 * kMp2Threads) {
                    T x1[4], x2[4];
                    int bb[4], jj[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {                              // 8 independent loads in flight
                        bb[u] = b; jj[u] = j;
                        if (e0 + u * kMp2Threads < total) {
                            x1[u] = row1[(b0 + b) * n + j];
                            x2[u] = row2[(b0 + b) * n * n * n + j];
                        }
                        b += l_db; j += l_dj;
                        if (j >= oi) { j -= oi; ++b; }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (e0 + u * kMp2Threads < total) {
                            s1[bb[u] * ld + jj[u]] = x1[u];
                            s2[bb[u] * ld + jj[u]] = x2[u];
                        }
                }
            }
            __syncthreads();
            {
                const int c_dj = kMp2Threads / nb, c_db = kMp2Threads % nb;
                int j = threadIdx.x / nb, b = threadIdx.x % nb;
                for (int e0 = threadIdx.x; e0 < total; e0 += 4 * kMp2Threads) {    // b fastest: runs of nb
                    T g[4], g2[4];
                    int bb[4], jj[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {                                  // integral loads first (MLP)
                        bb[u] = b; jj[u] = j;
                        if (!HERM && e0 + u * kMp2Threads < total) {
                            g[u] = rowg[(int64_t)j * n + b0 + b];                  // (ia|jb), contiguous in b
                            if (SO_ENERGY) g2[u] = rowh[(int64_t)j * n * n * n + b0 + b];   // (ja|ib)
                        }
                        j += c_dj; b += c_db;
                        if (b >= nb) { b -= nb; ++j; }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (e0 + u * kMp2Threads >= total) continue;
                        const int64_t B = b0 + bb[u];
                        const double rD = 1.0 / (Dia + eps[jj[u]] - eps[o + B]);
                        const T k1 = s1[bb[u] * ld + jj[u]], k2 = s2[bb[u] * ld + jj[u]];
                        if (HERM) { g[u] = scalar<T>::cj(k1); g2[u] = scalar<T>::cj(k2); }
                        const T t = rD * k1;
                        const T tx = rD * k2;
                        t2[((i * o + jj[u]) * v + a) * v + B] = t;
                        T ev;
                        if (SO_ENERGY) ev = t * (g[u] - 0.5 * g2[u]) + tx * (g2[u] - 0.5 * g[u]);
                        else ev = g[u] * (2.0 * t - tx);
                        acc[0] += scalar<T>::re(ev);
                        acc[1] += scalar<T>::im(ev);
                    }
                }
            }
            __syncthreads();
        }
    }
    grid_sum_finish<2, kMp2Threads>(acc, partials, [&](double(&tot)[2]) {
        E_out[0] = tot[0];
        E_out[1] = tot[1];
    });
}

// spin-orbital amplitudes from the spatial ones (one CTA per (I, J), B fastest):
//   t_IJAB = [(AI|BJ) - (AJ|BI)]/D = [sA==sI][sB==sJ] t_ijab - [sA==sJ][sB==sI] t_jiab
// with (PQ|RS) = (pq|rs)[sP==sQ][sR==sS] (utils.py:317-365; alpha = even, beta = odd).  Pure
// streaming: 16 o^2 v^2 elements written in runs of V, the spatial t2 is read in runs of v.
template <typename T> struct pair_t { T lo, hi; };
template <typename T>
__global__ void __launch_bounds__(kMp2Threads)
mp2_so_expand_kernel(const T *__restrict__ ts, int o, int v, T *__restrict__ t2) {
    const int O = 2 * o, V = 2 * v;
    const int d_a = kMp2Threads / v, d_b = kMp2Threads % v;
    for (int pr = blockIdx.x; pr < O * O; pr += gridDim.x) {
        const int I = pr / O, J = pr - I * O;
        const int i = I >> 1, si = I & 1, j = J >> 1, sj = J & 1;
        const T *tij = ts + ((int64_t)i * o + j) * v * v, *tji = ts + ((int64_t)j * o + i) * v * v;
        T *out = t2 + ((int64_t)I * O + J) * V * V;
        // every thread turns one spatial (a, b) into its four spin components: two 2-element stores
        const double c1 = 1.0, c2 = -1.0;
        int a = threadIdx.x / v, b = threadIdx.x % v;
        for (int e = threadIdx.x; e < v * v; e += kMp2Threads) {
            const T x1 = tij[e], x2 = tji[e];
            // value(sa, sb) = [sa==si][sb==sj] x1 - [sa==sj][sb==si] x2
            T val[2][2];
#pragma unroll
            for (int sa = 0; sa < 2; ++sa)
#pragma unroll
                for (int sb = 0; sb < 2; ++sb) {
                    T r = scalar<T>::zero();
                    if (sa == si && sb == sj) r = r + c1 * x1;
                    if (sa == sj && sb == si) r = r + c2 * x2;
                    val[sa][sb] = r;
                }
#pragma unroll
            for (int sa = 0; sa < 2; ++sa) {
                pair_t<T> p;
                p.lo = val[sa][0];
                p.hi = val[sa][1];
                *reinterpret_cast<pair_t<T> *>(out + (int64_t)(2 * a + sa) * V + 2 * b) = p;
            }
            a += d_a; b += d_b;
            if (b >= v) { b -= v; ++a; }
        }
    }
}

// --------------------------------------------------------------------------------------
// CI update  r <- r - E t ; t <- t + r / D
// --------------------------------------------------------------------------------------
// Streaming kernels keep U independent 16-byte loads per array and thread in flight: the loaded
// HBM latency on B200 is ~2 us, so one load per thread (32 KB per SM at full occupancy) caps a
// kernel at ~2.7 TB/s of reads (measured, profiles/r01_streaming_roofline_v1.json).
constexpr int kUnroll = 4;

template <typename T>
__global__ void __launch_bounds__(kThreads)
ci_update_kernel(T *__restrict__ r, T *__restrict__ t, const double *__restrict__ E, const double *__restrict__ eps_o,
                 const double *__restrict__ eps_v, int64_t o, int64_t v, int64_t n1, int64_t n2, int sh,
                 const int32_t *__restrict__ active, const double *__restrict__ E2, const double *__restrict__ E2off,
                 const T *__restrict__ tfix) {
    const int z = blockIdx.y;                      // finite-difference point of a batched solve
    if (active != nullptr && active[z] == 0) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, len = n1 + n2;
    r += z * len; t += z * len; E += z * 6; eps_o += z * (o >> sh); eps_v += z * (v >> sh);
    const T Ec = scalar<T>::make(E[0], E[1]);
    // linear-response form (analytic_aats.py:788-836): r <- r - E t - (E2 + E2off) t_fixed
    T E2c = scalar<T>::zero();
    if (tfix != nullptr) {
        tfix += z * len;
        E2c = scalar<T>::make(E2[z * 6] + E2off[z * 6], E2[z * 6 + 1] + E2off[z * 6 + 1]);
    }
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < len; base += kUnroll * stride) {
        T tv[kUnroll], rv[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t idx = base + u * stride;
            if (idx < len) { tv[u] = t[idx]; rv[u] = r[idx]; }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t idx = base + u * stride;
            if (idx >= len) continue;
            double D;
            if (idx < n1) {
                const int64_t a = idx % v, i = idx / v;
                D = eps_o[i >> sh] - eps_v[a >> sh];
            } else {
                int64_t q = idx - n1;
                const int64_t b = q % v; q /= v;
                const int64_t a = q % v; q /= v;
                const int64_t j = q % o;
                const int64_t i = q / o;
                D = eps_o[i >> sh] + eps_o[j >> sh] - eps_v[a >> sh] - eps_v[b >> sh];
            }
            T rn = rv[u] - Ec * tv[u];
            if (tfix != nullptr) rn = rn - E2c * tfix[idx];
            r[idx] = rn;
            t[idx] = tv[u] + scalar<T>::div_real(rn, D);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) symmetrize_kernel(const T *h, T *out, int64_t o, int64_t v) {
    const int64_t total = o * o * v * v;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t q = idx;
        const int64_t b = q % v; q /= v;
        const int64_t a = q % v; q /= v;
        const int64_t j = q % o;
        const int64_t i = q / o;
        out[idx] = h[idx] + h[((j * o + i) * v + b) * v + a];
    }
}

// --------------------------------------------------------------------------------------
// round-2 additions: batched streaming helpers
// --------------------------------------------------------------------------------------
// float64 -> complex128 (imaginary part 0): the real AO integrals of a magnetic-field point are uploaded once as
// float64 (half the PCIe bytes) and widened on the device for the complex128 contraction kernels.
__global__ void __launch_bounds__(kThreads) widen_kernel(cplx *__restrict__ dst, const double *__restrict__ src, int64_t len) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < len; base += kUnroll * stride) {
        double v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (base + u * stride < len) v[u] = src[base + u * stride];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (base + u * stride < len) dst[base + u * stride] = make_cplx(v[u], 0.0);
    }
}

// nb rows of `len` elements, row pitches dst_stride / src_stride (elements): dst[s][:] = alpha * src[s][:]
template <typename T>
__global__ void __launch_bounds__(kThreads)
copy_rows_kernel(T *__restrict__ dst, int64_t dst_stride, const T *__restrict__ src, int64_t src_stride, int64_t len, double alpha,
                 const int32_t *__restrict__ active) {
    const int s = blockIdx.y;
    if (active != nullptr && !active[s]) return;
    T *d = dst + (int64_t)s * dst_stride;
    const T *x = src + (int64_t)s * src_stride;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < len; base += kUnroll * stride) {
        T v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (base + u * stride < len) v[u] = x[base + u * stride];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (base + u * stride < len) d[base + u * stride] = alpha * v[u];
    }
}

// out[s][i,j,a,b] = h[s][i,j,a,b] + h[s][j,i,b,a] for nb points; one CTA per (point, i <= j pair): the (a,b) block
// and its transpose partner are staged through shared memory, so both global reads and both writes are coalesced.
template <typename T>
__global__ void __launch_bounds__(kThreads)
symmetrize_batch_kernel(const T *__restrict__ h, int64_t h_stride, T *__restrict__ out, int64_t out_stride, int o, int v,
                        const int32_t *__restrict__ active) {
    const int s = blockIdx.y;
    if (active != nullptr && !active[s]) return;
    // blockIdx.x enumerates pairs i <= j
    int p = blockIdx.x, i = 0;
    while (p >= o - i) { p -= o - i; ++i; }
    const int j = i + p;
    const T *hs = h + (int64_t)s * h_stride;
    T *os = out + (int64_t)s * out_stride;
    const int64_t vv = (int64_t)v * v;
    const T *hij = hs + ((int64_t)i * o + j) * vv;
    const T *hji = hs + ((int64_t)j * o + i) * vv;
    T *oij = os + ((int64_t)i * o + j) * vv;
    T *oji = os + ((int64_t)j * o + i) * vv;
    constexpr int TS = 32;
    __shared__ T tile_ij[TS][TS + 1];
    __shared__ T tile_ji[TS][TS + 1];
    const int tx = threadIdx.x % TS, ty = threadIdx.x / TS;          // 32 x 8
    for (int a0 = 0; a0 < v; a0 += TS)
        for (int b0 = 0; b0 < v; b0 += TS) {
            // tile (a0.., b0..) of h_ij and tile (b0.., a0..) of h_ji
            for (int r = ty; r < TS; r += kThreads / TS) {
                const int a = a0 + r, b = b0 + tx;
                tile_ij[r][tx] = (a < v && b < v) ? hij[(int64_t)a * v + b] : scalar<T>::zero();
                const int bb = b0 + r, aa = a0 + tx;
                tile_ji[r][tx] = (bb < v && aa < v) ? hji[(int64_t)bb * v + aa] : scalar<T>::zero();
            }
            __syncthreads();
            for (int r = ty; r < TS; r += kThreads / TS) {
                const int a = a0 + r, b = b0 + tx;
                if (a < v && b < v) oij[(int64_t)a * v + b] = tile_ij[r][tx] + tile_ji[tx][r];
                const int bb = b0 + r, aa = a0 + tx;
                if (i != j && bb < v && aa < v) oji[(int64_t)bb * v + aa] = tile_ji[r][tx] + tile_ij[tx][r];
            }
            __syncthreads();
        }
}

// pair packing for P-symmetric contractions (ladder <ab|cd> t_ijcd, ci_wfn.py:476): tp[s][p][:] = w_p * t[s][i_p, j_p][:]
// over the pairs i <= j (w = 1 for i < j, 1/2 for i == j), and the scatter-add back: h[s][i_p, j_p][:] += hp[s][p][:].
// With r2 = h + P h (symmetrize) the i > j blocks and the other half of the diagonal blocks come from the partner.
template <typename T>
__global__ void __launch_bounds__(kThreads)
pack_pairs_kernel(const T *__restrict__ t, int64_t t_stride, T *__restrict__ tp, int64_t tp_stride, int o, int64_t vv,
                  const int32_t *__restrict__ active) {
    const int s = blockIdx.z;
    if (active != nullptr && !active[s]) return;
    int p = blockIdx.y, i = 0;
    while (p >= o - i) { p -= o - i; ++i; }
    const int j = i + p;
    const double w = (i == j) ? 0.5 : 1.0;
    const T *src = t + (int64_t)s * t_stride + ((int64_t)i * o + j) * vv;
    T *dst = tp + (int64_t)s * tp_stride + (int64_t)blockIdx.y * vv;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < vv; e += (int64_t)gridDim.x * blockDim.x) dst[e] = w * src[e];
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
unpack_pairs_add_kernel(const T *__restrict__ hp, int64_t hp_stride, T *__restrict__ h, int64_t h_stride, int o, int64_t vv,
                        const int32_t *__restrict__ active) {
    const int s = blockIdx.z;
    if (active != nullptr && !active[s]) return;
    int p = blockIdx.y, i = 0;
    while (p >= o - i) { p -= o - i; ++i; }
    const int j = i + p;
    const T *src = hp + (int64_t)s * hp_stride + (int64_t)blockIdx.y * vv;
    T *dst = h + (int64_t)s * h_stride + ((int64_t)i * o + j) * vv;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < vv; e += (int64_t)gridDim.x * blockDim.x) dst[e] = dst[e] + src[e];
}

// --------------------------------------------------------------------------------------
// dots: out[j] = sum_i op(x_j[i]) * y[i]
// --------------------------------------------------------------------------------------
constexpr int kMaxVec = 8;

// single-vector dot (energies, norms): 2 accumulators, 8 independent element pairs in flight
template <typename T>
__global__ void __launch_bounds__(kThreads)
dot1_kernel(const T *__restrict__ x, const T *__restrict__ y, int64_t len, int conj_x, double *out, double *partials) {
    double acc[2] = {0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    constexpr int U = 8;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < len; base += U * stride) {
        T xv[U], yv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * stride;
            xv[u] = (i < len) ? x[i] : scalar<T>::zero();
            yv[u] = (i < len) ? y[i] : scalar<T>::zero();
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const T p = (conj_x ? scalar<T>::cj(xv[u]) : xv[u]) * yv[u];
            acc[0] += scalar<T>::re(p);
            acc[1] += scalar<T>::im(p);
        }
    }
    grid_sum_finish<2, kThreads>(acc, partials, [&](double(&tot)[2]) {
        out[0] = tot[0];
        out[1] = tot[1];
    });
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
dots_kernel(const T *__restrict__ x, int64_t xs, int nvec, const T *__restrict__ y, int64_t len, int conj_x,
            double *out, double *partials) {
    double acc[2 * kMaxVec];
#pragma unroll
    for (int k = 0; k < 2 * kMaxVec; ++k) acc[k] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const T yv = y[i];
        T xv[kMaxVec];
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) xv[j] = (j < nvec) ? x[(int64_t)j * xs + i] : scalar<T>::zero();
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
            const T p = (conj_x ? scalar<T>::cj(xv[j]) : xv[j]) * yv;
            acc[2 * j] += scalar<T>::re(p);
            acc[2 * j + 1] += scalar<T>::im(p);
        }
    }
    grid_sum_finish<2 * kMaxVec, kThreads>(acc, partials, [&](double(&tot)[2 * kMaxVec]) {
        for (int j = 0; j < nvec; ++j) {
            out[2 * j] = tot[2 * j];
            out[2 * j + 1] = tot[2 * j + 1];
        }
    });
}

// DIIS push: copy (r, t) into ring slot (it-1)%8 and refresh row/column `slot` of the Gram
// matrix B[m][n] = sum conj(e_m) e_n (utils.py:117-121; the reference rebuilds all of B every
// iteration, only one row/column actually changes).
template <typename T>
__global__ void __launch_bounds__(kThreads)
diis_push_kernel(const T *__restrict__ r, const T *__restrict__ t, T *__restrict__ hist_e, T *__restrict__ hist_t,
                 int64_t len, const int *iter, double *B, double *partials, const int32_t *__restrict__ active) {
    const int z = blockIdx.y;
    if (active != nullptr && active[z] == 0) return;
    r += z * len; t += z * len; hist_e += (int64_t)z * kMaxVec * len; hist_t += (int64_t)z * kMaxVec * len;
    B += z * 2 * kMaxVec * kMaxVec; partials += (int64_t)z * (kReduceMaxBlocks * kReduceMaxVals + 2);
    const int it = *iter;
    const int slot = (it - 1) % kMaxVec;
    const int m = it < kMaxVec ? it : kMaxVec;
    double acc[2 * kMaxVec];
#pragma unroll
    for (int k = 0; k < 2 * kMaxVec; ++k) acc[k] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    constexpr int U = 2;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < len; base += U * stride) {
        T rv[U], tv[U], ev[U][kMaxVec];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * stride;
            const bool ok = i < len;
            rv[u] = ok ? r[i] : scalar<T>::zero();
            tv[u] = ok ? t[i] : scalar<T>::zero();
#pragma unroll
            for (int j = 0; j < kMaxVec; ++j)
                ev[u][j] = (ok && j < m && j != slot) ? hist_e[(int64_t)j * len + i] : scalar<T>::zero();
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * stride;
            if (i < len) {
                hist_e[(int64_t)slot * len + i] = rv[u];
                hist_t[(int64_t)slot * len + i] = tv[u];
            }
#pragma unroll
            for (int j = 0; j < kMaxVec; ++j) {
                const T e = (j == slot) ? rv[u] : ev[u][j];
                const T p = scalar<T>::cj(e) * rv[u];     // <e_j | e_slot>
                acc[2 * j] += scalar<T>::re(p);
                acc[2 * j + 1] += scalar<T>::im(p);
            }
        }
    }
    grid_sum_finish<2 * kMaxVec, kThreads>(acc, partials, [&](double(&tot)[2 * kMaxVec]) {
        // B stored as complex (re,im) 8x8 row-major regardless of dtype
        for (int j = 0; j < m; ++j) {
            B[2 * (j * kMaxVec + slot)] = tot[2 * j];
            B[2 * (j * kMaxVec + slot) + 1] = tot[2 * j + 1];
            B[2 * (slot * kMaxVec + j)] = tot[2 * j];
            B[2 * (slot * kMaxVec + j) + 1] = -tot[2 * j + 1];
        }
    });
}

// Bordered system  [B -1; -1 0] c = (0,..,0,-1)  with partial pivoting (np.linalg.solve,
// utils.py:122-135).  m <= 8 -> at most 9x9; one thread is plenty.
__global__ void diis_solve_kernel(const double *B, int ldb, const int *iter, int m_fixed, double *c,
                                  const int32_t *__restrict__ active) {
    if (threadIdx.x != 0) return;
    const int z = blockIdx.x;
    if (active != nullptr && active[z] == 0) return;
    B += z * 2 * kMaxVec * kMaxVec; c += z * 2 * kMaxVec;
    int m = m_fixed;
    if (iter != nullptr) m = (*iter < kMaxVec) ? *iter : kMaxVec;
    const int n = m + 1;
    cplx Amat[(kMaxVec + 1) * (kMaxVec + 2)];
    const int ld = kMaxVec + 2;
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {
            cplx v;
            if (i < m && j < m) v = make_cplx(B[2 * (i * ldb + j)], B[2 * (i * ldb + j) + 1]);
            else if (i == m && j == m) v = make_cplx(0.0, 0.0);
            else v = make_cplx(-1.0, 0.0);
            Amat[i * ld + j] = v;
        }
        Amat[i * ld + n] = make_cplx(i == m ? -1.0 : 0.0, 0.0);
    }
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double best = abs2(Amat[k * ld + k]);
        for (int i = k + 1; i < n; ++i) {
            double a = abs2(Amat[i * ld + k]);
            if (a > best) { best = a; piv = i; }
        }
        if (piv != k)
            for (int j = k; j <= n; ++j) {
                cplx tmp = Amat[k * ld + j];
                Amat[k * ld + j] = Amat[piv * ld + j];
                Amat[piv * ld + j] = tmp;
            }
        const cplx pv = Amat[k * ld + k];
        for (int i = k + 1; i < n; ++i) {
            const cplx f = cdiv(Amat[i * ld + k], pv);
            for (int j = k + 1; j <= n; ++j) Amat[i * ld + j] = Amat[i * ld + j] - f * Amat[k * ld + j];
        }
    }
    cplx x[kMaxVec + 1];
    for (int i = n - 1; i >= 0; --i) {
        cplx s = Amat[i * ld + n];
        for (int j = i + 1; j < n; ++j) s = s - Amat[i * ld + j] * x[j];
        x[i] = cdiv(s, Amat[i * ld + i]);
    }
    for (int j = 0; j < m; ++j) {
        c[2 * j] = x[j].x;
        c[2 * j + 1] = x[j].y;
    }
}

// t <- sum_j c_j T_j (m from *iter or fixed; m == 0 keeps t), E = sum w t, rms1/rms2.
template <typename T>
__global__ void __launch_bounds__(kThreads)
lincomb_kernel(const T *hist, int64_t hs, const int *iter, int m_fixed, const double *c, T *t, const T *t_old,
               const T *w, int64_t n1, int64_t len, double *out, double *partials, const int32_t *__restrict__ active) {
    const int z = blockIdx.y;
    if (active != nullptr && active[z] == 0) return;
    hist += (int64_t)z * kMaxVec * hs; c += z * 2 * kMaxVec; t += z * len; t_old += z * len; w += z * len;
    out += z * 6; partials += (int64_t)z * (kReduceMaxBlocks * kReduceMaxVals + 2);
    int m = m_fixed;
    if (iter != nullptr) m = (*iter < kMaxVec) ? *iter : kMaxVec;
    T cj[kMaxVec];
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) cj[j] = (j < m) ? scalar<T>::make(c[2 * j], c[2 * j + 1]) : scalar<T>::zero();
    double acc[6] = {0, 0, 0, 0, 0, 0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    constexpr int U = 2;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < len; base += U * stride) {
        T hv[U][kMaxVec], wv[U], ov[U], tcur[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * stride;
            const bool ok = i < len;
#pragma unroll
            for (int j = 0; j < kMaxVec; ++j) hv[u][j] = (ok && j < m) ? hist[(int64_t)j * hs + i] : scalar<T>::zero();
            wv[u] = ok ? w[i] : scalar<T>::zero();
            ov[u] = ok ? t_old[i] : scalar<T>::zero();
            tcur[u] = (ok && m == 0) ? t[i] : scalar<T>::zero();
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * stride;
            if (i >= len) continue;
            T tv;
            if (m > 0) {
                tv = scalar<T>::zero();
#pragma unroll
                for (int j = 0; j < kMaxVec; ++j) tv = tv + cj[j] * hv[u][j];
                t[i] = tv;
            } else {
                tv = tcur[u];
            }
            const T e = wv[u] * tv;
            acc[0] += scalar<T>::re(e);
            acc[1] += scalar<T>::im(e);
            const T d = ov[u] - tv;
            const T d2 = d * d;
            const int k = (i < n1) ? 2 : 4;
            acc[k] += scalar<T>::re(d2);
            acc[k + 1] += scalar<T>::im(d2);
        }
    }
    grid_sum_finish<6, kThreads>(acc, partials, [&](double(&tot)[6]) {
#pragma unroll
        for (int k = 0; k < 6; ++k) out[k] = tot[k];
    });
}

__global__ void iter_advance_kernel(int *iter) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *iter += 1;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) copy_kernel(T *__restrict__ dst, const T *__restrict__ src, int64_t len) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < len; base += kUnroll * stride) {
        T v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (base + u * stride < len) v[u] = src[base + u * stride];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (base + u * stride < len) dst[base + u * stride] = v[u];
    }
}

// y <- alpha * op(x) + beta * y   (beta == 0: y is write-only)
template <typename T>
__global__ void __launch_bounds__(kThreads)
axpby_kernel(int64_t len, T alpha, const T *x, int conj_x, T beta, int has_beta, T *y) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        T xv = x[i];
        if (conj_x) xv = scalar<T>::cj(xv);
        T v = alpha * xv;
        if (has_beta) v = v + beta * y[i];
        y[i] = v;
    }
}

static int stream_grid(int64_t n) {
    int64_t b = (n + kThreads - 1) / kThreads;
    const int64_t cap = 148 * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace apyib

using namespace apyib;

extern "C" int64_t apyib_reduce_scratch_len(void) { return (int64_t)kReduceMaxBlocks * kReduceMaxVals + 2; }

static int gather4_impl(int dtype, const void *d_src, const int64_t src_dims[4], int spin, void *d_out,
                        const int64_t out_dims[4], const int32_t perm1[4], const int64_t start1[4], double c1,
                        const int32_t perm2[4], const int64_t start2[4], double c2, void *stream, int nb,
                        int64_t src_bstride) {
    APYIB_REQUIRE(nb >= 1 && nb <= 65535, "batch");
    APYIB_REQUIRE(dtype == APYIB_F64 || dtype == APYIB_C128, "dtype");
    APYIB_REQUIRE(d_src && d_out, "null pointer");
    Gather4Args g;
    int64_t total = 1;
    for (int k = 0; k < 4; ++k) {
        g.sd[k] = src_dims[k];
        g.od[k] = out_dims[k];
        g.perm1[k] = perm1[k];
        g.st1[k] = start1[k];
        g.perm2[k] = perm2 ? perm2[k] : perm1[k];
        g.st2[k] = start2 ? start2[k] : start1[k];
        APYIB_REQUIRE(perm1[k] >= 0 && perm1[k] < 4, "perm1");
        APYIB_REQUIRE(g.perm2[k] >= 0 && g.perm2[k] < 4, "perm2");
        total *= out_dims[k];
    }
    // bounds: every source index must stay inside the (spin-expanded) source
    for (int k = 0; k < 4; ++k) {
        const int64_t lim = spin ? 2 * src_dims[k] : src_dims[k];
        APYIB_REQUIRE(start1[k] >= 0 && start1[k] + out_dims[perm1[k]] <= lim, "block 1 out of range");
        if (c2 != 0.0) APYIB_REQUIRE(g.st2[k] >= 0 && g.st2[k] + out_dims[g.perm2[k]] <= lim, "block 2 out of range");
    }
    g.c1 = c1; g.c2 = c2; g.spin = spin;
    if (total == 0) return APYIB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    APYIB_REQUIRE(out_dims[2] * out_dims[3] < 2147483647LL, "inner block too large");
    // leading-index groups: (x0, x1) if that already gives >= 2 CTAs per SM, else (x0, x1, x2)
    const int lead = (out_dims[0] * out_dims[1] >= 2 * 148 || out_dims[2] == 1) ? 2 : 3;
    const int64_t nouter = lead == 3 ? out_dims[0] * out_dims[1] * out_dims[2] : out_dims[0] * out_dims[1];
    const int64_t cap = nb > 1 ? (148 * 16 + nb - 1) / nb : 148 * 16;
    const dim3 grid((unsigned)(nouter < cap ? nouter : cap), (unsigned)nb);
    if (dtype == APYIB_C128)
        gather4_kernel<cplx><<<grid, kThreads, 0, st>>>((const cplx *)d_src, (cplx *)d_out, g, lead, src_bstride, total);
    else
        gather4_kernel<double><<<grid, kThreads, 0, st>>>((const double *)d_src, (double *)d_out, g, lead, src_bstride, total);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_gather4(int dtype, const void *d_src, const int64_t src_dims[4], int spin, void *d_out,
                             const int64_t out_dims[4], const int32_t perm1[4], const int64_t start1[4], double c1,
                             const int32_t perm2[4], const int64_t start2[4], double c2, void *stream) {
    return gather4_impl(dtype, d_src, src_dims, spin, d_out, out_dims, perm1, start1, c1, perm2, start2, c2, stream, 1, 0);
}

// the same block of nb tensors that sit src_bstride elements apart (the MO integrals of a stack of
// finite-difference points), out[nb][...] contiguous: one launch, grid.y = point
extern "C" int apyib_gather4_batch(int dtype, const void *d_src, const int64_t src_dims[4], int spin, int nb,
                                   int64_t src_bstride, void *d_out, const int64_t out_dims[4], const int32_t perm1[4],
                                   const int64_t start1[4], double c1, const int32_t perm2[4], const int64_t start2[4],
                                   double c2, void *stream) {
    return gather4_impl(dtype, d_src, src_dims, spin, d_out, out_dims, perm1, start1, c1, perm2, start2, c2, stream, nb,
                        src_bstride);
}

extern "C" int apyib_gather2(int dtype, const void *d_src, const int64_t src_dims[2], int spin, void *d_out,
                             const int64_t out_dims[2], const int32_t perm[2], const int64_t start[2], void *stream) {
    APYIB_REQUIRE(dtype == APYIB_F64 || dtype == APYIB_C128, "dtype");
    APYIB_REQUIRE(d_src && d_out, "null pointer");
    for (int k = 0; k < 2; ++k) {
        const int64_t lim = spin ? 2 * src_dims[k] : src_dims[k];
        APYIB_REQUIRE(perm[k] == 0 || perm[k] == 1, "perm");
        APYIB_REQUIRE(start[k] >= 0 && start[k] + out_dims[perm[k]] <= lim, "block out of range");
    }
    const int64_t total = out_dims[0] * out_dims[1];
    if (total == 0) return APYIB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == APYIB_C128)
        gather2_kernel<cplx><<<stream_grid(total), kThreads, 0, st>>>((const cplx *)d_src, (cplx *)d_out, src_dims[0],
                                                                     src_dims[1], out_dims[0], out_dims[1], perm[0],
                                                                     perm[1], start[0], start[1], spin);
    else
        gather2_kernel<double><<<stream_grid(total), kThreads, 0, st>>>((const double *)d_src, (double *)d_out,
                                                                       src_dims[0], src_dims[1], out_dims[0],
                                                                       out_dims[1], perm[0], perm[1], start[0],
                                                                       start[1], spin);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_mp2_t2_energy(int dtype, const void *d_eri_mo, int64_t n, int64_t o, const double *d_eps,
                                   int spin_orbital, void *d_t2, double *d_E, double *d_partials, void *d_work,
                                   void *stream) {
    APYIB_REQUIRE(dtype == APYIB_F64 || dtype == APYIB_C128, "dtype");
    APYIB_REQUIRE(d_eri_mo && d_eps && d_t2 && d_E && d_partials, "null pointer");
    APYIB_REQUIRE(n > 0 && o >= 0 && o <= n, "sizes");
    APYIB_REQUIRE(spin_orbital >= 0 && spin_orbital <= 3, "flags: bit 0 = spin-orbital, bit 1 = Hermitian integrals");
    const bool hermitian = (spin_orbital & 2) != 0;
    spin_orbital &= 1;
    APYIB_REQUIRE(!spin_orbital || d_work, "spin-orbital MP2 needs a workspace of o*o*v*v elements");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t v = n - o;
    if (o == 0 || v == 0) {
        APYIB_CUDA_CHECK(cudaMemsetAsync(d_E, 0, 2 * sizeof(double), st));
        return APYIB_OK;
    }
    APYIB_REQUIRE(4 * v * v < 2147483647LL && 4 * o * o < 2147483647LL, "index range");
    const size_t es = dtype == APYIB_C128 ? 16 : 8;
    const int ld = (int)o | 1;
    const size_t budget = 24 * 1024;      // small slabs -> many CTAs per SM -> enough loads in flight
    int bt = (int)(budget / (2 * (size_t)ld * es));
    if (bt > v) bt = (int)v;
    if (bt < 1) bt = 1;
    const size_t smem = 2 * (size_t)bt * ld * es;
    APYIB_REQUIRE(smem <= 200 * 1024, "too many occupied orbitals for the shared-memory slab");
    const int64_t work = o * v;
    int grid = (int)(work < kReduceMaxBlocks ? work : kReduceMaxBlocks);
    void *t_spatial = spin_orbital ? d_work : d_t2;
#define MP2_LAUNCH2(T, SOE, HM)                                                                                   \
    do {                                                                                                          \
        APYIB_CUDA_CHECK(cudaFuncSetAttribute(mp2_spatial_kernel<T, SOE, HM>,                                      \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));          \
        mp2_spatial_kernel<T, SOE, HM><<<grid, kMp2Threads, smem, st>>>((const T *)d_eri_mo, n, o, d_eps,          \
                                                                        (T *)t_spatial, d_E, d_partials, bt);     \
    } while (0)
#define MP2_LAUNCH(T, SOE)                                                                                        \
    do {                                                                                                          \
        if (hermitian) MP2_LAUNCH2(T, SOE, true); else MP2_LAUNCH2(T, SOE, false);                                \
    } while (0)
    if (dtype == APYIB_C128) {
        if (spin_orbital) MP2_LAUNCH(cplx, true); else MP2_LAUNCH(cplx, false);
    } else {
        if (spin_orbital) MP2_LAUNCH(double, true); else MP2_LAUNCH(double, false);
    }
#undef MP2_LAUNCH
#undef MP2_LAUNCH2
    APYIB_LAUNCH_CHECK();
    if (spin_orbital) {
        const int64_t pairs = 4 * o * o;
        const int g2 = (int)(pairs < 148 * 16 ? pairs : 148 * 16);
        if (dtype == APYIB_C128)
            mp2_so_expand_kernel<cplx><<<g2, kMp2Threads, 0, st>>>((const cplx *)d_work, (int)o, (int)v, (cplx *)d_t2);
        else
            mp2_so_expand_kernel<double><<<g2, kMp2Threads, 0, st>>>((const double *)d_work, (int)o, (int)v, (double *)d_t2);
        APYIB_LAUNCH_CHECK();
    }
    return APYIB_OK;
}

extern "C" int apyib_ci_update(int dtype, void *d_r, void *d_t, const double *d_E, const double *d_eps_o,
                               const double *d_eps_v, int64_t o, int64_t v, int has_singles, int spin_orbital,
                               int nb, const int32_t *d_active, const double *d_E2, const double *d_E2_offset,
                               const void *d_t_fixed, void *stream) {
    APYIB_REQUIRE(d_t_fixed == nullptr || (d_E2 && d_E2_offset), "linear-response form needs E2 and its offset");
    APYIB_REQUIRE(nb >= 1 && nb <= 65535, "batch");
    APYIB_REQUIRE(dtype == APYIB_F64 || dtype == APYIB_C128, "dtype");
    APYIB_REQUIRE(d_r && d_t && d_E && d_eps_o && d_eps_v, "null pointer");
    const int64_t n1 = has_singles ? o * v : 0, n2 = o * o * v * v;
    if (n1 + n2 == 0) return APYIB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(stream_grid(n1 + n2), nb);
    const int sh = spin_orbital ? 1 : 0;
    if (dtype == APYIB_C128)
        ci_update_kernel<cplx><<<grid, kThreads, 0, st>>>((cplx *)d_r, (cplx *)d_t, d_E, d_eps_o, d_eps_v, o, v, n1, n2, sh, d_active, d_E2, d_E2_offset, (const cplx *)d_t_fixed);
    else
        ci_update_kernel<double><<<grid, kThreads, 0, st>>>((double *)d_r, (double *)d_t, d_E, d_eps_o, d_eps_v, o, v, n1, n2, sh, d_active, d_E2, d_E2_offset, (const double *)d_t_fixed);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_symmetrize_ijab(int dtype, const void *d_half, void *d_out, int64_t o, int64_t v, void *stream) {
    APYIB_REQUIRE(dtype == APYIB_F64 || dtype == APYIB_C128, "dtype");
    APYIB_REQUIRE(d_half && d_out && d_half != d_out, "pointers (must be out of place)");
    const int64_t total = o * o * v * v;
    if (total == 0) return APYIB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == APYIB_C128)
        symmetrize_kernel<cplx><<<stream_grid(total), kThreads, 0, st>>>((const cplx *)d_half, (cplx *)d_out, o, v);
    else
        symmetrize_kernel<double><<<stream_grid(total), kThreads, 0, st>>>((const double *)d_half, (double *)d_out, o, v);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_dots(int dtype, const void *d_x, int64_t x_stride, int nvec, const void *d_y, int64_t len,
                          int conj_x, double *d_out, double *d_partials, void *stream) {
    APYIB_REQUIRE(dtype == APYIB_F64 || dtype == APYIB_C128, "dtype");
    APYIB_REQUIRE(d_x && d_y && d_out && d_partials, "null pointer");
    APYIB_REQUIRE(nvec >= 1 && nvec <= kMaxVec, "nvec");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = stream_grid(len);
    if (nvec == 1) {
        if (dtype == APYIB_C128)
            dot1_kernel<cplx><<<grid, kThreads, 0, st>>>((const cplx *)d_x, (const cplx *)d_y, len, conj_x, d_out, d_partials);
        else
            dot1_kernel<double><<<grid, kThreads, 0, st>>>((const double *)d_x, (const double *)d_y, len, conj_x, d_out, d_partials);
        APYIB_LAUNCH_CHECK();
        return APYIB_OK;
    }
    if (dtype == APYIB_C128)
        dots_kernel<cplx><<<grid, kThreads, 0, st>>>((const cplx *)d_x, x_stride, nvec, (const cplx *)d_y, len, conj_x, d_out, d_partials);
    else
        dots_kernel<double><<<grid, kThreads, 0, st>>>((const double *)d_x, x_stride, nvec, (const double *)d_y, len, conj_x, d_out, d_partials);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_diis_push(int dtype, const void *d_r, const void *d_t, void *d_hist_e, void *d_hist_t,
                               int64_t len, const int32_t *d_iter, double *d_B, double *d_partials, int nb,
                               const int32_t *d_active, void *stream) {
    APYIB_REQUIRE(nb >= 1 && nb <= 65535, "batch");
    APYIB_REQUIRE(dtype == APYIB_F64 || dtype == APYIB_C128, "dtype");
    APYIB_REQUIRE(d_r && d_t && d_hist_e && d_hist_t && d_iter && d_B && d_partials, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(stream_grid(len), nb);
    if (dtype == APYIB_C128)
        diis_push_kernel<cplx><<<grid, kThreads, 0, st>>>((const cplx *)d_r, (const cplx *)d_t, (cplx *)d_hist_e, (cplx *)d_hist_t, len, d_iter, d_B, d_partials, d_active);
    else
        diis_push_kernel<double><<<grid, kThreads, 0, st>>>((const double *)d_r, (const double *)d_t, (double *)d_hist_e, (double *)d_hist_t, len, d_iter, d_B, d_partials, d_active);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_diis_solve(int dtype, const double *d_B, int ldb, int m, const int32_t *d_iter, double *d_c,
                                int nb, const int32_t *d_active, void *stream) {
    (void)dtype;
    APYIB_REQUIRE(nb >= 1 && nb <= 65535 && (nb == 1 || ldb == kMaxVec), "batch (batched B must be 8x8 blocks)");
    APYIB_REQUIRE(d_B && d_c, "null pointer");
    APYIB_REQUIRE(d_iter != nullptr || (m >= 1 && m <= kMaxVec), "m");
    diis_solve_kernel<<<nb, 32, 0, (cudaStream_t)stream>>>(d_B, ldb, d_iter, m, d_c, d_active);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_lincomb_energy_rms(int dtype, const void *d_hist, int64_t hist_stride, int m,
                                        const int32_t *d_iter, const double *d_c, void *d_t, const void *d_t_old,
                                        const void *d_w, int64_t n1, int64_t len, double *d_out, double *d_partials,
                                        int nb, const int32_t *d_active, void *stream) {
    APYIB_REQUIRE(nb >= 1 && nb <= 65535, "batch");
    APYIB_REQUIRE(dtype == APYIB_F64 || dtype == APYIB_C128, "dtype");
    APYIB_REQUIRE(d_t && d_t_old && d_w && d_out && d_partials, "null pointer");
    APYIB_REQUIRE(m >= 0 && m <= kMaxVec, "m");
    APYIB_REQUIRE((m == 0 && d_iter == nullptr) || (d_hist && d_c), "history");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(stream_grid(len), nb);
    if (dtype == APYIB_C128)
        lincomb_kernel<cplx><<<grid, kThreads, 0, st>>>((const cplx *)d_hist, hist_stride, d_iter, m, d_c, (cplx *)d_t, (const cplx *)d_t_old, (const cplx *)d_w, n1, len, d_out, d_partials, d_active);
    else
        lincomb_kernel<double><<<grid, kThreads, 0, st>>>((const double *)d_hist, hist_stride, d_iter, m, d_c, (double *)d_t, (const double *)d_t_old, (const double *)d_w, n1, len, d_out, d_partials, d_active);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_iter_advance(int32_t *d_iter, void *stream) {
    APYIB_REQUIRE(d_iter, "null pointer");
    iter_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_iter);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_copy(int dtype, void *d_dst, const void *d_src, int64_t len, void *stream) {
    APYIB_REQUIRE(dtype == APYIB_F64 || dtype == APYIB_C128, "dtype");
    if (len == 0) return APYIB_OK;
    APYIB_REQUIRE(d_dst && d_src, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == APYIB_C128)
        copy_kernel<cplx><<<stream_grid(len), kThreads, 0, st>>>((cplx *)d_dst, (const cplx *)d_src, len);
    else
        copy_kernel<double><<<stream_grid(len), kThreads, 0, st>>>((double *)d_dst, (const double *)d_src, len);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_axpby(int dtype, int64_t len, double alpha_re, double alpha_im, const void *d_x, int conj_x,
                           double beta_re, double beta_im, void *d_y, void *stream) {
    APYIB_REQUIRE(dtype == APYIB_F64 || dtype == APYIB_C128, "dtype");
    APYIB_REQUIRE(d_x && d_y, "null pointer");
    if (len == 0) return APYIB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int hb = (beta_re != 0.0 || beta_im != 0.0) ? 1 : 0;
    if (dtype == APYIB_C128)
        axpby_kernel<cplx><<<stream_grid(len), kThreads, 0, st>>>(len, make_cplx(alpha_re, alpha_im), (const cplx *)d_x,
                                                                 conj_x, make_cplx(beta_re, beta_im), hb, (cplx *)d_y);
    else
        axpby_kernel<double><<<stream_grid(len), kThreads, 0, st>>>(len, alpha_re, (const double *)d_x, 0, beta_re, hb,
                                                                   (double *)d_y);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_widen(void *d_dst, const void *d_src, int64_t len, void *stream) {
    if (len == 0) return APYIB_OK;
    APYIB_REQUIRE(d_dst && d_src, "null pointer");
    widen_kernel<<<stream_grid(len), kThreads, 0, (cudaStream_t)stream>>>((cplx *)d_dst, (const double *)d_src, len);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_copy_rows(int dtype, void *d_dst, int64_t dst_stride, const void *d_src, int64_t src_stride, int64_t len,
                               int nb, double alpha, const int32_t *d_active, void *stream) {
    APYIB_REQUIRE(dtype == APYIB_F64 || dtype == APYIB_C128, "dtype");
    if (len == 0 || nb == 0) return APYIB_OK;
    APYIB_REQUIRE(d_dst && d_src && nb > 0 && nb <= 65535, "arguments");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(std::max(1, std::min(stream_grid(len), std::max(1, 1184 / nb))), nb);
    if (dtype == APYIB_C128)
        copy_rows_kernel<cplx><<<grid, kThreads, 0, st>>>((cplx *)d_dst, dst_stride, (const cplx *)d_src, src_stride, len, alpha, d_active);
    else
        copy_rows_kernel<double><<<grid, kThreads, 0, st>>>((double *)d_dst, dst_stride, (const double *)d_src, src_stride, len, alpha, d_active);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_symmetrize_ijab_batch(int dtype, const void *d_half, int64_t half_stride, void *d_out, int64_t out_stride,
                                           int64_t o, int64_t v, int nb, const int32_t *d_active, void *stream) {
    APYIB_REQUIRE(dtype == APYIB_F64 || dtype == APYIB_C128, "dtype");
    APYIB_REQUIRE(d_half && d_out && d_half != d_out, "pointers (must be out of place)");
    if (o == 0 || v == 0 || nb == 0) return APYIB_OK;
    APYIB_REQUIRE(nb > 0 && nb <= 65535 && o < 46340, "batch / size");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)(o * (o + 1) / 2), nb);
    if (dtype == APYIB_C128)
        symmetrize_batch_kernel<cplx><<<grid, kThreads, 0, st>>>((const cplx *)d_half, half_stride, (cplx *)d_out, out_stride, (int)o, (int)v, d_active);
    else
        symmetrize_batch_kernel<double><<<grid, kThreads, 0, st>>>((const double *)d_half, half_stride, (double *)d_out, out_stride, (int)o, (int)v, d_active);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_pack_pairs(int dtype, const void *d_t, int64_t t_stride, void *d_tp, int64_t tp_stride, int64_t o, int64_t vv,
                                int nb, const int32_t *d_active, void *stream) {
    APYIB_REQUIRE(dtype == APYIB_F64 || dtype == APYIB_C128, "dtype");
    if (o == 0 || vv == 0 || nb == 0) return APYIB_OK;
    APYIB_REQUIRE(d_t && d_tp && nb > 0 && nb <= 65535 && o * (o + 1) / 2 <= 65535, "arguments");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)std::min<int64_t>((vv + kThreads - 1) / kThreads, 8), (unsigned)(o * (o + 1) / 2), nb);
    if (dtype == APYIB_C128)
        pack_pairs_kernel<cplx><<<grid, kThreads, 0, st>>>((const cplx *)d_t, t_stride, (cplx *)d_tp, tp_stride, (int)o, vv, d_active);
    else
        pack_pairs_kernel<double><<<grid, kThreads, 0, st>>>((const double *)d_t, t_stride, (double *)d_tp, tp_stride, (int)o, vv, d_active);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}

extern "C" int apyib_unpack_pairs_add(int dtype, const void *d_hp, int64_t hp_stride, void *d_h, int64_t h_stride, int64_t o,
                                      int64_t vv, int nb, const int32_t *d_active, void *stream) {
    APYIB_REQUIRE(dtype == APYIB_F64 || dtype == APYIB_C128, "dtype");
    if (o == 0 || vv == 0 || nb == 0) return APYIB_OK;
    APYIB_REQUIRE(d_hp && d_h && nb > 0 && nb <= 65535 && o * (o + 1) / 2 <= 65535, "arguments");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)std::min<int64_t>((vv + kThreads - 1) / kThreads, 8), (unsigned)(o * (o + 1) / 2), nb);
    if (dtype == APYIB_C128)
        unpack_pairs_add_kernel<cplx><<<grid, kThreads, 0, st>>>((const cplx *)d_hp, hp_stride, (cplx *)d_h, h_stride, (int)o, vv, d_active);
    else
        unpack_pairs_add_kernel<double><<<grid, kThreads, 0, st>>>((const double *)d_hp, hp_stride, (double *)d_h, h_stride, (int)o, vv, d_active);
    APYIB_LAUNCH_CHECK();
    return APYIB_OK;
}
