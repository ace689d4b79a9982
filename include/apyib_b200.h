/*
 * apyib_b200 -- C-ABI of the B200 (sm_100a) kernels behind apyib's correlated
 * wavefunction hot path.
 *
 * The reference (bshumberger/apyib) is pure Python: it has no FFI.  The
 * "interface each entry point replaces" is therefore the numpy / opt_einsum /
 * LAPACK call site inside the reference's Python classes; each prototype below
 * cites it (paths relative to the reference root).  INTEGRATION.md shows the
 * ctypes stubs a maintainer would add to those classes.
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer owned by the caller
 *     (PyTorch tensors on the host side); h_* are host pointers.
 *   - dtype: APYIB_F64 (double) or APYIB_C128 (interleaved re,im doubles,
 *     numpy complex128 / torch.complex128 layout).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - return value: 0 on success, negative error code otherwise;
 *     apyib_last_error() returns a static description for the calling thread.
 *   - no entry point allocates device memory, synchronises the device or
 *     throws; all work is enqueued on `stream`.
 */
#ifndef APYIB_B200_H
#define APYIB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APYIB_F64 0
#define APYIB_C128 1

#define APYIB_OK 0
#define APYIB_ERR_ARG -1
#define APYIB_ERR_CUDA -2
#define APYIB_ERR_UNSUPPORTED -3

/* ---- library ------------------------------------------------------------ */
int apyib_version(void);
const char *apyib_last_error(void);
/* sm count, compute capability major/minor, total global memory (bytes). */
int apyib_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, int64_t *mem_bytes);

/* ---- CUDA-graph capture of an iteration (replaces the Python `while` body of
 *      ci_wfn.py:78-130 / 205-257 / 302-378 / 451-527 being re-interpreted
 *      every trip) ------------------------------------------------------------ */
int apyib_graph_begin(void *stream);
int apyib_graph_end(void *stream, void **graph_exec_out);
int apyib_graph_launch(void *graph_exec, void *stream);
int apyib_graph_destroy(void *graph_exec);

/* ---- tensor contraction on the FP64 tensor path (DMMA) -------------------------
 * C[c_m[m] + c_n[n]] = alpha * sum_k opA(A[a_m[m] + a_k[k]]) * opB(B[b_k[k] + b_n[n]])
 *                      + beta * C[...]
 * Offsets are in ELEMENTS; the six tables are device int64 arrays of length
 * M, K, K, N, M, N.  This is one `opt_einsum.contract` call of the reference
 * (every oe.contract in ci_wfn.py:84-90, 211-218, 310-333, 458-482,
 * utils.py:240-252, 274-277, 381) with the index regrouping folded into the
 * tables instead of a transposed copy.  batch > 1 repeats the contraction for
 * `batch` problems whose bases differ by *_bstride elements (finite-difference
 * points solved together); d_active (nullable, int32[batch]) skips entries
 * whose flag is 0.  alpha/beta are given as (re, im); im ignored for F64.
 * conj_a / conj_b conjugate the operand (np.conjugate(C) in utils.py:252,275,277).
 * ksplit > 1 splits the K range over ksplit CTAs per output tile (dot-product-like contractions
 * with tiny M x N, e.g. the scalar energy/AAT sums); partial tiles go to d_work (batch * ksplit *
 * M * N elements) and are summed in fixed order by a second kernel -- deterministic.
 */
int apyib_contract(int dtype, const void *d_A, const void *d_B, void *d_C,
                   int64_t M, int64_t N, int64_t K,
                   const int64_t *d_a_m, const int64_t *d_a_k,
                   const int64_t *d_b_k, const int64_t *d_b_n,
                   const int64_t *d_c_m, const int64_t *d_c_n,
                   int a_kfast, int b_kfast, int conj_a, int conj_b,
                   double alpha_re, double alpha_im, double beta_re, double beta_im,
                   int batch, int64_t a_bstride, int64_t b_bstride, int64_t c_bstride,
                   const int32_t *d_active, int ksplit, void *d_work, void *stream);

/* TMA-fed variant (cp.async.bulk.tensor + mbarrier ring, 128B-swizzled tiles) for operands that
 * are k-contiguous strided matrices: opA(A[z*a_bstride + m*lda + k]), opB(B[z*b_bstride + n*ldb + k]);
 * C through offset tables as in apyib_contract.  Used for the ladder term <ab|cd> t_ijcd and the
 * first AO->MO quarter transform.  Returns APYIB_ERR_UNSUPPORTED (nothing launched) when the
 * operands do not meet the TMA alignment rules; the caller then uses apyib_contract.        */
int apyib_contract_tma(int dtype, const void *d_A, const void *d_B, void *d_C,
                       int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                       const int64_t *d_c_m, const int64_t *d_c_n, int conj_a, int conj_b,
                       double alpha_re, double alpha_im, double beta_re, double beta_im,
                       int batch, int64_t a_bstride, int64_t b_bstride, int64_t c_bstride,
                       const int32_t *d_active, void *stream);

/* ---- 4-index block gather with optional on-the-fly spin blocking ----------------
 * out[x0,x1,x2,x3] = c1 * G(start1 + x[perm1]) + c2 * G(start2 + x[perm2])   (dense out)
 * where G is the dense chemists' tensor src[n0,n1,n2,n3] (spin==0), or its
 * spin-orbital expansion G(p,q,r,s) = src[p/2,q/2,r/2,s/2] * [p%2==q%2][r%2==s%2]
 * (spin==1; utils.py:317-365).  perm*[k] says which OUTPUT axis runs along source
 * axis k.  Replaces the slice/swapaxes/copy expressions
 * `ERI.swapaxes(1,2)[blk] - ERI.swapaxes(1,2).swapaxes(2,3)[blk]` and
 * `2*ERI[blk] - ERI.swapaxes(2,3)[blk]` (ci_wfn.py:64-70, 191-197, 210-218, 309-334,
 * 433-442, 457-483; mp2_wfn.py:51-57, 79-85).  c2 == 0 skips the second term.
 */
int apyib_gather4(int dtype, const void *d_src, const int64_t src_dims[4], int spin,
                  void *d_out, const int64_t out_dims[4],
                  const int32_t perm1[4], const int64_t start1[4], double c1,
                  const int32_t perm2[4], const int64_t start2[4], double c2,
                  void *stream);
/* The same block of nb source tensors src_bstride elements apart (the MO integrals of a stack of
 * finite-difference points: ci_wfn.py builds these slices once per point), d_out[nb][...] contiguous.       */
int apyib_gather4_batch(int dtype, const void *d_src, const int64_t src_dims[4], int spin, int nb,
                        int64_t src_bstride, void *d_out, const int64_t out_dims[4], const int32_t perm1[4],
                        const int64_t start1[4], double c1, const int32_t perm2[4], const int64_t start2[4],
                        double c2, void *stream);
/* 2-index analogue (utils.py:283-313 compute_F_SO, :393-422 compute_so_overlap):
 * out[x0,x1] = src-or-spin-blocked(start + x[perm]).                                */
int apyib_gather2(int dtype, const void *d_src, const int64_t src_dims[2], int spin,
                  void *d_out, const int64_t out_dims[2],
                  const int32_t perm[2], const int64_t start[2], void *stream);

/* ---- MP2 amplitudes + energy, one streaming pass (mp2_wfn.py:42-59, 64-87) ------
 * d_eri_mo: chemists' (pq|rs) over the n = nbf - nfzc active MOs, occupied first.
 * spatial:  t2[i,j,a,b] = (ai|bj)/D_ijab ,  E = sum (2(ia|jb) - (ib|ja)) t2
 * spin_orbital: same on the spin-blocked, antisymmetrised integrals,
 *           t2 = <ab||ij>/D , E = 1/4 sum <ij||ab> t2   (O = 2o, V = 2v)
 * spin_orbital is a flag word: bit 0 = spin-orbital equations, bit 1 = the caller guarantees
 * (pq|rs) = conj((qp|sr)) (any MO tensor transformed from real AO integrals, utils.py:274-277);
 * the energy then reuses the integrals already read for t2 (2 reads + 1 write per amplitude).
 * d_eps: n orbital energies (double).  d_E: 2 doubles (re, im), written.
 * d_work: spin_orbital only, scratch of o*o*v*v elements (the spatial amplitudes).
 * d_partials: zero-initialised scratch of apyib_reduce_scratch_len() doubles.         */
int apyib_mp2_t2_energy(int dtype, const void *d_eri_mo, int64_t n, int64_t o,
                        const double *d_eps, int spin_orbital,
                        void *d_t2, double *d_E, double *d_partials, void *d_work, void *stream);
int64_t apyib_reduce_scratch_len(void);

/* ---- CI amplitude update (ci_wfn.py:93-94, 220-221, 316+334+336-337, 464+483+485-486)
 *   r <- r - E*t ;  t <- t + r / D      for the concatenated vector [t1 | t2]
 * D built on the fly from eps: D_ia = e_i - e_a, D_ijab = e_i + e_j - e_a - e_b
 * (spin_orbital: eps index p/2).  n1 = o*v (0 for CID), n2 = o*o*v*v, in the
 * solver's own (O,V) sizes.  d_E holds (re,im) of the current energy.
 * symmetrize: CID spatial only, r <- r + r^T(ij)(ab) is applied first
 * (ci_wfn.py:92; out-of-place into d_r from d_r_half).
 * Batched solves: nb finite-difference points are advanced by one launch; every per-point
 * array is the single-point array with a leading [nb] dimension (vectors [nb][len], energies
 * [nb][6], eps_o [nb][o_spatial], eps_v [nb][v_spatial], DIIS history [nb][8][len], Gram matrices
 * [nb][8][8], coefficients [nb][8], reduction scratch [nb][apyib_reduce_scratch_len()]); d_active
 * (nullable, int32[nb]) freezes converged points exactly where the reference `break`s.
 * Linear-response form (perturbed-amplitude iterations, analytic_aats.py:788-836, 1040-1088):
 * with d_t_fixed != NULL the update is r <- r - E*t - (E2 + E2_offset)*t_fixed, where t is the
 * perturbed amplitude vector, E = E_CISD (fixed), t_fixed the unperturbed amplitudes and
 * E2 + E2_offset the running projected energy derivative ([nb][6] layout like d_E).             */
int apyib_ci_update(int dtype, void *d_r, void *d_t, const double *d_E,
                    const double *d_eps_o, const double *d_eps_v,
                    int64_t o, int64_t v, int has_singles, int spin_orbital,
                    int nb, const int32_t *d_active,
                    const double *d_E2, const double *d_E2_offset, const void *d_t_fixed, void *stream);
int apyib_symmetrize_ijab(int dtype, const void *d_half, void *d_out, int64_t o, int64_t v,
                          void *stream);

/* ---- fused reductions ---------------------------------------------------------
 * apyib_dots: out[j] = sum_i opx(x_j[i]) * y[i]  for j < nvec, x_j = d_x + j*x_stride
 * (conj_x: B_mn = conj(e_m).e_n of utils.py:121; energy / norm dots of
 * ci_wfn.py:110, 355, 504, aats.py:144-155, 659-669).  out is 2*nvec doubles.       */
int apyib_dots(int dtype, const void *d_x, int64_t x_stride, int nvec, const void *d_y,
               int64_t len, int conj_x, double *d_out, double *d_partials, void *stream);
/* ---- DIIS, device-resident state machine (utils.py:104-140; call sites ci_wfn.py:97-107,
 * 224-234, 340-352, 489-501).  The iteration counter lives on the device (d_iter, 1-based) so
 * that one captured CUDA graph replays every iteration: history length m = min(iter, 8)
 * (the reference truncates to 7 before appending -> at most 8 vectors), ring slot =
 * (iter-1) % 8.
 *
 * apyib_diis_push: hist_e[slot] <- r, hist_t[slot] <- t, and row/column `slot` of the Gram
 *   matrix B[m][n] = sum conj(e_m) e_n (8x8 complex (re,im) row-major, device) is refreshed
 *   -- the reference rebuilds all of B each iteration (utils.py:121), only this row changes.
 * apyib_diis_solve: bordered system [B -1; -1 0] c = (0..0,-1) by Gaussian elimination with
 *   partial pivoting (np.linalg.solve, utils.py:122-135).  m from *d_iter, or `m` if d_iter
 *   is NULL.  c: m (re,im) pairs.
 * apyib_lincomb_energy_rms: t <- sum_j c[j] T_j (utils.py:138) fused with the new energy
 *   E = sum w[i] t[i] (ci_wfn.py:110, 237, 355, 504; w = energy weights laid out like t) and
 *   the unconjugated rms sums of ci_wfn.py:116, 243, 361-364, 510-513:
 *   d_out = { E.re, E.im, S1.re, S1.im, S2.re, S2.im },  S1 = sum_{i<n1} (t_old-t)^2, S2 = rest.
 *   m == 0 with d_iter == NULL keeps t (DIIS disabled).                                       */
int apyib_diis_push(int dtype, const void *d_r, const void *d_t, void *d_hist_e, void *d_hist_t,
                    int64_t len, const int32_t *d_iter, double *d_B, double *d_partials,
                    int nb, const int32_t *d_active, void *stream);
int apyib_diis_solve(int dtype, const double *d_B, int ldb, int m, const int32_t *d_iter, double *d_c,
                     int nb, const int32_t *d_active, void *stream);
int apyib_lincomb_energy_rms(int dtype, const void *d_hist, int64_t hist_stride, int m,
                             const int32_t *d_iter, const double *d_c, void *d_t, const void *d_t_old,
                             const void *d_w, int64_t n1, int64_t len, double *d_out,
                             double *d_partials, int nb, const int32_t *d_active, void *stream);
int apyib_iter_advance(int32_t *d_iter, void *stream);
/* y <- alpha*op(x) + beta*y: the scaled amplitude combinations t2_dH = N_mp*T(B+) - N_mn*T(B-),
 * conj(N_np*T(R+) - N_nn*T(R-)) of aats.py:690-711 and the 2*F_ov energy weights (ci_wfn.py:504). */
int apyib_axpby(int dtype, int64_t len, double alpha_re, double alpha_im, const void *d_x, int conj_x,
                double beta_re, double beta_im, void *d_y, void *stream);
/* t_old = t.copy() (ci_wfn.py:80, 207, 304-305, 453-454)                                      */
int apyib_copy(int dtype, void *d_dst, const void *d_src, int64_t len, void *stream);
/* Batched row copy with scale: dst[s][0:len] = alpha * src[s][0:len], s < nb, row pitches in elements
 * (the per-point `r = K.copy()` / `0.5 * <ab|ij>` set-ups of ci_wfn.py:83, 466 for a stack of points).     */
int apyib_copy_rows(int dtype, void *d_dst, int64_t dst_stride, const void *d_src, int64_t src_stride,
                    int64_t len, int nb, double alpha, const int32_t *d_active, void *stream);
/* float64 -> complex128: the AO integrals of a magnetic-field point are real (hamiltonian.py:29-35; only V picks
 * up the imaginary field term, :57-70); numpy upcasts them implicitly inside utils.py:274, here it is explicit. */
int apyib_widen(void *d_dst_c128, const void *d_src_f64, int64_t len, void *stream);
/* r_T2 = r_T2 + r_T2.swapaxes(0,1).swapaxes(2,3) (ci_wfn.py:92) for nb points in one launch (out of place).  */
int apyib_symmetrize_ijab_batch(int dtype, const void *d_half, int64_t half_stride, void *d_out,
                                int64_t out_stride, int64_t o, int64_t v, int nb, const int32_t *d_active,
                                void *stream);
/* Pair packing for the (i<->j, a<->b)-symmetric ladder term <ab|cd> t_ijcd (ci_wfn.py:87, 476):
 * tp[s][p][:] = w_p t[s][i_p, j_p][:] over the o(o+1)/2 pairs i <= j (w = 1/2 on the diagonal), and the
 * scatter-add h[s][i_p, j_p][:] += hp[s][p][:]; together with the symmetrisation above the contraction then
 * runs over o(o+1)/2 instead of o^2 occupied pairs.  vv = elements per (i, j) block.                        */
int apyib_pack_pairs(int dtype, const void *d_t, int64_t t_stride, void *d_tp, int64_t tp_stride, int64_t o,
                     int64_t vv, int nb, const int32_t *d_active, void *stream);
int apyib_unpack_pairs_add(int dtype, const void *d_hp, int64_t hp_stride, void *d_h, int64_t h_stride,
                           int64_t o, int64_t vv, int nb, const int32_t *d_active, void *stream);

/* ---- determinants of substituted occupied-overlap matrices ----------------------
 * out[r*ncol + c] = det( S[rows[r, :], cols[c, :]] ),  n x n, LU with partial
 * pivoting, one thread (n <= 12) or one sub-warp per matrix (np.linalg.det in aats.py:128,
 * 578-618, 672-675).
 * d_S: (ns x ns) complex128 row-major overlap (device); rows/cols: int32 index
 * lists (nrow x n), (ncol x n) as produced by apyib_det_index_lists / apyib_so_index_lists. */
int apyib_det_outer(const void *d_S, int ns, int n,
                    const int32_t *d_rows, int64_t nrow,
                    const int32_t *d_cols, int64_t ncol,
                    void *d_out, void *stream);
/* Fused determinant-table x amplitude-vector product; the table (the reference's 8-index
 * tensor, aats.py:575) is never materialised:
 *   Z[iy*nrow + r] = sum_c det(S[rows[r],cols[c]]) * Y[iy*ncol + c]          (ny <= 4)
 * (the einsum lines of aats.py:718-737 etc. restricted to i<j,a<b / k<l,c<d, with the
 * antisymmetric completion of aats.py:620-630 folded into the amplitude vectors).
 * d_work: apyib_det_matvec_work_len(...) complex128 elements of scratch.  Deterministic.       */
int apyib_det_matvec(const void *d_S, int ns, int n,
                     const int32_t *d_rows, int64_t nrow,
                     const int32_t *d_cols, int64_t ncol,
                     const void *d_Y, int ny, void *d_Z, void *d_work, void *stream);
int64_t apyib_det_matvec_work_len(int64_t nrow, int64_t ncol, int ny, int n);
/* Factorisation reuse.  The thread-per-matrix kernel runs a left-looking LU, which touches column j only
 * after the columns before it: two consecutive column lists that start with the same columns share those
 * panels, their L factors, the pivoting and the partial determinant, and only the last panel is redone.
 * apyib_det_sort_lists (host) re-orders an enumeration for that: inside every list the substituted
 * entries (value >= n) go last, the lists are sorted; sign[c] = +-1 relates the determinants and index[c]
 * is the position in the original enumeration.  The *_sorted entry points take the re-ordered lists and
 * still index d_Y / d_out by the ORIGINAL enumeration, so results equal apyib_det_outer / apyib_det_matvec
 * (same partial-pivoting LU of a column-permuted matrix).  2 <= n <= 12 only (APYIB_ERR_UNSUPPORTED).   */
int apyib_det_sort_lists(int n, const int32_t *h_lists, int64_t count, int32_t *h_sorted, double *h_sign,
                         int32_t *h_index);
int apyib_det_outer_sorted(const void *d_S, int ns, int n, const int32_t *d_rows, int64_t nrow,
                           const int32_t *d_cols_sorted, const double *d_col_sign, const int32_t *d_col_index,
                           int64_t ncol, void *d_out, void *stream);
int apyib_det_matvec_sorted(const void *d_S, int ns, int n, const int32_t *d_rows, int64_t nrow,
                            const int32_t *d_cols_sorted, const double *d_col_sign, const int32_t *d_col_index,
                            int64_t ncol, const void *d_Y, int ny, void *d_Z, void *d_work, void *stream);
/* Prefix-shared LU (csrc/dets_pairs.cu): the re-ordered lists of a singly (k = 1) or doubly (k = 2)
 * substituted table come in GROUPS of `group_len` consecutive lists that share their first n-k columns
 * (the unsubstituted ones) and end in every candidate column (k = 1) / every pair c < d (k = 2,
 * lexicographic) of the nc columns d_cand[] (ascending) -- what apyib_det_sort_lists produces for the
 * enumeration of aats.py:581-618.  The prefix is factorised ONCE per (row list, group) with partial
 * pivoting; every candidate column then costs one k-vector of the Schur complement (n k complex MACs) and
 * every determinant of the group a k x k determinant of those vectors.  Same quantity as
 * apyib_det_matvec_sorted (Z[q,r] = sum_c det(r,c) Y[q,c], Y indexed by the ORIGINAL enumeration), ~12x
 * fewer flops at n = 9, nc = 13.  Replaces the same np.linalg.det calls (aats.py:587-618).  Returns
 * APYIB_ERR_UNSUPPORTED when (n, k) is not instantiated (2 <= n <= 12) or the footprint does not fit.
 * d_work: apyib_det_matvec_pairs_work_len(...) complex128 elements per overlap (partial sums of the group chunks +
 * a copy of Y in sorted-list order with the list signs folded in, made by a small gather launch, so that the
 * determinant loop reads one contiguous warp-uniform amplitude per determinant).                            */
int64_t apyib_det_matvec_pairs_work_len(int64_t nrow, int64_t ngroup, int ny, int n, int k, int ns, int nc);
int apyib_det_matvec_pairs(const void *d_S, int ns, int n, int k, const int32_t *d_rows, int64_t nrow,
                           const int32_t *d_cols_sorted, const double *d_col_sign, const int32_t *d_col_index,
                           int64_t ncol, int64_t group_len, const int32_t *d_cand, int nc, const void *d_Y, int ny,
                           void *d_Z, void *d_work, void *stream);
/* Kernel variant of the prefix-shared LU: 0 (default) = one kernel for 1..4 amplitude vectors (count predicated
 * at run time); 1 = additionally a single-vector specialisation for k = 2 (no predicated issue slots in the
 * pair loop).  Same results either way.                                                                   */
int apyib_det_set_pairs_variant(int which);
/* Stacks of overlaps: the same row / column lists applied to nS overlap matrices d_S[nS][ns][ns] in ONE launch
 * (grid.y = overlap) -- the 12 (beta x pp/pn/np/nn) overlaps of one nuclear coordinate, the 6N pu/nu overlaps, ...
 * (aats.py:714-1008 evaluates compute_all_dets overlap by overlap).  Outputs are [nS][...] contiguous; overlap s
 * reads d_Y + s*y_stride (0: one Y for all); d_work is nS x the single-overlap work length; d_col_sign /
 * d_col_index may both be NULL (plain lists, apyib_det_outer / apyib_det_matvec semantics).                  */
int apyib_det_outer_stack(const void *d_S, int nS, int ns, int n, const int32_t *d_rows, int64_t nrow,
                          const int32_t *d_cols, const double *d_col_sign, const int32_t *d_col_index, int64_t ncol,
                          void *d_out, void *stream);
int apyib_det_matvec_stack(const void *d_S, int nS, int ns, int n, const int32_t *d_rows, int64_t nrow,
                           const int32_t *d_cols, const double *d_col_sign, const int32_t *d_col_index, int64_t ncol,
                           const void *d_Y, int64_t y_stride, int ny, void *d_Z, void *d_work, void *stream);
int apyib_det_matvec_pairs_stack(const void *d_S, int nS, int ns, int n, int k, const int32_t *d_rows, int64_t nrow,
                                 const int32_t *d_cols_sorted, const double *d_col_sign, const int32_t *d_col_index,
                                 int64_t ncol, int64_t group_len, const int32_t *d_cand, int nc, const void *d_Y,
                                 int64_t y_stride, int ny, void *d_Z, void *d_work, void *stream);
/* Which LU kernel apyib_det_outer / apyib_det_matvec launch: 0 (default) = one thread per matrix,
 * column panels in registers + L in shared memory, for 2 <= n <= 12 and the sub-warp kernel above
 * that; 1 = the sub-warp (one lane per row) kernel for every n.  Same results either way.       */
int apyib_det_set_kernel(int which);

/* Restricted-pair packing of doubles amplitudes, complex128 only:
 *   out[q*P + r] = 2 (x_q[i,j,a,b] - x_q[i,j,b,a] - x_q[j,i,a,b] + x_q[j,i,b,a]),  r = (i,a,j,b)
 * for the P rows of the doubles enumeration (apyib_det_enumeration; i,j carry the frozen-core
 * offset nf).  This is `t - t.swapaxes(2,3)` (aats.py:732 etc.) summed over the four images
 * that the antisymmetric completion of the determinant tensors (aats.py:620-630) generates.   */
int apyib_pack_doubles(const void *d_x, int64_t x_stride, int nq, int o, int v, int nf,
                       const int32_t *d_doubles, int64_t P, void *d_out, void *stream);

/* ---- determinant-lemma path: the same determinants from A = S_oo,
 * P = S_vo A^-1, Q = A^-1 S_ov, R = S_vv - S_vo A^-1 S_ov:
 *   det = det(A) (-1)^c det[[P[a,i'], R[a,c']], [A^-1[k,i'], -Q[k,c']]]   ((r+c) x (r+c) <= 4 x 4)
 * apyib_lemma_prepare: for a stack of nS overlaps (ns x ns, occupied block no x no) writes, per
 *   overlap, apyib_lemma_prep_len() complex numbers: det(A) | A^-1 | P | Q | R.
 * apyib_lemma_outer / apyib_lemma_matvec: as apyib_det_outer / apyib_det_matvec, but rows and
 *   columns are given as substitution tuples (rk, ck in {0,1,2} pairs (occupied, virtual) per
 *   entry, the enumeration of apyib_det_enumeration) and all nS overlaps are processed in one
 *   launch: out[(s*nrow + r)*ncol + c], Z[(s*ny + q)*nrow + r]; overlap s reads its vectors at
 *   Y + s*y_sstride (y_sstride = 0: the same Y for every overlap).                              */
int64_t apyib_lemma_prep_len(int ns, int no);
int apyib_lemma_prepare(const void *d_S, int nS, int ns, int no, void *d_prep, void *stream);
int apyib_lemma_outer(const void *d_prep, int nS, int ns, int no, int rk, const int32_t *d_rsub, int64_t nrow,
                      int ck, const int32_t *d_csub, int64_t ncol, void *d_out, void *stream);
int64_t apyib_lemma_matvec_work_len(int64_t nrow, int64_t ncol, int ny, int nS);
int apyib_lemma_matvec(const void *d_prep, int nS, int ns, int no, int rk, const int32_t *d_rsub, int64_t nrow,
                       int ck, const int32_t *d_csub, int64_t ncol, const void *d_Y, int64_t y_sstride, int ny,
                       void *d_Z, void *d_work, void *stream);

/* Host-side, bit-exact index tables ---------------------------------------------
 * get_slices (utils.py:184-213): bounds[0..7] = C_list f,o,v,t (start,stop pairs
 * flattened f0,f1,o0,o1,...) ; bounds[8..15] = I_list.                               */
int apyib_get_slices(int nbf, int ndocc, int nfzc, int spin_orbital, int32_t bounds[16]);
/* compute_all_dets enumeration (aats.py:581-618): singles (i,a) in loop order and
 * doubles (i,a,j,b) with i<j, a<b.  Pass NULL to query counts.                        */
int apyib_det_enumeration(int ndocc, int nfzc, int nvirt,
                          int32_t *h_singles, int64_t *n_singles,
                          int32_t *h_doubles, int64_t *n_doubles);
/* Row (or column) index lists for substitution lists: out[q, 0:n] = 0..n-1 with
 * out[q, sub[q,2t]] = sub[q,2t+1] + n  for t < nsub  (aats.py:583-611).               */
int apyib_det_index_lists(int n, const int32_t *h_sub, int64_t count, int nsub, int32_t *h_out);
/* Sequential pair-swap semantics of compute_SO_det (aats.py:120-130).                 */
int apyib_so_index_lists(int nso, int nocc, const int32_t *h_sub, int64_t count, int nsub,
                         int32_t *h_out);

/* ---- calibration microbenchmarks (roofline denominators; profiles/) ---------------
 * Register-resident DMMA / DFMA loops, returns achieved FLOP/s in *flops.             */
int apyib_peak_fp64(int use_dmma, int iters, double *flops, float *ms);
/* Device copy bandwidth (read+write bytes / s) over `bytes` with our own kernel.      */
int apyib_peak_copy(void *d_dst, const void *d_src, int64_t bytes, int iters, double *bytes_per_s);

#ifdef __cplusplus
}
#endif
#endif /* APYIB_B200_H */
