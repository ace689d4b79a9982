// Prefix-shared LU: dispatch over (n, k) and the K = 1 instantiations; kernel in dets_pairs_impl.cuh, the K = 2
// instantiations in dets_pairs_k2_small.cu / dets_pairs_k2_large.cu (three translation units compile in parallel).
#include "dets_pairs_impl.cuh"

namespace apyib {

int g_pairs_variant = 0;

int launch_det_pairs_k1(int n, APYIB_PAIRS_ARGS_DECL) {
    switch (n) {
        APYIB_PFX_CASE(2, 1) APYIB_PFX_CASE(3, 1) APYIB_PFX_CASE(4, 1) APYIB_PFX_CASE(5, 1) APYIB_PFX_CASE(6, 1)
        APYIB_PFX_CASE(7, 1) APYIB_PFX_CASE(8, 1) APYIB_PFX_CASE(9, 1) APYIB_PFX_CASE(10, 1) APYIB_PFX_CASE(11, 1)
        APYIB_PFX_CASE(12, 1)
    }
    return APYIB_ERR_UNSUPPORTED;
}

// warps resident on the device (one block per SM) -- the host sizes the group chunks with it
int pairs_total_warps(int n, int k, int ns, int nc) {
    bool ssm;
    int slots;
    size_t smem;
    int T = 0;
    switch (n * 4 + k) {
#define APYIB_PFX_T(NN, KK) case NN * 4 + KK: T = pfx_block<NN, KK>(ns, nc, &ssm, &slots, &smem); break;
        APYIB_PFX_T(3, 2) APYIB_PFX_T(4, 2) APYIB_PFX_T(5, 2) APYIB_PFX_T(6, 2) APYIB_PFX_T(7, 2) APYIB_PFX_T(8, 2)
        APYIB_PFX_T(9, 2) APYIB_PFX_T(10, 2) APYIB_PFX_T(11, 2) APYIB_PFX_T(12, 2)
        APYIB_PFX_T(2, 1) APYIB_PFX_T(3, 1) APYIB_PFX_T(4, 1) APYIB_PFX_T(5, 1) APYIB_PFX_T(6, 1) APYIB_PFX_T(7, 1)
        APYIB_PFX_T(8, 1) APYIB_PFX_T(9, 1) APYIB_PFX_T(10, 1) APYIB_PFX_T(11, 1) APYIB_PFX_T(12, 1)
#undef APYIB_PFX_T
    }
    return 148 * (T / 32);
}

int launch_det_pairs(int n, int k, APYIB_PAIRS_ARGS_DECL) {
    int rc = APYIB_ERR_UNSUPPORTED;
    if (k == 1 && n >= 2 && n <= 12) rc = launch_det_pairs_k1(n, APYIB_PAIRS_ARGS);
    else if (k == 2 && n >= 3 && n <= 9) rc = launch_det_pairs_k2_small(n, APYIB_PAIRS_ARGS);
    else if (k == 2 && n >= 10 && n <= 12) rc = launch_det_pairs_k2_large(n, APYIB_PAIRS_ARGS);
    else set_error("prefix-shared LU: (n, k) = (%d, %d) not instantiated", n, k);
    return rc;
}

}  // namespace apyib
