// Prefix-shared LU (dets_pairs_impl.cuh): K = 2 instantiations, n = 3..9
#include "dets_pairs_impl.cuh"

namespace apyib {

int launch_det_pairs_k2_small(int n, APYIB_PAIRS_ARGS_DECL) {
    switch (n) {
        APYIB_PFX_CASE(3, 2) APYIB_PFX_CASE(4, 2) APYIB_PFX_CASE(5, 2) APYIB_PFX_CASE(6, 2) APYIB_PFX_CASE(7, 2)
        APYIB_PFX_CASE(8, 2) APYIB_PFX_CASE(9, 2)
    }
    return APYIB_ERR_UNSUPPORTED;
}

}  // namespace apyib
