// Prefix-shared LU (dets_pairs_impl.cuh): K = 2 instantiations, n = 10..12
#include "dets_pairs_impl.cuh"

namespace apyib {

int launch_det_pairs_k2_large(int n, APYIB_PAIRS_ARGS_DECL) {
    switch (n) {
        APYIB_PFX_CASE(10, 2) APYIB_PFX_CASE(11, 2) APYIB_PFX_CASE(12, 2)
    }
    return APYIB_ERR_UNSUPPORTED;
}

}  // namespace apyib
