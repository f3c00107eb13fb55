// mbl_mask_reads (include/metabuli_b200.h): tantan masking of a batch of reads on the host, for `--mask 1`.
// Compiled with g++ -ffp-contract=off -mfma (csrc/Makefile): see tantan_mask.hpp for why the rounding is pinned.
#include "tantan_mask.hpp"

#include "../../../include/metabuli_b200.h"

extern "C" int mbl_mask_reads(char* bases, const uint64_t* offsets, uint32_t n_reads, float mask_prob, int threads) {
    if ((!bases && n_reads && offsets && offsets[n_reads] > 0) || !offsets) return MBL_E_BAD_ARG;
    try {
        mblhost::tantan_mask_reads(bases, offsets, n_reads, mask_prob, threads > 0 ? (unsigned)threads : std::thread::hardware_concurrency());
    } catch (...) {
        return MBL_E_HOST;
    }
    return MBL_OK;
}
