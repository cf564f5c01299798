// peer_swap.h — index arithmetic of the global-qubit exchange over NVLink peer memory (K5), shared by the sm_100a
// kernel (kernels.cu) and the host emulation harness (tests/emu) so the CPU suite checks it at every world size.
//
// An EXCHANGE step swaps rank bit j with local bit partner[j] (ascending), j < g, 2^g ranks.  For one pair of ranks
// that is: the block of rank r whose partner bits spell the peer p is swapped, element for element, with the block
// of rank p whose partner bits spell r.  Each rank of the pair moves one half of the block.
#pragma once
#include <stdint.h>

#include "qsv_types.h"

namespace qsv {

struct SwapArgs {
    uint64_t block_len;      // amplitudes per block: 2^(n_local - g)
    uint64_t first, count;   // sub-range of the block handled by this rank
    uint64_t local_spell;    // partner-bit pattern that spells the peer (in this rank's shard)
    uint64_t remote_spell;   // partner-bit pattern that spells this rank (in the peer's shard)
    uint32_t g;
    uint8_t partner[8];      // ascending
};

inline SwapArgs make_swap_args(uint32_t n_local, const uint8_t* partner, uint32_t g, int rank, int peer) {
    SwapArgs a{};
    a.g = g;
    a.block_len = (1ull << n_local) >> g;
    for (uint32_t k = 0; k < g; ++k) {
        a.partner[k] = partner[k];
        if ((peer >> k) & 1) a.local_spell |= 1ull << partner[k];
        if ((rank >> k) & 1) a.remote_spell |= 1ull << partner[k];
    }
    const uint64_t half = a.block_len / 2;  // the lower rank of the pair moves the first half, the higher rank the second
    a.first = rank < peer ? 0 : half;
    a.count = rank < peer ? half : a.block_len - half;
    return a;
}

// j-th element of a block (partner bits zero): j's bits spread around the partner positions
QSV_HD uint64_t insert_zero_bits(uint64_t j, const SwapArgs& a) {
    for (uint32_t k = 0; k < a.g; ++k) {
        const uint32_t p = a.partner[k];
        j = ((j >> p) << (p + 1)) | (j & ((1ull << p) - 1ull));
    }
    return j;
}

}  // namespace qsv
