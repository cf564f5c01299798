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

// A *slice* restricts the swap to the amplitudes whose index bits slice_bit[] spell slice_value (the pipelined exchange
// moves the shard slice by slice): the slice bits join the partner bits as fixed bits of both blocks.
struct SwapArgs {
    uint64_t block_len;      // amplitudes per block (and slice): 2^(n_local - g - slice bits)
    uint64_t first, count;   // sub-range of the block handled by this rank
    uint64_t local_spell;    // partner-bit pattern that spells the peer (in this rank's shard)
    uint64_t remote_spell;   // partner-bit pattern that spells this rank (in the peer's shard)
    uint32_t g;              // fixed bits: partner bits + slice bits
    uint8_t partner[8];      // their positions, ascending
};

inline SwapArgs make_swap_args(uint32_t n_local, const uint8_t* partner, uint32_t g, int rank, int peer, uint32_t n_slice = 0, const uint8_t* slice_bit = nullptr,
                               uint32_t slice_value = 0) {
    SwapArgs a{};
    a.g = g + n_slice;  // <= 8
    a.block_len = (1ull << n_local) >> a.g;
    for (uint32_t k = 0; k < g; ++k) {
        a.partner[k] = partner[k];
        if ((peer >> k) & 1) a.local_spell |= 1ull << partner[k];
        if ((rank >> k) & 1) a.remote_spell |= 1ull << partner[k];
    }
    for (uint32_t k = 0; k < n_slice; ++k) {
        a.partner[g + k] = slice_bit[k];
        if ((slice_value >> k) & 1u) {
            a.local_spell |= 1ull << slice_bit[k];
            a.remote_spell |= 1ull << slice_bit[k];
        }
    }
    for (uint32_t i = 1; i < a.g; ++i)  // insert_zero_bits needs ascending positions
        for (uint32_t j = i; j > 0 && a.partner[j] < a.partner[j - 1]; --j) { const uint8_t t = a.partner[j]; a.partner[j] = a.partner[j - 1]; a.partner[j - 1] = t; }
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
