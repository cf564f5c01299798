// tma_tile.h — how a pass's tile (a box of T index bits, qsv_types.h) maps onto a TMA tensor map.
//
// The local register is described to the TMA unit as a row-major tensor of f64 whose dimensions are cut at the
// positions where the tile's bit segments start: dimension i covers the index bits [lo_i, lo_{i+1}) and the tile
// takes the lowest w_i of them (a box of 2^w_i at an aligned coordinate); the bits between the box and the next cut
// belong to the tile id and become the box's coordinate.  Dimension 0 is the 128-byte row (index bits 0-2, re/im
// interleaved = 16 doubles), which is also the span CU_TENSOR_MAP_SWIZZLE_128B swizzles over (qsv_types.h swz()).
// A tensor map has at most five dimensions and a box extent of at most 256: wider segments are split, and tiles with
// more than four segments above the row are moved as several boxes (the highest segments are enumerated; their bits
// fold into the last dimension's coordinate).  Tile-local order = box traversal order, so a tile lands in shared
// memory exactly where the non-TMA kernels put it.
//
// Host code (no CUDA types here): the kernel launcher turns TmaTileDesc into a CUtensorMap, the host emulation
// (tests/emu) walks the same description with a software model of the box copy.
#pragma once
#include <stdint.h>

#include "qsv_types.h"

namespace qsv {

constexpr int kTmaRank = 5;        // every tensor map is encoded with five dimensions (unused ones have extent 1)
constexpr int kTmaMaxBoxes = 32;   // boxes per tile the pipelined kernel accepts
constexpr int kTmaLastSegs = 8;

// By-value kernel argument: coordinates of a tile's boxes, straight from the tile id t.  The part of dimension i above
// its box is one contiguous field of tile-id bits (the index bits between two tile segments belong to the tile id, in
// order), so   coordinate_i(t) = ((t >> tshift[i]) & tmask[i]) << cshift[i]   - no bit deposit per tile.
struct TmaTile {
    uint32_t n_boxes;              // power of two
    uint32_t box_bytes;            // bytes per box (a multiple of 1024 when n_boxes > 1: keeps the swizzle pattern aligned)
    uint32_t tshift[kTmaRank];     // first tile-id bit of dimension i's coordinate field
    uint32_t tmask[kTmaRank];      // width mask of that field (0 for unused dimensions)
    uint32_t cshift[kTmaRank];     // log2 of the box extent of dimension i (coordinates are box-aligned); dimension 0 counts doubles
    uint32_t bshift[kTmaRank];     // index bit where dimension i's coordinate field starts (tile base = fields shifted back in place)
    uint32_t last_lo;              // index bit where the last dimension starts
    uint32_t last_dim;             // dimension that absorbs the enumerated tile bits of boxes 1..n_boxes-1
    uint32_t n_ext_ops;            // DIAG ops with external-phase tables
    uint32_t tbl_a_bits;           // table A is indexed by the low tbl_a_bits bits of the tile id, table B by the rest
    uint32_t n_last_segs;          // tile-id segments of the last dimension's coordinate (dst_lo relative to the dimension)
    Seg last_segs[kTmaLastSegs];
    int32_t box_add[kTmaMaxBoxes]; // added to coordinate last_dim for box j
    uint64_t box_off[kTmaMaxBoxes];// element offset of box j relative to the tile base (host emulation, checks)
    uint8_t ext_diag[kMaxOps];     // DevOp::diag_index of external-phase slot j
    // A launch may cover only a *slice* of the register: the tiles whose tile-id bits slice_pos[] (ascending) spell
    // slice_val.  Used to pipeline a pass, slice by slice, against a global-qubit exchange (state_api.cu run_overlapped).
    uint32_t slice_n;              // 0 = the whole register
    uint32_t slice_pos[3];
    uint32_t slice_val;            // the fixed bits, in place
};

// k-th tile id of a slice: k's bits spread around the fixed positions
QSV_HD uint32_t slice_tile_id(const TmaTile& t, uint32_t k) {
    for (uint32_t i = 0; i < t.slice_n; ++i) {
        const uint32_t p = t.slice_pos[i];
        k = ((k >> p) << (p + 1u)) | (k & ((1u << p) - 1u));
    }
    return k | t.slice_val;
}

// A slice of the local register: the amplitudes whose index bits `bit[]` (ascending, none of them a tile bit of the
// pass) spell `value` (bit i of value <-> bit[i]).
struct PassSlice {
    uint32_t n;
    uint8_t bit[3];
    uint32_t value;
};

// Fills the slice fields of a TmaTile; false if a slice bit is a tile bit of the pass.
inline bool set_tma_slice(const DevPass& hdr, const PassSlice* sl, TmaTile& t) {
    t.slice_n = 0;
    t.slice_val = 0;
    if (!sl || sl->n == 0) return true;
    if (sl->n > 3) return false;
    for (uint32_t i = 0; i < sl->n; ++i) {
        const uint64_t id_bit = extract(1ull << sl->bit[i], hdr.ext_segs, hdr.n_ext_segs);  // the tile-id bit this index bit feeds
        if (id_bit == 0 || (id_bit & (id_bit - 1)) != 0 || (i && sl->bit[i] <= sl->bit[i - 1])) return false;
        const uint32_t pos = (uint32_t)__builtin_ctzll(id_bit);
        t.slice_pos[i] = pos;
        if ((sl->value >> i) & 1u) t.slice_val |= 1u << pos;
    }
    t.slice_n = sl->n;
    return true;
}

// Host side of the tensor map (arguments of cuTensorMapEncodeTiled, f64 elements).
struct TmaTileDesc {
    uint64_t dim[kTmaRank];        // extent of every dimension in elements
    uint64_t stride_bytes[kTmaRank];  // [0] unused (dimension 0 is contiguous)
    uint32_t box[kTmaRank];
    TmaTile tile;
};

// Returns false when the tile cannot be expressed (tile does not contain index bits 0-2, too many boxes, extents beyond
// the tensor-map limits); such passes run on the synchronous kernel.
inline bool make_tma_tile(const DevPass& hdr, uint32_t n_alloc, TmaTileDesc& out) {
    out = TmaTileDesc();
    const uint32_t T = hdr.tile_bits;
    if (T < 6 || T > (uint32_t)kMaxTileBits || n_alloc < T) return false;
    // tile bits, ascending
    uint32_t bits[kMaxTileBits + 1], nb = 0;
    for (uint32_t s = 0; s < hdr.n_tile_segs; ++s)
        for (uint32_t b = 0; b < hdr.tile_segs[s].width; ++b) {
            if (nb >= T) return false;
            bits[nb++] = hdr.tile_segs[s].dst_lo + b;
        }
    if (nb != T || bits[0] != 0 || bits[1] != 1 || bits[2] != 2) return false;
    // segments above the row: maximal runs, at most 8 bits wide (box extent <= 256)
    struct S { uint32_t lo, w; } segs[2 * kMaxTileBits + 8];
    uint32_t ns = 0;
    for (uint32_t i = 3; i < T;) {
        uint32_t j = i + 1;
        while (j < T && bits[j] == bits[j - 1] + 1 && j - i < 8) ++j;
        segs[ns++] = S{bits[i], j - i};
        i = j;
    }
    // cuts of zero width where a dimension would span more than 30 index bits (coordinates are int32)
    S cut[2 * kMaxTileBits + 16];
    uint32_t nc = 0, prev_lo = 0;
    auto push = [&](S s) {
        while (s.lo - prev_lo > 30) { prev_lo += 30; cut[nc++] = S{prev_lo, 0}; }
        cut[nc++] = s;
        prev_lo = s.lo;
    };
    for (uint32_t i = 0; i < ns; ++i) push(segs[i]);
    while (n_alloc - prev_lo > 30) { prev_lo += 30; cut[nc++] = S{prev_lo, 0}; }
    // dimension 0 = the row; the next (up to) four cuts are box dimensions; the rest are enumerated
    TmaTile& t = out.tile;
    const uint32_t in_box = nc < (uint32_t)(kTmaRank - 1) ? nc : (uint32_t)(kTmaRank - 1);
    uint32_t lo[kTmaRank + 1], w[kTmaRank];
    lo[0] = 0; w[0] = 3;
    for (uint32_t i = 0; i < in_box; ++i) { lo[i + 1] = cut[i].lo; w[i + 1] = cut[i].w; }
    const uint32_t rank = in_box + 1;
    uint32_t enum_bits = 0;
    for (uint32_t i = in_box; i < nc; ++i) enum_bits += cut[i].w;
    if (enum_bits > 5) return false;  // more than kTmaMaxBoxes boxes
    t.n_boxes = 1u << enum_bits;
    uint32_t box_amp_bits = 0;
    for (uint32_t i = 0; i < rank; ++i) box_amp_bits += w[i];
    t.box_bytes = (uint32_t)sizeof(cplx) << box_amp_bits;
    if (t.n_boxes > 1 && (t.box_bytes % 1024u) != 0) return false;
    for (uint32_t j = 0; j < t.n_boxes; ++j) {
        uint64_t off = 0;
        uint32_t src = 0;
        for (uint32_t i = in_box; i < nc; ++i)
            for (uint32_t b = 0; b < cut[i].w; ++b, ++src)
                if ((j >> src) & 1u) off |= 1ull << (cut[i].lo + b);
        t.box_off[j] = off;
    }
    uint64_t ext_mask = 0;  // index bits that belong to the tile id
    for (uint32_t sgi = 0; sgi < hdr.n_ext_segs; ++sgi) ext_mask |= ((1ull << hdr.ext_segs[sgi].width) - 1ull) << hdr.ext_segs[sgi].dst_lo;
    t.last_dim = rank - 1;
    for (uint32_t i = 0; i < (uint32_t)kTmaRank; ++i) {
        if (i < rank) {
            const uint32_t hi = (i + 1 < rank) ? lo[i + 1] : n_alloc;  // the last dimension runs to the top of the register
            const uint32_t span = hi - lo[i];
            if (span > 30 || span < w[i]) return false;
            const uint32_t field_lo = lo[i] + w[i];
            // every bit of the field [field_lo, hi) except the enumerated tile bits (last dimension only) is a tile-id bit
            t.tshift[i] = (uint32_t)__builtin_popcountll(ext_mask & ((1ull << field_lo) - 1ull));
            const uint32_t field_bits = (uint32_t)__builtin_popcountll(ext_mask & (((1ull << hi) - 1ull) ^ ((1ull << field_lo) - 1ull)));
            if (i + 1 < rank && field_bits != hi - field_lo) return false;  // (cannot happen: a tile bit inside would have started a dimension)
            t.tmask[i] = field_bits >= 32 ? 0xffffffffu : ((1u << field_bits) - 1u);
            t.cshift[i] = w[i] + (i == 0 ? 1u : 0u);
            t.bshift[i] = field_lo;
            out.dim[i] = (1ull << span) * (i == 0 ? 2ull : 1ull);
            out.box[i] = (1u << w[i]) * (i == 0 ? 2u : 1u);
            out.stride_bytes[i] = (uint64_t)sizeof(cplx) << lo[i];
        } else {
            t.tshift[i] = 0;
            t.tmask[i] = 0;  // coordinate 0
            t.cshift[i] = 0;
            t.bshift[i] = 0;
            out.dim[i] = 1;
            out.box[i] = 1;
            out.stride_bytes[i] = (uint64_t)sizeof(cplx) << n_alloc;
        }
    }
    // Last dimension: with several boxes per tile the enumerated tile bits lie inside its field, between tile-id bits, so
    // its coordinate is assembled from the tile-id segments of the field (one segment for single-box tiles).
    for (uint32_t j = 0; j < t.n_boxes; ++j) t.box_add[j] = (int32_t)(t.box_off[j] >> lo[rank - 1]);
    t.n_last_segs = 0;
    const uint32_t last_field_lo = lo[rank - 1] + w[rank - 1];
    for (uint32_t sgi = 0; sgi < hdr.n_ext_segs; ++sgi) {
        const Seg& e = hdr.ext_segs[sgi];
        const uint32_t e_lo = e.dst_lo, e_hi = e_lo + e.width;
        if (e_hi <= last_field_lo) continue;
        const uint32_t start = e_lo > last_field_lo ? e_lo : last_field_lo;  // a zero-width cut may split a tile-id segment
        if (t.n_last_segs >= (uint32_t)kTmaLastSegs) return false;
        t.last_segs[t.n_last_segs++] = Seg{(uint8_t)(e.src_lo + (start - e_lo)), (uint8_t)(e_hi - start), (uint8_t)(start - lo[rank - 1]), 0};
    }
    t.tmask[rank - 1] = 0;  // handled by last_segs
    t.last_lo = lo[rank - 1];
    return true;
}

// Coordinates (in elements of each dimension) of box j of tile t.
QSV_HD void tma_tile_coords(const TmaTile& t, uint32_t tile_id, uint32_t j, int32_t (&c)[kTmaRank]) {
#pragma unroll
    for (int i = 0; i < kTmaRank; ++i) c[i] = (int32_t)(((tile_id >> t.tshift[i]) & t.tmask[i]) << t.cshift[i]);
    int32_t last = t.box_add[j];
    for (uint32_t sgi = 0; sgi < t.n_last_segs; ++sgi)
        last += (int32_t)(((tile_id >> t.last_segs[sgi].src_lo) & ((1u << t.last_segs[sgi].width) - 1u)) << t.last_segs[sgi].dst_lo);
#pragma unroll
    for (int i = 0; i < kTmaRank; ++i)
        if ((uint32_t)i == t.last_dim) c[i] = last;
}

// Element offset of tile t in the register (= deposit(t, ext_segs)), from the same fields.
QSV_HD uint64_t tma_tile_base(const TmaTile& t, uint32_t tile_id) {
    uint64_t base = 0;
#pragma unroll
    for (int i = 0; i < kTmaRank; ++i) base |= (uint64_t)((tile_id >> t.tshift[i]) & t.tmask[i]) << t.bshift[i];
    for (uint32_t sgi = 0; sgi < t.n_last_segs; ++sgi)
        base |= (uint64_t)((tile_id >> t.last_segs[sgi].src_lo) & ((1u << t.last_segs[sgi].width) - 1u)) << (t.last_segs[sgi].dst_lo + t.last_lo);
    return base;
}

// Software model of one tiled-mode box copy with CU_TENSOR_MAP_SWIZZLE_128B (host emulation and unit tests):
// calls f(global element offset in doubles, shared-memory byte offset) for every f64 of the box.
template <class F>
inline void tma_box_walk(const TmaTileDesc& d, const int32_t (&c)[kTmaRank], F f) {
    uint64_t n = 1;
    for (int i = 0; i < kTmaRank; ++i) n *= d.box[i];
    for (uint64_t lin = 0; lin < n; ++lin) {
        uint64_t rem = lin, goff = 0;
        for (int i = 0; i < kTmaRank; ++i) {
            const uint64_t b = rem % d.box[i];
            rem /= d.box[i];
            const uint64_t coord = (uint64_t)c[i] + b;
            goff += i == 0 ? coord : coord * (d.stride_bytes[i] / sizeof(double));
        }
        const uint64_t byte = lin * sizeof(double);
        f(goff, byte ^ (((byte >> 7) & 7ull) << 4));
    }
}

}  // namespace qsv
