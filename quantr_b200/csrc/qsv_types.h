// qsv_types.h — POD layouts shared by the host scheduler and the sm_100a kernels.
//
// A *plan* is a list of *passes*.  One pass streams the whole local state once:
// every CTA pulls a tile of 2^T amplitudes (T "tile bits" of the amplitude index,
// always including the lowest `low_bits` so global accesses stay coalesced) into
// shared memory, runs the pass's *rounds* on it and writes it back in place.
// A register round keeps 16 amplitudes (4 "register bits") per thread in registers
// and applies a list of lowered ops to them; a dense round applies one k-qubit
// Custom matrix (CSR) to the tile.
//
// Index convention (reference: src/circuit/states/product_states.rs:205-215):
// wire q of an n-qubit circuit is physical index bit n-1-q.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define QSV_HD __host__ __device__ __forceinline__
#else
#define QSV_HD inline
#endif

namespace qsv {

#ifndef QSV_REG_BITS
#define QSV_REG_BITS 4  // amplitudes per thread = 2^QSV_REG_BITS (build parameter; host and device must agree)
#endif
constexpr int kRegBits = QSV_REG_BITS;
constexpr int kSlots = 1 << kRegBits;
constexpr int kMaxSlots = 16;         // array extents in the blob layout (independent of the build parameter)
static_assert(kRegBits == 3 || kRegBits == 4, "QSV_REG_BITS must be 3 or 4");
constexpr int kMinQubits = 4;      // states smaller than one register group are padded with idle top bits
constexpr int kMaxTileBits = 13;
constexpr int kMaxSegs = 16;
constexpr int kMaxThrSegs = 8;
constexpr int kMaxRounds = 32;
constexpr int kMaxOps = 96;
constexpr int kSmallRounds = 8;   // capacity classes of the kernel-parameter pass descriptor
constexpr int kSmallOps = 24;
constexpr int kMaxLoads = kMaxSlots; // global loads per thread per tile (= kSlots: one CTA thread per register group)
constexpr int kSmallTileThreads = 64; // CTA size for tiles below 2^10 amplitudes

// CTA size for a tile of 2^T amplitudes: one thread per register group of 16 amplitudes.
QSV_HD constexpr uint32_t tile_threads(uint32_t T) { return T >= 10 ? (1u << (T - kRegBits)) : (uint32_t)kSmallTileThreads; }
constexpr uint32_t kPassMagic = 0x51535632u;  // "QSV2"
constexpr uint32_t kNoExtSlot = 0xffffffffu;

struct alignas(16) cplx {
    double x, y;
};

// deposit: result |= ((v >> src_lo) & mask(width)) << dst_lo
struct Seg {
    uint8_t src_lo, width, dst_lo, pad;
};

enum OpType : uint32_t {
    OP_MAT_GENERAL = 0,   // complex 2x2 on one register bit
    OP_MAT_REAL = 1,      // real 2x2 (Ry)
    OP_MAT_ANTIDIAG = 2,  // a0' = m01*a1, a1' = m10*a0 (Y, X90, CY, ...)
    OP_MAT_XSWAP = 3,     // a0 <-> a1 (X, CNot, Toffoli)
    OP_DIAG = 4,          // amp *= exp(i*pi*(theta0 + sum_b coef_b*bit_b)) where control mask holds
    OP_DENSE = 5,         // k-qubit Custom matrix (dense round)
    OP_MAT_HADAMARD = 6   // unnormalised butterfly a0' = a0+a1, a1' = a0-a1; the 1/sqrt2 factors of a pass are
                          // collected in DevPass::final_scale and applied once when the tile is stored
    ,
    OP_QFT4 = 7           // macro-op, see kCodeQft4
};

// Fully resolved dispatch code of an op inside a register round (one jump-table entry per routine):
//   MAT : 1 + kind*8 + ctrl*4 + slot, kind = 0 HADAMARD, 1 XSWAP, 2 REAL, 3 GENERAL, 4 ANTIDIAG; ctrl = has register controls
//   DIAG: 41 + has_reg*6 + sel,        sel = 0 all slots, 1..4 slots with register bit sel-1 set, 5 runtime mask
constexpr int kDiagTblLen = 64;  // lo[32] (thread-index bits 0-4) + hi[32] (bits 5-9)
//   HD  : 53 + has_reg*4 + slot: an uncontrolled Hadamard on register bit `slot` fused with the DIAG op that follows it
//         and is controlled by exactly that bit (one stage of a QFT / phase-estimation ladder)
//   QFT4: 61: a whole register round of a QFT / phase-estimation ladder as one op - four HD stages on four physically
//         adjacent register bits, highest first, whose register-bit phases are the QFT ones (pi/2, pi/4, pi/8).  The
//         DevOp carries no constants; the DIAG ops of the (3 or 4, DevOp::slot) stages follow it in the op array outside
//         the round's op range and supply the tile/thread factors (pass_core.h qft4_apply).
//   DUAL: a REAL or GENERAL MAT code | kCodeDualFlag: a 2x2 op with two matrices (MatFlags::MAT_DUAL) - m[0..8) where all
//         its controls hold, m[8..16) everywhere else; such an op is never skipped
constexpr uint32_t kCodeNop = 0, kCodeMatBase = 1, kCodeDiagBase = 41, kCodeHdBase = 53, kCodeQft4 = 61, kCodeCount = 62, kCodeDualFlag = 128;
enum PassFlags : uint32_t {
    PASS_DIRECT_STORE = 2,  // the last round writes its registers straight to global memory (coalesced: its register bits
                            // exclude the three lowest tile bits)
    PASS_UNCONDITIONAL = 4  // no op of the pass has a control among the thread or tile-index bits and every round is a
                            // register round: every thread of every tile runs the whole op list (the FAST builds of the
                            // kernels: no control masks, no dense or permutation rounds compiled in)
};
constexpr int kMaxRoundOps = 32;  // ops per register round (one 32-bit active mask)

enum DiagFlags : uint32_t {
    DIAG_HAS_THR_LO = 1,
    DIAG_HAS_THR_HI = 2,
    DIAG_HAS_REG = 4,   // some register bit carries a linear term
    DIAG_HAS_W = 8,     // the tile/thread factor w is not identically 1
    // bits 8..22: which entries of the register-constant table are not 1 (see DevOp::m)
    DIAG_NONTRIVIAL_SHIFT = 8
};
enum MatFlags : uint32_t {
    MAT_DUAL = 1  // DevOp::m holds two matrices: m[0..8) where every control holds, m[8..16) elsewhere
};
// ROUND_PERM: a run of X / CNot / Toffoli / multi-controlled X gates (OP_MAT_XSWAP) as one gather through the tile: thread
// and slot layout as in a register round; every thread fetches, for each of its 16 tile-local indices l, the amplitude at
// P^-1(l) (the ops applied to the index in reverse order), and writes it to l after a barrier.  Targets and controls may
// be any tile bits (DevOp::slot = tile-local target position, DevOp::cmask_thr = all tile-local controls).  Ordinary ops on
// the round's register bits may follow the gather (DevRound::n_ops of them), as in a register round.
enum RoundType : uint32_t { ROUND_REG = 0, ROUND_DENSE = 1, ROUND_PERM = 2 };

struct DiagExtTerm {
    uint32_t bit;  // physical bit (outside the tile, may be a rank bit)
    uint32_t pad;
    double coef;   // half-turns
};

// 192 bytes
struct DevOp {
    uint32_t type;
    uint32_t slot;        // MAT: register slot 0..3 of the target bit
    uint32_t cmask_reg;   // controls among the register slots (4 bits)
    uint32_t cmask_thr;   // controls among the tile-local, non-register bits (tile-local positions)
    uint64_t cmask_ext;   // controls outside the tile (physical bit positions, rank bits included)
    uint32_t flags;       // DIAG: DiagFlags; MAT: MatFlags
    uint32_t diag_index;  // DIAG: slot in the per-tile external-phase array
    double m[16];         // MAT: m00 m01 m10 m11 (re, im) in m[0..8); MAT_DUAL: the matrix for unsatisfied controls in m[8..16).
                          // DIAG with one register control bit c: m[2(q-1)], m[2(q-1)+1] = exp(i*pi*sum of the coefs of the
                          //   free register bits in q), q = 1..7 over the register bits other than c in ascending order;
                          // DIAG otherwise: m[2k], m[2k+1] = r_k = exp(i*pi*coef of register bit k), k = 0..3.
                          // flags bit (DIAG_NONTRIVIAL_SHIFT + q - 1) resp. (+ k): the entry differs from 1.
    uint32_t ext_off;     // DIAG: byte offset in the pass blob of DiagExtTerm[n_ext]
    uint32_t n_ext;
    uint32_t tbl_off;     // DIAG: byte offset of the thread-phase table cplx lo[32], hi[16] (kDiagTblLen entries)
    uint32_t dense_off;   // DENSE: byte offset of DevDense
    uint32_t code;        // dispatch code (see kCode*)
    uint32_t ext_slot;    // DIAG with n_ext > 0: index of its pair of external-phase tables (ext_tables.h); else kNoExtSlot
    double theta0;        // DIAG: constant term (half-turns)
};
static_assert(sizeof(DevOp) == 192, "DevOp layout");

// 128 bytes
struct DevRound {
    uint32_t type;
    uint32_t first_op;
    uint32_t n_ops;
    uint32_t n_thr_segs;
    uint8_t reg_pos[4];  // tile-local positions of the 4 register bits, ascending
    uint32_t perm_first;  // ROUND_PERM: ops [perm_first, perm_first + n_perm) are the gather (X gates); ops [first_op, first_op + n_ops)
    uint32_t n_perm;      // are ordinary register-round ops applied to the gathered amplitudes before they are stored
    uint32_t pad;
    Seg thr_segs[kMaxThrSegs];  // thread index e -> tile-local index with register bits zero
    uint32_t xoff[kMaxSlots];   // byte offset of swz(slot s's tile-local offset): address = (swz(lb) << 4) ^ xoff[s]
};
static_assert(sizeof(DevRound) == 128, "DevRound layout");

struct DevDense {
    uint32_t k;
    uint32_t gate_mask;   // tile-local mask of the gate bits
    uint32_t rowptr_off;  // byte offsets in the pass blob
    uint32_t coloff_off;  // uint32 tile-local offset of each non-zero's input sub-state
    uint32_t val_off;     // cplx value of each non-zero
    uint32_t nnz;
    uint8_t gpos[16];     // tile-local position of sub-index bit e (e = 0 is the MSB = first control)
};

struct DevPass {
    uint32_t magic;
    uint32_t tile_bits;
    uint32_t n_rounds;
    uint32_t n_ops;
    uint32_t n_diag;
    uint32_t n_tile_segs;
    uint32_t n_ext_segs;
    uint32_t flags;
    uint64_t n_tiles;
    uint32_t rounds_off;  // byte offsets in the pass blob
    uint32_t ops_off;
    uint32_t blob_bytes;
    uint32_t threads;         // CTA size = tile_threads(tile_bits)
    double final_scale;       // product of the deferred 1/sqrt2 factors of the pass's Hadamards
    uint32_t ext_ctrl_mask[3];// bit o set: op o has controls outside the tile (evaluated once per tile)
    uint32_t max_ext;         // largest DevOp::n_ext of the pass (stride of the shared-memory copy of the term lists)
    uint32_t n_ext_ops;       // DIAG ops whose phase depends on bits outside the tile (DevOp::ext_slot = 0..n_ext_ops-1)
    uint32_t reserved0;
    Seg tile_segs[kMaxSegs];  // tile-local index -> physical (local) offset
    Seg ext_segs[kMaxSegs];   // tile id -> physical (local) base
};
static_assert(sizeof(DevPass) == 216, "DevPass layout");

// Per-load constants of the tile load/store phase: thread `tid` moves tile-local elements l = i*threads + tid.
// deposit() and swz() are linear over disjoint bit sets, so both split into a per-thread and a per-i part.
struct DevLoads {
    uint64_t goff[kMaxLoads];  // deposit(i * threads, tile_segs): element offset in the state
    uint32_t soff[kMaxLoads];  // swz(i * threads) << 4: byte offset in the shared-memory tile
    uint32_t pad[2];
    uint64_t store_goff[kMaxSlots];  // PASS_DIRECT_STORE: deposit(slot s's tile-local offset in the last round, tile_segs)
};
static_assert(sizeof(DevLoads) == 328, "DevLoads layout");

// The part of a pass blob the kernel receives by value (kernel-parameter constant bank): header, load constants,
// rounds and ops.  Tables, external phase terms and Custom matrices stay in the blob in global memory.
template <int NR, int NO>
struct PassParams {
    DevPass hdr;
    DevLoads loads;
    DevRound rounds[NR];
    DevOp ops[NO];
};
static_assert(sizeof(PassParams<kMaxRounds, kMaxOps>) < 32000, "kernel parameter limit");

QSV_HD uint64_t deposit(uint64_t v, const Seg* segs, uint32_t n) {
    uint64_t r = 0;
    for (uint32_t i = 0; i < n; ++i) r |= ((v >> segs[i].src_lo) & ((1ull << segs[i].width) - 1ull)) << segs[i].dst_lo;
    return r;
}

// inverse of deposit on the bits the segments cover
QSV_HD uint64_t extract(uint64_t v, const Seg* segs, uint32_t n) {
    uint64_t r = 0;
    for (uint32_t i = 0; i < n; ++i) r |= ((v >> segs[i].dst_lo) & ((1ull << segs[i].width) - 1ull)) << segs[i].src_lo;
    return r;
}

// Shared-memory swizzle of a tile-local index.  It is the layout the TMA unit writes for a tensor map with
// CU_TENSOR_MAP_SWIZZLE_128B and a 128-byte inner box (8 amplitudes): the 16-byte unit inside each 128-byte row is
// XORed with the row number mod 8 (byte-address bits 4-6 ^= bits 7-9).  A quarter-warp's 128-bit accesses hit all 32
// banks when its three varying tile-local bits lie below bit 6 with distinct (position mod 3); the scheduler orders the
// thread bits of every round accordingly (emit_pass).  Every kernel - TMA or not - and the host emulation use this one
// layout.
QSV_HD uint32_t swz(uint32_t l) { return l ^ ((l >> 3) & 7u); }

QSV_HD cplx cmul(cplx a, cplx b) { return cplx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }

}  // namespace qsv
