// pass_core.h — per-thread body of a fused pass, written __host__ __device__.
//
// The sm_100a kernel in kernels.cu calls these functions from its 256 threads with the
// tile in shared memory and the pass descriptor in the kernel-parameter constant bank.
// tests/emu/ compiles the same functions with g++ and walks the threads sequentially so the
// host scheduler + blob encoding can be unit-tested without a GPU; that harness is test
// infrastructure and is never linked into libqsv.so.
//
// Replaces, for a fused list of gates: Circuit::apply_gate (src/circuit/simulation.rs:64-135),
// insert_gate_image_into_product_state (:158-180) and the gate columns of
// src/circuit/standard_gate_ops.rs:37-267 (as lowered by plan.cpp).
#pragma once
#include <math.h>
#include <string.h>

#include "qsv_types.h"

namespace qsv {

// sin(pi*x), cos(pi*x) with exact results at multiples of 1/2.
QSV_HD void sincospi_hd(double x, double* s, double* c) {
#ifdef __CUDA_ARCH__
    ::sincospi(x, s, c);
#else
    double r = fmod(x, 2.0);  // exact
    if (r < 0) r += 2.0;      // [0, 2)
    // quadrant q in 0..4 with r = q/2 + f, f in [-1/4, 1/4]
    const double q = floor(r * 2.0 + 0.5);
    const double f = r - q * 0.5;  // exact
    const double sf = sin(M_PI * f), cf = cos(M_PI * f);
    switch (((int)q) & 3) {
        case 0: *s = sf; *c = cf; break;
        case 1: *s = cf; *c = -sf; break;
        case 2: *s = -sf; *c = -cf; break;
        default: *s = -cf; *c = sf; break;
    }
    if (f == 0.0) {  // exact multiples of 1/2: no signed zeros
        if (*s == 0.0) *s = 0.0;
        if (*c == 0.0) *c = 0.0;
    }
#endif
}

QSV_HD uint32_t ctz32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return (uint32_t)(__ffs((int)v) - 1);
#else
    return (uint32_t)__builtin_ctz(v);
#endif
}

// Copies the by-value part of a pass blob into the kernel-parameter struct (host side).
template <int NR, int NO>
inline bool fill_params(const uint8_t* blob, PassParams<NR, NO>& out) {
    const DevPass* h = reinterpret_cast<const DevPass*>(blob);
    if (h->magic != kPassMagic || h->n_rounds > (uint32_t)NR || h->n_ops > (uint32_t)NO) return false;
    memcpy(&out.hdr, blob, sizeof(DevPass));
    memcpy(&out.loads, blob + sizeof(DevPass), sizeof(DevLoads));
    if (h->n_rounds) memcpy(out.rounds, blob + h->rounds_off, sizeof(DevRound) * h->n_rounds);
    if (h->n_ops) memcpy(out.ops, blob + h->ops_off, sizeof(DevOp) * h->n_ops);
    return true;
}

// External phase of a DIAG op for one tile: exp(i*pi*(theta0 + sum over bits outside the tile)).
QSV_HD cplx diag_ext_phase_terms(double theta0, const DiagExtTerm* terms, uint32_t n_ext, uint64_t base_full) {
    double ang = theta0;
    for (uint32_t i = 0; i < n_ext; ++i)
        if ((base_full >> terms[i].bit) & 1ull) ang += terms[i].coef;
    cplx r;
    sincospi_hd(ang, &r.y, &r.x);
    return r;
}
QSV_HD cplx diag_ext_phase(const DevOp& op, const uint8_t* blob, uint64_t base_full) {
    return diag_ext_phase_terms(op.theta0, reinterpret_cast<const DiagExtTerm*>(blob + op.ext_off), op.n_ext, base_full);
}

// External-phase tables of the pipelined kernel: the tile id t is split into its low `a` bits and the rest, and
//   exp(i*pi*(theta0 + sum over set bits outside the tile)) = A[t & (2^a - 1)] * B[t >> a]
// with A carrying theta0 and the rank bits.  One entry (built once per plan upload, on the device):
//   part   = the tile-id bits of this half, in place (low half: i, high half: i << a)
//   is_low = the entry belongs to table A
QSV_HD cplx ext_table_entry(const DevPass& hdr, double theta0, const DiagExtTerm* terms, uint32_t n_ext, uint64_t part, uint64_t rank_hi, bool is_low) {
    const uint64_t base = deposit(part, hdr.ext_segs, hdr.n_ext_segs) | (is_low ? rank_hi : 0ull);
    return diag_ext_phase_terms(is_low ? theta0 : 0.0, terms, n_ext, base);
}
QSV_HD uint32_t ext_table_low_bits(uint64_t n_tiles) {
    uint32_t nt = 0;
    while ((1ull << nt) < n_tiles) ++nt;
    return (nt + 1) / 2;
}
// entries of table A + table B for a pass with n_tiles tiles
QSV_HD uint64_t ext_table_len(uint64_t n_tiles) {
    const uint32_t a = ext_table_low_bits(n_tiles);
    return (1ull << a) + ((n_tiles + (1ull << a) - 1) >> a);
}

// ---- in-place FP64 primitives ---------------------------------------------------------------------
// Every op below updates the thread's 16 amplitudes strictly in place: the device versions are single
// PTX instructions with tied ("+d") operands so the 64 data registers keep one home through the op
// loop (an out-of-place formulation costs a 64-register copy per op plus spills); the host versions
// perform the same IEEE operations (fma is correctly rounded on both sides).
#ifdef __CUDA_ARCH__
QSV_HD double f_mul(double a, double b) { double d; asm("mul.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b)); return d; }
QSV_HD void f_mul_ip(double& x, double c) { asm("mul.f64 %0, %0, %1;" : "+d"(x) : "d"(c)); }
QSV_HD void f_fma_acc(double& acc, double a, double b) { asm("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc) : "d"(a), "d"(b)); }
QSV_HD void f_fma_self(double& x, double c, double t) { asm("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(c), "d"(t)); }
QSV_HD void f_add_ip(double& x, double y) { asm("add.f64 %0, %0, %1;" : "+d"(x) : "d"(y)); }
QSV_HD void f_mov(double& x, double y) { asm("mov.f64 %0, %1;" : "=d"(x) : "d"(y)); }
#else
QSV_HD double f_mul(double a, double b) { return a * b; }
QSV_HD void f_mul_ip(double& x, double c) { x = x * c; }
QSV_HD void f_fma_acc(double& acc, double a, double b) { acc = fma(a, b, acc); }
QSV_HD void f_fma_self(double& x, double c, double t) { x = fma(x, c, t); }
QSV_HD void f_add_ip(double& x, double y) { x = x + y; }
QSV_HD void f_mov(double& x, double y) { x = y; }
#endif

// x' = x + y, y' = x - y without a temporary: y' = x' - 2y (one extra rounding of x' enters y').
QSV_HD void f_bfly(double& x, double& y) {
    f_add_ip(x, y);
    f_fma_self(y, -2.0, x);
}

// a *= f (complex), in place with two temporaries and no moves.
QSV_HD void c_mul_ip(cplx& a, double fr, double fi) {
    const double u = f_mul(a.y, -fi), v = f_mul(a.x, fi);
    f_fma_self(a.x, fr, u);  // a.x*fr - a.y*fi
    f_fma_self(a.y, fr, v);  // a.y*fr + a.x*fi
}

// ---- 2x2 ops on register slot J of the 16 amplitudes a thread holds -------------------------
// CTRL = false: no control among the register slots (the common case): straight-line code.
// m = {m00.re, m00.im, m01.re, m01.im, m10.re, m10.im, m11.re, m11.im}

// DUAL (MatFlags::MAT_DUAL): no pair is skipped; a pair takes m[0..8) where every control of the op holds (`sat`: its
// controls outside the registers, per thread; CTRL: and the pair's register controls) and m[8..16) otherwise.
template <int J, bool CTRL, bool DUAL = false>
QSV_HD void mat_general(cplx (&a)[kSlots], const double* m, uint32_t cm, bool sat = true) {
    if constexpr (J < kRegBits) {
    if (DUAL && !CTRL && !sat) m += 8;
#pragma unroll
    for (int s0 = 0; s0 < kSlots; ++s0) {
        if ((s0 >> J) & 1) continue;
        if (!DUAL && CTRL && (s0 & cm) != cm) continue;
        const double* mm = (DUAL && CTRL && !(sat && (s0 & cm) == cm)) ? m + 8 : m;
        const double ar = mm[0], ai = mm[1], br = mm[2], bi = mm[3], cr = mm[4], ci = mm[5], dr = mm[6], di = mm[7];
        cplx& x = a[s0];
        cplx& y = a[s0 | (1 << J)];
        // four chains of three into temporaries; each output's last FMA consumes (and overwrites) its own input
        double t0 = f_mul(x.y, -ai);      // x'.re = ar*xr - ai*xi + br*yr - bi*yi
        f_fma_acc(t0, br, y.x);
        f_fma_acc(t0, -bi, y.y);
        double t1 = f_mul(x.x, ai);       // x'.im = ar*xi + ai*xr + br*yi + bi*yr
        f_fma_acc(t1, br, y.y);
        f_fma_acc(t1, bi, y.x);
        double t2 = f_mul(y.y, -di);      // y'.re = dr*yr - di*yi + cr*xr - ci*xi
        f_fma_acc(t2, cr, x.x);
        f_fma_acc(t2, -ci, x.y);
        double t3 = f_mul(y.x, di);       // y'.im = dr*yi + di*yr + cr*xi + ci*xr
        f_fma_acc(t3, cr, x.y);
        f_fma_acc(t3, ci, x.x);
        f_fma_self(x.x, ar, t0);
        f_fma_self(x.y, ar, t1);
        f_fma_self(y.x, dr, t2);
        f_fma_self(y.y, dr, t3);
    }
    }  // J < kRegBits
}

template <int J, bool CTRL, bool DUAL = false>
QSV_HD void mat_real(cplx (&a)[kSlots], const double* m, uint32_t cm, bool sat = true) {
    if constexpr (J < kRegBits) {
    if (DUAL && !CTRL && !sat) m += 8;
#pragma unroll
    for (int s0 = 0; s0 < kSlots; ++s0) {
        if ((s0 >> J) & 1) continue;
        if (!DUAL && CTRL && (s0 & cm) != cm) continue;
        const double* mm = (DUAL && CTRL && !(sat && (s0 & cm) == cm)) ? m + 8 : m;
        const double ar = mm[0], br = mm[2], cr = mm[4], dr = mm[6];
        cplx& x = a[s0];
        cplx& y = a[s0 | (1 << J)];
        const double t0 = f_mul(br, y.x), t1 = f_mul(br, y.y);
        f_mul_ip(y.x, dr);
        f_mul_ip(y.y, dr);
        f_fma_acc(y.x, cr, x.x);
        f_fma_acc(y.y, cr, x.y);
        f_fma_self(x.x, ar, t0);
        f_fma_self(x.y, ar, t1);
    }
    }  // J < kRegBits
}

template <int J, bool CTRL>
QSV_HD void mat_hadamard(cplx (&a)[kSlots], const double*, uint32_t cm) {
    if constexpr (J < kRegBits) {
#pragma unroll
    for (int s0 = 0; s0 < kSlots; ++s0) {
        if ((s0 >> J) & 1) continue;
        if (CTRL && (s0 & cm) != cm) continue;
        cplx& x = a[s0];
        cplx& y = a[s0 | (1 << J)];
        f_bfly(x.x, y.x);
        f_bfly(x.y, y.y);
    }
    }  // J < kRegBits
}

template <int J, bool CTRL>
QSV_HD void mat_antidiag(cplx (&a)[kSlots], const double* m, uint32_t cm) {
    if constexpr (J < kRegBits) {
    const double br = m[2], bi = m[3], cr = m[4], ci = m[5];
#pragma unroll
    for (int s0 = 0; s0 < kSlots; ++s0) {
        if ((s0 >> J) & 1) continue;
        if (CTRL && (s0 & cm) != cm) continue;
        cplx& x = a[s0];
        cplx& y = a[s0 | (1 << J)];
        // x' = b*y, y' = c*x
        const double t0 = f_mul(br, y.x), t1 = f_mul(br, y.y), s0v = f_mul(-bi, y.y), s1v = f_mul(bi, y.x);
        y.x = f_mul(cr, x.x);
        f_fma_acc(y.x, -ci, x.y);
        y.y = f_mul(cr, x.y);
        f_fma_acc(y.y, ci, x.x);
        f_mov(x.x, t0);
        f_add_ip(x.x, s0v);
        f_mov(x.y, t1);
        f_add_ip(x.y, s1v);
    }
    }  // J < kRegBits
}

template <int J, bool CTRL>
QSV_HD void mat_xswap(cplx (&a)[kSlots], const double*, uint32_t cm) {
    if constexpr (J < kRegBits) {
#pragma unroll
    for (int s0 = 0; s0 < kSlots; ++s0) {
        if ((s0 >> J) & 1) continue;
        if (CTRL && (s0 & cm) != cm) continue;
        cplx& x = a[s0];
        cplx& y = a[s0 | (1 << J)];
        double t;
        f_mov(t, x.x); f_mov(x.x, y.x); f_mov(y.x, t);
        f_mov(t, x.y); f_mov(x.y, y.y); f_mov(y.y, t);
    }
    }  // J < kRegBits
}

// Which ops of a pass keep thread-dependent phases where: see choose_diag_mode() below.
#ifndef QSV_OCC_NUM
#define QSV_OCC_NUM 2  // CTAs per SM at T = 12 (x2 at T = 11, x4 at T = 10, /2 at T = 13); build parameter for occupancy experiments
#endif
#ifndef QSV_OCC_T11
#define QSV_OCC_T11 (2u * QSV_OCC_NUM)
#endif
QSV_HD constexpr uint32_t tile_min_blocks(uint32_t T) { return T >= 13 ? (QSV_OCC_NUM / 2 ? QSV_OCC_NUM / 2 : 1u) : T == 12 ? QSV_OCC_NUM : T == 11 ? QSV_OCC_T11 : 4u * QSV_OCC_NUM; }
//   2: per-thread phases [n_diag][threads] in shared memory, computed once per launch
//   1: the lo/hi tables [n_diag][kDiagTblLen] in shared memory      0: tables read from global memory
inline size_t pass_smem_bytes(uint32_t T, uint32_t n_diag, int mode) {
    size_t b = (sizeof(cplx) << T) + sizeof(cplx) * (n_diag + 1);
    if (mode == 2) b += sizeof(cplx) * (size_t)n_diag * tile_threads(T);
    if (mode == 1) b += sizeof(cplx) * (size_t)n_diag * kDiagTblLen;
    return b;
}
inline int choose_diag_mode(uint32_t T, uint32_t n_diag) {
    if (n_diag == 0) return 0;
    const size_t budget = (size_t)(227 * 1024) / tile_min_blocks(T) - 1024;
    if (pass_smem_bytes(T, n_diag, 1) <= budget) return 1;  // measured faster than mode 2 on B200 (less shared memory per CTA)
    if (pass_smem_bytes(T, n_diag, 2) <= budget) return 2;
    return 0;
}

// The uniform fast path of the pass kernel: every thread of every tile runs the whole op list (no thread-bit or tile-index controls), the pass fits
// the small parameter class and the per-thread phase table fits next to the tile.  Launcher and emulator agree on it.
inline bool pass_is_fast(const DevPass& hdr) {
    if (!(hdr.flags & PASS_UNCONDITIONAL) || hdr.tile_bits < 10) return false;
    if (hdr.n_rounds > (uint32_t)kSmallRounds || hdr.n_ops > (uint32_t)kSmallOps) return false;
    const size_t budget = (size_t)(227 * 1024) / tile_min_blocks(hdr.tile_bits) - 1024;
    return pass_smem_bytes(hdr.tile_bits, hdr.n_diag, 2) <= budget;
}

// Basis-state initialisation fused into the first pass of a plan: the pass does not read the register; every tile is
// synthesised (all zero, except the one tile that holds the basis amplitude).
//   mode 1: every tile is synthesised and computed;  mode 2: all-zero tiles are written as zeros without arithmetic
//   (a linear pass maps a zero tile to a zero tile).
struct PassInit {
    uint64_t base_full;  // tile base (rank bits included) of the tile holding the basis amplitude
    uint32_t local;      // its tile-local index
    uint32_t mode;
    uint64_t ext_mask;   // index bits (of the local shard) that belong to the tile id; base_full & ext_mask identifies the tile's rows
    uint32_t holds_here; // the tile lies in this rank's shard
    uint32_t n_alloc;    // log2 of the shard length
    double amp_re, amp_im;  // the amplitude (1 unless a sharded plan folded leading gates into the initial state, plan.h Plan::prefix)
    // A folded prefix on the top `sup_bits` local index bits (Plan::prefix_local_bits): the register holds 2^sup_bits
    // amplitudes, amp_tbl[j] at the local index whose top sup_bits bits spell j and whose other bits are those of `x_local`.
    uint32_t sup_bits;
    uint32_t sup_local_mask;  // the support bits that are tile bits of this pass, as a tile-local mask
    uint64_t sup_mask;        // the support bits (local index positions)
    uint64_t x_local;         // local index of the basis state
    const cplx* amp_tbl;
    // mode 2: the tiles that hold amplitudes, enumerated directly - tile id = hold_id_val | pdep(q, hold_id_mask), q < 2^popc(mask)
    // (dealt evenly to the compute groups of all CTAs; a CTA scanning only its own tile ids would find them on a few CTAs)
    uint32_t hold_id_mask, hold_id_val;
};
QSV_HD uint32_t pdep32(uint32_t v, uint32_t mask) {
    uint32_t r = 0;
    for (uint32_t m = mask; m; m &= m - 1u) {
        if (v & 1u) r |= m & (0u - m);
        v >>= 1;
    }
    return r;
}
// does the tile at `base_full` hold a non-zero amplitude of the initial state?
QSV_HD bool init_tile_holds(const PassInit& pi, uint64_t base_full) { return ((base_full ^ pi.base_full) & ~pi.sup_mask) == 0; }
// initial amplitude of tile-local element l of a holding tile at local base `base` (tile_segs: the pass's tile bits)
QSV_HD cplx init_tile_element(const PassInit& pi, const DevPass& hdr, uint64_t base, uint32_t l) {
    if (((l ^ pi.local) & ~pi.sup_local_mask) != 0) return cplx{0.0, 0.0};
    if (pi.sup_bits == 0) return cplx{pi.amp_re, pi.amp_im};
    const uint64_t p = base | (pi.sup_local_mask ? deposit(l, hdr.tile_segs, hdr.n_tile_segs) : 0ull);
    return pi.amp_tbl[(p & pi.sup_mask) >> (pi.n_alloc - pi.sup_bits)];
}
inline PassInit make_pass_init(const DevPass& hdr, uint64_t phys_index, uint32_t n_local, uint32_t mode, uint32_t sup_bits = 0, const cplx* amp_tbl = nullptr) {
    const uint64_t local_mask = (1ull << n_local) - 1ull;
    const uint64_t ext_mask = deposit(hdr.n_tiles - 1ull, hdr.ext_segs, hdr.n_ext_segs);
    PassInit pi;
    pi.sup_bits = sup_bits;
    pi.sup_mask = sup_bits ? (((1ull << sup_bits) - 1ull) << (n_local - sup_bits)) : 0ull;
    pi.sup_local_mask = (uint32_t)extract(pi.sup_mask, hdr.tile_segs, hdr.n_tile_segs);
    pi.x_local = phys_index & local_mask;
    pi.amp_tbl = amp_tbl;
    pi.hold_id_mask = (uint32_t)extract(pi.sup_mask & ext_mask, hdr.ext_segs, hdr.n_ext_segs);
    pi.hold_id_val = (uint32_t)extract(phys_index & local_mask & ext_mask & ~pi.sup_mask, hdr.ext_segs, hdr.n_ext_segs);
    pi.base_full = (phys_index & local_mask & ext_mask) | (phys_index & ~local_mask);
    pi.local = (uint32_t)extract(phys_index & local_mask, hdr.tile_segs, hdr.n_tile_segs);
    pi.mode = mode;
    pi.ext_mask = ext_mask;
    pi.holds_here = 1;  // the launcher compares the rank bits (launch_pass knows rank_hi)
    pi.n_alloc = n_local;
    pi.amp_re = 1.0;
    pi.amp_im = 0.0;
    return pi;
}

// Thread phase of a DIAG op: product of its lo/hi table entries for thread-group e (tile-independent).
QSV_HD cplx diag_thread_phase(const DevOp& op, const cplx* tbl, uint32_t e) {
    cplx w{1.0, 0.0};
    if (op.flags & DIAG_HAS_THR_LO) w = tbl[e & 31u];
    if (op.flags & DIAG_HAS_THR_HI) w = cmul(w, tbl[32u + (e >> 5)]);
    return w;
}

// ctx.thr_phase != null: per-thread phases precomputed once per launch ([diag_index * threads + e], shared memory);
// otherwise ctx.thr_tbl (shared memory) or the blob (global memory) holds the lo/hi tables.
struct DiagCtx {
    const uint8_t* blob;
    const cplx* ext_phase;
    const cplx* thr_tbl;
    const cplx* thr_phase;
    uint32_t threads;
};

// Tile/thread factor of a DIAG op for thread-group e.  FAST: external phase x precomputed thread phase, no flag tests.
template <bool FAST>
QSV_HD cplx diag_w(const DevOp& op, const DiagCtx& ctx, uint32_t e) {
    cplx w = ctx.ext_phase[op.diag_index];
    if constexpr (FAST) {
        return cmul(w, ctx.thr_phase[op.diag_index * ctx.threads + e]);
    } else {
        if (ctx.thr_phase) {
            if (op.flags & (DIAG_HAS_THR_LO | DIAG_HAS_THR_HI)) w = cmul(w, ctx.thr_phase[op.diag_index * ctx.threads + e]);
        } else {
            const cplx* tbl = ctx.thr_tbl ? ctx.thr_tbl + op.diag_index * kDiagTblLen : reinterpret_cast<const cplx*>(ctx.blob + op.tbl_off);
            if (op.flags & DIAG_HAS_THR_LO) w = cmul(w, tbl[e & 31u]);
            if (op.flags & DIAG_HAS_THR_HI) w = cmul(w, tbl[32u + (e >> 5)]);
        }
        return w;
    }
}

// q-th subset (q = 0..7) of the three register bits other than bit C, as a slot mask
template <int C>
QSV_HD constexpr int free_subset(int q) {
    int s = 0, k = 0;
    for (int b = 0; b < 4; ++b) {
        if (b == C) continue;
        if ((q >> k) & 1) s |= 1 << b;
        ++k;
    }
    return s;
}

// amp[s] *= w * R[s] on the slots selected by the register-control mask, R[s] = prod_{k: bit k of s} r_k.
// SEL = 0: all slots, SEL = 1..4: the slots with register bit SEL-1 set, SEL = 5: generic runtime mask.
// Two waves of independent in-place complex multiplies: first the tile/thread factor w on every selected slot, then
// the register-bit constants (uniform operands from the constant bank, trivial entries skipped).
template <int SEL, bool HAS_REG, bool FAST>
QSV_HD void diag_apply(cplx (&a)[kSlots], const DevOp& op, const DiagCtx& ctx, uint32_t e) {
    constexpr bool kSingle = SEL >= 1 && SEL <= kRegBits;
    constexpr int kCtl = kSingle ? SEL - 1 : 0;
    constexpr int kBit = kSingle ? (1 << kCtl) : 0;
    if constexpr (SEL > kRegBits && SEL <= 4) return;  // no such register bit in this build
    const uint32_t cm = op.cmask_reg;
    if (op.flags & DIAG_HAS_W) {
        const cplx w = diag_w<FAST>(op, ctx, e);
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
            if (kSingle && !(s & kBit)) continue;
            if (SEL == 5 && (s & cm) != cm) continue;
            c_mul_ip(a[s], w.x, w.y);
        }
    }
    if (!HAS_REG) return;
    const uint32_t nontrivial = op.flags >> DIAG_NONTRIVIAL_SHIFT;
    if (kSingle) {
#pragma unroll
        for (int q = 1; q < (1 << (kRegBits - 1)); ++q) {
            constexpr int dummy = 0;
            (void)dummy;
            const int s = kBit | free_subset<kCtl>(q);
            if (s >= kSlots) continue;
            if ((nontrivial >> (q - 1)) & 1u) c_mul_ip(a[s], op.m[2 * (q - 1)], op.m[2 * (q - 1) + 1]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < kRegBits; ++k) {
            if (!((nontrivial >> k) & 1u)) continue;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                if (!((s >> k) & 1)) continue;
                if (SEL == 5 && (s & cm) != cm) continue;
                c_mul_ip(a[s], op.m[2 * k], op.m[2 * k + 1]);
            }
        }
    }
}

// Hadamard on register bit J fused with the DIAG op controlled by that bit: the phase factor is fetched first so its
// latency hides behind the butterflies.
template <int J, bool HAS_REG, bool FAST>
QSV_HD void hd_apply(cplx (&a)[kSlots], const DevOp& op, const DiagCtx& ctx, uint32_t e) {
    if constexpr (J < kRegBits) {
        cplx w{1.0, 0.0};
        // FAST: fetched unconditionally (the tables hold 1 where the op has no tile/thread dependence) so the loads
        // issue ahead of the butterflies
        const bool has_w = FAST || (op.flags & DIAG_HAS_W) != 0;
        if (has_w) w = diag_w<FAST>(op, ctx, e);
#pragma unroll
        for (int s0 = 0; s0 < kSlots; ++s0) {
            if ((s0 >> J) & 1) continue;
            cplx& x = a[s0];
            cplx& y = a[s0 | (1 << J)];
            f_bfly(x.x, y.x);
            f_bfly(x.y, y.y);
        }
        if (has_w) {
#pragma unroll
            for (int s = 0; s < kSlots; ++s)
                if ((s >> J) & 1) c_mul_ip(a[s], w.x, w.y);
        }
        if (HAS_REG) {
            const uint32_t nontrivial = op.flags >> DIAG_NONTRIVIAL_SHIFT;
#pragma unroll
            for (int q = 1; q < (1 << (kRegBits - 1)); ++q) {
                const int s = (1 << J) | free_subset<J>(q);
                if (s >= kSlots) continue;
                if ((nontrivial >> (q - 1)) & 1u) c_mul_ip(a[s], op.m[2 * (q - 1)], op.m[2 * (q - 1) + 1]);
            }
        }
    }
}

// One register round of a QFT ladder on four physically adjacent register bits (kCodeQft4): a radix-16
// decimation-in-frequency butterfly.  Stage j (register bit j, highest first) is an unnormalised Hadamard followed, on the
// half with bit j set, by exp(i*pi*(lower register bits)/2^(j-i)) - compile-time constants, two of them powers of i and
// free - times the stage's tile/thread factor w_j.  The w_j commute with the later stages (those act on lower bits only),
// so they are applied once at the end as w3^s3 w2^s2 w1^s1 w0^s0: 26 complex multiplies instead of 32, and no per-stage
// table of register constants.  288 FP64 instructions per 16 amplitudes against 4 x 81 for four hd_apply calls.
// `diag` = the stages' DIAG ops (n_diag = 4, or 3 when the lowest stage is a bare Hadamard), highest stage first.
QSV_HD void cmul_to(cplx& a, double cr, double ci) {  // a *= (cr, ci)
    const double x = a.x * cr - a.y * ci, y = a.x * ci + a.y * cr;
    a.x = x;
    a.y = y;
}
//   n_stages = 4, or 3: the ladder covers register bits 2..0 only and bit 3 is a passenger.
template <bool FAST>
QSV_HD void qft4_apply(cplx (&a)[kSlots], const DevOp* diag, uint32_t n_diag, uint32_t n_stages, const DiagCtx& ctx, uint32_t e) {
    if constexpr (kRegBits == 4) {
        constexpr double kR2 = 0.70710678118654752440;  // cos(pi/4)
        constexpr double kC8 = 0.92387953251128675613;  // cos(pi/8)
        constexpr double kS8 = 0.38268343236508977173;  // sin(pi/8)
        // tile/thread factors first: their loads overlap the butterflies
        cplx w[4];  // w[3 - j] = factor of the stage on register bit j
        const uint32_t skip = 4u - n_stages;
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = ((uint32_t)j >= skip && (uint32_t)j - skip < n_diag) ? diag_w<FAST>(diag[(uint32_t)j - skip], ctx, e) : cplx{1.0, 0.0};
        // stage 3: pairs (s, s|8); odd half times W16^s, W16 = exp(i*pi/8)
        if (n_stages == 4)
#pragma unroll
        for (int s0 = 0; s0 < 8; ++s0) {
            const cplx x = a[s0], y = a[s0 | 8];
            a[s0] = cplx{x.x + y.x, x.y + y.y};
            const double tr = x.x - y.x, ti = x.y - y.y;
            cplx& o = a[s0 | 8];
            switch (s0) {
                case 0: o = cplx{tr, ti}; break;
                case 1: o = cplx{tr * kC8 - ti * kS8, tr * kS8 + ti * kC8}; break;
                case 2: o = cplx{(tr - ti) * kR2, (tr + ti) * kR2}; break;
                case 3: o = cplx{tr * kS8 - ti * kC8, tr * kC8 + ti * kS8}; break;
                case 4: o = cplx{-ti, tr}; break;
                case 5: o = cplx{-tr * kS8 - ti * kC8, tr * kC8 - ti * kS8}; break;
                case 6: o = cplx{(-tr - ti) * kR2, (tr - ti) * kR2}; break;
                default: o = cplx{-tr * kC8 - ti * kS8, tr * kS8 - ti * kC8}; break;
            }
        }
        // stage 2: pairs (s, s|4); odd half times W8^(s&3), W8 = exp(i*pi/4)
#pragma unroll
        for (int s0 = 0; s0 < 16; ++s0) {
            if (s0 & 4) continue;
            const cplx x = a[s0], y = a[s0 | 4];
            a[s0] = cplx{x.x + y.x, x.y + y.y};
            const double tr = x.x - y.x, ti = x.y - y.y;
            cplx& o = a[s0 | 4];
            switch (s0 & 3) {
                case 0: o = cplx{tr, ti}; break;
                case 1: o = cplx{(tr - ti) * kR2, (tr + ti) * kR2}; break;
                case 2: o = cplx{-ti, tr}; break;
                default: o = cplx{(-tr - ti) * kR2, (tr - ti) * kR2}; break;
            }
        }
        // stage 1: pairs (s, s|2); odd half times i^(s&1)
#pragma unroll
        for (int s0 = 0; s0 < 16; ++s0) {
            if (s0 & 2) continue;
            const cplx x = a[s0], y = a[s0 | 2];
            a[s0] = cplx{x.x + y.x, x.y + y.y};
            const double tr = x.x - y.x, ti = x.y - y.y;
            a[s0 | 2] = (s0 & 1) ? cplx{-ti, tr} : cplx{tr, ti};
        }
        // stage 0: pairs (s, s|1)
#pragma unroll
        for (int s0 = 0; s0 < 16; s0 += 2) {
            const cplx x = a[s0], y = a[s0 | 1];
            a[s0] = cplx{x.x + y.x, x.y + y.y};
            a[s0 | 1] = cplx{x.x - y.x, x.y - y.y};
        }
        // tile/thread factors: slot s gets w3^s3 w2^s2 w1^s1 w0^s0, two bits at a time (three factors live per half)
        {
            const cplx w1 = w[2], w0 = w[3], w10 = cmul(w1, w0);
#pragma unroll
            for (int s = 0; s < 16; ++s) {
                if ((s & 3) == 1) cmul_to(a[s], w0.x, w0.y);
                if ((s & 3) == 2) cmul_to(a[s], w1.x, w1.y);
                if ((s & 3) == 3) cmul_to(a[s], w10.x, w10.y);
            }
            const cplx w3 = w[0], w2 = w[1], w32 = cmul(w3, w2);
#pragma unroll
            for (int s = 0; s < 16; ++s) {
                if ((s >> 2) == 1) cmul_to(a[s], w2.x, w2.y);
                if ((s >> 2) == 2) cmul_to(a[s], w3.x, w3.y);
                if ((s >> 2) == 3) cmul_to(a[s], w32.x, w32.y);
            }
        }
    }
}

// Dispatch code of a lowered op inside a register round (host side; see kCode* in qsv_types.h).
inline uint32_t op_dispatch_code(const DevOp& op) {
    int kind = -1;
    switch (op.type) {
        case OP_MAT_HADAMARD: kind = 0; break;
        case OP_MAT_XSWAP: kind = 1; break;
        case OP_MAT_REAL: kind = 2; break;
        case OP_MAT_GENERAL: kind = 3; break;
        case OP_MAT_ANTIDIAG: kind = 4; break;
        default: break;
    }
    if (kind >= 0) return (kCodeMatBase + (uint32_t)kind * 8u + (op.cmask_reg ? 4u : 0u) + op.slot) | ((op.flags & MAT_DUAL) ? kCodeDualFlag : 0u);  // dual: REAL or GENERAL
    if (op.type == OP_DIAG) {
        uint32_t sel = 5;
        if (op.cmask_reg == 0) sel = 0;
        else if ((op.cmask_reg & (op.cmask_reg - 1)) == 0) sel = 1 + (op.cmask_reg == 1 ? 0 : op.cmask_reg == 2 ? 1 : op.cmask_reg == 4 ? 2 : 3);
        return kCodeDiagBase + ((op.flags & DIAG_HAS_REG) ? 6u : 0u) + sel;
    }
    return kCodeNop;
}

// c = the op's matrix, already in registers (for a dual op without register controls: the one this thread applies)
#define QSV_MAT_CASES(KIND, FN)                                                              \
    case kCodeMatBase + KIND * 8 + 0: FN<0, false>(a, c, 0u); break;                         \
    case kCodeMatBase + KIND * 8 + 1: FN<1, false>(a, c, 0u); break;                         \
    case kCodeMatBase + KIND * 8 + 2: FN<2, false>(a, c, 0u); break;                         \
    case kCodeMatBase + KIND * 8 + 3: FN<3, false>(a, c, 0u); break;                         \
    case kCodeMatBase + KIND * 8 + 4: FN<0, true>(a, c, op.cmask_reg); break;                \
    case kCodeMatBase + KIND * 8 + 5: FN<1, true>(a, c, op.cmask_reg); break;                \
    case kCodeMatBase + KIND * 8 + 6: FN<2, true>(a, c, op.cmask_reg); break;                \
    case kCodeMatBase + KIND * 8 + 7: FN<3, true>(a, c, op.cmask_reg); break;
// REAL and GENERAL may be dual: with register controls the choice of matrix is per pair
#define QSV_MAT2_CTRL_CASE(KIND, FN, J)                                                      \
    case kCodeMatBase + KIND * 8 + 4 + J:                                                    \
        if (dual) FN<J, true, true>(a, op.m, op.cmask_reg, sat);                             \
        else FN<J, true>(a, c, op.cmask_reg);                                                \
        break;
#define QSV_MAT2_CASES(KIND, FN)                                                             \
    case kCodeMatBase + KIND * 8 + 0: FN<0, false>(a, c, 0u); break;                         \
    case kCodeMatBase + KIND * 8 + 1: FN<1, false>(a, c, 0u); break;                         \
    case kCodeMatBase + KIND * 8 + 2: FN<2, false>(a, c, 0u); break;                         \
    case kCodeMatBase + KIND * 8 + 3: FN<3, false>(a, c, 0u); break;                         \
    QSV_MAT2_CTRL_CASE(KIND, FN, 0)                                                          \
    QSV_MAT2_CTRL_CASE(KIND, FN, 1)                                                          \
    QSV_MAT2_CTRL_CASE(KIND, FN, 2)                                                          \
    QSV_MAT2_CTRL_CASE(KIND, FN, 3)

#define QSV_HD_CASES(HAS_REG)                                                                   \
    case kCodeHdBase + (HAS_REG ? 4 : 0) + 0: hd_apply<0, HAS_REG, FAST>(a, op, ctx, e); break;       \
    case kCodeHdBase + (HAS_REG ? 4 : 0) + 1: hd_apply<1, HAS_REG, FAST>(a, op, ctx, e); break;       \
    case kCodeHdBase + (HAS_REG ? 4 : 0) + 2: hd_apply<2, HAS_REG, FAST>(a, op, ctx, e); break;       \
    case kCodeHdBase + (HAS_REG ? 4 : 0) + 3: hd_apply<3, HAS_REG, FAST>(a, op, ctx, e); break;

// (the FAST build tests for the macro-op ahead of the dispatch trees: no second copy of its body there)
#define QSV_QFT4_CASE case kCodeQft4: if constexpr (!FAST) qft4_apply<FAST>(a, &op + 1, op.slot, op.cmask_reg, ctx, e); break;

#define QSV_DIAG_CASES(HAS_REG)                                                                                                    \
    case kCodeDiagBase + (HAS_REG ? 6 : 0) + 0: diag_apply<0, HAS_REG, FAST>(a, op, ctx, e); break;           \
    case kCodeDiagBase + (HAS_REG ? 6 : 0) + 1: diag_apply<1, HAS_REG, FAST>(a, op, ctx, e); break;           \
    case kCodeDiagBase + (HAS_REG ? 6 : 0) + 2: diag_apply<2, HAS_REG, FAST>(a, op, ctx, e); break;           \
    case kCodeDiagBase + (HAS_REG ? 6 : 0) + 3: diag_apply<3, HAS_REG, FAST>(a, op, ctx, e); break;           \
    case kCodeDiagBase + (HAS_REG ? 6 : 0) + 4: diag_apply<4, HAS_REG, FAST>(a, op, ctx, e); break;           \
    case kCodeDiagBase + (HAS_REG ? 6 : 0) + 5: diag_apply<5, HAS_REG, FAST>(a, op, ctx, e); break;

// Which ops of the pass act on thread-group e (controls among the thread's fixed tile-local bits).  Tile-independent:
// the kernel evaluates it once per launch.  W words of 32 op bits.
template <int W>
QSV_HD void thread_active_mask(const DevPass& hdr, const DevRound* rounds, const DevOp* ops, uint32_t e, uint32_t (&act)[W]) {
#pragma unroll
    for (int w = 0; w < W; ++w) act[w] = 0xffffffffu;
    for (uint32_t r = 0; r < hdr.n_rounds; ++r) {
        if (rounds[r].type == ROUND_DENSE) continue;  // (the gather ops of a permutation round lie outside its op range)
        const uint32_t lb = (uint32_t)deposit(e, rounds[r].thr_segs, rounds[r].n_thr_segs);
        for (uint32_t o = rounds[r].first_op; o < rounds[r].first_op + rounds[r].n_ops; ++o)
            if ((lb & ops[o].cmask_thr) != ops[o].cmask_thr) {
#pragma unroll
                for (int w = 0; w < W; ++w)
                    if ((int)(o >> 5) == w) act[w] &= ~(1u << (o & 31u));
            }
    }
}

// Clears the ops whose controls outside the tile are not satisfied for this tile (uniform over the CTA).
template <int W>
QSV_HD void tile_active_mask(const DevPass& hdr, const DevOp* ops, uint64_t base_full, uint32_t (&act)[W]) {
#pragma unroll
    for (int w = 0; w < W; ++w) {
        uint32_t m = hdr.ext_ctrl_mask[w];
        while (m) {
            const uint32_t b = ctz32(m);
            m &= m - 1;
            const uint32_t o = 32u * w + b;
            if ((base_full & ops[o].cmask_ext) != ops[o].cmask_ext) act[w] &= ~(1u << b);
        }
    }
}

// A register round for thread-group e (0 <= e < 2^(T-4)) is: round_load (16 amplitudes from the swizzled tile into
// registers), round_ops (the round's ops, in place), then round_store_tile or - for the last round of a pass flagged
// PASS_DIRECT_STORE - round_store_global.
QSV_HD uint32_t round_thread_base(const DevRound& R, uint32_t e) { return (uint32_t)deposit(e, R.thr_segs, R.n_thr_segs); }

QSV_HD void round_load(const DevRound& R, uint32_t lb, const cplx* tile, cplx (&a)[kSlots]) {
    const uint32_t sb = swz(lb) << 4;
    const char* tb = reinterpret_cast<const char*>(tile);
#pragma unroll
    for (int s = 0; s < kSlots; ++s) a[s] = *reinterpret_cast<const cplx*>(tb + (sb ^ R.xoff[s]));
}

// Permutation round (ROUND_PERM): a[s] = tile[P^-1(l_s)], l_s = the thread's tile-local index of slot s.  The round's ops
// are applied to the 16 indices in reverse order (each op is its own inverse): idx ^= target bit where the op's tile-local
// controls hold.  act: as in round_ops (an op whose controls outside the tile fail for this tile is skipped).
template <int W, bool FAST = false>
QSV_HD void round_perm_load(const DevRound& R, const DevOp* ops, const uint32_t (&act)[W], uint32_t lb, const cplx* tile, cplx (&a)[kSlots]) {
    uint32_t idx[kSlots];
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
        uint32_t off = 0;
#pragma unroll
        for (int j = 0; j < kRegBits; ++j)
            if ((s >> j) & 1) off |= 1u << R.reg_pos[j];
        idx[s] = lb | off;
    }
    const uint32_t first = R.perm_first, n = R.n_perm;
    for (uint32_t j = n; j-- > 0;) {
        const uint32_t o = first + j;
        if constexpr (!FAST) {
            uint32_t word = act[0];
#pragma unroll
            for (int w = 1; w < W; ++w)
                if ((int)(o >> 5) == w) word = act[w];
            if (!((word >> (o & 31u)) & 1u)) continue;
        }
        const uint32_t cm = ops[o].cmask_thr, tb = 1u << ops[o].slot;
#pragma unroll
        for (int s = 0; s < kSlots; ++s)
            if ((idx[s] & cm) == cm) idx[s] ^= tb;
    }
#pragma unroll
    for (int s = 0; s < kSlots; ++s) a[s] = tile[swz(idx[s])];
}

QSV_HD void round_store_tile(const DevRound& R, uint32_t lb, cplx* tile, const cplx (&a)[kSlots]) {
    const uint32_t sb = swz(lb) << 4;
    char* tb = reinterpret_cast<char*>(tile);
#pragma unroll
    for (int s = 0; s < kSlots; ++s) *reinterpret_cast<cplx*>(tb + (sb ^ R.xoff[s])) = a[s];
}

// last round of a pass that leaves through the tile buffer (TMA store): the deferred 1/sqrt2 factors are applied here
QSV_HD void round_store_tile_scaled(const DevRound& R, uint32_t lb, cplx* tile, const cplx (&a)[kSlots], double scale) {
    const uint32_t sb = swz(lb) << 4;
    char* tb = reinterpret_cast<char*>(tile);
#pragma unroll
    for (int s = 0; s < kSlots; ++s) *reinterpret_cast<cplx*>(tb + (sb ^ R.xoff[s])) = cplx{a[s].x * scale, a[s].y * scale};
}

//   act : bit o set = every control of op o outside the registers holds for this thread-group and tile (W words).  A plain
//         op whose bit is clear is skipped; a dual op (kCodeDualFlag) runs either way and picks its matrix by the bit.
//   FAST: the uniform fast path (pass_is_fast): act is ignored, every control holds
// The op list is walked with a counted loop: op index, dispatch code and the ops' constants are uniform over the CTA;
// only the test of the thread's act bit diverges.  The op's matrix is fetched ahead of the dispatch tree so that the
// constant-bank latency overlaps it (fetching the next op's dispatch code ahead as well costs a register the kernel
// does not have: it spills inside this loop).
#ifndef QSV_PRELOAD
#define QSV_PRELOAD 0  // 1: the 2x2 routines get their constants loaded ahead of the dispatch tree - measured slower on B200 (2.60 s vs 2.01 s on config 3: the 16 extra live registers spill inside the op loop)
#endif
template <int W, bool FAST = false>
QSV_HD void round_ops(const DevRound& R, const DevOp* ops, const DiagCtx& ctx, const uint32_t (&act)[W], uint32_t e, cplx (&a)[kSlots]) {
    const uint32_t first = R.first_op, n = R.n_ops;  // n <= kMaxRoundOps
    if (n == 0) return;
    for (uint32_t o = first; o < first + n; ++o) {
        const DevOp& op = ops[o];
        const uint32_t full_code = op.code;
        const bool dual = (full_code & kCodeDualFlag) != 0;
        const uint32_t code = full_code & (kCodeDualFlag - 1u);
        bool sat = true;
        if constexpr (!FAST) {
            uint32_t word = act[0];
#pragma unroll
            for (int w = 1; w < W; ++w)
                if ((int)(o >> 5) == w) word = act[w];
            sat = ((word >> (o & 31u)) & 1u) != 0;
            if (!sat && !dual) continue;
        }
        if constexpr (FAST) {
            if (code == kCodeQft4) {  // the hot case of QFT passes: tested ahead of the dispatch tree
                qft4_apply<true>(a, &op + 1, op.slot, op.cmask_reg, ctx, e);
                continue;
            }
        }
        // two dispatch trees: the 2x2 routines need nothing but the op's constants; the phase routines share the
        // DiagCtx address arithmetic, which the compiler would otherwise hoist in front of every op
        if (code < kCodeDiagBase) {
#if QSV_PRELOAD
            double c[8];
            {
                const double* src = op.m + ((dual && !sat) ? 8 : 0);
#pragma unroll
                for (int i = 0; i < 8; ++i) c[i] = src[i];
            }
#else
            const double* c = op.m + ((dual && !sat) ? 8 : 0);
#endif
            switch (code) {
                QSV_MAT_CASES(0, mat_hadamard)
                QSV_MAT_CASES(1, mat_xswap)
                QSV_MAT2_CASES(2, mat_real)
                QSV_MAT2_CASES(3, mat_general)
                QSV_MAT_CASES(4, mat_antidiag)
                default: break;
            }
        } else {
            switch (code) {
                QSV_QFT4_CASE
                QSV_DIAG_CASES(false)
                QSV_DIAG_CASES(true)
                QSV_HD_CASES(false)
                QSV_HD_CASES(true)
                default: break;
            }
        }
    }
}

// Dense (Custom) round, phase 1: thread-group e computes outputs l = 16*e .. 16*e+15 from the
// tile; phase 2 (after a barrier) writes them back.  Implements
//   out[t] = sum_s M'[t][s] * in[s]
// where M' already carries the reference's None rule (src/circuit/simulation.rs:120-133):
// rows/columns of untouched sub-states are replaced by identity rows / zero columns on the host.
QSV_HD void dense_compute(const DevDense& D, const uint8_t* blob, uint32_t e, const cplx* tile, cplx (&out)[kSlots]) {
    const uint32_t* rowptr = reinterpret_cast<const uint32_t*>(blob + D.rowptr_off);
    const uint32_t* coloff = reinterpret_cast<const uint32_t*>(blob + D.coloff_off);
    const cplx* val = reinterpret_cast<const cplx*>(blob + D.val_off);
#pragma unroll
    for (int j = 0; j < kSlots; ++j) {
        const uint32_t l = e * kSlots + j;
        uint32_t t = 0;
        for (uint32_t b = 0; b < D.k; ++b) t = (t << 1) | ((l >> D.gpos[b]) & 1u);
        const uint32_t gb = l & ~D.gate_mask;
        cplx acc{0.0, 0.0};
        for (uint32_t p = rowptr[t]; p < rowptr[t + 1]; ++p) {
            const cplx v = val[p], x = tile[swz(gb | coloff[p])];
            acc.x += v.x * x.x - v.y * x.y;
            acc.y += v.x * x.y + v.y * x.x;
        }
        out[j] = acc;
    }
}

QSV_HD void dense_store(uint32_t e, cplx* tile, const cplx (&out)[kSlots]) {
#pragma unroll
    for (int j = 0; j < kSlots; ++j) tile[swz(e * kSlots + j)] = out[j];
}

}  // namespace qsv
