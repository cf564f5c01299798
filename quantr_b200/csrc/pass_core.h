// pass_core.h — per-thread body of a fused pass, written __host__ __device__.
//
// The sm_100a kernel in kernels.cu calls these functions from its 256 threads with the
// tile in shared memory.  tests/emu/ compiles the same functions with g++ and walks the
// threads sequentially so the host scheduler + blob encoding can be unit-tested without a
// GPU; that harness is test infrastructure and is never linked into libqsv.so.
//
// Replaces, for a fused list of gates: Circuit::apply_gate (src/circuit/simulation.rs:64-135),
// insert_gate_image_into_product_state (:158-180) and the gate columns of
// src/circuit/standard_gate_ops.rs:37-267 (as lowered by lower.cpp).
#pragma once
#include <math.h>

#include "qsv_types.h"

namespace qsv {

// sin(pi*x), cos(pi*x) with exact results at multiples of 1/2.
QSV_HD void sincospi_hd(double x, double* s, double* c) {
#ifdef __CUDA_ARCH__
    ::sincospi(x, s, c);
#else
    double r = fmod(x, 2.0);  // exact
    if (r < 0) r += 2.0;      // [0, 2)
    // quadrant q in 0..3 with r = q/2 + f, f in [-1/4, 1/4]
    const double q = floor(r * 2.0 + 0.5);
    const double f = r - q * 0.5;  // exact
    const double sf = sin(M_PI * f), cf = cos(M_PI * f);
    switch (((int)q) & 3) {
        case 0: *s = sf; *c = cf; break;
        case 1: *s = cf; *c = -sf; break;
        case 2: *s = -sf; *c = -cf; break;
        default: *s = -cf; *c = sf; break;
    }
    if (f == 0.0) {  // exact multiples of 1/2: kill the signed zeros' noise
        if (*s == 0.0) *s = 0.0;
        if (*c == 0.0) *c = 0.0;
    }
#endif
}

// External phase of a DIAG op for one tile: exp(i*pi*(theta0 + sum over bits outside the tile)).
QSV_HD cplx diag_ext_phase(const DevOp& op, const uint8_t* blob, uint64_t base_full) {
    double ang = op.m[0];
    const DiagExtTerm* terms = reinterpret_cast<const DiagExtTerm*>(blob + op.ext_off);
    for (uint32_t i = 0; i < op.n_ext; ++i)
        if ((base_full >> terms[i].bit) & 1ull) ang += terms[i].coef;
    cplx r;
    sincospi_hd(ang, &r.y, &r.x);
    return r;
}

// ---- 2x2 ops on register slot J of the 16 amplitudes a thread holds -------------------------

template <int J>
QSV_HD void mat_general(cplx (&a)[kSlots], const double* m, uint32_t cm) {
#pragma unroll
    for (int s0 = 0; s0 < kSlots; ++s0) {
        if ((s0 >> J) & 1) continue;
        if ((s0 & cm) != cm) continue;
        const int s1 = s0 | (1 << J);
        const cplx x = a[s0], y = a[s1];
        a[s0].x = m[0] * x.x - m[1] * x.y + m[2] * y.x - m[3] * y.y;
        a[s0].y = m[0] * x.y + m[1] * x.x + m[2] * y.y + m[3] * y.x;
        a[s1].x = m[4] * x.x - m[5] * x.y + m[6] * y.x - m[7] * y.y;
        a[s1].y = m[4] * x.y + m[5] * x.x + m[6] * y.y + m[7] * y.x;
    }
}

template <int J>
QSV_HD void mat_real(cplx (&a)[kSlots], const double* m, uint32_t cm) {
#pragma unroll
    for (int s0 = 0; s0 < kSlots; ++s0) {
        if ((s0 >> J) & 1) continue;
        if ((s0 & cm) != cm) continue;
        const int s1 = s0 | (1 << J);
        const cplx x = a[s0], y = a[s1];
        a[s0].x = m[0] * x.x + m[2] * y.x;
        a[s0].y = m[0] * x.y + m[2] * y.y;
        a[s1].x = m[4] * x.x + m[6] * y.x;
        a[s1].y = m[4] * x.y + m[6] * y.y;
    }
}

template <int J>
QSV_HD void mat_antidiag(cplx (&a)[kSlots], const double* m, uint32_t cm) {
#pragma unroll
    for (int s0 = 0; s0 < kSlots; ++s0) {
        if ((s0 >> J) & 1) continue;
        if ((s0 & cm) != cm) continue;
        const int s1 = s0 | (1 << J);
        const cplx x = a[s0], y = a[s1];
        a[s0].x = m[2] * y.x - m[3] * y.y;
        a[s0].y = m[2] * y.y + m[3] * y.x;
        a[s1].x = m[4] * x.x - m[5] * x.y;
        a[s1].y = m[4] * x.y + m[5] * x.x;
    }
}

template <int J>
QSV_HD void mat_xswap(cplx (&a)[kSlots], uint32_t cm) {
#pragma unroll
    for (int s0 = 0; s0 < kSlots; ++s0) {
        if ((s0 >> J) & 1) continue;
        if ((s0 & cm) != cm) continue;
        const int s1 = s0 | (1 << J);
        const cplx x = a[s0];
        a[s0] = a[s1];
        a[s1] = x;
    }
}

#define QSV_SLOT_SWITCH(FN, ...)           \
    switch (op.slot) {                     \
        case 0: FN<0>(__VA_ARGS__); break; \
        case 1: FN<1>(__VA_ARGS__); break; \
        case 2: FN<2>(__VA_ARGS__); break; \
        default: FN<3>(__VA_ARGS__); break;\
    }

QSV_HD void apply_diag(cplx (&a)[kSlots], const DevOp& op, const uint8_t* blob, uint32_t e, cplx w) {
    const cplx* tbl = reinterpret_cast<const cplx*>(blob + op.tbl_off);
    if (op.flags & DIAG_HAS_THR_LO) w = cmul(w, tbl[e & 31u]);
    if (op.flags & DIAG_HAS_THR_HI) w = cmul(w, tbl[32u + (e >> 5)]);
    const uint32_t cm = op.cmask_reg;
    if (op.flags & DIAG_HAS_REG) {
#pragma unroll
        for (int s = 0; s < kSlots; ++s)
            if ((s & cm) == cm) a[s] = cmul(a[s], cmul(w, tbl[64 + s]));
    } else {
#pragma unroll
        for (int s = 0; s < kSlots; ++s)
            if ((s & cm) == cm) a[s] = cmul(a[s], w);
    }
}

// One register round for thread-group index e (0 <= e < 2^(T-4)).
//   tile      : the tile in (swizzled) shared memory
//   ops       : the pass's op array
//   ext_phase : per-tile external phases of the pass's DIAG ops
//   base_full : physical index of the tile's first amplitude, rank bits included
QSV_HD void reg_round(const DevRound& R, const DevOp* ops, const uint8_t* blob, const cplx* ext_phase,
                      uint64_t base_full, uint32_t e, cplx* tile) {
    const uint32_t lb = (uint32_t)deposit(e, R.thr_segs, R.n_thr_segs);
    const uint32_t st0 = 1u << R.reg_pos[0], st1 = 1u << R.reg_pos[1], st2 = 1u << R.reg_pos[2], st3 = 1u << R.reg_pos[3];
    cplx a[kSlots];
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
        const uint32_t l = lb + ((s & 1) ? st0 : 0u) + ((s & 2) ? st1 : 0u) + ((s & 4) ? st2 : 0u) + ((s & 8) ? st3 : 0u);
        a[s] = tile[swz(l)];
    }
    for (uint32_t o = 0; o < R.n_ops; ++o) {
        const DevOp& op = ops[R.first_op + o];
        if ((base_full & op.cmask_ext) != op.cmask_ext) continue;  // uniform over the tile
        if ((lb & op.cmask_thr) != op.cmask_thr) continue;         // uniform over the thread's 16 amplitudes
        switch (op.type) {
            case OP_MAT_GENERAL: QSV_SLOT_SWITCH(mat_general, a, op.m, op.cmask_reg) break;
            case OP_MAT_REAL: QSV_SLOT_SWITCH(mat_real, a, op.m, op.cmask_reg) break;
            case OP_MAT_ANTIDIAG: QSV_SLOT_SWITCH(mat_antidiag, a, op.m, op.cmask_reg) break;
            case OP_MAT_XSWAP: QSV_SLOT_SWITCH(mat_xswap, a, op.cmask_reg) break;
            case OP_DIAG: apply_diag(a, op, blob, e, ext_phase[op.diag_index]); break;
            default: break;
        }
    }
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
        const uint32_t l = lb + ((s & 1) ? st0 : 0u) + ((s & 2) ? st1 : 0u) + ((s & 4) ? st2 : 0u) + ((s & 8) ? st3 : 0u);
        tile[swz(l)] = a[s];
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
    for (int j = 0; j < kSlots; ++j) tile[swz(e * kSlots + j)] = out[j];
}

}  // namespace qsv
