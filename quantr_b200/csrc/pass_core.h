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
QSV_HD cplx diag_ext_phase(const DevOp& op, const uint8_t* blob, uint64_t base_full) {
    double ang = op.m[0];
    const DiagExtTerm* terms = reinterpret_cast<const DiagExtTerm*>(blob + op.ext_off);
    for (uint32_t i = 0; i < op.n_ext; ++i)
        if ((base_full >> terms[i].bit) & 1ull) ang += terms[i].coef;
    cplx r;
    sincospi_hd(ang, &r.y, &r.x);
    return r;
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

template <int J, bool CTRL>
QSV_HD void mat_general(cplx (&a)[kSlots], const double (&m)[8], uint32_t cm) {
    const double ar = m[0], ai = m[1], br = m[2], bi = m[3], cr = m[4], ci = m[5], dr = m[6], di = m[7];
#pragma unroll
    for (int s0 = 0; s0 < kSlots; ++s0) {
        if ((s0 >> J) & 1) continue;
        if (CTRL && (s0 & cm) != cm) continue;
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
}

template <int J, bool CTRL>
QSV_HD void mat_real(cplx (&a)[kSlots], const double (&m)[8], uint32_t cm) {
    const double ar = m[0], br = m[2], cr = m[4], dr = m[6];
#pragma unroll
    for (int s0 = 0; s0 < kSlots; ++s0) {
        if ((s0 >> J) & 1) continue;
        if (CTRL && (s0 & cm) != cm) continue;
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
}

template <int J, bool CTRL>
QSV_HD void mat_hadamard(cplx (&a)[kSlots], const double (&)[8], uint32_t cm) {
#pragma unroll
    for (int s0 = 0; s0 < kSlots; ++s0) {
        if ((s0 >> J) & 1) continue;
        if (CTRL && (s0 & cm) != cm) continue;
        cplx& x = a[s0];
        cplx& y = a[s0 | (1 << J)];
        f_bfly(x.x, y.x);
        f_bfly(x.y, y.y);
    }
}

template <int J, bool CTRL>
QSV_HD void mat_antidiag(cplx (&a)[kSlots], const double (&m)[8], uint32_t cm) {
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
}

template <int J, bool CTRL>
QSV_HD void mat_xswap(cplx (&a)[kSlots], const double (&)[8], uint32_t cm) {
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
}

#define QSV_MAT_DISPATCH(FN)                                   \
    if (op.cmask_reg == 0) {                                   \
        switch (op.slot) {                                     \
            case 0: FN<0, false>(a, op.m, 0u); break;          \
            case 1: FN<1, false>(a, op.m, 0u); break;          \
            case 2: FN<2, false>(a, op.m, 0u); break;          \
            default: FN<3, false>(a, op.m, 0u); break;         \
        }                                                      \
    } else {                                                   \
        switch (op.slot) {                                     \
            case 0: FN<0, true>(a, op.m, op.cmask_reg); break; \
            case 1: FN<1, true>(a, op.m, op.cmask_reg); break; \
            case 2: FN<2, true>(a, op.m, op.cmask_reg); break; \
            default: FN<3, true>(a, op.m, op.cmask_reg); break;\
        }                                                      \
    }

// amp[s] *= f for the slots selected by the register-control mask; MASK_BIT < 0: all 16 slots,
// MASK_BIT = j: the 8 slots with register bit j set, MASK_BIT = 4: generic runtime mask.
template <int MASK_BIT, bool HAS_REG>
QSV_HD void diag_apply(cplx (&a)[kSlots], cplx w, const cplx* reg_tbl, uint32_t cm) {
    constexpr bool kSingle = MASK_BIT >= 0 && MASK_BIT < 4;
    constexpr int kShift = kSingle ? MASK_BIT : 0;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
        if (kSingle && !((s >> kShift) & 1)) continue;
        if (MASK_BIT == 4 && (s & cm) != cm) continue;
        if (HAS_REG) {
            const cplx f = cmul(w, reg_tbl[s]);
            c_mul_ip(a[s], f.x, f.y);
        } else {
            c_mul_ip(a[s], w.x, w.y);
        }
    }
}

QSV_HD void apply_diag(cplx (&a)[kSlots], const DevOp& op, const uint8_t* blob, uint32_t e, cplx w) {
    const cplx* tbl = reinterpret_cast<const cplx*>(blob + op.tbl_off);
    if (op.flags & DIAG_HAS_THR_LO) w = cmul(w, tbl[e & 31u]);
    if (op.flags & DIAG_HAS_THR_HI) w = cmul(w, tbl[32u + (e >> 5)]);
    const uint32_t cm = op.cmask_reg;
    const cplx* rt = tbl + 64;
    if (op.flags & DIAG_HAS_REG) {
        switch (cm) {
            case 0: diag_apply<-1, true>(a, w, rt, cm); break;
            case 1: diag_apply<0, true>(a, w, rt, cm); break;
            case 2: diag_apply<1, true>(a, w, rt, cm); break;
            case 4: diag_apply<2, true>(a, w, rt, cm); break;
            case 8: diag_apply<3, true>(a, w, rt, cm); break;
            default: diag_apply<4, true>(a, w, rt, cm); break;
        }
    } else {
        switch (cm) {
            case 0: diag_apply<-1, false>(a, w, rt, cm); break;
            case 1: diag_apply<0, false>(a, w, rt, cm); break;
            case 2: diag_apply<1, false>(a, w, rt, cm); break;
            case 4: diag_apply<2, false>(a, w, rt, cm); break;
            case 8: diag_apply<3, false>(a, w, rt, cm); break;
            default: diag_apply<4, false>(a, w, rt, cm); break;
        }
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
    const uint32_t sb = swz(lb) << 4;
    char* tb = reinterpret_cast<char*>(tile);
    cplx a[kSlots];
#pragma unroll
    for (int s = 0; s < kSlots; ++s) a[s] = *reinterpret_cast<const cplx*>(tb + (sb ^ R.xoff[s]));
    for (uint32_t o = 0; o < R.n_ops; ++o) {
        const DevOp& op = ops[R.first_op + o];
        if ((base_full & op.cmask_ext) != op.cmask_ext) continue;  // uniform over the tile
        if ((lb & op.cmask_thr) != op.cmask_thr) continue;         // uniform over the thread's 16 amplitudes
        switch (op.type) {
            case OP_MAT_HADAMARD: QSV_MAT_DISPATCH(mat_hadamard) break;
            case OP_DIAG: apply_diag(a, op, blob, e, ext_phase[op.diag_index]); break;
            case OP_MAT_XSWAP: QSV_MAT_DISPATCH(mat_xswap) break;
            case OP_MAT_REAL: QSV_MAT_DISPATCH(mat_real) break;
            case OP_MAT_GENERAL: QSV_MAT_DISPATCH(mat_general) break;
            case OP_MAT_ANTIDIAG: QSV_MAT_DISPATCH(mat_antidiag) break;
            default: break;
        }
    }
#pragma unroll
    for (int s = 0; s < kSlots; ++s) *reinterpret_cast<cplx*>(tb + (sb ^ R.xoff[s])) = a[s];
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
