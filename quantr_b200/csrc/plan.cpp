// plan.cpp — gate lowering, diagonal merging and the fusion scheduler (host, no CUDA).
//
// Reference behaviour being encoded:
//   gate walk / order             src/circuit/simulation.rs:37-56
//   gate -> positions dispatch    src/circuit/gate.rs:140-168, src/circuit/simulation.rs:75-112
//   gate columns (the matrices)   src/circuit/standard_gate_ops.rs:37-267
//   Custom None rule              src/circuit/simulation.rs:120-133
#include "plan.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <sstream>
#include <stdexcept>

#include "pass_core.h"

namespace qsv {

void host_sincospi(double x, double* s, double* c) { sincospi_hd(x, s, c); }

uint64_t LOp::support() const {
    uint64_t m = 0;
    if (kind == MAT) m = cmask | (1ull << target);
    else if (kind == DIAG) { m = cmask; for (auto& t : lin) m |= 1ull << t.first; }
    else for (int b : bits) m |= 1ull << b;
    return m;
}
uint64_t LOp::targets() const {
    uint64_t m = 0;
    if (kind == MAT) m = 1ull << target;
    else if (kind == DENSE) for (int b : bits) m |= 1ull << b;
    return m;
}

namespace {

[[noreturn]] void fail(const std::string& msg) { throw std::runtime_error(msg); }

LOp make_mat(uint32_t src, OpType t, int target, uint64_t cmask, double m00r, double m00i, double m01r, double m01i,
             double m10r, double m10i, double m11r, double m11i) {
    LOp o;
    o.kind = LOp::MAT; o.src_gate = src; o.mtype = t; o.target = target; o.cmask = cmask;
    const double mm[8] = {m00r, m00i, m01r, m01i, m10r, m10i, m11r, m11i};
    memcpy(o.m, mm, sizeof(mm));
    return o;
}
LOp make_xswap(uint32_t src, int target, uint64_t cmask) { return make_mat(src, OP_MAT_XSWAP, target, cmask, 0, 0, 1, 0, 1, 0, 0, 0); }
LOp make_diag(uint32_t src, uint64_t cmask, double theta0, std::vector<std::pair<int, double>> lin) {
    LOp o;
    o.kind = LOp::DIAG; o.src_gate = src; o.cmask = cmask; o.theta0 = theta0; o.lin = std::move(lin);
    return o;
}

// developer A/B switch: QSV_STRUCTURED_CUSTOM=0 sends every Custom gate through the dense round
bool plan_structured_custom() {
    const char* env = getenv("QSV_STRUCTURED_CUSTOM");
    return !env || atoi(env) != 0;
}

int expected_controls(uint32_t kind) {
    if (kind == QSV_GATE_TOFFOLI) return 2;
    if (kind >= QSV_GATE_CR && kind <= QSV_GATE_SWAP) return 1;
    if (kind == QSV_GATE_CUSTOM) return -1;
    return 0;
}

}  // namespace

void lower_gates(uint32_t n, const qsv_op* ops, size_t n_ops, std::vector<LOp>& out, uint64_t* n_gates) {
    if (n == 0 || n > 62) fail("n_qubits must be in 1..62");
    if (n_ops && !ops) fail("ops is NULL");
    uint64_t gates = 0;
    const double S2 = 0.70710678118654752440;  // FRAC_1_SQRT_2
    for (size_t g = 0; g < n_ops; ++g) {
        const qsv_op& op = ops[g];
        if (op.kind == QSV_GATE_ID) continue;  // simulation.rs:38-41
        std::ostringstream where;
        where << "op " << g << ": ";
        if (op.kind >= QSV_GATE_KIND_COUNT) fail(where.str() + "unknown gate kind");
        if (op.target >= n) fail(where.str() + "target wire out of range");
        const int want = expected_controls(op.kind);
        if (want >= 0 && (int)op.n_controls != want) fail(where.str() + "wrong number of control wires for this gate kind");
        if (op.n_controls && !op.controls) fail(where.str() + "controls is NULL");
        if (op.n_controls > 29) fail(where.str() + "too many control wires");
        for (uint32_t i = 0; i < op.n_controls; ++i) {
            if (op.controls[i] >= n) fail(where.str() + "control wire out of range");
            if (op.controls[i] == op.target) fail(where.str() + "control wire equals the gate's position");  // circuit.rs:272-293
            for (uint32_t j = 0; j < i; ++j)
                if (op.controls[i] == op.controls[j]) fail(where.str() + "overlapping control wires");
        }
        ++gates;
        const uint32_t src = (uint32_t)g;
        const int t = (int)(n - 1 - op.target);
        const int c0 = op.n_controls > 0 ? (int)(n - 1 - op.controls[0]) : -1;
        const int c1 = op.n_controls > 1 ? (int)(n - 1 - op.controls[1]) : -1;
        const uint64_t cm0 = c0 >= 0 ? (1ull << c0) : 0, cm1 = c1 >= 0 ? (1ull << c1) : 0;
        switch (op.kind) {
            case QSV_GATE_H: out.push_back(make_mat(src, OP_MAT_HADAMARD, t, 0, S2, 0, S2, 0, S2, 0, -S2, 0)); break;
            case QSV_GATE_X: out.push_back(make_xswap(src, t, 0)); break;
            case QSV_GATE_Y: out.push_back(make_mat(src, OP_MAT_ANTIDIAG, t, 0, 0, 0, 0, -1, 0, 1, 0, 0)); break;
            case QSV_GATE_Z: out.push_back(make_diag(src, 0, 0, {{t, 1.0}})); break;
            case QSV_GATE_S: out.push_back(make_diag(src, 0, 0, {{t, 0.5}})); break;
            case QSV_GATE_SDAG: out.push_back(make_diag(src, 0, 0, {{t, -0.5}})); break;
            case QSV_GATE_T: out.push_back(make_diag(src, 0, 0, {{t, 0.25}})); break;
            case QSV_GATE_TDAG: out.push_back(make_diag(src, 0, 0, {{t, -0.25}})); break;
            case QSV_GATE_RX: {
                const double c = cos(0.5 * op.param), s = sin(0.5 * op.param);
                out.push_back(make_mat(src, OP_MAT_GENERAL, t, 0, c, 0, 0, -s, 0, -s, c, 0));
                break;
            }
            case QSV_GATE_RY: {
                const double c = cos(0.5 * op.param), s = sin(0.5 * op.param);
                out.push_back(make_mat(src, OP_MAT_REAL, t, 0, c, 0, -s, 0, s, 0, c, 0));
                break;
            }
            case QSV_GATE_RZ: out.push_back(make_diag(src, 0, -0.5 * op.param / M_PI, {{t, op.param / M_PI}})); break;
            case QSV_GATE_PHASE: out.push_back(make_diag(src, 0, 0.5 * op.param / M_PI, {})); break;
            case QSV_GATE_X90: out.push_back(make_mat(src, OP_MAT_ANTIDIAG, t, 0, 0, 0, 0, -1, 0, -1, 0, 0)); break;
            case QSV_GATE_MX90: out.push_back(make_mat(src, OP_MAT_ANTIDIAG, t, 0, 0, 0, 0, 1, 0, 1, 0, 0)); break;
            case QSV_GATE_Y90: out.push_back(make_mat(src, OP_MAT_ANTIDIAG, t, 0, 0, 0, 1, 0, -1, 0, 0, 0)); break;
            case QSV_GATE_MY90: out.push_back(make_mat(src, OP_MAT_ANTIDIAG, t, 0, 0, 0, -1, 0, 1, 0, 0, 0)); break;
            case QSV_GATE_CR: out.push_back(make_diag(src, 1ull << t, 0, {{c0, op.param / M_PI}})); break;
            case QSV_GATE_CRK: out.push_back(make_diag(src, 1ull << t, 0, {{c0, ldexp(1.0, 1 - op.iparam)}})); break;
            case QSV_GATE_CZ: out.push_back(make_diag(src, 1ull << t, 0, {{c0, 1.0}})); break;
            case QSV_GATE_CY: out.push_back(make_mat(src, OP_MAT_ANTIDIAG, t, cm0, 0, 0, 0, -1, 0, 1, 0, 0)); break;
            case QSV_GATE_CNOT: out.push_back(make_xswap(src, t, cm0)); break;
            case QSV_GATE_SWAP:
                out.push_back(make_xswap(src, t, cm0));
                out.push_back(make_xswap(src, c0, 1ull << t));
                out.push_back(make_xswap(src, t, cm0));
                break;
            case QSV_GATE_TOFFOLI: out.push_back(make_xswap(src, t, cm0 | cm1)); break;
            case QSV_GATE_CUSTOM: {
                if (!op.matrix) fail(where.str() + "Custom gate without a matrix");
                const uint32_t k = op.n_controls + 1;
                const bool compact = op.iparam == 1;  // matrix = the columns of the sub-states the closure answered for (include/qsv.h)
                if (k > 30) fail(where.str() + "Custom gates on more than 30 wires are not supported");
                if (!compact && k > (uint32_t)kMaxTileBits) fail(where.str() + "Custom gates on more than 13 wires are not supported unless they act on one basis state or on one pair of them (pass compact columns, qsv.h)");
                const uint64_t dim = 1ull << k;
                if (compact && !op.none_mask) fail(where.str() + "compact Custom columns need a none_mask");
                std::vector<int> gbits;  // bits[0] = MSB of the sub-index = first control
                for (uint32_t i = 0; i < op.n_controls; ++i) gbits.push_back((int)(n - 1 - op.controls[i]));
                gbits.push_back(t);
                // The reference's rule (simulation.rs:120-133): sub-states the closure returned None for keep their own
                // amplitude (identity row, overwrite) and contribute nothing else; every other column is the closure's image
                // with its entries on None rows dropped.  Collect the rows that differ from the identity.
                struct Ent { uint64_t r, s; cplx v; };
                std::vector<Ent> ents;            // entries of the folded matrix outside its None rows
                std::vector<uint64_t> active;     // sub-states with a closure image
                if (compact) {
                    for (uint64_t s0 = 0; s0 < dim; ++s0) if (!op.none_mask[s0]) active.push_back(s0);
                    if (active.size() > 64) fail(where.str() + "compact Custom columns: at most 64 sub-states may have an image");
                } else {
                    for (uint64_t s0 = 0; s0 < dim; ++s0) if (!(op.none_mask && op.none_mask[s0])) active.push_back(s0);
                }
                for (size_t ci = 0; ci < active.size(); ++ci) {
                    const uint64_t s0 = active[ci];
                    const double* col = compact ? op.matrix + ci * dim * 2 : nullptr;
                    for (uint64_t r = 0; r < dim; ++r) {
                        if (op.none_mask && op.none_mask[r]) continue;
                        const cplx v = compact ? cplx{col[2 * r], col[2 * r + 1]} : cplx{op.matrix[(r * dim + s0) * 2], op.matrix[(r * dim + s0) * 2 + 1]};
                        if (v.x != 0.0 || v.y != 0.0) ents.push_back(Ent{r, s0, v});
                    }
                }
                // Structured gates: the folded matrix is the identity except on one basis sub-state (a multi-controlled
                // phase / scaling) or on one pair of sub-states that differ in a single wire (a multi-controlled 2x2 gate,
                // e.g. the reference's multicnot::<N>, tests/grovers.rs:157-172).  These become controlled ops of the fused
                // pass - no dense round, no limit of 13 wires.
                std::vector<uint64_t> moved;  // sub-states whose row or column is not the identity's
                auto note = [&](uint64_t x) { if (std::find(moved.begin(), moved.end(), x) == moved.end()) moved.push_back(x); };
                std::vector<char> has_diag(active.size(), 0);
                for (const Ent& e : ents) {
                    if (e.r == e.s && e.v.x == 1.0 && e.v.y == 0.0) { has_diag[std::find(active.begin(), active.end(), e.s) - active.begin()] = 1; continue; }
                    note(e.r);
                    note(e.s);
                    if (moved.size() > 2) break;
                }
                for (size_t ci = 0; ci < active.size() && moved.size() <= 2; ++ci) {
                    if (has_diag[ci]) continue;
                    bool unit = false;
                    for (const Ent& e : ents) unit |= (e.r == active[ci] && e.s == active[ci] && e.v.x == 1.0 && e.v.y == 0.0);
                    if (!unit) note(active[ci]);  // the column lost its diagonal 1 (e.g. maps to zero)
                }
                bool structured = moved.size() <= 2 && plan_structured_custom();
                if (structured && moved.size() == 2 && __builtin_popcountll(moved[0] ^ moved[1]) != 1) structured = false;
                if (structured && moved.empty()) break;  // the identity
                if (structured) {
                    std::sort(moved.begin(), moved.end());
                    auto entry = [&](uint64_t r, uint64_t s0) {
                        for (const Ent& e : ents) if (e.r == r && e.s == s0) return e.v;
                        return cplx{0.0, 0.0};
                    };
                    const uint64_t s_lo = moved[0];
                    const uint64_t diff = moved.size() == 2 ? (moved[0] ^ moved[1]) : 0;
                    // sub-index bit e (MSB first) <-> physical bit gbits[e]
                    auto phys_of_sub_bit = [&](uint64_t sub_bit_mask) { int e = 0; while (!((sub_bit_mask >> (k - 1 - e)) & 1ull)) ++e; return gbits[e]; };
                    uint64_t cmask = 0, flip = 0;  // controls (all other wires); wires whose control value is 0 are X-conjugated
                    for (uint32_t e = 0; e < k; ++e) {
                        const uint64_t sb = 1ull << (k - 1 - e);
                        if (sb == diff) continue;
                        cmask |= 1ull << gbits[e];
                        if (!(s_lo & sb)) flip |= 1ull << gbits[e];
                    }
                    for (int b = 0; b < 64; ++b) if ((flip >> b) & 1ull) out.push_back(make_xswap(src, b, 0));
                    if (moved.size() == 1) {
                        // one basis sub-state scaled by v: a multi-controlled 2x2 diag(1, v) on any one of its wires
                        const cplx v = entry(s_lo, s_lo);
                        const int tb = gbits[k - 1];
                        const uint64_t cm = cmask & ~(1ull << tb);
                        out.push_back(make_mat(src, OP_MAT_GENERAL, tb, cm, 1, 0, 0, 0, 0, 0, v.x, v.y));
                    } else {
                        const int tb = phys_of_sub_bit(diff);
                        const uint64_t s_hi = moved[1];  // target bit set
                        const cplx m00 = entry(s_lo, s_lo), m01 = entry(s_lo, s_hi), m10 = entry(s_hi, s_lo), m11 = entry(s_hi, s_hi);
                        const bool is_x = m00.x == 0 && m00.y == 0 && m11.x == 0 && m11.y == 0 && m01.x == 1 && m01.y == 0 && m10.x == 1 && m10.y == 0;
                        if (is_x) out.push_back(make_xswap(src, tb, cmask));
                        else out.push_back(make_mat(src, OP_MAT_GENERAL, tb, cmask, m00.x, m00.y, m01.x, m01.y, m10.x, m10.y, m11.x, m11.y));
                    }
                    for (int b = 63; b >= 0; --b) if ((flip >> b) & 1ull) out.push_back(make_xswap(src, b, 0));
                    break;
                }
                if (compact) fail(where.str() + "a Custom gate given as compact columns must act on one basis state or on one pair that differs in one wire");
                LOp o;
                o.kind = LOp::DENSE; o.src_gate = src;
                o.bits = gbits;
                o.rowptr.assign(dim + 1, 0);
                for (uint64_t r = 0; r < dim; ++r) {
                    const bool none_r = op.none_mask && op.none_mask[r];
                    for (uint64_t s = 0; s < dim; ++s) {
                        cplx v;
                        if (none_r) v = cplx{r == s ? 1.0 : 0.0, 0.0};  // untouched state keeps its own amplitude (overwrite)
                        else if (op.none_mask && op.none_mask[s]) v = cplx{0.0, 0.0};
                        else v = cplx{op.matrix[(r * dim + s) * 2], op.matrix[(r * dim + s) * 2 + 1]};
                        if (v.x == 0.0 && v.y == 0.0) continue;
                        o.cols.push_back((uint32_t)s);
                        o.vals.push_back(v);
                    }
                    o.rowptr[r + 1] = (uint32_t)o.cols.size();
                }
                out.push_back(std::move(o));
                break;
            }
            default: fail(where.str() + "unhandled gate kind");
        }
    }
    if (n_gates) *n_gates = gates;
}

// A diagonal op commutes with every other diagonal op and with any op that has no *target*
// in its support; pull it back to the nearest earlier diagonal with the same control mask.
void merge_diagonals(std::vector<LOp>& lops) {
    std::vector<char> dead(lops.size(), 0);
    for (size_t i = 0; i < lops.size(); ++i) {
        if (lops[i].kind != LOp::DIAG) continue;
        const uint64_t sup = lops[i].support();
        const size_t lo = i > 512 ? i - 512 : 0;
        for (size_t j = i; j-- > lo;) {
            if (dead[j]) continue;
            LOp& p = lops[j];
            if (p.kind == LOp::DIAG) {
                if (p.cmask != lops[i].cmask) continue;
                p.theta0 += lops[i].theta0;
                for (auto& t : lops[i].lin) {
                    bool found = false;
                    for (auto& q : p.lin) if (q.first == t.first) { q.second += t.second; found = true; break; }
                    if (!found) p.lin.push_back(t);
                }
                dead[i] = 1;
                break;
            }
            if (p.targets() & sup) break;
        }
    }
    std::vector<LOp> kept;
    kept.reserve(lops.size());
    for (size_t i = 0; i < lops.size(); ++i)
        if (!dead[i]) kept.push_back(std::move(lops[i]));
    lops.swap(kept);
}

// Peephole over commuting ops (SURVEY.md 8f "circuit-level optimiser").  Every 2x2 op is a pair of matrices: S, applied
// where all its control bits are 1, and U, applied everywhere else (U = 1 for a plain controlled gate, U = S for an
// uncontrolled one).  A 2x2 gate is multiplied into the nearest earlier 2x2 gate on the same target when every op
// between them commutes with it and their controls agree or one of them has none: H.H, X.X and the like vanish, runs
// of rotations become one matrix, and (merge_ctrl) a CNot/Toffoli absorbs the one-wire gates around its target,
//   U2 . X^c . U1  =  (c ? U2.X.U1 : U2.U1),
// a *dual* op: one 2x2 routine with a per-thread choice of constants instead of three ops, and no amplitude moves for
// the X.  One-wire phase gates on the target fold in the same way.
namespace {
void mul2(const double* x, const double* y, double* out) {  // out = x . y (y acts first), 2x2 complex, row-major (re, im)
    auto mul = [](const double* a, const double* b, double* o) { o[0] = a[0] * b[0] - a[1] * b[1]; o[1] = a[0] * b[1] + a[1] * b[0]; };
    double t0[2], t1[2];
    mul(x + 0, y + 0, t0); mul(x + 2, y + 4, t1); out[0] = t0[0] + t1[0]; out[1] = t0[1] + t1[1];
    mul(x + 0, y + 2, t0); mul(x + 2, y + 6, t1); out[2] = t0[0] + t1[0]; out[3] = t0[1] + t1[1];
    mul(x + 4, y + 0, t0); mul(x + 6, y + 4, t1); out[4] = t0[0] + t1[0]; out[5] = t0[1] + t1[1];
    mul(x + 4, y + 2, t0); mul(x + 6, y + 6, t1); out[6] = t0[0] + t1[0]; out[7] = t0[1] + t1[1];
}
// H.H and friends: products that are 0 or +-1 up to rounding (|x| < 2^-50) become exact
void snap(double* r) {
    for (int q = 0; q < 8; ++q) {
        if (fabs(r[q]) < 8.9e-16) r[q] = 0.0;
        if (fabs(r[q] - 1.0) < 8.9e-16) r[q] = 1.0;
        if (fabs(r[q] + 1.0) < 8.9e-16) r[q] = -1.0;
    }
}
const double kIdentity2[8] = {1, 0, 0, 0, 0, 0, 1, 0};
bool is_identity2(const double* m) { return memcmp(m, kIdentity2, sizeof(kIdentity2)) == 0 || (m[0] == 1.0 && m[1] == 0.0 && m[2] == 0.0 && m[3] == 0.0 && m[4] == 0.0 && m[5] == 0.0 && m[6] == 1.0 && m[7] == 0.0); }
bool is_real2(const double* m) { return m[1] == 0.0 && m[3] == 0.0 && m[5] == 0.0 && m[7] == 0.0; }
// picks the cheapest op type for a matrix; false if it is the identity
bool classify_mat(const double* m, OpType* t) {
    const bool off_zero = m[2] == 0.0 && m[3] == 0.0 && m[4] == 0.0 && m[5] == 0.0;
    const bool diag_zero = m[0] == 0.0 && m[1] == 0.0 && m[6] == 0.0 && m[7] == 0.0;
    if (off_zero && m[0] == 1.0 && m[1] == 0.0 && m[6] == 1.0 && m[7] == 0.0) return false;
    if (diag_zero && m[2] == 1.0 && m[3] == 0.0 && m[4] == 1.0 && m[5] == 0.0) *t = OP_MAT_XSWAP;
    else if (diag_zero) *t = OP_MAT_ANTIDIAG;
    else if (is_real2(m)) *t = OP_MAT_REAL;
    else *t = OP_MAT_GENERAL;
    return true;
}
void get_su(const LOp& o, double* s, double* u) {
    memcpy(s, o.m, sizeof(o.m));
    memcpy(u, o.dual ? o.m2 : (o.cmask ? kIdentity2 : o.m), sizeof(o.m));
}
// Stores the pair (S where the controls hold, U elsewhere) in its cheapest form; false if the op became the identity.
bool set_su(LOp& o, double* s, double* u) {
    snap(s);
    snap(u);
    if (o.cmask && memcmp(s, u, sizeof(o.m)) == 0) o.cmask = 0;  // the controls no longer matter
    o.dual = false;
    memcpy(o.m, s, sizeof(o.m));
    memcpy(o.m2, kIdentity2, sizeof(o.m2));
    if (o.cmask == 0 || is_identity2(u)) return classify_mat(o.m, &o.mtype);
    o.dual = true;
    memcpy(o.m2, u, sizeof(o.m2));
    o.mtype = (is_real2(s) && is_real2(u)) ? OP_MAT_REAL : OP_MAT_GENERAL;
    return true;
}
}  // namespace

void merge_single_qubit_gates(std::vector<LOp>& lops, bool merge_ctrl) {
    std::vector<char> dead(lops.size(), 0);
    // an uncontrolled one-wire phase gate (Rz, Z, S, T, ...) as the diagonal matrix it is
    auto diag_1q = [](const LOp& o, int* t, double* m) {
        if (o.kind != LOp::DIAG || o.cmask != 0 || o.lin.size() != 1) return false;
        *t = o.lin[0].first;
        double s0, c0, s1, c1;
        sincospi_hd(o.theta0, &s0, &c0);
        sincospi_hd(o.theta0 + o.lin[0].second, &s1, &c1);
        const double mm[8] = {c0, s0, 0, 0, 0, 0, c1, s1};
        memcpy(m, mm, sizeof(mm));
        return true;
    };
    for (size_t i = 0; i < lops.size(); ++i) {
        int dt = -1;
        double dm[8];
        const bool i_diag = diag_1q(lops[i], &dt, dm);
        if (lops[i].kind != LOp::MAT && !i_diag) continue;
        const size_t lo = i > 256 ? i - 256 : 0;
        for (size_t j = i; j-- > lo;) {
            if (dead[j]) continue;
            LOp& p = lops[j];
            const uint64_t tg = i_diag ? 0 : lops[i].targets(), sup = lops[i].support();
            const bool conflict = (p.targets() & sup) || (tg & p.support());
            if (!conflict) continue;
            int pt = -1;
            double pm[8], s[8], u[8], rs[8], ru[8];
            if (i_diag) {
                // a phase gate right after (in commutation order) a 2x2 gate on its wire: fold it into the matrices
                if (p.kind == LOp::MAT && p.target == dt && (p.cmask == 0 || merge_ctrl)) {
                    get_su(p, s, u);
                    mul2(dm, s, rs);
                    mul2(dm, u, ru);
                    dead[i] = 1;
                    if (!set_su(p, rs, ru)) dead[j] = 1;
                }
                break;
            }
            if ((lops[i].cmask == 0 || merge_ctrl) && diag_1q(p, &pt, pm) && pt == lops[i].target) {
                // a 2x2 gate right after a phase gate on its wire: absorb the phase gate and keep looking back
                get_su(lops[i], s, u);
                mul2(s, pm, rs);
                mul2(u, pm, ru);
                dead[j] = 1;
                if (!set_su(lops[i], rs, ru)) { dead[i] = 1; break; }
                continue;
            }
            if (p.kind == LOp::MAT && p.target == lops[i].target &&
                (p.cmask == lops[i].cmask || (merge_ctrl && (p.cmask == 0 || lops[i].cmask == 0)))) {
                // product = M_i * M_j (j acts first), for the controlled and the uncontrolled half
                double ps[8], pu[8];
                get_su(lops[i], s, u);
                get_su(p, ps, pu);
                mul2(s, ps, rs);
                mul2(u, pu, ru);
                p.cmask |= lops[i].cmask;
                dead[i] = 1;
                if (!set_su(p, rs, ru)) dead[j] = 1;
            }
            break;  // the nearest op that does not commute decides
        }
    }
    std::vector<LOp> kept;
    kept.reserve(lops.size());
    for (size_t i = 0; i < lops.size(); ++i)
        if (!dead[i]) kept.push_back(std::move(lops[i]));
    lops.swap(kept);
}

namespace {

struct RoundB {
    bool dense = false;
    bool perm = false;         // permutation round (ROUND_PERM): the first n_perm ops are X / CNot / Toffoli / multi-controlled X gates
                               // on any tile bits (the gather); ordinary ops on the round's register bits may follow
    size_t n_perm = 0;
    std::vector<int> reg;      // physical bits held in registers (targets first, then filler)
    std::vector<size_t> ops;   // indices into lops
};
struct PassB {
    uint64_t req = 0;          // physical bits that must be tile bits
    bool relaxed_low = false;  // a wide Custom gate took the low passenger bits
    bool big = false;          // all required bits lie in the contiguous low block: the pass uses the larger tile (kBigTileBits)
    std::vector<RoundB> rounds;
    size_t n_ops = 0;
    bool empty() const { return rounds.empty(); }
};

int popcnt(uint64_t v) { return __builtin_popcountll(v); }

std::vector<Seg> runs_to_segs(const std::vector<int>& phys_sorted) {
    std::vector<Seg> segs;
    size_t i = 0;
    while (i < phys_sorted.size()) {
        size_t j = i + 1;
        while (j < phys_sorted.size() && phys_sorted[j] == phys_sorted[j - 1] + 1) ++j;
        segs.push_back(Seg{(uint8_t)i, (uint8_t)(j - i), (uint8_t)phys_sorted[i], 0});
        i = j;
    }
    return segs;
}

// same for a list that is not ascending (thread index bit i -> tile-local bit list[i]): runs of consecutive positions
std::vector<Seg> runs_to_segs_ordered(const std::vector<int>& list) { return runs_to_segs(list); }

template <class T>
size_t append(std::vector<uint8_t>& blob, const T* data, size_t count) {
    while (blob.size() % 16) blob.push_back(0);
    const size_t off = blob.size();
    const uint8_t* p = reinterpret_cast<const uint8_t*>(data);
    blob.insert(blob.end(), p, p + sizeof(T) * count);
    return off;
}

cplx unit_phase(double half_turns) {
    cplx r;
    sincospi_hd(half_turns, &r.y, &r.x);
    return r;
}

constexpr int kBigTileBits = 12;  // tile of a pass over the contiguous low index bits (no short runs to pay for)
constexpr int kReorderWindow = 4096;  // ops a pass looks ahead for work that commutes with what it left behind

void emit_pass(Plan& plan, const PassB& pb, const std::vector<LOp>& lops) {
    const int T = pb.big ? kBigTileBits : std::min<int>(plan.opt.tile_bits, (int)plan.n_alloc);
    const int nloc = (int)plan.n_alloc;
    // tile bits = required bits + lowest free bits
    uint64_t tile_mask = pb.req;
    for (int b = 0; b < nloc && popcnt(tile_mask) < T; ++b) tile_mask |= 1ull << b;
    if (popcnt(tile_mask) != T) fail("internal: tile bit count");
    std::vector<int> tile_phys, ext_phys;
    for (int b = 0; b < nloc; ++b) ((tile_mask >> b) & 1 ? tile_phys : ext_phys).push_back(b);
    std::vector<int> local_of(64, -1);
    for (size_t i = 0; i < tile_phys.size(); ++i) local_of[tile_phys[i]] = (int)i;

    DevPass hdr;
    memset(&hdr, 0, sizeof(hdr));
    hdr.magic = kPassMagic;
    hdr.tile_bits = (uint32_t)T;
    hdr.n_tiles = 1ull << (nloc - T);
    auto tsegs = runs_to_segs(tile_phys), esegs = runs_to_segs(ext_phys);
    if (tsegs.size() > (size_t)kMaxSegs || esegs.size() > (size_t)kMaxSegs) fail("internal: too many tile segments");
    hdr.n_tile_segs = (uint32_t)tsegs.size();
    hdr.n_ext_segs = (uint32_t)esegs.size();
    for (size_t i = 0; i < tsegs.size(); ++i) hdr.tile_segs[i] = tsegs[i];
    for (size_t i = 0; i < esegs.size(); ++i) hdr.ext_segs[i] = esegs[i];

    DevLoads loads;
    memset(&loads, 0, sizeof(loads));
    const uint32_t threads = tile_threads((uint32_t)T);
    hdr.threads = threads;
    for (uint32_t i = 0; i < (uint32_t)kMaxLoads && (uint64_t)i * threads < (1ull << T); ++i) {
        loads.goff[i] = deposit((uint64_t)i * threads, hdr.tile_segs, hdr.n_tile_segs);
        loads.soff[i] = swz(i * threads) << 4;
    }
    uint32_t n_hadamard = 0;

    std::vector<DevRound> rounds;
    std::vector<DevOp> ops;
    std::vector<uint8_t> aux;  // appended after header/rounds/ops; offsets fixed up below
    struct Fix { size_t op; int field; };  // field: 0 ext_off, 1 tbl_off, 2 dense_off
    std::vector<Fix> fixes;
    std::vector<std::pair<size_t, std::vector<size_t>>> dense_fix;  // aux offset of DevDense -> needs internal fix
    uint32_t n_diag = 0, n_ext_ops = 0;

    // register bits of every round: the round's targets, padded with the highest free tile bits
    std::vector<std::vector<int>> round_regs;
    for (const RoundB& rb : pb.rounds) {
        std::vector<int> reg_local;
        if (!rb.dense) {
            for (int b : rb.reg) reg_local.push_back(local_of[b]);
            for (int lp = T - 1; lp >= 0 && (int)reg_local.size() < kRegBits; --lp)
                if (std::find(reg_local.begin(), reg_local.end(), lp) == reg_local.end()) reg_local.push_back(lp);
            std::sort(reg_local.begin(), reg_local.end());
        }
        round_regs.push_back(std::move(reg_local));
    }
    // Lane bits: thread bits 0-2 (the lanes of a quarter-warp) should address all eight 16-byte units of the swizzled
    // rows (qsv_types.h swz()): three non-register bits below 6 with distinct positions mod 3, lowest first.  When the
    // register bits leave tile bits 0-2 to the threads this is the identity order (and global stores coalesce).
    auto pick_lane_bits = [&](const std::vector<int>& reg_local, uint32_t excluded, int (&pick)[3]) {
        std::vector<int> low;
        for (int lp = 0; lp < T && lp < 6; ++lp)
            if (std::find(reg_local.begin(), reg_local.end(), lp) == reg_local.end() && !((excluded >> lp) & 1u)) low.push_back(lp);
        for (size_t x = 0; x < low.size(); ++x)
            for (size_t y = x + 1; y < low.size(); ++y)
                for (size_t z = y + 1; z < low.size(); ++z)
                    if (low[x] % 3 != low[y] % 3 && low[x] % 3 != low[z] % 3 && low[y] % 3 != low[z] % 3) {
                        pick[0] = low[x]; pick[1] = low[y]; pick[2] = low[z];
                        return true;
                    }
        return false;
    };
    size_t round_index = 0;
    for (const RoundB& rb : pb.rounds) {
        const std::vector<int>& reg_local_pre = round_regs[round_index++];
        DevRound dr;
        memset(&dr, 0, sizeof(dr));
        dr.first_op = (uint32_t)ops.size();
        if (rb.dense) {
            dr.type = ROUND_DENSE;
            dr.n_ops = 1;
            const LOp& lop = lops[rb.ops[0]];
            DevOp dop;
            memset(&dop, 0, sizeof(dop));
            dop.type = OP_DENSE;
            dop.ext_slot = kNoExtSlot;
            DevDense dd;
            memset(&dd, 0, sizeof(dd));
            dd.k = (uint32_t)lop.bits.size();
            for (size_t e = 0; e < lop.bits.size(); ++e) {
                const int lp = local_of[lop.bits[e]];
                if (lp < 0) fail("internal: dense bit outside tile");
                dd.gpos[e] = (uint8_t)lp;
                dd.gate_mask |= 1u << lp;
            }
            dd.nnz = (uint32_t)lop.cols.size();
            std::vector<uint32_t> coloff(lop.cols.size());
            for (size_t p = 0; p < lop.cols.size(); ++p) {
                uint32_t off = 0;
                for (uint32_t e = 0; e < dd.k; ++e)
                    if ((lop.cols[p] >> (dd.k - 1 - e)) & 1u) off |= 1u << dd.gpos[e];
                coloff[p] = off;
            }
            dd.rowptr_off = (uint32_t)append(aux, lop.rowptr.data(), lop.rowptr.size());
            dd.coloff_off = (uint32_t)append(aux, coloff.data(), coloff.size());
            dd.val_off = (uint32_t)append(aux, lop.vals.data(), lop.vals.size());
            const size_t dd_off = append(aux, &dd, 1);
            dense_fix.push_back({dd_off, {}});
            dop.dense_off = (uint32_t)dd_off;
            fixes.push_back({ops.size(), 2});
            ops.push_back(dop);
            rounds.push_back(dr);
            continue;
        }
        dr.type = rb.perm ? ROUND_PERM : ROUND_REG;
        // register bits: the round's targets, padded with the highest free tile bits
        const std::vector<int>& reg_local = reg_local_pre;
        std::vector<int> slot_of(16, -1), thr_local, thr_index(16, -1);
        for (int j = 0; j < kRegBits; ++j) { dr.reg_pos[j] = (uint8_t)reg_local[j]; slot_of[reg_local[j]] = j; }
        for (int lp = 0; lp < T; ++lp)
            if (slot_of[lp] < 0) thr_local.push_back(lp);
        {   // thread order: conflict-free lane bits first (pick_lane_bits)
            int pick[3] = {-1, -1, -1};
            if (pick_lane_bits(reg_local, 0u, pick)) {
                std::vector<int> ordered(pick, pick + 3);
                for (int lp : thr_local)
                    if (lp != pick[0] && lp != pick[1] && lp != pick[2]) ordered.push_back(lp);
                if (runs_to_segs_ordered(ordered).size() <= (size_t)kMaxThrSegs) thr_local.swap(ordered);
            }
        }
        for (size_t i = 0; i < thr_local.size(); ++i) thr_index[thr_local[i]] = (int)i;
        auto thsegs = runs_to_segs_ordered(thr_local);
        if (thsegs.size() > (size_t)kMaxThrSegs) fail("internal: too many thread segments");
        dr.n_thr_segs = (uint32_t)thsegs.size();
        for (size_t i = 0; i < thsegs.size(); ++i) dr.thr_segs[i] = thsegs[i];
        for (int sl = 0; sl < kSlots; ++sl) {
            uint32_t off = 0;
            for (int j = 0; j < kRegBits; ++j) if ((sl >> j) & 1) off |= 1u << reg_local[j];
            dr.xoff[sl] = swz(off) << 4;
        }

        auto split_cmask = [&](uint64_t cmask, DevOp& d) {
            for (int b = 0; b < 64; ++b) {
                if (!((cmask >> b) & 1)) continue;
                const int lp = b < nloc ? local_of[b] : -1;
                if (lp < 0) d.cmask_ext |= 1ull << b;
                else if (slot_of[lp] >= 0) d.cmask_reg |= 1u << slot_of[lp];
                else d.cmask_thr |= 1u << lp;
            }
        };
        // A whole round of a QFT / phase-estimation ladder (kCodeQft4): H + ladder on four physically adjacent register
        // bits, highest first, the ladders' register-bit phases being pi/2^(distance); the lowest stage may be a bare H.
        // Emitted as one macro-op followed - outside the round's op range - by the stages' DIAG ops, which keep supplying
        // the tile/thread factors.
        uint32_t qft4_diags = 0;
        bool qft4 = false;
        const int n_st = (int)rb.reg.size();  // stages = register bits that are targets (3: the fourth, highest, bit is a passenger)
        if (!rb.perm && plan.opt.fuse && plan.opt.qft4 && kRegBits == 4 && (n_st == 4 || n_st == 3) &&
            (rb.ops.size() == (size_t)(2 * n_st) || rb.ops.size() == (size_t)(2 * n_st - 1))) {
            bool ok = true;
            for (int j = 0; j + 1 < n_st; ++j) ok &= tile_phys[reg_local[j + 1]] == tile_phys[reg_local[j]] + 1;
            for (int st = 0; st < n_st && ok; ++st) {  // stage st works on slot n_st - 1 - st
                const int j = n_st - 1 - st;
                const LOp& h = lops[rb.ops[2 * st]];
                ok &= h.kind == LOp::MAT && h.mtype == OP_MAT_HADAMARD && h.cmask == 0 && slot_of[local_of[h.target]] == j;
                if (!ok || (size_t)(2 * st + 1) >= rb.ops.size()) break;  // bare H as the last stage
                const LOp& d = lops[rb.ops[2 * st + 1]];
                ok &= d.kind == LOp::DIAG && d.cmask == (1ull << h.target);
                if (!ok) break;
                double want[4] = {0, 0, 0, 0};
                for (int i = 0; i < j; ++i) want[i] = ldexp(1.0, -(j - i));
                double have[4] = {0, 0, 0, 0};
                for (auto& t : d.lin) {
                    const int lp = t.first < nloc ? local_of[t.first] : -1;
                    if (lp >= 0 && slot_of[lp] >= 0) have[slot_of[lp]] += t.second;
                }
                for (int i = 0; i < 4; ++i) ok &= have[i] == want[i];
            }
            if (ok) {
                qft4 = true;
                qft4_diags = (uint32_t)rb.ops.size() / 2;  // one per stage, or one less with a bare H at the end
                DevOp mo;
                memset(&mo, 0, sizeof(mo));
                mo.type = OP_QFT4;
                mo.code = kCodeQft4;
                mo.slot = qft4_diags;
                mo.cmask_reg = (uint32_t)n_st;
                mo.ext_slot = kNoExtSlot;
                ops.push_back(mo);
                n_hadamard += (uint32_t)n_st;
            }
        }
        if (rb.perm) { dr.perm_first = (uint32_t)ops.size(); dr.n_perm = (uint32_t)rb.n_perm; }
        for (size_t ri = 0; rb.perm && ri < rb.n_perm; ++ri) {  // the gather of a permutation round: targets and controls are tile-local positions
            const LOp& lop = lops[rb.ops[ri]];
            DevOp d;
            memset(&d, 0, sizeof(d));
            d.ext_slot = kNoExtSlot;
            d.type = OP_MAT_XSWAP;
            d.code = kCodeNop;
            const int lp = local_of[lop.target];
            if (lp < 0 || lop.kind != LOp::MAT || lop.mtype != OP_MAT_XSWAP || lop.dual) fail("internal: permutation round holds an op that is not an X gate inside the tile");
            d.slot = (uint32_t)lp;
            for (int b = 0; b < 64; ++b) {
                if (!((lop.cmask >> b) & 1)) continue;
                const int cl = b < nloc ? local_of[b] : -1;
                if (cl < 0) d.cmask_ext |= 1ull << b;
                else d.cmask_thr |= 1u << cl;
            }
            if (d.cmask_ext) hdr.ext_ctrl_mask[ops.size() >> 5] |= 1u << (ops.size() & 31);
            ops.push_back(d);
        }
        if (rb.perm) dr.first_op = (uint32_t)ops.size();  // the ordinary ops behind the gather
        for (size_t ri = rb.perm ? rb.n_perm : 0; ri < rb.ops.size(); ++ri) {
            size_t oi = rb.ops[ri];
            if (qft4 && lops[oi].kind == LOp::MAT) continue;  // the macro-op's Hadamards
            // Peephole: an uncontrolled Hadamard followed by a diagonal op controlled by exactly the Hadamard's bit
            // (one stage of a QFT-style ladder) becomes one fused op: emitted as the DIAG op with an HD dispatch code.
            int fused_hadamard_slot = -1;
            if (plan.opt.fuse && lops[oi].kind == LOp::MAT && lops[oi].mtype == OP_MAT_HADAMARD && lops[oi].cmask == 0 && ri + 1 < rb.ops.size()) {
                const LOp& nx = lops[rb.ops[ri + 1]];
                if (nx.kind == LOp::DIAG && nx.cmask == (1ull << lops[oi].target)) {
                    fused_hadamard_slot = slot_of[local_of[lops[oi].target]];
                    ++n_hadamard;
                    oi = rb.ops[++ri];
                }
            }
            const LOp& lop = lops[oi];
            DevOp d;
            memset(&d, 0, sizeof(d));
            d.ext_slot = kNoExtSlot;
            split_cmask(lop.cmask, d);
            if (lop.kind == LOp::MAT) {
                d.type = lop.mtype;
                const int lp = local_of[lop.target];
                if (lp < 0 || slot_of[lp] < 0) fail("internal: MAT target not a register bit");
                d.slot = (uint32_t)slot_of[lp];
                memcpy(d.m, lop.m, sizeof(lop.m));
                if (lop.dual) {  // second matrix: where the controls do not hold
                    memcpy(d.m + 8, lop.m2, sizeof(lop.m2));
                    d.flags |= MAT_DUAL;
                }
                if (lop.mtype == OP_MAT_HADAMARD) ++n_hadamard;
            } else {
                d.type = OP_DIAG;
                d.diag_index = n_diag++;
                d.theta0 = lop.theta0;
                double thr_coef[16] = {0}, reg_coef[4] = {0};
                std::vector<DiagExtTerm> ext;
                for (auto& t : lop.lin) {
                    if (t.second == 0.0) continue;
                    const int lp = t.first < nloc ? local_of[t.first] : -1;
                    if (lp < 0) ext.push_back(DiagExtTerm{(uint32_t)t.first, 0, t.second});
                    else if (slot_of[lp] >= 0) reg_coef[slot_of[lp]] += t.second;
                    else thr_coef[thr_index[lp]] += t.second;
                }
                std::vector<cplx> tbl(kDiagTblLen);
                bool has_lo = false, has_hi = false, has_reg = false;
                for (int i = 0; i < 32; ++i) {
                    double alo = 0;
                    for (int b = 0; b < 5; ++b) if ((i >> b) & 1) alo += thr_coef[b];
                    tbl[i] = unit_phase(alo);
                }
                for (int i = 0; i < 32; ++i) {
                    double ahi = 0;
                    for (int b = 0; b < 5; ++b) if ((i >> b) & 1) ahi += thr_coef[5 + b];
                    tbl[32 + i] = unit_phase(ahi);
                }
                for (int b = 0; b < 5; ++b) has_lo |= thr_coef[b] != 0.0;
                for (int b = 5; b < 10; ++b) has_hi |= thr_coef[b] != 0.0;
                // register-bit constants (see DevOp::m): one control bit among the register bits -> table over the
                // subsets of the other three; otherwise one phase per register bit
                uint32_t nontrivial = 0;
                const bool single_ctl = d.cmask_reg && (d.cmask_reg & (d.cmask_reg - 1)) == 0;
                if (single_ctl) {
                    int ctl = 0;
                    while (!((d.cmask_reg >> ctl) & 1)) ++ctl;
                    int free_bits[3], nf = 0;
                    for (int b = 0; b < 4; ++b) if (b != ctl) free_bits[nf++] = b;
                    for (int q = 1; q < 8; ++q) {
                        double ang = 0;
                        for (int k = 0; k < 3; ++k) if ((q >> k) & 1) ang += reg_coef[free_bits[k]];
                        const cplx r = unit_phase(ang);
                        d.m[2 * (q - 1)] = r.x;
                        d.m[2 * (q - 1) + 1] = r.y;
                        if (r.x != 1.0 || r.y != 0.0) nontrivial |= 1u << (q - 1);
                    }
                } else {
                    for (int j = 0; j < 4; ++j) {  // unused register slots (j >= kRegBits) stay at phase 1
                        const cplx r = unit_phase(reg_coef[j]);
                        d.m[2 * j] = r.x;
                        d.m[2 * j + 1] = r.y;
                        if (r.x != 1.0 || r.y != 0.0) nontrivial |= 1u << j;
                    }
                }
                has_reg = nontrivial != 0;
                const bool has_w = lop.theta0 != 0.0 || !ext.empty() || has_lo || has_hi;
                d.flags = (has_lo ? (uint32_t)DIAG_HAS_THR_LO : 0u) | (has_hi ? (uint32_t)DIAG_HAS_THR_HI : 0u) | (has_reg ? (uint32_t)DIAG_HAS_REG : 0u) |
                          (has_w ? (uint32_t)DIAG_HAS_W : 0u) | (nontrivial << DIAG_NONTRIVIAL_SHIFT);
                d.n_ext = (uint32_t)ext.size();
                d.ext_slot = ext.empty() ? kNoExtSlot : n_ext_ops++;
                if (!ext.empty()) { d.ext_off = (uint32_t)append(aux, ext.data(), ext.size()); fixes.push_back({ops.size(), 0}); }
                d.tbl_off = (uint32_t)append(aux, tbl.data(), tbl.size());  // always present: the kernel stages it in shared memory
                fixes.push_back({ops.size(), 1});
            }
            d.code = op_dispatch_code(d);
            if (fused_hadamard_slot >= 0) d.code = kCodeHdBase + ((d.flags & DIAG_HAS_REG) ? 4u : 0u) + (uint32_t)fused_hadamard_slot;
            if (qft4) d.code = kCodeNop;  // applied by the macro-op in front of it
            if (d.cmask_ext) hdr.ext_ctrl_mask[ops.size() >> 5] |= 1u << (ops.size() & 31);
            ops.push_back(d);
        }
        dr.n_ops = qft4 ? 1u : (uint32_t)ops.size() - dr.first_op;  // the macro-op's DIAG ops sit outside the range
        rounds.push_back(dr);
    }
    if (rounds.size() > (size_t)kMaxRounds || ops.size() > (size_t)kMaxOps) fail("internal: pass exceeds round/op caps");

    // Last round of the pass: if it is a register round whose register bits leave the three lowest tile bits to the
    // threads, its 16 amplitudes per thread go straight from registers to global memory (still 128-byte coalesced).
    if (!rounds.empty() && rounds.back().type != ROUND_DENSE && T >= 7 && plan.opt.direct_store) {
        const DevRound& lr = rounds.back();
        bool ok = true;
        for (int j = 0; j < kRegBits; ++j) ok &= lr.reg_pos[j] >= 3;
        if (ok) {
            hdr.flags |= PASS_DIRECT_STORE;
            for (int sl = 0; sl < kSlots; ++sl) {
                uint32_t off = 0;
                for (int j = 0; j < kRegBits; ++j) if ((sl >> j) & 1) off |= 1u << lr.reg_pos[j];
                loads.store_goff[sl] = deposit(off, hdr.tile_segs, hdr.n_tile_segs);
            }
        }
    }
    bool conditional = false;
    for (const DevOp& d : ops) conditional |= d.cmask_thr != 0 || d.cmask_ext != 0;
    for (const DevRound& r : rounds) conditional |= r.type != ROUND_REG;  // the fast path is built without dense and permutation rounds
    if (!conditional) hdr.flags |= PASS_UNCONDITIONAL;
    for (const DevOp& d : ops) if (d.type == OP_DIAG && d.n_ext > hdr.max_ext) hdr.max_ext = d.n_ext;
    hdr.n_rounds = (uint32_t)rounds.size();
    hdr.n_ops = (uint32_t)ops.size();
    hdr.n_diag = n_diag;
    hdr.n_ext_ops = n_ext_ops;
    hdr.final_scale = ldexp((n_hadamard & 1) ? 0.70710678118654752440 : 1.0, -(int)(n_hadamard / 2));
    hdr.rounds_off = (uint32_t)(sizeof(DevPass) + sizeof(DevLoads));
    hdr.ops_off = hdr.rounds_off + (uint32_t)(rounds.size() * sizeof(DevRound));
    const uint32_t aux_base = hdr.ops_off + (uint32_t)(ops.size() * sizeof(DevOp));
    for (auto& f : fixes) {
        DevOp& d = ops[f.op];
        if (f.field == 0) d.ext_off += aux_base;
        else if (f.field == 1) d.tbl_off += aux_base;
        else d.dense_off += aux_base;
    }
    for (auto& df : dense_fix) {
        DevDense* dd = reinterpret_cast<DevDense*>(aux.data() + df.first);
        dd->rowptr_off += aux_base;
        dd->coloff_off += aux_base;
        dd->val_off += aux_base;
    }
    while (aux.size() % 16) aux.push_back(0);
    hdr.blob_bytes = aux_base + (uint32_t)aux.size();

    std::vector<uint8_t> blob(hdr.blob_bytes);
    memcpy(blob.data(), &hdr, sizeof(hdr));
    memcpy(blob.data() + sizeof(hdr), &loads, sizeof(loads));
    if (!rounds.empty()) memcpy(blob.data() + hdr.rounds_off, rounds.data(), rounds.size() * sizeof(DevRound));
    if (!ops.empty()) memcpy(blob.data() + hdr.ops_off, ops.data(), ops.size() * sizeof(DevOp));
    if (!aux.empty()) memcpy(blob.data() + aux_base, aux.data(), aux.size());
    plan.passes.push_back(std::move(blob));
    plan.n_rounds += rounds.size();
}

}  // namespace

// Maps a lowered op from logical to physical bit space.
static LOp remap_lop(const LOp& in, const std::vector<uint8_t>& layout) {
    LOp o = in;
    auto mask = [&](uint64_t m) {
        uint64_t r = 0;
        for (int b = 0; b < 64; ++b) if ((m >> b) & 1) r |= 1ull << layout[b];
        return r;
    };
    o.cmask = mask(in.cmask);
    if (in.kind == LOp::MAT) o.target = layout[in.target];
    for (auto& t : o.lin) t.first = layout[t.first];
    for (auto& b : o.bits) b = layout[b];
    return o;
}

void build_plan(Plan& plan, uint32_t n_qubits, uint32_t n_local, const qsv_op* ops, size_t n_ops, const PlanOptions& opt_in,
                const uint8_t* initial_layout, bool free_layout) {
    if (n_local == 0 || n_local > n_qubits) fail("n_local_qubits must be in 1..n_qubits");
    // shards below 2^kMinQubits amplitudes are padded with idle index bits, which would collide with the rank bits
    if (n_local < n_qubits && n_local < (uint32_t)kMinQubits) fail("sharded registers need at least " + std::to_string(kMinQubits) + " qubits per rank");
    plan.n_qubits = n_qubits;
    plan.n_local = n_local;
    plan.n_alloc = std::max<uint32_t>(n_local, kMinQubits);
    plan.opt = opt_in;
    plan.passes.clear();
    plan.steps.clear();
    plan.n_rounds = 0;
    {
        static std::atomic<uint64_t> next_stamp{1};
        plan.stamp = next_stamp.fetch_add(1);
    }
    if (ops || n_ops == 0 || plan.lops.empty()) {  // (build_prefix_subplan hands over ready-made lowered ops instead of gates)
        plan.lops.clear();
        lower_gates(n_qubits, ops, n_ops, plan.lops, &plan.n_gates);
        if (plan.opt.fuse && plan.opt.merge_1q) merge_single_qubit_gates(plan.lops, plan.opt.merge_ctrl != 0);
        if (plan.opt.fuse) merge_diagonals(plan.lops);
    }
    // tile size: 11 for registers the pipelined kernel serves (measured on B200, DESIGN.md 6: four compute groups of
    // 128 threads overlap better than two of 256), else 12; widened when a Custom gate needs more tile bits
    int tile_bits_auto = plan.n_alloc >= 23 ? 11 : 12;
    for (const LOp& lop : plan.lops)
        if (lop.kind == LOp::DENSE) tile_bits_auto = std::max(tile_bits_auto, std::min<int>(popcnt(lop.targets()), kMaxTileBits));
    plan.opt.tile_bits = std::max<int>(kRegBits, std::min<int>(opt_in.tile_bits > 0 ? opt_in.tile_bits : tile_bits_auto, kMaxTileBits));

    const int nloc = (int)plan.n_alloc;
    const int T = std::min<int>(plan.opt.tile_bits, nloc);
    const int n = (int)n_qubits, g = n - (int)n_local;
    for (const LOp& lop : plan.lops)
        if (lop.kind == LOp::DENSE && popcnt(lop.targets()) > T) fail("Custom gate is wider than the tile (" + std::to_string(T) + " bits)");

    // ---- layout: physical position of every logical index bit ------------------------------------------------
    size_t n_lops = plan.lops.size();
    auto next_target_use = [&](int bit, size_t from) {  // first op >= from that needs `bit` as a tile bit
        for (size_t i = from; i < n_lops; ++i)
            if ((plan.lops[i].targets() >> bit) & 1) return i;
        return n_lops;
    };
    std::vector<uint8_t> layout(64);
    for (int b = 0; b < 64; ++b) layout[b] = (uint8_t)b;
    plan.free_initial_layout = false;
    if (initial_layout) {
        for (int b = 0; b < n; ++b) layout[b] = initial_layout[b];
    } else if (free_layout && g > 0) {
        // the g logical bits whose first use as a target comes last go to the rank id; the rest keep their order
        std::vector<std::pair<size_t, int>> first_use;
        for (int b = 0; b < n; ++b) first_use.push_back({next_target_use(b, 0), b});
        std::stable_sort(first_use.begin(), first_use.end(), [](const std::pair<size_t, int>& x, const std::pair<size_t, int>& y) { return x.first > y.first; });
        std::vector<int> global_bits;
        for (int j = 0; j < g; ++j) global_bits.push_back(first_use[j].second);
        std::sort(global_bits.begin(), global_bits.end());
        int next_local = 0;
        for (int b = 0; b < n; ++b) {
            auto it = std::find(global_bits.begin(), global_bits.end(), b);
            if (it == global_bits.end()) layout[b] = (uint8_t)next_local++;
            else layout[b] = (uint8_t)(n_local + (it - global_bits.begin()));
        }
        plan.free_initial_layout = true;
    }
    {  // validate: a permutation of 0..n-1
        uint64_t seen = 0;
        for (int b = 0; b < n; ++b) {
            if (layout[b] >= n || ((seen >> layout[b]) & 1)) fail("initial layout is not a permutation of the index bits");
            seen |= 1ull << layout[b];
        }
    }
    // Prefix folding (sharded basis states): with the canonical layout the rank id holds the top g logical bits (wires
    // 0..g-1, the first ones a QFT touches).  Leading ops whose targets are all rank bits - diagonal ops included - see a
    // product state, so the host applies them to the 2^g rank amplitudes at run time; if that prefix contains a
    // non-diagonal gate the canonical layout is kept and the prefix leaves the schedule.  QFT-n on 2^g ranks then needs
    // no remap at all (its first g stages are the prefix, every later stage is local).
    // The same holds for the top LOCAL qubits (large registers only, PlanOptions::prefix_min_local): after leading ops on
    // the top g + k index bits the state has 2^(g+k) non-zero amplitudes, which the host computes and the first pass
    // synthesises instead of reading the register (k <= kMaxPrefixLocalBits).  QFT-33 from a basis state: 14 stages folded,
    // the remaining 19 are one write-only pass (8 stages) and one pass over the contiguous low bits (11 stages).
    plan.prefix.clear();
    plan.prefix_local_bits = 0;
    if (free_layout && initial_layout == nullptr && plan.opt.fold_prefix && plan.opt.fuse) {
        int k_max = 0;
        // up to kHostPrefixBits the host computes the table; beyond, the prefix runs on a sub-register of g + k qubits on the
        // device (build_prefix_subplan, state_api.cu): up to all but the lowest eight qubits
        if ((int)n_local >= plan.opt.prefix_min_local && (int)n_local == nloc)
            k_max = std::min<int>(plan.opt.prefix_subregister ? kMaxPrefixLocalBits : kHostPrefixBits, (int)n_local - plan.opt.prefix_keep_bits);  // (default 16: at most 1/32 of the first pass's 2^11-amplitude tiles hold amplitudes)
        if (k_max < 0) k_max = 0;
        if (g > 0 || k_max > 0) {
            const uint64_t support = (((1ull << (g + k_max)) - 1ull) << ((int)n_local - k_max));
            size_t len = 0;
            bool has_mat = false;
            while (len < plan.lops.size()) {
                const LOp& lop = plan.lops[len];
                if (lop.kind == LOp::DENSE || (lop.targets() & ~support)) break;
                has_mat |= lop.kind == LOp::MAT;
                ++len;
            }
            while (len > 0 && plan.lops[len - 1].kind == LOp::DIAG) --len;  // trailing diagonals stay in the schedule (they fuse for free)
            if (has_mat && len > 0) {
                uint64_t touched = 0;
                for (size_t i = 0; i < len; ++i) touched |= plan.lops[i].targets();
                int lowest = n;  // lowest targeted bit: the support is the contiguous block of top bits down to it
                for (int b = 0; b < n; ++b)
                    if ((touched >> b) & 1) { lowest = b; break; }
                plan.prefix_local_bits = lowest < (int)n_local ? (uint32_t)((int)n_local - lowest) : 0u;
                plan.prefix.assign(plan.lops.begin(), plan.lops.begin() + (long)len);
                plan.lops.erase(plan.lops.begin(), plan.lops.begin() + (long)len);
                for (int b = 0; b < 64; ++b) layout[b] = (uint8_t)b;  // canonical layout
                plan.free_initial_layout = g > 0;
            }
        }
    }
    plan.initial_layout.assign(layout.begin(), layout.begin() + n);
    n_lops = plan.lops.size();

    // Greedy in-order grouping of (physical-space) ops into passes for a given number of low passenger bits.
    // L_first: passenger bits of the segment's first pass (it may differ: on a basis state the plan's first pass is
    // write-only, its run length does not matter, so it can hold one more target)
    auto schedule = [&](const std::vector<LOp>& lops, int L_first, int L, std::vector<PassB>& out) {
        out.clear();
        uint64_t low_mask = (T == nloc) ? 0 : ((1ull << L_first) - 1);  // single-tile states: every bit is a tile bit
        const bool allow_big = plan.opt.fuse && T == 11 && nloc >= 23 && plan.opt.big_low_pass;
        PassB cur;
        // Rounds of a finished pass, re-packed with the same commutation rule one level down: a register round takes, in
        // order, every op of the pass whose target is one of its (at most four) register bits and that commutes with the
        // ops it left for later rounds.  Every round is a trip of the tile through shared memory, so fewer, fuller rounds.
        auto repack_rounds = [&](PassB& pb) {
            std::vector<size_t> order;
            for (const RoundB& r : pb.rounds) order.insert(order.end(), r.ops.begin(), r.ops.end());
            std::sort(order.begin(), order.end());
            std::vector<char> used(order.size(), 0);
            std::vector<RoundB> packed;
            size_t left = order.size(), first = 0;
            while (left) {
                while (used[first]) ++first;
                RoundB r;
                if (lops[order[first]].kind == LOp::DENSE) {
                    r.dense = true;
                    r.ops.push_back(order[first]);
                    used[first] = 1;
                    --left;
                    packed.push_back(std::move(r));
                    continue;
                }
                uint64_t blocked_t = 0, blocked_s = 0;
                for (size_t k = first; k < order.size(); ++k) {
                    if (used[k]) continue;
                    const LOp& lop = lops[order[k]];
                    const uint64_t tg = lop.targets(), sup = lop.support();
                    bool ok = lop.kind != LOp::DENSE && !(tg & blocked_s) && !(sup & blocked_t) && r.ops.size() < (size_t)kMaxRoundOps;
                    if (ok && lop.kind == LOp::MAT && std::find(r.reg.begin(), r.reg.end(), lop.target) == r.reg.end()) {
                        if ((int)r.reg.size() == kRegBits) ok = false;
                        else r.reg.push_back(lop.target);
                    }
                    if (ok) {
                        r.ops.push_back(order[k]);
                        used[k] = 1;
                        --left;
                    } else {
                        blocked_t |= tg;
                        blocked_s |= sup;
                    }
                }
                packed.push_back(std::move(r));
            }
            if (packed.size() < pb.rounds.size()) pb.rounds.swap(packed);
        };
        // Permutation rounds: a run of X / CNot / Toffoli / multi-controlled X ops on more than four targets (it would take
        // several register rounds, each a trip through shared memory plus 96 register moves per op) becomes one gather
        // through the tile (ROUND_PERM).  The pass's ops keep their order; everything else is packed into register rounds
        // in that order.
        auto form_perm_rounds = [&](PassB& pb) {
            auto permable = [&](size_t i) { return lops[i].kind == LOp::MAT && lops[i].mtype == OP_MAT_XSWAP && !lops[i].dual; };
            std::vector<size_t> order;
            for (const RoundB& r : pb.rounds) order.insert(order.end(), r.ops.begin(), r.ops.end());
            // maximal runs of X ops; a run qualifies when it has more than kRegBits distinct targets
            std::vector<char> in_perm(order.size(), 0);
            bool any = false;
            for (size_t a = 0; a < order.size();) {
                if (!permable(order[a])) { ++a; continue; }
                size_t b = a;
                uint64_t tg = 0;
                while (b < order.size() && permable(order[b])) tg |= lops[order[b++]].targets();
                if (popcnt(tg) > kRegBits) {
                    for (size_t k = a; k < b; ++k) in_perm[k] = 1;
                    any = true;
                }
                a = b;
            }
            if (!any) return;
            std::vector<RoundB> out_rounds;
            for (size_t k = 0; k < order.size(); ++k) {
                const LOp& lop = lops[order[k]];
                if (in_perm[k]) {
                    if (out_rounds.empty() || !out_rounds.back().perm || out_rounds.back().n_perm != out_rounds.back().ops.size() ||
                        out_rounds.back().ops.size() >= (size_t)kMaxRoundOps) {
                        out_rounds.push_back(RoundB());
                        out_rounds.back().perm = true;
                    }
                    out_rounds.back().ops.push_back(order[k]);
                    out_rounds.back().n_perm = out_rounds.back().ops.size();
                } else if (lop.kind == LOp::DENSE) {
                    out_rounds.push_back(RoundB());
                    out_rounds.back().dense = true;
                    out_rounds.back().ops.push_back(order[k]);
                } else {
                    // (ordinary ops may follow the gather of a permutation round: they work on what it left in the registers)
                    bool fresh = out_rounds.empty() || out_rounds.back().dense || out_rounds.back().ops.size() - out_rounds.back().n_perm >= (size_t)kMaxRoundOps;
                    if (!fresh && lop.kind == LOp::MAT) {
                        RoundB& r = out_rounds.back();
                        if (std::find(r.reg.begin(), r.reg.end(), lop.target) == r.reg.end() && (int)r.reg.size() == kRegBits) fresh = true;
                    }
                    if (fresh) out_rounds.push_back(RoundB());
                    RoundB& r = out_rounds.back();
                    if (lop.kind == LOp::MAT && std::find(r.reg.begin(), r.reg.end(), lop.target) == r.reg.end()) r.reg.push_back(lop.target);
                    r.ops.push_back(order[k]);
                }
            }
            if (out_rounds.size() < pb.rounds.size() && out_rounds.size() <= (size_t)kMaxRounds) pb.rounds.swap(out_rounds);
        };
        auto close_pass = [&]() {
            if (!cur.empty()) {
                if (plan.opt.fuse && plan.opt.reorder && cur.rounds.size() > 2) repack_rounds(cur);
                if (plan.opt.fuse && plan.opt.perm_rounds && cur.rounds.size() > 1) form_perm_rounds(cur);
                out.push_back(std::move(cur));
                low_mask = (T == nloc) ? 0 : ((1ull << L) - 1);
            }
            cur = PassB();
        };
        // Commutation-aware greedy: a pass takes, in circuit order, every op that fits its tile and commutes with all
        // the ops it had to leave behind (an op left behind blocks the qubits it acts on: later ops commute with it iff
        // they share none of its targets and it shares none of theirs - controls and diagonal phases on common qubits
        // are fine).  Ops left behind start the next pass.  With reorder off (or fuse off) this is the plain in-order
        // grouping: the first op that does not fit closes the pass.
        const size_t n_all = lops.size();
        std::vector<char> done(n_all, 0);
        std::vector<uint64_t> tgs(n_all), sups(n_all);
        for (size_t i = 0; i < n_all; ++i) { tgs[i] = lops[i].targets(); sups[i] = lops[i].support(); }
        const bool reorder = plan.opt.fuse && plan.opt.reorder;
        const uint64_t all_bits = nloc >= 64 ? ~0ull : ((1ull << n) - 1ull);
        size_t first_undone = 0;
        auto take = [&](size_t i, bool relaxed, bool now_big) {
            const LOp& lop = lops[i];
            cur.req |= tgs[i];
            cur.big |= now_big;
            cur.relaxed_low |= relaxed;
            if (lop.kind == LOp::DENSE) {
                RoundB r;
                r.dense = true;
                r.ops.push_back(i);
                cur.rounds.push_back(std::move(r));
            } else {
                if (cur.rounds.empty() || cur.rounds.back().dense) cur.rounds.push_back(RoundB());
                if (cur.rounds.back().ops.size() >= (size_t)kMaxRoundOps) cur.rounds.push_back(RoundB());
                if (lop.kind == LOp::MAT) {
                    RoundB* r = &cur.rounds.back();
                    if (std::find(r->reg.begin(), r->reg.end(), lop.target) == r->reg.end()) {
                        if ((int)r->reg.size() == kRegBits) { cur.rounds.push_back(RoundB()); r = &cur.rounds.back(); }
                        r->reg.push_back(lop.target);
                    }
                    r->ops.push_back(i);
                } else {
                    cur.rounds.back().ops.push_back(i);
                }
            }
            cur.n_ops++;
            done[i] = 1;
        };
        while (first_undone < n_all) {
            uint64_t blocked_t = 0, blocked_s = 0;  // targets / supports of the ops left behind by the open pass
            size_t scanned = 0;
            for (size_t i = first_undone; i < n_all; ++i) {
                if (done[i]) continue;
                if (++scanned > (size_t)kReorderWindow) break;
                const uint64_t tg = tgs[i];
                const bool commutes = !(tg & blocked_s) && !(sups[i] & blocked_t);
                // Can the op join the open pass?  Its targets must fit next to the pass's tile bits and the low
                // passenger bits, and the pass must stay within the kernel's round/op caps.
                const bool caps_ok = cur.n_ops + 1 <= (size_t)kMaxOps && cur.rounds.size() + 2 <= (size_t)kMaxRounds;
                const bool fits = !cur.big && popcnt(cur.req | tg | low_mask) <= T;
                // a pass whose targets all lie in the lowest kBigTileBits index bits streams contiguous 64 KiB tiles
                // whatever its size, so it may hold more targets than the strided passes (large registers only)
                const bool fits_big = allow_big && ((cur.req | tg) >> kBigTileBits) == 0;
                const bool can_join = plan.opt.fuse && commutes && !cur.relaxed_low && (fits || fits_big) && caps_ok;
                if (cur.empty() && commutes) {
                    // the first op of a pass is always taken; a Custom gate wider than T - low_bits gets a pass of its
                    // own without passenger bits
                    const bool now_big = allow_big && (tg >> kBigTileBits) == 0 && popcnt(tg | low_mask) > T;
                    const bool relaxed = !now_big && popcnt(tg | low_mask) > T;
                    take(i, relaxed, now_big);
                    if (relaxed || !plan.opt.fuse) break;  // fuse off: one gate per pass
                    continue;
                }
                if (can_join) {
                    const bool now_big = allow_big && ((cur.req | tg) >> kBigTileBits) == 0 && popcnt(cur.req | tg | low_mask) > T;
                    take(i, false, now_big);
                    continue;
                }
                if (!reorder) break;  // in-order grouping: the pass ends at the first op it cannot take
                blocked_t |= tg;
                blocked_s |= sups[i];
                if ((blocked_s & all_bits) == all_bits && (blocked_t & all_bits) == all_bits) break;  // nothing later can commute
            }
            close_pass();
            while (first_undone < n_all && done[first_undone]) ++first_undone;
        }
        close_pass();
    };

    // Cost model, in units of one pass streamed at the HBM roofline.  Memory term: measured streaming efficiency of a
    // tile whose contiguous runs are 16 B << run_bits; compute term: measured per-op and per-round costs of the pass
    // kernel (tools/stream_probe.py, tools/diag_probe.py on B200, n = 30).  The two overlap only partly.
    // write_only_first: the first pass of the plan runs on a basis state that was never written to HBM (fused
    // initialisation): zeros are streamed contiguously whatever the tile shape, one tile is computed
    auto plan_cost = [&](const std::vector<LOp>& lops, const std::vector<PassB>& passes, bool write_only_first) {
        double total = 0;
        bool first = true;
        for (const PassB& pb : passes) {
            if (first && write_only_first) {
                first = false;
                total += 0.55;
                continue;
            }
            first = false;
            const int Tp = pb.big ? kBigTileBits : T;
            uint64_t tile_mask = pb.req;
            for (int b = 0; b < nloc && popcnt(tile_mask) < Tp; ++b) tile_mask |= 1ull << b;
            int run_bits = 0;
            while (run_bits < nloc && ((tile_mask >> run_bits) & 1)) ++run_bits;
            // measured with the pipelined TMA kernel (profiles/r02_stream_probe.txt, r02_qft33_tile_configs.txt): tiles whose
            // rows are one 128-byte line stream at ~60 % of the copy peak, two lines at ~88 %
            const double eff = run_bits <= 3 ? 0.60 : run_bits == 4 ? 0.88 : run_bits == 5 ? 0.90 : 0.93;
            const double mem = 1.0 / eff;
            double compute = 0.6 + 0.12 * (pb.rounds.empty() ? 0.0 : (double)pb.rounds.size() - 1.0);
            for (const RoundB& r : pb.rounds)
                for (size_t oi : r.ops) {
                    const LOp& lop = lops[oi];
                    if (lop.kind == LOp::DENSE) compute += 0.5;
                    else if (lop.kind == LOp::DIAG) compute += 0.035;
                    else compute += lop.mtype == OP_MAT_GENERAL ? 0.09 : lop.mtype == OP_MAT_HADAMARD ? 0.035 : 0.05;
                }
            // the 2^12-tile kernel keeps one tile in flight per SM (three 64 KiB buffers, two compute groups): measured
            // ~1.3x slower than the 2^11-tile kernel on the same work
            total += (std::max(mem, compute) + 0.1 * std::min(mem, compute)) * (Tp == 12 && T == 11 ? 1.3 : 1.0);
        }
        return total;
    };

    // Schedules one segment (ops whose targets are all local under the current layout) and emits its passes.
    int chosen_L = plan.opt.low_bits;
    auto emit_segment = [&](size_t i0, size_t i1) {
        if (i0 == i1) return;
        std::vector<LOp> seg;
        seg.reserve(i1 - i0);
        for (size_t i = i0; i < i1; ++i) seg.push_back(remap_lop(plan.lops[i], layout));
        std::vector<PassB> best;
        const bool write_only_first = free_layout && i0 == 0 && plan.passes.empty() && nloc >= 23;  // fused initialisation will apply
        if (plan.opt.low_bits > 0 || T == nloc || !plan.opt.fuse) {
            const int L = std::max(0, std::min<int>(plan.opt.low_bits > 0 ? plan.opt.low_bits : 3, T - 1));
            chosen_L = L;
            schedule(seg, L, L, best);
        } else {
            // low_bits = 0: pick the number of passenger bits that minimises the modelled cost (longer runs stream
            // better, fewer passenger bits fuse more gates per pass); ties go to the longer runs.  A write-only first
            // pass chooses its own.
            double best_cost = 0;
            std::vector<PassB> cand;
            for (int L = std::min(8, T - 1); L >= 3; --L)
                for (int Lf = write_only_first ? 3 : L; Lf <= (write_only_first ? std::min(8, T - 1) : L); ++Lf) {
                    schedule(seg, Lf, L, cand);
                    const double c = plan_cost(seg, cand, write_only_first);
                    if (best.empty() || c < best_cost - 1e-9) { best_cost = c; chosen_L = L; best.swap(cand); }
                }
        }
        for (const PassB& pb : best) {
            emit_pass(plan, pb, seg);
            PlanStep st;
            st.kind = PlanStep::PASS;
            st.pass_index = (uint32_t)plan.passes.size() - 1;
            plan.steps.push_back(std::move(st));
        }
    };

    size_t seg_start = 0;
    for (size_t i = 0; i < n_lops; ++i) {
        bool needs_global = false;
        const uint64_t tg = plan.lops[i].targets();
        for (int b = 0; b < n; ++b)
            if (((tg >> b) & 1) && layout[b] >= n_local) needs_global = true;
        if (!needs_global) continue;
        if (g == 0) fail("internal: global target without ranks");
        emit_segment(seg_start, i);
        seg_start = i;
        // Global-qubit remap: all g rank bits are swapped with the g local physical bits whose logical bits are
        // needed as targets furthest in the future (Belady), preferring high positions so the exchanged chunks are large.
        std::vector<int> logical_at(n, -1);
        for (int b = 0; b < n; ++b) logical_at[layout[b]] = b;
        // candidates: the top 10 local bits (large exchanged chunks); widened downwards if those do not suffice (wide
        // Custom gates, tiny shards)
        std::vector<std::pair<size_t, int>> cand;  // (next use, physical position)
        for (int p = (int)n_local - 1; p >= 0 && (p >= (int)n_local - 10 || (int)cand.size() < g); --p) {
            const int lb = logical_at[p];
            if ((tg >> lb) & 1) continue;  // needed right now
            cand.push_back({next_target_use(lb, i), p});
        }
        if ((int)cand.size() < g) fail("gate " + std::to_string(plan.lops[i].src_gate) + ": cannot bring its qubits onto one rank");
        std::stable_sort(cand.begin(), cand.end(), [](const std::pair<size_t, int>& x, const std::pair<size_t, int>& y) {
            return x.first != y.first ? x.first > y.first : x.second > y.second;
        });
        std::vector<int> partners;
        for (int j = 0; j < g; ++j) partners.push_back(cand[j].second);
        std::sort(partners.begin(), partners.end());
        PlanStep st;
        st.kind = PlanStep::EXCHANGE;
        for (int j = 0; j < g; ++j) {
            st.partner_bits.push_back((uint8_t)partners[j]);
            const int lg = logical_at[n_local + j], ll = logical_at[partners[j]];
            std::swap(layout[lg], layout[ll]);
        }
        plan.steps.push_back(std::move(st));
        // the op must be local now
        for (int b = 0; b < n; ++b)
            if (((tg >> b) & 1) && layout[b] >= n_local) fail("gate " + std::to_string(plan.lops[i].src_gate) + " needs more qubits on one rank than a remap provides");
    }
    emit_segment(seg_start, n_lops);
    plan.opt.low_bits = chosen_L;
    plan.final_layout.assign(layout.begin(), layout.begin() + n);
}

void prefix_amplitudes(const Plan& plan, uint64_t basis_index, std::vector<cplx>& out) {
    const uint32_t n = plan.n_qubits, g = n - plan.n_local;
    const uint32_t nl = plan.n_local - plan.prefix_local_bits;  // bits below the support: constants of the basis state
    const uint64_t P = 1ull << (n - nl);
    out.assign(P, cplx{0.0, 0.0});
    // physical index of the basis state under the initial layout
    uint64_t phys = 0;
    for (uint32_t b = 0; b < n; ++b) phys |= ((basis_index >> b) & 1ull) << plan.initial_layout[b];
    out[phys >> nl] = cplx{1.0, 0.0};
    (void)g;
    if (plan.prefix.empty()) return;
    // the prefix exists only under the canonical layout: logical bit = physical bit; the bits below the support are constants
    const uint64_t local = phys & ((1ull << nl) - 1ull), rank_mask = (P - 1ull) << nl;
    for (const LOp& op : plan.prefix) {
        if (op.kind == LOp::MAT) {
            const bool local_sat = (local & op.cmask & ~rank_mask) == (op.cmask & ~rank_mask);  // controls on the constant bits
            if (!local_sat && !op.dual) continue;
            const uint64_t cm = (op.cmask & rank_mask) >> nl, tb = 1ull << (op.target - (int)nl);
            for (uint64_t r = 0; r < P; ++r) {
                if (r & tb) continue;
                const bool sat = local_sat && (r & cm) == cm;
                if (!sat && !op.dual) continue;
                const double* m = sat ? op.m : op.m2;  // dual ops: m2 where the controls do not hold
                const cplx m00{m[0], m[1]}, m01{m[2], m[3]}, m10{m[4], m[5]}, m11{m[6], m[7]};
                const cplx a0 = out[r], a1 = out[r | tb];
                const cplx p00 = cmul(m00, a0), p01 = cmul(m01, a1), p10 = cmul(m10, a0), p11 = cmul(m11, a1);
                out[r] = cplx{p00.x + p01.x, p00.y + p01.y};
                out[r | tb] = cplx{p10.x + p11.x, p10.y + p11.y};
            }
        } else if (op.kind == LOp::DIAG) {
            for (uint64_t r = 0; r < P; ++r) {
                const uint64_t idx = (r << nl) | local;
                if ((idx & op.cmask) != op.cmask) continue;
                double ang = op.theta0;
                for (auto& t : op.lin)
                    if ((idx >> t.first) & 1ull) ang += t.second;
                out[r] = cmul(out[r], unit_phase(ang));
            }
        }
    }
}

// The folded prefix as a circuit of its own on the g + k support qubits (index bit b of the register = bit b - nf of the
// sub-register, nf = n - g - k): controls on the constant low bits are resolved against the basis state (an op whose
// control is 0 there vanishes), phases that depend on them become constants.  The sub-register's final state is the table
// prefix_amplitudes computes on the host - here for prefixes too wide for that (k > kHostPrefixBits).
void build_prefix_subplan(const Plan& plan, uint64_t basis_index, Plan& sub) {
    const uint32_t n = plan.n_qubits, nf = plan.n_local - plan.prefix_local_bits, ns = n - nf;
    const uint64_t low_mask = (1ull << nf) - 1ull, x_low = basis_index & low_mask;  // canonical layout: physical = logical
    sub = Plan();
    for (const LOp& op : plan.prefix) {
        LOp o = op;
        if (op.kind == LOp::DENSE) fail("internal: dense op in a folded prefix");
        const uint64_t c_low = op.cmask & low_mask;
        const bool low_sat = (x_low & c_low) == c_low;
        o.cmask = op.cmask >> nf;
        if (op.kind == LOp::MAT) {
            o.target = op.target - (int)nf;
            if (!low_sat) {  // the controls on the constant bits fail: the op is its "elsewhere" matrix (dual) or nothing
                if (!op.dual) continue;
                memcpy(o.m, op.m2, sizeof(o.m));
                o.dual = false;
                o.cmask = 0;
                OpType t;
                if (!classify_mat(o.m, &t)) continue;
                o.mtype = t;
            }
        } else {
            if (!low_sat) continue;
            o.lin.clear();
            for (auto& t : op.lin) {
                if (t.first < (int)nf) { if ((x_low >> t.first) & 1ull) o.theta0 += t.second; }
                else o.lin.push_back({t.first - (int)nf, t.second});
            }
        }
        sub.lops.push_back(std::move(o));
    }
    PlanOptions opt = plan.opt;
    opt.fold_prefix = 0;
    opt.tile_bits = 0;
    opt.low_bits = 0;
    sub.n_gates = plan.prefix.size();
    build_plan(sub, ns, ns, nullptr, sub.lops.empty() ? 0 : 1, opt, nullptr, false);  // (ops = NULL with n_ops = 1: keep sub.lops)
}

bool plan_overlap_group(const Plan& plan, size_t step, const std::vector<char>& sliceable, uint32_t log2_slices, OverlapGroup& out) {
    out = OverlapGroup();
    if (step >= plan.steps.size() || plan.steps[step].kind != PlanStep::EXCHANGE || log2_slices == 0) return false;
    if (log2_slices > 3) log2_slices = 3;
    const uint64_t local_mask = (1ull << plan.n_local) - 1ull;
    auto tile_mask_of = [&](size_t st) {
        const DevPass& h = *reinterpret_cast<const DevPass*>(plan.passes[plan.steps[st].pass_index].data());
        uint64_t m = 0;
        for (uint32_t sgi = 0; sgi < h.n_tile_segs; ++sgi) m |= ((1ull << h.tile_segs[sgi].width) - 1ull) << h.tile_segs[sgi].dst_lo;
        return m;
    };
    uint64_t partners = 0;
    for (uint8_t b : plan.steps[step].partner_bits) partners |= 1ull << b;
    const bool has_prev = step > 0 && plan.steps[step - 1].kind == PlanStep::PASS && sliceable[step - 1];
    const bool has_next = step + 1 < plan.steps.size() && plan.steps[step + 1].kind == PlanStep::PASS && sliceable[step + 1];
    // candidates, best first: both neighbours sliced, then the one after, then the one before
    for (int attempt = 0; attempt < 3; ++attempt) {
        const bool use_prev = has_prev && attempt != 1, use_next = has_next && attempt != 2;
        if ((attempt == 0 && !(has_prev && has_next)) || (!use_prev && !use_next)) continue;
        uint64_t blocked = partners;
        if (use_prev) blocked |= tile_mask_of(step - 1);
        if (use_next) blocked |= tile_mask_of(step + 1);
        const uint64_t free_bits = local_mask & ~blocked;
        if ((uint32_t)popcnt(free_bits) < log2_slices) continue;
        // the highest free bits: slices are as contiguous as the passes allow, and far above the groups' tile interleave
        uint32_t n = 0;
        uint8_t picked[3];
        auto id_pos_ok = [&](size_t st, int b) {  // at least three tile-id bits below the slice bit (kernel: the groups' interleaved tile ids)
            return popcnt(~tile_mask_of(st) & ((1ull << b) - 1ull)) >= 3;
        };
        for (int b = (int)plan.n_local - 1; b >= 6 && n < log2_slices; --b)
            if (((free_bits >> b) & 1ull) && (!use_prev || id_pos_ok(step - 1, b)) && (!use_next || id_pos_ok(step + 1, b))) picked[n++] = (uint8_t)b;
        if (n < log2_slices) continue;
        out.slice_prev = use_prev;
        out.slice_next = use_next;
        out.n_bits = n;
        for (uint32_t i = 0; i < n; ++i) out.bits[i] = picked[n - 1 - i];  // ascending
        return true;
    }
    return false;
}

std::string describe_plan(const Plan& plan) {
    std::ostringstream os;
    os << "{\"n_qubits\":" << plan.n_qubits << ",\"n_local\":" << plan.n_local << ",\"n_alloc\":" << plan.n_alloc
       << ",\"tile_bits\":" << std::min<int>(plan.opt.tile_bits, (int)plan.n_alloc) << ",\"low_bits\":" << plan.opt.low_bits
       << ",\"n_gates\":" << plan.n_gates << ",\"n_lowered_ops\":" << plan.lops.size() << ",\"prefix_ops\":" << plan.prefix.size()
       << ",\"prefix_local_bits\":" << plan.prefix_local_bits;
    size_t n_dual = 0;
    for (const LOp& lop : plan.lops) n_dual += lop.kind == LOp::MAT && lop.dual;
    os << ",\"n_dual_ops\":" << n_dual << ",\"passes\":[";
    for (size_t p = 0; p < plan.passes.size(); ++p) {
        const uint8_t* blob = plan.passes[p].data();
        const DevPass* h = reinterpret_cast<const DevPass*>(blob);
        const DevRound* rounds = reinterpret_cast<const DevRound*>(blob + h->rounds_off);
        const DevOp* ops = reinterpret_cast<const DevOp*>(blob + h->ops_off);
        os << (p ? "," : "") << "{\"tile\":[";
        bool first = true;
        for (uint32_t s = 0; s < h->n_tile_segs; ++s)
            for (uint32_t b = 0; b < h->tile_segs[s].width; ++b) { os << (first ? "" : ",") << (int)h->tile_segs[s].dst_lo + (int)b; first = false; }
        os << "],\"n_tiles\":" << h->n_tiles << ",\"bytes\":" << h->blob_bytes << ",\"flags\":" << h->flags << ",\"rounds\":[";
        for (uint32_t r = 0; r < h->n_rounds; ++r) {
            os << (r ? "," : "") << "{\"type\":" << rounds[r].type << ",\"n_perm\":" << (rounds[r].type == ROUND_PERM ? rounds[r].n_perm : 0u) << ",\"reg\":[" << (int)rounds[r].reg_pos[0] << "," << (int)rounds[r].reg_pos[1]
               << "," << (int)rounds[r].reg_pos[2] << "," << (int)rounds[r].reg_pos[3] << "],\"ops\":[";
            for (uint32_t o = 0; o < rounds[r].n_ops; ++o) os << (o ? "," : "") << ops[rounds[r].first_op + o].type;
            os << "]}";
        }
        os << "]}";
    }
    os << "]}";
    return os.str();
}

}  // namespace qsv
