// quantr_oracle.cpp — CPU restatement of quantr's `Circuit::simulate` hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (quantr_b200/, libqsv.so) may
// link, load or call this file.  It is used by tests/, by __graft_entry__.smoke()
// as the checker, and by bench.py's `cpu_baseline` / `--impl reference` legs as
// the timed CPU baseline ("kind": "port" — the reference is Rust and there is no
// Rust toolchain in this image, so it cannot be compiled into oracle/_ref).
//
// Parity pinning: tests/test_oracle_golden.py checks both entry points below
// against the reference's own 24 golden state vectors (src/circuit.rs:603-982,
// tests/qft.rs:30-39, tests/grovers.rs:45-54; restated as data in
// tests/golden/reference_vectors.py) and its statistical assertions.
//
// Two implementations of the same semantics:
//   oracle_simulate_faithful  follows src/circuit/simulation.rs:21-180 step for
//       step, including the two per-gate hash maps keyed by a heap-allocated
//       qubit vector, the ascending-index accumulation order, the "first
//       contribution assigns, later ones add" rule and the None-overwrite rule.
//       This is the timed "reference CPU path".
//   oracle_simulate_dense     the same arithmetic in the same order, computed per
//       group of 2^k coupled amplitudes without hashing, so parity can be checked
//       at 20-28 qubits.  Validated against the faithful one at small n.
//
// Third-party arithmetic restated (crates not under /root/reference):
//   num-complex 0.4.6 (Cargo.toml:16): Complex mul = (ac-bd, ad+bc), add is
//   component-wise, exp(i*theta) = (cos theta, sin theta), norm_sqr = re^2+im^2.
//   fastrand 2.1.0 (Cargo.toml:15): only its f64() in [0,1) is used; the caller
//   supplies the uniforms, the WyRand stream itself is not restated.

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "../include/qsv.h"

namespace {

struct C64 {
    double re, im;
};
// num-complex Mul: (a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re); no FMA (build uses -ffp-contract=off).
inline C64 cmul(C64 a, C64 b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
inline C64 cadd(C64 a, C64 b) { return {a.re + b.re, a.im + b.im}; }
inline C64 cexp_i(double theta) { return {std::cos(theta), std::sin(theta)}; }  // Complex64::exp of (0, theta)

constexpr double S2 = 0.70710678118654752440;  // std::f64::consts::FRAC_1_SQRT_2
constexpr C64 ZERO{0.0, 0.0};
constexpr C64 ONE{1.0, 0.0};

// ---- standard_gate_ops.rs:37-267: image of one basis state = one column ----------------

// One-wire gates: q = the wire's qubit (0/1); out[0..2).
void column_single(uint32_t kind, double angle, int q, C64* out) {
    switch (kind) {
        case QSV_GATE_H:  // hadamard :37
            out[0] = {S2, 0}; out[1] = {q ? -S2 : S2, 0}; break;
        case QSV_GATE_RX: {  // rx :45
            C64 re{std::cos(0.5 * angle), 0}, im{0, -std::sin(0.5 * angle)};
            out[0] = q ? im : re; out[1] = q ? re : im; break;
        }
        case QSV_GATE_RY: {  // ry :58
            C64 c{std::cos(0.5 * angle), 0}, sp{std::sin(0.5 * angle), 0}, sn{-std::sin(0.5 * angle), 0};
            out[0] = q ? sn : c; out[1] = q ? c : sp; break;
        }
        case QSV_GATE_RZ: {  // rz :72
            out[0] = q ? ZERO : cexp_i(-angle * 0.5); out[1] = q ? cexp_i(angle * 0.5) : ZERO; break;
        }
        case QSV_GATE_PHASE: {  // global_phase :85
            C64 e = cexp_i(angle * 0.5);
            out[0] = q ? ZERO : e; out[1] = q ? e : ZERO; break;
        }
        case QSV_GATE_X90:  // x90 :97
            out[0] = q ? C64{0, -1} : ZERO; out[1] = q ? ZERO : C64{0, -1}; break;
        case QSV_GATE_Y90:  // y90 :105
            out[0] = q ? ONE : ZERO; out[1] = q ? ZERO : C64{-1, 0}; break;
        case QSV_GATE_MX90:  // mx90 :113
            out[0] = q ? C64{0, 1} : ZERO; out[1] = q ? ZERO : C64{0, 1}; break;
        case QSV_GATE_MY90:  // my90 :121
            out[0] = q ? C64{-1, 0} : ZERO; out[1] = q ? ZERO : ONE; break;
        case QSV_GATE_T:  // tgate :129
            out[0] = q ? ZERO : ONE; out[1] = q ? C64{S2, S2} : ZERO; break;
        case QSV_GATE_TDAG:  // tgatedag :137
            out[0] = q ? ZERO : ONE; out[1] = q ? C64{S2, -S2} : ZERO; break;
        case QSV_GATE_S:  // phase :145
            out[0] = q ? ZERO : ONE; out[1] = q ? C64{0, 1} : ZERO; break;
        case QSV_GATE_SDAG:  // phasedag :153
            out[0] = q ? ZERO : ONE; out[1] = q ? C64{0, -1} : ZERO; break;
        case QSV_GATE_X:  // pauli_x :161
            out[0] = q ? ONE : ZERO; out[1] = q ? ZERO : ONE; break;
        case QSV_GATE_Y:  // pauli_y :169
            out[0] = q ? C64{0, -1} : ZERO; out[1] = q ? ZERO : C64{0, 1}; break;
        case QSV_GATE_Z:  // pauli_z :177
            out[0] = q ? ZERO : ONE; out[1] = q ? C64{-1, 0} : ZERO; break;
        default: out[0] = out[1] = ZERO;
    }
}

// Two-wire gates on |c t>: sub = 2*c + t; out[0..4).
void column_double(uint32_t kind, double angle, int32_t k, int sub, C64* out) {
    for (int i = 0; i < 4; ++i) out[i] = ZERO;
    switch (kind) {
        case QSV_GATE_CNOT: {  // cnot :189
            static const int to[4] = {0, 1, 3, 2};
            out[to[sub]] = ONE; break;
        }
        case QSV_GATE_CY:  // cy :199
            if (sub == 2) out[3] = {0, 1};
            else if (sub == 3) out[2] = {0, -1};
            else out[sub] = ONE;
            break;
        case QSV_GATE_CZ:  // cz :209
            out[sub] = (sub == 3) ? C64{-1, 0} : ONE; break;
        case QSV_GATE_SWAP: {  // swap :219
            static const int to[4] = {0, 2, 1, 3};
            out[to[sub]] = ONE; break;
        }
        case QSV_GATE_CR:  // cr :229
            out[sub] = (sub == 3) ? cexp_i(angle) : ONE; break;
        case QSV_GATE_CRK:  // crk :240 — exp(i * (2*pi) / 2^k), 2f64.powi(k)
            out[sub] = (sub == 3) ? cexp_i((2.0 * M_PI) / std::pow(2.0, k)) : ONE; break;
        default: break;
    }
}

// toffoli :256 on |c1 c2 t>.
void column_triple(int sub, C64* out) {
    for (int i = 0; i < 8; ++i) out[i] = ZERO;
    static const int to[8] = {0, 1, 2, 3, 4, 5, 7, 6};
    out[to[sub]] = ONE;
}

int arity(uint32_t kind, uint32_t n_controls) {
    if (kind == QSV_GATE_CUSTOM) return (int)n_controls + 1;
    if (kind == QSV_GATE_TOFFOLI) return 3;
    if (kind >= QSV_GATE_CR && kind <= QSV_GATE_SWAP) return 2;
    return 1;
}

bool op_is_valid(uint32_t n, const qsv_op& op) {
    if (op.kind == QSV_GATE_ID) return true;
    if (op.kind >= QSV_GATE_KIND_COUNT || op.target >= n) return false;
    int k = arity(op.kind, op.n_controls);
    if (op.kind != QSV_GATE_CUSTOM && (int)op.n_controls != k - 1) return false;
    if (op.n_controls && !op.controls) return false;
    for (uint32_t i = 0; i < op.n_controls; ++i) {
        if (op.controls[i] >= n || op.controls[i] == op.target) return false;
        for (uint32_t j = 0; j < i; ++j)
            if (op.controls[i] == op.controls[j]) return false;
    }
    if (op.kind == QSV_GATE_CUSTOM && (k > 20)) return false;
    return true;
}

// Column of a gate for sub-state `sub` over positions [controls..., target]
// (gate.rs:140-168 dispatch + simulation.rs:75-106).  Returns false for a Custom None.
typedef int (*oracle_custom_fn)(void* ctx, uint32_t op_index, const uint8_t* qubits, uint32_t k, double* out_amps);

bool gate_column(const qsv_op& op, uint32_t op_index, int k, uint64_t sub, C64* out, oracle_custom_fn cb, void* ctx) {
    if (op.kind == QSV_GATE_CUSTOM) {
        const uint64_t dim = 1ull << k;
        if (cb) {
            uint8_t qubits[64];
            for (int e = 0; e < k; ++e) qubits[e] = (sub >> (k - 1 - e)) & 1;  // [controls..., target], simulation.rs:144-150
            return cb(ctx, op_index, qubits, (uint32_t)k, reinterpret_cast<double*>(out)) != 0;
        }
        if (op.none_mask && op.none_mask[sub]) return false;
        for (uint64_t t = 0; t < dim; ++t) out[t] = {op.matrix[(t * dim + sub) * 2], op.matrix[(t * dim + sub) * 2 + 1]};
        return true;
    }
    if (k == 1) column_single(op.kind, op.param, (int)sub, out);
    else if (k == 2) column_double(op.kind, op.param, op.iparam, (int)sub, out);
    else column_triple((int)sub, out);
    return true;
}

// ---- SipHash-1-3 over the key the way Rust's derived Hash sees a Vec<Qubit>:
// a usize length prefix, then one isize discriminant (8 bytes) per qubit.
struct SipHasher13 {
    uint64_t v0, v1, v2, v3;
    static inline uint64_t rotl(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
    inline void round() {
        v0 += v1; v1 = rotl(v1, 13); v1 ^= v0; v0 = rotl(v0, 32);
        v2 += v3; v3 = rotl(v3, 16); v3 ^= v2;
        v0 += v3; v3 = rotl(v3, 21); v3 ^= v0;
        v2 += v1; v1 = rotl(v1, 17); v1 ^= v2; v2 = rotl(v2, 32);
    }
    SipHasher13(uint64_t k0, uint64_t k1)
        : v0(k0 ^ 0x736f6d6570736575ull), v1(k1 ^ 0x646f72616e646f6dull), v2(k0 ^ 0x6c7967656e657261ull), v3(k1 ^ 0x7465646279746573ull) {}
    inline void word(uint64_t m) { v3 ^= m; round(); v0 ^= m; }
    inline uint64_t finish(uint64_t total_bytes) {
        word(total_bytes << 56);
        v2 ^= 0xff; round(); round(); round();
        return v0 ^ v1 ^ v2 ^ v3;
    }
};

typedef std::vector<uint8_t> ProductState;  // one byte per qubit, wire 0 first (product_states.rs:18-23)

struct ProductStateHash {
    size_t operator()(const ProductState& p) const {
        SipHasher13 h(0x0706050403020100ull, 0x0f0e0d0c0b0a0908ull);
        h.word(p.size());
        for (uint8_t q : p) h.word(q);
        return (size_t)h.finish(8 * (p.size() + 1));
    }
};

typedef std::unordered_map<ProductState, C64, ProductStateHash> StateMap;

// product_states.rs:205-215
ProductState binary_basis(uint64_t index, uint32_t n) {
    ProductState p(n);
    for (uint32_t q = 0; q < n; ++q) p[q] = (index >> (n - 1 - q)) & 1;
    return p;
}

void positions_of(const qsv_op& op, std::vector<uint32_t>& pos) {
    pos.clear();
    for (uint32_t i = 0; i < op.n_controls; ++i) pos.push_back(op.controls[i]);
    pos.push_back(op.target);  // simulation.rs:108-112
}

// simulation.rs:64-135
void apply_gate_faithful(uint32_t n, const qsv_op& op, uint32_t op_index, std::vector<C64>& reg, oracle_custom_fn cb, void* ctx) {
    const int k = arity(op.kind, op.n_controls);
    const uint64_t dim = 1ull << k;
    StateMap mapped, untouched;
    std::vector<uint32_t> pos;
    positions_of(op, pos);
    std::vector<C64> image(dim);
    const uint64_t len = 1ull << n;
    for (uint64_t i = 0; i < len; ++i) {  // zeros included (super_position_iter.rs:56-72)
        ProductState prod = binary_basis(i, n);
        const C64 amp = reg[i];
        uint64_t sub = 0;
        for (int e = 0; e < k; ++e) sub = (sub << 1) | prod[pos[e]];
        if (gate_column(op, op_index, k, sub, image.data(), cb, ctx)) {
            // insert_gate_image_into_product_state, simulation.rs:158-180
            for (uint64_t t = 0; t < dim; ++t) {
                ProductState swapped = prod;  // clone
                for (int e = 0; e < k; ++e) swapped[pos[e]] = (t >> (k - 1 - e)) & 1;  // insert_qubits, product_states.rs:112-123
                const C64 contrib = cmul(image[t], amp);  // state_amp.mul(amp)
                auto it = mapped.find(swapped);
                if (it == mapped.end()) mapped.emplace(std::move(swapped), contrib);
                else it->second = cadd(it->second, contrib);
            }
        } else {
            untouched.emplace(std::move(prod), amp);  // simulation.rs:120-122
        }
    }
    for (auto& kv : untouched) mapped[kv.first] = kv.second;  // overwrite, simulation.rs:126-133
    // set_amplitudes_from_states_unchecked, super_positions_unchecked.rs:50-60
    for (uint64_t i = 0; i < len; ++i) {
        auto it = mapped.find(binary_basis(i, n));
        if (it == mapped.end()) reg[i] = ZERO;
        else { reg[i] = it->second; mapped.erase(it); }
    }
}

// Same arithmetic, same order, per group of coupled amplitudes.
void apply_gate_dense(uint32_t n, const qsv_op& op, uint32_t op_index, C64* reg, oracle_custom_fn cb, void* ctx, int threads) {
    const int k = arity(op.kind, op.n_controls);
    const uint64_t dim = 1ull << k;
    std::vector<uint32_t> pos;
    positions_of(op, pos);
    std::vector<int> bit(k);
    uint64_t gate_mask = 0;
    for (int e = 0; e < k; ++e) { bit[e] = (int)(n - 1 - pos[e]); gate_mask |= 1ull << bit[e]; }
    // columns once per gate (pure functions of the sub-state)
    std::vector<C64> cols(dim * dim);
    std::vector<uint8_t> none(dim, 0);
    for (uint64_t s = 0; s < dim; ++s) none[s] = gate_column(op, op_index, k, s, &cols[s * dim], cb, ctx) ? 0 : 1;
    std::vector<uint64_t> offs(dim);
    for (uint64_t t = 0; t < dim; ++t) {
        uint64_t o = 0;
        for (int e = 0; e < k; ++e) if ((t >> (k - 1 - e)) & 1) o |= 1ull << bit[e];
        offs[t] = o;
    }
    const uint64_t groups = 1ull << (n - k);
    std::vector<int> sorted_bits(bit);
    for (size_t a = 0; a < sorted_bits.size(); ++a)
        for (size_t b = a + 1; b < sorted_bits.size(); ++b)
            if (sorted_bits[b] < sorted_bits[a]) std::swap(sorted_bits[a], sorted_bits[b]);
    const int nthreads = threads > 0 ? threads : 1;
    // r enumerates a group's members in ascending canonical index (by sorted bit positions); s_of_r[r] = its sub-state
    std::vector<uint32_t> s_of_r(dim);
    for (uint64_t r = 0; r < dim; ++r) {
        uint64_t o = 0;
        for (int e = 0; e < k; ++e) if ((r >> e) & 1) o |= 1ull << sorted_bits[e];
        uint64_t s = 0;
        for (int e = 0; e < k; ++e) s = (s << 1) | ((o >> bit[e]) & 1);
        s_of_r[r] = (uint32_t)s;
    }
    // Accumulate in ascending canonical index of the inputs (the order the reference's iterator visits them).
    // `DIM` > 0: compile-time group size for the standard 1/2/3-wire gates (same operations, fixed-size buffers).
    auto run_groups = [&](auto dim_tag, int tid) {
        constexpr uint64_t DIM = decltype(dim_tag)::value;
        const uint64_t d = DIM ? DIM : dim;
        C64 in_fix[DIM ? DIM : 1], out_fix[DIM ? DIM : 1];
        uint8_t seen_fix[DIM ? DIM : 1];
        std::vector<C64> in_v(DIM ? 0 : dim), out_v(DIM ? 0 : dim);
        std::vector<uint8_t> seen_v(DIM ? 0 : dim);
        C64* in = DIM ? in_fix : in_v.data();
        C64* out = DIM ? out_fix : out_v.data();
        uint8_t* seen = DIM ? seen_fix : seen_v.data();
        const uint64_t g0 = groups * (uint64_t)tid / (uint64_t)nthreads, g1 = groups * (uint64_t)(tid + 1) / (uint64_t)nthreads;
        for (uint64_t g = g0; g < g1; ++g) {
            uint64_t base = g;
            for (int b : sorted_bits) base = ((base >> b) << (b + 1)) | (base & ((1ull << b) - 1));
            for (uint64_t s = 0; s < d; ++s) { in[s] = reg[base | offs[s]]; seen[s] = 0; out[s] = ZERO; }
            for (uint64_t r = 0; r < d; ++r) {
                const uint64_t s = s_of_r[r];
                if (none[s]) continue;
                const C64* col = &cols[s * d];
                for (uint64_t t = 0; t < d; ++t) {
                    const C64 c = cmul(col[t], in[s]);
                    if (!seen[t]) { out[t] = c; seen[t] = 1; } else out[t] = cadd(out[t], c);
                }
            }
            for (uint64_t s = 0; s < d; ++s) if (none[s]) out[s] = in[s];  // None-overwrite, simulation.rs:126-133
            for (uint64_t s = 0; s < d; ++s) reg[base | offs[s]] = out[s];
        }
    };
    auto worker = [&](int tid) {
        if (k == 1) run_groups(std::integral_constant<uint64_t, 2>(), tid);
        else if (k == 2) run_groups(std::integral_constant<uint64_t, 4>(), tid);
        else if (k == 3) run_groups(std::integral_constant<uint64_t, 8>(), tid);
        else run_groups(std::integral_constant<uint64_t, 0>(), tid);
    };
    if (nthreads == 1) { worker(0); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) pool.emplace_back(worker, t);
    for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

// amps: 2^n complex f64 interleaved, canonical order, updated in place.
// `custom_cb` may be NULL; then Custom ops use op.matrix / op.none_mask as the closure's table.
int oracle_simulate_faithful(uint32_t n, const qsv_op* ops, size_t n_ops, double* amps, oracle_custom_fn custom_cb, void* ctx) {
    if (n == 0 || n > 26 || !amps) return 1;
    for (size_t g = 0; g < n_ops; ++g) if (!op_is_valid(n, ops[g])) return 1;
    std::vector<C64> reg(1ull << n);
    std::memcpy(reg.data(), amps, sizeof(C64) << n);
    for (size_t g = 0; g < n_ops; ++g) {
        if (ops[g].kind == QSV_GATE_ID) continue;  // simulation.rs:38-41
        apply_gate_faithful(n, ops[g], (uint32_t)g, reg, custom_cb, ctx);
    }
    std::memcpy(amps, reg.data(), sizeof(C64) << n);
    return 0;
}

int oracle_simulate_dense(uint32_t n, const qsv_op* ops, size_t n_ops, double* amps, oracle_custom_fn custom_cb, void* ctx, int threads) {
    if (n == 0 || n > 34 || !amps) return 1;
    for (size_t g = 0; g < n_ops; ++g) if (!op_is_valid(n, ops[g])) return 1;
    for (size_t g = 0; g < n_ops; ++g) {
        if (ops[g].kind == QSV_GATE_ID) continue;
        apply_gate_dense(n, ops[g], (uint32_t)g, reinterpret_cast<C64*>(amps), custom_cb, ctx, threads);
    }
    return 0;
}

// SuperPosition::measure (super_positions.rs:332-342) for a caller-supplied dice roll per shot.
void oracle_measure_all(uint32_t n, const double* amps, const double* uniforms, uint64_t shots, uint64_t* out_indices) {
    const uint64_t len = 1ull << n;
    for (uint64_t s = 0; s < shots; ++s) {
        double cumulative = 0.0;
        const double dice = uniforms[s];
        uint64_t found = UINT64_MAX;
        for (uint64_t i = 0; i < len; ++i) {
            cumulative += amps[2 * i] * amps[2 * i] + amps[2 * i + 1] * amps[2 * i + 1];  // norm_sqr
            if (dice < cumulative) { found = i; break; }
        }
        out_indices[s] = found;
    }
}

// Fast equivalent for many shots: one cumulative array, same sequential f64 sums, binary search
// for the first i with dice < cum[i] (cum is non-decreasing, so this is the same index).
void oracle_measure_all_cdf(uint32_t n, const double* amps, const double* uniforms, uint64_t shots, uint64_t* out_indices) {
    const uint64_t len = 1ull << n;
    std::vector<double> cum(len);
    double c = 0.0;
    for (uint64_t i = 0; i < len; ++i) { c += amps[2 * i] * amps[2 * i] + amps[2 * i + 1] * amps[2 * i + 1]; cum[i] = c; }
    for (uint64_t s = 0; s < shots; ++s) {
        const double dice = uniforms[s];
        uint64_t lo = 0, hi = len;  // first index with dice < cum[idx]
        while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (dice < cum[mid]) hi = mid; else lo = mid + 1; }
        out_indices[s] = (lo == len) ? UINT64_MAX : lo;
    }
}

int oracle_version(void) { return 1; }

}  // extern "C"
