// quantr.hpp — C++ host mirror of quantr's public API over the C ABI (include/qsv.h).
//
// The reference's host code is Rust (src/circuit.rs, src/circuit/gate.rs, src/simulated_circuit.rs,
// src/circuit/states/*).  This image has no Rust toolchain, so the compiled host layer above the C ABI
// is written in C++ with the same names, argument meaning and error behaviour; the Rust shim a
// maintainer would add is shown in INTEGRATION.md and rust/.  Everything here is host-side bookkeeping
// (builder, validation, column layout, Custom-closure expansion, binning); the amplitudes live in HBM
// behind a qsv_state handle and every gate application happens in libqsv.so.
//
// Rust `Result<T, QuantrError>` maps to "returns T or throws QuantrError" (`.unwrap()` == let it throw).
#pragma once
#include <complex>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/qsv.h"

namespace quantr {

using Complex64 = std::complex<double>;

// src/error.rs:17-33
struct QuantrError : std::runtime_error {
    std::string message;
    explicit QuantrError(const std::string& m) : std::runtime_error("\x1b[91m[Quantr Error] " + m + "\x1b[0m "), message(m) {}
};

namespace states {

enum class Qubit : uint8_t { Zero = 0, One = 1 };  // src/circuit/states/qubit.rs:15

class SuperPosition;

// src/circuit/states/product_states.rs:18-23 — wire 0 first
class ProductState {
public:
    std::vector<Qubit> qubits;
    ProductState() = default;
    explicit ProductState(std::vector<Qubit> q) : qubits(std::move(q)) {}
    static ProductState make(const std::vector<Qubit>& q) {  // ProductState::new, :37
        if (q.empty()) throw QuantrError("The slice of qubits is empty, it needs to at least have one element.");
        return ProductState(q);
    }
    static ProductState new_unchecked(const std::vector<Qubit>& q) { return ProductState(q); }
    const std::vector<Qubit>& get_qubits() const { return qubits; }
    std::vector<Qubit>& get_mut_qubits() { return qubits; }
    size_t num_qubits() const { return qubits.size(); }
    ProductState& invert_digit(size_t place) {  // :154
        if (place >= qubits.size())
            throw QuantrError("The position of the binary digit, " + std::to_string(place) + ", is out of bounds. The product dimension is " +
                              std::to_string(qubits.size()) + ", and so the position must be strictly less.");
        qubits[place] = qubits[place] == Qubit::Zero ? Qubit::One : Qubit::Zero;
        return *this;
    }
    ProductState kronecker_prod(Qubit other) const { ProductState p = *this; p.qubits.push_back(other); return p; }  // :180
    uint64_t comp_basis() const {  // :191 (without the reference's u32 overflow)
        uint64_t v = 0;
        for (Qubit q : qubits) v = (v << 1) | (uint64_t)q;
        return v;
    }
    static ProductState binary_basis(uint64_t index, size_t basis_size) {  // :205-215
        std::vector<Qubit> q(basis_size);
        for (size_t w = 0; w < basis_size; ++w) q[w] = ((index >> (basis_size - 1 - w)) & 1) ? Qubit::One : Qubit::Zero;
        return ProductState(std::move(q));
    }
    std::string to_string() const {  // :218-240
        std::string s;
        for (Qubit q : qubits) s += q == Qubit::One ? '1' : '0';
        return s;
    }
    bool operator==(const ProductState& o) const { return qubits == o.qubits; }
    bool operator<(const ProductState& o) const { return qubits < o.qubits; }
};

constexpr double ZERO_MARGIN = 1e-6;  // src/circuit/states/super_positions.rs:18

// src/circuit/states/super_positions.rs:22-25 — dense amplitude vector in canonical order (host IO type)
class SuperPosition {
public:
    std::vector<Complex64> amplitudes;
    size_t product_dim = 0;

    static SuperPosition make(size_t prod_dimension) {  // SuperPosition::new, :41
        if (prod_dimension == 0) throw QuantrError("The number of qubits must be non-zero.");
        return new_unchecked(prod_dimension);
    }
    static SuperPosition new_unchecked(size_t n) {  // super_positions_unchecked.rs:39-46
        SuperPosition s;
        s.amplitudes.assign((size_t)1 << n, Complex64(0, 0));
        s.amplitudes[0] = Complex64(1, 0);
        s.product_dim = n;
        return s;
    }
    static bool equal_within_error(double num, double compare_num) { return num < compare_num + ZERO_MARGIN && num > compare_num - ZERO_MARGIN; }  // :246-248
    static void check(const std::vector<Complex64>& a) {  // :69-73, :236-240
        double total = 0;
        for (auto& x : a) total += std::norm(x);
        if (!equal_within_error(total, 1.0))
            throw QuantrError("Slice given to set amplitudes in super position does not conserve probability, the absolute square sum of the coefficents must be one.");
    }
    static SuperPosition new_with_amplitudes(const std::vector<Complex64>& a) {  // :68-88 (probability first, then length)
        check(a);
        if (a.size() & (a.size() - 1)) throw QuantrError("The length of the array must be of the form 2**n where n is an integer.");
        return new_with_amplitudes_unchecked(a);
    }
    // :113-123 / :279-290: every key has `dim` qubits and the squares sum to one; then from_hash_to_array (:344-357)
    static std::vector<Complex64> from_states(const std::map<ProductState, Complex64>& h, size_t dim, bool report_total) {
        if (h.empty()) throw QuantrError("An empty HashMap was given. A superposition must have at least one non-zero state.");
        double total = 0;
        for (auto& kv : h) {
            if (kv.first.num_qubits() != dim)
                throw QuantrError("The first state has product dimension of " + std::to_string(dim) + ", whilst the state, |" + kv.first.to_string() +
                                  ">, found as a key in the HashMap has dimension " + std::to_string(kv.first.num_qubits()) + ".");
            total += std::norm(kv.second);
        }
        if (!equal_within_error(total, 1.0))
            throw QuantrError("The total sum of the absolute square of all amplitudes" + (report_total ? ", " + std::to_string(total) + "," : std::string()) +
                              " does not equal 1. That is, the superpositon does not conserve probability.");
        std::vector<Complex64> a((size_t)1 << dim, Complex64(0, 0));
        for (auto& kv : h) a[kv.first.comp_basis()] = kv.second;
        return a;
    }
    static SuperPosition new_with_hash_amplitudes(const std::map<ProductState, Complex64>& h) {  // :105-131
        SuperPosition s;
        s.product_dim = h.empty() ? 0 : h.begin()->first.num_qubits();
        s.amplitudes = from_states(h, s.product_dim, true);
        return s;
    }
    static SuperPosition new_with_amplitudes_unchecked(const std::vector<Complex64>& a) {  // super_positions_unchecked.rs:64
        SuperPosition s;
        s.amplitudes = a;
        size_t len = a.size(), tz = 0;
        while (len > 1 && !(len & 1)) { len >>= 1; ++tz; }
        s.product_dim = tz;
        return s;
    }
    static SuperPosition from(const ProductState& p) {  // impl From<ProductState>, :360
        SuperPosition s;
        s.product_dim = p.num_qubits();
        s.amplitudes.assign((size_t)1 << s.product_dim, Complex64(0, 0));
        s.amplitudes[p.comp_basis()] = Complex64(1, 0);
        return s;
    }
    size_t get_num_qubits() const { return product_dim; }
    size_t get_dimension() const { return amplitudes.size(); }
    const std::vector<Complex64>& get_amplitudes() const { return amplitudes; }
    std::optional<Complex64> get_amplitude(size_t pos) const { return pos < amplitudes.size() ? std::optional<Complex64>(amplitudes[pos]) : std::nullopt; }
    Complex64 get_amplitude_from_state(const ProductState& p) const {  // :207
        if (p.num_qubits() != product_dim) throw QuantrError("Unable to retreive product state, |\"" + p.to_string() + "\"> with dimension " + std::to_string(p.num_qubits()) +
                              ". The superposition is a linear combination of states with different dimension. These dimensions should be equal.");
        return amplitudes[p.comp_basis()];
    }
    SuperPosition& set_amplitudes(const std::vector<Complex64>& a) {  // :229
        if (a.size() != amplitudes.size())
            throw QuantrError("The slice given to set the amplitudes in the computational basis has length " + std::to_string(a.size()) +
                              ", when it should have length " + std::to_string(amplitudes.size()) + ".");
        check(a);
        amplitudes = a;
        return *this;
    }
    SuperPosition& set_amplitudes_from_states(const std::map<ProductState, Complex64>& h) {  // :270-295
        amplitudes = from_states(h, product_dim, false);
        return *this;
    }
    std::map<ProductState, Complex64> to_hash_map() const {  // :315-323 (amplitudes with |a|^2 within 1e-6 of zero are left out)
        std::map<ProductState, Complex64> m;
        for (size_t i = 0; i < amplitudes.size(); ++i)
            if (!equal_within_error(std::norm(amplitudes[i]), 0.0)) m[ProductState::binary_basis(i, product_dim)] = amplitudes[i];
        return m;
    }
    // :332-342: first state whose running probability exceeds the roll (strict <); nullopt when the squares fall short of it
    std::optional<ProductState> measure(double dice_roll) const {
        double cumulative = 0;
        for (size_t i = 0; i < amplitudes.size(); ++i) {
            cumulative += std::norm(amplitudes[i]);
            if (dice_roll < cumulative) return ProductState::binary_basis(i, product_dim);
        }
        return std::nullopt;
    }
    std::optional<ProductState> measure() const;  // roll from the package generator (fastrand::f64 in the reference)
};

}  // namespace states

using states::ProductState;
using states::Qubit;
using states::SuperPosition;

// src/circuit/measurement.rs:16-28
template <class T>
struct Measurement {
    enum Kind { Observable, NonObservable } kind;
    T value;
    T take() { return std::move(value); }
};

using CustomFn = std::function<std::optional<SuperPosition>(ProductState)>;

// src/circuit/gate.rs:18-106
struct Gate {
    uint32_t kind = QSV_GATE_ID;
    double param = 0;
    int32_t iparam = 0;
    std::vector<uint32_t> controls;
    CustomFn func;
    std::string name;

    static Gate simple(uint32_t k) { Gate g; g.kind = k; return g; }
    static Gate Id() { return simple(QSV_GATE_ID); }
    static Gate H() { return simple(QSV_GATE_H); }
    static Gate X() { return simple(QSV_GATE_X); }
    static Gate Y() { return simple(QSV_GATE_Y); }
    static Gate Z() { return simple(QSV_GATE_Z); }
    static Gate S() { return simple(QSV_GATE_S); }
    static Gate Sdag() { return simple(QSV_GATE_SDAG); }
    static Gate T() { return simple(QSV_GATE_T); }
    static Gate Tdag() { return simple(QSV_GATE_TDAG); }
    static Gate X90() { return simple(QSV_GATE_X90); }
    static Gate Y90() { return simple(QSV_GATE_Y90); }
    static Gate MX90() { return simple(QSV_GATE_MX90); }
    static Gate MY90() { return simple(QSV_GATE_MY90); }
    static Gate Rx(double a) { Gate g = simple(QSV_GATE_RX); g.param = a; return g; }
    static Gate Ry(double a) { Gate g = simple(QSV_GATE_RY); g.param = a; return g; }
    static Gate Rz(double a) { Gate g = simple(QSV_GATE_RZ); g.param = a; return g; }
    static Gate Phase(double a) { Gate g = simple(QSV_GATE_PHASE); g.param = a; return g; }
    static Gate CR(double a, uint32_t c) { Gate g = simple(QSV_GATE_CR); g.param = a; g.controls = {c}; return g; }
    static Gate CRk(int32_t k, uint32_t c) { Gate g = simple(QSV_GATE_CRK); g.iparam = k; g.controls = {c}; return g; }
    static Gate CZ(uint32_t c) { Gate g = simple(QSV_GATE_CZ); g.controls = {c}; return g; }
    static Gate CY(uint32_t c) { Gate g = simple(QSV_GATE_CY); g.controls = {c}; return g; }
    static Gate CNot(uint32_t c) { Gate g = simple(QSV_GATE_CNOT); g.controls = {c}; return g; }
    static Gate Swap(uint32_t c) { Gate g = simple(QSV_GATE_SWAP); g.controls = {c}; return g; }
    static Gate Toffoli(uint32_t c1, uint32_t c2) { Gate g = simple(QSV_GATE_TOFFOLI); g.controls = {c1, c2}; return g; }
    static Gate Custom(CustomFn f, std::vector<uint32_t> ctrls, std::string nm) {
        Gate g = simple(QSV_GATE_CUSTOM);
        g.func = std::move(f); g.controls = std::move(ctrls); g.name = std::move(nm);
        return g;
    }
    bool is_id() const { return kind == QSV_GATE_ID; }
    bool is_single_gate() const { return kind <= QSV_GATE_PHASE; }  // gate.rs:173-201
    bool is_custom_gate() const { return kind == QSV_GATE_CUSTOM; }  // gate.rs:203-208
    // structural equality (Custom closures compare by name + controls: std::function has no identity)
    bool operator==(const Gate& o) const { return kind == o.kind && param == o.param && iparam == o.iparam && controls == o.controls && name == o.name; }
};

inline std::mt19937_64& rng() { static std::mt19937_64 g(0x9E3779B97F4A7C15ull); return g; }
inline void seed(uint64_t s) { rng().seed(s); }  // the reference's fastrand::seed
inline double next_f64() { return std::uniform_real_distribution<double>(0.0, 1.0)(rng()); }  // fastrand::f64(): [0,1)

namespace detail {

[[noreturn]] inline void ffi_panic(qsv_state* s, int code, const char* what) {
    // simulate()/measure_all() are infallible in the reference (no Result): FFI failures panic
    throw std::runtime_error(std::string("quantr-b200 FFI failure in ") + what + " (code " + std::to_string(code) + "): " + qsv_last_error(s));
}

struct StateHandle {  // owns the qsv_state* ; freed on drop
    qsv_state* h = nullptr;
    // QSV_DEVICES=0,1,..: one handle over several GPUs of the process (qsv_create_multi); registers too small to shard
    // (fewer than four qubits per device) stay on the first device
    explicit StateHandle(uint32_t n, int device = 0) {
        std::vector<int> devices;
        if (const char* env = getenv("QSV_DEVICES"))
            for (const char* p = env; *p;) {
                char* end = nullptr;
                const long v = strtol(p, &end, 10);
                if (end == p) break;
                devices.push_back((int)v);
                p = *end == ',' ? end + 1 : end;
            }
        auto log2_of = [](size_t c) { uint32_t g = 0; while (((size_t)1 << (g + 1)) <= c) ++g; return g; };
        while (devices.size() > 1 && n < 4 + log2_of(devices.size())) devices.resize(devices.size() / 2);
        const int rc = devices.size() > 1 ? qsv_create_multi(&h, n, devices.data(), (int)devices.size()) : qsv_create(&h, n, devices.empty() ? device : devices[0]);
        if (rc != QSV_OK) ffi_panic(nullptr, rc, "qsv_create");
    }
    ~StateHandle() { if (h) qsv_destroy(h); }
    StateHandle(const StateHandle&) = delete;
    StateHandle& operator=(const StateHandle&) = delete;
};

// Gate list -> qsv_op[] exactly as simulation.rs:37-56 walks it; Custom closures are evaluated on the 2^k basis
// states of [controls..., target] (simulation.rs:137-156) into a matrix + none mask.
constexpr size_t kCompactCustomWires = 11;  // same threshold as the Python and Rust hosts
struct EncodedOps {
    std::vector<qsv_op> ops;
    std::vector<std::vector<uint32_t>> controls;
    std::vector<std::vector<double>> matrices;
    std::vector<std::vector<uint8_t>> masks;
};

inline void encode(const std::vector<Gate>& gates, size_t num_qubits, EncodedOps& out) {
    size_t non_id = 0;
    for (auto& g : gates) non_id += !g.is_id();
    out.ops.reserve(non_id);
    out.controls.reserve(non_id);
    out.matrices.reserve(non_id);
    out.masks.reserve(non_id);
    for (size_t counter = 0; counter < gates.size(); ++counter) {
        const Gate& g = gates[counter];
        if (g.is_id()) continue;
        qsv_op op{};
        op.kind = g.kind;
        op.target = (uint32_t)(counter % num_qubits);
        op.n_controls = (uint32_t)g.controls.size();
        out.controls.push_back(g.controls);
        op.controls = out.controls.back().empty() ? nullptr : out.controls.back().data();
        op.param = g.param;
        op.iparam = g.iparam;
        if (g.kind == QSV_GATE_CUSTOM) {
            const size_t k = g.controls.size() + 1, dim = (size_t)1 << k;
            // wide gates (multi-controlled gates such as multicnot::<N>, tests/grovers.rs:157-172) go over as compact
            // columns: one 2^k column per sub-state the closure answered for (qsv.h, iparam = 1)
            const bool compact = k >= kCompactCustomWires;
            out.matrices.emplace_back(compact ? 0 : 2 * dim * dim, 0.0);
            out.masks.emplace_back(dim, 0);
            auto& m = out.matrices.back();
            auto& none = out.masks.back();
            size_t n_cols = 0;
            for (size_t s = 0; s < dim; ++s) {
                std::optional<SuperPosition> image = g.func(ProductState::binary_basis(s, k));
                if (!image) { none[s] = 1; continue; }
                if (image->get_dimension() != dim)
                    throw QuantrError("The custom gate '" + g.name + "' returned a superposition of the wrong dimension.");
                if (compact) {
                    if (++n_cols > 64)
                        throw QuantrError("The custom gate '" + g.name + "' acts on " + std::to_string(k) + " wires and answers for more than 64 basis states; dense Custom gates are limited to 13 wires.");
                    for (size_t t = 0; t < dim; ++t) { m.push_back(image->amplitudes[t].real()); m.push_back(image->amplitudes[t].imag()); }
                    continue;
                }
                for (size_t t = 0; t < dim; ++t) {
                    m[(t * dim + s) * 2] = image->amplitudes[t].real();
                    m[(t * dim + s) * 2 + 1] = image->amplitudes[t].imag();
                }
            }
            if (compact) op.iparam = 1;
            if (m.empty()) m.push_back(0.0);  // a closure that answers None everywhere: no column, but a non-NULL pointer
            op.matrix = m.data();
            op.none_mask = none.data();
        }
        out.ops.push_back(op);
    }
}

}  // namespace detail

inline std::optional<states::ProductState> states::SuperPosition::measure() const { return measure(next_f64()); }

// src/simulated_circuit.rs:20-188 with the register held in HBM
class SimulatedCircuit {
public:
    SimulatedCircuit(std::vector<Gate> gates, size_t n, std::unique_ptr<detail::StateHandle> st, bool progress)
        : circuit_gates(std::move(gates)), num_qubits(n), config_progress(progress), state_(std::move(st)) {}

    Measurement<std::map<ProductState, size_t>> measure_all(size_t shots) {  // :63-73
        bool any_custom = false;
        for (auto& g : circuit_gates) any_custom |= g.is_custom_gate();
        if (any_custom && !disable_warnings)
            fprintf(stderr, "\x1b[93m[Quantr Warning] Custom gates were detected in the circuit. Measurements will be taken from a cached register in memory, "
                            "and so if the Custom gate does NOT implement a unitary mapping, the measure_all method will most likely lead to wrong results. "
                            "To simulate a circuit without cache, see SimulatedCircuit::measure_all_without_cache.\x1b[0m\n");
        std::vector<double> u(shots);
        for (auto& x : u) x = next_f64();  // one dice roll per shot, in shot order (super_positions.rs:334)
        std::vector<uint64_t> idx(shots);
        int rc = qsv_sample(state_->h, u.data(), shots, idx.data());
        if (rc != QSV_OK) detail::ffi_panic(state_->h, rc, "qsv_sample");
        std::map<ProductState, size_t> bins;
        for (uint64_t i : idx) {
            if (i == UINT64_MAX) {  // add_to_bin, :116-130
                if (!disable_warnings)
                    fprintf(stderr, "\x1b[93m[Quantr Warning] The superposition failed to collapse to a state during repeat measurements. This is likely "
                                    "due to the use of Gate::Custom where the mapping is not unitary.\x1b[0m\n");
                continue;
            }
            bins[ProductState::binary_basis(i, num_qubits)] += 1;
        }
        return {Measurement<std::map<ProductState, size_t>>::Observable, std::move(bins)};
    }

    Measurement<const SuperPosition*> get_state() {  // :158-160
        if (!host_) {
            host_ = std::make_unique<SuperPosition>();
            host_->product_dim = num_qubits;
            host_->amplitudes.resize((size_t)1 << num_qubits);
            int rc = qsv_download(state_->h, reinterpret_cast<double*>(host_->amplitudes.data()), 0, (uint64_t)1 << num_qubits);
            if (rc != QSV_OK) detail::ffi_panic(state_->h, rc, "qsv_download");
        }
        return {Measurement<const SuperPosition*>::NonObservable, host_.get()};
    }
    Measurement<SuperPosition> take_state() {  // :185-187
        SuperPosition s = *get_state().value;
        return {Measurement<SuperPosition>::NonObservable, std::move(s)};
    }
    void print_warnings(bool printing) { disable_warnings = printing; }  // :163-165 (as upstream: sets disable_warnings = printing)
    const std::vector<Gate>& get_circuit_gates() const { return circuit_gates; }
    size_t get_num_qubits() const { return num_qubits; }
    void set_print_progress(bool p) { config_progress = p; }
    qsv_state* device_handle() { return state_->h; }  // escape hatch: range downloads / gathers for large registers

    std::vector<Gate> circuit_gates;
    size_t num_qubits;
    bool config_progress;
    bool disable_warnings = false;
    qsv_stats stats{};

private:
    std::unique_ptr<detail::StateHandle> state_;
    std::unique_ptr<SuperPosition> host_;
};

// src/circuit.rs:27-474
class Circuit {
public:
    explicit Circuit(size_t n) : num_qubits(n) {  // Circuit::new, :48
        if (n == 0) throw QuantrError("The initialised circuit must have at least one wire.");
    }
    size_t get_num_qubits() const { return num_qubits; }
    void set_print_progress(bool p) { config_progress = p; }
    const std::vector<Gate>& get_gates() const { return circuit_gates; }

    Circuit& add_gate(const Gate& g, size_t position) { return add_gates_with_positions({{position, g}}); }  // :124

    Circuit& add_gates_with_positions(const std::map<size_t, Gate>& gp) {  // :152-188
        for (auto& kv : gp)
            if (kv.first >= num_qubits)
                throw QuantrError("The position, " + std::to_string(kv.first) + ", is out of bounds for the circuit with " + std::to_string(num_qubits) + " qubits.");
        std::vector<Gate> col(num_qubits, Gate::Id());
        for (auto& kv : gp) col[kv.first] = kv.second;
        has_overlapping_controls_and_target(col);
        push_multi_gates(col);
        circuit_gates.insert(circuit_gates.end(), col.begin(), col.end());
        return *this;
    }

    Circuit& add_gates(const std::vector<Gate>& gates) {  // :210-226
        if (gates.size() != num_qubits)
            throw QuantrError("The number of gates, " + std::to_string(gates.size()) + ", does not match the number of wires, " + std::to_string(num_qubits) +
                              ". All wires must have gates added.");
        has_overlapping_controls_and_target(gates);
        std::vector<Gate> col = gates;
        push_multi_gates(col);
        circuit_gates.insert(circuit_gates.end(), col.begin(), col.end());
        return *this;
    }

    Circuit& add_repeating_gate(const Gate& g, const std::vector<size_t>& positions) {  // :324-341
        std::vector<bool> seen(num_qubits, false);
        for (size_t p : positions) {
            if (p >= num_qubits) throw QuantrError("The position, " + std::to_string(p) + ", is out of bounds for the circuit with " + std::to_string(num_qubits) + " qubits.");
            if (seen[p]) throw QuantrError("Attempted to add more than one gate onto a single wire. The positions must all differ.");
            seen[p] = true;
        }
        std::vector<Gate> col(num_qubits, Gate::Id());
        for (size_t p : positions) col[p] = g;
        return add_gates(col);
    }

    Circuit& change_register(const SuperPosition& sp) {  // :463-473
        if (sp.product_dim != num_qubits)
            throw QuantrError("The custom register has a product state dimension of " + std::to_string(sp.product_dim) + ", while the number of qubits is " +
                              std::to_string(num_qubits) + ". These must equal each other.");
        register_ = sp;
        return *this;
    }
    Circuit& change_register(const ProductState& p) { return change_register(SuperPosition::from(p)); }

    // :364-388 — consumes the circuit in the reference; here the gate list is moved out
    SimulatedCircuit simulate() {
        std::optional<SuperPosition> reg = std::move(register_);
        register_.reset();
        std::vector<Gate> gates = std::move(circuit_gates);
        circuit_gates.clear();
        return run(std::move(gates), reg);
    }
    SimulatedCircuit clone_and_simulate() const { return run(circuit_gates, register_); }  // :411-435

private:
    SimulatedCircuit run(std::vector<Gate> gates, const std::optional<SuperPosition>& reg) const {
        detail::EncodedOps enc;
        detail::encode(gates, num_qubits, enc);
        if (config_progress) {  // simulation.rs:32-34,45-47,183-200
            printf("Starting circuit simulation...\n");
            for (size_t c = 0; c < gates.size(); ++c)
                if (!gates[c].is_id()) {
                    printf("Applying gate kind %u on wire %zu # %zu/%zu \n", gates[c].kind, c % num_qubits, c + 1, gates.size());
                    if (c + 1 == gates.size()) printf("Finished circuit simulation.\n");
                }
        }
        auto st = std::make_unique<detail::StateHandle>((uint32_t)num_qubits);
        int rc = reg ? qsv_upload(st->h, reinterpret_cast<const double*>(reg->amplitudes.data()), 0, reg->amplitudes.size())
                     : qsv_init_basis(st->h, 0);  // SuperPosition::new_unchecked, super_positions_unchecked.rs:39-46
        if (rc != QSV_OK) detail::ffi_panic(st->h, rc, "register set-up");
        qsv_stats stats{};
        rc = qsv_apply(st->h, enc.ops.data(), enc.ops.size(), &stats);
        if (rc != QSV_OK) detail::ffi_panic(st->h, rc, "qsv_apply");
        SimulatedCircuit sim(std::move(gates), num_qubits, std::move(st), config_progress);
        sim.stats = stats;
        return sim;
    }

    static void push_multi_gates(std::vector<Gate>& gates) {  // :230-270
        size_t non_id = 0;
        for (auto& g : gates) {
            if (g.is_custom_gate())
                for (unsigned char ch : g.name)
                    if (ch >= 0x80)
                        throw QuantrError("The custom function name, " + g.name + ", does not only use ASCII chars. This could lead to problems in printing "
                                          "the circuit diagram. This warning will be promoted to an Error in the next major release.");
            non_id += !g.is_id();
        }
        if (non_id < 2) return;
        std::vector<Gate> extended;
        const size_t n = gates.size();
        for (size_t pos = 0; pos < n; ++pos)
            if (!gates[pos].is_single_gate()) {
                std::vector<Gate> col(n, Gate::Id());
                col[pos] = gates[pos];
                extended.insert(extended.end(), col.begin(), col.end());
                gates[pos] = Gate::Id();
            }
        gates.insert(gates.end(), extended.begin(), extended.end());
    }

    void has_overlapping_controls_and_target(const std::vector<Gate>& gates) const {  // :272-293
        for (size_t pos = 0; pos < gates.size(); ++pos) {
            const Gate& g = gates[pos];
            if (g.is_single_gate()) continue;
            std::vector<bool> seen(num_qubits, false);
            for (uint32_t node : g.controls) {
                if (node >= num_qubits)
                    throw QuantrError("The control node at position " + std::to_string(node) + ", is greater than the umnber of qubits " + std::to_string(num_qubits) + ".");
                if (seen[node]) throw QuantrError("The gate has overlapping control nodes.");
                seen[node] = true;
                if (node == pos) throw QuantrError("The gate has a control node that equals the gate's position " + std::to_string(pos) + ".");
            }
        }
    }

    std::vector<Gate> circuit_gates;
    size_t num_qubits;
    std::optional<SuperPosition> register_;
    bool config_progress = false;
};

}  // namespace quantr
