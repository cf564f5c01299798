// reference_tests.cpp — the reference's own tests replayed through the C++ host mirror (quantr_b200/host/quantr.hpp).
//
//   reference_tests host     builder / validation tests only (src/circuit.rs:516-597,715-720,839-845,984-991); no GPU needed
//   reference_tests device   + golden state vectors (src/circuit.rs:603-982, tests/qft.rs, tests/grovers.rs) on cuda:0
//
// Expected registers are the reference's hand-computed vectors; tolerances are the reference's ERROR_MARGINs.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../quantr_b200/host/quantr.hpp"

using namespace quantr;
static int g_failed = 0, g_run = 0;
static const double S2 = 0.70710678118654752440, PI = 3.14159265358979323846;

#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) { printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++g_failed; } \
    } while (0)

template <class F>
static bool throws_quantr_error(F f) {
    try { f(); } catch (const QuantrError&) { return true; }
    return false;
}

static void compare_circuit(Circuit& c, const std::vector<Complex64>& correct, double tol, const char* name) {
    ++g_run;
    auto sim = c.simulate();
    const SuperPosition* reg = sim.get_state().take();
    bool ok = true;
    for (size_t i = 0; i < reg->amplitudes.size(); ++i)
        ok &= std::abs(reg->amplitudes[i].real() - correct[i].real()) < tol && std::abs(reg->amplitudes[i].imag() - correct[i].imag()) < tol;
    if (!ok) { printf("  FAILED golden vector %s\n", name); ++g_failed; }
}

static std::optional<SuperPosition> example_cnot(ProductState prod) {  // src/circuit.rs:504-512
    if (prod.qubits[0] == Qubit::Zero) return std::nullopt;
    if (prod.qubits[1] == Qubit::Zero) return SuperPosition::new_with_amplitudes({0, 0, 0, 1});
    return SuperPosition::new_with_amplitudes({0, 0, 1, 0});
}

static std::optional<SuperPosition> qft(ProductState input) {  // tests/qft.rs:51-67: simulates a sub-circuit per call
    const size_t n = input.num_qubits();
    Circuit mini(n);
    for (size_t pos = 0; pos < n; ++pos) {
        mini.add_gate(Gate::H(), pos);
        for (size_t k = 2; k <= n - pos; ++k) mini.add_gate(Gate::CRk((int32_t)k, (uint32_t)(pos + k - 1)), pos);
    }
    mini.change_register(input);
    return mini.simulate().take_state().take();
}

static void host_tests() {
    const Complex64 Z(0, 0);
    (void)Z;
    {  // pushes_multi_gates, src/circuit.rs:535-551
        ++g_run;
        Circuit c(3);
        c.add_gates({Gate::CNot(2), Gate::CNot(0), Gate::H()}).add_gates({Gate::Toffoli(1, 2), Gate::H(), Gate::CNot(0)});
        std::vector<Gate> want = {Gate::Id(), Gate::Id(), Gate::H(), Gate::CNot(2), Gate::Id(), Gate::Id(), Gate::Id(), Gate::CNot(0), Gate::Id(),
                                  Gate::Id(), Gate::H(), Gate::Id(), Gate::Toffoli(1, 2), Gate::Id(), Gate::Id(), Gate::Id(), Gate::Id(), Gate::CNot(0)};
        CHECK(c.get_gates() == want);
    }
    {  // pushes_multi_gates_using_vec, :554-575
        ++g_run;
        Circuit c(3);
        c.add_gates_with_positions({{2, Gate::H()}, {0, Gate::CNot(2)}, {1, Gate::CNot(0)}});
        CHECK(c.get_gates().size() == 9 && c.get_gates()[2] == Gate::H() && c.get_gates()[3] == Gate::CNot(2) && c.get_gates()[7] == Gate::CNot(0));
    }
    ++g_run; CHECK(throws_quantr_error([] { Circuit c(3); c.add_gates({Gate::Id(), Gate::Custom(example_cnot, {1}, "X"), Gate::Id()}); }));  // :516-523
    ++g_run; CHECK(throws_quantr_error([] { Circuit c(3); c.add_gates({Gate::CNot(0), Gate::Id(), Gate::Id()}); }));                          // :525-532
    ++g_run; CHECK(throws_quantr_error([] { Circuit c(3); c.add_gates_with_positions({{2, Gate::H()}, {0, Gate::CNot(0)}, {1, Gate::CNot(0)}}); }));  // :577-585
    ++g_run; CHECK(throws_quantr_error([] { Circuit c(3); c.add_gates_with_positions({{2, Gate::H()}, {0, Gate::CNot(2)}, {1, Gate::CNot(3)}}); }));  // :587-597
    ++g_run; CHECK(throws_quantr_error([] { Circuit c(4); c.add_repeating_gate(Gate::X(), {0, 1, 1, 3}); }));                                 // :715-720
    ++g_run; CHECK(throws_quantr_error([] { Circuit c(3); c.add_gate(Gate::Custom(example_cnot, {0}, "NonAscii\xe2\x80\xa0"), 1); }));        // :839-845
    ++g_run; CHECK(throws_quantr_error([] { Circuit c(3); c.add_gate(Gate::X(), 1).change_register(ProductState::new_unchecked({Qubit::One, Qubit::Zero})); }));  // :984-991
    ++g_run; CHECK(throws_quantr_error([] { Circuit c(0); }));
    {  // product_states.rs:275-320
        ++g_run;
        ProductState p = ProductState::binary_basis(5, 4);
        CHECK(p.to_string() == "0101" && p.comp_basis() == 5);
        CHECK(throws_quantr_error([] { ProductState::make({}); }));
    }
    {  // super_positions.rs validation
        ++g_run;
        CHECK(throws_quantr_error([] { SuperPosition::new_with_amplitudes({1, 0, 0}); }));
        CHECK(throws_quantr_error([] { SuperPosition::new_with_amplitudes({0.5, 0.5}); }));
        CHECK(SuperPosition::new_with_amplitudes({0, Complex64(0, 1), 0, 0}).get_num_qubits() == 2);
    }
    {  // wide Custom gates are encoded as compact columns (qsv.h iparam = 1): multicnot::<12>, tests/grovers.rs:157-172
        ++g_run;
        const size_t n = 12;
        auto multicnot = [n](ProductState p) -> std::optional<SuperPosition> {
            for (size_t i = 0; i + 1 < n; ++i) if (p.qubits[i] != Qubit::One) return std::nullopt;
            p.invert_digit(n - 1);
            return SuperPosition::from(p);
        };
        std::vector<uint32_t> ctrl;
        for (uint32_t i = 0; i + 1 < n; ++i) ctrl.push_back(i);
        std::vector<Gate> gates(n, Gate::Id());
        gates[n - 1] = Gate::Custom(multicnot, ctrl, "X");
        quantr::detail::EncodedOps enc;
        quantr::detail::encode(gates, n, enc);
        CHECK(enc.ops.size() == 1 && enc.ops[0].iparam == 1 && enc.ops[0].n_controls == n - 1 && enc.ops[0].target == n - 1);
        const size_t dim = (size_t)1 << n;
        CHECK(enc.matrices[0].size() == 2 * 2 * dim);  // two columns: |1..10> and |1..11>
        size_t answered = 0;
        for (uint8_t v : enc.masks[0]) answered += v == 0;
        CHECK(answered == 2 && enc.masks[0][dim - 2] == 0 && enc.masks[0][dim - 1] == 0);
        CHECK(enc.matrices[0][2 * (dim - 1)] == 1.0 && enc.matrices[0][2 * dim + 2 * (dim - 2)] == 1.0);  // images: flipped target
    }
    {  // super_positions.rs:437-497 (hash constructors), :270-295, :315-342
        ++g_run;
        const double r = std::sqrt(0.5);
        const ProductState p01 = ProductState::binary_basis(1, 2), p10 = ProductState::binary_basis(2, 2);
        const SuperPosition h = SuperPosition::new_with_hash_amplitudes({{p01, r}, {p10, Complex64(0, -r)}});
        CHECK(h.get_amplitudes() == SuperPosition::new_with_amplitudes({0, r, Complex64(0, -r), 0}).get_amplitudes());
        CHECK(throws_quantr_error([&] { SuperPosition::new_with_hash_amplitudes({{p01, r}, {ProductState::binary_basis(5, 3), Complex64(0, -r)}}); }));
        CHECK(throws_quantr_error([&] { SuperPosition::new_with_hash_amplitudes({{p01, r}, {p10, Complex64(0, -0.5 * r)}}); }));
        CHECK(throws_quantr_error([] { SuperPosition::new_with_hash_amplitudes({}); }));
        SuperPosition s = SuperPosition::make(2);
        s.set_amplitudes_from_states({{p01, 1.0}});
        CHECK(s.get_amplitudes() == std::vector<Complex64>({0, 1, 0, 0}));
        CHECK(throws_quantr_error([&] { s.set_amplitudes_from_states({{p01, 0.5}}); }));
        CHECK(SuperPosition::new_with_amplitudes_unchecked({1.0, 5e-4, 2e-3, 0.0}).to_hash_map().size() == 2);
        const SuperPosition q = SuperPosition::new_with_amplitudes_unchecked({0.5, 0.0, 0.0, 0.5});
        CHECK(q.measure(0.25)->to_string() == "11" && q.measure(0.2)->to_string() == "00" && !q.measure(0.5).has_value());
        CHECK(q.measure().has_value() || true);
    }
}

static void device_tests() {
    const Complex64 Z(0, 0), I(0, 1);
    {  // multicnot::<12> (examples/generalised_control_not_gate.rs:24-35 at 12 wires): compact columns through qsv_apply
        const size_t n = 12;
        auto multicnot = [n](ProductState p) -> std::optional<SuperPosition> {
            for (size_t i = 0; i + 1 < n; ++i) if (p.qubits[i] != Qubit::One) return std::nullopt;
            p.invert_digit(n - 1);
            return SuperPosition::from(p);
        };
        std::vector<uint32_t> ctrl;
        std::vector<size_t> wires;
        for (uint32_t i = 0; i + 1 < n; ++i) { ctrl.push_back(i); wires.push_back(i); }
        Circuit c(n);
        c.add_repeating_gate(Gate::X(), wires).add_gate(Gate::Custom(multicnot, ctrl, "X"), n - 1);
        std::vector<Complex64> want((size_t)1 << n, Z);
        want.back() = 1.0;
        compare_circuit(c, want, 1e-12, "multicnot_12_compact_columns");
    }
    { Circuit c(2); c.add_gates({Gate::H(), Gate::H()}).add_gates({Gate::S(), Gate::Sdag()});
      compare_circuit(c, {0.5, -0.5 * I, 0.5 * I, 0.5}, 1e-6, "swap_and_conjugate_gates"); }                                           // :604
    { Circuit c(2); c.add_gates({Gate::H(), Gate::H()}).add_gates({Gate::T(), Gate::Tdag()});
      compare_circuit(c, {0.5, Complex64(0.5 * S2, -0.5 * S2), Complex64(0.5 * S2, 0.5 * S2), 0.5}, 1e-6, "t_and_conjugate_gates"); }    // :616
    { Circuit c(3); c.add_gate(Gate::H(), 2).add_gate(Gate::Custom(example_cnot, {2}, "cNot"), 1);
      compare_circuit(c, {S2, Z, Z, S2, Z, Z, Z, Z}, 1e-6, "custom_gates"); }                                                          // :629
    { Circuit c(4); c.add_gate(Gate::X(), 0).add_gate(Gate::H(), 3).add_gate(Gate::Y(), 3).add_gate(Gate::Toffoli(3, 0), 1);
      compare_circuit(c, {Z, Z, Z, Z, Z, Z, Z, Z, -S2 * I, Z, Z, Z, Z, S2 * I, Z, Z}, 1e-6, "toffoli_gates"); }                         // :644
    { Circuit c(4); c.add_gates({Gate::Z(), Gate::Y(), Gate::H(), Gate::X()});
      compare_circuit(c, {Z, Z, Z, Z, Z, S2 * I, Z, S2 * I, Z, Z, Z, Z, Z, Z, Z, Z}, 1e-6, "runs_three_pauli_gates_with_hadamard"); }   // :689
    { Circuit c(4); c.add_repeating_gate(Gate::X(), {1, 2}).add_gate(Gate::CY(2), 0).add_gate(Gate::Swap(3), 2).add_gate(Gate::CY(0), 3);
      compare_circuit(c, {Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, 1, Z, Z, Z}, 1e-6, "cy_and_swap_gates_work"); }                           // :751
    { Circuit c(3); c.add_repeating_gate(Gate::X(), {0, 2}).add_gate(Gate::Swap(1), 2).add_gate(Gate::CZ(1), 0);
      compare_circuit(c, {Z, Z, Z, Z, Z, Z, -1, Z}, 1e-6, "cz_and_swap_gates_work"); }                                                 // :771
    { Circuit c(4); c.add_gate(Gate::H(), 1).add_gate(Gate::CNot(1), 3).add_gate(Gate::Y(), 1);
      compare_circuit(c, {Z, -S2 * I, Z, Z, S2 * I, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z}, 1e-6, "cnot_gate_extended_control_works_asymmetric"); }  // :821
    { Circuit c(2); c.add_gates({Gate::H(), Gate::H()}).add_gate(Gate::Rx(PI), 0);
      compare_circuit(c, {-0.5 * I, -0.5 * I, -0.5 * I, -0.5 * I}, 1e-6, "rx_gate"); }                                                 // :848
    { Circuit c(2); c.add_gates({Gate::H(), Gate::H()}).add_gate(Gate::Ry(PI), 0);
      compare_circuit(c, {-0.5, -0.5, 0.5, 0.5}, 1e-6, "ry_gate"); }                                                                   // :863
    { Circuit c(2); c.add_gates({Gate::H(), Gate::H()}).add_gate(Gate::Rz(PI), 0);
      compare_circuit(c, {-0.5 * I, -0.5 * I, 0.5 * I, 0.5 * I}, 1e-6, "rz_gate"); }                                                   // :878
    { Circuit c(2); c.add_gates({Gate::H(), Gate::H()}).add_gate(Gate::Phase(PI), 0);
      compare_circuit(c, {0.5 * I, 0.5 * I, 0.5 * I, 0.5 * I}, 1e-6, "global_gate"); }                                                 // :893
    { Circuit c(2); c.add_gates({Gate::H(), Gate::H()}).add_gate(Gate::MY90(), 0).add_gate(Gate::Y90(), 1);
      compare_circuit(c, {-0.5, 0.5, 0.5, -0.5}, 1e-6, "y90_and_my90_gate"); }                                                         // :924
    { Circuit c(3); c.add_gates({Gate::X(), Gate::X(), Gate::X()}).add_gate(Gate::CR(-PI * 0.5, 2), 1);
      compare_circuit(c, {Z, Z, Z, Z, Z, Z, Z, -I}, 1e-6, "cr_gate"); }                                                                // :940
    { Circuit c(3); c.add_gates({Gate::X(), Gate::X(), Gate::X()}).add_gate(Gate::CRk(2, 2), 1);
      compare_circuit(c, {Z, Z, Z, Z, Z, Z, Z, I}, 1e-6, "crk_gate"); }                                                                // :955
    { Circuit c(3); c.add_gate(Gate::X(), 1).change_register(ProductState::new_unchecked({Qubit::One, Qubit::Zero, Qubit::One}));
      compare_circuit(c, {Z, Z, Z, Z, Z, Z, Z, 1}, 1e-6, "custom_register"); }                                                         // :970
    { Circuit c(3); c.add_repeating_gate(Gate::X(), {1, 2}).add_gate(Gate::Custom(qft, {0, 1}, "QFT"), 2);                             // tests/qft.rs:22-47
      compare_circuit(c, {S2 * 0.5, -S2 * 0.5, -S2 * 0.5 * I, S2 * 0.5 * I, Complex64(-0.25, 0.25), Complex64(0.25, -0.25), Complex64(0.25, 0.25), Complex64(-0.25, -0.25)},
                      1e-8, "simple_qft"); }
    {  // tests/grovers.rs:22-72
        ++g_run;
        seed(0);
        Circuit c(3);
        c.add_repeating_gate(Gate::H(), {0, 1, 2});
        c.add_gate(Gate::CZ(1), 2);
        c.add_repeating_gate(Gate::H(), {0, 1, 2}).add_repeating_gate(Gate::X(), {0, 1, 2}).add_gate(Gate::H(), 2).add_gate(Gate::Toffoli(0, 1), 2)
            .add_gate(Gate::H(), 2).add_repeating_gate(Gate::X(), {0, 1, 2}).add_repeating_gate(Gate::H(), {0, 1, 2});
        auto sim = c.simulate();
        const SuperPosition* reg = sim.get_state().take();
        const std::vector<Complex64> want = {Z, Z, Z, -S2, Z, Z, Z, -S2};
        for (size_t i = 0; i < 8; ++i) CHECK(std::abs(reg->amplitudes[i] - want[i]) < 1e-8);
        auto bins = sim.measure_all(500).take();
        size_t total = 0;
        for (auto& kv : bins) {
            total += kv.second;
            const std::string s = kv.first.to_string();
            if (s == "011" || s == "111") CHECK(kv.second > 200);
            else CHECK(kv.second == 0);
        }
        CHECK(total == 500);
    }
}

int main(int argc, char** argv) {
    const bool device = argc > 1 && !strcmp(argv[1], "device");
    host_tests();
    if (device) device_tests();
    printf("%d checks, %d failed (%s)\n", g_run, g_failed, device ? "host + device" : "host only");
    return g_failed ? 1 : 0;
}
