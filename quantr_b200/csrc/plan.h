// plan.h — host-side lowering and fusion scheduler (no CUDA here).
//
// Takes the gate list exactly as Circuit::simulate_with_register walks it
// (src/circuit/simulation.rs:37-56) and produces pass blobs (qsv_types.h) for the kernels.
#pragma once
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/qsv.h"
#include "qsv_types.h"

namespace qsv {

constexpr int kHostPrefixBits = 14;      // folded prefixes up to this many local bits are applied on the host (a 256 KiB table per rank)
constexpr int kMaxPrefixLocalBits = 26;  // ... wider ones run on a sub-register on the device (1 GiB at most)

// A gate lowered to physical-bit space (bit = n-1-wire).
struct LOp {
    enum Kind { MAT, DIAG, DENSE } kind = MAT;
    uint32_t src_gate = 0;
    // MAT: 2x2 matrix on `target`, applied where all bits of cmask are 1
    int target = -1;
    uint64_t cmask = 0;
    OpType mtype = OP_MAT_GENERAL;
    double m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // MAT, dual: `m` is applied where all bits of cmask are 1 and `m2` everywhere else (a controlled gate multiplied
    // together with its uncontrolled neighbours on the target: U2.X^c.U1 = U2.U1 or U2.X.U1, merge_single_qubit_gates)
    bool dual = false;
    double m2[8] = {1, 0, 0, 0, 0, 0, 1, 0};
    // DIAG: amp *= exp(i*pi*(theta0 + sum coef*bit)) where all bits of cmask are 1 (half-turns)
    double theta0 = 0;
    std::vector<std::pair<int, double>> lin;
    // DENSE: CSR over sub-index (bits[0] is the MSB of the sub-index)
    std::vector<int> bits;
    std::vector<uint32_t> rowptr, cols;
    std::vector<cplx> vals;

    uint64_t support() const;  // bits the op's action depends on or changes
    uint64_t targets() const;  // bits that must be tile bits
};

struct PlanOptions {
    int tile_bits = 0;   // 0 = auto: 11 for registers the pipelined kernel serves (>= 2^23 local amplitudes), else 12
    int low_bits = 0;     // contiguous low index bits kept in every tile; 0 = chosen per circuit by the cost model
    int fuse = 1;
    int direct_store = 1; // last round stores registers straight to global memory when that stays coalesced
    int qft4 = 1;         // whole QFT-ladder rounds become one radix-16 macro-op (pass_core.h qft4_apply)
    int big_low_pass = 1; // a pass over the contiguous low index bits may use a 2^12 tile next to 2^11 strided passes
    int fold_prefix = 1;  // basis states: leading gates on the top qubits (rank id + up to kMaxPrefixLocalBits local ones) are applied on the host (build_plan)
    int prefix_subregister = 1;  // prefixes wider than kHostPrefixBits local qubits run on a device sub-register (state_api.cu); 0: stop at kHostPrefixBits
    int prefix_keep_bits = 16;   // ... and never over the lowest prefix_keep_bits local qubits: the first pass computes the tiles that hold
                                 // amplitudes outside its pipeline, so they must stay a small fraction (tests lower it)
    int prefix_min_local = 23;  // ... local qubits are folded only on registers of at least this many local qubits (small ones: not worth a table upload per run)
    int reorder = 1;      // passes take later ops that commute with the ops they had to leave behind (plan.cpp schedule)
    int merge_1q = 1;     // 2x2 gates on the same target and controls are multiplied together across commuting ops; identities vanish
    int perm_rounds = 1;  // runs of X / CNot / Toffoli gates on more than four targets become one gather through the tile (ROUND_PERM)
    int merge_ctrl = 1;   // ... and a controlled 2x2 gate absorbs its uncontrolled neighbours on the target (dual-matrix ops)
};

// One step of a plan: a fused pass over the local shard, or a global-qubit remap that swaps the index bits held in
// the rank id with `n_global` local physical bits (pairwise amplitude exchange between ranks).
struct PlanStep {
    enum Kind { PASS = 0, EXCHANGE = 1 } kind = PASS;
    uint32_t pass_index = 0;             // PASS: index into Plan::passes
    std::vector<uint8_t> partner_bits;   // EXCHANGE: local physical bit swapped with rank bit j, ascending
};

struct Plan {
    uint32_t n_qubits = 0;        // logical qubits of the circuit
    uint32_t n_local = 0;         // logical index bits held by this rank
    uint32_t n_alloc = 0;         // allocated local bits (>= kMinQubits)
    PlanOptions opt;
    std::vector<LOp> lops;                        // after lowering + diagonal merging, in LOGICAL bit space (bit = n-1-wire)
    std::vector<std::vector<uint8_t>> passes;     // device blobs (physical bit space)
    std::vector<PlanStep> steps;                  // execution order
    // layout[b] = physical position of logical index bit b (positions >= n_local live in the rank id)
    std::vector<uint8_t> initial_layout, final_layout;
    bool free_initial_layout = false;             // the scheduler chose initial_layout (register must be a basis state)
    // Sharded plans on a basis state: the leading lowered ops whose targets all lie in the rank id act on a product
    // state (one amplitude per rank, all at the same local index), so they are applied on the host to the 2^g-vector
    // of rank amplitudes when the plan runs (prefix_amplitudes) instead of forcing a global-qubit remap.  Logical bit
    // space; empty = none.  A plan with a prefix needs a register that is a basis state.
    std::vector<LOp> prefix;
    // ... and, on registers of prefix_min_local qubits or more, on the top `prefix_local_bits` LOCAL index bits as well: the
    // state after the prefix has 2^(g + prefix_local_bits) non-zero amplitudes - every combination of those top bits, the
    // other bits as in the basis state - which the host computes (prefix_amplitudes) and the first pass synthesises
    // (PassInit::amp_tbl) or a scatter kernel writes.  QFT-33 from a basis state: 14 of its 33 stages cost nothing.
    uint32_t prefix_local_bits = 0;
    uint64_t stamp = 0;  // unique per built plan (caches keyed by a plan's address tell a new plan at an old address apart)
    uint64_t n_gates = 0, n_rounds = 0;
    // device residency (owned by the state API)
    void* dev_blob = nullptr;
    std::vector<size_t> dev_offsets;
    std::vector<size_t> dev_tbl_offsets;  // external-phase tables of pass i inside dev_blob (SIZE_MAX: the pass has none)
    int dev_device = -1;
    int dev_rank = -1;                     // the tables carry the rank bits of the handle that uploaded the plan
};

// Throws std::runtime_error with a message on invalid input / unsupported circuits.
void lower_gates(uint32_t n_qubits, const qsv_op* ops, size_t n_ops, std::vector<LOp>& out, uint64_t* n_gates);
void merge_diagonals(std::vector<LOp>& lops);
void merge_single_qubit_gates(std::vector<LOp>& lops, bool merge_ctrl = true);
// initial_layout: nullptr = identity; free_layout: let the scheduler choose the initial layout (sharded basis states).
void build_plan(Plan& plan, uint32_t n_qubits, uint32_t n_local, const qsv_op* ops, size_t n_ops, const PlanOptions& opt,
                const uint8_t* initial_layout = nullptr, bool free_layout = false);
std::string describe_plan(const Plan& plan);
// The 2^(g + prefix_local_bits) amplitudes after the plan's prefix has been applied to the basis state `basis_index`
// (canonical index): entry j belongs to the physical index whose top g + prefix_local_bits bits spell j (rank id first)
// and whose other bits are the basis state's.  Without a prefix: 1 on the rank that holds the basis state (2^g entries).
void prefix_amplitudes(const Plan& plan, uint64_t basis_index, std::vector<cplx>& out);
// The prefix as a plan over the g + prefix_local_bits support qubits, for the basis state `basis_index` (its low bits are
// folded into the ops as constants); the sub-register's final state equals prefix_amplitudes' table.
void build_prefix_subplan(const Plan& plan, uint64_t basis_index, Plan& sub);

// Pipelined exchange (state_api.cu run_overlapped): an EXCHANGE step can run slice by slice against the pass before it
// and the pass after it when some local index bits are neither partner bits of the exchange nor tile bits of those
// passes: the passes then run once per slice (the amplitudes whose slice bits spell v) and slice v is exchanged while the
// passes work on other slices.
struct OverlapGroup {
    bool slice_prev = false, slice_next = false;  // the neighbouring PASS steps that run slice by slice
    uint32_t n_bits = 0;                          // log2 of the number of slices
    uint8_t bits[3] = {0, 0, 0};                  // slice bits (local physical index bits), ascending
};
// step: index of an EXCHANGE step; sliceable[i] != 0: the PASS of step i may run over a slice (pipelined kernel, not the
// fused initialisation, not already part of another group).  log2_slices: 1..3.  False when nothing can overlap.
bool plan_overlap_group(const Plan& plan, size_t step, const std::vector<char>& sliceable, uint32_t log2_slices, OverlapGroup& out);

void host_sincospi(double x, double* s, double* c);

}  // namespace qsv
