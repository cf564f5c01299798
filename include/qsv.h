/*
 * qsv.h — C ABI of the B200-native state-vector engine behind quantr's
 * `Circuit::simulate` hot path.
 *
 * The reference (a-barlow/quantr v0.6.0, pure Rust) has no FFI of its own; the
 * seam this library replaces is Rust-internal.  Each entry point below names the
 * reference interface it stands in for (paths relative to the reference root):
 *
 *   qsv_create / qsv_destroy     SuperPosition storage owned by SimulatedCircuit
 *                                (src/simulated_circuit.rs:20-27, src/circuit/states/super_positions.rs:22-25)
 *   qsv_init_basis               SuperPosition::new_unchecked  (src/circuit/states/super_positions_unchecked.rs:39-46)
 *                                and ProductState -> SuperPosition registers (src/circuit.rs:463-473)
 *   qsv_upload                   Circuit::change_register       (src/circuit.rs:463-473)
 *   qsv_download / qsv_gather    SimulatedCircuit::get_state / take_state (src/simulated_circuit.rs:158-160,185-187)
 *   qsv_apply                    Circuit::simulate_with_register + Circuit::apply_gate
 *                                (src/circuit/simulation.rs:21-57, 64-135) over the whole gate list
 *   qsv_plan_* / qsv_run_plan    the same, split into "lower + schedule" and "launch" so a
 *                                circuit can be re-run with its schedule resident on the device
 *                                (SimulatedCircuit::measure_all_without_cache re-simulates per shot,
 *                                src/simulated_circuit.rs:81-114)
 *   qsv_sample                   SuperPosition::measure x shots  (src/circuit/states/super_positions.rs:332-342,
 *                                src/simulated_circuit.rs:63-73,116-130)
 *   qsv_norm_sqr                 the probability-conservation sums (src/circuit/states/super_positions.rs:236-248)
 *
 * Conventions
 *   - wire q of an n-qubit circuit is bit (n-1-q) of the amplitude index
 *     (src/circuit/states/product_states.rs:205-215): wire 0 is the MSB.
 *   - amplitudes are complex f64, interleaved (re, im), 16 bytes each, canonical index order.
 *   - every function returns 0 on success or a QSV_ERR_* code; the message is
 *     available from qsv_last_error().  No C++ exception crosses this boundary.
 *   - pointers passed in are borrowed for the duration of the call only.
 *   - a handle is not internally synchronised: one call at a time per handle.
 *   - there is NO CPU fallback: without a usable CUDA device qsv_create fails.
 */
#ifndef QSV_H_
#define QSV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QSV_VERSION 100

/* error codes */
enum {
    QSV_OK = 0,
    QSV_ERR_INVALID_ARG = 1,
    QSV_ERR_OUT_OF_MEMORY = 2,
    QSV_ERR_CUDA = 3,
    QSV_ERR_NCCL = 4,
    QSV_ERR_UNSUPPORTED = 5,
    QSV_ERR_INTERNAL = 6
};

/* Gate kinds: the 24 non-Id variants of `enum Gate` (src/circuit/gate.rs:19-106),
 * in declaration order.  QSV_GATE_ID is accepted and skipped, as in
 * src/circuit/simulation.rs:38-41. */
enum {
    QSV_GATE_ID = 0,
    QSV_GATE_H = 1,
    QSV_GATE_X = 2,
    QSV_GATE_Y = 3,
    QSV_GATE_Z = 4,
    QSV_GATE_S = 5,
    QSV_GATE_SDAG = 6,
    QSV_GATE_T = 7,
    QSV_GATE_TDAG = 8,
    QSV_GATE_RX = 9,      /* param = angle */
    QSV_GATE_RY = 10,     /* param = angle */
    QSV_GATE_RZ = 11,     /* param = angle */
    QSV_GATE_X90 = 12,
    QSV_GATE_Y90 = 13,
    QSV_GATE_MX90 = 14,
    QSV_GATE_MY90 = 15,
    QSV_GATE_PHASE = 16,  /* param = angle; exp(i*angle/2) * Identity */
    QSV_GATE_CR = 17,     /* param = angle, controls[0] */
    QSV_GATE_CRK = 18,    /* iparam = k,    controls[0] */
    QSV_GATE_CZ = 19,     /* controls[0] */
    QSV_GATE_CY = 20,     /* controls[0] */
    QSV_GATE_CNOT = 21,   /* controls[0] */
    QSV_GATE_SWAP = 22,   /* controls[0] */
    QSV_GATE_TOFFOLI = 23,/* controls[0], controls[1] */
    QSV_GATE_CUSTOM = 24, /* controls[0..n_controls), matrix, none_mask */
    QSV_GATE_KIND_COUNT = 25
};

/* One non-Id gate as it reaches Circuit::apply_gate: the POD image of
 * `GateInfo { cat_gate: GateCategory, position }` (src/circuit/gate.rs:243-259).
 *
 * Custom gates: the host evaluates the user closure on the 2^k basis states of
 * the sub-register [controls..., target] (k = n_controls + 1, first control =
 * MSB of the sub-index; src/circuit/simulation.rs:137-156) and passes
 *   matrix    : 2^k x 2^k complex f64, row-major, interleaved;
 *               matrix[(t * 2^k + s) * 2 + {0,1}] = amplitude of output sub-state t
 *               in the image of input sub-state s (one closure result per column).
 *   none_mask : 2^k bytes; non-zero where the closure returned None
 *               ("leave this basis state untouched", src/circuit/simulation.rs:120-133).
 *               May be NULL (no None results).  Column s of `matrix` is ignored
 *               where none_mask[s] != 0.
 * Dense Custom gates are applied by a small-dense-gate round and limited to 13 wires.  A gate
 * that is the identity except on one basis sub-state, or on one pair of sub-states that differ
 * in a single wire (multi-controlled gates: the reference's multicnot::<N>, tests/grovers.rs:157-172),
 * is recognised and lowered to controlled ops of the fused pass instead; for these the limit is
 * 30 wires, and beyond 13 wires the host passes compact columns:
 *   iparam = 1: `matrix` holds one column of 2^k complex f64 per sub-state with none_mask == 0, in
 *               ascending sub-state order (at most 64 of them); none_mask is then mandatory.
 */
typedef struct qsv_op {
    uint32_t kind;            /* QSV_GATE_* */
    uint32_t target;          /* wire the gate sits on (`position`) */
    uint32_t n_controls;
    uint32_t reserved;        /* must be 0 */
    const uint32_t* controls; /* wires, reference order */
    double param;
    int32_t iparam;
    int32_t reserved2;        /* must be 0 */
    const double* matrix;
    const uint8_t* none_mask;
} qsv_op;

/* Filled by qsv_apply / qsv_run_plan when non-NULL. */
typedef struct qsv_stats {
    uint64_t n_gates;          /* non-Id gates consumed */
    uint64_t n_passes;         /* fused passes over the local state */
    uint64_t n_rounds;         /* register rounds summed over passes */
    uint64_t n_kernel_launches;/* kernels of this library launched by the call */
    uint64_t bytes_per_pass;   /* algorithmic bytes of one pass: 32 * 2^n_local */
    uint64_t n_exchanges;      /* global-qubit remaps (sharded handles only) */
    uint64_t exchange_bytes;   /* bytes sent per rank over all remaps */
    double device_ms;          /* CUDA-event time of all passes (0 unless timing enabled) */
    double exchange_ms;        /* CUDA-event time of the remaps */
} qsv_stats;

typedef struct qsv_state qsv_state; /* opaque */
typedef struct qsv_plan qsv_plan;   /* opaque */

/* ---- lifetime ---------------------------------------------------------- */

/* Allocates a 2^n_qubits complex-f64 state on CUDA device `device` (HBM) and
 * sets it to |0...0>.  Fails (QSV_ERR_CUDA) if no device is usable. */
int qsv_create(qsv_state** out, uint32_t n_qubits, int device);

/* Sharded state: this process owns the 2^(n_qubits - log2(world)) amplitudes
 * whose top log2(world) index bits equal `rank`.  `nccl_unique_id` is the
 * 128-byte ncclUniqueId produced by qsv_nccl_unique_id() on rank 0 and
 * distributed by the caller.  Collective: every rank must call it.
 * Limits (QSV_ERR_UNSUPPORTED from qsv_apply / qsv_plan_create otherwise): a shard
 * holds at least 4 qubits, and a remap swaps ALL log2(world) rank bits with local
 * bits, so a gate that moves amplitudes across w wires (w = 1 for every standard
 * gate but Swap, 2 for Swap, all wires of a dense Custom gate) needs
 * w + log2(world) <= local qubits ("cannot bring its qubits onto one rank").
 * Controls and diagonal gates on rank-held qubits never need a remap. */
int qsv_create_sharded(qsv_state** out, uint32_t n_qubits, int device, int rank, int world,
                       const void* nccl_unique_id, size_t nccl_unique_id_bytes);
int qsv_nccl_unique_id(void* out, size_t out_bytes);

/* In-library multi-GPU ("single-process, multi-device inside the library, invisible to the caller"; the
 * reference's Circuit::simulate has no notion of devices, src/circuit.rs:364-388): ONE handle for a register sharded
 * over n_devices GPUs of this process (a power of two, at most 16; the devices must reach each other's memory).  Every
 * call below works on it as on a single-GPU handle - qsv_upload / qsv_download / qsv_gather take any range of the
 * register, qsv_apply runs the circuit on all devices with the global-qubit remaps going over NVLink peer memory,
 * qsv_sample / qsv_norm_sqr see the whole register, qsv_save / qsv_load use one file per shard (<path>.r<k>).  Not
 * available on it: qsv_run_plan (a plan object lives on one device; qsv_apply caches its plans per shard),
 * qsv_device_pointer, qsv_peer_export / qsv_peer_import.  Inside, the library keeps one sharded handle and one host thread
 * per device.  n_devices = 1 is qsv_create.  qsv_get_info "devices" returns the number of GPUs behind a handle. */
int qsv_create_multi(qsv_state** out, uint32_t n_qubits, const int* devices, int n_devices);

/* Optional, sharded handles: direct NVLink exchange.  Every rank exports an opaque 64-byte handle of its shard
 * (qsv_peer_export), the caller all-gathers them, and every rank imports the table (qsv_peer_import: `handles` holds
 * world x 64 bytes, entry r = rank r's export).  Afterwards global-qubit remaps swap amplitudes in place through
 * peer-mapped memory (one kernel per peer, loads/stores over NVLink) instead of NCCL send/recv through staging.
 * Every rank of the communicator must be in the same mode when a remap runs: if any rank's import fails, all ranks
 * call qsv_peer_import(s, NULL, 0), which drops the mappings and returns to the NCCL transport.  Teardown order (CUDA
 * IPC): every rank drops its mappings (qsv_peer_import(s, NULL, 0) or qsv_destroy) before any rank's shard is freed -
 * drop, synchronise the ranks, then destroy. */
#define QSV_PEER_HANDLE_BYTES 64
int qsv_peer_export(qsv_state* s, void* out_handle, size_t out_bytes);
int qsv_peer_import(qsv_state* s, const void* handles, size_t n_handles);

int qsv_destroy(qsv_state* s);

/* Thread-local message of the last failed call made with `s` (or with no
 * handle, e.g. a failed qsv_create, when s == NULL). */
const char* qsv_last_error(const qsv_state* s);

/* ---- options ----------------------------------------------------------- */

/* key: "tile_bits" (4..13), "low_bits" (contiguous low index bits kept in every
 * tile; 0 = chosen per circuit by the scheduler's cost model), "timing" (0/1: record CUDA-event times in qsv_stats),
 * "fuse" (0: one pass per gate, 1: fused passes).
 * Sharded handles with peer-mapped shards (qsv_peer_import): "overlap" (default 1: a global-qubit remap runs slice by
 * slice on a second stream while the passes before and after it work on the other slices; 0: every remap on its own),
 * "exchange_slices_log2" (1..3, default 2: 2^k slices), "exchange_sms" (SMs left to the swap kernels while a pass runs
 * next to them, default 32).  With "timing" set, remaps run on their own so that every step can be timed.
 * qsv_get_info also answers "overlapped_exchanges": the remaps of the last plan run that were pipelined. */
int qsv_set_option(qsv_state* s, const char* key, int64_t value);
int qsv_get_info(const qsv_state* s, const char* key, int64_t* value);

/* ---- register access --------------------------------------------------- */

/* amp[index] = 1, everything else 0 (index is a canonical, i.e. global, index). */
int qsv_init_basis(qsv_state* s, uint64_t index);

/* Copies `count` amplitudes starting at canonical index `first` from / to host
 * memory (interleaved f64).  On a sharded handle the range must lie inside the
 * rank's shard.  Exception: after a plan with global-qubit remaps (qsv_get_layout is
 * not the identity) qsv_download undoes the qubit permutation on the device and is
 * collective - every rank passes the same range, which may be any part of the register
 * and arrives on every rank. */
int qsv_upload(qsv_state* s, const double* host_amps, uint64_t first, uint64_t count);
int qsv_download(qsv_state* s, double* host_amps, uint64_t first, uint64_t count);

/* host_amps[2*i..2*i+1] = amplitude at canonical index indices[i]. */
int qsv_gather(qsv_state* s, const uint64_t* indices, uint64_t count, double* host_amps);

/* ---- simulation -------------------------------------------------------- */

/* Applies ops[0..n_ops) in order (Id entries skipped), exactly as
 * Circuit::simulate_with_register walks the gate list. */
int qsv_apply(qsv_state* s, const qsv_op* ops, size_t n_ops, qsv_stats* stats);

/* Lowering + scheduling only (host work, needs no GPU).  `n_local_qubits` is
 * the shard size (== n_qubits when unsharded). */
int qsv_plan_create(qsv_plan** out, uint32_t n_qubits, uint32_t n_local_qubits,
                    const qsv_op* ops, size_t n_ops, uint32_t tile_bits, uint32_t low_bits,
                    int fuse);
/* Same, for sharded registers.  `layout[b]` = physical position of logical index bit b (bit b of the canonical
 * amplitude index; positions >= n_local_qubits live in the rank id); NULL = identity.  With `free_layout` != 0 the
 * scheduler chooses the initial layout itself (valid when the register is a basis state, which has no data to move):
 * it parks the qubits that are targeted last in the rank id.  Gates that target a qubit held in the rank id make the
 * plan insert a global-qubit remap (QSV_STEP_EXCHANGE) before them.  `free_layout` also tells the scheduler that the
 * plan will start from a basis state that has not been written to HBM yet (qsv_init_basis is lazy), so its first pass is
 * write-only (the initialisation is fused into it) and is sized accordingly; qsv_apply passes it whenever that holds. */
int qsv_plan_create_ex(qsv_plan** out, uint32_t n_qubits, uint32_t n_local_qubits, const qsv_op* ops, size_t n_ops,
                       uint32_t tile_bits, uint32_t low_bits, int fuse, const uint8_t* layout, int free_layout);
int qsv_plan_destroy(qsv_plan* p);
/* Plans built with `free_layout`: the amplitudes the plan starts from when the register is the basis state
 * `basis_index`.  The scheduler folds the circuit's leading gates on the top qubits - the ones held in the rank id and, on
 * registers of 2^23 amplitudes per rank or more, up to `n_local - 16` local ones - into the initial state: they act on a
 * product state, so they are applied (on every run) to its 2^(g+k) non-zero amplitudes only, and the plan's first pass
 * starts from those.  out[2j], out[2j+1] = re, im of the amplitude at the physical index whose top g + k bits spell j
 * (rank id first) and whose other bits are the basis state's; cap >= 2^(g+k); k = "prefix_local_bits" of
 * qsv_plan_serialize.  Without a folded prefix: one entry per rank, 1 on the rank that holds the basis state.  QFT-n on
 * 2^g ranks needs no global-qubit remap because of it.  qsv_run_plan / qsv_apply use it internally; it is exported for
 * callers that drive the ranks themselves. */
int qsv_plan_initial_amplitudes(const qsv_plan* p, uint64_t basis_index, double* out, size_t cap);

/* Plan introspection (host-side tests, sharded drivers).  Step kinds: */
enum { QSV_STEP_PASS = 0, QSV_STEP_EXCHANGE = 1 };
int qsv_plan_num_steps(const qsv_plan* p, size_t* n_steps);
/* kind: QSV_STEP_*; pass_index: for PASS; partner_bits[0..n_global): for EXCHANGE, the local physical bit swapped
 * with rank bit j (cap = capacity of partner_bits). */
int qsv_plan_get_step(const qsv_plan* p, size_t i, int* kind, uint32_t* pass_index, uint8_t* partner_bits, size_t cap);
/* which = 0: layout the plan starts from, 1: layout after the last step.  out_layout holds n_qubits bytes. */
int qsv_plan_get_layout(const qsv_plan* p, int which, uint8_t* out_layout, size_t cap);
int qsv_plan_stats(const qsv_plan* p, qsv_stats* stats);
/* Serialised schedule (for inspection and for the host-side schedule tests):
 * writes up to `cap` bytes, returns the full size in *size. */
int qsv_plan_serialize(const qsv_plan* p, void* out, size_t cap, size_t* size);
const char* qsv_plan_last_error(void);

/* Runs a plan made for this handle's (n_qubits, n_local_qubits).  The first
 * run uploads the schedule to the device; later runs reuse it. */
int qsv_run_plan(qsv_state* s, qsv_plan* p, qsv_stats* stats);

/* ---- measurement ------------------------------------------------------- */

/* For each uniform u in [0,1): the first canonical index i with
 * u < sum_{j<=i} |amp_j|^2, or UINT64_MAX if u >= total ("failed to collapse",
 * src/circuit/states/super_positions.rs:341).  The caller draws the uniforms
 * (fastrand::f64() per shot in the reference, src/circuit/states/super_positions.rs:334).
 * On a register in a remapped qubit layout (sharded plans with global-qubit remaps) the
 * cumulative sums run in the physical order of the amplitudes: the distribution of the
 * returned canonical indices is the same, the map u -> index is a different inverse CDF. */
int qsv_sample(qsv_state* s, const double* uniforms, uint64_t shots, uint64_t* out_indices);

/* sum |amp|^2 over the local state (sharded: over all ranks). */
int qsv_norm_sqr(qsv_state* s, double* out);

/* Current layout of the handle: out_layout[b] = physical position of logical index bit b (identity unless a sharded
 * plan with remaps has run). */
int qsv_get_layout(const qsv_state* s, uint8_t* out_layout, size_t cap);

/* Blocks until all work queued on the handle's stream has finished. */
int qsv_synchronize(qsv_state* s);

/* Device time of every step (fused pass or exchange) of the last qsv_apply / qsv_run_plan, in plan order, in ms.
 * Recorded only while the "timing" option is on (CUDA events around every launch; the host then waits for each step).
 * Writes min(*n_steps, cap) entries; out_ms may be NULL to query the count. */
int qsv_last_step_ms(const qsv_state* s, double* out_ms, size_t cap, size_t* n_steps);

/* State checkpoint (the reference carries a state between circuits by hand: SimulatedCircuit::take_state +
 * Circuit::change_register, src/simulated_circuit.rs:185-187, src/circuit.rs:463-473).  File = 256-byte header
 * ("QSVCKPT1", n_qubits, n_local_qubits, rank, world, the qubit layout) + the rank's amplitudes, raw little-endian
 * interleaved f64 in physical order.  A sharded register is one file per rank (the caller names them); qsv_load needs a
 * handle of the same shape (n_qubits, rank, world) and restores the layout as well. */
int qsv_save(qsv_state* s, const char* path);
int qsv_load(qsv_state* s, const char* path);

/* Raw device pointer / stream of the handle, for callers that time with their
 * own CUDA events or wrap the memory (no ownership transfer). */
int qsv_device_pointer(qsv_state* s, void** dev_ptr, void** cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* QSV_H_ */
