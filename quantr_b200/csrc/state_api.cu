// state_api.cu — the device-resident register (`qsv_state`) and the C ABI around it.
//
// Replaces the `register: SuperPosition` owned by SimulatedCircuit (src/simulated_circuit.rs:20-27)
// with a handle to 2^n complex-f64 amplitudes in HBM.  See include/qsv.h for the reference
// interface each export stands in for.  No exception crosses the boundary.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "kernels.h"
#include "pass_core.h"
#include "plan.h"
#include "plan_handle.h"
#include "shard.h"

using namespace qsv;

struct qsv_state {
    uint32_t n_qubits = 0;  // logical qubits of the circuit
    uint32_t n_local = 0;   // index bits held by this rank
    uint32_t n_alloc = 0;   // allocated local bits (>= kMinQubits)
    int device = 0;
    int rank = 0, world = 1;
    int sm_count = 148;
    cplx* d_state = nullptr;
    cudaStream_t stream = nullptr;
    PlanOptions opt;
    int timing = 0;
    // measurement cache (K6): per-block sums + exclusive scan, valid until the register changes
    double* d_sums = nullptr;
    double* d_prefix = nullptr;
    uint64_t n_blocks = 0;
    uint32_t block_bits = 0;
    bool prefix_valid = false;
    double total_prob = 0.0;
    ShardComm* comm = nullptr;
    // layout[b] = physical position of logical index bit b (positions >= n_local live in the rank id)
    uint8_t layout[64];
    bool layout_identity = true;
    // a freshly initialised register is a basis state that has not been written to HBM yet: the next plan may pick
    // its own initial layout (nothing to move) and the memset is issued just before the first pass
    bool lazy_basis = false;
    uint64_t lazy_index = 0;
    void* d_staging = nullptr;  // exchange staging (sharded handles, NCCL path)
    size_t staging_bytes = 0;
    // scratch for qsv_sample / qsv_gather (uniforms, indices, results): grown on demand, never per call
    void* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    bool total_valid = false;  // total_prob has been read back since the prefix was rebuilt
    // plans of recent qsv_apply calls, keyed by the serialised gate list + options + layout: re-running a circuit
    // (measure_all_without_cache, repeated simulate) skips lowering, scheduling and the schedule upload
    struct CachedPlan {
        std::vector<uint8_t> key;
        qsv_plan* plan = nullptr;
        uint64_t stamp = 0;
    };
    std::vector<CachedPlan> plan_cache;
    uint64_t plan_stamp = 0;
    std::vector<double> last_step_ms;  // "timing" option: device time of every step of the last plan run
    std::vector<cudaEvent_t> tev;      // pooled events of the "timing" option: one before every step + one after the last
    std::vector<cplx*> peer_ptr;  // peer-mapped shards (qsv_peer_import); empty = NCCL send/recv exchange
    // Pipelined exchange (run_overlapped): the remap runs slice by slice on a second stream while the passes next to it
    // work on the other slices.  Cross-GPU hand-shakes are flag words behind the shard (kFlagBytes after the amplitudes,
    // so the peers reach them through the same mapping): ready[v][r] = rank r has finished the pass before the remap on
    // slice v, done[v][r] = rank r has finished writing slice v of this shard; values are the running exchange number.
    cudaStream_t xstream = nullptr;
    std::vector<cudaEvent_t> xev;
    uint32_t xchg_epoch = 0;
    int overlap = 1;         // option "overlap": 0 = every remap runs on its own between the passes
    int xchg_sms = 32;       // SMs left to the swap kernels while a pass runs next to them (option "exchange_sms")
    int xchg_slices_log2 = 2;  // option "exchange_slices_log2": 2^k slices per remap
    uint64_t n_overlapped = 0;   // remaps of the last plan run that were pipelined
    // A folded prefix wider than the host table (Plan::prefix_local_bits > kHostPrefixBits) runs on a sub-register of the
    // support qubits on this device; its final state is the table the plan's first pass synthesises its tiles from.
    qsv_state* sub = nullptr;
    qsv_plan* sub_plan = nullptr;
    const void* sub_plan_of = nullptr;  // the plan and basis state sub_plan was made for
    uint64_t sub_plan_basis = 0, sub_plan_stamp = 0;
    cudaEvent_t sub_done = nullptr;
    bool peers_local = false;    // peer_ptr holds plain device pointers of this process (qsv_create_multi), not IPC mappings
    // In-library multi-GPU (qsv_create_multi): this handle is a front for one sharded handle per device, each driven by its
    // own host thread of `pool`; every call is forwarded to all of them (see the "multi" section below).
    std::vector<qsv_state*> children;
    struct MultiPool* pool = nullptr;
    std::string error;
};

namespace {

thread_local std::string g_global_error;

int set_error(qsv_state* s, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (s) s->error = buf;
    g_global_error = buf;
    return code;
}

#define QSV_CUDA(s, call)                                                                                            \
    do {                                                                                                             \
        cudaError_t e__ = (call);                                                                                    \
        if (e__ != cudaSuccess) {                                                                                    \
            cudaGetLastError();                                                                                      \
            return set_error(s, e__ == cudaErrorMemoryAllocation ? QSV_ERR_OUT_OF_MEMORY : QSV_ERR_CUDA, "%s: %s", #call, \
                             cudaGetErrorString(e__));                                                               \
        }                                                                                                            \
    } while (0)

constexpr size_t kFlagBytes = 4096;     // 2 x 8 slices x 16 ranks x 4 bytes, rounded up
constexpr uint32_t kFlagRanks = 16, kFlagDone = 8 * kFlagRanks;
uint32_t* flag_words(const qsv_state* s, cplx* shard) { return reinterpret_cast<uint32_t*>(shard + (1ull << s->n_alloc)); }
uint64_t local_len(const qsv_state* s) { return 1ull << s->n_local; }
uint64_t rank_base(const qsv_state* s) { return (uint64_t)s->rank << s->n_local; }

// canonical (logical) index -> physical index (rank bits on top) under the handle's layout, and back
uint64_t to_physical(const qsv_state* s, uint64_t logical) {
    if (s->layout_identity) return logical;
    uint64_t p = 0;
    for (uint32_t b = 0; b < s->n_qubits; ++b) p |= ((logical >> b) & 1ull) << s->layout[b];
    return p;
}
uint64_t to_logical(const qsv_state* s, uint64_t physical) {
    if (s->layout_identity) return physical;
    uint64_t l = 0;
    for (uint32_t b = 0; b < s->n_qubits; ++b) l |= ((physical >> s->layout[b]) & 1ull) << b;
    return l;
}
void set_layout(qsv_state* s, const uint8_t* layout) {
    s->layout_identity = true;
    for (uint32_t b = 0; b < s->n_qubits; ++b) {
        s->layout[b] = layout ? layout[b] : (uint8_t)b;
        if (s->layout[b] != b) s->layout_identity = false;
    }
}

// Writes a pending basis state to HBM (under the current layout).
// `amp` (optional): this rank's amplitude at the basis state's local index when a sharded plan folded its leading
// gates into the initial state (Plan::prefix); default: 1 on the rank that holds the basis state.
// `tbl` / `sup_bits` (optional): the plan folded gates on the top sup_bits local qubits as well (Plan::prefix_local_bits);
// tbl = this rank's 2^sup_bits amplitudes in device memory.
int materialize(qsv_state* s, const cplx* amp = nullptr, const cplx* tbl = nullptr, uint32_t sup_bits = 0) {
    if (!s->lazy_basis) return QSV_OK;
    s->lazy_basis = false;
    QSV_CUDA(s, cudaMemsetAsync(s->d_state, 0, sizeof(cplx) << s->n_alloc, s->stream));
    const uint64_t phys = to_physical(s, s->lazy_index);
    if (tbl && sup_bits) {
        const uint64_t low = phys & ((1ull << (s->n_local - sup_bits)) - 1ull);
        QSV_CUDA(s, launch_scatter_prefix(s->d_state, tbl, sup_bits, s->n_local, low, s->stream));
    } else if (amp) {
        if (amp->x != 0.0 || amp->y != 0.0) QSV_CUDA(s, launch_set_amp(s->d_state, phys & (local_len(s) - 1), amp->x, amp->y, s->stream));
    } else if ((phys >> s->n_local) == (uint64_t)s->rank) {
        QSV_CUDA(s, launch_set_amp(s->d_state, phys & (local_len(s) - 1), 1.0, 0.0, s->stream));
    }
    s->prefix_valid = false;
    return QSV_OK;
}

// Device scratch of at least `bytes` (256-byte aligned sub-ranges are carved by the callers).
int ensure_scratch(qsv_state* s, size_t bytes) {
    if (bytes <= s->scratch_bytes) return QSV_OK;
    if (s->d_scratch) {
        QSV_CUDA(s, cudaStreamSynchronize(s->stream));
        cudaFree(s->d_scratch);
        s->d_scratch = nullptr;
        s->scratch_bytes = 0;
    }
    size_t want = 1 << 16;
    while (want < bytes) want <<= 1;
    QSV_CUDA(s, cudaMalloc(&s->d_scratch, want));
    s->scratch_bytes = want;
    return QSV_OK;
}

int create_common(qsv_state** out, uint32_t n_qubits, int device, int rank, int world) {
    if (!out) return set_error(nullptr, QSV_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (n_qubits == 0 || n_qubits > 62) return set_error(nullptr, QSV_ERR_INVALID_ARG, "n_qubits must be in 1..62");
    uint32_t g = 0;
    while ((1 << g) < world) ++g;
    if (world < 1 || (1 << g) != world || rank < 0 || rank >= world) return set_error(nullptr, QSV_ERR_INVALID_ARG, "world must be a power of two and 0 <= rank < world");
    if (g >= n_qubits) return set_error(nullptr, QSV_ERR_INVALID_ARG, "more ranks than amplitudes");
    if (g > 0 && n_qubits - g < (uint32_t)kMinQubits) return set_error(nullptr, QSV_ERR_UNSUPPORTED, "sharded registers need at least %d qubits per rank", kMinQubits);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_error(nullptr, QSV_ERR_CUDA, "no usable CUDA device (%s); quantr_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= count) return set_error(nullptr, QSV_ERR_INVALID_ARG, "device %d out of range (%d devices)", device, count);
    qsv_state* s = new (std::nothrow) qsv_state();
    if (!s) return set_error(nullptr, QSV_ERR_OUT_OF_MEMORY, "host allocation failed");
    s->n_qubits = n_qubits;
    s->n_local = n_qubits - g;
    s->n_alloc = s->n_local < (uint32_t)kMinQubits ? (uint32_t)kMinQubits : s->n_local;
    s->device = device;
    s->rank = rank;
    s->world = world;
    auto fail = [&](int code) { qsv_destroy(s); return code; };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(set_error(nullptr, QSV_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e)));
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(set_error(nullptr, QSV_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)));
    s->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(set_error(nullptr, QSV_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)));
    const size_t bytes = (sizeof(cplx) << s->n_alloc) + kFlagBytes;  // amplitudes + the exchange flags behind them
    if ((e = cudaMalloc(&s->d_state, bytes)) != cudaSuccess) {
        cudaGetLastError();
        return fail(set_error(nullptr, e == cudaErrorMemoryAllocation ? QSV_ERR_OUT_OF_MEMORY : QSV_ERR_CUDA, "cudaMalloc of %zu bytes for a %u-qubit register: %s", bytes, s->n_local, cudaGetErrorString(e)));
    }
    s->block_bits = s->n_alloc < 12 ? s->n_alloc : 12;
    s->n_blocks = 1ull << (s->n_alloc - s->block_bits);
    if ((e = cudaMalloc(&s->d_sums, sizeof(double) * s->n_blocks)) != cudaSuccess || (e = cudaMalloc(&s->d_prefix, sizeof(double) * (s->n_blocks + 2 + 2 * kScanMaxChunks))) != cudaSuccess) {
        cudaGetLastError();
        return fail(set_error(nullptr, QSV_ERR_OUT_OF_MEMORY, "cudaMalloc of the measurement scratch: %s", cudaGetErrorString(e)));
    }
    if ((e = cudaMemset(reinterpret_cast<uint8_t*>(s->d_state) + (sizeof(cplx) << s->n_alloc), 0, kFlagBytes)) != cudaSuccess)
        return fail(set_error(nullptr, QSV_ERR_CUDA, "cudaMemset of the exchange flags: %s", cudaGetErrorString(e)));
    if (const char* env = getenv("QSV_OVERLAP")) s->overlap = atoi(env) != 0;
    if (const char* env = getenv("QSV_XCHG_SMS")) s->xchg_sms = atoi(env);
    if (const char* env = getenv("QSV_XCHG_SLICES_LOG2")) s->xchg_slices_log2 = atoi(env);
    set_layout(s, nullptr);
    *out = s;
    return QSV_OK;
}

int upload_plan(qsv_state* s, qsv_plan* p);

void release_plan_device(qsv_plan* p) {
    if (p->plan.dev_blob) {
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(p->plan.dev_device);
        cudaFree(p->plan.dev_blob);
        if (prev >= 0) cudaSetDevice(prev);
        p->plan.dev_blob = nullptr;
    }
}

int upload_plan(qsv_state* s, qsv_plan* p) {
    Plan& plan = p->plan;
    if (plan.dev_blob && plan.dev_device == s->device && plan.dev_rank == s->rank) return QSV_OK;
    if (plan.dev_blob) release_plan_device(p);
    size_t total = 0;
    plan.dev_offsets.clear();
    plan.dev_tbl_offsets.clear();
    for (auto& b : plan.passes) {
        plan.dev_offsets.push_back(total);
        total += (b.size() + 255) & ~size_t(255);
    }
    if (total == 0) return QSV_OK;
    const size_t blob_bytes = total;
    // external-phase tables of the passes the pipelined kernel will run (two half-index tables per DIAG op whose phase
    // depends on bits outside the tile; filled on the device below)
    for (auto& b : plan.passes) {
        const DevPass& hdr = *reinterpret_cast<const DevPass*>(b.data());
        if (hdr.n_ext_ops && pass_uses_tma(b.data(), s->n_alloc, s->sm_count)) {
            plan.dev_tbl_offsets.push_back(total);
            total += (sizeof(cplx) * hdr.n_ext_ops * ext_table_len(hdr.n_tiles) + 255) & ~size_t(255);
        } else {
            plan.dev_tbl_offsets.push_back(SIZE_MAX);
        }
    }
    std::vector<uint8_t> host(blob_bytes, 0);
    for (size_t i = 0; i < plan.passes.size(); ++i) memcpy(host.data() + plan.dev_offsets[i], plan.passes[i].data(), plan.passes[i].size());
    QSV_CUDA(s, cudaMalloc(&plan.dev_blob, total));
    plan.dev_device = s->device;
    plan.dev_rank = s->rank;
    p->release_device = release_plan_device;
    QSV_CUDA(s, cudaMemcpyAsync(plan.dev_blob, host.data(), blob_bytes, cudaMemcpyHostToDevice, s->stream));
    for (size_t i = 0; i < plan.passes.size(); ++i)
        if (plan.dev_tbl_offsets[i] != SIZE_MAX)
            QSV_CUDA(s, launch_build_ext_tables(static_cast<const uint8_t*>(plan.dev_blob) + plan.dev_offsets[i], plan.passes[i].data(),
                                                reinterpret_cast<cplx*>(static_cast<uint8_t*>(plan.dev_blob) + plan.dev_tbl_offsets[i]), rank_base(s), s->stream));
    QSV_CUDA(s, cudaStreamSynchronize(s->stream));  // `host` goes out of scope
    return QSV_OK;
}

// Global-qubit remap: rank bit j <-> local physical bit partner[j] (ascending).  The amplitudes whose partner bits
// spell peer rank v go to rank v and land where the partner bits spell this rank; the block that spells this rank
// stays.  Contiguous chunks of 2^partner[0] amplitudes, staged through a bounded buffer, grouped ncclSend/ncclRecv.
int run_exchange(qsv_state* s, const PlanStep& st, double* ms_out) {
    if (!s->comm) return set_error(s, QSV_ERR_INTERNAL, "exchange step on an unsharded handle");
    const uint32_t g = s->n_qubits - s->n_local;
    if (st.partner_bits.size() != g) return set_error(s, QSV_ERR_INTERNAL, "exchange step does not match the handle");
    if (!s->d_staging && s->peer_ptr.empty()) {
        size_t want = (size_t)1 << 30;  // 1 GiB per direction, double-buffered below
        const size_t shard = sizeof(cplx) << s->n_local;
        if (want > shard / 4) want = shard / 4 ? shard / 4 : sizeof(cplx);
        if (const char* env = getenv("QSV_STAGING_BYTES")) want = (size_t)atoll(env);
        QSV_CUDA(s, cudaMalloc(&s->d_staging, want * 2));
        s->staging_bytes = want;
    }
    (void)ms_out;  // timed by the caller (run_plan_impl records an event before every step)
    std::string err;
    if (!s->peer_ptr.empty()) {
        // peer-memory path: barrier (every rank has finished the passes before the remap), one in-place swap kernel
        // per peer (round-robin pairing), barrier (every peer has finished writing into this shard)
        static const bool trace_exchange = getenv("QSV_TRACE_EXCHANGE") != nullptr;  // developer aid: barrier / swap / barrier times on stderr
        cudaEvent_t te[4] = {nullptr, nullptr, nullptr, nullptr};
        if (trace_exchange) for (auto& e : te) cudaEventCreate(&e);
        if (trace_exchange) cudaEventRecord(te[0], s->stream);
        if (!shard_barrier(s->comm, err)) return set_error(s, QSV_ERR_NCCL, "%s", err.c_str());
        if (trace_exchange) cudaEventRecord(te[1], s->stream);
        for (int step = 1; step < s->world; ++step) {
            const int peer = s->rank ^ step;
            QSV_CUDA(s, launch_peer_swap(s->d_state, s->peer_ptr[peer], s->n_local, st.partner_bits.data(), g, s->rank, peer, s->sm_count, s->stream));
        }
        if (trace_exchange) cudaEventRecord(te[2], s->stream);
        if (!shard_barrier(s->comm, err)) return set_error(s, QSV_ERR_NCCL, "%s", err.c_str());
        if (trace_exchange) {
            cudaEventRecord(te[3], s->stream);
            cudaEventSynchronize(te[3]);
            float a = 0, b = 0, c = 0;
            cudaEventElapsedTime(&a, te[0], te[1]);
            cudaEventElapsedTime(&b, te[1], te[2]);
            cudaEventElapsedTime(&c, te[2], te[3]);
            fprintf(stderr, "[qsv] rank %d remap: barrier %.3f ms, swap %.3f ms, barrier %.3f ms\n", s->rank, a, b, c);
            for (auto& e : te) cudaEventDestroy(e);
        }
    } else if (!shard_exchange_bits(s->comm, s->d_state, s->n_local, st.partner_bits.data(), g, s->d_staging, s->staging_bytes, err)) {
        return set_error(s, QSV_ERR_NCCL, "%s", err.c_str());
    }
    return QSV_OK;
}

// One pass step on the handle's stream, whole or over a slice.
int launch_pass_step(qsv_state* s, const Plan& plan, uint32_t pass_index, const PassInit* init, const PassSlice* slice, int grid_sms) {
    const uint8_t* dev_pass = static_cast<const uint8_t*>(plan.dev_blob) + plan.dev_offsets[pass_index];
    const cplx* ext_tbl = plan.dev_tbl_offsets[pass_index] == SIZE_MAX
                              ? nullptr
                              : reinterpret_cast<const cplx*>(static_cast<const uint8_t*>(plan.dev_blob) + plan.dev_tbl_offsets[pass_index]);
    QSV_CUDA(s, launch_pass(s->d_state, dev_pass, plan.passes[pass_index].data(), ext_tbl, rank_base(s), s->n_alloc, s->sm_count, init, s->stream, slice, grid_sms));
    return QSV_OK;
}

// Pipelined global-qubit remap (peer-memory transport): the shard is cut into 2^k slices along index bits that neither the
// remap nor the passes next to it touch.  The pass before the remap runs slice by slice on the handle's stream; slice v
// is exchanged on the second stream as soon as every rank has finished it (flag words in peer memory, no host round
// trip, no NCCL call); the pass after the remap starts on slice v as soon as every peer has finished writing it.  While
// both kinds of kernels are in flight the persistent pass kernel leaves `xchg_sms` SMs to the swap kernels (one fat
// CTA per SM each, so the two grids fit next to each other whatever the launch order).
//   step: index of the EXCHANGE step; grp: what plan_overlap_group decided.
int run_overlapped(qsv_state* s, const Plan& plan, size_t step, const OverlapGroup& grp) {
    const PlanStep& st = plan.steps[step];
    const uint32_t g = s->n_qubits - s->n_local, n_slices = 1u << grp.n_bits;
    if (!s->xstream) QSV_CUDA(s, cudaStreamCreateWithFlags(&s->xstream, cudaStreamNonBlocking));
    while (s->xev.size() < 2 * (size_t)n_slices) {
        cudaEvent_t ev;
        QSV_CUDA(s, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        s->xev.push_back(ev);
    }
    FlagPeers peers{};
    for (int r = 0; r < s->world && r < (int)kFlagRanks; ++r) peers.flags[r] = r == s->rank ? nullptr : flag_words(s, s->peer_ptr[r]);
    const uint32_t* my_flags = flag_words(s, s->d_state);
    const uint32_t epoch = ++s->xchg_epoch;
    const int pass_sms = s->sm_count - s->xchg_sms, swap_sms = s->xchg_sms > 4 ? s->xchg_sms - 2 : s->xchg_sms;  // two SMs stay free for the flag kernels
    PassSlice sl{};
    sl.n = grp.n_bits;
    for (uint32_t i = 0; i < grp.n_bits; ++i) sl.bit[i] = grp.bits[i];
    for (uint32_t v = 0; v < n_slices; ++v) {
        sl.value = v;
        if (grp.slice_prev) {
            int rc = launch_pass_step(s, plan, plan.steps[step - 1].pass_index, nullptr, &sl, v == 0 ? 0 : pass_sms);  // nothing else runs next to slice 0
            if (rc != QSV_OK) return rc;
        }
        if (grp.slice_prev || v == 0) QSV_CUDA(s, cudaEventRecord(s->xev[v], s->stream));
        QSV_CUDA(s, cudaStreamWaitEvent(s->xstream, s->xev[grp.slice_prev ? v : 0], 0));
        QSV_CUDA(s, launch_flag_signal(peers, s->world, v * kFlagRanks + (uint32_t)s->rank, epoch, s->xstream));
        QSV_CUDA(s, launch_flag_wait(my_flags, s->world, s->rank, v * kFlagRanks, epoch, s->xstream));
        for (int k = 1; k < s->world; ++k) {
            const int peer = s->rank ^ k;
            QSV_CUDA(s, launch_peer_swap(s->d_state, s->peer_ptr[peer], s->n_local, st.partner_bits.data(), g, s->rank, peer, swap_sms, s->xstream, &sl));
        }
        QSV_CUDA(s, launch_flag_signal(peers, s->world, kFlagDone + v * kFlagRanks + (uint32_t)s->rank, epoch, s->xstream));
        QSV_CUDA(s, cudaEventRecord(s->xev[n_slices + v], s->xstream));
    }
    for (uint32_t v = 0; v < n_slices; ++v) {
        sl.value = v;
        QSV_CUDA(s, cudaStreamWaitEvent(s->stream, s->xev[n_slices + v], 0));
        QSV_CUDA(s, launch_flag_wait(my_flags, s->world, s->rank, kFlagDone + v * kFlagRanks, epoch, s->stream));
        if (grp.slice_next) {
            int rc = launch_pass_step(s, plan, plan.steps[step + 1].pass_index, nullptr, &sl, v + 1 == n_slices ? 0 : pass_sms);  // the last slice has the GPU to itself
            if (rc != QSV_OK) return rc;
        }
    }
    return QSV_OK;
}

int run_plan_impl(qsv_state* s, qsv_plan* p, qsv_stats* stats);

// The folded prefix of `p` on a sub-register of the g + k support qubits (plan.cpp build_prefix_subplan): every rank runs the
// whole prefix (2^(g+k) amplitudes, at most 1 GiB) and uses its own 2^k slice.  The sub-plan is kept per (plan, basis state);
// the sub-register is simulated again on every run.  *d_tbl = this rank's slice, ready on s->stream.
int run_prefix_subregister(qsv_state* s, qsv_plan* p, cplx** d_tbl) {
    const Plan& plan = p->plan;
    const uint32_t g = s->n_qubits - s->n_local, k = plan.prefix_local_bits, ns = g + k;
    if (s->sub && s->sub->n_qubits != ns) {
        qsv_destroy(s->sub);
        s->sub = nullptr;
    }
    if (!s->sub) {
        int rc = create_common(&s->sub, ns, s->device, 0, 1);
        if (rc != QSV_OK) return set_error(s, rc, "sub-register of %u qubits for the folded prefix: %s", ns, qsv_last_error(nullptr));
    }
    if (!s->sub_done) QSV_CUDA(s, cudaEventCreateWithFlags(&s->sub_done, cudaEventDisableTiming));
    if (!s->sub_plan || s->sub_plan_of != (const void*)p || s->sub_plan_basis != s->lazy_index || s->sub_plan_stamp != p->plan.stamp) {
        if (s->sub_plan) qsv_plan_destroy(s->sub_plan);
        s->sub_plan = new (std::nothrow) qsv_plan();
        if (!s->sub_plan) return set_error(s, QSV_ERR_OUT_OF_MEMORY, "host allocation failed");
        try {
            build_prefix_subplan(plan, s->lazy_index, s->sub_plan->plan);
        } catch (const std::exception& e) {
            qsv_plan_destroy(s->sub_plan);
            s->sub_plan = nullptr;
            return set_error(s, QSV_ERR_INTERNAL, "prefix sub-plan: %s", e.what());
        }
        s->sub_plan_of = p;
        s->sub_plan_basis = s->lazy_index;
        s->sub_plan_stamp = p->plan.stamp;
    }
    const uint32_t nf = s->n_local - k;
    int rc = qsv_init_basis(s->sub, s->lazy_index >> nf);
    if (rc == QSV_OK) rc = run_plan_impl(s->sub, s->sub_plan, nullptr);
    if (rc == QSV_OK) rc = materialize(s->sub);  // (a prefix that lowered to nothing: the basis state itself)
    if (rc != QSV_OK) return set_error(s, rc, "folded prefix on the sub-register: %s", s->sub->error.c_str());
    QSV_CUDA(s, cudaSetDevice(s->device));
    QSV_CUDA(s, cudaEventRecord(s->sub_done, s->sub->stream));
    QSV_CUDA(s, cudaStreamWaitEvent(s->stream, s->sub_done, 0));
    *d_tbl = s->sub->d_state + ((size_t)s->rank << k);
    return QSV_OK;
}

int run_plan_impl(qsv_state* s, qsv_plan* p, qsv_stats* stats) {
    Plan& plan = p->plan;
    if (plan.n_qubits != s->n_qubits || plan.n_local != s->n_local)
        return set_error(s, QSV_ERR_INVALID_ARG, "plan was made for %u/%u qubits, the handle holds %u/%u", plan.n_qubits, plan.n_local, s->n_qubits, s->n_local);
    // layout handshake: a pending basis state adopts the plan's initial layout (nothing to move); otherwise they must agree
    if (s->lazy_basis) {
        set_layout(s, plan.initial_layout.data());
    } else if (memcmp(s->layout, plan.initial_layout.data(), s->n_qubits) != 0) {
        return set_error(s, QSV_ERR_INVALID_ARG, "the plan starts from a different qubit layout than the register is in");
    }
    // A sharded plan may have folded the circuit's leading gates on the rank-id qubits into the initial state: every rank
    // then starts from its own amplitude at the basis state's local index
    cplx rank_amp{1.0, 0.0};
    const bool folded = !plan.prefix.empty();
    const uint32_t sup_bits = folded ? plan.prefix_local_bits : 0u;
    cplx* d_tbl = nullptr;
    int rc = QSV_OK;
    if (folded) {
        if (!s->lazy_basis) return set_error(s, QSV_ERR_INVALID_ARG, "this plan was built for a register that is a basis state (qsv_init_basis)");
        // (QSV_HOST_PREFIX_BITS: tests send narrower prefixes through the sub-register as well)
        static const uint32_t host_prefix_bits = getenv("QSV_HOST_PREFIX_BITS") ? (uint32_t)atoi(getenv("QSV_HOST_PREFIX_BITS")) : (uint32_t)kHostPrefixBits;
        if (sup_bits > host_prefix_bits) {
            rc = run_prefix_subregister(s, p, &d_tbl);
            if (rc != QSV_OK) return rc;
        } else {
        std::vector<cplx> amps;
        prefix_amplitudes(plan, s->lazy_index, amps);
        if (sup_bits == 0) {
            rank_amp = amps[(size_t)s->rank];
        } else {  // this rank's 2^sup_bits amplitudes go to the device: the first pass synthesises its tiles from them
            const size_t count = (size_t)1 << sup_bits;
            rc = ensure_scratch(s, sizeof(cplx) * count);
            if (rc != QSV_OK) return rc;
            d_tbl = static_cast<cplx*>(s->d_scratch);
            QSV_CUDA(s, cudaMemcpyAsync(d_tbl, amps.data() + (size_t)s->rank * count, sizeof(cplx) * count, cudaMemcpyHostToDevice, s->stream));
            QSV_CUDA(s, cudaStreamSynchronize(s->stream));  // `amps` goes out of scope
        }
        }
    }
    // Fused initialisation (pass_kernel_tma.cu): a pending basis state is not written to HBM when the plan's first step is
    // a pass the pipelined kernel runs; that pass synthesises the one tile holding the amplitude and writes every other
    // tile as zeros (mode 2).  QSV_FUSED_INIT=0 turns it off (memset + ordinary first pass), 1 computes every tile.
    static const int fused_init_mode = getenv("QSV_FUSED_INIT") ? atoi(getenv("QSV_FUSED_INIT")) : 2;
    bool fused_init = false;
    PassInit pass_init{};
    if (s->lazy_basis && fused_init_mode > 0 && !plan.steps.empty() && plan.steps[0].kind == PlanStep::PASS && s->n_alloc == s->n_local &&
        pass_init_supported(plan.passes[plan.steps[0].pass_index].data(), s->n_alloc, s->sm_count)) {
        const DevPass& h0 = *reinterpret_cast<const DevPass*>(plan.passes[plan.steps[0].pass_index].data());
        uint64_t phys = to_physical(s, s->lazy_index);
        if (folded) phys = (phys & (local_len(s) - 1)) | rank_base(s);  // every rank holds an amplitude at that local index
        pass_init = make_pass_init(h0, phys, s->n_local, fused_init_mode >= 2 ? 2u : 1u, sup_bits, d_tbl);
        pass_init.amp_re = rank_amp.x;
        pass_init.amp_im = rank_amp.y;
        if (sup_bits == 0 && rank_amp.x == 0.0 && rank_amp.y == 0.0) pass_init.base_full = ~0ull;  // nothing to synthesise: the shard is all zero
        fused_init = true;
        s->lazy_basis = false;
    }
    rc = materialize(s, folded ? &rank_amp : nullptr, d_tbl, sup_bits);
    if (rc != QSV_OK) return rc;
    rc = upload_plan(s, p);
    if (rc != QSV_OK) return rc;
    s->prefix_valid = false;
    static const bool trace_passes = getenv("QSV_TRACE_PASSES") != nullptr;  // developer aid: per-step device times on stderr
    double pass_ms = 0.0, exch_ms = 0.0;
    uint64_t n_exch = 0;
    s->last_step_ms.clear();
    // Which remaps run pipelined against their neighbouring passes (peer-memory transport, no per-step timing): decided
    // left to right; a pass belongs to at most one group.
    std::vector<OverlapGroup> groups(plan.steps.size());
    std::vector<char> in_group(plan.steps.size(), 0);  // PASS steps that run inside run_overlapped
    s->n_overlapped = 0;
    if (s->overlap && !s->peer_ptr.empty() && !s->timing && !trace_passes && s->world <= (int)kFlagRanks && s->n_alloc == s->n_local) {
        std::vector<char> sliceable(plan.steps.size(), 0);
        for (size_t i = 0; i < plan.steps.size(); ++i)
            sliceable[i] = plan.steps[i].kind == PlanStep::PASS && !(i == 0 && fused_init) &&
                           pass_uses_tma(plan.passes[plan.steps[i].pass_index].data(), s->n_alloc, s->sm_count);
        for (size_t i = 0; i < plan.steps.size(); ++i) {
            if (plan.steps[i].kind != PlanStep::EXCHANGE) continue;
            OverlapGroup grp;
            if (!plan_overlap_group(plan, i, sliceable, (uint32_t)s->xchg_slices_log2, grp)) continue;
            groups[i] = grp;
            if (grp.slice_prev) in_group[i - 1] = 1, sliceable[i - 1] = 0;
            if (grp.slice_next) in_group[i + 1] = 1, sliceable[i + 1] = 0;
        }
    }
    const bool timed = s->timing || trace_passes;  // (no pipelined groups then: every step is timed on its own)
    while (timed && s->tev.size() < plan.steps.size() + 1) {
        cudaEvent_t ev;
        QSV_CUDA(s, cudaEventCreate(&ev));
        s->tev.push_back(ev);
    }
    for (size_t i = 0; i < plan.steps.size(); ++i) {
        const PlanStep& st = plan.steps[i];
        if (in_group[i]) continue;  // runs slice by slice inside its remap's group
        if (st.kind == PlanStep::EXCHANGE && groups[i].n_bits) {
            rc = run_overlapped(s, plan, i, groups[i]);
            if (rc != QSV_OK) return rc;
            ++n_exch;
            ++s->n_overlapped;
            continue;
        }
        if (timed) QSV_CUDA(s, cudaEventRecord(s->tev[i], s->stream));
        if (st.kind == PlanStep::PASS) {
            rc = launch_pass_step(s, plan, st.pass_index, (i == 0 && fused_init) ? &pass_init : nullptr, nullptr, 0);
            if (rc != QSV_OK) return rc;
        } else {
            rc = run_exchange(s, st, nullptr);
            if (rc != QSV_OK) return rc;
            ++n_exch;
        }
    }
    if (timed) {
        // "timing": events around every step, read after ONE host synchronisation at the end (the steps run back to back
        // as in an untimed run; round 1 synchronised after every pass)
        QSV_CUDA(s, cudaEventRecord(s->tev[plan.steps.size()], s->stream));
        QSV_CUDA(s, cudaEventSynchronize(s->tev[plan.steps.size()]));
        for (size_t i = 0; i < plan.steps.size(); ++i) {
            const PlanStep& st = plan.steps[i];
            float ms = 0.f;
            QSV_CUDA(s, cudaEventElapsedTime(&ms, s->tev[i], s->tev[i + 1]));
            if (st.kind == PlanStep::PASS) pass_ms += ms; else exch_ms += ms;
            if (s->timing) s->last_step_ms.push_back(ms);
            if (trace_passes && st.kind == PlanStep::PASS) {
                const DevPass& h = *reinterpret_cast<const DevPass*>(plan.passes[st.pass_index].data());
                int run_bits = 0;
                if (h.n_tile_segs && h.tile_segs[0].dst_lo == 0) run_bits = h.tile_segs[0].width;
                fprintf(stderr, "[qsv] pass %u: T=%u run=%uB rounds=%u ops=%u diag=%u flags=%u  %.3f ms  %.0f GB/s\n", st.pass_index, h.tile_bits, 16u << run_bits,
                        h.n_rounds, h.n_ops, h.n_diag, h.flags, ms, 32.0 * (double)(1ull << s->n_alloc) / (ms * 1e-3) / 1e9);
            }
        }
    }
    set_layout(s, plan.final_layout.data());
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->n_gates = plan.n_gates;
        stats->n_passes = plan.passes.size();
        stats->n_rounds = plan.n_rounds;
        stats->n_kernel_launches = plan.passes.size();
        stats->bytes_per_pass = 32ull << s->n_alloc;
        stats->n_exchanges = n_exch;
        const uint32_t g = s->n_qubits - s->n_local;
        stats->exchange_bytes = n_exch * (((16ull << s->n_local) >> g) * ((1ull << g) - 1));
        stats->device_ms = pass_ms;
        stats->exchange_ms = exch_ms;
    }
    return QSV_OK;
}

int ensure_prefix(qsv_state* s) {
    if (s->prefix_valid) return QSV_OK;
    QSV_CUDA(s, launch_prob_block_sums(s->d_state, s->d_sums, s->n_blocks, s->block_bits, s->sm_count, s->stream));
    QSV_CUDA(s, launch_scan_block_sums(s->d_sums, s->d_prefix, s->n_blocks, s->stream));
    s->prefix_valid = true;
    s->total_valid = false;
    return QSV_OK;
}

// total probability of this rank's shard on the host (one 8-byte read-back per rebuilt prefix)
int ensure_total(qsv_state* s) {
    int rc = ensure_prefix(s);
    if (rc != QSV_OK || s->total_valid) return rc;
    QSV_CUDA(s, cudaMemcpyAsync(&s->total_prob, s->d_prefix + s->n_blocks, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    QSV_CUDA(s, cudaStreamSynchronize(s->stream));
    s->total_valid = true;
    return QSV_OK;
}

// qsv_download on a register whose qubits were remapped by a sharded plan: the canonical range is assembled by undoing
// the index-bit permutation on the device (every rank contributes the amplitudes it holds, zeros elsewhere) and summing
// over ranks, chunk by chunk.  Collective: every rank must ask for the same range.
int download_permuted(qsv_state* s, double* host_amps, uint64_t first, uint64_t count) {
    const uint64_t chunk = 1ull << 20;  // 16 MiB of amplitudes per round trip
    int rc = ensure_scratch(s, sizeof(cplx) * (count < chunk ? count : chunk));
    if (rc != QSV_OK) return rc;
    cplx* d_out = static_cast<cplx*>(s->d_scratch);
    for (uint64_t done = 0; done < count; done += chunk) {
        const uint64_t m = count - done < chunk ? count - done : chunk;
        QSV_CUDA(s, launch_gather_range(s->d_state, d_out, first + done, m, s->layout, s->n_qubits, s->n_local, (uint64_t)s->rank, s->sm_count, s->stream));
        if (s->comm) {
            std::string err;
            if (!shard_allreduce_sum_f64(s->comm, reinterpret_cast<double*>(d_out), 2 * m, err)) return set_error(s, QSV_ERR_NCCL, "%s", err.c_str());
        }
        QSV_CUDA(s, cudaMemcpyAsync(host_amps + 2 * done, d_out, sizeof(cplx) * m, cudaMemcpyDeviceToHost, s->stream));
        QSV_CUDA(s, cudaStreamSynchronize(s->stream));
    }
    return QSV_OK;
}

// Serialises everything a plan depends on (qsv_apply's cache key).  Returns false when the key would be too large to be
// worth keeping (huge Custom matrices).
bool plan_cache_key(const qsv_state* s, const qsv_op* ops, size_t n_ops, bool free_layout, std::vector<uint8_t>& key) {
    constexpr size_t kMaxKey = 8u << 20;
    key.clear();
    auto put = [&](const void* p, size_t n) { const uint8_t* b = static_cast<const uint8_t*>(p); key.insert(key.end(), b, b + n); };
    const uint32_t head[6] = {s->n_qubits, s->n_local, (uint32_t)s->opt.tile_bits, (uint32_t)s->opt.low_bits, (uint32_t)s->opt.fuse, free_layout ? 1u : 0u};
    put(head, sizeof(head));
    if (!free_layout) put(s->layout, s->n_qubits);
    for (size_t g = 0; g < n_ops; ++g) {
        const qsv_op& op = ops[g];
        const uint32_t f[3] = {op.kind, op.target, op.n_controls};
        put(f, sizeof(f));
        put(&op.param, sizeof(op.param));
        put(&op.iparam, sizeof(op.iparam));
        if (op.n_controls > 20 || (op.n_controls && !op.controls)) return false;  // invalid: let the scheduler report it
        if (op.n_controls) put(op.controls, sizeof(uint32_t) * op.n_controls);
        if (op.kind == QSV_GATE_CUSTOM) {
            if (!op.matrix) return false;
            const size_t dim = (size_t)1 << (op.n_controls + 1);
            size_t n_cols = dim;
            if (op.iparam == 1) {  // compact columns (qsv.h): one column per sub-state with none_mask == 0
                if (!op.none_mask) return false;
                n_cols = 0;
                for (size_t i = 0; i < dim; ++i) n_cols += op.none_mask[i] == 0;
            }
            if (key.size() + n_cols * dim * 16 + dim > kMaxKey) return false;
            put(op.matrix, n_cols * dim * 16);
            const uint8_t has_none = op.none_mask ? 1 : 0;
            put(&has_none, 1);
            if (op.none_mask) put(op.none_mask, dim);
        }
        if (key.size() > kMaxKey) return false;
    }
    return true;
}

}  // namespace

// ---- in-library multi-GPU ("multi" handles) -------------------------------------------------------------------------
// "single-process, multi-device inside the library, invisible to the caller".  qsv_create_multi builds one
// sharded handle per device (rank r on devices[r], same NCCL communicator, shards mapped into each other by plain peer
// access) and a host thread per device; every call on the front handle runs on all of them at once - exactly what one
// process per GPU would do, so the sharded code paths above (remaps, collectives) are used unchanged.
struct MultiPool {
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    std::function<int(int)> job;
    uint64_t generation = 0;
    int pending = 0;
    bool stop = false;
    std::vector<int> rc;

    explicit MultiPool(int n) : rc((size_t)n, QSV_OK) {
        for (int r = 0; r < n; ++r) threads.emplace_back([this, r] { loop(r); });
    }
    ~MultiPool() {
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
        }
        cv_work.notify_all();
        for (auto& t : threads) t.join();
    }
    void loop(int r) {
        uint64_t seen = 0;
        for (;;) {
            std::function<int(int)> f;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_work.wait(lk, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
                f = job;
            }
            int code;
            try {
                code = f(r);
            } catch (...) {
                code = QSV_ERR_INTERNAL;
            }
            {
                std::lock_guard<std::mutex> lk(m);
                rc[(size_t)r] = code;
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    // runs f(rank) on every worker at the same time; returns the first failing rank's code and its index
    int run(const std::function<int(int)>& f, int* failed = nullptr) {
        std::unique_lock<std::mutex> lk(m);
        job = f;
        pending = (int)threads.size();
        ++generation;
        cv_work.notify_all();
        cv_done.wait(lk, [&] { return pending == 0; });
        for (size_t r = 0; r < rc.size(); ++r)
            if (rc[r] != QSV_OK) {
                if (failed) *failed = (int)r;
                return rc[r];
            }
        return QSV_OK;
    }
};

namespace {

bool is_multi(const qsv_state* s) { return s && !s->children.empty(); }

// forwards a call to every child; the front handle takes over the failing child's message
int multi_run(qsv_state* s, const std::function<int(int)>& f) {
    int failed = -1;
    const int rc = s->pool->run(f, &failed);
    if (rc != QSV_OK && failed >= 0) set_error(s, rc, "device %d: %s", s->children[(size_t)failed]->device, s->children[(size_t)failed]->error.c_str());
    return rc;
}

// maps the siblings' shards into child `s` (same process: plain peer access, no IPC handles)
int attach_local_peers(qsv_state* s, const std::vector<qsv_state*>& all) {
    QSV_CUDA(s, cudaSetDevice(s->device));
    std::vector<cplx*> ptrs(all.size(), nullptr);
    for (size_t r = 0; r < all.size(); ++r) {
        ptrs[r] = all[r]->d_state;
        if ((int)r == s->rank) continue;
        int can = 0;
        QSV_CUDA(s, cudaDeviceCanAccessPeer(&can, s->device, all[r]->device));
        if (!can) return set_error(s, QSV_ERR_UNSUPPORTED, "device %d cannot access device %d's memory", s->device, all[r]->device);
        const cudaError_t e = cudaDeviceEnablePeerAccess(all[r]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return set_error(s, QSV_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", all[r]->device, cudaGetErrorString(e));
        cudaGetLastError();
    }
    s->peer_ptr.swap(ptrs);
    s->peers_local = true;
    return QSV_OK;
}

}  // namespace

#define QSV_ENTER(s)                                                          \
    if (!(s)) return set_error(nullptr, QSV_ERR_INVALID_ARG, "handle is NULL"); \
    {                                                                         \
        cudaError_t e0__ = cudaSetDevice((s)->device);                        \
        if (e0__ != cudaSuccess) return set_error(s, QSV_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e0__)); \
    }

extern "C" {

int qsv_create(qsv_state** out, uint32_t n_qubits, int device) {
    try {
        int rc = create_common(out, n_qubits, device, 0, 1);
        if (rc != QSV_OK) return rc;
        return qsv_init_basis(*out, 0);
    } catch (...) {
        return set_error(nullptr, QSV_ERR_INTERNAL, "unexpected exception in qsv_create");
    }
}

// Second half of creating a sharded handle: joins the communicator (collective over all ranks).
static int attach_comm(qsv_state* s, const void* nccl_unique_id, size_t nccl_unique_id_bytes) {
    QSV_CUDA(s, cudaSetDevice(s->device));
    std::string err;
    s->comm = shard_comm_create(s->rank, s->world, nccl_unique_id, nccl_unique_id_bytes, s->stream, err);
    if (!s->comm) return set_error(s, QSV_ERR_NCCL, "%s", err.c_str());
    // NCCL sets its connections up lazily inside the first collective (hundreds of ms): pay for it here, collectively,
    // instead of inside whichever remap, norm or sample happens to come first
    if (!shard_barrier(s->comm, err) || cudaStreamSynchronize(s->stream) != cudaSuccess)
        return set_error(s, QSV_ERR_NCCL, "first collective on the new communicator failed: %s", err.c_str());
    return qsv_init_basis(s, 0);
}

int qsv_create_sharded(qsv_state** out, uint32_t n_qubits, int device, int rank, int world, const void* nccl_unique_id, size_t nccl_unique_id_bytes) {
    try {
        if (world == 1) return qsv_create(out, n_qubits, device);
        int rc = create_common(out, n_qubits, device, rank, world);
        if (rc != QSV_OK) return rc;
        qsv_state* s = *out;
        rc = attach_comm(s, nccl_unique_id, nccl_unique_id_bytes);
        if (rc != QSV_OK) {
            const std::string msg = s->error;
            qsv_destroy(s);
            *out = nullptr;
            return set_error(nullptr, rc, "%s", msg.c_str());
        }
        return QSV_OK;
    } catch (...) {
        return set_error(nullptr, QSV_ERR_INTERNAL, "unexpected exception in qsv_create_sharded");
    }
}

int qsv_nccl_unique_id(void* out, size_t out_bytes) {
    std::string err;
    if (!shard_unique_id(out, out_bytes, err)) return set_error(nullptr, QSV_ERR_NCCL, "%s", err.c_str());
    return QSV_OK;
}

int qsv_create_multi(qsv_state** out, uint32_t n_qubits, const int* devices, int n_devices) {
    if (!out) return set_error(nullptr, QSV_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (!devices || n_devices < 1) return set_error(nullptr, QSV_ERR_INVALID_ARG, "qsv_create_multi needs at least one device");
    if (n_devices == 1) return qsv_create(out, n_qubits, devices[0]);
    try {
        uint32_t g = 0;
        while ((1 << g) < n_devices) ++g;
        if ((1 << g) != n_devices || n_devices > 16) return set_error(nullptr, QSV_ERR_INVALID_ARG, "the number of devices must be a power of two (at most 16)");
        for (int a = 0; a < n_devices; ++a)
            for (int b = a + 1; b < n_devices; ++b)
                if (devices[a] == devices[b]) return set_error(nullptr, QSV_ERR_INVALID_ARG, "device %d is listed twice", devices[a]);
        uint8_t id[128];
        std::string err;
        if (!shard_unique_id(id, sizeof(id), err)) return set_error(nullptr, QSV_ERR_NCCL, "%s", err.c_str());
        qsv_state* front = new qsv_state();
        front->n_qubits = n_qubits;
        front->n_local = n_qubits > g ? n_qubits - g : 0;
        front->n_alloc = front->n_local;
        front->world = n_devices;
        front->device = devices[0];
        front->children.assign((size_t)n_devices, nullptr);
        front->pool = new MultiPool(n_devices);
        std::vector<std::string> msgs((size_t)n_devices);
        int failed = -1;
        // in two steps, so that a device that cannot even hold its shard does not leave the others waiting inside NCCL
        int rc = front->pool->run([&](int r) {
            qsv_state* c = nullptr;
            const int code = create_common(&c, n_qubits, devices[r], r, n_devices);
            front->children[(size_t)r] = c;
            if (code != QSV_OK) msgs[(size_t)r] = qsv_last_error(nullptr);
            return code;
        }, &failed);
        if (rc == QSV_OK)
            rc = front->pool->run([&](int r) {  // collective: the ranks of one NCCL communicator, created side by side
                const int code = attach_comm(front->children[(size_t)r], id, sizeof(id));
                if (code != QSV_OK) msgs[(size_t)r] = front->children[(size_t)r]->error;
                return code;
            }, &failed);
        if (rc == QSV_OK)
            rc = front->pool->run([&](int r) {
                const int code = attach_local_peers(front->children[(size_t)r], front->children);
                if (code != QSV_OK) msgs[(size_t)r] = front->children[(size_t)r]->error;
                return code;
            }, &failed);
        if (rc != QSV_OK) {
            const std::string msg = failed >= 0 ? msgs[(size_t)failed] : std::string("unknown error");
            qsv_destroy(front);
            return set_error(nullptr, rc, "device %d: %s", failed >= 0 ? devices[failed] : -1, msg.c_str());
        }
        *out = front;
        return QSV_OK;
    } catch (...) {
        return set_error(nullptr, QSV_ERR_INTERNAL, "unexpected exception in qsv_create_multi");
    }
}

int qsv_peer_export(qsv_state* s, void* out_handle, size_t out_bytes) {
    if (is_multi(s)) return set_error(s, QSV_ERR_INVALID_ARG, "a multi-device handle maps its shards itself");
    QSV_ENTER(s);
    if (!out_handle || out_bytes < sizeof(cudaIpcMemHandle_t)) return set_error(s, QSV_ERR_INVALID_ARG, "handle buffer must hold %zu bytes", sizeof(cudaIpcMemHandle_t));
    cudaIpcMemHandle_t h;
    QSV_CUDA(s, cudaIpcGetMemHandle(&h, s->d_state));
    memcpy(out_handle, &h, sizeof(h));
    return QSV_OK;
}

int qsv_peer_import(qsv_state* s, const void* handles, size_t n_handles) {
    if (is_multi(s)) return set_error(s, QSV_ERR_INVALID_ARG, "a multi-device handle maps its shards itself");
    QSV_ENTER(s);
    if (!s->comm) return set_error(s, QSV_ERR_INVALID_ARG, "peer import needs a sharded handle");
    if (n_handles == 0) {  // drop the peer mappings: remaps go through NCCL send/recv again
        QSV_CUDA(s, cudaStreamSynchronize(s->stream));
        for (int r = 0; r < (int)s->peer_ptr.size(); ++r)
            if (r != s->rank && s->peer_ptr[r]) cudaIpcCloseMemHandle(s->peer_ptr[r]);
        s->peer_ptr.clear();
        return QSV_OK;
    }
    if (!s->peer_ptr.empty()) return set_error(s, QSV_ERR_INVALID_ARG, "peers already imported");
    if (!handles || n_handles != (size_t)s->world) return set_error(s, QSV_ERR_INVALID_ARG, "expected one handle per rank");
    std::vector<cplx*> ptrs(s->world, nullptr);
    for (int r = 0; r < s->world; ++r) {
        if (r == s->rank) { ptrs[r] = s->d_state; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < r; ++q) if (q != s->rank && ptrs[q]) cudaIpcCloseMemHandle(ptrs[q]);
            return set_error(s, QSV_ERR_CUDA, "cudaIpcOpenMemHandle for rank %d: %s", r, cudaGetErrorString(e));
        }
        ptrs[r] = static_cast<cplx*>(p);
    }
    s->peer_ptr.swap(ptrs);
    return QSV_OK;
}

int qsv_destroy(qsv_state* s) {
    if (!s) return QSV_OK;
    if (s->sub_plan) qsv_plan_destroy(s->sub_plan);
    if (s->sub) qsv_destroy(s->sub);
    if (s->sub_done) cudaEventDestroy(s->sub_done);
    s->sub_plan = nullptr;
    s->sub = nullptr;
    s->sub_done = nullptr;
    if (s->pool) {  // multi-device front: every shard stops being used by its siblings before any of them is freed
        s->pool->run([&](int r) {
            qsv_state* c = s->children[(size_t)r];
            if (!c) return (int)QSV_OK;
            cudaSetDevice(c->device);
            cudaStreamSynchronize(c->stream);
            if (c->xstream) cudaStreamSynchronize(c->xstream);
            c->peer_ptr.clear();
            return (int)QSV_OK;
        });
        s->pool->run([&](int r) { return qsv_destroy(s->children[(size_t)r]); });
        delete s->pool;
        delete s;
        return QSV_OK;
    }
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (!s->peers_local)
        for (int r = 0; r < (int)s->peer_ptr.size(); ++r)
            if (r != s->rank && s->peer_ptr[r]) cudaIpcCloseMemHandle(s->peer_ptr[r]);
    if (s->comm) shard_comm_destroy(s->comm);
    if (s->d_state) cudaFree(s->d_state);
    if (s->d_sums) cudaFree(s->d_sums);
    if (s->d_prefix) cudaFree(s->d_prefix);
    if (s->d_staging) cudaFree(s->d_staging);
    if (s->d_scratch) cudaFree(s->d_scratch);
    for (auto& e : s->plan_cache) qsv_plan_destroy(e.plan);
    for (cudaEvent_t ev : s->xev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : s->tev) cudaEventDestroy(ev);
    if (s->xstream) cudaStreamDestroy(s->xstream);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return QSV_OK;
}

const char* qsv_last_error(const qsv_state* s) { return s ? s->error.c_str() : g_global_error.c_str(); }

int qsv_set_option(qsv_state* s, const char* key, int64_t value) {
    if (!s || !key) return set_error(s, QSV_ERR_INVALID_ARG, "NULL argument");
    if (is_multi(s)) {
        for (qsv_state* c : s->children) {
            const int rc = qsv_set_option(c, key, value);
            if (rc != QSV_OK) return set_error(s, rc, "%s", c->error.c_str());
        }
        return QSV_OK;
    }
    const std::string k(key);
    if (k == "tile_bits") {
        if (value < kRegBits || value > kMaxTileBits) return set_error(s, QSV_ERR_INVALID_ARG, "tile_bits must be in %d..%d", kRegBits, kMaxTileBits);
        s->opt.tile_bits = (int)value;
    } else if (k == "low_bits") {
        if (value < 0 || value > 8) return set_error(s, QSV_ERR_INVALID_ARG, "low_bits must be in 0..8");
        s->opt.low_bits = (int)value;
    } else if (k == "fuse") {
        s->opt.fuse = value ? 1 : 0;
    } else if (k == "timing") {
        s->timing = value ? 1 : 0;
    } else if (k == "overlap") {
        s->overlap = value ? 1 : 0;
    } else if (k == "exchange_sms") {
        if (value < 4 || value > s->sm_count / 2) return set_error(s, QSV_ERR_INVALID_ARG, "exchange_sms must be in 4..%d", s->sm_count / 2);
        s->xchg_sms = (int)value;
    } else if (k == "exchange_slices_log2") {
        if (value < 1 || value > 3) return set_error(s, QSV_ERR_INVALID_ARG, "exchange_slices_log2 must be in 1..3");
        s->xchg_slices_log2 = (int)value;
    } else {
        return set_error(s, QSV_ERR_INVALID_ARG, "unknown option '%s'", key);
    }
    return QSV_OK;
}

int qsv_get_info(const qsv_state* s, const char* key, int64_t* value) {
    if (!s || !key || !value) return set_error(nullptr, QSV_ERR_INVALID_ARG, "NULL argument");
    const std::string k(key);
    if (k == "devices") { *value = is_multi(s) ? (int64_t)s->children.size() : 1; return QSV_OK; }  // GPUs behind this handle
    if (is_multi(s)) {  // the front of a multi-device handle: one register, spread over `devices` shards
        if (k == "rank") { *value = 0; return QSV_OK; }
        if (k == "world") { *value = 1; return QSV_OK; }
        return qsv_get_info(s->children[0], key, value);
    }
    if (k == "n_qubits") *value = s->n_qubits;
    else if (k == "n_local_qubits") *value = s->n_local;
    else if (k == "n_alloc_qubits") *value = s->n_alloc;
    else if (k == "sm_count") *value = s->sm_count;
    else if (k == "tile_bits") *value = s->opt.tile_bits;
    else if (k == "low_bits") *value = s->opt.low_bits;
    else if (k == "device") *value = s->device;
    else if (k == "rank") *value = s->rank;
    else if (k == "world") *value = s->world;
    else if (k == "overlap") *value = s->overlap;
    else if (k == "exchange_sms") *value = s->xchg_sms;
    else if (k == "exchange_slices_log2") *value = s->xchg_slices_log2;
    else if (k == "overlapped_exchanges") *value = (int64_t)s->n_overlapped;  // remaps of the last plan run that were pipelined against their neighbouring passes
    else return QSV_ERR_INVALID_ARG;
    return QSV_OK;
}

int qsv_init_basis(qsv_state* s, uint64_t index) {
    if (is_multi(s)) {
        for (qsv_state* c : s->children) {
            const int rc = qsv_init_basis(c, index);
            if (rc != QSV_OK) return set_error(s, rc, "%s", c->error.c_str());
        }
        return QSV_OK;
    }
    QSV_ENTER(s);
    if (index >> s->n_qubits) return set_error(s, QSV_ERR_INVALID_ARG, "basis index out of range");
    // lazy: written to HBM right before the first pass (or first read), so a sharded plan can still pick its layout
    s->lazy_basis = true;
    s->lazy_index = index;
    set_layout(s, nullptr);
    s->prefix_valid = false;
    return QSV_OK;
}

int qsv_upload(qsv_state* s, const double* host_amps, uint64_t first, uint64_t count) {
    if (is_multi(s)) {  // any range of the register: every shard takes its part
        if (!host_amps && count) return set_error(s, QSV_ERR_INVALID_ARG, "host_amps is NULL");
        if ((first >> s->n_qubits) || count > (1ull << s->n_qubits) - first) return set_error(s, QSV_ERR_INVALID_ARG, "range is outside the register");
        for (qsv_state* c : s->children) {
            const uint64_t lo = std::max(first, rank_base(c)), hi = std::min(first + count, rank_base(c) + local_len(c));
            if (lo >= hi) continue;
            const int rc = qsv_upload(c, host_amps + 2 * (lo - first), lo, hi - lo);
            if (rc != QSV_OK) return set_error(s, rc, "%s", c->error.c_str());
        }
        return QSV_OK;
    }
    QSV_ENTER(s);
    if (!host_amps && count) return set_error(s, QSV_ERR_INVALID_ARG, "host_amps is NULL");
    if (first < rank_base(s) || first - rank_base(s) + count > local_len(s)) return set_error(s, QSV_ERR_INVALID_ARG, "range is outside this rank's shard");
    if (!s->lazy_basis && !s->layout_identity) return set_error(s, QSV_ERR_UNSUPPORTED, "the register is in a remapped qubit layout: re-initialise it before uploading ranges");
    { int rc0 = materialize(s); if (rc0 != QSV_OK) return rc0; }
    QSV_CUDA(s, cudaMemcpyAsync(s->d_state + (first - rank_base(s)), host_amps, sizeof(cplx) * count, cudaMemcpyHostToDevice, s->stream));
    QSV_CUDA(s, cudaStreamSynchronize(s->stream));
    s->prefix_valid = false;
    return QSV_OK;
}

int qsv_download(qsv_state* s, double* host_amps, uint64_t first, uint64_t count) {
    if (is_multi(s)) {
        if (!host_amps && count) return set_error(s, QSV_ERR_INVALID_ARG, "host_amps is NULL");
        if ((first >> s->n_qubits) || count > (1ull << s->n_qubits) - first) return set_error(s, QSV_ERR_INVALID_ARG, "range is outside the register");
        bool canonical = true;
        for (qsv_state* c : s->children) canonical &= c->lazy_basis || c->layout_identity;
        if (canonical) {  // every shard hands over its part of the range
            for (qsv_state* c : s->children) {
                const uint64_t lo = std::max(first, rank_base(c)), hi = std::min(first + count, rank_base(c) + local_len(c));
                if (lo >= hi) continue;
                const int rc = qsv_download(c, host_amps + 2 * (lo - first), lo, hi - lo);
                if (rc != QSV_OK) return set_error(s, rc, "%s", c->error.c_str());
            }
            return QSV_OK;
        }
        // remapped layout: the collective, un-permuting download, in bounded pieces; rank 0's copy is the answer
        const uint64_t piece = 1ull << 22;
        std::vector<std::vector<double>> scratch(s->children.size());
        for (uint64_t done = 0; done < count; done += piece) {
            const uint64_t m = std::min(piece, count - done);
            const int rc = multi_run(s, [&](int r) {
                double* dst = host_amps + 2 * done;
                if (r != 0) {
                    scratch[(size_t)r].resize(2 * m);
                    dst = scratch[(size_t)r].data();
                }
                return qsv_download(s->children[(size_t)r], dst, first + done, m);
            });
            if (rc != QSV_OK) return rc;
        }
        return QSV_OK;
    }
    QSV_ENTER(s);
    if (!host_amps && count) return set_error(s, QSV_ERR_INVALID_ARG, "host_amps is NULL");
    if (!s->lazy_basis && !s->layout_identity) {
        // remapped layout (a sharded plan with global-qubit remaps has run): any canonical range, collectively
        if ((first >> s->n_qubits) || count > (1ull << s->n_qubits) - first) return set_error(s, QSV_ERR_INVALID_ARG, "range is outside the register");
        return download_permuted(s, host_amps, first, count);
    }
    if (first < rank_base(s) || first - rank_base(s) + count > local_len(s)) return set_error(s, QSV_ERR_INVALID_ARG, "range is outside this rank's shard");
    { int rc0 = materialize(s); if (rc0 != QSV_OK) return rc0; }
    if (!s->layout_identity) return download_permuted(s, host_amps, first, count);
    QSV_CUDA(s, cudaMemcpyAsync(host_amps, s->d_state + (first - rank_base(s)), sizeof(cplx) * count, cudaMemcpyDeviceToHost, s->stream));
    QSV_CUDA(s, cudaStreamSynchronize(s->stream));
    return QSV_OK;
}

int qsv_gather(qsv_state* s, const uint64_t* indices, uint64_t count, double* host_amps) {
    if (is_multi(s)) {  // collective on the shards; rank 0's copy is the answer
        std::vector<std::vector<double>> scratch(s->children.size());
        return multi_run(s, [&](int r) {
            double* dst = host_amps;
            if (r != 0) {
                scratch[(size_t)r].resize(2 * count + 2);
                dst = scratch[(size_t)r].data();
            }
            return qsv_gather(s->children[(size_t)r], indices, count, dst);
        });
    }
    QSV_ENTER(s);
    if (count == 0) return QSV_OK;
    if (!indices || !host_amps) return set_error(s, QSV_ERR_INVALID_ARG, "NULL argument");
    { int rc0 = materialize(s); if (rc0 != QSV_OK) return rc0; }
    try {
        // canonical index -> (rank, local) under the current layout; entries held by other ranks come back as 0 here
        // and are filled in by the sum over ranks below (collective on sharded handles)
        std::vector<uint64_t> local(count);
        std::vector<uint8_t> mine(count);
        for (uint64_t i = 0; i < count; ++i) {
            if (indices[i] >> s->n_qubits) return set_error(s, QSV_ERR_INVALID_ARG, "index %llu out of range", (unsigned long long)indices[i]);
            const uint64_t phys = to_physical(s, indices[i]);
            mine[i] = (phys >> s->n_local) == (uint64_t)s->rank;
            if (!mine[i] && !s->comm) return set_error(s, QSV_ERR_INVALID_ARG, "index %llu is outside this rank's shard", (unsigned long long)indices[i]);
            local[i] = mine[i] ? (phys & (local_len(s) - 1)) : 0;
        }
        const size_t idx_bytes = (sizeof(uint64_t) * count + 255) & ~size_t(255);
        int rc = ensure_scratch(s, idx_bytes + sizeof(cplx) * count);
        if (rc != QSV_OK) return rc;
        uint64_t* d_idx = static_cast<uint64_t*>(s->d_scratch);
        cplx* d_out = reinterpret_cast<cplx*>(static_cast<uint8_t*>(s->d_scratch) + idx_bytes);
        cudaError_t e;
        if ((e = cudaMemcpyAsync(d_idx, local.data(), sizeof(uint64_t) * count, cudaMemcpyHostToDevice, s->stream)) != cudaSuccess ||
            (e = launch_gather(s->d_state, d_idx, d_out, count, s->stream)) != cudaSuccess ||
            (e = cudaMemcpyAsync(host_amps, d_out, sizeof(cplx) * count, cudaMemcpyDeviceToHost, s->stream)) != cudaSuccess ||
            (e = cudaStreamSynchronize(s->stream)) != cudaSuccess)
            rc = set_error(s, QSV_ERR_CUDA, "gather: %s", cudaGetErrorString(e));
        if (rc == QSV_OK && s->comm) {
            for (uint64_t i = 0; i < count; ++i)
                if (!mine[i]) host_amps[2 * i] = host_amps[2 * i + 1] = 0.0;
            std::string err;
            if ((e = cudaMemcpyAsync(d_out, host_amps, sizeof(cplx) * count, cudaMemcpyHostToDevice, s->stream)) != cudaSuccess ||
                !shard_allreduce_sum_f64(s->comm, reinterpret_cast<double*>(d_out), 2 * count, err) ||
                (e = cudaMemcpyAsync(host_amps, d_out, sizeof(cplx) * count, cudaMemcpyDeviceToHost, s->stream)) != cudaSuccess ||
                (e = cudaStreamSynchronize(s->stream)) != cudaSuccess)
                rc = set_error(s, QSV_ERR_NCCL, "gather across ranks: %s", err.empty() ? cudaGetErrorString(e) : err.c_str());
        }
        return rc;
    } catch (const std::bad_alloc&) {
        return set_error(s, QSV_ERR_OUT_OF_MEMORY, "host allocation failed");
    }
}

int qsv_apply(qsv_state* s, const qsv_op* ops, size_t n_ops, qsv_stats* stats) {
    if (is_multi(s)) {  // every device lowers and schedules the same gate list for its shard and runs it; remaps meet over NVLink
        std::vector<qsv_stats> st(s->children.size());
        const int rc = multi_run(s, [&](int r) { return qsv_apply(s->children[(size_t)r], ops, n_ops, &st[(size_t)r]); });
        if (rc == QSV_OK && stats) *stats = st[0];
        return rc;
    }
    QSV_ENTER(s);
    try {
        // a pending basis state has no data to move, so the scheduler may choose the initial layout of a sharded register
        const bool free_layout = s->lazy_basis;  // also tells the scheduler that the plan's first pass will be write-only
        std::vector<uint8_t> key;
        const bool cacheable = plan_cache_key(s, ops, n_ops, free_layout, key);
        qsv_plan* p = nullptr;
        if (cacheable)
            for (auto& e : s->plan_cache)
                if (e.key == key) {
                    e.stamp = ++s->plan_stamp;
                    p = e.plan;
                    break;
                }
        bool owned = false;
        if (!p) {
            int rc = qsv_plan_create_ex(&p, s->n_qubits, s->n_local, ops, n_ops, (uint32_t)s->opt.tile_bits, (uint32_t)s->opt.low_bits, s->opt.fuse,
                                        s->lazy_basis ? nullptr : s->layout, free_layout);
            if (rc != QSV_OK) return set_error(s, rc, "%s", qsv_plan_last_error());
            if (cacheable) {
                constexpr size_t kCacheSlots = 4;
                if (s->plan_cache.size() >= kCacheSlots) {  // evict the least recently used plan (its device copy may still be in use)
                    size_t lru = 0;
                    for (size_t i = 1; i < s->plan_cache.size(); ++i)
                        if (s->plan_cache[i].stamp < s->plan_cache[lru].stamp) lru = i;
                    cudaStreamSynchronize(s->stream);
                    qsv_plan_destroy(s->plan_cache[lru].plan);
                    s->plan_cache.erase(s->plan_cache.begin() + (long)lru);
                }
                qsv_state::CachedPlan e;
                e.key.swap(key);
                e.plan = p;
                e.stamp = ++s->plan_stamp;
                s->plan_cache.push_back(std::move(e));
            } else {
                owned = true;
            }
        }
        int rc = run_plan_impl(s, p, stats);
        if (owned) {
            if (rc == QSV_OK) {
                cudaError_t e = cudaStreamSynchronize(s->stream);  // the plan's device copy is freed below
                if (e != cudaSuccess) rc = set_error(s, QSV_ERR_CUDA, "fused pass failed: %s", cudaGetErrorString(e));
            }
            qsv_plan_destroy(p);
        }
        return rc;
    } catch (const std::bad_alloc&) {
        return set_error(s, QSV_ERR_OUT_OF_MEMORY, "host allocation failed");
    } catch (const std::exception& e) {
        return set_error(s, QSV_ERR_INTERNAL, "%s", e.what());
    } catch (...) {
        return set_error(s, QSV_ERR_INTERNAL, "unexpected exception in qsv_apply");
    }
}

int qsv_run_plan(qsv_state* s, qsv_plan* p, qsv_stats* stats) {
    if (is_multi(s)) return set_error(s, QSV_ERR_UNSUPPORTED, "a plan object lives on one device; use qsv_apply on a multi-device handle (it caches its plans per shard)");
    QSV_ENTER(s);
    if (!p) return set_error(s, QSV_ERR_INVALID_ARG, "plan is NULL");
    try {
        return run_plan_impl(s, p, stats);
    } catch (const std::exception& e) {
        return set_error(s, QSV_ERR_INTERNAL, "%s", e.what());
    } catch (...) {
        return set_error(s, QSV_ERR_INTERNAL, "unexpected exception in qsv_run_plan");
    }
}

int qsv_sample(qsv_state* s, const double* uniforms, uint64_t shots, uint64_t* out_indices) {
    if (is_multi(s)) {  // collective on the shards; rank 0's copy is the answer
        std::vector<std::vector<uint64_t>> scratch(s->children.size());
        return multi_run(s, [&](int r) {
            uint64_t* dst = out_indices;
            if (r != 0) {
                scratch[(size_t)r].resize(shots + 1);
                dst = scratch[(size_t)r].data();
            }
            return qsv_sample(s->children[(size_t)r], uniforms, shots, dst);
        });
    }
    QSV_ENTER(s);
    if (shots == 0) return QSV_OK;
    if (!uniforms || !out_indices) return set_error(s, QSV_ERR_INVALID_ARG, "NULL argument");
    { int rc0 = materialize(s); if (rc0 != QSV_OK) return rc0; }
    int rc = s->comm ? ensure_total(s) : ensure_prefix(s);
    if (rc != QSV_OK) return rc;
    // Sharded: the cumulative sums run over the physical order (rank 0's shard first); every rank receives the same
    // uniforms, answers the shots that fall into its interval and the minimum over ranks assembles the result.
    double offset = 0.0;
    if (s->comm) {
        std::string err;
        std::vector<double> totals(s->world, 0.0);
        if (!shard_allgather_f64(s->comm, s->total_prob, totals.data(), err)) return set_error(s, QSV_ERR_NCCL, "%s", err.c_str());
        for (int r = 0; r < s->rank; ++r) offset += totals[r];
    }
    const size_t u_bytes = (sizeof(double) * shots + 255) & ~size_t(255);
    rc = ensure_scratch(s, u_bytes + sizeof(uint64_t) * shots);
    if (rc != QSV_OK) return rc;
    double* d_u = static_cast<double*>(s->d_scratch);
    uint64_t* d_out = reinterpret_cast<uint64_t*>(static_cast<uint8_t*>(s->d_scratch) + u_bytes);
    std::vector<double> shifted;
    const double* src = uniforms;
    if (s->comm) {  // shots below this rank's interval get a negative uniform (-> not mine), shots above fall off the end
        shifted.resize(shots);
        for (uint64_t i = 0; i < shots; ++i) shifted[i] = uniforms[i] - offset;
        src = shifted.data();
    }
    std::string err;
    cudaError_t e;
    if ((e = cudaMemcpyAsync(d_u, src, sizeof(double) * shots, cudaMemcpyHostToDevice, s->stream)) != cudaSuccess ||
        (e = launch_sample_shots(s->d_state, s->d_prefix, s->n_blocks, s->block_bits, d_u, shots, rank_base(s), d_out, s->sm_count, s->stream)) != cudaSuccess ||
        (s->comm && !shard_allreduce_min_u64(s->comm, d_out, shots, err)) ||
        (e = cudaMemcpyAsync(out_indices, d_out, sizeof(uint64_t) * shots, cudaMemcpyDeviceToHost, s->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(s->stream)) != cudaSuccess)
        rc = set_error(s, err.empty() ? QSV_ERR_CUDA : QSV_ERR_NCCL, "sample: %s", err.empty() ? cudaGetErrorString(e) : err.c_str());
    if (rc == QSV_OK && !s->layout_identity)
        for (uint64_t i = 0; i < shots; ++i)
            if (out_indices[i] != UINT64_MAX) out_indices[i] = to_logical(s, out_indices[i]);
    return rc;
}

int qsv_norm_sqr(qsv_state* s, double* out) {
    if (is_multi(s)) {
        std::vector<double> v(s->children.size(), 0.0);
        const int rc = multi_run(s, [&](int r) { return qsv_norm_sqr(s->children[(size_t)r], &v[(size_t)r]); });
        if (rc == QSV_OK && out) *out = v[0];
        return rc;
    }
    QSV_ENTER(s);
    if (!out) return set_error(s, QSV_ERR_INVALID_ARG, "out is NULL");
    { int rc0 = materialize(s); if (rc0 != QSV_OK) return rc0; }
    int rc = ensure_total(s);
    if (rc != QSV_OK) return rc;
    double total = s->total_prob;
    if (s->comm) {
        std::string err;
        if (!shard_allreduce_sum(s->comm, &total, err)) return set_error(s, QSV_ERR_NCCL, "%s", err.c_str());
    }
    *out = total;
    return QSV_OK;
}

int qsv_get_layout(const qsv_state* s, uint8_t* out_layout, size_t cap) {
    if (!s || !out_layout || cap < s->n_qubits) return set_error(nullptr, QSV_ERR_INVALID_ARG, "bad argument");
    if (is_multi(s)) return qsv_get_layout(s->children[0], out_layout, cap);
    memcpy(out_layout, s->layout, s->n_qubits);
    return QSV_OK;
}

int qsv_last_step_ms(const qsv_state* s, double* out_ms, size_t cap, size_t* n_steps) {
    if (!s || !n_steps) return set_error(nullptr, QSV_ERR_INVALID_ARG, "NULL argument");
    if (is_multi(s)) return qsv_last_step_ms(s->children[0], out_ms, cap, n_steps);
    *n_steps = s->last_step_ms.size();
    if (out_ms)
        for (size_t i = 0; i < s->last_step_ms.size() && i < cap; ++i) out_ms[i] = s->last_step_ms[i];
    return QSV_OK;
}

namespace {
struct CkptHeader {
    char magic[8];
    uint32_t n_qubits, n_local, rank, world;
    uint8_t layout[64];
    uint8_t pad[256 - 8 - 16 - 64];
};
static_assert(sizeof(CkptHeader) == 256, "checkpoint header");
constexpr size_t kCkptChunk = (size_t)64 << 20;
}  // namespace

int qsv_save(qsv_state* s, const char* path) {
    if (is_multi(s)) {  // one file per shard: <path>.r<rank>
        if (!path) return set_error(s, QSV_ERR_INVALID_ARG, "path is NULL");
        return multi_run(s, [&](int r) { return qsv_save(s->children[(size_t)r], (std::string(path) + ".r" + std::to_string(r)).c_str()); });
    }
    QSV_ENTER(s);
    if (!path) return set_error(s, QSV_ERR_INVALID_ARG, "path is NULL");
    { int rc0 = materialize(s); if (rc0 != QSV_OK) return rc0; }
    FILE* f = fopen(path, "wb");
    if (!f) return set_error(s, QSV_ERR_INVALID_ARG, "cannot open %s for writing", path);
    CkptHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, "QSVCKPT1", 8);
    h.n_qubits = s->n_qubits; h.n_local = s->n_local; h.rank = (uint32_t)s->rank; h.world = (uint32_t)s->world;
    memcpy(h.layout, s->layout, sizeof(h.layout));
    int rc = fwrite(&h, sizeof(h), 1, f) == 1 ? QSV_OK : set_error(s, QSV_ERR_INTERNAL, "write to %s failed", path);
    void* host = nullptr;
    if (rc == QSV_OK && cudaMallocHost(&host, kCkptChunk) != cudaSuccess) { cudaGetLastError(); rc = set_error(s, QSV_ERR_OUT_OF_MEMORY, "pinned staging buffer"); }
    const size_t total = sizeof(cplx) << s->n_local;
    for (size_t off = 0; rc == QSV_OK && off < total; off += kCkptChunk) {
        const size_t m = total - off < kCkptChunk ? total - off : kCkptChunk;
        cudaError_t e = cudaMemcpyAsync(host, reinterpret_cast<const char*>(s->d_state) + off, m, cudaMemcpyDeviceToHost, s->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) rc = set_error(s, QSV_ERR_CUDA, "checkpoint read-back: %s", cudaGetErrorString(e));
        else if (fwrite(host, 1, m, f) != m) rc = set_error(s, QSV_ERR_INTERNAL, "write to %s failed", path);
    }
    if (host) cudaFreeHost(host);
    if (fclose(f) != 0 && rc == QSV_OK) rc = set_error(s, QSV_ERR_INTERNAL, "closing %s failed", path);
    return rc;
}

int qsv_load(qsv_state* s, const char* path) {
    if (is_multi(s)) {
        if (!path) return set_error(s, QSV_ERR_INVALID_ARG, "path is NULL");
        return multi_run(s, [&](int r) { return qsv_load(s->children[(size_t)r], (std::string(path) + ".r" + std::to_string(r)).c_str()); });
    }
    QSV_ENTER(s);
    if (!path) return set_error(s, QSV_ERR_INVALID_ARG, "path is NULL");
    FILE* f = fopen(path, "rb");
    if (!f) return set_error(s, QSV_ERR_INVALID_ARG, "cannot open %s", path);
    CkptHeader h;
    int rc = QSV_OK;
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "QSVCKPT1", 8) != 0) rc = set_error(s, QSV_ERR_INVALID_ARG, "%s is not a checkpoint", path);
    else if (h.n_qubits != s->n_qubits || h.n_local != s->n_local || h.rank != (uint32_t)s->rank || h.world != (uint32_t)s->world)
        rc = set_error(s, QSV_ERR_INVALID_ARG, "%s holds rank %u/%u of a %u-qubit register; the handle is rank %d/%d of %u qubits", path, h.rank, h.world, h.n_qubits, s->rank,
                       s->world, s->n_qubits);
    void* host = nullptr;
    if (rc == QSV_OK && cudaMallocHost(&host, kCkptChunk) != cudaSuccess) { cudaGetLastError(); rc = set_error(s, QSV_ERR_OUT_OF_MEMORY, "pinned staging buffer"); }
    const size_t total = sizeof(cplx) << s->n_local;
    for (size_t off = 0; rc == QSV_OK && off < total; off += kCkptChunk) {
        const size_t m = total - off < kCkptChunk ? total - off : kCkptChunk;
        if (fread(host, 1, m, f) != m) { rc = set_error(s, QSV_ERR_INVALID_ARG, "%s is truncated", path); break; }
        cudaError_t e = cudaMemcpyAsync(reinterpret_cast<char*>(s->d_state) + off, host, m, cudaMemcpyHostToDevice, s->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) rc = set_error(s, QSV_ERR_CUDA, "checkpoint upload: %s", cudaGetErrorString(e));
    }
    if (host) cudaFreeHost(host);
    fclose(f);
    if (rc == QSV_OK) {
        s->lazy_basis = false;
        set_layout(s, h.layout);
        s->prefix_valid = false;
    }
    return rc;
}

int qsv_synchronize(qsv_state* s) {
    if (is_multi(s)) {
        for (qsv_state* c : s->children) {
            const int rc = qsv_synchronize(c);
            if (rc != QSV_OK) return set_error(s, rc, "%s", c->error.c_str());
        }
        return QSV_OK;
    }
    QSV_ENTER(s);
    QSV_CUDA(s, cudaStreamSynchronize(s->stream));
    return QSV_OK;
}

int qsv_device_pointer(qsv_state* s, void** dev_ptr, void** cuda_stream) {
    if (!s) return set_error(nullptr, QSV_ERR_INVALID_ARG, "handle is NULL");
    if (is_multi(s)) return set_error(s, QSV_ERR_UNSUPPORTED, "a multi-device handle has one shard per device, not one device pointer");
    { cudaSetDevice(s->device); int rc0 = materialize(s); if (rc0 != QSV_OK) return rc0; }
    if (dev_ptr) *dev_ptr = s->d_state;
    if (cuda_stream) *cuda_stream = s->stream;
    return QSV_OK;
}

}  // extern "C"
