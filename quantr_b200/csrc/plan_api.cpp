// plan_api.cpp — C ABI for the host-only part of the library (lowering + scheduling).
// Needs no GPU; see include/qsv.h.
#include <stdlib.h>
#include <string.h>

#include <stdexcept>
#include <string>

#include "plan.h"
#include "plan_handle.h"

namespace {
thread_local std::string g_plan_error;
}

extern "C" {

const char* qsv_plan_last_error(void) { return g_plan_error.c_str(); }

int qsv_plan_create_ex(qsv_plan** out, uint32_t n_qubits, uint32_t n_local_qubits, const qsv_op* ops, size_t n_ops,
                       uint32_t tile_bits, uint32_t low_bits, int fuse, const uint8_t* layout, int free_layout) {
    if (!out) { g_plan_error = "out is NULL"; return QSV_ERR_INVALID_ARG; }
    *out = nullptr;
    try {
        qsv_plan* p = new qsv_plan();
        qsv::PlanOptions opt;
        if (tile_bits) opt.tile_bits = (int)tile_bits;
        if (low_bits) opt.low_bits = (int)low_bits;
        opt.fuse = fuse;
        if (const char* env = getenv("QSV_DIRECT_STORE")) opt.direct_store = atoi(env) != 0;
        if (const char* env = getenv("QSV_QFT4")) opt.qft4 = atoi(env) != 0;  // developer A/B switch
        if (const char* env = getenv("QSV_BIG_LOW_PASS")) opt.big_low_pass = atoi(env) != 0;  // developer A/B switch
        if (const char* env = getenv("QSV_FOLD_PREFIX")) opt.fold_prefix = atoi(env) != 0;    // developer A/B switch
        if (const char* env = getenv("QSV_PREFIX_SUBREG")) opt.prefix_subregister = atoi(env) != 0;  // developer A/B switch
        if (const char* env = getenv("QSV_PREFIX_KEEP_BITS")) opt.prefix_keep_bits = atoi(env);      // tests
        if (const char* env = getenv("QSV_PREFIX_MIN_LOCAL")) opt.prefix_min_local = atoi(env);  // tests: fold local qubits on small registers too
        if (const char* env = getenv("QSV_REORDER")) opt.reorder = atoi(env) != 0;            // developer A/B switch
        if (const char* env = getenv("QSV_MERGE_1Q")) opt.merge_1q = atoi(env) != 0;          // developer A/B switch
        if (const char* env = getenv("QSV_PERM_ROUNDS")) opt.perm_rounds = atoi(env) != 0;    // developer A/B switch
        if (const char* env = getenv("QSV_MERGE_CTRL")) opt.merge_ctrl = atoi(env) != 0;      // developer A/B switch
        try {
            qsv::build_plan(p->plan, n_qubits, n_local_qubits, ops, n_ops, opt, layout, free_layout != 0);
        } catch (...) {
            delete p;
            throw;
        }
        *out = p;
        return QSV_OK;
    } catch (const std::bad_alloc&) {
        g_plan_error = "out of host memory while building the plan";
        return QSV_ERR_OUT_OF_MEMORY;
    } catch (const std::exception& e) {
        g_plan_error = e.what();
        const bool unsupported = g_plan_error.find("not supported") != std::string::npos ||
                                 g_plan_error.find("cannot bring") != std::string::npos ||
                                 g_plan_error.find("needs more qubits") != std::string::npos ||
                                 g_plan_error.find("wider than the tile") != std::string::npos ||
                                 g_plan_error.find("qubits per rank") != std::string::npos;
        return unsupported ? QSV_ERR_UNSUPPORTED : QSV_ERR_INVALID_ARG;
    } catch (...) {
        g_plan_error = "unknown error";
        return QSV_ERR_INTERNAL;
    }
}

int qsv_plan_create(qsv_plan** out, uint32_t n_qubits, uint32_t n_local_qubits, const qsv_op* ops, size_t n_ops,
                    uint32_t tile_bits, uint32_t low_bits, int fuse) {
    return qsv_plan_create_ex(out, n_qubits, n_local_qubits, ops, n_ops, tile_bits, low_bits, fuse, nullptr, 0);
}

int qsv_plan_num_steps(const qsv_plan* p, size_t* n_steps) {
    if (!p || !n_steps) { g_plan_error = "NULL argument"; return QSV_ERR_INVALID_ARG; }
    *n_steps = p->plan.steps.size();
    return QSV_OK;
}

int qsv_plan_get_step(const qsv_plan* p, size_t i, int* kind, uint32_t* pass_index, uint8_t* partner_bits, size_t cap) {
    if (!p || !kind || i >= p->plan.steps.size()) { g_plan_error = "bad step index"; return QSV_ERR_INVALID_ARG; }
    const qsv::PlanStep& st = p->plan.steps[i];
    *kind = st.kind == qsv::PlanStep::PASS ? QSV_STEP_PASS : QSV_STEP_EXCHANGE;
    if (pass_index) *pass_index = st.pass_index;
    if (partner_bits)
        for (size_t j = 0; j < st.partner_bits.size() && j < cap; ++j) partner_bits[j] = st.partner_bits[j];
    return QSV_OK;
}

int qsv_plan_get_layout(const qsv_plan* p, int which, uint8_t* out_layout, size_t cap) {
    if (!p || !out_layout) { g_plan_error = "NULL argument"; return QSV_ERR_INVALID_ARG; }
    const std::vector<uint8_t>& l = which ? p->plan.final_layout : p->plan.initial_layout;
    if (cap < l.size()) { g_plan_error = "layout buffer too small"; return QSV_ERR_INVALID_ARG; }
    memcpy(out_layout, l.data(), l.size());
    return QSV_OK;
}

int qsv_plan_initial_amplitudes(const qsv_plan* p, uint64_t basis_index, double* out, size_t cap) {
    if (!p || !out) { g_plan_error = "NULL argument"; return QSV_ERR_INVALID_ARG; }
    if (basis_index >> p->plan.n_qubits) { g_plan_error = "basis index out of range"; return QSV_ERR_INVALID_ARG; }
    try {
        std::vector<qsv::cplx> amps;
        qsv::prefix_amplitudes(p->plan, basis_index, amps);
        if (cap < amps.size()) { g_plan_error = "amplitude buffer too small"; return QSV_ERR_INVALID_ARG; }
        for (size_t r = 0; r < amps.size(); ++r) { out[2 * r] = amps[r].x; out[2 * r + 1] = amps[r].y; }
        return QSV_OK;
    } catch (const std::exception& e) {
        g_plan_error = e.what();
        return QSV_ERR_INTERNAL;
    }
}

int qsv_plan_destroy(qsv_plan* p) {
    if (!p) return QSV_OK;
    if (p->release_device) p->release_device(p);
    delete p;
    return QSV_OK;
}

int qsv_plan_stats(const qsv_plan* p, qsv_stats* stats) {
    if (!p || !stats) { g_plan_error = "NULL argument"; return QSV_ERR_INVALID_ARG; }
    memset(stats, 0, sizeof(*stats));
    stats->n_gates = p->plan.n_gates;
    stats->n_passes = p->plan.passes.size();
    stats->n_rounds = p->plan.n_rounds;
    stats->n_kernel_launches = p->plan.passes.size();
    stats->bytes_per_pass = 32ull << p->plan.n_alloc;
    const uint32_t g = p->plan.n_qubits - p->plan.n_local;
    for (const auto& st : p->plan.steps)
        if (st.kind == qsv::PlanStep::EXCHANGE) {
            stats->n_exchanges++;
            stats->exchange_bytes += ((16ull << p->plan.n_local) >> g) * ((1ull << g) - 1);  // (P-1)/P of the shard
        }
    return QSV_OK;
}

int qsv_plan_serialize(const qsv_plan* p, void* out, size_t cap, size_t* size) {
    if (!p || !size) { g_plan_error = "NULL argument"; return QSV_ERR_INVALID_ARG; }
    try {
        const std::string s = qsv::describe_plan(p->plan);
        *size = s.size();
        if (out && cap) memcpy(out, s.data(), s.size() < cap ? s.size() : cap);
        return QSV_OK;
    } catch (const std::exception& e) {
        g_plan_error = e.what();
        return QSV_ERR_INTERNAL;
    }
}

}  // extern "C"
