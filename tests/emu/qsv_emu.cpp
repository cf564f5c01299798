// qsv_emu.cpp — TEST INFRASTRUCTURE ONLY.
//
// Walks a plan's pass blobs on the CPU by calling the very same __host__ __device__
// per-thread functions (quantr_b200/csrc/pass_core.h) the sm_100a kernel runs, one
// "thread" after another with the barriers where the kernel has them.  This lets the
// `-m "not gpu"` suite check the host scheduler, the blob encoding and the per-thread
// arithmetic against the oracle without a GPU.  It is compiled only into
// tests/emu/libqsv_emu.so and never linked or loaded by the product library.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../quantr_b200/csrc/pass_core.h"
#include "../../quantr_b200/csrc/peer_swap.h"
#include "../../quantr_b200/csrc/plan_handle.h"
#include "../../quantr_b200/csrc/tma_tile.h"

using namespace qsv;

// 1: passes whose tile is a tensor-map box are walked the way pass_kernel_tma.cu moves them (box copies through the
// software TMA model of tma_tile.h, tile phases from the two half-index tables, scaled last round + box store);
// 0: the way the synchronous kernel does (per-thread loads, term-loop phases).
static int g_tma_mode = 0;
extern "C" void qsv_emu_set_tma_mode(int on) { g_tma_mode = on; }
// number of passes the last qsv_emu_run_* call walked in TMA mode, and the largest number of boxes per tile among them
static uint32_t g_tma_passes = 0, g_tma_max_boxes = 0;
extern "C" uint32_t qsv_emu_tma_passes(void) { return g_tma_passes; }
extern "C" uint32_t qsv_emu_tma_max_boxes(void) { return g_tma_max_boxes; }

static void run_pass(const uint8_t* blob, cplx* state, uint64_t rank_hi, const PassInit* init = nullptr, uint32_t n_alloc = 0, const PassSlice* slice = nullptr) {
    static PassParams<kMaxRounds, kMaxOps> P;  // what the kernel receives by value
    if (!fill_params(blob, P)) return;
    constexpr int W = kMaxOps / 32;
    const uint32_t T = P.hdr.tile_bits;
    const uint32_t tile_len = 1u << T, groups = 1u << (T - kRegBits), threads = P.hdr.threads;
    const uint32_t n_loads = (tile_len + threads - 1) / threads;
    std::vector<cplx> tile(tile_len);
    std::vector<cplx> ext_phase(kMaxOps);
    std::vector<cplx> dense_out((size_t)groups * kSlots);
    const bool fast = pass_is_fast(P.hdr);
    std::vector<uint32_t> thr_act((size_t)groups * W);
    for (uint32_t e = 0; e < groups; ++e) {  // the kernel does this once per launch
        uint32_t act[W];
        if (fast) for (int w = 0; w < W; ++w) act[w] = 0xffffffffu;
        else thread_active_mask<W>(P.hdr, P.rounds, P.ops, e, act);
        memcpy(&thr_act[(size_t)e * W], act, sizeof(act));
    }
    // DIAG thread phases: same placement decision as the launcher
    const int mode = fast ? 2 : choose_diag_mode(T, P.hdr.n_diag);
    std::vector<cplx> thr_tbl((size_t)kMaxOps * kDiagTblLen), thr_phase((size_t)kMaxOps * threads);
    for (uint32_t o = 0; o < P.hdr.n_ops; ++o)
        if (P.ops[o].type == OP_DIAG) {
            const cplx* src = reinterpret_cast<const cplx*>(blob + P.ops[o].tbl_off);
            memcpy(&thr_tbl[(size_t)P.ops[o].diag_index * kDiagTblLen], src, sizeof(cplx) * kDiagTblLen);
            for (uint32_t e = 0; e < groups; ++e) thr_phase[(size_t)P.ops[o].diag_index * threads + e] = diag_thread_phase(P.ops[o], src, e);
        }
    DiagCtx ctx{blob, ext_phase.data(), mode == 1 ? thr_tbl.data() : nullptr, mode == 2 ? thr_phase.data() : nullptr, threads};
    const bool direct = (P.hdr.flags & PASS_DIRECT_STORE) != 0;
    const double sc = P.hdr.final_scale;
    char* tb = reinterpret_cast<char*>(tile.data());
    // TMA mode: tensor-map description of the tile + the external-phase tables exactly as the device builds them
    TmaTileDesc desc;
    const bool tma = g_tma_mode && n_alloc && T >= 6 && make_tma_tile(P.hdr, n_alloc, desc);
    const uint32_t a_bits = ext_table_low_bits(P.hdr.n_tiles);
    const uint64_t tbl_len = ext_table_len(P.hdr.n_tiles);
    std::vector<cplx> ext_tbl;
    if (tma) {
        ++g_tma_passes;
        if (desc.tile.n_boxes > g_tma_max_boxes) g_tma_max_boxes = desc.tile.n_boxes;
        ext_tbl.resize((size_t)P.hdr.n_ext_ops * tbl_len);
        for (uint32_t o = 0; o < P.hdr.n_ops; ++o)
            if (P.ops[o].type == OP_DIAG && P.ops[o].ext_slot != kNoExtSlot) {
                const DiagExtTerm* terms = reinterpret_cast<const DiagExtTerm*>(blob + P.ops[o].ext_off);
                for (uint64_t i = 0; i < tbl_len; ++i) {
                    const bool low = i < (1ull << a_bits);
                    ext_tbl[(size_t)P.ops[o].ext_slot * tbl_len + i] =
                        ext_table_entry(P.hdr, P.ops[o].theta0, terms, P.ops[o].n_ext, low ? i : (i - (1ull << a_bits)) << a_bits, rank_hi, low);
                }
            }
    }
    auto tma_move = [&](uint64_t tile_id, bool store) {  // every box of tile `tile_id`, coordinates as the kernel computes them
        double* gd = reinterpret_cast<double*>(state);
        for (uint32_t j = 0; j < desc.tile.n_boxes; ++j) {
            int32_t c[kTmaRank];
            tma_tile_coords(desc.tile, (uint32_t)tile_id, j, c);
            char* sb = tb + (size_t)j * desc.tile.box_bytes;
            tma_box_walk(desc, c, [&](uint64_t goff, uint64_t soff) {
                if (store) gd[goff] = *reinterpret_cast<double*>(sb + soff);
                else *reinterpret_cast<double*>(sb + soff) = gd[goff];
            });
        }
    };
    // a launch over a slice of the register enumerates the slice's tiles exactly as the kernel does (tma_tile.h slice_tile_id)
    TmaTile slice_tile{};
    if (!set_tma_slice(P.hdr, slice, slice_tile)) __builtin_trap();
    for (uint64_t k = 0; k < (P.hdr.n_tiles >> slice_tile.slice_n); ++k) {
        const uint64_t t = slice_tile_id(slice_tile, (uint32_t)k);
        const uint64_t base = deposit(t, P.hdr.ext_segs, P.hdr.n_ext_segs);
        if (tma && tma_tile_base(desc.tile, (uint32_t)t) != base) __builtin_trap();  // the kernel derives the tile base from the tile id fields
        const uint64_t base_full = base | rank_hi;
        const bool holds = init && init_tile_holds(*init, base_full);
        if (init && init->mode == 2 && !holds && tma) {  // bulk store from the zeroed buffer
            memset(tile.data(), 0, sizeof(cplx) * tile_len);
            tma_move(t, true);
            continue;
        }
        if (init && init->mode == 2 && !holds) {  // fused initialisation, zero tile in -> zero tile out
            for (uint32_t tid = 0; tid < threads; ++tid) {
                const uint64_t goff_t = deposit(tid, P.hdr.tile_segs, P.hdr.n_tile_segs);
                for (uint32_t i = 0; i < n_loads; ++i)
                    if (i * threads + tid < tile_len) state[base + goff_t + P.loads.goff[i]] = cplx{0.0, 0.0};
            }
            continue;
        }
        if (tma && !init) tma_move(t, false);
        else
        for (uint32_t tid = 0; tid < threads; ++tid) {  // load phase exactly as the kernel addresses it
            const uint64_t goff_t = deposit(tid, P.hdr.tile_segs, P.hdr.n_tile_segs);
            const uint32_t soff_t = swz(tid) << 4;
            for (uint32_t i = 0; i < n_loads; ++i)
                if (i * threads + tid < tile_len) {
                    cplx v = init ? (holds ? init_tile_element(*init, P.hdr, base, i * threads + tid) : cplx{0.0, 0.0})  // synthesised, the register is not read
                                  : state[base + goff_t + P.loads.goff[i]];
                    *reinterpret_cast<cplx*>(tb + (soff_t ^ P.loads.soff[i])) = v;
                }
        }
        for (uint32_t o = 0; o < P.hdr.n_ops; ++o)
            if (P.ops[o].type == OP_DIAG) {
                if (tma && !init && P.ops[o].ext_slot != kNoExtSlot) {
                    const cplx* tbl = &ext_tbl[(size_t)P.ops[o].ext_slot * tbl_len];
                    ext_phase[P.ops[o].diag_index] = cmul(tbl[t & ((1ull << a_bits) - 1)], tbl[(1ull << a_bits) + (t >> a_bits)]);
                } else if (tma && !init) {
                    ext_phase[P.ops[o].diag_index] = diag_ext_phase_terms(P.ops[o].theta0, nullptr, 0, 0);
                } else {
                    ext_phase[P.ops[o].diag_index] = diag_ext_phase(P.ops[o], blob, base_full);
                }
            }
        for (uint32_t r = 0; r < P.hdr.n_rounds; ++r) {
            const DevRound& R = P.rounds[r];
            if (R.type == ROUND_PERM) {  // every thread gathers from the tile as it was before the round (the kernel has a barrier there)
                std::vector<cplx> before(tile);
                for (uint32_t e = 0; e < groups; ++e) {
                    uint32_t act[W];
                    memcpy(act, &thr_act[(size_t)e * W], sizeof(act));
                    tile_active_mask<W>(P.hdr, P.ops, base_full, act);
                    const uint32_t lb = round_thread_base(R, e);
                    cplx a[kSlots];
                    if (fast) round_perm_load<W, true>(R, P.ops, act, lb, before.data(), a);
                    else round_perm_load<W, false>(R, P.ops, act, lb, before.data(), a);
                    if (fast) round_ops<W, true>(R, P.ops, ctx, act, e, a);  // ordinary ops behind the gather, if any
                    else round_ops<W, false>(R, P.ops, ctx, act, e, a);
                    if (direct && r + 1 == P.hdr.n_rounds) {
                        const uint64_t g = base + deposit(lb, P.hdr.tile_segs, P.hdr.n_tile_segs);
                        for (int s = 0; s < kSlots; ++s) state[g + P.loads.store_goff[s]] = cplx{a[s].x * sc, a[s].y * sc};
                    } else if (tma && r + 1 == P.hdr.n_rounds) {
                        round_store_tile_scaled(R, lb, tile.data(), a, sc);
                    } else {
                        round_store_tile(R, lb, tile.data(), a);
                    }
                }
            } else if (R.type == ROUND_REG) {
                for (uint32_t e = 0; e < groups; ++e) {
                    uint32_t act[W];
                    memcpy(act, &thr_act[(size_t)e * W], sizeof(act));
                    tile_active_mask<W>(P.hdr, P.ops, base_full, act);
                    const uint32_t lb = round_thread_base(R, e);
                    cplx a[kSlots];
                    round_load(R, lb, tile.data(), a);
                    if (fast) round_ops<W, true>(R, P.ops, ctx, act, e, a);
                    else round_ops<W, false>(R, P.ops, ctx, act, e, a);
                    if (direct && r + 1 == P.hdr.n_rounds) {
                        const uint64_t g = base + deposit(lb, P.hdr.tile_segs, P.hdr.n_tile_segs);
                        for (int s = 0; s < kSlots; ++s) {
                            cplx v = a[s];
                            if (sc != 1.0) { v.x *= sc; v.y *= sc; }
                            state[g + P.loads.store_goff[s]] = v;
                        }
                    } else if (tma && r + 1 == P.hdr.n_rounds) {
                        round_store_tile_scaled(R, lb, tile.data(), a, sc);
                    } else {
                        round_store_tile(R, lb, tile.data(), a);
                    }
                }
            } else {
                const DevDense* D = reinterpret_cast<const DevDense*>(blob + P.ops[R.first_op].dense_off);
                for (uint32_t e = 0; e < groups; ++e) {
                    cplx out[kSlots];
                    dense_compute(*D, blob, e, tile.data(), out);
                    memcpy(&dense_out[(size_t)e * kSlots], out, sizeof(out));
                }
                for (uint32_t e = 0; e < groups; ++e) {
                    cplx out[kSlots];
                    memcpy(out, &dense_out[(size_t)e * kSlots], sizeof(out));
                    if (tma && r + 1 == P.hdr.n_rounds)
                        for (int sl = 0; sl < kSlots; ++sl) { out[sl].x *= sc; out[sl].y *= sc; }
                    dense_store(e, tile.data(), out);
                }
            }
        }
        if (direct) continue;
        if (tma) { tma_move(t, true); continue; }
        for (uint32_t tid = 0; tid < threads; ++tid) {
            const uint64_t goff_t = deposit(tid, P.hdr.tile_segs, P.hdr.n_tile_segs);
            const uint32_t soff_t = swz(tid) << 4;
            for (uint32_t i = 0; i < n_loads; ++i)
                if (i * threads + tid < tile_len) {
                    cplx v = *reinterpret_cast<const cplx*>(tb + (soff_t ^ P.loads.soff[i]));
                    if (sc != 1.0) { v.x *= sc; v.y *= sc; }
                    state[base + goff_t + P.loads.goff[i]] = v;
                }
        }
    }
}

// Runs every PASS step in order on one rank's shard (plans without EXCHANGE steps).
extern "C" int qsv_emu_run_plan(const qsv_plan* p, double* amps, uint64_t rank) {
    if (!p || !amps) return 1;
    g_tma_passes = g_tma_max_boxes = 0;
    const uint64_t rank_hi = rank << p->plan.n_local;
    for (const auto& st : p->plan.steps) {
        if (st.kind != PlanStep::PASS) return 2;  // the caller must drive exchanges (qsv_emu_run_pass per step)
        run_pass(p->plan.passes[st.pass_index].data(), reinterpret_cast<cplx*>(amps), rank_hi, nullptr, p->plan.n_alloc);
    }
    return 0;
}

extern "C" int qsv_emu_run_pass(const qsv_plan* p, uint32_t pass_index, double* amps, uint64_t rank) {
    if (!p || !amps || pass_index >= p->plan.passes.size()) return 1;
    run_pass(p->plan.passes[pass_index].data(), reinterpret_cast<cplx*>(amps), rank << p->plan.n_local, nullptr, p->plan.n_alloc);
    return 0;
}

extern "C" uint32_t qsv_emu_alloc_qubits(const qsv_plan* p) { return p ? p->plan.n_alloc : 0; }

// The pipelined exchange (state_api.cu run_overlapped), piece by piece: one pass over one slice of the shard ...
extern "C" int qsv_emu_run_pass_slice(const qsv_plan* p, uint32_t pass_index, double* amps, uint64_t rank, uint32_t n_slice, const uint8_t* slice_bits, uint32_t value) {
    if (!p || !amps || pass_index >= p->plan.passes.size() || n_slice > 3) return 1;
    PassSlice sl{};
    sl.n = n_slice;
    for (uint32_t i = 0; i < n_slice; ++i) sl.bit[i] = slice_bits[i];
    sl.value = value;
    TmaTile probe{};
    if (!set_tma_slice(*reinterpret_cast<const DevPass*>(p->plan.passes[pass_index].data()), &sl, probe)) return 2;
    run_pass(p->plan.passes[pass_index].data(), reinterpret_cast<cplx*>(amps), rank << p->plan.n_local, nullptr, p->plan.n_alloc, &sl);
    return 0;
}
// ... and what plan_overlap_group decides for an EXCHANGE step: out = {slice_prev, slice_next, n_bits, bits[0..2]}; returns 1 if it overlaps
extern "C" int qsv_emu_overlap_group(const qsv_plan* p, uint32_t step, const uint8_t* sliceable, uint32_t log2_slices, uint32_t* out) {
    if (!p || !sliceable || !out) return -1;
    std::vector<char> sl(p->plan.steps.size());
    for (size_t i = 0; i < sl.size(); ++i) sl[i] = (char)sliceable[i];
    OverlapGroup g;
    const bool ok = plan_overlap_group(p->plan, step, sl, log2_slices, g);
    out[0] = g.slice_prev; out[1] = g.slice_next; out[2] = g.n_bits; out[3] = g.bits[0]; out[4] = g.bits[1]; out[5] = g.bits[2];
    return ok ? 1 : 0;
}

// One EXCHANGE step the way the peer-memory path runs it (state_api.cu run_exchange + peer_swap_kernel): every rank,
// for every step of the round-robin pairing, swaps its half of the pair's blocks in place.  shards[r] = rank r's amplitudes.
static int peer_exchange(double** shards, uint32_t n_local, const uint8_t* partner, uint32_t g, uint32_t n_slice, const uint8_t* slice_bits, uint32_t value);
extern "C" int qsv_emu_peer_exchange(double** shards, uint32_t n_local, const uint8_t* partner, uint32_t g) {
    return peer_exchange(shards, n_local, partner, g, 0, nullptr, 0);
}
// one slice of it (the amplitudes whose index bits slice_bits[] spell value)
extern "C" int qsv_emu_peer_exchange_slice(double** shards, uint32_t n_local, const uint8_t* partner, uint32_t g, uint32_t n_slice, const uint8_t* slice_bits, uint32_t value) {
    return peer_exchange(shards, n_local, partner, g, n_slice, slice_bits, value);
}
static int peer_exchange(double** shards, uint32_t n_local, const uint8_t* partner, uint32_t g, uint32_t n_slice, const uint8_t* slice_bits, uint32_t value) {
    const int world = 1 << g;
    for (int step = 1; step < world; ++step)
        for (int rank = 0; rank < world; ++rank) {
            const int peer = rank ^ step;
            const SwapArgs a = make_swap_args(n_local, partner, g, rank, peer, n_slice, slice_bits, value);
            cplx* local = reinterpret_cast<cplx*>(shards[rank]);
            cplx* remote = reinterpret_cast<cplx*>(shards[peer]);
            for (uint64_t j = 0; j < a.count; ++j) {
                const uint64_t idx = insert_zero_bits(a.first + j, a);
                const cplx mine = local[idx | a.local_spell], theirs = remote[idx | a.remote_spell];
                local[idx | a.local_spell] = theirs;
                remote[idx | a.remote_spell] = mine;
            }
        }
    return 0;
}

// The folded prefix run as a plan of its own on the support qubits (plan.cpp build_prefix_subplan; on the device this is how
// prefixes too wide for the host table are applied): out = the sub-register's final state, 2^(g + prefix_local_bits) amplitudes.
extern "C" int qsv_emu_prefix_subplan_table(const qsv_plan* p, uint64_t basis_index, double* out, size_t cap) {
    if (!p || !out) return 1;
    try {
        Plan sub;
        build_prefix_subplan(p->plan, basis_index, sub);
        const uint32_t nf = p->plan.n_local - p->plan.prefix_local_bits;
        if (cap < ((size_t)1 << sub.n_alloc)) return 2;
        std::vector<cplx> reg((size_t)1 << sub.n_alloc, cplx{0.0, 0.0});
        reg[basis_index >> nf] = cplx{1.0, 0.0};
        for (const PlanStep& st : sub.steps) {
            if (st.kind != PlanStep::PASS) return 3;
            run_pass(sub.passes[st.pass_index].data(), reg.data(), 0, nullptr, sub.n_alloc);
        }
        memcpy(out, reg.data(), sizeof(cplx) * reg.size());
        return 0;
    } catch (...) {
        return 4;
    }
}

// A plan on a basis state with the initialisation fused into its first pass (state_api.cu, QSV_FUSED_INIT): `amps` is
// never read by that pass.  phys_index = the basis index under the plan's initial layout, rank bits included.
extern "C" int qsv_emu_run_plan_fused_init(const qsv_plan* p, double* amps, uint64_t rank, uint64_t phys_index, uint32_t mode) {
    if (!p || !amps || p->plan.steps.empty() || p->plan.steps[0].kind != PlanStep::PASS) return 1;
    for (const PlanStep& st : p->plan.steps)
        if (st.kind != PlanStep::PASS) return 1;
    const uint64_t rank_hi = rank << p->plan.n_local;
    bool first = true;
    for (const PlanStep& st : p->plan.steps) {
        const uint8_t* blob = p->plan.passes[st.pass_index].data();
        if (first) {
            // a plan with a folded prefix starts from the prefix's amplitudes (canonical layout: phys_index is the basis index)
            std::vector<cplx> tbl;
            PassInit pi;
            if (!p->plan.prefix.empty()) {
                prefix_amplitudes(p->plan, phys_index, tbl);
                const uint32_t k = p->plan.prefix_local_bits;
                pi = make_pass_init(*reinterpret_cast<const DevPass*>(blob), (phys_index & ((1ull << p->plan.n_local) - 1ull)) | rank_hi, p->plan.n_local, mode, k,
                                    k ? tbl.data() + ((size_t)rank << k) : nullptr);
                if (k == 0) { pi.amp_re = tbl[(size_t)rank].x; pi.amp_im = tbl[(size_t)rank].y; }
            } else {
                pi = make_pass_init(*reinterpret_cast<const DevPass*>(blob), phys_index, p->plan.n_local, mode);
            }
            run_pass(blob, reinterpret_cast<cplx*>(amps), rank_hi, &pi, p->plan.n_alloc);
            first = false;
        } else {
            run_pass(blob, reinterpret_cast<cplx*>(amps), rank_hi, nullptr, p->plan.n_alloc);
        }
    }
    return 0;
}
