// pass_kernel_tma.cu — the fused pass as a persistent, software-pipelined kernel built on the TMA unit: one CTA per SM,
// G independent compute groups (one thread per 16 amplitudes each) and a ring of tile buffers in shared memory.
//
//   tile movement   A tile is a box of a five-dimensional tensor map over the register (tma_tile.h).  One elected
//                   thread per group issues cp.async.bulk.tensor (SASS: UTMALDG) for the tile a buffer will hold next,
//                   completing on the buffer's mbarrier (expect_tx); passes whose last round cannot store its registers
//                   coalesced leave through the buffer with a bulk tensor store (UTMASTG).  Shared memory is laid out by
//                   the TMA unit itself (CU_TENSOR_MAP_SWIZZLE_128B = qsv_types.h swz()).
//   tile phases     exp(i*pi*(theta0 + bits outside the tile)) of every DIAG op is the product of two table entries
//                   indexed by the halves of the tile id (pass_core.h ext_table_entry); the two 16-byte entries per op
//                   ride along with the tile as cp.async.bulk copies (UBLKCP) on the same mbarrier.
//   initialisation  On a register that is a basis state not yet written to HBM (qsv_init_basis is lazy) the first pass
//                   of a plan does not read it: the one tile holding the amplitude is synthesised in shared memory and
//                   every other tile - all zero in, all zero out, the pass is linear - is covered by one contiguous stream
//                   of zero stores over the shard (init.mode 2), or synthesised and computed like any other (mode 1, tests).
//
// Ring protocol: tile k of the CTA lives in buffer k mod NB and is worked on by group k mod G; the group that has moved
// tile k into registers (or out through a bulk store) refills that buffer with tile k + NB.  A consumer polls the
// buffer's mbarrier by parity only once the load it waits for has been issued (per-buffer issue counter, st.release /
// ld.acquire): a group can run two tiles ahead of the group that refills its next buffer, and a parity wait placed
// before that refill would be satisfied by the phase before last (tests/test_ring_protocol.py).
//
// Compiled per tile size (-DQSV_TILE_BITS=11|12).  Replaces Circuit::apply_gate (src/circuit/simulation.rs:64-135) for a
// fused list of gates and, in init mode, SuperPosition::new_unchecked (super_positions_unchecked.rs:39-46); the
// per-thread arithmetic is the pass_core.h code shared with the synchronous kernel and the host emulation.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "kernels.h"
#include "pass_core.h"
#include "tma_tile.h"

#ifndef QSV_TILE_BITS
#error "compile with -DQSV_TILE_BITS=<11|12>"
#endif

namespace qsv {

template <int T>
struct TmaCfg {
    static constexpr uint32_t kGroupThreads = 1u << (T - kRegBits);
    static constexpr uint32_t kGroups = (T >= 12) ? 2u : 4u;
    static constexpr uint32_t kBuffers = (T >= 12) ? 3u : 6u;
    static constexpr uint32_t kThreads = kGroupThreads * kGroups;
    static constexpr uint32_t kTileBytes = (uint32_t)sizeof(cplx) << T;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Spins on the phase with the given parity.  A wait that outlasts any legitimate tile load (about 2^22 timed-out
// try_waits, seconds) traps, so a protocol error surfaces as a launch failure instead of a hung device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spins = 0;; ++spins) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (spins > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void st_release_shared(uint32_t* p, uint32_t v) {
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// Spins until *p >= want (acquire); same watchdog as mbar_wait.
__device__ __forceinline__ void wait_issued(const uint32_t* p, uint32_t want) {
    const uint32_t addr = smem_u32(p);
    for (uint32_t spins = 0;; ++spins) {
        uint32_t v;
        asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
        if (v >= want) return;
        __nanosleep(64);  // the refill is issued by another group of the CTA: leave it the issue slots
        if (spins > (1u << 24)) __trap();
    }
}
// one box of the tile: global -> shared, completing `box bytes` on the mbarrier
__device__ __forceinline__ void tma_load_box(void* smem_dst, const CUtensorMap* map, uint64_t* bar, const int32_t (&c)[kTmaRank]) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])
                 : "memory");
}
// one box of the tile: shared -> global (bulk async-group of the issuing thread)
__device__ __forceinline__ void tma_store_box(const CUtensorMap* map, const void* smem_src, const int32_t (&c)[kTmaRank]) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem_src)), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// 16-byte bulk copy global -> shared on the same mbarrier (the table entries of a tile's phases)
__device__ __forceinline__ void bulk_copy16(void* smem_dst, const void* gmem_src, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(smem_u32(bar))
                 : "memory");
}
// generic-proxy writes to shared memory become visible to the async proxy (bulk stores read them)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void group_barrier(uint32_t group, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1u), "r"(threads) : "memory");
}
__device__ __forceinline__ void st_stream(cplx* p, cplx v) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y)); }

constexpr uint32_t kLbTabRounds = 8;  // rounds whose per-thread tile-local base is kept in a shared-memory table

// diag_mode bits 0-1: where the DIAG thread phases live (pass_core.h DiagCtx); bit 2: every pass leaves through its tile
// buffer with a bulk tensor store, PASS_DIRECT_STORE is ignored (A/B switch QSV_TMA_STORE).
// KIND: 0 = everything; 1 = FAST (pass_core.h pass_is_fast: no control masks, register rounds only); 2 = LEAN: control masks,
// but register rounds only (no dense or permutation round compiled in).  The kernel sits at the 128-register ceiling and
// every feature in the shared source perturbs its register allocation, hence the leaner builds for the common passes.
template <int T, int NR, int NO, int KIND>
__global__ void __launch_bounds__(TmaCfg<T>::kThreads, 1)
pass_kernel_tma(const __grid_constant__ CUtensorMap tmap, cplx* __restrict__ state, const uint8_t* __restrict__ blob, const cplx* __restrict__ ext_tbl,
                uint64_t rank_hi, int diag_mode, PassInit init, const __grid_constant__ TmaTile tt, const __grid_constant__ PassParams<NR, NO> P) {
    using Cfg = TmaCfg<T>;
    constexpr bool FAST = KIND == 1, LEAN = KIND != 0;
    constexpr uint32_t kGT = Cfg::kGroupThreads, kNB = Cfg::kBuffers, kG = Cfg::kGroups;
    constexpr uint32_t kTileLen = 1u << T;
    constexpr int W = (NO + 31) / 32;
    extern __shared__ uint8_t smem_raw[];
    // the swizzle pattern repeats every 1024 bytes of shared-memory address: tile buffers start on that boundary
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    // layout: [kNB tiles][kNB mbarriers][kNB issue counters][per-round thread bases][kNB x 2 n_ext_ops table entries]
    //         [per group 2 x tile phases][DIAG tables]
    cplx* tiles = reinterpret_cast<cplx*>(smem);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)kNB * Cfg::kTileBytes);
    uint32_t* issued = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(full_bar) + 64);
    uint32_t* lbtab = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(full_bar) + 128);
    cplx* ext_raw = reinterpret_cast<cplx*>(lbtab + kLbTabRounds * kGT);
    const uint32_t n_diag = P.hdr.n_diag, n_ext_ops = tt.n_ext_ops;
    // per group two sets of tile phases (tile parity): a warp that is already on the group's next tile must not
    // overwrite what a slower warp of the group still reads in the last round of the current one
    cplx* ext_all = ext_raw + (size_t)kNB * 2u * n_ext_ops;
    cplx* diag_smem = ext_all + (size_t)kG * 2u * (n_diag + 1);

    const uint32_t tid = threadIdx.x, group = tid / kGT, gtid = tid % kGT, lane = tid & 31u;
    cplx* ext_phase2 = ext_all + (size_t)group * 2u * (n_diag + 1);
    const uint32_t n_ext_segs = P.hdr.n_ext_segs, n_rounds = P.hdr.n_rounds;
    const double final_scale = P.hdr.final_scale;
    const bool last_is_reg = n_rounds && (LEAN || P.rounds[n_rounds - 1].type != ROUND_DENSE);  // register or permutation round
    const bool direct = (P.hdr.flags & PASS_DIRECT_STORE) != 0 && !(diag_mode & 4) && last_is_reg;
    const bool need_base = direct || (P.hdr.ext_ctrl_mask[0] | P.hdr.ext_ctrl_mask[1] | P.hdr.ext_ctrl_mask[2]) != 0 || init.mode != 0;
    // Tiles of this CTA (tile ids and per-CTA counts fit 32 bits): the k-th tile it works on is
    //   t_k = ((blockIdx.x + (k >> ilog) * gridDim.x) << ilog) | (k & (2^ilog - 1)),   k < n_my.
    // With ilog = log2(kG) the kG groups of the CTA work on kG tiles with consecutive ids at the same time: tile ids count
    // the index bits right above the tile's 128-byte rows, so the CTA's loads and stores cover kG adjacent rows (512
    // contiguous bytes) at every row position instead of one - measured, DRAM streams 128-byte runs at ~58 % of the
    // copy peak and 256-byte runs at ~90 % (profiles/r02_stream_probe.txt).
    // A launch over a slice of the register (tt.slice_n fixed tile-id bits) enumerates the slice's tiles and spreads their
    // ids around the fixed bits.
    const uint32_t n_tiles_all = (uint32_t)P.hdr.n_tiles, n_tiles = n_tiles_all >> tt.slice_n;
    const uint32_t ilog = ((uint32_t)diag_mode >> 8) & 7u, imask = (1u << ilog) - 1u;
    const uint32_t n_super = n_tiles >> ilog;
    const uint32_t n_my = (n_super > blockIdx.x ? (n_super - blockIdx.x + gridDim.x - 1) / gridDim.x : 0) << ilog;
    auto tile_of = [&](uint32_t k) { return slice_tile_id(tt, ((blockIdx.x + (k >> ilog) * gridDim.x) << ilog) | (k & imask)); };
    const uint32_t tbl_a_mask = (1u << tt.tbl_a_bits) - 1u;
    const uint32_t tbl_stride = (tbl_a_mask + 1u) + ((n_tiles_all + tbl_a_mask) >> tt.tbl_a_bits);  // entries per op: table A then table B

    // warp 0 of the calling group: tile t_id -> buffer `slot` (its use number `use`), its table entries -> the buffer's
    // slot of ext_raw
    auto issue_load = [&](uint32_t t_id, uint32_t slot, uint32_t use) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&full_bar[slot], Cfg::kTileBytes + 32u * n_ext_ops);
            uint8_t* dst = reinterpret_cast<uint8_t*>(tiles + (size_t)slot * kTileLen);
            for (uint32_t j = 0; j < tt.n_boxes; ++j) {
                int32_t c[kTmaRank];
                tma_tile_coords(tt, t_id, j, c);
                tma_load_box(dst + (size_t)j * tt.box_bytes, &tmap, &full_bar[slot], c);
            }
        }
        for (uint32_t j = lane; j < n_ext_ops; j += 32u) {
            const cplx* tbl = ext_tbl + (size_t)j * tbl_stride;
            cplx* dst = ext_raw + ((size_t)slot * n_ext_ops + j) * 2u;
            bulk_copy16(dst, tbl + (t_id & tbl_a_mask), &full_bar[slot]);
            bulk_copy16(dst + 1, tbl + tbl_a_mask + 1u + (t_id >> tt.tbl_a_bits), &full_bar[slot]);
        }
        __syncwarp();
        // published after the issuing group itself saw the buffer's previous phase complete (it consumed that tile)
        if (lane == 0) st_release_shared(&issued[slot], use + 1u);
    };
    // thread 0 of a group: the finished tile leaves through its buffer
    auto store_tile = [&](const cplx* tile, uint32_t t_id) {
        for (uint32_t j = 0; j < tt.n_boxes; ++j) {
            int32_t c[kTmaRank];
            tma_tile_coords(tt, t_id, j, c);
            tma_store_box(&tmap, reinterpret_cast<const uint8_t*>(tile) + (size_t)j * tt.box_bytes, c);
        }
        bulk_commit();
    };

    if (tid == 0) {
        for (uint32_t b = 0; b < kNB; ++b) {
            mbar_init(&full_bar[b], 1);
            issued[b] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // init mode 2: every tile but one is all zero in, all zero out.  The zeros are written here as one contiguous,
    // coalesced stream over the shard - independent of the pass's tile shape, which would cost short-run write efficiency
    // for nothing - skipping the 128-byte rows of the one tile that holds the amplitude; that tile is then synthesised
    // and computed by the group it falls to in the loop below.
    if (init.mode == 2) {
        const uint64_t n_rows = 1ull << (init.n_alloc - 3);
        // (with a folded prefix on the top local bits several tiles hold amplitudes: their rows differ in the support bits only)
        const uint64_t hold_mask = init.ext_mask & ~init.sup_mask;
        const uint64_t row_mask = hold_mask >> 3, row_hold = (init.base_full & hold_mask & ((1ull << init.n_alloc) - 1ull)) >> 3;
        const bool here = (init.base_full >> init.n_alloc) == (rank_hi >> init.n_alloc);
        const uint64_t rows_per_cta = (n_rows + gridDim.x - 1) / gridDim.x;
        const uint64_t r0 = rows_per_cta * blockIdx.x, r1 = r0 + rows_per_cta < n_rows ? r0 + rows_per_cta : n_rows;
        const uint32_t sub = tid & 7u;  // amplitude within the row
        if (!(here && row_mask == 0))  // (a prefix folded over every tile-id bit: all tiles hold amplitudes, nothing to zero)
            for (uint64_t r = r0 + (tid >> 3); r < r1; r += Cfg::kThreads >> 3)
                if (!(here && (r & row_mask) == row_hold)) st_stream(state + (r << 3) + sub, cplx{0.0, 0.0});
    }
    __syncthreads();
    // prologue: the first kNB tiles, spread over the groups
    if (!init.mode && gtid < 32u)
        for (uint32_t k = group; k < kNB; k += kG)
            if (k < n_my) issue_load(tile_of(k), k, 0u);

    // ---- once per launch: thread-dependent pieces that do not depend on the tile ----------------------
    uint32_t thr_act[W];
    if constexpr (FAST) {
#pragma unroll
        for (int w = 0; w < W; ++w) thr_act[w] = 0xffffffffu;
    } else {
        thread_active_mask<W>(P.hdr, P.rounds, P.ops, gtid, thr_act);
    }
    for (uint32_t o = 0; o < P.hdr.n_ops; ++o)
        if (P.ops[o].type == OP_DIAG) {
            const cplx* src = reinterpret_cast<const cplx*>(blob + P.ops[o].tbl_off);
            if ((diag_mode & 3) == 2) {  // per-thread phases, shared by the groups
                for (uint32_t i = tid; i < kGT; i += Cfg::kThreads) diag_smem[P.ops[o].diag_index * kGT + i] = diag_thread_phase(P.ops[o], src, i);
            } else if ((diag_mode & 3) == 1) {
                for (uint32_t i = tid; i < (uint32_t)kDiagTblLen; i += Cfg::kThreads) diag_smem[P.ops[o].diag_index * kDiagTblLen + i] = src[i];
            }
            // a phase without terms outside the tile is the same for every tile
            if (P.ops[o].ext_slot == kNoExtSlot && gtid == (o & 31u))
                ext_phase2[P.ops[o].diag_index] = ext_phase2[n_diag + 1 + P.ops[o].diag_index] = diag_ext_phase_terms(P.ops[o].theta0, nullptr, 0, 0);
        }
    for (uint32_t r = tid / kGT; r < n_rounds && r < kLbTabRounds; r += kG)  // group g fills rounds g, g + kG, ...
        if (P.rounds[r].type != ROUND_DENSE) lbtab[r * kGT + gtid] = round_thread_base(P.rounds[r], gtid);
    const uint64_t gstore_t = direct ? deposit(round_thread_base(P.rounds[n_rounds - 1], gtid), P.hdr.tile_segs, P.hdr.n_tile_segs) : 0;
    const uint32_t soff_t = swz(gtid) << 4;
    // warp 0 of the group: a tile that left through its buffer; the refill of that buffer waits until the bulk store has
    // read it, which is checked one tile later so that the warp does not sit on the store
    bool refill_pending = false;
    uint32_t refill_t = 0, refill_slot = 0, refill_use = 0;
    __syncthreads();

    // tiles are dealt round-robin to the groups (every tile of a pass costs the same); buffer and use number of the
    // group's current tile are carried along instead of recomputed (k mod kNB, k / kNB)
    uint32_t slot = group % kNB, use = group / kNB, parity = 0;
    // init mode 2 walks the tiles that hold amplitudes (PassInit::hold_id_*), dealt evenly over all groups of all CTAs
    const bool hold_enum = init.mode == 2;
    const uint64_t n_hold = 1ull << __popc(init.hold_id_mask);
    for (uint32_t it = 0;; ++it, parity ^= 1u) {
        const uint32_t k = group + it * kG;
        uint32_t t_id;
        if (hold_enum) {
            const uint64_t q = (uint64_t)blockIdx.x * kG + group + (uint64_t)it * gridDim.x * kG;
            if (q >= n_hold) break;
            t_id = init.hold_id_val | pdep32((uint32_t)q, init.hold_id_mask);
        } else {
            if (k >= n_my) break;
            t_id = tile_of(k);
        }
        uint64_t base = 0;
        if (need_base) base = tma_tile_base(tt, t_id);
        const uint64_t base_full = base | rank_hi;
        cplx* ext_phase = ext_phase2 + (size_t)parity * (n_diag + 1);
        const DiagCtx ctx{blob, ext_phase, (diag_mode & 3) == 1 ? diag_smem : nullptr, (diag_mode & 3) == 2 ? diag_smem : nullptr, kGT};
        cplx* tile;
        if (init.mode) {
            const bool holds = init_tile_holds(init, base_full);  // uniform over the group
            if (init.mode == 2 && !holds) continue;  // zero tile in, zero tile out: written by the stream above
            tile = tiles + (size_t)group * kTileLen;
            if (gtid < 32u && refill_pending) {  // the group's previous tile left through this buffer: wait until the bulk store has read it
                if (lane == 0) bulk_wait_read_all();
                refill_pending = false;
            }
            group_barrier(group, kGT);
            // synthesise the tile where a load would have put it (thread gtid owns tile-local elements i*kGT + gtid)
            char* tb = reinterpret_cast<char*>(tile);
#pragma unroll
            for (uint32_t i = 0; i < (uint32_t)kSlots; ++i)
                *reinterpret_cast<cplx*>(tb + (soff_t ^ P.loads.soff[i])) = holds ? init_tile_element(init, P.hdr, base, i * kGT + gtid) : cplx{0.0, 0.0};
            for (uint32_t o = gtid; o < P.hdr.n_ops; o += kGT)
                if (P.ops[o].type == OP_DIAG) ext_phase[P.ops[o].diag_index] = diag_ext_phase(P.ops[o], blob, base_full);
        } else {
            tile = tiles + (size_t)slot * kTileLen;
            wait_issued(&issued[slot], use + 1u);     // the load of this tile has been issued ...
            mbar_wait(&full_bar[slot], use & 1u);     // ... and has landed in shared memory
            // tile phases from the two table entries per op that arrived with the tile (warp 1: warp 0 issues the loads)
            const cplx* raw = ext_raw + (size_t)slot * n_ext_ops * 2u;
            if (gtid >= 32u && gtid < 64u)
                for (uint32_t j = lane; j < n_ext_ops; j += 32u) ext_phase[tt.ext_diag[j]] = cmul(raw[2 * j], raw[2 * j + 1]);
        }
        uint32_t act[W];
#pragma unroll
        for (int w = 0; w < W; ++w) act[w] = thr_act[w];
        if constexpr (!FAST) tile_active_mask<W>(P.hdr, P.ops, base_full, act);
        group_barrier(group, kGT);  // the tile phases (and a synthesised tile) are written
        if (refill_pending && !init.mode && gtid < 32u) {  // the previous tile's bulk store has had a barrier's time to read its buffer
            if (lane == 0) bulk_wait_read_all();
            __syncwarp();
            issue_load(refill_t, refill_slot, refill_use);
            refill_pending = false;
        }

        for (uint32_t r = 0; r < n_rounds; ++r) {
            const bool last = r + 1 == n_rounds;
            if (LEAN || P.rounds[r].type != ROUND_DENSE) {  // (FAST and LEAN passes hold register rounds only)
                const bool perm = !LEAN && P.rounds[r].type == ROUND_PERM;
                const uint32_t lb = r < kLbTabRounds ? lbtab[r * kGT + gtid] : round_thread_base(P.rounds[r], gtid);
                cplx a[kSlots];
                if constexpr (!LEAN) {
                    if (perm) round_perm_load<W, false>(P.rounds[r], P.ops, act, lb, tile, a);
                    else round_load(P.rounds[r], lb, tile, a);
                } else {
                    round_load(P.rounds[r], lb, tile, a);
                }
                // a permutation round gathers from all over the tile: everybody has read before anybody writes
                if (perm && !(direct && last && !init.mode)) group_barrier(group, kGT);
                if (direct && last && !init.mode) {
                    // the tile now lives in registers: hand the buffer to the tile that will use it next, a whole
                    // round of arithmetic before this group comes back for more
                    group_barrier(group, kGT);
                    if (gtid < 32u && k + kNB < n_my) issue_load(tile_of(k + kNB), slot, use + 1u);
                }
                round_ops<W, FAST>(P.rounds[r], P.ops, ctx, act, gtid, a);  // (a permutation round may carry ordinary ops behind its gather)
                if (direct && last) {
                    cplx* g = state + base + gstore_t;
#pragma unroll
                    for (int s = 0; s < kSlots; ++s) {
                        cplx v = a[s];
                        if (final_scale != 1.0) {
                            v.x *= final_scale;
                            v.y *= final_scale;
                        }
                        st_stream(g + P.loads.store_goff[s], v);
                    }
                } else if (last) {
                    round_store_tile_scaled(P.rounds[r], lb, tile, a, final_scale);
                } else {
                    round_store_tile(P.rounds[r], lb, tile, a);
                }
            } else if constexpr (!LEAN) {
                const DevDense& D = *reinterpret_cast<const DevDense*>(blob + P.ops[P.rounds[r].first_op].dense_off);
                cplx out[kSlots];
                dense_compute(D, blob, gtid, tile, out);
                group_barrier(group, kGT);
                if (last) {
#pragma unroll
                    for (int s = 0; s < kSlots; ++s) {
                        out[s].x *= final_scale;
                        out[s].y *= final_scale;
                    }
                }
                dense_store(gtid, tile, out);
            }
            if (!(direct && last)) {
                if (last) fence_proxy_async();  // the bulk store below reads what this thread wrote
                group_barrier(group, kGT);
            }
        }
        if (!direct && gtid < 32u) {
            // the finished tile leaves through its buffer; the buffer's refill is issued one tile later (see refill_pending)
            if (lane == 0) store_tile(tile, t_id);
            if (init.mode) {
                refill_pending = true;  // init mode: only the wait before the group's buffer is synthesised again
            } else if (k + kNB < n_my) {
                refill_pending = true;
                refill_t = tile_of(k + kNB);
                refill_slot = slot;
                refill_use = use + 1u;
            }
        }
        slot += kG;
        if (slot >= kNB) {
            slot -= kNB;
            ++use;
        }
    }
    if (refill_pending && !init.mode && gtid < 32u) {  // the group's last tiles: their buffers are still owed a refill
        if (lane == 0) bulk_wait_read_all();
        __syncwarp();
        issue_load(refill_t, refill_slot, refill_use);
    }
    // bulk stores of this thread must have read their shared-memory source before the CTA's shared memory goes away
    if (lane == 0 && gtid < 32u) bulk_wait_all();
}

// ---- host side ---------------------------------------------------------------------------------------------------

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libqsv.so links the runtime only; the driver's encoder is looked up through it
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

cudaError_t encode_tile_map(const TmaTileDesc& d, cplx* state, CUtensorMap* map) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return cudaErrorNotSupported;
    cuuint64_t dims[kTmaRank], strides[kTmaRank - 1];
    cuuint32_t box[kTmaRank], estr[kTmaRank];
    for (int i = 0; i < kTmaRank; ++i) {
        dims[i] = d.dim[i];
        box[i] = d.box[i];
        estr[i] = 1;
        if (i) strides[i - 1] = d.stride_bytes[i];
    }
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, kTmaRank, state, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int T>
size_t tma_fixed_smem(const DevPass& hdr) {
    using Cfg = TmaCfg<T>;
    return 1024 /* alignment slack */ + (size_t)Cfg::kBuffers * Cfg::kTileBytes + 64 /* mbarriers */ + 64 /* counters */ +
           sizeof(uint32_t) * kLbTabRounds * Cfg::kGroupThreads +
           sizeof(cplx) * ((size_t)Cfg::kBuffers * 2u * hdr.n_ext_ops + (size_t)Cfg::kGroups * 2u * (hdr.n_diag + 1));
}
constexpr size_t kSmemLimit = (size_t)227 * 1024;

}  // namespace

template <int T, int NR, int NO, int KIND>
static cudaError_t launch_tma_t(cplx* state, const uint8_t* dev_blob, const uint8_t* host_blob, const cplx* ext_tbl, uint64_t rank_hi, uint32_t n_alloc, int sm_count,
                                const PassInit& init, cudaStream_t stream, const PassSlice* slice, int grid_sms) {
    using Cfg = TmaCfg<T>;
    static thread_local PassParams<NR, NO> params;  // one staging buffer per host thread (include/qsv.h threading contract)
    if (!fill_params(host_blob, params)) return cudaErrorInvalidValue;
    const DevPass& hdr = params.hdr;
    if (hdr.tile_bits != (uint32_t)T) return cudaErrorInvalidValue;
    TmaTileDesc desc;
    if (!make_tma_tile(hdr, n_alloc, desc)) return cudaErrorInvalidValue;
    if (!set_tma_slice(hdr, slice, desc.tile) || (desc.tile.slice_n && init.mode)) return cudaErrorInvalidValue;
    desc.tile.n_ext_ops = hdr.n_ext_ops;
    desc.tile.tbl_a_bits = ext_table_low_bits(hdr.n_tiles);
    for (uint32_t o = 0; o < hdr.n_ops; ++o)
        if (params.ops[o].type == OP_DIAG && params.ops[o].ext_slot != kNoExtSlot) desc.tile.ext_diag[params.ops[o].ext_slot] = (uint8_t)params.ops[o].diag_index;
    if (hdr.n_ext_ops && !ext_tbl && !init.mode) return cudaErrorInvalidValue;
    alignas(64) CUtensorMap map;
    cudaError_t err = encode_tile_map(desc, state, &map);
    if (err != cudaSuccess) return err;
    const size_t fixed = tma_fixed_smem<T>(hdr);
    int mode = (fixed + sizeof(cplx) * kDiagTblLen * hdr.n_diag <= kSmemLimit) ? 1 : 0;  // DIAG tables in shared memory when they fit
    constexpr bool FAST = KIND == 1;
    if (FAST) mode = 2;
    if (hdr.n_diag == 0) mode = 0;
    const size_t smem = fixed + (mode == 1 ? sizeof(cplx) * kDiagTblLen * hdr.n_diag : mode == 2 ? sizeof(cplx) * Cfg::kGroupThreads * hdr.n_diag : 0);
    if (smem > kSmemLimit) return cudaErrorInvalidValue;
    static std::atomic<uint64_t> configured{0};
    err = ensure_dynamic_smem(pass_kernel_tma<T, NR, NO, KIND>, (int)kSmemLimit, configured);
    if (err != cudaSuccess) return err;
    // tile interleave (see the kernel): the groups of a CTA take tiles with consecutive ids; QSV_TILE_INTERLEAVE=0 turns it off
    static const int interleave = getenv("QSV_TILE_INTERLEAVE") ? atoi(getenv("QSV_TILE_INTERLEAVE")) : 1;
    const uint64_t n_tiles = hdr.n_tiles >> desc.tile.slice_n;  // tiles of this launch
    uint32_t ilog = 0;
    if (interleave)
        while ((1u << (ilog + 1)) <= Cfg::kGroups && (n_tiles >> (ilog + 1)) >= 1) ++ilog;
    for (uint32_t i = 0; i < desc.tile.slice_n; ++i)
        if (desc.tile.slice_pos[i] < ilog) return cudaErrorInvalidValue;  // the groups' consecutive tile ids must stay inside the slice
    uint64_t grid = (uint64_t)(grid_sms > 0 && grid_sms < sm_count ? grid_sms : sm_count);
    if (grid > (n_tiles >> ilog)) grid = n_tiles >> ilog;
    static const int tma_store = getenv("QSV_TMA_STORE") ? atoi(getenv("QSV_TMA_STORE")) : 0;  // developer A/B switch: 1 = no register->global stores
    pass_kernel_tma<T, NR, NO, KIND><<<(unsigned)grid, Cfg::kThreads, smem, stream>>>(map, state, dev_blob, ext_tbl, rank_hi, mode | (tma_store ? 4 : 0) | (int)(ilog << 8), init, desc.tile, params);
    return cudaGetLastError();
}

template <>
bool pass_tma_supported_tile<QSV_TILE_BITS>(const uint8_t* host_blob, uint32_t n_alloc) {
    const DevPass& hdr = *reinterpret_cast<const DevPass*>(host_blob);
    TmaTileDesc desc;
    if (hdr.tile_bits != (uint32_t)QSV_TILE_BITS || !make_tma_tile(hdr, n_alloc, desc)) return false;
    return tma_fixed_smem<QSV_TILE_BITS>(hdr) <= kSmemLimit && encode_tiled_fn() != nullptr;
}

template <>
cudaError_t launch_pass_tma_tile<QSV_TILE_BITS>(cplx* state, const uint8_t* dev_blob, const uint8_t* host_blob, const cplx* ext_tbl, uint64_t rank_hi, uint32_t n_alloc,
                                                int sm_count, const PassInit& init, cudaStream_t stream, const PassSlice* slice, int grid_sms) {
    const DevPass& hdr = *reinterpret_cast<const DevPass*>(host_blob);
    using Cfg = TmaCfg<QSV_TILE_BITS>;
    const bool small = hdr.n_rounds <= (uint32_t)kSmallRounds && hdr.n_ops <= (uint32_t)kSmallOps;
    const size_t fast_smem = tma_fixed_smem<QSV_TILE_BITS>(hdr) + sizeof(cplx) * Cfg::kGroupThreads * hdr.n_diag;
    static const bool no_fast = getenv("QSV_NO_FAST") != nullptr;  // developer A/B switches
    static const bool no_lean = getenv("QSV_NO_LEAN") != nullptr;
    if (!no_fast && small && (hdr.flags & PASS_UNCONDITIONAL) && fast_smem <= kSmemLimit)
        return launch_tma_t<QSV_TILE_BITS, kSmallRounds, kSmallOps, 1>(state, dev_blob, host_blob, ext_tbl, rank_hi, n_alloc, sm_count, init, stream, slice, grid_sms);
    bool reg_only = true;  // every round a register round: the build without dense and permutation rounds
    const DevRound* rounds = reinterpret_cast<const DevRound*>(host_blob + hdr.rounds_off);
    for (uint32_t r = 0; r < hdr.n_rounds; ++r) reg_only &= rounds[r].type == ROUND_REG;
    if (small && reg_only && !no_lean)
        return launch_tma_t<QSV_TILE_BITS, kSmallRounds, kSmallOps, 2>(state, dev_blob, host_blob, ext_tbl, rank_hi, n_alloc, sm_count, init, stream, slice, grid_sms);
    if (small) return launch_tma_t<QSV_TILE_BITS, kSmallRounds, kSmallOps, 0>(state, dev_blob, host_blob, ext_tbl, rank_hi, n_alloc, sm_count, init, stream, slice, grid_sms);
    return launch_tma_t<QSV_TILE_BITS, kMaxRounds, kMaxOps, 0>(state, dev_blob, host_blob, ext_tbl, rank_hi, n_alloc, sm_count, init, stream, slice, grid_sms);
}

}  // namespace qsv
