// pass_kernel_async.cu — the fused pass as a persistent, software-pipelined kernel: one CTA per SM, G independent
// compute groups (one thread per 16 amplitudes each) and a ring of G+1 (T = 12) or G+2 (T = 11) tile buffers in
// shared memory that are refilled with cp.async while the groups compute.  Compiled per tile size
// (-DQSV_TILE_BITS=11|12); launch_pass() prefers it for large registers.
//
// Why: with the synchronous kernel (pass_kernel.cu) a CTA's load phase (global -> shared + first barrier) is ~30 %
// of its time and only 2-4 CTAs fit per SM (128 registers/thread), so memory and arithmetic add up instead of
// overlapping (profiles/r01*).  Here the tile a group works on was requested while the other group(s) computed:
//   tile k of the CTA lives in buffer k mod NB; the group that finishes tile k refills that buffer with tile k+NB
//   (cp.async.cg, 16 B per thread-instruction, arriving on the buffer's mbarrier) and then takes the next tile in
//   sequence (shared-memory counter); it blocks on that buffer's mbarrier only if the copy has not landed yet.
//
// Replaces Circuit::apply_gate (src/circuit/simulation.rs:64-135) for a fused list of gates; the per-thread
// arithmetic is the same pass_core.h code as the synchronous kernel and the host emulation.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "kernels.h"
#include "pass_core.h"

#ifndef QSV_TILE_BITS
#error "compile with -DQSV_TILE_BITS=<11|12>"
#endif

namespace qsv {

template <int T>
struct AsyncCfg {
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
        if (spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
// the mbarrier receives one arrival from this thread once all its earlier cp.async have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void group_barrier(uint32_t group, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1u), "r"(threads) : "memory");
}
__device__ __forceinline__ void st_stream(cplx* p, cplx v) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y)); }

template <int T, int NR, int NO, bool FAST>
__global__ void __launch_bounds__(AsyncCfg<T>::kThreads, 1)
pass_kernel_async(cplx* __restrict__ state, const uint8_t* __restrict__ blob, uint64_t rank_hi, int diag_mode, const __grid_constant__ PassParams<NR, NO> P) {
    using Cfg = AsyncCfg<T>;
    constexpr uint32_t kGT = Cfg::kGroupThreads, kNB = Cfg::kBuffers, kG = Cfg::kGroups;
    constexpr uint32_t kTileLen = 1u << T;
    constexpr int W = (NO + 31) / 32;
    extern __shared__ __align__(128) uint8_t smem[];
    // layout: [kNB tiles][kNB mbarriers][kNB issue counters][per-group external phases][DIAG tables][external term lists]
    cplx* tiles = reinterpret_cast<cplx*>(smem);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)kNB * Cfg::kTileBytes);
    // issued[b] = number of tile loads issued into buffer b so far.  A consumer may only poll the buffer's mbarrier by
    // parity once the load it waits for has been issued: a group can run two tiles ahead of the group that refills
    // its next buffer, and a parity wait placed before that refill would be satisfied by the phase before last
    // (same parity) - stale data, then a miscounted phase and a wait that never ends.
    uint32_t* issued = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(full_bar) + 64);
    cplx* ext_all = reinterpret_cast<cplx*>(reinterpret_cast<uint8_t*>(full_bar) + 128);
    const uint32_t n_diag = P.hdr.n_diag;
    cplx* diag_smem = ext_all + (size_t)kG * (n_diag + 1);
    const uint32_t tbl_len = (diag_mode & 3) == 2 ? kGT : (diag_mode & 3) == 1 ? (uint32_t)kDiagTblLen : 0u;
    DiagExtTerm* ext_terms = reinterpret_cast<DiagExtTerm*>(diag_smem + (size_t)n_diag * tbl_len);  // diag_mode & 16: [n_diag][max_ext]

    const uint32_t tid = threadIdx.x, group = tid / kGT, gtid = tid % kGT;
    cplx* ext_phase = ext_all + (size_t)group * (n_diag + 1);
    const uint32_t n_tile_segs = P.hdr.n_tile_segs, n_ext_segs = P.hdr.n_ext_segs, n_rounds = P.hdr.n_rounds;
    const uint64_t goff_t = deposit(gtid, P.hdr.tile_segs, n_tile_segs);
    const uint32_t soff_t = swz(gtid) << 4;
    const double final_scale = P.hdr.final_scale;
    const bool direct = (P.hdr.flags & PASS_DIRECT_STORE) != 0;
    // tiles of this CTA: t_k = blockIdx.x + k * gridDim.x
    const uint64_t n_my = P.hdr.n_tiles > blockIdx.x ? (P.hdr.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    auto issue_load = [&](uint64_t k) {  // all kGT threads of the calling group: tile k -> buffer k mod kNB
        const uint32_t slot = (uint32_t)(k % kNB);
        const uint64_t base = deposit(blockIdx.x + k * gridDim.x, P.hdr.ext_segs, n_ext_segs);
        const cplx* gtile = state + base + goff_t;
        char* tb = reinterpret_cast<char*>(tiles + (size_t)slot * kTileLen);
#pragma unroll
        for (uint32_t i = 0; i < (uint32_t)kSlots; ++i) cp_async16(tb + (soff_t ^ P.loads.soff[i]), gtile + P.loads.goff[i]);
        cp_async_arrive(&full_bar[slot]);
        // published after the issuing group itself saw the buffer's previous phase complete (it consumed that tile)
        if (gtid == 0) st_release_shared(&issued[slot], (uint32_t)(k / kNB) + 1u);
    };

    if (tid == 0) {
        for (uint32_t b = 0; b < kNB; ++b) {
            mbar_init(&full_bar[b], kGT);
            issued[b] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // prologue: the first kNB tiles, spread over the groups
    for (uint32_t k = group; k < kNB; k += kG)
        if (k < n_my) issue_load(k);

    // ---- once per launch: thread-dependent pieces that do not depend on the tile ----------------------
    uint32_t thr_act[W];
    if constexpr (FAST) {
#pragma unroll
        for (int w = 0; w < W; ++w) thr_act[w] = 0xffffffffu;
    } else {
        thread_active_mask<W>(P.hdr, P.rounds, P.ops, gtid, thr_act);
    }
    if (diag_mode & 3) {
        for (uint32_t o = 0; o < P.hdr.n_ops; ++o)
            if (P.ops[o].type == OP_DIAG) {
                const cplx* src = reinterpret_cast<const cplx*>(blob + P.ops[o].tbl_off);
                if ((diag_mode & 3) == 2) {  // per-thread phases, shared by the groups
                    for (uint32_t i = tid; i < kGT; i += Cfg::kThreads) diag_smem[P.ops[o].diag_index * kGT + i] = diag_thread_phase(P.ops[o], src, i);
                } else {
                    for (uint32_t i = tid; i < (uint32_t)kDiagTblLen; i += Cfg::kThreads) diag_smem[P.ops[o].diag_index * kDiagTblLen + i] = src[i];
                }
            }
    }
    if (diag_mode & 16) {  // the term lists of the external phases, read once per tile by diag_ext_phase_terms
        for (uint32_t o = 0; o < P.hdr.n_ops; ++o)
            if (P.ops[o].type == OP_DIAG) {
                const DiagExtTerm* src = reinterpret_cast<const DiagExtTerm*>(blob + P.ops[o].ext_off);
                for (uint32_t i = tid; i < P.ops[o].n_ext; i += Cfg::kThreads) ext_terms[P.ops[o].diag_index * P.hdr.max_ext + i] = src[i];
            }
    }
    const DiagCtx ctx{blob, ext_phase, (diag_mode & 3) == 1 ? diag_smem : nullptr, (diag_mode & 3) == 2 ? diag_smem : nullptr, kGT};
    const uint64_t gstore_t = (direct && n_rounds) ? deposit(round_thread_base(P.rounds[n_rounds - 1], gtid), P.hdr.tile_segs, n_tile_segs) : 0;
    __syncthreads();

    // tiles are dealt round-robin to the groups (every tile of a pass costs the same)
    for (uint64_t k = group; k < n_my; k += kG) {
        const uint32_t slot = (uint32_t)(k % kNB);
        cplx* tile = tiles + (size_t)slot * kTileLen;
        const uint64_t base = deposit(blockIdx.x + k * gridDim.x, P.hdr.ext_segs, n_ext_segs);
        const uint64_t base_full = base | rank_hi;
        // external phases: op o on lane o / warps of warp o % warps, so no warp of the group lags behind
        for (uint32_t o = (gtid >> 5) + (kGT >> 5) * (gtid & 31u); o < P.hdr.n_ops; o += kGT)
            if (P.ops[o].type == OP_DIAG) {
                const DevOp& op = P.ops[o];
                const DiagExtTerm* terms = (diag_mode & 16) ? ext_terms + op.diag_index * P.hdr.max_ext : reinterpret_cast<const DiagExtTerm*>(blob + op.ext_off);
                ext_phase[op.diag_index] = diag_ext_phase_terms(op.theta0, terms, op.n_ext, base_full);
            }
        uint32_t act[W];
#pragma unroll
        for (int w = 0; w < W; ++w) act[w] = thr_act[w];
        tile_active_mask<W>(P.hdr, P.ops, base_full, act);
        wait_issued(&issued[slot], (uint32_t)(k / kNB) + 1u);   // the load of this tile has been issued ...
        mbar_wait(&full_bar[slot], (uint32_t)((k / kNB) & 1));  // ... and has landed in shared memory
        group_barrier(group, kGT);                              // ... and the external phases are written

        for (uint32_t r = 0; r < n_rounds; ++r) {
            if (P.rounds[r].type == ROUND_REG) {
                const uint32_t lb = round_thread_base(P.rounds[r], gtid);
                cplx a[kSlots];
                round_load(P.rounds[r], lb, tile, a);
                if (direct && r + 1 == n_rounds) {
                    // the tile now lives in registers: hand the buffer to the tile that will use it next, a whole
                    // round of arithmetic before this group comes back for more
                    group_barrier(group, kGT);
                    if (k + kNB < n_my) issue_load(k + kNB);
                }
                round_ops<W, FAST>(P.rounds[r], P.ops, ctx, act, gtid, a);
                if (direct && r + 1 == n_rounds) {
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
                } else {
                    round_store_tile(P.rounds[r], lb, tile, a);
                }
            } else {
                const DevDense& D = *reinterpret_cast<const DevDense*>(blob + P.ops[P.rounds[r].first_op].dense_off);
                cplx out[kSlots];
                dense_compute(D, blob, gtid, tile, out);
                group_barrier(group, kGT);
                dense_store(gtid, tile, out);
            }
            if (!(direct && r + 1 == n_rounds)) group_barrier(group, kGT);
        }
        if (!direct) {
            cplx* gtile = state + base + goff_t;
            const char* tb = reinterpret_cast<const char*>(tile);
#pragma unroll
            for (uint32_t i = 0; i < (uint32_t)kSlots; ++i) {
                cplx v = *reinterpret_cast<const cplx*>(tb + (soff_t ^ P.loads.soff[i]));
                if (final_scale != 1.0) {
                    v.x *= final_scale;
                    v.y *= final_scale;
                }
                st_stream(gtile + P.loads.goff[i], v);
            }
            group_barrier(group, kGT);  // every thread has read its part of the buffer
        }
        // refill this buffer with the tile that will use it next (direct-store passes did so in their last round)
        if (!(direct && n_rounds && P.rounds[n_rounds - 1].type == ROUND_REG) && k + kNB < n_my) issue_load(k + kNB);
    }
    // A group may run out of tiles while loads it issued for other groups are still in flight: their copies and
    // mbarrier arrivals must not be abandoned by an exiting thread.
    asm volatile("cp.async.wait_all;" ::: "memory");
}

template <int T, int NR, int NO, bool FAST>
static cudaError_t launch_async_t(cplx* state, const uint8_t* dev_blob, const uint8_t* host_blob, uint64_t rank_hi, int sm_count, cudaStream_t stream) {
    using Cfg = AsyncCfg<T>;
    static PassParams<NR, NO> params;
    if (!fill_params(host_blob, params)) return cudaErrorInvalidValue;
    const DevPass& hdr = params.hdr;
    if (hdr.tile_bits != (uint32_t)T) return cudaErrorInvalidValue;
    const size_t fixed = (size_t)Cfg::kBuffers * Cfg::kTileBytes + 64 /* mbarriers */ + 64 /* counters */ + sizeof(cplx) * Cfg::kGroups * (hdr.n_diag + 1);
    const size_t limit = (size_t)227 * 1024 - 1024;
    int mode = (fixed + sizeof(cplx) * kDiagTblLen * hdr.n_diag <= limit) ? 1 : 0;  // DIAG tables in shared memory when they fit
    if (FAST) mode = 2;
    if (hdr.n_diag == 0) mode = 0;
    size_t smem = fixed + (mode == 1 ? sizeof(cplx) * kDiagTblLen * hdr.n_diag : mode == 2 ? sizeof(cplx) * Cfg::kGroupThreads * hdr.n_diag : 0);
    const size_t terms_bytes = sizeof(DiagExtTerm) * (size_t)hdr.n_diag * hdr.max_ext;
    if (terms_bytes && smem + terms_bytes <= limit) {
        smem += terms_bytes;
        mode |= 16;
    }
    if (smem > limit) return cudaErrorInvalidValue;
    static bool configured = false;
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(pass_kernel_async<T, NR, NO, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit);
        if (err != cudaSuccess) return err;
        configured = true;
    }
    uint64_t grid = (uint64_t)sm_count;
    if (grid > hdr.n_tiles) grid = hdr.n_tiles;
    pass_kernel_async<T, NR, NO, FAST><<<(unsigned)grid, Cfg::kThreads, smem, stream>>>(state, dev_blob, rank_hi, mode | (getenv("QSV_SKIP_EXT") ? 8 : 0), params);
    return cudaGetLastError();
}

template <>
cudaError_t launch_pass_async_tile<QSV_TILE_BITS>(cplx* state, const uint8_t* dev_blob, const uint8_t* host_blob, uint64_t rank_hi, int sm_count, cudaStream_t stream) {
    const DevPass& hdr = *reinterpret_cast<const DevPass*>(host_blob);
    using Cfg = AsyncCfg<QSV_TILE_BITS>;
    const bool small = hdr.n_rounds <= (uint32_t)kSmallRounds && hdr.n_ops <= (uint32_t)kSmallOps;
    const size_t fast_smem = (size_t)Cfg::kBuffers * Cfg::kTileBytes + 128 + sizeof(cplx) * Cfg::kGroups * (hdr.n_diag + 1) + sizeof(cplx) * Cfg::kGroupThreads * hdr.n_diag;
    static const bool no_fast = getenv("QSV_NO_FAST") != nullptr;  // developer A/B switch
    if (!no_fast && small && (hdr.flags & PASS_UNCONDITIONAL) && fast_smem <= (size_t)227 * 1024 - 1024)
        return launch_async_t<QSV_TILE_BITS, kSmallRounds, kSmallOps, true>(state, dev_blob, host_blob, rank_hi, sm_count, stream);
    if (small) return launch_async_t<QSV_TILE_BITS, kSmallRounds, kSmallOps, false>(state, dev_blob, host_blob, rank_hi, sm_count, stream);
    return launch_async_t<QSV_TILE_BITS, kMaxRounds, kMaxOps, false>(state, dev_blob, host_blob, rank_hi, sm_count, stream);
}

}  // namespace qsv
