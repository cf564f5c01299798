// pass_kernel_init.cu — the first pass of a plan on a basis state, with the initialisation fused in: the register is
// never read; every tile is synthesised in shared memory (all zero, except the one tile that holds the basis
// amplitude), run through the pass's rounds and stored.  In mode 2 an all-zero tile is written as zeros without any
// arithmetic (a linear pass maps a zero tile to a zero tile), so the pass costs one write of the register.
// Compiled per tile size (-DQSV_TILE_BITS=11|12).  OPT-IN (QSV_FUSED_INIT, state_api.cu): written at the end of round
// 1 without GPU time left to measure it; the host side and the tile synthesis are checked through the host emulation
// (tests/test_schedule_emu.py).  The default path (memset + set_amp + ordinary first pass) does not touch this file.
//
// Replaces SuperPosition::new_unchecked (src/circuit/states/super_positions_unchecked.rs:39-46) followed by the
// first gates of Circuit::apply_gate (src/circuit/simulation.rs:64-135).
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
struct InitCfg {
    static constexpr uint32_t kGroupThreads = 1u << (T - kRegBits);
    static constexpr uint32_t kGroups = (T >= 12) ? 2u : 4u;
    static constexpr uint32_t kThreads = kGroupThreads * kGroups;
    static constexpr uint32_t kTileBytes = (uint32_t)sizeof(cplx) << T;
};

__device__ __forceinline__ void init_group_barrier(uint32_t group, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1u), "r"(threads) : "memory");
}
__device__ __forceinline__ void init_st_stream(cplx* p, cplx v) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y)); }

// One 512-thread CTA per SM, kGroups independent compute groups with one tile buffer each (nothing to prefetch).
template <int T, int NR, int NO, bool FAST>
__global__ void __launch_bounds__(InitCfg<T>::kThreads, 1)
pass_kernel_init(cplx* __restrict__ state, const uint8_t* __restrict__ blob, uint64_t rank_hi, int diag_mode, PassInit init, const __grid_constant__ PassParams<NR, NO> P) {
    using Cfg = InitCfg<T>;
    constexpr uint32_t kGT = Cfg::kGroupThreads, kG = Cfg::kGroups;
    constexpr uint32_t kTileLen = 1u << T;
    constexpr int W = (NO + 31) / 32;
    extern __shared__ __align__(128) uint8_t smem[];
    // layout: [kG tiles][per-group external phases][DIAG tables][external term lists]
    cplx* tiles = reinterpret_cast<cplx*>(smem);
    cplx* ext_all = reinterpret_cast<cplx*>(smem + (size_t)kG * Cfg::kTileBytes);
    const uint32_t n_diag = P.hdr.n_diag;
    cplx* diag_smem = ext_all + (size_t)kG * (n_diag + 1);
    const uint32_t tbl_len = (diag_mode & 3) == 2 ? kGT : (diag_mode & 3) == 1 ? (uint32_t)kDiagTblLen : 0u;
    DiagExtTerm* ext_terms = reinterpret_cast<DiagExtTerm*>(diag_smem + (size_t)n_diag * tbl_len);

    const uint32_t tid = threadIdx.x, group = tid / kGT, gtid = tid % kGT;
    cplx* ext_phase = ext_all + (size_t)group * (n_diag + 1);
    cplx* tile = tiles + (size_t)group * kTileLen;
    char* tb = reinterpret_cast<char*>(tile);
    const uint32_t n_tile_segs = P.hdr.n_tile_segs, n_ext_segs = P.hdr.n_ext_segs, n_rounds = P.hdr.n_rounds;
    const uint64_t goff_t = deposit(gtid, P.hdr.tile_segs, n_tile_segs);
    const uint32_t soff_t = swz(gtid) << 4;
    const double final_scale = P.hdr.final_scale;
    const bool direct = (P.hdr.flags & PASS_DIRECT_STORE) != 0;
    const uint64_t n_my = P.hdr.n_tiles > blockIdx.x ? (P.hdr.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    // ---- once per launch (as in pass_kernel_async) ---------------------------------------------------------
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
                if ((diag_mode & 3) == 2) {
                    for (uint32_t i = tid; i < kGT; i += Cfg::kThreads) diag_smem[P.ops[o].diag_index * kGT + i] = diag_thread_phase(P.ops[o], src, i);
                } else {
                    for (uint32_t i = tid; i < (uint32_t)kDiagTblLen; i += Cfg::kThreads) diag_smem[P.ops[o].diag_index * kDiagTblLen + i] = src[i];
                }
            }
    }
    if (diag_mode & 16) {
        for (uint32_t o = 0; o < P.hdr.n_ops; ++o)
            if (P.ops[o].type == OP_DIAG) {
                const DiagExtTerm* src = reinterpret_cast<const DiagExtTerm*>(blob + P.ops[o].ext_off);
                for (uint32_t i = tid; i < P.ops[o].n_ext; i += Cfg::kThreads) ext_terms[P.ops[o].diag_index * P.hdr.max_ext + i] = src[i];
            }
    }
    const DiagCtx ctx{blob, ext_phase, (diag_mode & 3) == 1 ? diag_smem : nullptr, (diag_mode & 3) == 2 ? diag_smem : nullptr, kGT};
    const uint64_t gstore_t = (direct && n_rounds) ? deposit(round_thread_base(P.rounds[n_rounds - 1], gtid), P.hdr.tile_segs, n_tile_segs) : 0;
    __syncthreads();

    for (uint64_t k = group; k < n_my; k += kG) {
        const uint64_t base = deposit(blockIdx.x + k * gridDim.x, P.hdr.ext_segs, n_ext_segs);
        const uint64_t base_full = base | rank_hi;
        const bool holds = base_full == init.base_full;  // uniform over the group
        cplx* gtile = state + base + goff_t;
        if (init.mode == 2 && !holds) {  // zero tile in, zero tile out
#pragma unroll
            for (uint32_t i = 0; i < (uint32_t)kSlots; ++i) init_st_stream(gtile + P.loads.goff[i], cplx{0.0, 0.0});
            continue;
        }
        // synthesise the tile where the ordinary kernels load it (thread gtid owns tile-local elements i*kGT + gtid)
#pragma unroll
        for (uint32_t i = 0; i < (uint32_t)kSlots; ++i)
            *reinterpret_cast<cplx*>(tb + (soff_t ^ P.loads.soff[i])) = cplx{(holds && i * kGT + gtid == init.local) ? 1.0 : 0.0, 0.0};
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
        init_group_barrier(group, kGT);  // tile and external phases are written

        for (uint32_t r = 0; r < n_rounds; ++r) {
            if (P.rounds[r].type == ROUND_REG) {
                const uint32_t lb = round_thread_base(P.rounds[r], gtid);
                cplx a[kSlots];
                round_load(P.rounds[r], lb, tile, a);
                // last round of a direct-store pass: once every thread holds its amplitudes the buffer is free for the
                // next tile of this group
                if (direct && r + 1 == n_rounds) init_group_barrier(group, kGT);
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
                        init_st_stream(g + P.loads.store_goff[s], v);
                    }
                } else {
                    round_store_tile(P.rounds[r], lb, tile, a);
                }
            } else {
                const DevDense& D = *reinterpret_cast<const DevDense*>(blob + P.ops[P.rounds[r].first_op].dense_off);
                cplx out[kSlots];
                dense_compute(D, blob, gtid, tile, out);
                init_group_barrier(group, kGT);
                dense_store(gtid, tile, out);
            }
            if (!(direct && r + 1 == n_rounds)) init_group_barrier(group, kGT);
        }
        if (!direct) {
#pragma unroll
            for (uint32_t i = 0; i < (uint32_t)kSlots; ++i) {
                cplx v = *reinterpret_cast<const cplx*>(tb + (soff_t ^ P.loads.soff[i]));
                if (final_scale != 1.0) {
                    v.x *= final_scale;
                    v.y *= final_scale;
                }
                init_st_stream(gtile + P.loads.goff[i], v);
            }
            init_group_barrier(group, kGT);  // every thread has read its part of the buffer
        }
    }
}

template <int T, int NR, int NO, bool FAST>
static cudaError_t launch_init_t(cplx* state, const uint8_t* dev_blob, const uint8_t* host_blob, uint64_t rank_hi, int sm_count, const PassInit& init, cudaStream_t stream) {
    using Cfg = InitCfg<T>;
    static PassParams<NR, NO> params;
    if (!fill_params(host_blob, params)) return cudaErrorInvalidValue;
    const DevPass& hdr = params.hdr;
    if (hdr.tile_bits != (uint32_t)T) return cudaErrorInvalidValue;
    const size_t fixed = (size_t)Cfg::kGroups * Cfg::kTileBytes + sizeof(cplx) * Cfg::kGroups * (hdr.n_diag + 1);
    const size_t limit = (size_t)227 * 1024 - 1024;
    int mode = (fixed + sizeof(cplx) * kDiagTblLen * hdr.n_diag <= limit) ? 1 : 0;
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
        cudaError_t err = cudaFuncSetAttribute(pass_kernel_init<T, NR, NO, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit);
        if (err != cudaSuccess) return err;
        configured = true;
    }
    uint64_t grid = (uint64_t)sm_count;
    if (grid > hdr.n_tiles) grid = hdr.n_tiles;
    pass_kernel_init<T, NR, NO, FAST><<<(unsigned)grid, Cfg::kThreads, smem, stream>>>(state, dev_blob, rank_hi, mode, init, params);
    return cudaGetLastError();
}

template <>
cudaError_t launch_pass_init_tile<QSV_TILE_BITS>(cplx* state, const uint8_t* dev_blob, const uint8_t* host_blob, uint64_t rank_hi, int sm_count, const PassInit& init,
                                                 cudaStream_t stream) {
    using Cfg = InitCfg<QSV_TILE_BITS>;
    const DevPass& hdr = *reinterpret_cast<const DevPass*>(host_blob);
    const bool small = hdr.n_rounds <= (uint32_t)kSmallRounds && hdr.n_ops <= (uint32_t)kSmallOps;
    const size_t fast_smem = (size_t)Cfg::kGroups * Cfg::kTileBytes + sizeof(cplx) * Cfg::kGroups * (hdr.n_diag + 1) + sizeof(cplx) * Cfg::kGroupThreads * hdr.n_diag;
    if (small && (hdr.flags & PASS_UNCONDITIONAL) && fast_smem <= (size_t)227 * 1024 - 1024)
        return launch_init_t<QSV_TILE_BITS, kSmallRounds, kSmallOps, true>(state, dev_blob, host_blob, rank_hi, sm_count, init, stream);
    if (small) return launch_init_t<QSV_TILE_BITS, kSmallRounds, kSmallOps, false>(state, dev_blob, host_blob, rank_hi, sm_count, init, stream);
    return launch_init_t<QSV_TILE_BITS, kMaxRounds, kMaxOps, false>(state, dev_blob, host_blob, rank_hi, sm_count, init, stream);
}

}  // namespace qsv
