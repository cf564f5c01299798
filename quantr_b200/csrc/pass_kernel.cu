// pass_kernel.cu — the fused-pass kernel (K1/K2/K3), compiled once per tile size: -DQSV_TILE_BITS=10|11|12|13
// (static tile) or 0 (tile size read from the header, states below 2^10 amplitudes).  One object file per
// tile size keeps the build parallel; launch_pass() in kernels.cu dispatches between them.
//
// Replaces Circuit::apply_gate (src/circuit/simulation.rs:64-135) for a fused list of gates.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "kernels.h"
#include "pass_core.h"

#ifndef QSV_TILE_BITS
#error "compile with -DQSV_TILE_BITS=<0|10|11|12|13>"
#endif

namespace qsv {

__device__ __forceinline__ cplx ld_stream(const cplx* p) {
    const double2 v = __ldcs(reinterpret_cast<const double2*>(p));
    return cplx{v.x, v.y};
}
__device__ __forceinline__ void st_stream(cplx* p, cplx v) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y)); }

// ---------------------------------------------------------------------------------------------
// Fused pass.  One thread per register group of 16 amplitudes: a tile of 2^T amplitudes is worked on by
// 2^(T-4) threads (T = 12: 256 threads, 2 CTAs/SM; T = 11: 128 threads, 4 CTAs/SM; T = 13: 512 threads, 1 CTA/SM).
// The pass descriptor (header, load constants, rounds, ops) arrives by value in the kernel-parameter
// constant bank, so op fields are uniform constant operands instead of shared- or global-memory loads.
// Shared memory: [tile: 2^T cplx][external phases of the DIAG ops].
// T_STATIC = 0: tile size read from the header (states below 2^10 amplitudes, latency-bound anyway).
// ---------------------------------------------------------------------------------------------
template <int T_STATIC>
struct TileCfg {
    static constexpr uint32_t kThreads = T_STATIC ? (1u << (T_STATIC - kRegBits)) : (uint32_t)kSmallTileThreads;
    static constexpr uint32_t kMinBlocks = T_STATIC ? tile_min_blocks(T_STATIC) : 8;
    static constexpr uint32_t kLoads = T_STATIC ? (uint32_t)kSlots : 8u;  // runtime T <= 9: 2^9 / 64
};

// FAST (pass_is_fast): no op has thread-bit controls, so the op walk is CTA-uniform (scalar registers, uniform branches,
// op constants as uniform operands) and DIAG thread phases always come from the per-launch table (diag_mode 2).
template <int T_STATIC, int NR, int NO, bool FAST>
__global__ void __launch_bounds__(TileCfg<T_STATIC>::kThreads, TileCfg<T_STATIC>::kMinBlocks)
pass_kernel(cplx* __restrict__ state, const uint8_t* __restrict__ blob, uint64_t rank_hi, int diag_mode, const __grid_constant__ PassParams<NR, NO> P) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr uint32_t kThreads = TileCfg<T_STATIC>::kThreads;
    constexpr uint32_t kLoads = TileCfg<T_STATIC>::kLoads;
    constexpr int W = (NO + 31) / 32;
    const uint32_t T = T_STATIC ? (uint32_t)T_STATIC : P.hdr.tile_bits;
    const uint32_t tile_len = 1u << T;
    const uint32_t groups = tile_len >> kRegBits;
    cplx* tile = reinterpret_cast<cplx*>(smem);
    char* tb = reinterpret_cast<char*>(smem);
    cplx* ext_phase = reinterpret_cast<cplx*>(smem + sizeof(cplx) * tile_len);
    cplx* diag_smem = ext_phase + (P.hdr.n_diag + 1);  // mode 2: [n_diag][threads] thread phases; mode 1: [n_diag][48] tables
    const uint32_t tid = threadIdx.x;
    const bool active = T_STATIC || tid < groups;      // runtime-T kernel: more threads than register groups
    const uint32_t n_tile_segs = P.hdr.n_tile_segs, n_ext_segs = P.hdr.n_ext_segs, n_rounds = P.hdr.n_rounds;
    const uint64_t goff_t = deposit(tid, P.hdr.tile_segs, n_tile_segs);
    const uint32_t soff_t = swz(tid) << 4;
    const double final_scale = P.hdr.final_scale;
    const bool direct = (P.hdr.flags & PASS_DIRECT_STORE) != 0;

    // ---- once per launch: thread-dependent pieces that do not depend on the tile ----------------------
    uint32_t thr_act[W];
    if constexpr (FAST) {
#pragma unroll
        for (int w = 0; w < W; ++w) thr_act[w] = 0xffffffffu;
    } else {
        thread_active_mask<W>(P.hdr, P.rounds, P.ops, tid, thr_act);
    }
    if (diag_mode & 3) {
        for (uint32_t o = 0; o < P.hdr.n_ops; ++o)
            if (P.ops[o].type == OP_DIAG) {
                const cplx* src = reinterpret_cast<const cplx*>(blob + P.ops[o].tbl_off);
                if ((diag_mode & 3) == 2) {
                    if (active) diag_smem[P.ops[o].diag_index * kThreads + tid] = diag_thread_phase(P.ops[o], src, tid);
                } else {
                    for (uint32_t i = tid; i < (uint32_t)kDiagTblLen; i += kThreads) diag_smem[P.ops[o].diag_index * kDiagTblLen + i] = src[i];
                }
            }
    }
    const DiagCtx ctx{blob, ext_phase, (diag_mode & 3) == 1 ? diag_smem : nullptr, (diag_mode & 3) == 2 ? diag_smem : nullptr, kThreads};
    // direct store: element offset of this thread's 16 amplitudes in the last round
    const uint64_t gstore_t = (direct && n_rounds) ? deposit(round_thread_base(P.rounds[n_rounds - 1], tid), P.hdr.tile_segs, n_tile_segs) : 0;

    for (uint64_t t = blockIdx.x; t < P.hdr.n_tiles; t += gridDim.x) {
        const uint64_t base = deposit(t, P.hdr.ext_segs, n_ext_segs);
        const uint64_t base_full = base | rank_hi;
        cplx* gtile = state + base + goff_t;
        {
            cplx v[kLoads];
#pragma unroll
            for (uint32_t i = 0; i < kLoads; ++i)
                if (T_STATIC || i * kThreads + tid < tile_len) v[i] = ld_stream(gtile + P.loads.goff[i]);
#pragma unroll
            for (uint32_t i = 0; i < kLoads; ++i)
                if (T_STATIC || i * kThreads + tid < tile_len) *reinterpret_cast<cplx*>(tb + (soff_t ^ P.loads.soff[i])) = v[i];
        }
        for (uint32_t o = tid; o < P.hdr.n_ops; o += kThreads)
            if (P.ops[o].type == OP_DIAG) ext_phase[P.ops[o].diag_index] = (diag_mode & 8) ? cplx{1.0, 0.0} : diag_ext_phase(P.ops[o], blob, base_full);
        uint32_t act[W];
#pragma unroll
        for (int w = 0; w < W; ++w) act[w] = thr_act[w];
        tile_active_mask<W>(P.hdr, P.ops, base_full, act);
        __syncthreads();

        for (uint32_t r = 0; r < n_rounds; ++r) {
            if (P.rounds[r].type == ROUND_PERM) {  // gather through the tile: everybody reads before anybody writes
                const uint32_t lb = active ? round_thread_base(P.rounds[r], tid) : 0u;
                cplx a[kSlots];
                if (active) round_perm_load<W, FAST>(P.rounds[r], P.ops, act, lb, tile, a);
                __syncthreads();
                if (active) {
                    round_ops<W, FAST>(P.rounds[r], P.ops, ctx, act, tid, a);  // ordinary ops behind the gather, if any
                    if (direct && r + 1 == n_rounds) {
                        cplx* g = state + base + gstore_t;
#pragma unroll
                        for (int s = 0; s < kSlots; ++s) st_stream(g + P.loads.store_goff[s], cplx{a[s].x * final_scale, a[s].y * final_scale});
                    } else {
                        round_store_tile(P.rounds[r], lb, tile, a);
                    }
                }
            } else if (P.rounds[r].type == ROUND_REG) {
                if (active) {
                    const uint32_t lb = round_thread_base(P.rounds[r], tid);
                    cplx a[kSlots];
                    round_load(P.rounds[r], lb, tile, a);
                    round_ops<W, FAST>(P.rounds[r], P.ops, ctx, act, tid, a);
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
                }
            } else {
                const DevDense& D = *reinterpret_cast<const DevDense*>(blob + P.ops[P.rounds[r].first_op].dense_off);
                cplx out[kSlots];
                if (active) dense_compute(D, blob, tid, tile, out);
                __syncthreads();
                if (active) dense_store(tid, tile, out);
            }
            __syncthreads();
        }
        if (direct) continue;  // the barrier after the last round already protects the tile buffer

#pragma unroll
        for (uint32_t i = 0; i < kLoads; ++i) {
            if (T_STATIC || i * kThreads + tid < tile_len) {
                cplx v = *reinterpret_cast<const cplx*>(tb + (soff_t ^ P.loads.soff[i]));
                if (final_scale != 1.0) {
                    v.x *= final_scale;
                    v.y *= final_scale;
                }
                st_stream(gtile + P.loads.goff[i], v);
            }
        }
        __syncthreads();
    }
}

template <int T_STATIC, int NR, int NO, bool FAST>
static cudaError_t launch_pass_t(cplx* state, const uint8_t* dev_blob, const uint8_t* host_blob, uint64_t rank_hi, int sm_count, cudaStream_t stream) {
    // one staging buffer per host thread (handles may be driven from different threads, include/qsv.h); zero-initialised,
    // only the used prefix of rounds/ops is rewritten per launch
    static thread_local PassParams<NR, NO> params;
    if (!fill_params(host_blob, params)) return cudaErrorInvalidValue;
    const DevPass& hdr = params.hdr;
    constexpr uint32_t kThreads = TileCfg<T_STATIC>::kThreads;
    if (hdr.threads != kThreads) return cudaErrorInvalidValue;
    static const int mode_cap = getenv("QSV_DIAG_MODE") ? atoi(getenv("QSV_DIAG_MODE")) : 2;  // developer A/B switch
    int mode = choose_diag_mode(hdr.tile_bits, hdr.n_diag);
    if (mode > mode_cap) mode = mode_cap;
    if (FAST) mode = hdr.n_diag ? 2 : 0;
    const size_t smem = pass_smem_bytes(hdr.tile_bits, hdr.n_diag, mode);
    constexpr size_t smem_cfg = (size_t)(227 * 1024) - 1024;  // opt-in maximum; the per-launch size decides the occupancy
    static std::atomic<uint64_t> configured{0};
    cudaError_t err = ensure_dynamic_smem(pass_kernel<T_STATIC, NR, NO, FAST>, (int)smem_cfg, configured);
    if (err != cudaSuccess) return err;
    if (smem > smem_cfg) return cudaErrorInvalidValue;
    static const size_t pad = getenv("QSV_SMEM_PAD") ? (size_t)atol(getenv("QSV_SMEM_PAD")) : 0;  // occupancy experiments
    const size_t smem_launch = (smem + pad <= smem_cfg) ? smem + pad : smem;
    int nb = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pass_kernel<T_STATIC, NR, NO, FAST>, (int)kThreads, smem_launch);
    if (err != cudaSuccess) return err;
    uint64_t grid = (uint64_t)sm_count * (uint64_t)(nb > 0 ? nb : 1);
    if (grid > hdr.n_tiles) grid = hdr.n_tiles;
    pass_kernel<T_STATIC, NR, NO, FAST><<<(unsigned)grid, kThreads, smem_launch, stream>>>(state, dev_blob, rank_hi, mode | (getenv("QSV_SKIP_EXT") ? 8 : 0), params);
    return cudaGetLastError();
}

template <>
cudaError_t launch_pass_tile<QSV_TILE_BITS>(cplx* state, const uint8_t* dev_blob, const uint8_t* host_blob, uint64_t rank_hi, int sm_count, cudaStream_t stream) {
    const DevPass& hdr = *reinterpret_cast<const DevPass*>(host_blob);
#if QSV_TILE_BITS >= 10
    static const bool no_fast = getenv("QSV_NO_FAST") != nullptr;  // developer A/B switch
    if (!no_fast && pass_is_fast(hdr)) return launch_pass_t<QSV_TILE_BITS, kSmallRounds, kSmallOps, true>(state, dev_blob, host_blob, rank_hi, sm_count, stream);
#endif
    if (hdr.n_rounds <= (uint32_t)kSmallRounds && hdr.n_ops <= (uint32_t)kSmallOps)
        return launch_pass_t<QSV_TILE_BITS, kSmallRounds, kSmallOps, false>(state, dev_blob, host_blob, rank_hi, sm_count, stream);
    return launch_pass_t<QSV_TILE_BITS, kMaxRounds, kMaxOps, false>(state, dev_blob, host_blob, rank_hi, sm_count, stream);
}

}  // namespace qsv
