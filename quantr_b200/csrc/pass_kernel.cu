// pass_kernel.cu — the fused-pass kernel (K1/K2/K3), compiled once per tile size: -DQSV_TILE_BITS=10|11|12|13
// (static tile) or 0 (tile size read from the header, states below 2^10 amplitudes).  One object file per
// tile size keeps the build parallel; launch_pass() in kernels.cu dispatches between them.
//
// Replaces Circuit::apply_gate (src/circuit/simulation.rs:64-135) for a fused list of gates.
#include <cuda_runtime.h>
#include <stdint.h>

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

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

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
    static constexpr uint32_t kMinBlocks = T_STATIC >= 13 ? 1 : T_STATIC == 12 ? 2 : T_STATIC == 11 ? 4 : 8;
    static constexpr uint32_t kLoads = T_STATIC ? (uint32_t)kSlots : 8u;  // runtime T <= 9: 2^9 / 64
};

template <int T_STATIC, int NR, int NO>
__global__ void __launch_bounds__(TileCfg<T_STATIC>::kThreads, TileCfg<T_STATIC>::kMinBlocks)
pass_kernel(cplx* __restrict__ state, const uint8_t* __restrict__ blob, uint64_t rank_hi, const __grid_constant__ PassParams<NR, NO> P) {
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
    // small passes keep the DIAG thread-phase tables in shared memory (no global loads in the op loop)
    constexpr bool kTblSmem = (NO == kSmallOps);
    cplx* thr_tbl = kTblSmem ? ext_phase + (P.hdr.n_diag + 1) : nullptr;
    const uint32_t tid = threadIdx.x;
    if (kTblSmem) {
        for (uint32_t o = 0; o < P.hdr.n_ops; ++o)
            if (P.ops[o].type == OP_DIAG) {
                const cplx* src = reinterpret_cast<const cplx*>(blob + P.ops[o].tbl_off);
                for (uint32_t i = tid; i < (uint32_t)kDiagTblLen; i += kThreads) thr_tbl[P.ops[o].diag_index * kDiagTblLen + i] = src[i];
            }
    }
    const uint32_t n_tile_segs = P.hdr.n_tile_segs, n_ext_segs = P.hdr.n_ext_segs;
    const uint64_t goff_t = deposit(tid, P.hdr.tile_segs, n_tile_segs);
    const uint64_t pf_t = deposit(8ull * tid, P.hdr.tile_segs, n_tile_segs);
    const uint32_t soff_t = swz(tid) << 4;
    const double final_scale = P.hdr.final_scale;
    const bool l2_prefetch = (P.hdr.flags & PASS_L2_PREFETCH) != 0;

    uint32_t thr_act[W];
    thread_active_mask<W>(P.hdr, P.rounds, P.ops, tid, thr_act);

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
        if (l2_prefetch && t + gridDim.x < P.hdr.n_tiles && (T_STATIC || 8u * tid < tile_len)) {
            // pull the CTA's next tile into L2 while this one is processed: two 128-byte lines per thread
            const cplx* nxt = state + deposit(t + gridDim.x, P.hdr.ext_segs, n_ext_segs) + pf_t;
            prefetch_l2(nxt);
            if (T_STATIC) prefetch_l2(nxt + P.hdr.pf_step);
        }
        for (uint32_t o = tid; o < P.hdr.n_ops; o += kThreads)
            if (P.ops[o].type == OP_DIAG) ext_phase[P.ops[o].diag_index] = diag_ext_phase(P.ops[o], blob, base_full);
        uint32_t act[W];
#pragma unroll
        for (int w = 0; w < W; ++w) act[w] = thr_act[w];
        tile_active_mask<W>(P.hdr, P.ops, base_full, act);
        __syncthreads();

        for (uint32_t r = 0; r < P.hdr.n_rounds; ++r) {
            if (P.rounds[r].type == ROUND_REG) {
                if (T_STATIC || tid < groups) reg_round<W>(P.rounds[r], P.ops, blob, ext_phase, thr_tbl, act, tid, tile);
            } else {
                const DevDense& D = *reinterpret_cast<const DevDense*>(blob + P.ops[P.rounds[r].first_op].dense_off);
                cplx out[kSlots];
                if (T_STATIC || tid < groups) dense_compute(D, blob, tid, tile, out);
                __syncthreads();
                if (T_STATIC || tid < groups) dense_store(tid, tile, out);
            }
            __syncthreads();
        }

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

template <int T_STATIC, int NR, int NO>
static cudaError_t launch_pass_t(cplx* state, const uint8_t* dev_blob, const uint8_t* host_blob, uint64_t rank_hi, int sm_count, cudaStream_t stream) {
    static PassParams<NR, NO> params;  // zero-initialised; only the used prefix of rounds/ops is rewritten per launch
    if (!fill_params(host_blob, params)) return cudaErrorInvalidValue;
    const DevPass& hdr = params.hdr;
    constexpr uint32_t kThreads = TileCfg<T_STATIC>::kThreads;
    if (hdr.threads != kThreads) return cudaErrorInvalidValue;
    constexpr bool kTblSmem = (NO == kSmallOps);
    const size_t smem = sizeof(cplx) * (size_t(1) << hdr.tile_bits) + sizeof(cplx) * (hdr.n_diag + 1) + (kTblSmem ? sizeof(cplx) * kDiagTblLen * hdr.n_diag : 0);
    static int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        const size_t want = sizeof(cplx) * (size_t(1) << (T_STATIC ? T_STATIC : 9)) + sizeof(cplx) * (NO + 1) + (kTblSmem ? sizeof(cplx) * kDiagTblLen * NO : 0);
        cudaError_t err = cudaFuncSetAttribute(pass_kernel<T_STATIC, NR, NO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
        if (err != cudaSuccess) return err;
        int nb = 0;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pass_kernel<T_STATIC, NR, NO>, (int)kThreads, want);
        if (err != cudaSuccess) return err;
        blocks_per_sm = nb > 0 ? nb : 1;
    }
    uint64_t grid = (uint64_t)sm_count * (uint64_t)blocks_per_sm;
    if (grid > hdr.n_tiles) grid = hdr.n_tiles;
    pass_kernel<T_STATIC, NR, NO><<<(unsigned)grid, kThreads, smem, stream>>>(state, dev_blob, rank_hi, params);
    return cudaGetLastError();
}

template <>
cudaError_t launch_pass_tile<QSV_TILE_BITS>(cplx* state, const uint8_t* dev_blob, const uint8_t* host_blob, uint64_t rank_hi, int sm_count, cudaStream_t stream) {
    const DevPass& hdr = *reinterpret_cast<const DevPass*>(host_blob);
    if (hdr.n_rounds <= (uint32_t)kSmallRounds && hdr.n_ops <= (uint32_t)kSmallOps)
        return launch_pass_t<QSV_TILE_BITS, kSmallRounds, kSmallOps>(state, dev_blob, host_blob, rank_hi, sm_count, stream);
    return launch_pass_t<QSV_TILE_BITS, kMaxRounds, kMaxOps>(state, dev_blob, host_blob, rank_hi, sm_count, stream);
}

}  // namespace qsv
