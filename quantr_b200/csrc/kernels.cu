// kernels.cu — hand-written sm_100a kernels of the state-vector engine.
//
//   pass_kernel<T>        K1/K2/K3: one fused pass (tile of 2^T amplitudes per CTA in shared memory,
//                         16 amplitudes per thread in registers per round); replaces
//                         Circuit::apply_gate (src/circuit/simulation.rs:64-135) for a list of gates
//   set_amp / gather      K4: SuperPosition::new_unchecked (super_positions_unchecked.rs:39-46), get_state
//   prob_block_sums,      K6: the cumulative |amp|^2 loop of SuperPosition::measure
//   scan_block_sums           (src/circuit/states/super_positions.rs:335-336), hierarchically
//   sample_shots          K7: the inverse-CDF search of measure (:333-341), one warp per shot
//
// All kernels are HBM-bound streaming kernels: 128-bit coalesced accesses, grids sized to the SM count
// (persistent CTAs), no tensor-core work.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "kernels.h"
#include "peer_swap.h"
#include "pass_core.h"

namespace qsv {

__device__ __forceinline__ cplx ld_stream(const cplx* p) {
    const double2 v = __ldcs(reinterpret_cast<const double2*>(p));
    return cplx{v.x, v.y};
}
__device__ __forceinline__ void st_stream(cplx* p, cplx v) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y)); }

// ---------------------------------------------------------------------------------------------
// Fused pass: dispatch to the per-tile-size objects built from pass_kernel.cu.
// ---------------------------------------------------------------------------------------------
// QSV_ASYNC (read once): 0 = the synchronous kernel at every size, 2 = the pipelined TMA kernel wherever the tile is
// expressible (parity tests on small registers), default = the pipelined kernel from 16 tiles per SM upwards.
static int async_mode() {
    static const int mode = getenv("QSV_ASYNC") ? atoi(getenv("QSV_ASYNC")) : 1;
    return mode;
}

bool pass_uses_tma(const uint8_t* host_blob, uint32_t n_alloc, int sm_count) {
    const DevPass& hdr = *reinterpret_cast<const DevPass*>(host_blob);
    const int mode = async_mode();
    if (mode == 0 || (mode < 2 && hdr.n_tiles < 16ull * (uint64_t)sm_count)) return false;
    if (hdr.tile_bits == 12) return pass_tma_supported_tile<12>(host_blob, n_alloc);
    if (hdr.tile_bits == 11) return pass_tma_supported_tile<11>(host_blob, n_alloc);
    return false;
}

bool pass_init_supported(const uint8_t* host_blob, uint32_t n_alloc, int sm_count) { return pass_uses_tma(host_blob, n_alloc, sm_count); }

cudaError_t launch_pass(cplx* state, const uint8_t* dev_blob, const uint8_t* host_blob, const cplx* ext_tbl, uint64_t rank_hi, uint32_t n_alloc, int sm_count,
                        const PassInit* init, cudaStream_t stream, const PassSlice* slice, int grid_sms) {
    const DevPass& hdr = *reinterpret_cast<const DevPass*>(host_blob);
    const uint32_t tile_bits = hdr.tile_bits;
    if (pass_uses_tma(host_blob, n_alloc, sm_count)) {
        const PassInit none{0, 0, 0, 0, 0, 0, 1.0, 0.0};
        if (init && slice && slice->n) return cudaErrorNotSupported;
        if (tile_bits == 12) return launch_pass_tma_tile<12>(state, dev_blob, host_blob, ext_tbl, rank_hi, n_alloc, sm_count, init ? *init : none, stream, slice, grid_sms);
        return launch_pass_tma_tile<11>(state, dev_blob, host_blob, ext_tbl, rank_hi, n_alloc, sm_count, init ? *init : none, stream, slice, grid_sms);
    }
    if (init || (slice && slice->n)) return cudaErrorNotSupported;
    switch (tile_bits) {
        case 10: return launch_pass_tile<10>(state, dev_blob, host_blob, rank_hi, sm_count, stream);
        case 11: return launch_pass_tile<11>(state, dev_blob, host_blob, rank_hi, sm_count, stream);
        case 12: return launch_pass_tile<12>(state, dev_blob, host_blob, rank_hi, sm_count, stream);
        case 13: return launch_pass_tile<13>(state, dev_blob, host_blob, rank_hi, sm_count, stream);
        default:
            if (tile_bits >= (uint32_t)kRegBits && tile_bits <= 9) return launch_pass_tile<0>(state, dev_blob, host_blob, rank_hi, sm_count, stream);
            return cudaErrorInvalidValue;
    }
}

// External-phase tables of one pass (pipelined kernel): grid.y = table slot, grid.x strides over its 2^a + 2^b entries.
struct ExtSlotOps {
    uint8_t op[kMaxOps];
};
__global__ void __launch_bounds__(256) build_ext_tables_kernel(const uint8_t* __restrict__ blob, cplx* __restrict__ tbl, uint64_t rank_hi, uint32_t a_bits, uint64_t len_a,
                                                             uint64_t len, ExtSlotOps slots) {
    const DevPass& hdr = *reinterpret_cast<const DevPass*>(blob);
    const DevOp& op = reinterpret_cast<const DevOp*>(blob + hdr.ops_off)[slots.op[blockIdx.y]];
    const DiagExtTerm* terms = reinterpret_cast<const DiagExtTerm*>(blob + op.ext_off);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < len; i += (uint64_t)gridDim.x * blockDim.x) {
        const bool low = i < len_a;
        tbl[(uint64_t)blockIdx.y * len + i] = ext_table_entry(hdr, op.theta0, terms, op.n_ext, low ? i : (i - len_a) << a_bits, rank_hi, low);
    }
}

cudaError_t launch_build_ext_tables(const uint8_t* dev_blob, const uint8_t* host_blob, cplx* tbl, uint64_t rank_hi, cudaStream_t stream) {
    const DevPass& hdr = *reinterpret_cast<const DevPass*>(host_blob);
    if (hdr.n_ext_ops == 0) return cudaSuccess;
    const DevOp* ops = reinterpret_cast<const DevOp*>(host_blob + hdr.ops_off);
    ExtSlotOps slots{};
    for (uint32_t o = 0; o < hdr.n_ops; ++o)
        if (ops[o].type == OP_DIAG && ops[o].ext_slot != kNoExtSlot) slots.op[ops[o].ext_slot] = (uint8_t)o;
    const uint32_t a_bits = ext_table_low_bits(hdr.n_tiles);
    const uint64_t len = ext_table_len(hdr.n_tiles);
    uint64_t gx = (len + 255) / 256;
    if (gx > 64) gx = 64;
    build_ext_tables_kernel<<<dim3((unsigned)gx, hdr.n_ext_ops), 256, 0, stream>>>(dev_blob, tbl, rank_hi, a_bits, 1ull << a_bits, len, slots);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// K4: register access
// ---------------------------------------------------------------------------------------------
__global__ void set_amp_kernel(cplx* state, uint64_t index, cplx v) { state[index] = v; }

__global__ void gather_kernel(const cplx* __restrict__ state, const uint64_t* __restrict__ idx, cplx* __restrict__ out, uint64_t count) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) out[i] = state[idx[i]];
}

cudaError_t launch_set_amp(cplx* state, uint64_t index, double re, double im, cudaStream_t stream) {
    set_amp_kernel<<<1, 1, 0, stream>>>(state, index, cplx{re, im});
    return cudaGetLastError();
}

// Canonical range [first, first + count) of a register whose index bits are permuted (layout[b] = physical position of
// logical bit b, rank bits on top): this rank's amplitudes land in `out`, the others' entries are zero (the caller sums
// over ranks).  K4 "applying the logical->physical permutation" (SURVEY.md 2.2).
struct LayoutArg {
    uint8_t pos[64];
};
__global__ void __launch_bounds__(256) gather_range_kernel(const cplx* __restrict__ state, cplx* __restrict__ out, uint64_t first, uint64_t count, LayoutArg layout,
                                                         uint32_t n_qubits, uint32_t n_local, uint64_t rank) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t logical = first + i;
        uint64_t phys = 0;
        for (uint32_t b = 0; b < n_qubits; ++b) phys |= ((logical >> b) & 1ull) << layout.pos[b];
        out[i] = (phys >> n_local) == rank ? state[phys & ((1ull << n_local) - 1ull)] : cplx{0.0, 0.0};
    }
}

cudaError_t launch_gather_range(const cplx* state, cplx* out, uint64_t first, uint64_t count, const uint8_t* layout, uint32_t n_qubits, uint32_t n_local, uint64_t rank,
                                int sm_count, cudaStream_t stream) {
    LayoutArg la{};
    for (uint32_t b = 0; b < n_qubits && b < 64; ++b) la.pos[b] = layout[b];
    uint64_t grid = (count + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count * 8;
    if (grid > cap) grid = cap;
    gather_range_kernel<<<(unsigned)(grid ? grid : 1), 256, 0, stream>>>(state, out, first, count, la, n_qubits, n_local, rank);
    return cudaGetLastError();
}

cudaError_t launch_gather(const cplx* state, const uint64_t* idx, cplx* out, uint64_t count, cudaStream_t stream) {
    const unsigned grid = (unsigned)((count + 255) / 256 > 1184 ? 1184 : (count + 255) / 256);
    gather_kernel<<<grid ? grid : 1, 256, 0, stream>>>(state, idx, out, count);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// K5: global-qubit exchange over NVLink peer memory.  Swaps, for one peer, the block of this rank's shard whose
// partner bits spell the peer with the peer's block whose partner bits spell this rank - in place, no staging:
// each rank of the pair moves one half of the block, reading the remote half over NVLink and writing its own
// amplitudes back into the peer's memory with 128-bit loads/stores.
// ---------------------------------------------------------------------------------------------
template <int THREADS>
__global__ void __launch_bounds__(THREADS) peer_swap_kernel(cplx* __restrict__ local, cplx* __restrict__ remote, SwapArgs a) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j0 < a.count; j0 += 4 * stride) {
        double2 mine[4], theirs[4];
        uint64_t idx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t j = j0 + u * stride;
            idx[u] = insert_zero_bits(a.first + (j < a.count ? j : j0), a);
            mine[u] = *reinterpret_cast<const double2*>(local + (idx[u] | a.local_spell));
            theirs[u] = *reinterpret_cast<const double2*>(remote + (idx[u] | a.remote_spell));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (j0 + u * stride >= a.count) continue;
            *reinterpret_cast<double2*>(local + (idx[u] | a.local_spell)) = theirs[u];
            *reinterpret_cast<double2*>(remote + (idx[u] | a.remote_spell)) = mine[u];
        }
    }
}

// slice != null: the pipelined exchange - one 1024-thread CTA per SM on `sm_count` SMs (a persistent pass kernel owns the
// others; fat CTAs keep the two grids from sharing an SM, which the pass kernel's register file does not allow anyway)
cudaError_t launch_peer_swap(cplx* local, cplx* remote, uint32_t n_local, const uint8_t* partner, uint32_t g, int rank, int peer, int sm_count, cudaStream_t stream,
                             const PassSlice* slice) {
    const SwapArgs a = make_swap_args(n_local, partner, g, rank, peer, slice ? slice->n : 0u, slice ? slice->bit : nullptr, slice ? slice->value : 0u);
    if (a.count == 0) return cudaSuccess;
    if (slice && slice->n) {
        uint64_t grid = (a.count + 1024 * 4 - 1) / (1024 * 4);
        if (grid > (uint64_t)sm_count) grid = (uint64_t)sm_count;
        peer_swap_kernel<1024><<<(unsigned)grid, 1024, 0, stream>>>(local, remote, a);
        return cudaGetLastError();
    }
    uint64_t grid = (a.count + 256 * 4 - 1) / (256 * 4);
    const uint64_t cap = (uint64_t)sm_count * 8;
    if (grid > cap) grid = cap;
    peer_swap_kernel<256><<<(unsigned)grid, 256, 0, stream>>>(local, remote, a);
    return cudaGetLastError();
}

// A folded prefix written to HBM (registers whose first pass cannot synthesise it): state[(j << shift) | low] = tbl[j].
__global__ void scatter_prefix_kernel(cplx* __restrict__ state, const cplx* __restrict__ tbl, uint64_t count, uint32_t shift, uint64_t low) {
    for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < count; j += (uint64_t)gridDim.x * blockDim.x) state[(j << shift) | low] = tbl[j];
}
cudaError_t launch_scatter_prefix(cplx* state, const cplx* tbl, uint32_t sup_bits, uint32_t n_local, uint64_t low, cudaStream_t stream) {
    const uint64_t count = 1ull << sup_bits;
    scatter_prefix_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(state, tbl, count, n_local - sup_bits, low);
    return cudaGetLastError();
}

// Cross-GPU flags of the pipelined exchange.  A wait that outlasts any legitimate exchange (seconds) traps so that a
// protocol error surfaces as a launch failure instead of a hung device.
__global__ void flag_signal_kernel(FlagPeers peers, int world, uint32_t index, uint32_t value) {
    const int r = (int)threadIdx.x;
    if (r < world && peers.flags[r]) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers.flags[r] + index), "r"(value) : "memory");
    }
}
__global__ void flag_wait_kernel(const uint32_t* __restrict__ local_flags, int world, int rank, uint32_t index, uint32_t value) {
    const int r = (int)threadIdx.x;
    if (r < world && r != rank) {
        const uint32_t* p = local_flags + index + r;
        for (uint64_t spins = 0;; ++spins) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
            if ((int32_t)(v - value) >= 0) break;
            __nanosleep(200);
            if (spins > (1ull << 24)) __trap();
        }
    }
    __threadfence_system();
}
cudaError_t launch_flag_signal(const FlagPeers& peers, int world, uint32_t index, uint32_t value, cudaStream_t stream) {
    flag_signal_kernel<<<1, 32, 0, stream>>>(peers, world, index, value);
    return cudaGetLastError();
}
cudaError_t launch_flag_wait(const uint32_t* local_flags, int world, int rank, uint32_t index, uint32_t value, cudaStream_t stream) {
    flag_wait_kernel<<<1, 32, 0, stream>>>(local_flags, world, rank, index, value);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// K6: per-block probability sums and their exclusive scan.  Fixed summation tree -> deterministic.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One CTA per block of 2^block_bits amplitudes (grid-stride).  256 threads.
__global__ void __launch_bounds__(256) prob_block_sums_kernel(const cplx* __restrict__ state, double* __restrict__ sums, uint64_t n_blocks, uint32_t block_bits) {
    __shared__ double warp_part[8];
    const uint32_t len = 1u << block_bits;
    for (uint64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const cplx* p = state + (b << block_bits);
        double acc = 0.0;
        for (uint32_t l = threadIdx.x; l < len; l += 256) {
            const cplx a = ld_stream(p + l);
            acc += a.x * a.x + a.y * a.y;
        }
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < 8; ++w) s += warp_part[w];
            sums[b] = s;
        }
        __syncthreads();
    }
}

// Single CTA, 1024 threads: exclusive scan of `n` doubles with a running carry; prefix[n] = total.
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(const double* __restrict__ sums, double* __restrict__ prefix, uint64_t n) {
    __shared__ double warp_tot[32];
    __shared__ double carry_s;
    if (threadIdx.x == 0) carry_s = 0.0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint64_t base = 0; base < n; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const double v = i < n ? sums[i] : 0.0;
        double inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += y;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            double w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= (uint32_t)o) w += y;
            }
            warp_tot[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const double carry = carry_s;
        const double before = carry + (warp ? warp_tot[warp - 1] : 0.0) + (inc - v);
        if (i < n) prefix[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) prefix[n] = carry_s;
}

// ---------------------------------------------------------------------------------------------
// K7: one warp per shot.  prefix[b] = sum of the blocks before b (b = 0..n_blocks), prefix[n_blocks] = total.
// Returns the first canonical index i with u < cumulative(i), UINT64_MAX if u >= total
// (super_positions.rs:341 "None").
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_shots_kernel(const cplx* __restrict__ state, const double* __restrict__ prefix, uint64_t n_blocks,
                                                         uint32_t block_bits, const double* __restrict__ uniforms, uint64_t shots,
                                                         uint64_t index_or, uint64_t* __restrict__ out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp_global = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t len = 1u << block_bits;
    for (uint64_t s = warp_global; s < shots; s += n_warps) {
        const double u = uniforms[s];
        uint64_t result = UINT64_MAX;
        if (u >= 0.0 && u < prefix[n_blocks]) {  // negative: a shot that belongs to a lower rank of a sharded register
            // largest b with prefix[b] <= u  (prefix is non-decreasing, prefix[0] = 0 <= u)
            uint64_t lo = 0, hi = n_blocks;  // invariant: prefix[lo] <= u, (hi == n_blocks or prefix[hi] > u)
            while (hi - lo > 1) {
                const uint64_t mid = (lo + hi) >> 1;
                if (prefix[mid] <= u) lo = mid; else hi = mid;
            }
            for (uint64_t b = lo; b < n_blocks && result == UINT64_MAX; ++b) {
                double carry = prefix[b];
                const cplx* p = state + (b << block_bits);
                for (uint32_t l0 = 0; l0 < len; l0 += 32) {
                    double inc = 0.0;
                    if (l0 + lane < len) {
                        const cplx a = p[l0 + lane];
                        inc = a.x * a.x + a.y * a.y;
                    }
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const double y = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= (uint32_t)o) inc += y;
                    }
                    const double cum = carry + inc;
                    const uint32_t hit = __ballot_sync(0xffffffffu, u < cum);
                    if (hit) {
                        result = (b << block_bits) + l0 + (uint32_t)(__ffs(hit) - 1);
                        break;
                    }
                    carry = __shfl_sync(0xffffffffu, cum, 31);
                }
            }
        }
        if (lane == 0) out[s] = (result == UINT64_MAX) ? result : (result | index_or);
    }
}

cudaError_t launch_prob_block_sums(const cplx* state, double* sums, uint64_t n_blocks, uint32_t block_bits, int sm_count, cudaStream_t stream) {
    uint64_t grid = (uint64_t)sm_count * 8;
    if (grid > n_blocks) grid = n_blocks;
    prob_block_sums_kernel<<<(unsigned)grid, 256, 0, stream>>>(state, sums, n_blocks, block_bits);
    return cudaGetLastError();
}

// Large registers (2^21 block sums at 33 qubits): three short launches instead of one CTA walking the whole array.
//   chunk_scan   one CTA per chunk of 4096 sums: exclusive scan inside the chunk, chunk total to `chunk_tot`
//   (the single-CTA kernel above scans the <= 1024 chunk totals in place; its total is prefix[n])
//   chunk_add    prefix[i] += offset of i's chunk
constexpr uint32_t kScanChunk = 4096;
__global__ void __launch_bounds__(1024) chunk_scan_kernel(const double* __restrict__ sums, double* __restrict__ prefix, double* __restrict__ chunk_tot, uint64_t n) {
    __shared__ double warp_tot[32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t i0 = (uint64_t)blockIdx.x * kScanChunk + 4ull * threadIdx.x;
    double v[4], run = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        v[j] = i0 + j < n ? sums[i0 + j] : 0.0;
        run += v[j];
    }
    double inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += y;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        double w = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= (uint32_t)o) w += y;
        }
        warp_tot[lane] = w;
    }
    __syncthreads();
    double before = (warp ? warp_tot[warp - 1] : 0.0) + (inc - run);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (i0 + j < n) prefix[i0 + j] = before;
        before += v[j];
    }
    if (threadIdx.x == 1023) chunk_tot[blockIdx.x] = warp_tot[31];
}
__global__ void __launch_bounds__(1024) chunk_add_kernel(double* __restrict__ prefix, const double* __restrict__ chunk_off, uint64_t n) {
    const double off = chunk_off[blockIdx.x];
    const uint64_t i0 = (uint64_t)blockIdx.x * kScanChunk + 4ull * threadIdx.x;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (i0 + j < n) prefix[i0 + j] += off;
}

// prefix must have room for n + 2 + 2 * kScanMaxChunks doubles: prefix[n] = total, the rest is chunk scratch
cudaError_t launch_scan_block_sums(const double* sums, double* prefix, uint64_t n, cudaStream_t stream) {
    const uint64_t chunks = (n + kScanChunk - 1) / kScanChunk;
    if (chunks <= 2 || chunks > (uint64_t)kScanMaxChunks) {
        scan_block_sums_kernel<<<1, 1024, 0, stream>>>(sums, prefix, n);
        return cudaGetLastError();
    }
    double* chunk_tot = prefix + n + 1;
    double* chunk_off = chunk_tot + kScanMaxChunks;  // exclusive scan of the chunk totals; chunk_off[chunks] = total
    chunk_scan_kernel<<<(unsigned)chunks, 1024, 0, stream>>>(sums, prefix, chunk_tot, n);
    scan_block_sums_kernel<<<1, 1024, 0, stream>>>(chunk_tot, chunk_off, chunks);
    chunk_add_kernel<<<(unsigned)chunks, 1024, 0, stream>>>(prefix, chunk_off, n);
    cudaError_t e = cudaMemcpyAsync(prefix + n, chunk_off + chunks, sizeof(double), cudaMemcpyDeviceToDevice, stream);
    return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_sample_shots(const cplx* state, const double* prefix, uint64_t n_blocks, uint32_t block_bits, const double* uniforms,
                                uint64_t shots, uint64_t index_or, uint64_t* out, int sm_count, cudaStream_t stream) {
    uint64_t grid = (shots + 7) / 8;
    const uint64_t cap = (uint64_t)sm_count * 8;
    if (grid > cap) grid = cap;
    if (grid == 0) grid = 1;
    sample_shots_kernel<<<(unsigned)grid, 256, 0, stream>>>(state, prefix, n_blocks, block_bits, uniforms, shots, index_or, out);
    return cudaGetLastError();
}

}  // namespace qsv
