// kernels.h — launchers of the sm_100a kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "qsv_types.h"
#include "tma_tile.h"

namespace qsv { struct PassInit; }

namespace qsv {
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: opt in once per (kernel, device).  `mask` is a
// per-kernel static; bit d = done on device d.  Thread-safe (a lost race repeats an idempotent call).
template <class Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes, std::atomic<uint64_t>& mask) {
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    const uint64_t bit = 1ull << (dev & 63);
    if (mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (err != cudaSuccess) return err;
    mask.fetch_or(bit, std::memory_order_release);
    return cudaSuccess;
}
}  // namespace qsv

namespace qsv {

// host_blob: the pass blob in host memory (its header/rounds/ops travel as kernel parameters);
// dev_blob: the same blob in device memory (tables, external phase terms, Custom matrices);
// ext_tbl: the pass's external-phase tables in device memory (launch_build_ext_tables), may be null for passes without;
// init: null, or the basis state whose initialisation is fused into this pass (the register is not read);
// slice: null, or the part of the register this launch covers (pipelined kernel only, tma_tile.h PassSlice);
// grid_sms: 0, or the number of SMs the persistent kernel may occupy (the rest is left to a concurrent exchange).
cudaError_t launch_pass(cplx* state, const uint8_t* dev_blob, const uint8_t* host_blob, const cplx* ext_tbl, uint64_t rank_hi, uint32_t n_alloc, int sm_count,
                        const PassInit* init, cudaStream_t stream, const PassSlice* slice = nullptr, int grid_sms = 0);
// true when launch_pass sends this pass to the pipelined TMA kernel (large registers whose tile is a tensor-map box)
bool pass_uses_tma(const uint8_t* host_blob, uint32_t n_alloc, int sm_count);
// true when launch_pass accepts `init` for this pass
bool pass_init_supported(const uint8_t* host_blob, uint32_t n_alloc, int sm_count);
// fills the pass's external-phase tables: hdr.n_ext_ops x ext_table_len(hdr.n_tiles) entries (pass_core.h ext_table_entry)
cudaError_t launch_build_ext_tables(const uint8_t* dev_blob, const uint8_t* host_blob, cplx* tbl, uint64_t rank_hi, cudaStream_t stream);

// defined in pass_kernel.cu, one explicit specialisation per tile size (0 = runtime tile size <= 9 bits): the synchronous kernel
template <int TILE_BITS>
cudaError_t launch_pass_tile(cplx* state, const uint8_t* dev_blob, const uint8_t* host_blob, uint64_t rank_hi, int sm_count, cudaStream_t stream);

// defined in pass_kernel_tma.cu for TILE_BITS = 11 and 12: the software-pipelined TMA kernel
template <int TILE_BITS>
cudaError_t launch_pass_tma_tile(cplx* state, const uint8_t* dev_blob, const uint8_t* host_blob, const cplx* ext_tbl, uint64_t rank_hi, uint32_t n_alloc, int sm_count,
                                 const PassInit& init, cudaStream_t stream, const PassSlice* slice, int grid_sms);
template <int TILE_BITS>
bool pass_tma_supported_tile(const uint8_t* host_blob, uint32_t n_alloc);

cudaError_t launch_set_amp(cplx* state, uint64_t index, double re, double im, cudaStream_t stream);
// state[(j << (n_local - sup_bits)) | low] = tbl[j], j < 2^sup_bits: the amplitudes a folded prefix left (Plan::prefix_local_bits)
cudaError_t launch_scatter_prefix(cplx* state, const cplx* tbl, uint32_t sup_bits, uint32_t n_local, uint64_t low, cudaStream_t stream);
// in-place swap of this rank's block 'spelled' peer with the peer's block 'spelled' rank (peer-mapped memory, NVLink)
// slice: null, or the part of the shard to swap (the blocks' amplitudes whose index bits slice->bit[] spell slice->value)
cudaError_t launch_peer_swap(cplx* local, cplx* remote, uint32_t n_local, const uint8_t* partner, uint32_t g, int rank, int peer, int sm_count, cudaStream_t stream,
                             const PassSlice* slice = nullptr);
// Cross-GPU flags of the pipelined exchange (words in peer-mapped memory): every peer's word `index` is set to `value`
// (system-scope release) / the kernel returns once the local words index + r, r != rank, are all >= value (acquire).
struct FlagPeers {
    uint32_t* flags[16];  // flags[r] = rank r's flag words as mapped here (null for this rank and beyond the world size)
};
cudaError_t launch_flag_signal(const FlagPeers& peers, int world, uint32_t index, uint32_t value, cudaStream_t stream);
cudaError_t launch_flag_wait(const uint32_t* local_flags, int world, int rank, uint32_t index, uint32_t value, cudaStream_t stream);
cudaError_t launch_gather(const cplx* state, const uint64_t* idx, cplx* out, uint64_t count, cudaStream_t stream);
// canonical range of a register in a permuted qubit layout; entries held by other ranks come back as zero
cudaError_t launch_gather_range(const cplx* state, cplx* out, uint64_t first, uint64_t count, const uint8_t* layout, uint32_t n_qubits, uint32_t n_local, uint64_t rank,
                                int sm_count, cudaStream_t stream);
cudaError_t launch_prob_block_sums(const cplx* state, double* sums, uint64_t n_blocks, uint32_t block_bits, int sm_count, cudaStream_t stream);
// `prefix` must hold n + 2 + 2 * kScanMaxChunks doubles (prefix[n] = total, the rest is scratch of the chunked scan)
constexpr int kScanMaxChunks = 1024;
cudaError_t launch_scan_block_sums(const double* sums, double* prefix, uint64_t n, cudaStream_t stream);
cudaError_t launch_sample_shots(const cplx* state, const double* prefix, uint64_t n_blocks, uint32_t block_bits, const double* uniforms,
                                uint64_t shots, uint64_t index_or, uint64_t* out, int sm_count, cudaStream_t stream);

}  // namespace qsv
