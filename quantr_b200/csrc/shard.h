// shard.h — NCCL plumbing for sharded registers (one process per GPU).
//
// The reference has no distributed code (README.md:96 "No parallelisation option"); this is new.
// NCCL is loaded with dlopen at first use so that single-GPU users need no NCCL at all and so that
// a process that already carries torch's bundled libnccl.so.2 shares that copy.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <string>

namespace qsv {

struct ShardComm;

bool shard_unique_id(void* out, size_t out_bytes, std::string& err);
ShardComm* shard_comm_create(int rank, int world, const void* unique_id, size_t unique_id_bytes, cudaStream_t stream, std::string& err);
void shard_comm_destroy(ShardComm* c);
bool shard_allreduce_sum(ShardComm* c, double* value, std::string& err);
// stream-ordered barrier over all ranks (a one-element all-reduce enqueued on the comm's stream)
bool shard_barrier(ShardComm* c, std::string& err);
// device buffers, enqueued on the comm's stream
bool shard_allreduce_sum_f64(ShardComm* c, double* d_buf, size_t count, std::string& err);
bool shard_allreduce_min_u64(ShardComm* c, uint64_t* d_buf, size_t count, std::string& err);
// host in/out: out[r] = rank r's value (synchronises the stream)
bool shard_allgather_f64(ShardComm* c, double value, double* out, std::string& err);
// Global-qubit remap on a shard of 2^n_local 16-byte amplitudes at `base`: rank bit j <-> local bit partner[j]
// (ascending).  Staged through two buffers of `staging_bytes` each.  Enqueued on the comm's stream.
bool shard_exchange_bits(ShardComm* c, void* base, uint32_t n_local, const uint8_t* partner, uint32_t g, void* staging, size_t staging_bytes,
                         std::string& err);

}  // namespace qsv
