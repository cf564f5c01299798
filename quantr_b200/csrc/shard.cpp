// shard.cpp — dlopen'ed NCCL: communicator set-up, scalar all-reduce and the global-qubit slot exchange.
#include "shard.h"

#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

namespace qsv {

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

bool load_nccl(std::string& err) {
    if (g_nccl.handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {  // prefer a copy the process already holds (torch bundles its own)
        h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h)
        for (const char* n : names) {
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
    if (!h) {
        err = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found");
        return false;
    }
#define QSV_SYM(field, name)                                                   \
    *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, name);                 \
    if (!g_nccl.field) { err = std::string("NCCL symbol missing: ") + name; return false; }
    QSV_SYM(GetUniqueId, "ncclGetUniqueId");
    QSV_SYM(CommInitRank, "ncclCommInitRank");
    QSV_SYM(CommDestroy, "ncclCommDestroy");
    QSV_SYM(AllReduce, "ncclAllReduce");
    QSV_SYM(AllGather, "ncclAllGather");
    QSV_SYM(Send, "ncclSend");
    QSV_SYM(Recv, "ncclRecv");
    QSV_SYM(GroupStart, "ncclGroupStart");
    QSV_SYM(GroupEnd, "ncclGroupEnd");
    QSV_SYM(GetErrorString, "ncclGetErrorString");
#undef QSV_SYM
    g_nccl.handle = h;
    return true;
}

bool nccl_ok(ncclResult_t r, const char* what, std::string& err) {
    if (r == ncclSuccess) return true;
    err = std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
    return false;
}

}  // namespace

struct ShardComm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    cudaStream_t stream = nullptr;
    double* d_scalar = nullptr;  // 1 + world doubles
    // exchange pipeline: staging -> shard copies run on their own stream so they overlap the next chunk's transfer
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t recv_done[2] = {nullptr, nullptr}, copy_done[2] = {nullptr, nullptr};
};

bool shard_unique_id(void* out, size_t out_bytes, std::string& err) {
    if (!out || out_bytes < sizeof(ncclUniqueId)) { err = "unique-id buffer must hold 128 bytes"; return false; }
    if (!load_nccl(err)) return false;
    ncclUniqueId id;
    if (!nccl_ok(g_nccl.GetUniqueId(&id), "ncclGetUniqueId", err)) return false;
    memcpy(out, &id, sizeof(id));
    return true;
}

ShardComm* shard_comm_create(int rank, int world, const void* unique_id, size_t unique_id_bytes, cudaStream_t stream, std::string& err) {
    if (!unique_id || unique_id_bytes < sizeof(ncclUniqueId)) { err = "nccl_unique_id must be the 128 bytes from qsv_nccl_unique_id"; return nullptr; }
    if (!load_nccl(err)) return nullptr;
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ShardComm* c = new ShardComm();
    c->rank = rank;
    c->world = world;
    c->stream = stream;
    if (!nccl_ok(g_nccl.CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank", err)) { delete c; return nullptr; }
    if (cudaMalloc(&c->d_scalar, sizeof(double) * (size_t)(1 + world)) != cudaSuccess) { err = "cudaMalloc failed"; g_nccl.CommDestroy(c->comm); delete c; return nullptr; }
    cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2; ++i) {
        cudaEventCreateWithFlags(&c->recv_done[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->copy_done[i], cudaEventDisableTiming);
    }
    return c;
}

void shard_comm_destroy(ShardComm* c) {
    if (!c) return;
    if (c->d_scalar) cudaFree(c->d_scalar);
    for (int i = 0; i < 2; ++i) {
        if (c->recv_done[i]) cudaEventDestroy(c->recv_done[i]);
        if (c->copy_done[i]) cudaEventDestroy(c->copy_done[i]);
    }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    delete c;
}

bool shard_allreduce_sum(ShardComm* c, double* value, std::string& err) {
    if (cudaMemcpyAsync(c->d_scalar, value, sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { err = "cudaMemcpyAsync failed"; return false; }
    if (!nccl_ok(g_nccl.AllReduce(c->d_scalar, c->d_scalar, 1, ncclFloat64, ncclSum, c->comm, c->stream), "ncclAllReduce", err)) return false;
    if (cudaMemcpyAsync(value, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) {
        err = "all-reduce of a scalar failed";
        return false;
    }
    return true;
}

bool shard_barrier(ShardComm* c, std::string& err) {
    return nccl_ok(g_nccl.AllReduce(c->d_scalar, c->d_scalar, 1, ncclFloat64, ncclSum, c->comm, c->stream), "ncclAllReduce (barrier)", err);
}

bool shard_allreduce_sum_f64(ShardComm* c, double* d_buf, size_t count, std::string& err) {
    return nccl_ok(g_nccl.AllReduce(d_buf, d_buf, count, ncclFloat64, ncclSum, c->comm, c->stream), "ncclAllReduce", err);
}

bool shard_allreduce_min_u64(ShardComm* c, uint64_t* d_buf, size_t count, std::string& err) {
    return nccl_ok(g_nccl.AllReduce(d_buf, d_buf, count, ncclUint64, ncclMin, c->comm, c->stream), "ncclAllReduce", err);
}

bool shard_allgather_f64(ShardComm* c, double value, double* out, std::string& err) {
    if (cudaMemcpyAsync(c->d_scalar, &value, sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { err = "cudaMemcpyAsync failed"; return false; }
    if (!nccl_ok(g_nccl.AllGather(c->d_scalar, c->d_scalar + 1, 1, ncclFloat64, c->comm, c->stream), "ncclAllGather", err)) return false;
    if (cudaMemcpyAsync(out, c->d_scalar + 1, sizeof(double) * (size_t)c->world, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
        err = "all-gather of a scalar failed";
        return false;
    }
    return true;
}

bool shard_exchange_bits(ShardComm* c, void* base, uint32_t n_local, const uint8_t* partner, uint32_t g, void* staging, size_t staging_bytes,
                         std::string& err) {
    if ((1 << g) != c->world) { err = "exchange: partner count does not match the communicator"; return false; }
    char* b = static_cast<char*>(base);
    char* stage[2] = {static_cast<char*>(staging), static_cast<char*>(staging) + staging_bytes};
    const uint64_t amp = 16;
    const uint64_t run = (uint64_t)1 << partner[0];             // contiguous amplitudes below the lowest partner bit
    uint64_t partner_mask = 0;
    for (uint32_t j = 0; j < g; ++j) partner_mask |= (uint64_t)1 << partner[j];
    const uint64_t n_runs = ((uint64_t)1 << n_local) >> (partner[0] + g);  // runs per peer block
    int slot = 0;
    bool slot_used[2] = {false, false};
    // Round-robin pairing (peer = rank ^ step): every step is a perfect matching over NVSwitch.
    for (int step = 1; step < c->world; ++step) {
        const int peer = c->rank ^ step;
        uint64_t peer_bits = 0;  // the block whose partner bits spell `peer`
        for (uint32_t j = 0; j < g; ++j) if ((peer >> j) & 1) peer_bits |= (uint64_t)1 << partner[j];
        for (uint64_t r = 0; r < n_runs; ++r) {
            // r enumerates the index bits above partner[0] that are not partner bits
            uint64_t idx = 0, rest = r;
            for (uint32_t bit = partner[0]; bit < n_local; ++bit) {
                if ((partner_mask >> bit) & 1) continue;
                idx |= (rest & 1) << bit;
                rest >>= 1;
            }
            char* chunk = b + (idx | peer_bits) * amp;
            const uint64_t bytes = run * amp;
            for (uint64_t off = 0; off < bytes; off += staging_bytes) {
                const size_t len = (size_t)(bytes - off < staging_bytes ? bytes - off : staging_bytes);
                char* st = stage[slot];
                // the staging buffer is free again once the copy that last read it has finished
                if (slot_used[slot] && cudaStreamWaitEvent(c->stream, c->copy_done[slot], 0) != cudaSuccess) { err = "cudaStreamWaitEvent failed"; return false; }
                if (!nccl_ok(g_nccl.GroupStart(), "ncclGroupStart", err)) return false;
                if (!nccl_ok(g_nccl.Send(chunk + off, len, ncclUint8, peer, c->comm, c->stream), "ncclSend", err)) return false;
                if (!nccl_ok(g_nccl.Recv(st, len, ncclUint8, peer, c->comm, c->stream), "ncclRecv", err)) return false;
                if (!nccl_ok(g_nccl.GroupEnd(), "ncclGroupEnd", err)) return false;
                // staging -> shard on the copy stream: overlaps the next chunk's transfer (the send of this chunk is complete)
                if (cudaEventRecord(c->recv_done[slot], c->stream) != cudaSuccess || cudaStreamWaitEvent(c->copy_stream, c->recv_done[slot], 0) != cudaSuccess ||
                    cudaMemcpyAsync(chunk + off, st, len, cudaMemcpyDeviceToDevice, c->copy_stream) != cudaSuccess ||
                    cudaEventRecord(c->copy_done[slot], c->copy_stream) != cudaSuccess) {
                    err = "staging copy failed";
                    return false;
                }
                slot_used[slot] = true;
                slot ^= 1;
            }
        }
    }
    // whatever follows on the comm's stream (the next pass) must see the last staging -> shard copies
    for (int i = 0; i < 2; ++i)
        if (slot_used[i] && cudaStreamWaitEvent(c->stream, c->copy_done[i], 0) != cudaSuccess) { err = "cudaStreamWaitEvent failed"; return false; }
    return true;
}


}  // namespace qsv
