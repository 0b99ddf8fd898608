// NCCL behind the C ABI: the hash shuffle of uniqueness / foreign-key keys (DataFusion's RepartitionExec(Hash)
// under COUNT(DISTINCT ..) and the foreign-key join, constraints/uniqueness.rs:549-718, foreign_key.rs:165-172) as ONE
// call a Rust host can make — partition on the device, exchange the counts, ncclSend / ncclRecv of every part inside one
// group straight into the column buffer of the receiving rank's shard table. No tensors, no staging copies.
//
// libnccl is bound at run time (dlopen of libnccl.so.2 the first time a communicator is created): libtermgpu.so has no
// link-time dependency on it, single-GPU users never load it, and inside a PyTorch process the already loaded copy is
// the one that resolves. The host layer only has to carry the 128-byte ncclUniqueId from rank 0 to the other ranks.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include "engine.hpp"
#include "hashpart.hpp"

namespace tg {

namespace {

typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId {
    char internal[128];
};
enum { NCCL_INT64 = 4, NCCL_UINT8 = 1 };  // ncclDataType_t values of nccl.h (ncclInt8 0, ncclUint8 1, ncclInt32 2, ncclUint32 3, ncclInt64 4)

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    if (api.lib) return api;
    void* lib = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) throw Error(TG_ERR_NCCL, std::string("libnccl.so.2 could not be loaded: ") + (dlerror() ? dlerror() : "not found"));
    auto sym = [&](const char* n) {
        void* p = dlsym(lib, n);
        if (!p) throw Error(TG_ERR_NCCL, std::string("libnccl: symbol not found: ") + n);
        return p;
    };
    api.GetUniqueId = (int (*)(NcclUniqueId*))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(ncclComm_t*, int, NcclUniqueId, int))sym("ncclCommInitRank");
    api.CommDestroy = (int (*)(ncclComm_t))sym("ncclCommDestroy");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.Send = (int (*)(const void*, size_t, int, int, ncclComm_t, cudaStream_t))sym("ncclSend");
    api.Recv = (int (*)(void*, size_t, int, int, ncclComm_t, cudaStream_t))sym("ncclRecv");
    api.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))sym("ncclAllGather");
    api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    api.lib = lib;
    return api;
}

void nccl_check(int rc, const char* what) {
    if (rc != 0) throw Error(TG_ERR_NCCL, std::string(what) + ": " + nccl().GetErrorString(rc));
}
#define TG_NCCL(expr) nccl_check((expr), #expr)

size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

void comm_unique_id(void* id128) {
    NcclUniqueId id;
    TG_NCCL(nccl().GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
}

void comm_destroy(Engine& e) {
    if (e.comm) {
        nccl().CommDestroy((ncclComm_t)e.comm);
        e.comm = nullptr;
    }
    if (e.d_comm_counts) cudaFree(e.d_comm_counts);
    e.d_comm_counts = nullptr;
    e.comm_world = 0;
}

void comm_init(Engine& e, const void* id128, int world, int rank) {
    if (world < 1 || world > 64 || rank < 0 || rank >= world) throw Error(TG_ERR_INVALID_ARG, "comm: bad world / rank");
    std::lock_guard<std::mutex> g(e.mu);
    TG_CUDA(cudaSetDevice(e.device));
    comm_destroy(e);
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    TG_NCCL(nccl().CommInitRank(&c, world, id, rank));
    e.comm = c;
    e.comm_world = world;
    e.comm_rank = rank;
    TG_CUDA(cudaMalloc(&e.d_comm_counts, (size_t)(world + 1) * 8 * (size_t)(world + 1) + 256));
}

// records grouped by destination rank (engine-owned send buffer) -> this rank's shard: a new engine-owned table `shard_name`
// with ONE column `column` of dtype `dtype`. rec_bytes = 8 (keys) or 24 (fingerprint records). n_nulls: this rank's NULL
// rows; rank 0 accounts for the NULL rows of every rank (they become trailing NULL rows of its shard). Returns the rows.
static int64_t exchange_into_table(Engine& e, const uint8_t* d_send, const int64_t* counts, int64_t n_nulls, size_t rec_bytes,
                                   const std::string& shard_name, const std::string& column, int32_t dtype) {
    if (!e.comm) throw Error(TG_ERR_NCCL, "no communicator: call tg_comm_init first");
    const int world = e.comm_world, rank = e.comm_rank;
    ncclComm_t comm = (ncclComm_t)e.comm;
    if (e.tables.count(shard_name)) throw Error(TG_ERR_INVALID_ARG, "table '" + shard_name + "' already exists");
    // ---- counts: every rank's (world + 1) numbers (keys per destination, NULL rows) to every rank
    const int w1 = world + 1;
    std::vector<long long> mine((size_t)w1), all((size_t)w1 * world);
    for (int r = 0; r < world; ++r) mine[r] = counts[r];
    mine[world] = n_nulls;
    long long* d_mine = (long long*)e.d_comm_counts;
    long long* d_all = d_mine + w1;
    TG_CUDA(cudaMemcpyAsync(d_mine, mine.data(), (size_t)w1 * 8, cudaMemcpyHostToDevice, e.stream));
    TG_NCCL(nccl().AllGather(d_mine, d_all, (size_t)w1, NCCL_INT64, comm, e.stream));
    TG_CUDA(cudaMemcpyAsync(all.data(), d_all, (size_t)w1 * world * 8, cudaMemcpyDeviceToHost, e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    int64_t n_recv = 0, nulls_total = 0;
    std::vector<int64_t> recv((size_t)world);
    for (int r = 0; r < world; ++r) {
        recv[r] = all[(size_t)r * w1 + rank];
        n_recv += recv[r];
        nulls_total += all[(size_t)r * w1 + world];
    }
    const int64_t my_nulls = rank == 0 ? nulls_total : 0;
    const int64_t n_rows = n_recv + my_nulls;
    // ---- the shard table, engine-owned
    auto tab = std::make_unique<Table>();
    tab->eng = &e;
    tab->name = shard_name;
    Column* c = table_get_or_add(*tab, column, dtype);
    c->n_rows = n_rows;
    tab->n_rows = n_rows;
    const size_t val_b = round_up((size_t)std::max<int64_t>(n_rows, 1) * rec_bytes + 256, 256);
    c->values.p = e.dev_alloc(val_b);
    c->values.cap = val_b;
    c->values.owned = true;
    c->value_bytes = n_rows * (int64_t)rec_bytes;
    c->null_count = my_nulls;
    if (my_nulls) {
        const size_t bits_b = round_up((size_t)(n_rows + 7) / 8 + 256, 256);
        c->validity.p = e.dev_alloc(bits_b);
        c->validity.cap = bits_b;
        c->validity.owned = true;
        TG_CUDA(cudaMemsetAsync(c->validity.p, 0, bits_b, e.stream));
        TG_CUDA(cudaMemsetAsync(c->validity.p, 0xFF, (size_t)(n_recv / 8), e.stream));
        if (n_recv % 8) {
            const uint8_t last = (uint8_t)((1u << (n_recv % 8)) - 1u);
            TG_CUDA(cudaMemcpyAsync(c->validity.p + n_recv / 8, &last, 1, cudaMemcpyHostToDevice, e.stream));
        }
        // the NULL rows' value slots: zero (deterministic)
        TG_CUDA(cudaMemsetAsync(c->values.p + (size_t)n_recv * rec_bytes, 0, (size_t)my_nulls * rec_bytes, e.stream));
    }
    TG_CUDA(cudaMemsetAsync(c->values.p + (size_t)n_rows * rec_bytes, 0, val_b - (size_t)n_rows * rec_bytes, e.stream));  // tail padding
    // ---- the data: every part straight into its place
    TG_NCCL(nccl().GroupStart());
    size_t so = 0, ro = 0;
    for (int r = 0; r < world; ++r) {
        if (counts[r]) TG_NCCL(nccl().Send(d_send + so, (size_t)counts[r] * rec_bytes, NCCL_UINT8, r, comm, e.stream));
        if (recv[r]) TG_NCCL(nccl().Recv(c->values.p + ro, (size_t)recv[r] * rec_bytes, NCCL_UINT8, r, comm, e.stream));
        so += (size_t)counts[r] * rec_bytes;
        ro += (size_t)recv[r] * rec_bytes;
    }
    TG_NCCL(nccl().GroupEnd());
    TG_CUDA(cudaStreamSynchronize(e.stream));
    e.tables[shard_name] = std::move(tab);
    e.comm_bytes_sent += (uint64_t)(so - (size_t)counts[rank] * rec_bytes);
    return n_rows;
}

int64_t comm_shuffle_column(Engine& e, const std::string& table, const std::string& column, const std::string& shard_name) {
    std::lock_guard<std::mutex> g(e.mu);
    TG_CUDA(cudaSetDevice(e.device));
    if (!e.comm) throw Error(TG_ERR_NCCL, "no communicator: call tg_comm_init first");
    e.sync_copies();
    auto it = e.tables.find(table);
    if (it == e.tables.end()) throw Error(TG_ERR_TABLE_NOT_FOUND, "table '" + table + "' not found");
    Table& t = *it->second;
    Column* c = t.find(column);
    if (!c) throw Error(TG_ERR_COLUMN_NOT_FOUND, "Schema error: No field named " + column + ". Valid fields are " + t.valid_fields() + ".");
    if (c->dtype != TG_INT64 && c->dtype != TG_FLOAT64) throw Error(TG_ERR_TYPE_MISMATCH, "the multi-GPU key shuffle supports Int64 / Float64 key columns");
    uint64_t* keys = nullptr;
    std::vector<int64_t> counts((size_t)e.comm_world, 0);
    int64_t nulls = 0;
    int launches = 0;
    partition_keys_by_rank(e, *c, t.n_rows, e.comm_world, &keys, counts.data(), &nulls, launches);
    e.launches += launches;
    return exchange_into_table(e, (const uint8_t*)keys, counts.data(), nulls, 8, shard_name, column, c->dtype);
}

int64_t comm_shuffle_fingerprints(Engine& e, const std::string& table, const std::vector<std::string>& columns, const std::string& shard_name) {
    std::lock_guard<std::mutex> g(e.mu);
    TG_CUDA(cudaSetDevice(e.device));
    if (!e.comm) throw Error(TG_ERR_NCCL, "no communicator: call tg_comm_init first");
    e.sync_copies();
    auto it = e.tables.find(table);
    if (it == e.tables.end()) throw Error(TG_ERR_TABLE_NOT_FOUND, "table '" + table + "' not found");
    Table& t = *it->second;
    void* recs = nullptr;
    std::vector<int64_t> counts((size_t)e.comm_world, 0);
    int launches = 0;
    partition_fingerprints_by_rank(e, t, columns, e.comm_world, &recs, counts.data(), launches);
    e.launches += launches;
    return exchange_into_table(e, (const uint8_t*)recs, counts.data(), 0, 24, shard_name, "tg_fp", TG_FP128);
}

}  // namespace tg
