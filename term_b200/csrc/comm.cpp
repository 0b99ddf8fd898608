// NCCL behind the C ABI: the hash shuffle of uniqueness / foreign-key keys (DataFusion's RepartitionExec(Hash)
// under COUNT(DISTINCT ..) and the foreign-key join, constraints/uniqueness.rs:549-718, foreign_key.rs:165-172) as ONE
// call a Rust host can make — partition on the device, exchange the counts, ncclSend / ncclRecv of every part inside one
// group straight into the column buffer of the receiving rank's shard table. No tensors, no staging copies.
//
// libnccl is bound at run time (dlopen of libnccl.so.2 the first time a communicator is created): libtermgpu.so has no
// link-time dependency on it, single-GPU users never load it, and inside a PyTorch process the already loaded copy is
// the one that resolves. The host layer only has to carry the 128-byte ncclUniqueId from rank 0 to the other ranks.
#include <dlfcn.h>

#include <chrono>
#include <cstdio>

#include <algorithm>
#include <cstdlib>
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

void comm_push_destroy(Engine& e);

void comm_destroy(Engine& e) {
    comm_push_destroy(e);
    if (e.comm) {
        nccl().CommDestroy((ncclComm_t)e.comm);
        e.comm = nullptr;
    }
    if (e.d_comm_counts) cudaFree(e.d_comm_counts);
    e.d_comm_counts = nullptr;
    if (e.d_comm_samples) cudaFree(e.d_comm_samples);
    e.d_comm_samples = nullptr;
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

// ---------------------------------------------------------------------------------------------------------------------
// Push shuffle: partition and all-to-all as ONE kernel. Every rank owns receive buffers that all peers of the node have
// mapped through CUDA IPC; after the ranks have exchanged their per-destination counts (a (world + 1)-number all-gather)
// each rank knows where its part for rank d starts in d's buffer, and the partition's scatter kernel writes every part
// straight there — the coalesced runs leave the SM as NVLink stores, the transfer overlaps the partition tile by tile and
// no intermediate copy of the keys exists anywhere. A second tiny collective is the barrier after which a rank may read
// its buffer. Falls back to partition + ncclSend / ncclRecv when peer mapping is not available (TG_NO_PUSH_SHUFFLE).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int PUSH_SHUFFLE_SLOTS = 2;  // shards alive at the same time (a foreign key shuffles the child and the parent column)
constexpr int PUSH_SLOTS = PUSH_SHUFFLE_SLOTS + 4;  // + keys / payload receive buffers of the two phases of the distributed sort

struct PushSlot {
    uint8_t* local = nullptr;
    size_t cap = 0;
    std::vector<uint8_t*> peers;      // every rank's buffer as seen from this device
    uint64_t** d_ptrs = nullptr;      // the same as a device array (argument of the scatter kernel)
    std::string user;                 // shard table living in the buffer
};
struct PushState {
    PushSlot slot[PUSH_SLOTS];
    bool disabled = false;
};

static void push_slot_release(Engine& e, PushSlot& s) {
    for (size_t r = 0; r < s.peers.size(); ++r)
        if ((int)r != e.comm_rank && s.peers[r]) cudaIpcCloseMemHandle(s.peers[r]);
    if (s.local) cudaFree(s.local);
    if (s.d_ptrs) cudaFree(s.d_ptrs);
    s = PushSlot{};
}
void comm_push_destroy(Engine& e) {
    if (!e.push) return;
    PushState* ps = (PushState*)e.push;
    for (auto& s : ps->slot) push_slot_release(e, s);
    delete ps;
    e.push = nullptr;
}

// (re)allocate slot `s` with `cap` bytes on every rank and map the peers' buffers: collective
static void push_slot_grow(Engine& e, PushSlot& s, size_t cap) {
    const int world = e.comm_world, rank = e.comm_rank;
    ncclComm_t comm = (ncclComm_t)e.comm;
    TG_CUDA(cudaStreamSynchronize(e.stream));
    push_slot_release(e, s);
    TG_CUDA(cudaMalloc(&s.local, cap));
    s.cap = cap;
    cudaIpcMemHandle_t mine;
    TG_CUDA(cudaIpcGetMemHandle(&mine, s.local));
    uint8_t* d_mine = e.d_comm_counts;  // staging: (world + 1) x 64 bytes fit ((world + 1)^2 x 8 + 256 were allocated; world >= 7 ... checked below)
    uint8_t* d_all = nullptr;
    TG_CUDA(cudaMalloc(&d_all, (size_t)(world + 1) * 64));
    std::vector<cudaIpcMemHandle_t> all((size_t)world);
    try {
        TG_CUDA(cudaMemcpyAsync(d_all + (size_t)world * 64, &mine, 64, cudaMemcpyHostToDevice, e.stream));
        TG_NCCL(nccl().AllGather(d_all + (size_t)world * 64, d_all, 64, NCCL_UINT8, comm, e.stream));
        TG_CUDA(cudaMemcpyAsync(all.data(), d_all, (size_t)world * 64, cudaMemcpyDeviceToHost, e.stream));
        TG_CUDA(cudaStreamSynchronize(e.stream));
    } catch (...) {
        cudaFree(d_all);
        throw;
    }
    cudaFree(d_all);
    (void)d_mine;
    s.peers.assign((size_t)world, nullptr);
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            s.peers[r] = s.local;
            continue;
        }
        void* p = nullptr;
        TG_CUDA(cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess));
        s.peers[r] = (uint8_t*)p;
    }
    TG_CUDA(cudaMalloc(&s.d_ptrs, (size_t)world * 8));
    TG_CUDA(cudaMemcpy(s.d_ptrs, s.peers.data(), (size_t)world * 8, cudaMemcpyHostToDevice));
}

// returns -1 when the push path is not available (the caller then takes the send / recv path)
static int64_t push_shuffle_column(Engine& e, Table& t, Column& c, const std::string& shard_name, bool allow_range) {
    static const bool off = getenv("TG_NO_PUSH_SHUFFLE") != nullptr;
    if (off) return -1;
    if (!e.push) e.push = new PushState();
    PushState& ps = *(PushState*)e.push;
    if (ps.disabled) return -1;
    const int world = e.comm_world, rank = e.comm_rank;
    ncclComm_t comm = (ncclComm_t)e.comm;
    if (e.tables.count(shard_name)) throw Error(TG_ERR_INVALID_ARG, "table '" + shard_name + "' already exists");
    // a free slot: one whose shard table is gone (the same choice on every rank: the call sequences are the same)
    int si = -1;
    for (int i = 0; i < PUSH_SHUFFLE_SLOTS && si < 0; ++i)
        if (ps.slot[i].user.empty() || !e.tables.count(ps.slot[i].user)) si = i;
    if (si < 0) return -1;
    PushSlot& s = ps.slot[si];
    s.user.clear();
    std::vector<int64_t> counts((size_t)world, 0);
    int64_t nulls = 0;
    int launches = 0;
    // TG_SHUFFLE_TRACE=1: wall-clock milestones of one shuffle on stderr (every milestone follows a stream synchronisation)
    static const bool trace = getenv("TG_SHUFFLE_TRACE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!trace || rank != 0) return;
        cudaStreamSynchronize(e.stream);
        fprintf(stderr, "[shuffle %s] %-14s +%.3f ms\n", c.name.c_str(), what,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
    };
    const int w1 = world + 1;
    std::vector<long long> mine((size_t)w1), all((size_t)w1 * world);
    long long* d_mine = (long long*)e.d_comm_counts;
    long long* d_all = d_mine + w1;
    auto gather = [&](size_t n_words) {  // `mine` -> `all` on every rank
        TG_CUDA(cudaMemcpyAsync(d_mine, mine.data(), n_words * 8, cudaMemcpyHostToDevice, e.stream));
        TG_NCCL(nccl().AllGather(d_mine, d_all, n_words, NCCL_INT64, comm, e.stream));
        TG_CUDA(cudaMemcpyAsync(all.data(), d_all, n_words * world * 8, cudaMemcpyDeviceToHost, e.stream));
        TG_CUDA(cudaStreamSynchronize(e.stream));
    };
    // ---- dense Int64 keys (the 1 B-key permutation of C4, surrogate ids): split by VALUE RANGE, so that every rank gets a
    // contiguous slice of the key space and its de-duplication runs on the L2-resident bitmap path instead of hash tables.
    // Decided from the global min / max / count — the same numbers, hence the same decision, on every rank.
    bool by_range = false;
    long long range_min = 0;
    unsigned long long range_span = 1;
    if (allow_range && c.dtype == TG_INT64 && !getenv("TG_NO_RANGE_SHUFFLE")) {
        long long mn = 0, mx = 0;
        unsigned long long nv = 0;
        column_minmax_i64(e, c, t.n_rows, &mn, &mx, &nv, launches);
        mark("minmax");
        mine[0] = mn;
        mine[1] = mx;
        mine[2] = (long long)nv;
        gather(3);
        mark("gather minmax");
        long long gmn = INT64_MAX, gmx = INT64_MIN;
        unsigned long long gn = 0;
        for (int r = 0; r < world; ++r) {
            if (all[(size_t)r * 3 + 2] == 0) continue;
            gmn = std::min<long long>(gmn, all[(size_t)r * 3]);
            gmx = std::max<long long>(gmx, all[(size_t)r * 3 + 1]);
            gn += (unsigned long long)all[(size_t)r * 3 + 2];
        }
        if (gn > 0) {
            // the ends come from samples on large columns: pad them (1/64 of the width) so that practically every key lies
            // inside; the few that do not go to the first / last rank (still one rank per key value)
            const unsigned long long w0 = (unsigned long long)gmx - (unsigned long long)gmn, pad = w0 / 64 + 4096;
            if (gmn >= INT64_MIN + (long long)pad && gmx <= INT64_MAX - (long long)pad) {
                gmn -= (long long)pad;
                gmx += (long long)pad;
            }
            const unsigned long long range = (unsigned long long)gmx - (unsigned long long)gmn;
            const unsigned long long span = range / (unsigned long long)world + 1;
            if (range <= 32ull * gn + 8192ull && span < (1ull << 28)) {
                by_range = true;
                range_min = gmn;
                range_span = span;
            }
        }
    }
    // ---- counts (a value-range split that turns out badly balanced is abandoned for the hash split)
    for (int attempt = 0; attempt < 2; ++attempt) {
        push_partition_hist(e, c, t.n_rows, world, counts.data(), &nulls, launches, by_range ? &range_min : nullptr, range_span);
        mark("hist");
        for (int r = 0; r < world; ++r) mine[r] = counts[r];
        mine[world] = nulls;
        gather((size_t)w1);
        mark("gather counts");
        if (!by_range) break;
        long long total = 0, worst = 0;
        for (int d = 0; d < world; ++d) {
            long long tot = 0;
            for (int sr = 0; sr < world; ++sr) tot += all[(size_t)sr * w1 + d];
            total += tot;
            worst = std::max(worst, tot);
        }
        if (worst <= 2 * (total / world) + 4096) break;
        by_range = false;
    }
    int64_t nulls_total = 0, max_rows = 0, n_recv = 0;
    std::vector<unsigned long long> first((size_t)world, 0);  // where my part starts in rank d's buffer
    for (int sr = 0; sr < world; ++sr) nulls_total += all[(size_t)sr * w1 + world];
    for (int d = 0; d < world; ++d) {
        int64_t tot = 0;
        for (int sr = 0; sr < world; ++sr) {
            if (sr == rank) first[d] = (unsigned long long)tot;
            tot += all[(size_t)sr * w1 + d];
        }
        if (d == rank) n_recv = tot;
        max_rows = std::max(max_rows, tot + (d == 0 ? nulls_total : 0));
    }
    const int64_t my_nulls = rank == 0 ? nulls_total : 0;
    const int64_t n_rows = n_recv + my_nulls;
    // ---- capacity: the same decision on every rank (everyone holds the whole matrix)
    const size_t need = round_up((size_t)std::max<int64_t>(max_rows, 1) * 8 + 512, 256);
    if (need > s.cap) {
        try {
            push_slot_grow(e, s, round_up(need + need / 8, (size_t)1 << 20));
        } catch (Error&) {
            // peer mapping refused (no IPC / no P2P): NOT rank-consistent in general, but CUDA IPC availability is a
            // property of the node, so every rank lands here together; remember it and use send / recv from now on
            ps.disabled = true;
            cudaGetLastError();
            push_slot_release(e, s);
            return -2;  // the counts collective was consumed: tell the caller to rerun the fallback from the start
        }
    }
    mark("slot");
    // ---- the scatter IS the all-to-all
    push_partition_scatter(e, c, t.n_rows, world, first.data(), (uint64_t* const*)s.d_ptrs, launches, by_range ? &range_min : nullptr, range_span);
    mark("scatter");
    // tail padding + NULL rows of my own buffer (nobody else writes behind n_recv)
    TG_CUDA(cudaMemsetAsync(s.local + (size_t)n_recv * 8, 0, std::min(s.cap - (size_t)n_recv * 8, (size_t)my_nulls * 8 + 512), e.stream));
    // ---- barrier: every rank's scatter has completed (kernel completion makes its peer stores visible) before anyone reads
    TG_NCCL(nccl().AllGather(d_mine, d_all, 1, NCCL_INT64, comm, e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    mark("barrier");
    e.launches += launches;
    // ---- the shard table lives in the receive buffer
    auto tab = std::make_unique<Table>();
    tab->eng = &e;
    tab->name = shard_name;
    Column* sc = table_get_or_add(*tab, c.name, c.dtype);
    sc->n_rows = n_rows;
    tab->n_rows = n_rows;
    sc->values.p = s.local;
    sc->values.cap = 0;
    sc->values.owned = false;
    sc->adopted = true;
    sc->value_bytes = n_rows * 8;
    sc->null_count = my_nulls;
    if (my_nulls) {
        const size_t bits_b = round_up((size_t)(n_rows + 7) / 8 + 256, 256);
        sc->validity.p = e.dev_alloc(bits_b);
        sc->validity.cap = bits_b;
        sc->validity.owned = true;
        TG_CUDA(cudaMemsetAsync(sc->validity.p, 0, bits_b, e.stream));
        TG_CUDA(cudaMemsetAsync(sc->validity.p, 0xFF, (size_t)(n_recv / 8), e.stream));
        if (n_recv % 8) {
            const uint8_t last = (uint8_t)((1u << (n_recv % 8)) - 1u);
            TG_CUDA(cudaMemcpyAsync(sc->validity.p + n_recv / 8, &last, 1, cudaMemcpyHostToDevice, e.stream));
        }
        TG_CUDA(cudaStreamSynchronize(e.stream));
    }
    e.tables[shard_name] = std::move(tab);
    s.user = shard_name;
    int64_t sent = 0;
    for (int d = 0; d < world; ++d)
        if (d != rank) sent += counts[d];
    e.comm_bytes_sent += (uint64_t)sent * 8;
    return n_rows;
}

int64_t comm_shuffle_column(Engine& e, const std::string& table, const std::string& column, const std::string& shard_name, bool allow_range) {
    std::lock_guard<std::mutex> g(e.mu);
    TG_CUDA(cudaSetDevice(e.device));
    if (!e.comm) throw Error(TG_ERR_NCCL, "no communicator: call tg_comm_init first");
    e.sync_copies();
    auto it = e.tables.find(table);
    if (it == e.tables.end()) throw Error(TG_ERR_TABLE_NOT_FOUND, "table '" + table + "' not found");
    Table& t = *it->second;
    Column* c = t.find(column);
    if (!c) throw Error(TG_ERR_COLUMN_NOT_FOUND, "Schema error: No field named " + column + ". Valid fields are " + t.valid_fields() + ".");
    c = numeric_view(e, c);  // (Int32 / Float32 keys travel as their exactly widened values)
    if (!c) throw Error(TG_ERR_TYPE_MISMATCH, "the multi-GPU key shuffle supports numeric key columns (strings travel as fingerprints)");
    {
        const int64_t pushed = push_shuffle_column(e, t, *c, shard_name, allow_range);
        if (pushed >= 0) return pushed;
    }
    uint64_t* keys = nullptr;
    std::vector<int64_t> counts((size_t)e.comm_world, 0);
    int64_t nulls = 0;
    int launches = 0;
    partition_keys_by_rank(e, *c, t.n_rows, e.comm_world, &keys, counts.data(), &nulls, launches);
    e.launches += launches;
    return exchange_into_table(e, (const uint8_t*)keys, counts.data(), nulls, 8, shard_name, column, c->dtype);
}

// ---------------------------------------------------------------------------------------------------------------------
// The exchange of the distributed sort (Spearman's global ranks, ranks.cu) inside the library: evenly spaced sample keys of
// every rank -> the same world - 1 splitters everywhere -> the session's (key, payload) pairs pushed to the rank whose
// key range holds them by the partition's scatter kernel (NVLink peer stores, as in the key shuffle above). The receiving
// rank then sorts its range (rank_finish_x / _y). Returns false when peer mapping is unavailable: the host layer then
// runs the same exchange with tg_rank_sample / tg_rank_split and its own all-to-all.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int RANK_SAMPLES = 2048;

bool comm_rank_exchange(Engine& e, int64_t* n_recv_out, uint64_t* rank_base, uint64_t* total_out) {
    std::lock_guard<std::mutex> g(e.mu);
    TG_CUDA(cudaSetDevice(e.device));
    if (!e.comm) return false;
    static const bool off = getenv("TG_NO_PUSH_SHUFFLE") != nullptr;
    if (off) return false;
    if (!e.push) e.push = new PushState();
    PushState& ps = *(PushState*)e.push;
    if (ps.disabled) return false;
    const int world = e.comm_world, rank = e.comm_rank;
    if (world > 64) return false;
    ncclComm_t comm = (ncclComm_t)e.comm;
    const uint64_t* keys = nullptr;
    const void* payload = nullptr;
    int pay_bytes = 8;
    int64_t n = 0;
    rank_current_unlocked(e, &keys, &payload, &pay_bytes, &n);
    int launches = 0;
    static const bool trace = getenv("TG_SHUFFLE_TRACE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!trace || rank != 0) return;
        cudaStreamSynchronize(e.stream);
        fprintf(stderr, "[rank exchange %d] %-14s +%.3f ms\n", pay_bytes, what,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
    };
    // ---- samples -> splitters (identical on every rank)
    const size_t sw = RANK_SAMPLES + 1;
    if (!e.d_comm_samples) TG_CUDA(cudaMalloc(&e.d_comm_samples, sw * 8 * (size_t)(world + 1) + 1024));
    std::vector<uint64_t> mine(sw, 0), all(sw * (size_t)world);
    mine[0] = (uint64_t)rank_sample_unlocked(e, RANK_SAMPLES, mine.data() + 1);
    uint64_t* d_mine = (uint64_t*)e.d_comm_samples;
    uint64_t* d_all = d_mine + sw;
    TG_CUDA(cudaMemcpyAsync(d_mine, mine.data(), sw * 8, cudaMemcpyHostToDevice, e.stream));
    TG_NCCL(nccl().AllGather(d_mine, d_all, sw, NCCL_INT64, comm, e.stream));
    TG_CUDA(cudaMemcpyAsync(all.data(), d_all, sw * world * 8, cudaMemcpyDeviceToHost, e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    std::vector<uint64_t> samples;
    for (int r = 0; r < world; ++r) {
        const uint64_t c = std::min<uint64_t>(all[(size_t)r * sw], RANK_SAMPLES);
        samples.insert(samples.end(), all.begin() + (size_t)r * sw + 1, all.begin() + (size_t)r * sw + 1 + c);
    }
    std::sort(samples.begin(), samples.end());
    std::vector<uint64_t> splitters((size_t)std::max(world - 1, 1), 0);
    for (int p = 0; p + 1 < world && !samples.empty(); ++p) {
        const size_t L = samples.size();
        const size_t idx = std::min(L - 1, (size_t)std::max<long long>(0, (long long)((size_t)(p + 1) * L / (size_t)world) - 1));
        splitters[p] = samples[idx];
    }
    uint64_t* d_split = d_all + sw * world;  // behind the gathered samples (1 KB of slack was allocated: world <= 64 splitters)
    TG_CUDA(cudaMemcpyAsync(d_split, splitters.data(), splitters.size() * 8, cudaMemcpyHostToDevice, e.stream));
    mark("splitters");
    // ---- counts
    std::vector<int64_t> counts((size_t)world, 0);
    split_partition_hist(e, keys, n, world, d_split, counts.data(), launches);
    mark("hist");
    std::vector<long long> cm((size_t)world), call((size_t)world * world);
    for (int r = 0; r < world; ++r) cm[r] = counts[r];
    long long* dc_mine = (long long*)e.d_comm_counts;
    long long* dc_all = dc_mine + (world + 1);
    TG_CUDA(cudaMemcpyAsync(dc_mine, cm.data(), (size_t)world * 8, cudaMemcpyHostToDevice, e.stream));
    TG_NCCL(nccl().AllGather(dc_mine, dc_all, (size_t)world, NCCL_INT64, comm, e.stream));
    TG_CUDA(cudaMemcpyAsync(call.data(), dc_all, (size_t)world * world * 8, cudaMemcpyDeviceToHost, e.stream));
    TG_CUDA(cudaStreamSynchronize(e.stream));
    std::vector<unsigned long long> first((size_t)world, 0);
    int64_t n_recv = 0, max_rows = 0;
    uint64_t base = 0, total = 0;
    for (int d = 0; d < world; ++d) {
        int64_t tot = 0;
        for (int sr = 0; sr < world; ++sr) {
            if (sr == rank) first[d] = (unsigned long long)tot;
            tot += call[(size_t)sr * world + d];
        }
        if (d == rank) n_recv = tot;
        if (d < rank) base += (uint64_t)tot;
        total += (uint64_t)tot;
        max_rows = std::max(max_rows, tot);
    }
    if (max_rows >= ((int64_t)1 << 30)) throw Error(TG_ERR_UNSUPPORTED, "Spearman: 2^30 or more keys in one rank's range");
    // ---- receive buffers of this phase (keys, payload), grown collectively
    const int phase = rank_phase_unlocked(e);
    PushSlot& sk = ps.slot[PUSH_SHUFFLE_SLOTS + 2 * phase];
    PushSlot& sp = ps.slot[PUSH_SHUFFLE_SLOTS + 2 * phase + 1];
    const size_t need = round_up((size_t)std::max<int64_t>(max_rows, 1) * 8 + 512, 256);
    try {
        if (need > sk.cap) push_slot_grow(e, sk, round_up(need + need / 8, (size_t)1 << 20));
        if (need > sp.cap) push_slot_grow(e, sp, round_up(need + need / 8, (size_t)1 << 20));
    } catch (Error&) {
        ps.disabled = true;
        cudaGetLastError();
        push_slot_release(e, sk);
        push_slot_release(e, sp);
        return false;  // every rank of the node fails alike (IPC is a property of the node) and takes the host-layer path
    }
    mark("counts+slots");
    // ---- the scatter is the exchange
    split_partition_scatter(e, keys, payload, pay_bytes, n, world, d_split, first.data(), (uint64_t* const*)sk.d_ptrs, (uint8_t* const*)sp.d_ptrs, launches);
    TG_NCCL(nccl().AllGather(dc_mine, dc_all, 1, NCCL_INT64, comm, e.stream));  // barrier: every rank's stores have landed
    TG_CUDA(cudaStreamSynchronize(e.stream));
    mark("scatter+barrier");
    e.launches += launches;
    e.comm_bytes_sent += (uint64_t)(n - counts[rank]) * (uint64_t)(8 + pay_bytes);
    rank_adopt_received_unlocked(e, (uint64_t*)sk.local, sp.local, n_recv);
    mark("adopt");
    *n_recv_out = n_recv;
    *rank_base = base;
    *total_out = total;
    return true;
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
