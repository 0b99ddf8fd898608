// Peer mailboxes: the partial-state exchange of a multi-GPU step over NVLink / NVSwitch peer memory.
//
// Every rank owns a small device buffer [2 parities][world slots][slot bytes] + flags, exported to the other ranks of
// the node through CUDA IPC. After its partial execute a rank PUBLISHES: one kernel stores its serialised partial
// states into slot[rank] of every peer's mailbox (plain stores that travel over NVLink), fences system-wide and then
// writes the step's sequence number into the peer's flag[rank]. COLLECT is a second tiny kernel (one block per rank)
// that spins on the local flag of its rank and then copies that rank's payload — only the bytes it holds — straight
// into pinned host memory. Compared with H2D -> ncclAllGather -> D2H this removes two copies and the collective's launch and
// rendezvous latency from every step (the payload is a few KB). Double buffering by step parity is enough: a rank
// can only be one step ahead of the slowest rank, because it cannot finish collecting step s+1 before everyone has
// published s+1, which they do after collecting s.
#include <cstring>

#include "engine.hpp"

namespace tg {

struct MailboxHeader {                    // device layout: header, then data
    unsigned long long flag[2][MAILBOX_MAX_WORLD];
};

__global__ void mailbox_publish_kernel(const uint4* __restrict__ src, uint32_t n16, MailboxPeers peers, int world, int rank,
                                       uint32_t slot_bytes, int parity, unsigned long long seq) {
    // block b serves peer b (including this rank's own mailbox)
    const int peer = blockIdx.x;
    if (peer >= world) return;
    uint8_t* base = peers.p[peer];
    uint4* dst = reinterpret_cast<uint4*>(base + sizeof(MailboxHeader) + ((size_t)parity * world + rank) * slot_bytes);
    for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned long long* f = &reinterpret_cast<MailboxHeader*>(base)->flag[parity][rank];
        *f = seq;
        __threadfence_system();
    }
}

// COLLECT, one block per rank: wait for rank r's sequence number, then copy its payload (8-byte length + bytes, only as
// much as it holds) from the local mailbox straight into pinned host memory — no device-to-host copy call, and a
// large slot costs nothing when the payload is small. A length of ~0 marks "my payload did not fit the slot".
__global__ void mailbox_collect_kernel(const uint8_t* base, int world, uint32_t slot_bytes, int parity, unsigned long long seq,
                                       uint8_t* host_out /* pinned, [world][slot_bytes] + 64 */, unsigned long long* timeout_flag /* pinned */) {
    const int r = blockIdx.x;
    if (r >= world) return;
    __shared__ unsigned long long s_len;
    if (threadIdx.x == 0) {
        const volatile unsigned long long* f = &reinterpret_cast<const MailboxHeader*>(base)->flag[parity][r];
        unsigned long long spins = 0;
        bool ok = true;
        while (*f != seq) {
            if (++spins > (1ull << 31)) {  // ~ seconds: a peer died; report instead of hanging the GPU
                *timeout_flag = 1;
                ok = false;
                break;
            }
        }
        __threadfence_system();
        const uint8_t* slot = base + sizeof(MailboxHeader) + ((size_t)parity * world + r) * slot_bytes;
        unsigned long long len = ok ? *reinterpret_cast<const volatile unsigned long long*>(slot) : 0ull;
        if (len != ~0ull && len + 8 > slot_bytes) len = 0;
        s_len = len;
    }
    __syncthreads();
    const unsigned long long len = s_len;
    const uint8_t* slot = base + sizeof(MailboxHeader) + ((size_t)parity * world + r) * slot_bytes;
    uint4* dst = reinterpret_cast<uint4*>(host_out + (size_t)r * slot_bytes);
    const uint32_t n16 = len == ~0ull ? 1u : (uint32_t)((len + 8 + 15) / 16);
    const uint4* src = reinterpret_cast<const uint4*>(slot);
    for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}

void mailbox_create(Engine& e, int world, int rank, size_t slot_bytes, void* handle_out) {
    if (world < 1 || world > MAILBOX_MAX_WORLD || rank < 0 || rank >= world) throw Error(TG_ERR_INVALID_ARG, "mailbox: bad world / rank");
    slot_bytes = (slot_bytes + 255) / 256 * 256;
    mailbox_destroy(e);
    Mailbox& m = e.mailbox;
    m.world = world;
    m.rank = rank;
    m.slot_bytes = slot_bytes;
    m.bytes = sizeof(MailboxHeader) + 2 * (size_t)world * slot_bytes + 256;
    TG_CUDA(cudaMalloc(&m.local, m.bytes));
    TG_CUDA(cudaMemset(m.local, 0, m.bytes));
    TG_CUDA(cudaMalloc(&m.d_stage, slot_bytes + 256));
    TG_CUDA(cudaMallocHost(&m.h_stage, slot_bytes));
    TG_CUDA(cudaMallocHost(&m.h_all, (size_t)world * slot_bytes + 64));  // the collect kernel stores into it directly (UVA); last 64 bytes: timeout flag
    TG_CUDA(cudaMemset(m.d_stage, 0, slot_bytes + 256));
    cudaIpcMemHandle_t h;
    TG_CUDA(cudaIpcGetMemHandle(&h, m.local));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle_out, &h, sizeof(h));
    for (auto& p : m.peers.p) p = nullptr;
    m.peers.p[rank] = m.local;
    m.seq = 0;
    m.open = false;
}

void mailbox_open(Engine& e, const void* handles) {
    Mailbox& m = e.mailbox;
    if (!m.local) throw Error(TG_ERR_INVALID_ARG, "mailbox: create it first");
    for (int r = 0; r < m.world; ++r) {
        if (r == m.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t*)handles + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        TG_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        m.peers.p[r] = (uint8_t*)p;
    }
    m.open = true;
}

void mailbox_destroy(Engine& e) {
    Mailbox& m = e.mailbox;
    for (int r = 0; r < m.world; ++r)
        if (r != m.rank && m.peers.p[r]) cudaIpcCloseMemHandle(m.peers.p[r]);
    if (m.local) cudaFree(m.local);
    if (m.d_stage) cudaFree(m.d_stage);
    if (m.h_stage) cudaFreeHost(m.h_stage);
    if (m.h_all) cudaFreeHost(m.h_all);
    m = Mailbox{};
}

// queue PUBLISH (d_stage -> slot[rank] of every peer + flag) and COLLECT (every rank's payload -> m.h_all) on the engine's
// stream; `send` = bytes of d_stage to publish (16-byte multiple; the payload starts with its 8-byte length)
static void mailbox_queue(Engine& e, size_t send, unsigned long long seq) {
    Mailbox& m = e.mailbox;
    const int parity = (int)(seq & 1ull);
    mailbox_publish_kernel<<<m.world, 128, 0, e.stream>>>((const uint4*)m.d_stage, (uint32_t)(send / 16), m.peers, m.world, m.rank,
                                                          (uint32_t)m.slot_bytes, parity, seq);
    unsigned long long* h_timeout = (unsigned long long*)(m.h_all + (size_t)m.world * m.slot_bytes);
    *h_timeout = 0;
    mailbox_collect_kernel<<<m.world, 128, 0, e.stream>>>(m.local, m.world, (uint32_t)m.slot_bytes, parity, seq, m.h_all, h_timeout);
    TG_CUDA(cudaGetLastError());
    e.launches += 2;
}
// after the stream was synchronised: rank r's payload; false when some rank could not fit its payload (then nobody uses
// this exchange: every rank sees the same markers and takes the fallback)
static bool mailbox_read(Engine& e, std::vector<std::vector<uint8_t>>& out) {
    Mailbox& m = e.mailbox;
    const unsigned long long timeout = *(const volatile unsigned long long*)(m.h_all + (size_t)m.world * m.slot_bytes);
    if (timeout) throw Error(TG_ERR_NCCL, "mailbox: a peer did not publish its partial state in time");
    out.resize(m.world);
    bool all_fit = true;
    for (int r = 0; r < m.world; ++r) {
        const uint8_t* s = m.h_all + (size_t)r * m.slot_bytes;
        uint64_t l;
        memcpy(&l, s, 8);
        if (l == ~0ull) {
            all_fit = false;
            continue;
        }
        if (l + 8 > m.slot_bytes) throw Error(TG_ERR_INTERNAL, "mailbox: corrupt slot");
        out[r].assign(s + 8, s + 8 + l);
    }
    return all_fit;
}

// publish `n` bytes (this rank's partial blob, prefixed with its length) and collect every rank's; out[r] = rank r's blob.
// Returns false when some rank's blob did not fit its slot: every rank gets false and the host layer falls back to NCCL.
bool mailbox_exchange(Engine& e, const uint8_t* blob, size_t n, std::vector<std::vector<uint8_t>>& out) {
    Mailbox& m = e.mailbox;
    if (!m.open) throw Error(TG_ERR_INVALID_ARG, "mailbox: not open");
    std::lock_guard<std::mutex> g(e.mu);
    TG_CUDA(cudaSetDevice(e.device));
    const unsigned long long seq = ++m.seq;
    const bool fits = n + 8 <= m.slot_bytes;
    const uint64_t len = fits ? (uint64_t)n : ~0ull;
    memcpy(m.h_stage, &len, 8);
    if (fits) memcpy(m.h_stage + 8, blob, n);
    const size_t send = fits ? (n + 8 + 15) / 16 * 16 : 16;
    TG_CUDA(cudaMemcpyAsync(m.d_stage, m.h_stage, send, cudaMemcpyHostToDevice, e.stream));
    mailbox_queue(e, send, seq);
    TG_CUDA(cudaStreamSynchronize(e.stream));
    return mailbox_read(e, out);
}

// the fused path of a scan-only plan (engine.cu): the payload was assembled in d_stage by kernels already queued on the
// stream — publish / collect behind them, ONE synchronisation for the whole step
void mailbox_exchange_device(Engine& e, size_t payload_bytes, std::vector<std::vector<uint8_t>>& out) {
    Mailbox& m = e.mailbox;
    if (!m.open) throw Error(TG_ERR_INVALID_ARG, "mailbox: not open");
    if (payload_bytes > m.slot_bytes) throw Error(TG_ERR_INVALID_ARG, "mailbox: payload larger than the slot");
    const unsigned long long seq = ++m.seq;
    mailbox_queue(e, (payload_bytes + 15) / 16 * 16, seq);
    TG_CUDA(cudaStreamSynchronize(e.stream));
    if (!mailbox_read(e, out)) throw Error(TG_ERR_INTERNAL, "mailbox: a rank did not take the fused path");
}

}  // namespace tg
