// C-ABI engine (include/hsk_capi.h): context, buffers, stage orchestration, supermer exchange.
//
// Replaces the body of the reference's hysortk::kmer_count (src/hysortk.cpp:36-95):
//   prepare_supermer  (kmerops.cpp:23-126)   -> reads.cu (read table), extract.cu (count pass, bin scan, scatter pass)
//   exchange_supermer (kmerops.cpp:130-195)  -> all-gather of the bin totals + CUDA IPC: the bin kernel reads the peers'
//                                               supermer streams in place over NVLink (TaskManager::exchange 80 KB
//                                               rounds, :814-1007); HSK_EXCHANGE=nccl: grouped ncclSend/ncclRecv
//   filter_kmer       (kmerops.cpp:198-250)  -> bins.cu: one persistent kernel expands, counts, sorts and emits every
//                                               bin; skewed bins go through expand.cu, radix.cu, count.cu (HBM path)
// The reference's task system (TaskManager, classifier, dispatcher) has no equivalent here: bins are owned by rank in
// contiguous ranges.  hsk_count / hsk_count_stream additionally pipeline the host side around the kernels: the input goes
// up in chunks under the extraction count pass (pageable memory through a ring of page-locked slots filled by the
// context's host threads), the result leaves in groups of bins while the bin kernel runs and is handed to the caller's
// sink by the context's delivery threads.  Memory: arena and staging area are limited by what is free; when that is short
// the run list and the arena share one block (count_device).
#include "../../include/hsk_capi.h"
#include "kernels.cuh"

#include <nccl.h>
#include <dlfcn.h>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define HSK_HAVE_SSE2 1
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <deque>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <unistd.h>
#include <string>
#include <thread>
#include <vector>

using namespace hsk;

static thread_local std::string g_err;

static int fail(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) return fail("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
// NCCL is bound at run time (dlopen) and only when a context spans more than one rank: a single-GPU
// caller needs no NCCL at all, and inside a process that already loaded an NCCL (e.g. PyTorch's
// bundled copy) the same library instance is reused instead of a second, older one being pulled in.
namespace {
struct NcclApi {
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
}

static int load_nccl()
{
    if (g_nccl.ok) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail("cannot load libnccl.so.2: %s", dlerror());
#define HSK_SYM(field, name)                                                            \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, #name));           \
    if (!g_nccl.field) return fail("libnccl: missing symbol %s", #name)
    HSK_SYM(GetUniqueId, ncclGetUniqueId);
    HSK_SYM(CommInitRank, ncclCommInitRank);
    HSK_SYM(CommDestroy, ncclCommDestroy);
    HSK_SYM(AllGather, ncclAllGather);
    HSK_SYM(AllReduce, ncclAllReduce);
    HSK_SYM(Send, ncclSend);
    HSK_SYM(Recv, ncclRecv);
    HSK_SYM(GroupStart, ncclGroupStart);
    HSK_SYM(GroupEnd, ncclGroupEnd);
    HSK_SYM(GetErrorString, ncclGetErrorString);
#undef HSK_SYM
    g_nccl.ok = true;
    return 0;
}

#define NK(call)                                                                                        \
    do {                                                                                                \
        ncclResult_t r_ = (call);                                                                       \
        if (r_ != ncclSuccess) return fail("%s:%d: %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
    } while (0)

namespace {

// HSK_TRACE=1: host-side timeline of a call on stderr (milliseconds since the first mark)
struct Trace {
    bool on = false;
    std::chrono::steady_clock::time_point t0;
    void start() { const char *e = getenv("HSK_TRACE"); on = e && *e == '1'; if (on) t0 = std::chrono::steady_clock::now(); }
    void mark(const char *what) const
    {
        if (!on) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[hsk trace] %8.3f ms  %s\n", ms, what);
    }
};
Trace g_trace;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool view = false;   // p points into another allocation (set_view): never freed, never grown
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (view) return cudaErrorMemoryAllocation;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 16 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    // exactly `bytes` (no slack): for the one block that takes most of the free memory
    cudaError_t ensure_exact(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void set_view(void *q, size_t bytes) { if (p && !view) cudaFree(p); p = q; cap = bytes; view = true; }
    void release() { if (p && !view) cudaFree(p); p = nullptr; cap = 0; view = false; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct HostBuf {   // page-locked
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 16 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    // grow, keeping the first `used` bytes (the caller makes sure no copy into the old block is in flight)
    cudaError_t ensure_keep(size_t bytes, size_t used)
    {
        if (bytes <= cap) return cudaSuccess;
        void *q = nullptr;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaHostAlloc(&q, want, cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        if (p) { if (used) memcpy(q, p, used); cudaFreeHost(p); }
        p = q; cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct EvPair { cudaEvent_t a, b; };

// Copy into a page-locked staging buffer that the GPU reads next.  Non-temporal stores send the lines to memory instead
// of leaving them dirty in the caches of the copying cores: a DMA read of a buffer just written by 16 threads with plain
// stores runs at 22 GB/s on this platform, after streaming stores at the full 52 GB/s (tools/hostbench/h2d_after_write.cu).
static void stream_copy(void *dst_, const void *src_, size_t n)
{
#ifdef HSK_HAVE_SSE2
    unsigned char *dst = static_cast<unsigned char *>(dst_);
    const unsigned char *src = static_cast<const unsigned char *>(src_);
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 15)) { dst[i] = src[i]; ++i; }
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 32));
        const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 48), d);
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
#else
    memcpy(dst_, src_, n);
#endif
}

// A few host threads of the context: copy a pageable DnaBuffer into the page-locked staging ring, fill caller memory from
// the result arrays.  parallel_for splits [0, n) into one contiguous piece per thread (the caller takes the first).
class HostPool {
public:
    explicit HostPool(unsigned nthreads) : n_(std::max(1u, nthreads))
    {
        for (unsigned t = 0; t < n_; ++t) workers_.emplace_back([this] { run(); });
    }
    ~HostPool()
    {
        wait();
        { std::lock_guard<std::mutex> l(m_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
    }
    // one job at a time: items 0 .. nitems-1 are handed to the threads in order; returns at once
    void start(u64 nitems, std::function<void(u64)> fn)
    {
        wait();
        if (nitems == 0) return;
        {
            std::lock_guard<std::mutex> l(m_);
            fn_ = std::move(fn); total_ = nitems; next_.store(0); active_ = n_; running_ = true; ++gen_;
        }
        cv_.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return !running_; });
    }
    // [0, n) in contiguous pieces, a few per thread; returns when all are done
    void parallel_for(u64 n, const std::function<void(u64, u64)> &fn)
    {
        if (n == 0) return;
        if (n_ == 1 || n < 2) { fn(0, n); return; }
        const u64 pieces = std::min<u64>(n, (u64)n_ * 2);
        start(pieces, [&fn, n, pieces](u64 i) { fn(n * i / pieces, n * (i + 1) / pieces); });
        wait();
    }
private:
    void run()
    {
        u64 seen = 0;
        while (true) {
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            for (u64 i = next_.fetch_add(1); i < total_; i = next_.fetch_add(1)) fn_(i);
            { std::lock_guard<std::mutex> l(m_); if (--active_ == 0) { running_ = false; done_.notify_all(); } }
        }
    }
    unsigned n_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::function<void(u64)> fn_;
    std::atomic<u64> next_{0};
    u64 total_ = 0, gen_ = 0;
    unsigned active_ = 0;
    bool running_ = false, stop_ = false;
};

// hsk_count_stream: parts of the result that have reached the page-locked host arrays are handed to the caller's sink by
// a few threads of the context (one part per thread at a time) while the GPU works on the rest.
// occurrences of a part: [occ0, occ1); ~0 = read it from the occurrence offsets of the entries once they have arrived
struct SinkPart { u64 first, n, occ0, occ1, hint; cudaEvent_t ready; };
constexpr u64 SINK_SUBPART = 1ull << 16;   // entries per delivery: small enough for the last ones to be short
class SinkPool {
public:
    explicit SinkPool(unsigned nthreads)
    {
        for (unsigned t = 0; t < std::max(1u, nthreads); ++t) th_.emplace_back([this] { run(); });
    }
    ~SinkPool()
    {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    void begin(hsk_sink_fn fn, void *user, const hsk_result *view, int device)
    {
        std::lock_guard<std::mutex> l(m_);
        fn_ = fn; user_ = user; view_ = view; device_ = device; rc_ = 0; busy_ = 0;
    }
    void push(const SinkPart &p)
    {
        { std::lock_guard<std::mutex> l(m_); q_.push_back(p); ++busy_; }
        cv_.notify_one();
    }
    // waits until every pushed part has been delivered; returns the first non-zero sink status
    int drain()
    {
        std::unique_lock<std::mutex> l(m_);
        idle_.wait(l, [this] { return busy_ == 0; });
        return rc_;
    }
private:
    void run()
    {
        int dev_set = -1;
        while (true) {
            SinkPart p;
            hsk_sink_fn fn; void *user; const hsk_result *view; int device; bool skip;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [this] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                p = q_.front(); q_.pop_front();
                fn = fn_; user = user_; view = view_; device = device_; skip = rc_ != 0;
            }
            if (dev_set != device) { cudaSetDevice(device); dev_set = device; }
            int rc = 0;
            if (p.ready && cudaEventSynchronize(p.ready) != cudaSuccess) rc = 2;
            if (!rc && !skip && fn) {
                const u64 o0 = p.occ0 != ~0ull ? p.occ0 : view->occ_off[p.first];
                const u64 o1 = p.occ1 != ~0ull ? p.occ1 : view->occ_off[p.first + p.n];
                rc = fn(user, view, p.first, p.n, o0, o1 - o0, p.hint);
            }
            {
                std::lock_guard<std::mutex> l(m_);
                if (rc && !rc_) rc_ = rc;
                if (--busy_ == 0) idle_.notify_all();
            }
        }
    }
    std::mutex m_;
    std::condition_variable cv_, idle_;
    std::deque<SinkPart> q_;
    hsk_sink_fn fn_ = nullptr;
    void *user_ = nullptr;
    const hsk_result *view_ = nullptr;
    int device_ = 0, rc_ = 0;
    unsigned busy_ = 0;
    bool stop_ = false;
    std::vector<std::thread> th_;
};

} // namespace

struct hsk_ctx {
    hsk_config cfg;
    int nwords = 1, m_eff = 0;
    u32 tg = 0, tt = 0;   // buckets per rank / total
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // host <-> device copies of hsk_count, overlapped with the kernels
    bool own_stream = false;
    ncclComm_t comm = nullptr;
    int sm_count = 148;

    // input staging (hsk_count)
    DevBuf d_packed, d_read_off, d_read_len, d_len64, d_rtscratch;
    // host pipeline of hsk_count: chunks of the packed reads in flight (tile bound + event), result streaming
    struct InChunk { u64 tile_end; cudaEvent_t ready; u64 off, n; u32 pieces; };
    std::vector<InChunk> in_chunks;
    bool stream_result = false;          // hsk_count: results go to the host buffers group by group
    u32 *d_in_flags = nullptr;           // hsk_count: read table checks (reads.cu), looked at after the first sync
    DevBuf d_dd;                         // per-CTA lists of distinct supermers (bins.cu: dedup_bin)
    DevBuf d_pend;                       // per-CTA stash of the last sorted bin (bins.cu: stash_bin)
    DevBuf d_grp;                        // per bin group: ticket, big-bin counter
    HostBuf h_grp;                       // per bin group: arena cursor after the group
    // extraction
    DevBuf d_bucket, d_run_list, d_tile_hdr, d_tile_read, d_bscratch;   // d_bucket: see run_extract
    HostBuf h_bucket;
    DevBuf d_slots;
    // exchange
    DevBuf d_alltot, d_rslots, d_seg, d_lb, d_peerinfo;
    HostBuf h_alltot, h_meta, h_peerinfo;
    std::vector<u64> rbase_idx;
    // peer supermer buffers mapped into this process (fused scatter + exchange over NVLink)
    struct PeerMap { unsigned char info[128]; void *mapped = nullptr; bool ipc = false; bool valid = false; };
    PeerMap peers[BN_MAX_SRC];
    bool use_p2p = true;
    std::vector<DevBuf> retired;         // former supermer buffers that peers may still have mapped (freed one call later)
    // host side of hsk_count_stream
    std::unique_ptr<HostPool> pool;
    std::unique_ptr<SinkPool> sink;
    unsigned host_threads = 1;
    hsk_sink_fn sink_fn = nullptr;
    void *sink_user = nullptr;
    hsk_result sink_view;
    // pageable input: pieces of the buffer are copied into a ring of page-locked buffers by the pool's threads, each of
    // which then sends its piece itself (stage_input / stage_chunk)
    static constexpr u64 PIECE = 1ull << 20;
    HostBuf h_ring, h_len;
    std::vector<cudaEvent_t> ring_free;              // per ring slot: its last piece has left
    std::unique_ptr<std::atomic<u32>[]> chunk_sent;  // per extraction chunk: its H2D copy + event are enqueued
    std::unique_ptr<std::atomic<u32>[]> chunk_done;  // per extraction chunk: pieces copied into its ring slot so far
    size_t flags_cap = 0;
    const u8 *in_host = nullptr;
    bool in_pageable = false;
    std::atomic<int> stage_err{0};
    std::atomic<long long> stage_ns[3];              // HSK_TRACE: time of the staging threads in memcpy / CUDA calls / waiting
    // extraction state between the count pass and the scatter pass
    ExtractParams xp;
    u32 x_nctas = 0;
    const u32 *src_slots[BN_MAX_SRC] = {};   // per source: first slot of the stream the bins read
    // batch buffers
    DevBuf d_keys[2][MAX_WORDS], d_val[2], d_rscratch, d_cscratch, d_tsum, d_tbase;
    // result arena
    DevBuf d_owords, d_ocnt, d_oocc_off, d_opos, d_orid, d_hist, d_cursor;
    DevBuf d_swords, d_scnt, d_spos, d_srid;   // staging (bins in completion order)
    HostBuf h_cursor, h_owords, h_ocnt, h_oocc_off, h_opos, h_orid, h_hist;
    u64 n_kept = 0, n_occ = 0;
    bool have_result = false;
    // memory-short mode: ONE block holds the run list during the extraction and the arena + staging area afterwards
    DevBuf d_big;
    bool big_mode = false;
    u64 big_budget = 0;                  // HSK_ARENA_BUDGET_MB (tests) that sized the block, 0: the free memory did
    u64 *run_list = nullptr;             // the run list of this call: d_run_list, or the start of d_big
    hsk_stats stats;

    std::vector<u64> h_dbg;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::vector<EvPair> ev_extract, ev_exchange, ev_expand, ev_sort, ev_count, ev_pass, ev_bins;

    cudaEvent_t ev()
    {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ev_pool.push_back(e);
        }
        return ev_pool[ev_used++];
    }
    EvPair begin(std::vector<EvPair> &v)
    {
        EvPair p{ev(), ev()};
        cudaEventRecord(p.a, stream);
        v.push_back(p);
        return p;
    }
    void end(std::vector<EvPair> &v) { cudaEventRecord(v.back().b, stream); }
    static float sum_ms(const std::vector<EvPair> &v)
    {
        float t = 0;
        for (auto &p : v) { float ms = 0; cudaEventElapsedTime(&ms, p.a, p.b); t += ms; }
        return t;
    }
};

extern "C" {

const char *hsk_last_error(void) { return g_err.c_str(); }
int hsk_version(void) { return HSK_VERSION; }

int hsk_get_unique_id(void *id_out)
{
    static_assert(sizeof(ncclUniqueId) <= HSK_NCCL_ID_BYTES, "nccl id size");
    ncclUniqueId id;
    if (load_nccl()) return 1;
    NK(g_nccl.GetUniqueId(&id));
    memset(id_out, 0, HSK_NCCL_ID_BYTES);
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

int hsk_create(hsk_ctx **out, const hsk_config *cfg)
{
    if (!out || !cfg) return fail("hsk_create: null argument");
    if (!(cfg->k > 2 && cfg->k < 96)) return fail("hsk_create: KMER_SIZE must satisfy 2 < k < 96 (got %d)", cfg->k);
    if (!(cfg->m > 0 && cfg->m < cfg->k)) return fail("hsk_create: MINIMIZER_SIZE must satisfy 0 < m < k (got %d)", cfg->m);
    if (!(cfg->lower > 0 && cfg->lower <= cfg->upper && cfg->upper <= 65535))
        return fail("hsk_create: need 0 < LOWER <= UPPER <= 65535 (got %d, %d)", cfg->lower, cfg->upper);
    if (cfg->nranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->nranks) return fail("hsk_create: bad rank/nranks");
    if (cfg->nranks > 1 && !cfg->nccl_id) return fail("hsk_create: nccl_id required when nranks > 1");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail("hsk_create: device %d not present (%d devices)", cfg->device, ndev);
    CK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) return fail("hsk_create: device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);

    hsk_ctx *c = new hsk_ctx;
    c->cfg = *cfg;
    c->cfg.nccl_id = nullptr;
    c->nwords = nwords_for_k(cfg->k);
    c->m_eff = std::min(cfg->m, 32);
    if (cfg->k - c->m_eff + 1 > XT_WMAX) c->m_eff = cfg->k + 1 - XT_WMAX;   // window of at most XT_WMAX m-mers
    c->sm_count = prop.multiProcessorCount;
    if (cfg->buckets_per_rank > 0 && (u64)cfg->buckets_per_rank * cfg->nranks > MAX_BINS) {
        delete c;
        return fail("hsk_create: buckets_per_rank * nranks exceeds %u", MAX_BINS);
    }
    c->tg = cfg->buckets_per_rank > 0 ? (u32)cfg->buckets_per_rank : 0;   // 0: chosen per call from the input size
    c->tt = c->tg * (u32)cfg->nranks;
    if (cfg->stream) { c->stream = (cudaStream_t)cfg->stream; c->own_stream = false; }
    else {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return fail("cudaStreamCreate: %s", cudaGetErrorString(e)); }
        c->own_stream = true;
    }
    {
        cudaError_t e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return fail("cudaStreamCreate: %s", cudaGetErrorString(e)); }
    }
    if (cfg->nranks > 1) {
        ncclUniqueId id;
        memcpy(&id, cfg->nccl_id, sizeof(id));
        if (load_nccl()) { delete c; return 1; }
        ncclResult_t r = g_nccl.CommInitRank(&c->comm, cfg->nranks, id, cfg->rank);
        if (r != ncclSuccess) { delete c; return fail("ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
    }
    memset(&c->stats, 0, sizeof(c->stats));
    memset(&c->sink_view, 0, sizeof(c->sink_view));
    if (const char *ev = getenv("HSK_EXCHANGE")) c->use_p2p = strcmp(ev, "nccl") != 0;
    {
        // host threads of the context (staging of pageable input, hsk_fill_entries): the ranks of a node share its cores
        unsigned nt = std::max(1u, std::thread::hardware_concurrency() / (unsigned)cfg->nranks);
        nt = std::min(nt, 16u);
        if (const char *ev = getenv("HSK_HOST_THREADS")) { const int v = atoi(ev); if (v >= 1 && v <= 64) nt = (unsigned)v; }
        c->host_threads = nt;
        c->pool.reset(new HostPool(nt));
    }
    *out = c;
    return 0;
}

void hsk_destroy(hsk_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    cudaStreamSynchronize(c->stream);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    c->sink.reset();
    c->pool.reset();
    // Collective when the context spans several ranks: nobody may still be reading a peer's supermers when they are
    // freed, and an exported buffer is only freed after every peer has closed its mapping of it.
    const bool fence = c->comm && c->use_p2p && c->d_cursor.p && !getenv("HSK_NO_DESTROY_BARRIER");
    auto barrier = [&]() {
        if (!fence) return;
        u64 *w = c->d_cursor.as<u64>() + 8;
        if (g_nccl.AllReduce(w, w + 1, 1, ncclUint64, ncclSum, c->comm, c->stream) == ncclSuccess) cudaStreamSynchronize(c->stream);
    };
    barrier();
    for (int p = 0; p < BN_MAX_SRC; ++p) {
        if (c->peers[p].valid && c->peers[p].ipc && c->peers[p].mapped) cudaIpcCloseMemHandle(c->peers[p].mapped);
        c->peers[p].valid = false;
    }
    barrier();
    if (c->comm) g_nccl.CommDestroy(c->comm);
    for (auto &r : c->retired) r.release();
    c->d_big.release();
    c->h_ring.release();
    for (auto e : c->ring_free) cudaEventDestroy(e);
    c->h_len.release();
    DevBuf *db[] = {&c->d_packed, &c->d_read_off, &c->d_read_len, &c->d_len64, &c->d_rtscratch, &c->d_run_list, &c->d_tile_hdr, &c->d_tile_read, &c->d_bscratch, &c->d_bucket, &c->d_slots,
                    &c->d_alltot, &c->d_rslots, &c->d_seg, &c->d_lb, &c->d_peerinfo, &c->d_dd, &c->d_pend, &c->d_grp, &c->d_val[0], &c->d_val[1], &c->d_rscratch,
                    &c->d_cscratch, &c->d_tsum, &c->d_tbase, &c->d_swords, &c->d_scnt, &c->d_spos, &c->d_srid, &c->d_owords, &c->d_ocnt, &c->d_oocc_off, &c->d_opos, &c->d_orid,
                    &c->d_hist, &c->d_cursor};
    for (auto *b : db) b->release();
    for (int h = 0; h < 2; ++h) for (int w = 0; w < MAX_WORDS; ++w) c->d_keys[h][w].release();
    HostBuf *hb[] = {&c->h_peerinfo, &c->h_bucket, &c->h_alltot, &c->h_meta, &c->h_cursor, &c->h_owords, &c->h_ocnt,
                     &c->h_oocc_off, &c->h_opos, &c->h_orid, &c->h_hist};
    for (auto *b : hb) b->release();
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

// ---- extraction (stages 1+2) -------------------------------------------------------------------------
// Bins per rank: fixed by the config, or sized so that a bin holds on average half of what the on-chip
// path is sized for (bins.cu: bin_target_kmers); every rank must use the same number, so the largest input of
// any rank decides.

static int choose_bins(hsk_ctx *c, u64 nbytes)
{
    if (c->cfg.buckets_per_rank > 0) return 0;
    u64 mx = nbytes;
    if (c->cfg.nranks > 1) {
        // words 10 / 11 of the cursor block: not shared with anything the kernels of this call use
        CK(c->d_cursor.ensure(128));
        CK(c->h_cursor.ensure(128));
        c->h_cursor.as<u64>()[10] = nbytes;
        CK(cudaMemcpyAsync(c->d_cursor.as<u64>() + 10, c->h_cursor.as<u64>() + 10, 8, cudaMemcpyHostToDevice, c->stream));
        NK(g_nccl.AllReduce(c->d_cursor.as<u64>() + 10, c->d_cursor.as<u64>() + 11, 1, ncclUint64, ncclMax, c->comm, c->stream));
        CK(cudaMemcpyAsync(c->h_cursor.as<u64>() + 11, c->d_cursor.as<u64>() + 11, 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        mx = c->h_cursor.as<u64>()[11];
    }
    // average occurrences per bin: sized so that the distinct k-mers of a bin fill about a third of its table
    u64 target = (u64)bin_target_kmers(c->nwords, c->cfg.ext != 0);
    if (const char *ev = getenv("HSK_TARGET_BIN")) { u64 v = strtoull(ev, nullptr, 10); if (v >= 64) target = v; }
    u64 tg = (mx * 4 + target - 1) / target;
    tg = std::max<u64>(64, (tg + 63) / 64 * 64);
    tg = std::min<u64>(tg, MAX_BINS / (u64)c->cfg.nranks);
    c->tg = (u32)tg;
    c->tt = c->tg * (u32)c->cfg.nranks;
    return 0;
}

// Device layout of d_bucket (u64 units): [bin_tot T][start T+1][meta 8: run_cursor, kmers_total, check_slots,
// check_kmers, -][cursor T].  bin_tot = bt_pack(slots, k-mers).  Host (h_meta): S, -, run cursor, k-mer total, the two
// independent totals of pass A (common.cuh: a bin that outgrew a field of its total is an error, not a corrupt stream).  With full_d2h bin_tot and
// start are also copied to h_bucket (debug entry point).  d_slots receives the bin-major supermer slots.
// extract_count: tile table, pass A (per-bin totals + run list), bin scan; `before_sync` may queue more work that
// only needs the totals (multi-rank bookkeeping) before the one host synchronisation.  extract_scatter: pass B into
// c->xp.out_base with the cursors in `d_cur`.
static int stage_chunk(hsk_ctx *c, size_t ci);
static void end_input(hsk_ctx *c);

static int extract_count(hsk_ctx *c, const u8 *d_packed, u64 nbytes, u64 nbytes_padded, const u64 *d_read_off,
                         const u32 *d_read_len, u64 nreads, int readid_base, const std::function<int()> &before_sync)
{
    if (choose_bins(c, nbytes)) return 1;
    const u32 T = c->tt;
    cudaStream_t s = c->stream;
    const bool ext = c->cfg.ext != 0;
    const int w = c->cfg.k - c->m_eff + 1;
    ExtractParams &P = c->xp;
    memset(&P, 0, sizeof(P));
    P.packed = d_packed; P.nbytes = nbytes; P.nbytes_padded = nbytes_padded;
    P.read_off = d_read_off; P.read_len = d_read_len; P.nreads = nreads;
    P.out_slots = (u32)xt_out_slots(w);
    const u64 nslots = nbytes * 4;
    P.ntiles = (nslots + P.out_slots - 1) / P.out_slots;
    const u32 nctas = c->x_nctas = extract_grid(w, c->sm_count);
    const u64 nwarps = (u64)nctas * XT_WARPS;
    P.tile_begin = 0; P.tile_end = P.ntiles;
    P.tiles_per_warp = (u32)((P.ntiles + nwarps - 1) / nwarps);
    P.k = c->cfg.k; P.m = c->m_eff; P.nbins = T; P.readid_base = readid_base;
    P.slot_nmax = (u32)(slot_max_bases(c->nwords, ext) - c->cfg.k + 1);
    P.slot_ninv = (u32)(((1ull << 32) + P.slot_nmax - 1) / P.slot_nmax);

    const size_t host_u64 = (size_t)T + ((size_t)T + 1);
    const size_t dev_u64 = host_u64 + XT_META + (size_t)T;
    CK(c->d_bucket.ensure(dev_u64 * 8));
    CK(c->h_meta.ensure(128 * 8));
    CK(c->d_tile_hdr.ensure((P.ntiles + 1) * sizeof(ulonglong2)));
    CK(c->d_tile_read.ensure((P.ntiles + 2) * sizeof(u32)));
    CK(c->d_bscratch.ensure(bin_scan_scratch_bytes(T)));
    P.tile_read = c->d_tile_read.as<u32>();
    u64 *d_tot = c->d_bucket.as<u64>();
    u64 *d_start = d_tot + T, *d_runcur = d_start + T + 1, *d_ktot = d_runcur + 1;
    u64 *d_cur = d_runcur + XT_META;
    u64 *hm = c->h_meta.as<u64>();

    // the run list: about one entry per 8 k-mers on reads; sized for one per 6 slots, with a retry at the worst case
    u64 run_cap = nslots / 6 + 1024;
    c->begin(c->ev_extract);
    CK(launch_tile_reads(P, c->d_tile_read.as<u32>(), s));
    c->stats.n_launches += 1;
    for (int attempt = 0;; ++attempt) {
        if (c->big_mode && run_cap * 8 <= c->d_big.cap) c->run_list = c->d_big.as<u64>();   // (the arena of the last call is over)
        else { CK(c->d_run_list.ensure(run_cap * 8)); c->run_list = c->d_run_list.as<u64>(); }
        CK(cudaMemsetAsync(d_tot, 0, (host_u64 + XT_META) * 8, s));
        // pass A over the tiles whose bytes have arrived (hsk_count uploads the reads in chunks); a retry and
        // hsk_count_device see the whole buffer at once
        u64 tb = 0;
        const size_t nchunks = attempt == 0 ? c->in_chunks.size() : 0;
        for (size_t ci = 0; ci <= nchunks; ++ci) {
            u64 te = P.ntiles;
            if (ci < nchunks) {
                if (stage_chunk(c, ci)) return 1;   // pageable input: copied to the staging ring and sent now
                CK(cudaStreamWaitEvent(s, c->in_chunks[ci].ready, 0));
                te = std::min<u64>(c->in_chunks[ci].tile_end, P.ntiles);
                if (ci + 1 == nchunks) te = P.ntiles;
            }
            if (te <= tb) continue;
            P.tile_begin = tb; P.tile_end = te;
            P.tiles_per_warp = (u32)((te - tb + nwarps - 1) / nwarps);
            CK(launch_supermer_count(P, nctas, d_tot, c->run_list, c->d_tile_hdr.as<ulonglong2>(), d_runcur, run_cap, s));
            c->stats.n_launches += 1;
            tb = te;
        }
        CK(launch_bin_scan(d_tot, T, d_start, d_cur, d_ktot, c->d_bscratch.as<u64>(), s));
        c->end(c->ev_extract);
        if (before_sync && before_sync()) return 1;
        CK(cudaMemcpyAsync(hm + 0, d_start + T, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(hm + 2, d_runcur, 32, cudaMemcpyDeviceToHost, s));   // run cursor, k-mer total, check totals
        hm[6] = 0;
        if (c->d_in_flags) CK(cudaMemcpyAsync(hm + 6, c->d_in_flags, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        c->stats.n_launches += 3;
        g_trace.mark("pass A + bin scan done (host sync)");
        if (hm[6] & 1) return fail("a read is longer than 2^32-1 bases");
        if (hm[6] & 2) return fail("DnaBuffer size %llu does not match the read lengths", (unsigned long long)nbytes);
        if (hm[2] <= run_cap) {
            if (hm[0] != hm[4] || hm[3] != hm[5])
                return fail("a minimizer bin holds more than 2^%d supermers or 2^%d k-mers of this rank's reads (totals %llu / %llu, summed "
                            "per bin %llu / %llu)", 64 - BT_KBITS, BT_KBITS, (unsigned long long)hm[4], (unsigned long long)hm[5],
                            (unsigned long long)hm[0], (unsigned long long)hm[3]);
            break;
        }
        if (attempt) return fail("internal: run list overflow after resize");
        run_cap = nslots + 1024;   // pathological input (runs shorter than 6 k-mers on average): worst-case list
        c->begin(c->ev_extract);
    }
    P.tile_begin = 0; P.tile_end = P.ntiles;
    P.tiles_per_warp = (u32)((P.ntiles + nwarps - 1) / nwarps);
    c->stats.n_supermers = hm[0];
    c->stats.supermer_bytes = hm[0] * (u64)slot_words(c->nwords, ext) * 4;
    c->stats.n_kmers_local = hm[3];
    return 0;
}

static int extract_scatter(hsk_ctx *c, u64 *d_cur)
{
    cudaStream_t s = c->stream;
    ExtractParams P = c->xp;
    // the scatter pass is latency-bound (one atomic + one store per supermer) and light on registers: more CTAs per SM
    // than the count pass
    u32 per_sm = 8;
    if (const char *ev = getenv("HSK_SCATTER_CTAS")) { const int v = atoi(ev); if (v >= 1 && v <= 8) per_sm = (u32)v; }
    const u32 nctas = (u32)c->sm_count * per_sm;
    const u64 nwarps = (u64)nctas * XT_WARPS;
    P.tiles_per_warp = (u32)((P.ntiles + nwarps - 1) / nwarps);
    c->begin(c->ev_extract);
    if (P.ntiles) CK(launch_supermer_scatter(P, nctas, c->nwords, c->cfg.ext != 0, c->run_list,
                                             c->d_tile_hdr.as<ulonglong2>(), d_cur, s));
    c->end(c->ev_extract);
    g_trace.mark("scatter enqueued");
    c->stats.n_launches += 1;
    return 0;
}

// single-rank extraction into c->d_slots (also the debug entry point)
static int run_extract(hsk_ctx *c, const u8 *d_packed, u64 nbytes, u64 nbytes_padded, const u64 *d_read_off,
                       const u32 *d_read_len, u64 nreads, int readid_base, bool full_d2h)
{
    if (extract_count(c, d_packed, nbytes, nbytes_padded, d_read_off, d_read_len, nreads, readid_base, nullptr)) return 1;
    const u32 T = c->tt;
    const int SW = slot_words(c->nwords, c->cfg.ext != 0);
    const u64 S = c->stats.n_supermers;
    CK(c->d_slots.ensure((S + 4) * (size_t)SW * 4));
    c->xp.out_stream = c->d_slots.as<u32>();
    u64 *d_cur = c->d_bucket.as<u64>() + (size_t)T + (T + 1) + XT_META;
    if (extract_scatter(c, d_cur)) return 1;
    if (full_d2h) {
        const size_t host_u64 = (size_t)T + ((size_t)T + 1);
        CK(c->h_bucket.ensure((host_u64 + 1) * 8));
        CK(cudaMemcpyAsync(c->h_bucket.p, c->d_bucket.p, host_u64 * 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

// ---- HBM path for the bins the on-chip path left over (skewed bins): expand -> radix sort -> count ------
struct OvfSeg { int src; u32 bin; u64 i0, nslots, kmers; };
struct CountLimits { u64 arena_cap, occ_cap; u32 *err; };

// segs: the segments of the overflow bins, all sources of a bin next to each other; a batch holds whole bins
static int run_hbm_path(hsk_ctx *c, const std::vector<OvfSeg> &segs, const CountLimits &lim)
{
    cudaStream_t s = c->stream;
    const int NW = c->nwords;
    const bool ext = c->cfg.ext != 0;
    u64 cap = c->cfg.batch_kmers ? c->cfg.batch_kmers : (1ull << 28);
    cap = std::min<u64>(cap, (1ull << 29) - 1);
    size_t first = 0;
    while (first < segs.size()) {   // batches of whole bins
        size_t last = first;
        u64 n = 0, max_sup = 0;
        while (last < segs.size()) {
            size_t e = last;
            u64 nb = 0, ms = 0;
            while (e < segs.size() && segs[e].bin == segs[last].bin) { nb += segs[e].kmers; ms = std::max(ms, segs[e].nslots); ++e; }
            if (n != 0 && n + nb > cap) break;
            n += nb; max_sup = std::max(max_sup, ms); last = e;
        }
        if (n > (1ull << 29) - 1)
            return fail("a minimizer bin holds %llu k-mers (> 2^29-1); raise buckets_per_rank", (unsigned long long)n);
        for (int h = 0; h < 2; ++h) {
            for (int w = 0; w < NW; ++w) CK(c->d_keys[h][w].ensure((n + 8) * 8));
            if (ext) CK(c->d_val[h].ensure((n + 8) * 8));
        }
        CK(c->d_rscratch.ensure(radix_scratch_bytes(n)));
        CK(c->d_cscratch.ensure(count_scratch_bytes(n)));
        const u64 seg_tiles = (max_sup + XP_TILE - 1) / XP_TILE + 1;
        CK(c->d_tsum.ensure(seg_tiles * sizeof(u32)));
        CK(c->d_tbase.ensure(seg_tiles * sizeof(u64)));
        Planes A, B;
        for (int w = 0; w < MAX_WORDS; ++w) {
            A.p[w] = w < NW ? c->d_keys[0][w].as<u64>() : nullptr;
            B.p[w] = w < NW ? c->d_keys[1][w].as<u64>() : nullptr;
        }
        u64 *VA = ext ? c->d_val[0].as<u64>() : nullptr, *VB = ext ? c->d_val[1].as<u64>() : nullptr;

        c->begin(c->ev_expand);
        u64 out_base = 0;
        for (size_t i = first; i < last; ++i) {
            const OvfSeg &g = segs[i];
            if (g.nslots == 0) continue;
            const int SW = slot_words(NW, ext);
            ExpandSegment seg;
            seg.slots = c->src_slots[g.src] + g.i0 * (u64)SW;
            seg.nslots = g.nslots;
            seg.out_base = out_base;
            CK(launch_expand(seg, c->cfg.k, NW, ext, c->d_tsum.as<u32>(), c->d_tbase.as<u64>(), A, VA, s));
            c->stats.n_launches += 3;
            out_base += g.kmers;
        }
        c->end(c->ev_expand);
        if (out_base != n) return fail("internal: batch k-mer count mismatch");

        c->begin(c->ev_sort);
        bool in_b = false; int np = 0, nl = 0;
        EvPair pp{c->ev(), c->ev()};
        c->ev_pass.push_back(pp);
        CK(launch_radix_sort(A, B, VA, VB, n, NW, c->cfg.k, c->d_rscratch.p, &in_b, &np, &nl, s, pp.a, pp.b));
        c->end(c->ev_sort);
        c->stats.n_sort_passes = (u64)np;
        c->stats.n_launches += (u64)nl;

        c->begin(c->ev_count);
        CountParams CP;
        CP.keys = in_b ? B : A;
        CP.val = ext ? (in_b ? VB : VA) : nullptr;
        CP.n = n; CP.nwords = NW; CP.lower = (u32)c->cfg.lower; CP.upper = (u32)c->cfg.upper;
        CP.out_words = c->d_owords.as<u64>(); CP.out_cnt = c->d_ocnt.as<u32>();
        CP.out_occ_off = c->d_oocc_off.as<u64>(); CP.out_pos = c->d_opos.as<u32>(); CP.out_rid = c->d_orid.as<int>();
        CP.histogram = c->d_hist.as<u64>(); CP.cursor = c->d_cursor.as<u64>();
        CP.arena_cap = lim.arena_cap; CP.occ_cap = lim.occ_cap; CP.err = lim.err;
        CK(launch_count_filter(CP, c->d_cscratch.p, s));
        c->end(c->ev_count);
        c->stats.n_launches += 3;
        c->stats.n_batches += 1;
        first = last;
    }
    return 0;
}

// ---- the whole path on device-resident reads -------------------------------------------------------
static int count_device(hsk_ctx *c, const u8 *d_packed, u64 nbytes, u64 nbytes_padded, const u64 *d_read_off,
                        const u32 *d_read_len, u64 nreads, int readid_base)
{
    CK(cudaSetDevice(c->cfg.device));
    cudaStream_t s = c->stream;
    const int G = c->cfg.nranks, me = c->cfg.rank;
    const int NW = c->nwords;
    const bool ext = c->cfg.ext != 0;
    if (G > BN_MAX_SRC) return fail("more than %d ranks are not supported yet", BN_MAX_SRC);
    c->ev_extract.clear(); c->ev_exchange.clear(); c->ev_expand.clear(); c->ev_sort.clear(); c->ev_count.clear(); c->ev_pass.clear();
    c->ev_bins.clear();
    hsk_stats keep = c->stats;
    memset(&c->stats, 0, sizeof(c->stats));
    c->stats.ms_h2d = keep.ms_h2d;
    c->have_result = false;
    cudaEvent_t ev_t0 = c->ev(), ev_t1 = c->ev();
    CK(cudaEventRecord(ev_t0, s));

    // ---- small device state of this call:
    //      [cursor 2][owned total 1][ticket, ovf_count (u32 x2)][stage cursor 2][big_count (u32)][error flags (u32)]
    //      [barrier words 2][choose_bins words 2]
    CK(c->d_cursor.ensure(128));
    CK(c->h_cursor.ensure(128));
    CK(cudaMemsetAsync(c->d_cursor.p, 0, 128, s));
    u64 *d_cursor = c->d_cursor.as<u64>();
    u64 *d_owned = d_cursor + 2;
    u32 *d_ticket = reinterpret_cast<u32 *>(d_cursor + 3), *d_ovfc = d_ticket + 1;
    u64 *d_stagecur = d_cursor + 4;
    u32 *d_bigc = reinterpret_cast<u32 *>(d_cursor + 6);
    u32 *d_err = reinterpret_cast<u32 *>(d_cursor + 7);

    BinParams BP;
    memset(&BP, 0, sizeof(BP));
    BP.k = c->cfg.k; BP.lower = (u32)c->cfg.lower; BP.upper = (u32)c->cfg.upper;
    BP.nsrc = G;
    c->rbase_idx.assign(G, 0);
    u64 owned = 0;
    const int SW = slot_words(NW, ext);
    u32 T = 0, TG = 0, b_lo = 0;
    u64 *d_tot = nullptr, *d_start = nullptr;
    u64 *hm = nullptr;

    if (G == 1) {
        if (run_extract(c, d_packed, nbytes, nbytes_padded, d_read_off, d_read_len, nreads, readid_base, false)) return 1;
        T = c->tt; TG = c->tg;
        d_tot = c->d_bucket.as<u64>(); d_start = d_tot + T;
        hm = c->h_meta.as<u64>();
        BP.slots[0] = c->d_slots.as<u32>();
        BP.seg_start[0] = d_start;
        BP.bin_kmers = d_tot;
        owned = c->stats.n_kmers_local;
    } else {
        // ---- stages 1-3 across ranks.  Every rank scatters its supermers into its own bin-major stream; after the
        //      count pass the bin totals of every rank are all-gathered, so that everybody knows where the bins it owns
        //      lie inside every rank's stream.  Then either
        //      (default) nothing is shipped at all: the streams are exchanged once as CUDA IPC handles and the bin
        //      kernel reads the supermers of its bins in place, out of the peers' memory over NVLink, while it counts
        //      (fused exchange + count: the all-to-all costs no pass and no buffer of its own), or
        //      (HSK_EXCHANGE=nccl) whole bin ranges are shipped with grouped ncclSend/ncclRecv first.
        u64 *d_seg_start = nullptr, *d_binkm = nullptr, *d_meta = nullptr;
        auto bookkeeping = [&]() -> int {
            T = c->tt; TG = c->tg; b_lo = (u32)me * TG;
            d_tot = c->d_bucket.as<u64>(); d_start = d_tot + T;
            hm = c->h_meta.as<u64>();
            CK(c->d_alltot.ensure((size_t)G * T * 8));
            CK(c->d_seg.ensure(((size_t)G * (TG + 1) + TG + G + 8) * 8));
            d_seg_start = c->d_seg.as<u64>();
            d_binkm = d_seg_start + (size_t)G * (TG + 1); d_meta = d_binkm + TG;
            c->begin(c->ev_exchange);
            NK(g_nccl.AllGather(d_tot, c->d_alltot.p, (size_t)T, ncclUint64, c->comm, s));
            CK(cudaMemsetAsync(d_owned, 0, 8, s));
            CK(launch_seg_scan(c->d_alltot.as<u64>(), T, me, TG, G, c->use_p2p, d_seg_start, d_meta, d_binkm, d_owned, s));
            c->end(c->ev_exchange);
            c->stats.n_launches += 3;
            CK(cudaMemcpyAsync(hm + 8, d_meta, (size_t)G * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(hm + 7, d_owned, 8, cudaMemcpyDeviceToHost, s));
            // local bin starts at the rank boundaries: what goes to every peer
            CK(cudaMemcpy2DAsync(hm + 8 + G, 8, d_start, (size_t)TG * 8, 8, (size_t)G + 1, cudaMemcpyDeviceToHost, s));
            return 0;
        };
        if (extract_count(c, d_packed, nbytes, nbytes_padded, d_read_off, d_read_len, nreads, readid_base, bookkeeping)) return 1;
        owned = hm[7];
        const u64 *rtot = hm + 8, *bounds = hm + 8 + (size_t)G;
        u64 *d_cur = d_start + T + 1 + XT_META;
        const u64 S = c->stats.n_supermers;
        // Every rank has passed the all-gather above, so every rank has finished its previous call: supermer buffers of
        // mine that peers had mapped then and that were replaced since are no longer mapped anywhere.
        for (auto &r : c->retired) r.release();
        c->retired.clear();
        {
            const size_t want = (S + S / 8 + 64) * (size_t)SW * 4;
            if (want > c->d_slots.cap && c->d_slots.p && c->use_p2p) {
                // peers still have the old buffer mapped (CUDA IPC) until they have seen the new record: keep it until
                // the next call instead of freeing it under them
                c->retired.push_back(c->d_slots);
                c->d_slots = DevBuf();
            }
            CK(c->d_slots.ensure(want));
        }
        c->xp.out_stream = c->d_slots.as<u32>();
        for (int peer = 0; peer < G; ++peer) {
            if (peer == me) continue;
            c->stats.bytes_sent += (bounds[peer + 1] - bounds[peer]) * SW * 4;
            c->stats.bytes_received += rtot[peer] * SW * 4;
        }
        if (c->use_p2p) {
            // tell the peers where my stream is: IPC handle + process id + pointer, all-gathered every call (128 bytes a
            // rank); a peer is (re)mapped only when its record changes
            CK(c->h_peerinfo.ensure((size_t)(G + 1) * 128));
            CK(c->d_peerinfo.ensure((size_t)(G + 1) * 128));
            unsigned char *mine = c->h_peerinfo.as<unsigned char>() + (size_t)G * 128;
            memset(mine, 0, 128);
            cudaIpcMemHandle_t hnd;
            CK(cudaIpcGetMemHandle(&hnd, c->d_slots.p));
            static_assert(sizeof(hnd) <= 64, "ipc handle size");
            memcpy(mine, &hnd, sizeof(hnd));
            const u64 pid = (u64)getpid(), ptr = (u64)(uintptr_t)c->d_slots.p, cap = (u64)c->d_slots.cap;
            memcpy(mine + 64, &pid, 8); memcpy(mine + 72, &ptr, 8); memcpy(mine + 80, &cap, 8);
            c->begin(c->ev_exchange);
            CK(cudaMemcpyAsync(c->d_peerinfo.as<unsigned char>() + (size_t)G * 128, mine, 128, cudaMemcpyHostToDevice, s));
            NK(g_nccl.AllGather(c->d_peerinfo.as<unsigned char>() + (size_t)G * 128, c->d_peerinfo.p, 128, ncclUint8, c->comm, s));
            CK(cudaMemcpyAsync(c->h_peerinfo.p, c->d_peerinfo.p, (size_t)G * 128, cudaMemcpyDeviceToHost, s));
            c->end(c->ev_exchange);
            // pass B into my own stream while the records travel
            if (extract_scatter(c, d_cur)) return 1;
            // every rank's stream has to be complete before anybody reads it: a one-word all-reduce is the barrier
            c->begin(c->ev_exchange);
            NK(g_nccl.AllReduce(d_cursor + 8, d_cursor + 9, 1, ncclUint64, ncclSum, c->comm, s));
            c->end(c->ev_exchange);
            CK(cudaStreamSynchronize(s));
            for (int p = 0; p < G; ++p) {
                const unsigned char *info = c->h_peerinfo.as<unsigned char>() + (size_t)p * 128;
                hsk_ctx::PeerMap &pm = c->peers[p];
                BP.seg_start[p] = d_seg_start + (size_t)p * (TG + 1);
                if (p == me) { BP.slots[p] = c->d_slots.as<u32>(); continue; }
                if (!pm.valid || memcmp(pm.info, info, 128) != 0) {
                    if (pm.valid && pm.ipc && pm.mapped) cudaIpcCloseMemHandle(pm.mapped);
                    pm.valid = false;
                    u64 ppid, pptr;
                    memcpy(&ppid, info + 64, 8); memcpy(&pptr, info + 72, 8);
                    if (ppid == pid) {
                        // a context of the same process: plain peer access
                        cudaPointerAttributes at;
                        CK(cudaPointerGetAttributes(&at, (void *)(uintptr_t)pptr));
                        if (at.device != c->cfg.device) {
                            cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
                            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail("cudaDeviceEnablePeerAccess(%d): %s", at.device, cudaGetErrorString(e));
                            (void)cudaGetLastError();
                        }
                        pm.mapped = (void *)(uintptr_t)pptr; pm.ipc = false;
                    } else {
                        cudaIpcMemHandle_t ph;
                        memcpy(&ph, info, sizeof(ph));
                        cudaError_t e = cudaIpcOpenMemHandle(&pm.mapped, ph, cudaIpcMemLazyEnablePeerAccess);
                        if (e != cudaSuccess) return fail("cudaIpcOpenMemHandle(rank %d): %s (set HSK_EXCHANGE=nccl where peer access is unavailable)", p, cudaGetErrorString(e));
                        pm.ipc = true;
                    }
                    memcpy(pm.info, info, 128);
                    pm.valid = true;
                }
                BP.slots[p] = reinterpret_cast<const u32 *>(pm.mapped);
            }
        } else {
            if (extract_scatter(c, d_cur)) return 1;
            u64 ri = 0;
            for (int src = 0; src < G; ++src) {
                if (src == me) continue;
                c->rbase_idx[src] = ri;
                ri += rtot[src];
            }
            CK(c->d_rslots.ensure((ri + 4) * (size_t)SW * 4));
            c->begin(c->ev_exchange);
            NK(g_nccl.GroupStart());
            for (int peer = 0; peer < G; ++peer) {
                if (peer == me) continue;
                const u64 si = bounds[peer], sn = bounds[peer + 1] - si;
                const u64 rn = rtot[peer];
                if (sn) NK(g_nccl.Send(c->d_slots.as<u32>() + si * SW, sn * SW, ncclUint32, peer, c->comm, s));
                if (rn) NK(g_nccl.Recv(c->d_rslots.as<u32>() + c->rbase_idx[peer] * SW, rn * SW, ncclUint32, peer, c->comm, s));
            }
            NK(g_nccl.GroupEnd());
            c->end(c->ev_exchange);
            for (int src = 0; src < G; ++src) {
                const bool local = (src == me);
                BP.slots[src] = local ? c->d_slots.as<u32>() : c->d_rslots.as<u32>() + c->rbase_idx[src] * (u64)SW;
                // the local stream is addressed with its own (absolute) bin starts, the received ones with the scanned tables
                BP.seg_start[src] = local ? d_start + b_lo : d_seg_start + (size_t)src * (TG + 1);
            }
        }
        BP.bin_kmers = d_binkm;
    }
    BP.nbins = TG;
    for (int src = 0; src < G; ++src) c->src_slots[src] = BP.slots[src];
    c->stats.n_kmers_owned = owned;

    // ---- result arena, staging area, per-bin records.  No more than owned / LOWER entries can be kept; when arena +
    //      staging area of that size do not fit the free memory, the run list (dead once the scatter pass has run) is
    //      given back and both are cut to what there is: the kernels check the limits (arena full: the call fails;
    //      staging area full: the bin takes the HBM path).
    const u64 bound = owned / (u64)c->cfg.lower + 8;
    u64 arena_cap = bound, stage_cap = bound, occ_cap = owned + 8, stage_occ_cap = owned + 8;
    {
        const size_t ent_b = (size_t)NW * 8 + 4 + (ext ? 8 : 0);
        const size_t need = 2 * (size_t)bound * ent_b + (ext ? 2 * (size_t)(owned + 8) * 8 : 0);
        u64 budget = 0;
        if (const char *ev = getenv("HSK_ARENA_BUDGET_MB")) budget = strtoull(ev, nullptr, 10) << 20;   // tests
        DevBuf *ar[] = {&c->d_owords, &c->d_ocnt, &c->d_swords, &c->d_scnt, &c->d_oocc_off, &c->d_opos, &c->d_orid, &c->d_spos, &c->d_srid};
        if (c->big_mode && c->big_budget != budget) {
            // (tests) the block was sized by HSK_ARENA_BUDGET_MB and the setting changed: start over.  The scatter pass, which
            // may be reading the run list in the block, is complete after the synchronisation.
            CK(cudaStreamSynchronize(s));
            for (auto *r : ar) r->release();
            c->d_big.release();
            c->big_mode = false;
        }
        if (!c->big_mode) {
            size_t have = 0, free_b = 0, total_b = 0;
            for (auto *r : ar) have += r->cap;
            if (need > have && !budget) CK(cudaMemGetInfo(&free_b, &total_b));   // (a slow driver call: only when something has to grow)
            if (budget || (need > have && (double)(need + need / 16) > 0.92 * (double)(free_b + have))) {
                // enter the memory-short mode: from now on the run list and the arena share one block
                CK(cudaStreamSynchronize(s));
                c->d_run_list.release();
                for (auto *r : ar) r->release();
                CK(cudaMemGetInfo(&free_b, &total_b));
                CK(c->d_big.ensure_exact(budget ? (size_t)budget : (size_t)(0.85 * (double)free_b)));
                c->big_mode = true;
                c->big_budget = budget;
                g_trace.mark("memory is short: the run list and the arena share one block from now on");
            }
        }
        if (c->big_mode) {
            // the arena is laid out in the block (the scatter pass, which reads the run list there, is ahead of the bin
            // kernel on the stream); the kernels check the limits
            const u64 B = c->d_big.cap > 8192 ? c->d_big.cap - 8192 : 0;
            if (!ext) {
                arena_cap = std::min<u64>(bound, B / 9 * 8 / ent_b);
                stage_cap = std::min<u64>(bound, B / 9 / ent_b);
            } else {
                arena_cap = std::min<u64>(bound, B / 100 * 40 / ent_b);
                stage_cap = std::min<u64>(bound, B / 100 * 10 / ent_b);
                occ_cap = std::min<u64>(owned + 8, B / 100 * 40 / 8);
                stage_occ_cap = std::min<u64>(owned + 8, B / 100 * 10 / 8);
            }
            u8 *at = c->d_big.as<u8>();
            auto carve = [&](DevBuf &b, size_t bytes) { b.set_view(at, bytes); at += (bytes + 255) & ~(size_t)255; };
            carve(c->d_owords, arena_cap * NW * 8);
            carve(c->d_ocnt, arena_cap * 4);
            carve(c->d_swords, stage_cap * NW * 8);
            carve(c->d_scnt, stage_cap * 4);
            if (ext) {
                carve(c->d_oocc_off, (arena_cap + 1) * 8);
                carve(c->d_opos, occ_cap * 4);
                carve(c->d_orid, occ_cap * 4);
                carve(c->d_spos, stage_occ_cap * 4);
                carve(c->d_srid, stage_occ_cap * 4);
            }
            if ((size_t)(at - c->d_big.as<u8>()) > c->d_big.cap) return fail("internal: arena layout exceeds its block");
        }
    }
    CK(c->d_owords.ensure(arena_cap * NW * 8));
    CK(c->d_ocnt.ensure(arena_cap * 4));
    CK(c->d_swords.ensure(stage_cap * NW * 8));
    CK(c->d_scnt.ensure(stage_cap * 4));
    if (ext) {
        CK(c->d_oocc_off.ensure((arena_cap + 1) * 8));
        CK(c->d_opos.ensure(occ_cap * 4));
        CK(c->d_orid.ensure(occ_cap * 4));
        CK(c->d_spos.ensure(stage_occ_cap * 4));
        CK(c->d_srid.ensure(stage_occ_cap * 4));
    }
    const size_t hist_bins = (size_t)c->cfg.upper + 1;
    CK(c->d_hist.ensure(hist_bins * 8));
    CK(cudaMemsetAsync(c->d_hist.p, 0, hist_bins * 8, s));
    // per-bin records: look-back cells (zeroed), staging records of the big bins, overflow / big lists
    const bool streaming = c->stream_result;
    int NG = (streaming && TG >= 2048) ? 32 : 1;
    if (const char *ev = getenv("HSK_GROUPS")) { const int v = atoi(ev); if (v >= 1 && v <= 64 && streaming) NG = v; }
    const u32 group_bins = (TG + (u32)NG - 1) / (u32)NG;
    NG = group_bins ? (int)((TG + group_bins - 1) / group_bins) : 1;
    CK(c->d_lb.ensure(((size_t)8 * TG + 16) * 8 + ((size_t)2 * TG + 16) * 4));
    CK(c->d_grp.ensure((size_t)NG * 24 + 64));
    CK(c->h_grp.ensure((size_t)NG * 32 + 64));
    CK(cudaMemsetAsync(c->d_lb.p, 0, (size_t)2 * TG * 8, s));
    CK(cudaMemsetAsync(c->d_grp.p, 0, (size_t)NG * 24 + 64, s));
    BP.lb_state = c->d_lb.as<u64>();
    BP.bin_rec = BP.lb_state + (size_t)2 * TG; BP.fin = BP.bin_rec + (size_t)4 * TG;
    BP.st_words = c->d_swords.as<u64>(); BP.st_cnt = c->d_scnt.as<u32>();
    BP.st_pos = c->d_spos.as<u32>(); BP.st_rid = c->d_srid.as<int>();
    BP.stage_cursor = d_stagecur;
    BP.stage_cap = stage_cap; BP.stage_occ_cap = stage_occ_cap;
    BP.out_words = c->d_owords.as<u64>(); BP.out_cnt = c->d_ocnt.as<u32>();
    BP.out_occ_off = c->d_oocc_off.as<u64>(); BP.out_pos = c->d_opos.as<u32>(); BP.out_rid = c->d_orid.as<int>();
    BP.histogram = c->d_hist.as<u64>(); BP.cursor = d_cursor;
    BP.arena_cap = arena_cap; BP.occ_cap = occ_cap; BP.err = d_err;
    BP.ticket = d_ticket; BP.ovf_count = d_ovfc;
    BP.ovf_list = reinterpret_cast<u32 *>(BP.fin + (size_t)2 * TG + 8);
    BP.big_list = BP.ovf_list + TG + 4; BP.big_count = d_bigc;
    BP.group_bins = group_bins ? group_bins : 1;
    {
        const char *ev = getenv("HSK_DEDUP");
        if (NW <= 2 && !ext && !(ev && *ev == '0')) {
            CK(c->d_dd.ensure(bin_dedup_scratch_bytes(c->sm_count, SW)));
            BP.dd_slots = c->d_dd.as<uint4>();
            BP.dd_mult = reinterpret_cast<u32 *>(BP.dd_slots + (size_t)c->sm_count * BN_MAX_CTAS * BN_DDLIMIT_MAX * (SW / 4));
        }
    }
    if (!ext) {
        CK(c->d_pend.ensure(bin_pending_scratch_bytes(c->sm_count, NW)));
        BP.pend_words = c->d_pend.as<u64>();
        BP.pend_cnt = reinterpret_cast<u32 *>(BP.pend_words + (size_t)c->sm_count * BN_MAX_CTAS * BN_SORTCAP * NW);
    }
    BP.walk_split = 1; BP.walk_min = 8;
    if (const char *ev = getenv("HSK_WALK_SPLIT")) { const int v = atoi(ev); if (v >= 1 && v <= 16) BP.walk_split = (u32)v; }
    if (const char *ev = getenv("HSK_WALK_MIN")) { const int v = atoi(ev); if (v >= 1 && v <= 32) BP.walk_min = (u32)v; }
    BP.grp_end = c->d_grp.as<u64>();
    BP.grp_done = reinterpret_cast<u32 *>(BP.grp_end + (size_t)2 * NG); BP.grp_big = BP.grp_done + NG;
    volatile u64 *snap = c->h_grp.as<u64>();
    BP.snap = streaming ? c->h_grp.as<u64>() : nullptr;
    for (int i = 0; i < 4 * NG; ++i) snap[i] = 0;

    // ---- stages 4+5 on chip: one persistent launch over all bins.  hsk_count streams the arena to the host group by
    //      group while the kernel is still running: the CTA that completes a group says so in page-locked memory.
    const size_t hist_bytes = hist_bins * 8;
    u64 sent_kept = 0, sent_occ = 0;
    cudaEvent_t ev_d0 = nullptr, ev_d1 = nullptr;
    if (streaming) { ev_d0 = c->ev(); ev_d1 = c->ev(); CK(cudaEventRecord(ev_d0, c->copy_stream)); }
    const bool sinking = streaming && c->sink_fn != nullptr;
    auto refresh_view = [&]() {
        hsk_result &v = c->sink_view;
        v.nwords = NW;
        v.kmer_words = c->h_owords.as<uint64_t>(); v.cnt = c->h_ocnt.as<u32>();
        v.occ_off = ext ? c->h_oocc_off.as<uint64_t>() : nullptr;
        v.pos = ext ? c->h_opos.as<u32>() : nullptr; v.rid = ext ? c->h_orid.as<int32_t>() : nullptr;
        v.histogram = c->h_hist.as<uint64_t>();
    };
    if (sinking) { refresh_view(); c->sink->begin(c->sink_fn, c->sink_user, &c->sink_view, c->cfg.device); }
    int sink_rc = 0;
    // copy arena entries [sent_kept, kept_end) / occurrences [sent_occ, occ_end) (after `ready`, if given) and hand them
    // to the sink
    auto send_result = [&](u64 kept_end, u64 occ_end, u64 kept_total_hint, cudaEvent_t ready, bool last) -> int {
        const u64 want = std::max<u64>(kept_end, kept_total_hint) + 1;
        const bool grow_k = want * NW * 8 > c->h_owords.cap || want * 4 > c->h_ocnt.cap || (ext && (want + 1) * 8 > c->h_oocc_off.cap);
        const bool grow_o = ext && ((occ_end + 1) * 4 > c->h_opos.cap);
        if (grow_k || grow_o) {
            // the arrays move: nobody may be reading or writing them
            CK(cudaStreamSynchronize(c->copy_stream));
            if (sinking) sink_rc |= c->sink->drain();
            if (grow_k) {
                CK(c->h_owords.ensure_keep(want * NW * 8, sent_kept * NW * 8));
                CK(c->h_ocnt.ensure_keep(want * 4, sent_kept * 4));
                if (ext) CK(c->h_oocc_off.ensure_keep((want + 1) * 8, sent_kept * 8));
            }
            if (grow_o) {
                const u64 wo = occ_end + (kept_end ? (u64)((double)occ_end / (double)kept_end * (double)(want - kept_end)) : 0) + 1;
                CK(c->h_opos.ensure_keep(wo * 4, sent_occ * 4));
                CK(c->h_orid.ensure_keep(wo * 4, sent_occ * 4));
            }
            refresh_view();
        }
        cudaStream_t cs = c->copy_stream;
        if (ready) CK(cudaStreamWaitEvent(cs, ready, 0));
        const u64 nk = kept_end - sent_kept, no = occ_end - sent_occ;
        if (nk) {
            CK(cudaMemcpyAsync(c->h_owords.as<u64>() + sent_kept * NW, c->d_owords.as<u64>() + sent_kept * NW, nk * NW * 8, cudaMemcpyDeviceToHost, cs));
            CK(cudaMemcpyAsync(c->h_ocnt.as<u32>() + sent_kept, c->d_ocnt.as<u32>() + sent_kept, nk * 4, cudaMemcpyDeviceToHost, cs));
            if (ext) CK(cudaMemcpyAsync(c->h_oocc_off.as<u64>() + sent_kept, c->d_oocc_off.as<u64>() + sent_kept, nk * 8, cudaMemcpyDeviceToHost, cs));
        }
        if (ext && no) {
            CK(cudaMemcpyAsync(c->h_opos.as<u32>() + sent_occ, c->d_opos.as<u32>() + sent_occ, no * 4, cudaMemcpyDeviceToHost, cs));
            CK(cudaMemcpyAsync(c->h_orid.as<int>() + sent_occ, c->d_orid.as<int>() + sent_occ, no * 4, cudaMemcpyDeviceToHost, cs));
        }
        if (sinking && (nk || last)) {
            cudaEvent_t arrived = c->ev();
            CK(cudaEventRecord(arrived, cs));
            const u64 hint = last ? kept_end : std::max<u64>(kept_end, kept_total_hint);
            u64 f = sent_kept;
            do {   // pieces of SINK_SUBPART entries, each to whichever delivery thread is free
                const u64 n = std::min<u64>(SINK_SUBPART, kept_end - f);
                const bool head = f == sent_kept, tail = f + n == kept_end;
                c->sink->push({f, n, !ext ? 0 : (head ? sent_occ : ~0ull), !ext ? 0 : (tail ? occ_end : ~0ull), hint, arrived});
                f += n;
            } while (f < kept_end);
        }
        sent_kept = kept_end; sent_occ = occ_end;
        return 0;
    };
    c->begin(c->ev_bins);
    CK(launch_bin_count(BP, NW, ext, c->sm_count, s));
    c->end(c->ev_bins);
    c->stats.n_launches += 1;
    cudaEvent_t ev_bins_done = c->ev();
    CK(cudaEventRecord(ev_bins_done, s));
    if (streaming) {
        // follow the kernel: a finished group whose bins are all in place goes out at once; after the first group
        // that waits for the big gather the rest is sent at the end
        for (int g = 0; g < NG; ++g) {
            while (snap[4 * g + 3] == 0) {
                if (cudaEventQuery(ev_bins_done) != cudaErrorNotReady) break;   // finished (or failed): the final sync tells
            }
            if (snap[4 * g + 3] == 0 || snap[4 * g + 2] != 0) break;
            g_trace.mark("bin group in the arena");
            const u64 kept_end = snap[4 * g], occ_end = snap[4 * g + 1];
            if (kept_end > arena_cap || occ_end > occ_cap) break;   // the arena overflowed: reported below
            const u64 hint = (u64)((double)kept_end * (double)NG / (double)(g + 1) * 1.15) + 4096;
            if (send_result(kept_end, occ_end, hint, nullptr, false)) return 1;
        }
    }
    c->stats.n_batches = 1;
    CK(cudaMemcpyAsync(c->h_cursor.p, c->d_cursor.p, 64, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const char *arena_msg = "the result (%llu entries) does not fit the memory of the GPU next to the reads and supermers (room for %llu): "
                            "split the input over more ranks";
    if (reinterpret_cast<u32 *>(c->h_cursor.as<u64>() + 7)[0] & 1u) {
        if (sinking) c->sink->drain();
        return fail(arena_msg, (unsigned long long)c->h_cursor.as<u64>()[0], (unsigned long long)arena_cap);
    }
    const u32 novf = reinterpret_cast<u32 *>(c->h_cursor.as<u64>() + 3)[1];
    c->stats.n_overflow_bins = novf;
    if (*reinterpret_cast<u32 *>(c->h_cursor.as<u64>() + 6)) {   // bins that keep more k-mers than a CTA sorts
        CK(launch_bin_gather_big(BP, NW, ext, c->sm_count, s));
        c->stats.n_launches += 1;
    }

    if (novf) {
        // ---- leftovers through HBM: fetch the tables of the overflow bins.  Segments are listed bin by bin (all sources
        //      of a bin next to each other), so that a batch never holds a part of a bin.
        std::vector<u32> ovf(novf);
        CK(cudaMemcpy(ovf.data(), BP.ovf_list, (size_t)novf * 4, cudaMemcpyDeviceToHost));
        std::sort(ovf.begin(), ovf.end());
        std::vector<OvfSeg> segs((size_t)novf * G);
        std::vector<u64> hs((size_t)TG + 1), hk((size_t)TG);
        for (int src = 0; src < G; ++src) {   // tables of all bins in one copy per source (only when something overflowed)
            CK(cudaMemcpy(hs.data(), BP.seg_start[src], ((size_t)TG + 1) * 8, cudaMemcpyDeviceToHost));
            const u64 *kp = (G == 1) ? d_tot : c->d_alltot.as<u64>() + (size_t)src * T + b_lo;
            CK(cudaMemcpy(hk.data(), kp, (size_t)TG * 8, cudaMemcpyDeviceToHost));
            for (u32 i = 0; i < novf; ++i) {
                const u32 lb = ovf[i];
                segs[(size_t)i * G + src] = {src, lb, hs[lb], hs[lb + 1] - hs[lb], bt_kmers(hk[lb])};
            }
        }
        CountLimits lim{arena_cap, occ_cap, d_err};
        if (run_hbm_path(c, segs, lim)) { if (sinking) c->sink->drain(); return 1; }
        CK(cudaMemcpyAsync(c->h_cursor.p, c->d_cursor.p, 64, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaEventRecord(ev_t1, s));
    CK(cudaStreamSynchronize(s));
    if (reinterpret_cast<u32 *>(c->h_cursor.as<u64>() + 7)[0] & 1u) {
        if (sinking) c->sink->drain();
        return fail(arena_msg, (unsigned long long)c->h_cursor.as<u64>()[0], (unsigned long long)arena_cap);
    }
    c->n_kept = c->h_cursor.as<u64>()[0];
    c->n_occ = c->h_cursor.as<u64>()[1];
    if (ext) {   // closing offset of the occurrence lists
        CK(cudaMemcpyAsync(c->d_oocc_off.as<u64>() + c->n_kept, c->h_cursor.as<u64>() + 1, 8, cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
    }
    if (streaming) {
        // what the HBM path appended, the histogram, and the end of the copies
        CK(c->h_hist.ensure(hist_bytes));
        if (sinking) refresh_view();
        CK(cudaMemcpyAsync(c->h_hist.p, c->d_hist.p, hist_bytes, cudaMemcpyDeviceToHost, c->copy_stream));
        if (send_result(c->n_kept, c->n_occ, c->n_kept, ev_t1, true)) return 1;
        CK(cudaEventRecord(ev_d1, c->copy_stream));
        g_trace.mark("last copies enqueued");
        CK(cudaStreamSynchronize(c->copy_stream));
        g_trace.mark("copies done");
        if (ext) c->h_oocc_off.as<u64>()[c->n_kept] = c->n_occ;
        CK(cudaEventElapsedTime(&c->stats.ms_d2h, ev_d0, ev_d1));
        if (sinking) {
            sink_rc |= c->sink->drain();
            g_trace.mark("sink done");
            if (sink_rc) return fail("hsk_count_stream: the sink failed (status %d)", sink_rc);
        }
    }
    c->stats.ms_extract = hsk_ctx::sum_ms(c->ev_extract);
    c->stats.ms_exchange = hsk_ctx::sum_ms(c->ev_exchange);
    c->stats.ms_expand = hsk_ctx::sum_ms(c->ev_expand);
    c->stats.ms_sort = hsk_ctx::sum_ms(c->ev_sort);
    c->stats.ms_count = hsk_ctx::sum_ms(c->ev_count);
    c->stats.ms_sort_passes = hsk_ctx::sum_ms(c->ev_pass);
    c->stats.ms_bins = hsk_ctx::sum_ms(c->ev_bins);
    CK(cudaEventElapsedTime(&c->stats.ms_total, ev_t0, ev_t1));
    c->have_result = true;
    return 0;
}

static void fill_device_result(hsk_ctx *c, hsk_device_result *out)
{
    out->nwords = c->nwords;
    out->n_kept = c->n_kept;
    out->n_occ = c->n_occ;
    out->d_kmer_words = c->d_owords.as<uint64_t>();
    out->d_cnt = c->d_ocnt.as<u32>();
    out->d_occ_off = c->cfg.ext ? c->d_oocc_off.as<uint64_t>() : nullptr;
    out->d_pos = c->cfg.ext ? c->d_opos.as<u32>() : nullptr;
    out->d_rid = c->cfg.ext ? c->d_orid.as<int32_t>() : nullptr;
    out->d_histogram = c->d_hist.as<uint64_t>();
    out->stats = c->stats;
}

int hsk_count_device(hsk_ctx *c, const uint8_t *d_packed, uint64_t nbytes, const uint64_t *d_read_off,
                     const uint32_t *d_read_len, uint64_t nreads, int32_t readid_base, hsk_device_result *out)
{
    if (!c || !out) return fail("hsk_count_device: null argument");
    if (((uintptr_t)d_packed & 15) != 0) return fail("hsk_count_device: d_packed must be 16-byte aligned");
    c->stats.ms_h2d = 0;
    c->ev_used = 0;
    end_input(c);
    if (count_device(c, d_packed, nbytes, (nbytes + 15) & ~15ull, (const u64 *)d_read_off, d_read_len, nreads, readid_base)) {
        const std::string msg = g_err;   // nothing of the failed call stays in flight; the first error is the one to report
        cudaStreamSynchronize(c->stream);
        cudaStreamSynchronize(c->copy_stream);
        (void)cudaGetLastError();
        g_err = msg;
        return 1;
    }
    fill_device_result(c, out);
    return 0;
}

int hsk_fetch_result(hsk_ctx *c, hsk_result *out)
{
    if (!c || !out) return fail("hsk_fetch_result: null argument");
    if (!c->have_result) return fail("hsk_fetch_result: no result on this context");
    CK(cudaSetDevice(c->cfg.device));
    cudaStream_t s = c->stream;
    const int NW = c->nwords;
    const bool ext = c->cfg.ext != 0;
    const size_t hist_bins = (size_t)c->cfg.upper + 1;
    CK(c->h_owords.ensure((c->n_kept + 1) * NW * 8));
    CK(c->h_ocnt.ensure((c->n_kept + 1) * 4));
    CK(c->h_hist.ensure(hist_bins * 8));
    if (ext) {
        CK(c->h_oocc_off.ensure((c->n_kept + 1) * 8));
        CK(c->h_opos.ensure((c->n_occ + 1) * 4));
        CK(c->h_orid.ensure((c->n_occ + 1) * 4));
    }
    cudaEvent_t e0 = c->ev(), e1 = c->ev();
    CK(cudaEventRecord(e0, s));
    if (c->n_kept) {
        CK(cudaMemcpyAsync(c->h_owords.p, c->d_owords.p, c->n_kept * NW * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(c->h_ocnt.p, c->d_ocnt.p, c->n_kept * 4, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaMemcpyAsync(c->h_hist.p, c->d_hist.p, hist_bins * 8, cudaMemcpyDeviceToHost, s));
    if (ext) {
        CK(cudaMemcpyAsync(c->h_oocc_off.p, c->d_oocc_off.p, (c->n_kept + 1) * 8, cudaMemcpyDeviceToHost, s));
        if (c->n_occ) {
            CK(cudaMemcpyAsync(c->h_opos.p, c->d_opos.p, c->n_occ * 4, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(c->h_orid.p, c->d_orid.p, c->n_occ * 4, cudaMemcpyDeviceToHost, s));
        }
    }
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&c->stats.ms_d2h, e0, e1));
    out->nwords = NW;
    out->n_kept = c->n_kept;
    out->n_occ = c->n_occ;
    out->kmer_words = c->h_owords.as<uint64_t>();
    out->cnt = c->h_ocnt.as<u32>();
    out->occ_off = ext ? c->h_oocc_off.as<uint64_t>() : nullptr;
    out->pos = ext ? c->h_opos.as<u32>() : nullptr;
    out->rid = ext ? c->h_orid.as<int32_t>() : nullptr;
    out->histogram = c->h_hist.as<uint64_t>();
    out->stats = c->stats;
    return 0;
}

// host -> device staging of a DnaBuffer.  The 64-bit read lengths go up as they are and reads.cu turns them into
// byte offsets + 32-bit lengths on the device; the packed bytes go up in chunks on the copy stream, every chunk
// with an event and the number of extraction tiles it completes, so that pass A of the extraction starts on the
// first chunk while the others are still on the bus.  Page-locked input is sent from where it is; pageable input
// (a DnaBuffer is a plain heap array) is copied chunk by chunk into a ring of page-locked buffers by the context's
// host threads, each chunk right before the extraction waits for it (stage_chunk).
static bool is_pageable(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

// pageable input: waits until the pool's threads have copied extraction chunk ci into its ring slot and sent it (the
// thread that finishes the last piece of a chunk enqueues the chunk's one H2D copy and records its event)
static int stage_chunk(hsk_ctx *c, size_t ci)
{
    if (!c->in_pageable) return 0;
    while (!c->chunk_sent[ci].load(std::memory_order_acquire)) {
        if (c->stage_err.load()) return fail("staging the input failed: %s", cudaGetErrorString((cudaError_t)c->stage_err.load()));
        std::this_thread::yield();
    }
    if (c->stage_err.load()) return fail("staging the input failed: %s", cudaGetErrorString((cudaError_t)c->stage_err.load()));
    if (ci + 1 == c->in_chunks.size()) {
        CK(cudaEventRecord(c->ev_h2d[1], c->copy_stream));
        if (g_trace.on) {
            char msg[200];
            snprintf(msg, sizeof(msg), "all input chunks staged + enqueued (threads: %.3f ms in memcpy, %.3f ms in CUDA calls, %.3f ms waiting for ring slots)",
                     c->stage_ns[0].load() * 1e-6, c->stage_ns[1].load() * 1e-6, c->stage_ns[2].load() * 1e-6);
            g_trace.mark(msg);
        }
    } else if (ci == 0) g_trace.mark("first input chunk staged");
    return 0;
}

static int stage_input(hsk_ctx *c, const uint8_t *packed, uint64_t nbytes, const uint64_t *read_len, uint64_t nreads)
{
    CK(cudaSetDevice(c->cfg.device));
    cudaStream_t s = c->stream, cs = c->copy_stream;
    const u64 padded = ((nbytes + 15) & ~15ull) + 64;
    CK(c->d_packed.ensure(padded));
    CK(c->d_len64.ensure((nreads + 1) * 8));
    CK(c->d_read_off.ensure((nreads + 2) * 8));
    CK(c->d_read_len.ensure((nreads + 2) * 4));
    const size_t rts = read_table_scratch_bytes(nreads);
    CK(c->d_rtscratch.ensure(rts + 16));
    c->d_in_flags = reinterpret_cast<u32 *>(c->d_rtscratch.as<u8>() + rts);
    c->in_pageable = nbytes > 0 && is_pageable(packed);
    c->in_host = packed;
    CK(cudaMemsetAsync(c->d_in_flags, 0, 16, s));
    if (nreads) {
        const uint64_t *lens = read_len;
        if (nreads >= 4096 && is_pageable(read_len)) {   // a large pageable array would be staged by the driver, one thread
            CK(c->h_len.ensure(nreads * 8));
            u64 *dst = c->h_len.as<u64>();
            c->pool->parallel_for((nreads + 511) / 512, [=](u64 lo, u64 hi) {
                const u64 i0 = lo * 512, i1 = std::min<u64>(hi * 512, nreads);
                if (i1 > i0) stream_copy(dst + i0, read_len + i0, (i1 - i0) * 8);
            });
            lens = reinterpret_cast<const uint64_t *>(dst);
        }
        CK(cudaMemcpyAsync(c->d_len64.p, lens, nreads * 8, cudaMemcpyHostToDevice, s));
    }
    CK(launch_read_table(c->d_len64.as<u64>(), nreads, nbytes, c->d_read_off.as<u64>(), c->d_read_len.as<u32>(),
                         c->d_rtscratch.as<u64>(), c->d_in_flags, s));
    c->stats.n_launches += 3;

    const u32 OL = (u32)xt_out_lanes(c->cfg.k - c->m_eff + 1);
    c->in_chunks.clear();
    cudaEvent_t e0 = c->ev(), e1 = c->ev();
    c->ev_h2d[0] = e0; c->ev_h2d[1] = e1;   // ms_h2d is read after the call has synchronised
    CK(cudaEventRecord(e0, cs));
    CK(cudaMemsetAsync(c->d_packed.as<u8>() + (nbytes & ~15ull), 0, padded - (nbytes & ~15ull), cs));
    const u64 PIECE = hsk_ctx::PIECE;
    u64 chunk = std::min<u64>(std::max<u64>((nbytes / 8 + 4095) & ~4095ull, 2ull << 20), 256ull << 20);
    if (c->in_pageable) chunk = (std::min<u64>(chunk, 32ull << 20) + PIECE - 1) / PIECE * PIECE;   // ring slots of whole pieces
    for (u64 o = 0; o < nbytes; o += chunk) {
        const u64 n = std::min<u64>(chunk, nbytes - o);
        cudaEvent_t ev = c->ev();
        if (!c->in_pageable) {
            CK(cudaMemcpyAsync(c->d_packed.as<u8>() + o, packed + o, n, cudaMemcpyHostToDevice, cs));
            CK(cudaEventRecord(ev, cs));
        }
        const u64 words = (o + n) / 4;   // a tile reads 34 words from its first one
        c->in_chunks.push_back({words >= 34 ? (words - 34) / OL + 1 : 0, ev, o, n, (u32)((n + PIECE - 1) / PIECE)});
    }
    if (!c->in_pageable || nbytes == 0) { CK(cudaEventRecord(e1, cs)); return 0; }

    // pageable: the pool's threads take 1 MiB pieces in order and copy them into the ring slot of their extraction chunk;
    // whoever completes a chunk sends it with ONE cudaMemcpyAsync (a copy per piece costs ~50 us of fixed latency each on
    // this platform) and records the chunk's event.  A slot is reused once the chunk that used it before has left.
    const u64 npieces = (nbytes + PIECE - 1) / PIECE;
    const u64 nchunks = c->in_chunks.size();
    const u64 nring = std::min<u64>(nchunks, 4);
    CK(c->h_ring.ensure(nring * chunk));
    while (c->ring_free.size() < nring) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->ring_free.push_back(e);
    }
    if (c->flags_cap < nchunks) {
        c->flags_cap = nchunks * 2;
        c->chunk_sent.reset(new std::atomic<u32>[c->flags_cap]);    // per chunk: its copy + event are enqueued
        c->chunk_done.reset(new std::atomic<u32>[c->flags_cap]);   // per chunk: pieces copied into the slot
    }
    for (u64 i = 0; i < nchunks; ++i) { c->chunk_sent[i].store(0, std::memory_order_relaxed); c->chunk_done[i].store(0, std::memory_order_relaxed); }
    c->stage_err.store(0);
    for (auto &x : c->stage_ns) x.store(0);
    const bool timing = g_trace.on;
    const u64 ppc = chunk / PIECE;
    const int device = c->cfg.device;
    c->pool->start(npieces, [c, nring, nbytes, PIECE, ppc, chunk, device, timing](u64 j) {
        auto tick = [] { return std::chrono::steady_clock::now(); };
        auto ns = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count();
        };
        const auto t0 = tick();
        static thread_local int dev_set = -1;
        if (dev_set != device) { cudaSetDevice(device); dev_set = device; }
        const u64 ci = j / ppc, slot = ci % nring, off = j * PIECE, n = std::min<u64>(PIECE, nbytes - off);
        const hsk_ctx::InChunk &ch = c->in_chunks[ci];
        cudaError_t e = cudaSuccess;
        if (ci >= nring && !c->stage_err.load(std::memory_order_relaxed)) {
            // the chunk that used the slot before must have left (pieces are taken in order: all of its pieces were taken earlier)
            while (!c->chunk_sent[ci - nring].load(std::memory_order_acquire)) std::this_thread::yield();
            e = cudaEventSynchronize(c->ring_free[slot]);
        }
        u8 *dst = c->h_ring.as<u8>() + slot * chunk;
        const auto t1 = tick();
        if (e == cudaSuccess && !c->stage_err.load(std::memory_order_relaxed)) stream_copy(dst + (off - ch.off), c->in_host + off, n);
        const auto t2 = tick();
        if (e != cudaSuccess) c->stage_err.store((int)e);
        if (c->chunk_done[ci].fetch_add(1, std::memory_order_acq_rel) + 1 == ch.pieces) {   // the chunk is complete: send it
            if (!c->stage_err.load()) {
                e = cudaMemcpyAsync(c->d_packed.as<u8>() + ch.off, dst, ch.n, cudaMemcpyHostToDevice, c->copy_stream);
                if (e == cudaSuccess) e = cudaEventRecord(ch.ready, c->copy_stream);
                if (e == cudaSuccess) e = cudaEventRecord(c->ring_free[slot], c->copy_stream);
                if (e != cudaSuccess) c->stage_err.store((int)e);
            }
            c->chunk_sent[ci].store(1, std::memory_order_release);
        }
        if (timing) { const auto t3 = tick(); c->stage_ns[0] += ns(t1, t2); c->stage_ns[1] += ns(t2, t3); c->stage_ns[2] += ns(t0, t1); }
    });
    return 0;
}

static void end_input(hsk_ctx *c)
{
    if (c->pool) c->pool->wait();   // (an error may have left the staging job running)
    c->in_chunks.clear();
    c->d_in_flags = nullptr;
    c->in_host = nullptr;
    c->in_pageable = false;
}

int hsk_count_stream(hsk_ctx *c, const uint8_t *packed, uint64_t nbytes, const uint64_t *read_len, uint64_t nreads,
                     int32_t readid_base, hsk_sink_fn sink, void *user, hsk_result *out)
{
    if (!c || !out) return fail("hsk_count: null argument");
    if (nreads && (!read_len)) return fail("hsk_count: null read_len");
    if (nbytes && !packed) return fail("hsk_count: null packed buffer");
    c->ev_used = 0;
    hsk_stats keep;
    memset(&keep, 0, sizeof(keep));
    c->stats = keep;
    g_trace.start();
    g_trace.mark("hsk_count begin");
    if (stage_input(c, packed, nbytes, read_len, nreads)) { end_input(c); return 1; }
    g_trace.mark("input copies enqueued");
    const u64 staged_launches = c->stats.n_launches;
    c->stream_result = true;
    c->sink_fn = sink; c->sink_user = user;
    if (sink && !c->sink) c->sink.reset(new SinkPool(c->host_threads));
    const int rc = count_device(c, c->d_packed.as<u8>(), nbytes, ((nbytes + 15) & ~15ull) + 64, c->d_read_off.as<u64>(),
                                c->d_read_len.as<u32>(), nreads, readid_base);
    c->stream_result = false;
    c->sink_fn = nullptr; c->sink_user = nullptr;
    end_input(c);
    if (rc) {
        const std::string msg = g_err;   // the first error is the one to report
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamSynchronize(c->stream);
        if (c->sink) c->sink->drain();
        (void)cudaGetLastError();
        g_err = msg;
        return 1;
    }
    g_trace.mark("hsk_count end");
    c->stats.n_launches += staged_launches;
    CK(cudaEventElapsedTime(&c->stats.ms_h2d, c->ev_h2d[0], c->ev_h2d[1]));
    if (g_trace.on) {
        char msg[200];
        snprintf(msg, sizeof(msg), "device: h2d %.3f ms, extract %.3f ms, bins %.3f ms, d2h %.3f ms, first kernel to last %.3f ms", c->stats.ms_h2d,
                 c->stats.ms_extract, c->stats.ms_bins, c->stats.ms_d2h, c->stats.ms_total);
        g_trace.mark(msg);
    }
    const bool ext = c->cfg.ext != 0;
    out->nwords = c->nwords;
    out->n_kept = c->n_kept;
    out->n_occ = c->n_occ;
    out->kmer_words = c->h_owords.as<uint64_t>();
    out->cnt = c->h_ocnt.as<u32>();
    out->occ_off = ext ? c->h_oocc_off.as<uint64_t>() : nullptr;
    out->pos = ext ? c->h_opos.as<u32>() : nullptr;
    out->rid = ext ? c->h_orid.as<int32_t>() : nullptr;
    out->histogram = c->h_hist.as<uint64_t>();
    out->stats = c->stats;
    return 0;
}

int hsk_count(hsk_ctx *c, const uint8_t *packed, uint64_t nbytes, const uint64_t *read_len, uint64_t nreads,
              int32_t readid_base, hsk_result *out)
{
    return hsk_count_stream(c, packed, nbytes, read_len, nreads, readid_base, nullptr, nullptr, out);
}

int hsk_host_register(void *p, size_t bytes, int32_t device)
{
    if (!p || !bytes) return fail("hsk_host_register: null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { (void)cudaGetLastError(); return fail("hsk_host_register: no such device"); }
    CK(cudaSetDevice(device));
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail("cudaHostRegister: %s", cudaGetErrorString(e)); }
    return 0;
}

int hsk_host_unregister(void *p)
{
    if (!p) return fail("hsk_host_unregister: null argument");
    cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail("cudaHostUnregister: %s", cudaGetErrorString(e)); }
    return 0;
}

int hsk_allreduce_histogram(hsk_ctx *c, uint64_t *hist)
{
    if (!c || !hist) return fail("hsk_allreduce_histogram: null argument");
    if (!c->have_result) return fail("hsk_allreduce_histogram: no result on this context");
    CK(cudaSetDevice(c->cfg.device));
    const size_t bins = (size_t)c->cfg.upper + 1;
    CK(c->h_hist.ensure(bins * 8));
    if (c->cfg.nranks > 1) {
        CK(c->d_alltot.ensure(bins * 8));
        NK(g_nccl.AllReduce(c->d_hist.p, c->d_alltot.p, bins, ncclUint64, ncclSum, c->comm, c->stream));
        CK(cudaMemcpyAsync(c->h_hist.p, c->d_alltot.p, bins * 8, cudaMemcpyDeviceToHost, c->stream));
    } else {
        CK(cudaMemcpyAsync(c->h_hist.p, c->d_hist.p, bins * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    memcpy(hist, c->h_hist.p, bins * 8);
    return 0;
}

int hsk_fill_entries(hsk_ctx *c, void *entries, uint64_t capacity_entries)
{
    if (!c || !entries) return fail("hsk_fill_entries: null argument");
    if (!c->have_result) return fail("hsk_fill_entries: no result on this context");
    if (capacity_entries < c->n_kept) return fail("hsk_fill_entries: capacity %llu < %llu entries", (unsigned long long)capacity_entries, (unsigned long long)c->n_kept);
    if (c->h_owords.cap < c->n_kept * c->nwords * 8) return fail("hsk_fill_entries: call hsk_count / hsk_fetch_result first");
    const int NW = c->nwords;
    const u64 *w = c->h_owords.as<u64>();
    const u32 *cnt = c->h_ocnt.as<u32>();
    u64 *dst = reinterpret_cast<u64 *>(entries);
    // SoA -> {TKmer kmer; uint64_t cnt;} entries (reference KmerListEntryS, include/kmer.hpp:368-407) by the context's
    // host threads: the loop is memory-bound and the list has millions of entries
    const u64 n = c->n_kept;
    c->pool->parallel_for((n + 4095) / 4096, [=](u64 lo, u64 hi) {
        const u64 i1 = std::min<u64>(hi * 4096, n);
        for (u64 i = lo * 4096; i < i1; ++i) {
            for (int l = 0; l < NW; ++l) dst[i * (NW + 1) + l] = w[i * NW + l];
            dst[i * (NW + 1) + NW] = cnt[i];
        }
    });
    return 0;
}

// ---- stage-level entry points -----------------------------------------------------------------------

int hsk_debug_sort(hsk_ctx *c, uint64_t *const *d_keys, uint64_t *const *d_tmp, uint64_t *d_val, uint64_t *d_val_tmp, uint64_t n,
                   int32_t nwords, int32_t k)
{
    if (!c) return fail("hsk_debug_sort: null ctx");
    if (nwords != nwords_for_k(k)) return fail("hsk_debug_sort: nwords does not match k");
    CK(cudaSetDevice(c->cfg.device));
    cudaStream_t s = c->stream;
    Planes A, B;
    for (int w = 0; w < MAX_WORDS; ++w) { A.p[w] = w < nwords ? (u64 *)d_keys[w] : nullptr; B.p[w] = w < nwords ? (u64 *)d_tmp[w] : nullptr; }
    CK(c->d_rscratch.ensure(radix_scratch_bytes(n)));
    bool in_b = false; int np = 0, nl = 0;
    c->ev_used = 0;
    cudaEvent_t e0 = c->ev(), e1 = c->ev();
    CK(cudaEventRecord(e0, s));
    CK(launch_radix_sort(A, B, (u64 *)d_val, (u64 *)d_val_tmp, n, nwords, k, c->d_rscratch.p, &in_b, &np, &nl, s));
    CK(cudaEventRecord(e1, s));
    if (in_b) {
        for (int w = 0; w < nwords; ++w) CK(cudaMemcpyAsync(d_keys[w], d_tmp[w], n * 8, cudaMemcpyDeviceToDevice, s));
        if (d_val) CK(cudaMemcpyAsync(d_val, d_val_tmp, n * 8, cudaMemcpyDeviceToDevice, s));
    }
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&c->stats.ms_sort, e0, e1));
    c->stats.n_sort_passes = (u64)np;
    return 0;
}

int hsk_debug_extract(hsk_ctx *c, const uint8_t *packed, uint64_t nbytes, const uint64_t *read_len, uint64_t nreads,
                      int32_t readid_base, hsk_supermers *out)
{
    if (!c || !out) return fail("hsk_debug_extract: null argument");
    c->ev_used = 0;
    c->ev_extract.clear();
    if (stage_input(c, packed, nbytes, read_len, nreads)) { end_input(c); return 1; }
    if (run_extract(c, c->d_packed.as<u8>(), nbytes, ((nbytes + 15) & ~15ull) + 64, c->d_read_off.as<u64>(), c->d_read_len.as<u32>(), nreads,
                    readid_base, true)) { end_input(c); cudaStreamSynchronize(c->copy_stream); return 1; }
    end_input(c);
    const u32 T = c->tt;
    const int SW = slot_words(c->nwords, c->cfg.ext != 0);
    const u64 *hb = c->h_bucket.as<u64>();
    const u64 S = hb[(size_t)T + T];
    c->h_dbg.resize(2 * (size_t)T);
    for (u32 b = 0; b < T; ++b) { c->h_dbg[b] = bt_slots(hb[b]); c->h_dbg[T + b] = bt_kmers(hb[b]); }
    CK(c->h_owords.ensure((S + 1) * (size_t)SW * 4));
    if (S) CK(cudaMemcpyAsync(c->h_owords.p, c->d_slots.p, S * (size_t)SW * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    out->n_bins = T;
    out->bin_slots = reinterpret_cast<const uint64_t *>(c->h_dbg.data());
    out->bin_kmers = reinterpret_cast<const uint64_t *>(c->h_dbg.data() + T);
    out->n_slots = S;
    out->slot_words = (uint32_t)SW;
    out->slots = c->h_owords.as<uint32_t>();
    c->have_result = false;
    return 0;
}

} // extern "C"
