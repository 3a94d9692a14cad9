/*
 * MPI stand-in for hosts without an MPI installation (this image has none): the subset of the MPI C API
 * that `hysortk.hpp`-style code uses, for ONE rank or for SEVERAL PROCESSES OF ONE NODE.
 *
 *   single rank     (default)  rank 0 of 1, collectives degenerate to local copies.
 *   several ranks   when the environment of every process carries
 *                       HSK_MPI_SIZE=<n>  HSK_MPI_RANK=<0..n-1>  HSK_MPI_SESSION=<unique string>
 *                   (set by hysortk_b200/shim/hsk_mpirun.cpp, by the tests, and by bench.py under torchrun).
 *                   The ranks meet in one POSIX shared-memory object /hsk_mpi_<session>; every rank owns a
 *                   staging slot in it (HSK_MPI_SLOT_MB, default 4 MiB).  A collective is a sequence of numbered
 *                   rounds: a rank publishes (descriptor, window of its send buffer) in its slot, the others
 *                   copy the byte range they need; a slot is rewritten only after every rank has finished
 *                   reading the previous round.  No helper threads, no sockets.
 *
 * With a real MPI on the include path this header is simply not used (Makefile: MPI_INC / MPI_LIB).
 *
 * Covers the calls made on the kmer_count path and its I/O neighbours (reference call sites:
 * kmerops.cpp:66,711,782,919,940,1172,1198,1287,1325; hysortk.cpp:104,115,156; fastaindex.cpp:137,168-185,
 * 223-224; logger.cpp:127,137; timer.hpp:26-51; memcheck.cpp:81), all on MPI_COMM_WORLD, called from one
 * thread per process (MPI_THREAD_FUNNELED), one non-blocking request outstanding at a time
 * (kmerops.cpp:919-968).
 *
 * Header-only C++17 (inline functions with one shared state per process); not usable from C.
 */
#ifndef HSK_MPI_SHIM_H_
#define HSK_MPI_SHIM_H_

#ifndef __cplusplus
#error "hysortk_b200/shim/mpi.h is a C++17 header"
#endif

#include <fcntl.h>
#include <sched.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <string>
#include <vector>

#define HSK_MPI_SHIM 2

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Info;
typedef int MPI_File;
typedef long long MPI_Offset;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_SUCCESS 0
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_IN_PLACE ((void *)-1)
#define MPI_INFO_NULL 0
#define MPI_MODE_RDONLY 2
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
/* datatype handle = class << 24 | element size in bytes; class: 0 opaque, 1 signed, 2 unsigned, 3 floating point */
#define HSK_MPI_TYPE(cls, size) (((cls) << 24) | (size))
#define MPI_BYTE HSK_MPI_TYPE(2, 1)
#define MPI_CHAR HSK_MPI_TYPE(1, 1)
#define MPI_INT HSK_MPI_TYPE(1, 4)
#define MPI_UNSIGNED HSK_MPI_TYPE(2, 4)
#define MPI_FLOAT HSK_MPI_TYPE(3, 4)
#define MPI_DOUBLE HSK_MPI_TYPE(3, 8)
#define MPI_LONG HSK_MPI_TYPE(1, 8)
#define MPI_UNSIGNED_LONG HSK_MPI_TYPE(2, 8)
#define MPI_LONG_LONG HSK_MPI_TYPE(1, 8)
#define MPI_UNSIGNED_LONG_LONG HSK_MPI_TYPE(2, 8)
#define MPI_UINT64_T HSK_MPI_TYPE(2, 8)
#define MPI_INT64_T HSK_MPI_TYPE(1, 8)

namespace hsk_mpi {

inline size_t type_size(MPI_Datatype t) { return (size_t)(t & 0xFFFFFF); }
inline int type_class(MPI_Datatype t) { return (t >> 24) & 0xF; }

/* ---- shared-memory object ------------------------------------------------------------------------------- */
struct alignas(64) RankCell {
    volatile uint64_t published;   /* last round whose slot contents are complete */
    volatile uint64_t consumed;    /* last round this rank has finished reading   */
    uint64_t len;                  /* bytes of the rank's whole contribution to the current collective */
    uint64_t desc_len;
};
struct Header {
    volatile uint64_t magic;
    volatile uint64_t abort_code;  /* non-zero: some rank called MPI_Abort */
    uint64_t nranks, slot_bytes;
    volatile uint64_t attached, detached;
};
constexpr uint64_t MAGIC = 0x48534B4D50493032ull;   /* "HSKMPI02" */
constexpr size_t DESC_MAX = 4096;                   /* descriptor area at the start of every slot */

struct State {
    bool init = false;
    int rank = 0, size = 1;
    std::string shm_name;
    Header *hdr = nullptr;
    RankCell *cells = nullptr;
    char *slots = nullptr;
    size_t slot_bytes = 0, map_bytes = 0;
    uint64_t seq = 0;              /* rounds completed so far (identical on every rank) */
    double timeout_s = 600.0;
    /* the one outstanding MPI_Ialltoall */
    bool pending = false;
    void *pend_recv = nullptr;
    size_t pend_bytes = 0;
};

inline State &state() { static State s; return s; }
inline void finalize();
inline void hsk_mpi_detach_at_exit() { finalize(); }

inline double now_s()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

[[noreturn]] inline void die(const char *what)
{
    State &S = state();
    fprintf(stderr, "[hsk_mpi rank %d/%d] %s\n", S.rank, S.size, what);
    if (S.hdr) S.hdr->abort_code = 134;
    _exit(134);
}

/* spin until cond(); yields the CPU (ranks may outnumber cores), watches the abort flag and a deadline */
template <typename F>
inline void wait_until(F cond, const char *what)
{
    State &S = state();
    unsigned spins = 0;
    double t0 = 0;
    while (!cond()) {
        if (++spins < 200) {
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
            continue;
        }
        if (S.hdr && S.hdr->abort_code) _exit((int)S.hdr->abort_code);
        sched_yield();
        if ((spins & 0x3FF) == 0) {
            const double t = now_s();
            if (t0 == 0) t0 = t;
            else if (t - t0 > S.timeout_s) die(what);
        }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
}

inline void ensure_init()
{
    State &S = state();
    if (S.init) return;
    S.init = true;
    const char *es = getenv("HSK_MPI_SIZE"), *er = getenv("HSK_MPI_RANK"), *ek = getenv("HSK_MPI_SESSION");
    const int n = es ? atoi(es) : 1;
    if (n <= 1) return;
    if (!er || !ek || !*ek) { fprintf(stderr, "[hsk_mpi] HSK_MPI_SIZE=%d needs HSK_MPI_RANK and HSK_MPI_SESSION\n", n); _exit(2); }
    S.size = n;
    S.rank = atoi(er);
    if (S.rank < 0 || S.rank >= n) { fprintf(stderr, "[hsk_mpi] bad HSK_MPI_RANK\n"); _exit(2); }
    if (const char *et = getenv("HSK_MPI_TIMEOUT_S")) S.timeout_s = atof(et);
    size_t slot_mb = 4;
    if (const char *em = getenv("HSK_MPI_SLOT_MB")) slot_mb = (size_t)atol(em);
    if (slot_mb < 1) slot_mb = 1;
    S.slot_bytes = slot_mb << 20;
    const size_t head = (sizeof(Header) + 63) / 64 * 64 + (size_t)n * sizeof(RankCell);
    const size_t head_pad = (head + 4095) / 4096 * 4096;
    S.map_bytes = head_pad + (size_t)n * S.slot_bytes;
    S.shm_name = std::string("/hsk_mpi_") + ek;
    int fd = -1;
    const double t0 = now_s();
    while (true) {   /* whoever comes first creates the object; a fresh object is zero-filled by the kernel */
        fd = shm_open(S.shm_name.c_str(), O_CREAT | O_RDWR, 0600);
        if (fd >= 0) break;
        if (now_s() - t0 > 30) { perror("[hsk_mpi] shm_open"); _exit(2); }
        sched_yield();
    }
    if (ftruncate(fd, (off_t)S.map_bytes) != 0) { perror("[hsk_mpi] ftruncate"); _exit(2); }
    void *p = mmap(nullptr, S.map_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { perror("[hsk_mpi] mmap"); _exit(2); }
    S.hdr = reinterpret_cast<Header *>(p);
    S.cells = reinterpret_cast<RankCell *>(reinterpret_cast<char *>(p) + (sizeof(Header) + 63) / 64 * 64);
    S.slots = reinterpret_cast<char *>(p) + head_pad;
    if (S.rank == 0) {
        S.hdr->nranks = (uint64_t)n;
        S.hdr->slot_bytes = S.slot_bytes;
        __atomic_thread_fence(__ATOMIC_RELEASE);
        S.hdr->magic = MAGIC;
    }
    wait_until([&] { return S.hdr->magic == MAGIC; }, "timeout waiting for rank 0 to create the session");
    if (S.hdr->nranks != (uint64_t)n || S.hdr->slot_bytes != S.slot_bytes) die("ranks disagree on HSK_MPI_SIZE / HSK_MPI_SLOT_MB");
    __atomic_add_fetch(&S.hdr->attached, 1, __ATOMIC_ACQ_REL);
    wait_until([&] { return S.hdr->attached >= (uint64_t)n; }, "timeout waiting for all ranks to attach");
    atexit([] { hsk_mpi_detach_at_exit(); });   /* a process that never calls MPI_Finalize still lets the object go */
}

inline char *slot_of(int p) { State &S = state(); return S.slots + (size_t)p * S.slot_bytes; }

/* ---- one collective = rounds of (publish window, read ranges) --------------------------------------------
 * Every rank contributes a byte array `data[0..len)` and a short descriptor; `plan(p, desc_p, len_p, want)` tells
 * which byte range [off, off+n) of rank p's array goes to which local address.  Rounds continue until the longest
 * contribution has been shown in full. */
struct Want { size_t off = 0, n = 0; char *dst = nullptr; };

inline void publish_round(const void *desc, size_t desc_len, const char *data, size_t len, size_t win_off)
{
    State &S = state();
    const uint64_t r = S.seq + 1;
    /* my slot may be rewritten once every rank has finished reading round r-1 */
    wait_until([&] { for (int p = 0; p < S.size; ++p) if (S.cells[p].consumed + 1 < r) return false; return true; },
               "timeout: a peer never finished reading the previous round");
    char *slot = slot_of(S.rank);
    if (desc_len) memcpy(slot, desc, desc_len);
    const size_t win = S.slot_bytes - DESC_MAX;
    if (win_off < len) memcpy(slot + DESC_MAX, data + win_off, len - win_off < win ? len - win_off : win);
    S.cells[S.rank].len = len;
    S.cells[S.rank].desc_len = desc_len;
    __atomic_thread_fence(__ATOMIC_RELEASE);
    S.cells[S.rank].published = r;
}

/* receive side of the outstanding MPI_Ialltoall.  Runs in MPI_Wait, or earlier when another collective is called while
 * the request is outstanding (the reference posts MPI_Ialltoall and then calls MPI_Barrier, kmerops.cpp:919 + 816): the
 * receive buffer belongs to the library until MPI_Wait, so filling it early is allowed, and every rank does so at the same
 * point of the common sequence of collectives. */
inline void complete_pending()
{
    State &S = state();
    if (!S.pending) return;
    const uint64_t r = S.seq + 1;
    for (int p = 0; p < S.size; ++p) {
        wait_until([&] { return S.cells[p].published >= r; }, "timeout: a peer never posted its MPI_Ialltoall");
        memcpy((char *)S.pend_recv + (size_t)p * S.pend_bytes, slot_of(p) + DESC_MAX + (size_t)S.rank * S.pend_bytes, S.pend_bytes);
    }
    __atomic_thread_fence(__ATOMIC_RELEASE);
    S.cells[S.rank].consumed = r;
    S.seq = r;
    S.pending = false;
}

template <typename Plan>
inline void collective(const void *desc, size_t desc_len, const void *data, size_t len, Plan plan)
{
    State &S = state();
    if (desc_len > DESC_MAX) die("descriptor too large");
    complete_pending();
    const size_t win = S.slot_bytes - DESC_MAX;
    std::vector<Want> want((size_t)S.size);
    size_t maxlen = 0;
    for (size_t round = 0;; ++round) {
        const size_t w0 = round * win;
        publish_round(round == 0 ? desc : nullptr, round == 0 ? desc_len : 0, reinterpret_cast<const char *>(data), len, w0);
        const uint64_t r = S.seq + 1;
        for (int p = 0; p < S.size; ++p) {
            wait_until([&] { return S.cells[p].published >= r; }, "timeout: a peer never reached this collective");
            if (round == 0) {
                const size_t lp = (size_t)S.cells[p].len;
                if (lp > maxlen) maxlen = lp;
                plan(p, slot_of(p), lp, want[(size_t)p]);
                if (want[(size_t)p].off + want[(size_t)p].n > lp) die("collective: peer sent fewer bytes than expected");
            }
            const Want &w = want[(size_t)p];
            const size_t a = w.off > w0 ? w.off : w0;
            const size_t b = (w.off + w.n) < (w0 + win) ? (w.off + w.n) : (w0 + win);
            if (a < b) memcpy(w.dst + (a - w.off), slot_of(p) + DESC_MAX + (a - w0), b - a);
        }
        __atomic_thread_fence(__ATOMIC_RELEASE);
        S.cells[S.rank].consumed = r;
        S.seq = r;
        if ((round + 1) * win >= maxlen) break;
    }
}

template <typename T>
inline void reduce_typed(T *acc, const T *x, size_t n, MPI_Op op)
{
    for (size_t i = 0; i < n; ++i) {
        if (op == MPI_SUM) acc[i] = (T)(acc[i] + x[i]);
        else if (op == MPI_MAX) acc[i] = x[i] > acc[i] ? x[i] : acc[i];
        else acc[i] = x[i] < acc[i] ? x[i] : acc[i];
    }
}

inline void reduce_into(void *acc, const void *x, size_t n, MPI_Datatype t, MPI_Op op)
{
    const int cls = type_class(t);
    const size_t sz = type_size(t);
    if (cls == 1 && sz == 4) reduce_typed((int32_t *)acc, (const int32_t *)x, n, op);
    else if (cls == 1 && sz == 8) reduce_typed((int64_t *)acc, (const int64_t *)x, n, op);
    else if (cls == 1 && sz == 1) reduce_typed((signed char *)acc, (const signed char *)x, n, op);
    else if (cls == 2 && sz == 4) reduce_typed((uint32_t *)acc, (const uint32_t *)x, n, op);
    else if (cls == 2 && sz == 8) reduce_typed((uint64_t *)acc, (const uint64_t *)x, n, op);
    else if (cls == 2 && sz == 1) reduce_typed((unsigned char *)acc, (const unsigned char *)x, n, op);
    else if (cls == 3 && sz == 4) reduce_typed((float *)acc, (const float *)x, n, op);
    else if (cls == 3 && sz == 8) reduce_typed((double *)acc, (const double *)x, n, op);
    else die("reduction on an unsupported datatype");
}

/* gathers `bytes` from every rank (rank order) into a temporary and folds them with `op` */
inline void reduce_all(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root /* -1: everybody */, bool exclusive_prefix)
{
    State &S = state();
    const size_t bytes = (size_t)n * type_size(t);
    std::vector<char> mine;
    const void *src = s;
    if (s == MPI_IN_PLACE) { mine.assign((const char *)r, (const char *)r + bytes); src = mine.data(); }
    const bool receiver = (root < 0 || root == S.rank);
    std::vector<char> all(receiver ? bytes * (size_t)S.size : 0);
    collective(nullptr, 0, src, bytes, [&](int p, const char *, size_t, Want &w) {
        if (receiver) { w.off = 0; w.n = bytes; w.dst = all.data() + (size_t)p * bytes; }
    });
    if (!receiver) return;
    if (exclusive_prefix) {
        if (S.rank == 0) return;   /* undefined on rank 0: left untouched */
        memcpy(r, all.data(), bytes);
        for (int p = 1; p < S.rank; ++p) reduce_into(r, all.data() + (size_t)p * bytes, (size_t)n, t, op);
        return;
    }
    memcpy(r, all.data(), bytes);
    for (int p = 1; p < S.size; ++p) reduce_into(r, all.data() + (size_t)p * bytes, (size_t)n, t, op);
}

inline void local_copy(const void *src, void *dst, size_t n)
{
    if (src != MPI_IN_PLACE && src != dst && n) memcpy(dst, src, n);
}

inline void finalize()
{
    State &S = state();
    if (!S.hdr) return;
    const uint64_t left = __atomic_add_fetch(&S.hdr->detached, 1, __ATOMIC_ACQ_REL);
    if (left >= (uint64_t)S.size) shm_unlink(S.shm_name.c_str());   /* the last one out removes the object */
    munmap((void *)S.hdr, S.map_bytes);
    S.hdr = nullptr;
    S.size = 1; S.rank = 0;
}

} // namespace hsk_mpi

/* ---- the MPI functions ---------------------------------------------------------------------------------- */

inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; hsk_mpi::ensure_init(); return MPI_SUCCESS; }
inline int MPI_Initialized(int *flag) { *flag = 1; return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm c, int *r) { hsk_mpi::ensure_init(); *r = c == MPI_COMM_SELF ? 0 : hsk_mpi::state().rank; return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm c, int *s) { hsk_mpi::ensure_init(); *s = c == MPI_COMM_SELF ? 1 : hsk_mpi::state().size; return MPI_SUCCESS; }
inline double MPI_Wtime(void) { return hsk_mpi::now_s(); }

inline int MPI_Barrier(MPI_Comm c)
{
    hsk_mpi::ensure_init();
    if (c == MPI_COMM_SELF || hsk_mpi::state().size == 1) return MPI_SUCCESS;
    hsk_mpi::collective(nullptr, 0, nullptr, 0, [](int, const char *, size_t, hsk_mpi::Want &) {});
    return MPI_SUCCESS;
}

inline int MPI_Finalize(void)
{
    hsk_mpi::ensure_init();
    if (hsk_mpi::state().size > 1) { MPI_Barrier(MPI_COMM_WORLD); hsk_mpi::finalize(); }
    return MPI_SUCCESS;
}

inline int MPI_Abort(MPI_Comm c, int code)
{
    (void)c;
    hsk_mpi::State &S = hsk_mpi::state();
    if (S.hdr) S.hdr->abort_code = (uint64_t)(code ? code : 1);   /* peers leave their wait loops */
    exit(code);
    return MPI_SUCCESS;
}

inline int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm c)
{
    hsk_mpi::ensure_init();
    hsk_mpi::State &S = hsk_mpi::state();
    if (c == MPI_COMM_SELF || S.size == 1) return MPI_SUCCESS;
    const size_t bytes = (size_t)n * hsk_mpi::type_size(t);
    const bool is_root = S.rank == root;
    hsk_mpi::collective(nullptr, 0, is_root ? buf : nullptr, is_root ? bytes : 0, [&](int p, const char *, size_t, hsk_mpi::Want &w) {
        if (p == root && !is_root) { w.off = 0; w.n = bytes; w.dst = (char *)buf; }
    });
    return MPI_SUCCESS;
}

inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{
    hsk_mpi::ensure_init();
    if (c == MPI_COMM_SELF || hsk_mpi::state().size == 1) { hsk_mpi::local_copy(s, r, (size_t)n * hsk_mpi::type_size(t)); return MPI_SUCCESS; }
    hsk_mpi::reduce_all(s, r, n, t, op, root, false);
    return MPI_SUCCESS;
}

inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
    hsk_mpi::ensure_init();
    if (c == MPI_COMM_SELF || hsk_mpi::state().size == 1) { hsk_mpi::local_copy(s, r, (size_t)n * hsk_mpi::type_size(t)); return MPI_SUCCESS; }
    hsk_mpi::reduce_all(s, r, n, t, op, -1, false);
    return MPI_SUCCESS;
}

/* rank 0's receive buffer is undefined after MPI_Exscan; it is left untouched */
inline int MPI_Exscan(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
    hsk_mpi::ensure_init();
    if (c == MPI_COMM_SELF || hsk_mpi::state().size == 1) return MPI_SUCCESS;
    hsk_mpi::reduce_all(s, r, n, t, op, -1, true);
    return MPI_SUCCESS;
}

inline int MPI_Gatherv(const void *s, int sn, MPI_Datatype st, void *r, const int *rn, const int *displs, MPI_Datatype rt,
                       int root, MPI_Comm c)
{
    hsk_mpi::ensure_init();
    hsk_mpi::State &S = hsk_mpi::state();
    const size_t sbytes = (size_t)sn * hsk_mpi::type_size(st);
    if (c == MPI_COMM_SELF || S.size == 1) {
        hsk_mpi::local_copy(s, (char *)r + (size_t)displs[0] * hsk_mpi::type_size(rt), sbytes);
        return MPI_SUCCESS;
    }
    const bool is_root = S.rank == root;
    hsk_mpi::collective(nullptr, 0, s, sbytes, [&](int p, const char *, size_t, hsk_mpi::Want &w) {
        if (is_root) { w.off = 0; w.n = (size_t)rn[p] * hsk_mpi::type_size(rt); w.dst = (char *)r + (size_t)displs[p] * hsk_mpi::type_size(rt); }
    });
    return MPI_SUCCESS;
}

inline int MPI_Gather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c)
{
    hsk_mpi::ensure_init();
    hsk_mpi::State &S = hsk_mpi::state();
    const size_t sbytes = (size_t)sn * hsk_mpi::type_size(st);
    if (c == MPI_COMM_SELF || S.size == 1) { hsk_mpi::local_copy(s, r, sbytes); return MPI_SUCCESS; }
    const bool is_root = S.rank == root;
    const size_t rbytes = (size_t)rn * hsk_mpi::type_size(rt);
    hsk_mpi::collective(nullptr, 0, s, sbytes, [&](int p, const char *, size_t, hsk_mpi::Want &w) {
        if (is_root) { w.off = 0; w.n = rbytes; w.dst = (char *)r + (size_t)p * rbytes; }
    });
    return MPI_SUCCESS;
}

inline int MPI_Allgather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c)
{
    hsk_mpi::ensure_init();
    hsk_mpi::State &S = hsk_mpi::state();
    const size_t sbytes = (size_t)sn * hsk_mpi::type_size(st);
    if (c == MPI_COMM_SELF || S.size == 1) { hsk_mpi::local_copy(s, r, sbytes); return MPI_SUCCESS; }
    const size_t rbytes = (size_t)rn * hsk_mpi::type_size(rt);
    std::vector<char> mine;
    const void *src = s;
    if (s == MPI_IN_PLACE) { mine.assign((char *)r + (size_t)S.rank * rbytes, (char *)r + (size_t)(S.rank + 1) * rbytes); src = mine.data(); }
    hsk_mpi::collective(nullptr, 0, src, s == MPI_IN_PLACE ? rbytes : sbytes, [&](int p, const char *, size_t, hsk_mpi::Want &w) {
        w.off = 0; w.n = rbytes; w.dst = (char *)r + (size_t)p * rbytes;
    });
    return MPI_SUCCESS;
}

inline int MPI_Scatterv(const void *s, const int *sn, const int *displs, MPI_Datatype st, void *r, int rn, MPI_Datatype rt,
                        int root, MPI_Comm c)
{
    hsk_mpi::ensure_init();
    hsk_mpi::State &S = hsk_mpi::state();
    const size_t rbytes = (size_t)rn * hsk_mpi::type_size(rt);
    if (c == MPI_COMM_SELF || S.size == 1) {
        hsk_mpi::local_copy((const char *)s + (size_t)displs[0] * hsk_mpi::type_size(st), r, rbytes);
        return MPI_SUCCESS;
    }
    const bool is_root = S.rank == root;
    /* the root shows its whole send array (up to the end of the last segment) and, as the descriptor, the byte
     * offset of every rank's segment */
    std::vector<uint64_t> offs;
    size_t total = 0;
    if (is_root) {
        offs.resize((size_t)S.size);
        for (int p = 0; p < S.size; ++p) {
            offs[(size_t)p] = (uint64_t)displs[p] * hsk_mpi::type_size(st);
            const size_t end = (size_t)offs[(size_t)p] + (size_t)sn[p] * hsk_mpi::type_size(st);
            if (end > total) total = end;
        }
        if (offs.size() * 8 > hsk_mpi::DESC_MAX) hsk_mpi::die("MPI_Scatterv: too many ranks for the descriptor area");
    }
    hsk_mpi::collective(is_root ? offs.data() : nullptr, is_root ? offs.size() * 8 : 0, is_root ? s : nullptr, is_root ? total : 0,
                        [&](int p, const char *desc, size_t, hsk_mpi::Want &w) {
        if (p == root) { uint64_t o; memcpy(&o, desc + (size_t)S.rank * 8, 8); w.off = (size_t)o; w.n = rbytes; w.dst = (char *)r; }
    });
    return MPI_SUCCESS;
}

inline int MPI_Alltoallv(const void *s, const int *sn, const int *sd, MPI_Datatype st, void *r, const int *rn, const int *rd,
                         MPI_Datatype rt, MPI_Comm c)
{
    hsk_mpi::ensure_init();
    hsk_mpi::State &S = hsk_mpi::state();
    if (c == MPI_COMM_SELF || S.size == 1) {
        hsk_mpi::local_copy((const char *)s + (size_t)sd[0] * hsk_mpi::type_size(st), (char *)r + (size_t)rd[0] * hsk_mpi::type_size(rt),
                            (size_t)sn[0] * hsk_mpi::type_size(st));
        return MPI_SUCCESS;
    }
    std::vector<uint64_t> offs((size_t)S.size);
    size_t total = 0;
    for (int p = 0; p < S.size; ++p) {
        offs[(size_t)p] = (uint64_t)sd[p] * hsk_mpi::type_size(st);
        const size_t end = (size_t)offs[(size_t)p] + (size_t)sn[p] * hsk_mpi::type_size(st);
        if (end > total) total = end;
    }
    if (offs.size() * 8 > hsk_mpi::DESC_MAX) hsk_mpi::die("MPI_Alltoallv: too many ranks for the descriptor area");
    hsk_mpi::collective(offs.data(), offs.size() * 8, s, total, [&](int p, const char *desc, size_t, hsk_mpi::Want &w) {
        uint64_t o;
        memcpy(&o, desc + (size_t)S.rank * 8, 8);
        w.off = (size_t)o; w.n = (size_t)rn[p] * hsk_mpi::type_size(rt); w.dst = (char *)r + (size_t)rd[p] * hsk_mpi::type_size(rt);
    });
    return MPI_SUCCESS;
}

inline int MPI_Alltoall(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c)
{
    hsk_mpi::ensure_init();
    hsk_mpi::State &S = hsk_mpi::state();
    const size_t sbytes = (size_t)sn * hsk_mpi::type_size(st), rbytes = (size_t)rn * hsk_mpi::type_size(rt);
    if (c == MPI_COMM_SELF || S.size == 1) { hsk_mpi::local_copy(s, r, sbytes); return MPI_SUCCESS; }
    hsk_mpi::collective(nullptr, 0, s, sbytes * (size_t)S.size, [&](int p, const char *, size_t, hsk_mpi::Want &w) {
        w.off = (size_t)S.rank * sbytes; w.n = rbytes; w.dst = (char *)r + (size_t)p * rbytes;
    });
    return MPI_SUCCESS;
}

/* Non-blocking all-to-all: the send buffer is published at once (it may be reused by the caller right away, as far as
 * this implementation is concerned it is copied), the receive side runs in MPI_Wait.  The whole exchange must fit one
 * window (size * bytes-per-destination <= slot - 4 KiB): the reference sends 80 000 bytes per destination. */
inline int MPI_Ialltoall(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c, MPI_Request *req)
{
    hsk_mpi::ensure_init();
    hsk_mpi::State &S = hsk_mpi::state();
    const size_t sbytes = (size_t)sn * hsk_mpi::type_size(st);
    *req = 1;
    if (c == MPI_COMM_SELF || S.size == 1) { hsk_mpi::local_copy(s, r, sbytes); return MPI_SUCCESS; }
    hsk_mpi::complete_pending();   /* one request at a time: a second one completes the first */
    if (sbytes * (size_t)S.size > S.slot_bytes - hsk_mpi::DESC_MAX) hsk_mpi::die("MPI_Ialltoall: raise HSK_MPI_SLOT_MB");
    (void)rn; (void)rt;
    hsk_mpi::publish_round(nullptr, 0, (const char *)s, sbytes * (size_t)S.size, 0);
    S.pending = true; S.pend_recv = r; S.pend_bytes = sbytes;
    return MPI_SUCCESS;
}

inline int MPI_Wait(MPI_Request *req, MPI_Status *st)
{
    (void)st;
    hsk_mpi::complete_pending();
    if (req) *req = 0;
    return MPI_SUCCESS;
}

inline int MPI_Type_contiguous(int n, MPI_Datatype old, MPI_Datatype *nw) { *nw = HSK_MPI_TYPE(0, n * (int)hsk_mpi::type_size(old)); return MPI_SUCCESS; }
inline int MPI_Type_commit(MPI_Datatype *t) { (void)t; return MPI_SUCCESS; }
inline int MPI_Type_free(MPI_Datatype *t) { (void)t; return MPI_SUCCESS; }

/* only the size query is used by the reference (fastaindex.cpp:223-224, to clamp a read range) */
inline int MPI_File_open(MPI_Comm c, const char *fn, int mode, MPI_Info info, MPI_File *fh)
{
    (void)c; (void)mode; (void)info;
    *fh = ::open(fn, O_RDONLY);
    return MPI_SUCCESS;
}
inline int MPI_File_get_size(MPI_File fh, MPI_Offset *sz)
{
    struct stat st;
    if (fh >= 0 && fstat(fh, &st) == 0) *sz = (MPI_Offset)st.st_size; else *sz = (MPI_Offset)1 << 62;
    return MPI_SUCCESS;
}
inline int MPI_File_close(MPI_File *fh) { if (fh && *fh >= 0) ::close(*fh); return MPI_SUCCESS; }

#endif /* HSK_MPI_SHIM_H_ */
