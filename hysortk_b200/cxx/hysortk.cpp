// The hysortk C++ API (reference include/hysortk.hpp:8-18) over the CUDA engine's C ABI
// (include/hsk_capi.h).  Same four functions, same argument meaning, same collectivity over
// `comm`, same output text; kmer_count's body is the GPU path.
//
//   read_dna_buffer      reference src/hysortk.cpp:18-33 + src/fastaindex.cpp (.fai parse :20-28,
//                        contiguous partition by bases :52-100, per-record line stripping :269-286)
//   kmer_count           reference src/hysortk.cpp:36-95  ->  hsk_create / hsk_count
//   print_kmer_histogram reference src/hysortk.cpp:98-136
//   write_output_file    reference src/hysortk.cpp:138-164
#include "hysortk.hpp"
#include "compiletime.h"
#include "hsk_capi.h"

#include <chrono>
#include <omp.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cstring>
#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <memory>
#include <mutex>
#include <new>
#include <utility>
#include <vector>
#include <type_traits>

namespace hysortk {

namespace {

struct FaiRecord { size_t len, pos, bases, width; };   /* .fai columns 2-5: length, offset, bases per line, bytes per line */

int local_device_for(int rank)
{
    for (const char *v : {"LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID"}) {
        const char *e = std::getenv(v);
        if (e && *e) return std::atoi(e);
    }
    const char *n = std::getenv("HSK_GPUS_PER_NODE");
    int per = n ? std::atoi(n) : 8;
    return per > 0 ? rank % per : 0;
}

/* One engine context per process, re-created when the communicator shape changes.  It is NOT torn down by a static
 * destructor: at that point MPI is finalised and the CUDA runtime may be gone, and with several ranks hsk_destroy is
 * collective.  hysortk::release_gpu_engine() frees it explicitly; otherwise it lives until the process exits. */
struct Engine {
    hsk_ctx *ctx = nullptr;
    int rank = -1, nranks = -1;
};
Engine g_engine;

/* Builds the KmerListS while the result is still arriving (sink of hsk_count_stream).  The entries are constructed in
 * place in memory obtained from std::allocator<KmerListEntryS>: the host threads of the engine deliver disjoint parts of
 * the result concurrently and each fills its part; the finished array is then handed to a std::vector without the
 * value-initialisation + copy that resize() + assignment would cost (the list has millions of entries; first touch of
 * its pages is the expensive part and is spread over the threads).  The array is allocated when the first part arrives,
 * from the engine's estimate of the total; should a part not fit after all, it is left out and the list is rebuilt
 * from the complete result after the call (complete()). */
struct ListBuilder {
    static constexpr int NW = TKmer::NBYTES / 8;
    std::allocator<KmerListEntryS> alloc;
    KmerListEntryS *base = nullptr;
    size_t cap = 0;
    std::mutex mu;
    std::vector<std::pair<uint64_t, uint64_t>> done;   /* parts constructed so far: (first, n) */
    bool overflow = false;
    int nthreads = 1;   /* the caller's OpenMP team size (complete() only) */
    double fill_seconds = 0.0;
    std::string error;

    ~ListBuilder() { release(); }
    void destroy_entries()
    {
#if EXTENSION == 1
        for (auto& r : done)
            for (uint64_t i = r.first; i < r.first + r.second; ++i) base[i].~KmerListEntryS();
#endif
        done.clear();
    }
    void release()
    {
        if (!base) return;
        destroy_entries();
        alloc.deallocate(base, cap);
        base = nullptr; cap = 0;
    }
    void allocate(size_t want)
    {
        cap = want + 64;
        base = alloc.allocate(cap);
#ifdef MADV_HUGEPAGE
        /* millions of entries: with 4 KiB pages the first touch of the array costs more than filling it; ask for
         * transparent huge pages on the 2 MiB-aligned interior (a no-op where they are disabled) */
        const uintptr_t huge = uintptr_t(2) << 20;
        const uintptr_t a0 = (reinterpret_cast<uintptr_t>(base) + huge - 1) & ~(huge - 1);
        const uintptr_t a1 = (reinterpret_cast<uintptr_t>(base) + cap * sizeof(KmerListEntryS)) & ~(huge - 1);
        if (a1 > a0) (void)::madvise(reinterpret_cast<void *>(a0), a1 - a0, MADV_HUGEPAGE);
#endif
    }
    /* entries [first, first + n) from the result arrays */
    void fill(const hsk_result *v, uint64_t first, uint64_t n, uint64_t occ_end)
    {
        const uint64_t *words = v->kmer_words;
        const uint32_t *cnt = v->cnt;
#if EXTENSION == 1
        for (uint64_t i = first; i < first + n; ++i) {
            KmerListEntryS *e = new (base + i) KmerListEntryS();
            e->kmer = TKmer(static_cast<const void *>(words + i * NW));
            e->cnt = cnt[i];
            const uint64_t o0 = v->occ_off[i], o1 = (i + 1 < first + n) ? v->occ_off[i + 1] : occ_end;
            e->pos.assign(v->pos + o0, v->pos + o1);
            e->rid.assign(v->rid + o0, v->rid + o1);
        }
#else
        (void)occ_end;
        static_assert(sizeof(KmerListEntryS) == 8 * (NW + 1), "entry layout: k-mer words followed by the 64-bit count");
        uint64_t *raw = reinterpret_cast<uint64_t *>(base);
        for (uint64_t i = first; i < first + n; ++i) {
            for (int l = 0; l < NW; ++l) raw[i * (NW + 1) + l] = words[i * NW + l];
            raw[i * (NW + 1) + NW] = cnt[i];
        }
#endif
    }
    int take(const hsk_result *v, uint64_t first, uint64_t n, uint64_t first_occ, uint64_t n_occ, uint64_t hint)
    {
        const auto t_in = std::chrono::steady_clock::now();
        try {
            {
                std::lock_guard<std::mutex> l(mu);
                if (!base) {
                    size_t want = std::max<size_t>(first + n, hint);
                    if (const char *t = std::getenv("HSK_TEST_LIST_HINT_DIV")) want = std::max<size_t>(first + n, want / std::max(1, std::atoi(t)));   /* tests: force the rebuild path */
                    allocate(want);
                }
                if (first + n > cap) { overflow = true; return 0; }
                if (n) done.emplace_back(first, n);
            }
            fill(v, first, n, first_occ + n_occ);
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_in).count();
            std::lock_guard<std::mutex> l(mu);
            fill_seconds += dt;
            return 0;
        } catch (const std::exception& e) {
            std::lock_guard<std::mutex> l(mu);
            error = e.what();
            return 1;
        }
    }
    static int sink(void *user, const hsk_result *v, uint64_t first, uint64_t n, uint64_t first_occ, uint64_t n_occ, uint64_t hint)
    {
        return static_cast<ListBuilder *>(user)->take(v, first, n, first_occ, n_occ, hint);
    }
    /* after the call: `res` is the complete result.  Nothing to do when the parts covered it; otherwise (the estimate
     * fell short, or the engine delivered nothing) the list is built from `res` by the caller's threads. */
    void complete(const hsk_result& res)
    {
        uint64_t covered = 0;
        for (auto& r : done) covered += r.second;
        if (!overflow && covered == res.n_kept) return;
        release();
        overflow = false;
        if (res.n_kept == 0) return;
        allocate(res.n_kept);
        const uint64_t n = res.n_kept, block = 16384;
        #pragma omp parallel for schedule(dynamic) num_threads(nthreads)
        for (uint64_t b = 0; b < (n + block - 1) / block; ++b) {
            const uint64_t first = b * block, cnt = std::min(block, n - first);
#if EXTENSION == 1
            fill(&res, first, cnt, res.occ_off[first + cnt]);   /* the complete result has the closing offset of every entry */
#else
            fill(&res, first, cnt, 0);
#endif
        }
        done.emplace_back(0, n);
    }
    /* the finished array becomes the storage of a std::vector */
    std::unique_ptr<KmerListS> finish(uint64_t n)
    {
        auto list = std::make_unique<KmerListS>();
        if (!base || n == 0) return list;
#if defined(__GLIBCXX__) || defined(_LIBCPP_VERSION)
        /* libstdc++ / libc++: a vector with the default allocator is three pointers (begin, end, end of storage) */
        static_assert(sizeof(KmerListS) == 3 * sizeof(void *), "std::vector layout");
        KmerListEntryS *triple[3] = {base, base + n, base + cap};
        std::memcpy(static_cast<void *>(list.get()), triple, sizeof(triple));
        if (list->data() != base || list->size() != n || list->capacity() != cap) {
            /* not the layout we took it for: undo and take the portable way */
            KmerListEntryS *none[3] = {nullptr, nullptr, nullptr};
            std::memcpy(static_cast<void *>(list.get()), none, sizeof(none));
            list->reserve(n);
            for (uint64_t i = 0; i < n; ++i) list->push_back(std::move(base[i]));
            release();
            return list;
        }
        base = nullptr; cap = 0;
        done.clear();
#else
        list->reserve(n);
        for (uint64_t i = 0; i < n; ++i) list->push_back(std::move(base[i]));
        release();
#endif
        return list;
    }
};

hsk_ctx *engine_for(MPI_Comm comm)
{
    int rank, nranks;
    MPI_Comm_rank(comm, &rank);
    MPI_Comm_size(comm, &nranks);
    if (g_engine.ctx && g_engine.rank == rank && g_engine.nranks == nranks) return g_engine.ctx;
    if (g_engine.ctx) { hsk_destroy(g_engine.ctx); g_engine.ctx = nullptr; }
    unsigned char id[HSK_NCCL_ID_BYTES] = {0};
    if (nranks > 1) {
        if (rank == 0 && hsk_get_unique_id(id)) throw std::runtime_error(hsk_last_error());
        MPI_Bcast(id, HSK_NCCL_ID_BYTES, MPI_BYTE, 0, comm);
    }
    hsk_config cfg{};
    cfg.k = KMER_SIZE; cfg.m = MINIMIZER_SIZE; cfg.lower = LOWER_KMER_FREQ; cfg.upper = UPPER_KMER_FREQ;
    cfg.ext = EXTENSION;
    cfg.device = nranks > 1 ? local_device_for(rank) : (std::getenv("HSK_DEVICE") ? std::atoi(std::getenv("HSK_DEVICE")) : 0);
    cfg.rank = rank; cfg.nranks = nranks;
    cfg.nccl_id = nranks > 1 ? id : nullptr;
    cfg.buckets_per_rank = std::getenv("HSK_BUCKETS_PER_RANK") ? std::atoi(std::getenv("HSK_BUCKETS_PER_RANK")) : 0;
    cfg.batch_kmers = std::getenv("HSK_BATCH_KMERS") ? std::strtoull(std::getenv("HSK_BATCH_KMERS"), nullptr, 10) : 0;
    cfg.stream = nullptr;
    if (hsk_create(&g_engine.ctx, &cfg)) throw std::runtime_error(hsk_last_error());
    g_engine.rank = rank; g_engine.nranks = nranks;
    return g_engine.ctx;
}

} // namespace

std::shared_ptr<DnaBuffer> read_dna_buffer(const std::string& fasta_fname, MPI_Comm comm)
{
    int rank, nranks;
    MPI_Comm_rank(comm, &rank);
    MPI_Comm_size(comm, &nranks);
    auto t0 = std::chrono::steady_clock::now();

    /* every rank parses the index itself (the reference parses on rank 0 and scatters) */
    std::vector<FaiRecord> rec;
    {
        std::ifstream fai(fasta_fname + ".fai");
        if (!fai) throw std::runtime_error("cannot open FASTA index " + fasta_fname + ".fai");
        std::string line, name;
        while (std::getline(fai, line)) {
            if (line.empty()) continue;
            FaiRecord r{};
            std::istringstream ls(line);
            ls >> name >> r.len >> r.pos >> r.bases;
            if (!ls) throw std::runtime_error("malformed line in " + fasta_fname + ".fai: " + line);
            if (!(ls >> r.width) || r.width <= r.bases) r.width = r.bases + 1;   /* the reference assumes 1-byte line ends (fastaindex.cpp:285) */
            rec.push_back(r);
        }
    }
    /* contiguous partition balanced by bases, same greedy rule as the reference (fastaindex.cpp:52-100) */
    size_t totbases = 0;
    for (auto& r : rec) totbases += r.len;
    const double avg = static_cast<double>(totbases) / nranks;
    std::vector<size_t> first(nranks + 1, rec.size());
    size_t id = 0;
    for (int p = 0; p < nranks - 1; ++p) {
        first[p] = id;
        size_t sofar = 0;
        if (id < rec.size()) {
            do { sofar += rec[id].len; ++id; } while (id < rec.size() && sofar + rec[id].len < avg);
        }
    }
    first[nranks - 1] = id;
    first[nranks] = rec.size();
    const size_t lo = first[rank], hi = first[rank + 1];

    /* One read() of this rank's byte range, then every record is 2-bit encoded straight into its place in the buffer
     * by the host threads (the reference encodes record by record on one thread per rank, fastaindex.cpp:269-286 +
     * dnaseq.cpp:9-31).  The bytes are exactly those of DnaSeq::compress. */
    std::vector<size_t> lens, byteoff;
    lens.reserve(hi - lo);
    byteoff.reserve(hi - lo + 1);
    size_t bufsize = 0;
    for (size_t i = lo; i < hi; ++i) {
        lens.push_back(rec[i].len);
        byteoff.push_back(bufsize);
        bufsize += DnaSeq::bytesneeded(rec[i].len);
    }
    byteoff.push_back(bufsize);
    uint8_t *buf = new uint8_t[bufsize ? bufsize : 1];
    if (hi > lo) {
        const size_t start = rec[lo].pos;
        const FaiRecord& last = rec[hi - 1];
        /* bytes of a record in the file: its bases + the line terminators between them */
        auto extent = [](const FaiRecord& r) { return r.len + (r.bases ? (r.len - (r.len ? 1 : 0)) / r.bases * (r.width - r.bases) : 0); };
        const size_t end = last.pos + extent(last);
        /* the rank's byte range: mapped read-only (the encoder threads read the page cache directly); a plain
         * read() into a string is the fallback */
        const char *text = nullptr;
        std::string chunk;
        void *map = MAP_FAILED;
        size_t map_len = 0, map_skew = 0;
        const int fd = ::open(fasta_fname.c_str(), O_RDONLY);
        if (fd < 0) { delete[] buf; throw std::runtime_error("cannot open FASTA file " + fasta_fname); }
        struct stat st;
        if (::fstat(fd, &st) != 0) { ::close(fd); delete[] buf; throw std::runtime_error("cannot stat FASTA file " + fasta_fname); }
        /* a stale or truncated index must not send the encoder threads past the end of the file */
        for (size_t i = lo; i < hi; ++i) {
            if (rec[i].pos + extent(rec[i]) > static_cast<size_t>(st.st_size)) {
                ::close(fd); delete[] buf;
                throw std::runtime_error("FASTA index " + fasta_fname + ".fai does not match the file: record " + std::to_string(i) +
                                         " ends beyond its " + std::to_string(st.st_size) + " bytes");
            }
        }
        if (end > start) {
            const size_t page = static_cast<size_t>(::sysconf(_SC_PAGESIZE));
            map_skew = start % page;
            map_len = end - start + map_skew;
            map = ::mmap(nullptr, map_len, PROT_READ, MAP_PRIVATE, fd, static_cast<off_t>(start - map_skew));
            if (map != MAP_FAILED) {
                ::madvise(map, map_len, MADV_SEQUENTIAL);
                text = static_cast<const char *>(map) + map_skew;
            }
        }
        if (!text) {
            chunk.assign(end - start, '\n');
            size_t got = 0;
            while (got < chunk.size()) {
                const ssize_t r = ::pread(fd, &chunk[got], chunk.size() - got, static_cast<off_t>(start + got));
                if (r <= 0) { ::close(fd); delete[] buf; throw std::runtime_error("short read from FASTA file " + fasta_fname); }
                got += static_cast<size_t>(r);
            }
            text = chunk.data();
        }
        ::close(fd);
        const long nrec = static_cast<long>(hi - lo);
        #pragma omp parallel for schedule(dynamic, 64)
        for (long ii = 0; ii < nrec; ++ii) {
            const FaiRecord& r = rec[lo + ii];
            const char *src = text + (r.pos - start);
            uint8_t *dst = buf + byteoff[ii];
            const size_t width = r.bases ? r.bases : r.len;   /* characters per FASTA line */
            const size_t eol = r.bases ? r.width - r.bases : 0; /* bytes of the line terminator (2 for CRLF files) */
            size_t col = 0;
            for (size_t p = 0; p < r.len;) {
                unsigned byte = 0;
                for (int j = 0; j < 4 && p < r.len; ++j, ++p) {
                    if (col == width) { src += eol; col = 0; }   /* the line end after every `bases` characters */
                    byte |= (static_cast<unsigned>(DnaSeq::getcharcode(*src++)) << (6 - 2 * j)) & 0xFFu;
                    ++col;
                }
                *dst++ = static_cast<uint8_t>(byte);
            }
        }
        if (map != MAP_FAILED) ::munmap(map, map_len);
    }
    auto dna = std::make_shared<DnaBuffer>(bufsize, lens.size(), buf, lens.data());   /* adopts buf */
    /* page-lock the bytes: kmer_count then sends them to the GPU from where they are instead of staging
     * them.  Failure is harmless (no GPU in this process, locked-memory limit): the buffer stays pageable. */
    {
        const char *pin = std::getenv("HSK_PIN_DNABUFFER");
        if (bufsize >= (size_t(1) << 14) && !(pin && *pin == '0')) {
            const int device = nranks > 1 ? local_device_for(rank) : (std::getenv("HSK_DEVICE") ? std::atoi(std::getenv("HSK_DEVICE")) : 0);
            if (hsk_host_register(buf, bufsize, device) == 0) dna->set_release_hook([](uint8_t *p) { (void)hsk_host_unregister(p); });
        }
    }
    MPI_Barrier(comm);
#if LOG_LEVEL >= 1
    if (rank == 0) {
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::cout << "reading and 2-bit encoding fasta sequences: " << s << " s" << std::endl;
    }
#else
    (void)t0;
#endif
    return dna;
}

void release_gpu_engine()
{
    if (g_engine.ctx) { hsk_destroy(g_engine.ctx); g_engine.ctx = nullptr; g_engine.rank = g_engine.nranks = -1; }
}

std::unique_ptr<KmerListS> kmer_count(const DnaBuffer& mydna, MPI_Comm comm)
{
    int rank, nranks;
    MPI_Comm_rank(comm, &rank);
    MPI_Comm_size(comm, &nranks);
    hsk_ctx *ctx = engine_for(comm);

#if LOG_LEVEL >= 1
    MPI_Barrier(comm);
    auto t0 = std::chrono::steady_clock::now();
#endif
    /* global id of this rank's first read (reference kmerops.cpp:65-70) */
    int numreads = static_cast<int>(mydna.size());
    int readoffset = 0;
    MPI_Exscan(&numreads, &readoffset, 1, MPI_INT, MPI_SUM, comm);
    if (rank == 0) readoffset = 0;

    /* read lengths as the C ABI takes them.  Kept between calls (a fresh 8 MB array per million reads would be first-touched
     * every time) and filled by this thread alone: an OpenMP team here would keep spinning on the cores that the engine's
     * host threads need a moment later */
    const size_t n = mydna.size();
    static thread_local std::vector<uint64_t> lens;
    lens.resize(n);
    for (size_t i = 0; i < n; ++i) lens[i] = mydna[i].size();
    const uint8_t *bytes = n ? mydna.getbufoffset(0) : nullptr;
    const size_t nbytes = n ? mydna.getrangebufsize(0, n) : 0;

    /* the DnaBuffer bytes go to the GPU as they are (pageable memory: staged by the engine's host threads); the entries
     * are built from the parts of the result as they arrive, while the GPU is still counting */
    hsk_result res;
    ListBuilder builder;
    builder.nthreads = std::max(1, omp_get_max_threads());
    if (hsk_count_stream(ctx, bytes, nbytes, lens.data(), n, readoffset, &ListBuilder::sink, &builder, &res))
        throw std::runtime_error(builder.error.empty() ? std::string(hsk_last_error()) : "kmer_count: " + builder.error);
    builder.complete(res);
    if (const char *tr = std::getenv("HSK_TRACE"); tr && *tr == '1')
        std::cerr << "[hsk trace] KmerListS: " << res.n_kept << " entries in " << builder.done.size() << " parts, "
                  << builder.fill_seconds * 1e3 << " ms (summed over the threads) in the sink" << std::endl;
    auto list = builder.finish(res.n_kept);

#if LOG_LEVEL >= 1
    MPI_Barrier(comm);
    if (rank == 0) {
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::cout << "Overall kmer counting (Excluding I/O): " << s << " s" << std::endl;
#if LOG_LEVEL >= 2
        const hsk_stats& st = res.stats;
        std::cout << "  device ms: h2d " << st.ms_h2d << " extract " << st.ms_extract << " exchange " << st.ms_exchange
                  << " expand " << st.ms_expand << " sort " << st.ms_sort << " count " << st.ms_count << " d2h " << st.ms_d2h
                  << " | k-mers " << st.n_kmers_local << " supermers " << st.n_supermers << " batches " << st.n_batches
                  << std::endl;
#endif
    }
#endif
    return list;
}

void print_kmer_histogram(const KmerListS& kmerlist, MPI_Comm comm)
{
    /* counts never exceed UPPER_KMER_FREQ, so the histogram has a fixed size and the reference's
     * max-allreduce (hysortk.cpp:102-104) is not needed */
    std::vector<unsigned long long> histo(UPPER_KMER_FREQ + 1, 0);
    const size_t nent = kmerlist.size();
    #pragma omp parallel if (nent > (size_t(1) << 16))
    {
        std::vector<unsigned long long> mine(UPPER_KMER_FREQ + 1, 0);   /* the list has millions of entries: one pass per thread */
        #pragma omp for schedule(static) nowait
        for (size_t i = 0; i < nent; ++i)
            if (kmerlist[i].cnt <= UPPER_KMER_FREQ) mine[kmerlist[i].cnt]++;
        #pragma omp critical
        for (size_t c = 0; c < mine.size(); ++c) histo[c] += mine[c];
    }
    MPI_Allreduce(MPI_IN_PLACE, histo.data(), static_cast<int>(histo.size()), MPI_UNSIGNED_LONG_LONG, MPI_SUM, comm);
    int rank;
    MPI_Comm_rank(comm, &rank);
    if (rank == 0) {
        std::ostringstream ss;
        ss << "#count\tnumkmers\n";
        for (size_t i = 1; i < histo.size(); ++i)
            if (histo[i] > 0) ss << i << "\t" << histo[i] << "\n";
        ss << "\n";
        std::cout << ss.str() << std::flush;
    }
    MPI_Barrier(comm);
}

void write_output_file(const KmerListS& kmerlist, const std::string& output_dir, MPI_Comm comm)
{
    int rank;
    MPI_Comm_rank(comm, &rank);
#if LOG_LEVEL >= 1
    if (rank == 0) std::cout << "Writing output files..." << std::endl;
#endif
    const std::string fname = output_dir + "/" + std::to_string(rank) + ".out";
    std::ofstream ofs(fname, std::ios::binary);
    if (!ofs) {
        std::cerr << "Error: cannot open output file " << fname << std::endl;
        MPI_Abort(comm, 1);
    }
    /* "<K bases>\t<cnt>\n" per entry, the reference's lines (hysortk.cpp:159-162) without its flush per line: blocks of
     * entries are formatted by the host threads into per-thread strings and written in order */
    constexpr size_t BLOCK = size_t(1) << 20;
    const size_t n = kmerlist.size();
    int nt = 1;
    #pragma omp parallel
    {
        #pragma omp single
        nt = omp_get_num_threads();
    }
    /* two sets of per-thread strings: while block i is being written (by thread 0, inside the parallel region), the
     * other threads already format block i + 1 */
    std::vector<std::string> part[2] = {std::vector<std::string>(static_cast<size_t>(nt)), std::vector<std::string>(static_cast<size_t>(nt))};
    auto format_block = [&](size_t b0, std::vector<std::string>& dst, size_t t, size_t nthr) {
        const size_t b1 = std::min(n, b0 + BLOCK);
        {
            const size_t lo = b0 + (b1 - b0) * t / nthr, hi = b0 + (b1 - b0) * (t + 1) / nthr;
            std::string& out = dst[t];
            out.clear();
            out.reserve((hi - lo) * (KMER_SIZE + 8));
            char line[KMER_SIZE + 24];
            for (size_t i = lo; i < hi; ++i) {
                const KmerListEntryS& e = kmerlist[i];
                const uint64_t *w = static_cast<const uint64_t *>(e.kmer.GetBytes());
                for (int j = 0; j < KMER_SIZE; ++j) line[j] = "ACGT"[(w[j >> 5] >> (2 * (31 - (j & 31)))) & 3];
                int len = KMER_SIZE;
                line[len++] = '\t';
                char digits[24];
                int nd = 0;
                unsigned long long c = e.cnt;
                do { digits[nd++] = static_cast<char>('0' + c % 10); c /= 10; } while (c);
                while (nd) line[len++] = digits[--nd];
                line[len++] = '\n';
                out.append(line, static_cast<size_t>(len));
            }
        }
    };
    const size_t nblocks = (n + BLOCK - 1) / BLOCK;
    #pragma omp parallel num_threads(nt)
    {
        const size_t t = static_cast<size_t>(omp_get_thread_num());
        const size_t nthr = static_cast<size_t>(omp_get_num_threads());
        if (nblocks) format_block(0, part[0], t, nthr);
        #pragma omp barrier
        for (size_t b = 0; b < nblocks; ++b) {
            /* thread 0 writes block b; everybody (thread 0 afterwards) formats its share of block b + 1 into the other set */
            if (t == 0) for (size_t u = 0; u < nthr; ++u) ofs.write(part[b & 1][u].data(), static_cast<std::streamsize>(part[b & 1][u].size()));
            if (b + 1 < nblocks) format_block((b + 1) * BLOCK, part[(b + 1) & 1], t, nthr);
            #pragma omp barrier
        }
    }
}

} // namespace hysortk
