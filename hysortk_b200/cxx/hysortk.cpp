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
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>

namespace hysortk {

namespace {

struct FaiRecord { size_t len, pos, bases; };

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

/* one engine context per process, re-created when the communicator shape changes */
struct Engine {
    hsk_ctx *ctx = nullptr;
    int rank = -1, nranks = -1;
    ~Engine() { if (ctx) hsk_destroy(ctx); }
};
Engine g_engine;

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
            std::istringstream(line) >> name >> r.len >> r.pos >> r.bases;
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

    std::vector<size_t> lens;
    lens.reserve(hi - lo);
    size_t maxlen = 0;
    for (size_t i = lo; i < hi; ++i) { lens.push_back(rec[i].len); maxlen = std::max(maxlen, rec[i].len); }
    auto dna = std::make_shared<DnaBuffer>(DnaBuffer::computebufsize(lens));
    if (hi > lo) {
        std::ifstream fa(fasta_fname, std::ios::binary);
        if (!fa) throw std::runtime_error("cannot open FASTA file " + fasta_fname);
        const size_t start = rec[lo].pos;
        const FaiRecord& last = rec[hi - 1];
        const size_t end = last.pos + last.len + (last.bases ? last.len / last.bases : 0) + 1;
        std::string chunk(end - start, '\n');
        fa.seekg(static_cast<std::streamoff>(start));
        fa.read(&chunk[0], static_cast<std::streamsize>(chunk.size()));
        std::string seq(maxlen, 'A');
        for (size_t i = lo; i < hi; ++i) {
            const FaiRecord& r = rec[i];
            size_t src = r.pos - start, dst = 0, remain = r.len;
            while (remain > 0) {   /* strip the newline after every `bases` characters */
                const size_t cnt = std::min(r.bases ? r.bases : remain, remain);
                std::memcpy(&seq[dst], &chunk[src], cnt);
                dst += cnt; remain -= cnt; src += cnt + 1;
            }
            dna->push_back(seq.data(), r.len);
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

    const size_t n = mydna.size();
    std::vector<uint64_t> lens(n);
    for (size_t i = 0; i < n; ++i) lens[i] = mydna[i].size();
    const uint8_t *bytes = n ? mydna.getbufoffset(0) : nullptr;
    const size_t nbytes = n ? mydna.getrangebufsize(0, n) : 0;

    hsk_result res;
    if (hsk_count(ctx, bytes, nbytes, lens.data(), n, readoffset, &res)) throw std::runtime_error(hsk_last_error());

    auto list = std::make_unique<KmerListS>();
    list->resize(res.n_kept);
    constexpr int NW = TKmer::NBYTES / 8;
#if EXTENSION == 0
    static_assert(sizeof(KmerListEntryS) == 8 * (NW + 1), "entry layout");
    if (hsk_fill_entries(ctx, list->data(), res.n_kept)) throw std::runtime_error(hsk_last_error());
#else
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < (size_t)res.n_kept; ++i) {
        KmerListEntryS& e = (*list)[i];
        e.kmer = TKmer(static_cast<const void *>(res.kmer_words + i * NW));
        e.cnt = res.cnt[i];
        e.pos.assign(res.pos + res.occ_off[i], res.pos + res.occ_off[i + 1]);
        e.rid.assign(res.rid + res.occ_off[i], res.rid + res.occ_off[i + 1]);
    }
#endif

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
    for (const auto& e : kmerlist) {
        if (e.cnt <= UPPER_KMER_FREQ) histo[e.cnt]++;
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
    std::string buf;
    buf.reserve(1 << 22);
    for (const auto& e : kmerlist) {   /* "<K bases>\t<cnt>\n" per entry (reference hysortk.cpp:159-162) */
        buf += e.kmer.GetString();
        buf += '\t';
        buf += std::to_string(e.cnt);
        buf += '\n';
        if (buf.size() > (1u << 22) - 256) { ofs.write(buf.data(), static_cast<std::streamsize>(buf.size())); buf.clear(); }
    }
    ofs.write(buf.data(), static_cast<std::streamsize>(buf.size()));
}

} // namespace hysortk
