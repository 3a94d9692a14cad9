// C entry points around the hysortk C++ API (include/hysortk.hpp) for callers that are not C++: the tests and bench.py
// reach hysortk::kmer_count — the function an ELBA-style caller links (reference include/hysortk.hpp:12,
// src/hysortk.cpp:36-95) — through these, so that what they time and check is the whole call: pageable DnaBuffer in,
// std::vector<KmerListEntryS> out.  Built per compile-time configuration (K, M, L, U, EXT) into
// hysortk_b200/_api/libhysortk_api_k<K>_m<M>_l<L>_u<U>_e<EXT>.so by hysortk_b200/cxxapi.py, on top of
// libhysortk_b200.so.  With HSK_MPI_SIZE / HSK_MPI_RANK / HSK_MPI_SESSION in the environment the process is one rank of
// a multi-rank job (bundled MPI stand-in, hysortk_b200/shim/mpi.h); with a real MPI the launcher decides.
#include "hysortk.hpp"

#include <omp.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <memory>
#include <string>
#include <unistd.h>

using namespace hysortk;

namespace {
struct api_handle {
    std::unique_ptr<KmerListS> list;
    double seconds = 0.0;
};
thread_local std::string g_api_error;

/* a DnaBuffer as a caller would hold it: plain heap memory (pageable), adopted by the buffer */
std::unique_ptr<DnaBuffer> make_buffer(const uint8_t *packed, size_t nbytes, const size_t *readlens, size_t nreads)
{
    uint8_t *buf = new uint8_t[nbytes ? nbytes : 1];
    std::memcpy(buf, packed, nbytes);
    return std::make_unique<DnaBuffer>(nbytes, nreads, buf, readlens);
}
} // namespace

extern "C" {

const char *hsk_api_last_error(void) { return g_api_error.c_str(); }

void hsk_api_params(int *k, int *m, int *l, int *u, int *ext, int *nwords)
{
    *k = KMER_SIZE; *m = MINIMIZER_SIZE; *l = LOWER_KMER_FREQ; *u = UPPER_KMER_FREQ; *ext = EXTENSION;
    *nwords = TKmer::NBYTES / 8;
}

/* host threads of the calling thread's OpenMP regions (the KmerListS is filled with that many) */
void hsk_api_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

int hsk_api_rank(void) { int r; MPI_Comm_rank(MPI_COMM_WORLD, &r); return r; }
int hsk_api_nranks(void) { int n; MPI_Comm_size(MPI_COMM_WORLD, &n); return n; }

/* hysortk::kmer_count on (packed, readlens); the time is that of the call alone, as the reference logs it
 * ("Overall kmer counting (Excluding I/O)", src/hysortk.cpp:58,91).  Returns null on error. */
api_handle *hsk_api_kmer_count(const uint8_t *packed, size_t nbytes, const size_t *readlens, size_t nreads)
{
    try {
        auto dna = make_buffer(packed, nbytes, readlens, nreads);
        auto h = std::make_unique<api_handle>();
        MPI_Barrier(MPI_COMM_WORLD);
        const auto t0 = std::chrono::steady_clock::now();
        h->list = kmer_count(*dna, MPI_COMM_WORLD);
        h->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return h.release();
    } catch (const std::exception& e) {
        g_api_error = e.what();
        return nullptr;
    }
}

/* the same through read_dna_buffer (needs <fasta>.fai) */
api_handle *hsk_api_kmer_count_fasta(const char *fasta)
{
    try {
        auto dna = read_dna_buffer(std::string(fasta), MPI_COMM_WORLD);
        auto h = std::make_unique<api_handle>();
        const auto t0 = std::chrono::steady_clock::now();
        h->list = kmer_count(*dna, MPI_COMM_WORLD);
        h->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return h.release();
    } catch (const std::exception& e) {
        g_api_error = e.what();
        return nullptr;
    }
}

/* `warmup` + `steps` calls of kmer_count on one DnaBuffer; seconds[i] = wall time of timed call i on this rank (every
 * call starts behind a barrier over the ranks; the list of the previous call is released outside the timed part).
 * n_kept / n_occ / checksum describe the last result.  Returns 0, or 1 on error. */
int hsk_api_bench(const uint8_t *packed, size_t nbytes, const size_t *readlens, size_t nreads, int warmup, int steps,
                  double *seconds, uint64_t *n_kept, uint64_t *n_occ, uint64_t *checksum)
{
    try {
        auto dna = make_buffer(packed, nbytes, readlens, nreads);
        std::unique_ptr<KmerListS> list;
        for (int i = 0; i < warmup + steps; ++i) {
            list.reset();
            MPI_Barrier(MPI_COMM_WORLD);
            const auto t0 = std::chrono::steady_clock::now();
            list = kmer_count(*dna, MPI_COMM_WORLD);
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (i >= warmup) seconds[i - warmup] = s;
        }
        uint64_t sum = 0, occ = 0;
        if (list) {
            for (const auto& e : *list) {
                const uint64_t *w = static_cast<const uint64_t *>(e.kmer.GetBytes());
                sum += (w[0] * 0x9E3779B97F4A7C15ULL ^ (w[0] >> 29)) * e.cnt;
#if EXTENSION == 1
                occ += e.pos.size();
#endif
            }
        }
        if (n_kept) *n_kept = list ? list->size() : 0;
        if (n_occ) *n_occ = occ;
        if (checksum) *checksum = sum;
        return 0;
    } catch (const std::exception& e) {
        g_api_error = e.what();
        return 1;
    }
}

double hsk_api_seconds(const api_handle *h) { return h->seconds; }
size_t hsk_api_size(const api_handle *h) { return h->list->size(); }

size_t hsk_api_total_occurrences(const api_handle *h)
{
#if EXTENSION == 1
    size_t n = 0;
    for (const auto& e : *h->list) n += e.pos.size();
    return n;
#else
    (void)h;
    return 0;
#endif
}

/* words: size() * nwords u64 (entry-major, word 0 first); cnt: size() u64; EXTENSION: occ_off size() + 1, pos / rid
 * one per occurrence, in the order of the entries' own vectors (KmerListEntryS::pos / rid, include/kmer.hpp) */
void hsk_api_export(const api_handle *h, uint64_t *words, uint64_t *cnt, uint64_t *occ_off, uint32_t *pos, int32_t *rid)
{
    const int nw = TKmer::NBYTES / 8;
    size_t o = 0;
    for (size_t i = 0; i < h->list->size(); ++i) {
        const auto& e = (*h->list)[i];
        std::memcpy(words + i * nw, e.kmer.GetBytes(), TKmer::NBYTES);
        cnt[i] = e.cnt;
#if EXTENSION == 1
        if (occ_off) occ_off[i] = o;
        for (size_t j = 0; j < e.pos.size(); ++j) { pos[o] = e.pos[j]; rid[o] = e.rid[j]; ++o; }
#endif
    }
    if (occ_off) occ_off[h->list->size()] = o;
    (void)pos; (void)rid;
}

/* print_kmer_histogram writes to stdout; capture it into `path` */
int hsk_api_print_histogram(const api_handle *h, const char *path)
{
    fflush(stdout);
    std::cout.flush();
    const int saved = dup(1);
    const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return -1;
    dup2(fd, 1);
    print_kmer_histogram(*h->list, MPI_COMM_WORLD);
    std::cout.flush();
    fflush(stdout);
    dup2(saved, 1);
    close(fd);
    close(saved);
    return 0;
}

/* write_output_file: <dir>/<rank>.out */
void hsk_api_write_output(const api_handle *h, const char *dir) { write_output_file(*h->list, std::string(dir), MPI_COMM_WORLD); }

void hsk_api_free(api_handle *h) { delete h; }

/* frees the GPU engine of the process (collective over the ranks) */
void hsk_api_release(void) { release_gpu_engine(); }

} /* extern "C" */
