/*
 * TEST INFRASTRUCTURE — not part of the product.
 *
 * C-ABI wrapper around the UNMODIFIED reference (CornellHPC/HySortK) so that tests and
 * the `cpu_baseline` / `--impl reference` bench arm can call the reference's own
 * `hysortk::kmer_count` (reference src/hysortk.cpp:36-95) on an in-memory packed read
 * buffer.  This file is ours; the reference sources are compiled from where they lie
 * under /root/reference by oracle/build_ref.sh and are never copied into this repo.
 *
 * Built once per compile-time configuration (K, M, L, U, EXT) into
 * oracle/_ref/libhysortk_ref_k<K>_m<M>_l<L>_u<U>_e<EXT>.so
 */
#include "hysortk.hpp"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <unistd.h>
#include <fcntl.h>

using namespace hysortk;

struct ref_handle {
    std::unique_ptr<KmerListS> list;
    double seconds = 0.0;
};

extern "C" {

void ref_params(int *k, int *m, int *l, int *u, int *ext, int *nwords)
{
    *k = KMER_SIZE; *m = MINIMIZER_SIZE; *l = LOWER_KMER_FREQ; *u = UPPER_KMER_FREQ; *ext = EXTENSION;
    *nwords = TKmer::NBYTES / 8;
}

/* Runs the reference kmer_count on (packed, readlens).  The buffer is copied because
 * DnaBuffer takes ownership of the pointer it is given (dnabuffer.hpp:40). */
ref_handle *ref_kmer_count(const uint8_t *packed, size_t nbytes, const size_t *readlens, size_t nreads)
{
    uint8_t *buf = new uint8_t[nbytes ? nbytes : 1];
    std::memcpy(buf, packed, nbytes);
    DnaBuffer dna(nbytes, nreads, buf, readlens);
    auto *h = new ref_handle;
    auto t0 = std::chrono::steady_clock::now();
    h->list = kmer_count(dna, MPI_COMM_WORLD);
    auto t1 = std::chrono::steady_clock::now();
    h->seconds = std::chrono::duration<double>(t1 - t0).count();
    return h;
}

/* Same, but through the reference's own FASTA reader (needs <fasta>.fai). */
ref_handle *ref_kmer_count_fasta(const char *fasta)
{
    auto dna = read_dna_buffer(std::string(fasta), MPI_COMM_WORLD);
    auto *h = new ref_handle;
    auto t0 = std::chrono::steady_clock::now();
    h->list = kmer_count(*dna, MPI_COMM_WORLD);
    auto t1 = std::chrono::steady_clock::now();
    h->seconds = std::chrono::duration<double>(t1 - t0).count();
    return h;
}

double ref_seconds(const ref_handle *h) { return h->seconds; }
size_t ref_size(const ref_handle *h) { return h->list->size(); }

size_t ref_total_occurrences(const ref_handle *h)
{
#if EXTENSION == 1
    size_t n = 0;
    for (const auto &e : *h->list) n += e.pos.size();
    return n;
#else
    (void)h; return 0;
#endif
}

/* words: size()*nwords u64 (entry-major, word 0 first); cnt: size() u64.
 * EXT: occ_off size()+1, pos/rid ref_total_occurrences() entries. */
void ref_export(const ref_handle *h, uint64_t *words, uint64_t *cnt, uint64_t *occ_off, uint32_t *pos, int32_t *rid)
{
    const int nw = TKmer::NBYTES / 8;
    size_t o = 0;
    for (size_t i = 0; i < h->list->size(); ++i) {
        const auto &e = (*h->list)[i];
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

/* print_kmer_histogram (hysortk.cpp:98-136) writes to stdout; capture it into `path`. */
int ref_print_histogram(const ref_handle *h, const char *path)
{
    fflush(stdout);
    std::cout.flush();
    int saved = dup(1);
    int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
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

/* write_output_file (hysortk.cpp:138-164): <dir>/0.out */
void ref_write_output(const ref_handle *h, const char *dir)
{
    write_output_file(*h->list, std::string(dir), MPI_COMM_WORLD);
}

void ref_free(ref_handle *h) { delete h; }

} /* extern "C" */
