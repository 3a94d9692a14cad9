/*
 * TEST INFRASTRUCTURE — CPU restatement of the reference's kmer_count path.
 *
 * This is the parity oracle: a plain-C restatement of what CornellHPC/HySortK computes on
 * the `kmer_count` hot path, function by function, each citing the reference file:line it
 * follows.  It is used ONLY by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg, and only as the checker.  The product (hysortk_b200/, include/)
 * never links, imports or executes anything in this directory.
 *
 * Parity pin: the reference ships no tests/golden vectors, so this oracle is
 * pinned against the reference ITSELF: oracle/build_ref.sh compiles the unmodified
 * reference sources into oracle/_ref/ and tests/test_oracle_vs_reference.py requires
 * identical (k-mer, count[, (ReadId, PosInRead) multiset]) results and histogram text on
 * seeded synthetic inputs; reference outputs generated in this container are committed as
 * fixtures under tests/golden/ (generator: tests/golden/make_golden.py).
 */
#ifndef HSK_ORACLE_H_
#define HSK_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_WORDS 3
#define ORC_MAX_SUPERMER_LEN 250 /* supermer.hpp:20 */

typedef struct { uint64_t w[ORC_MAX_WORDS]; } orc_kmer;

typedef struct {
    int k, m, lower, upper, ext, nwords;
    uint64_t total_kmers;    /* N = sum over reads of max(0, len-K+1) */
    uint64_t n_supermers;    /* supermer path only */
    uint64_t supermer_bytes; /* sum of cnt_bytes(len) + sizeof(length_t) */
    uint64_t n;              /* kept k-mers */
    uint64_t *words;         /* n * nwords, entry-major, word 0 first (kmer.hpp:165-185) */
    uint64_t *cnt;           /* n */
    uint64_t *occ_off;       /* n+1 (ext only) */
    uint32_t *pos;           /* occ_off[n] (ext only) */
    int32_t *rid;            /* occ_off[n] (ext only) */
    uint64_t hist_len;       /* upper+1 */
    uint64_t *hist;          /* hist[c] = #kept k-mers with count c (hysortk.cpp:106-113) */
} orc_result;

/* dnaseq.hpp:138-156 code table, dnaseq.cpp:9-31 packing */
int orc_char_code(char c);
size_t orc_bytes_needed(size_t len); /* dnaseq.hpp:126 */
void orc_pack_read(const char *ascii, size_t len, uint8_t *out);
int orc_base_at(const uint8_t *mem, size_t i); /* dnaseq.cpp:50-57 */

int orc_nwords(int k); /* kmer.hpp:343-345 */
void orc_kmer_set(const uint8_t *mem, size_t start, int k, orc_kmer *out);  /* kmer.hpp:165-185 */
void orc_kmer_extend(const orc_kmer *in, int k, int code, orc_kmer *out);   /* kmer.hpp:247-263 */
void orc_kmer_twin(const orc_kmer *in, int k, orc_kmer *out);               /* kmer.hpp:265-296 */
int orc_kmer_less(const orc_kmer *a, const orc_kmer *b, int nwords);        /* kmer.hpp:216-229 */
void orc_kmer_rep(const orc_kmer *in, int k, orc_kmer *out);                /* kmer.hpp:298-303 */
void orc_kmer_string(const orc_kmer *in, int k, char *out);                 /* kmer.hpp:147-163 */

uint64_t orc_murmur3_64(const void *key, uint32_t len); /* hashfuncs.cpp:42-119,233-238 */

/* kmerops.cpp:1010-1047: task id of every k-mer of one read (len-K+1 entries); returns count */
size_t orc_read_destinations(const uint8_t *mem, size_t len, int k, int m, int ntasks, int *dest);

/* Full path.  via_supermers=1 follows prepare_supermer -> exchange -> filter (kmerops.cpp:23-250)
 * with `ntasks` tasks; via_supermers=0 is the direct definition (every window of every read).
 * Results are returned in canonical order: k-mers ascending by Kmer::operator<, occurrences
 * ascending by (rid, pos). */
orc_result *orc_kmer_count(const uint8_t *packed, const uint64_t *readlens, uint64_t nreads, int k, int m, int lower,
                           int upper, int ext, int ntasks, int32_t readid_base, int via_supermers);
void orc_free(orc_result *r);

/* hysortk.cpp:98-136 text ("#count\tnumkmers\n" then "i\thisto[i]\n" for non-zero bins, blank line). */
size_t orc_histogram_text(const orc_result *r, char *out, size_t cap);
/* hysortk.cpp:149-162: one "KMER\tcnt\n" line per entry. Returns bytes written (or needed if out==NULL). */
size_t orc_output_text(const orc_result *r, char *out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
