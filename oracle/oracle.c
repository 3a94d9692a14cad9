/*
 * TEST INFRASTRUCTURE — see oracle.h.  Plain-C restatement of the reference's kmer_count
 * path (CornellHPC/HySortK); every function cites the reference lines it restates.
 * Clarity over speed: this is the checker, never the thing measured or shipped.
 */
#include "oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ packing */

/* dnaseq.hpp:138-156: A/a 0, C/c 1, G/g 2, T/t 3, N/n 0, everything else 4 (undefined) */
int orc_char_code(char c)
{
    switch (c) {
    case 'A': case 'a': case 'N': case 'n': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

/* dnaseq.hpp:126 */
size_t orc_bytes_needed(size_t len) { return (len + 3) / 4; }

/* dnaseq.cpp:9-31: base 4b+i of the read goes to bits 7-2i..6-2i of byte b, tail zero padded */
void orc_pack_read(const char *s, size_t len, uint8_t *out)
{
    size_t nbytes = orc_bytes_needed(len);
    for (size_t b = 0; b < nbytes; ++b) {
        uint8_t byte = 0;
        for (int i = 0; i < 4; ++i) {
            size_t p = 4 * b + (size_t)i;
            if (p >= len) break;
            byte |= (uint8_t)((orc_char_code(s[p]) & 3) << (6 - 2 * i));
        }
        out[b] = byte;
    }
}

/* dnaseq.cpp:50-57 */
int orc_base_at(const uint8_t *mem, size_t i) { return (mem[i / 4] >> (6 - 2 * (i % 4))) & 3; }

/* ------------------------------------------------------------------ k-mer words */

/* kmer.hpp:343-345 */
int orc_nwords(int k) { return k <= 32 ? 1 : (k <= 64 ? 2 : 3); }

/* kmer.hpp:165-185 (set_kmer): base i -> word i/32, shifted to 2*(31 - i%32) */
void orc_kmer_set(const uint8_t *mem, size_t start, int k, orc_kmer *out)
{
    memset(out, 0, sizeof(*out));
    for (int i = 0; i < k; ++i) {
        uint64_t code = (uint64_t)orc_base_at(mem, start + (size_t)i);
        out->w[i / 32] |= code << (2 * (31 - i % 32));
    }
}

/* kmer.hpp:247-263 (GetExtension): drop the first base, append `code` as the new last base.
 * The reference's shift 2*(32 - K%32) is 64 (undefined) for K%32==0; the intended position is bit 0. */
void orc_kmer_extend(const orc_kmer *in, int k, int code, orc_kmer *out)
{
    int nw = orc_nwords(k);
    orc_kmer e;
    memset(&e, 0, sizeof(e));
    e.w[0] = in->w[0] << 2;
    for (int i = 1; i < nw; ++i) {
        e.w[i - 1] |= (in->w[i] >> 62) & 3;
        e.w[i] = in->w[i] << 2;
    }
    e.w[nw - 1] |= (uint64_t)code << ((2 * (32 - (k % 32))) & 63);
    *out = e;
}

/* kmer.hpp:107-130: reverse complement of the four bases held in one byte */
static uint64_t tetramer_twin(uint8_t code)
{
    uint8_t r = 0;
    for (int i = 0; i < 4; ++i) {
        uint8_t base = (code >> (2 * i)) & 3;      /* i-th base from the right */
        r |= (uint8_t)((3 - base) << (6 - 2 * i)); /* complemented, now i-th from the left */
    }
    return r;
}

/* kmer.hpp:265-296 (GetTwin): byte-wise tetramer reverse complement with the word order
 * reversed, then left-align by 2*(32 - K%32) bits across the words. */
void orc_kmer_twin(const orc_kmer *in, int k, orc_kmer *out)
{
    int nw = orc_nwords(k);
    orc_kmer t;
    memset(&t, 0, sizeof(t));
    for (int l = 0; l < nw; ++l) {
        uint64_t longmer = in->w[l];
        for (int i = 0; i < 64; i += 8) {
            uint8_t bytemer = (uint8_t)((longmer >> i) & 0xff);
            t.w[nw - 1 - l] |= tetramer_twin(bytemer) << (56 - i);
        }
    }
    int shift = (k % 32) ? 2 * (32 - (k % 32)) : 0;
    if (shift) {
        uint64_t mask = ((1ULL << shift) - 1) << (64 - shift);
        t.w[0] <<= shift;
        for (int i = 1; i < nw; ++i) {
            t.w[i - 1] |= (t.w[i] & mask) >> (64 - shift);
            t.w[i] <<= shift;
        }
    }
    *out = t;
}

/* kmer.hpp:216-229 (operator<): word 0 is the most significant */
int orc_kmer_less(const orc_kmer *a, const orc_kmer *b, int nw)
{
    for (int i = 0; i < nw; ++i) {
        if (a->w[i] < b->w[i]) return 1;
        if (a->w[i] > b->w[i]) return 0;
    }
    return 0;
}

/* kmer.hpp:298-303 (GetRep) */
void orc_kmer_rep(const orc_kmer *in, int k, orc_kmer *out)
{
    orc_kmer t;
    orc_kmer_twin(in, k, &t);
    *out = orc_kmer_less(&t, in, orc_nwords(k)) ? t : *in;
}

/* kmer.hpp:147-163 (GetString) */
void orc_kmer_string(const orc_kmer *in, int k, char *out)
{
    for (int i = 0; i < k; ++i) out[i] = "ACGT"[(in->w[i / 32] >> (2 * (31 - i % 32))) & 3];
    out[k] = 0;
}

/* ------------------------------------------------------------------ hash */

static uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

/* hashfuncs.cpp:29-38 */
static uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

/* hashfuncs.cpp:42-119 (MurmurHash3 x64-128) with seed 313, low word (hashfuncs.cpp:233-238) */
uint64_t orc_murmur3_64(const void *key, uint32_t len)
{
    const uint8_t *data = (const uint8_t *)key;
    const uint32_t nblocks = len / 16;
    uint64_t h1 = 313, h2 = 313;
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    for (uint32_t i = 0; i < nblocks; ++i) {
        uint64_t k1, k2;
        memcpy(&k1, data + 16 * i, 8);
        memcpy(&k2, data + 16 * i + 8, 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *tail = data + nblocks * 16;
    uint64_t k1 = 0, k2 = 0;
    uint32_t rem = len & 15;
    for (uint32_t i = rem; i > 8; --i) k2 ^= (uint64_t)tail[i - 1] << (8 * (i - 9));
    if (rem > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    for (uint32_t i = (rem > 8 ? 8 : rem); i > 0; --i) k1 ^= (uint64_t)tail[i - 1] << (8 * (i - 1));
    if (rem > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= len; h2 ^= len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

/* ------------------------------------------------------------------ minimizers / destinations */

/* supermer.hpp:315-342 (GetRepMmers) + supermer.hpp:307-313 (GetHash): hash of every canonical
 * m-mer of the read.  An Mmer is a Kmer with MINIMIZER_SIZE in place of KMER_SIZE. */
static size_t read_mmer_hashes(const uint8_t *mem, size_t len, int m, uint64_t *hash)
{
    if (len < (size_t)m) return 0;
    size_t n = len - (size_t)m + 1;
    int mw = orc_nwords(m);
    orc_kmer cur, rep;
    orc_kmer_set(mem, 0, m, &cur);
    for (size_t i = 0; i < n; ++i) {
        if (i) orc_kmer_extend(&cur, m, orc_base_at(mem, i + (size_t)m - 1), &cur);
        orc_kmer_rep(&cur, m, &rep);
        hash[i] = orc_murmur3_64(rep.w, (uint32_t)(8 * mw));
    }
    return n;
}

/* kmerops.cpp:1010-1073: monotone deque sliding-window minimum of the m-mer hashes over the
 * K-M+1 m-mers of each k-mer; destination = min hash % tot_tasks (kmerops.cpp:1044-1047). */
size_t orc_read_destinations(const uint8_t *mem, size_t len, int k, int m, int ntasks, int *dest)
{
    if (len < (size_t)k) return 0;
    size_t nm = len - (size_t)m + 1;
    uint64_t *hash = (uint64_t *)malloc(nm * sizeof(uint64_t));
    uint64_t *dq_hash = (uint64_t *)malloc(nm * sizeof(uint64_t));
    long *dq_pos = (long *)malloc(nm * sizeof(long));
    read_mmer_hashes(mem, len, m, hash);
    size_t front = 0, back = 0, nd = 0; /* deque = [front, back) */
    long head = 0;
#define DQ_INSERT(h, p)                                                        \
    do {                                                                       \
        while (back > front && dq_hash[back - 1] > (h)) --back;                \
        dq_hash[back] = (h); dq_pos[back] = (p); ++back;                       \
    } while (0)
    for (; head < k - m; ++head) DQ_INSERT(hash[head], head);
    long tail = head - k + m - 1;
    for (; head < (long)nm; ++head, ++tail) {
        DQ_INSERT(hash[head], head);
        while (back > front && dq_pos[front] <= tail) ++front;
        dest[nd++] = (int)(dq_hash[front] % (uint64_t)ntasks);
    }
#undef DQ_INSERT
    free(hash); free(dq_hash); free(dq_pos);
    return nd;
}

/* ------------------------------------------------------------------ seeds, sort, count */

typedef struct {
    orc_kmer kmer;
    uint32_t pos;
    int32_t rid;
} seed_t; /* kmer.hpp:415-449 (KmerSeedStruct) */

typedef struct { seed_t *v; size_t n, cap; } seedvec;

static void seed_push(seedvec *s, const orc_kmer *km, uint32_t pos, int32_t rid)
{
    if (s->n == s->cap) {
        s->cap = s->cap ? 2 * s->cap : 1024;
        s->v = (seed_t *)realloc(s->v, s->cap * sizeof(seed_t));
    }
    s->v[s->n].kmer = *km; s->v[s->n].pos = pos; s->v[s->n].rid = rid;
    ++s->n;
}

static int g_nw; /* comparator context (single-threaded checker) */

/* kmerops.cpp:1382-1407 -> raduls.h:976-990 / paradissort.hpp:211-215: records ordered by key bytes
 * W-1 .. 0, i.e. word NLONGS-1 most significant. */
static int cmp_task_order(const void *a, const void *b)
{
    const seed_t *x = (const seed_t *)a, *y = (const seed_t *)b;
    for (int i = g_nw - 1; i >= 0; --i) {
        if (x->kmer.w[i] < y->kmer.w[i]) return -1;
        if (x->kmer.w[i] > y->kmer.w[i]) return 1;
    }
    return 0;
}

/* canonical comparison order: Kmer::operator< (word 0 first), then (rid, pos) */
static int cmp_canonical(const void *a, const void *b)
{
    const seed_t *x = (const seed_t *)a, *y = (const seed_t *)b;
    for (int i = 0; i < g_nw; ++i) {
        if (x->kmer.w[i] < y->kmer.w[i]) return -1;
        if (x->kmer.w[i] > y->kmer.w[i]) return 1;
    }
    if (x->rid != y->rid) return x->rid < y->rid ? -1 : 1;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    return 0;
}

static int kmer_eq(const orc_kmer *a, const orc_kmer *b, int nw)
{
    for (int i = 0; i < nw; ++i) if (a->w[i] != b->w[i]) return 0;
    return 1;
}

/* kept occurrences (k-mer repeated once per occurrence, tagged with the run's count) */
typedef struct { seedvec occ; uint64_t nkept; } keptlist;

/* kmerops.cpp:1410-1445 (count_sorted_kmers): run-length over the sorted seeds, keep runs with
 * LOWER <= cnt <= UPPER (kmerops.cpp:1428); with EXTENSION the run's (pos, rid) are kept too. */
static void count_sorted(const seed_t *v, size_t n, int nw, int lower, int upper, keptlist *out)
{
    size_t i = 0;
    while (i < n) {
        size_t j = i + 1;
        while (j < n && kmer_eq(&v[j].kmer, &v[i].kmer, nw)) ++j;
        uint64_t c = j - i;
        if (c >= (uint64_t)lower && c <= (uint64_t)upper) {
            for (size_t t = i; t < j; ++t) seed_push(&out->occ, &v[t].kmer, v[t].pos, v[t].rid);
            ++out->nkept;
        }
        i = j;
    }
}

/* ------------------------------------------------------------------ supermers */

/* kmerops.hpp:33-41: bytes used by a supermer of `len` bases on the wire (one spare byte when len%4==0) */
static int cnt_bytes(int len) { return (len + (4 - len % 4)) / 4; }

typedef struct {
    uint32_t *len; uint32_t *pos; int32_t *rid; size_t n, cap; /* length_t records (kmer.hpp:350-360) */
    uint8_t *bytes; size_t nbytes, bcap;                        /* packed supermer bases */
} task_t;

/* kmerops.cpp:1096-1107 (copy_bits): re-pack `len` bases starting at base `start` from bit 0 */
static void task_append(task_t *t, const uint8_t *src, uint32_t start, int len, int32_t rid)
{
    if (t->n == t->cap) {
        t->cap = t->cap ? 2 * t->cap : 256;
        t->len = (uint32_t *)realloc(t->len, t->cap * 4);
        t->pos = (uint32_t *)realloc(t->pos, t->cap * 4);
        t->rid = (int32_t *)realloc(t->rid, t->cap * 4);
    }
    t->len[t->n] = (uint32_t)len; t->pos[t->n] = start; t->rid[t->n] = rid; ++t->n;
    size_t nb = (size_t)cnt_bytes(len);
    if (t->nbytes + nb > t->bcap) {
        while (t->nbytes + nb > t->bcap) t->bcap = t->bcap ? 2 * t->bcap : 4096;
        t->bytes = (uint8_t *)realloc(t->bytes, t->bcap);
    }
    memset(t->bytes + t->nbytes, 0, nb);
    for (int i = 0; i < len; ++i) {
        int code = orc_base_at(src, (size_t)start + (size_t)i);
        t->bytes[t->nbytes + (size_t)i / 4] |= (uint8_t)(code << (6 - 2 * (i % 4)));
    }
    t->nbytes += nb;
}

/* kmerops.cpp:1109-1148 (SupermerEncoder::encode): maximal runs of equal destination, at most
 * MAX_SUPERMER_LEN bases (kmerops.cpp:1120), appended to the destination task. */
static void encode_read(task_t *tasks, const int *dest, size_t nd, const uint8_t *mem, int k, int32_t rid)
{
    uint32_t start_pos = 0;
    int cnt = 1;
    int last_dst = dest[0];
    for (size_t i = 1; i <= nd; ++i) {
        if (i == nd || dest[i] != last_dst || cnt == ORC_MAX_SUPERMER_LEN - k + 1) {
            task_append(&tasks[last_dst], mem, start_pos, cnt + k - 1, rid);
            if (i < nd) last_dst = dest[i];
            cnt = 0;
            start_pos = (uint32_t)i;
        }
        ++cnt;
    }
}

/* kmerops.cpp:484-521 (receive_from_buffer_stage2) + kmer.hpp:313-340 (GetRepKmers): every
 * canonical k-mer of each supermer, tagged pos+i / rid (kmerops.cpp:507). */
static void expand_task(const task_t *t, int k, seedvec *out)
{
    size_t off = 0;
    for (size_t s = 0; s < t->n; ++s) {
        int len = (int)t->len[s];
        const uint8_t *mem = t->bytes + off;
        orc_kmer cur, rep;
        orc_kmer_set(mem, 0, k, &cur);
        for (int i = 0; i < len - k + 1; ++i) {
            if (i) orc_kmer_extend(&cur, k, orc_base_at(mem, (size_t)(i + k - 1)), &cur);
            orc_kmer_rep(&cur, k, &rep);
            seed_push(out, &rep, t->pos[s] + (uint32_t)i, t->rid[s]);
        }
        off += (size_t)cnt_bytes(len);
    }
}

/* ------------------------------------------------------------------ whole path */

static orc_result *finish(keptlist *kept, orc_result *r)
{
    /* canonical order for comparison */
    g_nw = r->nwords;
    qsort(kept->occ.v, kept->occ.n, sizeof(seed_t), cmp_canonical);
    r->n = kept->nkept;
    r->words = (uint64_t *)calloc((size_t)(r->n ? r->n : 1) * (size_t)r->nwords, 8);
    r->cnt = (uint64_t *)calloc((size_t)(r->n ? r->n : 1), 8);
    r->hist_len = (uint64_t)r->upper + 1;
    r->hist = (uint64_t *)calloc((size_t)r->hist_len, 8);
    if (r->ext) {
        r->occ_off = (uint64_t *)calloc((size_t)r->n + 1, 8);
        r->pos = (uint32_t *)calloc(kept->occ.n ? kept->occ.n : 1, 4);
        r->rid = (int32_t *)calloc(kept->occ.n ? kept->occ.n : 1, 4);
    }
    size_t i = 0, e = 0;
    while (i < kept->occ.n) {
        size_t j = i + 1;
        while (j < kept->occ.n && kmer_eq(&kept->occ.v[j].kmer, &kept->occ.v[i].kmer, r->nwords)) ++j;
        for (int w = 0; w < r->nwords; ++w) r->words[e * (size_t)r->nwords + (size_t)w] = kept->occ.v[i].kmer.w[w];
        r->cnt[e] = j - i;
        r->hist[j - i]++; /* hysortk.cpp:106-113 */
        if (r->ext) {
            r->occ_off[e] = i;
            for (size_t t = i; t < j; ++t) { r->pos[t] = kept->occ.v[t].pos; r->rid[t] = kept->occ.v[t].rid; }
        }
        ++e;
        i = j;
    }
    if (r->ext) r->occ_off[e] = kept->occ.n;
    free(kept->occ.v);
    return r;
}

orc_result *orc_kmer_count(const uint8_t *packed, const uint64_t *readlens, uint64_t nreads, int k, int m, int lower,
                           int upper, int ext, int ntasks, int32_t readid_base, int via_supermers)
{
    orc_result *r = (orc_result *)calloc(1, sizeof(orc_result));
    r->k = k; r->m = m; r->lower = lower; r->upper = upper; r->ext = ext; r->nwords = orc_nwords(k);
    int nw = r->nwords;
    keptlist kept;
    memset(&kept, 0, sizeof(kept));
    g_nw = nw;

    if (!via_supermers) {
        /* direct definition: every length-K window of every read, canonicalised */
        seedvec all;
        memset(&all, 0, sizeof(all));
        size_t off = 0;
        for (uint64_t i = 0; i < nreads; ++i) {
            size_t len = (size_t)readlens[i];
            const uint8_t *mem = packed + off;
            for (size_t p = 0; p + (size_t)k <= len; ++p) {
                orc_kmer cur, rep;
                orc_kmer_set(mem, p, k, &cur);
                orc_kmer_rep(&cur, k, &rep);
                seed_push(&all, &rep, (uint32_t)p, (int32_t)i + readid_base);
                r->total_kmers++;
            }
            off += orc_bytes_needed(len);
        }
        qsort(all.v, all.n, sizeof(seed_t), cmp_task_order);
        count_sorted(all.v, all.n, nw, lower, upper, &kept);
        free(all.v);
        return finish(&kept, r);
    }

    /* prepare_supermer (kmerops.cpp:23-126) */
    if (ntasks < 1) ntasks = 1;
    task_t *tasks = (task_t *)calloc((size_t)ntasks, sizeof(task_t));
    size_t off = 0;
    for (uint64_t i = 0; i < nreads; ++i) {
        size_t len = (size_t)readlens[i];
        const uint8_t *mem = packed + off;
        if (len >= (size_t)k) { /* kmerops.cpp:1019-1020,1110 */
            size_t nk = len - (size_t)k + 1;
            int *dest = (int *)malloc(nk * sizeof(int));
            size_t nd = orc_read_destinations(mem, len, k, m, ntasks, dest);
            r->total_kmers += nd;
            encode_read(tasks, dest, nd, mem, k, (int32_t)i + readid_base); /* rid: kmerops.cpp:65-70,1018 */
            free(dest);
        }
        off += orc_bytes_needed(len);
    }
    /* exchange_supermer (single rank: every task stays local) + filter_kmer (kmerops.cpp:198-250):
     * per task expand, sort, count; results concatenated in task order (kmerops.cpp:883-904). */
    for (int t = 0; t < ntasks; ++t) {
        r->n_supermers += tasks[t].n;
        r->supermer_bytes += tasks[t].nbytes + tasks[t].n * (ext ? 12 : 4);
        seedvec seeds;
        memset(&seeds, 0, sizeof(seeds));
        expand_task(&tasks[t], k, &seeds);
        qsort(seeds.v, seeds.n, sizeof(seed_t), cmp_task_order);
        count_sorted(seeds.v, seeds.n, nw, lower, upper, &kept);
        free(seeds.v);
        free(tasks[t].len); free(tasks[t].pos); free(tasks[t].rid); free(tasks[t].bytes);
    }
    free(tasks);
    return finish(&kept, r);
}

void orc_free(orc_result *r)
{
    if (!r) return;
    free(r->words); free(r->cnt); free(r->occ_off); free(r->pos); free(r->rid); free(r->hist);
    free(r);
}

/* hysortk.cpp:98-136 */
size_t orc_histogram_text(const orc_result *r, char *out, size_t cap)
{
    size_t n = 0;
#define EMIT(...)                                                              \
    do {                                                                       \
        int w_ = snprintf(out ? out + n : NULL, out && cap > n ? cap - n : 0, __VA_ARGS__); \
        n += (size_t)w_;                                                       \
    } while (0)
    EMIT("#count\tnumkmers\n");
    for (uint64_t i = 1; i < r->hist_len; ++i)
        if (r->hist[i] > 0) EMIT("%llu\t%llu\n", (unsigned long long)i, (unsigned long long)r->hist[i]);
    EMIT("\n");
    return n;
}

/* hysortk.cpp:149-162 + kmer.hpp:75-81 */
size_t orc_output_text(const orc_result *r, char *out, size_t cap)
{
    size_t n = 0;
    char s[128];
    for (uint64_t i = 0; i < r->n; ++i) {
        orc_kmer km;
        memset(&km, 0, sizeof(km));
        for (int w = 0; w < r->nwords; ++w) km.w[w] = r->words[i * (uint64_t)r->nwords + (uint64_t)w];
        orc_kmer_string(&km, r->k, s);
        EMIT("%s\t%llu\n", s, (unsigned long long)r->cnt[i]);
    }
#undef EMIT
    return n;
}
