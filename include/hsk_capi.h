/*
 * hsk_capi.h — C ABI of the B200-native kmer_count engine (libhysortk_b200.so).
 *
 * This is the drop-in boundary of the hot path: plain pointers and sizes, no C++/torch types.
 * The reference (CornellHPC/HySortK) has no C layer; its boundary for this path is the C++
 * function
 *
 *     std::unique_ptr<KmerListS> hysortk::kmer_count(const DnaBuffer&, MPI_Comm)
 *                                       reference include/hysortk.hpp:12, src/hysortk.cpp:36-95
 *
 * whose body (prepare_supermer -> exchange_supermer -> filter_kmer, src/kmerops.cpp:23-250) this
 * library replaces.  include/hysortk.hpp + hysortk_b200/cxx/hysortk.cpp in this repo keep the
 * C++ signature and call the functions below; INTEGRATION.md shows the binding.
 *
 * One context per process per GPU (the reference is one MPI rank per NUMA domain; here one rank
 * per GPU).  All functions return 0 on success, non-zero on error with a message available from
 * hsk_last_error() (the reference throws / aborts: kmerops.cpp:357,472,532,725,1319).
 * Not re-entrant per context, like the reference.
 */
#ifndef HSK_CAPI_H_
#define HSK_CAPI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSK_VERSION 2
#define HSK_NCCL_ID_BYTES 128
#define HSK_MAX_KMER_WORDS 3 /* reference include/kmer.hpp:343-345: K<=32 -> 1, <=64 -> 2, <=95 -> 3 */

typedef struct hsk_ctx hsk_ctx;

/* Replaces the reference's compile-time -D parameters (Makefile:39-46, compiletime.h:7-22), which
 * the C++ shim forwards here at run time. */
typedef struct hsk_config {
    int32_t k;     /* KMER_SIZE, 2 < k < 96                                   */
    int32_t m;     /* MINIMIZER_SIZE, 0 < m < k (the engine uses min(m, 32), raised so that k - m + 1 <= 64; the
                      result does not depend on the minimizer)  */
    int32_t lower; /* LOWER_KMER_FREQ                                          */
    int32_t upper; /* UPPER_KMER_FREQ, lower <= upper <= 65535                 */
    int32_t ext;   /* EXTENSION: 1 = carry (ReadId, PosInRead) per occurrence  */
    int32_t device;           /* CUDA device ordinal                            */
    int32_t rank, nranks;     /* position in the job; replaces MPI_Comm_rank/size (hysortk.cpp:41-44) */
    const void *nccl_id;      /* HSK_NCCL_ID_BYTES from hsk_get_unique_id on rank 0, broadcast by the
                                 caller (MPI_Bcast / torch.distributed); NULL when nranks == 1 */
    int32_t buckets_per_rank; /* minimizer-hash buckets owned by each rank (reference: tasks per
                                 rank, kmerops.cpp:40-43); 0 = default */
    uint64_t batch_kmers;     /* k-mers expanded + sorted at once; 0 = default (fits HBM) */
    void *stream;             /* cudaStream_t to run on; NULL = the context's own stream */
} hsk_config;

/* Per-stage device time (CUDA events on the context stream, milliseconds) and algorithmic volume
 * of the last hsk_count call; what `roofline` in bench.py is computed from. */
typedef struct hsk_stats {
    uint64_t n_kmers_local;   /* k-mers extracted from this rank's reads                     */
    uint64_t n_kmers_owned;   /* k-mers this rank sorted after the exchange                  */
    uint64_t n_supermers;     /* supermers produced locally                                  */
    uint64_t supermer_bytes;  /* bytes of supermer records produced locally (wire volume)    */
    uint64_t bytes_sent;      /* supermer bytes sent to other ranks                          */
    uint64_t bytes_received;
    uint64_t n_batches;
    uint64_t n_sort_passes;   /* radix passes executed per batch                             */
    uint64_t n_launches;      /* kernels launched                                            */
    float ms_h2d, ms_extract, ms_exchange, ms_expand, ms_sort, ms_count, ms_d2h, ms_total; /* expand/sort/count: HBM path only */
    float ms_sort_passes;     /* part of ms_sort spent in the per-digit pass kernels (n_sort_passes * n_batches launches) */
    float ms_bins;            /* fused on-chip expand + sort + count kernel (stages 4+5)     */
    uint64_t n_overflow_bins; /* bins too large / too skewed for the on-chip path, handled through HBM */
} hsk_stats;

/* Result of one rank.  Arrays live in page-locked host memory owned by the context and stay valid
 * until the next hsk_count / hsk_destroy on it.  Entry i: k-mer words kmer_words[i*nwords + w]
 * (word 0 = bases 0..31, 2 bits per base from the most significant bit; reference
 * include/kmer.hpp:165-185), count cnt[i].  Entries are grouped by minimizer bin, bins in index order
 * (bins that went through the HBM path follow at the end), and ascend by k-mer inside a bin
 * (reference: per-task sorted runs, kmerops.cpp:883-904); the order is deterministic.
 * With ext, the occurrences of entry i are pos/rid[occ_off[i] .. occ_off[i+1]) (reference
 * KmerListEntryS::pos/rid, kmer.hpp:383-400).  histogram[c] = number of kept k-mers with count c
 * on this rank, upper+1 bins (hysortk.cpp:106-113). */
typedef struct hsk_result {
    int32_t nwords;
    uint64_t n_kept;
    uint64_t n_occ;
    const uint64_t *kmer_words;
    const uint32_t *cnt;
    const uint64_t *occ_off;
    const uint32_t *pos;
    const int32_t *rid;
    const uint64_t *histogram;
    hsk_stats stats;
} hsk_result;

/* Result left on the device (for callers that keep working on the GPU, and for timing the
 * HBM-resident path): same layout, device pointers, valid until the next call on the context. */
typedef struct hsk_device_result {
    int32_t nwords;
    uint64_t n_kept;
    uint64_t n_occ;
    const uint64_t *d_kmer_words; /* entry i, word w at d_kmer_words[i*nwords + w] */
    const uint32_t *d_cnt;
    const uint64_t *d_occ_off;
    const uint32_t *d_pos;
    const int32_t *d_rid;
    const uint64_t *d_histogram;
    hsk_stats stats;
} hsk_device_result;

const char *hsk_last_error(void);
int hsk_version(void);

/* ncclGetUniqueId for the collectives of the supermer exchange (bin totals, barrier, histogram; the
 * supermers themselves are read in place from the peers' memory, or shipped with ncclSend/ncclRecv when
 * HSK_EXCHANGE=nccl) — replaces the MPI communicator of exchange_supermer, kmerops.cpp:130-195. */
int hsk_get_unique_id(void *id_out /* HSK_NCCL_ID_BYTES */);

int hsk_create(hsk_ctx **ctx, const hsk_config *cfg);
void hsk_destroy(hsk_ctx *ctx);

/* kmer_count on host buffers — what hysortk::kmer_count(const DnaBuffer&, MPI_Comm) binds to.
 *   packed     DnaBuffer bytes (reference include/dnabuffer.hpp:14-47, src/dnaseq.cpp:9-31): reads
 *              back to back, each starting on a fresh byte, 4 bases per byte, first base in bits 7..6
 *   nbytes     DnaBuffer::getbufsize()
 *   read_len   DnaSeq::size() of every read, nreads entries
 *   readid_base  number of reads on lower ranks (the MPI_Exscan of kmerops.cpp:65-70)
 * Collective over the ranks of the context (every rank must call it).  H2D/D2H copies are part
 * of the call. */
int hsk_count(hsk_ctx *ctx, const uint8_t *packed, uint64_t nbytes, const uint64_t *read_len, uint64_t nreads,
              int32_t readid_base, hsk_result *out);

/* kmer_count on host buffers with the result handed over while it is being produced — what the C++ shim uses to build the
 * reference's KmerListS (std::vector<KmerListEntryS>, include/kmer.hpp:368-410; reference kmerops.cpp:883-904 copies the
 * task results into it at the end) while the GPU is still counting.  `sink` is called from the host threads of the
 * context — several calls may run at the same time, in any order — with disjoint parts [first_entry, first_entry +
 * n_entries) of the result that have reached the host; together they cover [0, n_kept) exactly.  `view` has the
 * hsk_result layout and is indexed with absolute entry / occurrence numbers (pointers valid during the call only; the
 * occurrences of the part are [first_occ, first_occ + n_occ)); `total_hint` is an estimate of the final number of entries
 * that does not fall short by more than a few percent (exact in the part that is produced last).  A non-zero return of
 * the sink stops the deliveries and makes hsk_count_stream fail.  All calls of the sink have returned when
 * hsk_count_stream returns; *out is filled as by hsk_count, so a sink may also leave parts to its caller.  `packed` may be pageable memory (a DnaBuffer is a plain heap array, reference
 * include/dnabuffer.hpp:40): it is then copied through a page-locked staging ring by the context's host threads
 * (HSK_HOST_THREADS, default: the hardware threads divided by the ranks, at most 8), chunk by chunk under the extraction. */
typedef int (*hsk_sink_fn)(void *user, const hsk_result *view, uint64_t first_entry, uint64_t n_entries, uint64_t first_occ,
                           uint64_t n_occ, uint64_t total_hint);
int hsk_count_stream(hsk_ctx *ctx, const uint8_t *packed, uint64_t nbytes, const uint64_t *read_len, uint64_t nreads,
                     int32_t readid_base, hsk_sink_fn sink, void *user, hsk_result *out);

/* Same with the reads already resident in HBM: d_packed (nbytes, 16-byte aligned, readable up to
 * nbytes rounded up to 16), d_read_off (nreads+1 byte offsets of the reads, uint64) and d_read_len
 * (nreads, uint32) are device pointers.  The result stays on the device. */
int hsk_count_device(hsk_ctx *ctx, const uint8_t *d_packed, uint64_t nbytes, const uint64_t *d_read_off,
                     const uint32_t *d_read_len, uint64_t nreads, int32_t readid_base, hsk_device_result *out);

/* Copies the last device result of the context to page-locked host arrays (hsk_result layout). */
int hsk_fetch_result(hsk_ctx *ctx, hsk_result *out);

/* Page-locks (cudaHostRegister, portable) / releases caller memory that will be handed to hsk_count repeatedly or is large:
 * page-locked input goes to the GPU from where it is, pageable input is copied through the context's staging ring first.
 * read_dna_buffer does this for the DnaBuffer it returns (the reference's FastaIndex::getmydna allocates it with new[],
 * src/fastaindex.cpp:204-305).  `device`: the CUDA device of this rank.  Return 0 on success; failure (no GPU, no
 * permission to lock that much memory) is harmless: the memory simply stays pageable. */
int hsk_host_register(void *p, size_t bytes, int32_t device);
int hsk_host_unregister(void *p);

/* Sum of the per-rank histograms over all ranks (replaces the MPI_Allreduce of
 * hysortk.cpp:104,115); hist has upper+1 bins.  Collective. */
int hsk_allreduce_histogram(hsk_ctx *ctx, uint64_t *hist);

/* Fills caller memory laid out as the reference's EXTENSION==0 KmerListEntryS array
 * ({uint64_t kmer[nwords]; uint64_t cnt;}, include/kmer.hpp:368-382) from the last result. */
int hsk_fill_entries(hsk_ctx *ctx, void *entries, uint64_t capacity_entries);

/* ---- stage-level entry points (tests, profiling); device pointers, run on the context stream ---- */

/* LSD radix sort of n keys held as nwords planes (plane 0 most significant) with an optional
 * 64-bit payload plane, over the significant bits of a k-mer of size k.  Result in the same
 * planes (tmp planes are scratch of the same size). */
int hsk_debug_sort(hsk_ctx *ctx, uint64_t *const *d_keys, uint64_t *const *d_tmp, uint64_t *d_val, uint64_t *d_val_tmp,
                   uint64_t n, int32_t nwords, int32_t k);

/* Extraction + binning only: returns the local supermer slots (host copies) so tests can check that the
 * supermers of every bin re-expand to exactly the k-mers of the input.  A slot is slot_words 32-bit words:
 * bases 16 per word from the top bits; the last payload word carries 12 bases in its upper 24 bits and the
 * length in bases in its low 8 bits; with ext two more words follow: PosInRead of the first base, ReadId.
 * Payload words = slot_words - (ext ? 2 : 0). */
typedef struct hsk_supermers {
    uint64_t n_bins;
    const uint64_t *bin_slots;   /* slots per bin                      */
    const uint64_t *bin_kmers;   /* k-mers per bin                     */
    uint64_t n_slots;
    uint32_t slot_words;
    const uint32_t *slots;       /* n_slots * slot_words, bin-major    */
} hsk_supermers;
int hsk_debug_extract(hsk_ctx *ctx, const uint8_t *packed, uint64_t nbytes, const uint64_t *read_len, uint64_t nreads,
                      int32_t readid_base, hsk_supermers *out);

#ifdef __cplusplus
}
#endif
#endif /* HSK_CAPI_H_ */
