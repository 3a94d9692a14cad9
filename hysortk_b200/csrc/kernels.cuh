// Launch interfaces of the hysortk_b200 kernels (one .cu per stage).
#pragma once
#include "common.cuh"

namespace hsk {

// ---- read table: reads.cu ---------------------------------------------------------------------------
size_t read_table_scratch_bytes(u64 nreads);
// read_off[nreads+1] / read_len[nreads] from 64-bit lengths; flags bit 0: a read longer than 2^32-1 bases, bit 1: the
// lengths do not add up to nbytes
cudaError_t launch_read_table(const u64 *len64, u64 nreads, u64 nbytes, u64 *read_off, u32 *read_len, u64 *scratch, u32 *flags,
                              cudaStream_t s);

constexpr int BN_MAX_SRC = 16;       // ranks (sources of a bin's supermers / destinations of the scatter)

// ---- stage 1+2: extract.cu -----------------------------------------------------------------------
struct ExtractParams {
    const u8 *packed;        // DnaBuffer bytes on the device, 16-byte aligned
    u64 nbytes;              // DnaBuffer::getbufsize()
    u64 nbytes_padded;       // readable extent (multiple of 16)
    const u64 *read_off;     // nreads+1 byte offsets (read_off[nreads] = nbytes)
    const u32 *read_len;     // nreads lengths in bases
    u64 nreads;
    u64 ntiles;              // warp tiles of out_slots k-mer start slots each
    u64 tile_begin, tile_end;   // tiles of this launch
    u32 tiles_per_warp;
    u32 out_slots;           // xt_out_slots(k - m + 1)
    const u32 *tile_read;    // ntiles+1: read holding the first byte of every tile (k_tile_reads)
    int k, m;                // m already clamped (m <= 32, k - m + 1 <= XT_WMAX)
    u32 nbins;               // all ranks' bins
    u32 slot_nmax;           // k-mers per supermer slot: slot_max_bases - k + 1
    u32 slot_ninv;           // ceil(2^32 / slot_nmax)
    int readid_base;
    u32 *out_stream;         // pass B destination: the bin-major supermer stream of this rank
};

constexpr int XT_META = 8;   // u64 words after the bin starts: run cursor, k-mer total, pass A's independent slot / k-mer totals
// grid of the two extraction passes (persistent warps, contiguous tile ranges)
u32 extract_grid(int w, int sm_count);
// tile_read[t] = read holding byte t * out_slots / 4 (t = 0..ntiles)
cudaError_t launch_tile_reads(const ExtractParams &P, u32 *tile_read, cudaStream_t s);
// pass A: bin_tot[b] += bt_pack(slots, k-mers) per run; run list + tile headers for pass B
cudaError_t launch_supermer_count(const ExtractParams &P, u32 nctas, u64 *bin_tot, u64 *run_list, ulonglong2 *tile_hdr,
                                  u64 *run_cursor, u64 run_capacity, cudaStream_t s);
// bin_start: nbins+1 exclusive prefix of the slot counts, bin_cursor[b] = bin_start[b]; *kmers_total += all k-mers
size_t bin_scan_scratch_bytes(u32 nbins);
cudaError_t launch_bin_scan(const u64 *bin_tot, u32 nbins, u64 *bin_start, u64 *bin_cursor, u64 *kmers_total, u64 *scratch,
                            cudaStream_t s);
// pass B: bin_cursor hands out slot indices inside P.out_stream
cudaError_t launch_supermer_scatter(const ExtractParams &P, u32 nctas, int nwords, bool ext, const u64 *run_list,
                                    const ulonglong2 *tile_hdr, u64 *bin_cursor, cudaStream_t s);

// ---- stage 4 (HBM path): expand.cu -------------------------------------------------------------------
constexpr int XP_THREADS = 256;
constexpr int XP_SPT = 4;                          // slots per thread in the scans
constexpr int XP_TILE = XP_THREADS * XP_SPT;       // 1024 slots per tile

struct ExpandSegment {
    const u32 *slots;        // nslots supermer slots of sw words
    u64 nslots;
    u64 out_base;            // first output k-mer index of the segment inside the batch
};

// scratch: tile_sums / tile_base hold ceil(nslots/XP_TILE) entries each
cudaError_t launch_expand(const ExpandSegment &seg, int k, int nwords, bool ext, u32 *tile_sums, u64 *tile_base,
                          Planes out_keys, u64 *out_val, cudaStream_t s);

// ---- stages 4+5 on chip: bins.cu ---------------------------------------------------------------------
constexpr int BN_MAX_CTAS = 4;          // CTAs of k_bin_count per SM, at most
constexpr int BN_SORTCAP = 2048;       // most kept k-mers of a bin that the CTA sorts itself
constexpr int BN_DDLIMIT_MAX = 2048;   // entries of a CTA's list of distinct supermers (bins with more slots skip the de-duplication)

struct BinParams {
    int k;
    u32 lower, upper;
    u32 nbins;                           // owned bins, local index 0..nbins-1
    int nsrc;
    const u32 *slots[BN_MAX_SRC];        // supermer slot stream per source rank
    const u64 *seg_start[BN_MAX_SRC];    // nbins+1: first slot of every bin inside the source's stream
    const u64 *bin_kmers;                // nbins: bt_pack(slots, k-mers) per bin over all sources; the k-mer field is used
    // final arena (bins in index order, ascending k-mers inside a bin)
    u64 *out_words; u32 *out_cnt; u64 *out_occ_off; u32 *out_pos; int *out_rid;
    u64 *histogram;
    u64 *cursor;                         // [0] entries, [1] occurrences in the arena once every bin is done
    u64 arena_cap, occ_cap;              // entries / occurrences the arena holds; a bin that would run past them sets *err
    u32 *err;                            // zeroed; bit 0: the arena is too small for the result
    u64 *lb_state;                       // 2 x nbins, zeroed: look-back cells (entries, occurrences)
    u32 *ticket;                         // zeroed
    u32 *ovf_list, *ovf_count;           // bins left to the HBM path
    // bins that keep more k-mers than a CTA sorts itself: unsorted in the staging area, listed for the big gather
    u64 *st_words; u32 *st_cnt; u32 *st_pos; int *st_rid;
    u64 *stage_cursor;                   // [0] entries, [1] occurrences claimed so far (zeroed)
    u64 stage_cap, stage_occ_cap;        // room in the staging area; a bin that finds none goes to the HBM path
    u64 *bin_rec;                        // nbins x {stage entry base, kept, stage occurrence base, occurrences}
    u64 *fin;                            // nbins x {final entry base, final occurrence base}
    u32 *big_list, *big_count;
    // groups of group_bins consecutive bins: completion is reported to the host, which streams the arena out
    u32 group_bins;
    u32 *grp_done, *grp_big;             // per group, zeroed: finished bins, bins left to the big gather
    u64 *grp_end;                        // per group: arena cursor (entries, occurrences) after its last bin
    u64 *snap;                           // page-locked host memory or null: per group {entries, occurrences, big bins, ready}
    // per CTA: list of the distinct supermer slots of the bin it works on and their weights (K <= 64 without EXTENSION)
    uint4 *dd_slots; u32 *dd_mult;
    // per CTA: the sorted (k-mer, count) entries of the bin it finished last, until their place in the arena is resolved
    // (without EXTENSION): BN_SORTCAP entries each
    u64 *pend_words; u32 *pend_cnt;
    u32 walk_split, walk_min;            // the walk cuts a bin into about walk_split batches per warp, of at least walk_min slots
};

size_t bin_dedup_scratch_bytes(int sm_count, int slot_words);   // dd_slots + dd_mult of every resident CTA
size_t bin_pending_scratch_bytes(int sm_count, int nwords);    // pend_words + pend_cnt of every resident CTA (3 per SM at most)
int bin_target_kmers(int nwords, bool ext);  // k-mer occurrences per bin the on-chip path is sized for
// k_bin_count: every bin counted, sorted and written to the arena (or listed for the big gather / the HBM path)
cudaError_t launch_bin_count(const BinParams &P, int nwords, bool ext, int sm_count, cudaStream_t s);
// k_bin_gather over big_list; needed only when *big_count != 0
cudaError_t launch_bin_gather_big(const BinParams &P, int nwords, bool ext, int sm_count, cudaStream_t s);
// multi-rank, from the all-gathered bin totals alltot[src][bin]:
//   seg_start[src][0..tg]  first slot of every owned bin inside the bin-major supermer stream of rank src: absolute
//                          (index in src's own buffer, read in place over NVLink) or relative to the first owned bin
//                          (index in the region received from src)
//   meta[src]              slots of the owned bins in src's stream;  bin_kmers / owned_total: k-mers per owned bin / in all
cudaError_t launch_seg_scan(const u64 *alltot, u32 T, int me, u32 tg, int nranks, bool absolute, u64 *seg_start, u64 *meta,
                            u64 *bin_kmers, u64 *owned_total, cudaStream_t s);

// ---- stage 5a: radix.cu ------------------------------------------------------------------------------
// scratch layout (u32 units): [RS_MAX_PASSES*256 bins][RS_MAX_PASSES tile counters][ntiles*256 look-back]
size_t radix_scratch_bytes(u64 n);
// Sorts n keys (planes `a`, optional payload va) using planes `b`/vb as the other half of the
// ping-pong.  On return *result_in_b tells where the sorted data is.
// ev_pass0/ev_pass1 (optional) are recorded around the per-digit pass kernels.
cudaError_t launch_radix_sort(Planes a, Planes b, u64 *va, u64 *vb, u64 n, int nwords, int k, void *scratch,
                              bool *result_in_b, int *npasses, int *nlaunches, cudaStream_t s,
                              cudaEvent_t ev_pass0 = nullptr, cudaEvent_t ev_pass1 = nullptr);

// ---- stage 5b/5c: count.cu ---------------------------------------------------------------------------
constexpr int CT_THREADS = 256;
constexpr int CT_IPT = 8;
constexpr int CT_TILE = CT_THREADS * CT_IPT;       // 2048 sorted k-mers per tile

struct CountParams {
    Planes keys;             // sorted keys of the batch
    const u64 *val;          // payload plane (ext) or null
    u64 n;
    int nwords;
    u32 lower, upper;
    // result arena (shared by all batches of a call)
    u64 *out_words;          // AoS: entry e -> out_words[e*nwords + w]
    u32 *out_cnt;
    u64 *out_occ_off;        // ext: exclusive occurrence offsets per entry
    u32 *out_pos;            // ext: PosInRead per occurrence
    int *out_rid;            // ext: ReadId per occurrence
    u64 *histogram;          // upper+1 bins
    u64 *cursor;             // [0] = entries emitted so far, [1] = occurrences emitted so far
    u64 arena_cap, occ_cap;  // room in the arena; what would run past it is not written and *err gets bit 0
    u32 *err;
};

// scratch: 2 * (ceil(n/CT_TILE)+1) u64
size_t count_scratch_bytes(u64 n);
cudaError_t launch_count_filter(const CountParams &P, void *scratch, cudaStream_t s);

} // namespace hsk
