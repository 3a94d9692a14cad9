// Stages 4+5 fused, on chip: one CTA takes one bin of supermers, expands it to canonical k-mers and
// counts them in shared memory — the k-mers themselves never touch HBM.
//
// Replaces, for bins that fit in shared memory, the reference's HOT LOOPS C+D+E:
//   receive_from_buffer_stage2 + GetRepKmers  (src/kmerops.cpp:484-521, include/kmer.hpp:313-340)
//   sort_task -> RADULS / PARADIS             (src/kmerops.cpp:1382-1407)
//   count_sorted_kmers                        (src/kmerops.cpp:1410-1445), histogram (hysortk.cpp:106-113)
//
// The reference sorts every k-mer occurrence and then run-length counts.  At 30x coverage a bin of
// ~8000 occurrences holds only ~270 genomic k-mers x 30 copies plus error singletons, and all of
// them are overlapping windows of a few loci (they share minimizers): sorting the occurrences is both
// unbalanced (30-copy lumps) and wasted work.  Here, k_bin_count (one CTA per bin, bins taken in index
// order through a ticket, every CTA resident):
//   0. (K <= 64 without EXTENSION) identical supermer slots of the bin are merged first (dedup_bin): a supermer
//      that was read 20 times is expanded once, with weight 20; the k-mers of the distinct slots are then dealt
//      out to the warps in equal shares (a share may begin and end inside a slot)
//   1. walk: a warp stages batches of up to 32 supermer slots; a warp scan of the k-mers per slot and a bitmap of
//      the slot starts map k-mer g of the batch to (slot, offset), so that in every round each lane extracts ONE
//      k-mer directly from the staged slot words (funnel shift, reverse complement by bit reversal, canonical
//      choice) — no per-thread walk, no divergence on supermer boundaries
//   2. count: the k-mer is inserted into an open-addressing table in shared memory: for K <= 32 the cell is the
//      k-mer itself (one 64-bit CAS, and only when a plain load did not already find it); for K > 32 a
//      32-bit fingerprint cell guards the full key words (claim = EMPTY -> LOCK -> words -> fingerprint), and
//      every fingerprint match is confirmed on the full words, so the table is exact.  A 32-bit counter per
//      slot counts the occurrences; the occurrence that lifts a counter to LOWER lists the slot as a candidate
//   3. filter: candidates with LOWER <= count <= UPPER are kept; the bin publishes how many (look-back cell)
//   4. sort: the kept k-mers of the bin in ascending order (sort_bin: one counting pass over the most
//      significant differing byte + rank inside the bucket; only slot numbers move)
//   5. emit: without EXTENSION the sorted (k-mer, count) entries are parked in the CTA's scratch (stash_bin) and
//      move to their place in the arena — known from a decoupled look-back over the bins before — while the
//      CTA already works on its next bin (flush_pending), so that nobody waits for the slowest bin in flight.
//      With EXTENSION the place is resolved at once (emit_now), the counters of the kept slots become cursors
//      into the bin's occurrence area and a second walk places (pos, rid) of every occurrence of a kept k-mer
//   6. the CTA that completes a group of bins reports the arena cursor to page-locked host memory: the host
//      streams that part of the result out while the kernel keeps counting (engine.cu)
// A bin that keeps more than BN_SORTCAP k-mers goes unsorted to a staging area and is sorted + moved by
// k_bin_gather afterwards (or takes the HBM path when the staging area is full).
//
// So the SORT is still there, but it runs over the distinct kept k-mers (D) instead of over every
// occurrence (N): D/N is ~4 % at 30x coverage with 1 % errors.  The arena holds the bins in index order
// and ascending k-mers inside each bin (the reference: per-task sorted runs in task order,
// kmerops.cpp:883-904), deterministically.  A bin of any number of occurrences is handled as long as its
// distinct k-mers fit the table; bins are sized by the extraction stage so that they fill about a third of
// it, and the rare bin that does not fit (skew) is reported in an overflow list and goes through the HBM
// path (expand.cu -> radix.cu -> count.cu).
#include "kernels.cuh"

#include <algorithm>
#include <cstdlib>

namespace hsk {

constexpr int BN_HCAP = 128;                   // shared-memory histogram bins (larger counts go to the global histogram)
constexpr u64 BN_EMPTY = ~0ull;                // never a canonical k-mer: a K-mer of all T is not canonical
constexpr u32 BN_EMPTY32 = 0xFFFFFFFFu;        // K > 32: state of a fingerprint cell
constexpr u32 BN_LOCK32 = 0xFFFFFFFEu;         //          claimed, key words not yet written
constexpr u32 BN_NOTKEPT = 0xFFFFFFFFu;
constexpr int BN_CAND = 1024;                  // candidate list; a bin with more candidates scans its whole table

// TH = threads per CTA: 512 (16 warps; two or three CTAs per SM for K <= 32), or 1024 for the one-CTA-per-SM tables of
// K > 32 (32 warps on the SM instead of 16)
template <int NW, bool EXT, int TH>
struct BinCfg {
    static constexpr int THREADS = TH, WARPS = TH / 32;
    static constexpr int SW = (NW == 1 ? 4 : 8) + (EXT ? 4 : 0);
    static constexpr int PW = SW - (EXT ? 2 : 0);
    // table slots: K <= 32: 8192 (two CTAs per SM), with EXTENSION 4096 (three CTAs); K > 32: one CTA per SM
    static constexpr int TS_BITS = NW == 1 ? (EXT ? 12 : 13) : (NW == 2 && !EXT ? 13 : 12);
    static constexpr int TS = 1 << TS_BITS;
    static constexpr int SLOTS_PT = TS / TH;
    static constexpr int CTAS = NW == 1 ? (EXT ? 3 : 2) : 1;
    // most k-mers one slot can hold (smallest K of the word count) = rounds of 32 k-mers a batch of 32 slots can need
    static constexpr int NMAX = 16 * (PW - 1) + 12 - (NW == 1 ? 3 : (NW == 2 ? 33 : 65)) + 1;
    // slots per walk batch (a warp stages them) and kept k-mers a CTA sorts itself; a bin that keeps more goes through
    // the staging area + big gather
    static constexpr int BATCH = TH != 512 ? 16 : 32;
    static constexpr int SORTCAP = BN_SORTCAP;
    static constexpr int EPT = SORTCAP / TH;   // kept k-mers per thread in the sort
    static constexpr int HEADW = ((BATCH * NMAX + 31) / 32 + 3) / 4 * 4;   // rounds of 32 k-mers a batch can need
    static constexpr int TARGET = NW == 1 ? (EXT ? 4096 : 8192) : (NW == 2 && !EXT ? 6144 : 3072);
    // de-duplication of the supermers of a bin (K <= 64 without EXTENSION): cells of the supermer table, most slots of
    // a bin that goes through it (a multiple of TH, below the number of cells)
    static constexpr int DDTS = TH > 512 ? 4096 : 2048;
    static constexpr int DDLIMIT = TH > 512 ? 2048 : 1536;
    static constexpr int DDPT = DDLIMIT / TH;
    static_assert(DDLIMIT <= BN_DDLIMIT_MAX && DDLIMIT % TH == 0 && DDLIMIT < DDTS && DDTS % TH == 0, "supermer table geometry");
    static_assert(SORTCAP % TH == 0 && SORTCAP <= TS, "sort geometry");
};

size_t bin_pending_scratch_bytes(int sm_count, int nwords) { return (size_t)sm_count * BN_MAX_CTAS * BN_SORTCAP * ((size_t)nwords * 8 + 4); }
size_t bin_dedup_scratch_bytes(int sm_count, int slot_words) { return (size_t)sm_count * BN_MAX_CTAS * BN_DDLIMIT_MAX * ((size_t)slot_words * 4 + sizeof(u32)); }

// threads per CTA of the kernel that comes in two shapes (HSK_BIN_THREADS): K in 33..64: 1024 or 512, one CTA per SM either
// way.  (For K <= 32 four CTAs of 256 threads with 4096-slot tables and bins half as large were measured: no gain.)
static int bin_threads_env(int dflt, int other)
{
    const char *e = getenv("HSK_BIN_THREADS");
    const int t = e ? atoi(e) : dflt;
    return t == other ? other : dflt;
}

int bin_target_kmers(int nwords, bool ext)
{
    if (nwords == 1) return ext ? BinCfg<1, true, 512>::TARGET : BinCfg<1, false, 512>::TARGET;
    if (nwords == 2) return ext ? BinCfg<2, true, 512>::TARGET : BinCfg<2, false, 512>::TARGET;
    return BinCfg<3, false, 512>::TARGET;
}

// scratch of a CTA, used by the walk (staged slots, scan, slot-start bitmap of every warp) and then by the sort of
// the kept k-mers (keys in bucket order, payloads in sorted order, bucket starts)
template <int NW, bool EXT, int TH>
struct BinScratchCfg {
    using Cfg = BinCfg<NW, EXT, TH>;
    static constexpr int STG_U4 = Cfg::BATCH * Cfg::SW / 4 + 2;                       // uint4 per warp (+ pad)
    static constexpr size_t WALK = (size_t)Cfg::WARPS * (STG_U4 * 16 + 32 * 2 + Cfg::HEADW * 4);
    static constexpr size_t SORT = (size_t)Cfg::SORTCAP * 4 + 264 * 4;
    static constexpr size_t BYTES = ((WALK > SORT ? WALK : SORT) + 15) / 16 * 16;
};

template <int NW, bool EXT, int TH>
struct BinSmem {
    using Cfg = BinCfg<NW, EXT, TH>;
    using Scr = BinScratchCfg<NW, EXT, TH>;
    alignas(16) u64 fp[NW == 1 ? Cfg::TS : 1];              // K <= 32: the k-mer itself is the CAS key
    alignas(16) u64 kw[NW > 1 ? NW : 1][NW > 1 ? Cfg::TS : 1];   // K > 32: full key words ...
    alignas(16) u32 fp32[NW > 1 ? Cfg::TS : 1];             //         ... guarded by a 32-bit fingerprint cell
    alignas(16) u32 cnt[Cfg::TS];                           // occurrences per slot; EXT pass 2: next occurrence offset
    alignas(16) unsigned char scratch[Scr::BYTES];
    u32 hist[BN_HCAP];
    u16 cand[BN_CAND];                                      // slots whose counter reached LOWER
    const u32 *src_ptr[BN_MAX_SRC];                         // first slot of the bin in every source stream
    u32 src_n[BN_MAX_SRC], src_sbase[BN_MAX_SRC + 1];
    u32 wa[Cfg::WARPS], wb[Cfg::WARPS];
    u64 base_k, base_o;                                     // where the bin's entries / occurrences go
    u32 *occ_pos; int *occ_rid;                             // EXT pass 2 target arrays (arena, or staging for big bins)
    u32 bin, nk, S, bail, next_batch, seen, ncand, skip_out;
    u32 pend_valid, pend_lb, pend_tk;                       // a sorted bin waiting in the CTA's global scratch for its place in the arena
    u32 xor_hi, xor_lo;                                     // bits in which the kept k-mers' first words differ
    u16 wlo[Cfg::WARPS + 1], wo0[Cfg::WARPS + 1];           // de-duplicated bin: first (slot, k-mer in it) of every warp's share of the walk
    u32 batch_slots;                                        // slots per walk batch: 32, fewer when the bin has few slots
    int nsrc;                                               // sources of the walk: P.nsrc, or 1 when the bin was de-duplicated
    const u32 *mult;                                        // weight per slot (de-duplicated bins) or null

    // walk layout
    __device__ uint4 *stg(int warp) { return reinterpret_cast<uint4 *>(scratch) + (size_t)warp * Scr::STG_U4; }
    __device__ u16 *scan(int warp) { return reinterpret_cast<u16 *>(scratch + (size_t)Cfg::WARPS * Scr::STG_U4 * 16) + warp * 32; }
    __device__ u32 *heads(int warp)
    {
        return reinterpret_cast<u32 *>(scratch + (size_t)Cfg::WARPS * (Scr::STG_U4 * 16 + 64)) + (size_t)warp * Cfg::HEADW;
    }
    // sort layout: kept slots (then the same slots in sorted order), slots in bucket order, bucket starts
    __device__ u16 *klist() { return reinterpret_cast<u16 *>(scratch); }                                   // [SORTCAP]
    __device__ u16 *xslot() { return reinterpret_cast<u16 *>(scratch) + Cfg::SORTCAP; }                    // [SORTCAP]
    __device__ u32 *bh() { return reinterpret_cast<u32 *>(scratch + (size_t)Cfg::SORTCAP * 4); }           // [257]
};

// block-wide exclusive scan of two u32 values (TH threads); returns exclusive prefixes and totals
template <int TH>
__device__ __forceinline__ void block_scan2(u32 a, u32 b, u32 *wa, u32 *wb, u32 &ea, u32 &eb, u32 &ta, u32 &tb)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 ia = a, ib = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 x = __shfl_up_sync(0xFFFFFFFFu, ia, d);
        u32 y = __shfl_up_sync(0xFFFFFFFFu, ib, d);
        if (lane >= d) { ia += x; ib += y; }
    }
    __syncthreads();   // protects wa/wb against the previous use
    if (lane == 31) { wa[warp] = ia; wb[warp] = ib; }
    __syncthreads();
    // every warp scans the TH / 32 warp totals itself (one per lane)
    u32 x = lane < TH / 32 ? wa[lane] : 0u, y = lane < TH / 32 ? wb[lane] : 0u;
    u32 ix = x, iy = y;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 p = __shfl_up_sync(0xFFFFFFFFu, ix, d);
        u32 q = __shfl_up_sync(0xFFFFFFFFu, iy, d);
        if (lane >= d) { ix += p; iy += q; }
    }
    ta = __shfl_sync(0xFFFFFFFFu, ix, 31);
    tb = __shfl_sync(0xFFFFFFFFu, iy, 31);
    const u32 oa = __shfl_sync(0xFFFFFFFFu, ix - x, warp), ob = __shfl_sync(0xFFFFFFFFu, iy - y, warp);
    ea = oa + ia - a;
    eb = ob + ib - b;
}

template <int NW>
__device__ __forceinline__ bool key_less(const u64 (&a)[NW], const u64 (&b)[NW])
{
#pragma unroll
    for (int l = 0; l < NW; ++l) {
        if (a[l] != b[l]) return a[l] < b[l];
    }
    return false;
}

// home slot and (K > 32) fingerprint of a k-mer
template <int NW>
__device__ __forceinline__ u64 key_mix(const u64 (&w)[NW])
{
    u64 h = w[0] * 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int l = 1; l < NW; ++l) {
        h ^= h >> 29;
        h = (h + w[l]) * 0xBF58476D1CE4E5B9ull;
    }
    return h;
}

// address of slot j of the bin (j counts over the per-source segments in rank order)
template <typename SM>
__device__ __forceinline__ const u32 *slot_ptr(const SM &sm, u32 j, int sw)
{
    if (sm.nsrc == 1) return sm.src_ptr[0] + (size_t)j * (u32)sw;
    int s = 0;
    while (j >= sm.src_sbase[s + 1]) ++s;
    return sm.src_ptr[s] + (size_t)(j - sm.src_sbase[s]) * (u32)sw;
}

// Find the slot of a k-mer in the table; INSERT claims an empty slot when the k-mer is new.  Returns TS when the
// table is full (INSERT) / the k-mer is absent (lookup).
//   K <= 32: the cell holds the k-mer; one 64-bit CAS claims it.
//   K > 32:  the cell holds a 32-bit fingerprint; a claim goes EMPTY -> LOCK (CAS), key words written, fence,
//            fingerprint published, so whoever reads a fingerprint also sees the words and compares them in full:
//            equal fingerprints of different k-mers just probe on.  Exact, no second pass.
template <bool INSERT, int NW, bool EXT, int TH>
__device__ __forceinline__ u32 table_find(BinSmem<NW, EXT, TH> &sm, const u64 (&key)[NW])
{
    using Cfg = BinCfg<NW, EXT, TH>;
    const u64 hmix = key_mix<NW>(key);
    u32 slot = (u32)(hmix >> (64 - Cfg::TS_BITS));
    if (NW == 1) {
        const u64 f = key[0];
        for (int probes = 0; probes < Cfg::TS; ++probes) {
            u64 old = *reinterpret_cast<volatile u64 *>(&sm.fp[slot]);
            if (old == f) return slot;
            if (old == BN_EMPTY) {
                if (!INSERT) return (u32)Cfg::TS;
                old = atomicCAS(&sm.fp[slot], BN_EMPTY, f);
                if (old == BN_EMPTY || old == f) return slot;
            }
            slot = (slot + 1) & (Cfg::TS - 1);
        }
    } else {
        const u32 f = (u32)hmix & 0x7FFFFFFFu;
        for (int probes = 0; probes < Cfg::TS; ++probes) {
            volatile u32 *cell = &sm.fp32[slot];
            u32 old = *cell;
            if (old == BN_EMPTY32) {
                if (!INSERT) return (u32)Cfg::TS;
                old = atomicCAS(&sm.fp32[slot], BN_EMPTY32, BN_LOCK32);
                if (old == BN_EMPTY32) {
#pragma unroll
                    for (int l = 0; l < NW; ++l) sm.kw[l][slot] = key[l];
                    __threadfence_block();
                    *cell = f;
                    return slot;
                }
            }
            while (old == BN_LOCK32) old = *cell;
            if (old == f) {
                bool same = true;
#pragma unroll
                for (int l = 0; l < NW; ++l) same = same && (*reinterpret_cast<volatile u64 *>(&sm.kw[l][slot]) == key[l]);
                if (same) return slot;
            }
            slot = (slot + 1) & (Cfg::TS - 1);
        }
    }
    return (u32)Cfg::TS;
}

// ---- identical supermers of a bin are counted once, with a weight ------------------------------------------
// At 30x coverage most supermers of a bin are exact copies of another one: the same locus read again without an error
// in those ~40 bases (the scatter pass stores every supermer in its canonical orientation, so the strand does not
// matter).  Before the walk, the slots of the bin go through a small table in the (still unused) k-mer table memory:
// key = the whole slot (16 bytes for K <= 32, 32 bytes for K in 33..64), guarded by a fingerprint cell with the same
// EMPTY -> LOCK -> publish claim as the K > 32 k-mer table; value = number of copies.  The distinct slots and their
// weights are compacted into the CTA's own list in global memory (L2-resident) and the walk then expands each of them
// once.  Without EXTENSION only: with EXTENSION every occurrence needs its own (pos, rid) anyway.
//   cells: sm.cnt[0 .. DDTS)   weights: sm.cnt[DDTS .. 2 DDTS)   keys: sm.fp (K <= 32) / sm.kw (K > 32), as uint4
template <int NW, bool EXT, int TH>
__device__ __forceinline__ uint4 *dedup_keys(BinSmem<NW, EXT, TH> &sm)
{
    return NW == 1 ? reinterpret_cast<uint4 *>(sm.fp) : reinterpret_cast<uint4 *>(&sm.kw[0][0]);
}

// the slots of a thread, fetched early (they may come over NVLink) while the table is being cleared
template <int NW, bool EXT, int TH>
__device__ __forceinline__ void dedup_fetch(const BinSmem<NW, EXT, TH> &sm, u32 S,
                                            uint4 (&v)[BinCfg<NW, EXT, TH>::DDPT][BinCfg<NW, EXT, TH>::SW / 4])
{
    constexpr int SW = BinCfg<NW, EXT, TH>::SW;
#pragma unroll
    for (int i = 0; i < BinCfg<NW, EXT, TH>::DDPT; ++i) {
        const u32 j = threadIdx.x + i * TH;
        if (j < S) {
            const uint4 *sp = reinterpret_cast<const uint4 *>(slot_ptr(sm, j, SW));
#pragma unroll
            for (int x = 0; x < SW / 4; ++x) v[i][x] = __ldg(sp + x);
        }
    }
}

template <int NW, bool EXT, int TH>
__device__ __forceinline__ u32 dedup_bin(BinSmem<NW, EXT, TH> &sm, const BinParams &P, u32 S, int k,
                                         const uint4 (&vs)[BinCfg<NW, EXT, TH>::DDPT][BinCfg<NW, EXT, TH>::SW / 4])
{
    using Cfg = BinCfg<NW, EXT, TH>;
    constexpr int Q = Cfg::SW / 4;   // uint4 per slot
    constexpr int DDTS = Cfg::DDTS;
    const u32 tid = threadIdx.x;
    uint4 *dk = dedup_keys<NW, EXT, TH>(sm);
    u32 *cell = sm.cnt, *wgt = sm.cnt + DDTS;
#pragma unroll
    for (int i = 0; i < Cfg::DDPT; ++i) {
        if (tid + i * TH >= S) break;
        u64 h = 0;
#pragma unroll
        for (int x = 0; x < Q; ++x) {
            const uint4 v = vs[i][x];
            const u64 a = ((u64)v.y << 32) | v.x, b = ((u64)v.w << 32) | v.z;
            h = (h ^ (h >> 31)) * 0x94D049BB133111EBull + ((a * 0x9E3779B97F4A7C15ull) ^ (b * 0xC2B2AE3D27D4EB4Full));
        }
        h ^= h >> 29;
        h *= 0xBF58476D1CE4E5B9ull;
        const u32 f = (u32)h & 0x7FFFFFFFu;
        u32 idx = (u32)(h >> 50) & (DDTS - 1);
        while (true) {   // S <= DDLIMIT < DDTS: an empty cell always exists
            volatile u32 *c = &cell[idx];
            u32 old = *c;
            if (old == BN_EMPTY32) {
                old = atomicCAS(&cell[idx], BN_EMPTY32, BN_LOCK32);
                if (old == BN_EMPTY32) {
#pragma unroll
                    for (int x = 0; x < Q; ++x) dk[idx * Q + x] = vs[i][x];
                    __threadfence_block();
                    *c = f;
                    break;
                }
            }
            while (old == BN_LOCK32) old = *c;
            if (old == f) {
                const volatile u32 *q = reinterpret_cast<const volatile u32 *>(&dk[idx * Q]);
                bool same = true;
#pragma unroll
                for (int x = 0; x < Q; ++x) {
                    const uint4 v = vs[i][x];
                    same = same && q[4 * x] == v.x && q[4 * x + 1] == v.y && q[4 * x + 2] == v.z && q[4 * x + 3] == v.w;
                }
                if (same) break;
            }
            idx = (idx + 1) & (DDTS - 1);
        }
        atomicAdd(&wgt[idx], 1u);
    }
    __syncthreads();
    // compact the used cells into the CTA's list; the k-mers of the distinct slots are dealt out evenly to the warps
    // of the walk: warp w starts with k-mer floor(w * total / WARPS) of the list, wherever in a slot that falls
    constexpr int PER = DDTS / TH, PW = Cfg::PW, W = Cfg::WARPS;
    u32 used = 0, mask = 0, ksum = 0, nk[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        nk[i] = 0;
        if (cell[tid * PER + i] != BN_EMPTY32) {
            mask |= 1u << i; ++used;
            const u32 lw = reinterpret_cast<const u32 *>(&dk[(tid * PER + i) * Q])[PW - 1];
            nk[i] = (lw & 0xFFu) - (u32)k + 1;
            ksum += nk[i];
        }
    }
    u32 ex, kp, total, ktot;
    block_scan2<TH>(used, ksum, sm.wa, sm.wb, ex, kp, total, ktot);
    uint4 *outs = P.dd_slots + (size_t)blockIdx.x * BN_DDLIMIT_MAX * Q;
    u32 *outm = P.dd_mult + (size_t)blockIdx.x * BN_DDLIMIT_MAX;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        if ((mask >> i) & 1) {
#pragma unroll
            for (int x = 0; x < Q; ++x) __stcg(&outs[(size_t)ex * Q + x], dk[(tid * PER + i) * Q + x]);
            __stcg(&outm[ex], wgt[tid * PER + i]);
            // warps whose first k-mer lies in this slot: floor(w * ktot / W) in [kp, kp + n)
            for (u32 w = (kp * W + ktot - 1) / ktot; w < (u32)W; ++w) {
                const u32 b = (u32)(((u64)w * ktot) / W);
                if (b >= kp + nk[i]) break;
                sm.wlo[w] = (u16)ex; sm.wo0[w] = (u16)(b - kp);
            }
            kp += nk[i];
            ++ex;
        }
    }
    if (tid == 0) { sm.wlo[W] = (u16)total; sm.wo0[W] = 0; }
    return total;
}

// One pass over the k-mers of the bin.  Warps take batches of up to 32 consecutive slots (ticket in shared memory): the
// lanes stage one slot each, a warp scan of the k-mers per slot and a bitmap of the slot starts map k-mer g of
// the batch to (slot, offset) without a search, and round r gives k-mer 32r + lane to every lane — all lanes
// work on every round but the last, whatever the supermer lengths are.
//   PASS2 == false: insert + count.   PASS2 == true (EXTENSION): place (pos, rid) of the occurrences of kept k-mers.
template <bool PASS2, int NW, bool EXT, int TH>
__device__ __forceinline__ void walk_bin(BinSmem<NW, EXT, TH> &sm, const BinParams &P, int k, int padbits, u32 S)
{
    using Cfg = BinCfg<NW, EXT, TH>;
    constexpr int SW = Cfg::SW, PW = Cfg::PW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 *stg4 = sm.stg(warp);
    u32 *stg = reinterpret_cast<u32 *>(stg4);
    u16 *scan = sm.scan(warp);
    u32 *heads = sm.heads(warp);
    u32 seen = 0;
    // the slot of the NEXT batch is fetched while the current one is processed: the supermers may sit in a peer's
    // memory (NVLink latency), and even local ones come from L2 / HBM
    uint4 pre[SW / 4];
    u32 pre_m = 1;
    const u32 *const mult = sm.mult;   // de-duplicated bin: slots come from the CTA's own list (written by this kernel:
                                       // coherent loads), each with the number of copies it stands for
    // de-duplicated bin: the warp walks its own share of the k-mers of the list (it may begin and end inside a slot);
    // otherwise the warps take batches of slots by ticket
    const bool ranged = !PASS2 && mult != nullptr;
    u32 r_lo = 0, r_hi = 0, r_o0 = 0, r_o1 = 0, lim = S, next_j = 0;
    if (ranged) {
        r_lo = sm.wlo[warp]; r_o0 = sm.wo0[warp]; r_hi = sm.wlo[warp + 1]; r_o1 = sm.wo0[warp + 1];
        lim = r_hi + (r_o1 ? 1u : 0u);
        next_j = r_lo;
    }
    const u32 BS = ranged ? min((u32)Cfg::BATCH, max(1u, lim - r_lo)) : sm.batch_slots;   // slots per batch
    auto claim = [&](u32 &j0) {
        if (ranged) { j0 = next_j; next_j += BS; }
        else {
            u32 bt = 0;
            if (lane == 0) bt = atomicAdd(&sm.next_batch, 1u);
            j0 = __shfl_sync(0xFFFFFFFFu, bt, 0) * BS;
        }
        if ((u32)lane < BS && j0 + lane < lim) {
            const uint4 *sp = reinterpret_cast<const uint4 *>(slot_ptr(sm, j0 + lane, SW));
            if (mult) {
#pragma unroll
                for (int x = 0; x < SW / 4; ++x) pre[x] = __ldcg(sp + x);
                pre_m = __ldcg(mult + j0 + lane);
            } else {
#pragma unroll
                for (int x = 0; x < SW / 4; ++x) pre[x] = __ldg(sp + x);
            }
        }
    };
    u32 j0;
    claim(j0);
    while (j0 < lim) {
        u32 n = 0, my_m = 1, start = 0;
        if ((u32)lane < BS && j0 + lane < lim) {
#pragma unroll
            for (int x = 0; x < SW / 4; ++x) stg4[lane * (SW / 4) + x] = pre[x];
            const uint4 v = pre[(PW - 1) / 4];
            const u32 lw = ((PW - 1) % 4 == 3) ? v.w : ((PW - 1) % 4 == 1 ? v.y : ((PW - 1) % 4 == 2 ? v.z : v.x));
            n = (lw & 0xFFu) - (u32)k + 1;
            if (ranged) {   // the first / last slot of the warp's share may be cut
                const u32 j = j0 + lane;
                const u32 end = (j == r_hi) ? r_o1 : n;
                start = (j == r_lo) ? r_o0 : 0u;
                n = end - start;
            }
            my_m = pre_m;
        }
        claim(j0);
        u32 inc = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += t;
        }
        const u32 T = __shfl_sync(0xFFFFFFFFu, inc, 31);
        const u32 ex = inc - n;
        const u32 rounds = (T + 31) >> 5;
        scan[lane] = (u16)(ex - start);   // k-mer g of the batch is k-mer g - scan (mod 2^16) of its slot
        for (u32 r = lane; r < rounds && r < (u32)Cfg::HEADW; r += 32) heads[r] = 0;
        __syncwarp();
        if (n) atomicOr(&heads[ex >> 5], 1u << (ex & 31));
        __syncwarp();
        u32 base = 0;   // slots started before this round
        for (u32 r = 0; r < rounds; ++r) {
            const u32 word = heads[r];
            const u32 g = 32 * r + lane;
            const u32 s = base + __popc(word & (0xFFFFFFFFu >> (31 - lane))) - 1;
            base += __popc(word);
            const bool act = g < T;
            const u32 o = act ? ((g - scan[s]) & 0xFFFFu) : 0u;
            const u32 *w = stg + s * SW;
            const u32 wi = o >> 4, sh = 2 * (o & 15);
            u64 key[NW];
#pragma unroll
            for (int l = 0; l < NW; ++l) {
                const u32 a = w[wi + 2 * l], b = w[wi + 2 * l + 1], c = w[wi + 2 * l + 2];
                key[l] = ((u64)__funnelshift_l(b, a, sh) << 32) | __funnelshift_l(c, b, sh);
            }
            if (padbits) key[NW - 1] &= ~0ull << padbits;
            kmer_canonical<NW>(key, k);
            if (!PASS2) {
                u32 slot = (u32)Cfg::TS;
                if (act) slot = table_find<true>(sm, key);
                __syncwarp();   // the probe loop diverges; everything after it runs once per warp
                const u32 m1 = mult ? __shfl_sync(0xFFFFFFFFu, my_m, s & 31) : 1u;   // copies of the k-mer's slot
                if (act) {
                    if (slot < (u32)Cfg::TS) {
                        // the occurrence that lifts a counter to LOWER makes its slot a candidate for the output
                        const u32 before = atomicAdd(&sm.cnt[slot], m1);
                        if (before < P.lower && before + m1 >= P.lower) {
                            const u32 ci = atomicAdd(&sm.ncand, 1u);
                            if (ci < (u32)BN_CAND) sm.cand[ci] = (u16)slot;
                        }
                    } else atomicOr(&sm.bail, 16u);
                }
            } else {
                u32 slot = (u32)Cfg::TS;
                if (act) slot = table_find<false>(sm, key);
                __syncwarp();
                if (slot < (u32)Cfg::TS && *reinterpret_cast<volatile u32 *>(&sm.cnt[slot]) != BN_NOTKEPT) {
                    const u64 p = sm.base_o + atomicAdd(&sm.cnt[slot], 1u);
                    sm.occ_pos[p] = w[SW - 2] + o;
                    sm.occ_rid[p] = (int)w[SW - 1];
                }
            }
        }
        seen += mult ? __reduce_add_sync(0xFFFFFFFFu, n * my_m) : T;
        __syncwarp();   // the staging area is reused by the next batch
    }
    if (!PASS2 && lane == 0 && seen) atomicAdd(&sm.seen, seen);
}

// ---- position of a bin in the arena: decoupled look-back over the bins -------------------------------
// Bins are taken in index order by CTAs that are all resident, so every bin before `lb` has been started and
// publishes the number of entries it keeps (AGG) before it waits for anybody; a bin adds up the aggregates
// behind it until it meets an inclusive prefix (INC) and then publishes its own.
constexpr u64 LB_AGG = 1ull << 62, LB_INC = 2ull << 62, LB_VAL = (1ull << 62) - 1;

__device__ __forceinline__ u64 warp_sum64(u64 v)
{
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}

// exclusive prefix over the bins before lb (whole warp; same value in every lane)
__device__ __forceinline__ u64 lookback(volatile u64 *st, u32 lb)
{
    const int lane = threadIdx.x & 31;
    u64 sum = 0;
    for (long long j = (long long)lb; j > 0; j -= 32) {
        const long long idx = j - 1 - lane;
        u64 v = LB_INC;   // before bin 0: inclusive 0
        if (idx >= 0) { do { v = st[idx]; } while ((v >> 62) == 0); }
        const u32 inc = __ballot_sync(0xFFFFFFFFu, (v >> 62) == 2);
        u64 val = v & LB_VAL;
        if (inc) {
            if (lane > __ffs(inc) - 1) val = 0;   // nothing behind the nearest inclusive prefix
            sum += warp_sum64(val);
            break;
        }
        sum += warp_sum64(val);
    }
    return sum;
}

// look-back of one bin (warp 0): arena position -> sm.base_k / sm.base_o, inclusive prefix published.  A bin that
// would run past the arena is not written (sm.skip_out) and the call fails (P.err).
template <int NW, bool EXT, int TH>
__device__ __forceinline__ void resolve_position(BinSmem<NW, EXT, TH> &sm, const BinParams &P, u32 lb, u32 tk, u32 to)
{
    if (threadIdx.x < 32) {
        volatile u64 *lbk = P.lb_state, *lbo = P.lb_state + P.nbins;
        const u64 exk = lookback(lbk, lb);
        const u64 exo = EXT ? lookback(lbo, lb) : 0;
        if (threadIdx.x == 0) {
            lbk[lb] = LB_INC | (exk + tk);
            if (EXT) lbo[lb] = LB_INC | (exo + to);
            sm.base_k = exk; sm.base_o = exo;
            const bool over = exk + tk > P.arena_cap || (EXT && exo + to > P.occ_cap);
            sm.skip_out = over ? 1u : 0u;
            if (over) atomicOr(P.err, 1u);
            const u32 g = lb / P.group_bins;
            if (lb + 1 == P.nbins || (lb + 1) % P.group_bins == 0) { P.grp_end[2 * g] = exk + tk; P.grp_end[2 * g + 1] = exo + to; }
            if (lb + 1 == P.nbins) { P.cursor[0] = exk + tk; P.cursor[1] = exo + to; }
        }
    }
    __syncthreads();
}

// ---- the sort of a bin: the kept k-mers in ascending order ------------------------------------------------
// The kept slots of the bin are listed in sm.klist()[0 .. tk), tk <= SORTCAP.  Their k-mers are distinct, so the place of
// a k-mer in the sorted bin is the number of smaller ones.  One counting pass over the most significant byte in which
// the k-mers differ at all (first word) splits them into 256 buckets in ascending order; inside its bucket (a handful
// of k-mers) every k-mer counts the smaller ones directly.  Five barriers and ~tk / 256 comparisons per k-mer, against
// the 45 compare-exchange stages of a bitonic network over 512 elements.  Only slot numbers move: the k-mers stay in
// the table.  Result: sm.klist()[e] = slot of the e-th smallest kept k-mer.
template <int NW, bool EXT, int TH>
__device__ __forceinline__ u64 slot_key0(const BinSmem<NW, EXT, TH> &sm, u32 slot) { return NW == 1 ? sm.fp[slot] : sm.kw[0][slot]; }

template <int NW, bool EXT, int TH>
__device__ __forceinline__ void sort_bin(BinSmem<NW, EXT, TH> &sm, u32 tk)
{
    using Cfg = BinCfg<NW, EXT, TH>;
    constexpr int EPT = Cfg::EPT;
    const u32 tid = threadIdx.x, lane = tid & 31;
    u32 *bh = sm.bh();
    u16 *klist = sm.klist(), *xslot = sm.xslot();
    for (u32 i = tid; i < 257; i += TH) bh[i] = 0;
    u32 slot[EPT];
    const u64 k0 = slot_key0(sm, klist[0]);
    u64 x = 0;
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        const u32 e = tid + r * TH;
        slot[r] = e < tk ? klist[e] : 0u;
        if (e < tk) x |= slot_key0(sm, slot[r]) ^ k0;
    }
    {
        const u32 xh = __reduce_or_sync(0xFFFFFFFFu, (u32)(x >> 32)), xl = __reduce_or_sync(0xFFFFFFFFu, (u32)x);
        if (lane == 0) { if (xh) atomicOr(&sm.xor_hi, xh); if (xl) atomicOr(&sm.xor_lo, xl); }
    }
    __syncthreads();   // the slot list has been read, the bucket counters are zero
    const u64 diff = ((u64)sm.xor_hi << 32) | sm.xor_lo;
    const int top = diff ? 63 - __clzll((long long)diff) : 7;   // highest bit in which two of the k-mers differ
    const int shift = top > 7 ? top - 7 : 0;
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        if (tid + r * TH < tk) {
            const u32 b = (u32)(slot_key0(sm, slot[r]) >> shift) & 255u;
            slot[r] |= atomicAdd(&bh[b], 1u) << 16;   // place inside the bucket (< SORTCAP <= 65535)
        }
    }
    __syncthreads();
    if (tid < 32) {   // exclusive scan of the 256 bucket sizes; bh[256] = tk
        u32 v[8], sum = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[i] = bh[tid * 8 + i]; sum += v[i]; }
        u32 inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += t;
        }
        u32 ex = inc - sum;
#pragma unroll
        for (int i = 0; i < 8; ++i) { bh[tid * 8 + i] = ex; ex += v[i]; }
        if (tid == 31) bh[256] = ex;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        if (tid + r * TH < tk) {
            const u32 sl = slot[r] & 0xFFFFu;
            const u32 b = (u32)(slot_key0(sm, sl) >> shift) & 255u;
            xslot[bh[b] + (slot[r] >> 16)] = (u16)sl;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        if (tid + r * TH < tk) {
            const u32 sl = slot[r] & 0xFFFFu;
            u64 key[NW];
#pragma unroll
            for (int l = 0; l < NW; ++l) key[l] = NW == 1 ? sm.fp[sl] : sm.kw[l][sl];
            const u32 b = (u32)(key[0] >> shift) & 255u;
            const u32 lo = bh[b], hi = bh[b + 1];
            u32 rank = lo;
            for (u32 j = lo; j < hi; ++j) {
                const u32 os = xslot[j];
                u64 o[NW];
#pragma unroll
                for (int l = 0; l < NW; ++l) o[l] = NW == 1 ? sm.fp[os] : sm.kw[l][os];
                rank += key_less<NW>(o, key) ? 1u : 0u;
            }
            klist[rank] = (u16)sl;
        }
    }
    __syncthreads();
}

// Sorted bin -> arena, at once (EXTENSION, where the occurrence pass needs the table of this bin: the look-back over the
// bins before it may have to wait for them): every entry gets its place, (k-mer, count, occurrence offset) are written in
// order and the occurrence cursors are left in sm.cnt.
template <int NW, bool EXT, int TH>
__device__ __forceinline__ void emit_now(BinSmem<NW, EXT, TH> &sm, const BinParams &P, u32 lb, u32 tk, u32 to)
{
    using Cfg = BinCfg<NW, EXT, TH>;
    constexpr int EPT = Cfg::EPT;
    const u32 tid = threadIdx.x;
    const u16 *klist = sm.klist();
    u32 off[EPT], cnt[EPT];
    if (EXT) {
        // occurrence lists follow the sorted order: offsets inside the bin, left in the slots' counters as cursors
        u32 carry = 0;
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            if ((u32)(r * TH) < tk) {
                const u32 e = tid + r * TH;
                cnt[r] = e < tk ? sm.cnt[klist[e]] : 0u;
                u32 ex, d0, t0, t1;
                block_scan2<TH>(cnt[r], 0u, sm.wa, sm.wb, ex, d0, t0, t1);
                off[r] = carry + ex;
                carry += t0;
            }
        }
        __syncthreads();   // every count has been read: the counters become cursors
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const u32 e = tid + r * TH;
            if (e < tk) sm.cnt[klist[e]] = off[r];
        }
    } else {
#pragma unroll
        for (int r = 0; r < EPT; ++r) { const u32 e = tid + r * TH; cnt[r] = e < tk ? sm.cnt[klist[e]] : 0u; }
    }
    resolve_position(sm, P, lb, tk, to);
    if (sm.skip_out) return;
    const u64 bk = sm.base_k, bo = sm.base_o;
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        const u32 e = tid + r * TH;
        if (e < tk) {
            const u32 slot = klist[e], c = cnt[r];
#pragma unroll
            for (int l = 0; l < NW; ++l) P.out_words[(bk + e) * NW + l] = NW == 1 ? sm.fp[slot] : sm.kw[l][slot];
            P.out_cnt[bk + e] = c;
            if (EXT) P.out_occ_off[bk + e] = bo + off[r];
            if (c < (u32)BN_HCAP) atomicAdd(&sm.hist[c], 1u); else atomicAdd(&P.histogram[c], 1ull);
        }
    }
}

// Sorted bin -> the CTA's scratch in global memory (L2).  Its place in the arena is resolved while the CTA works on its
// next bin (flush_pending): by then the bins before it have published how many entries they keep, so the look-back finds
// every one of them without waiting — emitting at once makes every CTA wait for the slowest of the bins in flight.
template <int NW, bool EXT, int TH>
__device__ __forceinline__ void stash_bin(BinSmem<NW, EXT, TH> &sm, const BinParams &P, u32 lb, u32 tk)
{
    using Cfg = BinCfg<NW, EXT, TH>;
    const u32 tid = threadIdx.x;
    const u16 *klist = sm.klist();
    u64 *pw = P.pend_words + (size_t)blockIdx.x * BN_SORTCAP * NW;
    u32 *pc = P.pend_cnt + (size_t)blockIdx.x * BN_SORTCAP;
    for (u32 e = tid; e < tk; e += TH) {
        const u32 slot = klist[e], c = sm.cnt[slot];
#pragma unroll
        for (int l = 0; l < NW; ++l) __stcg(&pw[(size_t)e * NW + l], NW == 1 ? sm.fp[slot] : sm.kw[l][slot]);
        __stcg(&pc[e], c);
        if (c < (u32)BN_HCAP) atomicAdd(&sm.hist[c], 1u); else atomicAdd(&P.histogram[c], 1ull);
    }
    if (tid == 0) { sm.pend_valid = 1; sm.pend_lb = lb; sm.pend_tk = tk; }
}

// "bin lb is in the arena": group bookkeeping for the host that streams the result out (one thread)
__device__ __forceinline__ void bin_done(const BinParams &P, u32 lb)
{
    __threadfence();
    const u32 g = lb / P.group_bins;
    const u32 gsize = min(P.group_bins, P.nbins - g * P.group_bins);
    if (atomicAdd(&P.grp_done[g], 1u) + 1 == gsize && P.snap) {
        __threadfence();
        volatile u64 *sn = P.snap + 4 * (size_t)g;
        sn[0] = P.grp_end[2 * g]; sn[1] = P.grp_end[2 * g + 1]; sn[2] = P.grp_big[g];
        __threadfence_system();
        sn[3] = 1;
    }
}

// the stashed bin of this CTA, if any: look-back, copy to the arena, bookkeeping.  Called by all threads between two
// barriers of the bin loop (sm.base_k / base_o / skip_out are not in use there).
template <int NW, bool EXT, int TH>
__device__ __forceinline__ void flush_pending(BinSmem<NW, EXT, TH> &sm, const BinParams &P)
{
    using Cfg = BinCfg<NW, EXT, TH>;
    if (!sm.pend_valid) return;
    const u32 tid = threadIdx.x, lb = sm.pend_lb, tk = sm.pend_tk;
    resolve_position(sm, P, lb, tk, 0u);
    if (!sm.skip_out) {
        const u64 bk = sm.base_k;
        const u64 *pw = P.pend_words + (size_t)blockIdx.x * BN_SORTCAP * NW;
        const u32 *pc = P.pend_cnt + (size_t)blockIdx.x * BN_SORTCAP;
        for (u32 i = tid; i < tk * NW; i += TH) P.out_words[bk * NW + i] = __ldcg(pw + i);
        for (u32 e = tid; e < tk; e += TH) P.out_cnt[bk + e] = __ldcg(pc + e);
    }
    __syncthreads();
    if (tid == 0) { sm.pend_valid = 0; bin_done(P, lb); }
}

// One CTA per bin, bins taken in index order through a ticket; a bin of any size is handled as long as its distinct
// k-mers fit the table (anything else is listed for the HBM path).
template <int NW, bool EXT, int TH>
__global__ void __launch_bounds__(TH, BinCfg<NW, EXT, TH>::CTAS) k_bin_count(BinParams P)
{
    using Cfg = BinCfg<NW, EXT, TH>;
    constexpr int SW = Cfg::SW;
    constexpr bool DEDUP = (NW <= 2) && !EXT;
    static_assert(!DEDUP || 2 * Cfg::DDTS <= Cfg::TS, "supermer table fits the k-mer table");
    extern __shared__ __align__(16) unsigned char smraw[];
    BinSmem<NW, EXT, TH> &sm = *reinterpret_cast<BinSmem<NW, EXT, TH> *>(smraw);
    const int tid = threadIdx.x;
    const int k = P.k;
    const int padbits = 2 * (32 * NW - k);

    for (int i = tid; i < BN_HCAP; i += TH) sm.hist[i] = 0;
    if (tid == 0) sm.pend_valid = 0;

    while (true) {
        __syncthreads();   // end of the previous bin: shared memory is free again
        if (tid == 0) {
            sm.bin = atomicAdd(P.ticket, 1u);
            sm.bail = 0; sm.next_batch = 0; sm.seen = 0; sm.ncand = 0; sm.skip_out = 0; sm.xor_hi = 0; sm.xor_lo = 0;
        }
        __syncthreads();
        const u32 lb = sm.bin;
        if (lb >= P.nbins) break;
        bool deferred = false;   // this bin's "in the arena" bookkeeping happens when its stash is flushed

        // ---- bin descriptor: one segment of slots per source rank
        if (tid < P.nsrc) {
            const u64 i0 = P.seg_start[tid][lb], i1 = P.seg_start[tid][lb + 1];
            sm.src_ptr[tid] = P.slots[tid] + i0 * (u64)SW;
            sm.src_n[tid] = (u32)min(i1 - i0, (u64)0xFFFFFFFFu);
        }
        __syncthreads();
        if (tid == 0) {
            u64 s = 0;
            for (int i = 0; i < P.nsrc; ++i) { sm.src_sbase[i] = (u32)min(s, (u64)0xFFFFFFFFu); s += sm.src_n[i]; }
            sm.src_sbase[P.nsrc] = (u32)min(s, (u64)0xFFFFFFFFu);
            const u64 nk = bt_kmers(P.bin_kmers[lb]);
            sm.nk = (u32)min(nk, (u64)0xFFFFFFFFu);
            sm.S = (u32)min(s, (u64)0xFFFFFFFFu);
            if (nk >= 0xFFFFFFFFull || s >= 0xFFFFFFFFull) sm.bail = 1;   // 32-bit counters
            sm.nsrc = P.nsrc; sm.mult = nullptr;
        }
        __syncthreads();
        const u32 nk = sm.nk;
        u32 S = sm.S;
        const bool dd = DEDUP && !sm.bail && S > 0 && S <= (u32)Cfg::DDLIMIT && P.dd_slots != nullptr;
        uint4 ddv[Cfg::DDPT][SW / 4];
        if constexpr (DEDUP) { if (dd) dedup_fetch<NW, EXT, TH>(sm, S, ddv); }

        // ---- empty table (the slots of a de-duplicated bin are on their way meanwhile)
        {
            const uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u), zero = make_uint4(0, 0, 0, 0);
            if (NW == 1) { for (int i = tid; i < Cfg::TS / 2; i += TH) reinterpret_cast<uint4 *>(sm.fp)[i] = ones; }
            else { for (int i = tid; i < Cfg::TS / 4; i += TH) reinterpret_cast<uint4 *>(sm.fp32)[i] = ones; }
            // de-duplication first uses the counters as its cells (empty = all ones) and weights
            for (int i = tid; i < Cfg::TS / 4; i += TH)
                reinterpret_cast<uint4 *>(sm.cnt)[i] = (dd && i < Cfg::DDTS / 4) ? ones : zero;
        }
        __syncthreads();
        if constexpr (DEDUP) {
            if (dd) {
                const u32 Sd = dedup_bin<NW, EXT, TH>(sm, P, S, k, ddv);
                __syncthreads();   // the list is complete; the table memory goes back to the k-mers
                {
                    const uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u), zero = make_uint4(0, 0, 0, 0);
                    // the supermer keys overlaid the first DDTS * SW * 4 bytes of the k-mer cells (K <= 32 only: for K > 32
                    // they lie in the key words, which need no clearing)
                    if (NW == 1) { for (int i = tid; i < Cfg::DDTS * (SW / 4); i += TH) reinterpret_cast<uint4 *>(sm.fp)[i] = ones; }
                    for (int i = tid; i < 2 * Cfg::DDTS / 4; i += TH) reinterpret_cast<uint4 *>(sm.cnt)[i] = zero;
                }
                if (tid == 0) {
                    sm.nsrc = 1;
                    sm.src_ptr[0] = reinterpret_cast<const u32 *>(P.dd_slots + (size_t)blockIdx.x * BN_DDLIMIT_MAX * (SW / 4));
                    sm.mult = P.dd_mult + (size_t)blockIdx.x * BN_DDLIMIT_MAX;
                }
                S = Sd;
                __syncthreads();
            }
        }

        // ---- expand + insert + count.  About two batches per warp, so that the warps finish together
        if (tid == 0) {
            const u32 parts = P.walk_split * Cfg::WARPS;
            u32 bs = (S + parts - 1) / parts;
            sm.batch_slots = min((u32)Cfg::BATCH, max(P.walk_min, bs));
        }
        __syncthreads();
        if (!sm.bail) walk_bin<false>(sm, P, k, padbits, S);
        __syncthreads();
        if (!sm.bail && sm.seen != nk && tid == 0) sm.bail = 2;   // inconsistent totals: never count from a corrupt table
        __syncthreads();
        bool bailed = sm.bail != 0;   // the bin goes to the HBM path; it still takes its (empty) place in the chain

        // ---- filter.  Usually the candidate list (slots that reached LOWER) names the few slots to look at; a bin
        //      with more candidates than the list holds (LOWER == 1, mostly) scans its whole table.
        const u32 ncand = sm.ncand;
        const bool listed = ncand <= (u32)BN_CAND;
        constexpr int LPT = BN_CAND / TH;   // listed candidates per thread, at most
        const u32 per_thread = bailed ? 0u : (listed ? (ncand + TH - 1) / TH : (u32)Cfg::SLOTS_PT);
        u32 kept = 0, occ = 0, keepmask = 0;
        u32 myslot[LPT];
        if (listed) {
#pragma unroll
            for (int i = 0; i < LPT; ++i) {
                const u32 idx = tid * per_thread + i;
                const bool in = (u32)i < per_thread && idx < ncand;
                const u32 slot = in ? sm.cand[idx] : 0u;
                myslot[i] = slot;
                const u32 c = sm.cnt[slot];
                if (in && c >= P.lower && c <= P.upper) { keepmask |= 1u << i; ++kept; occ += c; }
            }
        } else if (!bailed) {
#pragma unroll
            for (int i = 0; i < Cfg::SLOTS_PT; ++i) {
                const u32 c = sm.cnt[tid * Cfg::SLOTS_PT + i];
                if (c >= P.lower && c <= P.upper) { keepmask |= 1u << i; ++kept; occ += c; }
            }
        }
        u32 ek, eo, tk, to;
        block_scan2<TH>(kept, occ, sm.wa, sm.wb, ek, eo, tk, to);   // (its barriers: every candidate has been read)
        if (!EXT) to = 0;
        const bool big = tk > (u32)Cfg::SORTCAP;   // too many kept k-mers to sort here (never a listed bin): staging area + big gather
        u64 stage_k = 0, stage_o = 0;
        if (big) {
            // room in the staging area?  Otherwise the bin goes through the HBM path like an overflowing one.
            if (tid == 0) {
                const u64 sk = atomicAdd(P.stage_cursor, (u64)tk);
                const u64 so = (EXT && to) ? atomicAdd(P.stage_cursor + 1, (u64)to) : 0;
                if (sk + tk > P.stage_cap || (EXT && so + to > P.stage_occ_cap)) sm.bail = 32;
                sm.base_k = sk; sm.base_o = so;
            }
            __syncthreads();
            if (sm.bail) { bailed = true; tk = 0; to = 0; }
            stage_k = sm.base_k; stage_o = sm.base_o;
        }
        if (tid == 0) {
            volatile u64 *lbs = P.lb_state;
            lbs[lb] = LB_AGG | tk;
            if (EXT) lbs[P.nbins + lb] = LB_AGG | to;
            sm.next_batch = 0;
            if (bailed) P.ovf_list[atomicAdd(P.ovf_count, 1u)] = lb;
        }
        // the bin this CTA finished before this one goes to the arena now: every bin before it has long published its total
        if (!EXT) { __syncthreads(); flush_pending<NW, EXT, TH>(sm, P); }
        if (EXT && !bailed && !big) {
            // slots that are not kept: the occurrence pass tells by the mark (kept slots get their cursor below)
            __syncthreads();
            for (int i = tid; i < Cfg::TS / 4; i += TH) {
                uint4 v = reinterpret_cast<uint4 *>(sm.cnt)[i];
                if (v.x < P.lower || v.x > P.upper) v.x = BN_NOTKEPT;
                if (v.y < P.lower || v.y > P.upper) v.y = BN_NOTKEPT;
                if (v.z < P.lower || v.z > P.upper) v.z = BN_NOTKEPT;
                if (v.w < P.lower || v.w > P.upper) v.w = BN_NOTKEPT;
                reinterpret_cast<uint4 *>(sm.cnt)[i] = v;
            }
        }
        __syncthreads();
        if (bailed) {
            resolve_position(sm, P, lb, 0u, 0u);
        } else if (!big) {
            // compacted list of the kept slots (the walk's scratch is free), then the sort; the sorted bin goes to the
            // CTA's scratch (its place in the arena is resolved later) or, with EXTENSION, straight into the arena
            u32 g = ek;
            if (listed) {
#pragma unroll
                for (int i = 0; i < LPT; ++i) if ((keepmask >> i) & 1) sm.klist()[g++] = (u16)myslot[i];
            } else {
#pragma unroll
                for (int i = 0; i < Cfg::SLOTS_PT; ++i) if ((keepmask >> i) & 1) sm.klist()[g++] = (u16)(tid * Cfg::SLOTS_PT + i);
            }
            __syncthreads();
            if (tk) sort_bin<NW, EXT, TH>(sm, tk);
            if (EXT) {
                emit_now<NW, EXT, TH>(sm, P, lb, tk, to);
                if (tid == 0) { sm.occ_pos = P.out_pos; sm.occ_rid = P.out_rid; }
            } else {
                stash_bin<NW, EXT, TH>(sm, P, lb, tk);
                deferred = true;
            }
        } else {
            // unsorted into the staging area (the whole table was scanned); the big gather sorts and moves the bin to its
            // place in the arena
            const u64 sk = stage_k, so = stage_o;
            u64 g = sk + ek;
            u32 lo = eo;   // occurrence offset inside the bin
#pragma unroll
            for (int i = 0; i < Cfg::SLOTS_PT; ++i) {
                const u32 slot = tid * Cfg::SLOTS_PT + i;
                u32 mark = BN_NOTKEPT;
                if ((keepmask >> i) & 1) {
                    const u32 c = sm.cnt[slot];
                    if (NW == 1) P.st_words[g] = sm.fp[slot];
                    else {
#pragma unroll
                        for (int l = 0; l < NW; ++l) P.st_words[g * NW + l] = sm.kw[l][slot];
                    }
                    P.st_cnt[g] = c;
                    if (c < (u32)BN_HCAP) atomicAdd(&sm.hist[c], 1u); else atomicAdd(&P.histogram[c], 1ull);
                    mark = lo;
                    lo += c;
                    ++g;
                }
                if (EXT) sm.cnt[slot] = mark;
            }
            __syncthreads();
            resolve_position(sm, P, lb, tk, to);   // sm.base_k / base_o: now the place in the arena
            if (tid == 0) {
                P.bin_rec[4 * (size_t)lb + 0] = sk; P.bin_rec[4 * (size_t)lb + 1] = sm.skip_out ? 0 : tk;
                P.bin_rec[4 * (size_t)lb + 2] = so; P.bin_rec[4 * (size_t)lb + 3] = to;
                P.fin[2 * (size_t)lb] = sm.base_k; P.fin[2 * (size_t)lb + 1] = sm.base_o;
                P.big_list[atomicAdd(P.big_count, 1u)] = lb;
                atomicAdd(&P.grp_big[lb / P.group_bins], 1u);
                sm.base_o = so;   // pass 2 writes the occurrences to the staging area
                sm.occ_pos = P.st_pos; sm.occ_rid = P.st_rid;
            }
        }
        if (EXT) {
            // ---- occurrences: (pos, rid) of every occurrence of a kept k-mer, grouped per k-mer
            __syncthreads();
            if (to && !bailed && !sm.skip_out) walk_bin<true>(sm, P, k, padbits, S);
        }

        // ---- the bin is in the arena (unless it waits in the stash): group bookkeeping for the host that streams the result out
        __syncthreads();
        if (tid == 0 && !deferred) bin_done(P, lb);
    }

    __syncthreads();
    if (!EXT) { flush_pending<NW, EXT, TH>(sm, P); __syncthreads(); }
    for (int i = tid; i < BN_HCAP; i += TH)
        if (sm.hist[i]) atomicAdd(&P.histogram[i], (u64)sm.hist[i]);
}

// ---- per bin: sort the kept k-mers by key and move them (and their occurrences) to the arena ----------
// One CTA per bin.  CAPD = most kept entries this instantiation handles; bins with more than CAPD or at
// most skip_upto entries are left to the other launch.
template <int NW, bool EXT, int CAPD, int THREADS, bool FROM_LIST>
__global__ void __launch_bounds__(THREADS) k_bin_gather(BinParams P)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    u64 *s_key = reinterpret_cast<u64 *>(smraw);                       // [NW][CAPD]
    u32 *s_cnt = reinterpret_cast<u32 *>(s_key + (size_t)NW * CAPD);  // [CAPD]
    u32 *s_src = s_cnt + CAPD;                                        // [CAPD] occurrence start inside the bin (staging order)
    __shared__ u32 s_warp[THREADS / 32];
  for (u32 work = blockIdx.x; work < (FROM_LIST ? *P.big_count : P.nbins); work += gridDim.x) {
    const u32 lb = FROM_LIST ? P.big_list[work] : work;
    const u64 sk = P.bin_rec[4 * (size_t)lb + 0];
    const u32 D = (u32)P.bin_rec[4 * (size_t)lb + 1];
    const u64 so = P.bin_rec[4 * (size_t)lb + 2];
    if (D == 0 || D > (u32)CAPD) continue;
    __syncthreads();   // shared memory of the previous bin is free
    const u64 fk = P.fin[2 * (size_t)lb], fo = P.fin[2 * (size_t)lb + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u32 n2 = 1;
    while (n2 < D) n2 <<= 1;

    // load; entries beyond D are padding that sorts last
    for (u32 i = tid; i < n2; i += THREADS) {
        if (i < D) {
#pragma unroll
            for (int l = 0; l < NW; ++l) s_key[(size_t)l * CAPD + i] = P.st_words[(sk + i) * NW + l];
            s_cnt[i] = P.st_cnt[sk + i];
        } else {
#pragma unroll
            for (int l = 0; l < NW; ++l) s_key[(size_t)l * CAPD + i] = ~0ull;
            s_cnt[i] = 0;
        }
    }
    __syncthreads();
    if (EXT) {
        // occurrence start of every entry in staging order: exclusive scan of the counts
        u32 carry = 0;
        for (u32 base = 0; base < n2; base += THREADS) {
            const u32 i = base + tid;
            const u32 c = i < D ? s_cnt[i] : 0;
            u32 inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d) inc += t;
            }
            if (lane == 31) s_warp[warp] = inc;
            __syncthreads();
            u32 off = carry, tot = 0;
            for (int w = 0; w < THREADS / 32; ++w) { if (w < warp) off += s_warp[w]; tot += s_warp[w]; }
            if (i < n2) s_src[i] = off + inc - c;
            carry += tot;
            __syncthreads();
        }
    }

    // bitonic sort by key (ascending, word 0 most significant), payload (cnt, src) follows
    for (u32 size = 2; size <= n2; size <<= 1) {
        for (u32 stride = size >> 1; stride > 0; stride >>= 1) {
            for (u32 t = tid; t < n2 / 2; t += THREADS) {
                const u32 i = 2 * t - (t & (stride - 1));
                const u32 j = i + stride;
                const bool up = ((i & size) == 0);
                u64 a[NW], b[NW];
#pragma unroll
                for (int l = 0; l < NW; ++l) { a[l] = s_key[(size_t)l * CAPD + i]; b[l] = s_key[(size_t)l * CAPD + j]; }
                const bool swap = up ? key_less<NW>(b, a) : key_less<NW>(a, b);
                if (swap) {
#pragma unroll
                    for (int l = 0; l < NW; ++l) { s_key[(size_t)l * CAPD + i] = b[l]; s_key[(size_t)l * CAPD + j] = a[l]; }
                    const u32 c = s_cnt[i]; s_cnt[i] = s_cnt[j]; s_cnt[j] = c;
                    if (EXT) { const u32 s = s_src[i]; s_src[i] = s_src[j]; s_src[j] = s; }
                }
            }
            __syncthreads();
        }
    }

    // write in sorted order; occurrence offsets = exclusive scan of the counts in sorted order
    u32 carry = 0;
    for (u32 base = 0; base < n2; base += THREADS) {
        const u32 i = base + tid;
        const u32 c = i < D ? s_cnt[i] : 0;
        u32 dst = 0;
        if (EXT) {
            u32 inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d) inc += t;
            }
            if (lane == 31) s_warp[warp] = inc;
            __syncthreads();
            u32 off = carry, tot = 0;
            for (int w = 0; w < THREADS / 32; ++w) { if (w < warp) off += s_warp[w]; tot += s_warp[w]; }
            dst = off + inc - c;
            carry += tot;
            __syncthreads();
        }
        if (i < D) {
#pragma unroll
            for (int l = 0; l < NW; ++l) P.out_words[(fk + i) * NW + l] = s_key[(size_t)l * CAPD + i];
            P.out_cnt[fk + i] = c;
            if (EXT) {
                P.out_occ_off[fk + i] = fo + dst;
                const u64 src = so + s_src[i], dd = fo + dst;
                for (u32 t = 0; t < c; ++t) { P.out_pos[dd + t] = P.st_pos[src + t]; P.out_rid[dd + t] = P.st_rid[src + t]; }
            }
        }
    }
  }
}

// ---- multi-rank bookkeeping from the all-gathered bin totals -----------------------------------------
// alltot[src][b] = bt_pack(slots, k-mers) of bin b as extracted by rank src.  Rank r owns bins [r*tg, (r+1)*tg) and
// receives them in one buffer: one region per source rank (in rank order), bins in index order inside a region.

// block-wide exclusive scan helper for one u64 per thread (1024 threads), with a running carry
__device__ __forceinline__ u64 scan1024(u64 v, u64 *s_c, u64 &carry)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u64 x = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += x;
    }
    if (lane == 31) s_c[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const u64 x = s_c[lane];
        u64 ix = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u64 y = __shfl_up_sync(0xFFFFFFFFu, ix, d);
            if (lane >= d) ix += y;
        }
        s_c[lane] = ix - x;
    }
    __syncthreads();
    const u64 ex = carry + s_c[warp] + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = ex + v;
    __syncthreads();
    return ex;
}

// block src: where the owned bins start inside the bin-major supermer stream of rank src (absolute slot index in
// src's buffer when `absolute`, else relative to the first owned bin); meta[src] = slots of the owned bins in it
__global__ void __launch_bounds__(1024) k_seg_scan(const u64 *__restrict__ alltot, u32 T, u32 b_lo, u32 tg, int absolute,
                                                    u64 *__restrict__ seg_start, u64 *__restrict__ meta)
{
    __shared__ u64 s_c[32];
    __shared__ u64 carry, first;
    const int src = blockIdx.x;
    const u64 *row0 = alltot + (size_t)src * T;
    u64 before = 0;
    if (absolute) {
        for (u32 b = threadIdx.x; b < b_lo; b += 1024) before += bt_slots(row0[b]);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) before += __shfl_xor_sync(0xFFFFFFFFu, before, d);
    }
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = before;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 t = 0;
        for (int w = 0; w < 32; ++w) t += s_c[w];
        carry = t; first = t;
    }
    __syncthreads();
    const u64 *row = row0 + b_lo;
    u64 *os = seg_start + (size_t)src * (tg + 1);
    for (u32 base = 0; base < tg; base += 1024) {
        const u32 b = base + threadIdx.x;
        const u64 ex = scan1024(b < tg ? bt_slots(row[b]) : 0, s_c, carry);
        if (b < tg) os[b] = ex;
    }
    if (threadIdx.x == 0) { os[tg] = carry; meta[src] = carry - first; }
}

// k-mers per owned bin summed over the source ranks, and their grand total (atomicAdd into *owned_total)
__global__ void __launch_bounds__(256) k_sum_kmers(const u64 *__restrict__ alltot, u32 T, u32 b_lo, u32 tg, int nranks,
                                                    u64 *__restrict__ bin_kmers, u64 *__restrict__ owned_total)
{
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    u64 s = 0;
    if (b < tg) {
        for (int src = 0; src < nranks; ++src) s += bt_kmers(alltot[(size_t)src * T + b_lo + b]);
        bin_kmers[b] = s;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(owned_total, s);
}

cudaError_t launch_seg_scan(const u64 *alltot, u32 T, int me, u32 tg, int nranks, bool absolute, u64 *seg_start, u64 *meta,
                            u64 *bin_kmers, u64 *owned_total, cudaStream_t s)
{
    const u32 b_lo = (u32)me * tg;
    k_seg_scan<<<nranks, 1024, 0, s>>>(alltot, T, b_lo, tg, absolute ? 1 : 0, seg_start, meta);
    k_sum_kmers<<<(tg + 255) / 256, 256, 0, s>>>(alltot, T, b_lo, tg, nranks, bin_kmers, owned_total);
    return cudaGetLastError();
}

// 228 KB of shared memory per SM, 1 KB reserved per CTA
static_assert(sizeof(BinSmem<1, false, 512>) <= (233472 - 2 * 1024) / 2, "K <= 32: two CTAs per SM");
static_assert(sizeof(BinSmem<1, true, 512>) <= (233472 - 3 * 1024) / 3, "K <= 32 with EXTENSION: three CTAs per SM");
static_assert(sizeof(BinSmem<2, false, 512>) <= 232448 && sizeof(BinSmem<2, true, 512>) <= 232448 && sizeof(BinSmem<3, true, 512>) <= 232448, "K > 32: one CTA per SM");
static_assert(sizeof(BinSmem<2, false, 1024>) <= 232448, "K in 33..64: one CTA of 1024 threads per SM");
constexpr int GL_THREADS = 512;                    // gather: listed bins with more kept k-mers than a CTA sorts itself

template <int NW, bool EXT, int TH>
static cudaError_t launch_bins_t(const BinParams &P, int sm_count, cudaStream_t s)
{
    const size_t smem = sizeof(BinSmem<NW, EXT, TH>);
    cudaError_t e = cudaFuncSetAttribute(k_bin_count<NW, EXT, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bin_count<NW, EXT, TH>, TH, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    // every CTA must be resident: the look-back chain waits on bins that other CTAs hold
    const u32 grid = (u32)std::min<u64>((u64)sm_count * per_sm, std::max<u32>(P.nbins, 1u));
    k_bin_count<NW, EXT, TH><<<grid, TH, smem, s>>>(P);
    return cudaGetLastError();
}

template <int NW, bool EXT>
static cudaError_t launch_big_t(const BinParams &P, int sm_count, cudaStream_t s)
{
    constexpr int GL_CAP = BinCfg<NW, EXT, 512>::TS;   // a bin keeps at most one entry per table slot
    const size_t smem_l = ((size_t)8 * NW + 8) * GL_CAP;
    cudaError_t e = cudaFuncSetAttribute(k_bin_gather<NW, EXT, GL_CAP, GL_THREADS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l);
    if (e != cudaSuccess) return e;
    k_bin_gather<NW, EXT, GL_CAP, GL_THREADS, true><<<sm_count, GL_THREADS, smem_l, s>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_bin_count(const BinParams &P, int nwords, bool ext, int sm_count, cudaStream_t s)
{
    if (P.nbins == 0) return cudaSuccess;
    if (nwords == 1) {
        if (ext) return launch_bins_t<1, true, 512>(P, sm_count, s);
        return launch_bins_t<1, false, 512>(P, sm_count, s);
    }
    if (nwords == 2) {
        if (ext) return launch_bins_t<2, true, 512>(P, sm_count, s);
        return bin_threads_env(1024, 512) == 1024 ? launch_bins_t<2, false, 1024>(P, sm_count, s) : launch_bins_t<2, false, 512>(P, sm_count, s);
    }
    return ext ? launch_bins_t<3, true, 512>(P, sm_count, s) : launch_bins_t<3, false, 512>(P, sm_count, s);
}

// sorts + moves the bins listed in P.big_list (more kept k-mers than a CTA of k_bin_count sorts itself); launched only
// when the list is not empty
cudaError_t launch_bin_gather_big(const BinParams &P, int nwords, bool ext, int sm_count, cudaStream_t s)
{
    if (nwords == 1) return ext ? launch_big_t<1, true>(P, sm_count, s) : launch_big_t<1, false>(P, sm_count, s);
    if (nwords == 2) return ext ? launch_big_t<2, true>(P, sm_count, s) : launch_big_t<2, false>(P, sm_count, s);
    return ext ? launch_big_t<3, true>(P, sm_count, s) : launch_big_t<3, false>(P, sm_count, s);
}

} // namespace hsk
