// Stages 4+5 fused, on chip: one CTA takes one bin of supermers, expands it to canonical k-mers and
// counts them in shared memory — the k-mers themselves never touch HBM.
//
// Replaces, for bins that fit in shared memory, the reference's HOT LOOPS C+D+E:
//   receive_from_buffer_stage2 + GetRepKmers  (src/kmerops.cpp:484-521, include/kmer.hpp:313-340)
//   sort_task -> RADULS / PARADIS             (src/kmerops.cpp:1382-1407)
//   count_sorted_kmers                        (src/kmerops.cpp:1410-1445), histogram (hysortk.cpp:106-113)
//
// The reference sorts every k-mer occurrence and then run-length counts.  At 30x coverage a bin of
// ~3000 occurrences holds only ~100 genomic k-mers x 30 copies plus error singletons, and all of
// them are overlapping windows of a few loci (they share minimizers), so their leading bases take
// few values: sorting the occurrences is both unbalanced (30-copy lumps) and wasted work.  Here:
//
//   k_bin_count   (one CTA per bin, bins taken in index order through a ticket)
//     1. supermer table of the bin: block scan of k-mer / word offsets
//     2. every thread expands its share of consecutive k-mers (rolling forward / reverse words inside a
//        supermer) and inserts each into an open-addressing table in shared memory keyed by the
//        canonical k-mer (64-bit CAS; for K > 32 a 64-bit fingerprint of the words is the CAS key and
//        the full words are verified afterwards — a fingerprint clash sends the bin to the HBM path,
//        so the result stays exact).  The slot's 16-bit counter gives the count AND the index of this
//        occurrence among its k-mer's occurrences (used to place (pos, rid) when EXTENSION).
//     3. every thread filters its slots with LOWER <= count <= UPPER, a block scan compacts the kept
//        (k-mer, count) pairs and they are written to a staging area at an atomically claimed offset
//   k_bin_offsets  exclusive scan of the per-bin kept / occurrence totals -> final positions
//   k_bin_gather   (one CTA per bin) sorts the bin's kept k-mers by key (bitonic sort in shared
//                  memory over the few distinct kept keys) and writes them, with their occurrence
//                  lists, to the final arena.
//
// So the SORT is still there, but it runs over the distinct kept k-mers (D) instead of over every
// occurrence (N): D/N is ~4 % at 30x coverage with 1 % errors.  The arena holds the bins in index order
// and ascending k-mers inside each bin (the reference: per-task sorted runs in task order,
// kmerops.cpp:883-904), deterministically.  Bins are sized by the extraction stage so that they fit
// (CAP k-mers); the rare bin that does not (skew) is reported in an overflow list and goes through the
// HBM path (expand.cu -> radix.cu -> count.cu).
#include "kernels.cuh"

#include <algorithm>

namespace hsk {

constexpr int BN_SPT = BN_SCAP / BN_THREADS;   // supermers per thread in the table scan
constexpr int BN_HCAP = 2048;                  // shared-memory histogram bins
constexpr u64 BN_EMPTY = ~0ull;                // never a canonical k-mer: a K-mer of all T is not canonical

template <int NW>
struct BinCfg {
    static constexpr int KPT = NW == 1 ? 12 : (NW == 2 ? 6 : 4);
    static constexpr int CAP = BN_THREADS * KPT;           // 6144 / 3072 / 2048 k-mers per bin
    static constexpr int TS_BITS = NW == 1 ? 13 : 12;      // table slots: 8192 / 4096 / 4096
    static constexpr int TS = 1 << TS_BITS;
    static constexpr int SLOTS_PT = TS / BN_THREADS;       // slots per thread in the filter: 16 / 8 / 8
};

int bin_capacity(int nwords, bool ext)
{
    (void)ext;
    return BN_THREADS * (nwords == 1 ? 12 : (nwords == 2 ? 6 : 4));
}

template <int NW>
struct BinSmem {
    u64 fp[BinCfg<NW>::TS];                                 // CAS key: the k-mer (NW == 1) or its fingerprint
    u64 kw[NW > 1 ? NW : 1][NW > 1 ? BinCfg<NW>::TS : 1];   // full key words (NW > 1)
    u32 cnt[BinCfg<NW>::TS];                                // occurrences per slot; later: occurrence offset of the slot
    u16 koff[BN_SCAP + 2];                                  // k-mer offset of every slot of the chunk
    u32 hist[BN_HCAP];
    const u32 *src_ptr[BN_MAX_SRC];                         // first slot of the bin in every source stream
    u32 src_n[BN_MAX_SRC], src_sbase[BN_MAX_SRC + 1];
    u32 wa[BN_THREADS / 32], wb[BN_THREADS / 32];
    u64 stage_kept, stage_occ;
    u32 bin, nk, S, bail;
};

// block-wide exclusive scan of two u32 values (BN_THREADS threads); returns exclusive prefixes and totals
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
    u32 oa = 0, ob = 0;
    ta = 0; tb = 0;
#pragma unroll
    for (int i = 0; i < BN_THREADS / 32; ++i) {
        u32 x = wa[i], y = wb[i];
        if (i < warp) { oa += x; ob += y; }
        ta += x; tb += y;
    }
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

// CAS key of a k-mer: the k-mer itself when it is one word, else a 64-bit fingerprint of its words
template <int NW>
__device__ __forceinline__ u64 fingerprint(const u64 (&w)[NW])
{
    if (NW == 1) return w[0];
    u64 h = w[0] * 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int l = 1; l < NW; ++l) {
        h ^= h >> 29;
        h = (h + w[l]) * 0xBF58476D1CE4E5B9ull;
    }
    h ^= h >> 32;
    return h == BN_EMPTY ? 0x5851F42D4C957F2Dull : h;
}


// expansion state of one thread: the slot it is in (SW words in registers, shifted so that the current
// k-mer starts at the top), the k-mers left in the slot, and the position of the current k-mer
template <int SW>
struct Walk {
    u32 w[SW];
    u32 j;        // slot index inside the bin
    u32 left;     // k-mers of this slot not yet produced (0: load the next slot)
    u32 pos, rid; // EXTENSION: PosInRead of the current k-mer, ReadId
};

template <int SW, int PW>
__device__ __forceinline__ void shift_bases(u32 (&w)[SW], u32 nb)
{
    // left shift of the payload words by nb bases (nb < 16)
    const u32 sh = 2 * nb;
#pragma unroll
    for (int x = 0; x < PW - 1; ++x) w[x] = __funnelshift_l(w[x + 1], w[x], sh);
    w[PW - 1] <<= sh;
}

// address of slot j of the bin (j counts over the per-source segments in rank order)
template <int NW>
__device__ __forceinline__ const u32 *slot_ptr(const BinSmem<NW> &sm, const BinParams &P, u32 j, int sw)
{
    if (P.nsrc == 1) return sm.src_ptr[0] + j * (u32)sw;
    int s = 0;
    while (j >= sm.src_sbase[s + 1]) ++s;
    return sm.src_ptr[s] + (j - sm.src_sbase[s]) * (u32)sw;
}

// canonical k-mer at the current position, then advance by one k-mer
template <int NW, int SW, bool EXT>
__device__ __forceinline__ void next_kmer(Walk<SW> &r, const BinSmem<NW> &sm, const BinParams &P, int k, int padbits,
                                          u32 skip, u64 (&key)[NW], u64 &val)
{
    constexpr int PW = SW - (EXT ? 2 : 0);
    if (r.left == 0) {
        const uint4 *sp = reinterpret_cast<const uint4 *>(slot_ptr<NW>(sm, P, r.j, SW));
#pragma unroll
        for (int x = 0; x < SW / 4; ++x) {
            const uint4 v = __ldg(sp + x);
            r.w[4 * x] = v.x; r.w[4 * x + 1] = v.y; r.w[4 * x + 2] = v.z; r.w[4 * x + 3] = v.w;
        }
        r.left = (r.w[PW - 1] & 0xFFu) - (u32)k + 1 - skip;
        if (EXT) { r.pos = r.w[SW - 2] + skip; r.rid = r.w[SW - 1]; }
        // start inside the slot: drop `skip` bases (only the first slot of a thread's range)
        for (u32 t = skip; t > 0;) {
            const u32 step = min(t, 15u);
            shift_bases<SW, PW>(r.w, step);
            t -= step;
        }
    }
    u64 fwd[NW], rc[NW];
#pragma unroll
    for (int l = 0; l < NW; ++l) fwd[l] = ((u64)r.w[2 * l] << 32) | r.w[2 * l + 1];
    if (padbits) fwd[NW - 1] &= ~0ull << padbits;
    kmer_twin<NW>(fwd, k, rc);
    const bool use_rc = key_less<NW>(rc, fwd);
#pragma unroll
    for (int l = 0; l < NW; ++l) key[l] = use_rc ? rc[l] : fwd[l];
    if (EXT) { val = ((u64)r.pos << 32) | r.rid; ++r.pos; }
    shift_bases<SW, PW>(r.w, 1);
    if (--r.left == 0) ++r.j;
}

// claim or find the slot of a k-mer; a full table (more distinct k-mers than slots) sets the bail flag
template <int NW>
__device__ __forceinline__ u32 table_probe(BinSmem<NW> &sm, const u64 (&key)[NW])
{
    using Cfg = BinCfg<NW>;
    const u64 f = fingerprint<NW>(key);
    u32 slot = (u32)((f * 0x9E3779B97F4A7C15ull) >> (64 - Cfg::TS_BITS));
    for (int probes = 0;; ++probes) {
        const u64 old = atomicCAS(&sm.fp[slot], BN_EMPTY, f);
        if (old == BN_EMPTY) {
            if (NW > 1) {
#pragma unroll
                for (int l = 0; l < NW; ++l) sm.kw[l][slot] = key[l];
            }
            break;
        }
        if (old == f) break;
        if (probes >= Cfg::TS) { atomicOr(&sm.bail, 16u); break; }
        slot = (slot + 1) & (Cfg::TS - 1);
    }
    return slot;
}

// claim or find the slot of a k-mer and bump its counter; returns the slot and the counter word before the bump.
template <int NW>
__device__ __forceinline__ u32 table_insert(BinSmem<NW> &sm, const u64 (&key)[NW], u32 &prev)
{
    using Cfg = BinCfg<NW>;
    const u64 f = fingerprint<NW>(key);
    u32 slot = (u32)((f * 0x9E3779B97F4A7C15ull) >> (64 - Cfg::TS_BITS));
    for (int probes = 0;; ++probes) {
        const u64 old = atomicCAS(&sm.fp[slot], BN_EMPTY, f);
        if (old == BN_EMPTY) {
            if (NW > 1) {
#pragma unroll
                for (int l = 0; l < NW; ++l) sm.kw[l][slot] = key[l];
            }
            break;
        }
        if (old == f) break;
        if (probes >= Cfg::TS) { atomicOr(&sm.bail, 16u); prev = 0; return slot; }
        slot = (slot + 1) & (Cfg::TS - 1);
    }
    prev = atomicAdd(&sm.cnt[slot], 1u);
    return slot;
}

// One CTA per bin.  K <= 32 without EXTENSION needs no per-occurrence state after the insertion, so a bin of
// any size up to 65535 occurrences (16-bit counters / offsets) is handled, slots in chunks of BN_SCAP, as
// long as its distinct k-mers fit the table.  With EXTENSION or K > 32 every thread keeps its occurrences in
// registers for the second pass, which limits a bin to BinCfg::CAP occurrences and BN_SCAP slots.
template <int NW, bool EXT>
__global__ void __launch_bounds__(BN_THREADS, NW == 1 ? 2 : 1) k_bin_count(BinParams P)
{
    using Cfg = BinCfg<NW>;
    constexpr bool FREE = (NW == 1) && !EXT;
    constexpr int SW = (NW == 1 ? 4 : 8) + (EXT ? 4 : 0);
    constexpr int PW = SW - (EXT ? 2 : 0);
    extern __shared__ __align__(16) unsigned char smraw[];
    BinSmem<NW> &sm = *reinterpret_cast<BinSmem<NW> *>(smraw);
    const int tid = threadIdx.x;
    const int k = P.k;
    const int padbits = 2 * (32 * NW - k);

    for (int i = tid; i < BN_HCAP; i += BN_THREADS) sm.hist[i] = 0;

    while (true) {
        __syncthreads();   // end of the previous bin: shared memory is free again
        if (tid == 0) { sm.bin = atomicAdd(P.ticket, 1u); sm.bail = 0; }
        for (int i = tid; i < Cfg::TS; i += BN_THREADS) sm.fp[i] = BN_EMPTY;
        for (int i = tid; i < Cfg::TS; i += BN_THREADS) sm.cnt[i] = 0;
        __syncthreads();
        const u32 lb = sm.bin;
        if (lb >= P.nbins) break;

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
            const u64 nk = P.bin_kmers[lb] & ((1ull << 40) - 1);
            sm.nk = (u32)min(nk, (u64)0xFFFFFFFFu);
            sm.S = (u32)min(s, (u64)0xFFFFFFFFu);
            if (FREE) { if (nk > 65535ull || s > 65535ull) sm.bail = 1; }
            else { if (nk > (u64)Cfg::CAP || s > (u64)BN_SCAP) sm.bail = 1; }
        }
        __syncthreads();
        const u32 nk = sm.nk, S = sm.S;

        u64 kreg[FREE || NW == 1 ? 1 : Cfg::KPT][NW];   // full keys again for the K > 32 verification
        u64 vreg[EXT ? Cfg::KPT : 1];
        u16 slot_of[FREE ? 1 : Cfg::KPT], occ_idx[EXT ? Cfg::KPT : 1];
        u32 a = 0, e = 0;   // my occurrences [a, e) of the (single) chunk when !FREE
        u32 seen = 0;       // occurrences inserted so far (all chunks)

        for (u32 c0 = 0; c0 < S && !sm.bail; c0 += BN_SCAP) {
            const u32 Sc = min((u32)BN_SCAP, S - c0);
            // ---- k-mer offsets of the chunk's slots (block scan of len - K + 1)
            u32 n4[BN_SPT], tn = 0;
#pragma unroll
            for (int i = 0; i < BN_SPT; ++i) {
                const u32 jl = tid * BN_SPT + i;
                n4[i] = 0;
                if (jl < Sc) n4[i] = (__ldg(slot_ptr<NW>(sm, P, c0 + jl, SW) + (PW - 1)) & 0xFFu) - (u32)k + 1;
                tn += n4[i];
            }
            u32 en, dummy_e, totn, dummy_t;
            block_scan2(tn, 0u, sm.wa, sm.wb, en, dummy_e, totn, dummy_t);
#pragma unroll
            for (int i = 0; i < BN_SPT; ++i) {
                const u32 jl = tid * BN_SPT + i;
                if (jl < Sc) sm.koff[jl] = (u16)en;
                en += n4[i];
            }
            if (tid == 0) {
                sm.koff[Sc] = (u16)totn;
                if (seen + totn > nk) sm.bail = 2;   // inconsistent totals: never count from a corrupt table
            }
            __syncthreads();
            if (sm.bail) break;

            const u32 nkc = totn;
            const u32 q = (nkc + BN_THREADS - 1) / BN_THREADS;   // occurrences per thread in this chunk
            a = tid * q; e = min(nkc, a + q);
            const bool has = a < e;
            Walk<SW> r;
            r.j = 0; r.left = 0; r.pos = 0; r.rid = 0;
            u32 skip = 0;
            if (has) {
                u32 jl = 0;
                for (u32 step = BN_SCAP / 2; step >= 1; step >>= 1)
                    if (jl + step < Sc && sm.koff[jl + step] <= a) jl += step;
                r.j = c0 + jl;
                skip = a - sm.koff[jl];
            }
            if (FREE) {
                // every thread runs the same number of rounds so that the warp reconverges between the
                // (divergent) probe loop and the counter update
                for (u32 it = 0; it < q; ++it) {
                    const bool act = has && (a + it < e);
                    u32 slot = 0;
                    if (act) {
                        u64 key[NW], val;
                        next_kmer<NW, SW, EXT>(r, sm, P, k, padbits, skip, key, val);
                        skip = 0;
                        slot = table_probe<NW>(sm, key);
                    }
                    __syncwarp();
                    if (act) atomicAdd(&sm.cnt[slot], 1u);
                }
            } else if (has) {
#pragma unroll
                for (int i = 0; i < Cfg::KPT; ++i) {
                    if (a + i < e) {
                        u64 key[NW], val = 0;
                        next_kmer<NW, SW, EXT>(r, sm, P, k, padbits, skip, key, val);
                        skip = 0;
                        if (NW > 1) {
#pragma unroll
                            for (int l = 0; l < NW; ++l) kreg[i][l] = key[l];
                        }
                        if (EXT) vreg[i] = val;
                        u32 prev;
                        const u32 slot = table_insert<NW>(sm, key, prev);
                        slot_of[i] = (u16)slot;
                        if (EXT) occ_idx[i] = (u16)prev;
                    }
                }
            }
            seen += nkc;
            __syncthreads();   // koff may be overwritten by the next chunk
        }
        if (!sm.bail && seen != nk && tid == 0) sm.bail = 2;
        __syncthreads();

        if (NW > 1) {
            // ---- the CAS key was a fingerprint: every occurrence checks the full words of its slot
            if (!sm.bail && a < e) {
                bool ok = true;
#pragma unroll
                for (int i = 0; i < Cfg::KPT; ++i) {
                    if (a + i < e) {
#pragma unroll
                        for (int l = 0; l < NW; ++l) ok = ok && (sm.kw[l][slot_of[i]] == kreg[i][l]);
                    }
                }
                if (!ok) atomicOr(&sm.bail, 8u);
            }
            __syncthreads();
        }

        if (sm.bail) {
            // bin goes to the HBM path
            if (tid == 0) {
                P.ovf_list[atomicAdd(P.ovf_count, 1u)] = lb;
                P.bin_rec[4 * (size_t)lb + 0] = 0; P.bin_rec[4 * (size_t)lb + 1] = 0;
                P.bin_rec[4 * (size_t)lb + 2] = 0; P.bin_rec[4 * (size_t)lb + 3] = 0;
            }
            continue;
        }

        // ---- filter my slots, compact the kept (k-mer, count) pairs into the staging area
        u32 kept = 0, occ = 0, keepmask = 0;
#pragma unroll
        for (int i = 0; i < Cfg::SLOTS_PT; ++i) {
            const u32 slot = tid * Cfg::SLOTS_PT + i;
            const u32 c = sm.cnt[slot];
            if (c >= P.lower && c <= P.upper) { keepmask |= 1u << i; ++kept; occ += c; }
        }
        u32 ek, eo, tk, to;
        block_scan2(kept, occ, sm.wa, sm.wb, ek, eo, tk, to);
        if (tid == 0) {
            const u64 sk = tk ? atomicAdd(P.stage_cursor, (u64)tk) : 0;
            const u64 so = (EXT && to) ? atomicAdd(P.stage_cursor + 1, (u64)to) : 0;
            sm.stage_kept = sk; sm.stage_occ = so;
            P.bin_rec[4 * (size_t)lb + 0] = sk; P.bin_rec[4 * (size_t)lb + 1] = tk;
            P.bin_rec[4 * (size_t)lb + 2] = so; P.bin_rec[4 * (size_t)lb + 3] = EXT ? to : 0;
        }
        __syncthreads();
        {
            u64 g = sm.stage_kept + ek;
            u32 lo = eo;   // occurrence offset inside the bin
#pragma unroll
            for (int i = 0; i < Cfg::SLOTS_PT; ++i) {
                const u32 slot = tid * Cfg::SLOTS_PT + i;
                const u32 c = sm.cnt[slot];
                u32 mark = 0xFFFFFFFFu;   // "not kept" for the occurrence pass
                if ((keepmask >> i) & 1) {
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
        }
        if (EXT) {
            // ---- occurrences: (pos, rid) of every occurrence of a kept k-mer, grouped per k-mer
            __syncthreads();
            const u64 so = sm.stage_occ;
#pragma unroll
            for (int i = 0; i < Cfg::KPT; ++i) {
                if (a + i < e) {
                    const u32 slot = slot_of[i];
                    const u32 off = sm.cnt[slot];
                    if (off != 0xFFFFFFFFu) {
                        const u64 p = so + off + occ_idx[i];
                        P.st_pos[p] = (u32)(vreg[i] >> 32);
                        P.st_rid[p] = (int)(u32)vreg[i];
                    }
                }
            }
        }
    }

    __syncthreads();
    for (int i = tid; i < BN_HCAP; i += BN_THREADS)
        if (sm.hist[i]) atomicAdd(&P.histogram[i], (u64)sm.hist[i]);
}

// ---- final positions of the bins: exclusive scan of (kept, occurrences) over the bins; one block ----
__global__ void __launch_bounds__(1024) k_bin_offsets(const u64 *__restrict__ bin_rec, u32 nbins, u64 *__restrict__ fin,
                                                       u64 *__restrict__ cursor, u32 big_from, u32 *__restrict__ big_list,
                                                       u32 *__restrict__ big_count)
{
    __shared__ u64 s_a[32], s_b[32];
    __shared__ u64 carry_a, carry_b;
    if (threadIdx.x == 0) { carry_a = cursor[0]; carry_b = cursor[1]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u32 base = 0; base < nbins; base += 1024) {
        const u32 b = base + threadIdx.x;
        u64 x = 0, y = 0;
        if (b < nbins) { x = bin_rec[4 * (size_t)b + 1]; y = bin_rec[4 * (size_t)b + 3]; }
        if (x > (u64)big_from) big_list[atomicAdd(big_count, 1u)] = b;   // handled by the large gather launch
        u64 ix = x, iy = y;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 p = __shfl_up_sync(0xFFFFFFFFu, ix, d);
            u64 q = __shfl_up_sync(0xFFFFFFFFu, iy, d);
            if (lane >= d) { ix += p; iy += q; }
        }
        if (lane == 31) { s_a[warp] = ix; s_b[warp] = iy; }
        __syncthreads();
        if (warp == 0) {
            u64 p = s_a[lane], q = s_b[lane], ip = p, iq = q;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u64 u = __shfl_up_sync(0xFFFFFFFFu, ip, d);
                u64 v = __shfl_up_sync(0xFFFFFFFFu, iq, d);
                if (lane >= d) { ip += u; iq += v; }
            }
            s_a[lane] = ip - p; s_b[lane] = iq - q;
        }
        __syncthreads();
        const u64 ex = carry_a + s_a[warp] + ix - x, ey = carry_b + s_b[warp] + iy - y;
        if (b < nbins) { fin[2 * (size_t)b] = ex; fin[2 * (size_t)b + 1] = ey; }
        __syncthreads();
        if (threadIdx.x == 1023) { carry_a = ex + x; carry_b = ey + y; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { cursor[0] = carry_a; cursor[1] = carry_b; }
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

// ---- the usual bins (<= 512 kept k-mers): bitonic sort held in registers -----------------------------
// 128 threads x EPT elements (element i = tid * EPT + r).  Compare-exchange partners at distance < EPT are in
// the same thread, at thread distance < 32 they are reached with warp shuffles, and only the last stages
// (thread distance >= 32) go through shared memory: 3 of the 45 stages of a 512-element sort.
constexpr int GS_THREADS = 128;

template <int NW>
struct GsSmem {
    u64 key[NW][GS_THREADS * 4];
    u32 cnt[GS_THREADS * 4], src[GS_THREADS * 4];
    u32 warp[GS_THREADS / 32];
};

// block exclusive scan of one u32 per thread (GS_THREADS threads)
__device__ __forceinline__ u32 gs_scan(u32 v, u32 *warp_tot, u32 &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += t;
    }
    __syncthreads();
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    u32 off = 0;
    total = 0;
#pragma unroll
    for (int w = 0; w < GS_THREADS / 32; ++w) { if (w < warp) off += warp_tot[w]; total += warp_tot[w]; }
    return off + inc - v;
}

template <int NW, bool EXT, int EPT>
__device__ __forceinline__ void gather_small(const BinParams &P, GsSmem<NW> &sm, u32 lb, u32 D, u64 sk, u64 so, u64 fk, u64 fo)
{
    constexpr int N = GS_THREADS * EPT;
    const int tid = threadIdx.x;
    u64 key[EPT][NW];
    u32 cnt[EPT], src[EPT];
    // load my EPT consecutive entries (padding sorts last)
    u32 mysum = 0;
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        const u32 i = tid * EPT + r;
        if (i < D) {
#pragma unroll
            for (int l = 0; l < NW; ++l) key[r][l] = P.st_words[(sk + i) * NW + l];
            cnt[r] = P.st_cnt[sk + i];
        } else {
#pragma unroll
            for (int l = 0; l < NW; ++l) key[r][l] = ~0ull;
            cnt[r] = 0;
        }
        src[r] = mysum;
        mysum += cnt[r];
    }
    if (EXT) {   // occurrence start of every entry in staging order
        u32 tot;
        const u32 ex = gs_scan(mysum, sm.warp, tot);
#pragma unroll
        for (int r = 0; r < EPT; ++r) src[r] += ex;
    }

#pragma unroll
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride < EPT) {
#pragma unroll
                for (int r = 0; r < EPT; ++r) {
                    const int pr = r ^ stride;
                    if (pr > r) {
                        const bool asc = (((tid * EPT + r) & size) == 0);
                        const bool sw = asc ? key_less<NW>(key[pr], key[r]) : key_less<NW>(key[r], key[pr]);
                        if (sw) {
#pragma unroll
                            for (int l = 0; l < NW; ++l) { const u64 t = key[r][l]; key[r][l] = key[pr][l]; key[pr][l] = t; }
                            const u32 c = cnt[r]; cnt[r] = cnt[pr]; cnt[pr] = c;
                            if (EXT) { const u32 x = src[r]; src[r] = src[pr]; src[pr] = x; }
                        }
                    }
                }
            } else {
                const int ts = stride / EPT;   // partner thread distance
                if (ts >= 32) {
                    __syncthreads();
#pragma unroll
                    for (int r = 0; r < EPT; ++r) {
                        const int i = tid * EPT + r;
#pragma unroll
                        for (int l = 0; l < NW; ++l) sm.key[l][i] = key[r][l];
                        sm.cnt[i] = cnt[r];
                        if (EXT) sm.src[i] = src[r];
                    }
                    __syncthreads();
                }
#pragma unroll
                for (int r = 0; r < EPT; ++r) {
                    const int i = tid * EPT + r;
                    u64 pk[NW];
                    u32 pc, ps = 0;
                    if (ts >= 32) {
                        const int j = i ^ stride;
#pragma unroll
                        for (int l = 0; l < NW; ++l) pk[l] = sm.key[l][j];
                        pc = sm.cnt[j];
                        if (EXT) ps = sm.src[j];
                    } else {
#pragma unroll
                        for (int l = 0; l < NW; ++l) pk[l] = __shfl_xor_sync(0xFFFFFFFFu, key[r][l], ts);
                        pc = __shfl_xor_sync(0xFFFFFFFFu, cnt[r], ts);
                        if (EXT) ps = __shfl_xor_sync(0xFFFFFFFFu, src[r], ts);
                    }
                    const bool asc = ((i & size) == 0), lower = ((i & stride) == 0);
                    const bool want_min = (lower == asc);
                    const bool take = want_min ? key_less<NW>(pk, key[r]) : key_less<NW>(key[r], pk);
                    if (take) {
#pragma unroll
                        for (int l = 0; l < NW; ++l) key[r][l] = pk[l];
                        cnt[r] = pc;
                        if (EXT) src[r] = ps;
                    }
                }
            }
        }
    }

    // write in sorted order; occurrence offsets = exclusive scan of the counts in sorted order
    u32 dst[EPT], s2 = 0;
#pragma unroll
    for (int r = 0; r < EPT; ++r) { dst[r] = s2; s2 += cnt[r]; }
    if (EXT) {
        u32 tot;
        const u32 ex = gs_scan(s2, sm.warp, tot);
#pragma unroll
        for (int r = 0; r < EPT; ++r) dst[r] += ex;
    }
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        const u32 i = tid * EPT + r;
        if (i < D) {
#pragma unroll
            for (int l = 0; l < NW; ++l) P.out_words[(fk + i) * NW + l] = key[r][l];
            P.out_cnt[fk + i] = cnt[r];
            if (EXT) {
                P.out_occ_off[fk + i] = fo + dst[r];
                const u64 sp = so + src[r], dd = fo + dst[r];
                for (u32 t = 0; t < cnt[r]; ++t) { P.out_pos[dd + t] = P.st_pos[sp + t]; P.out_rid[dd + t] = P.st_rid[sp + t]; }
            }
        }
    }
    (void)lb;
}

template <int NW, bool EXT>
__global__ void __launch_bounds__(GS_THREADS) k_bin_gather_small(BinParams P)
{
    __shared__ GsSmem<NW> sm;
    const u32 lb = blockIdx.x;
    const u64 sk = P.bin_rec[4 * (size_t)lb + 0];
    const u32 D = (u32)P.bin_rec[4 * (size_t)lb + 1];
    const u64 so = P.bin_rec[4 * (size_t)lb + 2];
    if (D == 0 || D > (u32)(GS_THREADS * 4)) return;
    const u64 fk = P.fin[2 * (size_t)lb], fo = P.fin[2 * (size_t)lb + 1];
    if (D <= GS_THREADS) gather_small<NW, EXT, 1>(P, sm, lb, D, sk, so, fk, fo);
    else if (D <= 2 * GS_THREADS) gather_small<NW, EXT, 2>(P, sm, lb, D, sk, so, fk, fo);
    else gather_small<NW, EXT, 4>(P, sm, lb, D, sk, so, fk, fo);
}

// ---- per-source segment tables of the bins a rank owns (multi-rank) ---------------------------------
// alltot[src][b] = (slots << 40 | k-mers) of bin b as extracted by rank src (all-gathered).  Block src scans
// its row over the owned bins [b_lo, b_lo + tg): exclusive prefix of the slot counts = where the bin starts
// inside the stream received from src.  meta[src] = slots received from src; meta[G + p] = bin_start[p * tg]
// of the LOCAL stream = the boundaries of what this rank sends to rank p (p = 0..G).
__global__ void __launch_bounds__(1024) k_seg_scan(const u64 *__restrict__ alltot, u32 T, u32 b_lo, u32 tg, int nranks,
                                                    const u64 *__restrict__ local_start, u64 *__restrict__ seg_start,
                                                    u64 *__restrict__ meta)
{
    __shared__ u64 s_c[32];
    __shared__ u64 carry_c;
    const int src = blockIdx.x;
    const u64 *row = alltot + (size_t)src * T + b_lo;
    u64 *os = seg_start + (size_t)src * (tg + 1);
    if (threadIdx.x == 0) carry_c = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u32 base = 0; base < tg; base += 1024) {
        const u32 b = base + threadIdx.x;
        const u64 c = b < tg ? (row[b] >> 40) : 0;
        u64 ic = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 x = __shfl_up_sync(0xFFFFFFFFu, ic, d);
            if (lane >= d) ic += x;
        }
        if (lane == 31) s_c[warp] = ic;
        __syncthreads();
        if (warp == 0) {
            u64 x = s_c[lane], ix = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u64 p = __shfl_up_sync(0xFFFFFFFFu, ix, d);
                if (lane >= d) ix += p;
            }
            s_c[lane] = ix - x;
        }
        __syncthreads();
        const u64 ec = carry_c + s_c[warp] + ic - c;
        if (b < tg) os[b] = ec;
        __syncthreads();
        if (threadIdx.x == 1023) carry_c = ec + c;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        os[tg] = carry_c;
        meta[src] = carry_c;
        meta[nranks + src] = local_start[(size_t)src * tg];
        if (src == 0) meta[nranks + nranks] = local_start[(size_t)nranks * tg];
    }
}

// k-mers per owned bin summed over the source ranks, and their grand total (atomicAdd into *owned_total)
__global__ void __launch_bounds__(256) k_sum_kmers(const u64 *__restrict__ alltot, u32 T, u32 b_lo, u32 tg, int nranks,
                                                    u64 *__restrict__ bin_kmers, u64 *__restrict__ owned_total)
{
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    u64 s = 0;
    if (b < tg) {
        for (int src = 0; src < nranks; ++src) s += alltot[(size_t)src * T + b_lo + b] & ((1ull << 40) - 1);
        bin_kmers[b] = s;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(owned_total, s);
}

cudaError_t launch_seg_scan(const u64 *alltot, u32 T, u32 b_lo, u32 tg, int nranks, const u64 *local_start, u64 *seg_start,
                            u64 *meta, u64 *bin_kmers, u64 *owned_total, cudaStream_t s)
{
    k_seg_scan<<<nranks, 1024, 0, s>>>(alltot, T, b_lo, tg, nranks, local_start, seg_start, meta);
    k_sum_kmers<<<(tg + 255) / 256, 256, 0, s>>>(alltot, T, b_lo, tg, nranks, bin_kmers, owned_total);
    return cudaGetLastError();
}

constexpr int GS_CAP = GS_THREADS * 4;             // gather: the usual bins (register bitonic), one CTA each
constexpr int GL_THREADS = 512;                    // gather: listed bins with more kept k-mers (up to the bin capacity)

template <int NW, bool EXT>
static cudaError_t launch_bins_t(const BinParams &P, int sm_count, cudaStream_t s)
{
    constexpr int GL_CAP = NW == 1 ? 8192 : (NW == 2 ? 4096 : 2048);   // power of two >= bin capacity
    static_assert(BinCfg<NW>::CAP <= GL_CAP, "gather capacity");
    const size_t smem = sizeof(BinSmem<NW>);
    cudaError_t e = cudaFuncSetAttribute(k_bin_count<NW, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bin_count<NW, EXT>, BN_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const u32 grid = (u32)std::min<u64>((u64)sm_count * per_sm, std::max<u32>(P.nbins, 1u));
    k_bin_count<NW, EXT><<<grid, BN_THREADS, smem, s>>>(P);
    k_bin_offsets<<<1, 1024, 0, s>>>(P.bin_rec, P.nbins, P.fin, P.cursor, (u32)GS_CAP, P.big_list, P.big_count);
    // gather + sort: a small-footprint launch for the usual bins, a large one for bins with many kept k-mers
    const size_t per_entry = (size_t)8 * NW + 8;
    const size_t smem_l = per_entry * GL_CAP;
    k_bin_gather_small<NW, EXT><<<P.nbins, GS_THREADS, 0, s>>>(P);
    e = cudaFuncSetAttribute(k_bin_gather<NW, EXT, GL_CAP, GL_THREADS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l);
    if (e != cudaSuccess) return e;
    k_bin_gather<NW, EXT, GL_CAP, GL_THREADS, true><<<sm_count, GL_THREADS, smem_l, s>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_bin_count(const BinParams &P, int nwords, bool ext, int sm_count, cudaStream_t s)
{
    if (P.nbins == 0) return cudaSuccess;
    if (nwords == 1) return ext ? launch_bins_t<1, true>(P, sm_count, s) : launch_bins_t<1, false>(P, sm_count, s);
    if (nwords == 2) return ext ? launch_bins_t<2, true>(P, sm_count, s) : launch_bins_t<2, false>(P, sm_count, s);
    return ext ? launch_bins_t<3, true>(P, sm_count, s) : launch_bins_t<3, false>(P, sm_count, s);
}

} // namespace hsk
