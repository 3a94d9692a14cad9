// Stages 4+5 fused, on chip: one CTA takes one bin of supermers and expands it to canonical k-mers,
// sorts them, run-length counts, filters and emits — without the k-mers ever touching HBM.
//
// Replaces, for bins that fit in shared memory, the reference's HOT LOOPS C+D+E:
//   receive_from_buffer_stage2 + GetRepKmers  (src/kmerops.cpp:484-521, include/kmer.hpp:313-340)
//   sort_task -> RADULS / PARADIS             (src/kmerops.cpp:1382-1407)
//   count_sorted_kmers                        (src/kmerops.cpp:1410-1445), histogram (hysortk.cpp:106-113)
// Like RADULS (an MSD radix sort that finishes small buckets with small sorts, raduls.h:26-27) the
// sort is MSD-first, but staged entirely in shared memory: one counting pass on a 10-12 bit digit
// (shared-memory atomics, no stability needed for MSD), an exclusive scan, a scatter into sub-bucket
// order, then every thread finishes the handful of keys of its own consecutive sub-buckets with a
// selection sort over the DISTINCT keys (the k-mers of a bin are ~30x duplicated at 30x coverage, so
// cost is n * distinct, not n^2) and run-length counts them in place.
// The k-mers of one bin share a few minimizers, i.e. they are overlapping windows of a few genomic
// loci: their leading bases take only a few hundred values, so the MSD digit is taken from a
// bijective scramble of the first key word (odd multiplier); the bin is therefore sorted by
// (digit of scrambled key, key) — a total order under which equal k-mers are adjacent, which is all
// that counting needs.  Bins are sized by the extraction stage so that they fit (CAP k-mers); the rare
// bin that does not (skew) is reported in an overflow list and goes through the HBM path
// (expand.cu -> radix.cu -> count.cu).
//
// Output order is deterministic: bins are taken in index order through a ticket, and every bin
// resolves its position in the result arena by decoupled look-back over the preceding bins, so the
// arena holds the bins in index order and sorted k-mers inside each bin (the reference: per-task
// sorted runs in task order, kmerops.cpp:883-904).
#include "kernels.cuh"

#include <algorithm>

namespace hsk {

constexpr int BN_SPT = BN_SCAP / BN_THREADS;   // supermers per thread in the table scan
constexpr int BN_HCAP = 2048;                  // shared-memory histogram bins
constexpr u64 LBF_AGG = 1ull << 62, LBF_INC = 2ull << 62, LBF_MASK = (1ull << 62) - 1;

template <int NW, bool EXT>
struct BinCfg {
    static constexpr int REC = 8 * NW + (EXT ? 8 : 0);
    static constexpr int KPT = REC == 8 ? 12 : (REC == 16 ? 6 : (REC == 24 ? 4 : 3));
    static constexpr int CAP = BN_THREADS * KPT;
    static constexpr int NB_BITS = CAP >= 4096 ? 12 : (CAP >= 2048 ? 11 : 10);
    static constexpr int NB = 1 << NB_BITS;
    static constexpr int BPT = NB / BN_THREADS;
};

int bin_capacity(int nwords, bool ext)
{
    const int rec = 8 * nwords + (ext ? 8 : 0);
    return BN_THREADS * (rec == 8 ? 12 : (rec == 16 ? 6 : (rec == 24 ? 4 : 3)));
}

template <int NW, bool EXT>
struct BinSmem {
    u64 keys[NW][BinCfg<NW, EXT>::CAP];
    u64 val[EXT ? BinCfg<NW, EXT>::CAP : 1];
    u32 cnt[BinCfg<NW, EXT>::NB + 1];
    u32 cplx[BinCfg<NW, EXT>::NB / 32];   // sub-buckets holding more than one distinct key
    u32 woff[BN_SCAP + 1];
    u16 koff[BN_SCAP + 2];
    u8 ssrc[BN_SCAP];
    u32 hist[BN_HCAP];
    u64 src_i0[BN_MAX_SRC], src_w0[BN_MAX_SRC];
    u32 src_n[BN_MAX_SRC], src_sbase[BN_MAX_SRC + 1], src_wbase[BN_MAX_SRC + 1];
    u32 wa[BN_THREADS / 32], wb[BN_THREADS / 32];
    u64 gbase_kept, gbase_occ;
    u32 tot_kept, tot_occ;
    u32 bin, nk, S, bail;
};

__device__ __forceinline__ u64 ldv64(const u64 *p) { return *reinterpret_cast<const volatile u64 *>(p); }
__device__ __forceinline__ void stv64(u64 *p, u64 v) { *reinterpret_cast<volatile u64 *>(p) = v; }

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

// sub-bucket of a key: top bits of a bijective scramble of its first word
template <int BITS>
__device__ __forceinline__ u32 sub_bucket(u64 w0) { return (u32)((w0 * 0x9E3779B97F4A7C15ull) >> (64 - BITS)); }

template <int NW>
__device__ __forceinline__ bool key_less(const u64 (&a)[NW], const u64 (&b)[NW])
{
#pragma unroll
    for (int l = 0; l < NW; ++l) {
        if (a[l] != b[l]) return a[l] < b[l];
    }
    return false;
}

template <int NW, bool EXT>
__global__ void __launch_bounds__(BN_THREADS, 2) k_bin_sort_count(BinParams P)
{
    using Cfg = BinCfg<NW, EXT>;
    extern __shared__ __align__(16) unsigned char smraw[];
    BinSmem<NW, EXT> &sm = *reinterpret_cast<BinSmem<NW, EXT> *>(smraw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = P.k;
    const int padbits = 2 * (32 * NW - k);

    for (int i = tid; i < BN_HCAP; i += BN_THREADS) sm.hist[i] = 0;

    while (true) {
        __syncthreads();   // end of the previous bin: shared memory is free again
        if (tid == 0) { sm.bin = atomicAdd(P.ticket, 1u); sm.bail = 0; }
        for (int i = tid; i <= Cfg::NB; i += BN_THREADS) sm.cnt[i] = 0;
        for (int i = tid; i < Cfg::NB / 32; i += BN_THREADS) sm.cplx[i] = 0;
        __syncthreads();
        const u32 lb = sm.bin;
        if (lb >= P.nbins) break;

        // ---- bin descriptor: one segment of supermers per source rank
        if (tid < P.nsrc) {
            const u64 i0 = P.seg_start[tid][lb], i1 = P.seg_start[tid][lb + 1];
            sm.src_i0[tid] = i0;
            sm.src_n[tid] = (u32)min(i1 - i0, (u64)0xFFFFFFFFu);
            sm.src_w0[tid] = P.seg_wstart[tid][lb];
        }
        __syncthreads();
        if (tid == 0) {
            u64 s = 0;
            for (int i = 0; i < P.nsrc; ++i) { sm.src_sbase[i] = (u32)min(s, (u64)0xFFFFFFFFu); s += sm.src_n[i]; }
            sm.src_sbase[P.nsrc] = (u32)min(s, (u64)0xFFFFFFFFu);
            const u64 nk = P.bin_kmers[lb];
            sm.nk = (u32)min(nk, (u64)0xFFFFFFFFu);
            sm.S = (u32)min(s, (u64)0xFFFFFFFFu);
            if (nk > (u64)Cfg::CAP || s > (u64)BN_SCAP) sm.bail = 1;
        }
        __syncthreads();
        const u32 nk = sm.nk, S = sm.S;

        if (!sm.bail) {
            // ---- supermer table: k-mer and word offsets of every supermer of the bin
            u32 n4[BN_SPT], w4[BN_SPT], tn = 0, tw = 0;
#pragma unroll
            for (int i = 0; i < BN_SPT; ++i) {
                const u32 j = tid * BN_SPT + i;
                n4[i] = 0; w4[i] = 0;
                if (j < S) {
                    int s = 0;
                    while (j >= sm.src_sbase[s + 1]) ++s;
                    const u32 len = P.len[s][sm.src_i0[s] + (j - sm.src_sbase[s])];
                    n4[i] = len - (u32)k + 1;
                    w4[i] = (len + 15) >> 4;
                    sm.ssrc[j] = (u8)s;
                }
                tn += n4[i]; tw += w4[i];
            }
            u32 en, ew, totn, totw;
            block_scan2(tn, tw, sm.wa, sm.wb, en, ew, totn, totw);
#pragma unroll
            for (int i = 0; i < BN_SPT; ++i) {
                const u32 j = tid * BN_SPT + i;
                if (j < S) { sm.koff[j] = (u16)en; sm.woff[j] = ew; }
                en += n4[i]; ew += w4[i];
            }
            if (tid == 0) {
                sm.koff[S] = (u16)totn; sm.woff[S] = totw;
                if (totn != nk) sm.bail = 2;   // inconsistent totals: never emit from a corrupt table
            }
            __syncthreads();
            if (tid < P.nsrc) sm.src_wbase[tid] = sm.woff[min(sm.src_sbase[tid], S)];
            __syncthreads();
        }

        u64 kreg[Cfg::KPT][NW];
        u64 vreg[EXT ? Cfg::KPT : 1];
        u16 rank[Cfg::KPT];
        const u32 q = (nk + BN_THREADS - 1) / BN_THREADS;   // k-mers per thread, <= KPT
        const u32 a = tid * q, e = min(nk, a + q);

        if (!sm.bail && a < e) {
            // ---- expansion: my q consecutive k-mers, rolling inside a supermer
            u32 j = 0;
            for (u32 step = BN_SCAP / 2; step >= 1; step >>= 1)
                if (j + step < S && sm.koff[j + step] <= a) j += step;
            u32 o = a - sm.koff[j];
            u64 fwd[NW], rc[NW];
            const u32 *wp = nullptr;
            u32 nj = 0;
            u64 extv = 0;
            bool fresh = true;
#pragma unroll
            for (int i = 0; i < Cfg::KPT; ++i) {
                if (a + i < e) {
                    if (fresh) {
                        const int s = sm.ssrc[j];
                        wp = P.words[s] + sm.src_w0[s] + (sm.woff[j] - sm.src_wbase[s]);
                        nj = (u32)sm.koff[j + 1] - (u32)sm.koff[j];
                        if (EXT) extv = P.ext[s][sm.src_i0[s] + (j - sm.src_sbase[s])];
                        const u32 nws = sm.woff[j + 1] - sm.woff[j];
                        const u32 wi = o >> 4, sh = 2 * (o & 15);
                        u32 x[2 * NW + 1];
#pragma unroll
                        for (int t = 0; t < 2 * NW + 1; ++t) x[t] = (wi + t < nws) ? __ldg(wp + wi + t) : 0u;
#pragma unroll
                        for (int l = 0; l < NW; ++l) {
                            u32 hi = __funnelshift_l(x[2 * l + 1], x[2 * l], sh);
                            u32 lo = __funnelshift_l(x[2 * l + 2], x[2 * l + 1], sh);
                            fwd[l] = ((u64)hi << 32) | lo;
                        }
                        if (padbits) fwd[NW - 1] &= ~0ull << padbits;
                        kmer_twin<NW>(fwd, k, rc);
                        fresh = false;
                    } else {
                        // roll: drop the first base, append base (o + k - 1) of the supermer
                        const u32 bo = o + (u32)k - 1;
                        const u64 c = (__ldg(wp + (bo >> 4)) >> (30 - 2 * (bo & 15))) & 3u;
#pragma unroll
                        for (int l = 0; l < NW; ++l) {
                            fwd[l] <<= 2;
                            if (l + 1 < NW) fwd[l] |= fwd[l + 1] >> 62;
                        }
                        fwd[NW - 1] |= c << padbits;
#pragma unroll
                        for (int l = NW - 1; l >= 0; --l) {
                            rc[l] >>= 2;
                            if (l > 0) rc[l] |= rc[l - 1] << 62;
                        }
                        rc[0] |= (3 - c) << 62;
                        if (padbits) rc[NW - 1] &= ~0ull << padbits;
                    }
                    const bool use_rc = key_less<NW>(rc, fwd);
#pragma unroll
                    for (int l = 0; l < NW; ++l) kreg[i][l] = use_rc ? rc[l] : fwd[l];
                    if (EXT) vreg[i] = extv + ((u64)o << 32);
                    rank[i] = (u16)atomicAdd(&sm.cnt[sub_bucket<Cfg::NB_BITS>(kreg[i][0])], 1u);
                    ++o;
                    if (o >= nj) { ++j; o = 0; fresh = true; }
                }
            }
        }
        __syncthreads();

        // ---- exclusive scan of the sub-bucket counters; my consecutive sub-buckets = my sort range
        u32 rs = 0, re = 0;
        {
            u32 c[Cfg::BPT], sum = 0;
#pragma unroll
            for (int i = 0; i < Cfg::BPT; ++i) { c[i] = sm.cnt[tid * Cfg::BPT + i]; sum += c[i]; }
            u32 ex, dummy_e, tot, dummy_t;
            block_scan2(sum, 0u, sm.wa, sm.wb, ex, dummy_e, tot, dummy_t);
            rs = ex; re = ex + sum;
            (void)rs; (void)re;
#pragma unroll
            for (int i = 0; i < Cfg::BPT; ++i) { sm.cnt[tid * Cfg::BPT + i] = ex; ex += c[i]; }
            if (tid == BN_THREADS - 1) sm.cnt[Cfg::NB] = ex;
        }
        __syncthreads();

        // ---- scatter into sub-bucket order
        if (!sm.bail && a < e) {
#pragma unroll
            for (int i = 0; i < Cfg::KPT; ++i) {
                if (a + i < e) {
                    const u32 pos = sm.cnt[sub_bucket<Cfg::NB_BITS>(kreg[i][0])] + rank[i];
#pragma unroll
                    for (int l = 0; l < NW; ++l) sm.keys[l][pos] = kreg[i][l];
                    if (EXT) sm.val[pos] = vreg[i];
                }
            }
        }
        __syncthreads();

        if (sm.bail) {
            // bin goes to the HBM path; it contributes nothing here but must not block its successors
            if (tid == 0) {
                P.ovf_list[atomicAdd(P.ovf_count, 1u)] = lb;
                const u64 f = (lb == 0) ? LBF_INC : LBF_AGG;
                stv64(P.lb_occ + lb, f);
                stv64(P.lb_kept + lb, f);
                if (lb == P.nbins - 1) {   // still has to close the arena cursor: needs the prefix
                    u64 ek = 0, eo = 0;
                    for (long t = (long)lb - 1; t >= 0; --t) {
                        u64 vk, vo;
                        do { vk = ldv64(P.lb_kept + t); vo = ldv64(P.lb_occ + t); } while ((vk >> 62) == 0 || (vk >> 62) != (vo >> 62));
                        ek += vk & LBF_MASK; eo += vo & LBF_MASK;
                        if ((vk >> 62) == 2) break;
                    }
                    P.cursor[0] = ek; P.cursor[1] = eo;
                }
            }
            continue;
        }

        // ---- which sub-buckets hold more than one distinct key?  One comparison per k-mer, balanced:
        // every position checks its key against the first key of its sub-bucket.
        for (u32 p = tid; p < nk; p += BN_THREADS) {
            const u32 d = sub_bucket<Cfg::NB_BITS>(sm.keys[0][p]);
            const u32 h = sm.cnt[d];
            bool same = true;
#pragma unroll
            for (int l = 0; l < NW; ++l) same = same && (sm.keys[l][p] == sm.keys[l][h]);
            if (!same) atomicOr(&sm.cplx[d >> 5], 1u << (d & 31));
        }
        __syncthreads();

        // ---- my sub-buckets: a simple one is a single run (its size is the count); a complex one is
        // finished with a selection sort over its distinct keys (each round finds the smallest remaining
        // key and gathers its copies to the front).  Complex sub-buckets are re-marked as sorted.
        u32 kept = 0, occ = 0;
#pragma unroll 1
        for (int sb = 0; sb < Cfg::BPT; ++sb) {
            const u32 d = tid * Cfg::BPT + sb;
            const u32 s0 = sm.cnt[d], s1 = sm.cnt[d + 1];
            if (s1 == s0) continue;
            if (!((sm.cplx[d >> 5] >> (d & 31)) & 1)) {
                const u32 c = s1 - s0;
                if (c >= P.lower && c <= P.upper) { ++kept; occ += c; }
                continue;
            }
            for (u32 i = s0; i < s1;) {
                u64 mn[NW];
#pragma unroll
                for (int l = 0; l < NW; ++l) mn[l] = sm.keys[l][i];
                for (u32 t = i + 1; t < s1; ++t) {
                    u64 y[NW];
#pragma unroll
                    for (int l = 0; l < NW; ++l) y[l] = sm.keys[l][t];
                    if (key_less<NW>(y, mn)) {
#pragma unroll
                        for (int l = 0; l < NW; ++l) mn[l] = y[l];
                    }
                }
                u32 j = i;
                for (u32 t = i; t < s1; ++t) {
                    bool eq = true;
#pragma unroll
                    for (int l = 0; l < NW; ++l) eq = eq && (sm.keys[l][t] == mn[l]);
                    if (eq) {
                        if (t != j) {
#pragma unroll
                            for (int l = 0; l < NW; ++l) { sm.keys[l][t] = sm.keys[l][j]; sm.keys[l][j] = mn[l]; }
                            if (EXT) { const u64 v = sm.val[t]; sm.val[t] = sm.val[j]; sm.val[j] = v; }
                        }
                        ++j;
                    }
                }
                const u32 c = j - i;
                if (c >= P.lower && c <= P.upper) { ++kept; occ += c; }
                i = j;
            }
        }
        u32 ek, eo, tk, to;
        block_scan2(kept, occ, sm.wa, sm.wb, ek, eo, tk, to);

        // ---- position of the bin in the arena: decoupled look-back over the preceding bins
        if (warp == 0) {
            u64 bk = 0, bo = 0;
            if (lb == 0) {
                if (lane == 0) { stv64(P.lb_occ, LBF_INC | to); stv64(P.lb_kept, LBF_INC | tk); }
            } else {
                if (lane == 0) { stv64(P.lb_occ + lb, LBF_AGG | to); stv64(P.lb_kept + lb, LBF_AGG | tk); }
                long look = (long)lb - 1;
                while (true) {
                    const long idx = look - lane;
                    u64 vk = LBF_INC, vo = LBF_INC;
                    if (idx >= 0) {
                        do { vk = ldv64(P.lb_kept + idx); vo = ldv64(P.lb_occ + idx); } while ((vk >> 62) == 0 || (vk >> 62) != (vo >> 62));
                    }
                    const u32 inc = __ballot_sync(0xFFFFFFFFu, (vk >> 62) == 2);
                    const int first = inc ? __ffs(inc) - 1 : 32;
                    u64 ck = (lane <= first) ? (vk & LBF_MASK) : 0, co = (lane <= first) ? (vo & LBF_MASK) : 0;
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) {
                        ck += __shfl_xor_sync(0xFFFFFFFFu, ck, d);
                        co += __shfl_xor_sync(0xFFFFFFFFu, co, d);
                    }
                    bk += ck; bo += co;
                    if (inc) break;
                    look -= 32;
                }
                if (lane == 0) { stv64(P.lb_occ + lb, LBF_INC | (bo + to)); stv64(P.lb_kept + lb, LBF_INC | (bk + tk)); }
            }
            if (lane == 0) {
                sm.gbase_kept = bk; sm.gbase_occ = bo;
                if (lb == P.nbins - 1) { P.cursor[0] = bk + tk; P.cursor[1] = bo + to; }
            }
        }
        __syncthreads();

        // ---- emit the kept runs of my sub-buckets, in order
        {
            u64 g = sm.gbase_kept + ek, go = sm.gbase_occ + eo;
#pragma unroll 1
            for (int sb = 0; sb < Cfg::BPT; ++sb) {
                const u32 d = tid * Cfg::BPT + sb;
                const u32 s0 = sm.cnt[d], s1 = sm.cnt[d + 1];
                const bool complex_sb = (sm.cplx[d >> 5] >> (d & 31)) & 1;
                for (u32 i = s0; i < s1;) {
                    u32 r = s1;
                    if (complex_sb) {
                        r = i + 1;
                        while (r < s1) {
                            bool eq = true;
#pragma unroll
                            for (int l = 0; l < NW; ++l) eq = eq && (sm.keys[l][r] == sm.keys[l][i]);
                            if (!eq) break;
                            ++r;
                        }
                    }
                    const u32 c = r - i;
                    if (c >= P.lower && c <= P.upper) {
#pragma unroll
                        for (int l = 0; l < NW; ++l) P.out_words[g * NW + l] = sm.keys[l][i];
                        P.out_cnt[g] = c;
                        if (c < (u32)BN_HCAP) atomicAdd(&sm.hist[c], 1u); else atomicAdd(&P.histogram[c], 1ull);
                        if (EXT) {
                            P.out_occ_off[g] = go;
                            for (u32 t = 0; t < c; ++t) {
                                const u64 v = sm.val[i + t];
                                P.out_pos[go + t] = (u32)(v >> 32);
                                P.out_rid[go + t] = (int)(u32)v;
                            }
                            go += c;
                        }
                        ++g;
                    }
                    i = r;
                }
            }
        }
    }

    __syncthreads();
    for (int i = tid; i < BN_HCAP; i += BN_THREADS)
        if (sm.hist[i]) atomicAdd(&P.histogram[i], (u64)sm.hist[i]);
}

// ---- per-source segment tables of the bins a rank owns (multi-rank) ---------------------------------
// alltot[src][0][b] = k-mers, alltot[src][1][b] = (supermers << 32 | words) of bin b as extracted by rank src
// (all-gathered).  Block src scans its row over the owned bins [b_lo, b_lo + tg): exclusive prefixes of
// supermers / words = where the bin starts inside the stream received from src.  meta[2*src..] = totals
// received from src; meta[2*G + 2*p..] = (bin_start, word_start)[p * tg] of the LOCAL streams = the
// boundaries of what this rank sends to rank p.
__global__ void __launch_bounds__(1024) k_seg_scan(const u64 *__restrict__ alltot, u32 T, u32 b_lo, u32 tg, int nranks,
                                                    const u64 *__restrict__ local_start, const u64 *__restrict__ local_wstart,
                                                    u64 *__restrict__ seg_start, u64 *__restrict__ seg_wstart,
                                                    u64 *__restrict__ meta)
{
    __shared__ u64 s_c[32], s_w[32];
    __shared__ u64 carry_c, carry_w;
    const int src = blockIdx.x;
    const u64 *cw = alltot + ((size_t)src * 2 + 1) * T + b_lo;
    u64 *os = seg_start + (size_t)src * (tg + 1), *ow = seg_wstart + (size_t)src * (tg + 1);
    if (threadIdx.x == 0) { carry_c = 0; carry_w = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u32 base = 0; base < tg; base += 1024) {
        const u32 b = base + threadIdx.x;
        const u64 v = b < tg ? cw[b] : 0;
        const u64 c = v >> 32, w = v & 0xFFFFFFFFull;
        u64 ic = c, iw = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 x = __shfl_up_sync(0xFFFFFFFFu, ic, d);
            u64 y = __shfl_up_sync(0xFFFFFFFFu, iw, d);
            if (lane >= d) { ic += x; iw += y; }
        }
        if (lane == 31) { s_c[warp] = ic; s_w[warp] = iw; }
        __syncthreads();
        if (warp == 0) {
            u64 x = s_c[lane], y = s_w[lane], ix = x, iy = y;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u64 p = __shfl_up_sync(0xFFFFFFFFu, ix, d);
                u64 q = __shfl_up_sync(0xFFFFFFFFu, iy, d);
                if (lane >= d) { ix += p; iy += q; }
            }
            s_c[lane] = ix - x; s_w[lane] = iy - y;
        }
        __syncthreads();
        const u64 ec = carry_c + s_c[warp] + ic - c, ew = carry_w + s_w[warp] + iw - w;
        if (b < tg) { os[b] = ec; ow[b] = ew; }
        __syncthreads();
        if (threadIdx.x == 1023) { carry_c = ec + c; carry_w = ew + w; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        os[tg] = carry_c; ow[tg] = carry_w;
        meta[2 * src] = carry_c; meta[2 * src + 1] = carry_w;
        meta[2 * nranks + 2 * src] = local_start[(size_t)src * tg];
        meta[2 * nranks + 2 * src + 1] = local_wstart[(size_t)src * tg];
        if (src == 0) {
            meta[2 * nranks + 2 * nranks] = local_start[(size_t)nranks * tg];
            meta[2 * nranks + 2 * nranks + 1] = local_wstart[(size_t)nranks * tg];
        }
    }
}

// k-mers per owned bin summed over the source ranks, and their grand total (atomicAdd into *owned_total)
__global__ void __launch_bounds__(256) k_sum_kmers(const u64 *__restrict__ alltot, u32 T, u32 b_lo, u32 tg, int nranks,
                                                    u64 *__restrict__ bin_kmers, u64 *__restrict__ owned_total)
{
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    u64 s = 0;
    if (b < tg) {
        for (int src = 0; src < nranks; ++src) s += alltot[((size_t)src * 2) * T + b_lo + b];
        bin_kmers[b] = s;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(owned_total, s);
}

cudaError_t launch_seg_scan(const u64 *alltot, u32 T, u32 b_lo, u32 tg, int nranks, const u64 *local_start,
                            const u64 *local_wstart, u64 *seg_start, u64 *seg_wstart, u64 *meta, u64 *bin_kmers,
                            u64 *owned_total, cudaStream_t s)
{
    k_seg_scan<<<nranks, 1024, 0, s>>>(alltot, T, b_lo, tg, nranks, local_start, local_wstart, seg_start, seg_wstart, meta);
    k_sum_kmers<<<(tg + 255) / 256, 256, 0, s>>>(alltot, T, b_lo, tg, nranks, bin_kmers, owned_total);
    return cudaGetLastError();
}

template <int NW, bool EXT>
static cudaError_t launch_bins_t(const BinParams &P, int sm_count, cudaStream_t s)
{
    const size_t smem = sizeof(BinSmem<NW, EXT>);
    cudaError_t e = cudaFuncSetAttribute(k_bin_sort_count<NW, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bin_sort_count<NW, EXT>, BN_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const u32 grid = (u32)std::min<u64>((u64)sm_count * per_sm, std::max<u32>(P.nbins, 1u));
    k_bin_sort_count<NW, EXT><<<grid, BN_THREADS, smem, s>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_bin_sort_count(const BinParams &P, int nwords, bool ext, int sm_count, cudaStream_t s)
{
    if (P.nbins == 0) return cudaSuccess;
    if (nwords == 1) return ext ? launch_bins_t<1, true>(P, sm_count, s) : launch_bins_t<1, false>(P, sm_count, s);
    if (nwords == 2) return ext ? launch_bins_t<2, true>(P, sm_count, s) : launch_bins_t<2, false>(P, sm_count, s);
    return ext ? launch_bins_t<3, true>(P, sm_count, s) : launch_bins_t<3, false>(P, sm_count, s);
}

} // namespace hsk
