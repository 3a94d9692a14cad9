// Stage 1+2: minimizer scan, supermer extraction and owner-hash bucketing, straight from the
// 2-bit packed DnaBuffer.
//
// Replaces the reference's HOT LOOPS A+B: FindKmerDestinationsParallel (src/kmerops.cpp:1010-1041,
// canonical m-mers supermer.hpp:315-342, sliding-window minimum :1058-1073, owner :1044-1047) and
// SupermerEncoder::encode / copy_bits (:1096-1148), plus the per-thread ScatteredSupermers staging
// (:253-358).  Not a translation: the reference walks each read with a deque and materialises one
// int per k-mer; here the whole packed buffer is treated as one flat sequence of 2-bit slots,
// processed in tiles by persistent CTAs:
//
//   A. tile bytes -> shared memory as big-endian 32-bit words (16 bases per word)
//   B. every thread rolls forward/reverse m-mers over 16 consecutive slots, hashes the canonical
//      m-mer of every slot into shared memory, and derives a 16-bit "valid k-mer start" mask from
//      the read offset table
//   C. window minimum over the K-M+1 hashes of every k-mer slot -> bucket id (or INVALID)
//   D. run boundaries (bucket change / validity change / tile edge) -> bitmap -> compacted run list
//   E. one supermer per valid run: count pass accumulates per-CTA per-bucket totals, scatter pass
//      claims (index, word offset) from a shared-memory cursor and writes the length, optional
//      (pos, rid) and the re-packed bases into the bucket's region.
//
// A supermer is a run of consecutive k-mers of one read with the same bucket, stored as
// len (u16 bases) + ceil(len/16) 32-bit words, 16 bases per word from the top bits.  Where the
// reference splits supermers (250-base cap, kmerops.cpp:1120) and how it hashes are free choices:
// only the multiset of k-mers per bucket matters, and a canonical k-mer always lands in the same
// bucket because the bucket is a function of its set of canonical m-mers.
#include "kernels.cuh"

namespace hsk {

struct ExtractSmem {
    u32 wbe[EX_WORDS];                 // bases of the tile, 16 per word, first base in the top bits
    u32 hs[EX_TS + EX_TS / 32 + 8];    // m-mer hashes, padded 1 word per 32 to spread banks
    u16 st[EX_TS];                     // per k-mer slot: bucket or EX_INVALID
    u16 runs[EX_TSK + 8];              // compacted run starts (+ sentinel)
    u32 bm[EX_TS / 32];                // run-boundary bitmap
    u32 woff[EX_TS / 32 + 1];          // exclusive popcount prefix of bm
    u16 vm[EX_THREADS];                // valid-start mask of the thread's 16 slots
    u64 rlo, rhi;                      // reads overlapping the tile
    u32 nruns;
};

__device__ __forceinline__ int hs_idx(int q) { return q + (q >> 5); }

// largest r in [lo, hi] with off[r] <= byte (off is non-decreasing; caller guarantees off[lo] <= byte)
__device__ __forceinline__ u64 find_read(const u64 *__restrict__ off, u64 lo, u64 hi, u64 byte)
{
    while (lo < hi) {
        u64 mid = lo + (hi - lo + 1) / 2;
        if (__ldg(off + mid) <= byte) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Phases A-D for one tile.  On return sm.runs[0..nruns] / sm.st describe the runs of the tile.
__device__ __forceinline__ void tile_runs(ExtractSmem &sm, const ExtractParams &P, u64 tile)
{
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const u64 byte0 = tile * (u64)EX_TILE_BYTES;
    const u64 slot0 = byte0 * 4;

    // ---- A: stage bytes (16-byte vectors, byte-swapped to big-endian words)
    if (tid < EX_WORDS / 4) {
        u64 b = byte0 + (u64)tid * 16;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (b < P.nbytes_padded) v = __ldg(reinterpret_cast<const uint4 *>(P.packed + b));
        sm.wbe[4 * tid + 0] = __byte_perm(v.x, 0, 0x0123);
        sm.wbe[4 * tid + 1] = __byte_perm(v.y, 0, 0x0123);
        sm.wbe[4 * tid + 2] = __byte_perm(v.z, 0, 0x0123);
        sm.wbe[4 * tid + 3] = __byte_perm(v.w, 0, 0x0123);
    }
    if (tid == 0) {
        // reads overlapping [byte0, byte0 + tile bytes): read_off[nreads] = nbytes
        u64 lo = 0;
        if (P.nreads > 0 && byte0 < P.nbytes) lo = find_read(P.read_off, 0, P.nreads - 1, byte0);
        u64 last = byte0 + EX_TILE_BYTES;
        u64 hi = lo;
        if (P.nreads > 0 && byte0 < P.nbytes) hi = find_read(P.read_off, lo, P.nreads - 1, last);
        sm.rlo = lo; sm.rhi = hi;
    }
    __syncthreads();

    // ---- B: rolling canonical m-mer hashes of 16 consecutive slots
    {
        const int m = P.m;
        u64 hi = ((u64)sm.wbe[tid] << 32) | sm.wbe[tid + 1];
        u64 lo = ((u64)sm.wbe[tid + 2] << 32);
        const u64 mask = (m == 32) ? ~0ull : ((1ull << (2 * m)) - 1);
        const int rcs = 2 * (m - 1);
        u64 fwd = 0, rc = 0;
        for (int j = 0; j < m - 1; ++j) {
            u64 c = hi >> 62;
            hi = (hi << 2) | (lo >> 62); lo <<= 2;
            fwd = ((fwd << 2) | c) & mask;
            rc = (rc >> 2) | ((3 - c) << rcs);
        }
#pragma unroll
        for (int j = 0; j < EX_R; ++j) {
            u64 c = hi >> 62;
            hi = (hi << 2) | (lo >> 62); lo <<= 2;
            fwd = ((fwd << 2) | c) & mask;
            rc = (rc >> 2) | ((3 - c) << rcs);
            u64 canon = fwd < rc ? fwd : rc;
            sm.hs[hs_idx(tid * EX_R + j)] = mmer_hash(canon);
        }

        // valid k-mer starts among my 16 slots
        u32 vmask = 0;
        const u64 p0 = slot0 + (u64)tid * EX_R;
        if (tid * EX_R < EX_TSK && P.nreads > 0 && (p0 >> 2) < P.nbytes) {
            u64 r = find_read(P.read_off, sm.rlo, sm.rhi, p0 >> 2);
            u64 rstart = __ldg(P.read_off + r) * 4;
            u64 rnext = __ldg(P.read_off + r + 1) * 4;
            u64 rend = rstart + __ldg(P.read_len + r);
#pragma unroll
            for (int j = 0; j < EX_R; ++j) {
                u64 p = p0 + j;
                while (p >= rnext && r + 1 < P.nreads) {
                    ++r;
                    rstart = rnext;
                    rnext = __ldg(P.read_off + r + 1) * 4;
                    rend = rstart + __ldg(P.read_len + r);
                }
                if (p >= rstart && p + (u64)P.k <= rend) vmask |= 1u << j;
            }
        }
        sm.vm[tid] = (u16)vmask;
    }
    __syncthreads();

    // ---- C: minimizer = window minimum of the hashes of the k-mer's m-mers -> bucket
    {
        const int w = P.k - P.m + 1;
#pragma unroll 4
        for (int j = 0; j < EX_R; ++j) {
            int q = j * EX_THREADS + tid;
            if (q < EX_TSK) {
                u32 st = EX_INVALID;
                if ((sm.vm[q >> 4] >> (q & 15)) & 1) {
                    u32 mn = 0xFFFFFFFFu;
                    for (int x = 0; x < w; ++x) mn = min(mn, sm.hs[hs_idx(q + x)]);
                    st = hash_bucket(mn, P.nbuckets);
                }
                sm.st[q] = (u16)st;
            }
        }
    }
    __syncthreads();

    // ---- D: run boundaries -> bitmap -> compacted run starts
#pragma unroll 4
    for (int j = 0; j < EX_R; ++j) {
        int q = j * EX_THREADS + tid;
        bool b = false;
        if (q < EX_TSK) b = (q == 0) || (sm.st[q] != sm.st[q - 1]);
        u32 bal = __ballot_sync(0xFFFFFFFFu, b);
        if (lane == 0) sm.bm[j * (EX_THREADS / 32) + warp] = bal;
    }
    __syncthreads();
    if (warp == 0) {
        // exclusive prefix of popcounts over EX_TS/32 = 128 words, 4 per lane
        u32 c[4], s = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { c[i] = __popc(sm.bm[lane * 4 + i]); s += c[i]; }
        u32 inc = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += t;
        }
        u32 ex = inc - s;
#pragma unroll
        for (int i = 0; i < 4; ++i) { sm.woff[lane * 4 + i] = ex; ex += c[i]; }
        if (lane == 31) { sm.woff[EX_TS / 32] = inc; sm.nruns = inc; }
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < EX_R; ++j) {
        int q = j * EX_THREADS + tid;
        int wi = j * (EX_THREADS / 32) + warp;
        u32 bits = sm.bm[wi];
        if ((bits >> lane) & 1) sm.runs[sm.woff[wi] + __popc(bits & ((1u << lane) - 1))] = (u16)q;
    }
    if (tid == 0) sm.runs[sm.nruns] = (u16)EX_TSK;
    __syncthreads();
}

// ---- count pass: per-CTA per-bucket totals (supermers, words) + global k-mers per bucket --------
__global__ void __launch_bounds__(EX_THREADS) k_supermer_count(ExtractParams P, uint2 *__restrict__ cta_totals,
                                                                u64 *__restrict__ bucket_kmers)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    ExtractSmem &sm = *reinterpret_cast<ExtractSmem *>(smraw);
    u32 *s_cnt = reinterpret_cast<u32 *>(smraw + sizeof(ExtractSmem));
    u32 *s_words = s_cnt + P.nbuckets;
    u32 *s_kmers = s_words + P.nbuckets;
    for (u32 b = threadIdx.x; b < 3 * P.nbuckets; b += EX_THREADS) s_cnt[b] = 0;
    __syncthreads();

    const u64 t0 = (u64)blockIdx.x * P.tiles_per_cta;
    const u64 t1 = min(t0 + P.tiles_per_cta, P.ntiles);
    for (u64 tile = t0; tile < t1; ++tile) {
        tile_runs(sm, P, tile);
        const u32 nruns = sm.nruns;
        for (u32 j = threadIdx.x; j < nruns; j += EX_THREADS) {
            u32 start = sm.runs[j], end = sm.runs[j + 1];
            u32 b = sm.st[start];
            if (b == EX_INVALID) continue;
            u32 n = end - start;
            u32 len = n + P.k - 1;
            atomicAdd(&s_cnt[b], 1u);
            atomicAdd(&s_words[b], (len + 15) >> 4);
            atomicAdd(&s_kmers[b], n);
        }
        __syncthreads();
    }
    for (u32 b = threadIdx.x; b < P.nbuckets; b += EX_THREADS) {
        cta_totals[(u64)blockIdx.x * P.nbuckets + b] = make_uint2(s_cnt[b], s_words[b]);
        if (s_kmers[b]) atomicAdd(&bucket_kmers[b], (u64)s_kmers[b]);
    }
}

// ---- bucket scan: per bucket, exclusive prefix over CTAs (in place) and bucket totals; then
// exclusive prefix over buckets -> bucket starts.  One block.
__global__ void __launch_bounds__(1024) k_bucket_scan(uint2 *__restrict__ cta_totals, u32 nctas, u32 nbuckets,
                                                       u64 *__restrict__ bucket_count, u64 *__restrict__ bucket_words,
                                                       u64 *__restrict__ bucket_start, u64 *__restrict__ word_start)
{
    __shared__ u64 s_c[32], s_w[32];
    __shared__ u64 carry_c, carry_w;
    if (threadIdx.x == 0) { carry_c = 0; carry_w = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u32 base = 0; base < nbuckets; base += 1024) {
        u32 b = base + threadIdx.x;
        u64 tc = 0, tw = 0;
        if (b < nbuckets) {
            for (u32 c = 0; c < nctas; ++c) {
                uint2 v = cta_totals[(u64)c * nbuckets + b];
                cta_totals[(u64)c * nbuckets + b] = make_uint2((u32)tc, (u32)tw);
                tc += v.x; tw += v.y;
            }
            bucket_count[b] = tc;
            bucket_words[b] = tw;
        }
        // block exclusive scan of (tc, tw)
        u64 ic = tc, iw = tw;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 a = __shfl_up_sync(0xFFFFFFFFu, ic, d);
            u64 e = __shfl_up_sync(0xFFFFFFFFu, iw, d);
            if (lane >= d) { ic += a; iw += e; }
        }
        if (lane == 31) { s_c[warp] = ic; s_w[warp] = iw; }
        __syncthreads();
        if (warp == 0) {
            u64 a = s_c[lane], e = s_w[lane];
            u64 ia = a, ie = e;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u64 x = __shfl_up_sync(0xFFFFFFFFu, ia, d);
                u64 y = __shfl_up_sync(0xFFFFFFFFu, ie, d);
                if (lane >= d) { ia += x; ie += y; }
            }
            s_c[lane] = ia - a; s_w[lane] = ie - e;
        }
        __syncthreads();
        u64 ec = carry_c + s_c[warp] + ic - tc;
        u64 ew = carry_w + s_w[warp] + iw - tw;
        if (b < nbuckets) { bucket_start[b] = ec; word_start[b] = ew; }
        __syncthreads();
        if (threadIdx.x == 1023) { carry_c = ec + tc; carry_w = ew + tw; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { bucket_start[nbuckets] = carry_c; word_start[nbuckets] = carry_w; }
}

// ---- scatter pass ------------------------------------------------------------------------------
template <bool EXT>
__global__ void __launch_bounds__(EX_THREADS) k_supermer_scatter(ExtractParams P, const uint2 *__restrict__ cta_base,
                                                                  const u64 *__restrict__ bucket_start,
                                                                  const u64 *__restrict__ word_start,
                                                                  u16 *__restrict__ out_len, u32 *__restrict__ out_words,
                                                                  u64 *__restrict__ out_ext)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    ExtractSmem &sm = *reinterpret_cast<ExtractSmem *>(smraw);
    u64 *s_cur = reinterpret_cast<u64 *>(smraw + ((sizeof(ExtractSmem) + 7) & ~(size_t)7));
    for (u32 b = threadIdx.x; b < P.nbuckets; b += EX_THREADS) s_cur[b] = 0;
    __syncthreads();
    const uint2 *my_base = cta_base + (u64)blockIdx.x * P.nbuckets;

    const u64 t0 = (u64)blockIdx.x * P.tiles_per_cta;
    const u64 t1 = min(t0 + P.tiles_per_cta, P.ntiles);
    for (u64 tile = t0; tile < t1; ++tile) {
        tile_runs(sm, P, tile);
        const u32 nruns = sm.nruns;
        const u64 slot0 = tile * (u64)EX_TSK;
        for (u32 j = threadIdx.x; j < nruns; j += EX_THREADS) {
            u32 start = sm.runs[j], end = sm.runs[j + 1];
            u32 b = sm.st[start];
            if (b == EX_INVALID) continue;
            u32 n = end - start;
            u32 len = n + P.k - 1;
            u32 nw = (len + 15) >> 4;
            u64 old = atomicAdd(&s_cur[b], (1ull << 40) | (u64)nw);
            uint2 cb = __ldg(my_base + b);
            u64 gi = __ldg(bucket_start + b) + cb.x + (old >> 40);
            u64 gw = __ldg(word_start + b) + cb.y + (old & ((1ull << 40) - 1));
            out_len[gi] = (u16)len;
            if (EXT) {
                u64 p = slot0 + start;
                u64 r = find_read(P.read_off, sm.rlo, sm.rhi, p >> 2);
                u64 pos = p - __ldg(P.read_off + r) * 4;
                u32 rid = (u32)((long long)r + (long long)P.readid_base);
                out_ext[gi] = (pos << 32) | (u64)rid;
            }
            for (u32 x = 0; x < nw; ++x) {
                u32 q = start + 16 * x;
                u32 wv = __funnelshift_l(sm.wbe[(q >> 4) + 1], sm.wbe[q >> 4], 2 * (q & 15));
                u32 rem = len - 16 * x;
                if (rem < 16) wv &= ~0u << (32 - 2 * rem);
                out_words[gw + x] = wv;
            }
        }
        __syncthreads();
    }
}

size_t extract_count_smem(u32 nbuckets) { return sizeof(ExtractSmem) + 3 * (size_t)nbuckets * sizeof(u32); }
size_t extract_scatter_smem(u32 nbuckets) { return ((sizeof(ExtractSmem) + 7) & ~(size_t)7) + (size_t)nbuckets * sizeof(u64); }

cudaError_t launch_supermer_count(const ExtractParams &P, u32 nctas, uint2 *cta_totals, u64 *bucket_kmers, cudaStream_t s)
{
    size_t smem = extract_count_smem(P.nbuckets);
    cudaError_t e = cudaFuncSetAttribute(k_supermer_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_supermer_count<<<nctas, EX_THREADS, smem, s>>>(P, cta_totals, bucket_kmers);
    return cudaGetLastError();
}

cudaError_t launch_bucket_scan(uint2 *cta_totals, u32 nctas, u32 nbuckets, u64 *bucket_count, u64 *bucket_words,
                               u64 *bucket_start, u64 *word_start, cudaStream_t s)
{
    k_bucket_scan<<<1, 1024, 0, s>>>(cta_totals, nctas, nbuckets, bucket_count, bucket_words, bucket_start, word_start);
    return cudaGetLastError();
}

cudaError_t launch_supermer_scatter(const ExtractParams &P, u32 nctas, bool ext, const uint2 *cta_base,
                                    const u64 *bucket_start, const u64 *word_start, u16 *out_len, u32 *out_words,
                                    u64 *out_ext, cudaStream_t s)
{
    size_t smem = extract_scatter_smem(P.nbuckets);
    cudaError_t e;
    if (ext) {
        e = cudaFuncSetAttribute(k_supermer_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_supermer_scatter<true><<<nctas, EX_THREADS, smem, s>>>(P, cta_base, bucket_start, word_start, out_len, out_words, out_ext);
    } else {
        e = cudaFuncSetAttribute(k_supermer_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_supermer_scatter<false><<<nctas, EX_THREADS, smem, s>>>(P, cta_base, bucket_start, word_start, out_len, out_words, out_ext);
    }
    return cudaGetLastError();
}

} // namespace hsk
