// Stage 5b/5c: run-length counting of the sorted k-mers, [LOWER, UPPER] frequency filter, result
// compaction and the count histogram.
//
// Replaces the reference's HOT LOOP E: count_sorted_kmers (src/kmerops.cpp:1410-1445: serial
// run-length, keep runs with LOWER <= cnt <= UPPER at :1428, copy the run's pos/rid when EXTENSION)
// and the histogram loop of print_kmer_histogram (src/hysortk.cpp:106-113).  Parallel formulation:
// per tile of 2048 sorted keys a head-of-run bitmap is built with coalesced loads; a run's length is
// the distance to the next head (bitmap search inside the tile, one binary search in the sorted array
// for the run that leaves the tile); kept runs are compacted with a tile scan whose bases come from a
// scan over the tile totals.  Entries are emitted in sorted order at a device-side cursor so that the
// batches of one kmer_count call append to one result arena without host synchronisation.
#include "kernels.cuh"

namespace hsk {

constexpr int CT_WORDS = CT_TILE / 32;   // 64 bitmap words per tile
constexpr int CT_HCAP = 2048;            // shared-memory histogram bins

struct CountSmem {
    u32 hb[CT_WORDS];   // head-of-run bitmap
    u64 tail_end;       // global index where the run covering the tile's last key ends
    u32 wsum_c[CT_THREADS / 32], wsum_o[CT_THREADS / 32];
};

template <int NW>
__device__ __forceinline__ bool key_eq(const Planes &k, u64 a, u64 b)
{
    bool eq = true;
#pragma unroll
    for (int l = 0; l < NW; ++l) eq = eq && (k.p[l][a] == k.p[l][b]);
    return eq;
}

// Builds the head bitmap of the tile and the end of the run that crosses the tile's end.
template <int NW>
__device__ __forceinline__ void tile_heads(CountSmem &sm, const CountParams &P, u64 tbase, u32 valid)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int j = 0; j < CT_IPT; ++j) {
        u32 q = j * CT_THREADS + tid;
        bool head = false;
        if (q < valid) {
            u64 g = tbase + q;
            head = (g == 0) || !key_eq<NW>(P.keys, g, g - 1);
        }
        u32 bal = __ballot_sync(0xFFFFFFFFu, head);
        if (lane == 0) sm.hb[j * (CT_THREADS / 32) + warp] = bal;
    }
    if (tid == 0) {
        u64 tend = tbase + valid;
        u64 lo = tend, hi = P.n;
        if (tend < P.n && key_eq<NW>(P.keys, tend, tend - 1)) {
            lo = tend + 1;
            while (lo < hi) {
                u64 mid = lo + (hi - lo) / 2;
                if (key_eq<NW>(P.keys, mid, tend - 1)) lo = mid + 1; else hi = mid;
            }
        }
        sm.tail_end = lo;
    }
    __syncthreads();
}

// length of the run starting at local slot i (which is a head)
__device__ __forceinline__ u64 run_length(const CountSmem &sm, u64 tbase, u32 i)
{
    u32 word = i >> 5;
    u32 bits = sm.hb[word] & ((~0u << 1) << (i & 31));
    while (bits == 0 && ++word < CT_WORDS) bits = sm.hb[word];
    if (word < CT_WORDS) return (u64)(word * 32 + __ffs(bits) - 1 - i);
    return sm.tail_end - (tbase + i);
}

template <int NW>
__global__ void __launch_bounds__(CT_THREADS) k_count_tiles(CountParams P, u64 *__restrict__ tile_counts)
{
    __shared__ CountSmem sm;
    const u64 tbase = (u64)blockIdx.x * CT_TILE;
    const u32 valid = (u32)min((u64)CT_TILE, P.n - tbase);
    tile_heads<NW>(sm, P, tbase, valid);
    const int tid = threadIdx.x;
    u32 c = 0, o = 0;
    u32 bits = (sm.hb[tid >> 2] >> ((tid & 3) * 8)) & 0xFFu;
    while (bits) {
        int b = __ffs(bits) - 1;
        bits &= bits - 1;
        u64 len = run_length(sm, tbase, tid * CT_IPT + b);
        if (len >= P.lower && len <= P.upper) { ++c; o += (u32)len; }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
        o += __shfl_xor_sync(0xFFFFFFFFu, o, d);
    }
    if ((tid & 31) == 0) { sm.wsum_c[tid >> 5] = c; sm.wsum_o[tid >> 5] = o; }
    __syncthreads();
    if (tid == 0) {
        u64 tc = 0, to = 0;
        for (int i = 0; i < CT_THREADS / 32; ++i) { tc += sm.wsum_c[i]; to += sm.wsum_o[i]; }
        tile_counts[2 * (u64)blockIdx.x] = tc;
        tile_counts[2 * (u64)blockIdx.x + 1] = to;
    }
}

// exclusive scan of the tile totals starting at the arena cursor; advances the cursor (one block)
__global__ void __launch_bounds__(1024) k_count_scan(u64 *__restrict__ tile_counts, u64 ntiles, u64 *__restrict__ cursor,
                                                     u64 arena_cap, u64 occ_cap, u32 *__restrict__ err)
{
    __shared__ u64 s_a[32], s_b[32];
    __shared__ u64 carry_a, carry_b;
    if (threadIdx.x == 0) { carry_a = cursor[0]; carry_b = cursor[1]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u64 base = 0; base < ntiles; base += 1024) {
        u64 t = base + threadIdx.x;
        u64 a = 0, b = 0;
        if (t < ntiles) { a = tile_counts[2 * t]; b = tile_counts[2 * t + 1]; }
        u64 ia = a, ib = b;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 x = __shfl_up_sync(0xFFFFFFFFu, ia, d);
            u64 y = __shfl_up_sync(0xFFFFFFFFu, ib, d);
            if (lane >= d) { ia += x; ib += y; }
        }
        if (lane == 31) { s_a[warp] = ia; s_b[warp] = ib; }
        __syncthreads();
        if (warp == 0) {
            u64 x = s_a[lane], y = s_b[lane];
            u64 ix = x, iy = y;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u64 p = __shfl_up_sync(0xFFFFFFFFu, ix, d);
                u64 q = __shfl_up_sync(0xFFFFFFFFu, iy, d);
                if (lane >= d) { ix += p; iy += q; }
            }
            s_a[lane] = ix - x; s_b[lane] = iy - y;
        }
        __syncthreads();
        u64 ea = carry_a + s_a[warp] + ia - a;
        u64 eb = carry_b + s_b[warp] + ib - b;
        if (t < ntiles) { tile_counts[2 * t] = ea; tile_counts[2 * t + 1] = eb; }
        __syncthreads();
        if (threadIdx.x == 1023) { carry_a = ea + a; carry_b = eb + b; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        cursor[0] = carry_a; cursor[1] = carry_b;
        if (carry_a > arena_cap || carry_b > occ_cap) atomicOr(err, 1u);
    }
}

template <int NW, bool EXT>
__global__ void __launch_bounds__(CT_THREADS) k_count_emit(CountParams P, const u64 *__restrict__ tile_base)
{
    __shared__ CountSmem sm;
    __shared__ u32 s_hist[CT_HCAP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < CT_HCAP; i += CT_THREADS) s_hist[i] = 0;
    const u64 tbase = (u64)blockIdx.x * CT_TILE;
    const u32 valid = (u32)min((u64)CT_TILE, P.n - tbase);
    tile_heads<NW>(sm, P, tbase, valid);

    // my 8 consecutive slots: lengths of the kept runs
    u32 keep = 0, c = 0, o = 0;
    u32 lens[CT_IPT];
    u32 bits = (sm.hb[tid >> 2] >> ((tid & 3) * 8)) & 0xFFu;
#pragma unroll
    for (int b = 0; b < CT_IPT; ++b) {
        lens[b] = 0;
        if ((bits >> b) & 1) {
            u64 len = run_length(sm, tbase, tid * CT_IPT + b);
            if (len >= P.lower && len <= P.upper) { keep |= 1u << b; lens[b] = (u32)len; ++c; o += (u32)len; }
        }
    }
    // block exclusive scan of (c, o)
    u32 ic = c, io = o;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 x = __shfl_up_sync(0xFFFFFFFFu, ic, d);
        u32 y = __shfl_up_sync(0xFFFFFFFFu, io, d);
        if (lane >= d) { ic += x; io += y; }
    }
    if (lane == 31) { sm.wsum_c[warp] = ic; sm.wsum_o[warp] = io; }
    __syncthreads();
    u32 bc = 0, bo = 0;
    for (int i = 0; i < warp; ++i) { bc += sm.wsum_c[i]; bo += sm.wsum_o[i]; }
    u64 e = tile_base[2 * (u64)blockIdx.x] + bc + ic - c;
    u64 oc = tile_base[2 * (u64)blockIdx.x + 1] + bo + io - o;

#pragma unroll
    for (int b = 0; b < CT_IPT; ++b) {
        if ((keep >> b) & 1) {
            const u64 g = tbase + (u64)(tid * CT_IPT + b);
            const u32 len = lens[b];
            if (e >= P.arena_cap || (EXT && oc + len > P.occ_cap)) { ++e; oc += len; continue; }   // reported by k_count_scan
#pragma unroll
            for (int l = 0; l < NW; ++l) P.out_words[e * NW + l] = P.keys.p[l][g];
            P.out_cnt[e] = len;
            if (len < CT_HCAP) atomicAdd(&s_hist[len], 1u); else atomicAdd(&P.histogram[len], 1ull);
            if (EXT) {
                P.out_occ_off[e] = oc;
                for (u32 t = 0; t < len; ++t) {
                    u64 v = P.val[g + t];
                    P.out_pos[oc + t] = (u32)(v >> 32);
                    P.out_rid[oc + t] = (int)(u32)v;
                }
                oc += len;
            }
            ++e;
        }
    }
    __syncthreads();
    for (int i = tid; i < CT_HCAP; i += CT_THREADS)
        if (s_hist[i]) atomicAdd(&P.histogram[i], (u64)s_hist[i]);
}

size_t count_scratch_bytes(u64 n)
{
    u64 ntiles = (n + CT_TILE - 1) / CT_TILE;
    return (size_t)(2 * (ntiles + 1)) * sizeof(u64);
}

cudaError_t launch_count_filter(const CountParams &P, void *scratch, cudaStream_t s)
{
    if (P.n == 0) return cudaSuccess;
    const u64 ntiles = (P.n + CT_TILE - 1) / CT_TILE;
    u64 *tiles = reinterpret_cast<u64 *>(scratch);
    const bool ext = P.val != nullptr;
    if (P.nwords == 1) k_count_tiles<1><<<(unsigned)ntiles, CT_THREADS, 0, s>>>(P, tiles);
    else if (P.nwords == 2) k_count_tiles<2><<<(unsigned)ntiles, CT_THREADS, 0, s>>>(P, tiles);
    else k_count_tiles<3><<<(unsigned)ntiles, CT_THREADS, 0, s>>>(P, tiles);
    k_count_scan<<<1, 1024, 0, s>>>(tiles, ntiles, P.cursor, P.arena_cap, P.occ_cap, P.err);
#define HSK_CT(NW_, EXT_) k_count_emit<NW_, EXT_><<<(unsigned)ntiles, CT_THREADS, 0, s>>>(P, tiles)
    if (P.nwords == 1) { if (ext) HSK_CT(1, true); else HSK_CT(1, false); }
    else if (P.nwords == 2) { if (ext) HSK_CT(2, true); else HSK_CT(2, false); }
    else { if (ext) HSK_CT(3, true); else HSK_CT(3, false); }
#undef HSK_CT
    return cudaGetLastError();
}

} // namespace hsk
