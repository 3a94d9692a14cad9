// Stage 4: expansion of supermers back into canonical k-mer words (+ PosInRead/ReadId payload).
//
// Replaces the reference's HOT LOOP C: GatheredSupermer::receive_from_buffer_stage2
// (src/kmerops.cpp:484-521) with TKmer::GetRepKmers / GetTwin / GetRep (include/kmer.hpp:313-340,
// 265-303), and the (pos + i, rid) tagging of kmerops.cpp:507.  The reference walks each supermer
// with a rolling k-mer; here the work is output-centric: a tile of 1024 supermers is scanned in
// shared memory (k-mer and word offsets), then every thread produces one k-mer at a time by a
// load-balanced search over the scanned offsets and a direct funnel-shift extract from the
// big-endian packed words, so that the key planes are written fully coalesced.
#include "kernels.cuh"

namespace hsk {

__device__ __forceinline__ void xp_load4(const ExpandSegment &seg, u64 tile, int k, u32 (&n)[XP_SPT], u32 (&nw)[XP_SPT])
{
    const u64 base = tile * XP_TILE + (u64)threadIdx.x * XP_SPT;
#pragma unroll
    for (int i = 0; i < XP_SPT; ++i) {
        n[i] = 0; nw[i] = 0;
        if (base + i < seg.nsup) {
            u32 len = seg.len[base + i];
            n[i] = len - (u32)k + 1;
            nw[i] = (len + 15) >> 4;
        }
    }
}

__global__ void __launch_bounds__(XP_THREADS) k_expand_tile_sums(ExpandSegment seg, int k, uint2 *__restrict__ tile_sums)
{
    __shared__ u32 s_n[XP_THREADS / 32], s_w[XP_THREADS / 32];
    u32 n[XP_SPT], nw[XP_SPT];
    xp_load4(seg, blockIdx.x, k, n, nw);
    u32 a = 0, b = 0;
#pragma unroll
    for (int i = 0; i < XP_SPT; ++i) { a += n[i]; b += nw[i]; }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        a += __shfl_xor_sync(0xFFFFFFFFu, a, d);
        b += __shfl_xor_sync(0xFFFFFFFFu, b, d);
    }
    if ((threadIdx.x & 31) == 0) { s_n[threadIdx.x >> 5] = a; s_w[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 ta = 0, tb = 0;
        for (int i = 0; i < XP_THREADS / 32; ++i) { ta += s_n[i]; tb += s_w[i]; }
        tile_sums[blockIdx.x] = make_uint2(ta, tb);
    }
}

// exclusive scan of the tile sums (one block, sequential chunks with carry)
__global__ void __launch_bounds__(1024) k_expand_tile_scan(const uint2 *__restrict__ tile_sums, u64 ntiles,
                                                            ulonglong2 *__restrict__ tile_base)
{
    __shared__ u64 s_a[32], s_b[32];
    __shared__ u64 carry_a, carry_b;
    if (threadIdx.x == 0) { carry_a = 0; carry_b = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u64 base = 0; base < ntiles; base += 1024) {
        u64 t = base + threadIdx.x;
        u64 a = 0, b = 0;
        if (t < ntiles) { uint2 v = tile_sums[t]; a = v.x; b = v.y; }
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
        if (t < ntiles) tile_base[t] = make_ulonglong2(ea, eb);
        __syncthreads();
        if (threadIdx.x == 1023) { carry_a = ea + a; carry_b = eb + b; }
        __syncthreads();
    }
}

template <int NW, bool EXT>
__global__ void __launch_bounds__(XP_THREADS) k_expand(ExpandSegment seg, int k, const ulonglong2 *__restrict__ tile_base,
                                                        Planes out, u64 *__restrict__ out_val)
{
    __shared__ u32 s_koff[XP_TILE + 1], s_woff[XP_TILE + 1];
    __shared__ u32 s_wn[XP_THREADS / 32], s_ww[XP_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 tile = blockIdx.x;

    u32 n[XP_SPT], nw[XP_SPT];
    xp_load4(seg, tile, k, n, nw);
    u32 tn = 0, tw = 0;
#pragma unroll
    for (int i = 0; i < XP_SPT; ++i) { tn += n[i]; tw += nw[i]; }
    u32 in_ = tn, iw = tw;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 x = __shfl_up_sync(0xFFFFFFFFu, in_, d);
        u32 y = __shfl_up_sync(0xFFFFFFFFu, iw, d);
        if (lane >= d) { in_ += x; iw += y; }
    }
    if (lane == 31) { s_wn[warp] = in_; s_ww[warp] = iw; }
    __syncthreads();
    u32 bn = 0, bw = 0;
    for (int i = 0; i < warp; ++i) { bn += s_wn[i]; bw += s_ww[i]; }
    u32 en = bn + in_ - tn, ew = bw + iw - tw;
#pragma unroll
    for (int i = 0; i < XP_SPT; ++i) {
        s_koff[tid * XP_SPT + i] = en; s_woff[tid * XP_SPT + i] = ew;
        en += n[i]; ew += nw[i];
    }
    if (tid == XP_THREADS - 1) { s_koff[XP_TILE] = en; s_woff[XP_TILE] = ew; }
    __syncthreads();

    const u32 nk = s_koff[XP_TILE];
    const ulonglong2 tb = tile_base[tile];
    const u64 obase = seg.out_base + tb.x;
    const u32 *__restrict__ words = seg.words + tb.y;
    const int kb = k - 32 * (NW - 1);   // bases in the last word

    for (u32 j = tid; j < nk; j += XP_THREADS) {
        u32 s = 0;
#pragma unroll
        for (u32 step = XP_TILE / 2; step >= 1; step >>= 1)
            if (s_koff[s + step] <= j) s += step;
        const u32 i = j - s_koff[s];
        const u32 wb = s_woff[s];
        const u32 nws = s_woff[s + 1] - wb;
        const u32 a = i >> 4, sh = 2 * (i & 15);
        u32 x[2 * NW + 1];
#pragma unroll
        for (int t = 0; t < 2 * NW + 1; ++t) x[t] = (a + t < nws) ? __ldg(words + wb + a + t) : 0u;
        u64 w[NW];
#pragma unroll
        for (int l = 0; l < NW; ++l) {
            u32 hi = __funnelshift_l(x[2 * l + 1], x[2 * l], sh);
            u32 lo = __funnelshift_l(x[2 * l + 2], x[2 * l + 1], sh);
            w[l] = ((u64)hi << 32) | lo;
        }
        if (kb < 32) w[NW - 1] &= ~0ull << (64 - 2 * kb);
        kmer_canonical<NW>(w, k);
#pragma unroll
        for (int l = 0; l < NW; ++l) out.p[l][obase + j] = w[l];
        if (EXT) out_val[obase + j] = seg.ext[tile * XP_TILE + s] + ((u64)i << 32);
    }
}

cudaError_t launch_expand(const ExpandSegment &seg, int k, int nwords, bool ext, uint2 *tile_sums, ulonglong2 *tile_base,
                          Planes out_keys, u64 *out_val, cudaStream_t s)
{
    if (seg.nsup == 0) return cudaSuccess;
    const u64 ntiles = (seg.nsup + XP_TILE - 1) / XP_TILE;
    k_expand_tile_sums<<<(unsigned)ntiles, XP_THREADS, 0, s>>>(seg, k, tile_sums);
    k_expand_tile_scan<<<1, 1024, 0, s>>>(tile_sums, ntiles, tile_base);
#define HSK_XP(NW_, EXT_) k_expand<NW_, EXT_><<<(unsigned)ntiles, XP_THREADS, 0, s>>>(seg, k, tile_base, out_keys, out_val)
    if (nwords == 1) { if (ext) HSK_XP(1, true); else HSK_XP(1, false); }
    else if (nwords == 2) { if (ext) HSK_XP(2, true); else HSK_XP(2, false); }
    else { if (ext) HSK_XP(3, true); else HSK_XP(3, false); }
#undef HSK_XP
    return cudaGetLastError();
}

} // namespace hsk
