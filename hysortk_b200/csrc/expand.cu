// Stage 4, HBM path: expansion of supermer slots back into canonical k-mer words (+ PosInRead/ReadId
// payload) in key planes, for the bins the on-chip path (bins.cu) leaves over.
//
// Replaces the reference's HOT LOOP C: GatheredSupermer::receive_from_buffer_stage2
// (src/kmerops.cpp:484-521) with TKmer::GetRepKmers / GetTwin / GetRep (include/kmer.hpp:313-340,
// 265-303), and the (pos + i, rid) tagging of kmerops.cpp:507.  The reference walks each supermer
// with a rolling k-mer; here the work is output-centric: a tile of 1024 slots is scanned in shared
// memory (k-mer offsets), then every thread produces one k-mer at a time by a load-balanced search
// over the scanned offsets and a direct funnel-shift extract from the slot's big-endian words, so that
// the key planes are written fully coalesced.
#include "kernels.cuh"

namespace hsk {

template <int SW, bool EXT>
__device__ __forceinline__ void xp_load4(const ExpandSegment &seg, u64 tile, int k, u32 (&n)[XP_SPT])
{
    constexpr int PW = SW - (EXT ? 2 : 0);
    const u64 base = tile * XP_TILE + (u64)threadIdx.x * XP_SPT;
#pragma unroll
    for (int i = 0; i < XP_SPT; ++i) {
        n[i] = 0;
        if (base + i < seg.nslots) {
            const u32 len = __ldg(seg.slots + (base + i) * SW + (PW - 1)) & 0xFFu;
            n[i] = len - (u32)k + 1;
        }
    }
}

template <int SW, bool EXT>
__global__ void __launch_bounds__(XP_THREADS) k_expand_tile_sums(ExpandSegment seg, int k, u32 *__restrict__ tile_sums)
{
    __shared__ u32 s_n[XP_THREADS / 32];
    u32 n[XP_SPT];
    xp_load4<SW, EXT>(seg, blockIdx.x, k, n);
    u32 a = 0;
#pragma unroll
    for (int i = 0; i < XP_SPT; ++i) a += n[i];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, d);
    if ((threadIdx.x & 31) == 0) s_n[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 ta = 0;
        for (int i = 0; i < XP_THREADS / 32; ++i) ta += s_n[i];
        tile_sums[blockIdx.x] = ta;
    }
}

// exclusive scan of the tile sums (one block, sequential chunks with carry)
__global__ void __launch_bounds__(1024) k_expand_tile_scan(const u32 *__restrict__ tile_sums, u64 ntiles, u64 *__restrict__ tile_base)
{
    __shared__ u64 s_a[32];
    __shared__ u64 carry_a;
    if (threadIdx.x == 0) carry_a = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u64 base = 0; base < ntiles; base += 1024) {
        u64 t = base + threadIdx.x;
        u64 a = t < ntiles ? tile_sums[t] : 0;
        u64 ia = a;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 x = __shfl_up_sync(0xFFFFFFFFu, ia, d);
            if (lane >= d) ia += x;
        }
        if (lane == 31) s_a[warp] = ia;
        __syncthreads();
        if (warp == 0) {
            u64 x = s_a[lane], ix = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u64 p = __shfl_up_sync(0xFFFFFFFFu, ix, d);
                if (lane >= d) ix += p;
            }
            s_a[lane] = ix - x;
        }
        __syncthreads();
        u64 ea = carry_a + s_a[warp] + ia - a;
        if (t < ntiles) tile_base[t] = ea;
        __syncthreads();
        if (threadIdx.x == 1023) carry_a = ea + a;
        __syncthreads();
    }
}

template <int NW, int SW, bool EXT>
__global__ void __launch_bounds__(XP_THREADS) k_expand(ExpandSegment seg, int k, const u64 *__restrict__ tile_base, Planes out,
                                                        u64 *__restrict__ out_val)
{
    constexpr int PW = SW - (EXT ? 2 : 0);
    __shared__ u32 s_koff[XP_TILE + 1];
    __shared__ u32 s_wn[XP_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 tile = blockIdx.x;

    u32 n[XP_SPT];
    xp_load4<SW, EXT>(seg, tile, k, n);
    u32 tn = 0;
#pragma unroll
    for (int i = 0; i < XP_SPT; ++i) tn += n[i];
    u32 in_ = tn;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 x = __shfl_up_sync(0xFFFFFFFFu, in_, d);
        if (lane >= d) in_ += x;
    }
    if (lane == 31) s_wn[warp] = in_;
    __syncthreads();
    u32 bn = 0;
    for (int i = 0; i < warp; ++i) bn += s_wn[i];
    u32 en = bn + in_ - tn;
#pragma unroll
    for (int i = 0; i < XP_SPT; ++i) { s_koff[tid * XP_SPT + i] = en; en += n[i]; }
    if (tid == XP_THREADS - 1) s_koff[XP_TILE] = en;
    __syncthreads();

    const u32 nk = s_koff[XP_TILE];
    const u64 obase = seg.out_base + tile_base[tile];
    const int kb = k - 32 * (NW - 1);   // bases in the last word

    for (u32 j = tid; j < nk; j += XP_THREADS) {
        u32 s = 0;
#pragma unroll
        for (u32 step = XP_TILE / 2; step >= 1; step >>= 1)
            if (s_koff[s + step] <= j) s += step;
        const u32 i = j - s_koff[s];
        const u32 *__restrict__ sw = seg.slots + (tile * XP_TILE + s) * SW;
        const u32 a = i >> 4, sh = 2 * (i & 15);
        u32 x[2 * NW + 1];
#pragma unroll
        for (int t = 0; t < 2 * NW + 1; ++t) x[t] = (a + t < (u32)PW) ? __ldg(sw + a + t) : 0u;
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
        if (EXT) out_val[obase + j] = ((u64)(__ldg(sw + SW - 2) + i) << 32) | (u64)__ldg(sw + SW - 1);
    }
}

cudaError_t launch_expand(const ExpandSegment &seg, int k, int nwords, bool ext, u32 *tile_sums, u64 *tile_base,
                          Planes out_keys, u64 *out_val, cudaStream_t s)
{
    if (seg.nslots == 0) return cudaSuccess;
    const u64 ntiles = (seg.nslots + XP_TILE - 1) / XP_TILE;
#define HSK_XP(NW_, SW_, EXT_)                                                                              \
    do {                                                                                                      \
        k_expand_tile_sums<SW_, EXT_><<<(unsigned)ntiles, XP_THREADS, 0, s>>>(seg, k, tile_sums);           \
        k_expand_tile_scan<<<1, 1024, 0, s>>>(tile_sums, ntiles, tile_base);                                  \
        k_expand<NW_, SW_, EXT_><<<(unsigned)ntiles, XP_THREADS, 0, s>>>(seg, k, tile_base, out_keys, out_val); \
    } while (0)
    if (nwords == 1) { if (ext) HSK_XP(1, 8, true); else HSK_XP(1, 4, false); }
    else if (nwords == 2) { if (ext) HSK_XP(2, 12, true); else HSK_XP(2, 8, false); }
    else { if (ext) HSK_XP(3, 12, true); else HSK_XP(3, 8, false); }
#undef HSK_XP
    return cudaGetLastError();
}

} // namespace hsk
