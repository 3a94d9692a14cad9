// Stage 5a: shared-memory-staged LSD radix sort of k-mer words (+ optional 64-bit payload).
//
// Replaces the reference's HOT LOOP D: sort_task -> raduls::RadixSortMSD / paradis::sort
// (src/kmerops.cpp:1382-1407, dependency/Raduls/raduls.h:1061-1107, dependency/Paradis/
// paradissort.hpp:211-215), which are CPU MSD sorts with software write-combining.  Here:
//   * one histogram kernel reads the keys once and produces the digit histograms of ALL passes
//   * one kernel per 8-bit digit ("onesweep"): each CTA takes a tile of 6144 keys, ranks them per
//     warp with match.any, publishes its per-digit counts and resolves its global offsets by
//     decoupled look-back over the preceding tiles (no separate scan kernel, one read + one write
//     of the keys per pass), stages the tile in shared memory in sorted order and writes each digit
//     run coalesced
//   * keys are held as planes of 64-bit words (plane 0 most significant); digits never straddle a
//     word and the unused low bits of the last word are skipped (62 significant bits for K=31 ->
//     8 passes, 110 bits for K=55 -> 14 passes).
#include "kernels.cuh"

namespace hsk {

constexpr u32 LB_SHIFT = 29;
constexpr u32 LB_MASK = (1u << LB_SHIFT) - 1;

struct RadixSmem {
    u64 stage[RS_TILE];
    u32 wc[RS_WARPS][256];
    u32 dexcl[256];
    u32 gbase[256];
    u32 warp_tot[8];
    u32 tile;
    u8 dig[RS_TILE];
};

__device__ __forceinline__ u32 ld_volatile(const u32 *p) { return *reinterpret_cast<const volatile u32 *>(p); }
__device__ __forceinline__ void st_volatile(u32 *p, u32 v) { *reinterpret_cast<volatile u32 *>(p) = v; }

// ---- digit histograms of all passes in one read of the keys -----------------------------------------
template <int NW>
__global__ void __launch_bounds__(512) k_radix_hist(Planes in, u32 n, int k, u32 *__restrict__ bins)
{
    extern __shared__ u32 h[];   // npasses * 256
    const int kb = k - 32 * (NW - 1);
    const int npasses = (2 * kb + 7) / 8 + 8 * (NW - 1);
    for (int i = threadIdx.x; i < npasses * 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        u64 w[NW];
#pragma unroll
        for (int l = 0; l < NW; ++l) w[l] = in.p[l][i];
        int p = 0;
#pragma unroll
        for (int pl = NW - 1; pl >= 0; --pl) {
            const int low = (pl == NW - 1) ? 64 - 2 * kb : 0;
            for (int s = low; s < 64; s += 8, ++p) atomicAdd(&h[p * 256 + (int)((w[pl] >> s) & 255)], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npasses * 256; i += blockDim.x)
        if (h[i]) atomicAdd(&bins[i], h[i]);
}

// exclusive scan of every pass's 256 bins (one block of 256 threads per pass)
__global__ void __launch_bounds__(256) k_radix_scan_bins(u32 *__restrict__ bins)
{
    __shared__ u32 wt[8];
    u32 *b = bins + blockIdx.x * 256;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 v = b[threadIdx.x], inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) wt[warp] = inc;
    __syncthreads();
    u32 off = 0;
    for (int i = 0; i < warp; ++i) off += wt[i];
    b[threadIdx.x] = off + inc - v;
}

// ---- one LSD pass ------------------------------------------------------------------------------------
template <int NW, bool HAS_VAL>
__global__ void __launch_bounds__(RS_THREADS, 2) k_onesweep(Planes in, Planes out, const u64 *__restrict__ vin,
                                                             u64 *__restrict__ vout, u32 n, int plane, int shift,
                                                             const u32 *__restrict__ bins, u32 *lookback,
                                                             u32 *tile_counter, u32 flag_base)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    RadixSmem &sm = *reinterpret_cast<RadixSmem *>(smraw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool MULTI = (NW > 1) || HAS_VAL;
    const u32 FLAG_P = flag_base + 1, FLAG_I = flag_base + 2;

    if (tid == 0) sm.tile = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < RS_WARPS * 256; i += RS_THREADS) (&sm.wc[0][0])[i] = 0;
    __syncthreads();
    const u32 tile = sm.tile;
    const u32 tbase = tile * RS_TILE;
    const u32 valid = min((u32)RS_TILE, n - tbase);
    const u32 wbase = tbase + warp * (32 * RS_IPT) + lane;

    // load the digit plane, warp-striped (every load instruction of a warp covers 256 contiguous bytes)
    const u64 *__restrict__ kin = in.p[plane];
    u64 key[RS_IPT];
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
        u32 idx = wbase + j * 32;
        key[j] = idx < n ? kin[idx] : ~0ull;
    }

    // rank inside the warp: lanes holding the same digit form a group, the lowest lane claims
    // the group's slots from the warp's digit counter
    u32 lp[RS_IPT];
    const u32 lt = (1u << lane) - 1;
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
        u32 d = (u32)(key[j] >> shift) & 255u;
        u32 peers = __match_any_sync(0xFFFFFFFFu, d);
        int leader = __ffs(peers) - 1;
        u32 pre = 0;
        if (lane == leader) pre = atomicAdd(&sm.wc[warp][d], (u32)__popc(peers));
        pre = __shfl_sync(0xFFFFFFFFu, pre, leader);
        lp[j] = pre + __popc(peers & lt);
    }
    __syncthreads();

    // per digit: exclusive prefix over warps (in place), tile total, publish, scan over digits
    u32 total = 0, total_real = 0, inc = 0;
    if (tid < 256) {
        u32 sum = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { u32 c = sm.wc[w][tid]; sm.wc[w][tid] = sum; sum += c; }
        total = sum;
        total_real = (tid == 255) ? total - (RS_TILE - valid) : total;   // padding keys rank last in digit 255
        st_volatile(&lookback[tile * 256 + tid], ((tile == 0 ? FLAG_I : FLAG_P) << LB_SHIFT) | total_real);
        inc = total;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) sm.warp_tot[warp] = inc;
    }
    __syncthreads();
    if (tid < 256) {
        u32 off = 0;
        for (int i = 0; i < warp; ++i) off += sm.warp_tot[i];
        sm.dexcl[tid] = off + inc - total;
    }
    __syncthreads();

    // local sorted position of every key; stage the tile in sorted order
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
        u32 d = (u32)(key[j] >> shift) & 255u;
        lp[j] += sm.dexcl[d] + sm.wc[warp][d];
        sm.stage[lp[j]] = key[j];
    }

    // decoupled look-back: sum of this digit's counts over all preceding tiles
    if (tid < 256) {
        u32 prev = 0;
        if (tile > 0) {
            int t = (int)tile - 1;
            while (true) {
                u32 e = ld_volatile(&lookback[(u32)t * 256 + tid]);
                u32 f = e >> LB_SHIFT;
                if (f == FLAG_I) { prev += e & LB_MASK; break; }
                if (f == FLAG_P) { prev += e & LB_MASK; --t; }
            }
            st_volatile(&lookback[tile * 256 + tid], (FLAG_I << LB_SHIFT) | (prev + total_real));
        }
        sm.gbase[tid] = __ldg(bins + tid) + prev - sm.dexcl[tid];
    }
    __syncthreads();

    // write the digit plane: consecutive threads write consecutive addresses inside each digit run
    {
        u64 *__restrict__ kout = out.p[plane];
        for (u32 i = tid; i < valid; i += RS_THREADS) {
            u64 kk = sm.stage[i];
            u32 d = (u32)(kk >> shift) & 255u;
            if (MULTI) sm.dig[i] = (u8)d;
            kout[sm.gbase[d] + i] = kk;
        }
    }

    // the other planes follow the same permutation
    if (MULTI) {
#pragma unroll
        for (int q = 0; q < NW + (HAS_VAL ? 1 : 0); ++q) {
            if (q == plane) continue;
            const u64 *__restrict__ src = (q < NW) ? in.p[q] : vin;
            u64 *__restrict__ dst = (q < NW) ? out.p[q] : vout;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < RS_IPT; ++j) {
                u32 idx = wbase + j * 32;
                if (idx < n) sm.stage[lp[j]] = src[idx];
            }
            __syncthreads();
            for (u32 i = tid; i < valid; i += RS_THREADS) dst[sm.gbase[sm.dig[i]] + i] = sm.stage[i];
        }
    }
}

size_t radix_scratch_bytes(u64 n)
{
    u64 ntiles = (n + RS_TILE - 1) / RS_TILE;
    return (size_t)(RS_MAX_PASSES * 256 + 32 + ntiles * 256) * sizeof(u32);
}

template <int NW, bool HAS_VAL>
static cudaError_t run_passes(Planes a, Planes b, u64 *va, u64 *vb, u32 n, int k, u32 *bins, u32 *counters, u32 *lookback,
                              const PassTable &pt, bool *result_in_b, cudaStream_t s, cudaEvent_t ev0, cudaEvent_t ev1)
{
    const u32 ntiles = (n + RS_TILE - 1) / RS_TILE;
    const size_t smem = sizeof(RadixSmem);
    cudaError_t e = cudaFuncSetAttribute(k_onesweep<NW, HAS_VAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_radix_hist<NW><<<148 * 4, 512, pt.npasses * 256 * sizeof(u32), s>>>(a, n, k, bins);
    k_radix_scan_bins<<<pt.npasses, 256, 0, s>>>(bins);
    bool in_b = false;
    if (ev0) cudaEventRecord(ev0, s);
    for (int p = 0; p < pt.npasses; ++p) {
        Planes &src = in_b ? b : a;
        Planes &dst = in_b ? a : b;
        k_onesweep<NW, HAS_VAL><<<ntiles, RS_THREADS, smem, s>>>(src, dst, in_b ? vb : va, in_b ? va : vb, n, pt.d[p].plane,
                                                                 pt.d[p].shift, bins + p * 256, lookback, counters + p,
                                                                 (p & 1) ? 2u : 0u);
        in_b = !in_b;
    }
    if (ev1) cudaEventRecord(ev1, s);
    *result_in_b = in_b;
    return cudaGetLastError();
}

cudaError_t launch_radix_sort(Planes a, Planes b, u64 *va, u64 *vb, u64 n, int nwords, int k, void *scratch,
                              bool *result_in_b, int *npasses, int *nlaunches, cudaStream_t s, cudaEvent_t ev0,
                              cudaEvent_t ev1)
{
    *result_in_b = false;
    PassTable pt = make_pass_table(k);
    if (npasses) *npasses = pt.npasses;
    if (nlaunches) *nlaunches = 0;
    if (n == 0) return cudaSuccess;
    if (n > LB_MASK) return cudaErrorInvalidValue;
    u32 *bins = reinterpret_cast<u32 *>(scratch);
    u32 *counters = bins + RS_MAX_PASSES * 256;
    u32 *lookback = counters + 32;
    cudaError_t e = cudaMemsetAsync(scratch, 0, radix_scratch_bytes(n), s);
    if (e != cudaSuccess) return e;
    if (nlaunches) *nlaunches = 2 + pt.npasses;
    const bool hv = va != nullptr;
    if (nwords == 1) return hv ? run_passes<1, true>(a, b, va, vb, (u32)n, k, bins, counters, lookback, pt, result_in_b, s, ev0, ev1)
                              : run_passes<1, false>(a, b, va, vb, (u32)n, k, bins, counters, lookback, pt, result_in_b, s, ev0, ev1);
    if (nwords == 2) return hv ? run_passes<2, true>(a, b, va, vb, (u32)n, k, bins, counters, lookback, pt, result_in_b, s, ev0, ev1)
                              : run_passes<2, false>(a, b, va, vb, (u32)n, k, bins, counters, lookback, pt, result_in_b, s, ev0, ev1);
    return hv ? run_passes<3, true>(a, b, va, vb, (u32)n, k, bins, counters, lookback, pt, result_in_b, s, ev0, ev1)
              : run_passes<3, false>(a, b, va, vb, (u32)n, k, bins, counters, lookback, pt, result_in_b, s, ev0, ev1);
}

} // namespace hsk
