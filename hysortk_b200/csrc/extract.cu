// Stage 1+2: minimizer scan, supermer extraction and owner-hash bucketing, straight from the
// 2-bit packed DnaBuffer.
//
// Replaces the reference's HOT LOOPS A+B: FindKmerDestinationsParallel (src/kmerops.cpp:1010-1041,
// canonical m-mers supermer.hpp:315-342, sliding-window minimum :1058-1073, owner :1044-1047) and
// SupermerEncoder::encode / copy_bits (:1096-1148), plus the per-thread ScatteredSupermers staging
// (:253-358).  Not a translation: the reference walks each read with a deque and materialises one
// int per k-mer; here the whole packed buffer is treated as one flat sequence of 2-bit slots, cut
// into warp tiles that persistent warps process on their own (no block-wide barrier anywhere):
//
//   k_tile_reads: the read holding the first byte of every tile (one binary search per tile)
//   pass A (k_supermer_count<W>), per warp tile, every lane owning 16 consecutive slots
//     1. one coalesced 32-bit load per lane brings the tile's bases; neighbours' words come by shuffle
//     2. the lane rolls forward / reverse m-mers over its 16 slots and hashes the canonical one: 16 hashes
//        in registers
//     3. minimizer of every k-mer slot = minimum over its W = K-M+1 hashes: suffix minima of the own
//        hashes, prefix minima of the following lanes fetched by shuffle (W is a template parameter, so
//        every index is static and nothing leaves the register file)
//     4. state of a slot = its minimizer hash, or 0 when no k-mer starts there (read table lookup);
//        run boundaries = state changes -> per-lane bit mask -> warp scan -> compacted boundary list in the
//        warp's shared memory
//     5. one lane per run: per-bin totals (one 64-bit reduction per run) and the run list entry
//   k_bin_scan: exclusive prefix of the per-bin totals -> bin starts / cursors in the supermer stream
//   pass B (k_supermer_scatter), per warp tile: one lane per run claims the run's slot(s) in its bin with
//     one atomic and writes the re-packed bases, length and optional (pos, rid).  No hashing is repeated.
//
// A supermer is a run of consecutive k-mers of one read with the same minimizer, stored in fixed-size
// slots (common.cuh: 16 bytes for K <= 32: 60 bases + length; longer runs are split into overlapping
// pieces), so that the scatter is one atomic + one 128-bit store per supermer.  Bins are fine-grained (a
// few thousand k-mers each) so that a bin can later be expanded, sorted and counted entirely inside one
// CTA's shared memory.  Where the reference splits supermers (250-base cap, kmerops.cpp:1120) and how it
// hashes are free choices: only the multiset of k-mers per bin matters, and a canonical k-mer always
// lands in the same bin because the bin is a function of its set of canonical m-mers.
#include "kernels.cuh"

#include <utility>

namespace hsk {

constexpr u32 FULL = 0xFFFFFFFFu;

// largest r in [lo, hi] with off[r] <= byte (off is non-decreasing; caller guarantees off[lo] <= byte)
__device__ __forceinline__ u64 find_read(const u64 *__restrict__ off, u64 lo, u64 hi, u64 byte)
{
    while (lo < hi) {
        u64 mid = lo + (hi - lo + 1) / 2;
        if (__ldg(off + mid) <= byte) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_tile_reads(ExtractParams P, u32 *__restrict__ tile_read)
{
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > P.ntiles) return;
    const u64 byte = min(t * (u64)(P.out_slots / 4), P.nbytes);
    tile_read[t] = P.nreads ? (u32)find_read(P.read_off, 0, P.nreads - 1, byte) : 0u;
}

// ---- minimizer windows in registers ------------------------------------------------------------------
// h[0..15] = hashes of the lane's own slots.  Window of slot i = hashes i .. i+W-1 of the concatenation of
// the lanes' chunks = suffix of the own chunk, whole chunks of the next lanes, prefix of a later lane.
template <int W, int I>
__device__ __forceinline__ u32 window_at(const u32 (&h)[XT_R], const u32 (&suf)[XT_R], const u32 (&pre)[XT_R])
{
    constexpr int END = I + W - 1;
    if constexpr (END < XT_R) {
        if constexpr (I == 0) return pre[END];
        else if constexpr (END == XT_R - 1) return suf[I];
        else {
            u32 r = h[I];
#pragma unroll
            for (int q = I + 1; q <= END; ++q) r = min(r, h[q]);
            return r;
        }
    } else {
        constexpr int C = END >> 4, J = END & 15;
        u32 r = min(suf[I], __shfl_down_sync(FULL, pre[J], C));
#pragma unroll
        for (int t = 1; t < C; ++t) r = min(r, __shfl_down_sync(FULL, pre[XT_R - 1], t));
        return r;
    }
}

template <int W, int... I>
__device__ __forceinline__ void window_all(const u32 (&h)[XT_R], const u32 (&suf)[XT_R], const u32 (&pre)[XT_R], u32 (&out)[XT_R],
                                           std::integer_sequence<int, I...>)
{
    ((out[I] = window_at<W, I>(h, suf, pre)), ...);
}

template <int W>
__device__ __forceinline__ void window_min(const u32 (&h)[XT_R], u32 (&out)[XT_R])
{
    u32 suf[XT_R], pre[XT_R];
    pre[0] = h[0];
#pragma unroll
    for (int i = 1; i < XT_R; ++i) pre[i] = min(pre[i - 1], h[i]);
    suf[XT_R - 1] = h[XT_R - 1];
#pragma unroll
    for (int i = XT_R - 2; i >= 0; --i) suf[i] = min(suf[i + 1], h[i]);
    window_all<W>(h, suf, pre, out, std::make_integer_sequence<int, XT_R>{});
}

// big-endian words (16 bases each, first base in the top bits) w0, w1, w2 of a lane: word lane, lane+1, lane+2
// of the 34 words starting at word wi0 of the packed buffer
__device__ __forceinline__ void load_lane_words(const u32 *__restrict__ packed32, u64 wi0, u64 nwords_readable, int lane,
                                                u32 &w0, u32 &w1, u32 &w2)
{
    u32 wa = 0, wb = 0;
    if (wi0 + lane < nwords_readable) wa = __byte_perm(__ldg(packed32 + wi0 + lane), 0, 0x0123);
    if (lane < 2 && wi0 + 32 + lane < nwords_readable) wb = __byte_perm(__ldg(packed32 + wi0 + 32 + lane), 0, 0x0123);
    const u32 a1 = __shfl_sync(FULL, wa, (lane + 1) & 31), b1 = __shfl_sync(FULL, wb, 0);
    const u32 a2 = __shfl_sync(FULL, wa, (lane + 2) & 31), b2 = __shfl_sync(FULL, wb, (lane + 2) & 31);
    w0 = wa;
    w1 = lane == 31 ? b1 : a1;
    w2 = lane >= 30 ? b2 : a2;
}

// ---- pass A: per-bin totals + the run list of every tile ------------------------------------------
// bin_tot[b] += bt_pack(slots, k-mers) per valid run.  Run list entry = (n << 48 | start << 32 | bin);
// tile_hdr[tile] = (first entry, number of entries).
template <int W>
__global__ void __launch_bounds__(XT_THREADS, XT_CTAS_PER_SM) k_supermer_count(ExtractParams P, u64 *__restrict__ bin_tot,
                                                                u64 *__restrict__ run_list,
                                                                ulonglong2 *__restrict__ tile_hdr,
                                                                u64 *__restrict__ run_cursor, u64 run_capacity)
{
    // run_cursor[0] = entries of the run list; run_cursor[2], [3] = grand totals of slots / k-mers, kept apart from the
    // per-bin words so that a field overflow there is noticed (common.cuh: bt_pack)
    constexpr int OL = xt_out_lanes(W), OUT = OL * XT_R;
    u64 chk_slots = 0, chk_kmers = 0;
    __shared__ u32 s_st[XT_WARPS][OUT + 2];
    __shared__ u16 s_pos[XT_WARPS][OUT + 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 gw = (u64)blockIdx.x * XT_WARPS + warp;
    const u64 t0 = P.tile_begin + gw * P.tiles_per_warp, t1 = min(t0 + P.tiles_per_warp, P.tile_end);
    const u32 *packed32 = reinterpret_cast<const u32 *>(P.packed);
    const u64 nwords_readable = P.nbytes_padded >> 2;
    const int m = P.m;
    const u64 mask = (m == 32) ? ~0ull : ((1ull << (2 * m)) - 1);
    const int rcs = 2 * (m - 1);
    u32 *my_st = s_st[warp];
    u16 *my_pos = s_pos[warp];

    for (u64 tile = t0; tile < t1; ++tile) {
        // ---- 1. bases of my 16 slots + the next 32 (m-mers reach up to 31 bases further)
        u32 w0, w1, w2;
        load_lane_words(packed32, tile * OL, nwords_readable, lane, w0, w1, w2);

        // ---- 2. rolling canonical m-mer hashes
        u32 h[XT_R];
        {
            const u64 hi = ((u64)w0 << 32) | w1;
            u64 fwd = hi >> (64 - 2 * m);
            u64 rc = revcomp64(hi) & mask;
            // the 15 bases after the first m-mer, left-aligned
            const u32 nxt = (m == 32) ? w2 : (u32)(((hi << (2 * m)) | (((u64)w2 << 32) >> (64 - 2 * m))) >> 32);
            h[0] = mmer_hash(fwd < rc ? fwd : rc);
#pragma unroll
            for (int j = 1; j < XT_R; ++j) {
                const u64 c = (nxt >> (32 - 2 * j)) & 3u;
                fwd = ((fwd << 2) | c) & mask;
                rc = (rc >> 2) | ((3 - c) << rcs);
                h[j] = mmer_hash(fwd < rc ? fwd : rc);
            }
        }

        // ---- 3. minimizer hash of every slot
        u32 mn[XT_R];
        window_min<W>(h, mn);

        // ---- 4. valid k-mer starts among my slots -> states -> run boundaries
        u32 vmask = 0;
        const u64 p0 = tile * OUT + (u64)lane * XT_R;
        if (lane < OL && P.nreads > 0 && (p0 >> 2) < P.nbytes) {
            u64 r = find_read(P.read_off, __ldg(P.tile_read + tile), __ldg(P.tile_read + tile + 1), p0 >> 2);
            u64 rstart = __ldg(P.read_off + r) * 4;
            u64 rnext = __ldg(P.read_off + r + 1) * 4;
            u64 rend = rstart + __ldg(P.read_len + r);
            if (p0 + XT_R + (u64)P.k <= rend + 1 && p0 >= rstart) {
                vmask = 0xFFFFu;   // common case: all 16 windows inside the read
            } else {
#pragma unroll
                for (int j = 0; j < XT_R; ++j) {
                    const u64 p = p0 + j;
                    while (p >= rnext && r + 1 < P.nreads) {
                        ++r;
                        rstart = rnext;
                        rnext = __ldg(P.read_off + r + 1) * 4;
                        rend = rstart + __ldg(P.read_len + r);
                    }
                    if (p >= rstart && p + (u64)P.k <= rend) vmask |= 1u << j;
                }
            }
        }
        // state: minimizer hash with the low bit forced to 1, or 0 when no k-mer starts at the slot
        u32 st[XT_R];
#pragma unroll
        for (int i = 0; i < XT_R; ++i) st[i] = ((vmask >> i) & 1) ? (mn[i] | 1u) : 0u;
        u32 bnd = 0, nz = 0;
        {
            const u32 prev = __shfl_up_sync(FULL, st[XT_R - 1], 1);
            if (lane == 0 || st[0] != prev) bnd |= 1u;
#pragma unroll
            for (int i = 1; i < XT_R; ++i) if (st[i] != st[i - 1]) bnd |= 1u << i;
#pragma unroll
            for (int i = 0; i < XT_R; ++i) if (st[i] != 0) nz |= 1u << i;
            if (lane >= OL) bnd = 0;
        }
        const u32 cnt = __popc(bnd);
        u32 inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 t = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += t;
        }
        const u32 nent = __shfl_sync(FULL, inc, 31);               // boundaries in the tile
        const u32 nv = __reduce_add_sync(FULL, __popc(bnd & nz));  // runs that hold k-mers
        {
            u32 off = inc - cnt;
#pragma unroll
            for (int i = 0; i < XT_R; ++i) {
                if ((bnd >> i) & 1) { my_st[off] = st[i]; my_pos[off] = (u16)(lane * XT_R + i); ++off; }
            }
            if (lane == 0) my_pos[nent] = (u16)OUT;
        }
        u64 rb = 0;
        if (lane == 0) {
            if (nv) rb = atomicAdd(run_cursor, (u64)nv);
            tile_hdr[tile] = make_ulonglong2(rb, (u64)nv);
        }
        rb = __shfl_sync(FULL, rb, 0);
        __syncwarp();
        const bool fits = (rb + nv <= run_capacity);   // otherwise the host sees run_cursor > capacity and retries

        // ---- 5. one lane per run
        u32 base = 0;
        for (u32 j0 = 0; j0 < nent; j0 += 32) {
            const u32 j = j0 + lane;
            u32 s = 0, start = 0, n = 0;
            if (j < nent) { s = my_st[j]; start = my_pos[j]; n = my_pos[j + 1] - start; }
            const bool valid = (s != 0);
            const u32 bal = __ballot_sync(FULL, valid);
            if (valid) {
                const u32 b = hash_bucket(s, P.nbins);
                const u32 pieces = __umulhi(n + P.slot_nmax - 1, P.slot_ninv);
                atomicAdd(&bin_tot[b], bt_pack(pieces, n));
                chk_slots += pieces; chk_kmers += n;
                const u32 idx = base + __popc(bal & ((1u << lane) - 1));
                if (fits) run_list[rb + idx] = ((u64)(start | (n << 16)) << 32) | b;
            }
            base += __popc(bal);
        }
        __syncwarp();
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        chk_slots += __shfl_xor_sync(FULL, chk_slots, d);
        chk_kmers += __shfl_xor_sync(FULL, chk_kmers, d);
    }
    if (lane == 0 && chk_kmers) { atomicAdd(run_cursor + 2, chk_slots); atomicAdd(run_cursor + 3, chk_kmers); }
}

// ---- bin scan: exclusive prefix over bins of the slot counts -> bin starts / cursors; k-mer total ---------
// Three small launches (tile sums, scan of the tile sums, per-tile scan with its base): the number of bins grows with
// the number of ranks (150 K at 8 GPUs), so one block must not walk them all.
constexpr int BS_THREADS = 1024, BS_PER = 4, BS_TILE = BS_THREADS * BS_PER;

__device__ __forceinline__ u64 block_sum_1024(u64 v, u64 *s_c)
{
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = v;
    __syncthreads();
    u64 t = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) t += s_c[w];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(BS_THREADS) k_bin_tile_sums(const u64 *__restrict__ bin_tot, u32 nbins, u64 *__restrict__ tile_sums,
                                                               u64 *__restrict__ kmers_total)
{
    __shared__ u64 s_c[32];
    u64 slots = 0, kmers = 0;
#pragma unroll
    for (int i = 0; i < BS_PER; ++i) {
        const u32 b = blockIdx.x * BS_TILE + threadIdx.x * BS_PER + i;
        const u64 v = b < nbins ? bin_tot[b] : 0;
        slots += bt_slots(v);
        kmers += bt_kmers(v);
    }
    slots = block_sum_1024(slots, s_c);
    kmers = block_sum_1024(kmers, s_c);
    if (threadIdx.x == 0) {
        tile_sums[blockIdx.x] = slots;
        if (kmers) atomicAdd(kmers_total, kmers);
    }
}

// one block: exclusive scan of the tile sums in place; bin_start[nbins] = total
__global__ void __launch_bounds__(BS_THREADS) k_bin_tile_scan(u64 *__restrict__ tile_sums, u32 ntiles, u64 *__restrict__ total_out)
{
    __shared__ u64 s_c[32];
    __shared__ u64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u32 base = 0; base < ntiles; base += BS_THREADS) {
        const u32 t = base + threadIdx.x;
        const u64 v = t < ntiles ? tile_sums[t] : 0;
        u64 inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u64 x = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += x;
        }
        if (lane == 31) s_c[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const u64 x = s_c[lane];
            u64 ix = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u64 y = __shfl_up_sync(FULL, ix, d);
                if (lane >= d) ix += y;
            }
            s_c[lane] = ix - x;
        }
        __syncthreads();
        const u64 ex = carry + s_c[warp] + inc - v;
        if (t < ntiles) tile_sums[t] = ex;
        __syncthreads();
        if (threadIdx.x == BS_THREADS - 1) carry = ex + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(BS_THREADS) k_bin_starts(const u64 *__restrict__ bin_tot, u32 nbins, const u64 *__restrict__ tile_base,
                                                            u64 *__restrict__ bin_start, u64 *__restrict__ bin_cursor)
{
    __shared__ u64 s_c[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 c[BS_PER], tc = 0;
#pragma unroll
    for (int i = 0; i < BS_PER; ++i) {
        const u32 b = blockIdx.x * BS_TILE + threadIdx.x * BS_PER + i;
        c[i] = b < nbins ? bt_slots(bin_tot[b]) : 0;
        tc += c[i];
    }
    u64 ic = tc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u64 a = __shfl_up_sync(FULL, ic, d);
        if (lane >= d) ic += a;
    }
    if (lane == 31) s_c[warp] = ic;
    __syncthreads();
    if (warp == 0) {
        const u64 a = s_c[lane];
        u64 ia = a;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u64 x = __shfl_up_sync(FULL, ia, d);
            if (lane >= d) ia += x;
        }
        s_c[lane] = ia - a;
    }
    __syncthreads();
    u64 ec = tile_base[blockIdx.x] + s_c[warp] + ic - tc;
#pragma unroll
    for (int i = 0; i < BS_PER; ++i) {
        const u32 b = blockIdx.x * BS_TILE + threadIdx.x * BS_PER + i;
        if (b < nbins) { bin_start[b] = ec; bin_cursor[b] = ec; }
        ec += c[i];
    }
}

// ---- pass B: one slot per piece of every valid run ----------------------------------------------------
template <int SW, bool EXT>
__global__ void __launch_bounds__(XT_THREADS) k_supermer_scatter(ExtractParams P, const u64 *__restrict__ run_list,
                                                                  const ulonglong2 *__restrict__ tile_hdr,
                                                                  u64 *__restrict__ bin_cursor)
{
    constexpr int PW = SW - (EXT ? 2 : 0);
    __shared__ u32 s_w[XT_WARPS][XT_STAGE_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 gw = (u64)blockIdx.x * XT_WARPS + warp;
    const u64 t0 = P.tile_begin + gw * P.tiles_per_warp, t1 = min(t0 + P.tiles_per_warp, P.tile_end);
    const u32 *packed32 = reinterpret_cast<const u32 *>(P.packed);
    const u64 nwords_readable = P.nbytes_padded >> 2;
    const u32 OL = P.out_slots / XT_R;
    u32 *wbe = s_w[warp];

    for (u64 tile = t0; tile < t1; ++tile) {
        const ulonglong2 hdr = __ldg(tile_hdr + tile);
        const u32 nv = (u32)hdr.y;
        if (nv == 0) continue;
        const u64 wi0 = tile * OL;
        {
            u32 wa = 0, wb = 0;
            if (wi0 + lane < nwords_readable) wa = __byte_perm(__ldg(packed32 + wi0 + lane), 0, 0x0123);
            if (lane < XT_STAGE_WORDS - 32 && wi0 + 32 + lane < nwords_readable) wb = __byte_perm(__ldg(packed32 + wi0 + 32 + lane), 0, 0x0123);
            wbe[lane] = wa;
            if (lane < XT_STAGE_WORDS - 32) wbe[32 + lane] = wb;
        }
        u64 rlo = 0, rhi = 0;
        if (EXT) { rlo = __ldg(P.tile_read + tile); rhi = __ldg(P.tile_read + tile + 1); }
        __syncwarp();
        const u64 slot0 = tile * (u64)P.out_slots;
        for (u32 j = lane; j < nv; j += 32) {
            const u64 e = __ldg(run_list + hdr.x + j);
            const u32 b = (u32)e;
            const u32 start = (u32)(e >> 32) & 0xFFFFu, n = (u32)(e >> 48);
            const u32 pieces = __umulhi(n + P.slot_nmax - 1, P.slot_ninv);
            u64 gs = atomicAdd(&bin_cursor[b], (u64)pieces);
            u32 pos0 = 0, rid = 0;
            if (EXT) {
                const u64 p = slot0 + start;
                const u64 r = find_read(P.read_off, rlo, rhi, p >> 2);
                pos0 = (u32)(p - __ldg(P.read_off + r) * 4);
                rid = (u32)((long long)r + (long long)P.readid_base);
            }
            for (u32 pc = 0; pc < pieces; ++pc, ++gs) {
                const u32 ps = start + pc * P.slot_nmax;
                const u32 np = min(P.slot_nmax, n - pc * P.slot_nmax);
                const u32 len = np + P.k - 1;
                u32 w[SW];
#pragma unroll
                for (int x = 0; x < PW; ++x) {
                    const u32 q = ps + 16 * x;
                    const int rem = (int)len - 16 * x;
                    u32 wv = 0;
                    if (rem > 0) {
                        wv = __funnelshift_l(wbe[(q >> 4) + 1], wbe[q >> 4], 2 * (q & 15));
                        if (rem < 16) wv &= ~0u << (32 - 2 * rem);
                    }
                    w[x] = wv;
                }
                w[PW - 1] &= 0xFFFFFF00u;
                if (SW == 4 && !EXT) {
                    // Orientation: of the supermer and its reverse complement keep the smaller string.  Both hold the
                    // same canonical k-mers, and copies of a locus read from either strand become bit-identical slots,
                    // which the bin kernel counts once with a weight (bins.cu: dedup_bin).
                    const u64 fh = ((u64)w[0] << 32) | w[1], fl = ((u64)w[2] << 32) | w[3];
                    u64 rh = revcomp64(fl), rl = revcomp64(fh);   // all 64 positions reversed: the string is now at the end
                    const u32 sh = 2 * (64 - len);                // >= 8: a slot holds at most 60 bases
                    if (sh >= 64) { rh = rl << (sh - 64); rl = 0; }
                    else { rh = (rh << sh) | (rl >> (64 - sh)); rl <<= sh; }
                    if (rh < fh || (rh == fh && rl < fl)) {
                        w[0] = (u32)(rh >> 32); w[1] = (u32)rh; w[2] = (u32)(rl >> 32); w[3] = (u32)rl;
                    }
                } else if (!EXT) {
                    // the same for the 32-byte slots of K in 33..64  Both hold the
                    // same canonical k-mers, and copies of a locus read from either strand become bit-identical slots,
                    // which the bin kernel counts once with a weight (bins.cu: dedup_bin).
                    constexpr int NQ = SW / 2;   // 64-bit words of the slot
                    u64 f[NQ], r[NQ];
#pragma unroll
                    for (int x = 0; x < NQ; ++x) f[x] = ((u64)w[2 * x] << 32) | w[2 * x + 1];
                    // all 32 NQ positions reversed and complemented: the string is now at the end; shift it to the front
#pragma unroll
                    for (int x = 0; x < NQ; ++x) r[x] = revcomp64(f[NQ - 1 - x]);
                    const u32 sh = 2 * (32 * NQ - len);   // >= 8: the last 4 positions of a slot are never bases
                    const u32 ws = sh >> 6, bs = sh & 63;
#pragma unroll
                    for (int x = 0; x < NQ; ++x) {
                        // word x of (r << sh) = r[x + ws] << bs | r[x + ws + 1] >> (64 - bs)
                        u64 a = 0, b = 0;
#pragma unroll
                        for (int y = 0; y < NQ; ++y) {
                            if ((u32)y == x + ws) a = r[y];
                            if ((u32)y == x + ws + 1) b = r[y];
                        }
                        f[x] = bs ? ((a << bs) | (b >> (64 - bs))) : a;   // f is re-used below: keep the original in w
                    }
                    bool less = false, decided = false;
#pragma unroll
                    for (int x = 0; x < NQ; ++x) {
                        const u64 o = ((u64)w[2 * x] << 32) | w[2 * x + 1];
                        if (!decided && f[x] != o) { less = f[x] < o; decided = true; }
                    }
                    if (less) {
#pragma unroll
                        for (int x = 0; x < NQ; ++x) { w[2 * x] = (u32)(f[x] >> 32); w[2 * x + 1] = (u32)f[x]; }
                    }
                }
                w[PW - 1] |= len;
                if (EXT) { w[SW - 2] = pos0 + pc * P.slot_nmax; w[SW - 1] = rid; }
                uint4 *dst = reinterpret_cast<uint4 *>(P.out_stream + gs * SW);
#pragma unroll
                for (int x = 0; x < SW / 4; ++x) dst[x] = make_uint4(w[4 * x], w[4 * x + 1], w[4 * x + 2], w[4 * x + 3]);
            }
        }
        __syncwarp();
    }
}

// ---- launchers ---------------------------------------------------------------------------------------
template <int W>
static cudaError_t launch_count_w(const ExtractParams &P, u32 nctas, u64 *bin_tot, u64 *run_list, ulonglong2 *tile_hdr,
                                  u64 *run_cursor, u64 run_capacity, cudaStream_t s)
{
    k_supermer_count<W><<<nctas, XT_THREADS, 0, s>>>(P, bin_tot, run_list, tile_hdr, run_cursor, run_capacity);
    return cudaGetLastError();
}

template <int... W>
static cudaError_t dispatch_count(int w, const ExtractParams &P, u32 nctas, u64 *bin_tot, u64 *run_list, ulonglong2 *tile_hdr,
                                  u64 *run_cursor, u64 run_capacity, cudaStream_t s, std::integer_sequence<int, W...>)
{
    cudaError_t e = cudaErrorInvalidValue;
    ((w == W + 1 ? (void)(e = launch_count_w<W + 1>(P, nctas, bin_tot, run_list, tile_hdr, run_cursor, run_capacity, s)) : (void)0), ...);
    return e;
}

u32 extract_grid(int w, int sm_count)
{
    (void)w;
    return (u32)sm_count * (u32)XT_CTAS_PER_SM;   // every CTA resident; tiles are spread evenly over the warps
}

cudaError_t launch_tile_reads(const ExtractParams &P, u32 *tile_read, cudaStream_t s)
{
    k_tile_reads<<<(unsigned)((P.ntiles + 1 + 255) / 256), 256, 0, s>>>(P, tile_read);
    return cudaGetLastError();
}

cudaError_t launch_supermer_count(const ExtractParams &P, u32 nctas, u64 *bin_tot, u64 *run_list, ulonglong2 *tile_hdr,
                                  u64 *run_cursor, u64 run_capacity, cudaStream_t s)
{
    return dispatch_count(P.k - P.m + 1, P, nctas, bin_tot, run_list, tile_hdr, run_cursor, run_capacity, s,
                          std::make_integer_sequence<int, XT_WMAX>{});
}

size_t bin_scan_scratch_bytes(u32 nbins) { return ((size_t)(nbins + BS_TILE - 1) / BS_TILE + 2) * sizeof(u64); }

cudaError_t launch_bin_scan(const u64 *bin_tot, u32 nbins, u64 *bin_start, u64 *bin_cursor, u64 *kmers_total, u64 *scratch,
                            cudaStream_t s)
{
    const u32 ntiles = (nbins + BS_TILE - 1) / BS_TILE;
    k_bin_tile_sums<<<ntiles, BS_THREADS, 0, s>>>(bin_tot, nbins, scratch, kmers_total);
    k_bin_tile_scan<<<1, BS_THREADS, 0, s>>>(scratch, ntiles, bin_start + nbins);
    k_bin_starts<<<ntiles, BS_THREADS, 0, s>>>(bin_tot, nbins, scratch, bin_start, bin_cursor);
    return cudaGetLastError();
}

cudaError_t launch_supermer_scatter(const ExtractParams &P, u32 nctas, int nwords, bool ext, const u64 *run_list,
                                    const ulonglong2 *tile_hdr, u64 *bin_cursor, cudaStream_t s)
{
#define HSK_SC(SW_, EXT_) k_supermer_scatter<SW_, EXT_><<<nctas, XT_THREADS, 0, s>>>(P, run_list, tile_hdr, bin_cursor)
    if (nwords == 1) { if (ext) HSK_SC(8, true); else HSK_SC(4, false); }
    else { if (ext) HSK_SC(12, true); else HSK_SC(8, false); }
#undef HSK_SC
    return cudaGetLastError();
}

} // namespace hsk
