// Stage 1+2: minimizer scan, supermer extraction and owner-hash bucketing, straight from the
// 2-bit packed DnaBuffer.
//
// Replaces the reference's HOT LOOPS A+B: FindKmerDestinationsParallel (src/kmerops.cpp:1010-1041,
// canonical m-mers supermer.hpp:315-342, sliding-window minimum :1058-1073, owner :1044-1047) and
// SupermerEncoder::encode / copy_bits (:1096-1148), plus the per-thread ScatteredSupermers staging
// (:253-358).  Not a translation: the reference walks each read with a deque and materialises one
// int per k-mer; here the whole packed buffer is treated as one flat sequence of 2-bit slots,
// processed in tiles by persistent CTAs, in two passes:
//
//   pass A (k_supermer_count), per tile
//     A. tile bytes -> shared memory as big-endian 32-bit words (16 bases per word)
//     B. every thread rolls forward/reverse m-mers over 16 consecutive slots, hashes the canonical
//        m-mer of every slot into shared memory, and derives a 16-bit "valid k-mer start" mask from
//        the read offset table
//     C. window minimum over the K-M+1 hashes of every k-mer slot (block-wise suffix/prefix minima,
//        O(1) shared-memory traffic per slot) -> bin id (or INVALID)
//     D. run boundaries (bin change / validity change / tile edge) -> bitmap -> compacted run list
//     E. per valid run: reduce (count, words, k-mers) into the global per-bin totals, and store the
//        run list (start, bin) of the tile for pass B
//   k_bin_scan: exclusive prefix of the per-bin totals -> bin starts in the supermer streams
//   pass B (k_supermer_scatter), per tile: re-stage the bytes, read the run list, and for every
//     valid run claim its slot(s) in the bin with one atomic and write the re-packed bases, length and
//     optional (pos, rid).  No hashing is repeated.
//
// A supermer is a run of consecutive k-mers of one read with the same bin, stored in fixed-size slots
// (common.cuh: 16 bytes for K <= 32: 60 bases + length; longer runs are split into overlapping pieces),
// so that the scatter is one atomic + one 128-bit store per supermer.  Bins are
// fine-grained (a few thousand k-mers each) so that a bin can later be expanded, sorted and counted
// entirely inside one CTA's shared memory.  Where the reference splits supermers (250-base cap,
// kmerops.cpp:1120) and how it hashes are free choices: only the multiset of k-mers per bin matters,
// and a canonical k-mer always lands in the same bin because the bin is a function of its set of
// canonical m-mers.
#include "kernels.cuh"

namespace hsk {

struct ExtractSmem {
    u32 wbe[EX_WORDS];                 // bases of the tile, 16 per word, first base in the top bits
    u32 hs[EX_TS + EX_TS / 32 + 8];    // m-mer hashes, then in-place suffix minima (1 pad word per 32)
    u32 pm[EX_TS + EX_TS / 32 + 8];    // prefix minima inside blocks of w
    u16 runs[EX_TSK + 8];              // compacted run starts (+ sentinel)
    u32 bm[EX_TS / 32];                // run-boundary bitmap
    u32 woff[EX_TS / 32 + 1];          // exclusive popcount prefix of bm
    u16 vm[EX_THREADS];                // valid-start mask of the thread's 16 slots
    u64 rlo, rhi;                      // reads overlapping the tile
    u64 run_base;                      // where this tile's run list starts in the global list
    u32 nruns;
};

__device__ __forceinline__ int hx(int q) { return q + (q >> 5); }

// state of a k-mer slot once the block-wise minima are in place: its minimizer hash (low bit forced to 1),
// or 0 when no k-mer starts there.  Runs are maximal stretches of equal state; the bin is derived from the
// state once per run.
__device__ __forceinline__ u32 slot_state(const ExtractSmem &sm, int q, int w)
{
    if (!((sm.vm[q >> 4] >> (q & 15)) & 1)) return 0u;
    return min(sm.hs[hx(q)], sm.pm[hx(q + w - 1)]) | 1u;
}

// largest r in [lo, hi] with off[r] <= byte (off is non-decreasing; caller guarantees off[lo] <= byte)
__device__ __forceinline__ u64 find_read(const u64 *__restrict__ off, u64 lo, u64 hi, u64 byte)
{
    while (lo < hi) {
        u64 mid = lo + (hi - lo + 1) / 2;
        if (__ldg(off + mid) <= byte) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ void stage_tile(ExtractSmem &sm, const ExtractParams &P, u64 tile)
{
    const int tid = threadIdx.x;
    const u64 byte0 = tile * (u64)EX_TILE_BYTES;
    if (tid < EX_WORDS / 4) {
        u64 b = byte0 + (u64)tid * 16;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (b < P.nbytes_padded) v = __ldg(reinterpret_cast<const uint4 *>(P.packed + b));
        sm.wbe[4 * tid + 0] = __byte_perm(v.x, 0, 0x0123);
        sm.wbe[4 * tid + 1] = __byte_perm(v.y, 0, 0x0123);
        sm.wbe[4 * tid + 2] = __byte_perm(v.z, 0, 0x0123);
        sm.wbe[4 * tid + 3] = __byte_perm(v.w, 0, 0x0123);
    }
    if (tid == 32) {
        // reads overlapping [byte0, byte0 + tile bytes): read_off[nreads] = nbytes
        u64 lo = 0, hi = 0;
        if (P.nreads > 0 && byte0 < P.nbytes) {
            lo = find_read(P.read_off, 0, P.nreads - 1, byte0);
            hi = find_read(P.read_off, lo, P.nreads - 1, byte0 + EX_TILE_BYTES);
        }
        sm.rlo = lo; sm.rhi = hi;
    }
}

// Phases A-D for one tile.  On return sm.runs[0..nruns] / sm.st describe the runs of the tile.
__device__ __forceinline__ void tile_runs(ExtractSmem &sm, const ExtractParams &P, u64 tile)
{
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const u64 slot0 = tile * (u64)EX_TSK;

    stage_tile(sm, P, tile);
    __syncthreads();

    // ---- B: rolling canonical m-mer hashes of 16 consecutive slots
    {
        const int m = P.m;
        u64 hi = ((u64)sm.wbe[tid] << 32) | sm.wbe[tid + 1];
        u64 lo = ((u64)sm.wbe[tid + 2] << 32);
        const u64 mask = (m == 32) ? ~0ull : ((1ull << (2 * m)) - 1);
        const int rcs = 2 * (m - 1);
        // m-mer at my first slot: the top m bases of the window, and their reverse complement
        u64 fwd = hi >> (64 - 2 * m);
        u64 rc = revcomp64(hi) & mask;
        // consume the m bases
        if (m == 32) { hi = lo; lo = 0; } else { hi = (hi << (2 * m)) | (lo >> (64 - 2 * m)); lo <<= 2 * m; }
        {
            u64 canon = fwd < rc ? fwd : rc;
            sm.hs[hx(tid * EX_R)] = mmer_hash(canon);
        }
#pragma unroll
        for (int j = 1; j < EX_R; ++j) {
            u64 c = hi >> 62;
            hi = (hi << 2) | (lo >> 62); lo <<= 2;
            fwd = ((fwd << 2) | c) & mask;
            rc = (rc >> 2) | ((3 - c) << rcs);
            u64 canon = fwd < rc ? fwd : rc;
            sm.hs[hx(tid * EX_R + j)] = mmer_hash(canon);
        }

        // valid k-mer starts among my 16 slots
        u32 vmask = 0;
        const u64 p0 = slot0 + (u64)tid * EX_R;
        if (tid * EX_R < EX_TSK && P.nreads > 0 && (p0 >> 2) < P.nbytes) {
            u64 r = find_read(P.read_off, sm.rlo, sm.rhi, p0 >> 2);
            u64 rstart = __ldg(P.read_off + r) * 4;
            u64 rnext = __ldg(P.read_off + r + 1) * 4;
            u64 rend = rstart + __ldg(P.read_len + r);
            if (p0 + EX_R + (u64)P.k <= rend + 1 && p0 >= rstart) {
                vmask = 0xFFFFu;   // common case: all 16 windows inside the read
            } else {
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
        }
        sm.vm[tid] = (u16)vmask;
    }
    __syncthreads();

    // ---- C: minimizer = minimum of the w = K-M+1 hashes starting at the slot.
    // Blocks of w slots: pm = prefix minima inside a block, hs becomes suffix minima inside a block;
    // min over [q, q+w) = min(suffix[q], prefix[q+w-1]).
    {
        const int w = P.k - P.m + 1;
        const int nblocks = (EX_TS + w - 1) / w;
        for (int b = tid; b < nblocks; b += EX_THREADS) {
            const int s = b * w, e = min(s + w, (int)EX_TS);
            u32 run = 0xFFFFFFFFu;
            for (int q = s; q < e; ++q) { run = min(run, sm.hs[hx(q)]); sm.pm[hx(q)] = run; }
            run = 0xFFFFFFFFu;
            for (int q = e - 1; q >= s; --q) { run = min(run, sm.hs[hx(q)]); sm.hs[hx(q)] = run; }
        }
    }
    __syncthreads();

    // ---- D: run boundaries (state change / tile edge) -> bitmap -> compacted run starts
    {
        const int w = P.k - P.m + 1;
#pragma unroll 4
        for (int j = 0; j < EX_R; ++j) {
            const int q = j * EX_THREADS + tid;
            const u32 st = (q < EX_TSK) ? slot_state(sm, q, w) : 0u;
            u32 prev = __shfl_up_sync(0xFFFFFFFFu, st, 1);
            if (lane == 0 && q > 0 && q < EX_TSK) prev = slot_state(sm, q - 1, w);
            const bool b = (q < EX_TSK) && (q == 0 || st != prev);
            const u32 bal = __ballot_sync(0xFFFFFFFFu, b);
            if (lane == 0) sm.bm[j * (EX_THREADS / 32) + warp] = bal;
        }
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

// ---- pass A: per-bin totals + the run list of every tile ------------------------------------------
// bin_tot[b] = (slots << 40) | k-mers.  Run list entry = (n << 48 | start << 32 | bin) for valid runs
// only; tile_hdr[tile] = (first entry, number of entries).
__global__ void __launch_bounds__(EX_THREADS) k_supermer_count(ExtractParams P, u64 *__restrict__ bin_tot,
                                                                u64 *__restrict__ run_list,
                                                                ulonglong2 *__restrict__ tile_hdr,
                                                                u64 *__restrict__ run_cursor, u64 run_capacity)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    ExtractSmem &sm = *reinterpret_cast<ExtractSmem *>(smraw);
    __shared__ u32 s_nvalid, s_claim;
    if (threadIdx.x == 0) { s_nvalid = 0; s_claim = 0; }
    const u64 t0 = (u64)blockIdx.x * P.tiles_per_cta;
    const u64 t1 = min(t0 + P.tiles_per_cta, P.ntiles);
    for (u64 tile = t0; tile < t1; ++tile) {
        tile_runs(sm, P, tile);
        const u32 nruns = sm.nruns;
        const int w = P.k - P.m + 1;
        // pass 1 over the runs: number of valid runs of the tile -> its place in the global run list
        {
            u32 cntv = 0;
            for (u32 j = threadIdx.x; j < nruns; j += EX_THREADS) cntv += (slot_state(sm, sm.runs[j], w) != 0) ? 1u : 0u;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) cntv += __shfl_xor_sync(0xFFFFFFFFu, cntv, d);
            if ((threadIdx.x & 31) == 0 && cntv) atomicAdd(&s_nvalid, cntv);
        }
        __syncthreads();
        const u32 nv = s_nvalid;
        if (threadIdx.x == 0) {
            sm.run_base = atomicAdd(run_cursor, (u64)nv);
            tile_hdr[tile] = make_ulonglong2(sm.run_base, (u64)nv);
            s_nvalid = 0; s_claim = 0;
        }
        __syncthreads();
        const u64 rb = sm.run_base;
        const bool fits = (rb + nv <= run_capacity);   // otherwise the host sees run_cursor > capacity and retries
        // pass 2: per-bin totals and the run list entries (order inside a tile is irrelevant)
        for (u32 base = 0; base < nruns; base += EX_THREADS) {
            const u32 j = base + threadIdx.x;
            bool valid = false;
            u32 start = 0, st = 0, n = 0;
            if (j < nruns) {
                start = sm.runs[j];
                st = slot_state(sm, start, w);
                valid = (st != 0);
                n = sm.runs[j + 1] - start;
            }
            const u32 bal = __ballot_sync(0xFFFFFFFFu, valid);
            u32 wbase = 0;
            if ((threadIdx.x & 31) == 0 && bal) wbase = atomicAdd(&s_claim, (u32)__popc(bal));
            wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
            if (valid) {
                const u32 b = hash_bucket(st, P.nbins);
                const u32 pieces = (n + P.slot_nmax - 1) / P.slot_nmax;
                atomicAdd(&bin_tot[b], ((u64)pieces << 40) | (u64)n);
                const u32 slot = wbase + __popc(bal & ((1u << (threadIdx.x & 31)) - 1));
                if (fits) run_list[rb + slot] = ((u64)(start | (n << 16)) << 32) | b;
            }
        }
        __syncthreads();
    }
}

// ---- bin scan: exclusive prefix over bins of the slot counts -> bin starts; k-mer total.  One block. ---
__global__ void __launch_bounds__(1024) k_bin_scan(const u64 *__restrict__ bin_tot, u32 nbins, u64 *__restrict__ bin_start,
                                                    u64 *__restrict__ kmers_total)
{
    __shared__ u64 s_c[32];
    __shared__ u64 carry_c;
    u64 ksum = 0;
    if (threadIdx.x == 0) carry_c = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int PER = 4;
    for (u32 base = 0; base < nbins; base += 1024 * PER) {
        u64 c[PER], tc = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            u32 b = base + threadIdx.x * PER + i;
            u64 v = b < nbins ? bin_tot[b] : 0;
            ksum += v & ((1ull << 40) - 1);
            c[i] = v >> 40;
            tc += c[i];
        }
        u64 ic = tc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 a = __shfl_up_sync(0xFFFFFFFFu, ic, d);
            if (lane >= d) ic += a;
        }
        if (lane == 31) s_c[warp] = ic;
        __syncthreads();
        if (warp == 0) {
            u64 a = s_c[lane], ia = a;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u64 x = __shfl_up_sync(0xFFFFFFFFu, ia, d);
                if (lane >= d) ia += x;
            }
            s_c[lane] = ia - a;
        }
        __syncthreads();
        u64 ec = carry_c + s_c[warp] + ic - tc;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            u32 b = base + threadIdx.x * PER + i;
            if (b < nbins) bin_start[b] = ec;
            ec += c[i];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_c = ec;
        __syncthreads();
    }
    if (threadIdx.x == 0) bin_start[nbins] = carry_c;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) ksum += __shfl_xor_sync(0xFFFFFFFFu, ksum, d);
    if (lane == 0 && ksum) atomicAdd(kmers_total, ksum);
}

// ---- pass B: one slot per piece of every valid run ----------------------------------------------------
template <int SW, bool EXT>
__global__ void __launch_bounds__(EX_THREADS) k_supermer_scatter(ExtractParams P, const u64 *__restrict__ run_list,
                                                                  const ulonglong2 *__restrict__ tile_hdr,
                                                                  u32 *__restrict__ bin_cursor,
                                                                  const u64 *__restrict__ bin_start,
                                                                  u32 *__restrict__ out_slots)
{
    constexpr int PW = SW - (EXT ? 2 : 0);
    __shared__ u32 wbe[EX_WORDS];
    __shared__ u64 s_rlo, s_rhi;
    const int tid = threadIdx.x;
    const u64 t0 = (u64)blockIdx.x * P.tiles_per_cta;
    const u64 t1 = min(t0 + P.tiles_per_cta, P.ntiles);
    for (u64 tile = t0; tile < t1; ++tile) {
        const u64 byte0 = tile * (u64)EX_TILE_BYTES;
        if (tid < EX_WORDS / 4) {
            u64 b = byte0 + (u64)tid * 16;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (b < P.nbytes_padded) v = __ldg(reinterpret_cast<const uint4 *>(P.packed + b));
            wbe[4 * tid + 0] = __byte_perm(v.x, 0, 0x0123);
            wbe[4 * tid + 1] = __byte_perm(v.y, 0, 0x0123);
            wbe[4 * tid + 2] = __byte_perm(v.z, 0, 0x0123);
            wbe[4 * tid + 3] = __byte_perm(v.w, 0, 0x0123);
        }
        if (EXT && tid == 32) {
            u64 lo = 0, hi = 0;
            if (P.nreads > 0 && byte0 < P.nbytes) {
                lo = find_read(P.read_off, 0, P.nreads - 1, byte0);
                hi = find_read(P.read_off, lo, P.nreads - 1, byte0 + EX_TILE_BYTES);
            }
            s_rlo = lo; s_rhi = hi;
        }
        const ulonglong2 hdr = tile_hdr[tile];
        __syncthreads();
        const u64 slot0 = tile * (u64)EX_TSK;
        for (u32 j = tid; j < (u32)hdr.y; j += EX_THREADS) {
            const u64 e = __ldg(run_list + hdr.x + j);
            const u32 b = (u32)e;
            const u32 start = (u32)(e >> 32) & 0xFFFFu, n = (u32)(e >> 48);
            const u32 pieces = (n + P.slot_nmax - 1) / P.slot_nmax;
            u64 gs = __ldg(bin_start + b) + atomicAdd(&bin_cursor[b], pieces);
            u32 pos0 = 0, rid = 0;
            if (EXT) {
                const u64 p = slot0 + start;
                const u64 r = find_read(P.read_off, s_rlo, s_rhi, p >> 2);
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
                    u32 wv = __funnelshift_l(wbe[(q >> 4) + 1], wbe[q >> 4], 2 * (q & 15));
                    const int rem = (int)len - 16 * x;
                    if (rem <= 0) wv = 0; else if (rem < 16) wv &= ~0u << (32 - 2 * rem);
                    w[x] = wv;
                }
                w[PW - 1] = (w[PW - 1] & 0xFFFFFF00u) | len;
                if (EXT) { w[SW - 2] = pos0 + pc * P.slot_nmax; w[SW - 1] = rid; }
                uint4 *dst = reinterpret_cast<uint4 *>(out_slots + gs * SW);
#pragma unroll
                for (int x = 0; x < SW / 4; ++x) dst[x] = make_uint4(w[4 * x], w[4 * x + 1], w[4 * x + 2], w[4 * x + 3]);
            }
        }
        __syncthreads();
    }
}

cudaError_t launch_supermer_count(const ExtractParams &P, u32 nctas, u64 *bin_tot, u64 *run_list, ulonglong2 *tile_hdr,
                                  u64 *run_cursor, u64 run_capacity, cudaStream_t s)
{
    cudaError_t e = cudaFuncSetAttribute(k_supermer_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ExtractSmem));
    if (e != cudaSuccess) return e;
    k_supermer_count<<<nctas, EX_THREADS, sizeof(ExtractSmem), s>>>(P, bin_tot, run_list, tile_hdr, run_cursor, run_capacity);
    return cudaGetLastError();
}

cudaError_t launch_bin_scan(const u64 *bin_tot, u32 nbins, u64 *bin_start, u64 *kmers_total, cudaStream_t s)
{
    k_bin_scan<<<1, 1024, 0, s>>>(bin_tot, nbins, bin_start, kmers_total);
    return cudaGetLastError();
}

cudaError_t launch_supermer_scatter(const ExtractParams &P, u32 nctas, int nwords, bool ext, const u64 *run_list,
                                    const ulonglong2 *tile_hdr, u32 *bin_cursor, const u64 *bin_start, u32 *out_slots,
                                    cudaStream_t s)
{
#define HSK_SC(SW_, EXT_) k_supermer_scatter<SW_, EXT_><<<nctas, EX_THREADS, 0, s>>>(P, run_list, tile_hdr, bin_cursor, bin_start, out_slots)
    if (nwords == 1) { if (ext) HSK_SC(8, true); else HSK_SC(4, false); }
    else { if (ext) HSK_SC(12, true); else HSK_SC(8, false); }
#undef HSK_SC
    return cudaGetLastError();
}

} // namespace hsk
