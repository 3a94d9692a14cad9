// Read table of a DnaBuffer on the device: byte offset and 32-bit length of every read from the 64-bit
// lengths the host API hands over (reference: DnaBuffer keeps one DnaSeq per read, each starting on a
// fresh byte, include/dnabuffer.hpp:14-47; bytes per read = (len + 3) / 4, include/dnaseq.hpp:126).
// Three small kernels (tile sums, scan of the sums, offsets) so that hsk_count uploads the lengths as they
// are and the prefix sum does not run on the host.
#include "kernels.cuh"

namespace hsk {

constexpr int RD_THREADS = 256;
constexpr int RD_IPT = 8;
constexpr int RD_TILE = RD_THREADS * RD_IPT;

__global__ void __launch_bounds__(RD_THREADS) k_read_tile_sums(const u64 *__restrict__ len64, u64 n, u64 *__restrict__ tile_sums,
                                                                u32 *__restrict__ flags)
{
    __shared__ u64 s_w[RD_THREADS / 32];
    const u64 base = (u64)blockIdx.x * RD_TILE;
    u64 sum = 0;
    bool bad = false;
#pragma unroll
    for (int i = 0; i < RD_IPT; ++i) {
        const u64 r = base + (u64)i * RD_THREADS + threadIdx.x;
        if (r < n) {
            const u64 l = len64[r];
            bad = bad || (l > 0xFFFFFFFFull);
            sum += (l + 3) >> 2;
        }
    }
    if (bad) atomicOr(flags, 1u);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 t = 0;
        for (int w = 0; w < RD_THREADS / 32; ++w) t += s_w[w];
        tile_sums[blockIdx.x] = t;
    }
}

// exclusive scan of the tile sums in place (one block); total -> *total_out, compared with the buffer size
__global__ void __launch_bounds__(1024) k_read_tile_scan(u64 *__restrict__ tile_sums, u64 ntiles, u64 nbytes, u64 *__restrict__ total_out,
                                                          u32 *__restrict__ flags)
{
    __shared__ u64 s_c[32];
    __shared__ u64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u64 base = 0; base < ntiles; base += 1024) {
        const u64 t = base + threadIdx.x;
        const u64 v = t < ntiles ? tile_sums[t] : 0;
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
        if (t < ntiles) tile_sums[t] = ex;
        __syncthreads();
        if (threadIdx.x == 1023) carry = ex + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *total_out = carry;
        if (carry != nbytes) atomicOr(flags, 2u);
    }
}

__global__ void __launch_bounds__(RD_THREADS) k_read_offsets(const u64 *__restrict__ len64, u64 n, const u64 *__restrict__ tile_base,
                                                              u64 *__restrict__ read_off, u32 *__restrict__ read_len)
{
    __shared__ u64 s_w[RD_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 first = (u64)blockIdx.x * RD_TILE + (u64)threadIdx.x * RD_IPT;   // RD_IPT consecutive reads per thread
    u64 b[RD_IPT], sum = 0;
#pragma unroll
    for (int i = 0; i < RD_IPT; ++i) {
        const u64 r = first + i;
        u64 l = 0;
        if (r < n) { l = len64[r]; read_len[r] = (u32)l; }
        b[i] = (l + 3) >> 2;
        sum += b[i];
    }
    u64 inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u64 x = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += x;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    u64 off = tile_base[blockIdx.x] + inc - sum;
    for (int w = 0; w < warp; ++w) off += s_w[w];
#pragma unroll
    for (int i = 0; i < RD_IPT; ++i) {
        const u64 r = first + i;
        if (r < n) read_off[r] = off;
        if (r == n) read_off[n] = off;   // closing offset = total bytes
        off += b[i];
    }
}

size_t read_table_scratch_bytes(u64 nreads) { return ((nreads + 1 + RD_TILE - 1) / RD_TILE + 2) * sizeof(u64); }

// flags: bit 0 = a read longer than 2^32-1 bases, bit 1 = the lengths do not add up to nbytes
cudaError_t launch_read_table(const u64 *len64, u64 nreads, u64 nbytes, u64 *read_off, u32 *read_len, u64 *scratch, u32 *flags,
                              cudaStream_t s)
{
    const u64 ntiles = (nreads + 1 + RD_TILE - 1) / RD_TILE;   // the closing offset is element nreads
    k_read_tile_sums<<<(unsigned)ntiles, RD_THREADS, 0, s>>>(len64, nreads, scratch, flags);
    k_read_tile_scan<<<1, 1024, 0, s>>>(scratch, ntiles, nbytes, scratch + ntiles, flags);
    k_read_offsets<<<(unsigned)ntiles, RD_THREADS, 0, s>>>(len64, nreads, scratch, read_off, read_len);
    return cudaGetLastError();
}

} // namespace hsk
