// Shared definitions of the hysortk_b200 CUDA engine (sm_100a).
//
// Data contract (reference include/kmer.hpp:165-185, src/dnaseq.cpp:9-31):
//   * reads: 4 bases per byte, first base in bits 7..6, codes A0 C1 G2 T3, each read on a fresh byte
//   * k-mer: base i in word i/32 at bits 2*(31 - i%32)+1..; word 0 most significant; unused low bits 0
//   * canonical k-mer = min(forward, reverse complement) under word-0-first comparison
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace hsk {

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;

constexpr int MAX_WORDS = 3;

struct Planes {           // key planes of one k-mer array: p[w][i] = word w of k-mer i
    u64 *p[MAX_WORDS];
};

__host__ __device__ __forceinline__ int nwords_for_k(int k) { return k <= 32 ? 1 : (k <= 64 ? 2 : 3); }

// ---- bit helpers -----------------------------------------------------------------------------

// reverse complement of 32 bases held in a u64 (base 0 in the top bits): complement every base and
// reverse the base order.
__host__ __device__ __forceinline__ u64 revcomp64(u64 x)
{
    x = ~x;
#ifdef __CUDA_ARCH__
    x = __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0f0f0f0f0f0f0f0full) | ((x & 0x0f0f0f0f0f0f0f0full) << 4);
    x = ((x >> 8) & 0x00ff00ff00ff00ffull) | ((x & 0x00ff00ff00ff00ffull) << 8);
    x = ((x >> 16) & 0x0000ffff0000ffffull) | ((x & 0x0000ffff0000ffffull) << 16);
    x = (x >> 32) | (x << 32);
#endif
    // the bit reversal also swapped the two bits of every base; swap them back
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

// Reverse complement of a K-mer held in NW words (layout above).  Restates the effect of the
// reference's Kmer::GetTwin (include/kmer.hpp:265-296) with bit reversal instead of its tetramer table.
template <int NW>
__host__ __device__ __forceinline__ void kmer_twin(const u64 (&in)[NW], int k, u64 (&out)[NW])
{
    u64 t[NW];
#pragma unroll
    for (int l = 0; l < NW; ++l) t[NW - 1 - l] = revcomp64(in[l]);
    const int shift = 2 * (32 * NW - k);   // left-align: 0..62
    if (shift == 0) {
#pragma unroll
        for (int l = 0; l < NW; ++l) out[l] = t[l];
    } else {
#pragma unroll
        for (int l = 0; l < NW; ++l) {
            u64 v = t[l] << shift;
            if (l + 1 < NW) v |= t[l + 1] >> (64 - shift);
            out[l] = v;
        }
    }
}

// canonical representative (reference Kmer::GetRep, include/kmer.hpp:298-303; operator< :216-229)
template <int NW>
__host__ __device__ __forceinline__ void kmer_canonical(u64 (&w)[NW], int k)
{
    u64 t[NW];
    kmer_twin<NW>(w, k, t);
    bool less = false, decided = false;
#pragma unroll
    for (int l = 0; l < NW; ++l) {
        if (!decided && t[l] != w[l]) { less = t[l] < w[l]; decided = true; }
    }
    if (less) {
#pragma unroll
        for (int l = 0; l < NW; ++l) w[l] = t[l];
    }
}

// 32-bit hash of a canonical m-mer value (right-aligned, <= 64 bits).  The choice of hash only
// decides which bucket counts a k-mer (reference uses MurmurHash3, supermer.hpp:307-313); the
// result of kmer_count does not depend on it.
__host__ __device__ __forceinline__ u32 mmer_hash(u64 x)
{
    u32 lo = (u32)x, hi = (u32)(x >> 32);
    u32 h = lo * 0x9E3779B1u;
    h ^= (hi + 0x7F4A7C15u) * 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du;
    h ^= h >> 13;
    return h;
}

// bucket of a minimizer hash among nb buckets.  The minimum of a window of hashes is a small number
// (its top bits are almost always zero), so it is re-scrambled by an odd multiplier before the
// multiplicative range reduction; the reference takes hash % tot_tasks (kmerops.cpp:1044-1047).
__host__ __device__ __forceinline__ u32 hash_bucket(u32 h, u32 nb)
{
    h *= 0x9E3779B1u;
    h ^= h >> 15;
    return (u32)(((u64)h * nb) >> 32);
}

// ---- extraction geometry ---------------------------------------------------------------------
// The flat slot space (slot = 2 bits of the packed buffer, padding slots included) is cut into warp tiles:
// every lane owns XT_R consecutive slots; the last xt_halo_lanes(w) lanes of a warp only supply m-mer hashes
// for the minimizer windows (w = K - M + 1 hashes per k-mer) of the lanes before them.
constexpr int XT_R = 16;                       // slots per lane (= bases per 32-bit word of packed input)
constexpr int XT_WARPS = 8;                    // warps per CTA
constexpr int XT_THREADS = XT_WARPS * 32;
constexpr int XT_CTAS_PER_SM = 5;              // register budget of the extraction kernels: 40 warps per SM
constexpr int XT_WMAX = 64;                    // widest minimizer window
constexpr int XT_STAGE_WORDS = 40;             // words of bases a scatter warp stages: 32 + ceil((K_max - 1) / 16) + 1
__host__ __device__ constexpr int xt_halo_lanes(int w) { return (w + 14) >> 4; }
__host__ __device__ constexpr int xt_out_lanes(int w) { return 32 - xt_halo_lanes(w); }
__host__ __device__ constexpr int xt_out_slots(int w) { return xt_out_lanes(w) * XT_R; }
constexpr u32 MAX_BINS = 1u << 26;

// ---- per-bin totals ---------------------------------------------------------------------------
// One 64-bit word per bin and rank, accumulated with one reduction per run: supermer slots in the upper 28 bits,
// k-mers in the lower 36.  Pass A also keeps independent grand totals; the bin scan compares them with the sums of the
// two fields, so a bin that outgrows a field is reported instead of corrupting the stream (engine.cu: extract_count).
constexpr int BT_KBITS = 36;
constexpr u64 BT_KMASK = (1ull << BT_KBITS) - 1;
__host__ __device__ __forceinline__ u64 bt_pack(u64 slots, u64 kmers) { return (slots << BT_KBITS) | kmers; }
__host__ __device__ __forceinline__ u64 bt_slots(u64 v) { return v >> BT_KBITS; }
__host__ __device__ __forceinline__ u64 bt_kmers(u64 v) { return v & BT_KMASK; }

// ---- supermer slots ---------------------------------------------------------------------------
// A supermer is stored in one fixed-size slot of SW 32-bit words (16-byte multiples, so a slot is
// written and read with 128-bit accesses and addressed by its index alone):
//   words 0 .. PW-1   bases, 16 per word from the top bits; the last payload word holds 12 bases in its
//                     upper 24 bits and the supermer length (bases) in its low 8 bits
//   words PW, PW+1    PosInRead of the first base, ReadId (EXTENSION only)
// K <= 32: SW = 4 (60 bases), with EXTENSION SW = 8 (92 bases); K > 32: SW = 8 (124 bases), with
// EXTENSION SW = 12 (156 bases).  Runs longer than a slot are split into pieces that overlap by K-1.
__host__ __device__ __forceinline__ int slot_words(int nwords, bool ext) { return (nwords == 1 ? 4 : 8) + (ext ? 4 : 0); }
__host__ __device__ __forceinline__ int slot_payload_words(int nwords, bool ext) { return slot_words(nwords, ext) - (ext ? 2 : 0); }
__host__ __device__ __forceinline__ int slot_max_bases(int nwords, bool ext) { return 16 * (slot_payload_words(nwords, ext) - 1) + 12; }

// ---- radix geometry --------------------------------------------------------------------------
constexpr int RS_THREADS = 384;
constexpr int RS_IPT = 16;
constexpr int RS_TILE = RS_THREADS * RS_IPT;   // 6144 keys per tile
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_MAX_PASSES = 24;

struct PassDesc { int plane; int shift; };
struct PassTable { int npasses; PassDesc d[RS_MAX_PASSES]; };

// LSD pass schedule over the significant bits of a k-mer of size k: 8-bit digits, never
// straddling a word; the last word's unused low bits are skipped.
inline PassTable make_pass_table(int k)
{
    PassTable t;
    t.npasses = 0;
    int nw = nwords_for_k(k);
    for (int pl = nw - 1; pl >= 0; --pl) {
        int bases = (pl == nw - 1) ? (k - 32 * (nw - 1)) : 32;
        int low = 64 - 2 * bases;
        for (int s = low; s < 64; s += 8) {
            t.d[t.npasses].plane = pl;
            t.d[t.npasses].shift = s;
            ++t.npasses;
        }
    }
    return t;
}

} // namespace hsk
