// K-mer value type and result structs of the hysortk API.
//
// Source- and layout-compatible with the reference's include/kmer.hpp: Kmer<NLONGS> is NLONGS
// 64-bit words, base i in word i/32 at bits 2*(31 - i%32)+1 .. 2*(31 - i%32), word 0 most
// significant, unused low bits of the last word zero (reference kmer.hpp:165-185); TKmer picks
// NLONGS from KMER_SIZE (:343-345); KmerListEntryS = { TKmer kmer; uint64_t cnt; } plus
// std::vector<PosInRead> pos and std::vector<ReadId> rid when EXTENSION == 1 (:368-407).  Callers
// reinterpret these structs (reference README.md:67-70), so the layout is ABI.
// The arithmetic is written independently (bit-reversal reverse complement instead of the
// reference's tetramer table); the CUDA kernels implement the same functions on the device.
#ifndef HYSORTK_KMER_H_
#define HYSORTK_KMER_H_

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <string>
#include <type_traits>
#include <vector>

#include "compiletime.h"
#include "dnaseq.hpp"
#include "hashfuncs.hpp"

namespace hysortk {

namespace detail {
/* reverse complement of the 32 bases of one word (base 0 in the top two bits) */
inline uint64_t revcomp_word(uint64_t x)
{
    x = ~x;
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return __builtin_bswap64(x);
}
} // namespace detail

template <int NLONGS>
class Kmer
{
public:
    static_assert(NLONGS != 0, "unsupported KMER_SIZE");

    static constexpr int NBYTES = 8 * NLONGS;
    /* bits by which a full-width reverse complement must be shifted left to be left-aligned */
    static constexpr int PAD_BITS = 2 * (32 * NLONGS - KMER_SIZE);

    typedef std::array<uint64_t, NLONGS> MERARR;
    typedef std::array<uint8_t, NBYTES> BYTEARR;

    Kmer() : longs{} {}
    Kmer(const DnaSeq& s) : Kmer()
    {
        for (int i = 0; i < KMER_SIZE; ++i) put(i, static_cast<uint64_t>(s[i]));
    }
    Kmer(char const *s) : Kmer()
    {
        for (int i = 0; i < KMER_SIZE; ++i) put(i, static_cast<uint64_t>(DnaSeq::getcharcode(s[i]) & 3));
    }
    Kmer(const void *mem) : Kmer() { CopyDataFrom(mem); }
    Kmer(const Kmer& o) : longs(o.longs) {}

    Kmer& operator=(Kmer o) { longs = o.longs; return *this; }

    std::string GetString() const
    {
        std::string s(KMER_SIZE, 'A');
        for (int i = 0; i < KMER_SIZE; ++i) s[i] = "ACGT"[(longs[i >> 5] >> (2 * (31 - (i & 31)))) & 3];
        return s;
    }

    bool operator<(const Kmer& o) const { return longs < o.longs; }   /* word 0 first */
    bool operator==(const Kmer& o) const { return longs == o.longs; }
    bool operator!=(const Kmer& o) const { return !(longs == o.longs); }

    /* drop the first base, append `code` as the last one */
    Kmer GetExtension(int code) const
    {
        Kmer e;
        for (int l = 0; l < NLONGS; ++l) {
            e.longs[l] = longs[l] << 2;
            if (l + 1 < NLONGS) e.longs[l] |= longs[l + 1] >> 62;
        }
        e.longs[NLONGS - 1] |= static_cast<uint64_t>(code & 3) << (PAD_BITS % 64);
        return e;
    }

    /* reverse complement */
    Kmer GetTwin() const
    {
        std::array<uint64_t, NLONGS> t;
        for (int l = 0; l < NLONGS; ++l) t[NLONGS - 1 - l] = detail::revcomp_word(longs[l]);
        Kmer r;
        if (PAD_BITS == 0) {
            r.longs = t;
        } else {
            for (int l = 0; l < NLONGS; ++l) {
                r.longs[l] = t[l] << PAD_BITS;
                if (l + 1 < NLONGS) r.longs[l] |= t[l + 1] >> (64 - PAD_BITS);
            }
        }
        return r;
    }

    /* canonical representative: the smaller of the k-mer and its reverse complement */
    Kmer GetRep() const
    {
        Kmer t = GetTwin();
        return t < *this ? t : *this;
    }

    uint64_t GetHash() const
    {
        uint64_t h;
        murmurhash3_64(longs.data(), NBYTES, &h);
        return h;
    }

    const void* GetBytes() const { return reinterpret_cast<const void*>(longs.data()); }
    int getByte(int &i) const { return bytes[i]; }

    void CopyDataInto(void *mem) const { std::memcpy(mem, longs.data(), NBYTES); }
    void CopyDataFrom(const void *mem) { std::memcpy(longs.data(), mem, NBYTES); }

    static std::vector<Kmer> GetKmers(const DnaSeq& s)
    {
        std::vector<Kmer> out;
        const long n = static_cast<long>(s.size()) - KMER_SIZE + 1;
        if (n <= 0) return out;
        out.reserve(n);
        out.emplace_back(s);
        for (long i = 1; i < n; ++i) out.push_back(out.back().GetExtension(s[i + KMER_SIZE - 1]));
        return out;
    }

    static std::vector<Kmer> GetRepKmers(const DnaSeq& s)
    {
        std::vector<Kmer> out = GetKmers(s);
        for (auto& k : out) k = k.GetRep();
        return out;
    }

    template <int N>
    friend std::ostream& operator<<(std::ostream& os, const Kmer<N>& kmer);

private:
    union { MERARR  longs;
            BYTEARR bytes; };

    void put(int i, uint64_t code) { longs[i >> 5] |= code << (2 * (31 - (i & 31))); }
};

template <int NLONGS>
std::ostream& operator<<(std::ostream& os, const Kmer<NLONGS>& kmer)
{
    os << kmer.GetString();
    return os;
}

} // namespace hysortk

namespace std
{
    template <int NLONGS> struct hash<hysortk::Kmer<NLONGS>>
    {
        size_t operator()(const hysortk::Kmer<NLONGS>& kmer) const { return kmer.GetHash(); }
    };

    template <int NLONGS> struct less<hysortk::Kmer<NLONGS>>
    {
        bool operator()(const hysortk::Kmer<NLONGS>& a, const hysortk::Kmer<NLONGS>& b) const { return a < b; }
    };
}

namespace hysortk {

using TKmer = typename std::conditional<(KMER_SIZE <= 32), Kmer<1>,
              typename std::conditional<(KMER_SIZE <= 64), Kmer<2>,
              typename std::conditional<(KMER_SIZE <= 96), Kmer<3>, Kmer<0>>::type>::type>::type;

typedef uint32_t PosInRead;
typedef  int32_t ReadId;

/// One counted k-mer (and, with EXTENSION, where it occurs).
struct KmerListEntryS {
    TKmer kmer;
    uint64_t cnt;
#if EXTENSION == 1
    std::vector<PosInRead> pos;
    std::vector<ReadId> rid;
    KmerListEntryS(TKmer kmer, int cnt, PosInRead pos, ReadId rid) : kmer(kmer), cnt(cnt), pos({pos}), rid({rid}) {}
#endif
    KmerListEntryS(TKmer kmer, int cnt) : kmer(kmer), cnt(cnt) {}
    KmerListEntryS() {}
    KmerListEntryS(const KmerListEntryS&) = default;
    KmerListEntryS(KmerListEntryS&&) = default;
    KmerListEntryS& operator=(const KmerListEntryS&) = default;
    KmerListEntryS& operator=(KmerListEntryS&&) = default;

    bool operator < (const KmerListEntryS& o) const { return kmer < o.kmer; }
    bool operator == (const KmerListEntryS& o) const { return kmer == o.kmer; }
    bool operator != (const KmerListEntryS& o) const { return kmer != o.kmer; }
    int GetByte(int &i) const { return kmer.getByte(i); }
};

typedef std::vector<KmerListEntryS> KmerListS;
typedef std::vector<KmerListS> KmerListSVec;

/// One k-mer occurrence before counting.
struct KmerSeedStruct {
    TKmer kmer;
#if EXTENSION == 1
    PosInRead pos;
    ReadId rid;
    KmerSeedStruct(TKmer kmer, PosInRead pos, ReadId rid) : kmer(kmer), pos(pos), rid(rid) {}
#else
    KmerSeedStruct(TKmer kmer) : kmer(kmer) {}
#endif
    KmerSeedStruct() {}
    KmerSeedStruct(const KmerSeedStruct&) = default;
    KmerSeedStruct& operator=(const KmerSeedStruct&) = default;

    int GetByte(int &i) const { return kmer.getByte(i); }
    bool operator < (const KmerSeedStruct& o) const { return kmer < o.kmer; }
    bool operator == (const KmerSeedStruct& o) const { return kmer == o.kmer; }
    bool operator != (const KmerSeedStruct& o) const { return kmer != o.kmer; }
};

typedef std::vector<std::vector<KmerSeedStruct>> KmerSeedBuckets;
typedef std::vector<std::vector<std::vector<KmerSeedStruct>>> KmerSeedVecs;

static_assert(sizeof(TKmer) == TKmer::NBYTES, "TKmer must be exactly its words");
#if EXTENSION == 0
static_assert(sizeof(KmerListEntryS) == TKmer::NBYTES + 8, "KmerListEntryS layout is ABI: {words, cnt}");
#endif

} // namespace hysortk

#endif
