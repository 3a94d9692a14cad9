// MurmurHash3 x64-128 (Austin Appleby's published, public-domain algorithm) with the seed the
// hysortk API uses (313; reference src/hashfuncs.cpp:226-245).  Only exposed through Kmer::GetHash.
#include "hashfuncs.hpp"
#include <cstring>

namespace hysortk {

namespace {
inline uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t avalanche(uint64_t k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}
constexpr uint64_t C1 = 0x87c37b91114253d5ULL, C2 = 0x4cf5ad432745937fULL;

void mm3_x64_128(const void *key, uint32_t len, uint32_t seed, uint64_t out[2])
{
    const uint8_t *p = static_cast<const uint8_t *>(key);
    uint64_t h1 = seed, h2 = seed;
    uint32_t left = len;
    for (; left >= 16; left -= 16, p += 16) {
        uint64_t k1, k2;
        std::memcpy(&k1, p, 8);
        std::memcpy(&k2, p + 8, 8);
        k1 *= C1; k1 = rotl(k1, 31); k1 *= C2; h1 ^= k1;
        h1 = rotl(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= C2; k2 = rotl(k2, 33); k2 *= C1; h2 ^= k2;
        h2 = rotl(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    uint64_t t1 = 0, t2 = 0;   /* little-endian tail words */
    if (left > 8) std::memcpy(&t2, p + 8, left - 8);
    if (left > 0) std::memcpy(&t1, p, left > 8 ? 8 : left);
    if (left > 8) { t2 *= C2; t2 = rotl(t2, 33); t2 *= C1; h2 ^= t2; }
    if (left > 0) { t1 *= C1; t1 = rotl(t1, 31); t1 *= C2; h1 ^= t1; }
    h1 ^= len; h2 ^= len;
    h1 += h2; h2 += h1;
    h1 = avalanche(h1); h2 = avalanche(h2);
    h1 += h2; h2 += h1;
    out[0] = h1; out[1] = h2;
}
} // namespace

void murmurhash3_128(const void *key, uint32_t numbytes, void *out)
{
    uint64_t v[2];
    mm3_x64_128(key, numbytes, 313, v);
    std::memcpy(out, v, 16);
}

void murmurhash3_64(const void *key, uint32_t numbytes, void *out)
{
    uint64_t v[2];
    mm3_x64_128(key, numbytes, 313, v);
    std::memcpy(out, v, 8);
}

void murmurhash3_32(const void *key, uint32_t numbytes, void *out)
{
    uint64_t v[2];
    mm3_x64_128(key, numbytes, 313, v);
    uint32_t lo = static_cast<uint32_t>(v[0]);
    std::memcpy(out, &lo, 4);
}

uint32_t murmurhash3(const void *key, size_t len, uint32_t seed)
{
    uint64_t v[2];
    mm3_x64_128(key, static_cast<uint32_t>(len), seed, v);
    return static_cast<uint32_t>(v[0]);
}

} // namespace hysortk
