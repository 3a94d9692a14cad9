// Hash functions callers of the hysortk API can reach (reference include/hashfuncs.hpp).  Kmer::GetHash is
// MurmurHash3 x64-128, seed 313, low word — restated in hysortk_b200/cxx/hashfuncs.cpp from the published algorithm.
// The CUDA engine does not use them: which bin counts a k-mer is free to differ from the reference.
// (The reference also declares wanghash64 / wanghash64_inv; nothing on the kmer_count path uses them.)
//
//   `key`       the bytes to hash (for a k-mer: its 8 * NLONGS word bytes, Kmer::GetBytes)
//   `numbytes`  how many of them
//   `out`       receives 16, 8 or 4 bytes, in the byte order of the host
// tests/test_cxx_api.py checks murmurhash3_64 of every k-mer word against the oracle's independent restatement.
#ifndef HYSORTK_HASH_FUNCS_H
#define HYSORTK_HASH_FUNCS_H

#include <cstddef>
#include <cstdint>

namespace hysortk {

/* x64-128 variant, seed 313: 16 bytes to `out` */
void murmurhash3_128(const void *key, uint32_t numbytes, void *out);
/* the low 8 bytes of murmurhash3_128: what Kmer::GetHash and the reference's minimizer order use */
void murmurhash3_64(const void *key, uint32_t numbytes, void *out);
/* the low 4 bytes */
void murmurhash3_32(const void *key, uint32_t numbytes, void *out);
/* x86-32 variant with an explicit seed */
uint32_t murmurhash3(const void *key, size_t len, uint32_t seed);

} // namespace hysortk

#endif
