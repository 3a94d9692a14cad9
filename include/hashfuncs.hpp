// Hash functions of the hysortk API (reference include/hashfuncs.hpp): MurmurHash3 x64-128 with
// seed 313 is what Kmer::GetHash exposes to callers.  The CUDA engine does not use it (the bucket
// hash is free to differ).
#ifndef HYSORTK_HASH_FUNCS_H
#define HYSORTK_HASH_FUNCS_H

#include <cstddef>
#include <cstdint>

namespace hysortk {

void murmurhash3_128(const void *key, uint32_t numbytes, void *out);
void murmurhash3_64(const void *key, uint32_t numbytes, void *out);
void murmurhash3_32(const void *key, uint32_t numbytes, void *out);
uint32_t murmurhash3(const void *key, size_t len, uint32_t seed);

} // namespace hysortk

#endif
