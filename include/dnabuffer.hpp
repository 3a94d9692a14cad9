// DnaBuffer — all reads of one rank, 2-bit packed back to back in ONE contiguous byte array, plus a DnaSeq view
// per read.  This is the input of hysortk::kmer_count, and its byte array is what the GPU reads, unchanged.
//
// The class is source-compatible with the reference's (reference include/dnabuffer.hpp:14-47: same public
// members with the same meaning), so callers written against HySortK keep compiling; the implementation is in
// hysortk_b200/cxx/dnabuffer.cpp.
//
//   layout      read i occupies DnaSeq::bytesneeded(len_i) = (len_i + 3) / 4 bytes and starts on a fresh byte
//               right after read i-1 (reference src/dnabuffer.cpp:7-16, src/dnaseq.cpp:9-31)
//   ownership   the byte array is new[]-allocated and released by the destructor; the four-argument
//               constructor ADOPTS the array it is given
#ifndef HYSORTK_DNABUFFER_H_
#define HYSORTK_DNABUFFER_H_

#include <cstddef>
#include <cstdint>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "dnaseq.hpp"

namespace hysortk {

class DnaBuffer
{
    uint8_t *store_;                // the packed bytes of every read
    const size_t capacity_;         // size of store_ in bytes
    size_t used_;                   // bytes taken by the reads appended so far
    std::vector<DnaSeq> reads_;     // views into store_, one per read
    void (*on_release_)(uint8_t *) = nullptr;   // called with store_ before it is freed (see set_release_hook)

public:
    // ---- construction ------------------------------------------------------------------------------
    /* empty buffer with room for `bufsize` packed bytes (see computebufsize); fill it with push_back */
    DnaBuffer(size_t bufsize);
    /* a buffer that was packed elsewhere: `buf` holds `numreads` reads of the given lengths back to back and
     * now belongs to this object */
    DnaBuffer(size_t bufsize, size_t numreads, uint8_t *buf, const size_t *readlens);
    /* deep copy (read_dna_buffer hands out a copy, reference src/hysortk.cpp:28) */
    DnaBuffer(const DnaBuffer& other);
    DnaBuffer& operator=(const DnaBuffer&) = delete;
    ~DnaBuffer();

    /* bytes needed for reads of these lengths */
    static size_t computebufsize(const std::vector<size_t>& seqlens);

    /* 2-bit encodes the `len` characters at `s` (ACGT, N -> A) as the next read */
    void push_back(char const *s, size_t len);

    // ---- access ------------------------------------------------------------------------------------
    size_t size() const { return reads_.size(); }                              /* number of reads */
    const DnaSeq& operator[](size_t i) const { return reads_[i]; }
    size_t getbufsize() const { return capacity_; }
    const uint8_t* getbufoffset(size_t i) const { return reads_[i].data(); }   /* first byte of read i */
    size_t getrangebufsize(size_t start, size_t count) const;                  /* bytes of reads [start, start+count) */

    /* hysortk_b200 addition: `f(bytes)` is called right before the destructor frees the byte array.  read_dna_buffer
     * page-locks the array of the buffer it returns (so that kmer_count sends it to the GPU from where it is) and uses
     * the hook to undo that.  Not copied by the copy constructor. */
    void set_release_hook(void (*f)(uint8_t *)) { on_release_ = f; }

    /* the reads as text, one per line */
    std::string getasciifilecontents() const;
};

} // namespace hysortk

#endif
