// DnaBuffer: all reads of a rank in one contiguous 2-bit packed byte array plus the DnaSeq views.
// Source-compatible with the reference's class (reference include/dnabuffer.hpp:14-47): same
// constructors, accessors and ownership (the buffer is new[]-allocated and deleted by the
// destructor; the (bufsize, numreads, buf, readlens) constructor adopts `buf`).
// This is the input format of hysortk::kmer_count; its bytes go to the GPU unchanged.
#ifndef HYSORTK_DNABUFFER_H_
#define HYSORTK_DNABUFFER_H_

#include "dnaseq.hpp"
#include <cstddef>
#include <cstdint>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

namespace hysortk {

class DnaBuffer
{
public:
    DnaBuffer(size_t bufsize) : bufhead(0), bufsize(bufsize), buf(new uint8_t[bufsize]) {}
    DnaBuffer(size_t bufsize, size_t numreads, uint8_t *buf, const size_t *readlens);
    DnaBuffer(const DnaBuffer& other);
    DnaBuffer& operator=(const DnaBuffer&) = delete;

    void push_back(char const *s, size_t len);
    size_t size() const { return sequences.size(); }
    size_t getbufsize() const { return bufsize; }
    size_t getrangebufsize(size_t start, size_t count) const;
    const uint8_t* getbufoffset(size_t i) const { return sequences[i].data(); }
    const DnaSeq& operator[](size_t i) const { return sequences[i]; }

    std::string getasciifilecontents() const;

    static size_t computebufsize(const std::vector<size_t>& seqlens);

    ~DnaBuffer() { delete[] buf; }

private:
    size_t bufhead;
    const size_t bufsize;
    uint8_t *buf;
    std::vector<DnaSeq> sequences;
};

} // namespace hysortk

#endif
