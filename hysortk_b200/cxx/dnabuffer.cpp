// DnaBuffer host code (reference src/dnabuffer.cpp).
#include "dnabuffer.hpp"
#include <cassert>
#include <stdexcept>

namespace hysortk {

/* adopts `buf`: reads lie back to back, each on a fresh byte (reference src/dnabuffer.cpp:7-16) */
DnaBuffer::DnaBuffer(size_t bufsize, size_t numreads, uint8_t *buf, const size_t *readlens)
    : bufhead(0), bufsize(bufsize), buf(buf)
{
    sequences.reserve(numreads);
    for (size_t i = 0; i < numreads; ++i) {
        sequences.emplace_back(readlens[i], buf + bufhead);
        bufhead += DnaSeq::bytesneeded(readlens[i]);
    }
}

/* deep copy; the views are rebuilt against the new storage (reference include/dnabuffer.hpp:19-27) */
DnaBuffer::DnaBuffer(const DnaBuffer& other) : bufhead(other.bufhead), bufsize(other.bufsize), buf(new uint8_t[other.bufsize])
{
    std::memcpy(buf, other.buf, bufsize);
    sequences.reserve(other.size());
    size_t at = 0;
    for (size_t i = 0; i < other.size(); ++i) {
        sequences.emplace_back(other[i].size(), buf + at);
        at += other[i].numbytes();
    }
}

size_t DnaBuffer::computebufsize(const std::vector<size_t>& seqlens)
{
    size_t total = 0;
    for (size_t l : seqlens) total += DnaSeq::bytesneeded(l);
    return total;
}

void DnaBuffer::push_back(char const *s, size_t len)
{
    const size_t nbytes = DnaSeq::bytesneeded(len);
    if (bufhead + nbytes > bufsize) throw std::length_error("DnaBuffer::push_back: buffer full");
    sequences.emplace_back(s, len, buf + bufhead);
    bufhead += nbytes;
}

size_t DnaBuffer::getrangebufsize(size_t start, size_t count) const
{
    if (count == 0) return 0;
    const DnaSeq& last = sequences[start + count - 1];
    return static_cast<size_t>((last.data() + last.numbytes()) - sequences[start].data());
}

std::string DnaBuffer::getasciifilecontents() const
{
    std::string out;
    for (const auto& s : sequences) { out += s.ascii(); out += '\n'; }
    return out;
}

} // namespace hysortk
