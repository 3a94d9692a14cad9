// DnaBuffer host code (behaviour of the reference's src/dnabuffer.cpp + the inline members of its header).
#include "dnabuffer.hpp"

#include <cstring>
#include <stdexcept>

namespace hysortk {

DnaBuffer::DnaBuffer(size_t bufsize) : store_(new uint8_t[bufsize]), capacity_(bufsize), used_(0) {}

/* adopts `buf`: reads lie back to back, each on a fresh byte (reference src/dnabuffer.cpp:7-16) */
DnaBuffer::DnaBuffer(size_t bufsize, size_t numreads, uint8_t *buf, const size_t *readlens)
    : store_(buf), capacity_(bufsize), used_(0)
{
    reads_.reserve(numreads);
    for (size_t i = 0; i < numreads; ++i) {
        reads_.emplace_back(readlens[i], store_ + used_);
        used_ += DnaSeq::bytesneeded(readlens[i]);
    }
}

/* deep copy; the views are rebuilt against the new storage (reference include/dnabuffer.hpp:19-27) */
DnaBuffer::DnaBuffer(const DnaBuffer& other) : store_(new uint8_t[other.capacity_]), capacity_(other.capacity_), used_(other.used_)
{
    std::memcpy(store_, other.store_, capacity_);
    reads_.reserve(other.size());
    size_t at = 0;
    for (size_t i = 0; i < other.size(); ++i) {
        reads_.emplace_back(other[i].size(), store_ + at);
        at += other[i].numbytes();
    }
}

DnaBuffer::~DnaBuffer()
{
    if (on_release_) on_release_(store_);
    delete[] store_;
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
    if (used_ + nbytes > capacity_) throw std::length_error("DnaBuffer::push_back: buffer full");
    reads_.emplace_back(s, len, store_ + used_);
    used_ += nbytes;
}

size_t DnaBuffer::getrangebufsize(size_t start, size_t count) const
{
    if (count == 0) return 0;
    const DnaSeq& last = reads_[start + count - 1];
    return static_cast<size_t>((last.data() + last.numbytes()) - reads_[start].data());
}

std::string DnaBuffer::getasciifilecontents() const
{
    std::string out;
    for (const auto& s : reads_) { out += s.ascii(); out += '\n'; }
    return out;
}

} // namespace hysortk
