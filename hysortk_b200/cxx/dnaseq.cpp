// DnaSeq host code (reference src/dnaseq.cpp): 2-bit encoding / decoding of one read.
#include "dnaseq.hpp"
#include <algorithm>

namespace hysortk {

constexpr char DnaSeq::chartab[];
constexpr std::array<uint8_t, 256> DnaSeq::codetab;

/* 4 bases per byte, first base in the two most significant bits, unused tail bits zero
 * (reference src/dnaseq.cpp:9-31).  Characters outside ACGTN map to code 4, of which the two low
 * bits are kept, as in the reference (undefined input there). */
void DnaSeq::compress(char const *s)
{
    const size_t nbytes = numbytes();
    size_t p = 0;
    for (size_t b = 0; b < nbytes; ++b) {
        unsigned byte = 0;
        for (int i = 0; i < 4 && p < len; ++i, ++p) byte |= (getcharcode(s[p]) << (6 - 2 * i)) & 0xFFu;
        memory[b] = static_cast<uint8_t>(byte);
    }
}

std::string DnaSeq::ascii() const
{
    std::string s(len, 'A');
    for (size_t i = 0; i < len; ++i) s[i] = getcodechar((*this)[i]);
    return s;
}

bool DnaSeq::operator==(const DnaSeq& rhs)
{
    if (len != rhs.len) return false;
    for (size_t i = 0; i < len; ++i)
        if ((*this)[i] != rhs[i]) return false;
    return true;
}

bool DnaSeq::operator<(const DnaSeq& rhs)
{
    const size_t n = std::min(len, rhs.len);
    for (size_t i = 0; i < n; ++i) {
        const int a = (*this)[i], b = rhs[i];
        if (a != b) return a < b;
    }
    return false;
}

} // namespace hysortk
