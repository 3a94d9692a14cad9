// DnaSeq: a (length, pointer) view of one 2-bit packed read inside a DnaBuffer.
// Source-compatible with the reference's class (reference include/dnaseq.hpp:33-172): same public
// members, same packing (src/dnaseq.cpp:9-31: 4 bases per byte, first base in bits 7..6, codes
// A0 C1 G2 T3, N -> A, tail bits zero).  The packed bytes are what the CUDA engine reads directly.
#ifndef HYSORTK_DNASEQ_H_
#define HYSORTK_DNASEQ_H_

#include <array>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <string>

namespace hysortk {

namespace detail {
constexpr std::array<uint8_t, 256> make_codetab()
{
    std::array<uint8_t, 256> t{};
    for (int i = 0; i < 256; ++i) t[i] = 4;
    t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3;
    t['N'] = t['n'] = 0;
    return t;
}
} // namespace detail

class DnaSeq
{
public:
    DnaSeq() : len(0), memory(nullptr) {}
    DnaSeq(size_t len, uint8_t *mem) : len(len), memory(mem) {}
    /* encodes the ASCII sequence s into mem (which must hold bytesneeded(len) bytes) */
    DnaSeq(char const *s, size_t len, uint8_t *mem) : DnaSeq(len, mem) { compress(s); }
    DnaSeq(const DnaSeq& rhs) : len(rhs.len), memory(rhs.memory) {}
    DnaSeq& operator=(const DnaSeq& rhs) = default;

    std::string ascii() const;
    size_t size() const { return len; }
    size_t numbytes() const { return bytesneeded(len); }
    int remainder() const { return static_cast<int>(4 * numbytes() - len); }
    const uint8_t* data() const { return memory; }

    void copyto(size_t *readlen, uint8_t *mem) const
    {
        *readlen = len;
        std::memcpy(mem, memory, numbytes());
    }

    int operator[](size_t i) const { return (memory[i >> 2] >> (6 - 2 * (i & 3))) & 3; }
    bool operator<(const DnaSeq& rhs);
    bool operator==(const DnaSeq& rhs);
    bool operator!=(const DnaSeq& rhs) { return !(*this == rhs); }

    int regular_at(size_t i) const { return (*this)[i]; }
    int revcomp_at(size_t i) const { return 3 - (*this)[len - 1 - i]; }

    static size_t bytesneeded(size_t n) { return (n + 3) / 4; }

    static char    getcodechar(int c)  { return chartab[c]; }
    static uint8_t getcharcode(char c) { return codetab[static_cast<unsigned char>(c)]; }
    static char    getcharchar(char c) { return getcodechar(getcharcode(c)); }

    static constexpr char chartab[4 + 1] = {'A', 'C', 'G', 'T', 'X'};
    static constexpr std::array<uint8_t, 256> codetab = detail::make_codetab();

    friend std::ostream& operator<<(std::ostream& stream, const DnaSeq& s)
    {
        stream << s.ascii();
        return stream;
    }

private:
    size_t len;      /* number of bases */
    uint8_t *memory; /* 4 bases per byte, owned by the DnaBuffer */

    void compress(char const *s);
};

} // namespace hysortk

#endif
