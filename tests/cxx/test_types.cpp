// Host-only exercise of the public value types (include/kmer.hpp, dnaseq.hpp, dnabuffer.hpp):
// prints, for every k-mer of a few reads, "<kmer> <twin> <rep> <hash>" so the Python test can compare
// against the oracle's restatement of the reference's arithmetic.
#include <iostream>
#include <string>
#include <vector>
#include "dnabuffer.hpp"
#include "kmer.hpp"

using namespace hysortk;

int main(int argc, char **argv)
{
    std::vector<std::string> reads;
    for (int i = 1; i < argc; ++i) reads.push_back(argv[i]);
    std::vector<size_t> lens;
    for (auto& r : reads) lens.push_back(r.size());
    DnaBuffer buf(DnaBuffer::computebufsize(lens));
    for (auto& r : reads) buf.push_back(r.c_str(), r.size());
    DnaBuffer copy(buf);
    std::cout << "reads " << copy.size() << " bytes " << copy.getbufsize() << " range " << copy.getrangebufsize(0, copy.size()) << "\n";
    for (size_t i = 0; i < copy.size(); ++i) {
        std::cout << "read " << copy[i].ascii() << " " << copy[i].numbytes() << " " << copy[i].remainder() << "\n";
        auto kmers = TKmer::GetKmers(copy[i]);
        auto reps = TKmer::GetRepKmers(copy[i]);
        for (size_t j = 0; j < kmers.size(); ++j) {
            TKmer fromstr(kmers[j].GetString().c_str());
            if (fromstr != kmers[j]) { std::cout << "MISMATCH string ctor\n"; return 1; }
            std::cout << kmers[j] << " " << kmers[j].GetTwin() << " " << reps[j] << " " << reps[j].GetHash() << "\n";
        }
    }
    std::cout << "sizeof " << sizeof(TKmer) << " " << sizeof(KmerListEntryS) << " " << sizeof(KmerSeedStruct) << "\n";
    return 0;
}
