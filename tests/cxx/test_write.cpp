// Host-only exercise of hysortk::write_output_file: a KmerListS built from k-mer strings given on stdin
// ("<kmer> <count>" per line) is written to argv[1]/0.out; the Python test compares the file with the expected lines.
#include <iostream>
#include <string>
#include <mpi.h>
#include "hysortk.hpp"

int main(int argc, char **argv)
{
    MPI_Init(&argc, &argv);
    hysortk::KmerListS list;
    std::string km;
    unsigned long long cnt;
    while (std::cin >> km >> cnt) {
        hysortk::KmerListEntryS e;
        e.kmer = hysortk::TKmer(km.c_str());
        e.cnt = cnt;
        list.push_back(e);
    }
    hysortk::write_output_file(list, argv[1], MPI_COMM_WORLD);
    MPI_Finalize();
    return 0;
}
