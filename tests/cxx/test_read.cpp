// Host-only exercise of hysortk::read_dna_buffer: prints "<nreads> <bufsize>" and then "<len> <hex bytes>" per read,
// so the Python test can compare with the packing of the same reads (reference src/dnaseq.cpp:9-31).
#include <cstdio>
#include <mpi.h>
#include "hysortk.hpp"

int main(int argc, char **argv)
{
    MPI_Init(&argc, &argv);
    const double t0 = MPI_Wtime();
    auto dna = hysortk::read_dna_buffer(argv[1], MPI_COMM_WORLD);
    const double t1 = MPI_Wtime();
    std::printf("%zu %zu\n", dna->size(), dna->getbufsize());
    if (argc > 2) {   // timing only
        std::printf("read_dna_buffer: %.3f s, %.1f Mbases/s\n", t1 - t0, dna->getbufsize() * 4.0 / (t1 - t0) / 1e6);
        MPI_Finalize();
        return 0;
    }
    for (size_t i = 0; i < dna->size(); ++i) {
        const hysortk::DnaSeq& s = (*dna)[i];
        std::printf("%zu ", s.size());
        for (size_t b = 0; b < s.numbytes(); ++b) std::printf("%02x", s.data()[b]);
        std::printf("\n");
    }
    MPI_Finalize();
    return 0;
}
