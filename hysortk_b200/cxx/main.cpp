// Standalone driver, same command line as the reference CLI (reference standalone/main.cpp:9-72):
//   hysortk <fasta file> [output dir]
#include <iomanip>
#include <iostream>
#include <mpi.h>
#include "hysortk.hpp"

int main(int argc, char **argv)
{
    MPI_Init(&argc, &argv);
    if (argc < 2) {
        std::cerr << "Usage: " << argv[0] << " <fasta file> <output dir>(Optional)" << std::endl;
        return 1;
    }
    const std::string fasta = argv[1];
    const std::string outdir = argc >= 3 ? argv[2] : "";
    int rank, nranks;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &nranks);
    if (rank == 0) {
        std::cout << "Compiling Parameters:\n"
                  << "      KMER_SIZE: " << KMER_SIZE << "\n      EXTENSION: " << EXTENSION
                  << "\n      MINIMIZER_SIZE: " << MINIMIZER_SIZE << "\n      LOWER_KMER_FREQ: " << LOWER_KMER_FREQ
                  << "\n      UPPER_KMER_FREQ: " << UPPER_KMER_FREQ << "\n      LOGGING_LEVEL: " << LOG_LEVEL << "\n\n"
                  << "Runtime Parameters:\n      Fasta File: " << std::quoted(fasta) << "\n      Output Directory: "
                  << std::quoted(outdir) << "\n      Nprocs:" << nranks << "\n      Engine: hysortk_b200 (sm_100a)\n" << std::endl;
    }
    try {
        auto dna = hysortk::read_dna_buffer(fasta, MPI_COMM_WORLD);
        auto kmer_list = hysortk::kmer_count(*dna, MPI_COMM_WORLD);
        hysortk::print_kmer_histogram(*kmer_list, MPI_COMM_WORLD);
        if (!outdir.empty()) hysortk::write_output_file(*kmer_list, outdir, MPI_COMM_WORLD);
    } catch (const std::exception& e) {
        std::cerr << "hysortk: " << e.what() << std::endl;
        MPI_Abort(MPI_COMM_WORLD, 1);
    }
    hysortk::release_gpu_engine();
    MPI_Finalize();
    return 0;
}
