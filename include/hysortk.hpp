// hysortk.hpp — the public API of HySortK, kept as it is (reference include/hysortk.hpp:8-18) so that ELBA-style
// callers switch libraries without touching their code.  Only the body of kmer_count is different: it runs on the
// B200 through the C ABI of include/hsk_capi.h (hysortk_b200/cxx/hysortk.cpp -> libhysortk_b200).
//
// Build-time parameters are the reference's -D macros (KMER_SIZE, MINIMIZER_SIZE, LOWER_KMER_FREQ, UPPER_KMER_FREQ,
// EXTENSION, LOG_LEVEL; include/compiletime.h) and must be the same for the caller and the library.
// Every function is collective over `comm` (one rank per GPU) and reports errors by throwing std::runtime_error.
//
// Typical caller (the reference's standalone/main.cpp does the same):
//
//     MPI_Init(&argc, &argv);
//     auto dna  = hysortk::read_dna_buffer("reads.fa", MPI_COMM_WORLD);     // needs reads.fa.fai
//     auto list = hysortk::kmer_count(*dna, MPI_COMM_WORLD);                // the GPU path
//     hysortk::print_kmer_histogram(*list, MPI_COMM_WORLD);
//     hysortk::write_output_file(*list, "outdir", MPI_COMM_WORLD);          // outdir/<rank>.out
//     MPI_Finalize();
//
// Link: obj/libhysortk.o (make K= M= L= U= EXT=) -L$CUDA/lib64 -lcudart -ldl -lpthread -fopenmp, plus the MPI library.
// Rank -> GPU: LOCAL_RANK / OMPI_COMM_WORLD_LOCAL_RANK / SLURM_LOCALID, see INTEGRATION.md.
#ifndef HYSORTK_H_
#define HYSORTK_H_

#include <mpi.h>

#include "dnabuffer.hpp"
#include "kmer.hpp"

namespace hysortk {

/* Reads this rank's contiguous share (balanced by bases) of an indexed FASTA file (`fasta_fname` + ".fai") and returns
 * it 2-bit packed.  reference src/hysortk.cpp:18-33, src/fastaindex.cpp */
std::shared_ptr<DnaBuffer> read_dna_buffer(const std::string& fasta_fname, MPI_Comm comm);

/* Counts the canonical k-mers of all ranks' reads; returns the entries this rank owns with
 * LOWER_KMER_FREQ <= count <= UPPER_KMER_FREQ (with EXTENSION: plus ReadId / PosInRead of every occurrence), as
 * sorted runs.  `mydna` is only read.  reference src/hysortk.cpp:36-95 */
std::unique_ptr<KmerListS> kmer_count(const DnaBuffer& mydna, MPI_Comm comm);

/* Prints "#count\tnumkmers" + one line per non-empty count on rank 0, summed over the ranks.
 * reference src/hysortk.cpp:98-136 */
void print_kmer_histogram(const KmerListS& kmerlist, MPI_Comm comm);

/* Writes "<k-mer>\t<count>" per entry to <output_dir>/<rank>.out.  reference src/hysortk.cpp:138-164 */
void write_output_file(const KmerListS& kmerlist, const std::string& output_dir, MPI_Comm comm);

/* hysortk_b200 addition (not in the reference): frees the GPU engine of this process — device memory, streams, NCCL
 * communicator.  Collective over the communicator of the last kmer_count; call it before MPI_Finalize.  Optional: without
 * it the engine lives until the process exits (it is never torn down from a static destructor). */
void release_gpu_engine();

} // namespace hysortk

#endif
