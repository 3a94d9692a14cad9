// Compile-time parameters of the hysortk API, same macro names and limits as the reference
// (reference include/compiletime.h:7-22, Makefile:39-46).  They only configure the C++ shim; the
// CUDA engine behind it takes them at run time (include/hsk_capi.h: hsk_config).
#ifndef HYSORTK_COMPILE_TIME_H_
#define HYSORTK_COMPILE_TIME_H_

#include <cstdint>
#include <limits>

#if !defined(KMER_SIZE)
#error "KMER_SIZE must be defined (make K=...)"
#endif
#if !defined(MINIMIZER_SIZE)
#error "MINIMIZER_SIZE must be defined (make M=...)"
#endif
#if !defined(LOWER_KMER_FREQ) || !defined(UPPER_KMER_FREQ)
#error "LOWER_KMER_FREQ and UPPER_KMER_FREQ must be defined (make L=... U=...)"
#endif
#ifndef EXTENSION
#define EXTENSION 0
#endif
#ifndef LOG_LEVEL
#define LOG_LEVEL 0
#endif

static_assert(KMER_SIZE > 2 && KMER_SIZE < 96, "2 < KMER_SIZE < 96");
static_assert(MINIMIZER_SIZE > 0 && MINIMIZER_SIZE < KMER_SIZE, "0 < MINIMIZER_SIZE < KMER_SIZE");
static_assert(LOWER_KMER_FREQ > 0 && LOWER_KMER_FREQ <= UPPER_KMER_FREQ &&
                  UPPER_KMER_FREQ <= std::numeric_limits<uint16_t>::max(),
              "0 < LOWER_KMER_FREQ <= UPPER_KMER_FREQ <= 65535");

namespace hysortk {
typedef int32_t MPI_Count_t;
typedef int32_t MPI_Offset_t;
#define MPI_COUNT_TYPE MPI_INT
} // namespace hysortk

#endif
