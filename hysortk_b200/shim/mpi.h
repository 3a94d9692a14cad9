/*
 * Single-rank MPI stand-in, used only when the build host has no MPI installation
 * (this image has none).  It lets `hysortk.hpp`-style code that is written against
 * the MPI C API compile and run as a 1-rank job: rank = 0, size = 1, collectives
 * degenerate to local copies.  With a real MPI on the include path this header is
 * simply not used.
 *
 * Covers the calls made on the kmer_count path and its I/O neighbours
 * (reference call sites: kmerops.cpp:66,711,782,919,940,1172,1198,1287,1325;
 * hysortk.cpp:104,115,156; fastaindex.cpp:137,168-185,223-224; logger.cpp:127,137;
 * timer.hpp:26-51; memcheck.cpp:81).
 *
 * A datatype handle is the element size in bytes.
 */
#ifndef HSK_MPI_SHIM_H_
#define HSK_MPI_SHIM_H_

#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSK_MPI_SHIM 1

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Info;
typedef int MPI_File;
typedef long long MPI_Offset;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_SUCCESS 0
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_IN_PLACE ((void *)-1)
#define MPI_INFO_NULL 0
#define MPI_MODE_RDONLY 2
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_BYTE 1
#define MPI_CHAR 1
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_FLOAT 4
#define MPI_DOUBLE 8
#define MPI_LONG 8
#define MPI_UNSIGNED_LONG 8
#define MPI_LONG_LONG 8
#define MPI_UNSIGNED_LONG_LONG 8
#define MPI_UINT64_T 8

static inline void hsk_shim_copy(const void *src, void *dst, size_t n)
{
    if (src != MPI_IN_PLACE && src != dst && n) memcpy(dst, src, n);
}

static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
static inline int MPI_Initialized(int *flag) { *flag = 1; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code); return MPI_SUCCESS; }
static inline double MPI_Wtime(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static inline int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm c)
{ (void)buf; (void)n; (void)t; (void)root; (void)c; return MPI_SUCCESS; }

static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{ (void)op; (void)root; (void)c; hsk_shim_copy(s, r, (size_t)n * (size_t)t); return MPI_SUCCESS; }

static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{ (void)op; (void)c; hsk_shim_copy(s, r, (size_t)n * (size_t)t); return MPI_SUCCESS; }

/* rank 0's receive buffer is undefined after MPI_Exscan; leave it untouched */
static inline int MPI_Exscan(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{ (void)s; (void)r; (void)n; (void)t; (void)op; (void)c; return MPI_SUCCESS; }

static inline int MPI_Gather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c)
{ (void)rn; (void)rt; (void)root; (void)c; hsk_shim_copy(s, r, (size_t)sn * (size_t)st); return MPI_SUCCESS; }

static inline int MPI_Allgather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c)
{ (void)rn; (void)rt; (void)c; hsk_shim_copy(s, r, (size_t)sn * (size_t)st); return MPI_SUCCESS; }

static inline int MPI_Gatherv(const void *s, int sn, MPI_Datatype st, void *r, const int *rn, const int *displs,
                              MPI_Datatype rt, int root, MPI_Comm c)
{
    (void)rn; (void)root; (void)c;
    hsk_shim_copy(s, (char *)r + (size_t)displs[0] * (size_t)rt, (size_t)sn * (size_t)st);
    return MPI_SUCCESS;
}

static inline int MPI_Scatterv(const void *s, const int *sn, const int *displs, MPI_Datatype st, void *r, int rn,
                               MPI_Datatype rt, int root, MPI_Comm c)
{
    (void)sn; (void)root; (void)c;
    hsk_shim_copy((const char *)s + (size_t)displs[0] * (size_t)st, r, (size_t)rn * (size_t)rt);
    return MPI_SUCCESS;
}

static inline int MPI_Alltoallv(const void *s, const int *sn, const int *sd, MPI_Datatype st, void *r, const int *rn,
                                const int *rd, MPI_Datatype rt, MPI_Comm c)
{
    (void)rn; (void)c;
    hsk_shim_copy((const char *)s + (size_t)sd[0] * (size_t)st, (char *)r + (size_t)rd[0] * (size_t)rt,
                  (size_t)sn[0] * (size_t)st);
    return MPI_SUCCESS;
}

static inline int MPI_Alltoall(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c)
{ (void)rn; (void)rt; (void)c; hsk_shim_copy(s, r, (size_t)sn * (size_t)st); return MPI_SUCCESS; }

static inline int MPI_Ialltoall(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c,
                                MPI_Request *req)
{ (void)rn; (void)rt; (void)c; *req = 0; hsk_shim_copy(s, r, (size_t)sn * (size_t)st); return MPI_SUCCESS; }

static inline int MPI_Wait(MPI_Request *req, MPI_Status *st) { (void)req; (void)st; return MPI_SUCCESS; }

static inline int MPI_Type_contiguous(int n, MPI_Datatype old, MPI_Datatype *nw) { *nw = n * old; return MPI_SUCCESS; }
static inline int MPI_Type_commit(MPI_Datatype *t) { (void)t; return MPI_SUCCESS; }
static inline int MPI_Type_free(MPI_Datatype *t) { (void)t; return MPI_SUCCESS; }

/* only the size query is used (to clamp a read range); report "unbounded" */
static inline int MPI_File_open(MPI_Comm c, const char *fn, int mode, MPI_Info info, MPI_File *fh)
{ (void)c; (void)fn; (void)mode; (void)info; *fh = 0; return MPI_SUCCESS; }
static inline int MPI_File_get_size(MPI_File fh, MPI_Offset *sz) { (void)fh; *sz = (MPI_Offset)1 << 62; return MPI_SUCCESS; }
static inline int MPI_File_close(MPI_File *fh) { (void)fh; return MPI_SUCCESS; }

#ifdef __cplusplus
}
#endif
#endif /* HSK_MPI_SHIM_H_ */
