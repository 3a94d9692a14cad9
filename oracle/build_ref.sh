#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference (sources read in place from
# $HSK_REFERENCE, default /root/reference; never copied) against the single-rank MPI
# shim into oracle/_ref/ as a shared library with the C wrapper in ref_wrapper.cpp,
# plus the reference's own standalone CLI.  One build per compile-time configuration.
#
#   oracle/build_ref.sh K M L U EXT [SORT]
#
# Flags are the reference's own (Makefile:59-63: -O3 -mavx2 -fopenmp -std=c++17); the
# reference's build system itself is not run.  kmerops.cpp takes ~3-4 min (small_sort.h).
set -euo pipefail
K=${1:-31}; M=${2:-17}; L=${3:-2}; U=${4:-50}; EXT=${5:-0}; SORT=${6:-2}
REF=${HSK_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
TAG=k${K}_m${M}_l${L}_u${U}_e${EXT}
OBJ=$OUT/obj_$TAG
if [ ! -d "$REF/src" ]; then echo "reference tree not found at $REF (skipping)"; exit 0; fi
mkdir -p "$OBJ"
DEFS="-DKMER_SIZE=$K -DMINIMIZER_SIZE=$M -DLOWER_KMER_FREQ=$L -DUPPER_KMER_FREQ=$U -DLOG_LEVEL=${LOG:-0} -DDEBUG=0 \
 -DTHREAD_PER_WORKER=4 -DMAX_SEND_BATCH=80000 -DMAX_THREAD_MEMORY_BOUNDED=16 -DSORT=$SORT -DAVG_TASK_PER_WORKER=3 \
 -DDISPATCH_UPPER_COE=1.5 -DDISPATCH_STEP=0.05 -DUNBALANCED_RATIO=2.3 -DPLAIN_CLASSIFIER=0 -DPLAIN_DISPATCHER=0 \
 -DEXTENSION=$EXT"
FLAGS="-O3 -pthread -m64 -mavx2 -DTHREADED -fopenmp -std=c++17 -fPIC -w \
 -I$HERE/../hysortk_b200/shim -I$REF/include -I$REF/src -I$REF/dependency/Raduls -I$REF/dependency/Paradis"
pids=()
for f in logger dnaseq dnabuffer fastaindex hashfuncs kmerops memcheck hysortk; do
  if [ ! -f "$OBJ/$f.o" ]; then g++ $DEFS $FLAGS -c -o "$OBJ/$f.o" "$REF/src/$f.cpp" & pids+=($!); fi
done
# RADULS' sorting networks (dependency/Raduls/Makefile builds this one file)
if [ ! -f "$OBJ/sorting_network.o" ]; then
  g++ -O3 -mavx2 -fPIC -w -std=c++17 -c -o "$OBJ/sorting_network.o" "$REF/dependency/Raduls/sorting_network.cpp" & pids+=($!)
fi
if [ ! -f "$OBJ/standalone.o" ]; then g++ $DEFS $FLAGS -c -o "$OBJ/standalone.o" "$REF/standalone/main.cpp" & pids+=($!); fi
g++ $DEFS $FLAGS -c -o "$OBJ/ref_wrapper.o" "$HERE/ref_wrapper.cpp" & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
OBJS="$OBJ/logger.o $OBJ/dnaseq.o $OBJ/dnabuffer.o $OBJ/fastaindex.o $OBJ/hashfuncs.o $OBJ/kmerops.o $OBJ/memcheck.o $OBJ/hysortk.o $OBJ/sorting_network.o"
g++ -shared -fopenmp -o "$OUT/libhysortk_ref_$TAG.so" $OBJ/ref_wrapper.o $OBJS -lpthread
g++ -fopenmp -o "$OUT/hysortk_ref_$TAG" $OBJ/standalone.o $OBJS -lpthread
echo "built $OUT/libhysortk_ref_$TAG.so and $OUT/hysortk_ref_$TAG"
