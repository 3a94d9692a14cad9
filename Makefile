# hysortk_b200 — same `make K= M= L= U= EXT= LOG=` interface and the same artefacts as the reference
# (reference Makefile:1-46,99-109): `make` -> obj/libhysortk.o (relocatable: C++ API + CUDA engine),
# `make standalone` -> ./hysortk.  The reference's CPU tuning knobs (T, T2, TPW, SORT, BATCH,
# DISPATCH_*, UNBALANCED_THRESHOLD, PLAIN_*) are accepted and ignored: there are no tasks, worker
# threads or CPU sorters here.  Final links need:  -L$(CUDA)/lib64 -lcudart -ldl -lpthread
K?=31
M?=17
L?=15
U?=40
EXT?=0
LOG?=2
D?=0
T?=4
T2?=16
TPW?=3
SORT?=0
BATCH?=80000

OBJ?=obj
BIN?=hysortk
CUDA?=/usr/local/cuda
NVCC?=$(CUDA)/bin/nvcc
CXX=g++
# real MPI: make MPI_INC=-I/path/to/mpi/include MPI_LIB="-L... -lmpi"; default: bundled single-rank shim
MPI_INC?=-I./hysortk_b200/shim
MPI_LIB?=

ifneq ($(shell test $(M) -lt $(K) && echo 0 || echo 1), 0)
$(error ERROR: MINIMIZER_SIZE (M) must be less than KMER_SIZE (K))
endif

PARAMS=-DKMER_SIZE=$(K) -DMINIMIZER_SIZE=$(M) -DLOWER_KMER_FREQ=$(L) -DUPPER_KMER_FREQ=$(U) -DEXTENSION=$(EXT) \
       -DLOG_LEVEL=$(LOG) -DDEBUG=$(D)
ifeq ($(D), 1)
OPT=-g -O2 -fsanitize=address -fno-omit-frame-pointer
else
OPT=-O3
endif
CXXFLAGS=$(OPT) -std=c++17 -fPIC -fopenmp -pthread -Wall $(PARAMS) -I./include $(MPI_INC)
NVFLAGS=-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC

# the CUDA engine does not depend on K/M/L/U/EXT (they reach it at run time): CUOBJ=<dir> lets several configurations share
# one set of engine objects
CUOBJ?=$(OBJ)
HOST_OBJ=$(OBJ)/hysortk.o $(OBJ)/dnaseq.o $(OBJ)/dnabuffer.o $(OBJ)/hashfuncs.o
CUDA_OBJ=$(CUOBJ)/reads.o $(CUOBJ)/extract.o $(CUOBJ)/expand.o $(CUOBJ)/radix.o $(CUOBJ)/count.o $(CUOBJ)/bins.o $(CUOBJ)/engine.o

all: print lib

print:
	$(info ------ hysortk_b200 compile-time parameters ------ )
	$(info KMER_SIZE: $(K), MINIMIZER_SIZE: $(M), EXTENSION: $(EXT))
	$(info LOWER_KMER_FREQ: $(L), UPPER_KMER_FREQ: $(U), LOG_LEVEL: $(LOG), DEBUG: $(D))
	$(info -------------------------------------------------- )

lib: $(HOST_OBJ) $(CUDA_OBJ)
	ld -r -o $(OBJ)/libhysortk.o $(HOST_OBJ) $(CUDA_OBJ)

$(OBJ)/%.o: hysortk_b200/cxx/%.cpp include/*.hpp include/*.h
	@mkdir -p $(OBJ)
	$(CXX) $(CXXFLAGS) -c -o $@ $<

$(CUOBJ)/%.o: hysortk_b200/csrc/%.cu hysortk_b200/csrc/*.cuh include/hsk_capi.h
	@mkdir -p $(CUOBJ)
	$(NVCC) $(NVFLAGS) -c -o $@ $<

standalone: all
	$(CXX) $(CXXFLAGS) -c -o $(OBJ)/standalone.o hysortk_b200/cxx/main.cpp
	$(CXX) $(OPT) -fopenmp -o $(BIN) $(OBJ)/standalone.o $(OBJ)/libhysortk.o -L$(CUDA)/lib64 -lcudart -ldl -lpthread $(MPI_LIB)

# launcher for several ranks on one node with the bundled MPI stand-in (hysortk_b200/shim/mpi.h); with a real MPI use mpirun
mpirun:
	$(CXX) -O2 -std=c++17 -o hsk_mpirun hysortk_b200/shim/hsk_mpirun.cpp

clean:
	rm -rf $(OBJ) $(BIN) hsk_mpirun

.PHONY: all print lib standalone mpirun clean
