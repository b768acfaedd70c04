# Build everything in-tree (artefacts are git-ignored but travel to the GPU box with gpurun).
#   make            -> peregrine_b200/libpgb200.so, bin/shmr_index, bin/shmr_overlap, build/simreads
#   make oracle     -> oracle/_build/liboracle.so (+ oracle/_ref/* when /root/reference is present)
#   make hostsim    -> build/hostsim (CPU simulator of kernel logic; development/test tool)
NVCC ?= nvcc
NVFLAGS = -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --extended-lambda -Xcompiler -fPIC
CSRC = peregrine_b200/csrc
LIB = peregrine_b200/libpgb200.so

all: $(LIB) bin/shmr_index bin/shmr_overlap bin/shmr_dedup bin/shmr_mkseqdb bin/shmr_map build/simreads

$(LIB): $(wildcard $(CSRC)/*.cu $(CSRC)/*.cuh $(CSRC)/*.inc $(CSRC)/*.hpp) include/pgb200.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/pgb200.cu -lz

bin/shmr_index: cli/shmr_index.c $(LIB)
	mkdir -p bin
	gcc -O2 -o $@ cli/shmr_index.c -Lperegrine_b200 -lpgb200 -Wl,-rpath,'$$ORIGIN/../peregrine_b200'
bin/shmr_overlap: cli/shmr_overlap.c $(LIB)
	mkdir -p bin
	gcc -O2 -o $@ cli/shmr_overlap.c -Lperegrine_b200 -lpgb200 -Wl,-rpath,'$$ORIGIN/../peregrine_b200'

bin/shmr_dedup: cli/shmr_dedup.c $(LIB)
	mkdir -p bin
	gcc -O2 -o $@ cli/shmr_dedup.c -Lperegrine_b200 -lpgb200 -Wl,-rpath,'$$ORIGIN/../peregrine_b200'

bin/shmr_mkseqdb: cli/shmr_mkseqdb.c $(LIB)
	mkdir -p bin
	gcc -O2 -o $@ cli/shmr_mkseqdb.c -Lperegrine_b200 -lpgb200 -Wl,-rpath,'$$ORIGIN/../peregrine_b200'

bin/shmr_map: cli/shmr_map.c $(LIB)
	mkdir -p bin
	gcc -O2 -o $@ cli/shmr_map.c -Lperegrine_b200 -lpgb200 -Wl,-rpath,'$$ORIGIN/../peregrine_b200'

build/simreads: tools/simreads.c
	mkdir -p build
	gcc -O3 -fopenmp -o $@ tools/simreads.c -lm

hostsim: build/hostsim
build/hostsim: tests/hostsim/hostsim.cpp $(CSRC)/shimmer_core.cuh $(CSRC)/ovlp_match_lean_body.inc $(CSRC)/sketch_tile.cuh $(CSRC)/khash_small.cuh $(CSRC)/dedup.cuh $(CSRC)/fasta_reader.hpp $(CSRC)/host_util.hpp
	mkdir -p build
	g++ -O2 -std=c++17 -x c++ -o $@ tests/hostsim/hostsim.cpp -ldl -lz

oracle:
	$(MAKE) -C oracle all
	@if [ -d /root/reference/src ]; then $(MAKE) -C oracle ref; else echo "no /root/reference: keeping prebuilt oracle/_ref"; fi

clean:
	rm -rf $(LIB) bin build
.PHONY: all oracle hostsim clean
