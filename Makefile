# Builds the sm_100a library (the product) and the CPU oracle (the checker).
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: the f64 epilogues keep the reference's expression order (no FMA contraction).
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -cudart static
LIBDIR := distance_b200/_lib
LIB := $(LIBDIR)/libdistance_gpu.so
CSRC := distance_b200/csrc

all: lib cli oracle tools

lib: $(LIB)

$(LIB): $(CSRC)/dg_api.cu $(CSRC)/kernels.cuh $(CSRC)/tc_engine.cuh include/distance_gpu.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) $(PTXAS_V) -shared -o $@.tmp $(CSRC)/dg_api.cu && mv -f $@.tmp $@

# The `distance` command line (C++ host over the C ABI; the image has no Rust toolchain).
CXX ?= g++
BINDIR := distance_b200/_bin
CLI := $(BINDIR)/distance
HOST := $(CSRC)/host

cli: $(CLI)

$(CLI): $(HOST)/main.cpp $(HOST)/fasta.cpp $(HOST)/tsv.cpp $(HOST)/fasta.hpp $(HOST)/tsv.hpp $(HOST)/measures.hpp include/distance_gpu.h $(LIB)
	@mkdir -p $(BINDIR)
	$(CXX) -O2 -std=c++17 -ffp-contract=off -Wall -Wextra -pthread -o $@ $(HOST)/main.cpp $(HOST)/fasta.cpp $(HOST)/tsv.cpp \
		-L$(LIBDIR) -ldistance_gpu -Wl,-rpath,'$$ORIGIN/../_lib'

oracle:
	$(MAKE) -C oracle

# micro-benchmarks behind the roofline denominators and the host-side figures (profiles/*.json)
TOOLS := tools/ubench_int tools/ubench_tc tools/ubench_fp4 tools/ubench_pcie tools/ubench_init tools/tsv_bench
tools: $(TOOLS)
tools/ubench_%: tools/ubench_%.cu
	$(NVCC) $(ARCH) -O3 -lineinfo -o $@ $< -lcuda -ldl
tools/tsv_bench: tools/tsv_bench.cpp $(HOST)/tsv.cpp $(HOST)/fasta.cpp $(HOST)/tsv.hpp $(HOST)/fasta.hpp
	$(CXX) -O2 -std=c++17 -ffp-contract=off -pthread -o $@ tools/tsv_bench.cpp $(HOST)/tsv.cpp $(HOST)/fasta.cpp

clean:
	rm -rf $(LIBDIR) $(BINDIR) && $(MAKE) -C oracle clean

.PHONY: all lib cli oracle tools clean
