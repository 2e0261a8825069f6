# Builds the sm_100a library (the product) and the CPU oracle (the checker).
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: the f64 epilogues keep the reference's expression order (no FMA contraction).
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -cudart static
LIBDIR := distance_b200/_lib
LIB := $(LIBDIR)/libdistance_gpu.so
CSRC := distance_b200/csrc

all: lib oracle

lib: $(LIB)

$(LIB): $(CSRC)/dg_api.cu $(CSRC)/kernels.cuh include/distance_gpu.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) $(PTXAS_V) -shared -o $@ $(CSRC)/dg_api.cu

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(LIBDIR) && $(MAKE) -C oracle clean

.PHONY: all lib oracle clean
