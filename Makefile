# Builds the CUDA library (the product) and the CPU oracle (test infrastructure).
#   make            -> ctsm_b200/lib/libctsm_b200.so  +  oracle/liboracle.so
# The device code is FP64 and compiled with -fmad=false: the reference is built
# with -ffp-contract=off (ccs_config/machines/cmake_macros/gnu.cmake:51-53), so
# multiply-add contraction would change results.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo -fmad=false $(ARCH) -Xcompiler -fPIC -ccbin /usr/bin/g++ $(EXTRA)
CSRC      := ctsm_b200/csrc
CU        := $(wildcard $(CSRC)/*.cu)
OBJ       := $(CU:.cu=.o)
HDR       := $(wildcard $(CSRC)/*.cuh) include/ctsm_b200.h include/ctsm_b200_fields.def include/ctsm_b200_defaults.h
LIB       ?= ctsm_b200/lib/libctsm_b200.so

all: $(LIB) oracle

$(CSRC)/%.o: $(CSRC)/%.cu $(HDR)
	$(NVCC) $(NVCCFLAGS) $(PTXAS) -c $< -o $@

$(LIB): $(OBJ)
	mkdir -p ctsm_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -ccbin /usr/bin/g++

oracle:
	$(MAKE) -C oracle liboracle.so

clean:
	rm -f $(OBJ) $(LIB); $(MAKE) -C oracle clean

.PHONY: all oracle clean
