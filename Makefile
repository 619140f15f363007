# Builds libswrt.so (sm_100a only) and the CPU oracle's C restatement.
NVCC      ?= nvcc
CC        := gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
# EXTRA=-DSWRT_DEV_TUNING enables the developer environment overrides of the dense kernel's geometry (tools/sweep.sh)
EXTRA     ?=
NVFLAGS   := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall $(EXTRA)
CSRC      := swraytracing_b200/csrc
OBJ       := build
LIB       := swraytracing_b200/libswrt.so
ORACLE    := oracle/build/liboracle.so

all: $(LIB) $(ORACLE)

$(OBJ)/%.o: $(CSRC)/%.cu $(CSRC)/swrt_internal.h $(CSRC)/spectral_common.cuh $(wildcard $(CSRC)/*.inc) include/swrt.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@

# LAGRANGE6 = the reference's arithmetic, operation for operation: no fused multiply-add contraction
$(OBJ)/lagrange_kernels.o: $(CSRC)/lagrange_kernels.cu $(CSRC)/swrt_internal.h include/swrt.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -fmad=false -c $< -o $@

$(LIB): $(OBJ)/spectral_kernels.o $(OBJ)/spectral_rk4_kernels.o $(OBJ)/lagrange_kernels.o $(OBJ)/nufft_kernels.o $(OBJ)/misc_kernels.o $(OBJ)/swrt_api.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -lcufft -ldl -Xlinker -rpath -Xlinker /usr/local/cuda/lib64

$(ORACLE): oracle/swrt_oracle.c
	@mkdir -p oracle/build
	$(CC) -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp -fPIC -shared -o $@ $< -lm

clean:
	rm -rf $(OBJ) $(LIB) oracle/build

.PHONY: all clean
