# Builds the sm_100a C-ABI library and the CPU oracle.  No GPU needed (nvcc cross-compiles).
NVCC      ?= /usr/local/cuda/bin/nvcc
CC        := gcc
PKG       := hs-pose_b200
CSRC      := $(PKG)/csrc
LIBDIR    := $(PKG)/lib
LIB       := $(LIBDIR)/libhspose_b200.so
CU        := $(wildcard $(CSRC)/*.cu)
OBJ       := $(patsubst $(CSRC)/%.cu,build/%.o,$(CU))
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
             -Xcompiler -fPIC -Xptxas -v -Iinclude
ORACLE    := oracle/_build/libhsp_oracle.so

all: $(LIB) $(ORACLE)

build/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/hspose_b200.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	@mkdir -p $(LIBDIR)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $(OBJ)

$(ORACLE): oracle/hsp_oracle.c
	@mkdir -p oracle/_build
	$(CC) -O2 -std=c11 -fPIC -shared -mfma -ffp-contract=off -o $@ $< -lm

clean:
	rm -rf build $(LIBDIR) oracle/_build

.PHONY: all clean
