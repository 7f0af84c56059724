# Builds the C-ABI shared library (sm_100a only) and the oracle's C pieces.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -cudart static
SRC       := $(wildcard msau_b200/csrc/*.cu)
OBJ       := $(patsubst msau_b200/csrc/%.cu,build/%.o,$(SRC))
LIB       := msau_b200/lib/libmsau_b200.so

all: $(LIB)

build/%.o: msau_b200/csrc/%.cu $(wildcard msau_b200/csrc/*.cuh) include/msau_b200.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	@mkdir -p msau_b200/lib
	$(NVCC) $(ARCH) -shared -cudart static -o $@ $(OBJ)

clean:
	rm -rf build $(LIB)

.PHONY: all clean
