# Build the C-ABI shared library for sm_100a (cross-compiles without a GPU).
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
             --expt-relaxed-constexpr -Xptxas -v
SRC       := $(wildcard lass_b200/csrc/*.cu)
HDR       := $(wildcard lass_b200/csrc/*.cuh) include/lass_b200.h
OBJ       := $(patsubst lass_b200/csrc/%.cu,build/%.o,$(SRC))
LIB       := lass_b200/_lib/liblass_b200.so

all: $(LIB)

build/%.o: lass_b200/csrc/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	@mkdir -p lass_b200/_lib
	$(NVCC) -shared $(ARCH) -o $@ $(OBJ) -cudart static

clean:
	rm -rf build $(LIB)

.PHONY: all clean
