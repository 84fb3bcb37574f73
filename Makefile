# Build the C-ABI shared library for sm_100a (cross-compiles without a GPU).
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
             --expt-relaxed-constexpr -Xptxas -v
# probe.cu (descriptor probes, issue-rate microbenchmarks) is NOT product code: it goes into its own debug library
SRC       := $(filter-out lass_b200/csrc/probe.cu,$(wildcard lass_b200/csrc/*.cu))
HDR       := $(wildcard lass_b200/csrc/*.cuh) include/lass_b200.h
OBJ       := $(patsubst lass_b200/csrc/%.cu,build/%.o,$(SRC))
LIB       := lass_b200/_lib/liblass_b200.so

DBG_LIB   := lass_b200/_lib/liblass_b200_debug.so

all: $(LIB) $(DBG_LIB)

build/%.o: lass_b200/csrc/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	@mkdir -p lass_b200/_lib
	$(NVCC) -shared $(ARCH) -o $@ $(OBJ) -cudart static

$(DBG_LIB): build/probe.o build/common.o
	@mkdir -p lass_b200/_lib
	$(NVCC) -shared $(ARCH) -o $@ build/probe.o build/common.o -cudart static

# second library with the conv role profiler compiled in (tools/gpu_conv_timing.py; LASS_B200_LIB selects it)
PROF_OBJ  := $(patsubst lass_b200/csrc/%.cu,build/prof/%.o,$(SRC))
PROF_LIB  := lass_b200/_lib/liblass_b200_prof.so
build/prof/%.o: lass_b200/csrc/%.cu $(HDR)
	@mkdir -p build/prof
	$(NVCC) $(NVCCFLAGS) -DLASS_CONV_PROFILE -c $< -o $@ 2> build/prof/$*.ptxas.log || (cat build/prof/$*.ptxas.log; exit 1)
$(PROF_LIB): $(PROF_OBJ)
	@mkdir -p lass_b200/_lib
	$(NVCC) -shared $(ARCH) -o $@ $(PROF_OBJ) -cudart static
prof: $(PROF_LIB)

clean:
	rm -rf build $(LIB) $(DBG_LIB)

.PHONY: all clean prof
