# Builds the C-ABI shared library for sm_100a (cross-compiles without a GPU).
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall
SRC := $(wildcard demovlp_b200/csrc/*.cu)
OBJ := $(patsubst demovlp_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := demovlp_b200/lib/libdemovlp_b200.so

all: $(LIB)

build/%.o: demovlp_b200/csrc/%.cu $(wildcard demovlp_b200/csrc/*.cuh) include/demovlp_b200.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	@mkdir -p demovlp_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart_static -lpthread -ldl -lrt

clean:
	rm -rf build $(LIB)

.PHONY: all clean
