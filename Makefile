# Builds libgdn_b200.so (C ABI, sm_100a only) in-tree.  nvcc cross-compiles without a GPU.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr
SRC := $(wildcard gdn_pytorch_b200/csrc/*.cu)
OBJ := $(SRC:.cu=.o)
LIB := gdn_pytorch_b200/libgdn_b200.so

all: $(LIB)

%.o: %.cu $(wildcard gdn_pytorch_b200/csrc/*.cuh) include/gdn_b200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -cudart static

clean:
	rm -f $(OBJ) $(LIB)
.PHONY: all clean
