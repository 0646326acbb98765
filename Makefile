# Builds the C-ABI CUDA library (sm_100a) in-tree and the oracle's C restatement.
NVCC ?= nvcc
NVFLAGS = -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
CSRC = meanfieldvargp_b200/csrc
LIB = meanfieldvargp_b200/lib/libmfvgp.so

all: $(LIB)

$(LIB): $(wildcard $(CSRC)/*.cu $(CSRC)/*.cuh $(CSRC)/*.h) include/mfvgp.h
	mkdir -p meanfieldvargp_b200/lib
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/mfvgp.cu

clean:
	rm -f $(LIB)
