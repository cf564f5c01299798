# quantr-b200 build: the sm_100a C-ABI library, the CPU oracle and the host-emulation test harness.
#   make lib      -> quantr_b200/libqsv.so        (nvcc, sm_100a; the product)
#   make oracle   -> oracle/liboracle.so          (g++; test infrastructure)
#   make emu      -> tests/emu/libqsv_emu.so      (g++; test infrastructure)
NVCC ?= /usr/local/cuda/bin/nvcc
HOSTCXX := $(shell command -v /usr/bin/g++ || echo g++)
CSRC := quantr_b200/csrc
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unknown-pragmas -ccbin $(HOSTCXX)
HOSTFLAGS := -O2 -std=c++17 -fPIC -Wall -Wextra -Wno-unknown-pragmas -pthread

HOST_SRCS := $(CSRC)/plan.cpp $(CSRC)/plan_api.cpp
CUDA_SRCS := $(CSRC)/kernels.cu $(CSRC)/state_api.cu $(CSRC)/shard.cpp
HDRS := include/qsv.h $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh)

all: lib oracle emu

lib: quantr_b200/libqsv.so
oracle:
	$(MAKE) -C oracle
emu: tests/emu/libqsv_emu.so

quantr_b200/libqsv.so: $(HOST_SRCS) $(CUDA_SRCS) $(HDRS)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(HOST_SRCS) $(CUDA_SRCS) -lcudart -ldl

tests/emu/libqsv_emu.so: tests/emu/qsv_emu.cpp $(HOST_SRCS) $(HDRS)
	$(HOSTCXX) $(HOSTFLAGS) -shared -Wl,-Bsymbolic -o $@ tests/emu/qsv_emu.cpp $(HOST_SRCS)

clean:
	rm -f quantr_b200/libqsv.so tests/emu/libqsv_emu.so
	$(MAKE) -C oracle clean

.PHONY: all lib oracle emu clean
