# quantr-b200 build: the sm_100a C-ABI library, the CPU oracle and the host-emulation test harness.
#   make lib      -> quantr_b200/libqsv.so        (nvcc, sm_100a; the product)
#   make oracle   -> oracle/liboracle.so          (g++; test infrastructure)
#   make emu      -> tests/emu/libqsv_emu.so      (g++; test infrastructure)
NVCC ?= /usr/local/cuda/bin/nvcc
HOSTCXX := $(shell command -v /usr/bin/g++ || echo g++)
CSRC := quantr_b200/csrc
# amplitudes per thread = 2^QSV_REG_BITS (3 or 4); host scheduler, kernels and the emulation harness must agree
QSV_REG_BITS ?= 4
OBJ ?= build/obj$(QSV_REG_BITS)
LIB ?= quantr_b200/libqsv.so
QSV_OCC_NUM ?= 2
EXTRA_DEFS ?=
NVCCFLAGS := -DQSV_REG_BITS=$(QSV_REG_BITS) -DQSV_OCC_NUM=$(QSV_OCC_NUM) $(EXTRA_DEFS) -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unknown-pragmas -ccbin $(HOSTCXX)
HOSTFLAGS := -DQSV_REG_BITS=$(QSV_REG_BITS) -DQSV_OCC_NUM=$(QSV_OCC_NUM) $(EXTRA_DEFS) -O2 -std=c++17 -fPIC -Wall -Wextra -Wno-unknown-pragmas -pthread

HOST_SRCS := $(CSRC)/plan.cpp $(CSRC)/plan_api.cpp
HDRS := include/qsv.h $(wildcard $(CSRC)/*.h)
TILE_BITS := 0 10 11 12 13
PASS_OBJS := $(foreach t,$(TILE_BITS),$(OBJ)/pass_kernel_t$(t).o) $(OBJ)/pass_kernel_tma_t11.o $(OBJ)/pass_kernel_tma_t12.o
LIB_OBJS := $(OBJ)/plan.o $(OBJ)/plan_api.o $(OBJ)/kernels.o $(OBJ)/state_api.o $(OBJ)/shard.o $(PASS_OBJS)

all:
	$(MAKE) -j8 lib oracle emu

lib: $(LIB)
oracle:
	$(MAKE) -C oracle
emu: tests/emu/libqsv_emu.so

$(OBJ)/pass_kernel_tma_t%.o: $(CSRC)/pass_kernel_tma.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVCCFLAGS) -DQSV_TILE_BITS=$* -c -o $@ $<

$(OBJ)/pass_kernel_t%.o: $(CSRC)/pass_kernel.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVCCFLAGS) -DQSV_TILE_BITS=$* -c -o $@ $<

$(OBJ)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVCCFLAGS) -c -o $@ $<

$(OBJ)/%.o: $(CSRC)/%.cpp $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVCCFLAGS) -c -o $@ $<

$(LIB): $(LIB_OBJS)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(LIB_OBJS) -lcudart -ldl

tests/emu/libqsv_emu.so: tests/emu/qsv_emu.cpp $(HOST_SRCS) $(HDRS)
	$(HOSTCXX) $(HOSTFLAGS) -shared -Wl,-Bsymbolic -o $@ tests/emu/qsv_emu.cpp $(HOST_SRCS)

clean:
	rm -rf build quantr_b200/libqsv.so tests/emu/libqsv_emu.so
	$(MAKE) -C oracle clean

.PHONY: all lib oracle emu clean
