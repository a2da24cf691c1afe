# Builds the product library (libmprb.so, sm_100a) and the test oracles.
NVCC      ?= nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Impr_b200/shim -Impr_b200/csrc
CXXFLAGS  := -O2 -std=c++17 -fPIC -Impr_b200/shim -Impr_b200/csrc/host
BUILD     := build

CU_SRCS   := mpr_b200/csrc/kernels.cu mpr_b200/csrc/api.cu
CXX_SRCS  := mpr_b200/csrc/host/tree.cpp mpr_b200/csrc/host/tape_build.cpp
OBJS      := $(patsubst %.cu,$(BUILD)/%.o,$(CU_SRCS)) $(patsubst %.cpp,$(BUILD)/%.o,$(CXX_SRCS))

all: mpr_b200/libmprb.so

mpr_b200/libmprb.so: $(OBJS)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS)

$(BUILD)/%.o: %.cu $(wildcard mpr_b200/csrc/*.cuh) include/mprb.h
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(BUILD)/%.o: %.cpp
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) -c $< -o $@

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(BUILD) mpr_b200/libmprb.so

.PHONY: all oracle clean
