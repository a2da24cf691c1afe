# Builds the product library (libmprb.so, sm_100a) and the test oracles.
NVCC      ?= nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Impr_b200/shim -Impr_b200/csrc
CXXFLAGS  := -O2 -std=c++17 -fPIC -Impr_b200/shim -Impr_b200/inc -Impr_b200/csrc/host
BUILD     := build

CU_SRCS   := mpr_b200/csrc/kernels.cu mpr_b200/csrc/api.cu
CXX_SRCS  := mpr_b200/csrc/host/tree.cpp mpr_b200/csrc/host/tape_build.cpp mpr_b200/csrc/host/cxx_api.cpp
OBJS      := $(patsubst %.cu,$(BUILD)/%.o,$(CU_SRCS)) $(patsubst %.cpp,$(BUILD)/%.o,$(CXX_SRCS))

all: mpr_b200/libmprb.so drivers

mpr_b200/libmprb.so: $(OBJS)
	$(NVCC) -shared $(ARCH) -Xlinker -Bsymbolic -o $@ $(OBJS)

$(BUILD)/%.o: %.cu $(wildcard mpr_b200/csrc/*.cuh) include/mprb.h
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(BUILD)/%.o: %.cpp
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) -c $< -o $@

oracle:
	$(MAKE) -C oracle

# The reference's own benchmark drivers, compiled UNCHANGED from where they lie under $(REF)
# against mpr_b200/inc + mpr_b200/shim and linked with libmprb.so (drop-in check).
REF ?= /root/reference
DRIVERS := render_2d_table render_3d_table print_tape_table tape_building_time circle tape_shortening
drivers: mpr_b200/libmprb.so
	@if [ -d $(REF)/benchmark ]; then mkdir -p $(BUILD)/drivers && for d in $(DRIVERS); do \
	  $(CXX) -O2 -std=c++17 -Impr_b200/inc -Impr_b200/shim -I$(REF)/benchmark \
	    $(REF)/benchmark/$$d.cpp $(REF)/benchmark/stats.cpp -Lmpr_b200 -lmprb \
	    -Wl,-rpath,'$$ORIGIN/../../mpr_b200' -o $(BUILD)/drivers/$$d || exit 1; done; \
	else echo "reference sources not present; keeping prebuilt $(BUILD)/drivers (if any)"; fi

clean:
	rm -rf $(BUILD) mpr_b200/libmprb.so

.PHONY: all oracle drivers clean
