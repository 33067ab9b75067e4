# Builds the product: libb200scan.so (sm_100a kernels + C ABI), libblammhost.so (C++ host model + C ABI for
# ctypes) and the blamm-b200 command line.  `make oracle` builds the test-only oracle.  Everything is built
# in-tree under blamm_b200/lib so that it travels to the GPU box.
NVCC     ?= nvcc
CXX      := /usr/bin/g++
LIBDIR   := blamm_b200/lib
NVFLAGS  := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 186
CXXFLAGS := -O2 -std=c++17 -fPIC -Wall -Wextra -ffp-contract=off
CSRC     := $(wildcard blamm_b200/csrc/*.cu blamm_b200/csrc/*.cuh) include/b200scan.h
HSRC     := blamm_b200/host/motifs.cpp blamm_b200/host/sequence.cpp
HHDR     := blamm_b200/host/host.h include/blamm_host.h include/b200scan.h

all: $(LIBDIR)/libb200scan.so $(LIBDIR)/libblammhost.so $(LIBDIR)/blamm-b200 $(LIBDIR)/i8_peak

$(LIBDIR)/libb200scan.so: $(CSRC)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -shared blamm_b200/csrc/b200scan.cu -o $@

# diagnostic build with clock64 probes in the filter kernel (tools/tc_trace.py); never used by the product
$(LIBDIR)/libb200scan_trace.so: $(CSRC)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -DB200_TRACE -shared blamm_b200/csrc/b200scan.cu -o $@

# diagnostic build with per-role cycle totals (tools/tc_phase.py); never used by the product
$(LIBDIR)/libb200scan_phase.so: $(CSRC)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -DB200_PHASE -shared blamm_b200/csrc/b200scan.cu -o $@

$(LIBDIR)/libblammhost.so: $(HSRC) blamm_b200/host/host_abi.cpp $(HHDR)
	@mkdir -p $(LIBDIR)
	$(CXX) $(CXXFLAGS) -shared $(HSRC) blamm_b200/host/host_abi.cpp -o $@

$(LIBDIR)/blamm-b200: $(HSRC) blamm_b200/host/cli.cpp $(HHDR) $(LIBDIR)/libb200scan.so
	$(CXX) $(CXXFLAGS) $(HSRC) blamm_b200/host/cli.cpp -o $@ -L$(LIBDIR) -lb200scan -Wl,-rpath,'$$ORIGIN' -lpthread

# dense tcgen05.mma peak of the INT8 / FP16 tensor pipe: the measured denominator of bench.py's roofline
$(LIBDIR)/i8_peak: tools/micro/i8_peak.cu blamm_b200/csrc/filter_tc.cuh blamm_b200/csrc/common.cuh
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) tools/micro/i8_peak.cu -o $@

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(LIBDIR)

.PHONY: all oracle clean
