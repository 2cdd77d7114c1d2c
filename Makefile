# Builds the in-tree native artefacts.  `make` = engine + Triton backend + oracle; nvcc cross-compiles
# sm_100a without a GPU.
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
CSRC      := hugectr_backend_b200/csrc
LIBDIR    := hugectr_backend_b200/lib
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wextra,-pthread -Iinclude
ENGINE_SRC := $(CSRC)/kernels.cu $(CSRC)/shard_kernels.cu $(CSRC)/dense_mlp.cu $(CSRC)/hpsx.cpp $(CSRC)/shard_group.cpp $(CSRC)/peer_tier.cpp $(CSRC)/mlp_abi.cpp $(CSRC)/host_ps.cpp $(CSRC)/ps_config.cpp
ENGINE_HDR := $(wildcard $(CSRC)/*.h $(CSRC)/*.hpp $(CSRC)/*.cuh include/*.h)

all: $(LIBDIR)/libhpsx.so $(LIBDIR)/libtriton_hps.so tests/fake_triton/libfake_triton.so oracle

$(LIBDIR)/libhpsx.so: $(ENGINE_SRC) $(ENGINE_HDR)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(ENGINE_SRC) -lpthread

# The Triton `hps` backend shell: host C++ only, every device operation goes through libhpsx.so.
# Exports only TRITONBACKEND_* (version script); the TRITONSERVER_*/TRITONBACKEND_* imports are
# resolved by the process that loads it (tritonserver, or tests/fake_triton).
$(LIBDIR)/libtriton_hps.so: $(CSRC)/triton_hps.cpp $(CSRC)/ps_config.cpp $(CSRC)/libtriton_hps.ldscript $(ENGINE_HDR) $(LIBDIR)/libhpsx.so
	$(CXX) -O2 -std=c++17 -Wall -Wextra -fPIC -fvisibility=hidden -shared -pthread -Iinclude -o $@ \
	  $(CSRC)/triton_hps.cpp $(CSRC)/ps_config.cpp -L$(LIBDIR) -lhpsx \
	  -Wl,-rpath,'$$ORIGIN' -Wl,--version-script=$(CSRC)/libtriton_hps.ldscript

# Test infrastructure: a stand-in for the Triton server process (defines every symbol the backend imports).
tests/fake_triton/libfake_triton.so: tests/fake_triton/fake_triton.cpp include/triton_compat.h
	$(CXX) -O1 -g -std=c++17 -Wall -Wextra -fPIC -shared -pthread -Iinclude -o $@ tests/fake_triton/fake_triton.cpp -ldl

oracle: oracle/libhps_oracle.so
oracle/libhps_oracle.so: oracle/hps_oracle.c oracle/hps_oracle.h
	$(CC) -O2 -std=c11 -Wall -Wextra -fPIC -shared -pthread -o $@ oracle/hps_oracle.c -lm

clean:
	rm -f $(LIBDIR)/*.so oracle/*.so tests/fake_triton/*.so

.PHONY: all oracle clean
