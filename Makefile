# Builds the in-tree native artefacts.  `make` = engine + Triton backend + oracle; nvcc cross-compiles
# sm_100a without a GPU.
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
CSRC      := hugectr_backend_b200/csrc
LIBDIR    := hugectr_backend_b200/lib
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wextra,-pthread -Iinclude
ENGINE_SRC := $(CSRC)/kernels.cu $(CSRC)/hpsx.cpp $(CSRC)/host_ps.cpp $(CSRC)/ps_config.cpp
ENGINE_HDR := $(wildcard $(CSRC)/*.h $(CSRC)/*.hpp include/*.h)

all: $(LIBDIR)/libhpsx.so oracle

$(LIBDIR)/libhpsx.so: $(ENGINE_SRC) $(ENGINE_HDR)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(ENGINE_SRC) -lpthread

oracle: oracle/libhps_oracle.so
oracle/libhps_oracle.so: oracle/hps_oracle.c oracle/hps_oracle.h
	$(CC) -O2 -std=c11 -Wall -Wextra -fPIC -shared -pthread -o $@ oracle/hps_oracle.c -lm

clean:
	rm -f $(LIBDIR)/*.so oracle/*.so

.PHONY: all oracle clean
