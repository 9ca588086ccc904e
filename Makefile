# Builds the product (libfastc_gpu.so: CUDA kernels + C ABI, sm_100a only), the
# C++ host layer that mirrors the reference's Core API (libFasTCCore.so + tc),
# and -- as test infrastructure -- the CPU oracle and the compiled reference.
NVCC      ?= nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: the parity paths replay the reference's SSE2 scalar float math,
# which never fuses a multiply with an add (SURVEY.md trap T4).
NVCCFLAGS := $(ARCH) -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC $(EXTRA_NVCCFLAGS)
CSRC      := fastc_b200/csrc
GPU_SRCS  := $(CSRC)/capi.cu $(CSRC)/dxt.cu $(CSRC)/etc1.cu $(CSRC)/bc7.cu $(CSRC)/decode.cu $(CSRC)/pvrtc.cu
GPU_OBJS  := $(GPU_SRCS:.cu=.o)
GPU_SO    := fastc_b200/libfastc_gpu.so

.PHONY: all gpu core oracle clean
all: gpu core oracle

gpu: $(GPU_SO)

$(CSRC)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/fastc_gpu.h
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(GPU_SO): $(GPU_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(GPU_OBJS) -lcudart

core: gpu
	@if [ -f fastc_b200/core/Makefile ]; then $(MAKE) -C fastc_b200/core; fi

oracle:
	$(MAKE) -s -C oracle all

clean:
	rm -f $(GPU_OBJS) $(GPU_SO)
	$(MAKE) -C oracle clean
