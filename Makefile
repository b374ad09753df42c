# Builds libcorrla_b200.so (sm_100a only) and the standalone tools.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Iinclude -Icorrla_rs_b200/csrc
CSRC      := corrla_rs_b200/csrc
OBJDIR    := build
LIB       := corrla_rs_b200/lib/libcorrla_b200.so
OBJS      := $(OBJDIR)/skinny_gemm.o $(OBJDIR)/small_kernels.o $(OBJDIR)/engine.o $(OBJDIR)/comm.o $(OBJDIR)/hostcopy.o $(OBJDIR)/rom.o $(OBJDIR)/gradients.o $(OBJDIR)/jacobi_cluster.o $(OBJDIR)/jacobi_ring.o $(OBJDIR)/fused_small.o

all: $(LIB)

$(OBJDIR)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/corrla_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p $(dir $(LIB))
	$(NVCC) $(ARCH) -shared -cudart static -o $@ $(OBJS) -ldl -lpthread

tools: tools/test_gemm tools/peaks2
tools/test_gemm: tools/test_gemm.cu $(CSRC)/skinny_gemm.cu $(CSRC)/skinny_gemm.cuh $(CSRC)/ptx.cuh
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -I$(CSRC) -o $@ tools/test_gemm.cu $(CSRC)/skinny_gemm.cu
tools/peaks2: tools/peaks2.cu
	$(NVCC) $(ARCH) -O3 -o $@ $<

clean:
	rm -rf $(OBJDIR) $(LIB) tools/test_gemm tools/peaks tools/peaks2
.PHONY: all tools clean
