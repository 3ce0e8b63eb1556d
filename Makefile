# Top-level build: the CUDA library (sm_100a only), the C host programs, the oracle.
NVCC      ?= nvcc
CC        ?= gcc
ARCH       = -gencode arch=compute_100a,code=sm_100a
NVFLAGS    = $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -cudart static
CFLAGS     = -std=c11 -Wall -Wextra -O2
LIBDIR     = tiny_mc_b200/lib
BINDIR     = tiny_mc_b200/bin
CSRC       = tiny_mc_b200/csrc
HOST       = tiny_mc_b200/host
LIB        = $(LIBDIR)/libtinymc_b200.so
# -D overrides for the host programs, e.g. make headless DEFS="-DPHOTONS=67108864 -DSEED=7"
DEFS      ?=

all: lib host oracle

lib: $(LIB) $(LIBDIR)/libtmc_report.so $(BINDIR)/tmc_microbench

$(LIB): $(CSRC)/tmc_api.cu $(CSRC)/walk_kernel.cuh $(CSRC)/philox.cuh include/tiny_mc_b200.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/tmc_api.cu -ldl

# the printout formatter alone, so tests can feed it reference tallies (no GPU needed)
$(LIBDIR)/libtmc_report.so: $(HOST)/report.c $(HOST)/report.h
	@mkdir -p $(LIBDIR)
	$(CC) $(CFLAGS) -fPIC -shared -o $@ $(HOST)/report.c -lm

$(BINDIR)/tmc_microbench: $(CSRC)/microbench.cu
	@mkdir -p $(BINDIR)
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -o $@ $<

host: $(BINDIR)/headless $(BINDIR)/frames $(LIBDIR)/libphoton_compat.so configs

# the named configurations of BASELINE.json as host programs (compile-time macros, like the reference)
CONFIG2 = -DPHOTONS=67108864ULL -DSEED=24301
CONFIG3 = -DPHOTONS=4294967296ULL -DSEED=24301
CONFIG4 = -DPHOTONS=1073741824ULL -DSEED=24301 -DMU_A=0.1f -DMU_S=100.0f
CONFIG5 = -DPHOTONS=1073741824ULL -DSEED=24301 -DSHELLS=16384 -DMICRONS_PER_SHELL=5
configs: $(BINDIR)/headless_config2 $(BINDIR)/headless_config3 $(BINDIR)/headless_config4 $(BINDIR)/headless_config5

$(BINDIR)/headless_config%: $(HOST)/tiny_mc.c $(HOST)/report.c $(HOST)/wtime.c $(LIB)
	@mkdir -p $(BINDIR)
	$(CC) $(CFLAGS) $(CONFIG$*) -Iinclude -I$(HOST) -o $@ $(HOST)/tiny_mc.c $(HOST)/report.c $(HOST)/wtime.c \
	    -L$(LIBDIR) -ltinymc_b200 -Wl,-rpath,'$$ORIGIN/../lib' -lm

$(BINDIR)/headless: $(HOST)/tiny_mc.c $(HOST)/report.c $(HOST)/wtime.c $(LIB)
	@mkdir -p $(BINDIR)
	$(CC) $(CFLAGS) $(DEFS) -Iinclude -I$(HOST) -o $@ $(HOST)/tiny_mc.c $(HOST)/report.c $(HOST)/wtime.c \
	    -L$(LIBDIR) -ltinymc_b200 -Wl,-rpath,'$$ORIGIN/../lib' -lm

# the viewer's incremental use of the tallies (reference cg_mc.c:71-87) without the OpenGL part
$(BINDIR)/frames: $(HOST)/frames.c $(LIB)
	@mkdir -p $(BINDIR)
	$(CC) $(CFLAGS) -DSEED=4242 $(DEFS) -Iinclude -I$(HOST) -o $@ $(HOST)/frames.c -L$(LIBDIR) -ltinymc_b200 -Wl,-rpath,'$$ORIGIN/../lib'

$(LIBDIR)/libphoton_compat.so: $(HOST)/photon_compat.c $(LIB)
	$(CC) $(CFLAGS) $(DEFS) -Iinclude -I$(HOST) -fPIC -shared -o $@ $(HOST)/photon_compat.c \
	    -L$(LIBDIR) -ltinymc_b200 -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle all
	@if [ -d /root/reference ]; then $(MAKE) -C oracle ref; fi

# CPU-side tests here; the GPU-side ones need a B200 (python -m pytest tests -m gpu)
test: all
	python -m pytest tests -q -m "not gpu"

clean:
	rm -rf $(LIBDIR) $(BINDIR)
	$(MAKE) -C oracle clean

.PHONY: all lib host configs oracle test clean
