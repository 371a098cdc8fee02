# Builds libbsr.so (sm_100a only) and nothing else; `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
CSRC := blindshadowremoval_b200/csrc
OUT  := blindshadowremoval_b200/libbsr.so
HDRS := $(wildcard $(CSRC)/*.cuh) include/bsr.h

$(OUT): $(CSRC)/bsr_api.cu $(HDRS)
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared -Xptxas -v \
	    --expt-relaxed-constexpr $(EXTRA) -o $@ $(CSRC)/bsr_api.cu 2> build.log || (cat build.log; exit 1)
	@grep -E "error|warning" build.log | grep -v "ptxas info" | head -20 || true

# A/B build with bfloat16 activation storage (error measurements only; select it with BSR_LIB=<path>)
bf16: $(CSRC)/bsr_api.cu $(HDRS)
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared --expt-relaxed-constexpr -DBSR_ACT_BF16 \
	    -o blindshadowremoval_b200/libbsr_bf16.so $(CSRC)/bsr_api.cu 2> build_bf16.log || (cat build_bf16.log; exit 1)

# role-timer build of the same ABI for tools/role_timers.py (select it with BSR_LIB=<path>)
timers: $(CSRC)/bsr_api.cu $(HDRS)
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared --expt-relaxed-constexpr -DBSR_ROLE_TIMERS \
	    -o blindshadowremoval_b200/libbsr_timers.so $(CSRC)/bsr_api.cu 2> build_timers.log || (cat build_timers.log; exit 1)

# role-timer build for tools/role_timers.py:  make clean && make EXTRA=-DBSR_ROLE_TIMERS
clean:
	rm -f $(OUT) blindshadowremoval_b200/libbsr_bf16.so blindshadowremoval_b200/libbsr_timers.so build.log build_bf16.log build_timers.log
