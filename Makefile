# Builds the product library and the C++ drop-in example without Python (the same commands biorseo_b200/build.py and
# tests/test_dropin_cpp.py run). nvcc cross-compiles for sm_100a without a GPU.
#
#   make lib      -> biorseo_b200/lib/libbiorseo_b200.so   (CUDA kernels + C ABI + C++ host side)
#   make dropin   -> tests/cpp/_build/dropin               (the patched-MOIP stand-in of INTEGRATION.md)
#   make oracle   -> oracle/_build/biorseo_port (+ oracle/_ref/biorseo_ref where /root/reference exists): TEST infrastructure
NVCC     ?= nvcc
CXX      ?= g++
LIB      := biorseo_b200/lib/libbiorseo_b200.so
CUDA_SRC := $(addprefix biorseo_b200/csrc/,bss_kernels.cu bss_mask.cu bss_index.cu bss_probe.cu bss_exchange.cu bss_sharded.cu bss_capi.cu bss_blob.cpp)
HOST_SRC := $(addprefix biorseo_b200/host/,Motif.cpp Pool.cpp mini_json.cpp module_library.cpp site_search.cpp fasta.cpp bsh_capi.cpp)
HEADERS  := $(wildcard biorseo_b200/csrc/*.cuh biorseo_b200/csrc/*.h biorseo_b200/host/*.h include/*.h)

.PHONY: lib dropin oracle clean
lib: $(LIB)

$(LIB): $(CUDA_SRC) $(HOST_SRC) $(HEADERS)
	@mkdir -p $(dir $@)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC,-Wall,-Wno-unused-function \
	    -Xptxas -O3 -I include -I biorseo_b200/host -I biorseo_b200/csrc -o $@ $(CUDA_SRC) $(HOST_SRC)

dropin: tests/cpp/_build/dropin

tests/cpp/_build/dropin: tests/cpp/dropin_main.cpp $(LIB)
	@mkdir -p $(dir $@)
	$(CXX) -std=c++17 -O1 $< -I include -I biorseo_b200/host -L $(dir $(LIB)) -lbiorseo_b200 -Wl,-rpath,$(abspath $(dir $(LIB))) -o $@

oracle:
	$(MAKE) -C oracle port $(if $(wildcard /root/reference/cppsrc),ref,)

clean:
	rm -rf biorseo_b200/lib tests/cpp/_build
