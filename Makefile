# Builds libcough_b200.so (sm_100a only) and nothing else.  `python -c "import __graft_entry__ as g; g.build()"`
# runs this.  nvcc cross-compiles without a GPU.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v
SRC_DIR   := cough_detector_b200/csrc
OUT_DIR   := cough_detector_b200/lib
SRCS      := $(SRC_DIR)/api.cu $(SRC_DIR)/feat_kernel.cu $(SRC_DIR)/cnn_fp32.cu $(SRC_DIR)/cnn_tc.cu $(SRC_DIR)/cnn_plain.cu $(SRC_DIR)/resample.cu $(SRC_DIR)/spec_augment.cu
OBJS      := $(patsubst $(SRC_DIR)/%.cu,build/%.o,$(SRCS))
LIB       := $(OUT_DIR)/libcough_b200.so

all: $(LIB)

build/%.o: $(SRC_DIR)/%.cu $(SRC_DIR)/cd_common.cuh $(SRC_DIR)/cd_tables.h include/cough_b200.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

$(LIB): $(OBJS)
	@mkdir -p $(OUT_DIR)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS)

# experimental builds of the feature kernel (occupancy / register trade-offs): lib/exp/lib<tag>.so, timed by
# `CD_LIB=... python tools/feat_time.py`.  EXP="tag1:-DFLAG=.. tag2:-DFLAG=.." (lib/ travels to the GPU box, build/ does not)
EXP ?= c2:-DCD_FEAT_CTAS=2 c3:-DCD_FEAT_CTAS=3
exp: $(OBJS)
	@mkdir -p build/exp $(OUT_DIR)/exp
	for e in $(EXP); do tag=$${e%%:*}; flags=$$(echo $${e#*:} | tr ',' ' '); \
	  $(NVCC) $(NVCCFLAGS) $$flags -c $(SRC_DIR)/feat_kernel.cu -o build/exp/feat_$$tag.o 2> build/exp/feat_$$tag.ptxas.log && \
	  $(NVCC) -shared $(ARCH) -o $(OUT_DIR)/exp/lib$$tag.so build/api.o build/exp/feat_$$tag.o build/cnn_fp32.o build/cnn_tc.o build/cnn_plain.o build/resample.o build/spec_augment.o; \
	done

# timing-only ablations of the feature kernel (profiles/r2_feat_ablation.txt): CD_LIB=.../libab<N>.so python tools/feat_time.py
ablations:
	$(MAKE) exp EXP="ab1:-DCD_ABLATE=1 ab2:-DCD_ABLATE=2 ab3:-DCD_ABLATE=3 ab4:-DCD_ABLATE=4"

# EXPERIMENTAL, never run on a GPU (see CD_MMA_ELECT in cnn_tc.cu, profiles/r2_mma_issue_bound.txt): the single-CTA conv kernels with a
# convergent MMA warp and elect.sync-predicated tcgen05 instructions -> lib/exp/libelect.so.  First thing to do with a GPU:
#   make elect && CD_LIB=cough_detector_b200/lib/exp/libelect.so python -m pytest tests -m gpu -x -q   (tests/conftest.py honours CD_LIB)
elect: $(OBJS)
	@mkdir -p build/exp $(OUT_DIR)/exp
	$(NVCC) $(NVCCFLAGS) -DCD_MMA_ELECT=1 -c $(SRC_DIR)/cnn_tc.cu -o build/exp/cnn_tc_elect.o 2> build/exp/cnn_tc_elect.ptxas.log || (cat build/exp/cnn_tc_elect.ptxas.log; false)
	$(NVCC) -shared $(ARCH) -o $(OUT_DIR)/exp/libelect.so build/api.o build/feat_kernel.o build/cnn_fp32.o build/exp/cnn_tc_elect.o build/cnn_plain.o build/resample.o build/spec_augment.o

# C++ callers of the C ABI (no Python, no torch)
examples: $(LIB)
	$(NVCC) -O2 -std=c++17 $(ARCH) -Iinclude examples/e2e_minimal.cu -o build/e2e_minimal -L$(OUT_DIR) -lcough_b200 -Xlinker -rpath,'$$ORIGIN/../$(OUT_DIR)'

clean:
	rm -rf build $(LIB)

.PHONY: all clean exp examples ablations elect
