# oracle/ref_ptx.mk -- TEST INFRASTRUCTURE ONLY.
#
# Compiles the reference's OWN CUDA sources of the path -- unmodified, from where they lie under /root/reference -- to PTX
# with this image's nvcc (default -fmad=true, as the reference builds them).  Nothing is run on a GPU: the PTX is executed
# on the CPU by oracle/ptx_interp.py (exact binary32 arithmetic) and compared with oracle-G bit for bit
# (tests/test_oracle_ptx.py).  This pins what the compiler does to the reference's expressions (which products are fused
# into an fma, which are rounded on their own) by compilation instead of by reasoning.
# Outputs go to oracle/_ref/ptx/ only (git-ignored; travels with gpurun like the rest of oracle/_ref).
#
#   360_stitcher/resize.cu                                   -> app_resize.ptx       (kernel `resize`, custom_resize)
#   sources/modules/stitching/src/cuda/multiband_blend.cu    -> multiband_blend.ptx  (addSrcWeight / normalizeUsingWeight kernels)
#   sources/modules/cudawarping/src/cuda/pyr_down.cu, pyr_up.cu -> pyr_down.ptx, pyr_up.ptx
#   sources/modules/cudawarping/src/cuda/resize.cu           -> resize.ptx           (cuda::resize: seam-scale frames, seam masks)
#   sources/modules/core/src/cuda/gpu_mat.cu                 -> gpu_mat.ptx          (GpuMat::convertTo with a scale: the gain)
#   sources/modules/cudaarithm/src/cuda/copy_make_border.cu  -> copy_make_border.ptx (cuda::copyMakeBorder: the REFLECT border of feed_online)
#   sources/modules/stitching/src/cuda/build_warp_maps.cu    -> build_warp_maps.ptx  (inspected for its fma pattern only: sinf / cosf)
#   sources/modules/cudawarping/src/cuda/remap.cu            -> remap.ptx            (cuda::remap; its 1- / 4-channel texture paths are stubbed, ptx_stubs.h)

REF  := /root/reference
MOD  := $(REF)/sources/modules
OUT  := _ref/ptx
GEN  := _ref/gen
NVCC ?= /usr/local/cuda/bin/nvcc
# any virtual architecture gives the same floating-point code (checked: sm_61, sm_75, sm_90, sm_100a); sm_100a is what this image targets
# HAVE_OPENCV_CUDEV / __OPENCV_BUILD: what opencv_modules.hpp and the module build define for the cudev-based files
FLAGS := -w -arch=sm_100a -ptx -DHAVE_OPENCV_CUDEV -D__OPENCV_BUILD=1 -include ptx_stubs.h -I$(GEN) -I$(MOD)/core/include -I$(MOD)/cudaarithm/include \
         -I$(MOD)/cudawarping/include -I$(MOD)/cudev/include

all: $(OUT)/app_resize.ptx $(OUT)/multiband_blend.ptx $(OUT)/pyr_down.ptx $(OUT)/pyr_up.ptx $(OUT)/build_warp_maps.ptx $(OUT)/remap.ptx \
     $(OUT)/resize.ptx $(OUT)/gpu_mat.ptx $(OUT)/copy_make_border.ptx

$(OUT)/app_resize.ptx: $(REF)/360_stitcher/resize.cu ptx_stubs.h
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) $< -o $@
$(OUT)/multiband_blend.ptx: $(MOD)/stitching/src/cuda/multiband_blend.cu ptx_stubs.h
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) $< -o $@
$(OUT)/build_warp_maps.ptx: $(MOD)/stitching/src/cuda/build_warp_maps.cu ptx_stubs.h
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) $< -o $@
$(OUT)/pyr_down.ptx: $(MOD)/cudawarping/src/cuda/pyr_down.cu ptx_stubs.h
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) $< -o $@
$(OUT)/pyr_up.ptx: $(MOD)/cudawarping/src/cuda/pyr_up.cu ptx_stubs.h
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) $< -o $@
$(OUT)/remap.ptx: $(MOD)/cudawarping/src/cuda/remap.cu ptx_stubs.h
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) $< -o $@
$(OUT)/resize.ptx: $(MOD)/cudawarping/src/cuda/resize.cu ptx_stubs.h
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) $< -o $@
$(OUT)/gpu_mat.ptx: $(MOD)/core/src/cuda/gpu_mat.cu ptx_stubs.h
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) $< -o $@
$(OUT)/copy_make_border.ptx: $(MOD)/cudaarithm/src/cuda/copy_make_border.cu ptx_stubs.h
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) $< -o $@
