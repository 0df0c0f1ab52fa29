# oracle/ref.mk -- TEST INFRASTRUCTURE ONLY.
#
# Compiles the reference's OWN vendored OpenCV 3.4.0 CPU sources (core, imgproc and the unmodified
# stitching sources warpers.cpp / seam_finders.cpp / util.cpp / camera.cpp) directly from where they
# lie under /root/reference/sources -- nothing is copied into this repository, and the reference's
# build system (CMake) is not run -- plus oracle/ref_shim.cpp (a C-ABI veneer + the restated CPU
# branch of MultiBandBlender), into oracle/_ref/libvsref.so.
#
# The few header files OpenCV's CMake would have generated (cvconfig.h, cv_cpu_config.h,
# custom_hal.hpp, opencv_modules.hpp, version_string.inc, *.simd_declarations.hpp,
# opencl_kernels_*.hpp) are written below as minimal stubs into oracle/_ref/gen: CPU baseline
# SSE2/SSE3 only (no dispatched AVX variants), pthreads parallel_for_, no OpenCL/IPP/CUDA/TBB.
#
# sources/modules/stitching/src/blenders.cpp (the authors' edited MultiBandBlender) cannot be part of
# this build: it includes highgui and uses cv::cuda members unconditionally; its CPU
# branch (:463-695, :835-941, :954-1050) is restated in ref_shim.cpp on top of the vendored primitives.

REF   := /root/reference/sources
MOD   := $(REF)/modules
OUT   := _ref
GEN   := $(OUT)/gen
OBJ   := $(OUT)/obj
CXX   := $(shell command -v /usr/bin/g++ || echo g++)
CC    ?= gcc

CORE_SKIP := convert.avx2.cpp convert.fp16.cpp convert.sse4_1.cpp
CORE_SRCS := $(filter-out $(addprefix $(MOD)/core/src/,$(CORE_SKIP)),$(wildcard $(MOD)/core/src/*.cpp)) \
             $(wildcard $(MOD)/core/src/utils/*.cpp)
IMG_SKIP  := imgwarp.avx2.cpp imgwarp.sse4_1.cpp resize.avx2.cpp resize.sse4_1.cpp undistort.avx2.cpp filter.avx2.cpp corner.avx.cpp accum.cpp accum.dispatch.cpp
IMG_SRCS  := $(filter-out $(addprefix $(MOD)/imgproc/src/,$(IMG_SKIP)),$(wildcard $(MOD)/imgproc/src/*.cpp))
ST_SRCS   := $(addprefix $(MOD)/stitching/src/,warpers.cpp seam_finders.cpp util.cpp camera.cpp exposure_compensate.cpp)

CORE_OBJS := $(patsubst $(MOD)/core/src/%.cpp,$(OBJ)/core/%.o,$(CORE_SRCS))
IMG_OBJS  := $(patsubst $(MOD)/imgproc/src/%.cpp,$(OBJ)/imgproc/%.o,$(IMG_SRCS))
ST_OBJS   := $(patsubst $(MOD)/stitching/src/%.cpp,$(OBJ)/stitching/%.o,$(ST_SRCS))

INC := -I$(GEN) -I$(MOD)/core/include -I$(MOD)/imgproc/include -I$(MOD)/stitching/include \
       -I$(MOD)/features2d/include -I$(MOD)/calib3d/include -I$(MOD)/flann/include -I$(REF)/3rdparty/include
CXXFLAGS := -O2 -fPIC -std=c++11 -w -msse3 -pthread -D__OPENCV_BUILD=1 -DCVAPI_EXPORTS -DNDEBUG -fvisibility=default $(INC)

GEN_FILES := $(GEN)/cvconfig.h $(GEN)/cv_cpu_config.h $(GEN)/custom_hal.hpp $(GEN)/opencv2/opencv_modules.hpp \
             $(GEN)/version_string.inc $(GEN)/stat.simd_declarations.hpp $(GEN)/mathfuncs_core.simd_declarations.hpp \
             $(GEN)/accum.simd_declarations.hpp $(GEN)/opencl_kernels_core.hpp $(GEN)/opencl_kernels_imgproc.hpp \
             $(GEN)/opencl_kernels_stitching.hpp

.PHONY: all
all: $(OUT)/libvsref.so

$(GEN)/.stamp:
	@mkdir -p $(GEN)/opencv2 $(OBJ)/core/utils $(OBJ)/imgproc $(OBJ)/stitching
	@printf '#ifndef OPENCV_CVCONFIG_H_INCLUDED\n#define OPENCV_CVCONFIG_H_INCLUDED\n#define CV_ENABLE_INTRINSICS\n#define HAVE_PTHREAD\n#define HAVE_PTHREADS_PF\n#endif\n' > $(GEN)/cvconfig.h
	@printf '#define CV_CPU_COMPILE_SSE 1\n#define CV_CPU_BASELINE_COMPILE_SSE 1\n#define CV_CPU_COMPILE_SSE2 1\n#define CV_CPU_BASELINE_COMPILE_SSE2 1\n#define CV_CPU_COMPILE_SSE3 1\n#define CV_CPU_BASELINE_COMPILE_SSE3 1\n#define CV_CPU_BASELINE_FEATURES 0, CV_CPU_SSE, CV_CPU_SSE2, CV_CPU_SSE3\n' > $(GEN)/cv_cpu_config.h
	@printf '#ifndef _CUSTOM_HAL_INCLUDED_\n#define _CUSTOM_HAL_INCLUDED_\n#endif\n' > $(GEN)/custom_hal.hpp
	@printf '#define HAVE_OPENCV_CORE\n#define HAVE_OPENCV_IMGPROC\n#define HAVE_OPENCV_STITCHING\n' > $(GEN)/opencv2/opencv_modules.hpp
	@printf '"OpenCV 3.4.0 (vendored by ultravideo/video-stitcher), core+imgproc CPU subset built by oracle/ref.mk\\n"\n' > $(GEN)/version_string.inc
	@printf '#define CV_CPU_SIMD_FILENAME "stat.simd.hpp"\n#define CV_CPU_DISPATCH_MODES_ALL BASELINE\n' > $(GEN)/stat.simd_declarations.hpp
	@printf '#define CV_CPU_SIMD_FILENAME "mathfuncs_core.simd.hpp"\n#define CV_CPU_DISPATCH_MODES_ALL BASELINE\n' > $(GEN)/mathfuncs_core.simd_declarations.hpp
	@printf '#define CV_CPU_SIMD_FILENAME "accum.simd.hpp"\n#define CV_CPU_DISPATCH_MODES_ALL BASELINE\n' > $(GEN)/accum.simd_declarations.hpp
	@printf '#include "opencv2/core/ocl.hpp"\n#include "opencv2/core/ocl_genbase.hpp"\n#include "opencv2/core/opencl/ocl_defs.hpp"\n' > $(GEN)/opencl_kernels_core.hpp
	@printf '#include "opencv2/core/ocl.hpp"\n#include "opencv2/core/ocl_genbase.hpp"\n#include "opencv2/core/opencl/ocl_defs.hpp"\n' > $(GEN)/opencl_kernels_imgproc.hpp
	@printf '#include "opencv2/core/ocl.hpp"\n#include "opencv2/core/ocl_genbase.hpp"\n#include "opencv2/core/opencl/ocl_defs.hpp"\n' > $(GEN)/opencl_kernels_stitching.hpp
	@touch $@

$(OBJ)/core/%.o: $(MOD)/core/src/%.cpp $(GEN)/.stamp
	$(CXX) $(CXXFLAGS) -I$(MOD)/core/src -c $< -o $@
$(OBJ)/imgproc/%.o: $(MOD)/imgproc/src/%.cpp $(GEN)/.stamp
	$(CXX) $(CXXFLAGS) -I$(MOD)/imgproc/src -c $< -o $@
$(OBJ)/stitching/%.o: $(MOD)/stitching/src/%.cpp $(GEN)/.stamp
	$(CXX) $(CXXFLAGS) -I$(MOD)/stitching/src -c $< -o $@
$(OBJ)/ref_shim.o: ref_shim.cpp $(GEN)/.stamp
	$(CXX) $(CXXFLAGS) -I$(MOD)/cudawarping/test -ffp-contract=off -fopenmp -c $< -o $@

$(OUT)/libvsref.so: $(CORE_OBJS) $(IMG_OBJS) $(ST_OBJS) $(OBJ)/ref_shim.o
	$(CXX) -shared -o $@ $^ -pthread -fopenmp -lz -ldl -lm
