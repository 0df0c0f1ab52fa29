// vsb_internal.h -- shared host/device declarations of libvsb200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/vsb200.h"
#include "vsb_device.cuh"

namespace vsb {

struct Mat3 { float m[9]; };

// thread-local error text behind vsb_last_error(); returns `code`
int fail(int code, const char *fmt, ...);
int check_launch(const char *what);
int check_cuda(cudaError_t e, const char *what);

// ProjectorBase::setCameraParams (sources/modules/stitching/src/warpers.cpp:49-79), host side
void projector_setup(const float K[9], const float R[9], float k_rinv[9], float r_kinv[9], float rinv[9]);

// one output sample of the application's `resize` kernel (360_stitcher/resize.cu:12-26)
__device__ __forceinline__ float custom_resize_at(const float *__restrict__ in, int cols, int rows, size_t in_pitch,
                                                  int tx, int ty, int u, int v)
{
    const int left = u * (cols - 1) / tx;
    const int top = v * (rows - 1) / ty;
    const float uu = __fsub_rn(__fdiv_rn(__fmul_rn((float)u, (float)(cols - 1)), (float)tx), (float)left);
    const float vv = __fsub_rn(__fdiv_rn(__fmul_rn((float)v, (float)(rows - 1)), (float)ty), (float)top);
    const float *r0 = (const float *)((const char *)in + (size_t)top * in_pitch) + left;
    const float *r1 = (const float *)((const char *)in + (size_t)(top + 1) * in_pitch) + left;
    const float iu = __fsub_rn(1.f, uu), iv = __fsub_rn(1.f, vv);
    // nvcc's contraction of the reference's four-term sum (read off the PTX of 360_stitcher/resize.cu): the second
    // term is a rounded multiply, the first one is fused onto it, then the third and the fourth
    float a = __fmul_rn(__fmul_rn(uu, iv), __ldg(r0 + 1));
    a = __fmaf_rn(__fmul_rn(iu, iv), __ldg(r0), a);
    a = __fmaf_rn(__fmul_rn(iu, vv), __ldg(r1), a);
    a = __fmaf_rn(__fmul_rn(uu, vv), __ldg(r1 + 1), a);
    return a;
}

}  // namespace vsb

// internal: records the rig parameters after vsb_calibrate_rig
extern "C" int vsb_note_rig(vsb_stitcher *s, int projection, float scale, int src_w, int src_h);
// internal: seam-scale state of vsb_calibrate_rig_device (owned by vsb_calib.cu, released with the handle)
extern "C" void vsb_attach_calib(vsb_stitcher *s, void *state, void (*dtor)(void *));
extern "C" void *vsb_get_calib(const vsb_stitcher *s);
extern "C" int vsb_handle_device(const vsb_stitcher *s);
// internal: marks a view as a column window of a camera's warped image (vsb_calibrate_rig_split)
extern "C" int vsb_set_view_window(vsb_stitcher *s, int view, int camera, int x0, int full_w);
