// oracle/ptx_wrap_device.cu -- TEST INFRASTRUCTURE.  Thin __global__ wrappers around device functions of the PRODUCT
// (video-stitcher_b200/csrc), compiled to PTX with the product's own flags (--fmad=false) so that tests/test_oracle_ptx.py can
// execute them on the CPU (oracle/ptx_interp.py) next to the PTX of the reference's kernels: a device-vs-reference
// comparison that needs no GPU.  Nothing here is linked into libvsb200.so.
#include "../video-stitcher_b200/csrc/vsb_internal.h"

extern "C" __global__ void w_custom_resize(int tx, int ty, int cols, int rows, const float *in, size_t in_pitch, float *out, size_t out_pitch)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u < tx && v < ty) *(float *)((char *)out + (size_t)v * out_pitch + (size_t)u * 4) = vsb::custom_resize_at(in, cols, rows, in_pitch, tx, ty, u, v);
}
