// vsb_device.cuh -- rounding / border helpers shared by every kernel (sm_100a).
//
// The arithmetic contract mirrors the reference's CUDA kernels:
//   * cvt.rni.sat for remap / gain / pyramids (sources/modules/core/include/opencv2/core/cuda/saturate_cast.hpp:96-101,221-226)
//   * truncation (cvt.rzi) in the weighted add / normalise (sources/modules/stitching/src/cuda/multiband_blend.cu:46-48,95-98)
//   * every fp32 product/sum that nvcc would contract in the reference is an explicit __fmaf_rn here, every
//     other op an explicit _rn intrinsic, so the result does not depend on compiler flags.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vsb {

__device__ __forceinline__ unsigned rni_sat_u8(float v)
{
    unsigned r;
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ short rni_sat_s16(float v)
{
    short r;
    asm("cvt.rni.sat.s16.f32 %0, %1;" : "=h"(r) : "f"(v));
    return r;
}

// static_cast<short>(float) of device code
__device__ __forceinline__ int rz_s16(float v)
{
    short r;
    asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(r) : "f"(v));
    return (int)r;
}

__device__ __forceinline__ int sat_s16(int v) { return max(-32768, min(32767, v)); }

// round-half-even of v / 2^SH for two's-complement v (exact twin of cvt.rni on the exact fp32 value)
template <int SH>
__device__ __forceinline__ int rhe_shift(int v)
{
    return (v + ((1 << (SH - 1)) - 1) + ((v >> SH) & 1)) >> SH;
}

// BORDER_REFLECT (fedcba|abcdefgh|hgfedcb): sources/modules/cudev/include/opencv2/cudev/ptr2d/extrapolation.hpp:171-183
__device__ __forceinline__ int reflect_idx(int i, int len)
{
    const int last = len - 1;
    int j = last - abs(last - i) + (i > last);
    return (abs(j) - (j < 0)) % len;
}

// BORDER_REFLECT_101 as used by pyrDown: sources/modules/core/include/opencv2/core/cuda/border_interpolate.hpp:351-380
//   idx_low(i) = abs(i) % len, idx_high(i) = abs(last - abs(last - i)) % len, idx = idx_low(idx_high(i)).
// Within one fold (-(len-1) <= i <= 2*(len-1)) that is the plain mirror below; the modulo only matters for planes
// smaller than the tap reach, handled by the (rare) second branch so the common case costs four instructions.
__device__ __forceinline__ int r101_idx(int i, int len)
{
    const int last = len - 1;
    int j = abs(i);
    j = j > last ? 2 * last - j : j;
    if ((unsigned)j > (unsigned)last) j = abs(abs(last - abs(last - i)) % len) % len;
    return j;
}

// pyrUp source index: abs() at the low edge, clamp at the high edge (sources/modules/cudawarping/src/cuda/pyr_up.cu:70-74)
__device__ __forceinline__ int up_idx(int i, int n) { return min(abs(i), n - 1); }

// Bilinear tap weights of LinearFilter (sources/modules/core/include/opencv2/core/cuda/filters.hpp:95-112)
struct BilinearTaps {
    int x1, y1;
    float w11, w12, w21, w22;
};

__device__ __forceinline__ BilinearTaps make_taps(float x, float y)
{
    BilinearTaps t;
    t.x1 = __float2int_rd(x);
    t.y1 = __float2int_rd(y);
    const float fx2 = __fsub_rn((float)(t.x1 + 1), x), fx1 = __fsub_rn(x, (float)t.x1);
    const float fy2 = __fsub_rn((float)(t.y1 + 1), y), fy1 = __fsub_rn(y, (float)t.y1);
    t.w11 = __fmul_rn(fx2, fy2);
    t.w12 = __fmul_rn(fx1, fy2);
    t.w21 = __fmul_rn(fx2, fy1);
    t.w22 = __fmul_rn(fx1, fy1);
    return t;
}

__device__ __forceinline__ float bilerp(float s11, float s12, float s21, float s22, const BilinearTaps &t)
{
    float o = __fmul_rn(s11, t.w11);
    o = __fmaf_rn(s12, t.w12, o);
    o = __fmaf_rn(s21, t.w21, o);
    o = __fmaf_rn(s22, t.w22, o);
    return o;
}

// remap LINEAR / BORDER_CONSTANT(0) of one interleaved u8x3 pixel; returns packed b | g<<8 | r<<16
__device__ __forceinline__ unsigned remap_px_u8c3(const uint8_t *__restrict__ src, size_t pitch, int sw, int sh, float x, float y)
{
    const BilinearTaps t = make_taps(x, y);
    const bool inx1 = (unsigned)t.x1 < (unsigned)sw, inx2 = (unsigned)(t.x1 + 1) < (unsigned)sw;
    const bool iny1 = (unsigned)t.y1 < (unsigned)sh, iny2 = (unsigned)(t.y1 + 1) < (unsigned)sh;
    float s11[3] = {0.f, 0.f, 0.f}, s12[3] = {0.f, 0.f, 0.f}, s21[3] = {0.f, 0.f, 0.f}, s22[3] = {0.f, 0.f, 0.f};
    const uint8_t *r1 = src + (size_t)t.y1 * pitch + (size_t)t.x1 * 3;
    const uint8_t *r2 = r1 + pitch;
    if (iny1 && inx1) { s11[0] = __ldg(r1); s11[1] = __ldg(r1 + 1); s11[2] = __ldg(r1 + 2); }
    if (iny1 && inx2) { s12[0] = __ldg(r1 + 3); s12[1] = __ldg(r1 + 4); s12[2] = __ldg(r1 + 5); }
    if (iny2 && inx1) { s21[0] = __ldg(r2); s21[1] = __ldg(r2 + 1); s21[2] = __ldg(r2 + 2); }
    if (iny2 && inx2) { s22[0] = __ldg(r2 + 3); s22[1] = __ldg(r2 + 4); s22[2] = __ldg(r2 + 5); }
    unsigned out = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) out |= rni_sat_u8(bilerp(s11[c], s12[c], s21[c], s22[c], t)) << (8 * c);
    return out;
}

}  // namespace vsb
