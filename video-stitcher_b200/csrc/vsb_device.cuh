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

// One output sample of cuda::resize INTER_LINEAR on CV_8UC1 / CV_8UC3 (sources/modules/cudawarping/src/cuda/resize.cu:71-106): no
// half-pixel centre, fp32 weights in the kernel's order, out += src * w (an fma on the device), cvt.rni.sat.u8.
template <int CN>
__device__ __forceinline__ void resize_linear_px(const uint8_t *__restrict__ src, int sw, int sh, size_t sp, uint8_t *__restrict__ dst, size_t dp,
                                                 int dx, int dy, float fx, float fy)
{
    const float sx = __fmul_rn((float)dx, fx), sy = __fmul_rn((float)dy, fy);
    const int x1 = __float2int_rd(sx), y1 = __float2int_rd(sy), x2 = x1 + 1, y2 = y1 + 1;
    const int x2r = min(x2, sw - 1), y2r = min(y2, sh - 1);
    const float w11 = __fmul_rn(__fsub_rn((float)x2, sx), __fsub_rn((float)y2, sy)), w12 = __fmul_rn(__fsub_rn(sx, (float)x1), __fsub_rn((float)y2, sy));
    const float w21 = __fmul_rn(__fsub_rn((float)x2, sx), __fsub_rn(sy, (float)y1)), w22 = __fmul_rn(__fsub_rn(sx, (float)x1), __fsub_rn(sy, (float)y1));
#pragma unroll
    for (int c = 0; c < CN; ++c) {
        float o = __fmaf_rn((float)src[(size_t)y1 * sp + (size_t)x1 * CN + c], w11, 0.f);
        o = __fmaf_rn((float)src[(size_t)y1 * sp + (size_t)x2r * CN + c], w12, o);
        o = __fmaf_rn((float)src[(size_t)y2r * sp + (size_t)x1 * CN + c], w21, o);
        o = __fmaf_rn((float)src[(size_t)y2r * sp + (size_t)x2r * CN + c], w22, o);
        dst[(size_t)dy * dp + (size_t)dx * CN + c] = (uint8_t)rni_sat_u8(o);
    }
}

// round-half-even of v / 2^SH for two's-complement v (exact twin of cvt.rni on the exact fp32 value)
template <int SH>
__device__ __forceinline__ int rhe_shift(int v)
{
    return (v + ((1 << (SH - 1)) - 1) + ((v >> SH) & 1)) >> SH;
}

// BORDER_REFLECT (fedcba|abcdefgh|hgfedcb): sources/modules/cudev/include/opencv2/cudev/ptr2d/extrapolation.hpp:171-183
//   idx_high(i) = last - abs(last - i) + (i > last), idx_low(j) = (abs(j) - (j < 0)) % len.
// Within one fold (-len <= i <= 2*len - 1) that is the mirror below; the modulo form only runs for borders wider than
// the image (the reference's formula, not OpenCV's CPU borderInterpolate, is what the CUDA path computes there).
__device__ __forceinline__ int reflect_idx(int i, int len)
{
    const int last = len - 1;
    int j = i < 0 ? -i - 1 : (i > last ? 2 * last - i + 1 : i);
    if ((unsigned)j > (unsigned)last) {
        j = last - abs(last - i) + (i > last);
        j = (abs(j) - (j < 0)) % len;
    }
    return j;
}

// BORDER_REFLECT_101 as used by pyrDown: sources/modules/core/include/opencv2/core/cuda/border_interpolate.hpp:351-380
//   idx_low(i) = abs(i) % len, idx_high(i) = abs(last - abs(last - i)) % len, idx = idx_low(idx_high(i)).
// Within one fold (-(len-1) <= i <= 2*(len-1)) that is the plain mirror below; the modulo only matters for planes
// smaller than the tap reach, handled by the (rare) second branch so the common case costs four instructions.
__device__ __forceinline__ int r101_idx(int i, int len)
{
    const int last = len - 1;
    if ((unsigned)(i + last) > (unsigned)(3 * last)) return abs(abs(last - abs(last - i)) % len) % len;  // beyond one fold (or len == 1)
    const int j = abs(i);
    return j > last ? 2 * last - j : j;
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


// ---- fast exact bilinear tap machinery (same fp32 operations in the same order as remap_px_u8c3 above) ---------------
// u8 -> fp32 without the conversion pipe: PRMT builds 0x4B0000xx (= 2^23 + xx), one FADD removes the bias (exact).
__device__ __forceinline__ float u8_to_f32(unsigned packed, int byte)
{
    return __fsub_rn(__uint_as_float(__byte_perm(packed, 0x4B000000u, 0x7540u | (unsigned)byte)), 8388608.f);
}
// cvt.rni of a value known to lie in [0, 2^22): adding 1.5 * 2^23 rounds to nearest-even in the mantissa's low bits
__device__ __forceinline__ float rni_biased(float v) { return __fadd_rn(v, 12582912.f); }

// six consecutive bytes (two interleaved BGR pixels) starting at an arbitrary address, as lo = bytes 0..3, hi = bytes 4..5
__device__ __forceinline__ void load6(const uint8_t *__restrict__ p, bool bytewise, unsigned &lo, unsigned &hi)
{
    if (bytewise) {  // the last bytes of the image: never touch memory past the caller's buffer
        lo = (unsigned)__ldg(p) | ((unsigned)__ldg(p + 1) << 8) | ((unsigned)__ldg(p + 2) << 16) | ((unsigned)__ldg(p + 3) << 24);
        hi = (unsigned)__ldg(p + 4) | ((unsigned)__ldg(p + 5) << 8);
        return;
    }
    const size_t a = (size_t)p & ~(size_t)3;
    const unsigned sh = ((unsigned)(size_t)p & 3u) * 8u;
    const unsigned w0 = __ldg((const unsigned *)a), w1 = __ldg((const unsigned *)(a + 4));
    const unsigned w2 = sh == 24u ? __ldg((const unsigned *)(a + 8)) : 0u;
    lo = __funnelshift_r(w0, w1, sh);
    hi = __funnelshift_r(w1, w2, sh);
}

// remap LINEAR / BORDER_CONSTANT(0) of one interleaved u8x3 pixel followed by the gain convertTo:
//   p = sat_u8(rni(gain * sat_u8(rni(bilinear))))   (360_stitcher/timed.cpp:90-94); gain = 1 gives the plain remap.
// Returns b | g << 8 | r << 16.  Requires sw >= 2.  Bit-identical to remap_px_u8c3 + rni_sat_u8(gain * v): the taps,
// weights and the fmul/fma chain are the same; only the u8<->fp32 conversions take a cheaper route.
// General form: any coordinate (taps outside the image read 0, NaN gives 0, the last bytes of the image are read bytewise).
template <bool GAIN>
__device__ __noinline__ unsigned remap_gain_px_edge(const uint8_t *__restrict__ src, size_t pitch, int sw, int sh, float x, float y, float gain)
{
    const int x1 = __float2int_rd(x), y1 = __float2int_rd(y);
    // no tap in range (or NaN coordinates: every product is NaN and cvt.sat gives 0)
    if (x1 < -1 || x1 >= sw || y1 < -1 || y1 >= sh || !(x == x) || !(y == y)) return 0u;
    const float fx2 = __fsub_rn((float)(x1 + 1), x), fx1 = __fsub_rn(x, (float)x1);
    const float fy2 = __fsub_rn((float)(y1 + 1), y), fy1 = __fsub_rn(y, (float)y1);
    const float w11 = __fmul_rn(fx2, fy2), w12 = __fmul_rn(fx1, fy2), w21 = __fmul_rn(fx2, fy1), w22 = __fmul_rn(fx1, fy1);
    const int xs = min(max(x1, 0), sw - 2);
    unsigned l1 = 0, r1 = 0, l2 = 0, r2 = 0;  // 24-bit BGR of the four taps
    const uint8_t *row = src + (size_t)max(y1, 0) * pitch + (size_t)xs * 3;
    const bool tail = xs + 4 >= sw;  // word loads could run up to 3 bytes past the end of the last row
    if (y1 >= 0) {
        unsigned lo, hi;
        load6(row, tail && y1 == sh - 1, lo, hi);
        l1 = lo & 0xffffffu; r1 = (lo >> 24) | ((hi & 0xffffu) << 8);
    }
    if (y1 + 1 < sh) {
        unsigned lo, hi;
        load6(y1 >= 0 ? row + pitch : row, tail && y1 + 1 == sh - 1, lo, hi);
        l2 = lo & 0xffffffu; r2 = (lo >> 24) | ((hi & 0xffffu) << 8);
    }
    if (x1 != xs) {  // x1 == -1: left taps are outside; x1 == sw - 1: right taps are outside
        if (x1 < 0) { r1 = l1; r2 = l2; l1 = l2 = 0u; } else { l1 = r1; l2 = r2; r1 = r2 = 0u; }
    }
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = __fmul_rn(u8_to_f32(l1, c), w11);
        v = __fmaf_rn(u8_to_f32(r1, c), w12, v);
        v = __fmaf_rn(u8_to_f32(l2, c), w21, v);
        v = __fmaf_rn(u8_to_f32(r2, c), w22, v);
        v = rni_biased(v);                                                 // sat_u8(rni(.)): integer in the low mantissa bits
        o[c] = GAIN ? rni_biased(fminf(__fmul_rn(gain, __fsub_rn(v, 12582912.f)), 255.f)) : v;  // sat_u8(rni(gain * v))
    }
    return __byte_perm(__byte_perm(__float_as_uint(o[0]), __float_as_uint(o[1]), 0x0040u), __float_as_uint(o[2]), 0x0410u) & 0xffffffu;
}

// Hot form.  Taps outside the image contribute exactly +0 in the reference (0 * w with w >= 0), which is what a ZERO
// WEIGHT on a clamped, in-bounds tap gives too -- so image borders need no selects on the pixel data: a short, rarely
// taken block zeroes the 1-D fractions and clamps the row pointers.  Only the few pixels whose byte window would leave
// the caller's buffer (first / last bytes of the image) take the general form above.  (The kernels are bound by the
// ALU pipe, so the floor / int->float conversions deliberately stay on the otherwise idle conversion unit.)
template <bool GAIN>
__device__ __forceinline__ unsigned remap_gain_px(const uint8_t *__restrict__ src, size_t pitch, int sw, int sh, float x, float y, float gain)
{
    const int x1 = __float2int_rd(x), y1 = __float2int_rd(y);
    // no tap in range: x1 outside [-1, sw - 1] or y1 outside [-1, sh - 1]; NaN coordinates give 0 as well
    if ((unsigned)(x1 + 1) > (unsigned)sw || (unsigned)(y1 + 1) > (unsigned)sh || !(x == x) || !(y == y)) return 0u;
    float fx2 = __fsub_rn((float)(x1 + 1), x), fx1 = __fsub_rn(x, (float)x1);
    float fy2 = __fsub_rn((float)(y1 + 1), y), fy1 = __fsub_rn(y, (float)y1);
    const uint8_t *p1 = src + (ptrdiff_t)y1 * (ptrdiff_t)pitch + (ptrdiff_t)x1 * 3, *p2 = p1 + pitch;
    if ((unsigned)x1 >= (unsigned)(sw - 1) || (unsigned)y1 >= (unsigned)(sh - 2)) {  // a tap outside, or the last image row
        const int ya = max(y1, 0), yb = min(y1 + 1, sh - 1);
        if ((ya == 0 && x1 < 0) || (yb == sh - 1 && x1 + 5 >= sw)) return remap_gain_px_edge<GAIN>(src, pitch, sw, sh, x, y, gain);
        fx2 = x1 >= 0 ? fx2 : 0.f;
        fx1 = x1 + 1 < sw ? fx1 : 0.f;
        fy2 = y1 >= 0 ? fy2 : 0.f;
        fy1 = y1 + 1 < sh ? fy1 : 0.f;
        p1 = src + (size_t)ya * pitch + (ptrdiff_t)x1 * 3;
        p2 = src + (size_t)yb * pitch + (ptrdiff_t)x1 * 3;
    }
    const float w11 = __fmul_rn(fx2, fy2), w12 = __fmul_rn(fx1, fy2), w21 = __fmul_rn(fx2, fy1), w22 = __fmul_rn(fx1, fy1);
    unsigned lo1, hi1, lo2, hi2;
    {
        const size_t a = (size_t)p1 & ~(size_t)3;
        const unsigned s8 = ((unsigned)(size_t)p1 & 3u) * 8u;
        const unsigned w0 = __ldg((const unsigned *)a), w1 = __ldg((const unsigned *)(a + 4));
        const unsigned w2 = s8 == 24u ? __ldg((const unsigned *)(a + 8)) : 0u;
        lo1 = __funnelshift_r(w0, w1, s8); hi1 = __funnelshift_r(w1, w2, s8);
    }
    {
        const size_t a = (size_t)p2 & ~(size_t)3;
        const unsigned s8 = ((unsigned)(size_t)p2 & 3u) * 8u;
        const unsigned w0 = __ldg((const unsigned *)a), w1 = __ldg((const unsigned *)(a + 4));
        const unsigned w2 = s8 == 24u ? __ldg((const unsigned *)(a + 8)) : 0u;
        lo2 = __funnelshift_r(w0, w1, s8); hi2 = __funnelshift_r(w1, w2, s8);
    }
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {  // left pixel = bytes 0..2 of lo, right pixel = byte 3 of lo, bytes 0..1 of hi
        float v = __fmul_rn(u8_to_f32(lo1, c), w11);
        v = __fmaf_rn(c == 0 ? u8_to_f32(lo1, 3) : u8_to_f32(hi1, c - 1), w12, v);
        v = __fmaf_rn(u8_to_f32(lo2, c), w21, v);
        v = __fmaf_rn(c == 0 ? u8_to_f32(lo2, 3) : u8_to_f32(hi2, c - 1), w22, v);
        v = rni_biased(v);
        o[c] = GAIN ? rni_biased(fminf(__fmul_rn(gain, __fsub_rn(v, 12582912.f)), 255.f)) : v;
    }
    return __byte_perm(__byte_perm(__float_as_uint(o[0]), __float_as_uint(o[1]), 0x0040u), __float_as_uint(o[2]), 0x0410u) & 0xffffffu;
}

// ---- table-driven taps ----------------------------------------------------------------------------------------------
// The maps are static between calibrations, so everything remap_gain_px derives from the coordinates alone -- floor, the
// four weight products, the range tests, the clamped tap address -- is computed ONCE per map (make_tap_entry, same fp32
// operations as the reference) and stored as {byte offset of the 2x2 window, wa, wb, wc, wd}.  The per-frame kernel is
// then loads + the fmul / fma chain.  Taps outside the image keep their place in the chain with weight 0 on an
// in-bounds sample (s * 0 = +0 exactly, like the reference's 0 * w), see DESIGN.md section 3.
struct TapEntry {
    int off;              // byte offset of the window's first sample; TAP_SLOW: take the coordinate-driven path
    float wa, wb, wc, wd; // weights of (row 0, px 0), (row 0, px 1), (row 1, px 0), (row 1, px 1) of the window
};
constexpr int TAP_SLOW = (int)0x80000000;

// Image of sw x sh pixels (sw, sh >= 2) whose pixel (0, 0) sits at byte `origin` and whose rows are `pitch` bytes apart.
// `guard` = true: taps one pixel outside the image exist in memory and read 0 (zero frame around the buffer), so no
// clamping is needed; false: clamp into the image and move the weights (order of the non-zero terms is preserved).
__device__ __forceinline__ TapEntry make_tap_entry(float x, float y, int sw, int sh, unsigned pitch, unsigned origin, bool guard, unsigned bpp = 3u)
{
    TapEntry e;
    e.off = (int)origin; e.wa = e.wb = e.wc = e.wd = 0.f;
    if (!(x == x) || !(y == y)) return e;                         // NaN coordinates: cvt.sat(NaN) = 0
    if (!(x >= -1.f && x < (float)sw && y >= -1.f && y < (float)sh)) return e;  // no tap in range
    const int x1 = __float2int_rd(x), y1 = __float2int_rd(y);
    const float fx2 = __fsub_rn((float)(x1 + 1), x), fx1 = __fsub_rn(x, (float)x1);
    const float fy2 = __fsub_rn((float)(y1 + 1), y), fy1 = __fsub_rn(y, (float)y1);
    float wa = __fmul_rn(fx2, fy2), wb = __fmul_rn(fx1, fy2), wc = __fmul_rn(fx2, fy1), wd = __fmul_rn(fx1, fy1);
    int xs = x1, ys = y1;
    if (!guard) {
        if (x1 < 0) { xs = 0; wa = wb; wb = 0.f; wc = wd; wd = 0.f; }
        else if (x1 > sw - 2) { xs = sw - 2; wb = wa; wa = 0.f; wd = wc; wc = 0.f; }
        if (y1 < 0) { ys = 0; wa = wc; wb = wd; wc = wd = 0.f; }
        else if (y1 > sh - 2) { ys = sh - 2; wc = wa; wd = wb; wa = wb = 0.f; }
    }
    e.off = (int)(origin + (unsigned)(ys * (int)pitch) + (unsigned)xs * bpp);  // bpp: 3 = interleaved BGR, 1 = the luma plane of an NV12 frame
    e.wa = wa; e.wb = wb; e.wc = wc; e.wd = wd;
    return e;
}

// Table weights are stored pre-multiplied by 2^126 (TAP_WSCALE) and the taps enter the chain as DENORMALS: PRMT puts the byte
// into the low mantissa bits of a zero word, i.e. the float b * 2^-149 -- no bias to remove.  Every product and partial sum of
// the chain is then the reference's value times 2^-23 exactly (scaling by a power of two commutes with rounding as long as
// nothing leaves the normal range: the builders route entries with a non-zero weight below TAP_WMIN to the coordinate path),
// and ONE fused multiply-add (v * 2^23 + 1.5 * 2^23) undoes the scale and rounds half-to-even.  12 FADDs fewer per pixel.
constexpr float TAP_WSCALE = 8.507059173023462e37f;   // 2^126
constexpr float TAP_WMIN = 8.077935669463161e-28f;    // 2^-90: products stay normal (255 * w * 2^-23 >= 2^-113)
__device__ __forceinline__ float u8_den(unsigned packed, int byte)
{
    return __uint_as_float(__byte_perm(packed, 0u, 0x4440u | (unsigned)byte));
}

// One pixel from a table entry: `base` is 4-byte aligned, `pitch` a multiple of 4 (both rows share the byte shift).
template <bool GAIN>
__device__ __forceinline__ unsigned remap_tab_px(const uint8_t *__restrict__ base, unsigned pitch, unsigned off, float wa, float wb, float wc, float wd, float gain)
{
    const uint8_t *a = base + (off & ~3u);
    const unsigned s8 = (off & 3u) * 8u;
    const unsigned t0 = __ldg((const unsigned *)a), t1 = __ldg((const unsigned *)(a + 4));
    const unsigned u0 = __ldg((const unsigned *)(a + pitch)), u1 = __ldg((const unsigned *)(a + pitch + 4));
    unsigned t2 = 0u, u2 = 0u;
    if (s8 == 24u) { t2 = __ldg((const unsigned *)(a + 8)); u2 = __ldg((const unsigned *)(a + pitch + 8)); }
    const unsigned lo1 = __funnelshift_r(t0, t1, s8), hi1 = __funnelshift_r(t1, t2, s8);
    const unsigned lo2 = __funnelshift_r(u0, u1, s8), hi2 = __funnelshift_r(u1, u2, s8);
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {  // left pixel = bytes 0..2 of lo, right pixel = byte 3 of lo, bytes 0..1 of hi
        float v = __fmul_rn(u8_den(lo1, c), wa);
        v = __fmaf_rn(c == 0 ? u8_den(lo1, 3) : u8_den(hi1, c - 1), wb, v);
        v = __fmaf_rn(u8_den(lo2, c), wc, v);
        v = __fmaf_rn(c == 0 ? u8_den(lo2, 3) : u8_den(hi2, c - 1), wd, v);
        v = __fmaf_rn(v, 8388608.f, 12582912.f);   // undo the 2^-23 scale and round: sat_u8(rni(.)) as an integer in the low mantissa bits
        o[c] = GAIN ? rni_biased(fminf(__fmul_rn(gain, __fsub_rn(v, 12582912.f)), 255.f)) : v;
    }
    return __byte_perm(__byte_perm(__float_as_uint(o[0]), __float_as_uint(o[1]), 0x0040u), __float_as_uint(o[2]), 0x0410u) & 0xffffffu;
}

}  // namespace vsb
