// vsb_primitives.cu -- the B7 "device-launcher layer": one CUDA primitive per reference kernel, on the
// reference's own interleaved OpenCV layouts (CV_8UC3 / CV_16SC3 / CV_32FC1, byte pitches).  These are the
// drop-ins for code that still calls the unfused stages (calibration, tests, recalibration); the per-frame
// product path is the fused pipeline in vsb_pipeline.cu.
#include "vsb_internal.h"

namespace vsb {

// ---------------------------------------------------------------------------------------------------------
// cuda::remap LINEAR / BORDER_CONSTANT, CV_8UC3 (sources/modules/cudawarping/src/cuda/remap.cu:56-68)
__global__ void k_remap_linear_u8c3(const uint8_t *__restrict__ src, int sw, int sh, size_t sp,
                                    const float *__restrict__ xmap, const float *__restrict__ ymap, size_t mp,
                                    uint8_t *__restrict__ dst, int dw, int dh, size_t dp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const float fx = *(const float *)((const char *)xmap + (size_t)y * mp + (size_t)x * 4);
    const float fy = *(const float *)((const char *)ymap + (size_t)y * mp + (size_t)x * 4);
    // the fused path's tap routine (word loads, biased conversions); the plain byte-wise form for 1-pixel-wide sources
    const unsigned v = sw >= 2 ? remap_gain_px<false>(src, sp, sw, sh, fx, fy, 1.f) : remap_px_u8c3(src, sp, sw, sh, fx, fy);
    uint8_t *d = dst + (size_t)y * dp + (size_t)x * 3;
    d[0] = v & 0xff; d[1] = (v >> 8) & 0xff; d[2] = (v >> 16) & 0xff;
}

// GpuMat::convertTo(CV_8U, alpha) (sources/modules/core/src/cuda/gpu_mat.cu:488-498)
__global__ void k_gain_u8(uint8_t *buf, int wbytes, int h, size_t pitch, float gain)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= wbytes || y >= h) return;
    uint8_t *p = buf + (size_t)y * pitch + x;
    *p = (uint8_t)rni_sat_u8(__fmul_rn(gain, (float)*p));
}

// cuda::copyMakeBorder(BORDER_REFLECT) + convertTo(CV_16S)
__global__ void k_border_reflect_u8c3_s16c3(const uint8_t *__restrict__ src, int w, int h, size_t sp, int top, int left,
                                            int16_t *__restrict__ dst, int dw, int dh, size_t dp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const uint8_t *s = src + (size_t)reflect_idx(y - top, h) * sp + (size_t)reflect_idx(x - left, w) * 3;
    int16_t *d = (int16_t *)((char *)dst + (size_t)y * dp) + (size_t)x * 3;
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
}

// cuda::pyrDown CV_16SC3 (sources/modules/cudawarping/src/cuda/pyr_down.cu:55-174): exact integer form of the fp32 taps
__global__ void k_pyr_down_s16c3(const int16_t *__restrict__ src, int w, int h, size_t sp, int16_t *__restrict__ dst,
                                 int dw, int dh, size_t dp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const int K5[5] = {1, 4, 6, 4, 1};
    int acc[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int16_t *row = (const int16_t *)((const char *)src + (size_t)r101_idx(2 * y + j - 2, h) * sp);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int16_t *p = row + (size_t)r101_idx(2 * x + i - 2, w) * 3;
            const int kw = K5[j] * K5[i];
            acc[0] += kw * p[0]; acc[1] += kw * p[1]; acc[2] += kw * p[2];
        }
    }
    int16_t *d = (int16_t *)((char *)dst + (size_t)y * dp) + (size_t)x * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c] = (int16_t)sat_s16(rhe_shift<8>(acc[c]));
}

// cuda::pyrUp CV_16SC3 (sources/modules/cudawarping/src/cuda/pyr_up.cu:55-145)
__global__ void k_pyr_up_s16c3(const int16_t *__restrict__ src, int w, int h, size_t sp, int16_t *__restrict__ dst, size_t dp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= 2 * w || y >= 2 * h) return;
    const int ix = x >> 1, iy = y >> 1;
    int wx[3], wy[3];
    if (x & 1) { wx[0] = 0; wx[1] = 4; wx[2] = 4; } else { wx[0] = 1; wx[1] = 6; wx[2] = 1; }
    if (y & 1) { wy[0] = 0; wy[1] = 4; wy[2] = 4; } else { wy[0] = 1; wy[1] = 6; wy[2] = 1; }
    int acc[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int16_t *row = (const int16_t *)((const char *)src + (size_t)up_idx(iy + j - 1, h) * sp);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int16_t *p = row + (size_t)up_idx(ix + i - 1, w) * 3;
            const int kw = wy[j] * wx[i];
            acc[0] += kw * p[0]; acc[1] += kw * p[1]; acc[2] += kw * p[2];
        }
    }
    int16_t *d = (int16_t *)((char *)dst + (size_t)y * dp) + (size_t)x * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c] = (int16_t)sat_s16(rhe_shift<6>(acc[c]));
}

// cuda::pyrDown CV_32FC1: not exact in fp32 -> keep the reference's order (vertical pass, then horizontal) and its
// contraction pattern sum = 0.0625*a; sum = fma(0.25,b,sum); ...
__global__ void k_pyr_down_f32(const float *__restrict__ src, int w, int h, size_t sp, float *__restrict__ dst, int dw, int dh, size_t dp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const float *r[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) r[j] = (const float *)((const char *)src + (size_t)r101_idx(2 * y + j - 2, h) * sp);
    float col[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int sx = r101_idx(2 * x + i - 2, w);
        float s = __fmul_rn(0.0625f, r[0][sx]);
        s = __fmaf_rn(0.25f, r[1][sx], s);
        s = __fmaf_rn(0.375f, r[2][sx], s);
        s = __fmaf_rn(0.25f, r[3][sx], s);
        s = __fmaf_rn(0.0625f, r[4][sx], s);
        col[i] = s;
    }
    float s = __fmul_rn(0.0625f, col[0]);
    s = __fmaf_rn(0.25f, col[1], s);
    s = __fmaf_rn(0.375f, col[2], s);
    s = __fmaf_rn(0.25f, col[3], s);
    s = __fmaf_rn(0.0625f, col[4], s);
    *(float *)((char *)dst + (size_t)y * dp + (size_t)x * 4) = s;
}

// addSrcWeightKernel32F (sources/modules/stitching/src/cuda/multiband_blend.cu:36-50)
__global__ void k_add_src_weight_32f(const int16_t *__restrict__ src, size_t sp, const float *__restrict__ wgt, size_t wp,
                                     int16_t *dst, size_t dp, float *dstw, size_t dwp, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int16_t *s = (const int16_t *)((const char *)src + (size_t)y * sp) + (size_t)x * 3;
    const float wv = *(const float *)((const char *)wgt + (size_t)y * wp + (size_t)x * 4);
    int16_t *d = (int16_t *)((char *)dst + (size_t)y * dp) + (size_t)x * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c] = (int16_t)(d[c] + rz_s16(__fmul_rn((float)s[c], wv)));
    float *pw = (float *)((char *)dstw + (size_t)y * dwp + (size_t)x * 4);
    *pw = __fadd_rn(*pw, wv);
}

// normalizeUsingWeightKernel32F (multiband_blend.cu:85-99)
__global__ void k_normalize_32f(const float *__restrict__ wgt, size_t wp, int16_t *src, size_t sp, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    int16_t *s = (int16_t *)((char *)src + (size_t)y * sp) + (size_t)x * 3;
    const float wv = __fadd_rn(*(const float *)((const char *)wgt + (size_t)y * wp + (size_t)x * 4), 1e-5f);
#pragma unroll
    for (int c = 0; c < 3; ++c) s[c] = (int16_t)rz_s16(__fdiv_rn((float)s[c], wv));
}

// kernel `resize` of the application (360_stitcher/resize.cu:9-27)
__global__ void k_custom_resize(const float *__restrict__ in, int cols, int rows, size_t ip, float *__restrict__ out, int tx, int ty, size_t op)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u >= tx || v >= ty) return;
    *(float *)((char *)out + (size_t)v * op + (size_t)u * 4) = custom_resize_at(in, cols, rows, ip, tx, ty, u, v);
}

// buildWarpMapsKernel<SphericalMapper|CylindricalMapper> (sources/modules/stitching/src/cuda/build_warp_maps.cu:88-152)
__global__ void k_build_maps(int proj, float scale, Mat3 k, int tl_u, int tl_v, int cols, int rows, float *mx, float *my, size_t pitch)
{
    const int du = blockIdx.x * blockDim.x + threadIdx.x, dv = blockIdx.y * blockDim.y + threadIdx.y;
    if (du >= cols || dv >= rows) return;
    float u = (float)(tl_u + du), v = (float)(tl_v + dv);
    // What nvcc makes of  x = k[0]*x_ + k[1]*y_ + k[2]*z_  (read off the PTX of the reference's build_warp_maps.cu):
    // spherical, where y_ = -cos(v): the negation is folded into a subtraction of two separately rounded products and only the
    // third term is fused; cylindrical: the second product is rounded on its own, the first and the third are fused onto it.
    float x, y, z;
    if (proj == VSB_PROJ_SPHERICAL) {
        v = __fdiv_rn(v, scale); u = __fdiv_rn(u, scale);
        const float sinv = sinf(v);
        const float x_ = __fmul_rn(sinv, sinf(u));
        const float cosv = cosf(v);
        const float z_ = __fmul_rn(sinv, cosf(u));
        x = __fmaf_rn(k.m[2], z_, __fsub_rn(__fmul_rn(x_, k.m[0]), __fmul_rn(cosv, k.m[1])));
        y = __fmaf_rn(k.m[5], z_, __fsub_rn(__fmul_rn(x_, k.m[3]), __fmul_rn(cosv, k.m[4])));
        z = __fmaf_rn(k.m[8], z_, __fsub_rn(__fmul_rn(x_, k.m[6]), __fmul_rn(cosv, k.m[7])));
    } else {
        u = __fdiv_rn(u, scale);
        const float x_ = sinf(u);
        const float y_ = __fdiv_rn(v, scale);
        const float z_ = cosf(u);
        x = __fmaf_rn(k.m[2], z_, __fmaf_rn(x_, k.m[0], __fmul_rn(y_, k.m[1])));
        y = __fmaf_rn(k.m[5], z_, __fmaf_rn(x_, k.m[3], __fmul_rn(y_, k.m[4])));
        z = __fmaf_rn(k.m[8], z_, __fmaf_rn(x_, k.m[6], __fmul_rn(y_, k.m[7])));
    }
    if (z > 0) { x = __fdiv_rn(x, z); y = __fdiv_rn(y, z); } else { x = y = -1.f; }
    *(float *)((char *)mx + (size_t)dv * pitch + (size_t)du * 4) = x;
    *(float *)((char *)my + (size_t)dv * pitch + (size_t)du * 4) = y;
}

// cuda::remap as RotationWarperGpu::warp calls it (sources/modules/stitching/src/warpers_cuda.cpp:279-298 ->
// sources/modules/cudawarping/src/cuda/remap.cu:56-68): PointFilter / LinearFilter over BrdConstant(0) / BrdReflect readers
// (sources/modules/core/include/opencv2/core/cuda/filters.hpp:59-117, border_interpolate.hpp:484-534,698-717), CV_8UC1 / CV_8UC3.
// Calibration-time code (seam-scale images and masks, 360_stitcher/calibration.cpp:118,122,227): one thread per pixel.
template <int CN>
__global__ void k_warp_remap(const uint8_t *__restrict__ src, int sw, int sh, size_t sp, const float *__restrict__ xmap,
                             const float *__restrict__ ymap, size_t mp, uint8_t *__restrict__ dst, int dw, int dh, size_t dp, int interp, int border)
{
    const int xx = blockIdx.x * blockDim.x + threadIdx.x, yy = blockIdx.y * blockDim.y + threadIdx.y;
    if (xx >= dw || yy >= dh) return;
    const float x = *(const float *)((const char *)xmap + (size_t)yy * mp + (size_t)xx * 4);
    const float y = *(const float *)((const char *)ymap + (size_t)yy * mp + (size_t)xx * 4);
    auto rd = [&](int r, int c, int ch) -> float {
        if (border == 0) return ((unsigned)c < (unsigned)sw && (unsigned)r < (unsigned)sh) ? (float)src[(size_t)r * sp + (size_t)c * CN + ch] : 0.f;
        return (float)src[(size_t)reflect_idx(r, sh) * sp + (size_t)reflect_idx(c, sw) * CN + ch];
    };
    uint8_t *d = dst + (size_t)yy * dp + (size_t)xx * CN;
    if (interp == 0) {
        const int c = __float2int_rz(x), r = __float2int_rz(y);
#pragma unroll
        for (int ch = 0; ch < CN; ++ch) d[ch] = (uint8_t)rd(r, c, ch);
        return;
    }
    const BilinearTaps t = make_taps(x, y);
#pragma unroll
    for (int ch = 0; ch < CN; ++ch)
        d[ch] = (uint8_t)rni_sat_u8(bilerp(rd(t.y1, t.x1, ch), rd(t.y1, t.x1 + 1, ch), rd(t.y1 + 1, t.x1, ch), rd(t.y1 + 1, t.x1 + 1, ch), t));
}

}  // namespace vsb

using namespace vsb;

static inline dim3 grid2d(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y); }

#define VSB_REQUIRE(cond, msg) do { if (!(cond)) return vsb::fail(VSB_ERR_INVALID, msg); } while (0)

extern "C" {

int vsb_remap_linear_u8c3(const uint8_t *d_src, int sw, int sh, size_t src_pitch, const float *d_xmap, const float *d_ymap,
                          size_t map_pitch, uint8_t *d_dst, int dw, int dh, size_t dst_pitch, void *stream)
{
    VSB_REQUIRE(d_src && d_xmap && d_ymap && d_dst && sw > 0 && sh > 0 && dw > 0 && dh > 0, "remap: bad arguments");
    const dim3 b(32, 8);
    k_remap_linear_u8c3<<<grid2d(dw, dh, b), b, 0, (cudaStream_t)stream>>>(d_src, sw, sh, src_pitch, d_xmap, d_ymap, map_pitch, d_dst, dw, dh, dst_pitch);
    return vsb::check_launch("k_remap_linear_u8c3");
}

int vsb_gain_u8(uint8_t *d_buf, int width_bytes, int h, size_t pitch, float gain, void *stream)
{
    VSB_REQUIRE(d_buf && width_bytes > 0 && h > 0, "gain: bad arguments");
    const dim3 b(32, 8);
    k_gain_u8<<<grid2d(width_bytes, h, b), b, 0, (cudaStream_t)stream>>>(d_buf, width_bytes, h, pitch, gain);
    return vsb::check_launch("k_gain_u8");
}

int vsb_border_reflect_u8c3_to_s16c3(const uint8_t *d_src, int w, int h, size_t src_pitch, int top, int bottom, int left, int right,
                                     int16_t *d_dst, size_t dst_pitch, void *stream)
{
    VSB_REQUIRE(d_src && d_dst && w > 0 && h > 0 && top >= 0 && bottom >= 0 && left >= 0 && right >= 0, "border: bad arguments");
    const dim3 b(32, 8);
    const int dw = w + left + right, dh = h + top + bottom;
    k_border_reflect_u8c3_s16c3<<<grid2d(dw, dh, b), b, 0, (cudaStream_t)stream>>>(d_src, w, h, src_pitch, top, left, d_dst, dw, dh, dst_pitch);
    return vsb::check_launch("k_border_reflect_u8c3_s16c3");
}

int vsb_pyr_down_s16c3(const int16_t *d_src, int w, int h, size_t src_pitch, int16_t *d_dst, size_t dst_pitch, void *stream)
{
    VSB_REQUIRE(d_src && d_dst && w > 0 && h > 0, "pyrDown: bad arguments");
    const dim3 b(32, 8);
    const int dw = (w + 1) / 2, dh = (h + 1) / 2;
    k_pyr_down_s16c3<<<grid2d(dw, dh, b), b, 0, (cudaStream_t)stream>>>(d_src, w, h, src_pitch, d_dst, dw, dh, dst_pitch);
    return vsb::check_launch("k_pyr_down_s16c3");
}

int vsb_pyr_up_s16c3(const int16_t *d_src, int w, int h, size_t src_pitch, int16_t *d_dst, size_t dst_pitch, void *stream)
{
    VSB_REQUIRE(d_src && d_dst && w > 0 && h > 0, "pyrUp: bad arguments");
    const dim3 b(32, 8);
    k_pyr_up_s16c3<<<grid2d(2 * w, 2 * h, b), b, 0, (cudaStream_t)stream>>>(d_src, w, h, src_pitch, d_dst, dst_pitch);
    return vsb::check_launch("k_pyr_up_s16c3");
}

int vsb_pyr_down_f32(const float *d_src, int w, int h, size_t src_pitch, float *d_dst, size_t dst_pitch, void *stream)
{
    VSB_REQUIRE(d_src && d_dst && w > 0 && h > 0, "pyrDown f32: bad arguments");
    const dim3 b(32, 8);
    const int dw = (w + 1) / 2, dh = (h + 1) / 2;
    k_pyr_down_f32<<<grid2d(dw, dh, b), b, 0, (cudaStream_t)stream>>>(d_src, w, h, src_pitch, d_dst, dw, dh, dst_pitch);
    return vsb::check_launch("k_pyr_down_f32");
}

int vsb_add_src_weight_32f(const int16_t *d_src, size_t src_pitch, const float *d_w, size_t w_pitch, int16_t *d_dst, size_t dst_pitch,
                           float *d_dst_w, size_t dst_w_pitch, int w, int h, void *stream)
{
    VSB_REQUIRE(d_src && d_w && d_dst && d_dst_w && w > 0 && h > 0, "addSrcWeight: bad arguments");
    const dim3 b(32, 8);
    k_add_src_weight_32f<<<grid2d(w, h, b), b, 0, (cudaStream_t)stream>>>(d_src, src_pitch, d_w, w_pitch, d_dst, dst_pitch, d_dst_w, dst_w_pitch, w, h);
    return vsb::check_launch("k_add_src_weight_32f");
}

int vsb_normalize_32f(const float *d_w, size_t w_pitch, int16_t *d_src, size_t src_pitch, int w, int h, void *stream)
{
    VSB_REQUIRE(d_w && d_src && w > 0 && h > 0, "normalize: bad arguments");
    const dim3 b(32, 8);
    k_normalize_32f<<<grid2d(w, h, b), b, 0, (cudaStream_t)stream>>>(d_w, w_pitch, d_src, src_pitch, w, h);
    return vsb::check_launch("k_normalize_32f");
}

int vsb_custom_resize(const float *d_in, int cols, int rows, size_t in_pitch_bytes, float *d_out, int tx, int ty, size_t out_pitch_bytes, void *stream)
{
    VSB_REQUIRE(d_in && d_out && cols > 1 && rows > 1 && tx > 0 && ty > 0, "custom_resize: bad arguments");
    const dim3 b(32, 8);
    k_custom_resize<<<grid2d(tx, ty, b), b, 0, (cudaStream_t)stream>>>(d_in, cols, rows, in_pitch_bytes, d_out, tx, ty, out_pitch_bytes);
    return vsb::check_launch("k_custom_resize");
}

int vsb_build_maps(int projection, float scale, const float K[9], const float R[9], int src_w, int src_h, float *d_xmap, float *d_ymap,
                   size_t pitch_bytes, int roi[4], void *stream)
{
    VSB_REQUIRE(K && R && d_xmap && d_ymap && roi, "build_maps: bad arguments");
    int r = vsb_warp_roi(projection, scale, K, R, src_w, src_h, roi);
    if (r != VSB_OK) return r;
    Mat3 k;
    float r_kinv[9], rinv[9];
    vsb::projector_setup(K, R, k.m, r_kinv, rinv);
    const dim3 b(32, 8);
    k_build_maps<<<grid2d(roi[2], roi[3], b), b, 0, (cudaStream_t)stream>>>(projection, scale, k, roi[0], roi[1], roi[2], roi[3], d_xmap, d_ymap, pitch_bytes);
    return vsb::check_launch("k_build_maps");
}

int vsb_warp(int projection, float scale, const float K[9], const float R[9], const uint8_t *d_src, int src_w, int src_h, size_t src_pitch,
             int channels, int interp, int border, uint8_t *d_dst, size_t dst_pitch, int roi[4], void *stream)
{
    VSB_REQUIRE(K && R && d_src && d_dst && roi, "warp: bad arguments");
    VSB_REQUIRE(channels == 1 || channels == 3, "warp: CV_8UC1 or CV_8UC3 only");
    VSB_REQUIRE((interp == VSB_INTER_NEAREST || interp == VSB_INTER_LINEAR) && (border == VSB_BORDER_CONSTANT || border == VSB_BORDER_REFLECT),
                "warp: interp must be NEAREST / LINEAR, border CONSTANT / REFLECT");
    int r = vsb_warp_roi(projection, scale, K, R, src_w, src_h, roi);
    if (r != VSB_OK) return r;
    VSB_REQUIRE(src_pitch >= (size_t)src_w * channels && dst_pitch >= (size_t)roi[2] * channels, "warp: pitch too small");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t mp = ((size_t)roi[2] * 4 + 15) / 16 * 16;
    float *maps = nullptr;  // d_xmap_ / d_ymap_ of the warper object (warpers.hpp:547-549): scratch here, stream-ordered
    cudaError_t e = cudaMallocAsync(&maps, 2 * mp * roi[3], st);
    if (e != cudaSuccess) return vsb::check_cuda(e, "warp: map scratch");
    float *xm = maps, *ym = (float *)((char *)maps + mp * roi[3]);
    Mat3 k;
    float r_kinv[9], rinv[9];
    vsb::projector_setup(K, R, k.m, r_kinv, rinv);
    const dim3 b(32, 8);
    k_build_maps<<<grid2d(roi[2], roi[3], b), b, 0, st>>>(projection, scale, k, roi[0], roi[1], roi[2], roi[3], xm, ym, mp);
    if (channels == 1) k_warp_remap<1><<<grid2d(roi[2], roi[3], b), b, 0, st>>>(d_src, src_w, src_h, src_pitch, xm, ym, mp, d_dst, roi[2], roi[3], dst_pitch, interp, border);
    else k_warp_remap<3><<<grid2d(roi[2], roi[3], b), b, 0, st>>>(d_src, src_w, src_h, src_pitch, xm, ym, mp, d_dst, roi[2], roi[3], dst_pitch, interp, border);
    r = vsb::check_launch("k_warp_remap");
    cudaFreeAsync(maps, st);
    return r;
}

}  // extern "C"
