/*
 * oracle_g.c -- TEST INFRASTRUCTURE ONLY.  See oracle_g.h.
 *
 * Scalar CPU restatement ("oracle-G") of the arithmetic of the reference's CUDA compose path.
 * Paths below are relative to /root/reference; abbreviations:
 *   A/  = 360_stitcher/                      S/  = sources/modules/stitching/
 *   CW/ = sources/modules/cudawarping/       CA/ = sources/modules/cudaarithm/
 *   CORE/ = sources/modules/core/            IMG/ = sources/modules/imgproc/
 *
 * Floating-point contract: compiled with -ffp-contract=off; every place where nvcc's default
 * -fmad=true would contract "a*b + c" in the reference kernel is written as an explicit fmaf(),
 * every other operation is a separately rounded fp32 op.  The s16 pyramids are exact in fp32
 * (see og_pyr_down_s16_int) so contraction is immaterial there.
 *
 * Parity pin status: PINNED.  The reference ships no vendored golden vectors for this path, so the pin
 * is the reference's own code run here: (1) oracle/_ref = the reference's vendored OpenCV 3.4.0 CPU sources compiled in
 * place by oracle/ref.mk when /root/reference is present; its outputs are committed as tests/golden/reference_cpu.npz
 * (generator: tests/golden/make_golden.py) and checked live and from the fixture by tests/test_oracle_pin.py;
 * (2) the reference's float gold for cuda::remap (CW/test/interpolation.hpp:66-84, compiled into oracle/ref_shim.cpp)
 * on the CW/test/test_remap.cpp:158-177 recipe, and for cuda::resize (CW/test/test_resize.cpp:54-74) on that test's recipe;
 * (3) the reference's own CUDA KERNELS of the path -- cuda::remap, pyrDown, pyrUp, cuda::resize, copyMakeBorder, convertTo (gain),
 * addSrcWeightKernel32F, normalizeUsingWeightKernel32F, the application's resize -- compiled to PTX from the reference's unmodified .cu files (oracle/ref_ptx.mk) and executed on the CPU with
 * exact binary32 arithmetic (oracle/ptx_interp.py): oracle-G equals them bit for bit (tests/test_oracle_ptx.py, fixture
 * tests/golden/reference_ptx.npz); (4) replayed recipes of the reference's own tests
 * (CW/test/test_pyramids.cpp, S/test/test_blenders.cpp).
 */
#include "oracle_g.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;
void og_set_num_threads(int n) { g_threads = n > 0 ? n : 1; }

/* ---------------------------------------------------------------- rounding helpers */

/* cvt.rni.sat.u8.f32 (CORE/include/opencv2/core/cuda/saturate_cast.hpp:96-101): NaN -> 0 */
static inline uint8_t rni_sat_u8(float v)
{
    if (!(v == v)) return 0;
    if (v <= 0.f) return 0;
    if (v >= 255.f) return 255;
    return (uint8_t)lrintf(v); /* default rounding mode: nearest-even */
}

/* cvt.rni.sat.s16.f32 (saturate_cast.hpp:221-226) */
static inline int16_t rni_sat_s16(float v)
{
    if (!(v == v)) return 0;
    if (v <= -32768.f) return -32768;
    if (v >= 32767.f) return 32767;
    return (int16_t)lrintf(v);
}

/* static_cast<short>(float) in device code = cvt.rzi.s16.f32 (saturating, NaN -> 0) */
static inline int16_t rz_s16(float v)
{
    if (!(v == v)) return 0;
    if (v <= -32768.f) return -32768;
    if (v >= 32767.f) return 32767;
    return (int16_t)v;
}

/* __float2int_rd : cvt.rmi.s32.f32 (saturating, NaN -> 0) */
static inline int f2i_rd(float v)
{
    if (!(v == v)) return 0;
    if (v <= -2147483648.f) return INT_MIN;
    if (v >= 2147483648.f) return INT_MAX;
    return (int)floorf(v);
}

/* __float2int_rz */
static inline int f2i_rz(float v)
{
    if (!(v == v)) return 0;
    if (v <= -2147483648.f) return INT_MIN;
    if (v >= 2147483648.f) return INT_MAX;
    return (int)v;
}

static inline int16_t sat_s16_i(int v) { return (int16_t)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

/* ---------------------------------------------------------------- camera rig + projector */

/* A/calibration.cpp:28-68 calibrateCameras, generalised to n views (yaw = 2*pi*i/n), work_scale=1 */
void og_rig_camera(int n_views, int i, int src_w, int src_h, double hfov_deg, float K[9], float R[9])
{
    const double PI = 3.1415926535897932384626; /* A/defs.h:76 */
    double fov = hfov_deg * PI / 180.0;
    double focal_tmp = 1.0 / tan(fov * 0.5);
    float rot = (float)(2.0 * PI * (float)i / n_views);
    double ppx = src_w / 2.0, ppy = src_h / 2.0;
    double focal = focal_tmp * ppx;
    /* CameraParams::K() (S/src/camera.cpp) then convertTo(CV_32F) */
    K[0] = (float)focal; K[1] = 0.f; K[2] = (float)ppx;
    K[3] = 0.f; K[4] = (float)(focal * 1.0); K[5] = (float)ppy;
    K[6] = 0.f; K[7] = 0.f; K[8] = 1.f;
    /* Ry only (Rz = Rx = I): A/calibration.cpp:42-52 */
    R[0] = (float)cos(rot); R[1] = 0.f; R[2] = (float)sin(rot);
    R[3] = 0.f; R[4] = 1.f; R[5] = 0.f;
    R[6] = (float)-sin(rot); R[7] = 0.f; R[8] = (float)cos(rot);
}

/* the same camera after  cameras[i].focal *= compose_work_aspect; ppx *= ...; ppy *= ...  (A/calibration.cpp:168-172; doubles,
 * work_scale = 1 so compose_work_aspect = compose_scale), then K().convertTo(CV_32F) (:175-176) */
void og_rig_camera_scaled(int n_views, int i, int src_w, int src_h, double hfov_deg, double compose_work_aspect, float K[9], float R[9])
{
    const double PI = 3.1415926535897932384626;
    og_rig_camera(n_views, i, src_w, src_h, hfov_deg, K, R);
    double ppx = src_w / 2.0, ppy = src_h / 2.0;
    double focal = (1.0 / tan(hfov_deg * PI / 180.0 * 0.5)) * ppx;
    focal *= compose_work_aspect; ppx *= compose_work_aspect; ppy *= compose_work_aspect;
    K[0] = (float)focal; K[2] = (float)ppx; K[4] = (float)(focal * 1.0); K[5] = (float)ppy;
}

/* calibrateCameras at a work scale (A/calibration.cpp:54-60: ppx = (full.width * work_scale) / 2, focal = focal_tmp * ppx, doubles),
 * then focal, ppx, ppy *= aspect (:168-172: compose_work_aspect = compose_scale / work_scale; 1 for the work-scale camera itself) and
 * K().convertTo(CV_32F).  work_scale = aspect = 1 gives og_rig_camera; work_scale = 1 gives og_rig_camera_scaled. */
void og_rig_camera_work(int n_views, int i, int src_w, int src_h, double hfov_deg, double work_scale, double aspect, float K[9], float R[9])
{
    const double PI = 3.1415926535897932384626;
    og_rig_camera(n_views, i, src_w, src_h, hfov_deg, K, R);
    double ppx = (src_w * work_scale) / 2.0, ppy = (src_h * work_scale) / 2.0;
    double focal = (1.0 / tan(hfov_deg * PI / 180.0 * 0.5)) * ppx;
    focal *= aspect; ppx *= aspect; ppy *= aspect;
    K[0] = (float)focal; K[2] = (float)ppx; K[4] = (float)(focal * 1.0); K[5] = (float)ppy;
}

static void mat3_mul_d(const float *a, const float *b, float *c)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += (double)a[i * 3 + k] * (double)b[k * 3 + j];
            c[i * 3 + j] = (float)s;
        }
}

/* ProjectorBase::setCameraParams, S/src/warpers.cpp:49-79 (products accumulated in double as cv::gemm does) */
void og_projector(const float K[9], const float R[9], float k_rinv[9], float r_kinv[9], float rinv[9])
{
    float Rt[9], Kinv[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Rt[i * 3 + j] = R[j * 3 + i];
    double k[9];
    for (int i = 0; i < 9; ++i) k[i] = K[i];
    double det = k[0] * (k[4] * k[8] - k[5] * k[7]) - k[1] * (k[3] * k[8] - k[5] * k[6]) + k[2] * (k[3] * k[7] - k[4] * k[6]);
    double d = 1.0 / det;
    Kinv[0] = (float)((k[4] * k[8] - k[5] * k[7]) * d);
    Kinv[1] = (float)((k[2] * k[7] - k[1] * k[8]) * d);
    Kinv[2] = (float)((k[1] * k[5] - k[2] * k[4]) * d);
    Kinv[3] = (float)((k[5] * k[6] - k[3] * k[8]) * d);
    Kinv[4] = (float)((k[0] * k[8] - k[2] * k[6]) * d);
    Kinv[5] = (float)((k[2] * k[3] - k[0] * k[5]) * d);
    Kinv[6] = (float)((k[3] * k[7] - k[4] * k[6]) * d);
    Kinv[7] = (float)((k[1] * k[6] - k[0] * k[7]) * d);
    Kinv[8] = (float)((k[0] * k[4] - k[1] * k[3]) * d);
    memcpy(rinv, Rt, sizeof(Rt));
    mat3_mul_d(R, Kinv, r_kinv);
    mat3_mul_d(K, Rt, k_rinv);
}

/* SphericalProjector::mapForward / CylindricalProjector::mapForward, S/include/opencv2/stitching/detail/warpers_inl.hpp:243-253,274-283 */
static void map_forward(int proj, float scale, const float *r_kinv, float x, float y, float *u, float *v)
{
    float x_ = r_kinv[0] * x + r_kinv[1] * y + r_kinv[2];
    float y_ = r_kinv[3] * x + r_kinv[4] * y + r_kinv[5];
    float z_ = r_kinv[6] * x + r_kinv[7] * y + r_kinv[8];
    if (proj == OG_PROJ_SPHERICAL) {
        *u = scale * atan2f(x_, z_);
        float w = y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_);
        *v = scale * ((float)M_PI - acosf(w == w ? w : 0));
    } else {
        *u = scale * atan2f(x_, z_);
        *v = scale * y_ / sqrtf(x_ * x_ + z_ * z_);
    }
}

/* detectResultRoiByBorder (warpers_inl.hpp:176-210) + SphericalWarper::detectResultRoi pole test (S/src/warpers.cpp:277-318) */
void og_warp_roi(int proj, float scale, const float K[9], const float R[9], int src_w, int src_h, int roi[4])
{
    float k_rinv[9], r_kinv[9], rinv[9];
    og_projector(K, R, k_rinv, r_kinv, rinv);
    float tl_uf = 3.402823466e+38f, tl_vf = 3.402823466e+38f, br_uf = -3.402823466e+38f, br_vf = -3.402823466e+38f;
    float u, v;
#define ACC() do { tl_uf = fminf(tl_uf, u); tl_vf = fminf(tl_vf, v); br_uf = fmaxf(br_uf, u); br_vf = fmaxf(br_vf, v); } while (0)
    for (float x = 0; x < src_w; ++x) {
        map_forward(proj, scale, r_kinv, x, 0, &u, &v); ACC();
        map_forward(proj, scale, r_kinv, x, (float)(src_h - 1), &u, &v); ACC();
    }
    for (int y = 0; y < src_h; ++y) {
        map_forward(proj, scale, r_kinv, 0, (float)y, &u, &v); ACC();
        map_forward(proj, scale, r_kinv, (float)(src_w - 1), (float)y, &u, &v); ACC();
    }
#undef ACC
    int tlx = (int)tl_uf, tly = (int)tl_vf, brx = (int)br_uf, bry = (int)br_vf;
    if (proj == OG_PROJ_SPHERICAL) {
        tl_uf = (float)tlx; tl_vf = (float)tly; br_uf = (float)brx; br_vf = (float)bry;
        for (int pass = 0; pass < 2; ++pass) {
            float x = rinv[1], y = pass == 0 ? rinv[4] : -rinv[4], z = rinv[7];
            if (y > 0.f) {
                float x_ = (K[0] * x + K[1] * y) / z + K[2];
                float y_ = K[4] * y / z + K[5];
                if (x_ > 0.f && x_ < src_w && y_ > 0.f && y_ < src_h) {
                    float pole = pass == 0 ? (float)(M_PI * scale) : 0.f;
                    tl_uf = fminf(tl_uf, 0.f); tl_vf = fminf(tl_vf, pole);
                    br_uf = fmaxf(br_uf, 0.f); br_vf = fmaxf(br_vf, pole);
                }
            }
        }
        tlx = (int)tl_uf; tly = (int)tl_vf; brx = (int)br_uf; bry = (int)br_vf;
    }
    roi[0] = tlx; roi[1] = tly; roi[2] = brx - tlx + 1; roi[3] = bry - tly + 1;
}

/* SphericalMapper / CylindricalMapper::mapBackward + buildWarpMapsKernel, S/src/cuda/build_warp_maps.cu:88-152.
 * What nvcc makes of  x = k[0]*x_ + k[1]*y_ + k[2]*z_  (PTX of the reference's own file, oracle/ref_ptx.mk; executed by
 * tests/test_oracle_ptx.py): spherical, where y_ = -cos(v): the negation is folded into a SUBTRACTION of two separately rounded
 * products and only the third term is fused; cylindrical: the second product is rounded on its own, the first and the third
 * are fused onto it.  sinf / cosf are the host libm's here (the device's differ by ulps: map parity is a 2e-3 px bound). */
void og_build_maps(int proj, float scale, const float K[9], const float R[9], int tl_x, int tl_y,
                   int w, int h, float *xmap, float *ymap)
{
    float k[9], r_kinv[9], rinv[9];
    og_projector(K, R, k, r_kinv, rinv);
#pragma omp parallel for num_threads(g_threads)
    for (int dv = 0; dv < h; ++dv) {
        for (int du = 0; du < w; ++du) {
            float u = (float)(tl_x + du), v = (float)(tl_y + dv);
            float x, y, z;
            if (proj == OG_PROJ_SPHERICAL) {
                v = v / scale; u = u / scale;
                const float sinv = sinf(v);
                const float x_ = sinv * sinf(u);
                const float cosv = cosf(v);
                const float z_ = sinv * cosf(u);
                const float x0 = x_ * k[0], x1 = cosv * k[1], y0 = x_ * k[3], y1 = cosv * k[4], z0 = x_ * k[6], z1 = cosv * k[7];
                x = fmaf(k[2], z_, x0 - x1);
                y = fmaf(k[5], z_, y0 - y1);
                z = fmaf(k[8], z_, z0 - z1);
            } else {
                u = u / scale;
                const float x_ = sinf(u);
                const float y_ = v / scale;
                const float z_ = cosf(u);
                x = fmaf(k[2], z_, fmaf(x_, k[0], y_ * k[1]));
                y = fmaf(k[5], z_, fmaf(x_, k[3], y_ * k[4]));
                z = fmaf(k[8], z_, fmaf(x_, k[6], y_ * k[7]));
            }
            if (z > 0) { x = x / z; y = y / z; } else { x = y = -1.f; }
            xmap[(size_t)dv * w + du] = x;
            ymap[(size_t)dv * w + du] = y;
        }
    }
}

/* ---------------------------------------------------------------- remap / gain / resize */

/* cuda::remap LINEAR + BORDER_CONSTANT(0): CW/src/cuda/remap.cu:56-68; LinearFilter CORE/include/opencv2/core/cuda/filters.hpp:90-114;
 * BorderReader<BrdConstant> CORE/include/opencv2/core/cuda/border_interpolate.hpp:708-711 */
void og_remap_linear_u8(const uint8_t *src, int sw, int sh, int cn, size_t sstep,
                        const float *xmap, const float *ymap, size_t mstep,
                        uint8_t *dst, int dw, int dh, size_t dstep)
{
#pragma omp parallel for num_threads(g_threads)
    for (int yy = 0; yy < dh; ++yy) {
        for (int xx = 0; xx < dw; ++xx) {
            float x = xmap[(size_t)yy * mstep + xx], y = ymap[(size_t)yy * mstep + xx];
            int x1 = f2i_rd(x), y1 = f2i_rd(y);
            int x2 = x1 + 1, y2 = y1 + 1;
            float w11 = ((float)x2 - x) * ((float)y2 - y);
            float w12 = (x - (float)x1) * ((float)y2 - y);
            float w21 = ((float)x2 - x) * (y - (float)y1);
            float w22 = (x - (float)x1) * (y - (float)y1);
            int inx1 = x1 >= 0 && x1 < sw, inx2 = x2 >= 0 && x2 < sw;
            int iny1 = y1 >= 0 && y1 < sh, iny2 = y2 >= 0 && y2 < sh;
            for (int c = 0; c < cn; ++c) {
                float s11 = (inx1 && iny1) ? (float)src[(size_t)y1 * sstep + (size_t)x1 * cn + c] : 0.f;
                float s12 = (inx2 && iny1) ? (float)src[(size_t)y1 * sstep + (size_t)x2 * cn + c] : 0.f;
                float s21 = (inx1 && iny2) ? (float)src[(size_t)y2 * sstep + (size_t)x1 * cn + c] : 0.f;
                float s22 = (inx2 && iny2) ? (float)src[(size_t)y2 * sstep + (size_t)x2 * cn + c] : 0.f;
                float out = fmaf(s11, w11, 0.f);
                out = fmaf(s12, w12, out);
                out = fmaf(s21, w21, out);
                out = fmaf(s22, w22, out);
                dst[(size_t)yy * dstep + (size_t)xx * cn + c] = rni_sat_u8(out);
            }
        }
    }
}

/* cuda::remap NEAREST + BORDER_CONSTANT(0): PointFilter filters.hpp:64-78 (__float2int_rz) */
void og_remap_nearest_u8c1(const uint8_t *src, int sw, int sh, size_t sstep,
                           const float *xmap, const float *ymap, size_t mstep,
                           uint8_t *dst, int dw, int dh, size_t dstep)
{
    for (int yy = 0; yy < dh; ++yy)
        for (int xx = 0; xx < dw; ++xx) {
            int x = f2i_rz(xmap[(size_t)yy * mstep + xx]), y = f2i_rz(ymap[(size_t)yy * mstep + xx]);
            dst[(size_t)yy * dstep + xx] = (x >= 0 && x < sw && y >= 0 && y < sh) ? src[(size_t)y * sstep + x] : 0;
        }
}

/* cuda::remap as the warpers call it (RotationWarperGpu::warp, S/src/warpers_cuda.cpp:279-298 -> CW/src/cuda/remap.cu:56-68):
 * interp 0 = PointFilter (filters.hpp:64-78, __float2int_rz), 1 = LinearFilter (:90-114); border 0 = BrdConstant(0)
 * (border_interpolate.hpp:698-717), 2 = BrdReflect (:484-534: idx_low(idx_high(i)), low = (|i| - (i < 0)) % len,
 * high = last - |last - i| + (i > last)).  The application uses LINEAR/REFLECT for the seam-scale images and
 * NEAREST/CONSTANT for the masks (360_stitcher/calibration.cpp:118,122,227). */
static int brd_reflect(int i, int len)
{
    const int last = len - 1;
    int j = last - abs(last - i) + (i > last);
    return (abs(j) - (j < 0)) % len;
}
static float read_brd(const uint8_t *src, int sw, int sh, int cn, size_t sstep, int y, int x, int c, int border)
{
    if (border == 0) return (x >= 0 && x < sw && y >= 0 && y < sh) ? (float)src[(size_t)y * sstep + (size_t)x * cn + c] : 0.f;
    return (float)src[(size_t)brd_reflect(y, sh) * sstep + (size_t)brd_reflect(x, sw) * cn + c];
}
void og_remap_u8_border(const uint8_t *src, int sw, int sh, int cn, size_t sstep, const float *xmap, const float *ymap, size_t mstep,
                        uint8_t *dst, int dw, int dh, size_t dstep, int interp, int border)
{
#pragma omp parallel for num_threads(g_threads)
    for (int yy = 0; yy < dh; ++yy)
        for (int xx = 0; xx < dw; ++xx) {
            const float x = xmap[(size_t)yy * mstep + xx], y = ymap[(size_t)yy * mstep + xx];
            for (int c = 0; c < cn; ++c) {
                uint8_t *d = dst + (size_t)yy * dstep + (size_t)xx * cn + c;
                if (interp == 0) {
                    *d = (uint8_t)read_brd(src, sw, sh, cn, sstep, f2i_rz(y), f2i_rz(x), c, border);
                    continue;
                }
                const int x1 = f2i_rd(x), y1 = f2i_rd(y), x2 = x1 + 1, y2 = y1 + 1;
                float out = fmaf(read_brd(src, sw, sh, cn, sstep, y1, x1, c, border), ((float)x2 - x) * ((float)y2 - y), 0.f);
                out = fmaf(read_brd(src, sw, sh, cn, sstep, y1, x2, c, border), (x - (float)x1) * ((float)y2 - y), out);
                out = fmaf(read_brd(src, sw, sh, cn, sstep, y2, x1, c, border), ((float)x2 - x) * (y - (float)y1), out);
                out = fmaf(read_brd(src, sw, sh, cn, sstep, y2, x2, c, border), (x - (float)x1) * (y - (float)y1), out);
                *d = rni_sat_u8(out);
            }
        }
}

/* GpuMat::convertTo(type, alpha): Convertor, CORE/src/cuda/gpu_mat.cu:488-498 (A/timed.cpp:94) */
void og_gain_u8(uint8_t *buf, size_t n, float gain)
{
    uint8_t lut[256];
    for (int i = 0; i < 256; ++i) lut[i] = rni_sat_u8(fmaf(gain, (float)i, 0.f));
    for (size_t i = 0; i < n; ++i) buf[i] = lut[buf[i]];
}

/* The capture boards send NV12; the reference converts every received frame with cv::cvtColor(mat, mat, CV_YUV2BGR_NV12)
 * (360_stitcher/networking.cpp:46).  Integer BT.601 arithmetic of YUV420sp2RGB888Invoker<bIdx = 0, uIdx = 0>:
 * sources/modules/imgproc/src/color.cpp:8741-8746 (constants), :8793-8818 (per 2x2 block). */
static uint8_t sat_u8_int(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
void og_nv12_to_bgr(const uint8_t *nv12, int w, int h, size_t step, uint8_t *bgr, size_t bgr_step)
{
    const int CY = 1220542, CUB = 2116026, CUG = -409993, CVG = -852492, CVR = 1673527, SHIFT = 20;
    const uint8_t *uvp = nv12 + step * (size_t)h;
    for (int j = 0; j < h; ++j) {
        const uint8_t *yr = nv12 + step * (size_t)j, *uv = uvp + step * (size_t)(j / 2);
        uint8_t *o = bgr + bgr_step * (size_t)j;
        for (int i = 0; i < w; ++i) {
            const int u = (int)uv[(i & ~1)] - 128, v = (int)uv[(i & ~1) + 1] - 128;
            const int ruv = (1 << (SHIFT - 1)) + CVR * v;
            const int guv = (1 << (SHIFT - 1)) + CVG * v + CUG * u;
            const int buv = (1 << (SHIFT - 1)) + CUB * u;
            const int yy = ((int)yr[i] - 16 > 0 ? (int)yr[i] - 16 : 0) * CY;
            o[3 * i + 0] = sat_u8_int((yy + buv) >> SHIFT);
            o[3 * i + 1] = sat_u8_int((yy + guv) >> SHIFT);
            o[3 * i + 2] = sat_u8_int((yy + ruv) >> SHIFT);
        }
    }
}

/* consumer: mat.convertTo(mat_8u, CV_8U) before the download (360_stitcher/timed.cpp:250); cuda convertTo without scale is
 * saturate_cast<uchar>(short) (sources/modules/core/include/opencv2/core/cuda/saturate_cast.hpp) */
void og_s16_to_u8(const int16_t *src, size_t n, uint8_t *dst)
{
    for (size_t i = 0; i < n; ++i) dst[i] = sat_u8_int(src[i]);
}

/* ---- consumer epilogue (CPU code in the reference, downstream of the download: 360_stitcher/timed.cpp:254-315) ---- */
/* coefficient tables of cv::hal::resize for INTER_LINEAR on 8U: IMG/src/resize.cpp:3933-3958 (x), 3991-4016 (y) */
static void resize_tables(int ssize, int dsize, int clamp_x, int *ofs, short *coef)
{
    const double inv_scale = (double)dsize / ssize;   /* cv::resize: inv_scale_x = (double)dsize.width / ssize.width */
    const double scale = 1. / inv_scale;              /* hal::resize: scale_x = 1. / inv_scale_x */
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (clamp_x) {
            if (s < 0) { f = 0; s = 0; }
            if (s >= ssize - 1) { f = 0; s = ssize - 1; }
        }
        ofs[d] = s;
        float c0 = 1.f - f, c1 = f;
        long r0 = lrintf(c0 * 2048), r1 = lrintf(c1 * 2048);   /* saturate_cast<short>(float) = cvRound, round-half-even */
        coef[2 * d] = (short)(r0 > 32767 ? 32767 : (r0 < -32768 ? -32768 : r0));
        coef[2 * d + 1] = (short)(r1 > 32767 ? 32767 : (r1 < -32768 ? -32768 : r1));
    }
}
void og_resize_linear_u8c3(const uint8_t *src, int sw, int sh, size_t sstep, uint8_t *dst, int dw, int dh, size_t dstep)
{
    int *xofs = (int *)malloc(sizeof(int) * (dw + dh)), *yofs = xofs + dw;
    short *ia = (short *)malloc(sizeof(short) * 2 * (dw + dh)), *ib = ia + 2 * dw;
    resize_tables(sw, dw, 1, xofs, ia);
    resize_tables(sh, dh, 0, yofs, ib);
#pragma omp parallel for
    for (int dy = 0; dy < dh; ++dy) {
        /* resizeGeneric_Invoker: rows sy0 - ksize2 + 1 + k clipped into the image (IMG/src/resize.cpp:2200-2240) */
        int s0 = yofs[dy], s1 = yofs[dy] + 1;
        s0 = s0 < 0 ? 0 : (s0 >= sh ? sh - 1 : s0);
        s1 = s1 < 0 ? 0 : (s1 >= sh ? sh - 1 : s1);
        const uint8_t *r0 = src + sstep * (size_t)s0, *r1 = src + sstep * (size_t)s1;
        const int b0 = ib[2 * dy], b1 = ib[2 * dy + 1];
        for (int dx = 0; dx < dw; ++dx) {
            const int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
            const int a0 = ia[2 * dx], a1 = ia[2 * dx + 1];
            for (int c = 0; c < 3; ++c) {
                const int S0 = r0[sx * 3 + c] * a0 + r0[sx1 * 3 + c] * a1;   /* HResizeLinear (:1923-1941) */
                const int S1 = r1[sx * 3 + c] * a0 + r1[sx1 * 3 + c] * a1;
                dst[dstep * (size_t)dy + dx * 3 + c] = (uint8_t)((((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2);  /* VResizeLinear (:2013) */
            }
        }
    }
    free(xofs); free(ia);
}

void og_bgr_to_i420(const uint8_t *bgr, int w, int h, size_t step, uint8_t *yuv)
{
    const int CRY = 269484, CGY = 528482, CBY = 102760, CRU = -155188, CGU = -305135, CBU = 460324, CGV = -385875, CBV = -74448;
    const int SHIFT = 20, half = 1 << (SHIFT - 1), s16 = 16 << SHIFT, s128 = 128 << SHIFT;
    uint8_t *yp = yuv, *up = yuv + (size_t)w * h, *vp = up + (size_t)(w / 2) * (h / 2);
    for (int j = 0; j < h; ++j)
        for (int i = 0; i < w; ++i) {
            const uint8_t *p = bgr + step * (size_t)j + 3 * i;
            const int b = p[0], g = p[1], r = p[2];
            yp[(size_t)j * w + i] = sat_u8_int((CRY * r + CGY * g + CBY * b + half + s16) >> SHIFT);
            if (!(j & 1) && !(i & 1)) {
                up[(size_t)(j / 2) * (w / 2) + i / 2] = sat_u8_int((CRU * r + CGU * g + CBU * b + half + s128) >> SHIFT);
                vp[(size_t)(j / 2) * (w / 2) + i / 2] = sat_u8_int((CBU * r + CGV * g + CBV * b + half + s128) >> SHIFT);
            }
        }
}

int og_consumer_image_height(int src_w, int src_h, int out_w, int out_h, int keep_aspect)
{
    if (!keep_aspect) return out_h;
    int ih = (int)((double)out_w / (double)src_w * src_h + 0.5);   /* timed.cpp:260 */
    return ih > out_h ? out_h : ih;
}

/* cuda::resize INTER_LINEAR on CV_8UC1: CW/src/cuda/resize.cu:71-106, host CW/src/resize.cpp:76-105 */
void og_resize_linear_u8c1(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh)
{
    if (dw == sw && dh == sh) { memcpy(dst, src, (size_t)sw * sh); return; }
    double fxd = (double)dw / sw, fyd = (double)dh / sh;
    float fx = (float)(1.0 / fxd), fy = (float)(1.0 / fyd);
    for (int dy = 0; dy < dh; ++dy)
        for (int dx = 0; dx < dw; ++dx) {
            float src_x = (float)dx * fx, src_y = (float)dy * fy;
            int x1 = f2i_rd(src_x), y1 = f2i_rd(src_y);
            int x2 = x1 + 1, y2 = y1 + 1;
            int x2r = x2 < sw - 1 ? x2 : sw - 1, y2r = y2 < sh - 1 ? y2 : sh - 1;
            float out = fmaf((float)src[(size_t)y1 * sw + x1], ((float)x2 - src_x) * ((float)y2 - src_y), 0.f);
            out = fmaf((float)src[(size_t)y1 * sw + x2r], (src_x - (float)x1) * ((float)y2 - src_y), out);
            out = fmaf((float)src[(size_t)y2r * sw + x1], ((float)x2 - src_x) * (src_y - (float)y1), out);
            out = fmaf((float)src[(size_t)y2r * sw + x2r], (src_x - (float)x1) * (src_y - (float)y1), out);
            dst[(size_t)dy * dw + dx] = rni_sat_u8(out);
        }
}

/* cuda::resize INTER_LINEAR on CV_8UC1 / CV_8UC3 with explicit scale factors (CW/src/resize.cpp:76-105, CW/src/cuda/resize.cu:71-106):
 * fx = fy = 0 derives them from the sizes (dsize given); the application resizes the camera frames to seam scale with
 * cuda::resize(img, seam_img, Size(), seam_scale, seam_scale, INTER_LINEAR) (360_stitcher/calibration.cpp:95). */
void og_cuda_resize_linear_u8(const uint8_t *src, int sw, int sh, int cn, uint8_t *dst, int dw, int dh, double fx, double fy)
{
    if (!(fx > 0) || !(fy > 0)) { fx = (double)dw / sw; fy = (double)dh / sh; }
    if (dw == sw && dh == sh) { memcpy(dst, src, (size_t)sw * sh * cn); return; }
    const float kx = (float)(1.0 / fx), ky = (float)(1.0 / fy);
    for (int dy = 0; dy < dh; ++dy)
        for (int dx = 0; dx < dw; ++dx) {
            float src_x = (float)dx * kx, src_y = (float)dy * ky;
            int x1 = f2i_rd(src_x), y1 = f2i_rd(src_y);
            int x2 = x1 + 1, y2 = y1 + 1;
            int x2r = x2 < sw - 1 ? x2 : sw - 1, y2r = y2 < sh - 1 ? y2 : sh - 1;
            for (int c = 0; c < cn; ++c) {
                float out = fmaf((float)src[((size_t)y1 * sw + x1) * cn + c], ((float)x2 - src_x) * ((float)y2 - src_y), 0.f);
                out = fmaf((float)src[((size_t)y1 * sw + x2r) * cn + c], (src_x - (float)x1) * ((float)y2 - src_y), out);
                out = fmaf((float)src[((size_t)y2r * sw + x1) * cn + c], ((float)x2 - src_x) * (src_y - (float)y1), out);
                out = fmaf((float)src[((size_t)y2r * sw + x2r) * cn + c], (src_x - (float)x1) * (src_y - (float)y1), out);
                dst[((size_t)dy * dw + dx) * cn + c] = rni_sat_u8(out);
            }
        }
}

/* GainCompensator::feed (S/src/exposure_compensate.cpp:71-142): pairwise overlap counts N and mean intensities I (double sums of
 * sqrt(b^2 + g^2 + r^2) in row-major order over the overlap, masks == 255), then A g = b with alpha = 0.01, beta = 100, solved by
 * cv::solve(DECOMP_LU) = hal::LU64f (CORE/src/matrix_decomp.cpp:52-107) for n >= 4.  Returns 0, or -1 for a singular system / n < 4. */
int og_gain_compensator_feed(int n, const uint8_t *const *imgs, const uint8_t *const *masks, const int *sizes, const int *corners, double *gains)
{
    if (n < 4 || n > 64) return -1;
    int *N = (int *)calloc((size_t)n * n, sizeof(int));
    double *I = (double *)calloc((size_t)n * n, sizeof(double)), *A = (double *)calloc((size_t)n * n, sizeof(double));
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) {
            const int wi = sizes[2 * i], hi = sizes[2 * i + 1], wj = sizes[2 * j], hj = sizes[2 * j + 1];
            const int xi = corners[2 * i], yi = corners[2 * i + 1], xj = corners[2 * j], yj = corners[2 * j + 1];
            const int x_tl = xi > xj ? xi : xj, y_tl = yi > yj ? yi : yj;
            const int x_br = xi + wi < xj + wj ? xi + wi : xj + wj, y_br = yi + hi < yj + hj ? yi + hi : yj + hj;
            if (!(x_tl < x_br && y_tl < y_br)) continue;
            int cnt = 0;
            double s1 = 0, s2 = 0;
            for (int y = y_tl; y < y_br; ++y)
                for (int x = x_tl; x < x_br; ++x) {
                    const size_t o1 = (size_t)(y - yi) * wi + (x - xi), o2 = (size_t)(y - yj) * wj + (x - xj);
                    if (masks[i][o1] != 255 || masks[j][o2] != 255) continue;
                    ++cnt;
                    const uint8_t *p = imgs[i] + o1 * 3, *q = imgs[j] + o2 * 3;
                    s1 += sqrt((double)(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]));
                    s2 += sqrt((double)(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]));
                }
            const int nn = cnt > 1 ? cnt : 1;
            N[i * n + j] = N[j * n + i] = nn;
            I[i * n + j] = s1 / nn;
            I[j * n + i] = s2 / nn;
        }
    const double alpha = 0.01, beta = 100;
    for (int i = 0; i < n; ++i) gains[i] = 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            gains[i] += beta * N[i * n + j];
            A[i * n + i] += beta * N[i * n + j];
            if (j == i) continue;
            A[i * n + i] += 2 * alpha * I[i * n + j] * I[i * n + j] * N[i * n + j];
            A[i * n + j] -= 2 * alpha * I[i * n + j] * I[j * n + i] * N[i * n + j];
        }
    int ok = 0;
    const double eps = 2.220446049250313e-16 * 100;
    for (int i = 0; i < n && ok == 0; ++i) {  /* LUImpl: partial pivoting, d = -1 / pivot */
        int k = i;
        for (int j = i + 1; j < n; ++j) if (fabs(A[j * n + i]) > fabs(A[k * n + i])) k = j;
        if (fabs(A[k * n + i]) < eps) { ok = -1; break; }
        if (k != i) {
            for (int j = i; j < n; ++j) { double t = A[i * n + j]; A[i * n + j] = A[k * n + j]; A[k * n + j] = t; }
            double t = gains[i]; gains[i] = gains[k]; gains[k] = t;
        }
        const double d = -1 / A[i * n + i];
        for (int j = i + 1; j < n; ++j) {
            const double al = A[j * n + i] * d;
            for (int q = i + 1; q < n; ++q) A[j * n + q] += al * A[i * n + q];
            gains[j] += al * gains[i];
        }
    }
    if (ok == 0)
        for (int i = n - 1; i >= 0; --i) {
            double s = gains[i];
            for (int q = i + 1; q < n; ++q) s -= A[i * n + q] * gains[q];
            gains[i] = s / A[i * n + i];
        }
    free(N); free(I); free(A);
    return ok;
}

/* cuda::createMorphologyFilter(MORPH_DILATE, CV_8U, Mat(), {-1,-1}, 1): 3x3 rect max, border REFLECT_101
 * (sources/modules/cudafilters/src/filtering.cpp:543-606; A/calibration.cpp:209,232) */
void og_dilate3x3_u8c1(const uint8_t *src, int w, int h, uint8_t *dst)
{
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            uint8_t m = 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    int yy = y + dy, xx = x + dx;
                    if (yy < 0) yy = -yy; if (yy >= h) yy = 2 * (h - 1) - yy;
                    if (xx < 0) xx = -xx; if (xx >= w) xx = 2 * (w - 1) - xx;
                    if (yy < 0) yy = 0; if (xx < 0) xx = 0;
                    uint8_t s = src[(size_t)yy * w + xx];
                    if (s > m) m = s;
                }
            dst[(size_t)y * w + x] = m;
        }
}

/* ---------------------------------------------------------------- CPW mesh -> backward map */

/* kernel `resize`, A/resize.cu:9-27.  What nvcc makes of
 *   (1-uu)*(1-vv)*in00 + uu*(1-vv)*in01 + (1-uu)*vv*in10 + uu*vv*in11
 * (checked on the PTX of the reference's own file, oracle/ref_ptx.mk, same on sm_61 ... sm_100a): the four weight products are
 * rounded on their own; the SECOND term w01*in01 is a rounded multiply, the first term is fused onto it, then the third and
 * the fourth:  fma(w11,in11, fma(w10,in10, fma(w00,in00, rn(w01*in01)))).  tests/test_oracle_ptx.py executes that PTX. */
void og_custom_resize(const float *in, int cols, int rows, float *out, int tx, int ty)
{
#pragma omp parallel for num_threads(g_threads)
    for (int v = 0; v < ty; ++v)
        for (int u = 0; u < tx; ++u) {
            int left = u * (cols - 1) / tx;
            int top = v * (rows - 1) / ty;
            float uu = ((float)u * (float)(cols - 1)) / (float)tx - (float)left;
            float vv = ((float)v * (float)(rows - 1)) / (float)ty - (float)top;
            float a = (uu * (1.f - vv)) * in[(size_t)top * cols + left + 1];
            a = fmaf((1.f - uu) * (1.f - vv), in[(size_t)top * cols + left], a);
            a = fmaf((1.f - uu) * vv, in[(size_t)(top + 1) * cols + left], a);
            a = fmaf(uu * vv, in[(size_t)(top + 1) * cols + left + 1], a);
            out[(size_t)v * tx + u] = a;
        }
}

/* MeshWarper::convertMeshesToMap, A/meshwarper.cpp:823-876 (steps m1-m3 of SURVEY appendix A) */
void og_mesh_to_half_table(const float *mesh_x, const float *mesh_y, int mesh_rows, int mesh_cols,
                           int W, int H, float *warp_x, float *warp_y)
{
    const int scale = 2;
    int hw = W / scale, hh = H / scale;
    float *big_x = (float *)malloc(sizeof(float) * W * H), *big_y = (float *)malloc(sizeof(float) * W * H);
    float *sum_x = (float *)calloc((size_t)hw * hh, sizeof(float)), *sum_y = (float *)calloc((size_t)hw * hh, sizeof(float));
    float *cnt = (float *)calloc((size_t)hw * hh, sizeof(float));
    og_custom_resize(mesh_x, mesh_cols, mesh_rows, big_x, W, H);
    og_custom_resize(mesh_y, mesh_cols, mesh_rows, big_y, W, H);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float bx = big_x[(size_t)y * W + x], by = big_y[(size_t)y * W + x];
            /* (int)float is UB out of range / NaN; x86 gives INT_MIN which fails the range test */
            int xi = (bx == bx && fabsf(bx) < 2147483648.f) ? (int)bx : INT_MIN;
            int yi = (by == by && fabsf(by) < 2147483648.f) ? (int)by : INT_MIN;
            int x_ = xi / scale, y_ = yi / scale;
            if (x_ >= 0 && y_ >= 0 && x_ < hw && y_ < hh) {
                sum_x[(size_t)y_ * hw + x_] += (float)x;
                sum_y[(size_t)y_ * hw + x_] += (float)y;
                cnt[(size_t)y_ * hw + x_] += 1.f;
            }
        }
    for (size_t i = 0; i < (size_t)hw * hh; ++i) {
        warp_x[i] = sum_x[i] / cnt[i];
        warp_y[i] = sum_y[i] / cnt[i];
    }
    free(big_x); free(big_y); free(sum_x); free(sum_y); free(cnt);
}

/* A/meshwarper.cpp:877-884: upsample the half table back to W x H */
void og_mesh_to_map(const float *mesh_x, const float *mesh_y, int mesh_rows, int mesh_cols,
                    int W, int H, float *map_x, float *map_y)
{
    int hw = W / 2, hh = H / 2;
    float *wx = (float *)malloc(sizeof(float) * hw * hh), *wy = (float *)malloc(sizeof(float) * hw * hh);
    og_mesh_to_half_table(mesh_x, mesh_y, mesh_rows, mesh_cols, W, H, wx, wy);
    og_custom_resize(wx, hw, hh, map_x, W, H);
    og_custom_resize(wy, hw, hh, map_y, W, H);
    free(wx); free(wy);
}

/* ---------------------------------------------------------------- pyramid primitives */

/* BrdReflect, sources/modules/cudev/include/opencv2/cudev/ptr2d/extrapolation.hpp:105-113,171-183 */
static inline int reflect_idx(int i, int len)
{
    int last = len - 1;
    int j = last - abs(last - i) + (i > last);       /* idx_high */
    return (abs(j) - (j < 0)) % len;                 /* idx_low  */
}

/* Reflect101, CORE/include/opencv2/core/cuda/border_interpolate.hpp:351-380 */
static inline int r101_low(int i, int len) { return abs(i) % len; }
static inline int r101_high(int i, int len) { int last = len - 1; return abs(last - abs(last - i)) % len; }
static inline int r101(int i, int len) { return r101_low(r101_high(i, len), len); }

/* cuda::copyMakeBorder(BORDER_REFLECT) CA/src/cuda/copy_make_border.cu:105-113 then convertTo(CV_16S) (S/src/blenders.cpp:711-713) */
void og_border_reflect_u8c3_to_s16(const uint8_t *src, int w, int h, size_t sstep,
                                   int top, int bottom, int left, int right, int16_t *dst)
{
    int dw = w + left + right, dh = h + top + bottom;
#pragma omp parallel for num_threads(g_threads)
    for (int y = 0; y < dh; ++y) {
        int sy = reflect_idx(y - top, h);
        for (int x = 0; x < dw; ++x) {
            int sx = reflect_idx(x - left, w);
            for (int c = 0; c < 3; ++c)
                dst[((size_t)y * dw + x) * 3 + c] = (int16_t)src[(size_t)sy * sstep + (size_t)sx * 3 + c];
        }
    }
}

void og_border_constant_f32(const float *src, int w, int h, int top, int bottom, int left, int right, float *dst)
{
    int dw = w + left + right, dh = h + top + bottom;
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            int sy = y - top, sx = x - left;
            dst[(size_t)y * dw + x] = (sy >= 0 && sy < h && sx >= 0 && sx < w) ? src[(size_t)sy * w + sx] : 0.f;
        }
}

/* cuda::pyrDown<short3, BrdReflect101>: CW/src/cuda/pyr_down.cu:55-174 (vertical 5 taps into smem, then horizontal),
 * dst size CW/src/pyramids.cpp:88 */
void og_pyr_down_s16(const int16_t *src, int w, int h, int cn, int16_t *dst)
{
    int dw = (w + 1) / 2, dh = (h + 1) / 2;
#pragma omp parallel for num_threads(g_threads)
    for (int y = 0; y < dh; ++y) {
        int sy = 2 * y;
        const int16_t *r0 = src + (size_t)r101_low(sy - 2, h) * w * cn;
        const int16_t *r1 = src + (size_t)r101_low(sy - 1, h) * w * cn;
        const int16_t *r2 = src + (size_t)sy * w * cn;
        const int16_t *r3 = src + (size_t)r101_high(sy + 1, h) * w * cn;
        const int16_t *r4 = src + (size_t)r101_high(sy + 2, h) * w * cn;
        for (int x = 0; x < dw; ++x) {
            for (int c = 0; c < cn; ++c) {
                float col[5];
                for (int t = 0; t < 5; ++t) {
                    int sx = r101(2 * x + t - 2, w) * cn + c;
                    float sum = 0.0625f * r0[sx];
                    sum = sum + 0.25f * r1[sx];
                    sum = sum + 0.375f * r2[sx];
                    sum = sum + 0.25f * r3[sx];
                    sum = sum + 0.0625f * r4[sx];
                    col[t] = sum;
                }
                float sum = 0.0625f * col[0];
                sum = sum + 0.25f * col[1];
                sum = sum + 0.375f * col[2];
                sum = sum + 0.25f * col[3];
                sum = sum + 0.0625f * col[4];
                dst[((size_t)y * dw + x) * cn + c] = rni_sat_s16(sum);
            }
        }
    }
}

/* exact integer twin: round-half-even of (sum of binomial weights)/256 */
static inline int rhe_shift(int v, int sh)
{
    int q = v >> sh, r = v & ((1 << sh) - 1), half = 1 << (sh - 1);
    if (r > half || (r == half && (q & 1))) q += 1;
    return q;
}

void og_pyr_down_s16_int(const int16_t *src, int w, int h, int cn, int16_t *dst)
{
    static const int K5[5] = {1, 4, 6, 4, 1};
    int dw = (w + 1) / 2, dh = (h + 1) / 2;
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x)
            for (int c = 0; c < cn; ++c) {
                int s = 0;
                for (int j = 0; j < 5; ++j) {
                    int sy = r101(2 * y + j - 2, h);
                    for (int i = 0; i < 5; ++i) {
                        int sx = r101(2 * x + i - 2, w);
                        s += K5[j] * K5[i] * src[((size_t)sy * w + sx) * cn + c];
                    }
                }
                dst[((size_t)y * dw + x) * cn + c] = sat_s16_i(rhe_shift(s, 8));
            }
}

/* cuda::pyrUp<short3>: CW/src/cuda/pyr_up.cu:55-145; dst = 2x src (CW/src/pyramids.cpp:126) */
static inline int up_idx(int i, int n) { i = abs(i); return i < n - 1 ? i : n - 1; }

void og_pyr_up_s16(const int16_t *src, int w, int h, int cn, int16_t *dst)
{
    int dw = 2 * w, dh = 2 * h;
#pragma omp parallel for num_threads(g_threads)
    for (int y = 0; y < dh; ++y) {
        for (int x = 0; x < dw; ++x) {
            for (int c = 0; c < cn; ++c) {
                /* horizontal pass on the (up to 3) source rows that feed dst row y */
                float hrow[3];
                int iy = y >> 1;
                int rows[3] = {up_idx(iy - 1, h), up_idx(iy, h), up_idx(iy + 1, h)};
                int ix = x >> 1;
                for (int r = 0; r < 3; ++r) {
                    const int16_t *s = src + (size_t)rows[r] * w * cn + c;
                    float sum = 0.f;
                    if ((x & 1) == 0) {
                        sum = sum + 0.0625f * s[(size_t)up_idx(ix - 1, w) * cn];
                        sum = sum + 0.375f * s[(size_t)up_idx(ix, w) * cn];
                        sum = sum + 0.0625f * s[(size_t)up_idx(ix + 1, w) * cn];
                    } else {
                        sum = sum + 0.25f * s[(size_t)up_idx(ix, w) * cn];
                        sum = sum + 0.25f * s[(size_t)up_idx(ix + 1, w) * cn];
                    }
                    hrow[r] = sum;
                }
                float sum = 0.f;
                if ((y & 1) == 0) {
                    sum = sum + 0.0625f * hrow[0];
                    sum = sum + 0.375f * hrow[1];
                    sum = sum + 0.0625f * hrow[2];
                } else {
                    sum = sum + 0.25f * hrow[1];
                    sum = sum + 0.25f * hrow[2];
                }
                dst[((size_t)y * dw + x) * cn + c] = rni_sat_s16(4.0f * sum);
            }
        }
    }
}

void og_pyr_up_s16_int(const int16_t *src, int w, int h, int cn, int16_t *dst)
{
    int dw = 2 * w, dh = 2 * h;
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x)
            for (int c = 0; c < cn; ++c) {
                int iy = y >> 1, ix = x >> 1;
                int wy[3], wx[3];
                if ((y & 1) == 0) { wy[0] = 1; wy[1] = 6; wy[2] = 1; } else { wy[0] = 0; wy[1] = 4; wy[2] = 4; }
                if ((x & 1) == 0) { wx[0] = 1; wx[1] = 6; wx[2] = 1; } else { wx[0] = 0; wx[1] = 4; wx[2] = 4; }
                int s = 0;
                for (int j = 0; j < 3; ++j)
                    for (int i = 0; i < 3; ++i)
                        s += wy[j] * wx[i] * src[((size_t)up_idx(iy + j - 1, h) * w + up_idx(ix + i - 1, w)) * cn + c];
                dst[((size_t)y * dw + x) * cn + c] = sat_s16_i(rhe_shift(s, 6));
            }
}

/* CPU-rounding twins: the vendored CPU cv::pyrDown / cv::pyrUp on CV_16S are exact integer sums with round-half-UP
 * (FixPtCast<short,8>: (v + 128) >> 8, FixPtCast<short,6>: (v + 32) >> 6; IMG/src/pyramids.cpp:52-57,1375,1483).  They
 * differ from the CUDA fp32 + cvt.rni forms above only on exact .5 ties.  Used to pin the blender restatement
 * bit-exactly against oracle/_ref (tests/test_oracle_pin.py). */
void og_pyr_down_s16_halfup(const int16_t *src, int w, int h, int cn, int16_t *dst)
{
    static const int K5[5] = {1, 4, 6, 4, 1};
    int dw = (w + 1) / 2, dh = (h + 1) / 2;
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x)
            for (int c = 0; c < cn; ++c) {
                int s = 0;
                for (int j = 0; j < 5; ++j) {
                    int sy = r101(2 * y + j - 2, h);
                    for (int i = 0; i < 5; ++i) {
                        int sx = r101(2 * x + i - 2, w);
                        s += K5[j] * K5[i] * src[((size_t)sy * w + sx) * cn + c];
                    }
                }
                dst[((size_t)y * dw + x) * cn + c] = sat_s16_i((s + 128) >> 8);
            }
}

void og_pyr_up_s16_halfup(const int16_t *src, int w, int h, int cn, int16_t *dst)
{
    int dw = 2 * w, dh = 2 * h;
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x)
            for (int c = 0; c < cn; ++c) {
                int iy = y >> 1, ix = x >> 1;
                int wy[3], wx[3];
                if ((y & 1) == 0) { wy[0] = 1; wy[1] = 6; wy[2] = 1; } else { wy[0] = 0; wy[1] = 4; wy[2] = 4; }
                if ((x & 1) == 0) { wx[0] = 1; wx[1] = 6; wx[2] = 1; } else { wx[0] = 0; wx[1] = 4; wx[2] = 4; }
                int s = 0;
                for (int j = 0; j < 3; ++j)
                    for (int i = 0; i < 3; ++i)
                        s += wy[j] * wx[i] * src[((size_t)up_idx(iy + j - 1, h) * w + up_idx(ix + i - 1, w)) * cn + c];
                dst[((size_t)y * dw + x) * cn + c] = sat_s16_i((s + 32) >> 6);
            }
}

/* cuda::pyrDown<float, BrdReflect101> for the weight pyramids (S/src/blenders.cpp:422-423): not exact in fp32,
 * so nvcc's contraction pattern matters: sum = 0.0625f*a; sum = fma(0.25f,b,sum); ... */
void og_pyr_down_f32(const float *src, int w, int h, float *dst)
{
    int dw = (w + 1) / 2, dh = (h + 1) / 2;
#pragma omp parallel for num_threads(g_threads)
    for (int y = 0; y < dh; ++y) {
        int sy = 2 * y;
        const float *r0 = src + (size_t)r101_low(sy - 2, h) * w;
        const float *r1 = src + (size_t)r101_low(sy - 1, h) * w;
        const float *r2 = src + (size_t)sy * w;
        const float *r3 = src + (size_t)r101_high(sy + 1, h) * w;
        const float *r4 = src + (size_t)r101_high(sy + 2, h) * w;
        for (int x = 0; x < dw; ++x) {
            float col[5];
            for (int t = 0; t < 5; ++t) {
                int sx = r101(2 * x + t - 2, w);
                float sum = 0.0625f * r0[sx];
                sum = fmaf(0.25f, r1[sx], sum);
                sum = fmaf(0.375f, r2[sx], sum);
                sum = fmaf(0.25f, r3[sx], sum);
                sum = fmaf(0.0625f, r4[sx], sum);
                col[t] = sum;
            }
            float sum = 0.0625f * col[0];
            sum = fmaf(0.25f, col[1], sum);
            sum = fmaf(0.375f, col[2], sum);
            sum = fmaf(0.25f, col[3], sum);
            sum = fmaf(0.0625f, col[4], sum);
            dst[(size_t)y * dw + x] = sum;
        }
    }
}

/* ---------------------------------------------------------------- Voronoi seam finder */

/* distanceTransform(src, dst, DIST_L1, 3): IMG/src/distransform.cpp:68-140 (distanceTransform_3x3, metrics {1,2}) */
static void dist_l1_3x3(const uint8_t *src, int w, int h, float *dist)
{
    const int INIT = INT_MAX >> 2, HV = 1 << 16, DG = 2 << 16;
    int step = w + 2;
    int *temp = (int *)malloc(sizeof(int) * (size_t)step * (h + 2));
    for (int j = 0; j < step; ++j) { temp[j] = INIT; temp[(size_t)(h + 1) * step + j] = INIT; }
    for (int i = 0; i < h; ++i) {
        const uint8_t *s = src + (size_t)i * w;
        int *tmp = temp + (size_t)(i + 1) * step + 1;
        tmp[-1] = tmp[w] = INIT;
        for (int j = 0; j < w; ++j) {
            if (!s[j]) tmp[j] = 0;
            else {
                int t0 = tmp[j - step - 1] + DG, t = tmp[j - step] + HV;
                if (t0 > t) t0 = t;
                t = tmp[j - step + 1] + DG; if (t0 > t) t0 = t;
                t = tmp[j - 1] + HV; if (t0 > t) t0 = t;
                tmp[j] = t0;
            }
        }
    }
    for (int i = h - 1; i >= 0; --i) {
        float *d = dist + (size_t)i * w;
        int *tmp = temp + (size_t)(i + 1) * step + 1;
        for (int j = w - 1; j >= 0; --j) {
            int t0 = tmp[j];
            if (t0 > HV) {
                int t = tmp[j + step + 1] + DG; if (t0 > t) t0 = t;
                t = tmp[j + step] + HV; if (t0 > t) t0 = t;
                t = tmp[j + step - 1] + DG; if (t0 > t) t0 = t;
                t = tmp[j + 1] + HV; if (t0 > t) t0 = t;
                tmp[j] = t0;
            }
            d[j] = (float)((float)t0 * (1.f / 65536.f));
        }
    }
    free(temp);
}

/* VoronoiSeamFinder::find(sizes, corners, masks) -> PairwiseSeamFinder::run -> findInPair,
 * S/src/seam_finders.cpp:72-162; overlapRoi S/src/util.cpp:101-113 */
void og_voronoi_find(int n, const int *sizes_wh, const int *corners_xy, uint8_t **masks)
{
    const int gap = 10;
    for (int a = 0; a < n - 1; ++a)
        for (int b = a + 1; b < n; ++b) {
            int w1 = sizes_wh[2 * a], h1 = sizes_wh[2 * a + 1], w2 = sizes_wh[2 * b], h2 = sizes_wh[2 * b + 1];
            int tl1x = corners_xy[2 * a], tl1y = corners_xy[2 * a + 1], tl2x = corners_xy[2 * b], tl2y = corners_xy[2 * b + 1];
            int x_tl = tl1x > tl2x ? tl1x : tl2x, y_tl = tl1y > tl2y ? tl1y : tl2y;
            int x_br = (tl1x + w1 < tl2x + w2) ? tl1x + w1 : tl2x + w2;
            int y_br = (tl1y + h1 < tl2y + h2) ? tl1y + h1 : tl2y + h2;
            if (!(x_tl < x_br && y_tl < y_br)) continue;
            int rw = x_br - x_tl, rh = y_br - y_tl;
            int sw = rw + 2 * gap, sh = rh + 2 * gap;
            uint8_t *sub1 = (uint8_t *)malloc((size_t)sw * sh), *sub2 = (uint8_t *)malloc((size_t)sw * sh);
            uint8_t *z1 = (uint8_t *)malloc((size_t)sw * sh), *z2 = (uint8_t *)malloc((size_t)sw * sh);
            float *d1 = (float *)malloc(sizeof(float) * sw * sh), *d2 = (float *)malloc(sizeof(float) * sw * sh);
            uint8_t *m1 = masks[a], *m2 = masks[b];
            for (int y = -gap; y < rh + gap; ++y)
                for (int x = -gap; x < rw + gap; ++x) {
                    int y1 = y_tl - tl1y + y, x1 = x_tl - tl1x + x;
                    int y2 = y_tl - tl2y + y, x2 = x_tl - tl2x + x;
                    size_t o = (size_t)(y + gap) * sw + (x + gap);
                    sub1[o] = (y1 >= 0 && x1 >= 0 && y1 < h1 && x1 < w1) ? m1[(size_t)y1 * w1 + x1] : 0;
                    sub2[o] = (y2 >= 0 && x2 >= 0 && y2 < h2 && x2 < w2) ? m2[(size_t)y2 * w2 + x2] : 0;
                }
            for (size_t o = 0; o < (size_t)sw * sh; ++o) {
                int coll = sub1[o] != 0 && sub2[o] != 0;
                uint8_t u1 = coll ? 0 : sub1[o], u2 = coll ? 0 : sub2[o];
                z1[o] = (u1 == 0) ? 255 : 0;
                z2[o] = (u2 == 0) ? 255 : 0;
            }
            dist_l1_3x3(z1, sw, sh, d1);
            dist_l1_3x3(z2, sw, sh, d2);
            for (int y = 0; y < rh; ++y)
                for (int x = 0; x < rw; ++x) {
                    size_t o = (size_t)(y + gap) * sw + (x + gap);
                    if (d1[o] < d2[o]) m2[(size_t)(y_tl - tl2y + y) * w2 + (x_tl - tl2x + x)] = 0;
                    else m1[(size_t)(y_tl - tl1y + y) * w1 + (x_tl - tl1x + x)] = 0;
                }
            free(sub1); free(sub2); free(z1); free(z2); free(d1); free(d2);
        }
}

/* ---------------------------------------------------------------- MultiBandBlender (authors' GPU variant) */

#define OG_MAX_VIEWS 64
#define OG_MAX_LEVELS 16

typedef struct {
    int top, bottom, left, right;
    int x_tl, y_tl, x_br, y_br;    /* dst rect in padded-canvas coords at level 0 */
    int bw, bh;                    /* bordered size at level 0 */
    float *weight[OG_MAX_LEVELS];  /* gpu_weight_pyr_gauss_vec_[i][k] */
    int16_t *lap[OG_MAX_LEVELS];   /* gpu_src_pyr_laplace_vec[i][k] */
} og_view;

struct og_blender {
    int actual_num_bands, num_bands;
    int roi_final[4];              /* dst_roi_final_ */
    int roi[4];                    /* dst_roi_ (padded) */
    int lw[OG_MAX_LEVELS], lh[OG_MAX_LEVELS];
    int16_t *dst[OG_MAX_LEVELS];   /* gpu_dst_pyr_laplace_ */
    float *dstw[OG_MAX_LEVELS];    /* gpu_dst_band_weights_ */
    int n_views;
    int cpu_pyramids;              /* test hook: half-up integer pyramids (the CPU twins) instead of fp32 + rni */
    og_view views[OG_MAX_VIEWS];
};

og_blender *og_blender_create(int num_bands)
{
    og_blender *b = (og_blender *)calloc(1, sizeof(og_blender));
    b->actual_num_bands = num_bands;   /* setNumBands, S/include/opencv2/stitching/detail/blenders.hpp:131 */
    return b;
}

static void free_view(og_view *v)
{
    for (int k = 0; k < OG_MAX_LEVELS; ++k) { free(v->weight[k]); free(v->lap[k]); v->weight[k] = NULL; v->lap[k] = NULL; }
}

void og_blender_destroy(og_blender *b)
{
    if (!b) return;
    for (int k = 0; k < OG_MAX_LEVELS; ++k) { free(b->dst[k]); free(b->dstw[k]); }
    for (int i = 0; i < b->n_views; ++i) free_view(&b->views[i]);
    free(b);
}

/* Blender::prepare(corners,sizes) -> resultRoi (S/src/util.cpp:125-138) -> MultiBandBlender::prepare(Rect), S/src/blenders.cpp:237-274 */
int og_blender_prepare(og_blender *b, int n, const int *corners_xy, const int *sizes_wh)
{
    int tlx = INT_MAX, tly = INT_MAX, brx = INT_MIN, bry = INT_MIN;
    for (int i = 0; i < n; ++i) {
        if (corners_xy[2 * i] < tlx) tlx = corners_xy[2 * i];
        if (corners_xy[2 * i + 1] < tly) tly = corners_xy[2 * i + 1];
        if (corners_xy[2 * i] + sizes_wh[2 * i] > brx) brx = corners_xy[2 * i] + sizes_wh[2 * i];
        if (corners_xy[2 * i + 1] + sizes_wh[2 * i + 1] > bry) bry = corners_xy[2 * i + 1] + sizes_wh[2 * i + 1];
    }
    int W = brx - tlx, H = bry - tly;
    b->roi_final[0] = tlx; b->roi_final[1] = tly; b->roi_final[2] = W; b->roi_final[3] = H;
    double max_len = (double)(W > H ? W : H);
    int nb = (int)ceil(log(max_len) / log(2.0));
    b->num_bands = b->actual_num_bands < nb ? b->actual_num_bands : nb;
    int m = 1 << b->num_bands;
    W += (m - W % m) % m;
    H += (m - H % m) % m;
    b->roi[0] = tlx; b->roi[1] = tly; b->roi[2] = W; b->roi[3] = H;
    for (int i = 0; i < b->n_views; ++i) free_view(&b->views[i]);
    b->n_views = 0;
    for (int k = 0; k <= b->num_bands; ++k) {
        b->lw[k] = k == 0 ? W : (b->lw[k - 1] + 1) / 2;
        b->lh[k] = k == 0 ? H : (b->lh[k - 1] + 1) / 2;
        free(b->dst[k]); free(b->dstw[k]);
        b->dst[k] = (int16_t *)calloc((size_t)b->lw[k] * b->lh[k] * 3, sizeof(int16_t));
        b->dstw[k] = (float *)calloc((size_t)b->lw[k] * b->lh[k], sizeof(float));
    }
    return 0;
}

int og_blender_num_bands(const og_blender *b) { return b->num_bands; }

static void level_size(int w0, int h0, int k, int *w, int *h);
/* test hooks (tests/test_oracle_pin.py): CPU-rounding pyramids, and injection of a weight level computed elsewhere */
void og_blender_set_cpu_pyramids(og_blender *b, int on) { b->cpu_pyramids = on; }
void og_blender_set_view_weight(og_blender *b, int i, int level, const float *w)
{
    int lw, lh;
    level_size(b->views[i].bw, b->views[i].bh, level, &lw, &lh);
    memcpy(b->views[i].weight[level], w, sizeof(float) * lw * lh);
}

void og_blender_dst_roi(const og_blender *b, int roi_final[4], int roi_padded[4])
{
    memcpy(roi_final, b->roi_final, sizeof(int) * 4);
    memcpy(roi_padded, b->roi, sizeof(int) * 4);
}

/* MultiBandBlender::init_gpu, S/src/blenders.cpp:344-434 */
int og_blender_init_view(og_blender *b, const uint8_t *mask, int mw, int mh, size_t mstep, int tl_x, int tl_y)
{
    if (b->n_views >= OG_MAX_VIEWS) return -1;
    og_view *v = &b->views[b->n_views];
    int nb = b->num_bands, m = 1 << nb;
    int rx = b->roi[0], ry = b->roi[1], rbx = b->roi[0] + b->roi[2], rby = b->roi[1] + b->roi[3];
    int gap = 3 * m;
    int tnx = rx > tl_x - gap ? rx : tl_x - gap, tny = ry > tl_y - gap ? ry : tl_y - gap;
    int bnx = rbx < tl_x + mw + gap ? rbx : tl_x + mw + gap, bny = rby < tl_y + mh + gap ? rby : tl_y + mh + gap;
    tnx = rx + (((tnx - rx) >> nb) << nb);
    tny = ry + (((tny - ry) >> nb) << nb);
    int width = bnx - tnx, height = bny - tny;
    width += (m - width % m) % m;
    height += (m - height % m) % m;
    bnx = tnx + width; bny = tny + height;
    int dy = bny - rby > 0 ? bny - rby : 0, dx = bnx - rbx > 0 ? bnx - rbx : 0;
    tnx -= dx; bnx -= dx; tny -= dy; bny -= dy;
    v->top = tl_y - tny; v->left = tl_x - tnx;
    v->bottom = bny - tl_y - mh; v->right = bnx - tl_x - mw;
    v->y_tl = tny - ry; v->y_br = bny - ry; v->x_tl = tnx - rx; v->x_br = bnx - rx;
    v->bw = width; v->bh = height;
    /* weight_map = mask * (1/255) (convertTo CV_32F, S/src/blenders.cpp:412), then BORDER_CONSTANT, then pyrDown chain */
    float alpha = (float)(1. / 255.);
    float *wm = (float *)malloc(sizeof(float) * mw * mh);
    for (int y = 0; y < mh; ++y)
        for (int x = 0; x < mw; ++x) wm[(size_t)y * mw + x] = fmaf(alpha, (float)mask[(size_t)y * mstep + x], 0.f);
    v->weight[0] = (float *)malloc(sizeof(float) * width * height);
    og_border_constant_f32(wm, mw, mh, v->top, v->bottom, v->left, v->right, v->weight[0]);
    free(wm);
    int w = width, h = height;
    for (int k = 0; k < nb; ++k) {
        int nw = (w + 1) / 2, nh = (h + 1) / 2;
        v->weight[k + 1] = (float *)malloc(sizeof(float) * nw * nh);
        og_pyr_down_f32(v->weight[k], w, h, v->weight[k + 1]);
        w = nw; h = nh;
    }
    b->n_views++;
    return b->n_views - 1;
}

void og_blender_view_geom(const og_blender *b, int i, int out[8])
{
    const og_view *v = &b->views[i];
    out[0] = v->top; out[1] = v->bottom; out[2] = v->left; out[3] = v->right;
    out[4] = v->x_tl; out[5] = v->y_tl; out[6] = v->x_br; out[7] = v->y_br;
}

static void level_size(int w0, int h0, int k, int *w, int *h)
{
    for (int i = 0; i < k; ++i) { w0 = (w0 + 1) / 2; h0 = (h0 + 1) / 2; }
    *w = w0; *h = h0;
}

const float *og_blender_view_weight(const og_blender *b, int i, int level, int *w, int *h)
{
    level_size(b->views[i].bw, b->views[i].bh, level, w, h);
    return b->views[i].weight[level];
}

const float *og_blender_dst_weight(const og_blender *b, int level, int *w, int *h)
{
    *w = b->lw[level]; *h = b->lh[level];
    return b->dstw[level];
}

const int16_t *og_blender_dst_level(const og_blender *b, int level, int *w, int *h)
{
    *w = b->lw[level]; *h = b->lh[level];
    return b->dst[level];
}

const int16_t *og_blender_src_level(const og_blender *b, int i, int level, int *w, int *h)
{
    level_size(b->views[i].bw, b->views[i].bh, level, w, h);
    return b->views[i].lap[level];
}

/* one pixel of addSrcWeightKernel32F (S/src/cuda/multiband_blend.cu:36-50): dst.c += static_cast<short>(v.c * w) -- one rounded
 * multiply, cvt.rzi, 16-bit wrap-around add (the PTX of the reference's file says add.s16) -- and dst_weight += w */
static inline void add_src_weight_px(const int16_t *s, float wgt, int16_t *d, float *dw)
{
    for (int c = 0; c < 3; ++c) d[c] = (int16_t)(d[c] + rz_s16((float)s[c] * wgt));
    *dw += wgt;
}
/* one pixel of normalizeUsingWeightKernel32F (S/src/cuda/multiband_blend.cu:85-99): v.c / (w + WEIGHT_EPS), IEEE division, cvt.rzi */
static inline void normalize_px(int16_t *d, float w)
{
    const float wv = w + 1e-5f;
    for (int c = 0; c < 3; ++c) d[c] = rz_s16((float)d[c] / wv);
}
/* the two kernels over a rows x cols rectangle of densely packed arrays (what addSrcWeightGpu32F / normalizeUsingWeightMapGpu32F launch) */
void og_add_src_weight_32f(const int16_t *src, const float *weight, int16_t *dst, float *dst_weight, int rows, int cols)
{
    for (size_t j = 0; j < (size_t)rows * cols; ++j) add_src_weight_px(src + 3 * j, weight[j], dst + 3 * j, dst_weight + j);
}
void og_normalize_32f(const float *weight, int16_t *src, int rows, int cols)
{
    for (size_t j = 0; j < (size_t)rows * cols; ++j) normalize_px(src + 3 * j, weight[j]);
}

/* MultiBandBlender::feed_online, S/src/blenders.cpp:700-749; addSrcWeightKernel32F S/src/cuda/multiband_blend.cu:36-50;
 * subtract = saturating s16 (CA/src/cuda/sub_mat.cu:59-65) */
void og_blender_feed_online(og_blender *b, int i, const uint8_t *img, int w, int h, size_t step)
{
    og_view *v = &b->views[i];
    int nb = b->num_bands;
    int lw[OG_MAX_LEVELS], lh[OG_MAX_LEVELS];
    for (int k = 0; k <= nb; ++k) {
        level_size(v->bw, v->bh, k, &lw[k], &lh[k]);
        if (!v->lap[k]) v->lap[k] = (int16_t *)malloc(sizeof(int16_t) * 3 * lw[k] * lh[k]);
    }
    og_border_reflect_u8c3_to_s16(img, w, h, step, v->top, v->bottom, v->left, v->right, v->lap[0]);
    for (int k = 0; k < nb; ++k) (b->cpu_pyramids ? og_pyr_down_s16_halfup : og_pyr_down_s16)(v->lap[k], lw[k], lh[k], 3, v->lap[k + 1]);
    for (int k = 0; k < nb; ++k) {
        int16_t *up = (int16_t *)malloc(sizeof(int16_t) * 3 * lw[k] * lh[k]);
        (b->cpu_pyramids ? og_pyr_up_s16_halfup : og_pyr_up_s16)(v->lap[k + 1], lw[k + 1], lh[k + 1], 3, up);
        size_t n = (size_t)3 * lw[k] * lh[k];
        for (size_t j = 0; j < n; ++j) v->lap[k][j] = sat_s16_i((int)v->lap[k][j] - (int)up[j]);
        free(up);
    }
    int x_tl = v->x_tl, y_tl = v->y_tl, x_br = v->x_br, y_br = v->y_br;
    for (int k = 0; k <= nb; ++k) {
        int rw = x_br - x_tl, rh = y_br - y_tl;
#pragma omp parallel for num_threads(g_threads)
        for (int y = 0; y < rh; ++y)
            for (int x = 0; x < rw; ++x) {
                float wgt = v->weight[k][(size_t)y * lw[k] + x];
                const int16_t *s = v->lap[k] + ((size_t)y * lw[k] + x) * 3;
                int16_t *d = b->dst[k] + ((size_t)(y_tl + y) * b->lw[k] + (x_tl + x)) * 3;
                add_src_weight_px(s, wgt, d, &b->dstw[k][(size_t)(y_tl + y) * b->lw[k] + (x_tl + x)]);
            }
        x_tl /= 2; y_tl /= 2; x_br /= 2; y_br /= 2;
    }
}

/* MultiBandBlender::blend(dst, dst_mask, gpuOut, true), S/src/blenders.cpp:758-832;
 * normalizeUsingWeightKernel32F S/src/cuda/multiband_blend.cu:85-99; add = saturating s16 (CA/src/cuda/add_mat.cu:59-65) */
void og_blender_blend(og_blender *b, int16_t *out, uint8_t *mask_out)
{
    const float WEIGHT_EPS = 1e-5f;
    int nb = b->num_bands;
    for (int k = 0; k <= nb; ++k) {
        size_t n = (size_t)b->lw[k] * b->lh[k];
#pragma omp parallel for num_threads(g_threads)
        for (size_t j = 0; j < n; ++j) normalize_px(b->dst[k] + j * 3, b->dstw[k][j]);
    }
    for (int k = nb; k > 0; --k) {
        size_t n = (size_t)3 * b->lw[k - 1] * b->lh[k - 1];
        int16_t *up = (int16_t *)malloc(sizeof(int16_t) * n);
        (b->cpu_pyramids ? og_pyr_up_s16_halfup : og_pyr_up_s16)(b->dst[k], b->lw[k], b->lh[k], 3, up);
        for (size_t j = 0; j < n; ++j) b->dst[k - 1][j] = sat_s16_i((int)up[j] + (int)b->dst[k - 1][j]);
        free(up);
    }
    int W = b->roi_final[2], H = b->roi_final[3];
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int m = b->dstw[0][(size_t)y * b->lw[0] + x] > WEIGHT_EPS;
            if (mask_out) mask_out[(size_t)y * W + x] = m ? 255 : 0;
            for (int c = 0; c < 3; ++c)
                out[((size_t)y * W + x) * 3 + c] = m ? b->dst[0][((size_t)y * b->lw[0] + x) * 3 + c] : 0;
        }
    for (int k = 0; k <= nb; ++k) {
        memset(b->dst[k], 0, sizeof(int16_t) * 3 * b->lw[k] * b->lh[k]);
        memset(b->dstw[k], 0, sizeof(float) * b->lw[k] * b->lh[k]);
    }
}
