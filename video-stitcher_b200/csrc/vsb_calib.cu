// vsb_calib.cu -- host-side calibration that produces every STATIC input of the compose path, mirroring the
// reference's fixed-rig calibration (360_stitcher/calibration.cpp:28-249) generalised to N views:
//   calibrateCameras -> seam-scale mask warp -> VoronoiSeamFinder -> compose-scale warpRoi/buildMaps ->
//   prepare -> per view: mask warp, dilate, linear resize, AND, init_gpu.
// It runs once, on the host (as most of it does in the reference), with the host libm, so its products are
// bit-reproducible; the per-frame work is in vsb_pipeline.cu.  Gain ESTIMATION is out of scope:
// gains are an input.
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>

#include "vsb_internal.h"

namespace vsb {

static inline int f2i_rz(float v)
{
    if (!(v == v)) return 0;
    if (v <= -2147483648.f) return INT_MIN;
    if (v >= 2147483648.f) return INT_MAX;
    return (int)v;
}
static inline int f2i_rd(float v) { return f2i_rz(std::floor(v)); }
static inline uint8_t rni_sat_u8_host(float v)
{
    if (!(v == v) || v <= 0.f) return 0;
    if (v >= 255.f) return 255;
    return (uint8_t)std::nearbyint(v);
}

// {Spherical,Cylindrical}Mapper::mapBackward with the device's contraction pattern, on the host
// (sources/modules/stitching/src/cuda/build_warp_maps.cu:88-134; the pattern is what nvcc emits for the reference's file, see
// k_build_maps in vsb_primitives.cu: spherical = two rounded products subtracted, third term fused; cylindrical = second product
// rounded, first and third fused).  Host libm sinf / cosf.
static void host_build_maps(int proj, float scale, const float K[9], const float R[9], int tl_x, int tl_y, int w, int h,
                            float *xmap, float *ymap)
{
    float k[9], r_kinv[9], rinv[9];
    projector_setup(K, R, k, r_kinv, rinv);
    for (int dv = 0; dv < h; ++dv)
        for (int du = 0; du < w; ++du) {
            float u = (float)(tl_x + du), v = (float)(tl_y + dv);
            float x, y, z;
            if (proj == VSB_PROJ_SPHERICAL) {
                v = v / scale; u = u / scale;
                const float sinv = sinf(v);
                const float x_ = sinv * sinf(u);
                const float cosv = cosf(v);
                const float z_ = sinv * cosf(u);
                // (volatile: each product and the difference are rounded on their own whatever the host compiler's contraction setting)
                volatile float x0 = x_ * k[0], x1 = cosv * k[1], y0 = x_ * k[3], y1 = cosv * k[4], z0 = x_ * k[6], z1 = cosv * k[7];
                volatile float xs = x0 - x1, ys = y0 - y1, zs = z0 - z1;
                x = std::fmaf(k[2], z_, xs);
                y = std::fmaf(k[5], z_, ys);
                z = std::fmaf(k[8], z_, zs);
            } else {
                u = u / scale;
                const float x_ = sinf(u);
                const float y_ = v / scale;
                const float z_ = cosf(u);
                volatile float x1 = y_ * k[1], y1 = y_ * k[4], z1 = y_ * k[7];
                x = std::fmaf(k[2], z_, std::fmaf(x_, k[0], x1));
                y = std::fmaf(k[5], z_, std::fmaf(x_, k[3], y1));
                z = std::fmaf(k[8], z_, std::fmaf(x_, k[6], z1));
            }
            if (z > 0) { x = x / z; y = y / z; } else { x = y = -1.f; }
            xmap[(size_t)dv * w + du] = x;
            ymap[(size_t)dv * w + du] = y;
        }
}

// warp of an all-255 mask with INTER_NEAREST / BORDER_CONSTANT (PointFilter, __float2int_rz)
static void host_warp_full_mask(const float *xmap, const float *ymap, int w, int h, int src_w, int src_h, uint8_t *dst)
{
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        const int x = f2i_rz(xmap[i]), y = f2i_rz(ymap[i]);
        dst[i] = (x >= 0 && x < src_w && y >= 0 && y < src_h) ? 255 : 0;
    }
}

// distanceTransform(DIST_L1, 3) (sources/modules/imgproc/src/distransform.cpp:68-140)
static void dist_l1_3x3(const uint8_t *src, int w, int h, float *dist)
{
    const int INIT = INT_MAX >> 2, HV = 1 << 16, DG = 2 << 16;
    const int step = w + 2;
    std::vector<int> tbuf((size_t)step * (h + 2));
    int *temp = tbuf.data();
    for (int j = 0; j < step; ++j) { temp[j] = INIT; temp[(size_t)(h + 1) * step + j] = INIT; }
    for (int i = 0; i < h; ++i) {
        const uint8_t *s = src + (size_t)i * w;
        int *tmp = temp + (size_t)(i + 1) * step + 1;
        tmp[-1] = tmp[w] = INIT;
        for (int j = 0; j < w; ++j) {
            if (!s[j]) { tmp[j] = 0; continue; }
            int t0 = tmp[j - step - 1] + DG, t = tmp[j - step] + HV;
            if (t0 > t) t0 = t;
            t = tmp[j - step + 1] + DG; if (t0 > t) t0 = t;
            t = tmp[j - 1] + HV; if (t0 > t) t0 = t;
            tmp[j] = t0;
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
            d[j] = (float)t0 * (1.f / 65536.f);
        }
    }
}

// VoronoiSeamFinder (sources/modules/stitching/src/seam_finders.cpp:72-162)
static void voronoi(int n, const int *sizes, const int *corners, uint8_t *const *masks)
{
    const int gap = 10;
    for (int a = 0; a < n - 1; ++a)
        for (int b = a + 1; b < n; ++b) {
            const int w1 = sizes[2 * a], h1 = sizes[2 * a + 1], w2 = sizes[2 * b], h2 = sizes[2 * b + 1];
            const int t1x = corners[2 * a], t1y = corners[2 * a + 1], t2x = corners[2 * b], t2y = corners[2 * b + 1];
            const int x_tl = std::max(t1x, t2x), y_tl = std::max(t1y, t2y);
            const int x_br = std::min(t1x + w1, t2x + w2), y_br = std::min(t1y + h1, t2y + h2);
            if (!(x_tl < x_br && y_tl < y_br)) continue;  // overlapRoi, sources/modules/stitching/src/util.cpp:101-113
            const int rw = x_br - x_tl, rh = y_br - y_tl, sw = rw + 2 * gap, sh = rh + 2 * gap;
            std::vector<uint8_t> z1((size_t)sw * sh), z2((size_t)sw * sh);
            std::vector<float> d1((size_t)sw * sh), d2((size_t)sw * sh);
            uint8_t *m1 = masks[a], *m2 = masks[b];
            for (int y = -gap; y < rh + gap; ++y)
                for (int x = -gap; x < rw + gap; ++x) {
                    const int y1 = y_tl - t1y + y, x1 = x_tl - t1x + x, y2 = y_tl - t2y + y, x2 = x_tl - t2x + x;
                    const uint8_t s1 = (y1 >= 0 && x1 >= 0 && y1 < h1 && x1 < w1) ? m1[(size_t)y1 * w1 + x1] : 0;
                    const uint8_t s2 = (y2 >= 0 && x2 >= 0 && y2 < h2 && x2 < w2) ? m2[(size_t)y2 * w2 + x2] : 0;
                    const bool coll = s1 && s2;
                    const size_t o = (size_t)(y + gap) * sw + (x + gap);
                    z1[o] = ((coll ? 0 : s1) == 0) ? 255 : 0;  // unique1 == 0
                    z2[o] = ((coll ? 0 : s2) == 0) ? 255 : 0;
                }
            dist_l1_3x3(z1.data(), sw, sh, d1.data());
            dist_l1_3x3(z2.data(), sw, sh, d2.data());
            for (int y = 0; y < rh; ++y)
                for (int x = 0; x < rw; ++x) {
                    const size_t o = (size_t)(y + gap) * sw + (x + gap);
                    if (d1[o] < d2[o]) m2[(size_t)(y_tl - t2y + y) * w2 + (x_tl - t2x + x)] = 0;
                    else m1[(size_t)(y_tl - t1y + y) * w1 + (x_tl - t1x + x)] = 0;
                }
        }
}

// MORPH_DILATE 3x3 rect, BORDER_REFLECT_101 (sources/modules/cudafilters/src/filtering.cpp:543-606)
static void dilate3x3(const uint8_t *src, int w, int h, uint8_t *dst)
{
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            uint8_t m = 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    int yy = y + dy, xx = x + dx;
                    if (yy < 0) yy = -yy;
                    if (yy >= h) yy = 2 * (h - 1) - yy;
                    if (xx < 0) xx = -xx;
                    if (xx >= w) xx = 2 * (w - 1) - xx;
                    yy = std::max(yy, 0); xx = std::max(xx, 0);
                    m = std::max(m, src[(size_t)yy * w + xx]);
                }
            dst[(size_t)y * w + x] = m;
        }
}

// cuda::resize INTER_LINEAR CV_8UC1 (sources/modules/cudawarping/src/cuda/resize.cu:71-106, src/resize.cpp:76-105)
static void resize_linear_u8(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh)
{
    if (dw == sw && dh == sh) { std::memcpy(dst, src, (size_t)sw * sh); return; }
    const float fx = (float)(1.0 / ((double)dw / sw)), fy = (float)(1.0 / ((double)dh / sh));
    for (int dy = 0; dy < dh; ++dy)
        for (int dx = 0; dx < dw; ++dx) {
            const float sx = (float)dx * fx, sy = (float)dy * fy;
            const int x1 = f2i_rd(sx), y1 = f2i_rd(sy), x2 = x1 + 1, y2 = y1 + 1;
            const int x2r = std::min(x2, sw - 1), y2r = std::min(y2, sh - 1);
            float o = (float)src[(size_t)y1 * sw + x1] * (((float)x2 - sx) * ((float)y2 - sy));
            o = std::fmaf((float)src[(size_t)y1 * sw + x2r], (sx - (float)x1) * ((float)y2 - sy), o);
            o = std::fmaf((float)src[(size_t)y2r * sw + x1], ((float)x2 - sx) * (sy - (float)y1), o);
            o = std::fmaf((float)src[(size_t)y2r * sw + x2r], (sx - (float)x1) * (sy - (float)y1), o);
            dst[(size_t)dy * dw + dx] = rni_sat_u8_host(o);
        }
}


// ============================================================================================ calibration on the device
// the per-pixel work of the static half of the path as CUDA kernels, so that seams and gains can be
// refreshed at run time without a host loop over pixels.  Integer / index work (Voronoi labels, masks) is bit-exact against the
// host functions above and the reference's golden vectors; the fp32 pieces use the same operation order as the host twins.

// warp of an all-255 mask, INTER_NEAREST / BORDER_CONSTANT (A/calibration.cpp:122,227): 255 where the map addresses the image
__global__ void k_full_mask(const float *__restrict__ xm, const float *__restrict__ ym, size_t mp, int w, int h, int sw, int sh, uint8_t *__restrict__ dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int sx = __float2int_rz(*(const float *)((const char *)xm + (size_t)y * mp + (size_t)x * 4));
    const int sy = __float2int_rz(*(const float *)((const char *)ym + (size_t)y * mp + (size_t)x * 4));
    dst[(size_t)y * w + x] = ((unsigned)sx < (unsigned)sw && (unsigned)sy < (unsigned)sh) ? 255 : 0;
}

// VoronoiSeamFinder::findInPair (sources/modules/stitching/src/seam_finders.cpp:111-162) for one pair of views.
// distanceTransform(DIST_L1, 3) (sources/modules/imgproc/src/distransform.cpp:68-140) is the exact city-block distance to the
// nearest zero pixel (chamfer weights 1 / 2), which separates: a column pass (distance along the column), then for every pixel
// the minimum over the row of |dx| + column distance.  "No zero pixel anywhere" keeps one shared INF, which orders exactly like
// the reference's INIT-based values (larger than every finite distance, equal on both sides).
struct VorPair {
    uint8_t *m1, *m2;
    int w1, h1, t1x, t1y, w2, h2, t2x, t2y;
    int x_tl, y_tl, rw, rh;  // overlap rect (canvas coordinates) and size
};
constexpr int VOR_GAP = 10, VOR_INF = 1 << 28;

__device__ __forceinline__ void vor_unique(const VorPair &P, int x, int y, bool &u1, bool &u2)  // (x, y) relative to the overlap origin
{
    const int y1 = P.y_tl - P.t1y + y, x1 = P.x_tl - P.t1x + x, y2 = P.y_tl - P.t2y + y, x2 = P.x_tl - P.t2x + x;
    const bool s1 = (unsigned)y1 < (unsigned)P.h1 && (unsigned)x1 < (unsigned)P.w1 && P.m1[(size_t)y1 * P.w1 + x1] != 0;
    const bool s2 = (unsigned)y2 < (unsigned)P.h2 && (unsigned)x2 < (unsigned)P.w2 && P.m2[(size_t)y2 * P.w2 + x2] != 0;
    u1 = s1 && !s2; u2 = s2 && !s1;  // pixels that belong to one image only: the zeros the distances are measured to
}

// one thread per column of the padded region: distance along the column to the nearest unique pixel of image 1 / image 2
__global__ void k_vor_columns(const VorPair P, int *__restrict__ g1, int *__restrict__ g2)
{
    const int sw = P.rw + 2 * VOR_GAP, sh = P.rh + 2 * VOR_GAP;
    const int xc = blockIdx.x * blockDim.x + threadIdx.x;
    if (xc >= sw) return;
    int d1 = VOR_INF, d2 = VOR_INF;
    for (int yc = 0; yc < sh; ++yc) {
        bool u1, u2;
        vor_unique(P, xc - VOR_GAP, yc - VOR_GAP, u1, u2);
        d1 = u1 ? 0 : min(d1 + 1, VOR_INF); d2 = u2 ? 0 : min(d2 + 1, VOR_INF);
        g1[(size_t)yc * sw + xc] = d1; g2[(size_t)yc * sw + xc] = d2;
    }
    d1 = d2 = VOR_INF;
    for (int yc = sh - 1; yc >= 0; --yc) {
        const size_t o = (size_t)yc * sw + xc;
        d1 = g1[o] == 0 ? 0 : min(d1 + 1, VOR_INF); d2 = g2[o] == 0 ? 0 : min(d2 + 1, VOR_INF);
        g1[o] = min(g1[o], d1); g2[o] = min(g2[o], d2);
    }
}
// one thread per pixel of the overlap: row minimum, then the pixel leaves the mask of the farther image (seam_finders.cpp:153-161)
__global__ void k_vor_decide(const VorPair P, const int *__restrict__ g1, const int *__restrict__ g2)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.rw || y >= P.rh) return;
    const int sw = P.rw + 2 * VOR_GAP, xc = x + VOR_GAP;
    const int *r1 = g1 + (size_t)(y + VOR_GAP) * sw, *r2 = g2 + (size_t)(y + VOR_GAP) * sw;
    int d1 = VOR_INF, d2 = VOR_INF;
    for (int q = 0; q < sw; ++q) {
        const int dx = abs(q - xc);
        d1 = min(d1, min(r1[q] + dx, VOR_INF)); d2 = min(d2, min(r2[q] + dx, VOR_INF));
    }
    if (d1 < d2) P.m2[(size_t)(P.y_tl - P.t2y + y) * P.w2 + (P.x_tl - P.t2x + x)] = 0;
    else P.m1[(size_t)(P.y_tl - P.t1y + y) * P.w1 + (P.x_tl - P.t1x + x)] = 0;
}

// MORPH_DILATE 3x3 rect, BORDER_REFLECT_101 (sources/modules/cudafilters/src/filtering.cpp:543-606; A/calibration.cpp:209,232)
__global__ void k_dilate3x3(const uint8_t *__restrict__ src, int w, int h, uint8_t *__restrict__ dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    unsigned m = 0;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            int yy = y + dy, xx = x + dx;
            if (yy < 0) yy = -yy;
            if (yy >= h) yy = 2 * (h - 1) - yy;
            if (xx < 0) xx = -xx;
            if (xx >= w) xx = 2 * (w - 1) - xx;
            yy = max(yy, 0); xx = max(xx, 0);
            m = max(m, (unsigned)src[(size_t)yy * w + xx]);
        }
    dst[(size_t)y * w + x] = (uint8_t)m;
}

// cuda::resize INTER_LINEAR on CV_8UC1 / CV_8UC3 (sources/modules/cudawarping/src/cuda/resize.cu:71-106); the per-sample
// arithmetic is resize_linear_px (vsb_device.cuh).  fx / fy are the factors the host wrapper passes (src/resize.cpp:76-105).
template <int CN>
__global__ void k_resize_linear_u8(const uint8_t *__restrict__ src, int sw, int sh, size_t sp, uint8_t *__restrict__ dst, int dw, int dh, size_t dp, float fx, float fy)
{
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y * blockDim.y + threadIdx.y;
    if (dx >= dw || dy >= dh) return;
    resize_linear_px<CN>(src, sw, sh, sp, dst, dp, dx, dy, fx, fy);
}

__global__ void k_and_u8(uint8_t *__restrict__ a, const uint8_t *__restrict__ b, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] &= b[i];
}

// GainCompensator::feed, the reductions (sources/modules/stitching/src/exposure_compensate.cpp:89-121): one thread per pair
// (i <= j) walks the overlap in the reference's order (rows, then columns), so the double sums are the reference's bit for bit.
struct GainView { const uint8_t *img, *mask; int w, h, tx, ty; };  // warped seam-scale image (CV_8UC3, tight rows), its warped mask, corner
struct GainParams { int n; GainView v[VSB_MAX_VIEWS]; };
__global__ void k_gain_pairs(const GainParams P, int *__restrict__ N, double *__restrict__ I)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.n * P.n) return;
    const int i = t / P.n, j = t - i * P.n;
    if (j < i) return;
    const GainView &A = P.v[i], &B = P.v[j];
    const int x_tl = max(A.tx, B.tx), y_tl = max(A.ty, B.ty), x_br = min(A.tx + A.w, B.tx + B.w), y_br = min(A.ty + A.h, B.ty + B.h);
    if (!(x_tl < x_br && y_tl < y_br)) return;  // overlapRoi (util.cpp:101-113): N and I stay 0
    int cnt = 0;
    double s1 = 0, s2 = 0;
    for (int y = y_tl; y < y_br; ++y)
        for (int x = x_tl; x < x_br; ++x) {
            const size_t o1 = (size_t)(y - A.ty) * A.w + (x - A.tx), o2 = (size_t)(y - B.ty) * B.w + (x - B.tx);
            if (A.mask[o1] != 255 || B.mask[o2] != 255) continue;
            ++cnt;
            const uint8_t *p = A.img + o1 * 3, *q = B.img + o2 * 3;
            s1 += sqrt((double)((int)p[0] * p[0] + (int)p[1] * p[1] + (int)p[2] * p[2]));
            s2 += sqrt((double)((int)q[0] * q[0] + (int)q[1] * q[1] + (int)q[2] * q[2]));
        }
    const int nn = max(1, cnt);
    N[i * P.n + j] = N[j * P.n + i] = nn;
    I[i * P.n + j] = s1 / nn;
    I[j * P.n + i] = s2 / nn;
}

// seam-scale state kept with the handle (vsb_calibrate_rig_device): what a gain refresh needs
struct CalibState {
    int n = 0, projection = 0, src_w = 0, src_h = 0, seam_w = 0, seam_h = 0;
    double seam_scale = 1.0;
    float seam_warp_scale = 0.f;
    float Ks[VSB_MAX_VIEWS][9], R[VSB_MAX_VIEWS][9];
    int roi[VSB_MAX_VIEWS][4];
    uint8_t *warped_mask[VSB_MAX_VIEWS] = {};  // warped all-255 masks at seam scale, BEFORE the seam finder (what the compensator gets)
};
static void calib_state_free(void *p)
{
    CalibState *c = static_cast<CalibState *>(p);
    for (int i = 0; i < VSB_MAX_VIEWS; ++i) cudaFree(c->warped_mask[i]);
    delete c;
}

// cv::solve(A, b, x, DECOMP_LU) for n >= 4 (sources/modules/core/src/lapack.cpp -> hal::LU64f, matrix_decomp.cpp:52-107)
static bool lu_solve(std::vector<double> &A, std::vector<double> &b, int m)
{
    const double eps = 2.220446049250313e-16 * 100;
    for (int i = 0; i < m; ++i) {
        int k = i;
        for (int j = i + 1; j < m; ++j) if (std::abs(A[j * m + i]) > std::abs(A[k * m + i])) k = j;
        if (std::abs(A[k * m + i]) < eps) return false;
        if (k != i) { for (int j = i; j < m; ++j) std::swap(A[i * m + j], A[k * m + j]); std::swap(b[i], b[k]); }
        const double d = -1 / A[i * m + i];
        for (int j = i + 1; j < m; ++j) {
            const double alpha = A[j * m + i] * d;
            for (int q = i + 1; q < m; ++q) A[j * m + q] += alpha * A[i * m + q];
            b[j] += alpha * b[i];
        }
    }
    for (int i = m - 1; i >= 0; --i) {
        double sum = b[i];
        for (int q = i + 1; q < m; ++q) sum -= A[i * m + q] * b[q];
        b[i] = sum / A[i * m + i];
    }
    return true;
}

static inline dim3 grid2(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y); }

// VoronoiSeamFinder::find on device masks (pairs in the reference's order: every pair sees the masks the earlier ones left)
static int voronoi_device(int n, const int *sizes, const int *corners, uint8_t *const *d_masks, cudaStream_t st)
{
    int *g = nullptr;
    size_t cap = 0;
    int rc = VSB_OK;
    for (int a = 0; a < n - 1 && rc == VSB_OK; ++a)
        for (int b = a + 1; b < n && rc == VSB_OK; ++b) {
            VorPair P;
            P.m1 = d_masks[a]; P.m2 = d_masks[b];
            P.w1 = sizes[2 * a]; P.h1 = sizes[2 * a + 1]; P.w2 = sizes[2 * b]; P.h2 = sizes[2 * b + 1];
            P.t1x = corners[2 * a]; P.t1y = corners[2 * a + 1]; P.t2x = corners[2 * b]; P.t2y = corners[2 * b + 1];
            P.x_tl = std::max(P.t1x, P.t2x); P.y_tl = std::max(P.t1y, P.t2y);
            const int x_br = std::min(P.t1x + P.w1, P.t2x + P.w2), y_br = std::min(P.t1y + P.h1, P.t2y + P.h2);
            if (!(P.x_tl < x_br && P.y_tl < y_br)) continue;
            P.rw = x_br - P.x_tl; P.rh = y_br - P.y_tl;
            const size_t need = (size_t)(P.rw + 2 * VOR_GAP) * (P.rh + 2 * VOR_GAP) * 2 * sizeof(int);
            if (need > cap) {
                if (g) cudaFreeAsync(g, st);
                rc = check_cuda(cudaMallocAsync(&g, need, st), "voronoi scratch");
                if (rc != VSB_OK) break;
                cap = need;
            }
            int *g1 = g, *g2 = g + (size_t)(P.rw + 2 * VOR_GAP) * (P.rh + 2 * VOR_GAP);
            k_vor_columns<<<(P.rw + 2 * VOR_GAP + 127) / 128, 128, 0, st>>>(P, g1, g2);
            k_vor_decide<<<grid2(P.rw, P.rh, dim3(32, 8)), dim3(32, 8), 0, st>>>(P, g1, g2);
            rc = check_launch("voronoi kernels");
        }
    if (g) cudaFreeAsync(g, st);
    return rc;
}

// ---- modular wrap-around ROI (wrapAround, 360_stitcher/defs.h:25; SURVEY.md section 7: "never allocate the full-width ROI") --------
// A camera that looks across +-pi gets the reference's full-panorama-width ROI: its two parts sit at the two ends of one image
// with zeros between them.  plan_parts replaces such a camera by TWO views -- the column ranges [0, a + m) and [b - m rounded down to a
// multiple of 2^num_bands, w), where [a, b) is the widest run of columns no projection-map entry fills from the camera image and
// m = 3 * 2^num_bands + 8 (the REFLECT gap of init_gpu, sources/modules/stitching/src/blenders.cpp:355, plus the reach of the mesh
// remap).  With that margin and that origin the Gaussian, Laplacian and weight levels of the parts equal the full-width view's
// wherever a weight is non-zero, so the panorama is the same bit for bit (a cut AT the content edge is not
// exact).  A camera whose empty run is shorter than 4 m stays one view.
struct ViewPart { int cam, x0, x1; };

static int plan_parts(int projection, float scale, const float *K, const float *R, int n, int src_w, int src_h, int num_bands_cfg,
                      const int *corners, const int *sizes, std::vector<ViewPart> &parts)
{
    int tlx = INT_MAX, tly = INT_MAX, brx = INT_MIN, bry = INT_MIN;
    for (int i = 0; i < n; ++i) {
        tlx = std::min(tlx, corners[2 * i]); tly = std::min(tly, corners[2 * i + 1]);
        brx = std::max(brx, corners[2 * i] + sizes[2 * i]); bry = std::max(bry, corners[2 * i + 1] + sizes[2 * i + 1]);
    }
    const int W = brx - tlx, H = bry - tly;
    // MultiBandBlender::prepare (sources/modules/stitching/src/blenders.cpp:237-247): bands limited by the panorama size
    const int nb = std::min(num_bands_cfg, (int)std::ceil(std::log((double)std::max(W, H)) / std::log(2.0)));
    const int unit = 1 << nb, m = 3 * unit + 8;
    parts.clear();
    for (int i = 0; i < n; ++i) {
        const int w = sizes[2 * i], h = sizes[2 * i + 1];
        if (w < W - 1) { parts.push_back({i, 0, w}); continue; }
        std::vector<float> xm((size_t)w * h), ym((size_t)w * h);
        host_build_maps(projection, scale, K + 9 * i, R + 9 * i, corners[2 * i], corners[2 * i + 1], w, h, xm.data(), ym.data());
        std::vector<uint8_t> col(w, 0);
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const int sx = f2i_rz(xm[(size_t)y * w + x]), sy = f2i_rz(ym[(size_t)y * w + x]);
                if (sx >= 0 && sx < src_w && sy >= 0 && sy < src_h) col[x] = 1;
            }
        int best_a = -1, best_b = -1, run_start = -1;   // widest interior run [a, b) of empty columns between two filled ones
        bool seen = false;
        for (int x = 0; x < w; ++x) {
            if (col[x]) {
                if (seen && run_start >= 0 && x - run_start > best_b - best_a) { best_a = run_start; best_b = x; }
                seen = true; run_start = -1;
            } else if (seen && run_start < 0) run_start = x;
        }
        if (best_a < 0 || best_b - best_a <= 4 * m) { parts.push_back({i, 0, w}); continue; }
        parts.push_back({i, 0, best_a + m});
        parts.push_back({i, (best_b - m) / unit * unit, w});
    }
    return nb;
}

}  // namespace vsb



extern "C" {

// calibrateCameras (360_stitcher/calibration.cpp:28-68) for view i of n, work_scale = 1
int vsb_rig_camera(int n_views, int i, int src_w, int src_h, double hfov_deg, float K[9], float R[9])
{
    if (!K || !R || n_views < 1 || i < 0 || i >= n_views || src_w <= 0 || src_h <= 0 || !(hfov_deg > 0 && hfov_deg < 180))
        return vsb::fail(VSB_ERR_INVALID, "rig_camera: bad arguments");
    const double PI = 3.1415926535897932384626;
    const double fov = hfov_deg * PI / 180.0;
    const double focal_tmp = 1.0 / std::tan(fov * 0.5);
    const float rot = static_cast<float>(2.0 * PI * static_cast<float>(i) / n_views);
    const double ppx = src_w / 2.0, ppy = src_h / 2.0, focal = focal_tmp * ppx;
    const float k[9] = {(float)focal, 0.f, (float)ppx, 0.f, (float)focal, (float)ppy, 0.f, 0.f, 1.f};
    const float r[9] = {(float)std::cos(rot), 0.f, (float)std::sin(rot), 0.f, 1.f, 0.f, (float)-std::sin(rot), 0.f, (float)std::cos(rot)};
    std::memcpy(K, k, sizeof(k));
    std::memcpy(R, r, sizeof(r));
    return VSB_OK;
}

// The camera of vsb_rig_camera after A/calibration.cpp:168-172 (focal, ppx, ppy *= compose_work_aspect, in double; work_scale = 1) and
// K().convertTo(CV_32F) (:175-176)
int vsb_rig_camera_scaled(int n_views, int i, int src_w, int src_h, double hfov_deg, double compose_work_aspect, float K[9], float R[9])
{
    int r = vsb_rig_camera(n_views, i, src_w, src_h, hfov_deg, K, R);
    if (r != VSB_OK) return r;
    if (!(compose_work_aspect > 0)) return vsb::fail(VSB_ERR_INVALID, "rig_camera_scaled: compose_work_aspect must be > 0");
    const double PI = 3.1415926535897932384626;
    double ppx = src_w / 2.0, ppy = src_h / 2.0;
    double focal = (1.0 / std::tan(hfov_deg * PI / 180.0 * 0.5)) * ppx;
    focal *= compose_work_aspect; ppx *= compose_work_aspect; ppy *= compose_work_aspect;
    K[0] = (float)focal; K[2] = (float)ppx; K[4] = (float)focal; K[5] = (float)ppy;
    return VSB_OK;
}

// calibrateCameras at a work scale (A/calibration.cpp:54-60: ppx = (full.width * work_scale) / 2, focal = focal_tmp * ppx, in double),
// then focal, ppx, ppy *= aspect (:168-172: compose_work_aspect = compose_scale / work_scale; 1 for the work-scale camera itself) and
// K().convertTo(CV_32F).  work_scale = aspect = 1 is vsb_rig_camera, work_scale = 1 is vsb_rig_camera_scaled.
int vsb_rig_camera_work(int n_views, int i, int src_w, int src_h, double hfov_deg, double work_scale, double aspect, float K[9], float R[9])
{
    int r = vsb_rig_camera(n_views, i, src_w, src_h, hfov_deg, K, R);
    if (r != VSB_OK) return r;
    if (!(work_scale > 0) || !(aspect > 0)) return vsb::fail(VSB_ERR_INVALID, "rig_camera_work: work_scale and aspect must be > 0");
    const double PI = 3.1415926535897932384626;
    double ppx = (src_w * work_scale) / 2.0, ppy = (src_h * work_scale) / 2.0;
    double focal = (1.0 / std::tan(hfov_deg * PI / 180.0 * 0.5)) * ppx;
    focal *= aspect; ppx *= aspect; ppy *= aspect;
    K[0] = (float)focal; K[2] = (float)ppx; K[4] = (float)focal; K[5] = (float)ppy;
    return VSB_OK;
}

// work_scale and compose_scale as stitch_calib / warpImages derive them from WORK_MEGAPIX / COMPOSE_MEGAPIX (A/calibration.cpp:270-277,
// 140-143; defaults 0.6 / 1.4, A/defs.h:51-53): min(1, sqrt(MEGAPIX * 1e6 / area)); a negative value means scale 1.
int vsb_ref_scales(int src_w, int src_h, double work_megapix, double compose_megapix, double *work_scale, double *compose_scale)
{
    if (src_w <= 0 || src_h <= 0) return vsb::fail(VSB_ERR_INVALID, "ref_scales: bad frame size");
    const double area = (double)(src_w * src_h);
    if (work_scale) *work_scale = work_megapix < 0 ? 1.0 : std::min(1.0, std::sqrt(work_megapix * 1e6 / area));
    if (compose_scale) *compose_scale = compose_megapix <= 0 ? 1.0 : std::min(1.0, std::sqrt(compose_megapix * 1e6 / area));
    return VSB_OK;
}

// The sizes compose_scale implies, exactly as the reference derives them: frame[2] = the frame remap #1 reads -- the caller's
// full frame, or cvRound(full * scale) when |scale - 1| > 0.1 (A/calibration.cpp:157-161; the same test decides the per-frame
// cuda::resize, A/timed.cpp:75-77, whose dsize is that cvRound) -- which also sizes the blender (:176-178); map_src[2] = (int)(full *
// scale), the img_size the maps and the warped masks are ALWAYS built for (:203-204).  For most scales the two agree; where they do
// not (cvRound != truncation, or a scale within 0.1 of 1) the reference blends views whose maps were built for a slightly
// different frame, and so does this library.
int vsb_compose_size(int full_w, int full_h, double compose_scale, int frame[2], int map_src[2], int *resized)
{
    if (full_w <= 0 || full_h <= 0 || !(compose_scale > 0) || !frame || !map_src) return vsb::fail(VSB_ERR_INVALID, "compose_size: bad arguments");
    const bool scaled = std::fabs(compose_scale - 1) > 1e-1;
    frame[0] = scaled ? (int)std::nearbyint(full_w * compose_scale) : full_w;
    frame[1] = scaled ? (int)std::nearbyint(full_h * compose_scale) : full_h;
    map_src[0] = (int)(full_w * compose_scale);
    map_src[1] = (int)(full_h * compose_scale);
    if (resized) *resized = scaled ? 1 : 0;
    if (frame[0] < 2 || frame[1] < 2 || map_src[0] < 2 || map_src[1] < 2) return vsb::fail(VSB_ERR_INVALID, "compose_size: compose_scale %g leaves no frame", compose_scale);
    return VSB_OK;
}

int vsb_host_build_maps(int projection, float scale, const float K[9], const float R[9], int tl_x, int tl_y, int w, int h, float *xmap, float *ymap)
{
    if (!K || !R || !xmap || !ymap || w <= 0 || h <= 0 || !(scale > 0) || (projection != VSB_PROJ_SPHERICAL && projection != VSB_PROJ_CYLINDRICAL))
        return vsb::fail(VSB_ERR_INVALID, "host_build_maps: bad arguments");
    vsb::host_build_maps(projection, scale, K, R, tl_x, tl_y, w, h, xmap, ymap);
    return VSB_OK;
}

int vsb_voronoi_seams(int n, const int *sizes_wh, const int *corners_xy, uint8_t *const *masks)
{
    if (n < 1 || !sizes_wh || !corners_xy || !masks) return vsb::fail(VSB_ERR_INVALID, "voronoi_seams: bad arguments");
    vsb::voronoi(n, sizes_wh, corners_xy, masks);
    return VSB_OK;
}

// n_cameras = 0: one view per camera (cfg.num_views cameras).  n_cameras > 0: the split calibration -- cameras that wrap around +-pi
// become two views (plan_parts), cfg.num_views must equal the number of parts (vsb_split_plan tells it before vsb_create).
// pano_width > 0: sphere radius pano_width / 2 pi (work_scale must be 1).  pano_width = 0: the reference's own warped_image_scale =
// (float)cameras[0].focal at work_scale (stitch_calib, A/calibration.cpp:283-289).
static int calibrate_rig_host(vsb_stitcher *s, int projection, int pano_width, int src_w, int src_h, double hfov_deg, const float *gains,
                              double compose_scale, int n_cameras = 0, double work_scale = 1.0)
{
    using namespace vsb;
    if (!s || pano_width < 0 || src_w <= 0 || src_h <= 0 || !(work_scale > 0) || (pano_width > 0 && work_scale != 1.0))
        return fail(VSB_ERR_INVALID, "calibrate_rig: bad arguments");
    vsb_config cfg;
    int r = vsb_get_config(s, &cfg);
    if (r != VSB_OK) return r;
    int frame_sz[2], map_src[2];
    r = vsb_compose_size(src_w, src_h, compose_scale, frame_sz, map_src, nullptr);
    if (r != VSB_OK) return r;
    const bool split = n_cameras > 0;
    const int n = split ? n_cameras : cfg.num_views;
    if (split && (compose_scale != 1.0 || n > VSB_MAX_VIEWS)) return fail(VSB_ERR_INVALID, "calibrate_rig_split: bad arguments (compose_scale must be 1)");
    std::vector<float> K(9 * n), R(9 * n);
    for (int i = 0; i < n; ++i) {   // the cameras at work scale (for work_scale = 1 exactly vsb_rig_camera's)
        r = vsb_rig_camera_work(n, i, src_w, src_h, hfov_deg, work_scale, 1.0, &K[9 * i], &R[9 * i]);
        if (r != VSB_OK) return r;
    }
    float scale = pano_width > 0 ? (float)(pano_width / (2.0 * 3.1415926535897932384626))  // sphere radius: pano_width px per 2*pi
                                 : K[0];                                                     // warped_image_scale = (float)cameras[0].focal
    // ---- seam scale (360_stitcher/calibration.cpp:92-135; SEAM_MEAGPIX = 0.01, 360_stitcher/defs.h:52)
    const double seam_scale = std::min(1.0, std::sqrt(0.01 * 1e6 / ((double)src_w * src_h)));
    const double seam_work_aspect = seam_scale / work_scale;   // :280
    const int seam_w = (int)std::nearbyint(src_w * seam_scale), seam_h = (int)std::nearbyint(src_h * seam_scale);
    const float seam_warp_scale = static_cast<float>(scale * seam_work_aspect);
    const float swa = (float)seam_work_aspect;
    std::vector<std::vector<uint8_t>> seam_masks(n);
    std::vector<int> seam_sizes(2 * n), seam_corners(2 * n);
    for (int i = 0; i < n; ++i) {
        float Ks[9];
        std::memcpy(Ks, &K[9 * i], sizeof(Ks));
        Ks[0] *= swa; Ks[2] *= swa; Ks[4] *= swa; Ks[5] *= swa;
        int roi[4];
        r = vsb_warp_roi(projection, seam_warp_scale, Ks, &R[9 * i], seam_w, seam_h, roi);
        if (r != VSB_OK) return r;
        std::vector<float> xm((size_t)roi[2] * roi[3]), ym((size_t)roi[2] * roi[3]);
        host_build_maps(projection, seam_warp_scale, Ks, &R[9 * i], roi[0], roi[1], roi[2], roi[3], xm.data(), ym.data());
        seam_masks[i].resize((size_t)roi[2] * roi[3]);
        host_warp_full_mask(xm.data(), ym.data(), roi[2], roi[3], seam_w, seam_h, seam_masks[i].data());
        seam_corners[2 * i] = roi[0]; seam_corners[2 * i + 1] = roi[1];
        seam_sizes[2 * i] = roi[2]; seam_sizes[2 * i + 1] = roi[3];
    }
    {
        std::vector<uint8_t *> ptrs(n);
        for (int i = 0; i < n; ++i) ptrs[i] = seam_masks[i].data();
        voronoi(n, seam_sizes.data(), seam_corners.data(), ptrs.data());
    }
    // ---- compose scale (360_stitcher/calibration.cpp:137-246)
    const int full_w = src_w, full_h = src_h;
    const double compose_work_aspect = compose_scale / work_scale;   // :148
    if (compose_work_aspect != 1.0) {
        // warper scale: warped_image_scale * static_cast<float>(compose_work_aspect) (:151); cameras: focal, ppx, ppy *= aspect (:168-172)
        scale = scale * static_cast<float>(compose_work_aspect);
        for (int i = 0; i < n; ++i) {
            r = vsb_rig_camera_work(n, i, full_w, full_h, hfov_deg, work_scale, compose_work_aspect, &K[9 * i], &R[9 * i]);
            if (r != VSB_OK) return r;
        }
    }
    // corners + the sizes prepare() gets: warpRoi of the frame remap #1 reads (:176-178); maps and masks: img_size (:203-204, vsb_compose_size)
    std::vector<int> corners(2 * n), sizes(2 * n), map_roi(4 * n);
    for (int i = 0; i < n; ++i) {
        int roi[4];
        r = vsb_warp_roi(projection, scale, &K[9 * i], &R[9 * i], frame_sz[0], frame_sz[1], roi);
        if (r != VSB_OK) return r;
        corners[2 * i] = roi[0]; corners[2 * i + 1] = roi[1]; sizes[2 * i] = roi[2]; sizes[2 * i + 1] = roi[3];
        r = vsb_warp_roi(projection, scale, &K[9 * i], &R[9 * i], map_src[0], map_src[1], &map_roi[4 * i]);
        if (r != VSB_OK) return r;
    }
    if (split) {
        // ---- the split calibration: the same per-camera products, installed as column windows of the cameras (plan_parts)
        std::vector<ViewPart> parts;
        plan_parts(projection, scale, K.data(), R.data(), n, src_w, src_h, cfg.num_bands, corners.data(), sizes.data(), parts);
        const int nv = (int)parts.size();
        if (nv != cfg.num_views)
            return fail(VSB_ERR_INVALID, "calibrate_rig_split: this rig needs %d views (%d cameras, %d of them split), the handle has %d: size it with vsb_split_plan", nv, n, nv - n, cfg.num_views);
        std::vector<int> vc(2 * nv), vs(2 * nv);
        for (int k = 0; k < nv; ++k) {
            const ViewPart &P = parts[k];
            vc[2 * k] = corners[2 * P.cam] + P.x0; vc[2 * k + 1] = corners[2 * P.cam + 1];
            vs[2 * k] = P.x1 - P.x0; vs[2 * k + 1] = sizes[2 * P.cam + 1];
        }
        r = vsb_prepare(s, vc.data(), vs.data());
        if (r != VSB_OK) return r;
        std::vector<float> xm, ym;
        std::vector<uint8_t> warped, seam;
        int have = -1;
        for (int k = 0; k < nv; ++k) {
            const ViewPart &P = parts[k];
            const int i = P.cam, w = sizes[2 * i], h = sizes[2 * i + 1];
            if (have != i) {   // the camera's full-width products (both parts of a camera are adjacent in the list)
                xm.resize((size_t)w * h); ym.resize((size_t)w * h); warped.resize((size_t)w * h); seam.resize((size_t)w * h);
                host_build_maps(projection, scale, &K[9 * i], &R[9 * i], corners[2 * i], corners[2 * i + 1], w, h, xm.data(), ym.data());
                host_warp_full_mask(xm.data(), ym.data(), w, h, src_w, src_h, warped.data());
                const int sw = seam_sizes[2 * i], sh = seam_sizes[2 * i + 1];
                std::vector<uint8_t> dil(seam_masks[i].size());
                if (cfg.enable_local) dilate3x3(seam_masks[i].data(), sw, sh, dil.data());
                else dil = seam_masks[i];
                resize_linear_u8(dil.data(), sw, sh, seam.data(), w, h);
                for (size_t j = 0; j < seam.size(); ++j) seam[j] &= warped[j];
                have = i;
            }
            const int pw = P.x1 - P.x0;
            r = vsb_init_view(s, k, seam.data() + P.x0, pw, h, (size_t)w, vc[2 * k], vc[2 * k + 1], 0);
            if (r != VSB_OK) return r;
            r = vsb_set_maps(s, k, xm.data() + P.x0, ym.data() + P.x0, pw, h, (size_t)w * 4, 0, src_w, src_h);
            if (r != VSB_OK) return r;
            r = vsb_set_view_window(s, k, i, P.x0, w);
            if (r != VSB_OK) return r;
            if (gains) { r = vsb_set_gain(s, k, gains[i]); if (r != VSB_OK) return r; }
        }
        r = vsb_set_compose_scale(s, 1.0, src_w, src_h);
        if (r != VSB_OK) return r;
        return vsb_note_rig(s, projection, scale, src_w, src_h);
    }
    r = vsb_prepare(s, corners.data(), sizes.data());
    if (r != VSB_OK) return r;
    for (int i = 0; i < n; ++i) {
        const int w = map_roi[4 * i + 2], h = map_roi[4 * i + 3];
        std::vector<float> xm((size_t)w * h), ym((size_t)w * h);
        host_build_maps(projection, scale, &K[9 * i], &R[9 * i], map_roi[4 * i], map_roi[4 * i + 1], w, h, xm.data(), ym.data());
        std::vector<uint8_t> warped((size_t)w * h), seam((size_t)w * h);
        host_warp_full_mask(xm.data(), ym.data(), w, h, map_src[0], map_src[1], warped.data());
        const int sw = seam_sizes[2 * i], sh = seam_sizes[2 * i + 1];
        std::vector<uint8_t> dil(seam_masks[i].size());
        if (cfg.enable_local) dilate3x3(seam_masks[i].data(), sw, sh, dil.data());
        else dil = seam_masks[i];
        resize_linear_u8(dil.data(), sw, sh, seam.data(), w, h);
        for (size_t j = 0; j < seam.size(); ++j) seam[j] &= warped[j];
        r = vsb_init_view(s, i, seam.data(), w, h, (size_t)w, corners[2 * i], corners[2 * i + 1], 0);
        if (r != VSB_OK) return r;
        r = vsb_set_maps(s, i, xm.data(), ym.data(), w, h, (size_t)w * 4, 0, frame_sz[0], frame_sz[1]);   // cuda::remap tests against the frame it is given
        if (r != VSB_OK) return r;
        if (gains) { r = vsb_set_gain(s, i, gains[i]); if (r != VSB_OK) return r; }
    }
    r = vsb_set_compose_scale(s, compose_scale, full_w, full_h);
    if (r != VSB_OK) return r;
    return vsb_note_rig(s, projection, scale, full_w, full_h);
}

int vsb_calibrate_rig(vsb_stitcher *s, int projection, int pano_width, int src_w, int src_h, double hfov_deg, const float *gains)
{
    if (pano_width <= 0) return vsb::fail(VSB_ERR_INVALID, "calibrate_rig: bad arguments");
    return calibrate_rig_host(s, projection, pano_width, src_w, src_h, hfov_deg, gains, 1.0);
}

// The views vsb_calibrate_rig_split makes of n_cameras cameras (host only, no device): view k is the column window
// [view_x0[k], view_x0[k] + view_w[k]) of camera view_camera[k]'s warped image.  num_bands is the configured band count.
int vsb_split_plan(int projection, int pano_width, int n_cameras, int src_w, int src_h, double hfov_deg, int num_bands, int *n_views,
                   int *view_camera, int *view_x0, int *view_w)
{
    using namespace vsb;
    if (n_cameras < 1 || n_cameras > VSB_MAX_VIEWS || pano_width <= 0 || src_w <= 0 || src_h <= 0 || num_bands < 1 || num_bands > VSB_MAX_BANDS || !n_views)
        return fail(VSB_ERR_INVALID, "split_plan: bad arguments");
    const int n = n_cameras;
    const float scale = (float)(pano_width / (2.0 * 3.1415926535897932384626));
    std::vector<float> K(9 * n), R(9 * n);
    std::vector<int> corners(2 * n), sizes(2 * n);
    for (int i = 0; i < n; ++i) {
        int roi[4];
        int r = vsb_rig_camera(n, i, src_w, src_h, hfov_deg, &K[9 * i], &R[9 * i]);
        if (r == VSB_OK) r = vsb_warp_roi(projection, scale, &K[9 * i], &R[9 * i], src_w, src_h, roi);
        if (r != VSB_OK) return r;
        corners[2 * i] = roi[0]; corners[2 * i + 1] = roi[1]; sizes[2 * i] = roi[2]; sizes[2 * i + 1] = roi[3];
    }
    std::vector<ViewPart> parts;
    plan_parts(projection, scale, K.data(), R.data(), n, src_w, src_h, num_bands, corners.data(), sizes.data(), parts);
    *n_views = (int)parts.size();
    if (*n_views > VSB_MAX_VIEWS) return fail(VSB_ERR_INVALID, "split_plan: %d views exceed VSB_MAX_VIEWS", *n_views);
    for (int k = 0; k < *n_views; ++k) {
        if (view_camera) view_camera[k] = parts[k].cam;
        if (view_x0) view_x0[k] = parts[k].x0;
        if (view_w) view_w[k] = parts[k].x1 - parts[k].x0;
    }
    return VSB_OK;
}

static int calibrate_rig_dev(vsb_stitcher *s, int projection, int pano_width, int src_w, int src_h, double hfov_deg, const float *gains,
                             double compose_scale, int n_cameras = 0, double work_scale = 1.0);

int vsb_calibrate_rig_split(vsb_stitcher *s, int projection, int pano_width, int n_cameras, int src_w, int src_h, double hfov_deg,
                            const float *gains, int on_device)
{
    if (n_cameras < 1 || pano_width <= 0) return vsb::fail(VSB_ERR_INVALID, "calibrate_rig_split: n_cameras must be >= 1, pano_width > 0");
    return on_device ? calibrate_rig_dev(s, projection, pano_width, src_w, src_h, hfov_deg, gains, 1.0, n_cameras)
                     : calibrate_rig_host(s, projection, pano_width, src_w, src_h, hfov_deg, gains, 1.0, n_cameras);
}

// ---- device-side calibration entry points -------------------------------------------------------------------------------------
int vsb_voronoi_seams_device(int n, const int *sizes_wh, const int *corners_xy, uint8_t *const *d_masks, void *stream)
{
    if (n < 1 || n > VSB_MAX_VIEWS || !sizes_wh || !corners_xy || !d_masks) return vsb::fail(VSB_ERR_INVALID, "voronoi_seams_device: bad arguments");
    return vsb::voronoi_device(n, sizes_wh, corners_xy, d_masks, (cudaStream_t)stream);
}

int vsb_dilate3x3_u8(const uint8_t *d_src, int w, int h, uint8_t *d_dst, void *stream)
{
    if (!d_src || !d_dst || w <= 0 || h <= 0) return vsb::fail(VSB_ERR_INVALID, "dilate3x3: bad arguments");
    const dim3 b(32, 8);
    vsb::k_dilate3x3<<<vsb::grid2(w, h, b), b, 0, (cudaStream_t)stream>>>(d_src, w, h, d_dst);
    return vsb::check_launch("k_dilate3x3");
}

int vsb_resize_linear_u8(const uint8_t *d_src, int sw, int sh, size_t src_pitch, int channels, uint8_t *d_dst, int dw, int dh, size_t dst_pitch,
                         double fx, double fy, void *stream)
{
    if (!d_src || !d_dst || sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0 || (channels != 1 && channels != 3))
        return vsb::fail(VSB_ERR_INVALID, "resize_linear: bad arguments");
    // cuda::resize (sources/modules/cudawarping/src/resize.cpp:76-105): with fx = fy = 0 the factors follow from the sizes
    if (!(fx > 0) || !(fy > 0)) { fx = (double)dw / sw; fy = (double)dh / sh; }
    cudaStream_t st = (cudaStream_t)stream;
    if (dw == sw && dh == sh) return vsb::check_cuda(cudaMemcpy2DAsync(d_dst, dst_pitch, d_src, src_pitch, (size_t)sw * channels, sh, cudaMemcpyDeviceToDevice, st), "resize (copy)");
    const dim3 b(32, 8);
    const float kx = (float)(1.0 / fx), ky = (float)(1.0 / fy);
    if (channels == 1) vsb::k_resize_linear_u8<1><<<vsb::grid2(dw, dh, b), b, 0, st>>>(d_src, sw, sh, src_pitch, d_dst, dw, dh, dst_pitch, kx, ky);
    else vsb::k_resize_linear_u8<3><<<vsb::grid2(dw, dh, b), b, 0, st>>>(d_src, sw, sh, src_pitch, d_dst, dw, dh, dst_pitch, kx, ky);
    return vsb::check_launch("k_resize_linear_u8");
}

// vsb_calibrate_rig with every per-pixel loop on the device: seam-scale and compose-scale maps (k_build_maps), mask warps,
// Voronoi seams, dilate / resize / AND, weight pyramids (vsb_init_view with a device mask).  The ROIs stay on the host like in the
// reference (RotationWarperBase::detectResultRoi runs on the CPU there too: a walk along the image border).  The device evaluates
// sinf / cosf itself, so the projection maps agree with the host path to ~1e-3 px, not bit for bit (the reference's own maps come
// from the same kind of device code); everything downstream of the maps is exact.
static int calibrate_rig_dev(vsb_stitcher *s, int projection, int pano_width, int src_w, int src_h, double hfov_deg, const float *gains,
                             double compose_scale, int n_cameras, double work_scale)
{
    using namespace vsb;
    if (!s || pano_width < 0 || src_w <= 0 || src_h <= 0 || !(work_scale > 0) || (pano_width > 0 && work_scale != 1.0))
        return fail(VSB_ERR_INVALID, "calibrate_rig_device: bad arguments");
    vsb_config cfg;
    int r = vsb_get_config(s, &cfg);
    if (r != VSB_OK) return r;
    int frame_sz[2], map_src[2];
    r = vsb_compose_size(src_w, src_h, compose_scale, frame_sz, map_src, nullptr);
    if (r != VSB_OK) return r;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(vsb_handle_device(s));
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev};
    const bool split = n_cameras > 0;   // cameras that wrap around +-pi become two views, as in calibrate_rig_host
    const int n = split ? n_cameras : cfg.num_views;
    if (split && (compose_scale != 1.0 || n > VSB_MAX_VIEWS)) return fail(VSB_ERR_INVALID, "calibrate_rig_split: bad arguments (compose_scale must be 1)");
    CalibState *cs = new CalibState();
    cs->n = n; cs->projection = projection; cs->src_w = src_w; cs->src_h = src_h;
    std::vector<float> K(9 * n), R(9 * n);
    for (int i = 0; i < n; ++i) {   // the cameras at work scale; scales as in calibrate_rig_host
        r = vsb_rig_camera_work(n, i, src_w, src_h, hfov_deg, work_scale, 1.0, &K[9 * i], &R[9 * i]);
        if (r != VSB_OK) { calib_state_free(cs); return r; }
    }
    float scale = pano_width > 0 ? (float)(pano_width / (2.0 * 3.1415926535897932384626)) : K[0];
    cudaStream_t st = nullptr;  // calibration time: the default stream, synchronised at the end of every phase
    const dim3 b(32, 8);
    auto bail = [&](int code) { cudaDeviceSynchronize(); calib_state_free(cs); return code; };
    // ---- seam scale (360_stitcher/calibration.cpp:92-135)
    const double seam_scale = std::min(1.0, std::sqrt(0.01 * 1e6 / ((double)src_w * src_h)));
    const int seam_w = (int)std::nearbyint(src_w * seam_scale), seam_h = (int)std::nearbyint(src_h * seam_scale);
    const double seam_work_aspect = seam_scale / work_scale;
    const float seam_warp_scale = static_cast<float>(scale * seam_work_aspect), swa = (float)seam_work_aspect;
    cs->seam_scale = seam_scale; cs->seam_w = seam_w; cs->seam_h = seam_h; cs->seam_warp_scale = seam_warp_scale;
    std::vector<int> seam_sizes(2 * n), seam_corners(2 * n);
    std::vector<uint8_t *> seam_masks(n, nullptr);
    auto free_seams = [&]() { for (uint8_t *p : seam_masks) cudaFree(p); };
    for (int i = 0; i < n; ++i) {
        float *Ks = cs->Ks[i];
        std::memcpy(Ks, &K[9 * i], sizeof(float) * 9);
        std::memcpy(cs->R[i], &R[9 * i], sizeof(float) * 9);
        Ks[0] *= swa; Ks[2] *= swa; Ks[4] *= swa; Ks[5] *= swa;
        int *roi = cs->roi[i];
        r = vsb_warp_roi(projection, seam_warp_scale, Ks, &R[9 * i], seam_w, seam_h, roi);
        if (r != VSB_OK) { free_seams(); return bail(r); }
        const size_t mp = ((size_t)roi[2] * 4 + 15) / 16 * 16;
        float *xm = nullptr, *ym = nullptr;
        if (cudaMalloc(&xm, mp * roi[3]) != cudaSuccess || cudaMalloc(&ym, mp * roi[3]) != cudaSuccess ||
            cudaMalloc(&seam_masks[i], (size_t)roi[2] * roi[3]) != cudaSuccess || cudaMalloc(&cs->warped_mask[i], (size_t)roi[2] * roi[3]) != cudaSuccess) {
            cudaFree(xm); cudaFree(ym); free_seams();
            return bail(fail(VSB_ERR_NOMEM, "calibrate_rig_device: out of device memory"));
        }
        int roi2[4];
        r = vsb_build_maps(projection, seam_warp_scale, Ks, &R[9 * i], seam_w, seam_h, xm, ym, mp, roi2, st);
        if (r == VSB_OK) {
            k_full_mask<<<grid2(roi[2], roi[3], b), b, 0, st>>>(xm, ym, mp, roi[2], roi[3], seam_w, seam_h, seam_masks[i]);
            r = check_launch("k_full_mask");
        }
        if (r == VSB_OK) r = check_cuda(cudaMemcpyAsync(cs->warped_mask[i], seam_masks[i], (size_t)roi[2] * roi[3], cudaMemcpyDeviceToDevice, st), "seam mask copy");
        cudaStreamSynchronize(st);
        cudaFree(xm); cudaFree(ym);
        if (r != VSB_OK) { free_seams(); return bail(r); }
        seam_corners[2 * i] = roi[0]; seam_corners[2 * i + 1] = roi[1]; seam_sizes[2 * i] = roi[2]; seam_sizes[2 * i + 1] = roi[3];
    }
    r = voronoi_device(n, seam_sizes.data(), seam_corners.data(), seam_masks.data(), st);
    if (r != VSB_OK) { free_seams(); return bail(r); }
    // ---- compose scale (360_stitcher/calibration.cpp:137-246); the sizes as in calibrate_rig_host
    const int full_w = src_w, full_h = src_h;
    const double compose_work_aspect = compose_scale / work_scale;
    if (compose_work_aspect != 1.0) {
        scale = scale * static_cast<float>(compose_work_aspect);
        for (int i = 0; i < n; ++i) {
            r = vsb_rig_camera_work(n, i, full_w, full_h, hfov_deg, work_scale, compose_work_aspect, &K[9 * i], &R[9 * i]);
            if (r != VSB_OK) { free_seams(); return bail(r); }
        }
    }
    std::vector<int> corners(2 * n), sizes(2 * n), map_roi(4 * n);
    for (int i = 0; i < n; ++i) {
        int roi[4];
        r = vsb_warp_roi(projection, scale, &K[9 * i], &R[9 * i], frame_sz[0], frame_sz[1], roi);
        if (r == VSB_OK) r = vsb_warp_roi(projection, scale, &K[9 * i], &R[9 * i], map_src[0], map_src[1], &map_roi[4 * i]);
        if (r != VSB_OK) { free_seams(); return bail(r); }
        corners[2 * i] = roi[0]; corners[2 * i + 1] = roi[1]; sizes[2 * i] = roi[2]; sizes[2 * i + 1] = roi[3];
    }
    // the views: one per camera, or (split) the column windows plan_parts makes of the cameras that wrap around
    std::vector<ViewPart> parts;
    if (split) plan_parts(projection, scale, K.data(), R.data(), n, src_w, src_h, cfg.num_bands, corners.data(), sizes.data(), parts);
    else for (int i = 0; i < n; ++i) parts.push_back({i, 0, map_roi[4 * i + 2]});
    const int nv = (int)parts.size();
    if (nv != cfg.num_views) {
        free_seams();
        return bail(fail(VSB_ERR_INVALID, "calibrate_rig_split: this rig needs %d views (%d cameras, %d of them split), the handle has %d: size it with vsb_split_plan", nv, n, nv - n, cfg.num_views));
    }
    if (split) {
        std::vector<int> vc(2 * nv), vs(2 * nv);
        for (int k = 0; k < nv; ++k) {
            vc[2 * k] = corners[2 * parts[k].cam] + parts[k].x0; vc[2 * k + 1] = corners[2 * parts[k].cam + 1];
            vs[2 * k] = parts[k].x1 - parts[k].x0; vs[2 * k + 1] = sizes[2 * parts[k].cam + 1];
        }
        r = vsb_prepare(s, vc.data(), vs.data());
    } else {
        r = vsb_prepare(s, corners.data(), sizes.data());
    }
    if (r != VSB_OK) { free_seams(); return bail(r); }
    for (int i = 0, k = 0; i < n && r == VSB_OK; ++i) {
        const int w = map_roi[4 * i + 2], h = map_roi[4 * i + 3], sw = seam_sizes[2 * i], sh = seam_sizes[2 * i + 1];
        const size_t mp = ((size_t)w * 4 + 15) / 16 * 16;
        float *xm = nullptr, *ym = nullptr;
        uint8_t *warped = nullptr, *seam = nullptr, *dil = nullptr;
        if (cudaMalloc(&xm, mp * h) != cudaSuccess || cudaMalloc(&ym, mp * h) != cudaSuccess || cudaMalloc(&warped, (size_t)w * h) != cudaSuccess ||
            cudaMalloc(&seam, (size_t)w * h) != cudaSuccess || cudaMalloc(&dil, (size_t)sw * sh) != cudaSuccess)
            r = fail(VSB_ERR_NOMEM, "calibrate_rig_device: out of device memory");
        int roi2[4];
        if (r == VSB_OK) r = vsb_build_maps(projection, scale, &K[9 * i], &R[9 * i], map_src[0], map_src[1], xm, ym, mp, roi2, st);
        if (r == VSB_OK) {
            k_full_mask<<<grid2(w, h, b), b, 0, st>>>(xm, ym, mp, w, h, map_src[0], map_src[1], warped);
            const uint8_t *small = seam_masks[i];
            if (cfg.enable_local) { k_dilate3x3<<<grid2(sw, sh, b), b, 0, st>>>(seam_masks[i], sw, sh, dil); small = dil; }
            r = check_launch("mask kernels");
            if (r == VSB_OK) r = vsb_resize_linear_u8(small, sw, sh, (size_t)sw, 1, seam, w, h, (size_t)w, 0.0, 0.0, st);
            if (r == VSB_OK) {
                k_and_u8<<<(unsigned)(((size_t)w * h + 255) / 256), 256, 0, st>>>(seam, warped, (size_t)w * h);
                r = check_launch("k_and_u8");
            }
        }
        if (r == VSB_OK) r = check_cuda(cudaStreamSynchronize(st), "calibrate_rig_device");
        for (; k < nv && parts[k].cam == i && r == VSB_OK; ++k) {   // the camera's view(s): windows of its mask and maps, in place
            const int x0 = parts[k].x0, pw = parts[k].x1 - x0;
            r = vsb_init_view(s, k, seam + x0, pw, h, (size_t)w, corners[2 * i] + x0, corners[2 * i + 1], 1);
            if (r == VSB_OK) r = vsb_set_maps(s, k, xm + x0, ym + x0, pw, h, mp, 1, frame_sz[0], frame_sz[1]);
            if (r == VSB_OK && split) r = vsb_set_view_window(s, k, i, x0, w);
            if (r == VSB_OK && gains) r = vsb_set_gain(s, k, gains[i]);
        }
        cudaDeviceSynchronize();
        cudaFree(xm); cudaFree(ym); cudaFree(warped); cudaFree(seam); cudaFree(dil);
    }
    free_seams();
    if (r != VSB_OK) return bail(r);
    vsb_attach_calib(s, cs, calib_state_free);
    r = vsb_set_compose_scale(s, compose_scale, full_w, full_h);
    if (r != VSB_OK) return r;
    return vsb_note_rig(s, projection, scale, full_w, full_h);
}

int vsb_calibrate_rig_device(vsb_stitcher *s, int projection, int pano_width, int src_w, int src_h, double hfov_deg, const float *gains)
{
    if (pano_width <= 0) return vsb::fail(VSB_ERR_INVALID, "calibrate_rig_device: bad arguments");
    return calibrate_rig_dev(s, projection, pano_width, src_w, src_h, hfov_deg, gains, 1.0);
}

// Both calibrations with the reference's compose_scale (A/calibration.cpp:137-205): the cameras and the warper are scaled, the ROIs,
// maps and masks are built for the scaled frame, and every frame handed to vsb_feed / vsb_compose / vsb_submit_host (still
// src_w x src_h) goes through cuda::resize(..., Size(), compose_scale, compose_scale, INTER_LINEAR) first (A/timed.cpp:74-81).
// stitch_calib with its own constants (A/calibration.cpp:256-305; A/defs.h:51-53): work_scale from WORK_MEGAPIX, compose_scale from
// COMPOSE_MEGAPIX (vsb_ref_scales), the cameras of calibrateCameras at work scale, warped_image_scale = (float)cameras[0].focal,
// seam_work_aspect = seam_scale / work_scale, compose_work_aspect = compose_scale / work_scale -- the reference's default panorama
// geometry, which a pano_width cannot express (the sphere radius is a float derived from the focal length, not from an integer width).
int vsb_calibrate_rig_megapix(vsb_stitcher *s, int projection, int src_w, int src_h, double hfov_deg, const float *gains,
                              double work_megapix, double compose_megapix, int on_device)
{
    double ws = 1.0, cs = 1.0;
    int r = vsb_ref_scales(src_w, src_h, work_megapix, compose_megapix, &ws, &cs);
    if (r != VSB_OK) return r;
    return on_device ? calibrate_rig_dev(s, projection, 0, src_w, src_h, hfov_deg, gains, cs, 0, ws)
                     : calibrate_rig_host(s, projection, 0, src_w, src_h, hfov_deg, gains, cs, 0, ws);
}

int vsb_calibrate_rig_scaled(vsb_stitcher *s, int projection, int pano_width, int src_w, int src_h, double hfov_deg, const float *gains,
                             double compose_scale, int on_device)
{
    if (pano_width <= 0) return vsb::fail(VSB_ERR_INVALID, "calibrate_rig_scaled: bad arguments");
    return on_device ? calibrate_rig_dev(s, projection, pano_width, src_w, src_h, hfov_deg, gains, compose_scale)
                     : calibrate_rig_host(s, projection, pano_width, src_w, src_h, hfov_deg, gains, compose_scale);
}

// GainCompensator::feed (sources/modules/stitching/src/exposure_compensate.cpp:71-142) on warped images already on the device:
// d_imgs[i] CV_8UC3 (tight rows), d_masks[i] CV_8U (255 = valid), sizes / corners as the reference passes them.  The pairwise
// reductions run on the device (one thread per pair, the reference's summation order), the n x n solve on the host.
int vsb_gain_compensator_feed(int n, const uint8_t *const *d_imgs, const uint8_t *const *d_masks, const int *sizes_wh, const int *corners_xy,
                              double *gains_out, void *stream)
{
    using namespace vsb;
    if (n < 4 || n > VSB_MAX_VIEWS || !d_imgs || !d_masks || !sizes_wh || !corners_xy || !gains_out)
        return fail(VSB_ERR_INVALID, "gain_compensator_feed: bad arguments (4..%d views: cv::solve takes closed forms below 4, not restated)", VSB_MAX_VIEWS);
    cudaStream_t st = (cudaStream_t)stream;
    GainParams P;
    std::memset(&P, 0, sizeof(P));
    P.n = n;
    for (int i = 0; i < n; ++i) {
        P.v[i].img = d_imgs[i]; P.v[i].mask = d_masks[i];
        P.v[i].w = sizes_wh[2 * i]; P.v[i].h = sizes_wh[2 * i + 1]; P.v[i].tx = corners_xy[2 * i]; P.v[i].ty = corners_xy[2 * i + 1];
    }
    int *dN = nullptr;
    double *dI = nullptr;
    std::vector<int> N((size_t)n * n, 0);
    std::vector<double> I((size_t)n * n, 0.0);
    int r = VSB_OK;
    if (cudaMalloc(&dN, sizeof(int) * n * n) != cudaSuccess || cudaMalloc(&dI, sizeof(double) * n * n) != cudaSuccess)
        r = fail(VSB_ERR_NOMEM, "gain_compensator_feed: out of device memory");
    if (r == VSB_OK) {
        cudaMemsetAsync(dN, 0, sizeof(int) * n * n, st);
        cudaMemsetAsync(dI, 0, sizeof(double) * n * n, st);
        k_gain_pairs<<<(n * n + 63) / 64, 64, 0, st>>>(P, dN, dI);
        r = check_launch("k_gain_pairs");
        if (r == VSB_OK) r = check_cuda(cudaMemcpyAsync(N.data(), dN, sizeof(int) * n * n, cudaMemcpyDeviceToHost, st), "gain_compensator_feed");
        if (r == VSB_OK) r = check_cuda(cudaMemcpyAsync(I.data(), dI, sizeof(double) * n * n, cudaMemcpyDeviceToHost, st), "gain_compensator_feed");
        if (r == VSB_OK) r = check_cuda(cudaStreamSynchronize(st), "gain_compensator_feed");
    }
    cudaFree(dN); cudaFree(dI);
    if (r != VSB_OK) return r;
    const double alpha = 0.01, beta = 100;
    std::vector<double> A((size_t)n * n, 0.0), bb(n, 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            bb[i] += beta * N[i * n + j];
            A[i * n + i] += beta * N[i * n + j];
            if (j == i) continue;
            A[i * n + i] += 2 * alpha * I[i * n + j] * I[i * n + j] * N[i * n + j];
            A[i * n + j] -= 2 * alpha * I[i * n + j] * I[j * n + i] * N[i * n + j];
        }
    if (!lu_solve(A, bb, n)) return fail(VSB_ERR_INVALID, "gain_compensator_feed: singular system");
    for (int i = 0; i < n; ++i) gains_out[i] = bb[i];
    return VSB_OK;
}

// GainCompensator::feed (sources/modules/stitching/src/exposure_compensate.cpp:71-142) on the device, callable at run time
// ("Dynamically update the gain compensation" is an open TODO of the reference): the frames are resized to seam scale
// (cuda::resize, A/calibration.cpp:95) and warped LINEAR / BORDER_REFLECT (:118) with the seam-scale maps, the pairwise
// overlap counts and mean intensities are reduced on the device in the reference's summation order, the n x n system is solved
// on the host (hal::LU64f restated: bit-identical gains for n >= 4).  apply != 0 installs the gains (vsb_set_gain).
int vsb_estimate_gains(vsb_stitcher *s, const uint8_t *const *d_frames, size_t pitch, float *gains_out, int apply, void *stream)
{
    using namespace vsb;
    if (!s || !d_frames) return fail(VSB_ERR_INVALID, "estimate_gains: null argument");
    CalibState *cs = static_cast<CalibState *>(vsb_get_calib(s));
    if (!cs) return fail(VSB_ERR_STATE, "estimate_gains: the handle was not calibrated with vsb_calibrate_rig_device");
    const int n = cs->n;
    if (n < 4) return fail(VSB_ERR_INVALID, "estimate_gains: needs >= 4 views (cv::solve takes closed forms below that; not restated)");
    if (pitch < (size_t)cs->src_w * 3) return fail(VSB_ERR_INVALID, "estimate_gains: pitch too small");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(vsb_handle_device(s));
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev};
    cudaStream_t st = (cudaStream_t)stream;
    GainParams P;
    std::memset(&P, 0, sizeof(P));
    P.n = n;
    std::vector<uint8_t *> tmp;
    int r = VSB_OK;
    uint8_t *small = nullptr;
    auto cleanup = [&]() { cudaStreamSynchronize(st); for (uint8_t *p : tmp) cudaFree(p); cudaFree(small); };
    if (cudaMalloc(&small, (size_t)cs->seam_w * cs->seam_h * 3) != cudaSuccess) return fail(VSB_ERR_NOMEM, "estimate_gains: out of device memory");
    for (int i = 0; i < n && r == VSB_OK; ++i) {
        if (!d_frames[i]) { r = fail(VSB_ERR_INVALID, "estimate_gains: frame %d is null", i); break; }
        const int *roi = cs->roi[i];
        uint8_t *img = nullptr;
        if (cudaMalloc(&img, (size_t)roi[2] * roi[3] * 3) != cudaSuccess) { r = fail(VSB_ERR_NOMEM, "estimate_gains: out of device memory"); break; }
        tmp.push_back(img);
        // cuda::resize(img, seam_img, Size(), seam_scale, seam_scale, INTER_LINEAR): the factors are the scale itself
        r = vsb_resize_linear_u8(d_frames[i], cs->src_w, cs->src_h, pitch, 3, small, cs->seam_w, cs->seam_h, (size_t)cs->seam_w * 3, cs->seam_scale, cs->seam_scale, st);
        int roi2[4];
        if (r == VSB_OK) r = vsb_warp(cs->projection, cs->seam_warp_scale, cs->Ks[i], cs->R[i], small, cs->seam_w, cs->seam_h, (size_t)cs->seam_w * 3, 3,
                                      VSB_INTER_LINEAR, VSB_BORDER_REFLECT, img, (size_t)roi[2] * 3, roi2, st);
        P.v[i].img = img; P.v[i].mask = cs->warped_mask[i]; P.v[i].w = roi[2]; P.v[i].h = roi[3]; P.v[i].tx = roi[0]; P.v[i].ty = roi[1];
    }
    std::vector<double> g(n, 1.0);
    if (r == VSB_OK) {
        std::vector<const uint8_t *> imgs(n), masks(n);
        std::vector<int> sz(2 * n), co(2 * n);
        for (int i = 0; i < n; ++i) { imgs[i] = P.v[i].img; masks[i] = P.v[i].mask; sz[2 * i] = P.v[i].w; sz[2 * i + 1] = P.v[i].h; co[2 * i] = P.v[i].tx; co[2 * i + 1] = P.v[i].ty; }
        r = vsb_gain_compensator_feed(n, imgs.data(), masks.data(), sz.data(), co.data(), g.data(), st);
    }
    cleanup();
    if (r != VSB_OK) return r;
    std::vector<double> &bb = g;
    for (int i = 0; i < n; ++i)
        if (gains_out) gains_out[i] = (float)bb[i];
    if (apply) {   // one gain per CAMERA, installed on each of its views (a camera split by vsb_calibrate_rig_split has two)
        vsb_config cfg;
        r = vsb_get_config(s, &cfg);
        for (int k = 0; k < cfg.num_views && r == VSB_OK; ++k) {
            int cam = k;
            r = vsb_view_window(s, k, &cam, nullptr, nullptr);
            if (r == VSB_OK && cam >= 0 && cam < n) r = vsb_set_gain(s, k, (float)bb[cam]);
        }
        if (r != VSB_OK) return r;
    }
    return VSB_OK;
}

}  // extern "C"
