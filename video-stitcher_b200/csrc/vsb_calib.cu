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
// (sources/modules/stitching/src/cuda/build_warp_maps.cu:88-134)
static void host_build_maps(int proj, float scale, const float K[9], const float R[9], int tl_x, int tl_y, int w, int h,
                            float *xmap, float *ymap)
{
    float k[9], r_kinv[9], rinv[9];
    projector_setup(K, R, k, r_kinv, rinv);
    for (int dv = 0; dv < h; ++dv)
        for (int du = 0; du < w; ++du) {
            float u = (float)(tl_x + du), v = (float)(tl_y + dv);
            float x_, y_, z_;
            if (proj == VSB_PROJ_SPHERICAL) {
                v = v / scale; u = u / scale;
                const float sinv = sinf(v);
                x_ = sinv * sinf(u);
                y_ = -cosf(v);
                z_ = sinv * cosf(u);
            } else {
                u = u / scale;
                x_ = sinf(u);
                y_ = v / scale;
                z_ = cosf(u);
            }
            float x = std::fmaf(k[2], z_, std::fmaf(k[1], y_, k[0] * x_));
            float y = std::fmaf(k[5], z_, std::fmaf(k[4], y_, k[3] * x_));
            const float z = std::fmaf(k[8], z_, std::fmaf(k[7], y_, k[6] * x_));
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

int vsb_voronoi_seams(int n, const int *sizes_wh, const int *corners_xy, uint8_t *const *masks)
{
    if (n < 1 || !sizes_wh || !corners_xy || !masks) return vsb::fail(VSB_ERR_INVALID, "voronoi_seams: bad arguments");
    vsb::voronoi(n, sizes_wh, corners_xy, masks);
    return VSB_OK;
}

int vsb_calibrate_rig(vsb_stitcher *s, int projection, int pano_width, int src_w, int src_h, double hfov_deg, const float *gains)
{
    using namespace vsb;
    if (!s || pano_width <= 0 || src_w <= 0 || src_h <= 0) return fail(VSB_ERR_INVALID, "calibrate_rig: bad arguments");
    vsb_config cfg;
    int r = vsb_get_config(s, &cfg);
    if (r != VSB_OK) return r;
    const int n = cfg.num_views;
    const float scale = (float)(pano_width / (2.0 * 3.1415926535897932384626));  // sphere radius: pano_width px per 2*pi
    std::vector<float> K(9 * n), R(9 * n);
    for (int i = 0; i < n; ++i) {
        r = vsb_rig_camera(n, i, src_w, src_h, hfov_deg, &K[9 * i], &R[9 * i]);
        if (r != VSB_OK) return r;
    }
    // ---- seam scale (360_stitcher/calibration.cpp:92-135; SEAM_MEAGPIX = 0.01, 360_stitcher/defs.h:52)
    const double seam_scale = std::min(1.0, std::sqrt(0.01 * 1e6 / ((double)src_w * src_h)));
    const int seam_w = (int)std::nearbyint(src_w * seam_scale), seam_h = (int)std::nearbyint(src_h * seam_scale);
    const float seam_warp_scale = static_cast<float>(scale * seam_scale);
    const float swa = (float)seam_scale;
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
    // ---- compose scale (360_stitcher/calibration.cpp:137-246), compose_scale = 1
    std::vector<int> corners(2 * n), sizes(2 * n);
    for (int i = 0; i < n; ++i) {
        int roi[4];
        r = vsb_warp_roi(projection, scale, &K[9 * i], &R[9 * i], src_w, src_h, roi);
        if (r != VSB_OK) return r;
        corners[2 * i] = roi[0]; corners[2 * i + 1] = roi[1]; sizes[2 * i] = roi[2]; sizes[2 * i + 1] = roi[3];
    }
    r = vsb_prepare(s, corners.data(), sizes.data());
    if (r != VSB_OK) return r;
    for (int i = 0; i < n; ++i) {
        const int w = sizes[2 * i], h = sizes[2 * i + 1];
        std::vector<float> xm((size_t)w * h), ym((size_t)w * h);
        host_build_maps(projection, scale, &K[9 * i], &R[9 * i], corners[2 * i], corners[2 * i + 1], w, h, xm.data(), ym.data());
        std::vector<uint8_t> warped((size_t)w * h), seam((size_t)w * h);
        host_warp_full_mask(xm.data(), ym.data(), w, h, src_w, src_h, warped.data());
        const int sw = seam_sizes[2 * i], sh = seam_sizes[2 * i + 1];
        std::vector<uint8_t> dil(seam_masks[i].size());
        if (cfg.enable_local) dilate3x3(seam_masks[i].data(), sw, sh, dil.data());
        else dil = seam_masks[i];
        resize_linear_u8(dil.data(), sw, sh, seam.data(), w, h);
        for (size_t j = 0; j < seam.size(); ++j) seam[j] &= warped[j];
        r = vsb_init_view(s, i, seam.data(), w, h, (size_t)w, corners[2 * i], corners[2 * i + 1], 0);
        if (r != VSB_OK) return r;
        r = vsb_set_maps(s, i, xm.data(), ym.data(), w, h, (size_t)w * 4, 0, src_w, src_h);
        if (r != VSB_OK) return r;
        if (gains) { r = vsb_set_gain(s, i, gains[i]); if (r != VSB_OK) return r; }
    }
    return vsb_note_rig(s, projection, scale, src_w, src_h);
}

}  // extern "C"
