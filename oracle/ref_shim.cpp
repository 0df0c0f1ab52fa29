/*
 * ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (built by oracle/ref.mk into oracle/_ref/libvsref.so).
 *
 * A C-ABI veneer over the reference's own vendored OpenCV 3.4.0 CPU code (compiled in place from
 * /root/reference/sources, nothing copied), used
 *   (1) to pin oracle-G: tests/ compare the restatement with these genuine primitives and with golden
 *       vectors generated from them (tests/golden/make_golden.py), and
 *   (2) as the CPU baseline ("kind": "reference") that bench.py times next to the CUDA path.
 *
 * "oracle-C" below is the CPU compose sequence of the reference: the application's per-view order
 * (360_stitcher/timed.cpp:84-116: remap#1 -> gain convertTo -> remap#2 -> feed) executed with the CPU
 * twins of each call (cv::remap / Mat::convertTo / cv::copyMakeBorder / cv::pyrDown / cv::pyrUp), and the
 * CPU branch of MultiBandBlender (sources/modules/stitching/src/blenders.cpp:463-695 feed, :835-852 blend,
 * :880-941 normalizeUsingWeightMap, :954-1009 createLaplacePyr, :1040-1050 restoreImageFromLaplacePyr),
 * which has to be restated here because the authors' edited blenders.cpp does not compile without the
 * OpenCV CUDA modules.  As in the authors' init_gpu (:344-461) the per-view border geometry
 * and the weight pyramid are computed once, not per frame.
 */
#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>
#include <opencv2/stitching/detail/exposure_compensate.hpp>
#include <opencv2/stitching/detail/seam_finders.hpp>
#include <opencv2/stitching/detail/util.hpp>
#include <opencv2/stitching/detail/warpers.hpp>

/* the reference's own float gold for cuda::remap (LinearInterpolator / readVal), included from where it lies:
 * sources/modules/cudawarping/test/interpolation.hpp:50-84 (ref.mk adds that directory to the include path) */
#include "interpolation.hpp"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace cv;

namespace {

const float kWeightEps = 1e-5f; /* WEIGHT_EPS, blenders.cpp:67 */

int cv_type(int code)
{
    switch (code) {
    case 0: return CV_8UC3;
    case 1: return CV_16SC3;
    case 2: return CV_32FC1;
    case 3: return CV_8UC1;
    default: return -1;
    }
}

struct ViewState {
    int top, bottom, left, right;
    Rect rc0;               /* rect of the bordered view inside the padded dst, level 0 */
    std::vector<Mat> weight; /* Gaussian pyramid of the bordered mask / 255 */
    Point tl;
    Size size;
};

struct BlenderC {
    int want_bands, nb;
    Rect roi_final, roi;
    std::vector<Mat> dst_lap, dst_w;
    std::vector<ViewState> views;
};

void blender_zero(BlenderC *b)
{
    for (int k = 0; k <= b->nb; ++k) { b->dst_lap[k].setTo(Scalar::all(0)); b->dst_w[k].setTo(0); }
}

/* weighted add of one view's Laplacian pyramid (blenders.cpp:646-665) */
void add_weighted(const Mat &lap, const Mat &w, Mat dst, Mat dst_w)
{
    for (int y = 0; y < lap.rows; ++y) {
        const Point3_<short> *s = lap.ptr<Point3_<short> >(y);
        Point3_<short> *d = dst.ptr<Point3_<short> >(y);
        const float *wr = w.ptr<float>(y);
        float *dw = dst_w.ptr<float>(y);
        for (int x = 0; x < lap.cols; ++x) {
            d[x].x += static_cast<short>(s[x].x * wr[x]);
            d[x].y += static_cast<short>(s[x].y * wr[x]);
            d[x].z += static_cast<short>(s[x].z * wr[x]);
            dw[x] += wr[x];
        }
    }
}

/* Laplacian pyramid of a bordered CV_16SC3 view (createLaplacePyr, non-8U branch) */
void laplace_pyr(const Mat &bordered, int nb, std::vector<Mat> &pyr)
{
    pyr.resize(nb + 1);
    pyr[0] = bordered;
    for (int i = 0; i < nb; ++i) pyrDown(pyr[i], pyr[i + 1]);
    Mat tmp;
    for (int i = 0; i < nb; ++i) {
        pyrUp(pyr[i + 1], tmp, pyr[i].size());
        subtract(pyr[i], tmp, pyr[i]);
    }
}

struct RigC {
    int n, src_w, src_h, enable_local;
    std::vector<Mat> xmap, ymap, xmesh, ymesh;
    std::vector<float> gain;
    BlenderC *bl;
};

/* one view of stitch_online on the CPU, up to the weighted add; returns the view's Laplacian pyramid */
void view_front(const RigC *r, int i, const uint8_t *bgr, size_t step, std::vector<Mat> &lap)
{
    const Mat src(r->src_h, r->src_w, CV_8UC3, const_cast<uint8_t *>(bgr), step);
    Mat p, q, s16, bordered;
    remap(src, p, r->xmap[i], r->ymap[i], INTER_LINEAR, BORDER_CONSTANT);
    p.convertTo(p, CV_8U, r->gain[i]);
    if (r->enable_local) remap(p, q, r->xmesh[i], r->ymesh[i], INTER_LINEAR, BORDER_CONSTANT);
    else q = p;
    q.convertTo(s16, CV_16S);
    const ViewState &v = r->bl->views[i];
    copyMakeBorder(s16, bordered, v.top, v.bottom, v.left, v.right, BORDER_REFLECT);
    laplace_pyr(bordered, r->bl->nb, lap);
}

}  // namespace

extern "C" {

const char *vr_build_info(void) { static String s = getBuildInformation(); return s.c_str(); }
void vr_set_num_threads(int n) { setNumThreads(n); }
int vr_get_num_threads(void) { return getNumThreads(); }

/* ---- primitives (tightly packed buffers) */
void vr_pyr_down(const void *src, int w, int h, int type, void *dst)
{
    const int t = cv_type(type);
    Mat s(h, w, t, const_cast<void *>(src)), d((h + 1) / 2, (w + 1) / 2, t, dst), o;
    pyrDown(s, o);
    o.copyTo(d);
}
void vr_pyr_up(const void *src, int w, int h, int type, void *dst)
{
    const int t = cv_type(type);
    Mat s(h, w, t, const_cast<void *>(src)), d(h * 2, w * 2, t, dst), o;
    pyrUp(s, o);
    o.copyTo(d);
}
void vr_remap_u8(const uint8_t *src, int sw, int sh, int cn, const float *xmap, const float *ymap, uint8_t *dst, int dw, int dh, int nearest)
{
    Mat s(sh, sw, CV_8UC(cn), const_cast<uint8_t *>(src)), xm(dh, dw, CV_32F, const_cast<float *>(xmap)),
        ym(dh, dw, CV_32F, const_cast<float *>(ymap)), d(dh, dw, CV_8UC(cn), dst), o;
    remap(s, o, xm, ym, nearest ? INTER_NEAREST : INTER_LINEAR, BORDER_CONSTANT);
    o.copyTo(d);
}
/* remapGold(INTER_LINEAR, BORDER_CONSTANT, 0) of the CUDA remap test: the per-pixel loop of remapImpl<uchar, LinearInterpolator>
 * (sources/modules/cudawarping/test/test_remap.cpp:54-71) over the header's LinearInterpolator<uchar>::getValue */
void vr_remap_gold_u8(const uint8_t *src, int sw, int sh, int cn, const float *xmap, const float *ymap, uint8_t *dst, int dw, int dh)
{
    Mat s(sh, sw, CV_8UC(cn), const_cast<uint8_t *>(src)), xm(dh, dw, CV_32F, const_cast<float *>(xmap)),
        ym(dh, dw, CV_32F, const_cast<float *>(ymap)), d(dh, dw, CV_8UC(cn), dst);
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x)
            for (int c = 0; c < cn; ++c)
                d.at<uchar>(y, x * cn + c) = LinearInterpolator<uchar>::getValue(s, ym.at<float>(y, x), xm.at<float>(y, x), c, BORDER_CONSTANT, Scalar());
}
/* resizeGold(INTER_LINEAR) of the CUDA resize test: resizeImpl<uchar, LinearInterpolator>
 * (sources/modules/cudawarping/test/test_resize.cpp:54-74) -- the gold cuda::resize is tested against with a bound of 1.0 (:152) */
void vr_resize_gold_u8(const uint8_t *src, int sw, int sh, int cn, double fx, double fy, uint8_t *dst, int dw, int dh)
{
    Mat s(sh, sw, CV_8UC(cn), const_cast<uint8_t *>(src)), d(dh, dw, CV_8UC(cn), dst);
    const float ifx = static_cast<float>(1.0 / fx), ify = static_cast<float>(1.0 / fy);
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x)
            for (int c = 0; c < cn; ++c)
                d.at<uchar>(y, x * cn + c) = LinearInterpolator<uchar>::getValue(s, y * ify, x * ifx, c, BORDER_REPLICATE);
}
/* the reference's own GainCompensator::feed + gains() (S/src/exposure_compensate.cpp:71-142,162-168), compiled from its source:
 * imgs[i] CV_8UC3 / masks[i] CV_8U (255 = valid) of sizes_wh[i] at corners_xy[i] */
void vr_gain_compensator_feed(int n, const uint8_t *const *imgs, const uint8_t *const *masks, const int *sizes_wh, const int *corners_xy, double *gains)
{
    std::vector<Point> corners(n);
    std::vector<UMat> images(n);
    std::vector<std::pair<UMat, uchar> > ms(n);
    for (int i = 0; i < n; ++i) {
        corners[i] = Point(corners_xy[2 * i], corners_xy[2 * i + 1]);
        Mat(sizes_wh[2 * i + 1], sizes_wh[2 * i], CV_8UC3, const_cast<uint8_t *>(imgs[i])).copyTo(images[i]);
        Mat(sizes_wh[2 * i + 1], sizes_wh[2 * i], CV_8U, const_cast<uint8_t *>(masks[i])).copyTo(ms[i].first);
        ms[i].second = 255;
    }
    detail::GainCompensator comp;
    comp.feed(corners, images, ms);
    const std::vector<double> g = comp.gains();
    for (int i = 0; i < n; ++i) gains[i] = g[i];
}
void vr_copy_make_border(const void *src, int w, int h, int type, int top, int bottom, int left, int right, int reflect, void *dst)
{
    const int t = cv_type(type);
    Mat s(h, w, t, const_cast<void *>(src)), d(h + top + bottom, w + left + right, t, dst), o;
    copyMakeBorder(s, o, top, bottom, left, right, reflect ? BORDER_REFLECT : BORDER_CONSTANT);
    o.copyTo(d);
}
void vr_gain_u8(uint8_t *buf, size_t n, float gain)
{
    Mat m(1, (int)n, CV_8U, buf);
    m.convertTo(m, CV_8U, gain);
}
void vr_cvt_nv12_bgr(const uint8_t *nv12, int w, int h, uint8_t *bgr)
{
    Mat s(h * 3 / 2, w, CV_8UC1, const_cast<uint8_t *>(nv12)), d(h, w, CV_8UC3, bgr), o;
    cvtColor(s, o, COLOR_YUV2BGR_NV12);  // 360_stitcher/networking.cpp:46
    o.copyTo(d);
}
void vr_convert_s16_u8(const int16_t *src, int w, int h, int cn, uint8_t *dst)
{
    Mat s(h, w, CV_16SC(cn), const_cast<int16_t *>(src)), d(h, w, CV_8UC(cn), dst), o;
    s.convertTo(o, CV_8U);  // 360_stitcher/timed.cpp:250 (CPU twin of GpuMat::convertTo)
    o.copyTo(d);
}
/* the consumer thread after the download, with the reference's own OpenCV calls (360_stitcher/timed.cpp:254-315):
   fmt 0: resize + COLOR_BGR2RGB (:281, :291); fmt 1: resize, BGR2RGB, copy into the black out_w x out_h frame, BGR2RGB back,
   COLOR_BGR2YUV_I420 (:281-289, :310-311).  Returns the image height. */
int vr_consume(const uint8_t *pano, int w, int h, int out_w, int out_h, int keep_aspect, int fmt, uint8_t *out)
{
    Mat original_8u(h, w, CV_8UC3, const_cast<uint8_t *>(pano)), resized_bgr, resized_rgb;
    int image_height = out_h;
    if (keep_aspect) {
        image_height = (double)out_w / (double)original_8u.cols * original_8u.rows + 0.5;
        if (image_height > out_h) image_height = out_h;
    }
    resize(original_8u, resized_bgr, Size(out_w, image_height), 0, 0, INTER_LINEAR);
    if (fmt == 0) {
        Mat final_result(image_height, out_w, CV_8UC3, out), o;
        cvtColor(resized_bgr, o, COLOR_BGR2RGB);
        o.copyTo(final_result);
        return image_height;
    }
    Mat final_result = Mat(Size(out_w, out_h), CV_8UC3, cv::Scalar(0)), final_result_yuv;
    cvtColor(resized_bgr, resized_rgb, COLOR_BGR2RGB);
    uchar *row_ptr = final_result.ptr(final_result.rows / 2 - resized_rgb.rows / 2);
    memcpy(row_ptr, resized_rgb.data, (size_t)resized_rgb.rows * resized_rgb.cols * 3);
    cvtColor(final_result, final_result, COLOR_BGR2RGB);
    cvtColor(final_result, final_result_yuv, COLOR_BGR2YUV_I420);
    memcpy(out, final_result_yuv.data, (size_t)out_w * out_h * 3 / 2);
    return image_height;
}
void vr_resize_linear_u8c1(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh)
{
    Mat s(sh, sw, CV_8U, const_cast<uint8_t *>(src)), d(dh, dw, CV_8U, dst), o;
    resize(s, o, Size(dw, dh), 0, 0, INTER_LINEAR);
    o.copyTo(d);
}
void vr_dilate3x3_u8c1(const uint8_t *src, int w, int h, uint8_t *dst)
{
    Mat s(h, w, CV_8U, const_cast<uint8_t *>(src)), d(h, w, CV_8U, dst), o;
    dilate(s, o, getStructuringElement(MORPH_RECT, Size(3, 3)));
    o.copyTo(d);
}
void vr_distance_l1(const uint8_t *src, int w, int h, float *dst)
{
    Mat s(h, w, CV_8U, const_cast<uint8_t *>(src)), d(h, w, CV_32F, dst), o;
    distanceTransform(s, o, DIST_L1, 3);
    o.copyTo(d);
}

/* ---- warpers: detail::{Spherical,Cylindrical}Warper (CPU), warpers.hpp:249-290, warpers_inl.hpp */
static Ptr<detail::RotationWarper> make_warper(int proj, float scale)
{
    if (proj == 0) return makePtr<detail::SphericalWarper>(scale);
    return makePtr<detail::CylindricalWarper>(scale);
}
int vr_warp_roi(int proj, float scale, const float K[9], const float R[9], int sw, int sh, int roi[4])
{
    Mat k(3, 3, CV_32F, const_cast<float *>(K)), r(3, 3, CV_32F, const_cast<float *>(R));
    const Rect rc = make_warper(proj, scale)->warpRoi(Size(sw, sh), k, r);
    roi[0] = rc.x; roi[1] = rc.y; roi[2] = rc.width; roi[3] = rc.height;
    return 0;
}
/* xmap/ymap sized roi[2] x roi[3] as returned by vr_warp_roi (buildMaps allocates (br - tl + 1)) */
int vr_build_maps(int proj, float scale, const float K[9], const float R[9], int sw, int sh, float *xmap, float *ymap, int roi[4])
{
    Mat k(3, 3, CV_32F, const_cast<float *>(K)), r(3, 3, CV_32F, const_cast<float *>(R)), xm, ym;
    const Rect rc = make_warper(proj, scale)->buildMaps(Size(sw, sh), k, r, xm, ym);
    roi[0] = rc.x; roi[1] = rc.y; roi[2] = xm.cols; roi[3] = xm.rows;
    if (xmap) xm.copyTo(Mat(xm.rows, xm.cols, CV_32F, xmap));
    if (ymap) ym.copyTo(Mat(ym.rows, ym.cols, CV_32F, ymap));
    return 0;
}

/* ---- VoronoiSeamFinder::find(sizes, corners, masks), seam_finders.cpp:72-162; masks updated in place */
void vr_voronoi_find(int n, const int *sizes_wh, const int *corners_xy, uint8_t **masks)
{
    std::vector<Size> sizes(n);
    std::vector<Point> corners(n);
    std::vector<UMat> um(n);
    for (int i = 0; i < n; ++i) {
        sizes[i] = Size(sizes_wh[2 * i], sizes_wh[2 * i + 1]);
        corners[i] = Point(corners_xy[2 * i], corners_xy[2 * i + 1]);
        Mat(sizes[i], CV_8U, masks[i]).copyTo(um[i]);
    }
    detail::VoronoiSeamFinder().find(sizes, corners, um);
    for (int i = 0; i < n; ++i) um[i].getMat(ACCESS_READ).copyTo(Mat(sizes[i], CV_8U, masks[i]));
}
void vr_result_roi(int n, const int *corners_xy, const int *sizes_wh, int roi[4])
{
    std::vector<Size> sizes(n);
    std::vector<Point> corners(n);
    for (int i = 0; i < n; ++i) { sizes[i] = Size(sizes_wh[2 * i], sizes_wh[2 * i + 1]); corners[i] = Point(corners_xy[2 * i], corners_xy[2 * i + 1]); }
    const Rect r = detail::resultRoi(corners, sizes);
    roi[0] = r.x; roi[1] = r.y; roi[2] = r.width; roi[3] = r.height;
}

/* ---- oracle-C blender */
void *vr_blender_create(int num_bands)
{
    BlenderC *b = new BlenderC();
    b->want_bands = num_bands; b->nb = 0;
    return b;
}
void vr_blender_destroy(void *h) { delete static_cast<BlenderC *>(h); }

/* MultiBandBlender::prepare, blenders.cpp:237-295 */
void vr_blender_prepare(void *h, int n, const int *corners_xy, const int *sizes_wh)
{
    BlenderC *b = static_cast<BlenderC *>(h);
    int roi[4];
    vr_result_roi(n, corners_xy, sizes_wh, roi);
    Rect dst(roi[0], roi[1], roi[2], roi[3]);
    b->roi_final = dst;
    const double max_len = static_cast<double>(std::max(dst.width, dst.height));
    b->nb = std::min(b->want_bands, static_cast<int>(ceil(std::log(max_len) / std::log(2.0))));
    const int m = 1 << b->nb;
    dst.width += (m - dst.width % m) % m;
    dst.height += (m - dst.height % m) % m;
    b->roi = dst;
    b->dst_lap.assign(b->nb + 1, Mat());
    b->dst_w.assign(b->nb + 1, Mat());
    b->dst_lap[0].create(dst.size(), CV_16SC3);
    b->dst_w[0].create(dst.size(), CV_32F);
    for (int i = 1; i <= b->nb; ++i) {
        b->dst_lap[i].create((b->dst_lap[i - 1].rows + 1) / 2, (b->dst_lap[i - 1].cols + 1) / 2, CV_16SC3);
        b->dst_w[i].create((b->dst_w[i - 1].rows + 1) / 2, (b->dst_w[i - 1].cols + 1) / 2, CV_32F);
    }
    blender_zero(b);
    b->views.clear();
}
int vr_blender_num_bands(void *h) { return static_cast<BlenderC *>(h)->nb; }
void vr_blender_roi(void *h, int roi_final[4], int roi_padded[4])
{
    BlenderC *b = static_cast<BlenderC *>(h);
    roi_final[0] = b->roi_final.x; roi_final[1] = b->roi_final.y; roi_final[2] = b->roi_final.width; roi_final[3] = b->roi_final.height;
    roi_padded[0] = b->roi.x; roi_padded[1] = b->roi.y; roi_padded[2] = b->roi.width; roi_padded[3] = b->roi.height;
}

/* static half of feed(): geometry (blenders.cpp:476-505) and the weight pyramid (:604-625) */
void vr_blender_add_view(void *h, const uint8_t *mask, int mw, int mh, int tl_x, int tl_y, int out8[8])
{
    BlenderC *b = static_cast<BlenderC *>(h);
    const int nb = b->nb, m = 1 << nb, gap = 3 * m;
    const Rect &roi = b->roi;
    Point tl(tl_x, tl_y);
    Point tl_new(std::max(roi.x, tl.x - gap), std::max(roi.y, tl.y - gap));
    Point br_new(std::min(roi.br().x, tl.x + mw + gap), std::min(roi.br().y, tl.y + mh + gap));
    tl_new.x = roi.x + (((tl_new.x - roi.x) >> nb) << nb);
    tl_new.y = roi.y + (((tl_new.y - roi.y) >> nb) << nb);
    int width = br_new.x - tl_new.x, height = br_new.y - tl_new.y;
    width += (m - width % m) % m;
    height += (m - height % m) % m;
    br_new.x = tl_new.x + width; br_new.y = tl_new.y + height;
    const int dy = std::max(br_new.y - roi.br().y, 0), dx = std::max(br_new.x - roi.br().x, 0);
    tl_new.x -= dx; br_new.x -= dx; tl_new.y -= dy; br_new.y -= dy;
    ViewState v;
    v.top = tl.y - tl_new.y; v.left = tl.x - tl_new.x;
    v.bottom = br_new.y - tl.y - mh; v.right = br_new.x - tl.x - mw;
    v.rc0 = Rect(tl_new.x - roi.x, tl_new.y - roi.y, br_new.x - tl_new.x, br_new.y - tl_new.y);
    v.tl = tl; v.size = Size(mw, mh);
    Mat mk(mh, mw, CV_8U, const_cast<uint8_t *>(mask)), wmap;
    mk.convertTo(wmap, CV_32F, 1. / 255.);
    v.weight.resize(nb + 1);
    copyMakeBorder(wmap, v.weight[0], v.top, v.bottom, v.left, v.right, BORDER_CONSTANT);
    for (int i = 0; i < nb; ++i) pyrDown(v.weight[i], v.weight[i + 1]);
    if (out8) {
        out8[0] = v.top; out8[1] = v.bottom; out8[2] = v.left; out8[3] = v.right;
        out8[4] = v.rc0.x; out8[5] = v.rc0.y; out8[6] = v.rc0.x + v.rc0.width; out8[7] = v.rc0.y + v.rc0.height;
    }
    b->views.push_back(v);
}

/* weight level k of view i (tightly packed copy); returns 0 on success */
int vr_blender_view_weight(void *h, int i, int level, float *out, int *w, int *hgt)
{
    BlenderC *b = static_cast<BlenderC *>(h);
    const Mat &m = b->views[i].weight[level];
    if (w) *w = m.cols;
    if (hgt) *hgt = m.rows;
    if (out) m.copyTo(Mat(m.rows, m.cols, CV_32F, out));
    return 0;
}

static void blender_accumulate(BlenderC *b, int i, const std::vector<Mat> &lap)
{
    const ViewState &v = b->views[i];
    int x_tl = v.rc0.x, y_tl = v.rc0.y, x_br = v.rc0.x + v.rc0.width, y_br = v.rc0.y + v.rc0.height;
    for (int k = 0; k <= b->nb; ++k) {
        const Rect rc(x_tl, y_tl, x_br - x_tl, y_br - y_tl);
        add_weighted(lap[k], v.weight[k], b->dst_lap[k](rc), b->dst_w[k](rc));
        x_tl /= 2; y_tl /= 2; x_br /= 2; y_br /= 2;
    }
}

/* feed(img CV_16SC3 | CV_8UC3 converted to 16S as Stitcher::composePanorama does, stitcher.cpp:357) */
void vr_blender_feed(void *h, int i, const uint8_t *img_u8c3, int w, int hgt)
{
    BlenderC *b = static_cast<BlenderC *>(h);
    const ViewState &v = b->views[i];
    Mat u8(hgt, w, CV_8UC3, const_cast<uint8_t *>(img_u8c3)), s16, bordered;
    u8.convertTo(s16, CV_16S);
    copyMakeBorder(s16, bordered, v.top, v.bottom, v.left, v.right, BORDER_REFLECT);
    std::vector<Mat> lap;
    laplace_pyr(bordered, b->nb, lap);
    blender_accumulate(b, i, lap);
}

/* blend(): normalise, collapse, mask, crop (blenders.cpp:835-852 + Blender::blend :123-131); out CV_16SC3 of roi_final size */
void vr_blender_blend(void *h, int16_t *out, uint8_t *mask_out)
{
    BlenderC *b = static_cast<BlenderC *>(h);
    for (int k = 0; k <= b->nb; ++k) {
        Mat &src = b->dst_lap[k];
        const Mat &wt = b->dst_w[k];
        for (int y = 0; y < src.rows; ++y) {
            Point3_<short> *row = src.ptr<Point3_<short> >(y);
            const float *wr = wt.ptr<float>(y);
            for (int x = 0; x < src.cols; ++x) {
                row[x].x = static_cast<short>(row[x].x / (wr[x] + kWeightEps));
                row[x].y = static_cast<short>(row[x].y / (wr[x] + kWeightEps));
                row[x].z = static_cast<short>(row[x].z / (wr[x] + kWeightEps));
            }
        }
    }
    Mat tmp;
    for (int k = b->nb; k > 0; --k) {
        pyrUp(b->dst_lap[k], tmp, b->dst_lap[k - 1].size());
        add(tmp, b->dst_lap[k - 1], b->dst_lap[k - 1]);
    }
    const Rect rc(0, 0, b->roi_final.width, b->roi_final.height);
    Mat mask;
    compare(b->dst_w[0](rc), kWeightEps, mask, CMP_GT);
    Mat res = b->dst_lap[0](rc).clone();
    Mat inv;
    compare(mask, 0, inv, CMP_EQ);
    res.setTo(Scalar::all(0), inv); /* Blender::blend: dst_.setTo(0, dst_mask_ == 0) */
    res.copyTo(Mat(rc.height, rc.width, CV_16SC3, out));
    if (mask_out) mask.copyTo(Mat(rc.height, rc.width, CV_8U, mask_out));
    blender_zero(b);
}

/* ---- whole-frame CPU compose (the timed CPU baseline) */
void *vr_rig_create(void *blender, int n, int src_w, int src_h, int enable_local)
{
    RigC *r = new RigC();
    r->n = n; r->src_w = src_w; r->src_h = src_h; r->enable_local = enable_local;
    r->xmap.resize(n); r->ymap.resize(n); r->xmesh.resize(n); r->ymesh.resize(n);
    r->gain.assign(n, 1.f);
    r->bl = static_cast<BlenderC *>(blender);
    return r;
}
void vr_rig_destroy(void *h) { delete static_cast<RigC *>(h); }
void vr_rig_set_view(void *h, int i, const float *xmap, const float *ymap, int w, int hgt, float gain)
{
    RigC *r = static_cast<RigC *>(h);
    Mat(hgt, w, CV_32F, const_cast<float *>(xmap)).copyTo(r->xmap[i]);
    Mat(hgt, w, CV_32F, const_cast<float *>(ymap)).copyTo(r->ymap[i]);
    r->gain[i] = gain;
}
void vr_rig_set_mesh_maps(void *h, int i, const float *xmesh, const float *ymesh, int w, int hgt)
{
    RigC *r = static_cast<RigC *>(h);
    Mat(hgt, w, CV_32F, const_cast<float *>(xmesh)).copyTo(r->xmesh[i]);
    Mat(hgt, w, CV_32F, const_cast<float *>(ymesh)).copyTo(r->ymesh[i]);
}
/* frames[i] = BGR u8 of view i (pitch src_pitch).  parallel_views = 0: views in sequence as stitch_one does
 * (timed.cpp:127-132), OpenCV's own parallel_for_ threads inside each call; 1: one OpenMP thread per view for
 * the front half (the way the reference parallelises its calibration loop, calibration.cpp:91), accumulation
 * into the shared destination pyramid kept in view order. */
void vr_rig_compose(void *h, const uint8_t *const *frames, size_t src_pitch, int16_t *out, uint8_t *mask_out, int parallel_views)
{
    RigC *r = static_cast<RigC *>(h);
    if (!parallel_views) {
        for (int i = 0; i < r->n; ++i) {
            std::vector<Mat> lap;
            view_front(r, i, frames[i], src_pitch, lap);
            blender_accumulate(r->bl, i, lap);
        }
    } else {
        std::vector<std::vector<Mat> > laps(r->n);
#pragma omp parallel for schedule(dynamic, 1)
        for (int i = 0; i < r->n; ++i) view_front(r, i, frames[i], src_pitch, laps[i]);
        for (int i = 0; i < r->n; ++i) blender_accumulate(r->bl, i, laps[i]);
    }
    vr_blender_blend(r->bl, out, mask_out);
}

}  /* extern "C" */
