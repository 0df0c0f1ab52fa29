/*
 * vsb200.hpp -- C++ adapter over the C ABI (vsb200.h) that mirrors the reference's own class / function names and call
 * order for the per-frame compose path, so host code shaped like 360_stitcher/timed.cpp + calibration.cpp keeps its
 * structure (the north star: "host code stays C++").  Header-only; needs nothing but libvsb200.so.
 *
 * Reference interfaces mirrored (paths relative to the reference root):
 *   cv::detail::MultiBandBlender   sources/modules/stitching/include/opencv2/stitching/detail/blenders.hpp:126-176
 *       MultiBandBlender(try_gpu, num_bands, weight_type) / setNumBands / prepare / init_gpu / feed_online / blend
 *   stitch_online / stitch_one     360_stitcher/timed.cpp:56-152
 *   custom_resize                  360_stitcher/resize.cu:30-45 (decl 360_stitcher/calibration.h:15)
 *   MeshWarper::convertMeshesToMap 360_stitcher/meshwarper.h:34-36, meshwarper.cpp:823-886
 *   {Spherical,Cylindrical}WarperGpu::warpRoi / buildMaps   .../detail/warpers.hpp:489-550
 *
 * Types: the reference passes cv::cuda::GpuMat / cv::cuda::Stream.  This header does not depend on OpenCV: DeviceMat is
 * the (data, step, rows, cols) quadruple of a GpuMat / PtrStepSz and Stream is a cudaStream_t.  When OpenCV's core/cuda.hpp
 * has been included first, from_gpumat() / from_stream() convert without copying.
 *
 * Errors: the reference throws cv::Exception from CV_Assert / cudaSafeCall (core/cuda/common.hpp:66-74); here every
 * non-zero status of the C ABI is re-thrown as vsb::Error (std::runtime_error) carrying vsb_last_error().
 */
#ifndef VSB200_HPP
#define VSB200_HPP

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "vsb200.h"

namespace vsb {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};
inline void check(int rc)
{
    if (rc != VSB_OK) throw Error(rc, std::string("vsb200: ") + vsb_last_error());
}

typedef void *Stream;  /* cudaStream_t */

struct Point { int x, y; Point(int x_ = 0, int y_ = 0) : x(x_), y(y_) {} };
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect(int x_ = 0, int y_ = 0, int w = 0, int h = 0) : x(x_), y(y_), width(w), height(h) {} };

/* a pitched device matrix the caller owns (what a cv::cuda::GpuMat header describes) */
struct DeviceMat {
    void *data;
    size_t step;  /* bytes */
    int rows, cols;
    DeviceMat() : data(0), step(0), rows(0), cols(0) {}
    DeviceMat(void *d, size_t s, int r, int c) : data(d), step(s), rows(r), cols(c) {}
};
#ifdef OPENCV_CORE_CUDA_HPP
inline DeviceMat from_gpumat(const cv::cuda::GpuMat &m) { return DeviceMat(m.data, m.step, m.rows, m.cols); }
#endif

/* ---- rotation warpers: warpRoi / buildMaps (static inputs of remap #1) -------------------------------------------- */
class RotationWarperGpu {
public:
    RotationWarperGpu(int projection, float scale) : projection_(projection), scale_(scale) {}
    /* Rect warpRoi(Size src_size, InputArray K, InputArray R): K, R = 3x3 CV_32F, row major */
    Rect warpRoi(Size src, const float K[9], const float R[9]) const
    {
        int roi[4];
        check(vsb_warp_roi(projection_, scale_, K, R, src.width, src.height, roi));
        return Rect(roi[0], roi[1], roi[2], roi[3]);
    }
    /* Rect buildMaps(Size, K, R, GpuMat &xmap, GpuMat &ymap): the maps must be roi-sized CV_32FC1 device matrices */
    Rect buildMaps(Size src, const float K[9], const float R[9], DeviceMat xmap, DeviceMat ymap, Stream stream = 0) const
    {
        int roi[4];
        if (xmap.step != ymap.step) throw Error(VSB_ERR_INVALID, "vsb200: buildMaps: xmap and ymap must share one pitch");
        check(vsb_build_maps(projection_, scale_, K, R, src.width, src.height, (float *)xmap.data, (float *)ymap.data, xmap.step, roi, stream));
        return Rect(roi[0], roi[1], roi[2], roi[3]);
    }
    /* Point warp(const GpuMat &src, K, R, int interp_mode, int border_mode, GpuMat &dst) (S/src/warpers_cuda.cpp:279-298):
     * dst must be a warpRoi-sized matrix of src's type (CV_8UC1 / CV_8UC3); returns the roi's top-left corner */
    Point warp(const DeviceMat &src, int channels, const float K[9], const float R[9], int interp_mode, int border_mode, DeviceMat &dst, Stream stream = 0) const
    {
        int roi[4];
        check(vsb_warp(projection_, scale_, K, R, (const uint8_t *)src.data, src.cols, src.rows, src.step, channels, interp_mode, border_mode,
                       (uint8_t *)dst.data, dst.step, roi, stream));
        return Point(roi[0], roi[1]);
    }
private:
    int projection_;
    float scale_;
};
struct SphericalWarperGpu : RotationWarperGpu { explicit SphericalWarperGpu(float scale) : RotationWarperGpu(VSB_PROJ_SPHERICAL, scale) {} };
struct CylindricalWarperGpu : RotationWarperGpu { explicit CylindricalWarperGpu(float scale) : RotationWarperGpu(VSB_PROJ_CYLINDRICAL, scale) {} };

/* ---- extern void custom_resize(GpuMat &in, GpuMat &out, Size t_size), 360_stitcher/resize.cu:30 --------------------
 * `out` must already be a t_size CV_32FC1 device matrix (the reference creates it); asynchronous on `stream` instead of
 * the reference's cudaDeviceSynchronize(). */
inline void custom_resize(const DeviceMat &in, DeviceMat &out, Size t_size, Stream stream = 0)
{
    if (out.cols != t_size.width || out.rows != t_size.height) throw Error(VSB_ERR_INVALID, "vsb200: custom_resize: out must be t_size");
    check(vsb_custom_resize((const float *)in.data, in.cols, in.rows, in.step, (float *)out.data, t_size.width, t_size.height, out.step, stream));
}

/* ---- MultiBandBlender (the authors' GPU variant) + the stitch_online / stitch_one drivers ------------------------- */
class MultiBandBlender {
public:
    /* MultiBandBlender(int try_gpu = false, int num_bands = 5, int weight_type = CV_32F): only the CV_32F weights the
     * application uses exist here; try_gpu is implied.  num_views replaces the reference's hard-coded vector<...>(6). */
    explicit MultiBandBlender(int num_views, int num_bands = 5, bool enable_local = true, int max_batch = 1, int device = -1)
        : h_(0), n_(num_views), next_view_(0)
    {
        cfg_.num_views = num_views; cfg_.num_bands = num_bands; cfg_.enable_local = enable_local ? 1 : 0;
        cfg_.max_batch = max_batch; cfg_.device = device;
        check(vsb_create(&cfg_, &h_));
    }
    ~MultiBandBlender() { vsb_destroy(h_); }

    int numBands() const { return cfg_.num_bands; }
    /* setNumBands must precede prepare(), as in the reference (calibration.cpp:193-196) */
    void setNumBands(int val)
    {
        cfg_.num_bands = val;
        vsb_destroy(h_); h_ = 0; next_view_ = 0;
        check(vsb_create(&cfg_, &h_));
    }
    /* Blender::prepare(const std::vector<Point> &corners, const std::vector<Size> &sizes) */
    void prepare(const std::vector<Point> &corners, const std::vector<Size> &sizes)
    {
        if ((int)corners.size() != n_ || (int)sizes.size() != n_) throw Error(VSB_ERR_INVALID, "vsb200: prepare: one corner and one size per view");
        std::vector<int> c(2 * n_), s(2 * n_);
        for (int i = 0; i < n_; ++i) { c[2 * i] = corners[i].x; c[2 * i + 1] = corners[i].y; s[2 * i] = sizes[i].width; s[2 * i + 1] = sizes[i].height; }
        check(vsb_prepare(h_, c.data(), s.data()));
        next_view_ = 0;
    }
    /* void init_gpu(GpuMat img (unused by the reference), GpuMat mask, Point tl): views in push_back order */
    void init_gpu(const DeviceMat &mask, Point tl)
    {
        check(vsb_init_view(h_, next_view_, (const uint8_t *)mask.data, mask.cols, mask.rows, mask.step, tl.x, tl.y, 1));
        ++next_view_;
    }
    void init_host_mask(const uint8_t *mask, int w, int h, size_t step, Point tl)
    {
        check(vsb_init_view(h_, next_view_, mask, w, h, step, tl.x, tl.y, 0));
        ++next_view_;
    }
    /* static inputs of stitch_online: x_maps[i] / y_maps[i] (calibration.cpp:221), gains (timed.cpp:94) */
    void setMaps(int i, const DeviceMat &xmap, const DeviceMat &ymap, Size src)
    {
        if (xmap.step != ymap.step) throw Error(VSB_ERR_INVALID, "vsb200: setMaps: xmap and ymap must share one pitch");
        check(vsb_set_maps(h_, i, (const float *)xmap.data, (const float *)ymap.data, xmap.cols, xmap.rows, xmap.step, 1, src.width, src.height));
    }
    void setGain(int i, double gain) { check(vsb_set_gain(h_, i, (float)gain)); }
    /* compose_scale of stitch_online (A/timed.cpp:56,74-81): with a scale other than 1 the maps address the resized frame and
       stitch_online / stitch resize the full_size frames on the device first (cuda::resize, INTER_LINEAR); call after setMaps */
    void setComposeScale(double compose_scale, Size full_size) { check(vsb_set_compose_scale(h_, compose_scale, full_size.width, full_size.height)); }
    /* wire format in / consumer format out: VSB_IN_NV12 = the capture boards' NV12 frames (cv::cvtColor(CV_YUV2BGR_NV12),
       A/networking.cpp:46, runs on the device); VSB_OUT_U8C3 = the consumer's mat.convertTo(mat_8u, CV_8U) (A/timed.cpp:250)
       fused into blend(): gpuOut is then CV_8UC3 */
    void setFormats(int input_format, int output_format) { check(vsb_set_formats(h_, input_format, output_format)); }

    /* void feed_online(cuda::GpuMat &gpu_img, int img_num, cuda::Stream &stream): gpu_img = warped CV_8UC3 view */
    void feed_online(const DeviceMat &gpu_img, int img_num, Stream stream)
    {
        check(vsb_feed_warped(h_, img_num, (const uint8_t *)gpu_img.data, gpu_img.step, stream));
    }
    /* stitch_online(compose_scale, img, x_map, y_map, x_mesh, y_mesh, ..., mb, gc, thread_num) minus the H2D upload:
     * remap #1 -> gain -> remap #2 -> feed_online in one call, from the camera frame already on the device */
    void stitch_online(const DeviceMat &bgr_frame, int img_num, Stream stream)
    {
        check(vsb_feed(h_, img_num, (const uint8_t *)bgr_frame.data, bgr_frame.step, stream));
    }
    /* void blend(InputOutputArray dst, InputOutputArray dst_mask, cuda::GpuMat &gpuOut, bool outputGpu = true): gpuOut is
     * CV_16SC3 of resultRoi() size, CALLER-owned here (the reference allocates it and moves it into its result queue) */
    void blend(DeviceMat &gpuOut, Stream stream)
    {
        check(vsb_blend(h_, (int16_t *)gpuOut.data, gpuOut.step, stream));
    }
    /* stitch_one for a batch of frames: srcs[f * num_views + i] */
    void stitch(const std::vector<DeviceMat> &srcs, std::vector<DeviceMat> &outs, Stream stream)
    {
        if (outs.empty() || srcs.size() != outs.size() * (size_t)n_) throw Error(VSB_ERR_INVALID, "vsb200: stitch: need num_views sources per output");
        std::vector<const uint8_t *> sp(srcs.size());
        std::vector<int16_t *> op(outs.size());
        for (size_t k = 0; k < srcs.size(); ++k) { sp[k] = (const uint8_t *)srcs[k].data; if (srcs[k].step != srcs[0].step) throw Error(VSB_ERR_INVALID, "vsb200: stitch: sources must share one pitch"); }
        for (size_t k = 0; k < outs.size(); ++k) { op[k] = (int16_t *)outs[k].data; if (outs[k].step != outs[0].step) throw Error(VSB_ERR_INVALID, "vsb200: stitch: outputs must share one pitch"); }
        check(vsb_compose(h_, (int)outs.size(), sp.data(), srcs[0].step, op.data(), outs[0].step, stream));
    }
    /* dst_roi_final_ (the Rect `blend` crops to) and the padded dst_roi_ */
    /* the consumer thread after blend (360_stitcher/timed.cpp:254-315), on the device: resize(INTER_LINEAR) of the CV_8UC3
       panorama (setFormats(..., VSB_OUT_U8C3)) to out.width x consumerImageHeight, then VSB_CONSUME_RGB (cvtColor BGR2RGB) or
       VSB_CONSUME_I420 (black bars to out.height + cvtColor BGR2YUV_I420: the frame the encoder is fed) */
    int consumerImageHeight(Size out, bool keep_aspect_ratio = true) const
    {
        const Rect r = resultRoi();
        return vsb_consumer_image_height(r.width, r.height, out.width, out.height, keep_aspect_ratio ? 1 : 0);
    }
    void consume(const DeviceMat &pano_8u, Size out, bool keep_aspect_ratio, int format, void *d_out, size_t out_pitch, Stream stream)
    {
        check(vsb_consume(h_, (const uint8_t *)pano_8u.data, pano_8u.step, out.width, out.height, keep_aspect_ratio ? 1 : 0, format,
                          (uint8_t *)d_out, out_pitch, stream));
    }
    Rect resultRoi() const { int a[4], b[4], nb; check(vsb_get_roi(h_, a, b, &nb)); return Rect(a[0], a[1], a[2], a[3]); }
    Rect paddedRoi() const { int a[4], b[4], nb; check(vsb_get_roi(h_, a, b, &nb)); return Rect(b[0], b[1], b[2], b[3]); }

    /* calibrateCameras + warpImages for the fixed rig (calibration.cpp:28-249), N-generic */
    void calibrateRig(int projection, int pano_width, Size src, double hfov_deg = 90.0, const float *gains = 0)
    {
        check(vsb_calibrate_rig(h_, projection, pano_width, src.width, src.height, hfov_deg, gains));
        next_view_ = n_;
    }
    /* the same with every per-pixel loop on the device (maps, mask warps, VoronoiSeamFinder, dilate / resize / and, weight pyramids);
       keeps the seam-scale state estimateGains() needs */
    void calibrateRigDevice(int projection, int pano_width, Size src, double hfov_deg = 90.0, const float *gains = 0)
    {
        check(vsb_calibrate_rig_device(h_, projection, pano_width, src.width, src.height, hfov_deg, gains));
        next_view_ = n_;
    }
    /* either calibration at the reference's compose_scale (calibration.cpp:137-205): scaled cameras and warper, the reference's
       cvRound / (int) sizes; stitch_online / stitch then take FULL-size frames and resize them on the device first (timed.cpp:74-81) */
    void calibrateRigScaled(int projection, int pano_width, Size src, double compose_scale, bool on_device = false, double hfov_deg = 90.0,
                            const float *gains = 0)
    {
        check(vsb_calibrate_rig_scaled(h_, projection, pano_width, src.width, src.height, hfov_deg, gains, compose_scale, on_device ? 1 : 0));
        next_view_ = n_;
    }
    /* stitch_calib with its own WORK_MEGAPIX / COMPOSE_MEGAPIX constants (defs.h:51-53): the reference's default panorama geometry */
    void calibrateRigMegapix(int projection, Size src, double work_megapix = 0.6, double compose_megapix = 1.4, bool on_device = false,
                             double hfov_deg = 90.0, const float *gains = 0)
    {
        check(vsb_calibrate_rig_megapix(h_, projection, src.width, src.height, hfov_deg, gains, work_megapix, compose_megapix, on_device ? 1 : 0));
        next_view_ = n_;
    }
    /* wrapAround (defs.h:25) without a panorama-wide ROI: how many views a rig of n_cameras needs when every camera that looks across
       +-pi is installed as two column windows of its warped image (construct the blender with that many views), and the calibration
       that does it (gains per camera).  viewCamera(v) = the camera whose frame -- and whose mesh -- view v takes. */
    static int splitViewCount(int projection, int pano_width, int n_cameras, Size src, int num_bands, double hfov_deg = 90.0)
    {
        int n = 0;
        check(vsb_split_plan(projection, pano_width, n_cameras, src.width, src.height, hfov_deg, num_bands, &n, 0, 0, 0));
        return n;
    }
    void calibrateRigSplit(int projection, int pano_width, int n_cameras, Size src, bool on_device = false, double hfov_deg = 90.0, const float *gains = 0)
    {
        check(vsb_calibrate_rig_split(h_, projection, pano_width, n_cameras, src.width, src.height, hfov_deg, gains, on_device ? 1 : 0));
        next_view_ = n_;
    }
    int viewCamera(int view) const { int c = 0; check(vsb_view_window(h_, view, &c, 0, 0)); return c; }
    /* GainCompensator::feed + gains() (S/src/exposure_compensate.cpp:71-142,162-168) from the current camera frames, at run time;
       apply = true installs them (what A/timed.cpp:94 multiplies by) */
    std::vector<float> estimateGains(const std::vector<DeviceMat> &frames, bool apply, Stream stream = 0)
    {
        /* one frame per CAMERA (= per view, unless calibrateRigSplit made two views of a camera) */
        int n_cam = 0;
        for (int v = 0; v < n_; ++v) n_cam = viewCamera(v) + 1 > n_cam ? viewCamera(v) + 1 : n_cam;
        if ((int)frames.size() != n_cam) throw Error(VSB_ERR_INVALID, "vsb200: estimateGains: one frame per camera");
        std::vector<const uint8_t *> fp(n_cam);
        for (int i = 0; i < n_cam; ++i) { fp[i] = (const uint8_t *)frames[i].data; if (frames[i].step != frames[0].step) throw Error(VSB_ERR_INVALID, "vsb200: estimateGains: frames must share one pitch"); }
        std::vector<float> g(n_cam);
        check(vsb_estimate_gains(h_, fp.data(), frames[0].step, g.data(), apply ? 1 : 0, stream));
        return g;
    }
    /* host buffers in / out, asynchronous: submit returns once the copies and kernels are enqueued (pinned buffers), wait blocks until
       the OLDEST outstanding submission's panoramas are in host memory -- the role of BlockingQueue<GpuMat> results (A/timed.cpp:150,243) */
    void submitHost(const std::vector<const uint8_t *> &h_srcs, size_t src_pitch, const std::vector<int16_t *> &h_outs, size_t out_pitch)
    {
        if (h_outs.empty() || h_srcs.size() != h_outs.size() * (size_t)n_) throw Error(VSB_ERR_INVALID, "vsb200: submitHost: need num_views sources per output");
        check(vsb_submit_host(h_, (int)h_outs.size(), h_srcs.data(), src_pitch, h_outs.data(), out_pitch));
    }
    void waitHost() { check(vsb_wait_host(h_)); }
    /* one frame stream on several GPUs (one process and one calibrated blender per GPU): id = the bytes vsb_shard_unique_id() returned
       on rank 0, handed to every rank by the host's control plane; stitchSharded composes this rank's strip of every output */
    void shardInit(int rank, int world, const void *id128) { check(vsb_shard_init(h_, rank, world, id128)); }
    void stitchSharded(const std::vector<DeviceMat> &srcs /* [f * num_views + i]; data = 0 for views of other ranks */, std::vector<DeviceMat> &outs, Stream stream)
    {
        if (outs.empty() || srcs.size() != outs.size() * (size_t)n_) throw Error(VSB_ERR_INVALID, "vsb200: stitchSharded: need num_views sources per output");
        std::vector<const uint8_t *> sp(srcs.size());
        std::vector<int16_t *> op(outs.size());
        size_t pitch = 0;
        for (size_t k = 0; k < srcs.size(); ++k) { sp[k] = (const uint8_t *)srcs[k].data; if (srcs[k].data) pitch = srcs[k].step; }
        for (size_t k = 0; k < outs.size(); ++k) op[k] = (int16_t *)outs[k].data;
        check(vsb_shard_compose(h_, (int)outs.size(), sp.data(), pitch, op.data(), outs[0].step, stream));
    }
    Size viewSize(int i) const { vsb_rig_info info; check(vsb_rig_info_get(h_, &info)); return Size(info.view_roi[i][2], info.view_roi[i][3]); }

    vsb_stitcher *handle() { return h_; }
private:
    MultiBandBlender(const MultiBandBlender &);
    MultiBandBlender &operator=(const MultiBandBlender &);
    vsb_stitcher *h_;
    vsb_config cfg_;
    int n_, next_view_;
};

/* ---- MeshWarper::convertMeshesToMap(mesh_x, mesh_y, map_x, map_y, mesh_sizes), 360_stitcher/meshwarper.cpp:823-886 ---
 * mesh_x[i] / mesh_y[i]: N x M vertex positions (host, row major).  The maps live inside the blender handle (double
 * buffered; the reference rewrites the live maps under LockableVector mutexes): callable from the recalibration thread
 * while another thread composes. */
struct MeshCpu { const float *x, *y; int rows, cols; };
inline void convertMeshesToMap(MultiBandBlender &mb, const std::vector<MeshCpu> &meshes)
{
    for (size_t i = 0; i < meshes.size(); ++i) check(vsb_set_mesh(mb.handle(), (int)i, meshes[i].x, meshes[i].y, meshes[i].rows, meshes[i].cols));
}

}  /* namespace vsb */
#endif /* VSB200_HPP */
