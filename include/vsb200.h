/*
 * vsb200.h -- C ABI of the B200-native 360-degree compose path (libvsb200.so).
 *
 * Drop-in boundary for the per-frame compose path of ultravideo/video-stitcher.  The reference has
 * no FFI: its boundary is a set of C++ call sites.  Each entry point below names the
 * reference interface it stands behind (paths relative to the reference root; A/ = 360_stitcher/,
 * S/ = sources/modules/stitching/, CW/ = sources/modules/cudawarping/, CA/ = .../cudaarithm/).
 *
 * Conventions: plain C types only; every call returns an int status (VSB_OK = 0, negative = error,
 * text via vsb_last_error()); no exceptions cross the ABI; device buffers are caller-owned or
 * handle-owned, never returned; every per-frame call takes an explicit cudaStream_t (as void*) and is
 * asynchronous on it; one handle per GPU; vsb_set_mesh may be called from a second host thread while
 * another thread runs vsb_feed / vsb_blend / vsb_compose.
 *
 * There is NO CPU fallback: every compute entry point fails with VSB_ERR_CUDA when no device is present.
 */
#ifndef VSB200_H
#define VSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSB_OK 0
#define VSB_ERR_INVALID (-1) /* bad argument / call order (the reference would CV_Assert) */
#define VSB_ERR_CUDA (-2)    /* CUDA runtime error (the reference would throw via cudaSafeCall) */
#define VSB_ERR_NOMEM (-3)
#define VSB_ERR_STATE (-4)   /* handle not fully calibrated for this call */

#define VSB_MAX_VIEWS 16
#define VSB_MAX_BANDS 7
#define VSB_MAX_STAGES 16

enum { VSB_PROJ_SPHERICAL = 0, VSB_PROJ_CYLINDRICAL = 1 };

typedef struct vsb_stitcher vsb_stitcher;

typedef struct vsb_config {
    int num_views;     /* NUM_IMAGES, A/defs.h:37 (compile-time 6 in the reference; any 1..VSB_MAX_VIEWS here) */
    int num_bands;     /* MultiBandBlender::setNumBands, S/include/opencv2/stitching/detail/blenders.hpp:132 */
    int enable_local;  /* enable_local, A/defs.h:27 : CPW-mesh remap#2 on/off */
    int max_batch;     /* frames per vsb_compose submission (>=1); the reference is 1 frame at a time */
    int device;        /* CUDA device ordinal; -1 = current device (reference: cuda::setDevice(0), A/timed.cpp:496) */
} vsb_config;

const char *vsb_last_error(void);
const char *vsb_version(void);
int vsb_device_count(void);

/* ---- lifetime ------------------------------------------------------------------------------------ */
int vsb_create(const vsb_config *cfg, vsb_stitcher **out);
int vsb_destroy(vsb_stitcher *s);

/* ---- B3: projection maps.  {Spherical,Cylindrical}WarperGpu::warpRoi / buildMaps,
 *      S/src/warpers_cuda.cpp:210-231,255-277; kernel S/src/cuda/build_warp_maps.cu:88-152 ----------- */
/* roi = {tl_x, tl_y, width, height}; width/height are the size of the maps buildMaps produces */
int vsb_warp_roi(int projection, float scale, const float K[9], const float R[9], int src_w, int src_h, int roi[4]);
/* d_xmap / d_ymap: device, roi[2] x roi[3] floats, pitch in bytes */
int vsb_build_maps(int projection, float scale, const float K[9], const float R[9], int src_w, int src_h,
                   float *d_xmap, float *d_ymap, size_t pitch_bytes, int roi[4], void *stream);

/* RotationWarperGpu::warp(src, K, R, interp_mode, border_mode, dst) (S/src/warpers_cuda.cpp:279-298; used by the
 * application at A/calibration.cpp:118,122,227): buildMaps into scratch + cuda::remap.  d_src: CV_8UC1 / CV_8UC3 (channels = 1 / 3);
 * d_dst: roi[2] x roi[3] pixels of the same type (size it with vsb_warp_roi first); returns the roi whose tl is warp()'s Point. */
enum { VSB_INTER_NEAREST = 0, VSB_INTER_LINEAR = 1 };      /* cv::INTER_NEAREST, cv::INTER_LINEAR */
enum { VSB_BORDER_CONSTANT = 0, VSB_BORDER_REFLECT = 2 };   /* cv::BORDER_CONSTANT (value 0), cv::BORDER_REFLECT */
int vsb_warp(int projection, float scale, const float K[9], const float R[9], const uint8_t *d_src, int src_w, int src_h,
             size_t src_pitch, int channels, int interp, int border, uint8_t *d_dst, size_t dst_pitch, int roi[4], void *stream);

/* ---- B6: static setup.  MultiBandBlender::prepare / init_gpu (S/src/blenders.cpp:237-295,344-461),
 *      x_maps/y_maps (A/calibration.cpp:221), GainCompensator::gains (A/timed.cpp:94) ----------------- */
int vsb_prepare(vsb_stitcher *s, const int *corners_xy, const int *sizes_wh);
int vsb_get_roi(const vsb_stitcher *s, int roi_final[4], int roi_padded[4], int *num_bands);
/* views must be initialised in order i = 0..num_views-1 exactly once after vsb_prepare (the reference push_backs).
 * mask: CV_8U seam mask of the warped view (host or device memory), tl = corner of the warped view. */
int vsb_init_view(vsb_stitcher *s, int i, const uint8_t *mask, int w, int h, size_t pitch_bytes,
                  int tl_x, int tl_y, int mask_on_device);
/* out8 = {top, bottom, left, right, x_tl, y_tl, x_br, y_br} (S/src/blenders.cpp:383-434) */
int vsb_get_view_geometry(const vsb_stitcher *s, int i, int out8[8]);
/* projection maps of remap #1 for view i (size must equal the view's mask size); src_w/src_h = camera frame size */
int vsb_set_maps(vsb_stitcher *s, int i, const float *xmap, const float *ymap, int w, int h, size_t pitch_bytes,
                 int maps_on_device, int src_w, int src_h);
int vsb_set_gain(vsb_stitcher *s, int i, float gain);

/* ---- fixed-rig calibration (the static half of the path): calibrateCameras + warpImages,
 *      360_stitcher/calibration.cpp:28-249, generalised to N views (yaw_i = 2*pi*i/N, pitch = roll = 0), work_scale =
 *      compose_scale = 1 (vsb_calibrate_rig_scaled below takes a compose_scale), sphere radius = pano_width / 2*pi.  Host-side, runs once; computes K/R, seam-scale Voronoi
 *      masks (VoronoiSeamFinder, S/src/seam_finders.cpp:72-162), compose-scale ROIs and maps, then calls
 *      vsb_prepare / vsb_init_view / vsb_set_maps / vsb_set_gain.  gains may be NULL (all 1). -------------------- */
typedef struct vsb_rig_info {
    int projection;
    float scale;
    int src_w, src_h, num_views, num_bands;
    int roi_final[4], roi_padded[4];
    int view_roi[VSB_MAX_VIEWS][4]; /* {tl_x, tl_y, w, h} of each warped view */
} vsb_rig_info;
int vsb_rig_camera(int n_views, int i, int src_w, int src_h, double hfov_deg, float K[9], float R[9]);
int vsb_voronoi_seams(int n, const int *sizes_wh, const int *corners_xy, uint8_t *const *masks);
/* HOST buffers, no device needed: the projection maps exactly as vsb_calibrate_rig builds them -- buildWarpMapsKernel's arithmetic
 * (S/src/cuda/build_warp_maps.cu:88-152) with the host's sinf / cosf -- for the w x h rectangle at (tl_x, tl_y); dense rows. */
int vsb_host_build_maps(int projection, float scale, const float K[9], const float R[9], int tl_x, int tl_y, int w, int h,
                        float *xmap, float *ymap);
int vsb_calibrate_rig(vsb_stitcher *s, int projection, int pano_width, int src_w, int src_h, double hfov_deg,
                      const float *gains);
/* ---- the same calibration with every per-pixel loop on the DEVICE (the reference runs these stages on the GPU
 *      too, A/calibration.cpp:92-246): projection maps (buildWarp*Maps), mask warps, VoronoiSeamFinder, dilate / cuda::resize /
 *      bitwise_and, weight pyramids.  ROIs stay on the host (as RotationWarperBase::detectResultRoi does in the reference).  Maps
 *      agree with vsb_calibrate_rig to ~1e-3 px (device sinf / cosf), everything downstream of the maps is exact.  Keeps the
 *      seam-scale state vsb_estimate_gains needs. -------------------------------------------------------------------------- */
int vsb_calibrate_rig_device(vsb_stitcher *s, int projection, int pano_width, int src_w, int src_h, double hfov_deg,
                             const float *gains);
/* GainCompensator::feed (S/src/exposure_compensate.cpp:71-142) at run time: the n camera frames (device, CV_8UC3) are resized to
 * seam scale and warped like A/calibration.cpp:95,118, the pairwise overlap statistics are reduced on the device in the
 * reference's summation order, the n x n system is solved on the host (hal::LU64f restated).  n >= 4.  apply != 0 installs the
 * gains for the following frames (the reference's docx lists "dynamically update the gain compensation" as an open TODO). */
int vsb_estimate_gains(vsb_stitcher *s, const uint8_t *const *d_frames, size_t pitch_bytes, float *gains_out, int apply, void *stream);
/* the building blocks, on device buffers: VoronoiSeamFinder::find (S/src/seam_finders.cpp:72-162; masks modified in place),
 * MORPH_DILATE 3x3 (A/calibration.cpp:209,232), cuda::resize INTER_LINEAR on CV_8UC1 / CV_8UC3 (fx = fy = 0: factors from the
 * sizes; CW/src/resize.cpp:76-105), GainCompensator::feed on warped images (tight rows; gains as float64) */
int vsb_voronoi_seams_device(int n, const int *sizes_wh, const int *corners_xy, uint8_t *const *d_masks, void *stream);
int vsb_dilate3x3_u8(const uint8_t *d_src, int w, int h, uint8_t *d_dst, void *stream);
int vsb_resize_linear_u8(const uint8_t *d_src, int sw, int sh, size_t src_pitch, int channels, uint8_t *d_dst, int dw, int dh,
                         size_t dst_pitch, double fx, double fy, void *stream);
int vsb_gain_compensator_feed(int n, const uint8_t *const *d_imgs, const uint8_t *const *d_masks, const int *sizes_wh,
                              const int *corners_xy, double *gains_out, void *stream);
/* ---- compose_scale != 1 (A/calibration.cpp:137-205: COMPOSE_MEGAPIX; A/timed.cpp:74-81: the per-frame
 *      cuda::resize(full_imgs[i], images[i], Size(), compose_scale, compose_scale, INTER_LINEAR) in front of remap #1).
 *      vsb_compose_size: the sizes the reference derives from a compose_scale -- frame[2] = the frame remap #1 reads (the full
 *      frame, or cvRound(full * scale) when |scale - 1| > 0.1: the reference's test for resizing at all, and the dsize cuda::resize
 *      computes), which also sizes the blender (:159-178); map_src[2] = (int)(full * scale), the img_size its maps and masks are
 *      ALWAYS built for (:203-204).  Where the two differ (cvRound != truncation -- e.g. the reference's default COMPOSE_MEGAPIX
 *      = 1.4 on 1080p frames: 1578 vs 1577 columns -- or a scale within 0.1 of 1) the reference blends views whose maps were built
 *      for a slightly different frame; this library reproduces that.  vsb_rig_camera_scaled: the rig camera after focal, ppx,
 *      ppy *= compose_work_aspect (:168-172, doubles; work_scale = 1).  vsb_set_compose_scale (after vsb_set_maps; vsb_prepare
 *      resets it to 1): the maps address frame[], the frames handed to vsb_feed / vsb_compose / vsb_submit_host are full_w x
 *      full_h and go through the resize on the device first (bit-exact with the reference's kernel, CW/src/cuda/resize.cu:71-106).
 *      vsb_calibrate_rig_scaled: vsb_calibrate_rig (on_device = 0) / vsb_calibrate_rig_device (1) at that compose_scale: warper
 *      scale * (float)compose_scale, scaled cameras, ROIs / maps / masks as above, seam scale from the full frame. -------- */
int vsb_compose_size(int full_w, int full_h, double compose_scale, int frame[2], int map_src[2], int *resized);
int vsb_rig_camera_scaled(int n_views, int i, int src_w, int src_h, double hfov_deg, double compose_work_aspect, float K[9], float R[9]);
int vsb_set_compose_scale(vsb_stitcher *s, double compose_scale, int full_w, int full_h);
int vsb_calibrate_rig_scaled(vsb_stitcher *s, int projection, int pano_width, int src_w, int src_h, double hfov_deg,
                             const float *gains, double compose_scale, int on_device);
/* ---- modular wrap-around ROI (wrapAround, A/defs.h:25; A/meshwarper.cpp:93-102,620-626).  A camera that looks across +-pi gets
 *      the reference's full-panorama-width ROI (RotationWarperBase::detectResultRoi): two image parts at the two ends of one
 *      image, zeros between them -- four times the memory of its neighbours at 6 cameras, eight times at 12.  The split
 *      calibration installs such a camera as TWO views, column windows of its warped image (margins and origins chosen so that
 *      every pyramid level equals the full-width view's wherever a weight is non-zero: the panorama is the same bit for bit),
 *      so no buffer is ever panorama-wide.  vsb_split_plan (host only) tells how many views a rig needs -- create the handle with
 *      that num_views -- and which camera / columns each view shows; vsb_calibrate_rig_split is vsb_calibrate_rig (on_device = 0) or
 *      vsb_calibrate_rig_device (1; vsb_estimate_gains then installs each camera's gain on all of its views) for it (gains per
 *      CAMERA).  Per frame the caller passes, for every VIEW, the frame of its camera (vsb_view_window); vsb_set_mesh on a window
 *      view takes the camera's mesh.  Meshes must not move image content across the zero margin (3 * 2^num_bands + 8 px). ------ */
int vsb_split_plan(int projection, int pano_width, int n_cameras, int src_w, int src_h, double hfov_deg, int num_bands, int *n_views,
                   int *view_camera, int *view_x0, int *view_w);
int vsb_calibrate_rig_split(vsb_stitcher *s, int projection, int pano_width, int n_cameras, int src_w, int src_h, double hfov_deg,
                            const float *gains, int on_device);
int vsb_view_window(const vsb_stitcher *s, int view, int *camera, int *x0, int *full_w);
/* ---- stitch_calib with its own constants (A/calibration.cpp:256-305; A/defs.h:51-53: WORK_MEGAPIX 0.6, SEAM_MEAGPIX 0.01,
 *      COMPOSE_MEGAPIX 1.4).  vsb_ref_scales: work_scale / compose_scale = min(1, sqrt(MEGAPIX * 1e6 / area)) (negative: 1).
 *      vsb_rig_camera_work: calibrateCameras at a work scale (ppx = w * work_scale / 2, focal = ppx / tan(hfov / 2), doubles), then
 *      focal, ppx, ppy *= aspect.  vsb_calibrate_rig_megapix: the calibration with the sphere radius the REFERENCE uses --
 *      warped_image_scale = (float)cameras[0].focal at work scale, seam_work_aspect = seam_scale / work_scale, compose_work_aspect =
 *      compose_scale / work_scale -- instead of a pano_width; with (0.6, 1.4) it is the reference's default panorama, frames resized
 *      on the device as vsb_calibrate_rig_scaled does. ------------------------------------------------------------------------ */
int vsb_ref_scales(int src_w, int src_h, double work_megapix, double compose_megapix, double *work_scale, double *compose_scale);
int vsb_rig_camera_work(int n_views, int i, int src_w, int src_h, double hfov_deg, double work_scale, double aspect, float K[9], float R[9]);
int vsb_calibrate_rig_megapix(vsb_stitcher *s, int projection, int src_w, int src_h, double hfov_deg, const float *gains,
                              double work_megapix, double compose_megapix, int on_device);
int vsb_rig_info_get(const vsb_stitcher *s, vsb_rig_info *out);
int vsb_get_config(const vsb_stitcher *s, vsb_config *out);

/* ---- B1 + B2: custom_resize (A/resize.cu:9-45) and MeshWarper::convertMeshesToMap for one view
 *      (A/meshwarper.cpp:823-886).  mesh_x/mesh_y: host, rows x cols floats (vertex positions).
 *      Thread-safe w.r.t. feed/blend/compose; the new maps are used by the first compose submitted
 *      after this call returns. ------------------------------------------------------------------------ */
int vsb_set_mesh(vsb_stitcher *s, int i, const float *mesh_x, const float *mesh_y, int rows, int cols);
int vsb_custom_resize(const float *d_in, int cols, int rows, size_t in_pitch_bytes,
                      float *d_out, int tx, int ty, size_t out_pitch_bytes, void *stream);

/* ---- B4: per-view feed = stitch_online minus the H2D upload (A/timed.cpp:56-121):
 *      remap#1 -> gain -> remap#2 -> MultiBandBlender::feed_online (S/src/blenders.cpp:700-749) -------- */
int vsb_feed(vsb_stitcher *s, int i, const uint8_t *d_bgr, size_t pitch_bytes, void *stream);
/* ---- B4, inner boundary: MultiBandBlender::feed_online(cuda::GpuMat &gpu_img, int img_num, cuda::Stream &stream)
 *      (S/include/opencv2/stitching/detail/blenders.hpp:138, S/src/blenders.cpp:700-749) for callers that keep their own
 *      cuda::remap calls: d_warped = the warped CV_8UC3 view, size = the view's seam mask size. ------------------ */
int vsb_feed_warped(vsb_stitcher *s, int i, const uint8_t *d_warped, size_t pitch_bytes, void *stream);
/* ---- B5: MultiBandBlender::blend(dst, dst_mask, gpuOut, true) (S/src/blenders.cpp:758-832).
 *      d_out: caller-owned CV_16SC3 of roi_final size. ------------------------------------------------- */
int vsb_blend(vsb_stitcher *s, int16_t *d_out, size_t out_pitch_bytes, void *stream);
/* ---- stitch_one (A/timed.cpp:123-152) for n_frames frames in one submission.
 *      d_srcs[f*num_views + i] = device BGR frame of view i, frame f; d_outs[f] = device CV_16SC3 output. */
int vsb_compose(vsb_stitcher *s, int n_frames, const uint8_t *const *d_srcs, size_t src_pitch_bytes,
                int16_t *const *d_outs, size_t out_pitch_bytes, void *stream);
/* Same with HOST buffers (what A/timed.cpp:68 upload + the consumer's download do): H2D, compose, D2H;
 * returns after the outputs are in host memory. */
int vsb_compose_host(vsb_stitcher *s, int n_frames, const uint8_t *const *h_srcs, size_t src_pitch_bytes,
                     int16_t *const *h_outs, size_t out_pitch_bytes);
/* Asynchronous form: vsb_submit_host enqueues one submission (pinned host buffers must stay valid) and returns; up to two
 * submissions are in flight, so the download of one overlaps the upload of the next; vsb_wait_host blocks until the OLDEST
 * outstanding submission's panoramas are in host memory (the reference's result queue, A/timed.cpp:150,243, plays this role). */
int vsb_submit_host(vsb_stitcher *s, int n_frames, const uint8_t *const *h_srcs, size_t src_pitch_bytes,
                    int16_t *const *h_outs, size_t out_pitch_bytes);
int vsb_wait_host(vsb_stitcher *s);
/* ---- wire format in, consumer format out (both default to the reference's stage boundary).
 *      VSB_IN_NV12: the source pointers of vsb_feed / vsb_compose / vsb_compose_host are NV12 frames as the capture
 *      boards send them (A/defs.h:10-17: h rows of Y then h/2 rows of interleaved U,V; `pitch` = row pitch of both) and
 *      cv::cvtColor(mat, mat, CV_YUV2BGR_NV12) (A/networking.cpp:46) runs on the device, bit-exact.
 *      VSB_OUT_U8C3: the consumer's mat.convertTo(mat_8u, CV_8U) (A/timed.cpp:250) is fused into the final store; the
 *      output pointers are then CV_8UC3 buffers (pitch >= 3 * width; pass them cast to int16_t*). ---------------- */
enum { VSB_IN_BGR8 = 0, VSB_IN_NV12 = 1 };
enum { VSB_OUT_S16C3 = 0, VSB_OUT_U8C3 = 1 };
int vsb_set_formats(vsb_stitcher *s, int input_format, int output_format);
/* cv::cvtColor(CV_YUV2BGR_NV12) on device buffers: IMG/src/color.cpp:8759-8819 (YUV420sp2RGB888Invoker<0,0>) */
int vsb_nv12_to_bgr(const uint8_t *d_nv12, int w, int h, size_t pitch, uint8_t *d_bgr, size_t bgr_pitch, void *stream);
/* ---- consumer epilogue: what the reference's consumer thread does on the CPU after its download
 *      (A/timed.cpp:254-315), on the device: cv::resize(INTER_LINEAR) of the CV_8UC3 panorama to out_w x image_height
 *      (vsb_consumer_image_height = A/timed.cpp:254-270), then VSB_CONSUME_RGB: COLOR_BGR2RGB (:291), d_out = image_height rows
 *      of out_w RGB pixels; or VSB_CONSUME_I420: the image centred in a black out_w x out_h frame (:283-289) through
 *      COLOR_BGR2YUV_I420 (:310-315), d_out = out_w * out_h * 3 / 2 contiguous bytes (what kvazaar is fed).  Bit-exact. ---- */
enum { VSB_CONSUME_RGB = 0, VSB_CONSUME_I420 = 1 };
int vsb_consumer_image_height(int src_w, int src_h, int out_w, int out_h, int keep_aspect);
int vsb_consume(vsb_stitcher *s, const uint8_t *d_pano_u8, size_t pitch_bytes, int out_w, int out_h, int keep_aspect, int format,
                uint8_t *d_out, size_t out_pitch_bytes, void *stream);
/* ---- view-sharded multi-GPU mode (the reference is single-GPU, A/timed.cpp:495-496).  One process and
 *      one calibrated handle per GPU.  Rank r owns a canvas strip (its part of `blend`) and the views whose seam masks lie
 *      mostly inside it (their `stitch_online`).  Per frame: vsb_feed the owned views -> exchange the Gaussian sub-planes
 *      vsb_shard_rect lists (u8; NCCL send/recv by the caller, video-stitcher_b200/dist.py) -> vsb_blend writes the strip. */
int vsb_shard_set(vsb_stitcher *s, int rank, int world);
int vsb_shard_info(const vsb_stitcher *s, int *strip_x0, int *strip_x1, unsigned *owned_view_mask);
/* rect = {x0, y0, w, h} of Gaussian level `level` of `view` (plane coordinates) that rank dst_rank reads; w = 0: nothing */
int vsb_shard_rect(const vsb_stitcher *s, int dst_rank, int view, int level, int rect[4]);
/* batched form: n_frames frames per exchange and ONE message per peer.  vsb_shard_plan(owners[num_views]) fixes the per-peer
 * rectangle lists (same order on both sides); vsb_shard_pack gathers what `peer` reads of this rank's planes into a contiguous
 * buffer (n_frames * send_bytes_per_frame), vsb_shard_unpack scatters what this rank reads of `peer`'s.  Per submission:
 * vsb_feed_batch over the owned view range(s) -> pack -> send / recv (caller's transport) -> unpack -> vsb_blend_batch. */
int vsb_shard_plan(vsb_stitcher *s, const int *owners);
int vsb_shard_peer_bytes(const vsb_stitcher *s, int peer, size_t *send_bytes_per_frame, size_t *recv_bytes_per_frame);
int vsb_shard_pack(vsb_stitcher *s, int peer, int n_frames, void *d_buf, void *stream);
int vsb_shard_unpack(vsb_stitcher *s, int peer, int n_frames, void *d_buf, void *stream);
/* Native transport (no caller-side exchange): vsb_shard_unique_id on rank 0 returns VSB_SHARD_ID_BYTES bytes (an ncclUniqueId) that
 * the host hands to every rank (MPI / TCP / torch.distributed -- control plane only); vsb_shard_init(rank, world, id) = vsb_shard_set +
 * vsb_shard_plan with the ownership rule every rank evaluates identically + ncclCommInitRank.  libnccl.so.2 is loaded with dlopen
 * at that point (no link-time dependency).  vsb_shard_compose is stitch_one (A/timed.cpp:123-152) for n_frames frames of ONE frame
 * stream on `world` GPUs: front half of the owned views -> one grouped ncclSend / ncclRecv per peer (u8 Gaussian sub-planes over
 * NVLink, stream-ordered) -> back half of the owned strip, written into the FULL-SIZE buffers d_outs[f] (every rank writes its
 * strip; gather only if the consumer needs one contiguous frame).  d_srcs[f * num_views + v]: entries of views this rank does not own
 * may be NULL.  Submissions alternate between two halves of the frame slots when 2 * n_frames <= max_batch, so the exchange and back
 * half of submission k overlap the front half of submission k + 1 (use two caller streams alternately to let them). */
#define VSB_SHARD_ID_BYTES 128
int vsb_shard_unique_id(void *id128);
int vsb_shard_init(vsb_stitcher *s, int rank, int world, const void *id128);
int vsb_shard_compose(vsb_stitcher *s, int n_frames, const uint8_t *const *d_srcs, size_t src_pitch_bytes,
                      int16_t *const *d_outs, size_t out_pitch_bytes, void *stream);
int vsb_shard_exchange_bytes(const vsb_stitcher *s, size_t *send_bytes_per_frame, size_t *recv_bytes_per_frame);
/* stitch_online (A/timed.cpp:56-121) for views [v0, v1) of n_frames frames; d_srcs[f * (v1 - v0) + (i - v0)];
 * MultiBandBlender::blend (S/src/blenders.cpp:758-832) for n_frames frames */
int vsb_feed_batch(vsb_stitcher *s, int v0, int v1, int n_frames, const uint8_t *const *d_srcs, size_t pitch_bytes, void *stream);
int vsb_blend_batch(vsb_stitcher *s, int n_frames, int16_t *const *d_outs, size_t out_pitch_bytes, void *stream);
/* device address of Gaussian level `level` (0, 1, 2..num_bands) of `view`, frame slot `frame`: [3][h][w] u8 */
int vsb_get_plane(vsb_stitcher *s, int view, int level, int frame, void **ptr, int *w, int *h);
/* number of kernels the last vsb_compose / vsb_feed+vsb_blend submission launched */
int vsb_last_launch_count(const vsb_stitcher *s);

/* ---- measurement: with profiling on, CUDA events bracket every kernel of a submission on the caller's stream
 *      (the reference keeps std::chrono stamps per stage in times[5], A/timed.cpp:43-44,61-63).  vsb_get_profile
 *      returns, for the last submission, each kernel's name, its device time and its ALGORITHMIC bytes (its own
 *      compulsory input + output, static tables excluded). ------------------------------- */
int vsb_set_profiling(vsb_stitcher *s, int on);
int vsb_get_profile(vsb_stitcher *s, int max_stages, int *n_stages, const char **names, float *ms, double *bytes);

/* ---- B7: device-launcher layer (the .cpp -> .cu seam inside the reference's OpenCV), one primitive per
 *      reference kernel, interleaved OpenCV layouts, pitches in bytes ---------------------------------- */
/* cuda::remap LINEAR/BORDER_CONSTANT(0) on CV_8UC3: CW/src/cuda/remap.cu:56-86 */
int vsb_remap_linear_u8c3(const uint8_t *d_src, int sw, int sh, size_t src_pitch,
                          const float *d_xmap, const float *d_ymap, size_t map_pitch,
                          uint8_t *d_dst, int dw, int dh, size_t dst_pitch, void *stream);
/* GpuMat::convertTo(type, alpha) on CV_8U: CORE/src/cuda/gpu_mat.cu:488-512 */
int vsb_gain_u8(uint8_t *d_buf, int width_bytes, int h, size_t pitch, float gain, void *stream);
/* cuda::copyMakeBorder(BORDER_REFLECT)+convertTo(CV_16S): CA/src/cuda/copy_make_border.cu:92-124 */
int vsb_border_reflect_u8c3_to_s16c3(const uint8_t *d_src, int w, int h, size_t src_pitch, int top, int bottom,
                                     int left, int right, int16_t *d_dst, size_t dst_pitch, void *stream);
/* cuda::pyrDown / cuda::pyrUp on CV_16SC3: CW/src/cuda/pyr_down.cu:55-188, pyr_up.cu:55-157 */
int vsb_pyr_down_s16c3(const int16_t *d_src, int w, int h, size_t src_pitch, int16_t *d_dst, size_t dst_pitch, void *stream);
int vsb_pyr_up_s16c3(const int16_t *d_src, int w, int h, size_t src_pitch, int16_t *d_dst, size_t dst_pitch, void *stream);
/* cuda::pyrDown on CV_32FC1 (weight pyramids, S/src/blenders.cpp:422-423) */
int vsb_pyr_down_f32(const float *d_src, int w, int h, size_t src_pitch, float *d_dst, size_t dst_pitch, void *stream);
/* addSrcWeightGpu32F / normalizeUsingWeightMapGpu32F: S/src/cuda/multiband_blend.cu:36-60,85-108 */
int vsb_add_src_weight_32f(const int16_t *d_src, size_t src_pitch, const float *d_w, size_t w_pitch,
                           int16_t *d_dst, size_t dst_pitch, float *d_dst_w, size_t dst_w_pitch,
                           int w, int h, void *stream);
int vsb_normalize_32f(const float *d_w, size_t w_pitch, int16_t *d_src, size_t src_pitch, int w, int h, void *stream);

/* ---- introspection for parity tests: copy an internal buffer of frame slot `frame` to host.
 *      what: 0 = warped view Q (u8x3 interleaved, roi size), 1 = Gaussian level k (s16x3 interleaved, k>=0,
 *      bordered size >> k), 2 = weight level k (f32), 3 = canvas weight sum level k (f32, view ignored),
 *      4 = x mesh map (f32, roi size), 5 = y mesh map, 8 = x projection map, 9 = y projection map (f32, roi size).
 *      `bytes` must match exactly. ------------------------------------------------------------------------- */
int vsb_debug_read(vsb_stitcher *s, int what, int view, int level, int frame, void *h_dst, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* VSB200_H */
