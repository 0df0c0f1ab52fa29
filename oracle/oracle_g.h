/*
 * oracle_g.h -- TEST INFRASTRUCTURE ONLY (never linked, imported or called by the product path).
 *
 * "oracle-G": a scalar CPU restatement of the arithmetic of the reference's per-frame CUDA
 * compose path (ultravideo/video-stitcher): projection maps -> bilinear remap -> gain ->
 * CPW-mesh remap -> REFLECT border -> Gaussian/Laplacian pyramid -> seam-masked weighted
 * add -> normalise -> collapse -> masked crop.  Every function cites the reference file:line
 * it follows (paths relative to /root/reference).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * use this library.
 */
#ifndef ORACLE_G_H
#define ORACLE_G_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { OG_PROJ_SPHERICAL = 0, OG_PROJ_CYLINDRICAL = 1 };

/* ---------------------------------------------------------------- camera rig + projector */
void og_rig_camera(int n_views, int i, int src_w, int src_h, double hfov_deg, float K[9], float R[9]);
void og_rig_camera_work(int n_views, int i, int src_w, int src_h, double hfov_deg, double work_scale, double aspect, float K[9], float R[9]);
void og_rig_camera_scaled(int n_views, int i, int src_w, int src_h, double hfov_deg, double compose_work_aspect, float K[9], float R[9]);
void og_projector(const float K[9], const float R[9], float k_rinv[9], float r_kinv[9], float rinv[9]);
/* roi = {tl_x, tl_y, width, height} with width = br_x - tl_x + 1 (size of the maps buildMaps makes) */
void og_warp_roi(int proj, float scale, const float K[9], const float R[9], int src_w, int src_h, int roi[4]);
void og_build_maps(int proj, float scale, const float K[9], const float R[9], int tl_x, int tl_y,
                   int w, int h, float *xmap, float *ymap);

/* ---------------------------------------------------------------- remap / gain / resize */
void og_remap_linear_u8(const uint8_t *src, int sw, int sh, int cn, size_t sstep,
                        const float *xmap, const float *ymap, size_t mstep,
                        uint8_t *dst, int dw, int dh, size_t dstep);
void og_remap_nearest_u8c1(const uint8_t *src, int sw, int sh, size_t sstep,
                           const float *xmap, const float *ymap, size_t mstep,
                           uint8_t *dst, int dw, int dh, size_t dstep);
void og_remap_u8_border(const uint8_t *src, int sw, int sh, int cn, size_t sstep, const float *xmap, const float *ymap, size_t mstep,
                        uint8_t *dst, int dw, int dh, size_t dstep, int interp, int border);
void og_gain_u8(uint8_t *buf, size_t n, float gain);
void og_cuda_resize_linear_u8(const uint8_t *src, int sw, int sh, int cn, uint8_t *dst, int dw, int dh, double fx, double fy);
int og_gain_compensator_feed(int n, const uint8_t *const *imgs, const uint8_t *const *masks, const int *sizes, const int *corners, double *gains);
void og_resize_linear_u8c1(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh);
void og_dilate3x3_u8c1(const uint8_t *src, int w, int h, uint8_t *dst);

/* ---------------------------------------------------------------- wire format in / consumer format out */
/* cv::cvtColor(CV_YUV2BGR_NV12): h rows of Y (step bytes apart) followed by h/2 rows of interleaved U,V; w and h even */
void og_nv12_to_bgr(const uint8_t *nv12, int w, int h, size_t step, uint8_t *bgr, size_t bgr_step);
/* GpuMat::convertTo(CV_8U) of the CV_16SC3 panorama: saturate_cast<uchar>(short) */
void og_s16_to_u8(const int16_t *src, size_t n, uint8_t *dst);

/* consumer epilogue on the CPU in the reference (360_stitcher/timed.cpp:254-315) */
/* cv::resize(src, dst, Size(dw, dh), 0, 0, INTER_LINEAR) on CV_8UC3 (fixed point, IMG/src/resize.cpp:3930-4021, 1907-1957, 1993-2016) */
void og_resize_linear_u8c3(const uint8_t *src, int sw, int sh, size_t sstep, uint8_t *dst, int dw, int dh, size_t dstep);
/* cv::cvtColor(COLOR_BGR2YUV_I420): w, h even; dst = w*h Y bytes, then (w/2)*(h/2) U, then V (IMG/src/color.cpp:9082-9160) */
void og_bgr_to_i420(const uint8_t *bgr, int w, int h, size_t step, uint8_t *yuv);
/* image height the consumer resizes to (timed.cpp:254-270) */
int og_consumer_image_height(int src_w, int src_h, int out_w, int out_h, int keep_aspect);

/* ---------------------------------------------------------------- CPW mesh -> backward map */
void og_custom_resize(const float *in, int cols, int rows, float *out, int tx, int ty);
/* half-res table (W/2 x H/2) of step m1-m3; then og_custom_resize gives the full map (m4) */
void og_mesh_to_half_table(const float *mesh_x, const float *mesh_y, int mesh_rows, int mesh_cols,
                           int W, int H, float *warp_x, float *warp_y);
void og_mesh_to_map(const float *mesh_x, const float *mesh_y, int mesh_rows, int mesh_cols,
                    int W, int H, float *map_x, float *map_y);

/* ---------------------------------------------------------------- pyramid primitives */
void og_border_reflect_u8c3_to_s16(const uint8_t *src, int w, int h, size_t sstep,
                                   int top, int bottom, int left, int right, int16_t *dst);
void og_border_constant_f32(const float *src, int w, int h, int top, int bottom, int left, int right, float *dst);
void og_pyr_down_s16(const int16_t *src, int w, int h, int cn, int16_t *dst);   /* dst ((h+1)/2, (w+1)/2) */
void og_pyr_up_s16(const int16_t *src, int w, int h, int cn, int16_t *dst);     /* dst (2h, 2w) */
void og_pyr_down_f32(const float *src, int w, int h, float *dst);
/* exact-integer twins used to prove the fp32 forms are rounding-order independent on s16 data */
void og_pyr_down_s16_int(const int16_t *src, int w, int h, int cn, int16_t *dst);
void og_pyr_up_s16_int(const int16_t *src, int w, int h, int cn, int16_t *dst);
/* round-half-up twins = the vendored CPU cv::pyrDown / cv::pyrUp on CV_16S (IMG/src/pyramids.cpp:52-57) */
void og_pyr_down_s16_halfup(const int16_t *src, int w, int h, int cn, int16_t *dst);
void og_pyr_up_s16_halfup(const int16_t *src, int w, int h, int cn, int16_t *dst);

/* ---------------------------------------------------------------- seam masks */
void og_voronoi_find(int n, const int *sizes_wh, const int *corners_xy, uint8_t **masks);

/* ---------------------------------------------------------------- MultiBandBlender (authors' GPU variant) */
typedef struct og_blender og_blender;
og_blender *og_blender_create(int num_bands);
void og_blender_destroy(og_blender *b);
/* prepare(corners, sizes): returns 0; fills geometry */
int og_blender_prepare(og_blender *b, int n, const int *corners_xy, const int *sizes_wh);
int og_blender_num_bands(const og_blender *b);
void og_blender_set_cpu_pyramids(og_blender *b, int on);
void og_blender_set_view_weight(og_blender *b, int i, int level, const float *w);
void og_blender_dst_roi(const og_blender *b, int roi_final[4], int roi_padded[4]);
/* init_gpu(mask, tl): views must be added in order */
int og_blender_init_view(og_blender *b, const uint8_t *mask, int mw, int mh, size_t mstep, int tl_x, int tl_y);
/* geometry of view i: out = {top, bottom, left, right, x_tl, y_tl, x_br, y_br} */
void og_blender_view_geom(const og_blender *b, int i, int out[8]);
const float *og_blender_view_weight(const og_blender *b, int i, int level, int *w, int *h);
const float *og_blender_dst_weight(const og_blender *b, int level, int *w, int *h); /* valid after feeds */
const int16_t *og_blender_dst_level(const og_blender *b, int level, int *w, int *h);
const int16_t *og_blender_src_level(const og_blender *b, int i, int level, int *w, int *h); /* laplacian of last feed */
/* addSrcWeightKernel32F / normalizeUsingWeightKernel32F (S/src/cuda/multiband_blend.cu:36-50, 85-99) on densely packed CV_16SC3 / CV_32F arrays */
void og_add_src_weight_32f(const int16_t *src, const float *weight, int16_t *dst, float *dst_weight, int rows, int cols);
void og_normalize_32f(const float *weight, int16_t *src, int rows, int cols);
void og_blender_feed_online(og_blender *b, int i, const uint8_t *img, int w, int h, size_t step);
/* out: CV_16SC3 of dst_roi_final size, tightly packed; mask_out (optional) u8 */
void og_blender_blend(og_blender *b, int16_t *out, uint8_t *mask_out);

void og_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
