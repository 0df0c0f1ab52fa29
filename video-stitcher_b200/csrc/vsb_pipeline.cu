// vsb_pipeline.cu -- the fused per-frame compose path (sm_100a) and the handle behind the C ABI.
//
// Reference path replaced: stitch_online x N -> MultiBandBlender::feed_online x N ->
// MultiBandBlender::blend, ~190 kernel launches and ~1.1 GB of HBM traffic per frame in the reference
// (360_stitcher/timed.cpp:56-152, sources/modules/stitching/src/blenders.cpp:700-832).
//
// Here one frame (or a batch of F frames) is five launches (num_bands >= 3):
//   K1 k_remap_stage1   src (u8x3)            -> P  = gain(remap#1(src))          u8x3 interleaved, ROI size
//   K2 k_remap_stage2   P + CPW mesh maps     -> G0 = REFLECT-bordered remap#2(P) u8 planar, bordered size
//   K3 k_down2          G0                    -> G2 (two pyrDown levels, G1 on chip)              vsb_blend_kernels.cuh
//   K4 k_coarse         G2 of every view      -> C2 = collapsed blended levels 2..nb (canvas)     vsb_blend_kernels.cuh
//   K5 k_blend          G0, G2, masks, C2     -> out (CV_16SC3): levels 0/1, collapse, mask, crop vsb_blend_kernels.cuh
// The destination pyramid, the per-view Laplacian pyramids, the `ups` buffers and the per-frame clears of the
// reference never exist; which view touches which tile is a static table (build_plan).
// num_bands < 3 (bordered sizes not multiples of 8) takes the generic per-level kernels further below.
//
// HBM layout: all per-view intermediates are PLANAR u8 (one plane per colour channel) with the bordered
// width (a multiple of 2^nb) as row length, so rows of every level start word aligned.
#include <dlfcn.h>
#include <sys/resource.h>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: named ranges per stage for nsys / ncu timelines (no cost without a tool attached)
#include <nccl.h>  // types only: the library is loaded with dlopen when the view-sharded mode is initialised (single-GPU use needs no NCCL)

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "vsb_internal.h"
#include "vsb_blend_kernels.cuh"

namespace vsb {

// ============================================================================================ device side

constexpr int MAX_SPLIT = 4;  // sub-batches of one submission's back half (internal streams)
constexpr int HOST_DEPTH = 2;  // vsb_submit_host: submissions in flight (each with its own device staging)
constexpr int HOST_SUB = 2;    // ... frames per upload -> compose -> download pipeline stage

struct TileMap {
    int n;
    int start[MAXV + 1];  // prefix sums of tiles per view
    int tiles_x[MAXV];
};

__device__ __forceinline__ int find_view(const TileMap &tm, int tile)
{
    int v = 0;
#pragma unroll 1
    while (v + 1 < tm.n && tile >= tm.start[v + 1]) ++v;
    return v;
}

// ---- K1: remap#1 (projection maps) + gain ------------------------------------------------------------
// One CTA = one 128 x 8 tile of one view's warped ROI (4 consecutive pixels per thread).  Tiles whose maps address
// nothing inside the camera frame are not in the list: P is zeroed once and stays 0 there, which is what the
// reference's BORDER_CONSTANT remap writes every frame.
struct Stage1View {
    const float *xmap, *ymap;   // roi_w x roi_h, pitch map_pitch bytes
    uint8_t *P;                 // frame 0
    size_t map_pitch, p_pitch, p_frame_stride;
    int w, h, src_w, src_h;
    float gain;
};
struct Stage1Params {
    const uint32_t *tiles;      // view | tile_x << 8 | tile_y << 20
    Stage1View v[MAXV];
    const uint8_t *src[MAX_BATCH * MAXV];  // [frame slot][view - v0]
    size_t src_pitch;
    int v0, n_views, f0;                   // f0: first frame slot of this launch
};

constexpr int RM_BX = 32, RM_BY = 8, RM_PX = 4;  // 4 consecutive pixels per thread

__global__ void __launch_bounds__(RM_BX *RM_BY) k_remap_stage1(const __grid_constant__ Stage1Params p)
{
    const unsigned tile = __ldg(p.tiles + blockIdx.x);
    const int vi = tile & 0xff;
    const Stage1View &V = p.v[vi];
    const int x0 = ((int)((tile >> 8) & 0xfff) * RM_BX + threadIdx.x) * RM_PX, y = (int)(tile >> 20) * RM_BY + threadIdx.y;
    if (x0 >= V.w || y >= V.h) return;
    const int f = blockIdx.y + p.f0;
    const uint8_t *src = p.src[f * p.n_views + vi - p.v0];
    const float *mx = (const float *)((const char *)V.xmap + (size_t)y * V.map_pitch) + x0;
    const float *my = (const float *)((const char *)V.ymap + (size_t)y * V.map_pitch) + x0;
    uint8_t *dst = V.P + (size_t)f * V.p_frame_stride + (size_t)y * V.p_pitch + (size_t)x0 * 3;
    const int n = min(RM_PX, V.w - x0);
    float fx[RM_PX], fy[RM_PX];
    if (n == RM_PX) {  // map rows are 16-byte aligned and x0 % 4 == 0
        const float4 a = __ldg((const float4 *)mx), b = __ldg((const float4 *)my);
        fx[0] = a.x; fx[1] = a.y; fx[2] = a.z; fx[3] = a.w;
        fy[0] = b.x; fy[1] = b.y; fy[2] = b.z; fy[3] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < RM_PX; ++i) { fx[i] = i < n ? __ldg(mx + i) : -1.f; fy[i] = i < n ? __ldg(my + i) : -1.f; }
    }
    unsigned px[RM_PX];
#pragma unroll
    for (int i = 0; i < RM_PX; ++i) px[i] = remap_gain_px<true>(src, p.src_pitch, V.src_w, V.src_h, fx[i], fy[i], V.gain);
    if (n == RM_PX) {  // 12 bytes = three aligned 32-bit stores (p_pitch % 4 == 0, x0 % 4 == 0)
        unsigned *d32 = (unsigned *)dst;
        d32[0] = px[0] | (px[1] << 24);
        d32[1] = (px[1] >> 8) | (px[2] << 16);
        d32[2] = (px[2] >> 16) | (px[3] << 8);
    } else {
        for (int i = 0; i < n; ++i) { dst[3 * i] = px[i] & 0xff; dst[3 * i + 1] = (px[i] >> 8) & 0xff; dst[3 * i + 2] = (px[i] >> 16) & 0xff; }
    }
}

// ---- K2: CPW-mesh remap#2 + REFLECT border + interleaved -> planar ------------------------------------
// One CTA = one 128 x 8 tile of one view's BORDERED plane; only tiles some later kernel reads are in the list.
struct Stage2View {
    const uint8_t *P;
    const float *xmesh, *ymesh;  // roi_w x roi_h (null when enable_local == 0)
    uint8_t *G0;                 // planar 3 x (bw*bh), frame 0
    size_t p_pitch, p_frame_stride, map_pitch, g0_frame_stride;
    int w, h, bw, bh, top, left;
};
struct Stage2Params {
    const uint32_t *tiles;
    Stage2View v[MAXV];
    int f0;
};

__global__ void __launch_bounds__(RM_BX *RM_BY) k_remap_stage2(const __grid_constant__ Stage2Params p)
{
    const unsigned tile = __ldg(p.tiles + blockIdx.x);
    const Stage2View &V = p.v[tile & 0xff];
    const int bx0 = ((int)((tile >> 8) & 0xfff) * RM_BX + threadIdx.x) * RM_PX, by = (int)(tile >> 20) * RM_BY + threadIdx.y;
    if (bx0 >= V.bw || by >= V.bh) return;
    const int f = blockIdx.y + p.f0;
    const uint8_t *P = V.P + (size_t)f * V.p_frame_stride;
    const int y = reflect_idx(by - V.top, V.h);
    const int n = min(RM_PX, V.bw - bx0);
    const bool inside = bx0 >= V.left && bx0 + RM_PX <= V.left + V.w;  // no reflection in x for this thread's pixels
    unsigned px[RM_PX];
#pragma unroll
    for (int i = 0; i < RM_PX; ++i) {
        px[i] = 0u;
        if (i >= n) continue;
        const int x = inside ? bx0 + i - V.left : reflect_idx(bx0 + i - V.left, V.w);
        if (V.xmesh) {
            const float fx = __ldg((const float *)((const char *)V.xmesh + (size_t)y * V.map_pitch) + x);
            const float fy = __ldg((const float *)((const char *)V.ymesh + (size_t)y * V.map_pitch) + x);
            px[i] = remap_gain_px<false>(P, V.p_pitch, V.w, V.h, fx, fy, 1.f);
        } else {
            const uint8_t *s = P + (size_t)y * V.p_pitch + (size_t)x * 3;
            px[i] = (unsigned)__ldg(s) | ((unsigned)__ldg(s + 1) << 8) | ((unsigned)__ldg(s + 2) << 16);
        }
    }
    // interleaved -> planar: byte c of the four pixels
    const unsigned lo01 = __byte_perm(px[0], px[1], 0x5140u), lo23 = __byte_perm(px[2], px[3], 0x5140u);  // b0 b1 g0 g1 | b2 b3 g2 g3
    const unsigned c0 = __byte_perm(lo01, lo23, 0x5410u), c1 = __byte_perm(lo01, lo23, 0x7632u);
    const unsigned c2 = __byte_perm(__byte_perm(px[0], px[1], 0x0062u), __byte_perm(px[2], px[3], 0x0062u), 0x5410u);
    const size_t plane = (size_t)V.bw * V.bh;
    uint8_t *g = V.G0 + (size_t)f * V.g0_frame_stride + (size_t)by * V.bw + bx0;
    if (n == RM_PX && (V.bw & 3) == 0) {
        *(unsigned *)g = c0;
        *(unsigned *)(g + plane) = c1;
        *(unsigned *)(g + 2 * plane) = c2;
    } else {
        for (int i = 0; i < n; ++i) { g[i] = (c0 >> (8 * i)) & 0xff; g[plane + i] = (c1 >> (8 * i)) & 0xff; g[2 * plane + i] = (c2 >> (8 * i)) & 0xff; }
    }
}

// ---- K1 / K2, table-driven form (the one the compose path runs whenever the caller's frames are 4-byte aligned) -------
// Tap tables: per pixel one int32 window offset + four fp32 weights, stored as five planes of `tab_pitch` elements per
// row (16-byte aligned rows: one 128-bit load per plane and thread).  Built by k_build_taps1 when the projection maps or
// the caller's row pitch change, by k_build_taps2 when a CPW mesh is published (vsb_set_mesh, double buffered with it).
struct TapTable {
    const int *off;     // [rows][tab_pitch]
    const float *w;     // [4][rows][tab_pitch]
    size_t plane;       // elements per weight plane
    int tab_pitch;
};

// a non-zero weight so small that the 2^126-scaled, denormal-tap chain of remap_tab_px could round differently from the reference
__device__ __forceinline__ bool tap_weight_tiny(const TapEntry &e)
{
    return (e.wa != 0.f && e.wa < TAP_WMIN) || (e.wb != 0.f && e.wb < TAP_WMIN) || (e.wc != 0.f && e.wc < TAP_WMIN) || (e.wd != 0.f && e.wd < TAP_WMIN);
}

// One thread per table entry of remap #1.  Entries whose 32-bit window loads would leave the caller's image buffer
// (last bytes of the last row) are marked TAP_SLOW and take the coordinate-driven edge routine in the frame kernel.
// nv12 != 0: the table addresses the luma plane of an NV12 frame (one byte per pixel, byte loads: no end-of-buffer entries);
// a tiny weight then raises *unsafe and the host keeps the BGR staging path for this view.
__global__ void k_build_taps1(const float *__restrict__ xmap, const float *__restrict__ ymap, size_t map_pitch, int w, int h,
                              int sw, int sh, unsigned pitch, int *__restrict__ off, float *__restrict__ wgt, size_t plane, int tab_pitch,
                              int nv12, int *unsafe)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= tab_pitch || y >= h) return;
    TapEntry e;
    e.off = 0; e.wa = e.wb = e.wc = e.wd = 0.f;
    if (x < w) {
        const float fx = *((const float *)((const char *)xmap + (size_t)y * map_pitch) + x);
        const float fy = *((const float *)((const char *)ymap + (size_t)y * map_pitch) + x);
        e = make_tap_entry(fx, fy, sw, sh, pitch, 0u, false, nv12 ? 1u : 3u);
        if (nv12) {
            if (tap_weight_tiny(e)) atomicOr(unsafe, 1);
        } else {
            const unsigned p2 = (unsigned)e.off + pitch;                       // second row of the window
            const unsigned end = (p2 & ~3u) + ((p2 & 3u) == 3u ? 12u : 8u);    // one past the last byte the word loads touch
            if (end > (unsigned)(sh - 1) * pitch + (unsigned)sw * 3u) e.off = TAP_SLOW;
            if (tap_weight_tiny(e)) e.off = TAP_SLOW;                          // the scaled chain could leave the normal range
        }
    }
    const size_t i = (size_t)y * tab_pitch + x;
    off[i] = e.off;
    wgt[i] = __fmul_rn(e.wa, TAP_WSCALE); wgt[plane + i] = __fmul_rn(e.wb, TAP_WSCALE);
    wgt[2 * plane + i] = __fmul_rn(e.wc, TAP_WSCALE); wgt[3 * plane + i] = __fmul_rn(e.wd, TAP_WSCALE);
}

// One thread per BORDERED pixel of remap #2: the REFLECT border is resolved here, the taps address the zero-framed P.
__global__ void k_build_taps2(const float *__restrict__ xmesh, const float *__restrict__ ymesh, size_t map_pitch, int w, int h,
                              int bw, int bh, int top, int left, unsigned p_pitch, unsigned origin,
                              int *__restrict__ off, float *__restrict__ wgt, size_t plane, int tab_pitch, int *unsafe)
{
    const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y * blockDim.y + threadIdx.y;
    if (bx >= tab_pitch || by >= bh) return;
    TapEntry e;
    e.off = (int)origin; e.wa = e.wb = e.wc = e.wd = 0.f;
    if (bx < bw) {
        const int x = reflect_idx(bx - left, w), y = reflect_idx(by - top, h);
        const float fx = *((const float *)((const char *)xmesh + (size_t)y * map_pitch) + x);
        const float fy = *((const float *)((const char *)ymesh + (size_t)y * map_pitch) + x);
        e = make_tap_entry(fx, fy, w, h, p_pitch, origin, true);
        if (tap_weight_tiny(e)) atomicOr(unsafe, 1);  // (never seen: needs a map coordinate within 2^-90 of an integer) -> coordinate kernel
    }
    const size_t i = (size_t)by * tab_pitch + bx;
    off[i] = e.off;
    wgt[i] = __fmul_rn(e.wa, TAP_WSCALE); wgt[plane + i] = __fmul_rn(e.wb, TAP_WSCALE);
    wgt[2 * plane + i] = __fmul_rn(e.wc, TAP_WSCALE); wgt[3 * plane + i] = __fmul_rn(e.wd, TAP_WSCALE);
}

struct Stage1TabView {
    TapTable tab;
    const float *xmap, *ymap;   // only for TAP_SLOW entries
    uint8_t *P;
    size_t map_pitch, p_pitch, p_frame_stride;
    int w, h, src_w, src_h;
    float gain;
};
struct Stage1TabParams {
    const uint32_t *tiles;
    Stage1TabView v[MAXV];
    const uint8_t *src[MAX_BATCH * MAXV];  // [frame slot][view - v0]
    unsigned src_pitch;
    int v0, n_views, n_frames, f0;         // frame slots f0 .. f0 + n_frames - 1
};

// One CTA = one 128 x 8 tile for ALL frames of the submission: the table entries of the thread's 4 pixels are loaded once
// and stay in registers while the frames stream through (window loads -> fmul / fma chain -> store per frame).
// LANES = false: a thread owns 4 consecutive pixels (128-bit table loads, one 12-byte store).
// LANES = true : a thread owns pixels lane, lane + 32, lane + 64, lane + 96 of the tile row, so the 32 window loads of one
//                instruction are ONE source-pixel step apart instead of four and fall into two or three cache lines instead
//                of five or six (the kernel is bound by L1 wavefronts, not by issue slots); the row is transposed through
//                shared memory so that the stores stay full, coalesced words.
#ifndef VSB_RM1_MINB
#define VSB_RM1_MINB 4
#endif
#ifndef VSB_RM_UNROLL
#define VSB_RM_UNROLL 1
#endif
#ifndef VSB_K1_PREFETCH
#define VSB_K1_PREFETCH 0  // 1: prefetch.global.L2 of the next frame's tap windows, 2: prefetch.global.L1
#endif
constexpr int RM_UNROLL = VSB_RM_UNROLL;  // frames of the per-tile loop in flight per thread
template <bool LANES>
__global__ void __launch_bounds__(RM_BX *RM_BY, VSB_RM1_MINB) k_remap_stage1_tab(const __grid_constant__ Stage1TabParams p)
{
    __shared__ unsigned sP[LANES ? RM_BY : 1][LANES ? RM_BX * RM_PX + 1 : 1];
    const unsigned tile = __ldg(p.tiles + blockIdx.x);
    const int vi = tile & 0xff;
    const Stage1TabView &V = p.v[vi];
    const int tx0 = (int)((tile >> 8) & 0xfff) * (RM_BX * RM_PX), y = (int)(tile >> 20) * RM_BY + threadIdx.y;
    const int x0 = LANES ? tx0 + threadIdx.x : tx0 + threadIdx.x * RM_PX;
    constexpr int XSTEP = LANES ? RM_BX : 1;
    if (y >= V.h || (!LANES && x0 >= V.w)) return;  // LANES: the whole warp stays for the shared-memory transpose
    const size_t i = (size_t)y * V.tab.tab_pitch + x0;
    int off[RM_PX];
    float wa[RM_PX], wb[RM_PX], wc[RM_PX], wd[RM_PX];
    if (LANES) {
#pragma unroll
        for (int k = 0; k < RM_PX; ++k) {
            const bool in = x0 + k * XSTEP < V.tab.tab_pitch;
            off[k] = in ? __ldg(V.tab.off + i + k * XSTEP) : 0;
            wa[k] = in ? __ldg(V.tab.w + i + k * XSTEP) : 0.f;
            wb[k] = in ? __ldg(V.tab.w + V.tab.plane + i + k * XSTEP) : 0.f;
            wc[k] = in ? __ldg(V.tab.w + 2 * V.tab.plane + i + k * XSTEP) : 0.f;
            wd[k] = in ? __ldg(V.tab.w + 3 * V.tab.plane + i + k * XSTEP) : 0.f;
        }
    } else {
        const int4 o = __ldg((const int4 *)(V.tab.off + i));
        const float4 A = __ldg((const float4 *)(V.tab.w + i)), B = __ldg((const float4 *)(V.tab.w + V.tab.plane + i));
        const float4 C = __ldg((const float4 *)(V.tab.w + 2 * V.tab.plane + i)), D = __ldg((const float4 *)(V.tab.w + 3 * V.tab.plane + i));
        off[0] = o.x; off[1] = o.y; off[2] = o.z; off[3] = o.w;
        wa[0] = A.x; wa[1] = A.y; wa[2] = A.z; wa[3] = A.w; wb[0] = B.x; wb[1] = B.y; wb[2] = B.z; wb[3] = B.w;
        wc[0] = C.x; wc[1] = C.y; wc[2] = C.z; wc[3] = C.w; wd[0] = D.x; wd[1] = D.y; wd[2] = D.z; wd[3] = D.w;
    }
    const bool slow = (off[0] | off[1] | off[2] | off[3]) < 0;
    // LANES: this thread stores words lane, lane + 32, lane + 64 of the 96-word (128-pixel) row; word j starts in pixel 4j / 3
    int wp[3], wr[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) { const int j = threadIdx.x + 32 * m; wp[m] = (4 * j) / 3; wr[m] = 8 * (4 * j - 3 * wp[m]); }
    const int row_bytes = 3 * min(RM_BX * RM_PX, V.w - tx0);  // valid bytes of this tile row
    const int n = min(RM_PX, V.w - x0);
    uint8_t *dst = V.P + (size_t)p.f0 * V.p_frame_stride + (size_t)y * V.p_pitch + (size_t)(LANES ? tx0 : x0) * 3;
#pragma unroll RM_UNROLL
    for (int f = p.f0; f < p.f0 + p.n_frames; ++f, dst += V.p_frame_stride) {
        const uint8_t *src = p.src[f * p.n_views + vi - p.v0];
#if VSB_K1_PREFETCH
        // the next frame's windows are known now (same table entries, next source buffer): pull their lines towards the SM while
        // this frame is computed -- no registers held, the demand loads of the next iteration hit L2 / L1 instead of HBM
        if (f + 1 < p.f0 + p.n_frames) {
            const uint8_t *nsrc = p.src[(f + 1) * p.n_views + vi - p.v0];
#pragma unroll
            for (int k = 0; k < RM_PX; ++k) {
                const uint8_t *a = nsrc + ((unsigned)off[k] & 0x7ffffffcu);
#if VSB_K1_PREFETCH == 1
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a + p.src_pitch));
#else
                asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(a + p.src_pitch));
#endif
            }
        }
#endif
        unsigned px[RM_PX];
#pragma unroll
        for (int k = 0; k < RM_PX; ++k) px[k] = remap_tab_px<true>(src, p.src_pitch, (unsigned)off[k] & 0x7fffffffu, wa[k], wb[k], wc[k], wd[k], V.gain);
        if (slow) {  // TAP_SLOW entries: the window loads above were redirected to offset 0, the result comes from the coordinates
#pragma unroll 1
            for (int k = 0; k < RM_PX; ++k) {
                if (off[k] >= 0) continue;
                const float fx = __ldg((const float *)((const char *)V.xmap + (size_t)y * V.map_pitch) + x0 + k * XSTEP);
                const float fy = __ldg((const float *)((const char *)V.ymap + (size_t)y * V.map_pitch) + x0 + k * XSTEP);
                px[k] = remap_gain_px_edge<true>(src, p.src_pitch, V.src_w, V.src_h, fx, fy, V.gain);
            }
        }
        if (LANES) {
            unsigned *row = sP[threadIdx.y];
            __syncwarp();  // the previous frame's reads of this row are done
#pragma unroll
            for (int k = 0; k < RM_PX; ++k) row[threadIdx.x + k * XSTEP] = px[k];
            __syncwarp();
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const int b0 = 4 * (threadIdx.x + 32 * m);
                if (b0 >= row_bytes) continue;
                const unsigned word = (row[wp[m]] >> wr[m]) | (row[wp[m] + 1] << (24 - wr[m]));
                if (b0 + 4 <= row_bytes) *(unsigned *)(dst + b0) = word;
                else for (int e = 0; b0 + e < row_bytes; ++e) dst[b0 + e] = (word >> (8 * e)) & 0xff;
            }
        } else if (n == RM_PX) {  // 12 bytes = three aligned 32-bit stores
            unsigned *d32 = (unsigned *)dst;
            d32[0] = px[0] | (px[1] << 24);
            d32[1] = (px[1] >> 8) | (px[2] << 16);
            d32[2] = (px[2] >> 16) | (px[3] << 8);
        } else {
            for (int k = 0; k < n; ++k) { dst[3 * k] = px[k] & 0xff; dst[3 * k + 1] = (px[k] >> 8) & 0xff; dst[3 * k + 2] = (px[k] >> 16) & 0xff; }
        }
    }
}

// K1, higher-occupancy form: the same 128 x 8 tile by 512 threads, TWO lane-interleaved pixels per thread.  A warp owns half a
// tile row (64 pixels = 48 output words): half the table entries and tap values live per thread, 40 registers instead of 64,
// so six instead of four warps per scheduler hide the latency of the window gathers (the kernel's dominant stall).
constexpr int S1H_PX = 2, S1H_SEG = 32 * S1H_PX;  // pixels per thread / per warp
__global__ void __launch_bounds__(RM_BX * RM_BY * 2, 3) k_remap_stage1_tab_h(const __grid_constant__ Stage1TabParams p)
{
    __shared__ unsigned sP[RM_BY * 2][S1H_SEG + 1];
    const unsigned tile = __ldg(p.tiles + blockIdx.x);
    const int vi = tile & 0xff;
    const Stage1TabView &V = p.v[vi];
    const int half = threadIdx.y & 1, y = (int)(tile >> 20) * RM_BY + (threadIdx.y >> 1);
    const int sx0 = (int)((tile >> 8) & 0xfff) * (RM_BX * RM_PX) + half * S1H_SEG;  // first pixel of this warp's half row
    const int x0 = sx0 + threadIdx.x;
    if (y >= V.h || sx0 >= V.w) return;  // (warp-uniform)
    const size_t i = (size_t)y * V.tab.tab_pitch + x0;
    int off[S1H_PX];
    float wa[S1H_PX], wb[S1H_PX], wc[S1H_PX], wd[S1H_PX];
#pragma unroll
    for (int k = 0; k < S1H_PX; ++k) {
        const bool in = x0 + k * 32 < V.tab.tab_pitch;
        off[k] = in ? __ldg(V.tab.off + i + k * 32) : 0;
        wa[k] = in ? __ldg(V.tab.w + i + k * 32) : 0.f;
        wb[k] = in ? __ldg(V.tab.w + V.tab.plane + i + k * 32) : 0.f;
        wc[k] = in ? __ldg(V.tab.w + 2 * V.tab.plane + i + k * 32) : 0.f;
        wd[k] = in ? __ldg(V.tab.w + 3 * V.tab.plane + i + k * 32) : 0.f;
    }
    const bool slow = (off[0] | off[1]) < 0;
    // this thread stores words lane and lane + 32 (< 48) of the half row; word j starts in pixel 4j / 3
    int wp[2], wr[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) { const int j = threadIdx.x + 32 * m; wp[m] = (4 * j) / 3; wr[m] = 8 * (4 * j - 3 * wp[m]); }
    const int row_bytes = 3 * min(S1H_SEG, V.w - sx0);
    uint8_t *dst = V.P + (size_t)p.f0 * V.p_frame_stride + (size_t)y * V.p_pitch + (size_t)sx0 * 3;
    unsigned *row = sP[threadIdx.y];
#pragma unroll 1
    for (int f = p.f0; f < p.f0 + p.n_frames; ++f, dst += V.p_frame_stride) {
        const uint8_t *src = p.src[f * p.n_views + vi - p.v0];
        unsigned px[S1H_PX];
#pragma unroll
        for (int k = 0; k < S1H_PX; ++k) px[k] = remap_tab_px<true>(src, p.src_pitch, (unsigned)off[k] & 0x7fffffffu, wa[k], wb[k], wc[k], wd[k], V.gain);
        if (slow) {
#pragma unroll 1
            for (int k = 0; k < S1H_PX; ++k) {
                if (off[k] >= 0) continue;
                const float fx = __ldg((const float *)((const char *)V.xmap + (size_t)y * V.map_pitch) + x0 + k * 32);
                const float fy = __ldg((const float *)((const char *)V.ymap + (size_t)y * V.map_pitch) + x0 + k * 32);
                px[k] = remap_gain_px_edge<true>(src, p.src_pitch, V.src_w, V.src_h, fx, fy, V.gain);
            }
        }
        __syncwarp();  // the previous frame's reads of this row are done
#pragma unroll
        for (int k = 0; k < S1H_PX; ++k) row[threadIdx.x + k * 32] = px[k];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const int b0 = 4 * (threadIdx.x + 32 * m);
            if (b0 >= row_bytes) continue;
            const unsigned word = (row[wp[m]] >> wr[m]) | (row[wp[m] + 1] << (24 - wr[m]));
            if (b0 + 4 <= row_bytes) *(unsigned *)(dst + b0) = word;
            else for (int e = 0; b0 + e < row_bytes; ++e) dst[b0 + e] = (word >> (8 * e)) & 0xff;
        }
    }
}

// ---- K1, shared-memory staged form (default whenever the caller's frames are 16-byte aligned) ------------------------------------
// The scalar window gathers of the kernels above are bound by the L1 data pipe: a warp's 32 four-byte loads touch ~4 cache lines,
// i.e. ~4 wavefronts per instruction, ~20 per pixel.  Here one CTA owns a 32 x 32 tile of the warped ROI whose SOURCE FOOTPRINT --
// the bounding box of every 2x2 window its table entries address, a static property of the tile computed once per map
// (k_s1_boxes) -- is brought into shared memory with asynchronous 16-byte copies (cp.async: global -> shared without passing
// through registers, perfectly coalesced rows), once per frame of the submission, and the gathers run against shared memory.
// Boxes that fit twice into the buffer (most) are double buffered: the copies of frame f + 1 fly while frame f is computed.
// A thread owns 4 consecutive pixels (one 12-byte store) and keeps its table entries in registers while the frames stream
// through.  Tiles whose footprint exceeds the box (never at the BASELINE configurations) take the global-memory gathers inside
// the same kernel.  (A bulk-tensor copy would need one descriptor per caller frame buffer and a fixed box; cp.async.bulk per
// row is a uniform-datapath instruction and serialises over the lanes that issue it.)
constexpr int S1S_T = 32;                       // tile edge (pixels)
constexpr int S1S_BW = 320, S1S_BH = 104;       // footprint box: bytes per row x rows (33 280 B of shared memory)
struct S1STile { uint32_t id, xw, yh, magic; };  // id = view | tx << 8 | ty << 20; xw = x0 (bytes) | row bytes << 16; yh = y0 | rows << 16 (rows == 0: not staged);
                                                // magic = ceil(2^32 / (row bytes / 16)): chunk index -> box row by one multiply-high

__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// footprint boxes of the 32 x 32 tiles of one view for the current table (one CTA per tile, one thread per 4 entries)
__global__ void __launch_bounds__(256) k_s1_boxes(const uint32_t *__restrict__ ids, S1STile *__restrict__ out, TapTable tab, int w, int h,
                                                  unsigned pitch, int sw, int sh)
{
    __shared__ int mn_x, mx_x, mn_y, mx_y;
    if (threadIdx.x == 0) { mn_x = mn_y = 0x7fffffff; mx_x = mx_y = -1; }
    __syncthreads();
    const uint32_t id = ids[blockIdx.x];
    const int x0 = (int)((id >> 8) & 0xfff) * S1S_T + (threadIdx.x & 7) * 4, y = (int)(id >> 20) * S1S_T + (threadIdx.x >> 3);
    if (y < h) {
        for (int k = 0; k < 4 && x0 + k < w; ++k) {
            const size_t i = (size_t)y * tab.tab_pitch + x0 + k;
            const int off = tab.off[i];
            if (off < 0) continue;  // TAP_SLOW: coordinate path
            if (tab.w[i] == 0.f && tab.w[tab.plane + i] == 0.f && tab.w[2 * tab.plane + i] == 0.f && tab.w[3 * tab.plane + i] == 0.f) continue;
            const int r = (int)((unsigned)off / pitch), xb = (int)((unsigned)off - (unsigned)r * pitch);
            atomicMin(&mn_x, xb); atomicMax(&mx_x, xb); atomicMin(&mn_y, r); atomicMax(&mx_y, r);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        S1STile T;
        T.id = id; T.magic = 0;
        if (mx_x < 0) { T.xw = 16u << 16; T.yh = 1u << 16; }  // nothing addresses the image: a dummy 16-byte box (row 0)
        else {
            int bx0 = mn_x & ~15, bx1 = (mx_x + 10 + 15) & ~15;  // word loads of a window reach 9 bytes past its first byte
            if (bx1 - bx0 < 32) { if ((unsigned)bx0 + 32u <= pitch) bx1 = bx0 + 32; else if (bx0 >= 16) bx0 -= 16; }  // >= 2 chunks per row (multiply-high row index)
            const int bw = bx1 - bx0, bh = mx_y + 2 - mn_y;
            // the last image row must not be read past the end of the caller's buffer
            const bool overrun = mx_y + 1 == sh - 1 && (unsigned)bx1 > (unsigned)sw * 3u;
            const bool fits = bw >= 32 && bw <= S1S_BW && bh <= S1S_BH && (unsigned)bx1 <= pitch && !overrun;
            T.xw = (unsigned)bx0 | ((unsigned)bw << 16);
            T.yh = (unsigned)mn_y | ((fits ? (unsigned)bh : 0u) << 16);
        }
        const unsigned cpr = (T.xw >> 16) / 16u;
        T.magic = cpr <= 1u ? 0u : (unsigned)((0x100000000ull + cpr - 1u) / cpr);  // exact for chunk indices < 2^16 (<= 2080 here); the dummy box has one chunk
        out[blockIdx.x] = T;
    }
}

// remap_tab_px against a shared-memory image (plain loads: the table offset is relative to the staged box)
template <bool GAIN>
__device__ __forceinline__ unsigned remap_tab_px_smem(const uint8_t *base, unsigned pitch, unsigned off, float wa, float wb, float wc, float wd, float gain)
{
    const uint8_t *a = base + (off & ~3u);
    const unsigned s8 = (off & 3u) * 8u;
    const unsigned t0 = *(const unsigned *)a, t1 = *(const unsigned *)(a + 4);
    const unsigned u0 = *(const unsigned *)(a + pitch), u1 = *(const unsigned *)(a + pitch + 4);
    const unsigned t2 = *(const unsigned *)(a + 8), u2 = *(const unsigned *)(a + pitch + 8);  // inside the box by construction
    const unsigned lo1 = __funnelshift_r(t0, t1, s8), hi1 = __funnelshift_r(t1, t2, s8);
    const unsigned lo2 = __funnelshift_r(u0, u1, s8), hi2 = __funnelshift_r(u1, u2, s8);
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = __fmul_rn(u8_den(lo1, c), wa);
        v = __fmaf_rn(c == 0 ? u8_den(lo1, 3) : u8_den(hi1, c - 1), wb, v);
        v = __fmaf_rn(u8_den(lo2, c), wc, v);
        v = __fmaf_rn(c == 0 ? u8_den(lo2, 3) : u8_den(hi2, c - 1), wd, v);
        v = __fmaf_rn(v, 8388608.f, 12582912.f);
        o[c] = GAIN ? rni_biased(fminf(__fmul_rn(gain, __fsub_rn(v, 12582912.f)), 255.f)) : v;
    }
    return __byte_perm(__byte_perm(__float_as_uint(o[0]), __float_as_uint(o[1]), 0x0040u), __float_as_uint(o[2]), 0x0410u) & 0xffffffu;
}

struct Stage1StParams {
    Stage1TabParams t;
    const S1STile *tiles_s;
};

#ifndef VSB_S1S_MINB
#define VSB_S1S_MINB 4
#endif
__global__ void __launch_bounds__(256, VSB_S1S_MINB) k_remap_stage1_st(const __grid_constant__ Stage1StParams pp)
{
    __shared__ __align__(128) uint8_t box[S1S_BW * S1S_BH + 16];
    const Stage1TabParams &p = pp.t;
    const S1STile T = pp.tiles_s[blockIdx.x];
    const int vi = T.id & 0xff;
    const Stage1TabView &V = p.v[vi];
    const int t = threadIdx.x;
    const int x0 = (int)((T.id >> 8) & 0xfff) * S1S_T + (t & 7) * RM_PX, y = (int)(T.id >> 20) * S1S_T + (t >> 3);
    const bool live = y < V.h && x0 < V.w;
    const unsigned bx0 = T.xw & 0xffffu, bw = T.xw >> 16, by0 = T.yh & 0xffffu, bh = T.yh >> 16;
    const bool staged = bh != 0u;  // (uniform)
    const unsigned box_bytes = bw * bh, cpr = bw >> 4, n_chunks = cpr * bh;
    const bool dbl = staged && 2u * box_bytes <= (unsigned)(S1S_BW * S1S_BH);  // (uniform) two half buffers, 128-byte aligned
    const unsigned half = dbl ? ((box_bytes + 127u) & ~127u) : 0u;
    int off[RM_PX];
    float wa[RM_PX], wb[RM_PX], wc[RM_PX], wd[RM_PX];
#pragma unroll
    for (int k = 0; k < RM_PX; ++k) { off[k] = 0; wa[k] = wb[k] = wc[k] = wd[k] = 0.f; }
    if (live) {  // table rows are 16-byte aligned and x0 % 4 == 0; the padding entries of the last vector are zero
        const size_t i = (size_t)y * V.tab.tab_pitch + x0;
        const int4 o = __ldg((const int4 *)(V.tab.off + i));
        const float4 A = __ldg((const float4 *)(V.tab.w + i)), B = __ldg((const float4 *)(V.tab.w + V.tab.plane + i));
        const float4 C = __ldg((const float4 *)(V.tab.w + 2 * V.tab.plane + i)), D = __ldg((const float4 *)(V.tab.w + 3 * V.tab.plane + i));
        off[0] = o.x; off[1] = o.y; off[2] = o.z; off[3] = o.w;
        wa[0] = A.x; wa[1] = A.y; wa[2] = A.z; wa[3] = A.w; wb[0] = B.x; wb[1] = B.y; wb[2] = B.z; wb[3] = B.w;
        wc[0] = C.x; wc[1] = C.y; wc[2] = C.z; wc[3] = C.w; wd[0] = D.x; wd[1] = D.y; wd[2] = D.z; wd[3] = D.w;
    }
    const bool slow = (off[0] | off[1] | off[2] | off[3]) < 0;
    unsigned so[RM_PX];  // staged: offset of the window inside the box; entries without any tap in the image read the box origin (weights 0)
#pragma unroll
    for (int k = 0; k < RM_PX; ++k) {
        so[k] = (unsigned)off[k] & 0x7fffffffu;
        if (staged) {
            const bool null = off[k] < 0 || (wa[k] == 0.f && wb[k] == 0.f && wc[k] == 0.f && wd[k] == 0.f);
            const unsigned r = so[k] / p.src_pitch, xb = so[k] - r * p.src_pitch;
            so[k] = null ? 0u : (r - by0) * bw + (xb - bx0);
        }
    }
    // the box of one frame: 16-byte chunks, chunk c = row c / cpr, column c % cpr (rows of the box are contiguous in the source row)
    auto fetch = [&](const uint8_t *src, uint8_t *dstbox) {
        const uint8_t *g = src + (size_t)by0 * p.src_pitch + bx0;
        for (unsigned c = t; c < n_chunks; c += 256) {
            const unsigned r = __umulhi(c, T.magic), q = c - r * cpr;
            cp_async16(dstbox + c * 16u, g + (size_t)r * p.src_pitch + q * 16u);
        }
        cp_async_commit();
    };
    const int n = min(RM_PX, V.w - x0);
    uint8_t *dst = V.P + (size_t)p.f0 * V.p_frame_stride + (size_t)y * V.p_pitch + (size_t)x0 * 3;
    const int f_end = p.f0 + p.n_frames;
    if (staged) fetch(p.src[p.f0 * p.n_views + vi - p.v0], box);
#pragma unroll 1
    for (int f = p.f0; f < f_end; ++f, dst += V.p_frame_stride) {
        const uint8_t *src = p.src[f * p.n_views + vi - p.v0];
        const uint8_t *cur = box + ((f - p.f0) & 1) * half;
        if (staged) {
            if (dbl && f + 1 < f_end) {  // the next frame's box flies while this one is computed
                fetch(p.src[(f + 1) * p.n_views + vi - p.v0], box + ((f + 1 - p.f0) & 1) * half);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();  // every thread's copies of this frame have landed
        }
        if (live) {
            unsigned px[RM_PX];
            if (staged) {
#pragma unroll
                for (int k = 0; k < RM_PX; ++k) px[k] = remap_tab_px_smem<true>(cur, bw, so[k], wa[k], wb[k], wc[k], wd[k], V.gain);
            } else {
#pragma unroll
                for (int k = 0; k < RM_PX; ++k) px[k] = remap_tab_px<true>(src, p.src_pitch, so[k], wa[k], wb[k], wc[k], wd[k], V.gain);
            }
            if (slow) {  // TAP_SLOW entries: the result comes from the coordinates
#pragma unroll 1
                for (int k = 0; k < RM_PX; ++k) {
                    if (off[k] >= 0) continue;
                    const float fx = __ldg((const float *)((const char *)V.xmap + (size_t)y * V.map_pitch) + x0 + k);
                    const float fy = __ldg((const float *)((const char *)V.ymap + (size_t)y * V.map_pitch) + x0 + k);
                    px[k] = remap_gain_px_edge<true>(src, p.src_pitch, V.src_w, V.src_h, fx, fy, V.gain);
                }
            }
            if (n == RM_PX) {  // 12 bytes = three aligned 32-bit stores
                unsigned *d32 = (unsigned *)dst;
                d32[0] = px[0] | (px[1] << 24);
                d32[1] = (px[1] >> 8) | (px[2] << 16);
                d32[2] = (px[2] >> 16) | (px[3] << 8);
            } else {
                for (int k = 0; k < n; ++k) { dst[3 * k] = px[k] & 0xff; dst[3 * k + 1] = (px[k] >> 8) & 0xff; dst[3 * k + 2] = (px[k] >> 16) & 0xff; }
            }
        }
        if (staged) {
            __syncthreads();  // every reader is done before this buffer is filled again
            if (!dbl && f + 1 < f_end) fetch(p.src[(f + 1) * p.n_views + vi - p.v0], box);
        }
    }
}

// ---- K1 on NV12 frames: cv::cvtColor(CV_YUV2BGR_NV12) (A/networking.cpp:46) inside remap #1's tap fetch ----------
// The table addresses the luma plane; every tap is converted in registers with the integer BT.601 arithmetic of
// YUV420sp2RGB888Invoker<0, 0> (sources/modules/imgproc/src/color.cpp:8741-8818, the same constants as k_nv12_to_bgr) and enters
// the bilinear chain as an exact denormal (the integer's bit pattern IS the float b * 2^-149): no BGR staging image is written or
// read -- 1.5 instead of 3 + 3 bytes of HBM traffic per source pixel.  Lane-interleaved pixels and transposed stores like
// k_remap_stage1_tab<true>.  Bit-exact against nv12_to_bgr followed by the BGR kernels.
struct Nv12Tap { int b, g, r; };
__device__ __forceinline__ Nv12Tap nv12_tap(unsigned y, unsigned uv)
{
    const int u = (int)(uv & 0xffu) - 128, v = (int)(uv >> 8) - 128;
    const int ruv = (1 << 19) + 1673527 * v, guv = (1 << 19) - 852492 * v - 409993 * u, buv = (1 << 19) + 2116026 * u;
    const int yy = max(0, (int)y - 16) * 1220542;
    Nv12Tap t;
    t.b = min(255, max(0, (yy + buv) >> 20)); t.g = min(255, max(0, (yy + guv) >> 20)); t.r = min(255, max(0, (yy + ruv) >> 20));
    return t;
}
__global__ void __launch_bounds__(RM_BX *RM_BY, 4) k_remap_stage1_nv12(const __grid_constant__ Stage1TabParams p, int src_h)
{
    __shared__ unsigned sP[RM_BY][RM_BX * RM_PX + 1];
    const unsigned tile = __ldg(p.tiles + blockIdx.x);
    const int vi = tile & 0xff;
    const Stage1TabView &V = p.v[vi];
    const int tx0 = (int)((tile >> 8) & 0xfff) * (RM_BX * RM_PX), y = (int)(tile >> 20) * RM_BY + threadIdx.y;
    const int x0 = tx0 + threadIdx.x;
    if (y >= V.h) return;
    const size_t i = (size_t)y * V.tab.tab_pitch + x0;
    unsigned offy[RM_PX], offuv[RM_PX];  // luma window; chroma pair of its first tap, bit 0 = window starts on an odd column, bit 31 of offy = odd row
    float wa[RM_PX], wb[RM_PX], wc[RM_PX], wd[RM_PX];
#pragma unroll
    for (int k = 0; k < RM_PX; ++k) {
        const bool in = x0 + k * RM_BX < V.tab.tab_pitch;
        const unsigned o = in ? (unsigned)__ldg(V.tab.off + i + k * RM_BX) : 0u;
        wa[k] = in ? __ldg(V.tab.w + i + k * RM_BX) : 0.f;
        wb[k] = in ? __ldg(V.tab.w + V.tab.plane + i + k * RM_BX) : 0.f;
        wc[k] = in ? __ldg(V.tab.w + 2 * V.tab.plane + i + k * RM_BX) : 0.f;
        wd[k] = in ? __ldg(V.tab.w + 3 * V.tab.plane + i + k * RM_BX) : 0.f;
        const unsigned ys = o / p.src_pitch, xs = o - ys * p.src_pitch;
        offy[k] = o | ((ys & 1u) << 31);
        offuv[k] = ((unsigned)src_h + (ys >> 1)) * p.src_pitch + (xs & ~1u) + (xs & 1u);
    }
    int wp[3], wr[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) { const int j = threadIdx.x + 32 * m; wp[m] = (4 * j) / 3; wr[m] = 8 * (4 * j - 3 * wp[m]); }
    const int row_bytes = 3 * min(RM_BX * RM_PX, V.w - tx0);
    uint8_t *dst = V.P + (size_t)p.f0 * V.p_frame_stride + (size_t)y * V.p_pitch + (size_t)tx0 * 3;
#pragma unroll 1
    for (int f = p.f0; f < p.f0 + p.n_frames; ++f, dst += V.p_frame_stride) {
        const uint8_t *src = p.src[f * p.n_views + vi - p.v0];
        unsigned px[RM_PX];
#pragma unroll
        for (int k = 0; k < RM_PX; ++k) {
            const uint8_t *py = src + (offy[k] & 0x7fffffffu);
            const unsigned dx2 = (offuv[k] & 1u) * 2u, dyp = (offy[k] >> 31) * p.src_pitch;
            const uint8_t *puv = src + (offuv[k] & ~1u);
            const unsigned y00 = __ldg(py), y01 = __ldg(py + 1), y10 = __ldg(py + p.src_pitch), y11 = __ldg(py + p.src_pitch + 1);
            const unsigned uv00 = __ldg((const unsigned short *)puv), uv01 = __ldg((const unsigned short *)(puv + dx2));
            const unsigned uv10 = __ldg((const unsigned short *)(puv + dyp)), uv11 = __ldg((const unsigned short *)(puv + dyp + dx2));
            const Nv12Tap a = nv12_tap(y00, uv00), b = nv12_tap(y01, uv01), c = nv12_tap(y10, uv10), d = nv12_tap(y11, uv11);
            float o[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const int ta = ch == 0 ? a.b : (ch == 1 ? a.g : a.r), tb = ch == 0 ? b.b : (ch == 1 ? b.g : b.r);
                const int tc = ch == 0 ? c.b : (ch == 1 ? c.g : c.r), td = ch == 0 ? d.b : (ch == 1 ? d.g : d.r);
                float v = __fmul_rn(__int_as_float(ta), wa[k]);
                v = __fmaf_rn(__int_as_float(tb), wb[k], v);
                v = __fmaf_rn(__int_as_float(tc), wc[k], v);
                v = __fmaf_rn(__int_as_float(td), wd[k], v);
                v = __fmaf_rn(v, 8388608.f, 12582912.f);
                o[ch] = rni_biased(fminf(__fmul_rn(V.gain, __fsub_rn(v, 12582912.f)), 255.f));
            }
            px[k] = __byte_perm(__byte_perm(__float_as_uint(o[0]), __float_as_uint(o[1]), 0x0040u), __float_as_uint(o[2]), 0x0410u) & 0xffffffu;
        }
        unsigned *row = sP[threadIdx.y];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < RM_PX; ++k) row[threadIdx.x + k * RM_BX] = px[k];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            const int b0 = 4 * (threadIdx.x + 32 * m);
            if (b0 >= row_bytes) continue;
            const unsigned word = (row[wp[m]] >> wr[m]) | (row[wp[m] + 1] << (24 - wr[m]));
            if (b0 + 4 <= row_bytes) *(unsigned *)(dst + b0) = word;
            else for (int e = 0; b0 + e < row_bytes; ++e) dst[b0 + e] = (word >> (8 * e)) & 0xff;
        }
    }
}

struct Stage2TabView {
    TapTable tab;
    const uint8_t *Pbase;        // frame 0 of the zero-framed P allocation (table offsets are relative to it)
    uint8_t *G0;
    size_t p_frame_stride, g0_frame_stride;
    unsigned p_pitch;
    int bw, bh;
};
struct Stage2TabParams {
    const uint32_t *tiles;
    Stage2TabView v[MAXV];
    int n_frames, f0;
};

__global__ void __launch_bounds__(RM_BX *RM_BY) k_remap_stage2_tab(const __grid_constant__ Stage2TabParams p)
{
    const unsigned tile = __ldg(p.tiles + blockIdx.x);
    const Stage2TabView &V = p.v[tile & 0xff];
    const int bx0 = ((int)((tile >> 8) & 0xfff) * RM_BX + threadIdx.x) * RM_PX, by = (int)(tile >> 20) * RM_BY + threadIdx.y;
    if (bx0 >= V.bw || by >= V.bh) return;
    const size_t i = (size_t)by * V.tab.tab_pitch + bx0;
    const int4 o = __ldg((const int4 *)(V.tab.off + i));
    const float4 A = __ldg((const float4 *)(V.tab.w + i)), B = __ldg((const float4 *)(V.tab.w + V.tab.plane + i));
    const float4 C = __ldg((const float4 *)(V.tab.w + 2 * V.tab.plane + i)), D = __ldg((const float4 *)(V.tab.w + 3 * V.tab.plane + i));
    const size_t plane = (size_t)V.bw * V.bh;
    const int n = min(RM_PX, V.bw - bx0);
    const bool vec = n == RM_PX && (V.bw & 3) == 0;
    const uint8_t *P = V.Pbase + (size_t)p.f0 * V.p_frame_stride;
    uint8_t *g = V.G0 + (size_t)p.f0 * V.g0_frame_stride + (size_t)by * V.bw + bx0;
#pragma unroll RM_UNROLL
    for (int f = 0; f < p.n_frames; ++f, P += V.p_frame_stride, g += V.g0_frame_stride) {
        unsigned px[RM_PX];
        px[0] = remap_tab_px<false>(P, V.p_pitch, (unsigned)o.x, A.x, B.x, C.x, D.x, 1.f);
        px[1] = remap_tab_px<false>(P, V.p_pitch, (unsigned)o.y, A.y, B.y, C.y, D.y, 1.f);
        px[2] = remap_tab_px<false>(P, V.p_pitch, (unsigned)o.z, A.z, B.z, C.z, D.z, 1.f);
        px[3] = remap_tab_px<false>(P, V.p_pitch, (unsigned)o.w, A.w, B.w, C.w, D.w, 1.f);
        // interleaved -> planar: byte c of the four pixels
        const unsigned lo01 = __byte_perm(px[0], px[1], 0x5140u), lo23 = __byte_perm(px[2], px[3], 0x5140u);
        const unsigned c0 = __byte_perm(lo01, lo23, 0x5410u), c1 = __byte_perm(lo01, lo23, 0x7632u);
        const unsigned c2 = __byte_perm(__byte_perm(px[0], px[1], 0x0062u), __byte_perm(px[2], px[3], 0x0062u), 0x5410u);
        if (vec) {
            *(unsigned *)g = c0;
            *(unsigned *)(g + plane) = c1;
            *(unsigned *)(g + 2 * plane) = c2;
        } else {
            for (int k = 0; k < n; ++k) { g[k] = (c0 >> (8 * k)) & 0xff; g[plane + k] = (c1 >> (8 * k)) & 0xff; g[2 * plane + k] = (c2 >> (8 * k)) & 0xff; }
        }
    }
}

// ---- K0 (optional): NV12 wire format -> BGR ------------------------------------------------------------------------------
// The capture boards send NV12 and the reference converts every received frame on the CPU with
// cv::cvtColor(mat, mat, CV_YUV2BGR_NV12) before the upload (360_stitcher/networking.cpp:46, A/defs.h:10-17).  Here the NV12
// frame is what crosses PCIe (half the bytes) and this kernel restates the integer BT.601 arithmetic of
// compose_scale != 1: cuda::resize(full_img, img, Size(), compose_scale, compose_scale, INTER_LINEAR) of every camera frame of a
// submission in one launch (A/timed.cpp:74-77; kernel sources/modules/cudawarping/src/cuda/resize.cu:71-106, arithmetic in
// resize_linear_px).  blockIdx.z = frame * views + view; entries of views this rank does not own are null and skipped.
struct PrescaleParams {
    const uint8_t *src[MAX_BATCH * MAXV];
    uint8_t *dst[MAX_BATCH * MAXV];
    size_t pitch, dst_pitch;
    int sw, sh, dw, dh;
    float fx, fy;   // static_cast<float>(1.0 / compose_scale), as the host wrapper passes it (src/resize.cpp:104)
};

__global__ void __launch_bounds__(256) k_prescale(const __grid_constant__ PrescaleParams p)
{
    const int dx = blockIdx.x * 32 + threadIdx.x, dy = blockIdx.y * 8 + threadIdx.y;
    const uint8_t *src = p.src[blockIdx.z];
    if (dx >= p.dw || dy >= p.dh || !src) return;
    resize_linear_px<3>(src, p.sw, p.sh, p.pitch, p.dst[blockIdx.z], p.dst_pitch, dx, dy, p.fx, p.fy);
}

// YUV420sp2RGB888Invoker<bIdx = 0, uIdx = 0> (sources/modules/imgproc/src/color.cpp:8741-8746, 8793-8818) bit for bit.
// One thread = 4 pixels x 2 rows (two chroma pairs): three 32-bit loads, six 32-bit stores.
struct Nv12Params {
    const uint8_t *src[MAX_BATCH * MAXV];  // NV12 frames: h rows of Y, then h / 2 rows of interleaved U, V; rows `pitch` bytes apart
    uint8_t *dst[MAX_BATCH * MAXV];        // BGR staging image of each frame
    size_t pitch, dst_pitch;
    int w, h;
};

__device__ __forceinline__ unsigned nv12_px(int y, int ruv, int guv, int buv)
{
    const int yy = max(0, y - 16) * 1220542;
    const int b = min(255, max(0, (yy + buv) >> 20)), g = min(255, max(0, (yy + guv) >> 20)), r = min(255, max(0, (yy + ruv) >> 20));
    return (unsigned)b | ((unsigned)g << 8) | ((unsigned)r << 16);
}

__global__ void __launch_bounds__(256) k_nv12_to_bgr(const __grid_constant__ Nv12Params p)
{
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4, y0 = (blockIdx.y * 8 + threadIdx.y) * 2;
    if (x0 >= p.w || y0 >= p.h) return;
    const uint8_t *src = p.src[blockIdx.z];
    if (!src) return;  // a view another rank owns
    const uint8_t *y1 = src + (size_t)y0 * p.pitch + x0, *uv = src + (size_t)p.h * p.pitch + (size_t)(y0 >> 1) * p.pitch + x0;
    uint8_t *d = p.dst[blockIdx.z] + (size_t)y0 * p.dst_pitch + (size_t)x0 * 3;
    const int n = min(4, p.w - x0);  // w is even: n is 2 or 4
    unsigned ya, yb, c;
    if (n == 4 && ((((size_t)src) | p.pitch) & 3) == 0) {
        ya = __ldg((const unsigned *)y1); yb = __ldg((const unsigned *)(y1 + p.pitch)); c = __ldg((const unsigned *)uv);
    } else {
        ya = yb = c = 0;
        for (int i = 0; i < n; ++i) { ya |= (unsigned)__ldg(y1 + i) << (8 * i); yb |= (unsigned)__ldg(y1 + p.pitch + i) << (8 * i); c |= (unsigned)__ldg(uv + i) << (8 * i); }
    }
    unsigned pa[4], pb[4];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int u = (int)((c >> (16 * k)) & 0xff) - 128, v = (int)((c >> (16 * k + 8)) & 0xff) - 128;
        const int ruv = (1 << 19) + 1673527 * v, guv = (1 << 19) - 852492 * v - 409993 * u, buv = (1 << 19) + 2116026 * u;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            pa[2 * k + q] = nv12_px((int)((ya >> (8 * (2 * k + q))) & 0xff), ruv, guv, buv);
            pb[2 * k + q] = nv12_px((int)((yb >> (8 * (2 * k + q))) & 0xff), ruv, guv, buv);
        }
    }
    if (n == 4 && ((((size_t)p.dst[blockIdx.z]) | p.dst_pitch) & 3) == 0) {  // 12 bytes per row = three aligned words
        unsigned *d0 = (unsigned *)d, *d1 = (unsigned *)(d + p.dst_pitch);
        d0[0] = pa[0] | (pa[1] << 24); d0[1] = (pa[1] >> 8) | (pa[2] << 16); d0[2] = (pa[2] >> 16) | (pa[3] << 8);
        d1[0] = pb[0] | (pb[1] << 24); d1[1] = (pb[1] >> 8) | (pb[2] << 16); d1[2] = (pb[2] >> 16) | (pb[3] << 8);
    } else {
        for (int i = 0; i < n; ++i) {
            d[3 * i] = pa[i] & 0xff; d[3 * i + 1] = (pa[i] >> 8) & 0xff; d[3 * i + 2] = (pa[i] >> 16) & 0xff;
            uint8_t *e = d + p.dst_pitch;
            e[3 * i] = pb[i] & 0xff; e[3 * i + 1] = (pb[i] >> 8) & 0xff; e[3 * i + 2] = (pb[i] >> 16) & 0xff;
        }
    }
}

// ---- consumer epilogue: what the reference's consumer thread does on the CPU after the download --------
// cv::resize(original_8u, resized_bgr, Size(OUTPUT_WIDTH, image_height), 0, 0, INTER_LINEAR) in 11-bit fixed point
// (sources/modules/imgproc/src/resize.cpp:3930-4021 tables, :1923-1941 horizontal pass, :2013 vertical pass), then either
// COLOR_BGR2RGB (360_stitcher/timed.cpp:291) or the letter-boxed frame through COLOR_BGR2YUV_I420 (:283-289, :310-315,
// sources/modules/imgproc/src/color.cpp:9137-9157).  Integer work: bit-exact.
struct ConsumeParams {
    const uint8_t *src;   // CV_8UC3 panorama
    size_t src_pitch;
    int sw, sh;
    const int *xofs, *yofs;          // source column / row of every destination column / row (resize tables)
    const int *ia, *ib;              // the two 11-bit coefficients of every column / row, packed lo | hi << 16
    int out_w, out_h, ih, row0;      // output frame, resized image height, first frame row of the image
    uint8_t *dst;
    size_t dst_pitch;
};

__device__ __forceinline__ unsigned consume_px(const ConsumeParams &p, int dx, int dy)  // resized pixel as b | g << 8 | r << 16
{
    const int sy = __ldg(p.yofs + dy), cb = __ldg(p.ib + dy), sx = __ldg(p.xofs + dx), ca = __ldg(p.ia + dx);
    const int b0 = (short)(cb & 0xffff), b1 = cb >> 16, a0 = (short)(ca & 0xffff), a1 = ca >> 16;
    const uint8_t *r0 = p.src + (size_t)min(max(sy, 0), p.sh - 1) * p.src_pitch, *r1 = p.src + (size_t)min(max(sy + 1, 0), p.sh - 1) * p.src_pitch;
    const int x0 = sx * 3, x1 = min(sx + 1, p.sw - 1) * 3;
    unsigned out = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int S0 = (int)__ldg(r0 + x0 + c) * a0 + (int)__ldg(r0 + x1 + c) * a1;
        const int S1 = (int)__ldg(r1 + x0 + c) * a0 + (int)__ldg(r1 + x1 + c) * a1;
        out |= (unsigned)((((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2 & 0xff) << (8 * c);
    }
    return out;
}

__global__ void __launch_bounds__(256) k_consume_rgb(const __grid_constant__ ConsumeParams p)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= p.out_w || y >= p.ih) return;
    const unsigned v = consume_px(p, x, y);
    uint8_t *d = p.dst + (size_t)y * p.dst_pitch + (size_t)x * 3;
    d[0] = (v >> 16) & 0xff; d[1] = (v >> 8) & 0xff; d[2] = v & 0xff;  // COLOR_BGR2RGB
}

__device__ __forceinline__ int sat_u8i(int v) { return min(255, max(0, v)); }

// one thread = one 2 x 2 block of the out_w x out_h frame: four Y samples, one U and one V (from the block's top-left pixel)
__global__ void __launch_bounds__(256) k_consume_i420(const __grid_constant__ ConsumeParams p)
{
    const int x = (blockIdx.x * 32 + threadIdx.x) * 2, y = (blockIdx.y * 8 + threadIdx.y) * 2;
    if (x >= p.out_w || y >= p.out_h) return;
    uint8_t *yp = p.dst, *up = p.dst + (size_t)p.out_w * p.out_h, *vp = up + (size_t)(p.out_w / 2) * (p.out_h / 2);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int iy = y + j - p.row0;  // row of the resized image (the rest of the frame is black)
        unsigned yy = 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const unsigned v = (unsigned)iy < (unsigned)p.ih ? consume_px(p, x + i, iy) : 0u;
            const int b = v & 0xff, g = (v >> 8) & 0xff, r = (v >> 16) & 0xff;
            yy |= (unsigned)sat_u8i((269484 * r + 528482 * g + 102760 * b + (1 << 19) + (16 << 20)) >> 20) << (8 * i);
            if (i == 0 && j == 0) {
                up[(size_t)(y / 2) * (p.out_w / 2) + x / 2] = (uint8_t)sat_u8i((-155188 * r - 305135 * g + 460324 * b + (1 << 19) + (128 << 20)) >> 20);
                vp[(size_t)(y / 2) * (p.out_w / 2) + x / 2] = (uint8_t)sat_u8i((460324 * r - 385875 * g - 74448 * b + (1 << 19) + (128 << 20)) >> 20);
            }
        }
        *(uint16_t *)(yp + (size_t)(y + j) * p.out_w + x) = (uint16_t)yy;
    }
}

// ---- view-sharded mode: pack / unpack of the Gaussian sub-planes one peer reads (one message per peer and submission) ---------
struct ShardRect {
    uint8_t *plane;        // plane 0 of frame slot 0 of that Gaussian level: [3][ph][pw] u8
    size_t frame_stride;   // bytes between frame slots
    size_t off;            // byte offset of this rectangle inside one frame's packed block
    int pw, ph, x0, y0, w, h;
};
// block = 32 x 8 threads: rows over threadIdx.y, 32-bit words (rectangles are widened to word boundaries at plan time whenever
// the plane rows are word aligned) or bytes over threadIdx.x
template <bool PACK>
__global__ void __launch_bounds__(256) k_shard_copy(const ShardRect *__restrict__ tab, uint8_t *__restrict__ buf, size_t frame_bytes, int f0)
{
    const ShardRect R = tab[blockIdx.x];
    const int c = blockIdx.y, f = blockIdx.z;  // f counts frames of the submission (packed buffer); the planes start at frame slot f0
    uint8_t *pl = R.plane + (size_t)(f0 + f) * R.frame_stride + ((size_t)c * R.ph + R.y0) * R.pw + R.x0;
    uint8_t *pk = buf + (size_t)f * frame_bytes + R.off + (size_t)c * R.w * R.h;
    const bool words = ((R.pw | R.x0 | R.w) & 3) == 0;
    for (int r = threadIdx.y; r < R.h; r += 8) {
        uint8_t *a = pl + (size_t)r * R.pw, *b = pk + (size_t)r * R.w;
        if (words) {
            for (int q = threadIdx.x; q < (R.w >> 2); q += 32) {
                if (PACK) ((unsigned *)b)[q] = ((const unsigned *)a)[q];
                else ((unsigned *)a)[q] = ((const unsigned *)b)[q];
            }
        } else {
            for (int q = threadIdx.x; q < R.w; q += 32) {
                if (PACK) b[q] = a[q];
                else a[q] = b[q];
            }
        }
    }
}

// ---- K3: pyrDown on planes ----------------------------------------------------------------------------
// out(y,x) = rhe( sum_{j,i} k5[j] k5[i] in(r101(2y+j-2), r101(2x+i-2)) / 256 ): the exact integer form of the
// reference's fp32 vertical-then-horizontal 5-tap passes (every partial sum is exactly representable).
struct PyrView {
    const void *in;   // plane 0 of frame 0 (u8 when level 0, else s16)
    int16_t *out;
    size_t in_frame_stride, out_frame_stride;  // in elements
    int w, h;                                  // input plane size
};
struct PyrParams {
    TileMap tm;
    PyrView v[MAXV];
};

constexpr int PD_TX = 32, PD_TY = 8;  // output tile per block (one thread per output sample)

template <typename TIn>
__global__ void __launch_bounds__(PD_TX *PD_TY) k_pyr_down(const __grid_constant__ PyrParams p)
{
    __shared__ int16_t tile[2 * PD_TY + 3][2 * PD_TX + 4];
    const int vi = find_view(p.tm, blockIdx.x);
    const PyrView &V = p.v[vi];
    const int t = blockIdx.x - p.tm.start[vi];
    const int tx = t % p.tm.tiles_x[vi], ty = t / p.tm.tiles_x[vi];
    const int f = blockIdx.y, c = blockIdx.z;
    const int w = V.w, h = V.h, ow = (w + 1) >> 1, oh = (h + 1) >> 1;
    const TIn *in = (const TIn *)V.in + (size_t)f * V.in_frame_stride + (size_t)c * w * h;
    const int ox0 = tx * PD_TX, oy0 = ty * PD_TY;
    const int ix0 = 2 * ox0 - 2, iy0 = 2 * oy0 - 2;
    const int tid = threadIdx.y * PD_TX + threadIdx.x;
    for (int i = tid; i < (2 * PD_TY + 3) * (2 * PD_TX + 3); i += PD_TX * PD_TY) {
        const int ly = i / (2 * PD_TX + 3), lx = i - ly * (2 * PD_TX + 3);
        const int sy = r101_idx(iy0 + ly, h), sx = r101_idx(ix0 + lx, w);
        tile[ly][lx] = (int16_t)in[(size_t)sy * w + sx];
    }
    __syncthreads();
    const int ox = ox0 + threadIdx.x, oy = oy0 + threadIdx.y;
    if (ox >= ow || oy >= oh) return;
    int acc = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int16_t *r = &tile[2 * threadIdx.y + j][2 * threadIdx.x];
        const int rs = r[0] + 4 * r[1] + 6 * r[2] + 4 * r[3] + r[4];
        const int kj = (j == 0 || j == 4) ? 1 : ((j == 2) ? 6 : 4);
        acc += kj * rs;
    }
    V.out[(size_t)f * V.out_frame_stride + (size_t)c * ow * oh + (size_t)oy * ow + ox] = (int16_t)sat_s16(rhe_shift<8>(acc));
}

// ---- K4: fused Laplacian + weighted add + normalise + collapse + mask + crop ---------------------------
struct LgBlendView {
    const uint8_t *g0;       // level 0 planes (u8), frame 0
    const int16_t *g[MAXL];  // level k >= 1 planes (s16), frame 0
    const float *w[MAXL];    // static weight pyramid
    size_t g0_frame_stride;
    size_t g_frame_stride[MAXL];
    int x_tl, y_tl, bw, bh;  // level-0 canvas rect origin and bordered size (multiples of 2^nb)
};
struct LgBlendPlan {
    int n_views, nb;
    int cw[MAXL], ch[MAXL];   // canvas (padded dst roi) size per level
    const float *dw[MAXL];    // static sum of weights per level (accumulated in view order)
    int out_w, out_h;         // dst_roi_final_
    LgBlendView v[MAXV];
};

constexpr int LG_TW = 64, LG_TH = 32, LG_THREADS = 256;
// region of level k needed by a LG_TW x LG_TH tile: r(k) = r(k-1)/2 + 3 (upper bound)
constexpr int LG_R1W = LG_TW / 2 + 2, LG_R1H = LG_TH / 2 + 2;
constexpr int LG_SMEM_ELEMS = LG_R1W * LG_R1H;  // largest stored region (level 1)

struct GlobalS16Plane {
    const int16_t *p;
    int w;
    __device__ __forceinline__ int operator()(int x, int y) const { return (int)__ldg(p + (size_t)y * w + x); }
};
struct SmemRegion {
    const int16_t *p;
    int x0, y0, w;
    __device__ __forceinline__ int operator()(int x, int y) const { return (int)p[(y - y0) * w + (x - x0)]; }
};

__global__ void __launch_bounds__(LG_THREADS) k_legacy_blend_collapse(const LgBlendPlan *__restrict__ plan, const __grid_constant__ OutPtrs outs, size_t out_pitch)
{
    __shared__ int16_t sD[2][3][LG_SMEM_ELEMS];
    __shared__ LgBlendPlan P;
    for (int i = threadIdx.x; i < (int)(sizeof(LgBlendPlan) / 4); i += LG_THREADS) ((int *)&P)[i] = ((const int *)plan)[i];
    __syncthreads();
    const int nb = P.nb, f = blockIdx.z;
    int lo_x[MAXL], hi_x[MAXL], lo_y[MAXL], hi_y[MAXL];
    lo_x[0] = blockIdx.x * LG_TW; hi_x[0] = min(lo_x[0] + LG_TW, P.cw[0]) - 1;
    lo_y[0] = blockIdx.y * LG_TH; hi_y[0] = min(lo_y[0] + LG_TH, P.ch[0]) - 1;
#pragma unroll
    for (int k = 1; k < MAXL; ++k) {
        if (k <= nb) {
            lo_x[k] = max(0, (lo_x[k - 1] >> 1) - 1); hi_x[k] = min(P.cw[k] - 1, (hi_x[k - 1] >> 1) + 1);
            lo_y[k] = max(0, (lo_y[k - 1] >> 1) - 1); hi_y[k] = min(P.ch[k] - 1, (hi_y[k - 1] >> 1) + 1);
        }
    }
#pragma unroll 1
    for (int k = nb; k >= 0; --k) {
        const int rw = hi_x[k] - lo_x[k] + 1, rh = hi_y[k] - lo_y[k] + 1;
        const int cwk = P.cw[k];
        int16_t(*cur)[LG_SMEM_ELEMS] = sD[k & 1];
        const int16_t(*prev)[LG_SMEM_ELEMS] = sD[(k + 1) & 1];
#pragma unroll 1
        for (int idx = threadIdx.x; idx < rw * rh; idx += LG_THREADS) {
            const int ry = idx / rw, rx = idx - ry * rw;
            const int px = lo_x[k] + rx, py = lo_y[k] + ry;
            int a0 = 0, a1 = 0, a2 = 0;
#pragma unroll 1
            for (int vi = 0; vi < P.n_views; ++vi) {
                const LgBlendView &V = P.v[vi];
                const int bwk = V.bw >> k, bhk = V.bh >> k;
                const int qx = px - (V.x_tl >> k), qy = py - (V.y_tl >> k);
                if ((unsigned)qx >= (unsigned)bwk || (unsigned)qy >= (unsigned)bhk) continue;
                const float wv = __ldg(V.w[k] + (size_t)qy * bwk + qx);
                if (wv == 0.f) continue;  // (short)(L * 0) == 0 exactly: nothing to add
                const size_t plane = (size_t)bwk * bhk;
                int g[3];
                if (k == 0) {
                    const uint8_t *s = V.g0 + (size_t)f * V.g0_frame_stride + (size_t)qy * bwk + qx;
                    g[0] = __ldg(s); g[1] = __ldg(s + plane); g[2] = __ldg(s + 2 * plane);
                } else {
                    const int16_t *s = V.g[k] + (size_t)f * V.g_frame_stride[k] + (size_t)qy * bwk + qx;
                    g[0] = __ldg(s); g[1] = __ldg(s + plane); g[2] = __ldg(s + 2 * plane);
                }
                if (k < nb) {  // Laplacian: G_k - pyrUp(G_{k+1}), saturating
                    const int nx = bwk >> 1, ny = bhk >> 1;
                    const int16_t *up = V.g[k + 1] + (size_t)f * V.g_frame_stride[k + 1];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        GlobalS16Plane acc{up + (size_t)c * nx * ny, nx};
                        g[c] = sat_s16(g[c] - pyr_up_sample(acc, qx, qy, nx, ny));
                    }
                }
                a0 += rz_s16(__fmul_rn((float)g[0], wv));
                a1 += rz_s16(__fmul_rn((float)g[1], wv));
                a2 += rz_s16(__fmul_rn((float)g[2], wv));
            }
            // normalise (truncating) by the static weight sum
            const float dwv = __ldg(P.dw[k] + (size_t)py * cwk + px);
            const float den = __fadd_rn(dwv, 1e-5f);
            int d[3] = {(int)(short)a0, (int)(short)a1, (int)(short)a2};
#pragma unroll
            for (int c = 0; c < 3; ++c) d[c] = rz_s16(__fdiv_rn((float)d[c], den));
            if (k < nb) {  // collapse: D_k += pyrUp(D_{k+1}), saturating
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    SmemRegion acc{prev[c], lo_x[k + 1], lo_y[k + 1], hi_x[k + 1] - lo_x[k + 1] + 1};
                    d[c] = sat_s16(d[c] + pyr_up_sample(acc, px, py, P.cw[k + 1], P.ch[k + 1]));
                }
            }
            if (k > 0) {
                cur[0][idx] = (int16_t)d[0]; cur[1][idx] = (int16_t)d[1]; cur[2][idx] = (int16_t)d[2];
            } else if (px < P.out_w && py < P.out_h) {
                const bool m = dwv > 1e-5f;  // dst_mask = dst_band_weights_[0] > WEIGHT_EPS
                int16_t *o = (int16_t *)((char *)outs.out[f] + (size_t)py * out_pitch) + (size_t)px * 3;
                o[0] = m ? (int16_t)d[0] : (int16_t)0;
                o[1] = m ? (int16_t)d[1] : (int16_t)0;
                o[2] = m ? (int16_t)d[2] : (int16_t)0;
            }
        }
        __syncthreads();
    }
}

// ---- static setup kernels -------------------------------------------------------------------------------
// weight level 0: copyMakeBorder(mask * (1/255), BORDER_CONSTANT 0) (sources/modules/stitching/src/blenders.cpp:410-421)
__global__ void k_weight_level0(const uint8_t *__restrict__ mask, int mw, int mh, size_t mp, int top, int left, float *__restrict__ w0,
                                uint8_t *__restrict__ m0, int bw, int bh)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= bw || y >= bh) return;
    const int sx = x - left, sy = y - top;
    uint8_t m = 0;
    if ((unsigned)sx < (unsigned)mw && (unsigned)sy < (unsigned)mh) m = mask[(size_t)sy * mp + sx];
    w0[(size_t)y * bw + x] = m ? __fmul_rn((float)(1. / 255.), (float)m) : 0.f;
    m0[(size_t)y * bw + x] = m;  // k_blend recomputes W0 from this byte (same product)
}

// dst_band_weights_[k](rc) += weight (the `dst_weight += w` half of addSrcWeightKernel32F), one view at a time
__global__ void k_accum_weight(const float *__restrict__ w, int bw, int bh, float *dw, int cw, int x_tl, int y_tl)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= bw || y >= bh) return;
    float *d = dw + (size_t)(y_tl + y) * cw + (x_tl + x);
    *d = __fadd_rn(*d, w[(size_t)y * bw + x]);
}

// interleave / de-interleave helpers for vsb_debug_read
__global__ void k_planar_to_interleaved_s16(const void *__restrict__ in, int is_u8, int w, int h, int16_t *__restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t plane = (size_t)w * h, o = (size_t)y * w + x;
    for (int c = 0; c < 3; ++c)
        out[o * 3 + c] = is_u8 ? (int16_t)((const uint8_t *)in)[c * plane + o] : ((const int16_t *)in)[c * plane + o];
}

// ---- CPW mesh -> backward map (360_stitcher/meshwarper.cpp:823-886) on the device ----------------------
// m1+m2: forward-splat every pixel of the warped view into the half-resolution accumulators.  The sums are of
// integer pixel coordinates (< 2^24) so fp32 atomics reproduce the CPU loop bit-exactly in any order.
__global__ void k_mesh_splat(const float *__restrict__ mesh_x, const float *__restrict__ mesh_y, int mrows, int mcols,
                             int W, int H, float *sum_x, float *sum_y, float *cnt)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float bx = custom_resize_at(mesh_x, mcols, mrows, (size_t)mcols * 4, W, H, x, y);
    const float by = custom_resize_at(mesh_y, mcols, mrows, (size_t)mcols * 4, W, H, x, y);
    if (!(fabsf(bx) < 2147483648.f) || !(fabsf(by) < 2147483648.f)) return;  // (int) of NaN / out of range fails the test below
    const int x_ = (int)bx / 2, y_ = (int)by / 2;  // C cast (toward zero) then integer divide
    const int hw = W / 2, hh = H / 2;
    if (x_ >= 0 && y_ >= 0 && x_ < hw && y_ < hh) {
        atomicAdd(sum_x + (size_t)y_ * hw + x_, (float)x);
        atomicAdd(sum_y + (size_t)y_ * hw + x_, (float)y);
        atomicAdd(cnt + (size_t)y_ * hw + x_, 1.f);
    }
}

// m3: warp = sum / count (0/0 = NaN for empty cells, as in the reference)
__global__ void k_mesh_divide(float *sum_x, float *sum_y, const float *__restrict__ cnt, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sum_x[i] = __fdiv_rn(sum_x[i], cnt[i]);
    sum_y[i] = __fdiv_rn(sum_y[i], cnt[i]);
}

// m4: custom_resize of the half table back to W x H
__global__ void k_mesh_upsample(const float *__restrict__ wx, const float *__restrict__ wy, int hw, int hh, int W, int H,
                                float *mx, float *my, size_t pitch)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u >= W || v >= H) return;
    *((float *)((char *)mx + (size_t)v * pitch) + u) = custom_resize_at(wx, hw, hh, (size_t)hw * 4, W, H, u, v);
    *((float *)((char *)my + (size_t)v * pitch) + u) = custom_resize_at(wy, hw, hh, (size_t)hw * 4, W, H, u, v);
}

// m4 for a column window of the view (split calibration): columns [x0, x0 + w) of the W x H map, x coordinates re-based to the
// window.  x - x0 is exact wherever the map points into or near the window (both are multiples of ulp(x) and the difference is the
// smaller number), so floor and fraction -- all the remap uses -- are those of the full-width map; far outside it only has to stay outside.
__global__ void k_mesh_upsample_win(const float *__restrict__ wx, const float *__restrict__ wy, int hw, int hh, int W, int H, int x0, int w,
                                    float *mx, float *my, size_t pitch)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u >= w || v >= H) return;
    *((float *)((char *)mx + (size_t)v * pitch) + u) = __fsub_rn(custom_resize_at(wx, hw, hh, (size_t)hw * 4, W, H, u + x0, v), (float)x0);
    *((float *)((char *)my + (size_t)v * pitch) + u) = custom_resize_at(wy, hw, hh, (size_t)hw * 4, W, H, u + x0, v);
}

// ============================================================================================ host side

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline dim3 grid2d(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y); }

struct View {
    bool inited = false, has_maps = false;
    int roi_w = 0, roi_h = 0, tl_x = 0, tl_y = 0;
    int top = 0, bottom = 0, left = 0, right = 0, x_tl = 0, y_tl = 0, x_br = 0, y_br = 0, bw = 0, bh = 0;
    int src_w = 0, src_h = 0;
    int cam = -1, win_x0 = 0, win_full_w = 0;  // split calibration: this view is columns [win_x0, win_x0 + roi_w) of camera cam's win_full_w-wide warped image (0: the whole image)
    float gain = 1.f;
    float *xmap = nullptr, *ymap = nullptr;
    size_t map_pitch = 0;
    float *mesh[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [buffer][x|y]
    int mesh_cur = -1;                                              // buffer the compose path reads (-1: none yet)
    int mesh_pending = -1;                                          // buffer published by vsb_set_mesh, not yet adopted
    cudaEvent_t mesh_ready = nullptr;
    float *mesh_scratch = nullptr;                                  // sum_x | sum_y | cnt (half res) + device copy of the vertex mesh
    float *weight[MAXL] = {};
    uint8_t *P = nullptr;               // pixel (0, 0) of frame 0 inside P_alloc
    uint8_t *P_alloc = nullptr;         // P with a zero frame (1 row above / below, 4 px left, >= 1 px right): taps of remap #2 just outside read 0
    size_t p_pitch = 0, p_frame_stride = 0;
    unsigned p_origin = 0;              // P - P_alloc
    // tap tables (see TapTable): remap #1, valid for caller row pitch t1_src_pitch; remap #2, one per mesh buffer
    int *t1_off = nullptr; float *t1_w = nullptr; size_t t1_plane = 0; int t1_pitch = 0; size_t t1_src_pitch = 0;
    bool t1_nv12 = false, t1_nv12_unsafe = false;  // the table addresses an NV12 luma plane (fused conversion) / it holds a weight that path cannot take
    int *t1_flag = nullptr;                         // device int raised by the builder
    int *t2_off[2] = {nullptr, nullptr}; float *t2_w[2] = {nullptr, nullptr}; size_t t2_plane = 0; int t2_pitch = 0;
    bool t2_unsafe[2] = {false, false};  // the table of that mesh buffer holds a weight the scaled tap chain cannot take: coordinate kernel
    uint8_t *G0 = nullptr;
    size_t g0_frame_stride = 0;
    CUtensorMap g0_map;                 // TMA descriptor of G0 for k_down2 (valid when g0_map_ok)
    bool g0_map_ok = false;
    int16_t *G[MAXL] = {};              // generic path (num_bands < 3) only: s16 Gaussian levels >= 1
    size_t g_frame_stride[MAXL] = {};
    uint8_t *G2 = nullptr;              // fast path: u8 Gaussian level 2
    size_t g2_frame_stride = 0;
    uint8_t *G1 = nullptr;              // fast path: u8 Gaussian level 1 (written by k_down2, read by k_blend)
    size_t g1_frame_stride = 0;
    uint8_t *Gu[MAXL] = {};             // fast path: u8 Gaussian levels 3..nb (Gu[2] aliases G2)
    size_t gu_frame_stride[MAXL] = {};
    uint8_t *M0 = nullptr;              // bordered seam mask (u8, bw x bh): W0 = M0 * (1/255)
    std::vector<uint8_t> g2_needed;     // per k_down2 tile: computed (1) or skipped (0); host copy for vsb_debug_read
    int d2_tiles_x = 0, d2_tiles_y = 0;
    std::vector<uint32_t> s1_tiles, s2_tiles;  // 128 x 8 tiles k_remap_stage1 / k_remap_stage2 compute (packed view|tx|ty)
    std::vector<uint32_t> s1s_tiles;           // 32 x 32 tiles of the staged remap #1 (k_remap_stage1_st), same packing
    std::vector<int> w0_cols;                  // per plane column: number of non-zero level-0 weights (view -> strip ownership)
};

}  // namespace vsb

struct vsb_stitcher {
    vsb_config cfg;
    int device = 0;
    bool prepared = false, finalized = false;
    int nb = 0;
    int roi_final[4] = {0, 0, 0, 0}, roi[4] = {0, 0, 0, 0};
    int cw[vsb::MAXL] = {}, ch[vsb::MAXL] = {};
    float *dw[vsb::MAXL] = {};
    int views_inited = 0;
    vsb::View v[vsb::MAXV];
    vsb::LgBlendPlan *d_plan = nullptr;  // generic path
    // fast path (num_bands >= 3): static tile tables + per-frame canvas buffer
    bool fast = false;
    vsb::CoarseGeo cgeo;
    int ct = 64;                          // k_coarse tile edge (level-2 samples)
    size_t coarse_smem = 0;
    uint32_t *d_blend_views = nullptr, *d_coarse_views = nullptr, *d_down2_tiles = nullptr;
    uint32_t *d_blend_lists = nullptr;   // interior tiles (tile_x | tile_y << 12 | view << 24), then the other tiles (k_blend_int / k_blend_seam)
    int n_blend_int = 0, n_blend_seam = 0;
    int blend_tiles_x = 0, blend_tiles_y = 0, coarse_tiles_x = 0, coarse_tiles_y = 0, n_down2_tiles = 0;
    int16_t *C2 = nullptr;
    size_t c2_frame_stride = 0;
    uint32_t *d_s1_tiles = nullptr, *d_s2_tiles = nullptr;  // concatenated per-view lists, view order
    uint32_t *d_s1s_ids = nullptr;                           // ... of the staged remap #1
    vsb::S1STile *d_s1s_tiles = nullptr;                     // its tile records (footprint boxes: k_s1_boxes, rebuilt with the tap tables)
    vsb::CoarseView *d_coarse_desc = nullptr;
    // view-sharded mode (vsb_shard_set): this rank's views and canvas strip; host copies of the tile tables
    std::vector<uint32_t> h_bviews, h_cviews, h_d2tiles;
    int shard_rank = -1, shard_world = 1, strip_w = 0;
    bool owned[vsb::MAXV] = {};
    // batched exchange plan (vsb_shard_plan): per peer, the rectangles this rank sends / receives, on the device
    vsb::ShardRect *d_send[vsb::MAXV] = {}, *d_recv[vsb::MAXV] = {};
    int n_send[vsb::MAXV] = {}, n_recv[vsb::MAXV] = {};
    size_t send_bytes[vsb::MAXV] = {}, recv_bytes[vsb::MAXV] = {};  // per frame
    // native transport of the view-sharded mode (vsb_shard_init): NCCL communicator, per-peer exchange buffers (two sets:
    // submissions alternate between two halves of the frame slots so that exchange k overlaps front half k + 1)
    ncclComm_t comm = nullptr;
    int owners[vsb::MAXV] = {};
    uint8_t *x_send[2][vsb::MAXV] = {}, *x_recv[2][vsb::MAXV] = {};
    int x_frames = 0;                     // frames per submission the buffers are sized for
    cudaStream_t sh_front = nullptr, sh_comm = nullptr, sh_back = nullptr;
    cudaEvent_t ev_call = nullptr, ev_packed[2] = {}, ev_recv[2] = {}, ev_back[2] = {};
    bool ev_back_valid[2] = {false, false};
    unsigned shard_seq = 0;
    bool tiles_dirty = true;
    cudaStream_t setup_stream = nullptr, mesh_stream = nullptr, io_stream = nullptr, in_stream = nullptr, out_stream = nullptr;
    cudaEvent_t ev_in[vsb::MAX_BATCH] = {}, ev_done[vsb::MAX_BATCH] = {};
    cudaEvent_t last_compose = nullptr;
    bool last_compose_valid = false;
    std::mutex mu;  // guards mesh publication
    int f0 = 0;                         // first frame slot the launch helpers address (vsb_compose splits a batch over two streams)
    cudaStream_t sub[vsb::MAX_SPLIT] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[vsb::MAX_SPLIT] = {}, ev_forks[vsb::MAX_SPLIT] = {};
    int launches = 0, launches_last = 0;  // running count of the submission in flight / count of the last finished one
    // vsb_feed / vsb_blend bookkeeping (frame slot 0)
    // host-buffer path staging
    uint8_t *stage_src[vsb::HOST_DEPTH] = {};
    int16_t *stage_out[vsb::HOST_DEPTH] = {};
    cudaEvent_t ev_host[vsb::HOST_DEPTH] = {};
    unsigned host_seq = 0;      // submissions enqueued by vsb_submit_host
    int host_pending = 0;       // ... of which not yet waited for
    size_t stage_src_pitch = 0, stage_src_frame = 0, stage_out_pitch = 0, stage_out_frame = 0;
    int stage_src_w = 0, stage_src_h = 0, stage_views = 0, stage_batch = 0;  // what stage_src / stage_nv12 were sized for
    void *calib_state = nullptr;            // seam-scale state of vsb_calibrate_rig_device (vsb_calib.cu)
    void (*calib_dtor)(void *) = nullptr;
    int rig_projection = -1, rig_src_w = 0, rig_src_h = 0;
    float rig_scale = 0.f;
    // wire / consumer formats (vsb_set_formats): NV12 input goes through k_nv12_to_bgr into nv_bgr; CV_8UC3 output is k_blend<true>
    int in_format = VSB_IN_BGR8, out_format = VSB_OUT_S16C3;
    uint8_t *nv_bgr = nullptr;          // [max_batch][num_views] BGR images, rows nv_pitch bytes apart
    size_t nv_pitch = 0, nv_stride = 0;
    int nv_w = 0, nv_h = 0;
    // compose_scale != 1 (vsb_set_compose_scale): the caller's full_w x full_h frames go through cuda::resize into cs_buf first
    bool prescale = false;
    double compose_scale = 1.0;
    int full_w = 0, full_h = 0, comp_w = 0, comp_h = 0;
    uint8_t *cs_buf = nullptr;          // [max_batch][num_views] resized BGR frames, rows cs_pitch bytes apart
    size_t cs_pitch = 0, cs_stride = 0;
    int cs_w = 0, cs_h = 0;
    int *cons_tab = nullptr;            // consumer resize tables for (cons_w x cons_ih): xofs | yofs | ia | ib
    int cons_w = 0, cons_ih = 0;
    uint8_t *stage_nv12[vsb::HOST_DEPTH] = {};  // host path: device copy of the caller's NV12 frames
    size_t stage_nv12_frame = 0;
    // optional per-kernel timing (vsb_set_profiling): events bracket every launch of the last submission
    bool profiling = false;
    int n_stages = 0;
    cudaEvent_t prof_ev[VSB_MAX_STAGES + 1] = {};
    const char *stage_name[VSB_MAX_STAGES] = {};
    double stage_bytes[VSB_MAX_STAGES] = {};
};

namespace vsb {

// NVTX range per stage (the reference keeps std::chrono stamps per stage in times[5], A/timed.cpp:43-44, never printed)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

#define CK(expr) do { int _r = check_cuda((expr), #expr); if (_r != VSB_OK) return _r; } while (0)
#define REQ(cond, code, ...) do { if (!(cond)) return fail((code), __VA_ARGS__); } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static void free_view(View &V)
{
    cudaFree(V.xmap); cudaFree(V.ymap);
    for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c) cudaFree(V.mesh[b][c]);
    cudaFree(V.mesh_scratch);
    for (int k = 0; k < MAXL; ++k) { cudaFree(V.weight[k]); cudaFree(V.G[k]); }
    cudaFree(V.P_alloc); cudaFree(V.G0); cudaFree(V.G1); cudaFree(V.G2); cudaFree(V.M0);
    cudaFree(V.t1_off); cudaFree(V.t1_w); cudaFree(V.t1_flag);
    for (int b = 0; b < 2; ++b) { cudaFree(V.t2_off[b]); cudaFree(V.t2_w[b]); }
    for (int k = 3; k < MAXL; ++k) cudaFree(V.Gu[k]);
    if (V.mesh_ready) cudaEventDestroy(V.mesh_ready);
    V = View();
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
        cudaGetLastError();
    }
    return fn;
}

// G0 of one view as a 3-D u8 tensor {bw, bh, 3 * F}; the box is one k_down2 region.  Needs 16-byte aligned rows.
static void make_g0_map(View &V, int F)
{
    V.g0_map_ok = false;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || (V.bw & 15) != 0 || std::getenv("VSB_NO_TMA")) return;
    const cuuint64_t dims[3] = {(cuuint64_t)V.bw, (cuuint64_t)V.bh, (cuuint64_t)3 * F};
    const cuuint64_t strides[2] = {(cuuint64_t)V.bw, (cuuint64_t)V.bw * V.bh};
    const cuuint32_t box[3] = {(cuuint32_t)(D2_R0VEC * 16), (cuuint32_t)D2_R0H, 1u}, estr[3] = {1u, 1u, 1u};
    if (V.bw < (int)box[0] || V.bh < (int)box[1]) return;  // planes smaller than one region keep the plain loads
    const CUresult r = enc(&V.g0_map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, V.G0, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    V.g0_map_ok = r == CUDA_SUCCESS;
}

static int build_plan(vsb_stitcher *s)
{
    LgBlendPlan h;
    std::memset(&h, 0, sizeof(h));
    h.n_views = s->cfg.num_views; h.nb = s->nb;
    h.out_w = s->roi_final[2]; h.out_h = s->roi_final[3];
    for (int k = 0; k <= s->nb; ++k) { h.cw[k] = s->cw[k]; h.ch[k] = s->ch[k]; h.dw[k] = s->dw[k]; }
    for (int i = 0; i < s->cfg.num_views; ++i) {
        const View &V = s->v[i];
        LgBlendView &B = h.v[i];
        B.g0 = V.G0; B.g0_frame_stride = V.g0_frame_stride;
        for (int k = 0; k <= s->nb; ++k) { B.g[k] = V.G[k]; B.w[k] = V.weight[k]; B.g_frame_stride[k] = V.g_frame_stride[k]; }
        B.x_tl = V.x_tl; B.y_tl = V.y_tl; B.bw = V.bw; B.bh = V.bh;
    }
    if (!s->d_plan) CK(cudaMalloc(&s->d_plan, sizeof(LgBlendPlan)));
    CK(cudaMemcpyAsync(s->d_plan, &h, sizeof(h), cudaMemcpyHostToDevice, s->setup_stream));
    CK(cudaStreamSynchronize(s->setup_stream));
    return VSB_OK;
}


// ---- fast path (num_bands >= 3): static tables -------------------------------------------------------------------
static inline int fdiv2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }  // floor(v / 2)

// region offsets of k_coarse, relative to the tile origin at each level (see vsb_blend_kernels.cuh)
static void coarse_geometry(int nb, int CT, CoarseGeo &g)
{
    std::memset(&g, 0, sizeof(g));
    const int nlev = nb - 1;
    g.nlev = nlev;
    int a_lo[C_MAXJ], a_hi[C_MAXJ], g_lo[C_MAXJ], g_hi[C_MAXJ];
    a_lo[0] = 0; a_hi[0] = CT - 1;
    for (int j = 1; j < nlev; ++j) { a_lo[j] = fdiv2(a_lo[j - 1]) - 1; a_hi[j] = fdiv2(a_hi[j - 1]) + 1; }
    g_lo[nlev - 1] = a_lo[nlev - 1]; g_hi[nlev - 1] = a_hi[nlev - 1];
    for (int j = nlev - 2; j >= 0; --j) { g_lo[j] = std::min(a_lo[j], 2 * g_lo[j + 1] - 2); g_hi[j] = std::max(a_hi[j], 2 * g_hi[j + 1] + 2); }
    int a_total = 0, g_total = 0;
    for (int j = 0; j < nlev; ++j) {
        g.a_lo[j] = a_lo[j]; g.a_n[j] = a_hi[j] - a_lo[j] + 1;   // even by construction (64, 34, 20, 12, 8, 6)
        g.g_lo[j] = g_lo[j]; g.g_n[j] = g_hi[j] - g_lo[j] + 1;
        g.a_off[j] = a_total; a_total += (int)align_up((size_t)g.a_n[j] * g.a_n[j], 8);
        g.g_off[j] = g_total; g_total += (int)align_up((size_t)g.a_n[j] * g.a_n[j], 16);
    }
    g.a_total = a_total;
    int f_total = 0;
    for (int j = 0; j < nlev; ++j) { g.f_off[j] = f_total; f_total += (int)align_up((size_t)g.a_n[j] * g.a_n[j], 4); }
    for (int j = 1; j < nlev; ++j) { g.d_off[j] = f_total; f_total += (int)align_up((size_t)g.a_n[j] * g.a_n[j], 4); }
    g.f_total = f_total;
}
static size_t coarse_smem_bytes(const CoarseGeo &g)
{
    const int j = g.nlev - 1;
    (void)j;
    return (size_t)g.f_total * 4;
}

struct HostWeights {  // nonzero structure of one view's static weight pyramid
    std::vector<std::vector<uint8_t>> nz;  // [level][h * w]: weight != 0
    std::vector<uint8_t> one[2];           // levels 0 and 1: weight == 1.0f exactly
    bool all_one(int k, int x0, int y0, int x1, int y1) const  // half-open plane rect; false when it leaves the plane
    {
        if (x0 < 0 || y0 < 0 || x1 > w[k] || y1 > h[k]) return false;
        for (int y = y0; y < y1; ++y) {
            const uint8_t *r = one[k].data() + (size_t)y * w[k];
            for (int x = x0; x < x1; ++x) if (!r[x]) return false;
        }
        return true;
    }
    int w[MAXL], h[MAXL];
    bool any(int k, int x0, int y0, int x1, int y1) const  // half-open plane rect, clipped here
    {
        x0 = std::max(x0, 0); y0 = std::max(y0, 0); x1 = std::min(x1, w[k]); y1 = std::min(y1, h[k]);
        for (int y = y0; y < y1; ++y) {
            const uint8_t *r = nz[k].data() + (size_t)y * w[k];
            for (int x = x0; x < x1; ++x) if (r[x]) return true;
        }
        return false;
    }
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (device, function), not per handle: raise it once per device to the
// architectural maximum so that handles with different needs can coexist (a smaller later request must never lower it).
template <typename Kernel>
static int raise_dynamic_smem(Kernel kernel, int device)
{
    static std::mutex mu;
    static bool done[64] = {};
    std::lock_guard<std::mutex> lk(mu);
    if (device >= 0 && device < 64 && done[device]) return VSB_OK;
    int optin = 0;
    CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    if (device >= 0 && device < 64) done[device] = true;
    return VSB_OK;
}

// The two k_blend launches run over static tile lists derived from the per-tile words (view bits, bit 30 = interior,
// bit 31 = tile of another rank's strip in view-sharded mode): interior tiles carry their single view in the list entry.
static int upload_blend_lists(vsb_stitcher *s, const std::vector<uint32_t> &words)
{
    std::vector<uint32_t> li, ls;
    for (int ty = 0; ty < s->blend_tiles_y; ++ty)
        for (int tx = 0; tx < s->blend_tiles_x; ++tx) {
            const uint32_t w = words[(size_t)ty * s->blend_tiles_x + tx];
            if (w & 0x80000000u) continue;
            const uint32_t id = (uint32_t)tx | ((uint32_t)ty << 12);
            if (w & 0x40000000u) {
                int v = 0;
                while (!(w >> v & 1)) ++v;
                li.push_back(id | ((uint32_t)v << 24));
            } else {
                ls.push_back(id);
            }
        }
    cudaFree(s->d_blend_lists); s->d_blend_lists = nullptr;
    s->n_blend_int = (int)li.size(); s->n_blend_seam = (int)ls.size();
    li.insert(li.end(), ls.begin(), ls.end());
    CK(cudaMalloc(&s->d_blend_lists, std::max<size_t>(li.size(), 1) * 4));
    if (!li.empty()) CK(cudaMemcpy(s->d_blend_lists, li.data(), li.size() * 4, cudaMemcpyHostToDevice));
    return VSB_OK;
}

static int build_fast_plan(vsb_stitcher *s)
{
    const int n = s->cfg.num_views, nb = s->nb, F = s->cfg.max_batch;
    // nonzero structure of every weight level (read back once; calibration time)
    std::vector<HostWeights> hw(n);
    for (int i = 0; i < n; ++i) {
        const View &V = s->v[i];
        hw[i].nz.resize(nb + 1);
        for (int k = 0; k <= nb; ++k) {
            const int w = V.bw >> k, h = V.bh >> k;
            hw[i].w[k] = w; hw[i].h[k] = h;
            std::vector<float> tmp((size_t)w * h);
            CK(cudaMemcpy(tmp.data(), V.weight[k], tmp.size() * sizeof(float), cudaMemcpyDeviceToHost));
            hw[i].nz[k].resize(tmp.size());
            for (size_t j = 0; j < tmp.size(); ++j) hw[i].nz[k][j] = tmp[j] != 0.f;
            if (k <= 1) {
                hw[i].one[k].resize(tmp.size());
                for (size_t j = 0; j < tmp.size(); ++j) hw[i].one[k][j] = tmp[j] == 1.0f;
            }
        }
    }
    for (int i = 0; i < n; ++i) {
        View &V = s->v[i];
        V.w0_cols.assign(V.bw, 0);
        for (int y = 0; y < V.bh; ++y)
            for (int x = 0; x < V.bw; ++x) V.w0_cols[x] += hw[i].nz[0][(size_t)y * V.bw + x];
    }
    // ---- k_blend: views with level-0 / level-1 weight per 64 x 32 canvas tile
    s->blend_tiles_x = (s->cw[0] + BL_TW - 1) / BL_TW; s->blend_tiles_y = (s->ch[0] + BL_TH - 1) / BL_TH;
    std::vector<uint32_t> bviews((size_t)s->blend_tiles_x * s->blend_tiles_y, 0);
    for (int ty = 0; ty < s->blend_tiles_y; ++ty)
        for (int tx = 0; tx < s->blend_tiles_x; ++tx)
            for (int i = 0; i < n; ++i) {
                const View &V = s->v[i];
                const int x0 = tx * BL_TW - V.x_tl, y0 = ty * BL_TH - V.y_tl;
                const int x1 = fdiv2(tx * BL_TW) - 1 - (V.x_tl >> 1), y1 = fdiv2(ty * BL_TH) - 1 - (V.y_tl >> 1);
                if (hw[i].any(0, x0, y0, x0 + BL_TW, y0 + BL_TH) || hw[i].any(1, x1, y1, x1 + BL_R1W, y1 + BL_R1H))
                    bviews[(size_t)ty * s->blend_tiles_x + tx] |= 1u << i;
            }
    // interior tiles (bit 30): exactly one view, whose level-0 and level-1 weights are exactly 1 over everything the tile reads --
    // k_blend then skips the mask / weight / weight-sum loads (most of the panorama away from the seams)
    for (int ty = 0; ty < s->blend_tiles_y; ++ty)
        for (int tx = 0; tx < s->blend_tiles_x; ++tx) {
            uint32_t &b = bviews[(size_t)ty * s->blend_tiles_x + tx];
            if (b == 0 || (b & (b - 1)) != 0) continue;
            int i = 0;
            while (!(b >> i & 1)) ++i;
            const View &V = s->v[i];
            const int x0 = tx * BL_TW - V.x_tl, y0 = ty * BL_TH - V.y_tl;
            const int x1 = fdiv2(tx * BL_TW) - 1 - (V.x_tl >> 1), y1 = fdiv2(ty * BL_TH) - 1 - (V.y_tl >> 1);
            if ((tx + 1) * BL_TW <= s->cw[0] && (ty + 1) * BL_TH <= s->ch[0] && x1 >= 0 && y1 >= 0 &&
                hw[i].all_one(0, x0, y0, x0 + BL_TW, y0 + BL_TH) && hw[i].all_one(1, x1, y1, x1 + BL_R1W, y1 + BL_R1H))
                b |= 0x40000000u;
        }
    // ---- k_coarse: views with weight at any level >= 2 per 64 x 64 level-2 canvas tile
    // k_coarse tile edge: 32 level-2 samples while every region size stays even (num_bands <= 6), else 64
    static const int ct_env = [] { const char *e = std::getenv("VSB_COARSE_TILE"); return e ? std::atoi(e) : 0; }();
    s->ct = ct_env == 32 && nb <= 6 ? 32 : (ct_env == 64 || nb > 6 || s->cfg.max_batch > 2 ? 64 : 32);  // 32 pays off at 1-2 frames per launch (measured: +8 % single-frame rate, -10 % at 8 frames)
    const int CT = s->ct;
    coarse_geometry(nb, CT, s->cgeo);
    s->coarse_smem = coarse_smem_bytes(s->cgeo);
    s->coarse_tiles_x = (s->cw[2] + CT - 1) / CT; s->coarse_tiles_y = (s->ch[2] + CT - 1) / CT;
    std::vector<uint32_t> cviews((size_t)s->coarse_tiles_x * s->coarse_tiles_y, 0);
    for (int ty = 0; ty < s->coarse_tiles_y; ++ty)
        for (int tx = 0; tx < s->coarse_tiles_x; ++tx)
            for (int i = 0; i < n; ++i) {
                const View &V = s->v[i];
                bool any = false;
                for (int j = 0; j < s->cgeo.nlev && !any; ++j) {
                    const int k = 2 + j;
                    const int x0 = ((tx * CT) >> j) + s->cgeo.a_lo[j] - (V.x_tl >> k), y0 = ((ty * CT) >> j) + s->cgeo.a_lo[j] - (V.y_tl >> k);
                    any = hw[i].any(k, x0, y0, x0 + s->cgeo.a_n[j], y0 + s->cgeo.a_n[j]);
                }
                if (any) cviews[(size_t)ty * s->coarse_tiles_x + tx] |= 1u << i;
            }
    // ---- k_down2: the G2 tiles some k_coarse / k_blend tile reads
    std::vector<uint32_t> d2tiles;
    for (int i = 0; i < n; ++i) {
        View &V = s->v[i];
        const int w2 = V.bw >> 2, h2 = V.bh >> 2;
        V.d2_tiles_x = (w2 + D2_TW - 1) / D2_TW; V.d2_tiles_y = (h2 + D2_TH - 1) / D2_TH;
        V.g2_needed.assign((size_t)V.d2_tiles_x * V.d2_tiles_y, 0);
        auto mark = [&](int x0, int y0, int x1, int y1) {  // half-open rect in level-2 plane coordinates
            x0 = std::max(x0, 0); y0 = std::max(y0, 0); x1 = std::min(x1, w2); y1 = std::min(y1, h2);
            if (x0 >= x1 || y0 >= y1) return;
            for (int ty = y0 / D2_TH; ty <= (y1 - 1) / D2_TH; ++ty)
                for (int tx = x0 / D2_TW; tx <= (x1 - 1) / D2_TW; ++tx) V.g2_needed[(size_t)ty * V.d2_tiles_x + tx] = 1;
        };
        for (int ty = 0; ty < s->coarse_tiles_y; ++ty)
            for (int tx = 0; tx < s->coarse_tiles_x; ++tx)
                if (cviews[(size_t)ty * s->coarse_tiles_x + tx] >> i & 1) {
                    // the region is read through BORDER_REFLECT_101 (at most 2 samples beyond the loaded range fold back inside it)
                    const int x0 = tx * CT + s->cgeo.g_lo[0] - (V.x_tl >> 2), y0 = ty * CT + s->cgeo.g_lo[0] - (V.y_tl >> 2);
                    mark(x0, y0, x0 + s->cgeo.g_n[0], y0 + s->cgeo.g_n[0]);
                }
        for (int ty = 0; ty < s->blend_tiles_y; ++ty)
            for (int tx = 0; tx < s->blend_tiles_x; ++tx)
                if (bviews[(size_t)ty * s->blend_tiles_x + tx] >> i & 1) {
                    const int x0 = fdiv2(fdiv2(tx * BL_TW)) - 2 - (V.x_tl >> 2), y0 = fdiv2(fdiv2(ty * BL_TH)) - 2 - (V.y_tl >> 2);
                    mark(x0, y0, x0 + BL_R2W, y0 + BL_R2H);
                }
        for (int ty = 0; ty < V.d2_tiles_y; ++ty)
            for (int tx = 0; tx < V.d2_tiles_x; ++tx)
                if (V.g2_needed[(size_t)ty * V.d2_tiles_x + tx]) d2tiles.push_back((uint32_t)i | ((uint32_t)tx << 8) | ((uint32_t)ty << 20));
    }
    // ---- k_remap_stage2: the 128 x 8 tiles of G0 that a computed k_down2 tile or a k_blend tile reads
    for (int i = 0; i < n; ++i) {
        View &V = s->v[i];
        const int tw = RM_BX * RM_PX, ntx = (V.bw + tw - 1) / tw, nty = (V.bh + RM_BY - 1) / RM_BY;
        std::vector<uint8_t> need((size_t)ntx * nty, 0);
        auto mark = [&](int x0, int y0, int x1, int y1) {  // half-open rect in level-0 plane coordinates
            x0 = std::max(x0, 0); y0 = std::max(y0, 0); x1 = std::min(x1, V.bw); y1 = std::min(y1, V.bh);
            if (x0 >= x1 || y0 >= y1) return;
            for (int ty = y0 / RM_BY; ty <= (y1 - 1) / RM_BY; ++ty)
                for (int tx = x0 / tw; tx <= (x1 - 1) / tw; ++tx) need[(size_t)ty * ntx + tx] = 1;
        };
        for (int ty = 0; ty < V.d2_tiles_y; ++ty)
            for (int tx = 0; tx < V.d2_tiles_x; ++tx)
                if (V.g2_needed[(size_t)ty * V.d2_tiles_x + tx])
                    mark(4 * tx * D2_TW - 16, 4 * ty * D2_TH - 6, 4 * tx * D2_TW - 16 + 16 * D2_R0VEC, 4 * ty * D2_TH - 6 + D2_R0H);
        for (int ty = 0; ty < s->blend_tiles_y; ++ty)
            for (int tx = 0; tx < s->blend_tiles_x; ++tx)
                if (bviews[(size_t)ty * s->blend_tiles_x + tx] >> i & 1)
                    mark(tx * BL_TW - V.x_tl, ty * BL_TH - V.y_tl, (tx + 1) * BL_TW - V.x_tl, (ty + 1) * BL_TH - V.y_tl);
        V.s2_tiles.clear();
        for (int ty = 0; ty < nty; ++ty)
            for (int tx = 0; tx < ntx; ++tx)
                if (need[(size_t)ty * ntx + tx]) V.s2_tiles.push_back((uint32_t)i | ((uint32_t)tx << 8) | ((uint32_t)ty << 20));
    }
    s->tiles_dirty = true;
    s->n_down2_tiles = (int)d2tiles.size();
    s->h_d2tiles = d2tiles;
    cudaFree(s->d_blend_views); cudaFree(s->d_coarse_views); cudaFree(s->d_down2_tiles); cudaFree(s->C2);
    s->d_blend_views = s->d_coarse_views = s->d_down2_tiles = nullptr; s->C2 = nullptr;
    CK(cudaMalloc(&s->d_blend_views, bviews.size() * 4));
    CK(cudaMalloc(&s->d_coarse_views, cviews.size() * 4));
    CK(cudaMalloc(&s->d_down2_tiles, std::max<size_t>(d2tiles.size(), 1) * 4));
    s->h_bviews = bviews; s->h_cviews = cviews; s->shard_rank = -1; s->shard_world = 1;
    for (int i = 0; i < MAXV; ++i) s->owned[i] = i < n;
    CK(cudaMemcpy(s->d_blend_views, bviews.data(), bviews.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->d_coarse_views, cviews.data(), cviews.size() * 4, cudaMemcpyHostToDevice));
    { int r = upload_blend_lists(s, bviews); if (r != VSB_OK) return r; }
    if (!d2tiles.empty()) CK(cudaMemcpy(s->d_down2_tiles, d2tiles.data(), d2tiles.size() * 4, cudaMemcpyHostToDevice));
    s->c2_frame_stride = (size_t)3 * s->cw[2] * s->ch[2];
    CK(cudaMalloc(&s->C2, s->c2_frame_stride * sizeof(int16_t) * F));
    if (s->coarse_smem > 48 * 1024) { int r = raise_dynamic_smem(k_coarse<64>, s->device); if (r != VSB_OK) return r; }
    std::vector<CoarseView> desc(n);
    std::memset(desc.data(), 0, sizeof(CoarseView) * n);
    for (int i = 0; i < n; ++i) {
        const View &V = s->v[i];
        for (int j = 0; j < s->cgeo.nlev; ++j) { desc[i].g[j] = V.Gu[2 + j]; desc[i].g_fs[j] = V.gu_frame_stride[2 + j]; desc[i].w[j] = V.weight[2 + j]; }
        desc[i].x_tl = V.x_tl; desc[i].y_tl = V.y_tl; desc[i].bw = V.bw; desc[i].bh = V.bh;
    }
    cudaFree(s->d_coarse_desc); s->d_coarse_desc = nullptr;
    CK(cudaMalloc(&s->d_coarse_desc, sizeof(CoarseView) * n));
    CK(cudaMemcpy(s->d_coarse_desc, desc.data(), sizeof(CoarseView) * n, cudaMemcpyHostToDevice));
    return VSB_OK;
}

// static weight sums, accumulated in view order like successive feed_online calls would
static int finalize(vsb_stitcher *s)
{
    const dim3 b(32, 8);
    for (int k = 0; k <= s->nb; ++k) CK(cudaMemsetAsync(s->dw[k], 0, sizeof(float) * s->cw[k] * s->ch[k], s->setup_stream));
    for (int i = 0; i < s->cfg.num_views; ++i) {
        const View &V = s->v[i];
        for (int k = 0; k <= s->nb; ++k) {
            const int bwk = V.bw >> k, bhk = V.bh >> k;
            k_accum_weight<<<grid2d(bwk, bhk, b), b, 0, s->setup_stream>>>(V.weight[k], bwk, bhk, s->dw[k], s->cw[k], V.x_tl >> k, V.y_tl >> k);
        }
    }
    int r = check_launch("k_accum_weight");
    if (r != VSB_OK) return r;
    CK(cudaStreamSynchronize(s->setup_stream));
    s->fast = s->nb >= 3;
    r = s->fast ? build_fast_plan(s) : build_plan(s);
    if (r != VSB_OK) return r;
    s->finalized = true;
    return VSB_OK;
}

// view-sharded handles keep only the tiles of their own views in the device tile lists, so that ONE launch of each front-half
// kernel covers an arbitrary (not necessarily contiguous) set of owned views
static inline bool front_view(const vsb_stitcher *s, int i) { return s->shard_rank < 0 || s->owned[i]; }

static int ready_for_frames(vsb_stitcher *s)
{
    REQ(s->finalized, VSB_ERR_STATE, "compose: prepare + init_view for every view must come first");
    for (int i = 0; i < s->cfg.num_views; ++i) {
        REQ(s->v[i].has_maps, VSB_ERR_STATE, "compose: view %d has no projection maps (vsb_set_maps)", i);
        REQ(!s->cfg.enable_local || s->v[i].mesh_cur >= 0 || s->v[i].mesh_pending >= 0, VSB_ERR_STATE,
            "compose: enable_local is set but view %d has no mesh (vsb_set_mesh)", i);
    }
    return VSB_OK;
}

// Host-side wait for the frame work this handle has submitted (the event every per-frame entry point records), instead of
// cudaDeviceSynchronize(): other handles, other streams and the mesh builder keep running.
static int wait_own_frames(vsb_stitcher *s)
{
    cudaEvent_t ev = nullptr;
    {
        std::lock_guard<std::mutex> lk(s->mu);
        if (s->last_compose_valid) ev = s->last_compose;
    }
    if (ev) CK(cudaEventSynchronize(ev));
    return VSB_OK;
}

// adopt meshes published by vsb_set_mesh: the compose stream waits for the builder's event, then flips buffers
static int adopt_meshes(vsb_stitcher *s, cudaStream_t st)
{
    std::lock_guard<std::mutex> lk(s->mu);
    for (int i = 0; i < s->cfg.num_views; ++i) {
        View &V = s->v[i];
        if (V.mesh_pending >= 0) {
            CK(cudaStreamWaitEvent(st, V.mesh_ready, 0));
            V.mesh_cur = V.mesh_pending;
            V.mesh_pending = -1;
        }
    }
    return VSB_OK;
}

// profiling: record an event before the first kernel and after every kernel of a submission
static void prof_begin(vsb_stitcher *s, cudaStream_t st)
{
    if (!s->profiling) return;
    s->n_stages = 0;
    cudaEventRecord(s->prof_ev[0], st);
}
static void prof_stage(vsb_stitcher *s, cudaStream_t st, const char *name, double bytes)
{
    if (!s->profiling || s->n_stages >= VSB_MAX_STAGES) return;
    s->stage_name[s->n_stages] = name;
    s->stage_bytes[s->n_stages] = bytes;
    ++s->n_stages;
    cudaEventRecord(s->prof_ev[s->n_stages], st);
}

static void fill_tilemap(TileMap &tm, int n, const int *w, const int *h, int tile_w, int tile_h)
{
    tm.n = n;
    tm.start[0] = 0;
    for (int i = 0; i < n; ++i) {
        tm.tiles_x[i] = (w[i] + tile_w - 1) / tile_w;
        tm.start[i + 1] = tm.start[i] + tm.tiles_x[i] * ((h[i] + tile_h - 1) / tile_h);
    }
}



// (re)uploads the concatenated stage-1 / stage-2 tile lists after calibration changed them
static int sync_tile_lists(vsb_stitcher *s)
{
    if (!s->tiles_dirty) return VSB_OK;
    std::vector<uint32_t> a, b, c;
    for (int i = 0; i < s->cfg.num_views; ++i) {
        s->v[i].t1_src_pitch = 0;  // the footprint boxes live next to the lists: rebuilt (with the tap table) by the next compose
        if (!front_view(s, i)) continue;
        a.insert(a.end(), s->v[i].s1_tiles.begin(), s->v[i].s1_tiles.end());
        b.insert(b.end(), s->v[i].s2_tiles.begin(), s->v[i].s2_tiles.end());
        c.insert(c.end(), s->v[i].s1s_tiles.begin(), s->v[i].s1s_tiles.end());
    }
    { int r = wait_own_frames(s); if (r != VSB_OK) return r; }  // submissions in flight still read the old lists
    cudaFree(s->d_s1_tiles); cudaFree(s->d_s2_tiles); cudaFree(s->d_s1s_ids); cudaFree(s->d_s1s_tiles);
    s->d_s1_tiles = s->d_s2_tiles = s->d_s1s_ids = nullptr; s->d_s1s_tiles = nullptr;
    CK(cudaMalloc(&s->d_s1_tiles, std::max<size_t>(a.size(), 1) * 4));
    CK(cudaMalloc(&s->d_s2_tiles, std::max<size_t>(b.size(), 1) * 4));
    CK(cudaMalloc(&s->d_s1s_ids, std::max<size_t>(c.size(), 1) * 4));
    CK(cudaMalloc(&s->d_s1s_tiles, std::max<size_t>(c.size(), 1) * sizeof(S1STile)));
    if (!a.empty()) CK(cudaMemcpy(s->d_s1_tiles, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
    if (!b.empty()) CK(cudaMemcpy(s->d_s2_tiles, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
    if (!c.empty()) CK(cudaMemcpy(s->d_s1s_ids, c.data(), c.size() * 4, cudaMemcpyHostToDevice));
    s->tiles_dirty = false;
    return VSB_OK;
}

// every 128 x 8 tile of a w x h plane
static void all_tiles(int view, int w, int h, std::vector<uint32_t> &out)
{
    out.clear();
    for (int ty = 0; ty < (h + RM_BY - 1) / RM_BY; ++ty)
        for (int tx = 0; tx < (w + RM_BX * RM_PX - 1) / (RM_BX * RM_PX); ++tx) out.push_back((uint32_t)view | ((uint32_t)tx << 8) | ((uint32_t)ty << 20));
}

// K3: G0 -> G2 for the needed tiles of views [v0, v1)
static int launch_down2(vsb_stitcher *s, int v0, int v1, int n_frames, cudaStream_t st)
{
    NvtxRange nvtx("vsb: pyrDown levels 0-2 (k_down2)");
    Down2Params p;
    std::memset(&p, 0, sizeof(p));
    for (int i = 0; i < s->cfg.num_views; ++i) {
        const View &V = s->v[i];
        p.v[i].g0 = V.G0; p.v[i].g1 = V.G1; p.v[i].g2 = V.G2;
        p.v[i].g0_fs = V.g0_frame_stride; p.v[i].g1_fs = V.g1_frame_stride; p.v[i].g2_fs = V.g2_frame_stride;
        p.v[i].bw = V.bw; p.v[i].bh = V.bh;
    }
    // the tile list is sorted by view: launch the sub-range that belongs to [v0, v1)
    int first = 0, count = 0;
    double bytes = 0;
    for (int i = 0; i < s->cfg.num_views; ++i) {
        int nt = 0;
        if (!front_view(s, i)) continue;  // (view-sharded: the device list holds the owned views' tiles only)
        for (uint8_t b : s->v[i].g2_needed) nt += b;
        if (i < v0) first += nt;
        else if (i < v1) { count += nt; bytes += (double)nt * 3 * (16.0 * D2_TW * D2_TH + 4.0 * D2_TW * D2_TH + D2_TW * D2_TH); }
    }
    p.tiles = s->d_down2_tiles + first;
    p.f0 = s->f0;
    bool tma = true;
    Down2Maps maps;
    std::memset(&maps, 0, sizeof(maps));
    for (int i = v0; i < v1; ++i) { if (!front_view(s, i)) continue; tma = tma && s->v[i].g0_map_ok; maps.g0[i] = s->v[i].g0_map; }
    if (count > 0) {
        if (tma) k_down2<true><<<dim3(count, 3, n_frames), D2_THREADS, 0, st>>>(p, maps);
        else k_down2<false><<<dim3(count, 3, n_frames), D2_THREADS, 0, st>>>(p, maps);
    }
    ++s->launches;
    prof_stage(s, st, "down2", bytes * n_frames);  // G0 of the needed tiles in once + G1 and G2 out once
    return check_launch("k_down2");
}

// one more pyrDown level (k -> k + 1) of views [v0, v1) on the u8 planes
static int launch_down1_level(vsb_stitcher *s, int k, int v0, int v1, int n_frames, cudaStream_t st)
{
    {
        Down1Params p;
        std::memset(&p, 0, sizeof(p));
        double bytes = 0;
        int m = 0;
        for (int i = v0; i < v1; ++i) {
            if (!front_view(s, i)) continue;
            const View &V = s->v[i];
            p.v[m].src = V.Gu[k]; p.v[m].dst = V.Gu[k + 1]; p.v[m].src_fs = V.gu_frame_stride[k]; p.v[m].dst_fs = V.gu_frame_stride[k + 1];
            p.v[m].w = V.bw >> k; p.v[m].h = V.bh >> k;
            const int wd = V.bw >> (k + 1), hd = V.bh >> (k + 1);
            p.tiles_x[m] = (wd + D1_TX - 1) / D1_TX;
            p.start[m + 1] = p.start[m] + p.tiles_x[m] * ((hd + D1_TY - 1) / D1_TY);
            bytes += 3.0 * (V.bw >> k) * (V.bh >> k) + 3.0 * wd * hd;
            ++m;
        }
        p.n = m; p.f0 = s->f0;
        if (m > 0) k_down1<<<dim3(p.start[m], n_frames, 3), dim3(D1_TX, D1_TY), 0, st>>>(p);
        ++s->launches;
        static const char *names[MAXL] = {"", "", "down1_L3", "down1_L4", "down1_L5", "down1_L6", "down1_L7", ""};
        prof_stage(s, st, names[k], bytes * n_frames);
    }
    return VSB_OK;
}

// K3b: Gaussian levels 3..nb of views [v0, v1) (tiny planes), one launch per level
static int launch_down1(vsb_stitcher *s, int v0, int v1, int n_frames, cudaStream_t st)
{
    NvtxRange nvtx("vsb: pyrDown levels 3..nb (k_down_tail)");
    const int nb = s->nb;
    // Levels whose outputs fit in shared memory together (all of them up to ~8k-wide panoramas, all but level 3 at 16k) are
    // produced by ONE k_down_tail launch starting at level k0; the levels before it take one k_down1 launch each.
    int k0 = 2;
    size_t smem = 0;
    for (; k0 < nb; ++k0) {
        smem = 0;
        for (int i = v0; i < v1; ++i) {
            size_t need = 0;
            for (int k = k0 + 1; k <= nb; ++k) need += align_up((size_t)(s->v[i].bw >> k) * (s->v[i].bh >> k), 16);
            smem = std::max(smem, need);
        }
        if (smem <= 200 * 1024) break;
    }
    if (k0 < nb && v1 > v0) {
        double bytes = 0;
        for (int i = v0; i < v1; ++i)
            for (int k = k0; k <= nb; ++k) bytes += 3.0 * (s->v[i].bw >> k) * (s->v[i].bh >> k);
        if (smem > 48 * 1024) { int r = raise_dynamic_smem(k_down_tail, s->device); if (r != VSB_OK) return r; }
        DownTailParams p;
        std::memset(&p, 0, sizeof(p));
        p.nb = nb; p.k0 = k0; p.f0 = s->f0;
        int m = 0;
        for (int i = v0; i < v1; ++i) {
            if (!front_view(s, i)) continue;
            const View &V = s->v[i];
            DownTailView &D = p.v[m++];
            D.g2 = V.Gu[k0]; D.g2_fs = V.gu_frame_stride[k0]; D.w2 = V.bw >> k0; D.h2 = V.bh >> k0;
            for (int k = k0 + 1; k <= nb; ++k) { D.g[k] = V.Gu[k]; D.fs[k] = V.gu_frame_stride[k]; }
        }
        // stream order: the k_down1 launches for levels 2 .. k0 - 1 (below) come first
        for (int k = 2; k < k0; ++k) {
            int r = launch_down1_level(s, k, v0, v1, n_frames, st);
            if (r != VSB_OK) return r;
        }
        if (m > 0) k_down_tail<<<dim3(m, 3, n_frames), dim3(DT_TX, DT_TY), smem, st>>>(p);
        ++s->launches;
        prof_stage(s, st, "down_tail", bytes * n_frames);  // level k0 in once, levels k0 + 1 .. nb out once
        return check_launch("k_down_tail");
    }
    for (int k = 2; k < nb; ++k) {
        int r = launch_down1_level(s, k, v0, v1, n_frames, st);
        if (r != VSB_OK) return r;
    }
    return check_launch("k_down1");
}

// K4 + K5
static int launch_back_fast(vsb_stitcher *s, int n_frames, int16_t *const *d_outs, size_t out_pitch, cudaStream_t st)
{
    NvtxRange nvtx("vsb: blend (k_coarse, k_blend_seam, k_blend_int)");
    const int n = s->cfg.num_views, nb = s->nb;
    {
        CoarseParams p;
        std::memset(&p, 0, sizeof(p));
        p.geo = s->cgeo;
        for (int j = 0; j < s->cgeo.nlev; ++j) { p.cw[j] = s->cw[2 + j]; p.ch[j] = s->ch[2 + j]; p.dw[j] = s->dw[2 + j]; }
        p.c2 = s->C2; p.c2_fs = s->c2_frame_stride; p.tile_views = s->d_coarse_views; p.tiles_x = s->coarse_tiles_x;
        p.views = s->d_coarse_desc; p.f0 = s->f0;
        double bytes = 6.0 * s->cw[2] * s->ch[2];  // Gaussian levels >= 2 of every view in once + C2 (s16 x 3) out once
        for (int i = 0; i < n; ++i)
            for (int k = 2; k <= nb; ++k) bytes += 3.0 * (s->v[i].bw >> k) * (s->v[i].bh >> k);
        if (s->ct == 64) k_coarse<64><<<dim3(s->coarse_tiles_x * s->coarse_tiles_y, 3, n_frames), C_THREADS, s->coarse_smem, st>>>(p);
        else k_coarse<32><<<dim3(s->coarse_tiles_x * s->coarse_tiles_y, 3, n_frames), C_THREADS, s->coarse_smem, st>>>(p);
        ++s->launches;
        prof_stage(s, st, "coarse", bytes * n_frames);
    }
    {
        BlendParams p;
        std::memset(&p, 0, sizeof(p));
        p.nb = nb; p.tiles_x = s->blend_tiles_x; p.f0 = s->f0;
        p.cw0 = s->cw[0]; p.ch0 = s->ch[0]; p.cw1 = s->cw[1]; p.ch1 = s->ch[1]; p.cw2 = s->cw[2]; p.ch2 = s->ch[2];
        p.out_w = s->roi_final[2]; p.out_h = s->roi_final[3];
        p.c2 = s->C2; p.c2_fs = s->c2_frame_stride; p.tile_views = s->d_blend_views;
        p.dw0 = s->dw[0]; p.dw1 = s->dw[1];
        double bytes = (s->out_format == VSB_OUT_U8C3 ? 3.0 : 6.0) * s->roi_final[2] * s->roi_final[3] + 6.0 * s->cw[2] * s->ch[2];
        for (int i = 0; i < n; ++i) {
            const View &V = s->v[i];
            BlendView &B = p.v[i];
            B.g0 = V.G0; B.g1 = V.G1; B.g2 = V.G2; B.m0 = V.M0; B.w1 = V.weight[1];
            B.g0_fs = V.g0_frame_stride; B.g1_fs = V.g1_frame_stride; B.g2_fs = V.g2_frame_stride;
            B.x_tl = V.x_tl; B.y_tl = V.y_tl; B.bw = V.bw; B.bh = V.bh;
            bytes += 3.0 * V.bw * V.bh + 3.0 * (V.bw >> 1) * (V.bh >> 1) + 3.0 * (V.bw >> 2) * (V.bh >> 2);
        }
        OutPtrs o;
        std::memset(&o, 0, sizeof(o));
        for (int f = 0; f < n_frames; ++f) o.out[s->f0 + f] = d_outs[f];
        const bool u8 = s->out_format == VSB_OUT_U8C3;
        // the seam tiles first (two to three views each: the long CTAs), the interior tiles fill in behind them
        if (s->n_blend_seam > 0) {
            p.tiles = s->d_blend_lists + s->n_blend_int;
            const dim3 g(s->n_blend_seam, n_frames);
            if (u8) k_blend_seam<true><<<g, BL_THREADS, 0, st>>>(p, o, out_pitch);
            else k_blend_seam<false><<<g, BL_THREADS, 0, st>>>(p, o, out_pitch);
        }
        const double share = (double)s->n_blend_seam / std::max(1, s->n_blend_seam + s->n_blend_int);
        ++s->launches;
        prof_stage(s, st, "blend_seam", bytes * n_frames * share);
        if (s->n_blend_int > 0) {
            p.tiles = s->d_blend_lists;
            const dim3 g(s->n_blend_int, n_frames);
            if (u8) k_blend_int<true><<<g, BL_THREADS, 0, st>>>(p, o, out_pitch);
            else k_blend_int<false><<<g, BL_THREADS, 0, st>>>(p, o, out_pitch);
        }
        ++s->launches;
        prof_stage(s, st, "blend_int", bytes * n_frames * (1.0 - share));  // G0 + G1 + G2 of every view once, C2 once, CV_16SC3 pano out once (split by tile count)
    }
    return check_launch("k_coarse / k_blend");
}

// which form of the remap kernels runs (all bit-identical): VSB_REMAP_VARIANT = -1 forces the coordinate-driven kernels that
// also serve unaligned caller frames; 0 / 1 = table-driven with 4 consecutive pixels per thread / lane-interleaved pixels
// (default, ~3 % faster: fewer cache lines per window load); 2 = lane-interleaved, 2 pixels per thread, 512-thread CTAs (measured
// 4 % slower: the kernel is bound by L1 data-pipe wavefronts, not by latency); 3 = remap #1 gathers from a shared-memory copy of the
// tile's source footprint filled by cp.async (k_remap_stage1_st; measured 3 % slower: the shared-memory gathers of a 4-pixel-per-thread
// mapping take ~3 wavefronts each through bank conflicts, and the fill adds instructions to an issue-bound loop)
static int remap_variant()
{
    static int v = -2;
    if (v < -1) { const char *e = std::getenv("VSB_REMAP_VARIANT"); v = e ? std::max(-1, std::min(3, std::atoi(e))) : 1; }
    return v;
}

// size of the frames the caller hands in for view i: the size the maps address, or the full size when the frames are resized first
static inline int frame_w(const vsb_stitcher *s, int i) { return s->prescale ? s->full_w : s->v[i].src_w; }
static inline int frame_h(const vsb_stitcher *s, int i) { return s->prescale ? s->full_h : s->v[i].src_h; }

// compose_scale != 1: resizes the caller's (or the NV12 stage's) BGR frames of views [v0, v1) into cs_buf and returns pointers to it
static int launch_prescale(vsb_stitcher *s, int v0, int v1, int n_frames, const uint8_t *const *d_srcs, size_t src_pitch, cudaStream_t st,
                           const uint8_t **out_ptrs)
{
    const int n = v1 - v0, nv = s->cfg.num_views;
    for (int i = v0; i < v1; ++i)
        REQ(s->v[i].src_w == s->comp_w && s->v[i].src_h == s->comp_h, VSB_ERR_STATE,
            "compose_scale: the maps of view %d address %dx%d frames, the scaled frame is %dx%d", i, s->v[i].src_w, s->v[i].src_h, s->comp_w, s->comp_h);
    REQ(src_pitch >= (size_t)s->full_w * 3, VSB_ERR_INVALID, "compose_scale: pitch is smaller than a row of the %d-pixel-wide full frame", s->full_w);
    if (!s->cs_buf || s->cs_w != s->comp_w || s->cs_h != s->comp_h) {
        CK(cudaDeviceSynchronize());
        cudaFree(s->cs_buf); s->cs_buf = nullptr;
        s->cs_pitch = align_up((size_t)s->comp_w * 3, 16);
        s->cs_stride = align_up(s->cs_pitch * s->comp_h + 16, 256);
        CK(cudaMalloc(&s->cs_buf, s->cs_stride * nv * s->cfg.max_batch));
        s->cs_w = s->comp_w; s->cs_h = s->comp_h;
    }
    PrescaleParams p;
    std::memset(&p, 0, sizeof(p));
    for (int f = 0; f < n_frames; ++f)
        for (int j = 0; j < n; ++j) {
            p.src[f * n + j] = d_srcs[f * n + j];
            p.dst[f * n + j] = s->cs_buf + s->cs_stride * ((size_t)(s->f0 + f) * nv + v0 + j);
            out_ptrs[f * n + j] = p.dst[f * n + j];
        }
    p.pitch = src_pitch; p.dst_pitch = s->cs_pitch; p.sw = s->full_w; p.sh = s->full_h; p.dw = s->comp_w; p.dh = s->comp_h;
    p.fx = p.fy = static_cast<float>(1.0 / s->compose_scale);
    k_prescale<<<dim3((s->comp_w + 31) / 32, (s->comp_h + 7) / 8, n * n_frames), dim3(32, 8), 0, st>>>(p);
    ++s->launches;
    prof_stage(s, st, "prescale", 3.0 * ((double)s->full_w * s->full_h + (double)s->comp_w * s->comp_h) * n * n_frames);  // full frame in once, scaled frame out once
    return check_launch("k_prescale");
}

// NV12 input: converts the caller's frames of views [v0, v1) into the handle's BGR staging and returns pointers to it
static int launch_nv12(vsb_stitcher *s, int v0, int v1, int n_frames, const uint8_t *const *d_srcs, size_t src_pitch, cudaStream_t st,
                       const uint8_t **bgr_ptrs)
{
    const int n = v1 - v0, nv = s->cfg.num_views;
    const int w = frame_w(s, v0), h = frame_h(s, v0);
    for (int i = v0; i < v1; ++i) REQ(frame_w(s, i) == w && frame_h(s, i) == h, VSB_ERR_INVALID, "NV12 input: all views must share one source size");
    REQ((w & 1) == 0 && (h & 1) == 0, VSB_ERR_INVALID, "NV12 input: source width and height must be even (got %dx%d)", w, h);
    REQ(src_pitch >= (size_t)w, VSB_ERR_INVALID, "NV12 input: pitch is the Y-plane row pitch and must be >= width");
    if (!s->nv_bgr || s->nv_w != w || s->nv_h != h) {
        CK(cudaDeviceSynchronize());
        cudaFree(s->nv_bgr); s->nv_bgr = nullptr;
        s->nv_pitch = align_up((size_t)w * 3, 16);
        s->nv_stride = align_up(s->nv_pitch * h + 16, 256);
        CK(cudaMalloc(&s->nv_bgr, s->nv_stride * nv * s->cfg.max_batch));
        s->nv_w = w; s->nv_h = h;
    }
    Nv12Params p;
    std::memset(&p, 0, sizeof(p));
    for (int f = 0; f < n_frames; ++f)
        for (int j = 0; j < n; ++j) {
            p.src[f * n + j] = d_srcs[f * n + j];
            p.dst[f * n + j] = s->nv_bgr + s->nv_stride * ((size_t)(s->f0 + f) * nv + v0 + j);
            bgr_ptrs[f * n + j] = p.dst[f * n + j];
        }
    p.pitch = src_pitch; p.dst_pitch = s->nv_pitch; p.w = w; p.h = h;
    k_nv12_to_bgr<<<dim3(((w + 3) / 4 + 31) / 32, (h / 2 + 7) / 8, n * n_frames), dim3(32, 8), 0, st>>>(p);
    ++s->launches;
    prof_stage(s, st, "nv12_to_bgr", (1.5 + 3.0) * w * h * n * n_frames);  // NV12 in once, BGR out once
    return check_launch("k_nv12_to_bgr");
}

// (re)builds the remap #1 tap table of view i for the caller's row pitch; stream-ordered before the kernels that read it
static int build_taps1(vsb_stitcher *s, int i, size_t src_pitch, cudaStream_t st, bool nv12 = false)
{
    View &V = s->v[i];
    if (V.t1_src_pitch == src_pitch && V.t1_nv12 == nv12) return VSB_OK;
    if (V.t1_src_pitch != 0) {  // an earlier submission on another stream may still read the old table: order the rebuild after it
        std::lock_guard<std::mutex> lk(s->mu);
        if (s->last_compose_valid) CK(cudaStreamWaitEvent(st, s->last_compose, 0));
    }
    if (!V.t1_off) {
        V.t1_pitch = (int)align_up((size_t)V.roi_w, 4);
        V.t1_plane = (size_t)V.t1_pitch * V.roi_h;
        CK(cudaMalloc(&V.t1_off, V.t1_plane * sizeof(int)));
        CK(cudaMalloc(&V.t1_w, V.t1_plane * 4 * sizeof(float)));
        CK(cudaMalloc(&V.t1_flag, sizeof(int)));
    }
    const dim3 b(32, 8);
    CK(cudaMemsetAsync(V.t1_flag, 0, sizeof(int), st));
    k_build_taps1<<<grid2d(V.t1_pitch, V.roi_h, b), b, 0, st>>>(V.xmap, V.ymap, V.map_pitch, V.roi_w, V.roi_h, V.src_w, V.src_h,
                                                                (unsigned)src_pitch, V.t1_off, V.t1_w, V.t1_plane, V.t1_pitch, nv12 ? 1 : 0, V.t1_flag);
    V.t1_nv12_unsafe = false;
    if (nv12) {  // (calibration-time: the table is rebuilt only when the pitch or the input format changes)
        int flag = 0;
        CK(cudaMemcpyAsync(&flag, V.t1_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        V.t1_nv12_unsafe = flag != 0;
    } else {   // footprint boxes of the staged remap #1 tiles of this view (they follow the table: same maps, same pitch)
        int first = 0;
        for (int j = 0; j < i; ++j) if (front_view(s, j)) first += (int)s->v[j].s1s_tiles.size();
        const int nt = (int)V.s1s_tiles.size();
        TapTable tab;
        tab.off = V.t1_off; tab.w = V.t1_w; tab.plane = V.t1_plane; tab.tab_pitch = V.t1_pitch;
        if (nt > 0) k_s1_boxes<<<nt, 256, 0, st>>>(s->d_s1s_ids + first, s->d_s1s_tiles + first, tab, V.roi_w, V.roi_h, (unsigned)src_pitch, V.src_w, V.src_h);
    }
    V.t1_src_pitch = src_pitch; V.t1_nv12 = nv12;
    return check_launch("k_build_taps1 / k_s1_boxes");
}

// remap stages + pyramid for views [v0, v1) of n_frames frames
enum { FRONT_REMAP = 1, FRONT_PYRAMID = 2, FRONT_ALL = 3 };
static int launch_front(vsb_stitcher *s, int v0, int v1, int n_frames, const uint8_t *const *d_srcs, size_t src_pitch, cudaStream_t st,
                        const uint8_t *warped = nullptr, int stages = FRONT_ALL)
{
    NvtxRange nvtx(stages == FRONT_PYRAMID ? "vsb: feed (pyramid)" : "vsb: feed (remap #1 + gain, remap #2 + border, pyramid)");
    const int n = v1 - v0;
    int ws[MAXV], hs[MAXV];
    int r = sync_tile_lists(s);
    if (r != VSB_OK) return r;
    const uint8_t *bgr_ptrs[MAX_BATCH * MAXV], *cs_ptrs[MAX_BATCH * MAXV];
    bool nv12_fused = false;
    if (!(stages & FRONT_REMAP)) goto pyramid;
    if (!warped && s->in_format == VSB_IN_NV12 && s->prescale) {
        // the resize reads BGR: convert at full size first (the reference's capture thread does, A/networking.cpp:46), then resize
        r = launch_nv12(s, v0, v1, n_frames, d_srcs, src_pitch, st, bgr_ptrs);
        if (r != VSB_OK) return r;
        d_srcs = bgr_ptrs; src_pitch = s->nv_pitch;
    } else if (!warped && s->in_format == VSB_IN_NV12) {
        // fused form: remap #1 converts its taps itself (k_remap_stage1_nv12), no BGR staging image.  Needs 2-byte aligned frames
        // and an even pitch (16-bit chroma loads), 31-bit offsets, and tables free of weights its scaled chain cannot take.
        static const bool want = [] { const char *e = std::getenv("VSB_NV12_FUSED"); return !e || std::atoi(e) != 0; }();
        nv12_fused = want && remap_variant() >= 0 && src_pitch % 2 == 0;
        for (int i = v0; i < v1 && nv12_fused; ++i)
            nv12_fused = s->v[i].src_w >= 2 && s->v[i].src_h >= 2 && s->v[i].src_h % 2 == 0 && (unsigned long long)src_pitch * (s->v[i].src_h * 3 / 2) < 0x7fffffffull;
        for (int j = 0; j < n * n_frames && nv12_fused; ++j) nv12_fused = ((size_t)d_srcs[j] & 1) == 0;
        for (int i = v0; i < v1 && nv12_fused; ++i) {
            if (!front_view(s, i)) continue;
            r = build_taps1(s, i, src_pitch, st, true);
            if (r != VSB_OK) return r;
            nv12_fused = !s->v[i].t1_nv12_unsafe;
        }
        if (!nv12_fused) {
            r = launch_nv12(s, v0, v1, n_frames, d_srcs, src_pitch, st, bgr_ptrs);
            if (r != VSB_OK) return r;
            d_srcs = bgr_ptrs; src_pitch = s->nv_pitch;
        }
    }
    if (!warped && s->prescale) {
        r = launch_prescale(s, v0, v1, n_frames, d_srcs, src_pitch, st, cs_ptrs);
        if (r != VSB_OK) return r;
        d_srcs = cs_ptrs; src_pitch = s->cs_pitch;
    }
    if (!warped && nv12_fused) {
        int first = 0, count = 0;
        double bytes = 0;  // algorithmic: every NV12 source byte once + P once
        for (int i = 0; i < s->cfg.num_views; ++i) {
            const View &V = s->v[i];
            if (!front_view(s, i)) continue;
            if (i < v0) first += (int)V.s1_tiles.size();
            else if (i < v1) { count += (int)V.s1_tiles.size(); bytes += 1.5 * V.src_w * V.src_h + 3.0 * V.roi_w * V.roi_h; }
        }
        Stage1TabParams p;
        std::memset(&p, 0, sizeof(p));
        for (int i = 0; i < s->cfg.num_views; ++i) {
            const View &V = s->v[i];
            Stage1TabView &S = p.v[i];
            S.tab.off = V.t1_off; S.tab.w = V.t1_w; S.tab.plane = V.t1_plane; S.tab.tab_pitch = V.t1_pitch;
            S.xmap = V.xmap; S.ymap = V.ymap; S.P = V.P; S.gain = V.gain;
            S.map_pitch = V.map_pitch; S.p_pitch = V.p_pitch; S.p_frame_stride = V.p_frame_stride;
            S.w = V.roi_w; S.h = V.roi_h; S.src_w = V.src_w; S.src_h = V.src_h;
        }
        p.tiles = s->d_s1_tiles + first;
        p.v0 = v0; p.n_views = n; p.src_pitch = (unsigned)src_pitch; p.n_frames = n_frames; p.f0 = s->f0;
        for (int f = 0; f < n_frames; ++f) for (int j = 0; j < n; ++j) p.src[(s->f0 + f) * n + j] = d_srcs[f * n + j];
        if (count > 0) k_remap_stage1_nv12<<<(unsigned)count, dim3(RM_BX, RM_BY), 0, st>>>(p, s->v[v0].src_h);
        ++s->launches;
        prof_stage(s, st, "remap_stage1_nv12", bytes * n_frames);
    } else if (!warped) {
        // table-driven remap #1 whenever the caller's frames allow aligned 32-bit window loads
        bool tab = remap_variant() >= 0 && src_pitch % 4 == 0;
        for (int i = v0; i < v1 && tab; ++i) tab = s->v[i].src_h >= 2 && (unsigned long long)src_pitch * s->v[i].src_h < 0x7fffffffull;
        for (int j = 0; j < n * n_frames && tab; ++j) tab = ((size_t)d_srcs[j] & 3) == 0;  // (null entries of views another rank owns pass)
        int first = 0, count = 0;
        double bytes = 0;  // algorithmic: every source pixel once + P once
        for (int i = 0; i < s->cfg.num_views; ++i) {
            const View &V = s->v[i];
            if (!front_view(s, i)) continue;
            if (i < v0) first += (int)V.s1_tiles.size();
            else if (i < v1) { count += (int)V.s1_tiles.size(); bytes += 3.0 * V.src_w * V.src_h + 3.0 * V.roi_w * V.roi_h; }
        }
        if (tab) {
            for (int i = v0; i < v1; ++i) {
                if (!front_view(s, i)) continue;
                r = build_taps1(s, i, src_pitch, st);
                if (r != VSB_OK) return r;
            }
            Stage1TabParams p;
            std::memset(&p, 0, sizeof(p));
            for (int i = 0; i < s->cfg.num_views; ++i) {
                const View &V = s->v[i];
                Stage1TabView &S = p.v[i];
                S.tab.off = V.t1_off; S.tab.w = V.t1_w; S.tab.plane = V.t1_plane; S.tab.tab_pitch = V.t1_pitch;
                S.xmap = V.xmap; S.ymap = V.ymap; S.P = V.P; S.gain = V.gain;
                S.map_pitch = V.map_pitch; S.p_pitch = V.p_pitch; S.p_frame_stride = V.p_frame_stride;
                S.w = V.roi_w; S.h = V.roi_h; S.src_w = V.src_w; S.src_h = V.src_h;
            }
            p.tiles = s->d_s1_tiles + first;
            p.v0 = v0; p.n_views = n; p.src_pitch = (unsigned)src_pitch; p.n_frames = n_frames; p.f0 = s->f0;
            for (int f = 0; f < n_frames; ++f) for (int j = 0; j < n; ++j) p.src[(s->f0 + f) * n + j] = d_srcs[f * n + j];
            // staged form: 16-byte aligned frames and rows, box coordinates in 16 bits
            bool staged = remap_variant() == 3 && src_pitch % 16 == 0;
            for (int i = v0; i < v1 && staged; ++i) staged = (size_t)s->v[i].src_w * 3 < 65536 && s->v[i].src_h < 65536;
            for (int j = 0; j < n * n_frames && staged; ++j) staged = ((size_t)d_srcs[j] & 15) == 0;
            if (staged) {
                int sfirst = 0, scount = 0;
                for (int i = 0; i < s->cfg.num_views; ++i) {
                    if (!front_view(s, i)) continue;
                    if (i < v0) sfirst += (int)s->v[i].s1s_tiles.size();
                    else if (i < v1) scount += (int)s->v[i].s1s_tiles.size();
                }
                Stage1StParams ps;
                ps.t = p;
                ps.tiles_s = s->d_s1s_tiles + sfirst;
                if (scount > 0) k_remap_stage1_st<<<(unsigned)scount, 256, 0, st>>>(ps);
            } else if (count > 0) {
                if (remap_variant() == 2) k_remap_stage1_tab_h<<<(unsigned)count, dim3(RM_BX, RM_BY * 2), 0, st>>>(p);
                else if (remap_variant() & 1) k_remap_stage1_tab<true><<<(unsigned)count, dim3(RM_BX, RM_BY), 0, st>>>(p);
                else k_remap_stage1_tab<false><<<(unsigned)count, dim3(RM_BX, RM_BY), 0, st>>>(p);
            }
        } else {
            Stage1Params p;
            std::memset(&p, 0, sizeof(p));
            for (int i = 0; i < s->cfg.num_views; ++i) {
                const View &V = s->v[i];
                Stage1View &S = p.v[i];
                S.xmap = V.xmap; S.ymap = V.ymap; S.P = V.P; S.gain = V.gain;
                S.map_pitch = V.map_pitch; S.p_pitch = V.p_pitch; S.p_frame_stride = V.p_frame_stride;
                S.w = V.roi_w; S.h = V.roi_h; S.src_w = V.src_w; S.src_h = V.src_h;
            }
            p.tiles = s->d_s1_tiles + first;
            p.v0 = v0; p.n_views = n; p.src_pitch = src_pitch; p.f0 = s->f0;
            for (int f = 0; f < n_frames; ++f) for (int j = 0; j < n; ++j) p.src[(s->f0 + f) * n + j] = d_srcs[f * n + j];
            if (count > 0) k_remap_stage1<<<dim3(count, n_frames), dim3(RM_BX, RM_BY), 0, st>>>(p);
        }
        ++s->launches;
        prof_stage(s, st, "remap_stage1", bytes * n_frames);
    }
    {
        int first = 0, count = 0;
        double bytes = 0;  // P once + bordered planar G0 once
        bool tab = remap_variant() >= 0 && s->cfg.enable_local && !warped;
        for (int i = 0; i < s->cfg.num_views; ++i) {
            const View &V = s->v[i];
            if (!front_view(s, i)) continue;
            if (i < v0) first += (int)V.s2_tiles.size();
            else if (i < v1) {
                count += (int)V.s2_tiles.size(); bytes += 3.0 * V.roi_w * V.roi_h + 3.0 * V.bw * V.bh;
                tab = tab && V.mesh_cur >= 0 && V.t2_off[V.mesh_cur] != nullptr && !V.t2_unsafe[V.mesh_cur];
            }
        }
        if (tab) {
            Stage2TabParams p;
            std::memset(&p, 0, sizeof(p));
            for (int i = v0; i < v1; ++i) {
                const View &V = s->v[i];
                Stage2TabView &S = p.v[i];
                S.tab.off = V.t2_off[V.mesh_cur]; S.tab.w = V.t2_w[V.mesh_cur]; S.tab.plane = V.t2_plane; S.tab.tab_pitch = V.t2_pitch;
                S.Pbase = V.P_alloc; S.G0 = V.G0; S.p_frame_stride = V.p_frame_stride; S.g0_frame_stride = V.g0_frame_stride;
                S.p_pitch = (unsigned)V.p_pitch; S.bw = V.bw; S.bh = V.bh;
            }
            p.tiles = s->d_s2_tiles + first; p.n_frames = n_frames; p.f0 = s->f0;
            if (count > 0) {
                k_remap_stage2_tab<<<(unsigned)count, dim3(RM_BX, RM_BY), 0, st>>>(p);
            }
        } else {
            Stage2Params p;
            std::memset(&p, 0, sizeof(p));
            for (int i = 0; i < s->cfg.num_views; ++i) {
                const View &V = s->v[i];
                Stage2View &S = p.v[i];
                S.P = V.P; S.G0 = V.G0;
                if (s->cfg.enable_local && V.mesh_cur >= 0 && !warped) { S.xmesh = V.mesh[V.mesh_cur][0]; S.ymesh = V.mesh[V.mesh_cur][1]; }
                S.p_pitch = V.p_pitch; S.p_frame_stride = V.p_frame_stride;
                S.map_pitch = V.map_pitch; S.g0_frame_stride = V.g0_frame_stride;
                if (warped && i == v0) { S.P = warped; S.p_pitch = src_pitch; S.p_frame_stride = 0; }  // feed_online: the caller's warped view
                S.w = V.roi_w; S.h = V.roi_h; S.bw = V.bw; S.bh = V.bh; S.top = V.top; S.left = V.left;
            }
            p.tiles = s->d_s2_tiles + first; p.f0 = s->f0;
            if (count > 0) k_remap_stage2<<<dim3(count, n_frames), dim3(RM_BX, RM_BY), 0, st>>>(p);
        }
        ++s->launches;
        prof_stage(s, st, "remap_stage2", bytes * n_frames);
    }
pyramid:
    if (!(stages & FRONT_PYRAMID)) return check_launch("front half (remap)");
    if (s->fast) {
        r = launch_down2(s, v0, v1, n_frames, st);
        return r != VSB_OK ? r : launch_down1(s, v0, v1, n_frames, st);
    }
    for (int k = 0; k < s->nb; ++k) {
        PyrParams p;
        std::memset(&p, 0, sizeof(p));
        for (int j = 0; j < n; ++j) {
            const View &V = s->v[v0 + j];
            PyrView &S = p.v[j];
            S.in = k == 0 ? (const void *)V.G0 : (const void *)V.G[k];
            S.in_frame_stride = k == 0 ? V.g0_frame_stride : V.g_frame_stride[k];
            S.out = V.G[k + 1]; S.out_frame_stride = V.g_frame_stride[k + 1];
            S.w = V.bw >> k; S.h = V.bh >> k;
            ws[j] = (S.w + 1) / 2; hs[j] = (S.h + 1) / 2;
        }
        fill_tilemap(p.tm, n, ws, hs, PD_TX, PD_TY);
        const dim3 g(p.tm.start[n], n_frames, 3);
        if (k == 0) k_pyr_down<uint8_t><<<g, dim3(PD_TX, PD_TY), 0, st>>>(p);
        else k_pyr_down<int16_t><<<g, dim3(PD_TX, PD_TY), 0, st>>>(p);
        ++s->launches;
        double bytes = 0;  // level k in once + level k+1 out once
        for (int j = 0; j < n; ++j) {
            const View &V = s->v[v0 + j];
            bytes += 3.0 * (V.bw >> k) * (V.bh >> k) * (k == 0 ? 1 : 2) + 3.0 * (V.bw >> (k + 1)) * (V.bh >> (k + 1)) * 2;
        }
        static const char *names[VSB_MAX_BANDS] = {"pyr_down_0", "pyr_down_1", "pyr_down_2", "pyr_down_3", "pyr_down_4", "pyr_down_5", "pyr_down_6"};
        prof_stage(s, st, names[k], bytes * n_frames);
    }
    return check_launch("front half (remap + pyramid)");
}

static int launch_back(vsb_stitcher *s, int n_frames, int16_t *const *d_outs, size_t out_pitch, cudaStream_t st)
{
    if (s->fast) return launch_back_fast(s, n_frames, d_outs, out_pitch, st);
    OutPtrs o;
    std::memset(&o, 0, sizeof(o));
    for (int f = 0; f < n_frames; ++f) o.out[f] = d_outs[f];
    const dim3 g((s->cw[0] + LG_TW - 1) / LG_TW, (s->ch[0] + LG_TH - 1) / LG_TH, n_frames);
    k_legacy_blend_collapse<<<g, LG_THREADS, 0, st>>>(s->d_plan, o, out_pitch);
    ++s->launches;
    double bytes = 6.0 * s->roi_final[2] * s->roi_final[3];  // every Gaussian level of every view once + the CV_16SC3 pano once
    for (int i = 0; i < s->cfg.num_views; ++i) {
        const View &V = s->v[i];
        bytes += 3.0 * V.bw * V.bh;
        for (int k = 1; k <= s->nb; ++k) bytes += 6.0 * (V.bw >> k) * (V.bh >> k);
    }
    prof_stage(s, st, "blend_collapse", bytes * n_frames);
    return check_launch("k_legacy_blend_collapse");
}

// a front half (vsb_feed*) without its blend yet: later table / map updates must still be ordered after it
static int note_front_done(vsb_stitcher *s, cudaStream_t st, int r)
{
    if (r != VSB_OK) return r;
    std::lock_guard<std::mutex> lk(s->mu);
    CK(cudaEventRecord(s->last_compose, st));
    s->last_compose_valid = true;
    return VSB_OK;
}

static int note_compose_done(vsb_stitcher *s, cudaStream_t st)
{
    s->launches_last = s->launches;
    s->launches = 0;
    std::lock_guard<std::mutex> lk(s->mu);
    CK(cudaEventRecord(s->last_compose, st));
    s->last_compose_valid = true;
    return VSB_OK;
}

// ---- NCCL through dlopen (view-sharded mode) + view ownership -------------------------------------------------------
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommInitRankConfig)(ncclComm_t *, int, ncclUniqueId, int, ncclConfig_t *);  // optional
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
};
const NcclApi *nccl_api()
{
    static NcclApi api;
    static int state = 0;  // 0 = not tried, 1 = ok, -1 = missing
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (state == 0) {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        state = -1;
        if (h) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
            api.CommInitRankConfig = (decltype(api.CommInitRankConfig))dlsym(h, "ncclCommInitRankConfig");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
            api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
            api.Send = (decltype(api.Send))dlsym(h, "ncclSend");
            api.Recv = (decltype(api.Recv))dlsym(h, "ncclRecv");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
            if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send && api.Recv && api.GetErrorString) state = 1;
        }
    }
    return state == 1 ? &api : nullptr;
}
#define NC(expr) do { ncclResult_t _n = (expr); if (_n != ncclSuccess) return fail(VSB_ERR_CUDA, "%s: %s", #expr, nccl->GetErrorString(_n)); } while (0)

// the rank whose canvas strip holds most of the view's level-0 seam weight (same rule on every rank: no exchange needed)
int view_owner(const vsb_stitcher *s, int i, int world)
{
    const View &V = s->v[i];
    std::vector<long long> per(world, 0);
    for (int x = 0; x < V.bw; ++x) per[std::min((V.x_tl + x) / s->strip_w, world - 1)] += V.w0_cols[x];
    int best = 0;
    for (int r = 1; r < world; ++r) if (per[r] > per[best]) best = r;
    return best;
}
}  // namespace vsb

using namespace vsb;

extern "C" {

int vsb_create(const vsb_config *cfg, vsb_stitcher **out)
{
    REQ(cfg && out, VSB_ERR_INVALID, "create: null argument");
    REQ(cfg->num_views >= 1 && cfg->num_views <= VSB_MAX_VIEWS, VSB_ERR_INVALID, "create: num_views must be 1..%d", VSB_MAX_VIEWS);
    REQ(cfg->num_bands >= 0 && cfg->num_bands <= VSB_MAX_BANDS, VSB_ERR_INVALID, "create: num_bands must be 0..%d", VSB_MAX_BANDS);
    REQ(cfg->max_batch >= 1 && cfg->max_batch <= MAX_BATCH, VSB_ERR_INVALID, "create: max_batch must be 1..%d", MAX_BATCH);
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    REQ(ndev > 0, VSB_ERR_CUDA, "create: no CUDA device (this path has no CPU fallback)");
    int dev = cfg->device;
    if (dev < 0) CK(cudaGetDevice(&dev));
    REQ(dev < ndev, VSB_ERR_INVALID, "create: device %d out of range", dev);
    vsb_stitcher *s = new (std::nothrow) vsb_stitcher();
    REQ(s, VSB_ERR_NOMEM, "create: out of host memory");
    s->cfg = *cfg;
    s->device = dev;
    DeviceGuard g(dev);
    cudaError_t e = cudaStreamCreateWithFlags(&s->setup_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->mesh_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->io_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->last_compose, cudaEventDisableTiming);
    if (e != cudaSuccess) { vsb_destroy(s); return check_cuda(e, "create: streams"); }
    *out = s;
    return VSB_OK;
}

int vsb_destroy(vsb_stitcher *s)
{
    if (!s) return VSB_OK;
    DeviceGuard g(s->device);
    cudaDeviceSynchronize();
    if (s->calib_state && s->calib_dtor) s->calib_dtor(s->calib_state);
    for (int i = 0; i < MAXV; ++i) free_view(s->v[i]);
    for (int k = 0; k < MAXL; ++k) cudaFree(s->dw[k]);
    cudaFree(s->d_plan);
    cudaFree(s->d_blend_views); cudaFree(s->d_coarse_views); cudaFree(s->d_down2_tiles); cudaFree(s->C2);
    cudaFree(s->d_coarse_desc); cudaFree(s->d_s1_tiles); cudaFree(s->d_s2_tiles); cudaFree(s->d_blend_lists);
    cudaFree(s->d_s1s_ids); cudaFree(s->d_s1s_tiles);
    for (int i = 0; i < MAXV; ++i) { cudaFree(s->d_send[i]); cudaFree(s->d_recv[i]); }
    cudaFree(s->nv_bgr); cudaFree(s->cs_buf); cudaFree(s->cons_tab);
    for (int d = 0; d < HOST_DEPTH; ++d) { cudaFree(s->stage_src[d]); cudaFree(s->stage_out[d]); cudaFree(s->stage_nv12[d]); if (s->ev_host[d]) cudaEventDestroy(s->ev_host[d]); }
    for (int b = 0; b < 2; ++b)
        for (int p = 0; p < MAXV; ++p) { cudaFree(s->x_send[b][p]); cudaFree(s->x_recv[b][p]); }
    if (s->comm) { const NcclApi *api = nccl_api(); if (api) api->CommDestroy(s->comm); }
    if (s->sh_front) { cudaStreamDestroy(s->sh_front); cudaStreamDestroy(s->sh_comm); cudaStreamDestroy(s->sh_back); cudaEventDestroy(s->ev_call); }
    for (int b = 0; b < 2; ++b) { if (s->ev_packed[b]) cudaEventDestroy(s->ev_packed[b]); if (s->ev_recv[b]) cudaEventDestroy(s->ev_recv[b]); if (s->ev_back[b]) cudaEventDestroy(s->ev_back[b]); }
    if (s->setup_stream) cudaStreamDestroy(s->setup_stream);
    if (s->mesh_stream) cudaStreamDestroy(s->mesh_stream);
    if (s->io_stream) cudaStreamDestroy(s->io_stream);
    if (s->in_stream) cudaStreamDestroy(s->in_stream);
    if (s->out_stream) cudaStreamDestroy(s->out_stream);
    for (int h = 0; h < MAX_SPLIT; ++h) { if (s->sub[h]) cudaStreamDestroy(s->sub[h]); if (s->ev_join[h]) cudaEventDestroy(s->ev_join[h]); if (s->ev_forks[h]) cudaEventDestroy(s->ev_forks[h]); }
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    for (int f = 0; f < MAX_BATCH; ++f) { if (s->ev_in[f]) cudaEventDestroy(s->ev_in[f]); if (s->ev_done[f]) cudaEventDestroy(s->ev_done[f]); }
    if (s->last_compose) cudaEventDestroy(s->last_compose);
    for (int i = 0; i <= VSB_MAX_STAGES; ++i) if (s->prof_ev[i]) cudaEventDestroy(s->prof_ev[i]);
    cudaGetLastError();
    delete s;
    return VSB_OK;
}

// Blender::prepare(corners, sizes) -> resultRoi (sources/modules/stitching/src/util.cpp:125-138) ->
// MultiBandBlender::prepare(Rect) (sources/modules/stitching/src/blenders.cpp:237-274)
int vsb_prepare(vsb_stitcher *s, const int *corners_xy, const int *sizes_wh)
{
    REQ(s && corners_xy && sizes_wh, VSB_ERR_INVALID, "prepare: null argument");
    DeviceGuard g(s->device);
    const int n = s->cfg.num_views;
    int tlx = INT32_MAX, tly = INT32_MAX, brx = INT32_MIN, bry = INT32_MIN;
    for (int i = 0; i < n; ++i) {
        REQ(sizes_wh[2 * i] > 0 && sizes_wh[2 * i + 1] > 0, VSB_ERR_INVALID, "prepare: view %d has an empty size", i);
        tlx = std::min(tlx, corners_xy[2 * i]); tly = std::min(tly, corners_xy[2 * i + 1]);
        brx = std::max(brx, corners_xy[2 * i] + sizes_wh[2 * i]); bry = std::max(bry, corners_xy[2 * i + 1] + sizes_wh[2 * i + 1]);
    }
    int W = brx - tlx, H = bry - tly;
    s->roi_final[0] = tlx; s->roi_final[1] = tly; s->roi_final[2] = W; s->roi_final[3] = H;
    const double max_len = (double)std::max(W, H);
    s->nb = std::min(s->cfg.num_bands, (int)std::ceil(std::log(max_len) / std::log(2.0)));
    const int m = 1 << s->nb;
    W += (m - W % m) % m; H += (m - H % m) % m;
    s->roi[0] = tlx; s->roi[1] = tly; s->roi[2] = W; s->roi[3] = H;
    CK(cudaDeviceSynchronize());
    for (int i = 0; i < MAXV; ++i) free_view(s->v[i]);
    for (int k = 0; k < MAXL; ++k) { cudaFree(s->dw[k]); s->dw[k] = nullptr; }
    for (int k = 0; k <= s->nb; ++k) {
        s->cw[k] = k == 0 ? W : (s->cw[k - 1] + 1) / 2;
        s->ch[k] = k == 0 ? H : (s->ch[k - 1] + 1) / 2;
        CK(cudaMalloc(&s->dw[k], sizeof(float) * s->cw[k] * s->ch[k]));
    }
    for (int d = 0; d < HOST_DEPTH; ++d) {  // sized per calibration
        cudaFree(s->stage_src[d]); cudaFree(s->stage_nv12[d]); cudaFree(s->stage_out[d]);
        s->stage_src[d] = s->stage_nv12[d] = nullptr; s->stage_out[d] = nullptr;
    }
    s->stage_src_w = s->stage_src_h = 0; s->host_pending = 0;
    s->cons_w = s->cons_ih = 0;  // the consumer's resize tables depend on the panorama size
    s->prescale = false; s->compose_scale = 1.0;  // a new geometry starts at compose_scale 1 (vsb_set_compose_scale after the maps)
    s->views_inited = 0; s->prepared = true; s->finalized = false;
    return VSB_OK;
}

int vsb_get_roi(const vsb_stitcher *s, int roi_final[4], int roi_padded[4], int *num_bands)
{
    REQ(s && s->prepared, VSB_ERR_STATE, "get_roi: call vsb_prepare first");
    if (roi_final) std::memcpy(roi_final, s->roi_final, sizeof(int) * 4);
    if (roi_padded) std::memcpy(roi_padded, s->roi, sizeof(int) * 4);
    if (num_bands) *num_bands = s->nb;
    return VSB_OK;
}

// MultiBandBlender::init_gpu (sources/modules/stitching/src/blenders.cpp:344-434)
int vsb_init_view(vsb_stitcher *s, int i, const uint8_t *mask, int mw, int mh, size_t pitch, int tl_x, int tl_y, int on_device)
{
    REQ(s && mask, VSB_ERR_INVALID, "init_view: null argument");
    REQ(s->prepared, VSB_ERR_STATE, "init_view: call vsb_prepare first");
    REQ(i == s->views_inited && i < s->cfg.num_views, VSB_ERR_INVALID, "init_view: views must be initialised in order (expected %d, got %d)", s->views_inited, i);
    REQ(mw > 0 && mh > 0 && pitch >= (size_t)mw, VSB_ERR_INVALID, "init_view: bad mask size");
    DeviceGuard g(s->device);
    View &V = s->v[i];
    const int nb = s->nb, m = 1 << nb;
    const int rx = s->roi[0], ry = s->roi[1], rbx = s->roi[0] + s->roi[2], rby = s->roi[1] + s->roi[3];
    const int gap = 3 * m;
    int tnx = std::max(rx, tl_x - gap), tny = std::max(ry, tl_y - gap);
    int bnx = std::min(rbx, tl_x + mw + gap), bny = std::min(rby, tl_y + mh + gap);
    tnx = rx + (((tnx - rx) >> nb) << nb);
    tny = ry + (((tny - ry) >> nb) << nb);
    int width = bnx - tnx, height = bny - tny;
    width += (m - width % m) % m;
    height += (m - height % m) % m;
    bnx = tnx + width; bny = tny + height;
    const int dy = std::max(bny - rby, 0), dx = std::max(bnx - rbx, 0);
    tnx -= dx; bnx -= dx; tny -= dy; bny -= dy;
    V.top = tl_y - tny; V.left = tl_x - tnx; V.bottom = bny - tl_y - mh; V.right = bnx - tl_x - mw;
    REQ(V.top >= 0 && V.left >= 0 && V.bottom >= 0 && V.right >= 0, VSB_ERR_INVALID, "init_view: view %d lies outside the prepared roi", i);
    V.y_tl = tny - ry; V.y_br = bny - ry; V.x_tl = tnx - rx; V.x_br = bnx - rx;
    V.bw = width; V.bh = height; V.roi_w = mw; V.roi_h = mh; V.tl_x = tl_x; V.tl_y = tl_y;

    // static weight pyramid
    uint8_t *d_mask = nullptr;
    size_t d_pitch = pitch;
    if (!on_device) {
        CK(cudaMalloc(&d_mask, (size_t)mw * mh));
        // stream-ordered: a blocking cudaMemcpy2D from pageable memory may return before its DMA lands, and the
        // non-blocking setup stream does not serialise with the NULL stream
        CK(cudaMemcpy2DAsync(d_mask, mw, mask, pitch, mw, mh, cudaMemcpyHostToDevice, s->setup_stream));
        d_pitch = mw;
    }
    const dim3 b(32, 8);
    CK(cudaMalloc(&V.weight[0], sizeof(float) * width * height));
    CK(cudaMalloc(&V.M0, (size_t)width * height));
    k_weight_level0<<<grid2d(width, height, b), b, 0, s->setup_stream>>>(on_device ? mask : d_mask, mw, mh, d_pitch, V.top, V.left, V.weight[0], V.M0, width, height);
    for (int k = 0; k < nb; ++k) {
        const int w = width >> k, h = height >> k;
        CK(cudaMalloc(&V.weight[k + 1], sizeof(float) * (w / 2) * (h / 2)));
        int r = vsb_pyr_down_f32(V.weight[k], w, h, (size_t)w * 4, V.weight[k + 1], (size_t)(w / 2) * 4, s->setup_stream);
        if (r != VSB_OK) return r;
    }
    CK(cudaStreamSynchronize(s->setup_stream));
    cudaFree(d_mask);
    int r = check_launch("init_view weights");
    if (r != VSB_OK) return r;

    // per-frame buffers of this view
    const int F = s->cfg.max_batch;
    V.p_pitch = align_up((size_t)(mw + 5) * 3 + 8, 16);
    V.p_frame_stride = align_up(V.p_pitch * (mh + 2) + 16, 16);
    V.p_origin = (unsigned)(V.p_pitch + 12);
    CK(cudaMalloc(&V.P_alloc, V.p_frame_stride * F));
    V.P = V.P_alloc + V.p_origin;
    V.g0_frame_stride = (size_t)3 * width * height;
    CK(cudaMalloc(&V.G0, V.g0_frame_stride * F));
    make_g0_map(V, F);
    if (nb >= 3) {  // fast path keeps only G0 and G2 (u8)
        V.g2_frame_stride = (size_t)3 * (width >> 2) * (height >> 2);
        CK(cudaMalloc(&V.G2, V.g2_frame_stride * F));
        CK(cudaMemsetAsync(V.G2, 0, V.g2_frame_stride * F, s->setup_stream));  // skipped tiles stay defined
        V.g1_frame_stride = (size_t)3 * (width >> 1) * (height >> 1);
        CK(cudaMalloc(&V.G1, V.g1_frame_stride * F));
        CK(cudaMemsetAsync(V.G1, 0, V.g1_frame_stride * F, s->setup_stream));
        V.Gu[2] = V.G2; V.gu_frame_stride[2] = V.g2_frame_stride;
        for (int k = 3; k <= nb; ++k) {
            V.gu_frame_stride[k] = (size_t)3 * (width >> k) * (height >> k);
            CK(cudaMalloc(&V.Gu[k], V.gu_frame_stride[k] * F));
        }
    } else {
        for (int k = 1; k <= nb; ++k) {
            V.g_frame_stride[k] = (size_t)3 * (width >> k) * (height >> k);
            CK(cudaMalloc(&V.G[k], sizeof(int16_t) * V.g_frame_stride[k] * F));
        }
    }
    V.map_pitch = align_up((size_t)mw * 4, 16);
    CK(cudaEventCreateWithFlags(&V.mesh_ready, cudaEventDisableTiming));
    CK(cudaMemsetAsync(V.P_alloc, 0, V.p_frame_stride * F, s->setup_stream));  // tiles outside the camera frame and the zero frame are never written
    CK(cudaStreamSynchronize(s->setup_stream));
    all_tiles(i, V.bw, V.bh, V.s2_tiles);   // narrowed by build_fast_plan once every view is known
    V.s1_tiles.clear();                      // filled by vsb_set_maps
    s->tiles_dirty = true;
    V.inited = true;
    s->views_inited++;
    if (s->views_inited == s->cfg.num_views) return finalize(s);
    return VSB_OK;
}

int vsb_get_view_geometry(const vsb_stitcher *s, int i, int out8[8])
{
    REQ(s && out8 && i >= 0 && i < s->cfg.num_views && s->v[i].inited, VSB_ERR_INVALID, "get_view_geometry: view not initialised");
    const View &V = s->v[i];
    const int g[8] = {V.top, V.bottom, V.left, V.right, V.x_tl, V.y_tl, V.x_br, V.y_br};
    std::memcpy(out8, g, sizeof(g));
    return VSB_OK;
}

int vsb_set_maps(vsb_stitcher *s, int i, const float *xmap, const float *ymap, int w, int h, size_t pitch, int on_device, int src_w, int src_h)
{
    REQ(s && xmap && ymap, VSB_ERR_INVALID, "set_maps: null argument");
    REQ(i >= 0 && i < s->cfg.num_views && s->v[i].inited, VSB_ERR_STATE, "set_maps: view %d is not initialised", i);
    View &V = s->v[i];
    REQ(w == V.roi_w && h == V.roi_h, VSB_ERR_INVALID, "set_maps: maps are %dx%d but the view's mask is %dx%d", w, h, V.roi_w, V.roi_h);
    REQ(src_w >= 2 && src_h >= 1 && pitch >= (size_t)w * 4, VSB_ERR_INVALID, "set_maps: bad sizes");
    DeviceGuard g(s->device);
    {   // frames in flight may still read the old maps / P: the copies below are stream-ordered after them (no device-wide sync)
        std::lock_guard<std::mutex> lk(s->mu);
        if (s->last_compose_valid) CK(cudaStreamWaitEvent(s->setup_stream, s->last_compose, 0));
    }
    if (!V.xmap) { CK(cudaMalloc(&V.xmap, V.map_pitch * h)); CK(cudaMalloc(&V.ymap, V.map_pitch * h)); }
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CK(cudaMemcpy2DAsync(V.xmap, V.map_pitch, xmap, pitch, (size_t)w * 4, h, kind, s->setup_stream));
    CK(cudaMemcpy2DAsync(V.ymap, V.map_pitch, ymap, pitch, (size_t)w * 4, h, kind, s->setup_stream));
    CK(cudaMemsetAsync(V.P_alloc, 0, V.p_frame_stride * s->cfg.max_batch, s->setup_stream));
    CK(cudaStreamSynchronize(s->setup_stream));
    V.t1_src_pitch = 0;  // the remap #1 tap table follows the maps: rebuilt by the next compose
    // static tile table of remap #1: tiles in which at least one pixel has a tap inside the camera frame
    std::vector<float> hx((size_t)w * h), hy((size_t)w * h);
    CK(cudaMemcpy2D(hx.data(), (size_t)w * 4, V.xmap, V.map_pitch, (size_t)w * 4, h, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy2D(hy.data(), (size_t)w * 4, V.ymap, V.map_pitch, (size_t)w * 4, h, cudaMemcpyDeviceToHost));
    V.s1_tiles.clear();
    const int tw = RM_BX * RM_PX;
    for (int ty = 0; ty < (h + RM_BY - 1) / RM_BY; ++ty)
        for (int tx = 0; tx < (w + tw - 1) / tw; ++tx) {
            bool any = false;
            for (int y = ty * RM_BY; y < std::min(h, (ty + 1) * RM_BY) && !any; ++y)
                for (int x = tx * tw; x < std::min(w, (tx + 1) * tw); ++x) {
                    const float fx = hx[(size_t)y * w + x], fy = hy[(size_t)y * w + x];
                    if (fx >= -1.f && fx < (float)src_w && fy >= -1.f && fy < (float)src_h) { any = true; break; }
                }
            if (any) V.s1_tiles.push_back((uint32_t)i | ((uint32_t)tx << 8) | ((uint32_t)ty << 20));
        }
    V.s1s_tiles.clear();
    for (int ty = 0; ty < (h + S1S_T - 1) / S1S_T; ++ty)
        for (int tx = 0; tx < (w + S1S_T - 1) / S1S_T; ++tx) {
            bool any = false;
            for (int y = ty * S1S_T; y < std::min(h, (ty + 1) * S1S_T) && !any; ++y)
                for (int x = tx * S1S_T; x < std::min(w, (tx + 1) * S1S_T); ++x) {
                    const float fx = hx[(size_t)y * w + x], fy = hy[(size_t)y * w + x];
                    if (fx >= -1.f && fx < (float)src_w && fy >= -1.f && fy < (float)src_h) { any = true; break; }
                }
            if (any) V.s1s_tiles.push_back((uint32_t)i | ((uint32_t)tx << 8) | ((uint32_t)ty << 20));
        }
    s->tiles_dirty = true;
    V.src_w = src_w; V.src_h = src_h; V.has_maps = true;
    return VSB_OK;
}

int vsb_set_gain(vsb_stitcher *s, int i, float gain)
{
    REQ(s && i >= 0 && i < s->cfg.num_views && s->v[i].inited, VSB_ERR_STATE, "set_gain: view %d is not initialised", i);
    REQ(gain >= 0.f && gain < 1e6f, VSB_ERR_INVALID, "set_gain: gain must be a finite non-negative number");
    s->v[i].gain = gain;  // travels by value in the kernel parameters of every later launch: nothing in flight reads it
    return VSB_OK;
}

// MeshWarper::convertMeshesToMap for one view (360_stitcher/meshwarper.cpp:823-886), entirely on the device
int vsb_set_mesh(vsb_stitcher *s, int i, const float *mesh_x, const float *mesh_y, int rows, int cols)
{
    NvtxRange nvtx("vsb: convertMeshesToMap (vsb_set_mesh)");
    REQ(s && mesh_x && mesh_y, VSB_ERR_INVALID, "set_mesh: null argument");
    REQ(i >= 0 && i < s->cfg.num_views && s->v[i].inited, VSB_ERR_STATE, "set_mesh: view %d is not initialised", i);
    REQ(rows >= 2 && cols >= 2 && rows * cols <= 4096, VSB_ERR_INVALID, "set_mesh: mesh must be between 2x2 and 4096 vertices");
    DeviceGuard g(s->device);
    View &V = s->v[i];
    // a window view (split calibration) takes the mesh of its CAMERA: the splat and the half-resolution table are those of the
    // full-width image, only the window's columns of the map are written
    const bool win = V.win_full_w > 0;
    const int W = win ? V.win_full_w : V.roi_w, H = V.roi_h, hw = W / 2, hh = H / 2;
    REQ(hw >= 2 && hh >= 2, VSB_ERR_INVALID, "set_mesh: view too small");
    cudaStream_t st = s->mesh_stream;
    int target;
    {
        std::lock_guard<std::mutex> lk(s->mu);
        // write into the buffer the compose path is NOT reading; if a previous publication was never adopted, reuse it
        target = V.mesh_pending >= 0 ? V.mesh_pending : (V.mesh_cur < 0 ? 0 : 1 - V.mesh_cur);
        V.mesh_pending = -1;
        // frames already submitted may still read `target` (it was `cur` two publications ago): order after them
        if (s->last_compose_valid) CK(cudaStreamWaitEvent(st, s->last_compose, 0));
    }
    const size_t half = (size_t)hw * hh;
    // scratch: half-resolution accumulators + the vertex mesh.  A window view's accumulators are those of the panorama-wide image
    // it was cut from, which the split calibration exists not to keep: they come from the stream-ordered pool for this call only
    float *scratch = V.mesh_scratch;
    if (win) CK(cudaMallocAsync((void **)&scratch, sizeof(float) * (3 * half + 2 * 4096 + 1), st));
    else if (!scratch) { CK(cudaMalloc(&V.mesh_scratch, sizeof(float) * (3 * half + 2 * 4096 + 1))); scratch = V.mesh_scratch; }
    struct PoolScratch { float *p; cudaStream_t st; ~PoolScratch() { if (p) cudaFreeAsync(p, st); } } pool_scratch{win ? scratch : nullptr, st};
    for (int c = 0; c < 2; ++c)
        if (!V.mesh[target][c]) CK(cudaMalloc(&V.mesh[target][c], V.map_pitch * H));
    float *sum_x = scratch, *sum_y = sum_x + half, *cnt = sum_y + half, *d_mx = cnt + half, *d_my = d_mx + 4096;
    int *d_unsafe = (int *)(d_my + 4096);
    CK(cudaMemcpyAsync(d_mx, mesh_x, sizeof(float) * rows * cols, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_my, mesh_y, sizeof(float) * rows * cols, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(sum_x, 0, sizeof(float) * 3 * half, st));
    CK(cudaMemsetAsync(d_unsafe, 0, sizeof(int), st));
    const dim3 b(32, 8);
    k_mesh_splat<<<grid2d(W, H, b), b, 0, st>>>(d_mx, d_my, rows, cols, W, H, sum_x, sum_y, cnt);
    k_mesh_divide<<<(unsigned)((half + 255) / 256), 256, 0, st>>>(sum_x, sum_y, cnt, (int)half);
    if (win) k_mesh_upsample_win<<<grid2d(V.roi_w, H, b), b, 0, st>>>(sum_x, sum_y, hw, hh, W, H, V.win_x0, V.roi_w, V.mesh[target][0], V.mesh[target][1], V.map_pitch);
    else k_mesh_upsample<<<grid2d(W, H, b), b, 0, st>>>(sum_x, sum_y, hw, hh, W, H, V.mesh[target][0], V.mesh[target][1], V.map_pitch);
    {   // tap table of remap #2 for this mesh buffer (REFLECT border resolved, offsets into the zero-framed P)
        if (!V.t2_off[target]) {
            V.t2_pitch = (int)align_up((size_t)V.bw, 4);
            V.t2_plane = (size_t)V.t2_pitch * V.bh;
            CK(cudaMalloc(&V.t2_off[target], V.t2_plane * sizeof(int)));
            CK(cudaMalloc(&V.t2_w[target], V.t2_plane * 4 * sizeof(float)));
        }
        k_build_taps2<<<grid2d(V.t2_pitch, V.bh, b), b, 0, st>>>(V.mesh[target][0], V.mesh[target][1], V.map_pitch, V.roi_w, H, V.bw, V.bh, V.top, V.left,
                                                                 (unsigned)V.p_pitch, V.p_origin, V.t2_off[target], V.t2_w[target], V.t2_plane, V.t2_pitch, d_unsafe);
    }
    int r = check_launch("set_mesh kernels");
    if (r != VSB_OK) return r;
    int h_unsafe = 0;
    CK(cudaMemcpyAsync(&h_unsafe, d_unsafe, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(V.mesh_ready, st));
    CK(cudaStreamSynchronize(st));  // host mesh arrays may be reused by the caller after return (pageable H2D)
    {
        std::lock_guard<std::mutex> lk(s->mu);
        V.t2_unsafe[target] = h_unsafe != 0;
        V.mesh_pending = target;
    }
    return VSB_OK;
}

int vsb_set_formats(vsb_stitcher *s, int input_format, int output_format)
{
    REQ(s, VSB_ERR_INVALID, "set_formats: null handle");
    REQ(input_format == VSB_IN_BGR8 || input_format == VSB_IN_NV12, VSB_ERR_INVALID, "set_formats: unknown input format %d", input_format);
    REQ(output_format == VSB_OUT_S16C3 || output_format == VSB_OUT_U8C3, VSB_ERR_INVALID, "set_formats: unknown output format %d", output_format);
    REQ(output_format == VSB_OUT_S16C3 || !s->finalized || s->fast, VSB_ERR_STATE, "set_formats: CV_8UC3 output needs num_bands >= 3");
    DeviceGuard g(s->device);
    CK(cudaDeviceSynchronize());
    s->in_format = input_format; s->out_format = output_format;
    return VSB_OK;
}

int vsb_nv12_to_bgr(const uint8_t *d_nv12, int w, int h, size_t pitch, uint8_t *d_bgr, size_t bgr_pitch, void *stream)
{
    REQ(d_nv12 && d_bgr, VSB_ERR_INVALID, "nv12_to_bgr: null argument");
    REQ(w > 0 && h > 0 && (w & 1) == 0 && (h & 1) == 0 && pitch >= (size_t)w && bgr_pitch >= (size_t)w * 3, VSB_ERR_INVALID, "nv12_to_bgr: bad sizes");
    Nv12Params p;
    std::memset(&p, 0, sizeof(p));
    p.src[0] = d_nv12; p.dst[0] = d_bgr; p.pitch = pitch; p.dst_pitch = bgr_pitch; p.w = w; p.h = h;
    k_nv12_to_bgr<<<dim3(((w + 3) / 4 + 31) / 32, (h / 2 + 7) / 8, 1), dim3(32, 8), 0, (cudaStream_t)stream>>>(p);
    return check_launch("k_nv12_to_bgr");
}

// coefficient tables of cv::hal::resize, INTER_LINEAR on 8U (sources/modules/imgproc/src/resize.cpp:3933-3958, 3991-4016)
static void consumer_tables(int ssize, int dsize, bool clamp_x, int *ofs, int *coef)
{
    const double inv_scale = (double)dsize / ssize, scale = 1. / inv_scale;
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int sidx = (int)floorf(f);
        f -= sidx;
        if (clamp_x) {
            if (sidx < 0) { f = 0; sidx = 0; }
            if (sidx >= ssize - 1) { f = 0; sidx = ssize - 1; }
        }
        ofs[d] = sidx;
        const float c0 = 1.f - f, c1 = f;
        const long r0 = std::min(32767L, std::max(-32768L, lrintf(c0 * 2048))), r1 = std::min(32767L, std::max(-32768L, lrintf(c1 * 2048)));
        coef[d] = (int)((unsigned)(r0 & 0xffff) | ((unsigned)(r1 & 0xffff) << 16));
    }
}

int vsb_consumer_image_height(int src_w, int src_h, int out_w, int out_h, int keep_aspect)
{
    if (!keep_aspect || src_w <= 0) return out_h;
    const int ih = (int)((double)out_w / (double)src_w * src_h + 0.5);  // 360_stitcher/timed.cpp:260
    return std::min(ih, out_h);
}

int vsb_consume(vsb_stitcher *s, const uint8_t *d_pano_u8, size_t pitch, int out_w, int out_h, int keep_aspect, int format,
                uint8_t *d_out, size_t out_pitch, void *stream)
{
    REQ(s && d_pano_u8 && d_out, VSB_ERR_INVALID, "consume: null argument");
    REQ(s->prepared, VSB_ERR_STATE, "consume: call vsb_prepare first");
    REQ(format == VSB_CONSUME_RGB || format == VSB_CONSUME_I420, VSB_ERR_INVALID, "consume: unknown format %d", format);
    const int sw = s->roi_final[2], sh = s->roi_final[3];
    REQ(out_w > 0 && out_h > 0 && pitch >= (size_t)sw * 3, VSB_ERR_INVALID, "consume: bad sizes");
    REQ(format != VSB_CONSUME_I420 || ((out_w | out_h) & 1) == 0, VSB_ERR_INVALID, "consume: I420 needs even output sizes");
    REQ(format != VSB_CONSUME_RGB || out_pitch >= (size_t)out_w * 3, VSB_ERR_INVALID, "consume: output pitch too small");
    DeviceGuard g(s->device);
    const int ih = vsb_consumer_image_height(sw, sh, out_w, out_h, keep_aspect);
    REQ(ih > 0, VSB_ERR_INVALID, "consume: empty image");
    cudaStream_t st = (cudaStream_t)stream;
    if (!s->cons_tab || s->cons_w != out_w || s->cons_ih != ih) {
        CK(cudaDeviceSynchronize());
        cudaFree(s->cons_tab); s->cons_tab = nullptr;
        std::vector<int> tab((size_t)2 * (out_w + ih));
        consumer_tables(sw, out_w, true, tab.data(), tab.data() + out_w + ih);
        consumer_tables(sh, ih, false, tab.data() + out_w, tab.data() + 2 * out_w + ih);
        CK(cudaMalloc(&s->cons_tab, tab.size() * sizeof(int)));
        CK(cudaMemcpy(s->cons_tab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice));
        s->cons_w = out_w; s->cons_ih = ih;
    }
    ConsumeParams p;
    std::memset(&p, 0, sizeof(p));
    p.src = d_pano_u8; p.src_pitch = pitch; p.sw = sw; p.sh = sh;
    p.xofs = s->cons_tab; p.yofs = s->cons_tab + out_w; p.ia = s->cons_tab + out_w + ih; p.ib = s->cons_tab + 2 * out_w + ih;
    p.out_w = out_w; p.out_h = out_h; p.ih = ih; p.row0 = out_h / 2 - ih / 2; p.dst = d_out; p.dst_pitch = out_pitch;
    if (format == VSB_CONSUME_RGB) k_consume_rgb<<<dim3((out_w + 31) / 32, (ih + 7) / 8), dim3(32, 8), 0, st>>>(p);
    else k_consume_i420<<<dim3((out_w / 2 + 31) / 32, (out_h / 2 + 7) / 8), dim3(32, 8), 0, st>>>(p);
    return check_launch("k_consume");
}

int vsb_feed(vsb_stitcher *s, int i, const uint8_t *d_bgr, size_t pitch, void *stream)
{
    REQ(s && d_bgr, VSB_ERR_INVALID, "feed: null argument");
    REQ(i >= 0 && i < s->cfg.num_views, VSB_ERR_INVALID, "feed: view index out of range");
    int r = ready_for_frames(s);
    if (r != VSB_OK) return r;
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (s->launches == 0) prof_begin(s, st);
    r = adopt_meshes(s, st);
    if (r != VSB_OK) return r;
    const uint8_t *srcs[1] = {d_bgr};
    return note_front_done(s, st, launch_front(s, i, i + 1, 1, srcs, pitch, st));
}

// MultiBandBlender::feed_online(gpu_img, img_num, stream) itself (sources/modules/stitching/src/blenders.cpp:700-749):
// the caller has already warped the view (its own cuda::remap calls); border + pyramid + weighted add from there.
int vsb_feed_warped(vsb_stitcher *s, int i, const uint8_t *d_warped, size_t pitch, void *stream)
{
    REQ(s && d_warped, VSB_ERR_INVALID, "feed_warped: null argument");
    REQ(i >= 0 && i < s->cfg.num_views, VSB_ERR_INVALID, "feed_warped: view index out of range");
    REQ(s->finalized, VSB_ERR_STATE, "feed_warped: prepare + init_view for every view must come first");
    REQ(pitch >= (size_t)s->v[i].roi_w * 3, VSB_ERR_INVALID, "feed_warped: pitch too small for a %d-pixel-wide CV_8UC3 view", s->v[i].roi_w);
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (s->launches == 0) prof_begin(s, st);
    return note_front_done(s, st, launch_front(s, i, i + 1, 1, nullptr, pitch, st, d_warped));
}

int vsb_blend(vsb_stitcher *s, int16_t *d_out, size_t out_pitch, void *stream)
{
    REQ(s && d_out, VSB_ERR_INVALID, "blend: null argument");
    REQ(s->finalized, VSB_ERR_STATE, "blend: prepare + init_view for every view must come first");
    int r;
    REQ(s->out_format == VSB_OUT_S16C3 || s->fast, VSB_ERR_STATE, "blend: CV_8UC3 output needs num_bands >= 3");
    REQ(out_pitch >= (size_t)s->roi_final[2] * (s->out_format == VSB_OUT_U8C3 ? 3 : 6), VSB_ERR_INVALID, "blend: output pitch too small");
    DeviceGuard g(s->device);
    int16_t *outs[1] = {d_out};
    r = launch_back(s, 1, outs, out_pitch, (cudaStream_t)stream);
    if (r != VSB_OK) return r;
    return note_compose_done(s, (cudaStream_t)stream);
}

int vsb_compose(vsb_stitcher *s, int n_frames, const uint8_t *const *d_srcs, size_t src_pitch, int16_t *const *d_outs, size_t out_pitch, void *stream)
{
    NvtxRange nvtx("vsb_compose");
    REQ(s && d_srcs && d_outs, VSB_ERR_INVALID, "compose: null argument");
    REQ(n_frames >= 1 && n_frames <= s->cfg.max_batch, VSB_ERR_INVALID, "compose: n_frames must be 1..max_batch (%d)", s->cfg.max_batch);
    REQ(s->shard_rank < 0, VSB_ERR_STATE, "compose: handle is view-sharded; feed the owned views, exchange, then blend");
    REQ(s->out_format == VSB_OUT_S16C3 || s->fast, VSB_ERR_STATE, "compose: CV_8UC3 output needs num_bands >= 3");
    REQ(out_pitch >= (size_t)s->roi_final[2] * (s->out_format == VSB_OUT_U8C3 ? 3 : 6), VSB_ERR_INVALID, "compose: output pitch too small");
    int r = ready_for_frames(s);
    if (r != VSB_OK) return r;
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    s->launches = 0;
    r = adopt_meshes(s, st);
    if (r != VSB_OK) return r;
    // VSB_SPLIT = number of sub-batches of the back half (default 2; 1 = everything on the caller's stream)
    static const int n_split = [] { const char *e = std::getenv("VSB_SPLIT"); return e ? std::max(1, std::min(MAX_SPLIT, std::atoi(e))) : (std::getenv("VSB_NO_SPLIT") ? 1 : 2); }();
    if (n_frames >= 4 && s->fast && !s->profiling && n_split > 1) {
        // The remap kernels run ONCE for the whole submission on the caller's stream (their tap tables -- 20 B per pixel -- are
        // read once and stay in registers while all frames stream through).  From the pyramid on, the submission continues as
        // sub-batches on internal streams: the short kernels (k_down_tail: 144 CTAs, k_coarse: ~1.2 waves) and the last partial
        // wave of every kernel of one sub-batch overlap with the other's work.  Frame slots are disjoint.
        const int n = s->cfg.num_views, ns = std::min(n_split, n_frames / 2);
        if (!s->sub[0]) {
            for (int h = 0; h < MAX_SPLIT; ++h) {
                CK(cudaStreamCreateWithFlags(&s->sub[h], cudaStreamNonBlocking));
                CK(cudaEventCreateWithFlags(&s->ev_join[h], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&s->ev_forks[h], cudaEventDisableTiming));
            }
            CK(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
        }
        // VSB_STAGGER=1: the remap launches are split per sub-batch too and issued back to back on the caller's stream, so the
        // remaps of sub-batch h + 1 overlap the pyramid / blend of sub-batch h (at the price of reading the tap tables once per
        // sub-batch).  Default: one remap launch pair for the whole submission.
        static const bool stagger = [] { const char *e = std::getenv("VSB_STAGGER"); return e && std::atoi(e) != 0; }();
        if (!stagger) {
            s->f0 = 0;
            r = launch_front(s, 0, n, n_frames, d_srcs, src_pitch, st, nullptr, FRONT_REMAP);
            if (r != VSB_OK) return r;
            CK(cudaEventRecord(s->ev_fork, st));
        }
        for (int h = 0; h < ns && r == VSB_OK; ++h) {
            const int b0 = n_frames * h / ns, b1 = n_frames * (h + 1) / ns;
            s->f0 = b0;
            if (stagger) {
                r = launch_front(s, 0, n, b1 - b0, d_srcs + (size_t)b0 * n, src_pitch, st, nullptr, FRONT_REMAP);
                if (r != VSB_OK) break;
                CK(cudaEventRecord(s->ev_forks[h], st));
                CK(cudaStreamWaitEvent(s->sub[h], s->ev_forks[h], 0));
            } else {
                CK(cudaStreamWaitEvent(s->sub[h], s->ev_fork, 0));
            }
            r = launch_front(s, 0, n, b1 - b0, nullptr, src_pitch, s->sub[h], nullptr, FRONT_PYRAMID);
            if (r == VSB_OK) r = launch_back(s, b1 - b0, d_outs + b0, out_pitch, s->sub[h]);
            if (r == VSB_OK) CK(cudaEventRecord(s->ev_join[h], s->sub[h]));
        }
        for (int h = 0; h < ns && r == VSB_OK; ++h) CK(cudaStreamWaitEvent(st, s->ev_join[h], 0));
        s->f0 = 0;
        if (r != VSB_OK) return r;
        return note_compose_done(s, st);
    }
    prof_begin(s, st);
    r = launch_front(s, 0, s->cfg.num_views, n_frames, d_srcs, src_pitch, st);
    if (r != VSB_OK) return r;
    r = launch_back(s, n_frames, d_outs, out_pitch, st);
    if (r != VSB_OK) return r;
    return note_compose_done(s, st);
}

// Host-buffer entry points (what A/timed.cpp:68 `upload` and the consumer thread's `download`, :252, do around the path).
// vsb_submit_host enqueues ONE submission and returns: upload (copy engine, in_stream) -> compose (io_stream) -> download (second
// copy engine, out_stream), pipelined over sub-batches of HOST_SUB frames -- sub-batch g + 1 uploads while g composes and g - 1
// downloads -- and over submissions: up to HOST_DEPTH submissions are in flight, each with its own device staging, so the tail of
// one (last compose + download) overlaps the head of the next (first uploads).  vsb_wait_host blocks until the OLDEST outstanding
// submission's panoramas are in host memory (the reference hands results to its consumer through a queue the same way,
// A/timed.cpp:150,243).  Composing sub-batches instead of single frames reads the remap tap tables once per sub-batch.
static int host_stage_setup(vsb_stitcher *s, bool nv12, int sw, int sh, int src_rows, size_t out_pitch)
{
    const int n = s->cfg.num_views, F = s->cfg.max_batch;
    if (s->stage_src_w != sw || s->stage_src_h != sh || s->stage_views != n || s->stage_batch != F) {
        CK(cudaDeviceSynchronize());
        for (int d = 0; d < HOST_DEPTH; ++d) { cudaFree(s->stage_src[d]); cudaFree(s->stage_nv12[d]); s->stage_src[d] = s->stage_nv12[d] = nullptr; }
        s->stage_src_w = sw; s->stage_src_h = sh; s->stage_views = n; s->stage_batch = F;
    }
    for (int d = 0; d < HOST_DEPTH; ++d) {
        if (!nv12 && !s->stage_src[d]) {
            s->stage_src_pitch = align_up((size_t)sw * 3, 4);  // tight rows: a packed host frame moves as ONE contiguous DMA
            s->stage_src_frame = align_up(s->stage_src_pitch * sh + 16, 256);
            CK(cudaMalloc(&s->stage_src[d], s->stage_src_frame * n * F));
        }
        if (nv12 && !s->stage_nv12[d]) {
            s->stage_nv12_frame = align_up(align_up((size_t)sw, 4) * src_rows + 16, 256);
            CK(cudaMalloc(&s->stage_nv12[d], s->stage_nv12_frame * n * F));
        }
    }
    // the device-side panorama staging uses the HOST pitch, so each download is one contiguous DMA (a 2-D copy of 600+
    // rows whose pitches differ by a few bytes runs at a fraction of the link rate)
    if (!s->stage_out[0] || s->stage_out_pitch != out_pitch) {
        CK(cudaDeviceSynchronize());
        for (int d = 0; d < HOST_DEPTH; ++d) { cudaFree(s->stage_out[d]); s->stage_out[d] = nullptr; }
        s->stage_out_pitch = out_pitch;
        s->stage_out_frame = align_up(out_pitch * s->roi_final[3], 256);
        for (int d = 0; d < HOST_DEPTH; ++d) CK(cudaMalloc(&s->stage_out[d], s->stage_out_frame * F));
    }
    if (!s->in_stream) {
        CK(cudaStreamCreateWithFlags(&s->in_stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s->out_stream, cudaStreamNonBlocking));
        for (int f = 0; f < MAX_BATCH; ++f) {
            CK(cudaEventCreateWithFlags(&s->ev_in[f], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&s->ev_done[f], cudaEventDisableTiming));
        }
        for (int d = 0; d < HOST_DEPTH; ++d) CK(cudaEventCreateWithFlags(&s->ev_host[d], cudaEventDisableTiming));
    }
    return VSB_OK;
}

int vsb_wait_host(vsb_stitcher *s)
{
    REQ(s, VSB_ERR_INVALID, "wait_host: null handle");
    REQ(s->host_pending > 0, VSB_ERR_STATE, "wait_host: no submission outstanding");
    DeviceGuard g(s->device);
    const int d = (int)((s->host_seq - (unsigned)s->host_pending) % HOST_DEPTH);
    CK(cudaEventSynchronize(s->ev_host[d]));
    --s->host_pending;
    return VSB_OK;
}

int vsb_submit_host(vsb_stitcher *s, int n_frames, const uint8_t *const *h_srcs, size_t src_pitch, int16_t *const *h_outs, size_t out_pitch)
{
    NvtxRange nvtx("vsb_submit_host (H2D, compose, D2H)");
    REQ(s && h_srcs && h_outs, VSB_ERR_INVALID, "submit_host: null argument");
    REQ(n_frames >= 1 && n_frames <= s->cfg.max_batch, VSB_ERR_INVALID, "submit_host: n_frames must be 1..max_batch (%d)", s->cfg.max_batch);
    int r = ready_for_frames(s);
    if (r != VSB_OK) return r;
    DeviceGuard g(s->device);
    const int n = s->cfg.num_views;
    const int sw = frame_w(s, 0), sh = frame_h(s, 0);  // (the full-size frame when compose_scale != 1)
    for (int i = 1; i < n; ++i) REQ(frame_w(s, i) == sw && frame_h(s, i) == sh, VSB_ERR_INVALID, "submit_host: all views must share one source size");
    const bool nv12 = s->in_format == VSB_IN_NV12;
    const size_t opx = s->out_format == VSB_OUT_U8C3 ? 3 : 6;     // bytes per output pixel
    const size_t src_row = nv12 ? (size_t)sw : (size_t)sw * 3;    // bytes per source row, rows per source frame
    const int src_rows = nv12 ? sh * 3 / 2 : sh;
    REQ(src_pitch >= src_row && out_pitch >= (size_t)s->roi_final[2] * opx, VSB_ERR_INVALID, "submit_host: pitch too small");
    REQ(!nv12 || ((sw | sh) & 1) == 0, VSB_ERR_INVALID, "submit_host: NV12 needs even source sizes");
    while (s->host_pending >= HOST_DEPTH) { r = vsb_wait_host(s); if (r != VSB_OK) return r; }  // this slot's staging is still in use
    r = host_stage_setup(s, nv12, sw, sh, src_rows, out_pitch);
    if (r != VSB_OK) return r;
    const int d = (int)(s->host_seq % HOST_DEPTH);
    const size_t st_pitch = nv12 ? align_up((size_t)sw, 4) : s->stage_src_pitch, st_frame = nv12 ? s->stage_nv12_frame : s->stage_src_frame;
    uint8_t *st_base = nv12 ? s->stage_nv12[d] : s->stage_src[d];
    static const int host_sub = [] { const char *e = std::getenv("VSB_HOST_SUB"); return e ? std::max(1, std::min(MAX_BATCH, std::atoi(e))) : HOST_SUB; }();
    for (int f0 = 0, gidx = 0; f0 < n_frames; f0 += host_sub, ++gidx) {
        const int nf = std::min(host_sub, n_frames - f0);
        const uint8_t *d_srcs[MAX_BATCH * MAXV];
        int16_t *d_outs[MAX_BATCH];
        for (int f = f0; f < f0 + nf; ++f) {
            for (int i = 0; i < n; ++i) {
                uint8_t *dd = st_base + st_frame * (size_t)(f * n + i);
                if (src_pitch == st_pitch)
                    CK(cudaMemcpyAsync(dd, h_srcs[f * n + i], src_pitch * src_rows, cudaMemcpyHostToDevice, s->in_stream));
                else
                    CK(cudaMemcpy2DAsync(dd, st_pitch, h_srcs[f * n + i], src_pitch, src_row, src_rows, cudaMemcpyHostToDevice, s->in_stream));
                d_srcs[(f - f0) * n + i] = dd;
            }
            d_outs[f - f0] = (int16_t *)((char *)s->stage_out[d] + s->stage_out_frame * f);
        }
        CK(cudaEventRecord(s->ev_in[gidx], s->in_stream));
        CK(cudaStreamWaitEvent(s->io_stream, s->ev_in[gidx], 0));
        r = vsb_compose(s, nf, d_srcs, st_pitch, d_outs, s->stage_out_pitch, s->io_stream);
        if (r != VSB_OK) { cudaDeviceSynchronize(); return r; }
        CK(cudaEventRecord(s->ev_done[gidx], s->io_stream));
        CK(cudaStreamWaitEvent(s->out_stream, s->ev_done[gidx], 0));
        for (int f = f0; f < f0 + nf; ++f) {
            if (out_pitch == (size_t)s->roi_final[2] * opx)  // packed host rows: one contiguous DMA
                CK(cudaMemcpyAsync(h_outs[f], d_outs[f - f0], out_pitch * (size_t)s->roi_final[3], cudaMemcpyDeviceToHost, s->out_stream));
            else                                              // padded host rows: leave the caller's padding untouched
                CK(cudaMemcpy2DAsync(h_outs[f], out_pitch, d_outs[f - f0], s->stage_out_pitch, (size_t)s->roi_final[2] * opx, s->roi_final[3], cudaMemcpyDeviceToHost, s->out_stream));
        }
    }
    CK(cudaEventRecord(s->ev_host[d], s->out_stream));
    // the next submission's uploads reuse the OTHER staging set; this one is reused two submissions from now, after its wait
    ++s->host_seq;
    ++s->host_pending;
    return VSB_OK;
}

int vsb_compose_host(vsb_stitcher *s, int n_frames, const uint8_t *const *h_srcs, size_t src_pitch, int16_t *const *h_outs, size_t out_pitch)
{
    REQ(s, VSB_ERR_INVALID, "compose_host: null handle");
    while (s->host_pending > 0) { int r = vsb_wait_host(s); if (r != VSB_OK) return r; }
    int r = vsb_submit_host(s, n_frames, h_srcs, src_pitch, h_outs, out_pitch);
    if (r != VSB_OK) return r;
    return vsb_wait_host(s);
}


// ---- view-sharded multi-GPU mode: one process per GPU, every rank calibrates the same rig ---------
// Rank r owns (1) the canvas strip [r * strip_w, (r + 1) * strip_w) -- its k_coarse / k_blend tiles -- and (2) the views
// whose level-0 weight lies mostly inside that strip -- their remap / pyramid front half.  Between the two halves the
// ranks exchange the Gaussian sub-planes (u8) that foreign strips read; vsb_shard_rect names them, the transport is the
// caller's (NCCL send/recv in video-stitcher_b200/dist.py).
// k_blend tiles read C2 two level-2 samples (8 level-0 columns) beyond their own columns, so a rank also runs the k_coarse
// tiles that touch its strip widened by that margin (the neighbour computes the same tile for itself).
static bool coarse_tile_of_rank(const vsb_stitcher *s, int tx, int rank)
{
    const int unit = s->ct * 4, x0 = rank * s->strip_w - 8, x1 = std::min((rank + 1) * s->strip_w, s->cw[0]) + 8;
    return tx * unit < x1 && (tx + 1) * unit > x0;
}

int vsb_shard_set(vsb_stitcher *s, int rank, int world)
{
    REQ(s, VSB_ERR_INVALID, "shard_set: null handle");
    REQ(s->finalized && s->fast, VSB_ERR_STATE, "shard_set: needs a calibrated handle with num_bands >= 3");
    REQ(world >= 1 && rank >= 0 && rank < world, VSB_ERR_INVALID, "shard_set: bad rank / world");
    DeviceGuard g(s->device);
    CK(cudaDeviceSynchronize());
    const int n = s->cfg.num_views, cw0 = s->cw[0];
    const int unit = 256;  // strips are multiples of 256 level-0 columns (whole k_coarse tiles at either tile size)
    s->strip_w = (int)(align_up((size_t)(cw0 + world - 1) / world, unit));
    s->shard_rank = rank; s->shard_world = world;
    for (int i = 0; i < n; ++i) {
        const View &V = s->v[i];
        std::vector<long long> per(world, 0);
        for (int x = 0; x < V.bw; ++x) per[std::min((V.x_tl + x) / s->strip_w, world - 1)] += V.w0_cols[x];
        int best = 0;
        for (int r = 1; r < world; ++r) if (per[r] > per[best]) best = r;
        s->owned[i] = best == rank;
    }
    std::vector<uint32_t> b = s->h_bviews, c = s->h_cviews;
    for (int ty = 0; ty < s->blend_tiles_y; ++ty)
        for (int tx = 0; tx < s->blend_tiles_x; ++tx)
            if (std::min(tx * BL_TW / s->strip_w, world - 1) != rank) b[(size_t)ty * s->blend_tiles_x + tx] = 0x80000000u;
    for (int ty = 0; ty < s->coarse_tiles_y; ++ty)
        for (int tx = 0; tx < s->coarse_tiles_x; ++tx)
            if (!coarse_tile_of_rank(s, tx, rank)) c[(size_t)ty * s->coarse_tiles_x + tx] = 0x80000000u;
    CK(cudaMemcpy(s->d_blend_views, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->d_coarse_views, c.data(), c.size() * 4, cudaMemcpyHostToDevice));
    {   // front-half tile lists of the owned views only (one launch per kernel for any set of views)
        std::vector<uint32_t> d2;
        for (uint32_t tl : s->h_d2tiles) if (s->owned[tl & 0xff]) d2.push_back(tl);
        if (!d2.empty()) CK(cudaMemcpy(s->d_down2_tiles, d2.data(), d2.size() * 4, cudaMemcpyHostToDevice));
        s->tiles_dirty = true;
    }
    return upload_blend_lists(s, b);
}

int vsb_shard_info(const vsb_stitcher *s, int *strip_x0, int *strip_x1, unsigned *owned_mask)
{
    REQ(s && s->shard_rank >= 0, VSB_ERR_STATE, "shard_info: vsb_shard_set first");
    if (strip_x0) *strip_x0 = s->shard_rank * s->strip_w;
    if (strip_x1) *strip_x1 = std::min((s->shard_rank + 1) * s->strip_w, s->cw[0]);
    if (owned_mask) { *owned_mask = 0; for (int i = 0; i < s->cfg.num_views; ++i) if (s->owned[i]) *owned_mask |= 1u << i; }
    return VSB_OK;
}

// Bounding box {x0, y0, w, h} (plane coordinates of `level`) of what rank `dst_rank`'s tiles read from Gaussian level
// `level` of `view`; w = h = 0 when nothing.  Pure function of the static tables: every rank computes every rank's needs.
int vsb_shard_rect(const vsb_stitcher *s, int dst_rank, int view, int level, int rect[4])
{
    REQ(s && rect, VSB_ERR_INVALID, "shard_rect: null argument");
    REQ(s->shard_rank >= 0, VSB_ERR_STATE, "shard_rect: vsb_shard_set first");
    REQ(dst_rank >= 0 && dst_rank < s->shard_world && view >= 0 && view < s->cfg.num_views && level >= 0 && level <= s->nb, VSB_ERR_INVALID, "shard_rect: bad argument");
    const View &V = s->v[view];
    const int pw = V.bw >> level, ph = V.bh >> level, world = s->shard_world;
    int x0 = INT32_MAX, y0 = INT32_MAX, x1 = INT32_MIN, y1 = INT32_MIN;
    auto add = [&](int ax0, int ay0, int w, int h) {  // plane coordinates, clipped
        const int bx0 = std::max(ax0, 0), by0 = std::max(ay0, 0), bx1 = std::min(ax0 + w, pw), by1 = std::min(ay0 + h, ph);
        if (bx0 >= bx1 || by0 >= by1) return;
        x0 = std::min(x0, bx0); y0 = std::min(y0, by0); x1 = std::max(x1, bx1); y1 = std::max(y1, by1);
    };
    if (level <= 2)
        for (int ty = 0; ty < s->blend_tiles_y; ++ty)
            for (int tx = 0; tx < s->blend_tiles_x; ++tx) {
                if (std::min(tx * BL_TW / s->strip_w, world - 1) != dst_rank || !(s->h_bviews[(size_t)ty * s->blend_tiles_x + tx] >> view & 1)) continue;
                if (level == 0) add(tx * BL_TW - V.x_tl, ty * BL_TH - V.y_tl, BL_TW, BL_TH);
                else if (level == 1) add((tx * BL_TW >> 1) - 1 - (V.x_tl >> 1), (ty * BL_TH >> 1) - 1 - (V.y_tl >> 1), BL_R1W, BL_R1H);
                else add((tx * BL_TW >> 2) - 2 - (V.x_tl >> 2), (ty * BL_TH >> 2) - 2 - (V.y_tl >> 2), BL_R2W, BL_R2H);
            }
    if (level >= 2) {
        const int j = level - 2;
        for (int ty = 0; ty < s->coarse_tiles_y; ++ty)
            for (int tx = 0; tx < s->coarse_tiles_x; ++tx) {
                if (!coarse_tile_of_rank(s, tx, dst_rank) || !(s->h_cviews[(size_t)ty * s->coarse_tiles_x + tx] >> view & 1)) continue;
                add(((tx * s->ct) >> j) + s->cgeo.a_lo[j] - (V.x_tl >> level), ((ty * s->ct) >> j) + s->cgeo.a_lo[j] - (V.y_tl >> level), s->cgeo.a_n[j], s->cgeo.a_n[j]);
            }
    }
    if (x0 > x1) { rect[0] = rect[1] = rect[2] = rect[3] = 0; return VSB_OK; }
    rect[0] = x0; rect[1] = y0; rect[2] = x1 - x0; rect[3] = y1 - y0;
    return VSB_OK;
}

// device address of plane 0 of Gaussian level `level` of `view` in frame slot `frame` ([3][h][w] u8, rows of w bytes)
int vsb_get_plane(vsb_stitcher *s, int view, int level, int frame, void **ptr, int *w, int *h)
{
    REQ(s && ptr, VSB_ERR_INVALID, "get_plane: null argument");
    REQ(s->finalized && s->fast, VSB_ERR_STATE, "get_plane: needs a calibrated handle with num_bands >= 3");
    REQ(view >= 0 && view < s->cfg.num_views && level >= 0 && level <= s->nb && frame >= 0 && frame < s->cfg.max_batch, VSB_ERR_INVALID, "get_plane: bad argument");
    const View &V = s->v[view];
    if (level == 0) *ptr = V.G0 + V.g0_frame_stride * frame;
    else if (level == 1) *ptr = V.G1 + V.g1_frame_stride * frame;
    else *ptr = V.Gu[level] + V.gu_frame_stride[level] * frame;
    if (w) *w = V.bw >> level;
    if (h) *h = V.bh >> level;
    return VSB_OK;
}

// ---- view-sharded mode, batched: F frames per exchange and ONE message per peer -------------------------------------------------
// vsb_shard_plan fixes, from the owner of every view, what this rank sends to / receives from each peer (the rectangles of
// vsb_shard_rect, ordered by (view, level) on both sides); vsb_shard_pack / vsb_shard_unpack move them between the planes and a
// contiguous buffer with one kernel each.  Per submission: vsb_feed_batch (owned views) -> pack -> send/recv -> unpack -> vsb_blend_batch.
int vsb_shard_plan(vsb_stitcher *s, const int *owners)
{
    REQ(s && owners, VSB_ERR_INVALID, "shard_plan: null argument");
    REQ(s->shard_rank >= 0, VSB_ERR_STATE, "shard_plan: vsb_shard_set first");
    DeviceGuard g(s->device);
    CK(cudaDeviceSynchronize());
    const int n = s->cfg.num_views, me = s->shard_rank;
    for (int p = 0; p < s->shard_world && p < MAXV; ++p) {
        cudaFree(s->d_send[p]); cudaFree(s->d_recv[p]);
        s->d_send[p] = s->d_recv[p] = nullptr; s->n_send[p] = s->n_recv[p] = 0; s->send_bytes[p] = s->recv_bytes[p] = 0;
        if (p == me) continue;
        for (int dir = 0; dir < 2; ++dir) {  // 0: what p reads of MY views; 1: what I read of p's views
            std::vector<ShardRect> tab;
            size_t off = 0;
            for (int v = 0; v < n; ++v) {
                REQ(owners[v] >= 0 && owners[v] < s->shard_world, VSB_ERR_INVALID, "shard_plan: bad owner of view %d", v);
                if (owners[v] != (dir == 0 ? me : p)) continue;
                for (int k = 0; k <= s->nb; ++k) {
                    int rc[4];
                    int r = vsb_shard_rect(s, dir == 0 ? p : me, v, k, rc);
                    if (r != VSB_OK) return r;
                    if (rc[2] <= 0 || rc[3] <= 0) continue;
                    const View &V = s->v[v];
                    ShardRect R;
                    R.plane = k == 0 ? V.G0 : (k == 1 ? V.G1 : V.Gu[k]);
                    R.frame_stride = k == 0 ? V.g0_frame_stride : (k == 1 ? V.g1_frame_stride : V.gu_frame_stride[k]);
                    R.pw = V.bw >> k; R.ph = V.bh >> k; R.x0 = rc[0]; R.y0 = rc[1]; R.w = rc[2]; R.h = rc[3]; R.off = off;
                    if ((R.pw & 3) == 0) {  // widen to word boundaries (still inside the plane): both sides do the same, the extra columns are valid data
                        const int x1 = std::min(R.pw, (R.x0 + R.w + 3) & ~3);
                        R.x0 &= ~3; R.w = x1 - R.x0;
                    }
                    // every block starts on a 16-byte boundary of the packed buffer (k_shard_copy moves words when the plane side
                    // allows it; blocks of odd-sized coarse levels sit between them), and so does every frame of a submission
                    off = align_up(off + (size_t)3 * R.w * R.h, 16);
                    tab.push_back(R);
                }
            }
            ShardRect **d = dir == 0 ? &s->d_send[p] : &s->d_recv[p];
            (dir == 0 ? s->n_send[p] : s->n_recv[p]) = (int)tab.size();
            (dir == 0 ? s->send_bytes[p] : s->recv_bytes[p]) = off;
            if (!tab.empty()) {
                CK(cudaMalloc(d, tab.size() * sizeof(ShardRect)));
                CK(cudaMemcpy(*d, tab.data(), tab.size() * sizeof(ShardRect), cudaMemcpyHostToDevice));
            }
        }
    }
    return VSB_OK;
}

int vsb_shard_peer_bytes(const vsb_stitcher *s, int peer, size_t *send_bytes_per_frame, size_t *recv_bytes_per_frame)
{
    REQ(s && s->shard_rank >= 0 && peer >= 0 && peer < s->shard_world && peer < MAXV, VSB_ERR_INVALID, "shard_peer_bytes: bad argument");
    if (send_bytes_per_frame) *send_bytes_per_frame = s->send_bytes[peer];
    if (recv_bytes_per_frame) *recv_bytes_per_frame = s->recv_bytes[peer];
    return VSB_OK;
}

static int shard_copy(vsb_stitcher *s, int peer, int n_frames, void *d_buf, void *stream, bool pack)
{
    REQ(s && d_buf, VSB_ERR_INVALID, "shard_pack/unpack: null argument");
    REQ(s->shard_rank >= 0 && peer >= 0 && peer < s->shard_world && peer < MAXV && peer != s->shard_rank, VSB_ERR_INVALID, "shard_pack/unpack: bad peer");
    REQ(n_frames >= 1 && n_frames <= s->cfg.max_batch, VSB_ERR_INVALID, "shard_pack/unpack: n_frames must be 1..max_batch");
    DeviceGuard g(s->device);
    const int n = pack ? s->n_send[peer] : s->n_recv[peer];
    if (n == 0) return VSB_OK;
    if (pack) k_shard_copy<true><<<dim3(n, 3, n_frames), dim3(32, 8), 0, (cudaStream_t)stream>>>(s->d_send[peer], (uint8_t *)d_buf, s->send_bytes[peer], 0);
    else k_shard_copy<false><<<dim3(n, 3, n_frames), dim3(32, 8), 0, (cudaStream_t)stream>>>(s->d_recv[peer], (uint8_t *)d_buf, s->recv_bytes[peer], 0);
    return check_launch("k_shard_copy");
}
int vsb_shard_pack(vsb_stitcher *s, int peer, int n_frames, void *d_buf, void *stream) { return shard_copy(s, peer, n_frames, d_buf, stream, true); }
int vsb_shard_unpack(vsb_stitcher *s, int peer, int n_frames, void *d_buf, void *stream) { return shard_copy(s, peer, n_frames, d_buf, stream, false); }

// stitch_online for views [v0, v1) of n_frames frames in one submission; d_srcs[f * (v1 - v0) + (i - v0)]
int vsb_feed_batch(vsb_stitcher *s, int v0, int v1, int n_frames, const uint8_t *const *d_srcs, size_t pitch, void *stream)
{
    REQ(s && d_srcs, VSB_ERR_INVALID, "feed_batch: null argument");
    REQ(v0 >= 0 && v1 <= s->cfg.num_views && v0 < v1, VSB_ERR_INVALID, "feed_batch: bad view range");
    REQ(n_frames >= 1 && n_frames <= s->cfg.max_batch, VSB_ERR_INVALID, "feed_batch: n_frames must be 1..max_batch (%d)", s->cfg.max_batch);
    int r = ready_for_frames(s);
    if (r != VSB_OK) return r;
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (s->launches == 0) prof_begin(s, st);
    r = adopt_meshes(s, st);
    if (r != VSB_OK) return r;
    return note_front_done(s, st, launch_front(s, v0, v1, n_frames, d_srcs, pitch, st));
}

int vsb_blend_batch(vsb_stitcher *s, int n_frames, int16_t *const *d_outs, size_t out_pitch, void *stream)
{
    REQ(s && d_outs, VSB_ERR_INVALID, "blend_batch: null argument");
    REQ(s->finalized, VSB_ERR_STATE, "blend_batch: prepare + init_view for every view must come first");
    REQ(n_frames >= 1 && n_frames <= s->cfg.max_batch, VSB_ERR_INVALID, "blend_batch: n_frames must be 1..max_batch (%d)", s->cfg.max_batch);
    REQ(s->out_format == VSB_OUT_S16C3 || s->fast, VSB_ERR_STATE, "blend_batch: CV_8UC3 output needs num_bands >= 3");
    REQ(out_pitch >= (size_t)s->roi_final[2] * (s->out_format == VSB_OUT_U8C3 ? 3 : 6), VSB_ERR_INVALID, "blend_batch: output pitch too small");
    DeviceGuard g(s->device);
    int r = launch_back(s, n_frames, d_outs, out_pitch, (cudaStream_t)stream);
    if (r != VSB_OK) return r;
    return note_compose_done(s, (cudaStream_t)stream);
}

// ---- view-sharded mode, native transport: the exchange lives in the library, so a C++ host needs nothing else ------
// NCCL is reached through dlopen (nccl_api above): libvsb200 has no link-time dependency on it, and inside a process that
// already loaded an NCCL (torch) the same library instance is used.

int vsb_shard_unique_id(void *id128)
{
    REQ(id128, VSB_ERR_INVALID, "shard_unique_id: null argument");
    const NcclApi *nccl = nccl_api();
    REQ(nccl, VSB_ERR_STATE, "shard_unique_id: libnccl.so.2 not found (%s)", dlerror());
    ncclUniqueId id;
    NC(nccl->GetUniqueId(&id));
    static_assert(sizeof(id) == VSB_SHARD_ID_BYTES, "ncclUniqueId size");
    std::memcpy(id128, &id, sizeof(id));
    return VSB_OK;
}

static void shard_free_transport(vsb_stitcher *s)
{
    for (int b = 0; b < 2; ++b)
        for (int p = 0; p < MAXV; ++p) { cudaFree(s->x_send[b][p]); cudaFree(s->x_recv[b][p]); s->x_send[b][p] = s->x_recv[b][p] = nullptr; }
    s->x_frames = 0;
}

int vsb_shard_init(vsb_stitcher *s, int rank, int world, const void *id128)
{
    REQ(s && id128, VSB_ERR_INVALID, "shard_init: null argument");
    REQ(world >= 1 && world <= MAXV && rank >= 0 && rank < world, VSB_ERR_INVALID, "shard_init: bad rank / world (at most %d ranks)", MAXV);
    const NcclApi *nccl = nccl_api();
    REQ(nccl, VSB_ERR_STATE, "shard_init: libnccl.so.2 not found");
    int r = vsb_shard_set(s, rank, world);
    if (r != VSB_OK) return r;
    DeviceGuard g(s->device);
    for (int i = 0; i < s->cfg.num_views; ++i) s->owners[i] = view_owner(s, i, world);
    r = vsb_shard_plan(s, s->owners);
    if (r != VSB_OK) return r;
    shard_free_transport(s);
    if (!s->sh_front) {
        CK(cudaStreamCreateWithFlags(&s->sh_front, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s->sh_comm, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s->sh_back, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&s->ev_call, cudaEventDisableTiming));
        for (int b = 0; b < 2; ++b) {
            CK(cudaEventCreateWithFlags(&s->ev_packed[b], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&s->ev_recv[b], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&s->ev_back[b], cudaEventDisableTiming));
        }
    }
    s->ev_back_valid[0] = s->ev_back_valid[1] = false;
    s->shard_seq = 0;
    if (s->comm) { nccl->CommDestroy(s->comm); s->comm = nullptr; }
    if (world > 1) {
        ncclUniqueId id;
        std::memcpy(&id, id128, sizeof(id));
        // This communicator may live next to the host application's own.  NCCL's peer-to-peer transport passes file descriptors
        // between ranks while it connects, so the soft descriptor limit is lifted to the hard one; and the exchange moves a few
        // MB per frame, which eight CTAs saturate, so the communicator is asked for few channels (less memory per peer).
        struct rlimit rl;
        if (getrlimit(RLIMIT_NOFILE, &rl) == 0 && rl.rlim_cur < rl.rlim_max) { rl.rlim_cur = rl.rlim_max; setrlimit(RLIMIT_NOFILE, &rl); }
        static const int max_ctas = [] { const char *e = std::getenv("VSB_NCCL_MAXCTAS"); return e ? std::atoi(e) : 8; }();
        if (nccl->CommInitRankConfig && max_ctas > 0) {
            ncclConfig_t nc = NCCL_CONFIG_INITIALIZER;
            nc.maxCTAs = max_ctas;
            NC(nccl->CommInitRankConfig(&s->comm, world, id, rank, &nc));
        } else {
            NC(nccl->CommInitRank(&s->comm, world, id, rank));
        }
        // Connect the peers of the plan now, one peer pair per group (step d: send to rank + d, receive from rank - d), instead
        // of all at once inside the first exchange: a transport that cannot be set up fails here, with the peer named, and the
        // first vsb_shard_compose does not pay the set-up.
        static const int eager = [] { const char *e = std::getenv("VSB_NCCL_EAGER"); return e ? std::atoi(e) : 1; }();
        if (eager) {
            uint8_t *d_tok = nullptr;
            CK(cudaMalloc(&d_tok, 2 * 16));
            CK(cudaMemsetAsync(d_tok, 0, 2 * 16, s->sh_comm));
            for (int d = 1; d < world; ++d) {
                const int to = (rank + d) % world, from = (rank - d + world) % world;
                const bool tx = s->send_bytes[to] != 0, rx = s->recv_bytes[from] != 0;
                if (!tx && !rx) continue;
                ncclResult_t e = nccl->GroupStart();
                if (e == ncclSuccess && tx) e = nccl->Send(d_tok, 16, ncclUint8, to, s->comm, s->sh_comm);
                if (e == ncclSuccess && rx) e = nccl->Recv(d_tok + 16, 16, ncclUint8, from, s->comm, s->sh_comm);
                const ncclResult_t e2 = nccl->GroupEnd();
                if (e == ncclSuccess) e = e2;
                if (e != ncclSuccess) {
                    cudaFree(d_tok);
                    return fail(VSB_ERR_CUDA, "shard_init: connecting rank %d -> %d / %d -> %d: %s", rank, to, from, rank, nccl->GetErrorString(e));
                }
            }
            cudaError_t ce = cudaStreamSynchronize(s->sh_comm);
            cudaFree(d_tok);
            CK(ce);
        }
    }
    return VSB_OK;
}

// stitch_one (A/timed.cpp:123-152) for n_frames frames of ONE frame stream composed by `world` GPUs: this rank remaps / builds
// the pyramids of its views, the ranks exchange the Gaussian sub-planes foreign strips read (one grouped ncclSend / ncclRecv
// per peer, stream-ordered, no host wait), and this rank blends its canvas strip into d_outs (full-size buffers).
// d_srcs[f * num_views + v]: only the entries of owned views are read.  Three internal streams (front / exchange / back) and
// two alternating halves of the frame slots (when 2 * n_frames <= max_batch): submission k + 1's front half overlaps
// submission k's exchange and back half.  Outputs are ordered on `stream`.
int vsb_shard_compose(vsb_stitcher *s, int n_frames, const uint8_t *const *d_srcs, size_t src_pitch, int16_t *const *d_outs, size_t out_pitch, void *stream)
{
    NvtxRange nvtx("vsb_shard_compose (front, exchange, back)");
    REQ(s && d_srcs && d_outs, VSB_ERR_INVALID, "shard_compose: null argument");
    REQ(s->shard_rank >= 0 && s->sh_front, VSB_ERR_STATE, "shard_compose: vsb_shard_init first");
    REQ(n_frames >= 1 && n_frames <= s->cfg.max_batch, VSB_ERR_INVALID, "shard_compose: n_frames must be 1..max_batch (%d)", s->cfg.max_batch);
    REQ(out_pitch >= (size_t)s->roi_final[2] * (s->out_format == VSB_OUT_U8C3 ? 3 : 6), VSB_ERR_INVALID, "shard_compose: output pitch too small");
    const NcclApi *nccl = nccl_api();
    REQ(nccl || s->shard_world == 1, VSB_ERR_STATE, "shard_compose: libnccl.so.2 not found");
    int r = ready_for_frames(s);
    if (r != VSB_OK) return r;
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int n = s->cfg.num_views, me = s->shard_rank, world = s->shard_world, F = n_frames;
    if (s->x_frames < F) {  // exchange buffers: per peer and direction, two sets
        CK(cudaDeviceSynchronize());
        shard_free_transport(s);
        for (int b = 0; b < 2; ++b)
            for (int p = 0; p < world; ++p) {
                if (p == me) continue;
                if (s->send_bytes[p]) CK(cudaMalloc(&s->x_send[b][p], s->send_bytes[p] * F));
                if (s->recv_bytes[p]) CK(cudaMalloc(&s->x_recv[b][p], s->recv_bytes[p] * F));
            }
        s->x_frames = F;
    }
    const bool two = 2 * F <= s->cfg.max_batch;
    const int b = two ? (int)(s->shard_seq & 1u) : 0;
    const int f0 = b * F;
    ++s->shard_seq;
    s->launches = 0;
    // ---- front half on sh_front: inputs ready when `stream` reaches this call; the slots are free once their previous back half is done
    CK(cudaEventRecord(s->ev_call, st));
    CK(cudaStreamWaitEvent(s->sh_front, s->ev_call, 0));
    if (s->ev_back_valid[b]) CK(cudaStreamWaitEvent(s->sh_front, s->ev_back[b], 0));
    r = adopt_meshes(s, s->sh_front);
    if (r != VSB_OK) return r;
    s->f0 = f0;
    for (int f = 0; f < F; ++f)
        for (int v = 0; v < n; ++v)
            if (s->owned[v] && !d_srcs[f * n + v]) { s->f0 = 0; return fail(VSB_ERR_INVALID, "shard_compose: frame %d of owned view %d is null", f, v); }
    // one launch per front-half kernel: the device tile lists hold this rank's views only (vsb_shard_set)
    r = launch_front(s, 0, n, F, d_srcs, src_pitch, s->sh_front);
    for (int p = 0; p < world && r == VSB_OK; ++p)
        if (p != me && s->n_send[p] > 0) {
            k_shard_copy<true><<<dim3(s->n_send[p], 3, F), dim3(32, 8), 0, s->sh_front>>>(s->d_send[p], s->x_send[b][p], s->send_bytes[p], f0);
            ++s->launches;
        }
    if (r == VSB_OK) r = check_launch("shard front half");
    if (r != VSB_OK) { s->f0 = 0; return r; }
    CK(cudaEventRecord(s->ev_packed[b], s->sh_front));
    // ---- exchange on sh_comm: one grouped send / recv per peer, then scatter into the local copies of the foreign planes
    CK(cudaStreamWaitEvent(s->sh_comm, s->ev_packed[b], 0));
    if (s->ev_back_valid[b]) CK(cudaStreamWaitEvent(s->sh_comm, s->ev_back[b], 0));  // the planes the scatter overwrites were read by that back half
    if (world > 1) {
        NC(nccl->GroupStart());
        for (int p = 0; p < world; ++p) {
            if (p == me) continue;
            if (s->send_bytes[p]) NC(nccl->Send(s->x_send[b][p], s->send_bytes[p] * F, ncclUint8, p, s->comm, s->sh_comm));
            if (s->recv_bytes[p]) NC(nccl->Recv(s->x_recv[b][p], s->recv_bytes[p] * F, ncclUint8, p, s->comm, s->sh_comm));
        }
        NC(nccl->GroupEnd());
        for (int p = 0; p < world; ++p)
            if (p != me && s->n_recv[p] > 0) {
                k_shard_copy<false><<<dim3(s->n_recv[p], 3, F), dim3(32, 8), 0, s->sh_comm>>>(s->d_recv[p], s->x_recv[b][p], s->recv_bytes[p], f0);
                ++s->launches;
            }
        r = check_launch("shard exchange");
        if (r != VSB_OK) { s->f0 = 0; return r; }
    }
    CK(cudaEventRecord(s->ev_recv[b], s->sh_comm));
    // ---- back half on sh_back: this rank's canvas strip
    CK(cudaStreamWaitEvent(s->sh_back, s->ev_recv[b], 0));
    r = launch_back(s, F, d_outs, out_pitch, s->sh_back);
    s->f0 = 0;
    if (r != VSB_OK) return r;
    CK(cudaEventRecord(s->ev_back[b], s->sh_back));
    s->ev_back_valid[b] = true;
    CK(cudaStreamWaitEvent(st, s->ev_back[b], 0));
    return note_compose_done(s, st);
}

int vsb_shard_exchange_bytes(const vsb_stitcher *s, size_t *send_bytes_per_frame, size_t *recv_bytes_per_frame)
{
    REQ(s && s->shard_rank >= 0, VSB_ERR_STATE, "shard_exchange_bytes: vsb_shard_init first");
    size_t a = 0, b = 0;
    for (int p = 0; p < s->shard_world && p < MAXV; ++p) { a += s->send_bytes[p]; b += s->recv_bytes[p]; }
    if (send_bytes_per_frame) *send_bytes_per_frame = a;
    if (recv_bytes_per_frame) *recv_bytes_per_frame = b;
    return VSB_OK;
}

int vsb_last_launch_count(const vsb_stitcher *s) { return s ? s->launches_last : 0; }

int vsb_set_profiling(vsb_stitcher *s, int on)
{
    REQ(s, VSB_ERR_INVALID, "set_profiling: null handle");
    DeviceGuard g(s->device);
    if (on && !s->prof_ev[0])
        for (int i = 0; i <= VSB_MAX_STAGES; ++i) CK(cudaEventCreate(&s->prof_ev[i]));
    s->profiling = on != 0;
    s->n_stages = 0;
    return VSB_OK;
}

int vsb_get_profile(vsb_stitcher *s, int max_stages, int *n_stages, const char **names, float *ms, double *bytes)
{
    REQ(s && n_stages && names && ms && bytes, VSB_ERR_INVALID, "get_profile: null argument");
    REQ(s->profiling, VSB_ERR_STATE, "get_profile: profiling is off");
    DeviceGuard g(s->device);
    const int n = std::min(max_stages, s->n_stages);
    if (n > 0) CK(cudaEventSynchronize(s->prof_ev[s->n_stages]));
    for (int i = 0; i < n; ++i) {
        CK(cudaEventElapsedTime(&ms[i], s->prof_ev[i], s->prof_ev[i + 1]));
        names[i] = s->stage_name[i];
        bytes[i] = s->stage_bytes[i];
    }
    *n_stages = n;
    return VSB_OK;
}

int vsb_get_config(const vsb_stitcher *s, vsb_config *out)
{
    REQ(s && out, VSB_ERR_INVALID, "get_config: null argument");
    *out = s->cfg;
    return VSB_OK;
}

// compose_scale of the reference (A/timed.cpp:74-81, A/calibration.cpp:137-205).  The maps installed with vsb_set_maps address the
// SCALED frame; the frames handed to vsb_feed / vsb_compose / vsb_submit_host are full_w x full_h and are resized on the device
// first -- when |compose_scale - 1| > 0.1, the reference's own condition; otherwise the stage is off.  Takes effect with the next submission.
int vsb_set_compose_scale(vsb_stitcher *s, double compose_scale, int full_w, int full_h)
{
    REQ(s, VSB_ERR_INVALID, "set_compose_scale: null handle");
    if (compose_scale == 1.0) { s->prescale = false; s->compose_scale = 1.0; return VSB_OK; }
    int frame[2], map_src[2], resized = 0;
    int r = vsb_compose_size(full_w, full_h, compose_scale, frame, map_src, &resized);
    if (r != VSB_OK) return r;
    s->compose_scale = compose_scale;
    s->prescale = resized != 0;   // within 0.1 of 1 the reference scales its cameras, not its frames (A/timed.cpp:75)
    s->full_w = full_w; s->full_h = full_h; s->comp_w = frame[0]; s->comp_h = frame[1];
    return VSB_OK;
}

// split calibration (vsb_calibrate_rig_split): view `view` shows columns [x0, x0 + its width) of camera `camera`'s full_w-wide
// warped image; vsb_set_mesh then takes the camera's mesh.  Internal (the calibration calls it after vsb_set_maps).
int vsb_set_view_window(vsb_stitcher *s, int view, int camera, int x0, int full_w)
{
    REQ(s && view >= 0 && view < s->cfg.num_views && s->v[view].inited, VSB_ERR_INVALID, "set_view_window: bad view");
    View &V = s->v[view];
    REQ(camera >= 0 && x0 >= 0 && full_w >= x0 + V.roi_w, VSB_ERR_INVALID, "set_view_window: window outside the camera's image");
    V.cam = camera;
    const bool whole = x0 == 0 && full_w == V.roi_w;
    V.win_x0 = whole ? 0 : x0; V.win_full_w = whole ? 0 : full_w;
    return VSB_OK;
}

int vsb_view_window(const vsb_stitcher *s, int view, int *camera, int *x0, int *full_w)
{
    REQ(s && view >= 0 && view < s->cfg.num_views, VSB_ERR_INVALID, "view_window: bad view");
    REQ(s->v[view].inited, VSB_ERR_STATE, "view_window: view %d is not initialised", view);
    const View &V = s->v[view];
    if (camera) *camera = V.cam >= 0 ? V.cam : view;
    if (x0) *x0 = V.win_x0;
    if (full_w) *full_w = V.win_full_w > 0 ? V.win_full_w : V.roi_w;
    return VSB_OK;
}

void vsb_attach_calib(vsb_stitcher *s, void *state, void (*dtor)(void *))
{
    if (s->calib_state && s->calib_dtor) s->calib_dtor(s->calib_state);
    s->calib_state = state; s->calib_dtor = dtor;
}
void *vsb_get_calib(const vsb_stitcher *s) { return s ? s->calib_state : nullptr; }
int vsb_handle_device(const vsb_stitcher *s) { return s ? s->device : -1; }

int vsb_note_rig(vsb_stitcher *s, int projection, float scale, int src_w, int src_h)
{
    s->rig_projection = projection; s->rig_scale = scale; s->rig_src_w = src_w; s->rig_src_h = src_h;
    return VSB_OK;
}

int vsb_rig_info_get(const vsb_stitcher *s, vsb_rig_info *out)
{
    REQ(s && out, VSB_ERR_INVALID, "rig_info: null argument");
    REQ(s->finalized, VSB_ERR_STATE, "rig_info: not calibrated");
    std::memset(out, 0, sizeof(*out));
    out->projection = s->rig_projection; out->scale = s->rig_scale; out->src_w = s->rig_src_w; out->src_h = s->rig_src_h;
    out->num_views = s->cfg.num_views; out->num_bands = s->nb;
    std::memcpy(out->roi_final, s->roi_final, sizeof(int) * 4);
    std::memcpy(out->roi_padded, s->roi, sizeof(int) * 4);
    for (int i = 0; i < s->cfg.num_views; ++i) {
        out->view_roi[i][0] = s->v[i].tl_x; out->view_roi[i][1] = s->v[i].tl_y;
        out->view_roi[i][2] = s->v[i].roi_w; out->view_roi[i][3] = s->v[i].roi_h;
    }
    return VSB_OK;
}

int vsb_debug_read(vsb_stitcher *s, int what, int view, int level, int frame, void *h_dst, size_t bytes)
{
    REQ(s && h_dst, VSB_ERR_INVALID, "debug_read: null argument");
    REQ(s->finalized, VSB_ERR_STATE, "debug_read: not calibrated");
    REQ(frame >= 0 && frame < s->cfg.max_batch && level >= 0 && level <= s->nb, VSB_ERR_INVALID, "debug_read: bad frame/level");
    REQ(what == 3 || (view >= 0 && view < s->cfg.num_views), VSB_ERR_INVALID, "debug_read: bad view");
    DeviceGuard g(s->device);
    CK(cudaDeviceSynchronize());
    const View &V = s->v[what == 3 ? 0 : view];
    switch (what) {
    case 0: {  // warped view (after remap #2): crop of G0 planes -> interleaved u8
        REQ(bytes == (size_t)V.roi_w * V.roi_h * 3, VSB_ERR_INVALID, "debug_read: size mismatch");
        std::vector<uint8_t> tmp(V.g0_frame_stride);
        CK(cudaMemcpy(tmp.data(), V.G0 + V.g0_frame_stride * frame, tmp.size(), cudaMemcpyDeviceToHost));
        uint8_t *o = (uint8_t *)h_dst;
        const size_t plane = (size_t)V.bw * V.bh;
        for (int y = 0; y < V.roi_h; ++y)
            for (int x = 0; x < V.roi_w; ++x)
                for (int c = 0; c < 3; ++c) o[((size_t)y * V.roi_w + x) * 3 + c] = tmp[c * plane + (size_t)(y + V.top) * V.bw + (x + V.left)];
        return VSB_OK;
    }
    case 1: {
        const int w = V.bw >> level, h = V.bh >> level;
        REQ(bytes == (size_t)w * h * 6, VSB_ERR_INVALID, "debug_read: size mismatch");
        int16_t *d_tmp = nullptr;
        CK(cudaMalloc(&d_tmp, bytes));
        if (s->fast && level != 0 && level != 2) { cudaFree(d_tmp); return fail(VSB_ERR_STATE, "debug_read: Gaussian level %d is never materialised (only 0 and 2 are)", level); }
        const void *in = level == 0 ? (const void *)(V.G0 + V.g0_frame_stride * frame)
                         : (s->fast ? (const void *)(V.G2 + V.g2_frame_stride * frame) : (const void *)(V.G[level] + V.g_frame_stride[level] * frame));
        const dim3 b(32, 8);
        k_planar_to_interleaved_s16<<<grid2d(w, h, b), b>>>(in, level == 0 || s->fast, w, h, d_tmp);
        cudaError_t e = cudaMemcpy(h_dst, d_tmp, bytes, cudaMemcpyDeviceToHost);
        cudaFree(d_tmp);
        return check_cuda(e, "debug_read level");
    }
    case 2: {
        const int w = V.bw >> level, h = V.bh >> level;
        REQ(bytes == (size_t)w * h * 4, VSB_ERR_INVALID, "debug_read: size mismatch");
        return check_cuda(cudaMemcpy(h_dst, V.weight[level], bytes, cudaMemcpyDeviceToHost), "debug_read weight");
    }
    case 3:
        REQ(bytes == (size_t)s->cw[level] * s->ch[level] * 4, VSB_ERR_INVALID, "debug_read: size mismatch");
        return check_cuda(cudaMemcpy(h_dst, s->dw[level], bytes, cudaMemcpyDeviceToHost), "debug_read dw");
    case 4:
    case 5: {
        int buf;
        { std::lock_guard<std::mutex> lk(s->mu); buf = V.mesh_pending >= 0 ? V.mesh_pending : V.mesh_cur; }
        REQ(buf >= 0, VSB_ERR_STATE, "debug_read: no mesh set");
        REQ(bytes == (size_t)V.roi_w * V.roi_h * 4, VSB_ERR_INVALID, "debug_read: size mismatch");
        return check_cuda(cudaMemcpy2D(h_dst, (size_t)V.roi_w * 4, V.mesh[buf][what - 4], V.map_pitch, (size_t)V.roi_w * 4, V.roi_h, cudaMemcpyDeviceToHost), "debug_read mesh");
    }
    case 8:
    case 9:
        REQ(V.has_maps, VSB_ERR_STATE, "debug_read: no projection maps set");
        REQ(bytes == (size_t)V.roi_w * V.roi_h * 4, VSB_ERR_INVALID, "debug_read: size mismatch");
        return check_cuda(cudaMemcpy2D(h_dst, (size_t)V.roi_w * 4, what == 8 ? V.xmap : V.ymap, V.map_pitch, (size_t)V.roi_w * 4, V.roi_h, cudaMemcpyDeviceToHost), "debug_read map");
    case 6: {  // which level-2 samples k_down2 computes (1) / skips because no tile reads them (0)
        REQ(s->fast, VSB_ERR_STATE, "debug_read: the generic path computes every level");
        const int w = V.bw >> 2, h = V.bh >> 2;
        REQ(bytes == (size_t)w * h, VSB_ERR_INVALID, "debug_read: size mismatch");
        uint8_t *o = (uint8_t *)h_dst;
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) o[(size_t)y * w + x] = V.g2_needed[(size_t)(y / D2_TH) * V.d2_tiles_x + x / D2_TW];
        return VSB_OK;
    }
    case 7: {  // which bordered level-0 samples k_remap_stage2 computes (1) / skips because nothing reads them (0)
        REQ(bytes == (size_t)V.bw * V.bh, VSB_ERR_INVALID, "debug_read: size mismatch");
        std::memset(h_dst, 0, bytes);
        uint8_t *o = (uint8_t *)h_dst;
        for (uint32_t tl : V.s2_tiles) {
            const int x0 = (int)((tl >> 8) & 0xfff) * RM_BX * RM_PX, y0 = (int)(tl >> 20) * RM_BY;
            for (int y = y0; y < std::min(V.bh, y0 + RM_BY); ++y)
                for (int x = x0; x < std::min(V.bw, x0 + RM_BX * RM_PX); ++x) o[(size_t)y * V.bw + x] = 1;
        }
        return VSB_OK;
    }
    default:
        return fail(VSB_ERR_INVALID, "debug_read: unknown selector %d", what);
    }
}

}  // extern "C"
