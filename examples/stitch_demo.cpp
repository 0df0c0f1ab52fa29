// stitch_demo.cpp -- C++ host code shaped like the reference's 360_stitcher/timed.cpp main loop, on top of
// include/vsb200.hpp: calibrate once (stitch_calib), then per frame: upload -> stitch_online x N -> blend -> consume.
//
//   stitch_demo <n_views> <src_w> <src_h> <pano_width> <num_bands> <n_frames> <frames.bin> <out.bin> [<out_w> <out_h>]
//
// frames.bin : n_frames * n_views raw BGR frames (what capture/decoding would deliver, timed.cpp:577-586)
// out.bin    : n_frames raw CV_16SC3 panoramas (what `results.push(out)` hands to the consumer thread, timed.cpp:150)
// With <out_w> <out_h> the demo runs the wire-to-wire variant: frames.bin holds NV12 frames as the capture boards send them
// (networking.cpp:46 converts them on the CPU; here setFormats(VSB_IN_NV12, ...) does it on the device), the panorama leaves
// blend() as CV_8UC3 (timed.cpp:250) and consume() turns it into the letter-boxed out_w x out_h I420 frame kvazaar is fed
// (timed.cpp:281-315); out.bin then holds n_frames I420 frames.
// Build: g++ -std=c++11 -I include -I /usr/local/cuda/include examples/stitch_demo.cpp -o stitch_demo
//            -L video-stitcher_b200 -lvsb200 -L /usr/local/cuda/lib64 -lcudart
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "vsb200.hpp"

#define CUDA_OK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

int main(int argc, char **argv)
{
    if (argc != 9 && argc != 11) { std::fprintf(stderr, "usage: %s n_views src_w src_h pano_width num_bands n_frames frames.bin out.bin [out_w out_h]\n", argv[0]); return 1; }
    const bool wire = argc == 11;
    const int out_w = wire ? std::atoi(argv[9]) : 0, out_h = wire ? std::atoi(argv[10]) : 0;
    const int n = std::atoi(argv[1]), sw = std::atoi(argv[2]), sh = std::atoi(argv[3]), pano = std::atoi(argv[4]);
    const int bands = std::atoi(argv[5]), n_frames = std::atoi(argv[6]);
    try {
        // ---- stitch_calib (timed.cpp:547): fixed-rig calibration, gains, seam masks, maps, init_gpu per view
        vsb::MultiBandBlender mb(n, bands, /*enable_local=*/true);
        std::vector<float> gains(n);
        for (int i = 0; i < n; ++i) gains[i] = 1.f + 0.03f * ((i % 3) - 1);
        mb.calibrateRig(VSB_PROJ_SPHERICAL, pano, vsb::Size(sw, sh), 90.0, gains.data());
        // ---- MeshWarper::calibrateMeshWarp -> convertMeshesToMap (meshwarper.cpp:337-369): here a synthetic 10 x 10 mesh
        std::vector<std::vector<float> > mx(n), my(n);
        std::vector<vsb::MeshCpu> meshes(n);
        for (int v = 0; v < n; ++v) {
            const vsb::Size sz = mb.viewSize(v);
            mx[v].resize(100); my[v].resize(100);
            for (int i = 0; i < 10; ++i)
                for (int j = 0; j < 10; ++j) {
                    const float ix = (float)j * (float)sz.width / 9.f, iy = (float)i * (float)sz.height / 9.f;  // identity grid, meshwarper.cpp:76-77
                    mx[v][i * 10 + j] = ix + (float)(6.0 * std::sin(M_PI * i / 9) * std::cos(M_PI * j / 9));
                    my[v][i * 10 + j] = iy + (float)(4.0 * std::sin(M_PI * j / 9));
                }
            meshes[v].x = mx[v].data(); meshes[v].y = my[v].data(); meshes[v].rows = meshes[v].cols = 10;
        }
        vsb::convertMeshesToMap(mb, meshes);

        if (wire) mb.setFormats(VSB_IN_NV12, VSB_OUT_U8C3);
        const vsb::Rect roi = mb.resultRoi();
        const size_t frame_bytes = wire ? (size_t)sw * sh * 3 / 2 : (size_t)sw * sh * 3;
        const size_t pano_bytes = (size_t)roi.width * roi.height * (wire ? 3 : 6);
        const size_t out_bytes = wire ? (size_t)out_w * out_h * 3 / 2 : pano_bytes;
        std::vector<unsigned char> h_frame(frame_bytes * n), h_out(out_bytes);
        std::vector<vsb::DeviceMat> d_frames(n);
        for (int i = 0; i < n; ++i) { void *p; CUDA_OK(cudaMalloc(&p, frame_bytes)); d_frames[i] = vsb::DeviceMat(p, wire ? (size_t)sw : (size_t)sw * 3, sh, sw); }
        void *p_out, *p_yuv = 0; CUDA_OK(cudaMalloc(&p_out, pano_bytes));
        if (wire) CUDA_OK(cudaMalloc(&p_yuv, out_bytes));
        vsb::DeviceMat gpuOut(p_out, (size_t)roi.width * (wire ? 3 : 6), roi.height, roi.width);
        cudaStream_t stream; CUDA_OK(cudaStreamCreate(&stream));
        std::FILE *fin = std::fopen(argv[7], "rb"), *fout = std::fopen(argv[8], "wb");
        if (!fin || !fout) { std::fprintf(stderr, "cannot open files\n"); return 1; }
        // ---- while (1) { getImages(); stitch_one(); results.push(out); }  (timed.cpp:574-615)
        for (int f = 0; f < n_frames; ++f) {
            if (std::fread(h_frame.data(), 1, frame_bytes * n, fin) != frame_bytes * n) { std::fprintf(stderr, "short read\n"); return 1; }
            for (int i = 0; i < n; ++i) {  // stitch_online: upload (timed.cpp:68), then remap / gain / mesh remap / feed_online
                CUDA_OK(cudaMemcpyAsync(d_frames[i].data, h_frame.data() + frame_bytes * i, frame_bytes, cudaMemcpyHostToDevice, stream));
                mb.stitch_online(d_frames[i], i, stream);
            }
            mb.blend(gpuOut, stream);  // mb->blend(result, result_mask, out, true), timed.cpp:138
            if (wire) {  // consume(): convertTo(CV_8U) is already done; resize + black bars + BGR2YUV_I420 on the device, then download
                mb.consume(gpuOut, vsb::Size(out_w, out_h), /*keep_aspect_ratio=*/true, VSB_CONSUME_I420, p_yuv, (size_t)out_w, stream);
                CUDA_OK(cudaMemcpyAsync(h_out.data(), p_yuv, out_bytes, cudaMemcpyDeviceToHost, stream));
            } else
            CUDA_OK(cudaMemcpyAsync(h_out.data(), gpuOut.data, out_bytes, cudaMemcpyDeviceToHost, stream));  // consume(): download, timed.cpp:252
            CUDA_OK(cudaStreamSynchronize(stream));
            std::fwrite(h_out.data(), 1, out_bytes, fout);
        }
        std::fclose(fin); std::fclose(fout);
        std::printf("stitch_demo: %d frames, pano %dx%d\n", n_frames, roi.width, roi.height);
    } catch (const vsb::Error &e) {
        std::fprintf(stderr, "%s (code %d)\n", e.what(), e.code);
        return 3;
    }
    return 0;
}
