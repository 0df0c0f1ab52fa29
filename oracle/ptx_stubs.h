/* oracle/ptx_stubs.h -- TEST INFRASTRUCTURE.  Forced include for oracle/ref_ptx.mk: what the reference's CUDA sources need in
 * order to PARSE under CUDA 12, where legacy texture references no longer exist.  opencv2/core/cuda/common.hpp:99 names
 * `textureReference` in bindTexture(); cudawarping/src/cuda/remap.cu:112-121 declares `texture<T, cudaTextureType2D>` globals
 * and fetches them with tex2D() for the 1- and 4-channel types.  The kernels this repository compares against (3-channel remap
 * through BorderReader<PtrStep<uchar3>, ...>, the pyramids, the blender kernels, the application's resize) never touch a
 * texture; the stubs below only let the rest of each file compile (a stubbed fetch returns a value-initialised element). */
#pragma once
struct textureReference { int normalized; int filterMode; int addressMode[3]; };
template <class T, int texType = 1, int mode = 0> struct texture : public textureReference {
    __host__ texture(int norm = 0, int fMode = 0, int aMode = 0) { normalized = norm; filterMode = fMode; addressMode[0] = addressMode[1] = addressMode[2] = aMode; }
};
template <class Tex> struct vsb_tex_elem;
template <class T, int tt, int m> struct vsb_tex_elem<texture<T, tt, m> > { typedef T type; };
#define tex2D(t, x, y) (typename vsb_tex_elem<decltype(t)>::type())
static inline cudaError_t cudaBindTexture2D(size_t *, const textureReference *, const void *, const cudaChannelFormatDesc *, size_t, size_t, size_t) { return cudaSuccess; }
