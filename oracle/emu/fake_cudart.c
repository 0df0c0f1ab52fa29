/* oracle/emu/fake_cudart.c -- TEST INFRASTRUCTURE.
 *
 * A stand-in for the ~35 CUDA runtime entry points libvsb200 uses, so that the PRODUCT's unmodified host code (calibration, plans,
 * tile lists, launch sequences -- the same object files the real library is linked from) can run on a machine without a GPU:
 * "device" memory is host memory, copies are memcpy, streams and events are inert (everything executes in program order), and
 * every kernel launch is handed to a callback -- tests/test_emulated_pipeline.py registers the PTX interpreter there
 * (oracle/ptx_interp.py executes the kernel's PTX, compiled from the same .cu files with the product's flags).
 * Built by oracle/emu/Makefile into oracle/_build/libfakecudart.so; libvsb200_emu.so links against it instead of cudart.
 * Nothing here is part of, or reachable from, the shipped libvsb200.so. */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- registry of "device" allocations (bounds checks of the interpreter) ---------------------------------------------------- */
#define MAX_ALLOCS 4096
static struct { char *base; size_t size; } g_alloc[MAX_ALLOCS];
static int g_nalloc = 0;

void fake_register(void *p, size_t size)
{
    if (g_nalloc < MAX_ALLOCS) { g_alloc[g_nalloc].base = (char *)p; g_alloc[g_nalloc].size = size; ++g_nalloc; }
}
void fake_unregister(void *p)
{
    for (int i = 0; i < g_nalloc; ++i)
        if (g_alloc[i].base == (char *)p) { g_alloc[i] = g_alloc[--g_nalloc]; return; }
}
int fake_alloc_count(void) { return g_nalloc; }
void fake_alloc_get(int i, void **base, size_t *size) { *base = g_alloc[i].base; *size = g_alloc[i].size; }

/* ---- kernel registration (what nvcc's generated start-up code calls) and launches ------------------------------------------- */
#define MAX_FUNCS 512
static struct { const void *host; const char *name; } g_func[MAX_FUNCS];
static int g_nfunc = 0;
typedef int (*fake_launch_cb)(const char *name, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz, void **args, size_t smem);
static fake_launch_cb g_cb = 0;
static long g_launches = 0;
void fake_set_launch_callback(fake_launch_cb cb) { g_cb = cb; }
long fake_launch_count(void) { return g_launches; }

void **__cudaRegisterFatBinary(void *fatCubin) { static void *handle; (void)fatCubin; return &handle; }
void __cudaRegisterFatBinaryEnd(void **h) { (void)h; }
void __cudaUnregisterFatBinary(void **h) { (void)h; }
void __cudaRegisterFunction(void **h, const char *hostFun, char *deviceFun, const char *deviceName, int thread_limit, uint3 *tid, uint3 *bid,
                            dim3 *bDim, dim3 *gDim, int *wSize)
{
    (void)h; (void)deviceFun; (void)thread_limit; (void)tid; (void)bid; (void)bDim; (void)gDim; (void)wSize;
    if (g_nfunc < MAX_FUNCS) { g_func[g_nfunc].host = hostFun; g_func[g_nfunc].name = deviceName; ++g_nfunc; }
}
void __cudaRegisterVar(void **h, char *hostVar, char *deviceAddress, const char *deviceName, int ext, size_t size, int constant, int global)
{
    (void)h; (void)hostVar; (void)deviceAddress; (void)deviceName; (void)ext; (void)size; (void)constant; (void)global;
}
static struct { dim3 grid, block; size_t smem; void *stream; } g_cfg;
unsigned __cudaPushCallConfiguration(dim3 grid, dim3 block, size_t smem, struct CUstream_st *stream)
{
    g_cfg.grid = grid; g_cfg.block = block; g_cfg.smem = smem; g_cfg.stream = stream;
    return 0;
}
cudaError_t __cudaPopCallConfiguration(dim3 *grid, dim3 *block, size_t *smem, void *stream)
{
    *grid = g_cfg.grid; *block = g_cfg.block; *smem = g_cfg.smem; *(void **)stream = g_cfg.stream;
    return cudaSuccess;
}
static cudaError_t g_last = cudaSuccess;
cudaError_t cudaLaunchKernel(const void *func, dim3 grid, dim3 block, void **args, size_t smem, cudaStream_t stream)
{
    (void)stream;
    const char *name = 0;
    for (int i = 0; i < g_nfunc; ++i) if (g_func[i].host == func) { name = g_func[i].name; break; }
    if (!name || !g_cb) { fprintf(stderr, "fake_cudart: launch of an unregistered kernel or no callback\n"); return g_last = cudaErrorLaunchFailure; }
    ++g_launches;
    if (g_cb(name, grid.x, grid.y, grid.z, block.x, block.y, block.z, args, smem) != 0) return g_last = cudaErrorLaunchFailure;
    return cudaSuccess;
}

/* ---- memory ------------------------------------------------------------------------------------------------------------------- */
cudaError_t cudaMalloc(void **p, size_t size)
{
    *p = calloc(size ? size : 1, 1);   /* (real device memory is not zeroed; the product never relies on it -- the tests poison outputs) */
    if (!*p) return g_last = cudaErrorMemoryAllocation;
    fake_register(*p, size);
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) { if (p) { fake_unregister(p); free(p); } return cudaSuccess; }
cudaError_t cudaMallocAsync(void **p, size_t size, cudaStream_t s) { (void)s; return cudaMalloc(p, size); }
cudaError_t cudaFreeAsync(void *p, cudaStream_t s) { (void)s; return cudaFree(p); }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, enum cudaMemcpyKind k) { (void)k; memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, enum cudaMemcpyKind k, cudaStream_t st) { (void)st; return cudaMemcpy(d, s, n, k); }
cudaError_t cudaMemcpy2D(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, enum cudaMemcpyKind k)
{
    (void)k;
    for (size_t y = 0; y < h; ++y) memmove((char *)d + y * dp, (const char *)s + y * sp, w);
    return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, enum cudaMemcpyKind k, cudaStream_t st)
{
    (void)st;
    return cudaMemcpy2D(d, dp, s, sp, w, h, k);
}
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st) { (void)st; memset(d, v, n); return cudaSuccess; }

/* ---- device, streams, events: one inert device, program-order execution ------------------------------------------------------- */
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int *v, enum cudaDeviceAttr attr, int dev)
{
    (void)dev;
    switch (attr) {
    case cudaDevAttrMaxSharedMemoryPerBlockOptin: *v = 232448; break;
    case cudaDevAttrMultiProcessorCount: *v = 148; break;
    default: *v = 0; break;
    }
    return cudaSuccess;
}
cudaError_t cudaFuncSetAttribute(const void *f, enum cudaFuncAttribute a, int v) { (void)f; (void)a; (void)v; return cudaSuccess; }
cudaError_t cudaGetLastError(void) { cudaError_t e = g_last; g_last = cudaSuccess; return e; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "error (emulated runtime)"; }
cudaError_t cudaGetDriverEntryPoint(const char *sym, void **fn, unsigned long long flags, enum cudaDriverEntryPointQueryResult *q)
{
    (void)sym; (void)flags;
    *fn = 0;
    if (q) *q = cudaDriverEntryPointSymbolNotFound;   /* no tensor maps here: the product falls back to its plain-load pyramid kernel */
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned f) { (void)f; *s = (cudaStream_t)calloc(1, 8); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t s) { (void)s; return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned f) { (void)s; (void)e; (void)f; return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)calloc(1, 8); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned f) { (void)f; return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) { (void)e; (void)s; return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t e) { (void)e; return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { (void)a; (void)b; *ms = 0.f; return cudaSuccess; }
