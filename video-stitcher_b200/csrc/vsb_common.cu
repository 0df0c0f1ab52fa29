// vsb_common.cu -- error plumbing + host-side projector geometry (calibration-time, runs on the host exactly
// like the reference's detectResultRoi does).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>

#include "vsb_internal.h"

namespace vsb {

static thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_cuda(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return VSB_OK;
    return fail(VSB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

int check_launch(const char *what) { return check_cuda(cudaGetLastError(), what); }

static void mat3_mul(const float *a, const float *b, float *c)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += (double)a[i * 3 + k] * (double)b[k * 3 + j];
            c[i * 3 + j] = (float)s;
        }
}

// ProjectorBase::setCameraParams, sources/modules/stitching/src/warpers.cpp:49-79
void projector_setup(const float K[9], const float R[9], float k_rinv[9], float r_kinv[9], float rinv[9])
{
    float Rt[9], Kinv[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Rt[i * 3 + j] = R[j * 3 + i];
    double k[9];
    for (int i = 0; i < 9; ++i) k[i] = K[i];
    const double det = k[0] * (k[4] * k[8] - k[5] * k[7]) - k[1] * (k[3] * k[8] - k[5] * k[6]) + k[2] * (k[3] * k[7] - k[4] * k[6]);
    const double d = 1.0 / det;
    Kinv[0] = (float)((k[4] * k[8] - k[5] * k[7]) * d);
    Kinv[1] = (float)((k[2] * k[7] - k[1] * k[8]) * d);
    Kinv[2] = (float)((k[1] * k[5] - k[2] * k[4]) * d);
    Kinv[3] = (float)((k[5] * k[6] - k[3] * k[8]) * d);
    Kinv[4] = (float)((k[0] * k[8] - k[2] * k[6]) * d);
    Kinv[5] = (float)((k[2] * k[3] - k[0] * k[5]) * d);
    Kinv[6] = (float)((k[3] * k[7] - k[4] * k[6]) * d);
    Kinv[7] = (float)((k[1] * k[6] - k[0] * k[7]) * d);
    Kinv[8] = (float)((k[0] * k[4] - k[1] * k[3]) * d);
    std::memcpy(rinv, Rt, sizeof(Rt));
    mat3_mul(R, Kinv, r_kinv);
    mat3_mul(K, Rt, k_rinv);
}

// {Spherical,Cylindrical}Projector::mapForward, sources/modules/stitching/include/opencv2/stitching/detail/warpers_inl.hpp:243-253,274-283
static inline void map_forward(int proj, float scale, const float *r_kinv, float x, float y, float &u, float &v)
{
    const float x_ = r_kinv[0] * x + r_kinv[1] * y + r_kinv[2];
    const float y_ = r_kinv[3] * x + r_kinv[4] * y + r_kinv[5];
    const float z_ = r_kinv[6] * x + r_kinv[7] * y + r_kinv[8];
    u = scale * atan2f(x_, z_);
    if (proj == VSB_PROJ_SPHERICAL) {
        const float w = y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_);
        v = scale * (static_cast<float>(M_PI) - acosf(w == w ? w : 0));
    } else {
        v = scale * y_ / sqrtf(x_ * x_ + z_ * z_);
    }
}

}  // namespace vsb

extern "C" {

const char *vsb_last_error(void) { return vsb::g_err; }

const char *vsb_version(void) { return "vsb200 0.1 (sm_100a)"; }

int vsb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// RotationWarperBase::detectResultRoiByBorder (warpers_inl.hpp:176-210) and SphericalWarper::detectResultRoi
// (sources/modules/stitching/src/warpers.cpp:277-318); Rect as returned by warpRoi (warpers_inl.hpp:136-145).
int vsb_warp_roi(int projection, float scale, const float K[9], const float R[9], int src_w, int src_h, int roi[4])
{
    if (!K || !R || !roi || src_w <= 0 || src_h <= 0 || !(scale > 0) ||
        (projection != VSB_PROJ_SPHERICAL && projection != VSB_PROJ_CYLINDRICAL))
        return vsb::fail(VSB_ERR_INVALID, "warp_roi: bad arguments");
    float k_rinv[9], r_kinv[9], rinv[9];
    vsb::projector_setup(K, R, k_rinv, r_kinv, rinv);
    float tl_u = std::numeric_limits<float>::max(), tl_v = tl_u, br_u = -tl_u, br_v = -tl_u;
    auto acc = [&](float x, float y) {
        float u, v;
        vsb::map_forward(projection, scale, r_kinv, x, y, u, v);
        tl_u = std::min(tl_u, u); tl_v = std::min(tl_v, v);
        br_u = std::max(br_u, u); br_v = std::max(br_v, v);
    };
    for (float x = 0; x < src_w; ++x) { acc(x, 0.f); acc(x, static_cast<float>(src_h - 1)); }
    for (int y = 0; y < src_h; ++y) { acc(0.f, static_cast<float>(y)); acc(static_cast<float>(src_w - 1), static_cast<float>(y)); }
    int tlx = static_cast<int>(tl_u), tly = static_cast<int>(tl_v), brx = static_cast<int>(br_u), bry = static_cast<int>(br_v);
    if (projection == VSB_PROJ_SPHERICAL) {
        tl_u = static_cast<float>(tlx); tl_v = static_cast<float>(tly);
        br_u = static_cast<float>(brx); br_v = static_cast<float>(bry);
        for (int pass = 0; pass < 2; ++pass) {
            const float x = rinv[1], y = pass == 0 ? rinv[4] : -rinv[4], z = rinv[7];
            if (y > 0.f) {
                const float x_ = (K[0] * x + K[1] * y) / z + K[2];
                const float y_ = K[4] * y / z + K[5];
                if (x_ > 0.f && x_ < src_w && y_ > 0.f && y_ < src_h) {
                    const float pole = pass == 0 ? static_cast<float>(M_PI * scale) : 0.f;
                    tl_u = std::min(tl_u, 0.f); tl_v = std::min(tl_v, pole);
                    br_u = std::max(br_u, 0.f); br_v = std::max(br_v, pole);
                }
            }
        }
        tlx = static_cast<int>(tl_u); tly = static_cast<int>(tl_v); brx = static_cast<int>(br_u); bry = static_cast<int>(br_v);
    }
    roi[0] = tlx; roi[1] = tly; roi[2] = brx - tlx + 1; roi[3] = bry - tly + 1;
    return VSB_OK;
}

}  // extern "C"
