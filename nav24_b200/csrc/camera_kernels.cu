// camera_kernels.cu — device side of Calibration::undistort (core/sensor/camera/Calibration.cpp:135-149), the step
// between detect and matchV in the front end (core/frontEnd/FE_SlamMonoV.cpp:104-122).
// The reference delegates to OpenCV: cv::undistortPoints (models/PinholeRadTan.cpp:20) and
// cv::fisheye::undistortPoints (models/KannalaBrandt8.cpp:238) with K float, D 4 floats, R = I, P = K
// (models/GeometricCamera.h:62-66).  Both are restated here in IEEE double with the reference's operation order;
// THIS FILE IS COMPILED WITH -fmad=false so that no multiply-add is contracted (OpenCV's scalar code is not).
// RadTan is bit-exact against the oracle / cv2; KB8 ends in tan(), whose CUDA and glibc double results may differ in
// the last bit, which survives the final rounding to float only when it straddles a float rounding boundary
// (tests allow 1 float ulp there and report the count).
#include "orb_internal.cuh"

namespace nav24 {
namespace {

__device__ __forceinline__ float2 undistort_radtan(const nav24_camera& c, float px, float py) {
    const double fx = c.fx, fy = c.fy, cx = c.cx, cy = c.cy;
    const double ifx = 1. / fx, ify = 1. / fy;
    const double k0 = c.d[0], k1 = c.d[1], k2 = c.d[2], k3 = c.d[3];
    double x = px, y = py;
    const double u = x, v = y;
    x = (x - cx) * ifx;
    y = (y - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; ++j) {          // TermCriteria(MAX_ITER, 5, 0.01): five fixed-point iterations
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((0. * r2 + 0.) * r2 + 0.) * r2) / (1 + ((0. * r2 + k1) * r2 + k0) * r2);
        if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
        const double deltaX = 2 * k2 * x * y + k3 * (r2 + 2 * x * x) + 0. * r2 + 0. * r2 * r2;
        const double deltaY = k2 * (r2 + 2 * y * y) + 2 * k3 * x * y + 0. * r2 + 0. * r2 * r2;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    const double xx = fx * x + 0. * y + cx, yy = 0. * x + fy * y + cy, ww = 1. / (0. * x + 0. * y + 1.);
    return make_float2((float)(xx * ww), (float)(yy * ww));
}

__device__ __forceinline__ float2 undistort_kb8(const nav24_camera& c, float px, float py) {
    const double fx = c.fx, fy = c.fy, cx = c.cx, cy = c.cy;
    const double k0 = c.d[0], k1 = c.d[1], k2 = c.d[2], k3 = c.d[3];
    const double eps = 1e-8, kPi2 = 3.1415926535897932384626433832795 / 2.;
    const double pwx = ((double)px - cx) / fx, pwy = ((double)py - cy) / fy;
    double theta_d = sqrt(pwx * pwx + pwy * pwy);
    theta_d = fmin(fmax(-kPi2, theta_d), kPi2);
    bool converged = false;
    double theta = theta_d, scale = 0.0;
    if (fabs(theta_d) > eps) {
        for (int j = 0; j < 10; ++j) {     // TermCriteria(COUNT + EPS, 10, 1e-8): Newton on theta
            const double theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta6 * theta2;
            const double k0_theta2 = k0 * theta2, k1_theta4 = k1 * theta4, k2_theta6 = k2 * theta6, k3_theta8 = k3 * theta8;
            const double theta_fix = (theta * (1 + k0_theta2 + k1_theta4 + k2_theta6 + k3_theta8) - theta_d) /
                                     (1 + 3 * k0_theta2 + 5 * k1_theta4 + 7 * k2_theta6 + 9 * k3_theta8);
            theta = theta - theta_fix;
            if (fabs(theta_fix) < eps) { converged = true; break; }
        }
        scale = tan(theta) / theta_d;
    } else {
        converged = true;
    }
    const bool flipped = (theta_d < 0 && theta > 0) || (theta_d > 0 && theta < 0);
    if (!(converged && !flipped)) return make_float2(-1000000.f, -1000000.f);
    const double pux = pwx * scale, puy = pwy * scale;
    const double prx = fx * pux + 0. * puy + cx * 1.0, pry = 0. * pux + fy * puy + cy * 1.0, prz = 0. * pux + 0. * puy + 1.0 * 1.0;
    return make_float2((float)(prx / prz), (float)(pry / prz));
}

__device__ __forceinline__ float2 undistort_one(const nav24_camera& c, float x, float y) {
    if (c.model == NAV24_CAM_RADTAN) return undistort_radtan(c, x, y);
    if (c.model == NAV24_CAM_KB8) return undistort_kb8(c, x, y);
    return make_float2(x, y);               // Pinhole.hpp:75-78
}

// plain point list
__global__ void __launch_bounds__(128) undistort_points_kernel(const nav24_camera cam, const float2* __restrict__ xy, int n,
                                                               float2* __restrict__ out) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i < n) { const float2 p = xy[i]; out[i] = undistort_one(cam, p.x, p.y); }
}

// the keypoints of a batch of frames, in place on the device: ud[f][i] for i < nOut[f]
__global__ void __launch_bounds__(128) undistort_frames_kernel(const nav24_camera cam, const nav24_kp* __restrict__ kps,
                                                               const int* __restrict__ nOut, int cap, float2* __restrict__ ud) {
    const int f = blockIdx.y, i = blockIdx.x * 128 + threadIdx.x;
    if (i < min(nOut[f], cap)) {
        const nav24_kp k = kps[(long long)f * cap + i];
        ud[(long long)f * cap + i] = undistort_one(cam, k.x, k.y);
    }
}

// ------------------------------------------------------------------------------------------
// Two-view RANSAC scoring: TwoViewReconstruction::CheckHomography (core/operators/mapInit/
// OP_2ViewReconstruction.cpp:447-530) and ::CheckFundamental (:532-610) for ALL hypotheses of FindHomography /
// FindFundamental (:266-365, 200 iterations each, run by two std::threads at :133-134) in one launch.
// One CTA = one hypothesis of one model (blockIdx.y: 0 = H, 1 = F).  The per-match symmetric transfer errors are
// independent and computed by all threads in the reference's operation order (this file is compiled with -fmad=false:
// plain IEEE float multiplies, adds and divides, like the reference's SSE code); the score is a float sum in match order
// (`score += th - chiSquare1; ... score += th - chiSquare2`), so one thread adds the staged terms sequentially — bit-equal
// to the reference's loop (a skipped term is added as +0.0f, which leaves a float sum unchanged).
// ------------------------------------------------------------------------------------------
constexpr int kTvChunk = 1024;      // matches staged per round

__global__ void __launch_bounds__(128) two_view_score_kernel(const float2* __restrict__ p1, const float2* __restrict__ p2, int n,
                                                             const float* __restrict__ H21, const float* __restrict__ H12,
                                                             const float* __restrict__ F21, float invSigmaSquare, float thH,
                                                             float thF, float thScore, float* __restrict__ scoreH,
                                                             float* __restrict__ scoreF, uint8_t* __restrict__ inH,
                                                             uint8_t* __restrict__ inF) {
    __shared__ float2 s_term[kTvChunk];
    const int hyp = blockIdx.x, tid = threadIdx.x;
    const bool isF = blockIdx.y == 1;
    if (isF ? (F21 == nullptr) : (H21 == nullptr)) return;
    float m[9], mi[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        m[k] = isF ? F21[hyp * 9 + k] : H21[hyp * 9 + k];
        mi[k] = isF ? 0.f : H12[hyp * 9 + k];
    }
    uint8_t* inl = isF ? inF : inH;
    float score = 0.f;
    for (int base = 0; base < n; base += kTvChunk) {
        const int cnt = min(kTvChunk, n - base);
        for (int j = tid; j < cnt; j += blockDim.x) {
            const float2 a = p1[base + j], b = p2[base + j];
            const float u1 = a.x, v1 = a.y, u2 = b.x, v2 = b.y;
            bool bIn = true;
            float t1, t2;
            if (!isF) {
                // x2in1 = H12 * x2   (:492-505)
                const float w2in1inv = 1.0f / (mi[6] * u2 + mi[7] * v2 + mi[8]);
                const float u2in1 = (mi[0] * u2 + mi[1] * v2 + mi[2]) * w2in1inv;
                const float v2in1 = (mi[3] * u2 + mi[4] * v2 + mi[5]) * w2in1inv;
                const float squareDist1 = (u1 - u2in1) * (u1 - u2in1) + (v1 - v2in1) * (v1 - v2in1);
                const float chiSquare1 = squareDist1 * invSigmaSquare;
                if (chiSquare1 > thH) { bIn = false; t1 = 0.f; } else t1 = thH - chiSquare1;
                // x1in2 = H21 * x1   (:507-521)
                const float w1in2inv = 1.0f / (m[6] * u1 + m[7] * v1 + m[8]);
                const float u1in2 = (m[0] * u1 + m[1] * v1 + m[2]) * w1in2inv;
                const float v1in2 = (m[3] * u1 + m[4] * v1 + m[5]) * w1in2inv;
                const float squareDist2 = (u2 - u1in2) * (u2 - u1in2) + (v2 - v1in2) * (v2 - v1in2);
                const float chiSquare2 = squareDist2 * invSigmaSquare;
                if (chiSquare2 > thH) { bIn = false; t2 = 0.f; } else t2 = thH - chiSquare2;
            } else {
                // l2 = F21 x1   (:571-585)
                const float a2 = m[0] * u1 + m[1] * v1 + m[2];
                const float b2 = m[3] * u1 + m[4] * v1 + m[5];
                const float c2 = m[6] * u1 + m[7] * v1 + m[8];
                const float num2 = a2 * u2 + b2 * v2 + c2;
                const float squareDist1 = num2 * num2 / (a2 * a2 + b2 * b2);
                const float chiSquare1 = squareDist1 * invSigmaSquare;
                if (chiSquare1 > thF) { bIn = false; t1 = 0.f; } else t1 = thScore - chiSquare1;
                // l1 = x2^T F21   (:587-603)
                const float a1 = m[0] * u2 + m[3] * v2 + m[6];
                const float b1 = m[1] * u2 + m[4] * v2 + m[7];
                const float c1 = m[2] * u2 + m[5] * v2 + m[8];
                const float num1 = a1 * u1 + b1 * v1 + c1;
                const float squareDist2 = num1 * num1 / (a1 * a1 + b1 * b1);
                const float chiSquare2 = squareDist2 * invSigmaSquare;
                if (chiSquare2 > thF) { bIn = false; t2 = 0.f; } else t2 = thScore - chiSquare2;
            }
            s_term[j] = make_float2(t1, t2);
            if (inl) inl[(size_t)hyp * n + base + j] = bIn ? 1 : 0;
        }
        __syncthreads();
        if (tid == 0)
            for (int j = 0; j < cnt; ++j) { score += s_term[j].x; score += s_term[j].y; }
        __syncthreads();
    }
    if (tid == 0) (isF ? scoreF : scoreH)[hyp] = score;
}

}  // namespace

int launch_two_view_score(const float* xy1, const float* xy2, int n, const float* H21, const float* H12, const float* F21, int nHyp,
                          float sigma, float thH, float thF, float thScore, float* scoreH, float* scoreF, uint8_t* inH, uint8_t* inF,
                          cudaStream_t s) {
    if (n <= 0 || nHyp <= 0) return 0;
    const float invSigmaSquare = 1.0f / (sigma * sigma);      // :477, :553
    dim3 grid(nHyp, 2);
    two_view_score_kernel<<<grid, 128, 0, s>>>(reinterpret_cast<const float2*>(xy1), reinterpret_cast<const float2*>(xy2), n, H21, H12,
                                               F21, invSigmaSquare, thH, thF, thScore, scoreH, scoreF, inH, inF);
    return 1;
}

int launch_undistort_points(const nav24_camera& cam, const float* xy, int n, float* out, cudaStream_t s) {
    if (n <= 0) return 0;
    undistort_points_kernel<<<(n + 127) / 128, 128, 0, s>>>(cam, reinterpret_cast<const float2*>(xy), n, reinterpret_cast<float2*>(out));
    return 1;
}

int launch_undistort_frames(const nav24_camera& cam, const nav24_kp* kps, const int* nOut, int cap, int B, float* ud, cudaStream_t s) {
    dim3 grid((cap + 127) / 128, B);
    undistort_frames_kernel<<<grid, 128, 0, s>>>(cam, kps, nOut, cap, reinterpret_cast<float2*>(ud));
    return 1;
}

}  // namespace nav24
