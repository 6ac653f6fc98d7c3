// ref_driver_2v.cpp — TEST INFRASTRUCTURE.  extern "C" entry points over the REFERENCE's own
//   NAV24::OP::TwoViewReconstruction      core/operators/mapInit/OP_2ViewReconstruction.cpp
// compiled unchanged from /root/reference by oracle/Makefile.ref into oracle/_ref/libnav24_ref.so.
// What this pins: CheckHomography (:447-530) and CheckFundamental (:532-610) — plain float code of the reference that
// reads its matrices with at<float>() — and the `if (currentScore > score)` selection of FindHomography /
// FindFundamental (:266-365).  What it does NOT pin: the 8-point solvers' last bits; cv::SVD / cv::Mat algebra are this
// repo's stand-ins (oracle/ref_shim/opencv2/core_algebra.hpp), so the hypothesis matrices are test INPUTS produced by
// the reference's ComputeH21 / ComputeF21 over that algebra, not OpenCV-exact values.
// The members involved are private; they are reached through explicit template instantiation (ref_access_2v.hpp), so the
// reference header is included as it is.  Nothing in the product links, loads or imports this file.
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "ref_access_2v.hpp"

namespace {

cv::Mat mat3(const float* v) { cv::Mat m(3, 3, CV_32F); for (int i = 0; i < 9; ++i) m.at<float>(i / 3, i % 3) = v[i]; return m; }
void put3(const cv::Mat& m, float* v) { for (int i = 0; i < 9; ++i) v[i] = m.empty() ? 0.f : m.at<float>(i / 3, i % 3); }
void put_mask(const VB& b, uint8_t* out, int n) { for (int i = 0; i < n; ++i) out[i] = i < (int)b.size() && b[i] ? 1 : 0; }

}  // namespace

extern "C" {

void* ref_2v_create(const float* K, float sigma, int iterations) { return new TwoViewReconstruction(mat3(K), sigma, iterations); }
void ref_2v_destroy(void* h) { delete static_cast<TwoViewReconstruction*>(h); }

// The public entry point itself (:69-160): fills mvKeys1/2, mvMatches12 and the RANSAC sets (:100-127), runs
// FindHomography and FindFundamental in two threads (:133-134) and the model selection / reconstruction behind them.
// matches12[i] = index into the second view or -1.  Returns Reconstruct's bool; the number of putative matches in n_out.
int ref_2v_reconstruct(void* h, const float* xy1, int n1, const float* xy2, int n2, const int* matches12, int* n_out) {
    auto* t = static_cast<TwoViewReconstruction*>(h);
    VK k1(n1), k2(n2);
    for (int i = 0; i < n1; ++i) k1[i].pt = cv::Point2f(xy1[2 * i], xy1[2 * i + 1]);
    for (int i = 0; i < n2; ++i) k2[i].pt = cv::Point2f(xy2[2 * i], xy2[2 * i + 1]);
    std::vector<int> m12(matches12, matches12 + n1);
    cv::Mat R21, t21;
    std::vector<cv::Point3f> p3d;
    VB tri;
    const bool ok = t->Reconstruct(k1, k2, m12, R21, t21, p3d, tri);
    *n_out = (int)(t->*get(Matches())).size();
    return ok ? 1 : 0;
}

// The matched points in match order (what nav24_two_view_score takes) and the RANSAC sets of the last Reconstruct.
void ref_2v_get_matches(void* h, float* xy1, float* xy2, int* sets /* iterations x 8 */) {
    auto* t = static_cast<TwoViewReconstruction*>(h);
    const VM& m = t->*get(Matches());
    const VK& k1 = t->*get(Keys1());
    const VK& k2 = t->*get(Keys2());
    for (size_t i = 0; i < m.size(); ++i) {
        xy1[2 * i] = k1[m[i].first].pt.x; xy1[2 * i + 1] = k1[m[i].first].pt.y;
        xy2[2 * i] = k2[m[i].second].pt.x; xy2[2 * i + 1] = k2[m[i].second].pt.y;
    }
    const VS& s = t->*get(Sets());
    if (sets) for (size_t it = 0; it < s.size(); ++it) for (int j = 0; j < 8; ++j) sets[8 * it + j] = (int)s[it][j];
}

// CheckHomography / CheckFundamental on caller-given matrices (row-major 3 x 3), over the matches of the last Reconstruct.
float ref_2v_check_h(void* h, const float* H21, const float* H12, float sigma, uint8_t* inliers) {
    auto* t = static_cast<TwoViewReconstruction*>(h);
    VB inl;
    const float s = (t->*get(CheckH()))(mat3(H21), mat3(H12), inl, sigma);
    put_mask(inl, inliers, (int)(t->*get(Matches())).size());
    return s;
}
float ref_2v_check_f(void* h, const float* F21, float sigma, uint8_t* inliers) {
    auto* t = static_cast<TwoViewReconstruction*>(h);
    VB inl;
    const float s = (t->*get(CheckF()))(mat3(F21), inl, sigma);
    put_mask(inl, inliers, (int)(t->*get(Matches())).size());
    return s;
}

// The hypothesis matrices of every RANSAC iteration, produced by the reference's own Normalize / ComputeH21 / ComputeF21
// with the three statements FindHomography (:301-303) and FindFundamental (:352-354) wrap around them.
void ref_2v_hypotheses(void* h, float* H21, float* H12, float* F21) {
    auto* t = static_cast<TwoViewReconstruction*>(h);
    const VM& m = t->*get(Matches());
    const VS& sets = t->*get(Sets());
    VP pn1, pn2;
    cv::Mat T1, T2;
    (t->*get(Norm()))(t->*get(Keys1()), pn1, T1);
    (t->*get(Norm()))(t->*get(Keys2()), pn2, T2);
    cv::Mat T2inv = T2.inv(), T2t = T2.t();
    VP a(8), b(8);
    for (int it = 0; it < t->*get(MaxIt()); ++it) {
        for (int j = 0; j < 8; ++j) { const int idx = (int)sets[it][j]; a[j] = pn1[m[idx].first]; b[j] = pn2[m[idx].second]; }
        cv::Mat Hn = (t->*get(CompH()))(a, b);
        cv::Mat H21i = T2inv * Hn * T1;
        cv::Mat H12i = H21i.inv();
        cv::Mat Fn = (t->*get(CompF()))(a, b);
        cv::Mat F21i = T2t * Fn * T1;
        put3(H21i, H21 + 9 * it); put3(H12i, H12 + 9 * it); put3(F21i, F21 + 9 * it);
    }
}

// FindHomography / FindFundamental themselves (the RANSAC loops with their `if (currentScore > score)` selection).
void ref_2v_find_h(void* h, uint8_t* inliers, float* score, float* H21) {
    auto* t = static_cast<TwoViewReconstruction*>(h);
    VB inl; cv::Mat H;
    (t->*get(FindH()))(inl, *score, H);
    put_mask(inl, inliers, (int)(t->*get(Matches())).size());
    put3(H, H21);
}
void ref_2v_find_f(void* h, uint8_t* inliers, float* score, float* F21) {
    auto* t = static_cast<TwoViewReconstruction*>(h);
    VB inl; cv::Mat F;
    (t->*get(FindF()))(inl, *score, F);
    put_mask(inl, inliers, (int)(t->*get(Matches())).size());
    put3(F, F21);
}

}  // extern "C"
