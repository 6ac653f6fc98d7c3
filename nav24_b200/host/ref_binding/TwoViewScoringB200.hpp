// TwoViewScoringB200.hpp — the nav24-side binding of the two-view RANSAC scoring,
// in the reference's own types.  TwoViewReconstruction::FindHomography / FindFundamental
// (core/operators/mapInit/OP_2ViewReconstruction.cpp:266-365) keep their minimal-set solvers on the host, collect the
// hypotheses of all iterations and hand them to ONE call that replaces 2 x mMaxIterations runs of CheckHomography (:447-530)
// and CheckFundamental (:532-610) and the `if (currentScore > score)` selection (:307-312, :358-363).
// Compiled by oracle/Makefile.ref against the reference tree and checked against the reference's own FindHomography /
// FindFundamental in one process (tests/cpp/test_ref_binding.cpp).
#ifndef NAV24_TWOVIEWSCORINGB200_HPP
#define NAV24_TWOVIEWSCORINGB200_HPP

#include <cstdint>
#include <utility>
#include <vector>

#include <opencv2/core.hpp>

#include "OP_2ViewReconstruction.hpp"
#include "nav24_orb.h"

namespace NAV24::OP {

struct TwoViewScoresB200 {
    std::vector<float> SH, SF;                  // per iteration: CheckHomography / CheckFundamental scores
    std::vector<uint8_t> inliersH, inliersF;    // allMasks: per iteration x match (vbCurrentInliers); else the kept iteration's only
    int bestH = -1, bestF = -1;                 // the iteration the reference's selection loop keeps (-1: no score above 0)
    int nMatches = 0;
    bool allMasks = false;

    // what FindHomography / FindFundamental return through their reference parameters
    void keptHomography(const std::vector<cv::Mat>& H21s, std::vector<bool>& vbMatchesInliers, float& score, cv::Mat& H21) const {
        keep(bestH, SH, inliersH, H21s, vbMatchesInliers, score, H21);
    }
    void keptFundamental(const std::vector<cv::Mat>& F21s, std::vector<bool>& vbMatchesInliers, float& score, cv::Mat& F21) const {
        keep(bestF, SF, inliersF, F21s, vbMatchesInliers, score, F21);
    }

private:
    void keep(int best, const std::vector<float>& S, const std::vector<uint8_t>& inl, const std::vector<cv::Mat>& Ms,
              std::vector<bool>& vbMatchesInliers, float& score, cv::Mat& M) const {
        score = 0.f;                                                   // :279, :330
        vbMatchesInliers.assign((size_t)nMatches, false);
        if (best < 0) return;
        score = S[(size_t)best];
        M = Ms[(size_t)best].clone();                                  // :309, :360
        const size_t row = allMasks ? (size_t)best * (size_t)nMatches : 0;
        for (int i = 0; i < nMatches; ++i) vbMatchesInliers[(size_t)i] = inl[row + (size_t)i] != 0;
    }
};

// H21s[i] / H12s[i] = T2inv*Hn*T1 and its inverse (:302-303), F21s[i] = T2t*Fn*T1 (:354): CV_32F 3 x 3, one per iteration.
// Either model may be left out (empty vectors).  allMasks = false (the default) brings back only the kept iteration's inlier
// mask, which is all FindHomography / FindFundamental return (nav24_two_view_score_kept: n_matches bytes per model over PCIe
// instead of iterations x n_matches); true brings back every iteration's mask.  Returns NAV24_OK or a NAV24_E_* code.
inline int scoreHypothesesB200(nav24_orb* ctx, const std::vector<cv::KeyPoint>& vKeys1, const std::vector<cv::KeyPoint>& vKeys2,
                               const std::vector<std::pair<int, int>>& vMatches12, const std::vector<cv::Mat>& H21s,
                               const std::vector<cv::Mat>& H12s, const std::vector<cv::Mat>& F21s, float sigma,
                               const Params2VR& prm, TwoViewScoresB200& out, bool allMasks = false) {
    const int N = (int)vMatches12.size();
    const int nHyp = (int)(H21s.empty() ? F21s.size() : H21s.size());
    std::vector<float> xy1((size_t)2 * N), xy2((size_t)2 * N);
    for (int i = 0; i < N; ++i) {
        const cv::Point2f& p1 = vKeys1[(size_t)vMatches12[(size_t)i].first].pt;      // :483-484, :559-560
        const cv::Point2f& p2 = vKeys2[(size_t)vMatches12[(size_t)i].second].pt;
        xy1[(size_t)2 * i] = p1.x; xy1[(size_t)2 * i + 1] = p1.y;
        xy2[(size_t)2 * i] = p2.x; xy2[(size_t)2 * i + 1] = p2.y;
    }
    auto flatten = [](const std::vector<cv::Mat>& Ms) {
        std::vector<float> v(Ms.size() * 9);
        for (size_t h = 0; h < Ms.size(); ++h)
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) v[h * 9 + (size_t)(3 * r + c)] = Ms[h].at<float>(r, c);
        return v;
    };
    const std::vector<float> h21 = flatten(H21s), h12 = flatten(H12s), f21 = flatten(F21s);
    out = TwoViewScoresB200();
    out.nMatches = N;
    out.allMasks = allMasks;
    const size_t rows = allMasks ? (size_t)nHyp : 1;
    if (!H21s.empty()) { out.SH.resize((size_t)nHyp); out.inliersH.resize(rows * (size_t)N); }
    if (!F21s.empty()) { out.SF.resize((size_t)nHyp); out.inliersF.resize(rows * (size_t)N); }
    return (allMasks ? nav24_two_view_score : nav24_two_view_score_kept)(ctx, xy1.data(), xy2.data(), N, h21.empty() ? nullptr : h21.data(), h12.empty() ? nullptr : h12.data(),
                                f21.empty() ? nullptr : f21.data(), nHyp, sigma, prm.mThChiSqScore, prm.mThChiSqF, prm.mThChiSqScore,
                                out.SH.empty() ? nullptr : out.SH.data(), out.SF.empty() ? nullptr : out.SF.data(),
                                out.inliersH.empty() ? nullptr : out.inliersH.data(), out.inliersF.empty() ? nullptr : out.inliersF.data(),
                                &out.bestH, &out.bestF);
}

}  // namespace NAV24::OP

#endif  // NAV24_TWOVIEWSCORINGB200_HPP
