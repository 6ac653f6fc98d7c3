// test_ref_binding.cpp — the nav24-side binding (nav24_b200/host/ref_binding/OP_FtDtOrbB200.hpp, OP_FtAssocB200.hpp)
// compiled against the REFERENCE's own headers and run in one process next to the reference's own FtDtOrbSlam /
// FtAssocOrbSlam (compiled unchanged into oracle/_ref/libnav24_ref.so), on the reference's own Frame / KeyPoint2D /
// FeatureGrid / MatchedObs objects, in the order FE_SlamMonoV::handleImageMsg (core/frontEnd/FE_SlamMonoV.cpp:96-122)
// calls them: detect -> undistort (pinhole: identity) -> setObservations (grid) -> match(first, current).
// Built by oracle/Makefile.ref (needs /root/reference); the binary travels to the GPU box prebuilt.
//   usage: test_ref_binding in.raw [feature scale]      in.raw = int32 {n, H, W, nFeatures} + n grey frames
//          test_ref_binding --two-view in.raw           in.raw = int32 {n1, n2} + float xy1[n1][2], xy2[n2][2] + int32 matches12[n1]
//   exit 0 = every comparison held (a JSON summary on stdout), 1 = mismatch, 3 = no CUDA device (no CPU fallback)
// --two-view: TwoViewScoringB200.hpp (scoreHypothesesB200) against the reference's own TwoViewReconstruction — the
// hypotheses come from ITS Normalize / ComputeH21 / ComputeF21 on ITS RANSAC sets, the device scores all of them in one
// call, and the kept iteration / score / inliers / matrix must equal what ITS FindHomography / FindFundamental return.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "OP_FtDtOrbSlam.hpp"
#include "OP_FtAssocOrbSlam.hpp"
#include "OP_FtDtOrbB200.hpp"
#include "OP_FtAssocB200.hpp"
#include "TwoViewScoringB200.hpp"
#include "ref_access_2v.hpp"

using namespace NAV24;

namespace {
struct GridProbe : OB::FeatureGrid { static void reset() { mbInitImgBounds = false; } };

FramePtr frame_of(const uint8_t* px, int H, int W, double ts) {
    cv::Mat m;
    if (px) { m.create(H, W, CV_8UC1); for (int y = 0; y < H; ++y) memcpy(m.ptr(y), px + (size_t)y * W, (size_t)W); }
    auto pImg = std::make_shared<ImageTs>(m, ts, "");
    return std::make_shared<FrameMonoGrid>(ts, nullptr, std::vector<OB::ObsPtr>(), pImg);
}

// Calibration::undistort for a pinhole camera (Pinhole.hpp:75-78: identity) + the grid rebuild of FE_SlamMonoV.cpp:115
void undistort_identity(FramePtr& f) {
    auto obs = f->getObservations();
    for (auto& o : obs) { auto p = std::dynamic_pointer_cast<OB::Point2D>(o); p->setPointUd(p->getPoint()); p->updateDistorted(false); }
    f->setObservations(obs);
}

bool same_bits(float a, float b) { return memcmp(&a, &b, 4) == 0; }

int two_view_main(const char* path) {
    FILE* fi = fopen(path, "rb");
    if (!fi) { perror(path); return 2; }
    int hdr[2];
    if (fread(hdr, 4, 2, fi) != 2) return 2;
    const int n1 = hdr[0], n2 = hdr[1];
    std::vector<float> xy1((size_t)2 * n1), xy2((size_t)2 * n2);
    std::vector<int> m12((size_t)n1);
    if (fread(xy1.data(), 4, xy1.size(), fi) != xy1.size() || fread(xy2.data(), 4, xy2.size(), fi) != xy2.size() ||
        fread(m12.data(), 4, m12.size(), fi) != m12.size()) return 2;
    fclose(fi);
    std::shared_ptr<OP::FtDtOrbB200> det;
    try { det = std::make_shared<OP::FtDtOrbB200>(1000, 1.2f, 8, 20, 7); } catch (const std::exception& e) { fprintf(stderr, "%s\n", e.what()); return 3; }

    VK k1((size_t)n1), k2((size_t)n2);
    for (int i = 0; i < n1; ++i) k1[(size_t)i].pt = cv::Point2f(xy1[(size_t)2 * i], xy1[(size_t)2 * i + 1]);
    for (int i = 0; i < n2; ++i) k2[(size_t)i].pt = cv::Point2f(xy2[(size_t)2 * i], xy2[(size_t)2 * i + 1]);
    cv::Mat K = cv::Mat::eye(3, 3, CV_32F);
    K.at<float>(0, 0) = 458.f; K.at<float>(1, 1) = 457.f; K.at<float>(0, 2) = 367.f; K.at<float>(1, 2) = 248.f;
    OP::TwoViewReconstruction tvr(K, 1.f, 200);
    cv::Mat R21, t21;
    std::vector<cv::Point3f> p3d;
    VB tri;
    tvr.Reconstruct(k1, k2, m12, R21, t21, p3d, tri);      // fills mvKeys1/2, mvMatches12, mvSets (:69-127) and runs the reference pipeline once

    // the hypotheses of every iteration, as FindHomography (:286-303) / FindFundamental (:337-354) build them
    const VM& m = tvr.*get(Matches());
    const VS& sets = tvr.*get(Sets());
    const int iters = tvr.*get(MaxIt());
    const float sigma = tvr.*get(Sigma());
    VP pn1, pn2;
    cv::Mat T1, T2;
    (tvr.*get(Norm()))(tvr.*get(Keys1()), pn1, T1);
    (tvr.*get(Norm()))(tvr.*get(Keys2()), pn2, T2);
    const cv::Mat T2inv = T2.inv(), T2t = T2.t();
    std::vector<cv::Mat> H21s, H12s, F21s;
    VP a(8), b(8);
    for (int it = 0; it < iters; ++it) {
        for (int j = 0; j < 8; ++j) { const int idx = (int)sets[(size_t)it][(size_t)j]; a[(size_t)j] = pn1[(size_t)m[(size_t)idx].first]; b[(size_t)j] = pn2[(size_t)m[(size_t)idx].second]; }
        cv::Mat Hn = (tvr.*get(CompH()))(a, b);
        cv::Mat H21i = T2inv * Hn * T1;
        H21s.push_back(H21i); H12s.push_back(cv::Mat(H21i.inv()));
        cv::Mat Fn = (tvr.*get(CompF()))(a, b);
        F21s.push_back(cv::Mat(T2t * Fn * T1));
    }
    OP::TwoViewScoresB200 sc, kept;      // sc: every iteration's mask (for the per-hypothesis check); kept: the default call
    int rc = OP::scoreHypothesesB200(det->handle(), tvr.*get(Keys1()), tvr.*get(Keys2()), m, H21s, H12s, F21s, sigma, tvr.mParams2VR, sc, true);
    if (rc == NAV24_OK) rc = OP::scoreHypothesesB200(det->handle(), tvr.*get(Keys1()), tvr.*get(Keys2()), m, H21s, H12s, F21s, sigma, tvr.mParams2VR, kept);
    if (rc != NAV24_OK) { printf("{\"error\": \"nav24_two_view_score: %d %s\"}\n", rc, nav24_last_error_string(det->handle())); return 1; }

    const int N = (int)m.size();
    long scoreBad = 0, inlBad = 0, keptBad = 0;
    for (int it = 0; it < iters; ++it) {      // every hypothesis against the reference's own Check functions
        VB inl;
        const float sh = (tvr.*get(CheckH()))(H21s[(size_t)it], H12s[(size_t)it], inl, sigma);
        scoreBad += !same_bits(sh, sc.SH[(size_t)it]);
        for (int i = 0; i < N; ++i) inlBad += (inl[(size_t)i] ? 1 : 0) != sc.inliersH[(size_t)it * N + i];
        const float sf = (tvr.*get(CheckF()))(F21s[(size_t)it], inl, sigma);
        scoreBad += !same_bits(sf, sc.SF[(size_t)it]);
        for (int i = 0; i < N; ++i) inlBad += (inl[(size_t)i] ? 1 : 0) != sc.inliersF[(size_t)it * N + i];
    }
    // what FindHomography / FindFundamental return vs what the binding hands back in their place
    for (int model = 0; model < 4; ++model) {      // both result forms (all masks / kept mask only) x both models
        VB inlR, inlB;
        float sR = -1.f, sB = -1.f;
        cv::Mat MR, MB;
        const OP::TwoViewScoresB200& res = model < 2 ? sc : kept;
        if (model % 2 == 0) { (tvr.*get(FindH()))(inlR, sR, MR); res.keptHomography(H21s, inlB, sB, MB); }
        else { (tvr.*get(FindF()))(inlR, sR, MR); res.keptFundamental(F21s, inlB, sB, MB); }
        bool same = same_bits(sR, sB) && inlR == inlB && MR.empty() == MB.empty();
        if (same && !MR.empty()) for (int i = 0; i < 9; ++i) same = same && same_bits(MR.at<float>(i / 3, i % 3), MB.at<float>(i / 3, i % 3));
        keptBad += !same;
    }
    printf("{\"matches\": %d, \"iterations\": %d, \"best_h\": %d, \"best_f\": %d, \"score_h\": %.3f, \"score_f\": %.3f, "
           "\"score_mismatches\": %ld, \"inlier_mismatches\": %ld, \"kept_result_mismatches\": %ld}\n",
           N, iters, sc.bestH, sc.bestF, sc.bestH >= 0 ? sc.SH[(size_t)sc.bestH] : 0.f, sc.bestF >= 0 ? sc.SF[(size_t)sc.bestF] : 0.f, scoreBad, inlBad, keptBad);
    return scoreBad == 0 && inlBad == 0 && keptBad == 0 ? 0 : 1;
}
}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s in.raw [feature scale] | --two-view in.raw\n", argv[0]); return 2; }
    if (argc > 2 && strcmp(argv[1], "--two-view") == 0) return two_view_main(argv[2]);
    FILE* fi = fopen(argv[1], "rb");
    if (!fi) { perror(argv[1]); return 2; }
    int hdr[4];
    if (fread(hdr, 4, 4, fi) != 4) return 2;
    const int n = hdr[0], H = hdr[1], W = hdr[2], nf = hdr[3];
    std::vector<uint8_t> px((size_t)n * H * W);
    if (fread(px.data(), 1, px.size(), fi) != px.size()) return 2;
    fclose(fi);
    const float fscale = argc > 2 ? (float)atof(argv[2]) : 0.f;

    std::shared_ptr<OP::FtDtOrbB200> gpuDet;
    try {
        gpuDet = std::make_shared<OP::FtDtOrbB200>(nf, 1.2f, 8, 20, 7);
    } catch (const std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 3;
    }
    OP::FtDtPtr gpu = gpuDet;                                                        // used through the reference's interface
    OP::FtDtPtr cpu = std::make_shared<OP::FtDtOrbSlam>(nf, 1.2f, 8, 20, 7);
    if (fscale > 0.f) { gpu->scaleNumFeatures(fscale); cpu->scaleNumFeatures(fscale); }      // FE_SlamMonoV's x5 / x0.2 switch
    if (gpu->getNumFeatures() != cpu->getNumFeatures()) { printf("{\"error\": \"getNumFeatures\"}\n"); return 1; }
    std::shared_ptr<OP::FtAssoc> gpuMatcher = std::make_shared<OP::FtAssocB200>(gpuDet->handle());
    std::shared_ptr<OP::FtAssoc> cpuMatcher = std::make_shared<OP::FtAssocOrbSlam>();

    GridProbe::reset();      // FE_SlamMonoV.cpp:96-101; a pinhole camera's computeImageBounds is the image rectangle
    OB::FeatureGrid::setImageBounds(cv::Size(W, H), std::vector<float>{0.f, (float)W, 0.f, (float)H});

    long kpTotal = 0, kpBad = 0, descBad = 0, monoBad = 0, matchBad = 0, matchBadPure = 0, matchesTotal = 0, storedBad = 0;
    FramePtr firstG, firstC;
    for (int f = 0; f < n; ++f) {
        FramePtr fg = frame_of(px.data() + (size_t)f * H * W, H, W, f * 0.05), fc = frame_of(px.data() + (size_t)f * H * W, H, W, f * 0.05);
        const int mg = gpu->detect(fg), mc = cpu->detect(fc);
        monoBad += mg != mc;
        const auto& og = fg->getObservations();
        const auto& oc = fc->getObservations();
        if (og.size() != oc.size()) { printf("{\"error\": \"frame %d: %zu observations, the reference has %zu\"}\n", f, og.size(), oc.size()); return 1; }
        bool descEqual = true;
        for (size_t i = 0; i < og.size(); ++i) {
            auto a = std::dynamic_pointer_cast<OB::KeyPoint2D>(og[i]), b = std::dynamic_pointer_cast<OB::KeyPoint2D>(oc[i]);
            if (!a || !b) { kpBad++; continue; }
            const cv::KeyPoint &ka = a->getKeyPoint(), &kb = b->getKeyPoint();
            const bool same = same_bits(ka.pt.x, kb.pt.x) && same_bits(ka.pt.y, kb.pt.y) && same_bits(ka.size, kb.size) && same_bits(ka.angle, kb.angle) &&
                              same_bits(ka.response, kb.response) && ka.octave == kb.octave && ka.class_id == kb.class_id &&
                              same_bits(a->getPoint().x, b->getPoint().x) && same_bits(a->getPoint().y, b->getPoint().y);
            kpBad += !same;
            if (memcmp(a->getDescriptor().data, b->getDescriptor().data, 32) != 0) { descBad++; descEqual = false; }
            if (a->getFrame() != fg) kpBad++;      // setFrame (:929)
        }
        kpTotal += (long)og.size();
        undistort_identity(fg);
        undistort_identity(fc);
        if (!firstG) { firstG = fg; firstC = fc; continue; }
        // the B200 matcher against the reference's matcher on the SAME frames (the ones the B200 detector filled)
        std::vector<int> m = gpuMatcher->matchV(firstG, fg), r = cpuMatcher->matchV(firstG, fg);
        if (m.size() != r.size() || m.size() != firstG->getObservations().size()) { printf("{\"error\": \"matchV size\"}\n"); return 1; }
        for (size_t i = 0; i < m.size(); ++i) { matchBad += m[i] != r[i]; matchesTotal += m[i] >= 0; }
        // and against the all-reference pipeline (reference detector + reference matcher); only comparable when no
        // descriptor of the two frames differs (the 0.1 % budget of rotated-sample rounding)
        if (descEqual) {
            std::vector<int> p = cpuMatcher->matchV(firstC, fc);
            for (size_t i = 0; i < m.size(); ++i) matchBadPure += m[i] != p[i];
        }
        gpuMatcher->match(firstG, fg);      // stores a MatchedObs on frame 2 (OP_FtAssocOrbSlam.cpp:247-260)
        auto stored = std::dynamic_pointer_cast<FrameImgMono>(fg)->getMatches();
        int cnt = 0;
        for (int v : m) cnt += v >= 0;
        if (!stored || stored->mnMatches != cnt || stored->mvMatches12 != m || stored->mpMatchedFrame.lock() != firstG) storedBad++;
    }
    // empty image: -1 from both (OP_FtDtOrbSlam.cpp:851-852); a frame without a grid: {} from both (:103-107)
    FramePtr eg = frame_of(nullptr, 0, 0, 0.0), ec = frame_of(nullptr, 0, 0, 0.0);
    const int emptyG = gpu->detect(eg), emptyC = cpu->detect(ec);
    FramePtr plain = std::make_shared<FrameImgMono>(0.0, nullptr, std::vector<OB::ObsPtr>());
    const size_t noGridG = gpuMatcher->matchV(firstG, plain).size(), noGridC = cpuMatcher->matchV(firstG, plain).size();
    const bool edge = emptyG == -1 && emptyC == -1 && noGridG == 0 && noGridC == 0;

    printf("{\"frames\": %d, \"num_features\": %d, \"keypoints\": %ld, \"keypoint_mismatches\": %ld, \"mono_index_mismatches\": %ld, "
           "\"descriptor_mismatches\": %ld, \"matches\": %ld, \"match_mismatches\": %ld, \"match_mismatches_vs_all_reference\": %ld, "
           "\"stored_matchedobs_mismatches\": %ld, \"edge_cases_ok\": %s}\n",
           n, gpu->getNumFeatures(), kpTotal, kpBad, monoBad, descBad, matchesTotal, matchBad, matchBadPure, storedBad, edge ? "true" : "false");
    const bool ok = kpBad == 0 && monoBad == 0 && matchBad == 0 && matchBadPure == 0 && storedBad == 0 && edge && descBad * 1000 <= kpTotal;
    return ok ? 0 : 1;
}
