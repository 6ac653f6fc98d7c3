// test_ref_binding.cpp — the nav24-side binding (nav24_b200/host/ref_binding/OP_FtDtOrbB200.hpp, OP_FtAssocB200.hpp)
// compiled against the REFERENCE's own headers and run in one process next to the reference's own FtDtOrbSlam /
// FtAssocOrbSlam (compiled unchanged into oracle/_ref/libnav24_ref.so), on the reference's own Frame / KeyPoint2D /
// FeatureGrid / MatchedObs objects, in the order FE_SlamMonoV::handleImageMsg (core/frontEnd/FE_SlamMonoV.cpp:96-122)
// calls them: detect -> undistort (pinhole: identity) -> setObservations (grid) -> match(first, current).
// Built by oracle/Makefile.ref (needs /root/reference); the binary travels to the GPU box prebuilt.
//   usage: test_ref_binding in.raw [feature scale]      in.raw = int32 {n, H, W, nFeatures} + n grey frames
//   exit 0 = every comparison held (a JSON summary on stdout), 1 = mismatch, 3 = no CUDA device (no CPU fallback)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "OP_FtDtOrbSlam.hpp"
#include "OP_FtAssocOrbSlam.hpp"
#include "OP_FtDtOrbB200.hpp"
#include "OP_FtAssocB200.hpp"

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
}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s in.raw [feature scale]\n", argv[0]); return 2; }
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
