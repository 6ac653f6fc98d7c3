// test_host_ops.cpp — drives the C++ host operators exactly like the reference's only caller does
// (core/frontEnd/FE_SlamMonoV.cpp:104-122): per frame FtDt::detect, identity undistortion (pinhole), then
// FtAssoc::match(firstFrame, currFrame); with scaleNumFeatures(5.f) before the first frame like initOperators (:247).
// Input : raw file  [int32 n, h, w, nFeatures] + n*h*w bytes.   Output: raw file read back by tests/test_host_cpp.py
//         per frame: int32 monoIndex, int32 nObs, nObs*(nav24_kp 28 B), nObs*32 B; then per frame>0: int32 n1, n1*int32.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../nav24_b200/host/nav24_ops.hpp"

using namespace NAV24;

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s in.raw out.raw [scaleNumFeatures]\n", argv[0]); return 2; }
    FILE* fi = std::fopen(argv[1], "rb");
    if (!fi) return 2;
    int hdr[4];
    if (std::fread(hdr, 4, 4, fi) != 4) return 2;
    const int n = hdr[0], h = hdr[1], w = hdr[2], nf = hdr[3];
    std::vector<uint8_t> img((size_t)n * h * w);
    if (std::fread(img.data(), 1, img.size(), fi) != img.size()) return 2;
    std::fclose(fi);

    std::shared_ptr<OP::FtDtOrbB200> pOrbDetector;
    try {
        pOrbDetector = std::make_shared<OP::FtDtOrbB200>(nf, 1.2f, 8, 20, 7);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 3;
    }
    if (argc > 3) pOrbDetector->scaleNumFeatures((float)std::atof(argv[3]));
    OP::FtAssocB200 orbMatcher(pOrbDetector, FeatureGridCfg(w, h, 0.f, (float)w, 0.f, (float)h));

    FILE* fo = std::fopen(argv[2], "wb");
    std::vector<FramePtr> frames;
    for (int f = 0; f < n; ++f) {
        FramePtr pFrame = std::make_shared<FrameMonoGrid>(f * 0.05, img.data() + (size_t)f * h * w, w, h, (size_t)w);
        const int mono = pOrbDetector->detect(pFrame);
        const auto& obs = pFrame->getObservations();
        const int nObs = (int)obs.size();
        std::fwrite(&mono, 4, 1, fo); std::fwrite(&nObs, 4, 1, fo);
        for (const auto& o : obs) std::fwrite(&o->getKeyPoint(), sizeof(nav24_kp), 1, fo);
        for (const auto& o : obs) std::fwrite(o->getDescriptor().data(), 1, 32, fo);
        frames.push_back(pFrame);
    }
    for (int f = 1; f < n; ++f) {
        orbMatcher.match(frames[0], frames[f]);
        const auto pm = frames[f]->getMatches();
        const int n1 = (int)pm->mvMatches12.size();
        std::fwrite(&n1, 4, 1, fo);
        std::fwrite(pm->mvMatches12.data(), 4, n1, fo);
        if (pm->mpMatchedFrame.lock() != frames[0]) return 4;
    }
    // error behaviour of detect(): -1 on an empty image (OP_FtDtOrbSlam.cpp:851-852)
    FramePtr empty = std::make_shared<FrameMonoGrid>(0.0, nullptr, 0, 0, 0);
    const int rcEmpty = pOrbDetector->detect(empty);
    std::fwrite(&rcEmpty, 4, 1, fo);

    // the distorted-camera form of the same loop (FE_SlamMonoV.cpp:104-122): detect -> Calibration::undistort ->
    // grid bounds from the undistorted image corners -> matchV on the undistorted points
    CalibrationB200 calib(pOrbDetector, "radial-tangential", 458.654f, 457.296f, 367.215f, 248.375f,
                          {-0.28340811f, 0.07395907f, 0.00019359f, 1.76187114e-05f});
    const std::vector<float> b = calib.computeImageBounds(w, h);
    std::fwrite(b.data(), 4, 4, fo);
    for (int f = 0; f < 2 && f < n; ++f) {
        calib.undistort(frames[f]->getObservations());
        for (const auto& o : frames[f]->getObservations()) { const OB::Point2f p = o->getPointUd(); std::fwrite(&p, 4, 2, fo); }
    }
    if (n >= 2) {
        OP::FtAssocB200 udMatcher(pOrbDetector, FeatureGridCfg(w, h, b[0], b[1], b[2], b[3]));
        const std::vector<int> m = udMatcher.matchV(frames[0], frames[1]);
        const int n1 = (int)m.size();
        std::fwrite(&n1, 4, 1, fo);
        std::fwrite(m.data(), 4, n1, fo);
    }
    // struct-of-arrays observation store: same detector, undistortion and matcher, no per-keypoint
    // allocation; results must equal the object form above.  Host time of both forms is printed for the record.
    int soaOk = 1;
    {
        std::vector<FramePtr> soa;
        for (int f = 0; f < n; ++f) {
            FramePtr pFrame = std::make_shared<FrameMonoGrid>(f * 0.05, img.data() + (size_t)f * h * w, w, h, (size_t)w);
            pOrbDetector->detectSoA(pFrame);
            calib.undistort(*pFrame->getObservationStore());
            soa.push_back(pFrame);
            const auto& obs = frames[f]->getObservations();
            const OB::ObservationStore& st = *pFrame->getObservationStore();
            if (st.size() != obs.size()) soaOk = 0;
            for (size_t i = 0; i < st.size() && soaOk; ++i) {
                if (std::memcmp(&st[i].getKeyPoint(), &obs[i]->getKeyPoint(), sizeof(nav24_kp)) != 0) soaOk = 0;
                if (std::memcmp(st[i].getDescriptor(), obs[i]->getDescriptor().data(), 32) != 0) soaOk = 0;
                if (f < 2 && (st[i].getPointUd().x != obs[i]->getPointUd().x || st[i].getPointUd().y != obs[i]->getPointUd().y)) soaOk = 0;
            }
        }
        if (n >= 2) {
            OP::FtAssocB200 udMatcher(pOrbDetector, FeatureGridCfg(w, h, b[0], b[1], b[2], b[3]));
            if (udMatcher.matchV(soa[0], soa[1]) != udMatcher.matchV(frames[0], frames[1])) soaOk = 0;
        }
        // host-side materialisation cost per frame, object form vs store (the device work is identical)
        const int reps = 200;
        const OB::ObservationStore& st0 = *soa[0]->getObservationStore();
        auto t0 = std::chrono::steady_clock::now();
        size_t sink = 0;
        for (int r = 0; r < reps; ++r) {
            std::vector<OB::ObsPtr> v(st0.size());
            for (size_t i = 0; i < st0.size(); ++i) v[i] = std::make_shared<OB::KeyPoint2D>(st0.keypoints()[i], st0.descriptors() + 32 * i);
            sink += v.size();
        }
        auto t1 = std::chrono::steady_clock::now();
        OB::ObservationStore tmp; tmp.reserve(st0.size());
        for (int r = 0; r < reps; ++r) {
            std::memcpy(tmp.keypoints(), st0.keypoints(), st0.size() * sizeof(nav24_kp));
            std::memcpy(tmp.descriptors(), st0.descriptors(), st0.size() * 32);
            tmp.setSize(st0.size());
            sink += tmp.size();
        }
        auto t2 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[host] %zu keypoints/frame: KeyPoint2D objects %.1f us/frame, ObservationStore %.1f us/frame (%zu)\n",
                     st0.size(), std::chrono::duration<double, std::micro>(t1 - t0).count() / reps,
                     std::chrono::duration<double, std::micro>(t2 - t1).count() / reps, sink);
    }
    std::fwrite(&soaOk, 4, 1, fo);

    // image ingest: the frames go through the pinned slot ring (grey slots here; colour slots are covered
    // by tests/test_ingest.py) and must give the observations of the plain detect() above
    int ringOk = 1;
    {
        IngestRingB200 ring(pOrbDetector, w, h, 1, n + 1);
        for (int f = 0; f < n; ++f) std::memcpy(ring.slot(f + 1), img.data() + (size_t)f * h * w, (size_t)h * w);
        for (int f = 0; f < n; ++f) {
            FramePtr pFrame = std::make_shared<FrameMonoGrid>(f * 0.05, nullptr, w, h, (size_t)w);
            if (ring.detect(f + 1, pFrame) < 0) { ringOk = 0; break; }
            const OB::ObservationStore& st = *pFrame->getObservationStore();
            const auto& obs = frames[f]->getObservations();
            if (st.size() != obs.size()) ringOk = 0;
            for (size_t i = 0; i < st.size() && ringOk; ++i)
                if (std::memcmp(&st[i].getKeyPoint(), &obs[i]->getKeyPoint(), sizeof(nav24_kp)) != 0 ||
                    std::memcmp(st[i].getDescriptor(), obs[i]->getDescriptor().data(), 32) != 0) ringOk = 0;
        }
        if (ring.slot(n + 1) != nullptr) ringOk = 0;      // out of range
    }
    std::fwrite(&ringOk, 4, 1, fo);

    // two-view RANSAC scoring on the matches of frames 0 and 1: two homography and two fundamental
    // hypotheses (the true image shift of the synthetic sequence and a perturbed one); scores go to the checker
    if (n >= 2) {
        const std::vector<int>& m12 = frames[1]->getMatches()->mvMatches12;
        std::vector<float> xy1, xy2;
        const auto& o1 = frames[0]->getObservations(); const auto& o2 = frames[1]->getObservations();
        for (size_t i = 0; i < m12.size(); ++i)
            if (m12[i] >= 0) {
                xy1.push_back(o1[i]->getKeyPoint().x); xy1.push_back(o1[i]->getKeyPoint().y);
                xy2.push_back(o2[m12[i]]->getKeyPoint().x); xy2.push_back(o2[m12[i]]->getKeyPoint().y);
            }
        const float dx = -4.f, dy = -1.f;      // frame 1 is frame 0 shifted by (4, 1) px
        const std::vector<float> H21 = {1, 0, dx, 0, 1, dy, 0, 0, 1, 1.001f, 0.0005f, dx + 0.7f, -0.0004f, 0.999f, dy - 0.4f, 1e-6f, -1e-6f, 1};
        const std::vector<float> H12 = {1, 0, -dx, 0, 1, -dy, 0, 0, 1, 0.999f, -0.0005f, -dx - 0.7f, 0.0004f, 1.001f, -dy + 0.4f, -1e-6f, 1e-6f, 1};
        const std::vector<float> F21 = {0, 0, dy, 0, 0, -dx, -dy, dx, 0, 1e-7f, 0, dy * 1.1f, 0, 0, -dx, -dy, dx * 0.9f, 1e-3f};
        OP::TwoViewScorerB200 scorer(pOrbDetector, 1.f);
        OP::TwoViewScorerB200::Result r;
        const int ok = scorer.score(xy1, xy2, H21, H12, F21, r) ? 1 : 0;
        const int nm = (int)(xy1.size() / 2);
        std::fwrite(&ok, 4, 1, fo); std::fwrite(&nm, 4, 1, fo);
        std::fwrite(H21.data(), 4, 18, fo); std::fwrite(H12.data(), 4, 18, fo); std::fwrite(F21.data(), 4, 18, fo);
        std::fwrite(r.scoreH.data(), 4, 2, fo); std::fwrite(r.scoreF.data(), 4, 2, fo);
        std::fwrite(&r.bestH, 4, 1, fo); std::fwrite(&r.bestF, 4, 1, fo);
        std::fwrite(r.inliersH.data(), 1, 2 * (size_t)nm, fo); std::fwrite(r.inliersF.data(), 1, 2 * (size_t)nm, fo);
        // the kept-only form hands back the same scores, the same kept iteration and exactly its row of the masks
        OP::TwoViewScorerB200::Result k;
        int same = scorer.scoreKept(xy1, xy2, H21, H12, F21, k) && k.scoreH == r.scoreH && k.scoreF == r.scoreF && k.bestH == r.bestH &&
                   k.bestF == r.bestF && (int)k.inliersH.size() == nm && (int)k.inliersF.size() == nm ? 1 : 0;
        for (int i = 0; i < nm && same; ++i) {
            if (k.inliersH[(size_t)i] != (r.bestH >= 0 ? r.inliersH[(size_t)r.bestH * nm + i] : 0)) same = 0;
            if (k.inliersF[(size_t)i] != (r.bestF >= 0 ? r.inliersF[(size_t)r.bestF * nm + i] : 0)) same = 0;
        }
        std::fwrite(&same, 4, 1, fo);
    }
    std::fclose(fo);
    return 0;
}
