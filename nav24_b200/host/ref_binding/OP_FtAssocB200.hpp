// OP_FtAssocB200.hpp — the binding a nav24 maintainer adds as core/operators/objAssoc/OP_FtAssocB200.hpp
// (INTEGRATION.md §2).  Compiles against the REFERENCE's own headers (OP_FtAssoc.hpp, Frame.hpp, FeatureGrid.hpp,
// MatchedFeatures.hpp) and forwards FtAssoc::matchV (core/operators/objAssoc/OP_FtAssoc.hpp:20) to nav24_match_window;
// it replaces FtAssocOrbSlam (OP_FtAssocOrbSlam.cpp:91-260), constructed at FE_SlamMonoV.cpp:42.
// Compiled here by oracle/Makefile.ref against the reference tree and run next to the reference's FtAssocOrbSlam in one
// process (tests/cpp/test_ref_binding.cpp).
#ifndef NAV24_OP_FTASSOCB200_HPP
#define NAV24_OP_FTASSOCB200_HPP

#include <algorithm>
#include <cassert>
#include <cstring>
#include <memory>
#include <vector>

#include <glog/logging.h>

#include "OP_FtAssoc.hpp"
#include "FeatureGrid.hpp"
#include "Point2D.hpp"
#include "nav24_orb.h"

namespace NAV24::OP {

class FtAssocB200 : public FtAssoc {
public:
    // ctx: the detector's handle (FtDtOrbB200::handle()); the matcher shares its device and stream
    explicit FtAssocB200(nav24_orb* ctx, float nnratio = 0.6, bool checkOri = true)
        : mCtx(ctx), mfNNratio(nnratio), mbCheckOrientation(checkOri), windowSize(100.f) {}

    // FtAssocOrbSlam::matchV (OP_FtAssocOrbSlam.cpp:91-223): {} + warning without a grid frame (:103-107), else one
    // entry per observation of frame 1 (index into frame 2 or -1)
    std::vector<int> matchV(const FramePtr& pFrame1, const FramePtr& pFrame2) override {
        if (!std::dynamic_pointer_cast<FrameMonoGrid>(pFrame2)) {
            LOG(WARNING) << "FtAssocB200::match, no grid frame\n";
            return {};
        }
        const auto& vpObs1 = pFrame1->getObservations();
        const auto& vpObs2 = pFrame2->getObservations();
        pack(vpObs1, mK1, mUd1, mD1);
        pack(vpObs2, mK2, mUd2, mD2);
        const nav24_grid_cfg grid = GridConfig::get();      // FeatureGrid::setImageBounds' process-wide configuration
        std::vector<int> vnMatches12(vpObs1.size(), -1);
        const int rc = nav24_match_window(mCtx, mK1.data(), mUd1.data(), mD1.data(), (int)vpObs1.size(), mK2.data(), mUd2.data(),
                                          mD2.data(), (int)vpObs2.size(), &grid, windowSize, mfNNratio, /*TH_LOW*/ 50,
                                          mbCheckOrientation ? 1 : 0, vnMatches12.data());
        if (rc < 0) { LOG(ERROR) << "FtAssocB200::matchV: " << nav24_last_error_string(mCtx); return {}; }
        return vnMatches12;
    }

    // The two wrappers keep FtAssocOrbSlam's contracts (OP_FtAssocOrbSlam.cpp:225-260); both are matchV plus bookkeeping.
    // match(f1, f2, tracks): every accepted pair goes to the track container; returns the number of pairs (0 when matchV
    // had nothing to say).
    int match(const FramePtr& pFrame1, const FramePtr& pFrame2, OB::FtTracksPtr& pTracks) override {
        const std::vector<int> m12 = matchV(pFrame1, pFrame2);
        if (m12.empty()) return 0;
        const auto& obs1 = pFrame1->getObservations();
        const auto& obs2 = pFrame2->getObservations();
        assert(m12.size() == obs1.size());
        int nPairs = 0;
        for (size_t i1 = 0; i1 < m12.size(); ++i1) {
            if (m12[i1] < 0) continue;
            pTracks->addMatch(obs1[i1], obs2[(size_t)m12[i1]]);
            ++nPairs;
        }
        return nPairs;
    }

    // match(f1, f2): the result is left on frame 2 as a MatchedObs that refers (weakly) to frame 1.
    void match(const FramePtr& pFrame1, const FramePtr& pFrame2) override {
        const std::vector<int> m12 = matchV(pFrame1, pFrame2);
        const int nPairs = (int)std::count_if(m12.begin(), m12.end(), [](int v) { return v >= 0; });
        if (auto pImgFrame2 = std::dynamic_pointer_cast<FrameImgMono>(pFrame2))
            pImgFrame2->setMatches(std::make_shared<OB::MatchedObs>(pFrame1, m12, nPairs));
    }

protected:
    // FeatureGrid keeps its configuration in protected statics (FeatureGrid.hpp:36-46); a derived type may read them
    struct GridConfig : OB::FeatureGrid {
        static nav24_grid_cfg get() { return nav24_grid_cfg{mGridCols, mGridRows, mnMinX, mnMaxX, mnMinY, mnMaxY}; }
    };

    // getKeyPoint(), getPointUd(), getDescriptor() of every observation into the flat arrays of the C ABI.  An
    // observation that is no KeyPoint2D is neither a query (:116-119) nor a level-0 candidate: octave -1, far away.
    static void pack(const std::vector<OB::ObsPtr>& vpObs, std::vector<nav24_kp>& k, std::vector<float>& ud, std::vector<uint8_t>& d) {
        const size_t n = vpObs.size();
        k.resize(n); ud.resize(2 * n); d.resize(32 * n);
        for (size_t i = 0; i < n; i++) {
            auto p = std::dynamic_pointer_cast<OB::KeyPoint2D>(vpObs[i]);
            if (!p) {
                k[i] = nav24_kp{-1e9f, -1e9f, 0.f, 0.f, 0.f, 1 << 20, -1};
                ud[2 * i] = ud[2 * i + 1] = -1e9f;
                memset(&d[32 * i], 0, 32);
                continue;
            }
            const cv::KeyPoint& kp = p->getKeyPoint();
            k[i] = nav24_kp{kp.pt.x, kp.pt.y, kp.size, kp.angle, kp.response, kp.octave, kp.class_id};
            const cv::Point2f u = p->getPointUd();
            ud[2 * i] = u.x; ud[2 * i + 1] = u.y;
            memcpy(&d[32 * i], p->getDescriptor().data, 32);
        }
    }

    nav24_orb* mCtx;
    float mfNNratio;
    bool mbCheckOrientation;
    float windowSize;
    std::vector<nav24_kp> mK1, mK2;
    std::vector<float> mUd1, mUd2;
    std::vector<uint8_t> mD1, mD2;
};

}  // namespace NAV24::OP

#endif  // NAV24_OP_FTASSOCB200_HPP
