// OP_FtDtOrbB200.hpp — the binding a nav24 maintainer adds as core/operators/objDetection/OP_FtDtOrbB200.hpp
// (INTEGRATION.md §2).  Compiles against the REFERENCE's own headers (OP_FtDt.hpp, Frame.hpp, Image.hpp, Point2D.hpp)
// and forwards FtDt::detect (core/operators/objDetection/OP_FtDt.hpp:18) to the C ABI of libnav24orb.so; it replaces
// FtDtOrbSlam (OP_FtDtOrbSlam.cpp:844-934) under `type: "orb"` in FtDt::create (OP_FtDt.cpp:43-48).
// Not built into libnav24orb.so (the library has no OpenCV / nav24 types in it): this header lives on the nav24 side.
// In this repository it is compiled by oracle/Makefile.ref against the reference tree + the container stand-ins and
// run next to the reference's FtDtOrbSlam in one process (tests/cpp/test_ref_binding.cpp, tests/test_ref_binding.py).
#ifndef NAV24_OP_FTDTORBB200_HPP
#define NAV24_OP_FTDTORBB200_HPP

#include <cassert>
#include <memory>
#include <stdexcept>
#include <vector>

#include <glog/logging.h>

#include "OP_FtDt.hpp"
#include "Image.hpp"
#include "Point2D.hpp"
#include "nav24_orb.h"

namespace NAV24::OP {

class FtDtOrbB200 : public FtDt {
public:
    FtDtOrbB200(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device = 0) : FtDt(nfeatures) {
        nav24_orb_params p{nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, 0};
        const int rc = nav24_orb_create(&p, device, &mCtx);
        if (rc != NAV24_OK) {      // the B200 path has no CPU fallback: fail loudly
            mCtx = nullptr;
            LOG(ERROR) << "FtDtOrbB200: nav24_orb_create failed (" << rc << "): no CUDA device / bad parameters";
            throw std::runtime_error("FtDtOrbB200: nav24_orb_create failed (no CPU fallback)");
        }
    }
    ~FtDtOrbB200() { if (mCtx) nav24_orb_destroy(mCtx); }
    FtDtOrbB200(const FtDtOrbB200&) = delete;
    FtDtOrbB200& operator=(const FtDtOrbB200&) = delete;

    // FtDtOrbSlam::detect (OP_FtDtOrbSlam.cpp:844-934): -1 on an empty image (:851-852), CV_8UC1 asserted (:853), the
    // frame's observation vector REPLACED by freshly allocated KeyPoint2D objects (:924-931), returns monoIndex (:933).
    int detect(FramePtr& pFrame) override {
        auto pImgFrame = std::dynamic_pointer_cast<FrameImgMono>(pFrame);
        if (!pImgFrame || !pImgFrame->getImage()) return -1;
        const cv::Mat& image = pImgFrame->getImage()->mImage;
        if (image.empty()) return -1;
        assert(image.type() == CV_8UC1);
        const int cap = nav24_orb_max_keypoints(mCtx);
        int n = 0;
        mKps.resize((size_t)cap);
        cv::Mat descriptors(cap, 32, CV_8U);
        const int mono = nav24_orb_detect(mCtx, image.data, image.cols, image.rows, image.step, mKps.data(), descriptors.data, cap, &n);
        if (mono < 0) { LOG(ERROR) << "FtDtOrbB200::detect: " << nav24_last_error_string(mCtx); return -1; }
        std::vector<OB::ObsPtr> vpObservations((size_t)n);
        for (int i = 0; i < n; i++) {
            const nav24_kp& k = mKps[(size_t)i];
            cv::KeyPoint kp(k.x, k.y, k.size, k.angle, k.response, k.octave, k.class_id);
            auto pObs = std::make_shared<OB::KeyPoint2D>(kp, descriptors.row(i));      // clones the row (Point2D.hpp:39-40)
            pObs->setFrame(pFrame);
            vpObservations[(size_t)i] = pObs;
        }
        pFrame->setObservations(vpObservations);
        return mono;
    }

    nav24_orb* handle() const { return mCtx; }

protected:
    void setNumFeatures(const int nFt) override {      // FtDtOrbSlam::setNumFeatures (:962-976): quotas follow
        FtDt::setNumFeatures(nFt);
        if (nav24_orb_set_num_features(mCtx, nFt) != NAV24_OK) LOG(ERROR) << "FtDtOrbB200::setNumFeatures: " << nav24_last_error_string(mCtx);
    }

    nav24_orb* mCtx = nullptr;
    std::vector<nav24_kp> mKps;
};

}  // namespace NAV24::OP

#endif  // NAV24_OP_FTDTORBB200_HPP
