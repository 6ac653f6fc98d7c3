// ref_driver.cpp — TEST INFRASTRUCTURE.  extern "C" entry points over the REFERENCE's own classes, compiled unchanged
// from /root/reference by oracle/Makefile.ref into oracle/_ref/libnav24_ref.so:
//   NAV24::OP::FtDtOrbSlam        core/operators/objDetection/OP_FtDtOrbSlam.cpp (detect, DistributeOctTree, DivideNode, compareNodes)
//   NAV24::OP::FtAssocOrbSlam     core/operators/objAssoc/OP_FtAssocOrbSlam.cpp  (matchV, ComputeThreeMaxima, DescriptorDistance)
//   NAV24::OB::FeatureGrid        core/sensorData/observation/FeatureGrid.cpp    (assignFeaturesToGrid, getFeaturesInArea)
//   NAV24::FrameMonoGrid          core/dataTypes/frame/Frame.cpp                 (setObservations, getFeaturesInArea)
// Only the OpenCV *containers* are stand-ins (oracle/ref_shim); the pixel primitives forward to the oracle's routines
// that are pinned against live cv2.  tests/test_ref_build.py checks oracle == this library, which pins the oracle's
// restatement of nav24's own logic (cell loop, quadtree order, two-ended output order, grid, matcher) to the reference.
// Nothing in the product links, loads or imports this file.
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "OP_FtDtOrbSlam.hpp"
#include "OP_FtAssocOrbSlam.hpp"
#include "FeatureGrid.hpp"
#include "Point2D.hpp"

using namespace NAV24;

namespace {

struct OrbProbe : OP::FtDtOrbSlam {      // protected members of the reference class, for stage-level checks
    using OP::FtDtOrbSlam::FtDtOrbSlam;
    using OP::FtDtOrbSlam::DistributeOctTree;
    using OP::FtDtOrbSlam::mnFeaturesPerLevel;
    using OP::FtDtOrbSlam::mvScaleFactor;
    using OP::FtDtOrbSlam::mvInvScaleFactor;
    using OP::FtDtOrbSlam::umax;
    using OP::FtDtOrbSlam::mvImagePyramid;
    using OP::FtDtOrbSlam::ComputePyramid;
    using OP::FtDtOrbSlam::ComputeKeyPointsOctTree;
};

struct GridProbe : OB::FeatureGrid {      // the reference configures the grid once per process (FeatureGrid.cpp:100-113)
    static void reset() { mbInitImgBounds = false; }
};

struct ref_kp { float x, y, size, angle, response; int octave, class_id; };

FramePtr make_frame(const ref_kp* k, const float* ud, const uint8_t* d, int n) {
    auto fr = std::make_shared<FrameMonoGrid>(0.0, nullptr, std::vector<OB::ObsPtr>());
    std::vector<OB::ObsPtr> obs(n);
    for (int i = 0; i < n; ++i) {
        cv::KeyPoint kp(k[i].x, k[i].y, k[i].size, k[i].angle, k[i].response, k[i].octave, k[i].class_id);
        cv::Mat desc(1, 32, CV_8U);
        memcpy(desc.data, d + (size_t)32 * i, 32);
        auto p = std::make_shared<OB::KeyPoint2D>(kp, desc);
        p->setPointUd(cv::Point2f(ud[2 * i], ud[2 * i + 1]));      // Calibration::undistort (FE_SlamMonoV.cpp:115)
        obs[i] = p;
    }
    fr->setObservations(obs);      // builds the FeatureGrid (Frame.cpp:52-55)
    return fr;
}

}  // namespace

extern "C" {

void* ref_orb_create(int nfeatures, float scale, int nlevels, int iniTh, int minTh) {
    return new OrbProbe(nfeatures, scale, nlevels, iniTh, minTh);
}
void ref_orb_destroy(void* h) { delete (OrbProbe*)h; }
void ref_orb_scale_num_features(void* h, float s) { ((OrbProbe*)h)->scaleNumFeatures(s); }
int ref_orb_get_num_features(void* h) { return ((OrbProbe*)h)->getNumFeatures(); }
void ref_orb_tables(void* h, float* scale, float* inv, int* quota, int* umax) {
    OrbProbe* o = (OrbProbe*)h;
    for (size_t i = 0; i < o->mvScaleFactor.size(); ++i) { scale[i] = o->mvScaleFactor[i]; inv[i] = o->mvInvScaleFactor[i]; quota[i] = o->mnFeaturesPerLevel[i]; }
    for (size_t i = 0; i < o->umax.size(); ++i) umax[i] = o->umax[i];
}

// FtDtOrbSlam::detect on a grey image; returns monoIndex (or -1), *n_out = number of observations
int ref_orb_detect(void* h, const uint8_t* img, int w, int hh, size_t stride, ref_kp* kps, uint8_t* desc, int cap, int* n_out) {
    OrbProbe* o = (OrbProbe*)h;
    cv::Mat m;
    if (img && w > 0 && hh > 0) {
        m.create(hh, w, CV_8UC1);
        for (int y = 0; y < hh; ++y) memcpy(m.ptr(y), img + (size_t)y * stride, (size_t)w);
    }
    auto pImg = std::make_shared<ImageTs>(m, 0.0, "");
    FramePtr fr = std::make_shared<FrameMonoGrid>(0.0, nullptr, std::vector<OB::ObsPtr>(), pImg);
    const int mono = o->detect(fr);
    const auto& obs = fr->getObservations();
    *n_out = (int)obs.size();
    for (int i = 0; i < (int)obs.size() && i < cap; ++i) {
        auto p = std::dynamic_pointer_cast<OB::KeyPoint2D>(obs[i]);
        const cv::KeyPoint& k = p->getKeyPoint();
        kps[i] = ref_kp{k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id};
        memcpy(desc + (size_t)32 * i, p->getDescriptor().data, 32);
    }
    return mono;
}

void ref_orb_level_size(void* h, int l, int* w, int* hh) { OrbProbe* o = (OrbProbe*)h; *w = o->mvImagePyramid[l].cols; *hh = o->mvImagePyramid[l].rows; }
void ref_orb_get_level(void* h, int l, uint8_t* dst) {
    const cv::Mat& m = ((OrbProbe*)h)->mvImagePyramid[l];
    for (int y = 0; y < m.rows; ++y) memcpy(dst + (size_t)y * m.cols, m.ptr(y), (size_t)m.cols);
}

// FtDtOrbSlam::DistributeOctTree on a key list (x, y, response) in vToDistributeKeys order; returns the kept keys in
// the reference's output order as indices into the input
int ref_quadtree(const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N, int* kept, int cap) {
    OrbProbe o(1000, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> keys(n);
    for (int i = 0; i < n; ++i) { keys[i] = cv::KeyPoint(xyr[3 * i], xyr[3 * i + 1], 7.f, -1.f, xyr[3 * i + 2]); keys[i].class_id = i; }
    std::vector<cv::KeyPoint> out = o.DistributeOctTree(keys, minX, maxX, minY, maxY, N, 0);
    for (int i = 0; i < (int)out.size() && i < cap; ++i) kept[i] = out[i].class_id;
    return (int)out.size();
}

// FeatureGrid::setImageBounds + FtAssocOrbSlam::matchV; grid = {width, height, minX, maxX, minY, maxY} exactly as
// FE_SlamMonoV hands them over (Size of the image, Calibration::computeImageBounds)
int ref_match_window(const ref_kp* k1, const float* ud1, const uint8_t* d1, int n1, const ref_kp* k2, const float* ud2,
                     const uint8_t* d2, int n2, int imgW, int imgH, const float* bounds4, float nnratio, int checkOri,
                     int* matches12) {
    if (imgW > 0) {      // imgW <= 0: keep the grid configuration of the last call (worker threads of bench.py share one)
        GridProbe::reset();
        OB::FeatureGrid::setImageBounds(cv::Size(imgW, imgH), std::vector<float>(bounds4, bounds4 + 4));
    }
    FramePtr f1 = make_frame(k1, ud1, d1, n1), f2 = make_frame(k2, ud2, d2, n2);
    OP::FtAssocOrbSlam m(nnratio, checkOri != 0);
    std::vector<int> v = m.matchV(f1, f2);
    int nm = 0;
    for (int i = 0; i < (int)v.size(); ++i) { matches12[i] = v[i]; nm += v[i] >= 0; }
    // match(f1, f2) stores a MatchedObs on frame 2 (OP_FtAssocOrbSlam.cpp:247-260): exercised for the count
    m.match(f1, f2);
    auto mo = std::dynamic_pointer_cast<FrameImgMono>(f2)->getMatches();
    return mo && mo->mnMatches == nm ? nm : -1;
}

// FrameMonoGrid::getFeaturesInArea around (x, y) of frame-2 observations
int ref_grid_query(const ref_kp* k2, const float* ud2, int n2, int imgW, int imgH, const float* bounds4, float x, float y, float r,
                   int minLevel, int maxLevel, int* out, int cap) {
    GridProbe::reset();
    OB::FeatureGrid::setImageBounds(cv::Size(imgW, imgH), std::vector<float>(bounds4, bounds4 + 4));
    std::vector<uint8_t> d((size_t)32 * (n2 > 0 ? n2 : 1));
    FramePtr f2 = make_frame(k2, ud2, d.data(), n2);
    auto q = std::make_shared<OB::KeyPoint2D>(cv::KeyPoint(x, y, 31.f), cv::Mat(1, 32, CV_8U));
    q->setPointUd(cv::Point2f(x, y));
    std::vector<size_t> v = std::dynamic_pointer_cast<FrameMonoGrid>(f2)->getFeaturesInArea(q, r, minLevel, maxLevel);
    for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = (int)v[i];
    return (int)v.size();
}

}  // extern "C"
