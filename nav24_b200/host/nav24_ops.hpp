// nav24_ops.hpp — C++ host side of the B200 ORB front end: the reference's operator interface for the hot path,
// implemented on the C ABI of include/nav24_orb.h.
//
// The reference is compiled C++ (C++20, OpenCV containers).  OpenCV, Eigen and glog are not in this image, so this
// header mirrors the reference classes with dependency-free stand-ins of the same names, argument meaning and
// error behaviour; INTEGRATION.md shows the (shorter) subclasses a nav24 maintainer adds inside the real tree,
// where Frame / KeyPoint2D / MatchedObs are the reference's own types.
//
//   reference class (core/...)                                      here
//   OP::FtDt            operators/objDetection/OP_FtDt.hpp:14-29      OP::FtDt (same virtuals, same protected members)
//   OP::FtDtOrbSlam     operators/objDetection/OP_FtDtOrbSlam.hpp     OP::FtDtOrbB200
//   OP::FtAssoc         operators/objAssoc/OP_FtAssoc.hpp:15-22       OP::FtAssoc
//   OP::FtAssocOrbSlam  operators/objAssoc/OP_FtAssocOrbSlam.hpp      OP::FtAssocB200
//   OP::FtAssocOCV      operators/objAssoc/OP_FtAssoc.cpp:63-99       OP::FtAssocBfB200 (intended semantics)
//   OB::KeyPoint2D      sensorData/observation/Point2D.hpp:37-69      OB::KeyPoint2D (cv::KeyPoint -> nav24_kp, Mat -> 32 B)
//   OB::MatchedObs      sensorData/observation/MatchedFeatures.hpp    OB::MatchedObs
//   FrameMonoGrid       dataTypes/frame/Frame.hpp:59-74               FrameMonoGrid (image view instead of ImagePtr)
//   OB::FeatureGrid cfg sensorData/observation/FeatureGrid.cpp:100    FeatureGridCfg (explicit, not process-global)
//   Calibration         sensor/camera/Calibration.cpp:135-233         CalibrationB200 (undistort, computeImageBounds)
//   vector<ObsPtr>      OP_FtDtOrbSlam.cpp:925-931, Point2D.hpp:37-69 OB::ObservationStore (struct of arrays, same accessors)
//   imread + clones     sensor/camera/Camera.cpp:454-455, Image.hpp:22,   IngestRingB200 (pinned slots the decoder writes into, device
//                       frontEnd/FE_SlamMonoV.cpp:90-94                   BGR -> grey)
//   CheckHomography /   operators/mapInit/OP_2ViewReconstruction.cpp      OP::TwoViewScorerB200 (all RANSAC hypotheses in one call)
//   CheckFundamental    :447-610
//
// There is no CPU fallback: constructing FtDtOrbB200 without a CUDA device throws std::runtime_error.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nav24_orb.h"

namespace NAV24 {

namespace OB {

struct Point2f { float x = 0, y = 0; };

// = OB::KeyPoint2D: cv::KeyPoint + cloned 1x32 descriptor + the undistorted point (Point2D::mPointUd)
class KeyPoint2D {
public:
    KeyPoint2D(const nav24_kp& kpt, const uint8_t* desc) : mKPt(kpt) {
        std::memcpy(mDesc.data(), desc, 32);
        mPoint = {kpt.x, kpt.y};
        mPointUd = mPoint;
    }
    const nav24_kp& getKeyPoint() const { return mKPt; }
    const std::array<uint8_t, 32>& getDescriptor() const { return mDesc; }
    Point2f getPoint() const { return mPoint; }
    Point2f getPointUd() const { return mPointUd; }
    void setPointUd(const Point2f& p) { mPointUd = p; }      // Calibration::undistort writes this (Calibration.cpp:135-149)
private:
    nav24_kp mKPt;
    std::array<uint8_t, 32> mDesc;
    Point2f mPoint, mPointUd;
};
typedef std::shared_ptr<KeyPoint2D> ObsPtr;

// Struct-of-arrays observation store.  The reference materialises every keypoint as a
// make_shared<KeyPoint2D> holding a cloned 1x32 cv::Mat (OP_FtDtOrbSlam.cpp:925-931, Point2D.hpp:39-40): N small
// allocations per frame, which dominate the host time once detection runs on the GPU.  The store keeps the three
// arrays the C ABI reads and writes (nav24_kp, 32-byte descriptors, undistorted points) and hands out per-observation
// views with KeyPoint2D's accessors, so detect writes into it and undistort / matchV read it without any copy.
class ObservationStore {
public:
    class View {      // the accessors of OB::KeyPoint2D on element i
    public:
        View(const ObservationStore* s, size_t i) : mS(s), mI(i) {}
        const nav24_kp& getKeyPoint() const { return mS->mKps[mI]; }
        const uint8_t* getDescriptor() const { return &mS->mDesc[32 * mI]; }
        Point2f getPoint() const { return {mS->mKps[mI].x, mS->mKps[mI].y}; }
        Point2f getPointUd() const { return {mS->mUd[2 * mI], mS->mUd[2 * mI + 1]}; }
    private:
        const ObservationStore* mS; size_t mI;
    };
    size_t size() const { return mN; }
    View operator[](size_t i) const { return View(this, i); }
    void reserve(size_t cap) { if (mKps.size() < cap) { mKps.resize(cap); mDesc.resize(32 * cap); mUd.resize(2 * cap); } }
    void setSize(size_t n) {      // after the detector filled keypoints(): undistorted point = detected point (Point2D ctor)
        mN = n;
        for (size_t i = 0; i < n; ++i) { mUd[2 * i] = mKps[i].x; mUd[2 * i + 1] = mKps[i].y; }
    }
    nav24_kp* keypoints() { return mKps.data(); }
    const nav24_kp* keypoints() const { return mKps.data(); }
    uint8_t* descriptors() { return mDesc.data(); }
    const uint8_t* descriptors() const { return mDesc.data(); }
    float* pointsUd() { return mUd.data(); }
    const float* pointsUd() const { return mUd.data(); }
private:
    size_t mN = 0;
    std::vector<nav24_kp> mKps;
    std::vector<uint8_t> mDesc;
    std::vector<float> mUd;
};
typedef std::shared_ptr<ObservationStore> ObsStorePtr;

}  // namespace OB

class FrameMonoGrid;
typedef std::shared_ptr<FrameMonoGrid> FramePtr;

namespace OB {
// = OB::MatchedObs (MatchedFeatures.hpp:75-87): matches12 against a weakly referenced first frame
struct MatchedObs {
    int mnMatches = 0;
    std::weak_ptr<FrameMonoGrid> mpMatchedFrame;
    std::vector<int> mvMatches12;
};
typedef std::shared_ptr<MatchedObs> MatchedObsPtr;
}  // namespace OB

// FeatureGrid::setImageBounds (FeatureGrid.cpp:100-113) for an undistorted W x H image
inline nav24_grid_cfg FeatureGridCfg(int W, int H, float minX, float maxX, float minY, float maxY) {
    nav24_grid_cfg g;
    g.cols = W / 10; g.rows = H / 10; g.min_x = minX; g.max_x = maxX; g.min_y = minY; g.max_y = maxY;
    return g;
}

// = FrameMonoGrid (grey image + observations + matches); the image is a non-owning view like cv::Mat's header
class FrameMonoGrid : public std::enable_shared_from_this<FrameMonoGrid> {
public:
    FrameMonoGrid(double ts, const uint8_t* gray, int w, int h, size_t stride) : mTs(ts), mImg(gray), mW(w), mH(h), mStride(stride) {}
    double getTs() const { return mTs; }
    const uint8_t* image() const { return mImg; }
    int width() const { return mW; }
    int height() const { return mH; }
    size_t stride() const { return mStride; }
    const std::vector<OB::ObsPtr>& getObservations() const { return mvpObservations; }
    void setObservations(const std::vector<OB::ObsPtr>& v) { mvpObservations = v; }
    const OB::ObsStorePtr& getObservationStore() const { return mpStore; }      // struct-of-arrays form (detectSoA)
    void setObservationStore(const OB::ObsStorePtr& s) { mpStore = s; }
    void setMatches(const OB::MatchedObsPtr& m) { mpMatches12 = m; }
    OB::MatchedObsPtr getMatches() const { return mpMatches12; }
private:
    double mTs;
    const uint8_t* mImg;
    int mW, mH;
    size_t mStride;
    std::vector<OB::ObsPtr> mvpObservations;
    OB::ObsStorePtr mpStore;
    OB::MatchedObsPtr mpMatches12;
};

namespace OP {

// ---- detector -------------------------------------------------------------------------------------------
class FtDt {      // OP_FtDt.hpp:14-29
public:
    explicit FtDt(int nFt) : mnFeatures(nFt), mnIniNumFts(nFt) {}
    virtual ~FtDt() = default;
    virtual int detect(FramePtr& pFrame) = 0;
    int getNumFeatures() const { return mnFeatures; }
    void scaleNumFeatures(const float& scale) { this->setNumFeatures((int)(scale * mnIniNumFts)); }
protected:
    virtual void setNumFeatures(const int nFt) { mnFeatures = nFt; }
    int mnFeatures;
    const int mnIniNumFts;
};
typedef std::shared_ptr<FtDt> FtDtPtr;

class FtDtOrbB200 : public FtDt {      // replaces FtDtOrbSlam (OP_FtDtOrbSlam.hpp:27-84)
public:
    // same argument list and defaults as FtDtOrbSlam's constructor / FtDt::create (OP_FtDt.cpp:31-48)
    FtDtOrbB200(int nfeatures = 1000, float scaleFactor = 1.2f, int nlevels = 8, int iniThFAST = 20, int minThFAST = 7,
                int device = 0)
        : FtDt(nfeatures), mnLevels(nlevels) {
        nav24_orb_params p{nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, 0};
        const int rc = nav24_orb_create(&p, device, &mCtx);
        if (rc != NAV24_OK) throw std::runtime_error("FtDtOrbB200: nav24_orb_create failed (" + std::to_string(rc) +
                                                     "); a CUDA device is required, there is no CPU fallback");
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mnFeaturesPerLevel.resize(nlevels);
        refreshTables();
    }
    ~FtDtOrbB200() override { nav24_orb_destroy(mCtx); }
    FtDtOrbB200(const FtDtOrbB200&) = delete;
    FtDtOrbB200& operator=(const FtDtOrbB200&) = delete;

    // FtDtOrbSlam::detect (OP_FtDtOrbSlam.cpp:844-934): -1 on an empty image, otherwise monoIndex; replaces the
    // frame's observations with freshly allocated KeyPoint2D objects in the reference's two-ended order.
    int detect(FramePtr& pFrame) override {
        if (!pFrame || !pFrame->image() || pFrame->width() <= 0 || pFrame->height() <= 0) return -1;
        int cap = nav24_orb_max_keypoints(mCtx), n = 0;
        mKps.resize(cap); mDesc.resize((size_t)cap * 32);
        int rc = nav24_orb_detect(mCtx, pFrame->image(), pFrame->width(), pFrame->height(), pFrame->stride(), mKps.data(),
                                  mDesc.data(), cap, &n);
        if (rc == NAV24_E_CAPACITY) {
            cap = n; mKps.resize(cap); mDesc.resize((size_t)cap * 32);
            rc = nav24_orb_detect(mCtx, pFrame->image(), pFrame->width(), pFrame->height(), pFrame->stride(), mKps.data(),
                                  mDesc.data(), cap, &n);
        }
        if (rc < 0) { mLastError = nav24_last_error_string(mCtx); return rc == NAV24_E_BADARG ? -1 : rc; }
        std::vector<OB::ObsPtr> vpObservations(n);
        for (int i = 0; i < n; i++) vpObservations[i] = std::make_shared<OB::KeyPoint2D>(mKps[i], &mDesc[(size_t)i * 32]);
        pFrame->setObservations(vpObservations);
        return rc;      // monoIndex
    }

    // Same contract as detect(), but the frame receives a struct-of-arrays OB::ObservationStore that the library
    // writes directly (no per-keypoint allocation; the store of a recycled frame is reused).
    int detectSoA(FramePtr& pFrame) {
        if (!pFrame || !pFrame->image() || pFrame->width() <= 0 || pFrame->height() <= 0) return -1;
        OB::ObsStorePtr st = pFrame->getObservationStore();
        if (!st) { st = std::make_shared<OB::ObservationStore>(); pFrame->setObservationStore(st); }
        int cap = nav24_orb_max_keypoints(mCtx), n = 0;
        st->reserve(cap);
        int rc = nav24_orb_detect(mCtx, pFrame->image(), pFrame->width(), pFrame->height(), pFrame->stride(), st->keypoints(),
                                  st->descriptors(), cap, &n);
        if (rc == NAV24_E_CAPACITY) {
            cap = n; st->reserve(cap);
            rc = nav24_orb_detect(mCtx, pFrame->image(), pFrame->width(), pFrame->height(), pFrame->stride(), st->keypoints(),
                                  st->descriptors(), cap, &n);
        }
        if (rc < 0) { st->setSize(0); mLastError = nav24_last_error_string(mCtx); return rc == NAV24_E_BADARG ? -1 : rc; }
        st->setSize(n);
        return rc;      // monoIndex
    }

    int GetLevels() const { return mnLevels; }
    float GetScaleFactor() const { return mvScaleFactor.size() > 1 ? mvScaleFactor[1] : 1.f; }
    const std::vector<float>& GetScaleFactors() const { return mvScaleFactor; }
    const std::vector<float>& GetInverseScaleFactors() const { return mvInvScaleFactor; }
    const std::vector<int>& GetFeaturesPerLevel() const { return mnFeaturesPerLevel; }
    const std::string& lastError() const { return mLastError; }
    nav24_orb* handle() { return mCtx; }

protected:
    void setNumFeatures(const int nFt) override {      // FtDtOrbSlam::setNumFeatures (OP_FtDtOrbSlam.cpp:962-976)
        FtDt::setNumFeatures(nFt);
        nav24_orb_set_num_features(mCtx, nFt);
        refreshTables();
    }
    void refreshTables() { nav24_orb_get_tables(mCtx, mvScaleFactor.data(), mvInvScaleFactor.data(), mnFeaturesPerLevel.data()); }

    nav24_orb* mCtx = nullptr;
    int mnLevels;
    std::vector<float> mvScaleFactor, mvInvScaleFactor;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<nav24_kp> mKps;
    std::vector<uint8_t> mDesc;
    std::string mLastError;
};

// ---- matchers -------------------------------------------------------------------------------------------
class FtAssoc {      // OP_FtAssoc.hpp:15-22 (the FtTracks overload returns the match count; tracks stay host-side)
public:
    virtual ~FtAssoc() = default;
    virtual void match(const FramePtr& pFrame1, const FramePtr& pFrame2) = 0;
    virtual std::vector<int> matchV(const FramePtr& pFrame1, const FramePtr& pFrame2) = 0;
};

class FtAssocB200 : public FtAssoc {      // replaces FtAssocOrbSlam (OP_FtAssocOrbSlam.hpp:13-38)
public:
    // The reference keeps the grid configuration in process-global statics (FeatureGrid.cpp:15-18); here it is explicit.
    FtAssocB200(const std::shared_ptr<FtDtOrbB200>& ctxOwner, const nav24_grid_cfg& grid, float nnratio = 0.6f,
                bool checkOri = true)
        : mpOwner(ctxOwner), mGrid(grid), mfNNratio(nnratio), mbCheckOrientation(checkOri), windowSize(100.f) {}

    // FtAssocOrbSlam::matchV (OP_FtAssocOrbSlam.cpp:91-223): one int per observation of frame 1, -1 or an index into
    // frame 2; {} if a frame is missing (the reference returns {} + LOG(WARNING) when f2 has no grid, :103-107).
    std::vector<int> matchV(const FramePtr& pFrame1, const FramePtr& pFrame2) override {
        if (!pFrame1 || !pFrame2) return {};
        if (pFrame1->getObservationStore() && pFrame2->getObservationStore()) {      // struct-of-arrays frames: no packing
            const OB::ObservationStore& s1 = *pFrame1->getObservationStore();
            const OB::ObservationStore& s2 = *pFrame2->getObservationStore();
            std::vector<int> m12(s1.size(), -1);
            if (s1.size() == 0) return m12;
            const int rcS = nav24_match_window(mpOwner->handle(), s1.keypoints(), s1.pointsUd(), s1.descriptors(), (int)s1.size(),
                                               s2.keypoints(), s2.pointsUd(), s2.descriptors(), (int)s2.size(), &mGrid, windowSize,
                                               mfNNratio, TH_LOW, mbCheckOrientation ? 1 : 0, m12.data());
            if (rcS < 0) return {};
            return m12;
        }
        const auto& o1 = pFrame1->getObservations();
        const auto& o2 = pFrame2->getObservations();
        const int n1 = (int)o1.size(), n2 = (int)o2.size();
        std::vector<int> matches12(n1, -1);
        if (n1 == 0) return matches12;
        pack(o1, mK1, mU1, mD1); pack(o2, mK2, mU2, mD2);
        const int rc = nav24_match_window(mpOwner->handle(), mK1.data(), mU1.data(), mD1.data(), n1, mK2.data(), mU2.data(),
                                          mD2.data(), n2, &mGrid, windowSize, mfNNratio, TH_LOW, mbCheckOrientation ? 1 : 0,
                                          matches12.data());
        if (rc < 0) return {};
        return matches12;
    }
    // FtAssocOrbSlam::match(f1, f2) (:247-260): stores a MatchedObs (weak ref to f1) on f2
    void match(const FramePtr& pFrame1, const FramePtr& pFrame2) override {
        std::vector<int> matches12 = this->matchV(pFrame1, pFrame2);
        int nMatches = 0;
        for (const auto& m : matches12) nMatches += m >= 0;
        auto pMatchedObs = std::make_shared<OB::MatchedObs>();
        pMatchedObs->mnMatches = nMatches; pMatchedObs->mvMatches12 = matches12; pMatchedObs->mpMatchedFrame = pFrame1;
        if (pFrame2) pFrame2->setMatches(pMatchedObs);
    }
    static constexpr int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;      // OP_FtAssocOrbSlam.cpp:13-15

private:
    static void pack(const std::vector<OB::ObsPtr>& obs, std::vector<nav24_kp>& k, std::vector<float>& ud,
                     std::vector<uint8_t>& d) {
        const size_t n = obs.size();
        k.resize(n); ud.resize(2 * n); d.resize(32 * n);
        for (size_t i = 0; i < n; ++i) {
            k[i] = obs[i]->getKeyPoint();
            const OB::Point2f p = obs[i]->getPointUd();
            ud[2 * i] = p.x; ud[2 * i + 1] = p.y;
            std::memcpy(&d[32 * i], obs[i]->getDescriptor().data(), 32);
        }
    }
    std::shared_ptr<FtDtOrbB200> mpOwner;
    nav24_grid_cfg mGrid;
    float mfNNratio;
    bool mbCheckOrientation;
    float windowSize;
    std::vector<nav24_kp> mK1, mK2;
    std::vector<float> mU1, mU2;
    std::vector<uint8_t> mD1, mD2;
};

}  // namespace OP

// ---- camera -----------------------------------------------------------------------------------------------
// The part of Calibration the front end runs between detect and matchV (FE_SlamMonoV.cpp:115): undistort every
// observation and, once per sequence, the image bounds of the feature grid.  distType as in the reference's YAML:
// "radial-tangential" (cv::undistortPoints), "kannala-brandt8" (cv::fisheye::undistortPoints), anything else = calibrated
// input, identity (Calibration::isCalibrated, Calibration.cpp:230-233).
class CalibrationB200 {
public:
    CalibrationB200(const std::shared_ptr<OP::FtDtOrbB200>& ctxOwner, const std::string& distType, float fx, float fy, float cx,
                    float cy, const std::vector<float>& distCoefs = {})
        : mpOwner(ctxOwner), mDistType(distType) {
        mCam.model = distType == "radial-tangential" ? NAV24_CAM_RADTAN : distType == "kannala-brandt8" ? NAV24_CAM_KB8 : NAV24_CAM_PINHOLE;
        mCam.fx = fx; mCam.fy = fy; mCam.cx = cx; mCam.cy = cy;
        for (int i = 0; i < 4; ++i) mCam.d[i] = i < (int)distCoefs.size() ? distCoefs[i] : 0.f;      // mD_cv is 4x1 (GeometricCamera.h:64)
    }
    bool isCalibrated() const { return mCam.model == NAV24_CAM_PINHOLE; }
    const nav24_camera& camera() const { return mCam; }
    // Calibration::undistort(vector<ObsPtr>) (Calibration.cpp:135-159): sets every observation's undistorted point
    std::vector<OB::ObsPtr> undistort(const std::vector<OB::ObsPtr>& vpObs) {
        const size_t n = vpObs.size();
        mXY.resize(2 * n);
        for (size_t i = 0; i < n; ++i) { const OB::Point2f p = vpObs[i]->getPoint(); mXY[2 * i] = p.x; mXY[2 * i + 1] = p.y; }
        if (n && nav24_undistort_points(mpOwner->handle(), &mCam, mXY.data(), (int)n, mXY.data()) != NAV24_OK)
            throw std::runtime_error(nav24_last_error_string(mpOwner->handle()));
        for (size_t i = 0; i < n; ++i) vpObs[i]->setPointUd({mXY[2 * i], mXY[2 * i + 1]});
        return vpObs;
    }
    // the struct-of-arrays form: undistorted points written in place into the store
    void undistort(OB::ObservationStore& st) {
        const size_t n = st.size();
        mXY.resize(2 * n);
        for (size_t i = 0; i < n; ++i) { mXY[2 * i] = st.keypoints()[i].x; mXY[2 * i + 1] = st.keypoints()[i].y; }
        if (n && nav24_undistort_points(mpOwner->handle(), &mCam, mXY.data(), (int)n, st.pointsUd()) != NAV24_OK)
            throw std::runtime_error(nav24_last_error_string(mpOwner->handle()));
    }
    // Calibration::computeImageBounds (Calibration.cpp:196-228): {minX, maxX, minY, maxY}
    std::vector<float> computeImageBounds(int cols, int rows) {
        if (isCalibrated()) return {0.f, (float)cols, 0.f, (float)rows};
        float c[8] = {0.f, 0.f, (float)cols, 0.f, 0.f, (float)rows, (float)cols, (float)rows};
        if (nav24_undistort_points(mpOwner->handle(), &mCam, c, 4, c) != NAV24_OK)
            throw std::runtime_error(nav24_last_error_string(mpOwner->handle()));
        return {std::min(c[0], c[4]), std::max(c[2], c[6]), std::min(c[1], c[3]), std::max(c[5], c[7])};
    }
private:
    std::shared_ptr<OP::FtDtOrbB200> mpOwner;
    std::string mDistType;
    nav24_camera mCam{};
    std::vector<float> mXY;
};

namespace OP {

// Intended semantics of FtAssocOCV::match (OP_FtAssoc.cpp:63-99): kNN-2 + ratio 0.7, lowest train index wins ties.
class FtAssocBfB200 {
public:
    explicit FtAssocBfB200(const std::shared_ptr<FtDtOrbB200>& ctxOwner, int norm = NAV24_NORM_L2_U8, float ratio = 0.7f)
        : mpOwner(ctxOwner), mNorm(norm), mRatio(ratio) {}
    std::vector<int> matchV(const FramePtr& f1, const FramePtr& f2) {
        const auto& o1 = f1->getObservations();
        const auto& o2 = f2->getObservations();
        const int n1 = (int)o1.size(), n2 = (int)o2.size();
        std::vector<uint8_t> d1(32 * (size_t)n1), d2(32 * (size_t)n2), pass(n1);
        for (int i = 0; i < n1; ++i) std::memcpy(&d1[32 * (size_t)i], o1[i]->getDescriptor().data(), 32);
        for (int i = 0; i < n2; ++i) std::memcpy(&d2[32 * (size_t)i], o2[i]->getDescriptor().data(), 32);
        std::vector<int32_t> i0(n1), i1(n1);
        std::vector<float> f0(n1), f1v(n1);
        std::vector<int> m(n1, -1);
        if (n1 == 0) return m;
        if (nav24_match_bf_knn2(mpOwner->handle(), d1.data(), n1, d2.data(), n2, mNorm, mRatio, i0.data(), i1.data(), f0.data(),
                                f1v.data(), pass.data()) < 0)
            return {};
        for (int i = 0; i < n1; ++i) if (pass[i]) m[i] = i0[i];
        return m;
    }
private:
    std::shared_ptr<FtDtOrbB200> mpOwner;
    int mNorm;
    float mRatio;
};

}  // namespace OP
// ---- image ingest ---------------------------------------------------------------------------------
// The pinned slots CamOffline::run decodes into (cv::imdecode(..., &slotMat)) instead of imread + four clones; colour
// slots are converted to grey on the device.  detect(slot, frame) is FtDtOrbB200::detectSoA on a slot.
class IngestRingB200 {
public:
    IngestRingB200(const std::shared_ptr<OP::FtDtOrbB200>& det, int width, int height, int channels, int slots)
        : mpDet(det), mW(width), mH(height), mCh(channels), mSlots(slots) {
        if (nav24_ingest_create(det->handle(), width, height, channels, slots, &mRing) != NAV24_OK)
            throw std::runtime_error(std::string("IngestRingB200: ") + nav24_last_error_string(det->handle()));
    }
    ~IngestRingB200() { nav24_ingest_destroy(mRing); }
    IngestRingB200(const IngestRingB200&) = delete;
    IngestRingB200& operator=(const IngestRingB200&) = delete;

    uint8_t* slot(int k) { return nav24_ingest_slot(mRing, k); }
    size_t slotBytes() const { return nav24_ingest_slot_bytes(mRing); }
    int slots() const { return mSlots; }

    // detect on slot k into the frame's observation store; returns monoIndex or -1 (like FtDt::detect)
    int detect(int k, FramePtr& pFrame) {
        auto st = std::make_shared<OB::ObservationStore>();
        const int cap = nav24_orb_max_keypoints(mpDet->handle());
        st->reserve((size_t)cap);
        int n = 0, mono = 0;
        int rc = nav24_ingest_detect(mRing, k, 1, st->keypoints(), st->descriptors(), cap, &n, &mono);
        if (rc == NAV24_E_CAPACITY) {      // the first call of a shape learns the exact bound
            st->reserve((size_t)nav24_orb_max_keypoints(mpDet->handle()));
            rc = nav24_ingest_detect(mRing, k, 1, st->keypoints(), st->descriptors(), nav24_orb_max_keypoints(mpDet->handle()), &n, &mono);
        }
        if (rc < 0) return -1;
        st->setSize((size_t)n);
        pFrame->setObservationStore(st);
        return mono;
    }

private:
    std::shared_ptr<OP::FtDtOrbB200> mpDet;
    nav24_ingest* mRing = nullptr;
    int mW, mH, mCh, mSlots;
};

namespace OP {

// ---- two-view RANSAC scoring ------------------------------------------------------------------
// TwoViewReconstruction::CheckHomography / CheckFundamental (OP_2ViewReconstruction.cpp:447-610) for every iteration of
// FindHomography / FindFundamental at once; members mirror Params2VR (OP_2ViewReconstruction.hpp:46-65).
class TwoViewScorerB200 {
public:
    explicit TwoViewScorerB200(const std::shared_ptr<FtDtOrbB200>& det, float sigma = 1.f) : mpDet(det), mSigma(sigma) {}

    struct Result { std::vector<float> scoreH, scoreF; std::vector<uint8_t> inliersH, inliersF; int bestH = -1, bestF = -1; };

    // xy1 / xy2: n matched points (x, y); H21 / H12 / F21: nHyp x 9 row-major (F21 or the H pair may be empty)
    bool score(const std::vector<float>& xy1, const std::vector<float>& xy2, const std::vector<float>& H21,
               const std::vector<float>& H12, const std::vector<float>& F21, Result& r) const {
        const int n = (int)(xy1.size() / 2), nHyp = (int)(std::max(H21.size(), F21.size()) / 9);
        r = Result();
        if (!H21.empty()) { r.scoreH.resize(nHyp); r.inliersH.resize((size_t)nHyp * n); }
        if (!F21.empty()) { r.scoreF.resize(nHyp); r.inliersF.resize((size_t)nHyp * n); }
        return nav24_two_view_score(mpDet->handle(), xy1.data(), xy2.data(), n, H21.empty() ? nullptr : H21.data(),
                                    H12.empty() ? nullptr : H12.data(), F21.empty() ? nullptr : F21.data(), nHyp, mSigma, mThChiSqScore,
                                    mThChiSqF, mThChiSqScore, r.scoreH.data(), r.scoreF.data(), r.inliersH.data(), r.inliersF.data(),
                                    &r.bestH, &r.bestF) == NAV24_OK;
    }

    // What FindHomography / FindFundamental return: r.inliersH / r.inliersF hold ONLY the kept iteration's mask (n bytes each,
    // all 0 when r.bestH / r.bestF is -1) — nav24_two_view_score_kept, n bytes per model over PCIe instead of nHyp x n.
    bool scoreKept(const std::vector<float>& xy1, const std::vector<float>& xy2, const std::vector<float>& H21,
                   const std::vector<float>& H12, const std::vector<float>& F21, Result& r) const {
        const int n = (int)(xy1.size() / 2), nHyp = (int)(std::max(H21.size(), F21.size()) / 9);
        r = Result();
        if (!H21.empty()) { r.scoreH.resize(nHyp); r.inliersH.resize((size_t)n); }
        if (!F21.empty()) { r.scoreF.resize(nHyp); r.inliersF.resize((size_t)n); }
        return nav24_two_view_score_kept(mpDet->handle(), xy1.data(), xy2.data(), n, H21.empty() ? nullptr : H21.data(),
                                         H12.empty() ? nullptr : H12.data(), F21.empty() ? nullptr : F21.data(), nHyp, mSigma,
                                         mThChiSqScore, mThChiSqF, mThChiSqScore, r.scoreH.data(), r.scoreF.data(), r.inliersH.data(),
                                         r.inliersF.data(), &r.bestH, &r.bestF) == NAV24_OK;
    }

    float mThChiSqScore = 5.991f, mThChiSqF = 3.841f;      // DEF_TH_CHISQ_SCORE, DEF_TH_CHISQ_F

private:
    std::shared_ptr<FtDtOrbB200> mpDet;
    float mSigma;
};

}  // namespace OP
}  // namespace NAV24
