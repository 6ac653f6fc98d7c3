// ref_shim/Calibration.hpp — TEST INFRASTRUCTURE (oracle/_ref build only).  Shadows core/sensor/camera/Calibration.hpp,
// which FeatureGrid.hpp and MatchedFeatures.hpp include without using: the real header pulls GeometricCamera.h (boost
// serialization, Eigen geometry, cv::Mat_ initialisers), none of which is on the ORB path.  Only the names survive.
#pragma once
#include <memory>
#include <vector>
#include <opencv2/core.hpp>
#include "Parameter.hpp"
#include "WorldObject.hpp"
namespace NAV24 {
class Calibration;
typedef std::shared_ptr<Calibration> CalibPtr;
}
