// ref_access_2v.hpp — TEST INFRASTRUCTURE.  Access to the private members of the reference's
// NAV24::OP::TwoViewReconstruction (core/operators/mapInit/OP_2ViewReconstruction.hpp:150-186) without touching its
// header: an explicit template instantiation may name private members, and a friend function defined in the
// instantiated class hands the member pointer out.  Everything sits in an unnamed namespace, so every translation unit
// that includes this file (oracle/ref_driver_2v.cpp, tests/cpp/test_ref_binding.cpp) gets its own copies.
// Usage: (obj.*get(CheckH()))(H21, H12, inliers, sigma);   obj.*get(Matches())
#pragma once
#include <utility>
#include <vector>

#include "OP_2ViewReconstruction.hpp"

namespace {

using NAV24::OP::TwoViewReconstruction;

template <class Tag, typename Tag::type M> struct Rob { friend typename Tag::type get(Tag) { return M; } };
#define NAV24_ROB(tag, ...) struct tag { typedef __VA_ARGS__; friend type get(tag); }

typedef std::vector<bool> VB;
typedef std::vector<cv::KeyPoint> VK;
typedef std::vector<cv::Point2f> VP;
typedef std::vector<std::pair<int, int>> VM;
typedef std::vector<std::vector<size_t>> VS;

NAV24_ROB(CheckH, float (TwoViewReconstruction::*type)(const cv::Mat&, const cv::Mat&, VB&, float));
NAV24_ROB(CheckF, float (TwoViewReconstruction::*type)(const cv::Mat&, VB&, float));
NAV24_ROB(FindH, void (TwoViewReconstruction::*type)(VB&, float&, cv::Mat&));
NAV24_ROB(FindF, void (TwoViewReconstruction::*type)(VB&, float&, cv::Mat&));
NAV24_ROB(CompH, cv::Mat (TwoViewReconstruction::*type)(const VP&, const VP&));
NAV24_ROB(CompF, cv::Mat (TwoViewReconstruction::*type)(const VP&, const VP&));
NAV24_ROB(Norm, void (TwoViewReconstruction::*type)(const VK&, VP&, cv::Mat&));
NAV24_ROB(Keys1, VK TwoViewReconstruction::*type);
NAV24_ROB(Keys2, VK TwoViewReconstruction::*type);
NAV24_ROB(Matches, VM TwoViewReconstruction::*type);
NAV24_ROB(Sets, VS TwoViewReconstruction::*type);
NAV24_ROB(MaxIt, int TwoViewReconstruction::*type);
NAV24_ROB(Sigma, float TwoViewReconstruction::*type);
#undef NAV24_ROB

template struct Rob<CheckH, &TwoViewReconstruction::CheckHomography>;
template struct Rob<CheckF, &TwoViewReconstruction::CheckFundamental>;
template struct Rob<FindH, &TwoViewReconstruction::FindHomography>;
template struct Rob<FindF, &TwoViewReconstruction::FindFundamental>;
template struct Rob<CompH, &TwoViewReconstruction::ComputeH21>;
template struct Rob<CompF, &TwoViewReconstruction::ComputeF21>;
template struct Rob<Norm, &TwoViewReconstruction::Normalize>;
template struct Rob<Keys1, &TwoViewReconstruction::mvKeys1>;
template struct Rob<Keys2, &TwoViewReconstruction::mvKeys2>;
template struct Rob<Matches, &TwoViewReconstruction::mvMatches12>;
template struct Rob<Sets, &TwoViewReconstruction::mvSets>;
template struct Rob<MaxIt, &TwoViewReconstruction::mMaxIterations>;
template struct Rob<Sigma, &TwoViewReconstruction::mSigma>;

}  // namespace
