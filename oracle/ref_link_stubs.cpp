// ref_link_stubs.cpp — TEST INFRASTRUCTURE (oracle/_ref build only).  One symbol the reference's matcher TU references
// from code that is not on the ORB path: FtAssocOrbSlam::match(f1, f2, tracks) (OP_FtAssocOrbSlam.cpp:225-245) feeds a
// FeatureTracks, whose implementation (sensorData/observation/MatchedFeatures.cpp) needs C++20 std::map::contains
// while tools/parameters/Parameter.hpp:83 does not parse as C++20 with g++ 13 — the two cannot share one -std here.
// The driver never calls the tracks overload; it aborts loudly if anything does.
#include <cstdio>
#include <cstdlib>
#include <string>
#include "MatchedFeatures.hpp"
namespace NAV24::OB {
void FeatureTracks::addMatch(const ObsPtr&, const ObsPtr&) {
    fprintf(stderr, "oracle/_ref: FeatureTracks::addMatch is a link stub (not on the ORB path)\n");
    abort();
}
}  // namespace NAV24::OB
