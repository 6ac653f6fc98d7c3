// ref_shim/dbow2/thirdparty/DBoW2/DUtils/Random.h — TEST INFRASTRUCTURE (oracle/_ref build only).
// Stand-in for DBoW2's DUtils::Random (an un-vendored third-party dependency of the reference: the include at
// core/operators/mapInit/OP_2ViewReconstruction.cpp:22 points outside the tree).  Only the two members the reference
// calls (:109, :118) exist, with the library's published behaviour: SeedRandOnce seeds the C generator once per process,
// RandomInt(min, max) scales rand() into [min, max].
#pragma once
#include <cstdlib>
namespace DUtils {
class Random {
public:
    static void SeedRandOnce(int seed) { static bool seeded = false; if (!seeded) { srand((unsigned)seed); seeded = true; } }
    static int RandomInt(int min, int max) { const int d = max - min + 1; return int(((double)rand() / ((double)RAND_MAX + 1.0)) * d) + min; }
};
}
