// Pins nav24_b200/csrc/stdsort.cuh (the device restatement of libstdc++ std::sort) against the real
// std::sort on tie-heavy, adversarial and random inputs.  Host build: g++ -O2 test_stdsort.cpp.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../nav24_b200/csrc/stdsort.cuh"

struct Rec { unsigned key; unsigned id; };
static bool rec_less(const Rec& a, const Rec& b) { return a.key < b.key; }

static long g_heap_cases = 0;

static bool check(const std::vector<unsigned>& keys) {
    const int n = (int)keys.size();
    std::vector<Rec> a(n);
    std::vector<nav24::stdsort::rec_t> b(n);
    for (int i = 0; i < n; ++i) { a[i] = {keys[i], (unsigned)i}; b[i] = ((unsigned long long)keys[i] << 32) | (unsigned)i; }
    std::sort(a.begin(), a.end(), rec_less);
    nav24::stdsort::sort(b.data(), n);
    for (int i = 0; i < n; ++i)
        if ((unsigned)(b[i] & 0xffffffffu) != a[i].id || (unsigned)(b[i] >> 32) != a[i].key) return false;
    return true;
}

// median-of-3 killer (Musser): forces the depth limit and therefore the heap-sort fallback
static std::vector<unsigned> killer(int n) {
    std::vector<unsigned> v(n);
    int k = n / 2;
    for (int i = 1; i <= k; ++i) {
        if (i % 2) { v[i - 1] = i; v[i] = k + i; }
        v[k + i - 1] = 2 * i;
    }
    return v;
}

int main() {
    std::mt19937 rng(12345);
    long cases = 0;
    for (int n = 0; n <= 70; ++n)
        for (int rep = 0; rep < 40; ++rep) {
            std::vector<unsigned> k(n);
            unsigned range = 1 + rng() % 6;
            for (auto& x : k) x = rng() % range;
            if (!check(k)) { printf("FAIL small n=%d\n", n); return 1; }
            ++cases;
        }
    for (int rep = 0; rep < 600; ++rep) {
        int n = 17 + rng() % 6000;
        std::vector<unsigned> k(n);
        int mode = rep % 6;
        unsigned range = (mode == 0) ? 2 : (mode == 1) ? 8 : (mode == 2) ? 64 : (mode == 3) ? 4096 : 1u << 30;
        for (auto& x : k) x = rng() % range;
        if (mode == 5) std::sort(k.begin(), k.end());
        if (rep % 11 == 0) std::reverse(k.begin(), k.end());
        if (!check(k)) { printf("FAIL random n=%d mode=%d\n", n, mode); return 1; }
        ++cases;
    }
    // quadtree-like keys: (count << 12 | ulx) with few distinct counts
    for (int rep = 0; rep < 300; ++rep) {
        int n = 20 + rng() % 3000;
        std::vector<unsigned> k(n);
        for (auto& x : k) x = ((2 + rng() % 5) << 12) | ((rng() % 40) * 31);
        if (!check(k)) { printf("FAIL quadtree-like n=%d\n", n); return 1; }
        ++cases;
    }
    for (int n : {64, 100, 1000, 4096, 10000, 65536}) {
        if (!check(killer(n))) { printf("FAIL killer n=%d\n", n); return 1; }
        std::vector<unsigned> organ(n);
        for (int i = 0; i < n; ++i) organ[i] = std::min(i, n - 1 - i);
        if (!check(organ)) { printf("FAIL organ n=%d\n", n); return 1; }
        std::vector<unsigned> same(n, 7u);
        if (!check(same)) { printf("FAIL const n=%d\n", n); return 1; }
        cases += 3;
    }
    printf("OK %ld cases\n", cases);
    return 0;
}
