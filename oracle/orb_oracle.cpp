// orb_oracle.cpp — CPU ORACLE (test infrastructure, NOT a product path).
//
// A dependency-free scalar C++ restatement of the ORB front-end hot path of m-dayani/nav24
// (reference commit a6e8294), used ONLY by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs as the checker and the timed CPU baseline.  Nothing in
// nav24_b200/ may link or call this file.
//
// Parity status: PINNED.  The reference holds no golden vectors / known-answer tests for this path
//  and its own build system cannot run here (OpenCV C++, Eigen, glog, g2o are
// absent), so the oracle is pinned against
//   (i)   live cv2 4.13.0 for every OpenCV primitive (resize / FAST / GaussianBlur / fastAtan2 /
//         BFMatcher / cvtColor / undistortPoints; tests/test_oracle_vs_cv2.py, frozen in tests/golden/),
//   (ii)  the reference's OWN translation units compiled unchanged from /root/reference into
//         oracle/_ref (oracle/Makefile.ref, container stand-ins in oracle/ref_shim): extractor,
//         quadtree, matcher, grid, frame classes (tests/test_ref_build.py) and the two-view scoring
//         (TwoViewReconstruction::CheckHomography / CheckFundamental / FindHomography /
//         FindFundamental, tests/test_two_view.py) — oracle == reference build, bit for bit.
//
// Build: g++ -O3 -march=x86-64-v3 -ffp-contract=off -shared -fPIC (see oracle/Makefile).
// -ffp-contract=off matters: the reference is built without FMA contraction.
//
// Reference files restated (paths relative to /root/reference/core):
//   operators/objDetection/OP_FtDtOrbSlam.cpp   (detector; line ranges cited per function)
//   operators/objAssoc/OP_FtAssocOrbSlam.cpp    (windowed matcher)
//   operators/objAssoc/OP_FtAssoc.cpp           (brute-force kNN-2 + ratio)
//   sensorData/observation/FeatureGrid.cpp      (10-px grid, candidate order)
//   operators/mapInit/OP_2ViewReconstruction.cpp (CheckHomography / CheckFundamental)
// OpenCV primitives follow SURVEY.md Appendix A (verified bit-exact vs cv2 4.13.0 by tests).

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <utility>
#include <vector>

namespace {

const int kPatch = 31, kHalfPatch = 15, kEdge = 19;

const int kPattern[1024] = {
#include "pattern_31.inc"
};

struct Kp {
    float x, y, size, angle, response;
    int octave, class_id;
};

struct RawKey {  // one FAST survivor: coordinates relative to (minBorderX, minBorderY)
    float x, y, response;
};

inline int round_half_even(float v) { return (int)lrintf(v); }       // cvRound(float)
inline int round_half_even(double v) { return (int)lrint(v); }       // cvRound(double)

// ------------------------------------------------------------------------------------------
// OpenCV primitives
// ------------------------------------------------------------------------------------------

// A.1  cv::resize(..., INTER_LINEAR) for CV_8UC1: 11-bit fixed-point separable bilinear.
void resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep,
                      uint8_t* dst, int dw, int dh, size_t dstep) {
    const double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
    const double scale_x = 1.0 / inv_x, scale_y = 1.0 / inv_y;
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> xa(2 * dw), ya(2 * dh);
    for (int d = 0; d < dw; ++d) {
        float f = (float)((d + 0.5) * scale_x - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= sw - 1) { s = sw - 1; f = 0.f; }
        xofs[d] = s;
        xa[2 * d] = (short)round_half_even((1.f - f) * 2048.f);
        xa[2 * d + 1] = (short)round_half_even(f * 2048.f);
    }
    for (int d = 0; d < dh; ++d) {
        float f = (float)((d + 0.5) * scale_y - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= sh - 1) { s = sh - 1; f = 0.f; }
        yofs[d] = s;
        ya[2 * d] = (short)round_half_even((1.f - f) * 2048.f);
        ya[2 * d + 1] = (short)round_half_even(f * 2048.f);
    }
    std::vector<int> row0(dw), row1(dw);
    int cached0 = -1, cached1 = -1;
    auto hpass = [&](int sy, std::vector<int>& out) {
        const uint8_t* S = src + (size_t)sy * sstep;
        for (int d = 0; d < dw; ++d) {
            int s0 = xofs[d], s1 = std::min(s0 + 1, sw - 1);
            out[d] = S[s0] * xa[2 * d] + S[s1] * xa[2 * d + 1];
        }
    };
    for (int dy = 0; dy < dh; ++dy) {
        int sy0 = yofs[dy], sy1 = std::min(sy0 + 1, sh - 1);
        if (cached1 == sy0) { row0.swap(row1); std::swap(cached0, cached1); }
        if (cached0 != sy0) { hpass(sy0, row0); cached0 = sy0; }
        if (cached1 != sy1) {
            if (sy1 == sy0) { row1 = row0; } else { hpass(sy1, row1); }
            cached1 = sy1;
        }
        const int b0 = ya[2 * dy], b1 = ya[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * dstep;
        for (int d = 0; d < dw; ++d) {
            int v = (((b0 * (row0[d] >> 4)) >> 16) + ((b1 * (row1[d] >> 4)) >> 16) + 2) >> 2;
            D[d] = (uint8_t)v;
        }
    }
}

inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p; else p = 2 * n - 2 - p;
    }
    return p;
}

// A.2  cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) for CV_8UC1: exact integer kernel.
void gauss7_u8(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep) {
    static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
    std::vector<int> hbuf((size_t)w * h);
    std::vector<int> xi(w + 6);
    for (int x = -3; x < w + 3; ++x) xi[x + 3] = reflect101(x, w);
    for (int y = 0; y < h; ++y) {
        const uint8_t* S = src + (size_t)y * sstep;
        int* H = &hbuf[(size_t)y * w];
        for (int x = 0; x < w; ++x) {
            int a = 0;
            for (int i = 0; i < 7; ++i) a += k[i] * S[xi[x + i]];
            H[x] = a;
        }
    }
    for (int y = 0; y < h; ++y) {
        const int* R[7];
        for (int j = 0; j < 7; ++j) R[j] = &hbuf[(size_t)reflect101(y + j - 3, h) * w];
        uint8_t* D = dst + (size_t)y * dstep;
        for (int x = 0; x < w; ++x) {
            int v = 0;
            for (int j = 0; j < 7; ++j) v += k[j] * R[j][x];
            D[x] = (uint8_t)((v + 32768) >> 16);
        }
    }
}

// A.3  FAST-9/16 corner score: max over the 16 arcs of 9 contiguous ring pixels of
// max(min d, min -d), minus 1.  Independent of the detection threshold.
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

inline int fast_score_at(const uint8_t* p, const int* ofs) {
    int d[25];
    const int c = p[0];
    for (int k = 0; k < 16; ++k) d[k] = p[ofs[k]] - c;
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = -256;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
        best = std::max(best, std::max(mn, -mx));
    }
    return best - 1;
}

// cheap necessary condition for score >= t: an arc of 9 contains one pixel of every antipodal pair
inline bool fast_maybe(const uint8_t* p, const int* ofs, int t) {
    const int c = p[0];
    for (int k = 0; k < 8; ++k) {
        int a = p[ofs[k]] - c, b = p[ofs[k + 8]] - c;
        if (std::abs(a) <= t && std::abs(b) <= t) return false;
    }
    return true;
}

// cv::FAST(img, kps, t, nonmaxSuppression=true) on a sub-image; returns keypoints row-major.
// Coordinates are sub-image relative.  scratch must hold w*h ints.
void fast_nms(const uint8_t* img, int w, int h, size_t step, int t, std::vector<RawKey>& out,
              std::vector<int>& score) {
    out.clear();
    if (w < 7 || h < 7) return;
    int ofs[16];
    for (int k = 0; k < 16; ++k) ofs[k] = kRingDy[k] * (int)step + kRingDx[k];
    score.assign((size_t)w * h, 0);
    for (int y = 3; y < h - 3; ++y) {
        const uint8_t* row = img + (size_t)y * step;
        for (int x = 3; x < w - 3; ++x) {
            if (!fast_maybe(row + x, ofs, t)) continue;
            int s = fast_score_at(row + x, ofs);
            if (s >= t) score[(size_t)y * w + x] = s;
        }
    }
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const int s = score[(size_t)y * w + x];
            if (s == 0) continue;   // t >= 1 for every caller, so 0 means "not a corner"
            const int* r = &score[(size_t)y * w + x];
            if (s > r[-1] && s > r[1] && s > r[-w - 1] && s > r[-w] && s > r[-w + 1] &&
                s > r[w - 1] && s > r[w] && s > r[w + 1])
                out.push_back({(float)x, (float)y, (float)s});
        }
}

// A.4  cv::fastAtan2(y, x), degrees, scalar f32 arithmetic without contraction.
float fast_atan2_deg(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float eps = (float)2.2204460492503131e-16;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ------------------------------------------------------------------------------------------
// Quadtree distribution (OP_FtDtOrbSlam.cpp:358-373, :384-439, :502-725)
// ------------------------------------------------------------------------------------------
struct QNode {
    int x0, y0, x1, y1;            // UL.x, UL.y, UR.x (=BR.x), BL.y (=BR.y)
    std::vector<int> keys;         // indices into the level's raw key array, input order kept
    std::list<QNode>::iterator self;
    bool leaf = false;             // "bNoMore"
};

struct SizedNode {
    int count;
    QNode* node;
};

// compareNodes (:358-373): orders by (count, UL.x) only; everything else is "equivalent".
bool sized_less(const SizedNode& a, const SizedNode& b) {
    if (a.count < b.count) return true;
    if (a.count > b.count) return false;
    return a.node->x0 < b.node->x0;
}

// ExtractorNode::DivideNode (:384-439)
void split_node(const QNode& n, const std::vector<RawKey>& keys, QNode c[4]) {
    const int halfX = (int)std::ceil((float)(n.x1 - n.x0) / 2);
    const int halfY = (int)std::ceil((float)(n.y1 - n.y0) / 2);
    const int mx = n.x0 + halfX, my = n.y0 + halfY;
    c[0].x0 = n.x0; c[0].y0 = n.y0; c[0].x1 = mx;   c[0].y1 = my;
    c[1].x0 = mx;   c[1].y0 = n.y0; c[1].x1 = n.x1; c[1].y1 = my;
    c[2].x0 = n.x0; c[2].y0 = my;   c[2].x1 = mx;   c[2].y1 = n.y1;
    c[3].x0 = mx;   c[3].y0 = my;   c[3].x1 = n.x1; c[3].y1 = n.y1;
    for (int idx : n.keys) {
        const RawKey& k = keys[idx];
        int q = (k.x < (float)mx) ? ((k.y < (float)my) ? 0 : 2) : ((k.y < (float)my) ? 1 : 3);
        c[q].keys.push_back(idx);
    }
    for (int q = 0; q < 4; ++q) c[q].leaf = (c[q].keys.size() == 1);
}

// returns indices (into keys) of the retained keypoints, in the reference's output order.
// Returns false if the geometry would make the reference divide by zero / index out of range.
bool distribute_quadtree(const std::vector<RawKey>& keys, int minX, int maxX, int minY, int maxY,
                         int N, std::vector<int>& kept) {
    kept.clear();
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
    if (nIni < 1) return false;
    const float hX = (float)(maxX - minX) / (float)nIni;

    std::list<QNode> nodes;
    std::vector<QNode*> roots(nIni);
    for (int i = 0; i < nIni; ++i) {
        QNode n;
        n.x0 = (int)(hX * (float)i);
        n.x1 = (int)(hX * (float)(i + 1));
        n.y0 = 0;
        n.y1 = maxY - minY;
        nodes.push_back(n);
        roots[i] = &nodes.back();
    }
    for (int i = 0; i < (int)keys.size(); ++i) {
        size_t r = (size_t)(keys[i].x / hX);
        if (r >= roots.size()) return false;
        roots[r]->keys.push_back(i);
    }
    for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->keys.size() == 1) { it->leaf = true; ++it; }
        else if (it->keys.empty()) it = nodes.erase(it);
        else ++it;
    }

    std::vector<SizedNode> expandable;
    auto push_children = [&](QNode c[4], int& nExpand) {
        for (int q = 0; q < 4; ++q) {
            if (c[q].keys.empty()) continue;
            nodes.push_front(c[q]);
            if (c[q].keys.size() > 1) {
                ++nExpand;
                expandable.push_back({(int)c[q].keys.size(), &nodes.front()});
                nodes.front().self = nodes.begin();
            }
        }
    };

    bool done = false;
    while (!done) {
        int prevSize = (int)nodes.size();
        int nExpand = 0;
        expandable.clear();
        for (auto it = nodes.begin(); it != nodes.end();) {
            if (it->leaf) { ++it; continue; }
            QNode c[4];
            split_node(*it, keys, c);
            push_children(c, nExpand);
            it = nodes.erase(it);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) {
            done = true;
        } else if ((int)nodes.size() + nExpand * 3 > N) {
            while (!done) {
                prevSize = (int)nodes.size();
                std::vector<SizedNode> prev = expandable;
                expandable.clear();
                std::sort(prev.begin(), prev.end(), sized_less);
                for (int j = (int)prev.size() - 1; j >= 0; --j) {
                    QNode c[4];
                    split_node(*prev[j].node, keys, c);
                    int dummy = 0;
                    push_children(c, dummy);
                    nodes.erase(prev[j].node->self);
                    if ((int)nodes.size() >= N) break;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) done = true;
            }
        }
    }
    kept.reserve(nodes.size());
    for (const QNode& n : nodes) {
        int best = n.keys[0];
        float bestR = keys[best].response;
        for (size_t k = 1; k < n.keys.size(); ++k)
            if (keys[n.keys[k]].response > bestR) { best = n.keys[k]; bestR = keys[best].response; }
        kept.push_back(best);
    }
    return true;
}

// ------------------------------------------------------------------------------------------
// Detector (OP_FtDtOrbSlam.cpp:441-500 ctor, :727-842 keypoints, :844-934 detect, :936-960 pyramid)
// ------------------------------------------------------------------------------------------
struct Level {
    int w = 0, h = 0;
    std::vector<uint8_t> px, blurred;
    std::vector<RawKey> raw;       // vToDistributeKeys
    std::vector<Kp> kps;           // after quadtree + border + angle (level coordinates)
    std::vector<uint8_t> desc;
    bool has_blur = false;
};

struct Oracle {
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;
    std::vector<float> scale, invScale;
    std::vector<int> quota;
    int umax[kHalfPatch + 1];
    std::vector<Level> lv;
    std::vector<int> scratch;

    void set_quota(int n) {
        nfeatures = n;
        float factor = 1.0f / scaleFactor;
        float per = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
        int sum = 0;
        for (int l = 0; l < nlevels - 1; ++l) {
            quota[l] = round_half_even(per);
            sum += quota[l];
            per *= factor;
        }
        quota[nlevels - 1] = std::max(nfeatures - sum, 0);
    }

    Oracle(int nf, float sf, int nl, int ini, int mn) : nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(sf) {
        scale.resize(nl); invScale.resize(nl); quota.resize(nl); lv.resize(nl);
        scale[0] = 1.0f;
        for (int i = 1; i < nl; ++i) scale[i] = scale[i - 1] * scaleFactor;
        for (int i = 0; i < nl; ++i) invScale[i] = 1.0f / scale[i];
        set_quota(nf);
        int v, v0, vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
        int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
        const double hp2 = kHalfPatch * kHalfPatch;
        for (v = 0; v <= vmax; ++v) umax[v] = round_half_even(std::sqrt(hp2 - v * v));
        for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }

    // :936-960.  The 19-px reflected border of the reference is never read downstream.
    void pyramid(const uint8_t* img, int w, int h, size_t step) {
        for (int l = 0; l < nlevels; ++l) {
            Level& L = lv[l];
            L.w = round_half_even((float)w * invScale[l]);
            L.h = round_half_even((float)h * invScale[l]);
            L.px.resize((size_t)L.w * L.h);
            L.has_blur = false;
            if (l == 0) {
                for (int y = 0; y < h; ++y) memcpy(&L.px[(size_t)y * w], img + (size_t)y * step, w);
            } else {
                Level& P = lv[l - 1];
                resize_linear_u8(P.px.data(), P.w, P.h, P.w, L.px.data(), L.w, L.h, L.w);
            }
        }
    }

    // IC_Angle (:18-44)
    float ic_angle(const Level& L, float px, float py) const {
        int m01 = 0, m10 = 0;
        const int cx = round_half_even(px), cy = round_half_even(py);
        const uint8_t* c = &L.px[(size_t)cy * L.w + cx];
        for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
        for (int v = 1; v <= kHalfPatch; ++v) {
            int vsum = 0, d = umax[v];
            for (int u = -d; u <= d; ++u) {
                int p = c[u + v * L.w], m = c[u - v * L.w];
                vsum += p - m;
                m10 += u * (p + m);
            }
            m01 += v * vsum;
        }
        return fast_atan2_deg((float)m01, (float)m10);
    }

    // :727-842; returns false on degenerate geometry (image too small for one 35-px cell)
    bool keypoints() {
        const float W = 35;
        std::vector<RawKey> cell;
        std::vector<int> kept;
        for (int l = 0; l < nlevels; ++l) {
            Level& L = lv[l];
            const int minBX = kEdge - 3, minBY = minBX;
            const int maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
            const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
            const int nCols = (int)(width / W), nRows = (int)(height / W);
            if (nCols < 1 || nRows < 1) return false;
            const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
            L.raw.clear();
            for (int i = 0; i < nRows; ++i) {
                const float iniY = (float)(minBY + i * hCell);
                float maxY = iniY + hCell + 6;
                if (iniY >= maxBY - 3) continue;
                if (maxY > maxBY) maxY = (float)maxBY;
                for (int j = 0; j < nCols; ++j) {
                    const float iniX = (float)(minBX + j * wCell);
                    float maxX = iniX + wCell + 6;
                    if (iniX >= maxBX - 6) continue;
                    if (maxX > maxBX) maxX = (float)maxBX;
                    const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
                    const uint8_t* sub = &L.px[(size_t)y0 * L.w + x0];
                    fast_nms(sub, cw, ch, L.w, iniTh, cell, scratch);
                    if (cell.empty()) fast_nms(sub, cw, ch, L.w, minTh, cell, scratch);
                    for (RawKey k : cell) {
                        k.x += (float)(j * wCell);
                        k.y += (float)(i * hCell);
                        L.raw.push_back(k);
                    }
                }
            }
            if (!distribute_quadtree(L.raw, minBX, maxBX, minBY, maxBY, quota[l], kept)) return false;
            const int patch = (int)(kPatch * scale[l]);
            L.kps.clear();
            for (int idx : kept) {
                const RawKey& r = L.raw[idx];
                Kp k;
                k.x = r.x + (float)minBX; k.y = r.y + (float)minBY;
                k.size = (float)patch; k.angle = -1.f; k.response = r.response;
                k.octave = l; k.class_id = -1;
                L.kps.push_back(k);
            }
        }
        for (int l = 0; l < nlevels; ++l)
            for (Kp& k : lv[l].kps) k.angle = ic_angle(lv[l], k.x, k.y);
        return true;
    }

    // computeOrbDescriptor (:48-87)
    static void describe(const Kp& k, const uint8_t* img, int step, uint8_t* out) {
        const float factorPI = (float)(3.14159265358979323846 / 180.f);
        const float ang = k.angle * factorPI;
        const float a = cosf(ang), b = sinf(ang);
        const uint8_t* c = img + (size_t)round_half_even(k.y) * step + round_half_even(k.x);
        const int* p = kPattern;
        for (int i = 0; i < 32; ++i, p += 32) {
            int val = 0;
            for (int j = 0; j < 8; ++j) {
                const int x0 = p[4 * j], y0 = p[4 * j + 1], x1 = p[4 * j + 2], y1 = p[4 * j + 3];
                int t0 = c[round_half_even(x0 * b + y0 * a) * step + round_half_even(x0 * a - y0 * b)];
                int t1 = c[round_half_even(x1 * b + y1 * a) * step + round_half_even(x1 * a - y1 * b)];
                val |= (t0 < t1) << j;
            }
            out[i] = (uint8_t)val;
        }
    }

    // :844-934.  Returns monoIndex, -1 on empty image, -2 on degenerate geometry.
    int detect(const uint8_t* img, int w, int h, size_t step, std::vector<Kp>& outK, std::vector<uint8_t>& outD) {
        if (!img || w <= 0 || h <= 0) return -1;
        pyramid(img, w, h, step);
        if (!keypoints()) return -2;
        int n = 0;
        for (int l = 0; l < nlevels; ++l) n += (int)lv[l].kps.size();
        outK.assign(n, Kp());
        outD.assign((size_t)n * 32, 0);
        int mono = 0, stereo = n - 1;
        for (int l = 0; l < nlevels; ++l) {
            Level& L = lv[l];
            if (L.kps.empty()) continue;
            L.blurred.resize(L.px.size());
            gauss7_u8(L.px.data(), L.w, L.h, L.w, L.blurred.data(), L.w);
            L.has_blur = true;
            L.desc.resize(L.kps.size() * 32);
            for (size_t i = 0; i < L.kps.size(); ++i) describe(L.kps[i], L.blurred.data(), L.w, &L.desc[i * 32]);
            const float s = scale[l];
            for (size_t i = 0; i < L.kps.size(); ++i) {
                Kp k = L.kps[i];
                if (l != 0) { k.x *= s; k.y *= s; }
                int dst = (k.x >= 0 && k.x <= 1000) ? stereo-- : mono++;
                outK[dst] = k;
                memcpy(&outD[(size_t)dst * 32], &L.desc[i * 32], 32);
            }
        }
        return mono;
    }
};

// ------------------------------------------------------------------------------------------
// FeatureGrid (FeatureGrid.cpp:20-152) and windowed matcher (OP_FtAssocOrbSlam.cpp:28-223)
// ------------------------------------------------------------------------------------------
struct GridCfg {
    int cols, rows;
    float minX, maxX, minY, maxY;
};

struct Grid {
    GridCfg g;
    float invW, invH;
    std::vector<std::vector<int>> cells;   // [ix*rows + iy]
    Grid(const GridCfg& cfg, const float* ud, int n) : g(cfg) {
        invW = (float)g.cols / (g.maxX - g.minX);
        invH = (float)g.rows / (g.maxY - g.minY);
        cells.resize((size_t)g.cols * g.rows);
        for (int i = 0; i < n; ++i) {
            int px = (int)std::round((ud[2 * i] - g.minX) * invW);
            int py = (int)std::round((ud[2 * i + 1] - g.minY) * invH);
            if (px < 0 || px >= g.cols || py < 0 || py >= g.rows) continue;
            cells[(size_t)px * g.rows + py].push_back(i);
        }
    }
    void query(float x, float y, float r, int minLevel, int maxLevel, const float* ud, const Kp* kps,
               std::vector<int>& out) const {
        out.clear();
        const int cx0 = std::max(0, (int)std::floor((x - g.minX - r) * invW));
        if (cx0 >= g.cols) return;
        const int cx1 = std::min(g.cols - 1, (int)std::ceil((x - g.minX + r) * invW));
        if (cx1 < 0) return;
        const int cy0 = std::max(0, (int)std::floor((y - g.minY - r) * invH));
        if (cy0 >= g.rows) return;
        const int cy1 = std::min(g.rows - 1, (int)std::ceil((y - g.minY + r) * invH));
        if (cy1 < 0) return;
        const bool check = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = cx0; ix <= cx1; ++ix)
            for (int iy = cy0; iy <= cy1; ++iy)
                for (int j : cells[(size_t)ix * g.rows + iy]) {
                    if (check) {
                        if (kps[j].octave < minLevel) continue;
                        if (maxLevel >= 0 && kps[j].octave > maxLevel) continue;
                    }
                    const float dx = ud[2 * j] - x, dy = ud[2 * j + 1] - y;
                    if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(j);
                }
    }
};

inline int hamming256(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 32; i += 8) {
        uint64_t x, y;
        memcpy(&x, a + i, 8); memcpy(&y, b + i, 8);
        d += __builtin_popcountll(x ^ y);
    }
    return d;
}

// ComputeThreeMaxima (:28-69)
void three_maxima(const int* cnt, int L, int& i1, int& i2, int& i3) {
    int m1 = 0, m2 = 0, m3 = 0;
    for (int i = 0; i < L; ++i) {
        const int s = cnt[i];
        if (s > m1) { m3 = m2; m2 = m1; m1 = s; i3 = i2; i2 = i1; i1 = i; }
        else if (s > m2) { m3 = m2; m2 = s; i3 = i2; i2 = i; }
        else if (s > m3) { m3 = s; i3 = i; }
    }
    if (m2 < 0.1f * (float)m1) { i2 = -1; i3 = -1; }
    else if (m3 < 0.1f * (float)m1) { i3 = -1; }
}

}  // namespace

// ==========================================================================================
// C interface (ctypes) — oracle only
// ==========================================================================================
extern "C" {

struct orc_kp { float x, y, size, angle, response; int octave, class_id; };
struct orc_grid { int cols, rows; float minX, maxX, minY, maxY; };

void* orc_create(int nfeatures, float scale, int nlevels, int iniTh, int minTh) {
    return new Oracle(nfeatures, scale, nlevels, iniTh, minTh);
}
void orc_destroy(void* h) { delete (Oracle*)h; }
void orc_set_num_features(void* h, int n) { ((Oracle*)h)->set_quota(n); }
void orc_get_tables(void* h, float* scale, float* inv, int* quota, int* umax) {
    Oracle* o = (Oracle*)h;
    for (int l = 0; l < o->nlevels; ++l) { scale[l] = o->scale[l]; inv[l] = o->invScale[l]; quota[l] = o->quota[l]; }
    for (int i = 0; i <= kHalfPatch; ++i) umax[i] = o->umax[i];
}

int orc_detect(void* h, const uint8_t* img, int w, int hh, size_t stride, orc_kp* kps, uint8_t* desc, int cap,
               int* n_out) {
    Oracle* o = (Oracle*)h;
    std::vector<Kp> K;
    std::vector<uint8_t> D;
    int mono = o->detect(img, w, hh, stride, K, D);
    if (mono < 0) { *n_out = 0; return mono; }
    *n_out = (int)K.size();
    if ((int)K.size() > cap) return -3;
    if (!K.empty()) { memcpy(kps, K.data(), K.size() * sizeof(Kp)); memcpy(desc, D.data(), D.size()); }
    return mono;
}

// detect without copying results (timing loops)
int orc_detect_count(void* h, const uint8_t* img, int w, int hh, size_t stride) {
    Oracle* o = (Oracle*)h;
    std::vector<Kp> K;
    std::vector<uint8_t> D;
    int mono = o->detect(img, w, hh, stride, K, D);
    return mono < 0 ? mono : (int)K.size();
}

void orc_level_size(void* h, int l, int* w, int* hh) { Oracle* o = (Oracle*)h; *w = o->lv[l].w; *hh = o->lv[l].h; }
void orc_get_level(void* h, int l, uint8_t* dst) { Oracle* o = (Oracle*)h; memcpy(dst, o->lv[l].px.data(), o->lv[l].px.size()); }
int orc_get_blurred(void* h, int l, uint8_t* dst) {
    Oracle* o = (Oracle*)h;
    if (!o->lv[l].has_blur) return 0;
    memcpy(dst, o->lv[l].blurred.data(), o->lv[l].blurred.size());
    return 1;
}
int orc_raw_count(void* h, int l) { return (int)((Oracle*)h)->lv[l].raw.size(); }
void orc_get_raw(void* h, int l, float* xyr) {
    Oracle* o = (Oracle*)h;
    for (size_t i = 0; i < o->lv[l].raw.size(); ++i) {
        xyr[3 * i] = o->lv[l].raw[i].x; xyr[3 * i + 1] = o->lv[l].raw[i].y; xyr[3 * i + 2] = o->lv[l].raw[i].response;
    }
}
int orc_level_kp_count(void* h, int l) { return (int)((Oracle*)h)->lv[l].kps.size(); }
void orc_get_level_kps(void* h, int l, orc_kp* out, uint8_t* desc) {
    Oracle* o = (Oracle*)h;
    if (o->lv[l].kps.empty()) return;
    memcpy(out, o->lv[l].kps.data(), o->lv[l].kps.size() * sizeof(Kp));
    if (desc && !o->lv[l].desc.empty()) memcpy(desc, o->lv[l].desc.data(), o->lv[l].desc.size());
}

// --- primitives, exposed for pinning against cv2 ------------------------------------------
void orc_resize_u8(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst, int dw, int dh, size_t dstep) {
    resize_linear_u8(src, sw, sh, sstep, dst, dw, dh, dstep);
}
void orc_gauss7_u8(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep) {
    gauss7_u8(src, w, h, sstep, dst, dstep);
}
int orc_fast_u8(const uint8_t* img, int w, int h, size_t step, int t, float* xyr, int cap) {
    std::vector<RawKey> out;
    std::vector<int> scratch;
    fast_nms(img, w, h, step, t, out, scratch);
    int n = std::min((int)out.size(), cap);
    for (int i = 0; i < n; ++i) { xyr[3 * i] = out[i].x; xyr[3 * i + 1] = out[i].y; xyr[3 * i + 2] = out[i].response; }
    return (int)out.size();
}
float orc_fast_atan2(float y, float x) { return fast_atan2_deg(y, x); }
// cv::cvtColor(bgr, grey, COLOR_BGR2GRAY) on CV_8UC3 (FE_SlamMonoV.cpp:92-94): OpenCV's 15-bit fixed point
// (B 3735, G 19235, R 9798, rounding shift), pinned against live cv2 4.13 in tests/test_ingest.py.
void orc_bgr2gray(const uint8_t* bgr, int w, int h, size_t sstep, uint8_t* dst, size_t dstep) {
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = bgr + (size_t)y * sstep;
        uint8_t* d = dst + (size_t)y * dstep;
        for (int x = 0; x < w; ++x) d[x] = (uint8_t)((s[3 * x] * 3735 + s[3 * x + 1] * 19235 + s[3 * x + 2] * 9798 + 16384) >> 15);
    }
}
int orc_quadtree(const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N, int* kept, int cap) {
    std::vector<RawKey> keys(n);
    for (int i = 0; i < n; ++i) keys[i] = {xyr[3 * i], xyr[3 * i + 1], xyr[3 * i + 2]};
    std::vector<int> k;
    if (!distribute_quadtree(keys, minX, maxX, minY, maxY, N, k)) return -1;
    for (int i = 0; i < (int)k.size() && i < cap; ++i) kept[i] = k[i];
    return (int)k.size();
}
// std::sort with the quadtree's comparator on (count, ULx) pairs; returns the permutation.
// Exposed so the device restatement of libstdc++ introsort can be pinned against the real thing.
void orc_sort_sized(const int* count, const int* ulx, int n, int* perm) {
    std::vector<QNode> nodes(n);
    std::vector<SizedNode> v(n);
    for (int i = 0; i < n; ++i) { nodes[i].x0 = ulx[i]; v[i] = {count[i], &nodes[i]}; }
    std::sort(v.begin(), v.end(), sized_less);
    for (int i = 0; i < n; ++i) perm[i] = (int)(v[i].node - nodes.data());
}

// FtAssocOrbSlam::matchV (OP_FtAssocOrbSlam.cpp:91-223).  ud = undistorted (x,y) pairs.
int orc_match_window(const orc_kp* k1, const float* ud1, const uint8_t* d1, int n1,
                     const orc_kp* k2, const float* ud2, const uint8_t* d2, int n2,
                     const orc_grid* gc, float window, float nnratio, int th_low, int check_ori,
                     int* matches12) {
    const int HISTO = 30;
    GridCfg cfg{gc->cols, gc->rows, gc->minX, gc->maxX, gc->minY, gc->maxY};
    Grid grid(cfg, ud2, n2);
    const Kp* K1 = (const Kp*)k1; const Kp* K2 = (const Kp*)k2;
    std::vector<int> dist2(n2, INT_MAX), m21(n2, -1);
    std::vector<std::vector<int>> hist(HISTO);
    const float factor = 1.0f / HISTO;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    int nmatches = 0;
    std::vector<int> cand;
    for (int i1 = 0; i1 < n1; ++i1) {
        const int level1 = K1[i1].octave;
        if (level1 > 0) continue;
        grid.query(ud1[2 * i1], ud1[2 * i1 + 1], window, level1, level1, ud2, K2, cand);
        if (cand.empty()) continue;
        int best = INT_MAX, best2 = INT_MAX, bestIdx = -1;
        for (int i2 : cand) {
            const int d = hamming256(d1 + (size_t)i1 * 32, d2 + (size_t)i2 * 32);
            if (dist2[i2] <= d) continue;
            if (d < best) { best2 = best; best = d; bestIdx = i2; }
            else if (d < best2) best2 = d;
        }
        if (best <= th_low && best < (float)best2 * nnratio) {
            if (m21[bestIdx] >= 0) { matches12[m21[bestIdx]] = -1; --nmatches; }
            matches12[i1] = bestIdx;
            m21[bestIdx] = i1;
            dist2[bestIdx] = best;
            ++nmatches;
            if (check_ori) {
                float rot = K1[i1].angle - K2[bestIdx].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO) bin = 0;
                if (bin >= 0 && bin < HISTO) hist[bin].push_back(i1);
            }
        }
    }
    if (check_ori) {
        int cnt[HISTO], a = -1, b = -1, c = -1;
        for (int i = 0; i < HISTO; ++i) cnt[i] = (int)hist[i].size();
        three_maxima(cnt, HISTO, a, b, c);
        for (int i = 0; i < HISTO; ++i) {
            if (i == a || i == b || i == c) continue;
            for (int idx1 : hist[i])
                if (matches12[idx1] >= 0) { matches12[idx1] = -1; --nmatches; }
        }
    }
    return nmatches;
}

int orc_grid_query(const orc_kp* k2, const float* ud2, int n2, const orc_grid* gc, float x, float y, float r,
                   int minLevel, int maxLevel, int* out, int cap) {
    GridCfg cfg{gc->cols, gc->rows, gc->minX, gc->maxX, gc->minY, gc->maxY};
    Grid grid(cfg, ud2, n2);
    std::vector<int> c;
    grid.query(x, y, r, minLevel, maxLevel, ud2, (const Kp*)k2, c);
    for (int i = 0; i < (int)c.size() && i < cap; ++i) out[i] = c[i];
    return (int)c.size();
}

// Brute-force kNN-2 + ratio: the intended semantics of FtAssocOCV::match (OP_FtAssoc.cpp:63-99,
// SURVEY A12 / App. A.5).  norm: 0 = Hamming, 1 = L2 on the u8 bytes (BRUTEFORCE default).
// Lowest train index wins ties.  pass[i] = dist0 < ratio * dist1.
void orc_match_bf_knn2(const uint8_t* d1, int n1, const uint8_t* d2, int n2, int norm, float ratio,
                       int* idx0, int* idx1, float* dist0, float* dist1, uint8_t* pass) {
    for (int i = 0; i < n1; ++i) {
        int b0 = INT_MAX, b1 = INT_MAX, j0 = -1, j1 = -1;
        for (int j = 0; j < n2; ++j) {
            int d;
            if (norm == 0) d = hamming256(d1 + (size_t)i * 32, d2 + (size_t)j * 32);
            else {
                d = 0;
                for (int k = 0; k < 32; ++k) { int t = (int)d1[(size_t)i * 32 + k] - (int)d2[(size_t)j * 32 + k]; d += t * t; }
            }
            if (d < b0) { b1 = b0; j1 = j0; b0 = d; j0 = j; }
            else if (d < b1) { b1 = d; j1 = j; }
        }
        idx0[i] = j0; idx1[i] = j1;
        float f0 = (j0 < 0) ? 0.f : (norm == 0 ? (float)b0 : sqrtf((float)b0));
        float f1 = (j1 < 0) ? 0.f : (norm == 0 ? (float)b1 : sqrtf((float)b1));
        dist0[i] = f0; dist1[i] = f1;
        pass[i] = (j0 >= 0 && j1 >= 0 && f0 < ratio * f1) ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------
// Calibration::undistort (core/sensor/camera/Calibration.cpp:135-149) -> GeometricCamera::UndistortKeyPoints:
//   Pinhole        identity                                        (models/Pinhole.hpp:75-78)
//   PinholeRadTan  cv::undistortPoints(pts, pts, K, D, R=I, P=K)     (models/PinholeRadTan.cpp:11-25)
//   KannalaBrandt8 cv::fisheye::undistortPoints(pts, pts, K, D, I, K) (models/KannalaBrandt8.cpp:231-243)
// K = float (fx 0 cx; 0 fy cy; 0 0 1), D = 4 floats, R = eye, P = K.clone() (models/GeometricCamera.h:62-66).
// The arithmetic lives in OpenCV (un-vendored); restated here from its published algorithm and pinned bit-exact
// against cv2 4.13.0 by tests/test_oracle_vs_cv2.py.  All arithmetic is double, inputs and outputs float.
// ------------------------------------------------------------------------------------------
// cv::undistortPoints without criteria = TermCriteria(MAX_ITER, 5, 0.01): exactly five fixed-point iterations.
void orc_undistort_radtan(const float* K4 /*fx fy cx cy*/, const float* D4 /*k1 k2 p1 p2*/, const float* xy, int n, float* out) {
    const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
    const double ifx = 1. / fx, ify = 1. / fy;
    const double k0 = D4[0], k1 = D4[1], k2 = D4[2], k3 = D4[3];      // k[4..13] = 0
    for (int i = 0; i < n; ++i) {
        double x = xy[2 * i], y = xy[2 * i + 1];
        const double u = x, v = y;
        x = (x - cx) * ifx;
        y = (y - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; ++j) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((0. * r2 + 0.) * r2 + 0.) * r2) / (1 + ((0. * r2 + k1) * r2 + k0) * r2);
            if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
            const double deltaX = 2 * k2 * x * y + k3 * (r2 + 2 * x * x) + 0. * r2 + 0. * r2 * r2;
            const double deltaY = k2 * (r2 + 2 * y * y) + 2 * k3 * x * y + 0. * r2 + 0. * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        // RR = P * R = K: (xx, yy, ww) = RR * (x, y, 1)
        const double xx = fx * x + 0. * y + cx, yy = 0. * x + fy * y + cy, ww = 1. / (0. * x + 0. * y + 1.);
        out[2 * i] = (float)(xx * ww);
        out[2 * i + 1] = (float)(yy * ww);
    }
}

// cv::fisheye::undistortPoints with the default criteria (COUNT + EPS, 10, 1e-8): Newton iterations on theta.
void orc_undistort_kb8(const float* K4, const float* D4, const float* xy, int n, float* out) {
    const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
    const double k[4] = {D4[0], D4[1], D4[2], D4[3]};
    const double eps = 1e-8, kPi2 = 3.1415926535897932384626433832795 / 2.;
    for (int i = 0; i < n; ++i) {
        const double px = xy[2 * i], py = xy[2 * i + 1];
        const double pwx = (px - cx) / fx, pwy = (py - cy) / fy;
        double theta_d = std::sqrt(pwx * pwx + pwy * pwy);
        theta_d = std::min(std::max(-kPi2, theta_d), kPi2);
        bool converged = false;
        double theta = theta_d, scale = 0.0;
        if (std::fabs(theta_d) > eps) {
            for (int j = 0; j < 10; ++j) {
                const double theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta6 * theta2;
                const double k0_theta2 = k[0] * theta2, k1_theta4 = k[1] * theta4, k2_theta6 = k[2] * theta6, k3_theta8 = k[3] * theta8;
                const double theta_fix = (theta * (1 + k0_theta2 + k1_theta4 + k2_theta6 + k3_theta8) - theta_d) /
                                         (1 + 3 * k0_theta2 + 5 * k1_theta4 + 7 * k2_theta6 + 9 * k3_theta8);
                theta = theta - theta_fix;
                if (std::fabs(theta_fix) < eps) { converged = true; break; }
            }
            scale = std::tan(theta) / theta_d;
        } else {
            converged = true;
        }
        const bool flipped = (theta_d < 0 && theta > 0) || (theta_d > 0 && theta < 0);
        if (converged && !flipped) {
            const double pux = pwx * scale, puy = pwy * scale;
            const double prx = fx * pux + 0. * puy + cx * 1.0, pry = 0. * pux + fy * puy + cy * 1.0, prz = 0. * pux + 0. * puy + 1.0 * 1.0;
            out[2 * i] = (float)(prx / prz);
            out[2 * i + 1] = (float)(pry / prz);
        } else {
            out[2 * i] = -1000000.f;
            out[2 * i + 1] = -1000000.f;
        }
    }
}

// TwoViewReconstruction::CheckHomography (core/operators/mapInit/OP_2ViewReconstruction.cpp:447-530): H21 / H12 row-major
// 3x3, matched points in match order; returns the score, inl[i] = vbMatchesInliers[i].  Plain float arithmetic in the
// reference's operation order (this file is built with -ffp-contract=off).
float orc_check_homography(const float* H21, const float* H12, const float* xy1, const float* xy2, int n, float sigma, float th,
                           uint8_t* inl) {
    const float h11 = H21[0], h12 = H21[1], h13 = H21[2], h21 = H21[3], h22 = H21[4], h23 = H21[5], h31 = H21[6], h32 = H21[7], h33 = H21[8];
    const float h11inv = H12[0], h12inv = H12[1], h13inv = H12[2], h21inv = H12[3], h22inv = H12[4], h23inv = H12[5], h31inv = H12[6],
                h32inv = H12[7], h33inv = H12[8];
    float score = 0;
    const float invSigmaSquare = 1.0f / (sigma * sigma);
    for (int i = 0; i < n; i++) {
        bool bIn = true;
        const float u1 = xy1[2 * i], v1 = xy1[2 * i + 1], u2 = xy2[2 * i], v2 = xy2[2 * i + 1];
        const float w2in1inv = 1.0f / (h31inv * u2 + h32inv * v2 + h33inv);
        const float u2in1 = (h11inv * u2 + h12inv * v2 + h13inv) * w2in1inv;
        const float v2in1 = (h21inv * u2 + h22inv * v2 + h23inv) * w2in1inv;
        const float squareDist1 = (u1 - u2in1) * (u1 - u2in1) + (v1 - v2in1) * (v1 - v2in1);
        const float chiSquare1 = squareDist1 * invSigmaSquare;
        if (chiSquare1 > th) bIn = false; else score += th - chiSquare1;
        const float w1in2inv = 1.0f / (h31 * u1 + h32 * v1 + h33);
        const float u1in2 = (h11 * u1 + h12 * v1 + h13) * w1in2inv;
        const float v1in2 = (h21 * u1 + h22 * v1 + h23) * w1in2inv;
        const float squareDist2 = (u2 - u1in2) * (u2 - u1in2) + (v2 - v1in2) * (v2 - v1in2);
        const float chiSquare2 = squareDist2 * invSigmaSquare;
        if (chiSquare2 > th) bIn = false; else score += th - chiSquare2;
        if (inl) inl[i] = bIn ? 1 : 0;
    }
    return score;
}

// TwoViewReconstruction::CheckFundamental (:532-610)
float orc_check_fundamental(const float* F21, const float* xy1, const float* xy2, int n, float sigma, float th, float thScore,
                            uint8_t* inl) {
    const float f11 = F21[0], f12 = F21[1], f13 = F21[2], f21 = F21[3], f22 = F21[4], f23 = F21[5], f31 = F21[6], f32 = F21[7], f33 = F21[8];
    float score = 0;
    const float invSigmaSquare = 1.0f / (sigma * sigma);
    for (int i = 0; i < n; i++) {
        bool bIn = true;
        const float u1 = xy1[2 * i], v1 = xy1[2 * i + 1], u2 = xy2[2 * i], v2 = xy2[2 * i + 1];
        const float a2 = f11 * u1 + f12 * v1 + f13;
        const float b2 = f21 * u1 + f22 * v1 + f23;
        const float c2 = f31 * u1 + f32 * v1 + f33;
        const float num2 = a2 * u2 + b2 * v2 + c2;
        const float squareDist1 = num2 * num2 / (a2 * a2 + b2 * b2);
        const float chiSquare1 = squareDist1 * invSigmaSquare;
        if (chiSquare1 > th) bIn = false; else score += thScore - chiSquare1;
        const float a1 = f11 * u2 + f21 * v2 + f31;
        const float b1 = f12 * u2 + f22 * v2 + f32;
        const float c1 = f13 * u2 + f23 * v2 + f33;
        const float num1 = a1 * u1 + b1 * v1 + c1;
        const float squareDist2 = num1 * num1 / (a1 * a1 + b1 * b1);
        const float chiSquare2 = squareDist2 * invSigmaSquare;
        if (chiSquare2 > th) bIn = false; else score += thScore - chiSquare2;
        if (inl) inl[i] = bIn ? 1 : 0;
    }
    return score;
}

}  // extern "C"
