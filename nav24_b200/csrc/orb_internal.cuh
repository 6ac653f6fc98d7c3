// orb_internal.cuh — shared device-side layout of the B200 ORB front end (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nav24_orb.h"

namespace nav24 {

constexpr int kMaxLevels = 16;
constexpr int kEdge = 19;           // EDGE_THRESHOLD (OP_FtDtOrbSlam.cpp:14)
constexpr int kMinBorder = 16;      // EDGE_THRESHOLD-3 (:735)
constexpr int kBlurTileRows = 36;                    // rows per warp tile of blur_kernel (multiple of 6: the row-pair ring)
constexpr int kBlurCtaRows = 4 * kBlurTileRows;     // rows per CTA tile (four warp tiles stacked)
constexpr int kBlurBoxW = 160, kBlurBoxH = kBlurCtaRows + 6;   // TMA box: 128 px + 16-byte aligned halos, 3 halo rows each side
constexpr int kMaxCellTile = 76;    // wCell+6 <= 75 whenever nCols >= 1 (cell pitch < 70)
constexpr int kOriBoxW = 48, kOriBoxH = 31;    // TMA box of the orientation patch: 31 px + <= 15 px of alignment slack, 16-B multiple
constexpr int kDescBoxW = 80, kDescBoxH = 37;   // TMA box of the descriptor patch: 37 px + <= 15 px of alignment slack = 52 -> 64 would do, but a 20-word row pitch
                                               // spreads the hot central columns over all 32 banks (16 words: 4.8 wavefronts per gather instead of 3.1)
constexpr int kDescBoxWN = 48;                  // narrow box of the descriptor patch for keypoints whose alignment slack is <= 11 px (12-word pitch: 3.1 too)
constexpr int kFastThreads = 256;   // threads per CTA of fast_band_kernel
constexpr int kFastMaxSegCells = 8; // cells of one FAST segment (TMA boxes are <= 256 px wide, cells >= 35 px)
constexpr int kFastQueueCap = 3072; // stage A survivors one CTA queues (typ. 700 of 8400 pixels); beyond: the dense path
#ifndef NAV24_FAST_PITCH
#define NAV24_FAST_PITCH 256
#endif
constexpr int kFastPitch = NAV24_FAST_PITCH;     // row pitch of the FAST tile in shared memory = TMA box width (multiple of 16, <= 256), the same for every level, so
                                    // that every tap of the kernel is an immediate offset

// device error bits (ctx->d_err)
enum : int { ERR_RAW_OVERFLOW = 1, ERR_ROOT_RANGE = 2, ERR_NODE_OVERFLOW = 4, ERR_CELL_SIZE = 8, ERR_KP_OVERFLOW = 16 };

struct LevelGeom {
    int w, h;
    int pitch;               // row pitch in bytes of the level inside the pyramid / blurred buffers
    long long off;           // byte offset of the level inside one frame's pyramid slab (level 0: unused)
    long long boff;          // byte offset inside one frame's blurred slab
    // FAST cell grid (OP_FtDtOrbSlam.cpp:735-768)
    int nCols, nRows, wCell, hCell, maxBX, maxBY;
    int segCols, segsPerRow; // FAST segment = segCols horizontally adjacent cells of one cell row (the last may hold fewer)
    int boxW, boxH;          // TMA box of one FAST segment: kFastPitch x (hCell+6)
    unsigned magicW;         // 0xFFFFFFFF / wCell + 1: floor(n / wCell) = umulhi(n, magicW) for n < 65536
    unsigned magicH;         // the same for hCell
    int cellBase;            // first cell id of this level inside the per-frame cell table
    int blurTileBase;        // first CTA tile (128 px x kBlurCtaRows rows) of this level in blur_kernel
    int rawCap;              // raw-corner capacity of this level (records)
    int rawOff;              // record offset of the level inside one frame's raw slab
    // quadtree (:502-725)
    int quota;               // mnFeaturesPerLevel[level]
    int nIni;                // number of root nodes
    float hX;                // root width
    int nodeCap;             // node capacity
    int nodeOff;             // offset (in nodes) inside one frame's node slabs
    int kpOff;               // offset of the level inside one frame's level-keypoint slab
    int kpCap;               // quota + 3 ... (max(4*nIni, quota+3)+4 == nodeCap)
    float scale;             // mvScaleFactor[level]
    float patch;             // float(int(31*scale))
};

struct FrameGeom {
    int nlevels;
    int totalCells;
    int totalSegs;           // FAST segments per frame (grid.x of fast_band_kernel)
    int fastCutH, segsLow;   // the segments of the levels with boxH <= fastCutH come first in the table (segsLow of them) and are
                             // launched with a smaller shared-memory footprint (launch_fast); segsLow == totalSegs: one launch
    int rawPerFrame;         // records
    int nodesPerFrame;
    int kpPerFrame;          // level-keypoint slab size (sum of kpCap)
    int outCap;              // output capacity per frame
    int blurTiles;           // CTA tiles of blur_kernel per frame
    long long pyrFrameBytes, blurFrameBytes;
    LevelGeom lv[kMaxLevels];
};

struct RawRec {              // one FAST survivor, 8 bytes
    unsigned short x, y;     // relative to (minBorderX, minBorderY)
    unsigned short score;
    unsigned short pad;
};

struct FastSeg {             // one CTA of fast_band_kernel: a run of cells of one cell row (host-built, 32 bytes)
    short level, ci, cj0, nc;       // nc cells [cj0, cj0+nc) of cell row ci (entries of the cell table, skipped ones included)
    short nv, iniX0, iniY, ih;      // nv valid cells (a prefix of the nc), image position of the tile, interior rows
    short iw, ngx, nChunks, rc;     // interior columns of the nv cells together; stage A items: ngx word columns x
                                    // nChunks chunks of rc rows (one round of the CTA's threads)
    int cell0;                      // first cell of the segment in the per-frame cell table
    unsigned magicG;                // 0xFFFFFFFF / ngx + 1
};

struct QNode {               // 12 bytes
    short x0, y0, x1, y1;
    int count;
};

struct LevelKp {             // keypoint in level coordinates, quadtree order (16 bytes)
    unsigned short x, y;     // level pixel coordinates (border added)
    unsigned short score;
    unsigned short pad;
    int dst;                 // index in the frame's final output order
    float angle;
};

// all device pointers of one workspace, passed by value to kernels
struct DevPtrs {
    const uint8_t* l0; long long l0Pitch, l0Frame;   // level 0 (caller's buffer or staging copy)
    uint8_t* pyr;            // levels >= 1
    uint8_t* blur;           // blurred levels
    uint2* cellInfo;         // [B][totalCells] (offset inside level raw slab, count)
    int* cellDst;            // [B][totalCells] ordered start offset
    int* rawCount;           // [B][nlevels] atomic append cursors
    RawRec* raw;             // [B][rawPerFrame] append order
    RawRec* keys;            // [B][rawPerFrame] reference order (vToDistributeKeys)
    int* nodeOfKey;          // [B][rawPerFrame]
    QNode* nodesA; QNode* nodesB;   // [B][nodesPerFrame]
    int* childCnt;           // [B][4*nodesPerFrame]
    int* nodeAux;            // [B][nodesPerFrame]
    unsigned long long* best;       // [B][nodesPerFrame]
    unsigned long long* sortRec;    // [B][nodesPerFrame]
    LevelKp* lkp;            // [B][kpPerFrame]
    int* levelCount;         // [B][nlevels] keypoints per level after the quadtree
    int* rawTotal;           // [B][nlevels] raw keys per level
    int* frameDone;          // [B] levels of the frame whose quadtree CTA has finished (zero between launches)
    float* outUd;            // [B][outCap][2] undistorted keypoint coordinates (camera model != pinhole)
    nav24_kp* outKp;         // [B][outCap]
    uint8_t* outDesc;        // [B][outCap][32]
    int* nOut; int* monoOut; // [B]
    const unsigned* oriTab;  // orientation DP4A weights [4][279][2] (describe_kernel)
    const FastSeg* segs;     // [totalSegs] FAST segments of one frame
    int* err;                // device error bits
    int frameBase;           // first frame of this chunk inside the buffers the TMA maps were encoded over
};

struct TmaMaps { CUtensorMap m[kMaxLevels]; };   // one 3-D (x, y, frame) u8 tensor map per pyramid level

struct FastSmem { int offMap, offMask, offQueue, total; };   // dynamic shared memory layout of fast_band_kernel

struct ResizeTab {           // per level >= 1: source offsets and 11-bit coefficient pairs
    const int* xofs; const short2* xab; const int* yofs; const short2* yab;
    int rows;                // destination rows per warp strip (<= kResizeMaxRows)
    int boxW, boxH;          // TMA box over the SOURCE level that covers one CTA tile (128 x 4*rows destination pixels)
    int wide;                // 1: every group of eight destination pixels fits resize8_kernel's shared four-word window
};

constexpr int kResizeMaxRows = 32;     // destination rows of one warp strip of resize_kernel (one lane per row's constants)
#ifndef NAV24_RESIZE_CTA_ROWS
#define NAV24_RESIZE_CTA_ROWS 64
#endif
constexpr int kResizeCtaRows = NAV24_RESIZE_CTA_ROWS;     // target destination rows per CTA (4 warp strips)

// launchers (orb_kernels.cu); each returns the number of kernels launched
// tight (pitch = w) host-order frames -> 16-byte aligned pitch (TMA needs it); returns 1
int launch_repack(const uint8_t* src, int w, int h, uint8_t* dst, int dPitch, int B, cudaStream_t s);
// tight interleaved BGR host-order frames -> grey at the aligned pitch (cv::cvtColor COLOR_BGR2GRAY fixed point); returns 1
int launch_bgr2gray(const uint8_t* src, int w, int h, uint8_t* dst, int dPitch, int B, cudaStream_t s);
int launch_pyramid(const FrameGeom& g, const DevPtrs& p, const ResizeTab* tabs, const TmaMaps& mapsSrc, int B, cudaStream_t s);
void launch_fast_prepare(const FrameGeom& g, const DevPtrs& p, int B, cudaStream_t s);      // before launch_pyramid
int launch_fast(const FrameGeom& g, const DevPtrs& p, const TmaMaps& maps, int B, int iniTh, int minTh, cudaStream_t s);
int fast_smem_bytes(const FrameGeom& g, int minBoxH, int maxBoxH, FastSmem* out);      // levels with minBoxH < boxH <= maxBoxH
int launch_quadtree(const FrameGeom& g, const DevPtrs& p, int B, cudaStream_t s);
int launch_blur(const FrameGeom& g, const DevPtrs& p, const TmaMaps& mapsBlurSrc, int B, cudaStream_t s);
int launch_describe(const FrameGeom& g, const DevPtrs& p, const TmaMaps& mapsOri, const TmaMaps& mapsBlur, const TmaMaps& mapsBlurN, int B,
                    cudaStream_t s);
int launch_debug_sort(unsigned long long* d_recs, int n, cudaStream_t s);   // test hook (stdsort_warp.cuh)

// matchers (match_kernels.cu)
struct MatchArgs {
    const nav24_kp* k1; const float* ud1; const uint8_t* d1; const int* n1;   // [P][cap]...
    const nav24_kp* k2; const float* ud2; const uint8_t* d2; const int* n2;
    long long stride1, stride2;    // element stride between pairs (keypoints); 0 with index lists
    const int* pairs;              // optional [P][2] frame indices into k1/k2 (same arrays), else null
    const int* pairOrder;          // optional: launch block b handles pair pairOrder[pairBase + b] (chunked pipeline)
    int pairBase;                  // first pair (or first entry of pairOrder) of this launch
    nav24_grid_cfg grid; float invW, invH;
    float window, nnratio; int thLow, checkOri;
    int cap;
    int* cellOf; int* cellStart; int* cellFill; int* cellItems;  // grid scratch per pair
    int* cand; int* candCnt; int candCap;                          // pruned candidate lists per query
    int* dist2; int* m21; int* bins;                               // [P][cap]
    int* matches12; int* nMatches;
};
int launch_match_window(const MatchArgs& a, int P, cudaStream_t s);
int bf_knn2_segments(int n1, int n2);      // rows of the [segments][n1] int4 scratch launch_bf_knn2 needs
int launch_bf_knn2(const uint8_t* d1, int n1, const uint8_t* d2, int n2, int norm, float ratio, int* idx0, int* idx1,
                   float* dist0, float* dist1, uint8_t* pass, int4* part, cudaStream_t s);

// camera models (camera_kernels.cu): Calibration::undistort on the device
int launch_undistort_points(const nav24_camera& cam, const float* xy, int n, float* out, cudaStream_t s);
int launch_undistort_frames(const nav24_camera& cam, const nav24_kp* kps, const int* nOut, int cap, int B, float* ud, cudaStream_t s);

// two-view RANSAC scoring (camera_kernels.cu): CheckHomography / CheckFundamental of every hypothesis in one launch
int launch_two_view_score(const float* xy1, const float* xy2, int n, const float* H21, const float* H12, const float* F21, int nHyp,
                          float sigma, float thH, float thF, float thScore, float* scoreH, float* scoreF, uint8_t* inH, uint8_t* inF,
                          cudaStream_t s);

}  // namespace nav24
