// capi.cu — C ABI (include/nav24_orb.h) on top of the sm_100a kernels: context, device workspace,
// stream orchestration.  Host-side set-up math (scale tables, quotas, level sizes, cell grids,
// resize coefficient tables) restates the reference's float/double arithmetic exactly:
//   ctor / setNumFeatures   core/operators/objDetection/OP_FtDtOrbSlam.cpp:441-500, :962-976
//   level sizes             :940        cell grid  :735-768      quadtree roots  :505-527
// There is no CPU compute path: every entry point needs a CUDA device.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "orb_internal.cuh"

using namespace nav24;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return ctx->fail(NAV24_E_CUDA, #call, e_);                          \
    } while (0)

namespace {

struct DevBuf {
    void* ptr = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr; bytes = 0;
        size_t want = need + need / 8 + 256;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; bytes = 0; }
};

}  // namespace

struct nav24_orb {
    int device = 0;
    nav24_orb_params prm{};
    int nIniFeatures = 0;
    double scaleFactorD = 1.2;
    std::vector<float> scale, invScale;
    std::vector<int> quota;
    std::string err;
    static constexpr int kMaxStreams = 4;
    cudaStream_t stream = nullptr, copyStream = nullptr, outStream = nullptr;
    cudaStream_t xstream[kMaxStreams] = {nullptr, nullptr, nullptr, nullptr};      // [0] = stream; [1..] extra compute streams
    cudaEvent_t evJoinX[kMaxStreams] = {nullptr, nullptr, nullptr, nullptr};
    // one side stream per compute stream: the blur of a SMALL chunk (it needs only the pyramid) runs there, next to the
    // latency-bound quadtree of the same chunk: single KITTI frame 0.207 -> 0.198 ms.  Batches that fill the GPU gain nothing
    // from it (measured: 6.52 vs 6.54 ms per 1024 frames), so they keep the whole chain on one stream (NAV24_BLUR_FORK=0: never)
    cudaStream_t sstream[kMaxStreams] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t evFork[kMaxStreams] = {nullptr, nullptr, nullptr, nullptr}, evBlur[kMaxStreams] = {nullptr, nullptr, nullptr, nullptr};
    int blurFork = 1;
    int nStreams = 3;                          // compute streams the chunks rotate over (NAV24_STREAMS)
    int taper = 1;                             // host pipeline: short first and last chunks (NAV24_TAPER=0: uniform)
    std::vector<cudaEvent_t> evIn, evDone;      // per chunk of the host-buffer pipeline (no timing)
    cudaEvent_t evJoin = nullptr, evPrevEnd = nullptr, evPairs = nullptr;
    cudaEvent_t evSlotDone[2] = {nullptr, nullptr};      // end of the last call that used pair-table slot 0 / 1
    bool slotUsed[2] = {false, false};
    bool stagesValid = false;                  // the stage events belong to the last detect call (nav24_orb_detect_device only)
    bool trace = false;                      // NAV24_TRACE=1: per-chunk timeline of the host pipeline on stderr
    cudaEvent_t evT0 = nullptr, evT1 = nullptr;
    bool prevEndValid = false;
    unsigned long long prevSig = 0;           // (B, chunk, pairs) signature of the previous chunked call
    int callParity = 0;                       // ping-pong index of the pair tables
    int lastP = 0, lastMatchCap = 0;          // pairs / row capacity of the match results left on the device
    int chunkFrames = 64;                      // frames per pipeline chunk of the host-buffer entry points
    int residentChunk = 0;                     // 0 = whole batch in one chunk for device-resident frames
    int* hN = nullptr; int* hMono = nullptr; int* hNm = nullptr; int* hErr = nullptr; int hCap = 0;   // pinned result scratch
    static constexpr int kEvRing = 64;
    cudaEvent_t evRing[kEvRing][5]{};
    long long evCalls = 0;      // pipeline runs since the last stage-sum reset
    cudaEvent_t* ev = evRing[0];
    cudaEvent_t evT[2]{};
    long long launches = 0;

    // workspace keyed on (w, h, nFeatures); batch capacity grows on demand
    int wsW = 0, wsH = 0, wsB = 0, wsFeat = -1;
    FrameGeom g{};
    DevPtrs p{};
    std::vector<ResizeTab> tabs;
    // single-frame host calls are launch-bound (13 kernels of a few CTAs each): their kernel chain is captured once per
    // (shape, features, camera, workspace) as a CUDA graph and replayed (NAV24_GRAPH=0 turns it off)
    int useGraph = 1;
    cudaGraphExec_t graphExec = nullptr;
    unsigned long long graphKey = 0, wsGen = 0;
    long long graphLaunches = 0;
    nav24_camera cam{};       // camera of the fused paths (model 0 = pinhole: identity undistortion)
    TmaMaps maps{};           // FAST segment tiles over the un-blurred levels
    TmaMaps mapsRs{};         // m[l]: resize source tiles over level l-1 (resize_kernel producing level l)
    TmaMaps mapsBlurSrc{};    // 160 x 146 source tiles of blur_kernel over the un-blurred levels
    TmaMaps mapsOri{};        // 48 x 31 orientation patches over the un-blurred levels (describe_kernel)
    TmaMaps mapsBlur{};       // 80 x 37 descriptor patches over the blurred levels (describe_kernel)
    TmaMaps mapsBlurN{};      // 48 x 37: the narrow box for keypoints with little alignment slack
    const void* mapsBlurPtr = nullptr;
    int mapsB = 0;            // frame count the level>=1 maps were encoded for
    const void* mapsPyr = nullptr;
    DevBuf bL0Tight, bL0, bPyr, bBlur, bCell, bCellDst, bRawCount, bRaw, bKeys, bNodeOfKey, bNodesA, bNodesB, bChild, bAux, bBest,
        bSort, bLkp, bLevelCount, bRawTotal, bFrameDone, bOutKp, bOutDesc, bNOut, bMono, bErr, bTabs, bOriTab, bSegs, bOutUd, bUdTmp;
    // matcher scratch
    DevBuf mK1, mK2, mU1, mU2, mD1, mD2, mN1, mN2, mCellOf, mCellStart, mCellFill, mCellItems, mCand, mCandCnt, mDist2,
        mM21, mBins, mMatches, mNMatches, mPairs, mPairOrder, mI0, mI1, mF0, mF1, mPass;
    int lastB = 0;            // frames of the last detect call
    bool lastValid = false;
    int l0Pitch = 0;

    cudaError_t ensure_events(int n) {
        while ((int)evIn.size() < n) {
            cudaEvent_t a, b;
            const unsigned fl = trace ? cudaEventDefault : cudaEventDisableTiming;
            cudaError_t e = cudaEventCreateWithFlags(&a, fl);
            if (e != cudaSuccess) return e;
            e = cudaEventCreateWithFlags(&b, fl);
            if (e != cudaSuccess) return e;
            evIn.push_back(a); evDone.push_back(b);
        }
        return cudaSuccess;
    }
    cudaError_t ensure_host(int B) {
        if (B <= hCap) return cudaSuccess;
        if (hN) cudaFreeHost(hN);
        hN = nullptr; hCap = 0;
        const int want = B + B / 2 + 16;
        cudaError_t e = cudaHostAlloc((void**)&hN, sizeof(int) * (3 * (size_t)want + 4), cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        hMono = hN + want; hNm = hMono + want; hErr = hNm + want; hCap = want;
        return cudaSuccess;
    }

    int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
        err = what;
        if (e != cudaSuccess) { err += ": "; err += cudaGetErrorString(e); }
        return code;
    }

    void compute_quota(int n) {
        prm.n_features = n;
        const int nl = prm.n_levels;
        float factor = 1.0f / scaleFactorD;
        float per = n * (1 - factor) / (1 - (float)pow((double)factor, (double)nl));
        int sum = 0;
        for (int l = 0; l < nl - 1; ++l) {
            quota[l] = (int)lrintf(per);
            sum += quota[l];
            per *= factor;
        }
        quota[nl - 1] = std::max(n - sum, 0);
    }
};

namespace {

// Nothing throws across the C boundary: every entry point that can allocate host memory runs inside this guard
// (std::bad_alloc -> NAV24_E_NOMEM, anything else -> NAV24_E_CUDA; the context keeps a short message when it can).
template <class F> int guarded(nav24_orb* ctx, F&& body) noexcept {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        if (ctx) { try { ctx->err = "out of host memory"; } catch (...) {} }
        return NAV24_E_NOMEM;
    } catch (...) {
        if (ctx) { try { ctx->err = "unexpected C++ exception"; } catch (...) {} }
        return NAV24_E_CUDA;
    }
}

inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// Builds the frame geometry for (w,h). Returns NAV24_OK or NAV24_E_GEOMETRY.
int build_geometry(nav24_orb* ctx, int w, int h, FrameGeom& g) {
    memset(&g, 0, sizeof(g));
    const int nl = ctx->prm.n_levels;
    g.nlevels = nl;
    const int rawPerKpx = ctx->prm.raw_keys_per_kpx > 0 ? ctx->prm.raw_keys_per_kpx : 125;
    long long off = 0, boff = 0;
    int cellBase = 0, rawOff = 0, nodeOff = 0, kpOff = 0, stripBase = 0;
    for (int l = 0; l < nl; ++l) {
        LevelGeom& L = g.lv[l];
        L.w = (int)lrintf((float)w * ctx->invScale[l]);
        L.h = (int)lrintf((float)h * ctx->invScale[l]);
        if (L.w >= 8192 || L.h >= 8192) return ctx->fail(NAV24_E_GEOMETRY, "image larger than 8191 px not supported");
        L.pitch = align_up(L.w, 128);
        L.off = off;
        if (l > 0) off += (long long)L.pitch * L.h;
        L.boff = boff;
        boff += (long long)L.pitch * L.h;
        const int minBX = kMinBorder, minBY = kMinBorder;
        L.maxBX = L.w - kEdge + 3;
        L.maxBY = L.h - kEdge + 3;
        const float width = (float)(L.maxBX - minBX), height = (float)(L.maxBY - minBY);
        if (width < 35.f || height < 35.f) return ctx->fail(NAV24_E_GEOMETRY, "pyramid level smaller than one 35-px FAST cell");
        L.nCols = (int)(width / 35.f);
        L.nRows = (int)(height / 35.f);
        L.wCell = (int)std::ceil(width / L.nCols);
        L.hCell = (int)std::ceil(height / L.nRows);
        if (L.wCell + 6 > kMaxCellTile || L.hCell + 6 > kMaxCellTile) return ctx->fail(NAV24_E_GEOMETRY, "FAST cell larger than the kernel tile");
        // FAST segments: runs of cells of one cell row whose tile (interior + 3-px rim + <= 16 px of alignment slack)
        // fits one TMA box (<= 256 px wide); the cells of a row are dealt evenly over the segments
        // ... and whose cells x rows fit one round of the CTA's threads (the count + emit step has a thread per cell row)
        const int fit = std::max(1, std::min(std::min(kFastMaxSegCells, (kFastPitch - 22) / L.wCell), kFastThreads / L.hCell));
        L.segsPerRow = (L.nCols + fit - 1) / fit;
        L.segCols = (L.nCols + L.segsPerRow - 1) / L.segsPerRow;
        L.segsPerRow = (L.nCols + L.segCols - 1) / L.segCols;
        L.boxW = kFastPitch;
        L.boxH = L.hCell + 6;
        L.magicW = 0xFFFFFFFFu / (unsigned)L.wCell + 1u;
        L.magicH = 0xFFFFFFFFu / (unsigned)L.hCell + 1u;
        L.cellBase = cellBase;
        L.blurTileBase = stripBase;
        stripBase += ((L.w + 127) / 128) * ((L.h + kBlurCtaRows - 1) / kBlurCtaRows);
        cellBase += L.nCols * L.nRows;
        long long cap = ((long long)L.w * L.h * rawPerKpx + 999) / 1000 + 64;
        if (cap >= (1 << 19)) cap = (1 << 19) - 1;      // sort key packs count into 19 bits
        L.rawCap = (int)cap;
        L.rawOff = rawOff;
        rawOff += align_up(L.rawCap, 4);
        L.quota = ctx->quota[l];
        L.nIni = (int)std::round((float)(L.maxBX - minBX) / (L.maxBY - minBY));
        if (L.nIni < 1) return ctx->fail(NAV24_E_GEOMETRY, "image taller than 2:1, the quadtree has no root node");
        L.hX = (float)(L.maxBX - minBX) / (float)L.nIni;
        L.nodeCap = align_up(std::max(4 * L.nIni, L.quota + 3) + 4, 4);
        if (L.nodeCap > 65535) return ctx->fail(NAV24_E_GEOMETRY, "more than 65535 quadtree nodes per level (node indices are 16-bit)");
        L.nodeOff = nodeOff;
        nodeOff += L.nodeCap;
        L.kpOff = kpOff;
        L.kpCap = L.nodeCap;
        kpOff += L.kpCap;
        L.scale = ctx->scale[l];
        L.patch = (float)(int)(31 * ctx->scale[l]);
    }
    g.totalCells = cellBase;
    g.totalSegs = 0;
    for (int l = 0; l < nl; ++l) g.totalSegs += g.lv[l].segsPerRow * g.lv[l].nRows;
    // FAST launch groups (launch_fast): the cut below the tallest cell rows that lets the most segments run with one more
    // resident CTA per SM than a single launch sized for the tallest level would (228 KB per SM, 1 KB reserved per CTA)
    {
        auto ctas = [&](int lo, int hi) { const int b = fast_smem_bytes(g, lo, hi, nullptr); return b ? (228 * 1024) / (b + 1024) : 0; };
        const int all = ctas(0, 1 << 30);
        g.fastCutH = 1 << 30; g.segsLow = g.totalSegs;
        int bestSegs = 0;
        for (int l = 0; l < nl; ++l) {
            const int cut = g.lv[l].boxH;
            if (ctas(0, cut) <= all) continue;
            int segs = 0;
            for (int k = 0; k < nl; ++k) if (g.lv[k].boxH <= cut) segs += g.lv[k].segsPerRow * g.lv[k].nRows;
            if (segs > bestSegs && segs < g.totalSegs) { bestSegs = segs; g.fastCutH = cut; g.segsLow = segs; }
        }
    }
    g.blurTiles = stripBase;
    g.rawPerFrame = rawOff;
    g.nodesPerFrame = nodeOff;
    g.kpPerFrame = kpOff;
    g.outCap = kpOff;
    g.blurFrameBytes = (boff + 255) / 256 * 256;
    g.pyrFrameBytes = (off + 255) / 256 * 256;
    return NAV24_OK;
}

// cv::resize coefficient tables, computed exactly like OpenCV does on the host.
void build_resize_table(int ssize, int dsize, std::vector<int>& ofs, std::vector<short2>& ab) {
    ofs.resize(dsize); ab.resize(dsize);
    const double inv = (double)dsize / ssize, scale = 1.0 / inv;
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
        ofs[d] = s;
        ab[d].x = (short)lrintf((1.f - f) * 2048.f);
        ab[d].y = (short)lrintf(f * 2048.f);
    }
}

// FAST segment table of one frame (fast_band_kernel): the cell loop bounds of ComputeKeyPointsOctTree
// (OP_FtDtOrbSlam.cpp:751-768) evaluated once on the host.
void build_fast_segments(const FrameGeom& g, std::vector<FastSeg>& segs) {
    segs.clear();
    for (int pass = 0; pass < 2; ++pass)      // the levels of the low group first (g.fastCutH)
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom& L = g.lv[l];
        if ((L.boxH <= g.fastCutH) != (pass == 0)) continue;
        for (int ci = 0; ci < L.nRows; ++ci)
            for (int sj = 0; sj < L.segsPerRow; ++sj) {
                FastSeg s{};
                s.level = (short)l; s.ci = (short)ci; s.cj0 = (short)(sj * L.segCols);
                s.nc = (short)std::min(L.segCols, L.nCols - s.cj0);
                s.cell0 = L.cellBase + ci * L.nCols + s.cj0;
                const int iniY = kMinBorder + ci * L.hCell;
                const int maxY = std::min(iniY + L.hCell + 6, L.maxBY);
                s.iniY = (short)iniY; s.iniX0 = (short)(kMinBorder + s.cj0 * L.wCell);
                const bool rowSkip = iniY >= L.maxBY - 3 || maxY - iniY < 7;      // :756 (and cv::FAST on < 7 rows finds nothing)
                int nv = 0, iw = 0;
                for (int j = 0; j < s.nc && !rowSkip; ++j) {
                    const int iniX = kMinBorder + (s.cj0 + j) * L.wCell;
                    if (iniX >= L.maxBX - 6) break;                             // :765
                    const int maxX = std::min(iniX + L.wCell + 6, L.maxBX);
                    if (maxX - iniX < 7) break;
                    ++nv; iw += maxX - iniX - 6;
                }
                s.nv = (short)nv; s.iw = (short)iw; s.ih = (short)(nv ? maxY - iniY - 6 : 0);
                if (nv) {
                    const int o = (s.iniX0 - 1) & 15, c_lo = o + 4, c_hi = c_lo + iw;
                    s.ngx = (short)(((c_hi - 1) >> 2) - (c_lo >> 2) + 1);
                    s.nChunks = (short)std::min(std::max(kFastThreads / s.ngx, 1), (int)s.ih);
                    s.rc = (short)((s.ih + s.nChunks - 1) / s.nChunks);
                    s.magicG = 0xFFFFFFFFu / (unsigned)s.ngx + 1u;
                }
                segs.push_back(s);
            }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 3-D (x, y, frame) u8 tensor map over one pyramid level; box = one FAST cell tile
int encode_level_map(nav24_orb* ctx, CUtensorMap* m, const void* base, int w, int h, int frames, long long pitch,
                     long long frameStride, int boxW, int boxH) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return ctx->fail(NAV24_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frameStride};
    cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1u};
    cuuint32_t es[3] = {1u, 1u, 1u};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ctx->err = "cuTensorMapEncodeTiled failed, CUresult " + std::to_string((int)r);
        return NAV24_E_CUDA;
    }
    return NAV24_OK;
}

int ensure_workspace(nav24_orb* ctx, int w, int h, int B) {
    cudaSetDevice(ctx->device);
    const bool shapeChanged = (w != ctx->wsW || h != ctx->wsH || ctx->prm.n_features != ctx->wsFeat);
    if (shapeChanged) {
        FrameGeom g;
        int rc = build_geometry(ctx, w, h, g);
        if (rc != NAV24_OK) return rc;
        ctx->g = g;
        // resize tables
        const int nl = g.nlevels;
        std::vector<int> allOfs; std::vector<short2> allAb;
        std::vector<size_t> oX(nl), oY(nl);
        for (int l = 1; l < nl; ++l) {
            std::vector<int> o; std::vector<short2> a;
            build_resize_table(g.lv[l - 1].w, g.lv[l].w, o, a);
            oX[l] = allOfs.size(); allOfs.insert(allOfs.end(), o.begin(), o.end()); allAb.insert(allAb.end(), a.begin(), a.end());
            build_resize_table(g.lv[l - 1].h, g.lv[l].h, o, a);
            oY[l] = allOfs.size(); allOfs.insert(allOfs.end(), o.begin(), o.end()); allAb.insert(allAb.end(), a.begin(), a.end());
        }
        const size_t nT = allOfs.size();
        CK(ctx->bTabs.ensure(nT * 8 + 16));
        int* dOfs = (int*)ctx->bTabs.ptr;
        short2* dAb = (short2*)((char*)ctx->bTabs.ptr + nT * 4);
        if (nT) {
            CK(cudaMemcpyAsync(dOfs, allOfs.data(), nT * 4, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(dAb, allAb.data(), nT * 4, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
        ctx->tabs.assign(nl, ResizeTab{});
        for (int l = 1; l < nl; ++l) {
            ResizeTab& T = ctx->tabs[l];
            T = ResizeTab{dOfs + oX[l], dAb + oX[l], dOfs + oY[l], dAb + oY[l], 0, 0, 0, 0};
            // warp strip: `rows` destination rows; CTA tile: 128 x 4*rows destination pixels (the level's rows dealt evenly
            // over the CTA rows); the TMA box over the source level covers the CTA tile (x start rounded down to 16 bytes,
            // and the three aligned words each thread reads)
            const int* xo = allOfs.data() + oX[l]; const int* yo = allOfs.data() + oY[l];
            const int dw = g.lv[l].w, dh = g.lv[l].h;
            const int ctaRows = (dh + kResizeCtaRows - 1) / kResizeCtaRows;      // (64 rows per CTA measured best of 16 .. 128)
            T.rows = std::max(1, std::min(kResizeMaxRows, ((dh + ctaRows - 1) / ctaRows + 3) / 4));
            int needW = 0, needH = 0;
            for (int x0 = 0; x0 < dw; x0 += 128) {
                const int xl = std::min(x0 + 127, dw - 1) & ~3;
                needW = std::max(needW, (xo[xl] & ~3) + 12 - (xo[x0] & ~15));
            }
            for (int y0 = 0; y0 < dh; y0 += 4 * T.rows)
                needH = std::max(needH, yo[std::min(y0 + 4 * T.rows - 1, dh - 1)] + 2 - yo[y0]);
            T.boxW = needW <= 192 ? 192 : 256;
            T.boxH = needH;
            // resize8_kernel: pixels 0..3 of a group of eight tap bytes [d, d+1] of the shifted word pair (0, 1), pixels 4..7
            // of the pair (1, 2): d + 1 <= 7 resp. 4 <= d and d + 1 <= 11, with d = xofs[x] - xofs[group start]; and the
            // 128-column half must fit the 192-byte box.  The TMA box start of a half is its first column's source offset
            // rounded down to 16, so the two boxes of a CTA are encoded with the same map.
            bool wide = T.boxW == 192;
            for (int x8 = 0; x8 < dw && wide; x8 += 8) {
                for (int k = 0; k < 8 && x8 + k < dw; ++k) {
                    const int d = xo[x8 + k] - xo[x8];
                    if (k < 4 ? d + 1 > 7 : (d < 4 || d + 1 > 11)) wide = false;
                }
                if ((xo[x8] & ~3) + 16 - (xo[x8 & ~127] & ~15) > 192) wide = false;      // the four-word window inside the half's box
            }
            T.wide = wide ? 1 : 0;
            if (needW > 256 || needH > 256) return ctx->fail(NAV24_E_GEOMETRY, "scale factor too large for the resize tile");
        }
        std::vector<FastSeg> segs;
        build_fast_segments(g, segs);
        CK(ctx->bSegs.ensure(segs.size() * sizeof(FastSeg)));
        CK(cudaMemcpy(ctx->bSegs.ptr, segs.data(), segs.size() * sizeof(FastSeg), cudaMemcpyHostToDevice));
        ctx->wsW = w; ctx->wsH = h; ctx->wsFeat = ctx->prm.n_features; ctx->wsB = 0; ctx->mapsB = 0;
        ctx->lastValid = false;
    }
    if (B > ctx->wsB || shapeChanged) {
        const FrameGeom& g = ctx->g;
        const size_t b = (size_t)std::max(B, ctx->wsB);
        ctx->l0Pitch = align_up(w, 128);
        CK(ctx->bL0.ensure(b * (size_t)ctx->l0Pitch * h));
        CK(ctx->bPyr.ensure(b * (size_t)g.pyrFrameBytes + 256));
        CK(ctx->bBlur.ensure(b * (size_t)g.blurFrameBytes));
        CK(ctx->bCell.ensure(b * g.totalCells * sizeof(uint2)));
        CK(ctx->bCellDst.ensure(b * g.totalCells * sizeof(int)));
        CK(ctx->bRawCount.ensure(b * g.nlevels * sizeof(int)));
        CK(ctx->bRaw.ensure(b * g.rawPerFrame * sizeof(RawRec)));
        CK(ctx->bKeys.ensure(b * g.rawPerFrame * sizeof(RawRec)));
        CK(ctx->bNodeOfKey.ensure(b * g.rawPerFrame * sizeof(int)));
        CK(ctx->bNodesA.ensure(b * g.nodesPerFrame * sizeof(QNode)));
        CK(ctx->bNodesB.ensure(b * g.nodesPerFrame * sizeof(QNode)));
        CK(ctx->bChild.ensure(b * g.nodesPerFrame * 4 * sizeof(int)));
        CK(ctx->bAux.ensure(b * g.nodesPerFrame * sizeof(int)));
        CK(ctx->bBest.ensure(b * g.nodesPerFrame * sizeof(unsigned long long)));
        CK(ctx->bSort.ensure(b * g.nodesPerFrame * sizeof(unsigned long long)));
        CK(ctx->bLkp.ensure(b * g.kpPerFrame * sizeof(LevelKp)));
        CK(ctx->bLevelCount.ensure(b * g.nlevels * sizeof(int)));
        CK(ctx->bRawTotal.ensure(b * g.nlevels * sizeof(int)));
        CK(ctx->bFrameDone.ensure(b * sizeof(int)));
        CK(cudaMemset(ctx->bFrameDone.ptr, 0, ctx->bFrameDone.bytes));      // quadtree_kernel leaves the counters at zero
        CK(ctx->bOutKp.ensure(b * g.outCap * sizeof(nav24_kp)));
        CK(ctx->bOutDesc.ensure(b * g.outCap * 32));
        CK(ctx->bOutUd.ensure(b * g.outCap * 2 * sizeof(float)));
        CK(ctx->bNOut.ensure(b * sizeof(int)));
        CK(ctx->bMono.ensure(b * sizeof(int)));
        if (!ctx->bErr.ptr) {      // the device error word is sticky (cleared when read), so it starts at zero
            CK(ctx->bErr.ensure(sizeof(int) * 4));
            CK(cudaMemset(ctx->bErr.ptr, 0, sizeof(int) * 4));
        }
        if (!ctx->bOriTab.ptr) {      // orientation weights: umax (OP_FtDtOrbSlam.cpp:484-499) as DP4A byte weights
            static const int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
            std::vector<unsigned> t(4 * 279 * 2);
            for (int al = 0; al < 4; ++al)
                for (int r = 0; r < 31; ++r)
                    for (int c = 0; c < 9; ++c) {
                        unsigned wx = 0, wm = 0;
                        for (int b = 0; b < 4; ++b) {
                            const int u = 4 * c + b - al - 15, v = r - 15;
                            if (u >= -15 && u <= 15 && std::abs(u) <= umax[std::abs(v)]) {
                                wx |= (unsigned)(uint8_t)(int8_t)u << (8 * b);
                                wm |= 1u << (8 * b);
                            }
                        }
                        t[((al * 31 + r) * 9 + c) * 2] = wx; t[((al * 31 + r) * 9 + c) * 2 + 1] = wm;
                    }
            CK(ctx->bOriTab.ensure(t.size() * 4));
            CK(cudaMemcpy(ctx->bOriTab.ptr, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
        }
        ctx->wsB = (int)b;
        ctx->wsGen++;      // buffers moved: a captured graph holds stale pointers
        DevPtrs& p = ctx->p;
        p.pyr = (uint8_t*)ctx->bPyr.ptr; p.blur = (uint8_t*)ctx->bBlur.ptr;
        p.cellInfo = (uint2*)ctx->bCell.ptr; p.cellDst = (int*)ctx->bCellDst.ptr; p.rawCount = (int*)ctx->bRawCount.ptr;
        p.raw = (RawRec*)ctx->bRaw.ptr; p.keys = (RawRec*)ctx->bKeys.ptr; p.nodeOfKey = (int*)ctx->bNodeOfKey.ptr;
        p.nodesA = (QNode*)ctx->bNodesA.ptr; p.nodesB = (QNode*)ctx->bNodesB.ptr; p.childCnt = (int*)ctx->bChild.ptr;
        p.nodeAux = (int*)ctx->bAux.ptr; p.best = (unsigned long long*)ctx->bBest.ptr;
        p.sortRec = (unsigned long long*)ctx->bSort.ptr; p.lkp = (LevelKp*)ctx->bLkp.ptr;
        p.levelCount = (int*)ctx->bLevelCount.ptr; p.rawTotal = (int*)ctx->bRawTotal.ptr; p.frameDone = (int*)ctx->bFrameDone.ptr;
        p.outKp = (nav24_kp*)ctx->bOutKp.ptr; p.outDesc = (uint8_t*)ctx->bOutDesc.ptr; p.outUd = (float*)ctx->bOutUd.ptr;
        p.nOut = (int*)ctx->bNOut.ptr; p.monoOut = (int*)ctx->bMono.ptr; p.err = (int*)ctx->bErr.ptr;
        p.oriTab = (const unsigned*)ctx->bOriTab.ptr;
        p.segs = (const FastSeg*)ctx->bSegs.ptr;
        ctx->lastValid = false;
    }
    return NAV24_OK;
}

// pointers of the frames [f0, ...) of the workspace: every slab is [frame][...], so a chunk is an offset
DevPtrs chunk_ptrs(const nav24_orb* ctx, int f0) {
    const FrameGeom& g = ctx->g;
    DevPtrs q = ctx->p;
    const long long f = f0;
    q.l0 += f * q.l0Frame; q.pyr += f * g.pyrFrameBytes; q.blur += f * g.blurFrameBytes;
    q.cellInfo += f * g.totalCells; q.cellDst += f * g.totalCells;
    q.rawCount += f * g.nlevels; q.levelCount += f * g.nlevels; q.rawTotal += f * g.nlevels; q.frameDone += f;
    q.raw += f * g.rawPerFrame; q.keys += f * g.rawPerFrame; q.nodeOfKey += f * g.rawPerFrame;
    q.nodesA += f * g.nodesPerFrame; q.nodesB += f * g.nodesPerFrame; q.childCnt += 4 * f * g.nodesPerFrame;
    q.nodeAux += f * g.nodesPerFrame; q.best += f * g.nodesPerFrame; q.sortRec += f * g.nodesPerFrame;
    q.lkp += f * g.kpPerFrame; q.outKp += f * g.outCap; q.outDesc += f * g.outCap * 32; q.outUd += f * g.outCap * 2;
    q.nOut += f; q.monoOut += f;
    q.frameBase = f0;
    return q;
}

// (re)encode the tensor maps: levels >= 1 live in the workspace, level 0 is the caller's (or the staging) buffer
int encode_maps(nav24_orb* ctx, int B) {
    const FrameGeom& g = ctx->g;
    if (ctx->mapsB != ctx->wsB || ctx->mapsPyr != ctx->p.pyr || ctx->mapsBlurPtr != ctx->p.blur) {
        for (int l = 0; l < g.nlevels; ++l) {
            int rc = NAV24_OK;
            if (l > 0) {
                rc = encode_level_map(ctx, &ctx->maps.m[l], ctx->p.pyr + g.lv[l].off, g.lv[l].w, g.lv[l].h, ctx->wsB,
                                      g.lv[l].pitch, g.pyrFrameBytes, g.lv[l].boxW, g.lv[l].boxH);
                if (rc != NAV24_OK) return rc;
                rc = encode_level_map(ctx, &ctx->mapsOri.m[l], ctx->p.pyr + g.lv[l].off, g.lv[l].w, g.lv[l].h, ctx->wsB,
                                      g.lv[l].pitch, g.pyrFrameBytes, kOriBoxW, kOriBoxH);
                if (rc != NAV24_OK) return rc;
                rc = encode_level_map(ctx, &ctx->mapsBlurSrc.m[l], ctx->p.pyr + g.lv[l].off, g.lv[l].w, g.lv[l].h, ctx->wsB,
                                      g.lv[l].pitch, g.pyrFrameBytes, kBlurBoxW, kBlurBoxH);
                if (rc != NAV24_OK) return rc;
                if (l + 1 < g.nlevels) {
                    rc = encode_level_map(ctx, &ctx->mapsRs.m[l + 1], ctx->p.pyr + g.lv[l].off, g.lv[l].w, g.lv[l].h, ctx->wsB,
                                          g.lv[l].pitch, g.pyrFrameBytes, ctx->tabs[l + 1].boxW, ctx->tabs[l + 1].boxH);
                    if (rc != NAV24_OK) return rc;
                }
            }
            rc = encode_level_map(ctx, &ctx->mapsBlur.m[l], ctx->p.blur + g.lv[l].boff, g.lv[l].w, g.lv[l].h, ctx->wsB,
                                  g.lv[l].pitch, g.blurFrameBytes, kDescBoxW, kDescBoxH);
            if (rc != NAV24_OK) return rc;
            rc = encode_level_map(ctx, &ctx->mapsBlurN.m[l], ctx->p.blur + g.lv[l].boff, g.lv[l].w, g.lv[l].h, ctx->wsB,
                                  g.lv[l].pitch, g.blurFrameBytes, kDescBoxWN, kDescBoxH);
            if (rc != NAV24_OK) return rc;
        }
        ctx->mapsB = ctx->wsB; ctx->mapsPyr = ctx->p.pyr; ctx->mapsBlurPtr = ctx->p.blur;
    }
    const long long l0Frame = B > 1 ? ctx->p.l0Frame : ctx->p.l0Pitch * g.lv[0].h;
    int rc = encode_level_map(ctx, &ctx->mapsOri.m[0], ctx->p.l0, g.lv[0].w, g.lv[0].h, B, ctx->p.l0Pitch, l0Frame, kOriBoxW,
                              kOriBoxH);
    if (rc != NAV24_OK) return rc;
    rc = encode_level_map(ctx, &ctx->mapsBlurSrc.m[0], ctx->p.l0, g.lv[0].w, g.lv[0].h, B, ctx->p.l0Pitch, l0Frame, kBlurBoxW,
                          kBlurBoxH);
    if (rc != NAV24_OK) return rc;
    if (g.nlevels > 1) {
        rc = encode_level_map(ctx, &ctx->mapsRs.m[1], ctx->p.l0, g.lv[0].w, g.lv[0].h, B, ctx->p.l0Pitch, l0Frame,
                              ctx->tabs[1].boxW, ctx->tabs[1].boxH);
        if (rc != NAV24_OK) return rc;
    }
    return encode_level_map(ctx, &ctx->maps.m[0], ctx->p.l0, g.lv[0].w, g.lv[0].h, B, ctx->p.l0Pitch, l0Frame, g.lv[0].boxW,
                            g.lv[0].boxH);
}

// enqueue the per-frame chain for frames [f0, f0+C) on stream s; no host synchronisation.  With stages = true the
// five stage events of this run are recorded (nav24_orb_stage_ms).
int run_pipeline(nav24_orb* ctx, int f0, int C, cudaStream_t s, bool stages, int si = -1) {
    const FrameGeom& g = ctx->g;
    const DevPtrs q = chunk_ptrs(ctx, f0);
    if (stages) {
        ctx->ev = ctx->evRing[ctx->evCalls % nav24_orb::kEvRing];
        ctx->evCalls++;
        CK(cudaEventRecord(ctx->ev[0], s));
    }
    launch_fast_prepare(g, q, C, s);
    ctx->launches += launch_pyramid(g, q, ctx->tabs.data(), ctx->mapsRs, C, s);
    if (stages) CK(cudaEventRecord(ctx->ev[1], s));
    ctx->launches += launch_fast(g, q, ctx->maps, C, ctx->prm.ini_th_fast, ctx->prm.min_th_fast, s);
    if (stages) CK(cudaEventRecord(ctx->ev[2], s));
    // the blur reads only the pyramid: on the side stream of compute stream si it runs next to the quadtree (which is
    // bound by barrier and dependent-access latency); small chunks only, see nav24_orb::sstream
    const bool fork = ctx->blurFork && C <= 4 && si >= 0 && ctx->sstream[si] && !stages;
    if (fork) {
        CK(cudaEventRecord(ctx->evFork[si], s));
        CK(cudaStreamWaitEvent(ctx->sstream[si], ctx->evFork[si], 0));
        ctx->launches += launch_blur(g, q, ctx->mapsBlurSrc, C, ctx->sstream[si]);
        CK(cudaEventRecord(ctx->evBlur[si], ctx->sstream[si]));
    }
    ctx->launches += launch_quadtree(g, q, C, s);
    if (stages) CK(cudaEventRecord(ctx->ev[3], s));
    if (fork) CK(cudaStreamWaitEvent(s, ctx->evBlur[si], 0));
    else ctx->launches += launch_blur(g, q, ctx->mapsBlurSrc, C, s);
    ctx->launches += launch_describe(g, q, ctx->mapsOri, ctx->mapsBlur, ctx->mapsBlurN, C, s);
    if (ctx->cam.model != NAV24_CAM_PINHOLE)      // Calibration::undistort between detect and matchV (FE_SlamMonoV.cpp:115)
        ctx->launches += launch_undistort_frames(ctx->cam, q.outKp, q.nOut, g.outCap, C, q.outUd, s);
    if (stages) CK(cudaEventRecord(ctx->ev[4], s));
    CK(cudaGetLastError());
    return NAV24_OK;
}

int decode_device_error(nav24_orb* ctx, int e) {
    if (e & ERR_RAW_OVERFLOW) return ctx->fail(NAV24_E_OVERFLOW, "raw FAST corner buffer overflow (raise raw_keys_per_kpx)");
    if (e & ERR_ROOT_RANGE) return ctx->fail(NAV24_E_GEOMETRY, "keypoint outside the quadtree roots");
    if (e & ERR_NODE_OVERFLOW) return ctx->fail(NAV24_E_OVERFLOW, "quadtree node buffer overflow");
    if (e & ERR_CELL_SIZE) return ctx->fail(NAV24_E_GEOMETRY, "FAST cell larger than the kernel tile");
    if (e & ERR_KP_OVERFLOW) return ctx->fail(NAV24_E_OVERFLOW, "output keypoint buffer overflow");
    return NAV24_OK;
}

int check_device_error(nav24_orb* ctx) {
    int e = 0;
    CK(cudaMemcpyAsync(&e, ctx->p.err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemsetAsync(ctx->p.err, 0, sizeof(int), ctx->stream));      // sticky until read
    CK(cudaStreamSynchronize(ctx->stream));
    return decode_device_error(ctx, e);
}

// results of frames [f0, f0 + nf) of the last detect batch
int fetch_results(nav24_orb* ctx, int f0, int nf, nav24_kp* kps, uint8_t* desc, int cap, int* n_out, int* mono_out) {
    if (!ctx->lastValid) return ctx->fail(NAV24_E_BADARG, "no detect results to fetch");
    if (f0 < 0 || nf < 0 || f0 + nf > ctx->lastB) return ctx->fail(NAV24_E_BADARG, "frame range outside the last batch");
    const int B = nf;
    const FrameGeom& g = ctx->g;
    std::vector<int> n(B), m(B);
    if (B > 0) {
        CK(cudaMemcpyAsync(n.data(), ctx->p.nOut + f0, B * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(m.data(), ctx->p.monoOut + f0, B * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    }
    const int ccap = std::min(cap, g.outCap);
    if (kps && ccap > 0 && B > 0)
        CK(cudaMemcpy2DAsync(kps, (size_t)cap * sizeof(nav24_kp), ctx->p.outKp + (size_t)f0 * g.outCap, (size_t)g.outCap * sizeof(nav24_kp),
                             (size_t)ccap * sizeof(nav24_kp), B, cudaMemcpyDeviceToHost, ctx->stream));
    if (desc && ccap > 0 && B > 0)
        CK(cudaMemcpy2DAsync(desc, (size_t)cap * 32, ctx->p.outDesc + (size_t)f0 * g.outCap * 32, (size_t)g.outCap * 32, (size_t)ccap * 32, B,
                             cudaMemcpyDeviceToHost, ctx->stream));
    int rc = check_device_error(ctx);   // synchronises the stream
    if (rc != NAV24_OK) return rc;
    bool small = false;
    for (int f = 0; f < B; ++f) {
        if (n_out) n_out[f] = n[f];
        if (mono_out) mono_out[f] = m[f];
        if ((kps || desc) && n[f] > cap) small = true;
    }
    if (small) return ctx->fail(NAV24_E_CAPACITY, "output capacity too small");
    return NAV24_OK;
}

}  // namespace

// ---- matchers ---------------------------------------------------------------------------------
namespace {

int ensure_match_scratch(nav24_orb* ctx, int P, int cap, const nav24_grid_cfg* grid) {
    const size_t nCells = (size_t)grid->cols * grid->rows;
    const size_t pc = (size_t)P * cap;
    CK(ctx->mCellOf.ensure(pc * 4));
    CK(ctx->mCellStart.ensure((size_t)P * (nCells + 1) * 4));
    CK(ctx->mCellFill.ensure((size_t)P * nCells * 4));
    CK(ctx->mCellItems.ensure(pc * 4));
    CK(ctx->mCand.ensure(pc * 32 * 4));
    CK(ctx->mCandCnt.ensure(pc * 4));
    CK(ctx->mDist2.ensure(pc * 4));
    CK(ctx->mM21.ensure(pc * 4));
    CK(ctx->mBins.ensure(pc * 4));
    CK(ctx->mMatches.ensure(pc * 4));
    CK(ctx->mNMatches.ensure((size_t)P * 4));
    return NAV24_OK;
}

void fill_match_args(nav24_orb* ctx, MatchArgs& a, const nav24_grid_cfg* grid, float window, float nnratio, int th_low,
                     int check_ori, int cap) {
    a.grid = *grid;
    a.invW = (float)grid->cols / (grid->max_x - grid->min_x);      // FeatureGrid.cpp:110-111
    a.invH = (float)grid->rows / (grid->max_y - grid->min_y);
    a.window = window; a.nnratio = nnratio; a.thLow = th_low; a.checkOri = check_ori; a.cap = cap;
    a.cellOf = (int*)ctx->mCellOf.ptr; a.cellStart = (int*)ctx->mCellStart.ptr; a.cellFill = (int*)ctx->mCellFill.ptr;
    a.cellItems = (int*)ctx->mCellItems.ptr; a.cand = (int*)ctx->mCand.ptr; a.candCnt = (int*)ctx->mCandCnt.ptr; a.candCap = 32;
    a.dist2 = (int*)ctx->mDist2.ptr; a.m21 = (int*)ctx->mM21.ptr; a.bins = (int*)ctx->mBins.ptr;
    a.matches12 = (int*)ctx->mMatches.ptr; a.nMatches = (int*)ctx->mNMatches.ptr;
}

bool grid_ok(const nav24_grid_cfg* g) {
    return g && g->cols > 0 && g->rows > 0 && g->max_x > g->min_x && g->max_y > g->min_y;
}

}  // namespace

namespace {
struct MatchPlan;
int run_chunked(nav24_orb* ctx, int B, int w, int h, const uint8_t* hostGray, size_t stride, size_t frame_stride,
                const MatchPlan* mp, nav24_kp* kps, uint8_t* desc, int cap, int* n_out, int* mono_out,
                int32_t* matches12, int mcap, int* n_matches, int chunkOverride = 0, bool timedStages = false, int channels = 1);
}  // namespace

// ==========================================================================================
extern "C" {

void nav24_orb_destroy(nav24_orb* ctx);

int nav24_abi_version(void) { return NAV24_ABI_VERSION; }

int nav24_orb_create(const nav24_orb_params* params, int device, nav24_orb** out) {
    if (!params || !out) return NAV24_E_BADARG;
    *out = nullptr;
    if (params->n_levels < 1 || params->n_levels > kMaxLevels || params->n_features < 0 || params->ini_th_fast < 1 ||
        params->min_th_fast < 1 || !(params->scale_factor > 1.0f))
        return NAV24_E_BADARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return NAV24_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return NAV24_E_CUDA;
    nav24_orb* ctx = nullptr;
    const int rcCreate = guarded(nullptr, [&]() -> int {
    ctx = new nav24_orb();
    ctx->device = device;
    ctx->prm = *params;
    ctx->nIniFeatures = params->n_features;
    ctx->scaleFactorD = (double)params->scale_factor;      // the reference stores it in a double member
    const int nl = params->n_levels;
    ctx->scale.resize(nl); ctx->invScale.resize(nl); ctx->quota.resize(nl);
    ctx->scale[0] = 1.0f;
    for (int i = 1; i < nl; ++i) ctx->scale[i] = ctx->scale[i - 1] * ctx->scaleFactorD;
    for (int i = 0; i < nl; ++i) ctx->invScale[i] = 1.0f / ctx->scale[i];
    ctx->compute_quota(params->n_features);
    if (const char* e = getenv("NAV24_TRACE")) ctx->trace = atoi(e) != 0;
    if (const char* e = getenv("NAV24_GRAPH")) ctx->useGraph = atoi(e);
    if (ctx->trace) { cudaEventCreate(&ctx->evT0); cudaEventCreate(&ctx->evT1); }
    if (const char* e = getenv("NAV24_CHUNK_FRAMES")) { const int v = atoi(e); if (v > 0) ctx->chunkFrames = v; }
    if (const char* e = getenv("NAV24_STREAMS")) { const int v = atoi(e); if (v >= 1 && v <= nav24_orb::kMaxStreams) ctx->nStreams = v; }
    if (const char* e = getenv("NAV24_TAPER")) ctx->taper = atoi(e);
    if (const char* e = getenv("NAV24_RESIDENT_CHUNK")) { const int v = atoi(e); if (v > 0) ctx->residentChunk = v; }
    if (const char* e = getenv("NAV24_BLUR_FORK")) ctx->blurFork = atoi(e);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->xstream[1], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->xstream[2], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->xstream[3], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evJoinX[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evJoinX[2], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evJoinX[3], cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->outStream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evJoin, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evPrevEnd, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evPairs, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evSlotDone[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evSlotDone[1], cudaEventDisableTiming) != cudaSuccess)
        return NAV24_E_CUDA;
    ctx->xstream[0] = ctx->stream;
    for (int i = 0; i < nav24_orb::kMaxStreams; ++i)
        if (cudaStreamCreateWithFlags(&ctx->sstream[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->evFork[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->evBlur[i], cudaEventDisableTiming) != cudaSuccess)
            return NAV24_E_CUDA;
    for (auto& r : ctx->evRing) for (auto& e : r) cudaEventCreate(&e);
    for (auto& e : ctx->evT) cudaEventCreate(&e);
    return NAV24_OK;
    });
    if (rcCreate != NAV24_OK) {      // whatever was created so far goes through the one teardown path
        nav24_orb_destroy(ctx);
        return rcCreate;
    }
    *out = ctx;
    return NAV24_OK;
}

void nav24_orb_destroy(nav24_orb* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    DevBuf* bufs[] = {&ctx->bL0Tight, &ctx->bL0, &ctx->bPyr, &ctx->bBlur, &ctx->bCell, &ctx->bCellDst, &ctx->bRawCount, &ctx->bRaw, &ctx->bKeys,
                      &ctx->bNodeOfKey, &ctx->bNodesA, &ctx->bNodesB, &ctx->bChild, &ctx->bAux, &ctx->bBest, &ctx->bSort,
                      &ctx->bLkp, &ctx->bLevelCount, &ctx->bRawTotal, &ctx->bFrameDone, &ctx->bOutKp, &ctx->bOutDesc, &ctx->bNOut, &ctx->bMono,
                      &ctx->bErr, &ctx->bTabs, &ctx->bOriTab, &ctx->bSegs, &ctx->bOutUd, &ctx->bUdTmp, &ctx->mK1, &ctx->mK2, &ctx->mU1, &ctx->mU2, &ctx->mD1, &ctx->mD2, &ctx->mN1,
                      &ctx->mN2, &ctx->mCellOf, &ctx->mCellStart, &ctx->mCellFill, &ctx->mCellItems, &ctx->mCand, &ctx->mCandCnt,
                      &ctx->mDist2, &ctx->mM21, &ctx->mBins, &ctx->mMatches, &ctx->mNMatches, &ctx->mPairs, &ctx->mPairOrder, &ctx->mI0, &ctx->mI1,
                      &ctx->mF0, &ctx->mF1, &ctx->mPass};
    for (DevBuf* b : bufs) b->release();
    for (auto& r : ctx->evRing) for (auto& e : r) if (e) cudaEventDestroy(e);
    for (auto& e : ctx->evT) if (e) cudaEventDestroy(e);
    for (auto& e : ctx->evIn) cudaEventDestroy(e);
    for (auto& e : ctx->evDone) cudaEventDestroy(e);
    for (int i = 0; i < nav24_orb::kMaxStreams; ++i) {
        if (ctx->sstream[i]) cudaStreamDestroy(ctx->sstream[i]);
        if (ctx->evFork[i]) cudaEventDestroy(ctx->evFork[i]);
        if (ctx->evBlur[i]) cudaEventDestroy(ctx->evBlur[i]);
    }
    if (ctx->graphExec) cudaGraphExecDestroy(ctx->graphExec);
    if (ctx->evJoin) cudaEventDestroy(ctx->evJoin);
    if (ctx->evPrevEnd) cudaEventDestroy(ctx->evPrevEnd);
    if (ctx->evPairs) cudaEventDestroy(ctx->evPairs);
    for (auto& e : ctx->evSlotDone) if (e) cudaEventDestroy(e);
    if (ctx->evT0) cudaEventDestroy(ctx->evT0);
    if (ctx->evT1) cudaEventDestroy(ctx->evT1);
    if (ctx->hN) cudaFreeHost(ctx->hN);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    for (int i = 1; i < nav24_orb::kMaxStreams; ++i) {
        if (ctx->xstream[i]) cudaStreamDestroy(ctx->xstream[i]);
        if (ctx->evJoinX[i]) cudaEventDestroy(ctx->evJoinX[i]);
    }
    if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
    if (ctx->outStream) cudaStreamDestroy(ctx->outStream);
    delete ctx;
}

int nav24_orb_set_num_features(nav24_orb* ctx, int n) {
    if (!ctx || n < 0) return NAV24_E_BADARG;
    ctx->compute_quota(n);
    return NAV24_OK;
}
int nav24_orb_get_num_features(const nav24_orb* ctx) { return ctx ? ctx->prm.n_features : NAV24_E_BADARG; }

int nav24_orb_get_tables(const nav24_orb* ctx, float* scale, float* inv_scale, int32_t* fpl) {
    if (!ctx) return NAV24_E_BADARG;
    for (int l = 0; l < ctx->prm.n_levels; ++l) {
        if (scale) scale[l] = ctx->scale[l];
        if (inv_scale) inv_scale[l] = ctx->invScale[l];
        if (fpl) fpl[l] = ctx->quota[l];
    }
    return NAV24_OK;
}

const char* nav24_last_error_string(const nav24_orb* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int nav24_orb_max_keypoints(const nav24_orb* ctx) {
    if (!ctx) return NAV24_E_BADARG;
    // quota + 3 per level is the quadtree bound; wide images can start with 4*nIni > quota nodes
    int n = 0;
    for (int l = 0; l < ctx->prm.n_levels; ++l) n += std::max(ctx->quota[l] + 3, 64) + 4 + 4;
    if (ctx->wsFeat == ctx->prm.n_features && ctx->wsW > 0) n = std::max(n, ctx->g.outCap);
    return n;
}

int nav24_orb_detect_device(nav24_orb* ctx, const uint8_t* d_gray, int n_frames, int w, int h, size_t stride,
                            size_t frame_stride) {
    return guarded(ctx, [&]() -> int {
        if (!ctx) return NAV24_E_BADARG;
        if (!d_gray || n_frames <= 0 || w <= 0 || h <= 0 || stride < (size_t)w) return ctx->fail(NAV24_E_BADARG, "empty image");
        int rc = ensure_workspace(ctx, w, h, n_frames);
        if (rc != NAV24_OK) return rc;
        const bool aligned = (((uintptr_t)d_gray | stride | frame_stride) & 15) == 0;
        if (aligned) {
            ctx->p.l0 = d_gray; ctx->p.l0Pitch = (long long)stride; ctx->p.l0Frame = (long long)frame_stride;
        } else {
            for (int f = 0; f < n_frames; ++f)
                CK(cudaMemcpy2DAsync((uint8_t*)ctx->bL0.ptr + (size_t)f * ctx->l0Pitch * h, ctx->l0Pitch,
                                     d_gray + (size_t)f * frame_stride, stride, w, h, cudaMemcpyDeviceToDevice, ctx->stream));
            ctx->p.l0 = (const uint8_t*)ctx->bL0.ptr; ctx->p.l0Pitch = ctx->l0Pitch; ctx->p.l0Frame = (long long)ctx->l0Pitch * h;
        }
        return run_chunked(ctx, n_frames, w, h, nullptr, stride, frame_stride, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr,
                           0, nullptr, /*one chunk, one stream, stage events*/ n_frames, /*timedStages*/ true);
    });
}

int nav24_orb_fetch(nav24_orb* ctx, nav24_kp* kps, uint8_t* desc, int cap, int* n_out, int* mono_out) {
    return guarded(ctx, [&]() -> int {
        if (!ctx) return NAV24_E_BADARG;
        cudaSetDevice(ctx->device);
        return fetch_results(ctx, 0, ctx->lastB, kps, desc, cap, n_out, mono_out);
    });
}

int nav24_orb_fetch_range(nav24_orb* ctx, int first_frame, int n_frames, nav24_kp* kps, uint8_t* desc, int cap, int* n_out,
                          int* mono_out) {
    return guarded(ctx, [&]() -> int {
        if (!ctx) return NAV24_E_BADARG;
        cudaSetDevice(ctx->device);
        return fetch_results(ctx, first_frame, n_frames, kps, desc, cap, n_out, mono_out);
    });
}

int nav24_orb_sync(nav24_orb* ctx) {
    if (!ctx) return NAV24_E_BADARG;
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    return NAV24_OK;
}

}  // extern "C" (the chunked engine below is internal)

namespace {

struct MatchPlan {
    int P; const int* pairs; const nav24_grid_cfg* grid; float window, nnratio; int thLow, checkOri;
};

// The chunked engine behind every batched entry point.  The batch is cut into chunks of ctx->chunkFrames frames.
// Per chunk: [host->device copy on the copy stream] -> kernels on one of TWO compute streams (alternating, so the
// latency-bound tail of one chunk — quadtree, the sequential part of the matcher — overlaps the issue-bound head of
// the next) -> window matching of the pairs whose two frames lie in the chunk -> [device->host copy of keypoints and
// descriptors on the output stream].  Pairs spanning two chunks are matched after the streams joined.
//   hostGray != nullptr : frames come from host memory (one contiguous 1-D copy per chunk when rows are tight; the
//                         device re-pitches them) and results go back to the host; the call synchronises once.
//   hostGray == nullptr : frames are device-resident (ctx->p.l0 set by the caller); nothing is copied and the call
//                         returns after enqueueing.
int run_chunked(nav24_orb* ctx, int B, int w, int h, const uint8_t* hostGray, size_t stride, size_t frame_stride,
                const MatchPlan* mp, nav24_kp* kps, uint8_t* desc, int cap, int* n_out, int* mono_out,
                int32_t* matches12, int mcap, int* n_matches, int chunkOverride, bool timedStages, int channels) {
    const FrameGeom& g = ctx->g;
    const bool fromHost = hostGray != nullptr;
    // channels == 3: interleaved BGR frames (colour ingest), converted to grey on the device
    const bool tight = fromHost && (stride == (size_t)w * channels) && (B == 1 || frame_stride == stride * (size_t)h);
    const size_t tightFrame = (size_t)w * h * channels;
    if (fromHost && channels != 1 && !tight) return ctx->fail(NAV24_E_BADARG, "BGR frames must be tightly packed (stride = 3 * width)");
    uint8_t* l0 = (uint8_t*)ctx->bL0.ptr;
    if (fromHost) {
        if (tight) CK(ctx->bL0Tight.ensure((size_t)ctx->wsB * tightFrame + 16));
        ctx->p.l0 = l0; ctx->p.l0Pitch = ctx->l0Pitch; ctx->p.l0Frame = (long long)ctx->l0Pitch * h;
    }
    int rc = encode_maps(ctx, B);
    if (rc != NAV24_OK) return rc;
    // chunk schedule.  Host frames: the copy engine and the kernels run at about the same rate, so the call ends
    // (first copy) + (all kernels) + (last copy back): a short first chunk starts the kernels early, short last
    // chunks keep the tail after the last copy small, and the chunks in between are large enough to fill the GPU.
    const int C = std::max(1, std::min(chunkOverride > 0 ? chunkOverride : ctx->chunkFrames, B));
    std::vector<int> chunkStart;
    if (fromHost && ctx->taper && chunkOverride <= 0 && C >= 8 && B >= 3 * C) {
        const int head[2] = {C / 4, C / 2}, tail[2] = {C / 2, C / 4};
        const int rem = B - head[0] - head[1] - tail[0] - tail[1], nMid = (rem + C - 1) / C;
        int f = 0;
        for (int v : head) { chunkStart.push_back(f); f += v; }
        for (int i = 0; i < nMid; ++i) { chunkStart.push_back(f); f += rem / nMid + (i < rem % nMid ? 1 : 0); }
        for (int v : tail) { chunkStart.push_back(f); f += v; }
    } else {
        for (int f = 0; f < B; f += C) chunkStart.push_back(f);
    }
    const int nChunks = (int)chunkStart.size();
    chunkStart.push_back(B);
    std::vector<int> chunkOfFrame(B);
    for (int k = 0; k < nChunks; ++k)
        for (int f = chunkStart[k]; f < chunkStart[k + 1]; ++f) chunkOfFrame[f] = k;
    const int nS = std::max(1, std::min(ctx->nStreams, nChunks));
    CK(ctx->ensure_events(nChunks));
    CK(ctx->ensure_host(std::max(B, mp ? mp->P : 0)));
    unsigned long long sig = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { sig = (sig ^ v) * 1099511628211ull; };
    mix((unsigned long long)B); mix((unsigned long long)C); mix((unsigned long long)w); mix((unsigned long long)h);
    mix((unsigned long long)nChunks); mix((unsigned long long)nS);
    // everything that shapes the slabs or what the kernels read: workspace generation, feature count, camera, frames
    mix(ctx->wsGen); mix((unsigned long long)ctx->prm.n_features); mix((unsigned long long)(uintptr_t)ctx->p.l0);
    mix((unsigned long long)ctx->p.l0Pitch); mix((unsigned long long)ctx->p.l0Frame);
    { unsigned long long cb[5] = {0, 0, 0, 0, 0}; static_assert(sizeof(nav24_camera) <= sizeof(cb), "camera bytes"); memcpy(cb, &ctx->cam, sizeof(nav24_camera)); for (unsigned long long v : cb) mix(v); }

    // pairs grouped by chunk (pairs spanning chunks go last)
    MatchArgs ma{};
    std::vector<int> firstOfChunk(nChunks + 2, 0);
    const int P = mp ? mp->P : 0;
    if (P > 0) {
        rc = ensure_match_scratch(ctx, P, g.outCap, mp->grid);
        if (rc != NAV24_OK) return rc;
        // two slots (ping-pong) of FIXED size — half of the buffer, which only changes when it is re-allocated (cudaFree
        // synchronises the device) — so back-to-back asynchronous calls with different pair counts never overlap, and the
        // upload into a slot waits for the end of the call that used it last (two calls ago): a still-queued
        // match_window_kernel never sees its pair table overwritten.
        if ((size_t)P * 16 > ctx->mPairs.bytes || (size_t)P * 8 > ctx->mPairOrder.bytes) ctx->slotUsed[0] = ctx->slotUsed[1] = false;
        CK(ctx->mPairs.ensure((size_t)P * 16)); CK(ctx->mPairOrder.ensure((size_t)P * 8));
        ctx->callParity ^= 1;
        const int slot = ctx->callParity;
        int* dPairs = (int*)ctx->mPairs.ptr + (size_t)slot * (ctx->mPairs.bytes / 16) * 2;
        int* dOrder = (int*)ctx->mPairOrder.ptr + (size_t)slot * (ctx->mPairOrder.bytes / 8);
        if (ctx->slotUsed[slot]) CK(cudaStreamWaitEvent(ctx->copyStream, ctx->evSlotDone[slot], 0));
        std::vector<int> chunkOf(P), order(P);
        for (int q = 0; q < 2 * P; ++q) mix((unsigned long long)(unsigned)mp->pairs[q]);
        for (int q = 0; q < P; ++q) {
            const int ca = chunkOfFrame[mp->pairs[2 * q]], cb = chunkOfFrame[mp->pairs[2 * q + 1]];
            chunkOf[q] = ca == cb ? ca : nChunks;
            firstOfChunk[chunkOf[q] + 1]++;
        }
        for (int c = 0; c <= nChunks; ++c) firstOfChunk[c + 1] += firstOfChunk[c];
        std::vector<int> fill(firstOfChunk.begin(), firstOfChunk.end() - 1);
        for (int q = 0; q < P; ++q) order[fill[chunkOf[q]]++] = q;
        // (pageable sources: these two small copies return only after the data was staged, which is what we need)
        CK(cudaMemcpyAsync(dPairs, mp->pairs, (size_t)P * 8, cudaMemcpyHostToDevice, ctx->copyStream));
        CK(cudaMemcpyAsync(dOrder, order.data(), (size_t)P * 4, cudaMemcpyHostToDevice, ctx->copyStream));
        fill_match_args(ctx, ma, mp->grid, mp->window, mp->nnratio, mp->thLow, mp->checkOri, g.outCap);
        ma.k1 = ma.k2 = ctx->p.outKp; ma.d1 = ma.d2 = ctx->p.outDesc;
        ma.ud1 = ma.ud2 = ctx->cam.model != NAV24_CAM_PINHOLE ? ctx->p.outUd : nullptr;
        ma.n1 = ma.n2 = ctx->p.nOut; ma.stride1 = ma.stride2 = g.outCap;
        ma.pairs = dPairs; ma.pairOrder = dOrder;
    }
    const int nLate = P > 0 ? firstOfChunk[nChunks + 1] - firstOfChunk[nChunks] : 0;
    if (nLate > 0) mix(0x9e3779b97f4a7c15ull);
    if (fromHost) CK(cudaStreamSynchronize(ctx->stream));      // an earlier asynchronous call may still use the workspace
    // The device error word is sticky: it is read and cleared by the synchronising calls, never by an asynchronous one
    // (a memset here could race with kernels of the previous asynchronous call).
    const auto hostT0 = std::chrono::steady_clock::now();
    if (ctx->trace) CK(cudaEventRecord(ctx->evT0, ctx->copyStream));
    CK(cudaEventRecord(ctx->evPairs, ctx->copyStream));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->evPairs, 0));
    // the extra streams may run ahead into this call only if every slab is touched by the same stream as in the previous
    // call (same batch, chunking and pairs, and no pair spanning two chunks); otherwise they wait for the previous call's end
    for (int i = 1; i < nS; ++i) {
        CK(cudaStreamWaitEvent(ctx->xstream[i], ctx->evPairs, 0));
        if (ctx->prevEndValid && (sig != ctx->prevSig || nLate > 0)) CK(cudaStreamWaitEvent(ctx->xstream[i], ctx->evPrevEnd, 0));
    }
    ctx->prevSig = sig;
    const int ccap = std::min(cap, g.outCap);
    const bool stages = timedStages && nChunks == 1;      // per-stage events: nav24_orb_detect_device only, un-overlapped
    for (int k = 0; k < nChunks; ++k) {
        const int f0 = chunkStart[k], c = chunkStart[k + 1] - f0;
        cudaStream_t cs = ctx->xstream[k % nS];
        if (fromHost) {
            if (tight) {
                CK(cudaMemcpyAsync((uint8_t*)ctx->bL0Tight.ptr + f0 * tightFrame, hostGray + f0 * tightFrame, c * tightFrame,
                                   cudaMemcpyHostToDevice, ctx->copyStream));
            } else {
                for (int f = f0; f < f0 + c; ++f)
                    CK(cudaMemcpy2DAsync(l0 + (size_t)f * ctx->l0Pitch * h, ctx->l0Pitch, hostGray + (size_t)f * frame_stride,
                                         stride, w, h, cudaMemcpyHostToDevice, ctx->copyStream));
            }
            CK(cudaEventRecord(ctx->evIn[k], ctx->copyStream));
            CK(cudaStreamWaitEvent(cs, ctx->evIn[k], 0));
        }
        // one host frame, no matching: replay the captured kernel chain (re-captured when shape, feature count, camera or
        // the workspace changed)
        const bool graphed = ctx->useGraph && fromHost && B == 1 && P == 0 && !stages;
        bool replayed = false;
        if (graphed) {
            unsigned long long key = 1469598103934665603ull;
            for (unsigned long long v : {(unsigned long long)w, (unsigned long long)h, (unsigned long long)ctx->prm.n_features,
                                         ctx->wsGen, (unsigned long long)ctx->cam.model, (unsigned long long)tight + 2ull * channels,
                                         (unsigned long long)(uintptr_t)cs})
                key = (key ^ v) * 1099511628211ull;
            unsigned long long camBits = 0;
            memcpy(&camBits, &ctx->cam.fx, 8); key = (key ^ camBits) * 1099511628211ull;
            memcpy(&camBits, &ctx->cam.cx, 8); key = (key ^ camBits) * 1099511628211ull;
            memcpy(&camBits, &ctx->cam.d[0], 8); key = (key ^ camBits) * 1099511628211ull;
            memcpy(&camBits, &ctx->cam.d[2], 8); key = (key ^ camBits) * 1099511628211ull;
            if (!ctx->graphExec || key != ctx->graphKey) {
                if (ctx->graphExec) { cudaGraphExecDestroy(ctx->graphExec); ctx->graphExec = nullptr; }
                cudaGraph_t graph = nullptr;
                const long long before = ctx->launches;
                const cudaError_t eb = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
                if (eb != cudaSuccess && ctx->trace) fprintf(stderr, "[nav24 trace] begin capture failed: %s\n", cudaGetErrorString(eb));
                if (eb == cudaSuccess) {
                    if (tight)
                        ctx->launches += channels == 3 ? launch_bgr2gray((const uint8_t*)ctx->bL0Tight.ptr, w, h, l0, ctx->l0Pitch, 1, cs)
                                                       : launch_repack((const uint8_t*)ctx->bL0Tight.ptr, w, h, l0, ctx->l0Pitch, 1, cs);
                    const int prc = run_pipeline(ctx, 0, 1, cs, false, k % nS);
                    const cudaError_t ee = cudaStreamEndCapture(cs, &graph);
                    if (prc == NAV24_OK && ee == cudaSuccess && graph &&
                        cudaGraphInstantiate(&ctx->graphExec, graph, 0) == cudaSuccess) {
                        ctx->graphKey = key;
                        ctx->graphLaunches = ctx->launches - before;
                    } else {
                        ctx->graphExec = nullptr;
                        ctx->useGraph = 0;      // this driver / configuration cannot capture the chain: plain launches from now on
                        if (ctx->trace) fprintf(stderr, "[nav24 trace] graph capture failed: pipeline rc %d, end-capture %s\n", prc, cudaGetErrorString(ee));
                    }
                    if (graph) cudaGraphDestroy(graph);
                    cudaGetLastError();
                } else {
                    ctx->useGraph = 0;
                    cudaGetLastError();
                }
                ctx->launches = before;
            }
            if (ctx->graphExec) {
                if (ctx->trace) fprintf(stderr, "[nav24 trace] graph replay (%lld kernels)\n", ctx->graphLaunches);
                CK(cudaGraphLaunch(ctx->graphExec, cs));
                ctx->launches += ctx->graphLaunches;
                replayed = true;
            }
        }
        if (!replayed) {
            if (fromHost && tight) {
                const uint8_t* src = (const uint8_t*)ctx->bL0Tight.ptr + f0 * tightFrame;
                uint8_t* dst = l0 + (size_t)f0 * ctx->l0Pitch * h;
                ctx->launches += channels == 3 ? launch_bgr2gray(src, w, h, dst, ctx->l0Pitch, c, cs)
                                               : launch_repack(src, w, h, dst, ctx->l0Pitch, c, cs);
            }
            rc = run_pipeline(ctx, f0, c, cs, stages, k % nS);
            if (rc != NAV24_OK) return rc;
        }
        const int np = firstOfChunk[k + 1] - firstOfChunk[k];
        if (np > 0) {
            ma.pairBase = firstOfChunk[k];
            ctx->launches += launch_match_window(ma, np, cs);
        }
        CK(cudaEventRecord(ctx->evDone[k], cs));
        if (fromHost) {
            CK(cudaStreamWaitEvent(ctx->outStream, ctx->evDone[k], 0));
            CK(cudaMemcpyAsync(ctx->hN + f0, ctx->p.nOut + f0, c * sizeof(int), cudaMemcpyDeviceToHost, ctx->outStream));
            CK(cudaMemcpyAsync(ctx->hMono + f0, ctx->p.monoOut + f0, c * sizeof(int), cudaMemcpyDeviceToHost, ctx->outStream));
            if (kps && ccap > 0)
                CK(cudaMemcpy2DAsync(kps + (size_t)f0 * cap, (size_t)cap * sizeof(nav24_kp), ctx->p.outKp + (size_t)f0 * g.outCap,
                                     (size_t)g.outCap * sizeof(nav24_kp), (size_t)ccap * sizeof(nav24_kp), c,
                                     cudaMemcpyDeviceToHost, ctx->outStream));
            if (desc && ccap > 0)
                CK(cudaMemcpy2DAsync(desc + (size_t)f0 * cap * 32, (size_t)cap * 32, ctx->p.outDesc + (size_t)f0 * g.outCap * 32,
                                     (size_t)g.outCap * 32, (size_t)ccap * 32, c, cudaMemcpyDeviceToHost, ctx->outStream));
        }
    }
    for (int i = 1; i < nS; ++i) {      // later work on ctx->stream must see the chunks that ran on the other streams
        CK(cudaEventRecord(ctx->evJoinX[i], ctx->xstream[i]));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->evJoinX[i], 0));
    }
    if (nLate > 0) {
        ma.pairBase = firstOfChunk[nChunks];
        ctx->launches += launch_match_window(ma, nLate, ctx->stream);
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->evPrevEnd, ctx->stream));
    ctx->prevEndValid = true;
    if (P > 0) { CK(cudaEventRecord(ctx->evSlotDone[ctx->callParity], ctx->stream)); ctx->slotUsed[ctx->callParity] = true; }
    ctx->stagesValid = stages;
    ctx->lastB = B;
    ctx->lastValid = true;
    ctx->lastP = P; ctx->lastMatchCap = g.outCap;
    if (!fromHost) return NAV24_OK;

    // host mode: bring the small tails back and synchronise once
    CK(cudaStreamWaitEvent(ctx->outStream, ctx->evPrevEnd, 0));
    CK(cudaMemcpyAsync(ctx->hErr, ctx->p.err, sizeof(int), cudaMemcpyDeviceToHost, ctx->outStream));
    int* nm = ctx->hNm;      // pinned: a pageable destination would make the copy below block the host
    if (P > 0) {
        if (matches12)
            CK(cudaMemcpy2DAsync(matches12, (size_t)mcap * 4, ma.matches12, (size_t)g.outCap * 4,
                                 (size_t)std::min(mcap, g.outCap) * 4, P, cudaMemcpyDeviceToHost, ctx->outStream));
        CK(cudaMemcpyAsync(nm, ma.nMatches, (size_t)P * 4, cudaMemcpyDeviceToHost, ctx->outStream));
    }
    CK(cudaMemsetAsync(ctx->p.err, 0, sizeof(int), ctx->outStream));
    const auto hostT1 = std::chrono::steady_clock::now();
    if (ctx->trace) CK(cudaEventRecord(ctx->evT1, ctx->outStream));
    CK(cudaStreamSynchronize(ctx->outStream));
    if (ctx->trace) {
        float t = 0;
        fprintf(stderr, "[nav24 trace] enqueue %.3f ms on the host;", std::chrono::duration<double, std::milli>(hostT1 - hostT0).count());
        for (int k = 0; k < nChunks; ++k) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, ctx->evT0, ctx->evIn[k]); cudaEventElapsedTime(&b, ctx->evT0, ctx->evDone[k]);
            fprintf(stderr, " chunk %d: in %.3f done %.3f;", k, a, b);
        }
        cudaEventElapsedTime(&t, ctx->evT0, ctx->evT1);
        fprintf(stderr, " end %.3f ms\n", t);
    }
    rc = decode_device_error(ctx, *ctx->hErr);
    if (rc != NAV24_OK) { ctx->lastValid = false; return rc; }
    bool small = false;
    for (int f = 0; f < B; ++f) {
        if (n_out) n_out[f] = ctx->hN[f];
        if (mono_out) mono_out[f] = ctx->hMono[f];
        if ((kps || desc) && ctx->hN[f] > cap) small = true;
    }
    for (int q = 0; q < P; ++q) if (n_matches) n_matches[q] = nm[q];
    if (small) return ctx->fail(NAV24_E_CAPACITY, "output capacity too small");
    return NAV24_OK;
}

int check_pairs(nav24_orb* ctx, int P, const int* pairs_ab, int B, const nav24_grid_cfg* grid) {
    if (P < 0 || (P > 0 && (!pairs_ab || !grid_ok(grid)))) return ctx->fail(NAV24_E_BADARG, "bad matcher argument");
    for (int q = 0; q < 2 * P; ++q)
        if (pairs_ab[q] < 0 || pairs_ab[q] >= B) return ctx->fail(NAV24_E_BADARG, "frame index out of range");
    return NAV24_OK;
}

}  // namespace

extern "C" {

int nav24_orb_detect_batch(nav24_orb* ctx, const uint8_t* gray, int n_frames, int w, int h, size_t stride,
                           size_t frame_stride, nav24_kp* kps, uint8_t* desc, int cap, int* n_out, int* mono_out) {
    return guarded(ctx, [&]() -> int {
        if (!ctx) return NAV24_E_BADARG;
        if (!gray || n_frames <= 0 || w <= 0 || h <= 0 || stride < (size_t)w) return ctx->fail(NAV24_E_BADARG, "empty image");
        int rc = ensure_workspace(ctx, w, h, n_frames);
        if (rc != NAV24_OK) return rc;
        return run_chunked(ctx, n_frames, w, h, gray, stride, frame_stride, nullptr, kps, desc, cap, n_out, mono_out, nullptr, 0,
                           nullptr);
    });
}

int nav24_orb_detect_match_batch(nav24_orb* ctx, const uint8_t* gray, int n_frames, int w, int h, size_t stride,
                                 size_t frame_stride, nav24_kp* kps, uint8_t* desc, int cap, int* n_out, int* mono_out,
                                 int n_pairs, const int* pairs_ab, const nav24_grid_cfg* grid, float window, float nnratio,
                                 int th_low, int check_ori, int32_t* matches12, int mcap, int* n_matches) {
    return guarded(ctx, [&]() -> int {
        if (!ctx) return NAV24_E_BADARG;
        if (!gray || n_frames <= 0 || w <= 0 || h <= 0 || stride < (size_t)w) return ctx->fail(NAV24_E_BADARG, "empty image");
        int rc = check_pairs(ctx, n_pairs, pairs_ab, n_frames, grid);
        if (rc != NAV24_OK) return rc;
        rc = ensure_workspace(ctx, w, h, n_frames);
        if (rc != NAV24_OK) return rc;
        if (matches12 && mcap < ctx->g.outCap) return ctx->fail(NAV24_E_CAPACITY, "matches12 capacity below nav24_orb_max_keypoints");
        MatchPlan mp{n_pairs, pairs_ab, grid, window, nnratio, th_low, check_ori};
        return run_chunked(ctx, n_frames, w, h, gray, stride, frame_stride, n_pairs > 0 ? &mp : nullptr, kps, desc, cap, n_out,
                           mono_out, matches12, mcap, n_matches);
    });
}

int nav24_orb_detect_match_device(nav24_orb* ctx, const uint8_t* d_gray, int n_frames, int w, int h, size_t stride,
                                  size_t frame_stride, int n_pairs, const int* pairs_ab, const nav24_grid_cfg* grid,
                                  float window, float nnratio, int th_low, int check_ori) {
    return guarded(ctx, [&]() -> int {
        if (!ctx) return NAV24_E_BADARG;
        if (!d_gray || n_frames <= 0 || w <= 0 || h <= 0 || stride < (size_t)w) return ctx->fail(NAV24_E_BADARG, "empty image");
        if ((((uintptr_t)d_gray | stride | frame_stride) & 15) != 0)
            return ctx->fail(NAV24_E_BADARG, "device frames need a 16-byte aligned base, row stride and frame stride");
        int rc = check_pairs(ctx, n_pairs, pairs_ab, n_frames, grid);
        if (rc != NAV24_OK) return rc;
        rc = ensure_workspace(ctx, w, h, n_frames);
        if (rc != NAV24_OK) return rc;
        ctx->p.l0 = d_gray; ctx->p.l0Pitch = (long long)stride; ctx->p.l0Frame = (long long)frame_stride;
        MatchPlan mp{n_pairs, pairs_ab, grid, window, nnratio, th_low, check_ori};
        // Device-resident frames need no copy overlap, and cutting the batch only shortens the latency-bound launches
        // (quadtree, matcher) without making them cheaper: measured 3.24 ms/step as one chunk vs 3.47 ms in chunks of 64.
        // A large batch is still cut once, into two halves on two streams: the tails of one half's kernels are filled by the
        // other half (1024 KITTI frames: 5.99 -> 5.92 ms; three chunks 5.95, four and more slower).  NAV24_RESIDENT_CHUNK
        // overrides (a value >= the batch: one chunk).
        const int autoChunk = n_frames >= 512 ? ((n_frames + 1) / 2 + 1) & ~1 : n_frames;      // even: stereo pairs stay inside a chunk
        return run_chunked(ctx, n_frames, w, h, nullptr, stride, frame_stride, n_pairs > 0 ? &mp : nullptr, nullptr, nullptr, 0,
                           nullptr, nullptr, nullptr, 0, nullptr, ctx->residentChunk > 0 ? ctx->residentChunk : autoChunk);
    });
}

int nav24_match_fetch_range(nav24_orb* ctx, int first_pair, int n_pairs, int32_t* matches12, int mcap, int* n_matches) {
    return guarded(ctx, [&]() -> int {
        if (!ctx) return NAV24_E_BADARG;
        if (!ctx->lastValid || ctx->lastP <= 0) return ctx->fail(NAV24_E_BADARG, "no match results on the device");
        if (first_pair < 0 || n_pairs < 0 || first_pair + n_pairs > ctx->lastP) return ctx->fail(NAV24_E_BADARG, "pair range outside the last call");
        cudaSetDevice(ctx->device);
        const int P = n_pairs, oc = ctx->lastMatchCap;
        if (matches12 && mcap < oc) return ctx->fail(NAV24_E_CAPACITY, "matches12 capacity below nav24_orb_max_keypoints");
        cudaStream_t s = ctx->stream;
        if (matches12 && P > 0)
            CK(cudaMemcpy2DAsync(matches12, (size_t)mcap * 4, (const int*)ctx->mMatches.ptr + (size_t)first_pair * oc, (size_t)oc * 4,
                                 (size_t)oc * 4, P, cudaMemcpyDeviceToHost, s));
        std::vector<int> nm(P);
        if (P > 0) CK(cudaMemcpyAsync(nm.data(), (const int*)ctx->mNMatches.ptr + first_pair, (size_t)P * 4, cudaMemcpyDeviceToHost, s));
        int rc = check_device_error(ctx);      // synchronises ctx->stream (which joined the other compute streams)
        if (rc != NAV24_OK) return rc;
        int total = 0;
        for (int q = 0; q < P; ++q) { if (n_matches) n_matches[q] = nm[q]; total += nm[q]; }
        return total;
    });
}

int nav24_match_fetch(nav24_orb* ctx, int32_t* matches12, int mcap, int* n_matches) {
    if (!ctx) return NAV24_E_BADARG;
    return nav24_match_fetch_range(ctx, 0, ctx->lastP, matches12, mcap, n_matches);
}

int nav24_orb_detect(nav24_orb* ctx, const uint8_t* gray, int w, int h, size_t stride, nav24_kp* kps, uint8_t* desc,
                     int cap, int* n_out) {
    return guarded(ctx, [&]() -> int {
        int mono = 0, n = 0;
        int rc = nav24_orb_detect_batch(ctx, gray, 1, w, h, stride, stride * (size_t)(h > 0 ? h : 0), kps, desc, cap, &n, &mono);
        if (n_out) *n_out = n;
        return rc < 0 ? rc : mono;
    });
}

int nav24_orb_get_level(nav24_orb* ctx, int frame, int level, int which, uint8_t* dst, size_t dst_stride, int* w, int* h) {
    return guarded(ctx, [&]() -> int {
        if (!ctx) return NAV24_E_BADARG;
        if (!ctx->lastValid || frame < 0 || frame >= ctx->lastB || level < 0 || level >= ctx->g.nlevels)
            return ctx->fail(NAV24_E_BADARG, "bad frame/level");
        cudaSetDevice(ctx->device);
        const LevelGeom& L = ctx->g.lv[level];
        if (w) *w = L.w;
        if (h) *h = L.h;
        if (!dst) return NAV24_OK;
        const uint8_t* src; size_t pitch;
        if (which == 1) { src = ctx->p.blur + (size_t)frame * ctx->g.blurFrameBytes + L.boff; pitch = L.pitch; }
        else if (level == 0) { src = ctx->p.l0 + (size_t)frame * ctx->p.l0Frame; pitch = (size_t)ctx->p.l0Pitch; }
        else { src = ctx->p.pyr + (size_t)frame * ctx->g.pyrFrameBytes + L.off; pitch = L.pitch; }
        CK(cudaMemcpy2DAsync(dst, dst_stride, src, pitch, L.w, L.h, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return NAV24_OK;
    });
}

int nav24_orb_get_raw_keys(nav24_orb* ctx, int frame, int level, float* xyr, int cap) {
    return guarded(ctx, [&]() -> int {
        if (!ctx) return NAV24_E_BADARG;
        if (!ctx->lastValid || frame < 0 || frame >= ctx->lastB || level < 0 || level >= ctx->g.nlevels)
            return ctx->fail(NAV24_E_BADARG, "bad frame/level");
        cudaSetDevice(ctx->device);
        int n = 0;
        CK(cudaMemcpyAsync(&n, ctx->p.rawTotal + frame * ctx->g.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (!xyr || n == 0) return n;
        std::vector<RawRec> r(n);
        CK(cudaMemcpyAsync(r.data(), ctx->p.keys + (size_t)frame * ctx->g.rawPerFrame + ctx->g.lv[level].rawOff, n * sizeof(RawRec),
                           cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < n && i < cap; ++i) { xyr[3 * i] = r[i].x; xyr[3 * i + 1] = r[i].y; xyr[3 * i + 2] = r[i].score; }
        return n;
    });
}

int nav24_orb_get_level_keypoints(nav24_orb* ctx, int frame, int level, nav24_kp* kps, int cap) {
    return guarded(ctx, [&]() -> int {
        if (!ctx) return NAV24_E_BADARG;
        if (!ctx->lastValid || frame < 0 || frame >= ctx->lastB || level < 0 || level >= ctx->g.nlevels)
            return ctx->fail(NAV24_E_BADARG, "bad frame/level");
        cudaSetDevice(ctx->device);
        int n = 0;
        CK(cudaMemcpyAsync(&n, ctx->p.levelCount + frame * ctx->g.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (!kps || n == 0) return n;
        std::vector<LevelKp> r(n);
        const LevelGeom& L = ctx->g.lv[level];
        CK(cudaMemcpyAsync(r.data(), ctx->p.lkp + (size_t)frame * ctx->g.kpPerFrame + L.kpOff, n * sizeof(LevelKp),
                           cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < n && i < cap; ++i) {
            kps[i].x = r[i].x; kps[i].y = r[i].y; kps[i].size = L.patch; kps[i].angle = r[i].angle;
            kps[i].response = r[i].score; kps[i].octave = level; kps[i].class_id = -1;
        }
        return n;
    });
}

int nav24_orb_stage_ms(nav24_orb* ctx, float* ms5) {
    return guarded(ctx, [&]() -> int {
        if (!ctx || !ms5) return NAV24_E_BADARG;
        if (!ctx->lastValid || !ctx->stagesValid)
            return ctx->fail(NAV24_E_BADARG, "stage timers cover nav24_orb_detect_device only: the last call recorded none");
        cudaSetDevice(ctx->device);
        CK(cudaEventSynchronize(ctx->ev[4]));
        for (int i = 0; i < 4; ++i) CK(cudaEventElapsedTime(&ms5[i], ctx->ev[i], ctx->ev[i + 1]));
        CK(cudaEventElapsedTime(&ms5[4], ctx->ev[0], ctx->ev[4]));
        return NAV24_OK;
    });
}

int nav24_orb_stage_ms_sum(nav24_orb* ctx, float* ms5, int* calls, int reset) {
    return guarded(ctx, [&]() -> int {
        if (!ctx || !ms5) return NAV24_E_BADARG;
        cudaSetDevice(ctx->device);
        const int n = (int)std::min<long long>(ctx->evCalls, nav24_orb::kEvRing);
        for (int i = 0; i < 5; ++i) ms5[i] = 0.f;
        for (int c = 0; c < n; ++c) {
            cudaEvent_t* e = ctx->evRing[(ctx->evCalls - 1 - c) % nav24_orb::kEvRing];
            CK(cudaEventSynchronize(e[4]));
            float t;
            for (int i = 0; i < 4; ++i) { CK(cudaEventElapsedTime(&t, e[i], e[i + 1])); ms5[i] += t; }
            CK(cudaEventElapsedTime(&t, e[0], e[4])); ms5[4] += t;
        }
        if (calls) *calls = n;
        if (reset) ctx->evCalls = 0;
        return NAV24_OK;
    });
}

long long nav24_orb_launch_count(const nav24_orb* ctx) { return ctx ? ctx->launches : 0; }

int nav24_orb_timer_start(nav24_orb* ctx) {
    if (!ctx) return NAV24_E_BADARG;
    cudaSetDevice(ctx->device);
    CK(cudaEventRecord(ctx->evT[0], ctx->stream));
    return NAV24_OK;
}
int nav24_orb_timer_stop(nav24_orb* ctx, float* ms) {
    if (!ctx || !ms) return NAV24_E_BADARG;
    cudaSetDevice(ctx->device);
    CK(cudaEventRecord(ctx->evT[1], ctx->stream));
    CK(cudaEventSynchronize(ctx->evT[1]));
    CK(cudaEventElapsedTime(ms, ctx->evT[0], ctx->evT[1]));
    return NAV24_OK;
}

}  // extern "C"


extern "C" {

int nav24_match_window_batch(nav24_orb* ctx, int P, int cap, const nav24_kp* k1, const float* ud1, const uint8_t* d1,
                             const int* n1, const nav24_kp* k2, const float* ud2, const uint8_t* d2, const int* n2,
                             const nav24_grid_cfg* grid, float window, float nnratio, int th_low, int check_ori,
                             int32_t* matches12, int* n_matches) {
    return guarded(ctx, [&]() -> int {
        if (!ctx) return NAV24_E_BADARG;
        if (P <= 0 || cap <= 0 || !k1 || !k2 || !ud1 || !ud2 || !d1 || !d2 || !n1 || !n2 || !matches12 || !grid_ok(grid))
            return ctx->fail(NAV24_E_BADARG, "bad matcher argument");
        for (int p = 0; p < P; ++p)
            if (n1[p] < 0 || n1[p] > cap || n2[p] < 0 || n2[p] > cap) return ctx->fail(NAV24_E_BADARG, "n1/n2 exceed cap");
        cudaSetDevice(ctx->device);
        int rc = ensure_match_scratch(ctx, P, cap, grid);
        if (rc != NAV24_OK) return rc;
        const size_t pc = (size_t)P * cap;
        CK(ctx->mK1.ensure(pc * sizeof(nav24_kp))); CK(ctx->mK2.ensure(pc * sizeof(nav24_kp)));
        CK(ctx->mU1.ensure(pc * 8)); CK(ctx->mU2.ensure(pc * 8));
        CK(ctx->mD1.ensure(pc * 32)); CK(ctx->mD2.ensure(pc * 32));
        CK(ctx->mN1.ensure((size_t)P * 4)); CK(ctx->mN2.ensure((size_t)P * 4));
        cudaStream_t s = ctx->stream;
        CK(cudaMemcpyAsync(ctx->mK1.ptr, k1, pc * sizeof(nav24_kp), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->mK2.ptr, k2, pc * sizeof(nav24_kp), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->mU1.ptr, ud1, pc * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->mU2.ptr, ud2, pc * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->mD1.ptr, d1, pc * 32, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->mD2.ptr, d2, pc * 32, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->mN1.ptr, n1, (size_t)P * 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->mN2.ptr, n2, (size_t)P * 4, cudaMemcpyHostToDevice, s));
        MatchArgs a{};
        fill_match_args(ctx, a, grid, window, nnratio, th_low, check_ori, cap);
        a.k1 = (const nav24_kp*)ctx->mK1.ptr; a.k2 = (const nav24_kp*)ctx->mK2.ptr;
        a.ud1 = (const float*)ctx->mU1.ptr; a.ud2 = (const float*)ctx->mU2.ptr;
        a.d1 = (const uint8_t*)ctx->mD1.ptr; a.d2 = (const uint8_t*)ctx->mD2.ptr;
        a.n1 = (const int*)ctx->mN1.ptr; a.n2 = (const int*)ctx->mN2.ptr;
        a.stride1 = cap; a.stride2 = cap; a.pairs = nullptr;
        ctx->launches += launch_match_window(a, P, s);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(matches12, a.matches12, pc * 4, cudaMemcpyDeviceToHost, s));
        std::vector<int> nm(P);
        CK(cudaMemcpyAsync(nm.data(), a.nMatches, (size_t)P * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        int total = 0;
        for (int p = 0; p < P; ++p) { if (n_matches) n_matches[p] = nm[p]; total += nm[p]; }
        return total;
    });
}

int nav24_match_window(nav24_orb* ctx, const nav24_kp* k1, const float* ud1, const uint8_t* d1, int n1, const nav24_kp* k2,
                       const float* ud2, const uint8_t* d2, int n2, const nav24_grid_cfg* grid, float window, float nnratio,
                       int th_low, int check_ori, int32_t* matches12) {
    return guarded(ctx, [&]() -> int {
        if (!ctx) return NAV24_E_BADARG;
        if (n1 < 0 || n2 < 0) return ctx->fail(NAV24_E_BADARG, "negative count");
        if (n1 == 0) return 0;
        const int cap = std::max(std::max(n1, n2), 1);
        // pad both sides to a common capacity so that the batched entry point can be reused
        std::vector<nav24_kp> K1(cap), K2(cap);
        std::vector<float> U1(2 * (size_t)cap), U2(2 * (size_t)cap);
        std::vector<uint8_t> D1(32 * (size_t)cap), D2(32 * (size_t)cap);
        std::vector<int32_t> M(cap, -1);
        if (!k1 || !ud1 || !d1 || !matches12 || (n2 > 0 && (!k2 || !ud2 || !d2))) return ctx->fail(NAV24_E_BADARG, "null pointer");
        memcpy(K1.data(), k1, (size_t)n1 * sizeof(nav24_kp)); memcpy(U1.data(), ud1, (size_t)n1 * 8); memcpy(D1.data(), d1, (size_t)n1 * 32);
        if (n2) { memcpy(K2.data(), k2, (size_t)n2 * sizeof(nav24_kp)); memcpy(U2.data(), ud2, (size_t)n2 * 8); memcpy(D2.data(), d2, (size_t)n2 * 32); }
        int nm = 0;
        int rc = nav24_match_window_batch(ctx, 1, cap, K1.data(), U1.data(), D1.data(), &n1, K2.data(), U2.data(), D2.data(), &n2,
                                          grid, window, nnratio, th_low, check_ori, M.data(), &nm);
        if (rc < 0) return rc;
        memcpy(matches12, M.data(), (size_t)n1 * 4);
        return nm;
    });
}

int nav24_match_window_frames(nav24_orb* ctx, int P, const int* pairs_ab, const nav24_grid_cfg* grid, float window,
                              float nnratio, int th_low, int check_ori, int32_t* matches12, int cap, int* n_matches) {
    return guarded(ctx, [&]() -> int {
        if (!ctx) return NAV24_E_BADARG;
        if (!ctx->lastValid) return ctx->fail(NAV24_E_BADARG, "no detect results on the device");
        if (P <= 0 || !pairs_ab || !grid_ok(grid)) return ctx->fail(NAV24_E_BADARG, "bad matcher argument");
        const int oc = ctx->g.outCap;
        if (matches12 && cap < oc) return ctx->fail(NAV24_E_CAPACITY, "matches12 capacity below nav24_orb_max_keypoints");
        for (int p = 0; p < 2 * P; ++p)
            if (pairs_ab[p] < 0 || pairs_ab[p] >= ctx->lastB) return ctx->fail(NAV24_E_BADARG, "frame index out of range");
        cudaSetDevice(ctx->device);
        int rc = ensure_match_scratch(ctx, P, oc, grid);
        if (rc != NAV24_OK) return rc;
        CK(ctx->mPairs.ensure((size_t)P * 8));
        cudaStream_t s = ctx->stream;
        CK(cudaMemcpyAsync(ctx->mPairs.ptr, pairs_ab, (size_t)P * 8, cudaMemcpyHostToDevice, s));
        MatchArgs a{};
        fill_match_args(ctx, a, grid, window, nnratio, th_low, check_ori, oc);
        a.k1 = a.k2 = ctx->p.outKp; a.d1 = a.d2 = ctx->p.outDesc;
        a.ud1 = a.ud2 = ctx->cam.model != NAV24_CAM_PINHOLE ? ctx->p.outUd : nullptr;
        a.n1 = a.n2 = ctx->p.nOut; a.stride1 = a.stride2 = oc; a.pairs = (const int*)ctx->mPairs.ptr;
        ctx->launches += launch_match_window(a, P, s);
        CK(cudaGetLastError());
        if (!matches12 && !n_matches) return NAV24_OK;      // fully asynchronous: results stay on the device
        if (matches12)
            CK(cudaMemcpy2DAsync(matches12, (size_t)cap * 4, a.matches12, (size_t)oc * 4, (size_t)oc * 4, P, cudaMemcpyDeviceToHost, s));
        std::vector<int> nm(P);
        CK(cudaMemcpyAsync(nm.data(), a.nMatches, (size_t)P * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        int total = 0;
        for (int p = 0; p < P; ++p) { if (n_matches) n_matches[p] = nm[p]; total += nm[p]; }
        return total;
    });
}

static bool camera_ok(const nav24_camera* c) {
    return c && c->model >= NAV24_CAM_PINHOLE && c->model <= NAV24_CAM_KB8 && (c->model == NAV24_CAM_PINHOLE || (c->fx != 0.f && c->fy != 0.f));
}

int nav24_undistort_points(nav24_orb* ctx, const nav24_camera* cam, const float* xy, int n, float* ud_xy) {
    return guarded(ctx, [&]() -> int {
        if (!ctx) return NAV24_E_BADARG;
        if (!camera_ok(cam) || n < 0 || (n > 0 && (!xy || !ud_xy))) return ctx->fail(NAV24_E_BADARG, "bad undistort argument");
        if (n == 0) return NAV24_OK;
        cudaSetDevice(ctx->device);
        CK(ctx->bUdTmp.ensure((size_t)n * 16));
        float* dIn = (float*)ctx->bUdTmp.ptr; float* dOut = dIn + 2 * (size_t)n;
        cudaStream_t s = ctx->stream;
        CK(cudaMemcpyAsync(dIn, xy, (size_t)n * 8, cudaMemcpyHostToDevice, s));
        ctx->launches += launch_undistort_points(*cam, dIn, n, dOut, s);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ud_xy, dOut, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        return NAV24_OK;
    });
}

int nav24_orb_set_camera(nav24_orb* ctx, const nav24_camera* cam) {
    if (!ctx) return NAV24_E_BADARG;
    if (!cam) { ctx->cam = nav24_camera{}; return NAV24_OK; }
    if (!camera_ok(cam)) return ctx->fail(NAV24_E_BADARG, "bad camera model");
    ctx->cam = *cam;
    ctx->lastValid = false;      // coordinates on the device belong to the previous camera
    return NAV24_OK;
}

int nav24_orb_fetch_undistorted(nav24_orb* ctx, float* ud_xy, int cap) {
    return guarded(ctx, [&]() -> int {
        if (!ctx || !ud_xy) return NAV24_E_BADARG;
        if (!ctx->lastValid) return ctx->fail(NAV24_E_BADARG, "no detect results on the device");
        cudaSetDevice(ctx->device);
        int rc = check_device_error(ctx);      // synchronises ctx->stream (which joined the other compute streams)
        if (rc != NAV24_OK) return rc;
        const int oc = ctx->g.outCap, c = std::min(cap, oc), B = ctx->lastB;
        if (ctx->cam.model != NAV24_CAM_PINHOLE) {
            CK(cudaMemcpy2D(ud_xy, (size_t)cap * 8, ctx->p.outUd, (size_t)oc * 8, (size_t)c * 8, B, cudaMemcpyDeviceToHost));
        } else {                             // identity: the detected coordinates
            std::vector<nav24_kp> k((size_t)B * oc);
            CK(cudaMemcpy(k.data(), ctx->p.outKp, k.size() * sizeof(nav24_kp), cudaMemcpyDeviceToHost));
            std::vector<int> n(B);
            CK(cudaMemcpy(n.data(), ctx->p.nOut, (size_t)B * 4, cudaMemcpyDeviceToHost));
            for (int f = 0; f < B; ++f)
                for (int i = 0; i < std::min(n[f], c); ++i) {
                    ud_xy[((size_t)f * cap + i) * 2] = k[(size_t)f * oc + i].x;
                    ud_xy[((size_t)f * cap + i) * 2 + 1] = k[(size_t)f * oc + i].y;
                }
        }
        return NAV24_OK;
    });
}

int nav24_match_bf_knn2(nav24_orb* ctx, const uint8_t* d1, int n1, const uint8_t* d2, int n2, int norm, float ratio,
                        int32_t* idx0, int32_t* idx1, float* dist0, float* dist1, uint8_t* pass) {
    return guarded(ctx, [&]() -> int {
        if (!ctx) return NAV24_E_BADARG;
        if (n1 < 0 || n2 < 0 || (norm != NAV24_NORM_HAMMING && norm != NAV24_NORM_L2_U8)) return ctx->fail(NAV24_E_BADARG, "bad argument");
        if (n1 == 0) return 0;
        if (!d1 || (n2 > 0 && !d2) || !idx0 || !idx1 || !dist0 || !dist1 || !pass) return ctx->fail(NAV24_E_BADARG, "null pointer");
        cudaSetDevice(ctx->device);
        CK(ctx->mD1.ensure((size_t)n1 * 32)); CK(ctx->mD2.ensure((size_t)std::max(n2, 1) * 32));
        CK(ctx->mI0.ensure((size_t)n1 * 4)); CK(ctx->mI1.ensure((size_t)n1 * 4));
        CK(ctx->mF0.ensure((size_t)n1 * 4)); CK(ctx->mF1.ensure((size_t)n1 * 4)); CK(ctx->mPass.ensure((size_t)n1));
        cudaStream_t s = ctx->stream;
        CK(cudaMemcpyAsync(ctx->mD1.ptr, d1, (size_t)n1 * 32, cudaMemcpyHostToDevice, s));
        if (n2) CK(cudaMemcpyAsync(ctx->mD2.ptr, d2, (size_t)n2 * 32, cudaMemcpyHostToDevice, s));
        CK(ctx->bUdTmp.ensure((size_t)bf_knn2_segments(n1, n2) * n1 * sizeof(int4)));      // (scratch shared with nav24_undistort_points)
        CK(cudaEventRecord(ctx->evT[0], s));
        ctx->launches += launch_bf_knn2((const uint8_t*)ctx->mD1.ptr, n1, (const uint8_t*)ctx->mD2.ptr, n2, norm, ratio,
                                        (int*)ctx->mI0.ptr, (int*)ctx->mI1.ptr, (float*)ctx->mF0.ptr, (float*)ctx->mF1.ptr,
                                        (uint8_t*)ctx->mPass.ptr, (int4*)ctx->bUdTmp.ptr, s);
        CK(cudaEventRecord(ctx->evT[1], s));      // nav24_debug_last_kernel_ms: device time of the two kernels
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(idx0, ctx->mI0.ptr, (size_t)n1 * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(idx1, ctx->mI1.ptr, (size_t)n1 * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(dist0, ctx->mF0.ptr, (size_t)n1 * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(dist1, ctx->mF1.ptr, (size_t)n1 * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(pass, ctx->mPass.ptr, (size_t)n1, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        int np = 0;
        for (int i = 0; i < n1; ++i) np += pass[i];
        return np;
    });
}

int nav24_debug_last_kernel_ms(nav24_orb* ctx, float* ms) {
    if (!ctx || !ms) return NAV24_E_BADARG;
    cudaSetDevice(ctx->device);
    CK(cudaEventSynchronize(ctx->evT[1]));
    CK(cudaEventElapsedTime(ms, ctx->evT[0], ctx->evT[1]));
    return NAV24_OK;
}

int nav24_debug_sort_u32(nav24_orb* ctx, const uint32_t* keys, int n, int32_t* perm) {
    return guarded(ctx, [&]() -> int {
        if (!ctx) return NAV24_E_BADARG;
        if (!keys || !perm || n < 0) return ctx->fail(NAV24_E_BADARG, "bad argument");
        if (n == 0) return NAV24_OK;
        cudaSetDevice(ctx->device);
        std::vector<unsigned long long> r(n);
        for (int i = 0; i < n; ++i) r[i] = ((unsigned long long)keys[i] << 32) | (unsigned)i;
        CK(ctx->mCand.ensure((size_t)n * 8));
        CK(cudaMemcpyAsync(ctx->mCand.ptr, r.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (launch_debug_sort((unsigned long long*)ctx->mCand.ptr, n, ctx->stream) < 0) return ctx->fail(NAV24_E_CAPACITY, "too many records");
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(r.data(), ctx->mCand.ptr, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < n; ++i) perm[i] = (int32_t)(r[i] & 0xffffffffu);
        return NAV24_OK;
    });
}

int nav24_host_alloc(size_t bytes, void** out) {
    if (!out) return NAV24_E_BADARG;
    return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? NAV24_OK : NAV24_E_NOMEM;
}
int nav24_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? NAV24_OK : NAV24_E_CUDA; }
int nav24_device_alloc(size_t bytes, void** out) {
    if (!out) return NAV24_E_BADARG;
    return cudaMalloc(out, bytes) == cudaSuccess ? NAV24_OK : NAV24_E_NOMEM;
}
int nav24_device_free(void* p) { return cudaFree(p) == cudaSuccess ? NAV24_OK : NAV24_E_CUDA; }
int nav24_memcpy_h2d(void* dst, const void* src, size_t bytes) {
    return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess ? NAV24_OK : NAV24_E_CUDA;
}

}  // extern "C"

// ---- image ingest ---------------------------------------------------------------------------------
struct nav24_ingest {
    nav24_orb* ctx = nullptr;
    int w = 0, h = 0, ch = 1, slots = 0;
    size_t slotBytes = 0;
    uint8_t* host = nullptr;      // pinned: [slots][h][w * ch]
};

extern "C" {

int nav24_ingest_create(nav24_orb* ctx, int width, int height, int channels, int n_slots, nav24_ingest** out) {
    return guarded(ctx, [&]() -> int {
        if (!ctx || !out) return NAV24_E_BADARG;
        *out = nullptr;
        if (width <= 0 || height <= 0 || (channels != 1 && channels != 3) || n_slots <= 0)
            return ctx->fail(NAV24_E_BADARG, "ingest ring: width, height > 0, channels 1 (grey) or 3 (BGR), slots > 0");
        cudaSetDevice(ctx->device);
        nav24_ingest* r = new nav24_ingest();
        r->ctx = ctx; r->w = width; r->h = height; r->ch = channels; r->slots = n_slots;
        r->slotBytes = (size_t)width * height * channels;
        if (cudaHostAlloc((void**)&r->host, r->slotBytes * n_slots + 16, cudaHostAllocDefault) != cudaSuccess) {
            delete r;
            return ctx->fail(NAV24_E_NOMEM, "ingest ring: pinned allocation failed");
        }
        *out = r;
        return NAV24_OK;
    });
}

void nav24_ingest_destroy(nav24_ingest* ring) {
    if (!ring) return;
    if (ring->host) cudaFreeHost(ring->host);
    delete ring;
}

uint8_t* nav24_ingest_slot(nav24_ingest* ring, int slot) {
    return (ring && slot >= 0 && slot < ring->slots) ? ring->host + (size_t)slot * ring->slotBytes : nullptr;
}
size_t nav24_ingest_slot_bytes(const nav24_ingest* ring) { return ring ? ring->slotBytes : 0; }

int nav24_ingest_detect_match(nav24_ingest* ring, int first_slot, int n_frames, nav24_kp* kps, uint8_t* desc, int cap, int* n_out,
                              int* mono_out, int n_pairs, const int* pairs_ab, const nav24_grid_cfg* grid, float window,
                              float nnratio, int th_low, int check_ori, int32_t* matches12, int mcap, int* n_matches) {
    nav24_orb* ctx = ring ? ring->ctx : nullptr;
    return guarded(ctx, [&]() -> int {
        if (!ring || !ctx) return NAV24_E_BADARG;
        if (first_slot < 0 || n_frames <= 0 || first_slot + n_frames > ring->slots)
            return ctx->fail(NAV24_E_BADARG, "ingest: slot range outside the ring");
        int rc = check_pairs(ctx, n_pairs, pairs_ab, n_frames, grid);
        if (rc != NAV24_OK) return rc;
        rc = ensure_workspace(ctx, ring->w, ring->h, n_frames);
        if (rc != NAV24_OK) return rc;
        if (matches12 && mcap < ctx->g.outCap) return ctx->fail(NAV24_E_CAPACITY, "matches12 capacity below nav24_orb_max_keypoints");
        MatchPlan mp{n_pairs, pairs_ab, grid, window, nnratio, th_low, check_ori};
        const size_t stride = (size_t)ring->w * ring->ch;
        return run_chunked(ctx, n_frames, ring->w, ring->h, ring->host + (size_t)first_slot * ring->slotBytes, stride,
                           stride * ring->h, n_pairs > 0 ? &mp : nullptr, kps, desc, cap, n_out, mono_out, matches12, mcap,
                           n_matches, 0, false, ring->ch);
    });
}

int nav24_ingest_detect(nav24_ingest* ring, int first_slot, int n_frames, nav24_kp* kps, uint8_t* desc, int cap, int* n_out,
                        int* mono_out) {
    return nav24_ingest_detect_match(ring, first_slot, n_frames, kps, desc, cap, n_out, mono_out, 0, nullptr, nullptr, 0.f, 0.f, 0, 0,
                                     nullptr, 0, nullptr);
}

}  // extern "C"

// ---- two-view RANSAC scoring -------------------------------------------------------------------------
extern "C" {

// Shared body of nav24_two_view_score / nav24_two_view_score_kept.  all_h / all_f: n_hyp x n_matches masks of every iteration
// (may be NULL); kept_h / kept_f: n_matches bytes, the mask of the kept iteration only (may be NULL).
static int two_view_run(nav24_orb* ctx, const float* xy1, const float* xy2, int n_matches, const float* H21, const float* H12,
                        const float* F21, int n_hyp, float sigma, float th_h, float th_f, float th_score, float* score_h,
                        float* score_f, uint8_t* all_h, uint8_t* all_f, uint8_t* kept_h, uint8_t* kept_f, int* best_h, int* best_f) {
    return guarded(ctx, [&]() -> int {
        if (!ctx) return NAV24_E_BADARG;
        if (n_matches < 0 || n_hyp < 0 || !(sigma > 0.f) || (n_matches > 0 && (!xy1 || !xy2)) || (H21 && !H12) || (H21 && !score_h) ||
            (F21 && !score_f) || (!H21 && !F21))
            return ctx->fail(NAV24_E_BADARG, "bad two-view scoring argument");
        if (best_h) *best_h = -1;
        if (best_f) *best_f = -1;
        if (kept_h && n_matches > 0) memset(kept_h, 0, (size_t)n_matches);
        if (kept_f && n_matches > 0) memset(kept_f, 0, (size_t)n_matches);
        if (n_hyp == 0) return NAV24_OK;
        cudaSetDevice(ctx->device);
        const size_t nb = (size_t)std::max(n_matches, 1), hb = (size_t)n_hyp;
        // device scratch (shared with the undistortion helpers): points | matrices | scores | inlier masks
        const size_t oXy = 0, oM = oXy + 2 * nb * 8, oS = oM + 3 * hb * 36, oI = oS + 2 * hb * 4;
        CK(ctx->bUdTmp.ensure(oI + 2 * hb * nb + 64));
        char* d = (char*)ctx->bUdTmp.ptr;
        float* dXy1 = (float*)(d + oXy); float* dXy2 = dXy1 + 2 * nb;
        float* dH21 = (float*)(d + oM); float* dH12 = dH21 + 9 * hb; float* dF21 = dH12 + 9 * hb;
        float* dSH = (float*)(d + oS); float* dSF = dSH + hb;
        uint8_t* dIH = (uint8_t*)(d + oI); uint8_t* dIF = dIH + hb * nb;
        cudaStream_t s = ctx->stream;
        if (n_matches > 0) {
            CK(cudaMemcpyAsync(dXy1, xy1, (size_t)n_matches * 8, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(dXy2, xy2, (size_t)n_matches * 8, cudaMemcpyHostToDevice, s));
        }
        if (H21) {
            CK(cudaMemcpyAsync(dH21, H21, hb * 36, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(dH12, H12, hb * 36, cudaMemcpyHostToDevice, s));
        }
        if (F21) CK(cudaMemcpyAsync(dF21, F21, hb * 36, cudaMemcpyHostToDevice, s));
        const bool maskH = H21 && (all_h || kept_h), maskF = F21 && (all_f || kept_f);
        if (n_matches > 0) {
            ctx->launches += launch_two_view_score(dXy1, dXy2, n_matches, H21 ? dH21 : nullptr, H21 ? dH12 : nullptr, F21 ? dF21 : nullptr,
                                                   n_hyp, sigma, th_h, th_f, th_score, dSH, dSF, maskH ? dIH : nullptr,
                                                   maskF ? dIF : nullptr, s);
            CK(cudaGetLastError());
        } else {
            CK(cudaMemsetAsync(dSH, 0, 2 * hb * 4, s));      // no matches: every score is 0 (the reference's loops do not run)
        }
        if (H21) CK(cudaMemcpyAsync(score_h, dSH, hb * 4, cudaMemcpyDeviceToHost, s));
        if (F21) CK(cudaMemcpyAsync(score_f, dSF, hb * 4, cudaMemcpyDeviceToHost, s));
        if (H21 && all_h && n_matches > 0) CK(cudaMemcpyAsync(all_h, dIH, hb * n_matches, cudaMemcpyDeviceToHost, s));
        if (F21 && all_f && n_matches > 0) CK(cudaMemcpyAsync(all_f, dIF, hb * n_matches, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        // `if (currentScore > score) keep` over the iterations (:307-312, :358-363): the first hypothesis with the highest
        // score, and none when no score exceeds the initial 0
        auto pick = [&](const float* sc) { int b = -1; float best = 0.f; for (int i = 0; i < n_hyp; ++i) if (sc[i] > best) { best = sc[i]; b = i; } return b; };
        const int bh = H21 ? pick(score_h) : -1, bf = F21 ? pick(score_f) : -1;
        if (best_h) *best_h = bh;
        if (best_f) *best_f = bf;
        // the kept iteration's mask only: n_matches bytes per model instead of n_hyp x n_matches
        bool more = false;
        if (kept_h && bh >= 0 && n_matches > 0) { CK(cudaMemcpyAsync(kept_h, dIH + (size_t)bh * n_matches, (size_t)n_matches, cudaMemcpyDeviceToHost, s)); more = true; }
        if (kept_f && bf >= 0 && n_matches > 0) { CK(cudaMemcpyAsync(kept_f, dIF + (size_t)bf * n_matches, (size_t)n_matches, cudaMemcpyDeviceToHost, s)); more = true; }
        if (more) CK(cudaStreamSynchronize(s));
        return NAV24_OK;
    });
}

int nav24_two_view_score(nav24_orb* ctx, const float* xy1, const float* xy2, int n_matches, const float* H21, const float* H12,
                         const float* F21, int n_hyp, float sigma, float th_h, float th_f, float th_score, float* score_h,
                         float* score_f, uint8_t* inliers_h, uint8_t* inliers_f, int* best_h, int* best_f) {
    return two_view_run(ctx, xy1, xy2, n_matches, H21, H12, F21, n_hyp, sigma, th_h, th_f, th_score, score_h, score_f, inliers_h,
                        inliers_f, nullptr, nullptr, best_h, best_f);
}

int nav24_two_view_score_kept(nav24_orb* ctx, const float* xy1, const float* xy2, int n_matches, const float* H21, const float* H12,
                              const float* F21, int n_hyp, float sigma, float th_h, float th_f, float th_score, float* score_h,
                              float* score_f, uint8_t* kept_inliers_h, uint8_t* kept_inliers_f, int* best_h, int* best_f) {
    return two_view_run(ctx, xy1, xy2, n_matches, H21, H12, F21, n_hyp, sigma, th_h, th_f, th_score, score_h, score_f, nullptr,
                        nullptr, kept_inliers_h, kept_inliers_f, best_h, best_f);
}

}  // extern "C"
