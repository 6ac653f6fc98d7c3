// orb_kernels.cu — sm_100a kernels of the ORB detector (pyramid, FAST cells, quadtree, order,
// blur, orientation + rBRIEF).  Each kernel cites the reference lines whose results it reproduces
// (paths relative to /root/reference/core/operators/objDetection/).  Integer work is bit-exact;
// float work uses explicit round-to-nearest intrinsics so that nvcc cannot contract mul+add
// into FMA (the reference is built without contraction).
#include <algorithm>
#include <cstdlib>

#include "orb_internal.cuh"
#include "stdsort_warp.cuh"

// minimum-blocks argument of the launch bounds of the image kernels (0 = none: ptxas picks the register count for the plain
// thread bound); -DNAV24_*_MINB=n for A/B builds
#ifndef NAV24_FAST_MINB
#define NAV24_FAST_MINB 0
#endif
#ifndef NAV24_BLUR_MINB
#define NAV24_BLUR_MINB 0
#endif
#ifndef NAV24_RS_SPLIT
#define NAV24_RS_SPLIT 1
#endif
#ifndef NAV24_PDL
#define NAV24_PDL 1
#endif
#ifndef NAV24_RS_MINB
#define NAV24_RS_MINB 1      // resize: 62 registers instead of 40 (the loads of a source row are issued further ahead): 0.79 -> 0.75 ms
#endif
#if NAV24_FAST_MINB > 0
#define NAV24_FAST_LB __launch_bounds__(kFastThreads, NAV24_FAST_MINB)
#else
#define NAV24_FAST_LB __launch_bounds__(kFastThreads)
#endif
#if NAV24_BLUR_MINB > 0
#define NAV24_BLUR_LB __launch_bounds__(128, NAV24_BLUR_MINB)
#else
#define NAV24_BLUR_LB __launch_bounds__(128)
#endif
#if NAV24_RS_MINB > 0
#define NAV24_RS_LB __launch_bounds__(128, NAV24_RS_MINB)
#else
#define NAV24_RS_LB __launch_bounds__(128)
#endif

namespace nav24 {

namespace {

__device__ __forceinline__ const uint8_t* level_ptr(const FrameGeom& g, const DevPtrs& p, int f, int l) {
    return l == 0 ? p.l0 + (long long)f * p.l0Frame : p.pyr + (long long)f * g.pyrFrameBytes + g.lv[l].off;
}
__device__ __forceinline__ long long level_pitch(const FrameGeom& g, const DevPtrs& p, int l) {
    return l == 0 ? p.l0Pitch : (long long)g.lv[l].pitch;
}

// exclusive block scan of one int per thread; returns the exclusive prefix, *total = block sum.
// All threads of the block must call it. blockDim.x must be a multiple of 32, <= 1024.
// Fewest instructions (warp 0 alone scans the warp sums), three barriers: the form for the issue-bound FAST kernel
// (the one-barrier form below made it 1.5 % slower).
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* s_warp /* >= 33 ints */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();   // protect s_warp from the previous call's readers
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = lane < nw ? s_warp[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < nw) s_warp[lane] = wi - w;
        if (lane == nw - 1) s_warp[32] = wi;
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[wid] + incl - v;
}

// The same scan with ONE barrier per call, for the kernels that live on barrier latency (quadtree, order: 0.624 -> 0.606 ms): every warp leaves its sum in shared memory, and after the barrier every warp scans the (<= 32)
// warp sums itself.  The sums alternate between two buffers, so a warp that is already in the next call cannot overwrite
// what a slower warp still reads (it cannot be two calls ahead: the next call's barrier is in between).
struct BlockScan { int* buf; int par; };      // buf: 2 x 32 ints of shared memory; par = 0 at kernel start
__device__ __forceinline__ int block_excl_scan(int v, int* total, BlockScan& bs) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    int* w = bs.buf + 32 * bs.par;
    bs.par ^= 1;
    if (lane == 31) w[wid] = incl;
    __syncthreads();
    const int ws = lane < nw ? w[lane] : 0;
    int wi = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
    }
    *total = __shfl_sync(0xffffffffu, wi, nw - 1);
    return __shfl_sync(0xffffffffu, wi - ws, wid) + incl - v;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------
// K0  repack: frames arrive from the host as ONE contiguous copy (a strided cudaMemcpy2D of 1241-byte rows
// runs at 10 GB/s, the contiguous copy at 55 GB/s) and are re-pitched on the device to the 16-byte
// aligned row pitch TMA needs.  One thread = 16 destination bytes; unaligned source words are stitched
// with funnel shifts.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) repack_kernel(const uint8_t* __restrict__ src, int w, int h, uint8_t* __restrict__ dst,
                                                     int dPitch) {
    const int x16 = (blockIdx.x * 64 + threadIdx.x) * 16;
    const int y = blockIdx.y * 4 + threadIdx.y;
    if (x16 >= dPitch || y >= h) return;
    const long long f = blockIdx.z;
    const uint8_t* row = src + (f * h + y) * (long long)w;
    const uint8_t* rowEnd = row + w;
    const uint8_t* s0 = row + x16;
    const unsigned sh = ((unsigned)(uintptr_t)s0 & 3u) * 8u;
    const unsigned* a = reinterpret_cast<const unsigned*>((uintptr_t)s0 & ~(uintptr_t)3);
    unsigned wv[5];      // the staging buffer has 16 bytes of slack, so the last row may be over-read
#pragma unroll
    for (int k = 0; k < 5; ++k) wv[k] = (reinterpret_cast<const uint8_t*>(a + k) < rowEnd) ? __ldg(a + k) : 0u;
    uint4 o;
    o.x = __funnelshift_r(wv[0], wv[1], sh); o.y = __funnelshift_r(wv[1], wv[2], sh);
    o.z = __funnelshift_r(wv[2], wv[3], sh); o.w = __funnelshift_r(wv[3], wv[4], sh);
    *reinterpret_cast<uint4*>(dst + (f * h + y) * (long long)dPitch + x16) = o;
}

// ------------------------------------------------------------------------------------------
// K0b  colour ingest: interleaved 8-bit BGR frames, as cv::imread / cv::imdecode leave them and as they
// arrive from the host in ONE contiguous copy, -> grey level 0 at the 16-byte aligned pitch.  Replaces the
// cv::cvtColor(img, gray, COLOR_BGR2GRAY) of FE_SlamMonoV.cpp:92-94 (whose result the reference then forgets to hand to
// the detector, which asserts CV_8UC1 at OP_FtDtOrbSlam.cpp:853) with OpenCV's 15-bit fixed point
//   grey = (3735 B + 19235 G + 9798 R + 16384) >> 15          (pinned against cv2 4.13, tests/test_ingest.py).
// One thread = 4 pixels = 12 source bytes: the four aligned words covering them are shifted to byte 0 (3 SHF), every
// pixel's (B, G, R) is brought to bytes 0..2 by one PRMT, B and G weigh in by one DP2A, R by one IMAD.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bgr2gray_kernel(const uint8_t* __restrict__ src, int w, int h, uint8_t* __restrict__ dst,
                                                       int dPitch) {
    const int x4 = (blockIdx.x * 64 + threadIdx.x) * 4;
    const int y = blockIdx.y * 4 + threadIdx.y;
    if (x4 >= w || y >= h) return;
    const long long f = blockIdx.z;
    const uint8_t* s0 = src + ((f * h + y) * (long long)w + x4) * 3;
    const unsigned sh = ((unsigned)(uintptr_t)s0 & 3u) * 8u;
    const unsigned* a = reinterpret_cast<const unsigned*>((uintptr_t)s0 & ~(uintptr_t)3);
    // (the staging buffer has 16 bytes of slack: the last pixels of the last row may over-read)
    const unsigned w0 = __ldg(a), w1 = __ldg(a + 1), w2 = __ldg(a + 2), w3 = __ldg(a + 3);
    const unsigned lo = __funnelshift_r(w0, w1, sh), mid = __funnelshift_r(w1, w2, sh), hi = __funnelshift_r(w2, w3, sh);
    // lo = B0 G0 R0 B1 | mid = G1 R1 B2 G2 | hi = R2 B3 G3 R3
    const unsigned p0 = lo, p1 = __byte_perm(lo, mid, 0x0543), p2 = __byte_perm(mid, hi, 0x0432), p3 = hi >> 8;
    const unsigned kBG = 3735u | (19235u << 16);
    auto grey = [&](unsigned p) { return (__dp2a_lo(kBG, p, 16384u) + ((p >> 16) & 0xffu) * 9798u) >> 15; };
    const unsigned out = grey(p0) | (grey(p1) << 8) | (grey(p2) << 16) | (grey(p3) << 24);
    *reinterpret_cast<unsigned*>(dst + (f * h + y) * (long long)dPitch + x4) = out;      // pitch >= align128(w): the tail fits
}

// ------------------------------------------------------------------------------------------
// K1  pyramid level l-1 -> l.  cv::resize(INTER_LINEAR) on CV_8UC1 = 11-bit fixed-point separable
// bilinear; replaces ComputePyramid (OP_FtDtOrbSlam.cpp:936-960).
// One CTA = 128 destination columns x 4*rows destination rows; its source pixels arrive as ONE TMA box (x start rounded
// down to 16 bytes; the box sizes come from the host, ResizeTab).  One WARP walks down `rows` destination rows, a
// thread owns 4 adjacent columns of every row.  The horizontally interpolated values of the two source rows a
// destination row blends live in REGISTERS (top / bot): the source row index is monotone in the destination row, so
// stepping down one destination row either re-uses `bot` as the new `top` (source step 1) or computes both rows
// (step 2); no shared-memory strip, no dynamic row index — the only shared-memory traffic is the three aligned words
// per thread and source row.
//   horizontal, per source row: the three aligned words that cover the thread's <= 8-byte source window are shifted to
//     byte 0 (2 SHF), each pixel's two taps are picked by one PRMT (selectors are per-thread constants) and
//     S[s]*a0 + S[s+1]*a1 is one DP2A, >> 4
//   vertical, per destination row (row offset and coefficient pair are warp-uniform: lane j holds those of row j, one
//     SHFL each): ((b0*top + 2^17) >> 16) + ((b1*bot) >> 16) as two multiply-highs by the coefficients pre-shifted
//     left by 16, >> 2, pack, one 32-bit store.
// ------------------------------------------------------------------------------------------
template <int BOXW>
__global__ void NAV24_RS_LB resize_kernel(const __grid_constant__ CUtensorMap srcMap, int frameBase,
                                                     uint8_t* __restrict__ dst, int dPitch, long long dFrame, int dw, int dh,
                                                     ResizeTab t, int xBase) {
    extern __shared__ __align__(128) uint8_t s_rs[];       // [boxH][BOXW] source tile
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint2 s_yt[4][kResizeMaxRows];              // per warp strip: (source row inside the tile, b0 | b1 << 16) per destination row
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int rows = t.rows;                               // <= kResizeMaxRows (one lane per row of the warp's strip)
    const int x0c = xBase + blockIdx.x * 128, y0c = blockIdx.y * 4 * rows;      // (xBase: the remainder strip of a wide level)
    const int tileX0 = __ldg(t.xofs + x0c) & ~15, tileY0 = __ldg(t.yofs + y0c);
    const unsigned barAddr = smem_u32(&bar);
    // programmatic dependent launch (single-frame chains, launch_pyramid): the next level's CTAs may become resident now;
    // what they do before their own griddepcontrol.wait touches only the per-geometry tables, never a pyramid level
    asm volatile("griddepcontrol.launch_dependents;");
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barAddr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("griddepcontrol.wait;" ::: "memory");      // the source level is complete (no-op without the launch attribute)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr), "r"((unsigned)(BOXW * t.boxH)) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                smem_u32(s_rs)),
            "l"(&srcMap), "r"(tileX0), "r"(tileY0), "r"((int)blockIdx.z + frameBase), "r"(barAddr)
            : "memory");
    }
    const int x4 = x0c + lane * 4;
    const int y0 = y0c + wid * rows;
    if (y0 >= dh) return;                                  // warp-uniform
    const int nrows = min(rows, dh - y0);
    // vertical constants of the strip: lane j stores source row (inside the tile) and coefficient pair of destination
    // row y0 + j; the row loop reads them back with one broadcast 8-byte load
    if (lane < nrows) {
        const int yl = y0 + lane;
        s_yt[wid][lane] = make_uint2((unsigned)(__ldg(t.yofs + yl) - tileY0),
                                     __ldg(reinterpret_cast<const unsigned*>(t.yab) + yl));      // b0 | b1 << 16, both in [0, 2048]
    }
    // per-thread horizontal constants
    // (lanes beyond the level mirror the LAST ACTIVE lane: the host sizes the source box for the windows of the active
    // pixel groups; a lane clamped to column dw - 1 itself would start its window up to 3 source steps further right and
    // read past the last tile row — an out-of-range shared-memory access at e.g. scale 1.6, level width 116)
    const bool active = x4 < dw;
    const int x4c = min(x4, (dw - 1) & ~3);
    const int s0 = __ldg(t.xofs + x4c);
    const unsigned shift = (unsigned)(s0 & 3) * 8u;
    unsigned sel[4], ab[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = min(x4c + k, dw - 1);
        const int d = __ldg(t.xofs + x) - s0;              // 0..5
        sel[k] = (unsigned)d | ((unsigned)(d + 1) << 4);
        ab[k] = __ldg(reinterpret_cast<const unsigned*>(t.xab) + x);      // (a0, a1) as two u16 (both in [0, 2048])
    }
    const unsigned rpa = smem_u32(s_rs) + (unsigned)((s0 & ~3) - tileX0);      // the thread's window in tile row 0
    __syncwarp();
    {
        unsigned done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(barAddr), "r"(0u)
                : "memory");
        }
    }
    // horizontal pass of tile row i (bytes beyond the source image are zero-filled by TMA; their taps weigh 0).  PRMT is
    // written as PTX because the intrinsic re-masks the selector per use; the loads as PTX on a shared-window address so
    // that the three taps are immediate offsets of one register.
#define NAV24_HROW(i, h)                                                                                              \
    {                                                                                                                 \
        const unsigned a_ = rpa + (unsigned)(i) * (unsigned)BOXW;                                                     \
        unsigned w0_, w1_, w2_;                                                                                       \
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0_) : "r"(a_));                                                \
        asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1_) : "r"(a_));                                              \
        asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2_) : "r"(a_));                                              \
        const unsigned lo_ = __funnelshift_r(w0_, w1_, shift), hi_ = __funnelshift_r(w1_, w2_, shift);                \
        _Pragma("unroll") for (int k = 0; k < 4; ++k) {                                                               \
            unsigned tt_;                                                                                             \
            asm("prmt.b32 %0, %1, %2, %3;" : "=r"(tt_) : "r"(lo_), "r"(hi_), "r"(sel[k]));                            \
            h[k] = __dp2a_lo(ab[k], tt_, 0u) >> 4;                                                                    \
        }                                                                                                             \
    }
    // vertical pass of one destination row: ((b0*T + 0x20000) >> 16) + ((b1*B) >> 16) == umulhi(T, b0 << 16) + umulhi(B,
    // b1 << 16) + 2 (multiply-highs without addend: the fused form wants a 64-bit addend register pair per use), >> 2
    // on two pixels per register, bytes 0 and 2 of each picked by the final PRMT
#define NAV24_VROW(T, B, yc)                                                                                          \
    {                                                                                                                 \
        const unsigned c0_ = (yc) << 16, c1_ = (yc) & 0xffff0000u;                                                    \
        unsigned v_[4];                                                                                               \
        _Pragma("unroll") for (int k = 0; k < 4; ++k) v_[k] = __umulhi(T[k], c0_) + __umulhi(B[k], c1_) + 2u;         \
        const unsigned p01_ = (v_[0] | (v_[1] << 16)) >> 2, p23_ = (v_[2] | (v_[3] << 16)) >> 2;                      \
        const unsigned out_ = __byte_perm(p01_, p23_, 0x6420);                                                        \
        if (active) *reinterpret_cast<unsigned*>(dp) = out_;                                                          \
        dp += dPitch;                                                                                                 \
    }
    uint8_t* dp = dst + (long long)blockIdx.z * dFrame + (long long)y0 * dPitch + x4;
    const uint2* yt = s_yt[wid];
    unsigned A[4], B[4];
    int cur = (int)yt[0].x;
    NAV24_HROW(cur, A)
    NAV24_HROW(cur + 1, B)
    // two states so that "the old bottom row is the new top row" needs no register moves: (top, bot) = (A, B) or (B, A)
    int j = 0;
stateAB:
    for (; j < nrows; ++j) {
        const uint2 y = yt[j];
        const int sy = (int)y.x;
        if (sy == cur + 1) {                               // warp-uniform
            NAV24_HROW(sy + 1, A)
            cur = sy;
            NAV24_VROW(B, A, y.y)
            ++j;
            goto stateBA;
        }
        if (sy != cur) {
            NAV24_HROW(sy, A)
            NAV24_HROW(sy + 1, B)
            cur = sy;
        }
        NAV24_VROW(A, B, y.y)
    }
    return;
stateBA:
    for (; j < nrows; ++j) {
        const uint2 y = yt[j];
        const int sy = (int)y.x;
        if (sy == cur + 1) {
            NAV24_HROW(sy + 1, B)
            cur = sy;
            NAV24_VROW(A, B, y.y)
            ++j;
            goto stateAB;
        }
        if (sy != cur) {
            NAV24_HROW(sy, B)
            NAV24_HROW(sy + 1, A)
            cur = sy;
        }
        NAV24_VROW(B, A, y.y)
    }
#undef NAV24_HROW
#undef NAV24_VROW
}

// The same resize with EIGHT destination pixels per thread: a half-warp covers 128 destination columns, a warp 256, a CTA
// 256 columns x 4*rows rows fed by TWO TMA boxes side by side (one per 128-column half, each as in resize_kernel).  The
// eight pixels share one source window of four aligned words — pixels 0..3 tap the shifted word pair (0, 1), pixels 4..7
// the pair (1, 2) — and all the per-row control (row table load, state branch, pointer update, one 8-byte store), which was
// a third of resize_kernel's instructions.  The shared window needs a source step below ~1.28 px per destination pixel
// (true for the reference's 1.2 pyramid); the host checks every pixel group of a level (ResizeTab::wide) and launches
// resize_kernel for levels that do not qualify.
template <int BOXW>
__global__ void NAV24_RS_LB resize8_kernel(const __grid_constant__ CUtensorMap srcMap, int frameBase,
                                                      uint8_t* __restrict__ dst, int dPitch, long long dFrame, int dw, int dh,
                                                      ResizeTab t) {
    extern __shared__ __align__(128) uint8_t s_rs[];       // [2][boxH][BOXW] source tiles of the two 128-column halves
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint2 s_yt[4][kResizeMaxRows];              // per warp strip: (source row inside the tile, b0 | b1 << 16) per destination row
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int rows = t.rows;
    const int x0c = blockIdx.x * 256, y0c = blockIdx.y * 4 * rows;
    const bool twoBoxes = x0c + 128 < dw;
    const int half = lane >> 4;
    const int xh = x0c + 128 * half;                       // first destination column of this lane's half
    const int tileY0 = __ldg(t.yofs + y0c);
    const unsigned barAddr = smem_u32(&bar);
    const unsigned tileBytes = (unsigned)(BOXW * t.boxH);
    const unsigned tileStride = (tileBytes + 127u) & ~127u;      // TMA destinations are 128-byte aligned
    asm volatile("griddepcontrol.launch_dependents;");       // (see resize_kernel)
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barAddr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr), "r"(twoBoxes ? 2u * tileBytes : tileBytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                smem_u32(s_rs)),
            "l"(&srcMap), "r"(__ldg(t.xofs + x0c) & ~15), "r"(tileY0), "r"((int)blockIdx.z + frameBase), "r"(barAddr)
            : "memory");
        if (twoBoxes)
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                    smem_u32(s_rs) + tileStride),
                "l"(&srcMap), "r"(__ldg(t.xofs + x0c + 128) & ~15), "r"(tileY0), "r"((int)blockIdx.z + frameBase), "r"(barAddr)
                : "memory");
    }
    const int x8 = xh + (lane & 15) * 8;
    const int y0 = y0c + wid * rows;
    if (y0 >= dh) return;                                  // warp-uniform
    const int nrows = min(rows, dh - y0);
    if (lane < nrows) {
        const int yl = y0 + lane;
        s_yt[wid][lane] = make_uint2((unsigned)(__ldg(t.yofs + yl) - tileY0),
                                     __ldg(reinterpret_cast<const unsigned*>(t.yab) + yl));      // b0 | b1 << 16, both in [0, 2048]
    }
    // per-thread horizontal constants (columns beyond the level are clamped: their results are never stored)
    const bool active = x8 < dw;
    const int x8c = min(x8, (dw - 1) & ~7);                // (lanes beyond the level mirror the last active lane, see resize_kernel)
    const int s0 = __ldg(t.xofs + x8c);
    const int tileX0 = __ldg(t.xofs + min(xh, dw - 1)) & ~15;
    const unsigned shift = (unsigned)(s0 & 3) * 8u;
    unsigned sel[8], ab[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int x = min(x8c + k, dw - 1);
        int d = __ldg(t.xofs + x) - s0;                    // pixels 0..3: 0..6 inside the word pair (0, 1); 4..7: 4..10 -> 0..6 inside (1, 2)
        d = k < 4 ? min(d, 6) : min(max(d - 4, 0), 6);
        sel[k] = (unsigned)d | ((unsigned)(d + 1) << 4);
        ab[k] = __ldg(reinterpret_cast<const unsigned*>(t.xab) + x);      // (a0, a1) as two u16 (both in [0, 2048])
    }
    const unsigned rpa = smem_u32(s_rs) + (half ? tileStride : 0u) + (unsigned)((s0 & ~3) - tileX0);      // the thread's window in tile row 0
    __syncwarp();
    {
        unsigned done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(barAddr), "r"(0u)
                : "memory");
        }
    }
#define NAV24_HROW8(i, h)                                                                                             \
    {                                                                                                                 \
        const unsigned a_ = rpa + (unsigned)(i) * (unsigned)BOXW;                                                     \
        unsigned w0_, w1_, w2_, w3_;                                                                                  \
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0_) : "r"(a_));                                                \
        asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1_) : "r"(a_));                                              \
        asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2_) : "r"(a_));                                              \
        asm volatile("ld.shared.u32 %0, [%1+12];" : "=r"(w3_) : "r"(a_));                                             \
        const unsigned lo_ = __funnelshift_r(w0_, w1_, shift), mid_ = __funnelshift_r(w1_, w2_, shift),               \
                       hi_ = __funnelshift_r(w2_, w3_, shift);                                                        \
        _Pragma("unroll") for (int k = 0; k < 8; ++k) {                                                               \
            unsigned tt_;                                                                                             \
            if (k < 4) asm("prmt.b32 %0, %1, %2, %3;" : "=r"(tt_) : "r"(lo_), "r"(mid_), "r"(sel[k]));                \
            else asm("prmt.b32 %0, %1, %2, %3;" : "=r"(tt_) : "r"(mid_), "r"(hi_), "r"(sel[k]));                      \
            h[k] = __dp2a_lo(ab[k], tt_, 0u) >> 4;                                                                    \
        }                                                                                                             \
    }
#define NAV24_VROW8(T, B, yc)                                                                                         \
    {                                                                                                                 \
        const unsigned c0_ = (yc) << 16, c1_ = (yc) & 0xffff0000u;                                                    \
        unsigned v_[8];                                                                                               \
        _Pragma("unroll") for (int k = 0; k < 8; ++k) v_[k] = __umulhi(T[k], c0_) + __umulhi(B[k], c1_) + 2u;         \
        const unsigned p01_ = (v_[0] | (v_[1] << 16)) >> 2, p23_ = (v_[2] | (v_[3] << 16)) >> 2;                      \
        const unsigned p45_ = (v_[4] | (v_[5] << 16)) >> 2, p67_ = (v_[6] | (v_[7] << 16)) >> 2;                      \
        if (active) *reinterpret_cast<uint2*>(dp) = make_uint2(__byte_perm(p01_, p23_, 0x6420), __byte_perm(p45_, p67_, 0x6420)); \
        dp += dPitch;                                                                                                 \
    }
    uint8_t* dp = dst + (long long)blockIdx.z * dFrame + (long long)y0 * dPitch + x8;
    const uint2* yt = s_yt[wid];
    unsigned A[8], B[8];
    int cur = (int)yt[0].x;
    NAV24_HROW8(cur, A)
    NAV24_HROW8(cur + 1, B)
    int j = 0;
stateAB:
    for (; j < nrows; ++j) {
        const uint2 y = yt[j];
        const int sy = (int)y.x;
        if (sy == cur + 1) {                               // warp-uniform
            NAV24_HROW8(sy + 1, A)
            cur = sy;
            NAV24_VROW8(B, A, y.y)
            ++j;
            goto stateBA;
        }
        if (sy != cur) {
            NAV24_HROW8(sy, A)
            NAV24_HROW8(sy + 1, B)
            cur = sy;
        }
        NAV24_VROW8(A, B, y.y)
    }
    return;
stateBA:
    for (; j < nrows; ++j) {
        const uint2 y = yt[j];
        const int sy = (int)y.x;
        if (sy == cur + 1) {
            NAV24_HROW8(sy + 1, B)
            cur = sy;
            NAV24_VROW8(A, B, y.y)
            ++j;
            goto stateAB;
        }
        if (sy != cur) {
            NAV24_HROW8(sy, B)
            NAV24_HROW8(sy + 1, A)
            cur = sy;
        }
        NAV24_VROW8(B, A, y.y)
    }
#undef NAV24_HROW8
#undef NAV24_VROW8
}

// ------------------------------------------------------------------------------------------
// K2  FAST-9/16 per cell with threshold fallback.  One CTA per (segment, frame); a segment is a run of up to
// kFastMaxSegCells horizontally adjacent cells of one cell row (FastSeg, built on the host).
// Reproduces the cell loop of ComputeKeyPointsOctTree (OP_FtDtOrbSlam.cpp:751-818) and
// cv::FAST(cell, th, nms=true): the corner score is threshold independent, a
// keypoint at threshold t is a strict 3x3 local maximum of the score map with score >= t, the
// detection domain is the cell minus a 3-px rim (so the interiors of adjacent cells tile the image and
// only the rims overlap), scores outside the cell's own interior count as 0 in its NMS, and a cell
// with no survivor at iniThFAST is redone at minThFAST.
//   tile   : one TMA box (segment interior + rim), shared by the cells of the segment
//   stage A: antipodal-pair rejection, 4 pixels per 32-bit op; a thread owns one word column and walks down
//            its rows with the column's last seven words in registers (the +-3 row taps)
//   stage B: exact score of the survivors through a queue (balanced: every warp owns a slice of it and compacts the
//            pixels that reach the threshold to the front of its slice)
//   stage C: 3x3 NMS (cell-aware) -> bit mask in raster order, keypoints compacted in the slice again
//   count  : one warp per cell, one lane per row: keypoints of the cell and above each of its rows; a cell whose
//            count is final appends to the level's raw list with one global atomic; empty cells are redone at minTh
//   emit   : every keypoint writes itself at (cell offset + row prefix + keypoints to its left in the row)
//   A segment whose survivors overflow the queue (pure noise) takes a dense path: score every pixel, NMS over the map.
// ------------------------------------------------------------------------------------------
// Score of one pixel.  Ring differences are packed as biased u16x2 lanes (d+256, 256-d) by one IMAD;
// min over an arc of 9 = min3 of three min3's (VIMNMX3.U16x2); max over the 16 arcs by max3.
__device__ __forceinline__ int fast_score(const uint8_t* c, int pitch) {
    const unsigned cv = c[0];
    const unsigned cn = (256u - cv) + ((cv + 256u) << 16);
    const uint8_t* rm3 = c - 3 * pitch; const uint8_t* rm2 = c - 2 * pitch; const uint8_t* rm1 = c - pitch;
    const uint8_t* rp1 = c + pitch;     const uint8_t* rp2 = c + 2 * pitch; const uint8_t* rp3 = c + 3 * pitch;
    unsigned v[16];
#define NAV24_V(k, ptr, dx) v[k] = (unsigned)(ptr)[dx] * 0xFFFF0001u + cn;
    NAV24_V(0, rp3, 0)  NAV24_V(1, rp3, 1)   NAV24_V(2, rp2, 2)   NAV24_V(3, rp1, 3)
    NAV24_V(4, c, 3)    NAV24_V(5, rm1, 3)   NAV24_V(6, rm2, 2)   NAV24_V(7, rm3, 1)
    NAV24_V(8, rm3, 0)  NAV24_V(9, rm3, -1)  NAV24_V(10, rm2, -2) NAV24_V(11, rm1, -3)
    NAV24_V(12, c, -3)  NAV24_V(13, rp1, -3) NAV24_V(14, rp2, -2) NAV24_V(15, rp3, -1)
#undef NAV24_V
    unsigned m3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m3[k] = __vimin3_u16x2(v[k], v[(k + 1) & 15], v[(k + 2) & 15]);
    unsigned a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = __vimin3_u16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
    unsigned b0 = __vimax3_u16x2(a[0], a[1], a[2]), b1 = __vimax3_u16x2(a[3], a[4], a[5]);
    unsigned b2 = __vimax3_u16x2(a[6], a[7], a[8]), b3 = __vimax3_u16x2(a[9], a[10], a[11]);
    unsigned b4 = __vimax3_u16x2(a[12], a[13], a[14]);
    b0 = __vimax3_u16x2(b0, b1, b2);
    b3 = __vimax3_u16x2(b3, b4, a[15]);
    b0 = __vmaxu2(b0, b3);
    return (int)max(b0 & 0xffffu, b0 >> 16) - 257;
}

// Conservative per-byte (x1 > t || x2 > t) for t < 128: 0x80 in every byte that passes.  k = (0x7f - t) * 0x01010101.
// The adds run over the whole word: a byte >= 129 + t carries into its neighbour, which can only make the neighbour
// pass when it equals t (a false positive of the FILTER, removed by the exact score), and the overflowing byte
// itself has bit 7 set, so it passes through the "| x" term.
__device__ __forceinline__ unsigned gt_any2(unsigned x1, unsigned x2, unsigned k) { return (x1 + k) | x1 | (x2 + k) | x2; }

// set bits of kmask row `rowp` inside the column range [x0, x1), x1 > x0: popcount
__device__ __forceinline__ int mask_range_count(const unsigned* rowp, int x0, int x1) {
    const int wlo = x0 >> 5, whi = (x1 - 1) >> 5;
    int n = 0;
    for (int w = wlo; w <= whi; ++w) {
        unsigned m = rowp[w];
        if (w == wlo) m &= 0xffffffffu << (x0 & 31);
        if (w == whi) m &= 0xffffffffu >> (31 - ((x1 - 1) & 31));
        n += __popc(m);
    }
    return n;
}

// Dynamic shared memory: [tile | score map | keypoint bit mask | queue u16], sizes from FastSmem.
// Queue and winner entries are the byte offset of the pixel inside the tile (row * kFastPitch + column).
__global__ void NAV24_FAST_LB fast_band_kernel(const __grid_constant__ FrameGeom g, const DevPtrs p,
                                                                 const __grid_constant__ TmaMaps maps, const FastSmem sm,
                                                                 int iniTh, int minTh, int segBase) {
    extern __shared__ __align__(128) uint8_t s_dyn[];
    uint8_t* tile = s_dyn;
    uint8_t* smap = s_dyn + sm.offMap;
    unsigned* kmask = reinterpret_cast<unsigned*>(s_dyn + sm.offMask);
    unsigned short* queue = reinterpret_cast<unsigned short*>(s_dyn + sm.offQueue);
    __shared__ __align__(8) unsigned long long bar;
    __shared__ int s_q, s_ovf, s_empty;
    __shared__ int s_scan[33];
    __shared__ int s_cellFirst[kFastMaxSegCells], s_cellEnd[kFastMaxSegCells], s_cellBase[kFastMaxSegCells];
    constexpr int NT = kFastThreads, NW = kFastThreads / 32, P = kFastPitch, PW = kFastPitch / 4;

    const int tid = threadIdx.x, f = blockIdx.y;
    const int lane = tid & 31, wid = tid >> 5;
    asm volatile("griddepcontrol.launch_dependents;");
    FastSeg sg;                                                      // 2 x 16 B, the same for all threads
    {
        const uint4* sp = reinterpret_cast<const uint4*>(p.segs + segBase + blockIdx.x);
        uint4* dp = reinterpret_cast<uint4*>(&sg);
        dp[0] = __ldg(sp); dp[1] = __ldg(sp + 1);
    }
    const int l = sg.level;
    const LevelGeom& L = g.lv[l];
    uint2* info = p.cellInfo + (long long)f * g.totalCells + sg.cell0;
    const int nv = sg.nv, ih = sg.ih, iw = sg.iw;
    if (nv == 0) {
        if (tid < sg.nc) info[tid] = make_uint2(0u, 0u);
        if (tid == 0) asm volatile("griddepcontrol.wait;" ::: "memory");      // (every CTA orders itself behind the previous grid: completion stays transitive)
        return;
    }
    const int wCell = L.wCell;
    // TMA needs a 16-byte aligned start in x: the box starts at xa <= iniX0-1, the interior at tile column o+4
    const int iniX = sg.iniX0, iniY = sg.iniY;
    const int xa = (iniX - 1) & ~15, o = iniX - 1 - xa;
    const int boxH = L.boxH;
    const unsigned barAddr = smem_u32(&bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barAddr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_q = 0; s_ovf = 0; s_empty = 0;
    }
    __syncthreads();
    if (tid == 0) {
        // (programmatic dependent launch, launch_fast: everything above reads the segment table and writes this segment's own
        // cell entries only; the pyramid level is complete once this returns)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        const unsigned bytes = (unsigned)(P * boxH);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                smem_u32(tile)),
            "l"(&maps.m[l]), "r"(xa), "r"(iniY), "r"(f + p.frameBase), "r"(barAddr)
            : "memory");
    }
    // while the tile is in flight: clear the score map and the bit mask behind it (one range of 16-byte stores)
    const int wpr = (iw + 31) >> 5;                  // mask words per interior row
    {
        const int n16 = (sm.offMask - sm.offMap) / 16 + (wpr * ih + 3) / 4;
        uint4* z = reinterpret_cast<uint4*>(smap);
        for (int i = tid; i < n16; i += NT) z[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid >= nv && tid < sg.nc) info[tid] = make_uint2(0u, 0u);      // skipped cells (:756, :765)
    {
        unsigned done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(barAddr), "r"(0u)
                : "memory");
        }
    }
    __syncthreads();

    const int c_lo = o + 4, c_hi = o + 4 + iw;      // interior tile columns [c_lo, c_hi)
    const int gx0 = c_lo >> 2;
    const int ngx = sg.ngx, rc = sg.rc;             // stage A work items: (word column, chunk of rc rows), one round
    const int nItems = ngx * sg.nChunks;
    unsigned emptyCells = 0;                         // pass 1: cells without a keypoint at iniTh
    unsigned sqAddr = smem_u32(&s_q), qAddr = smem_u32(queue), tileAddr = smem_u32(tile);
    // (opaque to the compiler: otherwise it re-derives these shared-window addresses — S2R + LEA — at every use in stage A)
    asm volatile("mov.u32 %0, %0;" : "+r"(sqAddr));
    asm volatile("mov.u32 %0, %0;" : "+r"(qAddr));
    asm volatile("mov.u32 %0, %0;" : "+r"(tileAddr));
    RawRec* outL = p.raw + (long long)f * g.rawPerFrame + L.rawOff;
    const int xBase = 3 + sg.cj0 * wCell, yBase = 3 + sg.ci * L.hCell;      // :811-812, relative to (minBorderX, minBorderY)

    int th = iniTh;
    for (int pass = 0; pass < 2; ++pass) {
        th = pass == 0 ? iniTh : minTh;
        const unsigned kk = (unsigned)(0x7f - min(th, 127)) * 0x01010101u;
        const bool reject = th < 128;
        for (int it = tid; it < nItems; it += NT) {
            // it / ngx by multiply-high; a divisor of 1 has no 32-bit magic (2^32): a last segment that is ONE cell whose
            // interior is 1..4 px of one word (e.g. a 1149-px-wide level: 31 cells of 37, the last 7 px wide)
            const int chunk = ngx == 1 ? it : (int)__umulhi((unsigned)it, sg.magicG), gx = it - chunk * ngx;
            const int wc = gx0 + gx, cb = 4 * wc;                      // tile column of byte 0
            const int first = max(c_lo - cb, 0), last = min(c_hi - cb, 4);      // interior bytes [first, last)
            unsigned vmask = (0x80808080u >> (8 * (4 - last))) & (0x80808080u << (8 * first));
            if (pass) {                                                // only the pixels of the cells that came out empty
                unsigned m = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int xx = min(max(cb + b - c_lo, 0), iw - 1);
                    if ((emptyCells >> __umulhi((unsigned)xx, L.magicW)) & 1u) m |= 0x80u << (8 * b);
                }
                vmask &= m;
            }
            const int r0 = 3 + chunk * rc, r1 = min(r0 + rc, 3 + ih);
            if (vmask == 0u || r0 >= r1) continue;
            const uint8_t* tp = tile + r0 * P + cb;                    // this column's word of row r
            const uint8_t* tEnd = tile + r1 * P + cb;
#define NAV24_W(dy) (*reinterpret_cast<const unsigned*>(tp + (dy) * P))
            unsigned w0 = NAV24_W(-3), w1 = NAV24_W(-2), w2 = NAV24_W(-1), w3 = NAV24_W(0), w4 = NAV24_W(1), w5 = NAV24_W(2), w6;
            // one row: the word entering the window is row r+3; (a0 .. a6) = rows r-3 .. r+3 of this column.
            // (Collecting the survivors of seven rows in a register and queueing them with a per-lane loop measured
            // slower: vertical edges put all of a column's survivors into one lane.)
#define NAV24_STEP(a0, a1, a2, a3, a4, a5, a6)                                                                       \
            {                                                                                                        \
                a6 = NAV24_W(3);                                                                                     \
                const unsigned C = a3;                                                                               \
                unsigned alive = vmask;                                                                              \
                if (reject) {                                                                                        \
                    const unsigned Lw = *reinterpret_cast<const unsigned*>(tp - 4);                                  \
                    const unsigned Rw = *reinterpret_cast<const unsigned*>(tp + 4);                                  \
                    alive &= gt_any2(__vabsdiffu4(a0, C), __vabsdiffu4(a6, C), kk) &                                 \
                             gt_any2(__vabsdiffu4(__byte_perm(Lw, C, 0x4321), C),                                    \
                                     __vabsdiffu4(__byte_perm(C, Rw, 0x6543), C), kk);                               \
                }                                                                                                    \
                if (alive) {      /* (PTX: the C++ atomicAdd drags in the compiler's warp-aggregation path and */    \
                    unsigned pos;     /* generic-address set-up, 40 instructions per row instead of 16) */              \
                    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(pos) : "r"(sqAddr), "r"(__popc(alive)) : "memory"); \
                    const unsigned e = smem_u32(tp) - tileAddr;                                                      \
                    if (pos + 4 <= kFastQueueCap) {      /* else: s_q > cap -> the dense path below */                \
                        unsigned qa = qAddr + 2u * pos;                                                              \
                        if (alive & 0x80u) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(qa), "r"(e) : "memory"); qa += 2u; }        \
                        if (alive & 0x8000u) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(qa), "r"(e + 1u) : "memory"); qa += 2u; } \
                        if (alive & 0x800000u) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(qa), "r"(e + 2u) : "memory"); qa += 2u; } \
                        if (alive & 0x80000000u) asm volatile("st.shared.u16 [%0], %1;" ::"r"(qa), "r"(e + 3u) : "memory");           \
                    } else {                                                                                         \
                        s_ovf = 1;                                                                                   \
                    }                                                                                                \
                }                                                                                                    \
                tp += P;                                                                                             \
                if (tp == tEnd) break;                                                                               \
            }
            while (true) {
                NAV24_STEP(w0, w1, w2, w3, w4, w5, w6)
                NAV24_STEP(w1, w2, w3, w4, w5, w6, w0)
                NAV24_STEP(w2, w3, w4, w5, w6, w0, w1)
                NAV24_STEP(w3, w4, w5, w6, w0, w1, w2)
                NAV24_STEP(w4, w5, w6, w0, w1, w2, w3)
                NAV24_STEP(w5, w6, w0, w1, w2, w3, w4)
                NAV24_STEP(w6, w0, w1, w2, w3, w4, w5)
            }
#undef NAV24_STEP
#undef NAV24_W
        }
        __syncthreads();
        // stage C body: 3x3 non-max suppression of the corner at tile offset e; the neighbours in the adjacent cell's
        // interior count as 0
        auto nms = [&](int e) {
            const uint8_t* q = smap + e;
            const int er = (int)__umulhi((unsigned)e, 0xFFFFFFFFu / (unsigned)P + 1u);      // e / P: tile row (e < 65536)
            const int xx = (e - er * P) - c_lo;
            const int cl = xx - (int)__umulhi((unsigned)xx, L.magicW) * wCell;      // column inside the cell's interior
            const int s = q[0];
            int m = max((int)q[-P], (int)q[P]);
            const int ml = max(max((int)q[-P - 1], (int)q[-1]), (int)q[P - 1]);
            const int mr = max(max((int)q[-P + 1], (int)q[1]), (int)q[P + 1]);
            if (cl > 0) m = max(m, ml);
            if (cl < wCell - 1) m = max(m, mr);
            if (s > m) {
                const int bi = (er - 3) * (wpr << 5) + xx;
                atomicOr(&kmask[bi >> 5], 1u << (bi & 31));
                return true;
            }
            return false;
        };
        const bool dense = s_ovf != 0;
        if (!dense) {
            // stage B: exact score of the survivors.  Each warp owns a contiguous slice of the queue and compacts the
            // pixels that reach the threshold to the front of its slice (no block barrier, no atomics), then runs
            // stage C over them.
            const int nq = s_q;
            const int per = (nq + NW - 1) / NW, lo = wid * per, hi = min(lo + per, nq);
            int wpos = lo;
            for (int base = lo; base < hi; base += 32) {
                const int qi = base + lane;
                int e = 0;
                bool ok = false;
                if (qi < hi) {
                    e = queue[qi];
                    const int s = fast_score(tile + e, P);
                    ok = s >= th;
                    if (ok) smap[e] = (uint8_t)s;
                }
                const unsigned m = __ballot_sync(0xffffffffu, ok);      // (orders the round's reads before its writes)
                if (ok) queue[wpos + __popc(m & ((1u << lane) - 1u))] = (unsigned short)e;
                wpos += __popc(m);
            }
            __syncthreads();                                             // the score map is complete
            // stage C over the warp's corners: the keypoints (strict 3x3 maxima) are marked in the bit mask
            for (int qi = lo + lane; qi < wpos; qi += 32) nms(queue[qi]);
        } else {
            // dense path (rare: the survivors of stage A did not fit the queue, e.g. an image of pure noise): score every
            // interior pixel (in pass 2: of the empty cells), then NMS over the score map
            for (int row = 3; row < 3 + ih; ++row)
                for (int xx = tid; xx < iw; xx += NT) {
                    if (pass && !((emptyCells >> __umulhi((unsigned)xx, L.magicW)) & 1u)) continue;
                    const int e = row * P + c_lo + xx;
                    const int s = fast_score(tile + e, P);
                    if (s >= th) smap[e] = (uint8_t)s;
                }
            __syncthreads();
            for (int row = 3; row < 3 + ih; ++row)
                for (int xx = tid; xx < iw; xx += NT) {
                    if (pass && !((emptyCells >> __umulhi((unsigned)xx, L.magicW)) & 1u)) continue;
                    const int e = row * P + c_lo + xx;
                    if (smap[e]) nms(e);
                }
        }
        __syncthreads();
        // count and ordered emit by the whole CTA in ONE round: a thread per (cell, row) — the host sizes the segments so that
        // cells x rows <= 256 — reads its row of its cell out of the bit mask (<= 3 words, the cell's column range masked
        // in); one block scan gives every row its offset, the first and last row of a cell turn that into offsets inside
        // the cell and the cell's count; a cell whose count is final takes its place in the level's raw list with ONE global
        // atomic, and every thread writes its row's keypoints at (cell offset + keypoints above its row) by walking its
        // set bits — the row-major order inside the cell that cv::FAST reports, without a rank computation per keypoint.
        const bool lastPass = pass || minTh >= iniTh;
        {
            const int hC = L.hCell;
            const int k = (int)__umulhi((unsigned)tid, L.magicH), row = tid - k * hC;      // tid / hCell (tid < 65536)
            const bool item = k < nv && row < ih && !(pass && !((emptyCells >> k) & 1u));
            const int x0 = k * wCell, x1 = min(x0 + wCell, iw);
            const int wlo = x0 >> 5, nwd = ((x1 - 1) >> 5) - wlo + 1;                      // 1..3 mask words per row
            unsigned mw[3];
            int cnt = 0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                unsigned m = (item && j < nwd) ? kmask[row * wpr + wlo + j] : 0u;
                if (j == 0) m &= 0xffffffffu << (x0 & 31);
                if (j == nwd - 1) m &= 0xffffffffu >> (31 - ((x1 - 1) & 31));
                mw[j] = m;
                cnt += __popc(m);
            }
            int tot;
            const int ex = block_excl_scan(cnt, &tot, s_scan);
            if (k < nv && row == 0) s_cellFirst[k] = ex;                                   // (rows 0 and ih-1 exist for every valid cell)
            if (k < nv && row == ih - 1) s_cellEnd[k] = ex + cnt;
            __syncthreads();
            if (tid < nv && !(pass && !((emptyCells >> tid) & 1u))) {
                const int run = s_cellEnd[tid] - s_cellFirst[tid];
                if (run == 0 && !lastPass) {
                    atomicOr(&s_empty, 1 << tid);      // :787 "if(vKeysCell.empty())" -> retry at minThFAST
                    s_cellBase[tid] = -1;
                } else {
                    int base = 0;
                    if (run > 0) {
                        base = atomicAdd(p.rawCount + f * g.nlevels + l, run);
                        if (base + run > L.rawCap) { atomicOr(p.err, ERR_RAW_OVERFLOW); base = -1; }
                    }
                    s_cellBase[tid] = base;
                    info[tid] = base >= 0 ? make_uint2((unsigned)base, (unsigned)run) : make_uint2(0u, 0u);
                }
            }
            __syncthreads();
            if (cnt > 0) {
                const int base = s_cellBase[k];
                if (base >= 0) {
                    RawRec* o = outL + base + (ex - s_cellFirst[k]);
                    const uint8_t* srow = smap + (row + 3) * P + c_lo;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        unsigned m = mw[j];
                        while (m) {
                            const int xx = ((wlo + j) << 5) + __ffs(m) - 1;
                            m &= m - 1;
                            RawRec r;
                            r.x = (unsigned short)(xx + xBase);
                            r.y = (unsigned short)(row + yBase);
                            r.score = srow[xx]; r.pad = 0;
                            *o++ = r;
                        }
                    }
                }
            }
        }
        __syncthreads();
        emptyCells = (unsigned)s_empty;
        if (lastPass || emptyCells == 0u) break;
        __syncthreads();                             // the queue is refilled by the next pass
        if (tid == 0) { s_q = 0; s_ovf = 0; }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// K3  quadtree distribution, one CTA per (level, frame).  Reproduces DistributeOctTree
// (OP_FtDtOrbSlam.cpp:502-725) including std::list ordering and the unstable std::sort of the
// "largest first" phase.  The node list is an array in list order; every key carries the index
// of its node.  One pass = count children (parallel over keys) -> choose the parents to split
// (all non-leaf nodes, or the sorted order with the size>=N break) -> rebuild the list:
//   new list = reverse(children in creation order) ++ (old list minus split parents)
// which is what push_front of each child + erase of the parent produce.
// ------------------------------------------------------------------------------------------
// bytes of the sort scratch for `cap` records: records | two u16 position lists | leaf-start bit mask | two range lists of
// sort_block (>= the 192-int stack of sort_warp), rounded to 16 (see stdsort_warp.cuh)
__host__ __device__ inline size_t quadtree_sort_bytes(int cap) {
    const size_t rng = (size_t)(2 * (3 * (cap / 16 + 2) + 1)) > 192 ? (size_t)(2 * (3 * (cap / 16 + 2) + 1)) : 192;
    return ((size_t)cap * 12 + (size_t)((cap + 31) / 32) * 4 + rng * 4 + 16 + 15) / 16 * 16;
}

__device__ __forceinline__ int quadrant_of(const QNode& n, int x, int y) {
    const int mx = n.x0 + ((n.x1 - n.x0 + 1) >> 1);     // UL.x + ceil(w/2)  (:386)
    const int my = n.y0 + ((n.y1 - n.y0 + 1) >> 1);
    return (x < mx ? 0 : 1) + (y < my ? 0 : 2);
}

// Working tables of one (level, frame).  The ordered keys and the node of every key are STREAMED (coalesced, index i)
// and stay in global memory; the tables that are accessed at random through the node index — the two node lists, the
// child counts (atomics), per-node scratch, best key per node (atomics) — live in the CTA's shared memory when
// SM = true (typed through this template so that every access compiles to LDS / STS / ATOMS; a run-time choice
// between spaces makes the pointers generic: +24 registers and slower) and in the per-frame global slabs when SM = false
// (levels whose node tables do not fit, e.g. the x5 feature mode or 4K level 0).  Measured: with the keys in shared memory
// as well the CTA needs 72 KB, three CTAs per SM instead of seven, and the kernel — bound by barrier and dependent-access
// latency, not by bandwidth — got 23 % slower instead of faster.
template <int NT, bool SM>
__device__ __forceinline__ void quadtree_run(const FrameGeom& g, const DevPtrs& p, const int l, const int f, const int n,
                                             const int sortSmemCap, unsigned long long* s_sort, BlockScan& s_scan, int* s_K, int* s_nexp,
                                             const RawRec* __restrict__ keys, unsigned short* __restrict__ nodeOfKey, QNode* cur, QNode* nxt,
                                             int* childCnt, int* aux, unsigned* best, const RawRec* __restrict__ raw,
                                             const uint2* __restrict__ cinfo, const int* __restrict__ cdst, const int nC, RawRec* keysOut,
                                             const uint2* s_ci, const int* s_dst) {
    const int tid = threadIdx.x, nth = NT;
    const LevelGeom& L = g.lv[l];
    const long long fn = (long long)f * g.nodesPerFrame + L.nodeOff;
    unsigned long long* sortRec = p.sortRec + fn;
    LevelKp* lkp = p.lkp + (long long)f * g.kpPerFrame + L.kpOff;

    // (b) root nodes (:505-547)
    const int N = L.quota, nIni = L.nIni;
    const float hX = L.hX;
    for (int r = tid; r < nIni; r += nth) childCnt[r] = 0;
    __syncthreads();
    // The per-cell lists of the FAST kernel go into the reference order (cells row-major, row-major inside a cell: keysOut =
    // `keys`, read back only after the barrier below) and every key is counted into its root on the way.
    // Levels whose cell table does not fit the CTA's shared memory (4K frames): a warp takes 32 consecutive cells, lane j the
    // table entry of cell c0 + j.  Their keys are consecutive in the ordered
    // array, so entry t of the group goes to keys[first + t] (coalesced) and comes from the cell whose inclusive count
    // first exceeds t (binary search over the lanes); the trips are independent, four are in flight at a time.  (One warp
    // per cell — table entry, then keys, then the next cell — was a chain of dependent L2 latencies: 17 % of this kernel's
    // stall samples.)
    if (s_ci) {
        // the cell table (source offset, ordered offset) is in shared memory: a thread per key finds its cell by binary search
        // over the ordered offsets — the last cell whose offset is <= i — so that the only global round trip of the
        // gather is the key itself, four per thread in flight
        for (int i0 = tid; i0 < n; i0 += 4 * nth) {
            int src[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = min(i0 + u * nth, n - 1);
                int lo = 0, hi = nC;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s_dst[mid] <= i) lo = mid; else hi = mid;
                }
                src[u] = (int)s_ci[lo].x + (i - s_dst[lo]);
            }
            RawRec rr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) rr[u] = raw[src[u]];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * nth;
                if (i < n) {
                    keysOut[i] = rr[u];
                    int r = (int)__fdiv_rn((float)rr[u].x, hX);
                    if (r >= nIni || r < 0) { atomicOr(p.err, ERR_ROOT_RANGE); r = nIni - 1; }
                    nodeOfKey[i] = (unsigned short)r;
                    atomicAdd(&childCnt[r], 1);
                }
            }
        }
    } else {
        const int lane = tid & 31;
        for (int c0 = (tid >> 5) * 32; c0 < nC; c0 += nth) {
            const int c = c0 + lane;
            const uint2 ci = c < nC ? cinfo[c] : make_uint2(0u, 0u);
            const int first = cdst[c0];
            int incl = (int)ci.y;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int T = __shfl_sync(0xffffffffu, incl, 31);
            const int srcOff = (int)ci.x - (incl - (int)ci.y);      // source index of entry t of this lane's cell = srcOff + t
            for (int t0 = 0; t0 < T; t0 += 128) {
                int src[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = t0 + 32 * u + lane;
                    int col = 0;
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1) {
                        const int probe = __shfl_sync(0xffffffffu, incl, col + step - 1);
                        if (probe <= t) col += step;
                    }
                    src[u] = __shfl_sync(0xffffffffu, srcOff, min(col, 31)) + t;
                }
                RawRec rr[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (t0 + 32 * u + lane < T) rr[u] = raw[src[u]];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = t0 + 32 * u + lane;
                    if (t < T) {
                        keysOut[first + t] = rr[u];
                        int r = (int)__fdiv_rn((float)rr[u].x, hX);
                        if (r >= nIni || r < 0) { atomicOr(p.err, ERR_ROOT_RANGE); r = nIni - 1; }
                        nodeOfKey[first + t] = (unsigned short)r;
                        atomicAdd(&childCnt[r], 1);
                    }
                }
            }
        }
    }
    __syncthreads();
    int size = 0;
    for (int base = 0; base < nIni; base += nth) {
        const int r = base + tid;
        const int cnt = r < nIni ? childCnt[r] : 0;
        int tot;
        const int ex = block_excl_scan(cnt > 0, &tot, s_scan);
        if (cnt > 0) {
            QNode q;
            q.x0 = (short)(int)__fmul_rn(hX, (float)r);
            q.x1 = (short)(int)__fmul_rn(hX, (float)(r + 1));
            q.y0 = 0; q.y1 = (short)(L.maxBY - kMinBorder);
            q.count = cnt;
            cur[size + ex] = q;
        }
        if (r < nIni) aux[r] = size + ex;
        size += tot;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nth) nodeOfKey[i] = (unsigned short)aux[nodeOfKey[i]];
    __syncthreads();

    // (c) split passes (:555-700)
    bool finish = false, phase2 = false;
    while (!finish) {
        const int prevSize = size;
        for (int k = tid; k < 4 * size; k += nth) childCnt[k] = 0;
        if (tid == 0) { *s_nexp = 0; *s_K = 0x7fffffff; }
        __syncthreads();
        // (four keys per thread and trip: the global loads of a trip are issued together, so a pass pays the L2 latency
        // of the streamed key arrays once per four keys instead of once per key)
        for (int i0 = tid; i0 < n; i0 += 4 * nth) {
            int nd[4]; unsigned kxy[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * nth;
                nd[u] = i < n ? (int)nodeOfKey[i] : -1;
                kxy[u] = i < n ? *reinterpret_cast<const unsigned*>(keys + i) : 0u;      // x | y << 16
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (nd[u] >= 0) {
                    const QNode q = cur[nd[u]];
                    if (q.count > 1) atomicAdd(&childCnt[nd[u] * 4 + quadrant_of(q, (int)(kxy[u] & 0xffffu), (int)(kxy[u] >> 16))], 1);
                }
        }
        __syncthreads();

        int C = 0, U = 0;
        if (!phase2) {
            // every non-leaf node is split, in list order (:568-627)
            for (int base = 0; base < size; base += nth) {
                const int nd = base + tid;
                const bool valid = nd < size;
                const bool split = valid && cur[nd].count > 1;
                int ne = 0;
                if (split) {
                    const int4 cc = *reinterpret_cast<const int4*>(childCnt + nd * 4);
                    ne = (cc.x > 0) + (cc.y > 0) + (cc.z > 0) + (cc.w > 0);
                }
                // (one scan for both counts: per call they stay below 4 x 1024 and 1024, so they share a word)
                int tot2;
                const int ex2 = block_excl_scan(ne | ((int)(valid && !split) << 16), &tot2, s_scan);
                const int exNe = ex2 & 0xffff, exUn = ex2 >> 16, totNe = tot2 & 0xffff, totUn = tot2 >> 16;
                if (valid) aux[nd] = split ? (C + exNe) : -(U + exUn) - 1;
                C += totNe; U += totUn;
            }
        } else {
            // "largest first" (:635-700): candidates = non-leaf nodes in creation order (= reverse list
            // order), std::sort by (count, UL.x), processed from the back until size >= N.
            int m = 0;
            for (int base = 0; base < size; base += nth) {
                const int ridx = base + tid;
                const int nd = size - 1 - ridx;
                const bool split = ridx < size && cur[nd].count > 1;
                int tot;
                const int ex = block_excl_scan(split, &tot, s_scan);
                if (split) {
                    const QNode q = cur[nd];
                    const unsigned key = (unsigned)q.count * 8192u + (unsigned)q.x0;
                    const unsigned long long rec = ((unsigned long long)key << 32) | (unsigned)nd;
                    if (m + ex < sortSmemCap) s_sort[m + ex] = rec; else sortRec[m + ex] = rec;
                }
                m += tot;
            }
            __syncthreads();
            unsigned long long* sv = s_sort;
            if (m > sortSmemCap) {   // does not fit: sort in global memory (slow path, still exact)
                for (int k = tid; k < min(m, sortSmemCap); k += nth) sortRec[k] = s_sort[k];
                sv = sortRec;
                __syncthreads();
            }
            if (m <= sortSmemCap) {     // the CTA produces std::sort's exact permutation (stdsort_warp.cuh)
                unsigned short* sa = reinterpret_cast<unsigned short*>(s_sort + sortSmemCap);
                unsigned short* sb = sa + sortSmemCap;
                unsigned* bits = reinterpret_cast<unsigned*>(sb + sortSmemCap);
                int* rng = reinterpret_cast<int*>(bits + ((sortSmemCap + 31) >> 5));
                stdsort::sort_block(sv, m, sa, sb, bits, rng);
            } else if (tid == 0) {
                stdsort::sort(sv, m);
            }
            __syncthreads();
            // break point: first processed parent after which size >= N
            int runD = 0;
            for (int base = 0; base < m; base += nth) {
                const int pp = base + tid;
                int d = 0;
                if (pp < m) {
                    const int nd = (int)(sv[m - 1 - pp] & 0xffffffffu);
                    const int4 cc = *reinterpret_cast<const int4*>(childCnt + nd * 4);
                    d = (cc.x > 0) + (cc.y > 0) + (cc.z > 0) + (cc.w > 0) - 1;
                }
                int tot;
                const int ex = block_excl_scan(d, &tot, s_scan);
                if (pp < m && size + runD + ex + d >= N) atomicMin(s_K, pp + 1);
                runD += tot;
            }
            __syncthreads();
            const int K = min(*s_K, m);
            for (int nd = tid; nd < size; nd += nth) aux[nd] = -0x40000000;   // "not split" marker
            __syncthreads();
            for (int base = 0; base < K; base += nth) {
                const int pp = base + tid;
                int ne = 0, nd = 0;
                if (pp < K) {
                    nd = (int)(sv[m - 1 - pp] & 0xffffffffu);
                    const int4 cc = *reinterpret_cast<const int4*>(childCnt + nd * 4);
                    ne = (cc.x > 0) + (cc.y > 0) + (cc.z > 0) + (cc.w > 0);
                }
                int tot;
                const int ex = block_excl_scan(ne, &tot, s_scan);
                if (pp < K) aux[nd] = C + ex;
                C += tot;
            }
            __syncthreads();
            for (int base = 0; base < size; base += nth) {
                const int nd = base + tid;
                const bool un = nd < size && aux[nd] < 0;
                int tot;
                const int ex = block_excl_scan(un, &tot, s_scan);
                if (un) aux[nd] = -(U + ex) - 1;
                U += tot;
            }
        }
        __syncthreads();
        if (C + U > L.nodeCap) {     // cannot happen (size <= max(4*nIni, N+3)); fail loudly if it does
            if (tid == 0) { atomicOr(p.err, ERR_NODE_OVERFLOW); p.levelCount[f * g.nlevels + l] = 0; }
            return;
        }
        // rebuild the list
        int nexp = 0;
        for (int nd = tid; nd < size; nd += nth) {
            const int a = aux[nd];
            const QNode q = cur[nd];
            if (a >= 0) {
                const int mx = q.x0 + ((q.x1 - q.x0 + 1) >> 1), my = q.y0 + ((q.y1 - q.y0 + 1) >> 1);
                int cidx = a;
                const int4 cc = *reinterpret_cast<const int4*>(childCnt + nd * 4);
                const int cnts[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int cnt = cnts[k];
                    if (cnt > 0) {
                        QNode ch;
                        ch.x0 = (k & 1) ? (short)mx : q.x0;  ch.x1 = (k & 1) ? q.x1 : (short)mx;
                        ch.y0 = (k & 2) ? (short)my : q.y0;  ch.y1 = (k & 2) ? q.y1 : (short)my;
                        ch.count = cnt;
                        nxt[C - 1 - cidx] = ch;
                        ++cidx;
                        nexp += cnt > 1;
                    }
                }
            } else {
                nxt[C + (-a - 1)] = q;
            }
        }
        if (nexp) atomicAdd(s_nexp, nexp);
        for (int i0 = tid; i0 < n; i0 += 4 * nth) {
            int nd[4]; unsigned kxy[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * nth;
                nd[u] = i < n ? (int)nodeOfKey[i] : -1;
                kxy[u] = i < n ? *reinterpret_cast<const unsigned*>(keys + i) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (nd[u] >= 0) {
                    const int a = aux[nd[u]];
                    int to;
                    if (a >= 0) {
                        const int k = quadrant_of(cur[nd[u]], (int)(kxy[u] & 0xffffu), (int)(kxy[u] >> 16));
                        const int4 cc = *reinterpret_cast<const int4*>(childCnt + nd[u] * 4);      // one load instead of k dependent ones
                        to = C - 1 - (a + (k > 0 && cc.x > 0) + (k > 1 && cc.y > 0) + (k > 2 && cc.z > 0));
                    } else {
                        to = C + (-a - 1);
                    }
                    nodeOfKey[i0 + u * nth] = (unsigned short)to;
                }
        }
        __syncthreads();
        { QNode* t = cur; cur = nxt; nxt = t; }
        size = C + U;
        if (size >= N || size == prevSize) finish = true;
        else if (!phase2 && size + 3 * *s_nexp > N) phase2 = true;
        __syncthreads();   // s_nexp is reset at the top of the next pass
    }

    // (d) best response per node, first key wins ties (:703-722); add the border (:829-836).  One 32-bit atomicMax:
    // score (<= 255) in the top byte, 0xFFFFFF - key index below it (n < 2^19).
    for (int nd = tid; nd < size; nd += nth) best[nd] = 0u;
    __syncthreads();
    for (int i = tid; i < n; i += nth)
        atomicMax(&best[nodeOfKey[i]], ((unsigned)keys[i].score << 24) | (0xffffffu - (unsigned)i));
    __syncthreads();
    for (int nd = tid; nd < size; nd += nth) {
        const unsigned i = 0xffffffu - (best[nd] & 0xffffffu);
        const RawRec k = keys[i];
        LevelKp o;
        o.x = (unsigned short)(k.x + kMinBorder); o.y = (unsigned short)(k.y + kMinBorder);
        o.score = k.score; o.pad = 0; o.dst = -1; o.angle = -1.f;
        lkp[nd] = o;
    }
    if (tid == 0) p.levelCount[f * g.nlevels + l] = size;
}

// Shared-memory table sizes of quadtree_kernel (nodeCap 0: the node tables stay in global memory)
struct QtSmem { int sortCap, nodeCap, orderInside; };

__device__ __forceinline__ void order_frame(const FrameGeom& g, const DevPtrs& p, const int f, BlockScan& s_scan);      // K7, below

template <int NT>      // threads per CTA: 256 when the batch fills the GPU, 1024 for small batches (single-camera latency; one CTA
                       // per SM is enough there, so it may use 64 registers: the default bound of 32 spilled 616 bytes of loads)
#ifndef NAV24_QT_MINB
#define NAV24_QT_MINB 5
#endif
__global__ void __launch_bounds__(NT, NT >= 512 ? 1 : NAV24_QT_MINB) quadtree_kernel(const __grid_constant__ FrameGeom g, const DevPtrs p, const QtSmem qs) {
    extern __shared__ __align__(16) unsigned long long s_sort[];
    __shared__ int s_scanBuf[64];
    BlockScan s_scan{s_scanBuf, 0};
    __shared__ int s_K, s_nexp;

    const int tid = threadIdx.x, nth = NT;
    // (programmatic dependent launch: resident early, ordered behind the FAST kernel here)
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // grid = (frames, levels): CTAs are handed out level by level, i.e. the long ones first (level 0 has 13 x the keys of level
    // 7), so that the kernel does not end on a few level-0 CTAs that started late
    const int l = blockIdx.y, f = blockIdx.x;
    const LevelGeom& L = g.lv[l];
    const long long fr = (long long)f * g.rawPerFrame + L.rawOff;
    const long long fn = (long long)f * g.nodesPerFrame + L.nodeOff;
    const uint2* cinfo = p.cellInfo + (long long)f * g.totalCells + L.cellBase;
    int* cdst = p.cellDst + (long long)f * g.totalCells + L.cellBase;
    const RawRec* raw = p.raw + fr;
    RawRec* gkeys = p.keys + fr;

    // (a) bring the per-cell lists into the reference order: cells row-major, row-major inside a cell
    const int nC = L.nCols * L.nRows;
    // the cell table goes to shared memory (in the sort scratch, which is idle until the "largest first" phase) when it fits
    const bool cellsInSm = (size_t)nC * 12 <= quadtree_sort_bytes(qs.sortCap);
    uint2* s_ci = reinterpret_cast<uint2*>(s_sort);
    int* s_dst = reinterpret_cast<int*>(s_ci + nC);
    int n = 0;
    for (int base = 0; base < nC; base += nth) {
        const int c = base + tid;
        const uint2 ci = c < nC ? cinfo[c] : make_uint2(0u, 0u);
        int tot;
        const int ex = block_excl_scan((int)ci.y, &tot, s_scan);
        if (c < nC) {
            if (cellsInSm) { s_ci[c] = ci; s_dst[c] = n + ex; }
            else cdst[c] = n + ex;
        }
        n += tot;
    }
    __syncthreads();
    // shared-memory node tables behind the sort scratch: child counts (16-byte aligned) | node lists A, B | aux | best
    unsigned char* tb = reinterpret_cast<unsigned char*>(s_sort) + quadtree_sort_bytes(qs.sortCap);
    int* sChild = reinterpret_cast<int*>(tb);
    QNode* sA = reinterpret_cast<QNode*>(sChild + 4 * qs.nodeCap);
    QNode* sB = sA + qs.nodeCap;
    int* sAux = reinterpret_cast<int*>(sB + qs.nodeCap);
    unsigned* sBest = reinterpret_cast<unsigned*>(sAux + qs.nodeCap);
    const bool useSm = L.nodeCap <= qs.nodeCap;      // CTA-uniform
    if (tid == 0) p.rawTotal[f * g.nlevels + l] = n;
    __syncthreads();
    unsigned short* nok = reinterpret_cast<unsigned short*>(p.nodeOfKey + fr);
    if (n == 0) {
        if (tid == 0) p.levelCount[f * g.nlevels + l] = 0;
    } else if (useSm) {
        quadtree_run<NT, true>(g, p, l, f, n, qs.sortCap, s_sort, s_scan, &s_K, &s_nexp, gkeys, nok, sA, sB, sChild, sAux, sBest, raw, cinfo, cdst, nC, gkeys,
                               cellsInSm ? s_ci : nullptr, s_dst);
    } else {
        quadtree_run<NT, false>(g, p, l, f, n, qs.sortCap, s_sort, s_scan, &s_K, &s_nexp, gkeys, nok, p.nodesA + fn, p.nodesB + fn,
                                p.childCnt + 4 * fn, p.nodeAux + fn, reinterpret_cast<unsigned*>(p.best + fn), raw, cinfo, cdst, nC, gkeys,
                                cellsInSm ? s_ci : nullptr, s_dst);
    }
    // small batches: the CTA that finishes last among the levels of frame f puts the frame's keypoints into the output order
    if (!qs.orderInside) return;
    __threadfence();                             // this level's keypoints and count are visible before the counter moves
    __syncthreads();
    if (tid == 0) {
        const int done = atomicAdd(p.frameDone + f, 1);
        s_K = done == g.nlevels - 1;
        if (s_K) p.frameDone[f] = 0;             // ready for the next launch (nobody else touches it any more)
    }
    __syncthreads();
    if (s_K) {
        __threadfence();
        order_frame(g, p, f, s_scan);
    }
}

// ------------------------------------------------------------------------------------------
// K7  output order of one frame, by one CTA.  Reproduces the two-ended placement of detect()
// (OP_FtDtOrbSlam.cpp:877-921): keypoints with 0 <= x*scale <= 1000 fill the output from the back,
// the others from the front; monoIndex = number of the latter.
// Small batches (single-camera latency): runs at the end of quadtree_kernel in the CTA that finishes LAST among the levels
// of its frame (a counter per frame, release / acquire by __threadfence around the atomic) — no separate launch, and the
// step, a chain of block scans, overlaps the quadtree CTAs still running (single 4K frame: 0.288 -> 0.267 ms).  Batches
// that fill the GPU run it as order_kernel, one CTA per frame: there the tail inside the quadtree CTAs costs more than the
// launch (0.107 vs 0.07 ms per 1024 KITTI frames).  The level tables may have been written by other CTAs of the same
// launch, so they are read with ld.global.cg (L2), never through L1.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void order_frame(const FrameGeom& g, const DevPtrs& p, const int f, BlockScan& s_scan) {
    const int tid = threadIdx.x, nth = blockDim.x;
    int n = 0;
    for (int l = 0; l < g.nlevels; ++l) n += __ldcg(p.levelCount + f * g.nlevels + l);
    int nSt = 0, nMo = 0;
    for (int l = 0; l < g.nlevels; ++l) {
        const int cnt = __ldcg(p.levelCount + f * g.nlevels + l);
        LevelKp* lkp = p.lkp + (long long)f * g.kpPerFrame + g.lv[l].kpOff;
        const float sc = g.lv[l].scale;
        for (int base = 0; base < cnt; base += nth) {
            const int i = base + tid;
            bool st = false, mo = false;
            if (i < cnt) {
                const float x = (float)(__ldcg(reinterpret_cast<const unsigned*>(lkp + i)) & 0xffffu);      // LevelKp::x
                const float xs = l ? __fmul_rn(x, sc) : x;
                st = xs >= 0.f && xs <= 1000.f;
                mo = !st;
            }
            int t2;      // (one scan for both counts: each stays below 1024 per call, so they share a word)
            const int e2 = block_excl_scan((int)st | ((int)mo << 16), &t2, s_scan);
            const int eS = e2 & 0xffff, eM = e2 >> 16, tS = t2 & 0xffff, tM = t2 >> 16;
            if (i < cnt) lkp[i].dst = st ? n - 1 - (nSt + eS) : nMo + eM;
            nSt += tS; nMo += tM;
        }
    }
    if (tid == 0) {
        if (n > g.outCap) { atomicOr(p.err, ERR_KP_OVERFLOW); n = 0; nMo = 0; }
        p.nOut[f] = n; p.monoOut[f] = nMo;
    }
}

__global__ void __launch_bounds__(256) order_kernel(const __grid_constant__ FrameGeom g, const DevPtrs p) {
    __shared__ int s_scanBuf[64];
    BlockScan s_scan{s_scanBuf, 0};
    order_frame(g, p, blockIdx.x, s_scan);
}

// ------------------------------------------------------------------------------------------
// K5  7x7 Gaussian blur, sigma 2, BORDER_REFLECT_101: exact integer kernel [18 34 48 56 48 34 18]
// twice, (v + 32768) >> 16; replaces GaussianBlur at OP_FtDtOrbSlam.cpp:890-891.
// 64x16 output tile per CTA, horizontal pass into shared memory as u16 (max 255*256 fits).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int v, int n) {
    if (n == 1) return 0;
    while (v < 0 || v >= n) v = v < 0 ? -v : 2 * n - 2 - v;
    return v;
}

// One CTA owns a 128-pixel-wide, kBlurCtaRows-high tile whose source pixels (3-px halo, x start rounded down to 16 bytes)
// arrive as ONE TMA box; out-of-image bytes come back as zeros and BORDER_REFLECT_101 is restored by patching the few
// halo columns / rows of the tile in shared memory (uniform branches, edge tiles only), so the main loop has a single
// branch-free body.  One WARP owns kBlurTileRows rows of the tile, lane i the pixels [4i, 4i+4) of every row: per row
// three words from shared memory (immediate offsets), the four horizontal 7-tap sums with two DP4A each
// (coefficients (18,34,48,56 | 48,34,18,0)), the vertical pass on pairs of rows kept in a register ring (see below), one
// 32-bit store.
// All levels and frames in one launch; the CTA tiles of a frame are numbered level by level (blurTileBase).
__global__ void NAV24_BLUR_LB blur_kernel(const __grid_constant__ FrameGeom g, const DevPtrs p,
                                                   const __grid_constant__ TmaMaps maps) {
    __shared__ __align__(128) uint8_t tile[kBlurBoxW * kBlurBoxH];
    __shared__ __align__(8) unsigned long long bar;
    constexpr int P = kBlurBoxW;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int t = blockIdx.x, f = blockIdx.y;
    int l = 0;
    while (l + 1 < g.nlevels && t >= g.lv[l + 1].blurTileBase) ++l;
    const LevelGeom& L = g.lv[l];
    const int w = L.w, h = L.h;
    const int tilesX = (w + 127) >> 7;
    const int tt = t - L.blurTileBase;
    const int ty = tt / tilesX, tx = tt - ty * tilesX;
    const int tileX = tx * 128, rowBase = ty * kBlurCtaRows;
    const unsigned barAddr = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barAddr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr), "r"((unsigned)(P * kBlurBoxH)) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                smem_u32(tile)),
            "l"(&maps.m[l]), "r"(tileX - 16), "r"(rowBase - 3), "r"(f + p.frameBase), "r"(barAddr)
            : "memory");
    }
    {
        unsigned done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(barAddr), "r"(0u)
                : "memory");
        }
    }
    // BORDER_REFLECT_101.  Tile column of image column x: x - tileX + 16; tile row of image row y: y - rowBase + 3.
    const int rowsIn = min(h - rowBase, kBlurCtaRows + 3) + min(rowBase, 3);      // tile rows that hold image rows ...
    const int rFirst = rowBase == 0 ? 3 : 0;                                      // ... starting at this tile row
    const bool leftEdge = tx == 0, rightEdge = w < tileX + 131;
    if (leftEdge || rightEdge) {
        for (int r = threadIdx.x; r < rowsIn; r += 128) {
            uint8_t* row = tile + (rFirst + r) * P + 16 - tileX;                  // row[x] = image column x
            if (leftEdge) { row[-1] = row[1]; row[-2] = row[2]; row[-3] = row[3]; }
            if (rightEdge)
                for (int c = w; c < min(tileX + 131, w + 3); ++c) row[c] = row[2 * (w - 1) - c];
        }
        __syncthreads();
    }
    if (rowBase == 0) {                                                           // rows -1, -2, -3 = rows 1, 2, 3
        for (int i = threadIdx.x; i < 3 * (P / 4); i += 128) {
            const int k = i / (P / 4), c = i - k * (P / 4);
            reinterpret_cast<unsigned*>(tile + (2 - k) * P)[c] = reinterpret_cast<const unsigned*>(tile + (4 + k) * P)[c];
        }
    }
    if (h - rowBase < kBlurCtaRows + 3) {                                          // rows h, h+1, h+2 = rows h-2, h-3, h-4
        const int rh = h - rowBase + 3;                                           // tile row of image row h
        for (int i = threadIdx.x; i < 3 * (P / 4); i += 128) {
            const int k = i / (P / 4), c = i - k * (P / 4);
            if (rh + k < kBlurBoxH)
                reinterpret_cast<unsigned*>(tile + (rh + k) * P)[c] = reinterpret_cast<const unsigned*>(tile + (rh - 2 - k) * P)[c];
        }
    }
    __syncthreads();

    const int y0 = rowBase + wid * kBlurTileRows, yEnd = min(y0 + kBlurTileRows, h);
    const int x4 = tileX + lane * 4;
    if (y0 >= h) return;
    const bool active = x4 < w;
    uint8_t* dp = p.blur + (long long)f * g.blurFrameBytes + L.boff + (long long)y0 * L.pitch + x4;
    const int dPitch = L.pitch;
    unsigned half;
    asm("mov.u32 %0, 32768;" : "=r"(half));      // kept in a register: IMAD has one immediate slot
    // word 0 of the lane's window (pixels x4-4 .. x4-1) in the row that enters the ring first (image row y0 - 3)
    const uint8_t* rp = tile + (y0 - rowBase) * P + 12 + lane * 4;
    // Vertical pass on PAIRS of rows: the horizontal sums fit 16 bits (<= 255 * 256), so the sums of two consecutive rows
    // of one pixel share a register (one PRMT when the lower row arrives) and weigh in with ONE DP2A:
    //   out(y) = DP2A(pair(y-3, y-2), (18, 34)) + DP2A(pair(y-1, y), (48, 56)) + DP2A(pair(y+1, y+2), (48, 34)) + 18 * row(y+3)
    // i.e. 1 PRMT + 3 DP2A + 1 IMAD per pixel instead of 3 IADD + 4 IMAD.  Every pair of consecutive rows is used (by
    // the outputs of its own parity): ring of six pairs, pair j = rows (j, j+1) of the strip in slot j % 6, plus the last
    // two rows of sums (alternating, so that "the previous row" needs no register moves).
    unsigned pk[6][4], ob[2][4];
    // horizontal sums of tile row *rp
    auto hrow = [&](unsigned* o) {
        const unsigned w0 = *reinterpret_cast<const unsigned*>(rp), w1 = *reinterpret_cast<const unsigned*>(rp + 4),
                       w2 = *reinterpret_cast<const unsigned*>(rp + 8);
        rp += P;
        const unsigned K1 = 0x38302212u, K2 = 0x00122230u;      // (18,34,48,56) and (48,34,18,0), little endian
        // (weighing the three ALIGNED words with per-pixel coefficient words instead — 10 DP4A, no PRMT — measured slower:
        // the ten constants do not fit the instruction and cost a uniform move each)
        o[0] = __dp4a(__byte_perm(w0, w1, 0x4321), K1, __dp4a(__byte_perm(w1, w2, 0x4321), K2, 0u));
        o[1] = __dp4a(__byte_perm(w0, w1, 0x5432), K1, __dp4a(__byte_perm(w1, w2, 0x5432), K2, 0u));
        o[2] = __dp4a(__byte_perm(w0, w1, 0x6543), K1, __dp4a(__byte_perm(w1, w2, 0x6543), K2, 0u));
        o[3] = __dp4a(w1, K1, __dp4a(w2, K2, 0u));
    };
    auto pack = [&](int slot, const unsigned* lo, const unsigned* hi) {
#pragma unroll
        for (int i = 0; i < 4; ++i) pk[slot][i] = __byte_perm(lo[i], hi[i], 0x5410);
    };
    // the output row whose window ends with row `o`; m = (index of that row inside the strip) % 6
    auto emit = [&](int m, const unsigned* o) {
        const unsigned K01 = 0x2212u, K23 = 0x3830u, K45 = 0x2230u;      // (18,34), (48,56), (48,34)
        unsigned v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            v[i] = __dp2a_lo(pk[m % 6][i], K01, __dp2a_lo(pk[(m + 2) % 6][i], K23, __dp2a_lo(pk[(m + 4) % 6][i], K45, 18u * o[i] + half)));
        const unsigned word = __byte_perm(__byte_perm(v[0], v[1], 0x0062), __byte_perm(v[2], v[3], 0x0062), 0x5410);
        if (active) *reinterpret_cast<unsigned*>(dp) = word;
        dp += dPitch;
    };
    // six rows fill the window, then every row emits; full groups of six rows run without any per-row test
    hrow(ob[0]);
#pragma unroll
    for (int j = 1; j < 6; ++j) { hrow(ob[j & 1]); pack(j - 1, ob[(j - 1) & 1], ob[j & 1]); }
    const int nrows = yEnd - y0;
    int done = 0;
    for (; done + 6 <= nrows; done += 6) {
#pragma unroll
        for (int m = 0; m < 6; ++m) { hrow(ob[m & 1]); emit(m, ob[m & 1]); pack((m + 5) % 6, ob[(m + 1) & 1], ob[m & 1]); }
    }
#pragma unroll
    for (int m = 0; m < 6; ++m)
        if (done + m < nrows) { hrow(ob[m & 1]); emit(m, ob[m & 1]); pack((m + 5) % 6, ob[(m + 1) & 1], ob[m & 1]); }      // uniform: the bottom tile of a level
}

// ------------------------------------------------------------------------------------------
// K4+K6  orientation and rBRIEF-256, one warp per keypoint.
//   IC_Angle (OP_FtDtOrbSlam.cpp:18-44): integer moments over the 31-px disc of the un-blurred level,
//   angle = cv::fastAtan2 with un-contracted f32 arithmetic.
//   computeOrbDescriptor (:48-87): 256 rotated pair comparisons on the blurred level; lane i
//   produces descriptor byte i.  Writes the final nav24_kp (coordinates scaled to level 0, :905-907)
//   and the descriptor at the keypoint's slot of the two-ended output order.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// rBRIEF pattern as floats, transposed for the warp: entry [j][lane] = (x0, y0, x1, y1) of test pair 8*lane + j, so that
// the 32 lanes of one load read 512 consecutive bytes (lane-major order cost 32 L1 wavefronts per load instead of 4).
struct PatternT { float v[8][32][4]; };
constexpr PatternT make_pattern_t() {
    constexpr int src[1024] = {
#include "pattern_31.inc"
    };
    PatternT t{};
    for (int lane = 0; lane < 32; ++lane)
        for (int j = 0; j < 8; ++j)
            for (int c = 0; c < 4; ++c) t.v[j][lane][c] = (float)src[(lane * 8 + j) * 4 + c];
    return t;
}
__device__ const PatternT kPatternT = make_pattern_t();

// Orientation weights for the DP4A moments, [align 4][row 31][word 9][2]: for the aligned words that cover the 31-px
// row segment of a keypoint whose left end sits `align` bytes into its first word, (x) the four signed u offsets of the
// bytes inside the disc (0 outside), (y) their 0/1 mask.  m10 = sum dp4a(word, x), m01 = sum v * dp4a(word, y).
// Built once per context on the host (capi.cu: build_orientation_table) from umax[] (OP_FtDtOrbSlam.cpp:484-499).
//
// A warp owns kDescSlots consecutive keypoint slots of one frame and works in two passes over its live keypoints, so that
// everything that is ONE value per keypoint is computed by ONE LANE per keypoint, for all of the warp's keypoints at once,
// instead of redundantly by 32 lanes keypoint after keypoint (the first version spent 28 % of its instructions that way):
//   set-up  lane i < kDescSlots: level, liveness and record of slot i (the dead slots are the tail of every level's range)
//   pass 1  per keypoint: the 31-row patch of the un-blurred level arrives as a TMA box; DP4A moments of the disc (279 aligned
//           words, 9 per lane), two REDUX sums; lane i keeps (m01, m10) of keypoint i
//   between lane i: fastAtan2, sincosf, the nav24_kp record and the angle of keypoint i
//   pass 2  per keypoint: the 37-row patch of the blurred level arrives as a TMA box; (cos, sin) come from lane i by SHFL;
//           512 rotated samples, lane j produces descriptor byte j
// The boxes (x start rounded down to 16 bytes, hence the 48- and 80-byte box widths) land in a ring of kDescStages stages
// per warp, each holding one descriptor box or two orientation boxes: while a keypoint is computed, the boxes of the next
// kDescStages - 1 (pass 2) or 2 * kDescStages - 1 (pass 1, whose items are short) keypoints are in flight.  Warps never
// synchronise with each other (a CTA is four independent warps).
constexpr int kDescWarps = 4;                  // warps per CTA
// Measured on KITTI batches (this kernel, ms / 1024 frames; slots x stages): 16x2 1.40, 16x3 1.21, 16x4 1.34, 32x3 1.45,
// 8x3 1.18, 8x2 1.13, 10x2 1.11, 12x2 1.11 — warps in flight (8 CTAs of 4 warps per SM with a 6 KB ring each) matter more
// than ring depth: the kernel waits on box arrivals and L2 (67 % busy), not on issue slots.
#ifndef NAV24_DESC_SLOTS
#define NAV24_DESC_SLOTS 10
#endif
constexpr int kDescSlots = NAV24_DESC_SLOTS;   // keypoint slots per warp
#ifndef NAV24_DESC_STAGES
#define NAV24_DESC_STAGES 2
#endif
constexpr int kDescStages = NAV24_DESC_STAGES;
constexpr int kDescStage = (kDescBoxW * kDescBoxH + 255) / 256 * 256;      // bytes per stage
constexpr int kOriSlots = 2 * kDescStages;     // orientation boxes in the ring
static_assert(2 * ((kOriBoxW * kOriBoxH + 127) / 128 * 128) <= kDescStage, "one stage holds two orientation boxes");

// (no register bound: with the 64 registers ptxas picks for a plain 128-thread bound the kernel is 9 % slower than with the
// 72 it takes when left alone — the loads of a keypoint's test pairs are batched further ahead)
__global__ void __launch_bounds__(kDescWarps * 32, 1) describe_kernel(const __grid_constant__ FrameGeom g, const DevPtrs p,
                                                                   const __grid_constant__ TmaMaps mapsOri,
                                                                   const __grid_constant__ TmaMaps mapsBlur,
                                                                   const __grid_constant__ TmaMaps mapsBlurN) {
    extern __shared__ __align__(128) uint8_t s_desc[];      // [kDescWarps][kDescStages * kDescStage] box ring of every warp
    __shared__ __align__(8) unsigned long long s_bar[kDescWarps][kOriSlots + kDescStages];      // orientation slots, then descriptor stages
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int f = blockIdx.y;
    const int slot0 = (blockIdx.x * kDescWarps + wid) * kDescSlots;
    if (slot0 >= g.kpPerFrame) return;
    // set-up: lane i looks at slot slot0 + i
    LevelKp* lkp = p.lkp + (long long)f * g.kpPerFrame;
    int myL = 0;
    bool live = false;
    unsigned myXY = 0;
    {
        const int s = slot0 + lane;
        if (lane < kDescSlots && s < g.kpPerFrame) {
            for (int l = 1; l < g.nlevels; ++l)
                if (g.lv[l].kpOff <= s) myL = l;
            live = s - g.lv[myL].kpOff < p.levelCount[f * g.nlevels + myL];
            if (live) myXY = *reinterpret_cast<const unsigned*>(lkp + s);        // x | y << 16
        }
    }
    const unsigned liveMask = __ballot_sync(0xffffffffu, live);
    if (liveMask == 0u) return;
    const int nLive = __popc(liveMask);
    uint8_t* const s_ring = s_desc + wid * (kDescStages * kDescStage);
    const unsigned bar0 = smem_u32(&s_bar[wid][0]), buf0 = smem_u32(s_ring);
    if (lane < kOriSlots + kDescStages) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * lane));
    if (lane == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    // lane 0: the orientation box of a keypoint -> orientation slot i, or its descriptor box -> stage i.  The box start must
    // be 16-byte aligned in x (an unaligned start coordinate of a u8 tensor faults), so a 37-px patch needs 37 + slack
    // columns, slack = (cx - 18) & 15: the narrow 48-byte box when slack <= 11 (3 keypoints of 4), the 80-byte box
    // otherwise.  Both row pitches cost 3.1 shared-memory wavefronts per gather of the rotated pattern (a 64-byte pitch: 4.8).
    auto issue = [&](bool ori, int l, int cx, int cy, int i) {      // (called by lane 0 only)
        if (ori) {
            const unsigned bar = bar0 + 8 * i, dst = buf0 + (unsigned)i * (kDescStage / 2);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(kOriBoxW * kOriBoxH)) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                "l"(&mapsOri.m[l]), "r"((cx - 15) & ~15), "r"(cy - 15), "r"(f + p.frameBase), "r"(bar)
                : "memory");
        } else {
            const unsigned bar = bar0 + 8 * (kOriSlots + i), dst = buf0 + (unsigned)i * kDescStage;
            const bool narrow = ((cx - 18) & 15) <= kDescBoxWN - 37;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                         "r"((unsigned)((narrow ? kDescBoxWN : kDescBoxW) * kDescBoxH))
                         : "memory");
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                "l"(narrow ? &mapsBlurN.m[l] : &mapsBlur.m[l]), "r"((cx - 18) & ~15), "r"(cy - 18), "r"(f + p.frameBase), "r"(bar)
                : "memory");
        }
    };
    unsigned phase = 0;                          // bit i: parity of barrier i's next wait
    auto wait = [&](int i) {
        unsigned done = 0;
        const unsigned bar = bar0 + 8 * i, par = (phase >> i) & 1u;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar), "r"(par)
                : "memory");
        }
        phase ^= 1u << i;
    };
    // the next keypoint of `todo` (in slot order) gets its box issued into ring position i
    auto issue_next = [&](unsigned& todo, bool ori, int i) {
        const int kn = __ffs(todo) - 1;
        todo &= todo - 1;
        const unsigned xy = __shfl_sync(0xffffffffu, myXY, kn);
        const int l = __shfl_sync(0xffffffffu, myL, kn);
        if (lane == 0) issue(ori, l, (int)(xy & 0xffffu), (int)(xy >> 16), i);
    };
    // (keeping the lane's eight test pairs in registers across keypoints was measured slower: 96 registers per thread
    // cost more in occupancy than the 32 L1 wavefronts per keypoint cost in the load pipe)
    const float4* pat = reinterpret_cast<const float4*>(kPatternT.v) + lane;
    int wpos[9], wrow[9];                       // orientation: word k*32+lane of the 31 x 9-word patch -> (tile word, v)
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int idx = min(k * 32 + lane, 278);
        const int r = (idx * 57) >> 9, c = idx - r * 9;             // idx / 9 for idx < 288
        wpos[k] = r * (kOriBoxW / 4) + c;
        wrow[k] = r - 15;
    }

    int myM10 = 0, myM01 = 0;
    // ---- pass 1: moments ------------------------------------------------------------------------------------------
    unsigned todo = liveMask;                   // keypoints whose box is not yet issued
    int iss = 0;                                // boxes issued in this pass
    for (; iss < min(nLive, kOriSlots - 1); ++iss) issue_next(todo, true, iss);
    unsigned cur = liveMask;
    for (int t = 0; t < nLive; ++t) {
        const int k = __ffs(cur) - 1;           // the keypoint computed in this trip
        cur &= cur - 1;
        if (todo) {                             // (its ring position was released by keypoint t - 1)
            issue_next(todo, true, iss % kOriSlots);
            ++iss;
        }
        const int cx = (int)(__shfl_sync(0xffffffffu, myXY, k) & 0xffffu);
        const int sl = t % kOriSlots;
        // this lane's orientation weights while the box lands
        const int off = (cx - 15) & 15;                                            // 0..15
        const uint2* tab = reinterpret_cast<const uint2*>(p.oriTab) + (off & 3) * 279 + lane;
        const unsigned* ow = reinterpret_cast<const unsigned*>(s_ring + sl * (kDescStage / 2)) + (off >> 2);
        uint2 wt[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) wt[j] = __ldg(tab + min(j * 32, 278 - lane));
        wait(sl);
        // IC_Angle: integer moments of the 31-px disc by DP4A over the aligned words of the patch (279 words, 9 per lane;
        // the lane's word positions and row numbers do not depend on the keypoint: wpos / wrow, set up once per warp)
        int m10 = 0, m01 = 0;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            if (j * 32 + lane < 279) {
                const unsigned w = ow[wpos[j]];
                asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(m10) : "r"(w), "r"(wt[j].x));      // u8 pixels x s8 offsets
                m01 += wrow[j] * (int)__dp4a(w, wt[j].y, 0u);
            }
        }
        m10 = __reduce_add_sync(0xffffffffu, m10);
        m01 = __reduce_add_sync(0xffffffffu, m01);
        if (lane == k) { myM10 = m10; myM01 = m01; }
        __syncwarp();                               // every lane is done with the slot before it is refilled
    }
    // the first descriptor boxes fly while one lane per keypoint does the scalar work
    todo = liveMask;
    iss = 0;
    for (; iss < min(nLive, kDescStages - 1); ++iss) issue_next(todo, false, iss);
    // ---- between the passes: one lane per keypoint -------------------------------------------------------------
    float myA = 1.f, myB = 0.f;                 // cos, sin of the keypoint's angle
    int myDst = 0;
    if (live) {
        const float angle = fast_atan2_deg((float)myM01, (float)myM10);
        const float factorPI = (float)(3.14159265358979323846 / 180.f);
        sincosf(__fmul_rn(angle, factorPI), &myB, &myA);
        LevelKp* kp = lkp + slot0 + lane;
        myDst = kp->dst;
        kp->angle = angle;
        const LevelGeom& L = g.lv[myL];
        const int cx = (int)(myXY & 0xffffu), cy = (int)(myXY >> 16);
        nav24_kp o;
        o.x = myL ? __fmul_rn((float)cx, L.scale) : (float)cx;
        o.y = myL ? __fmul_rn((float)cy, L.scale) : (float)cy;
        o.size = L.patch; o.angle = angle; o.response = (float)kp->score; o.octave = myL; o.class_id = -1;
        p.outKp[(long long)f * g.outCap + myDst] = o;
    }
    // ---- pass 2: descriptors ---------------------------------------------------------------------------------------
    cur = liveMask;
    for (int t = 0; t < nLive; ++t) {
        const int k = __ffs(cur) - 1;
        cur &= cur - 1;
        if (todo) {
            issue_next(todo, false, iss % kDescStages);
            ++iss;
        }
        const int cx = (int)(__shfl_sync(0xffffffffu, myXY, k) & 0xffffu);
        const float a = __shfl_sync(0xffffffffu, myA, k), b = __shfl_sync(0xffffffffu, myB, k);
        const int dst = __shfl_sync(0xffffffffu, myDst, k);
        const int st = t % kDescStages;
        const uint8_t* s_blur = s_ring + st * kDescStage;
        wait(kOriSlots + st);
        // computeOrbDescriptor: the 512 sample points lie within +-18 px of the keypoint (pattern radius 18.38).
        // cvRound = round-half-even of an f32 in (-2^22, 2^22): adding 1.5 * 2^23 leaves the integer in the low mantissa
        // bits (float bits = 0x4B400000 + n), one full-rate FADD instead of a quarter-rate F2I per coordinate (1024 per
        // keypoint); the bias of row and column is folded into the base address.
        const float kMagic = 12582912.f;
        const unsigned bp = ((cx - 18) & 15) <= kDescBoxWN - 37 ? (unsigned)kDescBoxWN : (unsigned)kDescBoxW;      // row pitch of this keypoint's box
        const unsigned pcA = smem_u32(s_blur) + 18u * bp + (unsigned)(18 + ((cx - 18) & 15)) -
                             (bp + 1u) * 0x4B400000u;               // the keypoint, minus the biases
        unsigned val = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 q4 = __ldg(pat + j * 32);
            const unsigned r0 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(q4.x, b), __fmul_rn(q4.y, a)), kMagic));
            const unsigned q0 = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(q4.x, a), __fmul_rn(q4.y, b)), kMagic));
            const unsigned r1 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(q4.z, b), __fmul_rn(q4.w, a)), kMagic));
            const unsigned q1 = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(q4.z, a), __fmul_rn(q4.w, b)), kMagic));
            unsigned t0, t1;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t0) : "r"(pcA + r0 * bp + q0));
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t1) : "r"(pcA + r1 * bp + q1));
            val |= (unsigned)(t0 < t1) << j;
        }
        p.outDesc[((long long)f * g.outCap + dst) * 32 + lane] = (uint8_t)val;
        __syncwarp();                               // every lane is done with stage st before it is refilled
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
int launch_repack(const uint8_t* src, int w, int h, uint8_t* dst, int dPitch, int B, cudaStream_t s) {
    dim3 grid((dPitch / 16 + 63) / 64, (h + 3) / 4, B), block(64, 4);
    repack_kernel<<<grid, block, 0, s>>>(src, w, h, dst, dPitch);
    return 1;
}

int launch_bgr2gray(const uint8_t* src, int w, int h, uint8_t* dst, int dPitch, int B, cudaStream_t s) {
    dim3 grid((w + 255) / 256, (h + 3) / 4, B), block(64, 4);
    bgr2gray_kernel<<<grid, block, 0, s>>>(src, w, h, dst, dPitch);
    return 1;
}

// kernel launch with the programmatic-stream-serialization attribute: the kernel may start while its predecessor in the
// stream is still running and orders itself behind it with griddepcontrol.wait (works under stream capture too)
template <typename... KArgs, typename... Args>
static void launch_k(void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k, KArgs(std::forward<Args>(args))...);
}

int launch_pyramid(const FrameGeom& g, const DevPtrs& p, const ResizeTab* tabs, const TmaMaps& mapsSrc, int B, cudaStream_t s) {
    int n = 0;
    // (function attributes are per device and this library serves several devices and host threads: set on every call)
    cudaFuncSetAttribute(resize_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(resize_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(resize8_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int l = 1; l < g.nlevels; ++l) {
        const LevelGeom& D = g.lv[l];
        const ResizeTab& T = tabs[l];
        const size_t tile = (size_t)((T.boxW * T.boxH + 127) / 128 * 128);
        // programmatic dependent launch: levels 2.. may become resident while the previous level is still running and wait for
        // it inside the kernel (a single frame: the chain of seven launches is launch latency, 0.047 -> 0.037 ms; 1024 frames:
        // the tail of every launch is filled earlier, 0.74 -> 0.72 ms)
        const bool pdl = NAV24_PDL && l >= 2;
        if (T.wide) {      // eight pixels per thread: 256-column CTAs, two source boxes
            // a last column of <= 128 px would leave half of every warp of its CTAs idle: it goes to the four-pixel kernel
            // (128-column CTAs) instead — 16 % of the lanes of the 1.2 pyramid of a 1241-px frame were such idle halves
            const int rem = D.w % 256;
            // (batches only: a single frame pays more for the extra launches than for the idle lanes — 0.047 -> 0.071 ms)
            const bool split = NAV24_RS_SPLIT && B >= 16 && rem > 0 && rem <= 128 && T.boxW == 192 && D.w > 256;
            const int cols8 = split ? D.w / 256 : (D.w + 255) / 256;
            dim3 grid(cols8, (D.h + 4 * T.rows - 1) / (4 * T.rows), B);
            launch_k(resize8_kernel<192>, grid, dim3(128), 2 * tile, s, pdl, mapsSrc.m[l], p.frameBase, p.pyr + D.off, D.pitch,
                     g.pyrFrameBytes, split ? cols8 * 256 : D.w, D.h, T);
            if (split) {
                dim3 gridR(1, grid.y, B);
                launch_k(resize_kernel<192>, gridR, dim3(128), tile, s, NAV24_PDL != 0, mapsSrc.m[l], p.frameBase, p.pyr + D.off, D.pitch,
                         g.pyrFrameBytes, D.w, D.h, T, cols8 * 256);
                ++n;
            }
        } else {
            dim3 grid((D.w + 127) / 128, (D.h + 4 * T.rows - 1) / (4 * T.rows), B);
            if (T.boxW == 192)
                launch_k(resize_kernel<192>, grid, dim3(128), tile, s, pdl, mapsSrc.m[l], p.frameBase, p.pyr + D.off, D.pitch, g.pyrFrameBytes, D.w, D.h, T, 0);
            else
                launch_k(resize_kernel<256>, grid, dim3(128), tile, s, pdl, mapsSrc.m[l], p.frameBase, p.pyr + D.off, D.pitch, g.pyrFrameBytes, D.w, D.h, T, 0);
        }
        ++n;
    }
    return n;
}

// shared-memory layout of fast_band_kernel for the levels with minBoxH < boxH <= maxBoxH; returns the bytes (0: no such level)
int fast_smem_bytes(const FrameGeom& g, int minBoxH, int maxBoxH, FastSmem* out) {
    FastSmem sm{};
    int tileBytes = 0, mwords = 0, wcap = 0;
    const int qcap = kFastQueueCap;
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom& L = g.lv[l];
        if (L.boxH <= minBoxH || L.boxH > maxBoxH) continue;
        tileBytes = max(tileBytes, L.boxW * L.boxH + 64);
        mwords = max(mwords, ((L.segCols * L.wCell + 31) / 32) * L.hCell);
        wcap = max(wcap, L.segCols * ((L.wCell + 1) / 2) * ((L.hCell + 1) / 2));      // strict 3x3 maxima inside a cell: one per 2x2 block
    }
    if (tileBytes == 0) return 0;
    tileBytes = (tileBytes + 127) / 128 * 128;
    sm.offMap = tileBytes;
    sm.offMask = 2 * tileBytes;                                               // cleared together with the score map
    sm.offQueue = sm.offMask + (mwords * 4 + 15) / 16 * 16;
    sm.total = sm.offQueue + (max(qcap, wcap) * 2 + 15) / 16 * 16 + 16;      // (the dense path lists its keypoints in the queue)
    if (out) *out = sm;
    return sm.total;
}

// The tile and the score map are sized by the tallest cell row of the launch, and one level with tall cells (KITTI: 47-px
// cells at level 6, 39 elsewhere) would cost every CTA of the frame a resident CTA per SM (6 instead of 7).  The segment
// table therefore lists the levels with boxH <= g.fastCutH first (build_geometry picks the cut) and they are launched
// apart from the tall ones, each group with its own layout.
// (the raw-list counters are cleared by launch_fast_prepare BEFORE the pyramid launches, so that the FAST kernel follows the
// last resize kernel directly and can be launched as its programmatic dependent)
void launch_fast_prepare(const FrameGeom& g, const DevPtrs& p, int B, cudaStream_t s) {
    cudaMemsetAsync(p.rawCount, 0, sizeof(int) * (size_t)B * g.nlevels, s);
}

int launch_fast(const FrameGeom& g, const DevPtrs& p, const TmaMaps& maps, int B, int iniTh, int minTh, cudaStream_t s) {
    int n = 0;
    const bool grouped = g.segsLow < g.totalSegs && B >= 16;      // (a small batch does not fill the SMs: one launch)
    const int cutH = grouped ? g.fastCutH : 1 << 30, segsLow = grouped ? g.segsLow : g.totalSegs;
    const int lo[2] = {0, cutH}, hi[2] = {cutH, 1 << 30}, seg0[2] = {0, segsLow}, nSeg[2] = {segsLow, g.totalSegs - segsLow};
    int maxTotal = 0;
    FastSmem sm[2];
    for (int k = 0; k < 2; ++k) maxTotal = max(maxTotal, nSeg[k] > 0 ? fast_smem_bytes(g, lo[k], hi[k], &sm[k]) : 0);
    cudaFuncSetAttribute(fast_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max(maxTotal, 48 * 1024));
    for (int k = 0; k < 2; ++k) {
        if (nSeg[k] <= 0) continue;
        dim3 grid(nSeg[k], B);
        launch_k(fast_band_kernel, grid, dim3(kFastThreads), (size_t)sm[k].total, s, NAV24_PDL != 0, g, p, maps, sm[k], iniTh, minTh, seg0[k]);
        ++n;
    }
    return n;
}

// test hook: std::sort of n records (key = high 32 bits) by one warp, in shared memory
namespace {
__global__ void __launch_bounds__(256) debug_sort_kernel(unsigned long long* recs, int n, int cap) {
    extern __shared__ unsigned long long s_rec[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_rec[i] = recs[i];
    __syncthreads();
    unsigned short* sa = reinterpret_cast<unsigned short*>(s_rec + cap);
    unsigned short* sb = sa + cap;
    unsigned* bits = reinterpret_cast<unsigned*>(sb + cap);
    int* scratch = reinterpret_cast<int*>(bits + ((cap + 31) >> 5));
    if (blockDim.x == 32) stdsort::sort_warp(s_rec, n, sa, sb, bits, scratch);      // one warp, range stack
    else stdsort::sort_block(s_rec, n, sa, sb, bits, scratch);                      // the CTA, rounds of ranges (quadtree_kernel)
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) recs[i] = s_rec[i];
}
}  // namespace

int launch_debug_sort(unsigned long long* d_recs, int n, cudaStream_t s) {
    const int cap = (n + 3) & ~3;
    const size_t smem = quadtree_sort_bytes(cap);
    if (smem > 200 * 1024) return -1;
    cudaFuncSetAttribute(debug_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    debug_sort_kernel<<<1, (n & 1) ? 32 : 256, smem, s>>>(d_recs, n, cap);      // odd n: the one-warp version, even n: the CTA version
    return 1;
}

int launch_quadtree(const FrameGeom& g, const DevPtrs& p, int B, cudaStream_t s) {
    int maxNode = 0;
    for (int l = 0; l < g.nlevels; ++l) maxNode = max(maxNode, g.lv[l].nodeCap);
    QtSmem qs{};
    qs.sortCap = (min(maxNode, 12000) + 3) & ~3;      // records sorted in shared memory; larger levels sort in global memory
    size_t smem = quadtree_sort_bytes(qs.sortCap);
    // Node tables in shared memory (52 bytes per node) when those of the largest level fit the CTA's budget together with the
    // sort scratch.  A batch that fills the GPU (256-thread CTAs) gets 32 KB: seven CTAs per SM, the same occupancy as with
    // global tables (the kernel lives on overlapping barriers).  A small batch (1024-thread CTAs, single-camera latency) has
    // at most two CTAs per SM anyway: 100 KB each, or 200 KB when the whole grid is one wave of one CTA per SM — enough
    // for level 0 of a 4K frame (1744 nodes) or of the x5 feature mode.
    const int nCta = g.nlevels * B;
    // (measured: between one and two waves the 1024-thread CTAs win on 4K frames — 47 k keys at level 0 — and lose on
    // KITTI-sized ones, where seven 256-thread CTAs per SM overlap their barriers better)
    const bool big = nCta <= 148 || (nCta < 2 * 148 && (long long)g.lv[0].w * g.lv[0].h >= 1500000);
    const size_t budget = !big ? 32 * 1024 : (nCta <= 148 ? 200 * 1024 : 100 * 1024);
    const size_t nodeBytes = (size_t)maxNode * 52;
    if (smem + nodeBytes <= budget) {
        qs.nodeCap = maxNode;
        smem += nodeBytes;
    }
    cudaFuncSetAttribute(quadtree_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(quadtree_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(quadtree_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    dim3 grid(B, g.nlevels);
    // 256 threads per (level, frame) when the batch fills the GPU (128: 5 % faster alone, slower in the chunked host
    // pipeline); a small batch (single-camera latency) has few CTAs, so each gets 1024 threads for its parallel passes
    qs.orderInside = nCta < 2 * 148;
    // (small batches of megapixel-sized frames, a few thousand keys at level 0: 512 threads — the barriers are cheaper and the
    // parallel passes are one trip either way: single KITTI frame 0.065 -> 0.062 ms, EuRoC 0.055 -> 0.051; 4K frames with
    // 47 k keys at level 0 need the 1024: 0.264 vs 0.322 ms)
    if (big && (long long)g.lv[0].w * g.lv[0].h < 1500000) launch_k(quadtree_kernel<512>, grid, dim3(512), smem, s, NAV24_PDL != 0, g, p, qs);
    else if (big) launch_k(quadtree_kernel<1024>, grid, dim3(1024), smem, s, NAV24_PDL != 0, g, p, qs);
    else launch_k(quadtree_kernel<256>, grid, dim3(256), smem, s, NAV24_PDL != 0, g, p, qs);
    if (qs.orderInside) return 1;
    order_kernel<<<B, 256, 0, s>>>(g, p);
    return 2;
}

int launch_blur(const FrameGeom& g, const DevPtrs& p, const TmaMaps& mapsBlurSrc, int B, cudaStream_t s) {
    dim3 grid(g.blurTiles, B);
    blur_kernel<<<grid, 128, 0, s>>>(g, p, mapsBlurSrc);
    return 1;
}

int launch_describe(const FrameGeom& g, const DevPtrs& p, const TmaMaps& mapsOri, const TmaMaps& mapsBlur, const TmaMaps& mapsBlurN, int B,
                    cudaStream_t s) {
    dim3 grid((g.kpPerFrame + kDescWarps * kDescSlots - 1) / (kDescWarps * kDescSlots), B);
    constexpr int smem = kDescWarps * kDescStages * kDescStage;
    cudaFuncSetAttribute(describe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    describe_kernel<<<grid, kDescWarps * 32, smem, s>>>(g, p, mapsOri, mapsBlur, mapsBlurN);
    return 1;
}

}  // namespace nav24
