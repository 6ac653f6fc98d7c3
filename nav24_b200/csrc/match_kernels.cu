// match_kernels.cu — sm_100a kernels of the descriptor matchers (integer popc / warp-min; no
// tensor cores: Hamming matching is not a dense floating-point contraction).
//   window matcher  : FtAssocOrbSlam::matchV  core/operators/objAssoc/OP_FtAssocOrbSlam.cpp:91-223
//                     + FeatureGrid           core/sensorData/observation/FeatureGrid.cpp:20-152
//   brute force     : intended semantics of FtAssocOCV::match  core/operators/objAssoc/OP_FtAssoc.cpp:63-99
#include <algorithm>

#include "orb_internal.cuh"

namespace nav24 {
namespace {

constexpr int kCandCap = 32;       // pruned candidates kept per query (one warp round); more -> exact slow path
constexpr int kHisto = 30;         // HISTO_LENGTH (:15)

__device__ __forceinline__ int block_excl_scan_m(int v, int* total, int* s_warp) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = lane < nw ? s_warp[lane] : 0, wi = w;
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

__device__ __forceinline__ int hamming256(const uint4* a, const uint4* b) {
    const uint4 a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

struct PairView {
    const nav24_kp* k1; const float* ud1; const uint4* d1; int n1;
    const nav24_kp* k2; const float* ud2; const uint4* d2; int n2;
};

__device__ __forceinline__ float2 ud_of(const nav24_kp* k, const float* ud, int i) {
    return ud ? make_float2(ud[2 * i], ud[2 * i + 1]) : make_float2(k[i].x, k[i].y);
}

// Candidate cell range of FeatureGrid::getFeaturesInArea (FeatureGrid.cpp:38-60). false = empty.
__device__ __forceinline__ bool cell_range(const MatchArgs& a, float x, float y, int& cx0, int& cx1, int& cy0, int& cy1) {
    const float r = a.window;
    cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, a.grid.min_x), r), a.invW)));
    if (cx0 >= a.grid.cols) return false;
    cx1 = min(a.grid.cols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, a.grid.min_x), r), a.invW)));
    if (cx1 < 0) return false;
    cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, a.grid.min_y), r), a.invH)));
    if (cy0 >= a.grid.rows) return false;
    cy1 = min(a.grid.rows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, a.grid.min_y), r), a.invH)));
    if (cy1 < 0) return false;
    return true;
}

// One CTA (1024 threads) per frame pair.  Dynamic shared memory (sizes in MatchSmem; 0 = that table stays
// in global memory, the kernel is still exact): CSR cell starts | dist2 | m21 | active-query table | candidate pool.
struct MatchSmem { int cells, n2, act, pool, acc, total; };
constexpr int kAccSlots = 4;        // acceptors remembered per target in the parallel resolution; more -> sequential scan

// evaluates the entries [s_col + ...) of up to 32 grid columns for one query; calls emit(j, dist, lane_has) in
// the grid's iteration order, 32 candidates at a time.  All lanes of the warp must call it.
template <typename F>
__device__ __forceinline__ void for_each_candidate(const MatchArgs& a, const PairView& v, const int* cs, const int* items,
                                                   int i1, F&& emit) {
    const int lane = threadIdx.x & 31;
    const float2 q = ud_of(v.k1, v.ud1, i1);
    int cx0, cx1, cy0, cy1;
    if (!cell_range(a, q.x, q.y, cx0, cx1, cy0, cy1)) return;
    const uint4* dq = v.d1 + 2 * i1;
    const uint4 qa = dq[0], qb = dq[1];
    for (int cbase = cx0; cbase <= cx1; cbase += 32) {
        const int ncol = min(32, cx1 - cbase + 1);
        int s = 0, cnt = 0;
        if (lane < ncol) {
            s = cs[(cbase + lane) * a.grid.rows + cy0];
            cnt = cs[(cbase + lane) * a.grid.rows + cy1 + 1] - s;
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int T = __shfl_sync(0xffffffffu, incl, 31);
        for (int t0 = 0; t0 < T; t0 += 32) {
            const int t = t0 + lane;
            // column of entry t = number of columns whose inclusive count is <= t (binary search over the lanes)
            int col = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int probe = __shfl_sync(0xffffffffu, incl, col + step - 1);
                if (probe <= t) col += step;
            }
            col = min(col, 31);
            const int sc = __shfl_sync(0xffffffffu, s, col), ic = __shfl_sync(0xffffffffu, incl, col);
            const int cc = __shfl_sync(0xffffffffu, cnt, col);
            int j = -1, d = 0;
            if (t < T) {
                const int jj = items[sc + (t - (ic - cc))];
                const float2 pt = ud_of(v.k2, v.ud2, jj);
                if (fabsf(__fsub_rn(pt.x, q.x)) < a.window && fabsf(__fsub_rn(pt.y, q.y)) < a.window) {
                    const uint4 ta = __ldg(v.d2 + 2 * jj), tb = __ldg(v.d2 + 2 * jj + 1);
                    d = __popc(qa.x ^ ta.x) + __popc(qa.y ^ ta.y) + __popc(qa.z ^ ta.z) + __popc(qa.w ^ ta.w) +
                        __popc(qb.x ^ tb.x) + __popc(qb.y ^ tb.y) + __popc(qb.z ^ tb.z) + __popc(qb.w ^ tb.w);
                    j = jj;
                }
            }
            emit(j, d);
        }
    }
}

// rotation-histogram bin of an accepted pair (:175-185); kNoBin when it falls outside [0, HISTO_LENGTH)
constexpr int kNoBin = 31;
__device__ __forceinline__ int rot_bin(float angle1, float angle2) {
    float rot = __fsub_rn(angle1, angle2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, 1.0f / kHisto));
    if (bin == kHisto) bin = 0;
    return (bin >= 0 && bin < kHisto) ? bin : kNoBin;
}

__global__ void __launch_bounds__(1024) match_window_kernel(const MatchArgs a, const MatchSmem sm) {
    extern __shared__ int s_dynm[];
    __shared__ int s_scan[33];
    __shared__ int s_hist[kHisto];
    __shared__ int s_ind[3];
    __shared__ int s_nm, s_nA, s_poolUsed, s_big, s_changed, s_ovf, s_nq;
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nth >> 5;
    const int pr = a.pairOrder ? a.pairOrder[a.pairBase + blockIdx.x] : a.pairBase + (int)blockIdx.x;
    PairView v;
    {
        long long o1, o2;
        if (a.pairs) { o1 = (long long)a.pairs[2 * pr] * a.stride1; o2 = (long long)a.pairs[2 * pr + 1] * a.stride2; }
        else { o1 = (long long)pr * a.stride1; o2 = (long long)pr * a.stride2; }
        v.k1 = a.k1 + o1; v.k2 = a.k2 + o2;
        v.ud1 = a.ud1 ? a.ud1 + 2 * o1 : nullptr; v.ud2 = a.ud2 ? a.ud2 + 2 * o2 : nullptr;
        v.d1 = reinterpret_cast<const uint4*>(a.d1 + 32 * o1); v.d2 = reinterpret_cast<const uint4*>(a.d2 + 32 * o2);
        v.n1 = a.pairs ? a.n1[a.pairs[2 * pr]] : a.n1[pr];
        v.n2 = a.pairs ? a.n2[a.pairs[2 * pr + 1]] : a.n2[pr];
    }
    const int nCells = a.grid.cols * a.grid.rows;
    int* cellOf = a.cellOf + (long long)pr * a.cap;           // later reused as the active-query list
    int* cellStart = sm.cells ? s_dynm : a.cellStart + (long long)pr * (nCells + 1);
    int* cellFill = a.cellFill + (long long)pr * nCells;
    int* items = a.cellItems + (long long)pr * a.cap;
    int* cand = a.cand + (long long)pr * a.cap * kCandCap;
    int* candCnt = a.candCnt + (long long)pr * a.cap;
    const bool n2InSmem = sm.n2 >= v.n2 && sm.n2 > 0;
    int* dist2 = n2InSmem ? s_dynm + sm.cells : a.dist2 + (long long)pr * a.cap;
    int* m21 = n2InSmem ? s_dynm + sm.cells + sm.n2 : a.m21 + (long long)pr * a.cap;
    int* s_act = s_dynm + sm.cells + 2 * sm.n2;                // [4][sm.act]: i1, pool offset, count, decision
    int* s_pool = s_act + 4 * sm.act;
    int* accCnt = s_pool + sm.pool;                            // [sm.acc] acceptors per target, [sm.acc][kAccSlots] entries
    int* accList = accCnt + sm.acc;
    int* bins = a.bins + (long long)pr * a.cap;
    int* m12 = a.matches12 + (long long)pr * a.cap;

    // ---- FeatureGrid::assignFeaturesToGrid (:115-152) as CSR, cell id = ix*rows + iy ------------------
    // Only octave-0 keypoints of frame 2 enter the grid: every query has octave 0 (:120-122) and asks for
    // getFeaturesInArea(..., minLevel = maxLevel = its own octave) (:124), so no other keypoint can be a candidate;
    // the order inside a cell (insertion = index order) is unchanged for the ones that remain.
    for (int c = tid; c < nCells; c += nth) cellFill[c] = 0;
    for (int j = tid; j < v.n2; j += nth) { dist2[j] = 0x7fffffff; m21[j] = -1; }
    for (int i = tid; i < v.n1; i += nth) { m12[i] = -1; bins[i] = -1; candCnt[i] = 0; }
    if (tid < kHisto) s_hist[tid] = 0;
    if (tid == 0) { s_nA = 0; s_poolUsed = 0; s_big = 0; s_changed = 0; s_ovf = 0; s_nq = 0; }
    __syncthreads();
    // the queries — octave-0 keypoints of frame 1 (:120-122) — are listed up front (any order: their candidate lists are
    // independent), in the candidate pool, which is free until phase A is over; otherwise every warp of phase A would
    // discover its queries one dependent global load at a time (a quarter of this kernel's stall samples)
    for (int i = tid; i < v.n1; i += nth)
        if (v.k1[i].octave == 0) {
            const int q = atomicAdd(&s_nq, 1);
            if (q < sm.pool) s_pool[q] = i;
        }
    for (int j = tid; j < v.n2; j += nth) {
        int c = -1;
        if (v.k2[j].octave == 0) {
            const float2 pt = ud_of(v.k2, v.ud2, j);
            const int px = (int)roundf(__fmul_rn(__fsub_rn(pt.x, a.grid.min_x), a.invW));
            const int py = (int)roundf(__fmul_rn(__fsub_rn(pt.y, a.grid.min_y), a.invH));
            if (!(px < 0 || px >= a.grid.cols || py < 0 || py >= a.grid.rows)) { c = px * a.grid.rows + py; atomicAdd(&cellFill[c], 1); }
        }
        cellOf[j] = c;
    }
    __syncthreads();
    int run = 0;
    for (int base = 0; base < nCells; base += nth) {
        const int c = base + tid;
        const int cnt = c < nCells ? cellFill[c] : 0;
        int tot;
        const int ex = block_excl_scan_m(cnt, &tot, s_scan);
        if (c < nCells) cellStart[c] = run + ex;
        run += tot;
    }
    if (tid == 0) cellStart[nCells] = run;
    __syncthreads();
    for (int c = tid; c < nCells; c += nth) cellFill[c] = 0;
    __syncthreads();
    for (int j = tid; j < v.n2; j += nth) {
        const int c = cellOf[j];
        if (c >= 0) items[cellStart[c] + atomicAdd(&cellFill[c], 1)] = j;
    }
    __syncthreads();
    for (int c = tid; c < nCells; c += nth) {           // insertion order inside a cell = index order
        const int s = cellStart[c], e = cellStart[c + 1];
        for (int i = s + 1; i < e; ++i) {
            const int val = items[i];
            int k = i - 1;
            while (k >= s && items[k] > val) { items[k + 1] = items[k]; --k; }
            items[k + 1] = val;
        }
    }
    __syncthreads();

    // ---- phase A: per query (one warp each), pruned candidate list in the grid's iteration order -----------
    // A candidate whose distance can neither be an accepted best (dist > TH_LOW) nor fail the ratio
    // test as second best ((float)dist*ratio > TH_LOW >= best) is equivalent to "no candidate".
    // Entry = (j << 14) | (rotation bin << 9) | distance, so that the sequential phase touches no global memory.
    const float thLowF = (float)a.thLow;
    const int nq = s_nq;
    const bool listed = nq <= sm.pool;                           // (else: walk all keypoints, as the reference does)
    for (int qi = wid; qi < (listed ? nq : v.n1); qi += nw) {
        const int i1 = listed ? s_pool[qi] : qi;
        if (!listed && v.k1[i1].octave > 0) continue;            // :120-122
        const float ang1 = v.k1[i1].angle;
        int cnt = 0;
        for_each_candidate(a, v, cellStart, items, i1, [&](int j, int d) {
            const bool keep = j >= 0 && !(d > a.thLow && __fmul_rn((float)d, a.nnratio) > thLowF);
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
                if (pos < kCandCap) cand[(long long)i1 * kCandCap + pos] = (j << 14) | (rot_bin(ang1, v.k2[j].angle) << 9) | d;
            }
            cnt += __popc(bal);
        });
        if (lane == 0) { candCnt[i1] = cnt; if (cnt > kCandCap) s_big = 1; }
    }
    __syncthreads();

    // ---- ordered compaction of the active queries (+ their candidates into the shared pool) -----------------
    int* actList = cellOf;      // global list of active i1 (always), shared table when it fits
    {
        int nA = 0, poolUsed = 0;
        for (int base = 0; base < v.n1; base += nth) {
            const int i1 = base + tid;
            const int cnt = i1 < v.n1 ? candCnt[i1] : 0;
            const int stored = (cnt > 0 && cnt <= kCandCap) ? cnt : 0;
            int totA, totP;
            const int exA = block_excl_scan_m(cnt > 0, &totA, s_scan);
            const int exP = block_excl_scan_m(stored, &totP, s_scan);
            if (cnt > 0) {
                const int k = nA + exA;
                actList[k] = i1;
                if (k < sm.act) {
                    const bool fits = poolUsed + exP + stored <= sm.pool;
                    s_act[k] = i1; s_act[sm.act + k] = fits ? poolUsed + exP : -1; s_act[2 * sm.act + k] = cnt;
                    if (fits)
                        for (int c = 0; c < stored; ++c) s_pool[poolUsed + exP + c] = cand[(long long)i1 * kCandCap + c];
                }
            }
            nA += totA; poolUsed += totP;
        }
        if (tid == 0) s_nA = nA;
    }
    __syncthreads();

    // ---- phase B: the accept / steal scan over i1 (:114-187) --------------------------------------------
    // Sequentially, query k sees dist2[j] = min{ best(k') : k' < k accepted target j } (a later acceptor always has a
    // strictly smaller distance, :146) and decides from that alone; so the decisions D_k solve a triangular system
    // D_k = f(D_0..D_{k-1}).  Parallel form: iterate D <- F(D) from "nobody accepted" with all queries in parallel
    // (one warp each); after t rounds the first t decisions are final, and a round that changes nothing is the
    // unique solution, i.e. exactly the sequential result.  In practice a handful of rounds.  Per target the
    // acceptors of the previous round are kept in kAccSlots slots; an overflow, a query with more than kCandCap
    // candidates or tables that do not fit in shared memory fall back to the literal sequential scan below.
    const int nA = s_nA;
    bool parallelDone = false;
    if (sm.acc >= v.n2 && sm.acc > 0 && nA <= sm.act && !s_big) {
        int* dec = s_act + 3 * sm.act;
        for (int k = tid; k < nA; k += nth) dec[k] = -1;
        __syncthreads();
        for (int round = 0; round <= nA; ++round) {
            for (int j = tid; j < v.n2; j += nth) accCnt[j] = 0;
            __syncthreads();
            if (tid == 0) s_changed = 0;      // (everybody read the previous round's flag before the barrier above)
            for (int k = tid; k < nA; k += nth) {
                const int d = dec[k];
                if (d >= 0) {
                    const int j = d >> 9;
                    const int slot = atomicAdd(&accCnt[j], 1);
                    if (slot < kAccSlots) accList[j * kAccSlots + slot] = (k << 9) | (d & 511);
                    else s_ovf = 1;
                }
            }
            __syncthreads();
            if (s_ovf) break;
            for (int k = wid; k < nA; k += nw) {
                const int cnt = s_act[2 * sm.act + k], off = s_act[sm.act + k];
                int pk = 0, d = 0x7fffffff;
                if (lane < cnt) {
                    pk = off >= 0 ? s_pool[off + lane] : cand[(long long)s_act[k] * kCandCap + lane];
                    const int j = pk >> 14, dd = pk & 511;
                    const int c = min(accCnt[j], kAccSlots);
                    int before = 0x7fffffff;                              // dist2[j] as query k would see it
                    for (int t = 0; t < c; ++t) {
                        const int e = accList[j * kAccSlots + t];
                        if ((e >> 9) < k) before = min(before, e & 511);
                    }
                    if (!(before <= dd)) d = dd;                          // :146
                }
                int nd = -1;
                const int key = (d == 0x7fffffff) ? 0x7fffffff : ((d << 5) | lane);
                const int kmin = __reduce_min_sync(0xffffffffu, key);
                if (kmin != 0x7fffffff) {
                    const int bl = kmin & 31, best = kmin >> 5;
                    const int bpk = __shfl_sync(0xffffffffu, pk, bl);
                    const int best2 = __reduce_min_sync(0xffffffffu, (lane == bl) ? 0x7fffffff : d);
                    if (best <= a.thLow && (float)best < __fmul_rn((float)best2, a.nnratio)) nd = ((bpk >> 14) << 9) | best;   // :161-163
                }
                if (lane == 0 && nd != dec[k]) { dec[k] = nd; s_changed = 1; }
            }
            __syncthreads();
            if (!s_changed) { parallelDone = true; break; }
        }
        if (parallelDone) {
            // the acceptor lists now describe the final decisions: the owner of a target is its LAST acceptor (:165-169);
            // every acceptance, stolen later or not, left its rotation bin in the histogram (:175-185)
            for (int k = tid; k < nA; k += nth) {
                const int d = dec[k];
                if (d < 0) continue;
                const int j = d >> 9, i1 = s_act[k];
                const int c = min(accCnt[j], kAccSlots);
                bool owner = true;
                for (int t = 0; t < c; ++t) owner &= (accList[j * kAccSlots + t] >> 9) <= k;
                if (owner) m12[i1] = j;
                if (a.checkOri) {
                    const int bin = rot_bin(v.k1[i1].angle, v.k2[j].angle);
                    if (bin != kNoBin) { bins[i1] = bin; atomicAdd(&s_hist[bin], 1); }
                }
            }
        } else {
            for (int j = tid; j < v.n2; j += nth) { dist2[j] = 0x7fffffff; m21[j] = -1; }      // the tables share memory
        }
        __syncthreads();
    }
    if (wid == 0) {
        const int nSeq = parallelDone ? 0 : nA;
        auto fetch = [&](int k, int& i1, int& cnt, int& off, int& pk) {
            i1 = -1; cnt = 0; off = -1; pk = 0;
            if (k >= nSeq) return;
            if (k < sm.act) { i1 = s_act[k]; off = s_act[sm.act + k]; cnt = s_act[2 * sm.act + k]; }
            else { i1 = actList[k]; cnt = candCnt[i1]; }
            if (cnt <= kCandCap && lane < cnt) pk = off >= 0 ? s_pool[off + lane] : cand[(long long)i1 * kCandCap + lane];
        };
        int i1, cnt, off, pk;
        fetch(0, i1, cnt, off, pk);
        for (int k = 0; k < nSeq; ++k) {
            const int ci1 = i1, ccnt = cnt, cpk = pk;
            fetch(k + 1, i1, cnt, off, pk);
            int best = 0x7fffffff, best2 = 0x7fffffff, bestIdx = -1, bestBin = kNoBin;
            if (ccnt <= kCandCap) {
                int d = 0x7fffffff;
                const int j = cpk >> 14;
                if (lane < ccnt) {
                    const int dd = cpk & 511;
                    if (!(dist2[j] <= dd)) d = dd;               // :146
                }
                // first minimum in list order, then the second smallest of the multiset
                const int key = (d == 0x7fffffff) ? 0x7fffffff : ((d << 5) | lane);
                const int kmin = __reduce_min_sync(0xffffffffu, key);
                if (kmin != 0x7fffffff) {
                    const int bl = kmin & 31;
                    best = kmin >> 5;
                    const int bpk = __shfl_sync(0xffffffffu, cpk, bl);
                    bestIdx = bpk >> 14; bestBin = (bpk >> 9) & 31;
                    best2 = __reduce_min_sync(0xffffffffu, (lane == bl) ? 0x7fffffff : d);
                }
            } else {
                // exact slow path: re-enumerate every candidate of this query in order
                for_each_candidate(a, v, cellStart, items, ci1, [&](int j, int dd) {
                    int d = 0x7fffffff;
                    if (j >= 0 && !(dist2[j] <= dd)) d = dd;
                    const int key = (d == 0x7fffffff) ? 0x7fffffff : ((d << 5) | lane);
                    const int kmin = __reduce_min_sync(0xffffffffu, key);
                    if (kmin == 0x7fffffff) return;
                    const int bl = kmin & 31, cb = kmin >> 5;
                    const int cj = __shfl_sync(0xffffffffu, j, bl);
                    const int d2 = __reduce_min_sync(0xffffffffu, (lane == bl) ? 0x7fffffff : d);
                    if (cb < best) { best2 = min(best, d2); best = cb; bestIdx = cj; }
                    else best2 = min(best2, cb);
                });
                if (bestIdx >= 0) bestBin = rot_bin(v.k1[ci1].angle, v.k2[bestIdx].angle);
            }
            if (best <= a.thLow && (float)best < __fmul_rn((float)best2, a.nnratio)) {      // :161-163
                if (lane == 0) {
                    const int prev = m21[bestIdx];
                    if (prev >= 0) m12[prev] = -1;
                    m12[ci1] = bestIdx;
                    m21[bestIdx] = ci1;
                    dist2[bestIdx] = best;
                    if (a.checkOri && bestBin != kNoBin) { bins[ci1] = bestBin; s_hist[bestBin] += 1; }
                }
                __syncwarp();
            }
        }
        // ComputeThreeMaxima (:28-69)
        if (lane == 0) {
            int i1m = -1, i2 = -1, i3 = -1;
            if (a.checkOri) {
                int m1 = 0, m2 = 0, m3 = 0;
                for (int i = 0; i < kHisto; ++i) {
                    const int s = s_hist[i];
                    if (s > m1) { m3 = m2; m2 = m1; m1 = s; i3 = i2; i2 = i1m; i1m = i; }
                    else if (s > m2) { m3 = m2; m2 = s; i3 = i2; i2 = i; }
                    else if (s > m3) { m3 = s; i3 = i; }
                }
                if ((float)m2 < __fmul_rn(0.1f, (float)m1)) { i2 = -1; i3 = -1; }
                else if ((float)m3 < __fmul_rn(0.1f, (float)m1)) { i3 = -1; }
            }
            s_ind[0] = i1m; s_ind[1] = i2; s_ind[2] = i3;
            s_nm = 0;
        }
    }
    __syncthreads();
    int nm = 0;
    for (int i = tid; i < v.n1; i += nth) {
        int m = m12[i];
        if (a.checkOri && m >= 0) {
            const int b = bins[i];
            if (b >= 0 && b != s_ind[0] && b != s_ind[1] && b != s_ind[2]) { m = -1; m12[i] = -1; }
        }
        nm += m >= 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nm += __shfl_xor_sync(0xffffffffu, nm, o);
    if (lane == 0 && nm) atomicAdd(&s_nm, nm);
    __syncthreads();
    if (tid == 0) a.nMatches[pr] = s_nm;
}

// ------------------------------------------------------------------------------------------
// brute-force kNN-2: one thread per query, train descriptors staged through shared memory in
// tiles (every thread reads the same address -> broadcast).  Strict '<' keeps the lowest train
// index on ties, like cv::BFMatcher.
// ------------------------------------------------------------------------------------------
// The train set is cut into gridDim.y segments so that a 2000 x 2000 problem fills the GPU (16 query blocks alone would
// use 16 of 148 SMs); every (query, segment) writes its best two as (d0, j0, d1, j1) and bf_merge_kernel folds the
// segments.  "Best two" of the sequential scan = the two smallest (distance, index) pairs in lexicographic order (strict
// '<' keeps the earlier index on ties), so merging per-segment results in that order is exact.
__global__ void __launch_bounds__(128) bf_knn2_kernel(const uint4* __restrict__ d1, int n1, const uint4* __restrict__ d2,
                                                      int n2, int norm, int segLen, int4* __restrict__ part) {
    __shared__ uint4 s_t[128 * 2];
    const int q = blockIdx.x * 128 + threadIdx.x;
    uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
    if (q < n1) { qa = d1[2 * q]; qb = d1[2 * q + 1]; }
    int b0 = 0x7fffffff, b1 = 0x7fffffff, j0 = -1, j1 = -1;
    const int segBeg = blockIdx.y * segLen, segEnd = min(segBeg + segLen, n2);
    for (int base = segBeg; base < segEnd; base += 128) {
        const int nt = min(128, segEnd - base);
        __syncthreads();
        for (int k = threadIdx.x; k < nt * 2; k += 128) s_t[k] = d2[2 * base + k];
        __syncthreads();
        for (int t = 0; t < nt; ++t) {
            const uint4 ta = s_t[2 * t], tb = s_t[2 * t + 1];
            int d;
            if (norm == NAV24_NORM_HAMMING) {
                d = __popc(qa.x ^ ta.x) + __popc(qa.y ^ ta.y) + __popc(qa.z ^ ta.z) + __popc(qa.w ^ ta.w) +
                    __popc(qb.x ^ tb.x) + __popc(qb.y ^ tb.y) + __popc(qb.z ^ tb.z) + __popc(qb.w ^ tb.w);
            } else {
                unsigned acc = 0, df;
                df = __vabsdiffu4(qa.x, ta.x); acc = __dp4a(df, df, acc);
                df = __vabsdiffu4(qa.y, ta.y); acc = __dp4a(df, df, acc);
                df = __vabsdiffu4(qa.z, ta.z); acc = __dp4a(df, df, acc);
                df = __vabsdiffu4(qa.w, ta.w); acc = __dp4a(df, df, acc);
                df = __vabsdiffu4(qb.x, tb.x); acc = __dp4a(df, df, acc);
                df = __vabsdiffu4(qb.y, tb.y); acc = __dp4a(df, df, acc);
                df = __vabsdiffu4(qb.z, tb.z); acc = __dp4a(df, df, acc);
                df = __vabsdiffu4(qb.w, tb.w); acc = __dp4a(df, df, acc);
                d = (int)acc;
            }
            const int j = base + t;
            if (d < b0) { b1 = b0; j1 = j0; b0 = d; j0 = j; }
            else if (d < b1) { b1 = d; j1 = j; }
        }
    }
    if (q < n1) part[(size_t)blockIdx.y * n1 + q] = make_int4(b0, j0, b1, j1);
}

// folds the segments of one query in ascending train order (so equal distances keep the earlier index), then applies
// the ratio test of FtAssocOCV::match (OP_FtAssoc.cpp:73-85)
__global__ void __launch_bounds__(128) bf_merge_kernel(const int4* __restrict__ part, int n1, int nSeg, int norm, float ratio,
                                                       int* idx0, int* idx1, float* dist0, float* dist1, uint8_t* pass) {
    const int q = blockIdx.x * 128 + threadIdx.x;
    if (q >= n1) return;
    int b0 = 0x7fffffff, b1 = 0x7fffffff, j0 = -1, j1 = -1;
    for (int sgm = 0; sgm < nSeg; ++sgm) {
        const int4 c = part[(size_t)sgm * n1 + q];
        if (c.y >= 0) {
            if (c.x < b0) { b1 = b0; j1 = j0; b0 = c.x; j0 = c.y; }
            else if (c.x < b1) { b1 = c.x; j1 = c.y; }
        }
        if (c.w >= 0 && c.z < b1) { b1 = c.z; j1 = c.w; }
    }
    idx0[q] = j0; idx1[q] = j1;
    const float f0 = j0 < 0 ? 0.f : (norm == NAV24_NORM_HAMMING ? (float)b0 : __fsqrt_rn((float)b0));
    const float f1 = j1 < 0 ? 0.f : (norm == NAV24_NORM_HAMMING ? (float)b1 : __fsqrt_rn((float)b1));
    dist0[q] = f0; dist1[q] = f1;
    pass[q] = (j0 >= 0 && j1 >= 0 && f0 < __fmul_rn(ratio, f1)) ? 1 : 0;
}

}  // namespace

int launch_match_window(const MatchArgs& a, int P, cudaStream_t s) {
    // shared-memory budget: CSR cell starts, dist2/m21, active-query table, candidate pool (each optional)
    const int budget = 200 * 1024 / 4;      // ints
    MatchSmem sm{};
    const int nCells = a.grid.cols * a.grid.rows;
    int used = 0;
    if (nCells + 1 <= 40000) { sm.cells = (nCells + 1 + 3) & ~3; used += sm.cells; }
    if (2 * a.cap <= budget - used - 4096) { sm.n2 = (a.cap + 3) & ~3; used += 2 * sm.n2; }
    sm.act = min(max((budget - used) / 10, 0), (a.cap + 3) & ~3);
    used += 4 * sm.act;
    sm.pool = min(max(budget - used, 0), 8192);
    used += sm.pool;
    const int capR = (a.cap + 3) & ~3;
    if ((1 + kAccSlots) * capR <= budget - used) { sm.acc = capR; used += (1 + kAccSlots) * capR; }
    sm.total = used * 4;
    // (function attributes are per device and this library serves several devices and host threads: set on every launch)
    cudaFuncSetAttribute(match_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max(sm.total, 48 * 1024));
    match_window_kernel<<<P, 1024, sm.total, s>>>(a, sm);
    return 1;
}

int bf_knn2_segments(int n1, int n2) {      // train segments such that query blocks x segments cover the 148 SMs about twice
    const int qb = (n1 + 127) / 128;
    const int want = (2 * 148 + qb - 1) / qb;
    return std::max(1, std::min(want, (n2 + 127) / 128));
}

int launch_bf_knn2(const uint8_t* d1, int n1, const uint8_t* d2, int n2, int norm, float ratio, int* idx0, int* idx1,
                   float* dist0, float* dist1, uint8_t* pass, int4* part, cudaStream_t s) {
    const int nSeg = bf_knn2_segments(n1, n2);
    const int segLen = std::max(128, ((n2 + nSeg - 1) / nSeg + 127) / 128 * 128);
    dim3 grid((n1 + 127) / 128, nSeg);
    bf_knn2_kernel<<<grid, 128, 0, s>>>(reinterpret_cast<const uint4*>(d1), n1, reinterpret_cast<const uint4*>(d2), n2, norm,
                                        segLen, part);
    bf_merge_kernel<<<(n1 + 127) / 128, 128, 0, s>>>(part, n1, nSeg, norm, ratio, idx0, idx1, dist0, dist1, pass);
    return 2;
}

}  // namespace nav24
