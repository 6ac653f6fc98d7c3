// stdsort_warp.cuh — the SAME permutation as stdsort::sort (i.e. as libstdc++'s std::sort), produced by one warp.
//
// The single-thread restatement in stdsort.cuh made the quadtree kernel wait ~0.19 ms per launch for thread 0
// (ncu: 52 % of the warp samples of quadtree_kernel sat at the barrier behind it).  Introsort is sequential in
// its control flow but not in its data flow:
//   * __unguarded_partition(first+1, last, pivot): the k-th stop of the left scan is the k-th ORIGINAL position (from
//     the left) whose key is >= pivot, the k-th stop of the right scan the k-th original position (from the right)
//     whose key is <= pivot, for as long as left < right; exactly those pairs are swapped.  When the scans cross at
//     step K the left scan stops at min(a_K, b_{K-1}) (b_{K-1} now holds a swapped-in element >= pivot).  Both
//     position lists come from ballot/popc prefix ranks, K from one more ballot, the swaps are disjoint.
//   * the two sub-ranges of a partition never interact, so a range stack processed by the whole warp in any order
//     leaves the array exactly as the recursive original does;
//   * __final_insertion_sort never moves a record across a partition boundary (everything left of a boundary is
//     <= everything right of it and the insert stops at "not less"), so every leaf range (<= 16 records, or a
//     heap-sorted range) is insertion-sorted by its own lane.
// tests/test_gpu_parity.py::test_warp_sort_equals_std_sort pins it against the real std::sort (through the oracle).
#pragma once
#include "stdsort.cuh"

namespace nav24 {
namespace stdsort {

__device__ __forceinline__ unsigned key_of(rec_t r) { return (unsigned)(r >> 32); }

// v[first..last): median-of-3 to first, Hoare partition of [first+1, last) around v[first]; returns the cut.
// sa/sb: scratch for >= last-first-1 positions each.  All 32 lanes must call.
__device__ __forceinline__ int warp_partition(rec_t* v, int first, int last, unsigned short* sa, unsigned short* sb) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    if (lane == 0) median_to_first(v, first, first + 1, first + (last - first) / 2, last - 1);
    __syncwarp();
    const unsigned pk = key_of(v[first]);
    const int lo0 = first + 1, m = last - lo0;
    int nA = 0, nB = 0;
    for (int base = 0; base < m; base += 32) {
        const int i = base + lane;
        const bool in = i < m && key_of(v[lo0 + i]) >= pk;
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (in) sa[nA + __popc(bal & lt)] = (unsigned short)i;
        nA += __popc(bal);
    }
    for (int base = 0; base < m; base += 32) {
        const int i = m - 1 - (base + lane);
        const bool in = i >= 0 && key_of(v[lo0 + i]) <= pk;
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (in) sb[nB + __popc(bal & lt)] = (unsigned short)i;
        nB += __popc(bal);
    }
    __syncwarp();
    const int nMin = min(nA, nB);
    int K = 0;
    for (int base = 0; base < nMin; base += 32) {
        const int k = base + lane;
        const bool ok = k < nMin && sa[k] < sb[k];
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        K += __popc(bal);
        if (bal != 0xffffffffu) break;
    }
    const int aK = K < nA ? (int)sa[K] : 0x7fffffff;
    const int bK1 = K > 0 ? (int)sb[K - 1] : 0x7fffffff;
    for (int k = lane; k < K; k += 32) {
        const int i = lo0 + sa[k], j = lo0 + sb[k];
        const rec_t t = v[i];
        v[i] = v[j];
        v[j] = t;
    }
    __syncwarp();
    return lo0 + min(aK, bK1);
}

// std::sort(v, v+n) by one warp.  Scratch: sa, sb >= n entries each; startBits >= (n+31)/32 words; stk >= 192 ints.
__device__ __forceinline__ void sort_warp(rec_t* v, int n, unsigned short* sa, unsigned short* sb, unsigned* startBits,
                                          int* stk) {
    const int lane = threadIdx.x & 31;
    if (n <= 1) return;
    const int nW = (n + 31) >> 5;
    for (int w = lane; w < nW; w += 32) startBits[w] = 0u;
    if (lane == 0) { stk[0] = 0; stk[1] = n; stk[2] = 2 * floor_log2(n); }
    __syncwarp();
    int sp = 1;
    while (sp > 0) {
        --sp;
        int first = stk[3 * sp], last = stk[3 * sp + 1], depth = stk[3 * sp + 2];
        __syncwarp();
        while (last - first > 16) {
            if (depth == 0) {
                if (lane == 0) heap_sort(v + first, last - first);
                __syncwarp();
                break;
            }
            --depth;
            const int cut = warp_partition(v, first, last, sa, sb);
            if (lane == 0) { stk[3 * sp] = cut; stk[3 * sp + 1] = last; stk[3 * sp + 2] = depth; }
            ++sp;
            last = cut;
            __syncwarp();
        }
        if (last > first && lane == 0) startBits[first >> 5] |= 1u << (first & 31);
        __syncwarp();
    }
    // every leaf range is insertion-sorted by one lane
    for (int w = lane; w < nW; w += 32) {
        const unsigned word = startBits[w];
        unsigned bits = word;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int s = w * 32 + b;
            int e;
            if (bits) e = w * 32 + __ffs(bits) - 1;
            else {
                int ww = w + 1;
                while (ww < nW && startBits[ww] == 0u) ++ww;
                e = ww < nW ? ww * 32 + __ffs(startBits[ww]) - 1 : n;
            }
            for (int i = s + 1; i < e; ++i) {
                const rec_t val = v[i];
                int j = i - 1;
                while (j >= s && less(val, v[j])) { v[j + 1] = v[j]; --j; }
                v[j + 1] = val;
            }
        }
    }
    __syncwarp();
}

// std::sort(v, v+n) by a whole CTA (blockDim.x a multiple of 32), same permutation again.  The range stack of sort_warp
// becomes a list of ranges per ROUND: every range longer than 16 is partitioned by one warp (warp_partition works inside
// the range's own slice of sa / sb), its two halves go to the next round's list or are marked as leaves; rounds are
// separated by block barriers.  With w warps the critical path is about n + n/2 + n/4 + ... elements instead of
// n * log2(n / 16).  Scratch: sa, sb >= n entries each; startBits >= (n+31)/32 words; ranges >= 2 * (3 * (n / 16 + 2) + 1)
// ints.  All threads of the block must call.
__device__ __forceinline__ int block_sort_range_ints(int n) { return 2 * (3 * (n / 16 + 2) + 1); }

__device__ __forceinline__ void sort_block(rec_t* v, int n, unsigned short* sa, unsigned short* sb, unsigned* startBits,
                                           int* ranges) {
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nth >> 5;
    if (n <= 1) return;                                       // (uniform)
    const int nW = (n + 31) >> 5;
    const int listInts = 3 * (n / 16 + 2) + 1;                // [count | (first, last, depth) ...]
    int* cur = ranges;
    int* nxt = ranges + listInts;
    for (int w = tid; w < nW; w += nth) startBits[w] = 0u;
    __syncthreads();
    if (tid == 0) {
        nxt[0] = 0;
        if (n > 16) { cur[0] = 1; cur[1] = 0; cur[2] = n; cur[3] = 2 * floor_log2(n); }
        else { cur[0] = 0; startBits[0] = 1u; }
    }
    __syncthreads();
    while (cur[0] > 0) {                                      // (uniform: read after a barrier)
        const int nCur = cur[0];
        for (int r = wid; r < nCur; r += nw) {
            const int first = cur[1 + 3 * r], last = cur[2 + 3 * r];
            int depth = cur[3 + 3 * r];
            if (depth == 0) {
                if (lane == 0) { heap_sort(v + first, last - first); atomicOr(&startBits[first >> 5], 1u << (first & 31)); }
                __syncwarp();
                continue;
            }
            --depth;
            const int cut = warp_partition(v, first, last, sa + first, sb + first);
            if (lane == 0) {
                const int f2[2] = {first, cut}, l2[2] = {cut, last};
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int len = l2[c] - f2[c];
                    if (len > 16) {
                        const int slot = atomicAdd(&nxt[0], 1);
                        nxt[1 + 3 * slot] = f2[c]; nxt[2 + 3 * slot] = l2[c]; nxt[3 + 3 * slot] = depth;
                    } else if (len > 0) {
                        atomicOr(&startBits[f2[c] >> 5], 1u << (f2[c] & 31));
                    }
                }
            }
            __syncwarp();
        }
        __syncthreads();
        { int* t = cur; cur = nxt; nxt = t; }
        if (tid == 0) nxt[0] = 0;
        __syncthreads();
    }
    // every leaf range is insertion-sorted by one thread
    for (int w = tid; w < nW; w += nth) {
        unsigned bits = startBits[w];
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int s = w * 32 + b;
            int e;
            if (bits) e = w * 32 + __ffs(bits) - 1;
            else {
                int ww = w + 1;
                while (ww < nW && startBits[ww] == 0u) ++ww;
                e = ww < nW ? ww * 32 + __ffs(startBits[ww]) - 1 : n;
            }
            for (int i = s + 1; i < e; ++i) {
                const rec_t val = v[i];
                int j = i - 1;
                while (j >= s && less(val, v[j])) { v[j + 1] = v[j]; --j; }
                v[j + 1] = val;
            }
        }
    }
    __syncthreads();
}

}  // namespace stdsort
}  // namespace nav24
