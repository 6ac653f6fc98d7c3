// stdsort.cuh — exact restatement of libstdc++'s std::sort control flow for 64-bit records whose
// ordering key is the HIGH 32 bits (records with equal keys are "equivalent").
//
// Why: the reference's quadtree (core/operators/objDetection/OP_FtDtOrbSlam.cpp:646) calls
// std::sort with a comparator that only orders by (count, UL.x); which of several equivalent
// nodes ends up last decides which node is split next, so the retained keypoint set depends on
// the exact permutation libstdc++'s unstable introsort produces.
// This file re-states that procedure (introsort loop with median-of-3 + Hoare partition, depth
// limit 2*floor(log2 n) with heap-sort fallback, threshold 16, final insertion sort) so that a
// single GPU thread — or the host, for the CPU unit test against the real std::sort — produces
// the same permutation.  Written from the algorithm's published structure
// (bits/stl_algo.h:1848-1952, bits/stl_heap.h), iteratively, with an explicit range stack.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define NAV24_HD __host__ __device__ __forceinline__
#else
#define NAV24_HD inline
#endif

namespace nav24 {
namespace stdsort {

typedef unsigned long long rec_t;

NAV24_HD bool less(rec_t a, rec_t b) { return (uint32_t)(a >> 32) < (uint32_t)(b >> 32); }

NAV24_HD void swap_rec(rec_t* v, int i, int j) {
    rec_t t = v[i];
    v[i] = v[j];
    v[j] = t;
}

// heap helpers on the sub-array v[0..len)
NAV24_HD void sift_hole(rec_t* v, int hole, int len, rec_t value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less(v[child], v[child - 1])) child--;
        v[hole] = v[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        v[hole] = v[child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && less(v[parent], value)) {
        v[hole] = v[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    v[hole] = value;
}

// partial_sort(first, last, last): make_heap + sort_heap (the heap_select loop is empty)
NAV24_HD void heap_sort(rec_t* v, int len) {
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            rec_t value = v[parent];
            sift_hole(v, parent, len, value);
            if (parent == 0) break;
            parent--;
        }
    }
    int last = len;
    while (last > 1) {
        --last;
        rec_t value = v[last];
        v[last] = v[0];
        sift_hole(v, 0, last, value);
    }
}

NAV24_HD void median_to_first(rec_t* v, int result, int a, int b, int c) {
    if (less(v[a], v[b])) {
        if (less(v[b], v[c])) swap_rec(v, result, b);
        else if (less(v[a], v[c])) swap_rec(v, result, c);
        else swap_rec(v, result, a);
    } else if (less(v[a], v[c])) swap_rec(v, result, a);
    else if (less(v[b], v[c])) swap_rec(v, result, c);
    else swap_rec(v, result, b);
}

NAV24_HD int partition_pivot(rec_t* v, int first, int last) {
    const int mid = first + (last - first) / 2;
    median_to_first(v, first, first + 1, mid, last - 1);
    int lo = first + 1, hi = last;
    const rec_t pivot_unused = 0;
    (void)pivot_unused;
    while (true) {
        while (less(v[lo], v[first])) ++lo;
        --hi;
        while (less(v[first], v[hi])) --hi;
        if (!(lo < hi)) return lo;
        swap_rec(v, lo, hi);
        ++lo;
    }
}

NAV24_HD int floor_log2(int n) {
    int l = 0;
    while (n > 1) { n >>= 1; ++l; }
    return l;
}

// the introsort loop over [0,n): leaves every run of <=16 records unsorted but in place
NAV24_HD void introsort_loop(rec_t* v, int n) {
    int stk_first[64], stk_last[64], stk_depth[64];
    int sp = 0;
    stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = 2 * floor_log2(n);
    sp = 1;
    while (sp > 0) {
        --sp;
        int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
        while (last - first > 16) {
            if (depth == 0) {
                heap_sort(v + first, last - first);
                break;
            }
            --depth;
            const int cut = partition_pivot(v, first, last);
            // "recurse" on [cut,last) with the decremented depth, continue on [first,cut)
            stk_first[sp] = cut; stk_last[sp] = last; stk_depth[sp] = depth;
            ++sp;
            last = cut;
        }
    }
}

NAV24_HD void linear_insert_unguarded(rec_t* v, int pos) {
    const rec_t val = v[pos];
    int next = pos - 1;
    while (less(val, v[next])) {
        v[pos] = v[next];
        pos = next;
        --next;
    }
    v[pos] = val;
}

NAV24_HD void insertion_sort_guarded(rec_t* v, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (less(v[i], v[first])) {
            const rec_t val = v[i];
            for (int k = i; k > first; --k) v[k] = v[k - 1];
            v[first] = val;
        } else {
            linear_insert_unguarded(v, i);
        }
    }
}

NAV24_HD void final_insertion_sort(rec_t* v, int n) {
    if (n > 16) {
        insertion_sort_guarded(v, 0, 16);
        for (int i = 16; i != n; ++i) linear_insert_unguarded(v, i);
    } else {
        insertion_sort_guarded(v, 0, n);
    }
}

// std::sort(v, v+n, key-less)
NAV24_HD void sort(rec_t* v, int n) {
    if (n <= 0) return;
    introsort_loop(v, n);
    final_insertion_sort(v, n);
}

}  // namespace stdsort
}  // namespace nav24
