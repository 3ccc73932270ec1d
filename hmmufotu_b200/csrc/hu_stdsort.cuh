// libstdc++'s std::sort, restated so that its data movement can be replayed on the device.
//
// The reference orders seeds (src/HmmUFOtu_main.cpp:139), estimated placements (:164) and final placements
// (src/hmmufotu.cpp:725-731) with std::sort, which is unstable: WHICH of several equal keys ends up first is
// decided by the algorithm's data movement, and under SURVEY.md F6 (every final log-likelihood of a read is the
// same constant) that is what picks the reported placement.  The functions below follow bits/stl_algo.h and
// bits/stl_heap.h of GCC 13 (the toolchain of this image; the algorithm has not changed since GCC 4.x) step by
// step: __introsort_loop with its 2 lg n depth budget, __move_median_to_first, __unguarded_partition, the
// __partial_sort (heap sort) fallback when the budget runs out, and __final_insertion_sort.  `Less` is called
// exactly where the library calls its comparator, with the same operands in the same order, so keys that do not
// form a strict weak order (NaN) move as they do there.  Sequential code for one thread; used for the <= 64
// placements of a read, for ranges a warp finishes, and for the heap-sort fallback of the literal seed order.
// tests/test_abi.py compares it with the real std::sort through hu_debug_stdsort.
#pragma once
#include "hu_internal.cuh"

namespace hu {

#define HU_SS_THRESHOLD 16 // _S_threshold

template <typename T>
HU_HD void ssSwap(T& x, T& y) { const T t = x; x = y; y = t; }

// __move_median_to_first(result, a, b, c), bits/stl_algo.h
template <typename T, typename Less>
HU_HD void ssMedianToFirst(T* v, int result, int a, int b, int c, Less less) {
	if (less(v[a], v[b])) {
		if (less(v[b], v[c])) ssSwap(v[result], v[b]);
		else if (less(v[a], v[c])) ssSwap(v[result], v[c]);
		else ssSwap(v[result], v[a]);
	} else if (less(v[a], v[c])) ssSwap(v[result], v[a]);
	else if (less(v[b], v[c])) ssSwap(v[result], v[c]);
	else ssSwap(v[result], v[b]);
}

// __unguarded_partition(first, last, pivot)
template <typename T, typename Less>
HU_HD int ssUnguardedPartition(T* v, int first, int last, int pivot, Less less) {
	for (;;) {
		while (less(v[first], v[pivot])) ++first;
		--last;
		while (less(v[pivot], v[last])) --last;
		if (!(first < last)) return first;
		ssSwap(v[first], v[last]);
		++first;
	}
}

// __push_heap(first, holeIndex, topIndex, value), bits/stl_heap.h
template <typename T, typename Less>
HU_HD void ssPushHeap(T* v, int hole, int top, T value, Less less) {
	int parent = (hole - 1) / 2;
	while (hole > top && less(v[parent], value)) {
		v[hole] = v[parent];
		hole = parent;
		parent = (hole - 1) / 2;
	}
	v[hole] = value;
}
// __adjust_heap(first, holeIndex, len, value)
template <typename T, typename Less>
HU_HD void ssAdjustHeap(T* v, int hole, int len, T value, Less less) {
	const int top = hole;
	int second = hole;
	while (second < (len - 1) / 2) {
		second = 2 * (second + 1);
		if (less(v[second], v[second - 1])) second--;
		v[hole] = v[second];
		hole = second;
	}
	if ((len & 1) == 0 && second == (len - 2) / 2) {
		second = 2 * (second + 1);
		v[hole] = v[second - 1];
		hole = second - 1;
	}
	ssPushHeap(v, hole, top, value, less);
}
// __partial_sort(first, last, last): __heap_select with middle == last is __make_heap alone, then __sort_heap
template <typename T, typename Less>
HU_HD void ssHeapSort(T* v, int n, Less less) {
	if (n >= 2) {
		int parent = (n - 2) / 2;
		for (;;) {
			const T value = v[parent];
			ssAdjustHeap(v, parent, n, value, less);
			if (parent == 0) break;
			parent--;
		}
	}
	for (int last = n; last > 1;) { // __sort_heap: __pop_heap(first, last, last)
		--last;
		const T value = v[last];
		v[last] = v[0];
		ssAdjustHeap(v, 0, last, value, less);
	}
}

// __unguarded_linear_insert(last)
template <typename T, typename Less>
HU_HD void ssUnguardedLinearInsert(T* v, int last, Less less) {
	const T val = v[last];
	int next = last - 1;
	while (less(val, v[next])) {
		v[last] = v[next];
		last = next;
		--next;
	}
	v[last] = val;
}
// __insertion_sort(first, last)
template <typename T, typename Less>
HU_HD void ssInsertionSort(T* v, int first, int last, Less less) {
	if (first == last) return;
	for (int i = first + 1; i != last; ++i) {
		if (less(v[i], v[first])) {
			const T val = v[i];
			for (int k = i; k > first; --k) v[k] = v[k - 1]; // move_backward(first, i, i + 1)
			v[first] = val;
		} else
			ssUnguardedLinearInsert(v, i, less);
	}
}
// __final_insertion_sort(first, last)
template <typename T, typename Less>
HU_HD void ssFinalInsertionSort(T* v, int n, Less less) {
	if (n > HU_SS_THRESHOLD) {
		ssInsertionSort(v, 0, HU_SS_THRESHOLD, less);
		for (int i = HU_SS_THRESHOLD; i != n; ++i) ssUnguardedLinearInsert(v, i, less);
	} else
		ssInsertionSort(v, 0, n, less);
}

HU_HD int ssLg(int n) { // std::__lg
	int k = 0;
	for (; n > 1; n >>= 1) ++k;
	return k;
}

// __introsort_loop(first, last, depth): the recursion into [cut, last) becomes an explicit stack (the two parts of
// a partition are disjoint, so the order in which they are finished does not change the result)
template <typename T, typename Less>
HU_HD void ssIntrosortLoop(T* v, int first0, int last0, int depth0, Less less) {
	int st[64][3], top = 0;
	st[0][0] = first0; st[0][1] = last0; st[0][2] = depth0; top = 1;
	while (top > 0) {
		--top;
		int first = st[top][0], last = st[top][1], depth = st[top][2];
		while (last - first > HU_SS_THRESHOLD) {
			if (depth == 0) {
				ssHeapSort(v + first, last - first, less);
				break;
			}
			--depth;
			const int mid = first + (last - first) / 2;
			ssMedianToFirst(v, first, first + 1, mid, last - 1, less);
			const int cut = ssUnguardedPartition(v, first + 1, last, first, less);
			if (top < 64) { st[top][0] = cut; st[top][1] = last; st[top][2] = depth; ++top; }
			last = cut;
		}
	}
}

// std::sort(v, v + n, less)
template <typename T, typename Less>
HU_HD void stdSortReplay(T* v, int n, Less less) {
	if (n <= 0) return;
	ssIntrosortLoop(v, 0, n, ssLg(n) * 2, less);
	ssFinalInsertionSort(v, n, less);
}

struct SsKeyLess { // elements are indices into a key array
	const double* key;
	HU_HD bool operator()(int a, int b) const { return key[a] < key[b]; }
};

// std::sort(x.rbegin(), x.rend(), byKey) over the elements 0..n-1 of an array: out[k] = index of the element that
// ends up at position k (descending keys).  Sorting through reverse iterators is sorting the mirrored array.
HU_HD void stdSortDescendingPerm(const double* key, int n, int* out) {
	for (int i = 0; i < n; ++i) out[i] = n - 1 - i; // mirrored view: r[i] = x[n-1-i]
	SsKeyLess less;
	less.key = key;
	stdSortReplay(out, n, less);
	for (int i = 0; i < n / 2; ++i) ssSwap(out[i], out[n - 1 - i]); // back: x[k] = r[n-1-k]
}

} // namespace hu
