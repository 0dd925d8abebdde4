// cuemu shim of the few thrust entry points the library uses (TEST INFRASTRUCTURE).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <functional>
#include <numeric>
#include <vector>
namespace thrust {
struct cuemu_policy {
    cuemu_policy on(cudaStream_t) const { return *this; }
};
namespace cuda {
template <class A> inline cuemu_policy par_nosync(A &) { return cuemu_policy(); }
template <class A> inline cuemu_policy par(A &) { return cuemu_policy(); }
}   // namespace cuda
template <class T> using greater = std::greater<T>;
template <class T> using less = std::less<T>;
template <class P, class It> inline void sequence(const P &, It first, It last) { std::iota(first, last, 0); }
template <class P, class It, class Out> inline Out inclusive_scan(const P &, It first, It last, Out out)
{
    return std::partial_sum(first, last, out);
}
template <class P, class K, class V, class Cmp>
inline void stable_sort_by_key(const P &, K kfirst, K klast, V vfirst, Cmp cmp)
{
    const size_t n = (size_t)(klast - kfirst);
    std::vector<size_t> idx(n);
    std::iota(idx.begin(), idx.end(), (size_t)0);
    std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return cmp(kfirst[a], kfirst[b]); });
    std::vector<typename std::remove_reference<decltype(*kfirst)>::type> k(n);
    std::vector<typename std::remove_reference<decltype(*vfirst)>::type> v(n);
    for (size_t i = 0; i < n; ++i) { k[i] = kfirst[idx[i]]; v[i] = vfirst[idx[i]]; }
    for (size_t i = 0; i < n; ++i) { kfirst[i] = k[i]; vfirst[i] = v[i]; }
}
}   // namespace thrust
