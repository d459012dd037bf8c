// dense.cuh — K4 placeholder (implemented below in a later commit).
#pragma once
#include "common.cuh"
namespace pioran {
typedef int (*fail_fn)(int, const char*, ...);
inline int dense_logl_host(cudaStream_t, int64_t*, int64_t, const double*, const double*, const double*, int, int,
                           const double*, const double*, const double*, const double*, const double*, const double*,
                           double*, int*, fail_fn fail) { return fail(-5, "dense path not built yet"); }
}
