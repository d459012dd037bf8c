// scan.cuh — K3 placeholder (implemented below in a later commit).
#pragma once
#include "common.cuh"
namespace pioran {
typedef int (*fail_fn2)(int, const char*, ...);
inline int scan_logl_host(cudaStream_t, int, int64_t*, int64_t, const double*, const double*, const double*, int, int,
                          const double*, const double*, const double*, const double*, const double*, const double*,
                          double*, fail_fn2 fail) { return fail(-5, "scan path not built yet"); }
}
