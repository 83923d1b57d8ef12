// tooncrafter_b200 — host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/tooncrafter_b200.h"

namespace tc_host {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int check_cuda(cudaError_t e, const char* what);
void count_launch(int n = 1);
int sm_count();

// Cached cuTensorMapEncodeTiled: fp16 tensor of `rank` dims (dim 0 contiguous), 128B (or 64B) swizzle,
// zero fill out of bounds.  strides are in BYTES for dims 1..rank-1.
const CUtensorMap* get_tensor_map(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                  const uint32_t* box, int swizzle_bytes = 128);

}  // namespace tc_host

#define TC_CHECK_ARG(cond, msg)                                      \
    do {                                                             \
        if (!(cond)) return tc_host::fail(TC_ERR_INVALID, (msg));    \
    } while (0)

#define TC_CHECK_LAUNCH(what)                                        \
    do {                                                             \
        int _rc = tc_host::check_cuda(cudaGetLastError(), (what));   \
        if (_rc) return _rc;                                         \
    } while (0)
