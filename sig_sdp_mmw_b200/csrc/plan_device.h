// Device-side plan builder (plan_device.cu): the same HostPlan / device arrays as build_host_plan + upload,
// produced by sorts, scans and per-row merges on the GPU while the locality ordering runs on a host thread.
#pragma once
#include <functional>

#include "plan_host.h"

namespace sigsdp {

struct DevicePlanArrays {   // all inside one slab obtained from DevicePlanAlloc::device
    int *rowptr = nullptr, *col = nullptr, *eid = nullptr, *perm = nullptr, *dpos = nullptr, *apos = nullptr;
    double *tfwd = nullptr, *tbwd = nullptr, *S_sum = nullptr, *tnorm = nullptr, *h_max = nullptr;
    int *gi = nullptr, *gj = nullptr, *ai = nullptr, *aj = nullptr;   // edge lists, caller numbering
    double *tij = nullptr, *tji = nullptr;
};
struct DevicePlanAlloc {
    void* ctx = nullptr;
    void* (*pinned)(void* ctx, size_t bytes) = nullptr;   // process-wide staging buffer (valid until the next call)
    void* (*device)(void* ctx, size_t bytes) = nullptr;   // the plan's persistent slab (called once)
    int num_sms = 148;
    std::function<void()> overlap;                        // host work to do while the device sorts (may be empty)
};
// Fills P with what the host needs at once (sizes, rowptr, col, dpos, node vectors, perm / iperm); eid, tfwd, tbwd,
// apos and the edge lists stay on the device until fetch_device_plan_rest.  0 or a negative SIGSDP_E* code.
int build_device_plan(int64_t n, const int32_t* Sp, const int32_t* Si, const double* Sx, const int32_t* Qp, const int32_t* Qi,
                      const double* Qx, const double* h_max, int order, const DevicePlanAlloc& alloc, HostPlan& P, DevicePlanArrays& D,
                      std::string& err);
int fetch_device_plan_rest(const DevicePlanArrays& D, const DevicePlanAlloc& alloc, HostPlan& P, std::string& err);

}  // namespace sigsdp
