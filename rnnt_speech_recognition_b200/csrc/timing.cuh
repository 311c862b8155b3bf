// timing.cuh -- optional per-kernel CUDA-event instrumentation (off by default).  bench.py switches
// it on for ONE untimed pass to attribute the step time to kernels; events are recorded on the
// launching stream, so they see exactly the kernel's device-side duration.
#pragma once
#include <cuda_runtime.h>

#include <vector>

namespace rb {

struct TimingRec { const char* name; cudaEvent_t a, b; };
inline int& timing_on() { static int on = 0; return on; }
inline std::vector<TimingRec>& timing_recs() { static std::vector<TimingRec> v; return v; }

struct ScopedTimer {
    cudaStream_t s;
    int idx;
    ScopedTimer(const char* name, cudaStream_t stream) : s(stream), idx(-1) {
        if (!timing_on()) return;
        TimingRec r{name, nullptr, nullptr};
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        cudaEventRecord(r.a, s);
        timing_recs().push_back(r);
        idx = (int)timing_recs().size() - 1;
    }
    ~ScopedTimer() {
        if (idx >= 0) cudaEventRecord(timing_recs()[idx].b, s);
    }
};
inline void timing_reset(int on) {
    for (auto& r : timing_recs()) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    timing_recs().clear();
    timing_on() = on;
}

}  // namespace rb
