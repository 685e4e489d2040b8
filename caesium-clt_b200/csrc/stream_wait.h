// stream_wait.h -- how a caller thread waits for its stream.  The path blocks a handful of times per image or megabatch
// (sizes that the host needs before it can size the next launch, finished scans).  cudaStreamSynchronize either spins
// (burning a core per waiting thread -- the reference's rayon pool has one thread per core, all of them waiting) or blocks
// in the kernel (50-100 us of wake-up latency per wait, which dominates a 3 ms megabatch).  Hybrid: poll the stream for a
// short while, then back off to short sleeps.  B200_SYNC=spin | block | hybrid (default) selects the behaviour.
#pragma once
#include <cuda_runtime.h>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace b200 {

inline int stream_wait_mode()
{
    static const int mode = [] { const char *e = getenv("B200_SYNC"); return !e ? 2 : !strcmp(e, "spin") ? 0 : !strcmp(e, "block") ? 1 : 2; }();
    return mode;
}

// how long the hybrid mode polls before it starts sleeping (B200_SPIN_US; the runtime lowers it when the process has fewer
// cores than waiting threads, see runtime_init)
inline int &stream_spin_us()
{
    static int us = [] { const char *e = getenv("B200_SPIN_US"); const int v = e ? atoi(e) : -1; return v >= 0 && v <= 100000 ? v : 1500; }();
    return us;
}

inline cudaError_t stream_wait(cudaStream_t st)
{
    if (stream_wait_mode() != 2) return cudaStreamSynchronize(st);
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned i = 0;; i++) {
        const cudaError_t e = cudaStreamQuery(st);
        if (e != cudaErrorNotReady) return e;
        if ((i & 15) == 15 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(stream_spin_us())) break;
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    for (;;) {       // long wait (another worker's kernels are ahead of ours): stop burning the core
        std::this_thread::sleep_for(std::chrono::microseconds(40));
        const cudaError_t e = cudaStreamQuery(st);
        if (e != cudaErrorNotReady) return e;
    }
}

// Wait for an event recorded earlier on the caller's stream while later work of that stream keeps the GPU busy.
inline cudaError_t event_wait(cudaEvent_t ev)
{
    if (stream_wait_mode() != 2) return cudaEventSynchronize(ev);
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned i = 0;; i++) {
        const cudaError_t e = cudaEventQuery(ev);
        if (e != cudaErrorNotReady) return e;
        if ((i & 15) == 15 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(stream_spin_us())) break;
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    for (;;) {
        std::this_thread::sleep_for(std::chrono::microseconds(40));
        const cudaError_t e = cudaEventQuery(ev);
        if (e != cudaErrorNotReady) return e;
    }
}

} // namespace b200
