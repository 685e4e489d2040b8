// topology.h -- which host CPUs sit next to which GPU.  The per-image path is embarrassingly parallel ("expect
// scaling limited by host cores / PCIe / NUMA, not NVLink: pin each GPU's feeder threads to its NUMA node"): every worker thread
// that drives device d is bound to the CPUs of d's NUMA node, so its pinned staging buffers (first touched by that thread), its
// memcpy into them and the output assembly stay on the socket the GPU hangs off.  B200_NUMA=0 turns the binding off.
#pragma once
#include <sched.h>
#include <vector>

namespace b200 {

// CPUs (restricted to the process's current affinity mask) of the NUMA node CUDA device `ordinal` is attached to; empty when the
// node is unknown (no sysfs, single-node box, numa_node = -1) -- callers then leave the thread where it is.
std::vector<int> device_cpus(int ordinal);
int device_numa_node(int ordinal);          // -1 unknown

// Binds the calling thread to the CPUs of the device's node for the guard's lifetime and restores the previous mask afterwards
// (the batch call borrows its caller's thread as one of the workers).
class AffinityGuard {
public:
    explicit AffinityGuard(int ordinal);
    ~AffinityGuard();
    AffinityGuard(const AffinityGuard &) = delete;
    AffinityGuard &operator=(const AffinityGuard &) = delete;
    bool bound() const { return bound_; }
private:
    cpu_set_t old_; bool have_old_ = false, bound_ = false;
};

} // namespace b200
