// topology.cpp -- see topology.h.  Reads /sys/bus/pci/devices/<bus id>/numa_node and /sys/devices/system/node/nodeN/cpulist.
#include "topology.h"
#include <cuda_runtime.h>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <string>

namespace b200 {

namespace {
std::mutex g_mu;
std::map<int, std::pair<int, std::vector<int>>> g_cache;       // ordinal -> (node, cpus)

std::vector<int> parse_cpulist(const std::string &s)
{   // "0-47,96-143"
    std::vector<int> out;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && !isdigit((unsigned char)s[i])) i++;
        if (i >= s.size()) break;
        int a = 0; while (i < s.size() && isdigit((unsigned char)s[i])) a = a * 10 + (s[i++] - '0');
        int b = a;
        if (i < s.size() && s[i] == '-') { i++; b = 0; while (i < s.size() && isdigit((unsigned char)s[i])) b = b * 10 + (s[i++] - '0'); }
        for (int c = a; c <= b && c < CPU_SETSIZE; c++) out.push_back(c);
    }
    return out;
}

bool numa_enabled()
{
    static const bool on = [] { const char *e = getenv("B200_NUMA"); return !(e && !strcmp(e, "0")); }();
    return on;
}

const std::pair<int, std::vector<int>> &lookup(int ordinal)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(ordinal);
    if (it != g_cache.end()) return it->second;
    std::pair<int, std::vector<int>> r{-1, {}};
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, ordinal) == cudaSuccess) {
        for (char *p = bus; *p; p++) *p = (char)tolower((unsigned char)*p);
        std::ifstream f(std::string("/sys/bus/pci/devices/") + bus + "/numa_node");
        int node = -1;
        if (f && (f >> node) && node >= 0) {
            std::ifstream c("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
            std::string line;
            if (c && std::getline(c, line)) {
                cpu_set_t cur; CPU_ZERO(&cur);
                const bool have = sched_getaffinity(0, sizeof cur, &cur) == 0;
                for (int cpu : parse_cpulist(line)) if (!have || CPU_ISSET(cpu, &cur)) r.second.push_back(cpu);
                r.first = node;
            }
        }
    } else cudaGetLastError();
    return g_cache.emplace(ordinal, std::move(r)).first->second;
}
} // namespace

std::vector<int> device_cpus(int ordinal) { return numa_enabled() ? lookup(ordinal).second : std::vector<int>(); }
int device_numa_node(int ordinal) { return lookup(ordinal).first; }

AffinityGuard::AffinityGuard(int ordinal)
{
    const std::vector<int> cpus = device_cpus(ordinal);
    if (cpus.empty()) return;
    CPU_ZERO(&old_);
    if (sched_getaffinity(0, sizeof old_, &old_) != 0) return;
    have_old_ = true;
    cpu_set_t want; CPU_ZERO(&want);
    for (int c : cpus) CPU_SET(c, &want);
    bound_ = sched_setaffinity(0, sizeof want, &want) == 0;
}

AffinityGuard::~AffinityGuard() { if (bound_ && have_old_) sched_setaffinity(0, sizeof old_, &old_); }

} // namespace b200
