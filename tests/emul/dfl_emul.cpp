// dfl_emul.cpp -- CPU checks of the shared DEFLATE block coder (caesium-clt_b200/csrc/dfl_core.h): its two-queue code-length
// construction against the binary-heap formulation it replaced (the round-1 host writer), on arbitrary histograms.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>
#include "../../caesium-clt_b200/csrc/dfl_core.h"

namespace {
// round-1 host writer: plain Huffman on a heap ordered by (weight, creation index), IJG/zlib overflow repair, lengths by rank
void huff_lengths_heap(const uint32_t *freq, int n, int limit, uint8_t *len)
{
    struct Node { uint64_t w; int l, r; };
    std::vector<Node> nodes; std::vector<int> alive;
    for (int i = 0; i < n; i++) { len[i] = 0; if (freq[i]) { nodes.push_back({freq[i], -1, i}); alive.push_back((int)nodes.size() - 1); } }
    if (alive.empty()) return;
    if (alive.size() == 1) { len[nodes[alive[0]].r] = 1; return; }
    auto cmp = [&](int a, int b) { return nodes[a].w > nodes[b].w || (nodes[a].w == nodes[b].w && a > b); };
    std::make_heap(alive.begin(), alive.end(), cmp);
    while (alive.size() > 1) {
        std::pop_heap(alive.begin(), alive.end(), cmp); int a = alive.back(); alive.pop_back();
        std::pop_heap(alive.begin(), alive.end(), cmp); int b = alive.back(); alive.pop_back();
        nodes.push_back({nodes[a].w + nodes[b].w, a, b});
        alive.push_back((int)nodes.size() - 1); std::push_heap(alive.begin(), alive.end(), cmp);
    }
    std::vector<int> depth(nodes.size(), 0);
    std::vector<int> stack{alive[0]};
    int bl[64] = {0};
    std::vector<std::pair<uint32_t, int>> leaves;
    while (!stack.empty()) {
        int x = stack.back(); stack.pop_back();
        if (nodes[x].l < 0) { bl[std::min(depth[x], 63)]++; leaves.push_back({freq[nodes[x].r], nodes[x].r}); }
        else { depth[nodes[x].l] = depth[nodes[x].r] = depth[x] + 1; stack.push_back(nodes[x].l); stack.push_back(nodes[x].r); }
    }
    for (int i = 63; i > limit; i--) while (bl[i] > 0) {
        int j = i - 2; while (bl[j] == 0) j--;
        bl[i] -= 2; bl[i - 1]++; bl[j + 1] += 2; bl[j]--;
    }
    std::sort(leaves.begin(), leaves.end(), [](const std::pair<uint32_t, int> &a, const std::pair<uint32_t, int> &b) { return a.first > b.first || (a.first == b.first && a.second < b.second); });
    size_t k = 0;
    for (int l = 1; l <= limit; l++) for (int c = 0; c < bl[l]; c++) len[leaves[k++].second] = (uint8_t)l;
}
} // namespace

extern "C" {
// 0 = equal; otherwise 1 + the first differing symbol
int emul_huff_lengths_compare(const uint32_t *freq, int n, int limit)
{
    uint8_t a[b200::dfl::MAXSYM], b[b200::dfl::MAXSYM];
    static b200::dfl::HuffScratch S;
    std::vector<uint32_t> f(freq, freq + n);
    b200::dfl::huff_lengths(f.data(), n, limit, a, S);
    huff_lengths_heap(freq, n, limit, b);
    for (int i = 0; i < n; i++) if (a[i] != b[i]) return 1 + i;
    return 0;
}
int emul_len_dist_tables(void)
{   // closed forms against RFC 1951's tables
    static const uint16_t LB[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t LX[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t DB[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t DX[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    using namespace b200::dfl;
    for (int s = 0; s < 29; s++) if (len_base(s) != LB[s] || len_extra(s) != LX[s]) return 1;
    for (int s = 0; s < 30; s++) if (dist_base(s) != DB[s] || dist_extra(s) != DX[s]) return 2;
    for (int l = 3; l <= 258; l++) { int s = 0; while (s < 28 && LB[s + 1] <= l) s++; if (l == 258) s = 28; if (len_sym(l) != s) return 3; }
    for (int d = 1; d <= 32768; d++) { int s = 0; while (s < 29 && DB[s + 1] <= d) s++; if (dist_sym(d) != s) return 4; }
    return 0;
}
}
