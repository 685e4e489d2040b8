#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "vp8_decode.h"
#include "vp8l_decode.h"
#include "vp8l_alpha.h"
using namespace b200;
static std::vector<uint8_t> slurp(const char *p) { FILE *f = fopen(p, "rb"); std::vector<uint8_t> v; if (!f) return v; fseek(f, 0, SEEK_END); v.resize(ftell(f)); fseek(f, 0, SEEK_SET); if (fread(v.data(), 1, v.size(), f)) {} fclose(f); return v; }
int main(int argc, char **argv)
{
    std::mt19937 rng(12345);
    long ok = 0, bad = 0;
    for (int a = 1; a < argc; a++) {
        const std::vector<uint8_t> src = slurp(argv[a]);
        if (src.empty()) continue;
        for (int it = 0; it < 4000; it++) {
            std::vector<uint8_t> d = src;
            const int mode = rng() % 4;
            if (mode == 0) for (int k = 0; k < 1 + (int)(rng() % 6); k++) d[rng() % d.size()] = (uint8_t)rng();
            else if (mode == 1) d.resize(1 + rng() % d.size());
            else if (mode == 2) { const size_t i = rng() % d.size(); d.insert(d.begin() + i, (size_t)(1 + rng() % 40), (uint8_t)rng()); }
            else { for (int k = 0; k < 3; k++) { const size_t i = 30 + rng() % (d.size() > 31 ? d.size() - 30 : 1); if (i < d.size()) d[i] ^= (uint8_t)(1u << (rng() % 8)); } }
            // keep dimensions sane so that a flipped header bit does not ask for gigabytes
            WebpInfo info; std::vector<uint8_t> rgb, alpha; std::string err;
            WebpInfo probe; std::string e2;
            if (webp_probe(d.data(), d.size(), probe, e2) && (long long)probe.width * probe.height > 4000000) continue;
            if (d.size() > 25 && !memcmp(d.data() + 12, "VP8L", 4)) { const uint32_t h = d[21] | (d[22] << 8) | (d[23] << 16) | ((uint32_t)d[24] << 24); if ((long long)((h & 0x3FFF) + 1) * (((h >> 14) & 0x3FFF) + 1) > 4000000) continue; }
            const int rc = webp_decode_rgb(d.data(), d.size(), info, rgb, err, &alpha);
            if (rc == 0) ok++; else bad++;
        }
    }
    // alpha coder: random token streams must be rejected or coded, never crash
    for (int it = 0; it < 2000; it++) {
        const int w = 1 + rng() % 40, h = 1 + rng() % 40; std::vector<uint32_t> tok; size_t pos = 0; const size_t n = (size_t)w * h;
        while (pos < n) { if (pos > 0 && rng() % 3 == 0) { uint32_t len = 3 + rng() % 256; if (len > n - pos) len = (uint32_t)(n - pos); if (len < 3) { tok.push_back(rng() & 0xFF); pos++; continue; } const uint32_t dist = 1 + rng() % pos; tok.push_back(0x80000000u | ((len - 3) << 16) | ((dist - 1) & 0xFFFF)); pos += len; } else { tok.push_back(rng() & 0xFF); pos++; } }
        std::vector<uint8_t> alph; if (!vp8l_alpha_from_tokens(tok.data(), tok.size(), w, h, alph, rng() % 4)) { printf("alpha coder refused a valid stream\n"); return 1; }
        std::vector<uint8_t> back; std::string err;
        if (!webp_alpha_decode(alph.data(), alph.size(), w, h, back, err)) { printf("own decoder refused own alpha chunk: %s\n", err.c_str()); return 1; }
    }
    printf("decoded %ld, refused %ld\n", ok, bad);
    return 0;
}
