// TEST INFRASTRUCTURE: compiles the per-position logic of the K4 match finder (libagmv_b200/csrc/lzchain_core.h - the
// functions the CUDA kernels call per thread) for the host and checks it against the reference's brute-force search
// (AGMV_LZSS, src/agmv_encode.c:106-177, restated in brute() below) on synthetic buffers: for EVERY position the longest
// match length, and for every position with a match the earliest start among the longest.
//
//   g++ -O2 -std=c++17 -I libagmv_b200/csrc tests/lzchain_host_check.cpp -o /tmp/lzchain_host_check && /tmp/lzchain_host_check [file...]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "lzchain_core.h"

using namespace agmvb;

static const int HB = 13;

struct Result { std::vector<uint8_t> len; std::vector<uint32_t> off; };

// the pipeline as the kernels run it; frames = starts of the frames inside d (last entry = n)
static Result chain_pipeline(const std::vector<uint8_t>& data, const std::vector<uint32_t>& fs) {
    const size_t n = fs.back();
    std::vector<uint8_t> d(data);
    d.resize(n + 32, 0xA5);
    std::vector<uint32_t> lwh(n), lw[2];
    lw[0].resize(n); lw[1].resize(n);
    std::vector<uint16_t> rsd(n);
    Result R; R.len.assign(n, 0); R.off.assign(n, 0);
    std::vector<uint32_t> rem(n), fstart(n);
    for (size_t f = 0; f + 1 < fs.size(); f++) {
        const size_t base = fs[f], len = fs[f + 1] - fs[f];
        for (size_t i = 0; i < len; i++) fstart[base + i] = (uint32_t)base;
        std::vector<uint32_t> tab((size_t)1 << HB, 0);
        size_t runstart = 0;
        for (size_t i = 0; i < len; i++) {
            const size_t p = base + i;
            rem[p] = (uint32_t)(len - i);
            if (i == 0 || d[p] != d[p - 1]) runstart = i;
            rsd[p] = (uint16_t)std::min<size_t>(i - runstart, 65535);
            uint32_t dist = 0;
            if (i + 3 <= len) {
                const uint32_t h = lzc_hash(d[p] | d[p + 1] << 8 | d[p + 2] << 16, HB);
                const uint32_t old = tab[h];
                if (old && i - (old - 1) <= LZC_WINDOW) dist = (uint32_t)(i - (old - 1));
                tab[h] = (uint32_t)i + 1;
            }
            lwh[p] = dist | (uint32_t)d[p] << 16 | (uint32_t)d[p + 1] << 24;
        }
    }
    for (size_t p = 0; p < n; p++) {
        const uint32_t cap = std::min<uint32_t>(rem[p], 15u);
        LzcLink3Walk wk;
        uint32_t d3 = 0;
        if (wk.start((uint32_t)p, lwh[p], d[p + 2], cap)) {
            int r;
            while ((r = wk.hop(d.data(), lwh.data(), rsd.data())) == LZC_GO) {}
            if (r == LZC_FOUND) d3 = wk.acc;
        }
        lw[0][p] = d3 ? lzc_word(d3, d[p + 3], d[p + 2], cap) : lzc_dead(0u, 0u, d[p + 3]);
    }
    int cur = 0;
    for (uint32_t L = 3; L < 15; L++) {
        for (size_t p = 0; p < n; p++) {
            const uint32_t w = lw[cur][p];
            LzcLevelWalk wk;
            // the sweep's two shortcuts (lzchain.cuh, LzcLevelOp::sweep): first hop, then the right neighbour's candidate
            const uint32_t dist = lzc_link(w), cap = (w >> 24) & 0xFu;
            if (dist && L + 1u <= cap && ((lw[cur][p - dist] >> 16) & 0xFFu) == ((w >> 16) & 0xFFu)) {
                lw[cur ^ 1][p] = lzc_word(dist, d[p + L + 1], d[p + L], cap);
                continue;
            }
            if (dist && p + 1 < n) {
                const uint32_t wr = lw[cur][p + 1], xr = wr & 0xFFFFu;
                const bool cand = (wr & LZC_DEAD) ? (L >= 4u && ((wr >> 24) & 0xFu) == L - 1u) : (L + 1u <= cap);
                if (cand && xr && xr <= p && d[p - xr] == d[p] && p - xr >= fstart[p]) {
                    lw[cur ^ 1][p] = (wr & LZC_DEAD) ? lzc_dead(L, xr, d[p + L + 1]) : lzc_word(xr, d[p + L + 1], d[p + L], cap);
                    continue;
                }
            }
            if (wk.start((uint32_t)p, w, L)) {
                int r;
                while ((r = wk.hop(lw[cur].data(), rsd.data())) == LZC_GO) {}
                lw[cur ^ 1][p] = r == LZC_FOUND ? lzc_word(wk.acc, d[p + L + 1], d[p + L], (w >> 24) & 0xFu) : lzc_dead(L, wk.last, d[p + L + 1]);
            } else lw[cur ^ 1][p] = lzc_carry(w, d[p + L + 1]);
        }
        cur ^= 1;
    }
    for (size_t p = 0; p < n; p++) {   // level-15 words: dead = (length, offset), live = a 15-byte match whose earliest start ends its chain
        const uint32_t w = lw[cur][p];
        if (w & LZC_DEAD) { R.len[p] = (uint8_t)((w >> 24) & 0xFu); R.off[p] = R.len[p] ? (w & 0xFFFFu) : 0u; }
        else { R.len[p] = 15; R.off[p] = lzc_chain_end(lw[cur].data(), rsd.data(), (uint32_t)p); }
    }
    return R;
}

// src/agmv_encode.c:119-143 for every position (not only the parse positions)
static Result brute(const std::vector<uint8_t>& data, const std::vector<uint32_t>& fs) {
    const size_t n = fs.back();
    Result R; R.len.assign(n, 0); R.off.assign(n, 0);
    const uint8_t* d = data.data();
    for (size_t f = 0; f + 1 < fs.size(); f++) {
        const size_t base = fs[f], len = fs[f + 1] - fs[f];
        const uint8_t* q = d + base;
        for (size_t i = 0; i < len; i++) {
            const size_t cap = std::min<size_t>(15, len - i);
            if (cap < 3) continue;
            size_t best = 0, bs = 0;
            const size_t lo = i > LZC_WINDOW ? i - LZC_WINDOW : 0;
            for (size_t s = lo; s < i; s++) {
                if (q[s] != q[i] || q[s + 1] != q[i + 1] || q[s + 2] != q[i + 2]) continue;
                size_t l = 3;
                while (l < cap && q[s + l] == q[i + l]) l++;
                if (l > best) { best = l; bs = s; if (l == cap) break; }
            }
            if (best >= 3) { R.len[base + i] = (uint8_t)best; R.off[base + i] = (uint32_t)(i - bs); }
        }
    }
    return R;
}

static uint32_t rng_state = 12345;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

static int check(const char* name, const std::vector<uint8_t>& data, const std::vector<uint32_t>& fs) {
    Result a = chain_pipeline(data, fs), b = brute(data, fs);
    size_t nm = 0;
    for (size_t p = 0; p < fs.back(); p++) {
        if (a.len[p] != b.len[p] || (b.len[p] && a.off[p] != b.off[p])) {
            if (nm < 5) printf("  %s: p=%zu chain (%u,%u) brute (%u,%u)\n", name, p, a.len[p], a.off[p], b.len[p], b.off[p]);
            nm++;
        }
    }
    size_t matched = 0;
    for (size_t p = 0; p < fs.back(); p++) matched += b.len[p] != 0;
    printf("%-28s n=%8u frames=%2zu matched=%8zu %s\n", name, fs.back(), fs.size() - 1, matched, nm ? "MISMATCH" : "ok");
    return nm ? 1 : 0;
}

int main(int argc, char** argv) {
    int bad = 0;
    for (int k = 1; k < argc; k++) {   // files: one frame each
        FILE* f = fopen(argv[k], "rb");
        if (!f) { printf("cannot open %s\n", argv[k]); return 2; }
        std::vector<uint8_t> d;
        uint8_t buf[65536];
        size_t r;
        while ((r = fread(buf, 1, sizeof buf, f)) > 0) d.insert(d.end(), buf, buf + r);
        fclose(f);
        bad += check(argv[k], d, {0u, (uint32_t)d.size()});
    }
    if (argc > 1) return bad ? 1 : 0;
    auto gen = [&](size_t n, int alphabet) { std::vector<uint8_t> d(n); for (auto& x : d) x = (uint8_t)(rnd() % alphabet); return d; };
    bad += check("empty", {}, {0u, 0u});
    bad += check("tiny", gen(2, 2), {0u, 2u});
    bad += check("alphabet2", gen(30000, 2), {0u, 30000u});
    bad += check("alphabet3", gen(90000, 3), {0u, 90000u});
    bad += check("alphabet16", gen(100000, 16), {0u, 100000u});
    bad += check("alphabet256", gen(80000, 256), {0u, 80000u});
    {   // runs of random lengths (1..40) of few byte values, some very long
        std::vector<uint8_t> d;
        while (d.size() < 150000) { size_t r = 1 + rnd() % 40; if (rnd() % 200 == 0) r = 3000 + rnd() % 5000; uint8_t v = (uint8_t)(rnd() % 3); d.insert(d.end(), r, v); }
        bad += check("runs", d, {0u, (uint32_t)d.size()});
    }
    {   // one run longer than the window, then noise, then the run again
        std::vector<uint8_t> d(70000, 0x5E);
        auto t = gen(500, 4); d.insert(d.end(), t.begin(), t.end());
        d.insert(d.end(), 70000, 0x5E);
        d.insert(d.end(), 20, 7);
        bad += check("run > window", d, {0u, (uint32_t)d.size()});
    }
    {   // periodic zones (periods 1..6) with noise between them
        std::vector<uint8_t> d;
        while (d.size() < 140000) {
            size_t per = 1 + rnd() % 6, len = 10 + rnd() % 400;
            if (rnd() % 50 == 0) len = 20000;
            uint8_t pat[6]; for (auto& x : pat) x = (uint8_t)(rnd() % 4);
            for (size_t k = 0; k < len; k++) d.push_back(pat[k % per]);
            size_t nz = rnd() % 6; for (size_t k = 0; k < nz; k++) d.push_back((uint8_t)(rnd() % 8));
        }
        bad += check("periodic", d, {0u, (uint32_t)d.size()});
    }
    {   // repeats 65530..65540 bytes apart (window edge)
        std::vector<uint8_t> d = gen(140000, 256);
        for (size_t gap = 65530; gap <= 65540; gap++) {
            size_t at = 200 + (gap - 65530) * 40;
            for (int j = 0; j < 20; j++) d[at + gap + j] = d[at + j];
        }
        bad += check("window edge", d, {0u, (uint32_t)d.size()});
    }
    {   // several frames back to back: no match may cross a frame start; short frames included
        std::vector<uint8_t> d;
        std::vector<uint32_t> fs{0u};
        size_t lens[] = {5000, 0, 1, 2, 3, 17, 70000, 4000, 66000};
        for (size_t len : lens) {
            for (size_t k = 0; k < len; k++) d.push_back((uint8_t)((rnd() % 100 < 70) ? 0x5E : rnd() % 5));
            fs.push_back((uint32_t)d.size());
        }
        bad += check("multi-frame", d, fs);
    }
    {   // block-record-like data: flags, short codes, runs of 0x5E
        std::vector<uint8_t> d;
        while (d.size() < 200000) {
            int t = rnd() % 10;
            if (t < 4) { size_t r = 1 + rnd() % 300; d.insert(d.end(), r, 0x5E); }
            else if (t < 7) { d.push_back(0x4E); d.push_back((uint8_t)(0x80 | (rnd() % 6))); }
            else { d.push_back(0x2F); uint8_t c = (uint8_t)(rnd() % 40); for (int k = 0; k < 16; k++) { if (rnd() % 4 == 0) c = (uint8_t)(rnd() % 40); d.push_back(c); } }
        }
        bad += check("records", d, {0u, (uint32_t)d.size()});
    }
    printf(bad ? "FAILED\n" : "all ok\n");
    return bad ? 1 : 0;
}
