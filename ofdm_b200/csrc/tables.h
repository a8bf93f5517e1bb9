// tables.h -- host-side generation of the frame's constant tables (product code; the oracle has its own copy).
//
//   locking_signal::<80>   src/transmitter.rs:60-72   real ramp 0.25..0.497, fft_shift'ed
//   preamble::<80>         src/transmitter.rs:75-84   StdRng(seed 100), (U(-1,1) + jU(-1,1)) * 0.25, time domain
//   training_signals::<64> src/transmitter.rs:88-96   StdRng(seed 50),  U(-1,1) + jU(-1,1), frequency domain
//
// StdRng of rand 0.8 = ChaCha12 (rand_chacha 0.3) keyed by rand_core 0.6's PCG32 expansion of the u64 seed. The
// crates are not vendored in the reference and nothing in it pins the table values. Pinned by known answers
// (tests/test_abi_host.py::test_tables_stdrng_known_answers): the ChaCha12 block function (published zero-key vector) and
// StdRng::from_seed -> next_u64 (the two values of rand 0.8's own `test_stdrng_construction`). Not pinned offline: the
// PCG32 expansion of seed_from_u64 and the u64 -> f64 range conversion; the engine therefore also accepts the three
// tables as configuration (ofdm_cfg).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace ofdm_host {

class StdRng {
public:
    explicit StdRng(uint64_t seed)
    {
        for (int i = 0; i < 8; i++) {
            seed = seed * 6364136223846793005ULL + 11634580027462260723ULL;
            uint32_t xs = (uint32_t)(((seed >> 18) ^ seed) >> 27);
            uint32_t rot = (uint32_t)(seed >> 59);
            key_[i] = rot ? ((xs >> rot) | (xs << (32 - rot))) : xs;
        }
    }
    // StdRng::from_seed: the 32 seed bytes are the ChaCha key (little-endian words); used by the known-answer test
    static StdRng from_seed(const uint8_t seed[32])
    {
        StdRng g(0);
        for (int i = 0; i < 8; i++)
            g.key_[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
        return g;
    }
    uint64_t next_u64()
    {
        uint64_t lo = next_u32(), hi = next_u32();
        return lo | (hi << 32);
    }
    // gen_range(-1.0..1.0)
    double range_pm1()
    {
        for (;;) {
            uint64_t bits = (next_u64() >> 12) | 0x3FF0000000000000ULL;
            double v;
            std::memcpy(&v, &bits, 8);
            double r = (v - 1.0) * 2.0 - 1.0;
            if (r < 1.0) return r;
        }
    }

private:
    static uint32_t rl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
    static void quarter(uint32_t *x, int a, int b, int c, int d)
    {
        x[a] += x[b]; x[d] = rl(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rl(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rl(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rl(x[b] ^ x[c], 7);
    }
    void refill()
    {
        uint32_t in[16] = { 0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u };
        for (int i = 0; i < 8; i++) in[4 + i] = key_[i];
        in[12] = (uint32_t)ctr_; in[13] = (uint32_t)(ctr_ >> 32); in[14] = 0; in[15] = 0;
        uint32_t x[16];
        std::memcpy(x, in, sizeof x);
        for (int r = 0; r < 6; r++) {
            quarter(x, 0, 4, 8, 12); quarter(x, 1, 5, 9, 13); quarter(x, 2, 6, 10, 14); quarter(x, 3, 7, 11, 15);
            quarter(x, 0, 5, 10, 15); quarter(x, 1, 6, 11, 12); quarter(x, 2, 7, 8, 13); quarter(x, 3, 4, 9, 14);
        }
        for (int i = 0; i < 16; i++) buf_[i] = x[i] + in[i];
        ctr_++;
        pos_ = 0;
    }
    uint32_t next_u32()
    {
        if (pos_ >= 16) refill();
        return buf_[pos_++];
    }
    uint32_t key_[8];
    uint64_t ctr_ = 0;
    uint32_t buf_[16];
    int pos_ = 16;
};

struct cd { double re, im; };

inline void locking_signal(cd *out, int len)
{
    std::vector<cd> t(len);
    for (int i = 0; i < len; i++) t[i] = { 0.5 * ((double)i / (2.0 * len) + 0.5), 0.0 };
    int mid = (len + 1) / 2;                            // fft_shift, src/signals/mod.rs:61-77
    for (int i = 0; i < len; i++) out[i] = t[(i + mid) % len];
}
inline void preamble(cd *out, int len)
{
    StdRng g(100);
    for (int i = 0; i < len; i++) { double a = g.range_pm1(), b = g.range_pm1(); out[i] = { a * 0.25, b * 0.25 }; }
}
inline void training_signals(cd *out, int len)
{
    StdRng g(50);
    for (int i = 0; i < len; i++) { double a = g.range_pm1(), b = g.range_pm1(); out[i] = { a, b }; }
}

// scaled inverse DFT (src/signals/mod.rs:49-58), O(N^2): only used once per engine for the training symbol
inline void idft(const cd *in, cd *out, int n)
{
    const double PI = 3.14159265358979323846;
    for (int t = 0; t < n; t++) {
        double sr = 0, si = 0;
        for (int k = 0; k < n; k++) {
            double a = 2.0 * PI * (double)((long long)k * t % n) / n;
            double c = std::cos(a), s = std::sin(a);
            sr += in[k].re * c - in[k].im * s;
            si += in[k].re * s + in[k].im * c;
        }
        out[t] = { sr / n, si / n };
    }
}

}  // namespace ofdm_host
