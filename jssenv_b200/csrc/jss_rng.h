// Counter-based RNG shared by the device policy kernels and the host helper
// jss_host_masked_random().  One 32-bit draw per (seed, global env id, step
// index): a splitmix64-style finaliser over a linear combination of the three
// counters.  Stateless, so any env's action stream can be replayed on the host
// (the CPU oracle restates the same function for its replays).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) || defined(JSS_EMU)
#define JSS_HD __host__ __device__
#else
#define JSS_HD
#endif

JSS_HD static inline uint32_t jss_hash3(uint64_t seed, uint64_t env, uint64_t ctr) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (env + 1) + 0xD1B54A32D192ED03ull * (ctr + 1);
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

// index in [0, count) -- multiply-high, no modulo bias worth speaking of
JSS_HD static inline uint32_t jss_pick(uint32_t h, uint32_t count) {
    return (uint32_t)(((uint64_t)h * (uint64_t)count) >> 32);
}

// the rules' exploration coin (dispatching.py:113): u = h / 2^32 < 0.1
#define JSS_COIN_THRESHOLD 429496730u
