// zinc_b200/csrc/rand_compat.cpp -- host-side derivation of the RAA permutations from their seeds.
//
// RaaCode stores two u64 seeds (code_raa.rs:25-29,74-75) and calls shuffle_seeded(&mut row, seed) on every row
// (code_raa.rs:99,101; zip/utils.rs:139-142).  The GPU path wants the value-independent index form once per pp:
//     perm = shuffle_seeded applied to [0, 1, .., n)      =>      shuffled[i] == original[perm[i]]
// A Rust host obtains `perm` from the real `rand` crate and passes it to zipgpu_code_create; this file restates
// rand 0.9.2 (StdRng = ChaCha12 seeded through SeedableRng::seed_from_u64, SliceRandom::shuffle with the
// IncreasingUniform chooser and Canon's u32 range sampler) for hosts that do not have it.  It is host logic,
// executed once per pp, off the hot path.  Parity with the real crate is UNPINNED.
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

namespace zipgpu {

namespace {

inline uint32_t rol32(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
inline void quarter_round(uint32_t *x, int a, int b, int c, int d) {
    x[a] += x[b]; x[d] = rol32(x[d] ^ x[a], 16);
    x[c] += x[d]; x[b] = rol32(x[b] ^ x[c], 12);
    x[a] += x[b]; x[d] = rol32(x[d] ^ x[a], 8);
    x[c] += x[d]; x[b] = rol32(x[b] ^ x[c], 7);
}

}  // namespace

// One ChaCha block as rand_chacha 0.9 lays the state out: constants "expand 32-byte k", 256-bit key, 64-bit block
// counter in words 12-13, 64-bit stream id (0) in words 14-15.  Exposed (zipgpu_chacha_block) so that the core can be
// pinned to published vectors: 12 rounds, zero key and counter give the ChaCha12 test vector 9bf49a6a 0755f953 ...
void chacha_block(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16]) {
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    std::memcpy(in + 4, key, 8 * sizeof(uint32_t));
    in[12] = static_cast<uint32_t>(counter);
    in[13] = static_cast<uint32_t>(counter >> 32);
    in[14] = in[15] = 0;  // stream id 0
    uint32_t x[16];
    std::memcpy(x, in, sizeof x);
    for (int dr = 0; dr < rounds / 2; ++dr) {
        quarter_round(x, 0, 4, 8, 12); quarter_round(x, 1, 5, 9, 13); quarter_round(x, 2, 6, 10, 14); quarter_round(x, 3, 7, 11, 15);
        quarter_round(x, 0, 5, 10, 15); quarter_round(x, 1, 6, 11, 12); quarter_round(x, 2, 7, 8, 13); quarter_round(x, 3, 4, 9, 14);
    }
    for (int i = 0; i < 16; ++i) out[i] = x[i] + in[i];
}

namespace {

class ChaCha12Stream {
  public:
    explicit ChaCha12Stream(uint64_t seed) {
        // rand_core 0.9: the u64 is stretched to the 256-bit key with a PCG32 generator
        uint64_t st = seed;
        for (int i = 0; i < 8; ++i) {
            st = st * 6364136223846793005ull + 11634580027462260723ull;
            const uint32_t xs = static_cast<uint32_t>(((st >> 18) ^ st) >> 27);
            const unsigned rot = static_cast<unsigned>(st >> 59);
            key_[i] = (xs >> rot) | (xs << ((32u - rot) & 31u));
        }
    }

    uint32_t next() {
        if (pos_ == kBuf) refill();
        return buf_[pos_++];
    }

    // uniform in [0, bound), rand 0.9 `random_range(..bound)` for u32 (widening multiply, one bias-reducing retry)
    uint32_t below(uint32_t bound) {
        const uint64_t prod = static_cast<uint64_t>(next()) * bound;
        uint32_t hi = static_cast<uint32_t>(prod >> 32);
        const uint32_t lo = static_cast<uint32_t>(prod);
        if (lo > static_cast<uint32_t>(0u - bound)) {
            const uint32_t extra = static_cast<uint32_t>((static_cast<uint64_t>(next()) * bound) >> 32);
            if (static_cast<uint64_t>(lo) + extra > 0xffffffffull) ++hi;
        }
        return hi;
    }

  private:
    static constexpr int kBuf = 64;  // rand_chacha refills four 16-word blocks at a time

    void refill() {
        for (int blk = 0; blk < kBuf / 16; ++blk, ++counter_) chacha_block(key_, counter_, 12, buf_ + 16 * blk);
        pos_ = 0;
    }

    uint32_t key_[8];
    uint64_t counter_ = 0;
    uint32_t buf_[kBuf];
    int pos_ = kBuf;
};

}  // namespace

// Forward Fisher-Yates as rand 0.9 performs it: position i is swapped with an index drawn uniformly from
// [0, i]; several consecutive indices are carved out of one u32 draw whose range is the product
// (i+1)(i+2)...(i+c) for the largest c that still fits 32 bits.
void perm_from_seed(uint64_t seed, uint32_t n, uint32_t *perm) {
    for (uint32_t i = 0; i < n; ++i) perm[i] = i;
    if (n < 2) return;
    ChaCha12Stream rng(seed);
    uint32_t pending = 0;  // indices still to be carved out of `pool`
    uint32_t pool = 0;
    // the very first index (range [0,0]) is 0 and consumes no randomness
    for (uint32_t i = 1; i < n; ++i) {
        const uint32_t range = i + 1;
        uint32_t j;
        if (pending == 0) {
            uint64_t prod = range;
            uint32_t cnt = 1;
            while (prod * (static_cast<uint64_t>(range) + cnt) <= 0xffffffffull) {
                prod *= static_cast<uint64_t>(range) + cnt;
                ++cnt;
            }
            pool = rng.below(static_cast<uint32_t>(prod));
            pending = cnt;
        }
        if (pending == 1) {
            j = pool;
        } else {
            j = pool % range;
            pool /= range;
        }
        --pending;
        std::swap(perm[i], perm[j]);
    }
}

}  // namespace zipgpu
