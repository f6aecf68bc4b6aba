/*
 * oracle/chacha12.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Restatement of the start-vector random stream the reference draws from
 *   rand = "0.8.5" (Cargo.toml:16)  ->  rand::rngs::StdRng == rand_chacha::ChaCha12Rng (0.3.x)
 * The crates are NOT vendored under /root/reference, so the published algorithm is restated:
 *   - key        : 32 bytes, seed.to_le_bytes() in bytes 0..4, zeros elsewhere   (reference src/lib.rs:1109-1113)
 *   - state      : "expand 32-byte k" constants | key words | 64-bit block counter (from 0) | 64-bit stream id 0
 *   - rounds     : 12 (6 double rounds), output = permuted state + input state
 *   - next_u64   : two consecutive 32-bit output words, low word first
 *   - gen_range(-1.0..1.0) (reference src/lib.rs:1114) == UniformFloat<f64>::sample_single:
 *         v = f64::from_bits((u64 >> 12) | 1023<<52) - 1.0 ; res = v*2.0 + (-1.0) ; exactly one u64 per sample
 * Pinning: ChaCha20/12 zero-key known-answer vectors (tests/test_oracle_rng.py) and, end to end, the 17-digit
 * documented output of the reference for seed 3141 (reference src/lib.rs:259-281), which the oracle reproduces
 * bit-for-bit only if this stream is right.
 */
#include <stdint.h>
#include <string.h>
#include <stddef.h>

#define ROTL32(v, n) (((v) << (n)) | ((v) >> (32 - (n))))
#define QR(a, b, c, d)                                                                            \
    do {                                                                                          \
        a += b; d ^= a; d = ROTL32(d, 16);                                                        \
        c += d; b ^= c; b = ROTL32(b, 12);                                                        \
        a += b; d ^= a; d = ROTL32(d, 8);                                                         \
        c += d; b ^= c; b = ROTL32(b, 7);                                                         \
    } while (0)

/* One 64-byte ChaCha block with `rounds` rounds (20 or 12), 64-bit counter in words 12/13. */
void oracle_chacha_block(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16])
{
    uint32_t in[16], x[16];
    in[0] = 0x61707865u; in[1] = 0x3320646eu; in[2] = 0x79622d32u; in[3] = 0x6b206574u;
    for (int i = 0; i < 8; i++) in[4 + i] = key[i];
    in[12] = (uint32_t)(counter & 0xffffffffu);
    in[13] = (uint32_t)(counter >> 32);
    in[14] = 0; in[15] = 0;
    memcpy(x, in, sizeof x);
    for (int r = 0; r < rounds; r += 2) {
        QR(x[0], x[4], x[8], x[12]);  QR(x[1], x[5], x[9], x[13]);
        QR(x[2], x[6], x[10], x[14]); QR(x[3], x[7], x[11], x[15]);
        QR(x[0], x[5], x[10], x[15]); QR(x[1], x[6], x[11], x[12]);
        QR(x[2], x[7], x[8], x[13]);  QR(x[3], x[4], x[9], x[14]);
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + in[i];
}

typedef struct {
    uint32_t key[8];
    uint64_t counter;
    uint32_t buf[16];
    int pos; /* next unread word in buf, 16 == empty */
} oracle_rng_t;

void oracle_rng_seed_u32(oracle_rng_t *g, uint32_t seed)
{
    memset(g, 0, sizeof *g);
    g->key[0] = seed; /* le bytes 0..4 of the 32-byte key == word 0 */
    g->pos = 16;
}

static uint32_t rng_next_u32(oracle_rng_t *g)
{
    if (g->pos >= 16) {
        oracle_chacha_block(g->key, g->counter, 12, g->buf);
        g->counter += 1;
        g->pos = 0;
    }
    return g->buf[g->pos++];
}

uint64_t oracle_rng_next_u64(oracle_rng_t *g)
{
    uint64_t lo = rng_next_u32(g);
    uint64_t hi = rng_next_u32(g);
    return lo | (hi << 32);
}

/* gen_range(-1.0..1.0) */
double oracle_rng_unit_pm1(oracle_rng_t *g)
{
    uint64_t u = oracle_rng_next_u64(g);
    uint64_t bits = (u >> 12) | ((uint64_t)1023 << 52);
    double v12;
    memcpy(&v12, &bits, sizeof v12);
    double v01 = v12 - 1.0;
    return v01 * 2.0 + (-1.0);
}

/* exported helper for tests: the first n samples for `seed` */
void oracle_start_vector(uint32_t seed, size_t n, double *out)
{
    oracle_rng_t g;
    oracle_rng_seed_u32(&g, seed);
    for (size_t i = 0; i < n; i++) out[i] = oracle_rng_unit_pm1(&g);
}
