/* np_rng.h -- numpy's PCG64 Generator restated (CPU ORACLE, test infrastructure, NOT product code).
 * Generator(PCG64(SeedSequence(seed))): next_uint64 / next_double / uniform as in gym_oracle.c, plus
 * the buffered next_uint32 and Lemire's bounded integers that Generator.integers() uses for ranges
 * below 2^32 (BipedalWalker's terrain counters, bipedal_walker.py:301,328,342-343,379,381).
 * Checked against numpy 2.3.5 in tests/test_oracle_golden.py. */
#ifndef NP_RNG_H
#define NP_RNG_H
#include <stdint.h>
typedef unsigned __int128 u128;
typedef struct { u128 state, inc; int has_uint32; uint32_t uinteger; } pcg64_t;
#define PCG_MULT ((((u128)0x2360ED051FC65DA4ULL) << 64) | (u128)0x4385DF649FCCF645ULL)
static inline void pcg_adv(pcg64_t *g) { g->state = g->state * PCG_MULT + g->inc; }
static inline uint64_t pcg_next64(pcg64_t *g)
{
    pcg_adv(g);
    uint64_t hi = (uint64_t)(g->state >> 64), lo = (uint64_t)g->state, x = hi ^ lo;
    unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((-rot) & 63));
}
/* pcg64_next32: the low half of a fresh draw now, the high half on the next call */
static inline uint32_t pcg_next32(pcg64_t *g)
{
    if (g->has_uint32) { g->has_uint32 = 0; return g->uinteger; }
    uint64_t next = pcg_next64(g);
    g->has_uint32 = 1;
    g->uinteger = (uint32_t)(next >> 32);
    return (uint32_t)(next & 0xffffffffu);
}
static inline double pcg_double(pcg64_t *g) { return (double)(pcg_next64(g) >> 11) * (1.0 / 9007199254740992.0); }
static inline double rng_uniform(pcg64_t *g, double lo, double hi) { double r = hi - lo; return lo + r * pcg_double(g); }
/* Generator.integers(low, high) for high - low <= 2^32: Lemire's nearly-divisionless rejection */
static inline int64_t rng_integers(pcg64_t *g, int64_t low, int64_t high)
{
    uint32_t rng = (uint32_t)(high - low - 1); /* inclusive span */
    if (rng == 0) return low;
    const uint32_t rng_excl = rng + 1u;
    uint64_t m = (uint64_t)pcg_next32(g) * rng_excl;
    uint32_t leftover = (uint32_t)m;
    if (leftover < rng_excl) {
        const uint32_t threshold = (0xFFFFFFFFu - rng) % rng_excl;
        while (leftover < threshold) { m = (uint64_t)pcg_next32(g) * rng_excl; leftover = (uint32_t)m; }
    }
    return low + (int64_t)(m >> 32);
}
void orc_seed_sequence(const uint32_t ent[4], uint64_t out[4]); /* gym_oracle.c */
static inline void pcg_seed_from_words(pcg64_t *g, const uint32_t ent[4])
{
    uint64_t w[4];
    orc_seed_sequence(ent, w);
    g->state = 0;
    g->inc = (((((u128)w[2]) << 64) | w[3]) << 1) | 1;
    pcg_adv(g);
    g->state += (((u128)w[0]) << 64) | w[1];
    pcg_adv(g);
    g->has_uint32 = 0;
    g->uinteger = 0;
}
#endif
