// unrank_selftest.cu -- host-side exhaustive check of the seat-first unranking arithmetic
// (bridge_mc_b200/csrc/seatfirst.cuh: sf_split, sf_honours_of, sf_unrank_lows, make_lowtabs).
// Test infrastructure: compiled and run by tests/test_host.py with nvcc's HOST compiler; no GPU needed.
//
//   1. sf_split(d) == (floor(d / S) mod I, floor(d / S) / I) for every I that occurs, at random d and at
//      the boundaries d = k S - 1, k S, k S + 1 and the top of the class interval;
//   2. sf_unrank_lows is a bijection from [0, C(36,k)) onto the k-subsets of the 36 low cards for every
//      k with C(36,k) <= 2e6 (k <= 6), and injective with the right popcount on 2e6 ranks for larger k,
//      including the last rank C(36,k) - 1;
//   3. sf_honours_of is a bijection from [0, I) onto the honour holdings with the class's counts, for all
//      classes;
//   4. sum over the classes of I C(36, m) = C(52,13).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../bridge_mc_b200/csrc/seatfirst.cuh"

using namespace bmc;

static int fails = 0;
#define CHECK(cond, ...)                                   \
    do {                                                   \
        if (!(cond)) {                                     \
            if (fails < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } \
            ++fails;                                       \
        }                                                  \
    } while (0)

static const LowTabs T = make_lowtabs();

static uint64_t binom(int n, int k)
{
    if (k < 0 || k > n) return 0;
    unsigned __int128 r = 1;
    for (int i = 1; i <= k; ++i) r = r * (unsigned)(n - k + i) / (unsigned)i;
    return (uint64_t)r;
}

int main()
{
    std::mt19937_64 rng(12345);
    const uint64_t C4[5] = {1, 4, 6, 4, 1};

    // ---- 4. class weights -------------------------------------------------------------------
    unsigned __int128 total = 0;
    std::vector<uint32_t> Is;
    for (int id = 0; id < 625; ++id) {
        const int c[4] = {id % 5, (id / 5) % 5, (id / 25) % 5, id / 125}, n = c[0] + c[1] + c[2] + c[3];
        if (n > 13) continue;
        const uint64_t I = C4[c[0]] * C4[c[1]] * C4[c[2]] * C4[c[3]];
        total += (unsigned __int128)I * binom(36, 13 - n);
        Is.push_back((uint32_t)I);
    }
    CHECK(total == SF_HANDS, "class weights do not sum to C(52,13)");
    CHECK(SF_S == (uint64_t)((((unsigned __int128)1) << 64) / SF_HANDS), "SF_S");
    CHECK(SF_S_RECIP == (uint64_t)((((unsigned __int128)1) << 64) / SF_S), "SF_S_RECIP");
    std::sort(Is.begin(), Is.end());
    Is.erase(std::unique(Is.begin(), Is.end()), Is.end());

    // ---- 1. sf_split ------------------------------------------------------------------------
    for (uint32_t I : Is) {
        const uint32_t magic = I == 1 ? 0xFFFFFFFFu : (uint32_t)((1ull << 32) / I);
        const uint64_t Lmax = binom(36, 13);                       // q < I * L <= I * C(36,13)
        const unsigned __int128 dmax128 = (unsigned __int128)I * Lmax * SF_S;
        const uint64_t dmax = dmax128 > (unsigned __int128)~0ull ? ~0ull : (uint64_t)dmax128;
        auto check = [&](uint64_t d) {
            uint32_t i, ell;
            sf_split(d, I, magic, i, ell);
            const uint64_t q = d / SF_S;
            CHECK(i == q % I && ell == q / I, "sf_split d=%llu I=%u: got (%u,%u) want (%llu,%llu)", (unsigned long long)d, I, i, ell,
                  (unsigned long long)(q % I), (unsigned long long)(q / I));
        };
        for (int t = 0; t < 200000; ++t) check(rng() % dmax);
        for (int t = 0; t < 20000; ++t) {
            const uint64_t k = rng() % (dmax / SF_S);
            check(k * SF_S);
            if (k) check(k * SF_S - 1);
            check(k * SF_S + 1);
        }
        check(0); check(dmax - 1); check(SF_S - 1); check(SF_S);
    }

    // ---- 2. low cards -----------------------------------------------------------------------
    for (uint32_t k = 0; k <= 13; ++k) {
        const uint64_t L = binom(36, (int)k);
        const bool full = L <= 2000000ull;
        const uint64_t n = full ? L : 2000000ull;
        std::vector<uint64_t> seen;
        seen.reserve(n + 2);
        auto one = [&](uint32_t ell) {
            uint32_t m0, m1;
            sf_unrank_lows(T, ell, k, m0, m1);
            CHECK((m0 & ~0x01FF01FFu) == 0 && (m1 & ~0x01FF01FFu) == 0, "low cards outside the nine low ranks: k=%u ell=%u", k, ell);
            CHECK((uint32_t)(__builtin_popcount(m0) + __builtin_popcount(m1)) == k, "popcount: k=%u ell=%u", k, ell);
            seen.push_back(((uint64_t)m1 << 32) | m0);
        };
        if (full) for (uint64_t e = 0; e < L; ++e) one((uint32_t)e);
        else {
            // distinct random ranks + both ends
            std::vector<uint32_t> ranks;
            for (uint64_t t = 0; t < n; ++t) ranks.push_back((uint32_t)(rng() % L));
            ranks.push_back(0); ranks.push_back((uint32_t)(L - 1));
            std::sort(ranks.begin(), ranks.end());
            ranks.erase(std::unique(ranks.begin(), ranks.end()), ranks.end());
            for (uint32_t e : ranks) one(e);
        }
        const size_t cnt = seen.size();
        std::sort(seen.begin(), seen.end());
        seen.erase(std::unique(seen.begin(), seen.end()), seen.end());
        CHECK(seen.size() == cnt, "sf_unrank_lows is not injective for k=%u: %zu distinct of %zu", k, seen.size(), cnt);
        if (full) CHECK(cnt == L, "k=%u: %zu of %llu", k, cnt, (unsigned long long)L);
    }

    // ---- 2b. the tables behind it: prefix sums, radices, reciprocals, subset order --------------------
    for (int st = 0; st < 3; ++st)
        for (int k = 0; k <= 13; ++k) {
            const int R = 27 - 9 * st;
            uint64_t acc = 0;
            for (int l = 0; l <= 10; ++l) {
                CHECK(T.prefix[st][k][l] == (uint32_t)acc, "prefix[%d][%d][%d]", st, k, l);
                if (l <= 9) acc += binom(9, l) * binom(R, k - l);
            }
            CHECK(acc == binom(R + 9, k), "prefix total st=%d k=%d (Vandermonde)", st, k);
            for (int l = 11; l < 16; ++l) CHECK(T.prefix[st][k][l] == 0xFFFFFFFFu, "prefix padding");
        }
    for (int l = 0, pos = 0; l <= 9; ++l) {
        const uint32_t c = (uint32_t)binom(9, l);
        CHECK(T.c9[l][0] == c && T.c9[l][2] == (uint32_t)pos, "c9[%d]", l);
        CHECK(T.c9[l][1] == (c == 1 ? 0xFFFFFFFFu : (uint32_t)((1ull << 32) / c)), "c9 reciprocal %d", l);
        for (uint32_t r = 0; r < c; ++r) {
            CHECK(__builtin_popcount(T.subset9[pos + r]) == l && T.subset9[pos + r] < 512, "subset9 size");
            if (r) CHECK(T.subset9[pos + r] > T.subset9[pos + r - 1], "subset9 order");
        }
        pos += (int)c;
        // x / C(9,l) by the reciprocal, as sf_unrank_lows does it, for every x a stage can see (x < C(36,13))
        for (int t = 0; t < 300000; ++t) {
            const uint32_t x = (uint32_t)(rng() % 2310789600ull);
            uint32_t q = sf_mulhi32(x, T.c9[l][1]), r = x - q * c;
            if (r >= c) { r -= c; ++q; }
            CHECK(q == x / c && r == x % c, "division by C(9,%d) of %u", l, x);
        }
    }

    // ---- 3. honours -------------------------------------------------------------------------
    for (int id = 0; id < 625; ++id) {
        const int c[4] = {id % 5, (id / 5) % 5, (id / 25) % 5, id / 125};
        if (c[0] + c[1] + c[2] + c[3] > 13) continue;
        const uint32_t I = (uint32_t)(C4[c[0]] * C4[c[1]] * C4[c[2]] * C4[c[3]]);
        const uint32_t info = (uint32_t)c[0] | ((uint32_t)c[1] << 3) | ((uint32_t)c[2] << 6) | ((uint32_t)c[3] << 9) | (I << 12);
        std::vector<uint32_t> hs;
        for (uint32_t i = 0; i < I; ++i) {
            const uint32_t H = sf_honours_of(info, i);
            CHECK((H >> 16) == 0, "honour bits beyond 16");
            for (int lev = 0; lev < 4; ++lev)
                CHECK(__builtin_popcount(H & (0x1111u << lev)) == c[lev], "class %d level %d count", id, lev);
            hs.push_back(H);
        }
        std::sort(hs.begin(), hs.end());
        hs.erase(std::unique(hs.begin(), hs.end()), hs.end());
        CHECK(hs.size() == I, "sf_honours_of is not injective for class %d", id);
    }

    if (fails) { printf("unrank_selftest: %d failures\n", fails); return 1; }
    printf("unrank_selftest: ok\n");
    return 0;
}
