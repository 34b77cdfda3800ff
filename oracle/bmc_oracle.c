/*
 * bmc_oracle.c -- CPU restatement of bridge-mc's Monte-Carlo deal path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under bridge_mc_b200/ (the product) may
 * link, import or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker or the
 * timed CPU baseline.
 *
 * Parity status: PINNED.  The reference (Common Lisp, no SBCL in this image)
 * cannot be executed, so every function below is pinned against the
 * reference's own inline known-answer tests and its golden file
 * backend/test/distgen-test.dat (converted by tests/golden/make_golden.py into
 * the JSON fixtures under tests/golden; checked by the tests/test_oracle_ files).
 *
 * Each function cites the reference file:line it restates
 * (paths relative to /root/reference/backend/).
 *
 * Encoding (bridge.cl:519-547): suit c=0 d=1 h=2 s=3; rank 0..12 = 2..A;
 * card id = 13*suit + rank.  A hand here is four 13-bit rank masks.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

typedef struct { uint16_t s[4]; } orc_hand;

static inline int popc16(unsigned v) { return __builtin_popcount(v & 0x1FFFu); }

/* ------------------------------------------------------------------ */
/* Features                                                           */
/* ------------------------------------------------------------------ */

/* bridge.cl:667-668 rank-hcp: rank>8 -> rank-8 ; bridge.cl:2986-2987 suit-hcp */
ORC_API int orc_suit_hcp(unsigned ranks)
{
    int h = 0;
    for (int r = 9; r <= 12; ++r) if (ranks >> r & 1) h += r - 8;
    return h;
}

/* bidding.cl:20-23 suit-loosers */
ORC_API int orc_suit_losers(unsigned ranks)
{
    int len = popc16(ranks);
    return ((ranks >> 12 & 1) || len == 0 ? 0 : 1)
         + ((ranks >> 11 & 1) || len < 2 ? 0 : 1)
         + ((ranks >> 10 & 1) || len < 3 ? 0 : 1);
}

/* bidding.cl:40-46 balanced? (lens power loosers) */
ORC_API int orc_balanced(const int lens[4], const int power[4])
{
    for (int i = 0; i < 4; ++i) if (2 > lens[i]) return 0;          /* (curry #'> 2) */
    if (4 < lens[2] || 4 < lens[3]) return 0;                         /* hearts, spades */
    if (5 < lens[0] || 5 < lens[1]) return 0;                         /* clubs, diamonds */
    int hcp = power[0] + power[1] + power[2] + power[3];
    int weak = 0;
    for (int i = 0; i < 4; ++i) if (3 > power[i]) ++weak;
    return hcp < 12 || weak < 2;
}

/* bidding.cl:33-38 assess-hand -> (lens power losers); out[0..3] lens, [4..7]
 * power, [8] hcp, [9] losers, [10] balanced */
ORC_API void orc_assess(const uint16_t suits[4], int out[11])
{
    int losers = 0, hcp = 0;
    for (int i = 0; i < 4; ++i) {
        out[i] = popc16(suits[i]);
        out[4 + i] = orc_suit_hcp(suits[i]);
        hcp += out[4 + i];
        losers += orc_suit_losers(suits[i]);
    }
    out[8] = hcp;
    out[9] = losers;
    out[10] = orc_balanced(out, out + 4);
}

/* ------------------------------------------------------------------ */
/* bidding.cl rule tables, hard-coded line by line                    */
/* ------------------------------------------------------------------ */
#define C_  f[0]
#define D_  f[1]
#define H_  f[2]
#define S_  f[3]
#define CP  f[4]
#define DP  f[5]
#define HP  f[6]
#define SP  f[7]
#define HCP f[8]
#define LOS f[9]
#define BAL f[10]

/* bidding.cl:76-106 `openings`, index = table position:
 * 0 2C, 1 2NT, 2 1NT, 3 1S, 4 1H, 5 1D, 6 1C, 7 2D, 8 2H, 9 2S, 10 3C, 11 3D, 12 3H, 13 3S */
ORC_API int orc_opening_rule(int k, const int f[11])
{
    switch (k) {
    case 0: {   /* bidding.cl:77-80 */
        if (HCP > 22) return 1;
        if (BAL) return 0;
        if (4 < H_ || 4 < S_) return LOS <= 4;       /* cond clause 1: (subseq lens 2 4) */
        if (4 < C_ || 4 < D_) return LOS <= 3;       /* cond clause 2: (subseq lens 0 2) */
        return 0;
    }
    case 1:     /* bidding.cl:81-82 */
        return BAL && !(3 > CP || 3 > DP || 3 > HP || 3 > SP) && HCP > 18;
    case 2:     /* bidding.cl:83 */
        return BAL && HCP > 14 && HCP < 18;
    case 3:     /* bidding.cl:84-87 */
        return (LOS < 6 || HCP > 11) && LOS > 4 && HCP < 23 && S_ >= 5
            && !(HCP > 16 && (4 < C_ || 4 < D_ || 4 < H_));
    case 4:     /* bidding.cl:88-91 */
        return (LOS < 6 || HCP > 11) && LOS > 4 && HCP < 23 && H_ >= 5
            && !(HCP > 16 && (4 < C_ || 4 < D_));
    case 5:     /* bidding.cl:92-95 */
        return (LOS < 6 || HCP > 11) && LOS > 3 && HCP < 23 && D_ >= 4 && D_ >= C_
            && !(HCP > 16 && C_ >= 5);
    case 6:     /* bidding.cl:96-98 */
        return (LOS < 6 || HCP > 11) && LOS > 3 && HCP < 23 && C_ >= 3;
    case 7: case 8: case 9: {   /* bidding.cl:99-101  2D 2H 2S */
        int s = k - 6;
        return HCP > 5 && f[4 + s] >= 5 && (LOS > 5 || HCP < 11) && f[s] == 6;
    }
    case 10: case 11: case 12: case 13: {   /* bidding.cl:102-105  3C 3D 3H 3S */
        int s = k - 10;
        return HCP > 5 && f[4 + s] >= 5 && HCP < 11 && f[s] >= 7;
    }
    }
    return 0;
}
#define ORC_N_OPENINGS 14

/* bidding.cl:166-187 nt-responses.  level 1 table order:
 *   0 (2 C) 1 (2 D) 2 (2 H) 3 (2 S) 4 (3 C) 5 (2 NT) 6 (3 NT) 7 (4 NT) 8 (6 NT)
 * level 2: 0 (3 C) 1 (3 D) 2 (3 H) 3 (3 S) 4 (3 NT) 5 (4 NT) 6 (6 NT) */
ORC_API int orc_nt_response_count(int level) { return level == 1 ? 9 : 7; }

ORC_API int orc_nt_response_rule(int level, int k, const int f[11])
{
    int inv_lo = level == 1 ? 8 : 4;
    int game_lo = level == 1 ? 10 : 6, game_hi = level == 1 ? 15 : 10;
    int si_lo = level == 1 ? 16 : 12, si_hi = level == 1 ? 17 : 13;
    int slam = level == 1 ? 18 : 14;
    if (level != 1 && k >= 4) k += 2;      /* skip the level-1-only rows */
    switch (k) {
    case 0: return HCP >= inv_lo && (S_ == 4 || H_ == 4);
    case 1: return H_ >= 5;
    case 2: return S_ >= 5;
    case 3: return (HCP < inv_lo || HCP > game_hi - 2) && ((HCP > inv_lo && C_ >= 5) || C_ >= 6);
    case 4: return (HCP < 8 || HCP > 13) && ((HCP >= 7 && D_ >= 5) || D_ >= 6);
    case 5: return HCP >= 8 && HCP <= 9 && H_ < 5 && S_ < 5 && C_ < 6 && D_ < 6;
    case 6: return HCP >= game_lo && HCP <= game_hi;
    case 7: return HCP >= si_lo && HCP <= si_hi;
    case 8: return HCP >= slam;
    }
    return 0;
}

/* bidding.cl:204-210 2C-responses: 0 (2 D) 1 (2 H) 2 (2 S) 3 (2 NT) 4 (3 C) 5 (3 D) */
ORC_API int orc_2c_response_rule(int k, const int f[11])
{
    switch (k) {
    case 0: return HCP >= 6;
    case 1: return HCP <= 5 && H_ >= 4;
    case 2: return HCP <= 5 && S_ >= 4;
    case 3: return HCP <= 5 && BAL;
    case 4: return HCP <= 5 && C_ >= 5 && !BAL;
    case 5: return HCP <= 5 && D_ >= 5 && !BAL;
    }
    return 0;
}

/* bidding.cl:146-150 choose-bid: first matching row, -1 = nil.
 * scheme: 0 openings, 1 (nt-responses 1), 2 (nt-responses 2), 3 2C-responses */
ORC_API int orc_scheme_size(int scheme)
{
    return scheme == 0 ? ORC_N_OPENINGS : scheme == 1 ? 9 : scheme == 2 ? 7 : 6;
}
ORC_API int orc_test_rule(int scheme, int k, const int f[11])
{
    switch (scheme) {
    case 0: return orc_opening_rule(k, f);
    case 1: return orc_nt_response_rule(1, k, f);
    case 2: return orc_nt_response_rule(2, k, f);
    default: return orc_2c_response_rule(k, f);
    }
}
ORC_API int orc_choose_bid(int scheme, const uint16_t suits[4])
{
    int f[11];
    orc_assess(suits, f);
    int n = orc_scheme_size(scheme);
    for (int k = 0; k < n; ++k) if (orc_test_rule(scheme, k, f)) return k;
    return -1;
}
/* bidding.cl:111-112 test-bid */
ORC_API int orc_test_bid(int scheme, int k, const uint16_t suits[4])
{
    int f[11];
    orc_assess(suits, f);
    return orc_test_rule(scheme, k, f);
}

/* ------------------------------------------------------------------ */
/* Scoring                                                            */
/* ------------------------------------------------------------------ */

/* bridge.cl:3045-3069 contract-score.  suit: 0 c 1 d 2 h 3 s 4 NT/nil;
 * level 0 = not given (defaults to max(tricks-6,1)); dbl 0 nil, 1 x, 2 xx. */
ORC_API int orc_contract_score(int suit, int tricks, int vulnerable, int dbl, int level)
{
    if (!level) level = tricks - 6 > 1 ? tricks - 6 : 1;
    int result = tricks - level - 6;
    if (result >= 0) {
        int ppl = suit <= 1 ? 20 : 30;
        int base = ((suit == 4 ? 10 : 0) + ppl * level) * (dbl == 2 ? 4 : dbl ? 2 : 1);
        int bonus = level == 7 ? (vulnerable ? 2000 : 1300)
                  : level == 6 ? (vulnerable ? 1250 : 800)
                  : base >= 100 ? (vulnerable ? 500 : 300) : 50;
        bonus += dbl ? (dbl == 2 ? 100 : 50) + (vulnerable ? 200 : 100) * result : ppl * result;
        return base + bonus;
    }
    if (dbl) {
        int v = vulnerable ? -200 + 300 * (result + 1)
              : result == -1 ? -100 : result == -2 ? -300 : result == -3 ? -500
              : -800 + 300 * (result + 3);
        return v * (dbl == 2 ? 2 : 1);
    }
    return result * (vulnerable ? 100 : 50);
}

/* compscore.cl:94-100 base-score; declarer 0 N 1 E 2 S 3 W, -1 none */
ORC_API int orc_base_score(int level, int suit, int declarer, int dbl, int tricks, int vulnerable)
{
    int sc = orc_contract_score(suit, tricks, vulnerable, dbl, level);
    return (declarer == 1 || declarer == 3) ? -sc : sc;
}

/* compscore.cl:123-131; returns 0 on success, -1 where position-if gives nil
 * (|diff| >= 4000 -> Lisp type error) */
static const int IMP_PTS[24] = {20, 50, 90, 130, 170, 220, 270, 320, 370, 430, 500, 600, 750, 900,
                                1100, 1300, 1500, 1750, 2000, 2300, 2500, 3000, 3500, 4000};
ORC_API int orc_imps(int diff, int *out)
{
    int a = diff < 0 ? -diff : diff;
    for (int i = 0; i < 24; ++i)
        if (a < IMP_PTS[i]) { *out = diff < 0 ? -i : i; return 0; }
    return -1;
}

/* ------------------------------------------------------------------ */
/* RNG for the reference-algorithm sampler (the reference uses SBCL's */
/* never-seeded MT19937; any uniform source restates it)              */
/* ------------------------------------------------------------------ */
typedef struct { uint64_t s; } orc_rng;
static uint64_t rng_next(orc_rng *r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* (random n) for 0 < n < 2^63, exact by rejection */
static uint64_t rng_below(orc_rng *r, uint64_t n)
{
    uint64_t lim = UINT64_MAX - (UINT64_MAX % n + 1) % n;
    uint64_t x;
    do x = rng_next(r); while (x > lim);
    return x % n;
}

/* ------------------------------------------------------------------ */
/* distgen (bridge.cl:866-1083)                                       */
/* ------------------------------------------------------------------ */

/* range spec: lo/hi, -1 = nil (open).  "absent" suit spec = both -1. */
typedef struct {
    int hcp_lo, hcp_hi, has_hcp;      /* :hcp ; has_hcp=0 -> no filter */
    int len_lo[4], len_hi[4];          /* :c :d :h :s first two elements */
    int shcp_lo[4], shcp_hi[4], has_shcp[4]; /* third element (list form) */
} orc_spec;

typedef struct {
    int lc_supply[4];          /* bridge.cl:1030 */
    int hon[4][4];             /* available honours [suit][J Q K A] */
    uint16_t *pat;             /* admissible honour patterns, 16 bits: bit 15 = club J ... bit 0 = spade A */
    int npat;
    orc_spec spec;
    uint64_t total;            /* total-weight (fits: <= C(52,13) < 2^40) */
    uint64_t cards;            /* 52-bit mask of the cards dealt from */
} orc_distgen;

static uint64_t binom_tab[14][14];
static void binom_init(void)
{
    if (binom_tab[0][0]) return;
    for (int n = 0; n < 14; ++n)
        for (int k = 0; k < 14; ++k) {
            /* bridge.cl:814-816 combinations: prod(i = s-p+1 .. s) / p!  (0 when p > s) */
            if (k > n) { binom_tab[n][k] = 0; continue; }
            uint64_t num = 1, den = 1;
            for (int i = n - k + 1; i <= n; ++i) num *= (uint64_t)i;
            for (int i = 2; i <= k; ++i) den *= (uint64_t)i;
            binom_tab[n][k] = num / den;
        }
}
ORC_API uint64_t orc_combinations(int supply, int amount)
{
    binom_init();
    if (supply < 0 || supply > 13 || amount < 0) return 0;
    if (amount > 13) return 0;
    return binom_tab[supply][amount];
}

/* bridge.cl:876-878 supply: per suit (n_low nJ nQ nK nA) */
ORC_API void orc_supply(uint64_t cards, int out[4][5])
{
    for (int s = 0; s < 4; ++s) {
        unsigned lane = (unsigned)(cards >> (13 * s)) & 0x1FFFu;
        out[s][0] = __builtin_popcount(lane & 0x1FFu);
        for (int h = 0; h < 4; ++h) out[s][1 + h] = lane >> (9 + h) & 1;
    }
}

static int pat_bit(unsigned pat, int s, int h) { return pat >> (15 - (4 * s + h)) & 1; }

/* bridge.cl:921-924 permut-hcp on the flat 16 list: weight (i mod 4)+1 */
static int pat_hcp(unsigned pat)
{
    int t = 0;
    for (int i = 0; i < 16; ++i) if (pat >> (15 - i) & 1) t += i % 4 + 1;
    return t;
}
static int pat_suit_hcp(unsigned pat, int s)
{
    int t = 0;
    for (int h = 0; h < 4; ++h) if (pat_bit(pat, s, h)) t += h + 1;
    return t;
}

/* bridge.cl:880-888 valtest-fun on (lo hi) with nil ends */
static int in_range(int v, int lo, int hi)
{
    if (lo >= 0 && v < lo) return 0;
    if (hi >= 0 && v > hi) return 0;
    return 1;
}

/* bridge.cl:990-1004 possible-lc-lens + 828-831 permut-weight, summed as in
 * `next` (bridge.cl:1043-1044).  If lc_out != NULL the admissible tuples and
 * their weights are stored (dist-from-hc, bridge.cl:1053-1059). Enumerates the
 * full product and filters on the sum, like list-prod + filter. */
static uint64_t pattern_weight(const orc_distgen *g, unsigned pat,
                               int (*lc_out)[4], uint64_t *w_out, int *n_out)
{
    int base[4], lo[4], hi[4], hc = 0;
    for (int s = 0; s < 4; ++s) {
        base[s] = 0;
        for (int h = 0; h < 4; ++h) base[s] += pat_bit(pat, s, h);
        hc += base[s];
        int mn = g->spec.len_lo[s] < 0 ? 0 : g->spec.len_lo[s];
        int mx = g->spec.len_hi[s] < 0 ? g->lc_supply[s] : g->spec.len_hi[s];  /* (or max supply): the 9-cap quirk */
        lo[s] = mn - base[s];
        hi[s] = mx - base[s];
        if (lo[s] < 0) lo[s] = 0;                /* (filter (curry #'<= 0) ...) */
    }
    int need = 13 - hc, n = 0;
    uint64_t total = 0;
    for (int a = lo[0]; a <= hi[0]; ++a)
        for (int b = lo[1]; b <= hi[1]; ++b)
            for (int c = lo[2]; c <= hi[2]; ++c) {
                /* list-prod + filter on the sum (bridge.cl:1002-1004): per (a b c) at most one d qualifies */
                const int d = need - a - b - c;
                if (d < lo[3] || d > hi[3]) continue;
                {
                    uint64_t w = orc_combinations(g->lc_supply[0], a) * orc_combinations(g->lc_supply[1], b)
                               * orc_combinations(g->lc_supply[2], c) * orc_combinations(g->lc_supply[3], d);
                    if (lc_out) { lc_out[n][0] = a; lc_out[n][1] = b; lc_out[n][2] = c; lc_out[n][3] = d; w_out[n] = w; }
                    ++n;
                    total += w;
                }
            }
    if (n_out) *n_out = n;
    return total;
}

/* bridge.cl:926-963 hc-permuts: all subsets of the available honours in binary
 * counting order (first available position most significant), filtered by the
 * total-HCP test and the per-suit HCP tests.  No available honour -> nil. */
static void build_patterns(orc_distgen *g)
{
    int pos[16], npos = 0;
    for (int i = 0; i < 16; ++i) if (g->hon[i / 4][i % 4]) pos[npos++] = i;
    g->npat = 0;
    g->pat = NULL;
    if (!npos) return;
    g->pat = (uint16_t *)malloc(sizeof(uint16_t) * ((size_t)1 << npos));
    for (unsigned m = 0; m < (1u << npos); ++m) {
        unsigned pat = 0;
        for (int j = 0; j < npos; ++j)
            if (m >> (npos - 1 - j) & 1) pat |= 1u << (15 - pos[j]);
        if (g->spec.has_hcp && !in_range(pat_hcp(pat), g->spec.hcp_lo, g->spec.hcp_hi)) continue;
        int ok = 1;
        for (int s = 0; s < 4 && ok; ++s)
            if (g->spec.has_shcp[s]) {
                int h = pat_suit_hcp(pat, s);   /* (or lo 0) .. (or hi 10), bridge.cl:936-937 */
                int l = g->spec.shcp_lo[s] < 0 ? 0 : g->spec.shcp_lo[s];
                int u = g->spec.shcp_hi[s] < 0 ? 10 : g->spec.shcp_hi[s];
                ok = h >= l && h <= u;
            }
        if (ok) g->pat[g->npat++] = (uint16_t)pat;
    }
}

/* bridge.cl:1027-1036 (make-instance 'distgen :cards .. :hcp .. :c .. :d .. :h .. :s ..) */
ORC_API orc_distgen *orc_distgen_new(uint64_t cards, const orc_spec *spec)
{
    binom_init();
    orc_distgen *g = (orc_distgen *)calloc(1, sizeof *g);
    int sup[4][5];
    orc_supply(cards, sup);
    for (int s = 0; s < 4; ++s) {
        g->lc_supply[s] = sup[s][0];
        for (int h = 0; h < 4; ++h) g->hon[s][h] = sup[s][1 + h];
    }
    g->spec = *spec;
    g->cards = cards;
    build_patterns(g);
    g->total = 0;
    for (int i = 0; i < g->npat; ++i) g->total += pattern_weight(g, g->pat[i], NULL, NULL, NULL);
    return g;
}
ORC_API void orc_distgen_free(orc_distgen *g) { if (g) { free(g->pat); free(g); } }
ORC_API int orc_distgen_npat(const orc_distgen *g) { return g->npat; }
ORC_API uint64_t orc_distgen_total(const orc_distgen *g) { return g->total; }
/* bridge.cl:1038-1045 next: i-th (weight pattern) */
ORC_API uint64_t orc_distgen_item(const orc_distgen *g, int i, unsigned *pat)
{
    *pat = g->pat[i];
    return pattern_weight(g, g->pat[i], NULL, NULL, NULL);
}

/* Exact distribution of one constrained hand, from the enumeration:
 * hcp_num[38]: number of admissible hands by HCP; len_num[4][14] by suit length. */
ORC_API void orc_distgen_pmf(const orc_distgen *g, uint64_t hcp_num[38], uint64_t len_num[4][14])
{
    static __thread int lc[20000][4];
    static __thread uint64_t w[20000];
    memset(hcp_num, 0, 38 * sizeof(uint64_t));
    memset(len_num, 0, 4 * 14 * sizeof(uint64_t));
    for (int i = 0; i < g->npat; ++i) {
        int n;
        unsigned pat = g->pat[i];
        pattern_weight(g, pat, lc, w, &n);
        int hcp = pat_hcp(pat);
        for (int k = 0; k < n; ++k) {
            hcp_num[hcp] += w[k];
            for (int s = 0; s < 4; ++s) {
                int b = 0;
                for (int h = 0; h < 4; ++h) b += pat_bit(pat, s, h);
                len_num[s][lc[k][s] + b] += w[k];
            }
        }
    }
}

/* bridge.cl:646-650 deal: `amount` uniform picks without replacement from a pile */
static void deal_from(orc_rng *r, int *pile, int *npile, int amount, int *out, int *nout)
{
    for (int k = 0; k < amount; ++k) {
        int idx = (int)rng_below(r, (uint64_t)*npile);          /* (random (length lst)) */
        out[(*nout)++] = pile[idx];
        memmove(pile + idx, pile + idx + 1, sizeof(int) * (size_t)(*npile - idx - 1));   /* take, bridge.cl:627-634 */
        --*npile;
    }
}

/* bridge.cl:289-296 randcar-weights: index of the element chosen with probability weight / total
 * (walk with `<`, subtracting); -1 when the total weight is 0 (the reference returns nil). */
ORC_API int orc_randcar_weights(const uint64_t *weights, int n, uint64_t *rng_state)
{
    uint64_t total = 0;
    for (int i = 0; i < n; ++i) total += weights[i];
    if (total == 0) return -1;
    orc_rng r = { *rng_state };
    uint64_t w = rng_below(&r, total);
    *rng_state = r.s;
    for (int i = 0; i < n; ++i) {
        if (w < weights[i]) return i;
        w -= weights[i];
    }
    return n - 1;
}

/* bridge.cl:1075-1083 rand-hand: re-walks the pattern list recomputing every
 * weight (the per-deal rescan), with the reference's `>` comparison; then
 * rand-dist (1061-1063) / dist-from-hc (1053-1059) / randcar-weights (289-296)
 * and distrib-hand (1065-1071).  Returns the hand as a 52-bit card mask. */
ORC_API uint64_t orc_rand_hand(orc_distgen *g, uint64_t *rng_state)
{
    static __thread int lc[20000][4];
    static __thread uint64_t w[20000];
    orc_rng r = { *rng_state };
    uint64_t x = rng_below(&r, g->total);
    int cur = 0;
    for (;; ++cur) {
        uint64_t wt = pattern_weight(g, g->pat[cur], NULL, NULL, NULL);
        if (!(x > wt)) break;
        x -= wt;
    }
    unsigned pat = g->pat[cur];
    int n;
    uint64_t tot = pattern_weight(g, pat, lc, w, &n);
    int pick = 0;
    if (tot > 0) {
        uint64_t y = rng_below(&r, tot);
        for (pick = 0; pick < n; ++pick) { if (y < w[pick]) break; y -= w[pick]; }
    }
    uint64_t hand = 0;
    for (int s = 0; s < 4; ++s) {
        unsigned lane = (unsigned)(g->cards >> (13 * s)) & 0x1FFFu;
        int pile[13], np = 0, got[13], ng = 0;
        for (int rk = 0; rk <= 8; ++rk) if (lane >> rk & 1) pile[np++] = rk;
        deal_from(&r, pile, &np, tot > 0 ? lc[pick][s] : 0, got, &ng);
        for (int k = 0; k < ng; ++k) hand |= 1ull << (13 * s + got[k]);
        for (int h = 0; h < 4; ++h)
            if (pat_bit(pat, s, h)) hand |= 1ull << (13 * s + 9 + h);
    }
    *rng_state = r.s;
    return hand;
}

/* bridge.cl:655-658 deal-all 13 on the remaining cards: out[] gets the hands
 * in the order they are dealt; returns their number */
ORC_API int orc_deal_all(uint64_t rest, uint64_t *rng_state, uint64_t *out)
{
    orc_rng r = { *rng_state };
    int pile[52], np = 0, nh = 0;
    for (int c = 0; c < 52; ++c) if (rest >> c & 1) pile[np++] = c;
    while (np > 0) {
        int got[13], ng = 0;
        deal_from(&r, pile, &np, np < 13 ? np : 13, got, &ng);
        uint64_t h = 0;
        for (int k = 0; k < ng; ++k) h |= 1ull << got[k];
        out[nh++] = h;
    }
    *rng_state = r.s;
    return nh;
}

/* bridge.cl:1292-1319 build for the two shapes bench/tests need: no const
 * hands, zero or one ranged seat (spec may be NULL = `(make-instance 'deal)`),
 * remaining seats by deal-all.  The root distgen is cached by the caller like
 * the reference's `rootgen` slot.  out[0] = the sampled seat, out[1..3] rest. */
ORC_API void orc_build_one(orc_distgen *root, uint64_t *rng_state, uint64_t out[4])
{
    uint64_t hand = orc_rand_hand(root, rng_state);
    out[0] = hand;
    orc_deal_all(root->cards & ~hand, rng_state, out + 1);
}

/* ------------------------------------------------------------------ */
/* CPU mirror of the GPU sampler (csrc/deal.cuh) for bit-exact checks */
/* ------------------------------------------------------------------ */
static inline void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
/* Philox4x32-10 (Salmon et al. 2011) */
ORC_API void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = { ctr[0], ctr[1], ctr[2], ctr[3] };
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof c);
}

/* digit groups: word g supplies the bounded draws n = GRP_HI[g] .. GRP_LO[g] */
static const int GRP_HI[12] = {52, 48, 44, 40, 36, 32, 28, 24, 20, 16, 11, 5};
static const int GRP_LO[12] = {49, 45, 41, 37, 33, 29, 25, 21, 17, 12, 6, 2};

/* One raw deal of the GPU sampler: returns 0 and the two owner bit-planes
 * (16-bit lane per suit, bit = rank), or 1 if the index is void (a biased
 * bounded draw was rejected).  Mirrors bmc::deal52() in csrc/deal.cuh. */
ORC_API int orc_gpu_deal(uint64_t seed, uint64_t index, uint64_t *lo_plane, uint64_t *hi_plane)
{
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t w[12];
    for (uint32_t b = 0; b < 3; ++b) {
        uint32_t ctr[4] = { (uint32_t)index, (uint32_t)(index >> 32), b, 0 };
        orc_philox4x32_10(ctr, key, w + 4 * b);
    }
    int digit[53];
    int voided = 0;
    for (int g = 0; g < 12; ++g) {
        uint32_t x = w[g];
        uint64_t N = 1;
        for (int n = GRP_HI[g]; n >= GRP_LO[g]; --n) {
            uint64_t p = (uint64_t)x * (uint32_t)n;
            digit[n] = (int)(p >> 32);
            x = (uint32_t)p;
            N *= (uint64_t)n;
        }
        uint32_t thresh = (uint32_t)((((uint64_t)1) << 32) % N);
        if (x < thresh) voided = 1;
    }
    digit[1] = 0;
    int rem[4] = {13, 13, 13, 13};
    uint64_t lo = 0, hi = 0;
    for (int p = 0; p < 52; ++p) {
        int n = 52 - p, u = digit[n], owner = 0, acc = 0;
        for (owner = 0; owner < 4; ++owner) { acc += rem[owner]; if (u < acc) break; }
        --rem[owner];
        int suit = p / 13, rank = p % 13, bit = 16 * suit + rank;
        lo |= (uint64_t)(owner & 1) << bit;
        hi |= (uint64_t)(owner >> 1) << bit;
    }
    *lo_plane = lo; *hi_plane = hi;
    return voided;
}

/* ------------------------------------------------------------------ */
/* Batch helpers (plain loops over the functions above)               */
/* ------------------------------------------------------------------ */
static void lanes64_to_suits(uint64_t m, uint16_t s[4])
{
    for (int i = 0; i < 4; ++i) s[i] = (uint16_t)((m >> (16 * i)) & 0x1FFFu);
}
ORC_API void orc_features_batch(const uint64_t *masks, uint64_t n, int32_t *out)
{
    for (uint64_t i = 0; i < n; ++i) {
        uint16_t s[4];
        int f[11];
        lanes64_to_suits(masks[i], s);
        orc_assess(s, f);
        for (int k = 0; k < 11; ++k) out[11 * i + k] = f[k];
    }
}
ORC_API void orc_choose_bid_batch(int scheme, const uint64_t *masks, uint64_t n, int32_t *out)
{
    for (uint64_t i = 0; i < n; ++i) {
        uint16_t s[4];
        lanes64_to_suits(masks[i], s);
        out[i] = orc_choose_bid(scheme, s);
    }
}
ORC_API void orc_gpu_deal_batch(uint64_t seed, uint64_t first, uint64_t n, uint64_t *rec, uint8_t *voided)
{
    for (uint64_t i = 0; i < n; ++i)
        voided[i] = (uint8_t)orc_gpu_deal(seed, first + i, &rec[2 * i], &rec[2 * i + 1]);
}
ORC_API void orc_build_batch(orc_distgen *root, uint64_t *rng_state, uint64_t n, uint64_t *out)
{
    for (uint64_t i = 0; i < n; ++i) orc_build_one(root, rng_state, out + 4 * i);
}

/* Reference-algorithm run for the bench's CPU arm: n x (`build` of a deal whose
 * single ranged seat uses `root`, bridge.cl:1292-1319) followed by the
 * `test-bid` filter of opening row `rule_k` on the sampled hand (the only way
 * the reference gets bidding-constrained deals, SURVEY finding 0.5).
 * Returns the number of deals that pass. */
ORC_API uint64_t orc_ref_run(orc_distgen *root, uint64_t *rng_state, uint64_t n, int rule_k, uint64_t *checksum)
{
    uint64_t pass = 0, sum = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t hands[4];
        orc_build_one(root, rng_state, hands);
        uint16_t s[4];
        for (int k = 0; k < 4; ++k) s[k] = (uint16_t)((hands[0] >> (13 * k)) & 0x1FFFu);
        if (rule_k < 0 || orc_test_bid(0, rule_k, s)) ++pass;
        sum += hands[0] ^ hands[1] ^ (hands[2] << 1) ^ (hands[3] << 2);
    }
    if (checksum) *checksum = sum;
    return pass;
}

/* CPU mirror of bmc::deal_fixed (csrc/deal.cuh): deals with fixed (`:=`) hands.
 * fixed[k] = lane-layout mask of seat k's fixed hand, 0 = seat is free. */
static int gpu_deal_fixed_stream(uint64_t seed, uint64_t index, const uint64_t fixed[4], uint32_t stream,
                                 uint64_t *lo_plane, uint64_t *hi_plane)
{
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint64_t lo = 0, hi = 0, used = 0;
    int rem[4];
    for (int k = 0; k < 4; ++k) {
        rem[k] = fixed[k] ? 0 : 13;
        used |= fixed[k];
        if (k & 1) lo |= fixed[k];
        if (k & 2) hi |= fixed[k];
    }
    uint64_t free_mask = 0x1FFF1FFF1FFF1FFFull & ~used;
    int n = __builtin_popcountll(free_mask), n_free = n, voided = 0;
    uint32_t w[4] = {0, 0, 0, 0}, x = 0;
    for (int v = 0; n >= 1; ++v, --n) {
        int pos = __builtin_ctzll(free_mask);
        free_mask &= free_mask - 1;
        int u = 0;
        if (n >= 2) {
            if ((v & 3) == 0) {
                if ((v & 15) == 0) {
                    uint32_t ctr[4] = { (uint32_t)index, (uint32_t)(index >> 32), (uint32_t)(v >> 4), stream };
                    orc_philox4x32_10(ctr, key, w);
                }
                x = w[(v >> 2) & 3];
            }
            uint64_t p = (uint64_t)x * (uint32_t)n;
            u = (int)(p >> 32);
            x = (uint32_t)p;
            if ((v & 3) == 3 || n == 2) {
                uint64_t N = 1;
                for (int d = 0; d < 4; ++d) { int m = n_free - 4 * (v >> 2) - d; if (m >= 2) N *= (uint64_t)m; }
                if (x < (uint32_t)((((uint64_t)1) << 32) % N)) voided = 1;
            }
        }
        int owner, acc = 0;
        for (owner = 0; owner < 4; ++owner) { acc += rem[owner]; if (u < acc) break; }
        --rem[owner];
        lo |= (uint64_t)(owner & 1) << pos;
        hi |= (uint64_t)(owner >> 1) << pos;
    }
    *lo_plane = lo; *hi_plane = hi;
    return voided;
}
ORC_API int orc_gpu_deal_fixed(uint64_t seed, uint64_t index, const uint64_t fixed[4],
                               uint64_t *lo_plane, uint64_t *hi_plane)
{
    return gpu_deal_fixed_stream(seed, index, fixed, 0, lo_plane, hi_plane);
}

/* CPU mirror of the seat-first sampler (csrc/seatfirst.cuh).
 * Stage A1: the honour holding of the primary seat is drawn BY RANK.  The 625 count classes
 * (nJ nQ nK nA), weight C(4,nJ) C(4,nQ) C(4,nK) C(4,nA) C(36, 13 - n), are laid out on
 * [0, C(52,13)) with the classes whose HCP lies in [hcp_lo, hcp_hi] first (ascending class id
 * nJ + 5 nQ + 25 nK + 125 nA within each group); u = (hi << 32) | lo, the words (index & 3) of the
 * Philox calls with counter index >> 2 on streams 4 (hi) and 5 (lo); rank = u / S with S = floor(2^64 /
 * C(52,13)); u >= S C(52,13) voids the index.  The identity of the honours comes from
 * (u - S cum) mod I, I = product of the C(4, n.), as mixed-radix digits J first.
 * Stage A2: with q = (u - S cum) / S, i = q mod I picks the honours and ell = q / I, uniform on
 * [0, C(36, 13 - n)), is the rank of the low cards (no further random numbers).
 * Returns the hand as a lane mask; *voided: bit 0 = void in A1 (bit 1 is no longer used);
 * *a1_pass = 1 when the class has an admissible HCP (the index survives stage A1). */
static uint64_t binom_small(int n, int k)
{
    if (k < 0 || k > n) return 0;
    uint64_t r = 1;
    for (int i = 1; i <= k; ++i) r = r * (uint64_t)(n - k + i) / (uint64_t)i;
    return r;
}
static const uint8_t SUBSETS4[5][6] = { {0}, {1, 2, 4, 8}, {3, 5, 6, 9, 10, 12}, {7, 11, 13, 14}, {15} };
ORC_API uint64_t orc_gpu_seatfirst_hand(uint64_t seed, uint64_t index, int hcp_lo, int hcp_hi, int *voided, int *a1_pass)
{
    static const uint64_t C4[5] = {1, 4, 6, 4, 1};
    const uint64_t N = 635013559600ull;
    const uint64_t S = (uint64_t)((((unsigned __int128)1) << 64) / N);
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t wh[4], wl[4];
    uint64_t call = index >> 2;
    uint32_t ch[4] = { (uint32_t)call, (uint32_t)(call >> 32), 0, 4 }, cl[4] = { (uint32_t)call, (uint32_t)(call >> 32), 0, 5 };
    orc_philox4x32_10(ch, key, wh);
    orc_philox4x32_10(cl, key, wl);
    const int j = (int)(index & 3);
    const uint64_t u = ((uint64_t)wh[j] << 32) | wl[j];
    int vd = 0;
    *a1_pass = 0;
    if (u >= N * S) { *voided = 1; return 0; }
    const uint64_t R = u / S;
    /* walk the classes in table order: admitted ones first */
    uint64_t cum = 0, hand = 0, ell = 0;
    int found = 0;
    for (int group = 0; group < 2 && !found; ++group)
        for (int id = 0; id < 625 && !found; ++id) {
            int c[4] = { id % 5, (id / 5) % 5, (id / 25) % 5, id / 125 };
            int n = c[0] + c[1] + c[2] + c[3];
            if (n > 13) continue;
            int hcp = c[0] + 2 * c[1] + 3 * c[2] + 4 * c[3];
            int admitted = hcp >= hcp_lo && hcp <= hcp_hi;
            if (admitted != (group == 0)) continue;
            uint64_t I = C4[c[0]] * C4[c[1]] * C4[c[2]] * C4[c[3]];
            uint64_t lows = 1;                                   /* C(36, 13 - n) */
            for (int t = 1; t <= 13 - n; ++t) lows = lows * (uint64_t)(36 - t + 1) / (uint64_t)t;
            uint64_t wgt = I * lows;
            if (R < cum + wgt) {
                uint64_t q = (u - cum * S) / S;                      /* uniform on [0, I C(36, 13 - n)) */
                uint64_t i = q % I;
                ell = q / I;
                for (int lev = 0; lev < 4; ++lev) {
                    uint64_t r = C4[c[lev]];
                    int m4 = SUBSETS4[c[lev]][i % r];
                    i /= r;
                    for (int su = 0; su < 4; ++su)
                        if ((m4 >> su) & 1) hand |= 1ull << (16 * su + 9 + lev);
                }
                *a1_pass = admitted;
                found = 1;
            }
            cum += wgt;
        }
    /* Stage A2: the low cards from ell, suit by suit (c d h s): the number of cards l of the suit is the
     * largest l with sum_{j<l} C(9,j) C(R,k-j) <= ell (R = low cards of the later suits, k = low cards
     * still to place), the cards are the (ell' mod C(9,l))-th l-subset of the nine in numeric order. */
    int k = 13 - __builtin_popcountll(hand);
    for (int su = 0; su < 4; ++su) {
        int R = 27 - 9 * su, l = 0;
        uint64_t rank;
        if (su < 3) {
            uint64_t before = 0, acc = 0;
            for (int j = 0; j <= 9; ++j) {
                if (acc <= ell) { l = j; before = acc; }
                acc += binom_small(9, j) * binom_small(R, k - j);
            }
            ell -= before;
            rank = ell % binom_small(9, l);
            ell /= binom_small(9, l);
        } else {
            l = k;
            rank = ell;
        }
        uint64_t seen = 0;
        for (unsigned v = 0; v < 512; ++v)
            if (__builtin_popcount(v) == l && seen++ == rank) { hand |= (uint64_t)v << (16 * su); break; }
        k -= l;
    }
    *voided = vd;
    return hand;
}
/* Full seat-first deal: primary seat P from stage A, the other 39 cards by the fixed-hand dealer on
 * Philox stream 3.  Return: bit 0 = void in A1, bit 1 = void in A2, bit 2 = void in B, bit 3 = the
 * index passes stage A1 (HCP window). */
ORC_API int orc_gpu_deal_seatfirst(uint64_t seed, uint64_t index, int primary, int hcp_lo, int hcp_hi,
                                   uint64_t *lo_plane, uint64_t *hi_plane)
{
    int va = 0, pass = 0;
    uint64_t fixed[4] = {0, 0, 0, 0};
    fixed[primary] = orc_gpu_seatfirst_hand(seed, index, hcp_lo, hcp_hi, &va, &pass);
    if (va & 1) { *lo_plane = 0; *hi_plane = 0; return 1; }
    int vb = gpu_deal_fixed_stream(seed, index, fixed, 3, lo_plane, hi_plane);
    return va | (vb << 2) | (pass << 3);
}
ORC_API void orc_gpu_deal_seatfirst_batch(uint64_t seed, uint64_t first, uint64_t n, int primary, int hcp_lo, int hcp_hi,
                                          uint64_t *rec, uint8_t *flags)
{
    for (uint64_t i = 0; i < n; ++i)
        flags[i] = (uint8_t)orc_gpu_deal_seatfirst(seed, first + i, primary, hcp_lo, hcp_hi, &rec[2 * i], &rec[2 * i + 1]);
}

/* Best-effort CPU comparator (SURVEY 8d "ii"; NOT the reference algorithm and not a mirror of the GPU
 * stream): what a careful scalar host implementation of the same job costs.  Per index: Philox4x32-10
 * words, a partial Fisher-Yates shuffle of 13 cards for South by multiply-shift draws with Lemire
 * rejection, popcount features, the test-bid (1 NT) test (balanced?, 15-17 HCP); accepted deals get the
 * remaining 38 draws (a full deal).  Returns the number accepted; *checksum folds the accepted deals. */
static inline uint32_t best_bounded(uint32_t *w, int *have, uint32_t ctr[4], const uint32_t key[2], uint32_t n)
{
    for (;;) {
        if (*have == 0) { orc_philox4x32_10(ctr, key, w); ++ctr[2]; *have = 4; }
        uint64_t p = (uint64_t)w[4 - (*have)--] * n;
        if ((uint32_t)p >= (uint32_t)(-n) % n) return (uint32_t)(p >> 32);
    }
}
ORC_API uint64_t orc_cpu_best_run(uint64_t seed, uint64_t first, uint64_t n, uint64_t *checksum)
{
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint64_t accepted = 0, sum = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t idx = first + i;
        uint32_t ctr[4] = { (uint32_t)idx, (uint32_t)(idx >> 32), 0, 7 }, w[4];
        int have = 0;
        uint8_t cards[52];
        for (int c = 0; c < 52; ++c) cards[c] = (uint8_t)c;
        uint64_t south = 0;
        for (int k = 0; k < 13; ++k) {
            const uint32_t j = (uint32_t)k + best_bounded(w, &have, ctr, key, 52u - (uint32_t)k);
            const uint8_t t = cards[k]; cards[k] = cards[j]; cards[j] = t;
            south |= 1ull << (16 * (cards[k] / 13) + cards[k] % 13);
        }
        /* features: lengths, suit power (J1 Q2 K3 A4), HCP, balanced? (bidding.cl:40-46) */
        int len[4], pw[4], hcp = 0, weak = 0, ok = 1;
        for (int su = 0; su < 4; ++su) {
            const uint32_t lane = (uint32_t)(south >> (16 * su)) & 0x1FFFu;
            len[su] = __builtin_popcount(lane);
            pw[su] = (int)((lane >> 9) & 1) + 2 * (int)((lane >> 10) & 1) + 3 * (int)((lane >> 11) & 1) + 4 * (int)((lane >> 12) & 1);
            hcp += pw[su];
            weak += pw[su] < 3;
            if (len[su] < 2 || len[su] > (su < 2 ? 5 : 4)) ok = 0;
        }
        if (!ok || hcp < 15 || hcp > 17 || weak >= 2) continue;
        for (int k = 13; k < 51; ++k) {
            const uint32_t j = (uint32_t)k + best_bounded(w, &have, ctr, key, 52u - (uint32_t)k);
            const uint8_t t = cards[k]; cards[k] = cards[j]; cards[j] = t;
        }
        uint64_t lo = 0, hi = 0;
        for (int k = 0; k < 52; ++k) {
            const int owner = k < 13 ? 2 : (k < 26 ? 0 : (k < 39 ? 1 : 3)), bit = 16 * (cards[k] / 13) + cards[k] % 13;
            lo |= (uint64_t)(owner & 1) << bit;
            hi |= (uint64_t)(owner >> 1) << bit;
        }
        sum += lo ^ (hi << 1);
        ++accepted;
    }
    if (checksum) *checksum = sum;
    return accepted;
}

ORC_API void orc_gpu_deal_fixed_batch(uint64_t seed, uint64_t first, uint64_t n, const uint64_t fixed[4],
                                      uint64_t *rec, uint8_t *voided)
{
    for (uint64_t i = 0; i < n; ++i)
        voided[i] = (uint8_t)orc_gpu_deal_fixed(seed, first + i, fixed, &rec[2 * i], &rec[2 * i + 1]);
}

/* ------------------------------------------------------------------ */
/* Stand-in trick source + compscore tallies (mirror of csrc/score.cuh) */
/* ------------------------------------------------------------------ */
/* Eight exactly uniform trick counts 0..13 from the Philox stream (c0, c1, c2 + block, c3): the mixed-radix
 * digits (multiply-shift by 14) of the first word whose final remainder passes Lemire's test
 * (>= 2^32 mod 14^8 = 1 343 389 184); at most 16 blocks. */
ORC_API void orc_tricks_of(const uint32_t key[2], uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t t[8])
{
    for (uint32_t b = 0;; ++b) {
        uint32_t ctr[4] = { c0, c1, c2 + b, c3 }, w[4];
        orc_philox4x32_10(ctr, key, w);
        for (int j = 0; j < 4; ++j) {
            uint32_t x = w[j], d[8];
            for (int k = 0; k < 8; ++k) { uint64_t p = (uint64_t)x * 14u; d[k] = (uint32_t)(p >> 32); x = (uint32_t)p; }
            if (x >= 1343389184u || (b == 15 && j == 3)) { memcpy(t, d, sizeof d); return; }
        }
    }
}

typedef struct { int8_t level, suit, declarer, dbl; } orc_contract;

/* hist: tricks[8][14] | imps[8][49] | dropped[8] (u64), ACCUMULATED.  base-score per compscore.cl:94-100,
 * match-points per compscore.cl:123-131 (|diff| >= 4000 -> dropped). */
static void score_tally_one(const orc_contract *c, const uint8_t *vul, int nc, const uint8_t *pairs, int np,
                            const uint32_t tr[8], uint64_t *hist)
{
    for (int k = 0; k < nc; ++k) hist[k * 14 + tr[k]]++;
    for (int p = 0; p < np; ++p) {
        int a = pairs[2 * p], b = pairs[2 * p + 1], im;
        int sa = orc_base_score(c[a].level, c[a].suit, c[a].declarer, c[a].dbl, (int)tr[a], vul[a]);
        int sb = orc_base_score(c[b].level, c[b].suit, c[b].declarer, c[b].dbl, (int)tr[b], vul[b]);
        if (orc_imps(sa - sb, &im) == 0) hist[8 * 14 + p * 49 + 24 + im]++;
        else hist[8 * 14 + 8 * 49 + p]++;
    }
}
/* deal-keyed form (BMC_TALLY_SCORE of the sampling kernels): counter = the record, key = seed ^ "tricks" */
ORC_API void orc_score_tally_records(uint64_t seed, const uint64_t *rec, uint64_t n, const orc_contract *c, const uint8_t *vul,
                                     int nc, const uint8_t *pairs, int np, uint64_t *hist)
{
    seed ^= 0x0000747269636B73ull;
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t tr[8];
        orc_tricks_of(key, (uint32_t)rec[2 * i], (uint32_t)(rec[2 * i] >> 32), (uint32_t)rec[2 * i + 1], (uint32_t)(rec[2 * i + 1] >> 32), tr);
        score_tally_one(c, vul, nc, pairs, np, tr, hist);
    }
}
/* index-keyed form (bmc_score_tally): counter = (index_lo, index_hi, block, 1), key = seed */
ORC_API void orc_score_tally_index(uint64_t seed, uint64_t first, uint64_t n, const orc_contract *c, const uint8_t *vul,
                                   int nc, const uint8_t *pairs, int np, uint64_t *hist)
{
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t tr[8];
        orc_tricks_of(key, (uint32_t)(first + i), (uint32_t)((first + i) >> 32), 0u, 1u, tr);
        score_tally_one(c, vul, nc, pairs, np, tr, hist);
    }
}

/* ------------------------------------------------------------------ */
/* infer-* (bridge.cl:1085-1269) and `build` with any number of ranged */
/* seats (bridge.cl:1292-1319), self-contained                         */
/* ------------------------------------------------------------------ */
/* A range def as `build` sees it after rest.cl's normdef (numbers are (n n) lists): -1 = nil. */
typedef struct {
    int has_hcp, hcp_lo, hcp_hi;                 /* :hcp (lo hi) */
    int has_suit[4], len_lo[4], len_hi[4];       /* :c :d :h :s (lo hi [third]) */
    int has_third[4], th_lo[4], th_hi[4];        /* third element: per-suit HCP range (lo hi) */
} orc_def;

#define NILV (-1)
/* bridge.cl:1101-1104 -? with nil = NILV on the right-hand side only (the accumulator is always a number here) */
static int minus_q(int a, int b) { return b == NILV ? a : a - b; }

/* hcp spec of one other def for the folds: present?, (first, second) */
typedef struct { int present, lo, hi; } hcp_ref;

/* bridge.cl:1126-1153 infer-hcp-range over `nsuit` supply rows.  base: present?, lo, hi.  Returns 0 and
 * *out_has / out_lo / out_hi (out_has = 0: the reference returns nil), or -1 where it signals bad-range. */
static int infer_hcp_range(int supply[][5], int nsuit, int base_has, int base_lo, int base_hi,
                           const hcp_ref *others, int n_others, int *out_has, int *out_lo, int *out_hi)
{
    int total = 0, cards = 0;
    for (int s = 0; s < nsuit; ++s) {
        for (int h = 1; h <= 4; ++h) total += h * supply[s][h];          /* permut-hcp of the flattened honour counts */
        for (int h = 0; h <= 4; ++h) cards += supply[s][h];
    }
    int min_hcp = 0;
    if (cards == 13 * (n_others + 1)) {                                    /* (= (hands-in-supply supply) (+ (length other-defs) 1)) */
        int acc = total;
        for (int i = 0; i < n_others; ++i) {
            if (!others[i].present) acc = 0;                               /* :default 0 */
            else if (acc != 0) { acc = minus_q(acc, others[i].hi); if (acc < 0) acc = 0; }
        }
        min_hcp = acc;
    }
    int max_hcp = total;
    for (int i = 0; i < n_others; ++i)
        if (others[i].present) max_hcp = minus_q(max_hcp, others[i].lo);   /* default: acc */
    const int up_limit = max_hcp < total ? max_hcp : total;
    if (!base_has) {
        *out_has = (up_limit < total || min_hcp > 0);
        *out_lo = min_hcp; *out_hi = up_limit;
        return 0;
    }
    int down = base_lo == NILV ? 0 : base_lo;
    if (min_hcp > down) down = min_hcp;
    int up = base_hi == NILV ? up_limit : (base_hi < up_limit ? base_hi : up_limit);
    if (down > up) return -1;
    *out_has = 1; *out_lo = down; *out_hi = up;
    return 0;
}

/* bridge.cl:1267-1269 infer-distgen-params = infer-hcp-range ++ infer-suit-ranges (1192-1234).
 * Returns 0 and the narrowed def, -1 on bad-range. */
ORC_API int orc_infer_distgen_params(int supply[4][5], const orc_def *base_in, const orc_def *others, int n_others, orc_def *out)
{
    orc_def nil_def;
    memset(&nil_def, 0, sizeof nil_def);
    const orc_def *base = base_in ? base_in : &nil_def;                    /* (car nil) = nil: no key at all */
    memset(out, 0, sizeof *out);
    out->hcp_lo = out->hcp_hi = NILV;
    for (int s = 0; s < 4; ++s) out->len_lo[s] = out->len_hi[s] = out->th_lo[s] = out->th_hi[s] = NILV;
    hcp_ref refs[4];
    for (int i = 0; i < n_others; ++i) { refs[i].present = others[i].has_hcp; refs[i].lo = others[i].hcp_lo; refs[i].hi = others[i].hcp_hi; }
    if (infer_hcp_range(supply, 4, base->has_hcp, base->hcp_lo, base->hcp_hi, refs, n_others,
                        &out->has_hcp, &out->hcp_lo, &out->hcp_hi)) return -1;
    int cards = 0, suitlen[4];
    for (int s = 0; s < 4; ++s) {
        suitlen[s] = 0;
        for (int h = 0; h <= 4; ++h) suitlen[s] += supply[s][h];
        cards += suitlen[s];
    }
    const int exhaust = cards == 13 * (n_others + 1);
    for (int s = 0; s < 4; ++s) {
        int min_len = 0;
        if (exhaust) {
            int acc = suitlen[s];
            for (int i = 0; i < n_others; ++i) {
                if (!others[i].has_suit[s]) acc = 0;                                       /* :default 0 */
                else if (acc != 0) acc = minus_q(acc, others[i].len_hi[s] != NILV ? others[i].len_hi[s] : others[i].len_lo[s]);
            }
            min_len = acc;
        }
        int max_len = suitlen[s];
        for (int i = 0; i < n_others; ++i)
            if (others[i].has_suit[s]) max_len = minus_q(max_len, others[i].len_lo[s]);
        const int up_limit = max_len < suitlen[s] ? max_len : suitlen[s];
        if (!base->has_suit[s]) {
            if (up_limit < suitlen[s]) { out->has_suit[s] = 1; out->len_lo[s] = min_len; out->len_hi[s] = up_limit; }
            continue;
        }
        int down = base->len_lo[s] != NILV ? (base->len_lo[s] > min_len ? base->len_lo[s] : min_len) : min_len;
        int up = base->len_hi[s] != NILV ? (base->len_hi[s] < up_limit ? base->len_hi[s] : up_limit) : up_limit;
        if (down > up_limit) return -1;
        out->has_suit[s] = 1; out->len_lo[s] = down; out->len_hi[s] = up;
        if (n_others > 0) {                                                                   /* suit-hcp is a non-empty list */
            for (int i = 0; i < n_others; ++i) {
                refs[i].present = others[i].has_suit[s] && others[i].has_third[s];
                refs[i].lo = others[i].th_lo[s]; refs[i].hi = others[i].th_hi[s];
            }
            int has = 0, lo = NILV, hi = NILV;
            if (infer_hcp_range(&supply[s], 1, base->has_third[s], base->th_lo[s], base->th_hi[s], refs, n_others, &has, &lo, &hi)) return -1;
            if (has) { out->has_third[s] = 1; out->th_lo[s] = lo; out->th_hi[s] = hi; }
        }
    }
    return 0;
}

static void def_to_spec(const orc_def *d, orc_spec *sp)
{
    memset(sp, 0, sizeof *sp);
    sp->has_hcp = d->has_hcp; sp->hcp_lo = d->hcp_lo; sp->hcp_hi = d->hcp_hi;
    for (int s = 0; s < 4; ++s) {
        sp->len_lo[s] = d->has_suit[s] ? d->len_lo[s] : NILV;
        sp->len_hi[s] = d->has_suit[s] ? d->len_hi[s] : NILV;
        sp->has_shcp[s] = d->has_suit[s] && d->has_third[s];
        sp->shcp_lo[s] = d->th_lo[s]; sp->shcp_hi[s] = d->th_hi[s];
    }
}

/* rand-hand on a generator whose weights are kept (same distribution as orc_rand_hand, which re-walks the
 * pattern list recomputing every weight -- that form is the TIMED baseline; this one serves the distribution checks) */
static uint64_t rand_hand_weights(orc_distgen *g, const uint64_t *wts, orc_rng *r)
{
    static __thread int lc[20000][4];
    static __thread uint64_t w[20000];
    uint64_t x = rng_below(r, g->total);
    int cur = 0;
    for (;; ++cur) { if (!(x > wts[cur])) break; x -= wts[cur]; }         /* bridge.cl:1075-1080, `>` kept */
    unsigned pat = g->pat[cur];
    int n;
    uint64_t tot = pattern_weight(g, pat, lc, w, &n);
    int pick = 0;
    if (tot > 0) { uint64_t y = rng_below(r, tot); for (pick = 0; pick < n; ++pick) { if (y < w[pick]) break; y -= w[pick]; } }
    uint64_t hand = 0;
    for (int s = 0; s < 4; ++s) {
        unsigned lane = (unsigned)(g->cards >> (13 * s)) & 0x1FFFu;
        int pile[13], np = 0, got[13], ng = 0;
        for (int rk = 0; rk <= 8; ++rk) if (lane >> rk & 1) pile[np++] = rk;
        deal_from(r, pile, &np, tot > 0 ? lc[pick][s] : 0, got, &ng);
        for (int k = 0; k < ng; ++k) hand |= 1ull << (13 * s + got[k]);
        for (int h = 0; h < 4; ++h) if (pat_bit(pat, s, h)) hand |= 1ull << (13 * s + 9 + h);
    }
    return hand;
}

/* One `build` (bridge.cl:1292-1319): consts = 52-bit card masks of the := hands, defs = the :~ defs in order.
 * hands[] gets consts, the sampled seats, then the deal-all hands; returns their number, or -1 where the
 * reference signals an error (bad-range, or a generator of total weight 0: `(random 0)`).
 * root / root_w: the cached root generator and its weights (NULL: built here). */
static int ref_build_one(const uint64_t *consts, int n_consts, const orc_def *defs, int n_defs, orc_rng *r,
                         orc_distgen *root, const uint64_t *root_w, uint64_t *hands)
{
    uint64_t rest = (1ull << 52) - 1;
    int nh = 0;
    for (int i = 0; i < n_consts; ++i) { rest &= ~consts[i]; hands[nh++] = consts[i]; }
    int first = 1, di = 0;
    /* the first def always goes through the root generator, even when there is none ((car nil) = nil) */
    while (first || (di < n_defs && __builtin_popcountll(rest) > 13)) {
        orc_distgen *g = root;
        uint64_t *wts = NULL;
        const uint64_t *use_w = root_w;
        if (!first || !root) {
            int sup[4][5];
            orc_def nar;
            orc_spec sp;
            orc_supply(rest, sup);
            const int n_others = n_defs - di - 1 > 0 ? n_defs - di - 1 : 0;
            if (orc_infer_distgen_params(sup, di < n_defs ? &defs[di] : NULL, di + 1 < n_defs ? &defs[di + 1] : NULL, n_others, &nar)) return -1;
            def_to_spec(&nar, &sp);
            g = orc_distgen_new(rest, &sp);
            wts = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(g->npat ? g->npat : 1));
            for (int i = 0; i < g->npat; ++i) wts[i] = pattern_weight(g, g->pat[i], NULL, NULL, NULL);
            use_w = wts;
        }
        if (g->total == 0) { if (g != root) { orc_distgen_free(g); free(wts); } return -1; }
        const uint64_t h = rand_hand_weights(g, use_w, r);
        if (g != root) { orc_distgen_free(g); free(wts); }
        hands[nh++] = h;
        rest &= ~h;
        first = 0;
        ++di;
    }
    if (rest) {
        uint64_t st = r->s;
        nh += orc_deal_all(rest, &st, hands + nh);
        r->s = st;
    }
    return nh;
}

#include <pthread.h>
typedef struct {
    const uint64_t *consts; int n_consts; const orc_def *defs; int n_defs;
    uint64_t seed, n; uint64_t *out; int8_t *count; orc_distgen *root; const uint64_t *root_w;
} ref_job;
static void *ref_worker(void *arg)
{
    ref_job *j = (ref_job *)arg;
    orc_rng r = { j->seed };
    for (uint64_t i = 0; i < j->n; ++i) {
        uint64_t h[8] = {0};
        const int k = ref_build_one(j->consts, j->n_consts, j->defs, j->n_defs, &r, j->root, j->root_w, h);
        j->count[i] = (int8_t)k;
        for (int q = 0; q < 4; ++q) j->out[4 * i + q] = (k > q) ? h[q] : 0;
    }
    return NULL;
}
/* n builds on `threads` host threads (independent generators seeded seed + 7919 t): out[4 i + k] = hand k of build i
 * (52-bit masks, `build`'s order), count[i] = hands returned or -1 = the reference signals an error. */
ORC_API int orc_ref_build_batch(const uint64_t *consts, int n_consts, const orc_def *defs, int n_defs, uint64_t n, uint64_t seed,
                                int threads, uint64_t *out, int8_t *count)
{
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    binom_init();
    /* the root generator is made once, like the reference's `rootgen` slot (bridge.cl:1297-1301) */
    uint64_t cards = (1ull << 52) - 1;
    for (int i = 0; i < n_consts; ++i) cards &= ~consts[i];
    int sup[4][5];
    orc_def nar;
    orc_spec sp;
    orc_supply(cards, sup);
    if (orc_infer_distgen_params(sup, n_defs ? &defs[0] : NULL, n_defs > 1 ? &defs[1] : NULL, n_defs > 1 ? n_defs - 1 : 0, &nar)) {
        for (uint64_t i = 0; i < n; ++i) count[i] = -1;
        return -1;
    }
    def_to_spec(&nar, &sp);
    orc_distgen *root = orc_distgen_new(cards, &sp);
    uint64_t *root_w = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(root->npat ? root->npat : 1));
    for (int i = 0; i < root->npat; ++i) root_w[i] = pattern_weight(root, root->pat[i], NULL, NULL, NULL);
    pthread_t th[64];
    ref_job jobs[64];
    uint64_t at = 0;
    for (int t = 0; t < threads; ++t) {
        const uint64_t share = n / (uint64_t)threads + ((uint64_t)t < n % (uint64_t)threads ? 1 : 0);
        jobs[t] = (ref_job){ consts, n_consts, defs, n_defs, seed + 7919ull * (uint64_t)t, share, out + 4 * at, count + at, root, root_w };
        at += share;
        pthread_create(&th[t], NULL, ref_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    orc_distgen_free(root);
    free(root_w);
    return 0;
}

/* ------------------------------------------------------------------ */
/* CPU mirror of the seat-first sampler's EXACT form (mode 4)          */
/* ------------------------------------------------------------------ */
/* The primary seat's hands are grouped into CLASSES -- the honours exactly (pattern p: bit 15 = club J, 14 = club Q
 * ... bit 0 = spade A, the order of hc-permuts, bridge.cl:926-957) and the number of low cards (2..10) per suit
 * (l_c l_d l_h l_s) -- like distgen's enumeration (bridge.cl:990-1045) over the full deck.  Classes run in ascending
 * (p, l_c, l_d, l_h); a class has prod C(9, l_s) hands.  The classes whose features satisfy the seat are the table;
 * the uniform u of an index (as in orc_gpu_seatfirst_hand) has rank R = u / S; the index passes stage A1 iff
 * R < W = hands in the table; then class = the one with cum <= R < cum + weight, r = R - cum, and suit by suit
 * (clubs first) t = r mod C(9, l), r = r / C(9, l): the low cards are the t-th l-subset of the nine in numeric order. */
typedef struct { uint64_t *cum; uint32_t *pat; uint8_t (*l)[4]; uint32_t n; uint64_t W; } orc_xtab;

/* pred: kind 0 = ranges only; 1 = ranges AND test-rule(scheme, k); 2 = ranges AND choose-bid(scheme) == k.
 * lo / hi: inclusive ranges over the 11 features (lens, power, hcp, losers, balanced). */
ORC_API orc_xtab *orc_xtab_build(int kind, int scheme, int k, const int lo[11], const int hi[11])
{
    orc_xtab *t = (orc_xtab *)calloc(1, sizeof *t);
    size_t capn = 1 << 16;
    t->cum = (uint64_t *)malloc(capn * sizeof(uint64_t));
    t->pat = (uint32_t *)malloc(capn * sizeof(uint32_t));
    t->l = (uint8_t (*)[4])malloc(capn * 4);
    for (uint32_t p = 0; p < 65536; ++p) {
        uint16_t hon[4];
        int b = 0, hcp = 0;
        for (int s = 0; s < 4; ++s) {
            hon[s] = 0;
            for (int h = 0; h < 4; ++h) if (p >> (15 - (4 * s + h)) & 1) { hon[s] |= (uint16_t)(1u << (9 + h)); ++b; hcp += h + 1; }
        }
        if (hcp < lo[8] || hcp > hi[8]) continue;                       /* every class of the pattern fails the HCP range */
        for (int l0 = 0; l0 <= 9; ++l0) for (int l1 = 0; l1 <= 9; ++l1) for (int l2 = 0; l2 <= 9; ++l2) {
            const int l3 = 13 - b - l0 - l1 - l2;
            if (l3 < 0 || l3 > 9) continue;
            const int lv[4] = { l0, l1, l2, l3 };
            uint16_t suits[4];
            int f[11], ok = 1;
            for (int s = 0; s < 4; ++s) suits[s] = (uint16_t)(hon[s] | ((1u << lv[s]) - 1u));
            orc_assess(suits, f);
            for (int i = 0; i < 11 && ok; ++i) ok = f[i] >= lo[i] && f[i] <= hi[i];
            if (ok && kind == 1) ok = orc_test_rule(scheme, k, f);
            if (ok && kind == 2) {
                int first = -1, n = orc_scheme_size(scheme);
                for (int j = 0; j < n && first < 0; ++j) if (orc_test_rule(scheme, j, f)) first = j;
                ok = first == k;
            }
            if (!ok) continue;
            if (t->n == capn) {
                capn *= 2;
                t->cum = (uint64_t *)realloc(t->cum, capn * sizeof(uint64_t));
                t->pat = (uint32_t *)realloc(t->pat, capn * sizeof(uint32_t));
                t->l = (uint8_t (*)[4])realloc(t->l, capn * 4);
            }
            t->cum[t->n] = t->W;
            t->pat[t->n] = p;
            for (int s = 0; s < 4; ++s) t->l[t->n][s] = (uint8_t)lv[s];
            t->W += binom_small(9, l0) * binom_small(9, l1) * binom_small(9, l2) * binom_small(9, l3);
            ++t->n;
        }
    }
    return t;
}
ORC_API void orc_xtab_free(orc_xtab *t) { if (t) { free(t->cum); free(t->pat); free(t->l); free(t); } }
ORC_API uint64_t orc_xtab_weight(const orc_xtab *t) { return t->W; }
ORC_API uint32_t orc_xtab_classes(const orc_xtab *t) { return t->n; }

/* bit 0 = void (u >= S C(52,13)), bit 2 = void in stage B, bit 3 = the index passes stage A1 (R < W) */
ORC_API int orc_gpu_deal_exact(const orc_xtab *t, uint64_t seed, uint64_t index, int primary, uint64_t *lo_plane, uint64_t *hi_plane)
{
    const uint64_t N = 635013559600ull;
    const uint64_t S = (uint64_t)((((unsigned __int128)1) << 64) / N);
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t wh[4], wl[4];
    uint64_t call = index >> 2;
    uint32_t ch[4] = { (uint32_t)call, (uint32_t)(call >> 32), 0, 4 }, cl[4] = { (uint32_t)call, (uint32_t)(call >> 32), 0, 5 };
    orc_philox4x32_10(ch, key, wh);
    orc_philox4x32_10(cl, key, wl);
    const int j = (int)(index & 3);
    const uint64_t u = ((uint64_t)wh[j] << 32) | wl[j];
    *lo_plane = 0; *hi_plane = 0;
    if (u >= N * S) return 1;
    const uint64_t R = u / S;
    if (R >= t->W) return 0;
    uint32_t a = 0, b = t->n - 1;
    while (a < b) { uint32_t mid = (a + b + 1) / 2; if (t->cum[mid] <= R) a = mid; else b = mid - 1; }
    uint64_t r = R - t->cum[a], hand = 0;
    for (int s = 0; s < 4; ++s) {
        const int l = t->l[a][s];
        const uint64_t c = binom_small(9, l), tt = r % c;
        r /= c;
        uint64_t seen = 0;
        for (unsigned v = 0; v < 512; ++v)
            if (__builtin_popcount(v) == l && seen++ == tt) { hand |= (uint64_t)v << (16 * s); break; }
        for (int h = 0; h < 4; ++h) if (t->pat[a] >> (15 - (4 * s + h)) & 1) hand |= 1ull << (16 * s + 9 + h);
    }
    uint64_t fixed[4] = {0, 0, 0, 0};
    fixed[primary] = hand;
    const int vb = gpu_deal_fixed_stream(seed, index, fixed, 3, lo_plane, hi_plane);
    return (vb << 2) | 8;
}
ORC_API void orc_gpu_deal_exact_batch(const orc_xtab *t, uint64_t seed, uint64_t first, uint64_t n, int primary, uint64_t *rec, uint8_t *flags)
{
    for (uint64_t i = 0; i < n; ++i)
        flags[i] = (uint8_t)orc_gpu_deal_exact(t, seed, first + i, primary, &rec[2 * i], &rec[2 * i + 1]);
}
