/*
 * bridgemc.h -- C ABI of libbridgemc.so, the B200 (sm_100a) Monte-Carlo deal
 * engine that sits behind bridge-mc's Lisp entry points.
 *
 * Every entry point names the reference interface it replaces (paths relative
 * to the reference's backend/ directory).  The Lisp side binds these with CFFI
 * (lisp/bridgemc-cffi.lisp); bridge_mc_b200/ is the same
 * binding in Python (ctypes) used by the tests and the benchmark.
 *
 * Conventions
 *  - plain pointers and sizes only; the caller owns every buffer it passes and
 *    the library never keeps a caller pointer after the call returns;
 *  - handles are opaque and freed by the matching *_destroy;
 *  - int return: 0 = BMC_OK, < 0 = error code; bmc_last_error() gives the
 *    message for the calling thread; nothing exits, aborts or throws;
 *  - there is no CPU fallback: without a CUDA device every compute call
 *    returns BMC_E_CUDA;
 *  - re-entrant: a bmc_ctx owns one CUDA stream; use one ctx per calling
 *    thread (hunchentoot workers, dealgen<..>/sim{..} threads -- rest.cl:403,
 *    bridge.cl:3100,3115) or serialise calls on a shared one.  A
 *    bmc_constraints / bmc_scheme may be shared by any number of contexts, on
 *    one device or several: its device-side tables are uploaded once per
 *    device, under a lock.
 *
 * Encoding (bridge.cl:519-547): suit c=0 d=1 h=2 s=3, rank 0..12 = 2..A.
 * Seats 0..3 = N E S W.  A 64-bit "lane mask" holds one 16-bit lane per suit,
 * bit r of lane s = card (s, r); bits 13..15 of every lane are zero.
 */
#ifndef BRIDGEMC_H
#define BRIDGEMC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define BMC_OK            0
#define BMC_E_INVALID    -1   /* bad argument */
#define BMC_E_CUDA       -2   /* no device / CUDA runtime error */
#define BMC_E_BAD_RANGE  -3   /* contradictory ranges: the reference's `bad-range` condition (bridge.cl:1109-1123) */
#define BMC_E_PARSE      -4   /* rule form not in the bidding.cl predicate language */
#define BMC_E_NOMEM      -5
#define BMC_E_CAPACITY   -6   /* program / table too large */

#define BMC_LANE_MASK 0x1FFF1FFF1FFF1FFFull

/* One deal = 128 bits (a CUDA uint4): two owner bit-planes over the 52 cards in
 * lane-mask layout.  Owner seat of a card = 2*hi_bit + lo_bit.  Replaces the
 * list of four `hand` objects `build` returns (bridge.cl:1316-1318). */
typedef struct { uint64_t lo, hi; } bmc_record;

/* Per-seat range definition: the `:~` plist of class `deal`
 * (bridge.cl:1283-1286; README "(:hcp (15 17) :s (5 nil (3 nil)))").
 * -1 = nil (open end).  fixed != 0: the seat holds exactly `fixed_mask`
 * (a `:=` const hand, bridge.cl:1285,1294). */
typedef struct {
    int8_t   hcp[2];
    int8_t   len[4][2];        /* c d h s */
    int8_t   suit_hcp[4][2];   /* third element of a suit spec */
    int8_t   losers[2];        /* losing-trick count (bidding.cl:20-23), extension */
    uint8_t  fixed;
    uint8_t  reserved[3];
    uint64_t fixed_mask;
} bmc_seat_spec;

/* assess-hand (bidding.cl:33-38) + balanced? (40-46) + choose-bid (146-150) for one seat */
typedef struct {
    uint8_t len[4];
    uint8_t suit_hcp[4];       /* "power" */
    uint8_t hcp;
    uint8_t losers;
    uint8_t balanced;
    int8_t  bid;               /* row of the scheme that matched first, -1 = nil / no scheme */
} bmc_seat_features;
typedef struct { bmc_seat_features seat[4]; } bmc_features;

/* hist-base pre-counts (bridge.cl:1325-1330 `base` argument) produced on the
 * device.  All counters are over ACCEPTED deals except raw/voided. */
#define BMC_HCP_BINS   40      /* 0..37 used */
#define BMC_LEN_BINS   14
#define BMC_SHAPE_BINS 40      /* 39 sorted shapes, see bmc_shape_table */
#define BMC_SCORE_CONTRACTS 8  /* contracts / IMP pairs of the fused compscore tally (bmc_constraints_set_scoring) */
#define BMC_SCORE_PAIRS     8
typedef struct {
    uint64_t raw;              /* deal indices drawn */
    uint64_t voided;           /* indices discarded by the exact-uniformity (Lemire) check */
    uint64_t accepted;
    uint64_t stored;           /* records written to the output buffer (<= cap) */
    uint64_t hcp[4][BMC_HCP_BINS];
    uint64_t len[4][4][BMC_LEN_BINS];     /* [seat][suit][length] */
    uint64_t shape[4][BMC_SHAPE_BINS];    /* [seat][sorted-shape class] */
    /* BMC_TALLY_SCORE: compscore.cl tallies over the ACCEPTED deals (base-score, match-points; compscore.cl:94-131) */
    uint64_t tricks[BMC_SCORE_CONTRACTS][14];   /* [contract][tricks] */
    uint64_t imps[BMC_SCORE_PAIRS][49];         /* [pair][24 + match-points(a, b)] */
    uint64_t imp_dropped[BMC_SCORE_PAIRS];      /* |score difference| >= 4000: match-points signals an error there */
} bmc_tallies;
#define BMC_TALLY_WORDS (sizeof(bmc_tallies) / sizeof(uint64_t))

#define BMC_TALLY_HCP   1u
#define BMC_TALLY_LEN   2u
#define BMC_TALLY_SHAPE 4u
#define BMC_TALLY_ALL   7u
#define BMC_TALLY_SCORE 8u     /* score the accepted deals against the contracts attached to the constraint */

typedef struct {
    uint64_t raw, voided, accepted, stored;
    uint32_t waves, launches;
    float    kernel_ms;        /* CUDA-event time of the sampling kernels */
    float    total_ms;         /* wall time of the call */
} bmc_stats;


/* compscore.cl:38-65 class `contract` after parsing (the string parser stays on
 * the Lisp side).  suit 0 c 1 d 2 h 3 s 4 NT; declarer 0 N 1 E 2 S 3 W, -1 none;
 * dbl 0 nil 1 x 2 xx. */
typedef struct { int8_t level, suit, declarer, dbl; } bmc_contract;

typedef struct bmc_ctx bmc_ctx;
typedef struct bmc_constraints bmc_constraints;
typedef struct bmc_scheme bmc_scheme;

/* ---- context ------------------------------------------------------------ */
int  bmc_init(int device, bmc_ctx **out);
void bmc_shutdown(bmc_ctx *ctx);
const char *bmc_last_error(void);
const char *bmc_version(void);
/* pinned host memory for record buffers (fast device->host copies); optional */
void *bmc_host_alloc(size_t bytes);
void  bmc_host_free(void *p);
/* the 39 sorted shapes in tally order (each row a>=b>=c>=d, sum 13) */
void  bmc_shape_table(uint8_t out[39][4]);

/* ---- constraints -------------------------------------------------------- */
/* Replaces (make-instance 'deal := const-hands :~ defs) (bridge.cl:1283-1286)
 * plus a per-seat bidding.cl predicate.  specs may be NULL (unconstrained, the
 * `full-random` deal of training.cl:1); rules[k] may be NULL or a rule FORM in
 * the language good-opening? evaluates (bidding.cl:60-74): and/or/not/cond,
 * < <= > >= =, variables C D H S Cpower Dpower Hpower Spower hcp loosers
 * balanced, and (find-if (curry #'< n) (subseq lens a b)) /
 * (find-if (curry #'> n) power).  BMC_E_BAD_RANGE where `build` would signal
 * bad-range; BMC_E_PARSE for a form outside the language. */
int  bmc_constraints_create(const bmc_seat_spec specs[4], const char *const rules[4],
                            bmc_constraints **out);
void bmc_constraints_destroy(bmc_constraints *c);
/* How raw deals are drawn (both are exactly uniform; they consume the Philox stream
 * differently, so the deal behind a given index depends on the mode):
 *   1 full deal   all 52 cards, then test (best when most deals pass the primary seat)
 *   2 seat-first  the primary (most selective) seat's honour holding (drawn by rank, holdings with
 *                 an admissible HCP first: one compare per raw index), then its low cards, then
 *                 the other three hands -- each stage only for survivors of the previous one
 *   4 seat-first, exact: the device enumerates every CLASS of hands (honours exactly, low-card counts per
 *                 suit -- the reference's distgen enumeration, bridge.cl:921-1045) the primary seat's ranges
 *                 AND rule form admit; the admitted hands come first in the rank order, so the one compare
 *                 per raw index accepts exactly the indices whose primary hand satisfies the seat, and the
 *                 survivors are unranked without any further test (a 13-card hand has 4 582 640 classes at most).
 *   0 auto        (default) at the first sampling call the enumeration gives the primary seat's EXACT
 *                 acceptance: 1 when it is 0.30 or more, else 4.
 * bmc_constraints_mode returns the resolved mode (0 before the first call),
 * bmc_constraints_primary the seat tested first, bmc_constraints_primary_hcp the HCP window of
 * that seat (its range after the rule form's ranges were absorbed) that stage A1 of mode 2 orders
 * the honour holdings by -- together with the seed they determine the deal behind an index. */
int  bmc_constraints_set_mode(bmc_constraints *c, int mode);
/* Mode 3, "reference order": reproduce `build`'s OWN multi-seat semantics (bridge.cl:1292-1319)
 * instead of the uniform distribution over valid deals: after the fixed hands (which must occupy
 * the first seats) the next `n_defs` seats are sampled one after the other, each uniform over the
 * hands of the remaining cards admitted by infer-distgen-params (bridge.cl:1126-1269) -- including
 * the low-card cap on suits without an explicit upper bound, the never-sampled last seat and the
 * dropped suit-HCP range of a lone def; the remaining seats are dealt by deal-all.  Ranges only (no
 * rule forms).  Every index yields one deal (raw == accepted) unless the narrowed ranges are empty
 * for it -- where the reference signals bad-range -- which is counted in `voided`. */
int  bmc_constraints_set_reference_order(bmc_constraints *c, int n_defs);
/* Rule forms are interpreted from bytecode by default.  on = 1: at the next sampling call the forms
 * are translated to C++, spliced into the same kernel source and compiled for sm_100a with NVRTC
 * (about a second or two, once per constraint and dealing mode) -- the way the reference `eval`s its
 * rule forms.  Results are bit-identical; worthwhile for long or repeated jobs on rule-heavy
 * constraints.  BMC_E_CUDA if libnvrtc cannot be loaded.  bmc_constraints_jit_ms: compile time spent.
 * on = 2 (the default): interpret, and switch to the compiled kernel by itself once the constraint has spent more
 * than a second of kernel time on interpreted rules (about what the build costs). */
int  bmc_constraints_set_jit(bmc_constraints *c, int on);
float bmc_constraints_jit_ms(const bmc_constraints *c);
int  bmc_constraints_mode(const bmc_constraints *c);
int  bmc_constraints_primary(const bmc_constraints *c);
/* Attaches the contracts to score to a constraint: with BMC_TALLY_SCORE in tally_mask the sampling kernels tally, for
 * every ACCEPTED deal, tricks[c][t] and imps[p][24 + match-points(contracts[pairs[2p]], contracts[pairs[2p+1]])] in the
 * launch that draws the deal -- `msr-<contract>` measures + base-score + histogram of rest.cl:287-303 / compscore.cl in
 * one pass ("compscore tallies over constrained samples").  The trick column of this fused path is a stand-in for the
 * play-deal estimator (bmc_play_deal below, about a millisecond of GPU time per deal): eight exactly uniform values 0..13 that are a function of the deal
 * alone (Philox keyed by seed, counter = the deal's record), so tallies do not depend on the dealing mode, the index
 * behind a deal or the split over devices.  n_contracts <= 8, n_pairs <= 8; vulnerable[c] != 0 => :vulnerable t. */
int  bmc_constraints_set_scoring(bmc_constraints *c, const bmc_contract *contracts, const uint8_t *vulnerable,
                                 uint32_t n_contracts, const uint8_t *pairs, uint32_t n_pairs);
int  bmc_constraints_primary_hcp(const bmc_constraints *c, int32_t *lo, int32_t *hi);
/* After the first sampling call in mode 0 / 4: the exact number of 13-card hands (of C(52,13) = 635 013 559 600) that
 * satisfy the primary seat's ranges and rule form, and the number of classes they fall into. */
int  bmc_constraints_primary_acceptance(const bmc_constraints *c, uint64_t *hands, uint64_t *classes);

/* A rule table as printed by Lisp: "(((2 C) . FORM) ((2 NT) . FORM) ...)" --
 * `openings`, (nt-responses n), 2C-responses, (basic-responses bid) ...
 * (bidding.cl:76-106,166-187,204-210,226-253).  Used by bmc_filter for
 * choose-bid (first row whose form is true, bidding.cl:146-150). */
int  bmc_scheme_create(const char *table_sexpr, bmc_scheme **out);
int  bmc_scheme_size(const bmc_scheme *s);
/* How bmc_filter evaluates the table: 0 interpret its rows from bytecode; 1 translate the rows to C++ and compile them
 * into the filter kernel with NVRTC (about a second, once per table and device; bit-identical results); 2 (default)
 * interpret small jobs, compile for large (>= 2^21 deals) or repeated ones. */
int  bmc_scheme_set_jit(bmc_scheme *s, int on);
float bmc_scheme_jit_ms(const bmc_scheme *s);
void bmc_scheme_destroy(bmc_scheme *s);

/* ---- sampling: replaces `build` in a loop / `sim` (bridge.cl:1292-1319, 3144-3162) */
/* Draws raw deals with Philox4x32-10 (key = seed, counter = deal index), keeps
 * those satisfying `cons`, writes them as records and tallies them.
 *   n_accept_target == 0: exactly the indices [first_index, first_index+n_raw).
 *   n_accept_target  > 0: whole waves of `wave` indices (0 = default 2^26)
 *       starting at first_index until accepted >= target or n_raw indices
 *       (0 = no limit) are used; the accepted SET depends only on
 *       (seed, first_index, wave, target); stats->raw tells how many indices were
 *       used (the next job continues at first_index + raw).  The waves are enqueued
 *       back to back -- one that starts after the target was met cancels itself on the
 *       device -- so the host synchronises a few times per job, not once per wave.
 *       Watchdog of the unlimited form: the call returns (BMC_OK, fewer than target
 *       accepted) after 2^36 indices without a single accepted deal, or 2^44 in all.
 * records may be NULL (tally only); at most `cap` records are stored.
 * Host-buffer form (what the Lisp shim calls): */
int bmc_sample(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index,
               uint64_t n_raw, uint64_t n_accept_target, uint64_t wave, uint32_t tally_mask,
               bmc_record *records, uint64_t cap, bmc_tallies *tallies, bmc_stats *stats);
/* Device-buffer form: records_dev / tallies_dev are device pointers (tallies_dev
 * is ACCUMULATED into, not cleared: zero it first); used by the multi-GPU
 * driver, which all-reduces tallies_dev with NCCL. */
int bmc_sample_dev(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index,
                   uint64_t n_raw, uint64_t n_accept_target, uint64_t wave, uint32_t tally_mask,
                   void *records_dev, uint64_t cap, void *tallies_dev, bmc_stats *stats);

/* Index records: the deal behind an index is a pure function of (seed, index, constraint, dealing mode), so an
 * accepted deal can leave the device as the 32-bit OFFSET of its index from first_index -- 4 bytes instead of 16 -- and
 * be regenerated where and when the hands are needed.  Same arguments and semantics as bmc_sample / bmc_sample_dev
 * with offsets[] in place of records[]; a call spans at most 2^32 indices (n_raw = 0 means 2^32).  offsets are in
 * no particular order (like records). */
int bmc_sample_indices(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index,
                       uint64_t n_raw, uint64_t n_accept_target, uint64_t wave, uint32_t tally_mask,
                       uint32_t *offsets, uint64_t cap, bmc_tallies *tallies, bmc_stats *stats);
int bmc_sample_indices_dev(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index,
                           uint64_t n_raw, uint64_t n_accept_target, uint64_t wave, uint32_t tally_mask,
                           void *offsets_dev, uint64_t cap, void *tallies_dev, bmc_stats *stats);
/* records[i] = the deal behind index first_index + offsets[i]; (cons, seed, first_index) as given to the sampling call
 * (and the same dealing mode: bmc_constraints_mode).  Any offset may be expanded, accepted or not. */
int bmc_expand(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index,
               const uint32_t *offsets, uint64_t n, bmc_record *records);
int bmc_expand_dev(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index,
                   const void *offsets_dev, uint64_t n, void *records_dev);

/* Packed index records: the accepted deals' Philox indices as GAP BYTES, about one byte per deal (1.03 at an acceptance of
 * 3.8 %) instead of four -- for hosts whose ingest bandwidth is shared by several GPUs.  The stream is a sequence of
 * chunks, in no particular order: 8-byte header (little endian: u32 offset of the chunk's first index from first_index,
 * u32 payload length), then the payload: with cur = that offset - 1, a byte b < 255 accepts index cur += b + 1 and
 * b = 255 skips: cur += 255.  *n_bytes = bytes written (bytes needed, with BMC_E_CAPACITY, if cap_bytes was too small;
 * cap_bytes >= 1.1 x the expected number of accepted deals + 8 per 65 536 raw indices is safe).  Same sampling, same
 * accepted set and tallies as bmc_sample_indices; bmc_unpack_indices (host, a format decoder) returns the 32-bit
 * offsets, which bmc_expand turns into records.  At most 2^32 indices per call. */
int bmc_sample_packed(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index, uint64_t n_raw,
                      uint64_t n_accept_target, uint64_t wave, uint32_t tally_mask, uint8_t *bytes, uint64_t cap_bytes,
                      uint64_t *n_bytes, bmc_tallies *tallies, bmc_stats *stats);
int bmc_unpack_indices(const uint8_t *bytes, uint64_t n_bytes, uint32_t *offsets, uint64_t cap, uint64_t *n);

/* One call, several GPUs of the box: devices[0..n_devices) each get a contiguous share of
 * [first_index, first_index + n_raw), one host thread and one internal context per device; the only exchange is the
 * sum of the tally buffers (device 0 reads its peers' buffers over NVLink when peer access is available, else the
 * host adds them).  Tallies and the accepted set are IDENTICAL for every number of devices.  records (host, may be
 * NULL): the devices' accepted deals, device after device; at most cap / n_devices per device.  With
 * n_accept_target > 0 every device instead runs whole waves of its own range first_index + g * 2^40 ... until it
 * holds ceil(target / n_devices) deals (that result depends on n_devices).  stats->kernel_ms: the slowest device. */
int bmc_sample_multi(const int *devices, int n_devices, const bmc_constraints *cons, uint64_t seed,
                     uint64_t first_index, uint64_t n_raw, uint64_t n_accept_target, uint64_t wave,
                     uint32_t tally_mask, bmc_record *records, uint64_t cap, bmc_tallies *tallies, bmc_stats *stats);

/* ---- distgen: the reference's exact enumerator on the device (bridge.cl:921-1045) -----------------------------
 * (make-instance 'distgen :cards CARDS :hcp .. :c .. :d .. :h .. :s ..) followed by `next` until exhausted: the
 * (weight pattern) rows in the generator's order.  cards = lane mask of the deck dealt from (BMC_LANE_MASK = all 52);
 * spec: hcp / len / suit_hcp ranges as distgen reads them (-1 = nil; fixed, losers ignored; an open upper length
 * bound means the number of LOW cards left in the suit, bridge.cl:1000-1001).  patterns[i]: bit 15 = club J, 14 =
 * club Q ... bit 0 = spade A (the reference's 16-element 0/1 list read as a binary number); weights[i] = number of
 * hands with exactly those honours.  Every pattern that passes the honour-level tests is a row, also with weight 0.
 * total_weight = the generator's total-weight; hcp_pmf[38] / len_pmf[4][14] (may be NULL) = hands by HCP / by suit
 * length: the exact distribution of one constrained hand.  patterns / weights may be NULL (count only). */
int bmc_distgen(bmc_ctx *ctx, uint64_t cards, const bmc_seat_spec *spec, uint16_t *patterns, uint64_t *weights,
                uint32_t cap, uint32_t *n_patterns, uint64_t *total_weight, uint64_t *hcp_pmf, uint64_t *len_pmf);

/* ---- filter / features on given deals: assess-hand, balanced?, test-bid,
 * choose-bid (bidding.cl:33-46,111-112,146-150) over a deal set ---------------- */
/* accept[i] = 1 iff deal i satisfies cons (NULL cons: all 1); feats[i] gets the
 * four seats' features, .bid = choose-bid over `scheme` (NULL: -1).  accept and
 * feats may each be NULL. */
int bmc_filter(bmc_ctx *ctx, const bmc_constraints *cons, const bmc_scheme *scheme,
               const bmc_record *records, uint64_t n, uint8_t *accept, bmc_features *feats);
/* device-buffer form (e.g. the records bmc_sample_dev just wrote): no host copies, asynchronous on the ctx's stream */
int bmc_filter_dev(bmc_ctx *ctx, const bmc_constraints *cons, const bmc_scheme *scheme,
                   const void *records_dev, uint64_t n, void *accept_dev, void *feats_dev);

/* ---- the auction driver: class `bidding` / `next` (bidding.cl:457-480) over a deal set --------------------------
 * For every deal: the first seat, from `dealer` on (N E S W order), that has a bid over `openings` (choose-bid,
 * first matching row) opens; its partner -- the deal "rolled" so that the opener sits first -- answers over the
 * scheme that opening selects: responses[row] for opening row `row` (what `next` installs as bid-scheme: 2c-responses
 * after (2 c), (nt-responses level) after NT, (basic-responses bid) after a one-level suit; NULL = none).
 * opener -1 = passed out; opening / response = row of the scheme (-1 = nil); status 1 / 2 = evaluation reached an
 * illegal form in the openings / the responses (the reference signals an error there). */
typedef struct { int8_t opener, opening, response, status; } bmc_auction_row;
int bmc_auction(bmc_ctx *ctx, const bmc_scheme *openings, const bmc_scheme *const *responses, int dealer,
                const bmc_record *records, uint64_t n, bmc_auction_row *out);
int bmc_auction_dev(bmc_ctx *ctx, const bmc_scheme *openings, const bmc_scheme *const *responses, int dealer,
                    const void *records_dev, uint64_t n, void *out_dev);

/* ---- scoring: contract-score (bridge.cl:3045-3069), base-score
 * (compscore.cl:94-100), match-points (compscore.cl:127-131) ------------------ */
/* scores[i*n_contracts + c] = base-score of contracts[c] making tricks[i*n_contracts + c]
 * tricks (0..13), vulnerable[c] != 0 => :vulnerable t. */
int bmc_score(bmc_ctx *ctx, const bmc_contract *contracts, const uint8_t *vulnerable,
              uint32_t n_contracts, const uint8_t *tricks, uint64_t n, int32_t *scores);
/* imps[i] = IMPs of score difference a[i]-b[i] (sign carried); status[i] = 1
 * where |diff| >= 4000 (the reference errors there), imps[i] = 0. */
int bmc_match_points(bmc_ctx *ctx, const int32_t *a, const int32_t *b, uint64_t n,
                     int8_t *imps, uint8_t *status);
/* Drill tally over an index range without dealing (the scoring arithmetic alone; the tally over CONSTRAINED samples
 * is BMC_TALLY_SCORE of the sampling calls): for index i in [first_index, first_index + n) and contract c, tricks =
 * the stand-in trick source keyed by (seed, i) -- eight exactly uniform values 0..13 (mixed-radix digits of one
 * Philox word, Lemire rejection); tricks_hist[c][t] += 1; for every pair p, imp_hist[p][24 + match-points(
 * contracts[pairs[2p]], contracts[pairs[2p+1]])] += 1, or imp_dropped[p] += 1 where |difference| >= 4000 (the
 * reference signals an error there).  score_table[c][t] = base-score (output).  imp_dropped and score_table may be
 * NULL.  n_contracts <= 8, n_pairs <= 8. */
int bmc_score_tally(bmc_ctx *ctx, const bmc_contract *contracts, const uint8_t *vulnerable,
                    uint32_t n_contracts, const uint8_t *pairs, uint32_t n_pairs,
                    uint64_t seed, uint64_t first_index, uint64_t n,
                    uint64_t *tricks_hist /*[n_contracts][14]*/, uint64_t *imp_hist /*[n_pairs][49]*/,
                    uint64_t *imp_dropped /*[n_pairs]*/, int32_t *score_table /*[n_contracts][14]*/);

/* ---- hist-base over measure tuples (bridge.cl:1325-1330; feeds hist-breakdown / hist-percent) ---- */
/* A measure yields one byte per deal; seat 0..3 = N E S W ("seat and its partner" for the pair measures).
 * kind: 0 HCP  1 suit length  2 suit HCP  3 losers  4 balanced?  5 line-hcp (bridge.cl:3238-3240)
 *       6 trump-length (bridge.cl:3083-3086)  7 sorted-shape class (suit-distribution, bridge.cl:1438-1439;
 *       see bmc_shape_table)  8 longest-suit of the pair, later suit wins ties (bridge.cl:3030-3036). */
typedef struct { uint8_t kind, seat, suit, reserved; } bmc_measure;
/* Evaluates the k (<= 8) measures on every deal, counts equal tuples and returns them in FIRST-SEEN order
 * of `records` -- hist-base's output for the rows (m1 .. mk): tuples[i*k + j], counts[i], i < *n_unique.
 * BMC_E_CAPACITY (with *n_unique set) if more than `cap` distinct tuples exist. */
int bmc_hist_base(bmc_ctx *ctx, const bmc_record *records, uint64_t n, const bmc_measure *measures, uint32_t k,
                  uint8_t *tuples, uint64_t *counts, uint64_t cap, uint64_t *n_unique);

/* ---- play-deal: the trick estimator (bridge.cl:2961-2963; best-tricks 3027-3028) ------------------ */
/* Result of `play-deal trump declarer left dummy right` for one deal.  `score` is the reference outcome's
 * (score ...) pair: score[0] = tricks of the side that leads first (the defence), score[1] = tricks of the
 * declaring side, i.e. what `best-tricks` returns (bridge.cl:3028).  cards[t][i] = i-th card played to trick t,
 * suit << 4 | rank (0xFF = nothing played), winner[t] = position in the trick of the winning card
 * (`winner-nos`).  status: 0 ok; 1 the deal's route lists outgrew the per-deal arena (only with an explicit
 * arena_kb, or beyond 8 MB); 2 the reference would signal a Lisp error on this deal; 3 an internal table overflowed;
 * 4 the reference returns nil. */
typedef struct {
    uint8_t status, n_tricks;
    uint8_t score[2];
    uint8_t cards[13][4];
    uint8_t winner[13];
    uint8_t reserved[3];
} bmc_play_result;   /* 72 bytes */
/* One warp plays one deal; batches of more than two rounds of the resident warps are played longest-expected first
 * (a probe pass scores every deal; environment BMC_PD_ORDER=0/1 forces it off/on; no result depends on it).  trump[i] in {-1 (no trump), 0..3 = c d h s}; declarer[i] = seat 0..3 (N E S W);
 * the left-hand opponent (seat + 1) leads.  per_deal = 0: trump[0] / declarer[0] apply to every deal.
 * arena_kb: per-deal scratch for the route lists; 0 = adaptive: 1 MB per resident deal in a first pass (environment
 * BMC_PD_ARENA_KB overrides), and the rare deals that outgrow it are played again with 8 MB each.  `reserved` is zero
 * unless BMC_PD_TIMING is set (diagnostic: SM cycles spent on the deal / 4096, 24 bits little endian).  `bmc_play_deal` takes host buffers,
 * `bmc_play_deal_dev` device pointers (no copies; runs on the context's stream and synchronises it). */
int bmc_play_deal(bmc_ctx *ctx, const bmc_record *records, uint64_t n, const int8_t *trump, const uint8_t *declarer,
                  int per_deal, uint32_t arena_kb, bmc_play_result *out);
int bmc_play_deal_dev(bmc_ctx *ctx, const void *d_records, uint64_t n, const int8_t *d_trump, const uint8_t *d_declarer,
                      int per_deal, uint32_t arena_kb, void *d_out);
/* The same for four explicit hands per deal -- hands[4*i + k] = lane mask of declarer, left, dummy, right -- which
 * need not cover the deck: the reference's own tests call play-deal's parts on endings of a few cards. */
int bmc_play_hands(bmc_ctx *ctx, const uint64_t *hands, uint64_t n, const int8_t *trump, int per_deal, uint32_t arena_kb,
                   bmc_play_result *out);
/* best-tricks (bridge.cl:3027-3028) for every deal: tricks[i] = score[1] of play-deal, 255 where status != 0 */
int bmc_best_tricks(bmc_ctx *ctx, const bmc_record *records, uint64_t n, const int8_t *trump, const uint8_t *declarer,
                    int per_deal, uint32_t arena_kb, uint8_t *tricks);

/* ---- measurement helpers -------------------------------------------------- */
/* INT32 issue-rate microbenchmark (IMAD + LOP3/IADD3 chains on every SM):
 * returns thread-level integer ops per second in *ops_per_s. */
int bmc_int_peak(bmc_ctx *ctx, double *ops_per_s, float *ms);
/* number of kernel launches this context has made (for bench.py's gpu_launches) */
uint64_t bmc_launch_count(const bmc_ctx *ctx);
/* use an external CUDA stream (e.g. torch's current stream) for subsequent calls; 0 restores the own stream */
int bmc_set_stream(bmc_ctx *ctx, void *cuda_stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* BRIDGEMC_H */
