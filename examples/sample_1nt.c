/* sample_1nt.c -- libbridgemc from plain C (C99): the calls a foreign-function binding makes.
 *
 *   gcc -std=c99 -Iinclude examples/sample_1nt.c -Lbridge_mc_b200 -lbridgemc -Wl,-rpath,$PWD/bridge_mc_b200 -o sample_1nt
 *   ./sample_1nt [n_accept]
 *
 * Draws deals in which South would open 1NT (the rule form of bidding.cl's `openings` row, given as
 * text exactly as Lisp prints it), prints the first deal and South's HCP split.  Without a CUDA device
 * bmc_init fails with BMC_E_CUDA and the program says so: there is no CPU fallback. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bridgemc.h"

static void print_deal(const bmc_record *r)
{
    static const char *seat_name[4] = {"N", "E", "S", "W"}, *suit_name[4] = {"C", "D", "H", "S"};
    static const char rank_name[] = "23456789TJQKA";
    for (int seat = 0; seat < 4; ++seat) {
        printf("%s:", seat_name[seat]);
        for (int suit = 3; suit >= 0; --suit) {
            printf(" %s ", suit_name[suit]);
            for (int rank = 12; rank >= 0; --rank) {
                const int bit = 16 * suit + rank;                       /* owner = 2 * hi bit + lo bit */
                const int owner = (int)(((r->hi >> bit) & 1u) << 1 | ((r->lo >> bit) & 1u));
                if (owner == seat) putchar(rank_name[rank]);
            }
        }
        putchar('\n');
    }
}

int main(int argc, char **argv)
{
    const uint64_t want = argc > 1 ? strtoull(argv[1], NULL, 10) : 100000;
    bmc_ctx *ctx = NULL;
    int rc = bmc_init(0, &ctx);
    if (rc != BMC_OK) {
        printf("bmc_init: error %d (%s) -- libbridgemc needs a CUDA device\n", rc, bmc_last_error());
        return rc == BMC_E_CUDA ? 3 : 1;
    }
    /* South: (and balanced (> hcp 14) (< hcp 18)) -- bidding.cl:83 */
    const char *rules[4] = {NULL, NULL, "(and balanced (> hcp 14) (< hcp 18))", NULL};
    bmc_constraints *cons = NULL;
    rc = bmc_constraints_create(NULL, rules, &cons);
    if (rc != BMC_OK) { printf("bmc_constraints_create: %s\n", bmc_last_error()); return 1; }

    bmc_record *records = (bmc_record *)bmc_host_alloc(want * sizeof(bmc_record));   /* pinned: fast copies */
    bmc_tallies *tallies = (bmc_tallies *)calloc(1, sizeof(bmc_tallies));
    bmc_stats stats;
    memset(&stats, 0, sizeof stats);
    /* accept-target form: waves of 2^26 indices from index 0 until `want` deals are accepted */
    rc = bmc_sample(ctx, cons, 0x6272696467656D63ull, 0, 0, want, 0, BMC_TALLY_ALL, records, want, tallies, &stats);
    if (rc != BMC_OK) { printf("bmc_sample: %s\n", bmc_last_error()); return 1; }

    int32_t lo = 0, hi = 0;
    bmc_constraints_primary_hcp(cons, &lo, &hi);
    printf("mode %d, primary seat %d, HCP window %d-%d\n", bmc_constraints_mode(cons), bmc_constraints_primary(cons), (int)lo, (int)hi);
    printf("raw %llu  accepted %llu  stored %llu  kernel %.3f ms  (%.3e raw deals/s)\n", (unsigned long long)stats.raw,
           (unsigned long long)stats.accepted, (unsigned long long)stats.stored, stats.kernel_ms,
           stats.kernel_ms > 0 ? (double)stats.raw / (stats.kernel_ms * 1e-3) : 0.0);
    printf("South HCP 15 / 16 / 17: %llu / %llu / %llu\n", (unsigned long long)tallies->hcp[2][15],
           (unsigned long long)tallies->hcp[2][16], (unsigned long long)tallies->hcp[2][17]);
    if (stats.stored) print_deal(&records[0]);

    bmc_host_free(records);
    free(tallies);
    bmc_constraints_destroy(cons);
    bmc_shutdown(ctx);
    return 0;
}
