// TEST ONLY -- compiles bridge_mc_b200/csrc/playdeal.cuh for the host so the CPU suite can check the device code of the
// trick estimator against the oracle (oracle/playdeal.py) without a GPU.  Not part of libbridgemc.so: the product has
// no CPU path.  stdin: one deal per line: trump (-1..3) then 16 suit masks (declarer c d h s, left, dummy, right);
// stdout: status score0 score1 n then n*4 cards (hex) and the winner numbers.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../bridge_mc_b200/csrc/playdeal.cuh"

int main(int argc, char **argv)
{
    const uint32_t arena_bytes = argc > 1 ? (uint32_t)atol(argv[1]) : (64u << 20);
    std::vector<uint8_t> arena(arena_bytes);
    static pd::Scratch S;
    int trump;
    while (scanf("%d", &trump) == 1) {
        pd::Hand h[4];
        for (int i = 0; i < 4; ++i) {
            h[i].s.m = 0;
            for (int s = 0; s < 4; ++s) { unsigned m; if (scanf("%u", &m) != 1) return 1; h[i].set(s, m); }
            h[i].set_id(i);
        }
        pd::Result r;
        if (argc > 2 && std::string(argv[2]) == "probe") {           // the scheduling probe: status score features...
            int feat[pd::ORDER_FEATURES] = {0};
            pd::play_deal(arena.data(), arena_bytes, trump, h, S, r, true, feat);
            printf("%d %u", r.status, r.nodes);
            for (int i = 0; i < pd::ORDER_FEATURES; ++i) printf(" %d", feat[i]);
            printf("\n");
            fflush(stdout);
            continue;
        }
        pd::play_deal(arena.data(), arena_bytes, trump, h, S, r);
        printf("%d %d %d %d %u %u", r.status, r.score[0], r.score[1], r.n, r.arena_peak, r.nodes);
        if (argc > 2) fprintf(stderr, "tmp_peak %u of %zu\n", r.tmp_peak, sizeof(S.tmp));
        if (r.status == 0) {
            for (int t = 0; t < r.n; ++t) for (int i = 0; i < 4; ++i) printf(" %02x", r.tricks[t][i]);
            for (int t = 0; t < r.n; ++t) printf(" %d", r.wno[t]);
        }
        printf("\n");
        fflush(stdout);
    }
    return 0;
}
