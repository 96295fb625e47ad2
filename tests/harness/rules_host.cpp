// rules_host.cpp -- TEST HARNESS ONLY (never loaded by the product package).
// Compiles azb_rules.cuh -- the exact header the CUDA kernels include -- with g++ so the packed
// bit-plane rules can be compared with the oracle on the CPU (tests/test_rules_host.py) before a
// GPU is involved.  Works on unpacked records: import -> operation on the packed game -> export.
#include <cstdint>
#include <cstring>
#include "../../azul_deep_reinforcement_learning_b200/csrc/azb_rules.cuh"

using namespace azb;

struct HostSink {
    int64_t* c;
    void add(int i, uint32_t v) { c[i] += v; }
    void add_group(int i, uint32_t v) { c[i] += v; }
    void pass_done() {}
};

template <int P, int POOL>
static int run_op(int32_t* rec, int op, int a, const int8_t* draws, uint64_t seed, uint32_t gid, int first_rule,
                  uint32_t* mask6, int32_t* preview)
{
    Game<P> g;
    bool ok = import_record<P>(g, [&](int i) { return rec[i]; });
    if (!ok) return -16;
    Philox rng{(uint32_t)seed, (uint32_t)(seed >> 32)};
    int rc = 0;
    switch (op) {
    case 0: apply_move<P, POOL>(g, (uint32_t)a); break;
    case 1: {   // step with legality/ended checks (azul.py:296-302)
        uint32_t m[6];
        if (g.ended()) { rc = -2; break; }
        legal_mask(g, m);
        if (move_is_legal(g, (uint32_t)a) != action_is_legal(m, (uint32_t)a)) return -99;
        {   // the rollout's register forms agree with the reference forms at every replayed state:
            // open-rows words -> the same mask; move_if_legal -> the same verdict and the same move
            uint32_t open[P], mo[6];
            for (int p = 0; p < P; p++) open[p] = open_rows(g, p);
            legal_mask_open(g, open, mo);
            for (int w = 0; w < 6; w++) if (mo[w] != m[w]) return -98;
            Game<P> h1 = g, h2 = g;
            const bool legal = move_if_legal<P, POOL>(h1, (uint32_t)a);
            if (legal != action_is_legal(m, (uint32_t)a)) return -97;
            if (legal) {
                apply_move<P, POOL>(h2, (uint32_t)a);
                const uint32_t pp = (uint32_t)a / 30u;
                apply_move_core<P, POOL, true>(g, pp, (uint32_t)a - 30u * pp, open);     // on the game itself: undone below
                uint32_t w1[Game<P>::WORDS], w2[Game<P>::WORDS], w3[Game<P>::WORDS];
                h1.store(w1, 1, 0); h2.store(w2, 1, 0); g.store(w3, 1, 0);
                for (int w = 0; w < Game<P>::WORDS; w++) if (w1[w] != w2[w] || w1[w] != w3[w]) return -96;
                for (int p = 0; p < P; p++) if (open[p] != open_rows(g, p)) return -95;  // the move kept the open words current
                import_record<P>(g, [&](int i) { return rec[i]; });                      // back to the pre-move state
            }
        }
        if (!action_is_legal(m, (uint32_t)a)) { rc = -1; break; }
        if (draws) advance<P, POOL>(g, (uint32_t)a, [&](Game<P>& gg) { new_round_injected<P, POOL>(gg, [&](int k) { return (int)draws[k]; }); });
        else advance<P, POOL>(g, (uint32_t)a, [&](Game<P>& gg) { new_round_philox<P, POOL>(gg, rng, gid, PURPOSE_REFILL); });
        break;
    }
    case 2: next_player(g); break;
    case 3: count_score<P, POOL>(g); break;
    case 5: new_round_injected<P, POOL>(g, [&](int k) { return (int)draws[k]; }); break;
    case 6: reset_game<P, POOL>(g, rng, gid, first_rule); break;
    case 7: { Game<P> h = g; count_score<P, POOL>(h); for (int p = 0; p < P; p++) preview[p] = (int32_t)(h.scf[p] & 0xFFFFu); break; }
    case 8: rc = is_end_of_round(g); break;
    case 9: rc = is_end_of_game(g); break;
    case 10: {  // single-action legality test against the full mask, all action bytes
        uint32_t m[6];
        legal_mask(g, m);
        for (uint32_t b = 0; b < 256; b++) {
            rc += move_is_legal(g, b) != action_is_legal(m, b);
            Game<P> h = g;
            rc += move_if_legal<P, POOL>(h, b) != action_is_legal(m, b);
        }
        break;
    }
    case 100: break;   // round trip
    default: return -100;
    }
    if (mask6) legal_mask(g, mask6);
    if (rc >= 0 || op == 8 || op == 9) export_record<P>(g, [&](int i, int32_t v) { rec[i] = v; });
    return rc;
}

template <int P, int POOL>
static void run_rollout(int32_t* recs, int64_t n, int first_rule, uint64_t seed, uint32_t gid0, int k, int64_t* counters)
{
    const int U = 48 + 58 * P;
    Philox rng{(uint32_t)seed, (uint32_t)(seed >> 32)};
    HostSink sink{counters};
    for (int64_t i = 0; i < n; i++) {
        int32_t* rec = recs + i * U;
        Game<P> g;
        import_record<P>(g, [&](int j) { return rec[j]; });
        InlineWords words;
        rollout_steps<P, POOL>(g, rng, gid0 + (uint32_t)i, first_rule, k, sink, SingleLane{}, true, 1, words);
        export_record<P>(g, [&](int j, int32_t v) { rec[j] = v; });
    }
}

#define DISPATCH(P, POOL, CALL)                                                         \
    if (P == 2 && POOL == 0) { return CALL<2, 0>; } if (P == 2 && POOL == 1) { return CALL<2, 1>; } \
    if (P == 3 && POOL == 0) { return CALL<3, 0>; } if (P == 3 && POOL == 1) { return CALL<3, 1>; } \
    if (P == 4 && POOL == 0) { return CALL<4, 0>; } if (P == 4 && POOL == 1) { return CALL<4, 1>; }

extern "C" int hh_op(int32_t* rec, int players, int pool, int op, int a, const int8_t* draws, uint64_t seed,
                     uint32_t gid, int first_rule, uint32_t* mask6, int32_t* preview)
{
#define CALL_OP(P, POOL) if (players == P && pool == POOL) return run_op<P, POOL>(rec, op, a, draws, seed, gid, first_rule, mask6, preview);
    CALL_OP(2, 0) CALL_OP(2, 1) CALL_OP(3, 0) CALL_OP(3, 1) CALL_OP(4, 0) CALL_OP(4, 1)
    return -101;
}

extern "C" int hh_rollout(int32_t* recs, int64_t n, int players, int pool, int first_rule, uint64_t seed,
                          uint32_t gid0, int k, int64_t* counters)
{
#define CALL_RO(P, POOL) if (players == P && pool == POOL) { run_rollout<P, POOL>(recs, n, first_rule, seed, gid0, k, counters); return 0; }
    CALL_RO(2, 0) CALL_RO(2, 1) CALL_RO(3, 0) CALL_RO(3, 1) CALL_RO(4, 0) CALL_RO(4, 1)
    return -101;
}

template <int P, int POOL>
static int run_opponent(int32_t* rec, uint64_t seed, uint32_t gid, int require_two, uint32_t* mask6, int32_t* diff)
{
    Game<P> g;
    if (!import_record<P>(g, [&](int i) { return rec[i]; })) return -16;
    Philox rng{(uint32_t)seed, (uint32_t)(seed >> 32)};
    *diff = opponent_random<P, POOL>(g, rng, gid, require_two != 0, mask6);
    export_record<P>(g, [&](int i, int32_t v) { rec[i] = v; });
    return 0;
}

extern "C" int hh_opponent(int32_t* rec, int players, int pool, uint64_t seed, uint32_t gid, int require_two,
                           uint32_t* mask6, int32_t* diff)
{
#define CALL_OPP(P, POOL) if (players == P && pool == POOL) return run_opponent<P, POOL>(rec, seed, gid, require_two, mask6, diff);
    CALL_OPP(2, 0) CALL_OPP(2, 1) CALL_OPP(3, 0) CALL_OPP(3, 1) CALL_OPP(4, 0) CALL_OPP(4, 1)
    return -101;
}

extern "C" int hh_random_action(const uint32_t* mask6, uint32_t word) { return (int)random_action(mask6, word); }
