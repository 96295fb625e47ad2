// variant_host.cpp -- TEST HARNESS ONLY (never loaded by the product package).
// Compiles azb_variant.cuh -- the header the variant's CUDA kernels include -- with g++ so the "factory count by player
// count" rules can be compared with the oracle (ao_set_factories) on the CPU.  Works on unpacked records.
#include <cstdint>
#include <cstring>
#include "../../azul_deep_reinforcement_learning_b200/csrc/azb_variant.cuh"

using namespace azb;

struct HostSink {
    int64_t* c;
    void add(int i, uint32_t v) { c[i] += v; }
    void add_group(int i, uint32_t v) { c[i] += v; }
};

// op 0: legal mask only; 1: step(action) with injected or Philox draws (rc -1 illegal / -2 ended); 6: reset; 100: round trip
template <int P, int F, int POOL>
static int run_op(int32_t* rec, int op, int a, const int8_t* draws, uint64_t seed, uint32_t gid, int first_rule, uint64_t* mask6)
{
    GameV<P, F> g;
    if (!import_record_v<P, F>(g, [&](int i) { return rec[i]; })) return -16;
    Philox rng{(uint32_t)seed, (uint32_t)(seed >> 32)};
    int rc = 0;
    if (op == 1) {
        uint64_t m[6];
        legal_mask_v(g, m);
        if (g.ended()) rc = -2;
        else if (!action_is_legal_v<F>(m, (uint32_t)a)) rc = -1;
        else if (draws) advance_v<P, F, POOL>(g, (uint32_t)a, [&](GameV<P, F>& gg) { new_round_injected_v<P, F, POOL>(gg, [&](int k) { return (int)draws[k]; }); });
        else advance_v<P, F, POOL>(g, (uint32_t)a, [&](GameV<P, F>& gg) { new_round_philox_v<P, F, POOL>(gg, rng, gid, PURPOSE_REFILL); });
    } else if (op == 6) {
        reset_game_v<P, F, POOL>(g, rng, gid, first_rule);
    }
    if (mask6) legal_mask_v(g, mask6);
    if (rc >= 0) export_record_v<P, F>(g, [&](int i, int32_t v) { rec[i] = v; });
    return rc;
}

template <int P, int F, int POOL>
static void run_rollout(int32_t* recs, int64_t n, int first_rule, uint64_t seed, uint32_t gid0, int k, int64_t* counters)
{
    const int U = 48 + 58 * P + 5 * (F - 5);
    Philox rng{(uint32_t)seed, (uint32_t)(seed >> 32)};
    HostSink sink{counters};
    for (int64_t i = 0; i < n; i++) {
        int32_t* rec = recs + i * U;
        GameV<P, F> g;
        import_record_v<P, F>(g, [&](int j) { return rec[j]; });
        rollout_steps_v<P, F, POOL>(g, rng, gid0 + (uint32_t)i, first_rule, k, sink);
        export_record_v<P, F>(g, [&](int j, int32_t v) { rec[j] = v; });
    }
}

#define FOR_ALL(X) X(2, 5, 0) X(2, 5, 1) X(3, 5, 0) X(3, 5, 1) X(4, 5, 0) X(4, 5, 1) X(3, 7, 0) X(3, 7, 1) X(4, 9, 0) X(4, 9, 1)

extern "C" int vh_op(int32_t* rec, int players, int factories, int pool, int op, int a, const int8_t* draws, uint64_t seed,
                     uint32_t gid, int first_rule, uint64_t* mask6)
{
#define CALL_OP(P, F, POOL) if (players == P && factories == F && pool == POOL) return run_op<P, F, POOL>(rec, op, a, draws, seed, gid, first_rule, mask6);
    FOR_ALL(CALL_OP)
    return -101;
}

extern "C" int vh_rollout(int32_t* recs, int64_t n, int players, int factories, int pool, int first_rule, uint64_t seed,
                          uint32_t gid0, int k, int64_t* counters)
{
#define CALL_RO(P, F, POOL) if (players == P && factories == F && pool == POOL) { run_rollout<P, F, POOL>(recs, n, first_rule, seed, gid0, k, counters); return 0; }
    FOR_ALL(CALL_RO)
    return -101;
}

extern "C" int vh_random_action(const uint64_t* mask6, uint32_t word, int factories)
{
    return factories == 5 ? (int)random_action_v<5>(mask6, word) : factories == 7 ? (int)random_action_v<7>(mask6, word) : (int)random_action_v<9>(mask6, word);
}
