// azb_variant.cu -- kernels and C ABI of the opt-in rule variant "factory count by player count" (azb_variant.cuh):
// 5 / 7 / 9 factory displays for 2 / 3 / 4 players, 180 / 240 / 300 actions.  One game per thread, structure-of-arrays
// state, coalesced loads and stores; deliberately simple (no staging, no deferred passes): the variant is opt-in and sits
// outside the measured hot path.  A handle created with azb_create serves both paths; the azb_v_* entry points take the
// number of displays explicitly (5 = the reference's rule, used by the tests to pin this code to the default engine).
#include <cuda_runtime.h>
#include <stdint.h>

#include "azb_internal.h"
#include "azb_variant.cuh"

using namespace azb;

namespace var {

struct SmemSink {
    unsigned long long* c;
    __device__ __forceinline__ void add(int i, uint32_t v)
    {
        if (v) atomicAdd(&c[i], (unsigned long long)v);
    }
    __device__ __forceinline__ void add_group(int i, uint32_t v) { add(i, v); }
};

template <int P, int F, int POOL>
__global__ void k_v_reset(Launch L, const uint8_t* __restrict__ which)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= L.n) return;
    if (which && which[g] == 0) return;
    GameV<P, F> gm;
    gm.steps = L.state[(GameV<P, F>::DW + 4) * L.n + g];
    const Philox rng{L.k0, L.k1};
    reset_game_v<P, F, POOL>(gm, rng, L.gid0 + (uint32_t)g, L.first_rule);
    gm.store(L.state, L.n, g);
}

template <int P, int F>
__global__ void k_v_legal_mask(const uint32_t* __restrict__ s, int64_t n, uint64_t* __restrict__ mask6)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n) return;
    GameV<P, F> gm;
    gm.load(s, n, g);
    uint64_t m[6];
    legal_mask_v(gm, m);
#pragma unroll
    for (int p = 0; p < 6; p++) mask6[p * n + g] = m[p];
}

// Azul.step (azul.py:296-313): IllegalMove / GameEnded become status bits with the state untouched
template <int P, int F, int POOL>
__global__ void k_v_step(Launch L, const uint16_t* __restrict__ action, const int8_t* __restrict__ draws,
                         uint64_t* __restrict__ mask6_out, uint8_t* __restrict__ done_out, uint8_t* __restrict__ status_out)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= L.n) return;
    GameV<P, F> gm;
    gm.load(L.state, L.n, g);
    const uint32_t a = action[g];
    uint32_t status = 0;
    uint64_t m[6];
    if (a != 0xFFFFu) {
        if (gm.ended()) {
            status = ST_ENDED;
        } else {
            legal_mask_v(gm, m);
            if (!action_is_legal_v<F>(m, a)) {
                status = ST_ILLEGAL;
            } else {
                const Philox rng{L.k0, L.k1};
                advance_v<P, F, POOL>(gm, a, [&](GameV<P, F>& gg) {
                    if (draws) {
                        const int8_t* d = draws + 4 * F * g;
                        new_round_injected_v<P, F, POOL>(gg, [&](int k) { return (int)d[k]; });
                    } else {
                        new_round_philox_v<P, F, POOL>(gg, rng, L.gid0 + (uint32_t)g, PURPOSE_REFILL);
                    }
                });
                gm.store(L.state, L.n, g);
            }
        }
    }
    legal_mask_v(gm, m);
    if (!gm.ended() && m[0] == 0ull && gm.current_player() != 0u) status |= ST_STUCK;
    if (mask6_out) {
#pragma unroll
        for (int p = 0; p < 6; p++) mask6_out[p * L.n + g] = m[p];
    }
    if (done_out) done_out[g] = gm.ended() ? 1 : 0;
    if (status_out) status_out[g] = (uint8_t)(status | gm.status());
}

template <int P, int F, int POOL>
__global__ void k_v_rollout(Launch L, int k_steps, unsigned long long* __restrict__ counters)
{
    __shared__ unsigned long long cnt[AZB_N_COUNTERS];
    if (threadIdx.x < AZB_N_COUNTERS) cnt[threadIdx.x] = 0ull;
    __syncthreads();
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g < L.n) {
        GameV<P, F> gm;
        gm.load(L.state, L.n, g);
        const Philox rng{L.k0, L.k1};
        SmemSink sink{cnt};
        rollout_steps_v<P, F, POOL>(gm, rng, L.gid0 + (uint32_t)g, L.first_rule, k_steps, sink);
        gm.store(L.state, L.n, g);
    }
    __syncthreads();
    if (counters && threadIdx.x < AZB_N_COUNTERS && cnt[threadIdx.x]) atomicAdd(&counters[threadIdx.x], cnt[threadIdx.x]);
}

template <int P, int F>
__global__ void k_v_import(const int32_t* __restrict__ rec, uint32_t* __restrict__ s, int64_t n, uint8_t* __restrict__ ok_out)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n) return;
    const int32_t* r = rec + g * (48 + 58 * P + 5 * (F - 5));
    GameV<P, F> gm;
    const bool ok = import_record_v<P, F>(gm, [&](int i) { return r[i]; });
    gm.store(s, n, g);
    if (ok_out) ok_out[g] = ok ? 1 : 0;
}

template <int P, int F>
__global__ void k_v_export(const uint32_t* __restrict__ s, int32_t* __restrict__ rec, int64_t n)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n) return;
    int32_t* r = rec + g * (48 + 58 * P + 5 * (F - 5));
    GameV<P, F> gm;
    gm.load(s, n, g);
    export_record_v<P, F>(gm, [&](int i, int32_t v) { r[i] = v; });
}

static bool valid_factories(int players, int f) { return f == 5 || f == 2 * players + 1; }

}  // namespace var

// (players, factories) pairs that exist: the reference's five displays for every player count (the pin against the default
// engine) and the board game's 2P + 1 (7 for three players, 9 for four; for two players that is 5 again)
#define DISPATCH_V(h, f, EXPR)                                                         \
    {                                                                                  \
        const int pf_ = (h)->players * 16 + (f);                                       \
        if (pf_ == 2 * 16 + 5) { constexpr int P = 2, F = 5; EXPR; }                   \
        else if (pf_ == 3 * 16 + 5) { constexpr int P = 3, F = 5; EXPR; }              \
        else if (pf_ == 4 * 16 + 5) { constexpr int P = 4, F = 5; EXPR; }              \
        else if (pf_ == 3 * 16 + 7) { constexpr int P = 3, F = 7; EXPR; }              \
        else { constexpr int P = 4, F = 9; EXPR; }                                     \
    }
#define DISPATCH_VP(h, f, EXPR)                                                        \
    if ((h)->tile_pool == AZB_POOL_LID) { constexpr int POOL = 1; DISPATCH_V(h, f, EXPR) } \
    else { constexpr int POOL = 0; DISPATCH_V(h, f, EXPR) }

#define CHECK_V(h, f)                                                                                      \
    CHECK_HANDLE(h);                                                                                       \
    if (!var::valid_factories((h)->players, (f)))                                                          \
        return azb_fail(AZB_E_INVALID, "factories must be 5 (the reference's rule) or 2 * players + 1%s");

static inline dim3 v_grid(const azb_t* h) { return dim3((unsigned)((h->n_games + 127) / 128)); }

extern "C" {

int azb_v_state_words(int players, int factories)
{
    if (players < 2 || players > 4 || !var::valid_factories(players, factories)) return AZB_E_INVALID;
    return (factories + 1) / 2 + 5 + 5 * players;
}
int azb_v_record_size(int players, int factories)
{
    if (players < 2 || players > 4 || !var::valid_factories(players, factories)) return AZB_E_INVALID;
    return 48 + 58 * players + 5 * (factories - 5);
}
int azb_v_n_actions(int factories) { return 30 * (factories + 1); }

int azb_v_reset(azb_t* h, int factories, uint32_t* state, const uint8_t* which, void* stream)
{
    CHECK_V(h, factories);
    if (!state) return azb_fail(AZB_E_INVALID, "state is null%s");
    const Launch L = make_launch(h, state);
    DISPATCH_VP(h, factories, (var::k_v_reset<P, F, POOL><<<v_grid(h), 128, 0, (cudaStream_t)stream>>>(L, which)));
    CHECK_LAUNCH();
    return 0;
}

int azb_v_legal_mask(azb_t* h, int factories, const uint32_t* state, uint64_t* mask6, void* stream)
{
    CHECK_V(h, factories);
    if (!state || !mask6) return azb_fail(AZB_E_INVALID, "null buffer%s");
    DISPATCH_V(h, factories, (var::k_v_legal_mask<P, F><<<v_grid(h), 128, 0, (cudaStream_t)stream>>>(state, h->n_games, mask6)));
    CHECK_LAUNCH();
    return 0;
}

int azb_v_step(azb_t* h, int factories, uint32_t* state, const uint16_t* action, const int8_t* draws, uint64_t* mask6_out,
               uint8_t* done_out, uint8_t* status_out, void* stream)
{
    CHECK_V(h, factories);
    if (!state || !action) return azb_fail(AZB_E_INVALID, "null buffer%s");
    const Launch L = make_launch(h, state);
    DISPATCH_VP(h, factories, (var::k_v_step<P, F, POOL><<<v_grid(h), 128, 0, (cudaStream_t)stream>>>(L, action, draws, mask6_out, done_out, status_out)));
    CHECK_LAUNCH();
    return 0;
}

int azb_v_rollout_random(azb_t* h, int factories, uint32_t* state, int k_steps, unsigned long long* counters, void* stream)
{
    CHECK_V(h, factories);
    if (!state) return azb_fail(AZB_E_INVALID, "state is null%s");
    if (k_steps < 0) return azb_fail(AZB_E_INVALID, "k_steps < 0%s");
    const Launch L = make_launch(h, state);
    DISPATCH_VP(h, factories, (var::k_v_rollout<P, F, POOL><<<v_grid(h), 128, 0, (cudaStream_t)stream>>>(L, k_steps, counters)));
    CHECK_LAUNCH();
    return 0;
}

int azb_v_import_state(azb_t* h, int factories, const int32_t* records, uint32_t* state, uint8_t* ok_out, void* stream)
{
    CHECK_V(h, factories);
    if (!state || !records) return azb_fail(AZB_E_INVALID, "null buffer%s");
    DISPATCH_V(h, factories, (var::k_v_import<P, F><<<v_grid(h), 128, 0, (cudaStream_t)stream>>>(records, state, h->n_games, ok_out)));
    CHECK_LAUNCH();
    return 0;
}

int azb_v_export_state(azb_t* h, int factories, const uint32_t* state, int32_t* records, void* stream)
{
    CHECK_V(h, factories);
    if (!state || !records) return azb_fail(AZB_E_INVALID, "null buffer%s");
    DISPATCH_V(h, factories, (var::k_v_export<P, F><<<v_grid(h), 128, 0, (cudaStream_t)stream>>>(state, records, h->n_games)));
    CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
