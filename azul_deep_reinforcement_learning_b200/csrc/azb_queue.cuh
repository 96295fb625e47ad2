// azb_queue.cuh -- warp-private shared-memory queue of packed games (device only).
// Used where a rare, long piece of work (the end-of-round pass) hits a few lanes of every warp: the lanes park
// their game here and the warp finishes 32 parked games at a time with every lane busy.
#pragma once
#include "azb_rules.cuh"

using namespace azb;

constexpr int STEP_QCAP = 64;            // entries per warp queue: < 32 waiting + up to 32 new ones
constexpr int STEP_WARPS = 4;            // warps per block

template <int P>
__device__ __forceinline__ void queue_put(uint32_t* q, int slot, const Game<P>& g, uint32_t gidx, uint32_t status)
{
    q[0 * STEP_QCAP + slot] = g.pl0; q[1 * STEP_QCAP + slot] = g.pl1; q[2 * STEP_QCAP + slot] = g.pl2;
    q[3 * STEP_QCAP + slot] = g.misc; q[4 * STEP_QCAP + slot] = g.box; q[5 * STEP_QCAP + slot] = g.lid;
    q[6 * STEP_QCAP + slot] = g.steps;
#pragma unroll
    for (int p = 0; p < P; p++) {
        q[(7 + 5 * p) * STEP_QCAP + slot] = g.pat[p];  q[(8 + 5 * p) * STEP_QCAP + slot] = g.wall[p];
        q[(9 + 5 * p) * STEP_QCAP + slot] = g.scf[p];  q[(10 + 5 * p) * STEP_QCAP + slot] = g.sta[p];
        q[(11 + 5 * p) * STEP_QCAP + slot] = g.stb[p];
    }
    q[(7 + 5 * P) * STEP_QCAP + slot] = gidx;
    q[(8 + 5 * P) * STEP_QCAP + slot] = status;
}
template <int P>
__device__ __forceinline__ void queue_get(const uint32_t* q, int slot, Game<P>& g, uint32_t& gidx, uint32_t& status)
{
    g.pl0 = q[0 * STEP_QCAP + slot]; g.pl1 = q[1 * STEP_QCAP + slot]; g.pl2 = q[2 * STEP_QCAP + slot];
    g.misc = q[3 * STEP_QCAP + slot]; g.box = q[4 * STEP_QCAP + slot]; g.lid = q[5 * STEP_QCAP + slot];
    g.steps = q[6 * STEP_QCAP + slot];
#pragma unroll
    for (int p = 0; p < P; p++) {
        g.pat[p] = q[(7 + 5 * p) * STEP_QCAP + slot];  g.wall[p] = q[(8 + 5 * p) * STEP_QCAP + slot];
        g.scf[p] = q[(9 + 5 * p) * STEP_QCAP + slot];  g.sta[p] = q[(10 + 5 * p) * STEP_QCAP + slot];
        g.stb[p] = q[(11 + 5 * p) * STEP_QCAP + slot];
    }
    gidx = q[(7 + 5 * P) * STEP_QCAP + slot];
    status = q[(8 + 5 * P) * STEP_QCAP + slot];
}

