// azb_queue.cuh -- warp-private shared-memory queue of packed games (device only).
// Used where a rare, long piece of work (the end-of-round pass) hits a few lanes of every warp: the lanes park
// their game here and the warp finishes 32 parked games at a time with every lane busy.
#pragma once
#include "azb_rules.cuh"

using namespace azb;


template <int P, int QCAP>
__device__ __forceinline__ void queue_put(uint32_t* q, int slot, const Game<P>& g, uint32_t gidx, uint32_t status)
{
    q[0 * QCAP + slot] = g.pl0; q[1 * QCAP + slot] = g.pl1; q[2 * QCAP + slot] = g.pl2;
    q[3 * QCAP + slot] = g.misc; q[4 * QCAP + slot] = g.box; q[5 * QCAP + slot] = g.lid;
    q[6 * QCAP + slot] = g.steps;
#pragma unroll
    for (int p = 0; p < P; p++) {
        q[(7 + 5 * p) * QCAP + slot] = g.pat[p];  q[(8 + 5 * p) * QCAP + slot] = g.wall[p];
        q[(9 + 5 * p) * QCAP + slot] = g.scf[p];  q[(10 + 5 * p) * QCAP + slot] = g.sta[p];
        q[(11 + 5 * p) * QCAP + slot] = g.stb[p];
    }
    q[(7 + 5 * P) * QCAP + slot] = gidx;
    q[(8 + 5 * P) * QCAP + slot] = status;
}
template <int P, int QCAP>
__device__ __forceinline__ void queue_get(const uint32_t* q, int slot, Game<P>& g, uint32_t& gidx, uint32_t& status)
{
    g.pl0 = q[0 * QCAP + slot]; g.pl1 = q[1 * QCAP + slot]; g.pl2 = q[2 * QCAP + slot];
    g.misc = q[3 * QCAP + slot]; g.box = q[4 * QCAP + slot]; g.lid = q[5 * QCAP + slot];
    g.steps = q[6 * QCAP + slot];
#pragma unroll
    for (int p = 0; p < P; p++) {
        g.pat[p] = q[(7 + 5 * p) * QCAP + slot];  g.wall[p] = q[(8 + 5 * p) * QCAP + slot];
        g.scf[p] = q[(9 + 5 * p) * QCAP + slot];  g.sta[p] = q[(10 + 5 * p) * QCAP + slot];
        g.stb[p] = q[(11 + 5 * p) * QCAP + slot];
    }
    gidx = q[(7 + 5 * P) * QCAP + slot];
    status = q[(8 + 5 * P) * QCAP + slot];
}

