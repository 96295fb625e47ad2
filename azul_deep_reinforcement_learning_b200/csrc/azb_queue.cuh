// azb_queue.cuh -- warp-private shared-memory queue of packed games (device only).
// Used where a rare, long piece of work (the end-of-round pass) hits a few lanes of every warp: the lanes park
// their game here and the warp finishes 32 parked games at a time with every lane busy.
#pragma once
#include "azb_rules.cuh"

using namespace azb;


template <int P, int QCAP>
__device__ __forceinline__ void queue_put(uint32_t* q, int slot, const Game<P>& g, uint32_t gidx, uint32_t status)
{
    g.store(q, QCAP, slot);                            // the packed words, [word][slot]
    q[(7 + 5 * P) * QCAP + slot] = gidx;
    q[(8 + 5 * P) * QCAP + slot] = status;
}
template <int P, int QCAP>
__device__ __forceinline__ void queue_get(const uint32_t* q, int slot, Game<P>& g, uint32_t& gidx, uint32_t& status)
{
    g.load(q, QCAP, slot);
    gidx = q[(7 + 5 * P) * QCAP + slot];
    status = q[(8 + 5 * P) * QCAP + slot];
}

