// azb_internal.h -- handle, error plumbing and launch record shared by the translation units of libazb.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/azb.h"

struct azb_handle {
    int device;
    int64_t n_games;
    int players;
    int tile_pool;
    int first_player;
    uint64_t seed;
    uint64_t game_id_base;
    int block_threads;
    int block_threads_set; // azb_set_block_threads was called: no automatic choice for the rollout kernel
    int defer;             // rollout: games of a warp that must be waiting before the end-of-round pass runs
    int sm_count;
    unsigned int* sched;   // device: azb_step's row counter and exit counter (one azb_step launch per handle at a time)
};

int azb_fail(int code, const char* fmt, const char* detail = "");

#define AZB_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) return azb_fail(AZB_E_CUDA, #call ": %s", cudaGetErrorString(e_)); \
    } while (0)

#define CHECK_HANDLE(h)                                           \
    if (!(h)) return azb_fail(AZB_E_INVALID, "null handle%s");    \
    AZB_CUDA(cudaSetDevice((h)->device));

#define CHECK_LAUNCH() AZB_CUDA(cudaGetLastError())

struct Launch {
    const uint32_t* __restrict__ state_in;
    uint32_t* __restrict__ state;
    int64_t n;
    uint32_t k0, k1;       // Philox key
    uint32_t gid0;         // global id of game 0
    int first_rule;
};

static inline Launch make_launch(const azb_handle* h, uint32_t* state)
{
    Launch L;
    L.state_in = state; L.state = state; L.n = h->n_games;
    L.k0 = (uint32_t)h->seed; L.k1 = (uint32_t)(h->seed >> 32);
    L.gid0 = (uint32_t)h->game_id_base;
    L.first_rule = h->first_player;
    return L;
}
