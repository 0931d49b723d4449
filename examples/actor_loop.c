/*
 * actor_loop.c — the C ABI (include/pom_batch.h) from plain C: an RL-style actor loop.
 *
 * My agent is agent 0 of every game (here: a trivial "walk right, sometimes bomb" policy written on the host);
 * agents 1-3 are the reference's heuristic SimpleAgent, computed on the GPU.  The games are split into two
 * half-batches that are stepped alternately, so the GPU works on one half while the host handles the other.
 *
 *   cc -std=c11 -O2 -I include examples/actor_loop.c -L pomcpp_b200 -lpom_b200 -Wl,-rpath,$PWD/pomcpp_b200 -o actor_loop
 *   ./actor_loop [envs_per_half] [ticks]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pom_batch.h"

#define CHECK(call) do { if((call) != POM_OK) { fprintf(stderr, "%s: %s\n", #call, pom_last_error()); return 1; } } while(0)

static void my_policy(uint8_t* moves, const uint8_t* status, uint64_t n, uint32_t tick)
{
    for(uint64_t e = 0; e < n; e++)
    {
        (void)status;                                   /* a real policy would look at observations / done flags here */
        moves[4 * e] = (uint8_t)(((e + tick) % 7 == 0) ? POM_MOVE_BOMB : POM_MOVE_RIGHT);
    }
}

int main(int argc, char** argv)
{
    const uint64_t n = argc > 1 ? strtoull(argv[1], NULL, 10) : 65536;
    const uint32_t ticks = argc > 2 ? (uint32_t)atoi(argv[2]) : 200;
    pom_batch* half[2];
    uint8_t* moves[2];
    uint8_t* status[2];
    for(int i = 0; i < 2; i++)
    {
        pom_init_desc d;
        memset(&d, 0, sizeof d);
        d.env_offset = (uint64_t)i * n;                 /* global env index: keys the boards and the RNG streams */
        d.n_templates = 1024;
        d.first_seed = 0x1337;
        d.max_ticks = 800;
        CHECK(pom_batch_init(&half[i], 0, n, &d));
        CHECK(pom_host_alloc(4 * n, (void**)&moves[i]));    /* pinned + mapped: the kernels read and write them directly */
        CHECK(pom_host_alloc(n, (void**)&status[i]));
        memset(moves[i], 0, 4 * n);
        memset(status[i], 0, n);
    }
    const uint64_t seed = 2024;
    /* per half and tick: my move (host) -> opponents' moves (device) -> step; results come back in status[] */
    for(uint32_t t = 0; t < ticks; t++)
    {
        for(int i = 0; i < 2; i++)
        {
            if(t > 0) CHECK(pom_batch_sync(half[i]));                               /* status[i] of tick t-1 is valid now */
            my_policy(moves[i], status[i], n, t);
            CHECK(pom_batch_policy_moves_host(half[i], moves[i], seed, t, 0xE));    /* fills bytes 1..3 of every env */
            CHECK(pom_batch_step_host_async(half[i], moves[i], status[i], POM_STEP_AUTORESET | POM_STEP_COUNT));
        }
    }
    unsigned long long steps = 0, episodes = 0, wins0 = 0;
    for(int i = 0; i < 2; i++)
    {
        pom_stats s;
        CHECK(pom_batch_sync(half[i]));
        CHECK(pom_batch_stats(half[i], &s));
        steps += s.env_steps; episodes += s.episodes; wins0 += s.wins[0];
        pom_batch_destroy(half[i]);
        pom_host_free(moves[i]);
        pom_host_free(status[i]);
    }
    printf("actor_loop: %llu env-steps, %llu episodes finished, agent 0 won %llu of them\n", steps, episodes, wins0);
    return steps == 2ull * n * ticks ? 0 : 1;
}
