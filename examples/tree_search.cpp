/*
 * tree_search.cpp — code written against the reference's C++ API (bboard.hpp) plus the batch extensions.
 *
 *  1. a single game with four SimpleAgents through bboard::Environment, exactly as with the reference
 *     (src/main.cpp, unit_test/bboard/live_testing.cpp) — Step and SimpleAgent::act run on the GPU;
 *  2. one ply of tree search on the batch engine: every root state is expanded into its 6^4 joint actions with one
 *     fused clone + Step per child (pom_batch_expand_step), and the children are rolled out for a few ticks with
 *     SimpleAgent opponents to score agent 0's moves.
 *
 *   g++ -std=c++17 -O2 -I include -I pomcpp_b200/host examples/tree_search.cpp pomcpp_b200/host/libpom_host.a \
 *       -L pomcpp_b200 -lpom_b200 -Wl,-rpath,$PWD/pomcpp_b200 -o tree_search
 */
#include <cstdio>
#include <cstring>
#include <vector>

#include "bboard.hpp"
#include "pom_agents.hpp"

#define CHECK(call) do { if((call) != POM_OK) { std::fprintf(stderr, "%s: %s\n", #call, pom_last_error()); return 1; } } while(0)

int main()
{
    /* ---- 1. the reference's own usage pattern */
    agents::SimpleAgent a[4] = {agents::SimpleAgent(1), agents::SimpleAgent(2), agents::SimpleAgent(3), agents::SimpleAgent(4)};
    bboard::Environment env;
    env.MakeGame({&a[0], &a[1], &a[2], &a[3]});
    env.StartGame(40, false);
    const bboard::State root = env.GetState();
    std::printf("tree_search: after %d ticks %d agents are alive, %d bombs on the board\n", root.timeStep, root.aliveAgents, root.bombs.count);

    /* ---- 2. expand the position reached above: 1296 children, each = root + one joint action */
    pom_init_desc d;
    std::memset(&d, 0, sizeof d);
    d.n_templates = 1; d.first_seed = 0x1337; d.flags = POM_INIT_EMPTY; d.max_ticks = 800;
    pom_batch* roots = nullptr; pom_batch* kids = nullptr;
    CHECK(pom_batch_init(&roots, 0, 1, &d));
    CHECK(pom_batch_init(&kids, 0, 1296, &d));
    CHECK(pom_batch_upload(roots, 0, 1, reinterpret_cast<const pom_state*>(&root), nullptr));
    const uint32_t idx[1] = {0};
    CHECK(pom_batch_expand_step(kids, roots, idx, 1, 1296, 0));
    /* play every child on for 30 ticks: agent 0 random, agents 1-3 SimpleAgent; finished games freeze */
    CHECK(pom_batch_rollout(kids, 30, 99, 0, POM_ROLL_NO_RESET | POM_ROLL_SIMPLE(0xE)));
    std::vector<bboard::State> out(1296);
    std::vector<uint8_t> status(1296);
    CHECK(pom_batch_download(kids, 0, 1296, reinterpret_cast<pom_state*>(out.data()), status.data()));
    /* child j carried the joint action a_k = (j / 6^k) % 6: score agent 0's six first moves by its survival rate */
    int alive[6] = {0, 0, 0, 0, 0, 0};
    for(int j = 0; j < 1296; j++) alive[j % 6] += out[size_t(j)].agents[0].dead ? 0 : 1;
    static const char* name[6] = {"IDLE", "UP", "DOWN", "LEFT", "RIGHT", "BOMB"};
    for(int m = 0; m < 6; m++) std::printf("  agent 0 plays %-5s -> alive in %3d of 216 continuations after 30 ticks\n", name[m], alive[m]);
    pom_batch_destroy(roots);
    pom_batch_destroy(kids);
    return 0;
}
