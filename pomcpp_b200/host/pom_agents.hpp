/*
 * pom_agents.hpp — the reference's policies (include/agents.hpp:18-76, src/agents/basic_agents.cpp,
 * src/agents/simple_agent.cpp) as host Agent objects: uniform{0..5}, uniform{0..4}, always IDLE, and the heuristic
 * SimpleAgent.  Unlike the reference they can be seeded.  SimpleAgent::act does not run on the host: like
 * bboard::Step it executes the device code (pom_policy.cuh) on the given State; for throughput use
 * BatchEnvironment::Rollout / Step with a simpleMask, which keep states and agent memories on the GPU.
 */
#ifndef POM_AGENTS_HPP_
#define POM_AGENTS_HPP_

#include <random>
#include "pom_bboard.hpp"

namespace agents
{

/* uniform{0 .. N_ACTIONS-1}; the members keep the reference's names (rng, intDist) */
template<int N_ACTIONS>
struct UniformAgent : bboard::Agent
{
    std::mt19937_64 rng;
    std::uniform_int_distribution<int> intDist{0, N_ACTIONS - 1};
    UniformAgent() : rng(std::random_device{}()) {}
    explicit UniformAgent(uint64_t seed) : rng(seed) {}
    bboard::Move act(const bboard::State*) override { return bboard::Move(intDist(rng)); }
};

struct RandomAgent : UniformAgent<6> { using UniformAgent<6>::UniformAgent; };      /* basic_agents.cpp:12-22: moves and bombs */
struct HarmlessAgent : UniformAgent<5> { using UniformAgent<5>::UniformAgent; };    /* basic_agents.cpp:28-38: never plants a bomb */

struct LazyAgent : bboard::Agent
{
    bboard::Move act(const bboard::State*) override { return bboard::Move::IDLE; }
};

/* agents::SimpleAgent (include/agents.hpp:55-76): same decisions as the reference's, one intDist draw per act */
struct SimpleAgent : bboard::Agent
{
    std::mt19937_64 rng;
    std::uniform_int_distribution<int> intDist{0, 4};
    pom_simple_agent memory{};        /* recentPositions + moveQueue (zero = a freshly constructed agent) */
    SimpleAgent() : rng(std::random_device{}()) {}
    explicit SimpleAgent(uint64_t seed) : rng(seed) {}
    bboard::Move act(const bboard::State* state) override
    {
        return bboard::Move(bboard::SimpleActOnDevice(state, id, &memory, intDist(rng)));
    }
};

}

#endif
