/*
 * agents.hpp — drop-in for the reference's include/agents.hpp (:18-76): RandomAgent, HarmlessAgent, LazyAgent and
 * SimpleAgent with the reference's member names and layout, so that agent code written against the reference - its
 * own src/agents/simple_agent.cpp included - compiles unchanged against include/ (tests/user_agent builds exactly
 * that file).  Each struct also takes an explicit seed (the reference seeds from std::random_device only).
 *
 * Definitions: libpom_host.a.  RandomAgent / HarmlessAgent / LazyAgent are three-line host functions
 * (pomcpp_b200/host/pom_agents.cpp).  SimpleAgent::act comes from pom_simple_agent_dev.cpp, which does not restate
 * the heuristic on the host: it runs the device policy code (pom_policy.cuh, pom_batch_policy_act) on the given State,
 * with the agent's moveQueue / recentPositions as its memory.  A program that links its own simple_agent.cpp gets that
 * one instead (the archive member is only pulled in when the symbols are still undefined); with include/strategy.hpp
 * and libpom_host.a it computes on the host, as in the reference.
 */
#ifndef RANDOM_AGENT_H
#define RANDOM_AGENT_H

#include <cstdint>
#include <random>

#include "bboard.hpp"
#include "strategy.hpp"

namespace agents
{

/* basic_agents.cpp:12-22: uniform over all six moves */
struct RandomAgent : bboard::Agent
{
    std::mt19937_64 rng;
    std::uniform_int_distribution<int> intDist;

    RandomAgent();
    explicit RandomAgent(uint64_t seed);
    bboard::Move act(const bboard::State* state) override;
};

/* basic_agents.cpp:28-38: uniform over the five moves that plant no bomb */
struct HarmlessAgent : bboard::Agent
{
    std::mt19937_64 rng;
    std::uniform_int_distribution<int> intDist;

    HarmlessAgent();
    explicit HarmlessAgent(uint64_t seed);
    bboard::Move act(const bboard::State* state) override;
};

/* basic_agents.cpp:44-47 */
struct LazyAgent : bboard::Agent
{
    bboard::Move act(const bboard::State* state) override;
};

/* agents.hpp:55-76 / simple_agent.cpp:12-141 */
struct SimpleAgent : bboard::Agent
{
    std::mt19937_64 rng;
    std::uniform_int_distribution<int> intDist;

    SimpleAgent();
    explicit SimpleAgent(uint64_t seed) : SimpleAgent() { rng.seed(seed); }

    int danger = 0;
    bboard::strategy::RMap r;
    /* zero-filled here; the reference leaves both queues' slots indeterminate and reads unwritten ones
     * (simple_agent.cpp:28,48,125 - DESIGN defect D6) */
    bboard::FixedQueue<bboard::Move, bboard::MOVE_COUNT> moveQueue{};
    static const int rpCapacity = 4;
    bboard::FixedQueue<bboard::Position, rpCapacity> recentPositions{};

    bboard::Move act(const bboard::State* state) override;
    void PrintDetailedInfo();
};

}

#endif
