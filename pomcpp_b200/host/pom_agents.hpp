/*
 * pom_agents.hpp — the reference's trivial policies (include/agents.hpp:18-48, src/agents/basic_agents.cpp)
 * as host Agent objects: uniform{0..5}, uniform{0..4}, always IDLE.  Unlike the reference they can be seeded.
 */
#ifndef POM_AGENTS_HPP_
#define POM_AGENTS_HPP_

#include <random>
#include "pom_bboard.hpp"

namespace agents
{

struct RandomAgent : bboard::Agent
{
    std::mt19937_64 rng;
    std::uniform_int_distribution<int> intDist{0, 5};
    RandomAgent() : rng(std::random_device{}()) {}
    explicit RandomAgent(uint64_t seed) : rng(seed) {}
    bboard::Move act(const bboard::State*) override { return bboard::Move(intDist(rng)); }
};

struct HarmlessAgent : bboard::Agent
{
    std::mt19937_64 rng;
    std::uniform_int_distribution<int> intDist{0, 4};
    HarmlessAgent() : rng(std::random_device{}()) {}
    explicit HarmlessAgent(uint64_t seed) : rng(seed) {}
    bboard::Move act(const bboard::State*) override { return bboard::Move(intDist(rng)); }
};

struct LazyAgent : bboard::Agent
{
    bboard::Move act(const bboard::State*) override { return bboard::Move::IDLE; }
};

}

#endif
