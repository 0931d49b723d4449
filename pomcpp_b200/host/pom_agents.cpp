/*
 * pom_agents.cpp — the reference's trivial policies (src/agents/basic_agents.cpp:12-47) behind include/agents.hpp.
 * Seeded from std::random_device like the reference, or from the caller's seed.
 */
#include "agents.hpp"

namespace agents
{

RandomAgent::RandomAgent() : rng(std::random_device{}()), intDist(0, 5) {}
RandomAgent::RandomAgent(uint64_t seed) : rng(seed), intDist(0, 5) {}
bboard::Move RandomAgent::act(const bboard::State*) { return bboard::Move(intDist(rng)); }

HarmlessAgent::HarmlessAgent() : rng(std::random_device{}()), intDist(0, 4) {}
HarmlessAgent::HarmlessAgent(uint64_t seed) : rng(seed), intDist(0, 4) {}
bboard::Move HarmlessAgent::act(const bboard::State*) { return bboard::Move(intDist(rng)); }

bboard::Move LazyAgent::act(const bboard::State*) { return bboard::Move::IDLE; }

}
