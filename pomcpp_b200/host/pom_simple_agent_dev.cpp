/*
 * pom_simple_agent_dev.cpp — agents::SimpleAgent (include/agents.hpp; reference simple_agent.cpp:12-141) whose act()
 * runs the DEVICE policy code (pom_policy.cuh through pom_batch_policy_act) on the State it is given.  The agent's
 * memory is the reference's own members: moveQueue and recentPositions travel to the device as an 8-byte
 * pom_simple_agent and back.  Kept in its own translation unit: a program that links its own simple_agent.cpp
 * defines the same three symbols, and the linker then never pulls this archive member.
 */
#include <cstdio>

#include "agents.hpp"

namespace agents
{

SimpleAgent::SimpleAgent() : rng(std::random_device{}()), intDist(0, 4) {}   /* simple_agent.cpp:17-22: no bombs from the draw */

bboard::Move SimpleAgent::act(const bboard::State* state)
{
    pom_simple_agent m{};
    unsigned mq = 0;
    for(int k = 0; k < 4; k++)
    {
        const bboard::Position& p = recentPositions.queue[k];         /* physical slots; -1 / 11 survive as nibbles */
        m.recent[k] = uint8_t((p.x & 15) | ((p.y & 15) << 4));
        mq |= (unsigned(moveQueue.queue[k]) & 7u) << (3 * k);
    }
    m.rp_index = uint8_t(recentPositions.index);
    m.rp_count = uint8_t(recentPositions.count);
    m.move_queue = uint16_t(mq);

    const int move = bboard::SimpleActOnDevice(state, id, &m, intDist(rng));

    for(int k = 0; k < 4; k++)
    {
        const int x = m.recent[k] & 15, y = m.recent[k] >> 4;
        recentPositions.queue[k] = { x == 15 ? -1 : x, y == 15 ? -1 : y };
        moveQueue.queue[k] = bboard::Move((m.move_queue >> (3 * k)) & 7);
    }
    recentPositions.index = m.rp_index;
    recentPositions.count = m.rp_count;
    return bboard::Move(move);
}

void SimpleAgent::PrintDetailedInfo()
{
    for(int i = 0; i < recentPositions.count; i++)
        std::printf("(%d, %d)\n", recentPositions[i].x, recentPositions[i].y);
}

}
