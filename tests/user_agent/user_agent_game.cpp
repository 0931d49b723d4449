/*
 * user_agent_game.cpp — TEST: agent code written against the reference runs unchanged on this library.
 *
 * Built (oracle/Makefile, target _ref/user_agent_game) from THIS file + the UNMODIFIED reference source
 * $(REF)/src/agents/simple_agent.cpp, compiled against include/ only (bboard.hpp, agents.hpp, strategy.hpp,
 * step_utility.hpp of this repo) and linked with libpom_host.a + libpom_b200.so.  So agents::SimpleAgent::act below is
 * the reference's own text, its strategy:: / util:: calls resolve to this repo's host helpers, and
 * BatchEnvironment::Step(agents) steps the game on the GPU.
 *
 * Check: the same games are played a second time by the DEVICE SimpleAgent (BatchEnvironment::Step(moves, mask, seed),
 * pom_policy.cuh).  Both get the same draws: the device takes byte a of pom_rng_moves(seed, game, tick, 5); the host
 * agent's engine is re-seeded before every act() with a seed whose first intDist(0,4) draw is that value.  Every tick the
 * two States must be identical field for field.
 *
 * usage: user_agent_game [games] [ticks]     exit code 0 = identical
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "bboard.hpp"
#include "agents.hpp"

using namespace bboard;

namespace
{

unsigned long long seedForDraw[5];

void findSeeds()
{
    bool have[5] = {false, false, false, false, false};
    int found = 0;
    for(unsigned long long sd = 1; found < 5; sd++)
    {
        std::mt19937_64 g(sd);
        std::uniform_int_distribution<int> d(0, 4);
        const int v = d(g);
        if(!have[v]) { have[v] = true; seedForDraw[v] = sd; found++; }
    }
}

/* the reference's agent, with the one thing a test must control: its random draw */
struct DrivenSimpleAgent : agents::SimpleAgent
{
    uint64_t seed = 0, game = 0;
    uint32_t tick = 0;
    Move act(const State* state) override
    {
        const uint32_t draws = pom_rng_moves(seed, game, tick, 5);
        rng.seed(seedForDraw[(draws >> (8 * id)) & 0xFFu]);
        return agents::SimpleAgent::act(state);
    }
};

bool sameState(const State& a, const State& b)
{
    if(std::memcmp(a.board, b.board, sizeof a.board) != 0 || a.timeStep != b.timeStep || a.aliveAgents != b.aliveAgents) return false;
    for(int i = 0; i < AGENT_COUNT; i++)
    {
        const AgentInfo &p = a.agents[i], &q = b.agents[i];
        if(p.x != q.x || p.y != q.y || p.bombCount != q.bombCount || p.maxBombCount != q.maxBombCount ||
           p.bombStrength != q.bombStrength || p.canKick != q.canKick || p.dead != q.dead) return false;
    }
    if(a.bombs.count != b.bombs.count || a.bombs.index != b.bombs.index || a.flames.count != b.flames.count || a.flames.index != b.flames.index) return false;
    for(int k = 0; k < a.bombs.count; k++) if(a.bombs[k] != b.bombs[k]) return false;
    for(int k = 0; k < a.flames.count; k++)
    {
        const Flame &f = a.flames[k], &g = b.flames[k];
        if(!(f.position == g.position) || f.timeLeft != g.timeLeft || f.strength != g.strength) return false;
    }
    return true;
}

}

int main(int argc, char** argv)
{
    const int games = argc > 1 ? std::atoi(argv[1]) : 6;
    const int ticks = argc > 2 ? std::atoi(argv[2]) : 120;
    const uint64_t seed = 4711;
    findSeeds();
    long acts = 0;
    for(int g = 0; g < games; g++)
    {
        /* one game per environment: the four agent objects of Step(agents) carry per-game memory */
        BatchEnvironment host(1, 0, uint64_t(g), 1, 0x1337 + 7 * g), dev(1, 0, uint64_t(g), 1, 0x1337 + 7 * g);
        DrivenSimpleAgent a[4];
        for(int i = 0; i < 4; i++) { a[i].id = i; a[i].seed = seed; a[i].game = uint64_t(g); }
        const Move idle[4] = {Move::IDLE, Move::IDLE, Move::IDLE, Move::IDLE};
        for(int t = 0; t < ticks; t++)
        {
            for(int i = 0; i < 4; i++) a[i].tick = uint32_t(t);
            for(int i = 0; i < 4; i++) acts += !host.States()[0].agents[i].dead;
            const size_t r1 = host.Step({&a[0], &a[1], &a[2], &a[3]});
            const size_t r2 = dev.Step(idle, 0xF, seed);
            if(r1 != r2 || !sameState(host.States()[0], dev.States()[0]) || host.Status()[0] != dev.Status()[0])
            {
                std::printf("user_agent_game: game %d differs at tick %d\n", g, t);
                return 1;
            }
            if(r1 == 0) break;
        }
    }
    std::printf("user_agent_game: %d games, %ld act() calls of the reference's simple_agent.cpp == device SimpleAgent\n", games, acts);
    return 0;
}
