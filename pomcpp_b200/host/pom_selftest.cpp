/*
 * pom_selftest.cpp — known-answer checks written against the bboard mirror (include/pom_bboard.hpp) in the
 * style of the reference's Catch2 suite (unit_test/bboard/board_logic.cpp, general_test.cpp): the fixtures
 * and REQUIREs of a representative subset, with bboard::Step running on the GPU.  The complete scenario list
 * runs from Python (tests/scenarios.py); this binary proves that code written against the reference header
 * compiles and behaves the same when pointed at this library.  Exit code 0 = all passed.
 */
#include <cstdio>
#include <cstring>
#include <memory>
#include <random>

#include "bboard.hpp"
#include "pom_agents.hpp"
#include "strategy.hpp"
#include "step_utility.hpp"

using namespace bboard;

static int failures = 0;
#define REQUIRE(cond) do { if(!(cond)) { std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); failures++; } } while(0)

static void REQUIRE_AGENT(State* s, int agent, int x, int y)          /* board_logic.cpp:11-17 */
{
    REQUIRE(s->agents[agent].x == x);
    REQUIRE(s->agents[agent].y == y);
    REQUIRE(s->board[y][x] == Item::AGENT0 + agent);
}

static void SeveralSteps(int times, State* s, Move* m) { for(int i = 0; i < times; i++) bboard::Step(s, m); }

static void TestQueue(FixedQueue<Bomb, 10>& q)                        /* general_test.cpp:8-40 */
{
    for(int i = 0; i < 10; i++) { q.NextPos() = i; q.count++; }
    REQUIRE(q.count == 10);
    q.PopElem(); q.PopElem(); q.PopElem();
    REQUIRE(q.count == 7); REQUIRE(q[0] == 3);
    q.RemoveAt(5);
    REQUIRE(q.count == 6); REQUIRE(q[4] == 7); REQUIRE(q[5] == 9);
    q.RemoveAt(0);
    REQUIRE(q[0] == 4);
    q.RemoveAt(4);
    REQUIRE(q.count == 4); REQUIRE(q[3] == 7);
}

int main()
{
    const Move id = Move::IDLE;
    {   /* "Fixed Size Queue", general_test.cpp:41-61 */
        for(int start : {0, 5, 2})
        {
            auto q = std::make_unique<FixedQueue<Bomb, 10>>();
            q->index = start;
            TestQueue(*q);
        }
    }
    {   /* "Basic Non-Obstacle Movement", board_logic.cpp:55-83 */
        auto s = std::make_unique<State>();
        s->PutAgentsInCorners(0, 1, 2, 3);
        Move m[4] = {id, id, id, id};
        m[0] = Move::RIGHT; Step(s.get(), m); REQUIRE_AGENT(s.get(), 0, 1, 0);
        m[0] = Move::DOWN;  Step(s.get(), m); REQUIRE_AGENT(s.get(), 0, 1, 1);
        m[0] = Move::LEFT;  Step(s.get(), m); REQUIRE_AGENT(s.get(), 0, 0, 1);
        m[0] = Move::UP;    Step(s.get(), m); REQUIRE_AGENT(s.get(), 0, 0, 0);
        m[3] = Move::UP;    Step(s.get(), m); REQUIRE_AGENT(s.get(), 3, 0, 9);
    }
    {   /* "Movement Against Flames", :104-119 */
        auto s = std::make_unique<State>();
        Move m[4] = {id, id, id, id};
        s->PutAgentsInCorners(0, 1, 2, 3);
        s->SpawnFlame(1, 1, 2);
        m[0] = Move::RIGHT;
        Step(s.get(), m);
        REQUIRE(s->agents[0].dead);
        REQUIRE(s->board[0][0] == Item::PASSAGE);
    }
    {   /* "Movement Dependency Handling" / "Move Ouroboros", :221-238 */
        auto s = std::make_unique<State>();
        s->PutAgent(0, 0, 0); s->PutAgent(1, 0, 1); s->PutAgent(1, 1, 2); s->PutAgent(0, 1, 3);
        Move m[4] = {Move::RIGHT, Move::DOWN, Move::LEFT, Move::UP};
        Step(s.get(), m);
        REQUIRE_AGENT(s.get(), 3, 0, 0); REQUIRE_AGENT(s.get(), 0, 1, 0);
        REQUIRE_AGENT(s.get(), 1, 1, 1); REQUIRE_AGENT(s.get(), 2, 0, 1);
    }
    {   /* "Bomb Explosion" / "Destroy Objects and Agents", :331-345 */
        auto s = std::make_unique<State>();
        Move m[4] = {id, id, id, id};
        s->Kill(2, 3);
        s->PutAgent(5, 5, 0);
        s->PutItem(6, 5, Item::WOOD);
        s->PutAgent(4, 5, 1);
        m[0] = Move::BOMB; Step(s.get(), m);
        m[0] = Move::UP; SeveralSteps(BOMB_LIFETIME, s.get(), m);
        REQUIRE(s->agents[1].dead);
        REQUIRE(IS_FLAME(s->board[5][4]));
        REQUIRE(IS_FLAME(s->board[5][6]));
    }
    {   /* "Chained Explosions" / "Two Bombs Covered By Agent", :450-469 */
        auto s = std::make_unique<State>();
        Move m[4] = {id, id, id, id};
        s->PutAgent(5, 5, 0); s->PutAgent(4, 5, 1);
        s->Kill(2, 3);
        m[0] = Move::BOMB; Step(s.get(), m);
        m[1] = Move::BOMB; Step(s.get(), m);
        m[0] = m[1] = Move::DOWN;
        SeveralSteps(BOMB_LIFETIME - 2, s.get(), m);
        REQUIRE(s->bombs.count == 2);
        Step(s.get(), m);
        REQUIRE(s->bombs.count == 0);
        REQUIRE(s->flames.count == 2);
    }
    {   /* "Bomb Kick Mechanics" / "Bounce Back Agent", :474-484, 549-562 */
        auto s = std::make_unique<State>();
        Move m[4] = {id, id, id, id};
        s->PutAgent(0, 1, 0);
        s->agents[0].canKick = true;
        s->PlantBomb(1, 1, 0, true);
        s->agents[0].maxBombCount = MAX_BOMBS_PER_AGENT;
        m[0] = Move::RIGHT;
        s->Kill(2, 3);
        s->PutAgent(0, 2, 1);
        m[1] = Move::UP;
        s->PlantBomb(2, 2, 0, true);
        SetBombDirection(s->bombs[1], Direction::UP);
        Step(s.get(), m);
        REQUIRE_AGENT(s.get(), 0, 0, 1); REQUIRE_AGENT(s.get(), 1, 0, 2);
        REQUIRE(BMB_POS_X(s->bombs[0]) == 1); REQUIRE(BMB_POS_X(s->bombs[1]) == 2);
    }
    {   /* Environment + trivial agents: a game runs to its end on the GPU (environment.cpp:68-88) */
        agents::RandomAgent r[4] = {agents::RandomAgent(1), agents::RandomAgent(2), agents::RandomAgent(3), agents::RandomAgent(4)};
        Environment env;
        env.MakeGame({&r[0], &r[1], &r[2], &r[3]});
        env.StartGame(800, false);
        REQUIRE(env.IsDone() || env.GetState().timeStep == 800);
        REQUIRE(env.GetState().aliveAgents <= 1 || !env.IsDone());
    }
    {   /* BatchEnvironment: host agents on 512 games, then a fused rollout */
        agents::HarmlessAgent h(7);
        BatchEnvironment be(512, 0, 0, 64);
        size_t running = 0;
        for(int t = 0; t < 20; t++) running = be.Step({&h, &h, &h, &h});
        REQUIRE(running == 512);                       /* harmless agents cannot end a game */
        REQUIRE(be.States()[0].timeStep == 20);
        pom_stats st = be.Rollout(200, 5);
        REQUIRE(st.env_steps == 512ull * 200ull);
        REQUIRE(st.episodes == st.wins[0] + st.wins[1] + st.wins[2] + st.wins[3] + st.draws + st.truncated + st.invalid);
    }
    {   /* "Test Simple Agent" (live_testing.cpp:18-36) without the console: four SimpleAgents play a game through
         * Environment; their first moves on the default board are known: nobody is in danger, no enemy within 7, wood
         * next to every corner => agents bomb or step safely, never walk into a wall */
        agents::SimpleAgent a[4] = {agents::SimpleAgent(11), agents::SimpleAgent(12), agents::SimpleAgent(13), agents::SimpleAgent(14)};
        Environment env;
        env.MakeGame({&a[0], &a[1], &a[2], &a[3]});
        for(int t = 0; t < 60 && !env.IsDone(); t++)
        {
            const State before = env.GetState();
            env.Step();
            for(int i = 0; i < 4; i++)
            {
                if(before.agents[i].dead) continue;
                const Move m = env.GetLastMove(i);
                REQUIRE(int(m) >= 0 && int(m) <= 5);
                if(m == Move::BOMB) REQUIRE(before.agents[i].bombCount < before.agents[i].maxBombCount);
            }
        }
        REQUIRE(env.GetState().timeStep > 0);
        REQUIRE(a[0].recentPositions.count > 0);              /* the agents remember where they went */
    }
    {   /* BatchEnvironment with device-side SimpleAgent opponents: per tick (agent 0 host-controlled) and fused */
        BatchEnvironment be(256, 0, 0, 32, 0x1337, 800);
        std::vector<Move> mv(4 * 256, Move::IDLE);
        size_t running = 256;
        for(int t = 0; t < 30; t++) running = be.Step(mv.data(), 0xE, 99);
        REQUIRE(running > 0);
        REQUIRE(be.States()[0].timeStep > 0);
        pom_stats st = be.Rollout(300, 5, false, 0xF);
        REQUIRE(st.env_steps > 0);
        REQUIRE(st.episodes == st.wins[0] + st.wins[1] + st.wins[2] + st.wins[3] + st.draws + st.truncated + st.invalid);
    }
    {   /* StepSequence (one launch, given moves) == the same moves tick by tick */
        const size_t n = 200; const uint32_t ticks = 25;
        BatchEnvironment a(n, 0, 0, 16), b(n, 0, 0, 16);
        std::vector<Move> seq(size_t(ticks) * n * 4);
        std::mt19937 rng(5);
        for(auto& m : seq) m = Move(int(rng() % 6));
        size_t ra = 0;
        for(uint32_t t = 0; t < ticks; t++) ra = a.Step(seq.data() + size_t(t) * n * 4);
        const size_t rb = b.StepSequence(seq.data(), ticks);
        REQUIRE(ra == rb);
        bool same = true;
        for(size_t i = 0; i < n; i++)
            same = same && std::memcmp(a.States()[i].board, b.States()[i].board, sizeof(a.States()[i].board)) == 0 &&
                   a.States()[i].timeStep == b.States()[i].timeStep && a.Status()[i] == b.Status()[i];
        REQUIRE(same);
    }
    {   /* fog of war: host agents see a 9x9 window; what lies outside is Item::FOG and hidden agents have no position */
        struct Peek : Agent
        {
            int fogCells = 0, hidden = 0, calls = 0;
            Move act(const State* s) override
            {
                calls++;
                for(int y = 0; y < BOARD_SIZE; y++) for(int x = 0; x < BOARD_SIZE; x++) fogCells += s->board[y][x] == Item::FOG;
                for(int i = 0; i < AGENT_COUNT; i++) hidden += (i != id && s->agents[i].x < 0);
                return Move::IDLE;
            }
        } peek;
        BatchEnvironment be(64, 0, 0, 8);
        be.SetViewRange(4);
        be.Step({&peek, &peek, &peek, &peek});
        REQUIRE(peek.calls == 64 * 4);
        REQUIRE(peek.fogCells == 64 * 4 * (121 - 25));       /* from a corner only 5 x 5 cells are in the window */
        REQUIRE(peek.hidden == 64 * 4 * 3);                   /* the other corners are 10 cells away */
        std::vector<State> all = be.Observe(0, 10);
        REQUIRE(all[0].agents[2].x == 10 && all[0].agents[2].y == 10);
    }
    {   /* fog of war + SimpleAgent (runs on the device): hidden agents (x = y = -1) are simply absent for the policy */
        agents::SimpleAgent a[4] = {agents::SimpleAgent(21), agents::SimpleAgent(22), agents::SimpleAgent(23), agents::SimpleAgent(24)};
        BatchEnvironment be(4, 0, 0, 4);
        be.SetViewRange(4);
        size_t running = 4;
        for(int t = 0; t < 25; t++) running = be.Step({&a[0], &a[1], &a[2], &a[3]});
        REQUIRE(running <= 4);
        REQUIRE(be.States()[0].timeStep > 0);
    }
    {   /* bboard::StartGame(State*, Agent*[], int) (bboard.hpp:677): act x 4 -> Step, `timeSteps` times */
        auto s = std::make_unique<State>();
        InitState(s.get(), 0, 1, 2, 3);
        agents::HarmlessAgent h[4] = {agents::HarmlessAgent(1), agents::HarmlessAgent(2), agents::HarmlessAgent(3), agents::HarmlessAgent(4)};
        Agent* four[4] = {&h[0], &h[1], &h[2], &h[3]};
        for(int i = 0; i < 4; i++) four[i]->id = i;
        StartGame(s.get(), four, 12);
        REQUIRE(s->aliveAgents == 4);
        int onBoard = 0;
        for(int i = 0; i < 4; i++) onBoard += s->board[s->agents[i].y][s->agents[i].x] == Item::AGENT0 + i;
        REQUIRE(onBoard == 4);
    }
    {   /* the host helpers of include/strategy.hpp / step_utility.hpp, as agent code calls them (strategy_test.cpp:16-37,
         * step_utility_test.cpp): known answers on a small fixture */
        auto s = std::make_unique<State>();
        s->PutAgentsInCorners(0, 1, 2, 3);
        s->PutItem(1, 0, Item::RIGID);
        strategy::RMap r;
        strategy::FillRMap(*s, r, 0);
        REQUIRE(r.GetDistance(0, 1) == 1 && r.GetDistance(1, 1) == 2 && r.GetDistance(1, 0) == 0);
        REQUIRE(r.GetDistance(10, 10) == 20);                 /* paths end AT agent cells */
        REQUIRE(strategy::MoveTowardsPosition(r, {3, 1}) == Move::DOWN);
        REQUIRE(strategy::IsAdjacentEnemy(*s, 0, 9) == false && strategy::IsAdjacentEnemy(*s, 0, 10) == true);
        s->PlantBomb(0, 3, 1, true);
        REQUIRE(strategy::IsInDanger(*s, 0, 2) == BOMB_LIFETIME && strategy::IsInDanger(*s, 1, 2) == 0);
        Position d[4];
        Move m[4] = {Move::RIGHT, Move::LEFT, Move::IDLE, Move::IDLE};
        s->PutAgent(4, 4, 0); s->PutAgent(5, 4, 1);
        util::FillDestPos(s.get(), m, d);
        REQUIRE((d[0] == Position{5, 4}) && (d[1] == Position{4, 4}));
        util::FixSwitchMove(s.get(), d);
        REQUIRE((d[0] == Position{4, 4}) && (d[1] == Position{5, 4}));
        int dep[4] = {-1, -1, -1, -1}, chain[4] = {-1, -1, -1, -1};
        m[1] = Move::IDLE;
        util::FillDestPos(s.get(), m, d);
        REQUIRE(util::ResolveDependencies(s.get(), d, dep, chain) == 3 && dep[1] == 0 && chain[0] == 1);
    }
    std::printf(failures ? "pom_selftest: %d FAILED\n" : "pom_selftest: all passed\n", failures);
    return failures ? 1 : 0;
}
