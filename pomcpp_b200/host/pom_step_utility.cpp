/*
 * pom_step_utility.cpp — host implementations of the read-only step helpers declared in include/step_utility.hpp
 * (reference: src/bboard/step_utility.cpp; each function cites the lines whose behaviour it reproduces).
 * Nothing here advances a game: the tick itself runs on the device (pomcpp_b200/csrc/pom_core.cuh).
 */
#include "step_utility.hpp"

namespace bboard::util
{

namespace
{
/* UP = y - 1, DOWN = y + 1, LEFT = x - 1, RIGHT = x + 1; IDLE and BOMB stay (step_utility.cpp:9-31) */
struct Delta { int dx, dy; };
inline Delta delta_of(int direction)
{
    static const Delta table[5] = { {0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0} };
    return (direction >= 1 && direction <= 4) ? table[direction] : table[0];
}
}

Position DesiredPosition(int x, int y, Move m)                        /* step_utility.cpp:9-31 */
{
    const Delta d = delta_of(int(m));
    return { x + d.dx, y + d.dy };
}

Position OriginPosition(int x, int y, Move m)                         /* :33-55 */
{
    const Delta d = delta_of(int(m));
    return { x - d.dx, y - d.dy };
}

Position DesiredPosition(const Bomb b)                                /* :57-60 */
{
    const Delta d = delta_of(BMB_DIR(b));
    return { BMB_POS_X(b) + d.dx, BMB_POS_Y(b) + d.dy };
}

void FillPositions(State* s, Position p[AGENT_COUNT])                 /* :130-136 */
{
    for(int a = 0; a < AGENT_COUNT; a++) p[a] = { s->agents[a].x, s->agents[a].y };
}

void FillDestPos(State* s, Move m[AGENT_COUNT], Position p[AGENT_COUNT])   /* :138-144 */
{
    for(int a = 0; a < AGENT_COUNT; a++) p[a] = DesiredPosition(s->agents[a].x, s->agents[a].y, m[a]);
}

void FillBombDestPos(State* s, Position p[MAX_BOMBS])                 /* :146-152 */
{
    for(int k = 0; k < s->bombs.count; k++) p[k] = DesiredPosition(s->bombs[k]);
}

/* :154-170.  Pairs are visited in the order (0,0),(0,1),..,(3,3); dead agents are not skipped; an earlier
 * fix changes what later pairs see. */
void FixSwitchMove(State* s, Position d[AGENT_COUNT])
{
    for(int a = 0; a < AGENT_COUNT; a++)
    {
        for(int b = a; b < AGENT_COUNT; b++)
        {
            const Position pa = { s->agents[a].x, s->agents[a].y }, pb = { s->agents[b].x, s->agents[b].y };
            if(d[a] == pb && d[b] == pa)
            {
                d[a] = pa;
                d[b] = pb;
            }
        }
    }
}

/* :172-205.  chain[] receives the roots in agent order; dependency[j] = i says that i wants the cell of j
 * (the caller pre-fills both with -1, step.cpp:28-31).  Dead agents are roots; a later i overwrites an earlier one. */
int ResolveDependencies(State* s, Position des[AGENT_COUNT], int dependency[AGENT_COUNT], int chain[AGENT_COUNT])
{
    int roots = 0;
    for(int i = 0; i < AGENT_COUNT; i++)
    {
        int blocker = -1;
        if(!s->agents[i].dead)
        {
            for(int j = 0; j < AGENT_COUNT && blocker < 0; j++)
            {
                if(j != i && !s->agents[j].dead && des[i].x == s->agents[j].x && des[i].y == s->agents[j].y) blocker = j;
            }
        }
        if(blocker < 0) chain[roots++] = i;
        else dependency[blocker] = i;
    }
    return roots;
}

void ConsumePowerup(State& state, int agentID, int powerUp)           /* :247-262 */
{
    AgentInfo& a = state.agents[agentID];
    switch(powerUp)
    {
    case Item::EXTRABOMB: a.maxBombCount++; break;
    case Item::INCRRANGE: a.bombStrength++; break;
    case Item::KICK:      a.canKick = true; break;
    default: break;
    }
}

bool HasDPCollision(const State& state, Position dp[AGENT_COUNT], int agentID)   /* :264-277 */
{
    for(int other = 0; other < AGENT_COUNT; other++)
    {
        if(other != agentID && !state.agents[other].dead && dp[other] == dp[agentID]) return true;
    }
    return false;
}

bool HasBombCollision(const State& state, const Bomb& b, int index)   /* :279-293: bombs are compared by VALUE */
{
    const Position mine = DesiredPosition(b);
    for(int k = index; k < state.bombs.count; k++)
    {
        const Bomb other = state.bombs[k];
        if(other != b && DesiredPosition(other) == mine) return true;
    }
    return false;
}

void ResetBombFlags(State& state)                                     /* :331-337 */
{
    for(int k = 0; k < state.bombs.count; k++) SetBombMovedFlag(state.bombs[k], false);
}

}
