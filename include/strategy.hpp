/*
 * strategy.hpp — drop-in for the reference's include/strategy.hpp (namespace bboard::strategy, :29-186): the
 * reachability map and the movement heuristics that agent code such as the reference's own SimpleAgent
 * (src/agents/simple_agent.cpp) is written against.  They only read the State handed to Agent::act, so they are plain
 * host code (pomcpp_b200/host/pom_strategy.cpp, libpom_host.a) and give the reference's answers cell for cell,
 * including the scan orders and the loop-bound slip of MoveTowardsSafePlace (strategy.cpp:126-144).  The batched
 * device counterpart (bitboard floods, pomcpp_b200/csrc/pom_policy.cuh) is what pom_batch_policy_moves runs.
 * PrintMap / PrintPath (console output) are out of scope.
 */
#ifndef STRATEGY_H
#define STRATEGY_H

#include "bboard.hpp"
#include "step_utility.hpp"

static_assert(sizeof(int) == 4 || sizeof(int) == 8, "32/64 bit integer");

namespace bboard::strategy
{

typedef unsigned int RMapInfo;
const int chalf = 0xFFFF;

/* strategy.hpp:37-62: per cell, low half = BFS distance from `source` (0 = not reached), high half = index x + 11 y of
 * the predecessor on the shortest path */
struct RMap
{
    int map[BOARD_SIZE][BOARD_SIZE] = {};
    RMapInfo info;
    Position source;

    int  GetDistance(int x, int y) const { return map[y][x] & chalf; }
    void SetDistance(int x, int y, int distance) { map[y][x] = (map[y][x] & ~chalf) + distance; }
    int  GetPredecessor(int x, int y) const { return map[y][x] >> 16; }
    void SetPredecessor(int x, int y, int xPredecessor, int yPredecessor)
    {
        map[y][x] = (map[y][x] & chalf) + ((xPredecessor + BOARD_SIZE * yPredecessor) << 16);
    }
};

void FillRMap(const State& s, RMap& r, int agentID);                  /* :68 */
inline bool IsReachable(RMap& r, int x, int y) { return r.GetDistance(x, y) != 0; }   /* :74-77 */

bool IsAdjacentEnemy(const State& state, int agentID, int distance);  /* :87 */
bool IsAdjacentItem(const State& state, int agentID, int distance, Item item);   /* :93 */
Move MoveTowardsPosition(const RMap& r, const Position& position);    /* :99 */
Move MoveTowardsSafePlace(const State& state, const RMap& r, int radius);   /* :112 */
Move MoveTowardsPowerup(const State& state, const RMap& r, int radius);     /* :122 */
Move MoveTowardsEnemy(const State& state, const RMap& r, int radius);       /* :132 */
void SafeDirections(const State& state, FixedQueue<Move, MOVE_COUNT>& q, int x, int y);   /* :138 */

/* strategy.hpp:144-168: directions that lead to a recently visited position go to the back of the queue.  The move put
 * back is the one that slid into slot i when slot i was removed (the reference reads q[i] after RemoveAt(i)); agents
 * depend on that order, so it is kept. */
template <int X>
void SortDirections(FixedQueue<Move, MOVE_COUNT>& q, FixedQueue<Position, X>& p, int x, int y)
{
    const int initial = q.count;
    int moved = 0;
    int i = 0;
    /* slot i is looked at again after a removal: only a direction that stays advances i */
    while(i < initial && moved < MOVE_COUNT)
    {
        const Position target = util::DesiredPosition(x, y, q[i]);
        bool recent = false;
        for(int j = 0; j < p.count && !recent; j++) recent = target == p[j];
        if(recent)
        {
            q.RemoveAt(i);
            q.AddElem(q[i]);
            moved++;
        }
        else
        {
            i++;
        }
    }
}

int IsInDanger(const State& state, int agentID);                      /* :174 */
int IsInDanger(const State& state, int x, int y);                     /* :175 */

inline bool IsInBombRange(int x, int y, int s, const Position& pos)   /* :181-186 */
{
    const bool sameRow = pos.y == y && pos.x >= x - s && pos.x <= x + s;
    const bool sameColumn = pos.x == x && pos.y >= y - s && pos.y <= y + s;
    return sameRow || sameColumn;
}

bool _safe_condition(int danger, int min = 2);

}

#endif
