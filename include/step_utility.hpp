/*
 * step_utility.hpp — drop-in for the reference's include/step_utility.hpp (namespace bboard::util, :16-166):
 * the read-only helpers of the step path that agent code calls (simple_agent.cpp:1-4,12-15 uses DesiredPosition and
 * IsOutOfBounds; strategy.hpp's SortDirections uses DesiredPosition), plus the two field-level writers that are not
 * simulation (ConsumePowerup, ResetBombFlags).  Host implementations in pomcpp_b200/host/pom_step_utility.cpp
 * (libpom_host.a); they read or edit the AoS State they are given and never advance a game.
 *
 * Not provided on the host: TickFlames, TickBombs, AgentBombChainReversion, ResolveBombCollision, MoveBombsForward
 * (declared but never defined in the reference either) - those ARE the tick and exist only as device code
 * (pomcpp_b200/csrc/pom_core.cuh) behind bboard::Step / pom_batch_step; PrintDependency* (console output) is out of scope.
 */
#ifndef STEP_UTILITY_H
#define STEP_UTILITY_H

#include "bboard.hpp"

namespace bboard::util
{

Position DesiredPosition(int x, int y, Move m);                       /* step_utility.hpp:16-23 */
Position OriginPosition(int x, int y, Move m);                        /* :25-29 */
Position DesiredPosition(const Bomb b);                               /* :31-35 */
void FillPositions(State* s, Position p[AGENT_COUNT]);                /* :47-51 */
void FillDestPos(State* s, Move m[AGENT_COUNT], Position p[AGENT_COUNT]);   /* :53-59 */
void FillBombDestPos(State* s, Position p[MAX_BOMBS]);                /* :61-65 */
void FixSwitchMove(State* s, Position desiredPositions[AGENT_COUNT]); /* :67-73 */
int  ResolveDependencies(State* s, Position des[AGENT_COUNT], int dependency[AGENT_COUNT], int chain[AGENT_COUNT]);   /* :75-80 */
void ConsumePowerup(State& state, int agentID, int powerUp);          /* :100-106 */
bool HasDPCollision(const State& state, Position dp[AGENT_COUNT], int agentID);   /* :124-130 */
bool HasBombCollision(const State& state, const Bomb& b, int index = 0);          /* :132-139 */
void ResetBombFlags(State& state);                                    /* :153-157 */

inline bool IsOutOfBounds(const Position& pos)                        /* :159-165 */
{
    return pos.x < 0 || pos.y < 0 || pos.x >= BOARD_SIZE || pos.y >= BOARD_SIZE;
}
inline bool IsOutOfBounds(const int& x, const int& y)                 /* :167-173 */
{
    return x < 0 || y < 0 || x >= BOARD_SIZE || y >= BOARD_SIZE;
}

}

#endif
