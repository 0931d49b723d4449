/*
 * pom_strategy.cpp — host implementations of include/strategy.hpp (reference: src/bboard/strategy.cpp:17-340).
 * Agent code calls these on the State it receives in Agent::act; they only read it.  Results are the reference's cell
 * for cell: the BFS visits neighbours in the order down, up, right, left and stops at agent cells, the scans run
 * y-outer / x-inner, MoveTowardsSafePlace keeps the reference's loop bounds (y < radius, x < radius).
 * Checked against the compiled reference and its [strategy] known-answer tests by tests/test_user_agent.py.
 */
#include <climits>
#include <cstdlib>

#include "strategy.hpp"

namespace bboard::strategy
{

namespace
{

inline int manhattan(int x0, int y0, int x1, int y1) { return std::abs(x0 - x1) + std::abs(y0 - y1); }

/* flat FIFO over board cells; a cell enters at most once */
struct CellQueue
{
    short cells[BOARD_SIZE * BOARD_SIZE];
    int head = 0, tail = 0;
    bool empty() const { return head == tail; }
    void push(int x, int y) { cells[tail++] = short(x + BOARD_SIZE * y); }
    Position pop() { const int c = cells[head++]; return { c % BOARD_SIZE, c / BOARD_SIZE }; }
};

}

/* strategy.cpp:37-95.  Distance 0 marks "not reached" - the source keeps 0, so the search never re-enters it only
 * because its four neighbours are skipped explicitly when they ARE the source.  Agent cells are labelled but not
 * expanded: paths end at agents. */
void FillRMap(const State& s, RMap& r, int agentID)
{
    for(int y = 0; y < BOARD_SIZE; y++)
        for(int x = 0; x < BOARD_SIZE; x++) r.map[y][x] = 0;
    const int sx = s.agents[agentID].x, sy = s.agents[agentID].y;
    r.source = { sx, sy };
    r.info = 0;
    CellQueue open;
    open.push(sx, sy);
    static const int step[4][2] = { {0, 1}, {0, -1}, {1, 0}, {-1, 0} };
    while(!open.empty())
    {
        const Position c = open.pop();
        const int next = r.GetDistance(c.x, c.y) + 1;
        for(const auto& d : step)
        {
            const int nx = c.x + d[0], ny = c.y + d[1];
            if((nx == sx && ny == sy) || util::IsOutOfBounds(nx, ny)) continue;
            const int item = s.board[ny][nx];
            if(r.GetDistance(nx, ny) != 0 || !(IS_WALKABLE(item) || item >= Item::AGENT0)) continue;
            r.SetPredecessor(nx, ny, c.x, c.y);
            r.SetDistance(nx, ny, next);
            if(item < Item::AGENT0) open.push(nx, ny);
        }
    }
}

/* strategy.cpp:101-124: walk the predecessor chain back from `position` to the cell next to the source */
Move MoveTowardsPosition(const RMap& r, const Position& position)
{
    Position at = position;
    for(;;)
    {
        const int pred = r.GetPredecessor(at.x, at.y);
        const Position before = { pred % BOARD_SIZE, pred / BOARD_SIZE };
        if(before == r.source)
        {
            if(at.x > r.source.x) return Move::RIGHT;
            if(at.x < r.source.x) return Move::LEFT;
            if(at.y > r.source.y) return Move::DOWN;
            if(at.y < r.source.y) return Move::UP;
        }
        else if(r.GetDistance(at.x, at.y) == 0)
        {
            return Move::IDLE;                                        /* not reachable */
        }
        at = before;
    }
}

bool _safe_condition(int danger, int min) { return danger == 0 || danger >= min; }   /* :190-193 */

Move MoveTowardsSafePlace(const State& state, const RMap& r, int radius)   /* :126-144 */
{
    const Position o = r.source;
    for(int y = o.y - radius; y < radius; y++)                        /* sic: not o.y + radius */
    {
        for(int x = o.x - radius; x < radius; x++)
        {
            if(util::IsOutOfBounds(x, y) || manhattan(x, y, o.x, o.y) > radius) continue;
            if(r.GetDistance(x, y) != 0 && _safe_condition(IsInDanger(state, x, y)))
                return MoveTowardsPosition(r, { x, y });
        }
    }
    return Move::IDLE;
}

Move MoveTowardsPowerup(const State& state, const RMap& r, int radius)     /* :146-163 */
{
    const Position o = r.source;
    for(int y = o.y - radius; y <= o.y + radius; y++)
    {
        for(int x = o.x - radius; x <= o.x + radius; x++)
        {
            if(util::IsOutOfBounds(x, y) || manhattan(x, y, o.x, o.y) > radius) continue;
            if(IS_POWERUP(state.board[y][x])) return MoveTowardsPosition(r, { x, y });
        }
    }
    return Move::IDLE;
}

Move MoveTowardsEnemy(const State& state, const RMap& r, int radius)       /* :165-183 */
{
    const Position o = r.source;
    for(int i = 0; i < AGENT_COUNT; i++)
    {
        const AgentInfo& e = state.agents[i];
        const bool self = e.x == o.x && e.y == o.y;                   /* identified by position, as in the reference */
        if(self || e.dead || manhattan(e.x, e.y, o.x, o.y) > radius) continue;
        return MoveTowardsPosition(r, { e.x, e.y });
    }
    return Move::IDLE;
}

/* :194-219: the four neighbours in the order right, left, down, up that can be entered and are not about to blow */
void SafeDirections(const State& state, FixedQueue<Move, MOVE_COUNT>& q, int x, int y)
{
    static const Move order[4] = { Move::RIGHT, Move::LEFT, Move::DOWN, Move::UP };
    for(Move m : order)
    {
        const Position p = util::DesiredPosition(x, y, m);
        const int danger = IsInDanger(state, p.x, p.y);
        if(!util::IsOutOfBounds(p.x, p.y) && IS_WALKABLE(state.board[p.y][p.x]) && _safe_condition(danger)) q.AddElem(m);
    }
}

int IsInDanger(const State& state, int agentID)                       /* :221-224 */
{
    return IsInDanger(state, state.agents[agentID].x, state.agents[agentID].y);
}

/* :225-246: shortest timer among the bombs whose cross covers (x, y); walls do not shield; 0 = safe */
int IsInDanger(const State& state, int x, int y)
{
    int soonest = INT_MAX;
    for(int k = 0; k < state.bombs.count; k++)
    {
        const Bomb b = state.bombs[k];
        if(IsInBombRange(BMB_POS_X(b), BMB_POS_Y(b), BMB_STRENGTH(b), { x, y }) && BMB_TIME(b) < soonest) soonest = BMB_TIME(b);
    }
    return soonest == INT_MAX ? 0 : soonest;
}

bool IsAdjacentEnemy(const State& state, int agentID, int distance)   /* :296-312 */
{
    const AgentInfo& me = state.agents[agentID];
    for(int i = 0; i < AGENT_COUNT; i++)
    {
        if(i != agentID && !state.agents[i].dead && manhattan(state.agents[i].x, state.agents[i].y, me.x, me.y) <= distance) return true;
    }
    return false;
}

bool IsAdjacentItem(const State& state, int agentID, int distance, Item item)   /* :314-338: any wood matches WOOD */
{
    const AgentInfo& me = state.agents[agentID];
    for(int y = me.y - distance; y <= me.y + distance; y++)
    {
        for(int x = me.x - distance; x <= me.x + distance; x++)
        {
            if(util::IsOutOfBounds(x, y) || manhattan(x, y, me.x, me.y) > distance) continue;
            const int c = state.board[y][x];
            if((IS_WOOD(item) && IS_WOOD(c)) || c == item) return true;
        }
    }
    return false;
}

}
