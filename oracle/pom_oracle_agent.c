/*
 * pom_oracle_agent.c — TEST INFRASTRUCTURE ONLY (the parity checker, never the product).
 *
 * Plain-C restatement of the reference's heuristic agent: agents::SimpleAgent
 * (src/agents/simple_agent.cpp:12-129) and the bboard::strategy helpers it calls
 * (src/bboard/strategy.cpp:17-340, include/strategy.hpp:134-185), on the AoS `pom_state`.
 *
 * Differences from the reference, both at the boundary only:
 *   * the agent's private std::mt19937_64 (seeded from std::random_device, simple_agent.cpp:17-22)
 *     is replaced by ONE caller-supplied draw d in 0..4 per act() — every path of _Decide consumes at
 *     most one intDist(rng) (simple_agent.cpp:48,86,125); the compiled reference is driven with the same
 *     draw by re-seeding its engine before each act (oracle/ref_shim.cpp ref_simple_act);
 *   * the agent object is taken as zero-initialised memory (the reference leaves moveQueue.queue and
 *     recentPositions.queue indeterminate and READS unwritten slots: simple_agent.cpp:28 with count==2,
 *     :48/:125 with count==1 — "defect D6"; zero is what the shim's placement-new on zeroed memory gives).
 *
 * Parity pinning: tests/test_oracle.py checks this file against the reference's [strategy] known-answer
 * tests (unit_test/bboard/strategy_test.cpp) and, where /root/reference exists, move-by-move against the
 * compiled reference on full SimpleAgent games; tests/golden/simple_agent.npz holds such games.
 */
#include <string.h>
#include <limits.h>
#include "pom_oracle.h"

#define BS POM_BOARD_SIZE
#define NB POM_MAX_BOMBS

typedef struct { int x, y; } pos_t;

static int oob(int x, int y) { return x < 0 || y < 0 || x >= BS || y >= BS; }
static int is_wood(int c)     { return (c >> 8) == 2; }
static int is_powerup(int c)  { return c > 5 && c < 9; }
static int is_walkable(int c) { return is_powerup(c) || c == 0; }

static pos_t desired_pos(int x, int y, int move)                     /* step_utility.cpp:9-31 */
{
    pos_t p = { x, y };
    if (move == POM_MOVE_UP) p.y--;
    else if (move == POM_MOVE_DOWN) p.y++;
    else if (move == POM_MOVE_LEFT) p.x--;
    else if (move == POM_MOVE_RIGHT) p.x++;
    return p;
}

/* ---- RMap (strategy.hpp:29-48, strategy.cpp:17-35): low half = distance, high half = predecessor index */
typedef struct { int map[BS][BS]; pos_t source; } rmap_t;

static int  r_dist(const rmap_t* r, int x, int y) { return r->map[y][x] & 0xFFFF; }
static int  r_pred(const rmap_t* r, int x, int y) { return r->map[y][x] >> 16; }
static void r_set_dist(rmap_t* r, int x, int y, int d) { r->map[y][x] = (r->map[y][x] & ~0xFFFF) + d; }
static void r_set_pred(rmap_t* r, int x, int y, int xp, int yp) { r->map[y][x] = (r->map[y][x] & 0xFFFF) + ((xp + BS * yp) << 16); }

typedef struct { pos_t q[BS * BS]; int index, count; } bfsq_t;

static void try_add(const pom_state* s, bfsq_t* q, rmap_t* r, pos_t c, int cx, int cy)   /* strategy.cpp:37-57 */
{
    /* the reference loads board[cy][cx] before the bounds test and never uses it when out of bounds */
    if (oob(cx, cy)) return;
    int dist = r_dist(r, c.x, c.y);
    int item = s->board[cy][cx];
    if (r_dist(r, cx, cy) == 0 && (is_walkable(item) || item >= POM_ITEM_AGENT0)) {
        r_set_pred(r, cx, cy, c.x, c.y);
        r_set_dist(r, cx, cy, dist + 1);
        if (item < POM_ITEM_AGENT0) {               /* paths END at agent cells */
            q->q[(q->index + q->count) % (BS * BS)] = (pos_t){ cx, cy };
            q->count++;
        }
    }
}

static void fill_rmap(const pom_state* s, rmap_t* r, int id)          /* strategy.cpp:58-95 */
{
    memset(r->map, 0, sizeof r->map);
    int x = s->agents[id].x, y = s->agents[id].y;
    r->source = (pos_t){ x, y };
    bfsq_t q; q.index = 0; q.count = 0;
    r_set_dist(r, x, y, 0);
    q.q[0] = (pos_t){ x, y }; q.count = 1;
    while (q.count != 0) {
        pos_t c = q.q[q.index % (BS * BS)];
        q.index = (q.index + 1) % (BS * BS);
        q.count--;
        /* (RMapInfo `result` is computed by the reference but read by nobody) */
        if (c.x != x || c.y + 1 != y) try_add(s, &q, r, c, c.x, c.y + 1);
        if (c.x != x || c.y - 1 != y) try_add(s, &q, r, c, c.x, c.y - 1);
        if (c.x + 1 != x || c.y != y) try_add(s, &q, r, c, c.x + 1, c.y);
        if (c.x - 1 != x || c.y != y) try_add(s, &q, r, c, c.x - 1, c.y);
    }
}

static int move_towards_position(const rmap_t* r, pos_t position)     /* strategy.cpp:101-124 */
{
    pos_t curr = position;
    for (int guard = 0; guard < 4 * BS * BS; guard++) {
        int idx = r_pred(r, curr.x, curr.y);
        int y = idx / BS, x = idx % BS;
        if (x == r->source.x && y == r->source.y) {
            if (curr.x > r->source.x) return POM_MOVE_RIGHT;
            if (curr.x < r->source.x) return POM_MOVE_LEFT;
            if (curr.y > r->source.y) return POM_MOVE_DOWN;
            if (curr.y < r->source.y) return POM_MOVE_UP;
        } else if (r_dist(r, curr.x, curr.y) == 0) {
            return POM_MOVE_IDLE;
        }
        curr = (pos_t){ x, y };
    }
    return -1;      /* position == source == (0,0): the reference spins forever; no caller can produce it */
}

static int in_bomb_range(int x, int y, int s, pos_t pos)              /* strategy.hpp:163-168 */
{
    return (pos.y == y && (x - s <= pos.x && pos.x <= x + s)) || (pos.x == x && (y - s <= pos.y && pos.y <= y + s));
}

static int is_in_danger(const pom_state* s, int x, int y)             /* strategy.cpp:225-246 */
{
    int minTime = INT_MAX;
    for (int i = 0; i < s->bombs_count; i++) {
        int b = s->bombs[(s->bombs_index + i) % NB];
        if (in_bomb_range(b & 0xF, (b >> 4) & 0xF, (b >> 12) & 0xF, (pos_t){ x, y })) {
            int t = (b >> 16) & 0xF;
            if (t < minTime) minTime = t;
        }
    }
    return minTime == INT_MAX ? 0 : minTime;
}

static int safe_condition(int danger, int min) { return danger == 0 || danger >= min; }   /* strategy.cpp:190-193 */

static int check_pos(const pom_state* s, int x, int y)                /* strategy.cpp:185-188 */
{
    return !oob(x, y) && is_walkable(s->board[y][x]);
}

static int move_towards_safe_place(const pom_state* s, const rmap_t* r, int radius)   /* strategy.cpp:126-144 */
{
    int ox = r->source.x, oy = r->source.y;
    for (int y = oy - radius; y < radius; y++) {          /* sic: the upper bounds are `radius`, not origin + radius */
        for (int x = ox - radius; x < radius; x++) {
            int dx = x - ox, dy = y - oy;
            if (oob(x, y) || (dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy) > radius) continue;
            if (r_dist(r, x, y) != 0 && safe_condition(is_in_danger(s, x, y), 2))
                return move_towards_position(r, (pos_t){ x, y });
        }
    }
    return POM_MOVE_IDLE;
}

static int move_towards_enemy(const pom_state* s, const rmap_t* r, int radius)        /* strategy.cpp:165-183 */
{
    pos_t a = r->source;
    for (int i = 0; i < POM_AGENT_COUNT; i++) {
        const pom_agent* inf = &s->agents[i];
        if ((inf->x == a.x && inf->y == a.y) || inf->dead) continue;
        int dx = inf->x - a.x, dy = inf->y - a.y;
        if ((dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy) > radius) continue;
        return move_towards_position(r, (pos_t){ inf->x, inf->y });
    }
    return POM_MOVE_IDLE;
}

static int move_towards_powerup(const pom_state* s, const rmap_t* r, int radius)      /* strategy.cpp:146-163 */
{
    pos_t a = r->source;
    for (int y = a.y - radius; y <= a.y + radius; y++) {
        for (int x = a.x - radius; x <= a.x + radius; x++) {
            int dx = x - a.x, dy = y - a.y;
            if (oob(x, y) || (dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy) > radius) continue;
            if (is_powerup(s->board[y][x])) return move_towards_position(r, (pos_t){ x, y });
        }
    }
    return POM_MOVE_IDLE;
}

static int is_adjacent_enemy(const pom_state* s, int id, int distance)                /* strategy.cpp:296-312 */
{
    const pom_agent* a = &s->agents[id];
    for (int i = 0; i < POM_AGENT_COUNT; i++) {
        if (i == id || s->agents[i].dead) continue;
        int dx = s->agents[i].x - a->x, dy = s->agents[i].y - a->y;
        if ((dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy) <= distance) return 1;
    }
    return 0;
}

static int is_adjacent_item(const pom_state* s, int id, int distance, int item)       /* strategy.cpp:314-338 */
{
    int ox = s->agents[id].x, oy = s->agents[id].y;
    for (int y = oy - distance; y <= oy + distance; y++) {
        for (int x = ox - distance; x <= ox + distance; x++) {
            int dx = x - ox, dy = y - oy;
            if (oob(x, y) || (dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy) > distance) continue;
            if (is_wood(item) && is_wood(s->board[y][x])) return 1;
            if (s->board[y][x] == item) return 1;
        }
    }
    return 0;
}

/* ---- the agent's two small FixedQueues, unpacked (bboard.hpp:115-188) */
typedef struct {
    int   mq[4]; int mq_count;                 /* moveQueue (index is never moved: only AddElem/RemoveAt/count=0) */
    pos_t rp[4]; int rp_index, rp_count;       /* recentPositions */
} agent_t;

static int nib(int v) { return v & 15; }
static int unnib(int v) { return v == 15 ? -1 : v; }

static void agent_unpack(const pom_simple_agent* p, agent_t* a)
{
    for (int k = 0; k < 4; k++) {
        a->rp[k].x = unnib(p->recent[k] & 15);
        a->rp[k].y = unnib(p->recent[k] >> 4);
        a->mq[k] = (p->move_queue >> (3 * k)) & 7;
    }
    a->rp_index = p->rp_index; a->rp_count = p->rp_count; a->mq_count = 0;
}

static void agent_pack(const agent_t* a, pom_simple_agent* p)
{
    p->move_queue = 0;
    for (int k = 0; k < 4; k++) {
        p->recent[k] = (uint8_t)(nib(a->rp[k].x) | (nib(a->rp[k].y) << 4));
        p->move_queue |= (uint16_t)((a->mq[k] & 7) << (3 * k));
    }
    p->rp_index = (uint8_t)a->rp_index; p->rp_count = (uint8_t)a->rp_count;
}

static void safe_directions(const pom_state* s, agent_t* a, int x, int y)             /* strategy.cpp:194-219 */
{
    static const int order[4] = { POM_MOVE_RIGHT, POM_MOVE_LEFT, POM_MOVE_DOWN, POM_MOVE_UP };
    for (int k = 0; k < 4; k++) {
        pos_t p = desired_pos(x, y, order[k]);
        int d = is_in_danger(s, p.x, p.y);
        if (check_pos(s, p.x, p.y) && safe_condition(d, 2)) a->mq[a->mq_count++ % 4] = order[k];
    }
}

static void sort_directions(agent_t* a, int x, int y)                                 /* strategy.hpp:134-158 */
{
    int moves = a->mq_count, totalRemoves = 0;
    for (int i = 0; i < moves && totalRemoves < 4; i++) {
        pos_t pos = desired_pos(x, y, a->mq[i % 4]);
        for (int j = 0; j < a->rp_count; j++) {
            pos_t rp = a->rp[(a->rp_index + j) % 4];
            if (pos.x == rp.x && pos.y == rp.y) {
                for (int k = i + 1; k < a->mq_count; k++) a->mq[(k + 3) % 4] = a->mq[k % 4];   /* RemoveAt(i) */
                a->mq_count--;
                int e = a->mq[i % 4];                                                       /* AddElem(q[i]): the NEXT element (sic) */
                a->mq[a->mq_count % 4] = e;
                a->mq_count++;
                i--;
                totalRemoves++;
                break;
            }
        }
    }
}

static int has_rp_loop(const agent_t* a)                                              /* simple_agent.cpp:24-35 */
{
    for (int i = 0; i < a->rp_count / 2; i++) {
        pos_t p = a->rp[(a->rp_index + i) % 4], q = a->rp[(a->rp_index + i + 2) % 4];
        if (!(p.x == q.x && p.y == q.y)) return 0;
    }
    return 1;
}

static int pick_safe_direction(const pom_state* s, agent_t* a, int id, int draw)      /* simple_agent.cpp:37-49,113-126 */
{
    const pom_agent* ag = &s->agents[id];
    a->mq_count = 0;
    safe_directions(s, a, ag->x, ag->y);
    sort_directions(a, ag->x, ag->y);
    if (a->mq_count == 0) return POM_MOVE_IDLE;
    return a->mq[draw % 2];
}

static int decide(const pom_state* s, agent_t* me, int id, int draw)                  /* simple_agent.cpp:52-127 */
{
    const pom_agent* a = &s->agents[id];
    rmap_t r;
    fill_rmap(s, &r, id);
    int danger = is_in_danger(s, a->x, a->y);
    if (danger > 0) {
        int m = move_towards_safe_place(s, &r, danger);
        pos_t p = desired_pos(a->x, a->y, m);
        if (!oob(p.x, p.y) && is_walkable(s->board[p.y][p.x]) && safe_condition(is_in_danger(s, p.x, p.y), 2)) return m;
        return pick_safe_direction(s, me, id, draw);
    }
    if (a->bombCount < a->maxBombCount) {
        if (is_adjacent_enemy(s, id, 1)) return POM_MOVE_BOMB;
        if (is_adjacent_enemy(s, id, 7) && has_rp_loop(me)) return draw % 4;
        if (is_adjacent_enemy(s, id, 7)) {
            int m = move_towards_enemy(s, &r, 7);
            pos_t p = desired_pos(a->x, a->y, m);
            if (!oob(p.x, p.y) && is_walkable(s->board[p.y][p.x]) && safe_condition(is_in_danger(s, p.x, p.y), 5)) return m;
        }
        if (is_adjacent_item(s, id, 1, POM_ITEM_WOOD)) return POM_MOVE_BOMB;
    }
    return pick_safe_direction(s, me, id, draw);
}

int pom_oracle_simple_act(const pom_state* s, int id, pom_simple_agent* st, int draw)  /* SimpleAgent::act, simple_agent.cpp:128-141 */
{
    agent_t me;
    agent_unpack(st, &me);
    const pom_agent* a = &s->agents[id];
    int m = decide(s, &me, id, draw);
    pos_t p = desired_pos(a->x, a->y, m);
    if (me.rp_count == 4) { me.rp_index = (me.rp_index + 1) % 4; me.rp_count--; }
    me.rp[(me.rp_index + me.rp_count) % 4] = p;
    me.rp_count++;
    agent_pack(&me, st);
    return m;
}

/* Environment::Step's collection loop (environment.cpp:137-146) for a batch: act() for every live agent in
 * `agent_mask` of every running env, draw a = lane a of pom_oracle_rng_moves(seed, env, tick, 5); the moves of
 * the other agents are left as the caller wrote them.  A dead agent's entry becomes IDLE (the reference
 * leaves it uninitialised). */
void pom_oracle_simple_moves_batch(const pom_state* S, const uint8_t* status, long n, pom_simple_agent* A,
                                   uint64_t seed, uint64_t env0, uint32_t tick, unsigned agent_mask, uint8_t* moves)
{
    for (long e = 0; e < n; e++) {
        if (status && (status[e] & (POM_STATUS_DONE | POM_STATUS_INVALID))) continue;
        uint32_t d = pom_oracle_rng_moves(seed, env0 + (uint64_t)e, tick, 5);
        for (int a = 0; a < 4; a++) {
            if (!((agent_mask >> a) & 1)) continue;
            if (S[e].agents[a].dead) { moves[4 * e + a] = POM_MOVE_IDLE; continue; }
            moves[4 * e + a] = (uint8_t)pom_oracle_simple_act(&S[e], a, &A[4 * e + a], (int)((d >> (8 * a)) & 0xFF));
        }
    }
}

/* strategy helpers exposed for the reference's [strategy] known-answer tests (unit_test/bboard/strategy_test.cpp) */
int pom_oracle_is_adjacent_enemy(const pom_state* s, int id, int distance) { return is_adjacent_enemy(s, id, distance); }
int pom_oracle_is_in_danger(const pom_state* s, int x, int y) { return is_in_danger(s, x, y); }
void pom_oracle_fill_rmap(const pom_state* s, int id, int32_t map_out[POM_BOARD_CELLS])
{
    rmap_t r; fill_rmap(s, &r, id); memcpy(map_out, r.map, sizeof r.map);
}
int pom_oracle_move_towards(const pom_state* s, int id, int kind, int a, int b)
{
    rmap_t r; fill_rmap(s, &r, id);
    if (kind == 0) return move_towards_position(&r, (pos_t){ a, b });
    if (kind == 1) return move_towards_powerup(s, &r, a);
    if (kind == 2) return move_towards_enemy(s, &r, a);
    return move_towards_safe_place(s, &r, a);
}
