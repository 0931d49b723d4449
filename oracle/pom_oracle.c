/*
 * pom_oracle.c — TEST INFRASTRUCTURE ONLY (see pom_oracle.h).
 *
 * Plain-C restatement of the reference step path.  Every function cites the reference
 * file:line it follows (paths relative to the dist1ll/pomcpp tree).  Recursion is kept
 * where the reference recurses (SpawnFlame <-> SpawnFlameItem <-> ExplodeBombAt), so the
 * order-dependent quirks fall out of the same control flow; the
 * CUDA path uses an explicit stack instead, which makes the two independent statements
 * of the same semantics.
 *
 * Canonical handling of the reference's undefined spots:
 *   D1: the dependency walk stops when the roots are exhausted (flag D1_UNREACHABLE).
 *   D3: a kicker on a BOMB cell without queue entry moves but sets no direction (flag).
 *   D4 / flame overflow: flagged; the ring wraps exactly as FixedQueue would.
 *   D5: unbounded AgentBombChainReversion recursion (found by this project's differential runs: the
 *       compiled reference hangs at -O3 and overflows the stack at -O0): flagged, chain stopped.
 * An env whose tick raised D3/D4/D5/overflow/bad-move is marked POM_STATUS_INVALID and frozen.
 *
 * Also here: InitBoardItems with libstdc++'s mt19937_64 / uniform_int_distribution restated, the shared counter RNG, the
 * multi-threaded "port" CPU baseline, and pom_oracle_fog (the DEFINITION of the fogged view, which the reference only
 * declares).  The heuristic agent (SimpleAgent + strategy) is restated in pom_oracle_agent.c.
 */
#include "pom_oracle.h"

#include <string.h>
#include <stdlib.h>
#include <pthread.h>
#include <time.h>

#define NB POM_MAX_BOMBS
#define BS POM_BOARD_SIZE

/* ---- Item predicates: include/bboard.hpp:73-109 ---- */
static int is_wood(int c)      { return (c >> 8) == 2; }
static int is_powerup(int c)   { return c > 5 && c < 9; }
static int is_walkable(int c)  { return is_powerup(c) || c == 0; }
static int is_flame(int c)     { return (c >> 16) == 4; }
static int is_agent(int c)     { return c >= (1 << 24); }
static int is_static_block(int c) { return is_wood(c) || is_powerup(c) || c == 1; }
static int flame_id(int c)     { return (c & 0xFFFF) >> 3; }

/* ---- Bomb word accessors: include/bboard.hpp:266-335 (identical add/sub arithmetic) ---- */
static int b_pos(int b)  { return b & 0xFF; }
static int b_x(int b)    { return b & 0xF; }
static int b_y(int b)    { return (b & 0xF0) >> 4; }
static int b_id(int b)   { return (b & 0xF00) >> 8; }
static int b_str(int b)  { return (b & 0xF000) >> 12; }
static int b_time(int b) { return (b & 0xF0000) >> 16; }
static int b_dir(int b)  { return (b & 0xF00000) >> 20; }
static void b_set_pos(int* b, int x, int y) { *b = (*b & ~0xF & ~0xF0) + x + (y << 4); }
static void b_set_id(int* b, int id)        { *b = (*b & ~0xF00) + (id << 8); }
static void b_set_str(int* b, int s)        { *b = (*b & ~0xF000) + (s << 12); }
static void b_set_time(int* b, int t)       { *b = (*b & ~0xF0000) + (t << 16); }
static void b_set_dir(int* b, int d)        { *b = (*b & ~0xF00000) + (d << 20); }
static void b_set_moved(int* b, int m)      { *b = (*b & ~0xF000000) + (m << 24); }

/* ---- FixedQueue: include/bboard.hpp:115-188 ---- */
static int* bomb_at(pom_state* s, int i) { return &s->bombs[(s->bombs_index + i) % NB]; }
static pom_flame* flame_at(pom_state* s, int i) { return &s->flames[(s->flames_index + i) % NB]; }

static void bombs_remove_at(pom_state* s, int removeAt)           /* :151-160 */
{
    for (int i = removeAt + 1; i < s->bombs_count; i++) {
        int t = (s->bombs_index + i) % NB;
        s->bombs[(t - 1 + NB) % NB] = s->bombs[t];
    }
    s->bombs_count--;
}

static int oob(int x, int y) { return x < 0 || y < 0 || x >= BS || y >= BS; }   /* step_utility.hpp:163-166 */

typedef struct { int x, y; } pos_t;

static pos_t desired_pos(int x, int y, int move)                    /* step_utility.cpp:9-31 */
{
    pos_t p = { x, y };
    if (move == POM_MOVE_UP) p.y -= 1;
    else if (move == POM_MOVE_DOWN) p.y += 1;
    else if (move == POM_MOVE_LEFT) p.x -= 1;
    else if (move == POM_MOVE_RIGHT) p.x += 1;
    return p;
}

static pos_t origin_pos(int x, int y, int move)                     /* step_utility.cpp:33-55 */
{
    pos_t p = { x, y };
    if (move == POM_MOVE_DOWN) p.y -= 1;
    else if (move == POM_MOVE_UP) p.y += 1;
    else if (move == POM_MOVE_RIGHT) p.x -= 1;
    else if (move == POM_MOVE_LEFT) p.x += 1;
    return p;
}

static pos_t bomb_desired(int b) { return desired_pos(b_x(b), b_y(b), b_dir(b)); }  /* :57-60 */

/* ---- State methods ---- */
void pom_oracle_zero_state(pom_state* s)
{
    memset(s, 0, sizeof(*s));
    s->aliveAgents = POM_AGENT_COUNT;
    for (int i = 0; i < 4; i++) {
        s->agents[i].maxBombCount = 1;
        s->agents[i].bombStrength = POM_BOMB_DEFAULT_STRENGTH;
    }
    for (int i = 0; i < NB; i++) s->flames[i].timeLeft = POM_FLAME_LIFETIME;
}

void pom_oracle_kill(pom_state* s, int id)                           /* bboard.hpp:474-481 */
{
    if (!s->agents[id].dead) {
        s->agents[id].dead = 1;
        s->aliveAgents--;
    }
}

void pom_oracle_put_agent(pom_state* s, int x, int y, int id)        /* bboard.cpp:313-320 */
{
    s->board[y][x] = POM_ITEM_AGENT0 + id;
    s->agents[id].x = x;
    s->agents[id].y = y;
}

void pom_oracle_put_agents_in_corners(pom_state* s, int a0, int a1, int a2, int a3) /* bboard.cpp:322-333 */
{
    s->board[0][0] = POM_ITEM_AGENT0 + a0;
    s->board[0][BS - 1] = POM_ITEM_AGENT0 + a1;
    s->board[BS - 1][BS - 1] = POM_ITEM_AGENT0 + a2;
    s->board[BS - 1][0] = POM_ITEM_AGENT0 + a3;
    s->agents[a1].x = s->agents[a2].x = BS - 1;
    s->agents[a2].y = s->agents[a3].y = BS - 1;
}

static int has_bomb(pom_state* s, int x, int y)                      /* bboard.cpp:265-275 */
{
    for (int i = 0; i < s->bombs_count; i++) {
        int b = *bomb_at(s, i);
        if (b_x(b) == x && b_y(b) == y) return 1;
    }
    return 0;
}

static int bomb_index(pom_state* s, int x, int y)                    /* bboard.cpp:277-287, 301-311 */
{
    for (int i = 0; i < s->bombs_count; i++) {
        int b = *bomb_at(s, i);
        if (b_x(b) == x && b_y(b) == y) return i;
    }
    return -1;
}

static int get_agent(pom_state* s, int x, int y)                     /* bboard.cpp:289-299 */
{
    for (int i = 0; i < 4; i++) {
        if (!s->agents[i].dead && s->agents[i].x == x && s->agents[i].y == y) return i;
    }
    return -1;
}

void pom_oracle_plant_bomb(pom_state* s, int x, int y, int id, int lifeTime, int setItem) /* bboard.cpp:125-146 */
{
    if (s->agents[id].bombCount >= s->agents[id].maxBombCount) return;
    int* b = bomb_at(s, s->bombs_count);        /* NextPos(): stale direction/moved bits survive (Q4) */
    b_set_id(b, id);
    b_set_pos(b, x, y);
    b_set_str(b, s->agents[id].bombStrength);
    b_set_time(b, lifeTime);
    if (setItem) s->board[y][x] = POM_ITEM_BOMB;
    s->agents[id].bombCount++;
    s->bombs_count++;
}

static int flag_item(int pwp)                                        /* bboard.cpp:182-189 */
{
    if (pwp == 1) return POM_ITEM_EXTRABOMB;
    if (pwp == 2) return POM_ITEM_INCRRANGE;
    if (pwp == 3) return POM_ITEM_KICK;
    return POM_ITEM_PASSAGE;
}

static void pop_flame(pom_state* s)                                  /* bboard.cpp:148-180 */
{
    pom_flame* f = flame_at(s, 0);
    int st = f->strength, x = f->x, y = f->y;
    int sig = (x + BS * y) & 0xFFFF;
    for (int i = -st; i <= st; i++) {
        if (!oob(x + i, y) && is_flame(s->board[y][x + i])) {
            int c = s->board[y][x + i];
            if (flame_id(c) == sig) s->board[y][x + i] = flag_item(c & 3);
        }
        if (!oob(x, y + i) && is_flame(s->board[y + i][x])) {
            int c = s->board[y + i][x];
            if (flame_id(c) == sig) s->board[y + i][x] = flag_item(c & 3);
        }
    }
    s->flames_index = (s->flames_index + 1) % NB;                    /* PopElem bboard.hpp:131-137 */
    s->flames_count--;
}

static void spawn_flame(pom_state* s, int x, int y, int strength, int* flags);

static void explode_bomb_at(pom_state* s, int i, int* flags)         /* bboard.cpp:111-118 */
{
    int b = *bomb_at(s, i);
    spawn_flame(s, b_x(b), b_y(b), s->agents[b_id(b)].bombStrength, flags);
    /* bombs[i] is re-read AFTER the nested explosions: an inner RemoveAt(j<i) has shifted the ring (Q6) */
    s->agents[b_id(*bomb_at(s, i))].bombCount--;
    bombs_remove_at(s, i);
}

static int spawn_flame_item(pom_state* s, int x, int y, int signature, int* flags) /* bboard.cpp:24-57 */
{
    if (s->board[y][x] >= POM_ITEM_AGENT0) pom_oracle_kill(s, s->board[y][x] - POM_ITEM_AGENT0);
    if (s->board[y][x] == POM_ITEM_BOMB || s->board[y][x] >= POM_ITEM_AGENT0) {
        for (int i = 0; i < s->bombs_count; i++) {
            if (b_pos(*bomb_at(s, i)) == (x + (y << 4))) {
                explode_bomb_at(s, i, flags);
                break;
            }
        }
    }
    if (s->board[y][x] != POM_ITEM_RIGID) {
        int old = s->board[y][x];
        int wasWood = is_wood(old);
        s->board[y][x] = POM_ITEM_FLAMES + signature;
        if (wasWood) s->board[y][x] += old & 3;
        return !wasWood;
    }
    return 0;
}

static void spawn_flame(pom_state* s, int x, int y, int strength, int* flags) /* bboard.cpp:198-263 */
{
    if (s->flames_count >= NB) *flags |= POM_ORC_FLAME_OVF;
    pom_flame* f = flame_at(s, s->flames_count);
    f->x = x; f->y = y; f->strength = strength; f->timeLeft = POM_FLAME_LIFETIME;
    int signature = ((x + BS * y) << 3) & 0xFFFF;
    s->flames_count++;
    if (s->board[y][x] >= POM_ITEM_AGENT0) pom_oracle_kill(s, s->board[y][x] - POM_ITEM_AGENT0);
    s->board[y][x] = POM_ITEM_FLAMES + signature;
    for (int i = 1; i <= strength; i++) {
        if (x + i >= BS) break;
        if (!spawn_flame_item(s, x + i, y, signature, flags)) break;
    }
    for (int i = 1; i <= strength; i++) {
        if (x - i < 0) break;
        if (!spawn_flame_item(s, x - i, y, signature, flags)) break;
    }
    for (int i = 1; i <= strength; i++) {
        if (y + i >= BS) break;
        if (!spawn_flame_item(s, x, y + i, signature, flags)) break;
    }
    for (int i = 1; i <= strength; i++) {
        if (y - i < 0) break;
        if (!spawn_flame_item(s, x, y - i, signature, flags)) break;
    }
}

int pom_oracle_spawn_flame(pom_state* s, int x, int y, int strength)
{
    int flags = 0;
    spawn_flame(s, x, y, strength, &flags);
    return flags;
}

static void explode_top_bomb(pom_state* s, int* flags)               /* bboard.cpp:191-196, PopBomb :93-97 */
{
    int c = *bomb_at(s, 0);
    spawn_flame(s, b_x(c), b_y(c), b_str(c), flags);
    s->agents[b_id(*bomb_at(s, 0))].bombCount--;
    s->bombs_index = (s->bombs_index + 1) % NB;
    s->bombs_count--;
}

/* ---- step utilities ---- */
static void tick_flames(pom_state* s)                                /* step_utility.cpp:208-222 */
{
    for (int i = 0; i < s->flames_count; i++) flame_at(s, i)->timeLeft--;
    int n = s->flames_count;
    for (int i = 0; i < n; i++) {
        if (flame_at(s, 0)->timeLeft == 0) pop_flame(s);
    }
}

static void tick_bombs(pom_state* s, int* flags)                     /* step_utility.cpp:224-245 */
{
    for (int i = 0; i < s->bombs_count; i++) *bomb_at(s, i) -= (1 << 16);   /* ReduceBombTimer bboard.hpp:308-311 */
    int n = s->bombs_count;
    for (int i = 0; i < n && s->bombs_count > 0; i++) {
        if (b_time(*bomb_at(s, 0)) == 0) explode_top_bomb(s, flags);
        else break;
    }
}

void pom_oracle_fill_dest_pos(const pom_state* s, const uint8_t moves[4], int pos8[8]) /* step_utility.cpp:138-144 */
{
    for (int i = 0; i < 4; i++) {
        pos_t p = desired_pos(s->agents[i].x, s->agents[i].y, moves[i]);
        pos8[2 * i] = p.x; pos8[2 * i + 1] = p.y;
    }
}

void pom_oracle_fix_switch_move(const pom_state* s, int d[8])        /* step_utility.cpp:154-170 (dead agents NOT skipped, Q1) */
{
    for (int i = 0; i < 4; i++) {
        for (int j = i; j < 4; j++) {
            if (d[2 * i] == s->agents[j].x && d[2 * i + 1] == s->agents[j].y &&
                d[2 * j] == s->agents[i].x && d[2 * j + 1] == s->agents[i].y) {
                d[2 * i] = s->agents[i].x; d[2 * i + 1] = s->agents[i].y;
                d[2 * j] = s->agents[j].x; d[2 * j + 1] = s->agents[j].y;
            }
        }
    }
}

int pom_oracle_resolve_dependencies(const pom_state* s, const int des[8], int dependency[4], int roots[4]) /* step_utility.cpp:172-205 */
{
    int rootCount = 0;
    for (int i = 0; i < 4; i++) { dependency[i] = -1; roots[i] = -1; }
    for (int i = 0; i < 4; i++) {
        if (s->agents[i].dead) { roots[rootCount++] = i; continue; }
        int isRoot = 1;
        for (int j = 0; j < 4; j++) {
            if (i == j || s->agents[j].dead) continue;
            if (des[2 * i] == s->agents[j].x && des[2 * i + 1] == s->agents[j].y) {
                dependency[j] = i;
                isRoot = 0;
                break;
            }
        }
        if (isRoot) roots[rootCount++] = i;
    }
    return rootCount;
}

static int has_dp_collision(const pom_state* s, const int dp[8], int id)      /* step_utility.cpp:264-277 */
{
    for (int i = 0; i < 4; i++) {
        if (id == i || s->agents[i].dead) continue;
        if (dp[2 * id] == dp[2 * i] && dp[2 * id + 1] == dp[2 * i + 1]) return 1;
    }
    return 0;
}

static void consume_powerup(pom_state* s, int id, int item)          /* step_utility.cpp:247-262 */
{
    if (item == POM_ITEM_EXTRABOMB) s->agents[id].maxBombCount++;
    else if (item == POM_ITEM_INCRRANGE) s->agents[id].bombStrength++;
    else if (item == POM_ITEM_KICK) s->agents[id].canKick = 1;
}

static int has_bomb_collision(pom_state* s, int b, int index)        /* step_utility.cpp:279-293 (compares bomb VALUES) */
{
    pos_t t = bomb_desired(b);
    for (int i = index; i < s->bombs_count; i++) {
        int o = *bomb_at(s, i);
        pos_t q = bomb_desired(o);
        if (b != o && q.x == t.x && q.y == t.y) return 1;
    }
    return 0;
}

/* D5: the reference recurses without bound when the chain reaches an agent whose move is IDLE/BOMB
 * (its "origin" is its own cell, where GetAgent finds the agent itself): stack overflow at -O0, endless
 * tail-call loop at -O2/-O3.  Canonical: flag the env and stop the chain there. */
static pos_t chain_reversion_d(pom_state* s, const uint8_t moves[4], const pos_t destBombs[NB], int agentID, int depth, int* flags);

static pos_t chain_reversion(pom_state* s, const uint8_t moves[4], const pos_t destBombs[NB], int agentID, int* flags)
{
    return chain_reversion_d(s, moves, destBombs, agentID, 0, flags);
}

static pos_t chain_reversion_d(pom_state* s, const uint8_t moves[4], const pos_t destBombs[NB], int agentID, int depth, int* flags) /* step_utility.cpp:62-128 */
{
    pom_agent* agent = &s->agents[agentID];
    if (depth >= 64) { *flags |= POM_ORC_D5_REVERT_LOOP; pos_t r0 = { agent->x, agent->y }; return r0; }
    pos_t origin = origin_pos(agent->x, agent->y, moves[agentID]);
    if (!oob(origin.x, origin.y)) {
        int indexOriginAgent = get_agent(s, origin.x, origin.y);
        int bombDestIndex = -1;
        for (int i = 0; i < s->bombs_count; i++) {
            if (destBombs[i].x == origin.x && destBombs[i].y == origin.y) { bombDestIndex = i; break; }
        }
        agent->x = origin.x;
        agent->y = origin.y;
        s->board[origin.y][origin.x] = POM_ITEM_AGENT0 + agentID;
        if (indexOriginAgent == agentID) { *flags |= POM_ORC_D5_REVERT_LOOP; return origin; }
        if (indexOriginAgent != -1) {
            return chain_reversion_d(s, moves, destBombs, indexOriginAgent, depth + 1, flags);
        } else if (bombDestIndex != -1) {
            int* b = bomb_at(s, bombDestIndex);
            pos_t bombDest = destBombs[bombDestIndex];
            pos_t originBomb = origin_pos(bombDest.x, bombDest.y, b_dir(*b));
            if (originBomb.x == bombDest.x && originBomb.y == bombDest.y) {
                s->board[originBomb.y][originBomb.x] = POM_ITEM_AGENT0 + agentID;
                return originBomb;
            }
            int hasAgent = get_agent(s, originBomb.x, originBomb.y);
            b_set_dir(b, 0);
            b_set_pos(b, originBomb.x, originBomb.y);
            s->board[originBomb.y][originBomb.x] = POM_ITEM_BOMB;
            if (hasAgent != -1) return chain_reversion_d(s, moves, destBombs, hasAgent, depth + 1, flags);
            return originBomb;
        }
        return origin;
    }
    pos_t r = { agent->x, agent->y };
    return r;
}

static void resolve_bomb_collision(pom_state* s, const uint8_t moves[4], const pos_t destBombs[NB], int index, int* flags) /* step_utility.cpp:295-329 */
{
    int* b = bomb_at(s, index);
    pos_t t = bomb_desired(*b);
    int hasCollided = 0;
    for (int i = index; i < s->bombs_count; i++) {
        int* o = bomb_at(s, i);
        pos_t q = bomb_desired(*o);
        if (*b != *o && q.x == t.x && q.y == t.y) {
            b_set_dir(o, 0);
            hasCollided = 1;
        }
    }
    if (hasCollided) {
        if (b_dir(*b) != 0) {
            b_set_dir(b, 0);
            int a = get_agent(s, b_x(*b), b_y(*b));
            if (a > -1 && moves[a] != POM_MOVE_IDLE && moves[a] != POM_MOVE_BOMB) {
                chain_reversion(s, moves, destBombs, a, flags);
                s->board[b_y(*b)][b_x(*b)] = POM_ITEM_BOMB;
            }
        }
    }
}

/* ---- bboard::Step, step.cpp:9-284 ---- */
int pom_oracle_step(pom_state* s, const uint8_t moves_in[4])
{
    int flags = 0;
    uint8_t moves[4];
    for (int i = 0; i < 4; i++) {
        moves[i] = moves_in[i];
        if (moves[i] > 5) { moves[i] = 0; flags |= POM_ORC_BAD_MOVE; }
    }

    tick_flames(s);                                                  /* :15 */

    pos_t oldPos[4];
    int dest[8];
    for (int i = 0; i < 4; i++) { oldPos[i].x = s->agents[i].x; oldPos[i].y = s->agents[i].y; }  /* :24 */
    pom_oracle_fill_dest_pos(s, moves, dest);                        /* :25 */
    pom_oracle_fix_switch_move(s, dest);                             /* :26 */

    int dependency[4], roots[4];
    int rootNumber = pom_oracle_resolve_dependencies(s, dest, dependency, roots);   /* :32 */
    int ouroboros = rootNumber == 0;

    int rootIdx = 0;
    int i = rootNumber == 0 ? 0 : roots[0];
    for (int k = 0; k < 4; k++, i = dependency[i]) {                 /* :39 */
        if (i == -1) {
            rootIdx++;
            /* D1: the reference reads roots[]/moves[]/agents[] at -1 here; canonical = stop */
            if (rootIdx > 3 || roots[rootIdx] == -1) { flags |= POM_ORC_D1_UNREACHABLE; break; }
            i = roots[rootIdx];
        }
        int m = moves[i];
        if (s->agents[i].dead || m == POM_MOVE_IDLE) continue;
        if (m == POM_MOVE_BOMB) {                                    /* :52-56 (lifetime 11, ticked to 10 below: Q3) */
            if (s->agents[i].bombCount < s->agents[i].maxBombCount && s->bombs_count >= NB) flags |= POM_ORC_D4_BOMB_OVF;
            pom_oracle_plant_bomb(s, s->agents[i].x, s->agents[i].y, i, POM_BOMB_LIFETIME + 1, 0);
            continue;
        }
        int x = s->agents[i].x, y = s->agents[i].y;
        int dx = dest[2 * i], dy = dest[2 * i + 1];
        if (oob(dx, dy)) continue;                                   /* :63 */
        int item = s->board[dy][dx];
        if (ouroboros) {                                             /* :71-82 */
            for (int j = 0; j < s->bombs_count; j++) {
                int b = *bomb_at(s, j);
                if (b_x(b) == dx && b_y(b) == dy) { item = POM_ITEM_BOMB; break; }
            }
        }
        if (is_flame(item)) {                                        /* :84-99 */
            pom_oracle_kill(s, i);
            if (s->board[y][x] == POM_ITEM_AGENT0 + i)
                s->board[y][x] = has_bomb(s, x, y) ? POM_ITEM_BOMB : POM_ITEM_PASSAGE;
            continue;
        }
        if (has_dp_collision(s, dest, i)) continue;                  /* :100 */
        if (is_powerup(item)) {                                      /* :111-115 */
            consume_powerup(s, i, item);
            item = POM_ITEM_PASSAGE;
        }
        if (item == POM_ITEM_PASSAGE || (ouroboros && item >= POM_ITEM_AGENT0)) {   /* :120-140 */
            if (s->board[y][x] == POM_ITEM_AGENT0 + i)
                s->board[y][x] = has_bomb(s, x, y) ? POM_ITEM_BOMB : POM_ITEM_PASSAGE;
            s->board[dy][dx] = POM_ITEM_AGENT0 + i;
            s->agents[i].x = dx; s->agents[i].y = dy;
        } else if (item == POM_ITEM_BOMB) {                          /* :147-184 (kick / step onto bomb, Q2) */
            s->board[y][x] = has_bomb(s, x, y) ? POM_ITEM_BOMB : POM_ITEM_PASSAGE;
            s->board[dy][dx] = POM_ITEM_AGENT0 + i;
            s->agents[i].x = dx; s->agents[i].y = dy;
            if (s->agents[i].canKick) {
                int bi = bomb_index(s, dx, dy);
                if (bi < 0) flags |= POM_ORC_D3_NULL_BOMB;           /* D3: reference dereferences nullptr */
                else b_set_dir(bomb_at(s, bi), m);
            }
        }
    }

    for (int k = 0; k < s->bombs_count; k++) b_set_moved(bomb_at(s, k), 0);   /* :188, step_utility.cpp:331-337 */

    pos_t bombDest[NB];
    if (s->bombs_count > NB) flags |= POM_ORC_D4_BOMB_OVF;
    for (int k = 0; k < s->bombs_count && k < NB; k++) bombDest[k] = bomb_desired(*bomb_at(s, k));   /* :191-192 */

    for (int k = 0; k < s->bombs_count; k++) {                       /* :195-227 */
        int* b = bomb_at(s, k);
        int bx = b_x(*b), by = b_y(*b);
        pos_t t = bomb_desired(*b);
        if (oob(t.x, t.y) || is_static_block(s->board[t.y][t.x]) || is_agent(s->board[t.y][t.x])) {
            b_set_dir(b, 0);
            int a = get_agent(s, bx, by);
            if (a > -1 && moves[a] != POM_MOVE_IDLE && moves[a] != POM_MOVE_BOMB &&
                !(s->agents[a].x == oldPos[a].x && s->agents[a].y == oldPos[a].y)) {
                chain_reversion(s, moves, bombDest, a, &flags);
                if (get_agent(s, bx, by) == -1) s->board[by][bx] = POM_ITEM_BOMB;
            }
        }
    }

    for (int k = 0; k < s->bombs_count; k++) {                       /* :230-278 */
        int* b = bomb_at(s, k);
        if (b_dir(*b) == 0) {
            if (has_bomb_collision(s, *b, k)) {
                resolve_bomb_collision(s, moves, bombDest, k, &flags);
                continue;
            }
        }
        int bx = b_x(*b), by = b_y(*b);
        pos_t t = bomb_desired(*b);
        /* the reference forms &board[t] before the bounds test but only reads it when in bounds */
        if (!oob(t.x, t.y) && !is_static_block(s->board[t.y][t.x])) {
            if (has_bomb_collision(s, *b, k)) {
                resolve_bomb_collision(s, moves, bombDest, k, &flags);
                continue;
            }
            b_set_pos(b, t.x, t.y);
            if (!has_bomb(s, bx, by) && s->board[by][bx] == POM_ITEM_BOMB) s->board[by][bx] = POM_ITEM_PASSAGE;
            int* tItem = &s->board[t.y][t.x];
            if (is_walkable(*tItem)) *tItem = POM_ITEM_BOMB;
            else if (is_flame(*tItem)) explode_bomb_at(s, bomb_index(s, t.x, t.y), &flags);   /* Q7: ring shrinks inside the loop */
        } else {
            b_set_dir(b, 0);
        }
    }

    tick_bombs(s, &flags);                                           /* :283 */
    return flags;
}

/* POM_STEP_CONTINUE_UNDEFINED (include/pom_batch.h): where the reference dereferences null (D3) or recurses for ever
 * (D5) the restatement already computes the canonical continuation (the kicker moves and no direction is set; the
 * reversion chain stops); with this switch on such a tick no longer takes the env out of the game. */
static int g_invalid_mask = POM_ORC_INVALID_MASK;
void pom_oracle_set_continue_undefined(int on)
{
    g_invalid_mask = on ? (POM_ORC_INVALID_MASK & ~(POM_ORC_D3_NULL_BOMB | POM_ORC_D5_REVERT_LOOP)) : POM_ORC_INVALID_MASK;
}

int pom_oracle_env_step(pom_state* s, uint8_t* status, const uint8_t moves[4])   /* environment.cpp:125-128,149-168 */
{
    if (*status & (POM_STATUS_DONE | POM_STATUS_INVALID)) return 0;   /* invalid envs freeze: the reference would have crashed */
    int flags = pom_oracle_step(s, moves);
    s->timeStep++;
    if (s->aliveAgents == 1) {
        int w = 0;
        for (int i = 0; i < 4; i++) if (!s->agents[i].dead) w = i;
        *status = (uint8_t)((*status & POM_STATUS_INVALID) | POM_STATUS_DONE | (w << POM_STATUS_WINNER_SHIFT));
    }
    if (s->aliveAgents == 0) *status = (uint8_t)((*status & POM_STATUS_INVALID) | POM_STATUS_DONE | POM_STATUS_DRAW);
    if (flags & g_invalid_mask) *status |= POM_STATUS_INVALID;
    return flags;
}

/* ---- InitBoardItems: bboard.cpp:346-382 with libstdc++'s generators restated ---- */
typedef struct { uint64_t mt[312]; int idx; } mt64_t;

static void mt64_seed(mt64_t* g, uint64_t seed)                      /* std::mt19937_64(seed): [rand.eng.mers] */
{
    g->mt[0] = seed;
    for (int i = 1; i < 312; i++) g->mt[i] = 6364136223846793005ULL * (g->mt[i - 1] ^ (g->mt[i - 1] >> 62)) + (uint64_t)i;
    g->idx = 312;
}

static uint64_t mt64_next(mt64_t* g)
{
    if (g->idx >= 312) {
        for (int i = 0; i < 312; i++) {
            uint64_t x = (g->mt[i] & 0xFFFFFFFF80000000ULL) | (g->mt[(i + 1) % 312] & 0x7FFFFFFFULL);
            uint64_t xa = x >> 1;
            if (x & 1ULL) xa ^= 0xB5026F5AA96619E9ULL;
            g->mt[i] = g->mt[(i + 156) % 312] ^ xa;
        }
        g->idx = 0;
    }
    uint64_t y = g->mt[g->idx++];
    y ^= (y >> 29) & 0x5555555555555555ULL;
    y ^= (y << 17) & 0x71D67FFFEDA60000ULL;
    y ^= (y << 37) & 0xFFF7EEE000000000ULL;
    y ^= (y >> 43);
    return y;
}

/* libstdc++ 13 uniform_int_distribution<int>(a,b) on a 64-bit URBG:
 * Lemire multiply-high with rejection (bits/uniform_int_dist.h:252-281,311-320) */
static int uniform_int(mt64_t* g, int a, int b)
{
    uint64_t range = (uint64_t)b - (uint64_t)a + 1;
    unsigned __int128 product = (unsigned __int128)mt64_next(g) * range;
    uint64_t low = (uint64_t)product;
    if (low < range) {
        uint64_t threshold = (0 - range) % range;
        while (low < threshold) {
            product = (unsigned __int128)mt64_next(g) * range;
            low = (uint64_t)product;
        }
    }
    return a + (int)(uint64_t)(product >> 64);
}

int pom_oracle_init_board_items(pom_state* s, int seed)
{
    mt64_t g;
    mt64_seed(&g, (uint64_t)(int64_t)seed);
    int q[POM_BOARD_CELLS + 1];
    int count = 0;
    for (int i = 0; i < BS; i++) {
        for (int j = 0; j < BS; j++) {
            int tmp = uniform_int(&g, 0, 6);
            int item = POM_ITEM_PASSAGE;                             /* ChooseItemOuter :59-74 */
            if (tmp == 2) item = POM_ITEM_WOOD;
            else if (tmp == 1) item = POM_ITEM_RIGID;
            s->board[i][j] = item;
            if (is_wood(item)) q[count++] = j + BS * i;
        }
    }
    int total = 0;
    int* flat = &s->board[0][0];
    for (;;) {
        int k = uniform_int(&g, 0, count);                           /* inclusive upper bound: D2 when k == count */
        if (k == count) return 1;
        int idx = q[k];
        if ((flat[idx] & 0xFF) == 0) {
            flat[idx] += uniform_int(&g, 1, 4);
            total++;
        }
        if ((float)total >= (float)count / 2) break;
    }
    return 0;
}

int pom_oracle_init_state(pom_state* s, int seed, int a0, int a1, int a2, int a3)
{
    int dirty = pom_oracle_init_board_items(s, seed);
    pom_oracle_put_agents_in_corners(s, a0, a1, a2, a3);
    return dirty;
}

/* ---- comparison helpers ---- */
int pom_oracle_state_diff(const pom_state* a, const pom_state* b)
{
    if (memcmp(a->board, b->board, sizeof(a->board))) return 1;
    if (a->timeStep != b->timeStep) return 2;
    if (a->aliveAgents != b->aliveAgents) return 3;
    for (int i = 0; i < 4; i++) {
        const pom_agent* p = &a->agents[i]; const pom_agent* q = &b->agents[i];
        if (p->x != q->x || p->y != q->y || p->bombCount != q->bombCount || p->maxBombCount != q->maxBombCount ||
            p->bombStrength != q->bombStrength || (p->canKick != 0) != (q->canKick != 0) || (p->dead != 0) != (q->dead != 0))
            return 4 + i;
    }
    if (memcmp(a->bombs, b->bombs, sizeof(a->bombs))) return 8;
    if (a->bombs_index != b->bombs_index || a->bombs_count != b->bombs_count) return 9;
    if (memcmp(a->flames, b->flames, sizeof(a->flames))) return 10;
    if (a->flames_index != b->flames_index || a->flames_count != b->flames_count) return 11;
    return 0;
}

static uint64_t fnv(uint64_t h, const void* p, size_t n)
{
    const uint8_t* c = (const uint8_t*)p;
    for (size_t i = 0; i < n; i++) { h ^= c[i]; h *= 1099511628211ULL; }
    return h;
}

uint64_t pom_oracle_state_hash(const pom_state* s)
{
    uint64_t h = 1469598103934665603ULL;
    h = fnv(h, s->board, sizeof(s->board));
    h = fnv(h, &s->timeStep, 8);
    for (int i = 0; i < 4; i++) {
        h = fnv(h, &s->agents[i], 20);
        uint8_t f[2] = { (uint8_t)(s->agents[i].canKick != 0), (uint8_t)(s->agents[i].dead != 0) };
        h = fnv(h, f, 2);
    }
    h = fnv(h, s->bombs, sizeof(s->bombs) + 8);
    h = fnv(h, s->flames, sizeof(s->flames) + 8);
    return h;
}

/* ---- shared stateless action source ---- */
static uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

uint32_t pom_oracle_rng_moves(uint64_t seed, uint64_t env, uint32_t tick, uint32_t n_actions)
{
    uint64_t h = splitmix64(splitmix64(seed ^ (env * 0xD6E8FEB86659FD93ULL)) + (uint64_t)tick);
    uint32_t out = 0;
    for (int a = 0; a < 4; a++) {
        uint32_t lane = (uint32_t)(h >> (16 * a)) & 0xFFFFu;
        out |= ((lane * n_actions) >> 16) << (8 * a);
    }
    return out;
}

/* ---- multi-threaded batch stepping (cpu_baseline kind "port") ---- */
typedef struct {
    pom_state* S; uint8_t* status; long n, lo, hi; const uint8_t* moves; int ticks;
    const pom_state* T; int nT; unsigned long long count;
} job_t;

static void* bench_worker(void* arg)
{
    job_t* j = (job_t*)arg;
    unsigned long long c = 0;
    uint32_t* episode = (uint32_t*)calloc((size_t)(j->hi - j->lo) + 1, sizeof(uint32_t));
    for (int k = 0; k < j->ticks; k++) {
        const uint8_t* mv = j->moves + ((size_t)k * (size_t)j->n) * 4;
        for (long e = j->lo; e < j->hi; e++) {
            if (j->status[e] & (POM_STATUS_DONE | POM_STATUS_INVALID)) continue;
            pom_oracle_env_step(&j->S[e], &j->status[e], mv + 4 * e);
            c++;
            if (j->T && (j->status[e] & (POM_STATUS_DONE | POM_STATUS_INVALID))) {
                uint32_t ep = ++episode[e - j->lo];
                j->S[e] = j->T[((uint64_t)e + ep) % (uint64_t)j->nT];
                j->status[e] = 0;
            }
        }
    }
    free(episode);
    j->count = c;
    return 0;
}

double pom_oracle_bench_steps(pom_state* states, uint8_t* status, long n, const uint8_t* moves,
                              int ticks, int nthreads, const pom_state* reset_templates,
                              int n_templates, unsigned long long* steps_out)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    job_t* jobs = (job_t*)calloc((size_t)nthreads, sizeof(job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < nthreads; t++) {
        job_t* j = &jobs[t];
        j->S = states; j->status = status; j->n = n; j->moves = moves; j->ticks = ticks;
        j->T = reset_templates; j->nT = n_templates;
        j->lo = n * t / nthreads; j->hi = n * (t + 1) / nthreads;
        pthread_create(&th[t], 0, bench_worker, j);
    }
    unsigned long long tot = 0;
    for (int t = 0; t < nthreads; t++) { pthread_join(th[t], 0); tot += jobs[t].count; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (steps_out) *steps_out = tot;
    free(jobs); free(th);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ---- batch helpers for the test harness ---- */
void pom_oracle_env_step_batch(pom_state* S, uint8_t* status, long n, const uint8_t* moves, uint8_t* flags_out)
{
    for (long e = 0; e < n; e++) {
        int f = pom_oracle_env_step(&S[e], &status[e], moves + 4 * e);
        if (flags_out) flags_out[e] = (uint8_t)f;
    }
}

void pom_oracle_step_batch(pom_state* S, long n, const uint8_t* moves, uint8_t* flags_out)
{
    for (long e = 0; e < n; e++) {
        int f = pom_oracle_step(&S[e], moves + 4 * e);
        if (flags_out) flags_out[e] = (uint8_t)f;
    }
}

/* first env whose states differ (envs with skip[e] != 0 ignored); -1 if none; *why = field group */
long pom_oracle_diff_batch(const pom_state* A, const pom_state* B, long n, const uint8_t* skip, int* why)
{
    for (long e = 0; e < n; e++) {
        if (skip && skip[e]) continue;
        int d = pom_oracle_state_diff(&A[e], &B[e]);
        if (d) { if (why) *why = d; return e; }
    }
    return -1;
}

void pom_oracle_hash_batch(const pom_state* S, long n, uint64_t* out)
{
    for (long e = 0; e < n; e++) out[e] = pom_oracle_state_hash(&S[e]);
}

void pom_oracle_rng_moves_batch(uint64_t seed, uint64_t env0, long n, uint32_t tick, uint32_t n_actions, uint8_t* moves_out)
{
    for (long e = 0; e < n; e++) {
        uint32_t m = pom_oracle_rng_moves(seed, env0 + (uint64_t)e, tick, n_actions);
        memcpy(moves_out + 4 * e, &m, 4);
    }
}

/* ---- partial observability: the State as agent `agent` sees it through a square window of `view` cells.
 * The reference only reserves the vocabulary (Item::FOG, bboard.hpp:62; "potentially fogged board state", :529; hidden
 * AgentInfo, :218-226); the rule is Pommerman's.  Written independently of the device code: a visibility map first,
 * then every component is filtered through it. */
void pom_oracle_fog(pom_state* s, int agent, int view)
{
    unsigned char vis[BS][BS];
    for (int y = 0; y < BS; y++)
        for (int x = 0; x < BS; x++) {
            int dx = x - s->agents[agent].x, dy = y - s->agents[agent].y;
            if (dx < 0) dx = -dx;
            if (dy < 0) dy = -dy;
            vis[y][x] = (unsigned char)((dx > dy ? dx : dy) <= view);
            if (!vis[y][x]) s->board[y][x] = POM_ITEM_FOG;
        }
    for (int i = 0; i < POM_AGENT_COUNT; i++) {
        pom_agent* g = &s->agents[i];
        if (i == agent) continue;
        int seen = !g->dead && !oob(g->x, g->y) && vis[g->y][g->x];
        if (!seen) { int d = g->dead; memset(g, 0, sizeof *g); g->x = -1; g->y = -1; g->dead = (uint8_t)d; }
    }
    pom_state t = *s;
    memset(s->bombs, 0, sizeof s->bombs);
    s->bombs_index = 0; s->bombs_count = 0;
    for (int i = 0; i < t.bombs_count && i < NB; i++) {
        int b = t.bombs[(t.bombs_index + i) % NB];
        if (!oob(b_x(b), b_y(b)) && vis[b_y(b)][b_x(b)]) s->bombs[s->bombs_count++] = b;
    }
    memset(s->flames, 0, sizeof s->flames);
    s->flames_index = 0; s->flames_count = 0;
    for (int i = 0; i < t.flames_count && i < NB; i++) {
        pom_flame f = t.flames[(t.flames_index + i) % NB];
        if (!oob(f.x, f.y) && vis[f.y][f.x]) s->flames[s->flames_count++] = f;
    }
}

void pom_oracle_fog_batch(pom_state* S, long n, int agent, int view)
{
    for (long e = 0; e < n; e++) pom_oracle_fog(&S[e], agent, view);
}

/* ---- observation planes: the DEFINITION the device code is checked against (the reference has no counterpart).
 * Built from the AoS State through pom_oracle_fog, i.e. independently of the packed record. */
void pom_oracle_observe_planes(const pom_state* full, int agent, int view, uint8_t out[512])
{
    pom_state s = *full;
    pom_oracle_fog(&s, agent, view);
    memset(out, 0, 512);
    for (int y = 0; y < BS; y++)
        for (int x = 0; x < BS; x++) {
            int v = s.board[y][x], id;
            if (is_agent(v)) id = 10 + (v - POM_ITEM_AGENT0);
            else if (is_flame(v)) id = 4;
            else if (is_wood(v)) id = 2;
            else id = v;                                   /* 0 1 3 5 6 7 8 9 keep the reference's ids */
            out[x + BS * y] = (uint8_t)id;
            if (is_flame(v)) {
                int o = flame_id(v), t = 0;
                for (int i = 0; i < full->flames_count && i < NB; i++) {       /* first entry of the FULL queue with this origin */
                    const pom_flame* f = &full->flames[(full->flames_index + i) % NB];
                    if (f->x + BS * f->y == o) { t = f->timeLeft; break; }
                }
                out[363 + x + BS * y] = (uint8_t)(t < 0 ? 0 : t);
            }
        }
    for (int i = 0; i < s.bombs_count; i++) {              /* the fogged queue: visible bombs only, in order */
        int b = s.bombs[i];
        out[121 + b_x(b) + BS * b_y(b)] = (uint8_t)b_str(b);
        out[242 + b_x(b) + BS * b_y(b)] = (uint8_t)b_time(b);
    }
    const pom_agent* a = &full->agents[agent];
    int ammo = a->maxBombCount - a->bombCount, alive = 0;
    for (int i = 0; i < 4; i++) alive |= full->agents[i].dead ? 0 : (1 << i);
    out[484] = (uint8_t)a->x; out[485] = (uint8_t)a->y;
    out[486] = (uint8_t)(ammo < 0 ? 0 : ammo > 255 ? 255 : ammo);
    out[487] = (uint8_t)a->bombStrength; out[488] = (uint8_t)(a->canKick ? 1 : 0); out[489] = (uint8_t)alive;
    out[490] = (uint8_t)(full->timeStep & 0xFF); out[491] = (uint8_t)((full->timeStep >> 8) & 0xFF);
    out[492] = (uint8_t)((alive >> agent) & 1);
}

/* the cropped layout (include/pom_batch.h, pom_batch_observe_planes_cropped): the window of the full planes, row-major,
 * plane after plane, then the twelve scalar bytes; window cells off the board hold 5 (fog) in the board plane, 0 elsewhere.
 * Defined THROUGH the full-layout definition above, so that the two layouts cannot drift apart. */
long pom_oracle_obs_cropped_bytes(int view) { long w = 2 * view + 1; return (4 * w * w + 12 + 31) & ~31L; }

void pom_oracle_observe_cropped(const pom_state* full, int agent, int view, uint8_t* out)
{
    uint8_t planes[512];
    const int W = 2 * view + 1, W2 = W * W;
    const int ax = full->agents[agent].x, ay = full->agents[agent].y;
    pom_oracle_observe_planes(full, agent, view, planes);
    memset(out, 0, (size_t)pom_oracle_obs_cropped_bytes(view));
    for (int p = 0; p < 4; p++)
        for (int row = 0; row < W; row++)
            for (int col = 0; col < W; col++) {
                int x = ax - view + col, y = ay - view + row;
                int inb = x >= 0 && x < BS && y >= 0 && y < BS;
                out[p * W2 + row * W + col] = inb ? planes[p * BS * BS + x + BS * y] : (uint8_t)(p == 0 ? 5 : 0);
            }
    memcpy(out + 4 * W2, planes + 484, 12);
}

void pom_oracle_observe_cropped_batch(const pom_state* S, long n, int agent, int view, uint8_t* out)
{
    const long rb = pom_oracle_obs_cropped_bytes(view);
    for (long e = 0; e < n; e++) pom_oracle_observe_cropped(&S[e], agent, view, out + rb * e);
}

void pom_oracle_observe_planes_batch(const pom_state* S, long n, int agent, int view, uint8_t* out)
{
    for (long e = 0; e < n; e++) pom_oracle_observe_planes(&S[e], agent, view, out + 512 * e);
}
