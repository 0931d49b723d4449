/*
 * pom_policy.cuh — the reference's heuristic agent as a device-side rollout policy.
 *
 * Re-states agents::SimpleAgent (reference src/agents/simple_agent.cpp:12-141) and the bboard::strategy helpers
 * it calls (src/bboard/strategy.cpp:37-338, include/strategy.hpp:134-168) on the packed record of pom_record.h:
 * one thread computes the move of one agent of its own env, reading the record in place.  It is the caller
 * of the step path in the reference's own benchmark (unit_test/bboard/performance_test.cpp:38,59-63).
 *
 * What is kept bit-exact: every decision of _Decide including the reference's quirks — the loop bounds of
 * MoveTowardsSafePlace (`y < radius`, strategy.cpp:130-132), SortDirections re-appending the element AFTER the
 * removed one (strategy.hpp:148-149), _HasRPLoop being vacuously true for fewer than two remembered positions
 * and reading an unwritten slot at two (simple_agent.cpp:24-35), moveQueue[1] being read when only one
 * direction is safe (:48,:125), paths ending at (but including) agent cells (strategy.cpp:44-52).
 *
 * What is restructured for the GPU:
 *   * the agent's persistent members are 8 bytes (pom_simple_agent, include/pom_state.h) instead of a 3 KB object;
 *     its private mt19937_64 is replaced by one caller-supplied draw in 0..4 per act (every path of _Decide
 *     consumes at most one intDist(rng));
 *   * FillRMap's 121-int distance/predecessor map becomes one byte per cell holding the FIRST MOVE of the BFS
 *     path (non-zero = reachable).  MoveTowardsPosition (strategy.cpp:101-124) only ever follows the predecessor
 *     chain back to the cell next to the source, so the label is all it needs; the BFS visiting order
 *     (down, up, right, left; FIFO) is kept because it decides which first move a cell inherits;
 *   * the BFS runs lazily: only the two branches that read the map (danger > 0, enemy within 7) pay for it;
 *   * IsInDanger for the agent's cell and its four neighbours (the nine calls of _Decide + SafeDirections) is one
 *     pass over the bomb ring.
 *
 * Compiles as plain C++ too (POM_HD), so tests/hostsim can differential-test it on the CPU against the oracle.
 */
#ifndef POM_POLICY_CUH_
#define POM_POLICY_CUH_

#include "pom_core.cuh"

namespace pompolicy
{
using namespace pomcore;

/* persistent agent state, same bytes as pom_simple_agent: w0 = recent[4]; w1 = rp_index | rp_count<<8 | move_queue<<16 */
struct SimpleSt { uint32_t w0, w1; };

/* DesiredPosition on a nibble-packed position; x or y leaving the board wraps inside its own nibble (-1 -> 15, 11 stays 11) */
POM_HD uint32_t pos_step(uint32_t p, uint32_t m)
{
    const int d = move_delta(m);
    return (d & 15) ? ((p & 0xF0u) | (uint32_t(int(p) + d) & 15u)) : (uint32_t(int(p) + d) & 0xFFu);
}
POM_HD bool pos_oob(uint32_t p) { return (p & 15u) > 10u || (p >> 4) > 10u; }

/* IsInDanger (strategy.cpp:225-246) at (x, y), which may lie one cell off the board */
POM_HD uint32_t danger_at(const uint8_t* r, int x, int y)
{
    const int n = r[R_BCOUNT];
    uint32_t slot = r[R_BINDEX];
    uint32_t best = 16u;
    POM_LOOP
    for(int i = 0; i < n; i++)
    {
        const uint32_t b = reinterpret_cast<const uint32_t*>(r + R_BOMBS)[slot];
        slot = ring_next(slot);
        const int bx = int(b & 15u), by = int((b >> 4) & 15u), s = int((b >> 12) & 15u);
        const uint32_t t = (b >> 16) & 15u;
        const int dx = x - bx, dy = y - by;
        const bool hit = (dy == 0 && dx >= -s && dx <= s) || (dx == 0 && dy >= -s && dy <= s);   /* IsInBombRange, strategy.hpp:163-168 */
        if(hit && t < best) best = t;
    }
    return best == 16u ? 0u : best;
}

/* IsInDanger at DesiredPosition(x, y, m) for m = IDLE, UP, DOWN, LEFT, RIGHT: nibble m of the result */
POM_HD uint32_t danger5(const uint8_t* r, int x, int y)
{
    const int n = r[R_BCOUNT];
    uint32_t slot = r[R_BINDEX];
    uint32_t b0 = 16u, b1 = 16u, b2 = 16u, b3 = 16u, b4 = 16u;
    POM_LOOP
    for(int i = 0; i < n; i++)
    {
        const uint32_t b = reinterpret_cast<const uint32_t*>(r + R_BOMBS)[slot];
        slot = ring_next(slot);
        const int s = int((b >> 12) & 15u);
        const uint32_t t = (b >> 16) & 15u;
        const int dx = x - int(b & 15u), dy = y - int((b >> 4) & 15u);
        const bool rowx = dx >= -s && dx <= s, coly = dy >= -s && dy <= s;
        if(((dy == 0 && rowx) || (dx == 0 && coly)) && t < b0) b0 = t;
        if(((dy == 1 && rowx) || (dx == 0 && dy - 1 >= -s && dy - 1 <= s)) && t < b1) b1 = t;       /* UP: y-1 */
        if(((dy == -1 && rowx) || (dx == 0 && dy + 1 >= -s && dy + 1 <= s)) && t < b2) b2 = t;      /* DOWN: y+1 */
        if(((dy == 0 && dx - 1 >= -s && dx - 1 <= s) || (dx == 1 && coly)) && t < b3) b3 = t;       /* LEFT: x-1 */
        if(((dy == 0 && dx + 1 >= -s && dx + 1 <= s) || (dx == -1 && coly)) && t < b4) b4 = t;      /* RIGHT: x+1 */
    }
    return (b0 & 15u) | ((b1 & 15u) << 4) | ((b2 & 15u) << 8) | ((b3 & 15u) << 12) | ((b4 & 15u) << 16);
}

POM_HD bool safe_condition(uint32_t danger, uint32_t min) { return danger == 0u || danger >= min; }   /* strategy.cpp:190-193 */

/* FillRMap (strategy.cpp:58-95) as first-move labels.  lab[cell] = 0 unreachable (or the source), else the Move of
 * the first step of the BFS path source -> cell.  `q` is the FIFO of cell ids. */
struct Reach { uint8_t lab[124]; };

POM_HD void fill_reach(const uint8_t* r, uint32_t src_pos, Reach& R)
{
    uint8_t q[124];
    uint32_t* lw = reinterpret_cast<uint32_t*>(R.lab);
    for(int k = 0; k < 31; k++) lw[k] = 0u;
    const int sx = int(src_pos & 15u), sy = int(src_pos >> 4);
    const int src = sx + 11 * sy;
    int head = 0, tail = 0;
    q[tail++] = uint8_t(src);
    POM_LOOP
    while(head != tail)
    {
        const int c = q[head++];
        const int cx = c % 11, cy = c / 11;
        const uint32_t inherit = R.lab[c];                     /* 0 only for the source */
        /* TryAdd order: (x, y+1), (x, y-1), (x+1, y), (x-1, y)  = DOWN, UP, RIGHT, LEFT */
        POM_LOOP
        for(int k = 0; k < 4; k++)
        {
            const uint32_t mv = 0x03040102u >> (8 * k) & 0xFFu;
            const int nx = cx + ((k == 2) ? 1 : (k == 3) ? -1 : 0);
            const int ny = cy + ((k == 0) ? 1 : (k == 1) ? -1 : 0);
            if(uint32_t(nx) > 10u || uint32_t(ny) > 10u) continue;
            const int nc = nx + 11 * ny;
            if(nc == src || R.lab[nc] != 0u) continue;
            const uint32_t code = r[R_BOARD + nc];
            const bool agent = c_is_agent(code);
            if(!(c_is_walkable(code) || agent)) continue;
            R.lab[nc] = uint8_t(inherit ? inherit : mv);
            if(!agent) q[tail++] = uint8_t(nc);                /* paths end at agent cells (strategy.cpp:50-52) */
        }
    }
}

/* MoveTowardsPosition (strategy.cpp:101-124) for a target != source */
POM_HD uint32_t move_towards(const Reach& R, uint32_t src_pos, int tx, int ty)
{
    const uint32_t l = R.lab[tx + 11 * ty];
    if(l) return l;
    /* unreachable target: its predecessor field is 0 = cell (0,0).  Only a source standing ON (0,0) takes the
     * "predecessor is the source" branch and walks towards the target's side; everybody else gets IDLE. */
    if(src_pos != 0u) return POM_MOVE_IDLE;
    return tx > 0 ? uint32_t(POM_MOVE_RIGHT) : uint32_t(POM_MOVE_DOWN);
}

/* MoveTowardsSafePlace (strategy.cpp:126-144); the loops stop at `radius`, not at origin + radius (sic) */
POM_HD uint32_t move_towards_safe_place(const uint8_t* r, const Reach& R, uint32_t src_pos, int radius)
{
    const int ox = int(src_pos & 15u), oy = int(src_pos >> 4);
    const int y0 = oy - radius < 0 ? 0 : oy - radius, y1 = radius < 11 ? radius : 11;
    const int x0 = ox - radius < 0 ? 0 : ox - radius, x1 = radius < 11 ? radius : 11;
    POM_LOOP
    for(int y = y0; y < y1; y++)
    {
        POM_LOOP
        for(int x = x0; x < x1; x++)
        {
            const int dx = x - ox, dy = y - oy;
            if((dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy) > radius) continue;
            if(R.lab[x + 11 * y] != 0u && safe_condition(danger_at(r, x, y), 2u)) return R.lab[x + 11 * y];
        }
    }
    return POM_MOVE_IDLE;
}

/* the 3-bit slots of moveQueue.queue */
POM_HD uint32_t mq_get(uint32_t mq, int i) { return (mq >> (3 * (i & 3))) & 7u; }
POM_HD uint32_t mq_set(uint32_t mq, int i, uint32_t v) { const int s = 3 * (i & 3); return (mq & ~(7u << s)) | (v << s); }

/* _MoveSafeOneSpace / the tail of _Decide (simple_agent.cpp:37-49,113-126): SafeDirections (strategy.cpp:194-219)
 * + SortDirections (strategy.hpp:134-158) + the random pick among the first two */
POM_HD uint32_t pick_safe_direction(const uint8_t* r, uint32_t pos, uint32_t dg, SimpleSt& st, uint32_t draw)
{
    uint32_t mq = st.w1 >> 16;
    int count = 0;
    POM_LOOP
    for(int k = 0; k < 4; k++)
    {
        const uint32_t mv = 4u - uint32_t(k);                  /* RIGHT, LEFT, DOWN, UP */
        const uint32_t p = pos_step(pos, mv);
        if(pos_oob(p)) continue;
        if(!c_is_walkable(r[R_BOARD + cell_of(p)])) continue;
        if(!safe_condition((dg >> (4u * mv)) & 15u, 2u)) continue;
        mq = mq_set(mq, count, mv);
        count++;
    }
    const int moves = count;
    const uint32_t rp_index = st.w1 & 0xFFu, rp_count = (st.w1 >> 8) & 0xFFu;
    int removes = 0;
    POM_LOOP
    for(int i = 0; i < moves && removes < 4; i++)
    {
        const uint32_t p = pos_step(pos, mq_get(mq, i));
        bool seen = false;
        for(uint32_t j = 0; j < rp_count; j++) seen = seen || byte_of(st.w0, int((rp_index + j) & 3u)) == p;
        if(seen)
        {
            for(int k = i + 1; k < count; k++) mq = mq_set(mq, k - 1, mq_get(mq, k));   /* RemoveAt(i) */
            count--;
            mq = mq_set(mq, count, mq_get(mq, i));                                      /* AddElem(q[i]) */
            count++;
            i--;
            removes++;
        }
    }
    st.w1 = (st.w1 & 0xFFFFu) | (mq << 16);
    if(count == 0) return POM_MOVE_IDLE;
    return mq_get(mq, int(draw & 1u));
}

/* _Decide, simple_agent.cpp:52-127 */
POM_HD uint32_t simple_decide(const uint8_t* r, int id, SimpleSt& st, uint32_t draw)
{
    const uint32_t pos = r[R_APOS + id];
    const int x = int(pos & 15u), y = int(pos >> 4);
    const uint32_t dg = danger5(r, x, y);
    const uint32_t danger = dg & 15u;
    Reach R;
    if(danger > 0u)
    {
        fill_reach(r, pos, R);
        const uint32_t m = move_towards_safe_place(r, R, pos, int(danger));
        const uint32_t p = pos_step(pos, m);
        if(!pos_oob(p) && c_is_walkable(r[R_BOARD + cell_of(p)]) && safe_condition((dg >> (4u * m)) & 15u, 2u)) return m;
        return pick_safe_direction(r, pos, dg, st, draw);
    }
    if(int(int8_t(r[R_ABCNT + id])) < int(r[R_AMAX + id]))
    {
        /* IsAdjacentEnemy (strategy.cpp:296-312): nearest live enemy, Manhattan */
        int dmin = 99, target = -1;
        for(int i = 0; i < 4; i++)
        {
            if(i == id || (r[R_AFLAGS + i] & AF_DEAD)) continue;
            const int dx = int(r[R_APOS + i] & 15u) - x, dy = int(r[R_APOS + i] >> 4) - y;
            const int d = (dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy);
            if(d < dmin) dmin = d;
            /* MoveTowardsEnemy (strategy.cpp:165-183) takes the FIRST agent within the radius that does not share
             * the source's cell */
            if(target < 0 && d <= 7 && d != 0) target = i;
        }
        if(dmin <= 1) return POM_MOVE_BOMB;
        if(dmin <= 7)
        {
            /* _HasRPLoop, simple_agent.cpp:24-35 */
            const uint32_t rp_index = st.w1 & 0xFFu, rp_count = (st.w1 >> 8) & 0xFFu;
            bool loop = true;
            for(uint32_t i = 0; i < rp_count / 2u; i++)
                loop = loop && byte_of(st.w0, int((rp_index + i) & 3u)) == byte_of(st.w0, int((rp_index + i + 2u) & 3u));
            if(loop) return draw & 3u;                          /* Move(intDist(rng) % 4) */
            uint32_t m = POM_MOVE_IDLE;
            if(target >= 0)
            {
                fill_reach(r, pos, R);
                m = move_towards(R, pos, int(r[R_APOS + target] & 15u), int(r[R_APOS + target] >> 4));
            }
            const uint32_t p = pos_step(pos, m);
            if(!pos_oob(p) && c_is_walkable(r[R_BOARD + cell_of(p)]) && safe_condition((dg >> (4u * m)) & 15u, 5u)) return m;
        }
        /* IsAdjacentItem(state, id, 1, WOOD), strategy.cpp:314-338: own cell and the four neighbours */
        bool wood = c_is_wood(r[R_BOARD + x + 11 * y]);
        for(uint32_t mv = 1; mv <= 4u; mv++)
        {
            const uint32_t p = pos_step(pos, mv);
            if(!pos_oob(p)) wood = wood || c_is_wood(r[R_BOARD + cell_of(p)]);
        }
        if(wood) return POM_MOVE_BOMB;
    }
    return pick_safe_direction(r, pos, dg, st, draw);
}

/* SimpleAgent::act, simple_agent.cpp:128-141 */
POM_HD uint32_t simple_act(const uint8_t* r, int id, SimpleSt& st, uint32_t draw)
{
    const uint32_t m = simple_decide(r, id, st, draw);
    const uint32_t p = pos_step(r[R_APOS + id], m);            /* BOMB and IDLE remember the agent's own cell */
    uint32_t idx = st.w1 & 0xFFu, cnt = (st.w1 >> 8) & 0xFFu;
    if(cnt == 4u) { idx = (idx + 1u) & 3u; cnt = 3u; }          /* PopElem when full */
    st.w0 = with_byte(st.w0, int((idx + cnt) & 3u), p);
    cnt++;
    st.w1 = (st.w1 & 0xFFFF0000u) | idx | (cnt << 8);
    return m;
}

/* Environment::Step's collection loop (environment.cpp:137-146) for the agents in `mask`: returns `moves` with the
 * bytes of the masked agents replaced (IDLE for a dead agent, whose slot the reference leaves unwritten).
 * `draws` = pom_rng_moves(seed, env, tick, 5): byte a is agent a's uniform{0..4} draw.  `S` gives access to the
 * four agent states: S.load(a) / S.store(a, st) (an array on the host, shared or global memory in the kernels). */
template<class Store>
POM_HD uint32_t simple_moves(const uint8_t* r, uint32_t mask, uint32_t moves, uint32_t draws, Store& S)
{
    POM_LOOP
    for(int a = 0; a < 4; a++)
    {
        if(!((mask >> a) & 1u)) continue;
        uint32_t m = POM_MOVE_IDLE;
        if(!(r[R_AFLAGS + a] & AF_DEAD))
        {
            SimpleSt st = S.load(a);
            m = simple_act(r, a, st, byte_of(draws, a));
            S.store(a, st);
        }
        moves = with_byte(moves, a, m);
    }
    return moves;
}

/* agent states in a plain array of four */
struct ArrayStore {
    SimpleSt* st;
    POM_HD SimpleSt load(int a) const { return st[a]; }
    POM_HD void store(int a, const SimpleSt& v) { st[a] = v; }
};

}
#endif
