/*
 * hostsim.cpp — TEST-ONLY host build of the kernel body (pomcpp_b200/csrc/pom_core.cuh).
 *
 * The tick logic that the CUDA kernels run is a header of `__host__ __device__` functions; this
 * file compiles the very same header with g++ so that the logic can be differential-tested against
 * the oracle on the CPU-only build container (millions of steps in seconds) before any GPU time is
 * spent.  It is NOT part of the product: libpom_b200.so contains no host stepping path and nothing
 * in pomcpp_b200/ loads this library.
 */
#include <cstdint>
#include <cstring>
#include "pom_core.cuh"
#include "pom_policy.cuh"

extern "C" {

int hostsim_record_bytes() { return POM_REC_BYTES; }

void hostsim_pack_batch(const pom_state* S, const uint8_t* status, long n, uint8_t* recs, uint8_t* bad)
{
    for(long e = 0; e < n; e++)
    {
        int b = pomcore::pack(&S[e], status ? status[e] : 0, recs + e * POM_REC_BYTES);
        if(bad) bad[e] = uint8_t(b);
    }
}

void hostsim_unpack_batch(const uint8_t* recs, long n, pom_state* S, uint8_t* status)
{
    for(long e = 0; e < n; e++)
    {
        uint8_t st = pomcore::unpack(recs + e * POM_REC_BYTES, &S[e]);
        if(status) status[e] = st;
    }
}

/* raw != 0: bboard::Step only; raw == 0: Environment::Step semantics.
 * by_rays != 0: the tick in the form the warp-cooperative kernels run it (ray-parallel explosions, arm-parallel flame
 * pops: the lanes' work done one after the other), so that the scan / commit decomposition itself is checked on the CPU */
static int g_by_rays = 0;
void hostsim_set_by_rays(int on) { g_by_rays = on; }
/* POM_STEP_CONTINUE_UNDEFINED: D3 / D5 ticks continue with the canonical result instead of freezing the env */
static int g_invalid_mask = pomcore::F_INVALID_MASK;
void hostsim_set_continue_undefined(int on) { g_invalid_mask = on ? pomcore::F_INVALID_MASK_CONTINUE : pomcore::F_INVALID_MASK; }

void hostsim_step_records(uint8_t* recs, long n, const uint8_t* moves, int raw, uint8_t* flags_out)
{
    for(long e = 0; e < n; e++)
    {
        uint32_t m;
        std::memcpy(&m, moves + 4 * e, 4);
        uint8_t* r = recs + e * POM_REC_BYTES;
        int f = 0;
        if(!g_by_rays) f = raw ? pomcore::step(r, m) : pomcore::env_step(r, m, g_invalid_mask);
        else if(raw || !(r[R_STATUS] & (POM_STATUS_DONE | POM_STATUS_INVALID)))
        {
            f = pomcore::step_by_rays(r, m);
            if(!raw)
            {
                if(f & g_invalid_mask) r[R_STATUS] |= POM_STATUS_INVALID;
                pomcore::env_post(r);
            }
        }
        if(flags_out) flags_out[e] = uint8_t(f);
    }
}

void hostsim_spawn_flame(uint8_t* rec, int x, int y, int strength)
{
    pomcore::Agents A;
    pomcore::load_agents(rec, A);
    int flags = 0;
    pomcore::explode(rec, A, uint32_t(x) | (uint32_t(y) << 4), uint32_t(strength), 31u, flags);
    pomcore::store_agents(rec, A);
}

/* the device rollout policy (pom_policy.cuh) on packed records: moves of the agents in `mask` are replaced;
 * A = [n][4] pom_simple_agent; draws from rng_moves(seed, env0 + e, tick, 5) */
void hostsim_simple_moves(const uint8_t* recs, long n, pom_simple_agent* A, uint64_t seed, uint64_t env0, uint32_t tick,
                          uint32_t mask, uint8_t* moves)
{
    for(long e = 0; e < n; e++)
    {
        const uint8_t* r = recs + e * POM_REC_BYTES;
        if(r[R_STATUS] & (POM_STATUS_DONE | POM_STATUS_INVALID)) continue;
        uint32_t m;
        std::memcpy(&m, moves + 4 * e, 4);
        pompolicy::SimpleSt st[4];
        std::memcpy(st, A + 4 * e, sizeof st);
        pompolicy::ArrayStore store{ st };
        m = pompolicy::simple_moves(r, mask, m, pomcore::rng_moves(seed, env0 + uint64_t(e), tick, 5), store);
        std::memcpy(A + 4 * e, st, sizeof st);
        std::memcpy(moves + 4 * e, &m, 4);
    }
}

/* MoveTowardsPosition as the device policy code computes it (a HUNT flood job towards cell (tx, ty)), for direct
 * comparison with the queue BFS of the oracle / the reference on arbitrary boards; returns the Move */
int hostsim_move_towards(const uint8_t* rec, int agent, int tx, int ty)
{
    const pompolicy::Boards B = pompolicy::make_boards(rec);
    pompolicy::Plan pl[4];
    for(int a = 0; a < 4; a++) pl[a].kind = pompolicy::K_NONE;
    pl[agent].kind = pompolicy::K_HUNT;
    pl[agent].target = uint8_t(tx + 11 * ty);
    pl[agent].move = POM_MOVE_IDLE;
    pompolicy::run_floods(rec, B, pl, 1u << agent);
    return pl[agent].move;
}

/* MoveTowardsSafePlace as the device policy code computes it (a FLEE job with the given radius) */
int hostsim_move_towards_safe_place(const uint8_t* rec, int agent, int radius)
{
    const pompolicy::Boards B = pompolicy::make_boards(rec);
    pompolicy::Plan pl[4];
    for(int a = 0; a < 4; a++) pl[a].kind = pompolicy::K_NONE;
    const uint32_t pos = rec[R_APOS + agent];
    pl[agent].kind = pompolicy::K_FLEE;
    pl[agent].target = 0xFFu;
    pl[agent].move = POM_MOVE_IDLE;
    pl[agent].dg = uint32_t(radius) & 15u;
    pl[agent].region = pompolicy::safe_place_region(int(pos & 15u), int(pos >> 4), radius);
    pompolicy::run_floods(rec, B, pl, 1u << agent);
    return pl[agent].move;
}

/* the device fog code (pomcore::fog_state) on AoS states */
void hostsim_fog_batch(pom_state* S, long n, int agent, int view)
{
    for(long e = 0; e < n; e++) pomcore::fog_state(&S[e], agent, view);
}

/* the device observation code (pomcore::observe_planes) on packed records */
void hostsim_observe_planes(const uint8_t* recs, long n, int agent, int view, uint8_t* out)
{
    for(long e = 0; e < n; e++) pomcore::observe_planes(recs + e * POM_REC_BYTES, agent, view, out + 512 * e);
}

/* the cropped layout (pomcore::observe_cropped) */
long hostsim_obs_cropped_bytes(int view) { return long(pomcore::obs_cropped_bytes(view)); }
void hostsim_observe_cropped(const uint8_t* recs, long n, int agent, int view, uint8_t* out)
{
    const long rb = long(pomcore::obs_cropped_bytes(view));
    for(long e = 0; e < n; e++) pomcore::observe_cropped(recs + e * POM_REC_BYTES, agent, view, out + rb * e);
}

uint32_t hostsim_rng_moves(uint64_t seed, uint64_t env, uint32_t tick, uint32_t n_actions)
{
    return pomcore::rng_moves(seed, env, tick, n_actions);
}

}
