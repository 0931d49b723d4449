/*
 * pom_oracle.h — TEST INFRASTRUCTURE ONLY (the parity checker, never the product).
 *
 * CPU restatement, in plain C, of the reference's step path (dist1ll/pomcpp
 * src/bboard/step.cpp, step_utility.cpp, bboard.cpp, environment.cpp:123-169) on the
 * AoS `pom_state` of include/pom_state.h.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity pinning: the restatement is checked (tests/test_oracle_vs_ref.py, run in the
 * build container where /root/reference exists) against the UNMODIFIED reference
 * compiled into oracle/_ref/libpomref.so, on every scenario of the reference's own unit
 * tests and on millions of random steps, field by field after every tick; the golden
 * vectors under tests/golden/ were produced by that compiled reference
 * (tests/golden/make_golden.py) and pin it where /root/reference is absent.
 */
#ifndef POM_ORACLE_H_
#define POM_ORACLE_H_

#include <stdint.h>
#include "pom_state.h"

#ifdef __cplusplus
extern "C" {
#endif

/* flags reported by the step functions (same meaning as POM_STATUS_INVALID reasons) */
enum {
    POM_ORC_D1_UNREACHABLE = 0x01, /* dependency walk ran out of roots (canonical: stop)   */
    POM_ORC_D3_NULL_BOMB   = 0x02, /* kicker on a BOMB cell without queue entry (ref: crash) */
    POM_ORC_D4_BOMB_OVF    = 0x04, /* bomb queue would exceed 20 entries                    */
    POM_ORC_FLAME_OVF      = 0x08, /* flame queue would exceed 20 entries                   */
    POM_ORC_BAD_MOVE       = 0x10, /* move byte outside 0..5 (treated as IDLE)              */
    POM_ORC_D5_REVERT_LOOP = 0x20, /* AgentBombChainReversion would recurse forever (ref: hang/stack overflow) */
    POM_ORC_INVALID_MASK   = 0x3E  /* everything but D1 takes the env out of the reference's defined domain */
};

/* std::make_unique<State>() : zero bytes + default member initialisers (bboard.hpp:225-240,372-373) */
void pom_oracle_zero_state(pom_state* s);

/* bboard::Step (step.cpp:9-284); returns POM_ORC_* flags */
int pom_oracle_step(pom_state* s, const uint8_t moves[4]);

/* Environment::Step on a bare state + status byte (environment.cpp:125-128,149-168) */
int pom_oracle_env_step(pom_state* s, uint8_t* status, const uint8_t moves[4]);
void pom_oracle_set_continue_undefined(int on);   /* D3 / D5 ticks keep the env running (canonical continuation) */

/* State primitives used by fixtures (bboard.cpp) */
void pom_oracle_put_agent(pom_state* s, int x, int y, int id);                 /* :313-320 */
void pom_oracle_put_agents_in_corners(pom_state* s, int a0, int a1, int a2, int a3); /* :322-333 */
void pom_oracle_kill(pom_state* s, int id);                                    /* bboard.hpp:474-481 */
void pom_oracle_plant_bomb(pom_state* s, int x, int y, int id, int lifeTime, int setItem); /* :125-146 */
int  pom_oracle_spawn_flame(pom_state* s, int x, int y, int strength);         /* :198-263 */

/* step utilities exposed for the reference's [step utilities] known-answer tests */
void pom_oracle_fill_dest_pos(const pom_state* s, const uint8_t moves[4], int pos8[8]);   /* step_utility.cpp:138-144 */
void pom_oracle_fix_switch_move(const pom_state* s, int pos8[8]);                          /* :154-170 */
int  pom_oracle_resolve_dependencies(const pom_state* s, const int pos8[8], int dependency[4], int roots[4]); /* :172-205 */

/* InitBoardItems (bboard.cpp:346-382) with libstdc++ 13's mt19937_64 +
 * uniform_int_distribution restated.  Returns 0 if the seed is clean, 1 if the
 * reference would read the uninitialised slot q[q.count] (defect D2; board then undefined). */
int  pom_oracle_init_board_items(pom_state* s, int seed);
/* InitState (bboard.cpp:339-344) on a zeroed state; returns the D2 flag */
int  pom_oracle_init_state(pom_state* s, int seed, int a0, int a1, int a2, int a3);

/* field-wise comparison; returns 0 if equal, else 1 + index of the first differing field group */
int  pom_oracle_state_diff(const pom_state* a, const pom_state* b);

/* FNV-1a over the meaningful fields (padding bytes skipped) */
uint64_t pom_oracle_state_hash(const pom_state* s);

/* the shared stateless action source (SURVEY §8d "Shared RNG"): splitmix64 of
 * (seed, env, tick) -> four moves, each mulhi32(x, n_actions)                            */
uint32_t pom_oracle_rng_moves(uint64_t seed, uint64_t env, uint32_t tick, uint32_t n_actions);

/* multi-threaded stepping of an AoS batch with this restatement (cpu_baseline kind "port").
 * Same contract as ref_bench_steps in ref_shim.cpp. */
double pom_oracle_bench_steps(pom_state* states, uint8_t* status, long n, const uint8_t* moves,
                              int ticks, int nthreads, const pom_state* reset_templates,
                              int n_templates, unsigned long long* steps_out);

/* batch helpers for the test harness (moves: [n][4] bytes) */
void pom_oracle_env_step_batch(pom_state* S, uint8_t* status, long n, const uint8_t* moves, uint8_t* flags_out);
void pom_oracle_step_batch(pom_state* S, long n, const uint8_t* moves, uint8_t* flags_out);
long pom_oracle_diff_batch(const pom_state* A, const pom_state* B, long n, const uint8_t* skip, int* why);
void pom_oracle_hash_batch(const pom_state* S, long n, uint64_t* out);
void pom_oracle_rng_moves_batch(uint64_t seed, uint64_t env0, long n, uint32_t tick, uint32_t n_actions, uint8_t* moves_out);

/* the State as agent `agent` observes it through a square window of `view` cells: Item::FOG outside (bboard.hpp:62),
 * agents / bombs / flames outside not exposed (bboard.hpp:218-226).  Declared but not implemented by the reference
 * (bboard.hpp:529), so this is the definition the device code is checked against, not a restatement. */
void pom_oracle_fog(pom_state* s, int agent, int view);
void pom_oracle_fog_batch(pom_state* S, long n, int agent, int view);

/* observation planes of agent `agent` (layout: include/pom_batch.h POM_OBS_BYTES); a definition like pom_oracle_fog */
void pom_oracle_observe_planes(const pom_state* s, int agent, int view, uint8_t out[512]);
void pom_oracle_observe_planes_batch(const pom_state* S, long n, int agent, int view, uint8_t* out);
long pom_oracle_obs_cropped_bytes(int view);
void pom_oracle_observe_cropped(const pom_state* s, int agent, int view, uint8_t* out);
void pom_oracle_observe_cropped_batch(const pom_state* S, long n, int agent, int view, uint8_t* out);

/* ---- agents::SimpleAgent + bboard::strategy (pom_oracle_agent.c) ----
 * act() of the reference's heuristic agent `id` on state s (simple_agent.cpp:128-141); `st` holds the
 * agent's persistent members, `draw` (0..4) replaces its one intDist(rng) call.  Returns the Move. */
int  pom_oracle_simple_act(const pom_state* s, int id, pom_simple_agent* st, int draw);
/* moves of the agents in agent_mask for a batch (A: [n][4] agent records; draws from
 * pom_oracle_rng_moves(seed, env0+e, tick, 5)); entries of other agents are kept, dead agents get IDLE */
void pom_oracle_simple_moves_batch(const pom_state* S, const uint8_t* status, long n, pom_simple_agent* A,
                                   uint64_t seed, uint64_t env0, uint32_t tick, unsigned agent_mask, uint8_t* moves);
/* strategy helpers for the reference's [strategy] known-answer tests */
int  pom_oracle_is_adjacent_enemy(const pom_state* s, int id, int distance);           /* strategy.cpp:296-312 */
int  pom_oracle_is_in_danger(const pom_state* s, int x, int y);                        /* strategy.cpp:225-246 */
void pom_oracle_fill_rmap(const pom_state* s, int id, int32_t map_out[POM_BOARD_CELLS]); /* strategy.cpp:58-95  */
/* kind 0: MoveTowardsPosition(a,b)  1: MoveTowardsPowerup(radius a)  2: MoveTowardsEnemy(radius a)
 * 3: MoveTowardsSafePlace(radius a)  (strategy.cpp:101-183) */
int  pom_oracle_move_towards(const pom_state* s, int id, int kind, int a, int b);

#ifdef __cplusplus
}
#endif
#endif
