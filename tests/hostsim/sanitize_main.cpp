/*
 * sanitize_main.cpp — TEST-ONLY.  Runs the kernel body (pom_core.cuh, host build) and the plain-C restatement side by
 * side under AddressSanitizer + UndefinedBehaviorSanitizer on seeded random traces (random agents, the all-kick stress
 * regime, and games in which agents 1-3 are played by the device policy code, pom_policy.cuh, against the oracle's
 * SimpleAgent), comparing every field after every tick, and every 16th env-tick the observation planes of one agent with
 * the oracle's definition.  compute-sanitizer is not available on the GPU pool,
 * so this is the memory-safety check of the exact code the CUDA kernels run: every record access of the tick
 * happens inside a 292-byte heap block of its own, so an out-of-bounds ring / board / stack index trips ASan.
 *
 *   sanitize_main [n_envs] [ticks]        exit 0 = no sanitizer report, no mismatch
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pom_core.cuh"
#include "pom_policy.cuh"
extern "C" {
#include "pom_oracle.h"
}

int main(int argc, char** argv)
{
    const int n = argc > 1 ? std::atoi(argv[1]) : 512;
    const int ticks = argc > 2 ? std::atoi(argv[2]) : 200;
    long steps = 0, invalid = 0;
    for(int stress = 0; stress < 3; stress++)      /* 0 random agents, 1 all-kick stress, 2 SimpleAgent opponents (policy code + its oracle) */
    {
        const bool simple = stress == 2;
        std::vector<pom_simple_agent> A(static_cast<size_t>(4 * n)), B(static_cast<size_t>(4 * n));
        std::memset(A.data(), 0, A.size() * sizeof(pom_simple_agent));
        std::memset(B.data(), 0, B.size() * sizeof(pom_simple_agent));
        std::vector<pom_state> S(static_cast<size_t>(n)), T(static_cast<size_t>(n));
        std::vector<uint8_t*> recs(static_cast<size_t>(n));
        std::vector<uint8_t> st(static_cast<size_t>(n), 0);
        int seed = 0x1337;
        for(int e = 0; e < n; e++)
        {
            do { pom_oracle_zero_state(&S[size_t(e)]); } while(pom_oracle_init_state(&S[size_t(e)], seed++, 0, 1, 2, 3));
            if(stress == 1)
            {
                for(int a = 0; a < 4; a++) { S[size_t(e)].agents[a].canKick = 1; S[size_t(e)].agents[a].maxBombCount = 5; S[size_t(e)].agents[a].bombStrength = 4; }
                S[size_t(e)].bombs_index = (e * 3) % 20;       /* rings that wrap */
                S[size_t(e)].flames_index = (e * 7) % 20;
            }
            recs[size_t(e)] = static_cast<uint8_t*>(std::malloc(POM_REC_BYTES));   /* exact-size heap block: ASan guards both ends */
            if(pomcore::pack(&S[size_t(e)], 0, recs[size_t(e)])) { std::printf("pack failed\n"); return 1; }
        }
        const std::vector<pom_state> S0 = S;
        for(int t = 0; t < ticks; t++)
        {
            for(int e = 0; e < n; e++)
            {
                uint32_t m = pom_oracle_rng_moves(99 + uint64_t(stress), uint64_t(e), uint32_t(t), 6);
                uint8_t mv[4];
                std::memcpy(mv, &m, 4);
                if(simple && !(st[size_t(e)] & 0x11))
                {
                    /* agents 1-3: the device policy code on the packed record vs the oracle's SimpleAgent on the AoS state */
                    const uint32_t draws = pom_oracle_rng_moves(7, uint64_t(e), uint32_t(t), 5);
                    pompolicy::ArrayStore store{ reinterpret_cast<pompolicy::SimpleSt*>(&B[size_t(4 * e)]) };
                    m = pompolicy::simple_moves(recs[size_t(e)], 0xEu, m, draws, store);
                    for(int a = 1; a < 4; a++)
                        mv[a] = S[size_t(e)].agents[a].dead ? uint8_t(0) : uint8_t(pom_oracle_simple_act(&S[size_t(e)], a, &A[size_t(4 * e + a)], int((draws >> (8 * a)) & 0xFFu)));
                    uint32_t m2;
                    std::memcpy(&m2, mv, 4);
                    if(m2 != m || std::memcmp(&A[size_t(4 * e)], &B[size_t(4 * e)], 4 * sizeof(pom_simple_agent)))
                    {
                        std::printf("POLICY MISMATCH tick %d env %d\n", t, e);
                        return 1;
                    }
                }
                if(!(st[size_t(e)] & 0x11)) steps++;
                pom_oracle_env_step(&S[size_t(e)], &st[size_t(e)], mv);
                pomcore::env_step(recs[size_t(e)], m);
                const uint8_t st2 = pomcore::unpack(recs[size_t(e)], &T[size_t(e)]);
                if(pom_oracle_state_diff(&S[size_t(e)], &T[size_t(e)]) || st2 != st[size_t(e)])
                {
                    std::printf("MISMATCH stress %d tick %d env %d\n", stress, t, e);
                    return 1;
                }
                if(((t + e) & 15) == 0)
                {
                    /* the observation code on the packed record, into an exact-size block, against the definition */
                    const int agent = (t + e) >> 4 & 3, view = 1 + ((t >> 2) & 7);
                    uint8_t* obs = static_cast<uint8_t*>(std::aligned_alloc(32, 512));
                    uint8_t want[512];
                    pomcore::observe_planes(recs[size_t(e)], agent, view, obs);
                    pom_oracle_observe_planes(&S[size_t(e)], agent, view, want);
                    const bool same = std::memcmp(obs, want, 496) == 0;
                    std::free(obs);
                    if(!same) { std::printf("OBSERVATION MISMATCH stress %d tick %d env %d\n", stress, t, e); return 1; }
                }
                if(st[size_t(e)] & 0x11)
                {
                    if(st[size_t(e)] & 0x10) invalid++;
                    S[size_t(e)] = S0[size_t(e)];
                    st[size_t(e)] = 0;
                    std::memset(&A[size_t(4 * e)], 0, 4 * sizeof(pom_simple_agent));
                    std::memset(&B[size_t(4 * e)], 0, 4 * sizeof(pom_simple_agent));
                    pomcore::pack(&S[size_t(e)], 0, recs[size_t(e)]);
                }
            }
        }
        for(int e = 0; e < n; e++) std::free(recs[size_t(e)]);
    }
    /* the deepest explosion the rings allow: 20 bombs, each in range of the next, set off by the first (nesting depth 20
     * on the explosion machine's explicit stack) */
    for(int layout = 0; layout < 2; layout++)
    {
        pom_state s, t;
        pom_oracle_zero_state(&s);
        pom_oracle_kill(&s, 1); pom_oracle_kill(&s, 2); pom_oracle_kill(&s, 3);
        pom_oracle_put_agent(&s, 5, 5, 0);
        s.agents[0].maxBombCount = 30; s.agents[0].bombStrength = 2;
        for(int k = 0; k < 20; k++)
        {
            const int a = k < 11 ? k : 10, b = k < 11 ? 0 : k - 10;
            pom_oracle_plant_bomb(&s, layout ? b : a, layout ? a : b, 0, k ? 10 : 1, 1);
        }
        uint8_t* rec = static_cast<uint8_t*>(std::malloc(POM_REC_BYTES));
        if(pomcore::pack(&s, 0, rec)) { std::printf("pack failed\n"); return 1; }
        const uint8_t idle[4] = {0, 0, 0, 0};
        pom_oracle_step(&s, idle);
        pomcore::step(rec, 0u);
        pomcore::unpack(rec, &t);
        std::free(rec);
        if(pom_oracle_state_diff(&s, &t) || s.bombs_count != 0 || s.flames_count != 20) { std::printf("CHAIN MISMATCH layout %d\n", layout); return 1; }
    }
    std::printf("sanitize_main: %ld env-steps, %ld aborted episodes, clean\n", steps, invalid);
    return 0;
}
