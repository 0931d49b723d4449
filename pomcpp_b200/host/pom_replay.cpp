/*
 * pom_replay.cpp — re-runs a POMTRC1 trace (pom_trace.hpp) on the GPU through the C ABI and checks the final
 * state hashes; optionally renders one env tick by tick.
 *
 *   pom_replay trace.pomtrc [--print ENV] [--device D]
 *
 * Exit code 0 = every final state hash and status byte matches the file.
 */
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "pom_batch.h"
#include "pom_trace.hpp"

int main(int argc, char** argv)
{
    if(argc < 2) { std::fprintf(stderr, "usage: pom_replay trace.pomtrc [--print ENV] [--device D]\n"); return 2; }
    long print_env = -1;
    int device = 0;
    for(int i = 2; i + 1 < argc; i += 2)
    {
        if(std::string(argv[i]) == "--print") print_env = std::atol(argv[i + 1]);
        else if(std::string(argv[i]) == "--device") device = std::atoi(argv[i + 1]);
    }
    pomtrace::Trace t;
    std::string err;
    if(!pomtrace::read(argv[1], t, err)) { std::fprintf(stderr, "pom_replay: %s\n", err.c_str()); return 2; }

    pom_batch* h = nullptr;
    pom_init_desc d = {};
    d.n_templates = 1;
    d.first_seed = 0x1337;
    d.flags = POM_INIT_EMPTY;
    if(pom_batch_init(&h, device, t.n_envs, &d)) { std::fprintf(stderr, "pom_replay: %s\n", pom_last_error()); return 2; }
    if(pom_batch_upload(h, 0, t.n_envs, t.initial.data(), nullptr)) { std::fprintf(stderr, "pom_replay: %s\n", pom_last_error()); return 2; }
    const uint32_t step_flags = (t.flags & 1u) ? uint32_t(POM_STEP_RAW) : 0u;
    pom_state one;
    uint8_t st1 = 0;
    if(print_env >= 0 && print_env < long(t.n_envs))
        std::fputs(pomtrace::render(t.initial[size_t(print_env)], 0).c_str(), stdout);
    for(uint32_t k = 0; k < t.n_ticks; k++)
    {
        if(pom_batch_step_host(h, t.moves.data() + size_t(k) * t.n_envs * 4, nullptr, step_flags))
        { std::fprintf(stderr, "pom_replay: tick %u: %s\n", k, pom_last_error()); return 2; }
        if(print_env >= 0 && print_env < long(t.n_envs))
        {
            pom_batch_download(h, uint64_t(print_env), 1, &one, &st1);
            const uint8_t* m = t.moves.data() + (size_t(k) * t.n_envs + size_t(print_env)) * 4;
            std::printf("--- moves %d %d %d %d\n", m[0], m[1], m[2], m[3]);
            std::fputs(pomtrace::render(one, st1).c_str(), stdout);
        }
    }
    std::vector<pom_state> fin(t.n_envs);
    std::vector<uint8_t> st(t.n_envs);
    if(pom_batch_download(h, 0, t.n_envs, fin.data(), st.data())) { std::fprintf(stderr, "pom_replay: %s\n", pom_last_error()); return 2; }
    uint32_t bad = 0;
    for(uint32_t e = 0; e < t.n_envs; e++)
    {
        if(pomtrace::hash_state(fin[e]) != t.final_hash[e] || ((st[e] ^ t.final_status[e]) & 0x1F))
        {
            if(bad < 5) std::printf("MISMATCH env %u\n", e);
            bad++;
        }
    }
    std::printf("pom_replay: %u envs x %u ticks, %u mismatches\n", t.n_envs, t.n_ticks, bad);
    pom_batch_destroy(h);
    return bad ? 1 : 0;
}
