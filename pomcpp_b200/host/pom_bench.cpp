/*
 * pom_bench.cpp — the repo's own C++ host driver of the hot path (north star: "called from the repo's own
 * C++ host code").  Single process, one host thread + one pom_batch handle + one stream per GPU; envs are
 * sharded contiguously over the GPUs with no collective on the data path; the only collective is one
 * ncclAllReduce of the episode counters after the run.
 *
 *   pom_bench [--gpus N] [--envs-per-gpu E] [--steps K] [--warmup W] [--mode step|rollout|expand|host|hostc] [--ticks T] [--simple MASK]
 *
 * mode step    : K launches of the per-tick kernel (pom_batch_step, auto-reset), moves pre-generated on device
 * mode rollout : K launches of the fused kernel (pom_batch_rollout), T ticks each, in-kernel RNG + auto-reset
 * mode expand  : tree-search expansion (BASELINE config 5): --roots R root states (taken from a 16-tick pre-roll)
 *                x 6^4 joint actions, one Step each, K repetitions of pom_batch_expand_step (GPU 0 only)
 * mode hostc   : mode host with compact I/O (pom_batch_step_compact: uint16 joint actions in, done bits + finished-env list out)
 * mode host    : end to end from host buffers (GPU 0): the envs as two half-batches stepped alternately with
 *                pom_batch_step_host_async + pom_batch_sync from pinned move / status buffers (what bench.py reports as e2e)
 * --no-overlap  : mode step without POM_STEP_OVERLAP (one launch per tick; default: two half-batch launches on two streams)
 * --simple MASK : the agents in MASK (bit a) are played by the device-side SimpleAgent (pom_batch_policy_moves before
 *                every step / POM_ROLL_SIMPLE in the rollout); 15 = the reference's own benchmark setting
 * Prints one JSON line.  (The CPU reference baseline is reported by bench.py, which alone may load oracle/.)
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#ifdef POM_WITH_NCCL
#include <nccl.h>
#endif

#include "pom_batch.h"

namespace
{

struct Args { int gpus = 1; uint64_t envs = 1u << 20; int steps = 200; int warmup = 10; std::string mode = "step"; uint32_t ticks = 800; uint64_t roots = 4096; uint32_t simple = 0; bool overlap = true; };

struct Shard { pom_batch* h = nullptr; void* moves = nullptr; float ms = 0.f; pom_stats stats; int rc = 0; std::string err; };

void die(const char* what) { std::fprintf(stderr, "pom_bench: %s: %s\n", what, pom_last_error()); std::exit(2); }

}

int main(int argc, char** argv)
{
    Args a;
    for(int i = 1; i < argc; i++)
    {
        std::string k = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : "0"; };
        if(k == "--gpus") a.gpus = std::atoi(next());
        else if(k == "--envs-per-gpu") a.envs = std::strtoull(next(), nullptr, 10);
        else if(k == "--steps") a.steps = std::atoi(next());
        else if(k == "--warmup") a.warmup = std::atoi(next());
        else if(k == "--mode") a.mode = next();
        else if(k == "--ticks") a.ticks = uint32_t(std::atoi(next()));
        else if(k == "--roots") a.roots = std::strtoull(next(), nullptr, 10);
        else if(k == "--simple") a.simple = uint32_t(std::atoi(next())) & 0xFu;
        else if(k == "--no-overlap") a.overlap = false;
    }
    if(a.mode == "expand")
    {
        /* config 5: R roots x 1296 joint actions, clone + one Step fused */
        pom_batch* src = nullptr; pom_batch* dst = nullptr;
        pom_init_desc d;
        std::memset(&d, 0, sizeof(d));
        d.n_templates = 1024; d.first_seed = 0x1337;
        if(pom_batch_init(&src, 0, a.roots, &d)) die("pom_batch_init(src)");
        if(pom_batch_rollout(src, 16, 7, 0, POM_ROLL_NO_RESET)) die("preroll");
        d.flags = POM_INIT_EMPTY; d.n_templates = 1;
        if(pom_batch_init(&dst, 0, a.roots * 1296, &d)) die("pom_batch_init(dst)");
        std::vector<uint32_t> idx(a.roots);
        for(uint64_t i = 0; i < a.roots; i++) idx[i] = uint32_t(i);
        for(int w = 0; w < a.warmup; w++) if(pom_batch_expand_step(dst, src, idx.data(), a.roots, 1296, 0)) die("expand");
        if(pom_batch_sync(dst)) die("sync");
        auto t0 = std::chrono::steady_clock::now();
        for(int k = 0; k < a.steps; k++) if(pom_batch_expand_step(dst, src, idx.data(), a.roots, 1296, 0)) die("expand");
        if(pom_batch_sync(dst)) die("sync");                         /* the calls only enqueue */
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const double children = double(a.roots) * 1296.0 * a.steps;
        std::printf("{\"metric\": \"env-steps/sec\", \"mode\": \"expand\", \"value\": %.6g, \"unit\": \"children (clone+Step)/s\", \"roots\": %llu, "
                    "\"fanout\": 1296, \"steps\": %d, \"ms_per_step\": %.6g, \"hbm_write_gbs\": %.6g}\n",
                    children / s, (unsigned long long)a.roots, a.steps, 1e3 * s / a.steps, children * 292.0 / s / 1e9);
        pom_batch_destroy(src); pom_batch_destroy(dst);
        return 0;
    }
    if(a.mode == "host")
    {
        /* every tick: 4 move bytes/env host -> device, 1 status byte/env device -> host, and the host waits for them */
        const uint64_t h = a.envs / 2;
        const int ring = 16;
        pom_batch* B[2] = { nullptr, nullptr };
        uint8_t* mv[2] = { nullptr, nullptr }; uint8_t* st[2] = { nullptr, nullptr };
        for(int i = 0; i < 2; i++)
        {
            pom_init_desc d;
            std::memset(&d, 0, sizeof(d));
            d.env_offset = uint64_t(i) * h; d.n_templates = 4096; d.first_seed = 0x1337; d.max_ticks = 800;
            if(pom_batch_init(&B[i], 0, h, &d)) die("pom_batch_init");
            if(pom_batch_rollout(B[i], 96, 20240229, 0, 0)) die("preroll");
            if(pom_host_alloc(4 * h * ring, reinterpret_cast<void**>(&mv[i])) || pom_host_alloc(h, reinterpret_cast<void**>(&st[i]))) die("pom_host_alloc");
            for(int t = 0; t < ring; t++)
                for(uint64_t e = 0; e < h; e++)
                {
                    const uint32_t m = pom_rng_moves(77, uint64_t(i) * h + e, uint32_t(t), 6);
                    std::memcpy(mv[i] + (size_t(t) * h + e) * 4, &m, 4);
                }
            if(pom_batch_sync(B[i])) die("sync");
        }
        unsigned long long done_seen = 0;
        auto run = [&](int steps)
        {
            if(pom_batch_step_host_async(B[0], mv[0], st[0], POM_STEP_AUTORESET)) die("step_host_async");
            for(int k = 0; k < steps; k++)
            {
                if(pom_batch_step_host_async(B[1], mv[1] + size_t(k % ring) * h * 4, st[1], POM_STEP_AUTORESET)) die("step_host_async");
                if(pom_batch_sync(B[0])) die("sync");
                done_seen += st[0][size_t(k) % h] & 1u;                 /* the host reads the results */
                if(k + 1 < steps && pom_batch_step_host_async(B[0], mv[0] + size_t((k + 1) % ring) * h * 4, st[0], POM_STEP_AUTORESET)) die("step_host_async");
                if(pom_batch_sync(B[1])) die("sync");
                done_seen += st[1][size_t(k) % h] & 1u;
            }
        };
        run(a.warmup);
        auto t0 = std::chrono::steady_clock::now();
        run(a.steps);
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("{\"metric\": \"env-steps/sec\", \"mode\": \"host\", \"value\": %.6g, \"unit\": \"env-steps/s\", \"envs\": %llu, \"steps\": %d, "
                    "\"us_per_step\": %.4g, \"h2d_bytes_per_step\": %llu, \"d2h_bytes_per_step\": %llu, \"done_flags_sampled\": %llu}\n",
                    double(2 * h) * a.steps / s, (unsigned long long)(2 * h), a.steps, 1e6 * s / a.steps,
                    (unsigned long long)(8 * h), (unsigned long long)(2 * h), done_seen);
        for(int i = 0; i < 2; i++) { pom_batch_destroy(B[i]); pom_host_free(mv[i]); pom_host_free(st[i]); }
        return 0;
    }
    if(a.mode == "hostc")
    {
        /* the same loop with compact I/O (pom_batch_step_compact): per env a uint16 joint action in; a done bit per env and
         * the list of finished envs (index + status byte) out; the host waits for every half-batch every tick and reads
         * the list.  Pinned buffers on the GPU's NUMA node, the calling thread bound there too. */
        const uint64_t h = a.envs / 2;
        const int ring = 16;
        pom_bind_thread_near(0);
        pom_batch* B[2] = { nullptr, nullptr };
        uint16_t* jt[2]; uint32_t* bits[2]; uint32_t* fenv[2]; uint8_t* fst[2]; uint32_t* fcnt[2];
        pom_step_compact_io io[2][16];
        for(int i = 0; i < 2; i++)
        {
            pom_init_desc d;
            std::memset(&d, 0, sizeof(d));
            d.env_offset = uint64_t(i) * h; d.n_templates = 4096; d.first_seed = 0x1337; d.max_ticks = 800;
            if(pom_batch_init(&B[i], 0, h, &d)) die("pom_batch_init");
            if(pom_batch_rollout(B[i], 96, 20240229, 0, 0)) die("preroll");
            if(pom_host_alloc_near(0, 2 * h * ring, reinterpret_cast<void**>(&jt[i])) ||
               pom_host_alloc_near(0, 4 * ((h + 31) / 32), reinterpret_cast<void**>(&bits[i])) ||
               pom_host_alloc_near(0, 4 * h, reinterpret_cast<void**>(&fenv[i])) ||
               pom_host_alloc_near(0, h, reinterpret_cast<void**>(&fst[i])) ||
               pom_host_alloc_near(0, 64, reinterpret_cast<void**>(&fcnt[i]))) die("pom_host_alloc_near");
            for(int t = 0; t < ring; t++)
                for(uint64_t e = 0; e < h; e++)
                {
                    const uint32_t m = pom_rng_moves(77, uint64_t(i) * h + e, uint32_t(t), 6);
                    jt[i][size_t(t) * h + e] = uint16_t((m & 0xFF) + 6 * ((m >> 8) & 0xFF) + 36 * ((m >> 16) & 0xFF) + 216 * (m >> 24));
                }
            for(int t = 0; t < ring; t++)
            {
                std::memset(&io[i][t], 0, sizeof(pom_step_compact_io));
                io[i][t].joint = jt[i] + size_t(t) * h; io[i][t].done_bits = bits[i]; io[i][t].fin_env = fenv[i];
                io[i][t].fin_status = fst[i]; io[i][t].fin_count = fcnt[i]; io[i][t].fin_capacity = uint32_t(h);
            }
            if(pom_batch_sync(B[i])) die("sync");
        }
        unsigned long long fin_seen = 0, wins0 = 0;
        const uint32_t fl = POM_STEP_AUTORESET | POM_STEP_COUNT;
        auto consume = [&](int i)
        {
            const uint32_t n = *fcnt[i];
            fin_seen += n;
            for(uint32_t k = 0; k < n; k += 64) wins0 += (fst[i][k] >> 2) & 1u;      /* the host reads the results */
        };
        auto run = [&](int steps)
        {
            if(pom_batch_step_compact(B[0], &io[0][0], fl)) die("step_compact");
            for(int k = 0; k < steps; k++)
            {
                if(pom_batch_step_compact(B[1], &io[1][k % ring], fl)) die("step_compact");
                if(pom_batch_sync(B[0])) die("sync");
                consume(0);
                if(k + 1 < steps && pom_batch_step_compact(B[0], &io[0][(k + 1) % ring], fl)) die("step_compact");
                if(pom_batch_sync(B[1])) die("sync");
                consume(1);
            }
        };
        run(a.warmup);
        fin_seen = 0;
        auto t0 = std::chrono::steady_clock::now();
        run(a.steps);
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("{\"metric\": \"env-steps/sec\", \"mode\": \"hostc\", \"value\": %.6g, \"unit\": \"env-steps/s\", \"envs\": %llu, \"steps\": %d, "
                    "\"us_per_step\": %.4g, \"h2d_bytes_per_step\": %llu, \"d2h_bytes_per_step\": %llu, \"finished_envs_seen\": %llu}\n",
                    double(2 * h) * a.steps / s, (unsigned long long)(2 * h), a.steps, 1e6 * s / a.steps,
                    (unsigned long long)(4 * h), (unsigned long long)(2 * (4 * ((h + 31) / 32) + 4) + 5 * fin_seen / (unsigned long long)a.steps), fin_seen);
        (void)wins0;
        for(int i = 0; i < 2; i++)
        {
            pom_batch_destroy(B[i]);
            pom_host_free(jt[i]); pom_host_free(bits[i]); pom_host_free(fenv[i]); pom_host_free(fst[i]); pom_host_free(fcnt[i]);
        }
        return 0;
    }
    if(a.gpus < 1 || a.gpus > pom_device_count()) { std::fprintf(stderr, "pom_bench: %d GPUs requested, %d present\n", a.gpus, pom_device_count()); return 2; }
    const bool rollout = a.mode == "rollout";
    const int ring = 64;
    const uint64_t seed = 20240229;
    std::vector<Shard> sh(size_t(a.gpus));

    /* set-up: one handle per GPU, shard g owns global envs [g*E, (g+1)*E) */
    for(int g = 0; g < a.gpus; g++)
    {
        pom_init_desc d;
        std::memset(&d, 0, sizeof(d));
        d.env_offset = uint64_t(g) * a.envs;
        d.n_templates = 4096;
        d.first_seed = 0x1337;
        d.max_ticks = 800;
        if(pom_batch_init(&sh[size_t(g)].h, g, a.envs, &d)) die("pom_batch_init");
        if(!rollout)
        {
            if(pom_device_alloc(g, 4 * a.envs * ring, &sh[size_t(g)].moves)) die("pom_device_alloc");
            for(int t = 0; t < ring; t++)
                if(pom_batch_generate_moves(sh[size_t(g)].h, static_cast<uint8_t*>(sh[size_t(g)].moves) + 4 * a.envs * size_t(t), seed, 100000u + uint32_t(t), 6)) die("generate_moves");
            if(pom_batch_rollout(sh[size_t(g)].h, 96, seed, 0, 0)) die("preroll");
        }
        if(pom_batch_sync(sh[size_t(g)].h)) die("sync");
        pom_batch_clear_stats(sh[size_t(g)].h);
    }

    /* timed region: one host thread per GPU, CUDA events on each handle's stream */
    auto work = [&](int g)
    {
        Shard& s = sh[size_t(g)];
        auto one = [&](int k) -> int
        {
            if(rollout) return pom_batch_rollout(s.h, a.ticks, seed, uint32_t(k) * a.ticks, POM_ROLL_SIMPLE(a.simple));
            if(a.simple)
            {
                const int rc = pom_batch_policy_moves(s.h, static_cast<uint8_t*>(s.moves) + 4 * a.envs * size_t(k % ring), seed, uint32_t(k), a.simple);
                if(rc) return rc;
            }
            return pom_batch_step(s.h, static_cast<uint8_t*>(s.moves) + 4 * a.envs * size_t(k % ring), POM_STEP_AUTORESET | POM_STEP_COUNT | (a.overlap && !a.simple ? POM_STEP_OVERLAP : 0));
        };
        for(int w = 0; w < a.warmup && !s.rc; w++) s.rc = one(w);
        if(!s.rc) s.rc = pom_batch_sync(s.h);
        if(!s.rc) s.rc = pom_batch_event_record(s.h, 0);
        for(int k = 0; k < a.steps && !s.rc; k++) s.rc = one(a.warmup + k);
        if(!s.rc) s.rc = pom_batch_event_record(s.h, 1);
        if(!s.rc) s.rc = pom_batch_event_elapsed_ms(s.h, &s.ms);
        if(s.rc) s.err = pom_last_error();
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for(int g = 0; g < a.gpus; g++) th.emplace_back(work, g);
    for(auto& t : th) t.join();
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for(int g = 0; g < a.gpus; g++) if(sh[size_t(g)].rc) { std::fprintf(stderr, "pom_bench: GPU %d: %s\n", g, sh[size_t(g)].err.c_str()); return 2; }

    /* the one collective: sum the episode counters over the GPUs */
    unsigned long long total[POM_STATS_WORDS] = {0};
    const char* reduced = "host sum";
#ifdef POM_WITH_NCCL
    if(a.gpus > 1)
    {
        std::vector<ncclComm_t> comms(size_t(a.gpus));
        std::vector<int> devs(size_t(a.gpus));
        for(int g = 0; g < a.gpus; g++) devs[size_t(g)] = g;
        if(ncclCommInitAll(comms.data(), a.gpus, devs.data()) == ncclSuccess)
        {
            ncclGroupStart();
            for(int g = 0; g < a.gpus; g++)
            {
                void* p = pom_batch_stats_device_ptr(sh[size_t(g)].h);
                ncclAllReduce(p, p, POM_STATS_WORDS, ncclUint64, ncclSum, comms[size_t(g)], static_cast<cudaStream_t>(pom_batch_stream(sh[size_t(g)].h)));
            }
            ncclGroupEnd();
            for(int g = 0; g < a.gpus; g++) pom_batch_sync(sh[size_t(g)].h);
            pom_stats s;
            pom_batch_stats(sh[0].h, &s);
            std::memcpy(total, &s, sizeof(total));
            for(auto c : comms) ncclCommDestroy(c);
            reduced = "ncclAllReduce";
        }
    }
#endif
    if(std::strcmp(reduced, "host sum") == 0)
    {
        for(int g = 0; g < a.gpus; g++)
        {
            pom_stats s;
            pom_batch_stats(sh[size_t(g)].h, &s);
            const unsigned long long* w = reinterpret_cast<const unsigned long long*>(&s);
            for(int i = 0; i < POM_STATS_WORDS; i++) total[i] += w[i];
        }
    }

    float ms_max = 0.f;
    for(int g = 0; g < a.gpus; g++) ms_max = sh[size_t(g)].ms > ms_max ? sh[size_t(g)].ms : ms_max;
    const double env_steps_timed = double(a.gpus) * double(a.envs) * double(a.steps) * (rollout ? double(a.ticks) : 1.0);
    const double value = env_steps_timed / (double(ms_max) * 1e-3);
    std::printf("{\"metric\": \"env-steps/sec\", \"value\": %.6g, \"unit\": \"env-steps/s\", \"n_gpus\": %d, \"mode\": \"%s\", \"simple_agent_mask\": %u, "
                "\"envs_per_gpu\": %llu, \"steps\": %d, \"warmup\": %d, \"ticks_per_step\": %u, \"ms_per_step\": %.6g, \"wall_s\": %.4g, "
                "\"hbm_gbs_algorithmic_per_gpu\": %.6g, \"episode_stats\": {\"env_steps\": %llu, \"episodes\": %llu, \"wins\": [%llu, %llu, %llu, %llu], "
                "\"draws\": %llu, \"truncated\": %llu, \"sum_episode_len\": %llu, \"invalid\": %llu, \"reduced_with\": \"%s\"}}\n",
                value, a.gpus, a.mode.c_str(), a.simple, (unsigned long long)a.envs, a.steps, a.warmup, rollout ? a.ticks : 1u, double(ms_max) / a.steps, wall,
                rollout ? 0.0 : 582.0 * double(a.envs) / (double(ms_max) / a.steps * 1e-3) / 1e9,
                total[0], total[1], total[2], total[3], total[4], total[5], total[6], total[7], total[8], total[9], reduced);
    for(int g = 0; g < a.gpus; g++)
    {
        if(sh[size_t(g)].moves) pom_device_free(g, sh[size_t(g)].moves);
        pom_batch_destroy(sh[size_t(g)].h);
    }
    return 0;
}
