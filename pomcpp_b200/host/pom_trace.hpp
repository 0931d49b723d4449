/*
 * pom_trace.hpp — on-disk trace format + a console renderer for single envs.
 *
 * The reference has no wire/disk format and no golden traces; its only renderer is PrintState
 * (src/bboard/bboard.cpp:403-489, ANSI colours).  A trace file pins a batch of games completely:
 *
 *   offset  size                     field
 *   0       8                        magic "POMTRC1\0"
 *   8       4                        n_envs
 *   12      4                        n_ticks
 *   16      4                        flags (bit 0: moves are POM_STEP_RAW ticks, i.e. bare bboard::Step)
 *   20      4                        reserved (0)
 *   24      n_envs * 1004            initial states, AoS `pom_state` (= reference bboard::State layout)
 *   ...     n_ticks * n_envs * 4     joint actions, tick-major, one byte per agent (Move, bboard.hpp:35-43)
 *   ...     n_envs * 8               FNV-1a hash of every final state (pom_trace_hash below)
 *   ...     n_envs                   final status bytes (POM_STATUS_*)
 *
 * All integers little-endian.  tests/golden/make_golden.py writes such files from the COMPILED REFERENCE;
 * pom_replay re-runs them on the GPU and checks the hashes.
 */
#ifndef POM_TRACE_HPP_
#define POM_TRACE_HPP_

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "pom_state.h"

namespace pomtrace
{

struct Trace {
    uint32_t n_envs = 0, n_ticks = 0, flags = 0;
    std::vector<pom_state> initial;
    std::vector<uint8_t> moves;         /* [tick][env][4] */
    std::vector<uint64_t> final_hash;
    std::vector<uint8_t> final_status;
};

/* FNV-1a over the meaningful fields of a state (the two padding bytes of each agent are skipped);
 * identical to the checker's pom_oracle_state_hash so that files written from either side agree */
inline uint64_t hash_state(const pom_state& s)
{
    uint64_t h = 1469598103934665603ull;
    auto mix = [&h](const void* p, size_t n)
    {
        const uint8_t* c = static_cast<const uint8_t*>(p);
        for(size_t i = 0; i < n; i++) { h ^= c[i]; h *= 1099511628211ull; }
    };
    mix(s.board, sizeof(s.board));
    mix(&s.timeStep, 8);
    for(int a = 0; a < 4; a++)
    {
        mix(&s.agents[a], 20);
        const uint8_t f[2] = { uint8_t(s.agents[a].canKick != 0), uint8_t(s.agents[a].dead != 0) };
        mix(f, 2);
    }
    mix(s.bombs, sizeof(s.bombs) + 8);
    mix(s.flames, sizeof(s.flames) + 8);
    return h;
}

inline bool read(const std::string& path, Trace& t, std::string& err)
{
    FILE* f = std::fopen(path.c_str(), "rb");
    if(!f) { err = "cannot open " + path; return false; }
    char magic[8];
    uint32_t hdr[4];
    bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "POMTRC1\0", 8) == 0 && std::fread(hdr, 4, 4, f) == 4;
    if(ok)
    {
        t.n_envs = hdr[0]; t.n_ticks = hdr[1]; t.flags = hdr[2];
        t.initial.resize(t.n_envs);
        t.moves.resize(size_t(t.n_ticks) * t.n_envs * 4);
        t.final_hash.resize(t.n_envs);
        t.final_status.resize(t.n_envs);
        ok = std::fread(t.initial.data(), sizeof(pom_state), t.n_envs, f) == t.n_envs &&
             std::fread(t.moves.data(), 1, t.moves.size(), f) == t.moves.size() &&
             std::fread(t.final_hash.data(), 8, t.n_envs, f) == t.n_envs &&
             std::fread(t.final_status.data(), 1, t.n_envs, f) == t.n_envs;
    }
    std::fclose(f);
    if(!ok) err = "not a POMTRC1 file or truncated: " + path;
    return ok;
}

inline bool write(const std::string& path, const Trace& t)
{
    FILE* f = std::fopen(path.c_str(), "wb");
    if(!f) return false;
    const uint32_t hdr[4] = { t.n_envs, t.n_ticks, t.flags, 0 };
    bool ok = std::fwrite("POMTRC1\0", 1, 8, f) == 8 && std::fwrite(hdr, 4, 4, f) == 4 &&
              std::fwrite(t.initial.data(), sizeof(pom_state), t.n_envs, f) == t.n_envs &&
              std::fwrite(t.moves.data(), 1, t.moves.size(), f) == t.moves.size() &&
              std::fwrite(t.final_hash.data(), 8, t.n_envs, f) == t.n_envs &&
              std::fwrite(t.final_status.data(), 1, t.n_envs, f) == t.n_envs;
    std::fclose(f);
    return ok;
}

/* one env as 11 text rows + a status line (plain ASCII; the reference's PrintState uses ANSI colours):
 *   .  passage   #  rigid   +  wood   Q  bomb   *  flames   b r k  powerups (extra bomb / range / kick)
 *   0-3 agents   ?  anything else */
inline std::string render(const pom_state& s, uint8_t status)
{
    std::string out;
    for(int y = 0; y < POM_BOARD_SIZE; y++)
    {
        for(int x = 0; x < POM_BOARD_SIZE; x++)
        {
            const int v = s.board[y][x];
            char c = '?';
            if(v == POM_ITEM_PASSAGE) c = '.';
            else if(v == POM_ITEM_RIGID) c = '#';
            else if((v >> 8) == 2) c = '+';
            else if(v == POM_ITEM_BOMB) c = 'Q';
            else if((v >> 16) == 4) c = '*';
            else if(v == POM_ITEM_EXTRABOMB) c = 'b';
            else if(v == POM_ITEM_INCRRANGE) c = 'r';
            else if(v == POM_ITEM_KICK) c = 'k';
            else if(v >= POM_ITEM_AGENT0 && v <= POM_ITEM_AGENT0 + 3) c = char('0' + (v - POM_ITEM_AGENT0));
            out += c;
            out += ' ';
        }
        out += '\n';
    }
    char line[160];
    std::snprintf(line, sizeof(line), "t=%d alive=%d bombs=%d flames=%d status=0x%02x%s\n", s.timeStep, s.aliveAgents,
                  s.bombs_count, s.flames_count, status, (status & POM_STATUS_DONE) ? " DONE" : "");
    out += line;
    for(int a = 0; a < 4; a++)
    {
        const pom_agent& g = s.agents[a];
        std::snprintf(line, sizeof(line), "  agent %d: (%d,%d) bombs %d/%d range %d%s%s\n", a, g.x, g.y, g.bombCount, g.maxBombCount,
                      g.bombStrength, g.canKick ? " kick" : "", g.dead ? " DEAD" : "");
        out += line;
    }
    return out;
}

}

#endif
