// stage_common.h — helpers shared by the drop-in stage executables.
// File contract: SURVEY.md 8(b)(1); size -> directory name: submission/src/help_fun.rs:3-10.
#pragma once
#include "cbs_b200.h"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <unistd.h>
#include <vector>

inline const char *size_string(long size)
{
    switch (size) {
        case 0: return "toy";
        case 1: return "small";
        case 2: return "medium";
        default: return "unknown";
    }
}

inline bool parse_size(int argc, char **argv, long *size)
{
    if (argc < 2) {
        fprintf(stderr, "Usage: %s <size>\n", argv[0]);
        exit(1);  // same as the reference mains (server_encrypted_aes_decryption.rs:600-604)
    }
    char *end = nullptr;
    *size = strtol(argv[1], &end, 10);
    if (!end || *end != '\0' || *size < 0) {
        fprintf(stderr, "Error: invalid size argument '%s'\n", argv[1]);
        return false;
    }
    return true;
}

inline bool read_hex_file(const std::string &path, std::vector<uint8_t> &out)
{
    std::ifstream f(path);
    if (!f) return false;
    std::stringstream ss;
    ss << f.rdbuf();
    std::string s = ss.str();
    while (!s.empty() && isspace((unsigned char)s.back())) s.pop_back();
    size_t b = 0;
    while (b < s.size() && isspace((unsigned char)s[b])) b++;
    s = s.substr(b);
    if (s.size() % 2) return false;
    out.clear();
    for (size_t i = 0; i < s.size(); i += 2) {
        unsigned v;
        if (sscanf(s.c_str() + i, "%2x", &v) != 1) return false;
        out.push_back((uint8_t)v);
    }
    return true;
}

// How many GPUs a one-shot stage process should use, decided BEFORE the first CUDA call.  Driver start-up dominates these
// processes: on an 8 x B200 box cuInit + primary contexts cost 5.4 - 8.6 s with all eight GPUs visible against 0.8 - 3 s
// with one (profiles/r02_harness_{1,8}gpu.jsonl), while a GPU transciphers ~62 blocks per second.  Minimising
// t(g) = 0.85 g + blocks / (62 g) gives g = sqrt(blocks / 53): one GPU up to ~120 blocks, 4 - 5 for 1024.  The choice is made
// effective by narrowing CUDA_VISIBLE_DEVICES (to the first g entries of the caller's list if there is one; CBS_GPUS asks
// for a specific count), so that the driver never touches the other devices.  block_equivalents = AES blocks (stage 7) or values / 8 (stage 8).
inline void plan_visible_gpus(long block_equivalents)
{
    // the devices this process may use: the caller's CUDA_VISIBLE_DEVICES list if there is one, else /dev/nvidia<i>
    std::vector<std::string> avail;
    if (const char *cvd = getenv("CUDA_VISIBLE_DEVICES")) {
        std::string tok;
        for (const char *p = cvd;; p++) {
            if (*p == ',' || *p == '\0') {
                if (!tok.empty()) avail.push_back(tok);
                tok.clear();
                if (*p == '\0') break;
            } else if (!isspace((unsigned char)*p)) {
                tok.push_back(*p);
            }
        }
    } else {
        for (int i = 0; i < 64; i++)
            if (access(("/dev/nvidia" + std::to_string(i)).c_str(), F_OK) == 0) avail.push_back(std::to_string(avail.size()));
    }
    const int present = (int)avail.size();
    if (present <= 1) return;
    int want = 1;
    if (const char *e = getenv("CBS_GPUS")) {
        want = atoi(e);
    } else {
        while ((want + 0.5) * (want + 0.5) * 53.0 < (double)block_equivalents) want++;
    }
    want = want < 1 ? 1 : (want > present ? present : want);
    std::string list;
    for (int i = 0; i < want; i++) list += (i ? "," : "") + avail[(size_t)i];
    setenv("CUDA_VISIBLE_DEVICES", list.c_str(), 1);
    if (getenv("CBS_PLAN_DEBUG")) fprintf(stderr, "[plan] %ld block equivalents, %d device(s) available -> CUDA_VISIBLE_DEVICES=%s\n",
                                          block_equivalents, present, list.c_str());
}

// Wall-clock breakdown of a stage process (the harness only records the total per stage, harness/utils.py:85-109).
// CBS_STAGE_TIMING=1 prints one JSON line to stderr at exit, any other value appends it to that file.
struct StageClock {
    const char *dest = getenv("CBS_STAGE_TIMING");
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(), last = t0;
    std::string name, json;
    explicit StageClock(const char *stage) : name(stage) {}
    void mark(const char *what)
    {
        if (!dest) return;
        const auto now = std::chrono::steady_clock::now();
        char buf[96];
        snprintf(buf, sizeof buf, "%s\"%s_ms\": %.1f", json.empty() ? "" : ", ", what,
                 std::chrono::duration<double, std::milli>(now - last).count());
        json += buf;
        last = now;
    }
    void add(const char *what, double ms)
    {
        if (!dest) return;
        char buf[96];
        snprintf(buf, sizeof buf, "%s\"%s_ms\": %.1f", json.empty() ? "" : ", ", what, ms);
        json += buf;
    }
    ~StageClock()
    {
        if (!dest) return;
        char buf[96];
        snprintf(buf, sizeof buf, ", \"total_ms\": %.1f}\n",
                 std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        const std::string line = "{\"stage\": \"" + name + "\", " + json + buf;
        FILE *f = strcmp(dest, "1") == 0 ? stderr : fopen(dest, "a");
        if (f) {
            fputs(line.c_str(), f);
            if (f != stderr) fclose(f);
        }
    }
};

#define STAGE_TRY(expr)                                                   \
    do {                                                                  \
        int _rc = (expr);                                                 \
        if (_rc != CBS_OK) {                                              \
            fprintf(stderr, "Error: %s: %s\n", #expr, cbs_last_error());  \
            return 1;                                                     \
        }                                                                 \
    } while (0)
