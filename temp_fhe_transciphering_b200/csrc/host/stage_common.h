// stage_common.h — helpers shared by the drop-in stage executables.
// File contract: SURVEY.md 8(b)(1); size -> directory name: submission/src/help_fun.rs:3-10.
#pragma once
#include "cbs_b200.h"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
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

#define STAGE_TRY(expr)                                                   \
    do {                                                                  \
        int _rc = (expr);                                                 \
        if (_rc != CBS_OK) {                                              \
            fprintf(stderr, "Error: %s: %s\n", #expr, cbs_last_error());  \
            return 1;                                                     \
        }                                                                 \
    } while (0)
