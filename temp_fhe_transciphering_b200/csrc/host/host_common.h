// host_common.h — shared between the host-only translation units and the CUDA API layer.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

namespace cbs_host {
extern thread_local std::string g_last_error;
void set_error(const std::string &s);
}  // namespace cbs_host

// Host-side key material in the standard domain (flat layouts of include/cbs_b200.h).
struct cbs_keyset {
    std::vector<uint64_t> bsk, ksk, autok, ss, lwe_sk_small, glwe_sk;
};
