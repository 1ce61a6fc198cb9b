// client_key_generation <size> [seed] — seeded stand-in for the reference's (unseeded) client key
// generator (submission/src/bin/client_key_generation.rs:89-132): writes the same six files under
// io/<s>/{secret_keys,public_keys}/ in the same bincode layout, so the reference's own server and
// client binaries accept them.  Client-side helper for reproducible tests and benches; the hot
// path never needs it.
#include "stage_common.h"

int main(int argc, char **argv)
{
    long size;
    if (!parse_size(argc, argv, &size)) return 1;
    uint64_t seed = 1;
    if (argc > 2) seed = strtoull(argv[2], nullptr, 10);
    else if (const char *e = getenv("CBS_SEED")) seed = strtoull(e, nullptr, 10);
    const std::string io_dir = std::string("io/") + size_string(size);
    cbs_keyset *ks = nullptr;
    STAGE_TRY(cbs_keyset_generate(seed, &ks));
    STAGE_TRY(cbs_keyset_save_dir(ks, io_dir.c_str(), 1));
    cbs_keyset_free(ks);
    return 0;
}
