// client_key_generation <size> [seed] — stand-in for the reference's client key generator
// (submission/src/bin/client_key_generation.rs:89-132): writes the same six files under
// io/<s>/{secret_keys,public_keys}/ in the same bincode layout, so the reference's own server and
// client binaries accept them.  Without a seed (no second argument, no CBS_SEED) all randomness comes
// from ChaCha20 keyed by getrandom(2), like the reference's new_seeder() (:91-96): fresh keys on every
// run.  An explicit seed selects the deterministic, NOT cryptographically secure test generator
// (reproducible tests, benches and golden vectors only).  Client-side; the hot path never needs it.
#include "stage_common.h"

int main(int argc, char **argv)
{
    long size;
    if (!parse_size(argc, argv, &size)) return 1;
    const char *seed_arg = argc > 2 ? argv[2] : getenv("CBS_SEED");
    const std::string io_dir = std::string("io/") + size_string(size);
    cbs_keyset *ks = nullptr;
    if (seed_arg) STAGE_TRY(cbs_keyset_generate(strtoull(seed_arg, nullptr, 10), &ks));
    else STAGE_TRY(cbs_keyset_generate_os_entropy(&ks));
    STAGE_TRY(cbs_keyset_save_dir(ks, io_dir.c_str(), 1));
    cbs_keyset_free(ks);
    return 0;
}
