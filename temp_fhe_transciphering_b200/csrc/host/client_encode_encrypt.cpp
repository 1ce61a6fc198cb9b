// client_encode_encrypt <size> [seed] — stand-in for submission/src/bin/client_encode_encrypt.rs (without a seed argument
// or CBS_SEED: ChaCha20 keyed by getrandom(2); with one: the deterministic test generator):
// reads datasets/<s>/aes_key.hex and io/<s>/secret_keys/glwe_sk.bin, writes
// io/<s>/ciphertexts_upload/trans_key.bin (AllRdKeys, src/data_struct.rs:11-26).  Size 0 (ECB) writes the
// reference's inverse-direction keys; sizes 1/2 (CTR in the harness) write the forward-direction keys.
#include "stage_common.h"

int main(int argc, char **argv)
{
    long size;
    if (!parse_size(argc, argv, &size)) return 1;
    bool seeded = false;
    uint64_t seed = 0;
    if (argc > 2) seed = strtoull(argv[2], nullptr, 10), seeded = true;
    else if (const char *e = getenv("CBS_SEED")) seed = strtoull(e, nullptr, 10) + 1, seeded = true;
    const std::string io_dir = std::string("io/") + size_string(size);
    const std::string data_dir = std::string("datasets/") + size_string(size);
    std::vector<uint8_t> key;
    if (!read_hex_file(data_dir + "/aes_key.hex", key) || key.size() != 16) {
        fprintf(stderr, "Error: cannot read %s/aes_key.hex\n", data_dir.c_str());
        return 1;
    }
    cbs_keyset *ks = nullptr;
    STAGE_TRY(cbs_keyset_load_dir(io_dir.c_str(), 1, &ks));
    const std::string out = io_dir + "/ciphertexts_upload/trans_key.bin";
    if (size == 0) {
        std::vector<uint64_t> k10_9(CBS_K10_9_WORDS), k8_1(CBS_K8_1_WORDS), k0(CBS_K0_WORDS);
        if (seeded) STAGE_TRY(cbs_trans_key_generate(ks, key.data(), seed, k10_9.data(), k8_1.data(), k0.data()));
        else STAGE_TRY(cbs_trans_key_generate_os_entropy(ks, key.data(), k10_9.data(), k8_1.data(), k0.data()));
        STAGE_TRY(cbs_trans_key_save(out.c_str(), k10_9.data(), k8_1.data(), k0.data()));
    } else {
        std::vector<uint64_t> kf(CBS_KF_FIRST_WORDS), km(CBS_KF_MID_WORDS), kl(CBS_KF_LAST_WORDS);
        if (seeded) STAGE_TRY(cbs_fwd_trans_key_generate(ks, key.data(), seed, kf.data(), km.data(), kl.data()));
        else STAGE_TRY(cbs_fwd_trans_key_generate_os_entropy(ks, key.data(), kf.data(), km.data(), kl.data()));
        STAGE_TRY(cbs_fwd_trans_key_save(out.c_str(), kf.data(), km.data(), kl.data()));
    }
    printf("Transciphering keys saved to %s/ciphertexts_upload\n", io_dir.c_str());
    cbs_keyset_free(ks);
    return 0;
}
