// client_decrypt_common.h — shared body of the two decrypt stages and the two postprocess stages
// (submission/src/bin/client_decrypt_decode{,_aes_decryption}.rs, client_postprocess{,_aes_decryption}.rs).
#pragma once
#include "stage_common.h"

// result.bin (LweCiphertextList) under io/<s>/<download_dir>/ -> bincode Vec<u64> of decrypted bits
inline int decrypt_stage(int argc, char **argv, const char *download_dir, const char *out_name)
{
    long size;
    if (!parse_size(argc, argv, &size)) return 1;
    const std::string io_dir = std::string("io/") + size_string(size);
    cbs_keyset *ks = nullptr;
    // only the secret key is needed, but the loader validates the whole directory the same way
    uint64_t *sk = nullptr, nsk = 0;
    {
        // LweSecretKey<Vec<u64>> = bincode Vec<u64>
        STAGE_TRY(cbs_u64_vec_load((io_dir + "/secret_keys/lwe_sk.bin").c_str(), &sk, &nsk));
    }
    (void)ks;
    uint64_t *lwe = nullptr, count = 0, words = 0;
    STAGE_TRY(cbs_lwe_list_load((io_dir + "/" + download_dir + "/result.bin").c_str(), &lwe, &count, &words));
    if (words != nsk + 1) {
        fprintf(stderr, "Error: ciphertext size %llu does not match the secret key (%llu)\n", (unsigned long long)words,
                (unsigned long long)nsk);
        return 1;
    }
    std::vector<uint64_t> bits(count);
    STAGE_TRY(cbs_lwe_decrypt_bits(sk, (int)nsk, lwe, count, bits.data()));
    STAGE_TRY(cbs_u64_vec_save((io_dir + "/intermediate/" + out_name).c_str(), bits.data(), count));
    cbs_free(sk);
    cbs_free(lwe);
    return 0;
}

// decoded bits (16 per value, MSB first) -> decimal u16 per line
inline int postprocess_stage(int argc, char **argv, const char *in_name, const char *out_name)
{
    long size;
    if (!parse_size(argc, argv, &size)) return 1;
    const std::string io_dir = std::string("io/") + size_string(size);
    uint64_t *bits = nullptr, n = 0;
    STAGE_TRY(cbs_u64_vec_load((io_dir + "/intermediate/" + in_name).c_str(), &bits, &n));
    if (n % 16 != 0) {
        fprintf(stderr, "Error: decrypted_result length is not a multiple of 16\n");
        return 1;
    }
    std::string out;
    for (uint64_t i = 0; i < n; i += 16) {
        unsigned v = 0;
        for (int b = 0; b < 16; b++) v = (v << 1) | (unsigned)(bits[i + b] & 1);
        out += std::to_string(v) + "\n";
    }
    std::ofstream f(io_dir + "/" + out_name);
    if (!f) {
        fprintf(stderr, "Error: cannot write %s/%s\n", io_dir.c_str(), out_name);
        return 1;
    }
    f << out;
    cbs_free(bits);
    return 0;
}
