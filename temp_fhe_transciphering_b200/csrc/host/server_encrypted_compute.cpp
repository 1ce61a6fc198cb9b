// server_encrypted_compute <size> — drop-in for the reference stage 8 executable
// (submission/src/bin/server_encrypted_compute.rs:99-359): reads
// io/<s>/ciphertext_aes_download/result.bin (16 LWE bits per u16, MSB first), computes the
// encrypted maximum and writes io/<s>/ciphertexts_download/result.bin (16 LWE).
// The reference hard-fails unless there are exactly 8 values (:207-210); any count >= 1 works here.
// The harness selects the mini-workload with --mini_workload but does not pass it on
// (harness/run_submission.py:97,129-131); a second argument "1" or CBS_MINI_WORKLOAD=1 selects the inner
// product mod 2^16 of the first half of the values with the second half instead of the maximum.
#include "stage_common.h"

int main(int argc, char **argv)
{
    long size;
    if (!parse_size(argc, argv, &size)) return 1;
    const std::string io_dir = std::string("io/") + size_string(size);

    uint64_t *data = nullptr, count = 0, words = 0;
    STAGE_TRY(cbs_lwe_list_load((io_dir + "/ciphertext_aes_download/result.bin").c_str(), &data, &count, &words));
    if (words != CBS_LWE_BIG_WORDS || count == 0 || count % 16 != 0) {
        fprintf(stderr, "Error: lwe_ciphertext_list length is not a multiple of 16\n");
        return 1;
    }
    cbs_keyset *ks = nullptr;
    STAGE_TRY(cbs_keyset_load_dir(io_dir.c_str(), 0, &ks));
    cbs_ctx *ctx = nullptr;
    STAGE_TRY(cbs_ctx_create(ks, 0, &ctx));
    std::vector<uint64_t> out((size_t)16 * CBS_LWE_BIG_WORDS);
    const char *mw = argc >= 3 ? argv[2] : getenv("CBS_MINI_WORKLOAD");
    if (mw && atoi(mw) == 1)
        STAGE_TRY(cbs_inner_product_u16(ctx, data, (int)(count / 16), out.data()));
    else
        STAGE_TRY(cbs_max_u16(ctx, data, (int)(count / 16), out.data()));
    STAGE_TRY(cbs_lwe_list_save((io_dir + "/ciphertexts_download/result.bin").c_str(), out.data(), 16, CBS_LWE_BIG_WORDS));
    cbs_ctx_destroy(ctx);
    cbs_keyset_free(ks);
    cbs_free(data);
    return 0;
}
