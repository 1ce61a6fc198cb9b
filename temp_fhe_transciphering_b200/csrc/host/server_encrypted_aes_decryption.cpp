// server_encrypted_aes_decryption <size> — drop-in for the reference stage 7 executable
// (submission/src/bin/server_encrypted_aes_decryption.rs:599-707): same name, argv, input files
// (datasets/<s>/db.hex, io/<s>/public_keys/{bsk,ksk,auto_keys,ss_key}.bin,
// io/<s>/ciphertexts_upload/trans_key.bin) and output (io/<s>/ciphertext_aes_download/result.bin).
// The reference transciphers only the first 16-byte block of db.hex (:613-617); this binary
// transciphers every block, sharded contiguously over the GPUs it decides to use (stage_common.h plan_visible_gpus:
// one GPU up to ~120 blocks because driver start-up costs more than the transciphering; CBS_GPUS / CUDA_VISIBLE_DEVICES
// override).
// Mode follows the harness (harness/aes_keygen_and_encrypt.py:45-55): size 0 = ECB block decryption
// with the reference's AllRdKeys; sizes 1/2 = CTR with datasets/<s>/aes_iv.hex and the forward-direction
// transciphering key written by our client_encode_encrypt (the reference has no CTR path).
#include "stage_common.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <thread>

int main(int argc, char **argv)
{
    long size;
    if (!parse_size(argc, argv, &size)) return 1;
    StageClock clk("server_encrypted_aes_decryption");
    const std::string io_dir = std::string("io/") + size_string(size);
    const std::string data_dir = std::string("datasets/") + size_string(size);

    std::vector<uint8_t> ct;
    if (!read_hex_file(data_dir + "/db.hex", ct) || ct.size() < 16) {
        fprintf(stderr, "Error: cannot read %s/db.hex\n", data_dir.c_str());
        return 1;
    }
    const int nblocks = (int)(ct.size() / 16);
    plan_visible_gpus(nblocks);  // before the first CUDA call

    cbs_keyset *ks = nullptr;
    STAGE_TRY(cbs_keyset_load_dir(io_dir.c_str(), 0, &ks));
    clk.mark("read_public_keys");
    const bool ctr_mode = size >= 1;
    std::vector<uint8_t> iv;
    if (ctr_mode && (!read_hex_file(data_dir + "/aes_iv.hex", iv) || iv.size() != 16)) {
        fprintf(stderr, "Error: cannot read %s/aes_iv.hex\n", data_dir.c_str());
        return 1;
    }
    std::vector<uint64_t> k10_9(ctr_mode ? CBS_KF_FIRST_WORDS : CBS_K10_9_WORDS),
        k8_1(ctr_mode ? CBS_KF_MID_WORDS : CBS_K8_1_WORDS), k0(CBS_K0_WORDS);
    const std::string tk_path = io_dir + "/ciphertexts_upload/trans_key.bin";
    if (ctr_mode) STAGE_TRY(cbs_fwd_trans_key_load(tk_path.c_str(), k10_9.data(), k8_1.data(), k0.data()));
    else STAGE_TRY(cbs_trans_key_load(tk_path.c_str(), k10_9.data(), k8_1.data(), k0.data()));

    clk.mark("read_trans_key");
    int ngpu = 0;
    if (cbs_device_count(&ngpu) != CBS_OK || ngpu == 0) {
        fprintf(stderr, "Error: no CUDA device (this executable has no CPU fallback)\n");
        return 1;
    }
    if (const char *e = getenv("CBS_GPUS")) ngpu = std::max(1, std::min(ngpu, atoi(e)));
    ngpu = std::min(ngpu, nblocks);

    std::vector<uint64_t> result((size_t)nblocks * 128 * CBS_LWE_BIG_WORDS);
    if (getenv("CBS_STAGE_TIMING")) {  // separate the driver / primary-context start-up from our own set-up
        std::vector<std::thread> init;  // in parallel, like the worker threads would do it implicitly
        for (int g = 0; g < ngpu; g++) init.emplace_back([g]() { cbs_device_init(g); });
        for (auto &th : init) th.join();
        clk.mark("cuda_init");
    }
    std::vector<double> t_ctx(ngpu, 0.0), t_run(ngpu, 0.0);
    std::vector<int> rc(ngpu, 0);
    std::vector<std::string> err(ngpu);
    std::vector<std::thread> workers;
    for (int g = 0; g < ngpu; g++) {
        workers.emplace_back([&, g]() {
            const int b0 = (int)((long)nblocks * g / ngpu), b1 = (int)((long)nblocks * (g + 1) / ngpu);
            cbs_ctx *ctx = nullptr;
            const auto w0 = std::chrono::steady_clock::now();
            rc[g] = cbs_ctx_create(ks, g, &ctx);
            const auto w1 = std::chrono::steady_clock::now();
            t_ctx[g] = std::chrono::duration<double, std::milli>(w1 - w0).count();
            if (rc[g] == CBS_OK && !ctr_mode)
                rc[g] = cbs_aes128_transcipher(ctx, ct.data() + (size_t)b0 * 16, b1 - b0, k10_9.data(), k8_1.data(), k0.data(),
                                               result.data() + (size_t)b0 * 128 * CBS_LWE_BIG_WORDS);
            if (rc[g] == CBS_OK && ctr_mode) {
                uint8_t civ[16];  // counter of this shard's first block: IV + b0 (128-bit big-endian)
                memcpy(civ, iv.data(), 16);
                unsigned carry = (unsigned)b0;
                for (int i = 15; i >= 0 && carry; i--) {
                    carry += civ[i];
                    civ[i] = (uint8_t)carry;
                    carry >>= 8;
                }
                rc[g] = cbs_aes128_ctr_transcipher(ctx, ct.data() + (size_t)b0 * 16, b1 - b0, civ, k10_9.data(), k8_1.data(),
                                                   k0.data(), result.data() + (size_t)b0 * 128 * CBS_LWE_BIG_WORDS);
            }
            if (rc[g] != CBS_OK) err[g] = cbs_last_error();
            t_run[g] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w1).count();
            // the context is NOT destroyed: this is a one-shot process, exit releases the device memory, and freeing ~6 GB of
            // workspaces was measured at up to 2 s on a shared host (profiles/r02_harness_1gpu.jsonl, medium instance)
            (void)ctx;
        });
    }
    for (auto &w : workers) w.join();
    clk.mark("workers");
    clk.add("ctx_create_max", *std::max_element(t_ctx.begin(), t_ctx.end()));
    clk.add("transcipher_max", *std::max_element(t_run.begin(), t_run.end()));
    clk.add("gpus", ngpu);
    for (int g = 0; g < ngpu; g++)
        if (rc[g] != CBS_OK) {
            fprintf(stderr, "Error: GPU %d: %s\n", g, err[g].c_str());
            return 1;
        }
    STAGE_TRY(cbs_lwe_list_save((io_dir + "/ciphertext_aes_download/result.bin").c_str(), result.data(),
                                (uint64_t)nblocks * 128, CBS_LWE_BIG_WORDS));
    clk.mark("write_result");
    cbs_keyset_free(ks);
    return 0;
}
