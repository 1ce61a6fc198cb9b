// cbs_api.cu — device context, workspaces and the compute entry points of include/cbs_b200.h.
//
// Orchestration mirrors the reference's server mains, batched over all ciphertexts of all blocks:
//   stage 7  src/bin/server_encrypted_aes_decryption.rs:28-191  (aes_to_lwe_trasnciphering)
//   stage 8  src/bin/server_encrypted_compute.rs:99-359
// Everything is enqueued on one CUDA stream; there is no host synchronisation inside a
// transciphering call except the final download in the host-buffer variants.
#include "cbs_b200.h"
#include "cbs_kernels.cuh"
#include "fft_tables.h"
#include "host/host_common.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "host/ip_plan.h"

using namespace cbs;
using cbs_host::set_error;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

}  // namespace

struct cbs_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    uint64_t launches = 0;
    DeviceKeys K{};
    std::vector<void *> key_allocs;
    uint64_t *d_k10_9 = nullptr, *d_k8_1 = nullptr, *d_k0 = nullptr;
    bool have_trans_key = false;
    uint64_t *d_kf_first = nullptr, *d_kf_mid = nullptr, *d_kf_last = nullptr;
    bool have_fwd_key = false;
    // every mask word of the round 8..1 / round 0 LUTs is zero (always true for AllRdKeys written by the
    // reference, src/data_struct.rs:145-151,258-263); checked at upload, enables the first-CMux shortcut
    int inv_luts_trivial = 0, fwd_luts_trivial = 0;  // host-side promise (unused since the device flags below)
    int *d_masks_nonzero = nullptr;                  // [0] round 8..0 LUTs, [1] forward (CTR) LUTs: 0 = every mask word is zero
    int jobs24_nblocks[5] = {-1, -1, -1, -1, -1};
    std::map<std::string, DevBuf> ws;
    // cached LUT job tables, keyed by block count
    int chunk_blocks = 64;
    // Lanes: a chunk is processed as up to kMaxLanes block-aligned parts on separate side streams so that
    // the second (partial) wave of one part's blind rotation overlaps the trace / scheme switch / LUT
    // kernels of another part.  lane = -1: work is issued on `stream` with un-prefixed workspaces.
    static constexpr int kMaxLanes = 4;
    int lanes = 2;
    int lane = -1;
    cudaStream_t side[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};
    int jobs_nblocks[kMaxLanes + 1] = {-1, -1, -1, -1, -1};
    // end-to-end calls upload the round 8..0 LUTs (26 MB) on a copy stream while the first rounds and the first blind
    // rotation already run; the first LUT launch of every lane waits for ev_keys
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_keys = nullptr, ev_copy_after = nullptr;
    bool keys_in_flight = false;
};

namespace {

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                               \
            return CBS_ERR_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error(std::string(what) + ": " + cudaGetErrorString(e));
        return CBS_ERR_CUDA;
    }
    return CBS_OK;
}

// stream the current lane issues on
inline cudaStream_t S(const cbs_ctx *ctx) { return ctx->lane >= 0 ? ctx->side[ctx->lane] : ctx->stream; }

int ws_get(cbs_ctx *ctx, const char *name, size_t bytes, void **out)
{
    DevBuf &b = ctx->ws[ctx->lane >= 0 ? std::string("lane") + std::to_string(ctx->lane) + ":" + name : std::string(name)];
    if (b.bytes < bytes) {
        // stream-ordered growth (no device-wide synchronisation inside a compute call): a workspace is only ever used on
        // the stream of the lane that owns it, so freeing and re-allocating on that stream orders correctly with its work;
        // the device's default pool keeps freed blocks (release threshold set in cbs_ctx_create)
        cudaStream_t st = S(ctx);
        if (b.p) {
            CUDA_TRY(cudaFreeAsync(b.p, st));
            b.p = nullptr;
            b.bytes = 0;
        }
        CUDA_TRY(cudaMallocAsync(&b.p, bytes, st));
        b.bytes = bytes;
    }
    *out = b.p;
    return CBS_OK;
}

template <typename T>
int ws_typed(cbs_ctx *ctx, const char *name, size_t count, T **out)
{
    void *p = nullptr;
    int rc = ws_get(ctx, name, count * sizeof(T), &p);
    *out = static_cast<T *>(p);
    return rc;
}

int key_alloc(cbs_ctx *ctx, size_t bytes, void **out)
{
    CUDA_TRY(cudaMalloc(out, bytes));
    ctx->key_allocs.push_back(*out);
    return CBS_OK;
}

struct Activate {
    int prev = -1;
    bool ok = true;
    explicit Activate(const cbs_ctx *ctx)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (cudaSetDevice(ctx->device) != cudaSuccess) ok = false;
    }
    ~Activate()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

#define ENTER(ctx)                                           \
    if (!(ctx)) {                                            \
        set_error("null context");                           \
        return CBS_ERR_ARG;                                  \
    }                                                        \
    Activate _act(ctx);                                      \
    if (!_act.ok) {                                          \
        set_error("cudaSetDevice failed");                   \
        return CBS_ERR_CUDA;                                 \
    }

int upload(cbs_ctx *ctx, void *d, const void *h, size_t bytes)
{
    CUDA_TRY(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return CBS_OK;
}
int download(cbs_ctx *ctx, void *h, const void *d, size_t bytes)
{
    CUDA_TRY(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return CBS_OK;
}

#define TRY(expr)                  \
    do {                           \
        int _rc = (expr);          \
        if (_rc != CBS_OK) return _rc; \
    } while (0)

// ---- device pipelines (all pointers device, async on S(ctx)) ----

int dev_keyswitch(cbs_ctx *ctx, const uint64_t *d_in, uint64_t *d_out, int count)
{
    launch_lwe_keyswitch(ctx->K, d_in, d_out, count, S(ctx));
    ctx->launches++;
    return check_launch("k_lwe_keyswitch");
}

int dev_blind_rotate(cbs_ctx *ctx, const uint64_t *d_in, uint64_t *d_acc, int count)
{
    launch_blind_rotate(ctx->K, d_in, d_acc, count, S(ctx));
    ctx->launches++;
    return check_launch("k_blind_rotate");
}

// lwe_msb_bit_to_glev_by_trace_with_preprocessing for `count` small LWE -> glev[count][7][3072]
int dev_msb_to_glev(cbs_ctx *ctx, const uint64_t *d_in, uint64_t *d_glev, int count)
{
    uint64_t *d_acc;
    TRY(ws_typed(ctx, "acc", (size_t)count * kGlweWords, &d_acc));
    TRY(dev_blind_rotate(ctx, d_in, d_acc, count));
    launch_trace(ctx->K, d_acc, d_glev, count * kCbsLevel, 1, S(ctx));
    ctx->launches++;
    return check_launch("k_trace");
}

// full circuit bootstrap: small LWE -> GGSW (Fourier and/or standard)
int dev_circuit_bootstrap(cbs_ctx *ctx, const uint64_t *d_in, uint64_t *d_ggsw_std, double *d_ggsw_f, int count)
{
    uint64_t *d_glev;
    TRY(ws_typed(ctx, "glev", (size_t)count * kGlevWords, &d_glev));
    TRY(dev_msb_to_glev(ctx, d_in, d_glev, count));
    launch_scheme_switch(ctx->K, d_glev, d_ggsw_std, d_ggsw_f, count, S(ctx));
    ctx->launches++;
    return check_launch("k_scheme_switch");
}

int ensure_job_tables(cbs_ctx *ctx, int nb, int **lut32, int **out32, int **lut8, int **out8)
{
    const int j32 = nb * 16 * 8, j8 = nb * 16 * 2;
    TRY(ws_typed(ctx, "job_lut32", (size_t)j32, lut32));
    TRY(ws_typed(ctx, "job_out32", (size_t)j32, out32));
    TRY(ws_typed(ctx, "job_lut8", (size_t)j8, lut8));
    TRY(ws_typed(ctx, "job_out8", (size_t)j8, out8));
    if (ctx->jobs_nblocks[ctx->lane + 1] == nb) return CBS_OK;
    std::vector<int> l32(j32), o32(j32), l8(j8), o8(j8);
    for (int blk = 0; blk < nb; blk++)
        for (int byte = 0; byte < 16; byte++) {
            for (int m = 0; m < 4; m++)
                for (int a = 0; a < 2; a++) {
                    const int job = (blk * 16 + byte) * 8 + m * 2 + a;
                    l32[job] = (m * 16 + byte) * 2 + a;
                    o32[job] = (m * nb + blk) * 128 + byte * 8 + 4 * a;
                }
            for (int a = 0; a < 2; a++) {
                const int job = (blk * 16 + byte) * 2 + a;
                l8[job] = byte * 2 + a;
                o8[job] = blk * 128 + byte * 8 + 4 * a;
            }
        }
    CUDA_TRY(cudaMemcpyAsync(*lut32, l32.data(), sizeof(int) * j32, cudaMemcpyHostToDevice, S(ctx)));
    CUDA_TRY(cudaMemcpyAsync(*out32, o32.data(), sizeof(int) * j32, cudaMemcpyHostToDevice, S(ctx)));
    CUDA_TRY(cudaMemcpyAsync(*lut8, l8.data(), sizeof(int) * j8, cudaMemcpyHostToDevice, S(ctx)));
    CUDA_TRY(cudaMemcpyAsync(*out8, o8.data(), sizeof(int) * j8, cudaMemcpyHostToDevice, S(ctx)));
    CUDA_TRY(cudaStreamSynchronize(S(ctx)));  // host vectors go out of scope
    ctx->jobs_nblocks[ctx->lane + 1] = nb;
    return CBS_OK;
}

// aes_to_lwe_trasnciphering for one chunk of nb blocks
int dev_transcipher_part(cbs_ctx *ctx, const uint8_t *d_ct, int nb, uint64_t *d_out)
{
    const int B = nb * 128;
    uint64_t *d_t4, *d_st, *d_ks;
    double *d_ggsw_f;
    int *lut32, *out32, *lut8, *out8;
    TRY(ws_typed(ctx, "t4", (size_t)4 * B * kLweBig, &d_t4));
    TRY(ws_typed(ctx, "st", (size_t)B * kLweBig, &d_st));
    TRY(ws_typed(ctx, "ks", (size_t)B * kLweSmall, &d_ks));
    TRY(ws_typed(ctx, "ggsw_f", (size_t)B * kGgswWords, &d_ggsw_f));
    TRY(ensure_job_tables(ctx, nb, &lut32, &out32, &lut8, &out8));
    // rounds 10 + 9 (server_encrypted_aes_decryption.rs:89-128)
    launch_known_rotate(d_ct, ctx->d_k10_9, d_t4, nb, 4, 1, S(ctx));
    launch_inv_linear(d_t4, d_st, nb, S(ctx));
    ctx->launches += 2;
    TRY(check_launch("first rounds"));
    // rounds 8..1 (:130-163)
    for (int round = 8; round >= 1; round--) {
        TRY(dev_keyswitch(ctx, d_st, d_ks, B));
        TRY(dev_circuit_bootstrap(ctx, d_ks, nullptr, d_ggsw_f, B));
        const uint64_t *luts = ctx->d_k8_1 + (size_t)(round - 1) * 4 * 16 * 2 * kGlweWords;
        if (ctx->keys_in_flight && round == 8) CUDA_TRY(cudaStreamWaitEvent(S(ctx), ctx->ev_keys, 0));
        launch_lut8(ctx->K, d_ggsw_f, luts, lut32, out32, d_t4, nb * 16 * 8, 8, ctx->inv_luts_trivial, ctx->d_masks_nonzero, S(ctx));
        launch_inv_linear(d_t4, d_st, nb, S(ctx));
        ctx->launches += 2;
        TRY(check_launch("round"));
    }
    // last round (:166-180) + per-byte bit reversal (:182-189)
    TRY(dev_keyswitch(ctx, d_st, d_ks, B));
    TRY(dev_circuit_bootstrap(ctx, d_ks, nullptr, d_ggsw_f, B));
    launch_lut8(ctx->K, d_ggsw_f, ctx->d_k0, lut8, out8, d_st, nb * 16 * 2, 2, ctx->inv_luts_trivial, ctx->d_masks_nonzero, S(ctx));
    launch_reverse_bits(d_st, d_out, nb, S(ctx));
    ctx->launches += 2;
    return check_launch("last round");
}

// CTR mode: forward AES of the public counter blocks (aes_he.rs:285-474), one chunk of nb blocks
int dev_ctr_part(cbs_ctx *ctx, const uint8_t *d_ctr, const uint8_t *d_ct, int nb, uint64_t *d_out)
{
    const int B = nb * 128;
    uint64_t *d_t3, *d_st, *d_ks;
    double *d_ggsw_f;
    int *lut24, *out24, *lut8, *out8, *dummy_a, *dummy_b;
    TRY(ws_typed(ctx, "t4", (size_t)4 * B * kLweBig, &d_t3));
    TRY(ws_typed(ctx, "st", (size_t)B * kLweBig, &d_st));
    TRY(ws_typed(ctx, "ks", (size_t)B * kLweSmall, &d_ks));
    TRY(ws_typed(ctx, "ggsw_f", (size_t)B * kGgswWords, &d_ggsw_f));
    TRY(ensure_job_tables(ctx, nb, &dummy_a, &dummy_b, &lut8, &out8));
    const int j24 = nb * 16 * 6;
    TRY(ws_typed(ctx, "job_lut24", (size_t)j24, &lut24));
    TRY(ws_typed(ctx, "job_out24", (size_t)j24, &out24));
    if (ctx->jobs24_nblocks[ctx->lane + 1] != nb) {
        std::vector<int> l(j24), o(j24);
        for (int blk = 0; blk < nb; blk++)
            for (int byte = 0; byte < 16; byte++)
                for (int m = 0; m < 3; m++)
                    for (int a = 0; a < 2; a++) {
                        const int job = (blk * 16 + byte) * 6 + m * 2 + a;
                        l[job] = (m * 16 + byte) * 2 + a;
                        o[job] = (m * nb + blk) * 128 + byte * 8 + 4 * a;
                    }
        CUDA_TRY(cudaMemcpyAsync(lut24, l.data(), sizeof(int) * j24, cudaMemcpyHostToDevice, S(ctx)));
        CUDA_TRY(cudaMemcpyAsync(out24, o.data(), sizeof(int) * j24, cudaMemcpyHostToDevice, S(ctx)));
        CUDA_TRY(cudaStreamSynchronize(S(ctx)));
        ctx->jobs24_nblocks[ctx->lane + 1] = nb;
    }
    // round 1: the counter block is public -> keyed LUTs are "rotated" by sample extraction
    launch_known_rotate(d_ctr, ctx->d_kf_first, d_t3, nb, 3, 0, S(ctx));
    launch_fwd_linear(d_t3, d_st, nb, S(ctx));
    ctx->launches += 2;
    TRY(check_launch("ctr round 1"));
    for (int round = 2; round <= 9; round++) {
        TRY(dev_keyswitch(ctx, d_st, d_ks, B));
        TRY(dev_circuit_bootstrap(ctx, d_ks, nullptr, d_ggsw_f, B));
        const uint64_t *luts = ctx->d_kf_mid + (size_t)(round - 2) * 3 * 16 * 2 * kGlweWords;
        launch_lut8(ctx->K, d_ggsw_f, luts, lut24, out24, d_t3, j24, 6, ctx->fwd_luts_trivial, ctx->d_masks_nonzero + 1, S(ctx));
        launch_fwd_linear(d_t3, d_st, nb, S(ctx));
        ctx->launches += 2;
        TRY(check_launch("ctr round"));
    }
    TRY(dev_keyswitch(ctx, d_st, d_ks, B));
    TRY(dev_circuit_bootstrap(ctx, d_ks, nullptr, d_ggsw_f, B));
    launch_lut8(ctx->K, d_ggsw_f, ctx->d_kf_last, lut8, out8, d_st, nb * 16 * 2, 2, ctx->fwd_luts_trivial, ctx->d_masks_nonzero + 1, S(ctx));
    launch_ctr_finish(d_st, d_ct, d_out, nb, S(ctx));
    ctx->launches += 2;
    return check_launch("ctr last round");
}

// Run `part(lane_index, first_block, block_count)` over a chunk split into block-aligned lanes on the side
// streams, forked from and joined back into ctx->stream with events (no host synchronisation).
template <typename Fn>
int run_in_lanes(cbs_ctx *ctx, int nb, Fn part)
{
    const int P = std::max(1, std::min(std::min(ctx->lanes, (int)cbs_ctx::kMaxLanes), nb));
    if (P == 1) return part(0, nb);
    CUDA_TRY(cudaEventRecord(ctx->ev_fork, ctx->stream));
    int rc = CBS_OK;
    for (int l = 0; l < P && rc == CBS_OK; l++) {
        const int b0 = (int)((long)nb * l / P), b1 = (int)((long)nb * (l + 1) / P);
        CUDA_TRY(cudaStreamWaitEvent(ctx->side[l], ctx->ev_fork, 0));
        ctx->lane = l;
        rc = part(b0, b1 - b0);
        ctx->lane = -1;
        if (rc != CBS_OK) break;
        CUDA_TRY(cudaEventRecord(ctx->ev_join[l], ctx->side[l]));
        CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[l], 0));
    }
    return rc;
}

int dev_transcipher_chunk(cbs_ctx *ctx, const uint8_t *d_ct, int nb, uint64_t *d_out)
{
    return run_in_lanes(ctx, nb, [&](int b0, int n) {
        return dev_transcipher_part(ctx, d_ct + (size_t)b0 * 16, n, d_out + (size_t)b0 * 128 * kLweBig);
    });
}

int dev_ctr_chunk(cbs_ctx *ctx, const uint8_t *d_ctr, const uint8_t *d_ct, int nb, uint64_t *d_out)
{
    return run_in_lanes(ctx, nb, [&](int b0, int n) {
        return dev_ctr_part(ctx, d_ctr + (size_t)b0 * 16, d_ct + (size_t)b0 * 16, n, d_out + (size_t)b0 * 128 * kLweBig);
    });
}

int dev_ctr(cbs_ctx *ctx, const uint8_t *d_ctr, const uint8_t *d_ct, int nblocks, uint64_t *d_out)
{
    if (!ctx->have_fwd_key) {
        set_error("forward transciphering key not uploaded (cbs_fwd_trans_key_upload)");
        return CBS_ERR_ARG;
    }
    for (int b0 = 0; b0 < nblocks; b0 += ctx->chunk_blocks) {
        const int nb = std::min(ctx->chunk_blocks, nblocks - b0);
        TRY(dev_ctr_chunk(ctx, d_ctr + (size_t)b0 * 16, d_ct + (size_t)b0 * 16, nb, d_out + (size_t)b0 * 128 * kLweBig));
    }
    return CBS_OK;
}

int dev_transcipher(cbs_ctx *ctx, const uint8_t *d_ct, int nblocks, uint64_t *d_out)
{
    if (!ctx->have_trans_key) {
        set_error("transciphering key not uploaded (cbs_trans_key_upload)");
        return CBS_ERR_ARG;
    }
    for (int b0 = 0; b0 < nblocks; b0 += ctx->chunk_blocks) {
        const int nb = std::min(ctx->chunk_blocks, nblocks - b0);
        TRY(dev_transcipher_chunk(ctx, d_ct + (size_t)b0 * 16, nb, d_out + (size_t)b0 * 128 * kLweBig));
    }
    return CBS_OK;
}

}  // namespace

static int run_lut_plan(cbs_ctx *ctx, const cbs_host::IpPlan &plan, const uint64_t *in, uint64_t *out);

extern "C" {

int cbs_device_count(int *count)
{
    if (!count) return CBS_ERR_ARG;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        set_error(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
        return CBS_ERR_CUDA;
    }
    *count = n;
    return CBS_OK;
}

// workspaces come from the stream-ordered allocator (ws_get): keep what it frees cached in the device's default pool
static void keep_pool_memory(int device)
{
    cudaMemPool_t pool = nullptr;
    uint64_t keep = UINT64_MAX;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    cudaGetLastError();
}

int cbs_device_init(int device)
{
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaFree(nullptr));
    keep_pool_memory(device);
    return CBS_OK;
}

int cbs_ctx_create(const cbs_keyset *ks, int device, cbs_ctx **out)
{
    if (!ks || !out) {
        set_error("cbs_ctx_create: null argument");
        return CBS_ERR_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: libcbs_b200 has no CPU fallback");
        return CBS_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        set_error("cbs_ctx_create: device index out of range");
        return CBS_ERR_ARG;
    }
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error(std::string("device ") + prop.name + " is not sm_100: kernels are built for sm_100a only");
        return CBS_ERR_CUDA;
    }
    auto *ctx = new cbs_ctx;
    ctx->device = device;
    if (const char *e = getenv("CBS_CHUNK_BLOCKS")) ctx->chunk_blocks = std::max(1, atoi(e));
    if (const char *e = getenv("CBS_LANES")) ctx->lanes = std::max(1, std::min((int)cbs_ctx::kMaxLanes, atoi(e)));
    Activate act(ctx);
    keep_pool_memory(device);
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        set_error(std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
        delete ctx;
        return CBS_ERR_CUDA;
    }
    auto fail = [&](int rc) {
        cbs_ctx_destroy(ctx);
        return rc;
    };
    for (int l = 0; l < cbs_ctx::kMaxLanes; l++) {
        if (cudaStreamCreateWithFlags(&ctx->side[l], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_join[l], cudaEventDisableTiming) != cudaSuccess) {
            set_error("cannot create lane streams");
            return fail(CBS_ERR_CUDA);
        }
    }
    if (cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess) {
        set_error("cannot create lane events");
        return fail(CBS_ERR_CUDA);
    }
    // twiddle tables
    std::vector<double> tw = make_twiddle_table();
    std::vector<double> tw128(256 + 128);
    for (int j = 0; j < 128; j++) {
        long double a = 3.14159265358979323846264338327950288L * j / 256.0L;
        tw128[2 * j] = (double)cosl(a);
        tw128[2 * j + 1] = (double)sinl(a);
    }
    for (int j = 0; j < 64; j++) {
        long double a = -2.0L * 3.14159265358979323846264338327950288L * j / 128.0L;
        tw128[256 + 2 * j] = (double)cosl(a);
        tw128[256 + 2 * j + 1] = (double)sinl(a);
    }
    void *d_tw, *d_tw128, *d_bsk_f, *d_auto_f, *d_ss_f, *d_ksk_f, *d_tmp;
    int rc;
    if ((rc = key_alloc(ctx, tw.size() * 8, &d_tw)) || (rc = key_alloc(ctx, tw128.size() * 8, &d_tw128)) ||
        (rc = key_alloc(ctx, (size_t)CBS_BSK_WORDS * 8, &d_bsk_f)) ||
        (rc = key_alloc(ctx, (size_t)CBS_AUTO_WORDS * 2 * 8, &d_auto_f)) ||
        (rc = key_alloc(ctx, (size_t)CBS_SS_WORDS * 8, &d_ss_f)) ||
        (rc = key_alloc(ctx, (size_t)CBS_KSK_WORDS * 8, &d_ksk_f)))
        return fail(rc);
    if ((rc = upload(ctx, d_tw, tw.data(), tw.size() * 8)) || (rc = upload(ctx, d_tw128, tw128.data(), tw128.size() * 8)))
        return fail(rc);
    {
        // device flags "some LUT mask word is non-zero" (k_masks_nonzero after every LUT upload); 1 = not trivial until known
        void *d_flags = nullptr;
        const int ones[2] = {1, 1};
        if ((rc = key_alloc(ctx, sizeof(ones), &d_flags))) return fail(rc);
        ctx->d_masks_nonzero = (int *)d_flags;
        if (cudaMemcpy(d_flags, ones, sizeof(ones), cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("cudaMemcpy(flags) failed");
            return fail(CBS_ERR_CUDA);
        }
    }
    ctx->K.tw = (const double *)d_tw;
    ctx->K.tw128 = (const double *)d_tw128;
    // staging buffer for the standard-domain keys (largest = bsk)
    if (cudaMalloc(&d_tmp, (size_t)CBS_BSK_WORDS * 8) != cudaSuccess) {
        set_error("cudaMalloc(staging) failed");
        return fail(CBS_ERR_NOMEM);
    }
    auto conv = [&](const uint64_t *h, size_t words, double *dst, int mode) -> int {
        TRY(upload(ctx, d_tmp, h, words * 8));
        launch_std_to_fourier((const uint64_t *)d_tmp, dst, (int)(words / 1024), mode, 41, ctx->K.tw, ctx->stream);
        ctx->launches++;
        return check_launch("k_std_to_fourier");
    };
    // bsk: convert_standard_lwe_bootstrap_key_to_fourier (server_encrypted_aes_decryption.rs:656-663)
    if ((rc = conv(ks->bsk.data(), CBS_BSK_WORDS, (double *)d_bsk_f, 0))) {
        cudaFree(d_tmp);
        return fail(rc);
    }
    // ss key (:665-687)
    if ((rc = conv(ks->ss.data(), CBS_SS_WORDS, (double *)d_ss_f, 0))) {
        cudaFree(d_tmp);
        return fail(rc);
    }
    // automorphism keys: Split(41) limbs, layout [10][2 in][2 split][3 level][3 col]
    // (convert_standard_glwe_keyswitch_key_to_fourier, fourier_glwe_keyswitch.rs:154-211)
    if ((rc = upload(ctx, d_tmp, ks->autok.data(), (size_t)CBS_AUTO_WORDS * 8))) {
        cudaFree(d_tmp);
        return fail(rc);
    }
    for (int idx = 0; idx < 10; idx++)
        for (int in = 0; in < 2; in++)
            for (int sp = 0; sp < 2; sp++) {
                const uint64_t *src = (const uint64_t *)d_tmp + ((size_t)idx * 2 + in) * 9 * 1024;
                double *dst = (double *)d_auto_f + (((size_t)idx * 2 + in) * 2 + sp) * 9 * kFourierPolyDoubles;
                launch_std_to_fourier(src, dst, 9, sp ? 2 : 1, 41, ctx->K.tw, ctx->stream);
                ctx->launches++;
            }
    if ((rc = check_launch("auto key conversion"))) {
        cudaFree(d_tmp);
        return fail(rc);
    }
    cudaStreamSynchronize(ctx->stream);
    // ksk over N' = 256 (:646-654)
    if ((rc = upload(ctx, d_tmp, ks->ksk.data(), (size_t)CBS_KSK_WORDS * 8))) {
        cudaFree(d_tmp);
        return fail(rc);
    }
    launch_ksk_to_fourier((const uint64_t *)d_tmp, (double *)d_ksk_f, 8 * 3 * 4, ctx->K.tw128, ctx->stream);
    ctx->launches++;
    rc = check_launch("k_ksk_to_fourier");
    cudaError_t se = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_tmp);
    if (rc) return fail(rc);
    if (se != cudaSuccess) {
        set_error(std::string("key conversion failed: ") + cudaGetErrorString(se));
        return fail(CBS_ERR_CUDA);
    }
    ctx->K.bsk_f = (const double *)d_bsk_f;
    ctx->K.auto_f = (const double *)d_auto_f;
    ctx->K.ss_f = (const double *)d_ss_f;
    ctx->K.ksk_f = (const double *)d_ksk_f;
    *out = ctx;
    return CBS_OK;
}

void cbs_ctx_destroy(cbs_ctx *ctx)
{
    if (!ctx) return;
    Activate act(ctx);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaDeviceSynchronize();
    for (void *p : ctx->key_allocs) cudaFree(p);
    for (auto &kv : ctx->ws)
        if (kv.second.p) cudaFree(kv.second.p);
    if (ctx->d_k10_9) cudaFree(ctx->d_k10_9);
    if (ctx->d_k8_1) cudaFree(ctx->d_k8_1);
    if (ctx->d_k0) cudaFree(ctx->d_k0);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->ev_keys) cudaEventDestroy(ctx->ev_keys);
    if (ctx->ev_copy_after) cudaEventDestroy(ctx->ev_copy_after);
    if (ctx->d_kf_first) cudaFree(ctx->d_kf_first);
    if (ctx->d_kf_mid) cudaFree(ctx->d_kf_mid);
    if (ctx->d_kf_last) cudaFree(ctx->d_kf_last);
    for (int l = 0; l < cbs_ctx::kMaxLanes; l++) {
        if (ctx->side[l]) cudaStreamDestroy(ctx->side[l]);
        if (ctx->ev_join[l]) cudaEventDestroy(ctx->ev_join[l]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int cbs_ctx_device(const cbs_ctx *ctx) { return ctx ? ctx->device : -1; }
uint64_t cbs_ctx_launch_count(const cbs_ctx *ctx) { return ctx ? ctx->launches : 0; }

int cbs_ctx_set_stream(cbs_ctx *ctx, void *cuda_stream)
{
    ENTER(ctx);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
        ctx->own_stream = false;
    } else {
        CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return CBS_OK;
}

int cbs_ctx_synchronize(cbs_ctx *ctx)
{
    ENTER(ctx);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return CBS_OK;
}

int cbs_measure_fp64_tflops(cbs_ctx *ctx, double *tflops)
{
    ENTER(ctx);
    if (!tflops) return CBS_ERR_ARG;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, ctx->device));
    const int blocks = prop.multiProcessorCount * 8, iters = 20000;
    double *d;
    TRY(ws_typed(ctx, "fp64_probe", (size_t)blocks * 256, &d));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CUDA_TRY(cudaEventRecord(e0, ctx->stream));
        launch_fp64_peak(d, blocks, iters, ctx->stream);
        CUDA_TRY(cudaEventRecord(e1, ctx->stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ctx->launches += 4;
    *tflops = 2.0 * 16.0 * (double)iters * 256.0 * blocks / (best * 1e-3) * 1e-12;
    return check_launch("k_fp64_peak");
}

int cbs_dev_alloc(cbs_ctx *ctx, size_t bytes, void **dptr)
{
    ENTER(ctx);
    CUDA_TRY(cudaMalloc(dptr, bytes));
    return CBS_OK;
}
int cbs_dev_free(cbs_ctx *ctx, void *dptr)
{
    ENTER(ctx);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaFree(dptr));
    return CBS_OK;
}
int cbs_dev_upload(cbs_ctx *ctx, void *dptr, const void *host, size_t bytes)
{
    ENTER(ctx);
    TRY(upload(ctx, dptr, host, bytes));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return CBS_OK;
}
int cbs_dev_download(cbs_ctx *ctx, void *host, const void *dptr, size_t bytes)
{
    ENTER(ctx);
    return download(ctx, host, dptr, bytes);
}

// ---- host-buffer stage entry points ----

int cbs_lwe_keyswitch(cbs_ctx *ctx, const uint64_t *in_big, uint64_t *out_small, int count)
{
    ENTER(ctx);
    if (count < 0 || (count && (!in_big || !out_small))) return set_error("cbs_lwe_keyswitch: bad argument"), CBS_ERR_ARG;
    if (!count) return CBS_OK;
    uint64_t *d_in, *d_out;
    TRY(ws_typed(ctx, "io_in", (size_t)count * kLweBig, &d_in));
    TRY(ws_typed(ctx, "ks", (size_t)count * kLweSmall, &d_out));
    TRY(upload(ctx, d_in, in_big, (size_t)count * kLweBig * 8));
    TRY(dev_keyswitch(ctx, d_in, d_out, count));
    return download(ctx, out_small, d_out, (size_t)count * kLweSmall * 8);
}

int cbs_blind_rotate(cbs_ctx *ctx, const uint64_t *in_small, uint64_t *acc_out, int count)
{
    ENTER(ctx);
    if (count < 0 || (count && (!in_small || !acc_out))) return set_error("cbs_blind_rotate: bad argument"), CBS_ERR_ARG;
    if (!count) return CBS_OK;
    uint64_t *d_in, *d_acc;
    TRY(ws_typed(ctx, "ks", (size_t)count * kLweSmall, &d_in));
    TRY(ws_typed(ctx, "acc", (size_t)count * kGlweWords, &d_acc));
    TRY(upload(ctx, d_in, in_small, (size_t)count * kLweSmall * 8));
    TRY(dev_blind_rotate(ctx, d_in, d_acc, count));
    return download(ctx, acc_out, d_acc, (size_t)count * kGlweWords * 8);
}

int cbs_blind_rotate_dev(cbs_ctx *ctx, const uint64_t *d_in_small, uint64_t *d_acc_out, int count)
{
    ENTER(ctx);
    if (count <= 0 || !d_in_small || !d_acc_out) return set_error("cbs_blind_rotate_dev: bad argument"), CBS_ERR_ARG;
    return dev_blind_rotate(ctx, d_in_small, d_acc_out, count);
}

int cbs_glev_from_acc(cbs_ctx *ctx, const uint64_t *acc, uint64_t *glev_out, int count)
{
    ENTER(ctx);
    if (count < 0 || (count && (!acc || !glev_out))) return set_error("cbs_glev_from_acc: bad argument"), CBS_ERR_ARG;
    if (!count) return CBS_OK;
    uint64_t *d_acc, *d_glev;
    TRY(ws_typed(ctx, "acc", (size_t)count * kGlweWords, &d_acc));
    TRY(ws_typed(ctx, "glev", (size_t)count * kGlevWords, &d_glev));
    TRY(upload(ctx, d_acc, acc, (size_t)count * kGlweWords * 8));
    launch_glev_from_acc(d_acc, d_glev, count, ctx->stream);
    ctx->launches++;
    TRY(check_launch("k_glev_from_acc"));
    return download(ctx, glev_out, d_glev, (size_t)count * kGlevWords * 8);
}

int cbs_trace(cbs_ctx *ctx, uint64_t *glwe_inout, int count)
{
    ENTER(ctx);
    if (count < 0 || (count && !glwe_inout)) return set_error("cbs_trace: bad argument"), CBS_ERR_ARG;
    if (!count) return CBS_OK;
    uint64_t *d_in, *d_out;
    TRY(ws_typed(ctx, "io_in", (size_t)count * kGlweWords, &d_in));
    TRY(ws_typed(ctx, "glev", (size_t)count * kGlweWords, &d_out));
    TRY(upload(ctx, d_in, glwe_inout, (size_t)count * kGlweWords * 8));
    launch_trace(ctx->K, d_in, d_out, count, 0, ctx->stream);
    ctx->launches++;
    TRY(check_launch("k_trace"));
    return download(ctx, glwe_inout, d_out, (size_t)count * kGlweWords * 8);
}

int cbs_lwe_msb_bit_to_glev(cbs_ctx *ctx, const uint64_t *in_small, uint64_t *glev_out, int count)
{
    ENTER(ctx);
    if (count < 0 || (count && (!in_small || !glev_out))) return set_error("cbs_lwe_msb_bit_to_glev: bad argument"), CBS_ERR_ARG;
    if (!count) return CBS_OK;
    uint64_t *d_in, *d_glev;
    TRY(ws_typed(ctx, "ks", (size_t)count * kLweSmall, &d_in));
    TRY(ws_typed(ctx, "glev", (size_t)count * kGlevWords, &d_glev));
    TRY(upload(ctx, d_in, in_small, (size_t)count * kLweSmall * 8));
    TRY(dev_msb_to_glev(ctx, d_in, d_glev, count));
    return download(ctx, glev_out, d_glev, (size_t)count * kGlevWords * 8);
}

int cbs_scheme_switch(cbs_ctx *ctx, const uint64_t *glev, uint64_t *ggsw_out, int count)
{
    ENTER(ctx);
    if (count < 0 || (count && (!glev || !ggsw_out))) return set_error("cbs_scheme_switch: bad argument"), CBS_ERR_ARG;
    if (!count) return CBS_OK;
    uint64_t *d_glev, *d_ggsw;
    TRY(ws_typed(ctx, "glev", (size_t)count * kGlevWords, &d_glev));
    TRY(ws_typed(ctx, "ggsw_std", (size_t)count * kGgswWords, &d_ggsw));
    TRY(upload(ctx, d_glev, glev, (size_t)count * kGlevWords * 8));
    launch_scheme_switch(ctx->K, d_glev, d_ggsw, nullptr, count, ctx->stream);
    ctx->launches++;
    TRY(check_launch("k_scheme_switch"));
    return download(ctx, ggsw_out, d_ggsw, (size_t)count * kGgswWords * 8);
}

int cbs_circuit_bootstrap(cbs_ctx *ctx, const uint64_t *in_small, uint64_t *ggsw_out, int count)
{
    ENTER(ctx);
    if (count < 0 || (count && (!in_small || !ggsw_out))) return set_error("cbs_circuit_bootstrap: bad argument"), CBS_ERR_ARG;
    if (!count) return CBS_OK;
    uint64_t *d_in, *d_ggsw;
    double *d_ggsw_f;
    TRY(ws_typed(ctx, "ks", (size_t)count * kLweSmall, &d_in));
    TRY(ws_typed(ctx, "ggsw_std", (size_t)count * kGgswWords, &d_ggsw));
    TRY(ws_typed(ctx, "ggsw_f", (size_t)count * kGgswWords, &d_ggsw_f));
    TRY(upload(ctx, d_in, in_small, (size_t)count * kLweSmall * 8));
    TRY(dev_circuit_bootstrap(ctx, d_in, d_ggsw, d_ggsw_f, count));
    return download(ctx, ggsw_out, d_ggsw, (size_t)count * kGgswWords * 8);
}

int cbs_circuit_bootstrap_dev(cbs_ctx *ctx, const uint64_t *d_in_small, int count)
{
    ENTER(ctx);
    if (count <= 0 || !d_in_small) return set_error("cbs_circuit_bootstrap_dev: bad argument"), CBS_ERR_ARG;
    double *d_ggsw_f;
    TRY(ws_typed(ctx, "ggsw_f", (size_t)count * kGgswWords, &d_ggsw_f));
    return dev_circuit_bootstrap(ctx, d_in_small, nullptr, d_ggsw_f, count);
}

int cbs_lut8_eval(cbs_ctx *ctx, const uint64_t *ggsw_bits, int nbytes, const uint64_t *luts, int nluts, uint64_t *out)
{
    ENTER(ctx);
    if (nbytes < 0 || nluts <= 0 || (nbytes && (!ggsw_bits || !luts || !out))) return set_error("cbs_lut8_eval: bad argument"), CBS_ERR_ARG;
    if (!nbytes) return CBS_OK;
    const int nbits = nbytes * 8, apb = nluts * 2, njobs = nbytes * apb;
    uint64_t *d_ggsw, *d_luts, *d_out;
    double *d_ggsw_f;
    int *d_li, *d_oi;
    TRY(ws_typed(ctx, "ggsw_std", (size_t)nbits * kGgswWords, &d_ggsw));
    TRY(ws_typed(ctx, "ggsw_f", (size_t)nbits * kGgswWords, &d_ggsw_f));
    TRY(ws_typed(ctx, "io_luts", (size_t)njobs * kGlweWords, &d_luts));
    TRY(ws_typed(ctx, "io_out", (size_t)njobs * 4 * kLweBig, &d_out));
    TRY(ws_typed(ctx, "io_li", (size_t)njobs, &d_li));
    TRY(ws_typed(ctx, "io_oi", (size_t)njobs, &d_oi));
    std::vector<int> li(njobs), oi(njobs);
    for (int j = 0; j < njobs; j++) {
        li[j] = j;      // luts[nbytes][nluts][2] in job order
        oi[j] = j * 4;  // out[nbytes][nluts][8]
    }
    TRY(upload(ctx, d_ggsw, ggsw_bits, (size_t)nbits * kGgswWords * 8));
    TRY(upload(ctx, d_luts, luts, (size_t)njobs * kGlweWords * 8));
    TRY(upload(ctx, d_li, li.data(), sizeof(int) * njobs));
    TRY(upload(ctx, d_oi, oi.data(), sizeof(int) * njobs));
    launch_ggsw_to_fourier(ctx->K, d_ggsw, d_ggsw_f, nbits, ctx->stream);
    launch_lut8(ctx->K, d_ggsw_f, d_luts, d_li, d_oi, d_out, njobs, apb, 0, nullptr, ctx->stream);
    ctx->launches += 2;
    TRY(check_launch("k_lut8"));
    return download(ctx, out, d_out, (size_t)njobs * 4 * kLweBig * 8);
}

int cbs_aes_first_rounds(cbs_ctx *ctx, const uint8_t *ct, int nblocks, const uint64_t *k10_9, uint64_t *state_out)
{
    ENTER(ctx);
    if (nblocks < 0 || (nblocks && (!ct || !k10_9 || !state_out))) return set_error("cbs_aes_first_rounds: bad argument"), CBS_ERR_ARG;
    if (!nblocks) return CBS_OK;
    uint8_t *d_ct;
    uint64_t *d_k, *d_t4, *d_st;
    TRY(ws_typed(ctx, "io_ct", (size_t)nblocks * 16, &d_ct));
    TRY(ws_typed(ctx, "io_luts", (size_t)CBS_K10_9_WORDS, &d_k));
    TRY(ws_typed(ctx, "t4", (size_t)4 * nblocks * 128 * kLweBig, &d_t4));
    TRY(ws_typed(ctx, "st", (size_t)nblocks * 128 * kLweBig, &d_st));
    TRY(upload(ctx, d_ct, ct, (size_t)nblocks * 16));
    TRY(upload(ctx, d_k, k10_9, (size_t)CBS_K10_9_WORDS * 8));
    launch_known_rotate(d_ct, d_k, d_t4, nblocks, 4, 1, ctx->stream);
    launch_inv_linear(d_t4, d_st, nblocks, ctx->stream);
    ctx->launches += 2;
    TRY(check_launch("first rounds"));
    return download(ctx, state_out, d_st, (size_t)nblocks * 128 * kLweBig * 8);
}

int cbs_aes_inv_linear(cbs_ctx *ctx, const uint64_t *t4, int nblocks, uint64_t *state_out)
{
    ENTER(ctx);
    if (nblocks < 0 || (nblocks && (!t4 || !state_out))) return set_error("cbs_aes_inv_linear: bad argument"), CBS_ERR_ARG;
    if (!nblocks) return CBS_OK;
    uint64_t *d_t4, *d_st;
    TRY(ws_typed(ctx, "t4", (size_t)4 * nblocks * 128 * kLweBig, &d_t4));
    TRY(ws_typed(ctx, "st", (size_t)nblocks * 128 * kLweBig, &d_st));
    TRY(upload(ctx, d_t4, t4, (size_t)4 * nblocks * 128 * kLweBig * 8));
    launch_inv_linear(d_t4, d_st, nblocks, ctx->stream);
    ctx->launches++;
    TRY(check_launch("k_inv_linear"));
    return download(ctx, state_out, d_st, (size_t)nblocks * 128 * kLweBig * 8);
}

int cbs_trans_key_upload(cbs_ctx *ctx, const uint64_t *k10_9, const uint64_t *k8_1, const uint64_t *k0)
{
    ENTER(ctx);
    if (!k10_9 || !k8_1 || !k0) return set_error("cbs_trans_key_upload: null argument"), CBS_ERR_ARG;
    // each buffer on its own: a failed allocation leaves its pointer null and the next call retries it
    if (!ctx->d_k10_9) CUDA_TRY(cudaMalloc(&ctx->d_k10_9, (size_t)CBS_K10_9_WORDS * 8));
    if (!ctx->d_k8_1) CUDA_TRY(cudaMalloc(&ctx->d_k8_1, (size_t)CBS_K8_1_WORDS * 8));
    if (!ctx->d_k0) CUDA_TRY(cudaMalloc(&ctx->d_k0, (size_t)CBS_K0_WORDS * 8));
    TRY(upload(ctx, ctx->d_k10_9, k10_9, (size_t)CBS_K10_9_WORDS * 8));
    TRY(upload(ctx, ctx->d_k8_1, k8_1, (size_t)CBS_K8_1_WORDS * 8));
    TRY(upload(ctx, ctx->d_k0, k0, (size_t)CBS_K0_WORDS * 8));
    // are the LUT accumulators trivial (zero masks)?  Answered on the device, right behind the copies
    CUDA_TRY(cudaMemsetAsync(ctx->d_masks_nonzero, 0, sizeof(int), ctx->stream));
    launch_masks_nonzero(ctx->d_k8_1, 8 * 4 * 16 * 2, ctx->d_masks_nonzero, ctx->stream);
    launch_masks_nonzero(ctx->d_k0, 16 * 2, ctx->d_masks_nonzero, ctx->stream);
    ctx->inv_luts_trivial = 0;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // the caller owns the host buffers again
    ctx->have_trans_key = true;
    return CBS_OK;
}

int cbs_aes128_transcipher_dev(cbs_ctx *ctx, const uint8_t *d_ct, int nblocks, uint64_t *d_out)
{
    ENTER(ctx);
    if (nblocks <= 0 || !d_ct || !d_out) return set_error("cbs_aes128_transcipher_dev: bad argument"), CBS_ERR_ARG;
    return dev_transcipher(ctx, d_ct, nblocks, d_out);
}

int cbs_aes128_transcipher(cbs_ctx *ctx, const uint8_t *ct, int nblocks, const uint64_t *k10_9, const uint64_t *k8_1,
                           const uint64_t *k0, uint64_t *out)
{
    ENTER(ctx);
    if (nblocks < 0 || (nblocks && (!ct || !out)) || !k10_9 || !k8_1 || !k0) return set_error("cbs_aes128_transcipher: bad argument"), CBS_ERR_ARG;
    if (!nblocks) return CBS_OK;
    // keys: rounds 10 + 9 need k10_9 (3 MB, on the compute stream); k8_1 and k0 (26 MB) are first read by the LUT ladders of
    // round 8, one keyswitch + blind rotation (>= 2.4 ms) into the call, so they travel on the copy stream meanwhile
    if (!ctx->d_k10_9) CUDA_TRY(cudaMalloc(&ctx->d_k10_9, (size_t)CBS_K10_9_WORDS * 8));
    if (!ctx->d_k8_1) CUDA_TRY(cudaMalloc(&ctx->d_k8_1, (size_t)CBS_K8_1_WORDS * 8));
    if (!ctx->d_k0) CUDA_TRY(cudaMalloc(&ctx->d_k0, (size_t)CBS_K0_WORDS * 8));
    if (!ctx->copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_keys, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_copy_after, cudaEventDisableTiming));
    }
    // the copy must not overtake earlier work on the compute stream that may still read the old keys
    CUDA_TRY(cudaEventRecord(ctx->ev_copy_after, ctx->stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy_after, 0));
    CUDA_TRY(cudaMemcpyAsync(ctx->d_k8_1, k8_1, (size_t)CBS_K8_1_WORDS * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
    CUDA_TRY(cudaMemcpyAsync(ctx->d_k0, k0, (size_t)CBS_K0_WORDS * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
    CUDA_TRY(cudaMemsetAsync(ctx->d_masks_nonzero, 0, sizeof(int), ctx->copy_stream));
    launch_masks_nonzero(ctx->d_k8_1, 8 * 4 * 16 * 2, ctx->d_masks_nonzero, ctx->copy_stream);
    launch_masks_nonzero(ctx->d_k0, 16 * 2, ctx->d_masks_nonzero, ctx->copy_stream);
    CUDA_TRY(cudaEventRecord(ctx->ev_keys, ctx->copy_stream));
    TRY(upload(ctx, ctx->d_k10_9, k10_9, (size_t)CBS_K10_9_WORDS * 8));
    uint8_t *d_ct;
    uint64_t *d_out;
    TRY(ws_typed(ctx, "io_ct", (size_t)nblocks * 16, &d_ct));
    TRY(ws_typed(ctx, "io_result", (size_t)nblocks * 128 * kLweBig, &d_out));
    TRY(upload(ctx, d_ct, ct, (size_t)nblocks * 16));
    ctx->inv_luts_trivial = 0;
    ctx->have_trans_key = true;
    ctx->keys_in_flight = true;
    int rc = dev_transcipher(ctx, d_ct, nblocks, d_out);
    // whatever happened, the compute stream is ordered after the key copy before the call returns (host buffers, later calls)
    cudaStreamWaitEvent(ctx->stream, ctx->ev_keys, 0);
    ctx->keys_in_flight = false;
    if (rc != CBS_OK) {
        cudaStreamSynchronize(ctx->stream);
        return rc;
    }
    return download(ctx, out, d_out, (size_t)nblocks * 128 * kLweBig * 8);
}

int cbs_fwd_trans_key_upload(cbs_ctx *ctx, const uint64_t *kf_first, const uint64_t *kf_mid, const uint64_t *kf_last)
{
    ENTER(ctx);
    if (!kf_first || !kf_mid || !kf_last) return set_error("cbs_fwd_trans_key_upload: null argument"), CBS_ERR_ARG;
    if (!ctx->d_kf_first) CUDA_TRY(cudaMalloc(&ctx->d_kf_first, (size_t)CBS_KF_FIRST_WORDS * 8));
    if (!ctx->d_kf_mid) CUDA_TRY(cudaMalloc(&ctx->d_kf_mid, (size_t)CBS_KF_MID_WORDS * 8));
    if (!ctx->d_kf_last) CUDA_TRY(cudaMalloc(&ctx->d_kf_last, (size_t)CBS_KF_LAST_WORDS * 8));
    TRY(upload(ctx, ctx->d_kf_first, kf_first, (size_t)CBS_KF_FIRST_WORDS * 8));
    TRY(upload(ctx, ctx->d_kf_mid, kf_mid, (size_t)CBS_KF_MID_WORDS * 8));
    TRY(upload(ctx, ctx->d_kf_last, kf_last, (size_t)CBS_KF_LAST_WORDS * 8));
    CUDA_TRY(cudaMemsetAsync(ctx->d_masks_nonzero + 1, 0, sizeof(int), ctx->stream));
    launch_masks_nonzero(ctx->d_kf_mid, 8 * 3 * 16 * 2, ctx->d_masks_nonzero + 1, ctx->stream);
    launch_masks_nonzero(ctx->d_kf_last, 16 * 2, ctx->d_masks_nonzero + 1, ctx->stream);
    ctx->fwd_luts_trivial = 0;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->have_fwd_key = true;
    return CBS_OK;
}

int cbs_aes128_ctr_transcipher_dev(cbs_ctx *ctx, const uint8_t *d_ctr, const uint8_t *d_ct, int nblocks, uint64_t *d_out)
{
    ENTER(ctx);
    if (nblocks <= 0 || !d_ctr || !d_ct || !d_out) return set_error("cbs_aes128_ctr_transcipher_dev: bad argument"), CBS_ERR_ARG;
    return dev_ctr(ctx, d_ctr, d_ct, nblocks, d_out);
}

int cbs_aes128_ctr_transcipher(cbs_ctx *ctx, const uint8_t *ct, int nblocks, const uint8_t iv[16], const uint64_t *kf_first,
                               const uint64_t *kf_mid, const uint64_t *kf_last, uint64_t *out)
{
    ENTER(ctx);
    if (nblocks < 0 || (nblocks && (!ct || !out)) || !iv || !kf_first || !kf_mid || !kf_last)
        return set_error("cbs_aes128_ctr_transcipher: bad argument"), CBS_ERR_ARG;
    if (!nblocks) return CBS_OK;
    TRY(cbs_fwd_trans_key_upload(ctx, kf_first, kf_mid, kf_last));
    // counter blocks: 128-bit big-endian IV + block index (pyaes.Counter, harness/aes_keygen_and_encrypt.py:52)
    std::vector<uint8_t> ctr((size_t)nblocks * 16);
    uint8_t cur[16];
    memcpy(cur, iv, 16);
    for (int b = 0; b < nblocks; b++) {
        memcpy(ctr.data() + (size_t)b * 16, cur, 16);
        for (int i = 15; i >= 0; i--)
            if (++cur[i] != 0) break;
    }
    uint8_t *d_ctr, *d_ct;
    uint64_t *d_out;
    TRY(ws_typed(ctx, "io_ctr", (size_t)nblocks * 16, &d_ctr));
    TRY(ws_typed(ctx, "io_ct", (size_t)nblocks * 16, &d_ct));
    TRY(ws_typed(ctx, "io_result", (size_t)nblocks * 128 * kLweBig, &d_out));
    TRY(upload(ctx, d_ctr, ctr.data(), (size_t)nblocks * 16));
    TRY(upload(ctx, d_ct, ct, (size_t)nblocks * 16));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // `ctr` is a local vector
    TRY(dev_ctr(ctx, d_ctr, d_ct, nblocks, d_out));
    return download(ctx, out, d_out, (size_t)nblocks * 128 * kLweBig * 8);
}

int cbs_max_u16(cbs_ctx *ctx, const uint64_t *in, int nvals, uint64_t *out)
{
    ENTER(ctx);
    if (nvals <= 0 || !in || !out) return set_error("cbs_max_u16: bad argument"), CBS_ERR_ARG;
    if (nvals == 1) {
        memcpy(out, in, (size_t)16 * kLweBig * 8);
        return CBS_OK;
    }
    // up to the reference's own 8 values: its max_of_two CMux ladder; above (where the reference cannot run and the
    // ladder's data-dependent noise becomes a real failure probability): the LUT circuit.  CBS_MAX_VARIANT=ladder|lut overrides.
    {
        const char *e = getenv("CBS_MAX_VARIANT");
        const bool lut = e ? (strcmp(e, "lut") == 0) : nvals > 8;
        if (lut) return run_lut_plan(ctx, cbs_host::max_make_plan(nvals), in, out);
    }
    // balanced tree: level buffers A (current) and B (next)
    uint64_t *d_lwe[2], *d_op[2], *d_ks, *d_glev;
    double *d_ggsw[2];
    int *d_a, *d_b;
    const size_t vals_lwe = (size_t)16 * kLweBig, vals_ggsw = (size_t)16 * kGgswWords, vals_op = (size_t)16 * kGlweWords;
    TRY(ws_typed(ctx, "max_lwe0", (size_t)nvals * vals_lwe, &d_lwe[0]));
    TRY(ws_typed(ctx, "max_lwe1", (size_t)((nvals + 1) / 2) * vals_lwe, &d_lwe[1]));
    TRY(ws_typed(ctx, "max_ggsw0", (size_t)nvals * vals_ggsw, &d_ggsw[0]));
    TRY(ws_typed(ctx, "max_ggsw1", (size_t)((nvals + 1) / 2) * vals_ggsw, &d_ggsw[1]));
    TRY(ws_typed(ctx, "max_op0", (size_t)nvals * vals_op, &d_op[0]));
    TRY(ws_typed(ctx, "max_op1", (size_t)((nvals + 1) / 2) * vals_op, &d_op[1]));
    TRY(ws_typed(ctx, "ks", (size_t)nvals * 16 * kLweSmall, &d_ks));
    TRY(ws_typed(ctx, "glev", (size_t)nvals * 16 * kGlevWords, &d_glev));  // the circuit bootstrap's own GLEV workspace
    TRY(ws_typed(ctx, "max_a", (size_t)nvals, &d_a));
    TRY(ws_typed(ctx, "max_b", (size_t)nvals, &d_b));
    std::vector<int> ha(nvals / 2), hb(nvals / 2);
    for (int i = 0; i < nvals / 2; i++) {
        ha[i] = 2 * i;
        hb[i] = 2 * i + 1;
    }
    TRY(upload(ctx, d_a, ha.data(), sizeof(int) * ha.size()));
    TRY(upload(ctx, d_b, hb.data(), sizeof(int) * hb.size()));
    TRY(upload(ctx, d_lwe[0], in, (size_t)nvals * vals_lwe * 8));
    int cnt = nvals, cur = 0, fresh = nvals;  // `fresh` leading values of the current level still need a CBS
    while (cnt > 1) {
        // keyswitch + circuit bootstrap the values produced by the previous level (server_encrypted_compute.rs:213-263,314-342);
        // the level-1 GLEV of the same bootstrap, doubled, is the refreshed data operand of the ladder
        TRY(dev_keyswitch(ctx, d_lwe[cur], d_ks, fresh * 16));
        TRY(dev_circuit_bootstrap(ctx, d_ks, nullptr, d_ggsw[cur], fresh * 16));
        launch_glev_to_operand(d_glev, d_op[cur], fresh * 16, ctx->stream);
        ctx->launches++;
        TRY(check_launch("k_glev_to_operand"));
        const int npairs = cnt / 2;
        launch_max_ladder(ctx->K, d_ggsw[cur], d_op[cur], d_a, d_b, d_lwe[cur ^ 1], npairs, ctx->stream);
        ctx->launches++;
        TRY(check_launch("k_max_ladder"));
        int next = npairs;
        if (cnt & 1) {  // odd one out is carried with its GGSW and operand (no new bootstrap needed)
            CUDA_TRY(cudaMemcpyAsync(d_lwe[cur ^ 1] + (size_t)npairs * vals_lwe, d_lwe[cur] + (size_t)(cnt - 1) * vals_lwe,
                                     vals_lwe * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(d_ggsw[cur ^ 1] + (size_t)npairs * vals_ggsw, d_ggsw[cur] + (size_t)(cnt - 1) * vals_ggsw,
                                     vals_ggsw * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(d_op[cur ^ 1] + (size_t)npairs * vals_op, d_op[cur] + (size_t)(cnt - 1) * vals_op,
                                     vals_op * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            next++;
        }
        fresh = npairs;
        cnt = next;
        cur ^= 1;
    }
    return download(ctx, out, d_lwe[cur], vals_lwe * 8);
}

}  // extern "C"

// Executes a host-planned LUT circuit (host/ip_plan.h) on `nin` input bit ciphertexts: every layer is
// [LWE additions] -> [gather -> LWE keyswitch -> circuit bootstrap] -> [fresh operands from the bootstrap's GLEV] ->
// [gathered LUT ladders] -> [LWE additions]; layers whose bootstraps exceed the workspace cap run in sub-batches.
static int run_lut_plan(cbs_ctx *ctx, const cbs_host::IpPlan &plan, const uint64_t *in, uint64_t *out)
{
    using namespace cbs_host;
    constexpr int kCap = 8192;  // circuit bootstraps per sub-batch (bounds the Fourier GGSW workspace to 4.2 GB)
    struct Batch {
        int idx_off, m, job_off, njobs, ref_off, nref;
    };
    struct LayerExec {
        int pre_off, npre, post_off, npost;
        std::vector<Batch> batches;
    };
    std::vector<LayerExec> layers;
    std::vector<int> h_idx, h_sel, h_lut, h_out, h_add, h_ref;
    for (const IpLayer &L : plan.layers) {
        LayerExec E;
        E.pre_off = (int)h_add.size() / 3;
        E.npre = (int)L.pre.size();
        for (const IpAdd &a : L.pre) h_add.insert(h_add.end(), {a.dst, a.a, a.b});
        E.post_off = (int)h_add.size() / 3;
        E.npost = (int)L.post.size();
        for (const IpAdd &a : L.post) h_add.insert(h_add.end(), {a.dst, a.a, a.b});
        std::vector<int> remap(L.cbs.size(), -1), ref_dst(L.cbs.size(), -1), touched;
        for (const IpRefresh &r : L.refresh) ref_dst[(size_t)r.pos] = r.dst;
        Batch b{(int)h_idx.size(), 0, (int)h_lut.size(), 0, (int)h_ref.size() / 2, 0};
        auto flush = [&]() {
            if (b.njobs) E.batches.push_back(b);
            for (int pos : touched) remap[(size_t)pos] = -1;
            touched.clear();
            b = Batch{(int)h_idx.size(), 0, (int)h_lut.size(), 0, (int)h_ref.size() / 2, 0};
        };
        for (const IpJob &j : L.jobs) {
            int fresh = 0;
            for (int i = 0; i < 8; i++)
                if (j.sel[i] >= 0 && remap[(size_t)j.sel[i]] < 0) fresh++;
            if (b.m + fresh > kCap) flush();
            for (int i = 0; i < 8; i++) {
                int v = -1;
                if (j.sel[i] >= 0) {
                    int &r = remap[(size_t)j.sel[i]];
                    if (r < 0) {
                        r = b.m++;
                        touched.push_back(j.sel[i]);
                        h_idx.push_back(L.cbs[(size_t)j.sel[i]]);
                        if (ref_dst[(size_t)j.sel[i]] >= 0) {
                            h_ref.insert(h_ref.end(), {r, ref_dst[(size_t)j.sel[i]]});
                            b.nref++;
                        }
                    }
                    v = r;
                }
                h_sel.push_back(v);
            }
            h_lut.push_back(j.lut * 2 + j.acc);
            h_out.push_back(j.out);
            b.njobs++;
        }
        flush();
        layers.push_back(std::move(E));
    }
    int max_m = 1;
    for (const LayerExec &E : layers)
        for (const Batch &b : E.batches) max_m = b.m > max_m ? b.m : max_m;
    for (int i = 0; i < 16; i++) h_idx.push_back(plan.result[15 - i]);  // final gather, MSB first

    // trivial LUT accumulators: coefficient i carries bit (4*acc + i/256) of table[i % 256] at 2^63
    // (same layout as generate_vec_keyed_lut_accumulator, cbs_lib/src/aes_he.rs:835-875)
    std::vector<uint64_t> h_luts((size_t)kIpNumLuts * 2 * kGlweWords, 0);
    for (int l = 0; l < kIpNumLuts; l++)
        for (int a = 0; a < 2; a++) {
            uint64_t *body = h_luts.data() + (size_t)(l * 2 + a) * kGlweWords + 2048;
            for (int i = 0; i < 1024; i++) body[i] = (uint64_t)((ip_lut_value(l, (unsigned)(i % 256)) >> (4 * a + i / 256)) & 1u) << 63;
        }

    uint64_t *d_pool, *d_rows, *d_ks, *d_luts, *d_glev;
    double *d_ggsw_f;
    int *d_idx, *d_sel, *d_lut, *d_out, *d_add, *d_ref;
    TRY(ws_typed(ctx, "ip_pool", (size_t)plan.pool_size * kLweBig, &d_pool));
    TRY(ws_typed(ctx, "ip_rows", (size_t)max_m * kLweBig, &d_rows));
    TRY(ws_typed(ctx, "ks", (size_t)max_m * kLweSmall, &d_ks));
    TRY(ws_typed(ctx, "ggsw_f", (size_t)max_m * kGgswWords, &d_ggsw_f));
    TRY(ws_typed(ctx, "glev", (size_t)max_m * kGlevWords, &d_glev));  // the circuit bootstrap's own GLEV workspace
    TRY(ws_typed(ctx, "ip_luts", h_luts.size(), &d_luts));
    TRY(ws_typed(ctx, "ip_idx", h_idx.size(), &d_idx));
    TRY(ws_typed(ctx, "ip_sel", h_sel.size() + 1, &d_sel));
    TRY(ws_typed(ctx, "ip_lut", h_lut.size() + 1, &d_lut));
    TRY(ws_typed(ctx, "ip_out", h_out.size() + 1, &d_out));
    TRY(ws_typed(ctx, "ip_add", h_add.size() + 1, &d_add));
    TRY(ws_typed(ctx, "ip_ref", h_ref.size() + 1, &d_ref));
    TRY(upload(ctx, d_pool, in, (size_t)plan.nin_bits * kLweBig * 8));
    TRY(upload(ctx, d_luts, h_luts.data(), h_luts.size() * 8));
    TRY(upload(ctx, d_idx, h_idx.data(), h_idx.size() * sizeof(int)));
    if (!h_lut.empty()) {
        TRY(upload(ctx, d_sel, h_sel.data(), h_sel.size() * sizeof(int)));
        TRY(upload(ctx, d_lut, h_lut.data(), h_lut.size() * sizeof(int)));
        TRY(upload(ctx, d_out, h_out.data(), h_out.size() * sizeof(int)));
    }
    if (!h_add.empty()) TRY(upload(ctx, d_add, h_add.data(), h_add.size() * sizeof(int)));
    if (!h_ref.empty()) TRY(upload(ctx, d_ref, h_ref.data(), h_ref.size() * sizeof(int)));
    for (const LayerExec &E : layers) {
        if (E.npre) {
            launch_lwe_add_rows(d_pool, d_add + (size_t)E.pre_off * 3, E.npre, ctx->stream);
            ctx->launches++;
            TRY(check_launch("k_lwe_add_rows"));
        }
        for (const Batch &b : E.batches) {
            launch_gather_lwe(d_pool, d_idx + b.idx_off, d_rows, b.m, ctx->stream);
            ctx->launches++;
            TRY(check_launch("k_gather_lwe"));
            TRY(dev_keyswitch(ctx, d_rows, d_ks, b.m));
            TRY(dev_circuit_bootstrap(ctx, d_ks, nullptr, d_ggsw_f, b.m));
            if (b.nref) {
                launch_glev_to_lwe(d_glev, d_ref + (size_t)b.ref_off * 2, d_pool, b.nref, ctx->stream);
                ctx->launches++;
                TRY(check_launch("k_glev_to_lwe"));
            }
            launch_lut8_gather(ctx->K, d_ggsw_f, d_sel + (size_t)b.job_off * 8, d_luts, d_lut + b.job_off, d_out + b.job_off, d_pool,
                               b.njobs, ctx->stream);
            ctx->launches++;
            TRY(check_launch("k_lut8_gather"));
        }
        if (E.npost) {
            launch_lwe_add_rows(d_pool, d_add + (size_t)E.post_off * 3, E.npost, ctx->stream);
            ctx->launches++;
            TRY(check_launch("k_lwe_add_rows"));
        }
    }
    launch_gather_lwe(d_pool, d_idx + (h_idx.size() - 16), d_rows, 16, ctx->stream);
    ctx->launches++;
    TRY(check_launch("k_gather_lwe"));
    // download() synchronises the stream, so the host tables above outlive every asynchronous copy
    return download(ctx, out, d_rows, (size_t)16 * kLweBig * 8);
}

extern "C" {

// Mini-workload #2 of the harness (harness/cleartext_impl.py:65-70; no reference implementation exists, SURVEY.md
// 8(f)2): sum_i (x_i * y_i mod 2^16) mod 2^16 with x = first half, y = second half of the values.  The circuit
// (nibble products, column compression by population counts, nibble adders) is planned on the host by
// host/ip_plan.h and executed by run_lut_plan.
int cbs_inner_product_u16(cbs_ctx *ctx, const uint64_t *in, int nvals, uint64_t *out)
{
    ENTER(ctx);
    if (nvals <= 0 || (nvals & 1) || !in || !out) return set_error("cbs_inner_product_u16: bad argument (need an even number of values)"), CBS_ERR_ARG;
    return run_lut_plan(ctx, cbs_host::ip_make_plan(nvals), in, out);
}

// Sum of nvals 16-bit values mod 2^16 (the column-compression stage of the inner product alone): combines per-GPU
// partial inner products when the pairs are sharded across GPUs.
int cbs_sum_u16(cbs_ctx *ctx, const uint64_t *in, int nvals, uint64_t *out)
{
    ENTER(ctx);
    if (nvals <= 0 || !in || !out) return set_error("cbs_sum_u16: bad argument"), CBS_ERR_ARG;
    return run_lut_plan(ctx, cbs_host::sum_make_plan(nvals), in, out);
}

// Maximum as a LUT circuit (host/ip_plan.h max_make_plan): noise independent of the data and of the tree depth.
int cbs_max_u16_lut(cbs_ctx *ctx, const uint64_t *in, int nvals, uint64_t *out)
{
    ENTER(ctx);
    if (nvals <= 0 || !in || !out) return set_error("cbs_max_u16_lut: bad argument"), CBS_ERR_ARG;
    return run_lut_plan(ctx, cbs_host::max_make_plan(nvals), in, out);
}

}  // extern "C"
