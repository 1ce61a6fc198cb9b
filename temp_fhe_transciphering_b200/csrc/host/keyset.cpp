// keyset.cpp — host side of libcbs_b200: the reference's io/ file formats (bincode 1.3) and the
// client-side helpers (seeded key generation, transciphering-key encoder, bit encryption).
//
// Formats: SURVEY.md 8(b); types serialised by src/bin/client_key_generation.rs:114-129,
// src/bin/client_encode_encrypt.rs:55-62, src/bin/server_encrypted_aes_decryption.rs:700-704.
// Key semantics: cbs_lib/src/keygen.rs:187-243 (bsk, ksk), cbs_lib/src/automorphism.rs:59-178
// (automorphism keys), cbs_lib/src/ggsw_conv.rs:15-77 (scheme-switching key),
// cbs_lib/src/glwe_keyswitch.rs:183-218 (GLWE keyswitch key = GLEV of -S_in),
// src/data_struct.rs:30-269 + src/aes_manager.rs:163-432 (AllRdKeys tables).
#include "cbs_b200.h"
#include "host_common.h"
#include "ip_plan.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <sys/random.h>
#include <sys/stat.h>
#include <thread>
#include <vector>

namespace cbs_host {

thread_local std::string g_last_error;
void set_error(const std::string &s) { g_last_error = s; }

// ------------------------------------------------------------------------------------------------
// bincode cursor
struct Reader {
    std::vector<uint8_t> buf;
    size_t off = 0;
    bool ok = true;
    bool load(const std::string &path)
    {
        std::ifstream f(path, std::ios::binary | std::ios::ate);
        if (!f) return false;
        std::streamsize n = f.tellg();
        f.seekg(0);
        buf.resize((size_t)n);
        return (bool)f.read(reinterpret_cast<char *>(buf.data()), n);
    }
    uint64_t u64()
    {
        if (off + 8 > buf.size()) {
            ok = false;
            return 0;
        }
        uint64_t v;
        memcpy(&v, buf.data() + off, 8);
        off += 8;
        return v;
    }
    bool vec(std::vector<uint64_t> &out, uint64_t expect)
    {
        uint64_t n = u64();
        if (!ok || n != expect || off + 8 * n > buf.size()) {
            ok = false;
            return false;
        }
        out.resize(n);
        memcpy(out.data(), buf.data() + off, 8 * n);
        off += 8 * n;
        return true;
    }
    bool vec_into(uint64_t *dst, uint64_t expect)
    {
        uint64_t n = u64();
        if (!ok || n != expect || off + 8 * n > buf.size()) {
            ok = false;
            return false;
        }
        memcpy(dst, buf.data() + off, 8 * n);
        off += 8 * n;
        return true;
    }
    void expect(uint64_t v)
    {
        if (u64() != v) ok = false;
    }
    void modulus()
    {  // CiphertextModulus<u64> native: u128 0, then scalar bits 64
        expect(0);
        expect(0);
        expect(64);
    }
    bool done() const { return ok && off == buf.size(); }
};

struct Writer {
    std::vector<uint8_t> buf;
    void u64(uint64_t v)
    {
        size_t o = buf.size();
        buf.resize(o + 8);
        memcpy(buf.data() + o, &v, 8);
    }
    void vec(const uint64_t *p, uint64_t n)
    {
        u64(n);
        size_t o = buf.size();
        buf.resize(o + 8 * n);
        memcpy(buf.data() + o, p, 8 * n);
    }
    void modulus()
    {
        u64(0);
        u64(0);
        u64(64);
    }
    bool save(const std::string &path) const
    {
        std::ofstream f(path, std::ios::binary);
        if (!f) return false;
        f.write(reinterpret_cast<const char *>(buf.data()), (std::streamsize)buf.size());
        return (bool)f;
    }
};

static void mkdirs(const std::string &path)
{
    std::string cur;
    for (size_t i = 0; i <= path.size(); i++) {
        if (i == path.size() || path[i] == '/') {
            if (!cur.empty()) mkdir(cur.c_str(), 0755);
        }
        if (i < path.size()) cur.push_back(path[i]);
    }
}

// ------------------------------------------------------------------------------------------------
// natural-order 512-point transform used only for the on-disk form of the automorphism keys
// (tfhe forward_as_torus convention, SURVEY.md 8(b) note 1)
using cd = std::complex<double>;
static void fft512_natural(std::vector<cd> &a, bool inverse)
{
    const int n = 512;
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    static std::vector<cd> w;
    if (w.empty()) {
        w.resize(n / 2);
        for (int k = 0; k < n / 2; k++) {
            long double ang = -2.0L * 3.14159265358979323846264338327950288L * k / n;
            w[k] = cd((double)cosl(ang), (double)sinl(ang));
        }
    }
    for (int len = 2; len <= n; len <<= 1) {
        int step = n / len;
        for (int i = 0; i < n; i += len)
            for (int j = 0; j < len / 2; j++) {
                cd tw = inverse ? std::conj(w[j * step]) : w[j * step];
                cd u = a[i + j], v = a[i + j + len / 2] * tw;
                a[i + j] = u + v;
                a[i + j + len / 2] = u - v;
            }
    }
    if (inverse)
        for (auto &x : a) x /= (double)n;
}
static const std::vector<cd> &twist1024()
{
    static std::vector<cd> t;
    if (t.empty()) {
        t.resize(512);
        for (int j = 0; j < 512; j++) {
            long double ang = 3.14159265358979323846264338327950288L * j / 1024.0L;
            t[j] = cd((double)cosl(ang), (double)sinl(ang));
        }
    }
    return t;
}
static void limb_to_fourier(const uint64_t *limb /*1024*/, double *out /*1024 (re,im)*/)
{
    const auto &tw = twist1024();
    std::vector<cd> z(512);
    for (int j = 0; j < 512; j++)
        z[j] = cd((double)(int64_t)limb[j] * 0x1p-64, (double)(int64_t)limb[j + 512] * 0x1p-64) * tw[j];
    fft512_natural(z, false);
    for (int j = 0; j < 512; j++) {
        out[2 * j] = z[j].real();
        out[2 * j + 1] = z[j].imag();
    }
}
// returns max distance of the recovered coefficients from integers (sanity of the FFT ordering)
static double fourier_to_limb(const double *in, uint64_t *limb, double max_value)
{
    const auto &tw = twist1024();
    std::vector<cd> z(512);
    for (int j = 0; j < 512; j++) z[j] = cd(in[2 * j], in[2 * j + 1]);
    fft512_natural(z, true);
    double worst = 0;
    for (int j = 0; j < 512; j++) {
        cd v = z[j] * std::conj(tw[j]) * 0x1p64;
        double re = v.real(), im = v.imag();
        double rr = std::nearbyint(re), ri = std::nearbyint(im);
        worst = std::max(worst, std::max(std::fabs(re - rr), std::fabs(im - ri)));
        if (rr < 0 || rr >= max_value || ri < 0 || ri >= max_value) worst = 1e30;
        limb[j] = (uint64_t)rr;
        limb[j + 512] = (uint64_t)ri;
    }
    return worst;
}

// ------------------------------------------------------------------------------------------------
// randomness.  Two engines behind one interface, both as independent streams keyed by a stream id so results do not
// depend on threading:
//   * RngSource(seed): xoshiro256++ - deterministic, NOT cryptographically secure; only for reproducible tests, benches
//     and golden vectors (explicit seed argument / CBS_SEED);
//   * RngSource::os_entropy(): ChaCha20 (RFC 8439 block function) keyed with 256 bits from getrandom(2), stream id as
//     nonce - what the stage executables use when no seed is given, like the reference's own client
//     (client_key_generation.rs:91-96 new_seeder(); harness/run_submission.py:74-77: keys differ on every run).
struct RngSource {
    bool secure = false;
    uint64_t seed = 0;
    uint32_t key[8] = {0};
    explicit RngSource(uint64_t s) : seed(s) {}
    static bool os_entropy(RngSource &out)
    {
        out.secure = true;
        size_t got = 0;
        unsigned char *p = reinterpret_cast<unsigned char *>(out.key);
        while (got < sizeof out.key) {
            ssize_t n = getrandom(p + got, sizeof out.key - got, 0);
            if (n <= 0) return false;
            got += (size_t)n;
        }
        return true;
    }
};

struct Rng {
    bool secure;
    uint64_t s[4];             // xoshiro256++ state
    uint32_t cc[16];           // ChaCha20 input block (constants, key, counter, nonce)
    uint64_t buf[8];
    int have = 0;
    static uint64_t splitmix(uint64_t &x)
    {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    Rng(const RngSource &src, uint64_t stream) : secure(src.secure)
    {
        if (secure) {
            cc[0] = 0x61707865, cc[1] = 0x3320646e, cc[2] = 0x79622d32, cc[3] = 0x6b206574;  // "expand 32-byte k"
            for (int i = 0; i < 8; i++) cc[4 + i] = src.key[i];
            cc[12] = 0;
            cc[13] = (uint32_t)stream;
            cc[14] = (uint32_t)(stream >> 32);
            cc[15] = 0;
        } else {
            uint64_t x = src.seed * 0xD1342543DE82EF95ull + stream * 0xA0761D6478BD642Full + 0x1234567ull;
            for (auto &v : s) v = splitmix(x);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    static uint32_t rotl32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }
    void chacha_block()
    {
        uint32_t x[16];
        memcpy(x, cc, sizeof x);
#define CBS_QR(a, b, c, d)                                                                       \
    x[a] += x[b], x[d] = rotl32(x[d] ^ x[a], 16), x[c] += x[d], x[b] = rotl32(x[b] ^ x[c], 12), \
        x[a] += x[b], x[d] = rotl32(x[d] ^ x[a], 8), x[c] += x[d], x[b] = rotl32(x[b] ^ x[c], 7)
        for (int r = 0; r < 10; r++) {
            CBS_QR(0, 4, 8, 12);
            CBS_QR(1, 5, 9, 13);
            CBS_QR(2, 6, 10, 14);
            CBS_QR(3, 7, 11, 15);
            CBS_QR(0, 5, 10, 15);
            CBS_QR(1, 6, 11, 12);
            CBS_QR(2, 7, 8, 13);
            CBS_QR(3, 4, 9, 14);
        }
#undef CBS_QR
        for (int i = 0; i < 8; i++) buf[i] = (uint64_t)(x[2 * i] + cc[2 * i]) | ((uint64_t)(x[2 * i + 1] + cc[2 * i + 1]) << 32);
        if (++cc[12] == 0) cc[15]++;  // 2^32 blocks of 64 bytes per (stream, cc[15]): far beyond any use here
        have = 8;
    }
    uint64_t next()
    {
        if (secure) {
            if (!have) chacha_block();
            return buf[--have];
        }
        uint64_t r = rotl(s[0] + s[3], 23) + s[0];
        uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return r;
    }
    double uniform() { return ((double)(next() >> 11) + 0.5) * 0x1p-53; }
    double gauss()
    {
        double u1 = uniform(), u2 = uniform();
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925 * u2);
    }
    uint64_t noise(double std_torus) { return (uint64_t)(int64_t)std::nearbyint(gauss() * std_torus * 0x1p64); }
};

constexpr double kStdLwe = 0.00000702047462940120;               // aes_instances.rs:78
constexpr double kStdGlwe = 0.00000000000000029403601535432533;  // aes_instances.rs:81

// body += a * s  (negacyclic, binary s)
static void add_mul_binary(uint64_t *body, const uint64_t *a, const uint64_t *s, int N)
{
    for (int i = 0; i < N; i++) {
        if (!s[i]) continue;
        for (int j = i; j < N; j++) body[j] += a[j - i];
        for (int j = 0; j < i; j++) body[j] -= a[N + j - i];
    }
}

// fresh GLWE encryption: ct = (a_0..a_{k-1}, sum a_c * s_c + e + pt)
static void glwe_encrypt(uint64_t *ct, const uint64_t *pt /*N or null*/, const uint64_t *sk, int k, int N, double std,
                         Rng &rng)
{
    for (int i = 0; i < k * N; i++) ct[i] = rng.next();
    uint64_t *body = ct + (size_t)k * N;
    for (int j = 0; j < N; j++) body[j] = rng.noise(std) + (pt ? pt[j] : 0ull);
    for (int c = 0; c < k; c++) add_mul_binary(body, ct + (size_t)c * N, sk + (size_t)c * N, N);
}

static void eval_x_k(uint64_t *out, const uint64_t *in, int N, unsigned kappa)
{
    for (int i = 0; i < N; i++) {
        unsigned long prod = (unsigned long)i * kappa;
        int j = (int)(prod % (unsigned)N);
        out[j] = ((prod / (unsigned)N) & 1) ? (0ull - in[i]) : in[i];
    }
}

static void parallel_for(int n, const std::function<void(int)> &fn)
{
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (unsigned w = 0; w < nt; w++)
        th.emplace_back([&, w]() {
            for (int i = (int)w; i < n; i += (int)nt) fn(i);
        });
    for (auto &t : th) t.join();
}

// ------------------------------------------------------------------------------------------------
// AES tables (src/aes_manager.rs)
static uint8_t gmul(uint8_t a, uint8_t b)
{
    uint8_t r = 0;
    while (b) {
        if (b & 1) r ^= a;
        a = (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1B : 0));
        b >>= 1;
    }
    return r;
}
struct AesTables {
    uint8_t sbox[256], inv[256];
    uint8_t rk[11][16];
    AesTables(const uint8_t key[16])
    {
        // S-box from the field inverse + affine map
        uint8_t invf[256] = {0};
        for (int x = 1; x < 256; x++)
            for (int y = 1; y < 256; y++)
                if (gmul((uint8_t)x, (uint8_t)y) == 1) invf[x] = (uint8_t)y;
        for (int x = 0; x < 256; x++) {
            uint8_t b = invf[x], r = 0;
            for (int i = 0; i < 8; i++) {
                int bit = ((b >> i) ^ (b >> ((i + 4) % 8)) ^ (b >> ((i + 5) % 8)) ^ (b >> ((i + 6) % 8)) ^
                           (b >> ((i + 7) % 8)) ^ (0x63 >> i)) & 1;
                r |= (uint8_t)(bit << i);
            }
            sbox[x] = r;
        }
        for (int x = 0; x < 256; x++) inv[sbox[x]] = (uint8_t)x;
        // key schedule, round key byte index = 4*col + row (aes_manager.rs:163-205)
        static const uint8_t rcon[11] = {0, 1, 2, 4, 8, 16, 32, 64, 128, 0x1B, 0x36};
        memcpy(rk[0], key, 16);
        for (int r = 1; r <= 10; r++) {
            const uint8_t *p = rk[r - 1];
            uint8_t t[4] = {sbox[p[13]], sbox[p[14]], sbox[p[15]], sbox[p[12]]};
            t[0] ^= rcon[r];
            for (int row = 0; row < 4; row++) rk[r][row] = p[row] ^ t[row];
            for (int col = 1; col < 4; col++)
                for (int row = 0; row < 4; row++) rk[r][4 * col + row] = p[4 * col + row] ^ rk[r][4 * (col - 1) + row];
        }
    }
    // get_10_9_round_lut (aes_manager.rs:338-396), before the GF multiples
    void lut_10_9(uint8_t t[16][256]) const
    {
        for (int col = 0; col < 4; col++)
            for (int row = 0; row < 4; row++) {
                int src = 4 * ((4 + col - row) % 4) + row;
                for (int x = 0; x < 256; x++) t[4 * col + row][x] = inv[x ^ rk[10][src]] ^ rk[9][4 * col + row];
            }
    }
    // forward keyed S-box of round r+1: S(x ^ rk_r[b]) (cbs_lib/src/aes_ref.rs:334-380 get_keyed_sbox*)
    void lut_fwd(int prev_round, uint8_t t[16][256]) const
    {
        for (int b = 0; b < 16; b++)
            for (int x = 0; x < 256; x++) t[b][x] = sbox[x ^ rk[prev_round][b]];
    }
    // last round: S(x ^ rk9[b]) ^ rk10[position b lands on after ShiftRows]
    void lut_fwd_last(uint8_t t[16][256]) const
    {
        for (int col = 0; col < 4; col++)
            for (int row = 0; row < 4; row++) {
                const int b = 4 * col + row, dst = 4 * ((col - row + 4) % 4) + row;
                for (int x = 0; x < 256; x++) t[b][x] = sbox[x ^ rk[9][b]] ^ rk[10][dst];
            }
    }
    // get_round_lut / get_0_round_lut (aes_manager.rs:398-432)
    void lut_round(int round, uint8_t t[16][256]) const
    {
        for (int b = 0; b < 16; b++)
            for (int x = 0; x < 256; x++) t[b][x] = inv[x] ^ rk[round][b];
    }
};

// accumulator plaintext of generate_vec_keyed_lut_accumulator (cbs_lib/src/aes_he.rs:835-875) and
// src/data_struct.rs:140-150: coefficient i carries bit (acc_idx*4 + i/256) of table[i % 256] at 2^63
static void lut_plaintext(uint64_t *pt, const uint8_t table[256], int acc_idx, uint8_t mult)
{
    for (int i = 0; i < 1024; i++) {
        int lut_idx = acc_idx * 4 + i / 256;
        uint8_t v = mult ? gmul(table[i % 256], mult) : table[i % 256];
        pt[i] = (uint64_t)((v >> lut_idx) & 1) << 63;
    }
}

}  // namespace cbs_host

using namespace cbs_host;

extern "C" {

const char *cbs_last_error(void) { return g_last_error.c_str(); }
const char *cbs_version(void) { return "cbs_b200 0.1 (sm_100a)"; }
void cbs_free(void *p) { free(p); }

void cbs_keyset_free(cbs_keyset *ks) { delete ks; }
const uint64_t *cbs_keyset_bsk(const cbs_keyset *ks) { return ks->bsk.data(); }
const uint64_t *cbs_keyset_ksk(const cbs_keyset *ks) { return ks->ksk.data(); }
const uint64_t *cbs_keyset_auto(const cbs_keyset *ks) { return ks->autok.data(); }
const uint64_t *cbs_keyset_ss(const cbs_keyset *ks) { return ks->ss.data(); }
const uint64_t *cbs_keyset_lwe_sk_small(const cbs_keyset *ks) { return ks->lwe_sk_small.empty() ? nullptr : ks->lwe_sk_small.data(); }
const uint64_t *cbs_keyset_glwe_sk(const cbs_keyset *ks) { return ks->glwe_sk.empty() ? nullptr : ks->glwe_sk.data(); }

int cbs_keyset_from_arrays(const uint64_t *bsk, const uint64_t *ksk, const uint64_t *auto_std, const uint64_t *ss,
                           const uint64_t *lwe_sk_small, const uint64_t *glwe_sk, cbs_keyset **out)
{
    if (!bsk || !ksk || !auto_std || !ss || !out) {
        set_error("cbs_keyset_from_arrays: null argument");
        return CBS_ERR_ARG;
    }
    auto *k = new cbs_keyset;
    k->bsk.assign(bsk, bsk + CBS_BSK_WORDS);
    k->ksk.assign(ksk, ksk + CBS_KSK_WORDS);
    k->autok.assign(auto_std, auto_std + CBS_AUTO_WORDS);
    k->ss.assign(ss, ss + CBS_SS_WORDS);
    if (lwe_sk_small) k->lwe_sk_small.assign(lwe_sk_small, lwe_sk_small + 768);
    if (glwe_sk) k->glwe_sk.assign(glwe_sk, glwe_sk + 2048);
    *out = k;
    return CBS_OK;
}

int cbs_keyset_load_dir(const char *io_dir, int with_secret, cbs_keyset **out)
{
    if (!io_dir || !out) {
        set_error("cbs_keyset_load_dir: null argument");
        return CBS_ERR_ARG;
    }
    const std::string pk = std::string(io_dir) + "/public_keys/";
    auto k = std::unique_ptr<cbs_keyset>(new cbs_keyset);
    {
        Reader r;
        if (!r.load(pk + "bsk.bin")) {
            set_error("cannot read " + pk + "bsk.bin");
            return CBS_ERR_IO;
        }
        r.vec(k->bsk, CBS_BSK_WORDS);
        r.expect(3);
        r.expect(1024);
        r.expect(23);
        r.expect(1);
        r.modulus();
        if (!r.done()) {
            set_error("bsk.bin: not an AES_TIGHT LweBootstrapKey<u64>");
            return CBS_ERR_FORMAT;
        }
    }
    {
        Reader r;
        if (!r.load(pk + "ksk.bin")) {
            set_error("cannot read " + pk + "ksk.bin");
            return CBS_ERR_IO;
        }
        r.vec(k->ksk, CBS_KSK_WORDS);
        r.expect(8);
        r.expect(3);
        r.expect(256);
        r.expect(4);
        r.expect(3);
        r.modulus();
        if (!r.done()) {
            set_error("ksk.bin: not an AES_TIGHT GlweKeyswitchKey<u64>");
            return CBS_ERR_FORMAT;
        }
    }
    {
        Reader r;
        if (!r.load(pk + "ss_key.bin")) {
            set_error("cannot read " + pk + "ss_key.bin");
            return CBS_ERR_IO;
        }
        r.vec(k->ss, CBS_SS_WORDS);
        r.expect(3);
        r.expect(1024);
        r.expect(17);
        r.expect(2);
        r.modulus();
        if (!r.done()) {
            set_error("ss_key.bin: not an AES_TIGHT GgswCiphertextList<u64>");
            return CBS_ERR_FORMAT;
        }
    }
    {
        Reader r;
        if (!r.load(pk + "auto_keys.bin")) {
            set_error("cannot read " + pk + "auto_keys.bin");
            return CBS_ERR_IO;
        }
        k->autok.assign(CBS_AUTO_WORDS, 0);
        if (r.u64() != 10) {
            set_error("auto_keys.bin: expected 10 automorphism keys");
            return CBS_ERR_FORMAT;
        }
        int seen = 0;
        std::vector<uint64_t> data, lo(1024), hi(1024);
        for (int e = 0; e < 10; e++) {
            uint64_t kappa = r.u64();
            int idx = -1;
            for (int i = 0; i < 10; i++)
                if (kappa == (uint64_t)((1024 >> i) + 1)) idx = i;
            if (idx < 0 || !r.vec(data, 36864)) {
                set_error("auto_keys.bin: malformed entry");
                return CBS_ERR_FORMAT;
            }
            r.expect(13);
            r.expect(3);
            r.expect(2);
            r.expect(1024);
            r.expect(kappa);
            // [in 2][split 2][level 3][poly 3][512 c64] -> std [in 2][level 3][poly 3][1024]
            const double *f = reinterpret_cast<const double *>(data.data());
            for (int in = 0; in < 2; in++)
                for (int p = 0; p < 9; p++) {
                    double e0 = fourier_to_limb(f + ((size_t)(in * 2 + 0) * 9 + p) * 1024, lo.data(), 0x1p41);
                    double e1 = fourier_to_limb(f + ((size_t)(in * 2 + 1) * 9 + p) * 1024, hi.data(), 0x1p23);
                    if (e0 > 0.05 || e1 > 1e-4) {
                        set_error("auto_keys.bin: Fourier data is not in natural DFT order "
                                  "(limbs do not invert to integers)");
                        return CBS_ERR_FORMAT;
                    }
                    uint64_t *dst = k->autok.data() + ((size_t)idx * 18 + in * 9 + p) * 1024;
                    for (int j = 0; j < 1024; j++) dst[j] = lo[j] | (hi[j] << 41);
                }
            seen |= 1 << idx;
        }
        if (!r.done() || seen != 1023) {
            set_error("auto_keys.bin: unexpected contents");
            return CBS_ERR_FORMAT;
        }
    }
    if (with_secret) {
        const std::string sk = std::string(io_dir) + "/secret_keys/";
        Reader r;
        if (!r.load(sk + "glwe_sk.bin") || !r.vec(k->glwe_sk, 2048)) {
            set_error("cannot read " + sk + "glwe_sk.bin");
            return CBS_ERR_IO;
        }
        Reader r2;  // optional: the small key is not written by the reference
        if (r2.load(sk + "lwe_sk_small.bin")) r2.vec(k->lwe_sk_small, 768);
    }
    *out = k.release();
    return CBS_OK;
}

int cbs_keyset_save_dir(const cbs_keyset *ks, const char *io_dir, int with_secret)
{
    if (!ks || !io_dir) {
        set_error("cbs_keyset_save_dir: null argument");
        return CBS_ERR_ARG;
    }
    const std::string pk = std::string(io_dir) + "/public_keys/";
    mkdirs(pk);
    bool ok = true;
    {
        Writer w;
        w.vec(ks->bsk.data(), CBS_BSK_WORDS);
        w.u64(3);
        w.u64(1024);
        w.u64(23);
        w.u64(1);
        w.modulus();
        ok &= w.save(pk + "bsk.bin");
    }
    {
        Writer w;
        w.vec(ks->ksk.data(), CBS_KSK_WORDS);
        w.u64(8);
        w.u64(3);
        w.u64(256);
        w.u64(4);
        w.u64(3);
        w.modulus();
        ok &= w.save(pk + "ksk.bin");
    }
    {
        Writer w;
        w.vec(ks->ss.data(), CBS_SS_WORDS);
        w.u64(3);
        w.u64(1024);
        w.u64(17);
        w.u64(2);
        w.modulus();
        ok &= w.save(pk + "ss_key.bin");
    }
    {
        Writer w;
        w.u64(10);
        std::vector<uint64_t> data(36864), limb(1024);
        for (int idx = 0; idx < 10; idx++) {
            uint64_t kappa = (uint64_t)((1024 >> idx) + 1);
            double *f = reinterpret_cast<double *>(data.data());
            for (int in = 0; in < 2; in++)
                for (int sp = 0; sp < 2; sp++)
                    for (int p = 0; p < 9; p++) {
                        const uint64_t *src = ks->autok.data() + ((size_t)idx * 18 + in * 9 + p) * 1024;
                        for (int j = 0; j < 1024; j++) limb[j] = sp ? (src[j] >> 41) : ((src[j] << 23) >> 23);
                        limb_to_fourier(limb.data(), f + ((size_t)(in * 2 + sp) * 9 + p) * 1024);
                    }
            w.u64(kappa);
            w.vec(data.data(), 36864);
            w.u64(13);
            w.u64(3);
            w.u64(2);
            w.u64(1024);
            w.u64(kappa);
        }
        ok &= w.save(pk + "auto_keys.bin");
    }
    if (with_secret && !ks->glwe_sk.empty()) {
        const std::string sk = std::string(io_dir) + "/secret_keys/";
        mkdirs(sk);
        Writer a, b, c;
        a.vec(ks->glwe_sk.data(), 2048);  // LweSecretKey (big) == flattened GlweSecretKey
        ok &= a.save(sk + "lwe_sk.bin");
        b.vec(ks->glwe_sk.data(), 2048);
        b.u64(1024);
        ok &= b.save(sk + "glwe_sk.bin");
        if (!ks->lwe_sk_small.empty()) {
            c.vec(ks->lwe_sk_small.data(), 768);
            ok &= c.save(sk + "lwe_sk_small.bin");
        }
    }
    if (!ok) {
        set_error(std::string("cannot write key files under ") + io_dir);
        return CBS_ERR_IO;
    }
    return CBS_OK;
}

static int keyset_generate_impl(const RngSource &src, cbs_keyset **out)
{
    if (!out) return CBS_ERR_ARG;
    auto *k = new cbs_keyset;
    k->bsk.assign(CBS_BSK_WORDS, 0);
    k->ksk.assign(CBS_KSK_WORDS, 0);
    k->autok.assign(CBS_AUTO_WORDS, 0);
    k->ss.assign(CBS_SS_WORDS, 0);
    k->lwe_sk_small.resize(768);
    k->glwe_sk.resize(2048);
    Rng srng(src, 0);
    for (auto &b : k->lwe_sk_small) b = srng.next() >> 63;
    for (auto &b : k->glwe_sk) b = srng.next() >> 63;
    const uint64_t *S = k->glwe_sk.data();

    // bootstrap key: GGSW(s_i) under glwe_sk, B = 2^23, l = 1 (keygen.rs:213-221)
    parallel_for(768, [&](int i) {
        Rng rng(src, 1000 + (uint64_t)i);
        const uint64_t m = k->lwe_sk_small[i];
        std::vector<uint64_t> pt(1024);
        for (int row = 0; row < 3; row++) {
            std::fill(pt.begin(), pt.end(), 0ull);
            if (m) {
                if (row < 2)
                    for (int j = 0; j < 1024; j++) pt[j] = (0ull - S[row * 1024 + j]) << 41;  // -s_i * S_row * 2^(64-23)
                else
                    pt[0] = 1ull << 41;
            }
            glwe_encrypt(k->bsk.data() + ((size_t)i * 3 + row) * 3072, pt.data(), S, 2, 1024, kStdGlwe, rng);
        }
    });

    // LWE keyswitch key over the common ring N' = 256: big key (8 polys) -> small key (3 polys),
    // GLEV of -S_in,i, B = 2^4, l = 3, LWE noise (keygen.rs:224-240, glwe_keyswitch.rs:183-218)
    {
        std::vector<uint64_t> pt(256);
        for (int i = 0; i < 8; i++)
            for (int lev = 0; lev < 3; lev++) {
                Rng rng(src, 5000 + (uint64_t)(i * 3 + lev));
                const int log_scale = 64 - 4 * (lev + 1);
                for (int j = 0; j < 256; j++) pt[j] = (0ull - S[i * 256 + j]) << log_scale;
                glwe_encrypt(k->ksk.data() + ((size_t)i * 3 + lev) * 4 * 256, pt.data(), k->lwe_sk_small.data(), 3, 256,
                             kStdLwe, rng);
            }
    }

    // automorphism keys: keyswitch from S(X^kappa) back to S, B = 2^13, l = 3 (automorphism.rs:59-178)
    parallel_for(10, [&](int idx) {
        const unsigned kappa = (unsigned)((1024 >> idx) + 1);
        std::vector<uint64_t> before(1024), pt(1024);
        for (int i = 0; i < 2; i++) {
            eval_x_k(before.data(), S + i * 1024, 1024, kappa);
            for (int lev = 0; lev < 3; lev++) {
                Rng rng(src, 6000 + (uint64_t)((idx * 2 + i) * 3 + lev));
                const int log_scale = 64 - 13 * (lev + 1);
                for (int j = 0; j < 1024; j++) pt[j] = (0ull - before[j]) << log_scale;
                glwe_encrypt(k->autok.data() + (((size_t)idx * 2 + i) * 3 + lev) * 3072, pt.data(), S, 2, 1024, kStdGlwe,
                             rng);
            }
        }
    });

    // scheme-switching key: GGSW(0) then -S_i * 2^(64-17*lev) added to mask poly `col` (col < 2) or to
    // the body (col == 2) of row (lev, col)  (ggsw_conv.rs:41-74)
    for (int i = 0; i < 2; i++)
        for (int lev = 0; lev < 2; lev++)
            for (int col = 0; col < 3; col++) {
                Rng rng(src, 7000 + (uint64_t)((i * 2 + lev) * 3 + col));
                uint64_t *ct = k->ss.data() + (((size_t)i * 2 + lev) * 3 + col) * 3072;
                glwe_encrypt(ct, nullptr, S, 2, 1024, kStdGlwe, rng);
                const int log_scale = 64 - 17 * (lev + 1);
                uint64_t *dst = ct + (size_t)col * 1024;
                for (int j = 0; j < 1024; j++) dst[j] += (0ull - S[i * 1024 + j]) << log_scale;
            }
    *out = k;
    return CBS_OK;
}

int cbs_keyset_generate(uint64_t seed, cbs_keyset **out) { return keyset_generate_impl(RngSource(seed), out); }

int cbs_keyset_generate_os_entropy(cbs_keyset **out)
{
    RngSource src(0);
    if (!RngSource::os_entropy(src)) {
        set_error("getrandom failed");
        return CBS_ERR_IO;
    }
    return keyset_generate_impl(src, out);
}

static int trans_key_generate_impl(const cbs_keyset *ks, const uint8_t aes_key[16], const RngSource &src, uint64_t *k10_9,
                                   uint64_t *k8_1, uint64_t *k0)
{
    if (!ks || ks->glwe_sk.empty() || !aes_key || !k10_9 || !k8_1 || !k0) {
        set_error("cbs_trans_key_generate: needs a keyset with the GLWE secret key");
        return CBS_ERR_ARG;
    }
    AesTables aes(aes_key);
    static const uint8_t mults[4] = {9, 11, 13, 14};  // AllRdKeys tuple order, data_struct.rs:76-81
    uint8_t tab[16][256];
    aes.lut_10_9(tab);
    const uint64_t *S = ks->glwe_sk.data();
    parallel_for(4 * 16 * 2, [&](int id) {
        const int a = id & 1, b = (id >> 1) & 15, m = id >> 5;
        Rng rng(src, 9000 + (uint64_t)id);
        std::vector<uint64_t> pt(1024);
        lut_plaintext(pt.data(), tab[b], a, mults[m]);
        glwe_encrypt(k10_9 + (size_t)id * 3072, pt.data(), S, 2, 1024, kStdGlwe, rng);
    });
    memset(k8_1, 0, sizeof(uint64_t) * CBS_K8_1_WORDS);
    for (int r = 1; r <= 8; r++) {
        aes.lut_round(r, tab);
        for (int m = 0; m < 4; m++)
            for (int b = 0; b < 16; b++)
                for (int a = 0; a < 2; a++)
                    lut_plaintext(k8_1 + ((((size_t)(r - 1) * 4 + m) * 16 + b) * 2 + a) * 3072 + 2048, tab[b], a, mults[m]);
    }
    memset(k0, 0, sizeof(uint64_t) * CBS_K0_WORDS);
    aes.lut_round(0, tab);
    for (int b = 0; b < 16; b++)
        for (int a = 0; a < 2; a++) lut_plaintext(k0 + ((size_t)b * 2 + a) * 3072 + 2048, tab[b], a, 0);
    return CBS_OK;
}

int cbs_trans_key_generate(const cbs_keyset *ks, const uint8_t aes_key[16], uint64_t seed, uint64_t *k10_9,
                           uint64_t *k8_1, uint64_t *k0)
{
    return trans_key_generate_impl(ks, aes_key, RngSource(seed), k10_9, k8_1, k0);
}

int cbs_trans_key_generate_os_entropy(const cbs_keyset *ks, const uint8_t aes_key[16], uint64_t *k10_9, uint64_t *k8_1,
                                      uint64_t *k0)
{
    RngSource src(0);
    if (!RngSource::os_entropy(src)) {
        set_error("getrandom failed");
        return CBS_ERR_IO;
    }
    return trans_key_generate_impl(ks, aes_key, src, k10_9, k8_1, k0);
}

int cbs_trans_key_load(const char *path, uint64_t *k10_9, uint64_t *k8_1, uint64_t *k0)
{
    Reader r;
    if (!path || !r.load(path)) {
        set_error(std::string("cannot read ") + (path ? path : "(null)"));
        return CBS_ERR_IO;
    }
    for (int m = 0; m < 4; m++) {
        r.expect(16);
        for (int b = 0; b < 16; b++) {  // GlweCiphertextList { data, glwe_size, poly, modulus }
            r.vec_into(k10_9 + ((size_t)m * 16 + b) * 2 * 3072, 2 * 3072);
            r.expect(3);
            r.expect(1024);
            r.modulus();
        }
    }
    r.expect(8);
    for (int rd = 0; rd < 8; rd++)
        for (int m = 0; m < 4; m++) {
            r.expect(16);
            for (int b = 0; b < 16; b++) {
                r.expect(2);
                for (int a = 0; a < 2; a++) {  // GlweCiphertext { data, poly, modulus }
                    r.vec_into(k8_1 + ((((size_t)rd * 4 + m) * 16 + b) * 2 + a) * 3072, 3072);
                    r.expect(1024);
                    r.modulus();
                }
            }
        }
    r.expect(16);
    for (int b = 0; b < 16; b++) {
        r.expect(2);
        for (int a = 0; a < 2; a++) {
            r.vec_into(k0 + ((size_t)b * 2 + a) * 3072, 3072);
            r.expect(1024);
            r.modulus();
        }
    }
    if (!r.done()) {
        set_error(std::string(path) + ": not an AllRdKeys file for AES_TIGHT");
        return CBS_ERR_FORMAT;
    }
    return CBS_OK;
}

int cbs_trans_key_save(const char *path, const uint64_t *k10_9, const uint64_t *k8_1, const uint64_t *k0)
{
    if (!path) return CBS_ERR_ARG;
    Writer w;
    w.buf.reserve(29200000);
    for (int m = 0; m < 4; m++) {
        w.u64(16);
        for (int b = 0; b < 16; b++) {
            w.vec(k10_9 + ((size_t)m * 16 + b) * 2 * 3072, 2 * 3072);
            w.u64(3);
            w.u64(1024);
            w.modulus();
        }
    }
    w.u64(8);
    for (int rd = 0; rd < 8; rd++)
        for (int m = 0; m < 4; m++) {
            w.u64(16);
            for (int b = 0; b < 16; b++) {
                w.u64(2);
                for (int a = 0; a < 2; a++) {
                    w.vec(k8_1 + ((((size_t)rd * 4 + m) * 16 + b) * 2 + a) * 3072, 3072);
                    w.u64(1024);
                    w.modulus();
                }
            }
        }
    w.u64(16);
    for (int b = 0; b < 16; b++) {
        w.u64(2);
        for (int a = 0; a < 2; a++) {
            w.vec(k0 + ((size_t)b * 2 + a) * 3072, 3072);
            w.u64(1024);
            w.modulus();
        }
    }
    std::string p(path);
    size_t slash = p.find_last_of('/');
    if (slash != std::string::npos) mkdirs(p.substr(0, slash));
    if (!w.save(p)) {
        set_error("cannot write " + p);
        return CBS_ERR_IO;
    }
    return CBS_OK;
}

// ---- forward direction (CTR mode) ----
static int fwd_trans_key_generate_impl(const cbs_keyset *ks, const uint8_t aes_key[16], const RngSource &src, uint64_t *kf_first,
                               uint64_t *kf_mid, uint64_t *kf_last)
{
    if (!ks || ks->glwe_sk.empty() || !aes_key || !kf_first || !kf_mid || !kf_last) {
        set_error("cbs_fwd_trans_key_generate: needs a keyset with the GLWE secret key");
        return CBS_ERR_ARG;
    }
    AesTables aes(aes_key);
    static const uint8_t mults[3] = {1, 2, 3};
    uint8_t tab[16][256];
    aes.lut_fwd(0, tab);
    const uint64_t *S = ks->glwe_sk.data();
    parallel_for(3 * 16 * 2, [&](int id) {
        const int a = id & 1, b = (id >> 1) & 15, m = id >> 5;
        Rng rng(src, 11000 + (uint64_t)id);
        std::vector<uint64_t> pt(1024);
        lut_plaintext(pt.data(), tab[b], a, mults[m] == 1 ? 0 : mults[m]);
        glwe_encrypt(kf_first + (size_t)id * 3072, pt.data(), S, 2, 1024, kStdGlwe, rng);
    });
    memset(kf_mid, 0, sizeof(uint64_t) * CBS_KF_MID_WORDS);
    for (int r = 2; r <= 9; r++) {
        aes.lut_fwd(r - 1, tab);
        for (int m = 0; m < 3; m++)
            for (int b = 0; b < 16; b++)
                for (int a = 0; a < 2; a++)
                    lut_plaintext(kf_mid + ((((size_t)(r - 2) * 3 + m) * 16 + b) * 2 + a) * 3072 + 2048, tab[b], a,
                                  mults[m] == 1 ? 0 : mults[m]);
    }
    memset(kf_last, 0, sizeof(uint64_t) * CBS_KF_LAST_WORDS);
    aes.lut_fwd_last(tab);
    for (int b = 0; b < 16; b++)
        for (int a = 0; a < 2; a++) lut_plaintext(kf_last + ((size_t)b * 2 + a) * 3072 + 2048, tab[b], a, 0);
    return CBS_OK;
}

// Forward keys use the AllRdKeys bincode shape with 3-tuples (x1, x2, x3) and 8 middle rounds.
int cbs_fwd_trans_key_generate(const cbs_keyset *ks, const uint8_t aes_key[16], uint64_t seed, uint64_t *kf_first,
                               uint64_t *kf_mid, uint64_t *kf_last)
{
    return fwd_trans_key_generate_impl(ks, aes_key, RngSource(seed), kf_first, kf_mid, kf_last);
}

int cbs_fwd_trans_key_generate_os_entropy(const cbs_keyset *ks, const uint8_t aes_key[16], uint64_t *kf_first, uint64_t *kf_mid,
                                          uint64_t *kf_last)
{
    RngSource src(0);
    if (!RngSource::os_entropy(src)) {
        set_error("getrandom failed");
        return CBS_ERR_IO;
    }
    return fwd_trans_key_generate_impl(ks, aes_key, src, kf_first, kf_mid, kf_last);
}

int cbs_fwd_trans_key_save(const char *path, const uint64_t *kf_first, const uint64_t *kf_mid, const uint64_t *kf_last)
{
    if (!path) return CBS_ERR_ARG;
    Writer w;
    w.buf.reserve(22000000);
    for (int m = 0; m < 3; m++) {
        w.u64(16);
        for (int b = 0; b < 16; b++) {
            w.vec(kf_first + ((size_t)m * 16 + b) * 2 * 3072, 2 * 3072);
            w.u64(3);
            w.u64(1024);
            w.modulus();
        }
    }
    w.u64(8);
    for (int rd = 0; rd < 8; rd++)
        for (int m = 0; m < 3; m++) {
            w.u64(16);
            for (int b = 0; b < 16; b++) {
                w.u64(2);
                for (int a = 0; a < 2; a++) {
                    w.vec(kf_mid + ((((size_t)rd * 3 + m) * 16 + b) * 2 + a) * 3072, 3072);
                    w.u64(1024);
                    w.modulus();
                }
            }
        }
    w.u64(16);
    for (int b = 0; b < 16; b++) {
        w.u64(2);
        for (int a = 0; a < 2; a++) {
            w.vec(kf_last + ((size_t)b * 2 + a) * 3072, 3072);
            w.u64(1024);
            w.modulus();
        }
    }
    std::string p(path);
    size_t slash = p.find_last_of('/');
    if (slash != std::string::npos) mkdirs(p.substr(0, slash));
    if (!w.save(p)) {
        set_error("cannot write " + p);
        return CBS_ERR_IO;
    }
    return CBS_OK;
}

int cbs_fwd_trans_key_load(const char *path, uint64_t *kf_first, uint64_t *kf_mid, uint64_t *kf_last)
{
    Reader r;
    if (!path || !r.load(path)) {
        set_error(std::string("cannot read ") + (path ? path : "(null)"));
        return CBS_ERR_IO;
    }
    for (int m = 0; m < 3; m++) {
        r.expect(16);
        for (int b = 0; b < 16; b++) {
            r.vec_into(kf_first + ((size_t)m * 16 + b) * 2 * 3072, 2 * 3072);
            r.expect(3);
            r.expect(1024);
            r.modulus();
        }
    }
    r.expect(8);
    for (int rd = 0; rd < 8; rd++)
        for (int m = 0; m < 3; m++) {
            r.expect(16);
            for (int b = 0; b < 16; b++) {
                r.expect(2);
                for (int a = 0; a < 2; a++) {
                    r.vec_into(kf_mid + ((((size_t)rd * 3 + m) * 16 + b) * 2 + a) * 3072, 3072);
                    r.expect(1024);
                    r.modulus();
                }
            }
        }
    r.expect(16);
    for (int b = 0; b < 16; b++) {
        r.expect(2);
        for (int a = 0; a < 2; a++) {
            r.vec_into(kf_last + ((size_t)b * 2 + a) * 3072, 3072);
            r.expect(1024);
            r.modulus();
        }
    }
    if (!r.done()) {
        set_error(std::string(path) + ": not a forward (CTR) transciphering key for AES_TIGHT");
        return CBS_ERR_FORMAT;
    }
    return CBS_OK;
}

int cbs_lwe_list_load(const char *path, uint64_t **data, uint64_t *count, uint64_t *lwe_words)
{
    Reader r;
    if (!path || !data || !count || !lwe_words || !r.load(path)) {
        set_error(std::string("cannot read ") + (path ? path : "(null)"));
        return CBS_ERR_IO;
    }
    uint64_t n = r.u64();
    // n comes from the file: bound it by the bytes actually present before any multiplication can wrap
    if (!r.ok || r.buf.size() < r.off + 32 || n != (r.buf.size() - r.off - 32) / 8 || (r.buf.size() - r.off - 32) % 8) {
        set_error(std::string(path) + ": not an LweCiphertextList<u64>");
        return CBS_ERR_FORMAT;
    }
    uint64_t *p = (uint64_t *)malloc(8 * (n ? n : 1));
    memcpy(p, r.buf.data() + r.off, 8 * n);
    r.off += 8 * n;
    uint64_t lw = r.u64();
    r.modulus();
    if (!r.done() || lw == 0 || n % lw) {
        free(p);
        set_error(std::string(path) + ": malformed LweCiphertextList trailer");
        return CBS_ERR_FORMAT;
    }
    *data = p;
    *count = n / lw;
    *lwe_words = lw;
    return CBS_OK;
}

int cbs_lwe_list_save(const char *path, const uint64_t *data, uint64_t count, uint64_t lwe_words)
{
    if (!path || (!data && count)) return CBS_ERR_ARG;
    Writer w;
    w.vec(data, count * lwe_words);
    w.u64(lwe_words);
    w.modulus();
    std::string p(path);
    size_t slash = p.find_last_of('/');
    if (slash != std::string::npos) mkdirs(p.substr(0, slash));
    if (!w.save(p)) {
        set_error("cannot write " + p);
        return CBS_ERR_IO;
    }
    return CBS_OK;
}

// decrypt_decode_lwe_list / decrypt, submission/src/help_fun.rs:12-42 (delta = 2^63)
int cbs_lwe_decrypt_bits(const uint64_t *sk, int n, const uint64_t *lwe, uint64_t count, uint64_t *bits_out)
{
    if (!sk || !lwe || !bits_out || n <= 0) return CBS_ERR_ARG;
    const uint64_t delta = 1ull << 63;
    for (uint64_t c = 0; c < count; c++) {
        const uint64_t *ct = lwe + c * (uint64_t)(n + 1);
        uint64_t phase = ct[n];
        for (int i = 0; i < n; i++)
            if (sk[i]) phase -= ct[i];
        const uint64_t rounding = (phase & (delta >> 1)) << 1;
        bits_out[c] = (phase + rounding) / delta;
    }
    return CBS_OK;
}

// bincode Vec<u64> (the intermediate/decoded_result*.txt files of the client stages)
int cbs_u64_vec_save(const char *path, const uint64_t *data, uint64_t n)
{
    if (!path || (!data && n)) return CBS_ERR_ARG;
    Writer w;
    w.vec(data, n);
    std::string p(path);
    size_t slash = p.find_last_of('/');
    if (slash != std::string::npos) mkdirs(p.substr(0, slash));
    if (!w.save(p)) {
        set_error("cannot write " + p);
        return CBS_ERR_IO;
    }
    return CBS_OK;
}

int cbs_u64_vec_load(const char *path, uint64_t **data, uint64_t *n)
{
    Reader r;
    if (!path || !data || !n || !r.load(path)) {
        set_error(std::string("cannot read ") + (path ? path : "(null)"));
        return CBS_ERR_IO;
    }
    uint64_t len = r.u64();
    if (!r.ok || r.off + 8 * len != r.buf.size()) {
        set_error(std::string(path) + ": not a bincode Vec<u64>");
        return CBS_ERR_FORMAT;
    }
    uint64_t *p = (uint64_t *)malloc(8 * (len ? len : 1));
    memcpy(p, r.buf.data() + r.off, 8 * len);
    *data = p;
    *n = len;
    return CBS_OK;
}

// dry run of the inner-product circuit plan on cleartext values (test hook of host/ip_plan.h)
int cbs_inner_product_plan_check(const uint16_t *vals, int nvals, uint16_t *result, int64_t *circuit_bootstraps, int *layers,
                                 int64_t *lut_ladders)
{
    if (nvals <= 0 || (nvals & 1) || (!vals && result)) {
        set_error("cbs_inner_product_plan_check: bad argument (need an even number of values)");
        return CBS_ERR_ARG;
    }
    const IpPlan plan = ip_make_plan(nvals);
    if (result) *result = ip_eval_clear(plan, vals);
    if (circuit_bootstraps) *circuit_bootstraps = plan.total_cbs();
    if (layers) *layers = (int)plan.layers.size();
    if (lut_ladders) {
        int64_t n = 0;
        for (const IpLayer &l : plan.layers) n += (int64_t)l.jobs.size();
        *lut_ladders = n;
    }
    return CBS_OK;
}

int cbs_max_plan_check(const uint16_t *vals, int nvals, uint16_t *result, int64_t *circuit_bootstraps, int *layers,
                       int64_t *lut_ladders)
{
    if (nvals <= 0 || (!vals && result)) {
        set_error("cbs_max_plan_check: bad argument");
        return CBS_ERR_ARG;
    }
    const IpPlan plan = max_make_plan(nvals);
    if (result) *result = ip_eval_clear(plan, vals);
    if (circuit_bootstraps) *circuit_bootstraps = plan.total_cbs();
    if (layers) *layers = (int)plan.layers.size();
    if (lut_ladders) {
        int64_t n = 0;
        for (const IpLayer &l : plan.layers) n += (int64_t)l.jobs.size();
        *lut_ladders = n;
    }
    return CBS_OK;
}

static int encrypt_bits(const uint64_t *sk, int n, double std, const uint8_t *bits, int count, uint64_t seed, uint64_t *out)
{
    const RngSource src(seed);
    for (int c = 0; c < count; c++) {
        Rng rng(src, 20000 + (uint64_t)c);
        uint64_t *ct = out + (size_t)c * (n + 1);
        uint64_t b = rng.noise(std) + ((uint64_t)(bits[c] & 1) << 63);
        for (int i = 0; i < n; i++) {
            ct[i] = rng.next();
            if (sk[i]) b += ct[i];
        }
        ct[n] = b;
    }
    return CBS_OK;
}

int cbs_encrypt_bits_big(const cbs_keyset *ks, const uint8_t *bits, int count, uint64_t seed, uint64_t *out)
{
    if (!ks || ks->glwe_sk.empty() || !bits || !out) {
        set_error("cbs_encrypt_bits_big: needs a keyset with secret keys");
        return CBS_ERR_ARG;
    }
    return encrypt_bits(ks->glwe_sk.data(), 2048, kStdGlwe, bits, count, seed, out);
}

int cbs_encrypt_bits_small(const cbs_keyset *ks, const uint8_t *bits, int count, uint64_t seed, uint64_t *out)
{
    if (!ks || ks->lwe_sk_small.empty() || !bits || !out) {
        set_error("cbs_encrypt_bits_small: needs a keyset with secret keys");
        return CBS_ERR_ARG;
    }
    return encrypt_bits(ks->lwe_sk_small.data(), 768, kStdLwe, bits, count, seed, out);
}

}  // extern "C"
