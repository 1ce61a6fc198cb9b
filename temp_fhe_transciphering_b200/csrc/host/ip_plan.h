// ip_plan.h — circuit plan of the encrypted inner product mod 2^16 (mini-workload #2 of the harness,
// harness/cleartext_impl.py:65-70: sum((x*y) % 2^16 for x, y in zip(first_half, second_half)) % 2^16).
//
// The reference submission has no implementation of this workload; the semantics are the
// harness's cleartext formula.  The circuit is built only from operations the transciphering path already
// has: LWE keyswitch -> circuit bootstrap of a bit -> 8-selector LUT ladder (evaluate_8_to_8_lut,
// cbs_lib/src/aes_he.rs:791-832).  Every value is 16 LWE bits (MSB first, delta = 2^63), exactly the output
// format of stage 7.  Arithmetic precision beyond one bit per ciphertext is not available at the noise level of
// this parameter set (LUT outputs carry ~2^57 of noise under a 2^63 bit), so the sum is a Boolean circuit:
//
//   layer 0  nibble products: x = sum_k x_k 16^k, y = sum_l y_l 16^l; the 10 products x_k * y_l with k + l <= 3
//            are one LUT each (4 + 4 selector bits -> 8 product bits, 4 for k + l = 3).  32 circuit bootstraps
//            and 16 accumulator ladders per pair; the product bits are filed under their weight ("columns").
//   phase A  column compression: while some column holds more than 2 bits, every column is cut into chunks of
//            7 bits (remainder >= 3: one smaller chunk; <= 2: passed on untouched, no bootstrap) and a chunk
//            is replaced by its population count (<= 3 bits, one LUT): 4 bits removed per 7 bootstraps.
//   phase B  when all columns hold <= 2 bits: nibble-wise addition, (a3..a0) + (b3..b0) -> 5 bits in one
//            LUT (two accumulators); the carries ripple over at most 4 layers.
// The plan is data independent, so it is computed once on the host; the GPU executes it layer by layer
// (cbs_inner_product_u16 in cbs_api.cu).  ip_eval_clear runs the same plan on cleartext bits and is
// what the CPU tests use to check the circuit itself against the harness formula.
#pragma once
#include <cstdint>
#include <vector>

namespace cbs_host {

// LUT tables of the plans; each yields <= 8 output bits = two accumulators of 4 bits
enum IpLut : int { kIpMul = 0, kIpPop = 1, kIpAdd4 = 2, kMaxCmp = 3, kMaxSel = 4, kIpNumLuts = 5 };

inline unsigned ip_lut_value(int lut, unsigned idx)
{
    switch (lut) {
        case kIpMul: return (idx & 15u) * (idx >> 4);
        case kIpPop: return (unsigned)__builtin_popcount(idx);
        case kIpAdd4: return (idx & 15u) + (idx >> 4);
        case kMaxCmp: {  // nibbles a = idx & 15, b = idx >> 4  ->  bit 0: a > b, bit 1: a == b
            const unsigned a = idx & 15u, b = idx >> 4;
            return (a > b ? 1u : 0u) | (a == b ? 2u : 0u);
        }
        default: {  // kMaxSel: selectors (gt0, gt1, eq1, gt2, eq2, gt3, eq3, d) -> d AND [a > b]
            const unsigned gt0 = idx & 1, gt1 = (idx >> 1) & 1, eq1 = (idx >> 2) & 1, gt2 = (idx >> 3) & 1, eq2 = (idx >> 4) & 1,
                           gt3 = (idx >> 5) & 1, eq3 = (idx >> 6) & 1, d = (idx >> 7) & 1;
            return d & (gt3 | (eq3 & (gt2 | (eq2 & (gt1 | (eq1 & gt0))))));
        }
    }
}

struct IpJob {
    int sel[8];  // position in the layer's bootstrap list (selector i <-> bit i of the table index), -1 = constant 0
    int lut;     // IpLut
    int acc;     // accumulator: output bits 4*acc .. 4*acc + 3 of the table value
    int out;     // pool index of the first of the 4 output bits
};

struct IpAdd {
    int dst, a, b;  // pool[dst] = pool[a] + pool[b]  (LWE addition = XOR of the encrypted bits at delta 2^63)
};
struct IpRefresh {
    int pos, dst;  // pool[dst] = a fresh encryption of the bit bootstrapped at position `pos` of this layer
};

struct IpLayer {
    std::vector<IpAdd> pre;  // executed before the layer's keyswitch
    std::vector<int> cbs;    // pool indices of the bits that are keyswitched + circuit-bootstrapped in this layer
    std::vector<IpRefresh> refresh;  // fresh LWEs taken from the bootstrap's level-1 GLEV (2 x, sample 0)
    std::vector<IpJob> jobs;
    std::vector<IpAdd> post;  // executed after the layer's ladders
};

struct IpPlan {
    int nvals = 0;
    int pool_size = 0;  // LWE slots: inputs first ([value][bit MSB first]), then LUT outputs
    std::vector<IpLayer> layers;
    int nin_bits = 0;   // leading pool slots holding the inputs
    int result[16];     // pool index of result bit with weight 2^c, or -1 for a constant 0
    long total_cbs() const
    {
        long n = 0;
        for (const IpLayer &l : layers) n += (long)l.cbs.size();
        return n;
    }
};

// Column compression + final addition (phases A and B above): reduces `col` (col[c] = pool indices of the bits of
// weight 2^c) to one bit per column, appending the layers to `plan`.
inline void ip_compress_columns(IpPlan &plan, std::vector<std::vector<int>> &col, int &pool)
{
    auto alloc4 = [&]() {
        int b = pool;
        pool += 4;
        return b;
    };
    for (;;) {
        size_t maxh = 0;
        for (int c = 0; c < 16; c++) maxh = col[c].size() > maxh ? col[c].size() : maxh;
        if (maxh <= 1) break;
        IpLayer L;
        std::vector<std::vector<int>> next(16);
        auto use = [&](int pool_idx) {
            L.cbs.push_back(pool_idx);
            return (int)L.cbs.size() - 1;
        };
        if (maxh > 2) {  // phase A: population counts inside a column
            for (int c = 0; c < 16; c++) {
                const std::vector<int> &bits = col[c];
                size_t pos = 0;
                while (pos < bits.size()) {
                    const size_t left = bits.size() - pos;
                    const size_t take = left >= 7 ? 7 : left;
                    if (take <= 2) {
                        for (; pos < bits.size(); pos++) next[c].push_back(bits[pos]);
                        break;
                    }
                    IpJob j;
                    for (int i = 0; i < 8; i++) j.sel[i] = i < (int)take ? use(bits[pos + i]) : -1;
                    pos += take;
                    j.lut = kIpPop;
                    j.acc = 0;
                    j.out = alloc4();
                    L.jobs.push_back(j);
                    const int nout = take >= 4 ? 3 : 2;  // bits of the count (take <= 7)
                    for (int q = 0; q < nout; q++)
                        if (c + q < 16) next[c + q].push_back(j.out + q);
                }
            }
        } else {  // phase B: 4-bit + 4-bit additions
            for (int nib = 0; nib < 4; nib++) {
                const int c0 = 4 * nib;
                bool two = false;
                for (int i = 0; i < 4; i++) two = two || col[c0 + i].size() == 2;
                if (!two) {
                    for (int i = 0; i < 4; i++)
                        for (int b : col[c0 + i]) next[c0 + i].push_back(b);
                    continue;
                }
                IpJob j;
                for (int i = 0; i < 4; i++) {
                    j.sel[i] = col[c0 + i].size() >= 1 ? use(col[c0 + i][0]) : -1;
                    j.sel[4 + i] = col[c0 + i].size() >= 2 ? use(col[c0 + i][1]) : -1;
                }
                j.lut = kIpAdd4;
                j.acc = 0;
                j.out = alloc4();
                L.jobs.push_back(j);
                for (int q = 0; q < 4; q++) next[c0 + q].push_back(j.out + q);
                if (c0 + 4 < 16) {
                    j.acc = 1;
                    j.out = alloc4();
                    L.jobs.push_back(j);
                    next[c0 + 4].push_back(j.out);
                }
            }
        }
        col.swap(next);
        plan.layers.push_back(std::move(L));
    }
    for (int c = 0; c < 16; c++) plan.result[c] = col[c].empty() ? -1 : col[c][0];
    plan.pool_size = pool;
}

inline IpPlan ip_make_plan(int nvals)
{
    IpPlan plan;
    plan.nvals = nvals;
    plan.nin_bits = nvals * 16;
    const int P = nvals / 2;
    int pool = nvals * 16;
    std::vector<std::vector<int>> col(16);
    auto in_bit = [](int value, int weight) { return value * 16 + (15 - weight); };

    // layer 0: nibble products
    if (P > 0) {
        IpLayer L;
        L.cbs.resize((size_t)nvals * 16);
        for (int i = 0; i < nvals * 16; i++) L.cbs[i] = i;  // position == pool index
        for (int i = 0; i < P; i++)
            for (int k = 0; k < 4; k++)
                for (int l = 0; k + l < 4; l++) {
                    IpJob j;
                    for (int b = 0; b < 4; b++) {
                        j.sel[b] = in_bit(i, 4 * k + b);
                        j.sel[4 + b] = in_bit(P + i, 4 * l + b);
                    }
                    j.lut = kIpMul;
                    for (int a = 0; a < 2; a++) {
                        const int c0 = 4 * (k + l) + 4 * a;
                        if (c0 >= 16) break;
                        j.acc = a;
                        j.out = pool;
                        pool += 4;
                        L.jobs.push_back(j);
                        for (int q = 0; q < 4; q++) col[c0 + q].push_back(j.out + q);
                    }
                }
        plan.layers.push_back(std::move(L));
    }
    ip_compress_columns(plan, col, pool);
    return plan;
}

// Sum of nvals 16-bit values mod 2^16: the compression stage alone.  Combines the per-GPU partial inner products when
// the pairs are sharded across GPUs (each GPU's partial sum is 16 bit ciphertexts).
inline IpPlan sum_make_plan(int nvals)
{
    IpPlan plan;
    plan.nvals = nvals;
    plan.nin_bits = nvals * 16;
    int pool = nvals * 16;
    std::vector<std::vector<int>> col(16);
    for (int v = 0; v < nvals; v++)
        for (int w = 0; w < 16; w++) col[(size_t)w].push_back(v * 16 + (15 - w));
    ip_compress_columns(plan, col, pool);
    return plan;
}

// Maximum of nvals 16-bit values as a LUT circuit (used above 8 values, where the reference's max_of_two CMux
// ladder cannot run at all and its data-dependent noise - 2^60.2 under the 2^62 decoding bound when the operands share a
// long prefix - becomes a real failure probability).  Balanced tree; one max_of_two(a, b) is two layers:
//   layer A  bootstrap the 32 bits; 4 ladders compare the nibbles -> (a_k > b_k, a_k == b_k); the b bits are also
//            taken back as FRESH ciphertexts from the same bootstraps (refresh);
//   layer B  d_i = a_i xor b_i (LWE additions of the inputs); bootstrap the 7 compare bits and the 16 d_i; one ladder
//            per output bit with selectors (gt0, gt1, eq1, gt2, eq2, gt3, eq3, d_i) -> z_i = d_i and [a > b];
//            out_i = fresh b_i + z_i  (= b_i xor z_i): a_i if a > b, else b_i.
// 55 bootstraps per max_of_two; every output is a fresh LUT result plus a fresh operand, so the noise (2^58) depends
// neither on the data nor on the depth of the tree.
inline IpPlan max_make_plan(int nvals)
{
    IpPlan plan;
    plan.nvals = nvals;
    plan.nin_bits = nvals * 16;
    int pool = nvals * 16;
    // a value = 16 pool indices, weight 2^w at [w]
    std::vector<std::vector<int>> cur((size_t)nvals, std::vector<int>(16));
    for (int v = 0; v < nvals; v++)
        for (int w = 0; w < 16; w++) cur[(size_t)v][(size_t)w] = v * 16 + (15 - w);
    while (cur.size() > 1) {
        IpLayer A, B;
        const size_t npairs = cur.size() / 2;
        std::vector<std::vector<int>> next;
        struct PairTmp {
            int fresh_b[16], d[16], cmp[7];
        };
        std::vector<PairTmp> tmp(npairs);
        for (size_t p = 0; p < npairs; p++) {
            const std::vector<int> &a = cur[2 * p], &b = cur[2 * p + 1];
            int pa[16], pb[16];
            for (int w = 0; w < 16; w++) {
                A.cbs.push_back(a[(size_t)w]);
                pa[w] = (int)A.cbs.size() - 1;
                A.cbs.push_back(b[(size_t)w]);
                pb[w] = (int)A.cbs.size() - 1;
                tmp[p].fresh_b[w] = pool++;
                A.refresh.push_back(IpRefresh{pb[w], tmp[p].fresh_b[w]});
                tmp[p].d[w] = pool++;
                B.pre.push_back(IpAdd{tmp[p].d[w], a[(size_t)w], b[(size_t)w]});
            }
            for (int k = 0; k < 4; k++) {
                IpJob j;
                for (int i = 0; i < 4; i++) {
                    j.sel[i] = pa[4 * k + i];
                    j.sel[4 + i] = pb[4 * k + i];
                }
                j.lut = kMaxCmp;
                j.acc = 0;
                j.out = pool;
                pool += 4;
                A.jobs.push_back(j);
                // cmp order (gt0, gt1, eq1, gt2, eq2, gt3, eq3)
                if (k == 0) tmp[p].cmp[0] = j.out;
                else {
                    tmp[p].cmp[2 * k - 1] = j.out;
                    tmp[p].cmp[2 * k] = j.out + 1;
                }
            }
        }
        for (size_t p = 0; p < npairs; p++) {
            int pc[7];
            for (int i = 0; i < 7; i++) {
                B.cbs.push_back(tmp[p].cmp[i]);
                pc[i] = (int)B.cbs.size() - 1;
            }
            std::vector<int> outv(16);
            for (int w = 0; w < 16; w++) {
                B.cbs.push_back(tmp[p].d[w]);
                IpJob j;
                for (int i = 0; i < 7; i++) j.sel[i] = pc[i];
                j.sel[7] = (int)B.cbs.size() - 1;
                j.lut = kMaxSel;
                j.acc = 0;
                j.out = pool;
                pool += 4;
                B.jobs.push_back(j);
                const int o = pool++;
                B.post.push_back(IpAdd{o, tmp[p].fresh_b[w], j.out});
                outv[(size_t)w] = o;
            }
            next.push_back(std::move(outv));
        }
        if (cur.size() & 1) next.push_back(cur.back());  // odd one out passes on untouched
        cur.swap(next);
        plan.layers.push_back(std::move(A));
        plan.layers.push_back(std::move(B));
    }
    for (int c = 0; c < 16; c++) plan.result[c] = cur[0][(size_t)c];
    plan.pool_size = pool;
    return plan;
}

// Run the plan on cleartext bits (tests of the circuit; nothing on the GPU path calls this).
inline uint16_t ip_eval_clear(const IpPlan &plan, const uint16_t *vals)
{
    std::vector<uint8_t> bit((size_t)plan.pool_size, 0);
    for (int v = 0; v < plan.nvals; v++)
        for (int w = 0; w < 16; w++) bit[(size_t)v * 16 + (15 - w)] = (vals[v] >> w) & 1;
    for (const IpLayer &L : plan.layers) {
        for (const IpAdd &a : L.pre) bit[(size_t)a.dst] = bit[(size_t)a.a] ^ bit[(size_t)a.b];
        for (const IpRefresh &r : L.refresh) bit[(size_t)r.dst] = bit[(size_t)L.cbs[(size_t)r.pos]];
        for (const IpJob &j : L.jobs) {
            unsigned idx = 0;
            for (int i = 0; i < 8; i++)
                if (j.sel[i] >= 0) idx |= (unsigned)bit[(size_t)L.cbs[(size_t)j.sel[i]]] << i;
            const unsigned val = ip_lut_value(j.lut, idx) >> (4 * j.acc);
            for (int q = 0; q < 4; q++) bit[(size_t)j.out + q] = (val >> q) & 1;
        }
        for (const IpAdd &a : L.post) bit[(size_t)a.dst] = bit[(size_t)a.a] ^ bit[(size_t)a.b];
    }
    uint16_t r = 0;
    for (int c = 0; c < 16; c++)
        if (plan.result[c] >= 0) r |= (uint16_t)(bit[(size_t)plan.result[c]] << c);
    return r;
}

}  // namespace cbs_host
