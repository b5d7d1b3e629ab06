// host_keys.cpp -- upstream's dictionary keys, computed in bulk on the host.
//
// Upstream keys its `_nodes` / `_edges` dicts by int(sha256(pickle.dumps(x)).hexdigest(), 16)
// (amira/construct_gene.py:5-10) where x is a tuple of signed 256-bit integers: the canonical genes'
// hashes for a node (construct_gene_mer.py:94-97), (source*sd, target*td) for an edge
// (construct_edge.py:104-124).  Downstream code depends on the numeric values, so the drop-in class
// must reproduce them; doing ~10^5 of them through hashlib + pickle costs a third of the Python
// materialisation.  This file restates the two ingredients -- the pickle protocol-4 byte stream of a
// tuple of ints and SHA-256 (FIPS 180-4) -- and the Python side checks the first results of every
// call against hashlib/pickle before trusting it (amira_b200/_keys.py).
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#if defined(__x86_64__) && defined(__GNUC__)
#define AMIRA_SHA_NI 1
#include <cpuid.h>
#include <immintrin.h>
#endif

#include "../../include/amira_gmg.h"

namespace {

static const uint32_t SHA_K[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};

#ifdef AMIRA_SHA_NI
// one SHA-256 compression with the x86 SHA extensions (the host side hashes two to three short messages per node
// and edge: ~4x faster than the scalar rounds); used when CPUID reports them, checked against hashlib by the caller
// like everything else in this file
__attribute__((target("sha,sse4.1,ssse3"))) static void sha256_block_ni(uint32_t state[8], const uint8_t *data) {
    const __m128i bswap = _mm_set_epi64x(0x0c0d0e0f08090a0bLL, 0x0405060700010203LL);
    __m128i tmp = _mm_loadu_si128((const __m128i *)&state[0]);     // DCBA
    __m128i st1 = _mm_loadu_si128((const __m128i *)&state[4]);     // HGFE
    tmp = _mm_shuffle_epi32(tmp, 0xB1);                            // CDAB
    st1 = _mm_shuffle_epi32(st1, 0x1B);                            // EFGH
    __m128i st0 = _mm_alignr_epi8(tmp, st1, 8);                    // ABEF
    st1 = _mm_blend_epi16(st1, tmp, 0xF0);                         // CDGH
    const __m128i save0 = st0, save1 = st1;
    __m128i M[4];
    for (int i = 0; i < 16; ++i) {
        if (i < 4) {
            M[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(data + 16 * i)), bswap);
        } else {
            __m128i t = _mm_sha256msg1_epu32(M[i & 3], M[(i + 1) & 3]);
            t = _mm_add_epi32(t, _mm_alignr_epi8(M[(i + 3) & 3], M[(i + 2) & 3], 4));
            M[i & 3] = _mm_sha256msg2_epu32(t, M[(i + 3) & 3]);
        }
        __m128i msg = _mm_add_epi32(M[i & 3], _mm_loadu_si128((const __m128i *)&SHA_K[4 * i]));
        st1 = _mm_sha256rnds2_epu32(st1, st0, msg);
        msg = _mm_shuffle_epi32(msg, 0x0E);
        st0 = _mm_sha256rnds2_epu32(st0, st1, msg);
    }
    st0 = _mm_add_epi32(st0, save0);
    st1 = _mm_add_epi32(st1, save1);
    tmp = _mm_shuffle_epi32(st0, 0x1B);                            // FEBA
    st1 = _mm_shuffle_epi32(st1, 0xB1);                            // DCHG
    st0 = _mm_blend_epi16(tmp, st1, 0xF0);                         // DCBA
    st1 = _mm_alignr_epi8(st1, tmp, 8);                            // HGFE
    _mm_storeu_si128((__m128i *)&state[0], st0);
    _mm_storeu_si128((__m128i *)&state[4], st1);
}

static bool cpu_has_sha_ni() {
    unsigned int a = 0, b = 0, c = 0, d = 0;
    if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return false;
    const bool sha = (b >> 29) & 1u;
    if (!__get_cpuid(1, &a, &b, &c, &d)) return false;
    return sha && ((c >> 19) & 1u) && ((c >> 9) & 1u);             // + SSE4.1, SSSE3
}
static const bool g_sha_ni = cpu_has_sha_ni();
#endif

struct Sha256 {
    uint32_t h[8];
    uint8_t buf[64];
    uint64_t len = 0;
    size_t fill = 0;
    Sha256() {
        static const uint32_t init[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au,
                                         0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
        memcpy(h, init, sizeof(h));
    }
    static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void block(const uint8_t *p) {
#ifdef AMIRA_SHA_NI
        if (g_sha_ni) {
            sha256_block_ni(h, p);
            return;
        }
#endif
        static const uint32_t K[64] = {
            0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
            0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
            0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
            0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
            0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
            0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
            0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
            0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
        uint32_t w[64];
        for (int i = 0; i < 16; ++i)
            w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
        for (int i = 16; i < 64; ++i) {
            const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            const uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; ++i) {
            const uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
            const uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    void update(const uint8_t *p, size_t n) {
        len += n;
        while (n) {
            const size_t take = n < 64 - fill ? n : 64 - fill;
            memcpy(buf + fill, p, take);
            fill += take; p += take; n -= take;
            if (fill == 64) {
                block(buf);
                fill = 0;
            }
        }
    }
    void finish(uint8_t out[32]) {
        const uint64_t bits = len * 8;
        const uint8_t one = 0x80, zero = 0;
        update(&one, 1);
        while (fill != 56) update(&zero, 1);
        uint8_t l[8];
        for (int i = 0; i < 8; ++i) l[i] = (uint8_t)(bits >> (56 - 8 * i));
        update(l, 8);
        for (int i = 0; i < 8; ++i) {
            out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16);
            out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i];
        }
    }
};

// pickle protocol 4 opcode stream of one Python int = sign * (32-byte big-endian magnitude)
void pickle_int(std::vector<uint8_t> &o, const uint8_t mag_be[32], int negative) {
    uint8_t le[33];
    int n = 0;  // magnitude bytes, little endian, no leading zeros
    for (int i = 31; i >= 0; --i) le[31 - i] = mag_be[i];
    n = 32;
    while (n > 0 && le[n - 1] == 0) --n;
    if (n == 0) negative = 0;
    if (n <= 4) {  // may fit the fixed-width opcodes
        uint64_t v = 0;
        for (int i = 0; i < n; ++i) v |= (uint64_t)le[i] << (8 * i);
        if (!negative && v < 256) { o.push_back('K'); o.push_back((uint8_t)v); return; }
        if (!negative && v < 65536) { o.push_back('M'); o.push_back((uint8_t)v); o.push_back((uint8_t)(v >> 8)); return; }
        if ((!negative && v <= 0x7FFFFFFFull) || (negative && v <= 0x80000000ull)) {
            const uint32_t s = negative ? (uint32_t)(0u - (uint32_t)v) : (uint32_t)v;
            o.push_back('J');
            for (int i = 0; i < 4; ++i) o.push_back((uint8_t)(s >> (8 * i)));
            return;
        }
    }
    // LONG1: little-endian two's complement in (bit_length >> 3) + 1 bytes, redundant sign byte trimmed
    int bit_length = (n - 1) * 8;
    for (uint8_t t = le[n - 1]; t; t >>= 1) ++bit_length;
    int nbytes = (bit_length >> 3) + 1;
    uint8_t tc[34];
    memset(tc, 0, sizeof(tc));
    memcpy(tc, le, n);
    if (negative) {
        int carry = 1;
        for (int i = 0; i < nbytes; ++i) {
            const int v = (uint8_t)~tc[i] + carry;
            tc[i] = (uint8_t)v;
            carry = v >> 8;
        }
        if (nbytes > 1 && tc[nbytes - 1] == 0xff && (tc[nbytes - 2] & 0x80)) --nbytes;
    }
    o.push_back(0x8a);
    o.push_back((uint8_t)nbytes);
    o.insert(o.end(), tc, tc + nbytes);
}

void sha_of_int_tuple(const uint8_t *mags, const int8_t *neg, int arity, std::vector<uint8_t> &payload, uint8_t out[32]) {
    payload.clear();
    if (arity > 3) payload.push_back('(');
    for (int i = 0; i < arity; ++i) pickle_int(payload, mags + 32 * i, neg[i]);
    payload.push_back(arity == 0 ? ')' : arity == 1 ? 0x85 : arity == 2 ? 0x86 : arity == 3 ? 0x87 : 't');
    if (arity > 0) payload.push_back(0x94);  // MEMOIZE (the empty tuple is a singleton and is not memoised)
    payload.push_back('.');
    uint8_t head[11] = {0x80, 0x04, 0x95};
    const uint64_t n = payload.size();
    for (int i = 0; i < 8; ++i) head[3 + i] = (uint8_t)(n >> (8 * i));
    Sha256 s;
    s.update(head, n >= 4 ? 11 : 2);  // frames shorter than 4 bytes are not framed
    s.update(payload.data(), payload.size());
    s.finish(out);
}

// [0, n) in contiguous pieces on the host cores (the keys are independent)
template <typename F>
void parallel_ranges(int64_t n, F f) {
    unsigned int hw = std::thread::hardware_concurrency();
    // (threads pay from ~30 000 keys each on: cores asleep take longer to wake than a small piece takes to hash)
    int64_t n_thr = std::min<int64_t>(hw ? hw : 1, std::min<int64_t>(16, n / 32768));
    if (n_thr <= 1) {
        f(0, n);
        return;
    }
    std::vector<std::thread> pool;
    const int64_t per = (n + n_thr - 1) / n_thr;
    for (int64_t t = 0; t < n_thr; ++t) {
        const int64_t lo = t * per, hi = std::min(n, lo + per);
        if (lo < hi) pool.emplace_back([=] { f(lo, hi); });
    }
    for (auto &th : pool) th.join();
}

}  // namespace

extern "C" {

// out[t] = sha256(pickle.dumps(tuple of `arity` ints)) for n_tuples tuples; item i of tuple t is
// (neg ? -1 : 1) * big-endian magnitude at mags[(t*arity + i)*32]
int amira_host_tuple_sha(const uint8_t *mags, const int8_t *neg, int64_t n_tuples, int32_t arity, uint8_t *out) {
    if (n_tuples < 0 || arity < 0 || (n_tuples > 0 && arity > 0 && (!mags || !neg)) || (n_tuples > 0 && !out)) return AMIRA_E_ARG;
    parallel_ranges(n_tuples, [=](int64_t lo, int64_t hi) {
        std::vector<uint8_t> payload;
        payload.reserve(64 + 40 * (size_t)arity);
        for (int64_t t = lo; t < hi; ++t)
            sha_of_int_tuple(mags + (size_t)t * arity * 32, neg + (size_t)t * arity, arity, payload, out + 32 * t);
    });
    return AMIRA_OK;
}

// Edge keys (construct_edge.py:104-124): min(H((s*sd, t*td)), H((-s*sd, -t*td))) with s, t the node keys
int amira_host_edge_keys(const uint8_t *node_sha, const int32_t *src, const int32_t *tgt, const int8_t *sd,
                         const int8_t *td, int64_t n_edges, uint8_t *out) {
    if (n_edges < 0 || (n_edges > 0 && (!node_sha || !src || !tgt || !sd || !td || !out))) return AMIRA_E_ARG;
    parallel_ranges(n_edges, [=](int64_t lo, int64_t hi) {
        std::vector<uint8_t> payload;
        payload.reserve(128);
        uint8_t mags[64], a[32], b[32];
        int8_t neg[2];
        for (int64_t e = lo; e < hi; ++e) {
            memcpy(mags, node_sha + 32 * (size_t)src[e], 32);
            memcpy(mags + 32, node_sha + 32 * (size_t)tgt[e], 32);
            neg[0] = sd[e] < 0; neg[1] = td[e] < 0;
            sha_of_int_tuple(mags, neg, 2, payload, a);
            neg[0] = !neg[0]; neg[1] = !neg[1];
            sha_of_int_tuple(mags, neg, 2, payload, b);
            memcpy(out + 32 * e, memcmp(a, b, 32) <= 0 ? a : b, 32);
        }
    });
    return AMIRA_OK;
}

}  // extern "C"
