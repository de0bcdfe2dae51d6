// md5.cuh -- RFC 1321 MD5, one digest stream per lane, for sm_100a.
//
// Replaces hashlib.md5().update()/digest() at skyplane/obj_store/s3_interface.py:181-192.
// One chunk is ONE serial chain (the digest must equal hashlib.md5(whole chunk)), so a warp
// carries 32 chunks, lane = chunk.  The per-step dependent chain is 3 SASS ops:
//   LOP3 (F/G/H/I of the newest b) -> IADD3 (+ a + M[g] + K[i], pre-added off the chain)
//   -> LEA.HI (b + rotl(t, s): ptxas fuses the funnel shift and the add).
// Message words are staged into shared memory with 16-byte cp.async SKY_MD5_SLOTS-1 blocks ahead of the
// chain (ring of SKY_MD5_SLOTS x 64 B per lane, laid out [slot][piece][lane] so LDS.128 is conflict-free).
#pragma once
#include <stdint.h>

#include <type_traits>

#ifndef SKY_MD5_SLOTS
#define SKY_MD5_SLOTS 4
#endif

namespace sky {

struct Md5State {
    uint32_t a, b, c, d;
};

__device__ __forceinline__ void md5_init(Md5State &s) {
    s.a = 0x67452301u;
    s.b = 0xefcdab89u;
    s.c = 0x98badcfeu;
    s.d = 0x10325476u;
}

// M[g] + K[i] is formed by an opaque add so ptxas cannot re-associate the constant onto the
// dependent chain (it otherwise emits LOP3 -> IADD3 -> VIADD(+K) -> LEA.HI: four ops per step).
__device__ __forceinline__ uint32_t md5_mk(uint32_t m, uint32_t k) {
    uint32_t r;
    asm("add.u32 %0, %1, %2;" : "=r"(r) : "r"(m), "r"(k));
    return r;
}
#define SKY_MD5_STEP(FN, a, b, c, d, m, k, s)                 \
    {                                                         \
        uint32_t t_ = a + md5_mk((m), (k)) + FN(b, c, d);     \
        a = b + __funnelshift_l(t_, t_, s);                   \
    }
#define SKY_F(b, c, d) ((d) ^ ((b) & ((c) ^ (d))))
#define SKY_G(b, c, d) ((c) ^ ((d) & ((b) ^ (c))))
#define SKY_H(b, c, d) ((b) ^ (c) ^ (d))
#define SKY_I(b, c, d) ((c) ^ ((b) | ~(d)))

// One 64-byte block.  w[16] = little-endian message words.
__device__ __forceinline__ void md5_block(Md5State &st, const uint32_t (&w)[16]) {
    uint32_t a = st.a, b = st.b, c = st.c, d = st.d;
    SKY_MD5_STEP(SKY_F, a, b, c, d, w[0], 0xd76aa478, 7)
    SKY_MD5_STEP(SKY_F, d, a, b, c, w[1], 0xe8c7b756, 12)
    SKY_MD5_STEP(SKY_F, c, d, a, b, w[2], 0x242070db, 17)
    SKY_MD5_STEP(SKY_F, b, c, d, a, w[3], 0xc1bdceee, 22)
    SKY_MD5_STEP(SKY_F, a, b, c, d, w[4], 0xf57c0faf, 7)
    SKY_MD5_STEP(SKY_F, d, a, b, c, w[5], 0x4787c62a, 12)
    SKY_MD5_STEP(SKY_F, c, d, a, b, w[6], 0xa8304613, 17)
    SKY_MD5_STEP(SKY_F, b, c, d, a, w[7], 0xfd469501, 22)
    SKY_MD5_STEP(SKY_F, a, b, c, d, w[8], 0x698098d8, 7)
    SKY_MD5_STEP(SKY_F, d, a, b, c, w[9], 0x8b44f7af, 12)
    SKY_MD5_STEP(SKY_F, c, d, a, b, w[10], 0xffff5bb1, 17)
    SKY_MD5_STEP(SKY_F, b, c, d, a, w[11], 0x895cd7be, 22)
    SKY_MD5_STEP(SKY_F, a, b, c, d, w[12], 0x6b901122, 7)
    SKY_MD5_STEP(SKY_F, d, a, b, c, w[13], 0xfd987193, 12)
    SKY_MD5_STEP(SKY_F, c, d, a, b, w[14], 0xa679438e, 17)
    SKY_MD5_STEP(SKY_F, b, c, d, a, w[15], 0x49b40821, 22)

    SKY_MD5_STEP(SKY_G, a, b, c, d, w[1], 0xf61e2562, 5)
    SKY_MD5_STEP(SKY_G, d, a, b, c, w[6], 0xc040b340, 9)
    SKY_MD5_STEP(SKY_G, c, d, a, b, w[11], 0x265e5a51, 14)
    SKY_MD5_STEP(SKY_G, b, c, d, a, w[0], 0xe9b6c7aa, 20)
    SKY_MD5_STEP(SKY_G, a, b, c, d, w[5], 0xd62f105d, 5)
    SKY_MD5_STEP(SKY_G, d, a, b, c, w[10], 0x02441453, 9)
    SKY_MD5_STEP(SKY_G, c, d, a, b, w[15], 0xd8a1e681, 14)
    SKY_MD5_STEP(SKY_G, b, c, d, a, w[4], 0xe7d3fbc8, 20)
    SKY_MD5_STEP(SKY_G, a, b, c, d, w[9], 0x21e1cde6, 5)
    SKY_MD5_STEP(SKY_G, d, a, b, c, w[14], 0xc33707d6, 9)
    SKY_MD5_STEP(SKY_G, c, d, a, b, w[3], 0xf4d50d87, 14)
    SKY_MD5_STEP(SKY_G, b, c, d, a, w[8], 0x455a14ed, 20)
    SKY_MD5_STEP(SKY_G, a, b, c, d, w[13], 0xa9e3e905, 5)
    SKY_MD5_STEP(SKY_G, d, a, b, c, w[2], 0xfcefa3f8, 9)
    SKY_MD5_STEP(SKY_G, c, d, a, b, w[7], 0x676f02d9, 14)
    SKY_MD5_STEP(SKY_G, b, c, d, a, w[12], 0x8d2a4c8a, 20)

    SKY_MD5_STEP(SKY_H, a, b, c, d, w[5], 0xfffa3942, 4)
    SKY_MD5_STEP(SKY_H, d, a, b, c, w[8], 0x8771f681, 11)
    SKY_MD5_STEP(SKY_H, c, d, a, b, w[11], 0x6d9d6122, 16)
    SKY_MD5_STEP(SKY_H, b, c, d, a, w[14], 0xfde5380c, 23)
    SKY_MD5_STEP(SKY_H, a, b, c, d, w[1], 0xa4beea44, 4)
    SKY_MD5_STEP(SKY_H, d, a, b, c, w[4], 0x4bdecfa9, 11)
    SKY_MD5_STEP(SKY_H, c, d, a, b, w[7], 0xf6bb4b60, 16)
    SKY_MD5_STEP(SKY_H, b, c, d, a, w[10], 0xbebfbc70, 23)
    SKY_MD5_STEP(SKY_H, a, b, c, d, w[13], 0x289b7ec6, 4)
    SKY_MD5_STEP(SKY_H, d, a, b, c, w[0], 0xeaa127fa, 11)
    SKY_MD5_STEP(SKY_H, c, d, a, b, w[3], 0xd4ef3085, 16)
    SKY_MD5_STEP(SKY_H, b, c, d, a, w[6], 0x04881d05, 23)
    SKY_MD5_STEP(SKY_H, a, b, c, d, w[9], 0xd9d4d039, 4)
    SKY_MD5_STEP(SKY_H, d, a, b, c, w[12], 0xe6db99e5, 11)
    SKY_MD5_STEP(SKY_H, c, d, a, b, w[15], 0x1fa27cf8, 16)
    SKY_MD5_STEP(SKY_H, b, c, d, a, w[2], 0xc4ac5665, 23)

    SKY_MD5_STEP(SKY_I, a, b, c, d, w[0], 0xf4292244, 6)
    SKY_MD5_STEP(SKY_I, d, a, b, c, w[7], 0x432aff97, 10)
    SKY_MD5_STEP(SKY_I, c, d, a, b, w[14], 0xab9423a7, 15)
    SKY_MD5_STEP(SKY_I, b, c, d, a, w[5], 0xfc93a039, 21)
    SKY_MD5_STEP(SKY_I, a, b, c, d, w[12], 0x655b59c3, 6)
    SKY_MD5_STEP(SKY_I, d, a, b, c, w[3], 0x8f0ccc92, 10)
    SKY_MD5_STEP(SKY_I, c, d, a, b, w[10], 0xffeff47d, 15)
    SKY_MD5_STEP(SKY_I, b, c, d, a, w[1], 0x85845dd1, 21)
    SKY_MD5_STEP(SKY_I, a, b, c, d, w[8], 0x6fa87e4f, 6)
    SKY_MD5_STEP(SKY_I, d, a, b, c, w[15], 0xfe2ce6e0, 10)
    SKY_MD5_STEP(SKY_I, c, d, a, b, w[6], 0xa3014314, 15)
    SKY_MD5_STEP(SKY_I, b, c, d, a, w[13], 0x4e0811a1, 21)
    SKY_MD5_STEP(SKY_I, a, b, c, d, w[4], 0xf7537e82, 6)
    SKY_MD5_STEP(SKY_I, d, a, b, c, w[11], 0xbd3af235, 10)
    SKY_MD5_STEP(SKY_I, c, d, a, b, w[2], 0x2ad7d2bb, 15)
    SKY_MD5_STEP(SKY_I, b, c, d, a, w[9], 0xeb86d391, 21)
    st.a += a;
    st.b += b;
    st.c += c;
    st.d += d;
}

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Digest of one chunk per lane.  `ring` = this warp's 8 KiB shared-memory area (2048 x uint32),
// `src` 16-byte aligned (or len == 0), `active` false for lanes without a chunk.
// Writes 16 digest bytes to `out` for active lanes.
// `progress` (may be null): lane 0 publishes how many 64 KiB rows the warp has consumed, so LZ4 warps can
// fetch a row just ahead of the MD5 lanes and the lanes then hit L2 instead of HBM.
// `gate(row, wants)`: called (warp-converged) before the first byte of 64 KiB row `row` is fetched; `wants` tells
// whether this lane has data in that row.  The receiver side uses it to wait until the row has been decoded.
struct Md5NoGate {
    __device__ __forceinline__ void operator()(uint64_t, bool) const {}
};

template <class Gate = Md5NoGate>
__device__ __forceinline__ void md5_warp(uint32_t *ring, const uint8_t *src, uint64_t len, bool active, uint8_t *out,
                                         unsigned lane, volatile uint32_t *progress, Gate gate = Gate()) {
    constexpr int kSlots = SKY_MD5_SLOTS;  // ring depth (blocks, power of two); prefetch distance = kSlots - 1
    constexpr bool kGated = !std::is_same<Gate, Md5NoGate>::value;
    const uint64_t nfull = active ? (len >> 6) : 0;
    uint64_t wmax = nfull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint64_t other = __shfl_xor_sync(0xffffffffu, wmax, o);
        wmax = other > wmax ? other : wmax;
    }
    Md5State st;
    md5_init(st);
    const uint32_t ring_base = (uint32_t)__cvta_generic_to_shared(ring);
    // slot s, piece q of this lane lives at ring[((s*4 + q)*32 + lane) * 4 words]
    auto slot_addr = [&](int s, int q) { return ring_base + (uint32_t)(((s * 4 + q) * 32 + lane) * 16); };

    gate(0, active && len > 0);
#pragma unroll
    for (int s = 0; s < kSlots - 1; s++) {
        if ((uint64_t)s < nfull) {
#pragma unroll
            for (int q = 0; q < 4; q++) cp_async16(slot_addr(s, q), src + (uint64_t)s * 64 + q * 16);
        }
        cp_async_commit();
    }
    auto load_words = [&](uint32_t (&w)[16], int cs) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint4 v = *reinterpret_cast<const uint4 *>(ring + ((cs * 4 + q) * 32 + lane) * 4);
            w[4 * q + 0] = v.x;
            w[4 * q + 1] = v.y;
            w[4 * q + 2] = v.z;
            w[4 * q + 3] = v.w;
        }
    };
    // software pipeline: the words of block i+1 are pulled from the ring (LDS) while block i's chain runs
    uint32_t wn[16];
#pragma unroll
    for (int k = 0; k < 16; k++) wn[k] = 0;
    cp_async_wait<kSlots - 2>();  // block 0 has landed
    if (nfull) load_words(wn, 0);
#define SKY_MD5_LOOP_BODY(i)                                                                              \
    {                                                                                                     \
        uint32_t w[16];                                                                                   \
        _Pragma("unroll") for (int k = 0; k < 16; k++) w[k] = wn[k];                                      \
        const uint64_t pf = (i) + (kSlots - 1);                                                           \
        if (pf < nfull) {                                                                                 \
            const int ps = (int)(pf & (kSlots - 1)); /* == slot of block i-1, last read one iteration ago */ \
            _Pragma("unroll") for (int q = 0; q < 4; q++) cp_async16(slot_addr(ps, q), src + pf * 64 + q * 16); \
        }                                                                                                 \
        cp_async_commit();                                                                                \
        cp_async_wait<kSlots - 2>(); /* block i+1 has landed */                                           \
        if ((i) + 1 < nfull) load_words(wn, (int)(((i) + 1) & (kSlots - 1)));                             \
        if ((i) < nfull) md5_block(st, w);                                                                \
    }
    if constexpr (kGated) {
        // Receiver side: a row-granular outer loop keeps the gate (a spin with a scheduling barrier) out of the
        // chain-bound inner loop; the prefetch runs kSlots-1 blocks ahead, so a row is gated one row early.
        for (uint64_t base = 0; base < wmax; base += 1024) {
            gate((base >> 10) + 1, (base + 1024) * 64 < len);
            const uint64_t iend = (wmax - base > 1024) ? base + 1024 : wmax;
            for (uint64_t i = base; i < iend; i++) SKY_MD5_LOOP_BODY(i)
        }
    } else {
        // Sender side: one flat loop (measured 1.027x faster per block than the nested form); lane 0 publishes the
        // rows consumed so LZ4 warps can stay just ahead of the digest lanes.
        for (uint64_t i = 0; i < wmax; i++) {
            SKY_MD5_LOOP_BODY(i)
            if (progress && lane == 0 && ((i + 1) & 1023) == 0) *progress = (uint32_t)((i + 1) >> 10) + 1u;  // 1 + 64 KiB rows done
        }
    }
#undef SKY_MD5_LOOP_BODY
    cp_async_wait<0>();
    if (active) {
        // tail: rem bytes + 0x80 + zeros + u64le bit length -> one or two more blocks (slow path, once per chunk)
        const uint32_t rem = (uint32_t)(len & 63);
        const uint8_t *tp = src + (nfull << 6);
        uint32_t tw[32];
#pragma unroll
        for (int k = 0; k < 32; k++) tw[k] = 0;
        for (uint32_t k = 0; k < rem; k++) tw[k >> 2] |= (uint32_t)tp[k] << (8 * (k & 3));
        tw[rem >> 2] |= 0x80u << (8 * (rem & 3));
        const bool two = rem >= 56;
        const uint64_t bits = len << 3;
        tw[two ? 30 : 14] = (uint32_t)bits;
        tw[two ? 31 : 15] = (uint32_t)(bits >> 32);
        for (int t = 0; t < (two ? 2 : 1); t++) {
            uint32_t w[16];
#pragma unroll
            for (int k = 0; k < 16; k++) w[k] = tw[16 * t + k];
            md5_block(st, w);
        }
        uint4 dg = make_uint4(st.a, st.b, st.c, st.d);
        *reinterpret_cast<uint4 *>(out) = dg;  // out is 16-byte aligned (md5 array base is 256-aligned)
    }
}

}  // namespace sky
