// lz4.cuh -- warp-per-block LZ4 block compressor for sm_100a.
//
// Replaces, per 64 KiB block, what lz4.frame.compress(data) does inside
// skyplane/gateway/operators/gateway_operator.py:358-361 (liblz4's level-0 "fast" compressor:
// single-candidate hash table, greedy, skip acceleration).  The GPU formulation:
//   * one warp owns one independent 64 KiB block (frame flag B.Indep, so no cross-block state);
//   * the match table is 4096 x u16 block positions in shared memory (8 KiB per warp);
//   * each iteration the 32 lanes probe 32 cursor positions (stride = LZ4's skip step, which grows
//     by one every 64 failed probes), the lowest matching lane wins (greedy = reference order),
//     the match is extended backwards/forwards cooperatively, and the sequence is emitted.
//   * blocks that do not shrink are stored raw (bit 31 of the block header), like LZ4F_makeBlock.
// Output is a standard LZ4 block: decodable by lz4.frame.decompress (gateway_receiver.py:196).
#pragma once
#include <stdint.h>

namespace sky {

constexpr uint32_t kBlock = 65536;       // BD = 0x40
constexpr uint32_t kSlot = kBlock + 4;   // worst-case block footprint in the frame (header + raw data)
#ifndef SKY_HASHLOG
#define SKY_HASHLOG 12
#endif
constexpr uint32_t kHashLog = SKY_HASHLOG;
constexpr uint32_t kHashSize = 1u << kHashLog;
constexpr uint32_t kMinMatch = 4;
constexpr uint32_t kMfLimit = 12;        // a match must start >= 12 bytes before the block end
constexpr uint32_t kLastLiterals = 5;    // the last 5 bytes are always literals
constexpr uint32_t kSkipTrigger = 6;
constexpr unsigned kFull = 0xffffffffu;

// ---- unaligned little-endian 32-bit read from global memory (base 4-byte aligned) --------------
__device__ __forceinline__ uint32_t load32(const uint8_t *base, uint32_t pos) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(base + (pos & ~3u));
    const uint32_t lo = __ldg(w), hi = __ldg(w + 1);
    return __funnelshift_r(lo, hi, (pos & 3u) * 8u);
}

__device__ __forceinline__ uint32_t lz4_hash(uint32_t v) { return (v * 2654435761u) >> (32 - kHashLog); }

// ---- warp copy: dst and src arbitrarily aligned; regions disjoint, or dst < src (forward move) ----
// Over-reads at most 3 bytes past src+n (inside the same 4-byte word group); never over-writes.
__device__ __forceinline__ void warp_copy(uint8_t *dst, const uint8_t *src, uint32_t n, unsigned lane) {
    if (n < 64) {
        for (uint32_t base = 0; base < n; base += 32) {
            const uint32_t k = base + lane;
            uint8_t b = 0;
            if (k < n) b = src[k];
            __syncwarp();
            if (k < n) dst[k] = b;
        }
        return;
    }
    const uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;  // < 16 <= n
    uint8_t hb = 0;
    if (lane < head) hb = src[lane];
    __syncwarp();
    if (lane < head) dst[lane] = hb;
    dst += head;
    src += head;
    n -= head;
    const uint32_t nvec = n >> 4;
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u) * 8u;
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3);
    uint4 *dv = reinterpret_cast<uint4 *>(dst);
    for (uint32_t base = 0; base < nvec; base += 32) {
        const uint32_t k = base + lane;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (k < nvec) {
            const uint32_t *q = sw + 4 * (size_t)k;
            const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3];
            const uint32_t w4 = sh ? q[4] : 0u;
            o.x = __funnelshift_r(w0, w1, sh);
            o.y = __funnelshift_r(w1, w2, sh);
            o.z = __funnelshift_r(w2, w3, sh);
            o.w = __funnelshift_r(w3, w4, sh);
        }
        __syncwarp();  // every lane has its source words before any lane overwrites (forward move)
        if (k < nvec) dv[k] = o;
    }
    const uint32_t done = nvec << 4, tail = n & 15u;
    uint8_t tb = 0;
    if (lane < tail) tb = src[done + lane];
    __syncwarp();
    if (lane < tail) dst[done + lane] = tb;
}

// ---- warp copy for DISJOINT regions whose source is kernel-read-only input (ld.global.nc), 4 x 16 B in
// flight per lane.  src 16-byte aligned; dst arbitrary.
__device__ __forceinline__ void warp_copy_input(uint8_t *dst, const uint8_t *__restrict__ src, uint32_t n, unsigned lane) {
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u);
    if (mis == 0) {
        const uint4 *sv = reinterpret_cast<const uint4 *>(src);
        uint4 *dv = reinterpret_cast<uint4 *>(dst);
        const uint32_t nvec = n >> 4;
        uint32_t k = lane;
        for (; k + 96 < nvec; k += 128) {
            const uint4 a = __ldg(sv + k), b = __ldg(sv + k + 32), c = __ldg(sv + k + 64), d = __ldg(sv + k + 96);
            __stcs(dv + k, a); __stcs(dv + k + 32, b); __stcs(dv + k + 64, c); __stcs(dv + k + 96, d);
        }
        for (; k < nvec; k += 32) __stcs(dv + k, __ldg(sv + k));
        const uint32_t done = nvec << 4;
        if (lane < (n & 15u)) dst[done + lane] = src[done + lane];
        return;
    }
    // dst = 16-byte aligned base + mis: build each aligned 16-byte store from two consecutive source vectors
    const uint32_t head = 16u - mis;  // bytes until dst is aligned
    if (n <= head + 16) {
        for (uint32_t k = lane; k < n; k += 32) dst[k] = src[k];
        return;
    }
    if (lane < head) dst[lane] = src[lane];
    uint4 *dv = reinterpret_cast<uint4 *>(dst + head);
    const uint32_t rem = n - head, nvec = rem >> 4;
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(src);  // src + head = word (head>>2), byte shift (head&3)
    const uint32_t wsh = head >> 2, bsh = (head & 3u) * 8u;
    for (uint32_t k = lane; k < nvec; k += 32) {
        const uint32_t *q = sw + wsh + 4 * (size_t)k;
        const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = __ldg(q + 3);
        const uint32_t w4 = bsh ? __ldg(q + 4) : 0u;
        uint4 o;
        o.x = __funnelshift_r(w0, w1, bsh);
        o.y = __funnelshift_r(w1, w2, bsh);
        o.z = __funnelshift_r(w2, w3, bsh);
        o.w = __funnelshift_r(w3, w4, bsh);
        __stcs(dv + k, o);  // streaming: the frame is never re-read here, keep L2 for the input rows
    }
    const uint32_t done = head + (nvec << 4);
    if (lane < (rem & 15u)) dst[done + lane] = src[done + lane];
}

// Ask L2 to fetch [p, p+bytes) (bytes multiple of 16, p 16-byte aligned); one thread issues it.
// -DSKY_L2_EVICT_LAST=1 tags the row evict-last so it survives until the MD5 lanes have read it (experiment for the
// 1.21x DRAM traffic of config 2; off by default, not yet measured).
__device__ __forceinline__ void l2_prefetch_bulk(const void *p, uint32_t bytes) {
#if SKY_L2_EVICT_LAST
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(p), "r"(bytes), "l"(policy) : "memory");
#else
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#endif
}

// ---- XXH32 of the frame descriptor (2 or 10 bytes), for the header checksum byte -----------------
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int s) { return __funnelshift_l(x, x, s); }
__device__ __forceinline__ uint32_t xxh32_small(const uint8_t *p, uint32_t len) {  // len < 16, seed 0
    constexpr uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    uint32_t h = P5 + len;
    uint32_t i = 0;
    for (; i + 4 <= len; i += 4) {
        const uint32_t v = p[i] | (p[i + 1] << 8) | (p[i + 2] << 16) | ((uint32_t)p[i + 3] << 24);
        h = rotl32(h + v * P3, 17) * P4;
    }
    for (; i < len; i++) h = rotl32(h + p[i] * P5, 11) * P1;
    h ^= h >> 15;
    h *= P2;
    h ^= h >> 13;
    h *= P3;
    h ^= h >> 16;
    return h;
}

// Writes the frame header for an n-byte chunk at dst; returns its size (15, or 7 when n == 0).
// Single thread.
__device__ __forceinline__ uint32_t write_frame_header(uint8_t *dst, uint64_t n) {
    uint8_t d[10];
    d[1] = 0x40;  // BD: 64 KiB blocks
    dst[0] = 0x04; dst[1] = 0x22; dst[2] = 0x4D; dst[3] = 0x18;
    if (n == 0) {
        d[0] = 0x60;  // v01 | B.Indep ; content size omitted (0 means "unknown" to LZ4F)
        dst[4] = d[0]; dst[5] = d[1];
        dst[6] = (uint8_t)(xxh32_small(d, 2) >> 8);
        return 7;
    }
    d[0] = 0x68;  // v01 | B.Indep | C.Size
#pragma unroll
    for (int i = 0; i < 8; i++) d[2 + i] = (uint8_t)(n >> (8 * i));
#pragma unroll
    for (int i = 0; i < 10; i++) dst[4 + i] = d[i];
    dst[14] = (uint8_t)(xxh32_small(d, 10) >> 8);
    return 15;
}

// ---- sequence emission ----------------------------------------------------------------------------
// Bytes a sequence with `ll` literals and match length `ml` (>= 4, or 0 for the final literal run) needs.
__device__ __forceinline__ uint32_t seq_bytes(uint32_t ll, uint32_t ml) {
    uint32_t s = 1 + ll + (ll >= 15 ? (ll - 15) / 255 + 1 : 0);
    if (ml) s += 2 + ((ml - 4) >= 15 ? (ml - 4 - 15) / 255 + 1 : 0);
    return s;
}

// Emits token, literal-length bytes, literals, offset, match-length bytes.  All lanes call it with
// warp-uniform arguments; returns the new output cursor.
__device__ __forceinline__ uint32_t emit_seq(uint8_t *out, uint32_t op, const uint8_t *src, uint32_t anchor, uint32_t ll,
                                            uint32_t ml, uint32_t offset, unsigned lane) {
    const uint32_t mcode = ml ? ml - kMinMatch : 0;
    if (lane == 0) out[op] = (uint8_t)(((ll < 15 ? ll : 15) << 4) | (mcode < 15 ? mcode : 15));
    op += 1;
    if (ll >= 15) {
        const uint32_t r = ll - 15, n255 = r / 255;
        for (uint32_t k = lane; k < n255; k += 32) out[op + k] = 255;
        if (lane == 0) out[op + n255] = (uint8_t)(r - n255 * 255);
        op += n255 + 1;
    }
    warp_copy(out + op, src + anchor, ll, lane);
    op += ll;
    if (ml) {
        if (lane == 0) {
            out[op] = (uint8_t)offset;
            out[op + 1] = (uint8_t)(offset >> 8);
        }
        op += 2;
        if (mcode >= 15) {
            const uint32_t r = mcode - 15, n255 = r / 255;
            for (uint32_t k = lane; k < n255; k += 32) out[op + k] = 255;
            if (lane == 0) out[op + n255] = (uint8_t)(r - n255 * 255);
            op += n255 + 1;
        }
    }
    return op;
}

// ---- cooperative forward extension: all lanes compare 4 bytes each per round, starting at match length `mlen`
__device__ __forceinline__ uint32_t extend_coop(const uint8_t *__restrict__ src, uint32_t mpos, uint32_t mcand, uint32_t mlen,
                                                uint32_t maxlen, unsigned lane) {
    for (;;) {
        const uint32_t o = mlen + lane * 4;
        uint32_t cnt = 0;
        if (o < maxlen) {
            const uint32_t x = load32(src, mpos + o) ^ load32(src, mcand + o);
            cnt = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
            cnt = min(cnt, maxlen - o);
        }
        const unsigned fullm = __ballot_sync(kFull, cnt == 4);
        if (fullm == kFull) {
            mlen += 128;
            continue;
        }
        const int f = __ffs(~fullm) - 1;
        return mlen + 4 * f + __shfl_sync(kFull, cnt, f);
    }
}

#ifndef SKY_BACK_EXT_ALWAYS
#define SKY_BACK_EXT_ALWAYS 0
#endif
#ifndef SKY_L2_EVICT_LAST
#define SKY_L2_EVICT_LAST 0
#endif
#ifndef SKY_EXT_ROUNDS
#define SKY_EXT_ROUNDS 7
#endif
#ifndef SKY_COOP_LIT
#define SKY_COOP_LIT 48
#endif
constexpr int kExtRounds = SKY_EXT_ROUNDS;     // per-lane extension: up to 4 + 4*rounds bytes before going cooperative
constexpr uint32_t kCoopLit = SKY_COOP_LIT;    // literal runs at least this long are copied by the whole warp

// ---- the block compressor -------------------------------------------------------------------------
// src: block start in the chunk (16-byte aligned), L: block length (1..65536), out: where compressed
// bytes may be written (capacity L bytes), ht: this warp's match table.
// Returns the compressed size (1..L-1), or 0 if the block does not shrink (caller stores it raw).
//
// One iteration handles a whole window of 32 cursor positions: every lane probes its position, lanes with a
// verified candidate extend their own match (lane-parallel), then matches are accepted greedily in position
// order (a match is skipped if it starts inside an accepted one -- exactly what the sequential reference
// does), and all accepted sequences are sized with a warp scan and written by their own lanes at once.
__device__ __forceinline__ uint32_t lz4_compress_block(const uint8_t *__restrict__ src, uint32_t L, uint8_t *__restrict__ out,
                                                      uint16_t *ht, unsigned lane) {
    {   // clear the table: 32 lanes x 16 B per round
        uint4 *t4 = reinterpret_cast<uint4 *>(ht);
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < (int)(kHashSize * 2 / 16 / 32); k++) t4[k * 32 + lane] = z;
    }
    __syncwarp();

    uint32_t ip = 0, anchor = 0, op = 0;
    const uint32_t limit = L - 1;  // accept only csize <= L-1 (LZ4F_makeBlock passes dstCapacity = srcSize-1)
    const unsigned lt_mask = (1u << lane) - 1u;

    if (L >= kMfLimit + 1) {
        const uint32_t mflimit = L - kMfLimit;          // last position a match may start at
        const uint32_t matchlimit = L - kLastLiterals;  // matches end at or before this
        uint32_t nprobe = 1u << kSkipTrigger;           // LZ4: searchMatchNb = acceleration << skipTrigger
        while (ip <= mflimit) {
            const uint32_t step = nprobe >> kSkipTrigger;
            uint32_t pos = ip + lane * step;
            const bool valid = pos <= mflimit;
            uint32_t v = 0, h = 0, cand = 0;
            if (valid) {
                v = load32(src, pos);
                h = lz4_hash(v);
                cand = ht[h];
            }
            bool hit = valid && cand < pos && load32(src, cand) == v;
            // Lanes that hash to the same slot are grouped with one match.any: (a) the nearest lower lane of the
            // group is the most recent occurrence inside the window (offsets below 32*step: runs, short periods,
            // repeated words), which the table cannot know yet -- like the sequential reference, the most recent
            // occurrence wins; (b) only the highest lane of a group stores its position, so the table update is
            // deterministic and free of same-instruction write-write conflicts.
            const unsigned vmask = __ballot_sync(kFull, valid);
            const unsigned grp = __match_any_sync(kFull, valid ? h : (0xffff0000u | lane)) & vmask;
            {
                const unsigned lower = grp & lt_mask;
                const int near = lower ? 31 - __clz(lower) : 0;
                const uint32_t v_near = __shfl_sync(kFull, v, near);
                if (valid && lower && v_near == v) {
                    cand = ip + (uint32_t)near * step;
                    hit = true;
                }
            }
            const unsigned hits = __ballot_sync(kFull, hit);  // (also orders the table reads before the writes)
            // every probed position enters the table: the cursor always moves past the whole window, so no
            // entry can point ahead of a later probe
            if (valid && (grp >> lane) == 1u) ht[h] = (uint16_t)pos;
            __syncwarp();
            if (hits == 0) {
                ip += 32 * step;
                nprobe += 32;
                continue;
            }
            // ---- lane-parallel forward extension (bounded); `more` = still matching at the bound
            uint32_t mlen = 0;
            bool more = false;
            if (hit) {
                const uint32_t maxlen = matchlimit - pos;  // >= 7
                mlen = kMinMatch;
                more = true;
#pragma unroll 1
                for (int r = 0; r < kExtRounds && mlen < maxlen; r++) {
                    const uint32_t x = load32(src, pos + mlen) ^ load32(src, cand + mlen);
                    if (x) {
                        mlen += (uint32_t)(__ffs(x) - 1) >> 3;
                        more = false;
                        break;
                    }
                    mlen += 4;
                }
                if (mlen >= maxlen) {
                    mlen = maxlen;
                    more = false;
                }
            }
            __syncwarp();
            // ---- greedy acceptance in position order
            unsigned sel = 0, rem = hits;
            while (rem) {
                const int l = __ffs(rem) - 1;
                const uint32_t p_l = ip + (uint32_t)l * step;
                uint32_t len_l = __shfl_sync(kFull, mlen, l);
                if (__shfl_sync(kFull, (int)more, l)) {  // long match: finish it with the whole warp
                    len_l = extend_coop(src, p_l, __shfl_sync(kFull, cand, l), len_l, matchlimit - p_l, lane);
                    if ((int)lane == l) mlen = len_l;
                }
                sel |= 1u << l;
                rem &= __ballot_sync(kFull, pos >= p_l + len_l);  // drop every hit that starts inside this match
            }
            const bool is_sel = (sel >> lane) & 1u;
            const unsigned before = sel & lt_mask;
            const int prev_l = before ? 31 - __clz(before) : -1;
            const uint32_t e_mine = pos + mlen;
            uint32_t prev_end = __shfl_sync(kFull, e_mine, prev_l < 0 ? 0 : prev_l);
            if (prev_l < 0) prev_end = anchor;
            // ---- backward extension ("catch up").  Default: only probes that skipped positions (step > 1).
            // -DSKY_BACK_EXT_ALWAYS=1 extends every accepted match: +1.5 % ratio on the Silesia-like set in the CPU
            // model (tools/ratio_study.py) for one more dependent load per window -- to be measured on the GPU.
            if ((SKY_BACK_EXT_ALWAYS || step > 1) && is_sel) {
                const uint32_t room = min(pos - prev_end, cand);
                uint32_t b = 0;
                while (b < room && src[pos - 1 - b] == src[cand - 1 - b]) b++;
                pos -= b;
                cand -= b;
                mlen += b;
            }
            const uint32_t ll = is_sel ? pos - prev_end : 0;
            const uint32_t sz = is_sel ? seq_bytes(ll, mlen) : 0;
            uint32_t incl = sz;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(kFull, incl, d);
                if ((int)lane >= d) incl += t;
            }
            const uint32_t total = __shfl_sync(kFull, incl, 31);
            if (op + total + 1 + kLastLiterals > limit) return 0;  // cannot end up smaller than the input
            uint32_t lit_o = 0;
            if (is_sel) {
                uint32_t o = op + incl - sz;
                const uint32_t mcode = mlen - kMinMatch;
                out[o++] = (uint8_t)(((ll < 15 ? ll : 15) << 4) | (mcode < 15 ? mcode : 15));
                if (ll >= 15) {
                    uint32_t r = ll - 15;
                    for (; r >= 255; r -= 255) out[o++] = 255;
                    out[o++] = (uint8_t)r;
                }
                lit_o = o;
                if (ll < kCoopLit) {
                    uint32_t k = 0;
                    for (; k + 4 <= ll; k += 4) {  // 4 literal bytes per round: one unaligned read, four byte stores
                        const uint32_t w = load32(src, prev_end + k);
                        out[o + k] = (uint8_t)w;
                        out[o + k + 1] = (uint8_t)(w >> 8);
                        out[o + k + 2] = (uint8_t)(w >> 16);
                        out[o + k + 3] = (uint8_t)(w >> 24);
                    }
                    for (; k < ll; k++) out[o + k] = src[prev_end + k];
                }
                o += ll;
                const uint32_t offset = pos - cand;
                out[o] = (uint8_t)offset;
                out[o + 1] = (uint8_t)(offset >> 8);
                o += 2;
                if (mcode >= 15) {
                    uint32_t r = mcode - 15;
                    for (; r >= 255; r -= 255) out[o++] = 255;
                    out[o++] = (uint8_t)r;
                }
            }
            __syncwarp();
            unsigned big = __ballot_sync(kFull, is_sel && ll >= kCoopLit);
            while (big) {  // long literal runs: one warp-wide copy each
                const int l = __ffs(big) - 1;
                big &= big - 1;
                warp_copy(out + __shfl_sync(kFull, lit_o, l), src + __shfl_sync(kFull, prev_end, l), __shfl_sync(kFull, ll, l), lane);
            }
            op += total;
            anchor = __shfl_sync(kFull, e_mine, 31 - __clz(sel));
            ip = max(ip + 32 * step, anchor);
            nprobe = 1u << kSkipTrigger;
        }
    }
    // last literals
    const uint32_t last = L - anchor;
    if (op + seq_bytes(last, 0) > limit) return 0;
    op = emit_seq(out, op, src, anchor, last, 0, 0, lane);
    return op;
}

}  // namespace sky
