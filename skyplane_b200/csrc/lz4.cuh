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
#ifndef SKY_LZ4_ENTRIES
#define SKY_LZ4_ENTRIES 4096
#endif
constexpr uint32_t kEntries = SKY_LZ4_ENTRIES;  // match-table entries per warp (u32 each: pos16 | tag16); any multiple of 128
constexpr uint32_t kTableBytes = kEntries * 4;
constexpr uint32_t kMinMatch = 4;
constexpr uint32_t kMfLimit = 12;        // a match must start >= 12 bytes before the block end
constexpr uint32_t kLastLiterals = 5;    // the last 5 bytes are always literals
constexpr unsigned kFull = 0xffffffffu;

// ---- unaligned little-endian 32-bit read from global memory (base 4-byte aligned) --------------
__device__ __forceinline__ uint32_t load32(const uint8_t *base, uint32_t pos) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(base + (pos & ~3u));
    const uint32_t lo = __ldg(w), hi = __ldg(w + 1);
    return __funnelshift_r(lo, hi, (pos & 3u) * 8u);
}

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

// ---- streaming warp copy for DISJOINT regions: 16-byte loads (4 in flight per lane), 16-byte streaming stores (the
// frame is never re-read here: keep L2 for the input rows and the scratch).  src 16-byte aligned; dst arbitrary.
// kReadOnly: the source is kernel-read-only input (ld.global.nc); otherwise it was written by this warp (ld.global.cg).
template <bool kReadOnly>
__device__ __forceinline__ uint32_t ld_stream32(const uint32_t *p) { return kReadOnly ? __ldg(p) : __ldcg(p); }
template <bool kReadOnly>
__device__ __forceinline__ uint4 ld_stream128(const uint4 *p) { return kReadOnly ? __ldg(p) : __ldcg(p); }

template <bool kReadOnly>
__device__ __forceinline__ void warp_copy_stream(uint8_t *dst, const uint8_t *__restrict__ src, uint32_t n, unsigned lane) {
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u);
    if (mis == 0) {
        const uint4 *sv = reinterpret_cast<const uint4 *>(src);
        uint4 *dv = reinterpret_cast<uint4 *>(dst);
        const uint32_t nvec = n >> 4;
        uint32_t k = lane;
        for (; k + 96 < nvec; k += 128) {
            const uint4 a = ld_stream128<kReadOnly>(sv + k), b = ld_stream128<kReadOnly>(sv + k + 32),
                        c = ld_stream128<kReadOnly>(sv + k + 64), d = ld_stream128<kReadOnly>(sv + k + 96);
            __stcs(dv + k, a); __stcs(dv + k + 32, b); __stcs(dv + k + 64, c); __stcs(dv + k + 96, d);
        }
        for (; k < nvec; k += 32) __stcs(dv + k, ld_stream128<kReadOnly>(sv + k));
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
        const uint32_t w0 = ld_stream32<kReadOnly>(q), w1 = ld_stream32<kReadOnly>(q + 1), w2 = ld_stream32<kReadOnly>(q + 2),
                       w3 = ld_stream32<kReadOnly>(q + 3);
        const uint32_t w4 = bsh ? ld_stream32<kReadOnly>(q + 4) : 0u;
        uint4 o;
        o.x = __funnelshift_r(w0, w1, bsh);
        o.y = __funnelshift_r(w1, w2, bsh);
        o.z = __funnelshift_r(w2, w3, bsh);
        o.w = __funnelshift_r(w3, w4, bsh);
        __stcs(dv + k, o);
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

#ifndef SKY_COOP_LIT
#define SKY_COOP_LIT 16
#endif
#ifndef SKY_MAX_STEP_LOG
#define SKY_MAX_STEP_LOG 4
#endif
constexpr uint32_t kCoopLit = SKY_COOP_LIT;          // literal runs at least this long are copied by the whole warp
constexpr uint32_t kMaxStepLog = SKY_MAX_STEP_LOG;   // probe stride doubles after a tile without a hit, up to 1 << this
constexpr int kGroups = 8;                           // a tile = 8 warp-wide groups = 256 probe slots
constexpr uint32_t kTile = kGroups * 32;
constexpr uint32_t kOffsBytes = kTile * 2;           // per-warp u16 offsets of the current tile's slots
constexpr uint32_t kLz4AreaBytes = kTableBytes + kOffsBytes;
constexpr uint32_t kScratchBytes = kBlock + 1024;    // per-warp compressed-block scratch (output can overshoot L by < 300 B)

__device__ __forceinline__ uint32_t div255(uint32_t x) { return (x * 0x8081u) >> 23; }  // exact for x < 65536

// ---- emission of up to 32 recorded sequences, one per lane ------------------------------------------
// q0 = literal length | match length << 16, q1 = literal start | offset << 16 (lane k = k-th sequence).
// Sizes go through a warp scan; every lane writes its own token, length bytes, literals (runs >= kCoopLit: one warp
// copy each), offset and match-length bytes.  Returns the new output cursor.  Not inlined: called from two places.
__device__ __noinline__ uint32_t flush_seqs(uint8_t *__restrict__ out, uint32_t op, const uint8_t *__restrict__ src, uint32_t q0,
                                            uint32_t q1, uint32_t nseq, unsigned lane) {
    const bool act = lane < nseq;
    const uint32_t ll = q0 & 0xffffu, ml = q0 >> 16, lit = q1 & 0xffffu, off = q1 >> 16;
    const uint32_t mcode = ml - kMinMatch;
    const uint32_t nl = ll >= 15 ? div255(ll - 15) + 1 : 0, nm = mcode >= 15 ? div255(mcode - 15) + 1 : 0;
    const uint32_t sz = act ? 3 + ll + nl + nm : 0;
    uint32_t incl = sz;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, d);
        if ((int)lane >= d) incl += t;
    }
    const uint32_t total = __shfl_sync(kFull, incl, 31);
    uint32_t lit_o = 0;
    if (act) {
        uint32_t o = op + incl - sz;
        out[o++] = (uint8_t)(((ll < 15 ? ll : 15) << 4) | (mcode < 15 ? mcode : 15));
        if (nl) {
            for (uint32_t k = 0; k + 1 < nl; k++) out[o + k] = 255;
            out[o + nl - 1] = (uint8_t)(ll - 15 - (nl - 1) * 255);
            o += nl;
        }
        lit_o = o;
        if (ll < kCoopLit) {
            uint32_t k = 0;
            for (; k + 4 <= ll; k += 4) {  // 4 literal bytes per round: one unaligned read, four byte stores
                const uint32_t w = load32(src, lit + k);
                out[o + k] = (uint8_t)w;
                out[o + k + 1] = (uint8_t)(w >> 8);
                out[o + k + 2] = (uint8_t)(w >> 16);
                out[o + k + 3] = (uint8_t)(w >> 24);
            }
            const uint32_t rem = ll - k;
            if (rem) {  // (literals of a match sequence end >= 12 bytes before the block end: the word read is in range)
                const uint32_t w = load32(src, lit + k);
                out[o + k] = (uint8_t)w;
                if (rem > 1) out[o + k + 1] = (uint8_t)(w >> 8);
                if (rem > 2) out[o + k + 2] = (uint8_t)(w >> 16);
            }
        }
        o += ll;
        out[o] = (uint8_t)off;
        out[o + 1] = (uint8_t)(off >> 8);
        o += 2;
        if (nm) {
            for (uint32_t k = 0; k + 1 < nm; k++) out[o + k] = 255;
            out[o + nm - 1] = (uint8_t)(mcode - 15 - (nm - 1) * 255);
        }
    }
    __syncwarp();
    unsigned big = __ballot_sync(kFull, act && ll >= kCoopLit);
    while (big) {  // long literal runs: one warp-wide copy each
        const int l = __ffs(big) - 1;
        big &= big - 1;
        warp_copy(out + __shfl_sync(kFull, lit_o, l), src + __shfl_sync(kFull, lit, l), __shfl_sync(kFull, ll, l), lane);
    }
    return op + total;
}

// ---- the block compressor -------------------------------------------------------------------------
// src: block start in the chunk (16-byte aligned), L: block length (1..65536), out: this warp's scratch (kScratchBytes),
// tab: this warp's shared-memory area (kLz4AreaBytes: match table, then the tile's slot offsets).
// Returns the compressed size (1..L-1), or 0 if the block does not shrink (caller stores it raw).
//
// The block is parsed tile by tile (kTile probe slots at stride 1 << slog), three decoupled phases per tile, each with
// all 32 lanes busy (tools/lz4_tile_model.c is the sequential twin of exactly these rules):
//   pass 1  every slot: read 5 bytes, hash, ONE shared-memory lookup of (pos16 | tag16); a hit is a tag match, no byte of
//           the candidate is read.  Lanes of a group that hash alike are ordered with one match.any so the result equals
//           sequential insertion (nearest lower lane = most recent occurrence; the highest lane stores).
//   pass 2  greedy parse over the hit bit masks: per accepted match one warp-wide round compares 24 bytes ahead and 8
//           bytes behind (coalesced byte loads) -- verification, forward and backward extension in one ballot; longer
//           matches continue 128 bytes per round.  Sequences are recorded one per lane.
//   pass 3  every 32 sequences: flush_seqs (scan of sizes, lane-parallel emission).
__device__ __forceinline__ uint32_t lz4_compress_block(const uint8_t *__restrict__ src, uint32_t L, uint8_t *__restrict__ out,
                                                      uint32_t *tab, unsigned lane) {
    {   // clear the table: 32 lanes x 16 B per round; entry 0 = (position 0, tag 0) doubles as "empty"
        uint4 *t4 = reinterpret_cast<uint4 *>(tab);
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll 4
        for (uint32_t k = lane; k < kEntries / 4; k += 32) t4[k] = z;
    }
    uint16_t *offs = reinterpret_cast<uint16_t *>(tab + kEntries);
    __syncwarp();

    uint32_t anchor = 0, cur = 0, op = 0, nseq = 0, q0 = 0, q1 = 0;
    const uint32_t limit = L - 1;  // accept only csize <= L-1 (LZ4F_makeBlock passes dstCapacity = srcSize-1)
    const unsigned lt_mask = (1u << lane) - 1u;

    if (L >= kMfLimit + 1) {
        const uint32_t mflimit = L - kMfLimit;          // last position a match may start at
        const uint32_t matchlimit = L - kLastLiterals;  // matches end at or before this
        const uint32_t *W = reinterpret_cast<const uint32_t *>(src);
        const int dl = lane < 24 ? (int)lane : 23 - (int)lane;  // byte this lane compares: +0..+23 ahead, -1..-8 behind
        uint32_t tb = 0, slog = 0;
        while (tb <= mflimit) {
            const uint32_t span = kTile << slog;
            if (lane == 0) {  // pull the tile after next towards L2 (the MD5 lanes of the chunk share it)
                const uint32_t pf = (tb + 2 * span) & ~15u;
                if (pf < L) l2_prefetch_bulk(src + pf, min(span, ((L - pf) + 15u) & ~15u));
            }
            // ---- pass 1
            uint32_t mymask = 0;
#pragma unroll
            for (int g = 0; g < kGroups; g++) {
                const uint32_t p = tb + ((uint32_t)(g * 32 + lane) << slog);
                const bool valid = p <= mflimit;
                uint32_t mine = 0, idx = 0xffff0000u | lane, e = 0;
                if (valid) {
                    const uint32_t w0 = __ldg(W + (p >> 2)), w1 = __ldg(W + (p >> 2) + 1);
                    const uint32_t sh = (p & 3u) * 8u;
                    uint32_t hf = __funnelshift_r(w0, w1, sh) * 2654435761u;
                    hf = ((w1 >> sh) & 0xffu) * 0x85EBCA6Bu + hf;  // fifth byte
                    idx = __umulhi(hf, kEntries);
                    mine = __byte_perm(p, hf, 0x6510);  // pos16 | hash bytes 1-2 as the tag
                    e = tab[idx];
                }
                const unsigned grp = __match_any_sync(kFull, idx);
                const unsigned lower = grp & lt_mask;
                const uint32_t e_near = __shfl_sync(kFull, mine, lower ? 31 - __clz(lower) : 0);
                if (lower) e = e_near;  // a lower lane of this group filled the slot more recently than the table knows
                const uint32_t x = e ^ mine;
                const unsigned hits = __ballot_sync(kFull, valid && x < 65536u && x != 0u);  // (also: table reads before writes)
                if ((int)lane == g) mymask = hits;
                offs[g * 32 + lane] = (uint16_t)(mine - e);
                if (valid && (grp >> lane) == 1u) tab[idx] = mine;
                __syncwarp();
            }
            // ---- pass 2
            unsigned nz = __ballot_sync(kFull, mymask != 0u);
            const bool anyhit = nz != 0u;
            bool accepted = false;
            uint32_t m = 0, gbase = 0;
            for (;;) {
                if (m == 0) {
                    if (nz == 0) break;
                    const int g = __ffs(nz) - 1;
                    nz &= nz - 1;
                    m = __shfl_sync(kFull, mymask, g);
                    gbase = tb + ((uint32_t)(g * 32) << slog);
                }
                if (cur > gbase) {  // drop the slots the cursor has passed
                    m &= __funnelshift_lc(0u, 0xffffffffu, (cur - gbase + (1u << slog) - 1u) >> slog);
                    if (m == 0) continue;
                }
                const uint32_t bit = (uint32_t)__ffs(m) - 1u;
                m &= m - 1;
                uint32_t pos = gbase + (bit << slog);
                const uint32_t off = offs[((gbase - tb) >> slog) + bit];
                const uint32_t cand = pos - off;
                const uint32_t maxlen = matchlimit - pos;  // >= 7
                const uint32_t room = min(min(pos - anchor, cand), 8u);
                bool ok = false;
                if (lane < 24 ? lane < maxlen : lane - 24 < room) ok = src[(int)pos + dl] == src[(int)cand + dl];
                const unsigned eq = __ballot_sync(kFull, ok);
                const int f = __ffs(~eq & 0xffffffu);
                uint32_t mlen = f ? (uint32_t)(f - 1) : 24u;
                if (mlen < kMinMatch) continue;  // tag collision (or only 4 of the 5 hashed bytes... still a valid match if >= 4)
                if (mlen == 24u && maxlen > 24u) mlen = extend_coop(src, pos, cand, 24u, maxlen, lane);
                const uint32_t back = (uint32_t)__ffs(~(eq >> 24)) - 1u;
                pos -= back;
                mlen += back;
                const uint32_t r0 = (pos - anchor) | (mlen << 16), r1 = anchor | (off << 16);
                if (lane == nseq) {
                    q0 = r0;
                    q1 = r1;
                }
                nseq++;
                anchor = cur = pos + mlen;
                accepted = true;
                if (nseq == 32) {
                    op = flush_seqs(out, op, src, q0, q1, 32, lane);
                    nseq = 0;
                    if (op + 1 + kLastLiterals > limit) return 0;  // cannot end up smaller than the input
                }
            }
            tb = max(tb + span, cur);
            if (accepted) slog = 0;
            else if (!anyhit && slog < kMaxStepLog) slog++;
        }
    }
    if (nseq) {
        op = flush_seqs(out, op, src, q0, q1, nseq, lane);
        if (op + 1 + kLastLiterals > limit) return 0;
    }
    // last literals
    const uint32_t last = L - anchor;
    if (op + seq_bytes(last, 0) > limit) return 0;
    op = emit_seq(out, op, src, anchor, last, 0, 0, lane);
    return op;
}

}  // namespace sky
