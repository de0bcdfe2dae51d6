/*
 * lz4_tile_model.c -- sequential CPU twin of the round-2 GPU block compressor (skyplane_b200/csrc/lz4.cuh).
 *
 * NOT the reference's algorithm (that is oracle/skyoracle.c) and NOT product code: a development tool.  It replays the
 * kernel's tile parse on the host, byte for byte, so (a) design options can be compared for compression ratio without a
 * GPU (tools/ratio_study.py) and (b) the GPU output can be diffed against a sequential implementation of the same rules
 * (tests/test_tools_model.py, tests/test_gpu_parity.py).
 *
 * The parse, per 64 KiB block:
 *   tile    = T probe slots, slot i at position tbase + i*step (step 1, doubled after a tile without any hit, up to
 *             max_step; back to 1 after a tile with an accepted match);
 *   pass 1  = every slot, in position order: hash the 4 bytes, look the table entry (pos16 | tag) up, hit = tag equal and
 *             entry older than the slot; the slot then replaces the entry.  No input byte of the candidate is read here.
 *   pass 2  = greedy parse over the tile's hits at or after the cursor: measure the real match length from byte 0 (this
 *             is also the verification), drop it if < 4, optionally extend backwards (<= 8 bytes, not into the previous
 *             sequence), record the sequence, move the cursor to the match end.
 *   next tile base = max(tile end, cursor): tiles wholly inside a long match are never probed.
 *   emission = standard LZ4 sequences; a block that would not shrink is stored raw (return 0).
 *
 * Build: gcc -O2 -shared -fPIC -o tools/bin/liblz4tile.so tools/lz4_tile_model.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MINMATCH 4
#define MFLIMIT 12
#define LASTLITERALS 5

typedef struct {
    int entries;    /* table entries (kernel: 4096; any value >= 2, index = mulhi(hash, entries)) */
    int tag_bits;   /* tag width kept beside the 16-bit position (kernel: 16) */
    int tile;       /* probe slots per tile (kernel: 1024) */
    int max_step;   /* largest probe stride (power of two; 1 = no skip acceleration) */
    int back_ext;   /* 0 = never, 1 = when step > 1, 2 = always */
    int ways;       /* 1 = direct mapped, 2 = two-way FIFO sets (entries/2 sets) */
    int policy;     /* 0 = every slot is inserted; 1 = a hit that continues its left neighbour's match (same offset) is not */
    int hash5;      /* 1 = liblz4's 5-byte hash (64-bit multiply); 2 = 5 bytes mixed with two 32-bit multiplies */
} tile_opts;

static uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }

static uint32_t emit(uint8_t *out, uint32_t op, const uint8_t *src, uint32_t anchor, uint32_t ll, uint32_t ml, uint32_t off) {
    uint32_t mcode = ml ? ml - MINMATCH : 0;
    out[op++] = (uint8_t)(((ll < 15 ? ll : 15) << 4) | (mcode < 15 ? mcode : 15));
    if (ll >= 15) { uint32_t r = ll - 15; for (; r >= 255; r -= 255) out[op++] = 255; out[op++] = (uint8_t)r; }
    memcpy(out + op, src + anchor, ll); op += ll;
    if (ml) {
        out[op++] = (uint8_t)off; out[op++] = (uint8_t)(off >> 8);
        if (mcode >= 15) { uint32_t r = mcode - 15; for (; r >= 255; r -= 255) out[op++] = 255; out[op++] = (uint8_t)r; }
    }
    return op;
}
static uint32_t seq_bytes(uint32_t ll, uint32_t ml) {
    uint32_t s = 1 + ll + (ll >= 15 ? (ll - 15) / 255 + 1 : 0);
    if (ml) s += 2 + ((ml - 4) >= 15 ? (ml - 4 - 15) / 255 + 1 : 0);
    return s;
}

typedef struct { uint64_t probes, hits, verified, accepted, tiles; } tile_stats;
static tile_stats g_stats;
void tile_model_stats(tile_stats *s, int reset) { *s = g_stats; if (reset) memset(&g_stats, 0, sizeof g_stats); }

/* returns compressed size, or 0 if the block does not shrink (store raw); out capacity >= L */
uint32_t tile_compress_block(const uint8_t *src, uint32_t L, uint8_t *out, const tile_opts *o) {
    const uint32_t T = (uint32_t)o->tile, limit = L - 1;
    const int nsets = o->entries / (o->ways > 1 ? o->ways : 1);
    uint32_t *tab = calloc((size_t)o->entries, 4);  /* pos | tag << 16 ; 0 = (pos 0, tag 0) = "empty" */
    uint16_t *off = malloc(2 * T);
    uint8_t *hit = malloc(T);
    const uint32_t tag_mask = o->tag_bits >= 16 ? 0xffffu : ((1u << o->tag_bits) - 1u);
    uint32_t anchor = 0, cur = 0, op = 0, result = 0;
    if (L >= MFLIMIT + 1) {
        const uint32_t mflimit = L - MFLIMIT, matchlimit = L - LASTLITERALS;
        uint32_t tbase = 0, step = 1;
        while (tbase <= mflimit) {
            g_stats.tiles++;
            /* ---- pass 1: groups of 32 slots = one warp instruction; lookups see the table as of the group start plus
             * the lower slots of the same group (nearest first), exactly like sequential insertion */
            int anyhit = 0;
            uint32_t prev_hit = 0, prev_off = 0;
            for (uint32_t g = 0; g < T; g += 32) {
                uint32_t idx[32], ent[32], valid[32], ins[32];
                for (uint32_t l = 0; l < 32; l++) {
                    const uint32_t i = g + l, p = tbase + i * step;
                    hit[i] = 0; valid[l] = p <= mflimit; ins[l] = 0;
                    if (!valid[l]) continue;
                    uint32_t hf;
                    if (o->hash5 == 1) {
                        uint64_t v = 0; memcpy(&v, src + p, p + 8 <= L ? 8 : L - p);
                        hf = (uint32_t)(((v << 24) * 889523592379ULL) >> 32);
                    } else if (o->hash5 == 2) {  /* 5 bytes with 32-bit arithmetic: two IMADs in the kernel */
                        hf = rd32(src + p) * 2654435761u + (p + 4 < L ? src[p + 4] : 0) * 0x85EBCA6Bu;
                    } else hf = rd32(src + p) * 2654435761u;
                    idx[l] = (uint32_t)(((uint64_t)hf * (uint32_t)nsets) >> 32);
                    const uint32_t tag = (hf >> 8) & tag_mask;  /* bytes 1-2 of the hash: one PRMT in the kernel */
                    ent[l] = p | (tag << 16);
                    g_stats.probes++;
                    /* set content at this slot's time: lower same-set slots (nearest first), then the stored ways */
                    uint32_t c[2], nc = 0;
                    for (int k = (int)l - 1; k >= 0 && nc < (uint32_t)o->ways; k--)
                        if (valid[k] && idx[k] == idx[l]) c[nc++] = ent[k];
                    for (int w = 0; w < o->ways && nc < (uint32_t)o->ways; w++) c[nc++] = tab[idx[l] * o->ways + w];
                    for (uint32_t w = 0; w < nc; w++) {
                        const uint32_t epos = c[w] & 0xffffu;
                        if ((c[w] >> 16) == tag && epos < p) { hit[i] = 1; off[i] = (uint16_t)(p - epos); anyhit = 1; g_stats.hits++; break; }
                    }
                    ins[l] = !(o->policy == 1 && hit[i] && prev_hit && off[i] == prev_off);
                    prev_hit = hit[i]; prev_off = off[i];
                }
                /* stores: per set, the highest slot decides (and stores only if it is an inserting slot) */
                for (int l = 31; l >= 0; l--) {
                    if (!valid[l]) continue;
                    int highest = 1;
                    for (int k = l + 1; k < 32; k++) if (valid[k] && idx[k] == idx[l]) highest = 0;
                    if (!highest || !ins[l]) continue;
                    if (o->ways == 2) {
                        uint32_t second = tab[idx[l] * 2];
                        for (int k = l - 1; k >= 0; k--) if (valid[k] && idx[k] == idx[l]) { second = ent[k]; break; }
                        tab[idx[l] * 2 + 1] = second;
                        tab[idx[l] * 2] = ent[l];
                    } else tab[idx[l]] = ent[l];
                }
            }
            /* ---- pass 2 */
            int accepted = 0;
            for (uint32_t i = 0; i < T; i++) {
                if (!hit[i]) continue;
                uint32_t pos = tbase + i * step;
                if (pos < cur) continue;
                uint32_t cand = pos - off[i];
                const uint32_t maxlen = matchlimit - pos;
                uint32_t mlen = 0;
                while (mlen < maxlen && src[pos + mlen] == src[cand + mlen]) mlen++;
                if (mlen < MINMATCH) continue;
                g_stats.verified++;
                if (o->back_ext == 2 || (o->back_ext == 1 && step > 1)) {
                    uint32_t room = pos - anchor;
                    if (cand < room) room = cand;
                    if (room > 8) room = 8;   /* the kernel checks 8 bytes behind the match in the same round as the first 24 ahead */
                    uint32_t b = 0;
                    while (b < room && src[pos - 1 - b] == src[cand - 1 - b]) b++;
                    pos -= b; cand -= b; mlen += b;
                }
                const uint32_t ll = pos - anchor;
                if (op + seq_bytes(ll, mlen) + 1 + LASTLITERALS > limit) goto done;  /* cannot end up smaller */
                op = emit(out, op, src, anchor, ll, mlen, pos - cand);
                anchor = cur = pos + mlen;
                accepted = 1;
                g_stats.accepted++;
            }
            const uint32_t tile_end = tbase + T * step;
            tbase = tile_end > cur ? tile_end : cur;
            if (accepted) step = 1;
            else if (!anyhit && step < (uint32_t)o->max_step) step <<= 1;
        }
    }
    {
        const uint32_t last = L - anchor;
        if (op + seq_bytes(last, 0) > limit) goto done;
        op = emit(out, op, src, anchor, last, 0, 0);
        result = op;
    }
done:
    free(tab); free(off); free(hit);
    return result;
}
