/*
 * lz4_window_model.c -- sequential CPU model of the GPU block compressor in skyplane_b200/csrc/lz4.cuh.
 *
 * NOT the reference's algorithm (that is oracle/skyoracle.c) and NOT product code: a development tool that replays
 * the kernel's windowed multi-match parse on the host so design options (table size, window width, insertion policy,
 * skip rule) can be compared for compression ratio without a GPU.  tools/ratio_study.py drives it.
 *
 * Per window of W positions (stride = skip step): every position reads the table as it was BEFORE the window; positions
 * that hash alike form a group -- the nearest lower member with equal 4 bytes is an in-window candidate and beats the
 * table's; every hit is extended to its full length; hits are accepted greedily in position order (a hit starting inside
 * an accepted match is dropped); step > 1 probes also extend backwards; the highest member of each group stores its
 * position; the cursor moves to max(ip + W*step, end of last accepted match).
 *
 * Build: gcc -O2 -shared -fPIC -o tools/bin/liblz4model.so tools/lz4_window_model.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MINMATCH 4
#define MFLIMIT 12
#define LASTLITERALS 5

typedef struct {
    int hash_log;      /* table entries = 1 << hash_log (kernel: 12) */
    int window;        /* positions per iteration (kernel: 32) */
    int skip_trigger;  /* LZ4: 6 */
    int in_window;     /* 1 = use in-window candidates (kernel: 1) */
    int back_ext;      /* 1 = backward extension when step > 1 (kernel: 1); 2 = always */
} model_opts;

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

/* returns compressed size, or 0 if the block does not shrink (store raw); out capacity >= L */
uint32_t model_compress_block(const uint8_t *src, uint32_t L, uint8_t *out, const model_opts *o) {
    const uint32_t tsize = 1u << o->hash_log;
    uint16_t *ht = calloc(tsize, 2);
    const int W = o->window > 64 ? 64 : o->window;
    uint32_t *pos = malloc(sizeof(uint32_t) * W * 6);
    uint32_t *cand = pos + W, *mlen = pos + 2 * W, *h = pos + 3 * W, *v = pos + 4 * W, *hit = pos + 5 * W;
    uint32_t ip = 0, anchor = 0, op = 0, limit = L - 1, result = 0;
    if (L >= MFLIMIT + 1) {
        const uint32_t mflimit = L - MFLIMIT, matchlimit = L - LASTLITERALS;
        uint32_t nprobe = 1u << o->skip_trigger;
        while (ip <= mflimit) {
            const uint32_t step = nprobe >> o->skip_trigger;
            int nvalid = 0, anyhit = 0;
            for (int l = 0; l < W; l++) {
                pos[l] = ip + l * step;
                if (pos[l] > mflimit) break;
                nvalid++;
                v[l] = rd32(src + pos[l]);
                h[l] = (v[l] * 2654435761u) >> (32 - o->hash_log);
                cand[l] = ht[h[l]];
                hit[l] = cand[l] < pos[l] && rd32(src + cand[l]) == v[l];
            }
            for (int l = 0; l < nvalid && o->in_window; l++)
                for (int k = l - 1; k >= 0; k--)
                    if (h[k] == h[l]) {  /* nearest lower member of the hash group */
                        if (v[k] == v[l]) { cand[l] = pos[k]; hit[l] = 1; }
                        break;
                    }
            for (int l = 0; l < nvalid; l++) anyhit |= hit[l];
            if (!anyhit) {
                for (int l = 0; l < nvalid; l++) {  /* highest member of each group wins = last writer in lane order */
                    ht[h[l]] = (uint16_t)pos[l];
                }
                ip += W * step;
                nprobe += W;
                continue;
            }
            for (int l = 0; l < nvalid; l++) {
                mlen[l] = 0;
                if (!hit[l]) continue;
                uint32_t maxlen = matchlimit - pos[l], m = MINMATCH;
                while (m < maxlen && src[pos[l] + m] == src[cand[l] + m]) m++;
                mlen[l] = m;
            }
            /* greedy acceptance in position order, on the un-extended positions (as the kernel does) */
            int acc[64], nacc = 0;
            uint32_t cur_end = anchor;
            for (int l = 0; l < nvalid; l++)
                if (hit[l] && pos[l] >= cur_end) { acc[nacc++] = l; cur_end = pos[l] + mlen[l]; }
            /* literal lengths, optional backward extension, size check, emission */
            uint32_t prev_end = anchor, total = 0;
            uint32_t P[64], C[64], M[64], LL[64];
            for (int a = 0; a < nacc; a++) {
                int l = acc[a];
                uint32_t p = pos[l], c = cand[l], m = mlen[l];
                if (o->back_ext == 2 || (o->back_ext == 1 && step > 1)) {
                    uint32_t room = p - prev_end < c ? p - prev_end : c, bk = 0;
                    while (bk < room && src[p - 1 - bk] == src[c - 1 - bk]) bk++;
                    p -= bk; c -= bk; m += bk;
                }
                P[a] = p; C[a] = c; M[a] = m; LL[a] = p - prev_end;
                total += seq_bytes(LL[a], m);
                prev_end = pos[l] + mlen[l];
            }
            if (op + total + 1 + LASTLITERALS > limit) goto done; /* cannot shrink: stored raw */
            prev_end = anchor;
            for (int a = 0; a < nacc; a++) {
                op = emit(out, op, src, prev_end, LL[a], M[a], P[a] - C[a]);
                prev_end = P[a] + M[a];
            }
            for (int l = 0; l < nvalid; l++) ht[h[l]] = (uint16_t)pos[l]; /* highest member of a group = last writer */
            anchor = prev_end;
            ip = ip + W * step > anchor ? ip + W * step : anchor;
            nprobe = 1u << o->skip_trigger;
        }
    }
    {
        uint32_t last = L - anchor;
        if (op + seq_bytes(last, 0) > limit) goto done;
        op = emit(out, op, src, anchor, last, 0, 0);
        result = op;
    }
done:
    free(ht);
    free(pos);
    return result;
}

/* whole chunk -> total frame size with independent 64 KiB blocks (header 15 + blocks + endmark 4) */
uint64_t model_frame_size(const uint8_t *src, uint64_t n, const model_opts *o, uint8_t *scratch /* >= 65536 */) {
    uint64_t total = n ? 15 : 7, pos = 0;
    while (pos < n) {
        uint32_t L = (uint32_t)(n - pos < 65536 ? n - pos : 65536);
        uint32_t c = model_compress_block(src + pos, L, scratch, o);
        total += 4 + (c ? c : L);
        pos += L;
    }
    return total + 4;
}
