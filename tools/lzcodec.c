/* Small LZ4-block and Snappy-raw COMPRESSORS (test and bench tooling, C twins of tools/blockcodecs.py):
 * pyarrow's ORC writer stores every LZ4 chunk uncompressed, so ORC files with real LZ4 chunks (and Snappy files with
 * the same chunking) are produced by tools/orc_recompress.py with these.  Greedy matcher over a hash table of 4-byte
 * sequences; the output is what any conforming decoder accepts, not what a particular encoder would write.
 * Built by tools/lzcodec.py with gcc. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HASH_BITS 15
static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t hash4(uint32_t v) { return (v * 2654435761u) >> (32 - HASH_BITS); }

/* worst case: LZ4 n + n/255 + 16, Snappy 32 + n + n/6 */
size_t lzc_bound(size_t n) { return n + n / 6 + 64; }

size_t lzc_lz4_compress(const uint8_t* src, size_t n, uint8_t* dst) {
    int32_t* table = (int32_t*)malloc(sizeof(int32_t) << HASH_BITS);
    for (int i = 0; i < (1 << HASH_BITS); i++) table[i] = -1;
    uint8_t* o = dst;
    size_t i = 0, lit = 0;
    /* LZ4 end-of-block rules: the last 5 bytes are literals, the last match starts >= 12 bytes before the end */
    const size_t match_limit = n >= 12 ? n - 12 : 0;
    while (n >= 13 && i <= match_limit) {
        const uint32_t h = hash4(rd32(src + i));
        const int32_t j = table[h];
        table[h] = (int32_t)i;
        if (j >= 0 && i - (size_t)j <= 65535 && rd32(src + j) == rd32(src + i)) {
            size_t ml = 4;
            const size_t max_ml = n - 5 - i;
            while (ml < max_ml && src[j + ml] == src[i + ml]) ml++;
            if (ml >= 4) {
                const size_t ll = i - lit, dist = i - (size_t)j;
                *o++ = (uint8_t)(((ll < 15 ? ll : 15) << 4) | (ml - 4 < 15 ? ml - 4 : 15));
                if (ll >= 15) { size_t v = ll - 15; while (v >= 255) { *o++ = 255; v -= 255; } *o++ = (uint8_t)v; }
                memcpy(o, src + lit, ll); o += ll;
                *o++ = (uint8_t)(dist & 255); *o++ = (uint8_t)(dist >> 8);
                if (ml - 4 >= 15) { size_t v = ml - 4 - 15; while (v >= 255) { *o++ = 255; v -= 255; } *o++ = (uint8_t)v; }
                i += ml;
                lit = i;
                continue;
            }
        }
        i++;
    }
    {
        const size_t ll = n - lit;
        *o++ = (uint8_t)((ll < 15 ? ll : 15) << 4);
        if (ll >= 15) { size_t v = ll - 15; while (v >= 255) { *o++ = 255; v -= 255; } *o++ = (uint8_t)v; }
        memcpy(o, src + lit, ll); o += ll;
    }
    free(table);
    return (size_t)(o - dst);
}

static uint8_t* snappy_literal(uint8_t* o, const uint8_t* p, size_t ll) {
    if (!ll) return o;
    if (ll <= 60) *o++ = (uint8_t)((ll - 1) << 2);
    else if (ll <= 256) { *o++ = 60 << 2; *o++ = (uint8_t)(ll - 1); }
    else if (ll <= 65536) { *o++ = 61 << 2; *o++ = (uint8_t)((ll - 1) & 255); *o++ = (uint8_t)((ll - 1) >> 8); }
    else { *o++ = 62 << 2; *o++ = (uint8_t)((ll - 1) & 255); *o++ = (uint8_t)(((ll - 1) >> 8) & 255); *o++ = (uint8_t)((ll - 1) >> 16); }
    memcpy(o, p, ll);
    return o + ll;
}

size_t lzc_snappy_compress(const uint8_t* src, size_t n, uint8_t* dst) {
    int32_t* table = (int32_t*)malloc(sizeof(int32_t) << HASH_BITS);
    for (int i = 0; i < (1 << HASH_BITS); i++) table[i] = -1;
    uint8_t* o = dst;
    { size_t v = n; while (v >= 128) { *o++ = (uint8_t)(v | 128); v >>= 7; } *o++ = (uint8_t)v; }
    size_t i = 0, lit = 0;
    while (i + 4 <= n) {
        const uint32_t h = hash4(rd32(src + i));
        const int32_t j = table[h];
        table[h] = (int32_t)i;
        if (j >= 0 && i - (size_t)j <= 65535 && rd32(src + j) == rd32(src + i)) {
            size_t ml = 4;
            while (i + ml < n && src[j + ml] == src[i + ml]) ml++;
            const size_t dist = i - (size_t)j;
            o = snappy_literal(o, src + lit, i - lit);
            size_t left = ml;
            while (left > 0) {
                size_t take = left < 64 ? left : 64;
                if (left - take >= 1 && left - take <= 3) take = left - 4 >= 4 ? left - 4 : take;  /* keep >= 4 for the next copy */
                if (take >= 4 && take <= 11 && dist < 2048) {
                    *o++ = (uint8_t)(1 | ((take - 4) << 2) | ((dist >> 8) << 5));
                    *o++ = (uint8_t)(dist & 255);
                } else {
                    *o++ = (uint8_t)(2 | ((take - 1) << 2));
                    *o++ = (uint8_t)(dist & 255); *o++ = (uint8_t)(dist >> 8);
                }
                left -= take;
            }
            i += ml;
            lit = i;
            continue;
        }
        i++;
    }
    o = snappy_literal(o, src + lit, n - lit);
    free(table);
    return (size_t)(o - dst);
}

/* ---- LZO1X (test inputs for the LZO chunk decoder) -------------------------------------------------
 * A greedy compressor that exercises every instruction of the published LZO1X stream format: the first-byte literal
 * forms, literal runs, M1 (2-byte match after 1..3 literals, 3-byte match at 2049..3072 after a run), M2, M3, M4, the
 * 1..3 literals riding on a match instruction, and the end marker. */
static uint8_t* lzo_run_len(uint8_t* o, size_t t) { /* t >= 1: zero bytes of 255 each, then the rest */
    while (t > 255) { *o++ = 0; t -= 255; }
    *o++ = (uint8_t)t;
    return o;
}

/* lzc_lzo_use_m1(0): leave out the two M1 instructions (what the LZO compressors of ORC writers produce; Apache ORC C++'s
 * decoder reads M1 distances 2048 bytes too far back, so files for that reader need this). */
static int lzo_m1 = 1;
void lzc_lzo_use_m1(int on) { lzo_m1 = on; }

size_t lzc_lzo_compress(const uint8_t* src, size_t n, uint8_t* dst) {
    enum { HB = 15 };
    static int32_t h3[1 << HB], h2[1 << 16];
    memset(h3, -1, sizeof h3);
    memset(h2, -1, sizeof h2);
    uint8_t* o = dst;
    uint8_t* sbits = NULL; /* byte holding the S bits of the previous match instruction */
    size_t ip = 0, lit_start = 0;
    int started = 0;
    unsigned state = 0;
    while (ip <= n) {
        size_t mlen = 0, mdist = 0;
        if (ip + 3 <= n) {
            uint32_t v = src[ip] | (src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16);
            uint32_t h = (v * 2654435761u) >> (32 - HB);
            int32_t c = h3[h];
            h3[h] = (int32_t)ip;
            if (c >= 0 && ip - (size_t)c <= 49151) {
                size_t l = 0, lim = n - ip < 3000 ? n - ip : 3000;
                while (l < lim && src[c + l] == src[ip + l]) l++;
                /* (plain mode: no 3- and 4-byte matches beyond 16 KiB either - Apache's decoder takes the 3-byte M4 opcode, 0x11, for
                 * the end marker) */
                if (l >= 3 && (lzo_m1 || l >= 5 || ip - (size_t)c <= 16384)) { mlen = l; mdist = ip - (size_t)c; }
            }
        }
        size_t ll = ip - lit_start;
        if (lzo_m1 && !mlen && ip + 2 <= n && started && ll >= 1 && ll <= 3) {
            uint32_t k = src[ip] | (src[ip + 1] << 8);
            int32_t c = h2[k];
            if (c >= 0 && ip - (size_t)c <= 1024) { mlen = 2; mdist = ip - (size_t)c; }
        }
        if (ip + 2 <= n) h2[src[ip] | (src[ip + 1] << 8)] = (int32_t)ip;
        if (!mlen && ip < n) { ip++; continue; }
        /* ---- pending literals */
        if (ll) {
            if (!started) {
                if (ll <= 238) { *o++ = (uint8_t)(17 + ll); state = ll < 4 ? (unsigned)ll : 4; }
                else { *o++ = 0; o = lzo_run_len(o, ll - 18); state = 4; }
            } else if (ll <= 3) {
                *sbits |= (uint8_t)ll;
                state = (unsigned)ll;
            } else {
                if (ll - 3 <= 15) *o++ = (uint8_t)(ll - 3);
                else { *o++ = 0; o = lzo_run_len(o, ll - 18); }
                state = 4;
            }
            memcpy(o, src + lit_start, ll);
            o += ll;
            started = 1;
        } else if (started) {
            state = 0;
        }
        if (ip >= n) break;
        /* ---- the match */
        if (mlen == 2) {
            size_t d = mdist - 1;
            sbits = o;
            *o++ = (uint8_t)((d & 3) << 2);
            *o++ = (uint8_t)(d >> 2);
        } else if (lzo_m1 && mlen == 3 && state == 4 && mdist >= 2049 && mdist <= 3072) {
            size_t d = mdist - 2049;
            sbits = o;
            *o++ = (uint8_t)((d & 3) << 2);
            *o++ = (uint8_t)(d >> 2);
        } else if (mlen <= 8 && mdist <= 2048) {
            size_t d = mdist - 1;
            sbits = o;
            *o++ = (uint8_t)(((mlen - 1) << 5) | ((d & 7) << 2));
            *o++ = (uint8_t)(d >> 3);
        } else if (mdist <= 16384) {
            if (mlen - 2 <= 31) *o++ = (uint8_t)(32 | (mlen - 2));
            else { *o++ = 32; o = lzo_run_len(o, mlen - 2 - 31); }
            size_t d = mdist - 1;
            sbits = o;
            *o++ = (uint8_t)((d << 2) & 0xff);
            *o++ = (uint8_t)(d >> 6);
        } else {
            size_t d = mdist - 16384;
            uint8_t hbit = (uint8_t)(((d >> 14) & 1) << 3);
            if (mlen - 2 <= 7) *o++ = (uint8_t)(16 | hbit | (mlen - 2));
            else { *o++ = (uint8_t)(16 | hbit); o = lzo_run_len(o, mlen - 2 - 7); }
            d &= 0x3fff;
            sbits = o;
            *o++ = (uint8_t)((d << 2) & 0xff);
            *o++ = (uint8_t)(d >> 6);
        }
        started = 1;
        ip += mlen;
        lit_start = ip;
    }
    *o++ = 0x11;
    *o++ = 0;
    *o++ = 0;
    return (size_t)(o - dst);
}
