// deflate_stats.c -- token statistics of the DEFLATE streams in a BGZF file (literals, matches, lengths, distances, blocks):
// the numbers the inflate kernels' design is sized by.  gcc -O2 -o tools/bin/deflate_stats tools/deflate_stats.c
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
typedef struct { const uint8_t *p; uint64_t bit; } BR;
static uint32_t bits(BR *b, int n) { uint32_t v = 0; for (int i = 0; i < n; ++i) { v |= (uint32_t)((b->p[b->bit >> 3] >> (b->bit & 7)) & 1) << i; b->bit++; } return v; }
typedef struct { uint16_t cnt[16], sym[320]; } H;
static void build(H *h, const uint8_t *l, int n) { uint16_t offs[16]; memset(h->cnt, 0, sizeof h->cnt); for (int i = 0; i < n; ++i) h->cnt[l[i]]++; h->cnt[0] = 0; offs[1] = 0; for (int i = 1; i < 15; ++i) offs[i + 1] = offs[i] + h->cnt[i]; for (int i = 0; i < n; ++i) if (l[i]) h->sym[offs[l[i]]++] = i; }
static int dec(BR *b, const H *h, int *clen) { int code = 0, first = 0, idx = 0; for (int l = 1; l < 16; ++l) { code |= bits(b, 1); int c = h->cnt[l]; if (code - c < first) { *clen = l; return h->sym[idx + (code - first)]; } idx += c; first += c; first <<= 1; code <<= 1; } return -1; }
static const uint16_t lb[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258}, le[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
static const uint16_t db[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577}, de[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
int main(int argc, char **argv) {
    FILE *f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); uint8_t *d = malloc(n + 8); if (fread(d, 1, n, f) != (size_t)n) return 1;
    int maxm = argc > 2 ? atoi(argv[2]) : 64;
    uint64_t llh[16]={0}, dlh[16]={0}; uint64_t lits = 0, matches = 0, mbytes = 0, blocks = 0, members = 0, out = 0, dh[6] = {0}, lh[5] = {0}, longlit = 0, longdist = 0, litbits = 0, inb = 0, dep = 0, litrun_tokens = 0;
    long p = 0;
    while (p + 18 <= n && (int)members < maxm) {
        int bsize = (d[p + 16] | d[p + 17] << 8) + 1; BR b = {d + p + 18, 0}; uint32_t pos = 0; int last, prev_lit = 0;
        do { last = bits(&b, 1); int t = bits(&b, 2); blocks++; H hl, hd; uint8_t l[320];
            if (t == 0) { b.bit = (b.bit + 7) & ~7ull; int len = bits(&b, 16); bits(&b, 16); b.bit += 8ull * len; pos += len; lits += len; continue; }
            if (t == 1) { for (int i = 0; i < 288; ++i) l[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8; build(&hl, l, 288); for (int i = 0; i < 30; ++i) l[i] = 5; build(&hd, l, 30); }
            else { int nl = bits(&b, 5) + 257, nd = bits(&b, 5) + 1, nc = bits(&b, 4) + 4; static const uint8_t ord[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15}; uint8_t cl[19] = {0}; for (int i = 0; i < nc; ++i) cl[ord[i]] = bits(&b, 3); H hc; build(&hc, cl, 19);
                int i = 0, k; while (i < nl + nd) { int s = dec(&b, &hc, &k); if (s < 16) l[i++] = s; else { int r, v = 0; if (s == 16) { v = l[i - 1]; r = 3 + bits(&b, 2); } else if (s == 17) r = 3 + bits(&b, 3); else r = 11 + bits(&b, 7); while (r--) l[i++] = v; } }
                build(&hl, l, nl); build(&hd, l + nl, nd); }
            for (;;) { int k; int s = dec(&b, &hl, &k); llh[k]++; if (s < 256) { lits++; pos++; litbits += k; if (k > 8) longlit++; if (!prev_lit) litrun_tokens++; prev_lit = 1; continue; } if (s == 256) break; prev_lit = 0; s -= 257; int len = lb[s] + bits(&b, le[s]); int ds = dec(&b, &hd, &k); dlh[k]++; if (k > 7) longdist++; int dist = db[ds] + bits(&b, de[ds]);
                matches++; mbytes += len; if ((uint32_t)dist < (uint32_t)len) dep++; pos += len; dh[dist <= 256 ? 0 : dist <= 1024 ? 1 : dist <= 4096 ? 2 : dist <= 8192 ? 3 : dist <= 16384 ? 4 : 5]++; lh[len <= 4 ? 0 : len <= 8 ? 1 : len <= 16 ? 2 : len <= 32 ? 3 : 4]++; }
        } while (!last);
        out += pos; members++; inb += bsize; p += bsize; }
    printf("members %llu blocks %llu in %llu out %llu\nliterals %llu (%.1f%% of out, %.2f bits each, %.1f%% longer than 8 bits) literal runs %llu\nmatches %llu bytes %llu (avg %.1f) overlapping %llu long-dist-codes %.1f%%\n", (unsigned long long)members, (unsigned long long)blocks, (unsigned long long)inb, (unsigned long long)out, (unsigned long long)lits, 100.0 * lits / out, (double)litbits / lits, 100.0 * longlit / lits, (unsigned long long)litrun_tokens, (unsigned long long)matches, (unsigned long long)mbytes, (double)mbytes / matches, (unsigned long long)dep, 100.0 * longdist / matches);
    printf("dist <=256 %.1f%% <=1K %.1f%% <=4K %.1f%% <=8K %.1f%% <=16K %.1f%% <=32K %.1f%%\n", 100.0 * dh[0] / matches, 100.0 * dh[1] / matches, 100.0 * dh[2] / matches, 100.0 * dh[3] / matches, 100.0 * dh[4] / matches, 100.0 * dh[5] / matches);
    printf("len <=4 %.1f%% <=8 %.1f%% <=16 %.1f%% <=32 %.1f%% >32 %.1f%%\n", 100.0 * lh[0] / matches, 100.0 * lh[1] / matches, 100.0 * lh[2] / matches, 100.0 * lh[3] / matches, 100.0 * lh[4] / matches);
    printf("lit/len code lengths:"); for (int i=1;i<16;++i) printf(" %d:%.1f%%", i, 100.0*llh[i]/(lits+matches+blocks)); printf("\ndist code lengths:"); for (int i=1;i<16;++i) printf(" %d:%.1f%%", i, 100.0*dlh[i]/matches); printf("\n");
    return 0; }
