/*
 * oracle/bam_oracle.c -- CPU restatement of the reference's BAM scan -> flag / MAPQ filter -> per-reference COUNT
 * path (BASELINE.json configs[3]).
 *
 * TEST INFRASTRUCTURE ONLY (same rule as vcf_oracle.c).
 *
 * Restates (paths relative to /root/reference):
 *   - exon/exon-core/src/datasources/bam/file_opener.rs:39          BGZF reader over the object, header + reference sequences
 *   - exon/exon-bam/src/batch_reader.rs:70-107                      read_record_buf per row, batch_size rows per batch
 *   - exon/exon-bam/src/array_builder.rs:102-218                    columns: name, flag (u16 -> i32), reference (name of
 *       refID, NULL when refID = -1), start (pos + 1, NULL when pos = -1), end, mapping_quality (decimal STRING,
 *       NULL when 255), cigar text, mate_reference
 *   - exon/exon-bam/src/indexed_async_batch_stream.rs:43-50         end = start + reference-consuming CIGAR length - 1
 *   - exon/exon-core/src/udfs/sam/samflags.rs:26-47, 111-141        is_unmapped / is_secondary / is_supplementary = flag bit tests
 *   - noodles-bam 0.72 / noodles-bgzf 0.34 (Cargo.lock, not vendored): BAM record layout per SAM spec 4.2, restated here
 * The config-4 query `WHERE NOT is_unmapped(flag) AND NOT is_secondary(flag) AND NOT is_supplementary(flag) AND
 * CAST(mapping_quality AS INT) >= 30 GROUP BY reference` is: (flag & 0x904) == 0, mapq != 255 (NULL never passes a
 * comparison), mapq >= 30; groups are reference NAMES, the NULL reference (refID -1) is its own group.
 * Pinned by tests/test_bam_golden.py: slt/bam-select-tests.slt:9-12 (first row READ_ID 83 chr1 12203704 12217173 NULL
 * 55M13394N21M chr1), :56-64 (61 / 122 rows) and SURVEY.md appendix A (flag histogram, all MAPQ 255).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

uint8_t *exo_gunzip_all(const uint8_t *data, int64_t len, int64_t *out_len); /* vcf_oracle.c */

typedef struct {
    uint8_t *buf; /* inflated stream (owned) */
    int64_t len;
    int32_t n_ref;
    const uint8_t **ref_name; /* pointers into buf, NUL terminated */
    int32_t *ref_len;
    int64_t records_at; /* offset of the first record */
} exo_bam;

static inline int32_t rd_i32(const uint8_t *p) { return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24)); }
static inline uint32_t rd_u16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

void exo_bam_close(exo_bam *b) {
    if (!b) return;
    free(b->buf);
    free((void *)b->ref_name);
    free(b->ref_len);
    free(b);
}

/* inflate + header: magic, l_text, text, n_ref, (l_name, name, l_ref)* ; NULL on malformed input */
exo_bam *exo_bam_open(const uint8_t *data, int64_t len) {
    exo_bam *b = (exo_bam *)calloc(1, sizeof(*b));
    b->buf = exo_gunzip_all(data, len, &b->len);
    if (!b->buf || b->len < 12 || memcmp(b->buf, "BAM\1", 4) != 0) { exo_bam_close(b); return NULL; }
    int64_t p = 4;
    const int32_t l_text = rd_i32(b->buf + p);
    p += 4 + l_text;
    if (l_text < 0 || p + 4 > b->len) { exo_bam_close(b); return NULL; }
    b->n_ref = rd_i32(b->buf + p);
    p += 4;
    if (b->n_ref < 0) { exo_bam_close(b); return NULL; }
    b->ref_name = (const uint8_t **)calloc((size_t)b->n_ref + 1, sizeof(uint8_t *));
    b->ref_len = (int32_t *)calloc((size_t)b->n_ref + 1, sizeof(int32_t));
    for (int32_t i = 0; i < b->n_ref; i++) {
        if (p + 4 > b->len) { exo_bam_close(b); return NULL; }
        const int32_t l_name = rd_i32(b->buf + p);
        p += 4;
        if (l_name < 1 || p + l_name + 4 > b->len) { exo_bam_close(b); return NULL; }
        b->ref_name[i] = b->buf + p;
        p += l_name;
        b->ref_len[i] = rd_i32(b->buf + p);
        p += 4;
    }
    b->records_at = p;
    return b;
}
int32_t exo_bam_n_ref(const exo_bam *b) { return b->n_ref; }
const char *exo_bam_ref_name(const exo_bam *b, int32_t i) { return (const char *)b->ref_name[i]; }

/* fields of one record (the columns the reference's slt goldens show) */
typedef struct {
    char name[256];
    int32_t flag, ref_id, mate_ref_id, mapq; /* mapq 255 = NULL */
    int64_t start, end;                     /* 0 = NULL */
    char cigar[1024];
    int32_t l_seq;
    int32_t first_quals[8];
} exo_bam_row;

/* walks the record chain; returns the number of records, -1 on a truncated / malformed chain.  If row_index >= 0 the
 * fields of that record are stored in *row.  counts (n_ref + 1 entries, last = NULL reference) receives the filtered
 * per-reference counts when non-NULL. */
/* region predicate state (bam_region_filter, exon-bam/src/indexed_async_batch_stream.rs:66-86): set before exo_bam_scan */
static _Thread_local int32_t g_region_ref = -3; /* -3: no region; -2: reference absent from this file */
static _Thread_local int64_t g_region_lo = 1, g_region_hi = INT64_MAX;
void exo_bam_set_region(const exo_bam *b, const char *name, int64_t lo, int64_t hi) {
    g_region_ref = -3;
    if (!name) return;
    g_region_ref = -2;
    for (int32_t i = 0; i < b->n_ref; i++)
        if (strcmp((const char *)b->ref_name[i], name) == 0) g_region_ref = i;
    g_region_lo = lo < 1 ? 1 : lo;
    g_region_hi = hi;
}

int64_t exo_bam_scan(const exo_bam *b, int32_t has_pred, uint32_t flag_exclude, uint32_t flag_require, int32_t min_mapq,
                     int64_t *counts, int64_t row_index, exo_bam_row *row) {
    int64_t p = b->records_at, n = 0;
    if (counts) memset(counts, 0, sizeof(int64_t) * ((size_t)b->n_ref + 1));
    while (p < b->len) {
        if (p + 4 > b->len) return -1;
        const int32_t block_size = rd_i32(b->buf + p);
        if (block_size < 32 || p + 4 + block_size > b->len) return -1;
        const uint8_t *r = b->buf + p + 4;
        const int32_t ref_id = rd_i32(r), pos = rd_i32(r + 4);
        const uint32_t l_read_name = r[8], mapq = r[9], n_cigar = rd_u16(r + 12), flag = rd_u16(r + 14);
        const int32_t l_seq = rd_i32(r + 16);
        if (ref_id < -1 || ref_id >= b->n_ref) return -1;
        if (counts) {
            int sel = 1;
            if (has_pred) {
                sel = (flag & flag_exclude) == 0 && (flag & flag_require) == flag_require;
                if (min_mapq >= 0) sel = sel && mapq != 255 && (int32_t)mapq >= min_mapq;
                if (sel && g_region_ref != -3) {
                    sel = ref_id >= 0 && ref_id == g_region_ref && pos >= 0;
                    if (sel) {
                        const uint8_t *c = r + 32 + l_read_name;
                        int64_t span = 0;
                        for (uint32_t i = 0; i < n_cigar; i++) {
                            const uint32_t v = (uint32_t)rd_i32(c + 4 * i), op = v & 15;
                            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += v >> 4;
                        }
                        const int64_t start = (int64_t)pos + 1, end = start + span - 1;
                        sel = end >= 1 && g_region_lo <= end && start <= g_region_hi; /* Interval::intersects */
                    }
                }
            }
            if (sel) counts[ref_id < 0 ? b->n_ref : ref_id]++;
        }
        if (row && n == row_index) {
            memset(row, 0, sizeof(*row));
            /* name: NUL terminated; "*" alone means no name (noodles maps it to None) */
            size_t nl = l_read_name ? l_read_name - 1 : 0;
            if (nl > 255) nl = 255;
            memcpy(row->name, r + 32, nl);
            row->flag = (int32_t)flag;
            row->ref_id = ref_id;
            row->mate_ref_id = rd_i32(r + 20);
            row->mapq = (int32_t)mapq;
            row->start = pos >= 0 ? (int64_t)pos + 1 : 0;
            const uint8_t *c = r + 32 + l_read_name;
            int64_t ref_span = 0;
            size_t w = 0;
            for (uint32_t i = 0; i < n_cigar; i++) {
                const uint32_t v = (uint32_t)rd_i32(c + 4 * i);
                const uint32_t op = v & 15, ln = v >> 4;
                static const char OPS[] = "MIDNSHP=X";
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_span += ln; /* M D N = X consume the reference */
                if (w + 16 < sizeof(row->cigar)) w += (size_t)snprintf(row->cigar + w, sizeof(row->cigar) - w, "%u%c", ln, op < 9 ? OPS[op] : '?');
            }
            row->end = row->start ? row->start + ref_span - 1 : 0;
            row->l_seq = l_seq;
            const uint8_t *q = c + 4 * n_cigar + (l_seq + 1) / 2;
            for (int i = 0; i < 8 && i < l_seq; i++) row->first_quals[i] = (int8_t)q[i];
        }
        p += 4 + (int64_t)block_size;
        n++;
    }
    return n;
}

/* Sequence and quality scores of record `row_index`, as BAMArrayBuilder::append materialises columns 8 and 9
 * (exon-bam/src/array_builder.rs:177-201) from a noodles RecordBuf: bases decoded from the 4-bit encoding with the SAM
 * alphabet "=ACMGRSVTWYHKDBN" (noodles-bam 0.7x record codec, un-vendored), quality scores as the raw bytes
 * reinterpreted as i8 and widened to i64; a quality string that is all 0xFF is "missing" and decodes to an empty list.
 * Returns l_seq (bases written to seq, NUL terminated), -1 when the row does not exist; *n_qual = list length. */
int32_t exo_bam_seq_qual(const exo_bam *b, int64_t row_index, char *seq, int32_t seq_cap, int64_t *qual, int32_t qual_cap, int32_t *n_qual) {
    static const char BASES[] = "=ACMGRSVTWYHKDBN";
    int64_t p = b->records_at, n = 0;
    while (p < b->len) {
        if (p + 4 > b->len) return -1;
        const int32_t block_size = rd_i32(b->buf + p);
        if (block_size < 32 || p + 4 + block_size > b->len) return -1;
        if (n == row_index) {
            const uint8_t *r = b->buf + p + 4;
            const uint32_t l_read_name = r[8], n_cigar = rd_u16(r + 12);
            const int32_t l_seq = rd_i32(r + 16);
            const uint8_t *s = r + 32 + l_read_name + 4 * n_cigar, *q = s + (l_seq + 1) / 2;
            if (l_seq < 0 || l_seq + 1 > seq_cap || l_seq > qual_cap) return -1;
            for (int32_t i = 0; i < l_seq; i++) seq[i] = BASES[(s[i >> 1] >> ((i & 1) ? 0 : 4)) & 15];
            seq[l_seq] = 0;
            int missing = 1;
            for (int32_t i = 0; i < l_seq; i++)
                if (q[i] != 0xFF) missing = 0;
            *n_qual = (missing || l_seq == 0) ? 0 : l_seq;
            for (int32_t i = 0; i < *n_qual; i++) qual[i] = (int64_t)(int8_t)q[i];
            return l_seq;
        }
        p += 4 + (int64_t)block_size;
        n++;
    }
    return -1;
}
