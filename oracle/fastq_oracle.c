/*
 * oracle/fastq_oracle.c -- CPU restatement of the reference's FASTQ scan -> mean-quality filter -> COUNT path
 * (BASELINE.json configs[1]).
 *
 * TEST INFRASTRUCTURE ONLY (same rule as vcf_oracle.c): only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may call this.
 *
 * Restates (paths relative to /root/reference):
 *   - exon/exon-fastq/src/batch_reader.rs:56-82      read_record / read_batch: batch_size records per batch
 *   - exon/exon-fastq/src/array_builder.rs:68-102    append: name, description (NULL when empty), sequence,
 *                                                    quality_scores as four Utf8 columns
 *   - noodles-fastq 0.16.0 (Cargo.lock; not vendored) `Reader::read_record`: a record is exactly four lines --
 *     '@' name [' ' description], sequence, '+' ..., quality scores; line terminator '\n' (a '\r' before it is
 *     kept as data here: UNPINNED, no reference test has CRLF input); the name ends at the first ' ' of the
 *     definition line (pinned by slt/fastq-scan-test.slt:6-10: "SEQ_ID" / "This is a description"); a '\t' does NOT
 *     split (UNPINNED); a definition line that does not start with '@' or a third line that does not start with '+'
 *     is an error; end of input before the third line is an error; a missing fourth line is an empty quality
 *     string (UNPINNED, follows the reader's read_line-returns-0 behaviour)
 *   - exon/exon-core/src/udfs/sequence/quality_score_string_to_list.rs:80-93  Phred score = byte - 33
 *   - the config-2 predicate `mean(quality) > T` restated over integers: sum(byte - 33) * den > num * len for
 *     T = num / den (an empty quality string has no mean: the row is not selected)
 * Pinned by tests/test_fastq_golden.py against slt/fastq-scan-test.slt (2 rows, the four column values, NULL
 * description) and slt/quality-score-udfs.slt ('###' -> [2, 2, 2]).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EXO_END 0
#define EXO_BATCH 1
#define EXO_ERR_PARSE (-2)

typedef struct {
    int32_t *offsets;
    uint8_t *values;
    uint8_t *valid; /* one byte per row; only `description` uses it */
    int64_t values_len, values_cap;
} exo_utf8_col;

typedef struct {
    int64_t rows;
    exo_utf8_col name, description, sequence, quality;
} exo_fastq_batch;

typedef struct {
    const uint8_t *text;
    int64_t len, at, batch_size;
    int64_t n_records, err_record;
    exo_fastq_batch b;
} exo_fastq_reader;

static void col_init(exo_utf8_col *c, int64_t rows) {
    c->offsets = (int32_t *)malloc(sizeof(int32_t) * (size_t)(rows + 1));
    c->valid = (uint8_t *)malloc((size_t)rows + 1);
    c->values_cap = 1 << 16;
    c->values = (uint8_t *)malloc((size_t)c->values_cap);
    c->values_len = 0;
}
static void col_free(exo_utf8_col *c) {
    free(c->offsets);
    free(c->values);
    free(c->valid);
}
static void col_append(exo_utf8_col *c, int64_t row, const uint8_t *p, int64_t n, int valid) {
    if (c->values_len + n > c->values_cap) {
        while (c->values_len + n > c->values_cap) c->values_cap *= 2;
        c->values = (uint8_t *)realloc(c->values, (size_t)c->values_cap);
    }
    c->offsets[row] = (int32_t)c->values_len;
    if (n) memcpy(c->values + c->values_len, p, (size_t)n);
    c->values_len += n;
    c->offsets[row + 1] = (int32_t)c->values_len;
    c->valid[row] = (uint8_t)valid;
}

exo_fastq_reader *exo_fastq_reader_open(const uint8_t *text, int64_t len, int64_t batch_size) {
    exo_fastq_reader *r = (exo_fastq_reader *)calloc(1, sizeof(*r));
    r->text = text;
    r->len = len;
    r->batch_size = batch_size > 0 ? batch_size : 8192;
    r->err_record = -1;
    col_init(&r->b.name, r->batch_size);
    col_init(&r->b.description, r->batch_size);
    col_init(&r->b.sequence, r->batch_size);
    col_init(&r->b.quality, r->batch_size);
    return r;
}
void exo_fastq_reader_close(exo_fastq_reader *r) {
    if (!r) return;
    col_free(&r->b.name);
    col_free(&r->b.description);
    col_free(&r->b.sequence);
    col_free(&r->b.quality);
    free(r);
}
int64_t exo_fastq_reader_err_record(const exo_fastq_reader *r) { return r->err_record; }

/* one line starting at r->at: [*p, *p + *n), terminator consumed; returns 0 at end of input */
static int next_line(exo_fastq_reader *r, const uint8_t **p, int64_t *n) {
    if (r->at >= r->len) return 0;
    const uint8_t *s = r->text + r->at;
    const uint8_t *nl = (const uint8_t *)memchr(s, '\n', (size_t)(r->len - r->at));
    *p = s;
    if (nl) {
        *n = nl - s;
        r->at += *n + 1;
    } else {
        *n = r->len - r->at;
        r->at = r->len;
    }
    return 1;
}

/* noodles-fastq Reader::read_record; 1 = record, 0 = end of input, < 0 = error */
static int read_record(exo_fastq_reader *r, const uint8_t **name, int64_t *name_len, const uint8_t **desc, int64_t *desc_len,
                       const uint8_t **seq, int64_t *seq_len, const uint8_t **qual, int64_t *qual_len) {
    const uint8_t *p;
    int64_t n;
    if (!next_line(r, &p, &n)) return 0;
    if (n < 1 || p[0] != '@') return EXO_ERR_PARSE; /* "invalid name prefix" */
    const uint8_t *sp = (const uint8_t *)memchr(p + 1, ' ', (size_t)(n - 1));
    *name = p + 1;
    if (sp) {
        *name_len = sp - (p + 1);
        *desc = sp + 1;
        *desc_len = (p + n) - (sp + 1);
    } else {
        *name_len = n - 1;
        *desc = p + n;
        *desc_len = 0;
    }
    if (!next_line(r, seq, seq_len)) {
        *seq = p + n;
        *seq_len = 0;
    }
    if (!next_line(r, &p, &n)) return EXO_ERR_PARSE; /* unexpected EOF where the '+' line must be */
    if (n < 1 || p[0] != '+') return EXO_ERR_PARSE;
    if (!next_line(r, qual, qual_len)) {
        *qual = p + n;
        *qual_len = 0;
    }
    return 1;
}

/* BatchReader::read_batch (exon-fastq/src/batch_reader.rs:63-82) + FASTQArrayBuilder::append */
int exo_fastq_reader_next(exo_fastq_reader *r, exo_fastq_batch *out) {
    int64_t rows = 0;
    r->b.name.values_len = r->b.description.values_len = r->b.sequence.values_len = r->b.quality.values_len = 0;
    while (rows < r->batch_size) {
        const uint8_t *nm, *ds, *sq, *ql;
        int64_t nn, dn, sn, qn;
        int rc = read_record(r, &nm, &nn, &ds, &dn, &sq, &sn, &ql, &qn);
        if (rc == 0) break;
        if (rc < 0) {
            r->err_record = r->n_records;
            return rc;
        }
        col_append(&r->b.name, rows, nm, nn, 1);
        col_append(&r->b.description, rows, ds, dn, dn > 0); /* empty description -> NULL (array_builder.rs:76-83) */
        col_append(&r->b.sequence, rows, sq, sn, 1);
        col_append(&r->b.quality, rows, ql, qn, 1);
        rows++;
        r->n_records++;
    }
    if (rows == 0) return EXO_END;
    r->b.rows = rows;
    *out = r->b;
    return EXO_BATCH;
}

/* mean(quality) > num / den over one batch, column at a time: sum(byte - phred_offset) * den > num * len */
int64_t exo_fastq_mean_quality_count_batch(const exo_fastq_batch *b, int32_t phred_offset, int64_t num, int64_t den) {
    int64_t c = 0;
    for (int64_t i = 0; i < b->rows; i++) {
        const int32_t s = b->quality.offsets[i], e = b->quality.offsets[i + 1];
        int64_t sum = 0;
        for (int32_t j = s; j < e; j++) sum += (int64_t)b->quality.values[j] - phred_offset;
        c += (e > s) && (sum * den > num * (int64_t)(e - s));
    }
    return c;
}

/* whole file: rows and the filtered count; < 0 on a malformed record.  has_pred == 0 -> COUNT(*) */
int64_t exo_fastq_filter_count(const uint8_t *text, int64_t len, int64_t batch_size, int32_t has_pred, int32_t phred_offset,
                               int64_t num, int64_t den, int64_t *n_rows) {
    exo_fastq_reader *r = exo_fastq_reader_open(text, len, batch_size);
    exo_fastq_batch b;
    int64_t count = 0, rows = 0;
    int rc;
    while ((rc = exo_fastq_reader_next(r, &b)) == EXO_BATCH) {
        rows += b.rows;
        count += has_pred ? exo_fastq_mean_quality_count_batch(&b, phred_offset, num, den) : b.rows;
    }
    exo_fastq_reader_close(r);
    if (n_rows) *n_rows = rows;
    return rc < 0 ? rc : count;
}

/* one worker per file partition, as FASTQScan::execute gives DataFusion one stream per file group
 * (exon-core/src/datasources/fastq/scanner.rs:126); files are dealt round-robin in the given order */
typedef struct {
    const uint8_t *const *texts;
    const int64_t *lens;
    int32_t n_files, me, parts, has_pred, phred_offset;
    int64_t batch_size, num, den, count, rows;
    int err;
} fq_worker;

static void *fq_worker_main(void *p) {
    fq_worker *a = (fq_worker *)p;
    for (int32_t i = a->me; i < a->n_files; i += a->parts) {
        int64_t rows = 0;
        int64_t c = exo_fastq_filter_count(a->texts[i], a->lens[i], a->batch_size, a->has_pred, a->phred_offset, a->num, a->den, &rows);
        if (c < 0) {
            a->err = (int)c;
            return NULL;
        }
        a->count += c;
        a->rows += rows;
    }
    return NULL;
}

int64_t exo_fastq_filter_count_files(const uint8_t *const *texts, const int64_t *lens, int32_t n_files, int32_t target_partitions,
                                     int64_t batch_size, int32_t has_pred, int32_t phred_offset, int64_t num, int64_t den,
                                     int64_t *n_rows) {
    int32_t parts = n_files < target_partitions ? n_files : target_partitions;
    if (parts < 1) {
        if (n_rows) *n_rows = 0;
        return 0;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)parts);
    fq_worker *w = (fq_worker *)calloc((size_t)parts, sizeof(fq_worker));
    for (int32_t g = 0; g < parts; g++) {
        w[g].texts = texts; w[g].lens = lens; w[g].n_files = n_files; w[g].me = g; w[g].parts = parts;
        w[g].has_pred = has_pred; w[g].phred_offset = phred_offset; w[g].batch_size = batch_size; w[g].num = num; w[g].den = den;
        pthread_create(&th[g], NULL, fq_worker_main, &w[g]);
    }
    int64_t total = 0, rows = 0;
    int err = 0;
    for (int32_t g = 0; g < parts; g++) {
        pthread_join(th[g], NULL);
        if (w[g].err) err = w[g].err;
        total += w[g].count;
        rows += w[g].rows;
    }
    free(th);
    free(w);
    if (n_rows) *n_rows = rows;
    return err ? err : total;
}

/* quality_scores_to_list (exon-core/src/udfs/sequence/quality_score_string_to_list.rs:80-93): out[i] = byte - 33 */
void exo_quality_scores_to_list(const uint8_t *s, int64_t n, int32_t *out) {
    for (int64_t i = 0; i < n; i++) out[i] = (int32_t)s[i] - 33;
}

/* ---------------------------------------------------------------------------------------------
 * FASTA row count (BASELINE configs[0]).  noodles-fasta reader as driven by exon-fasta/src/batch_reader.rs: a record is
 * a definition line that starts with '>' plus the sequence lines up to the next definition; COUNT(*) is the number of
 * definition lines.  A non-empty input whose first byte is not '>' fails ("invalid definition").  Pinned by
 * slt/fasta-scan-tests.slt:72-85 (2 / 4 / gzip 2).
 * ------------------------------------------------------------------------------------------- */
int64_t exo_fasta_count(const uint8_t *text, int64_t len) {
    if (len == 0) return 0;
    if (text[0] != '>') return EXO_ERR_PARSE;
    int64_t n = 1;
    const uint8_t *p = text, *end = text + len;
    while ((p = (const uint8_t *)memchr(p, '\n', (size_t)(end - p))) != NULL) {
        ++p;
        if (p < end && *p == '>') ++n;
    }
    return n;
}

/* ---------------------------------------------------------------------------------------------
 * GFF record count with the reference's region filter.  exon-gff/src/batch_reader.rs:56-130: every line is a noodles-gff
 * Line -- "##..." directive, "#..." comment, otherwise a record of 9 tab-separated fields (fewer is an error, as is an
 * empty line); BatchReader::filter (:70-96) keeps a record when seqname == region name and, if the region has an
 * interval, the interval contains the record's START (1-based, inclusive).  has_name / has_interval select the terms.
 * Returns the selected count (*n_rows = all records), EXO_ERR_PARSE on a malformed record.
 * Pinned by slt/gff-scan-tests.slt:80-92 (5000 / 10000 / gzip 5000) and its first row (sq0 caat 8 13).
 * ------------------------------------------------------------------------------------------- */
int64_t exo_gff_filter_count(const uint8_t *text, int64_t len, const uint8_t *name, int32_t name_len, int32_t has_name,
                             int32_t has_interval, int64_t lo, int64_t hi, int64_t *n_rows) {
    int64_t p = 0, rows = 0, count = 0;
    while (p < len) {
        const uint8_t *nl = (const uint8_t *)memchr(text + p, '\n', (size_t)(len - p));
        const int64_t e = nl ? nl - text : len;
        const uint8_t *s = text + p;
        const int64_t n = e - p;
        p = e + 1;
        if (n == 0) return EXO_ERR_PARSE;
        if (s[0] == '#') continue;
        const uint8_t *f[10];
        int nf = 0;
        f[nf++] = s;
        for (int64_t i = 0; i < n && nf < 10; i++)
            if (s[i] == '\t') f[nf++] = s + i + 1;
        if (nf < 9) return EXO_ERR_PARSE;
        if (f[1] - 1 == f[0]) return EXO_ERR_PARSE; /* empty seqname */
        int64_t start = 0;
        const uint8_t *q = f[3];
        if (q >= f[4] - 1) return EXO_ERR_PARSE;
        for (; q < f[4] - 1; q++) {
            if (*q < '0' || *q > '9') return EXO_ERR_PARSE;
            start = start * 10 + (*q - '0');
        }
        if (start < 1) return EXO_ERR_PARSE;
        rows++;
        int sel = 1;
        if (has_name) sel = (f[1] - 1 - f[0]) == name_len && memcmp(f[0], name, (size_t)name_len) == 0;
        if (sel && has_interval) sel = start >= lo && start <= hi;
        count += sel;
    }
    if (n_rows) *n_rows = rows;
    return count;
}

/* ---------------------------------------------------------------------------------------------
 * FASTA records -> {id, description, sequence}.  exon-fasta/src/batch_reader.rs:52-103 (`read_definition` +
 * `read_sequence` of the noodles-fasta 0.4x reader, un-vendored) and FASTAArrayBuilder::append,
 * exon-fasta/src/array_builder.rs:108-160, through noodles `Definition::from_str`: '>' prefix, name = up to the first ASCII
 * whitespace (required), description = the rest of the line trimmed (None when the line has no whitespace); the sequence is
 * every following line up to the next definition, line terminators ("\n" / "\r\n") removed.  Errors: first line not a
 * definition, empty name, a definition with no sequence line ("invalid sequence", batch_reader.rs:63-65).
 * Pinned by slt/fasta-scan-tests.slt:6-10 (`a description ATCG`, `b description2 ATCG`).
 * Flat output, all malloc'ed: per record name_len / desc_len (-1 = NULL) / seq_len, and the three byte streams.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int64_t rows;
    int32_t *name_len, *desc_len;
    int64_t *seq_len;
    uint8_t *names, *descs, *seqs;
    int64_t names_len, descs_len, seqs_len;
    int32_t err;
} exo_fasta_records;

static int fa_ascii_ws(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == 0x0C || c == '\r'; }
static int fa_trim_ws(uint8_t c) { return fa_ascii_ws(c) || c == 0x0B; }

void exo_fasta_records_free(exo_fasta_records *r) {
    if (!r) return;
    free(r->name_len); free(r->desc_len); free(r->seq_len);
    free(r->names); free(r->descs); free(r->seqs);
    free(r);
}

exo_fasta_records *exo_fasta_read(const uint8_t *text, int64_t len) {
    exo_fasta_records *r = (exo_fasta_records *)calloc(1, sizeof(*r));
    int64_t cap = 0, ncap = 0, dcap = 0, scap = 0, p = 0;
    int in_record = 0, seq_lines = 0;
    while (p < len) {
        const uint8_t *nl = (const uint8_t *)memchr(text + p, '\n', (size_t)(len - p));
        const uint8_t *s = text + p, *e = nl ? nl : text + len;
        p = (nl ? nl - text : len) + 1;
        if (e > s && e[-1] == '\r') e--;
        if (e > s && s[0] == '>') {
            if (in_record && !seq_lines) { r->err = 3; return r; }
            if (r->rows == cap) {
                cap = cap ? cap * 2 : 1024;
                r->name_len = (int32_t *)realloc(r->name_len, sizeof(int32_t) * (size_t)cap);
                r->desc_len = (int32_t *)realloc(r->desc_len, sizeof(int32_t) * (size_t)cap);
                r->seq_len = (int64_t *)realloc(r->seq_len, sizeof(int64_t) * (size_t)cap);
            }
            const uint8_t *q = s + 1;
            while (q < e && !fa_ascii_ws(*q)) q++;
            const int32_t nlen = (int32_t)(q - (s + 1));
            if (nlen == 0) { r->err = 2; return r; }
            if (r->names_len + nlen > ncap) { while (r->names_len + nlen > ncap) ncap = ncap ? ncap * 2 : 4096; r->names = (uint8_t *)realloc(r->names, (size_t)ncap); }
            memcpy(r->names + r->names_len, s + 1, (size_t)nlen);
            r->names_len += nlen;
            r->name_len[r->rows] = nlen;
            r->desc_len[r->rows] = -1;
            if (q < e) {
                const uint8_t *a = q + 1, *b = e;
                while (a < b && fa_trim_ws(*a)) a++;
                while (b > a && fa_trim_ws(b[-1])) b--;
                const int32_t dlen = (int32_t)(b - a);
                if (r->descs_len + dlen > dcap) { while (r->descs_len + dlen > dcap) dcap = dcap ? dcap * 2 : 4096; r->descs = (uint8_t *)realloc(r->descs, (size_t)dcap); }
                if (dlen) memcpy(r->descs + r->descs_len, a, (size_t)dlen);
                r->descs_len += dlen;
                r->desc_len[r->rows] = dlen;
            }
            r->seq_len[r->rows] = 0;
            r->rows++;
            in_record = 1;
            seq_lines = 0;
        } else {
            if (!in_record) { r->err = 1; return r; }
            const int64_t n = e - s;
            if (r->seqs_len + n > scap) { while (r->seqs_len + n > scap) scap = scap ? scap * 2 : 65536; r->seqs = (uint8_t *)realloc(r->seqs, (size_t)scap); }
            if (n) memcpy(r->seqs + r->seqs_len, s, (size_t)n);
            r->seqs_len += n;
            r->seq_len[r->rows - 1] += n;
            seq_lines++;
        }
    }
    if (in_record && !seq_lines) r->err = 3;
    return r;
}

/* ---------------------------------------------------------------------------------------------
 * GFF records -> columns 0..7.  GFFArrayBuilder::append, exon-gff/src/array_builder.rs:84-150 over noodles-gff lazy records
 * (un-vendored): seqname / source / type as text; start / end decimal, non-zero; score "." -> NULL else f32 (Rust grammar is
 * checked by the caller's data; strtof here); strand "+" / "-" (anything else: error -- "." and "?" become NULL in a
 * non-nullable column and the reference's batch construction fails); phase "." -> NULL else "0" / "1" / "2".
 * Pinned by slt/gff-scan-tests.slt:6-10 (`sq0 caat 8 13 NULL + NULL`).  One record per call: fields of record `row`
 * (0-based, comments and directives skipped).  Returns 1, 0 when the row does not exist, < 0 on a malformed file.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    char seqname[256], source[256], type[256];
    int64_t start, end;
    float score;
    int32_t score_valid;
    char strand[4], phase[4]; /* phase "" = NULL */
} exo_gff_row;

int32_t exo_gff_row_at(const uint8_t *text, int64_t len, int64_t row, exo_gff_row *out) {
    int64_t p = 0, rows = 0;
    while (p < len) {
        const uint8_t *nl = (const uint8_t *)memchr(text + p, '\n', (size_t)(len - p));
        const int64_t e = nl ? nl - text : len;
        const uint8_t *s = text + p;
        const int64_t n = e - p;
        p = e + 1;
        if (n == 0) return EXO_ERR_PARSE;
        if (s[0] == '#') continue;
        const uint8_t *f[10];
        int nf = 0;
        f[nf++] = s;
        for (int64_t i = 0; i < n && nf < 10; i++)
            if (s[i] == '\t') f[nf++] = s + i + 1;
        if (nf < 9) return EXO_ERR_PARSE;
        if (nf == 9) f[9] = s + n + 1;
        if (rows++ != row) continue;
        memset(out, 0, sizeof(*out));
#define GFLD(k, dst) do { size_t l = (size_t)(f[(k) + 1] - f[(k)] - 1); if (l >= sizeof(dst)) return EXO_ERR_PARSE; memcpy(dst, f[(k)], l); dst[l] = 0; } while (0)
        GFLD(0, out->seqname);
        GFLD(1, out->source);
        GFLD(2, out->type);
        char num[64];
        GFLD(3, num);
        out->start = strtoll(num, NULL, 10);
        GFLD(4, num);
        out->end = strtoll(num, NULL, 10);
        if (out->start <= 0 || out->end <= 0) return EXO_ERR_PARSE;
        GFLD(5, num);
        out->score_valid = strcmp(num, ".") != 0;
        out->score = out->score_valid ? strtof(num, NULL) : 0.0f;
        GFLD(6, out->strand);
        if (strcmp(out->strand, "+") != 0 && strcmp(out->strand, "-") != 0) return EXO_ERR_PARSE;
        char ph[8];
        GFLD(7, ph);
        if (strcmp(ph, ".") == 0) out->phase[0] = 0;
        else if (!strcmp(ph, "0") || !strcmp(ph, "1") || !strcmp(ph, "2")) strcpy(out->phase, ph);
        else return EXO_ERR_PARSE;
#undef GFLD
        return 1;
    }
    return 0;
}
