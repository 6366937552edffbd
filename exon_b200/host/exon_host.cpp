// exon_host.cpp -- see exon_host.hpp.  Planning and formatting only: record bytes are never parsed here.
#include "exon_host.hpp"

#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <thread>

#include "../../include/exon_gpu.h"

namespace exon::host {

namespace {

[[noreturn]] void gpu_fail(int rc) {
    const std::string msg = exon_gpu_last_error();
    // the reference surfaces reader errors as ArrowError::ExternalError / DataFusionError::External
    throw ExonError(rc == EXON_GPU_ERR_PARSE ? ExonError::External : ExonError::Execution, msg);
}
inline void check(int rc) {
    if (rc != EXON_GPU_OK) gpu_fail(rc);
}

std::string lower(std::string s) {
    for (auto &c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}
bool ends_with(const std::string &s, const std::string &suf) {
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

// ---- files ------------------------------------------------------------------------------------------
std::vector<uint8_t> read_file(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw ExonError(ExonError::Execution, "Object at location " + path + " not found");
    f.seekg(0, std::ios::end);
    std::vector<uint8_t> buf((size_t)f.tellg());
    f.seekg(0);
    f.read((char *)buf.data(), (std::streamsize)buf.size());
    return buf;
}

void list_dir(const std::string &dir, std::vector<std::string> &out) {
    DIR *d = opendir(dir.c_str());
    if (!d) return;
    std::vector<std::string> names;
    while (dirent *e = readdir(d)) {
        const std::string n = e->d_name;
        if (n != "." && n != "..") names.push_back(n);
    }
    closedir(d);
    std::sort(names.begin(), names.end());
    for (const auto &n : names) {
        const std::string p = dir + "/" + n;
        struct stat st;
        if (stat(p.c_str(), &st) != 0) continue;
        if (S_ISDIR(st.st_mode)) list_dir(p, out);
        else out.push_back(p);
    }
}

// ---- a tiny SQL reader for the statement shapes of the reference's VCF sqllogictests -----------------------
struct Tok {
    enum T { Ident, Str, Num, Sym, End } t = End;
    std::string s;
};

struct Lexer {
    std::vector<Tok> toks;
    size_t i = 0;
    explicit Lexer(const std::string &q) {
        size_t p = 0;
        while (p < q.size()) {
            const char c = q[p];
            if (std::isspace((unsigned char)c)) { ++p; continue; }
            if (c == ';') { ++p; continue; }
            if (c == '\'') {
                std::string v;
                ++p;
                while (p < q.size() && q[p] != '\'') v += q[p++];
                if (p >= q.size()) throw ExonError(ExonError::Plan, "SQL error: unterminated string literal");
                ++p;
                toks.push_back({Tok::Str, v});
            } else if (std::isdigit((unsigned char)c)) {
                std::string v;
                while (p < q.size() && (std::isalnum((unsigned char)q[p]) || q[p] == '.')) v += q[p++];
                toks.push_back({Tok::Num, v});
            } else if (std::isalpha((unsigned char)c) || c == '_') {
                std::string v;
                while (p < q.size() && (std::isalnum((unsigned char)q[p]) || q[p] == '_' || q[p] == '.')) v += q[p++];
                toks.push_back({Tok::Ident, v});
            } else {
                std::string v(1, c);
                if ((c == '>' || c == '<' || c == '!') && p + 1 < q.size() && q[p + 1] == '=') v += q[++p];
                ++p;
                toks.push_back({Tok::Sym, v});
            }
        }
        toks.push_back({Tok::End, ""});
    }
    const Tok &peek() const { return toks[i]; }
    Tok next() { return toks[i < toks.size() - 1 ? i++ : i]; }
    bool kw(const char *k) {
        if (peek().t == Tok::Ident && lower(peek().s) == k) { ++i; return true; }
        return false;
    }
    bool sym(const char *k) {
        if (peek().t == Tok::Sym && peek().s == k) { ++i; return true; }
        return false;
    }
    void expect_kw(const char *k) {
        if (!kw(k)) throw ExonError(ExonError::Plan, std::string("SQL error: expected ") + k + " near '" + peek().s + "'");
    }
    void expect_sym(const char *k) {
        if (!sym(k)) throw ExonError(ExonError::Plan, std::string("SQL error: expected '") + k + "' near '" + peek().s + "'");
    }
    std::string ident() {
        if (peek().t != Tok::Ident) throw ExonError(ExonError::Plan, "SQL error: expected an identifier near '" + peek().s + "'");
        return next().s;
    }
    std::string str() {
        if (peek().t != Tok::Str) throw ExonError(ExonError::Plan, "SQL error: expected a string literal near '" + peek().s + "'");
        return next().s;
    }
};

Expr lit_or_col(Lexer &lx) {
    Expr e;
    const Tok t = lx.next();
    if (t.t == Tok::Str) { e.kind = Expr::Utf8; e.name = t.s; }
    else if (t.t == Tok::Num) {
        if (t.s.find_first_not_of("0123456789") != std::string::npos)
            throw ExonError(ExonError::NotImplemented, "only integer literals are supported in predicates on this path (got " + t.s + "; SURVEY 2.2 #7)");
        e.kind = Expr::Int64;
        e.i64 = std::stoll(t.s);
    } else if (t.t == Tok::Ident && (lower(t.s) == "true" || lower(t.s) == "false")) {
        e.kind = Expr::Boolean;
        e.i64 = lower(t.s) == "true";
    } else if (t.t == Tok::Ident) {
        if (lx.sym("(")) {
            e.kind = Expr::ScalarFunction;
            e.name = lower(t.s);
            if (!lx.sym(")")) {
                do e.args.push_back(lit_or_col(lx));
                while (lx.sym(","));
                lx.expect_sym(")");
            }
        } else {
            e.kind = Expr::Column;
            e.name = lower(t.s);
        }
    } else {
        throw ExonError(ExonError::Plan, "SQL error: unexpected token '" + t.s + "'");
    }
    return e;
}

// One conjunct: f(...) [= true] | a <op> b | a BETWEEN x AND y
Expr conjunct(Lexer &lx) {
    Expr l = lit_or_col(lx);
    if (lx.kw("between")) {
        Expr b;
        b.kind = Expr::Between;
        b.args.push_back(l);
        b.args.push_back(lit_or_col(lx));
        lx.expect_kw("and");
        b.args.push_back(lit_or_col(lx));
        return b;
    }
    for (const char *op : {"=", ">=", "<=", ">", "<"}) {
        if (lx.sym(op)) {
            Expr r = lit_or_col(lx);
            if (l.kind == Expr::ScalarFunction && std::string(op) == "=" && r.kind == Expr::Boolean && r.i64) return l;  // f(..) = true
            Expr b;
            b.kind = Expr::Binary;
            b.name = op;
            b.args = {l, r};
            return b;
        }
    }
    return l;
}

std::vector<Expr> where_clause(Lexer &lx) {  // split_conjunction of the WHERE expression
    std::vector<Expr> out;
    do out.push_back(conjunct(lx));
    while (lx.kw("and"));
    return out;
}

// Region + partition-column equalities a conjunction of filters amounts to; anything else is unsupported here.
struct Residual {
    bool has_chrom = false, has_interval = false, empty = false;
    std::string chrom;
    int64_t lo = 1, hi = INT64_MAX;
    std::vector<std::pair<std::string, std::string>> partition_eq;
    void and_chrom(const std::string &c) {
        if (has_chrom && chrom != c) empty = true;
        has_chrom = true;
        chrom = c;
    }
    void and_interval(int64_t a, int64_t b) {
        has_interval = true;
        lo = std::max(lo, a);
        hi = std::min(hi, b);
    }
};

bool is_col(const Expr &e, const char *n) { return e.kind == Expr::Column && e.name == n; }

Region region_of_udf(const Expr &f) {  // infer_region_from_udf: the first argument is the region literal
    if (f.args.empty() || f.args[0].kind != Expr::Utf8)
        throw ExonError(ExonError::Execution, "vcf_region_filter: the first argument must be a region string");
    return Region::parse(f.args[0].name);
}

void fold_filter(const Expr &f, const std::vector<std::string> &partition_cols, Residual &r) {
    auto is_part = [&](const Expr &e) {
        return e.kind == Expr::Column && std::find(partition_cols.begin(), partition_cols.end(), e.name) != partition_cols.end();
    };
    if (f.kind == Expr::ScalarFunction && f.name == "vcf_region_filter") {
        if (f.args.size() != 2 && f.args.size() != 3) throw ExonError(ExonError::Plan, "vcf_region_filter takes 2 or 3 arguments");
        const Region g = region_of_udf(f);
        r.and_chrom(g.name);
        if (g.has_interval) r.and_interval(g.lo, g.hi);
        return;
    }
    if (f.kind == Expr::ScalarFunction && (f.name == "region_match" || f.name == "chrom_match" || f.name == "interval_match")) {
        const size_t want = f.name == "region_match" ? 3 : 2;
        if (f.args.size() != want) throw ExonError(ExonError::Plan, f.name + ": wrong number of arguments");
        if (f.args.back().kind != Expr::Utf8) throw ExonError(ExonError::Execution, "Failed to get region");
        if (f.name == "chrom_match") { r.and_chrom(f.args[1].name); return; }
        if (f.name == "interval_match") {
            exon_gpu_region g;
            check(exon_gpu_interval_parse(f.args[1].name.c_str(), &g));
            r.and_interval(g.lo, g.hi);
            return;
        }
        const Region g = Region::parse(f.args[2].name);
        r.and_chrom(g.name);
        r.and_interval(g.has_interval ? g.lo : 1, g.has_interval ? g.hi : INT64_MAX);
        return;
    }
    if (f.kind == Expr::Between && is_col(f.args[0], "pos") && f.args[1].kind == Expr::Int64 && f.args[2].kind == Expr::Int64) {
        r.and_interval(f.args[1].i64, f.args[2].i64);
        return;
    }
    if (f.kind == Expr::Binary) {
        const Expr &a = f.args[0], &b = f.args[1];
        if (f.name == "=" && is_col(a, "chrom") && b.kind == Expr::Utf8) { r.and_chrom(b.name); return; }
        if (f.name == "=" && is_part(a) && b.kind == Expr::Utf8) { r.partition_eq.push_back({a.name, b.name}); return; }
        if (is_col(a, "pos") && b.kind == Expr::Int64) {
            if (f.name == "=") { r.and_interval(b.i64, b.i64); return; }
            if (f.name == ">=") { r.and_interval(b.i64, INT64_MAX); return; }
            if (f.name == "<=") { r.and_interval(1, b.i64); return; }
            if (f.name == ">") { r.and_interval(b.i64 == INT64_MAX ? INT64_MAX : b.i64 + 1, INT64_MAX); if (b.i64 == INT64_MAX) r.empty = true; return; }
            if (f.name == "<") { r.and_interval(1, b.i64 - 1); return; }
        }
    }
    throw ExonError(ExonError::NotImplemented, "predicate not supported on the GPU VCF path: " + f.to_string());
}

}  // namespace

// ---- Expr / Region ------------------------------------------------------------------------------------
std::string Expr::to_string() const {
    switch (kind) {
        case Column: return name;
        case Utf8: return "'" + name + "'";
        case Int64: return std::to_string(i64);
        case Boolean: return i64 ? "true" : "false";
        case Binary: return args[0].to_string() + " " + name + " " + args[1].to_string();
        case Between: return args[0].to_string() + " BETWEEN " + args[1].to_string() + " AND " + args[2].to_string();
        case ScalarFunction: {
            std::string s = name + "(";
            for (size_t i = 0; i < args.size(); ++i) s += (i ? ", " : "") + args[i].to_string();
            return s + ")";
        }
    }
    return "?";
}

Region Region::parse(const std::string &s) {
    char buf[256];
    exon_gpu_region g;
    if (exon_gpu_region_parse(s.c_str(), buf, sizeof(buf), &g) != EXON_GPU_OK)
        throw ExonError(ExonError::Execution, std::string("Failed to parse region: ") + exon_gpu_last_error());
    Region r;
    r.name.assign(buf, (size_t)g.chrom_len);
    r.has_interval = g.has_interval != 0;
    r.lo = g.lo;
    r.hi = g.hi;
    return r;
}

ListingVCFTableOptions ListingVCFTableOptions::make(FileCompressionType c, bool indexed) {
    ListingVCFTableOptions o;
    o.file_compression_type = c;
    o.indexed = indexed;
    // ExonFileType::get_file_extension: "vcf" + ".gz" under GZIP (exon_file_type.rs)
    o.file_extension = c == FileCompressionType::GZIP ? "vcf.gz" : "vcf";
    return o;
}

// ---- TableProvider ------------------------------------------------------------------------------------
std::vector<PartitionedFile> ListingVCFTable::list_files() const {
    std::vector<std::string> paths;
    struct stat st;
    if (stat(table_path.c_str(), &st) != 0) throw ExonError(ExonError::Execution, "Object at location " + table_path + " not found");
    if (S_ISDIR(st.st_mode)) list_dir(table_path, paths);
    else paths.push_back(table_path);
    std::vector<PartitionedFile> out;
    for (const auto &p : paths) {
        if (S_ISDIR(st.st_mode) && !ends_with(p, "." + options.file_extension)) continue;
        PartitionedFile f;
        f.path = p;
        struct stat fs;
        if (stat(p.c_str(), &fs) == 0) f.size = (int64_t)fs.st_size;
        // hive partition values: path components `col=value` below the table path
        for (const auto &col : options.table_partition_cols) {
            const std::string key = "/" + col + "=";
            const size_t at = p.find(key, table_path.size() ? table_path.size() - 1 : 0);
            std::string v;
            if (at != std::string::npos) {
                const size_t b = at + key.size();
                v = p.substr(b, p.find('/', b) - b);
            }
            f.partition_values.push_back(v);
        }
        out.push_back(f);
    }
    return out;
}

std::vector<TableProviderFilterPushDown> ListingVCFTable::supports_filters_pushdown(const std::vector<Expr> &filters) const {
    std::vector<TableProviderFilterPushDown> out;
    for (const auto &f : filters) {
        if (f.kind == Expr::ScalarFunction && f.name == "vcf_region_filter" && (f.args.size() == 2 || f.args.size() == 3)) {
            out.push_back(TableProviderFilterPushDown::Exact);
            continue;
        }
        // filter_matches_partition_cols (hive_partition.rs:32-53): `partition_col = literal` is Exact
        bool part = false;
        if (f.kind == Expr::Binary && f.name == "=" && f.args[0].kind == Expr::Column && f.args[1].kind == Expr::Utf8)
            part = std::find(options.table_partition_cols.begin(), options.table_partition_cols.end(), f.args[0].name) !=
                   options.table_partition_cols.end();
        out.push_back(part ? TableProviderFilterPushDown::Exact : TableProviderFilterPushDown::Unsupported);
    }
    return out;
}

std::shared_ptr<VCFScan> ListingVCFTable::scan(const std::vector<int> *projection, const std::vector<Expr> &filters,
                                               const int64_t *) const {
    std::vector<Region> regions;
    for (const auto &f : filters)
        if (f.kind == Expr::ScalarFunction && f.name == "vcf_region_filter") regions.push_back(region_of_udf(f));
    regions.insert(regions.end(), options.regions.begin(), options.regions.end());
    if (regions.size() > 1) throw ExonError(ExonError::NotImplemented, "Multiple regions are not supported yet");
    if (regions.empty() && options.indexed)
        throw ExonError(ExonError::Plan, "INDEXED_VCF table requires a region filter. See the UDF 'vcf_region_filter'.");
    // pruned_partition_list: drop files whose hive partition values contradict `col = literal` filters
    std::vector<PartitionedFile> files;
    for (auto &f : list_files()) {
        bool keep = true;
        for (const auto &e : filters) {
            if (e.kind != Expr::Binary || e.name != "=" || e.args[0].kind != Expr::Column || e.args[1].kind != Expr::Utf8) continue;
            for (size_t c = 0; c < options.table_partition_cols.size(); ++c)
                if (options.table_partition_cols[c] == e.args[0].name && f.partition_values[c] != e.args[1].name) keep = false;
        }
        if (keep) files.push_back(f);
    }
    auto plan = std::make_shared<VCFScan>();
    plan->file_groups.push_back(files);  // one group; `repartitioned` splits it
    if (projection) plan->projection = *projection;
    else plan->projection = {0, 1};
    plan->compression = options.file_compression_type;
    if (!regions.empty()) {
        plan->has_region = true;
        plan->region = regions[0];
    }
    return plan;
}

std::shared_ptr<VCFScan> VCFScan::repartitioned(int target_partitions) const {
    std::vector<PartitionedFile> flat;
    for (const auto &g : file_groups) flat.insert(flat.end(), g.begin(), g.end());
    auto out = std::make_shared<VCFScan>(*this);
    out->file_groups.clear();
    if (flat.empty() || target_partitions <= 1) {
        out->file_groups.push_back(flat);
        return out;
    }
    std::vector<int64_t> sizes;
    for (const auto &f : flat) sizes.push_back(f.size);
    std::vector<int32_t> part(flat.size());
    int32_t n_parts = 0;
    check(exon_gpu_regroup_files_by_size(sizes.data(), (int32_t)flat.size(), target_partitions, part.data(), &n_parts));
    out->file_groups.resize((size_t)n_parts);
    for (size_t i = 0; i < flat.size(); ++i) out->file_groups[(size_t)part[i]].push_back(flat[i]);
    return out;
}

// ---- execution -----------------------------------------------------------------------------------------
struct Executor {
    ExonSession &s;
    explicit Executor(ExonSession &session) : s(session) {}

    // FileStream + VCFOpener::open / IndexedVCFOpener::open for one partition: every file of the group is fed into one
    // stream handle.  Compressed files are inflated on the device (exon_gpu_stream_feed_gzip); an indexed scan with a
    // .tbi next to the file asks the index for the region's chunks and feeds only those (indexed_bgzf_file.rs:52-83,
    // indexed_file_opener.rs:53-214), otherwise the whole file is scanned with the same region predicate.
    void feed_group(exon_gpu_stream *st, const VCFScan &plan, const std::vector<PartitionedFile> &group) {
        for (const auto &f : group) {
            std::vector<uint8_t> bytes = read_file(f.path);
            bool fed = false;
            if (plan.compression == FileCompressionType::GZIP && plan.has_region) {
                struct stat ts;
                const std::string tbi = f.path + ".tbi";
                if (stat(tbi.c_str(), &ts) == 0) {
                    const std::vector<uint8_t> index = read_file(tbi);
                    exon_gpu_region rg;
                    memset(&rg, 0, sizeof(rg));
                    rg.chrom = plan.region.name.c_str();
                    rg.chrom_len = (int32_t)plan.region.name.size();
                    rg.has_chrom = 1;
                    rg.has_interval = plan.region.has_interval;
                    rg.lo = plan.region.lo;
                    rg.hi = plan.region.hi;
                    int32_t n = 0;
                    check(exon_gpu_tabix_query(s.ctx_, index.data(), index.size(), &rg, nullptr, 0, &n));
                    std::vector<exon_gpu_chunk> chunks((size_t)std::max(n, 1));
                    check(exon_gpu_tabix_query(s.ctx_, index.data(), index.size(), &rg, chunks.data(), n, &n));
                    for (int32_t i = 0; i < n; ++i) {  // a ranged GET from the chunk's first member to the end of the object
                        const uint64_t lo = chunks[(size_t)i].start >> 16;
                        if (lo >= bytes.size()) throw ExonError(ExonError::External, "tabix chunk beyond the end of " + f.path);
                        check(exon_gpu_stream_feed_bgzf_chunk(st, bytes.data() + lo, bytes.size() - (size_t)lo, lo, &chunks[(size_t)i]));
                    }
                    fed = true;
                }
            }
            if (!fed) {
                if (plan.compression == FileCompressionType::GZIP) check(exon_gpu_stream_feed_gzip(st, bytes.data(), bytes.size(), 1));
                else check(exon_gpu_vcf_feed(st, bytes.data(), bytes.size(), 0, 1));
            }
            int64_t flushed = 0;
            check(exon_gpu_stream_body_bytes(st, &flushed));  // flushes pending inflates: `bytes` goes out of scope
            check(exon_gpu_ctx_synchronize(s.ctx_));
        }
    }

    // COUNT(*) of the other formats' table functions (fastq_scan, bam_scan, mzml_scan): every file under `path`.
    int64_t count_other(const std::string &fn, const std::string &path, bool gz) {
        std::vector<std::string> paths;
        struct stat stt;
        if (stat(path.c_str(), &stt) != 0) throw ExonError(ExonError::Execution, "Object at location " + path + " not found");
        if (S_ISDIR(stt.st_mode)) list_dir(path, paths);
        else paths.push_back(path);
        std::sort(paths.begin(), paths.end());
        exon_gpu_stream *st = nullptr;
        if (fn == "fastq_scan") check(exon_gpu_fastq_open(s.ctx_, nullptr, &st));
        else if (fn == "bam_scan") check(exon_gpu_bam_open(s.ctx_, &st));
        else if (fn == "fasta_scan") check(exon_gpu_fasta_open(s.ctx_, &st));
        else if (fn == "gff_scan") check(exon_gpu_gff_open(s.ctx_, &st));
        else check(exon_gpu_mzml_open(s.ctx_, &st));
        int64_t n = 0;
        try {
            for (const auto &p : paths) {
                const std::vector<uint8_t> bytes = read_file(p);
                const bool file_gz = gz || ends_with(p, ".gz");
                if (fn == "bam_scan") check(exon_gpu_bam_feed(st, bytes.data(), bytes.size(), 1));
                else if (file_gz) check(exon_gpu_stream_feed_gzip(st, bytes.data(), bytes.size(), 1));
                else if (fn == "fastq_scan") check(exon_gpu_fastq_feed(st, bytes.data(), bytes.size(), 0, 1));
                else if (fn == "fasta_scan") check(exon_gpu_fasta_feed(st, bytes.data(), bytes.size(), 0, 1));
                else if (fn == "gff_scan") check(exon_gpu_gff_feed(st, bytes.data(), bytes.size(), 0, 1));
                else check(exon_gpu_mzml_feed(st, bytes.data(), bytes.size(), 0, 1));
                int64_t flushed = 0;
                check(exon_gpu_stream_body_bytes(st, &flushed));
                check(exon_gpu_ctx_synchronize(s.ctx_));
            }
            if (fn == "fastq_scan") {
                check(exon_gpu_fastq_rows(st, &n));
            } else if (fn == "fasta_scan") {
                check(exon_gpu_fasta_rows(st, &n));
            } else if (fn == "gff_scan") {
                check(exon_gpu_gff_filter_count(st, nullptr, &n));
            } else if (fn == "bam_scan") {
                int32_t groups = 0;
                check(exon_gpu_bam_filter_count_by_reference(st, nullptr, nullptr, 0, &groups, &n));
            } else {
                double sum = 0;
                int64_t sel = 0;
                check(exon_gpu_mzml_filter_sum(st, nullptr, &sum, &sel, &n));
            }
        } catch (...) {
            exon_gpu_stream_close(st);
            throw;
        }
        exon_gpu_stream_close(st);
        return n;
    }

    int64_t count(const VCFScan &plan, const Residual &r) {
        exon_gpu_region reg;
        memset(&reg, 0, sizeof(reg));
        reg.chrom = r.chrom.c_str();
        reg.chrom_len = (int32_t)r.chrom.size();
        reg.has_chrom = r.has_chrom;
        reg.has_interval = r.has_interval;
        reg.lo = r.lo;
        reg.hi = r.hi;
        const bool any_pred = r.has_chrom || r.has_interval;
        int64_t total = 0;
        exon_gpu_partial *d_acc = nullptr;
        if (!s.config.gpu_fused) {
            check(exon_gpu_device_alloc(s.ctx_, 64, (void **)&d_acc));
            check(exon_gpu_memset(s.ctx_, d_acc, 0, 64));
        }
        for (const auto &group : plan.file_groups) {  // one partition stream per file group
            exon_gpu_vcf_opts o;
            memset(&o, 0, sizeof(o));
            o.batch_rows = s.config.batch_size;
            std::vector<int32_t> proj;
            if (r.has_chrom) proj.push_back(0);
            if (r.has_interval) proj.push_back(1);
            o.projection = proj.data();
            o.n_projection = (int32_t)proj.size();
            o.columns_on_device = 1;
            o.strict = s.config.gpu_strict;
            o.pushdown = (s.config.gpu_fused && any_pred) ? &reg : nullptr;
            exon_gpu_stream *st = nullptr;
            check(exon_gpu_vcf_open(s.ctx_, &o, &st));
            try {
                feed_group(st, plan, group);
                if (s.config.gpu_fused) {
                    int64_t n = 0;
                    check(exon_gpu_vcf_filter_count(st, any_pred ? &reg : nullptr, &n));
                    total += r.empty ? 0 : n;
                } else {  // VCFScan batches -> FilterExec + AggregateExec(Partial) on the columns
                    exon_gpu_pred p;
                    memset(&p, 0, sizeof(p));
                    p.chrom_col = r.has_chrom ? 0 : -1;
                    p.pos_col = r.has_interval ? (r.has_chrom ? 1 : 0) : -1;
                    p.region = reg;
                    exon_gpu_agg a{EXON_GPU_AGG_COUNT_STAR, -1};
                    for (;;) {
                        ArrowArray arr;
                        ArrowSchema sch;
                        check(exon_gpu_vcf_next_batch(st, &arr, &sch));
                        if (!arr.release) {
                            if (sch.release) sch.release(&sch);
                            break;
                        }
                        const int rc = exon_gpu_filter_agg_accumulate(s.ctx_, &arr, &sch, &p, &a, d_acc);
                        arr.release(&arr);
                        sch.release(&sch);
                        check(rc);
                    }
                }
            } catch (...) {
                exon_gpu_vcf_close(st);
                if (d_acc) exon_gpu_device_free(s.ctx_, d_acc);
                throw;
            }
            check(exon_gpu_vcf_close(st));
        }
        if (d_acc) {
            exon_gpu_partial out;
            check(exon_gpu_partial_read(s.ctx_, d_acc, 1, &out));
            check(exon_gpu_device_free(s.ctx_, d_acc));
            total = r.empty ? 0 : out.count;
        }
        return total;
    }

    std::vector<RecordBatch> collect(const VCFScan &plan, const Residual &r, int64_t limit) {
        std::vector<RecordBatch> out;
        int64_t taken = 0;
        exon_gpu_region reg;
        memset(&reg, 0, sizeof(reg));
        reg.chrom = r.chrom.c_str();
        reg.chrom_len = (int32_t)r.chrom.size();
        reg.has_chrom = r.has_chrom;
        reg.has_interval = r.has_interval;
        reg.lo = r.lo;
        reg.hi = r.hi;
        for (const auto &group : plan.file_groups) {
            exon_gpu_vcf_opts o;
            memset(&o, 0, sizeof(o));
            o.batch_rows = s.config.batch_size;
            const int32_t proj[2] = {0, 1};
            o.projection = proj;
            o.n_projection = 2;
            o.strict = s.config.gpu_strict;
            exon_gpu_stream *st = nullptr;
            check(exon_gpu_vcf_open(s.ctx_, &o, &st));
            try {
                feed_group(st, plan, group);
                while (limit < 0 || taken < limit) {
                    ArrowArray arr;
                    ArrowSchema sch;
                    check(exon_gpu_vcf_next_batch(st, &arr, &sch));
                    if (!arr.release) {
                        if (sch.release) sch.release(&sch);
                        break;
                    }
                    const int64_t n = arr.length;
                    std::vector<uint8_t> keep((size_t)std::max<int64_t>(n, 1), 1);
                    int rc = EXON_GPU_OK;
                    if ((r.has_chrom || r.has_interval) && !r.empty) {  // FilterExec: the mask comes from the GPU
                        exon_gpu_pred p;
                        memset(&p, 0, sizeof(p));
                        p.chrom_col = 0;
                        p.pos_col = 1;
                        p.region = reg;
                        const int kind = r.has_chrom ? (r.has_interval ? EXON_GPU_UDF_REGION_MATCH : EXON_GPU_UDF_CHROM_MATCH)
                                                     : EXON_GPU_UDF_INTERVAL_MATCH;
                        rc = exon_gpu_region_udf(s.ctx_, kind, &arr, &sch, 0, &p, keep.data(), nullptr);
                    } else if (r.empty) {
                        std::fill(keep.begin(), keep.end(), 0);
                    }
                    RecordBatch b;
                    if (rc == EXON_GPU_OK) {
                        const auto *off = (const int32_t *)arr.children[0]->buffers[1];
                        const auto *val = (const char *)arr.children[0]->buffers[2];
                        const auto *pos = (const int64_t *)arr.children[1]->buffers[1];
                        for (int64_t i = 0; i < n && (limit < 0 || taken < limit); ++i) {
                            if (!keep[(size_t)i]) continue;
                            b.chrom.emplace_back(val + off[i], (size_t)(off[i + 1] - off[i]));
                            b.pos.push_back(pos[i]);
                            ++taken;
                        }
                        b.num_rows = (int64_t)b.pos.size();
                    }
                    arr.release(&arr);
                    sch.release(&sch);
                    check(rc);
                    if (b.num_rows) out.push_back(std::move(b));
                }
            } catch (...) {
                exon_gpu_vcf_close(st);
                throw;
            }
            check(exon_gpu_vcf_close(st));
        }
        return out;
    }
};

// ---- session -------------------------------------------------------------------------------------------
std::unique_ptr<ExonSession> ExonSession::new_exon(int device) {
    std::unique_ptr<ExonSession> s(new ExonSession());
    check(exon_gpu_ctx_create(device, nullptr, &s->ctx_));
    return s;
}

ExonSession::~ExonSession() {
    if (ctx_) exon_gpu_ctx_destroy(ctx_);
}

int64_t ExonSession::gpu_launches() const {
    int64_t n = 0;
    exon_gpu_ctx_launch_count(ctx_, &n);
    return n;
}

std::shared_ptr<ListingVCFTable> ExonSession::read_vcf(const std::string &table_path, const ListingVCFTableOptions &options) {
    auto t = std::make_shared<ListingVCFTable>();
    t->table_path = table_path;
    t->options = options;
    return t;
}

static Residual residual_of(const ListingVCFTable &table, const std::vector<Expr> &filters) {
    Residual r;
    for (const auto &f : filters) fold_filter(f, table.options.table_partition_cols, r);
    if (r.has_interval && r.lo > r.hi) r.empty = true;
    return r;
}

static int effective_partitions(const SessionConfig &c) {
    if (c.target_partitions > 0) return c.target_partitions;
    const unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

int64_t ExonSession::count(const ListingVCFTable &table, const std::vector<Expr> &filters) {
    const std::vector<int> projection;  // COUNT(*): empty projection; predicates add their columns
    auto plan = table.scan(&projection, filters, nullptr)->repartitioned(effective_partitions(config));
    Residual r = residual_of(table, filters);
    return Executor(*this).count(*plan, r);
}

std::vector<RecordBatch> ExonSession::collect(const ListingVCFTable &table, const std::vector<Expr> &filters, int64_t limit) {
    auto plan = table.scan(nullptr, filters, limit >= 0 ? &limit : nullptr)->repartitioned(effective_partitions(config));
    Residual r = residual_of(table, filters);
    return Executor(*this).collect(*plan, r, limit);
}

std::string ResultSet::to_text() const {
    std::string out;
    for (const auto &row : rows) {
        for (size_t i = 0; i < row.size(); ++i) out += (i ? " " : "") + row[i];
        out += "\n";
    }
    return out;
}

ResultSet ExonSession::sql(const std::string &query) {
    Lexer lx(query);
    ResultSet rs;
    if (lx.kw("set")) {
        const std::string key = lower(lx.ident());
        lx.expect_sym("=");
        const Tok v = lx.next();
        config.options[key] = v.s;
        const bool on = lower(v.s) == "true" || v.s == "1";
        if (key == "exon.gpu_fused") config.gpu_fused = on;
        else if (key == "exon.gpu_strict") config.gpu_strict = on;
        else if (key == "datafusion.execution.batch_size") config.batch_size = std::stoi(v.s);
        else if (key == "datafusion.execution.target_partitions") config.target_partitions = std::stoi(v.s);
        return rs;
    }
    if (lx.kw("drop")) {
        lx.expect_kw("table");
        const std::string name = lower(lx.ident());
        if (!tables_.erase(name)) throw ExonError(ExonError::Plan, "Table '" + name + "' doesn't exist.");
        return rs;
    }
    if (lx.kw("create")) {
        lx.expect_kw("external");
        lx.expect_kw("table");
        const std::string name = lower(lx.ident());
        lx.expect_kw("stored");
        lx.expect_kw("as");
        const std::string type = lower(lx.ident());
        if (type != "vcf" && type != "indexed_vcf")
            throw ExonError(ExonError::NotImplemented, "STORED AS " + type + " is outside the GPU VCF path");
        std::vector<std::string> part_cols;
        std::string location;
        FileCompressionType comp = FileCompressionType::UNCOMPRESSED;
        for (;;) {
            if (lx.kw("partitioned")) {
                lx.expect_kw("by");
                lx.expect_sym("(");
                do part_cols.push_back(lower(lx.ident()));
                while (lx.sym(","));
                lx.expect_sym(")");
            } else if (lx.kw("location")) {
                location = lx.str();
            } else if (lx.kw("compression")) {
                lx.expect_kw("type");
                if (lower(lx.ident()) == "gzip") comp = FileCompressionType::GZIP;
            } else if (lx.kw("options")) {
                lx.expect_sym("(");
                do {
                    const std::string k = lower(lx.peek().t == Tok::Str ? lx.str() : lx.ident());
                    const Tok v = lx.next();
                    if (k == "compression" && lower(v.s) == "gzip") comp = FileCompressionType::GZIP;
                } while (lx.sym(","));
                lx.expect_sym(")");
            } else {
                break;
            }
        }
        if (location.empty()) throw ExonError(ExonError::Plan, "CREATE EXTERNAL TABLE needs a LOCATION");
        auto opts = ListingVCFTableOptions::make(comp, type == "indexed_vcf");
        opts.table_partition_cols = part_cols;
        tables_[name] = read_vcf(location, opts);
        return rs;
    }
    lx.expect_kw("select");
    bool count_star = false, star = false;
    std::vector<std::string> cols;
    do {
        if (lx.sym("*")) { star = true; continue; }
        const std::string id = lower(lx.ident());
        if (id == "count" && lx.sym("(")) {
            lx.expect_sym("*");
            lx.expect_sym(")");
            count_star = true;
            std::string alias = "count(*)";
            if (lx.kw("as")) alias = lx.ident();
            rs.columns.push_back(alias);
        } else {
            cols.push_back(id);
            rs.columns.push_back(id);
        }
    } while (lx.sym(","));
    lx.expect_kw("from");
    const std::string src = lower(lx.ident());
    std::shared_ptr<ListingVCFTable> table;
    std::vector<Expr> filters;
    if (lx.sym("(")) {  // table function: ScanFunction::try_from
        std::vector<Expr> args;
        if (!lx.sym(")")) {
            do args.push_back(lit_or_col(lx));
            while (lx.sym(","));
            lx.expect_sym(")");
        }
        if (src == "fastq_scan" || src == "bam_scan" || src == "mzml_scan" || src == "fasta_scan" || src == "gff_scan") {
            // FASTQ / BAM / mzML ScanFunctions (fastq/udtf.rs:49, bam/udtf.rs:55, mzml/udtf.rs:47): COUNT(*) runs on the GPU
            if (args.empty() || args[0].kind != Expr::Utf8)
                throw ExonError(ExonError::Internal, "this function requires the path to be specified as the first argument");
            if (!count_star || !cols.empty() || star || lx.peek().t != Tok::End)
                throw ExonError(ExonError::NotImplemented, src + ": only SELECT COUNT(*) runs on the GPU path of this round");
            const bool gz = args.size() > 1 && args[1].kind == Expr::Utf8 && lower(args[1].name) == "gzip";
            Executor ex(*this);
            rs.rows.push_back({std::to_string(ex.count_other(src, args[0].name, gz))});
            return rs;
        }
        if (src != "vcf_scan" && src != "vcf_indexed_scan")
            throw ExonError(ExonError::NotImplemented, "table function " + src + " is outside the GPU path");
        if (args.empty() || args[0].kind != Expr::Utf8)
            throw ExonError(ExonError::Internal, "this function requires the path to be specified as the first argument");
        FileCompressionType comp = ends_with(args[0].name, ".gz") ? FileCompressionType::GZIP : FileCompressionType::UNCOMPRESSED;
        if (src == "vcf_scan" && args.size() > 1 && args[1].kind == Expr::Utf8 && lower(args[1].name) == "gzip") comp = FileCompressionType::GZIP;
        if (src == "vcf_indexed_scan") {
            if (args.size() < 2 || args[1].kind != Expr::Utf8)
                throw ExonError(ExonError::Internal, "this function requires the region to be specified as the second argument");
            comp = FileCompressionType::GZIP;  // indexed files are BGZF (vcf/udtf.rs:106-145)
            auto opts = ListingVCFTableOptions::make(comp, true);
            opts.regions.push_back(Region::parse(args[1].name));
            table = read_vcf(args[0].name, opts);
            Expr f;  // the configured region is also the residual predicate of the indexed scan
            f.kind = Expr::ScalarFunction;
            f.name = "region_match";
            Expr c, p, l;
            c.kind = Expr::Column; c.name = "chrom";
            p.kind = Expr::Column; p.name = "pos";
            l.kind = Expr::Utf8; l.name = args[1].name;
            f.args = {c, p, l};
            filters.push_back(f);
        } else {
            table = read_vcf(args[0].name, ListingVCFTableOptions::make(comp, false));
        }
    } else {
        auto it = tables_.find(src);
        if (it == tables_.end()) throw ExonError(ExonError::Plan, "table '" + src + "' not found");
        table = it->second;
    }
    if (lx.kw("where")) {
        auto w = where_clause(lx);
        filters.insert(filters.end(), w.begin(), w.end());
    }
    int64_t limit = -1;
    if (lx.kw("group")) throw ExonError(ExonError::NotImplemented, "GROUP BY is outside the GPU VCF path of this round");
    if (lx.kw("limit")) limit = std::stoll(lx.next().s);
    if (lx.peek().t != Tok::End) throw ExonError(ExonError::Plan, "SQL error: unexpected '" + lx.peek().s + "'");

    if (count_star) {
        if (!cols.empty() || star) throw ExonError(ExonError::NotImplemented, "COUNT(*) mixed with columns needs GROUP BY");
        rs.rows.push_back({std::to_string(count(*table, filters))});
        return rs;
    }
    if (star) {
        cols = {"chrom", "pos"};  // the columns this round builds on the GPU
        rs.columns = cols;
    }
    for (const auto &c : cols)
        if (c != "chrom" && c != "pos")
            throw ExonError(ExonError::NotImplemented, "column " + c + " is not built on the GPU yet (supported: chrom, pos)");
    for (const auto &b : collect(*table, filters, limit))
        for (int64_t i = 0; i < b.num_rows; ++i) {
            std::vector<std::string> row;
            for (const auto &c : cols) row.push_back(c == "chrom" ? b.chrom[(size_t)i] : std::to_string(b.pos[(size_t)i]));
            rs.rows.push_back(row);
        }
    return rs;
}

}  // namespace exon::host

// ---- C shim for the Python tests (ctypes) -------------------------------------------------------------------
using exon::host::ExonError;
using exon::host::ExonSession;

static thread_local std::string g_host_error;
static thread_local std::string g_host_text;

extern "C" {

void *exon_host_session_new(int device) {
    try {
        return ExonSession::new_exon(device).release();
    } catch (const std::exception &e) {
        g_host_error = e.what();
        return nullptr;
    }
}
void exon_host_session_free(void *s) { delete static_cast<ExonSession *>(s); }
const char *exon_host_last_error(void) { return g_host_error.c_str(); }

// 0 = ok (*out_text -> rows, one per line), 1 = Plan, 2 = Execution, 3 = NotImplemented, 4 = External, 5 = Internal
int exon_host_sql(void *s, const char *query, const char **out_text) {
    try {
        g_host_text = static_cast<ExonSession *>(s)->sql(query).to_text();
        if (out_text) *out_text = g_host_text.c_str();
        return 0;
    } catch (const ExonError &e) {
        g_host_error = e.what();
        return 1 + (int)e.kind;
    } catch (const std::exception &e) {
        g_host_error = e.what();
        return 2;
    }
}

// supports_filters_pushdown of a registered table for the conjuncts of `where_sql`: writes one char per filter
// ('E'xact / 'U'nsupported / 'I'nexact) into buf.
int exon_host_pushdown(void *s, const char *table_sql_source, const char *where_sql, char *buf, int buf_len) {
    try {
        auto *sess = static_cast<ExonSession *>(s);
        (void)sess;
        exon::host::ListingVCFTable t;
        std::string src = table_sql_source;
        // "col1,col2" = partition columns of the probe table
        size_t p = 0;
        while (p < src.size()) {
            size_t q = src.find(',', p);
            if (q == std::string::npos) q = src.size();
            if (q > p) t.options.table_partition_cols.push_back(src.substr(p, q - p));
            p = q + 1;
        }
        exon::host::Lexer lx(where_sql);
        auto filters = exon::host::where_clause(lx);
        auto r = t.supports_filters_pushdown(filters);
        if ((int)r.size() + 1 > buf_len) return 2;
        for (size_t i = 0; i < r.size(); ++i)
            buf[i] = r[i] == exon::host::TableProviderFilterPushDown::Exact ? 'E'
                     : r[i] == exon::host::TableProviderFilterPushDown::Inexact ? 'I' : 'U';
        buf[r.size()] = 0;
        return 0;
    } catch (const std::exception &e) {
        g_host_error = e.what();
        return 1;
    }
}

int64_t exon_host_gpu_launches(void *s) { return static_cast<ExonSession *>(s)->gpu_launches(); }

}  // extern "C"
