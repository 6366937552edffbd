// exon_host.hpp -- C++ mirror of the reference's host-side operators for the VCF scan -> filter -> aggregate path,
// sitting above the C ABI of include/exon_gpu.h.  The reference is Rust on DataFusion; no cargo/rustc exists in this
// image, so the same interfaces are restated in C++ with the reference's names, argument meaning and error
// behaviour (paths relative to the reference repo root):
//
//   ExonSession::{new_exon, sql, read_vcf}       exon/exon-core/src/session_context/exon_context_ext.rs:103-721
//   ScanFunction (vcf_scan / vcf_indexed_scan)   exon/exon-core/src/datasources/scan_function.rs:32-64,
//                                                exon/exon-core/src/datasources/vcf/udtf.rs:55-145
//   ListingVCFTableOptions / ListingVCFTable     exon/exon-core/src/datasources/vcf/table_provider.rs:60-444
//     ::supports_filters_pushdown                 :299-320
//     ::scan                                      :322-443
//   VCFScan::{repartitioned, execute}            exon/exon-core/src/datasources/vcf/scanner.rs:103-162
//   ExonFileScanConfig::regroup_files_by_size    exon/exon-core/src/datasources/exon_file_scan_config.rs:79-110
//   VCFOpener::open                              exon/exon-core/src/datasources/vcf/file_opener/unindex_file_opener.rs:48-92
//   infer_region_from_udf                        exon/exon-core/src/physical_plan/infer_region.rs:25-42
//   hive partition pruning                       exon/exon-core/src/physical_plan/object_store/hive_partition.rs:32-157
//
//   IndexedVCFOpener::open + get_byte_range_for_file
//                                                exon/exon-core/src/datasources/vcf/file_opener/indexed_file_opener.rs:53-214,
//                                                exon/exon-core/src/datasources/indexed_file/indexed_bgzf_file.rs:52-83
//   fastq_scan / bam_scan / mzml_scan (COUNT(*)) exon/exon-core/src/datasources/{fastq,bam,mzml}/udtf.rs
//
// Everything that touches record bytes runs on the GPU through exon_gpu_* (BGZF members are inflated on the device,
// tabix chunks restrict which members); this layer only plans (which files,
// which partition, which predicate) and formats results.  DataFusion's SQL front end is third party and out of
// scope: `ExonSession::sql` understands exactly the statement shapes the reference's VCF sqllogictests use.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

struct exon_gpu_ctx;

namespace exon::host {

// DataFusionError variants the reference raises on this path.
struct ExonError : std::runtime_error {
    enum Kind { Plan, Execution, NotImplemented, External, Internal };
    Kind kind;
    ExonError(Kind k, const std::string &m) : std::runtime_error(m), kind(k) {}
};

enum class FileCompressionType { UNCOMPRESSED, GZIP };
enum class ExonFileType { VCF, INDEXED_VCF };
enum class TableProviderFilterPushDown { Unsupported, Inexact, Exact };

// noodles_core::Region: name + optional closed 1-based interval.
struct Region {
    std::string name;
    bool has_interval = false;
    int64_t lo = 1, hi = INT64_MAX;
    static Region parse(const std::string &s);  // exon_gpu_region_parse
};

// The slice of datafusion::logical_expr::Expr this path sees.
struct Expr {
    enum Kind { Column, Utf8, Int64, Boolean, Binary, Between, ScalarFunction } kind = Column;
    std::string name;         // Column / ScalarFunction name, Utf8 value, Binary operator ("=", ">=", "<=", ">", "<", "AND")
    int64_t i64 = 0;          // Int64 / Boolean value
    std::vector<Expr> args;   // Binary: [l, r]; Between: [expr, low, high]; ScalarFunction: arguments
    std::string to_string() const;
};

struct PartitionedFile {
    std::string path;
    int64_t size = 0;
    std::vector<std::string> partition_values;  // one per table partition column
};

struct ListingVCFTableOptions {
    std::string file_extension = "vcf";
    FileCompressionType file_compression_type = FileCompressionType::UNCOMPRESSED;
    bool indexed = false;
    std::vector<Region> regions;
    std::vector<std::string> table_partition_cols;
    static ListingVCFTableOptions make(FileCompressionType c, bool indexed);  // ListingVCFTableOptions::new
};

struct RecordBatch {  // the projected columns of one batch, copied to the host for display
    int64_t num_rows = 0;
    std::vector<std::string> chrom;
    std::vector<int64_t> pos;
};

struct SessionConfig {
    int batch_size = 8192;          // exon-core/src/config/mod.rs:24,36
    int target_partitions = 0;      // 0 = hardware threads (config/mod.rs:44)
    bool gpu_fused = true;          // exon.gpu_fused: fused K1 count vs K2 batches + K3 accumulate
    bool gpu_strict = true;         // exon.gpu_strict: validate every row like LazyVCFArrayBuilder::append does
    std::map<std::string, std::string> options;  // every `SET exon.x = v` seen
};

class ExonSession;

// ExecutionPlan of the non-indexed and indexed scans (VCFScan / IndexedVCFScanner).
class VCFScan {
public:
    std::vector<std::vector<PartitionedFile>> file_groups;
    std::vector<int> projection;  // file-schema column indices: 0 chrom, 1 pos
    FileCompressionType compression = FileCompressionType::UNCOMPRESSED;
    bool has_region = false;      // IndexedVCFScanner: records are filtered by `region` inside the scan
    Region region;
    // VCFScan::repartitioned: regroup whole files by size into min(target_partitions, n_files) groups.
    std::shared_ptr<VCFScan> repartitioned(int target_partitions) const;
    size_t output_partitioning() const { return file_groups.size(); }
};

// TableProvider.
class ListingVCFTable {
public:
    std::string table_path;
    ListingVCFTableOptions options;
    std::vector<PartitionedFile> list_files() const;  // every file under table_path with the expected extension
    std::vector<TableProviderFilterPushDown> supports_filters_pushdown(const std::vector<Expr> &filters) const;
    // Errors: NotImplemented("Multiple regions are not supported yet"), Plan("INDEXED_VCF table requires a region
    // filter. See the UDF 'vcf_region_filter'.")
    std::shared_ptr<VCFScan> scan(const std::vector<int> *projection, const std::vector<Expr> &filters, const int64_t *limit) const;
};

struct ResultSet {
    std::vector<std::string> columns;
    std::vector<std::vector<std::string>> rows;
    std::string to_text() const;  // one line per row, cells separated by one space (the slt runner's rendering)
};

class ExonSession {
public:
    static std::unique_ptr<ExonSession> new_exon(int device = 0);  // ExonSession::new_exon
    ~ExonSession();
    SessionConfig config;
    ResultSet sql(const std::string &query);                         // ExonSession::sql
    std::shared_ptr<ListingVCFTable> read_vcf(const std::string &table_path, const ListingVCFTableOptions &options);
    // COUNT(*) over a scan with the residual filters DataFusion would put in FilterExec above it.
    int64_t count(const ListingVCFTable &table, const std::vector<Expr> &filters);
    std::vector<RecordBatch> collect(const ListingVCFTable &table, const std::vector<Expr> &filters, int64_t limit);
    int64_t gpu_launches() const;

private:
    ExonSession() = default;
    exon_gpu_ctx *ctx_ = nullptr;
    std::map<std::string, std::shared_ptr<ListingVCFTable>> tables_;
    friend struct Executor;
};

}  // namespace exon::host
