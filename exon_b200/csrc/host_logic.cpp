// host_logic.cpp -- planning-time helpers that stay on the host in the reference too: region literal
// parsing (a11) and the file -> partition assignment (a3).  Pure C++, no CUDA.
#include <algorithm>
#include <climits>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/exon_gpu.h"
#include "f32_parse.cuh"

namespace exon {
int fail(int code, const char *fmt, ...);

// Rust `usize::from_str`: optional '+', one or more ASCII digits, nothing else; overflow is an error.
static bool parse_position(const char *s, size_t n, int64_t *out) {
    size_t i = 0;
    if (n && s[0] == '+') i = 1;
    if (i >= n) return false;
    uint64_t v = 0;
    for (; i < n; ++i) {
        if (s[i] < '0' || s[i] > '9') return false;
        const uint64_t d = (uint64_t)(s[i] - '0');
        if (v > (UINT64_MAX - d) / 10) return false;
        v = v * 10 + d;
    }
    if (v == 0 || v > (uint64_t)INT64_MAX) return false;  // Position is NonZeroUsize
    *out = (int64_t)v;
    return true;
}

// noodles-core Interval::from_str: "", "a", "a-", "-b", "a-b" (1-based, inclusive).
static bool parse_interval(const char *s, size_t n, int64_t *lo, int64_t *hi) {
    *lo = 1;
    *hi = INT64_MAX;
    if (n == 0) return true;
    const char *dash = (const char *)memchr(s, '-', n);
    if (!dash) return parse_position(s, n, lo);
    const size_t a = (size_t)(dash - s), b = n - a - 1;
    if (a && !parse_position(s, a, lo)) return false;
    if (b && !parse_position(dash + 1, b, hi)) return false;
    return true;
}
}  // namespace exon

using namespace exon;

extern "C" {

// noodles-core Region::from_str at its reference call sites
// (exon/exon-core/src/physical_plan/infer_region.rs:25-42, exon/exon-core/src/udfs/vcf/mod.rs:85-95):
// the text after the LAST ':' is the interval if it parses as one, otherwise the whole string is the name.  As in
// noodles-core 0.15 (`rsplit_once(':')`, `Interval::from_str("")` = unbounded) an EMPTY suffix parses: "chr1:" is the
// contig "chr1" with an open interval, ":5-9" an empty name with an interval (it matches no CHROM).
int exon_gpu_region_parse(const char *s, char *name_buf, size_t name_buf_len, exon_gpu_region *out) {
    if (!s || !name_buf || !out) return fail(EXON_GPU_ERR_ARG, "region_parse: NULL argument");
    const size_t n = strlen(s);
    if (n == 0) return fail(EXON_GPU_ERR_ARG, "region_parse: empty region");
    if (n >= name_buf_len) return fail(EXON_GPU_ERR_ARG, "region_parse: name buffer too small");
    size_t name_len = n;
    int64_t lo = 1, hi = INT64_MAX;
    int has_interval = 0;
    const char *colon = strrchr(s, ':');
    if (colon) {
        int64_t a, b;
        if (parse_interval(colon + 1, n - (size_t)(colon - s) - 1, &a, &b)) {
            name_len = (size_t)(colon - s);
            lo = a;
            hi = b;
            has_interval = 1;
        }
    }
    memcpy(name_buf, s, name_len);
    name_buf[name_len] = '\0';
    out->chrom = name_buf;
    out->chrom_len = (int32_t)name_len;
    out->has_chrom = 1;
    out->has_interval = has_interval;
    out->lo = lo;
    out->hi = hi;
    return EXON_GPU_OK;
}

// Interval literal of interval_match (exon/exon-core/src/udfs/vcf/mod.rs:246-252).
int exon_gpu_interval_parse(const char *s, exon_gpu_region *out) {
    if (!s || !out) return fail(EXON_GPU_ERR_ARG, "interval_parse: NULL argument");
    int64_t lo, hi;
    if (!parse_interval(s, strlen(s), &lo, &hi)) return fail(EXON_GPU_ERR_ARG, "interval_parse: bad interval '%s'", s);
    out->chrom = nullptr;
    out->chrom_len = 0;
    out->has_chrom = 0;
    out->has_interval = 1;
    out->lo = lo;
    out->hi = hi;
    return EXON_GPU_OK;
}

// ExonFileScanConfig::regroup_files_by_size, exon/exon-core/src/datasources/exon_file_scan_config.rs:79-110.
int exon_gpu_regroup_files_by_size(const int64_t *sizes, int32_t n_files, int32_t target_partitions,
                                   int32_t *out_partition, int32_t *out_n_partitions) {
    if (n_files < 0 || (n_files > 0 && (!sizes || !out_partition)) || !out_n_partitions)
        return fail(EXON_GPU_ERR_ARG, "regroup_files_by_size: bad argument");
    if (target_partitions < 1) return fail(EXON_GPU_ERR_ARG, "regroup_files_by_size: target_partitions must be >= 1");
    if (n_files == 0) {
        *out_n_partitions = 0;
        return EXON_GPU_OK;
    }
    std::vector<int32_t> order((size_t)n_files);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return sizes[a] < sizes[b]; });
    const int32_t parts = std::min(target_partitions, n_files);
    for (int32_t i = 0; i < n_files; ++i) out_partition[order[(size_t)i]] = i % parts;
    *out_n_partitions = parts;
    return EXON_GPU_OK;
}

// Rust f32::from_str as applied to QUAL (exon/exon-vcf/src/array_builder/lazy_array_builder.rs:205-208); f32_parse.cuh.
int exon_gpu_parse_f32(const char *s, size_t len, float *out) {
    if (!s || !out) return fail(EXON_GPU_ERR_ARG, "parse_f32: NULL argument");
    if (len > 4096) return fail(EXON_GPU_ERR_UNSUPPORTED, "parse_f32: literal longer than 4096 bytes");
    const int rc = parse_f32_rust(reinterpret_cast<const uint8_t *>(s), (int)len, out);
    if (rc == kF32Malformed) return fail(EXON_GPU_ERR_PARSE, "parse_f32: invalid float literal '%.*s'", (int)len, s);
    if (rc == kF32Unsupported) return fail(EXON_GPU_ERR_UNSUPPORTED, "parse_f32: more than 36 significant digits in '%.*s'", (int)len, s);
    return EXON_GPU_OK;
}

}  // extern "C"
