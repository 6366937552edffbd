// vcf_wide.cuh -- what the two builders of VCF columns 2..8 share: the row-parallel build on the line index (vcf_wide.cu: INFO /
// FORMAT, and columns 2..6 whenever one of those two is projected) and the tile-pipeline build of columns 2..6 inside K2's
// two passes (vcf_columns.cu).  Both fill the same column store; wide_export (vcf_wide.cu) turns it into Arrow arrays.
#pragma once
#include <cstdint>
#include <vector>

#include "internal.h"

namespace exon {

constexpr uint32_t kWErrFields = 1u;       // fewer than 8 tab-separated fields
constexpr uint32_t kWErrQual = 2u;         // QUAL is not a float literal
constexpr uint32_t kWErrQualDigits = 4u;   // QUAL has more than 36 significant digits
constexpr uint32_t kWErrFieldLen = 8u;     // a line of 2 GiB or more
constexpr uint32_t kWErrInfoValue = 16u;   // INFO: a non-flag key without a value (the reference unwraps a None there)
constexpr uint32_t kWErrInfoKey = 32u;     // INFO: a key the header does not define (reserved: such keys fall back to noodles' tables)
constexpr uint32_t kWErrInfoForm = 64u;    // INFO: a value noodles cannot parse as its declared type, or a flag with a value
constexpr uint32_t kWErrFmtValue = 128u;   // FORMAT: a missing sample value ('.'): the reference unwraps a None there
constexpr uint32_t kWErrFmtForm = 256u;    // FORMAT: a sample value noodles cannot parse as its declared type

enum { kIdE = 0, kIdB = 1, kRefB = 2, kFiE = 3, kFiB = 4, kInfoB = 5, kFmtB = 6, kNScan = 7 };

struct WideBuf {
    void *d = nullptr, *h = nullptr;
    size_t bytes = 0;
};

struct WideStore {
    int device = 0;
    bool on_device = false;
    int batch_rows = 8192, wpb = 256;
    int64_t n_batches = 0, n_rows = 0;
    bool want[9] = {false, false, false, false, false, false, false, false, false};
    WideBuf id_loff, id_coff, id_val, id_valid, ref_off, ref_val, alt_valid, zeros, qual, qual_valid, fi_loff, fi_coff, fi_val, info_off, info_val, info_tab,
        fmt_off, fmt_val, fmt_tab;
    std::vector<long long> batch_row0, base[kNScan];  // per batch (+ total): global item / byte offset of the batch's first row
    static constexpr int kBufs = 19;
    void all(WideBuf *out[kBufs]) {
        WideBuf *v[kBufs] = {&id_loff, &id_coff, &id_val, &id_valid, &ref_off, &ref_val, &alt_valid, &zeros, &qual, &qual_valid, &fi_loff, &fi_coff, &fi_val, &info_off, &info_val, &info_tab,
                             &fmt_off, &fmt_val, &fmt_tab};
        for (int i = 0; i < kBufs; ++i) out[i] = v[i];
    }
    template <class T>
    const T *p(const WideBuf &b) const { return static_cast<const T *>(on_device ? b.d : b.h); }
};

// One row whose QUAL is not a short unsigned integer: the exact parser (f32_parse.cuh) runs on these in its own kernel.
struct QualSlow {
    const uint8_t *p;
    uint32_t n, pad_;
    unsigned long long row;
};
// qual[row] / the row's validity bit (absolute row numbering, one bit per row) for every entry of the list; error bits into
// flags[0], the smallest failing row into first_bad_row (vcf_wide.cu)
int wide_qual_list(Ctx *ctx, const QualSlow *list, unsigned long long n, float *qual, uint32_t *valid_abs, uint32_t *flags,
                   unsigned long long *first_bad_row);
// the error message of a failed wide build
int wide_fail(uint32_t e, unsigned long long row);

}  // namespace exon
