// arrow_stream.cpp -- Arrow C STREAM interface over a partition stream (SURVEY.md 8b seam B4).
//
// The reference hands record batches across its FFI as an `FFI_ArrowArrayStream`
// (`create_dataset_stream_from_table_provider`, exon/exon-core/src/ffi/mod.rs:58-73; the stream object itself is built
// at :25-49 from a DataFusion `SendableRecordBatchStream`) and exon-r / exon-py consume exactly that struct:
// get_schema / get_next / get_last_error / release.  exon_gpu_stream_export gives the same object for a GPU-built
// partition stream, so a Rust host wraps it with `ArrowArrayStreamReader::from_raw` instead of
// looping over exon_gpu_*_next_batch itself.
//
// Every format's column builder sits behind the same call here (VCF, FASTQ, FASTA, BAM, GFF, mzML).  Host-resident
// batches only: device-resident columns would need the Arrow C *Device* stream, which nothing on the reference side reads.
#include <cerrno>
#include <cstring>
#include <string>

#include "internal.h"

namespace exon {
namespace {

struct StreamPriv {
    exon_gpu_stream *s = nullptr;
    bool owns = false;
    std::string last_error;
};

int next_any(exon_gpu_stream *s, ArrowArray *out) {
    switch (s->fmt) {
        case kFmtVcf: return exon_gpu_vcf_next_batch(s, out, nullptr);
        case kFmtFastq: return exon_gpu_fastq_next_batch(s, out, nullptr);
        case kFmtFasta: return exon_gpu_fasta_next_batch(s, out, nullptr);
        case kFmtBam: return exon_gpu_bam_next_batch(s, out, nullptr);
        case kFmtGff: return exon_gpu_gff_next_batch(s, out, nullptr);
        case kFmtMzml: return exon_gpu_mzml_next_batch(s, out, nullptr);
        default: return fail(EXON_GPU_ERR_ARG, "stream_export: unknown stream format");
    }
}

int to_errno(int rc) {
    switch (rc) {
        case EXON_GPU_OK: return 0;
        case EXON_GPU_ERR_ARG: return EINVAL;
        case EXON_GPU_ERR_OOM: return ENOMEM;
        case EXON_GPU_ERR_UNSUPPORTED: return ENOTSUP;
        default: return EIO;  // CUDA / parse / state / NCCL: the message says which
    }
}

int get_schema(ArrowArrayStream *st, ArrowSchema *out) {
    auto *p = static_cast<StreamPriv *>(st->private_data);
    memset(out, 0, sizeof(*out));
    const int rc = exon_gpu_stream_schema(p->s, out);
    if (rc != EXON_GPU_OK) p->last_error = exon_gpu_last_error();
    return to_errno(rc);
}

int get_next(ArrowArrayStream *st, ArrowArray *out) {
    auto *p = static_cast<StreamPriv *>(st->private_data);
    memset(out, 0, sizeof(*out));
    const int rc = next_any(p->s, out);  // release == NULL marks the end of the stream
    if (rc != EXON_GPU_OK) {
        p->last_error = exon_gpu_last_error();
        memset(out, 0, sizeof(*out));
    }
    return to_errno(rc);
}

const char *get_last_error(ArrowArrayStream *st) {
    auto *p = static_cast<StreamPriv *>(st->private_data);
    return p->last_error.empty() ? nullptr : p->last_error.c_str();
}

void release(ArrowArrayStream *st) {
    auto *p = static_cast<StreamPriv *>(st->private_data);
    if (p) {
        if (p->owns) exon_gpu_stream_close(p->s);
        delete p;
    }
    st->release = nullptr;
    st->private_data = nullptr;
}

}  // namespace
}  // namespace exon

using namespace exon;

extern "C" {

// Schema of the batches a stream produces, without producing one (the projection fixes it).
int exon_gpu_stream_schema(exon_gpu_stream *s, struct ArrowSchema *out) {
    if (!s || !out) return fail(EXON_GPU_ERR_ARG, "stream_schema: NULL argument");
    switch (s->fmt) {
        case kFmtVcf: vcf_stream_schema(s, out); return EXON_GPU_OK;
        case kFmtFastq:
        case kFmtFasta: fastq_stream_schema(s, out); return EXON_GPU_OK;
        case kFmtBam: bam_stream_schema(s, out); return EXON_GPU_OK;
        case kFmtGff: gff_stream_schema(s, out); return EXON_GPU_OK;
        case kFmtMzml: return mzml_stream_schema(s, out);
        default: return fail(EXON_GPU_ERR_ARG, "stream_schema: unknown stream format");
    }
}

int exon_gpu_stream_export(exon_gpu_stream *s, struct ArrowArrayStream *out, int take_ownership) {
    if (!s || !out) return fail(EXON_GPU_ERR_ARG, "stream_export: NULL argument");
    if (s->columns_on_device)
        return fail(EXON_GPU_ERR_UNSUPPORTED, "stream_export: the stream keeps its columns in device memory; the Arrow C stream carries host buffers");
    if (s->projection.empty() && s->fmt != kFmtVcf)
        return fail(EXON_GPU_ERR_STATE, "stream_export: the stream was opened without a projection (fused query only)");
    auto *p = new StreamPriv();
    p->s = s;
    p->owns = take_ownership != 0;
    memset(out, 0, sizeof(*out));
    out->get_schema = get_schema;
    out->get_next = get_next;
    out->get_last_error = get_last_error;
    out->release = release;
    out->private_data = p;
    return EXON_GPU_OK;
}

}  // extern "C"
