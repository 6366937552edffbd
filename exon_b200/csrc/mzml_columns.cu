// mzml_columns.cu -- mzML record batches (placeholder until the column build lands in this round)
#include "internal.h"

namespace exon {
int mzml_stream_schema(VcfStream *, ArrowSchema *) { return fail(EXON_GPU_ERR_UNSUPPORTED, "mzml: record batches are not built yet"); }
}  // namespace exon

extern "C" int exon_gpu_mzml_open_columns(exon_gpu_ctx *, const exon_gpu_fastq_opts *, exon_gpu_stream **) {
    return exon::fail(EXON_GPU_ERR_UNSUPPORTED, "mzml: record batches are not built yet");
}

extern "C" int exon_gpu_mzml_next_batch(exon_gpu_stream *, struct ArrowArray *, struct ArrowSchema *) {
    return exon::fail(EXON_GPU_ERR_UNSUPPORTED, "mzml: record batches are not built yet");
}
