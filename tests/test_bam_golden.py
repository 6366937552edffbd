"""The BAM oracle (oracle/bam_oracle.c) against the reference's known answers: slt/bam-select-tests.slt:9-12 (first row),
:17-31 (quality scores), :56-64 (61 / 122 rows) and the decoded fixture facts of SURVEY.md appendix A."""
import os

import oracle
from conftest import GOLDEN


def fixture():
    with open(os.path.join(GOLDEN, "test.bam"), "rb") as f:
        return f.read()


def test_first_row_and_counts():
    b = oracle.Bam(fixture())
    r = b.row(0)
    # READ_ID 83 chr1 12203704 12217173 NULL 55M13394N21M chr1
    assert (r["name"], r["flag"], r["reference"], r["start"], r["end"], r["mapping_quality"], r["cigar"], r["mate_reference"]) == \
        ("READ_ID", 83, "chr1", 12203704, 12217173, None, "55M13394N21M", "chr1")
    assert r["l_seq"] == 76                                                    # array_length(quality_score) = 76
    assert [b.row(i)["first_quals"][0] for i in range(5)] == [23, 20, 37, 34, 31]  # array_element(quality_score, 1)
    counts, rows = b.count_by_reference(all_rows=True)
    assert rows == 61 and counts["chr1"] == 61 and sum(counts.values()) == 61    # SELECT COUNT(*) -> 61
    assert len(b.refs) == 195 and b.refs[0] == "chr1"
    # every record has MAPQ 255 (NULL): a MAPQ comparison selects nothing; the flag tests alone keep the 58 primary ones
    assert sum(b.count_by_reference(flag_exclude=0x904, min_mapq=30)[0].values()) == 0
    flags = [b.row(i)["flag"] for i in range(61)]
    assert sorted(set(flags)) == [83, 97, 145, 147, 595, 659, 2177]
    assert sum(b.count_by_reference(flag_exclude=0x904)[0].values()) == sum(1 for f in flags if not f & 0x904) == 60
    # bam_region_filter('chr1:1-12209145', reference, start, end) -> 7 (slt/bam-indexed-select-tests.slt:11-14); two files -> 14 (:22-25)
    assert sum(b.count_by_reference(region=("chr1", 1, 12209145))[0].values()) == 7
    assert sum(b.count_by_reference(region=("chr1", None, None))[0].values()) == 61
    assert sum(b.count_by_reference(region=("chr2", 1, 10**9))[0].values()) == 0
    b.close()
    assert sum(oracle.bam_count_by_reference_files([fixture(), fixture()], region=("chr1", 1, 12209145))[0].values()) == 14
    assert oracle.bam_count_by_reference_files([fixture(), fixture()], all_rows=True)[1] == 122  # the bam-partition directory


def test_synthetic_truth():
    from synth import bam

    sh = bam.shards(30_000, 3)
    for kw in [dict(all_rows=True), dict(flag_exclude=0x904, min_mapq=30), dict(flag_require=0x10), dict(min_mapq=60)]:
        got, rows = oracle.bam_count_by_reference_files(sh.files, **kw)
        want = sh.truth(**{k: v for k, v in kw.items() if k != "all_rows"})
        assert rows == sh.n and got == want, kw
