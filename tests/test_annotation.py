"""Annotation reader (pingpongpro_b200/host/annotation.cpp) against the reference's readTransposonsFromFile
(pingpongpro.cpp:993-1175), run through the include-harness on hand-written files that exercise its quirks."""
import os

import numpy as np
import pytest

from golden_tools import GOLDEN
from oracle_tools import have_ref, run_harness
from pingpongpro_b200 import host

BAM = os.path.join(GOLDEN, "files", "files_tiny_weighted", "input.bam")

FILES = {
    "plain.tsv": "T1\t+\tctgA\t101\t200\nT2\t-\tctgB\t5\t50\nT3\t+\tctgA\t101\t150\n",
    "quirks.tsv": (
        "ID\tstrand\tcontig\tstart\tend\n"           # header line: strand 'strand' -> dropped
        "A\t+\tctgA\t10\t20\textra\tcolumns\n"
        "\t\tB\t-\tctgA\t30\t40\n"                   # empty leading fields are skipped, not counted
        "C\t*\tctgA\t1\t2\n"                         # strand neither + nor -
        "D\t+\tnewContig\t7\t9\n"                    # unknown contig is appended to the name store
        "E\t+\tctgC\t0\t5\n"                         # start 0 - offset 1 = -1 -> dropped
        "F\t+\tctgC\tabc\t5\n"                       # conversion error
        "G\t-\tctgC\t12x\t15\n"                      # atoi prefix
        "H\t+\tctgB\t100\t90\r\n"                    # CR LF, end < start is kept
        "I\t+\tctgB\t100\t100"                        # no trailing newline
    ),
    "a.bed": "ctgA 100 200 bedOne 0 +\nctgA\t50\t60\tbedTwo\t0\t-\ntrack name=x\nctgB  10   20  bedThree  5  +\nctgZ 1 2 n 0 -\n",
    "b.csv": 'c1,+,ctgA,11,20\n"quoted, id",-,ctgB,5,9\n"a ""b"" c",+,ctgA,3,4\nlast,+,ctgC,1,1\n"multi\nline",+,ctgA,8,9\n',
    "c.gff": "ctgA\tctgA\tte\t100\t200\t.\t+\t.\tgffOne\nx\tctgB\tte\t5\t6\t.\t-\t.\tgffTwo\n#comment\n",
    "d.gtf": "ctgC\tctgC\texon\t1\t10\t.\t+\t.\tgene_id \"g1\"; transcript_id \"t1\";\n",
    "dups.tsv": "X\t+\tctgA\t500\t600\nY\t-\tctgA\t500\t600\nZ\t+\tctgA\t500\t550\nW\t+\tctgA\t400\t9000\n",
}


@pytest.fixture(scope="module")
def annot_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("annot")
    for name, text in FILES.items():
        (d / name).write_bytes(text.encode())
    return d


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("files", [["plain.tsv"], ["quirks.tsv"], ["a.bed"], ["b.csv"], ["c.gff"], ["d.gtf"], ["dups.tsv"],
                                   ["quirks.tsv", "a.bed", "b.csv", "c.gff", "d.gtf", "dups.tsv", "plain.tsv"]])
def test_reader_equals_reference(annot_dir, files, tmp_path):
    paths = [str(annot_dir / f) for f in files]
    dump = run_harness(str(tmp_path / "d.bin"), [BAM], annots=paths)
    ref = dump.get("transposons")
    dec = host.decode_file(BAM)
    iv, ids, names = host.read_annotation(paths, dec.names)
    assert ref is not None and len(iv) == len(ref)
    for k in ("contig", "strand", "start", "end"):
        assert np.array_equal(iv[k], ref[k]), k


def test_specific_quirks(annot_dir):
    iv, ids, names = host.read_annotation([str(annot_dir / "quirks.tsv")], ["ctgA", "ctgB", "ctgC"])
    assert names == ["ctgA", "ctgB", "ctgC", "contig", "newContig"]  # header line's contig column and D's contig are appended
    got = {i: (int(v["contig"]), int(v["strand"]), int(v["start"]), int(v["end"])) for i, v in zip(ids, iv)}
    assert got == {"A": (0, 0, 9, 19), "B": (0, 1, 29, 39), "D": (4, 0, 6, 8), "G": (2, 1, 11, 14), "H": (1, 0, 99, 89), "I": (1, 0, 99, 99)}
    iv, ids, names = host.read_annotation([str(annot_dir / "b.csv")], ["ctgA", "ctgB", "ctgC"])
    # a line feed inside quotes still ends the record at the next comma (:1070-1074 set newLine before the quote test):
    # "multi\nline" becomes the lone identifier "multi,line" (dropped), the rest of the line is parsed as a new record
    assert ids == ['a "b" c', "c1", "quoted, id", "last"] and names[-1] == "8"
    iv, ids, _ = host.read_annotation([str(annot_dir / "dups.tsv")], ["ctgA"])
    assert ids == ["W", "Z", "X", "Y"]  # (start, end) order, ties in file order (stable sort, :1174)


def test_missing_file():
    with pytest.raises(OSError):
        host.read_annotation(["/nonexistent/x.tsv"], ["a"])
