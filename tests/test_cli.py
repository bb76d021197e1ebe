"""Command-line drop-in parity (-m gpu): the new `pingpongpro` binary against
  * the reference CLI's output files committed under tests/golden/files/*/expected_* (byte for byte), and
  * the reference binary itself (oracle/_ref/pingpongpro_ref, built from the unmodified source) on fresh seeded
    inputs, when it travelled to the box.
Transposon p-values come from a device exp() instead of the host's (1e-16 relative on the fp64 sum), so the
transposon tables are compared field by field with a one-ulp allowance on the printed 7 significant digits."""
import filecmp
import os
import subprocess

import pytest

from golden_tools import GOLDEN
from oracle_tools import have_ref, ref_cli_path
from pingpongpro_b200 import _build

pytestmark = pytest.mark.gpu

FILES = os.path.join(GOLDEN, "files", "files_tiny_weighted")
SYNTH = os.path.join(_build.BIN, "ppp_synth")


@pytest.fixture(scope="module")
def cli():
    path = _build.build_cli()
    assert path and os.path.exists(path)
    return path


def run(binary, args, cwd=None):
    return subprocess.run([binary, *args], cwd=cwd, capture_output=True, text=True)


def same_number(a: str, b: str) -> bool:
    if a == b:
        return True
    try:
        x, y = float(a), float(b)
    except ValueError:
        return False
    return abs(x - y) <= 2e-6 * max(abs(x), abs(y))


def assert_same_tree(got: str, want: str):
    want_files = sorted(f for f in os.listdir(want) if f != "ARGS" and not f.endswith(".pdf"))
    got_files = sorted(f for f in os.listdir(got) if not f.endswith(".pdf"))
    assert got_files == want_files
    for f in want_files:
        a, b = os.path.join(got, f), os.path.join(want, f)
        if filecmp.cmp(a, b, shallow=False):
            continue
        assert "transposons" in f, f"{f} differs"  # only the p/q columns of transposon tables may differ in the last digit
        la, lb = open(a).read().splitlines(), open(b).read().splitlines()
        assert len(la) == len(lb), f
        for x, y in zip(la, lb):
            fx, fy = x.split("\t"), y.split("\t")
            assert len(fx) == len(fy) and all(same_number(p, q) for p, q in zip(fx, fy)), (f, x, y)


@pytest.mark.parametrize("tag", ["default", "s2_b", "unique_T"])
def test_committed_reference_outputs(cli, tag, tmp_path):
    want = os.path.join(FILES, "expected_" + tag)
    extra = open(os.path.join(want, "ARGS")).read().split()
    out = tmp_path / "out"
    r = run(cli, ["-i", os.path.join(FILES, "input.bam"), "-o", str(out), "-t", os.path.join(FILES, "annotation.tsv"), *extra])
    assert r.returncode == 0, r.stderr
    assert_same_tree(str(out), want)


@pytest.mark.parametrize("threads", ["1", "2", "3", "7"])
def test_thread_counts_give_identical_output(cli, threads, tmp_path):
    """--threads only changes how many workers inflate / parse / extract; the file order of the reads (which fixes the
    float additions of -m weighted, :487) and therefore every output byte must not depend on it.  The input is large
    enough for several worker jobs per format, and the signature table spans several formatting chunks."""
    f = tmp_path / "in.bam"
    subprocess.check_call([SYNTH, "--contigs", "c1:900000,c2:300000,c3:40000", "--reads", "400000", "--seed", "91", "--out", str(f),
                           "--annot", str(tmp_path / "annot.tsv")])
    subprocess.check_call([SYNTH, "--contigs", "c1:900000,c2:300000,c3:40000", "--reads", "100000", "--seed", "92", "--out", str(tmp_path / "more.sam")])
    args = ["-i", str(f), "-i", str(tmp_path / "more.sam"), "-t", str(tmp_path / "annot.tsv"), "-b"]
    base, other = tmp_path / "base", tmp_path / "other"
    assert run(cli, [*args, "-o", str(base)]).returncode == 0
    assert run(cli, [*args, "-o", str(other), "--threads", threads]).returncode == 0
    for name in sorted(os.listdir(base)):
        assert filecmp.cmp(base / name, other / name, shallow=False), name


def test_sam_equals_bam(cli, tmp_path):
    a, b = tmp_path / "a", tmp_path / "b"
    assert run(cli, ["-i", os.path.join(FILES, "input.sam"), "-o", str(a), "-b"]).returncode == 0
    assert run(cli, ["-i", os.path.join(FILES, "input.bam"), "-o", str(b), "-b"]).returncode == 0
    assert_same_tree(str(a), str(b))


LIVE = [
    ["-m", "discard", "-b"],
    ["-l", "20", "-L", "35", "-s", "3"],
    ["-m", "unique", "-T", "300", "-b"],
    ["-T", "60", "-t", "{annot.tsv}", "-b"],          # -t before -T: the prediction loop stops at annotated contigs without signatures
    ["-t", "{annot.bed}", "-t", "{annot.csv}", "-t", "{annot.gff}", "-p"],
    ["-t", "{overlap.tsv}", "-T", "100", "-v"],        # nested / duplicate intervals: the list-iterator quirk; annotated contigs without signatures
]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/pingpongpro_ref did not travel to this box")
@pytest.mark.parametrize("extra", LIVE, ids=[" ".join(e) for e in LIVE])
def test_against_reference_binary(cli, extra, tmp_path):
    d = tmp_path
    subprocess.check_call([SYNTH, "--contigs", "c1:90000,c2:61000,c3:30000,c4:2000", "--reads", "70000", "--seed", "77", "--out", str(d / "x.bam"),
                           "--annot", str(d / "annot.tsv")])
    subprocess.check_call([SYNTH, "--contigs", "c1:90000,c2:61000,c3:30000,c4:2000", "--reads", "30000", "--seed", "78", "--out", str(d / "y.sam")])
    for ext in ("bed", "csv", "gff"):
        subprocess.check_call([SYNTH, "--contigs", "c1:90000,c2:61000,c3:30000,c4:2000", "--reads", "1", "--seed", "77", "--annot", str(d / f"annot.{ext}")])
    with open(d / "overlap.tsv", "w") as f:
        f.write("all\t+\tc1\t1\t90000\nin1\t-\tc1\t100\t5000\nin2\t+\tc1\t100\t5000\nlate\t+\tc1\t60000\t70000\nc2a\t+\tc2\t1\t61000\nc2b\t+\tc2\t20000\t30000\n"
                "other\t+\tcX\t5\t10\nempty\t-\tc4\t1\t100\n")
    args = [str(d / a[1:-1]) if a.startswith("{") else a for a in extra]
    inputs = ["-i", str(d / "x.bam"), "-i", str(d / "y.sam")]
    mine, ref = d / "mine", d / "ref"
    r1 = run(cli, [*inputs, "-o", str(mine), *args])
    r2 = run(ref_cli_path(), [*inputs, "-o", str(ref), *args])
    assert r1.returncode == r2.returncode == 0, (r1.stderr, r2.stderr)
    assert_same_tree(str(mine), str(ref))
    if "-v" in extra:  # same progress messages (seconds may differ)
        strip = lambda s: [ln.split("(")[0] for ln in s.splitlines()]
        assert strip(r1.stderr) == strip(r2.stderr)


def test_error_paths_and_exit_codes(cli, tmp_path):
    bam = os.path.join(FILES, "input.bam")
    assert run(cli, ["--help"]).returncode == 1          # :1465-1466: every non-OK parse result exits with 1
    assert run(cli, ["--version"]).returncode == 1
    assert run(cli, []).returncode == 1                  # -i is required
    r = run(cli, ["-i", bam, "-l", "30", "-L", "25"])
    assert r.returncode == 1 and "maximum alignment length (25) must not be lower than minimum alignment length (30)" in r.stderr
    assert run(cli, ["-i", str(tmp_path / "x.txt")]).returncode == 1  # extension check
    assert run(cli, ["-i", bam, "-T", "10"]).returncode == 1          # RANGE >= 30
    assert run(cli, ["-i", bam, "-m", "bogus"]).returncode == 1
    r = run(cli, ["-i", str(tmp_path / "missing.bam")])
    assert r.returncode == 1 and r.stderr.strip() == f"Failed to open input file: {tmp_path / 'missing.bam'}"
    r = run(cli, ["-i", bam, "-t", str(tmp_path / "missing.tsv"), "-o", str(tmp_path / "o")])
    assert r.returncode == 1 and f'Failed to open transposon file "{tmp_path / "missing.tsv"}".' in r.stderr
    other = tmp_path / "other.sam"
    other.write_text("@SQ\tSN:different\tLN:500\n")
    r = run(cli, ["-i", bam, "-i", str(other), "-o", str(tmp_path / "o")])
    assert r.returncode == 1 and f"@SQ header lines of '{other}' differ from those of previous input files" in r.stderr
    bad = tmp_path / "bad.sam"
    bad.write_text("@SQ\tSN:c\tLN:100\nr1\tnotanumber\tc\t5\t255\t25M\t*\t0\t0\tACGT\tIIII\n")
    r = run(cli, ["-i", str(bad), "-o", str(tmp_path / "o")])
    assert r.returncode == 1 and "Failed to read record" in r.stderr
    r = run(cli, ["-i", bam, "-o", str(tmp_path / "no" / "such" / "dir")])
    assert r.returncode == 1 and r.stderr.startswith("Failed to open output directory: ")


def test_combined_short_flags(cli, tmp_path):
    """-bv is -b -v (SeqAn's ArgumentParser accepts grouped flags); the last flag of a group may take a value."""
    src = os.path.join(GOLDEN, "files", "files_tiny_weighted", "input.sam")
    a, b, c = tmp_path / "a", tmp_path / "b", tmp_path / "c"
    assert run(cli, ["-b", "-v", "-s", "2", "-i", src, "-o", str(a)]).returncode == 0
    assert run(cli, ["-bv", "-s", "2", "-i", src, "-o", str(b)]).returncode == 0
    assert run(cli, ["-vbs", "2", "-i", src, "-o", str(c)]).returncode == 0
    names = sorted(os.listdir(a))
    assert len(names) == 4 and names == sorted(os.listdir(b)) == sorted(os.listdir(c))
    for n in names:
        assert (a / n).read_bytes() == (b / n).read_bytes() == (c / n).read_bytes()
    assert run(cli, ["-sb", "2", "-i", src, "-o", str(tmp_path / "d")]).returncode == 1  # a value flag inside a group


def test_output_goes_to_cwd_by_default(cli, tmp_path):
    r = run(cli, ["-i", os.path.join(FILES, "input.bam")], cwd=str(tmp_path))
    assert r.returncode == 0 and os.path.exists(tmp_path / "ping-pong_signatures.tsv")
    assert filecmp.cmp(tmp_path / "ping-pong_signatures.tsv", os.path.join(FILES, "expected_default", "ping-pong_signatures.tsv"), shallow=False)
