"""`pingpongpro --gpus N` (-m gpu): the genome cut into N position ranges, one CUDA context each (they share the
device when the box has fewer GPUs), driven by one host process -- ppp_group_* collective, routed uploads,
ppp_transposons_sharded.  Every output byte must equal the single-context run: loci concatenate in rank order, the
all-reduced histograms give the same bins and FDRs, and transposons that straddle a range boundary (or contain one or
several whole ranges, or overlap each other: the list-iterator quirk of :1215-1222) get the sequential
single-precision sums of the unsharded run."""
import filecmp
import os
import subprocess

import pytest

from pingpongpro_b200 import _build

pytestmark = pytest.mark.gpu

SYNTH = os.path.join(_build.BIN, "ppp_synth")
CONTIGS = "c1:900000,c2:300000,c3:40000"  # 1 240 000 positions: --gpus 2 / 3 / 5 all cut c1, --gpus 5 also cuts c2

# hand-written intervals around the cuts at 413 333 / 826 666 (3 shards), 620 000 (2), 248 000 / 496 000 / 744 000 / 992 000 (5)
STRADDLERS = (
    "whole1\t+\tc1\t1\t900000\n"        # contains every cut of c1 and overlaps everything below (iterator quirk)
    "cut3a\t-\tc1\t400000\t430000\n"
    "cut2\t+\tc1\t600000\t640000\n"
    "cut2dup\t+\tc1\t600000\t640000\n"  # duplicate interval
    "cut3b\t+\tc1\t820000\t830000\n"
    "tail1\t-\tc1\t890000\t900063\n"    # up to the keys beyond the contig end
    "whole2\t+\tc2\t1\t300000\n"
    "cut5\t-\tc2\t90000\t95000\n"       # 992 000 - 900 000 = 92 000
    "whole3\t+\tc3\t1\t40000\n"
    "nowhere\t+\tcX\t5\t10\n"
)


@pytest.fixture(scope="module")
def cli():
    path = _build.build_cli()
    assert path and os.path.exists(path)
    return path


@pytest.fixture(scope="module")
def inputs(tmp_path_factory):
    d = tmp_path_factory.mktemp("multi_gpu")
    subprocess.check_call([SYNTH, "--contigs", CONTIGS, "--reads", "300000", "--seed", "57", "--out", str(d / "x.bam"), "--annot", str(d / "annot.tsv")])
    subprocess.check_call([SYNTH, "--contigs", CONTIGS, "--reads", "80000", "--seed", "58", "--out", str(d / "y.sam")])
    (d / "straddle.tsv").write_text(STRADDLERS)
    return d


def run(binary, args):
    return subprocess.run([binary, *args], capture_output=True, text=True)


@pytest.mark.parametrize("gpus,extra", [("3", ["-T", "200", "-b"]), ("2", ["-m", "unique", "-s", "2"]), ("5", ["-T", "60"])])
def test_gpus_flag_reproduces_the_single_context_output(cli, inputs, gpus, extra, tmp_path):
    d = inputs
    args = ["-i", str(d / "x.bam"), "-i", str(d / "y.sam"), "-t", str(d / "straddle.tsv"), "-t", str(d / "annot.tsv"), *extra]
    one, many = tmp_path / "one", tmp_path / "many"
    r1 = run(cli, [*args, "-o", str(one)])
    assert r1.returncode == 0, r1.stderr
    rn = run(cli, [*args, "-o", str(many), "--gpus", gpus])
    assert rn.returncode == 0, rn.stderr
    names = sorted(os.listdir(one))
    assert names == sorted(os.listdir(many)) and "transposons.tsv" in names
    for name in names:
        assert filecmp.cmp(one / name, many / name, shallow=False), name
    assert sum(1 for _ in open(one / "ping-pong_signatures.tsv")) > 1000  # not vacuous
