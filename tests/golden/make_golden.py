"""Generates the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run here (the container that has /root/reference) after `make -C oracle ref`:

    python tests/golden/make_golden.py

For every case it writes a seeded synthetic input with pingpongpro_b200/bin/ppp_synth, runs
oracle/_ref/ppp_harness (the reference's own functions behind an include-harness) on it and stores
  * the packed upload records the host front end extracts from that file, and
  * every intermediate the reference produced (stacks, height frequencies, grouped counts before and
    after collapse, signatures with bins and FDR, transposon results)
in tests/golden/<case>.npz.  For the `files_*` cases the input files and the reference CLI's output
files are stored too (tests/golden/files/<case>/), for the front-end and command-line parity tests.
The reference has no golden vectors of its own; these are the project's KATs.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle_tools import ref_cli_path, run_harness  # noqa: E402
from pingpongpro_b200 import _build, host  # noqa: E402

SYNTH = os.path.join(_build.BIN, "ppp_synth")

CASES = [
    # name, synth args, mode, annot ext, max_overlap, keep files
    dict(name="files_tiny_weighted", synth=["--genome", "tiny", "--reads", "6000", "--seed", "11"], mode="weighted", annot="tsv", files=True),
    dict(name="tiny_unique", synth=["--genome", "tiny", "--reads", "6000", "--seed", "12"], mode="unique", annot="bed"),
    dict(name="tiny_discard", synth=["--genome", "tiny", "--reads", "6000", "--seed", "13"], mode="discard", annot="csv"),
    dict(name="tiny_sorted_nh1", synth=["--genome", "tiny", "--reads", "8000", "--seed", "14", "--nh1", "--sorted"], mode="weighted", annot="gff"),
    dict(name="four_contigs_overlapping", synth=["--contigs", "a:400000,b:250000,c:100000,d:30000", "--reads", "40000", "--seed", "5"],
         mode="weighted", annot="overlapping"),
    dict(name="tiny_max30", synth=["--genome", "tiny", "--reads", "6000", "--seed", "15"], mode="weighted", annot="tsv", max_overlap=30),
    dict(name="tiny_max26_unique", synth=["--genome", "tiny", "--reads", "6000", "--seed", "16"], mode="unique", annot=None, max_overlap=26),
    dict(name="dense_weighted", synth=["--contigs", "p:30000,q:26000", "--reads", "80000", "--seed", "21"], mode="weighted", annot="tsv"),
    dict(name="dense_unique_max30", synth=["--contigs", "p:30000,q:26000", "--reads", "60000", "--seed", "22"], mode="unique", annot="bed", max_overlap=30),
    dict(name="dm6_120k_weighted", synth=["--genome", "dm6", "--reads", "120000", "--seed", "1"], mode="weighted", annot=None),
]


def overlapping_annotation(path, names, lens):
    """Nested / duplicate / overlapping intervals: exercises the reference's list-iterator walk."""
    rng = np.random.RandomState(3)
    rows = []
    for i in range(300):
        c = rng.randint(len(names))
        s = rng.randint(0, lens[c] - 100)
        e = s + rng.randint(1, 20000)
        rows.append((f"T{i}", "+-"[rng.randint(2)], names[c], s + 1, e))
    rows += [("N1", "+", names[0], 1, lens[0]), ("N2", "-", names[0], 1, 5), ("N3", "+", names[0], 1000, 2000), ("N4", "+", names[0], 1000, 2000),
             ("N5", "+", names[-1], 1, lens[-1]), ("N6", "+", names[-1], lens[-1] - 1000, lens[-1] - 500)]
    with open(path, "w") as f:
        for r in rows:
            f.write("%s\t%s\t%s\t%d\t%d\n" % r)


def main():
    _build.build_host()
    _build.build_synth()
    tmp = tempfile.mkdtemp(prefix="ppp_golden.")
    for case in CASES:
        name, mode, maxov = case["name"], case["mode"], case.get("max_overlap", 23)
        d = os.path.join(tmp, name)
        os.makedirs(d)
        sam, bam = os.path.join(d, "input.sam"), os.path.join(d, "input.bam")
        annots = []
        annot_args = []
        if case["annot"] and case["annot"] != "overlapping":
            annots = [os.path.join(d, "annotation." + case["annot"])]
            annot_args = ["--annot", annots[0]]
        subprocess.check_call([SYNTH, *case["synth"], "--out", bam, *annot_args])
        dec = host.decode_file(bam, mode=mode)
        if case["annot"] == "overlapping":
            annots = [os.path.join(d, "annotation.tsv")]
            overlapping_annotation(annots[0], dec.names, dec.lengths)
        dump = run_harness(os.path.join(d, "dump.bin"), [bam], mode=mode, annots=annots, max_overlap=maxov)
        out = dict(
            mode=np.array(mode), min_ov=dump["min_ov"], max_ov=dump["max_ov"], pp_ov=dump["pp_ov"],
            contig_names=np.array(dec.names), contig_lengths=np.array(dec.lengths, dtype=np.uint64),
            reads=np.concatenate(dec.reads) if dec.reads else np.empty(0, host.READ_DTYPE),
            reads_per_contig=np.array([len(r) for r in dec.reads], dtype=np.uint64),
            total=dump["total"], stacks=dump["stacks"], freq=dump["freq"], sigs=dump["sigs"], grouped_post=dump["grouped_post"],
        )
        gp = dump["grouped_pre"]
        nz = np.flatnonzero(gp)
        out["grouped_pre_idx"] = nz.astype(np.uint32)
        out["grouped_pre_val"] = gp.reshape(-1)[nz]
        if "transposons" in dump:
            out["transposons"] = dump["transposons"]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(f"{name}: reads={len(out['reads'])} stacks={len(dump['stacks'])} sigs={len(dump['sigs'])} C={dump['grouped_post'].shape[1]} "
              f"transposons={len(dump.get('transposons', []))}")
        if case.get("files"):
            fdir = os.path.join(HERE, "files", name)
            shutil.rmtree(fdir, ignore_errors=True)
            os.makedirs(fdir)
            subprocess.check_call([SYNTH, *case["synth"], "--out", sam])
            shutil.copy(bam, fdir)
            shutil.copy(sam, fdir)
            for a in annots:
                shutil.copy(a, fdir)
            for tag, extra in (("default", []), ("s2_b", ["-s", "2", "-b"]), ("unique_T", ["-m", "unique", "-T", "100", "-b"])):
                odir = os.path.join(fdir, "expected_" + tag)
                os.makedirs(odir)
                cmd = [ref_cli_path(), "-i", bam, "-o", odir, *extra]
                for a in annots:
                    cmd += ["-t", a]
                subprocess.check_call(cmd)
                with open(os.path.join(odir, "ARGS"), "w") as f:
                    f.write(" ".join(extra) + "\n")
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
