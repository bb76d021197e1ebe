"""Command-line end-to-end timing (file open -> last output byte): the new binary against the reference binary on the
same seeded BAM/SAM.  python -m pingpongpro_b200.tools.cli_e2e --reads 10000000 [--genome dm6]"""
import argparse
import filecmp
import json
import os
import subprocess
import tempfile
import time

from pingpongpro_b200 import _build

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=10_000_000)
ap.add_argument("--genome", default="dm6")
ap.add_argument("--format", default="bam")
ap.add_argument("--skip-reference", action="store_true")
a = ap.parse_args()
ref = os.path.join(_build.ROOT, "oracle", "_ref", "pingpongpro_ref")
d = tempfile.mkdtemp()
f = os.path.join(d, "reads." + a.format)
t0 = time.perf_counter()
subprocess.check_call([os.path.join(_build.BIN, "ppp_synth"), "--genome", a.genome, "--reads", str(a.reads), "--seed", "1", "--out", f])
out = {"reads": a.reads, "genome": a.genome, "format": a.format, "file_bytes": os.path.getsize(f), "generate_s": time.perf_counter() - t0, "host_cores": os.cpu_count()}
for name, binary, extra in (("b200", os.path.join(_build.BIN, "pingpongpro"), []), ("b200_1thread", os.path.join(_build.BIN, "pingpongpro"), ["--threads", "1"]),
                            ("reference", ref, [])):
    if name == "reference" and (a.skip_reference or not os.path.exists(ref)):
        continue
    o = os.path.join(d, name)
    t0 = time.perf_counter()
    rc = subprocess.run([binary, "-i", f, "-o", o, *extra], capture_output=True).returncode
    dt = time.perf_counter() - t0
    out[name] = {"seconds": dt, "reads_per_s": a.reads / dt, "rc": rc}
if "reference" in out:
    out["identical_output"] = filecmp.cmp(os.path.join(d, "b200", "ping-pong_signatures.tsv"), os.path.join(d, "reference", "ping-pong_signatures.tsv"), shallow=False)
    out["speedup"] = out["reference"]["seconds"] / out["b200"]["seconds"]
print(json.dumps(out))
