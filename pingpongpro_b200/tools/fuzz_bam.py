"""Mutation fuzz of the BAM decode paths under AddressSanitizer / UBSan (CPU only).

    python pingpongpro_b200/tools/fuzz_bam.py [seed] [iterations]

Builds a small driver (sequential walk and batch mode with 3 and 8 workers over the same file) with
-fsanitize=address,undefined, then feeds it a synthetic BAM whose record stream was mutated (bit flips, random bytes,
corrupted block_size words, deleted spans; re-wrapped into BGZF blocks of several sizes).  Every mode must agree: same
records when the file decodes, a read error otherwise, no sanitizer report.  tests/test_frontend.py runs a short,
unsanitised version of the same loop.
"""
import os, random, struct, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import test_frontend as tf  # BGZF helpers
from pingpongpro_b200 import _build

WORK = tempfile.mkdtemp(prefix="ppp_fuzz_")
DRIVER_SRC = os.path.join(WORK, "driver.cpp")
DRIVER = os.path.join(WORK, "driver")
open(DRIVER_SRC, "w").write('''
#include <cstdio>
#include "frontend.h"
int main(int argc, char** argv) {
    for (int i = 1; i < argc; ++i)
        for (int th : {1, 3, 8}) {
            ppp::DecodedFile d;
            ppp::ExtractOptions o;
            int rc = ppp::decode_file(argv[i], o, d, th);
            printf("%s threads %d rc %d records %llu kept %llu\\n", argv[i], th, rc, (unsigned long long)d.nRecords, (unsigned long long)d.nKept);
        }
}
''')
H = _build.HOST
subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-std=gnu++17", "-I", _build.INC, "-I", H, "-I", _build.TOOLS, DRIVER_SRC,
                       *[os.path.join(H, f) for f in ("aln_reader.cpp", "fast_inflate.cpp", "frontend.cpp", "annotation.cpp")], "-lz", "-lpthread", "-o", DRIVER])
src = os.path.join(WORK, 'src.bam')
subprocess.check_call([os.path.join(_build.BIN, "ppp_synth"), "--genome", "tiny", "--reads", "20000", "--seed", "5", "--out", src])
payload = bytearray(tf._bgzf_payload(src))
l_text = struct.unpack_from("<I", payload, 4)[0]; p = 8 + l_text; n_ref = struct.unpack_from("<I", payload, p)[0]; p += 4
for _ in range(n_ref):
    l = struct.unpack_from("<I", payload, p)[0]; p += 4 + l + 4
rec_start = p
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 100
bad = 0
for it in range(n_iter):
    m = bytearray(payload)
    kind = rng.randrange(4)
    for _ in range(rng.randrange(1, 4)):
        pos = rng.randrange(rec_start, len(m))
        if kind == 0: m[pos] ^= 1 << rng.randrange(8)
        elif kind == 1: m[pos] = rng.randrange(256)
        elif kind == 2:  # corrupt a block_size word specifically: walk to a random record
            q = rec_start
            for _ in range(rng.randrange(1, 15000)):
                nq = q + 4 + struct.unpack_from("<I", payload, q)[0]
                if nq + 4 > len(payload): break
                q = nq
            struct.pack_into("<I", m, q, rng.choice([0, 31, 32, 33, 90, 200, 5000, 70000, 1 << 20, 0xffffffff, struct.unpack_from("<I", payload, q)[0] + rng.choice([-1, 1, 2, 16])]) & 0xffffffff)
        else:
            a = rng.randrange(rec_start, len(m) - 300); del m[a:a + rng.randrange(1, 300)]
    f = os.path.join(WORK, 'm.bam')
    tf._bgzf_write(f, bytes(m), rng.choice([601, 7919, 65280]), level=rng.choice([0, 1, 6]))
    out = subprocess.run([DRIVER, f], capture_output=True, text=True)
    lines = [l.split(' threads ')[1] for l in out.stdout.splitlines() if ' threads ' in l]
    rcs = [l.split()[2] for l in lines]  # "<t> rc <rc> records <n> kept <k>"
    tails = [l.split(' rc ')[1] for l in lines]
    ok = out.returncode == 0 and len(lines) == 3 and len(set(rcs)) == 1 and (rcs[0] != '0' or len(set(tails)) == 1) and 'ERROR' not in out.stderr and 'runtime error' not in out.stderr
    if not ok:
        bad += 1
        print("MISMATCH iter", it, "kind", kind, out.stdout, out.stderr[-2000:])
        os.replace(f, os.path.join(WORK, f'bad_{it}.bam'))
print("iterations", n_iter, "mismatches", bad, ("(failing inputs kept in " + WORK + ")") if bad else "")
sys.exit(1 if bad else 0)
