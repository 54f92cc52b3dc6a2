"""Generates tests/golden/ld_*: seeded .bed inputs with planted LD and what the REFERENCE ITSELF (oracle/_ref/ldcorr:
the unmodified reference objects running the call sequence of GWAS::getLessSignificantCorrelatedSNPs,
gwas.cpp:1163-1172) produced for them — the SNP x SNP correlation matrix and the index lists of
getGlobalIndexOutsideRange.  Run from the repo root in the build container:  python oracle/make_golden_ld.py
The fixtures are committed; the GPU box never runs this script."""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dissect_b200 import bedio  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
LDCORR = os.path.join(ROOT, "oracle", "_ref", "ldcorr")

# name, n_fam, m, seed, missing, keep_every, extract_every, threshold
CASES = [
    ("ld_n61_m40", 61, 40, 201, 0.03, 0, 0, 0.5),
    ("ld_n130_m90_keep_extract", 130, 90, 202, 0.02, 7, 6, 0.4),
    ("ld_n300_m270", 300, 270, 203, 0.02, 0, 0, 0.6),          # crosses the 256-row tile edge of the output
]


def plant_ld(pk, n, rng, pairs, flip=0.08):
    """row b := row a with a fraction `flip` of the genotypes redrawn from another SNP (correlation ~ 1 - 2 flip)."""
    m = pk.shape[0]
    for _ in range(pairs):
        a, b, c = rng.choice(m, 3, replace=False)
        for i in range(n):
            src = a if rng.random() > flip else c
            byte, sh = i // 4, 2 * (i % 4)
            v = (int(pk[src, byte]) >> sh) & 3
            pk[b, byte] = (int(pk[b, byte]) & (0xFF ^ (3 << sh))) | (v << sh)


def main():
    if not os.path.exists(LDCORR):
        raise SystemExit("oracle/_ref/ldcorr is not built: run `make -C oracle ref` first")
    for name, n, m, seed, missing, keep_every, extract_every, thr in CASES:
        tmp = tempfile.mkdtemp(prefix="golden_ld_")
        try:
            pk = bedio.synth_packed(n, m, seed, maf_lo=0.1, missing=missing)
            plant_ld(pk, n, np.random.default_rng(seed), max(3, m // 12))
            bf = os.path.join(tmp, name)
            bedio.write_bed(bf, pk, n)
            cmd = [LDCORR, "--bfile", bf, "--out", os.path.join(tmp, "ref")]
            if keep_every:
                with open(bf + ".keep", "w") as f:
                    f.write("".join(f"f{i + 1} i{i + 1}\n" for i in range(n) if (i + 1) % keep_every))
                cmd += ["--keep", bf + ".keep"]
            if extract_every:
                with open(bf + ".extract", "w") as f:
                    f.write("".join(f"rs{k + 1}\n" for k in range(m) if (k + 1) % extract_every))
                cmd += ["--extract", bf + ".extract"]
            r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, LDCORR_THRESHOLD=repr(thr)))
            if r.returncode != 0 or not os.path.exists(os.path.join(tmp, "ref.ldcorr")):
                raise SystemExit(f"{name}: ldcorr failed\n" + (r.stdout + r.stderr)[-3000:])
            dst = os.path.join(GOLD, name)
            os.makedirs(dst, exist_ok=True)
            for ext in (".bed", ".bim", ".fam", ".keep", ".extract"):
                if os.path.exists(bf + ext):
                    shutil.copy(bf + ext, os.path.join(dst, "in" + ext))
            shutil.copy(os.path.join(tmp, "ref.ldcorr"), os.path.join(dst, "ref.ldcorr"))
            with open(os.path.join(dst, "case.txt"), "w") as f:
                f.write(f"n_fam {n}\nm {m}\nseed {seed}\nmissing {missing}\nkeep_every {keep_every}\n"
                        f"extract_every {extract_every}\nthreshold {thr}\n")
            print(f"{name}: ok")
        finally:
            shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
