"""Generates tests/golden/*: small seeded .bed inputs and what the REFERENCE ITSELF (oracle/_ref, built from
the unmodified sources by oracle/Makefile) produced for them — .grm.dat, .grm.ids/.snps, the snpdump of
`struct SNP`, and for the tiniest case the standardised genotype matrix.  Run from the repo root in the
build container:  python oracle/make_golden.py
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
from oracle import ref_runner  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name, n_fam, m, seed, missing, maf_lo, keep_every (drop every k-th individual), extract_every, flat
CASES = [
    ("tiny7_m40", 7, 40, 101, 0.05, 0.2, 0, 0, False),     # 7 individuals like the reference's absent test/test3, enough SNPs to pass sanitizeKernel
    ("n37_m64", 37, 64, 102, 0.05, 0.05, 0, 0, False),      # n % 4 == 1
    ("n130_m300_keep_extract", 130, 300, 103, 0.02, 0.05, 7, 5, False),
    ("n258_m500_flat", 258, 500, 104, 0.01, 0.05, 0, 0, True),  # n % 4 == 2, crosses one 256 panel, --grm-flat-norm
    ("n64_m200_nomiss", 64, 200, 105, 0.0, 0.05, 0, 0, False),
]


def main():
    if not ref_runner.available():
        raise SystemExit("oracle/_ref is not built: run `make -C oracle ref` first")
    os.makedirs(GOLD, exist_ok=True)
    for name, n, m, seed, missing, maf_lo, keep_every, extract_every, flat in CASES:
        tmp = tempfile.mkdtemp(prefix="golden_")
        try:
            pk = bedio.synth_packed(n, m, seed, maf_lo=maf_lo, missing=missing)
            bf = os.path.join(tmp, name)
            bedio.write_bed(bf, pk, n)
            keep = extract = None
            if keep_every:
                keep = bf + ".keep"
                with open(keep, "w") as f:
                    for i in range(n):
                        if (i + 1) % keep_every != 0:
                            f.write(f"f{i + 1} i{i + 1}\n")
            if extract_every:
                extract = bf + ".extract"
                with open(extract, "w") as f:
                    for k in range(m):
                        if (k + 1) % extract_every != 0:
                            f.write(f"rs{k + 1}\n")
            out = os.path.join(tmp, "ref")
            rc, log = ref_runner.run_make_grm(bf, out, extract=extract, keep=keep, flat_norm=flat)
            if rc != 0:
                raise SystemExit(f"{name}: reference failed\n{log[-3000:]}")
            zpath = os.path.join(tmp, "z.bin") if n * m <= 4096 else None
            if name == "n64_m200_nomiss":
                pass
            ref_runner.run_snpdump(bf, out, extract=extract, keep=keep, dump_z=zpath)
            dst = os.path.join(GOLD, name)
            os.makedirs(dst, exist_ok=True)
            for ext in (".bed", ".bim", ".fam"):
                shutil.copy(bf + ext, os.path.join(dst, "in" + ext))
            if keep:
                shutil.copy(keep, os.path.join(dst, "in.keep"))
            if extract:
                shutil.copy(extract, os.path.join(dst, "in.extract"))
            for ext in (".grm.dat", ".grm.ids", ".grm.snps", ".snpdump"):
                shutil.copy(out + ext, os.path.join(dst, "ref" + ext))
            if zpath:
                shutil.copy(zpath, os.path.join(dst, "ref.z.bin"))
            with open(os.path.join(dst, "case.txt"), "w") as f:
                f.write(f"n_fam {n}\nm {m}\nseed {seed}\nmissing {missing}\nmaf_lo {maf_lo}\n"
                        f"keep_every {keep_every}\nextract_every {extract_every}\nflat {int(flat)}\n")
            print(f"{name}: ok ({sum(os.path.getsize(os.path.join(dst, x)) for x in os.listdir(dst))} bytes)")
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    # sanitizeKernel fixtures (kernel.cpp:1993-2037): (a) two individuals with ~90% missing genotypes get removed,
    # (b) a 7 x 7 cohort where every pair overlaps on < 10 SNPs: the reference skips the GRM altogether.
    tmp = tempfile.mkdtemp(prefix="golden_")
    try:
        n, m = 40, 60
        pk = bedio.synth_packed(n, m, 107, maf_lo=0.2, missing=0.02)
        rng = np.random.default_rng(1070)
        codes_missing = 0b01
        for ind in (3, 17):
            for k in range(m):
                if rng.random() < 0.9:
                    byte, sh = ind // 4, 2 * (ind % 4)
                    pk[k, byte] = (int(pk[k, byte]) & (0xFF ^ (3 << sh))) | (codes_missing << sh)
        bf = os.path.join(tmp, "san")
        bedio.write_bed(bf, pk, n)
        out = os.path.join(tmp, "ref")
        rc, log = ref_runner.run_make_grm(bf, out)
        if rc != 0 or not os.path.exists(out + ".grm.dat"):
            raise SystemExit("sanitize case failed\n" + log[-3000:])
        dst = os.path.join(GOLD, "sanitize_remove")
        os.makedirs(dst, exist_ok=True)
        for ext in (".bed", ".bim", ".fam"):
            shutil.copy(bf + ext, os.path.join(dst, "in" + ext))
        for ext in (".grm.dat", ".grm.ids", ".grm.snps"):
            shutil.copy(out + ext, os.path.join(dst, "ref" + ext))
        with open(os.path.join(dst, "case.txt"), "w") as f:
            f.write(f"n_fam {n}\nm {m}\n")
        print("sanitize_remove: kept", len(open(out + ".grm.ids").read().split("\n")) - 1, "of", n)
        n, m = 7, 7
        pk = bedio.synth_packed(n, m, 108, maf_lo=0.2, missing=0.1)
        bf = os.path.join(tmp, "rej")
        bedio.write_bed(bf, pk, n)
        out = os.path.join(tmp, "ref2")
        rc, log = ref_runner.run_make_grm(bf, out)
        dst = os.path.join(GOLD, "sanitize_reject")
        os.makedirs(dst, exist_ok=True)
        for ext in (".bed", ".bim", ".fam"):
            shutil.copy(bf + ext, os.path.join(dst, "in" + ext))
        with open(os.path.join(dst, "case.txt"), "w") as f:
            f.write(f"n_fam {n}\nm {m}\ngrm_written {int(os.path.exists(out + '.grm.dat'))}\n")
            f.write("log_line " + [l for l in log.splitlines() if "Skipping this GRM" in l][0] + "\n")
        print("sanitize_reject: grm written?", os.path.exists(out + ".grm.dat"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    # partition fixtures (the reference's test20 / test57 invariants): GRMs of SNP subsets a, b and of all SNPs
    tmp = tempfile.mkdtemp(prefix="golden_")
    try:
        n, m = 48, 160
        pk = bedio.synth_packed(n, m, 109, maf_lo=0.1, missing=0.03)
        bf = os.path.join(tmp, "p")
        bedio.write_bed(bf, pk, n)
        dst = os.path.join(GOLD, "parts")
        os.makedirs(dst, exist_ok=True)
        for name, rng in (("a", range(0, 70)), ("b", range(70, 160)), ("all", range(0, 160))):
            ext = os.path.join(tmp, name + ".extract")
            with open(ext, "w") as f:
                f.write("".join(f"rs{k + 1}\n" for k in rng))
            rc, log = ref_runner.run_make_grm(bf, os.path.join(tmp, name), extract=ext)
            if rc != 0:
                raise SystemExit("parts case failed\n" + log[-3000:])
            for e in (".grm.dat", ".grm.ids", ".grm.snps"):
                shutil.copy(os.path.join(tmp, name + e), os.path.join(dst, name + e))
        for e in (".bed", ".bim", ".fam"):
            shutil.copy(bf + e, os.path.join(dst, "in" + e))
        # the reference's own (S, N) addition: `--bfile-list` with `--grm-join-method 0` computes one unnormalised
        # kernel per genotype file, joins them with Kernel::addKernels and normalises (auxiliar.cpp:620-672)
        bim = open(bf + ".bim").read().splitlines()
        for name, (a, b) in (("pa", (0, 70)), ("pb", (70, 160))):
            sub = os.path.join(tmp, name)
            bedio.write_bed(sub, pk[a:b], n)
            shutil.copy(bf + ".fam", sub + ".fam")
            with open(sub + ".bim", "w") as f:
                f.write("\n".join(bim[a:b]) + "\n")
        with open(os.path.join(tmp, "list.txt"), "w") as f:
            f.write(os.path.join(tmp, "pa") + "\n" + os.path.join(tmp, "pb") + "\n")
        import subprocess
        r = subprocess.run([ref_runner.DISSECT, "--make-grm", "--bfile-list", os.path.join(tmp, "list.txt"),
                            "--grm-join-method", "0", "--out", os.path.join(tmp, "joined")], capture_output=True, text=True)
        if r.returncode != 0 or "then add GRMs" not in r.stdout:
            raise SystemExit("parts/joined case failed\n" + (r.stdout + r.stderr)[-3000:])
        for e in (".grm.dat", ".grm.ids", ".grm.snps"):
            shutil.copy(os.path.join(tmp, "joined" + e), os.path.join(dst, "joined" + e))
        # the reference's own LOCO subtraction and file-level addition (Kernel::addKernels, kernel.cpp:1705-1859;
        # singlereml.cpp:298) on the files above, through oracle/_ref/grmalg (same unmodified reference objects)
        grmalg = os.path.join(os.path.dirname(ref_runner.DISSECT), "grmalg")
        for op, x, y, name in (("sub", "all", "a", "loco_a"), ("sub", "all", "b", "loco_b"), ("add", "a", "b", "sum_ab")):
            r = subprocess.run([grmalg, op, os.path.join(dst, x), os.path.join(dst, y), "--out", os.path.join(tmp, name)],
                               capture_output=True, text=True)
            if r.returncode != 0:
                raise SystemExit(f"parts/{name} failed\n" + (r.stdout + r.stderr)[-3000:])
            for e in (".grm.dat", ".grm.ids", ".grm.snps"):
                shutil.copy(os.path.join(tmp, name + e), os.path.join(dst, name + e))
        with open(os.path.join(dst, "case.txt"), "w") as f:
            f.write(f"n_fam {n}\nm {m}\nsplit 70\n")
        print("parts: ok")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    # block-cyclic index arithmetic of the reference itself (Matrix::local2global / global2local, matrix.cpp:435-488)
    # for emulated processes of R x C grids, through oracle/_ref/layoutdump
    from dissect_b200.engine import numroc
    layoutdump = os.path.join(os.path.dirname(ref_runner.DISSECT), "layoutdump")
    dst = os.path.join(GOLD, "layout")
    os.makedirs(dst, exist_ok=True)
    with open(os.path.join(dst, "layout.txt"), "w") as out:
        for (n, bs, grid) in ((10, (3, 2), (2, 3)), (37, (4, 4), (2, 2)), (130, (16, 8), (3, 2)), (700, (256, 256), (1, 1)),
                              (515, (100, 64), (1, 2)), (258, (7, 5), (3, 1))):
            for r in range(grid[0]):
                for c in range(grid[1]):
                    nlr, nlc = numroc(n, bs[0], r, grid[0]), numroc(n, bs[1], c, grid[1])
                    res = subprocess.run([layoutdump] + [str(x) for x in (n, bs[0], bs[1], grid[0], grid[1], r, c, nlr, nlc)],
                                         capture_output=True, text=True)
                    if res.returncode != 0:
                        raise SystemExit("layoutdump failed\n" + (res.stdout + res.stderr)[-2000:])
                    out.write(f"CASE {n} {bs[0]} {bs[1]} {grid[0]} {grid[1]} {r} {c} {nlr} {nlc}\n")
                    out.write("".join(l + "\n" for l in res.stdout.splitlines() if l[:2] in ("R ", "C ", "G ")))
    print("layout: ok")
    # error-path fixture: a monomorphic SNP makes the reference abort and write <out>.badsnps
    tmp = tempfile.mkdtemp(prefix="golden_")
    try:
        n, m = 20, 12
        pk = bedio.synth_packed(n, m, 106, maf_lo=0.2, missing=0.0)
        pk[5, :] = 0x00   # SNP rs6: everyone hom A1 -> standardDev == 0
        pk[9, :] = 0xFF   # SNP rs10: everyone hom A2
        bf = os.path.join(tmp, "bad")
        bedio.write_bed(bf, pk, n)
        out = os.path.join(tmp, "ref")
        rc, log = ref_runner.run_make_grm(bf, out)
        dst = os.path.join(GOLD, "badsnp")
        os.makedirs(dst, exist_ok=True)
        for ext in (".bed", ".bim", ".fam"):
            shutil.copy(bf + ext, os.path.join(dst, "in" + ext))
        shutil.copy(out + ".badsnps", os.path.join(dst, "ref.badsnps"))
        with open(os.path.join(dst, "case.txt"), "w") as f:
            f.write(f"n_fam {n}\nm {m}\nreference_returncode_nonzero {int(rc != 0)}\n")
        print("badsnp: reference rc", rc, "badsnps:", open(out + ".badsnps").read().split())
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
