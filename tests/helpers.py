"""Shared test helpers: golden-fixture loading and the reference's text-file conventions."""
import os

import numpy as np

from dissect_b200 import bedio
from oracle import grm_oracle as orc
from oracle import ref_runner

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

GOLDEN_GRM_CASES = ["tiny7_m40", "n37_m64", "n130_m300_keep_extract", "n258_m500_flat", "n64_m200_nomiss"]


def load_case(name):
    d = os.path.join(GOLDEN, name)
    meta = {}
    for line in open(os.path.join(d, "case.txt")):
        k, v = line.rstrip("\n").split(" ", 1)
        meta[k] = v
    n_fam, m_file = int(meta["n_fam"]), int(meta["m"])
    packed = bedio.read_bed(os.path.join(d, "in"), n_fam)
    assert packed.shape[0] == m_file
    fam = [l.split()[:2] for l in open(os.path.join(d, "in.fam"))]
    bim = [l.split()[1] for l in open(os.path.join(d, "in.bim"))]
    keep_mask = None
    if os.path.exists(os.path.join(d, "in.keep")):
        keep = {tuple(l.split()[:2]) for l in open(os.path.join(d, "in.keep"))}
        keep_mask = np.array([1 if tuple(x) in keep else 0 for x in fam], dtype=np.int32)
    snp_keep = None
    if os.path.exists(os.path.join(d, "in.extract")):
        ext = {l.strip() for l in open(os.path.join(d, "in.extract"))}
        snp_keep = np.array([1 if s in ext else 0 for s in bim], dtype=np.uint8)
    case = {"dir": d, "meta": meta, "n_fam": n_fam, "m_file": m_file, "packed": packed, "fam": fam, "bim": bim,
            "keep_mask": keep_mask, "snp_keep": snp_keep, "flat": bool(int(meta.get("flat", "0")))}
    n = n_fam if keep_mask is None else int(keep_mask.sum())
    case["n"] = n
    dat = os.path.join(d, "ref.grm.dat")
    if os.path.exists(dat):
        ids = [l.split() for l in open(os.path.join(d, "ref.grm.ids"))]
        case["ref_ids"] = ids
        case["ref_A"], case["ref_N"] = orc.unpack_grm_dat(open(dat, "rb").read(), len(ids))
        case["ref_dat"] = open(dat, "rb").read()
    sd = os.path.join(d, "ref.snpdump")
    if os.path.exists(sd):
        case["ref_stats"] = ref_runner.parse_snpdump(sd)
    return case


def kept_rows(case):
    pk = case["packed"]
    return pk if case["snp_keep"] is None else pk[case["snp_keep"] != 0]


def lower_mask(n):
    return np.tril(np.ones((n, n), bool))


def load_ld_case(name):
    """tests/golden/ld_*: inputs + the reference's correlation matrix and outside-range index lists."""
    d = os.path.join(GOLDEN, name)
    meta = dict(l.rstrip("\n").split(" ", 1) for l in open(os.path.join(d, "case.txt")))
    n_fam, m_file = int(meta["n_fam"]), int(meta["m"])
    packed = bedio.read_bed(os.path.join(d, "in"), n_fam)
    fam = [l.split()[:2] for l in open(os.path.join(d, "in.fam"))]
    bim = [l.split()[1] for l in open(os.path.join(d, "in.bim"))]
    keep_mask = snp_keep = None
    if os.path.exists(os.path.join(d, "in.keep")):
        keep = {tuple(l.split()[:2]) for l in open(os.path.join(d, "in.keep"))}
        keep_mask = np.array([1 if tuple(x) in keep else 0 for x in fam], dtype=np.int32)
    if os.path.exists(os.path.join(d, "in.extract")):
        ext = {l.strip() for l in open(os.path.join(d, "in.extract"))}
        snp_keep = np.array([1 if s in ext else 0 for s in bim], dtype=np.uint8)
    blob = open(os.path.join(d, "ref.ldcorr"), "rb").read()
    m = int(np.frombuffer(blob, np.int64, 1, 0)[0])
    C = np.frombuffer(blob, np.float64, m * m, 8).reshape((m, m), order="F").copy()
    C[np.triu_indices(m, 1)] = np.nan                       # pdsyrk_('L') leaves the upper triangle undefined
    cnt = int(np.frombuffer(blob, np.int64, 1, 8 + 8 * m * m)[0])
    pairs = np.frombuffer(blob, np.int32, 2 * cnt, 16 + 8 * m * m).reshape(cnt, 2)
    return {"dir": d, "n_fam": n_fam, "m_file": m_file, "packed": packed, "keep_mask": keep_mask, "snp_keep": snp_keep,
            "rows": packed if snp_keep is None else packed[snp_keep != 0], "threshold": float(meta["threshold"]),
            "ref_C": C, "ref_pairs": [(int(r), int(c)) for r, c in pairs]}
