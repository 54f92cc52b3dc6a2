"""PLINK .bed/.bim/.fam writers and the seeded synthetic cohort generator.

Data plumbing only (numpy): the generator produces the *inputs* both the CUDA path and the
oracle consume.  .bed layout follows genotype.cpp:578-605, 662, 752, 784 of the reference:
3 header bytes 6C 1B 01, then one row of ceil(n/4) bytes per SNP, individual f in byte f/4,
bits 2(f%4)..2(f%4)+1, patterns 00 hom A1 / 01 missing / 10 het / 11 hom A2.
"""
import os

import numpy as np

BED_MAGIC = bytes([0x6C, 0x1B, 0x01])
# allele-2 dosage 0/1/2 -> .bed bit pattern; 3 encodes "missing"
_DOSAGE_TO_PATTERN = np.array([0b00, 0b10, 0b11, 0b01], dtype=np.uint8)


def pack_patterns(pat):
    """pat: uint8 [m, n] of 2-bit patterns -> packed rows uint8 [m, ceil(n/4)] (padding bits 0)."""
    m, n = pat.shape
    b = (n + 3) // 4
    padded = np.zeros((m, b * 4), dtype=np.uint8)
    padded[:, :n] = pat
    q = padded.reshape(m, b, 4)
    return (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8)


def synth_packed(n, m, seed, maf_lo=0.05, maf_hi=0.5, missing=0.0, rows_per_block=4096, ensure_poly=True,
                 return_maf=False):
    """Seeded cohort: MAF_k ~ U(maf_lo, maf_hi), dosage ~ Binomial(2, MAF_k) (HWE), each genotype
    independently missing with probability `missing`.  Returns packed rows uint8 [m, ceil(n/4)]
    (and the per-SNP sample MAF = min(p1, p2) over non-missing genotypes, analysis.cpp:534, when
    return_maf is set — used to write the `--extract` list that realises a MAF filter).
    ensure_poly rewrites the first two individuals of a SNP that came out monomorphic so that
    standardDev > 0 (the reference aborts on such SNPs, genotype.cpp:924-939)."""
    rng = np.random.default_rng(seed)
    out = np.empty((m, (n + 3) // 4), dtype=np.uint8)
    mafs = np.empty(m, dtype=np.float64)
    for r0 in range(0, m, rows_per_block):
        r1 = min(m, r0 + rows_per_block)
        maf = rng.uniform(maf_lo, maf_hi, size=(r1 - r0, 1))
        u = rng.random((r1 - r0, n), dtype=np.float32)
        dos = (u < maf).astype(np.uint8)
        u = rng.random((r1 - r0, n), dtype=np.float32)
        dos += (u < maf).astype(np.uint8)
        if missing > 0:
            u = rng.random((r1 - r0, n), dtype=np.float32)
            dos[u < missing] = 3
        if ensure_poly and n >= 2:
            for k in range(r1 - r0):
                row = dos[k]
                nm = row[row != 3]
                if nm.size == 0 or nm.min() == nm.max():
                    row[0], row[1] = 0, 2
        present = dos != 3
        nn = np.maximum(present.sum(axis=1), 1)
        p = np.where(present, dos, 0).sum(axis=1) / (2.0 * nn)
        mafs[r0:r1] = np.minimum(p, 1.0 - p)
        out[r0:r1] = pack_patterns(_DOSAGE_TO_PATTERN[dos])
    return (out, mafs) if return_maf else out


def write_bed(prefix, packed, n, snp_ids=None, fam_ids=None, chrom=None):
    """Write <prefix>.bed/.bim/.fam for packed rows [m, ceil(n/4)]."""
    m = packed.shape[0]
    assert packed.shape[1] == (n + 3) // 4
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + ".bed", "wb") as f:
        f.write(BED_MAGIC)
        f.write(np.ascontiguousarray(packed).tobytes())
    with open(prefix + ".bim", "w") as f:
        for k in range(m):
            sid = snp_ids[k] if snp_ids is not None else f"rs{k + 1}"
            c = int(chrom[k]) if chrom is not None else 1
            f.write(f"{c}\t{sid}\t0\t{k + 1}\tA\tG\n")
    with open(prefix + ".fam", "w") as f:
        for i in range(n):
            fid, iid = fam_ids[i] if fam_ids is not None else (f"f{i + 1}", f"i{i + 1}")
            f.write(f"{fid} {iid} 0 0 0 -9\n")


def read_bed(prefix, n):
    """Read packed rows back (validates the magic like genotype.cpp:581-605)."""
    with open(prefix + ".bed", "rb") as f:
        hdr = f.read(3)
        if hdr[:2] != BED_MAGIC[:2]:
            raise ValueError("Error: The data file is not a valid binary BED > v1.0")
        if hdr[2] != 1:
            raise ValueError("not SNP-major")
        raw = np.frombuffer(f.read(), dtype=np.uint8)
    b = (n + 3) // 4
    return raw.reshape(-1, b)
