"""BGEN container plumbing for the dosage ingest (dgrm_push_dosage_host).

The reference reads BGEN through the `genfile::bgen` library (genotypebgen.cpp:20-22, 284-559: ReadBGENData wraps
genfile::bgen::read_header_block / read_snp_identifying_data / read_and_parse_genotype_data_block), which is not
vendored in the reference tree.  This module restates the published BGEN v1.1 / v1.2 layout for exactly the subset the
reference accepts -- biallelic, diploid, unphased variants (ProbabilitiesSetter::initialise / set_min_max_ploidy /
set_number_of_entries, genotypebgen.cpp:99-172) -- and hands each variant on as the row the reference's setter forms:
    dosage_i  = P_i(AB) + 2 * P_i(BB)                       (set_value, genotypebgen.cpp:199-205)
    missing_i = layout 1: all three probabilities zero       (genotypebgen.cpp:208-214)
                layout 2: the sample's missingness bit       (set_value(MissingValue), genotypebgen.cpp:229-236)
Everything numerical after that (mean, std, standardisation, S, N) happens on the GPU.

Supported: layout 1 (16-bit probabilities / 32768) and layout 2 (B-bit probabilities / (2^B - 1), B = 1..32),
uncompressed or zlib.  zstd-compressed files are refused (no zstd in this image).  `write_bgen` produces such files
for the tests.
"""
import struct
import zlib

import numpy as np


class BgenError(ValueError):
    pass


class BgenHeader:
    def __init__(self, offset, n_variants, n_samples, compression, layout, sample_ids):
        self.offset, self.n_variants, self.n_samples = offset, n_variants, n_samples
        self.compression, self.layout, self.sample_ids = compression, layout, sample_ids


def _read_header(f):
    (offset,) = struct.unpack("<I", f.read(4))
    l_h, m, n = struct.unpack("<III", f.read(12))
    magic = f.read(4)
    if magic not in (b"bgen", b"\0\0\0\0"):
        raise BgenError("Error: The file is not a valid BGEN file (bad magic number).")
    f.read(l_h - 20)
    (flags,) = struct.unpack("<I", f.read(4))
    compression, layout = flags & 3, (flags >> 2) & 15
    if layout not in (1, 2):
        raise BgenError("Error: Unsupported BGEN layout %d." % layout)
    if compression == 2:
        raise BgenError("Error: zstd-compressed BGEN files are not supported here.")
    ids = None
    if flags >> 31:
        _l_si, n2 = struct.unpack("<II", f.read(8))
        if n2 != n:
            raise BgenError("Error: The BGEN sample identifier block disagrees with the header.")
        ids = []
        for _ in range(n):
            (ln,) = struct.unpack("<H", f.read(2))
            ids.append(f.read(ln).decode())
    f.seek(offset + 4)
    return BgenHeader(offset, m, n, compression, layout, ids)


def _read_str(f, fmt):
    (ln,) = struct.unpack(fmt, f.read(struct.calcsize(fmt)))
    return f.read(ln).decode()


def _unpack_bits(buf, count, bits):
    """`count` little-endian unsigned integers of `bits` bits each, packed back to back."""
    if bits == 8:
        return np.frombuffer(buf, dtype=np.uint8, count=count).astype(np.uint64)
    if bits == 16:
        return np.frombuffer(buf, dtype="<u2", count=count).astype(np.uint64)
    if bits == 32:
        return np.frombuffer(buf, dtype="<u4", count=count).astype(np.uint64)
    raw = np.frombuffer(buf, dtype=np.uint8, count=(count * bits + 7) // 8)
    b = np.unpackbits(raw, bitorder="little")[: count * bits].reshape(count, bits).astype(np.uint64)
    return (b << np.arange(bits, dtype=np.uint64)).sum(axis=1, dtype=np.uint64)


def _variant_blocks(f, hdr):
    """Yields (chromosome, position, rsid, alleles, block) per variant in file order; block = the inflated
    probability data (layout 2: from the N field on; layout 1: the 6N bytes), a bytes object."""
    n = hdr.n_samples
    for _ in range(hdr.n_variants):
        if hdr.layout == 1:
            (n_here,) = struct.unpack("<I", f.read(4))
            if n_here != n:
                raise BgenError("Error: Unexpected number of samples in a BGEN variant block.")
        _varid = _read_str(f, "<H")
        rsid = _read_str(f, "<H")
        chrom = _read_str(f, "<H")
        (pos,) = struct.unpack("<I", f.read(4))
        k = 2
        if hdr.layout == 2:
            (k,) = struct.unpack("<H", f.read(2))
        alleles = [_read_str(f, "<I") for _ in range(k)]
        if k != 2:  # genotypebgen.cpp:107-110
            raise BgenError("Error: The bgen file contains non-biallelic genotypes. DISSECT currently can only work "
                            "with biallelic genotypes.")
        if hdr.layout == 1:
            if hdr.compression:
                (c,) = struct.unpack("<I", f.read(4))
                data = zlib.decompress(f.read(c))
            else:
                data = f.read(6 * n)
        else:
            (c,) = struct.unpack("<I", f.read(4))
            if hdr.compression:
                (d,) = struct.unpack("<I", f.read(4))
                data = zlib.decompress(f.read(c - 4))
                if len(data) != d:
                    raise BgenError("Error: A BGEN genotype block does not decompress to its stated length.")
            else:
                data = f.read(c)
        yield chrom, pos, rsid, alleles, data


def _variant_rows(f, hdr):
    """Yields (chromosome, position, rsid, alleles, dosage[n], missing[n]) per variant, file order."""
    n = hdr.n_samples
    for chrom, pos, rsid, alleles, data in _variant_blocks(f, hdr):
        if hdr.layout == 1:
            v = np.frombuffer(data, dtype="<u2", count=3 * n).reshape(n, 3).astype(np.float64) / 32768.0
            missing = (v[:, 0] + v[:, 1] + v[:, 2]) == 0.0
            dosage = v[:, 1] + 2.0 * v[:, 2]
        else:
            n_here, k2, pmin, pmax = struct.unpack_from("<IHBB", data, 0)
            if n_here != n or k2 != 2:
                raise BgenError("Error: Unexpected sample or allele count in a BGEN genotype block.")
            if pmin != 2 or pmax != 2:  # genotypebgen.cpp:120-123
                raise BgenError("Error: The bgen file contains non-diploid genotypes. DISSECT currently can only work "
                                "with diploid genotypes.")
            ploidy = np.frombuffer(data, dtype=np.uint8, count=n, offset=8)
            phased, bits = struct.unpack_from("<BB", data, 8 + n)
            if phased:  # genotypebgen.cpp:168-171
                raise BgenError("Error: Currently can only load unphased genotypes. If you are interested on loading "
                                "phased genotypes, please, contact us.")
            if np.any((ploidy & 0x3F) != 2):
                raise BgenError("Error: The bgen file contains non-diploid genotypes. DISSECT currently can only work "
                                "with diploid genotypes.")
            missing = (ploidy & 0x80) != 0
            ints = _unpack_bits(data[10 + n:], 2 * n, bits).reshape(n, 2)
            scale = float((1 << bits) - 1)
            p_aa = ints[:, 0].astype(np.float64) / scale
            p_ab = ints[:, 1].astype(np.float64) / scale
            p_bb = 1.0 - (p_aa + p_ab)       # the implied last probability of the format
            dosage = p_ab + 2.0 * p_bb
        yield chrom, pos, rsid, alleles, dosage, missing


def read_header(path):
    with open(path, "rb") as f:
        return _read_header(f)


def iter_dosage_blocks(path, keep_mask=None, keep_snps=None, block=1024):
    """Blocks of variants as (meta, dosage[rows, n_kept] float64, missing[rows, n_kept] uint8); meta is a list of
    (chromosome, position, rsid, allele1, allele2).  keep_mask: one flag per sample of the file
    (individualBEDIdxsMaskToKeep); keep_snps: set of rsids to keep (--extract, genotypebgen.cpp:655-659)."""
    with open(path, "rb") as f:
        hdr = _read_header(f)
        cols = None if keep_mask is None else np.nonzero(np.asarray(keep_mask) != 0)[0]
        meta, dos, mis = [], [], []
        for chrom, pos, rsid, alleles, d, q in _variant_rows(f, hdr):
            if keep_snps is not None and rsid not in keep_snps:
                continue
            meta.append((chrom, pos, rsid, alleles[0], alleles[1]))
            dos.append(d if cols is None else d[cols])
            mis.append(q if cols is None else q[cols])
            if len(meta) == block:
                yield meta, np.vstack(dos), np.vstack(mis).astype(np.uint8)
                meta, dos, mis = [], [], []
        if meta:
            yield meta, np.vstack(dos), np.vstack(mis).astype(np.uint8)


def iter_raw_blocks(path, keep_snps=None, block=1024):
    """Blocks of variants as (meta, data uint8[...], offsets int64[rows + 1], header): the inflated probability blocks
    back to back, for GrmEngine.push_bgen_blocks_host (the device unpacks them)."""
    with open(path, "rb") as f:
        hdr = _read_header(f)
        meta, parts = [], []
        for chrom, pos, rsid, alleles, data in _variant_blocks(f, hdr):
            if keep_snps is not None and rsid not in keep_snps:
                continue
            meta.append((chrom, pos, rsid, alleles[0], alleles[1]))
            parts.append(data)
            if len(meta) == block:
                off = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.int64)
                yield meta, np.frombuffer(b"".join(parts), dtype=np.uint8), off, hdr
                meta, parts = [], []
        if meta:
            off = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.int64)
            yield meta, np.frombuffer(b"".join(parts), dtype=np.uint8), off, hdr


def push_bgen(engine, path, keep_mask=None, keep_snps=None, block=1024, device_unpack=True):
    """Feed every kept variant of a BGEN file to an engine that has been begun; returns the variants' metadata.
    device_unpack: hand the inflated probability blocks to the device (dgrm_push_bgen_blocks_host); otherwise the rows
    are unpacked here and pushed as dosages (dgrm_push_dosage_host).  Both give bit-identical results."""
    out = []
    if device_unpack:
        for meta, data, offsets, hdr in iter_raw_blocks(path, keep_snps, block):
            engine.push_bgen_blocks_host(data, offsets, hdr.layout, hdr.n_samples, keep_mask)
            out.extend(meta)
        return out
    for meta, dosage, missing in iter_dosage_blocks(path, keep_mask, keep_snps, block):
        engine.push_dosage_host(dosage, missing if missing.any() else None)
        out.extend(meta)
    return out


def write_bgen(path, p_aa, p_ab, missing=None, layout=2, bits=8, compressed=True, sample_ids=None, rsids=None):
    """Test-fixture writer.  p_aa, p_ab: integer arrays [m, n] of stored probability numerators (layout 2: value /
    (2^bits - 1); layout 1: value / 32768, P(BB) numerator = 32768 - p_aa - p_ab); missing: bool [m, n]."""
    p_aa, p_ab = np.asarray(p_aa, dtype=np.uint64), np.asarray(p_ab, dtype=np.uint64)
    m, n = p_aa.shape
    missing = np.zeros((m, n), bool) if missing is None else np.asarray(missing, bool)
    flags = (1 if compressed else 0) | (layout << 2) | ((1 << 31) if sample_ids is not None else 0)
    header = struct.pack("<III", 20, m, n) + b"bgen" + struct.pack("<I", flags)
    sample_block = b""
    if sample_ids is not None:
        body = b"".join(struct.pack("<H", len(s)) + s.encode() for s in sample_ids)
        sample_block = struct.pack("<II", len(body) + 8, n) + body
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(header) + len(sample_block)))
        f.write(header)
        f.write(sample_block)
        for k in range(m):
            rs = (rsids[k] if rsids is not None else "rs%d" % k).encode()
            ident = b""
            if layout == 1:
                ident += struct.pack("<I", n)
            ident += struct.pack("<H", len(rs)) + rs + struct.pack("<H", len(rs)) + rs + struct.pack("<H", 1) + b"1"
            ident += struct.pack("<I", 1000 + k)
            if layout == 2:
                ident += struct.pack("<H", 2)
            ident += struct.pack("<I", 1) + b"A" + struct.pack("<I", 1) + b"G"
            f.write(ident)
            if layout == 1:
                v = np.zeros((n, 3), dtype="<u2")
                v[:, 0], v[:, 1] = p_aa[k], p_ab[k]
                v[:, 2] = 32768 - p_aa[k] - p_ab[k]
                v[missing[k]] = 0
                data = v.tobytes()
                if compressed:
                    z = zlib.compress(data)
                    f.write(struct.pack("<I", len(z)) + z)
                else:
                    f.write(data)
            else:
                ploidy = np.full(n, 2, dtype=np.uint8)
                ploidy[missing[k]] |= 0x80
                vals = np.stack([p_aa[k], p_ab[k]], axis=1).reshape(-1)
                vals = np.where(np.repeat(missing[k], 2), 0, vals).astype(np.uint64)
                if bits in (8, 16, 32):
                    packed = vals.astype({8: np.uint8, 16: "<u2", 32: "<u4"}[bits]).tobytes()
                else:
                    b = ((vals[:, None] >> np.arange(bits, dtype=np.uint64)) & 1).astype(np.uint8).reshape(-1)
                    packed = np.packbits(b, bitorder="little").tobytes()
                data = struct.pack("<IHBB", n, 2, 2, 2) + ploidy.tobytes() + struct.pack("<BB", 0, bits) + packed
                if compressed:
                    z = zlib.compress(data)
                    f.write(struct.pack("<II", len(z) + 4, len(data)) + z)
                else:
                    f.write(struct.pack("<I", len(data)) + data)
