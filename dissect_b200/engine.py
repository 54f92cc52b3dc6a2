"""Python face of the C ABI (include/dgrm.h): one `GrmEngine` per GPU / per process.

This is plumbing for tests, bench.py and the multi-GPU driver — all arithmetic happens in
libdgrm.so's sm_100a kernels.  Method names follow the C entry points one to one.
"""
import ctypes as C

import numpy as np

from . import _lib

lib_path = _lib.LIB_PATH


class DgrmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dgrm error {code}: {msg}")
        self.code = code
        self.msg = msg


def numroc(n, nb, iproc, nprocs):
    """ScaLAPACK numroc_ with source process 0 (reference: matrix.cpp:1748-1749)."""
    nblocks = n // nb
    r = (nblocks // nprocs) * nb
    extra = nblocks % nprocs
    if iproc < extra:
        r += nb
    elif iproc == extra:
        r += n % nb
    return r


class PinnedArray:
    """numpy view over cudaHostAlloc memory (dgrm_host_alloc); freed with the object."""

    def __init__(self, shape, dtype, order="C"):
        self._lib = _lib.load()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = self._lib.dgrm_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise MemoryError(f"dgrm_host_alloc({self.nbytes}) failed")
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape, order=order)

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                self.array = None
                self._lib.dgrm_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class GrmEngine:
    def __init__(self, device=0, rank=0, world=1, operand="fp16", count_method="fp4", flat_norm=False,
                 chunk_snps=16384, cta_group=2, profile=False, resident_chunks=0, resident_mb=16384,
                 reserve_sms=-1, a_transport="fp64", n_transport="narrow"):
        self._lib = _lib.load()
        cfg = _lib.Config()
        self._lib.dgrm_default_config(C.byref(cfg))
        cfg.device, cfg.rank, cfg.world = device, rank, world
        cfg.operand = _lib.OPERANDS[operand]
        cfg.count_method = _lib.COUNT_METHODS[count_method]
        cfg.flat_norm = int(bool(flat_norm))
        cfg.chunk_snps = chunk_snps
        cfg.cta_group = cta_group
        cfg.profile = int(bool(profile))
        cfg.resident_chunks = resident_chunks
        cfg.resident_mb = resident_mb
        cfg.reserve_sms = reserve_sms
        cfg.a_transport = _lib.TRANSPORT[a_transport]
        cfg.n_transport = _lib.TRANSPORT[n_transport]
        self.cfg = cfg
        h = C.c_void_p()
        rc = self._lib.dgrm_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise DgrmError(rc, self._lib.dgrm_last_error(None).decode())
        self._h = h
        self._keep = []

    # -- helpers -------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise DgrmError(rc, self._lib.dgrm_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dgrm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- C ABI, one to one ---------------------------------------------------------------------
    def set_individuals(self, n_fam, keep_mask=None):
        km = None
        if keep_mask is not None:
            km = np.ascontiguousarray(keep_mask, dtype=np.int32)
            assert km.shape == (n_fam,)
        self._ck(self._lib.dgrm_set_individuals(self._h, n_fam, km.ctypes.data if km is not None else None))

    def ld_begin(self, n_fam, m_group, keep_mask=None):
        """LD mode: the following pushes deliver the m_group SNPs whose correlations are wanted (gwas.cpp:1156-1172)."""
        km = None
        if keep_mask is not None:
            km = np.ascontiguousarray(keep_mask, dtype=np.int32)
            assert km.shape == (n_fam,)
        self._ck(self._lib.dgrm_ld_begin(self._h, int(n_fam), km.ctypes.data if km is not None else None, int(m_group)))

    def ld_pairs(self, threshold, cap=1 << 20):
        pairs = np.zeros((cap, 2), np.int64)
        cnt = C.c_int64(0)
        self._ck(self._lib.dgrm_ld_pairs(self._h, float(threshold), pairs.ctypes.data, cap, C.byref(cnt)))
        return sorted(map(tuple, pairs[: min(cap, cnt.value)].tolist())), cnt.value

    def begin(self, m_capacity):
        self._ck(self._lib.dgrm_begin(self._h, int(m_capacity)))

    def push_bed_host(self, rows, row_keep=None, pitch=None, nrows=None):
        """rows: uint8 array [nrows, pitch] (C order) or a raw address with explicit pitch/nrows."""
        rk = None
        if row_keep is not None:
            rk = np.ascontiguousarray(row_keep, dtype=np.uint8)
        if isinstance(rows, np.ndarray):
            assert rows.dtype == np.uint8 and rows.ndim == 2 and rows.strides[1] == 1
            ptr, pitch, nrows = rows.ctypes.data, rows.strides[0], rows.shape[0]
        else:
            ptr = int(rows)
        self._ck(self._lib.dgrm_push_bed_host(self._h, ptr, pitch, nrows, rk.ctypes.data if rk is not None else None))
        self._keep.append((rows, rk))  # the copy is asynchronous: keep the sources alive

    def push_dosage_host(self, dosage, missing=None):
        """dosage: float64 [nrows, n] expected allele counts of the kept individuals; missing: uint8/bool [nrows, n],
        non-zero where the sample has no probabilities (the rows the reference's BGEN reader forms,
        genotypebgen.cpp:183-236).  The rows are copied before the call returns."""
        d = np.ascontiguousarray(dosage, dtype=np.float64)
        assert d.ndim == 2
        mk = None
        if missing is not None:
            mk = np.ascontiguousarray(np.asarray(missing) != 0, dtype=np.uint8)
            assert mk.shape == d.shape
        self._ck(self._lib.dgrm_push_dosage_host(self._h, d.ctypes.data, d.shape[1], mk.ctypes.data if mk is not None else None,
                                                 d.shape[1], d.shape[0]))

    def push_bgen_blocks_host(self, data, offsets, layout, n_samples, sample_keep=None):
        """data: uint8 array holding the inflated probability blocks of BGEN variants back to back; offsets: int64
        [nrows + 1]; sample_keep: one flag per sample of the file (None = all).  The device unpacks the blocks."""
        d = np.ascontiguousarray(data, dtype=np.uint8)
        o = np.ascontiguousarray(offsets, dtype=np.int64)
        sk = None if sample_keep is None else np.ascontiguousarray(np.asarray(sample_keep) != 0, dtype=np.int32)
        assert sk is None or sk.shape == (n_samples,)
        self._ck(self._lib.dgrm_push_bgen_blocks_host(self._h, d.ctypes.data, o.ctypes.data, o.shape[0] - 1, int(layout),
                                                      int(n_samples), sk.ctypes.data if sk is not None else None))

    def push_bed_device(self, ptr, pitch, nrows, row_keep=None, stream=None):
        rk = None
        if row_keep is not None:
            rk = np.ascontiguousarray(row_keep, dtype=np.uint8)
        self._ck(self._lib.dgrm_push_bed_device(self._h, int(ptr), pitch, nrows,
                                                rk.ctypes.data if rk is not None else None, stream))

    def release_inputs(self, stream=None):
        """`stream` waits until the engine has read every buffer pushed so far (no host synchronisation)."""
        self._ck(self._lib.dgrm_release_inputs(self._h, stream))

    def load_bed_file(self, path, m_file, snp_keep_idx=None):
        idx = None
        if snp_keep_idx is not None:
            idx = np.ascontiguousarray(snp_keep_idx, dtype=np.int64)
        self._ck(self._lib.dgrm_load_bed_file(self._h, str(path).encode(), m_file,
                                              idx.ctypes.data if idx is not None else None,
                                              len(idx) if idx is not None else 0))

    # ---- resident cohort + (S, N) algebra ------------------------------------------------------------
    def store_bed_file(self, path, m_file):
        self._ck(self._lib.dgrm_store_bed_file(self._h, str(path).encode(), int(m_file)))

    def store_bed_host(self, rows):
        assert rows.dtype == np.uint8 and rows.ndim == 2 and rows.strides[1] == 1
        self._ck(self._lib.dgrm_store_bed_host(self._h, rows.ctypes.data, rows.strides[0], rows.shape[0]))

    def store_bed_device(self, ptr, pitch, nrows, stream=None):
        self._ck(self._lib.dgrm_store_bed_device(self._h, int(ptr), int(pitch), int(nrows), stream))

    @property
    def last_d2h_bytes(self):
        return int(self._lib.dgrm_last_d2h_bytes(self._h))

    def store_clear(self):
        self._ck(self._lib.dgrm_store_clear(self._h))

    @property
    def stored_snps(self):
        return int(self._lib.dgrm_stored_snps(self._h))

    def push_stored(self, first=0, nrows=None, row_keep=None):
        if nrows is None:
            nrows = self.stored_snps - first
        rk = None
        if row_keep is not None:
            rk = np.ascontiguousarray(row_keep, dtype=np.uint8)
            assert rk.shape == (nrows,)
        self._ck(self._lib.dgrm_push_stored(self._h, int(first), int(nrows), rk.ctypes.data if rk is not None else None))

    def fetch_standardised(self, bs=(256, 256), grid=(1, 1), coords=(0, 0), out=None):
        """Local block of the m x n standardised genotype matrix (row = SNP), Fortran order like Matrix::m."""
        nr = numroc(self.num_snps, bs[0], coords[0], grid[0])
        nc = numroc(self.num_individuals, bs[1], coords[1], grid[1])
        if out is None:
            out = np.full((nr, nc), np.nan, dtype=np.float64, order="F")
        assert out.dtype == np.float64 and out.flags.f_contiguous and out.shape == (nr, nc)
        self._ck(self._lib.dgrm_fetch_standardised(self._h, out.ctypes.data, max(nr, 1), bs[0], bs[1], coords[0],
                                                   coords[1], grid[0], grid[1]))
        return out

    def save(self, mode="copy"):
        self._ck(self._lib.dgrm_save(self._h, _lib.SAVE[mode]))

    def set_view(self, view="current"):
        self._ck(self._lib.dgrm_set_view(self._h, _lib.VIEW[view]))

    def finish(self):
        try:
            self._ck(self._lib.dgrm_finish(self._h))
        finally:
            self._keep.clear()

    def sync(self):
        self._ck(self._lib.dgrm_sync(self._h))

    @property
    def num_snps(self):
        return self._lib.dgrm_num_snps(self._h)

    @property
    def num_individuals(self):
        return self._lib.dgrm_num_individuals(self._h)

    def snp_stats(self):
        m = self.num_snps
        out = {
            "freq": np.zeros((m, 4), np.int32), "nn": np.zeros(m, np.int32),
            "p1": np.zeros(m), "p2": np.zeros(m), "sd": np.zeros(m),
        }
        self._ck(self._lib.dgrm_snp_stats(self._h, out["freq"].ctypes.data, out["nn"].ctypes.data,
                                          out["p1"].ctypes.data, out["p2"].ctypes.data, out["sd"].ctypes.data))
        return out

    def bad_snps(self, cap=65536):
        idx = np.zeros(cap, np.int64)
        cnt = C.c_int64(0)
        self._ck(self._lib.dgrm_bad_snps(self._h, idx.ctypes.data, cap, C.byref(cnt)))
        return idx[: min(cap, cnt.value)].copy(), cnt.value

    def fetch_block_cyclic(self, which="A", bs=(256, 256), grid=(1, 1), coords=(0, 0), uplo="L", out=None,
                           fill=np.nan):
        """Returns the caller-side local block (Fortran order, like Matrix::m)."""
        n = getattr(self, "_n_out", None) or self.num_individuals
        nr = numroc(n, bs[0], coords[0], grid[0])
        nc = numroc(n, bs[1], coords[1], grid[1])
        if out is None:
            out = np.full((nr, nc), fill, dtype=np.float64, order="F")
        assert out.dtype == np.float64 and out.flags.f_contiguous and out.shape == (nr, nc)
        self._ck(self._lib.dgrm_fetch_block_cyclic(self._h, _lib.FETCH[which], out.ctypes.data, max(nr, 1),
                                                   bs[0], bs[1], coords[0], coords[1], grid[0], grid[1],
                                                   uplo.encode()))
        return out

    def fetch_into(self, which, ptr, lld, bs=(256, 256), grid=(1, 1), coords=(0, 0), uplo="L", cols=None):
        """Fill caller memory at `ptr` (pageable / pinned host, or device memory of this GPU); `cols` = (lc0, lc1)
        restricts to those local columns, `ptr` then addresses column lc0."""
        if cols is None:
            self._ck(self._lib.dgrm_fetch_block_cyclic(self._h, _lib.FETCH[which], int(ptr), lld, bs[0], bs[1],
                                                       coords[0], coords[1], grid[0], grid[1], uplo.encode()))
        else:
            self._ck(self._lib.dgrm_fetch_block_cyclic_cols(self._h, _lib.FETCH[which], int(ptr), lld, bs[0], bs[1],
                                                            coords[0], coords[1], grid[0], grid[1], uplo.encode(),
                                                            int(cols[0]), int(cols[1])))

    def bind_output(self, which, ptr, lld, bs=(256, 256), grid=(1, 1), coords=(0, 0), uplo="L"):
        self._ck(self._lib.dgrm_bind_output(self._h, _lib.FETCH[which], int(ptr), lld, bs[0], bs[1], coords[0],
                                            coords[1], grid[0], grid[1], uplo.encode()))

    def bind_tiles(self, which, ptr):
        """dgrm_finish streams this engine's tiles of `which` ([num_tiles][256][128] doubles) into `ptr`."""
        self._ck(self._lib.dgrm_bind_tiles(self._h, _lib.FETCH[which], int(ptr)))

    def unbind_outputs(self):
        self._ck(self._lib.dgrm_unbind_outputs(self._h))

    def low_overlap_pairs(self, min_overlap=10.0, cap=1 << 20):
        pairs = np.zeros((cap, 2), np.int64)
        cnt = C.c_int64(0)
        self._ck(self._lib.dgrm_low_overlap_pairs(self._h, float(min_overlap), pairs.ctypes.data, cap, C.byref(cnt)))
        return pairs[: min(cap, cnt.value)].copy(), cnt.value

    def set_output_subset(self, idx=None):
        if idx is None or len(idx) == 0:
            self._ck(self._lib.dgrm_set_output_subset(self._h, None, 0))
            self._n_out = None
        else:
            a = np.ascontiguousarray(idx, dtype=np.int64)
            self._ck(self._lib.dgrm_set_output_subset(self._h, a.ctypes.data, len(a)))
            self._n_out = len(a)

    def write_grm_dat(self, path):
        self._ck(self._lib.dgrm_write_grm_dat(self._h, str(path).encode()))

    def write_grm_dat_part(self, path, create):
        """This engine's entries of <path> (multi-engine writer); `create` on exactly one engine, first."""
        self._ck(self._lib.dgrm_write_grm_dat_part(self._h, str(path).encode(), int(bool(create))))

    @property
    def num_tiles(self):
        return self._lib.dgrm_num_tiles(self._h)

    def tile_coords(self):
        rc = np.zeros((self.num_tiles, 2), np.int32)
        if self.num_tiles:
            self._ck(self._lib.dgrm_tile_coords(self._h, rc.ctypes.data))
        return rc

    def fetch_tiles(self, which="A", out_ptr=None):
        """This engine's tiles as doubles, [tile][col][row]; out_ptr: caller buffer (e.g. pinned) to fill instead."""
        if out_ptr is not None:
            if self.num_tiles:
                self._ck(self._lib.dgrm_fetch_tiles(self._h, _lib.FETCH[which], int(out_ptr)))
            return None
        t = np.zeros((self.num_tiles, 256, 128), np.float64)
        if self.num_tiles:
            self._ck(self._lib.dgrm_fetch_tiles(self._h, _lib.FETCH[which], t.ctypes.data))
        return t

    def fetch_tiles_range(self, which, tile0, count):
        """Tiles [tile0, tile0 + count) of this engine as doubles, [tile][col][row]."""
        t = np.zeros((count, 256, 128), np.float64)
        if count:
            self._ck(self._lib.dgrm_fetch_tiles_range(self._h, _lib.FETCH[which], t.ctypes.data, int(tile0), int(count)))
        return t

    def kernel_times(self):
        ms = (C.c_double * _lib.K_NUM)()
        ln = (C.c_int64 * _lib.K_NUM)()
        self._ck(self._lib.dgrm_kernel_times(self._h, C.byref(ms), C.byref(ln)))
        return {k: (ms[i], ln[i]) for i, k in enumerate(_lib.K_NAMES)}


def assemble_from_tiles(n, tile_sets):
    """Merge [(tile_coords, tiles)] from several engines into a dense lower-triangular n x n array
    (NaN elsewhere).  Host-side helper for multi-engine gathers."""
    out = np.full((n, n), np.nan)
    for rc, tiles in tile_sets:
        for (rb, cp), t in zip(rc, tiles):
            r0, c0 = int(rb) * 128, int(cp) * 256
            r1, c1 = min(r0 + 128, n), min(c0 + 256, n)
            if r0 >= n or c0 >= n:
                continue
            out[r0:r1, c0:c1] = t[: c1 - c0, : r1 - r0].T
    iu = np.triu_indices(n, 1)
    out[iu] = np.nan
    return out
