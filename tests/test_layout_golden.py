"""CPU: the block-cyclic index arithmetic (the layout of `Matrix::m`) pinned against the reference's own
code.  tests/golden/layout/layout.txt was printed by oracle/_ref/layoutdump, which calls the reference's unmodified
Matrix::local2global and Matrix::global2local (matrix.cpp:435-488) for emulated processes (myRow, myCol) of R x C grids.
Checked against it: the oracle's restatement (oracle/grm_oracle.py), the binding's numroc (engine.py) and the C++ host
mirror use the same mapping; the GPU fetch kernels are then checked against the oracle on the GPU."""
import os

import numpy as np

from dissect_b200.engine import numroc
from oracle import grm_oracle as orc
from tests.helpers import GOLDEN


def _cases():
    cur = None
    for line in open(os.path.join(GOLDEN, "layout", "layout.txt")):
        t = line.split()
        if t[0] == "CASE":
            if cur:
                yield cur
            cur = {"args": tuple(int(x) for x in t[1:]), "R": {}, "C": {}, "G": {}}
        elif t[0] in ("R", "C"):
            cur[t[0]][int(t[1])] = int(t[2])
        else:
            cur["G"][int(t[1])] = tuple(int(x) for x in t[2:])
    yield cur


def test_local_to_global_and_numroc_match_the_reference():
    ncases = 0
    for c in _cases():
        n, bsr, bsc, R, C, r, col, nlr, nlc = c["args"]
        ncases += 1
        # ownership according to the reference's global2local: which globals live on (r, col), and where
        rows = [g for g in range(n) if c["G"][g][0] == r]
        cols = [g for g in range(n) if c["G"][g][2] == col]
        assert len(rows) == nlr == numroc(n, bsr, r, R)          # numroc_ (ScaLAPACK) == the reference's owner map
        assert len(cols) == nlc == numroc(n, bsc, col, C)
        assert [c["G"][g][1] for g in rows] == list(range(nlr))  # local positions are dense and ascending
        assert [c["G"][g][3] for g in cols] == list(range(nlc))
        # local -> global: reference vs restatement, and inverse of global2local
        for l in range(nlr):
            assert c["R"][l] == orc.local_to_global(l, bsr, r, R) == rows[l]
        for l in range(nlc):
            assert c["C"][l] == orc.local_to_global(l, bsc, col, C) == cols[l]
    assert ncases == 6 + 4 + 6 + 1 + 2 + 3


def test_block_cyclic_local_piece_matches_the_reference_indices():
    """oracle.block_cyclic_local (what the GPU fetch is compared with) selects exactly the reference's rows/columns."""
    for c in _cases():
        n, bsr, bsc, R, C, r, col, nlr, nlc = c["args"]
        if n > 300:
            continue
        G = np.arange(n * n, dtype=np.float64).reshape(n, n)
        loc = orc.block_cyclic_local(G, (bsr, bsc), (R, C), (r, col))
        assert loc.shape == (nlr, nlc)
        want = G[np.ix_([c["R"][l] for l in range(nlr)], [c["C"][l] for l in range(nlc)])]
        assert np.array_equal(loc, want)


def test_host_mirror_matrix_geometry_matches_the_reference():
    """The C++ host mirror's Matrix (dissect_b200/host: initParameters = numroc, local2globalRow/Col) against the same
    reference output."""
    import subprocess
    exe = os.path.join(os.path.dirname(GOLDEN), "..", "dissect_b200", "host", "selftest")
    exe = os.path.normpath(exe)
    if not os.path.exists(exe):
        import pytest
        pytest.skip("host selftest not built (build())")
    for c in _cases():
        n, bsr, bsc, R, C, r, col, nlr, nlc = c["args"]
        res = subprocess.run([exe, "layout"] + [str(x) for x in (n, bsr, bsc, R, C, r, col)], capture_output=True,
                             text=True, timeout=60)
        assert res.returncode == 0, res.stdout[-500:] + res.stderr[-500:]
        got = {"R": {}, "C": {}}
        for line in res.stdout.splitlines():
            t = line.split()
            if t[0] == "N":
                assert (int(t[1]), int(t[2])) == (nlr, nlc)
            elif t[0] in got:
                got[t[0]][int(t[1])] = int(t[2])
        assert got["R"] == c["R"] and got["C"] == c["C"]
