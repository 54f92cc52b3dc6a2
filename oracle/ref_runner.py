"""Runs the reference's own binaries built by oracle/Makefile (oracle/_ref/dissect, snpdump).
TEST INFRASTRUCTURE ONLY.  Never reads /root/reference at run time — only the prebuilt binaries."""
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DISSECT = os.path.join(_HERE, "_ref", "dissect")
SNPDUMP = os.path.join(_HERE, "_ref", "snpdump")


def available():
    return os.path.exists(DISSECT) and os.path.exists(SNPDUMP)


def _env(threads):
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
        env["OPENBLAS_NUM_THREADS"] = str(threads)
    return env


def run_make_grm(bfile, out, extract=None, keep=None, flat_norm=False, threads=None, timeout=3600, extra=()):
    """`dissect --make-grm --bfile <bfile> --out <out> [--extract F] [--keep F] [--grm-flat-norm]`.
    Returns (returncode, log text)."""
    cmd = [DISSECT, "--make-grm", "--bfile", bfile, "--out", out]
    if extract:
        cmd += ["--extract", extract]
    if keep:
        cmd += ["--keep", keep]
    if flat_norm:
        cmd += ["--grm-flat-norm"]
    cmd += list(extra)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=_env(threads))
    return r.returncode, r.stdout + r.stderr


def run_snpdump(bfile, out, extract=None, keep=None, dump_z=None, timeout=3600):
    cmd = [SNPDUMP, "--make-grm", "--bfile", bfile, "--out", out]
    if extract:
        cmd += ["--extract", extract]
    if keep:
        cmd += ["--keep", keep]
    env = _env(None)
    if dump_z:
        env["DGRM_DUMP_Z"] = dump_z
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    if r.returncode != 0:
        raise RuntimeError("snpdump failed: " + r.stdout[-2000:] + r.stderr[-2000:])
    return parse_snpdump(out + ".snpdump")


def parse_snpdump(path):
    rows = [l.split() for l in open(path) if not l.startswith("#")]
    return {
        "names": [r[0] for r in rows],
        "freq": np.array([[int(x) for x in r[1:5]] for r in rows], dtype=np.int32).reshape(-1, 4),
        "nn": np.array([int(r[5]) for r in rows], dtype=np.int32),
        "p1": np.array([float.fromhex(r[6]) for r in rows]),
        "p2": np.array([float.fromhex(r[7]) for r in rows]),
        "sd": np.array([float.fromhex(r[8]) for r in rows]),
    }
