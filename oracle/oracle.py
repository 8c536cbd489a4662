"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes loader for the C restatement (oracle/libfamseq_oracle.so) and a runner for the
reference harness (oracle/_ref/ref_harness, built from the unmodified reference sources).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; famseq_b200/ must never do so.
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess
import tempfile
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfamseq_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_HARNESS = os.path.join(REF_DIR, "ref_harness")
REF_CLI = os.path.join(REF_DIR, "FamSeq")

BN, ES, MCMC = 1, 2, 3
RNG_LIBC, RNG_PHILOX = 0, 1

DEFAULT_PRIORS = np.array(
    [[0.9985, 0.001, 0.0005], [0.45, 0.1, 0.45], [0.999, 0.0, 0.001], [0.5, 0.0, 0.5]], dtype=np.float64
)


def build(ref: bool = True) -> None:
    """Compile the C restatement and, when the reference sources are present, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(ref=False)
        L = ctypes.CDLL(LIB_PATH)
        P = ctypes.c_void_p
        L.fso_run.restype = ctypes.c_int
        L.fso_run.argtypes = [ctypes.c_int, ctypes.c_int, P, P, P, P, ctypes.c_int, P, ctypes.c_double,
                              ctypes.c_double, P, ctypes.c_int64, P, P, ctypes.c_int, ctypes.c_int,
                              ctypes.c_int, ctypes.c_int64, ctypes.c_int64, P, P, P, P, P, P]
        L.fso_tables.restype = None
        L.fso_tables.argtypes = [ctypes.c_double, P, P, P]
        L.fso_topology.restype = ctypes.c_int
        L.fso_topology.argtypes = [ctypes.c_int, P, P, P, P, P, P]
        L.fso_pl_decode.restype = ctypes.c_double
        L.fso_pl_decode.argtypes = [ctypes.c_double]
        L.fso_pl_table.restype = None
        L.fso_pl_table.argtypes = [P, ctypes.c_int]
        L.fso_phred_text.restype = ctypes.c_int
        L.fso_phred_text.argtypes = [ctypes.c_double, ctypes.c_char_p]
        L.fso_mcmc_sum_range.restype = None
        L.fso_mcmc_sum_range.argtypes = [P]
        L.fso_philox4x32.restype = None
        L.fso_philox4x32.argtypes = [P, P, P]
        _lib = L
    return _lib


@dataclass
class Pedigree:
    """Rows of a FamSeq ped file: id, mother id, father id, gender (1 male / 2 female), sample name."""
    ids: list
    mids: list
    fids: list
    genders: list
    names: list = field(default_factory=list)

    @property
    def n(self) -> int:
        return len(self.ids)

    @staticmethod
    def read(path: str) -> "Pedigree":
        """Same reading rules as the reference (file.cpp:24-62): skip one header line, stop at the
        first line shorter than two characters."""
        ids, mids, fids, genders, names = [], [], [], [], []
        with open(path) as fh:
            fh.readline()
            for line in fh:
                line = line.rstrip("\n")
                if len(line) < 2:
                    break
                t = line.split()
                ids.append(int(t[0])); mids.append(int(t[1])); fids.append(int(t[2]))
                genders.append(int(t[3])); names.append(t[4] if len(t) > 4 else "")
        return Pedigree(ids, mids, fids, genders, names)

    def write(self, path: str) -> None:
        with open(path, "w") as fh:
            fh.write("ID\tmID\tfID\tgender\tIndividualName\n")
            for r in zip(self.ids, self.mids, self.fids, self.genders, self.names or ["NA"] * self.n):
                fh.write("\t".join(str(x) for x in r) + "\n")

    def sequenced_cols(self) -> list:
        """Ped rows of members with a real sample name (everything but 'NA'), in ped order."""
        return [i for i, nm in enumerate(self.names) if nm != "NA"]


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def tables(mrate: float):
    a, xf, xm = (np.zeros(27) for _ in range(3))
    lib().fso_tables(mrate, _ptr(a), _ptr(xf), _ptr(xm))
    return a, xf, xm


def pl_table(n: int = 65536) -> np.ndarray:
    """pow(10, -fabs(pl)/10) for pl = 0..n-1 (file.cpp:588-590), evaluated by the C restatement (libm)."""
    out = np.empty(n, np.float64)
    lib().fso_pl_table(_ptr(out), n)
    return out


def run(ped: Pedigree, cols, lk, flags=None, method=ES, mrate=1e-7, lc=1.0, priors=None, burn=1000,
        rep=100000, rng=RNG_LIBC, seed=-1, v_offset=0):
    """Run the C restatement.  lk: [V][S][3] float64.  Returns dict(post, single, gt, status, post_full, single_full)."""
    lk = np.ascontiguousarray(lk, dtype=np.float64)
    V, S = lk.shape[0], lk.shape[1]
    N = ped.n
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    assert cols.shape[0] == S
    flags = np.zeros(V, np.uint8) if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
    priors = DEFAULT_PRIORS if priors is None else np.ascontiguousarray(priors, dtype=np.float64)
    i32 = lambda x: np.ascontiguousarray(x, dtype=np.int32)
    ids, mids, fids, gen = i32(ped.ids), i32(ped.mids), i32(ped.fids), i32(ped.genders)
    post = np.zeros((V, S, 3)); single = np.zeros((V, S, 3))
    gt = np.zeros((V, S), np.int32); status = np.zeros(V, np.uint8)
    pf = np.zeros((V, N, 3)); sf = np.zeros((V, N, 3))
    rc = lib().fso_run(method, N, _ptr(ids), _ptr(mids), _ptr(fids), _ptr(gen), S, _ptr(cols), mrate, lc,
                       _ptr(priors), V, _ptr(flags), _ptr(lk), burn, rep, rng, seed, v_offset, _ptr(post),
                       _ptr(single), _ptr(gt), _ptr(status), _ptr(pf), _ptr(sf))
    if rc != 0:
        raise RuntimeError(f"oracle error {rc}")
    return dict(post=post, single=single, gt=gt, status=status, post_full=pf, single_full=sf)


def phred_text(p: float) -> str:
    """The reference drivers' text of one posterior (file.cpp:702-749)."""
    buf = ctypes.create_string_buffer(40)
    n = lib().fso_phred_text(float(p), buf)
    return buf.raw[:n].decode()


def mcmc_sum_range():
    """(min, max) Gibbs weight sum met by the chains of the last run(method=MCMC) call (test probe)."""
    out = np.zeros(2)
    lib().fso_mcmc_sum_range(_ptr(out))
    return float(out[0]), float(out[1])


FAST_SUM_LO, FAST_SUM_HI = 2.0 ** -963, 2.0 ** 963  # range of weight sums the generated Gibbs kernel handles itself


def have_ref() -> bool:
    return os.path.exists(REF_HARNESS)


def run_ref(ped_path: str, cols, lk, flags=None, method=ES, mrate=1e-7, lc=1.0, priors=None, burn=1000,
            rep=100000, seed=-1, repeat=1, limit=-1, want_output=True):
    """Run the unmodified reference engine through oracle/_ref/ref_harness."""
    lk = np.ascontiguousarray(lk, dtype=np.float64)
    V, S = lk.shape[0], lk.shape[1]
    flags = np.zeros(V, np.uint8) if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as fh:
            fh.write(np.array([V, S], np.int32).tobytes())
            fh.write(flags.tobytes())
            fh.write(lk.tobytes())
        args = [REF_HARNESS, f"ped={ped_path}", f"method={method}", f"mrate={mrate!r}", f"lc={lc!r}",
                f"burn={burn}", f"rep={rep}", f"seed={seed}", "cols=" + ",".join(str(int(c)) for c in cols),
                f"repeat={repeat}", f"limit={limit}", f"in={fin}"]
        if want_output:
            args.append(f"out={fout}")
        if priors is not None:
            pr = np.asarray(priors, dtype=np.float64)
            for k, row in zip(("gpn", "gpk", "gpxn", "gpxk"), pr):
                args.append(k + "=" + ",".join(repr(float(x)) for x in row))
        r = subprocess.run(args, check=True, capture_output=True, text=True)
        info = json.loads(r.stdout.strip().splitlines()[-1])
        if not want_output:
            return info
        raw = open(fout, "rb").read()
    V2, S2, N = np.frombuffer(raw, np.int32, 3)
    o = 12
    status = np.frombuffer(raw, np.uint8, V2, o); o += V2
    post = np.frombuffer(raw, np.float64, V2 * S2 * 3, o).reshape(V2, S2, 3); o += V2 * S2 * 24
    single = np.frombuffer(raw, np.float64, V2 * S2 * 3, o).reshape(V2, S2, 3); o += V2 * S2 * 24
    gt = np.frombuffer(raw, np.int32, V2 * S2, o).reshape(V2, S2); o += V2 * S2 * 4
    pf = np.frombuffer(raw, np.float64, V2 * N * 3, o).reshape(V2, N, 3); o += V2 * N * 24
    sf = np.frombuffer(raw, np.float64, V2 * N * 3, o).reshape(V2, N, 3)
    return dict(post=post.copy(), single=single.copy(), gt=gt.copy(), status=status.copy(),
                post_full=pf.copy(), single_full=sf.copy(), info=info)
