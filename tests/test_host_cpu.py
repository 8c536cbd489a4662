"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol the header
declares, and the pedigree compilers (fs_create with device = -1: no CUDA call) agree with the oracle's view of
the same pedigree.  No compute call is made here (there is no CPU compute path)."""
import ctypes
import os
import re

import numpy as np
import pytest

import famseq_b200 as fs
from famseq_b200 import synth
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "famseq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fs_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = fs.lib()
    syms = header_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/famseq_b200.h but not exported"
    assert sorted(fs.engine.EXPORTED_SYMBOLS) == syms


def host_engine(ped, params=None):
    return fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, ped.sequenced_cols(), params=params, device=-1)


@pytest.mark.parametrize("mrate", [1e-7, 0.0, 1e-4, 0.01, 0.5])
def test_transmission_tables_bit_identical_to_oracle(mrate):
    prm = fs.Params.default()
    prm.mrate = mrate
    with host_engine(synth.trio(), prm) as e:
        a, xf, xm, _, _ = e.tables()
    oa, oxf, oxm = O.tables(mrate)
    assert np.array_equal(a, oa) and np.array_equal(xf, oxf) and np.array_equal(xm, oxm)


@pytest.mark.parametrize("name", sorted(synth.PEDIGREES))
def test_topology_matches_oracle(name):
    ped = synth.PEDIGREES[name]()
    with host_engine(ped) as e:
        _, _, _, mo, fa = e.tables()
        info = e.info()
    n = ped.n
    i32 = lambda x: np.ascontiguousarray(x, dtype=np.int32)
    omo, ofa = np.zeros(n, np.int32), np.zeros(n, np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    ids, mids, fids, gen = i32(ped.ids), i32(ped.mids), i32(ped.fids), i32(ped.genders)
    assert O.lib().fso_topology(n, p(ids), p(mids), p(fids), p(gen), p(omo), p(ofa)) == 0
    assert np.array_equal(mo, omo) and np.array_equal(fa, ofa)
    assert info["n"] == n and info["s"] == n
    assert info["n_founders"] == int((omo < 0).sum())
    assert info["mcmc_links"] == 2 * int((omo >= 0).sum())
    assert info["bn_levels"] == (n if n <= 31 else 0)


def test_loop_detection_and_es_refusal():
    for name, loop in (("trio", 0), ("ped14", 0), ("half_sibs", 0), ("three_wives", 0), ("cousins_loop", 1), ("ped40", 1)):
        with host_engine(synth.PEDIGREES[name]()) as e:
            info = e.info()
            assert info["has_loop"] == loop, name
            assert (info["es_ops"] == 0) == bool(loop), name


def test_es_program_sizes():
    with host_engine(synth.trio()) as e:
        i = e.info()
        # trio: pos(father->mother), pos(mother->father), ant(child) and one FIN per member; products ant*lk of the
        # two founders are shared
        assert i["es_ops"] == 8 and i["es_slots"] <= 5
    with host_engine(synth.ped14()) as e:
        i = e.info()
        assert 14 <= i["es_ops"] < 80 and i["es_slots"] < 40


def test_pedigree_errors_follow_the_reference():
    # one parent only: "This is not a fulfill family" (family.cpp:318-322)
    with pytest.raises(fs.FamSeqError) as ei:
        fs.Engine([1, 2, 3], [0, 0, 2], [0, 0, 0], [1, 2, 1], [0, 1, 2], device=-1)
    assert ei.value.code == -2 and "fulfill" in str(ei.value)
    # mother is male (family.cpp:204-219)
    with pytest.raises(fs.FamSeqError) as ei:
        fs.Engine([1, 2, 3], [0, 0, 1], [0, 0, 2], [1, 2, 1], [0, 1, 2], device=-1)
    assert ei.value.code == -3
    # BN size limit
    n = 40
    with host_engine(synth.ped40()) as e:
        assert e.info()["bn_levels"] == 0  # 3^40 configurations: refused, not attempted


def test_no_cpu_fallback():
    with host_engine(synth.trio()) as e:
        lk = np.ones((4, 3, 3))
        with pytest.raises(fs.FamSeqError) as ei:
            e.run(fs.ES, lk)
        assert ei.value.code == -6 and "no CPU fallback" in str(ei.value)


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "famseq_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), f"{os.path.join(dirpath, f)} mentions the oracle"


def test_synthetic_generator_is_sliceable():
    ped = synth.ped14()
    full, ff = synth.synth_likelihoods(ped, 70000, seed=3, x_fraction=0.1)
    part, pf = synth.synth_likelihoods(ped, 1000, seed=3, v0=65000, x_fraction=0.1)
    assert np.array_equal(full[65000:66000], part) and np.array_equal(ff[65000:66000], pf)
    assert full.min() > 0 and full.max() == 1.0


@pytest.mark.parametrize("name", ["trio", "ped14", "ped40"])
def test_generated_gibbs_kernel_compiles_for_sm_100a(name):
    """The pedigree-specialised Gibbs kernel (csrc/cuda/gibbs_jit.cu): the engine writes CUDA C++ for the pedigree and
    NVRTC compiles it to an sm_100a cubin; neither step needs a device.  Every member must appear as a block of its
    own, and the hot loop must not spill more than a few registers."""
    ped = synth.PEDIGREES[name]()
    with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, ped.sequenced_cols(), device=-1) as e:
        src, _ = e.gibbs_kernel()
        # one evaluation function per member and rule set (autosomal, chrX), called from the member-by-member path of its group
        for i in range(ped.n):
            assert src.count(f"Rc eval_a_{i}(") == 1 and src.count(f"Rc eval_x_{i}(") == 1 and src.count(f"= eval_a_{i}(") == 1
        assert 'extern "C" __global__' in src and "famseq_gibbs" in src
        log, cubin_bytes = e.gibbs_kernel(compile=True)
    assert cubin_bytes > 0
    m = re.search(r"(\d+) bytes spill stores", log)
    assert m and int(m.group(1)) <= 256, log


def test_fast_number_formatter_matches_printf_g():
    """The output writer's "%g" replacement (csrc/cli/format_g.hpp) against glibc's snprintf on a few million values:
    raw bit patterns, Phred values, decimal ties and their neighbours, neighbours of powers of ten."""
    import subprocess
    exe = os.path.join(ROOT, "famseq_b200", "bin", "format_check")
    r = subprocess.run([exe, "300000", "20261018"], capture_output=True, text=True)
    assert r.returncode == 0 and "0 mismatches" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("name", ["half_sibs", "ped14"])
def test_generated_es_kernel_compiles_for_sm_100a(name):
    """The Elston-Stewart message program as straight-line code (csrc/cuda/es_jit.cu): generated and compiled with NVRTC
    without a device; big pedigrees are refused (they stay with the interpreter)."""
    ped = synth.PEDIGREES[name]()
    with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, ped.sequenced_cols(), device=-1) as e:
        src, _ = e.es_kernel()
        assert "peel_a(" in src and "peel_x(" in src and 'extern "C" __global__' in src
        log, cubin_bytes = e.es_kernel(compile=True)
    assert cubin_bytes > 0
    m = re.search(r"(\d+) bytes spill stores", log)
    assert m and int(m.group(1)) <= 512, log
    big = synth.ped100()
    with fs.Engine(big.ids, big.mids, big.fids, big.genders, big.sequenced_cols(), device=-1) as e:
        with pytest.raises(fs.FamSeqError):
            e.es_kernel()


def test_generated_kernels_can_be_cached_on_disk(tmp_path, monkeypatch):
    """FAMSEQ_JIT_CACHE_DIR: the second engine on the same pedigree loads the cubin instead of compiling."""
    import time
    monkeypatch.setenv("FAMSEQ_JIT_CACHE_DIR", str(tmp_path))
    ped = synth.half_sibs()
    sizes, times, logs = [], [], []
    for _ in range(2):
        with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, ped.sequenced_cols(), device=-1) as e:
            t0 = time.time()
            log, n = e.gibbs_kernel(compile=True)
            times.append(time.time() - t0)
            sizes.append(n)
            logs.append(log)
    assert sizes[0] == sizes[1] > 0 and "cubin loaded from" in logs[1] and "cubin loaded from" not in logs[0]
    assert len([f for f in os.listdir(tmp_path) if f.endswith(".cubin")]) == 1
    assert times[1] < times[0]


def test_pl_decode_table_is_libm_pow_for_all_65536_values():
    """fs_run_pl decodes uint16 PLs through a table built on the host with libm: lut[pl] == pow(10, -pl/10), the
    expression of the reference's VCF driver (file.cpp:588-590), for every one of the 65 536 values -- checked against
    the oracle's restatement (C, libm) and against Python's math.pow (libm as well; numpy's own power() differs in the
    last bit for ~0.2 % of the values, which is why the table is not built with it)."""
    import math

    with host_engine(synth.trio()) as e:
        lut = e.pl_table()
    assert lut.shape == (65536,) and np.array_equal(lut, O.pl_table())
    assert np.array_equal(lut, np.array([math.pow(10.0, -abs(float(k)) / 10.0) for k in range(65536)]))
    assert lut[0] == 1.0 and lut[10] == 0.1 and lut[3236] > 0 and lut[3237] == 0.0 and not lut[3237:].any()
    assert np.array_equal(synth.pl_to_likelihood(np.arange(70000)), np.concatenate([lut, np.zeros(70000 - 65536)]))


def test_compact_entry_has_no_cpu_fallback_either():
    with host_engine(synth.trio()) as e:
        with pytest.raises(fs.FamSeqError) as ei:
            e.run_pl(fs.ES, np.zeros((4, 3, 3), np.uint16))
        assert ei.value.code == -6 and "no CPU fallback" in str(ei.value)
        assert e.info()["n_devices"] == 0 and e.info()["mcmc_fixups"] == 0


def test_multi_device_argument_checks():
    ped = synth.trio()
    for devices in ([], [-1], [0, -2]):
        with pytest.raises(fs.FamSeqError) as ei:
            fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, ped.sequenced_cols(), device=devices)
        assert ei.value.code == -1


def test_phred_code_text_is_printf_g():
    """fs_phred_text: the text of a packed Phred code (six digits m, decimal exponent e: value m * 10^(e-5)) must be what
    "%g" prints for that number -- every exponent the codes can carry, trailing-zero and notation rules included."""
    rng = np.random.default_rng(7)
    for e in range(-20, 8):
        for m in list(rng.integers(100000, 1000000, 200)) + [100000, 999999, 100001, 500000, 120000, 123000, 123400, 123450]:
            code = int(m) | ((e + 32) << 20)
            want = "%g" % float(f"{int(m)}e{e - 5}")
            assert fs.phred_text(code) == want, (m, e, fs.phred_text(code), want)
    assert fs.phred_text(1 << 30) == "0" and fs.phred_text(2 << 30) == "99999" and fs.phred_text(3 << 30) is None


def test_exact_phred_text_matches_the_oracle():
    """fs_phred_text_exact (what a caller formats fs_phred_fix entries with) against the oracle's restatement of
    file.cpp:702-749, including 0, 1, subnormals, NaN and values above one."""
    rng = np.random.default_rng(11)
    ps = np.concatenate([rng.random(3000), 10.0 ** (-320 * rng.random(3000)), 1 - 10.0 ** (-16 * rng.random(3000)),
                         [0.0, 1.0, 5e-324, 2.2e-308, 0.5, np.nan, -0.25, 1.5, np.inf, -0.0]])
    for p in ps:
        assert fs.phred_text_exact(p) == O.phred_text(p), p
