"""Random pedigrees through the pedigree compilers, on the CPU: several spouses per member, childless married-in
founders, unsequenced members, ped rows in random order (children before their parents), marriage loops for the Gibbs
sampler.  Every compiled artefact is executed on the host (the message program by the Python interpreter of
test_es_program_cpu.py, the generated CUDA C++ behind the one-thread shims of test_es_jit_cpu.py /
test_gibbs_jit_cpu.py) and must reproduce the oracle."""
import ctypes

import numpy as np
import pytest

import famseq_b200 as fs
from famseq_b200 import synth
from oracle import oracle as O
from tests.test_es_jit_cpu import build_host_peel
from tests.test_es_program_cpu import interpret
from tests.test_gibbs_jit_cpu import build_host_kernel, chain_stays_in_fast_range


def host_engine(ped, cols):
    return fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, cols, device=-1)


def likelihoods(S, V, seed):
    return synth.synth_likelihoods(synth._mk([(i, 0, 0, 1) for i in range(1, S + 1)]), V, seed=seed, x_fraction=0.4)


@pytest.mark.parametrize("seed", range(24))
def test_message_program_of_a_random_loop_free_pedigree(seed):
    ped = synth.random_pedigree(seed, 3 + seed % 17, shuffle=seed % 2 == 1, unsequenced=0.25 if seed % 3 == 0 else 0.0)
    cols = ped.sequenced_cols()
    if seed % 4 == 2:
        cols = list(reversed(cols))  # input columns in another order than the ped rows
    V = 12
    lk, fl = likelihoods(len(cols), V, 1000 + seed)
    want = O.run(ped, cols, lk, fl, method=O.ES, lc=5.0)
    with host_engine(ped, cols) as e:
        assert e.info()["has_loop"] == 0
        words, n_slots = e.es_program()
        a, xf, xm, _, _ = e.tables()
        priors = e.params.priors()
    checked = 0
    for v in range(V):
        got = interpret(words, n_slots, [a, xf, xm], priors, lk[v], int(fl[v]))
        if want["status"][v]:
            continue
        assert got is not None and np.array_equal(np.array(got), want["post"][v]), f"seed {seed} variant {v}"
        checked += 1
    assert checked > 0


@pytest.mark.parametrize("seed", range(6))
def test_generated_peeling_code_of_a_random_pedigree(seed, tmp_path):
    ped = synth.random_pedigree(50 + seed, 6 + 2 * seed, shuffle=seed % 2 == 0, unsequenced=0.2)
    cols = ped.sequenced_cols()
    S, V = len(cols), 16
    lk, fl = likelihoods(S, V, 2000 + seed)
    want = O.run(ped, cols, lk, fl, method=O.ES, lc=5.0)
    with host_engine(ped, cols) as e:
        src, _ = e.es_kernel()
        priors = e.params.priors()
    lib = build_host_peel(tmp_path, src, S)
    for v in range(V):
        known, chrx = int(fl[v]) & 1, (int(fl[v]) >> 1) & 1
        pa = np.ascontiguousarray(priors[1 if known else 0])
        pm = np.ascontiguousarray(priors[3 if known else 2]) if chrx else pa
        row, gt_row = np.zeros(S * 3), np.zeros(S, np.uint8)
        row_lk = np.ascontiguousarray(lk[v].reshape(-1))
        failed = lib.peel_host(chrx, row_lk.ctypes.data, pa.ctypes.data, pm.ctypes.data, row.ctypes.data, gt_row.ctypes.data)
        if want["status"][v]:
            continue
        assert not failed and np.array_equal(row.reshape(S, 3), want["post"][v]), f"seed {seed} variant {v}"


@pytest.mark.parametrize("seed", range(6))
def test_generated_gibbs_sampler_of_a_random_pedigree(seed, tmp_path):
    ped = synth.random_pedigree(80 + seed, 8 + 3 * seed, loops=True, shuffle=seed % 2 == 1, unsequenced=0.2 if seed % 3 == 0 else 0.0)
    cols = ped.sequenced_cols()
    S, V, burn, rep, rng_seed, v_offset = len(cols), 10, 10, 80, 77, 500
    lk, fl = likelihoods(S, V, 3000 + seed)
    want = O.run(ped, cols, lk, fl, method=O.MCMC, burn=burn, rep=rep, rng=O.RNG_PHILOX, seed=rng_seed, v_offset=v_offset)
    with host_engine(ped, cols) as e:
        src, _ = e.gibbs_kernel()
    lib = build_host_kernel(tmp_path, src)
    scratch = np.zeros(6 * ped.n * 8 + 64)
    checked = 0
    for v in range(V):
        row_lk, flag = np.ascontiguousarray(lk[v]), np.array([fl[v]], np.uint8)
        post, single = np.zeros((S, 3)), np.zeros((S, 3))
        gt, status = np.zeros(S, np.uint8), np.full(1, 9, np.uint8)
        lib.famseq_gibbs(row_lk.ctypes.data, flag.ctypes.data, post.ctypes.data, single.ctypes.data, gt.ctypes.data, status.ctypes.data,
                         1, burn, rep, rng_seed, v_offset + v, scratch.ctypes.data, 1, None)
        if status[0] == 2:  # legitimate only if a weight sum of this chain left the generated code's fast range
            assert not chain_stays_in_fast_range(ped, cols, lk[v:v + 1], fl[v:v + 1], burn, rep, rng_seed, v_offset + v), f"seed {seed} variant {v}"
            continue
        assert status[0] == want["status"][v], f"seed {seed} variant {v}"
        if status[0]:
            continue
        assert np.allclose(post, want["post"][v], rtol=1e-9, atol=0), f"seed {seed} variant {v}"
        assert np.array_equal(gt, want["gt"][v].astype(np.uint8))
        checked += 1
    assert checked > 0
