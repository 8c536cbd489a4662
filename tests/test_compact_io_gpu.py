"""GPU tests of the compact-I/O entry points (input half) and of the multi-device engine.

fs_run_pl ships integer Phred-scaled likelihoods (uint16) and decodes them on the device through a table the engine
builds on the host with libm's pow -- the expression of the reference's VCF driver (file.cpp:588-590) -- so every
output byte must equal what fs_run returns for the decoded doubles; `single = NULL` must change nothing else; an
engine over several GPUs must return the bytes of a single-GPU engine."""
import numpy as np
import pytest

import famseq_b200 as fs
from famseq_b200 import synth
from oracle import oracle as O
from tests.util import REL_TOL, CasePed, assert_parity, golden_cases, load_case

pytestmark = pytest.mark.gpu


def engine_for(ped, cols=None, params=None, device=0):
    cols = ped.sequenced_cols() if cols is None else cols
    return fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, cols, params=params, device=device)


def pl_batch(ped, V, seed, x_fraction=0.25):
    """PL integers as uint16 with the awkward values mixed in: beyond the synthetic clip (2550), in the subnormal range
    (3077..3236), decoding to exactly 0 (>= 3237), the uint16 maximum, and all-zero (missing) samples."""
    pl, fl = synth.synth_pl(ped, V, seed, x_fraction=x_fraction)
    pl = pl.astype(np.uint16)
    rng = np.random.default_rng(seed)
    odd = rng.random(pl.shape) < 0.01
    pl[odd] = rng.choice(np.array([2551, 3076, 3077, 3100, 3236, 3237, 5000, 65535], np.uint16), int(odd.sum()))
    pl[rng.random(pl.shape[:2]) < 0.01] = 0
    pl[rng.random(pl.shape[:2]) < 0.002] = 65535  # an impossible sample: the variant fails (status 1)
    return pl, fl


def same_bytes(a, b, single=True):
    assert np.array_equal(a.status, b.status) and np.array_equal(a.gt, b.gt)
    assert np.array_equal(a.post, b.post, equal_nan=True)
    if single:
        assert np.array_equal(a.single, b.single, equal_nan=True)


def test_decode_table_on_the_engine_is_the_oracles():
    ped = synth.trio()
    with engine_for(ped) as e:
        assert np.array_equal(e.pl_table(), O.pl_table())


@pytest.mark.parametrize("pedname,method,V,kw", [
    ("trio", fs.ES, 100_003, {}), ("trio", fs.BN, 20_001, {}), ("trio", fs.MCMC, 3000, dict(burn=10, rep=60, seed=5, v_offset=11)),
    ("ped14", fs.ES, 20_011, {}), ("half_sibs", fs.ES, 5003, {}), ("half_sibs", fs.BN, 40, {}),
    ("ped40", fs.MCMC, 300, dict(burn=10, rep=60, seed=5, v_offset=11)), ("ped100", fs.ES, 700, {}),
])
@pytest.mark.parametrize("jit", ["0", "1"])
def test_pl_entry_returns_the_bytes_of_the_fp64_entry(pedname, method, V, kw, jit, monkeypatch):
    if jit == "1" and (pedname == "trio" or method == fs.BN):
        pytest.skip("no generated kernel on this path")
    monkeypatch.setenv("FAMSEQ_MCMC_JIT", jit)
    monkeypatch.setenv("FAMSEQ_ES_JIT", jit)
    ped = synth.PEDIGREES[pedname]()
    pl, fl = pl_batch(ped, V, seed=900 + V)
    lk = synth.pl_table()[pl]
    with engine_for(ped) as e:
        want = e.run(method, lk, fl, **kw)
        got = e.run_pl(method, pl, fl, **kw)
        lean = e.run_pl(method, pl, fl, want_single=False, **kw)
        lean64 = e.run(method, lk, fl, want_single=False, **kw)
    assert want.status.sum() < V and (V < 5000 or want.status.sum() > 0)  # the impossible samples make some variants fail
    same_bytes(got, want)
    assert lean.single is None and lean64.single is None
    same_bytes(lean, want, single=False)
    same_bytes(lean64, want, single=False)


@pytest.mark.parametrize("n_children", [1, 2, 3, 4, 5])
def test_nuclear_kernel_with_compact_input_against_the_oracle(n_children):
    """The fused path (es_nuclear_kernel.cu gathers from the decode table itself): bit-identical to the oracle on the
    decoded doubles, chrX and Known flags, every sibship size, a ragged last tile."""
    rows = [(1, 0, 0, 1), (2, 0, 0, 2)] + [(3 + k, 2, 1, 1 + k % 2) for k in range(n_children)]
    ped = synth._mk(rows)
    V = 50_017
    pl, fl = pl_batch(ped, V, seed=40 + n_children, x_fraction=0.3)
    want = O.run(ped, ped.sequenced_cols(), O.pl_table()[pl], fl, method=O.ES)
    with engine_for(ped) as e:
        got = e.run_pl(fs.ES, pl, fl)
        lean = e.run_pl(fs.ES, pl, fl, want_single=False)
    assert_parity(got, want, 0.0, f"nuclear/pl C={n_children}")
    ok = want["status"] == 0
    assert np.array_equal(got.post[ok], want["post"][ok]) and np.array_equal(got.single[ok], want["single"][ok])
    same_bytes(lean, got, single=False)


def test_partially_sequenced_trio_with_compact_input():
    rows = [(7, 5, 9, 2), (9, 0, 0, 1), (4, 5, 9, 1), (5, 0, 0, 2)]  # children before parents, the father unsequenced
    ped = synth._mk(rows)
    cols = [3, 0, 2]
    pl, fl = pl_batch(synth.trio(), 20_000, seed=78, x_fraction=0.5)
    want = O.run(ped, cols, O.pl_table()[pl], fl, method=O.ES)
    with engine_for(ped, cols) as e:
        got = e.run_pl(fs.ES, pl, fl)
        lean = e.run(fs.ES, O.pl_table()[pl], fl, want_single=False)
    assert_parity(got, want, 0.0, "nuclear/pl/partial")
    same_bytes(lean, got, single=False)


def _as_pl(lk):
    """uint16 PLs whose decode is exactly `lk`, or None when some entry is not a table value."""
    table = O.pl_table()
    with np.errstate(divide="ignore"):
        guess = np.where(lk > 0, np.rint(-10.0 * np.log10(np.where(lk > 0, lk, 1.0))), 3237.0)
    guess = np.clip(guess, 0, 65535).astype(np.int64)
    return guess.astype(np.uint16) if np.array_equal(table[guess], lk) else None


def test_golden_vectors_through_the_compact_entry():
    """Every reference-generated golden case whose likelihoods are decoded integer PLs is re-run through fs_run_pl."""
    done = 0
    for name in golden_cases():
        c = load_case(name)
        if name.endswith("_mcmc"):
            continue
        pl = _as_pl(c["lk"])
        if pl is None:
            continue
        prm = fs.Params.default()
        prm.mrate, prm.lrc = float(c["mrate"]), float(c["lc"])
        for k, row in zip(("geno_prob_n", "geno_prob_k", "geno_prob_xn", "geno_prob_xk"), c["priors"]):
            for g in range(3):
                getattr(prm, k)[g] = float(row[g])
        method = int(c["method"])
        with engine_for(CasePed(c), c["cols"].tolist(), prm) as e:
            got = e.run_pl(method, pl, c["flags"])
        assert_parity(got, c, REL_TOL, name + " (pl)")
        if method == fs.ES:
            ok = c["status"] == 0
            assert np.array_equal(got.post[ok], c["post"][ok]), f"{name}: ES through fs_run_pl is not bit-identical to the reference"
        done += 1
    assert done >= 2, done  # the TestData VCF cases; the synthetic goldens were decoded with numpy's power(), 1 ulp off libm for a few PLs


def test_ragged_batches_through_the_compact_entry():
    ped = synth.trio()
    pl, fl = pl_batch(ped, 5000, seed=4)
    want = O.run(ped, ped.sequenced_cols(), O.pl_table()[pl], fl, method=O.ES)
    with engine_for(ped) as e:
        assert e.run_pl(fs.ES, pl[:0], fl[:0]).post.shape == (0, 3, 3)
        for V in (1, 2, 31, 32, 33, 1023, 1024, 1025, 4999):
            got = e.run_pl(fs.ES, pl[:V], fl[:V], want_single=False)
            ok = want["status"][:V] == 0
            assert np.array_equal(got.status, want["status"][:V])
            assert np.array_equal(got.post[ok], want["post"][:V][ok]) and np.array_equal(got.gt[ok], want["gt"][:V][ok].astype(np.uint8))
        assert np.array_equal(e.run_pl(fs.ES, pl[:100], None).post, e.run(fs.ES, O.pl_table()[pl[:100]], None).post)


def test_compact_input_on_device_buffers():
    torch = pytest.importorskip("torch")
    for ped, V in ((synth.trio(), 70_001), (synth.ped14(), 5001)):
        pl, fl = pl_batch(ped, V, seed=3)
        S = pl.shape[1]
        with engine_for(ped) as e:
            host = e.run_pl(fs.ES, pl, fl)
            d_pl, d_fl = torch.from_numpy(pl.view(np.int16)).cuda(), torch.from_numpy(fl).cuda()
            d_post = torch.empty((V, S, 3), dtype=torch.float64, device="cuda")
            d_single = torch.empty_like(d_post)
            d_gt = torch.empty((V, S), dtype=torch.uint8, device="cuda")
            d_st = torch.empty(V, dtype=torch.uint8, device="cuda")
            stream = torch.cuda.current_stream().cuda_stream
            e.run_pl_device(fs.ES, V, d_pl.data_ptr(), d_fl.data_ptr(), d_post.data_ptr(), d_single.data_ptr(), d_gt.data_ptr(),
                            d_st.data_ptr(), stream=stream)
            torch.cuda.synchronize()
            assert np.array_equal(d_post.cpu().numpy(), host.post, equal_nan=True) and np.array_equal(d_single.cpu().numpy(), host.single, equal_nan=True)
            assert np.array_equal(d_gt.cpu().numpy(), host.gt) and np.array_equal(d_st.cpu().numpy(), host.status)
            d_post.zero_()
            e.run_pl_device(fs.ES, V, d_pl.data_ptr(), d_fl.data_ptr(), d_post.data_ptr(), None, d_gt.data_ptr(), d_st.data_ptr(), stream=stream)
            torch.cuda.synchronize()
            assert np.array_equal(d_post.cpu().numpy(), host.post, equal_nan=True)


@pytest.mark.parametrize("n_children,stream", [(1, None), (1, "5"), (1, "-1"), (2, None), (3, "3")])
def test_tile_lists_of_the_compact_nuclear_kernel(n_children, stream, monkeypatch):
    """A device-resident batch large enough for the nuclear kernel to give every block a LIST of tiles (es_nuclear_kernel.cuh,
    es_nuclear_stream_kernel: double-buffered TMA input, stores left in flight): list lengths 4 / 8 (defaults), 5 and 3 (list
    ends that are no multiple of anything), the one-tile kernel, a ragged last tile, chrX / Known flags and failing variants,
    with and without `single`, a flags array that is not 16-byte aligned (the one-tile kernel takes that) -- the oracle's bytes."""
    torch = pytest.importorskip("torch")
    if stream is not None:
        monkeypatch.setenv("FAMSEQ_ES_STREAM", stream)
    rows = [(1, 0, 0, 1), (2, 0, 0, 2)] + [(3 + k, 2, 1, 1 + k % 2) for k in range(n_children)]
    ped = synth._mk(rows)
    V = 1_500_013
    pl, fl = pl_batch(ped, V, seed=60 + n_children, x_fraction=0.3)
    S = pl.shape[1]
    want = O.run(ped, ped.sequenced_cols(), O.pl_table()[pl], fl, method=O.ES)
    ok = want["status"] == 0
    assert 0 < (~ok).sum() < V // 10
    d_pl = torch.from_numpy(pl.view(np.int16)).cuda()
    d_fl_padded = torch.zeros(V + 16, dtype=torch.uint8, device="cuda")
    d_post = torch.empty((V, S, 3), dtype=torch.float64, device="cuda")
    d_single = torch.empty_like(d_post)
    d_gt = torch.empty((V, S), dtype=torch.uint8, device="cuda")
    d_st = torch.empty(V, dtype=torch.uint8, device="cuda")
    stream_h = torch.cuda.current_stream().cuda_stream
    with engine_for(ped) as e:
        for flag_offset, single in ((0, True), (0, False), (1, True)):
            d_fl = d_fl_padded[flag_offset:flag_offset + V]
            d_fl.copy_(torch.from_numpy(fl))
            assert d_fl.data_ptr() % 16 == flag_offset
            d_post.fill_(-1.0), d_single.fill_(-1.0), d_gt.fill_(77), d_st.fill_(77)
            e.run_pl_device(fs.ES, V, d_pl.data_ptr(), d_fl.data_ptr(), d_post.data_ptr(), d_single.data_ptr() if single else None, d_gt.data_ptr(),
                            d_st.data_ptr(), stream=stream_h)
            torch.cuda.synchronize()
            what = f"C={n_children} stream={stream} flags+{flag_offset} single={single}"
            assert np.array_equal(d_st.cpu().numpy(), want["status"]), what
            assert np.array_equal(d_post.cpu().numpy()[ok], want["post"][ok]), what
            assert np.array_equal(d_gt.cpu().numpy()[ok], want["gt"][ok].astype(np.uint8)), what
            assert not d_post.cpu().numpy()[~ok].any() and (d_gt.cpu().numpy()[~ok] == 255).all(), what
            if single:
                assert np.array_equal(d_single.cpu().numpy()[ok], want["single"][ok]), what
                assert not d_single.cpu().numpy()[~ok].any(), what


# ---- one engine over several GPUs (fs_create_multi) ---------------------------------------------------------
def _multi_devices():
    n = fs.device_count()
    return [[0]] + ([list(range(n))] if n > 1 else [])


@pytest.mark.parametrize("devices", _multi_devices())
def test_multi_device_engine_returns_the_bytes_of_one_gpu(devices):
    """fs_create_multi cuts every batch into one contiguous slice per GPU; outputs land in order in the caller's
    buffers and the bytes do not depend on the number of GPUs (MCMC: streams keyed by the global variant index)."""
    ped = synth.half_sibs()
    pl, fl = pl_batch(ped, 30_011, seed=17)
    lk = synth.pl_table()[pl]
    with engine_for(ped) as one:
        es, mc = one.run(fs.ES, lk, fl), one.run(fs.MCMC, lk[:2500], fl[:2500], burn=10, rep=50, seed=3, v_offset=40)
        bn = one.run(fs.BN, lk[:2100], fl[:2100])
    with engine_for(ped, device=devices) as many:
        assert many.info()["n_devices"] == len(devices)
        same_bytes(many.run(fs.ES, lk, fl), es)
        same_bytes(many.run_pl(fs.ES, pl, fl), es)
        same_bytes(many.run_pl(fs.ES, pl, fl, want_single=False), es, single=False)
        same_bytes(many.run(fs.MCMC, lk[:2500], fl[:2500], burn=10, rep=50, seed=3, v_offset=40), mc)
        same_bytes(many.run(fs.BN, lk[:2100], fl[:2100]), bn)
        same_bytes(many.run(fs.ES, lk[:5], fl[:5]), fs.Result(es.post[:5], es.single[:5], es.gt[:5], es.status[:5]))
        assert many.info()["kernel_launches"] >= 5
        if len(devices) > 1:
            with pytest.raises(fs.FamSeqError):  # device buffers belong to one GPU
                many.run_device(fs.ES, 1, 0, 0, 0, 0, 0, 0)


def test_multi_device_engine_rejects_bad_device_lists():
    ped = synth.trio()
    with pytest.raises(fs.FamSeqError):
        engine_for(ped, device=[0, -1])
    with pytest.raises(fs.FamSeqError):
        engine_for(ped, device=[])
    with engine_for(ped, device=[0, 0, 0]) as e:  # one GPU listed three times: three independent pipelines on it
        assert e.info()["n_devices"] == 3


def test_errors_leave_no_copy_in_flight():
    """A failing call (ES on a looped pedigree) in the middle of host-buffer traffic: the engine reports the error, and
    the next call on the same engine and buffers works and returns the right bytes."""
    ped = synth.cousins_loop()
    lk, fl = synth.synth_likelihoods(ped, 3000, seed=2)
    with engine_for(ped) as e:
        first = e.run(fs.BN, lk[:64], fl[:64])
        for _ in range(3):
            with pytest.raises(fs.FamSeqError) as ei:
                e.run(fs.ES, lk, fl)
            assert ei.value.code == -4
            with pytest.raises(fs.FamSeqError):
                e.run(fs.MCMC, lk, fl, burn=5, rep=0)  # rep must be >= 1
        again = e.run(fs.BN, lk[:64], fl[:64])
    same_bytes(again, first)
