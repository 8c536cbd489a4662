"""world_size-2 `gloo` test of the multi-rank host logic (no GPU): slice assignment, slice-wise input
generation, the global-site keying of the Gibbs stream, output gathering and the max-over-ranks timing.
The per-slice compute is done by the C oracle here (it is the checker, and shares the engine's
(seed, global site) Philox keying); on a GPU box the same logic drives one engine per rank (bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from famseq_b200 import sharding, synth
from oracle import oracle as O

V = 1501  # odd on purpose: the two ranks own slices of different sizes


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ped = synth.half_sibs()
    lo, hi = sharding.shard_range(V, rank, world)
    lk, fl = synth.synth_likelihoods(ped, hi - lo, seed=11, v0=lo, x_fraction=0.2)  # each rank generates only its slice
    es = O.run(ped, ped.sequenced_cols(), lk, fl, method=O.ES)["post"]
    mc = O.run(ped, ped.sequenced_cols(), lk[:40], fl[:40], method=O.MCMC, burn=10, rep=100, rng=O.RNG_PHILOX, seed=5,
               v_offset=lo)["post"]
    all_es = sharding.gather_to_rank0(dist, es, V)
    slowest = sharding.max_over_ranks(dist, 10.0 * (rank + 1))
    dist.barrier()
    if rank == 0:
        np.save(out + ".es.npy", all_es)
    np.save(out + f".mc{rank}.npy", mc)
    np.save(out + f".t{rank}.npy", np.array([slowest]))
    dist.destroy_process_group()


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 1500, 10_000_001):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_two_ranks_reproduce_the_single_process_result(tmp_path):
    out = str(tmp_path / "r")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    ped = synth.half_sibs()
    lk, fl = synth.synth_likelihoods(ped, V, seed=11, x_fraction=0.2)
    want = O.run(ped, ped.sequenced_cols(), lk, fl, method=O.ES)["post"]
    assert np.array_equal(np.load(out + ".es.npy"), want)  # gathered in site order, identical bytes
    # the Gibbs stream is keyed by the global site index: rank 1's first sites equal sites 751.. of a single run
    lo1, _ = sharding.shard_range(V, 1, 2)
    single = O.run(ped, ped.sequenced_cols(), lk[lo1:lo1 + 40], fl[lo1:lo1 + 40], method=O.MCMC, burn=10, rep=100,
                   rng=O.RNG_PHILOX, seed=5, v_offset=lo1)["post"]
    assert np.array_equal(np.load(out + ".mc1.npy"), single)
    assert not np.array_equal(np.load(out + ".mc0.npy"), np.load(out + ".mc1.npy"))
    # timing is the max over ranks on every rank
    assert np.load(out + ".t0.npy")[0] == 20.0 and np.load(out + ".t1.npy")[0] == 20.0
