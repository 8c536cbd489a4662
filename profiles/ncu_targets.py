"""Launches exactly the kernels the round-2 ncu captures look at, once each, device-resident, at the bench's sizes:
    python profiles/ncu_targets.py es      # es_nuclear_kernel<1,32,...>: canonical, compact_in, compact_in_no_single (10 M variants each)
    python profiles/ncu_targets.py bn      # bn_kernel on ped14 (20 000 variants)
    python profiles/ncu_targets.py mcmc    # famseq_gibbs on ped40, 1 000 + 10 000 sweeps (37 888 variants = one wave)
    python profiles/ncu_targets.py es14    # famseq_es on ped14 (4 M variants)
Run under `ncu -k regex:<kernel> ...` (profiles/ncu_capture_r2.sh); numbers printed by this script are not bench values."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import famseq_b200 as fs  # noqa: E402
from famseq_b200 import synth  # noqa: E402


def run(ped, method, V, compact, single, burn=0, rep=0):
    pl, fl = synth.synth_pl(ped, V, 20261018 + method)
    S = pl.shape[1]
    d_in = torch.from_numpy(pl.astype(np.uint16).view(np.int16)).cuda() if compact else torch.from_numpy(synth.pl_to_likelihood(pl)).cuda()
    d_fl = torch.from_numpy(fl).cuda()
    d_post = torch.empty((V, S, 3), dtype=torch.float64, device="cuda")
    d_single = torch.empty_like(d_post) if single else None
    d_gt = torch.empty((V, S), dtype=torch.uint8, device="cuda")
    d_st = torch.empty(V, dtype=torch.uint8, device="cuda")
    with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, ped.sequenced_cols(), device=0) as e:
        call = e.run_pl_device if compact else e.run_device
        for _ in range(2):
            call(method, V, d_in.data_ptr(), d_fl.data_ptr(), d_post.data_ptr(), d_single.data_ptr() if single else None, d_gt.data_ptr(),
                 d_st.data_ptr(), burn=burn, rep=rep, seed=1, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    print("ran", method, V, compact, single, "failed", int(d_st.sum().item()))


what = sys.argv[1]
if what == "es":
    for compact, single in ((False, True), (True, True), (True, False)):
        run(synth.trio(), fs.ES, 10_000_000, compact, single)
elif what == "bn":
    run(synth.ped14(), fs.BN, 20_000, False, True)
elif what == "mcmc":
    os.environ["FAMSEQ_MCMC_JIT"] = "1"
    run(synth.ped40(), fs.MCMC, 37_888, False, True, 1000, 10000)
elif what == "es14":
    os.environ["FAMSEQ_ES_JIT"] = "1"
    run(synth.ped14(), fs.ES, 4_000_000, False, True)
