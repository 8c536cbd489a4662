"""Times the MCMC kernels on one pedigree, device-resident (CUDA events), for tuning sweeps:
    FAMSEQ_JIT_PREG=.. FAMSEQ_JIT_TB=.. python profiles/mcmc_time.py [ped40] [variants] [burn] [rep] [lk: synth|flat|partial|random]
synth = the bench's data (genotypes dropped through the pedigree); flat = the same with flattened likelihoods; partial = every
third member sequenced; random = likelihoods of unrelated individuals (Mendelian inconsistencies everywhere: the chains
keep moving -- the worst case of the cached-conditional kernel).
Prints one line: pedigree, kernel, variants/s.  Not a bench value (no clocks sampling, short)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("FAMSEQ_MCMC_JIT", "1")
import famseq_b200 as fs  # noqa: E402
from famseq_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ped40"
V = int(sys.argv[2]) if len(sys.argv) > 2 else 300_000
burn = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
rep = int(sys.argv[4]) if len(sys.argv) > 4 else 10000
kind = sys.argv[5] if len(sys.argv) > 5 else "synth"
ped = synth.PEDIGREES[name]()
cols = ped.sequenced_cols()
if kind == "random":
    lk, fl = synth.synth_likelihoods(synth._mk([(i, 0, 0, 1) for i in range(1, len(cols) + 1)]), V, 20261018 + 3)
else:
    lk, fl = synth.synth_likelihoods(ped, V, 20261018 + 3)
if kind == "partial":  # every third member sequenced: the others flip all the time
    cols = cols[::3]
    lk = np.ascontiguousarray(lk[:, ::3])
if kind == "flat":
    lk = np.sqrt(np.sqrt(lk))
S = len(cols)
d_lk, d_fl = torch.from_numpy(lk).cuda(), torch.from_numpy(fl).cuda()
d_post = torch.empty((V, S, 3), dtype=torch.float64, device="cuda")
d_single = torch.empty_like(d_post)
d_gt = torch.empty((V, S), dtype=torch.uint8, device="cuda")
d_st = torch.empty(V, dtype=torch.uint8, device="cuda")
with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, cols, device=0) as e:
    def step():
        e.run_device(fs.MCMC, V, d_lk.data_ptr(), d_fl.data_ptr(), d_post.data_ptr(), d_single.data_ptr(), d_gt.data_ptr(), d_st.data_ptr(),
                     burn=burn, rep=rep, seed=1, stream=torch.cuda.current_stream().cuda_stream)
    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    info = e.info()
print(f"{name} {kind} V={V} {burn}+{rep}: {V / ms * 1e3:.4g} variants/s ({ms:.1f} ms), jit_launches={info['jit_launches']} fixups={info['mcmc_fixups']} "
      f"failed={int((d_st == 1).sum().item())} env={ {k: v for k, v in os.environ.items() if k.startswith('FAMSEQ_')} }")
