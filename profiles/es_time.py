"""Times ES kernels device-resident on nuclear families with C children (quads, sibships) and on named pedigrees:
    python profiles/es_time.py nuclear 2 5000000        python profiles/es_time.py ped14 1000000
    python profiles/es_time.py nuclear 1 10000000 compact | compact_no_single     (uint16 PL input: 55 S + 2 / 31 S + 2 bytes)
Prints variants/s and the fraction of the HBM roofline (73 S + 2 bytes per variant).  Tuning aid, not a bench value."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import famseq_b200 as fs  # noqa: E402
from famseq_b200 import synth  # noqa: E402

what = sys.argv[1]
if what == "nuclear":
    C, V = int(sys.argv[2]), int(sys.argv[3])
    ped = synth._mk([(1, 0, 0, 1), (2, 0, 0, 2)] + [(3 + k, 2, 1, 1 + k % 2) for k in range(C)])
else:
    ped, V = synth.PEDIGREES[what](), int(sys.argv[2])
layout = sys.argv[-1] if sys.argv[-1].startswith("compact") else "canonical"
if layout == "canonical":
    lk, fl = synth.synth_likelihoods(ped, V, 20261018 + 2)
    d_lk = torch.from_numpy(lk).cuda()
else:
    import numpy as np
    lk, fl = synth.synth_pl(ped, V, 20261018 + 2)
    d_lk = torch.from_numpy(lk.astype(np.uint16).view(np.int16)).cuda()
S = lk.shape[1]
d_fl = torch.from_numpy(fl).cuda()
d_post = torch.empty((V, S, 3), dtype=torch.float64, device="cuda")
d_single = torch.empty_like(d_post) if layout != "compact_no_single" else None
bytes_per_variant = {"canonical": 73 * S + 2, "compact": 55 * S + 2, "compact_no_single": 31 * S + 2}[layout]
d_gt = torch.empty((V, S), dtype=torch.uint8, device="cuda")
d_st = torch.empty(V, dtype=torch.uint8, device="cuda")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6538.0
with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, ped.sequenced_cols(), device=0) as e:
    def step():
        call = e.run_device if layout == "canonical" else e.run_pl_device
        call(fs.ES, V, d_lk.data_ptr(), d_fl.data_ptr(), d_post.data_ptr(), d_single.data_ptr() if d_single is not None else None, d_gt.data_ptr(),
             d_st.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    info = e.info()
gbs = bytes_per_variant * V / ms / 1e6
print(f"{' '.join(sys.argv[1:])}: S={S} {V / ms * 1e3:.4g} variants/s, {ms:.3f} ms, {gbs:.0f} GB/s = {gbs / peak:.3f} of HBM, jit_launches={info['jit_launches']}")
