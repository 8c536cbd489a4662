# the generated Gibbs kernel against the table-driven one on small pedigrees (is the automatic switch always a win?)
python - <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import famseq_b200 as fs
from famseq_b200 import synth
for name, V in (("trio", 400000), ("half_sibs", 200000), ("ped14", 150000), ("ped40", 60000)):
    ped = synth.PEDIGREES[name]()
    lk, fl = synth.synth_likelihoods(ped, V, seed=3, x_fraction=0.0)
    res = {}
    for mode in ("0", "1"):
        os.environ["FAMSEQ_MCMC_JIT"] = mode
        with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, ped.sequenced_cols(), device=0) as e:
            e.run(fs.MCMC, lk[:1000], fl[:1000], burn=10, rep=10)          # warm-up (and compile)
            r = e.run(fs.MCMC, lk, fl, burn=100, rep=1000, seed=5)
            ms = e.last_kernel_ms()
        res[mode] = (ms, r)
    same = np.array_equal(res["0"][1].post, res["1"][1].post, equal_nan=True)
    print(f"{name:10s} V={V} generic {res['0'][0]:.1f} ms, generated {res['1'][0]:.1f} ms, x{res['0'][0] / res['1'][0]:.2f}, same bytes: {same}")
PY
