python - <<'PY'
import sys, os, subprocess, time, json
sys.path.insert(0, os.getcwd())
from famseq_b200 import synth
ped = synth.ped14()
ped.write("/tmp/fam.ped")
pl, fl = synth.synth_pl(ped, 200000, seed=5); fl[:] = fl & 1
synth.write_vcf("/tmp/in.vcf", ped, pl, fl)
for m in ("1", "2"):
    for rep in range(2):
        t0 = time.perf_counter()
        r = subprocess.run(["famseq_b200/bin/FamSeq", "vcf", "-vcfFile", "/tmp/in.vcf", "-pedFile", "/tmp/fam.ped", "-method", m, "-output", "/tmp/out.vcf"],
                           capture_output=True, text=True, env=dict(os.environ, FAMSEQ_STATS="1"))
        print("method", m, "wall", round(time.perf_counter() - t0, 3), r.stderr.strip().splitlines()[-1])
PY
