python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -q --timeout 1200 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1_v.json 2> gpurun_out/bench_r1_v.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1v.csv python bench.py > gpurun_out/ncu_list_v.log 2>&1
echo rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r1_v.json").read().strip().splitlines()[-1])
print("ES", d["value"], d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["clocks"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "launches", d["gpu_launches"])
for k,m in d["methods"].items(): print(k, m["value"], m["ms_per_step"], m.get("roofline",{}).get("frac"), m.get("cpu_baseline",{}).get("value"), m.get("kernel"))
PY
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
