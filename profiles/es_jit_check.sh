# ES on the 14-member pedigree: message-program interpreter against the generated straight-line kernel
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "generated_kernel or es_program or random_vs_oracle or partial" 2>&1 | tail -3
for j in 0 1; do
  FAMSEQ_ES_JIT=$j FAMSEQ_JIT_VERBOSE=1 timeout 300 python bench.py --methods es14 --variants 1000000 --steps 3 --no-cpu-baseline 2> gpurun_out/esjit_$j.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); m=d['methods']['ES_ped14']; print('jit=$j', m['value'], m['ms_per_step'], m['roofline']['frac'], m.get('kernel'))"
done
grep -E "Used|spill" gpurun_out/esjit_1.err | head -3
