#!/bin/bash
# round-2 (second session) final 1-GPU evidence: full GPU suite, smoke, bench (driver flags), reference arm (short), ncu launch list of the
# same command, ncu --set full of the four Legendre kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_tests_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests_final.log
tail -4 gpurun_out/r02b_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02b_bench_final.json 2> gpurun_out/r02b_bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02b_bench_final.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'], {k: round(d['e2e'][k]['ms_per_step'], 2) for k in ('pageable', 'pinned')},
      'parity', d['parity']['rel_l2'], 'cg', d['cg']['value'], 'roofline', d['roofline']['frac'], d['roofline']['frac_executed'], d['clocks'])
print(d['roofline']['kernels'])
print(d['cpu_baseline'])
print({k: d.get(k) for k in ('batch', 'conviqt')})
PY
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.err; echo "reference rc=$?"; tail -c 600 gpurun_out/r02b_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_final.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cg --no-batch --no-conviqt --no-parity --e2e-steps 1 > gpurun_out/r02b_ncu_lf.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'anal2_kernel|anal0_kernel|synth2_kernel|synth0_kernel' -s 8 -c 4 \
  -o gpurun_out/r02b_legendre python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cg --no-batch --no-conviqt --no-parity --e2e-steps 1 > gpurun_out/r02b_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r02b_legendre.ncu-rep
python tools/pair_small.py 1024 2000 2>&1 | tail -1
python tools/pair_small.py 512 1500 2>&1 | tail -1
