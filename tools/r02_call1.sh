#!/bin/bash
# round-2 GPU call 1: library probe, full GPU test suite, bench, launch list, ncu capture of the analysis kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
  echo "== probe for a libsharp2 / healpy / ducc0 on the GPU box ($(date -u))"
  for mod in healpy ducc0 pysharp libsharp; do python -c "import $mod; print('$mod', getattr($mod,'__version__','?'), $mod.__file__)" 2>&1 | tail -1; done
  find / \( -name 'libsharp*' -o -name 'libhealpix*' -o -name 'libchealpix*' -o -name 'ducc0*' -o -name 'healpy*' \) -not -path '/proc/*' -not -path '/sys/*' 2>/dev/null | grep -v sharpyuv | head -20
  echo "-- mpirun / gfortran:"; which mpirun mpiexec gfortran 2>&1
  echo "-- pip list | grep -i -E 'healpy|ducc|sharp|astropy'"; python -m pip list 2>/dev/null | grep -i -E 'healpy|ducc|sharp|astropy'
  ls baseline/_ref 2>&1 | head -3
  echo "== end probe"
  nproc; lscpu | grep -E 'Model name|Socket|NUMA|Thread|Core' ; free -g | head -2
  nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
} > gpurun_out/r02_probe.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests1.log
tail -5 gpurun_out/r02_tests1.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02_bench1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches1.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cg --no-batch --no-conviqt --no-parity --e2e-steps 1 > gpurun_out/r02_ncu_l.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'anal2_kernel|anal0_kernel' -s 6 -c 2 -o gpurun_out/r02_anal_a -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cg --no-batch --no-conviqt --no-parity --e2e-steps 1 > gpurun_out/r02_ncu_a.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
