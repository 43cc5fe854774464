#!/bin/bash
# round 2, GPU call 2: smoke of the new fast kernel first (stop if it fails), GPU test-suite, bench, ncu captures of C2 and C3
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== smoke"
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
[ "${PIPESTATUS[0]}" = "0" ] || { echo "smoke failed: stopping"; exit 1; }
echo "== pytest -m gpu"
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2a_pytest.log
echo "== bench"
timeout -s KILL 600 python bench.py --steps 200 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 3500 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
echo "== ncu C2"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:ek_fast -s 3 -c 1 -f -o gpurun_out/r2a_prof_c2 \
    python bench.py --steps 2 --warmup 3 --skip-backward --skip-cpu --skip-e2e --skip-extras > gpurun_out/r2a_ncu_c2.log 2>&1
ncu -i gpurun_out/r2a_prof_c2.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > gpurun_out/r2a_ncu_c2_summary.txt; head -30 gpurun_out/r2a_ncu_c2_summary.txt
echo "== ncu C3"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:ek_fast -s 14 -c 1 -f -o gpurun_out/r2a_prof_c3 \
    python bench.py --steps 2 --warmup 3 --skip-backward --skip-cpu --skip-e2e > gpurun_out/r2a_ncu_c3.log 2>&1
ncu -i gpurun_out/r2a_prof_c3.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > gpurun_out/r2a_ncu_c3_summary.txt; head -30 gpurun_out/r2a_ncu_c3_summary.txt
