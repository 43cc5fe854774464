#!/bin/bash
# First hardware run after round 2 (one B200).  Round 2 itself never got this far: see DESIGN.md section 0.
#   1. kernel qualification in the foreground (what ek_init() would run in a child): agreement + timings of both sweep kernels
#   2. smoke, GPU test-suite
#   3. the bench line with the kernel ek_init() picked, and A/B lines with the choice forced (general / fast / fast at T=128)
#   4. ncu --set full of one C2 launch and one C3 launch of the fast kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== ek_qualify"
EK_FAST=0 timeout -s KILL 300 enoki_b200/ek_qualify gpurun_out/r3_qualify_timing.json 2>&1 | tail -20; echo "rc=${PIPESTATUS[0]}"
echo "== smoke"
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
[ "${PIPESTATUS[0]}" = "0" ] || { echo "smoke failed: stopping"; exit 1; }
echo "== pytest -m gpu"
timeout -s KILL 1800 python -m pytest tests -m gpu -q -rs --timeout 900 > gpurun_out/r3_pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/r3_pytest.log
echo "== bench (kernel chosen by ek_init)"
timeout -s KILL 900 python bench.py --steps 200 --warmup 5 > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; echo "bench rc=$?"; tail -c 4000 gpurun_out/r3_bench.json; tail -5 gpurun_out/r3_bench.err
for variant in "EK_FAST=0" "EK_FAST=1" "EK_FAST=1 EK_FAST_T=128"; do
  tag=$(echo "$variant" | tr ' =' '__')
  echo "== A/B: $variant"
  env $variant timeout -s KILL 400 python bench.py --steps 100 --warmup 5 --skip-backward --skip-cpu --skip-e2e > gpurun_out/r3_bench_$tag.json 2> gpurun_out/r3_bench_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r3_bench_$tag.json"))
    print("  ms/step", d["ms_per_step"], "kernel", d["roofline"]["kernel"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"],
          "C3 ms", (d.get("histogram") or {}).get("kernel_ms"), "C3 frac", ((d.get("histogram") or {}).get("roofline") or {}).get("frac"))
except Exception as e:
    print("  no line:", e)
PY
done
echo "== ncu C2 (fast kernel forced)"
EK_FAST=1 timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:ek_fast -s 3 -c 1 -f -o gpurun_out/r3_prof_c2 \
    python bench.py --steps 2 --warmup 3 --skip-backward --skip-cpu --skip-e2e --skip-extras > gpurun_out/r3_ncu_c2.log 2>&1
ncu -i gpurun_out/r3_prof_c2.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > gpurun_out/r3_ncu_c2_summary.txt; head -30 gpurun_out/r3_ncu_c2_summary.txt
echo "== ncu C3 (fast kernel forced)"
EK_FAST=1 timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:ek_fast -s 14 -c 1 -f -o gpurun_out/r3_prof_c3 \
    python bench.py --steps 2 --warmup 3 --skip-backward --skip-cpu --skip-e2e > gpurun_out/r3_ncu_c3.log 2>&1
ncu -i gpurun_out/r3_prof_c3.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > gpurun_out/r3_ncu_c3_summary.txt; head -30 gpurun_out/r3_ncu_c3_summary.txt
