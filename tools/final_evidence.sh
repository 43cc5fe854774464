# Round-end evidence run (one GPU): tests, default bench line, launch list and one full ncu capture per hot kernel.
# usage: bash tools/final_evidence.sh <tag>     -> files under gpurun_out/
tag=${1:-r1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_$tag.log 2>&1; tail -3 gpurun_out/pytest_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "bench ref rc=$?"
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/ncu_launches_$tag.log 2>&1
# full captures
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:ek_fast|ek_sweep' -s 3 -c 1 -f -o gpurun_out/prof_sweep_$tag \
    python bench.py --steps 2 --warmup 3 --skip-backward --skip-cpu > gpurun_out/ncu_sweep_$tag.log 2>&1
ncu -i gpurun_out/prof_sweep_$tag.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > gpurun_out/ncu_sweep_${tag}_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ek_adjoint -s 100 -c 1 -f -o gpurun_out/prof_adjoint_$tag \
    python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/ncu_adjoint_$tag.log 2>&1
ncu -i gpurun_out/prof_adjoint_$tag.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > gpurun_out/ncu_adjoint_${tag}_summary.txt
head -12 gpurun_out/ncu_adjoint_${tag}_summary.txt
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$tag.json"))
print("value", d["value"], "ms", d["ms_per_step"], "roofline", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "bwd", d["backward"]["roofline"]["frac"], "clocks", d["clocks"])
PY
