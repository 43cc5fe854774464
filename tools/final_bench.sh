# default bench line + reference arm, summarised (used for the last check of a round)
timeout 500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "reference rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_final.json")); r = json.load(open("gpurun_out/bench_ref_final.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["backward"]["roofline"]["frac"], d["histogram"]["kernel_ms"], d["clocks"])
print(d["cpu_baseline"]); print(d.get("small"))
print(r["value"], r["ms_per_step"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["sample"][:70])
PY
