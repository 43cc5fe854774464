# usage: bash tools/cfgsweep.sh "V,T,stages,ctas" ...   -- times the C2 sweep for forced configurations
for cfg in "$@"; do
  echo "== $cfg"
  EK_CFG=$cfg timeout 300 python bench.py --steps 20 --warmup 3 --skip-backward --skip-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
