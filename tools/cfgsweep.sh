mkdir -p gpurun_out
for cfg in "16,128,2,2" "16,128,1,3" "16,128,1,4" "16,64,1,6" "16,64,2,4" "16,256,1,2" "8,256,1,4" "8,256,2,2" "8,128,1,8" "8,128,2,4"; do
  echo "== $cfg"
  EK_CFG=$cfg timeout 300 python bench.py --steps 5 --warmup 3 --skip-backward --skip-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
