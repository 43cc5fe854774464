#!/bin/bash
# round 2, first GPU call: ek_partition under short timeouts (a hang costs 2 minutes, not a box), then the baseline bench
mkdir -p gpurun_out
export EK_ENABLE_PARTITION=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
echo "== partition (pytest, 150 s limit)"
timeout -s KILL 150 python -m pytest tests/test_gpu_eval.py -x -q -m gpu -k "partition" 2>&1 | tail -5
echo "rc=$?"
echo "== call_check 1000 under synccheck (120 s limit)"
timeout -s KILL 120 compute-sanitizer --tool synccheck tests/cpp/call_check 1000 2>&1 | tail -8
echo "== call_check 100003 (60 s limit)"
timeout -s KILL 60 tests/cpp/call_check 100003 2>&1 | tail -4
echo "== call_check 4194304 (60 s limit)"
timeout -s KILL 60 tests/cpp/call_check 4194304 2>&1 | tail -4
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader
echo "== baseline bench"
timeout -s KILL 400 python bench.py --steps 100 --warmup 5 > gpurun_out/r2_base_bench.json 2> gpurun_out/r2_base_bench.err
tail -c 3000 gpurun_out/r2_base_bench.json
