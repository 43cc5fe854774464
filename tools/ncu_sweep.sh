# usage: bash tools/ncu_sweep.sh <tag>   -- one full ncu capture of the C2 sweep kernel + raw csv summary into gpurun_out/
tag=${1:-cur}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ek_sweep -s 3 -c 1 -f -o gpurun_out/prof_sweep_$tag \
    python bench.py --steps 2 --warmup 3 --skip-backward --skip-cpu > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/prof_sweep_$tag.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > gpurun_out/ncu_${tag}_summary.txt
tail -40 gpurun_out/ncu_${tag}_summary.txt
