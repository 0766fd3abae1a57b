#!/bin/bash
# full ncu capture of the hot kernels (one launch each, after warm-up) + summaries; usage: tools/gpu_ncu.sh <tag> [kernel regex]
tag=${1:-ncu}
rx=${2:-k_pairs|k_rows2}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip 6 -c 2 -f -o $out/hot \
	python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu_full.log 2>&1
ncu -i $out/hot.ncu-rep --page raw --csv > $out/hot_raw.csv 2>/dev/null
python tools/ncu_summary.py $out/hot_raw.csv > $out/hot_summary.txt 2>&1
cat $out/hot_summary.txt
