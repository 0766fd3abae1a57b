#!/bin/bash
# usage: tools/gpu_visit.sh <tag> : quick parity (fuzz + parity tests), variants bench, ncu of the hot kernels
tag=${1:-visit}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -4 $out/pytest.log
tools/gpu_var.sh $tag
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_pairs|k_rows2' --launch-skip 6 -c 2 -f -o $out/hot \
	python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu_full.log 2>&1
ncu -i $out/hot.ncu-rep --page raw --csv > $out/hot_raw.csv 2>/dev/null
python tools/ncu_summary.py $out/hot_raw.csv > $out/hot_summary.txt 2>&1
cat $out/hot_summary.txt | head -60
