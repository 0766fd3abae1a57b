#!/bin/bash
# one GPU-box visit: parity tests, bench line, ncu launch list, one full capture of the two hot kernels
# usage (from the repo root, on the box): tools/gpu_check.sh <tag>
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -5 $out/pytest.log
timeout 600 python bench.py --steps 200 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench exit $?"
cat $out/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
cat $out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches.csv \
	python bench.py --steps 2 --warmup 3 --no-cpu > $out/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_pairs|k_rows2' --launch-skip 6 -c 2 -f -o $out/hot \
	python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu_full.log 2>&1
ncu -i $out/hot.ncu-rep --page raw --csv > $out/hot_raw.csv 2>/dev/null
python tools/ncu_summary.py $out/hot_raw.csv > $out/hot_summary.txt 2>&1
cat $out/hot_summary.txt
