#!/bin/bash
# smoke() + one full ncu capture of the k_pairs memory-system skeleton (the SKEL instantiation launched by bench.py's roofline leg)
tag=${1:-skel}
out=gpurun_out/$tag
mkdir -p $out

timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:.*k_pairs<[^0-9>]*1[^0-9>]*[01][^0-9>]*1[^0-9>]*0[^0-9>]*>.*' -c 1 -f -o $out/skeleton \
	python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu_skel.log 2>&1
ncu -i $out/skeleton.ncu-rep --page raw --csv > $out/skeleton_raw.csv 2>/dev/null
python tools/ncu_summary.py $out/skeleton_raw.csv > $out/skeleton_summary.txt 2>&1
head -30 $out/skeleton_summary.txt
tail -3 $out/ncu_skel.log | cut -c1-300


