#!/bin/bash
tag=${1:-ncu5}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k_filter|k_pairs|k_prim|k_cell|k_fill|k_rows|k_sort|k_count|k_final|k_correct|k_mat|k_spec' --launch-skip 60 -c 60 --csv --log-file $out/launches.csv \
	python tools/bench_strong.py c5 --steps 3 > $out/launches_run.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('$out/launches.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
for r in rows[h+2:h+2+45]:
    print('%-60s %s %s'%(r[4][:58], r[-2], r[-1]))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_filter' --launch-skip 6 -c 1 -f -o $out/filter \
	python tools/bench_strong.py c5 --steps 2 > $out/ncu_full.log 2>&1
ncu -i $out/filter.ncu-rep --page raw --csv > $out/filter_raw.csv 2>/dev/null
python tools/ncu_summary.py $out/filter_raw.csv 2>&1 | head -40
