#!/bin/bash
# the round's final single-GPU visit: all gpu tests, both bench arms, ncu launch list, full captures of the two hot kernels and of
# the k_pairs memory-system skeleton, the C3 variants, the small configurations.  usage: tools/gpu_final.sh <tag>
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -4 $out/pytest.log
cp gpurun_out/parity_report.txt $out/parity_report.txt 2>/dev/null
timeout 600 python bench.py --steps 200 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench exit $?"
cat $out/bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
cat $out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches.csv \
	python bench.py --steps 2 --warmup 3 --no-cpu > $out/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_pairs|k_rows2' --launch-skip 6 -c 2 -f -o $out/hot \
	python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu_full.log 2>&1
ncu -i $out/hot.ncu-rep --page raw --csv > $out/hot_raw.csv 2>/dev/null
python tools/ncu_summary.py $out/hot_raw.csv > $out/hot_summary.txt 2>&1
cat $out/hot_summary.txt
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:.*k_pairs<[^0-9>]*1[^0-9>]*[01][^0-9>]*1[^0-9>]*0[^0-9>]*>.*' -c 1 -f -o $out/skeleton \
	python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu_skel.log 2>&1
ncu -i $out/skeleton.ncu-rep --page raw --csv > $out/skeleton_raw.csv 2>/dev/null
python tools/ncu_summary.py $out/skeleton_raw.csv > $out/skeleton_summary.txt 2>&1
head -30 $out/skeleton_summary.txt
for flat in 0 1; do
	timeout 600 python tools/bench_variants.py --steps 50 --flat $flat > $out/variants_flat$flat.jsonl 2> $out/variants_flat$flat.err
done
python - <<PY
import json
for flat in (0, 1):
    for l in open('$out/variants_flat%d.jsonl' % flat):
        if l.startswith('{'):
            d=json.loads(l); s=d['stage_ms']
            print('flat=%d %-10s step %.4f ms k_pairs %.1f (skeleton %.1f) k_rows %.1f grid %.1f lists %.1f us rows %d' % (d['flat'], d['variant'], d['step_ms'], 1e3*s['k_pairs'], 1e3*d['k_pairs_skeleton_ms'], 1e3*s['k_rows'], 1e3*s['grid'], 1e3*s['lists'], d['rows']))
PY
timeout 600 python tools/bench_configs.py c1 c2 > $out/configs.jsonl 2> $out/configs.err; cut -c1-330 $out/configs.jsonl
timeout 600 python tools/bench_strong.py c4 c5 --steps 10 > $out/strong_n1.jsonl 2> $out/strong_n1.err
python - <<PY
import json
for l in open('$out/strong_n1.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print('%s N=%d %-20s %.3f ms rows %d  %s' % (d['config'], d['n_gpus'], d['mode'], d['device_ms'], d['rows'], {k: round(v, 3) for k, v in d['stage_ms_rank0'].items()}))
PY
