#!/bin/bash
# one GPU-box visit for an experiment round: (optional) quick parity, bench.py for the default library and every variant in
# build/variants, (optional) an ncu launch list of the default library.
# usage: [PYTEST="tests/test_gpu_parity.py tests/test_gpu_fuzz.py"] [LAUNCHES=1] [STEPS=100] tools/gpu_exp.sh <tag>
tag=${1:-exp}
out=gpurun_out/$tag
mkdir -p $out
if [ -n "$PYTEST" ]; then
	timeout 900 python -m pytest $PYTEST -m gpu -x -q --durations=5 > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
	tail -12 $out/pytest.log
fi
one() {
	python bench.py --steps ${STEPS:-100} --warmup 5 --no-cpu 2> $out/bench_$1.err | tee $out/bench_$1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['config']['stage_ms']
print('%-16s %.4e/s step %.4f ms  k_pairs %.1f k_rows %.1f grid %.1f lists %.1f total %.1f us  e2e %.2f ms rows %d pairs %d' % ('$1', d['value'], d['ms_per_step'], 1e3*s['k_pairs'], 1e3*s['k_rows'], 1e3*s['grid'], 1e3*s['lists'], 1e3*s['total'], d['e2e']['ms_per_step'], d['config']['rows_per_gpu'], d['config']['pairs_per_gpu']))" || tail -3 $out/bench_$1.err
}
one default
shopt -s nullglob
for lib in build/variants/lib_*.so; do
	name=$(basename $lib .so)
	NWB_LIB=$PWD/$lib one ${name#lib_}
done
if [ -n "$LAUNCHES" ]; then
	timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/launches.csv \
		python bench.py --steps 2 --warmup 3 --no-cpu > $out/launches_bench.log 2>&1
	python - <<PY
import csv
rows=list(csv.reader(open('$out/launches.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
d=rows[h+2:]
for r in d[40:75]:
    print('%-70s %-10s %s'%(r[4][:68], r[8], r[-1]))
PY
fi
