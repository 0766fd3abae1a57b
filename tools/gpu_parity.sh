#!/bin/bash
# one GPU-box visit: all GPU tests (parity report into gpurun_out/parity_report.txt) + a short bench line
# usage (from the repo root, on the box): tools/gpu_parity.sh <tag> [pytest args]
tag=${1:-parity}
shift
out=gpurun_out/$tag
mkdir -p $out
rm -f gpurun_out/parity_report.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=8 "$@" > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -25 $out/pytest.log
cp gpurun_out/parity_report.txt $out/parity_report.txt 2>/dev/null
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu > $out/bench.json 2> $out/bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.load(open('$out/bench.json')); s=d['config']['stage_ms']
    print('%.4e/s step %.4f ms  k_pairs %.1f k_rows %.1f grid %.1f lists %.1f total %.1f us  e2e %.2f ms rows %d' % (d['value'], d['ms_per_step'], 1e3*s['k_pairs'], 1e3*s['k_rows'], 1e3*s['grid'], 1e3*s['lists'], 1e3*s['total'], d['e2e']['ms_per_step'], d['config']['rows_per_gpu']))
except Exception as e:
    print('bench parse failed', e); print(open('$out/bench.err').read()[-2000:])
PY
