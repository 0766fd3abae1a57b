#!/bin/bash
tag=${1:-one2}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -14 $out/pytest.log
cp gpurun_out/parity_report.txt $out/parity_report.txt 2>/dev/null
timeout 900 python tools/bench_strong.py c4 c5 --steps 10 > $out/strong_n1.jsonl 2> $out/strong_n1.err; echo "strong exit $?"
python - <<PY
import json
for l in open('$out/strong_n1.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print('%s N=%d %-15s %.3f ms rows %d  %s' % (d['config'], d['n_gpus'], d['mode'], d['device_ms'], d['rows'], {k: round(v, 3) for k, v in d['stage_ms_rank0'].items()}))
PY
grep -v "^\[W\|^W1\|^$\|OMP_NUM\|\*\*\*\*" $out/strong_n1.err | tail -8
timeout 600 python tools/bench_configs.py c1 c2 > $out/configs.jsonl 2> $out/configs.err; cat $out/configs.jsonl | cut -c1-400
