#!/bin/bash
# one single-GPU visit: selected tests, strong-scaling tool at N = 1, bench
tag=${1:-one}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "score_rows or sharded or async" > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -5 $out/pytest.log
timeout 900 python tools/bench_strong.py c4 c5 --steps 10 > $out/strong_n1.jsonl 2> $out/strong_n1.err; echo "strong exit $?"
python - <<PY
import json
for l in open('$out/strong_n1.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print('%s N=%d %-15s %.3f ms rows %d  %s' % (d['config'], d['n_gpus'], d['mode'], d['device_ms'], d['rows'], {k: round(v, 3) for k, v in d['stage_ms_rank0'].items()}))
PY
grep -v "^\[W\|^W1\|^$\|OMP_NUM\|\*\*\*\*" $out/strong_n1.err | tail -8
timeout 600 python bench.py --steps 50 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench exit $?"; tail -c 1500 $out/bench.json; tail -3 $out/bench.err
