#!/bin/bash
tag=${1:-one3}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -x -q -k "c5 or c4 or allsky or sharded" > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -4 $out/pytest.log
timeout 900 python tools/bench_strong.py c4 c5 --steps 10 > $out/strong_n1.jsonl 2> $out/strong_n1.err; echo "strong exit $?"
python - <<PY
import json
for l in open('$out/strong_n1.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print('%s N=%d %-15s %.3f ms rows %d  %s' % (d['config'], d['n_gpus'], d['mode'], d['device_ms'], d['rows'], {k: round(v, 3) for k, v in d['stage_ms_rank0'].items()}))
PY
grep -v "^\[W\|^W1\|^$\|OMP_NUM\|\*\*\*\*" $out/strong_n1.err | tail -8
