#!/bin/bash
# N-GPU visit for the shard mode only: both sharded modes against the single-device table, strong scaling of C4 / C5.  usage: tools/gpu_strong.sh <tag> <ngpus>
tag=${1:-s8}
n=${2:-8}
out=gpurun_out/$tag
mkdir -p $out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 400 $R --master-port 29651 tests/run_sharded.py > $out/sharded.log 2>&1; echo "sharded exit $?"; tail -2 $out/sharded.log
timeout 600 $R --master-port 29653 tools/bench_strong.py c4 c5 --steps 10 > $out/strong_n$n.jsonl 2> $out/strong_n$n.err; echo "strong exit $?"
python - <<PY
import json
for l in open('$out/strong_n$n.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print('%s N=%d %-34s %.3f ms rows %d  %s' % (d['config'], d['n_gpus'], d['mode'], d['device_ms'], d['rows'], {k: round(v, 3) for k, v in d['stage_ms_rank0'].items()}))
PY
grep -v "^\[W\|^W1\|^$\|OMP_NUM\|\*\*\*\*" $out/strong_n$n.err | tail -6
