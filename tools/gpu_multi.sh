#!/bin/bash
# one multi-GPU box visit: the sharded-match equality test and bench.py under torchrun
# usage: tools/gpu_multi.sh <tag> <ngpus> [steps]
tag=${1:-multi}
n=${2:-2}
steps=${3:-20}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/smi.txt 2>&1
nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29641 tests/run_sharded.py > $out/sharded.log 2>&1; echo "sharded exit $?"
tail -3 $out/sharded.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus $n --steps $steps --warmup 5 > $out/bench_n$n.json 2> $out/bench_n$n.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=[json.loads(l) for l in open('$out/bench_n$n.json') if l.startswith('{')][-1]
    s=d['config']['stage_ms']
    print('N=%d value %.4e/s step %.4f ms | e2e %.4e/s %.2f ms | with table allgather: %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], json.dumps(d.get('with_table_allgather'))[:600]))
except Exception as e:
    print('bench parse failed', e); print(open('$out/bench_n$n.err').read()[-3000:])
PY
if [ -n "$STRONG" ]; then
	timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29643 tools/bench_strong.py $STRONG > $out/strong_n$n.jsonl 2> $out/strong_n$n.err; echo "strong exit $?"
	python - <<PY
import json
for l in open('$out/strong_n$n.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print('%s N=%d %-15s %.3f ms rows %d  %.3e rows/s' % (d['config'], d['n_gpus'], d['mode'], d['device_ms'], d['rows'], d['associations_per_s']))
PY
	grep -v "^\[W\|^W1\|^$\|OMP_NUM\|\*\*\*\*" $out/strong_n$n.err | tail -15
fi
