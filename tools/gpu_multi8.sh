#!/bin/bash
# one N-GPU visit with the final code: both sharded modes against the single-device table, bench.py, strong scaling of C4 / C5,
# the host <-> device ceiling.  usage: tools/gpu_multi8.sh <tag> <ngpus> [steps]
tag=${1:-m8}
n=${2:-8}
steps=${3:-20}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 600 $R --master-port 29651 tests/run_sharded.py > $out/sharded.log 2>&1; echo "sharded exit $?"; tail -2 $out/sharded.log
timeout 900 $R --master-port 29652 bench.py --gpus $n --steps $steps --warmup 5 > $out/bench_n$n.json 2> $out/bench_n$n.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=[json.loads(l) for l in open('$out/bench_n$n.json') if l.startswith('{')][-1]
    print('N=%d value %.4e/s step %.4f ms | e2e %.4e/s %.2f ms | with table allgather: %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], json.dumps({k: v for k, v in (d.get('with_table_allgather') or {}).items() if k != 'how'})))
except Exception as e:
    print('bench parse failed', e); print(open('$out/bench_n$n.err').read()[-3000:])
PY
timeout 1200 $R --master-port 29653 tools/bench_strong.py c4 c5 --steps 10 > $out/strong_n$n.jsonl 2> $out/strong_n$n.err; echo "strong exit $?"
python - <<PY
import json
for l in open('$out/strong_n$n.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print('%s N=%d %-15s %.3f ms rows %d  %s' % (d['config'], d['n_gpus'], d['mode'], d['device_ms'], d['rows'], {k: round(v, 3) for k, v in d['stage_ms_rank0'].items()}))
PY
grep -v "^\[W\|^W1\|^$\|OMP_NUM\|\*\*\*\*" $out/strong_n$n.err | tail -6
timeout 300 $R --master-port 29654 tools/pcie_ceiling.py > $out/pcie_n$n.json 2> $out/pcie_n$n.err; tail -1 $out/pcie_n$n.json
