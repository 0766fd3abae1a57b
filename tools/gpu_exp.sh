#!/bin/bash
# one GPU-box visit for an experiment round: quick parity (parity + fuzz files), then bench.py for the default library and
# every variant in build/variants.  usage: tools/gpu_exp.sh <tag> [pytest args]
tag=${1:-exp}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q --durations=8 > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -15 $out/pytest.log
one() {
	python bench.py --steps ${STEPS:-100} --warmup 5 --no-cpu 2> $out/bench_$1.err | tee $out/bench_$1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['config']['stage_ms']
print('%-22s value %.4e  step %.4f ms  k_pairs %.1f  k_rows %.1f  grid %.1f  lists %.1f  total %.1f us  e2e %.2f ms  aff %s' % ('$1', d['value'], d['ms_per_step'], 1e3*s['k_pairs'], 1e3*s['k_rows'], 1e3*s['grid'], 1e3*s['lists'], 1e3*s['total'], d['e2e']['ms_per_step'], d['config'].get('cpu_affinity')))" || tail -3 $out/bench_$1.err
}
one default
for lib in build/variants/lib_*.so; do
	name=$(basename $lib .so)
	NWB_LIB=$PWD/$lib one ${name#lib_}
done
