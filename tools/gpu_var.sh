#!/bin/bash
# one GPU-box visit: C3 variants (orderings, per-source sigma, k_pairs skeleton) for the default library and every
# library in build/variants; usage: tools/gpu_var.sh <tag> [variants...]
tag=${1:-var}
shift
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python tools/bench_variants.py --steps 50 "$@" > $out/variants_default.jsonl 2> $out/variants_default.err
timeout 300 python tools/bench_variants.py --steps 50 --flat 1 random >> $out/variants_default.jsonl 2>> $out/variants_default.err
python - <<PY
import json
for l in open('$out/variants_default.jsonl'):
    d=json.loads(l); s=d['stage_ms']
    print('flat=%d %-10s step %.4f ms k_pairs %.1f (skeleton %.1f) k_rows %.1f grid %.1f lists %.1f us rows %d' % (d['flat'], d['variant'], d['step_ms'], 1e3*s['k_pairs'], 1e3*d['k_pairs_skeleton_ms'], 1e3*s['k_rows'], 1e3*s['grid'], 1e3*s['lists'], d['rows']))
PY
tail -3 $out/variants_default.err
shopt -s nullglob
for lib in build/variants/lib_*.so; do
	name=$(basename $lib .so); name=${name#lib_}
	NWB_LIB=$PWD/$lib timeout 300 python tools/bench_variants.py --steps 50 random > $out/variants_$name.jsonl 2> $out/variants_$name.err
	NWB_LIB=$PWD/$lib timeout 300 python tools/bench_variants.py --steps 50 --flat 1 random >> $out/variants_$name.jsonl 2>> $out/variants_$name.err
	python - <<PY
import json
for l in open('$out/variants_$name.jsonl'):
    d=json.loads(l); s=d['stage_ms']
    print('%-14s flat=%d %-10s step %.4f ms k_pairs %.1f k_rows %.1f grid %.1f lists %.1f us' % ('$name', d['flat'], d['variant'], d['step_ms'], 1e3*s['k_pairs'], 1e3*s['k_rows'], 1e3*s['grid'], 1e3*s['lists']))
PY
done
