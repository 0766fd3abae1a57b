#!/bin/bash
# run bench.py once per experiment library in build/variants (on the GPU box)
for lib in build/variants/lib_*.so; do
	name=$(basename $lib .so)
	NWB_LIB=$PWD/$lib python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['config']['stage_ms']
print('%-28s value %.3e  step %.3f ms  k_pairs %.1f us  k_rows %.1f us  grid %.1f us' % ('$name', d['value'], d['ms_per_step'], 1e3*s['k_pairs'], 1e3*s['k_rows'], 1e3*s['grid']))"
done
