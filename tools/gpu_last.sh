#!/bin/bash
# the round's last, two-minute visit: the sharded test (the one GPU test that runs the set-up handshakes) and one bench line on the
# final tree (roofline.traffic from the capture whose source hash is the built library's).  usage: tools/gpu_last.sh <tag>
tag=${1:-last}
out=gpurun_out/$tag
mkdir -p $out
( timeout 55 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sharded" > $out/pytest_sharded.log 2>&1; echo "pytest exit $?" >> $out/pytest_sharded.log ) &
wait
tail -3 $out/pytest_sharded.log
timeout 60 python bench.py --steps 200 --warmup 5 --cpu-scale 0.01 > $out/bench.json 2> $out/bench.err; echo "bench exit $?"
cat $out/bench.json
