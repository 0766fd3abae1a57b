"""Build experiment variants of the library (compile-time knobs) in parallel; run them with tools/sweep.sh."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nway_b200 import build

VARIANTS = {}
for line in open(sys.argv[1]):
	line = line.strip()
	if line and not line.startswith('#'):
		name, *defs = line.split()
		VARIANTS[name] = defs

out_dir = os.path.join(ROOT, 'build', 'variants')
os.makedirs(out_dir, exist_ok=True)


def one(item):
	name, defs = item
	log = build.build_variant(os.path.join(out_dir, 'lib_%s.so' % name), defs)
	regs = [l for l in log.splitlines() if 'Used' in l]
	return name, defs


with ThreadPoolExecutor(max_workers=6) as ex:
	for name, defs in ex.map(one, VARIANTS.items()):
		print('built', name, defs)
