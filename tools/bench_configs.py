"""Device-time of the match path on BASELINE.json's other named configurations (bench.py measures configs[2]):

  C1  configs[0]: COSMOS XMM x OPTICAL, r = 20 arcsec                      } the committed subset of the demo catalogues
  C2  configs[1]: COSMOS XMM x OPTICAL x IRAC, r = 20 arcsec, automatic    } (tests/golden/cosmos_subset.npz: every source
      OPT MAG + IRAC mag_ch1 priors                                         } within 30 arcsec of an XMM source -- the same
      -> latency of the whole nway_match() call (these are 1797 groups: launch-latency bound)   rows as the full files)
  C4  configs[3]: synthetic 3-cat 1e6 x 1e7 x 1e7 all-sky, r = 10 arcsec, circular errors (one GPU's worth)
  C5  configs[4]: synthetic 4-cat all-sky 1e5 x 3 x 1e8 (--c5-scale of the secondaries), elliptical primary errors,
      one magnitude prior per secondary catalogue, command-line correction on

    python tools/bench_configs.py [c4] [c5] [--steps K] [--c5-scale 1.0]

Catalogues are uploaded once (nway_match, which also checks the table), then nwb_match is repeated on the resident
copies and timed with the library's CUDA events (nwb_timing).  Prints one JSON line per configuration.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(name, c5_scale):
	import nway_b200
	from tests import cases
	if name == 'c4':
		tables = cases.allsky(20260302, (1000000, 10000000, 10000000), (1.0, 0.3, 0.5))
		return tables, 10.0, dict()
	ns = int(round(1e8 * c5_scale))
	rng = np.random.default_rng(20260303)
	tables = cases.allsky(20260303, (100000, ns, ns, ns), (1.0, 0.3, 0.4, 0.5))
	n0 = 100000
	major = rng.uniform(0.5, 3.0, n0)
	tables[0]['error'] = nway_b200.ellipse_error(major, rng.uniform(0.2, 1.0, n0) * major, rng.uniform(0, 180, n0))
	tables = cases.with_mags(tables, 77, cats=(1, 2, 3), ncols=1)
	for t in tables[1:]:
		lo, hi, hs, ha = t['maghists'][0]
		t['maghists'][0] = (lo, hi, np.where(hs == 0, 0.05, hs), ha)
	return tables, 10.0, dict(unrelated_mode='cli')


def api_latency(name, steps):
	"""wall time of the public call nway_b200.nway_match() (host arrays in, pandas DataFrame out) on the COSMOS cases"""
	import nway_b200
	from nway_b200 import _lib
	from tests import cases
	case = 'cosmos2' if name == 'c1' else 'cosmos3_magauto'
	wall, rows, ms = [], 0, {}
	for k in range(steps + 1):
		tables = cases.build_case(case)
		t0 = time.perf_counter()
		res = nway_b200.nway_match(tables, 20.0, 0.9, logger=nway_b200.NullOutputLogger(), store_mag_hists=False)
		dt = time.perf_counter() - t0
		rows = len(res)
		if k >= 1:
			wall.append(dt)
			ms = _lib.get_context().timings()
	print(json.dumps(dict(config=name, case=case, sizes=[len(t['ra']) for t in tables], radius_arcsec=20.0, rows=rows,
		api_wall_ms=1e3 * float(np.median(wall)), api_wall_ms_min=1e3 * float(min(wall)), device_stage_ms_last_pass=ms,
		reference_cpu_s={'c1': 3.17, 'c2': 21.7}[name], note='reference time: SURVEY.md 8d, full demo catalogues, one core')))


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument('configs', nargs='*', default=['c4', 'c5'])
	ap.add_argument('--steps', type=int, default=10)
	ap.add_argument('--c5-scale', type=float, default=1.0)
	args = ap.parse_args()
	import nway_b200
	from nway_b200 import _lib
	for name in args.configs:
		if name in ('c1', 'c2'):
			api_latency(name, args.steps)
			continue
		t0 = time.time()
		tables, radius, kw = build(name, args.c5_scale)
		t_build = time.time() - t0
		t0 = time.time()
		res = nway_b200.nway_match(tables, radius, 0.9, logger=nway_b200.NullOutputLogger(), as_frame=False, keep_on_device=True, **kw)
		t_first = time.time() - t0
		ctx = _lib.get_context()
		fuse = True
		acc = {}
		wall = []
		rows = 0
		for k in range(args.steps + 2):
			t0 = time.perf_counter()
			rows = ctx.match(fuse_final=fuse)
			ctx.sync()
			dt = time.perf_counter() - t0
			if k >= 2:
				wall.append(dt)
				for s, v in ctx.timings().items():
					acc[s] = acc.get(s, 0.0) + v
		ms = {s: v / args.steps for s, v in acc.items()}
		b_in = sum(len(t['ra']) for t in tables) * 8 * 3
		n = len(tables)
		b_row = 8 * (n + n * (n - 1) // 2 + 9 + sum(len(t['mags']) for t in tables))
		line = dict(config=name, sizes=[len(t['ra']) for t in tables], radius_arcsec=radius, rows=rows, device_ms=ms['total'], stage_ms=ms,
			wall_ms=1e3 * float(np.median(wall)), associations_per_s=rows / (ms['total'] * 1e-3),
			sources_streamed_per_s=sum(len(t['ra']) for t in tables[1:]) / (ms['total'] * 1e-3),
			B_alg_bytes=b_in + rows * b_row, B_alg_GBs=(b_in + rows * b_row) / (ms['total'] * 1e-3) / 1e9,
			launches=ctx.launch_count(), host_build_s=t_build, first_call_s=t_first, mode=kw)
		print(json.dumps(line))


if __name__ == '__main__':
	main()
