"""Variants of the bench workload C3 (BASELINE.json configs[2]) on one GPU, device time of nwb_match with resident inputs:

  random     the workload as bench.py runs it (secondaries in random order, one sigma per catalogue)
  dec        the same sources, secondary catalogue stored sorted by declination
  cell       ... sorted by (5-arcsec declination strip, ra): the order of a tiled survey catalogue
  persource  random order, every source with its own positional error (defeats the constant-sigma shortcuts of k_rows2)

and, for each ordering, the memory-system skeleton of k_pairs (nwb_bench_skeleton: same loads / atomics / stores, no
arithmetic) -- the measured floor of the access pattern.  One JSON line per variant.

    python tools/bench_variants.py [--steps 50] [variants ...]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument('variants', nargs='*', default=['random', 'dec', 'cell', 'persource'])
	ap.add_argument('--steps', type=int, default=50)
	ap.add_argument('--scale', type=float, default=1.0)
	ap.add_argument('--flat', type=int, default=0, help='1: NWB_COMPAT_FLAT_HASH on (the reference-identical row set; C3 is a flat-sky field)')
	args = ap.parse_args()
	import torch
	import nway_b200
	from nway_b200 import _lib
	import bench
	dev = torch.device('cuda', 0)
	for name in args.variants:
		tables, n0 = bench.make_workload(1, scale=args.scale)
		sec = tables[1]
		if name == 'dec':
			o = np.argsort(sec['dec'], kind='stable')
		elif name == 'cell':
			o = np.lexsort((sec['ra'], np.floor(sec['dec'] * 720.0)))
		else:
			o = None
		if o is not None:
			sec['ra'], sec['dec'] = sec['ra'][o], sec['dec'][o]
		if name == 'persource':
			rng = np.random.default_rng(5)
			tables[0]['error'] = rng.uniform(0.5, 1.5, len(tables[0]['ra']))
			sec['error'] = rng.uniform(0.1, 0.3, len(sec['ra']))
		ctx = _lib.Context(0)
		keep = []
		for c, t in enumerate(tables):
			arrs = [torch.from_numpy(np.ascontiguousarray(t[k], dtype=np.float64)).to(dev) for k in ('ra', 'dec', 'error')]
			keep.append(arrs)
			ctx.set_catalogue_device(c, 2, len(t['ra']), arrs[0].data_ptr(), arrs[1].data_ptr(), arrs[2].data_ptr(), t['area'])
		tab = nway_b200._scalar_tables(tables, bench.COMPLETENESS, nway_b200.NullOutputLogger())
		ctx.set_params(bench.RADIUS, tab['pc'], 0.5, _lib.UNRELATED_API)
		ctx.set_tables(tab['norm'], tab['log10e'], tab['prior'], tab['log10prior'], tab['sub_log10prior'])
		ctx.set_compat(_lib.COMPAT_FLAT_HASH if args.flat else 0)
		rows = ctx.match(fuse_final=True)
		assert ctx.flat_hash_applied() == bool(args.flat)
		for _ in range(5):
			ctx.match_async(fuse_final=True)
		ctx.match_wait()
		torch.cuda.synchronize()
		acc = {}
		import time
		torch.cuda.synchronize()
		t0 = time.perf_counter()
		for k in range(args.steps):
			ctx.match_async(fuse_final=True)
		rows = ctx.match_wait()
		step_ms = (time.perf_counter() - t0) / args.steps * 1e3   # back-to-back async steps, collected once: host clock around the lot
		stage = ctx.timings()
		ctx.match(fuse_final=True)
		skel = ctx.bench_skeleton(1, 10)
		skel_by_blocks = {}
		for b in (2, 3, 4, 5):
			ctx.match(fuse_final=True)
			skel_by_blocks[b] = ctx.bench_skeleton(1, 5, b)
		n1 = len(sec['ra'])
		pairs = rows - n0
		kb = n1 * 16 + pairs * 16
		print(json.dumps(dict(variant=name, flat=args.flat, rows=rows, step_ms=step_ms, rows_per_s=rows / (step_ms * 1e-3), stage_ms=stage,
			k_pairs_ms=stage['k_pairs'], k_pairs_skeleton_ms=skel, skeleton_ms_by_blocks_per_sm=skel_by_blocks, k_pairs_vs_skeleton=stage['k_pairs'] / skel if skel else None,
			k_pairs_GBs=kb / (stage['k_pairs'] * 1e-3) / 1e9, skeleton_GBs=kb / (skel * 1e-3) / 1e9)))
		ctx.close()
		del keep


if __name__ == '__main__':
	main()
