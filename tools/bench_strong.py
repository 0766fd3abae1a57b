"""Strong scaling of BASELINE.json configs[3] / configs[4] (C4, C5) on 1 .. 8 GPUs of one node: total work fixed, the
STREAM of the secondaries divided between the ranks (shard mode: nwb_shard_*, nway_b200.parallel.ScatterMatcher), the
matches scattered to the owners of the primaries over NVLink peer memory, then the output table reassembled with one
all-gather-v.

    python tools/bench_strong.py c4 c5 [--steps 10]                                          one GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_strong.py c4 c5      N GPUs

  C4  1e6 x 1e7 x 1e7 all-sky, r = 10 arcsec, circular errors (sigma 1.0 / 0.3 / 0.5 arcsec)
  C5  1e5 x 3 x 1e8 all-sky, r = 10 arcsec, elliptical primary errors, one magnitude prior per secondary catalogue
      (fixed 16-bin histograms), command-line correction on

The catalogues are generated ON the device with torch's generator (same seed on every rank: identical replicas; 14 GB
of columns for C5 would otherwise have to come through every rank's host).  Device time per match from CUDA events
on the matching stream, maximum over the ranks; rank 0 prints one JSON line per configuration and mode:
  "plain"    (N = 1 only) nwb_match on one GPU
  "scatter"  shard mode, table left sharded by primary blocks
  "scatter+gather"  ... plus the all-gather-v of the table to every rank (torch.distributed / NCCL, parallel.allgather_table)
  "scatter+peer gather"  ... the same reassembly over peer memory (parallel.TableGather / nwb_gather_*), row counts and
                         barrier as flag words in peer memory too ("..., NCCL counts": those two through NCCL)
  "scatter, NCCL barrier"  shard mode with an NCCL all-reduce between its halves instead of the flags in peer memory
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def device_catalogue(gen, n, sigma, dev, nmag=0):
	import torch
	ra = 360.0 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen)
	dec = torch.rad2deg(torch.asin(2.0 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen) - 1.0))
	err = torch.full((n,), float(sigma), dtype=torch.float64, device=dev)
	mags = None
	if nmag:
		mags = 22.0 + 2.0 * torch.randn(n * nmag, dtype=torch.float64, device=dev, generator=gen)
		mags[torch.rand(n * nmag, dtype=torch.float64, device=dev, generator=gen) < 0.01] = -99.0
	return ra, dec, err, mags


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument('configs', nargs='*', default=['c4', 'c5'])
	ap.add_argument('--steps', type=int, default=10)
	ap.add_argument('--c5-scale', type=float, default=1.0)
	args = ap.parse_args()
	import torch
	import torch.distributed as dist
	import nway_b200
	from nway_b200 import _lib, parallel
	from tests import cases
	rank = int(os.environ.get('RANK', '0'))
	world = int(os.environ.get('WORLD_SIZE', '1'))
	local = int(os.environ.get('LOCAL_RANK', '0'))
	torch.cuda.set_device(local)
	dev = torch.device('cuda', local)
	if 'RANK' in os.environ:
		dist.init_process_group('nccl', device_id=dev)
	else:
		os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
		os.environ.setdefault('MASTER_PORT', '29677')
		dist.init_process_group('nccl', rank=0, world_size=1, device_id=dev)
	area = 41252.96124941928
	for name in args.configs:
		gen = torch.Generator(device=dev)
		gen.manual_seed(20260302 if name == 'c4' else 20260303)
		ctx = _lib.Context(local)
		keep = []
		if name == 'c4':
			sizes, sigmas, nmag, radius, mode = (1000000, 10000000, 10000000), (1.0, 0.3, 0.5), 0, 10.0, _lib.UNRELATED_API
		else:
			ns = int(round(1e8 * args.c5_scale))
			sizes, sigmas, nmag, radius, mode = (100000, ns, ns, ns), (1.0, 0.3, 0.4, 0.5), 1, 10.0, _lib.UNRELATED_CLI
		nc = len(sizes)
		tables = []
		for c, (n, sg) in enumerate(zip(sizes, sigmas)):
			ra, dec, err, mags = device_catalogue(gen, n, sg, dev, nmag if c > 0 else 0)
			if name == 'c5':
				# elliptical mode: every catalogue carries (sigma_x, sigma_y, rho); the primary's from random ellipses
				if c == 0:
					major = 0.5 + 2.5 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen)
					minor = (0.2 + 0.8 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen)) * major
					ang = 180.0 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen)
					sx, sy, rho = nway_b200.ellipse_error(major.cpu().numpy(), minor.cpu().numpy(), ang.cpu().numpy())
					err = torch.from_numpy(np.ascontiguousarray(np.stack([sx, sy, rho]))).to(dev).reshape(-1)
				else:
					err = torch.cat([err, err, torch.zeros_like(err)])
			keep.append((ra, dec, err, mags))
			ctx.set_catalogue_device(c, nc, n, ra.data_ptr(), dec.data_ptr(), err.data_ptr(), area,
				mags_ptr=mags.data_ptr() if mags is not None else None, m=nmag if c > 0 else 0,
				err_kind=_lib.ERR_ELLIPSE if name == 'c5' else _lib.ERR_CIRCULAR)
			tables.append(dict(name='ABCD'[c], ra=range(n), area=area))
		tab = nway_b200._scalar_tables(tables, 0.9, nway_b200.NullOutputLogger())
		ctx.set_params(radius, tab['pc'], 0.5, mode)
		ctx.set_tables(tab['norm'], tab['log10e'], tab['prior'], tab['log10prior'], tab['sub_log10prior'])
		ctx.set_compat(_lib.COMPAT_FLAT_HASH)   # all-sky: the reference hashes with HEALPix, the switch is inert
		if nmag:
			from nway_b200 import magnitudeweights
			for c in range(1, nc):
				lo, hi, hs, ha = cases.fixed_hist(77 + 17 * c)
				hs = np.where(hs == 0, 0.05, hs)
				ctx.set_maghist(c, 0, *magnitudeweights.step_tables(np.array(list(lo) + [hi[-1]]), hs, ha))
		results = []

		def timed(fn, label, steps):
			for _ in range(2):
				rows = fn()
			torch.cuda.synchronize()
			dist.barrier()
			e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			st = torch.cuda.current_stream()
			e0.record(st)
			for _ in range(steps):
				rows = fn()
			e1.record(st)
			torch.cuda.synchronize()
			t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
			r = torch.tensor([rows], dtype=torch.int64, device=dev)
			dist.all_reduce(t, op=dist.ReduceOp.MAX)
			dist.all_reduce(r, op=dist.ReduceOp.SUM)
			results.append((label, float(t.item()), int(r.item())))
			return rows

		if world == 1:
			s0 = torch.cuda.Stream(device=dev)
			ctx.set_stream(s0.cuda_stream)
			with torch.cuda.stream(s0):
				timed(lambda: ctx.match(fuse_final=True), 'plain', args.steps)
			stage_plain = ctx.timings()
			ctx.set_stream(None)
		else:
			stage_plain = None
		if world > 1:   # for comparison: the barrier between the two halves as an NCCL all-reduce
			m0 = parallel.ScatterMatcher(None, local, peer_barrier=False)
			m0.setup(ctx)
			with torch.cuda.stream(m0.stream):
				timed(lambda: m0(ctx, True), 'scatter, NCCL barrier', args.steps)
			m0.close(ctx)
		matcher = parallel.ScatterMatcher(None, local)
		xbytes = matcher.setup(ctx)
		with torch.cuda.stream(matcher.stream):
			timed(lambda: matcher(ctx, True), 'scatter', args.steps)
			stage_scatter = ctx.timings()
			gtab = {}

			def with_gather():
				nr = matcher(ctx, True)
				cnts = parallel.exchange_counts(nr, None, dev)
				if gtab.get('t') is None or gtab['t'].shape[1] != sum(cnts):
					gtab['t'] = torch.empty((ctx.table_layout()[2], sum(cnts)), dtype=torch.int64, device=dev)
				if sum(cnts):
					parallel.allgather_table(ctx.table_view() if nr else torch.empty((gtab['t'].shape[0], 0), dtype=torch.int64, device=dev), cnts, out=gtab['t'])
				return nr
			timed(with_gather, 'scatter+gather', args.steps)
			prim = gtab['t'][0]
			assert bool((prim[1:] >= prim[:-1]).all()), 'the gathered table is not in primary order'
			# ... and the reassembly over peer memory (nwb_gather_*): each GPU stores its rows into every rank's table
			tot = torch.tensor([matcher(ctx, True)], dtype=torch.int64, device=dev)
			dist.all_reduce(tot)
			for label, engine in (('scatter+peer gather, NCCL counts', 0), ('scatter+peer gather', 2)):
				tg = parallel.TableGather(None, local, stream=matcher.stream, engine=engine)
				tg.setup(ctx, int(tot.item()) + 4096, ctx.table_layout()[2])
				last = {}

				def with_peer_gather():
					nr = matcher(ctx, True)
					last['t'], last['c'] = tg(ctx)
					return nr
				timed(with_peer_gather, label, args.steps)
				assert last['t'].shape == gtab['t'].shape and torch.equal(last['t'], gtab['t']), 'the two reassemblies disagree'
				tg.close(ctx)
		matcher.close(ctx)
		if rank == 0:
			base = None
			for label, ms, rows in results:
				if base is None:
					base = rows
				assert rows == base, (label, rows, base)   # every mode produces the same number of rows
				print(json.dumps(dict(config=name, n_gpus=world, mode=label, sizes=list(sizes), radius_arcsec=radius, rows=rows, device_ms=ms,
					associations_per_s=rows / (ms * 1e-3), sources_streamed_per_s=sum(sizes[1:]) / (ms * 1e-3),
					exchange_buffer_bytes=xbytes, stage_ms_rank0=stage_plain if label == 'plain' else stage_scatter)))
		ctx.close()
		del keep
		torch.cuda.empty_cache()
	dist.destroy_process_group()


if __name__ == '__main__':
	main()
