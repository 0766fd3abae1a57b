"""The table reassembly alone (nwb_gather_* / parallel.TableGather) on the bench workload's shards: every rank matches
its C3 shard once, then the gather of the resulting 12 x ~6.15e6 table is timed by itself -- count exchange, push of the
shard into every rank's table over peer memory, barrier -- with CUDA events on the context's stream, max over the ranks.

    python -m torch.distributed.run --nproc-per-node N ... tools/bench_push.py [--steps 20]

NWB_PUSH_BLOCKS_PER_SM=1..4 (read by the library) sizes the push kernel's grid.  Rank 0 prints one JSON line per variant."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument('--steps', type=int, default=20)
	ap.add_argument('--scale', type=float, default=1.0)
	args = ap.parse_args()
	import torch
	import torch.distributed as dist
	import nway_b200
	from nway_b200 import _lib, parallel
	import bench
	local = int(os.environ.get('LOCAL_RANK', '0'))
	rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
	torch.cuda.set_device(local)
	dev = torch.device('cuda', local)
	dist.init_process_group('nccl', device_id=dev)
	tables, n0 = bench.make_workload(world, scale=args.scale)
	ctx = _lib.Context(local)
	stream = torch.cuda.Stream(device=dev)
	ctx.set_stream(stream.cuda_stream)
	keep = []
	for c, t in enumerate(tables):
		arrs = [torch.from_numpy(np.ascontiguousarray(t[k], dtype=np.float64)).to(dev) for k in ('ra', 'dec', 'error')]
		keep.append(arrs)
		ctx.set_catalogue_device(c, 2, len(t['ra']), arrs[0].data_ptr(), arrs[1].data_ptr(), arrs[2].data_ptr(), t['area'])
	tab = nway_b200._scalar_tables(tables, bench.COMPLETENESS, nway_b200.NullOutputLogger())
	ctx.set_params(bench.RADIUS, tab['pc'], 0.5, _lib.UNRELATED_API)
	ctx.set_tables(tab['norm'], tab['log10e'], tab['prior'], tab['log10prior'], tab['sub_log10prior'])
	ctx.set_compat(_lib.COMPAT_FLAT_HASH)
	ctx.set_primary_range(rank * n0, n0)
	with torch.cuda.stream(stream):
		nrows = ctx.match(fuse_final=True)
		counts = parallel.exchange_counts(nrows, None, dev)
		ncols = ctx.table_layout()[2]
		total = sum(counts)
		recv = (total - nrows) * 8 * ncols
		ref = None
		for label, engine, host_counts, per_sm, store, chunk in (('as shipped: streaming stores, chunk sized by the kernel, counts exchanged on the device', 0, False, 4, 1, 0),
				('as shipped, counts known', 0, True, 4, 1, 0), ('plain stores, 64 KB chunks', 0, True, 4, 0, 8192), ('plain stores, 64 KB chunks, 2 blocks per SM', 0, True, 2, 0, 8192),
				('streaming stores, 64 KB chunks', 0, True, 4, 1, 8192), ('write-through stores, 64 KB chunks', 0, True, 4, 2, 8192),
				('plain stores, 16 KB chunks', 0, True, 4, 0, 2048), ('plain stores, 512 KB chunks', 0, True, 4, 0, 65536), ('copy engines, counts known', 1, True, 4, 0, 0)):
			os.environ['NWB_PUSH_BLOCKS_PER_SM'] = str(per_sm)
			os.environ['NWB_PUSH_STORE'] = str(store)
			os.environ['NWB_PUSH_CHUNK'] = str(chunk)
			tg = parallel.TableGather(None, local, stream=stream, engine=engine)
			tg.setup(ctx, total + 4096, ncols)
			for _ in range(3):
				table, _c = tg(ctx, counts if host_counts else None)
			stream.synchronize()
			dist.barrier()
			e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			e0.record(stream)
			for _ in range(args.steps):
				table, _c = tg(ctx, counts if host_counts else None)
			e1.record(stream)
			stream.synchronize()
			t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
			dist.all_reduce(t, op=dist.ReduceOp.MAX)
			chk = torch.stack([table[k].sum() for k in range(ncols)])
			if ref is None:
				ref = chk
				prim = table[0]
				assert bool((prim[1:] >= prim[:-1]).all()) and int(prim[-1]) == world * n0 - 1
			assert torch.equal(chk, ref)
			if rank == 0:
				ms = float(t.item())
				print(json.dumps(dict(n_gpus=world, method=label, blocks_per_sm=per_sm, store=store, chunk_elements=chunk, ms=ms, rows=total, ncols=ncols,
					received_bytes_per_gpu=recv, received_GBs_per_gpu=recv / (ms * 1e-3) / 1e9)), flush=True)
			tg.close(ctx)
	ctx.close()
	dist.destroy_process_group()


if __name__ == '__main__':
	main()
