"""How to reassemble the sharded table fastest: the candidates for parallel.allgather_table measured side by side on the
bench table (12 columns x ~6.15e6 rows per rank, counts uneven by a few hundred rows) and on a small one (15 x 1.3e5).
Run under torch.distributed.run; rank 0 prints one JSON line per method."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
	from nway_b200 import parallel
	local = int(os.environ.get('LOCAL_RANK', '0'))
	world = int(os.environ.get('WORLD_SIZE', '1'))
	rank = int(os.environ.get('RANK', '0'))
	torch.cuda.set_device(local)
	dev = torch.device('cuda', local)
	dist.init_process_group('nccl', device_id=dev)
	for label, ncols, base in (('bench table', 12, 6149842), ('small table', 15, 126473)):
		counts = [base + 137 * ((r * 7) % 5) - 200 for r in range(world)]
		offs = parallel.row_offsets(counts)
		stride = counts[rank] + counts[rank] // 16 + 1024
		store = torch.zeros((ncols, stride), dtype=torch.int64, device=dev)
		store[:, :counts[rank]] = rank + 1
		local_t = store[:, :counts[rank]]
		total = sum(counts)
		out = torch.empty((ncols, total), dtype=torch.int64, device=dev)
		maxc = max(counts)
		methods = {}

		def grouped():   # what parallel.allgather_table did before this measurement
			out[:, offs[rank]:offs[rank] + counts[rank]].copy_(local_t)
			send, recv = [], []
			for peer in range(world):
				if peer != rank:
					send += [(local_t[k], peer) for k in range(ncols)]
					recv += [(out[k, offs[peer]:offs[peer] + counts[peer]], peer) for k in range(ncols)]
			parallel._exchange(send, recv)
		methods['grouped send/recv per (column, peer)'] = grouped

		def packed():
			parallel.allgather_table(local_t, counts, out=out)
		methods['packed: one message per peer + unpack copies'] = packed

		def uneven_allgather():
			out[:, offs[rank]:offs[rank] + counts[rank]].copy_(local_t)
			for k in range(ncols):
				dist.all_gather([out[k, offs[r]:offs[r] + counts[r]] for r in range(world)], local_t[k])
		methods['dist.all_gather with uneven views, per column'] = uneven_allgather

		pad = torch.empty((world, ncols, maxc), dtype=torch.int64, device=dev)
		mine = torch.zeros((ncols, maxc), dtype=torch.int64, device=dev)

		def padded_whole():
			mine[:, :counts[rank]].copy_(local_t)
			dist.all_gather_into_tensor(pad.view(-1), mine.view(-1))
			for r in range(world):
				out[:, offs[r]:offs[r] + counts[r]].copy_(pad[r, :, :counts[r]])
		methods['ONE padded ncclAllGather of the whole shard + unpack copies'] = padded_whole

		def padded_collective_only():
			dist.all_gather_into_tensor(pad.view(-1), mine.view(-1))
		methods['(the padded ncclAllGather alone, no copies)'] = padded_collective_only

		for name, fn in methods.items():
			try:
				for _ in range(2):
					fn()
				torch.cuda.synchronize()
				dist.barrier()
				e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
				e0.record()
				for _ in range(5):
					fn()
				e1.record()
				torch.cuda.synchronize()
				t = torch.tensor([e0.elapsed_time(e1) / 5], dtype=torch.float64, device=dev)
				dist.all_reduce(t, op=dist.ReduceOp.MAX)
				ok = bool((out[:, offs[-1]:offs[-1] + counts[-1]] == world).all()) if 'alone' not in name else None
				if rank == 0:
					ms = float(t.item())
					recv = (total - counts[rank]) * 8 * ncols
					print(json.dumps(dict(table=label, n_gpus=world, method=name, ms=ms, received_GBs_per_gpu=recv / (ms * 1e-3) / 1e9, correct=ok)))
			except Exception as e:
				if rank == 0:
					print(json.dumps(dict(table=label, method=name, error=repr(e)[:300])))
		del store, out, pad, mine
		torch.cuda.empty_cache()
	dist.destroy_process_group()


if __name__ == '__main__':
	main()
