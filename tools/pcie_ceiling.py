"""Host <-> device ceiling of the box, all ranks at once: aggregate bandwidth of pinned-memory copies in both directions,
alone and concurrently -- what bounds bench.py's end-to-end number at N GPUs (every rank downloads its 590 MB table per
step).  Run under torch.distributed.run with one rank per GPU; rank 0 prints one JSON line."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
	local = int(os.environ.get('LOCAL_RANK', '0'))
	world = int(os.environ.get('WORLD_SIZE', '1'))
	torch.cuda.set_device(local)
	dev = torch.device('cuda', local)
	if world > 1:
		dist.init_process_group('nccl', device_id=dev)
	nbytes = 512 << 20
	h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
	h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
	d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
	d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
	s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
	res = {}
	for name in ('h2d', 'd2h', 'both'):
		for rep in range(2):
			if world > 1:
				dist.barrier()
			torch.cuda.synchronize()
			t0 = time.perf_counter()
			for _ in range(4):
				if name in ('h2d', 'both'):
					with torch.cuda.stream(s1):
						d_a.copy_(h_in, non_blocking=True)
				if name in ('d2h', 'both'):
					with torch.cuda.stream(s2):
						h_out.copy_(d_b, non_blocking=True)
			torch.cuda.synchronize()
			if world > 1:
				dist.barrier()
			dt = time.perf_counter() - t0
		t = torch.tensor([dt], dtype=torch.float64, device=dev)
		if world > 1:
			dist.all_reduce(t, op=dist.ReduceOp.MAX)
		per_dir = 4 * nbytes * world / float(t.item()) / 1e9
		res[name] = {'aggregate_GBs_per_direction': per_dir, 'per_gpu_GBs_per_direction': per_dir / world}
	if int(os.environ.get('RANK', '0')) == 0:
		print(json.dumps({'n_gpus': world, 'copy_bytes': nbytes, 'pinned': True, 'result': res, 'cpus': os.cpu_count()}))
	if world > 1:
		dist.destroy_process_group()


if __name__ == '__main__':
	main()
