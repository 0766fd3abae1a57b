"""torchrun entry: nway_match_sharded over NCCL must equal the single-device table (run by the gpu tests)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
	import torch
	import torch.distributed as dist
	import nway_b200
	from nway_b200 import parallel
	from tests import cases
	local = int(os.environ.get('LOCAL_RANK', '0'))
	torch.cuda.set_device(local)
	dist.init_process_group('nccl', device_id=torch.device('cuda', local))
	rank = dist.get_rank()
	tables = cases.with_mags(cases.uniform_patch(12, (503, 9000, 7000), (1.0, 0.4, 0.6), 0.07), 4, cats=(1,))
	full = nway_b200.nway_match(tables, 7.0, 0.9, logger=nway_b200.NullOutputLogger(), as_frame=False, device=local)
	got = parallel.nway_match_sharded(tables, 7.0, 0.9, gather='all', device=local, logger=nway_b200.NullOutputLogger())
	assert list(got.keys()) == list(full.keys())
	for k in full:
		assert got[k].shape == full[k].shape, (k, got[k].shape, full[k].shape)
		assert np.array_equal(got[k], full[k], equal_nan=True), k
	shard, counts, offsets = parallel.nway_match_sharded(tables, 7.0, 0.9, gather='none', device=local, logger=nway_b200.NullOutputLogger())
	lo = offsets[rank]
	for k in full:
		assert np.array_equal(shard[k], full[k][lo:lo + counts[rank]], equal_nan=True), k
	dist.barrier()
	if rank == 0:
		print('SHARDED_OK world=%d rows=%d' % (dist.get_world_size(), len(full['A'])))
	dist.destroy_process_group()


if __name__ == '__main__':
	main()
