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
	# the reassembly over NVLink peer memory (nwb_gather_*, parallel.TableGather): SM stores and copy engines, two gathers
	# each (both buffer sets), against the single-device table
	from nway_b200 import _lib
	n0 = len(tables[0]['ra'])
	first, count = parallel.shard_range(n0, rank, dist.get_world_size())
	mine = nway_b200.nway_match(tables, 7.0, 0.9, logger=nway_b200.NullOutputLogger(), as_frame=False, device=local,
		primary_range=(first, count), allow_empty=True, keep_on_device=True)
	ctx = _lib.get_context(local)
	names = list(mine['selectors'].keys())
	assert names == list(full.keys())
	for engine in (0, 1, 2):   # 2: counts and barrier as flag words in peer memory, nothing of NCCL per gather
		tg = parallel.TableGather(None, local, engine=engine)
		tg.setup(ctx, len(full['A']) + 8, len(names))
		for again in range(3):
			table, cnts = tg(ctx, None if again else parallel.exchange_counts(mine['nrows'], None, torch.device('cuda', local)))
			tg.stream.synchronize()
			host = table.cpu().numpy()
			assert host.shape == (len(names), len(full['A'])), (host.shape, len(full['A']))
			for k, name in enumerate(names):
				assert np.array_equal(host[k].view(full[name].dtype), full[name], equal_nan=True), ('peer gather', engine, again, name)
		tg.close(ctx)
		ctx.set_stream(None)
	# automatic magnitude histograms in sharded mode: selected from the gathered rows of all shards -- the same table as
	# on one device, by radius and by posterior (the reference's two modes, nwaylib/__init__.py:324-375)
	synthetic = lambda: cases.with_mags(cases.uniform_patch(14, (900, 30000, 20000), (1.0, 0.4, 0.6), 0.12), 5, cats=(1, 2), hist=False)
	cosmos = lambda: cases.cosmos_subset(3, mags=True)   # BASELINE.json configs[1]: both priors automatic, selected by posterior
	for auto, radius, kw in ((synthetic, 7.0, dict(mag_include_radius=3.0)), (cosmos, 20.0, dict()), (cosmos, 20.0, dict(mag_include_radius=4.0))):
		one = nway_b200.nway_match(auto(), radius, 0.9, logger=nway_b200.NullOutputLogger(), as_frame=False, device=local, store_mag_hists=False, **kw)
		two = parallel.nway_match_sharded(auto(), radius, 0.9, gather='all', device=local, logger=nway_b200.NullOutputLogger(), store_mag_hists=False, **kw)
		assert list(one.keys()) == list(two.keys())
		for k in one:
			assert np.array_equal(two[k], one[k], equal_nan=True), ('auto histograms', sorted(kw), k)
	# the other multi-GPU mode: the ranks share the streaming of the secondaries, matches travel to the owner of the
	# primary over peer memory (nwb_shard_*); same table, bit for bit
	got2 = parallel.nway_match_scatter(tables, 7.0, 0.9, gather='all', device=local, logger=nway_b200.NullOutputLogger())
	assert list(got2.keys()) == list(full.keys())
	for k in full:
		assert got2[k].shape == full[k].shape, ('scatter', k, got2[k].shape, full[k].shape)
		assert np.array_equal(got2[k], full[k], equal_nan=True), ('scatter', k)
	# ... with the NCCL all-reduce as the barrier instead of the flags in peer memory
	got3 = parallel.nway_match_scatter(tables, 7.0, 0.9, gather='all', device=local, logger=nway_b200.NullOutputLogger(), peer_barrier=False)
	for k in full:
		assert np.array_equal(got3[k], full[k], equal_nan=True), ('scatter, NCCL barrier', k)
	# ... also for two catalogues (the specialised row kernel) and on an off-equator flat field (NWB_COMPAT_FLAT_HASH)
	for name in ('syn2', 'offeq3'):
		spec = cases.GOLDEN_CASES[name]
		one = nway_b200.nway_match(cases.build_case(name), spec['radius'], spec['completeness'], logger=nway_b200.NullOutputLogger(), as_frame=False, device=local)
		two = parallel.nway_match_scatter(cases.build_case(name), spec['radius'], spec['completeness'], gather='all', device=local, logger=nway_b200.NullOutputLogger())
		for k in one:
			assert np.array_equal(two[k], one[k], equal_nan=True), ('scatter', name, k)
	dist.barrier()
	if rank == 0:
		print('SHARDED_OK world=%d rows=%d' % (dist.get_world_size(), len(full['A'])))
	dist.destroy_process_group()


if __name__ == '__main__':
	main()
