"""CPU, world_size 2, gloo: the host-side logic of the sharded path (shard bounds, count exchange, the unpadded
all-gather-v of ragged shards as one grouped exchange, gather to rank 0, re-assembly order).  The per-shard tables come from the oracle, so this runs without
a GPU; the same code path runs over NCCL on device tensors in tests/test_gpu_parity.py / bench.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nway_b200 import parallel
from tests import cases


def free_port():
	s = socket.socket()
	s.bind(('127.0.0.1', 0))
	port = s.getsockname()[1]
	s.close()
	return port


def worker(rank, world, port, ret):
	os.environ['MASTER_ADDR'] = '127.0.0.1'
	os.environ['MASTER_PORT'] = str(port)
	dist.init_process_group('gloo', rank=rank, world_size=world)
	try:
		from oracle import nway_oracle as O
		tables = cases.uniform_patch(9, (301, 4000, 3000), (1.0, 0.4, 0.6), 0.05)
		full = O.nway_match(tables, 6.0, 0.9)
		n0 = len(tables[0]['ra'])
		first, count = parallel.shard_range(n0, rank, world)
		mine = (full['A'] >= first) & (full['A'] < first + count)
		cols = {k: torch.from_numpy(np.ascontiguousarray(v[mine])) for k, v in full.items() if not k.startswith('_')}
		counts = parallel.exchange_counts(int(mine.sum()), None, 'cpu')
		assert sum(counts) == len(full['A'])
		assert parallel.row_offsets(counts)[rank] == int((full['A'] < first).sum())
		got = parallel.allgather_columns(cols, counts)
		for k, v in got.items():
			a, b = v.numpy(), full[k]
			assert a.shape == b.shape, k
			assert np.array_equal(a, b, equal_nan=True), k
		# the table form: (ncols, rows) of 8-byte words sent from a strided view (what Context.table_view() is), to
		# every rank and to rank 0 only
		names = list(cols)
		stride = int(mine.sum()) + 37
		store = torch.zeros((len(names), stride), dtype=torch.int64)
		for k, name in enumerate(names):
			store[k, :int(mine.sum())] = cols[name].view(torch.int64)
		local = store[:, :int(mine.sum())]
		for mode in ('all', 'rank0'):
			tab = parallel.allgather_table(local, counts, gather=mode)
			if mode == 'rank0' and rank != 0:
				assert tab is None
				continue
			assert tab.shape == (len(names), len(full['A']))
			for k, name in enumerate(names):
				assert np.array_equal(tab[k].numpy().view(full[name].dtype), full[name], equal_nan=True), (mode, name)
		ret[rank] = 'ok'
	except Exception as e:   # surface the failure in the parent
		ret[rank] = repr(e)
	finally:
		dist.destroy_process_group()


class FakeContext(object):
	"""stands in for _lib.Context in the collective set-up handshakes: fails where it is told to"""

	def __init__(self, rank, fail_setup_on=None, fail_connect_on=None):
		self.rank, self.fail_setup_on, self.fail_connect_on = rank, fail_setup_on, fail_connect_on
		self.closed = 0
		self.stream = 'unset'

	def shard_setup(self, rank, world, spill_capacity):
		if rank == self.fail_setup_on:
			raise RuntimeError('no memory on rank %d' % rank)
		return bytes([rank]) * 64, 4096

	def gather_setup(self, rank, world, capacity_rows, ncols):
		return self.shard_setup(rank, world, 0)[0]

	def shard_connect(self, handles):
		assert [h[0] for h in handles] == list(range(len(handles))) and all(len(h) == 64 for h in handles)
		if self.rank == self.fail_connect_on:
			raise RuntimeError('no peer access on rank %d' % self.rank)

	gather_connect = shard_connect

	def shard_close(self):
		self.closed += 1

	gather_close = shard_close

	def set_stream(self, s):
		self.stream = s


class FakeStream(object):
	cuda_stream = 1234


def handshake_worker(rank, world, port, ret):
	os.environ['MASTER_ADDR'] = '127.0.0.1'
	os.environ['MASTER_PORT'] = str(port)
	dist.init_process_group('gloo', rank=rank, world_size=world)
	try:
		def bare(cls):
			m = object.__new__(cls)   # the constructors allocate on a CUDA device
			m.group, m.rank, m.world, m.stream, m.ready_for = None, rank, world, FakeStream(), None
			m.spill_capacity = 65536
			return m
		for cls, setup in ((parallel.ScatterMatcher, lambda m, c: m.setup(c)), (parallel.TableGather, lambda m, c: m.setup(c, 1000, 12))):
			# a failure on ONE rank, in either step: EVERY rank raises and names it; nobody is left in a collective
			for kw, culprit, word in ((dict(fail_setup_on=1), 1, 'no memory'), (dict(fail_connect_on=0), 0, 'no peer access')):
				m, ctx = bare(cls), FakeContext(rank, **kw)
				with pytest.raises(RuntimeError) as info:
					setup(m, ctx)
				assert 'rank %d' % culprit in str(info.value) and word in str(info.value), str(info.value)
				assert m.ready_for is None
				assert ctx.closed == (1 if 'fail_connect_on' in kw else 0)
			m, ctx = bare(cls), FakeContext(rank)
			setup(m, ctx)
			assert m.ready_for is ctx and ctx.closed == 0
			if cls is parallel.ScatterMatcher:
				assert ctx.stream == FakeStream.cuda_stream
		ret[rank] = 'ok'
	except BaseException as e:
		ret[rank] = repr(e)
	finally:
		dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_setup_failure_on_one_rank_raises_on_every_rank():
	world = 2
	port = free_port()
	with mp.Manager() as mgr:
		ret = mgr.dict()
		mp.spawn(handshake_worker, args=(world, port, ret), nprocs=world, join=True)
		assert dict(ret) == {0: 'ok', 1: 'ok'}, dict(ret)


def sharded_worker(rank, world, port, ret):
	os.environ['MASTER_ADDR'] = '127.0.0.1'
	os.environ['MASTER_PORT'] = str(port)
	dist.init_process_group('gloo', rank=rank, world_size=world)
	try:
		import nway_b200
		from tests import oraclectx, parity
		ctx = oraclectx.OracleContext()
		nway_b200._lib.get_context = lambda device=None: ctx   # this process only: the numeric stages on the CPU
		quiet = dict(logger=nway_b200.NullOutputLogger(), store_mag_hists=False, device=0)
		# automatic histograms (by posterior, two magnitude columns) are a property of the WHOLE table: every rank selects from
		# three gathered columns of all shards and must end with the table of the single-device match -- the reference's
		for name in ('cosmos3_magauto', 'cosmos2_magradius'):
			spec = cases.GOLDEN_CASES[name]
			got = parallel.nway_match_sharded(cases.build_case(name), spec['radius'], spec['completeness'], gather='all', **dict(quiet, **spec.get('kwargs', {})))
			parity.check_against_digest(name, got, [t['name'] for t in cases.build_case(name)])
			assert ctx.primary_range == parallel.shard_range(len(cases.build_case(name)[0]['ra']), rank, world)
		# the three ways to hand the table back, a completeness vector, truncation: against the whole match of one context
		spec = cases.GOLDEN_CASES['syn4_minprob']
		args = (spec['radius'], spec['completeness'])
		kw = dict(quiet, **spec['kwargs'])
		whole = parity.load_golden('ref_syn4_minprob.npz')
		everywhere = parallel.nway_match_sharded(cases.build_case('syn4_minprob'), *args, gather='all', **kw)
		parity.check_against_digest('syn4_minprob', everywhere, [t['name'] for t in cases.build_case('syn4_minprob')])
		at_root = parallel.nway_match_sharded(cases.build_case('syn4_minprob'), *args, gather='rank0', **kw)
		assert (at_root is None) == (rank != 0)
		if rank == 0:
			for c in everywhere:
				assert np.array_equal(at_root[c], everywhere[c], equal_nan=True), c
		mine, counts, offsets = parallel.nway_match_sharded(cases.build_case('syn4_minprob'), *args, gather='none', **kw)
		assert sum(counts) == int(whole['nrows']) == len(everywhere['A']) and offsets[0] == 0 and offsets[1] == counts[0]
		for c in everywhere:
			assert np.array_equal(mine[c], everywhere[c][offsets[rank]:offsets[rank] + counts[rank]], equal_nan=True), c
		ret[rank] = 'ok'
	except BaseException as e:
		import traceback
		ret[rank] = traceback.format_exc()[-1500:]
	finally:
		dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_world2_sharded_match_with_automatic_histograms_equals_the_reference():
	"""nway_b200.parallel.nway_match_sharded itself over gloo, two ranks, the oracle standing in for the library's numeric
	stages in each (tests/oraclectx.py): shard ranges, the count exchange, the gathered columns the automatic histograms
	are selected from, the all-gather-v / gather / no gather of the table -- equal to the committed outputs of the
	unmodified reference.  The same function runs over NCCL on B200 in tests/run_sharded.py."""
	world = 2
	port = free_port()
	with mp.Manager() as mgr:
		ret = mgr.dict()
		mp.spawn(sharded_worker, args=(world, port, ret), nprocs=world, join=True)
		assert dict(ret) == {0: 'ok', 1: 'ok'}, '\n'.join('rank %d: %s' % kv for kv in dict(ret).items())


def test_shard_ranges_cover_the_primaries():
	for n in (0, 1, 7, 100000, 1000003):
		for world in (1, 2, 3, 8):
			spans = [parallel.shard_range(n, r, world) for r in range(world)]
			assert spans[0][0] == 0 and sum(c for _, c in spans) == n
			for (f0, c0), (f1, _) in zip(spans, spans[1:]):
				assert f0 + c0 == f1
			assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


@pytest.mark.timeout(300)
def test_world2_gloo_allgather_reassembles_the_table():
	world = 2
	port = free_port()
	with mp.Manager() as mgr:
		ret = mgr.dict()
		mp.spawn(worker, args=(world, port, ret), nprocs=world, join=True)
		assert dict(ret) == {0: 'ok', 1: 'ok'}, dict(ret)
