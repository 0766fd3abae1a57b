"""Multi-GPU: one process per GPU, primary-catalogue rows sharded.

Every output row belongs to exactly one primary source and every reduction of the path (p_any, p_i,
match_flag, the CLI correction) stays inside one primary's rows, so the shards are independent: each rank
matches its contiguous block of primaries against the full secondary catalogues.  The only exchange is the
re-assembly of the table: an all-gather of the per-rank row counts (which fixes where each shard sits in the
global table) and, if the caller wants the whole table on every rank, ONE variable-length all-gather of the shards.
Two implementations: TableGather -- every GPU stores its shard straight into its final place in every rank's table
over NVLink peer memory (the library's nwb_gather_*: no padding, no staging, no unpacking; for the repeated matches of
a resident service) -- and allgather_table, the same exchange through torch.distributed (one exact-size message per
peer in one NCCL group, unpacked on arrival; gloo for the CPU tests; what the one-shot calls use).
"""
from collections import OrderedDict

import numpy


def shard_range(n_primary, rank, world):
	"""contiguous, balanced block of primary rows of this rank: (first, count)"""
	first = n_primary * rank // world
	return first, n_primary * (rank + 1) // world - first


def exchange_counts(nrows, group=None, device='cpu'):
	"""all-gather one int64 per rank; returns the list of row counts in rank order"""
	import torch
	import torch.distributed as dist
	world = dist.get_world_size(group)
	mine = torch.tensor([int(nrows)], dtype=torch.int64, device=device)
	out = torch.zeros(world, dtype=torch.int64, device=device)
	dist.all_gather_into_tensor(out, mine, group=group)
	return [int(x) for x in out.cpu().tolist()]


def row_offsets(counts):
	"""first global row of every shard"""
	return [int(x) for x in numpy.concatenate(([0], numpy.cumsum(counts)[:-1]))]


def _exchange(send, recv, group=None):
	"""ONE grouped exchange: send = [(tensor, peer)], recv = [(tensor, peer)], every message at its exact size.  With NCCL
	the whole list is a single ncclGroup (torch.distributed.batch_isend_irecv) -- no padding to the largest shard."""
	import torch.distributed as dist
	ops = [dist.P2POp(dist.irecv, t, p, group) for t, p in recv] + [dist.P2POp(dist.isend, t, p, group) for t, p in send]
	if ops:
		for w in dist.batch_isend_irecv(ops):
			w.wait()


def allgather_table(local, counts, group=None, gather='all', out=None):
	"""The reassembly of the sharded output table (SURVEY.md 8e "Collective") as one variable-length all-gather.

	local: (ncols, counts[rank]) tensor of 8-byte values (e.g. Context.table_view(): the context's own column
	allocation).  counts: rows of every rank's shard, in rank order (exchange_counts).  Returns the (ncols, sum(counts))
	table, shards in rank order, on every rank (gather='all') or on rank 0 only (gather='rank0': a real gather, the
	other ranks only send and return None).  ONE message per peer of exactly that shard's size -- the shard packed into
	a contiguous (ncols, rows) block -- all of them in one NCCL group, unpacked into the table's columns on arrival.
	Measured on 8 B200 (tools/bench_allgather.py, 8 x 590 MB): 9.1 ms, against 17.2 ms for one message per (column,
	peer) straight into place (168 messages per rank: NCCL serialises them over its point-to-point channels), 19.4 ms
	for torch's uneven all_gather per column, 8.6 ms for one ncclAllGather padded to the largest shard plus the same
	unpacking (6.4 ms of it the collective) -- and 2 (world - 1) messages instead of 2 ncols (world - 1) is what a
	small table needs anyway (15 x 1.3e5 rows: 0.36 ms against 2.2 ms)."""
	import torch
	import torch.distributed as dist
	rank, world = dist.get_rank(group), dist.get_world_size(group)
	assert len(counts) == world and local.shape[1] == counts[rank]
	ncols = local.shape[0]
	offs = row_offsets(counts)
	want = gather == 'all' or rank == 0
	if want:
		if out is None:
			out = torch.empty((ncols, sum(counts)), dtype=local.dtype, device=local.device)
		if counts[rank]:
			out[:, offs[rank]:offs[rank] + counts[rank]].copy_(local)
	send, recv = [], []
	packed = local.contiguous() if counts[rank] else None
	staging = {}
	for peer in range(world):
		if peer == rank:
			continue
		if counts[rank] and (gather == 'all' or peer == 0):
			send.append((packed.view(-1), peer))
		if want and counts[peer]:
			staging[peer] = torch.empty(ncols * counts[peer], dtype=local.dtype, device=local.device)
			recv.append((staging[peer], peer))
	_exchange(send, recv, group)
	for peer, buf in staging.items():
		out[:, offs[peer]:offs[peer] + counts[peer]].copy_(buf.view(ncols, counts[peer]))
	return out if want else None


class TableGather(object):
	"""The same reassembly over NVLink peer memory (include/nwayb200.h, nwb_gather_*): every rank's GPU stores its shard
	of the table, column by column from where the row kernels wrote it, straight into its final position in the
	gathered table of every rank.  No collective library on the data path, no padding, no staging, no unpacking: the
	bytes cross the link once and land in place.  What remains of torch.distributed is the exchange of the row counts
	(which fixes the positions) and ONE stream-ordered barrier behind the push.

	One instance per (process group, context).  setup() is collective and allocates two sets of ncols x capacity_rows
	on every rank (consecutive gathers alternate, so a rank may still read table e while its peers push e + 1); __call__
	returns this rank's complete (ncols, total) table as a torch view, valid until the next-but-one call.  The context
	works on `stream` (a torch stream; default: one of its own), and so do the count exchange and the barrier.  The row
	counts are all-gathered on the device straight into the push kernel's argument, so nothing between the match and
	the complete table waits for the host.  engine: 0 = SM stores, counts and barrier through NCCL; 1 = copy engines; 2 = SM
	stores with the counts and the barrier as flag words in peer memory -- nothing of torch.distributed per gather."""

	def __init__(self, group=None, device=None, stream=None, engine=0):
		import torch
		import torch.distributed as dist
		self.group = group
		self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
		self.device = torch.cuda.current_device() if device is None else device
		self.dev = torch.device('cuda', self.device)
		self.stream = stream
		self.engine = engine
		self.token = torch.zeros(1, dtype=torch.int32, device=self.dev)
		self.counts = torch.zeros(self.world, dtype=torch.int64, device=self.dev)
		self.ready_for = None
		self.capacity_rows = 0

	def setup(self, ctx, capacity_rows, ncols):
		import torch
		import torch.distributed as dist
		if self.stream is None:
			self.stream = torch.cuda.Stream(device=self.dev)
			ctx.set_stream(self.stream.cuda_stream)
		# a failure on one rank (no peer access, out of memory) must not leave the others waiting in a collective: every
		# step's outcome is exchanged, and all ranks raise together
		try:
			handle, err = ctx.gather_setup(self.rank, self.world, capacity_rows, ncols), None
		except Exception as e:
			handle, err = None, repr(e)
		handles = [None] * self.world
		dist.all_gather_object(handles, (handle, err), group=self.group)
		if any(h is None for h, _ in handles):
			raise RuntimeError('TableGather.setup failed: ' + '; '.join('rank %d: %s' % (r, e) for r, (h, e) in enumerate(handles) if h is None))
		try:
			ctx.gather_connect([h for h, _ in handles])
			err = None
		except Exception as e:
			err = repr(e)
		errs = [None] * self.world
		dist.all_gather_object(errs, err, group=self.group)
		if any(e is not None for e in errs):
			ctx.gather_close()
			raise RuntimeError('TableGather.setup failed: ' + '; '.join('rank %d: %s' % (r, e) for r, e in enumerate(errs) if e is not None))
		self.ready_for, self.capacity_rows = ctx, int(capacity_rows)

	def __call__(self, ctx, counts=None):
		"""one gather of the context's last table; returns (table view, counts).  counts: the shards' row counts if the
		caller has them on the host already (the copy engines need them there)."""
		import torch
		import torch.distributed as dist
		assert self.ready_for is ctx, 'TableGather.setup first'
		if self.engine == 2:
			# no collective library at all: counts and completion travel as flag words in peer memory (k_peer_sync)
			wide = ctx.gather_push(None, 2)
			counts = ctx.gather_counts(self.world)   # waits for the stream: the table is complete
			return wide[:, :sum(counts)], counts
		with torch.cuda.stream(self.stream):
			if counts is None and self.engine == 1:
				counts = exchange_counts(ctx.table_layout()[3], self.group, self.dev)
			if counts is None:
				mine = torch.tensor([ctx.table_layout()[3]], dtype=torch.int64, device=self.dev)
				dist.all_gather_into_tensor(self.counts, mine, group=self.group)
				wide = ctx.gather_push(self.counts.data_ptr(), 0)
			else:
				wide = ctx.gather_push([int(c) for c in counts], self.engine)
			dist.all_reduce(self.token, group=self.group)   # every rank's push has landed once this has run
			if counts is None:
				counts = [int(c) for c in self.counts.cpu().tolist()]   # the one wait for the host: the table is complete behind it
		total = sum(counts)
		if total > self.capacity_rows:
			raise RuntimeError('the gathered table has %d rows, TableGather.setup reserved %d' % (total, self.capacity_rows))
		return wide[:, :total], counts

	def close(self, ctx):
		ctx.gather_close()
		self.ready_for = None


def allgather_columns(cols, counts, group=None):
	"""cols: mapping name -> 1-D tensor (this rank's rows; same device, any dtypes).  Returns name -> tensor of
	sum(counts) rows, shards in rank order, identical on every rank: the same single grouped exchange as
	allgather_table, for callers that hold separate columns."""
	import torch
	import torch.distributed as dist
	rank, world = dist.get_rank(group), dist.get_world_size(group)
	assert len(counts) == world
	offs = row_offsets(counts)
	total = sum(counts)
	out = OrderedDict((name, torch.empty(total, dtype=col.dtype, device=col.device)) for name, col in cols.items())
	send, recv = [], []
	for name, col in cols.items():
		assert col.shape[0] == counts[rank]
		if counts[rank]:
			out[name][offs[rank]:offs[rank] + counts[rank]].copy_(col)
	for peer in range(world):
		if peer == rank:
			continue
		for name, col in cols.items():
			if counts[rank]:
				send.append((col.contiguous(), peer))
			if counts[peer]:
				recv.append((out[name][offs[peer]:offs[peer] + counts[peer]], peer))
	_exchange(send, recv, group)
	return out


def _table_device(ctx, device):
	"""the torch device a context's table lives on: CUDA device `device`, unless the context names its own (the CPU tests
	run this module's host logic over gloo with a context whose tables are host tensors)"""
	import torch
	return torch.device(getattr(ctx, 'torch_device', None) or 'cuda:%d' % device)


def _wait_for_current_stream(dev):
	import torch
	if dev.type == 'cuda':
		torch.cuda.current_stream(dev).synchronize()


def nway_match_sharded(match_tables, match_radius, prior_completeness, gather='all', group=None, device=None, **kwargs):
	"""nway_match() across the ranks of a torch.distributed process group (one rank per GPU).

	Every rank passes the same catalogues; rank r matches primaries shard_range(N0, r, world).  gather:
	  'all'   every rank returns the complete table (one unpadded all-gather-v straight from the context's columns),
	  'rank0' ranks send their shard to rank 0 (returns the table there, None elsewhere),
	  'none'  every rank returns only its own shard plus (counts, offsets) to place it.
	Automatic magnitude histograms (maghists=None) are selected from the rows of ALL shards: between the two passes
	every rank gathers three columns of the whole table (hist_rows below) and derives the same histogram from them, as
	the single-device match does from its own rows (nwaylib/__init__.py:324-375)."""
	import torch
	import torch.distributed as dist
	from . import nway_match, _lib, _column_names
	rank, world = dist.get_rank(group), dist.get_world_size(group)
	if device is None:
		device = torch.cuda.current_device()
	first, count = shard_range(len(match_tables[0]['ra']), rank, world)
	kwargs['as_frame'] = False
	kwargs['keep_on_device'] = True
	ncat = len(match_tables)
	held = []

	def hist_rows(ctx, c):
		"""the columns the automatic histogram of catalogue c is selected from: index of c, Separation_max and dist_post
		of ALL shards in global row order (one small all-gather-v) -- the selection is a property of the whole table:
		first occurrence of a source over all rows, the reference's weight indexing (nwaylib/__init__.py:324-366)"""
		base, stride, ncols, nrows = ctx.table_layout()
		dev = _table_device(ctx, device)
		counts = exchange_counts(nrows, group, dev)
		npairs = ncat * (ncat - 1) // 2
		pick = [c, ncat + npairs, ncat + npairs + 4]   # <name_c>, Separation_max, dist_post in the table's column order
		view = ctx.table_view()
		local3 = torch.stack([view[k] for k in pick]) if nrows else torch.empty((3, 0), dtype=torch.int64, device=dev)
		_wait_for_current_stream(dev)
		ctx.sync()
		full = allgather_table(local3, counts, group)
		_wait_for_current_stream(dev)
		held.append(full)   # stays alive while the library reads it
		return (full.shape[1], full[0].data_ptr(), full[1].data_ptr(), full[2].data_ptr())

	auto = any(h is None for t in match_tables for h in t.get('maghists', []))
	local = nway_match(match_tables, match_radius, prior_completeness, primary_range=(first, count), device=device,
		allow_empty=True, hist_rows=hist_rows if auto else None, **kwargs)
	nrows = local['nrows']
	ctx = _lib.get_context(device)
	dev = _table_device(ctx, device)
	counts = exchange_counts(nrows, group, dev)
	offsets = row_offsets(counts)
	names, seps, biases = _column_names(match_tables)
	int_cols = set(names) | {'ncat', 'match_flag'}
	colnames = list(local['selectors'].keys())
	shard = ctx.table_view() if nrows else torch.empty((len(colnames), 0), dtype=torch.int64, device=dev)

	def to_host(table):
		host = table.cpu().numpy()
		return OrderedDict((name, host[k] if name in int_cols else host[k].view(numpy.float64)) for k, name in enumerate(colnames))

	if gather == 'none':
		return to_host(shard), counts, offsets
	_wait_for_current_stream(dev)   # the context works on its own stream; it has been synchronised by nway_match
	full = allgather_table(shard, counts, group, gather=gather)
	return None if full is None else to_host(full)


class ScatterMatcher(object):
	"""The match of nway_b200.nway_match() in shard mode (include/nwayb200.h, nwb_shard_*): every rank holds all
	catalogues but STREAMS only its slice of every secondary catalogue against all primaries; a match is written by the
	streaming kernel straight into the pair store of the rank that owns the primary (blocks of ceil(n0 / world)
	primaries) over NVLink peer memory; each rank then produces the rows of its own primaries.  The time of the stream --
	the dominant cost when secondaries outnumber primaries by orders of magnitude -- divides by the number of GPUs.

	One instance per (process group, context); __call__(ctx, fuse_final) runs one match: phase 1 (grid + streaming,
	matches arrive from all ranks) | barrier | phase 2 (rows).  The barrier is stream-ordered, no host synchronisation,
	and one is enough because the exchange buffers are double-buffered (include/nwayb200.h): by default flag words in
	peer memory (peer_barrier=True: nwb_shard_match phase 3 enqueues phase 1, the flag exchange and phase 2 itself --
	nothing of torch.distributed per match), else an NCCL all-reduce of one word on the context's stream."""

	def __init__(self, group=None, device=None, spill_capacity=65536, peer_barrier=True):
		import torch
		import torch.distributed as dist
		self.group = group
		self.peer_barrier = peer_barrier   # the barrier as flags in peer memory (nwb_shard_match phase 3) instead of an NCCL all-reduce
		self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
		self.device = torch.cuda.current_device() if device is None else device
		self.dev = torch.device('cuda', self.device)
		self.stream = torch.cuda.Stream(device=self.dev)
		self.token = torch.zeros(1, dtype=torch.int32, device=self.dev)
		self.flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
		self.spill_capacity = spill_capacity
		self.ready_for = None

	def setup(self, ctx):
		"""exchange buffers + peer mapping for the catalogues / radius now set on the context (collective)"""
		import torch.distributed as dist
		# as in TableGather.setup: a failure on one rank (no peer access, out of memory) is exchanged, and all ranks
		# raise together instead of leaving the others waiting in a collective
		try:
			(handle, nbytes), err = ctx.shard_setup(self.rank, self.world, self.spill_capacity), None
		except Exception as e:
			(handle, nbytes), err = (None, 0), repr(e)
		handles = [None] * self.world
		dist.all_gather_object(handles, (handle, err), group=self.group)
		if any(h is None for h, _ in handles):
			raise RuntimeError('ScatterMatcher.setup failed: ' + '; '.join('rank %d: %s' % (r, e) for r, (h, e) in enumerate(handles) if h is None))
		try:
			ctx.shard_connect([h for h, _ in handles])
			err = None
		except Exception as e:
			err = repr(e)
		errs = [None] * self.world
		dist.all_gather_object(errs, err, group=self.group)
		if any(e is not None for e in errs):
			ctx.shard_close()
			raise RuntimeError('ScatterMatcher.setup failed: ' + '; '.join('rank %d: %s' % (r, e) for r, e in enumerate(errs) if e is not None))
		ctx.set_stream(self.stream.cuda_stream)
		self.ready_for = ctx
		return nbytes

	def barrier(self):
		import torch.distributed as dist
		dist.all_reduce(self.token, group=self.group)

	def __call__(self, ctx, fuse_final=True):
		import torch
		import torch.distributed as dist
		if self.ready_for is not ctx:
			self.setup(ctx)
		with torch.cuda.stream(self.stream):
			for attempt in range(4):
				if self.peer_barrier:
					nrows, retry = ctx.shard_match(3, fuse_final)
				else:
					ctx.shard_match(1)
					self.barrier()
					nrows, retry = ctx.shard_match(2, fuse_final)
				# "a grid buffer was too small": a property of the (replicated) primaries, so every rank takes the same
				# decision without asking the others
				if not retry:
					return nrows
		raise RuntimeError('shard mode: the grid buffers kept overflowing')

	def close(self, ctx):
		ctx.set_stream(None)
		ctx.shard_close()
		self.ready_for = None


def nway_match_scatter(match_tables, match_radius, prior_completeness, gather='all', group=None, device=None, peer_barrier=True, **kwargs):
	"""nway_match() with the streaming of the secondaries shared between the ranks (strong scaling; see ScatterMatcher).
	Every rank passes the same catalogues.  gather as in nway_match_sharded: 'all' / 'rank0' / 'none'.  Automatic
	magnitude histograms are not available in this mode (supply maghists)."""
	import torch
	import torch.distributed as dist
	from . import nway_match, _lib, _column_names
	for t in match_tables:
		if any(h is None for h in t.get('maghists', [])):
			raise NotImplementedError('automatic magnitude histograms are a global step; supply maghists in sharded mode')
	rank, world = dist.get_rank(group), dist.get_world_size(group)
	if device is None:
		device = torch.cuda.current_device()
	dev = torch.device('cuda', device)
	ctx = _lib.get_context(device)
	matcher = ScatterMatcher(group, device, peer_barrier=peer_barrier)
	kwargs['as_frame'] = False
	kwargs['keep_on_device'] = True
	try:
		local = nway_match(match_tables, match_radius, prior_completeness, device=device, allow_empty=True, matcher=matcher, **kwargs)
		nrows = local['nrows']
		with torch.cuda.stream(matcher.stream):
			counts = exchange_counts(nrows, group, dev)
			offsets = row_offsets(counts)
			names, seps, biases = _column_names(match_tables)
			int_cols = set(names) | {'ncat', 'match_flag'}
			colnames = list(local['selectors'].keys())
			shard = ctx.table_view() if nrows else torch.empty((len(colnames), 0), dtype=torch.int64, device=dev)

			def to_host(table):
				host = table.cpu().numpy()
				return OrderedDict((name, host[k] if name in int_cols else host[k].view(numpy.float64)) for k, name in enumerate(colnames))

			if gather == 'none':
				return to_host(shard), counts, offsets
			full = allgather_table(shard, counts, group, gather=gather)
			return None if full is None else to_host(full)
	finally:
		matcher.close(ctx)
