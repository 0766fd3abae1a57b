"""Multi-GPU: one process per GPU, primary-catalogue rows sharded.

Every output row belongs to exactly one primary source and every reduction of the path (p_any, p_i,
match_flag, the CLI correction) stays inside one primary's rows, so the shards are independent: each rank
matches its contiguous block of primaries against the full secondary catalogues.  The only exchange is the
re-assembly of the table: an all-gather of the per-rank row counts (which fixes where each shard sits in the
global table) and, if the caller wants the whole table on every rank, ONE variable-length all-gather of the shards:
every (column, peer) message at its exact size, sent from the context's own column allocation into its final place
in the gathered table, all in one NCCL group over NVLink (gloo for the CPU tests).
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


PACK_BELOW_BYTES = 1 << 20   # per (column, peer) message: below this the exchange is latency-bound and shards travel packed


def _exchange(send, recv, group=None):
	"""ONE grouped exchange: send = [(tensor, peer)], recv = [(tensor, peer)], every message at its exact size.  With NCCL
	the whole list is a single ncclGroup (torch.distributed.batch_isend_irecv) -- no padding, no staging copies, no
	concatenation afterwards: the receives land in their final place."""
	import torch.distributed as dist
	ops = [dist.P2POp(dist.irecv, t, p, group) for t, p in recv] + [dist.P2POp(dist.isend, t, p, group) for t, p in send]
	if ops:
		for w in dist.batch_isend_irecv(ops):
			w.wait()


def allgather_table(local, counts, group=None, gather='all', out=None):
	"""The reassembly of the sharded output table (SURVEY.md 8e "Collective") as one variable-length all-gather.

	local: (ncols, counts[rank]) tensor of 8-byte values whose rows are contiguous (e.g. Context.table_view(): the
	context's own column allocation, sent from where the row kernels wrote it).  counts: rows of every rank's shard, in
	rank order (exchange_counts).  Returns the (ncols, sum(counts)) table, shards in rank order, on every rank
	(gather='all') or on rank 0 only (gather='rank0': a real gather, the other ranks only send and return None).
	Per peer and column one message of exactly that shard's size, all of them in one NCCL group; the own shard is
	placed with one strided device copy.  Small shards (column pieces below PACK_BELOW_BYTES) travel as one packed
	message per peer instead: there the number of messages, not their bytes, is the cost."""
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
	if max(counts) * local.element_size() >= PACK_BELOW_BYTES:
		# bandwidth regime: every (column, peer) message goes from the context's column to its final place, no staging
		for peer in range(world):
			if peer == rank:
				continue
			if counts[rank] and (gather == 'all' or peer == 0):
				send += [(local[k], peer) for k in range(ncols)]
			if want and counts[peer]:
				recv += [(out[k, offs[peer]:offs[peer] + counts[peer]], peer) for k in range(ncols)]
		_exchange(send, recv, group)
		return out if want else None
	# latency regime (shards of a few MB: a sparse all-sky match): ONE message per peer -- the shard packed into a
	# contiguous (ncols, rows) block, unpacked into the table's columns on arrival; 2 (world - 1) messages instead of
	# 2 ncols (world - 1)
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
		dev = torch.device('cuda', device)
		counts = exchange_counts(nrows, group, dev)
		npairs = ncat * (ncat - 1) // 2
		pick = [c, ncat + npairs, ncat + npairs + 4]   # <name_c>, Separation_max, dist_post in the table's column order
		view = ctx.table_view()
		local3 = torch.stack([view[k] for k in pick]) if nrows else torch.empty((3, 0), dtype=torch.int64, device=dev)
		torch.cuda.current_stream(dev).synchronize()
		ctx.sync()
		full = allgather_table(local3, counts, group)
		torch.cuda.current_stream(dev).synchronize()
		held.append(full)   # stays alive while the library reads it
		return (full.shape[1], full[0].data_ptr(), full[1].data_ptr(), full[2].data_ptr())

	auto = any(h is None for t in match_tables for h in t.get('maghists', []))
	local = nway_match(match_tables, match_radius, prior_completeness, primary_range=(first, count), device=device,
		allow_empty=True, hist_rows=hist_rows if auto else None, **kwargs)
	nrows = local['nrows']
	dev = torch.device('cuda', device)
	counts = exchange_counts(nrows, group, dev)
	offsets = row_offsets(counts)
	ctx = _lib.get_context(device)
	names, seps, biases = _column_names(match_tables)
	int_cols = set(names) | {'ncat', 'match_flag'}
	colnames = list(local['selectors'].keys())
	shard = ctx.table_view() if nrows else torch.empty((len(colnames), 0), dtype=torch.int64, device=dev)

	def to_host(table):
		host = table.cpu().numpy()
		return OrderedDict((name, host[k] if name in int_cols else host[k].view(numpy.float64)) for k, name in enumerate(colnames))

	if gather == 'none':
		return to_host(shard), counts, offsets
	torch.cuda.current_stream(dev).synchronize()   # the context works on its own stream; it has been synchronised by nway_match
	full = allgather_table(shard, counts, group, gather=gather)
	return None if full is None else to_host(full)


class ScatterMatcher(object):
	"""The match of nway_b200.nway_match() in shard mode (include/nwayb200.h, nwb_shard_*): every rank holds all
	catalogues but STREAMS only its slice of every secondary catalogue against all primaries; a match is written by the
	streaming kernel straight into the pair store of the rank that owns the primary (blocks of ceil(n0 / world)
	primaries) over NVLink peer memory; each rank then produces the rows of its own primaries.  The time of the stream --
	the dominant cost when secondaries outnumber primaries by orders of magnitude -- divides by the number of GPUs.

	One instance per (process group, context); __call__(ctx, fuse_final) runs one match: phase 1 (grid + streaming,
	matches arrive from all ranks) | barrier | phase 2 (rows).  The barrier is an NCCL all-reduce of one word on the
	context's stream: stream-ordered, no host synchronisation; one is enough because the exchange buffers are double-
	buffered (include/nwayb200.h)."""

	def __init__(self, group=None, device=None, spill_capacity=65536):
		import torch
		import torch.distributed as dist
		self.group = group
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
		handle, nbytes = ctx.shard_setup(self.rank, self.world, self.spill_capacity)
		handles = [None] * self.world
		dist.all_gather_object(handles, handle, group=self.group)
		ctx.shard_connect(handles)
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


def nway_match_scatter(match_tables, match_radius, prior_completeness, gather='all', group=None, device=None, **kwargs):
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
	matcher = ScatterMatcher(group, device)
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
