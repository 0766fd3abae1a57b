"""Multi-GPU: one process per GPU, primary-catalogue rows sharded.

Every output row belongs to exactly one primary source and every reduction of the path (p_any, p_i,
match_flag, the CLI correction) stays inside one primary's rows, so the shards are independent: each rank
matches its contiguous block of primaries against the full secondary catalogues.  The only exchange is the
re-assembly of the table: an all-gather of the per-rank row counts (which fixes where each shard sits in the
global table) and, if the caller wants the whole table on every rank, one padded all-gather per column --
NCCL over NVLink for device tensors, gloo for the CPU tests.
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


def allgather_columns(cols, counts, group=None):
	"""cols: mapping name -> 1-D torch tensor (all of this rank's row count, 8-byte dtypes, same device).
	Returns name -> tensor of sum(counts) rows, shards in rank order, identical on every rank.  NCCL needs equal
	message sizes, so each column is padded to the longest shard and trimmed after the collective."""
	import torch
	import torch.distributed as dist
	world = dist.get_world_size(group)
	assert len(counts) == world
	longest = max(max(counts), 1)
	out = OrderedDict()
	for name, col in cols.items():
		n = col.shape[0]
		send = col
		if n != longest:
			send = torch.zeros(longest, dtype=col.dtype, device=col.device)
			send[:n] = col
		recv = torch.empty(world * longest, dtype=col.dtype, device=col.device)
		dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
		if all(c == longest for c in counts):
			out[name] = recv
		else:
			out[name] = torch.cat([recv[r * longest:r * longest + counts[r]] for r in range(world)])
	return out


def nway_match_sharded(match_tables, match_radius, prior_completeness, gather='all', group=None, device=None, **kwargs):
	"""nway_match() across the ranks of a torch.distributed process group (one rank per GPU).

	Every rank passes the same catalogues; rank r matches primaries shard_range(N0, r, world).  gather:
	  'all'   every rank returns the complete table (padded NCCL all-gather of every column),
	  'rank0' ranks send their shard to rank 0 (returns the table there, None elsewhere),
	  'none'  every rank returns only its own shard plus (counts, offsets) to place it.
	Automatic magnitude histograms (maghists=None) need the global first pass and are not supported here:
	supply the histograms."""
	import torch
	import torch.distributed as dist
	from . import nway_match, _lib, _column_names
	for t in match_tables:
		if any(h is None for h in t.get('maghists', [])):
			raise NotImplementedError('automatic magnitude histograms are a global step; supply maghists in sharded mode')
	rank, world = dist.get_rank(group), dist.get_world_size(group)
	if device is None:
		device = torch.cuda.current_device()
	first, count = shard_range(len(match_tables[0]['ra']), rank, world)
	kwargs['as_frame'] = False
	kwargs['keep_on_device'] = True
	local = nway_match(match_tables, match_radius, prior_completeness, primary_range=(first, count), device=device,
		allow_empty=True, **kwargs)
	nrows = local['nrows']
	dev = torch.device('cuda', device)
	counts = exchange_counts(nrows, group, dev)
	offsets = row_offsets(counts)
	ctx = _lib.get_context(device)
	names, seps, biases = _column_names(match_tables)
	int_cols = set(names) | {'ncat', 'match_flag'}
	cols = OrderedDict()
	for name, sel in local['selectors'].items():
		tns = torch.empty(nrows, dtype=torch.int64 if name in int_cols else torch.float64, device=dev)
		if nrows:
			ctx.fetch_device(sel, tns.data_ptr())
		cols[name] = tns
	ctx.sync()
	if gather == 'none':
		return OrderedDict((k, v.cpu().numpy()) for k, v in cols.items()), counts, offsets
	full = allgather_columns(cols, counts, group)
	if gather == 'rank0' and rank != 0:
		return None
	return OrderedDict((k, v.cpu().numpy()) for k, v in full.items())
