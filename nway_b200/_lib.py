"""ctypes binding of libnwayb200.so (include/nwayb200.h).  There is no CPU fallback: if the library is
missing or no CUDA device is visible, every entry point raises."""
import ctypes
import os

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('NWB_LIB') or os.path.join(_HERE, 'libnwayb200.so')   # $NWB_LIB: experiment builds

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int64_p = ctypes.POINTER(ctypes.c_int64)

# column selectors (nwayb200.h)
COL_IDX, COL_SEP, COL_BIAS = 0, 100, 200
COL_SEPMAX, COL_NCAT, COL_LOGBF_UNCORR, COL_LOGBF, COL_DIST_POST, COL_P_SINGLE, COL_MATCH_FLAG, COL_P_ANY, COL_P_I = range(300, 309)
ERR_CIRCULAR, ERR_ELLIPSE = 1, 3
UNRELATED_API, UNRELATED_CLI = 0, 1
MAX_CATALOGUES, MAX_MAG_COLUMNS, MAX_HIST_BINS = 8, 8, 64   # nwb::MAXC, MAXM, MAXB (nwb_device.cuh)
COMPAT_SEP_F32 = 1
COMPAT_FLAT_HASH = 2
T_GRID, T_PAIRS, T_LISTS, T_ROWS, T_FINAL, T_TOTAL, T_KPAIRS, T_KROWS = range(8)
STAGE_NAMES = ['grid', 'pairs', 'lists', 'rows', 'final', 'total', 'k_pairs', 'k_rows']
NWB_ERR_EMPTY = -3

EXPORTS = {
	# name: (restype, argtypes)
	'nwb_version': (ctypes.c_int, []),
	'nwb_create': (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
	'nwb_destroy': (None, [ctypes.c_void_p]),
	'nwb_last_error': (ctypes.c_char_p, [ctypes.c_void_p]),
	'nwb_set_stream': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
	'nwb_set_catalogue': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p,
		ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int]),
	'nwb_set_params': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, c_double_p, ctypes.c_double, ctypes.c_int]),
	'nwb_set_compat': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
	'nwb_flat_hash_applied': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]),
	'nwb_maghist_select': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int,
		c_int64_p, c_int64_p, c_double_p]),
	'nwb_maghist_select_rows': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
		ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, c_int64_p, c_int64_p, c_double_p]),
	'nwb_maghist_sample': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, c_double_p, c_double_p]),
	'nwb_maghist_count': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, c_int64_p]),
	'nwb_set_prefilter': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), c_double_p]),
	'nwb_set_tables': (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.c_double, c_double_p, c_double_p, c_double_p]),
	'nwb_set_maghist': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, c_double_p]),
	'nwb_set_primary_range': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64]),
	'nwb_shard_setup': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, c_int64_p]),
	'nwb_shard_connect': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
	'nwb_shard_match': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_int64_p]),
	'nwb_shard_close': (ctypes.c_int, [ctypes.c_void_p]),
	'nwb_gather_setup': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]),
	'nwb_gather_connect': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
	'nwb_gather_push': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), c_int64_p]),
	'nwb_gather_counts': (ctypes.c_int, [ctypes.c_void_p, c_int64_p]),
	'nwb_gather_close': (ctypes.c_int, [ctypes.c_void_p]),
	'nwb_match': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_int64_p]),
	'nwb_match_async': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
	'nwb_match_wait': (ctypes.c_int, [ctypes.c_void_p, c_int64_p]),
	'nwb_finalize': (ctypes.c_int, [ctypes.c_void_p]),
	'nwb_truncate': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, c_int64_p]),
	'nwb_fetch': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
	'nwb_nrows_device_ptr': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
	'nwb_fetch_device': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
	'nwb_column_ptr': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
	'nwb_table_layout': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), c_int64_p, ctypes.POINTER(ctypes.c_int), c_int64_p]),
	'nwb_sync': (ctypes.c_int, [ctypes.c_void_p]),
	'nwb_timing': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]),
	'nwb_launch_count': (ctypes.c_int, [ctypes.c_void_p, c_int64_p]),
	'nwb_bench_skeleton': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]),
	'nwb_stats': (ctypes.c_int, [ctypes.c_void_p, c_int64_p]),
	'nwb_dist': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
	'nwb_log_bf': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, c_double_p, c_double_p, c_double_p]),
	'nwb_score_rows': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, c_int64_p, c_double_p, c_double_p, c_int64_p, c_double_p, c_double_p]),
	'nwb_posterior': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, c_double_p, c_double_p, c_double_p]),
	'nwb_log_bf_elliptical': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
	'nwb_row_offsets': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p]),
}

_lib = None


class NwbError(RuntimeError):
	def __init__(self, code, msg):
		RuntimeError.__init__(self, 'libnwayb200 error %d: %s' % (code, msg))
		self.code = code


def load():
	"""dlopen the in-tree library and declare every prototype.  Raises if it has not been built."""
	global _lib
	if _lib is not None:
		return _lib
	if not os.path.exists(LIB_PATH):
		raise ImportError('%s not found: build it with `python -m nway_b200.build` (needs nvcc). '
			'There is no CPU fallback.' % LIB_PATH)
	lib = ctypes.CDLL(LIB_PATH)
	for name, (restype, argtypes) in EXPORTS.items():
		fn = getattr(lib, name)
		fn.restype = restype
		fn.argtypes = argtypes
	_lib = lib
	return lib


def dptr(a):
	return a.ctypes.data_as(c_double_p)


def f64(a):
	return numpy.ascontiguousarray(a, dtype=numpy.float64)


class Context(object):
	"""one libnwayb200 context == one CUDA device + stream + scratch"""

	def __init__(self, device=0):
		self.lib = load()
		h = ctypes.c_void_p()
		rc = self.lib.nwb_create(int(device), ctypes.byref(h))
		if rc != 0:
			raise NwbError(rc, (self.lib.nwb_last_error(None) or b'').decode())
		self.h = h
		self.device = int(device)
		self._keep = []

	def close(self):
		if getattr(self, 'h', None):
			self.lib.nwb_destroy(self.h)
			self.h = None

	def __del__(self):
		try:
			self.close()
		except Exception:
			pass

	def check(self, rc):
		if rc != 0:
			raise NwbError(rc, (self.lib.nwb_last_error(self.h) or b'').decode())

	def set_catalogue(self, c, ncat, ra, dec, err, area, mags=(), err_kind=ERR_CIRCULAR):
		"""host numpy arrays (copied to the device)"""
		ra, dec, err = f64(ra), f64(dec), f64(err)
		n = len(ra)
		assert len(dec) == n and err.size == n * err_kind
		m = len(mags)
		magbuf = numpy.ascontiguousarray(numpy.stack([f64(x) for x in mags])) if m else None
		self.check(self.lib.nwb_set_catalogue(self.h, c, ncat, n, ra.ctypes.data, dec.ctypes.data, err.ctypes.data,
			err_kind, magbuf.ctypes.data if m else None, m, float(area), 0))
		self.sync()   # the host arrays may be temporaries

	def set_catalogue_device(self, c, ncat, n, ra_ptr, dec_ptr, err_ptr, area, mags_ptr=None, m=0, err_kind=ERR_CIRCULAR):
		"""device pointers (e.g. torch tensors' data_ptr()); used in place"""
		self.check(self.lib.nwb_set_catalogue(self.h, c, ncat, int(n), ra_ptr, dec_ptr, err_ptr, err_kind,
			mags_ptr, m, float(area), 1))

	def set_stream(self, cuda_stream):
		self.check(self.lib.nwb_set_stream(self.h, cuda_stream))

	def set_params(self, radius, completeness, ratio_secondary=0.5, unrelated_mode=UNRELATED_API):
		pc = f64(completeness)
		self.check(self.lib.nwb_set_params(self.h, float(radius), dptr(pc), float(ratio_secondary), int(unrelated_mode)))

	def row_offsets(self, a, b, nrows):
		"""(dra, ddec) in arcsec between the members a < b of every row of the last result (NaN where absent)"""
		dra, ddec = numpy.empty(nrows), numpy.empty(nrows)
		if nrows:
			self.check(self.lib.nwb_row_offsets(self.h, int(a), int(b), dptr(dra), dptr(ddec)))
		return dra, ddec

	def set_prefilter(self, pairwise_errs):
		"""[(catalogue a, catalogue b, radius in arcsec), ...]; empty clears"""
		n = len(pairwise_errs)
		a = (ctypes.c_int * max(n, 1))(*[int(x[0]) for x in pairwise_errs])
		b = (ctypes.c_int * max(n, 1))(*[int(x[1]) for x in pairwise_errs])
		r = (ctypes.c_double * max(n, 1))(*[float(x[2]) for x in pairwise_errs])
		self.check(self.lib.nwb_set_prefilter(self.h, n, a, b, r))

	def maghist_select(self, c, k, by_radius, thr_select, thr_possible, weights_cli, rows=None):
		"""device half of the automatic histogram: returns (magnitudes, weights) of the unique selected sources in
		ascending source order, (n_possible, n_others, n_valid), (min, max) magnitude of the field sources.
		rows: (nrows, res_ptr, sepmax_ptr, dist_post_ptr) device columns to select from instead of the context's own table
		(the gathered rows of all shards, nwb_maghist_select_rows)"""
		nsel = ctypes.c_int64()
		counts = (ctypes.c_int64 * 3)()
		mm = (ctypes.c_double * 2)()
		if rows is None:
			self.check(self.lib.nwb_maghist_select(self.h, int(c), int(k), int(bool(by_radius)), float(thr_select), float(thr_possible),
				int(bool(weights_cli)), ctypes.byref(nsel), counts, mm))
		else:
			self.check(self.lib.nwb_maghist_select_rows(self.h, int(c), int(k), int(rows[0]), rows[1], rows[2], rows[3], int(bool(by_radius)),
				float(thr_select), float(thr_possible), int(bool(weights_cli)), ctypes.byref(nsel), counts, mm))
		mag, w = numpy.empty(nsel.value), numpy.empty(nsel.value)
		if nsel.value:
			self.check(self.lib.nwb_maghist_sample(self.h, nsel.value, dptr(mag), dptr(w)))
		return mag, w, tuple(int(x) for x in counts), (float(mm[0]), float(mm[1]))

	def maghist_count(self, c, k, edges):
		edges = f64(edges)
		counts = numpy.zeros(len(edges) - 1, dtype=numpy.int64)
		self.check(self.lib.nwb_maghist_count(self.h, int(c), int(k), len(edges) - 1, dptr(edges), counts.ctypes.data_as(c_int64_p)))
		return counts

	def set_compat(self, flags):
		self.check(self.lib.nwb_set_compat(self.h, int(flags)))

	def bench_skeleton(self, c=1, reps=5, blocks_per_sm=0):
		"""mean duration (ms) of the memory-system skeleton of k_pairs on the grid of the last match (invalidates its result);
		blocks_per_sm = 0: at the residency of the match kernel"""
		v = ctypes.c_float(0)
		self.check(self.lib.nwb_bench_skeleton(self.h, int(c), int(reps), int(blocks_per_sm), ctypes.byref(v)))
		return float(v.value)

	def table_layout(self):
		"""(base device pointer, column stride in bytes, number of columns, rows) of the last match's table"""
		base, stride, ncols, nrows = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int(), ctypes.c_int64()
		self.check(self.lib.nwb_table_layout(self.h, ctypes.byref(base), ctypes.byref(stride), ctypes.byref(ncols), ctypes.byref(nrows)))
		return int(base.value or 0), int(stride.value), int(ncols.value), int(nrows.value)

	def table_view(self):
		"""the last match's table as a (ncols, nrows) int64 torch view of the context's own allocation (no copy; float
		columns are read with .view(torch.float64)); valid until the next match on this context"""
		import torch
		base, stride, ncols, nrows = self.table_layout()
		if nrows == 0 or base == 0:
			return torch.empty((ncols, 0), dtype=torch.int64, device=torch.device('cuda', self.device))
		flat = torch.as_tensor(DeviceView(base, ncols * (stride // 8)), device=torch.device('cuda', self.device))
		return flat.view(ncols, stride // 8)[:, :nrows]

	def flat_hash_applied(self):
		"""did the last match apply the reference's flat-sky bucket predicate (COMPAT_FLAT_HASH)?"""
		v = ctypes.c_int(0)
		self.check(self.lib.nwb_flat_hash_applied(self.h, ctypes.byref(v)))
		return bool(v.value)

	def set_tables(self, norm, log10e, prior, log10prior, sub_log10prior):
		a, b, c, d = f64(norm), f64(prior), f64(log10prior), f64(sub_log10prior)
		self.check(self.lib.nwb_set_tables(self.h, dptr(a), float(log10e), dptr(b), dptr(c), dptr(d)))

	def set_maghist(self, c, k, edges, weight, bias):
		e, w, b = f64(edges), f64(weight), f64(bias)
		assert len(e) == len(w) + 1 == len(b) + 1
		self.check(self.lib.nwb_set_maghist(self.h, c, k, len(w), dptr(e), dptr(w), dptr(b)))

	def set_primary_range(self, first, count):
		self.check(self.lib.nwb_set_primary_range(self.h, int(first), int(count)))

	def match(self, fuse_final=True):
		n = ctypes.c_int64(0)
		rc = self.lib.nwb_match(self.h, 1 if fuse_final else 0, ctypes.byref(n))
		if rc == NWB_ERR_EMPTY:
			return 0
		self.check(rc)
		return n.value

	def shard_setup(self, rank, world, spill_capacity=65536):
		"""enter shard mode (nwb_shard_setup); returns this rank's 64-byte IPC handle and the size of its exchange buffer"""
		h = ctypes.create_string_buffer(64)
		nbytes = ctypes.c_int64(0)
		self.check(self.lib.nwb_shard_setup(self.h, int(rank), int(world), int(spill_capacity), h, ctypes.byref(nbytes)))
		return h.raw, nbytes.value

	def shard_connect(self, handles):
		"""handles: the ranks' IPC handles in rank order (bytes, 64 each)"""
		blob = ctypes.create_string_buffer(b''.join(handles), 64 * len(handles))
		self.check(self.lib.nwb_shard_connect(self.h, blob))

	def shard_match(self, phase, fuse_final=True):
		"""one phase of a shard-mode match; phase 2 returns (rows, retry)"""
		n = ctypes.c_int64(0)
		rc = self.lib.nwb_shard_match(self.h, int(phase), 1 if fuse_final else 0, ctypes.byref(n))
		if rc == 1:
			return 0, True
		if rc == NWB_ERR_EMPTY:
			return 0, False
		self.check(rc)
		return n.value, False

	def shard_close(self):
		self.check(self.lib.nwb_shard_close(self.h))

	def gather_setup(self, rank, world, capacity_rows, ncols):
		"""nwb_gather_setup: this rank's gathered-table buffer (two sets of ncols x capacity_rows); returns its 64-byte IPC handle"""
		h = ctypes.create_string_buffer(64)
		self.check(self.lib.nwb_gather_setup(self.h, int(rank), int(world), int(capacity_rows), int(ncols), h))
		return h.raw

	def gather_connect(self, handles):
		blob = ctypes.create_string_buffer(b''.join(handles), 64 * len(handles))
		self.check(self.lib.nwb_gather_connect(self.h, blob))

	def gather_push(self, counts, engine=0):
		"""nwb_gather_push: enqueue the push of the last match's table into every rank's gathered table.  counts: the
		shards' row counts, a list (host) or a device pointer to world int64 (int: no host round trip).  Returns this
		rank's gathered table as a (ncols, capacity_rows) int64 torch view -- the first sum(counts) rows of every column
		are the table, complete after the caller's barrier"""
		import torch
		table = ctypes.c_void_p(0)
		stride = ctypes.c_int64(0)
		if counts is None:   # engine 2: the library exchanges the counts itself
			rc = self.lib.nwb_gather_push(self.h, None, 0, int(engine), ctypes.byref(table), ctypes.byref(stride))
		elif isinstance(counts, int):
			rc = self.lib.nwb_gather_push(self.h, ctypes.c_void_p(counts), 1, int(engine), ctypes.byref(table), ctypes.byref(stride))
		else:
			c = (ctypes.c_int64 * len(counts))(*[int(x) for x in counts])
			rc = self.lib.nwb_gather_push(self.h, ctypes.cast(c, ctypes.c_void_p), 0, int(engine), ctypes.byref(table), ctypes.byref(stride))
		self.check(rc)
		ncols = self.table_layout()[2]
		dev = torch.device('cuda', self.device)
		return torch.as_tensor(DeviceView(table.value, ncols * (stride.value // 8)), device=dev).view(ncols, stride.value // 8)

	def gather_counts(self, world):
		"""nwb_gather_counts: the ranks' row counts of the last engine-2 push (waits for the context's stream)"""
		c = (ctypes.c_int64 * world)()
		self.check(self.lib.nwb_gather_counts(self.h, c))
		return [int(x) for x in c]

	def gather_close(self):
		self.check(self.lib.nwb_gather_close(self.h))

	def score_rows(self, idx):
		"""nwb_score_rows: idx (R, ncat) int64 -> dict(sep (npairs, R), sepmax, ncat, log_bf, dist_post)"""
		idx = numpy.ascontiguousarray(idx, dtype=numpy.int64)
		R, nc = idx.shape
		npairs = nc * (nc - 1) // 2
		sep = numpy.empty((npairs, R))
		sepmax, lbf, post = numpy.empty(R), numpy.empty(R), numpy.empty(R)
		ncat = numpy.empty(R, dtype=numpy.int64)
		self.check(self.lib.nwb_score_rows(self.h, R, idx.ctypes.data_as(c_int64_p), dptr(sep), dptr(sepmax), ncat.ctypes.data_as(c_int64_p), dptr(lbf), dptr(post)))
		return dict(sep=sep, sepmax=sepmax, ncat=ncat, log_bf=lbf, dist_post=post)

	def match_async(self, fuse_final=True):
		"""enqueue a match on the context's stream without waiting for it (nwb_match_async); collect it with match_wait()"""
		rc = self.lib.nwb_match_async(self.h, 1 if fuse_final else 0)
		if rc != NWB_ERR_EMPTY:
			self.check(rc)

	def match_wait(self):
		n = ctypes.c_int64(0)
		rc = self.lib.nwb_match_wait(self.h, ctypes.byref(n))
		if rc == NWB_ERR_EMPTY:
			return 0
		self.check(rc)
		return n.value

	def finalize(self):
		self.check(self.lib.nwb_finalize(self.h))

	def truncate(self, min_prob):
		n = ctypes.c_int64(0)
		self.check(self.lib.nwb_truncate(self.h, float(min_prob), ctypes.byref(n)))
		return n.value

	def fetch(self, column, nrows, dtype=numpy.float64, out=None):
		if out is None:
			out = numpy.empty(nrows, dtype=dtype)
		assert out.dtype.itemsize == 8 and out.flags.c_contiguous and len(out) >= nrows
		self.check(self.lib.nwb_fetch(self.h, int(column), out.ctypes.data))
		return out

	def fetch_device(self, column, dst_ptr):
		self.check(self.lib.nwb_fetch_device(self.h, int(column), dst_ptr))

	def nrows_device_ptr(self):
		p = ctypes.c_void_p()
		self.check(self.lib.nwb_nrows_device_ptr(self.h, ctypes.byref(p)))
		return p.value

	def column_ptr(self, column):
		p = ctypes.c_void_p()
		self.check(self.lib.nwb_column_ptr(self.h, int(column), ctypes.byref(p)))
		return p.value

	def sync(self):
		self.check(self.lib.nwb_sync(self.h))

	def timings(self):
		out = {}
		v = ctypes.c_float()
		for k, name in enumerate(STAGE_NAMES):
			self.check(self.lib.nwb_timing(self.h, k, ctypes.byref(v)))
			out[name] = v.value
		return out

	def launch_count(self):
		n = ctypes.c_int64(0)
		self.check(self.lib.nwb_launch_count(self.h, ctypes.byref(n)))
		return n.value

	def stats(self):
		a = (ctypes.c_int64 * 4)()
		self.check(self.lib.nwb_stats(self.h, a))
		return dict(pairs_kept=a[1], grid_cells=a[2], cell_entries=a[3])


class DeviceView(object):
	"""wrap a raw device pointer for torch.as_tensor(..., device='cuda') via the CUDA array interface"""

	def __init__(self, ptr, n, typestr='<i8'):
		self.__cuda_array_interface__ = {'shape': (n,), 'typestr': typestr, 'data': (int(ptr), False), 'version': 2}


_contexts = {}


def get_context(device=None):
	"""process-wide context per device (scratch buffers are reused between calls)"""
	if device is None:
		device = int(os.environ.get('NWB_DEVICE', os.environ.get('LOCAL_RANK', '0')))
	if device not in _contexts:
		_contexts[device] = Context(device)
	return _contexts[device]
