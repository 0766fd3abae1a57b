"""TEST INFRASTRUCTURE: a CPU stand-in for nway_b200._lib.Context whose numeric stages are the oracle's
(oracle/nway_oracle.py, piece by piece).  It exists so that the HOST layers above the C ABI -- the orchestration of
nway_b200.nway_match (scalar tables, histogram installation, the automatic-histogram host code, truncation, logging,
exceptions), nway_b200/cli.py (arguments, the output table, header keys, the stdout transcript) and the FITS writer --
run end to end under `-m "not gpu"`, against the same committed outputs of the unmodified reference as the GPU tests.
Nothing here is measured or shipped, and the product never sees it: a test installs it with
`monkeypatch.setattr(nway_b200._lib, 'get_context', ...)` (tests/test_host_layer_cpu.py).

The stage boundaries are the library's (include/nwayb200.h): set_catalogue / set_params / set_compat / set_prefilter /
set_maghist -> match(fuse_final) -> [maghist_select / maghist_count -> set_maghist -> finalize] -> truncate -> fetch."""
import ctypes

import numpy as np

from nway_b200 import _lib as L
from oracle import nway_oracle as O


class OracleContext(object):
	def __init__(self):
		self.tables = []
		self.hists = {}
		self.primary_range = (0, -1)
		self.flags = 0
		self.prefilter = []
		self.cols = None
		self.calls = []

	# ---- inputs -----------------------------------------------------------------------------------------------
	def set_catalogue(self, c, ncat, ra, dec, err, area, mags=(), err_kind=L.ERR_CIRCULAR):
		if len(self.tables) != ncat:
			self.tables = [None] * ncat
		e = np.asarray(err, dtype=float)
		n = len(ra)
		assert e.size == n * err_kind
		error = tuple(e.reshape(3, n)) if err_kind == L.ERR_ELLIPSE else e.reshape(n)
		self.tables[c] = dict(name='T%d' % c, ra=np.asarray(ra, dtype=float), dec=np.asarray(dec, dtype=float), error=error,
			area=float(area), mags=[np.asarray(m, dtype=float) for m in mags])   # the device column: the caller's values widened
		self.hists = {k: v for k, v in self.hists.items() if k[0] != c}
		self.cols = None

	def set_primary_range(self, first, count):
		self.primary_range = (int(first), int(count))

	def set_params(self, radius, completeness, ratio_secondary=0.5, unrelated_mode=L.UNRELATED_API):
		self.radius, self.pc = float(radius), np.asarray(completeness, dtype=float)
		self.ratio_secondary, self.mode = float(ratio_secondary), int(unrelated_mode)

	def set_compat(self, flags):
		self.flags = int(flags)

	def set_prefilter(self, pairwise_errs):
		self.prefilter = [(int(a), int(b), float(r)) for a, b, r in pairwise_errs]

	def set_tables(self, norm, log10e, prior, log10prior, sub_log10prior):
		# the host scalars of nway_b200._scalar_tables against the oracle's own: the prior of every presence pattern
		# (bit 0 of the pattern = catalogue 1 present, __init__.py:254)
		nu, nu_plus = O.source_densities(self.tables)
		n = len(self.tables)
		for mask in range(2 ** (n - 1)):
			sel = np.array([True] + [(mask >> (c - 1)) & 1 == 1 for c in range(1, n)])
			assert prior[mask] == nu[0] * np.prod(self.pc[sel]) / np.prod(nu_plus[sel]), mask
		assert log10e == O.LOG10_E

	def set_maghist(self, c, k, edges, weight, bias):
		edges, weight, bias = (np.asarray(x, dtype=float) for x in (edges, weight, bias))
		assert len(edges) == len(weight) + 1 == len(bias) + 1 <= L.MAX_HIST_BINS + 1
		self.hists[(c, k)] = (edges, weight, bias)

	# ---- stages -----------------------------------------------------------------------------------------------
	def match(self, fuse_final=True):
		self.calls.append('match(fuse_final=%d)' % bool(fuse_final))
		t = self.tables
		mt = O.create_match_table(t, self.radius, enumerator='reference' if self.flags & L.COMPAT_FLAT_HASH else 'complete',
			sep_f32=bool(self.flags & L.COMPAT_SEP_F32), pairwise_errs=self.prefilter)
		first, count = self.primary_range
		if count >= 0:   # a shard: the rows of these primaries; densities stay those of the whole catalogues
			p = mt['idx'][:, 0]
			mt = self._take(mt, (p >= first) & (p < first + count))
		self.mt = mt
		idx = mt['idx']
		self.cols = None
		if len(idx) == 0:
			return 0
		nu, nu_plus = O.source_densities(t)
		self.prior, lbf = O.single_log_bf(mt, nu, nu_plus, self.pc)
		self.starts = O.group_starts(idx[:, 0])
		corr = O.correct_unrelated_cli(mt, lbf, nu, nu_plus, self.starts) if self.mode == L.UNRELATED_CLI else lbf
		n = len(t)
		cols = {}
		for c in range(n):
			cols[L.COL_IDX + c] = idx[:, c]
		for k, (a, b) in enumerate((a, b) for a in range(n) for b in range(a + 1, n)):
			cols[L.COL_SEP + k] = mt['sep'][(a, b)].astype(float)
		cols[L.COL_SEPMAX] = mt['sepmax']
		cols[L.COL_NCAT] = mt['ncat'].astype(np.int64)
		cols[L.COL_LOGBF_UNCORR] = lbf
		cols[L.COL_LOGBF] = corr
		cols[L.COL_DIST_POST] = O.posterior(self.prior, corr)
		self.cols = cols
		self.finalized = False
		if fuse_final:
			self._finalize()
		return len(idx)

	@staticmethod
	def _take(mt, keep):
		out = dict(idx=mt['idx'][keep], sep={k: v[keep] for k, v in mt['sep'].items()}, sepmax=mt['sepmax'][keep], ncat=mt['ncat'][keep])
		out['errors'] = [tuple(x[keep] for x in e) if isinstance(e, tuple) else e[keep] for e in mt['errors']]
		if 'off' in mt:
			out['off'] = {k: (a[keep], b[keep]) for k, (a, b) in mt['off'].items()}
		return out

	def finalize(self):
		self.calls.append('finalize')
		self._finalize()

	def _finalize(self):
		assert self.cols is not None
		idx = self.mt['idx']
		wsum = 0   # ((0 + w1) + w2): the reference's sum(biases.values()) (__init__.py:394)
		nbias = 0
		for c, t in enumerate(self.tables):
			for k, magvals in enumerate(t['mags']):
				edges, weight, bias = self.hists[(c, k)]   # KeyError: a magnitude column whose histogram was never installed
				res = idx[:, c]
				m = magvals[res]
				with np.errstate(invalid='ignore'):
					inside = (res != -1) & np.isfinite(m) & (m >= edges[0]) & (m <= edges[-1])
				b = np.clip(np.searchsorted(edges, m, side='right') - 1, 0, len(weight) - 1)   # zero-order hold, last edge inclusive
				self.cols[L.COL_BIAS + nbias] = np.where(inside, bias[b], 1.0)
				wsum = wsum + np.where(inside, weight[b], 0.0)
				nbias += 1
		total = self.cols[L.COL_LOGBF] + wsum
		self.cols[L.COL_P_SINGLE] = O.posterior(self.prior, total)
		p_any, p_i, flag = O.group_statistics(total + np.log10(self.prior), self.starts, self.ratio_secondary)
		self.cols[L.COL_MATCH_FLAG], self.cols[L.COL_P_ANY], self.cols[L.COL_P_I] = flag, p_any, p_i
		self.finalized = True

	def truncate(self, min_prob):
		self.calls.append('truncate')
		assert self.finalized
		keep = ~(self.cols[L.COL_P_I] < min_prob)
		self.cols = {k: v[keep] for k, v in self.cols.items()}
		self.mt = self._take(self.mt, keep)
		return int(keep.sum())

	# ---- automatic histograms: the device half of nwaylib/__init__.py:324-366 (nwb_maghist_select / _count) ---------------
	def maghist_select(self, c, k, by_radius, thr_select, thr_possible, weights_cli, rows=None):
		magvals = self.tables[c]['mags'][k]
		if rows is None:
			res = self.mt['idx'][:, c]
			quantity = self.cols[L.COL_SEPMAX] if by_radius else self.cols[L.COL_DIST_POST]
		else:
			# nwb_maghist_select_rows: three columns of ALL shards' rows, handed over as addresses (nway_b200.parallel.hist_rows)
			n, res_ptr, sepmax_ptr, post_ptr = rows
			read = lambda ptr, ctype: np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctype)), (n,)).copy()
			res = read(res_ptr, ctypes.c_int64)
			quantity = read(sepmax_ptr if by_radius else post_ptr, ctypes.c_double)
		defined = res != -1
		selection = ((quantity < thr_select) if by_radius else (quantity > thr_select)) & defined
		possible = ((quantity < thr_possible) if by_radius else (quantity > thr_possible)) & defined
		sw = np.ones(len(res)) if by_radius else quantity
		sw = sw[selection] if weights_cli else sw[defined]   # the API indexes the compressed weights by positions in res[selection]
		sources, firsts = np.unique(res[selection], return_index=True)
		valid = np.isfinite(magvals)
		others = valid.copy()
		possible_sources = np.unique(res[possible])
		others[possible_sources] = False
		self.others = (c, k, others)
		field = magvals[others]
		return (magvals[sources], sw[firsts], (len(possible_sources), int(others.sum()), int(valid.sum())),
			(float(np.min(field)) if len(field) else np.nan, float(np.max(field)) if len(field) else np.nan))

	def maghist_count(self, c, k, edges):
		assert self.others[:2] == (c, k)
		return np.histogram(self.tables[c]['mags'][k][self.others[2]], bins=np.asarray(edges, dtype=float))[0].astype(np.int64)

	# ---- outputs ----------------------------------------------------------------------------------------------
	def fetch(self, column, nrows, dtype=np.float64, out=None):
		col = self.cols[int(column)]
		assert len(col) == nrows
		return np.ascontiguousarray(col, dtype=dtype)

	# ---- the table as nway_b200.parallel reads it: one (ncols, rows) block of 8-byte words in the output column order --------
	torch_device = 'cpu'

	def _selectors_in_order(self):
		n = len(self.tables)
		nbias = sum(len(t['mags']) for t in self.tables)
		return ([L.COL_IDX + c for c in range(n)] + [L.COL_SEP + k for k in range(n * (n - 1) // 2)] +
			[L.COL_SEPMAX, L.COL_NCAT, L.COL_LOGBF_UNCORR, L.COL_LOGBF, L.COL_DIST_POST] + [L.COL_BIAS + k for k in range(nbias)] +
			[L.COL_P_SINGLE, L.COL_MATCH_FLAG, L.COL_P_ANY, L.COL_P_I])

	def table_view(self):
		import torch
		sel = self._selectors_in_order()
		nrows = 0 if self.cols is None else len(self.cols[L.COL_IDX])
		block = np.zeros((len(sel), nrows), dtype=np.int64)
		for k, s in enumerate(sel):
			if self.cols is not None and s in self.cols:   # the columns of the final stage exist after finalize only
				col = self.cols[s]
				block[k] = col if col.dtype.kind in 'iu' else np.ascontiguousarray(col, dtype=np.float64).view(np.int64)
		self._block = torch.from_numpy(block)   # stays alive until the next view, like the library's allocation until the next match
		return self._block

	def table_layout(self):
		view = self.table_view()
		return view.data_ptr(), 8 * view.shape[1], view.shape[0], view.shape[1]

	def row_offsets(self, a, b, nrows):
		"""(dra, ddec) in arcsec between the members a < b of every row (nwb_row_offsets), NaN where one is absent"""
		dra, ddec = (x.copy() for x in self.mt['off'][(a, b)])
		absent = (self.mt['idx'][:, a] == -1) | (self.mt['idx'][:, b] == -1)
		dra[absent], ddec[absent] = np.nan, np.nan
		assert len(dra) == nrows
		return dra, ddec

	def flat_hash_applied(self):
		radec = [(t['ra'], t['dec']) for t in self.tables]
		return bool(self.flags & L.COMPAT_FLAT_HASH) and O.flat_sky_applicable(radec, self.radius / 60. / 60)

	def sync(self):
		pass

	def timings(self):
		return {}

	def launch_count(self):
		return 0
