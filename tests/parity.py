"""Column-wise parity metric, shared by the oracle and the CUDA tests.

integer / index columns : exact
prob_has_match (p_any)  : |d| <= 1e-10*|ref| + 2e-13     (absolute floor: p_any = 1 - 10^x cancels, SURVEY fact 4;
                          x = v0 - bfsum carries the rounding of log-weights of magnitude ~25, i.e. a few 1e-14,
                          and d p_any = ln(10) * 10^x * dx -- measured worst case 3.4e-14 at p_any = 1.9e-7)
prob_this_match (p_i)   : |d| <= 1e-10*|ref| for ref >= 1e-30, |d| <= 1e-40 below
dist_bayesfactor*       : |d| <= 1e-9 + 1e-12*|ref|
separations             : |d| <= 1e-9 arcsec relative 1e-10
dist_post, p_single, bias_* : relative 1e-10 above 1e-30
"""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

RTOL = 1e-10


def column_error(name, ref, got, rtol=None):
	"""returns (ok, worst absolute error, worst relative error, index of worst)"""
	ref = np.asarray(ref)
	got = np.asarray(got)
	assert ref.shape == got.shape, (name, ref.shape, got.shape)
	RTOL = globals()['RTOL'] if rtol is None else rtol
	if ref.size == 0:
		return True, 0.0, 0.0, -1
	if ref.dtype.kind in 'iub':
		bad = ref != got
		return (not bad.any()), float(bad.sum()), 0.0, int(np.argmax(bad))
	nan_r, nan_g = np.isnan(ref), np.isnan(got)
	if (nan_r != nan_g).any():
		return False, np.inf, np.inf, int(np.argmax(nan_r != nan_g))
	r = np.where(nan_r, 0.0, ref)
	g = np.where(nan_g, 0.0, got)
	same_inf = np.isinf(r) & (r == g)
	r = np.where(same_inf, 0.0, r)
	g = np.where(same_inf, 0.0, g)
	with np.errstate(invalid='ignore'):
		d = np.abs(r - g)
		d = np.where(np.isnan(d), np.inf, d)
	if name == 'prob_has_match':
		tol = RTOL * np.abs(r) + 2e-13
	elif name.startswith('dist_bayesfactor'):
		tol = max(1e-9, 10 * RTOL) + 1e-12 * np.abs(r)
	elif name.startswith('Separation'):
		tol = 1e-9 + RTOL * np.abs(r)
	else:
		tol = np.maximum(RTOL * np.abs(r), 1e-40)
	worst = int(np.argmax(d - tol))
	with np.errstate(divide='ignore', invalid='ignore'):
		rel = np.where(r != 0, d / np.abs(r), 0.0)
	return bool((d <= tol).all()), float(d.max()), float(rel[np.abs(r) >= 1e-30].max() if (np.abs(r) >= 1e-30).any() else 0.0), worst


def assert_tables_match(ref, got, columns=None, context='', rtol=None):
	"""ref / got: mappings column -> array.  Row set and order must be identical; see module docstring."""
	columns = columns or [c for c in ref.keys() if not c.startswith('_')]
	report = []
	failed = []
	for c in columns:
		assert c in got, '%s: column %s missing (have %s)' % (context, c, list(got.keys()))
		a, b = np.asarray(ref[c]), np.asarray(got[c])
		assert len(a) == len(b), '%s: column %s has %d rows, expected %d' % (context, c, len(b), len(a))
		ok, dabs, drel, worst = column_error(c, a, b, rtol)
		report.append('%-30s %s  max|d| %.3e  max rel %.3e' % (c, 'ok  ' if ok else 'FAIL', dabs, drel))
		if not ok:
			failed.append('%s: %s row %d ref %r got %r' % (context, c, worst, a[worst], b[worst]))
	assert not failed, '\n'.join(failed + report)
	return report


def load_golden(name):
	return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def check_against_digest(name, got, names, rtol=None):
	g = load_golden('ref_%s.npz' % name)
	idx = np.stack([got[n] for n in names], axis=1).astype(np.int64)
	assert len(idx) == int(g['nrows'])
	sha = np.frombuffer(hashlib.sha256(np.ascontiguousarray(idx).tobytes()).digest(), dtype=np.uint8)
	assert (sha == g['idx_sha256']).all(), 'row set / order differs from the reference'
	starts = np.concatenate(([0], np.flatnonzero(np.diff(idx[:, 0]) != 0) + 1)) if len(idx) else np.zeros(0, dtype=np.int64)
	ok, dabs, drel, worst = column_error('prob_has_match', g['p_any'], np.asarray(got['prob_has_match'])[starts], rtol)
	assert ok, ('p_any', name, worst, dabs, drel)
	sel = g['sample_rows']
	ref = {str(c): g['col_' + str(c)] for c in g['columns']}
	assert_tables_match(ref, {c: np.asarray(got[c])[sel] for c in ref}, context=name, rtol=rtol)
	for c in ref:
		v = np.asarray(got[c])
		s = np.nansum(v[np.isfinite(v)]) if v.dtype.kind == 'f' else v.sum()
		assert np.isclose(float(s), float(g['sum_' + c]), rtol=1e-9, atol=1e-9), (name, c, s, g['sum_' + c])


