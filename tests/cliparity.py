"""Parity helpers for the command-line surface: compare a CLI output table (dict column -> array in its FITS type)
with the digest of the real reference nway.py (tests/golden/ref_cli_<case>.npz).

The CLI writes float32 ('E') columns.  The oracle reproduces them bit for bit (exact=True).  The CUDA path computes
in fp64 to <= 1e-10 relative of the reference and is then rounded to float32 by the writer, so a value within 1e-10 of
a float32 rounding boundary may land on the neighbouring float32: tolerance 1 float32 ulp (1.2e-7 relative) --
2 ulp for dist_post / p_single / p_i, whose fp64 values differ by up to ~3e-11 relative -- and 2e-13 + 1 ulp absolute
for p_any (tests/parity.py explains that floor)."""
import hashlib
import os

import numpy as np

from tests import parity

ULP32 = float(np.finfo(np.float32).eps)


def load_cli_golden(name):
	return np.load(os.path.join(parity.GOLDEN, 'ref_cli_%s.npz' % name), allow_pickle=False)


def load_cli_transcript(name):
	"""stdout of the unmodified reference nway.py for the case, as lines (file names without their directory, the program
	called nway.py)"""
	return open(os.path.join(parity.GOLDEN, 'ref_cli_stdout_%s.txt' % name)).read().splitlines()


def normalise_transcript(text, directory):
	return text.replace(directory.rstrip(os.sep) + os.sep, '').splitlines()


def check_against_cli_digest(name, got, exact=False, check_layout=False, formats=None, header=None):
	"""got: mapping column -> array (the computed columns and the *_ID columns at least)"""
	g = load_cli_golden(name)
	names = [str(c) for c in g['columns']]
	ids = [str(c) for c in g['id_columns']]
	if check_layout:
		assert list(got.keys()) == names, (list(got.keys()), names)
		assert [formats[n] for n in names] == [str(f) for f in g['formats']]
		for k in ('TABLES', 'BIASING', 'COLS_RA', 'COLS_DEC', 'COL_PRIM', 'COLS_ERR'):
			assert str(header[k]) == str(g['hdr_' + k]), (k, header[k], str(g['hdr_' + k]))
	idx = np.stack([np.asarray(got[n]).astype(np.int64) for n in ids], axis=1)
	assert len(idx) == int(g['nrows']), (len(idx), int(g['nrows']))
	sha = np.frombuffer(hashlib.sha256(np.ascontiguousarray(idx).tobytes()).digest(), dtype=np.uint8)
	assert (sha == g['idx_sha256']).all(), 'row set / order differs from the reference command-line program'
	sel = g['sample_rows']
	report = []
	for n in names:
		if n not in got:
			continue
		ref = g['col_' + n]
		val = np.asarray(got[n])[sel]
		assert val.dtype == ref.dtype, (n, val.dtype, ref.dtype)
		if ref.dtype.kind in 'iubS':
			assert (val == ref).all(), (name, n)
			continue
		nan_ok = np.isnan(ref) == np.isnan(val)
		assert nan_ok.all(), (name, n, 'NaN pattern')
		r = np.where(np.isnan(ref), 0, ref).astype(np.float64)
		v = np.where(np.isnan(val), 0, val).astype(np.float64)
		d = np.abs(r - v)
		if exact:
			tol = np.zeros_like(d)
		else:
			k = 2.0 if n in ('dist_post', 'p_single', 'p_i') else 1.0
			tol = k * ULP32 * np.abs(r) + (2e-13 if n == 'p_any' else 0.0) + 1e-45
		bad = d > tol
		report.append('%-28s max |d| %.3e  (%d of %d sample rows differ at all)' % (n, d.max() if len(d) else 0, int((d > 0).sum()), len(d)))
		assert not bad.any(), (name, n, int(bad.sum()), ref[bad][:3], val[bad][:3])
		if 'sum_' + n in g.files:
			f = np.asarray(got[n]).astype(np.float64)
			s = np.nansum(f[np.isfinite(f)])
			assert np.isclose(s, float(g['sum_' + n]), rtol=0 if exact else 1e-6, atol=0 if exact else 1e-6), (name, n, s, float(g['sum_' + n]))
	return report


def oracle_cli_table(name):
	"""the case through oracle/nway_oracle.py in cli_compat mode, as CLI-named / CLI-typed columns"""
	from oracle import make_golden_cli
	paths = {'XMM': 'XMM.fits', 'OPT': 'OPT.fits', 'IRAC': 'IRAC.fits'}
	return make_golden_cli.oracle_table(name, paths)
