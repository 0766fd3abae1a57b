"""Generate tests/golden/ref_cli_*.npz from the UNMODIFIED reference command-line program (build container only).

    python oracle/make_golden_cli.py [--stdout-only]

For every case of tests/cases.CLI_CASES the real /root/reference/nway.py is executed (oracle/refcli.py) on FITS
files of the COSMOS subset; the digest holds the column names and FITS formats of its output table, the row count,
a SHA-256 of the three ID columns, float64 sums of every computed column and a strided sample of full rows.
The oracle (cli_compat mode) is checked against the same output, bit for bit after the cast to the column format.
It also asserts that the harness reproduces the reference's own published logs on the full demo catalogues
(doc/logs/match2:30 -- 37836 rows, 17 columns).
"""
import hashlib
import os
import re
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

from oracle import nway_oracle as O  # noqa: E402
from tests import cases  # noqa: E402

COMPUTED = re.compile(r'^(Separation_|ncat$|dist_|bias_|p_single$|p_any$|p_i$|match_flag$)')


def digest(tab, stride):
	cols = tab['columns']
	names = list(cols.keys())
	ids = [n for n in names if n.endswith('_ID')]
	idx = np.stack([cols[n].astype(np.int64) for n in ids], axis=1)
	out = dict(nrows=np.int64(len(idx)), id_columns=np.array(ids), columns=np.array(names),
		formats=np.array([tab['formats'][n] for n in names]),
		idx_sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(idx).tobytes()).digest(), dtype=np.uint8))
	sel = np.unique(np.concatenate((np.arange(0, min(len(idx), 400)), np.arange(0, len(idx), stride))))
	out['sample_rows'] = sel
	for n in names:
		v = cols[n]
		out['col_' + n] = v[sel]
		if COMPUTED.match(n):
			f = v.astype(np.float64)
			out['sum_' + n] = np.float64(np.nansum(f[np.isfinite(f)]))
	hdr = tab['primary_header']
	for k in ('TABLES', 'BIASING', 'COLS_RA', 'COLS_DEC', 'COL_PRIM', 'COLS_ERR'):
		out['hdr_' + k] = np.array(str(hdr[k]))
	return out


def oracle_table(name, paths):
	"""the same case through the oracle port in cli_compat mode, as CLI-named, CLI-typed columns"""
	import argparse
	argv = cases.cli_args(name, paths, 'x.fits')
	p = argparse.ArgumentParser()
	p.add_argument('--radius', type=float)
	p.add_argument('--mag-radius', type=float, default=None)
	p.add_argument('--mag-auto-minprob', type=float, default=0.9)
	p.add_argument('--prior-completeness', default='1')
	p.add_argument('--ignore-unrelated-associations', dest='unrel', action='store_false')
	p.add_argument('--mag', nargs=2, action='append', default=[])
	p.add_argument('--acceptable-prob', type=float, default=0.5)
	p.add_argument('--min-prob', type=float, default=0)
	p.add_argument('--out')
	p.add_argument('--prefilter-pair', nargs=3, action='append', default=[])
	p.add_argument('catalogues', nargs='+')
	a = p.parse_args(argv)
	tabs = cases.cosmos_subset(len(a.catalogues) // 2)
	names = [t['name'] for t in tabs]
	for t, spec in zip(tabs, a.catalogues[1::2]):
		if not spec.startswith(':'):
			t['error'] = float(spec) * np.ones(len(t['ra']))
	z = np.load(os.path.join(GOLDEN, 'cosmos_subset.npz'))
	for mag, magfile in a.mag:
		tn, cn = mag.split(':')
		t = tabs[names.index(tn)]
		t['mags'].append(z[tn + '_mag'].copy())
		t['magnames'].append(cn)
		t['maghists'].append(None)
	if ':' in a.prior_completeness:
		pc = np.array([1.0] + [float(x) for x in a.prior_completeness.split(':')])
	else:
		pc = float(a.prior_completeness)
	pw = [(names.index(x), names.index(y), 0.0) for x, y, r in a.prefilter_pair]   # the reference AS WRITTEN: radius 0 (Q8)
	out = O.nway_match(tabs, a.radius, pc, mag_include_radius=a.mag_radius, magauto_post_single_minvalue=a.mag_auto_minprob,
		prob_ratio_secondary=a.acceptable_prob, min_prob=a.min_prob, unrelated_mode='cli' if a.unrel else 'api', cli_compat=True,
		pairwise_errs=pw)
	m = {}
	n = len(names)
	for i in range(n):
		for j in range(i):
			m['Separation_%s_%s' % (names[i], names[j])] = out['Separation_%s_%s' % (names[j], names[i])].astype(np.float32)
	m['Separation_max'] = out['Separation_max'].astype(np.float32)
	m['ncat'] = out['ncat'].astype(np.int16)
	m['dist_bayesfactor'] = out['dist_bayesfactor_uncorrected'].astype(np.float32)
	m['dist_bayesfactor_corrected'] = out['dist_bayesfactor'].astype(np.float32)
	m['dist_post'] = out['dist_post'].astype(np.float32)
	for k in out:
		if k.startswith('bias_'):
			m[k] = out[k].astype(np.float32)
	m['p_single'] = out['p_single'].astype(np.float32)
	m['p_any'] = out['prob_has_match'].astype(np.float32)
	m['p_i'] = out['prob_this_match'].astype(np.float32)
	m['match_flag'] = out['match_flag'].astype(np.int16)
	for c, nm in enumerate(names):
		m[nm + '_ID'] = np.where(out[nm] >= 0, out[nm] + 1, -99).astype(np.int32)
	return m


def transcripts():
	"""tests/golden/ref_cli_stdout_<case>.txt: what the unmodified nway.py prints to stdout for every case (its progress
	bars and the matching stage's log go to stderr), with the scratch directory cut out of the file names"""
	from oracle import refcli, refrun
	work = tempfile.mkdtemp(prefix='nwb_refcli_')
	paths = cases.write_cosmos_subset_fits(work)
	for name in cases.CLI_CASES:
		tab, log = refcli.run_cli(cases.cli_args(name, paths, name + '.fits'), work)
		with open(os.path.join(GOLDEN, 'ref_cli_stdout_%s.txt' % name), 'w') as f:
			f.write(log.replace(work + os.sep, '').replace(os.path.join(refrun.REFERENCE_ROOT, 'nway.py'), 'nway.py'))
		print('%-16s %d lines of stdout' % (name, len(log.splitlines())))


def main():
	from oracle import refcli, refrun   # the real reference: build container only
	os.makedirs(GOLDEN, exist_ok=True)
	work = tempfile.mkdtemp(prefix='nwb_refcli_')
	# ---- the harness against the reference's own published log (full demo catalogues) ---------------------------
	D = os.path.join(refrun.REFERENCE_ROOT, 'doc')
	tab, log = refcli.run_cli(['--radius', '20', os.path.join(D, 'COSMOS_XMM.fits'), ':pos_err', os.path.join(D, 'COSMOS_OPTICAL.fits'), '0.1',
		'--out=example2.fits'], work)
	want = open(os.path.join(D, 'logs', 'match2')).read()
	line = [l for l in want.splitlines() if 'writing "example2.fits"' in l][0].strip()
	assert line in log, (line, log[-300:])
	print('doc/logs/match2 reproduced:', line)

	paths = cases.write_cosmos_subset_fits(work)
	for name in cases.CLI_CASES:
		tab, log = refcli.run_cli(cases.cli_args(name, paths, name + '.fits'), work)
		orc = oracle_table(name, paths)
		cols = tab['columns']
		for k, v in cols.items():
			if k not in orc:
				assert not COMPUTED.match(k), k
				continue
			a, b = v, orc[k]
			assert a.shape == b.shape and a.dtype == b.dtype, (name, k, a.shape, b.shape, a.dtype, b.dtype)
			same = (a == b) | ((a != a) & (b != b))
			assert same.all(), (name, k, int((~same).sum()), a[~same][:3], b[~same][:3])
		d = digest(tab, 101 if len(cols['p_any']) > 50000 else 23)
		np.savez_compressed(os.path.join(GOLDEN, 'ref_cli_%s.npz' % name), **d)
		print('%-16s rows %7d  cols %2d  sum p_any %.6f  (oracle: identical)' % (name, len(cols['p_any']), len(cols), d['sum_p_any']))


if __name__ == '__main__':
	if '--stdout-only' in sys.argv:
		transcripts()
	else:
		main()
		transcripts()
