"""GPU tests of the command-line surface (nway.py / nway_b200.cli) against golden outputs of the UNMODIFIED
reference nway.py (tests/golden/ref_cli_*.npz, oracle/make_golden_cli.py): same columns in the same order and FITS
formats, same header keys, same rows, values within 1-2 float32 ulp (tests/cliparity.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import cases, cliparity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_cli(name, tmp_path, extra=()):
	from nway_b200 import cli, fitsio
	paths = cases.write_cosmos_subset_fits(str(tmp_path))
	out = str(tmp_path / (name + '.fits'))
	cwd = os.getcwd()
	os.chdir(str(tmp_path))   # *_fit.txt files land in the working directory, like the reference's
	try:
		rc = cli.main(cases.cli_args(name, paths, out) + list(extra))
	finally:
		os.chdir(cwd)
	assert rc == 0
	t = fitsio.read_table(out)
	cards, _ = fitsio._read_header(open(out, 'rb').read(), 0)
	return t, cards


@pytest.mark.parametrize('name', ['cli2', 'cli3', 'cli3_magauto', 'cli3_bayes', 'cli3_minprob', 'cli3_prefilter'])
def test_cli_against_reference_cli(name, tmp_path, capsys):
	extra = []   # --prefilter-mode reference is the default: the drop-in entry point returns the reference's rows
	t, cards = run_cli(name, tmp_path, extra)
	# stdout, line for line what the unmodified script prints (tests/golden/ref_cli_stdout_*.txt) -- the histogram
	# populations in it ("... secure matches, ... insecure matches and ... secure non-matches") are counted on the device
	printed = capsys.readouterr().out
	assert cliparity.normalise_transcript(printed, str(tmp_path)) == cliparity.load_cli_transcript(name)
	got = {n: t.data[n] for n in t.columns}
	report = cliparity.check_against_cli_digest(name, got, check_layout=True, formats=dict(zip(t.columns, t.formats)), header=cards)
	print('\n'.join(report))
	assert t.name == 'NWAYMATCH'
	assert cards['METHOD'] == 'NWAY multi-way matching'


def test_cli_prefilter_fixed_against_oracle(tmp_path):
	"""--prefilter-pair as documented (the reference's own implementation is broken): against the oracle"""
	from oracle import nway_oracle as O
	t, cards = run_cli('cli3_prefilter', tmp_path, ['--prefilter-mode', 'fixed'])
	tabs = cases.cosmos_subset(3)
	tabs[1]['error'] = 0.1 * np.ones(len(tabs[1]['ra']))
	tabs[2]['error'] = 0.5 * np.ones(len(tabs[2]['ra']))
	ref = O.nway_match(tabs, 12., 1.0, unrelated_mode='cli', cli_compat=True, pairwise_errs=[(1, 2, 0.5)])
	assert len(t) == len(ref['XMM'])
	assert (np.where(ref['OPT'] >= 0, ref['OPT'] + 1, -99) == t.data['OPT_ID']).all()
	assert (np.where(ref['IRAC'] >= 0, ref['IRAC'] + 1, -99) == t.data['IRAC_ID']).all()
	both = (ref['OPT'] >= 0) & (ref['IRAC'] >= 0)
	assert both.any() and (t.data['Separation_IRAC_OPT'][both] < 0.5).all()
	for mine, theirs in (('p_any', 'prob_has_match'), ('p_i', 'prob_this_match'), ('dist_bayesfactor_corrected', 'dist_bayesfactor')):
		a, b = t.data[mine].astype(np.float64), ref[theirs].astype(np.float32).astype(np.float64)
		assert np.allclose(a, b, rtol=3e-7, atol=1e-12), mine


def test_cli_script_and_elliptical(tmp_path):
	"""the top-level nway.py script end to end, with :ra_err:dec_err and :major:minor:angle error specifications"""
	from nway_b200 import fitsio
	from oracle import nway_oracle as O
	rng = np.random.default_rng(11)
	tabs = cases.uniform_patch(3, (400, 6000, 5000), (1.0, 0.3, 0.5), 0.05)
	n0, n1 = len(tabs[0]['ra']), len(tabs[1]['ra'])
	maj = rng.uniform(0.5, 3, n0); mnr = rng.uniform(0.2, 1, n0) * maj; ang = rng.uniform(0, 180, n0)
	era = rng.uniform(0.2, 0.6, n1); edec = rng.uniform(0.2, 0.6, n1)
	files = []
	for t, extra in zip(tabs, ([('emaj', 'D', maj), ('emin', 'D', mnr), ('eang', 'D', ang)], [('era', 'D', era), ('edec', 'D', edec)], [])):
		n = len(t['ra'])
		cols = [fitsio.Column('ID', 'K', np.arange(n) + 100), fitsio.Column('RA', 'D', t['ra']), fitsio.Column('DEC', 'D', t['dec'])]
		cols += [fitsio.Column(*e) for e in extra]
		path = str(tmp_path / ('%s.fits' % t['name']))
		fitsio.write_table(path, cols, t['name'], table_header=[('SKYAREA', t['area'])])
		files.append(path)
	out = str(tmp_path / 'ell.fits')
	cmd = [sys.executable, os.path.join(ROOT, 'nway.py'), '--radius', '6', '--prior-completeness', '0.9', files[0], ':emaj:emin:eang',
		files[1], ':era:edec', files[2], '0.5', '--out', out]
	res = subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True)
	assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
	t = fitsio.read_table(out)
	names = [x['name'] for x in tabs]
	assert 'Separation_%s_%s_ra' % (names[1], names[0]) in t.columns and 'Separation_%s_%s_dec' % (names[2], names[1]) in t.columns
	tabs[0]['error'] = tuple(O.ellipse_from_cli(maj, mnr, ang))
	tabs[1]['error'] = (era, edec, np.zeros(n1))
	tabs[2]['error'] = 0.5 * np.ones(len(tabs[2]['ra']))
	ref = O.nway_match(tabs, 6., 0.9, unrelated_mode='cli', cli_compat=True)
	assert len(t) == len(ref[names[0]])
	for c, nm in enumerate(names):
		assert (np.where(ref[nm] >= 0, ref[nm] + 100, -99) == t.data[nm + '_ID']).all()
	# float32 offsets feed an fp64 computation on both sides; the oracle's offsets differ from the device's by ~1e-16
	# relative, which can move a float32 rounding: compare at float32 resolution of the offsets' effect
	for mine, theirs, tol in (('p_any', 'prob_has_match', 2e-5), ('p_i', 'prob_this_match', 2e-5), ('dist_bayesfactor', 'dist_bayesfactor_uncorrected', 2e-5)):
		a, b = t.data[mine].astype(np.float64), ref[theirs]
		assert np.allclose(a, b, rtol=tol, atol=1e-6), (mine, np.abs(a - b).max())
	assert (t.data['match_flag'] == ref['match_flag']).mean() > 0.999
