"""host-side pieces of the calibration loop (no GPU): the cut-off table and the great-arc interpolation"""
import numpy as np


def test_calibrate_cutoff_table():
	from nway_b200 import calibrate
	rng = np.random.default_rng(0)
	real = dict(ncat=np.r_[np.ones(1000), 2 * np.ones(500)], p_any=np.r_[rng.beta(5, 2, 1000), rng.uniform(size=500)])
	fake = dict(ncat=np.ones(1000), p_any=rng.beta(1, 6, 1000))
	cut, eff, err, lines = calibrate.calibrate_cutoff(real, fake)
	assert len(cut) == 101 and cut[0] == 0 and cut[-1] == 1
	for c, e, f in zip(cut, eff, err):   # nway-calibrate-cutoff.py:58-59
		assert e == (real['p_any'][:1000] > c).mean() and f == (fake['p_any'] > c).mean()
	text = '\n'.join(lines)
	i = np.min(np.where(err < 0.05)[0])
	assert 'For a false detection rate of <5%' in text
	assert '--> use only counterparts with p_any>%.2f (%.2f%% of matches)' % (cut[i], eff[i] * 100) in text
	_, _, _, lines = calibrate.calibrate_cutoff(real, dict(ncat=np.ones(10), p_any=np.ones(10) * 2))
	assert 'A false detection rate of <1% is not possible.' in '\n'.join(lines)


def test_greatarc_interpolate():
	from nway_b200 import calibrate
	from oracle import nway_oracle as O
	a, b = (150.0, 2.0), (150.01, 2.02)
	for f in (0.1, 0.5, 0.9):
		ra, dec = calibrate.greatarc_interpolate(a, b, f)
		d = O.dist(a, b)
		assert abs(O.dist(a, (ra, dec)) - f * d) < 1e-12 and abs(O.dist((ra, dec), b) - (1 - f) * d) < 1e-12


def test_helper_command_lines_without_a_gpu(tmp_path):
	"""the root scripts are thin wrappers around nway_b200/calibrate_cli.py: argument handling, FITS I/O and the cut-off
	table run on the host (the collision searches of the other two need the device and are covered by the GPU tests)"""
	import os
	import subprocess
	import sys
	from nway_b200 import fitsio
	from tests import cases
	root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
	paths = cases.write_cosmos_subset_fits(str(tmp_path))
	# no shift given: the reference's error message, exit code 1, nothing written
	res = subprocess.run([sys.executable, os.path.join(root, 'nway-create-shifted-catalogue.py'), '--radius', '40', paths['XMM'], str(tmp_path / 'out.fits')],
		capture_output=True, text=True)
	assert res.returncode == 1 and 'ERROR: You have to set either shift-ra or shift-dec to non-zero' in res.stdout, res.stderr[-500:]
	assert not os.path.exists(str(tmp_path / 'out.fits'))
	res = subprocess.run([sys.executable, os.path.join(root, 'nway-create-fake-catalogue.py'), '--help'], capture_output=True, text=True)
	assert res.returncode == 0 and '--seed' in res.stdout and '--radius' in res.stdout
	# cut-off table from two match tables on disk
	rng = np.random.default_rng(3)
	for name, a, b in (('real.fits', 5, 2), ('fake.fits', 1, 6)):
		n = 2000
		fitsio.write_table(str(tmp_path / name), [fitsio.Column('ncat', 'I', np.where(np.arange(n) % 3 == 0, 1, 2)),
			fitsio.Column('p_any', 'E', rng.beta(a, b, n)), fitsio.Column('p_i', 'E', rng.uniform(size=n)), fitsio.Column('match_flag', 'I', np.ones(n))], 'NWAYMATCH')
	res = subprocess.run([sys.executable, os.path.join(root, 'nway-calibrate-cutoff.py'), 'real.fits', 'fake.fits'], capture_output=True, text=True, cwd=str(tmp_path))
	assert res.returncode == 0, res.stderr[-1000:]
	assert 'For a false detection rate of <1%' in res.stdout and 'created table "real.fits_p_any_cutoffquality.txt"' in res.stdout
	tab = np.loadtxt(str(tmp_path / 'real.fits_p_any_cutoffquality.txt'))
	assert tab.shape == (101, 3) and tab[0, 0] == 0 and tab[-1, 0] == 1
