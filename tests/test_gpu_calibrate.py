"""GPU tests of the calibration loop (nway_b200/calibrate.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shifted_catalogue_reference_log(tmp_path):
	"""doc/logs/XMM-shift:4-5 of the reference: `--radius 40 --shift-ra 60 COSMOS_XMM.fits` removes 561 sources,
	1236 remain; and against the oracle's dist() source by source"""
	from nway_b200 import fitsio
	from oracle import nway_oracle as O
	paths = cases.write_cosmos_subset_fits(str(tmp_path))   # the XMM catalogue is complete in the subset
	out = str(tmp_path / 'XMM-shift.fits')
	res = subprocess.run([sys.executable, os.path.join(ROOT, 'nway-create-shifted-catalogue.py'), '--radius', '40', '--shift-ra', '60', paths['XMM'], out],
		capture_output=True, text=True)
	assert res.returncode == 0, res.stdout + res.stderr
	assert 'removed 561 sources which collide with original positions' in res.stdout
	assert 'writing "%s" (1236 rows)' % out in res.stdout
	t0, t1 = fitsio.read_table(paths['XMM']), fitsio.read_table(out)
	assert len(t1) == 1236 and t1.columns == t0.columns and t1.formats == t0.formats and t1.header['SKYAREA'] == 2.0
	ra, dec = t0.data['RA'], t0.data['DEC']
	ra2 = ra + 60 / 60. / 60
	ex = np.array([(O.dist((a, b), (ra, dec)) * 60 * 60 < 40).any() for a, b in zip(ra2, dec)])
	assert (t1.data['ID'] == t0.data['ID'][~ex]).all()
	assert (t1.data['RA'] == ra2[~ex]).all() and (t1.data['DEC'] == dec[~ex]).all()


def test_pairs_within_against_oracle():
	from nway_b200 import calibrate
	from oracle import nway_oracle as O
	rng = np.random.default_rng(3)
	ra1, dec1 = 10 + rng.uniform(size=500) * 0.2, -30 + rng.uniform(size=500) * 0.2
	ra2, dec2 = 10 + rng.uniform(size=4000) * 0.2, -30 + rng.uniform(size=4000) * 0.2
	i, j, s = calibrate.pairs_within(ra1, dec1, ra2, dec2, 30.0)
	d = O.dist((ra1[:, None], dec1[:, None]), (ra2[None, :], dec2[None, :])) * 3600
	ii, jj = np.nonzero(d < 30.0)
	assert len(i) == len(ii) and (i == ii).all() and (j == jj).all()
	assert np.allclose(s, d[ii, jj], rtol=1e-10, atol=1e-9)


def test_fake_catalogue_properties(tmp_path):
	"""same size, every fake source at least `radius` from every original and from every other fake source, and still
	inside the footprint (a point between two catalogue sources)"""
	from nway_b200 import calibrate, fitsio
	from oracle import nway_oracle as O
	paths = cases.write_cosmos_subset_fits(str(tmp_path))
	out = str(tmp_path / 'XMM-fake.fits')
	res = subprocess.run([sys.executable, os.path.join(ROOT, 'nway-create-fake-catalogue.py'), '--radius', '40', '--seed', '7', paths['XMM'], out],
		capture_output=True, text=True)
	assert res.returncode == 0, res.stdout + res.stderr
	t0, t1 = fitsio.read_table(paths['XMM']), fitsio.read_table(out)
	assert len(t1) == len(t0) and (t1.data['ID'] == t0.data['ID']).all() and (t1.data['pos_err'] == t0.data['pos_err']).all()
	ra, dec, fra, fdec = t0.data['RA'], t0.data['DEC'], t1.data['RA'], t1.data['DEC']
	d_orig = O.dist((fra[:, None], fdec[:, None]), (ra[None, :], dec[None, :])) * 3600
	assert d_orig.min() >= 40.0
	d_self = O.dist((fra[:, None], fdec[:, None]), (fra[None, :], fdec[None, :])) * 3600
	np.fill_diagonal(d_self, 1e9)
	assert d_self.min() >= 40.0
	assert fra.min() >= ra.min() and fra.max() <= ra.max() and fdec.min() >= dec.min() and fdec.max() <= dec.max()
	assert (np.abs(fra - ra) + np.abs(fdec - dec) > 0).all()


def test_calibration_loop_end_to_end(tmp_path):
	"""real match, fake catalogue, fake match, cut-off table: the workflow of doc/logs/cutoff2 on the COSMOS subset"""
	from nway_b200 import fitsio
	paths = cases.write_cosmos_subset_fits(str(tmp_path))
	run = lambda *a: subprocess.run([sys.executable] + list(a), cwd=str(tmp_path), capture_output=True, text=True)
	r = run(os.path.join(ROOT, 'nway.py'), '--radius', '20', '--prior-completeness', '0.9', paths['XMM'], ':pos_err', paths['OPT'], '0.1', '--out', 'real.fits')
	assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
	r = run(os.path.join(ROOT, 'nway-create-fake-catalogue.py'), '--radius', '40', '--seed', '1', paths['XMM'], 'XMM-fake.fits')
	assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
	r = run(os.path.join(ROOT, 'nway.py'), '--radius', '20', '--prior-completeness', '0.9', 'XMM-fake.fits', ':pos_err', paths['OPT'], '0.1', '--out', 'fake.fits')
	assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
	r = run(os.path.join(ROOT, 'nway-calibrate-cutoff.py'), 'real.fits', 'fake.fits')
	assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
	assert 'For a false detection rate of <10%' in r.stdout or 'A false detection rate of <10% is not possible.' in r.stdout
	tab = np.loadtxt(str(tmp_path / 'real.fits_p_any_cutoffquality.txt'))
	assert tab.shape == (101, 3) and (np.diff(tab[:, 1]) <= 0).all() and (np.diff(tab[:, 2]) <= 0).all()
	real, fake = fitsio.read_table(str(tmp_path / 'real.fits')).data, fitsio.read_table(str(tmp_path / 'fake.fits')).data
	# real counterparts exist, fake positions have none: the real catalogue must look better at every cut-off
	m0, m1 = real['ncat'] == 1, fake['ncat'] == 1
	assert real['p_any'][m0].mean() > fake['p_any'][m1].mean() + 0.2
