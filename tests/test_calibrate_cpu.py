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
