"""Shared, seeded input builders for the parity tests, the golden generator and bench.py.

Every catalogue is a dict in the reference's own format (nwaylib/__init__.py:38-48):
name, ra, dec (deg), error (arcsec), area (deg^2), mags, magnames, maghists.
"""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _cat(name, ra, dec, error, area, mags=(), magnames=(), maghists=()):
	return dict(name=name, ra=ra, dec=dec, error=error, area=float(area),
		mags=list(mags), magnames=list(magnames), maghists=list(maghists))


def cosmos_subset(ncat=2, mags=False):
	"""the reference's demo field (doc/COSMOS_{XMM,OPTICAL,IRAC}.fits) cut to what a 20 arcsec match can
	reach; see oracle/make_golden.py.  mags=True attaches MAG / mag_ch1 with maghists=None ('auto')."""
	z = np.load(os.path.join(GOLDEN_DIR, 'cosmos_subset.npz'))
	out = []
	for name, magname in (('XMM', None), ('OPT', 'MAG'), ('IRAC', 'mag_ch1'))[:ncat]:
		t = _cat(name, z[name + '_ra'].copy(), z[name + '_dec'].copy(), z[name + '_error'].copy(), 2.0)
		if mags and magname is not None:
			t['mags'] = [z[name + '_mag'].copy()]
			t['magnames'] = [magname]
			t['maghists'] = [None]
		out.append(t)
	return out


def uniform_patch(seed, counts, sigmas, side_deg, ra0=150.0, dec0=0.0, names='ABCDEFGH'):
	"""uniform catalogues on a small square centred on (ra0+side/2, dec0): the flat-sky regime in which the
	reference's own hash is complete, so its row set is comparable 1:1."""
	rng = np.random.default_rng(seed)
	out = []
	for c, (n, s) in enumerate(zip(counts, sigmas)):
		ra = ra0 + side_deg * rng.uniform(size=n)
		dec = dec0 - side_deg / 2 + side_deg * rng.uniform(size=n)
		out.append(_cat(names[c], ra, dec, s * np.ones(n), side_deg * side_deg))
	return out


def allsky(seed, counts, sigmas, names='ABCDEFGH'):
	"""uniform on the sphere."""
	rng = np.random.default_rng(seed)
	out = []
	for c, (n, s) in enumerate(zip(counts, sigmas)):
		ra = 360 * rng.uniform(size=n)
		dec = np.degrees(np.arcsin(2 * rng.uniform(size=n) - 1))
		out.append(_cat(names[c], ra, dec, s * np.ones(n), 41252.96124941928))
	return out


def allsky_hard(seed, counts, sigmas, radius_arcsec, ncluster=250):
	"""all-sky catalogues with the difficult primaries put in by hand -- the two poles and their surroundings, both sides
	of ra = 0 / 360, the corners where HEALPix base pixels meet -- and secondaries clustered around the primaries, so
	that the match has real groups everywhere the sphere is awkward.  Takes the HEALPix branch of the reference's
	crossproduct (fastskymatch.py:134-160)."""
	rng = np.random.default_rng(seed)
	tables = allsky(seed + 1, counts, sigmas)
	p = tables[0]
	hard_ra = [0.0, 359.99999, 0.00001, 123.0, 250.0, 180.0, 45.0, 135.0, 0.0, 90.0, 44.9999, 315.0, 90.0001]
	hard_dec = [0.0, 10.0, -45.0, 89.9999, -89.99995, 90.0, 41.8103149, -41.8103149, 41.81, 0.0, 0.0001, 89.99, -90.0]
	nh = len(hard_ra)
	p['ra'][:nh] = hard_ra
	p['dec'][:nh] = hard_dec
	n0 = len(p['ra'])
	for t in tables[1:]:
		n = min(ncluster * 4, len(t['ra']))
		k = np.concatenate((rng.integers(0, nh, n // 2), rng.integers(0, n0, n - n // 2)))
		rad = np.radians(rng.uniform(0, 1.3 * radius_arcsec / 3600.0, n))
		ang = rng.uniform(0, 2 * np.pi, n)
		dec = p['dec'][k] + np.degrees(rad * np.cos(ang))
		over, under = dec > 90, dec < -90
		dec[over] = 180 - dec[over]
		dec[under] = -180 - dec[under]
		cosd = np.maximum(np.cos(np.radians(p['dec'][k])), 1e-6)
		ra = (p['ra'][k] + np.degrees(rad * np.sin(ang)) / cosd + np.where(over | under, 180, 0)) % 360
		t['ra'][:n] = ra
		t['dec'][:n] = dec
	return tables


def config_c3(scale=1.0, seed=20260301):
	"""BASELINE.json configs[2]: 1e5 x 1e7 uniform on 1 deg^2, sigma 1.0 / 0.2 arcsec, r = 5 arcsec,
	completeness 0.9 (generator of SURVEY.md 8d).  scale < 1 shrinks the AREA at fixed surface density, so a
	sample has the same rows per primary as the full workload."""
	rng = np.random.default_rng(seed)
	side = np.sqrt(scale)
	n0, n1 = int(round(1e5 * scale)), int(round(1e7 * scale))
	prim = _cat('A', 150 + side * rng.uniform(size=n0), -side / 2 + side * rng.uniform(size=n0), 1.0 * np.ones(n0), scale)
	sec = _cat('B', 150 + side * rng.uniform(size=n1), -side / 2 + side * rng.uniform(size=n1), 0.2 * np.ones(n1), scale)
	return [prim, sec]


def fixed_hist(seed, lo=16.0, hi=28.0, nbins=16):
	"""a deterministic user-supplied magnitude prior (bins_lo, bins_hi, hist_sel, hist_all), with one empty
	hist_all bin (-> ratio 100, magnitudeweights.py:23) and one empty hist_sel bin (-> weight -inf)."""
	rng = np.random.default_rng(seed)
	edges = np.linspace(lo, hi, nbins + 1)
	hs = rng.uniform(0.01, 0.2, nbins)
	ha = rng.uniform(0.01, 0.2, nbins)
	ha[3] = 0.0
	hs[nbins - 2] = 0.0
	return edges[:-1], edges[1:], hs, ha


def with_mags(tables, seed, cats=(1,), ncols=1, hist=True):
	"""attach ~N(22,2) magnitudes with 1% -99 to the given catalogues."""
	rng = np.random.default_rng(seed)
	for c in cats:
		t = tables[c]
		n = len(t['ra'])
		for k in range(ncols):
			m = rng.normal(22, 2, n)
			m[rng.uniform(size=n) < 0.01] = -99
			t['mags'].append(m)
			t['magnames'].append('m%d' % k)
			t['maghists'].append(fixed_hist(seed + 17 * c + k) if hist else None)
	return tables


GOLDEN_CASES = {
	# name: how to build + how to call.  Outputs of the REAL reference are stored in golden/ref_<name>.npz
	'cosmos2': dict(radius=20, completeness=0.9),
	'cosmos2_magradius': dict(radius=20, completeness=0.9, kwargs=dict(mag_include_radius=4.0)),
	'cosmos3': dict(radius=20, completeness=0.9, stride=211),
	'cosmos3_magauto': dict(radius=20, completeness=0.9, stride=211),
	'syn2': dict(radius=5, completeness=0.9, stride=97),
	'syn2_sparse': dict(radius=5, completeness=0.7, stride=3),
	'syn2_maghist': dict(radius=5, completeness=0.9, stride=97),
	'syn3': dict(radius=8, completeness=0.9, stride=151),
	'syn3_pcvec': dict(radius=8, completeness=np.array([1.0, 0.8, 0.6]), stride=151, kwargs=dict(prob_ratio_secondary=0.1)),
	'syn4': dict(radius=6, completeness=0.95, stride=199),
	'syn4_minprob': dict(radius=6, completeness=0.95, stride=23, kwargs=dict(min_prob=0.01)),
	# flat-sky fields AWAY from the equator: the reference's hash bins ra without cos(dec) and leaves out part of the pairs
	# within the radius.  NWB_COMPAT_FLAT_HASH (the default of nway_b200.nway_match) returns exactly these
	# rows; the oracle applies the hash's bucket predicate (enumerator='reference', its default)
	'offeq2': dict(radius=6, completeness=0.9, stride=31),
	'offeq3': dict(radius=8, completeness=0.85, stride=53),
	'offeq3_south': dict(radius=10, completeness=np.array([1.0, 0.9, 0.7]), stride=41),
	# whole sphere: the reference takes its HEALPix branch (healpy restated in oracle/healpix_nest.py).  At |dec| ~ 90 the
	# reference's separation formula (fastskymatch.py:44) cancels to an absolute 1e-16 rad, which a 1-ulp difference in
	# sin / cos turns into ~1e-10 of the posteriors: the device evaluates them with the reference's bits (sin_ref /
	# cos_ref, nwb_device.cuh), so these hold the north star's 1e-10 like every other case
	'allsky2': dict(radius=120, completeness=0.9, stride=7),
	'allsky3': dict(radius=300, completeness=0.8, stride=7),
}


def build_case(name):
	if name == 'cosmos2':
		return cosmos_subset(2)
	if name == 'cosmos2_magradius':
		return cosmos_subset(2, mags=True)
	if name == 'cosmos3':
		return cosmos_subset(3)
	if name == 'cosmos3_magauto':
		return cosmos_subset(3, mags=True)
	if name == 'syn2':
		return config_c3(scale=0.02)
	if name == 'syn2_sparse':
		return uniform_patch(11, (1000, 700), (2.0, 1.0), 1.0)
	if name == 'syn2_maghist':
		return with_mags(config_c3(scale=0.01, seed=7), 99, cats=(1,), ncols=2)
	if name in ('syn3', 'syn3_pcvec'):
		return uniform_patch(5, (500, 20000, 15000), (1.0, 0.3, 0.5), 0.1)
	if name in ('syn4', 'syn4_minprob'):
		return with_mags(uniform_patch(6, (200, 3000, 3000, 2500), (1.0, 0.4, 0.5, 0.8), 0.05), 3, cats=(2,), ncols=1)
	if name == 'offeq2':
		return uniform_patch(21, (800, 60000), (1.0, 0.3), 0.2, ra0=210.0, dec0=38.0)
	if name == 'offeq3':
		return uniform_patch(22, (400, 15000, 12000), (1.0, 0.3, 0.5), 0.1, ra0=40.0, dec0=-41.0)
	if name == 'offeq3_south':
		return with_mags(uniform_patch(23, (300, 9000, 7000), (1.5, 0.4, 0.6), 0.1, ra0=300.0, dec0=-22.0), 8, cats=(1,))
	if name == 'allsky2':
		return allsky_hard(31, (600, 6000), (4.0, 2.0), 120.0)
	if name == 'allsky3':
		return allsky_hard(32, (400, 5000, 4000), (5.0, 3.0, 4.0), 300.0, ncluster=90)
	raise KeyError(name)


# ---- command-line cases: FITS files of the COSMOS subset + argument lists (oracle/make_golden_cli.py) ----------

def write_cosmos_subset_fits(directory):
	"""the COSMOS subset as three FITS catalogues with the columns of the reference's demo files
	(doc/COSMOS_{XMM,OPTICAL,IRAC}.fits: ID, RA, DEC + pos_err / MAG / mag_ch1; SKYAREA = 2 deg^2).  Written with
	the product's FITS writer; the reference run that produced the goldens read them with its own reader."""
	from nway_b200 import fitsio
	z = np.load(os.path.join(GOLDEN_DIR, 'cosmos_subset.npz'))
	paths = {}
	for name, extra, key in (('XMM', 'pos_err', 'error'), ('OPT', 'MAG', 'mag'), ('IRAC', 'mag_ch1', 'mag')):
		n = len(z[name + '_ra'])
		cols = [fitsio.Column('ID', 'J', np.arange(1, n + 1)), fitsio.Column('RA', 'D', z[name + '_ra']),
			fitsio.Column('DEC', 'D', z[name + '_dec']), fitsio.Column(extra, 'E', z['%s_%s' % (name, key)])]
		paths[name] = os.path.join(directory, 'SUB_%s.fits' % name)
		fitsio.write_table(paths[name], cols, name, table_header=[('SKYAREA', 2.0)])
	return paths


CLI_CASES = {
	# name: arguments after the catalogue list is substituted ({XMM}, {OPT}, {IRAC} = file paths)
	'cli2': ['--radius', '20', '--prior-completeness', '0.9', '{XMM}', ':pos_err', '{OPT}', '0.1'],
	'cli3': ['--radius', '20', '{XMM}', ':pos_err', '{OPT}', '0.1', '{IRAC}', '0.5'],
	'cli3_magauto': ['--radius', '20', '{XMM}', ':pos_err', '{OPT}', '0.1', '{IRAC}', '0.5',
		'--mag', 'OPT:MAG', 'auto', '--mag', 'IRAC:mag_ch1', 'auto', '--mag-radius', '4'],
	'cli3_bayes': ['--radius', '15', '--prior-completeness', '0.9:0.8', '--acceptable-prob', '0.3', '{XMM}', ':pos_err', '{OPT}', '0.1',
		'{IRAC}', '0.5', '--mag', 'OPT:MAG', 'auto', '--mag-auto-minprob', '0.8'],
	'cli3_minprob': ['--radius', '12', '--prior-completeness', '0.95', '--min-prob', '0.01', '--ignore-unrelated-associations',
		'{XMM}', ':pos_err', '{OPT}', '0.1', '{IRAC}', '0.5'],
	'cli3_prefilter': ['--radius', '12', '{XMM}', ':pos_err', '{OPT}', '0.1', '{IRAC}', '0.5', '--prefilter-pair', 'OPT', 'IRAC', '0.5'],
}


def cli_args(name, paths, outfile):
	return [a.format(**paths) for a in CLI_CASES[name]] + ['--out=' + outfile]
