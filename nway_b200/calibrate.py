"""The false-association calibration loop around the match path: the three helper programs every
real nway run is followed by --

  nway-create-shifted-catalogue.py   shift a catalogue, drop sources that land near an original one
  nway-create-fake-catalogue.py      a random-position twin of a catalogue with the same local structure
  nway-calibrate-cutoff.py           p_any cut-off for a given false-detection rate, from a real and a fake match

The reference does the neighbour searches with one vectorised dist() per source (O(N^2), nway-create-shifted-
catalogue.py:69-75, nway-create-fake-catalogue.py:137-186).  Here every "is anything within r of this position"
question is one pass of the GPU match path (k_pairs): the positions to test are the primary catalogue, the
positions to avoid the secondary one.  The bookkeeping around it (random draws, great-arc interpolation of O(N)
points, the 101-point cut-off table) is host numpy, as in the reference.
"""

import numpy
from numpy import arccos, arctan2, cos, pi, sin, sqrt

from .logger import NullOutputLogger


def pairs_within(ra1, dec1, ra2, dec2, radius_arcsec, device=None):
	"""all pairs (i of catalogue 1, j of catalogue 2) closer than radius_arcsec (strict, fastskymatch.dist
	arithmetic): arrays (i, j, separation in arcsec), ordered by i then j.  One 2-catalogue run of the match path."""
	from . import nway_match
	n1, n2 = len(ra1), len(ra2)
	if n1 == 0 or n2 == 0:
		return numpy.zeros(0, dtype=numpy.int64), numpy.zeros(0, dtype=numpy.int64), numpy.zeros(0)
	ra2, dec2 = numpy.asarray(ra2, dtype=float), numpy.asarray(dec2, dtype=float)
	# a rough sky area of the catalogue to avoid (it only sizes the per-source match buffers; overflow is handled)
	area = max((dec2.max() - dec2.min()) * min(ra2.max() - ra2.min(), 360.0) * max(cos(numpy.radians(numpy.abs(dec2).min())), 1e-3), 1e-6)
	tables = [
		dict(name='A', ra=numpy.asarray(ra1, dtype=float), dec=numpy.asarray(dec1, dtype=float), error=numpy.ones(n1), area=area, mags=[], magnames=[], maghists=[]),
		dict(name='B', ra=ra2, dec=dec2, error=numpy.ones(n2), area=area, mags=[], magnames=[], maghists=[]),
	]
	cols = nway_match(tables, radius_arcsec, 1.0, logger=NullOutputLogger(), store_mag_hists=False, as_frame=False, device=device,
		flat_hash_compat=False)   # a neighbour search: every pair within the radius, not the reference hash's subset
	has = cols['B'] >= 0
	return cols['A'][has], cols['B'][has], cols['Separation_A_B'][has]


def collides(ra, dec, ra_avoid, dec_avoid, radius_arcsec, device=None):
	"""boolean per (ra, dec): some (ra_avoid, dec_avoid) lies within radius_arcsec"""
	i, j, s = pairs_within(ra, dec, ra_avoid, dec_avoid, radius_arcsec, device=device)
	out = numpy.zeros(len(ra), dtype=bool)
	out[i] = True
	return out


def shifted_catalogue(ra, dec, shift_ra_arcsec, shift_dec_arcsec, radius_arcsec, device=None):
	"""nway-create-shifted-catalogue.py:66-76: (ra + shift, dec + shift, excluded) where excluded marks the shifted
	sources that collide with an original position"""
	ra = numpy.asarray(ra, dtype=float)
	dec = numpy.asarray(dec, dtype=float)
	ra_new = ra + shift_ra_arcsec / 60. / 60
	dec_new = dec + shift_dec_arcsec / 60. / 60
	return ra_new, dec_new, collides(ra_new, dec_new, ra, dec, radius_arcsec, device=device)


def greatarc_interpolate(posa, posb, f):
	"""the point a fraction f along the great arc from a to b (nway-create-fake-catalogue.py:103-122)"""
	(a_ra, a_dec), (b_ra, b_dec) = posa, posb
	lon1 = a_ra / 180 * pi
	lat1 = a_dec / 180 * pi
	lon2 = b_ra / 180 * pi
	lat2 = b_dec / 180 * pi
	d = arccos(numpy.clip(sin(lat1) * sin(lat2) + cos(lat1) * cos(lat2) * cos(lon1 - lon2), -1, 1))
	A = sin((1 - f) * d) / sin(d)
	B = sin(f * d) / sin(d)
	x = A * cos(lat1) * cos(lon1) + B * cos(lat2) * cos(lon2)
	y = A * cos(lat1) * sin(lon1) + B * cos(lat2) * sin(lon2)
	z = A * sin(lat1) + B * sin(lat2)
	lat_f = arctan2(z, sqrt(x**2 + y**2))
	lon_f = arctan2(y, x)
	return lon_f * 180 / pi, lat_f * 180 / pi


def fake_catalogue(ra, dec, radius_arcsec, seed=0, device=None, max_rounds=200, logger=None):
	"""A random-position twin of the catalogue (nway-create-fake-catalogue.py:124-186): every source moves to a random
	point on the great arc towards one of its nearest neighbours farther than `radius` (with probability 2/3 one of the
	10 nearest, else one of the 100 nearest), at least `radius` away from both ends, from every original source and
	from every fake source placed so far.

	The reference places the sources one after the other with one O(N) dist() per trial.  Here all still-unplaced
	sources draw a trial position at once and one GPU pass answers the collision questions; sources whose trial
	collides (with an original, or with a fake source of an earlier round / an earlier index of this round) draw again
	in the next round.  The output is random in both programs; the guarantees are the same."""
	ra = numpy.asarray(ra, dtype=float)
	dec = numpy.asarray(dec, dtype=float)
	n = len(ra)
	rng = numpy.random.RandomState(seed) if seed > 0 else numpy.random
	log = logger.log if logger is not None else (lambda *a: None)
	# neighbour lists: grow the search radius until (almost) every source has enough neighbours beyond `radius`
	# (a small or sparse catalogue cannot give every source 10 neighbours: min(10, n - 1) is what can be asked for, and the
	# search stops at the library's largest radius, just under 30 degrees -- nwb_set_params)
	search = max(4 * radius_arcsec, 60.0)
	search_max = 0.999 * 30 * 3600.0
	want = max(1, min(10, n - 1))
	for _ in range(24):
		search = min(search, search_max)
		i, j, s = pairs_within(ra, dec, ra, dec, search, device=device)
		far = s > radius_arcsec   # excludes the source itself (s = 0) and anything too close to interpolate towards
		counts = numpy.bincount(i[far], minlength=n)
		if (counts >= want).mean() > 0.98 and (counts >= 1).all() or search >= search_max:
			break
		search *= 2
	assert (counts >= 1).all(), 'Method failed: No sources found nearby, could not interpolate a fake source.'
	i, j, s = i[far], j[far], s[far]
	order = numpy.lexsort((s, i))   # per source, neighbours by increasing separation
	i, j, s = i[order], j[order], s[order]
	start = numpy.concatenate(([0], numpy.cumsum(counts)[:-1]))
	log('neighbour lists within %.0f arcsec: %d pairs' % (search, len(i)))

	ra_out, dec_out = ra.copy(), dec.copy()
	todo = numpy.arange(n)
	placed_ra, placed_dec = numpy.zeros(0), numpy.zeros(0)
	for rnd in range(max_rounds):
		if len(todo) == 0:
			break
		m = len(todo)
		c = numpy.minimum(counts[todo], 100)
		wide = rng.randint(0, 3, size=m) == 0
		limit = numpy.where(wide, c, numpy.minimum(c, 10))
		pick = (rng.uniform(size=m) * limit).astype(int)
		k = start[todo] + pick
		di = s[k] / 3600.
		uex = radius_arcsec / 3600. / di
		u = uex + rng.uniform(size=m) * (1 - 2 * uex)
		ok = uex < 0.5   # the neighbour must be more than 2 radii away for a point radius away from both ends
		ra_t, dec_t = greatarc_interpolate((ra[todo], dec[todo]), (ra[j[k]], dec[j[k]]), u)
		ra_t = numpy.fmod(ra_t + 360, 360)
		ok &= ~collides(ra_t, dec_t, ra, dec, radius_arcsec, device=device)
		if len(placed_ra):
			ok &= ~collides(ra_t, dec_t, placed_ra, placed_dec, radius_arcsec, device=device)
		# among this round's trials: a trial loses against any surviving trial of lower index that it collides with
		a, b, _ = pairs_within(ra_t, dec_t, ra_t, dec_t, radius_arcsec, device=device)
		clash = (a > b) & ok[a] & ok[b]
		lose = numpy.zeros(m, dtype=bool)
		lose[a[clash]] = True
		ok &= ~lose
		ra_out[todo[ok]] = ra_t[ok]
		dec_out[todo[ok]] = dec_t[ok]
		placed_ra = numpy.concatenate((placed_ra, ra_t[ok]))
		placed_dec = numpy.concatenate((placed_dec, dec_t[ok]))
		log('round %d: placed %d of %d' % (rnd, ok.sum(), m))
		todo = todo[~ok]
	assert len(todo) == 0, 'Method failed: %d sources could not be placed' % len(todo)
	return ra_out, dec_out


def calibrate_cutoff(real, fake, rates=(0.01, 0.03, 0.05, 0.1)):
	"""nway-calibrate-cutoff.py:41-94 without the plots.  real / fake: mappings with the columns ncat, p_any of a real
	match and of the match of the fake catalogue.  Returns (cutoffs, efficiency, error_rate, lines) where lines is the
	text the reference prints."""
	p_any0 = numpy.asarray(real['p_any'])[numpy.asarray(real['ncat']) == 1]
	p_any0_offset = numpy.asarray(fake['p_any'])[numpy.asarray(fake['ncat']) == 1]
	cutoffs = numpy.linspace(0, 1, 101)
	efficiency = numpy.array([(p_any0 > cutoff).mean() for cutoff in cutoffs])
	error_rate = numpy.array([(p_any0_offset > cutoff).mean() for cutoff in cutoffs])
	lines = []
	for rate in rates:
		lines.append('')
		mask = error_rate < rate
		if not mask.any():
			lines.append('A false detection rate of <%d%% is not possible.' % (rate * 100))
		else:
			i = numpy.min(numpy.where(mask)[0])
			lines.append('For a false detection rate of <%d%%' % (rate * 100))
			lines.append('--> use only counterparts with p_any>%.2f (%.2f%% of matches)' % (cutoffs[i], efficiency[i] * 100))
	return cutoffs, efficiency, error_rate, lines
