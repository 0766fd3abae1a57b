"""GPU: seeded random configurations against the oracle -- catalogue counts, field position (equator, mid-latitudes,
pole caps, across ra = 0), field size, radius, densities, duplicated positions, error columns, completeness vectors,
circular / elliptical errors, API / command-line semantics, float32 separations, pair pre-filters.  Every case is
small enough for the oracle's KD-tree enumerator; together they walk the grid code (band grids of very different
shapes, packed cell entries, the occupancy bitmap, spill lists, the one-thread-per-primary kernels)."""
import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu


def random_case(seed):
	rng = np.random.default_rng(seed)
	ncat = int(rng.integers(2, 5))
	kind = rng.choice(['equator', 'mid', 'north', 'south', 'wrap', 'allsky'])
	radius = float(10 ** rng.uniform(0, 3.2))               # 1 arcsec .. 1600 arcsec
	side = float(np.clip(radius / 3600 * rng.uniform(8, 200), 1e-3, 30))   # field side in degrees
	if kind == 'allsky':
		side = 360.0
	n0 = int(rng.integers(1, 400))
	tables = []
	dec0 = dict(equator=0.0, mid=rng.uniform(-70, 70), north=90 - side * rng.uniform(0, 0.6), south=-90 + side * rng.uniform(0, 0.6),
		wrap=rng.uniform(-60, 60), allsky=0.0)[kind]
	ra0 = 360 - side / 3 if kind == 'wrap' else rng.uniform(0, 360)
	for c in range(ncat):
		n = n0 if c == 0 else int(rng.integers(0 if ncat > 2 else 1, 3000))
		if kind == 'allsky':
			ra = 360 * rng.uniform(size=n)
			dec = np.degrees(np.arcsin(2 * rng.uniform(size=n) - 1))
			area = 41252.96124941928
		else:
			dec = dec0 + side * (rng.uniform(size=n) - 0.5)
			dec = np.where(dec > 90, 180 - dec, np.where(dec < -90, -180 - dec, dec))   # over the pole, not piled up on it
			cosd = max(np.cos(np.radians(min(abs(dec0) + side / 2, 89.9))), 0.02)
			ra = (ra0 + side * (rng.uniform(size=n) - 0.5) / (cosd if rng.uniform() < 0.5 else 1.0)) % 360
			area = max(side * side, 1e-6)
		if c > 0 and kind != 'allsky':
			# keep the cartesian product small enough for the oracle: at most ~3 expected neighbours per primary
			true_area = area * (cosd if False else 1.0)
			expect = n / max(true_area, 1e-12) * np.pi * (radius / 3600) ** 2
			if expect > 3:
				keep = max(1, int(n * 3 / expect))
				ra, dec, n = ra[:keep], dec[:keep], keep
		if c > 0 and n > 10 and len(tables[0]['ra']) > 0:
			# plant counterparts near primaries (some exactly on top, some just inside / outside the radius)
			k = rng.integers(0, len(tables[0]['ra']), n // 3)
			off = radius / 3600 * rng.choice([0.0, 0.2, 0.999, 1.001, 0.6], size=len(k)) * rng.uniform(0.9, 1.0, len(k))
			ang = rng.uniform(0, 2 * np.pi, len(k))
			pd = tables[0]['dec'][k]
			d2 = pd + off * np.cos(ang)
			d2 = np.where(d2 > 90, 180 - d2, np.where(d2 < -90, -180 - d2, d2))
			r2 = (tables[0]['ra'][k] + off * np.sin(ang) / np.maximum(np.cos(np.radians(pd)), 1e-3)) % 360
			ra[:len(k)] = r2
			dec[:len(k)] = d2
		# positional errors of a few per cent to a third of the search radius: |log BF| stays below a few hundred, the regime
		# in which a relative 1e-10 on the posteriors is meaningful (d p / p = ln 10 x d log BF, and log BF inherits the
		# conditioning of the reference's own separation formula)
		slo, shi = max(0.05, radius / 30), max(0.1, radius / 3)
		err = rng.uniform(slo, shi, n) if rng.uniform() < 0.5 else float(rng.uniform(slo, shi)) * np.ones(n)
		tables.append(dict(name='ABCD'[c], ra=ra, dec=dec, error=err, area=float(area), mags=[], magnames=[], maghists=[]))
	# keep the cartesian product small enough for the oracle: thin the densest secondary catalogue until the number of
	# candidate tuples (before the secondary-secondary filter) is below ~1e5
	from oracle import nway_oracle as O
	for _ in range(20):
		if len(tables[0]['ra']) == 0:
			break
		lists = O.neighbour_lists([(t['ra'], t['dec']) for t in tables], radius / 3600)
		counts = np.array([[len(js) for js in lc] for lc in lists], dtype=float)   # (ncat - 1, n0)
		if np.prod(counts + 1, axis=0).sum() <= 1e5:
			break
		c = 1 + int(np.argmax(counts.sum(axis=1)))
		t = tables[c]
		keep = np.sort(rng.choice(len(t['ra']), max(1, len(t['ra']) // 2), replace=False))
		t['ra'], t['dec'], t['error'] = t['ra'][keep], t['dec'][keep], t['error'][keep]
	kw = {}
	if ncat > 2 and rng.uniform() < 0.5:
		kw['unrelated_mode'] = 'cli'
		if rng.uniform() < 0.5:
			kw['cli_compat'] = True
	if rng.uniform() < 0.25:
		for t in tables:   # elliptical mode: every catalogue carries a triple
			n = len(t['ra'])
			a = rng.uniform(max(0.05, radius / 30), max(0.1, radius / 3), n)
			from oracle import nway_oracle as O
			t['error'] = tuple(O.convert_from_ellipse(a, a * rng.uniform(0.2, 1, n), rng.uniform(0, np.pi, n)))
	if ncat > 2 and rng.uniform() < 0.3:
		kw['pairwise_errs'] = [(1, 2, float(radius * rng.uniform(0, 0.8)))]
	pc = float(rng.uniform(0.3, 1.0)) if rng.uniform() < 0.6 else np.r_[1.0, rng.uniform(0.3, 1.0, ncat - 1)]
	return tables, radius, pc, kw, kind


@pytest.mark.parametrize('seed', list(range(1000, 1080)))
def test_random_configuration(seed):
	import nway_b200
	from oracle import nway_oracle as O
	tables, radius, pc, kw, kind = random_case(seed)
	try:
		ref = O.nway_match([dict(t) for t in tables], radius, pc, **kw)
	except ValueError:
		ref = None
	copy = [dict(t) for t in tables]
	got = nway_b200.nway_match(copy, radius, pc, logger=nway_b200.NullOutputLogger(), store_mag_hists=False, as_frame=False, **kw)
	assert ref is not None
	flat = O.flat_sky_applicable([(t['ra'], t['dec']) for t in tables], radius / 60. / 60)
	assert nway_b200._lib.get_context().flat_hash_applied() == flat   # the device takes the reference's decision (fastskymatch.py:94-98)
	if flat and seed % 2:
		# ... and with the switch off it returns the complete enumeration (odd seeds: both are checked)
		full = O.nway_match([dict(t) for t in tables], radius, pc, enumerator='complete', **kw)
		got2 = nway_b200.nway_match([dict(t) for t in tables], radius, pc, logger=nway_b200.NullOutputLogger(), store_mag_hists=False, as_frame=False,
			flat_hash_compat=False, **kw)
		parity.assert_tables_match(full, got2, columns=[c for c in full if not c.startswith('_')], context='fuzz seed %d, complete' % seed,
			rtol=5e-7 if kw.get('cli_compat') else None)
	cols = [c for c in ref if not c.startswith('_')]
	# The north star's tolerance (tests/parity.py: 1e-10 relative) everywhere the arithmetic is fp64: sin / cos carry the
	# reference's bits (sin_ref / cos_ref), so the cancelling part of the separation formula (fastskymatch.py:44) is
	# reproduced exactly, poles and sub-arcsecond separations included.
	rtol = None
	if kw.get('cli_compat'):
		# separations are rounded to float32 before they are scored: the few-ulp difference that hypot / atan2 leave in the
		# fp64 value moves one float32 rounding in ~1e7, and that one then differs by a float32 ulp
		rtol = 5e-7
	parity.assert_tables_match(ref, got, columns=cols, context='fuzz seed %d (%s, r=%.3g, ncat=%d, %s)' % (seed, kind, radius, len(tables), sorted(kw)), rtol=rtol)
