"""GPU: BASELINE.json's full-size synthetic configurations.  The oracle cannot enumerate these in seconds, so the
checks are (a) size-independent properties of the table and (b) the oracle on a random subset of the primaries
against the complete secondary catalogues (a primary's rows depend on no other primary)."""
import numpy as np
import pytest

from tests import cases, parity

pytestmark = pytest.mark.gpu


def check_properties(got, names, n_primary, radius):
	idx = np.stack([got[n] for n in names], axis=1)
	prim = idx[:, 0]
	# rows sorted lexicographically, -1 first (fastskymatch.py:181,217); every primary present exactly once as a group
	order = np.lexsort(tuple(idx[:, c] for c in range(idx.shape[1] - 1, -1, -1)))
	assert (order == np.arange(len(prim))).all(), 'rows are not in the reference order'
	starts = np.concatenate(([0], np.flatnonzero(np.diff(prim) != 0) + 1))
	assert len(starts) == n_primary and (prim[starts] == np.arange(n_primary)).all()
	# first row of each group is the no-counterpart row (__init__.py:435)
	assert (got['ncat'][starts] == 1).all() and (idx[starts, 1:] == -1).all()
	assert (got['ncat'] == (idx > -1).sum(axis=1)).all()
	# separations: strict radius cut, NaN exactly where a member is absent
	assert (got['Separation_max'] < radius).all() and (got['Separation_max'][starts] == 0).all()
	# probabilities
	p_any, p_i = got['prob_has_match'], got['prob_this_match']
	assert np.isfinite(p_any).all() and (p_any >= -1e-13).all() and (p_any <= 1).all()
	size = np.diff(np.concatenate((starts, [len(prim)])))
	sums = np.add.reduceat(p_i, starts)
	assert np.allclose(sums[size > 1], 1.0, rtol=0, atol=1e-12), 'p_i must sum to 1 over the counterparts of a primary'
	assert (sums[size == 1] == 0).all() and (p_any[starts][size == 1] == 0).all()
	assert (p_any == np.repeat(p_any[starts], size)).all(), 'p_any is a per-primary constant'
	flag = got['match_flag']
	best = np.add.reduceat((flag == 1).astype(np.int64), starts)
	assert (best >= 1).all(), 'every group flags its best row'
	assert (flag[starts][size > 1] == 0).all()
	return starts, size


def check_subset_against_oracle(tables, got, names, radius, completeness, pick, context, **kw):
	"""oracle on the picked primaries only, with the densities of the full catalogue"""
	from oracle import nway_oracle as O
	sub = [dict(t) for t in tables]
	for k in ('ra', 'dec', 'error'):
		sub[0][k] = tables[0][k][pick]
	# the prior uses n/area of the FULL primary catalogue: scale the area of the subset accordingly
	sub[0]['area'] = tables[0]['area'] * len(pick) / len(tables[0]['ra'])
	ref = O.nway_match(sub, radius, completeness, **kw)
	rows = np.flatnonzero(np.isin(got[names[0]], pick))
	part = {k: np.asarray(v)[rows] for k, v in got.items() if not k.startswith('_')}
	ref[names[0]] = pick[ref[names[0]]]
	cols = [c for c in ref if not c.startswith('_')]
	# nu_0 = n/area*A is the same number up to the rounding of the rescaled area (1e-16): the north star's 1e-10 holds
	return parity.assert_tables_match(ref, part, columns=cols, context=context)


def test_c3_full_size():
	"""configs[2]: 1e5 x 1e7 on 1 deg^2, r = 5 arcsec -- the bench workload"""
	import nway_b200
	tables = cases.config_c3(1.0)
	got = nway_b200.nway_match(tables, 5.0, 0.9, logger=nway_b200.NullOutputLogger(), as_frame=False)
	assert len(got['A']) == 6149842          # the reference's own row count for this seed
	check_properties(got, ['A', 'B'], 100000, 5.0)
	pick = np.sort(np.random.default_rng(1).choice(100000, 1500, replace=False))
	check_subset_against_oracle(tables, got, ['A', 'B'], 5.0, 0.9, pick, 'C3 subset')


def test_c4_all_sky_three_catalogues():
	"""configs[3] (one GPU's worth): 1e6 x 1e7 x 1e7 uniform on the sphere, r = 10 arcsec"""
	import nway_b200
	tables = cases.allsky(20260302, (1000000, 10000000, 10000000), (1.0, 0.3, 0.5))
	got = nway_b200.nway_match(tables, 10.0, 0.9, logger=nway_b200.NullOutputLogger(), as_frame=False)
	starts, size = check_properties(got, ['A', 'B', 'C'], 1000000, 10.0)
	assert 1.005e6 < len(got['A']) < 1.02e6   # expected ~1.012e6 rows
	pick = np.sort(np.random.default_rng(2).choice(1000000, 3000, replace=False))
	# make sure the subset contains matched primaries
	matched = np.unique(got['A'][got['ncat'] > 1])
	pick = np.unique(np.concatenate((pick, matched[:: max(1, len(matched) // 400)])))
	check_subset_against_oracle(tables, got, ['A', 'B', 'C'], 10.0, 0.9, pick, 'C4 subset')


def test_c5_like_four_catalogues_elliptical_and_magnitude_priors():
	"""configs[4] at 1/30 of the secondaries: all-sky 4-catalogue match, elliptical primary errors, one magnitude
	prior per secondary catalogue (fixed histograms), CLI correction on"""
	import nway_b200
	from oracle import nway_oracle as O
	def build():
		rng = np.random.default_rng(20260303)
		tables = cases.allsky(20260303, (100000, 3000000, 3000000, 3000000), (1.0, 0.3, 0.4, 0.5))
		n0 = 100000
		major = rng.uniform(0.5, 3.0, n0)
		tables[0]['error'] = nway_b200.ellipse_error(major, rng.uniform(0.2, 1.0, n0) * major, rng.uniform(0, 180, n0))
		tables = cases.with_mags(tables, 77, cats=(1, 2, 3), ncols=1)
		for t in tables[1:]:   # no empty hist_sel bin here: a lone counterpart with weight -inf makes the reference's p_i NaN
			lo, hi, hs, ha = t['maghists'][0]   # (covered by the syn2_maghist golden case)
			t['maghists'][0] = (lo, hi, np.where(hs == 0, 0.05, hs), ha)
		return tables
	tables = build()
	got = nway_b200.nway_match(tables, 10.0, 0.9, logger=nway_b200.NullOutputLogger(), as_frame=False, unrelated_mode='cli')
	check_properties(got, ['A', 'B', 'C', 'D'], 100000, 10.0)
	matched = np.unique(got['A'][got['ncat'] > 1])
	pick = np.unique(np.concatenate((np.arange(0, 100000, 997), matched[:: max(1, len(matched) // 300)])))
	sub = build()
	sub[0]['error'] = tuple(x[pick] for x in sub[0]['error'])
	for k in ('ra', 'dec'):
		sub[0][k] = sub[0][k][pick]
	sub[0]['area'] = sub[0]['area'] * len(pick) / 100000
	ref = O.nway_match(sub, 10.0, 0.9, unrelated_mode='cli')
	rows = np.flatnonzero(np.isin(got['A'], pick))
	part = {k: np.asarray(v)[rows] for k, v in got.items() if not k.startswith('_')}
	ref['A'] = pick[ref['A']]
	parity.assert_tables_match(ref, part, columns=[c for c in ref if not c.startswith('_')], context='C5-like subset')


@pytest.mark.parametrize('ra0,dec0', [(100.0, 35.0), (352.0, -20.0), (40.0, 80.0)])
def test_sparse_patch_streams_through_the_filter(ra0, dec0):
	"""few primaries, a long secondary catalogue, NOT the whole sky: the two-kernel stream (k_filter with the shared-memory
	bitmap on a grid that does not span the circle, across ra = 0, near a pole) against the oracle's whole table"""
	import nway_b200
	from nway_b200 import _lib
	from oracle import nway_oracle as O
	tables = cases.uniform_patch(int(ra0) + 7, (2500, 2300000), (1.0, 0.4), 16.0, ra0=ra0, dec0=dec0)
	for t in tables:   # positions beyond 360 deg / 90 deg folded back onto the sphere
		t['ra'] = np.mod(t['ra'], 360.0)
		over = t['dec'] > 90.0
		t['dec'] = np.where(over, 180.0 - t['dec'], t['dec'])
		t['ra'] = np.where(over, np.mod(t['ra'] + 180.0, 360.0), t['ra'])
	got = nway_b200.nway_match(tables, 6.0, 0.9, logger=nway_b200.NullOutputLogger(), as_frame=False)
	check_properties(got, ['A', 'B'], 2500, 6.0)
	ref = O.nway_match(tables, 6.0, 0.9)
	parity.assert_tables_match(ref, got, columns=[c for c in ref if not c.startswith('_')], context='sparse patch at (%g, %g)' % (ra0, dec0))
	# ... and it did take that route: the long catalogue costs launches (bitmap, k_filter, the fallback k_pairs) that the
	# same field with a short catalogue -- streamed by k_pairs alone -- does not
	ctx = _lib.get_context()
	def launches_of(tabs):
		nway_b200.nway_match(tabs, 6.0, 0.9, logger=nway_b200.NullOutputLogger(), as_frame=False)   # sizes the buffers
		nway_b200.nway_match(tabs, 6.0, 0.9, logger=nway_b200.NullOutputLogger(), as_frame=False)
		return ctx.launch_count()   # kernels of the last match
	short = [dict(tables[0]), dict(tables[1])]
	for k in ('ra', 'dec', 'error'):
		short[1][k] = tables[1][k][:200000]
	assert launches_of(tables) >= launches_of(short) + 2

