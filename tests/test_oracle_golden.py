"""CPU: the oracle (oracle/nway_oracle.py) against the committed outputs of the REAL reference."""

import numpy as np
import pytest

from oracle import nway_oracle as O
from tests import cases, parity

GOLDEN = cases.GOLDEN_DIR


load = parity.load_golden
check_against_digest = parity.check_against_digest


@pytest.mark.parametrize('name', list(cases.GOLDEN_CASES))
def test_oracle_reproduces_reference(name):
	spec = cases.GOLDEN_CASES[name]
	tables = cases.build_case(name)
	got = O.nway_match(tables, spec['radius'], spec['completeness'], enumerator=spec.get('enumerator', 'reference'), **spec.get('kwargs', {}))
	check_against_digest(name, got, [t['name'] for t in tables])


def test_reference_golden_row_counts():
	"""nway-apitest.py:66,109 and doc/logs/match2:30, match3:32."""
	assert int(load('ref_cosmos2.npz')['nrows']) == 37836
	assert int(load('ref_cosmos3.npz')['nrows']) == 387601
	assert int(load('ref_cosmos3_magauto.npz')['nrows']) == 387601


def test_refhash_equals_complete_enumeration_near_equator():
	"""the reference's own flat hash (restated) and the complete enumerator give the same rows where the
	flat hash is complete."""
	tables = cases.uniform_patch(21, (200, 6000, 5000), (1.0, 0.3, 0.5), 0.08)
	a = O.create_match_table(tables, 7.0, 'refhash')
	b = O.create_match_table(tables, 7.0, 'complete')
	assert a['idx'].shape == b['idx'].shape and (a['idx'] == b['idx']).all()


def test_kat_dist_logbf_posterior_elliptical():
	k = load('kat.npz')
	d = O.dist((k['dist_ra1'], k['dist_dec1']), (k['dist_ra2'], k['dist_dec2']))
	assert np.allclose(d, k['dist_out'], rtol=1e-14, atol=0)
	# SURVEY.md Appendix C literals (produced by the reference)
	assert abs(d[0] - 0.002098457623965017) < 1e-17
	assert abs(d[1] * 3600 - 5.089617584392069) < 1e-13
	for n in (1, 2, 3, 4):
		s, p = k['logbf%d_s' % n], k['logbf%d_p' % n]
		out = O.log_bf([[p[i][j] for j in range(n)] for i in range(n)], list(s))
		assert np.allclose(out, k['logbf%d_out' % n], rtol=1e-14, atol=1e-14)
	assert np.allclose(O.posterior(k['post_prior'], k['post_logbf']), k['post_out'], rtol=1e-14, atol=0)
	conv = [O.ellipse_from_cli(a, b, ang) for a, b, ang in k['ell_in']]
	assert np.allclose(np.array(conv), k['ell_conv'], rtol=1e-15, atol=0)
	out = O.log_bf_elliptical(k['ell_sra'], k['ell_sdec'], conv)
	assert np.allclose(out, k['ell_out'], rtol=1e-13, atol=1e-13)


def test_kat_literals_appendix_c():
	lb = lambda psi: float(O.log_bf([[None, np.array([psi])]], [np.array([0.1]), np.array([0.2])])[0])
	assert abs(lb(0.0) - 12.23091025768088) < 1e-13
	assert abs(lb(0.3) - 11.840045223967955) < 1e-13
	assert abs(lb(5.0) - -96.34271021813204) < 1e-12
	p3 = [[None, np.array([0.3]), np.array([0.3])], [None, None, np.array([0.3])], [None, None, None]]
	assert abs(float(O.log_bf(p3, [np.array([0.1]), np.array([0.2]), np.array([0.3])])[0]) - 23.61118582441539) < 1e-13
	assert float(O.log_bf([[None]], [np.array([0.7])])[0]) == 0.0
	assert abs(float(O.posterior(1e-5, 6.0)) - 0.9090917355379414) < 1e-15
	sx, sy, rho = O.convert_from_ellipse(2, 0.5, (30 - 90) * np.pi / 180)
	assert abs(sx - 1.7499999999999998) < 1e-15 and abs(sy - 1.0897247358851685) < 1e-15 and abs(rho - -0.8514850866846712) < 1e-15
	e = O.log_bf_elliptical([[None, np.array([0.8])]], [[None, np.array([-0.6])]],
		[(sx, sy, rho), (np.array([0.1]), np.array([0.1]), np.array([0.]))])
	assert abs(float(e[0]) - 10.44312181435299) < 1e-13


def test_cli_correction_reproduces_recorded_explain_log():
	"""doc/logs/explain:3-19 (CLI): XMM ID 422 -> p_any 0.41, p_i 0.71 0.21 0.04 0.03 0.01.  Needs the full
	catalogue sizes for the densities, which the subset fixture keeps in *_nfull; the subset changes n, so we
	check the published 2-decimal values with the area rescaled to restore the source densities."""
	z = load('cosmos_subset.npz')
	tables = cases.cosmos_subset(3)
	for t in tables:
		t['area'] = 2.0 * len(t['ra']) / int(z[t['name'] + '_nfull'])
	out = O.nway_match(tables, 20, 1.0, unrelated_mode='cli')
	g = out['XMM'] == 368
	assert g.sum() == 221
	assert abs(out['prob_has_match'][g][0] - 0.41) < 0.005
	top = np.sort(out['prob_this_match'][g])[::-1][:5]
	assert np.allclose(top, [0.71, 0.21, 0.04, 0.03, 0.01], atol=0.006)
	api = O.nway_match(tables, 20, 1.0, unrelated_mode='api')
	assert abs(api['prob_has_match'][g][0] - 0.9696) < 0.001


def test_flat_hash_predicate_reproduces_the_reference_row_set():
	"""enumerator='reference' -- the complete enumeration restricted to the tuples whose present members span at most
	one cell of `radius` degrees in int(ra / err) and in int(dec / err) (O.flat_hash_keeps; the same predicate the
	device applies under NWB_COMPAT_FLAT_HASH) -- IS the flat-sky hash restated loop by loop (enumerator='refhash',
	fastskymatch.py:119-133,164-218), table for table, also where that hash is incomplete (mid-latitudes); and the
	complete enumeration is a superset of it"""
	from tests.test_gpu_fuzz import random_case
	incomplete = 0
	for seed in (1003, 1004, 1019, 1038, 1071, 2059, 2074):
		tables, radius, pc, kw, kind = random_case(seed)
		assert not kw
		assert O.flat_sky_applicable([(t['ra'], t['dec']) for t in tables], radius / 3600)
		full = O.nway_match([dict(t) for t in tables], radius, pc, enumerator='complete')
		pred = O.nway_match([dict(t) for t in tables], radius, pc, enumerator='reference')
		ref = O.nway_match([dict(t) for t in tables], radius, pc, enumerator='refhash')
		assert list(pred) == list(ref)
		for c in ref:
			if c.startswith('_'):
				continue
			a, b = np.asarray(pred[c]), np.asarray(ref[c])
			assert a.shape == b.shape and ((a == b) | ((a != a) & (b != b))).all(), (seed, c)
		names = [t['name'] for t in tables]
		have = set(map(tuple, np.stack([full[n] for n in names], axis=1).tolist()))
		assert all(tuple(r) in have for r in np.stack([ref[n] for n in names], axis=1).tolist())
		incomplete += int(len(ref[names[0]]) < len(full[names[0]]))
	assert incomplete >= 4
	for name in ('offeq2', 'offeq3'):   # and on the off-equator golden cases (outputs of the real reference)
		spec = cases.GOLDEN_CASES[name]
		a = O.create_match_table(cases.build_case(name), spec['radius'], 'reference')
		b = O.create_match_table(cases.build_case(name), spec['radius'], 'refhash')
		assert np.array_equal(a['idx'], b['idx'])


def test_tangent_plane_offsets_follow_astropys_algorithm():
	"""dist3d (fastskymatch.py:50-74) is astropy's SkyOffsetFrame; astropy is absent, so the closed form the oracle and the
	device evaluate (O.offsets) is held against a step-by-step restatement of astropy's own code path
	(rotation matrices, cartesian round trip, longitude wrapped at 180 deg: O.offsets_skyoffsetframe) -- everywhere on the
	sphere, across ra = 0, next to the poles, for offsets from 0.01 arcsec to a degree"""
	rng = np.random.default_rng(4)
	n = 400000
	ra1 = rng.uniform(0, 360, n)
	dec1 = np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
	dec1[:2000] = np.sign(rng.uniform(-1, 1, 2000)) * (90 - 10 ** rng.uniform(-4, 0, 2000))
	ra1[2000:4000] = rng.choice([0.0, 359.9999, 1e-5, 180.0], 2000)
	step = 10 ** rng.uniform(-5.5, 0, n)
	ang = rng.uniform(0, 2 * np.pi, n)
	dec2 = np.clip(dec1 + step * np.cos(ang), -90, 90)
	ra2 = (ra1 + step * np.sin(ang) / np.maximum(np.cos(np.radians(dec1)), 1e-2)) % 360
	sep, dra, ddec = O.offsets((ra1, dec1), (ra2, dec2))
	dra2, ddec2 = O.offsets_skyoffsetframe((ra1, dec1), (ra2, dec2))
	far = np.abs(dec1) < 89.9
	assert np.abs(dra - dra2)[far].max() * 3600 < 1e-10 and np.abs(ddec - ddec2).max() * 3600 < 1e-10
	# next to a pole the offset longitude is ill-conditioned in both forms alike; the quantity that is used -- the offset
	# vector's length and direction on the sky -- still agrees
	d1 = np.hypot(dra * np.cos(np.radians(ddec)), ddec)
	d2 = np.hypot(dra2 * np.cos(np.radians(ddec2)), ddec2)
	assert np.abs(d1 - d2).max() * 3600 < 1e-9
	# and the length of the offset vector is the great-circle separation, to second order in the offset
	small = step < 1e-2
	assert np.abs(d1 - sep)[small].max() * 3600 < 1e-3


def correction_row_by_row(mt, lbf, nu, nu_plus, group_start):
	"""nway.py:366-421 as the script walks it: for every row i that lacks two or more catalogues, over every row j of the
	same primary with ncat > 2, one sub-association at a time (what oracle.correct_unrelated_cli computes array-wise)"""
	idx = mt['idx']
	n = idx.shape[1]
	out = lbf.copy()
	ncat = mt['ncat']
	starts = list(group_start) + [len(idx)]
	for g in range(len(starts) - 1):
		lo, hi = starts[g], starts[g + 1]
		rich = [j for j in range(lo, hi) if ncat[j] > 2]
		if not rich:
			continue
		cache = {}
		for i in range(lo, hi):
			if ncat[i] > n - 2:
				continue
			missing = [k for k in range(1, n) if idx[i, k] == -1]
			best = 0.0
			for j in rich:
				aug = tuple(k for k in missing if idx[j, k] != -1)
				if len(aug) < 2:
					continue
				key = (j, aug)
				if key not in cache:
					pr = nu[aug[0]] / np.prod(nu_plus[list(aug)])
					if 'off' in mt:   # nway.py:404-411
						sra = [[np.array([mt['off'][(a, b)][0][j]]) if a < b else None for b in aug] for a in aug]
						sde = [[np.array([mt['off'][(a, b)][1][j]]) if a < b else None for b in aug] for a in aug]
						errs = [tuple(np.array([x[j]]) for x in mt['errors'][k]) for k in aug]
						val = O.log_bf_elliptical(sra, sde, errs)[0]
					else:
						# nway.py:389-392 builds one numpy.array from float32 separations and float64 NaNs: float64
						p = [[[float(mt['sep'][(a, b)][j])] if a < b else None for b in aug] for a in aug]
						s = [[mt['errors'][k][j]] for k in aug]
						val = O.log_bf(p, s)[0]
					cache[key] = float(val + np.log10(pr))
				best = max(best, cache[key])
			if best > 0:
				out[i] += best
	return out


def test_array_wise_correction_equals_the_row_by_row_walk():
	"""the oracle's unrelated-association correction (one vectorised log_bf per pair of presence patterns, segmented
	maximum) against the nested loops of the script, bit for bit: 3 to 5 catalogues, circular and elliptical errors,
	float32 separations as the command line has them"""
	rng = np.random.default_rng(9)
	for counts, sigmas, side, radius, variants in (((30, 350, 300), (1.0, 0.4, 0.6), 0.014, 8.0, 4), ((16, 130, 130, 110), (1.0, 0.4, 0.5, 0.8), 0.01, 6.0, 4),
			((5, 30, 30, 27, 30), (1.0, 0.4, 0.5, 0.8, 0.6), 0.005, 6.0, 1)):   # the walk is quartic in the group size: five catalogues once
		for elliptical, f32 in ((False, False), (True, True), (False, True), (True, False))[:variants]:
			tables = cases.uniform_patch(len(counts), counts, sigmas, side)
			if elliptical:
				for t in tables:
					n = len(t['ra'])
					a = rng.uniform(0.3, 2, n)
					t['error'] = tuple(O.convert_from_ellipse(a, rng.uniform(0.2, 1, n) * a, rng.uniform(0, np.pi, n)))
			mt = O.create_match_table(tables, radius, sep_f32=f32)
			nu, nu_plus = O.source_densities(tables)
			prior, lbf = O.single_log_bf(mt, nu, nu_plus, O.completeness_vector(0.9, len(tables)))
			starts = O.group_starts(mt['idx'][:, 0])
			fast = O.correct_unrelated_cli(mt, lbf, nu, nu_plus, starts)
			assert (fast != lbf).sum() > 10, (counts, int((fast != lbf).sum()))
			assert np.array_equal(fast, correction_row_by_row(mt, lbf, nu, nu_plus, starts)), (counts, elliptical, f32)
