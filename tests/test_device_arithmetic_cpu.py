"""CPU: the device arithmetic, built for the host from the SAME source the CUDA library is compiled from
(nway_b200/csrc/nwb_device.cuh through tests/emu/arith_emu.cpp), against numpy / the oracle on millions of arguments.
What differs between this build and the device is only the math library underneath (glibc here, CUDA's there); what
is checked is everything the source adds on top of it: the operation order of the reference, the division-free
x/180 and x/pi, the polynomial shortcuts for small angles, the table-driven 10^x, the memo-free log Bayes factor."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import nway_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = ctypes.c_void_p


@pytest.fixture(scope='module')
def emu(tmp_path_factory):
	out = str(tmp_path_factory.mktemp('emu') / 'arith_emu.so')
	res = subprocess.run(['g++', '-O2', '-ffp-contract=off', '-std=c++17', '-fPIC', '-shared', '-w', '-I', os.path.join(ROOT, 'tests', 'emu'),
		'-o', out, os.path.join(ROOT, 'tests', 'emu', 'arith_emu.cpp')], capture_output=True, text=True)
	assert res.returncode == 0, res.stderr[-3000:]
	return ctypes.CDLL(out)


def ptr(a):
	return a.ctypes.data


def ulps(a, b):
	with np.errstate(invalid='ignore', divide='ignore'):
		return np.abs(a - b) / np.spacing(np.maximum(np.abs(a), np.abs(b)))


def test_division_by_180_and_pi_is_the_ieee_quotient(emu):
	rng = np.random.default_rng(0)
	x = np.concatenate((rng.uniform(-400, 400, 3000000), rng.uniform(-1e-3, 1e-3, 1000000), 10 ** rng.uniform(-300, 300, 1000000),
		np.array([0.0, -0.0, 180.0, 360.0, 90.0, np.pi, 1e-320, 5e-324])))
	a, b = np.empty_like(x), np.empty_like(x)
	emu.nwb_emu_div(ctypes.c_longlong(len(x)), P(ptr(x)), P(ptr(a)), P(ptr(b)))
	assert (a == x / 180).all() and (b == x / np.pi).all()


def test_quotient_from_memoised_reciprocal_is_the_ieee_quotient(emu):
	"""-q / 2 / wsum of the 2-catalogue Bayes factor (bayesdistance.py:84) without a division per row"""
	rng = np.random.default_rng(12)
	n = 6000000
	a = -(10 ** rng.uniform(-12, 12, n)) * rng.uniform(0.5, 1, n)
	b = 10 ** rng.uniform(-6, 8, n) * rng.uniform(0.5, 1, n)
	a[:1000] = -(10 ** rng.uniform(-300, 300, 1000))
	b[1000:2000] = 10 ** rng.uniform(-300, 300, 1000)
	a[2000:2010] = [0.0, -0.0, -1e-320, -5e-324, -1e308, -np.inf, np.nan, -1.0, -3.0, -1e-281]
	out = np.empty(n)
	emu.nwb_emu_quotient(ctypes.c_longlong(n), P(ptr(a)), P(ptr(b)), P(ptr(out)))
	with np.errstate(over='ignore', under='ignore', invalid='ignore'):
		ref = a / b
	assert ((out == ref) | (np.isnan(out) & np.isnan(ref))).all()


def test_sin_cos_have_the_bits_of_the_reference(emu):
	"""sin_ref / cos_ref (nwb_device.cuh) against libm -- which is what numpy's float64 sin / cos call, i.e. what the
	reference's separations are made of (fastskymatch.py:36-41).  Not "within an ulp": the same bits."""
	rng = np.random.default_rng(3)
	edges = [0.126, 0.85546875, 2.426265, np.pi / 2, np.pi, 2 * np.pi, 2.0 ** -26, 2.0 ** -27, 105414350.0]
	x = np.concatenate((rng.uniform(-np.pi / 2, np.pi / 2, 4000000),          # latitudes
		rng.uniform(-7, 7, 4000000),                                            # longitude differences
		10 ** rng.uniform(-12, -1, 1000000) * rng.choice([-1, 1], 1000000),     # ... of close pairs
		rng.uniform(-1.06e8, 1.06e8, 500000),
		np.arange(-400, 400) / 128.0, (np.arange(-400, 400) + 0.5) / 128.0,     # the table's nodes and mid-points
		np.concatenate([np.nextafter(s * v, [-np.inf, np.inf]) for v in edges for s in (1, -1)]), np.array(edges), -np.array(edges),
		np.array([0.0, -0.0, 1e-300, 5e-324])))
	x = np.concatenate((x, np.concatenate([np.nextafter(s * v, [-np.inf, np.inf]) for v in (2.0 ** -8, 2.0 ** -7, 1.5 / 128) for s in (1, -1)]),
		np.array([2.0 ** -8, -2.0 ** -8]), rng.uniform(-2.0 ** -7, 2.0 ** -7, 2000000)))
	s, c = np.empty_like(x), np.empty_like(x)
	# three entry points: latitudes (shared table row), small longitude differences (no table), every branch
	for fn in (emu.nwb_emu_sincos, emu.nwb_emu_sincos_small, emu.nwb_emu_sincos_any):
		s[:] = np.nan
		c[:] = np.nan
		fn(ctypes.c_longlong(len(x)), P(ptr(x)), P(ptr(s)), P(ptr(c)))
		assert (s == np.sin(x)).all() and (c == np.cos(x)).all()
		assert (np.signbit(s) == np.signbit(np.sin(x))).all()
	# the scalar libm entry points the table was read from
	for v in x[::200003]:
		assert math.sin(v) == s[np.flatnonzero(x == v)[0]] and math.cos(v) == c[np.flatnonzero(x == v)[0]]


def test_separation_follows_the_reference_formula(emu):
	rng = np.random.default_rng(1)
	n = 2000000
	ra1 = rng.uniform(0, 360, n)
	dec1 = np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
	dec1[:1000] = np.where(rng.uniform(size=1000) < 0.5, 90.0, -90.0) - rng.uniform(-1e-6, 1e-6, 1000) ** 2 * np.sign(rng.uniform(-1, 1, 1000))
	dec1 = np.clip(dec1, -90, 90)
	step = 10 ** rng.uniform(-7, 1.0, n)                      # 0.4 milli-arcsec .. 10 degrees
	ang = rng.uniform(0, 2 * np.pi, n)
	dec2 = np.clip(dec1 + step * np.cos(ang), -90, 90)
	ra2 = ra1 + step * np.sin(ang) / np.maximum(np.cos(np.radians(dec1)), 1e-3)
	ra2[::7] = ra1[::7]
	dec2[::11] = dec1[::11]
	out = np.empty(n)
	emu.nwb_emu_sep(ctypes.c_longlong(n), P(ptr(ra1)), P(ptr(dec1)), P(ptr(ra2)), P(ptr(dec2)), P(ptr(out)))
	ref = O.dist((ra1, dec1), (ra2, dec2)) * 60 * 60
	assert np.isfinite(out).all()
	# sin / cos carry the reference's bits (sin_ref / cos_ref), the products and sums keep its order: the cancelling
	# part of the formula (fastskymatch.py:44) is reproduced exactly, and what is left -- hypot / atan2 by their
	# small-angle polynomials against the library's -- is a few ulp of the value itself
	err = np.abs(out - ref)
	assert (err <= 4 * np.spacing(ref)).all(), (err.max(), ref[np.argmax(err)], ra1[np.argmax(err)], dec1[np.argmax(err)])
	assert (out == ref).mean() > 0.8
	same = ra2 == ra1
	coincident = same & (dec2 == dec1)
	assert (out[coincident] == 0).all()


def test_exp10_table(emu):
	rng = np.random.default_rng(2)
	x = np.concatenate((rng.uniform(-330, 310, 3000000), rng.uniform(-1, 1, 1000000), np.array([0.0, -0.0, 1.0, -1.0, 308.25, -323.3, -400.0, 400.0, np.nan, np.inf, -np.inf])))
	out = np.empty_like(x)
	emu.nwb_emu_exp10(ctypes.c_longlong(len(x)), P(ptr(x)), P(ptr(out)))
	with np.errstate(over='ignore', under='ignore'):
		ref = np.power(10.0, x.astype(np.longdouble)).astype(np.float64)   # 80-bit power, rounded once
	fin = np.isfinite(ref) & (ref > 1e-300)
	assert ulps(out[fin], ref[fin]).max() <= 1.5
	assert np.isnan(out[np.isnan(x)]).all() and (out[x == np.inf] == np.inf).all() and (out[x == -np.inf] == 0).all()
	assert (out[x > 308.3] == np.inf).all() and (out[(x < -323.4)] == 0).all()
	sub = np.isfinite(ref) & (ref <= 1e-300) & (ref > 0)
	assert ulps(out[sub], ref[sub]).max() <= 2   # results near / below the normal range are scaled in two steps


@pytest.mark.parametrize('ncat', [2, 3, 4])
def test_log_bayes_factor(emu, ncat):
	rng = np.random.default_rng(10 + ncat)
	n = 300000
	sig = rng.uniform(0.05, 5, (n, ncat))
	npair = ncat * (ncat - 1) // 2
	sep = rng.uniform(0, 30, (n, npair))
	norm = np.array([(k - 1) * math.log(2) + 2 * (k - 1) * O.LOG_ARCSEC2RAD for k in range(ncat + 1)])
	present = np.full(n, (1 << ncat) - 1, dtype=np.uint32)
	out = np.empty(n)
	emu.nwb_emu_log_bf(ctypes.c_longlong(n), ncat, P(ptr(norm)), ctypes.c_double(O.LOG10_E), P(ptr(present)), P(ptr(sig)), P(ptr(sep)), P(ptr(out)))
	p = [[None] * ncat for _ in range(ncat)]
	k = 0
	for a in range(ncat):
		for b in range(a + 1, ncat):
			p[a][b] = sep[:, k]
			k += 1
	ref = O.log_bf(p, [sig[:, c] for c in range(ncat)])
	# same operations in the same order, except w = sigma^-2: numpy's pow(sigma, -2.0) against 1 / (sigma * sigma) on the device
	# (two roundings: up to 1 ulp apart), so the sums agree to a few ulp of their largest term
	w = sig ** -2.
	big = sum(w[:, a] * w[:, b] * p[a][b] ** 2 for a in range(ncat) for b in range(a + 1, ncat)) / 2 / w.sum(axis=1) * O.LOG10_E
	tol = lambda r: 8 * np.spacing(np.maximum(np.maximum(np.abs(r), big), 64.0))   # the exponent term can dwarf the result
	assert (np.abs(out - ref) <= tol(ref)).all(), np.abs(out - ref).max()
	assert (out == ref).mean() > 0.5
	# a sub-association (catalogue 1 absent) equals the Bayes factor of the remaining catalogues
	if ncat >= 3:
		present[:] = ((1 << ncat) - 1) & ~2
		emu.nwb_emu_log_bf(ctypes.c_longlong(n), ncat, P(ptr(norm)), ctypes.c_double(O.LOG10_E), P(ptr(present)), P(ptr(sig)), P(ptr(sep)), P(ptr(out)))
		keep = [c for c in range(ncat) if c != 1]
		p2 = [[None] * len(keep) for _ in keep]
		for i, a in enumerate(keep):
			for j, b in enumerate(keep):
				if i < j:
					p2[i][j] = p[a][b]
		ref2 = O.log_bf(p2, [sig[:, c] for c in keep])
		assert (np.abs(out - ref2) <= tol(ref2)).all()


def test_posterior(emu):
	rng = np.random.default_rng(5)
	n = 1000000
	prior = 10 ** rng.uniform(-14, 0, n)
	prior[:10] = 1.0
	lbf = rng.uniform(-500, 40, n)
	l10p = np.log10(prior)
	out = np.empty(n)
	emu.nwb_emu_posterior(ctypes.c_longlong(n), P(ptr(prior)), P(ptr(l10p)), P(ptr(lbf)), P(ptr(out)))
	with np.errstate(invalid='ignore'):
		ref = O.posterior(prior, lbf)   # prior = 1 with an overflowing exponential is 0 * inf = NaN in the reference too
	assert np.allclose(out, ref, rtol=4e-16 * 3, atol=0, equal_nan=True)   # glibc's exp10 vs numpy's 10**x: a few ulp


def test_tangent_plane_offsets(emu):
	rng = np.random.default_rng(6)
	n = 1000000
	ra_o = rng.uniform(0, 360, n)
	dec_o = np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
	step = 10 ** rng.uniform(-6, -0.5, n)
	ang = rng.uniform(0, 2 * np.pi, n)
	dec_t = np.clip(dec_o + step * np.cos(ang), -90, 90)
	ra_t = (ra_o + step * np.sin(ang) / np.maximum(np.cos(np.radians(dec_o)), 1e-2)) % 360
	dra, ddec = np.empty(n), np.empty(n)
	emu.nwb_emu_offsets(ctypes.c_longlong(n), P(ptr(ra_o)), P(ptr(dec_o)), P(ptr(ra_t)), P(ptr(dec_t)), P(ptr(dra)), P(ptr(ddec)))
	_, rdra, rddec = O.offsets((ra_o, dec_o), (ra_t, dec_t))
	# the closed form cancels like the separation formula does: ~1e-16 rad absolute
	assert (np.abs(dra - rdra * 3600) <= 8e-11 + 8 * np.spacing(np.abs(rdra * 3600))).all()
	assert (np.abs(ddec - rddec * 3600) <= 8e-11 + 8 * np.spacing(np.abs(rddec * 3600))).all()


def test_elliptical_bayes_factor(emu):
	"""ell_rescaled_sep + log_bf_ref<2> against the oracle's log_bf_elliptical, itself pinned to the reference's
	(tests/golden/kat.npz, test_kat_dist_logbf_posterior_elliptical)"""
	rng = np.random.default_rng(7)
	n = 500000
	vx, vy = rng.normal(0, 3, n), rng.normal(0, 3, n)
	vx[:100] = 0
	vy[:100] = 0
	ell = []
	for _ in range(2):
		a = rng.uniform(0.2, 4, n)
		ell.append(O.convert_from_ellipse(a, a * rng.uniform(0.2, 1, n), rng.uniform(0, np.pi, n)))
	ea = np.ascontiguousarray(np.stack(ell[0], axis=1))
	eb = np.ascontiguousarray(np.stack(ell[1], axis=1))
	norm = np.array([(k - 1) * math.log(2) + 2 * (k - 1) * O.LOG_ARCSEC2RAD for k in range(3)])
	out = np.empty(n)
	emu.nwb_emu_log_bf_ell2(ctypes.c_longlong(n), P(ptr(norm)), ctypes.c_double(O.LOG10_E), P(ptr(vx)), P(ptr(vy)), P(ptr(ea)), P(ptr(eb)), P(ptr(out)))
	ref = O.log_bf_elliptical([[None, vx], [None, None]], [[None, vy], [None, None]], ell)
	assert (np.abs(out - ref) <= 1e-12 * np.maximum(np.abs(ref), 100.0)).all(), np.abs(out - ref).max()


def test_magnitude_prior_lookup(emu):
	from nway_b200 import magnitudeweights
	rng = np.random.default_rng(8)
	nb = 16
	edges = np.sort(rng.uniform(15, 28, nb + 1))
	hs, ha = rng.uniform(0.01, 0.3, nb), rng.uniform(0.01, 0.3, nb)
	ha[3] = 0.0    # ratio 100 (magnitudeweights.py:23)
	hs[11] = 0.0   # ratio 0 -> weight -inf
	e, weight, bias = magnitudeweights.step_tables(edges, hs, ha)
	m = np.concatenate((rng.uniform(10, 32, 200000), edges, np.nextafter(edges, 40), np.nextafter(edges, 0), np.array([-99.0, np.nan, np.inf, -np.inf])))
	w, b = np.empty_like(m), np.empty_like(m)
	emu.nwb_emu_mag_weight(ctypes.c_longlong(len(m)), nb, P(ptr(e)), P(ptr(weight)), P(ptr(bias)), P(ptr(m)), P(ptr(w)), P(ptr(b)))
	with np.errstate(divide='ignore', invalid='ignore'):
		ref = np.log10(O.bias_lookup(edges, hs, ha, m))   # __init__.py:386: log10 of the interpolated ratio ...
	ref[np.isnan(ref)] = 0                                    # ... NaN (outside the table, undefined magnitude) -> 0 (:388)
	assert ((w == ref) | (np.isinf(w) & (w == ref))).all()
	assert np.array_equal(b, 10 ** ref)


def test_fused_group_normalisation(emu):
	"""the one-exponential-per-row normalisation of k_rows2 (p_any without the logarithm and the "1 -", p_i from the
	shared exponentials, dist_post from the same exponentials) against the reference's formulas (__init__.py:423-457,
	bayesdistance.py:26-32) on random groups: ordinary ones, groups dominated by the no-counterpart row or by one
	counterpart, log-weights hundreds of decades apart, ties, lone rows"""
	rng = np.random.default_rng(9)
	prior1 = 3.7e-4
	l10p1 = float(np.log10(prior1))
	worst = dict(p_any=0.0, p_i=0.0, post=0.0)
	ngroups = 30000
	for g in range(ngroups):
		rows = int(rng.integers(1, 129))
		kind = g % 6
		spread = [3, 30, 300, 3, 3, 1000][kind]
		lbf = rng.normal(0, spread, rows)
		if kind == 3:
			lbf[1:] -= rng.uniform(0, 40)       # the no-counterpart row dominates
		if kind == 4 and rows > 2:
			lbf[2] = lbf[1]                         # a tie for the best counterpart
		lbf[0] = 0.0
		v = lbf + l10p1
		v[0] = 0.0                                  # row 0: log10(prior = 1) + log BF 0
		p_any = ctypes.c_double()
		p_i, post = np.empty(rows), np.empty(rows)
		flag = np.empty(rows, dtype=np.int64)
		emu.nwb_emu_group(rows, P(ptr(v)), P(ptr(lbf)), ctypes.c_double(prior1), ctypes.c_double(l10p1), ctypes.c_double(0.5),
			ctypes.byref(p_any), P(ptr(p_i)), P(ptr(flag)), P(ptr(post)))
		with np.errstate(over='ignore', divide='ignore', invalid='ignore'):
			rpa, rpi, rflag = O.group_statistics(v, [0], 0.5)
			rpost = O.posterior(np.r_[1.0, np.full(rows - 1, prior1)], lbf)
		if rows > 1 and not np.isfinite(rpi).all():
			continue   # the reference itself overflows here (10**(v - bfsum1) with bfsum1 = -inf)
		# the parity metric of tests/parity.py
		assert abs(p_any.value - rpa[0]) <= 1e-10 * abs(rpa[0]) + 2e-13, (g, rows, p_any.value, rpa[0])
		big = rpi >= 1e-30
		assert (np.abs(p_i[big] - rpi[big]) <= 1e-10 * rpi[big]).all() and (np.abs(p_i[~big] - rpi[~big]) <= 1e-40).all(), (g, rows)
		ok = np.abs(post - rpost) <= 1e-10 * np.abs(rpost) + 1e-300
		assert ok.all(), (g, rows, post[~ok], rpost[~ok])
		# flags: exact unless a p_i sits on the 0.5 x best threshold or ties with the best within rounding
		sure = (np.abs(rpi - 0.5 * rpi.max()) > 1e-9 * rpi.max()) & ((rpi == rpi.max()) | (np.abs(rpi - rpi.max()) > 1e-9 * rpi.max()))
		assert (flag[sure] == rflag[sure]).all(), (g, rows)
		worst['p_any'] = max(worst['p_any'], abs(p_any.value - rpa[0]))
		if big.any():
			worst['p_i'] = max(worst['p_i'], float(np.max(np.abs(p_i[big] - rpi[big]) / rpi[big])))
	print('worst differences from the reference formulas over %d groups: %s' % (ngroups, worst))
	assert worst['p_i'] < 5e-12 and worst['p_any'] < 2e-13, worst
