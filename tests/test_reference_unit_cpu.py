"""CPU: the host-side part of the reference's own unit tests (tests/bayesdistance_test.py:34-147,206-230) against
nway_b200.bayesdistance -- the ellipse conversion and the 2 x 2 covariance algebra are host numpy here as there.  The tests of
the functions that run on the device (log_bf, log_bf_elliptical, dist, match_multiple) are tests/test_gpu_reference_unit.py."""
import numpy
import numpy.testing as test

from nway_b200.bayesdistance import (apply_vABv, apply_vector_left, apply_vector_right, assert_possemdef, convert_from_ellipse, make_covmatrix,
	make_invcovmatrix, matrix_add, matrix_det, matrix_invert, matrix_multiply, vector_multiply, vector_normalised)


def test_ellipse_conversion():
	rng = numpy.random.RandomState(1)
	sigma_ra = rng.uniform(1, 100, size=100)
	sigma_dec = rng.uniform(1, 100, size=100)
	angles = rng.uniform(0, 180, size=100)
	sigma_x, sigma_y, rho = convert_from_ellipse(sigma_ra, sigma_ra, angles)   # a circle, whatever the angle
	test.assert_almost_equal(rho, 0.)
	test.assert_almost_equal(sigma_x, sigma_ra)
	test.assert_almost_equal(sigma_y, sigma_ra)
	sigma_x, sigma_y, rho = convert_from_ellipse(sigma_ra, sigma_dec, 0)   # aligned with the axes
	test.assert_almost_equal(rho, 0.)
	test.assert_almost_equal(sigma_y, sigma_ra)
	test.assert_almost_equal(sigma_x, sigma_dec)
	sigma_x, sigma_y, rho = convert_from_ellipse(sigma_ra, sigma_dec, numpy.pi / 2)   # a quarter turn
	test.assert_almost_equal(rho, 0.)
	test.assert_almost_equal(sigma_x, sigma_ra)
	test.assert_almost_equal(sigma_y, sigma_dec)
	sigma_x, sigma_y, rho_corr = convert_from_ellipse(sigma_ra, sigma_dec, angles)
	assert numpy.all(numpy.abs(rho_corr) > 1e-6), rho_corr


def test_corrmatrix_distances_circular():
	n = 10
	dra, ddec = numpy.arange(n), 0 * numpy.arange(n)
	s1, s2 = numpy.ones(n), numpy.ones(n) / 10
	by_hand = dra**2 / s1**2 + ddec**2 / s1**2 + dra**2 / s2**2 + ddec**2 / s2**2
	assert by_hand[0] == 0 and numpy.isclose(by_hand[1], 101)
	total = (s1**-2 + s2**-2)**-0.5
	test.assert_almost_equal(by_hand, (dra / total)**2 + (ddec / total)**2)
	A, B = make_invcovmatrix(s1, s1), make_invcovmatrix(s2, s2)
	test.assert_almost_equal(by_hand, apply_vABv((dra, ddec), A, B))
	test.assert_almost_equal(apply_vABv((dra, ddec), B, A), apply_vABv((dra, ddec), A, B))   # the order of the catalogues does not matter


def test_corrmatrix_distances_elliptical():
	dvec = numpy.array([0., 1., 1., 1.]), numpy.array([0., 0., 1., -1.])
	sigma_ra, sigma_dec = numpy.ones(4), 2 * numpy.ones(4)
	sx, sy, rho = convert_from_ellipse(sigma_ra, sigma_dec, -45 * numpy.pi / 180)
	assert numpy.all(numpy.abs(rho) > 0.001), rho
	wide = make_invcovmatrix(sigma_ra * 10, sigma_dec * 10)
	symmetric = apply_vABv(dvec, make_invcovmatrix(sigma_ra, sigma_dec), wide)
	tilted = apply_vABv(dvec, make_invcovmatrix(sx, sy, rho), wide)
	test.assert_almost_equal(tilted[0], symmetric[0])
	assert tilted[2] < symmetric[2] / 1.4, (tilted[2], symmetric[2])   # a step along the covariance counts for less
	assert tilted[3] > symmetric[3] * 1.4, (tilted[3], symmetric[3])   # a step across it for more


def test_diag_mult():
	dx = 1.0
	v = numpy.array([dx, 0.]), numpy.array([0., dx])
	s1, s2 = numpy.ones(2), 2 * numpy.ones(2)
	d = apply_vABv(v, make_invcovmatrix(s1, s1, 0), make_invcovmatrix(s2, s2, 0))
	test.assert_almost_equal(dx * (1 / s1**2 + 1 / s2**2), d)


def test_ell_semdef_pos():
	rng = numpy.random.RandomState(3)
	n = 400
	one = make_invcovmatrix(*convert_from_ellipse(1.0, rng.uniform(size=n), rng.uniform(0, 2 * numpy.pi)))
	two = make_invcovmatrix(*convert_from_ellipse(rng.uniform(size=n), rng.uniform(size=n), rng.uniform(0, 2 * numpy.pi)))
	assert_possemdef(one)
	assert_possemdef(two)
	assert_possemdef(matrix_add(one, two))
	v = rng.normal(0, 1, size=(2, n))
	assert (apply_vABv(v, one, two) >= 0).all()
	assert (vector_multiply(v, apply_vector_right(matrix_add(one, two), v)) >= 0).all()
	try:
		assert_possemdef(((numpy.ones(3), 2 * numpy.ones(3)), (2 * numpy.ones(3), numpy.ones(3))))   # eigenvalues 3 and -1
	except AssertionError:
		pass
	else:
		raise AssertionError('an indefinite matrix passed')


def test_the_algebra_is_an_algebra():
	"""what the reference's tests leave out: inverse, product, determinant, left and right application, unit vectors"""
	rng = numpy.random.RandomState(5)
	n = 50
	sx, sy, rho = rng.uniform(0.2, 3, n), rng.uniform(0.2, 3, n), rng.uniform(-0.9, 0.9, n)
	C, P = make_covmatrix(sx, sy, rho), make_invcovmatrix(sx, sy, rho)
	I = matrix_multiply(C, P)
	test.assert_almost_equal(I[0][0], 1) ; test.assert_almost_equal(I[1][1], 1) ; test.assert_almost_equal(I[0][1], 0) ; test.assert_almost_equal(I[1][0], 0)
	Q = matrix_invert(C)
	for i in (0, 1):
		for j in (0, 1):
			test.assert_almost_equal(Q[i][j], P[i][j])
	test.assert_almost_equal(matrix_det(C) * matrix_det(P), 1)
	v = rng.normal(size=(2, n))
	test.assert_almost_equal(vector_multiply(apply_vector_left(v, C), v), vector_multiply(v, apply_vector_right(C, v)))
	u = vector_normalised((numpy.array([3., 0., 0.]), numpy.array([4., 0., -2.])))
	test.assert_almost_equal(u[0], [0.6, 2**-0.5, 0.]) ; test.assert_almost_equal(u[1], [0.8, 2**-0.5, -1.])
