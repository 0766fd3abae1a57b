"""GPU-backed mirror of nwaylib/bayesdistance.py's array functions (log_bf, posterior)."""
import numpy
from numpy import log10

from . import _lib


def log_bf(p, s, device=None):
	"""log10 of the multi-way Bayes factor (bayesdistance.py:64-86).
	p: separations matrix (NxN nested list of arrays, only i<j is read), s: list of N error arrays."""
	n = len(s)
	errs = numpy.broadcast_arrays(*[numpy.asarray(si, dtype=float) for si in s])
	shape = errs[0].shape
	size = errs[0].size
	err = numpy.ascontiguousarray(numpy.stack([e.ravel() for e in errs]))
	sep = numpy.full((n, n, size), numpy.nan)
	for i in range(n):
		for j in range(i + 1, n):
			sep[i, j] = numpy.broadcast_to(numpy.asarray(p[i][j], dtype=float), shape).ravel()
	out = numpy.empty(size)
	ctx = _lib.get_context(device)
	ctx.check(ctx.lib.nwb_log_bf(ctx.h, size, n, _lib.dptr(sep), _lib.dptr(err), _lib.dptr(out)))
	return out.reshape(shape)


def log_bf_elliptical(separations_ra, separations_dec, pos_errors, device=None):
	"""log10 Bayes factor for elliptical positional errors (bayesdistance.py:207-240).
	separations_ra / separations_dec: NxN nested lists of offset arrays in arcsec (only i<j is read);
	pos_errors: list of N triples (sigma_ra, sigma_dec, rho)."""
	n = len(pos_errors)
	flat = [numpy.asarray(x, dtype=float) for trip in pos_errors for x in trip]
	flat = numpy.broadcast_arrays(*flat)
	shape = flat[0].shape
	size = flat[0].size
	err = numpy.ascontiguousarray(numpy.stack([e.ravel() for e in flat]))   # (3n, size): c*3 + {x, y, rho}
	sra = numpy.full((n, n, size), numpy.nan)
	sde = numpy.full((n, n, size), numpy.nan)
	for i in range(n):
		for j in range(i + 1, n):
			sra[i, j] = numpy.broadcast_to(numpy.asarray(separations_ra[i][j], dtype=float), shape).ravel()
			sde[i, j] = numpy.broadcast_to(numpy.asarray(separations_dec[i][j], dtype=float), shape).ravel()
	out = numpy.empty(size)
	ctx = _lib.get_context(device)
	ctx.check(ctx.lib.nwb_log_bf_elliptical(ctx.h, size, n, _lib.dptr(sra), _lib.dptr(sde), _lib.dptr(err), _lib.dptr(out)))
	return out.reshape(shape)


def convert_from_ellipse(a, b, phi):
	"""(sigma_x, sigma_y, rho) from major axis, minor axis and angle in radians (bayesdistance.py:190-204);
	one pass over a catalogue column, stays on the host"""
	from . import convert_from_ellipse as _c
	return _c(a, b, phi)


def posterior(prior, log_bf, device=None):
	"""posterior against the unrelated hypothesis (bayesdistance.py:26-32)"""
	prior, log_bf = numpy.broadcast_arrays(numpy.asarray(prior, dtype=float), numpy.asarray(log_bf, dtype=float))
	shape = prior.shape
	a, b = _lib.f64(prior).ravel(), _lib.f64(log_bf).ravel()
	out = numpy.empty(a.size)
	ctx = _lib.get_context(device)
	ctx.check(ctx.lib.nwb_posterior(ctx.h, a.size, _lib.dptr(a), _lib.dptr(b), _lib.dptr(out)))
	return out.reshape(shape)


def unnormalised_log_posterior(prior, log_bf, ncat):
	"""bayesdistance.py:35-39 (two flops: stays on the host)"""
	return log_bf + log10(prior)


# ---- 2 x 2 covariance algebra of the reference's module (bayesdistance.py:88-187), kept for callers -----------------------
# A matrix is ((m11, m12), (m21, m22)), a vector (v1, v2); every entry may be an array: one matrix / vector per element.
# A handful of flops per source on catalogue columns: host numpy, like convert_from_ellipse.  The match itself does this
# algebra on the device (ell_prepare in csrc/nwb_rows.cuh).

def make_covmatrix(sigma_x, sigma_y, rho=0):
	"""covariance matrix of standard deviations sigma_x, sigma_y and normalised correlation rho"""
	off = rho * sigma_x * sigma_y
	return (sigma_x**2, off), (off, sigma_y**2)


def make_invcovmatrix(sigma_x, sigma_y, rho=0):
	"""its inverse (the precision matrix), written out"""
	scale = 1.0 / (sigma_x**2 * sigma_y**2 * (1 - rho**2))
	off = scale * -rho * sigma_x * sigma_y
	return (scale * sigma_y**2, off), (off, scale * sigma_x**2)


def matrix_det(A):
	return A[0][0] * A[1][1] - A[0][1] * A[1][0]


def matrix_add(A, B):
	return tuple(tuple(a + b for a, b in zip(ra, rb)) for ra, rb in zip(A, B))


def matrix_multiply(A, B):
	return tuple(tuple(A[i][0] * B[0][j] + A[i][1] * B[1][j] for j in (0, 1)) for i in (0, 1))


def matrix_invert(A):
	scale = 1.0 / matrix_det(A)
	assert (scale > 0).all()
	return (scale * A[1][1], -scale * A[0][1]), (-scale * A[1][0], scale * A[0][0])


def apply_vector_right(A, b):
	"""A b"""
	return A[0][0] * b[0] + A[0][1] * b[1], A[1][0] * b[0] + A[1][1] * b[1]


def apply_vector_left(a, B):
	"""a^T B"""
	return a[0] * B[0][0] + a[1] * B[1][0], a[0] * B[0][1] + a[1] * B[1][1]


def vector_multiply(a, b):
	"""a . b"""
	return a[0] * b[0] + a[1] * b[1]


def vector_normalised(v):
	"""v / |v|; the diagonal direction for a null vector (bayesdistance.py:164-169)"""
	length = (v[0]**2 + v[1]**2)**0.5
	return tuple(numpy.where(length == 0, 2**-0.5, x / (length + 1e-300)) for x in v)


def apply_vABv(v, A, B):
	"""v^T (A + B) v"""
	return vector_multiply(v, apply_vector_right(matrix_add(A, B), v))


def assert_possemdef(M):
	"""raise AssertionError unless every 2 x 2 matrix of M is positive semi-definite (both eigenvalues real and >= 0)"""
	trace, det = M[0][0] + M[1][1], M[0][0] * M[1][1] - M[0][1] * M[0][1]
	degenerate = numpy.isclose(trace**2, 4 * det)   # a double eigenvalue trace / 2: nothing to take a root of
	if degenerate.all():
		return
	disc = (trace**2 - 4 * det)[~degenerate]
	assert (disc >= 0).all(), (trace, det, M)
	half = trace[~degenerate] / 2
	assert (half + disc**0.5 / 2 >= 0).all() and (half - disc**0.5 / 2 >= 0).all(), (trace, det, M)
