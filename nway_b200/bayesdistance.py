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
