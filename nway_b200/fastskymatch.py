"""GPU-backed mirror of the array-level surface of nwaylib/fastskymatch.py that other code calls
(SURVEY.md 8b): dist()."""
import numpy

from . import _lib


def dist(apos, bpos, device=None):
	"""Angular separation (degrees) between two points on a sphere; same arithmetic as fastskymatch.py:26-47,
	evaluated on the GPU (nwb_dist)."""
	(a_ra, a_dec), (b_ra, b_dec) = apos, bpos
	a_ra, a_dec, b_ra, b_dec = numpy.broadcast_arrays(*[numpy.asarray(x, dtype=float) for x in (a_ra, a_dec, b_ra, b_dec)])
	shape = a_ra.shape
	arrs = [_lib.f64(x).ravel() for x in (a_ra, a_dec, b_ra, b_dec)]
	out = numpy.empty(arrs[0].size)
	ctx = _lib.get_context(device)
	ctx.check(ctx.lib.nwb_dist(ctx.h, out.size, *[_lib.dptr(x) for x in arrs], _lib.dptr(out)))
	return out.reshape(shape)
