"""HEALPix NESTED pixelisation: the two healpy calls the reference's all-sky branch makes.

TEST INFRASTRUCTURE ONLY (like everything under oracle/).  The reference hashes sources on the whole sphere with

    healpy.pixelfunc.ang2pix(nside, phi=phi, theta=theta, nest=True)              nwaylib/fastskymatch.py:138
    healpy.pixelfunc.get_all_neighbours(nside, phi=phi, theta=theta, nest=True)   nwaylib/fastskymatch.py:139
    healpy.pixelfunc.nside2resol(nside)                                           nwaylib/fastskymatch.py:84

healpy is a third-party dependency of the reference (pyproject.toml:53, conda-requirements.txt:7; no version
pinned), it is not under /root/reference and not installed here.  This module restates the published algorithm
(Gorski et al. 2005, ApJ 622, 759; the `healpix_base` routines ang2pix_z_phi, xyf2nest, nest2xyf and neighbors of
the HEALPix C++ library that healpy wraps) in numpy, so that oracle/refrun.py can run the UNMODIFIED reference
through its HEALPix branch and pin the oracle's all-sky enumeration against it.  tests/test_healpix_stub.py checks
it against the known answers of healpy's own docstrings (ang2pix, get_all_neighbours) and against the structural
properties of the pixelisation (equal areas, symmetric neighbourhoods, 24 pixels with seven neighbours).
"""
import numpy as np

# base-pixel tables of the HEALPix C++ library (healpix_tables / healpix_base)
JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4])
JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7])
# neighbour directions in healpy's order: SW, W, NW, N, NE, E, SE, S
NB_XOFFSET = np.array([-1, -1, 0, 1, 1, 1, 0, -1])
NB_YOFFSET = np.array([0, 1, 1, 1, 0, -1, -1, -1])
# face reached when leaving a base pixel: index 3 * (y overflow + 1) + (x overflow + 1), by base pixel
NB_FACEARRAY = np.array([
	[8, 9, 10, 11, -1, -1, -1, -1, 10, 11, 8, 9],   # S
	[5, 6, 7, 4, 8, 9, 10, 11, 9, 10, 11, 8],       # SE
	[-1, -1, -1, -1, 5, 6, 7, 4, -1, -1, -1, -1],   # E
	[4, 5, 6, 7, 11, 8, 9, 10, 11, 8, 9, 10],       # SW
	[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11],         # centre
	[1, 2, 3, 0, 0, 1, 2, 3, 5, 6, 7, 4],           # NE
	[-1, -1, -1, -1, 7, 4, 5, 6, -1, -1, -1, -1],   # W
	[3, 0, 1, 2, 3, 0, 1, 2, 4, 5, 6, 7],           # NW
	[2, 3, 0, 1, -1, -1, -1, -1, 0, 1, 2, 3]])      # N
# how (x, y) transform on that crossing (bit 0: flip x, bit 1: flip y, bit 2: swap), by face >> 2
NB_SWAPARRAY = np.array([
	[0, 0, 3], [0, 0, 6], [0, 0, 0], [0, 0, 5], [0, 0, 0], [5, 0, 0], [0, 0, 0], [6, 0, 0], [3, 0, 0]])


def nside2resol(nside):
	"""sqrt of the pixel area, radians"""
	return np.sqrt(4 * np.pi / (12. * nside * nside))


def _spread_bits(v):
	"""interleave zeros: bit k of v -> bit 2k"""
	v = np.asarray(v, dtype=np.int64)
	out = np.zeros_like(v)
	for k in range(30):
		out |= ((v >> k) & 1) << (2 * k)
	return out


def _compress_bits(v):
	v = np.asarray(v, dtype=np.int64)
	out = np.zeros_like(v)
	for k in range(30):
		out |= ((v >> (2 * k)) & 1) << k
	return out


def xyf2nest(nside, ix, iy, face):
	return np.asarray(face, dtype=np.int64) * nside * nside + _spread_bits(ix) + 2 * _spread_bits(iy)


def nest2xyf(nside, pix):
	pix = np.asarray(pix, dtype=np.int64)
	npface = nside * nside
	face = pix // npface
	p = pix % npface
	return _compress_bits(p), _compress_bits(p >> 1), face


def xyf2ring(nside, ix, iy, face):
	"""RING index of (ix, iy, face): only used to compare with the RING-scheme examples of healpy's docstrings"""
	ix, iy, face = [np.asarray(a, dtype=np.int64) for a in (ix, iy, face)]
	nl4 = 4 * nside
	ncap = 2 * nside * (nside - 1)
	npix = 12 * nside * nside
	jr = JRLL[face] * nside - ix - iy - 1
	north = jr < nside
	south = jr > 3 * nside
	nr = np.where(north, jr, np.where(south, nl4 - jr, nside))
	n_before = np.where(north, 2 * nr * (nr - 1), np.where(south, npix - 2 * (nr + 1) * nr, ncap + (jr - nside) * nl4))
	kshift = np.where(north | south, 0, (jr - nside) & 1)
	jp = (JPLL[face] * nr + ix - iy + 1 + kshift) // 2
	jp = np.where(jp > nl4, jp - nl4, np.where(jp < 1, jp + nl4, jp))
	return n_before + jp - 1


def _ang2xyf(nside, theta, phi):
	"""ang2pix_z_phi of healpix_base: (theta, phi) -> (ix, iy, face)"""
	theta = np.atleast_1d(np.asarray(theta, dtype=float))
	phi = np.atleast_1d(np.asarray(phi, dtype=float))
	theta, phi = np.broadcast_arrays(theta, phi)
	assert ((theta >= 0) & (theta <= np.pi)).all(), 'theta out of range'
	z = np.cos(theta)
	za = np.abs(z)
	tt = np.mod(phi * (2 / np.pi), 4.0)   # in [0, 4)
	tt = np.where(tt >= 4.0, 0.0, tt)
	# equatorial region
	temp1 = nside * (0.5 + tt)
	temp2 = nside * z * 0.75
	jp = np.floor(temp1 - temp2).astype(np.int64)   # ascending edge line
	jm = np.floor(temp1 + temp2).astype(np.int64)   # descending edge line
	ifp = jp // nside
	ifm = jm // nside
	face_eq = np.where(ifp == ifm, (ifp & 3) | 4, np.where(ifp < ifm, ifp & 3, (ifm & 3) + 8))
	ix_eq = jm & (nside - 1)
	iy_eq = nside - (jp & (nside - 1)) - 1
	# polar caps; near the poles sin(theta) gives 1 - |z| without cancellation (as healpix_base does)
	ntt = np.minimum(3, tt.astype(np.int64))
	tp = tt - ntt
	sth = np.sin(theta)
	tmp = np.where(za >= 0.99, nside * sth / np.sqrt((1.0 + za) / 3.0), nside * np.sqrt(3 * (1 - za)))
	jp2 = np.minimum((tp * tmp).astype(np.int64), nside - 1)
	jm2 = np.minimum(((1.0 - tp) * tmp).astype(np.int64), nside - 1)
	face_po = np.where(z >= 0, ntt, ntt + 8)
	ix_po = np.where(z >= 0, nside - jm2 - 1, jp2)
	iy_po = np.where(z >= 0, nside - jp2 - 1, jm2)
	eq = za <= 2.0 / 3.0
	return np.where(eq, ix_eq, ix_po), np.where(eq, iy_eq, iy_po), np.where(eq, face_eq, face_po)


def ang2pix(nside, theta=None, phi=None, nest=False, lonlat=False):
	"""healpy.pixelfunc.ang2pix (colatitude theta, longitude phi, radians)"""
	assert not lonlat
	nside = int(nside)
	assert nside > 0 and nside & (nside - 1) == 0, 'nside must be a power of two'
	scalar = np.ndim(theta) == 0 and np.ndim(phi) == 0
	ix, iy, face = _ang2xyf(nside, theta, phi)
	out = xyf2nest(nside, ix, iy, face) if nest else xyf2ring(nside, ix, iy, face)
	return int(out[0]) if scalar else out


def neighbours_xyf(nside, ix, iy, face):
	"""the eight neighbours (SW, W, NW, N, NE, E, SE, S) of pixels given as (ix, iy, face), as NESTED indices,
	-1 where a neighbour does not exist; shape (8, n) -- `neighbors` of healpix_base"""
	ix, iy, face = [np.asarray(a, dtype=np.int64) for a in (ix, iy, face)]
	out = np.empty((8,) + ix.shape, dtype=np.int64)
	for m in range(8):
		x = ix + NB_XOFFSET[m]
		y = iy + NB_YOFFSET[m]
		nbnum = np.full(ix.shape, 4, dtype=np.int64)
		lo, hi = x < 0, x >= nside
		x = np.where(lo, x + nside, np.where(hi, x - nside, x))
		nbnum = nbnum - lo.astype(np.int64) + hi.astype(np.int64)
		lo, hi = y < 0, y >= nside
		y = np.where(lo, y + nside, np.where(hi, y - nside, y))
		nbnum = nbnum - 3 * lo.astype(np.int64) + 3 * hi.astype(np.int64)
		f = NB_FACEARRAY[nbnum, face]
		bits = NB_SWAPARRAY[nbnum, face >> 2]
		x = np.where(bits & 1, nside - x - 1, x)
		y = np.where(bits & 2, nside - y - 1, y)
		x, y = np.where(bits & 4, y, x), np.where(bits & 4, x, y)
		out[m] = np.where(f >= 0, xyf2nest(nside, x, y, np.maximum(f, 0)), -1)
	return out


def get_all_neighbours(nside, theta, phi=None, nest=False, lonlat=False):
	"""healpy.pixelfunc.get_all_neighbours: theta alone = pixel indices, (theta, phi) = angles; returns (8, n)"""
	assert not lonlat
	assert nest or nside == 1, 'only the NESTED scheme is restated here (for nside = 1 the two schemes coincide)'
	nside = int(nside)
	if phi is None:
		scalar = np.ndim(theta) == 0
		ix, iy, face = nest2xyf(nside, np.atleast_1d(theta))
	else:
		scalar = np.ndim(theta) == 0 and np.ndim(phi) == 0
		ix, iy, face = _ang2xyf(nside, theta, phi)
	out = neighbours_xyf(nside, ix, iy, face)
	return out[:, 0] if scalar else out
