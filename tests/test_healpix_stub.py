"""CPU: oracle/healpix_nest.py (the stand-in for the un-vendored healpy that lets the unmodified reference run its
all-sky branch in the build container) against the known answers of healpy's own docstrings and the structural
properties of the HEALPix pixelisation."""
import numpy as np

from oracle import healpix_nest as H


def test_ang2pix_docstring_examples_ring():
	# healpy.pixelfunc.ang2pix docstring
	assert H.ang2pix(16, np.pi / 2, 0) == 1440
	assert H.ang2pix(16, [np.pi / 2, np.pi / 4, np.pi / 2, 0, np.pi], [0., np.pi / 4, np.pi / 2, 0, 0]).tolist() == [1440, 427, 1520, 0, 3068]
	assert H.ang2pix(16, np.pi / 2, [0, np.pi / 2]).tolist() == [1440, 1520]
	assert [H.ang2pix(n, np.pi / 2, 0) for n in (1, 2, 4, 8, 16)] == [4, 12, 72, 336, 1440]


def test_get_all_neighbours_docstring_examples():
	# healpy.pixelfunc.get_all_neighbours docstring (nside = 1: RING and NESTED numbering coincide)
	assert H.get_all_neighbours(1, 4).tolist() == [11, 7, 3, -1, 0, 5, 8, -1]
	assert H.get_all_neighbours(1, np.pi / 2, np.pi / 2).tolist() == [8, 4, 0, -1, 1, 6, 9, -1]


def test_nested_indexing_round_trip_and_ring_is_a_permutation():
	for nside in (1, 2, 8, 64):
		pix = np.arange(12 * nside * nside)
		ix, iy, f = H.nest2xyf(nside, pix)
		assert (H.xyf2nest(nside, ix, iy, f) == pix).all()
		assert (np.sort(H.xyf2ring(nside, ix, iy, f)) == pix).all()


def test_neighbourhoods_are_symmetric_and_complete():
	for nside in (2, 4, 16, 32):
		npix = 12 * nside * nside
		nb = H.get_all_neighbours(nside, np.arange(npix), nest=True)
		assert nb.shape == (8, npix) and nb.max() < npix
		# exactly the 24 pixels at the eight vertices where only three base pixels meet have seven neighbours
		assert ((nb < 0).sum(axis=0) == 1).sum() == 24 and ((nb < 0).sum(axis=0) > 1).sum() == 0
		pairs = set()
		for m in range(8):
			ok = nb[m] >= 0
			pairs.update(zip(np.flatnonzero(ok).tolist(), nb[m][ok].tolist()))
		assert all((b, a) in pairs for a, b in pairs), 'neighbour relation is not symmetric'
		assert all(a != b for a, b in pairs)


def test_equal_area_and_locality():
	rng = np.random.default_rng(3)
	n = 400000
	theta = np.arccos(rng.uniform(-1, 1, n))
	phi = rng.uniform(0, 2 * np.pi, n)
	for nside in (1, 4, 16):
		npix = 12 * nside * nside
		pix = H.ang2pix(nside, theta, phi, nest=True)
		assert pix.min() >= 0 and pix.max() < npix
		counts = np.bincount(pix, minlength=npix)
		expect = n / npix
		assert (np.abs(counts - expect) < 6 * np.sqrt(expect) + 1).all(), 'pixels do not have equal areas'
	# a point moved by much less than a pixel stays in its pixel or lands in one of its eight neighbours
	nside = 64
	step = 0.2 * H.nside2resol(nside)
	theta2 = np.clip(theta + step * rng.normal(size=n) / 3, 0, np.pi)
	phi2 = phi + step * rng.normal(size=n) / 3 / np.maximum(np.sin(theta), 1e-3)
	a = H.ang2pix(nside, theta, phi, nest=True)
	b = H.ang2pix(nside, theta2, np.mod(phi2, 2 * np.pi), nest=True)
	nb = H.get_all_neighbours(nside, theta, phi, nest=True)
	near = (a == b) | (nb == b[None, :]).any(axis=0)
	far_ok = np.sin(theta) * np.abs(phi2 - phi) + np.abs(theta2 - theta) > 0.9 * H.nside2resol(nside)   # (none expected)
	assert (near | far_ok).all()
	# and the centre pixel of get_all_neighbours(theta, phi) is ang2pix(theta, phi): no neighbour equals it
	assert not (nb == a[None, :]).any()
