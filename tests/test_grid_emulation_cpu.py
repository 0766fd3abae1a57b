"""CPU: completeness of the device's candidate search, checked on the host with the device's own source.

tests/emu/grid_emu.cpp compiles nwb_grid.cuh / nwb_grid_host.h / nwb_device.cuh -- the headers the CUDA library is
built from -- with a plain host compiler (tests/emu/nwb_host_emu.h supplies the handful of CUDA names they use) and
replays, one thread after the other, what decides which (primary, secondary) pairs ever reach the exact fp64 test:
grid geometry, registration of the primaries in the cells their search boxes overlap, the packed 8-byte entries inside
the cell records and the fp32 overflow entries, the cell a secondary falls into, and the two fp32 pre-tests.  A pair
within the radius that is not found this way would be a row the device silently loses; the exact test can only reject.

Random fields at the equator, at mid-latitudes, over both poles, across ra = 0 and on the whole sphere, radii from
1 arcsec to 40 arcmin, grids from one cell per primary to a handful of huge cells (crowded cells: overflow entries);
secondaries are planted at 0 ... 0.99999 of the radius around primaries, where a pre-test that is too tight fails
first.  No GPU needed."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import nway_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def emu(tmp_path_factory):
	out = str(tmp_path_factory.mktemp('emu') / 'grid_emu.so')
	cmd = ['g++', '-O2', '-ffp-contract=off', '-std=c++17', '-fPIC', '-shared', '-w', '-I', os.path.join(ROOT, 'tests', 'emu'),
		'-o', out, os.path.join(ROOT, 'tests', 'emu', 'grid_emu.cpp')]
	res = subprocess.run(cmd, capture_output=True, text=True)
	assert res.returncode == 0, res.stderr[-3000:]
	lib = ctypes.CDLL(out)
	P = ctypes.c_void_p
	lib.nwb_emu_check.restype = ctypes.c_longlong
	lib.nwb_emu_check.argtypes = [ctypes.c_int, P, P, ctypes.c_int, P, P, ctypes.c_double, ctypes.c_longlong, P, P, ctypes.c_longlong, P, P]
	return lib


def offset(ra, dec, dist_deg, bearing):
	"""destination point on the sphere (exact spherical trigonometry, so that the planted separations are what they say)"""
	la, lo, d = np.radians(dec), np.radians(ra), np.radians(dist_deg)
	la2 = np.arcsin(np.clip(np.sin(la) * np.cos(d) + np.cos(la) * np.sin(d) * np.cos(bearing), -1, 1))
	lo2 = lo + np.arctan2(np.sin(bearing) * np.sin(d) * np.cos(la), np.cos(d) - np.sin(la) * np.sin(la2))
	return np.degrees(lo2) % 360, np.degrees(la2)


def random_field(seed):
	rng = np.random.default_rng(seed)
	kind = rng.choice(['equator', 'mid', 'north', 'south', 'wrap', 'allsky', 'pole_exact'])
	radius = float(10 ** rng.uniform(0, 3.4))                       # 1 arcsec .. 2500 arcsec
	n0 = int(rng.integers(1, 3000))
	side = float(np.clip(radius / 3600 * rng.uniform(3, 300), 1e-3, 40))
	if kind == 'allsky':
		ra = 360 * rng.uniform(size=n0)
		dec = np.degrees(np.arcsin(2 * rng.uniform(size=n0) - 1))
	else:
		dec0 = dict(equator=0.0, mid=rng.uniform(-75, 75), north=90 - side * rng.uniform(0, 0.6), south=-90 + side * rng.uniform(0, 0.6),
			wrap=rng.uniform(-70, 70), pole_exact=rng.choice([90.0, -90.0]))[kind]
		ra0 = 360 - side / 3 if kind == 'wrap' else rng.uniform(0, 360)
		dec = dec0 + side * (rng.uniform(size=n0) - 0.5)
		dec = np.where(dec > 90, 180 - dec, np.where(dec < -90, -180 - dec, dec))
		cosd = max(np.cos(np.radians(min(abs(dec0) + side / 2, 89.9))), 0.02)
		ra = (ra0 + side * (rng.uniform(size=n0) - 0.5) / (cosd if rng.uniform() < 0.5 else 1.0)) % 360
		if kind == 'pole_exact':
			dec[0] = dec0   # a primary exactly on the pole
	# secondaries: planted around primaries at fractions of the radius (the interesting ones sit just inside)
	k = rng.integers(0, n0, 6 * n0 + 200)
	frac = rng.choice([0.0, 0.3, 0.9, 0.99, 0.999, 0.9999, 0.99999, 1.00001, 1.01], size=len(k))
	sra, sdec = offset(ra[k], dec[k], frac * radius / 3600, rng.uniform(0, 2 * np.pi, len(k)))
	max_cells = int(rng.choice([0, 0, 0, 64, 1024, 65536]))   # 0: the library's own choice
	return kind, radius, ra, dec, sra, sdec, max_cells


def run(emu, radius, ra, dec, sra, sdec, max_cells):
	lists = O.neighbour_lists([(ra, dec), (sra, sdec)], radius * (1 + 1e-9) / 3600)[0]   # complete, by the exact formula
	pi = np.concatenate([np.full(len(js), i, dtype=np.int32) for i, js in enumerate(lists)] + [np.zeros(0, dtype=np.int32)])
	si = np.concatenate([np.asarray(js, dtype=np.int32) for js in lists] + [np.zeros(0, dtype=np.int32)])
	stats = np.zeros(8, dtype=np.int64)
	first = np.full(2, -1, dtype=np.int32)
	arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (ra, dec, sra, sdec)]
	miss = emu.nwb_emu_check(len(ra), arrs[0].ctypes.data, arrs[1].ctypes.data, len(sra), arrs[2].ctypes.data, arrs[3].ctypes.data,
		float(radius), len(pi), pi.ctypes.data, si.ctypes.data, int(max_cells), stats.ctypes.data, first.ctypes.data)
	return miss, stats, first


@pytest.mark.parametrize('block', range(5))
def test_no_pair_within_the_radius_is_filtered_out(emu, block):
	# both ways of building the cell lists: count + place in one pass (steady state, even blocks) and count / headers /
	# fill (first match of a context, odd blocks)
	emu.nwb_emu_set_one_pass(1 - block % 2)
	total = inline = overflow = 0
	for seed in range(5000 + 16 * block, 5000 + 16 * (block + 1)):
		kind, radius, ra, dec, sra, sdec, max_cells = random_field(seed)
		miss, stats, first = run(emu, radius, ra, dec, sra, sdec, max_cells)
		assert miss == 0, 'seed %d (%s, r = %.4g arcsec, %d primaries, %d cells): %d of %d pairs lost, first: primary %d (%.8f, %.8f) secondary %d (%.8f, %.8f)' % (
			seed, kind, radius, len(ra), stats[3], miss, stats[0], first[0], ra[first[0]], dec[first[0]], first[1], sra[first[1]], sdec[first[1]])
		total += stats[0]; inline += stats[1]; overflow += stats[2]
	assert total > 10000 and inline > 0
	if block == 0:
		print('pairs within the radius: %d (inline entries %d, overflow entries %d)' % (total, inline, overflow))


def test_crowded_cells_use_the_overflow_entries(emu):
	"""many primaries in few cells: most entries are fp32 overflow entries (k1_pretest)"""
	rng = np.random.default_rng(1)
	ra = 150 + 0.05 * rng.uniform(size=4000)
	dec = 30 + 0.05 * rng.uniform(size=4000)
	k = rng.integers(0, 4000, 30000)
	sra, sdec = offset(ra[k], dec[k], rng.choice([0.5, 0.999, 0.99999], size=len(k)) * 3.0 / 3600, rng.uniform(0, 2 * np.pi, len(k)))
	for one_pass in (1, 0):
		emu.nwb_emu_set_one_pass(one_pass)
		miss, stats, first = run(emu, 3.0, ra, dec, sra, sdec, 64)
		assert miss == 0 and stats[2] > 10 * stats[1] > 0, (one_pass, miss, stats)
	emu.nwb_emu_set_one_pass(1)
