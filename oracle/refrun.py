"""Run the UNMODIFIED reference (nway 4.7.1): from /root/reference in the build container, or from the pip-installed
copy oracle/_ref/ (oracle/install_ref.py; git-ignored, shipped to the GPU box with the snapshot).

TEST / BENCH INFRASTRUCTURE ONLY.  This module exists to (a) validate oracle/nway_oracle.py
against the real reference, (b) generate the golden vectors committed under
tests/golden/ (see oracle/make_golden.py) and (c) time the reference's own CPU path for
bench.py --impl reference / cpu_baseline (from oracle/_ref only -- /root/reference does not
exist on the GPU box).  The product never imports it.

The reference cannot be imported as-is here because astropy, healpy and matplotlib
are not installed.  We register inert stub modules for
those three packages; the real nwaylib code then runs unchanged: fastskymatch.crossproduct in its flat-sky branch
(fastskymatch.py:94-98) needs nothing else, and for its HEALPix branch (:134-160) the three healpy functions it
calls are provided by oracle/healpix_nest.py (a restatement of the published pixelisation).
"""
import os
import sys
import types

import numpy

REFERENCE_ROOT = '/root/reference'                                     # the mounted source tree (build container only)
INSTALLED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')   # pip --target copy (oracle/install_ref.py)


def reference_available():
	"""the source tree with its demo catalogues and scripts (tests that need doc/ or nway.py check this)"""
	return os.path.isdir(os.path.join(REFERENCE_ROOT, 'nwaylib'))


def package_root():
	"""where the unmodified nwaylib package can be imported from: the mounted tree, else the installed copy, else None"""
	for root in (REFERENCE_ROOT, INSTALLED_ROOT):
		if os.path.isfile(os.path.join(root, 'nwaylib', '__init__.py')):
			return root
	return None


def _install_stubs():
	if 'astropy' in sys.modules and not getattr(sys.modules['astropy'], '_nwb_stub', False):
		return  # a real astropy is there, nothing to do

	def mod(name):
		m = types.ModuleType(name)
		m._nwb_stub = True
		sys.modules[name] = m
		return m

	astropy = mod('astropy')
	io = mod('astropy.io')
	fits = mod('astropy.io.fits')
	units = mod('astropy.units')
	coords = mod('astropy.coordinates')
	astropy.io = io
	io.fits = fits
	astropy.units = units
	astropy.coordinates = coords

	class BinTableHDU(object):
		@staticmethod
		def from_columns(*args, **kwargs):
			raise NotImplementedError('astropy stub')
	fits.BinTableHDU = BinTableHDU

	def writeto(filename, data, header=None, output_verify='exception', overwrite=False, checksum=False):
		raise NotImplementedError('astropy stub')
	fits.writeto = writeto

	class _NoAstropy(object):
		def __init__(self, *args, **kwargs):
			raise NotImplementedError('astropy stub')
	coords.SkyCoord = _NoAstropy
	coords.SkyOffsetFrame = _NoAstropy

	healpy = mod('healpy')
	pixelfunc = mod('healpy.pixelfunc')
	healpy.pixelfunc = pixelfunc
	# healpy is un-vendored and unpinned (pyproject.toml:53): its three calls on the path are restated in
	# oracle/healpix_nest.py (published HEALPix algorithm, checked against healpy's docstring examples), so that the
	# real crossproduct() can also take its HEALPix branch here (fastskymatch.py:134-160)
	from oracle import healpix_nest
	pixelfunc.nside2resol = healpix_nest.nside2resol
	pixelfunc.ang2pix = healpix_nest.ang2pix
	pixelfunc.get_all_neighbours = healpix_nest.get_all_neighbours

	mpl = mod('matplotlib')
	plt = mod('matplotlib.pyplot')
	mpl.pyplot = plt


_nwaylib = None


def load_reference(scratch_dir='/tmp/nwb_refrun'):
	"""import the real nwaylib (cwd is moved to a scratch dir: the reference creates
	./cache at import, fastskymatch.py:21-22)."""
	global _nwaylib
	if _nwaylib is not None:
		return _nwaylib
	root = package_root()
	assert root is not None, 'reference neither mounted at %s nor installed in %s' % (REFERENCE_ROOT, INSTALLED_ROOT)
	_install_stubs()
	os.makedirs(scratch_dir, exist_ok=True)
	os.chdir(scratch_dir)
	if root not in sys.path:
		sys.path.insert(0, root)
	import nwaylib
	import nwaylib.fastskymatch
	import nwaylib.bayesdistance
	import nwaylib.logger
	_nwaylib = nwaylib
	return nwaylib


def read_fits_table(path):
	"""Minimal FITS BINTABLE reader (ext 1): fixed-width big-endian rows, formats J I D E.
	Returns (dict of native-endian numpy columns, header dict)."""
	with open(path, 'rb') as f:
		data = f.read()
	pos = 0
	ihdu = 0
	while pos < len(data):
		cards = {}
		done = False
		while not done:
			block = data[pos:pos + 2880]
			pos += 2880
			for i in range(36):
				c = block[i * 80:(i + 1) * 80].decode('ascii')
				k = c[:8].strip()
				if k == 'END':
					done = True
					break
				if c[8:10] == '= ':
					v = c[10:]
					if v.lstrip().startswith("'"):
						v = v.lstrip()[1:]
						v = v[:v.index("'")].strip()
					else:
						v = v.split('/')[0].strip()
					cards[k] = v
		size = 0
		naxis = int(cards.get('NAXIS', '0'))
		if naxis > 0:
			size = abs(int(cards['BITPIX'])) // 8
			for i in range(1, naxis + 1):
				size *= int(cards['NAXIS%d' % i])
		size += int(cards.get('PCOUNT', '0'))
		if ihdu == 1:
			fmts = {'J': '>i4', 'I': '>i2', 'D': '>f8', 'E': '>f4', 'K': '>i8'}
			nf = int(cards['TFIELDS'])
			dt = numpy.dtype([(cards['TTYPE%d' % i], fmts[cards['TFORM%d' % i].lstrip('1')]) for i in range(1, nf + 1)])
			assert dt.itemsize == int(cards['NAXIS1'])
			nrows = int(cards['NAXIS2'])
			raw = numpy.frombuffer(data, dtype=dt, count=nrows, offset=pos)
			cols = {n: numpy.ascontiguousarray(raw[n].astype(raw[n].dtype.newbyteorder('='))) for n in dt.names}
			return cols, cards
		pos += (size + 2879) // 2880 * 2880
		ihdu += 1
	raise ValueError('no extension 1 in %s' % path)


def run_reference(match_tables, match_radius, prior_completeness, **kwargs):
	"""real nwaylib.nway_match; joblib disk cache bypassed (fastskymatch.py:91) so that
	repeated calls and timings are honest.  Returns a flat DataFrame (primary column restored
	from the index that pandas>=2.2 groupby.apply creates)."""
	nwaylib = load_reference()
	m = nwaylib.fastskymatch
	if hasattr(m.crossproduct, 'func'):
		m.crossproduct = m.crossproduct.func
	kwargs.setdefault('logger', nwaylib.logger.NullOutputLogger())
	kwargs.setdefault('store_mag_hists', False)
	import contextlib, io
	with contextlib.redirect_stderr(io.StringIO()):  # tqdm bars bypass the logger
		res = nwaylib.nway_match(match_tables, match_radius, prior_completeness, **kwargs)
	return flatten_result(res, match_tables[0]['name'])


def flatten_result(res, primary_name):
	import pandas
	if primary_name not in res.columns:
		if isinstance(res.index, pandas.MultiIndex):
			prim = res.index.get_level_values(0).values
		else:
			prim = res.index.values
		res = res.reset_index(drop=True)
		res.insert(0, primary_name, prim)
	else:
		res = res.reset_index(drop=True)
	return res


def cosmos_tables(ncat=2, mags=False, root=os.path.join(REFERENCE_ROOT, 'doc')):
	"""The reference's own demo catalogues as nway_match dicts (nway-apitest.py:19-55)."""
	out = []
	spec = [('COSMOS_XMM.fits', None, []), ('COSMOS_OPTICAL.fits', 0.1, ['MAG']), ('COSMOS_IRAC.fits', 0.5, ['mag_ch1'])]
	for fname, poserr, magcols in spec[:ncat]:
		cols, hdr = read_fits_table(os.path.join(root, fname))
		n = len(cols['RA'])
		err = cols['pos_err'] if 'pos_err' in cols else poserr * numpy.ones(n)
		t = dict(name=hdr['EXTNAME'], ra=cols['RA'], dec=cols['DEC'], error=err, area=float(hdr['SKYAREA']),
			mags=[], maghists=[], magnames=[])
		if mags:
			for mc in magcols:
				mv = cols[mc].copy()
				mv[mv == -99] = numpy.nan
				t['mags'].append(mv)
				t['maghists'].append(None)
				t['magnames'].append(mc)
		out.append(t)
	return out
