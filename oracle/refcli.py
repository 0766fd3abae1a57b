"""Run the UNMODIFIED reference command-line program (/root/reference/nway.py) in this container.

TEST INFRASTRUCTURE ONLY (build container only: needs /root/reference).  nway.py is a top-level script
built on astropy.io.fits, which is not installed here.  This module registers a small *functional* stand-in
for the handful of astropy.io.fits names the script and nwaylib.fastskymatch.match_multiple touch
(open, Column, ColDefs, BinTableHDU.from_columns, PrimaryHDU, HDUList.writeto) -- FITS files are read with
the independent reader of oracle/refrun.py, written tables are captured in memory -- and then executes the
real script with runpy.  Every number in the captured table is computed by the reference's own code.

Column semantics follow astropy: `Column(name, format, array)` converts the array to the format's numpy type
at construction (`_convert_to_valid_data_type` -> `astype`), i.e. 'E' columns hold float32 COPIES; this is
what makes the CLI score float32 separations and keeps `dist_bayesfactor` (uncorrected)
distinct from `dist_bayesfactor_corrected`.

Only circular position errors can run: the elliptical branch needs astropy.coordinates (dist3d,
fastskymatch.py:50-74).
"""
import contextlib
import io
import os
import runpy
import sys
from collections import OrderedDict

import numpy

from . import refrun

_FMT = {'L': numpy.bool_, 'B': numpy.uint8, 'I': numpy.int16, 'J': numpy.int32, 'K': numpy.int64, 'E': numpy.float32, 'D': numpy.float64}

WRITTEN = {}   # output file name -> dict(columns=OrderedDict name -> array, formats=..., primary_header=..., table_header=...)


def _np_type(fmt):
	fmt = str(fmt).strip().lstrip('0123456789')
	return _FMT[fmt[0]]


class Column(object):
	def __init__(self, name=None, format=None, array=None):
		self.name = name
		self.format = format
		self.array = numpy.asarray(array).astype(_np_type(format))   # astype: always a copy


class ColDefs(list):
	def __init__(self, cols):
		if isinstance(cols, BinTableHDU):
			cols = cols.columns
		list.__init__(self, cols)


class Header(OrderedDict):
	def add_comment(self, text):
		self.setdefault('COMMENT', []).append(text)


class BinTableHDU(object):
	def __init__(self, columns, name='', header=None):
		self.columns = ColDefs(columns)
		self.name = name
		self.header = Header(header or {})
		dt = numpy.dtype([(c.name, c.array.dtype) for c in self.columns])
		n = len(self.columns[0].array) if len(self.columns) else 0
		self.data = numpy.empty(n, dtype=dt)
		for c in self.columns:
			self.data[c.name] = c.array

	@staticmethod
	def from_columns(cols, **kwargs):
		return BinTableHDU(cols)


class PrimaryHDU(object):
	def __init__(self):
		self.header = Header()


class HDUList(list):
	def writeto(self, filename, overwrite=False, **kwargs):
		tb = self[1]
		WRITTEN[filename] = dict(
			columns=OrderedDict((c.name, numpy.array(tb.data[c.name])) for c in tb.columns),
			formats=OrderedDict((c.name, c.format) for c in tb.columns),
			primary_header=dict(self[0].header), table_header=dict(tb.header))


def _open(path):
	cols, cards = refrun.read_fits_table(path)
	nf = int(cards['TFIELDS'])
	names = [cards['TTYPE%d' % i] for i in range(1, nf + 1)]
	formats = [cards['TFORM%d' % i].strip() for i in range(1, nf + 1)]
	columns = [Column(n, f, cols[n]) for n, f in zip(names, formats)]
	hdr = {}
	for k, v in cards.items():
		try:
			hdr[k] = int(v)
		except ValueError:
			try:
				hdr[k] = float(v)
			except ValueError:
				hdr[k] = v
	hdu = BinTableHDU(columns, name=cards.get('EXTNAME', ''), header=hdr)
	return [PrimaryHDU(), hdu]


def _writeto(filename, data, header=None, output_verify='exception', overwrite=False, checksum=False):
	raise NotImplementedError('astropy stand-in: pyfits.writeto')


def install():
	"""refrun's inert stubs, then make astropy.io.fits functional and matplotlib.pyplot a sink"""
	if 'astropy.io.fits' not in sys.modules:
		refrun._install_stubs()
	fits = sys.modules['astropy.io.fits']
	fits.open = _open
	fits.Column = Column
	fits.ColDefs = ColDefs
	fits.BinTableHDU = BinTableHDU
	fits.PrimaryHDU = PrimaryHDU
	fits.HDUList = HDUList
	fits.writeto = _writeto
	plt = sys.modules['matplotlib.pyplot']

	def sink(*args, **kwargs):
		return None
	plt.__getattr__ = lambda name: sink


def run_cli(argv, cwd):
	"""execute /root/reference/nway.py with these arguments (catalogue paths relative to cwd or absolute).
	Returns (captured table dict, stdout text).  *_fit.txt histogram files land in cwd."""
	os.makedirs(cwd, exist_ok=True)
	refrun.load_reference(scratch_dir=cwd)   # registers the inert stubs, imports nwaylib (creates ./cache in cwd)
	install()
	import nwaylib.fastskymatch as m
	if hasattr(m.crossproduct, 'func'):
		m.crossproduct = m.crossproduct.func   # bypass the joblib disk cache
	# nwaylib.fastskymatch bound fits_from_columns at import time, from the inert stub: rebind it
	m.fits_from_columns = BinTableHDU.from_columns
	m.pyfits = sys.modules['astropy.io.fits']
	os.makedirs(cwd, exist_ok=True)
	old_cwd, old_argv = os.getcwd(), sys.argv
	out = io.StringIO()
	outfile = [a.split('=', 1)[1] for a in argv if a.startswith('--out=')] or [argv[argv.index('--out') + 1]]
	try:
		os.chdir(cwd)
		sys.argv = ['nway.py'] + list(argv)
		with contextlib.redirect_stdout(out), contextlib.redirect_stderr(io.StringIO()):
			runpy.run_path(os.path.join(refrun.REFERENCE_ROOT, 'nway.py'), run_name='__main__')
	finally:
		os.chdir(old_cwd)
		sys.argv = old_argv
	return WRITTEN[outfile[0]], out.getvalue()
