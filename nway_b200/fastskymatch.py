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


def get_tablekeys(table, name, tablename=''):
	"""the column called `name`, else one starting with it, else one containing it (fastskymatch.py:77-80)"""
	from .cli import get_tablekeys as _g
	return _g(list(table.dtype.names), name, tablename=tablename)


def get_healpix_resolution_degrees(nside):
	"""0.7 x the HEALPix pixel resolution in degrees (fastskymatch.py:83-88).  The reference uses it to choose the
	nside of its HEALPix hash; this implementation bins on a band grid instead (the final association set does not
	depend on the hash) and keeps the function for callers."""
	resol = numpy.sqrt(4 * numpy.pi / (12. * nside * nside)) / numpy.pi * 180
	return 0.7 * resol


def healpix_nside_for(err):
	"""the nside the reference would pick for a search radius err in degrees (fastskymatch.py:102-116)"""
	nside = 1
	for nside_next in range(30):
		if get_healpix_resolution_degrees(2**nside_next) < err:
			break
		nside = 2**nside_next
	return nside


def _match_columns(radectables, err, pairwise_errs=(), device=None, names=None):
	from . import nway_match, NullOutputLogger, EmptyResultException
	names = names or ['T%d' % i for i in range(len(radectables))]
	tables = [dict(name=n, ra=numpy.asarray(ra, dtype=float), dec=numpy.asarray(dec, dtype=float), error=numpy.ones(len(ra)), area=1.0,
		mags=[], magnames=[], maghists=[]) for n, (ra, dec) in zip(names, radectables)]
	try:
		return nway_match(tables, err * 60 * 60, 1.0, logger=NullOutputLogger(), store_mag_hists=False, as_frame=False, device=device,
			pairwise_errs=[(a, b, e) for a, b, e in pairwise_errs]), names
	except EmptyResultException:
		return None, names


def crossproduct(radectables, err, logger=None, pairwise_errs=[], device=None):
	"""The associations of N catalogues within err (degrees): int64 array (R, N) of row indices, -1 = the catalogue
	takes no part, sorted like the reference's (fastskymatch.py:92-218).  The reference returns every tuple that shares
	a hash bucket and leaves the radius filter to its caller (__init__.py:180, fastskymatch.py:335); here the filter
	has already run (max pairwise separation < err), so this is the subset of the reference's array that survives it.
	pairwise_errs: [(i, j, radius in arcsec)], see nway_match."""
	cols, names = _match_columns(radectables, err, pairwise_errs, device)
	if cols is None:
		return numpy.zeros((0, len(radectables)), dtype=numpy.int64)
	return numpy.stack([cols[n] for n in names], axis=1)


def match_multiple(tables, table_names, err, fits_formats, logger, circular=True, pairwise_errs=[], device=None):
	"""fastskymatch.match_multiple (:228-342): all associations within err (degrees) of the FITS tables.
	Returns (results, cat_columns, header): results = structured array of row indices per table; cat_columns = the
	merged table as a list of nway_b200.fitsio.Column -- `<table>_<col>` copies of every input column (-99 where absent),
	`Separation_<b>_<a>` in arcsec ('E'; plus `_ra` / `_dec` offsets when circular=False), `Separation_max`, `ncat`;
	header = dict(COLS_RA, COLS_DEC)."""
	from . import _lib, fitsio
	from .cli import merged_input_columns
	logger.log('')
	logger.log('matching with %f arcsec radius' % (err * 60 * 60))
	logger.log('matching: %6d naive possibilities' % numpy.prod([float(len(t)) for t in tables]))
	ra_keys = [get_tablekeys(table, 'RA', tablename=tablename) for table, tablename in zip(tables, table_names)]
	logger.log('    using RA  columns: %s' % ', '.join(ra_keys))
	dec_keys = [get_tablekeys(table, 'DEC', tablename=tablename) for table, tablename in zip(tables, table_names)]
	logger.log('    using DEC columns: %s' % ', '.join(dec_keys))
	ratables = [(t[ra_key], t[dec_key]) for t, ra_key, dec_key in zip(tables, ra_keys, dec_keys)]
	cols, names = _match_columns(ratables, err, pairwise_errs, device, names=list(table_names))
	n = len(tables)
	if cols is None:
		cols = dict((nm, numpy.zeros(0, dtype=numpy.int64)) for nm in names)
		for a in range(n):
			for b in range(a + 1, n):
				cols['Separation_%s_%s' % (names[a], names[b])] = numpy.zeros(0)
		cols['Separation_max'] = numpy.zeros(0)
		cols['ncat'] = numpy.zeros(0, dtype=numpy.int64)
	nrows = len(cols[names[0]])
	results = numpy.empty(nrows, dtype=[(nm, numpy.int64) for nm in names])
	for nm in names:
		results[nm] = cols[nm]

	class _T(object):
		pass
	wrapped = []
	for t, fmts in zip(tables, fits_formats):
		w = _T()
		w.columns, w.formats, w.data = list(t.dtype.names), list(fmts), t
		wrapped.append(w)
	cat_columns = merged_input_columns(wrapped, names, cols)
	header = dict(
		COLS_RA=' '.join(["%s_%s" % (ti, ra_key) for ti, ra_key in zip(table_names, ra_keys)]),
		COLS_DEC=' '.join(["%s_%s" % (ti, dec_key) for ti, dec_key in zip(table_names, dec_keys)]))
	logger.log('    adding angular separation columns')
	ctx = _lib.get_context(device)
	for i in range(n):
		for j in range(i):
			k = "Separation_%s_%s" % (names[i], names[j])
			cat_columns.append(fitsio.Column(k, 'E', cols['Separation_%s_%s' % (names[j], names[i])]))
			if not circular:
				dra, ddec = ctx.row_offsets(j, i, nrows)
				cat_columns.append(fitsio.Column(k + '_ra', 'E', dra))
				cat_columns.append(fitsio.Column(k + '_dec', 'E', ddec))
	cat_columns.append(fitsio.Column('Separation_max', 'E', cols['Separation_max']))
	cat_columns.append(fitsio.Column('ncat', 'I', cols['ncat']))
	logger.log('matching: %6d matches after filtering by search radius' % nrows)
	logger.log('')
	return results, cat_columns, header
