"""nway_b200 -- B200-native implementation of nway's match-probability path.

Drop-in for nwaylib.nway_match() (reference nwaylib/__init__.py:31-120): same arguments, same output columns,
computed by hand-written sm_100a CUDA kernels behind the C ABI of include/nwayb200.h.  There is no CPU path:
without the built library and a CUDA device, calls raise.
"""
from collections import OrderedDict

import numpy
from numpy import log10, pi

from . import _lib
from . import magnitudeweights
from .logger import NormalLogger, NullOutputLogger

__version__ = '0.1.0'


class UndersampledException(Exception):
	pass


class EmptyResultException(Exception):
	pass


default_logger = NormalLogger()


def _scalar_tables(match_tables, prior_completeness, logger):
	"""host scalars feeding the kernels, with the reference's expressions: source densities
	(__init__.py:199-217), completeness vector (:224-229), prior per presence pattern (:254), the Bayes-factor
	normalisation (bayesdistance.py:15,76) and the CLI sub-association prior (nway.py:395)."""
	ncats = len(match_tables)
	source_densities = []
	source_densities_plus = []
	area_total = (4 * pi * (180 / pi)**2)
	for i, t in enumerate(match_tables):
		n = len(t['ra'])
		area = t['area'] * 1.0
		density = n / area * area_total
		logger.log('%s "%s" (%d), density gives %.2e objects on entire sky' % ('Primary catalogue' if i == 0 else 'Catalogue', t['name'], n, density))
		source_densities.append(density)
		source_densities_plus.append((n + 1) / area * area_total)
	source_densities_plus[0] = source_densities[0]
	source_densities = numpy.array(source_densities)
	source_densities_plus = numpy.array(source_densities_plus)

	if numpy.shape(prior_completeness) == ():
		prior_completeness = numpy.array([1.0] + [float(prior_completeness)**(1. / (ncats - 1)) for i in range(1, ncats)])
	prior_completeness = numpy.asarray(prior_completeness, dtype=float)
	if len(prior_completeness) != ncats:
		raise Exception('Prior completeness needs one value per catalog. Received "%s".' % prior_completeness)
	assert prior_completeness[0] == 1.0

	nmask = 2**(ncats - 1)
	prior = numpy.empty(nmask)
	sub = numpy.ones(nmask)
	for mask in range(nmask):
		table_mask = numpy.array([True] + [(mask >> (c - 1)) & 1 == 1 for c in range(1, ncats)])
		prior[mask] = source_densities[0] * numpy.prod(prior_completeness[table_mask]) / numpy.prod(source_densities_plus[table_mask])
		aug = [c for c in range(1, ncats) if (mask >> (c - 1)) & 1]
		if aug:
			sub[mask] = source_densities[aug[0]] / numpy.prod(source_densities_plus[aug])
	assert numpy.isfinite(prior).all(), (source_densities, prior_completeness, source_densities_plus)
	log_arcsec2rad = numpy.log(3600 * 180 / pi)
	norm = numpy.array([(n - 1) * numpy.log(2) + 2 * (n - 1) * log_arcsec2rad for n in range(ncats + 1)])
	with numpy.errstate(divide='ignore'):   # an empty secondary catalogue has density 0: its sub-association prior is 0, log10 -inf
		sub_log10prior = log10(sub)
	return dict(pc=prior_completeness, norm=norm, log10e=float(log10(numpy.e)), prior=prior, log10prior=log10(prior),
		sub_log10prior=sub_log10prior)


def _is_triple(error):
	return isinstance(error, (tuple, list)) and len(error) == 3 and numpy.ndim(error[0]) == 1


def _error_triple(error, n):
	"""(sigma_ra, sigma_dec, rho) columns stacked for nwb_set_catalogue(err_kind=ELLIPSE); a circular error
	column becomes (sigma, sigma, 0) as in nway.py:79-88"""
	if _is_triple(error):
		cols = [numpy.asarray(x, dtype=float) for x in error]
	else:
		e = numpy.asarray(error, dtype=float)
		cols = [e, e, numpy.zeros(n)]
	assert all(len(x) == n for x in cols)
	return numpy.ascontiguousarray(numpy.stack(cols))


def convert_from_ellipse(a, b, phi):
	"""covariance parameters (sigma_x, sigma_y, rho) from major axis a, minor axis b, angle phi in radians
	(bayesdistance.py:190-204; host side: one pass over a catalogue column)"""
	a2 = a**2
	b2 = b**2
	s = numpy.sin(phi)
	c = numpy.cos(phi)
	s2 = s**2
	c2 = c**2
	sigma_x = (a2 * s2 + b2 * c2)**0.5
	sigma_y = (a2 * c2 + b2 * s2)**0.5
	rho = c * s * (a2 - b2) / (sigma_x * sigma_y)
	return sigma_x, sigma_y, rho


def ellipse_error(major, minor, angle_deg):
	"""the CLI's `:maj:min:angle` error specification as an `error` triple (nway.py:56-65)"""
	return convert_from_ellipse(numpy.asarray(major, dtype=float), numpy.asarray(minor, dtype=float),
		(numpy.asarray(angle_deg, dtype=float) - 90) / 180 * pi)


def _column_names(match_tables):
	names = [t['name'] for t in match_tables]
	n = len(names)
	seps = ['Separation_%s_%s' % (names[a], names[b]) for a in range(n) for b in range(a + 1, n)]
	biases = ['bias_%s_%s' % (t['name'], magname) for t in match_tables for magname in t.get('magnames', [])]
	return names, seps, biases


def _column_selectors(match_tables):
	"""output column name -> libnwayb200 column selector, in the reference's column order (__init__.py:131-196,
	100-103,111,392,405,415-419)"""
	L = _lib
	names, seps, biases = _column_names(match_tables)
	sel = OrderedDict()
	for c, name in enumerate(names):
		sel[name] = L.COL_IDX + c
	for k, name in enumerate(seps):
		sel[name] = L.COL_SEP + k
	sel['Separation_max'] = L.COL_SEPMAX
	sel['ncat'] = L.COL_NCAT
	sel['dist_bayesfactor_uncorrected'] = L.COL_LOGBF_UNCORR
	sel['dist_bayesfactor'] = L.COL_LOGBF
	sel['dist_post'] = L.COL_DIST_POST
	for k, name in enumerate(biases):
		sel[name] = L.COL_BIAS + k
	sel['p_single'] = L.COL_P_SINGLE
	sel['match_flag'] = L.COL_MATCH_FLAG
	sel['prob_has_match'] = L.COL_P_ANY
	sel['prob_this_match'] = L.COL_P_I
	return sel


def _fetch_table(ctx, match_tables, nrows):
	"""device columns -> OrderedDict of numpy arrays"""
	names = set(t['name'] for t in match_tables) | {'ncat', 'match_flag'}
	cols = OrderedDict()
	for name, sel in _column_selectors(match_tables).items():
		cols[name] = ctx.fetch(sel, nrows, numpy.int64 if name in names else numpy.float64)
	ctx.sync()
	return cols


def nway_match(match_tables, match_radius, prior_completeness,
	mag_include_radius=None, mag_exclude_radius=None, magauto_post_single_minvalue=0.9,
	prob_ratio_secondary=0.5,
	min_prob=0., consider_unrelated_associations=True,
	store_mag_hists=True,
	logger=default_logger,
	unrelated_mode='api', device=None, primary_range=None, as_frame=True, keep_on_device=False, allow_empty=False,
	cli_compat=False, pairwise_errs=(), flat_hash_compat=True, matcher=None, hist_rows=None):
	"""Same contract as nwaylib.nway_match (nwaylib/__init__.py:31-83); see there for the arguments.

	match_tables: list of dicts with name, ra, dec (deg), error (arcsec), area (deg^2), mags, magnames, maghists
	(None = build the histogram automatically; else (bins_lo, bins_hi, hist_sel, hist_all)).
	error may also be a triple (sigma_ra, sigma_dec, rho) of columns (see ellipse_error()): this switches the match
	to the elliptical Bayes factor of the reference's CLI (nway.py:346-354, bayesdistance.py:207-240).

	Extra keyword arguments (not in the reference):
	  unrelated_mode  'api' (default): the behaviour of the reference API, whose correction for unrelated
	                  associations is inert; 'cli': the live algorithm of nway.py:366-421.
	                  consider_unrelated_associations=False disables both.
	  device          CUDA device index (default: $NWB_DEVICE, $LOCAL_RANK or 0)
	  primary_range   (first, count): match only these primary rows (multi-GPU sharding)
	  as_frame        False returns an OrderedDict of numpy columns instead of a pandas.DataFrame
	  keep_on_device  leave the table in device memory and return {'nrows', 'selectors'} (used by nway_b200.parallel)
	  allow_empty     do not raise EmptyResultException for an empty shard
	  pairwise_errs   [(catalogue index a, catalogue index b, radius in arcsec), ...]: nway.py's --prefilter-pair as
	                  intended -- associations containing sources of both a and b are formed only if those two are
	                  closer than the radius (fastskymatch.py:184-208; the reference's code as written drops all of
	                  them, SURVEY.md Q8, which is radius 0 here)
	  flat_hash_compat  True (default): return exactly the reference's rows.  Where nwaylib takes its flat-sky hash (every
	                  catalogue at |dec| < 45 deg, away from ra = 0, radius < 1 deg: fastskymatch.py:94-98) it bins ra
	                  without cos(dec) and, away from the equator, never forms some pairs that ARE within the radius
	                  (SURVEY.md Q3); those associations are left out here too, so that prob_has_match / prob_this_match /
	                  match_flag of a group are the reference's.  False: the complete enumeration on the whole sphere
	                  (a superset; identical near the equator and wherever the reference uses its HEALPix hash).
	  hist_rows       callable(ctx, c) -> (nrows, res_ptr, sepmax_ptr, dist_post_ptr): the device columns the automatic
	                  histograms of catalogue c are selected from instead of this context's own rows (nway_b200.parallel:
	                  the rows of all shards, gathered)
	  matcher         callable(ctx, fuse_final) -> rows that runs the match in place of ctx.match (nway_b200.parallel: the
	                  multi-GPU mode in which the ranks share the streaming of the secondaries)
	  cli_compat      the arithmetic quirks of the command-line program nway.py on top of unrelated_mode='cli':
	                  separations (elliptical: offsets) pass through float32 before they are scored
	                  and automatic histograms take the weights of the selected rows (nway.py:471)

	Returns one row per association, ordered by primary index, then by the secondary indices with -1 first.
	The primary index is an ordinary column (pandas < 2.2 shape of the reference's frame)."""
	if mag_exclude_radius is None:
		mag_exclude_radius = mag_include_radius
	if mag_include_radius is not None:
		if mag_include_radius >= match_radius:
			logger.warn('WARNING: magnitude radius is very large (>= matching radius). Consider using a smaller value.')
	if unrelated_mode not in ('api', 'cli'):
		raise ValueError("unrelated_mode must be 'api' or 'cli'")

	ncats = len(match_tables)
	elliptical = any(_is_triple(t['error']) for t in match_tables)
	# limits of the device tables, checked before anything is uploaded
	if not 2 <= ncats <= _lib.MAX_CATALOGUES:
		raise ValueError('between 2 and %d catalogues can be matched, got %d' % (_lib.MAX_CATALOGUES, ncats))
	nmagcols = sum(len(t.get('mags', [])) for t in match_tables)
	if nmagcols > _lib.MAX_MAG_COLUMNS:
		raise ValueError('at most %d magnitude columns (all catalogues together) are supported, got %d' % (_lib.MAX_MAG_COLUMNS, nmagcols))
	for t in match_tables:
		for maghist, magname in zip(t.get('maghists', []), t.get('magnames', [])):
			if maghist is not None and len(maghist[0]) > _lib.MAX_HIST_BINS:
				raise ValueError('magnitude histogram for "%s_%s" has %d bins; at most %d are supported' % (t['name'], magname, len(maghist[0]), _lib.MAX_HIST_BINS))
	ctx = _lib.get_context(device)
	mag_columns = []   # (catalogue, k, values in the caller's dtype with -99 -> NaN, maghist, name)
	for c, t in enumerate(match_tables):
		mags = []
		for k, (magvals, maghist, magname) in enumerate(zip(t.get('mags', []), t.get('maghists', []), t.get('magnames', []))):
			magvals = numpy.array(magvals)
			if magvals.dtype.kind != 'f':
				magvals = magvals.astype(float)
			magvals[magvals == -99] = numpy.nan
			mags.append(magvals)
			mag_columns.append((c, k, magvals, maghist, magname))
		if elliptical:
			err = _error_triple(t['error'], len(t['ra']))
			ctx.set_catalogue(c, ncats, t['ra'], t['dec'], err, t['area'], mags=mags, err_kind=_lib.ERR_ELLIPSE)
		else:
			ctx.set_catalogue(c, ncats, t['ra'], t['dec'], t['error'], t['area'], mags=mags)
	if primary_range is not None:
		ctx.set_primary_range(*primary_range)
	else:
		ctx.set_primary_range(0, -1)

	tab = _scalar_tables(match_tables, prior_completeness, logger)
	mode = _lib.UNRELATED_CLI if (unrelated_mode == 'cli' and consider_unrelated_associations) else _lib.UNRELATED_API
	ctx.set_params(match_radius, tab['pc'], prob_ratio_secondary, mode)
	ctx.set_compat((_lib.COMPAT_SEP_F32 if cli_compat else 0) | (_lib.COMPAT_FLAT_HASH if flat_hash_compat else 0))
	ctx.set_prefilter(list(pairwise_errs))
	ctx.set_tables(tab['norm'], tab['log10e'], tab['prior'], tab['log10prior'], tab['sub_log10prior'])

	def install_hist(c, k, bins, hist_sel, hist_all):
		ctx.set_maghist(c, k, *magnitudeweights.step_tables(bins, hist_sel, hist_all))

	auto = [mc for mc in mag_columns if mc[3] is None]
	for c, k, magvals, maghist, magname in mag_columns:
		if maghist is not None:
			logger.log('magnitude histogramming: using user-supplied histogram for "%s_%s"' % (match_tables[c]['name'], magname))
			bins_lo, bins_hi, hist_sel, hist_all = maghist
			install_hist(c, k, numpy.array(list(bins_lo) + [bins_hi[-1]]), hist_sel, hist_all)

	logger.log('Computing distance-based probabilities ...')
	nrows = ctx.match(fuse_final=not auto) if matcher is None else matcher(ctx, not auto)
	logger.log('matching: %6d matches after filtering by search radius' % nrows)
	if not nrows > 0:
		if allow_empty:
			return dict(nrows=0, selectors=_column_selectors(match_tables)) if keep_on_device else OrderedDict(
				(k, numpy.zeros(0)) for k in _column_selectors(match_tables))
		raise EmptyResultException('No matches.')

	if auto:
		# first pass done; select the secure counterparts / field sources on the device and build the <= 17-bin tables
		# from the compact sample (__init__.py:324-375)
		for c, k, magvals, maghist, magname in auto:
			table_name = match_tables[c]['name']
			col = '%s_%s' % (table_name, magname)
			mag = '%s:%s' % (table_name, magname)
			logger.log('Incorporating bias "%s" ...' % mag)
			bins, hist_sel, hist_all, nsel, npossible, nothers = magnitudeweights.auto_histogram_device(ctx, c, k, magvals.dtype,
				mag_include_radius, mag_exclude_radius, magauto_post_single_minvalue, cli=cli_compat,
				rows=None if hist_rows is None else hist_rows(ctx, c))
			logger.log('magnitude histogram of column "%s": %d secure matches, %d insecure matches and %d secure non-matches of %d total entries (%d valid)' % (
				col, nsel, npossible, nothers, len(magvals), numpy.isfinite(magvals).sum()))
			if store_mag_hists:
				fname = mag.replace(':', '_') + '_fit.txt'
				logger.log('magnitude histogram stored to "%s".' % fname)
				with open(fname, 'wb') as f:
					f.write(b'# lo hi selected others\n')
					numpy.savetxt(f, numpy.transpose([bins[:-1], bins[1:], hist_sel, hist_all]), fmt=["%10.5f"] * 4)
			if nsel < 100:
				raise UndersampledException('ERROR: too few secure matches (%d) to make a good histogram. If you are sure you want to use this poorly sampled histogram, replace "auto" with the filename. You can also decrease the mag-auto-minprob parameter.' % nsel)
			install_hist(c, k, bins, hist_sel, hist_all)
		logger.log('')
		logger.log('Computing final probabilities ...')
		ctx.finalize()

	if min_prob > 0:
		kept = ctx.truncate(min_prob)
		logger.log('    cutting away %d (below p_i minimum)' % (nrows - kept))
		nrows = kept

	if keep_on_device:
		return dict(nrows=nrows, selectors=_column_selectors(match_tables))
	cols = _fetch_table(ctx, match_tables, nrows)
	if not as_frame:
		return cols
	import pandas
	# the columns are fresh arrays nobody else holds: hand them to the frame as they are (the default consolidates them
	# into per-dtype 2-D blocks -- for the 387 601 x 27 table of the COSMOS 3-catalogue match that copy alone costs more
	# than everything else in this function)
	try:
		return pandas.DataFrame(cols, copy=False)
	except TypeError:   # pandas < 1.3
		return pandas.DataFrame(cols)
