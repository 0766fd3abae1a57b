"""Magnitude-prior histograms.  The per-row lookup runs on the GPU (nwb_set_maghist); so does the selection of the
secure counterparts / field sources for the automatic histograms and the counting of the field sources
(auto_histogram_device: nwb_maghist_select / _count) -- what stays on the host is the <= 17-bin
table itself: the quantile bins of a sample of a few thousand magnitudes, with the reference's own numpy / scipy
calls so that the bin edges are bit-identical.  Mirrors nwaylib/magnitudeweights.py:18-23,74-118 and the selection
logic of nwaylib/__init__.py:324-375; auto_histogram is the all-host version of the same (kept for callers that hold
the columns on the host; the product path uses the device version)."""
import numpy
import scipy.interpolate


def ratio(hist_sel, hist_all):
	"""hist_sel / hist_all, 100 where hist_all is empty (magnitudeweights.py:18-23)"""
	hist_sel = numpy.asarray(hist_sel, dtype=float)
	hist_all = numpy.asarray(hist_all, dtype=float)
	with numpy.errstate(divide='ignore', invalid='ignore'):
		return numpy.where(hist_all == 0, 100, hist_sel / hist_all)


def step_tables(bins, hist_sel, hist_all):
	"""(edges, weight, bias) for nwb_set_maghist: weight = log10(ratio) with NaN -> 0 (__init__.py:386-388),
	bias = 10**weight (__init__.py:392), evaluated per bin with the same numpy calls the reference applies per row."""
	y = ratio(hist_sel, hist_all)
	with numpy.errstate(divide='ignore'):
		weight = numpy.log10(y)
	weight[numpy.isnan(weight)] = 0
	return numpy.asarray(bins, dtype=float), weight, 10 ** weight


def fitfunc_histogram(bin_mag, hist_sel, hist_all):
	"""the biasing function as a python callable (magnitudeweights.py:74-87), for callers that want it"""
	y = ratio(hist_sel, hist_all)
	return scipy.interpolate.interp1d(bin_mag, list(y) + [y[-1]], bounds_error=False, kind='zero')


def adaptive_histograms(mag_all, mag_sel, weights=None):
	"""common adaptive binning of the two samples: 15 quantile points of the (weighted) selected sample,
	extended to cover mag_all, density-normalised (magnitudeweights.py:90-118)"""
	if weights is None:
		weights = numpy.ones(len(mag_sel))
	assert len(weights) == len(mag_sel), (len(weights), len(mag_sel))
	order = numpy.argsort(mag_sel)
	sorted_sel = mag_sel[order]
	cum = numpy.cumsum(weights[order]) / numpy.sum(weights)
	cum[0] = 0
	cum[-1] = 1
	quantile = scipy.interpolate.interp1d(cum, sorted_sel)
	x = numpy.unique(quantile(numpy.linspace(0, 1, 15)))
	lo, hi = numpy.nanmin(mag_all), numpy.nanmax(mag_all)
	if x[-1] < hi:
		x = numpy.asarray(list(x) + [hi + 1])
	if x[0] > lo:
		x = numpy.asarray([lo - 1] + list(x))
	hist_sel, bins = numpy.histogram(mag_sel, bins=x, density=True, weights=weights)
	hist_all, bins = numpy.histogram(mag_all, bins=bins, density=True)
	return bins, hist_sel, hist_all


def auto_histogram(res, magvals, separation_max, dist_post, mag_include_radius, mag_exclude_radius,
		magauto_post_single_minvalue, cli=False):
	"""Select secure counterparts / secure field sources from the first pass and histogram their magnitudes
	(__init__.py:324-366).  res: index column of the catalogue; magvals: its magnitude column with NaN for
	undefined, in the caller's dtype.  Includes the reference's weight indexing so that results
	are identical: the API takes the weights of the rows with a counterpart (__init__.py:337), the command-line
	program those of the selected rows (nway.py:471) -- cli=True.
	Returns bins, hist_sel, hist_all, n_secure, n_possible, n_others."""
	res_defined = res != -1
	mask_all = numpy.isfinite(magvals)
	if mag_include_radius is not None:
		selection = separation_max < mag_include_radius
		selection_possible = separation_max < mag_exclude_radius
		selection_weights = numpy.ones(len(selection))
	else:
		selection = dist_post > magauto_post_single_minvalue
		selection_weights = dist_post
		selection_possible = dist_post > 0.01
	selection = numpy.logical_and(selection, res_defined)
	selection_weights = selection_weights[selection] if cli else selection_weights[res_defined]
	selection_possible = numpy.logical_and(selection_possible, res_defined)
	rows, unique_indices = numpy.unique(res[selection], return_index=True)
	rows_weights = selection_weights[unique_indices]
	assert len(rows) > 0, 'No magnitude values within radius.'
	mag_sel = magvals[rows]
	rows_possible = numpy.unique(res[selection_possible])
	mask_others = mask_all.copy()
	mask_others[rows_possible] = False
	mask_sel = ~numpy.logical_or(numpy.isnan(mag_sel), numpy.isinf(mag_sel))
	bins, hist_sel, hist_all = adaptive_histograms(magvals[mask_others], mag_sel[mask_sel], weights=rows_weights[mask_sel])
	return bins, hist_sel, hist_all, int(mask_sel.sum()), len(rows_possible), int(mask_others.sum())


def auto_histogram_device(ctx, c, k, magdtype, mag_include_radius, mag_exclude_radius, magauto_post_single_minvalue, cli=False):
	"""auto_histogram() with the row / catalogue-sized work on the device: the selection, first-occurrence unique,
	weight indexing and the field-source histogram run in nwb_maghist_select / nwb_maghist_count; the host only sees
	the compact sample of selected sources.  magdtype: dtype of the caller's magnitude column (the reference bins in
	that dtype: float32 quantile edges differ from float64 ones).  Same return value as auto_histogram."""
	if mag_include_radius is not None:
		mag_sel, w, (npossible, nothers, nvalid), (lo, hi) = ctx.maghist_select(c, k, True, mag_include_radius, mag_exclude_radius, cli)
	else:
		mag_sel, w, (npossible, nothers, nvalid), (lo, hi) = ctx.maghist_select(c, k, False, magauto_post_single_minvalue, 0.01, cli)
	assert len(mag_sel) > 0, 'No magnitude values within radius.'
	magdtype = numpy.dtype(magdtype)
	if magdtype.kind != 'f':
		magdtype = numpy.dtype(float)
	mag_sel = mag_sel.astype(magdtype)   # lossless: the device column holds the caller's values widened to fp64
	ok = ~numpy.logical_or(numpy.isnan(mag_sel), numpy.isinf(mag_sel))
	mag_sel, weights = mag_sel[ok], w[ok]
	# adaptive_histograms() with the histogram of the field sources counted on the device
	order = numpy.argsort(mag_sel)
	sorted_sel = mag_sel[order]
	cum = numpy.cumsum(weights[order]) / numpy.sum(weights)
	cum[0] = 0
	cum[-1] = 1
	quantile = scipy.interpolate.interp1d(cum, sorted_sel)
	x = numpy.unique(quantile(numpy.linspace(0, 1, 15)))
	lo, hi = magdtype.type(lo), magdtype.type(hi)
	if x[-1] < hi:
		x = numpy.asarray(list(x) + [hi + 1])
	if x[0] > lo:
		x = numpy.asarray([lo - 1] + list(x))
	hist_sel, bins = numpy.histogram(mag_sel, bins=x, density=True, weights=weights)
	n = ctx.maghist_count(c, k, bins)
	db = numpy.array(numpy.diff(bins), float)
	hist_all = n / db / n.sum()   # numpy.histogram(..., density=True)
	return bins, hist_sel, hist_all, int(ok.sum()), npossible, nothers
