"""Magnitude-prior histograms.  The per-row lookup runs on the GPU (nwb_set_maghist); so does the selection of the
secure counterparts / field sources for the automatic histograms and the counting of the field sources
(auto_histogram_device: nwb_maghist_select / _count) -- what stays on the host is the <= 17-bin
table itself: the quantile bins of a sample of a few thousand magnitudes, with the reference's own numpy / scipy
calls so that the bin edges are bit-identical.  Mirrors nwaylib/magnitudeweights.py:18-23,74-118; the selection logic
of nwaylib/__init__.py:324-375 lives in the library (nwb_maghist_select)."""
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


def auto_histogram_device(ctx, c, k, magdtype, mag_include_radius, mag_exclude_radius, magauto_post_single_minvalue, cli=False, rows=None):
	"""The automatic histogram of one magnitude column (nwaylib/__init__.py:324-366, magnitudeweights.py:90-118) with the
	row / catalogue-sized work on the device: the selection of secure counterparts, first-occurrence unique, the
	reference's weight indexing and the field-source histogram run in nwb_maghist_select /
	nwb_maghist_count; the host only sees the compact sample of selected sources and builds the <= 17 quantile bins from
	it.  magdtype: dtype of the caller's magnitude column (the reference bins in that dtype: float32 quantile edges
	differ from float64 ones).  rows: select from these device columns (all shards' rows, see Context.maghist_select)
	instead of the context's own table.  Returns bins, hist_sel, hist_all, n_secure, n_possible, n_others."""
	if mag_include_radius is not None:
		mag_sel, w, (npossible, nothers, nvalid), (lo, hi) = ctx.maghist_select(c, k, True, mag_include_radius, mag_exclude_radius, cli, rows)
	else:
		mag_sel, w, (npossible, nothers, nvalid), (lo, hi) = ctx.maghist_select(c, k, False, magauto_post_single_minvalue, 0.01, cli, rows)
	assert len(mag_sel) > 0, 'No magnitude values within radius.'
	magdtype = numpy.dtype(magdtype)
	if magdtype.kind != 'f':
		magdtype = numpy.dtype(float)
	mag_sel = mag_sel.astype(magdtype)   # lossless: the device column holds the caller's values widened to fp64
	ok = ~numpy.logical_or(numpy.isnan(mag_sel), numpy.isinf(mag_sel))
	mag_sel, weights = mag_sel[ok], w[ok]
	# common adaptive binning: 15 quantile points of the weighted selected sample, extended to cover the field sources,
	# density-normalised (magnitudeweights.py:90-118); the field sources are counted on the device
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
