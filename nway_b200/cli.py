"""Command-line surface of the reference (`nway.py`, 653 lines, reference @ /root/reference/nway.py) on top of the
B200 match path: same arguments, same output table (column names, order and FITS formats), same header keys.

	python nway.py --radius 20 --prior-completeness 0.9 COSMOS_XMM.fits :pos_err COSMOS_OPTICAL.fits 0.1 --out=out.fits

What runs where: FITS I/O, argument handling, the copy of the input columns into the output table and the <= 17-bin
magnitude histograms are host Python (the reference's own layer L3); candidate enumeration,
separations, Bayes factors, the unrelated-association correction, magnitude-prior lookup and the per-primary
normalisation run on the GPU through nway_b200.nway_match(unrelated_mode='cli', cli_compat=True), which reproduces the
arithmetic of the command-line program (float32 separation columns, live correction: SURVEY.md Q1/Q2/Q7).
There is no CPU path: without a CUDA device the match raises.
"""
from __future__ import division, print_function

import argparse
import sys
from collections import OrderedDict

import numpy
from numpy import pi

from . import fitsio

__doc_cli__ = """Multiway association between astrometric catalogue. Use --help for usage.

Example: nway.py --radius 10 --prior-completeness 0.95 --mag GOODS:mag_H auto --mag IRAC:mag_irac1 auto cdfs4Ms_srclist_v3.fits :Pos_error CANDELS_irac1.fits 0.5 gs_short.fits 0.1 --out=out.fits
"""


class HelpfulParser(argparse.ArgumentParser):
	def error(self, message):
		sys.stderr.write('error: %s\n' % message)
		self.print_help()
		sys.exit(2)


def build_parser():
	"""the reference's arguments, verbatim (nway.py:103-158)"""
	parser = HelpfulParser(description=__doc_cli__,
		epilog="""B200-native implementation of the nway match path; arguments of nway.py by Johannes Buchner (C) 2013-2025""",
		formatter_class=argparse.ArgumentDefaultsHelpFormatter)
	parser.add_argument('--radius', type=float, required=True,
		help='exclusive search radius in arcsec for initial matching')
	parser.add_argument('--mag-radius', default=None, type=float,
		help='search radius for building the magnitude histogram of target sources. If not set, the Bayesian posterior is used.')
	parser.add_argument('--mag-auto-minprob', default=0.9, type=float,
		help='minimum posterior probability (default: 0.9) for the magnitude histogram of secure target sources. Used in the Bayesian procedure.')
	parser.add_argument('--mag-exclude-radius', default=None, type=float,
		help='exclusion radius for building the magnitude histogram of field sources. If not set, --mag-radius is used.')
	parser.add_argument('--prior-completeness', metavar='COMPLETENESS', default="1", type=str,
		help='expected matching completeness of sources (prior)')
	parser.add_argument('--ignore-unrelated-associations', dest='consider_unrelated_associations', action='store_false',
		help='Ignore in the calculation source pairings unrelated to the primary source (not recommended)')
	parser.set_defaults(consider_unrelated_associations=True)
	parser.add_argument('--mag', metavar='MAGCOLUMN+MAGFILE', type=str, nargs=2, action='append', default=[],
		help="""name of <table>:<column> for magnitude biasing, and filename for magnitude histogram
		(use auto for auto-computation within mag-radius).
		Example: --mag GOODS:mag_H auto --mag IRAC:mag_irac1 irac_histogram.txt""")
	parser.add_argument('--acceptable-prob', metavar='PROB', type=float, default=0.5,
		help='ratio limit up to which secondary solutions are flagged')
	parser.add_argument('--min-prob', type=float, default=0,
		help='lowest probability allowed in final catalogue. If 0, no trimming is performed (default).')
	parser.add_argument('--out', metavar='OUTFILE', help='output file name', required=True)
	parser.add_argument('catalogues', type=str, nargs='+',
		help="""input catalogue fits files and position errors.

		Example: cdfs4Ms_srclist_v3.fits :Pos_error CANDELS_irac1.fits 0.5 gs_short.fits 0.1
		""")
	parser.add_argument('--prefilter-pair', metavar='CATNAME1 CATNAME2 radius', type=str, nargs=3, action='append', default=[],
		help="""name of two <table>s where combinations more distant than radius (in arcsec)
		should not considered. This reduces the memory needs when several large catalogs
		with high accuracy are matched against some with low accuracy (high --radius).

		Example: --prefilter-pair GOODS IRAC 0.1""")
	parser.add_argument('--prefilter-mode', choices=['fixed', 'reference'], default='reference',
		help='reference (default: the same rows as nway.py 4.7.1): every association containing both catalogues is dropped, whatever '
		'the radius -- what the reference\'s code does (its mask assignment has no effect, fastskymatch.py:203). '
		'fixed: --prefilter-pair as documented (drop associations whose two members are farther apart than the radius)')
	parser.add_argument('--device', type=int, default=None, help='CUDA device index (default: $NWB_DEVICE, $LOCAL_RANK or 0)')
	return parser


class PrintLogger(object):
	"""logger duck type of nwaylib.logger printing to stdout (the helper programs report this way)"""

	def log(self, *msg):
		print(' '.join(str(m) for m in msg))

	def warn(self, msg):
		print(msg)

	def progress(self, *args, **kwargs):
		from .logger import _Identity
		return _Identity()


class TranscriptLogger(object):
	"""logger duck type of nwaylib.logger for the match of the command-line program.  nway.py computes everything in its own
	script body and prints to stdout as it goes; here the same work is one nway_match() call that reports through its
	logger with the API's wording.  This logger keeps what belongs to the script's transcript -- per magnitude column the
	histogram lines (nway.py:495-497), the truncation line (:583) -- for main() to print in nway.py's order and wording,
	and passes the rest to stderr, where the reference's matching stage (fastskymatch.py:243-252,335 through
	logger.NormalLogger) writes too."""

	def __init__(self):
		self.column = None
		self.histogram_lines = OrderedDict()   # 'TABLE:column' -> lines
		self.truncation_line = None

	def log(self, *msg):
		msg = ' '.join(str(m) for m in msg)
		if msg.startswith('Incorporating bias "') and msg.endswith('" ...'):
			self.column = msg[len('Incorporating bias "'):-len('" ...')]
			self.histogram_lines[self.column] = []
		elif msg.startswith('magnitude histogram of column ') or msg.startswith('magnitude histogram stored to '):
			self.histogram_lines[self.column].append('    ' + msg)
		elif msg.lstrip().startswith('cutting away '):
			self.truncation_line = msg
		elif msg.startswith('matching: '):
			sys.stderr.write(msg + '\n')
		# anything else (densities, stage headings, user-supplied histograms) is printed by main() itself, as nway.py does

	def warn(self, msg):
		pass   # the one warning of the path (magnitude radius >= match radius) is printed per column by main(), nway.py:458-459

	def progress(self, *args, **kwargs):
		from .logger import _Identity
		return _Identity()


def offset_run_command(argv, first_catalogue, shiftfile, shiftoutfile):
	"""this run's command line turned into the one for the offset catalogue of the calibration recipe (nway.py:597-622):
	the primary catalogue replaced by the offset one, every `--mag T:col auto` by the histogram file this run has written,
	the output by its own file"""
	out = []
	k = 0
	while k < len(argv):
		word = argv[k]
		if word == first_catalogue:
			out.append(shiftfile)
		elif word == '--mag' and k + 2 < len(argv):
			column, histogram = argv[k + 1], argv[k + 2]
			out += [word, column, column.replace(':', '_') + '_fit.txt' if histogram == 'auto' else histogram]
			k += 2
		elif word == '--out' and k + 1 < len(argv):
			out += [word, shiftoutfile]
			k += 1
		elif word.startswith('--out='):
			out.append('--out=' + shiftoutfile)
		else:
			out.append(word)
		k += 1
	return ' '.join(out)


def get_tablekeys(names, name, tablename=''):
	"""the column called `name`, else the first one starting with it, else containing it (fastskymatch.py:77-80)"""
	keys = sorted(names, key=lambda k: 0 if k.upper() == name else 1 if k.upper().startswith(name) else 2)
	assert len(keys) > 0 and name in keys[0].upper(), 'ERROR: No "%s" column found in input catalogue "%s". Only have: %s' % (name, tablename, ', '.join(names))
	return keys[0]


def parse_error_spec(table, table_name, pos_error, match_radius_arcsec, n):
	"""one `errspec` argument -> (error for nway_match, is_simple).  A number or `:col` is a circular error (arcsec);
	`:ra_err:dec_err` an axis-aligned ellipse; `:major:minor:angle` a rotated one (nway.py:25-98, 276-305)."""
	from . import convert_from_ellipse
	colnames = table.columns
	if pos_error[0] != ':':
		print('    Position error for "%s": using fixed value %f' % (table_name, float(pos_error)))
		value = float(pos_error)
		if value > match_radius_arcsec:
			print('WARNING: Given separation error for "%s" is larger than the match radius! Increase --radius to >> %s' % (table_name, value))
		return value * numpy.ones(n), True
	keys = pos_error[1:].split(':')
	if len(keys) > 3:
		raise AssertionError('Invalid column specifier: %s' % pos_error)
	meanings = ['ra_error', 'dec_error', 'ell_angle']
	for k, meaning in zip(keys, meanings):
		assert k in colnames, 'ERROR: Position error column "%s" not in table "%s". Have these columns: %s' % (k, table_name, ', '.join(colnames))
		print('    Position error for "%s": found column %s (for %s): Values are [%f..%f]' % (
			table_name, k, meaning, table.data[k].min(), table.data[k].max()))
	cols = [numpy.asarray(table.data[k], dtype=float) for k in keys]
	if len(keys) == 3:
		era, edec, erho = convert_from_ellipse(cols[0], cols[1], (cols[2] - 90) / 180 * pi)
		error, simple = (era, edec, erho), False
	elif len(keys) == 2:
		era, edec = cols
		error, simple = (era, edec, numpy.zeros_like(era)), False
	else:
		era = edec = cols[0]
		error, simple = era, True
	if era.min() <= 0 or edec.min() <= 0:
		print('WARNING: Some separation errors in "%s" are 0! This will give invalid results (%d rows).' % (
			keys[0], numpy.logical_and(era <= 0, edec <= 0).sum()))
	if era.max() > match_radius_arcsec or edec.max() > match_radius_arcsec:
		print('WARNING: Some separation errors in "%s" are larger than the match radius! Increase --radius to >> %s' % (
			keys[0], max(era.max(), edec.max())))
	return error, simple


def merged_input_columns(tables, table_names, index_columns):
	"""every input column as `<table>_<column>` in its own FITS format, gathered by the row indices of the match
	table, -99 where the catalogue has no counterpart in that row (fastskymatch.py:262-283)"""
	out = []
	for table, table_name in zip(tables, table_names):
		idx = index_columns[table_name]
		missing = idx == -1
		take = numpy.where(missing, 0, idx)
		for n, fmt in zip(table.columns, table.formats):
			col = table.data[n][take] if len(table.data) else numpy.zeros(len(idx), dtype=table.data[n].dtype)
			if missing.any():
				col = col.copy()
				try:
					col[missing] = -99
				except Exception as e:
					print('   setting "%s_%s" to -99 failed (%d affected; column format "%s"): %s' % (table_name, n, missing.sum(), fmt, e))
			out.append(fitsio.Column('%s_%s' % (table_name, n), fmt, col))
	return out


def main(argv=None):
	import nway_b200
	from . import _lib

	parser = build_parser()
	args = parser.parse_args(argv)
	words = list(sys.argv) if argv is None else ['nway.py'] + list(argv)
	cmdline = ' '.join(words)

	print('NWAY arguments:')
	diff_secondary = args.acceptable_prob
	outfile = args.out
	filenames = args.catalogues[::2]
	print('    catalogues: ', ', '.join(filenames))
	pos_errors = args.catalogues[1::2]
	print('    position errors/columns: ', ', '.join(pos_errors))
	if len(pos_errors) != len(filenames):
		parser.error('every catalogue needs a position error (a number or :column)')
	if len(filenames) < 2:
		parser.error('need at least two catalogues')

	tables, table_names = [], []
	area_total = (4 * pi * (180 / pi)**2)
	for fitsname in filenames:
		t = fitsio.read_table(fitsname)
		tables.append(t)
		table_names.append(t.name)
		n = len(t)
		assert 'SKYAREA' in t.header, 'file "%s", table "%s" does not have a field "SKYAREA", which should contain the area of the catalogue in square degrees' % (fitsname, t.name)
		area = t.header['SKYAREA'] * 1.0
		print('      from catalogue "%s" (%d), density gives %.2e on entire sky' % (t.name, n, n / area * area_total))

	if ':' in args.prior_completeness:
		prior_completeness = numpy.array([1.0] + [float(pc) for pc in args.prior_completeness.split(':')])
		if len(prior_completeness) != len(filenames):
			raise Exception('Prior completeness needs one value per catalog, like "%s". Received "%s".' % (':'.join(["0.9"] * (len(filenames) - 1)), args.prior_completeness))
	else:
		prior_completeness = numpy.array([1.0] + [float(args.prior_completeness)**(1. / (len(filenames) - 1)) for i in range(1, len(filenames))])

	min_prob = args.min_prob
	match_radius = args.radius   # arcsec (nway_match's unit; nway.py:209 converts to degrees for match_multiple)

	pairwise_errs = [(table_names.index(tablea), table_names.index(tableb), float(err) if args.prefilter_mode == 'fixed' else 0.0)
		for tablea, tableb, err in args.prefilter_pair]
	if len(pairwise_errs) > 0:
		print('    pair-wise pre-filtering on')
		if args.prefilter_mode == 'reference':
			sys.stderr.write('NOTE: as in nway.py 4.7.1, --prefilter-pair removes EVERY association that contains both catalogues of a pair, whatever their separation; --prefilter-mode fixed applies the radius as documented\n')

	mag_include_radius = args.mag_radius
	mag_exclude_radius = args.mag_exclude_radius
	if mag_exclude_radius is None:
		mag_exclude_radius = mag_include_radius
	magauto_post_single_minvalue = args.mag_auto_minprob
	assert 0 < magauto_post_single_minvalue <= 1, 'probability should be between 0 and 1'

	magnitude_columns = args.mag
	print('    magnitude columns: ', ', '.join([c for c, _ in magnitude_columns]))
	for mag, magfile in magnitude_columns:
		table_name, col_name = mag.split(':', 1)
		assert table_name in table_names, 'table name specified for magnitude ("%s") unknown. Known tables: %s' % (table_name, ', '.join(table_names))
		ti = table_names.index(table_name)
		assert col_name in tables[ti].columns, 'column name specified for magnitude ("%s") unknown. Known columns in table "%s": %s' % (mag, table_name, ', '.join(tables[ti].columns))

	print('Computing distance-based probabilities ...')
	print('  finding position error columns ...')
	errors = []
	simple_errors = True
	for t, table_name, pos_error in zip(tables, table_names, pos_errors):
		error, simple = parse_error_spec(t, table_name, pos_error, match_radius, len(t))
		errors.append(error)
		simple_errors = simple_errors and simple

	print('  finding position columns ...')
	ra_keys = [get_tablekeys(t.columns, 'RA', tablename=n) for t, n in zip(tables, table_names)]
	sys.stderr.write('    using RA  columns: %s\n' % ', '.join(ra_keys))   # fastskymatch.py:249,251 (the matching stage's logger)
	dec_keys = [get_tablekeys(t.columns, 'DEC', tablename=n) for t, n in zip(tables, table_names)]
	sys.stderr.write('    using DEC columns: %s\n' % ', '.join(dec_keys))
	print('  building primary_id index ...')
	primary_id_key = get_tablekeys(tables[0].columns, 'ID', tablename=table_names[0])
	assert len(numpy.unique(tables[0].data[primary_id_key])) == len(tables[0].data[primary_id_key]), "ERROR: ID column '%s' in primary catalog contains duplicates." % primary_id_key
	primary_id_key = '%s_%s' % (table_names[0], primary_id_key)

	# ---- the match path (GPU) ------------------------------------------------------------------------------
	match_tables = []
	for ti, (t, table_name) in enumerate(zip(tables, table_names)):
		d = dict(name=table_name, ra=numpy.asarray(t.data[ra_keys[ti]], dtype=float), dec=numpy.asarray(t.data[dec_keys[ti]], dtype=float),
			error=errors[ti], area=t.header['SKYAREA'] * 1.0, mags=[], magnames=[], maghists=[])
		match_tables.append(d)
	for mag, magfile in magnitude_columns:
		table_name, col_name = mag.split(':', 1)
		d = match_tables[table_names.index(table_name)]
		d['mags'].append(tables[table_names.index(table_name)].data[col_name])
		d['magnames'].append(col_name)
		if magfile == 'auto':
			d['maghists'].append(None)
		else:
			d['maghists'].append(tuple(numpy.loadtxt(magfile).transpose()))

	print('  computing probabilities ...')
	logger = TranscriptLogger()

	def transcript(failed=False):
		"""what nway.py prints between its Bayes-factor loop and the output table (nway.py:364-583), in its order"""
		if args.consider_unrelated_associations:
			# the correction runs when some row lacks two or more catalogues (nway.py:366-368): with three or more
			# catalogues every primary's no-counterpart row does
			print('    correcting for unrelated associations ...' if len(tables) >= 3 else '      correcting for unrelated associations ... not necessary')
		if magnitude_columns:
			print()
			print('Incorporating magnitude biases ...')
		for mag, magfile in magnitude_columns:
			if magfile == 'auto' and mag not in logger.histogram_lines:
				break   # the run ended at an earlier column
			print('    magnitude bias "%s" ...' % mag)
			if magfile == 'auto':
				if mag_include_radius is not None and mag_include_radius >= match_radius:
					print('WARNING: magnitude radius is very large (>= matching radius). Consider using a smaller value.')
				print('\n'.join(logger.histogram_lines[mag]))
			else:
				print('    magnitude histogramming: using histogram from "%s" for column "%s"' % (magfile, mag.replace(':', '_')))
		if failed:
			return
		print()
		print('Computing final probabilities ...')
		print('    grouping by column "%s" and flagging ...' % (primary_id_key))
		if logger.truncation_line is not None:
			print(logger.truncation_line)

	try:
		cols = nway_b200.nway_match(match_tables, match_radius, prior_completeness,
			mag_include_radius=mag_include_radius, mag_exclude_radius=mag_exclude_radius,
			magauto_post_single_minvalue=magauto_post_single_minvalue, prob_ratio_secondary=diff_secondary,
			min_prob=min_prob, consider_unrelated_associations=args.consider_unrelated_associations,
			store_mag_hists=True, logger=logger, unrelated_mode='cli', cli_compat=True, as_frame=False,
			device=args.device, pairwise_errs=pairwise_errs)
	except nway_b200.EmptyResultException:
		raise AssertionError('No matches.')
	except nway_b200.UndersampledException:
		transcript(failed=True)
		print('ERROR: too few secure matches to make a good histogram. If you are sure you want to use this poorly sampled histogram, replace "auto" with the filename. You can also decrease the mag-auto-minprob parameter.')
		return 1
	transcript()
	nrows = len(cols[table_names[0]])
	ncats = len(tables)

	# ---- the output table, in the reference's column order (fastskymatch.py:262-342, nway.py:361-586) ------------
	columns = merged_input_columns(tables, table_names, cols)
	ctx = _lib.get_context(args.device)
	for i in range(ncats):
		for j in range(i):
			k = 'Separation_%s_%s' % (table_names[i], table_names[j])
			columns.append(fitsio.Column(k, 'E', cols['Separation_%s_%s' % (table_names[j], table_names[i])]))
			if not simple_errors:
				dra, ddec = ctx.row_offsets(j, i, nrows)
				columns.append(fitsio.Column(k + '_ra', 'E', dra))
				columns.append(fitsio.Column(k + '_dec', 'E', ddec))
	columns.append(fitsio.Column('Separation_max', 'E', cols['Separation_max']))
	columns.append(fitsio.Column('ncat', 'I', cols['ncat']))
	columns.append(fitsio.Column('dist_bayesfactor', 'E', cols['dist_bayesfactor_uncorrected']))
	if args.consider_unrelated_associations:
		# the column exists when some row lacks two or more catalogues (nway.py:366-421): with three or more
		# catalogues every primary's no-counterpart row does
		if ncats >= 3:
			columns.append(fitsio.Column('dist_bayesfactor_corrected', 'E', cols['dist_bayesfactor']))
	columns.append(fitsio.Column('dist_post', 'E', cols['dist_post']))
	biases = []
	for mag, magfile in magnitude_columns:
		col = mag.replace(':', '_')
		biases.append(col)
		columns.append(fitsio.Column('bias_%s' % col, 'E', cols['bias_%s' % col]))
	columns.append(fitsio.Column('p_single', 'E', cols['p_single']))
	columns.append(fitsio.Column('p_any', 'E', cols['prob_has_match']))
	columns.append(fitsio.Column('p_i', 'E', cols['prob_this_match']))
	columns.append(fitsio.Column('match_flag', 'I', cols['match_flag']))

	if not filenames[0].endswith('shifted.fits'):
		print()
		print()
		print('  You can calibrate a p_any cut-off with the following steps:')
		print('   1) Create a offset catalogue to simulate random sky positions:')
		shiftfile = filenames[0].replace('.fits', '').replace('.FITS', '') + '-fake.fits'
		shiftoutfile = outfile + '-fake.fits'
		print('      nway-create-fake-catalogue.py --radius %d %s %s' % (args.radius * 2, filenames[0], shiftfile))
		print('   2) Match the offset catalogue in the same way as this run:')
		print('      ' + offset_run_command(words, filenames[0], shiftfile, shiftoutfile))
		print('   3) determining the p_any cutoff that corresponds to a false-detection rate')
		print('      nway-calibrate-cutoff.py %s %s' % (outfile, shiftoutfile))
		print()

	print()
	print('creating output FITS file ...')
	primary_header = [('ANALYSIS', 'NWAY matching'), ('METHOD', 'NWAY multi-way matching'), ('INPUT', ', '.join(filenames)),
		('TABLES', ', '.join(table_names)), ('BIASING', ', '.join(biases)), ('NWAYCMD', cmdline),
		('COLS_RA', ' '.join(['%s_%s' % (ti, k) for ti, k in zip(table_names, ra_keys)])),
		('COLS_DEC', ' '.join(['%s_%s' % (ti, k) for ti, k in zip(table_names, dec_keys)])),
		('COL_PRIM', primary_id_key),
		('COLS_ERR', ' '.join(['%s_%s' % (ti, poscol) for ti, poscol in zip(table_names, pos_errors)]))]
	comments = ['argument %s: %s' % (k, v) for k, v in args.__dict__.items()]
	print('    writing "%s" (%d rows, %d columns) ...' % (outfile, nrows, len(columns)))
	fitsio.write_table(outfile, columns, 'NWAYMATCH', primary_header=primary_header, comments=comments)
	return 0


if __name__ == '__main__':
	sys.exit(main())
