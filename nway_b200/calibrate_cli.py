"""Command lines of the helper programs around a match, with the reference's arguments and printed lines: the
calibration loop (nway-create-shifted-catalogue.py, nway-create-fake-catalogue.py, nway-calibrate-cutoff.py; the work is in
nway_b200/calibrate.py, every collision search is one pass of the GPU match path) and the two pure FITS helpers a run
starts and ends with (nway-write-header.py: name and sky area of a catalogue; nway-explain.py: the associations of one
source, as text).  The same-named scripts at the repository root call these."""
import argparse

SHIFTED_MAIN_DOC = """Create a shifted catalogue for testing the false association rate (arguments of the reference's
nway-create-shifted-catalogue.py; the collision search runs on the GPU, see nway_b200/calibrate.py).

Example: nway-create-shifted-catalogue.py --radius 20 --shift-ra 0 --shift-dec 60 COSMOS-XMM.fits shifted-COSMOS-XMM.fits
"""


def shifted_main(argv=None):
	from nway_b200 import calibrate, fitsio
	from nway_b200.cli import get_tablekeys
	parser = argparse.ArgumentParser(description=SHIFTED_MAIN_DOC, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
	parser.add_argument('--shift-dec', default=0, type=float, help='Shift to add in dec (arcsec)')
	parser.add_argument('--shift-ra', default=0, type=float, help='Shift to add in ra (arcsec)')
	parser.add_argument('--radius', type=float, required=True, help='Remove sources which are near original sources, within this radius (arcsec).')
	parser.add_argument('inputfile', type=str, help='input catalogue fits file')
	parser.add_argument('outputfile', help='output catalogue fits file')
	args = parser.parse_args(argv)
	print('opening', args.inputfile)
	t = fitsio.read_table(args.inputfile)
	if args.shift_ra == 0 and args.shift_dec == 0:
		print('ERROR: You have to set either shift-ra or shift-dec to non-zero')
		return 1
	ra_key = get_tablekeys(t.columns, 'RA')
	print('    using RA  column: %s' % ra_key)
	dec_key = get_tablekeys(t.columns, 'DEC')
	print('    using DEC column: %s' % dec_key)
	ra, dec, excluded = calibrate.shifted_catalogue(t.data[ra_key], t.data[dec_key], args.shift_ra, args.shift_dec, args.radius)
	print('removed %d sources which collide with original positions' % (excluded.sum()))
	data = t.data.copy()
	data[ra_key] = ra
	data[dec_key] = dec
	cols = [fitsio.Column(n, f, data[n][~excluded]) for n, f in zip(t.columns, t.formats)]
	print('writing "%s" (%d rows)' % (args.outputfile, (~excluded).sum()))
	fitsio.write_table(args.outputfile, cols, t.name, table_header=fitsio.extra_header(t))
	return 0


FAKE_MAIN_DOC = """Create a fake, random-position catalogue for testing the false association rate (arguments of the reference's
nway-create-fake-catalogue.py).  For each source, a new position is drawn on the great arc towards one of its nearest
neighbours (with 2/3 probability one of the 10 nearest, else one of the 100 nearest); positions within --radius (arcsec)
of an old or new source are drawn again.  The collision searches run on the GPU, see nway_b200/calibrate.py.

Example: nway-create-fake-catalogue.py --radius 20 COSMOS-XMM.fits fake-COSMOS-XMM.fits
"""


def fake_main(argv=None):
	from nway_b200 import calibrate, fitsio
	from nway_b200.cli import get_tablekeys, PrintLogger
	parser = argparse.ArgumentParser(description=FAKE_MAIN_DOC, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
	parser.add_argument('--radius', type=float, required=True, help='Remove sources which are near original sources, within this radius (arcsec).')
	parser.add_argument('--seed', type=int, default=0, help='Seed for deterministic output.')
	parser.add_argument('inputfile', type=str, help='input catalogue fits file')
	parser.add_argument('outputfile', help='output catalogue fits file')
	args = parser.parse_args(argv)
	print('opening', args.inputfile)
	t = fitsio.read_table(args.inputfile)
	ra_key = get_tablekeys(t.columns, 'RA')
	print('    using RA  column: %s' % ra_key)
	dec_key = get_tablekeys(t.columns, 'DEC')
	print('    using DEC column: %s' % dec_key)
	ra, dec = calibrate.fake_catalogue(t.data[ra_key], t.data[dec_key], args.radius, seed=args.seed, logger=PrintLogger())
	data = t.data.copy()
	data[ra_key] = ra
	data[dec_key] = dec
	cols = [fitsio.Column(n, f, data[n]) for n, f in zip(t.columns, t.formats)]
	print('writing "%s" (%d rows)' % (args.outputfile, len(data)))
	fitsio.write_table(args.outputfile, cols, t.name, table_header=fitsio.extra_header(t))
	return 0


CUTOFF_MAIN_DOC = """Characterise the false association rate and efficiency of a match with a offset (fake) match and a real match
(arguments and printed table of the reference's nway-calibrate-cutoff.py; the plots are not produced).

Example: nway-calibrate-cutoff.py example2.fits example2-shifted-match.fits
"""


def cutoff_main(argv=None):
	import numpy
	from nway_b200 import calibrate, fitsio
	parser = argparse.ArgumentParser(description=CUTOFF_MAIN_DOC, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
	parser.add_argument('realfile', help='match output using real catalogue')
	parser.add_argument('fakefile', help='match output using fake catalogue')
	args = parser.parse_args(argv)
	real = fitsio.read_table(args.realfile).data
	fake = fitsio.read_table(args.fakefile).data
	cutoffs, efficiency, error_rate, lines = calibrate.calibrate_cutoff(real, fake)
	numpy.savetxt(args.realfile + '_p_any_cutoffquality.txt', numpy.transpose([cutoffs, efficiency, error_rate]),
		header='p_any_cutoff selection_efficiency false_selection_rate', fmt='%.6f')
	print('created table "%s_p_any_cutoffquality.txt"' % args.realfile)
	print('\n'.join(lines))
	return 0


def write_header_main(argv=None):
	"""nway-write-header.py <catalogue.fits> <tablename> <skyarea>: the two keywords nway.py needs in a catalogue
	(nway.py:176-191: the extension's name, SKYAREA in square degrees), set in place -- nothing else of the file changes"""
	import sys
	from nway_b200 import fitsio
	argv = sys.argv[1:] if argv is None else list(argv)
	if len(argv) != 3:
		sys.stderr.write("""SYNOPSIS: nway-write-header.py <catalogue.fits> <tablename> <skyarea>

tablename: name of the catalogue
skyarea: catalogue area in square degrees
""")
		return 1
	path, name, area = argv
	assert '_' not in name, 'Table name must not contain underscore "_".'
	t = fitsio.read_table(path)
	print('current', t.name, 'SKYAREA:', t.header.get('SKYAREA', None))
	fitsio.set_table_keywords(path, [('EXTNAME', name), ('SKYAREA', float(area))])
	t = fitsio.read_table(path)
	print('new    ', t.name, 'SKYAREA:', t.header.get('SKYAREA', None))
	return 0


EXPLAIN_MAIN_DOC = """Explain the associations of one primary source in a match table written by nway.py: whether it has a
counterpart, and every candidate association with its probability, the catalogues involved and the priors that moved it
(arguments and printed text of the reference's nway-explain.py; its two plots are not drawn).

Example: nway-explain.py example3.fits 422
"""


def explain_main(argv=None):
	import numpy
	from nway_b200 import fitsio
	parser = argparse.ArgumentParser(description=EXPLAIN_MAIN_DOC, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
	parser.add_argument('matchcatalogue', type=str, help='nway output catalogue')
	parser.add_argument('id', type=str, help='ID to explain (from primary catalogue)')
	args = parser.parse_args(argv)
	t = fitsio.read_table(args.matchcatalogue)
	header, _ = fitsio._read_header(fitsio._file_bytes(args.matchcatalogue), 0)
	data = t.data
	key = header['COL_PRIM']
	kind = data.dtype[key].kind
	wanted = int(args.id) if kind in 'iu' else float(args.id) if kind == 'f' else args.id.encode() if kind == 'S' else args.id
	rows = numpy.flatnonzero(data[key] == wanted)
	if len(rows) == 0:
		print('ERROR: ID not found. Was searching for %s == %s' % (key, args.id))
		return 1
	group = data[rows]
	p_any = group['p_any'][0]
	print('NWAY results for Source %s:' % args.id)
	print()
	if p_any > 0.8:
		print('This source probably has a counterpart (p_any=%.2f)' % p_any)
	elif p_any < 0.1:
		print('This source probably does not a counterpart (p_any=%.2f)' % p_any)
	else:
		print('It is uncertain if this source has a counterpart (p_any=%.2f)' % p_any)
	print()
	print("Assuming it has a counterpart, we have the following possible associations:")
	print()
	order = numpy.argsort(group['p_i'])[::-1]
	# per catalogue, the distinct positions among the group's rows in that order; "absent" (-99) comes first.  The script names
	# a catalogue in an association only when the row holds the FIRST of them (nway-explain.py:200-205), an empty name when
	# the catalogue is absent, nothing at all otherwise
	position_lists = []
	for col_ra, col_dec in zip(header['COLS_RA'].split(' '), header['COLS_DEC'].split(' ')):
		seen = [(-99, -99)]
		for i in order:
			here = (group[col_ra][i], group[col_dec][i])
			if here not in seen:
				seen.append(here)
		position_lists.append((col_ra, col_dec, seen))
	priors = [c for c in str(header.get('BIASING', '')).split(', ') if c.strip() != '']
	for number, i in enumerate(order, 1):
		parts = []
		for (col_ra, col_dec, seen), tablename in zip(position_lists, header['TABLES'].split(', ')):
			k = seen.index((group[col_ra][i], group[col_dec][i]))
			if k == 0:
				parts.append('')
			elif k == 1:
				parts.append(tablename)
		flag = group['match_flag'][i]
		if flag == 0:
			print('Association %d: probability p_i=%.2f ' % (number, group['p_i'][i]))
		else:
			stars = '' if p_any < 0.1 else '**' if flag == 1 else '*' if flag == 2 else ''
			print('Association %d%s[match_flag==%d]: probability p_i=%.2f ' % (number, stars, flag, group['p_i'][i]))
		print('     Involved catalogues:  %s ' % '-'.join(parts))
		for col in priors:
			bias = group['bias_' + col][i]
			if bias >= 2:
				print('     prior %-15s increased the probability (bias_%s=%.2f)' % (col, col, bias))
			elif bias <= 0.5:
				print('     prior %-15s decreased the probability (bias_%s=%.2f)' % (col, col, bias))
		print()
	print()
	print("Disclaimer: These results assume that the input (sky densities, positional errors, and priors) are correct.")
	print()
	return 0
