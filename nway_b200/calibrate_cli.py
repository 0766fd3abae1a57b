"""Command lines of the calibration helpers, with the reference's arguments and printed lines
(nway-create-shifted-catalogue.py, nway-create-fake-catalogue.py, nway-calibrate-cutoff.py of the reference); the work
is in nway_b200/calibrate.py (every collision search is one pass of the GPU match path).  The same-named scripts at
the repository root call these."""
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
