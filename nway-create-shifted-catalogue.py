#!/usr/bin/env python
"""Create a shifted catalogue for testing the false association rate (arguments of the reference's
nway-create-shifted-catalogue.py; the collision search runs on the GPU, see nway_b200/calibrate.py).

Example: nway-create-shifted-catalogue.py --radius 20 --shift-ra 0 --shift-dec 60 COSMOS-XMM.fits shifted-COSMOS-XMM.fits
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main(argv=None):
	from nway_b200 import calibrate, fitsio
	from nway_b200.cli import get_tablekeys
	parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
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


if __name__ == '__main__':
	sys.exit(main())
