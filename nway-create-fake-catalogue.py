#!/usr/bin/env python
"""Create a fake, random-position catalogue for testing the false association rate (arguments of the reference's
nway-create-fake-catalogue.py).  For each source, a new position is drawn on the great arc towards one of its nearest
neighbours (with 2/3 probability one of the 10 nearest, else one of the 100 nearest); positions within --radius (arcsec)
of an old or new source are drawn again.  The collision searches run on the GPU, see nway_b200/calibrate.py.

Example: nway-create-fake-catalogue.py --radius 20 COSMOS-XMM.fits fake-COSMOS-XMM.fits
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main(argv=None):
	from nway_b200 import calibrate, fitsio
	from nway_b200.cli import get_tablekeys, PrintLogger
	parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
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


if __name__ == '__main__':
	sys.exit(main())
