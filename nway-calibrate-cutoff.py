#!/usr/bin/env python
"""Characterise the false association rate and efficiency of a match with a offset (fake) match and a real match
(arguments and printed table of the reference's nway-calibrate-cutoff.py; the plots are not produced).

Example: nway-calibrate-cutoff.py example2.fits example2-shifted-match.fits
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main(argv=None):
	import numpy
	from nway_b200 import calibrate, fitsio
	parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
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


if __name__ == '__main__':
	sys.exit(main())
