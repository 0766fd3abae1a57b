"""CPU, build container only (needs the reference's demo catalogues under /root/reference/doc): the numbers the
reference publishes in doc/logs/ for its calibration workflow (doc/Makefile:41-73), reproduced by the oracle in its
command-line mode and by the host-side calibration code of the product (nway_b200/calibrate.py, pure numpy):

  doc/logs/XMM-shift:4-5      --radius 40 --shift-ra 60: 561 sources removed, 1236 remain
  doc/logs/match2-offset:30   24614 rows     doc/logs/match3-offset:32   220645 rows
  doc/logs/cutoff2            p_any cut-offs 0.82 / 0.77 / 0.74 / 0.67 with 9.35 / 22.43 / 30.05 / 47.86 % of the matches
  doc/logs/cutoff3            0.94 / 0.85 / 0.76 / 0.55 with 36.78 / 55.15 / 64.94 % (the last percentage is 78.13 in the log
                              and 78.19 here -- one of the 1797 sources sits on the 0.55 cut-off; the UNMODIFIED nway.py,
                              run in this container through oracle/refcli.py on the same two command lines, also gives
                              78.19: the published log predates the current code / numpy / scipy)
"""
import io
import os

import numpy as np
import pytest

from oracle import nway_oracle as O
from oracle import refrun

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(refrun.REFERENCE_ROOT, 'doc', 'COSMOS_OPTICAL.fits')),
	reason='the demo catalogues are only there in the build container')


def shifted(x):
	"""nway-create-shifted-catalogue.py:66-81 with --radius 40 --shift-ra 60"""
	ra, dec = x['ra'] + 60 / 60. / 60, x['dec'] + 0 / 60. / 60
	excluded = np.array([(O.dist((ra[i], dec[i]), (x['ra'], x['dec'])) * 60 * 60 < 40).any() for i in range(len(ra))])
	out = dict(x)
	out['ra'], out['dec'], out['error'] = ra[~excluded], dec[~excluded], np.asarray(x['error'])[~excluded]
	return out, int(excluded.sum())


def cli_run(tables, **kw):
	out = O.nway_match(tables, 20.0, 1.0, unrelated_mode='cli', cli_compat=True, **kw)   # nway.py defaults: completeness 1
	return out, dict(ncat=out['ncat'], p_any=out['prob_has_match'].astype(np.float32))   # the FITS column is 'E'


def through_fit_file(edges, hs, ha):
	"""the *_fit.txt round trip of nway.py:498-503,509-510 (five decimals)"""
	buf = io.BytesIO()
	buf.write(b'# lo hi selected others\n')
	np.savetxt(buf, np.transpose([edges[:-1], edges[1:], hs, ha]), fmt=["%10.5f"] * 4)
	buf.seek(0)
	return tuple(np.loadtxt(buf).transpose())


def test_shift_offset_counts_and_cutoff2():
	from nway_b200 import calibrate
	full = refrun.cosmos_tables(3, mags=False)
	xs, removed = shifted(full[0])
	assert removed == 561 and len(xs['ra']) == 1236
	real2, real2_cols = cli_run([dict(t) for t in full[:2]])
	fake2, fake2_cols = cli_run([dict(xs), dict(full[1])])
	assert len(real2['ncat']) == 37836 and len(fake2['ncat']) == 24614
	fake3, _ = cli_run([dict(xs), dict(full[1]), dict(full[2])])
	assert len(fake3['ncat']) == 220645
	lines = calibrate.calibrate_cutoff(real2_cols, fake2_cols)[3]
	want = open(os.path.join(refrun.REFERENCE_ROOT, 'doc', 'logs', 'cutoff2')).read().splitlines()[2:]
	assert [l for l in lines if l] == [l for l in want if l]


def test_cutoff3_with_automatic_and_file_histograms():
	from nway_b200 import calibrate
	full = refrun.cosmos_tables(3, mags=True)
	xs, _ = shifted(full[0])
	real, real_cols = cli_run([dict(t, mags=list(t['mags']), maghists=list(t['maghists'])) for t in full], mag_include_radius=4.0)
	assert len(real['ncat']) == 387601
	fake_tables = [dict(xs)] + [dict(t, mags=list(t['mags']), maghists=[through_fit_file(*real['_hists']['%s_%s' % (t['name'], m)])
		for m in t['magnames']]) for t in full[1:]]
	fake, fake_cols = cli_run(fake_tables)
	assert len(fake['ncat']) == 220645
	lines = [l for l in calibrate.calibrate_cutoff(real_cols, fake_cols)[3] if l]
	want = [l for l in open(os.path.join(refrun.REFERENCE_ROOT, 'doc', 'logs', 'cutoff3')).read().splitlines()[2:] if l]
	assert lines[:7] == want[:7]
	assert lines[7].startswith('--> use only counterparts with p_any>0.55 (78.1')
