"""CPU: the host layers above the C ABI, end to end, with the oracle standing in for the library's numeric stages
(tests/oraclectx.py installs a stand-in for nway_b200._lib.Context -- TEST INFRASTRUCTURE, never the product path).

What this covers that the other CPU tests do not: the orchestration of nway_b200.nway_match (scalar tables, installation of
user histograms, the host half of the automatic histograms, the order of the stages, truncation, the logger, exceptions,
the DataFrame), and nway_b200/cli.py from argv to the FITS table -- compared with the oracle AND with the committed
outputs of the unmodified reference (tests/golden/ref_*.npz, ref_cli_*.npz), the same files the GPU tests use.
The CUDA kernels are not involved here; their parity is tests/test_gpu_*.py."""
import os

import numpy as np
import pytest

from tests import cases, cliparity, oraclectx, parity


@pytest.fixture
def hostctx(monkeypatch):
	import nway_b200
	ctx = oraclectx.OracleContext()
	monkeypatch.setattr(nway_b200._lib, 'get_context', lambda device=None: ctx)
	return ctx


class Recorder(object):
	def __init__(self):
		self.lines, self.warnings = [], []

	def log(self, *msg):
		self.lines.append(' '.join(str(m) for m in msg))

	def warn(self, msg):
		self.warnings.append(msg)


def run_host(tables, radius, completeness, **kw):
	import nway_b200
	kw.setdefault('logger', nway_b200.NullOutputLogger())
	kw.setdefault('store_mag_hists', False)
	kw.setdefault('as_frame', False)
	return nway_b200.nway_match(tables, radius, completeness, **kw)


@pytest.mark.parametrize('name', ['cosmos2_magradius', 'cosmos3_magauto', 'syn2_maghist', 'syn3_pcvec', 'syn4_minprob', 'offeq3_south'])
def test_nway_match_host_logic_against_oracle_and_reference(name, hostctx):
	"""automatic histograms by radius and by posterior (two columns), user histograms, completeness vector, truncation,
	a flat field off the equator: the product's host code around the stand-in == the oracle's own nway_match == the
	unmodified reference's committed output"""
	from oracle import nway_oracle as O
	spec = cases.GOLDEN_CASES[name]
	kw = spec.get('kwargs', {})
	tables = cases.build_case(name)
	kept = [[np.array(m) for m in t['mags']] for t in tables]
	got = run_host(tables, spec['radius'], spec['completeness'], **kw)
	for t, mags in zip(tables, kept):   # unlike the reference (__init__.py:318-319) the caller's magnitudes are left alone
		assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(t['mags'], mags))
	ref = O.nway_match(cases.build_case(name), spec['radius'], spec['completeness'], **kw)
	cols = [c for c in ref if not c.startswith('_')]
	assert list(got.keys()) == cols
	for c in cols:
		assert got[c].dtype == (np.int64 if ref[c].dtype.kind in 'iu' else np.float64), c
	parity.assert_tables_match(ref, got, columns=cols, context=name)
	parity.check_against_digest(name, got, [t['name'] for t in tables])
	auto = any(h is None for t in tables for h in t['maghists'])
	want = ['match(fuse_final=0)', 'finalize'] if auto else ['match(fuse_final=1)']
	assert hostctx.calls == want + (['truncate'] if kw.get('min_prob', 0) > 0 else [])


def test_log_lines_files_and_frame(hostctx, tmp_path, monkeypatch):
	"""the messages of nwaylib.nway_match in its order (__init__.py:86,188,207,221,312,363,368,400-401,468), the
	*_fit.txt tables (:369-373) and the returned frame"""
	import pandas
	monkeypatch.chdir(tmp_path)
	log = Recorder()
	tables = cases.cosmos_subset(3, mags=True)
	frame = run_host(tables, 20, 0.9, mag_include_radius=25.0, min_prob=0.001, logger=log, store_mag_hists=True, as_frame=True)
	assert isinstance(frame, pandas.DataFrame)
	assert list(frame.columns[:3]) == ['XMM', 'OPT', 'IRAC'] and list(frame.columns[-3:]) == ['match_flag', 'prob_has_match', 'prob_this_match']
	assert frame['XMM'].dtype == np.int64 and frame['match_flag'].dtype == np.int64 and frame['ncat'].dtype == np.int64
	assert frame['Separation_XMM_OPT'].dtype == np.float64 and not (frame['prob_this_match'] < 0.001).any()
	assert log.warnings == ['WARNING: magnitude radius is very large (>= matching radius). Consider using a smaller value.']
	want = ['Primary catalogue "XMM" (1797), density gives 3.71e+07 objects on entire sky',
		'Computing distance-based probabilities ...',
		'matching: 387601 matches after filtering by search radius',
		'Incorporating bias "OPT:MAG" ...',
		'magnitude histogram stored to "OPT_MAG_fit.txt".',
		'Incorporating bias "IRAC:mag_ch1" ...',
		'magnitude histogram stored to "IRAC_mag_ch1_fit.txt".',
		'',
		'Computing final probabilities ...']
	found = [l for l in log.lines if l in want]
	assert found == want, log.lines
	counts = [l for l in log.lines if l.startswith('magnitude histogram of column')]
	assert len(counts) == 2 and counts[0].startswith('magnitude histogram of column "OPT_MAG": ') and ' secure matches, ' in counts[0]
	assert log.lines[-1].startswith('    cutting away ') and log.lines[-1].endswith(' (below p_i minimum)')
	for col in ('OPT_MAG', 'IRAC_mag_ch1'):
		text = open('%s_fit.txt' % col).read().splitlines()
		assert text[0] == '# lo hi selected others'
		rows = np.loadtxt('%s_fit.txt' % col)
		assert 2 <= len(rows) <= 17 and rows.shape[1] == 4 and (rows[1:, 0] == rows[:-1, 1]).all()
		assert all(len(l) == 43 for l in text[1:])   # four '%10.5f' fields


def test_errors_before_and_after_the_match(hostctx):
	import nway_b200
	far = cases.uniform_patch(1, (20, 200), (1.0, 0.5), 0.01)
	far[1]['ra'] += 5.0
	alone = run_host(far, 5.0, 0.9)   # no secondary in reach: every primary keeps its no-counterpart row (fastskymatch.py:178-181)
	assert np.array_equal(alone['A'], np.arange(20)) and (alone['B'] == -1).all() and (alone['prob_has_match'] == 0).all()
	with pytest.raises(nway_b200.EmptyResultException):   # only a table without primaries is empty (__init__.py:92-93)
		run_host(far, 5.0, 0.9, primary_range=(5, 0))
	empty = run_host(far, 5.0, 0.9, primary_range=(5, 0), allow_empty=True)
	assert list(empty.keys())[:2] == ['A', 'B'] and all(len(v) == 0 for v in empty.values())
	few = cases.with_mags(cases.uniform_patch(2, (50, 1500), (1.0, 0.3), 0.02), 4, cats=(1,), hist=False)
	with pytest.raises(nway_b200.UndersampledException) as info:
		run_host(few, 5.0, 0.9, mag_include_radius=1.0)
	assert 'too few secure matches' in str(info.value)
	ok = cases.uniform_patch(3, (30, 300, 300), (1.0, 0.5, 0.5), 0.01)
	with pytest.raises(Exception) as info:
		run_host(ok, 5.0, [1.0, 0.9])
	assert 'Prior completeness needs one value per catalog' in str(info.value)
	with pytest.raises(ValueError):
		run_host(ok, 5.0, 0.9, unrelated_mode='both')
	with pytest.raises(ValueError):
		run_host(ok[:1], 5.0, 0.9)
	many = cases.with_mags(cases.uniform_patch(3, (30, 300, 300), (1.0, 0.5, 0.5), 0.01), 5, cats=(1, 2), ncols=5)
	with pytest.raises(ValueError) as info:
		run_host(many, 5.0, 0.9)
	assert 'magnitude columns' in str(info.value)
	wide = cases.with_mags(cases.uniform_patch(3, (30, 300), (1.0, 0.5), 0.01), 5, cats=(1,))
	wide[1]['maghists'] = [cases.fixed_hist(1, nbins=65)]
	with pytest.raises(ValueError) as info:
		run_host(wide, 5.0, 0.9)
	assert '65 bins' in str(info.value)


def test_primary_ranges_concatenate_to_the_whole_table(hostctx):
	"""what nway_b200.parallel relies on: the rows of a block of primaries are the same rows the whole match holds"""
	spec = cases.GOLDEN_CASES['syn3']
	whole = run_host(cases.build_case('syn3'), spec['radius'], spec['completeness'])
	n0 = len(cases.build_case('syn3')[0]['ra'])
	parts = [run_host(cases.build_case('syn3'), spec['radius'], spec['completeness'], primary_range=r, allow_empty=True)
		for r in ((0, 170), (170, 1), (171, n0 - 171))]
	for c in whole:
		assert np.array_equal(np.concatenate([p[c] for p in parts]), whole[c], equal_nan=True), c
	assert hostctx.primary_range == (171, n0 - 171)
	run_host(cases.build_case('syn3'), spec['radius'], spec['completeness'])
	assert hostctx.primary_range == (0, -1)   # a plain call clears the range of the previous one


def run_cli_host(name, tmp_path, extra=()):
	from nway_b200 import cli, fitsio
	paths = cases.write_cosmos_subset_fits(str(tmp_path))
	out = str(tmp_path / (name + '.fits'))
	cwd = os.getcwd()
	os.chdir(str(tmp_path))
	try:
		rc = cli.main(cases.cli_args(name, paths, out) + list(extra))
	finally:
		os.chdir(cwd)
	assert rc == 0
	t = fitsio.read_table(out)
	cards, _ = fitsio._read_header(open(out, 'rb').read(), 0)
	return t, cards


@pytest.mark.parametrize('name', ['cli2', 'cli3', 'cli3_magauto', 'cli3_bayes', 'cli3_minprob', 'cli3_prefilter'])
def test_cli_from_argv_to_fits_against_reference_cli(name, tmp_path, hostctx, capsys):
	"""nway.py's layer (arguments, error columns, the merged input columns, column order and FITS formats, header keys,
	the FITS writer and reader) around the stand-in: bit for bit the table of the unmodified reference script -- and
	line for line what that script prints to stdout (tests/golden/ref_cli_stdout_*.txt, oracle/make_golden_cli.py)"""
	t, cards = run_cli_host(name, tmp_path)
	printed = capsys.readouterr().out
	assert cliparity.normalise_transcript(printed, str(tmp_path)) == cliparity.load_cli_transcript(name)
	got = {n: t.data[n] for n in t.columns}
	cliparity.check_against_cli_digest(name, got, exact=True, check_layout=True, formats=dict(zip(t.columns, t.formats)), header=cards)
	assert t.name == 'NWAYMATCH' and cards['METHOD'] == 'NWAY multi-way matching'
	assert cards['NWAYCMD'].startswith('nway.py --radius')
	assert '    writing "%s" (%d rows, %d columns) ...' % (str(tmp_path / (name + '.fits')), len(t), len(t.columns)) in printed.splitlines()


def test_calibration_helpers_from_argv_to_fits(tmp_path, hostctx, capsys):
	"""nway-create-shifted-catalogue.py / nway-create-fake-catalogue.py (nway_b200/calibrate_cli.py, calibrate.py): the
	reference's published numbers for the shift (doc/logs/XMM-shift: 561 removed, 1236 left) and the guarantees of the fake
	catalogue -- same size and columns, every new position at least the radius away from every original and every other new
	one, inside the footprint"""
	from nway_b200 import calibrate_cli, fitsio
	from oracle import nway_oracle as O
	paths = cases.write_cosmos_subset_fits(str(tmp_path))   # the XMM catalogue is complete in the subset
	out = str(tmp_path / 'XMM-shift.fits')
	assert calibrate_cli.shifted_main(['--radius', '40', '--shift-ra', '60', paths['XMM'], out]) == 0
	printed = capsys.readouterr().out.splitlines()
	assert printed == ['opening ' + paths['XMM'], '    using RA  column: RA', '    using DEC column: DEC',
		'removed 561 sources which collide with original positions', 'writing "%s" (1236 rows)' % out]
	t0, t1 = fitsio.read_table(paths['XMM']), fitsio.read_table(out)
	assert len(t1) == 1236 and t1.columns == t0.columns and t1.formats == t0.formats and t1.header['SKYAREA'] == 2.0
	assert np.isin(t1.data['ID'], t0.data['ID']).all() and (t1.data['RA'] > t0.data['RA'].min()).all()

	out = str(tmp_path / 'XMM-fake.fits')
	assert calibrate_cli.fake_main(['--radius', '40', '--seed', '7', paths['XMM'], out]) == 0
	t1 = fitsio.read_table(out)
	assert len(t1) == len(t0) and (t1.data['ID'] == t0.data['ID']).all() and (t1.data['pos_err'] == t0.data['pos_err']).all()
	ra, dec, fra, fdec = t0.data['RA'], t0.data['DEC'], t1.data['RA'], t1.data['DEC']
	assert (O.dist((fra[:, None], fdec[:, None]), (ra[None, :], dec[None, :])) * 3600).min() >= 40.0
	d_self = O.dist((fra[:, None], fdec[:, None]), (fra[None, :], fdec[None, :])) * 3600
	np.fill_diagonal(d_self, 1e9)
	assert d_self.min() >= 40.0
	assert fra.min() >= ra.min() and fra.max() <= ra.max() and fdec.min() >= dec.min() and fdec.max() <= dec.max()
	assert (np.abs(fra - ra) + np.abs(fdec - dec) > 0).all()


@pytest.mark.parametrize('nfiles', [2, 4])
def test_match_multiple_and_crossproduct_mirrors(nfiles, tmp_path, hostctx):
	"""the reference's own test of match_multiple (tests/fastskymatch_test.py:31-72,109-119: seeded uniform float32
	catalogues on [0, 1] deg, err = 0.03 deg, through FITS files) on nway_b200.fastskymatch's mirror: the merged table's
	layout and, beyond the reference's `len > 20`, the oracle's row set; crossproduct() returns the same rows"""
	from nway_b200 import fitsio
	from nway_b200.fastskymatch import match_multiple, crossproduct, healpix_nside_for, get_healpix_resolution_degrees
	from nway_b200.logger import NullOutputLogger
	from oracle import nway_oracle as O
	np.random.seed(0)
	files = []
	for i in range(nfiles):
		ra, dec = np.random.uniform(size=40), np.random.uniform(size=40)
		files.append(str(tmp_path / ('test_input_%d.fits' % i)))
		fitsio.write_table(files[-1], [fitsio.Column('ra', 'E', ra), fitsio.Column('dec', 'E', dec)], 'test_input_%d' % i)
	tabs = [fitsio.read_table(f) for f in files]
	names = [t.name for t in tabs]
	err = 0.03
	results, columns, header = match_multiple([t.data for t in tabs], names, err, [t.formats for t in tabs], logger=NullOutputLogger())
	assert [c.name for c in columns][:2 * nfiles] == ['%s_%s' % (n, k) for n in names for k in ('ra', 'dec')]
	assert [c.name for c in columns][-2:] == ['Separation_max', 'ncat'] and [c.format for c in columns][-2:] == ['E', 'I']
	assert 'Separation_%s_%s' % (names[1], names[0]) in [c.name for c in columns]   # later catalogue first (fastskymatch.py:298)
	assert header == dict(COLS_RA=' '.join('%s_ra' % n for n in names), COLS_DEC=' '.join('%s_dec' % n for n in names))
	radec = [(x.data['ra'].astype(float), x.data['dec'].astype(float)) for x in tabs]
	mt = O.create_match_table([dict(ra=r, dec=d, error=np.ones(len(r))) for r, d in radec], err * 3600)
	assert len(results) == len(mt['idx']) > 20
	for c, name in enumerate(names):
		assert (results[name] == mt['idx'][:, c]).all()
	merged = {c.name: c.array for c in columns}
	absent = mt['idx'][:, 1] == -1
	assert absent.any() and (merged['%s_ra' % names[1]][absent] == -99).all()   # fastskymatch.py:271-279
	assert (merged['%s_ra' % names[1]][~absent] == tabs[1].data['ra'][mt['idx'][~absent, 1]]).all()
	assert (crossproduct(radec, err) == mt['idx']).all()
	assert healpix_nside_for(15. / 3600) == 8192 and healpix_nside_for(20. / 3600) == 4096   # doc/matching.rst:196
	assert get_healpix_resolution_degrees(8192) >= 15. / 3600 > get_healpix_resolution_degrees(16384)


def test_cli_error_specifications(tmp_path, hostctx, capsys):
	"""`:major:minor:angle`, `:ra_err:dec_err` and a fixed number as position errors (nway.py:25-98): the error triples the
	command-line layer builds from the FITS columns, the `_ra` / `_dec` offset columns it adds to the table
	(fastskymatch.py:299-331) and the lines it prints about them"""
	from nway_b200 import cli, fitsio
	from oracle import nway_oracle as O
	rng = np.random.default_rng(11)
	tabs = cases.uniform_patch(3, (400, 6000, 5000), (1.0, 0.3, 0.5), 0.05)
	n0, n1 = len(tabs[0]['ra']), len(tabs[1]['ra'])
	maj = rng.uniform(0.5, 3, n0); mnr = rng.uniform(0.2, 1, n0) * maj; ang = rng.uniform(0, 180, n0)
	era = rng.uniform(0.2, 0.6, n1); edec = rng.uniform(0.2, 0.6, n1)
	files = []
	for t, extra in zip(tabs, ([('emaj', 'D', maj), ('emin', 'D', mnr), ('eang', 'D', ang)], [('era', 'D', era), ('edec', 'D', edec)], [])):
		n = len(t['ra'])
		cols = [fitsio.Column('ID', 'K', np.arange(n) + 100), fitsio.Column('RA', 'D', t['ra']), fitsio.Column('DEC', 'D', t['dec'])]
		cols += [fitsio.Column(*e) for e in extra]
		files.append(str(tmp_path / ('%s.fits' % t['name'])))
		fitsio.write_table(files[-1], cols, t['name'], table_header=[('SKYAREA', t['area'])])
	out = str(tmp_path / 'ell.fits')
	assert cli.main(['--radius', '6', '--prior-completeness', '0.9', files[0], ':emaj:emin:eang', files[1], ':era:edec', files[2], '0.5', '--out', out]) == 0
	printed = capsys.readouterr().out
	assert '    Position error for "A": found column emaj (for ra_error): Values are [%f..%f]' % (maj.min(), maj.max()) in printed
	assert '    Position error for "A": found column eang (for ell_angle): Values are [%f..%f]' % (ang.min(), ang.max()) in printed
	assert '    Position error for "B": found column edec (for dec_error): Values are [%f..%f]' % (edec.min(), edec.max()) in printed
	assert '    Position error for "C": using fixed value 0.500000' in printed
	t = fitsio.read_table(out)
	names = [x['name'] for x in tabs]
	tabs[0]['error'] = tuple(O.ellipse_from_cli(maj, mnr, ang))
	tabs[1]['error'] = (era, edec, np.zeros(n1))
	tabs[2]['error'] = 0.5 * np.ones(len(tabs[2]['ra']))
	ref = O.nway_match(tabs, 6., 0.9, unrelated_mode='cli', cli_compat=True)
	assert len(t) == len(ref[names[0]])
	for nm in names:
		assert (np.where(ref[nm] >= 0, ref[nm] + 100, -99) == t.data[nm + '_ID']).all()
	for mine, theirs in (('p_any', 'prob_has_match'), ('p_i', 'prob_this_match'), ('dist_bayesfactor', 'dist_bayesfactor_uncorrected'),
			('dist_bayesfactor_corrected', 'dist_bayesfactor'), ('Separation_B_A', 'Separation_A_B')):
		assert np.array_equal(t.data[mine], ref[theirs].astype(np.float32), equal_nan=True), mine
	cards, _ = fitsio._read_header(open(out, 'rb').read(), 0)
	assert cards['COLS_ERR'] == 'A_:emaj:emin:eang B_:era:edec C_0.5'   # nway.py:541
	for b, a in ((1, 0), (2, 0), (2, 1)):
		k = 'Separation_%s_%s' % (names[b], names[a])
		i = t.columns.index(k)
		assert t.columns[i + 1:i + 3] == [k + '_ra', k + '_dec'] and t.formats[i:i + 3] == ['E', 'E', 'E']
		both = (ref[names[a]] >= 0) & (ref[names[b]] >= 0)
		sep = np.hypot(t.data[k + '_ra'][both].astype(float), t.data[k + '_dec'][both].astype(float))
		assert np.allclose(sep, t.data[k][both], rtol=1e-5, atol=1e-5)   # the offsets are the separation's two components
		assert np.isnan(t.data[k + '_ra'][~both]).all()


def test_explain_reproduces_the_published_known_answer(tmp_path, hostctx, capsys):
	"""nway-explain.py on the 3-catalogue COSMOS match: the reference publishes the result for XMM source 422
	(doc/logs/explain:1-19: p_any 0.41; p_i 0.71 / 0.21 / 0.04 / 0.03 / 0.01 with the catalogues involved).  The subset holds
	every source that can appear in that group (row 368 of the XMM catalogue: ID 369 in the subset's files); its SKYAREA is
	rescaled so that the source densities are those of the full catalogues (nway-write-header's keyword, set in place)"""
	from nway_b200 import calibrate_cli, cli, fitsio
	paths = cases.write_cosmos_subset_fits(str(tmp_path))
	z = np.load(os.path.join(cases.GOLDEN_DIR, 'cosmos_subset.npz'))
	for name, path in paths.items():
		fitsio.set_table_keywords(path, [('SKYAREA', 2.0 * len(z[name + '_ra']) / int(z[name + '_nfull']))])
	out = str(tmp_path / 'example3.fits')
	assert cli.main(cases.cli_args('cli3', paths, out)) == 0
	t = fitsio.read_table(out)
	capsys.readouterr()
	assert calibrate_cli.explain_main([out, '369']) == 0
	lines = capsys.readouterr().out.splitlines()
	assert lines[:6] == ['NWAY results for Source 369:', '', 'It is uncertain if this source has a counterpart (p_any=0.41)', '',
		'Assuming it has a counterpart, we have the following possible associations:', '']
	want = [('Association 1**[match_flag==1]: probability p_i=0.71 ', 'XMM-OPT-IRAC'), ('Association 2: probability p_i=0.21 ', 'XMM'),
		('Association 3: probability p_i=0.04 ', 'XMM--IRAC'), ('Association 4: probability p_i=0.03 ', 'XMM-OPT-'), ('Association 5: probability p_i=0.01 ', 'XMM-')]
	for k, (head, names) in enumerate(want):
		assert lines[6 + 3 * k:9 + 3 * k] == [head, '     Involved catalogues:  %s ' % names, '']
	assert sum(l.startswith('Association ') for l in lines) == 221 == int((t.data['XMM_ID'] == 369).sum())
	assert lines[-2] == 'Disclaimer: These results assume that the input (sky densities, positional errors, and priors) are correct.'
	assert calibrate_cli.explain_main([out, '99999']) == 1
	assert capsys.readouterr().out.strip() == 'ERROR: ID not found. Was searching for XMM_ID == 99999'


def test_reference_api_test_script(tmp_path, hostctx, monkeypatch):
	"""the reference's API test (nway-apitest.py:17-127) call for call: catalogues read from FITS files (float32 error and
	magnitude columns handed over as they are), two and three catalogues, automatic histograms by radius and by posterior,
	then the histogram files the previous call has written -- its assertions (37836 / 387601 rows, the columns a caller
	relies on) hold for the frames nway_b200.nway_match returns.  The committed subset of the demo catalogues keeps every
	source that can appear in these matches."""
	import nway_b200
	from nway_b200 import fitsio
	paths = cases.write_cosmos_subset_fits(str(tmp_path))
	monkeypatch.chdir(tmp_path)

	def table_from_fits(fitsname, poserr_value=None, area=None, magnitude_columns=[]):
		t = fitsio.read_table(fitsname)
		ra, dec = t.data['RA'], t.data['DEC']
		poserr = t.data['pos_err'] if 'pos_err' in t.columns else poserr_value * np.ones(len(ra))
		mags, maghists, magnames = [], [], []
		for col_name, magfile in magnitude_columns:
			mag_all = t.data[col_name].copy()
			mag_all[mag_all == -99] = np.nan
			mags.append(mag_all)
			magnames.append(col_name)
			maghists.append(None if magfile == 'auto' else tuple(np.loadtxt(magfile).transpose()))
		return dict(name=t.name, ra=ra, dec=dec, error=poserr, area=t.header['SKYAREA'] * 1.0 if area is None else area, mags=mags, maghists=maghists, magnames=magnames)

	minimum = ['Separation_max', 'ncat', 'dist_bayesfactor', 'dist_post', 'p_single', 'prob_has_match', 'prob_this_match']
	xmm = lambda: table_from_fits(paths['XMM'], area=2.0)
	opt = lambda mag=(): table_from_fits(paths['OPT'], poserr_value=0.1, area=2.0, magnitude_columns=[('MAG', m) for m in mag])
	irac = lambda mag=(): table_from_fits(paths['IRAC'], poserr_value=0.5, area=2.0, magnitude_columns=[('mag_ch1', m) for m in mag])
	quiet = nway_b200.NullOutputLogger()

	result = nway_b200.nway_match([xmm(), opt()], match_radius=20, prior_completeness=0.9, logger=quiet)
	assert len(result) == 37836 and all(c in result.columns for c in minimum + ['Separation_XMM_OPT'])
	result = nway_b200.nway_match([xmm(), opt(['auto'])], match_radius=20, prior_completeness=0.9, store_mag_hists=False, mag_include_radius=4.0, logger=quiet)
	assert len(result) == 37836 and all(c in result.columns for c in minimum + ['Separation_XMM_OPT', 'bias_OPT_MAG'])
	assert not os.path.exists('OPT_MAG_fit.txt')
	result = nway_b200.nway_match([xmm(), opt(['auto'])], match_radius=20, prior_completeness=0.9, mag_include_radius=4.0, logger=quiet)
	assert len(result) == 37836 and os.path.exists('OPT_MAG_fit.txt')
	extra = ['Separation_XMM_OPT', 'Separation_OPT_IRAC', 'Separation_XMM_IRAC', 'bias_OPT_MAG', 'bias_IRAC_mag_ch1']
	auto = nway_b200.nway_match([xmm(), opt(['auto']), irac(['auto'])], match_radius=20, prior_completeness=0.9, logger=quiet)
	assert len(auto) == 387601 and all(c in auto.columns for c in minimum + extra)
	assert os.path.exists('IRAC_mag_ch1_fit.txt')
	filed = nway_b200.nway_match([xmm(), opt(['OPT_MAG_fit.txt']), irac(['IRAC_mag_ch1_fit.txt'])], match_radius=20, prior_completeness=0.9, logger=quiet)
	assert len(filed) == 387601 and all(c in filed.columns for c in minimum + extra)
	# beyond the script: the same rows; the histograms read back from their five-decimal files are the ones of the run that
	# made them up to that rounding (a sparsely filled bin can move a single source's probability by several per cent)
	diff = np.abs(filed['prob_has_match'] - auto['prob_has_match'])
	assert (filed['XMM'] == auto['XMM']).all() and (filed['OPT'] == auto['OPT']).all() and np.median(diff) < 1e-4 and diff.max() < 0.2
