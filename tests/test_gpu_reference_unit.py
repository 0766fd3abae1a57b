"""The reference's own unit tests (tests/bayesdistance_test.py, tests/fastskymatch_test.py) run against the GPU mirrors
of the functions they exercise (nway_b200.bayesdistance / nway_b200.fastskymatch)."""
import numpy
import numpy.testing as test
import pytest

pytestmark = pytest.mark.gpu
log10e = numpy.log10(numpy.e)
L = numpy.log(3600 * 180 / numpy.pi)


def log_bf2(psi, s1, s2):
	"""closed form for two catalogues (bayesdistance.py:42-49)"""
	s = s1 * s1 + s2 * s2
	return (numpy.log(2) + 2 * L - numpy.log(s) - psi * psi / 2 / s) * log10e


def log_bf3(p12, p23, p31, s1, s2, s3):
	"""closed form for three catalogues (bayesdistance.py:52-61)"""
	ss1, ss2, ss3 = s1 * s1, s2 * s2, s3 * s3
	s = ss1 * ss2 + ss2 * ss3 + ss3 * ss1
	q = ss3 * p12**2 + ss1 * p23**2 + ss2 * p31**2
	return numpy.log10(4) + 4 * L * log10e - numpy.log10(s) - q / 2 / s * log10e


def test_log_bf_consistent2():
	from nway_b200.bayesdistance import log_bf
	for psi in numpy.array([0., 0.1, 0.2, 0.3, 0.4, 0.5]):
		test.assert_almost_equal(log_bf2(psi, 0.1, 0.2), log_bf([[None, psi]], [0.1, 0.2]))


def test_log_bf_consistent3():
	from nway_b200.bayesdistance import log_bf
	sep = numpy.array([0., 0.1, 0.2, 0.3, 0.4, 0.5])
	for psi in sep:
		bf3 = log_bf3(psi, psi, psi, 0.1, 0.2, 0.3)
		g = log_bf([[None, psi, psi], [psi, None, psi], [psi, psi, None]], [0.1, 0.2, 0.3])
		test.assert_almost_equal(bf3, g)
	q = numpy.zeros(len(sep))
	g = log_bf([[numpy.nan + sep, sep, sep], [sep, numpy.nan + sep, sep], [sep, sep, numpy.nan + sep]], [0.1 + q, 0.2 + q, 0.3 + q])
	test.assert_almost_equal(g, log_bf3(sep, sep, sep, 0.1, 0.2, 0.3))


def test_ell_circ_consistent():
	"""tests/bayesdistance_test.py:149-203"""
	from nway_b200.bayesdistance import log_bf, log_bf_elliptical, convert_from_ellipse
	sigma1 = 100 * numpy.ones(1)
	sigma2 = 1 * numpy.ones(1)
	z = numpy.zeros(1)
	A = [sigma1, sigma1, z]
	B = [sigma2, sigma2, z]
	step = 1.0 * numpy.ones(1)
	test.assert_almost_equal(log_bf_elliptical([[None, z]], [[None, z]], [B, B]), log_bf([[None, z]], [sigma2, sigma2]), decimal=5)
	test.assert_almost_equal(log_bf_elliptical([[None, z]], [[None, z]], [A, A]), log_bf([[None, z]], [sigma1, sigma1]), decimal=5)
	test.assert_almost_equal(log_bf_elliptical([[None, step]], [[None, z]], [A, A]), log_bf([[None, step]], [sigma1, sigma1]))
	test.assert_almost_equal(log_bf_elliptical([[None, step]], [[None, z]], [B, B]), log_bf([[None, step]], [sigma2, sigma2]))
	test.assert_almost_equal(log_bf_elliptical([[None, step]], [[None, z]], [A, B]), log_bf([[None, step]], [sigma1, sigma2]))
	e1 = [numpy.atleast_1d(x) for x in convert_from_ellipse(0.1, 0.1, 0)]
	e2 = [numpy.atleast_1d(x) for x in convert_from_ellipse(0.2, 0.2, 0)]
	test.assert_almost_equal(log_bf_elliptical([[None, step]], [[None, z]], [e1, e2]), log_bf([[None, step]], [0.1 + z, 0.2 + z]), decimal=5)
	test.assert_almost_equal(log_bf_elliptical([[None, 2 * step]], [[None, z]], [e1, e2]), log_bf([[None, 2 * step]], [0.1 + z, 0.2 + z]), decimal=5)
	test.assert_almost_equal(log_bf_elliptical([[None, step]], [[None, step]], [e1, e2]), log_bf([[None, 2**0.5 * step]], [0.1 + z, 0.2 + z]))


def test_ell_against_oracle_random():
	from nway_b200.bayesdistance import log_bf_elliptical
	from oracle import nway_oracle as O
	rng = numpy.random.default_rng(4)
	n = 500
	for ncat in (2, 3, 4):
		errs = []
		for c in range(ncat):
			a = rng.uniform(0.3, 3, n); b = rng.uniform(0.1, 1, n) * a; phi = rng.uniform(0, numpy.pi, n)
			errs.append(tuple(O.convert_from_ellipse(a, b, phi)))
		sra = [[rng.normal(size=n) * 2 if i < j else None for j in range(ncat)] for i in range(ncat)]
		sde = [[rng.normal(size=n) * 2 if i < j else None for j in range(ncat)] for i in range(ncat)]
		got = log_bf_elliptical(sra, sde, errs)
		ref = O.log_bf_elliptical(sra, sde, errs)
		assert numpy.allclose(got, ref, rtol=1e-11, atol=1e-9)


def test_dist():
	"""tests/fastskymatch_test.py:16-29 plus the values of the real function"""
	from nway_b200.fastskymatch import dist
	d = dist((53.15964508, -27.92927742), (53.15953445, -27.9313736))
	assert not numpy.isnan(d) and abs(d - 0.002098457623965017) < 1e-15
	ra = numpy.array([53.14784241, 53.14784241, 53.14749908, 53.16559982, 53.19423676, 53.1336441])
	dec = numpy.array([-27.79363823, -27.79363823, -27.81790352, -27.79622459, -27.70860672, -27.76327515])
	ra2 = numpy.array([53.14907837, 53.14907837, 53.1498642, 53.16150284, 53.19681549, 53.13626862])
	dec2 = numpy.array([-27.79297447, -27.79297447, -27.81404877, -27.79223251, -27.71365929, -27.76314735])
	d = dist((ra, dec), (ra2, dec2))
	assert not numpy.isnan(d).any()
	from oracle import nway_oracle as O
	assert numpy.allclose(d, O.dist((ra, dec), (ra2, dec2)), rtol=1e-11, atol=0)


@pytest.mark.parametrize('nfiles', [2, 3, 4, 5])
def test_match_multiple_toy_catalogues(nfiles, tmp_path):
	"""tests/fastskymatch_test.py:31-72,109-119 (run_match): seeded uniform float32 catalogues on [0, 1] deg, err = 0.03 deg,
	through FITS files and match_multiple; the reference asserts only len > 20 -- here also the oracle's row set"""
	from nway_b200 import fitsio
	from nway_b200.fastskymatch import match_multiple, crossproduct, healpix_nside_for
	from nway_b200.logger import NullOutputLogger
	from oracle import nway_oracle as O
	numpy.random.seed(0)
	ngen = 40
	files = []
	for i in range(nfiles):
		ra = numpy.random.uniform(size=ngen)
		dec = numpy.random.uniform(size=ngen)
		path = str(tmp_path / ('test_input_%d.fits' % i))
		fitsio.write_table(path, [fitsio.Column('ra', 'E', ra), fitsio.Column('dec', 'E', dec)], 'test_input_%d' % i,
			primary_header=[('GENERAT', 'match test table, random')])
		files.append(path)
	tabs = [fitsio.read_table(f) for f in files]
	table_names = [t.name for t in tabs]
	err = 0.03
	results, columns, header = match_multiple([t.data for t in tabs], table_names, err, [t.formats for t in tabs], logger=NullOutputLogger())
	out = str(tmp_path / ('test_match%d.fits' % nfiles))
	fitsio.write_table(out, columns, 'MATCH', primary_header=[('ANALYSIS', 'match table from' + ', '.join(table_names)), ('INPUT', ', '.join(files))])
	t = fitsio.read_table(out)
	for name in table_names:
		ra, dec = t.data['%s_ra' % name], t.data['%s_dec' % name]
		assert len(ra) > 20 and len(dec) == len(ra)
	assert header['COLS_RA'] == ' '.join('%s_ra' % n for n in table_names)
	# the row set against the oracle (complete enumeration + radius filter)
	radec = [(x.data['ra'].astype(float), x.data['dec'].astype(float)) for x in tabs]
	mt = O.create_match_table([dict(ra=r, dec=d, error=numpy.ones(len(r))) for r, d in radec], err * 3600)
	assert len(results) == len(mt['idx'])
	for c, name in enumerate(table_names):
		assert (results[name] == mt['idx'][:, c]).all()
	assert numpy.allclose(t.data['Separation_max'], mt['sepmax'].astype(numpy.float32), rtol=3e-7)
	assert (t.data['ncat'] == mt['ncat']).all()
	assert (crossproduct(radec, err) == mt['idx']).all()
	assert healpix_nside_for(15. / 3600) == 8192 and healpix_nside_for(20. / 3600) == 4096   # doc/matching.rst:196
