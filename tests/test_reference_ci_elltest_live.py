"""CPU, build container only (needs /root/reference/tests/elltest): the reference's own CI job "test elliptical errors"
(.github/workflows/tests.yml:125-135) with this repository's programs -- the four nway.py command lines on the catalogues the
reference ships for it (written by topcat / stilts, not by our FITS writer; circular, axis-aligned and rotated error
ellipses, two and three catalogues, --min-prob) each followed by nway-explain.py for source 95.  The reference's job only
checks that the programs run; here the circular case is also compared with the UNMODIFIED nway.py executed on the same files
(oracle/refcli.py: table bit for bit, stdout line for line), and the elliptical ones with the oracle (the unmodified
script cannot run them here: its offsets come from astropy, which is absent).  The oracle stands in for the library's
numeric stages (tests/oraclectx.py); tests/test_gpu_cli.py runs elliptical command lines on the device."""
import os
import shutil

import numpy as np
import pytest

from oracle import refrun
from tests import oraclectx

ELL = os.path.join(refrun.REFERENCE_ROOT, 'tests', 'elltest')
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(ELL, 'randomcatX.fits')), reason='the fixtures are only there in the build container')


@pytest.mark.timeout(900)
def test_elliptical_error_job(tmp_path, monkeypatch, capsys):
	import nway_b200
	from nway_b200 import calibrate_cli, cli, fitsio
	from oracle import nway_oracle as O
	from oracle import refcli
	ctx = oraclectx.OracleContext()
	monkeypatch.setattr(nway_b200._lib, 'get_context', lambda device=None: ctx)
	for f in ('randomcatX.fits', 'randomcatR.fits', 'randomcatO.fits'):
		shutil.copy(os.path.join(ELL, f), str(tmp_path / f))
	monkeypatch.chdir(tmp_path)
	x, r, o = (fitsio.read_table('randomcat%s.fits' % k) for k in 'XRO')
	assert (x.name, r.name, o.name) == ('CHANDRA', 'XMM', 'OPT') and x.header['SKYAREA'] == 0.01 and len(x) == 120 and len(o) == 129960
	assert x.columns == ['ID', 'RA', 'DEC', 'pos_err', 'a', 'b', 'phi']

	def job(argv, explain_id='95'):
		assert cli.main(argv) == 0
		printed = capsys.readouterr().out
		out = [a.split('=', 1)[1] for a in argv if a.startswith('--out=')][0]
		assert calibrate_cli.explain_main([out, explain_id]) == 0
		text = capsys.readouterr().out
		assert text.startswith('NWAY results for Source %s:' % explain_id) and 'Association 1' in text
		return fitsio.read_table(out), printed

	# ---- circular errors: against the unmodified script on the same files ------------------------------------------------
	argv = ['--radius=10.0', 'randomcatX.fits', ':pos_err', 'randomcatO.fits', '0.1', '--out=random_circtest.fits', '--min-prob=0.01']
	t, printed = job(argv)
	ref_dir = str(tmp_path / 'ref')
	os.makedirs(ref_dir)
	for f in ('randomcatX.fits', 'randomcatO.fits'):
		shutil.copy(os.path.join(ELL, f), os.path.join(ref_dir, f))
	ref, ref_stdout = refcli.run_cli(argv, ref_dir)
	os.chdir(str(tmp_path))   # the harness imports the reference from inside its scratch directory and may leave us there
	assert list(ref['columns']) == t.columns and [ref['formats'][n] for n in t.columns] == t.formats
	for n in t.columns:
		assert np.array_equal(np.asarray(ref['columns'][n]), t.data[n], equal_nan=True), n
	fix = lambda s: [l.replace(os.path.join(refrun.REFERENCE_ROOT, 'nway.py'), 'nway.py') for l in s.splitlines()]
	assert fix(printed) == fix(ref_stdout)
	assert not (t.data['p_i'] < 0.01).any() and 'cutting away' in printed

	# ---- axis-aligned and rotated ellipses, two and three catalogues: against the oracle ------------------------------------
	def tables(specs):
		out = []
		for tab, spec in specs:
			d = dict(name=tab.name, ra=tab.data['RA'].astype(float), dec=tab.data['DEC'].astype(float), area=tab.header['SKYAREA'] * 1.0)
			if spec == ':a:b':
				d['error'] = (tab.data['a'].astype(float), tab.data['b'].astype(float), np.zeros(len(tab)))
			elif spec == ':a:b:phi':
				d['error'] = tuple(O.ellipse_from_cli(tab.data['a'].astype(float), tab.data['b'].astype(float), tab.data['phi'].astype(float)))
			else:
				d['error'] = float(spec) * np.ones(len(tab))
			out.append(d)
		return out

	for out, specs in (('random_asymtest.fits', [(x, ':a:b'), (o, '0.1')]), ('random_elltest.fits', [(x, ':a:b:phi'), (o, '0.1')]),
			('random3_elltest.fits', [(x, ':a:b:phi'), (r, ':a:b:phi'), (o, '0.1')])):
		argv = ['--radius=10.0'] + [w for tab, spec in specs for w in ('randomcat%s.fits' % {'CHANDRA': 'X', 'XMM': 'R', 'OPT': 'O'}[tab.name], spec)]
		t, printed = job(argv + ['--out=' + out, '--min-prob=0.01'])
		names = [tab.name for tab, _ in specs]
		assert 'Separation_%s_%s_ra' % (names[1], names[0]) in t.columns
		assert ('dist_bayesfactor_corrected' in t.columns) == (len(names) >= 3)
		assert ((t.data['p_any'] >= 0) & (t.data['p_any'] <= 1)).all() and not (t.data['p_i'] < 0.01).any()
		want = O.nway_match(tables(specs), 10.0, 1.0, min_prob=0.01, unrelated_mode='cli', cli_compat=True)
		assert len(t) == len(want[names[0]]) > 100
		for nm, (tab, _) in zip(names, specs):
			assert np.array_equal(t.data[nm + '_ID'], np.where(want[nm] >= 0, tab.data['ID'][np.maximum(want[nm], 0)], -99)), nm
		for mine, theirs in (('p_any', 'prob_has_match'), ('p_i', 'prob_this_match'), ('dist_bayesfactor', 'dist_bayesfactor_uncorrected'), ('match_flag', 'match_flag')) + \
				((('dist_bayesfactor_corrected', 'dist_bayesfactor'),) if len(names) >= 3 else ()):
			assert np.array_equal(t.data[mine], want[theirs].astype(t.data[mine].dtype), equal_nan=True), (out, mine)
