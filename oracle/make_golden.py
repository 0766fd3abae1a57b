"""Generate tests/golden/* from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py                  everything
    python oracle/make_golden.py allsky2 allsky3  only the reference outputs of the named cases

Writes
  tests/golden/cosmos_subset.npz   inputs: the reference's demo catalogues (doc/COSMOS_*.fits) reduced to
                                   the columns the path reads and, for the two big catalogues, to the rows
                                   within 30 arcsec of any XMM source (none of the dropped rows can appear in
                                   a 20-arcsec match, so the reference's golden row counts 37836 / 387601 --
                                   nway-apitest.py:66,109 -- are preserved)
  tests/golden/ref_<case>.npz      outputs of the real nwaylib.nway_match on those inputs / on seeded
                                   synthetic inputs: index-table digest, per-primary p_any, column sums and
                                   a strided sample of full rows
  tests/golden/kat.npz             known-answer vectors from the real fastskymatch.dist and
                                   bayesdistance.{log_bf,log_bf_elliptical,posterior,convert_from_ellipse}
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

from oracle import refrun  # noqa: E402
from oracle import nway_oracle as O  # noqa: E402
from tests import cases  # noqa: E402


def digest(res, names, stride):
	idx = np.stack([res[n].values for n in names], axis=1).astype(np.int64)
	out = dict(nrows=np.int64(len(res)), idx_sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(idx).tobytes()).digest(), dtype=np.uint8))
	starts = O.group_starts(idx[:, 0])
	out['primary'] = idx[starts, 0]
	out['group_size'] = np.diff(np.concatenate((starts, [len(idx)])))
	out['p_any'] = res['prob_has_match'].values[starts]
	sel = np.arange(0, len(res), stride)
	# plus the complete first 8 groups
	head = np.arange(0, starts[min(8, len(starts) - 1)])
	sel = np.unique(np.concatenate((head, sel)))
	out['sample_rows'] = sel
	cols = [c for c in res.columns]
	out['columns'] = np.array(cols)
	for c in cols:
		v = res[c].values
		out['col_' + c] = v[sel]
		if v.dtype.kind == 'f':
			out['sum_' + c] = np.float64(np.nansum(v[np.isfinite(v)]))
		else:
			out['sum_' + c] = np.int64(v.sum())
	return out


def reference_outputs(only=None):
	for name, spec in cases.GOLDEN_CASES.items():
		if only and name not in only:
			continue
		tables = cases.build_case(name)
		names = [t['name'] for t in tables]
		res = refrun.run_reference(tables, spec['radius'], spec['completeness'], **spec.get('kwargs', {}))
		# the oracle must agree with the real reference here and now, to the bit for the row set and
		# to 1e-13 for the floats (it agrees to 0 ulp on this machine, but numpy SIMD paths may differ)
		orc = O.nway_match(cases.build_case(name), spec['radius'], spec['completeness'], enumerator=spec.get('enumerator', 'reference'),
			**spec.get('kwargs', {}))
		for c in res.columns:
			a, b = res[c].values, orc[c]
			assert len(a) == len(b), (name, c)
			if a.dtype.kind in 'iu':
				assert (a == b).all(), (name, c)
			else:
				assert np.allclose(a, b, rtol=1e-13, atol=1e-300, equal_nan=True), (name, c)
		d = digest(res, names, spec.get('stride', 37))
		np.savez_compressed(os.path.join(GOLDEN, 'ref_%s.npz' % name), **d)
		print('%-16s rows %8d  sum p_any %.12f' % (name, len(res), d['p_any'].sum()))


def main():
	os.makedirs(GOLDEN, exist_ok=True)
	refrun.load_reference()
	if len(sys.argv) > 1:
		reference_outputs(only=set(sys.argv[1:]))
		return

	# ---- COSMOS subset fixture ------------------------------------------------------------
	full = refrun.cosmos_tables(3, mags=True)
	from scipy.spatial import cKDTree
	u0 = O._unitvec(full[0]['ra'], full[0]['dec'])
	keep_r = 2 * np.sin(np.radians(30. / 3600) / 2)
	fixture = {}
	for t in full:
		n = len(t['ra'])
		if t['name'] == 'XMM':
			keep = np.ones(n, dtype=bool)
		else:
			d, _ = cKDTree(u0).query(O._unitvec(t['ra'], t['dec']))
			keep = d < keep_r
		fixture[t['name'] + '_ra'] = t['ra'][keep]
		fixture[t['name'] + '_dec'] = t['dec'][keep]
		fixture[t['name'] + '_error'] = np.asarray(t['error'])[keep].astype(np.float32)
		fixture[t['name'] + '_nfull'] = np.int64(n)
		if t['mags']:
			fixture[t['name'] + '_mag'] = t['mags'][0][keep]
		print(t['name'], n, '->', keep.sum())
	np.savez_compressed(os.path.join(GOLDEN, 'cosmos_subset.npz'), **fixture)

	# ---- reference outputs ---------------------------------------------------------------------
	reference_outputs()

	# ---- known-answer vectors -------------------------------------------------------------------
	nw = refrun.load_reference()
	B, M = nw.bayesdistance, nw.fastskymatch
	rng = np.random.default_rng(20260101)
	n = 2000
	kat = {}
	ra1 = rng.uniform(0, 360, n)
	dec1 = np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
	step = 10 ** rng.uniform(-5, -1, n)
	ra2 = ra1 + step * rng.normal(size=n) / np.maximum(np.cos(np.radians(dec1)), 1e-3)
	dec2 = np.clip(dec1 + step * rng.normal(size=n), -90, 90)
	# the literal pairs of tests/fastskymatch_test.py:16-29 first
	lit = np.array([[53.15964508, -27.92927742, 53.15953445, -27.9313736], [150, 2, 150.001, 2.001]])
	ra1[:2], dec1[:2], ra2[:2], dec2[:2] = lit[:, 0], lit[:, 1], lit[:, 2], lit[:, 3]
	kat.update(dist_ra1=ra1, dist_dec1=dec1, dist_ra2=ra2, dist_dec2=dec2, dist_out=M.dist((ra1, dec1), (ra2, dec2)))
	for ncat in (1, 2, 3, 4):
		s = [rng.uniform(0.05, 3, n) for _ in range(ncat)]
		p = [[rng.uniform(0, 20, n) if i < j else None for j in range(ncat)] for i in range(ncat)]
		kat['logbf%d_s' % ncat] = np.array(s)
		kat['logbf%d_p' % ncat] = np.array([[p[i][j] if i < j else np.full(n, np.nan) for j in range(ncat)] for i in range(ncat)])
		kat['logbf%d_out' % ncat] = B.log_bf(p, s) * np.ones(n)
	prior = 10 ** rng.uniform(-12, 0, n)
	lbf = rng.uniform(-400, 30, n)
	kat.update(post_prior=prior, post_logbf=lbf, post_out=B.posterior(prior, lbf))
	ell = [(rng.uniform(0.5, 3, n), rng.uniform(0.2, 0.5, n), rng.uniform(0, 180, n)) for _ in range(3)]
	conv = [B.convert_from_ellipse(a, b, (ang - 90) / 180 * np.pi) for a, b, ang in ell]
	sra = [[rng.normal(0, 2, n) for j in range(3)] for i in range(3)]
	sde = [[rng.normal(0, 2, n) for j in range(3)] for i in range(3)]
	sra[0][1][:4] = 0
	sde[0][1][:4] = 0
	kat.update(ell_in=np.array(ell), ell_conv=np.array(conv), ell_sra=np.array(sra), ell_sdec=np.array(sde),
		ell_out=B.log_bf_elliptical(sra, sde, conv))
	np.savez_compressed(os.path.join(GOLDEN, 'kat.npz'), **kat)
	print('kat written')


if __name__ == '__main__':
	main()
