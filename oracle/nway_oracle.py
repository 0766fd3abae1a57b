"""CPU oracle: a numpy restatement of the nway match-probability hot path.

TEST INFRASTRUCTURE ONLY -- the checker, never the product.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (nway_b200/) must never route through it.

Parity status: PINNED.  oracle/make_golden.py runs the unmodified reference
(/root/reference, nway 4.7.1 @ a37be1f) through oracle/refrun.py and this oracle on
the same inputs; tests/test_oracle_golden.py re-checks the oracle against the
committed outputs of the real reference (tests/golden/*.npz), against the
reference's own golden row counts (nway-apitest.py:66,109 -> 37836 / 387601) and
against the known-answer values of SURVEY.md Appendix C.

Every function cites the reference lines it restates (paths relative to
/root/reference).  Floating-point operation ORDER follows the reference so the
oracle agrees with it to the last few ulp; the structure of the code does not.
"""
import itertools
import math

import numpy as np

LOG_ARCSEC2RAD = math.log(3600 * 180 / math.pi)      # nwaylib/bayesdistance.py:15
LOG10_E = math.log10(math.e)
FULL_SKY_DEG2 = 4 * math.pi * (180 / math.pi) ** 2   # nwaylib/__init__.py:205


# --------------------------------------------------------------------------------------
# separations
# --------------------------------------------------------------------------------------

def dist(apos, bpos):
	"""Great-circle separation in degrees (Vincenty form).  nwaylib/fastskymatch.py:26-47."""
	ra1, dec1 = apos
	ra2, dec2 = bpos
	lam1 = ra1 / 180 * np.pi
	phi1 = dec1 / 180 * np.pi
	lam2 = ra2 / 180 * np.pi
	phi2 = dec2 / 180 * np.pi
	dl = lam2 - lam1
	sdl, cdl = np.sin(dl), np.cos(dl)
	s1, s2 = np.sin(phi1), np.sin(phi2)
	c1, c2 = np.cos(phi1), np.cos(phi2)
	y1 = c2 * sdl
	y2 = c1 * s2 - s1 * c2 * cdl
	x = s1 * s2 + c1 * c2 * cdl
	return np.arctan2(np.hypot(y1, y2), x) * 180 / np.pi


def offsets_skyoffsetframe(apos, bpos):
	"""(dra, ddec) in degrees as fastskymatch.dist3d forms them (fastskymatch.py:50-74), following astropy's OWN algorithm
	step by step -- astropy (un-vendored, unpinned: pyproject.toml) is absent here, so this restates its published code
	path: SkyOffsetFrame(origin=a) turns ICRS into the offset frame with the matrix
	    R_x(-rotation = 0) @ R_y(-lat_a) @ R_z(lon_a)
	(astropy/coordinates/builtin_frames/skyoffset.py, reference_to_skyoffset; rotation_matrix(angle, axis) of
	astropy/coordinates/matrix_utilities.py: R_z = [[c, s, 0], [-s, c, 0], [0, 0, 1]], R_y = [[c, 0, -s], [0, 1, 0],
	[s, 0, c]]) applied to the unit vector (cos d cos a, cos d sin a, sin d); back to angles with
	UnitSphericalRepresentation.from_cartesian: lon = atan2(y, x), lat = atan2(z, hypot(x, y)); the frame's longitude wraps
	at 180 degrees.  dist3d then takes dra = na.lon - nb.lon, ddec = na.lat - nb.lat with na = the origin in its own frame
	(0 up to rounding -- kept, as the reference keeps it).  tests/test_oracle_golden.py holds offsets() -- the closed
	form the device evaluates -- against this to 1e-10 arcsec everywhere on the sphere."""
	def to_frame(lon0, lat0, lon, lat):
		c0, s0 = np.cos(lon0), np.sin(lon0)
		cl, sl = np.cos(-lat0), np.sin(-lat0)
		rz = np.array([[c0, s0, 0 * c0], [-s0, c0, 0 * c0], [0 * c0, 0 * c0, 1 + 0 * c0]])
		ry = np.array([[cl, 0 * cl, -sl], [0 * cl, 1 + 0 * cl, 0 * cl], [sl, 0 * cl, cl]])
		m = np.einsum('ij...,jk...->ik...', ry, rz)
		v = np.array([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])
		x, y, z = np.einsum('ij...,j...->i...', m, v)
		lon_f = np.degrees(np.arctan2(y, x))
		lon_f = np.where(lon_f >= 180, lon_f - 360, lon_f)
		return lon_f, np.degrees(np.arctan2(z, np.hypot(x, y)))
	ra1, dec1 = np.radians(apos[0]), np.radians(apos[1])
	ra2, dec2 = np.radians(bpos[0]), np.radians(bpos[1])
	ra1, dec1, ra2, dec2 = np.broadcast_arrays(ra1, dec1, ra2, dec2)
	na_lon, na_lat = to_frame(ra1, dec1, ra1, dec1)
	nb_lon, nb_lat = to_frame(ra1, dec1, ra2, dec2)
	return na_lon - nb_lon, na_lat - nb_lat


def offsets(apos, bpos):
	"""(separation, dra, ddec) in degrees: what astropy's SkyOffsetFrame(origin=a) gives
	for b (fastskymatch.py:50-74 -> un-vendored astropy; closed form of SURVEY Appendix A.6: the product of the two
	rotations of offsets_skyoffsetframe() written out).  Pinned to that restatement of astropy's algorithm, not to
	astropy's output: astropy is absent here and no reference test pins dist3d (the command-line program rounds these
	offsets to float32 before use, 6e-8 relative; the two forms differ by ~1e-11 arcsec)."""
	ra1, dec1 = np.radians(apos[0]), np.radians(apos[1])
	ra2, dec2 = np.radians(bpos[0]), np.radians(bpos[1])
	dl = ra2 - ra1
	lon = np.arctan2(np.cos(dec2) * np.sin(dl), np.cos(dec1) * np.cos(dec2) * np.cos(dl) + np.sin(dec1) * np.sin(dec2))
	lat = np.arcsin(np.clip(np.cos(dec1) * np.sin(dec2) - np.sin(dec1) * np.cos(dec2) * np.cos(dl), -1, 1))
	return dist(apos, bpos), -np.degrees(lon), -np.degrees(lat)


# --------------------------------------------------------------------------------------
# candidate enumeration
# --------------------------------------------------------------------------------------

def flat_sky_applicable(radectables, err):
	"""fastskymatch.py:94-98."""
	for ra, dec in radectables:
		if not (err < 1 and (ra > 10 * err).all() and (ra < 360 - 10 * err).all() and (np.abs(dec) < 45).all()):
			return False
	return True


def crossproduct_refhash(radectables, err):
	"""The reference's own enumeration algorithm, flat-sky branch: square cells of `err` degrees,
	every source dropped into its cell and the three cells towards +ra/+dec, only the primary
	may open a bucket; per bucket the Cartesian product with -1 = "absent"; union; sort.
	fastskymatch.py:119-133 (hash) and :164-218 (product).  Interpreted loops on purpose: this is
	what the reference executes, and it is the CPU baseline bench.py times."""
	assert flat_sky_applicable(radectables, err), 'only the flat-sky branch is restated (healpy is un-vendored)'
	ncat = len(radectables)
	buckets = {}
	for ti, (ras, decs) in enumerate(radectables):
		ci = (ras / err).astype(np.int64)      # int() truncates toward zero, like astype on finite values
		cj = (decs / err).astype(np.int64)
		# astype(int64) of a negative float truncates toward zero as python int() does
		for ei, (i, j) in enumerate(zip(ci.tolist(), cj.tolist())):
			for key in ((i, j), (i + 1, j), (i, j + 1), (i + 1, j + 1)):
				slot = buckets.get(key)
				if slot is None:
					if ti != 0:
						continue
					slot = buckets[key] = [[] for _ in range(ncat)]
				slot[ti].append(ei)
	tuples = set()
	for lists in buckets.values():
		options = [sorted(lists[0])] + [[-1] + sorted(li) for li in lists[1:]]
		tuples.update(itertools.product(*options))
	if not tuples:
		return np.zeros((0, ncat), dtype=np.int64)
	return np.array(sorted(tuples), dtype=np.int64)


def flat_hash_keeps(radectables, err, idx):
	"""Which index tuples does the reference's flat-sky hash form at all?  Those whose present members were all put
	into one bucket: a source goes to the cells (i, j) .. (i+1, j+1) with i = int(ra / err), j = int(dec / err)
	(fastskymatch.py:125-132), so the members share a bucket iff their i span at most one step and their j likewise.
	Vectorised form of crossproduct_refhash (which restates the loops literally); equal to it on every row that
	survives the radius filter -- tests/test_oracle_golden.py -- and pinned to the real reference by the off-equator
	goldens (tests/golden/ref_offeq*.npz)."""
	big = np.iinfo(np.int64).max
	imin = np.full(len(idx), big); imax = np.full(len(idx), -big)
	jmin = np.full(len(idx), big); jmax = np.full(len(idx), -big)
	for c, (ras, decs) in enumerate(radectables):
		present = idx[:, c] >= 0
		ci = (ras / err).astype(np.int64)[idx[:, c]]
		cj = (decs / err).astype(np.int64)[idx[:, c]]
		imin = np.where(present, np.minimum(imin, ci), imin); imax = np.where(present, np.maximum(imax, ci), imax)
		jmin = np.where(present, np.minimum(jmin, cj), jmin); jmax = np.where(present, np.maximum(jmax, cj), jmax)
	return (imax - imin <= 1) & (jmax - jmin <= 1)


def _unitvec(ra, dec):
	lam, phi = np.radians(ra), np.radians(dec)
	return np.stack([np.cos(phi) * np.cos(lam), np.cos(phi) * np.sin(lam), np.sin(phi)], axis=1)


def neighbour_lists(radectables, radius_deg):
	"""For every primary source i and secondary catalogue c: the ascending list of j with
	dist(i, j)*3600 < radius (strict, the predicate of nwaylib/__init__.py:163,180).
	Complete on the whole sphere (KD-tree on unit vectors with a chord margin, then the exact
	reference formula) -- the stand-in for the HEALPix branch fastskymatch.py:134-160."""
	from scipy.spatial import cKDTree
	radius_arcsec = radius_deg * 60 * 60
	ra0, dec0 = radectables[0]
	u0 = _unitvec(ra0, dec0)
	chord = 2 * math.sin(math.radians(radius_deg) / 2) * (1 + 1e-6) + 1e-12
	out = []
	for ra, dec in radectables[1:]:
		tree = cKDTree(_unitvec(ra, dec))
		cand = tree.query_ball_point(u0, chord)
		lists = []
		for i, js in enumerate(cand):
			js = np.sort(np.asarray(js, dtype=np.int64))
			if len(js):
				sep = dist((ra0[i], dec0[i]), (ra[js], dec[js])) * 60 * 60
				js = js[sep < radius_arcsec]
			lists.append(js)
		out.append(lists)
	return out


def crossproduct_complete(radectables, radius_deg):
	"""All index tuples (i0, i1|-1, ...) whose present members all lie within the radius OF THE
	PRIMARY, in the reference's output order (lexicographic, -1 first: fastskymatch.py:178-181,217).
	A superset-free replacement for crossproduct(): the secondary-secondary filter is applied by
	create_match_table exactly as the reference does (__init__.py:180)."""
	ncat = len(radectables)
	lists = neighbour_lists(radectables, radius_deg)
	blocks = []
	for i in range(len(radectables[0][0])):
		opts = [np.concatenate(([-1], lists[c][i])) for c in range(ncat - 1)]
		sizes = [len(o) for o in opts]
		total = int(np.prod(sizes)) if sizes else 1
		block = np.empty((total, ncat), dtype=np.int64)
		block[:, 0] = i
		rep = total
		for c, o in enumerate(opts):
			rep //= sizes[c]
			block[:, c + 1] = np.tile(np.repeat(o, rep), total // (rep * sizes[c]))
		blocks.append(block)
	if not blocks:
		return np.zeros((0, ncat), dtype=np.int64)
	return np.concatenate(blocks, axis=0)


def is_elliptical(tables):
	"""a catalogue's 'error' may be a triple (sigma_ra, sigma_dec, rho) instead of one sigma column: the CLI's
	elliptical mode (nway.py:25-98, 346-354).  One elliptical catalogue switches the whole match to that mode,
	circular catalogues becoming (sigma, sigma, 0) (nway.py:79-88)."""
	return any(isinstance(t['error'], (tuple, list)) and len(t['error']) == 3 for t in tables)


def error_triples(tables):
	out = []
	for t in tables:
		e = t['error']
		if isinstance(e, (tuple, list)) and len(e) == 3:
			out.append(tuple(np.asarray(x, dtype=float) for x in e))
		else:
			e = np.asarray(e, dtype=float)
			out.append((e, e, np.zeros_like(e)))
	return out


def create_match_table(tables, match_radius, enumerator='reference', sep_f32=False, pairwise_errs=()):
	"""nwaylib/__init__.py:123-196.  Returns dict(idx (R,N) int64, sep {(a,b): (R,)}, sepmax, ncat,
	errors [N x (R,)]); in elliptical mode also off {(a,b): (dra, ddec)} in arcsec, measured like the CLI does
	(fastskymatch.py:299-331: offset frame centred on the source of the LATER catalogue b, the earlier one is the
	target) and errors as triples.  sep_f32: the command-line program keeps separations and offsets in float32
	FITS columns ('E', fastskymatch.py:328-333) and scores what it reads back (nway.py:269,302-305); the radius
	filter runs before that, on the fp64 values (fastskymatch.py:335)."""
	radec = [(np.asarray(t['ra'], dtype=float), np.asarray(t['dec'], dtype=float)) for t in tables]
	radius_deg = match_radius / 60. / 60
	if enumerator == 'refhash':
		idx = crossproduct_refhash(radec, radius_deg)
	else:
		idx = crossproduct_complete(radec, radius_deg)
		if enumerator == 'reference' and flat_sky_applicable(radec, radius_deg):
			idx = idx[flat_hash_keeps(radec, radius_deg, idx)]
	n = len(tables)
	nrows = len(idx)
	sep = {}
	sepmax = np.zeros(nrows)
	for a in range(n):
		for b in range(a + 1, n):
			ia, ib = idx[:, a], idx[:, b]
			col = dist((radec[a][0][ia], radec[a][1][ia]), (radec[b][0][ib], radec[b][1][ib]))
			col[(ia == -1) | (ib == -1)] = np.nan
			col = col * 60 * 60
			sep[(a, b)] = col
			with np.errstate(invalid='ignore'):
				sepmax = np.where(np.isnan(col), sepmax, np.maximum(col, sepmax))
	keep = sepmax < match_radius
	# --prefilter-pair as intended (fastskymatch.py:184-208): with both members present the pair must be closer than
	# its own radius.  The reference as written drops every tuple holding both (the mask assignment at :203 writes to
	# a temporary): that is radius 0 here -- pinned against the real nway.py in tests/golden/ref_cli_*.
	for a, b, rad in pairwise_errs:
		col = sep[(min(a, b), max(a, b))]
		with np.errstate(invalid='ignore'):
			keep &= np.isnan(col) | (col < rad)
	idx = idx[keep]
	# sep_f32: the separation arrays STAY float32, so that log_bf squares them in float32 exactly as numpy does for
	# the reference (`p[i][j]**2` on an 'E' column, bayesdistance.py:83); offsets are rounded and widened again
	# (the elliptical branch cannot run here, its float32 arithmetic is not pinned: DESIGN.md)
	f32 = (lambda x: x.astype(np.float32)) if sep_f32 else (lambda x: x)
	out = dict(idx=idx, sep={k: f32(v[keep]) for k, v in sep.items()}, sepmax=f32(sepmax[keep]).astype(float), ncat=(idx > -1).sum(axis=1))
	if is_elliptical(tables):
		trip = error_triples(tables)
		out['errors'] = [tuple(x[idx[:, c]] for x in trip[c]) for c in range(n)]
		out['off'] = {}
		for a in range(n):
			for b in range(a + 1, n):
				ia, ib = idx[:, a], idx[:, b]
				_, dra, ddec = offsets((radec[b][0][ib], radec[b][1][ib]), (radec[a][0][ia], radec[a][1][ia]))
				out['off'][(a, b)] = (f32(dra * 60 * 60).astype(float), f32(ddec * 60 * 60).astype(float))
	else:
		out['errors'] = [np.asarray(t['error'], dtype=float)[idx[:, c]] for c, t in enumerate(tables)]
	return out


# --------------------------------------------------------------------------------------
# Bayes factors, priors, posteriors
# --------------------------------------------------------------------------------------

def log_bf(p, s):
	"""log10 positional Bayes factor for n catalogues.  p[i][j] (i<j) separations in arcsec,
	s[i] 1-sigma errors in arcsec.  nwaylib/bayesdistance.py:64-86."""
	n = len(s)
	w = [np.asarray(si, dtype=float) ** -2. for si in s]
	norm = (n - 1) * math.log(2) + 2 * (n - 1) * LOG_ARCSEC2RAD
	wsum = w[0]
	for wi in w[1:]:
		wsum = wsum + wi
	slog = np.log(w[0])
	for wi in w[1:]:
		slog = slog + np.log(wi)
	slog = slog - np.log(wsum)
	q = 0
	for i in range(n):
		for j in range(i + 1, n):
			q = q + w[i] * w[j] * np.asarray(p[i][j]) ** 2
	return (norm + slog + (-q / 2 / wsum)) * LOG10_E


def posterior(prior, log_bf_):
	"""bayesdistance.py:26-32."""
	with np.errstate(over='ignore'):
		return 1. / (1 + (1 - prior) * 10 ** (-log_bf_ - np.log10(prior)))


def source_densities(tables):
	"""nwaylib/__init__.py:199-217."""
	nu = np.array([len(t['ra']) / (t['area'] * 1.0) * FULL_SKY_DEG2 for t in tables])
	nu_plus = np.array([(len(t['ra']) + 1) / (t['area'] * 1.0) * FULL_SKY_DEG2 for t in tables])
	nu_plus[0] = nu[0]
	return nu, nu_plus


def completeness_vector(prior_completeness, ncats):
	"""nwaylib/__init__.py:224-229."""
	if np.shape(prior_completeness) == ():
		return np.array([1.0] + [float(prior_completeness) ** (1. / (ncats - 1)) for _ in range(1, ncats)])
	pc = np.asarray(prior_completeness, dtype=float)
	if len(pc) != ncats:
		raise Exception('Prior completeness needs one value per catalog.')
	assert pc[0] == 1.0
	return pc


def presence_patterns(idx):
	"""group rows by which secondaries are present; yields (present tuple incl. 0, row mask).
	Order of cases as nwaylib/__init__.py:234-242."""
	n = idx.shape[1]
	for case in range(2 ** (n - 1)):
		present = [True] + [(case // 2 ** ti) % 2 == 0 for ti in range(n - 1)]
		mask = np.ones(len(idx), dtype=bool)
		for c in range(1, n):
			mask &= (idx[:, c] > -1) == present[c]
		yield [c for c in range(n) if present[c]], mask


def single_log_bf(mt, nu, nu_plus, pc):
	"""nwaylib/__init__.py:220-259: per-row log10 BF and prior."""
	idx = mt['idx']
	lbf = np.full(len(idx), np.nan)
	prior = np.full(len(idx), np.nan)
	for cats, mask in presence_patterns(idx):
		if not mask.any():
			continue
		if 'off' in mt:   # nway.py:346-354
			sra = [[mt['off'][(a, b)][0][mask] if a < b else None for b in cats] for a in cats]
			sde = [[mt['off'][(a, b)][1][mask] if a < b else None for b in cats] for a in cats]
			errs = [tuple(x[mask] for x in mt['errors'][c]) for c in cats]
			lbf[mask] = log_bf_elliptical(sra, sde, errs) if len(cats) > 1 else 0.0
		else:
			p = [[mt['sep'][(a, b)][mask] if a < b else None for b in cats] for a in cats]
			s = [mt['errors'][c][mask] for c in cats]
			lbf[mask] = log_bf(p, s)
		sel = np.zeros(idx.shape[1], dtype=bool)
		sel[cats] = True
		prior[mask] = nu[0] * np.prod(pc[sel]) / np.prod(nu_plus[sel])
	return prior, lbf


# --------------------------------------------------------------------------------------
# elliptical errors (CLI only in the reference)
# --------------------------------------------------------------------------------------

def convert_from_ellipse(a, b, phi):
	"""bayesdistance.py:190-204."""
	a2, b2 = a ** 2, b ** 2
	s, c = np.sin(phi), np.cos(phi)
	s2, c2 = s ** 2, c ** 2
	sx = (a2 * s2 + b2 * c2) ** 0.5
	sy = (a2 * c2 + b2 * s2) ** 0.5
	return sx, sy, c * s * (a2 - b2) / (sx * sy)


def ellipse_from_cli(major, minor, angle_deg):
	"""nway.py:56-65: the CLI feeds (angle-90)/180*pi."""
	return convert_from_ellipse(major, minor, (angle_deg - 90) / 180 * np.pi)


def _directional_precision(vx, vy, sx, sy, rho):
	"""v^T Sigma^-1 v for the unit vector v.  bayesdistance.py:150-161,183-187."""
	f = 1.0 / (sx ** 2 * sy ** 2 * (1 - rho ** 2))
	m11, m12, m22 = f * sy ** 2, f * -rho * sx * sy, f * sx ** 2
	l1 = vx * m11 + vy * m12
	l2 = vx * m12 + vy * m22
	return l1 * vx + l2 * vy


def log_bf_elliptical(sep_ra, sep_dec, pos_errors):
	"""bayesdistance.py:207-240."""
	n = len(pos_errors)
	circ = [((sx ** 2 + sy ** 2) / 2) ** 0.5 for sx, sy, rho in pos_errors]
	newsep = [[None] * n for _ in range(n)]
	for i in range(n):
		for j in range(i + 1, n):
			vx, vy = np.asarray(sep_ra[i][j], dtype=float), np.asarray(sep_dec[i][j], dtype=float)
			d = (vx * vx + vy * vy) ** 0.5
			ux = np.where(d == 0, 2 ** -0.5, vx / (d + 1e-300))
			uy = np.where(d == 0, 2 ** -0.5, vy / (d + 1e-300))
			wi = _directional_precision(ux, uy, *pos_errors[i])
			wj = _directional_precision(ux, uy, *pos_errors[j])
			ratio = (circ[i] ** 2 + circ[j] ** 2) / (1 / wi + 1 / wj)
			newsep[i][j] = d * ratio ** -0.5
	return log_bf(newsep, circ)


# --------------------------------------------------------------------------------------
# unrelated-association correction: the LIVE algorithm of the CLI
# --------------------------------------------------------------------------------------

def correct_unrelated_cli(mt, lbf, nu, nu_plus, group_start):
	"""nway.py:366-421.  For each row missing >= 2 catalogues, add the best positive log-posterior of a
	sub-association, made only of catalogues the row lacks, found in any row of the same primary with ncat > 2.
	The API's version (__init__.py:262-301) is inert and therefore not restated: API mode == no correction.

	The script walks rows i and j in two nested Python loops and scores one sub-association at a time; the same numbers
	are obtained here array-wise: the sub-association of row j depends only on (catalogues row i lacks) x (catalogues row
	j holds), i.e. on a pair of presence patterns, so for every such pair the log-posteriors of all rows j are one
	vectorised log_bf call, and "best of the group" is a segmented maximum."""
	idx = mt['idx']
	n = idx.shape[1]
	out = lbf.copy()
	nrows = len(idx)
	if nrows == 0 or n < 3:
		return out
	ncat = mt['ncat']
	present = idx[:, 1:] != -1
	pattern = (present * (1 << np.arange(n - 1))).sum(axis=1)   # bit k-1 set: catalogue k present
	starts = np.asarray(group_start, dtype=np.int64)
	group = np.repeat(np.arange(len(starts)), np.diff(np.concatenate((starts, [nrows]))))
	rich = ncat > 2
	full = (1 << (n - 1)) - 1
	score = {}   # (aug tuple) -> log-posterior of that sub-association for every row holding all of aug (NaN elsewhere)

	def sub_posterior(aug):
		if aug not in score:
			rows = np.flatnonzero(rich & present[:, [k - 1 for k in aug]].all(axis=1))
			val = np.full(nrows, np.nan)
			if len(rows):
				pr = nu[aug[0]] / np.prod(nu_plus[list(aug)])
				if 'off' in mt:   # nway.py:404-411
					sra = [[mt['off'][(a, b)][0][rows] if a < b else None for b in aug] for a in aug]
					sde = [[mt['off'][(a, b)][1][rows] if a < b else None for b in aug] for a in aug]
					errs = [tuple(x[rows] for x in mt['errors'][k]) for k in aug]
					v = log_bf_elliptical(sra, sde, errs)
				else:
					# nway.py:389-392 builds one numpy.array from float32 separations and float64 NaNs: float64
					p = [[mt['sep'][(a, b)][rows].astype(float) if a < b else None for b in aug] for a in aug]
					s = [mt['errors'][k][rows] for k in aug]
					v = log_bf(p, s)
				val[rows] = v + np.log10(pr)
			score[aug] = val
		return score[aug]

	for lacking in range(1, full + 1):   # the catalogues row i lacks, as a bit set
		missing = [k for k in range(1, n) if (lacking >> (k - 1)) & 1]
		if len(missing) < 2:
			continue
		needy = np.flatnonzero((pattern == (full & ~lacking)) & (ncat <= n - 2))
		if len(needy) == 0:
			continue
		best = np.zeros(len(starts))   # per primary: max(0, best sub-association posterior)
		for holds in np.unique(pattern[rich]):
			aug = tuple(k for k in missing if (holds >> (k - 1)) & 1)
			if len(aug) < 2:
				continue
			rows = np.flatnonzero(rich & (pattern == holds))
			np.maximum.at(best, group[rows], sub_posterior(aug)[rows])
		out[needy] += best[group[needy]]
	return out


# --------------------------------------------------------------------------------------
# magnitude priors
# --------------------------------------------------------------------------------------

def hist_ratio(hist_sel, hist_all):
	"""magnitudeweights.py:18-23."""
	with np.errstate(divide='ignore', invalid='ignore'):
		return np.where(hist_all == 0, 100, hist_sel / hist_all)


def bias_lookup(edges, hist_sel, hist_all, mag):
	"""Zero-order hold over the bin edges, last edge inclusive, NaN outside or for NaN input.
	magnitudeweights.py:74-87 (scipy interp1d kind='zero', bounds_error=False)."""
	y = hist_ratio(np.asarray(hist_sel, dtype=float), np.asarray(hist_all, dtype=float))
	edges = np.asarray(edges, dtype=float)
	mag = np.asarray(mag, dtype=float)
	k = np.searchsorted(edges, mag, side='right') - 1
	k = np.where(mag == edges[-1], len(edges) - 2, k)
	inside = (mag >= edges[0]) & (mag <= edges[-1])
	return np.where(inside, y[np.clip(k, 0, len(y) - 1)], np.nan)


def adaptive_histograms(mag_all, mag_sel, weights=None):
	"""magnitudeweights.py:90-118."""
	from scipy.interpolate import interp1d
	if weights is None:
		weights = np.ones(len(mag_sel))
	order = np.argsort(mag_sel)
	xs = mag_sel[order]
	cw = np.cumsum(weights[order]) / np.sum(weights)
	cw[0], cw[-1] = 0, 1
	edges = np.unique(interp1d(cw, xs)(np.linspace(0, 1, 15)))
	lo, hi = np.nanmin(mag_all), np.nanmax(mag_all)
	if edges[-1] < hi:
		edges = np.asarray(list(edges) + [hi + 1])
	if edges[0] > lo:
		edges = np.asarray([lo - 1] + list(edges))
	hs, edges = np.histogram(mag_sel, bins=edges, density=True, weights=weights)
	ha, edges = np.histogram(mag_all, bins=edges, density=True)
	return edges, hs, ha


def auto_histogram(res, magvals, sepmax, dist_post, mag_include_radius, mag_exclude_radius, minprob, cli=False):
	"""nwaylib/__init__.py:324-366, including quirk Q7 (weights compressed by res_defined but
	indexed by positions inside res[selection]); cli=True: nway.py:455-503 (weights compressed by the selection)."""
	res_defined = res != -1
	mask_all = np.isfinite(magvals)
	if mag_include_radius is not None:
		selection = sepmax < mag_include_radius
		possible = sepmax < mag_exclude_radius
		sw = np.ones(len(selection))
	else:
		selection = dist_post > minprob
		sw = dist_post
		possible = dist_post > 0.01
	selection = selection & res_defined
	sw = sw[selection] if cli else sw[res_defined]
	possible = possible & res_defined
	rows, first = np.unique(res[selection], return_index=True)
	rw = sw[first]
	assert len(rows) > 0
	mag_sel = magvals[rows]
	others = mask_all.copy()
	others[np.unique(res[possible])] = False
	ok = np.isfinite(mag_sel)
	edges, hs, ha = adaptive_histograms(magvals[others], mag_sel[ok], weights=rw[ok])
	return edges, hs, ha, int(ok.sum())


# --------------------------------------------------------------------------------------
# group normalisation
# --------------------------------------------------------------------------------------

def group_starts(primary_col):
	"""first row of every run of equal primary index (nway.py:308-322)."""
	if len(primary_col) == 0:
		return np.zeros(0, dtype=np.int64)
	change = np.flatnonzero(np.diff(primary_col) != 0) + 1
	return np.concatenate(([0], change)).astype(np.int64)


def group_statistics(v, starts, ratio_secondary):
	"""nwaylib/__init__.py:423-457 == nway.py:547-578.  v = log_post_weight per row."""
	nrows = len(v)
	p_any = np.zeros(nrows)
	p_i = np.zeros(nrows)
	flag = np.zeros(nrows, dtype=np.int64)
	bounds = list(starts) + [nrows]
	for g in range(len(bounds) - 1):
		vals = v[bounds[g]:bounds[g + 1]].copy()
		off = vals.max()
		bfsum = np.log10((10 ** (vals - off)).sum()) + off
		if len(vals) > 1:
			off = vals[1:].max()
			bfsum1 = np.log10((10 ** (vals[1:] - off)).sum()) + off
		else:
			bfsum1 = 0
		pa = 1 - 10 ** (vals[0] - bfsum)
		vals[0] = bfsum1
		pi = 10 ** (vals - bfsum1)
		pi[0] = 0
		best = pi.max()
		sl = slice(bounds[g], bounds[g + 1])
		p_any[sl] = pa
		p_i[sl] = pi
		flag[sl] = np.where(best == pi, 1, np.where(pi > ratio_secondary * best, 2, 0))
	return p_any, p_i, flag


# --------------------------------------------------------------------------------------
# the whole path
# --------------------------------------------------------------------------------------

def nway_match(tables, match_radius, prior_completeness, mag_include_radius=None, mag_exclude_radius=None,
		magauto_post_single_minvalue=0.9, prob_ratio_secondary=0.5, min_prob=0.,
		unrelated_mode='api', enumerator='reference', cli_compat=False, pairwise_errs=()):
	"""nwaylib.nway_match (nwaylib/__init__.py:31-120) as a dict of numpy columns.
	unrelated_mode 'api' reproduces the API (inert correction, Q1); 'cli' applies nway.py:366-421.
	cli_compat: float32 separations / offsets (Q2) and the CLI's histogram weights (Q7), as nway.py computes.
	enumerator: 'reference' (default) = the reference's row set: where nwaylib uses its flat-sky hash
	(fastskymatch.py:94-98) the complete search filtered by that hash's bucket predicate (flat_hash_keeps, incomplete
	away from the equator like the original), where it uses HEALPix the complete search; 'refhash' = the
	flat-sky hash restated loop by loop (asserts that it applies); 'complete' = every association within the radius."""
	if mag_exclude_radius is None:
		mag_exclude_radius = mag_include_radius
	n = len(tables)
	names = [t['name'] for t in tables]
	mt = create_match_table(tables, match_radius, enumerator=enumerator, sep_f32=cli_compat, pairwise_errs=pairwise_errs)
	idx = mt['idx']
	if len(idx) == 0:
		raise ValueError('No matches.')
	nu, nu_plus = source_densities(tables)
	pc = completeness_vector(prior_completeness, n)
	prior, lbf = single_log_bf(mt, nu, nu_plus, pc)
	starts = group_starts(idx[:, 0])
	lbf_corr = lbf
	if unrelated_mode == 'cli':
		lbf_corr = correct_unrelated_cli(mt, lbf, nu, nu_plus, starts)
	out = {}
	for c in range(n):
		out[names[c]] = idx[:, c]
	for (a, b), col in mt['sep'].items():
		out['Separation_%s_%s' % (names[a], names[b])] = col.astype(float)
	out['Separation_max'] = mt['sepmax']
	out['ncat'] = mt['ncat']
	out['dist_bayesfactor_uncorrected'] = lbf
	out['dist_bayesfactor'] = lbf_corr
	out['dist_post'] = posterior(prior, lbf_corr)
	wsum = 0   # the reference adds sum(biases.values()) to log_bf: ((0 + w1) + w2) first (__init__.py:394)
	hists = {}
	for c, t in enumerate(tables):
		for magvals, maghist, magname in zip(t.get('mags', []), t.get('maghists', []), t.get('magnames', [])):
			magvals = np.array(magvals)   # keep the caller's dtype: float32 edges differ from float64 ones
			magvals[magvals == -99] = np.nan
			res = idx[:, c]
			if maghist is None:
				edges, hs, ha, nsel = auto_histogram(res, magvals, mt['sepmax'], out['dist_post'],
					mag_include_radius, mag_exclude_radius, magauto_post_single_minvalue, cli=cli_compat)
			else:
				lo, hi, hs, ha = maghist
				edges = np.array(list(lo) + [hi[-1]])
			hists['%s_%s' % (names[c], magname)] = (edges, np.asarray(hs), np.asarray(ha))
			m = magvals[res]
			m[~((res != -1) & np.isfinite(m))] = -99
			with np.errstate(divide='ignore'):
				wgt = np.log10(bias_lookup(edges, hs, ha, m))
			wgt[np.isnan(wgt)] = 0
			out['bias_%s_%s' % (names[c], magname)] = 10 ** wgt
			wsum = wsum + wgt
	total = lbf_corr + wsum
	out['p_single'] = posterior(prior, total)
	v = total + np.log10(prior)
	p_any, p_i, flag = group_statistics(v, starts, prob_ratio_secondary)
	out['match_flag'] = flag
	out['prob_has_match'] = p_any
	out['prob_this_match'] = p_i
	if min_prob > 0:
		keep = ~(p_i < min_prob)
		out = {k: col[keep] for k, col in out.items()}
	out['_hists'] = hists
	return out
