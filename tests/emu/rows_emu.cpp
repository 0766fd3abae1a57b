// A whole two-catalogue match on the host, from the device's own per-element source: primary side and grid (emu_k0.h:
// prim_register, build_grid, ...), the secondary stream of k_pairs (cell look-up, packed and fp32 pre-tests, exact
// separation in the reference's arithmetic: k1_* / sep_arcsec_ref), and the rows of k_rows2<FUSE, SHARE> (rows2_write,
// the one-exponential group normalisation with its 32 lane sums, group_p_any, shared_post).  What is NOT the device's
// source here is only the orchestration a kernel adds around these functions (queues, slots, the in-group sort -- done
// with std::sort).  tests/test_rows_emulation_cpu.py compares the resulting tables with the oracle.
#define NWB_HOST_EMU 1
#include "emu_k0.h"
#include "../../nway_b200/csrc/nwb_rows.cuh"

#include <algorithm>
#include <cstring>

using namespace nwb;

extern "C" {

// returns the number of rows (or -(rows needed) - 1 if max_rows is too small).  Columns as in nwayb200.h, 8 bytes each.
long long nwb_emu_match2(int np, const double *pra, const double *pdec, const double *perr,
	int ns, const double *sra, const double *sdec, const double *serr, double radius_arcsec,
	const double *norm /* [3] */, double log10e, const double *prior /* [2] */, const double *log10prior /* [2] */, double ratio_secondary,
	int nmag, const double *const *mag /* [nmag] columns of catalogue 1 */, const int *nbins, const double *const *edges,
	const double *const *weight, const double *const *biasval, double *const *bias_out /* [nmag] */,
	long long max_rows, long long *idx0, long long *idx1, double *sep, double *sepmax, long long *ncat, double *lbf_u, double *lbf,
	double *dist_post, double *p_single, long long *flag, double *p_any, double *p_i)
{
	emu::K0 K;
	emu::build_k0(K, np, pra, pdec, radius_arcsec, 0);
	if (!K.ok) return -1;
	const Grid &G = K.G;
	// primary records as k_prim_prep writes them
	std::vector<PrimRec> prec(np);
	for (int i = 0; i < np; i++) {
		double sl, cl;
		sincos_ref(deg2rad_ref(pdec[i]), &sl, &cl);
		prec[i].lon = deg2rad_ref(pra[i]); prec[i].slat = sl; prec[i].clat = cl; prec[i].ij = 0;
	}
	// k_pairs: every secondary against the entries of its cell; survivors of the pre-test get the exact separation
	std::vector<std::vector<Slot16>> match(np);
	const double nbands_d = (double) G.nbands;
	for (int s = 0; s < ns; s++) {
		const double r = sra[s], d = sdec[s];
		const double t = k1_band_coord(G, d);
		if (!(t >= 0.0 && t < nbands_d)) continue;
		const double x = k1_ra_coord(G, r);
		if (!(G.full_circle || x <= G.ra_span)) continue;
		const int b = __double2int_rd(t);
		const BandRec B = load_band(G, b);
		const float kx = G.kx[b];
		int ic;
		const double xcells = k1_ra_cell(B, x, ic);
		const CellRec &cr = K.cells[B.base + ic];
		const int ecnt = (int) (unsigned) cr.q[0], estart = (int) (cr.q[0] >> 32);
		const float xr = (float) (xcells - (double) ic), yr = (float) (t - (double) b);
		for (int e = 0; e < ecnt; e++) {
			int p;
			bool pass;
			if (e < 3) {
				p = (int) (cr.q[1 + e] >> 32);
				pass = k1_pretest_packed(G, xr, yr, kx, (unsigned) cr.q[1 + e]);
			} else {
				const Entry &en = K.entries[estart + e];
				p = en.p;
				pass = k1_pretest(G, (float) k1_ra_coord(G, r), (float) (d - G.dec_lo), en.x, en.y, en.clat);
			}
			if (!pass) continue;
			// k1_flush
			double slat2, clat2;
			sincos_ref(deg2rad_ref(d), &slat2, &clat2);
			const double lon2 = deg2rad_ref(r);
			const double sp = sep_arcsec_ref(prec[p].lon, prec[p].slat, prec[p].clat, lon2, slat2, clat2);
			if (sp < radius_arcsec) {
				Slot16 m;
				m.s = s; m.pad = 0; m.sep = sp;
				match[p].push_back(m);
			}
		}
	}
	long long R = 0;
	for (int p = 0; p < np; p++) R += (long long) match[p].size() + 1;
	if (R > max_rows) return -R - 1;
	// k_rows2<FUSE = true, SHARE = true>
	ConstTables T;
	memset(&T, 0, sizeof(T));
	for (int k = 0; k < 3; k++) T.norm[k] = norm[k];
	T.log10e = log10e;
	T.prior[0] = prior[0]; T.prior[1] = prior[1];
	T.log10prior[0] = log10prior[0]; T.log10prior[1] = log10prior[1];
	RowParams RP;
	memset(&RP, 0, sizeof(RP));
	for (int j = 0; j < nmag; j++) {   // magnitude priors of the secondary catalogue (nwb_set_maghist)
		MagTable &MT = T.mag[j];
		MT.cat = 1; MT.nbins = nbins[j]; MT.mag = mag[j];
		for (int k = 0; k <= nbins[j]; k++) MT.edges[k] = edges[j][k];
		for (int k = 0; k < nbins[j]; k++) { MT.weight[k] = weight[j][k]; MT.bias[k] = biasval[j][k]; }
	}
	RP.ncat = 2; RP.nmag = nmag; RP.np = np; RP.first = 0;
	RP.radius = radius_arcsec; RP.ratio_secondary = ratio_secondary;
	RP.err[0] = perr; RP.err[1] = serr;
	RP.n[0] = np; RP.n[1] = ns;
	RP.ra[0] = pra; RP.ra[1] = sra; RP.dec[0] = pdec; RP.dec[1] = sdec;
	RP.T = &T;
	RP.C.idx[0] = idx0; RP.C.idx[1] = idx1; RP.C.sep[0] = sep; RP.C.sepmax = sepmax; RP.C.ncat = ncat;
	RP.C.lbf_u = lbf_u; RP.C.lbf = lbf; RP.C.dist_post = dist_post; RP.C.p_single = p_single; RP.C.flag = flag;
	RP.C.p_any = p_any; RP.C.p_i = p_i;
	for (int j = 0; j < nmag; j++) RP.C.bias[j] = bias_out[j];
	const bool share = nmag == 0;   // nwb_api.cu: k_rows2<true, true> without magnitude priors, <true, false> with
	R2Memo memo;
	double m_sig0 = -1.0, w0 = 0.0, lw0 = 0.0;
	std::vector<double> v, tt;
	long long rbase = 0;
	for (int p = 0; p < np; p++) {
		std::vector<Slot16> &M = match[p];
		std::sort(M.begin(), M.end(), [](const Slot16 &a, const Slot16 &b) { return a.s < b.s; });
		const int rows = (int) M.size() + 1;
		const double sig0 = perr[p];
		if (sig0 != m_sig0) { m_sig0 = sig0; w0 = 1.0 / (sig0 * sig0); lw0 = log(w0); }
		v.assign(rows, 0.0);
		tt.assign(rows, 0.0);
		double m_rest = -INFINITY;
		for (int k = 0; k < rows; k++) {
			double vk = 0.0;
			if (share) rows2_write<true, true>(RP, &T, rbase + k, (long long) p, k == 0 ? -1 : (long long) M[k - 1].s, k == 0 ? 0.0 : M[k - 1].sep, w0, lw0, memo, vk);
			else rows2_write<true, false>(RP, &T, rbase + k, (long long) p, k == 0 ? -1 : (long long) M[k - 1].s, k == 0 ? 0.0 : M[k - 1].sep, w0, lw0, memo, vk);
			v[k] = vk;
			if (k > 0) m_rest = fmax(m_rest, vk);
		}
		const double v0 = v[0];
		double lane_sum[32];
		for (int lane = 0; lane < 32; lane++) {
			double sacc = 0.0;
			for (int k = lane + (lane == 0 ? 32 : 0); k < rows; k += 32) {
				tt[k] = nwb_exp10(v[k] - m_rest);
				sacc += tt[k];
			}
			lane_sum[lane] = sacc;
		}
		for (int o = 16; o > 0; o >>= 1) {   // warp_sum
			double nxt[32];
			for (int i = 0; i < 32; i++) nxt[i] = lane_sum[i] + lane_sum[i ^ o];
			memcpy(lane_sum, nxt, sizeof(nxt));
		}
		double pa, rinv;
		group_p_any(rows, v0, m_rest, lane_sum[0], pa, rinv);
		const double best = rinv;
		const bool direct = !(fabs(m_rest) <= 250.0);
		const double oscale = (share && rows > 1 && !direct) ? (1 - T.prior[1]) * nwb_exp10(-m_rest) : 0.0;
		const double omp = 1 - T.prior[1], l10p1 = T.log10prior[1];
		for (int k = 0; k < rows; k++) {
			const long long row = rbase + k;
			const double tk = k == 0 ? v0 : tt[k];
			const double pi = k == 0 ? 0.0 : tk * rinv;
			p_i[row] = pi;
			p_any[row] = pa;
			flag[row] = (pi == best) ? 1 : (pi > ratio_secondary * best ? 2 : 0);
			if (share) {
				const double post = k == 0 ? 1.0 : shared_post(direct, tk, oscale, omp, lbf[row], l10p1);
				dist_post[row] = post;
				p_single[row] = post;
			}
		}
		rbase += rows;
	}
	return R;
}

}  // extern "C"
