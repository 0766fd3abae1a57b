// An N-catalogue match (N = 2 .. 4, circular errors, magnitude priors, the API's semantics) on the host, with the
// device's own functions wherever one exists: grid / registration / pre-tests / exact separation as in rows_emu.cpp
// for every secondary catalogue, the secondary-secondary separations (sep_arcsec_ref on the records k_sort_lists keeps),
// log_bf_ref<NC>, posterior_ref, row_bias.  The enumeration of the tuples (mixed-radix decode in lexicographic order,
// "absent" first, validity = every present pair within the radius) and the unfused group normalisation are restated
// sequentially from k_rows / group_normalise (nwb_kernels.cuh), the latter with the kernels' 32 lane sums.
#define NWB_HOST_EMU 1
#include "emu_k0.h"
#include "../../nway_b200/csrc/nwb_rows.cuh"

#include <algorithm>
#include <cstring>

using namespace nwb;

namespace {

struct Item { int s; double sep, lon, slat, clat; };

double lanes_sum(const std::vector<double> &x, long long first)   // sum over k >= first as 32 lane sums + butterfly
{
	double lane_sum[32];
	for (int lane = 0; lane < 32; lane++) {
		double a = 0.0;
		for (long long k = lane; k < (long long) x.size(); k += 32)
			if (k >= first) a += x[k];
		lane_sum[lane] = a;
	}
	for (int o = 16; o > 0; o >>= 1) {
		double nxt[32];
		for (int i = 0; i < 32; i++) nxt[i] = lane_sum[i] + lane_sum[i ^ o];
		memcpy(lane_sum, nxt, sizeof(nxt));
	}
	return lane_sum[0];
}

template <int NC>
long long run(const int *n, const double *const *ra, const double *const *dec, const double *const *err, double radius,
	const ConstTables &T, double ratio_secondary, int nmag, int flat_hash, long long max_rows, long long *const *idx, double *const *sep_out, double *sepmax,
	long long *ncat, double *lbf_u, double *lbf_c, double *dist_post, double *const *bias_out, double *p_single, long long *flag, double *p_any, double *p_i)
{
	const int np = n[0];
	emu::K0 K;
	emu::build_k0(K, np, ra[0], dec[0], radius, 0);
	if (!K.ok) return -1;
	const Grid &G = K.G;
	std::vector<PrimRec> prec(np);
	for (int i = 0; i < np; i++) {
		double sl, cl;
		sincos_ref(deg2rad_ref(dec[0][i]), &sl, &cl);
		prec[i].lon = deg2rad_ref(ra[0][i]); prec[i].slat = sl; prec[i].clat = cl; prec[i].ij = 0;
	}
	std::vector<std::vector<Item>> L[NC];
	const double nbands_d = (double) G.nbands;
	for (int c = 1; c < NC; c++) {
		L[c].resize(np);
		for (int s = 0; s < n[c]; s++) {
			const double r = ra[c][s], d = dec[c][s];
			const double t = k1_band_coord(G, d);
			if (!(t >= 0.0 && t < nbands_d)) continue;
			const double x = k1_ra_coord(G, r);
			if (!(G.full_circle || x <= G.ra_span)) continue;
			const int b = __double2int_rd(t);
			const BandRec B = load_band(G, b);
			int ic;
			const double xcells = k1_ra_cell(B, x, ic);
			const CellRec &cr = K.cells[B.base + ic];
			const int ecnt = (int) (unsigned) cr.q[0], estart = (int) (cr.q[0] >> 32);
			const float xr = (float) (xcells - (double) ic), yr = (float) (t - (double) b);
			for (int e = 0; e < ecnt; e++) {
				int p;
				bool pass;
				if (e < 3) { p = (int) (cr.q[1 + e] >> 32); pass = k1_pretest_packed(G, xr, yr, G.kx[b], (unsigned) cr.q[1 + e]); }
				else { const Entry &en = K.entries[estart + e]; p = en.p; pass = k1_pretest(G, (float) k1_ra_coord(G, r), (float) (d - G.dec_lo), en.x, en.y, en.clat); }
				if (!pass) continue;
				Item it;
				sincos_ref(deg2rad_ref(d), &it.slat, &it.clat);
				it.lon = deg2rad_ref(r);
				it.sep = sep_arcsec_ref(prec[p].lon, prec[p].slat, prec[p].clat, it.lon, it.slat, it.clat);
				it.s = s;
				if (it.sep < radius) L[c][p].push_back(it);
			}
		}
	}
	RowParams RP;
	memset(&RP, 0, sizeof(RP));
	RP.ncat = NC; RP.nmag = nmag; RP.np = np; RP.ratio_secondary = ratio_secondary; RP.T = &T;
	for (int j = 0; j < nmag; j++) RP.C.bias[j] = bias_out[j];
	long long row = 0;
	std::vector<double> v;
	for (int p = 0; p < np; p++) {
		int nl[NC];
		long long ntup = 1;
		nl[0] = 0;
		for (int c = 1; c < NC; c++) {
			std::sort(L[c][p].begin(), L[c][p].end(), [](const Item &a, const Item &b) { return a.s < b.s; });
			nl[c] = (int) L[c][p].size();
			ntup *= nl[c] + 1;
		}
		const long long r0 = row;
		for (long long t = 0; t < ntup; t++) {
			int dg[NC];
			long long sidx[NC];
			double sep[NC * (NC - 1) / 2 + 1], sig[NC];
			unsigned present = 1u;
			sidx[0] = p; sig[0] = err[0][p];
			long long rem = t;
			for (int c = NC - 1; c >= 1; c--) { dg[c] = (int) (rem % (nl[c] + 1)); rem /= (nl[c] + 1); }
			bool ok = true;
			for (int c = 1; c < NC; c++) {
				if (dg[c] > 0) { const Item &it = L[c][p][dg[c] - 1]; sidx[c] = it.s; sep[pair_index(0, c, NC)] = it.sep; present |= 1u << c; }
				else { sidx[c] = -1; sep[pair_index(0, c, NC)] = nan(""); }
			}
			for (int a = 1; a < NC; a++)
				for (int b = a + 1; b < NC; b++) {
					double s = nan("");
					if (dg[a] > 0 && dg[b] > 0) {
						const Item &A = L[a][p][dg[a] - 1], &B2 = L[b][p][dg[b] - 1];
						s = sep_arcsec_ref(A.lon, A.slat, A.clat, B2.lon, B2.slat, B2.clat);   // k_count_rows
						ok = ok && (s < radius);
					}
					sep[pair_index(a, b, NC)] = s;
				}
			if (ok && flat_hash) {
				// NWB_COMPAT_FLAT_HASH as planned: every pair of present members in the same or adjacent hash cells
				long long ci[NC], cj[NC];
				for (int c = 0; c < NC; c++)
					if (present >> c & 1u) { ci[c] = flat_hash_cell(ra[c][sidx[c]], radius / 3600.0); cj[c] = flat_hash_cell(dec[c][sidx[c]], radius / 3600.0); }
				for (int a = 0; a < NC; a++)
					for (int b = a + 1; b < NC; b++)
						if ((present >> a & 1u) && (present >> b & 1u)) ok = ok && flat_hash_same_bucket(ci[a], cj[a], ci[b], cj[b]);
			}
			if (!ok) continue;
			if (row >= max_rows) { row++; continue; }
			double smax = 0.0;
			for (int c = 0; c < NC; c++) idx[c][row] = sidx[c];
			for (int k = 0; k < NC * (NC - 1) / 2; k++) { sep_out[k][row] = sep[k]; if (sep[k] > smax) smax = sep[k]; }
			sepmax[row] = smax;
			ncat[row] = __popc(present);
			for (int c = 1; c < NC; c++) if (present >> c & 1u) sig[c] = err[c][sidx[c]];
			const double lb = log_bf_ref<NC>(&T, NC, present, sig, sep, false);
			const unsigned smask = present >> 1;
			const double prior = T.prior[smask], l10p = T.log10prior[smask];
			lbf_u[row] = lb; lbf_c[row] = lb;
			dist_post[row] = posterior_ref(prior, l10p, lb);
			const double total = lb + row_bias(RP, row, sidx);
			p_single[row] = posterior_ref(prior, l10p, total);
			p_i[row] = total + l10p;
			row++;
		}
		if (row > max_rows) continue;
		// group_normalise
		const long long nrow = row - r0;
		v.assign(p_i + r0, p_i + row);
		double m_all = -INFINITY, m_rest = -INFINITY;
		for (long long k = 0; k < nrow; k++) { m_all = fmax(m_all, v[k]); if (k > 0) m_rest = fmax(m_rest, v[k]); }
		std::vector<double> ea(nrow), er(nrow);
		for (long long k = 0; k < nrow; k++) { ea[k] = exp10(v[k] - m_all); er[k] = k > 0 ? exp10(v[k] - m_rest) : 0.0; }
		const double s_all = lanes_sum(ea, 0), s_rest = lanes_sum(er, 1);
		const double bfsum = log10(s_all) + m_all;
		const double bfsum1 = nrow > 1 ? log10(s_rest) + m_rest : 0.0;
		const double pa = 1 - exp10(v[0] - bfsum);
		double best = 0.0;
		for (long long k = 0; k < nrow; k++) {
			const double pi = k == 0 ? 0.0 : exp10(v[k] - bfsum1);
			p_i[r0 + k] = pi; p_any[r0 + k] = pa;
			best = fmax(best, pi);
		}
		for (long long k = 0; k < nrow; k++) {
			const double pi = p_i[r0 + k];
			flag[r0 + k] = (pi == best) ? 1 : (pi > ratio_secondary * best ? 2 : 0);
		}
	}
	return row <= max_rows ? row : -row - 1;
}

}  // namespace

extern "C" {

// mag tables: mag_cat[j] = catalogue of magnitude column j, columns in catalogue order
long long nwb_emu_matchn(int ncat, const int *n, const double *const *ra, const double *const *dec, const double *const *err, double radius,
	const double *norm, double log10e, const double *prior, const double *log10prior, double ratio_secondary,
	int nmag, const int *mag_cat, const double *const *mag, const int *nbins, const double *const *edges, const double *const *weight,
	const double *const *biasval, double *const *bias_out, int flat_hash, long long max_rows, long long *const *idx, double *const *sep_out, double *sepmax,
	long long *ncat_out, double *lbf_u, double *lbf_c, double *dist_post, double *p_single, long long *flag, double *p_any, double *p_i)
{
	static ConstTables T;
	memset(&T, 0, sizeof(T));
	for (int k = 0; k <= ncat; k++) T.norm[k] = norm[k];
	T.log10e = log10e;
	for (int m = 0; m < (1 << (ncat - 1)); m++) { T.prior[m] = prior[m]; T.log10prior[m] = log10prior[m]; }
	for (int j = 0; j < nmag; j++) {
		MagTable &MT = T.mag[j];
		MT.cat = mag_cat[j]; MT.nbins = nbins[j]; MT.mag = mag[j];
		for (int k = 0; k <= nbins[j]; k++) MT.edges[k] = edges[j][k];
		for (int k = 0; k < nbins[j]; k++) { MT.weight[k] = weight[j][k]; MT.bias[k] = biasval[j][k]; }
	}
#define NWB_RUN(NC) run<NC>(n, ra, dec, err, radius, T, ratio_secondary, nmag, flat_hash, max_rows, idx, sep_out, sepmax, ncat_out, lbf_u, lbf_c, dist_post, bias_out, p_single, flag, p_any, p_i)
	switch (ncat) {
		case 2: return NWB_RUN(2);
		case 3: return NWB_RUN(3);
		case 4: return NWB_RUN(4);
	}
	return -1;
}

}  // extern "C"
