// Host emulation of what decides WHICH pairs reach the exact test on the device: the grid chosen from the primaries
// (build_grid / pretest_constants, nwb_grid_host.h), the registration of every primary in the cells its search box
// overlaps and the packed / fp32 cell entries (prim_register, nwb_grid.cuh), the cell a secondary falls into (k1_band_coord /
// k1_ra_coord / k1_ra_cell: the functions k_pairs' first stage calls) and the two fp32 pre-tests.  The device
// functions are compiled here from the SAME headers the library is built from (tests/emu/nwb_host_emu.h provides the
// handful of CUDA names they use).
//
// check(): for every given (primary, secondary) pair whose exact separation (sep_arcsec_ref, the reference's
// arithmetic) is below the radius, the secondary's cell must hold an entry of that primary and the entry must pass its
// pre-test -- otherwise the device would silently lose a row.  Returns the number of such misses.
#define NWB_HOST_EMU 1
#include "../../nway_b200/csrc/nwb_grid_host.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace nwb;

namespace {

// order-preserving bounding-box reduction of k_prim_prep (nwb_kernels.cuh), as plain min / max
void bounding_box(int np, const double *ra, const double *dec, double rb, std::vector<double> &rn, std::vector<double> &dra, double red[6])
{
	double v[6] = {1e300, -1e300, 1e300, -1e300, 1e300, -1e300};
	rn.resize(np); dra.resize(np);
	for (int i = 0; i < np; i++) {
		double r = ra[i], d = dec[i];
		rn[i] = wrap360(r);
		dra[i] = search_box_dra(d, rb);
		double rn_b = wrap360(rn[i] + 180.0);
		v[0] = std::min(v[0], d); v[1] = std::max(v[1], d);
		v[2] = std::min(v[2], rn[i] - dra[i]); v[3] = std::max(v[3], rn[i] + dra[i]);
		v[4] = std::min(v[4], rn_b - dra[i]); v[5] = std::max(v[5], rn_b + dra[i]);
	}
	for (int k = 0; k < 6; k++) red[k] = v[k];
}

}  // namespace

extern "C" {

// Returns the number of pairs (pi[k], si[k]) with exact separation < radius that the grid + pre-tests would not hand to
// the exact stage; stats[0] = pairs within the radius, [1] = of those found through an inline (packed) entry, [2] =
// through an overflow entry, [3] = grid cells, [4] = bands, [5] = registrations.  first_miss[0..1] = the first missed pair.
long long nwb_emu_check(int np, const double *pra, const double *pdec, int ns, const double *sra, const double *sdec,
	double radius_arcsec, long long npairs, const int *pi, const int *si, long long max_cells_override, long long *stats, int *first_miss)
{
	const double r_deg = radius_arcsec / 3600.0;
	const double rb = r_deg * (1 + 1e-9) + 1e-12;
	const double rb_ins = rb + 1e-9, dra_eps = 1e-9;
	std::vector<double> rn, dra;
	double red[6];
	bounding_box(np, pra, pdec, rb, rn, dra, red);
	HostGrid HG;
	long long max_cells = std::min<long long>(2ll << 20, std::max<long long>(1ll << 16, 16 * (long long) np));   // nwb_api.cu match_impl
	if (max_cells_override > 0) max_cells = max_cells_override;
	build_grid(red, rb_ins, rb_ins, max_cells, HG);
	pretest_constants(HG, rb_ins);
	Grid G = HG.g;
	G.bands = HG.bands.data();
	G.kx = HG.kx.data();
	G.bits = nullptr;
	const double entry_tau_max = (rb_ins * M_PI / 180 > 0.02) ? -1.0 : 0.02;
	// K0: count, headers, fill (k_prim_prep<COUNT> / k_cell_headers / k_prim_cells<true>, one thread after the other)
	std::vector<int> cellcnt(G.ncells + 1, 0);
	std::vector<CellRec> cells(G.ncells);
	std::vector<double> clat(np);
	for (int i = 0; i < np; i++) {
		const double cl = cos(deg2rad_ref(pdec[i]));
		const double tau = (rb_ins / 180 * NWB_PI) * tan(fmin(fabs(pdec[i]), 89.9999) / 180 * NWB_PI);
		clat[i] = (tau > entry_tau_max || dra[i] >= 180.0) ? 0.0 : (double) __double2float_rd(cl);
		prim_register<false>(G, i, pdec[i], rn[i], dra[i], cl, rb_ins, dra_eps, 0, 1, cellcnt.data(), nullptr, nullptr);
	}
	long long total = 0, regs = 0;
	for (long long c = 0; c < G.ncells; c++) {
		const int cnt = cellcnt[c];
		const int start = (int) total - 3;
		total += cnt > 3 ? cnt - 3 : 0;
		regs += cnt;
		cells[c].q[0] = (unsigned long long) (unsigned) cnt | ((unsigned long long) (unsigned) start << 32);
	}
	std::vector<Entry> entries(total + 1);
	for (int i = 0; i < np; i++)
		for (int bslot = 0; bslot < 4; bslot++)   // the four threads of a primary in k_prim_cells
			prim_register<true>(G, i, pdec[i], rn[i], dra[i], clat[i], rb_ins, dra_eps, bslot, 4, cellcnt.data(), cells.data(), entries.data());
	for (long long c = 0; c < G.ncells; c++)
		if (cellcnt[c] != 0) { fprintf(stderr, "grid_emu: count and fill disagree in cell %lld\n", c); return -1; }
	stats[0] = stats[1] = stats[2] = 0;
	stats[3] = G.ncells; stats[4] = G.nbands; stats[5] = regs;
	long long misses = 0;
	const double nbands_d = (double) G.nbands;
	for (long long k = 0; k < npairs; k++) {
		const int p = pi[k], s = si[k];
		const double r = sra[s], d = sdec[s];
		// the exact stage of k1_flush
		double sl1, cl1, sl2, cl2;
		sincos(deg2rad_ref(pdec[p]), &sl1, &cl1);
		sincos(deg2rad_ref(d), &sl2, &cl2);
		const double sep = sep_arcsec_ref(deg2rad_ref(pra[p]), sl1, cl1, deg2rad_ref(r), sl2, cl2);
		if (!(sep < radius_arcsec)) continue;
		stats[0]++;
		// first stage of k_pairs (nwb_kernels.cuh): the cell of the secondary and its position inside it
		bool found = false;
		const double t = k1_band_coord(G, d);
		if (t >= 0.0 && t < nbands_d) {
			const double x = k1_ra_coord(G, r);
			if (G.full_circle || x <= G.ra_span) {
				const int b = __double2int_rd(t);
				const BandRec B = load_band(G, b);
				const float kx = G.kx[b];
				int ic;
				const double xcells = k1_ra_cell(B, x, ic);
				const int cell = B.base + ic;
				const CellRec &cr = cells[cell];
				const int ecnt = (int) (unsigned) cr.q[0], estart = (int) (cr.q[0] >> 32);
				const float xr = (float) (xcells - (double) ic), yr = (float) (t - (double) b);
				for (int e = 0; e < ecnt && !found; e++) {
					if (e < 3) {
						if ((int) (cr.q[1 + e] >> 32) == p && k1_pretest_packed(G, xr, yr, kx, (unsigned) cr.q[1 + e])) { found = true; stats[1]++; }
					} else {
						// k1_items
						const Entry &en = entries[estart + e];
						const double xx = k1_ra_coord(G, r);
						if (en.p == p && k1_pretest(G, (float) xx, (float) (d - G.dec_lo), en.x, en.y, en.clat)) { found = true; stats[2]++; }
					}
				}
			}
		}
		if (!found) {
			if (misses == 0) { first_miss[0] = p; first_miss[1] = s; }
			misses++;
		}
	}
	return misses;
}

}  // extern "C"
