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
#include "emu_k0.h"

#include <cstdlib>

using namespace nwb;

static int g_one_pass = 1;

extern "C" {

// which K0 path build_k0 emulates: 1 = count + place in one pass (the steady state), 0 = count, headers, fill (first match)
void nwb_emu_set_one_pass(int v) { g_one_pass = v; }

// Returns the number of pairs (pi[k], si[k]) with exact separation < radius that the grid + pre-tests would not hand to
// the exact stage; stats[0] = pairs within the radius, [1] = of those found through an inline (packed) entry, [2] =
// through an overflow entry, [3] = grid cells, [4] = bands, [5] = registrations.  first_miss[0..1] = the first missed pair.
long long nwb_emu_check(int np, const double *pra, const double *pdec, int ns, const double *sra, const double *sdec,
	double radius_arcsec, long long npairs, const int *pi, const int *si, long long max_cells_override, long long *stats, int *first_miss)
{
	emu::K0 K;
	emu::build_k0(K, np, pra, pdec, radius_arcsec, max_cells_override, g_one_pass != 0);
	if (!K.ok) return -1;
	const Grid &G = K.G;
	const std::vector<CellRec> &cells = K.cells;
	const std::vector<Entry> &entries = K.entries;
	const long long regs = K.registrations;
	stats[0] = stats[1] = stats[2] = 0;
	stats[3] = G.ncells; stats[4] = G.nbands; stats[5] = regs;
	long long misses = 0;
	const double nbands_d = (double) G.nbands;
	for (long long k = 0; k < npairs; k++) {
		const int p = pi[k], s = si[k];
		const double r = sra[s], d = sdec[s];
		// the exact stage of k1_flush
		double sl1, cl1, sl2, cl2;
		sincos_ref(deg2rad_ref(pdec[p]), &sl1, &cl1);
		sincos_ref(deg2rad_ref(d), &sl2, &cl2);
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
