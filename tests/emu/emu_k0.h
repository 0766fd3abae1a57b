// The primary side of the match on the host, one thread after the other: bounding box (k_prim_prep), grid geometry and
// pre-test constants (build_grid / pretest_constants, as nwb_api.cu calls them), count pass, cell headers, fill pass
// (prim_register: the device's own function).  Shared by grid_emu.cpp and rows_emu.cpp.
#pragma once
#include "../../nway_b200/csrc/nwb_grid_host.h"

#include <cstdio>
#include <vector>

namespace emu {

using namespace nwb;

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


struct K0 {
	HostGrid HG;
	Grid G;
	std::vector<CellRec> cells;
	std::vector<Entry> entries;
	std::vector<double> rn, dra;
	long long registrations = 0;
	bool ok = true;
};

// max_cells_override > 0 replaces the library's own choice of the cell budget (coarser grids: crowded cells)
inline void build_k0(K0 &K, int np, const double *pra, const double *pdec, double radius_arcsec, long long max_cells_override, bool one_pass = true)
{
	std::vector<double> &rn = K.rn, &dra = K.dra;
	HostGrid &HG = K.HG;
	std::vector<CellRec> &cells = K.cells;
	std::vector<Entry> &entries = K.entries;
	const double r_deg = radius_arcsec / 3600.0;
	const double rb = r_deg * (1 + 1e-9) + 1e-12;
	const double rb_ins = rb + 1e-9, dra_eps = 1e-9;
	double red[6];
	bounding_box(np, pra, pdec, rb, rn, dra, red);
	long long max_cells = std::min<long long>(2ll << 20, std::max<long long>(1ll << 16, 16 * (long long) np));   // nwb_api.cu match_impl
	if (max_cells_override > 0) max_cells = max_cells_override;
	build_grid(red, rb_ins, rb_ins, max_cells, HG);
	pretest_constants(HG, rb_ins);
	Grid &G = K.G;
	G = HG.g;
	G.bands = HG.bands.data();
	G.kx = HG.kx.data();
	G.bits = nullptr;
	const double entry_tau_max = (rb_ins * M_PI / 180 > 0.02) ? -1.0 : 0.02;
	// K0, one thread after the other.  one_pass (the steady state: grid geometry known from the previous match):
	// k_prim_prep<COUNT> counts AND places (prim_register<REG_COUNT_INLINE>: first three of a cell inline, the rest into a
	// work list), k_cell_headers, k_fill_overflow.  Otherwise (first match of a context): k_prim_cells<false> counts,
	// k_cell_headers, k_prim_cells<true> fills by counting back down.
	std::vector<int> cellcnt(G.ncells + 1, 0);
	cells.assign(G.ncells, CellRec());
	std::vector<double> clat(np);
	std::vector<OverflowItem> worklist((size_t) np * 64 + 64);
	int worklist_n = 0;
	for (int i = 0; i < np; i++) {
		const double cl = cos(deg2rad_ref(pdec[i]));
		const double tau = (rb_ins / 180 * NWB_PI) * tan(fmin(fabs(pdec[i]), 89.9999) / 180 * NWB_PI);
		clat[i] = (tau > entry_tau_max || dra[i] >= 180.0) ? 0.0 : (double) __double2float_rd(cl);
		if (one_pass)
			prim_register<REG_COUNT_INLINE>(G, i, pdec[i], rn[i], dra[i], cl, rb_ins, dra_eps, 0, 1, cellcnt.data(), cells.data(), nullptr,
				worklist.data(), &worklist_n, (long long) worklist.size());
		else
			prim_register<REG_COUNT>(G, i, pdec[i], rn[i], dra[i], cl, rb_ins, dra_eps, 0, 1, cellcnt.data(), nullptr, nullptr);
	}
	long long total = 0, regs = 0;
	for (long long c = 0; c < G.ncells; c++) {
		const int cnt = cellcnt[c];
		const int start = (int) total - 3;
		total += cnt > 3 ? cnt - 3 : 0;
		regs += cnt;
		cells[c].q[0] = (unsigned long long) (unsigned) cnt | ((unsigned long long) (unsigned) start << 32);   // q[1..3] stay as placed
	}
	entries.assign(total + 1, Entry());
	if (one_pass) {
		if ((long long) worklist_n != total || worklist_n > (long long) worklist.size()) { fprintf(stderr, "emu: work list %d, overflow entries %lld\n", worklist_n, total); K.ok = false; return; }
		for (int k = 0; k < worklist_n; k++) {   // k_fill_overflow
			const OverflowItem it = worklist[k];
			double x = rn[it.p] - G.ra_org_n;
			if (x < 0.0) x += 360.0;
			Entry en;
			en.x = (float) x; en.y = (float) (pdec[it.p] - G.dec_lo); en.clat = (float) clat[it.p]; en.p = it.p;
			entries[(int) (cells[it.cell].q[0] >> 32) + it.slot] = en;
		}
	} else {
		for (int i = 0; i < np; i++)
			for (int bslot = 0; bslot < 4; bslot++)   // the four threads of a primary in k_prim_cells
				prim_register<REG_FILL>(G, i, pdec[i], rn[i], dra[i], clat[i], rb_ins, dra_eps, bslot, 4, cellcnt.data(), cells.data(), entries.data());
		for (long long c = 0; c < G.ncells; c++)
			if (cellcnt[c] != 0) { fprintf(stderr, "emu: count and fill disagree in cell %lld\n", c); K.ok = false; return; }
	}
	K.registrations = regs;
}

}  // namespace emu
