// Host side of the grid: the band / cell geometry chosen from the primaries' bounding box and the constants of the fp32
// pre-tests.  Plain C++ (no CUDA calls), shared by the library (nwb_api.cu) and the host emulation that checks the
// pre-tests for completeness (tests/emu/grid_emu.cpp).
#pragma once
#include "nwb_grid.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace nwb {

// choose the band grid from the primaries' bounding box (host side, tiny)
struct HostGrid {
	Grid g;
	std::vector<BandRec> bands;
	std::vector<float> kx;   // per band, for the packed pre-test (struct PEntry)
};

inline void build_grid(const double red[6], double rb_ins, double cell_min_deg, long long max_cells, HostGrid &H)
{
	const double margin = 1e-6;
	double dec_lo = red[0] - rb_ins - margin, dec_hi = red[1] + rb_ins + margin;
	dec_lo = std::max(dec_lo, -90.0 - margin);
	dec_hi = std::min(dec_hi, 90.0 + margin);
	double spanA = red[3] - red[2], spanB = red[5] - red[4];
	Grid &g = H.g;
	if (std::min(spanA, spanB) + 2 * margin >= 359.0) {
		g.full_circle = 1; g.ra_org = 0.0; g.ra_span = 360.0;
	} else if (spanA <= spanB) {
		g.full_circle = 0; g.ra_org = red[2] - margin; g.ra_span = spanA + 2 * margin;
	} else {
		g.full_circle = 0; g.ra_org = red[4] - 180.0 - margin; g.ra_span = spanB + 2 * margin;
	}
	double dspan = dec_hi - dec_lo;
	double s = std::max(cell_min_deg, std::sqrt(dspan * g.ra_span / (double) max_cells));
	s = std::max(s, dspan / 1048576.0);
	int nb = (int) std::ceil(dspan / s);
	if (nb < 1) nb = 1;
	g.dec_lo = dec_lo;
	g.inv_h = 1.0 / s;
	g.nbands = nb;
	H.bands.resize(nb);
	long long tot = 0;
	for (int b = 0; b < nb; b++) {
		double mid = dec_lo + (b + 0.5) * s;
		double c = std::cos(std::min(std::fabs(mid), 90.0) * M_PI / 180);
		double w = s / std::max(c, 1e-6);
		long long n = (long long) std::floor(g.ra_span / w);
		n = std::max<long long>(1, std::min<long long>(n, 1 << 24));
		H.bands[b].nra = (int) n;
		H.bands[b].base = (int) tot;
		H.bands[b].inv_w = (double) n / g.ra_span;
		tot += n;
	}
	g.ncells = tot;
}

// constants of the fp32 flat pre-test (see struct Entry): rr2 and the pole cut-off tau_max
inline void pretest_constants(HostGrid &H, double rb_deg)
{
	Grid &g = H.g;
	const double theta = rb_deg * M_PI / 180;
	double tau_max = 0.02;
	double kappa;
	if (theta > 0.02) {
		tau_max = -1.0;   // large radii: every primary is pre-tested on declination only
		kappa = 0.0;
	} else {
		double f = (1 - theta * theta / 2 - tau_max) * (1 - (theta + tau_max) * (theta + tau_max) / 12);
		kappa = 1 / std::sqrt(f) - 1 + 1e-6;
	}
	const double dspan = (double) g.nbands / g.inv_h;
	const double slack = std::ldexp(std::max(g.ra_span, dspan), -21) + 1e-9;   // fp32 rounding of the relative coordinates
	const double rr = rb_deg * (1 + kappa) + 1.5 * slack;
	g.rr2 = std::nextafter((float) (rr * rr * (1 + 1e-6)), INFINITY);
	g.ra_org_n = g.ra_org - 360.0 * std::floor(g.ra_org / 360.0);
	if (g.ra_org_n >= 360.0 || g.ra_org_n < 0.0) g.ra_org_n = 0.0;
	g.tau_max = tau_max;
	// packed pre-test: per band kx = (smallest cos(dec) any primary registered in the band can have) x (cell width in
	// degrees), rounded down -- a smaller kx only makes the test more permissive.  Primaries registered in band b lie
	// within rb of it.  Bands that reach the zone where the flat metric is not a safe bound (tau > tau_max, or a search
	// box spanning all of ra) are tested on declination only: kx = 0.
	const double s = 1.0 / g.inv_h;
	H.kx.resize(g.nbands);
	double cell_true_max = s;
	for (int b = 0; b < g.nbands; b++) {
		double lo = g.dec_lo + b * s - (rb_deg + 1e-8), hi = g.dec_lo + (b + 1) * s + (rb_deg + 1e-8);
		double amax = std::max(std::fabs(lo), std::fabs(hi));
		bool polar = tau_max < 0 || amax + rb_deg >= 89.99 || theta * std::tan(std::min(amax, 89.9999) * M_PI / 180) > tau_max;
		float k = 0.f;
		if (!polar) {
			double v = std::cos(amax * M_PI / 180) * (1 - 1e-7) / H.bands[b].inv_w;
			k = (float) v;
			if ((double) k > v) k = std::nextafter(k, 0.f);
			const double amin = (lo < 0 && hi > 0) ? 0.0 : std::min(std::fabs(lo), std::fabs(hi));
			cell_true_max = std::max(cell_true_max, std::cos(amin * M_PI / 180) / H.bands[b].inv_w);
		}
		H.kx[b] = k;
	}
	g.hdeg = std::nextafter((float) s, 0.f);   // rounded down: only more permissive
	// quantisation of the packed entries: half a step of 4/32768 cell widths in x and of 4/65536 band heights in y, plus
	// the fp32 rounding of a handful of O(1) quantities
	const double quant = std::sqrt(6.2e-5 * 6.2e-5 + 3.1e-5 * 3.1e-5) * std::max(cell_true_max, s) + 4e-6 * std::max(cell_true_max, s);
	const double rrp = rb_deg * (1 + kappa) + quant + 1e-9;
	g.rr2p = std::nextafter((float) (rrp * rrp * (1 + 1e-6)), INFINITY);
}

}  // namespace nwb
