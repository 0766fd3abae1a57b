// Host build of the device arithmetic (nway_b200/csrc/nwb_device.cuh, the same source the CUDA library is compiled
// from; tests/emu/nwb_host_emu.h supplies the CUDA names): the separation in the reference's operation order with its
// small-angle polynomial shortcuts, the division-free x/180 and x/pi, the table-driven 10^x, the log Bayes factor and
// the posterior.  tests/test_device_arithmetic_cpu.py compares them with numpy / the oracle on millions of arguments.
#define NWB_HOST_EMU 1
#include "../../nway_b200/csrc/nwb_device.cuh"

using namespace nwb;

extern "C" {

void nwb_emu_sep(long long n, const double *ra1, const double *dec1, const double *ra2, const double *dec2, double *out)
{
	for (long long i = 0; i < n; i++) {
		double s1, c1, s2, c2;
		sincos_ref(deg2rad_ref(dec1[i]), &s1, &c1);
		sincos_ref(deg2rad_ref(dec2[i]), &s2, &c2);
		out[i] = sep_arcsec_ref(deg2rad_ref(ra1[i]), s1, c1, deg2rad_ref(ra2[i]), s2, c2);
	}
}

// sin_ref / cos_ref: glibc's double-precision sin / cos restated for the device
void nwb_emu_sincos(long long n, const double *x, double *s, double *c)
{
	for (long long i = 0; i < n; i++) sincos_ref(x[i], &s[i], &c[i]);
}

// the small-angle entry point (longitude differences) and the complete functions
void nwb_emu_sincos_small(long long n, const double *x, double *s, double *c)
{
	for (long long i = 0; i < n; i++) sincos_ref_small(x[i], &s[i], &c[i]);
}

void nwb_emu_sincos_any(long long n, const double *x, double *s, double *c)
{
	for (long long i = 0; i < n; i++) { s[i] = sin_ref(x[i]); c[i] = cos_ref(x[i]); }
}

void nwb_emu_div(long long n, const double *x, double *by180, double *bypi)
{
	for (long long i = 0; i < n; i++) {
		by180[i] = div_const(x[i], 180.0, NWB_INV180);
		bypi[i] = div_const(x[i], NWB_PI, NWB_INVPI);
	}
}

void nwb_emu_quotient(long long n, const double *a, const double *b, double *out)
{
	for (long long i = 0; i < n; i++) out[i] = quotient_by_reciprocal(a[i], b[i], 1.0 / b[i]);
}

void nwb_emu_exp10(long long n, const double *x, double *out)
{
	for (long long i = 0; i < n; i++) out[i] = nwb_exp10(x[i]);
}

// sig: n x ncat, sep: n x npairs (pair order of pair_index), present: n masks
void nwb_emu_log_bf(long long n, int ncat, const double *norm, double log10e, const unsigned *present, const double *sig, const double *sep, double *out)
{
	ConstTables T;
	memset(&T, 0, sizeof(T));
	for (int k = 0; k <= ncat; k++) T.norm[k] = norm[k];
	T.log10e = log10e;
	const int np = ncat * (ncat - 1) / 2;
	for (long long i = 0; i < n; i++) {
		const double *s = sig + i * ncat, *p = sep + i * np;
		switch (ncat) {
			case 2: out[i] = log_bf_ref<2>(&T, ncat, present[i], s, p); break;
			case 3: out[i] = log_bf_ref<3>(&T, ncat, present[i], s, p); break;
			default: out[i] = log_bf_ref<0>(&T, ncat, present[i], s, p); break;
		}
	}
}

void nwb_emu_posterior(long long n, const double *prior, const double *log10prior, const double *lbf, double *out)
{
	for (long long i = 0; i < n; i++) out[i] = posterior_ref(prior[i], log10prior[i], lbf[i]);
}

// offsets (arcsec) of the target in the tangent frame of the origin: offsets_ref
void nwb_emu_offsets(long long n, const double *ra_o, const double *dec_o, const double *ra_t, const double *dec_t, double *dra, double *ddec)
{
	for (long long i = 0; i < n; i++) offsets_ref(ra_o[i], dec_o[i], ra_t[i], dec_t[i], dra[i], ddec[i]);
}

// two catalogues with elliptical errors (sigma_x, sigma_y, rho): circularised errors, rescaled separation, log_bf_ref<2>
// -- the per-row work of ell_prepare + the row kernels (nwb_kernels.cuh)
void nwb_emu_log_bf_ell2(long long n, const double *norm, double log10e, const double *vx, const double *vy,
	const double *ea /* n x 3 */, const double *eb /* n x 3 */, double *out)
{
	ConstTables T;
	memset(&T, 0, sizeof(T));
	for (int k = 0; k <= 2; k++) T.norm[k] = norm[k];
	T.log10e = log10e;
	for (long long i = 0; i < n; i++) {
		const double *a = ea + 3 * i, *b = eb + 3 * i;
		double sig[2] = {sqrt((a[0] * a[0] + a[1] * a[1]) / 2), sqrt((b[0] * b[0] + b[1] * b[1]) / 2)};
		double sep[1] = {ell_rescaled_sep(vx[i], vy[i], a[0], a[1], a[2], b[0], b[1], b[2], sig[0], sig[1])};
		out[i] = log_bf_ref<2>(&T, 2, 3u, sig, sep);
	}
}

// zero-order-hold magnitude prior: mag_weight
void nwb_emu_mag_weight(long long n, int nbins, const double *edges, const double *weight, const double *bias, const double *m, double *w, double *b)
{
	static MagTable M;
	M.cat = 1; M.nbins = nbins; M.mag = nullptr;
	for (int k = 0; k <= nbins; k++) M.edges[k] = edges[k];
	for (int k = 0; k < nbins; k++) { M.weight[k] = weight[k]; M.bias[k] = bias[k]; }
	for (long long i = 0; i < n; i++) w[i] = mag_weight(M, m[i], b[i]);
}

// One group (one primary) through the fused normalisation of k_rows2<FUSE, SHARE> (nwb_kernels.cuh): rows 1.. are
// distributed over 32 lanes (k = lane + 32 j), every lane sums its own exponentials, warp_sum adds the lanes in its
// butterfly order; then group_p_any / shared_post (nwb_device.cuh) per row.  v = log-weights, lbf = log Bayes factors.
void nwb_emu_group(int rows, const double *v, const double *lbf, double prior1, double l10p1, double ratio_secondary,
	double *p_any_out, double *p_i, long long *flag, double *post)
{
	const double v0 = v[0];
	double m_rest = -INFINITY;
	for (int k = 1; k < rows; k++) m_rest = fmax(m_rest, v[k]);
	double lane_sum[32];
	double t[4096];
	for (int lane = 0; lane < 32; lane++) {
		double s = 0.0;
		for (int k = lane + (lane == 0 ? 32 : 0); k < rows; k += 32) {
			t[k] = nwb_exp10(v[k] - m_rest);
			s += t[k];
		}
		lane_sum[lane] = s;
	}
	for (int o = 16; o > 0; o >>= 1) {   // warp_sum: v += shfl_xor(v, o)
		double nxt[32];
		for (int i = 0; i < 32; i++) nxt[i] = lane_sum[i] + lane_sum[i ^ o];
		for (int i = 0; i < 32; i++) lane_sum[i] = nxt[i];
	}
	const double s_rest = lane_sum[0];
	double p_any, rinv;
	group_p_any(rows, v0, m_rest, s_rest, p_any, rinv);
	const double best = rinv;
	const bool direct = !(fabs(m_rest) <= 250.0);
	const double oscale = (rows > 1 && !direct) ? (1 - prior1) * nwb_exp10(-m_rest) : 0.0;
	const double omp = 1 - prior1;
	for (int k = 0; k < rows; k++) {
		const double tk = k == 0 ? v0 : t[k];
		const double pi = k == 0 ? 0.0 : tk * rinv;
		p_i[k] = pi;
		flag[k] = (pi == best) ? 1 : (pi > ratio_secondary * best ? 2 : 0);
		post[k] = k == 0 ? 1.0 : shared_post(direct, tk, oscale, omp, lbf[k], l10p1);
	}
	*p_any_out = p_any;
}

}  // extern "C"
