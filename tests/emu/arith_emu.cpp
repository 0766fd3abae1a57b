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
		sincos(deg2rad_ref(dec1[i]), &s1, &c1);
		sincos(deg2rad_ref(dec2[i]), &s2, &c2);
		out[i] = sep_arcsec_ref(deg2rad_ref(ra1[i]), s1, c1, deg2rad_ref(ra2[i]), s2, c2);
	}
}

void nwb_emu_div(long long n, const double *x, double *by180, double *bypi)
{
	for (long long i = 0; i < n; i++) {
		by180[i] = div_const(x[i], 180.0, NWB_INV180);
		bypi[i] = div_const(x[i], NWB_PI, NWB_INVPI);
	}
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

}  // extern "C"
