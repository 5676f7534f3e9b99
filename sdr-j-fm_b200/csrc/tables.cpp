// Host-side table design; see tables.hpp for what follows which reference lines.
#include "tables.hpp"
#include <cmath>
#include <cstring>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace sdrjfm {

// Blackman-windowed sinc prototype shared by the reference's three designers.
// `f` is the (float) normalised cut-off.  Promotion as in fir-filters.cpp:46-58 /
// :202-213 / :332-343: the sinc and window are evaluated in double, the product is
// narrowed to float once per tap.
std::vector<float> design_sinc_blackman (int ntaps, float f) {
std::vector<float> t (ntaps);
const int mid = ntaps / 2;
	for (int i = 0; i < ntaps; i ++) {
	   float v;
	   if (i == mid)
	      v = (float)(2 * M_PI * f);
	   else
	      v = (float)(sin (2 * M_PI * f * (i - mid)) / (i - mid));
	   const double w = 0.42
	                  - 0.50 * cos (2 * M_PI * (float)i / (float)ntaps)
	                  + 0.08 * cos (4 * M_PI * (float)i / (float)ntaps);
	   t [i] = (float)(v * w);
	}
	return t;
}

static float float_sum (const std::vector<float> &t) {
float s = 0.0f;                       // accumulated in float, as the reference does
	for (float v : t) s += v;
	return s;
}

// DecimatingFIR::newKernel (int32_t low), fir-filters.cpp:327-347.  Note the reference's
// quirk at :346 — the imaginary part is NOT normalised.
std::vector<cf32> design_decimating_lowpass (int ntaps, int32_t low, int32_t fs) {
const float f = (float)low / fs;
std::vector<float> t = design_sinc_blackman (ntaps, f);
const float s = float_sum (t);
std::vector<cf32> k (ntaps);
	for (int i = 0; i < ntaps; i ++)
	   k [i] = cf32 (t [i] / s, t [i]);
	return k;
}

// LowPassFIR::newKernel, fir-filters.cpp:41-62
std::vector<cf32> design_lowpass (int ntaps, int32_t fc, int32_t fs) {
const float f = (float)fc / fs;
std::vector<float> t = design_sinc_blackman (ntaps, f);
const float s = float_sum (t);
std::vector<cf32> k (ntaps);
	for (int i = 0; i < ntaps; i ++)
	   k [i] = cf32 (t [i] / s, 0);
	return k;
}

// BandPassFIR::newKernel, fir-filters.cpp:197-222: low-pass prototype of half the band
// width, shifted to the band centre.
std::vector<cf32> design_bandpass (int ntaps, int32_t low, int32_t high, int32_t fs) {
const float lo    = (float)((high - low) / 2) / fs;
const float shift = (float)((high + low) / 2) / fs;
std::vector<float> t = design_sinc_blackman (ntaps, lo);
const float s = float_sum (t);
std::vector<cf32> k (ntaps);
const int mid = ntaps / 2;
	for (int i = 0; i < ntaps; i ++) {
	   const float v = (float)((i - mid) * (2 * M_PI * shift));
	   // v is a float in the reference, so cos/sin resolve to the float overloads (:219-220)
	   k [i] = cf32 (t [i] * cosf (v) / s, t [i] * sinf (v) / s);
	}
	return k;
}

// exp table of Fft_transformRadix2, fft-complex.cpp:65-71: the angle is formed in double,
// narrowed to float, and std::exp (complex<float>) evaluates cosf/sinf of that float.
std::vector<cf32> fft_twiddles (int n) {
std::vector<cf32> w (n / 2);
	for (int i = 0; i < n / 2; i ++) {
	   const float a = (float)(-2 * M_PI * i / n);
	   w [i] = cf32 (cosf (a), sinf (a));
	}
	return w;
}

// Radix-2 decimation-in-time FFT with the operation order of fft-complex.cpp:73-98
// (bit reversal, then stages of size 2,4,..n; butterfly temp = v[l]*w; v[l] = v[j]-temp;
// v[j] += temp), float32 throughout.  Used on the host only to transform filter kernels.
void fft_radix2_reference_order (cf32 *v, int n) {
int levels = 0;
	while ((1 << levels) < n) levels ++;
std::vector<cf32> w = fft_twiddles (n);
	for (int i = 0; i < n; i ++) {
	   int j = 0;
	   for (int b = 0; b < levels; b ++)
	      j |= ((i >> b) & 1) << (levels - 1 - b);
	   if (j > i) std::swap (v [i], v [j]);
	}
	for (int size = 2; size <= n; size *= 2) {
	   const int half = size / 2, step = n / size;
	   for (int i = 0; i < n; i += size)
	      for (int j = i, k = 0; j < i + half; j ++, k += step) {
	         const cf32 a = v [j + half], b = w [k];
	         // std::complex<float> product, 4 multiplies + 2 adds, no contraction
	         const cf32 t (a.real () * b.real () - a.imag () * b.imag (),
	                       a.real () * b.imag () + a.imag () * b.real ());
	         v [j + half] = v [j] - t;
	         v [j] += t;
	      }
	}
}

static std::vector<cf32> filter_spectrum (const std::vector<cf32> &taps, int nfft) {
std::vector<cf32> v (nfft, cf32 (0, 0));      // fft-filters.cpp:74-81 / :87-94
	for (size_t i = 0; i < taps.size (); i ++) v [i] = taps [i];
	fft_radix2_reference_order (v.data (), nfft);
	return v;
}

// Rational polyphase resampler input_rate -> fm_rate in two stages (resample.cuh):
//   A: 49-tap Blackman-windowed sinc, cut-off 0.1 fs, /5
//   B: L / M = 5 fm_rate / input_rate (reduced), prototype = Blackman-windowed sinc at rate L fa with
//      cut-off fm_rate / 2, L * P taps, P = ceil (8 M / L); phase phi holds h[phi + L j], unit DC gain
// Designed in double, stored as float.
static double blackman (int i, int n) {
	return 0.42 - 0.5 * cos (2 * M_PI * i / (n - 1)) + 0.08 * cos (4 * M_PI * i / (n - 1));
}
static double sinc_lp (double fc, double t) {       // ideal low-pass, cut-off fc (cycles / sample), at time t
	return t == 0.0 ? 2 * fc : sin (2 * M_PI * fc * t) / (M_PI * t);
}
bool design_resampler (int32_t input_rate, int32_t fm_rate, int &L, int &M, int &P,
                       std::vector<float> &hA, std::vector<float> &hB) {
const int DA = 5;
	if (input_rate % DA != 0) return false;
int64_t a = (int64_t)DA * fm_rate, b = input_rate;
	while (b) { int64_t t = a % b; a = b; b = t; }
	L = (int)((int64_t)DA * fm_rate / a); M = (int)(input_rate / a);
	P = (8 * M + L - 1) / L;
	if (L < 1 || L > 16 || P > 128 || M < L) return false;
const int NA = 49;
	{  std::vector<double> d (NA); double sum = 0;
	   for (int i = 0; i < NA; i ++) { d [i] = sinc_lp (0.1, i - (NA - 1) / 2.0) * blackman (i, NA); sum += d [i]; }
	   hA.resize (NA);
	   for (int i = 0; i < NA; i ++) hA [i] = (float)(d [i] / sum); }
const int NB = L * P;
const double fc = 0.5 * (double)fm_rate / ((double)input_rate / DA * L);
	hB.assign ((size_t)NB, 0.f);
	for (int phi = 0; phi < L; phi ++) {
	   std::vector<double> d (P); double sum = 0;
	   for (int j = 0; j < P; j ++) {
	      const int k = phi + L * j;
	      d [j] = sinc_lp (fc, k - (NB - 1) / 2.0) * blackman (k, NB); sum += d [j];
	   }
	   for (int j = 0; j < P; j ++) hB [(size_t)phi * P + j] = (float)(d [j] / sum);
	}
	return true;
}

// ---- squelch IIR design ----------------------------------------------------------------------
// Restates, with the reference's float / double promotions (DSPFLOAT = float; `using namespace std`
// picks the float overloads of sqrt / log / sinh / cosh / sin / cos for float arguments), the design
// path the two squelch filters take through src/various/iir-filters.cpp:
//   newChebyshev (:173-229, even order)  ->  low-pass (:469-500) / high-pass (:510-548) un-normalisation
//   ->  Bilineair (:80-117).  order 20 => 10 biquads; apass = -1 dB.
namespace {
struct Quad { float A0, A1, A2, B0, B1, B2; };

float asinh_ref (float x) { return logf (x + sqrtf (x * x + 1)); }           // sinhm1, :35-38

float chebyshev_even (Quad *q, int n, int order, int apass) {
const float Eps = sqrt (pow (10.0, -0.1 * apass) - 1);
const float D = asinh_ref (1.0 / Eps) / order;
const float sinhD = sinhf (D), coshD = coshf (D);
	for (int i = 0; i < n; i ++) {
	   const float Phim = (M_PI * (2 * i + 1) / (2 * order));
	   const float sigma = - sinhD * sinf (Phim);
	   const float omega =   coshD * cosf (Phim);
	   q [i].A0 = 0; q [i].A1 = 0; q [i].A2 = (sigma * sigma + omega * omega);
	   q [i].B0 = 1; q [i].B1 = -2 * sigma; q [i].B2 = (sigma * sigma + omega * omega);
	}
	return pow (10.0, 0.05 * apass);
}

float warp_d_to_a (int fd, int fs) { return 2.0 * fs * tan ((2 * M_PI * fd) / (2 * fs)); }   // :120-122

float bilinear (Quad *q, int fs, int n) {
const float f2 = 2 * fs, f4 = f2 * f2;
float gain = 1.0;
	for (int i = 0; i < n; i ++) {
	   Quad &c = q [i];
	   const float N0 = c.A0 * f4 + c.A1 * f2 + c.A2;
	   const float N1 = 2 * (c.A2 - c.A0 * f4);
	   const float N2 = c.A0 * f4 - c.A1 * f2 + c.A2;
	   const float D0 = c.B0 * f4 + c.B1 * f2 + c.B2;
	   const float D1 = 2 * (c.B2 - c.B0 * f4);
	   const float D2 = c.B0 * f4 - c.B1 * f2 + c.B2;
	   c.A0 = 1.0; c.A1 = N1 / N0; c.A2 = N2 / N0;
	   c.B0 = 1.0; c.B1 = D1 / D0; c.B2 = D2 / D0;
	   gain *= (N0 / D0);
	}
	return gain;
}
}	// namespace

void design_squelch_iir (int32_t fm_rate, float *out) {
const int order = 20, n = kSquelchQuads, apass = -1;
const int key = 70000;                                   // keyFrequency, fm-processor.cpp:87
Quad q [kSquelchQuads];
int k = 0;
	{  // HighPassIIR (20, key - 100, fs, S_CHEBYSHEV)
	   int fpass = key - 100;
	   if (2 * fpass >= fm_rate) fpass = fm_rate / 4;
	   const float omega = warp_d_to_a (fpass, fm_rate);
	   float gain = chebyshev_even (q, n, order, apass);
	   for (int i = 0; i < n; i ++) {
	      const float A0 = q [i].A0, A1 = q [i].A1, A2 = q [i].A2, B0 = q [i].B0, B1 = q [i].B1, B2 = q [i].B2;
	      gain *= A2 / B2;
	      q [i].A0 = 1.0; q [i].B0 = 1.0;
	      q [i].A1 = (A1 / A2) * omega; q [i].B1 = (B1 / B2) * omega;
	      q [i].A2 = (A0 / A2) * omega * omega; q [i].B2 = (B0 / B2) * omega * omega;
	   }
	   gain *= bilinear (q, fm_rate, n);
	   out [k ++] = gain;
	   for (int i = 0; i < n; i ++) { out [k ++] = q [i].A1; out [k ++] = q [i].A2; out [k ++] = q [i].B1; out [k ++] = q [i].B2; }
	}
	{  // LowPassIIR (20, key, fs, S_CHEBYSHEV)
	   int fpass = key;
	   if (2 * fpass >= fm_rate) fpass = fm_rate / 4;
	   const float omega = warp_d_to_a (fpass, fm_rate);
	   float gain = chebyshev_even (q, n, order, apass);
	   for (int i = 0; i < n; i ++) {
	      q [i].A1 = q [i].A1 * omega; q [i].B1 = q [i].B1 * omega;
	      q [i].A2 = q [i].A2 * omega * omega; q [i].B2 = q [i].B2 * omega * omega;
	   }
	   gain *= bilinear (q, fm_rate, n);
	   out [k ++] = gain;
	   for (int i = 0; i < n; i ++) { out [k ++] = q [i].A1; out [k ++] = q [i].A2; out [k ++] = q [i].B1; out [k ++] = q [i].B2; }
	}
}

// ---- RDS symbol stage tables (rds-decoder-1.cpp:45-112) ------------------------------------------
namespace {
float butterworth_even (Quad *q, int n, int order, int apass) {          // newButterworth, iir-filters.cpp:124-171
const float Eps = sqrt (pow (10.0, -0.1 * apass) - 1);
const float R = 1.0 / pow (Eps, 1.0 / order);
	for (int i = 0; i < n; i ++) {
	   const float Phim = (M_PI * (2 * i + order + 1) / (2 * order));
	   const float sigma = R * cosf (Phim);
	   const float omega = R * sinf (Phim);
	   q [i].A0 = 0; q [i].A1 = 0; q [i].A2 = (sigma * sigma + omega * omega);
	   q [i].B0 = 1; q [i].B1 = -2 * sigma; q [i].B2 = (sigma * sigma + omega * omega);
	}
	return 1.0;
}
// roots of A s^2 + B s + C with complex<float> coefficients (cQuadratic, iir-filters.cpp:305-311)
void quad_roots (cf32 A, cf32 B, cf32 C, cf32 &D, cf32 &E) {
const cf32 AC = A * C;
const cf32 t = std::sqrt (B * B - cf32 (AC.real () * 4.0, AC.imag () * 4.0));
const cf32 A2 (A.real () * 2.0, A.imag () * 2.0);
	D = (-B + t) / A2;
	E = (-B - t) / A2;
}
}	// namespace

void design_rds_symbol_tables (int32_t rate, float *out) {
int k = 0;
//	matched filter, rds-decoder-1.cpp:56-99
	{
	   const double bitclk = 1187.5;
	   const float syncSamples = rate / (float)bitclk;
	   const int length = ((int)ceil (syncSamples) & ~01) + 1;            // 21
	   std::vector<float> kern (2 * length + 1, 0.f);
	   for (int i = 1; i <= length; i ++) {
	      const float x = ((float)i) / rate * bitclk;
	      const float v = 0.75 * cos (4 * M_PI * x) * ((1.0 / (1.0 / x - 64.01 * x)) - ((1.0 / (9.0 / x - 64.01 * x))));
	      const float w = - 0.75 * cos (4 * M_PI * x) * ((1.0 / (1.0 / x - 64.01 * x)) - ((1.0 / (9.0 / x - 64.01 * x))));
	      kern [length + i] = v;
	      kern [length - i] = w;
	   }
	   for (int i = 0; i < kRdsMatchTaps; i ++) out [k ++] = i < (int)kern.size () ? kern [i] : 0.f;
	}
//	rdsFilter = LowPassFIR (21, RDS_WIDTH = 4800, rate)
	{
	   std::vector<cf32> lp = design_lowpass (kRdsLpTaps, 2 * 2400, rate);
	   for (int i = 0; i < kRdsLpTaps; i ++) out [k ++] = lp [i].real ();
	}
//	sharpFilter = BandPassIIR (7, bitclk - 6, bitclk + 6, rate, S_BUTTERWORTH), iir-filters.cpp:555-596, 315-382
	{
	   const int order = (7 + 1) & 0176;                                  // 8
	   const int nq = order;                                              // Basic_IIR ((order + 1) & MAXORDER)
	   int flow = (int)(1187.5 - 6), fhigh = (int)(1187.5 + 6);
	   if (flow >= rate / 2) flow = (int)(0.2 * rate);
	   if (fhigh >= rate / 2) fhigh = (int)(0.3 * rate);
	   const float omegaL = warp_d_to_a (flow, rate), omegaH = warp_d_to_a (fhigh, rate);
	   const float Wo = sqrtf (omegaL * omegaH);
	   const float BW = omegaH - omegaL;
	   Quad temp [kRdsBpQuads], Q [kRdsBpQuads];
	   float gain = butterworth_even (temp, nq / 2, order, -1);
	   for (int i = 0; i < nq / 2; i ++) {                                // unnormalizeBP
	      cf32 A, B, C, D, E;
	      if (temp [i].A0 == 0.0) {
	         Q [2 * i].A0 = 0.0; Q [2 * i].A1 = sqrtf (temp [i].A2) * BW; Q [2 * i].A2 = 0.0;
	         Q [2 * i + 1].A0 = 0.0; Q [2 * i + 1].A1 = sqrtf (temp [i].A2) * BW; Q [2 * i + 1].A2 = 0.0;
	      }
	      else {
	         quad_roots (cf32 (temp [i].A0, 0.0), cf32 (temp [i].A1, 0.0), cf32 (temp [i].A2, 0.0), D, E);
	         quad_roots (cf32 (1.0, 0.0), cf32 (-D.real () * BW, -D.imag () * BW), cf32 (Wo * Wo, 0.0), D, E);
	         Q [2 * i].A0 = 1.0; Q [2 * i].A1 = -2.0 * D.real (); Q [2 * i].A2 = (D * std::conj (D)).real ();
	         Q [2 * i + 1].A0 = 1.0; Q [2 * i + 1].A1 = -2.0 * E.real (); Q [2 * i + 1].A2 = (E * std::conj (E)).real ();
	      }
	      quad_roots (cf32 (temp [i].B0, 0.0), cf32 (temp [i].B1, 0.0), cf32 (temp [i].B2, 0.0), D, E);
	      quad_roots (cf32 (1.0, 0.0), cf32 ((-D).real () * BW, (-D).imag () * BW), cf32 (Wo * Wo, 0), D, E);
	      Q [2 * i].B0 = 1.0; Q [2 * i].B1 = -2.0 * D.real (); Q [2 * i].B2 = (D * std::conj (D)).real ();
	      Q [2 * i + 1].B0 = 1.0; Q [2 * i + 1].B1 = -2.0 * E.real (); Q [2 * i + 1].B2 = (E * std::conj (E)).real ();
	   }
	   gain *= 1.0f;                                                       // unnormalizeBP returns 1.0
	   gain *= bilinear (Q, rate, nq);
	   out [k ++] = gain;
	   for (int i = 0; i < nq; i ++) { out [k ++] = Q [i].A1; out [k ++] = Q [i].A2; out [k ++] = Q [i].B1; out [k ++] = Q [i].B2; }
	}
}

namespace {
struct Packer {
	std::vector<float> f;
	int64_t put (const float *p, size_t n) {
	   while (f.size () % 4) f.push_back (0.0f);      // keep every table 16-byte aligned
	   int64_t off = (int64_t)f.size ();
	   f.insert (f.end (), p, p + n);
	   return off;
	}
	int64_t put (const std::vector<float> &v) { return put (v.data (), v.size ()); }
	int64_t put (const std::vector<cf32> &v) {
	   return put (reinterpret_cast<const float *>(v.data ()), v.size () * 2);
	}
};
}

void design_rds2_matched_filter (int32_t rate, float *out) {
//	root_raised_cosine (gain 1, sampling_freq rate, symbol_rate 2 * 1187.5, alpha 1, ntaps 45), every term in
//	double and in the reference's order (shaping_filter.cpp:9-54): with alpha = 1 the singular taps are -1
const int ntaps = 45 | 1;
const double gain = 1.0, alpha = 1.0;
const double spb = (double)rate / (2 * 1187.5);
double scale = 0;
std::vector<float> taps (ntaps);
	for (int i = 0; i < ntaps; i ++) {
	   double x1, x2, x3, num, den;
	   const double xindx = i - ntaps / 2;
	   x1 = M_PI * xindx / spb;
	   x2 = 4 * alpha * xindx / spb;
	   x3 = x2 * x2 - 1;
	   if (fabs (x3) >= 0.000001) {
	      if (i != ntaps / 2) num = cos ((1 + alpha) * x1) + sin ((1 - alpha) * x1) / (4 * alpha * xindx / spb);
	      else num = cos ((1 + alpha) * x1) + (1 - alpha) * M_PI / (4 * alpha);
	      den = x3 * M_PI;
	   }
	   else { taps [i] = -1; scale += taps [i]; continue; }        // alpha == 1
	   taps [i] = 4 * alpha * num / den;
	   scale += taps [i];
	}
	for (int i = 0; i < ntaps; i ++) out [i] = taps [i] * gain / scale;
}

// rdsDecoder_3's bit clock (src/rds/rds-decoder-3.cpp:62, 121-147): the sine table of SinCos (rate) (sincos.cpp:35-44),
// omegaRDS, and the slot of the correlation vector each of the 21 ring elements is added to — the index k of
// synchronizeOnBitClk restarts whenever the half-rate clock changes sign, which only depends on the element's number
void design_rds3_clock (int32_t rate, std::vector<float> &sin_tab, float *omega, int8_t *kmap /* [21] */) {
	sin_tab.resize (rate);
	for (int32_t i = 0; i < rate; i ++) sin_tab [i] = (float)sin (2 * M_PI * i / rate);
const double C = rate / (2 * M_PI);
const float w = (float)((2 * M_PI * 1187.5) / (float)rate);
	*omega = w;
const int ceiling = (int)ceil (rate / (float)1187.5);
bool isHigh = false;
int k = 0;
	for (int i = 0; i < ceiling && i < 21; i ++) {
	   const float phase = (float)fmod ((double)(i * (w / 2)), 2 * M_PI);
	   const float sn = phase < 0 ? -sin_tab [((int32_t)((double)(-phase) * C)) % rate] : sin_tab [((int32_t)((double)phase * C)) % rate];
	   if (sn > 0 && !isHigh) { isHigh = true; k = 0; }
	   else if (sn < 0 && isHigh) { isHigh = false; k = 0; }
	   kmap [i] = (int8_t)(k ++);
	}
}

TableBlob build_tables (int32_t input_rate, int32_t fm_rate, int32_t input_filter_hz,
                        int32_t audio_lp_hz) {
TableHeader h;
Packer pk;
	memset (&h, 0, sizeof (h));
	h.magic = 0x54464A53u; h.version = 1;
	h.input_rate = input_rate; h.fm_rate = fm_rate;
	h.input_filter_hz = input_filter_hz; h.audio_lp_hz = audio_lp_hz;

//	fmBand_1 / fmBand_2 constructor arguments, fm-processor.cpp:36,68-75
	const int32_t irate = input_rate / 6;
	h.decim1 = input_rate / irate;
	h.decim2 = irate / fm_rate;
	h.ntaps1 = 4 * input_rate / irate + 1;
	h.ntaps2 = irate / fm_rate + 1;
	std::vector<cf32> k1 = design_decimating_lowpass (h.ntaps1, fm_rate / 2, input_rate);
	std::vector<cf32> k2 = design_decimating_lowpass (h.ntaps2, fm_rate / 2, irate);
	std::vector<cf32> kr = design_decimating_lowpass (kRdsDecimTaps, 24000 / 2, fm_rate);
	h.off_fmband1 = pk.put (k1);
	h.off_fmband2 = pk.put (k2);
	h.off_rdsdecim = pk.put (kr);

//	Composite of the two decimators.  Each reference kernel is (t/sum, t) = t * (1/sum + j)
//	up to one float rounding of the real part, so the cascade equals a REAL FIR over the
//	input-rate samples followed by one constant complex gain G:
//	  z[m] = G * sum_n C[n] x[D*m + D - 1 - n],  C[d1*j + i] += t2[j]*t1[i],  D = d1*d2
	h.ncomp = h.ntaps1 + h.decim1 * (h.ntaps2 - 1);
	std::vector<double> cd (h.ncomp, 0.0);
	for (int j = 0; j < h.ntaps2; j ++)
	   for (int i = 0; i < h.ntaps1; i ++)
	      cd [h.decim1 * j + i] += (double)k2 [j].imag () * (double)k1 [i].imag ();
//	RF DC removal folded into the taps.  The reference subtracts the one-pole
//	estimate r[n] (fm-processor.cpp:425,444) BEFORE the FIR; inside the 37-sample window
//	r[n-i] = r[n] - alpha * sum_{j=n-i+1..n} (x[j] - r[j-1]), hence
//	  sum_i C[i] (x[n-i] - r[n-i]) = sum_i (C[i] + alpha g[i]) x[n-i] - r[n] (sumC + alpha sum g) + O(alpha^2)
//	with the tail sums g[i] = sum_{k>i} C[k].  The kernel therefore runs the taps
//	C'[i] = C[i] + alpha g[i]; while a component of r sits on the +-0.01 clamp the subtracted
//	value is constant instead, and the fm-rate stage takes alpha*sum g x back out using the
//	12-sample block sums (block means gbar[0..2] of g).
	const double alpha = (double)(1.0f / input_rate);      // rfDcAlpha, fm-processor.cpp:379
	std::vector<double> gt (h.ncomp, 0.0);
	for (int i = 0; i < h.ncomp; i ++)
	   for (int k = i + 1; k < h.ncomp; k ++) gt [i] += (double)(float)cd [k];
	std::vector<float> comp (h.ncomp);
	const int D = h.decim1 * h.decim2;
	double sumC = 0, sumCm = 0, gbar [3] = { 0, 0, 0 };
	for (int i = 0; i < h.ncomp; i ++) {
	   const double c = (double)(float)cd [i];
	   comp [i] = (float)(c + alpha * gt [i]);
	   sumC += c; sumCm += comp [i];
	   if (i / D < 3) gbar [i / D] += gt [i] / (double)D;      // means over the D-sample blocks S[m], S[m-1], S[m-2]
	}
	h.off_comp = pk.put (comp);
	std::complex<double> g1 ((double)k1 [h.ntaps1 / 2].real () / k1 [h.ntaps1 / 2].imag (), 1.0);
	std::complex<double> g2 ((double)k2 [h.ntaps2 / 2].real () / k2 [h.ntaps2 / 2].imag (), 1.0);
	std::complex<double> G = g1 * g2;
//	K_FM of fm_Demodulator::fm_Demodulator, fm-demodulator.cpp:58-64
	float F_G     = 0.65 * fm_rate / 2;
	float Delta_F = 0.95 * fm_rate / 2;
	float B_FM    = 2 * (Delta_F + F_G);
	float K_FM    = 2 * B_FM * M_PI / F_G;
	float consts [8] = { (float)sumC, (float)sumCm, (float)G.real (), (float)G.imag (),
	                     K_FM, (float)(alpha * gbar [0]), (float)(alpha * gbar [1]),
	                     (float)(alpha * gbar [2]) };
	h.off_comp_consts = pk.put (consts, 8);

//	compAtan, Xtan2.cpp:26-38 (Stretch is a float holding M_PI)
	{
	   const float Stretch = M_PI;
	   std::vector<float> t (kAtanTables * (kAtanSize + 1));
	   float *PPY = &t [0 * (kAtanSize + 1)], *PPX = &t [1 * (kAtanSize + 1)];
	   float *PNY = &t [2 * (kAtanSize + 1)], *PNX = &t [3 * (kAtanSize + 1)];
	   float *NPY = &t [4 * (kAtanSize + 1)], *NPX = &t [5 * (kAtanSize + 1)];
	   float *NNY = &t [6 * (kAtanSize + 1)], *NNX = &t [7 * (kAtanSize + 1)];
	   for (int i = 0; i <= kAtanSize; i ++) {
	      float f = (float)i / kAtanSize;
	      PPY [i] = atanf (f) * Stretch / M_PI;     // atan (float) is the float overload there
	      PPX [i] = Stretch * 0.5f - PPY [i];
	      PNY [i] = -PPY [i];
	      PNX [i] = PPY [i] - Stretch * 0.5f;
	      NPY [i] = Stretch - PPY [i];
	      NPX [i] = PPY [i] + Stretch * 0.5f;
	      NNY [i] = PPY [i] - Stretch;
	      NNX [i] = -Stretch * 0.5f - PPY [i];
	   }
	   h.off_atan = pk.put (t);
	}

//	SinCos (fmRate), sincos.cpp:36-45
	{
	   std::vector<cf32> t (fm_rate);
	   for (int i = 0; i < fm_rate; i ++)
	      t [i] = cf32 (cos (2 * M_PI * i / fm_rate), sin (2 * M_PI * i / fm_rate));
	   h.off_sincos = pk.put (t);
	}

//	Arcsine table, fm-demodulator.cpp:73-77
	{
	   std::vector<float> t (kArcsineSize + 1);
	   for (int i = 0; i <= kArcsineSize; i ++)
	      t [i] = asin (2.0 * i / kArcsineSize - 1.0) / 2.0;
	   h.off_arcsine = pk.put (t);
	}

	h.off_tw2048  = pk.put (fft_twiddles (kPssFftSize));
	h.off_tw8192  = pk.put (fft_twiddles (kAudioFftSize));
	h.off_tw32768 = pk.put (fft_twiddles (kRdsFftSize));

//	PerfectStereoSeparation: lpFilter (2048, 295).setLowPass (15000, rate),
//	stereo-separation.cpp:31-39
	h.off_pss_lp = pk.put (filter_spectrum (design_lowpass (kPssDegree, 15000, fm_rate), kPssFftSize));
//	rdsBandPassFilter.setBand (57000 -+ 2400, fmRate), fm-processor.cpp:166-168
	h.off_rds_bp = pk.put (filter_spectrum (
	         design_bandpass (kRdsDegree, 57000 - 2400, 57000 + 2400, fm_rate), kRdsFftSize));
//	fmAudioFilter.setLowPass (lowPassFrequency, fmRate), fm-processor.cpp:403-408
	if (audio_lp_hz > 0)
	   h.off_audio_lp = pk.put (filter_spectrum (
	         design_lowpass (kAudioDegree, audio_lp_hz, fm_rate), kAudioFftSize));
//	inputFilter.setLowPass (fmBandwidth / 2, inputRate), fm-processor.cpp:397-401.  The
//	65536-point overlap-add filter is a delayed linear convolution with these 251 real
//	taps; the GPU path folds them into the decimator composite.
	if (input_filter_hz > 0) {
	   std::vector<cf32> lp = design_lowpass (kInputDegree, input_filter_hz / 2, input_rate);
	   std::vector<float> t (kInputDegree);
	   for (int i = 0; i < kInputDegree; i ++) t [i] = lp [i].real ();
	   h.off_input_taps = pk.put (t);
	   h.ncomp_wide = h.ncomp + kInputDegree - 1;
	   std::vector<double> wd (h.ncomp_wide, 0.0);
	   for (int a = 0; a < kInputDegree; a ++)
	      for (int b = 0; b < h.ncomp; b ++)
	         wd [a + b] += (double)t [a] * cd [b];
	   std::vector<float> wide (h.ncomp_wide);
	   for (int i = 0; i < h.ncomp_wide; i ++) wide [i] = (float)wd [i];
	   h.off_comp_wide = pk.put (wide);
	}
//	rational polyphase resampler (new block, resample.cuh)
	{
	   int L, M, P; std::vector<float> hA, hB;
	   if (design_resampler (input_rate, fm_rate, L, M, P, hA, hB)) {
	      h.rs_L = L; h.rs_M = M; h.rs_P = P; h.rs_ntapsA = (int32_t)hA.size ();
	      h.off_rsA = pk.put (hA);
	      h.off_rsB = pk.put (hB);
	   }
	}
	{
	   float sq [kSquelchFloats];
	   design_squelch_iir (fm_rate, sq);
	   h.off_squelch = pk.put (sq, kSquelchFloats);
	   float rs [kRdsSymFloats];
	   design_rds_symbol_tables (24000, rs);                              // RDS_RATE, fm-constants.h
	   h.off_rds_sym = pk.put (rs, kRdsSymFloats);
	}
	while (pk.f.size () % 4) pk.f.push_back (0.0f);
	h.payload_floats = (int64_t)pk.f.size ();

TableBlob blob;
	blob.bytes.resize (sizeof (TableHeader) + pk.f.size () * sizeof (float));
	memcpy (blob.bytes.data (), &h, sizeof (h));
	memcpy (blob.bytes.data () + sizeof (h), pk.f.data (), pk.f.size () * sizeof (float));
	return blob;
}

}	// namespace sdrjfm
