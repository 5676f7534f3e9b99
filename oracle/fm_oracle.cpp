/*
 * TEST INFRASTRUCTURE — not product code.  Nothing under sdr-j-fm_b200/ links or calls this.
 *
 * CPU restatement ("port") of the reference's FM hot path in plain, flat C++: one struct,
 * one sample loop, no reference headers.  It follows, block by block and with the same
 * float/double promotions:
 *   RF DC removal, IQ gain, LO                 src/fm/fm-processor.cpp:423-446, 462-466
 *   Oscillator                                 src/various/oscillator.cpp:26-58
 *   fftFilter (overlap-add, complex / real)    src/various/fft-filters.cpp:29-163
 *   fftFilterHilbert                           src/various/fft-filters.cpp:166-201
 *   radix-2 FFT                                src/various/fft-complex.cpp:50-102
 *   tap designers                              src/various/fir-filters.cpp:41-62,197-222,327-347
 *   DecimatingFIR::Pass                        src/various/fir-filters.cpp:397-424
 *   fm_Demodulator (decoders 1..6)             src/fm/fm-demodulator.cpp:51-241
 *   pllC                                       src/various/pllC.cpp:38-90
 *   compAtan                                   src/various/Xtan2.cpp:12-100
 *   SinCos                                     src/various/sincos.cpp:36-91
 *   pilotRecovery                              src/fm/pilot-recover.cpp:28-83
 *   PerfectStereoSeparation                    src/fm/stereo-separation.cpp:27-110
 *   process_signal_with_rds, matrix, de-emph   src/fm/fm-processor.cpp:689-759, 517-549, 594-595, 303-306
 *
 * PINNING: the reference has no golden vectors or tests.  This port is
 * pinned by tests/test_oracle_vs_reference.py, which requires BIT-EXACT agreement of every
 * tap with oracle/_ref (the reference's own classes compiled from /root/reference) and by
 * the committed fixtures under tests/golden/ that were generated from oracle/_ref.
 * The 192 kHz -> 48 kHz step (libsamplerate) is outside both: parity unpinned there.
 */
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "chain_api.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef M_PI_4
#define M_PI_4 0.78539816339744830962
#endif

namespace {

typedef std::complex<float> cf;

// ---------------------------------------------------------------- FFT (fft-complex.cpp:50-102)
static void fft_r2 (cf *v, int n) {
int levels = 0;
	for (int t = n; t > 1; t >>= 1) levels ++;
std::vector<cf> w (n / 2);
	for (int i = 0; i < n / 2; i ++) {
	   const float ang = (float)(-2 * M_PI * i / n);          // narrowed by the complex<float> ctor
	   w [i] = std::exp (cf (0, ang));
	}
	for (int i = 0; i < n; i ++) {
	   int j = 0;
	   for (int b = 0, t = i; b < levels; b ++, t >>= 1) j = (j << 1) | (t & 1);
	   if (j > i) std::swap (v [i], v [j]);
	}
	for (int size = 2; size <= n; size *= 2) {
	   const int half = size / 2, step = n / size;
	   for (int i = 0; i < n; i += size)
	      for (int j = i, k = 0; j < i + half; j ++, k += step) {
	         const cf t = v [j + half] * w [k];
	         v [j + half] = v [j] - t;
	         v [j] += t;
	      }
	   if (size == n) break;
	}
}

// ---------------------------------------------------------------- tap designers
static std::vector<float> sinc_blackman (int n, float f) {     // common body of the three newKernel()s
std::vector<float> t (n);
	for (int i = 0; i < n; i ++) {
	   if (i == n / 2) t [i] = 2 * M_PI * f;
	   else t [i] = sin (2 * M_PI * f * (i - n / 2)) / (i - n / 2);
	   t [i] *= (0.42 - 0.50 * cos (2 * M_PI * (float)i / (float)n)
	                  + 0.08 * cos (4 * M_PI * (float)i / (float)n));
	}
	return t;
}

static std::vector<cf> lowpass_taps (int n, int32_t fc, int32_t fs) {      // fir-filters.cpp:41-62
std::vector<float> t = sinc_blackman (n, (float)fc / fs);
float sum = 0;
	for (float v : t) sum += v;
std::vector<cf> k (n);
	for (int i = 0; i < n; i ++) k [i] = cf (t [i] / sum, 0);
	return k;
}

static std::vector<cf> bandpass_taps (int n, int32_t low, int32_t high, int32_t fs) {   // :197-222
const float lo = (float)((high - low) / 2) / fs;
const float shift = (float)((high + low) / 2) / fs;
std::vector<float> t = sinc_blackman (n, lo);
float sum = 0;
	for (float v : t) sum += v;
std::vector<cf> k (n);
	for (int i = 0; i < n; i ++) {
	   float v = (i - n / 2) * (2 * M_PI * shift);
	   k [i] = cf (t [i] * cosf (v) / sum, t [i] * sinf (v) / sum);
	}
	return k;
}

// ---------------------------------------------------------------- DecimatingFIR
struct DecimFir {
	int n, dm, cnt, ip;
	std::vector<cf> k, buf;
	DecimFir (int n_, int32_t low, int32_t fs, int dm_) : n (n_), dm (dm_), cnt (0), ip (0), k (n_), buf (n_, cf (0, 0)) {
	   std::vector<float> t = sinc_blackman (n, (float)low / fs);          // :327-347
	   float sum = 0;
	   for (float v : t) sum += v;
	   for (int i = 0; i < n; i ++) k [i] = cf (t [i] / sum, t [i]);
	}
	bool pass (cf z, cf *out) {                                             // :397-424
	   buf [ip] = z;
	   if (++cnt < dm) { ip = (ip + 1) % n; return false; }
	   cnt = 0;
	   cf tmp = 0;
	   for (int i = 0; i <= ip; i ++) tmp += buf [ip - i] * k [i];
	   for (int i = ip + 1; i < n; i ++) tmp += buf [n + ip - i] * k [i];
	   ip = (ip + 1) % n;
	   *out = tmp;
	   return true;
	}
};

// ---------------------------------------------------------------- fftFilter
struct FftFilter {
	int nfft, degree, nsamp, inp;
	std::vector<cf> A, Cc, H, over;
	FftFilter (int size, int deg) : nfft (size), degree (deg), nsamp (size - deg), inp (0),
	      A (size, cf (0, 0)), Cc (size, cf (0, 0)), H (size, cf (0, 0)), over (deg, cf (0, 0)) {}
	void load (const std::vector<cf> &taps) {                               // :71-95
	   for (int i = 0; i < degree; i ++) H [i] = taps [i];
	   for (int i = degree; i < nfft; i ++) H [i] = cf (0, 0);
	   fft_r2 (H.data (), nfft);
	   inp = 0;
	}
	void set_hilbert () {                                                   // :177-201 (even size)
	   H [0] = 1.0f;
	   for (int i = 1; i < nfft / 2; i ++) H [i] = 2.0f;
	   H [nfft / 2] = 1.0f;
	   for (int i = nfft / 2 + 1; i < nfft; i ++) H [i] = 0.0f;
	   inp = 0;
	}
	void block (bool times3) {
	   for (int i = nsamp; i < nfft; i ++) A [i] = cf (0, 0);
	   fft_r2 (A.data (), nfft);
	   for (int j = 0; j < nfft; j ++) {
	      Cc [j] = A [j] * H [j];
	      if (times3) Cc [j] = cf (real (Cc [j]) * 3, imag (Cc [j]) * 3);
	   }
	   for (int j = 0; j < nfft; j ++) Cc [j] = conj (Cc [j]);
	   fft_r2 (Cc.data (), nfft);
	   const float factor = times3 ? (float)(1.0f / nfft) : (float)(1.0 / nfft);
	   for (int j = 0; j < nfft; j ++) Cc [j] = conj (Cc [j]) * factor;
	   for (int j = 0; j < degree; j ++) { Cc [j] += over [j]; over [j] = Cc [nsamp + j]; }
	}
	float pass_real (float x) {                                             // :97-130
	   float s = real (Cc [inp]);
	   A [inp] = x;
	   if (++inp >= nsamp) { inp = 0; block (true); }
	   return s;
	}
	cf pass_cplx (cf z) {                                                   // :132-163
	   cf s = Cc [inp];
	   A [inp] = cf (real (z), imag (z));
	   if (++inp >= nsamp) { inp = 0; block (false); }
	   return s;
	}
};

// ---------------------------------------------------------------- SinCos
struct SinCosT {
	int32_t rate; double C; std::vector<cf> tab;
	SinCosT (int32_t r) : rate (r), C (r / (2 * M_PI)), tab (r) {            // :36-45
	   for (int i = 0; i < r; i ++) tab [i] = cf (cos (2 * M_PI * i / r), sin (2 * M_PI * i / r));
	}
	int32_t index (float ph) const {                                        // :54-58
	   if (ph >= 0) return (int32_t (ph * C)) % rate;
	   return rate - (int32_t (ph * C)) % rate;
	}
	float get_sin (float ph) const {                                        // :75-79
	   if (ph < 0) return -get_sin (-ph);
	   return imag (tab [index (ph)]);
	}
	cf get_complex (float ph) const {                                       // :87-91 (getCos :81-85 takes its real part)
	   while (ph < 0) ph += 2 * M_PI;
	   ph = fmod (ph, 2 * M_PI);
	   return tab [(int32_t (ph * C)) % rate];
	}
};

// ---------------------------------------------------------------- compAtan
struct AtanLut {
	enum { SIZE = 8192 };
	std::vector<float> t [8];   // PPY PPX PNY PNX NPY NPX NNY NNX
	AtanLut () {                                                            // Xtan2.cpp:12-39
	   const float Stretch = M_PI;
	   for (auto &v : t) v.resize (SIZE + 1);
	   for (int i = 0; i <= SIZE; i ++) {
	      float f = (float)i / SIZE;
	      t [0][i] = atanf (f) * Stretch / M_PI;
	      t [1][i] = Stretch * 0.5f - t [0][i];
	      t [2][i] = -t [0][i];
	      t [3][i] = t [0][i] - Stretch * 0.5f;
	      t [4][i] = Stretch - t [0][i];
	      t [5][i] = t [0][i] + Stretch * 0.5f;
	      t [6][i] = t [0][i] - Stretch;
	      t [7][i] = -Stretch * 0.5f - t [0][i];
	   }
	}
	float atan2 (float y, float x) const {                                  // :56-100
	   const int EZIS = -SIZE;
	   if (std::isinf (x) || std::isinf (y)) return 0;
	   if (std::isnan (x) || std::isnan (y)) return 0;
	   if (x == 0) {
	      if (y == 0) return 0;
	      else if (y > 0) return M_PI / 2;
	      else return -M_PI / 2;
	   }
	   if (x > 0) {
	      if (y >= 0) {
	         if (x >= y) return t [0][(int)(SIZE * y / x + 0.5)];
	         else return t [1][(int)(SIZE * x / y + 0.5)];
	      }
	      else {
	         if (x >= -y) return t [2][(int)(EZIS * y / x + 0.5)];
	         else return t [3][(int)(EZIS * x / y + 0.5)];
	      }
	   }
	   else {
	      if (y >= 0) {
	         if (-x >= y) return t [4][(int)(EZIS * y / x + 0.5)];
	         else return t [5][(int)(EZIS * x / y + 0.5)];
	      }
	      else {
	         if (x <= y) return t [6][(int)(SIZE * y / x + 0.5)];
	         else return t [7][(int)(SIZE * x / y + 0.5)];
	      }
	   }
	}
};

static float pi_constrain (float val) {                                    // fm-constants.h:148-158
	if (0 <= val && val < 2 * M_PI) return val;
	if (val >= 2 * M_PI) return fmod (val, 2 * M_PI);
	if (val > -2 * M_PI) return val + 2 * M_PI;
	return 2 * M_PI - fmod (-val, 2 * M_PI);
}

// ---------------------------------------------------------------- the chain
struct Oracle {
	chain_cfg cfg;
	int32_t inRate, fmRate;
	std::vector<cf> loTab; int32_t loPhase;
	SinCosT sc;
	AtanLut at;
	DecimFir band1, band2, rdsDecim;
	FftFilter audioLp, inputLp, pssLp, rdsBp, rdsHil;
	bool inputOn, audioOn;
	// fm_Demodulator
	float K_FM, Imin1, Qmin1, Imin2, Qmin2, fm_afc, am_carr;
	std::vector<float> arcs;
	// pllC
	float pllBeta, pllNco, pllIncr, pllLo, pllHi, pllErr;
	// pilotRecovery
	float pOmega, pGain, pPhase, pOld, pLock, pQuad; bool pLocked; int32_t pStable;
	// PSS
	float psAlpha, psLockAlpha, psAcc, psMean; bool psMin; int32_t psLockCnt, psUnlockCnt;
	float pilotDelayPSS;
	std::vector<float> rdsPhaseBuf; int rdsPhaseIdx;
	cf lastAudio, RfDC; float rfAlpha, deAlpha, vol, pan, lch, rch;

	Oracle (const chain_cfg &c) : cfg (c), inRate (c.input_rate), fmRate (c.fm_rate),
	   loTab (c.input_rate), loPhase (0), sc (c.fm_rate),
	   band1 (4 * c.input_rate / (c.input_rate / 6) + 1, c.fm_rate / 2, c.input_rate,
	          c.input_rate / (c.input_rate / 6)),
	   band2 ((c.input_rate / 6) / c.fm_rate + 1, c.fm_rate / 2, c.input_rate / 6,
	          (c.input_rate / 6) / c.fm_rate),
	   rdsDecim (11, 24000 / 2, c.fm_rate, c.fm_rate / 24000),
	   audioLp (8192, 756), inputLp (65536, 251), pssLp (2048, 295),
	   rdsBp (32768, 768), rdsHil (32768, 768),
	   rdsPhaseBuf (2 * (32768 - 768), 0.0f), rdsPhaseIdx (0) {
	   for (int i = 0; i < inRate; i ++)                                    // oscillator.cpp:30-32
	      loTab [i] = cf (cos (2.0 * M_PI * i / inRate), sin (2.0 * M_PI * i / inRate));
	   inputLp. load (lowpass_taps (251, 0.95 * fmRate / 2, inRate));       // fm-processor.cpp:148
	   inputOn = c.input_filter_hz > 0;
	   if (inputOn) inputLp. load (lowpass_taps (251, c.input_filter_hz / 2, inRate));
	   audioOn = c.lf_cutoff_hz > 0;
	   if (audioOn) audioLp. load (lowpass_taps (756, c.lf_cutoff_hz, fmRate));
	   pssLp. load (lowpass_taps (295, 15000, fmRate));                     // stereo-separation.cpp:39
	   rdsBp. load (bandpass_taps (768, 57000 - 4800 / 2, 57000 + 4800 / 2, fmRate));
	   rdsHil. set_hilbert ();
	   // fm_Demodulator ctor, fm-demodulator.cpp:51-87
	   float F_G = 0.65 * fmRate / 2, Delta_F = 0.95 * fmRate / 2, B_FM = 2 * (Delta_F + F_G);
	   K_FM = 2 * B_FM * M_PI / F_G;
	   arcs. resize (4 * 8192 + 1);
	   for (int i = 0; i <= 4 * 8192; i ++) arcs [i] = asin (2.0 * i / (4 * 8192) - 1.0) / 2.0;
	   Imin1 = Qmin1 = Imin2 = Qmin2 = 0.01; fm_afc = 0; am_carr = 0;
	   {  // pllC ctor, pllC.cpp:38-58
	      float maxdev = 0.95 * (0.5 * fmRate);
	      float fac = 2.0 * M_PI / fmRate;
	      float bandwidth = 0.85 * fmRate;
	      pllBeta = exp (-2.0 * M_PI * bandwidth / 2 / fmRate);
	      pllNco = 0; pllErr = 0; pllIncr = 0 * fac;
	      pllLo = -maxdev * fac; pllHi = maxdev * fac;
	   }
	   pOmega = ((float (19000)) / fmRate) * (2 * M_PI);                    // fm-processor.cpp:34,78-80
	   pGain = 10 * (2 * M_PI) / fmRate;
	   pPhase = 0; pOld = 0; pLock = 0; pQuad = 0; pLocked = false; pStable = 0;
	   psAlpha = 10.0f / fmRate; psLockAlpha = 1.0f / fmRate;               // :81-82, stereo-separation.cpp:32
	   pss_reset ();
	   pilotDelayPSS = 0;
	   lastAudio = 0; RfDC = cf (0, 0); rfAlpha = 1.0f / inRate;
	   { float Tau = 1000000.0 / c.deemph_us; deAlpha = 1.0 / (float (fmRate) / Tau + 1.0); }
	   vol = std::pow (10.0f, c.volume_db / 20.0f);
	   pan = (float)c.panorama / 100.0f;
	   lch = (c.balance > 0 ? (100 - c.balance) / 100.0 : 1.0f);
	   rch = (c.balance < 0 ? (100 + c.balance) / 100.0 : 1.0f);
	}

	void pss_reset () { psAcc = 0; psMin = false; psMean = 0; psLockCnt = 0; psUnlockCnt = 0; }

	float pss_sample (float mux, float mixPhase) {                          // stereo-separation.cpp:60-110
	   cf p = sc. get_complex (mixPhase) * mux;
	   p = pssLp. pass_cplx (p);
	   float error = real (p) * imag (p);
	   if (!psMin) error *= 10.0f;
	   psAcc += psAlpha * error;
	   psMean = psLockAlpha * error + psMean * (1.0f - psLockAlpha);
	   const bool ok = (std::abs (psMean) < 0.001f);
	   if (ok) {
	      if (psMin || (++psLockCnt > 3 * fmRate)) psMin = true;
	      psUnlockCnt = 0;
	   }
	   else {
	      if (!psMin || (++psUnlockCnt > 3 * fmRate)) psMin = false;
	      psLockCnt = 0;
	   }
	   if (psAcc < -M_PI_4) psAcc = -M_PI_4;
	   else if (psAcc > M_PI_4) psAcc = M_PI_4;
	   return psAcc;
	}

	float pilot_phase (float pilot) {                                       // pilot-recover.cpp:54-83
	   float osc = sc. get_sin (pPhase);
	   float perr = pilot * osc;
	   constexpr float alpha = 1.0f / 3000.0f;
	   pPhase += perr * pGain;
	   const float cur = pi_constrain (pPhase);
	   pPhase = pi_constrain (pPhase + pOmega);
	   pQuad = (osc - pOld) / pOmega;
	   pOld = osc;
	   pLock = alpha * (-pQuad * pilot) + pLock * (1.0 - alpha);
	   if (pLock > 0.07f) {
	      if (pLocked || ++pStable > (fmRate >> 1)) pLocked = true;
	   }
	   else { pLocked = false; pStable = 0; }
	   return cur;
	}

	void do_pll (cf s) {                                                    // pllC.cpp:67-90
	   cf nco = sc. get_complex (pllNco);
	   cf d = conj (nco) * s;
	   pllErr = at. atan2 (imag (d), real (d));
	   pllIncr = (1 - pllBeta) * pllErr + pllBeta * pllIncr;
	   if (pllIncr < pllLo || pllIncr > pllHi) pllIncr = 0 * 2 * M_PI / fmRate;
	   pllNco += pllIncr;
	   if (pllNco >= 2 * M_PI) pllNco = fmod (pllNco, 2 * M_PI);
	   else while (pllNco < 0) pllNco += 2 * M_PI;
	}

	float demodulate (cf z) {                                               // fm-demodulator.cpp:111-205
	   float res, I, Q;
	   float carrierAlpha = 0.0010f, fmDcAlpha = 0.0001f;
	   float zAbs = std::abs (z);
	   if (zAbs <= 0.001) I = Q = 0.001;
	   else { I = real (z) / zAbs; Q = imag (z) / zAbs; }
	   am_carr = (1.0f - carrierAlpha) * am_carr + carrierAlpha * zAbs;
	   if (cfg.decoder == 1) {                                              // decodeAM, :215-241
	      do_pll (z);                                                       // on the RAW sample: AFC read-out only
	      res = pllIncr;
	      fm_afc = (1 - fmDcAlpha) * fm_afc + fmDcAlpha * res;
	      float gainLimit = 0.01f;
	      res = (std::abs (z) - am_carr) / (am_carr < gainLimit ? gainLimit : am_carr);
	      float audioLimit = 1.0f;
	      if (res > audioLimit) res = audioLimit;
	      else if (res < -audioLimit) res = -audioLimit;
	      return res;
	   }
	   z = cf (I, Q);
	   int index = 0;
	   float Scaler = sqrt (2);
	   switch (cfg.decoder) {
	      default:
	      case 2: do_pll (z); res = pllIncr; break;
	      case 3: res = at. atan2 (Q * Imin1 - I * Qmin1, I * Imin1 + Q * Qmin1); break;
	      case 4: { cf m = z * cf (Imin1, -Qmin1); res = at. atan2 (imag (m), real (m)); break; }
	      case 5:
	         res = (Imin1 * Q - Qmin1 * I + 1) / 2.0;
	         index = (int)floor (res * (4 * 8192));
	         if (index < 0) index = 0;
	         if (index >= 4 * 8192) index = 4 * 8192;
	         res = arcs [index];
	         break;
	      case 6:
	         res = (Imin1 * (Q - Qmin2) - Qmin1 * (I - Imin2));
	         res /= (Imin1 * Imin1 + Qmin1 * Qmin1) * Scaler;
	         Imin2 = Imin1; Qmin2 = Qmin1;
	         break;
	   }
	   fm_afc = (1 - fmDcAlpha) * fm_afc + fmDcAlpha * res;
	   res = 20.0f * (res - fm_afc) * 1.0f / K_FM;
	   Imin1 = I; Qmin1 = Q;
	   return res;
	}

//	one fm-rate sample after the discriminator (:515-648); also entered by orc_process_demod
	void after_demod (float demod, cf v, const chain_taps *t, int64_t &nfm, int64_t &nrds) {
	   // process_signal_with_rds, :689-759
	   float curPhase = pilot_phase (5 * demod);
	   const bool locked = pLocked;
	   if (!locked) { pilotDelayPSS = 0; pss_reset (); }
	   cf audio, rdsC (0, 0);
	   if (cfg.fm_mode != 2 && (locked || !cfg.auto_mono)) {
	      float ph = 2 * (curPhase + M_PI_4 + 0) - pilotDelayPSS;
	      if (ph < -2 * M_PI) ph += 4 * M_PI;
	      ph = fmod (ph, 2 * M_PI);
	      pilotDelayPSS = cfg.pss_on ? pss_sample (demod, ph) : 0;
	      float diff = 2.0 * (cfg.sound_sel == 6 ? sc. get_sin (ph) : real (sc. get_complex (ph))) * demod;
	      audio = cf (demod, diff);
	   }
	   else audio = cf (demod, 0);
	   if (cfg.rds_on) {
	      float bp = rdsBp. pass_real (demod);
	      cf hil = rdsHil. pass_cplx (cf (bp, 0));
	      float thePhase = 3 * (rdsPhaseBuf [rdsPhaseIdx] + 0);
	      rdsPhaseBuf [rdsPhaseIdx] = curPhase;
	      rdsPhaseIdx = (rdsPhaseIdx + 1) % (int)rdsPhaseBuf. size ();
	      cf osc (cosf (thePhase), -sinf (thePhase));
	      rdsC = osc * hil;
	   }
	   // matrix and selector, :517-549
	   const float sumLR = real (audio), diffLR = imag (audio);
	   const float dw = diffLR * (cfg.fm_mode == 1 ? pan : 1.0f);
	   const float left = sumLR + dw, right = sumLR - dw;
	   switch (cfg.sound_sel) {
	      default:
	      case 0: audio = cf (left, right); break;
	      case 1: audio = cf (right, left); break;
	      case 2: audio = cf (left, left); break;
	      case 3: audio = cf (right, right); break;
	      case 4: audio = cf (sumLR, sumLR); break;
	      case 5: case 6: audio = cf (dw, dw); break;
	   }
	   if (t -> fm_z) { t -> fm_z [2 * nfm] = real (v); t -> fm_z [2 * nfm + 1] = imag (v); }
	   if (t -> demod) t -> demod [nfm] = demod;
	   if (t -> pilot_phase) t -> pilot_phase [nfm] = curPhase;
	   if (t -> locked) t -> locked [nfm] = locked ? 1 : 0;
	   if (t -> pss_delay) t -> pss_delay [nfm] = pilotDelayPSS;
	   if (t -> lr) { t -> lr [2 * nfm] = real (audio); t -> lr [2 * nfm + 1] = imag (audio); }
	   if (t -> rds_cplx) { t -> rds_cplx [2 * nfm] = real (rdsC); t -> rds_cplx [2 * nfm + 1] = imag (rdsC); }
	   if (cfg.rds_on) {
	      cf r24;
	      if (rdsDecim. pass (rdsC, &r24)) {                             // :553
	         if (t -> rds24) { t -> rds24 [2 * nrds] = real (r24); t -> rds24 [2 * nrds + 1] = imag (r24); }
	         nrds ++;
	      }
	   }
	   if (audioOn) audio = audioLp. pass_cplx (audio);                  // :589-591
	   audio = lastAudio = (audio - lastAudio) * deAlpha + lastAudio;    // :594-595
	   const float gl = vol * lch * real (audio), gr = vol * rch * imag (audio);   // :304-305
	   if (t -> audio192) { t -> audio192 [2 * nfm] = gl; t -> audio192 [2 * nfm + 1] = gr; }
	}

	int64_t process (const float *iq, int64_t n_in, const chain_taps *t, int64_t *n_rds24) {
	   int64_t nfm = 0, nrds = 0;
	   for (int64_t i = 0; i < n_in; i ++) {
	      cf x (iq [2 * i], iq [2 * i + 1]);
	      if (cfg.dc_remove) {                                              // fm-processor.cpp:423-446
	         RfDC = (x - RfDC) * rfAlpha + RfDC;
	         float dr = real (RfDC), di = imag (RfDC);
	         if (dr > 0.01f) dr = 0.01f; else if (dr < -0.01f) dr = -0.01f;
	         if (di > 0.01f) di = 0.01f; else if (di < -0.01f) di = -0.01f;
	         x -= cf (dr, di);
	      }
	      cf v (real (x) * cfg.lgain, imag (x) * cfg.rgain);                // :462-464
	      loPhase -= cfg.lo_hz;                                             // oscillator.cpp:49-58
	      if (loPhase < 0) loPhase += inRate; else if (loPhase >= inRate) loPhase -= inRate;
	      v = v * loTab [loPhase];
	      if (inputOn) v = inputLp. pass_cplx (v);                          // :469-470
	      if (inRate / fmRate > 1) {                                        // :471 (192000: bypass)
	         if (!band1. pass (v, &v)) continue;                            // :472-475
	         if (!band2. pass (v, &v)) continue;
	      }
	      float demod = demodulate (v);                                     // :497
	      after_demod (demod, v, t, nfm, nrds);
	      nfm ++;
	   }
	   if (n_rds24) *n_rds24 = nrds;
	   return nfm;
	}
};
}	// namespace

extern "C" {
void	*orc_create (const chain_cfg *cfg) { return new Oracle (*cfg); }
void	orc_destroy (void *h) { delete (Oracle *)h; }
int64_t	orc_process (void *h, const float *iq, int64_t n_in, const chain_taps *taps, int64_t *n_rds24) {
	return ((Oracle *)h) -> process (iq, n_in, taps, n_rds24);
}
int64_t	orc_process_demod (void *h, const float *demod, int64_t n_fm, const chain_taps *taps,
	                   int64_t *n_rds24) {
Oracle *c = (Oracle *)h;
int64_t nfm = 0, nrds = 0;
	for (int64_t i = 0; i < n_fm; i ++) { c -> after_demod (demod [i], cf (0, 0), taps, nfm, nrds); nfm ++; }
	if (n_rds24) *n_rds24 = nrds;
	return nfm;
}

void	orc_get_meta (void *h, chain_meta *m) {
Oracle *c = (Oracle *)h;
	m -> dc_rf_re = real (c -> RfDC); m -> dc_rf_im = imag (c -> RfDC);
	m -> dc_if = c -> fm_afc; m -> carrier_ampl = c -> am_carr;
	m -> pss_phase_shift = c -> pilotDelayPSS; m -> pss_mean_error = c -> psMean;
	m -> pss_minimized = c -> psMin; m -> pilot_lock_strength = c -> pLock;
	m -> pilot_locked = c -> pLocked;
	m -> squelch_active = 0;          // the port has no squelch (ref_ only; chain_api.h)
}
int32_t	orc_dump_taps (void *h, int which, float *out, int32_t cap) {
Oracle *c = (Oracle *)h;
const cf *src = nullptr; int32_t n = 0;
	switch (which) {
	   case DUMP_FMBAND1: src = c -> band1.k.data (); n = c -> band1.n; break;
	   case DUMP_FMBAND2: src = c -> band2.k.data (); n = c -> band2.n; break;
	   case DUMP_RDSDECIM: src = c -> rdsDecim.k.data (); n = c -> rdsDecim.n; break;
	   case DUMP_INPUT_FILTER_FREQ: src = c -> inputLp.H.data (); n = c -> inputLp.nfft; break;
	   case DUMP_RDS_BP_FREQ: src = c -> rdsBp.H.data (); n = c -> rdsBp.nfft; break;
	   case DUMP_PSS_LP_FREQ: src = c -> pssLp.H.data (); n = c -> pssLp.nfft; break;
	   case DUMP_AUDIO_LP_FREQ: src = c -> audioLp.H.data (); n = c -> audioLp.nfft; break;
	   case DUMP_SINCOS: src = c -> sc.tab.data (); n = c -> sc.rate; break;
	   case DUMP_ATAN:
	      if (cap < 8 * 8193 / 2) return -1;
	      for (int t = 0; t < 8; t ++) memcpy (out + t * 8193, c -> at.t [t].data (), 8193 * sizeof (float));
	      return 8 * 8193 / 2;
	   case DUMP_CONSTS:
	      if (cap < 4) return -1;
	      out [0] = c -> K_FM; out [1] = c -> deAlpha; out [2] = c -> vol; out [3] = c -> pOmega;
	      out [4] = c -> pGain; out [5] = c -> psAlpha; out [6] = c -> psLockAlpha; out [7] = c -> rfAlpha;
	      return 4;
	   default: return -1;
	}
	if (n > cap) n = cap;
	memcpy (out, src, (size_t)n * 2 * sizeof (float));
	return n;
}
}
