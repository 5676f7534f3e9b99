/*
 * TEST INFRASTRUCTURE — not product code.
 *
 * Qt-free driver for the reference's OWN DSP classes.  Every arithmetic block called
 * here (Oscillator, SinCos, DecimatingFIR, fftFilter, fftFilterHilbert, pilotRecovery,
 * PerfectStereoSeparation, fm_Demodulator, pllC, compAtan) is the reference's source
 * file, compiled where it lies under /root/reference by oracle/Makefile.  Only the
 * sequencing below is ours: it restates, in the same order and with the same member
 * construction arguments,
 *     fmProcessor::fmProcessor             src/fm/fm-processor.cpp:48-198
 *     fmProcessor::run (per-sample loop)   src/fm/fm-processor.cpp:373-687
 *     fmProcessor::process_signal_with_rds src/fm/fm-processor.cpp:689-759
 *     the setters                          src/fm/fm-processor.cpp:232-301,762-770
 * fm-processor.cpp itself cannot be compiled here (QThread, signals, sndfile.h,
 * samplerate.h), which is why this harness exists.  The chain ends at the 192 kHz
 * de-emphasised, gain-corrected stereo stream: the next reference step is libsamplerate
 * (newconverter.cpp:55-80), which is not installed and not vendored.
 */
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <vector>
#include <complex>
#include <cmath>

// The table dumps below read private members of the reference classes (SinCos::Table,
// compAtan's tables, PerfectStereoSeparation::lpFilter, fm_Demodulator::K_FM).  The
// access-specifier override is confined to this harness translation unit; the reference
// .cpp files are compiled untouched, and class layout does not depend on access.
#define private public
#define protected public
#include "fm-constants.h"
#include "fir-filters.h"
#include "fft-filters.h"
#include "sincos.h"
#include "oscillator.h"
#include "pilot-recover.h"
#include "stereo-separation.h"
#include "fm-demodulator.h"
#include "squelchClass.h"
#include "fft-complex.h"
#include "costas.h"
#define private public       /* the dump below reads rdsDecoder_1's tables; nothing else is touched */
#include "rds-decoder-1.h"
#include "rds-decoder-2.h"
#include "rds-decoder-3.h"
#include "rds-blocksynchronizer.h"
#include "rds-group.h"
#undef private
#undef private
#undef protected

#include "chain_api.h"

// the signal body moc would generate for `signals: void setSquelchIsActive (bool)` (squelchClass.h):
// the GUI indicator is not part of the arithmetic
void	squelch::setSquelchIsActive (bool) {}


#define PILOT_FREQUENCY 19000
#define RDS_FREQUENCY (3 * PILOT_FREQUENCY)
#define RDS_RATE 24000
#define RDS_SAMPLE_DELAY (2 * (FFT_SIZE - PILOTFILTER_SIZE))

namespace {

struct fftFilterPeek : public fftFilter {
	fftFilterPeek (int32_t s, int d) : fftFilter (s, d) {}
	std::complex<float> *freq () { return filterVector; }
	int32_t size () const { return fftSize; }
};

enum { MODE_STEREO = 0, MODE_PANO = 1, MODE_MONO = 2 };
enum { S_STEREO, S_STEREO_SWAPPED, S_LEFT, S_RIGHT, S_LEFTplusRIGHT,
       S_LEFTminusRIGHT, S_LEFTminusRIGHT_Test };

static const char *decoderName (int code) {
	switch (code) {          // fm-demodulator.cpp:36-44
	   case 1: return "AM";
	   case 2: return "FM PLL Decoder";
	   case 3: return "FM Mixed Demod";
	   case 4: return "FM Complex Baseband Delay";
	   case 5: return "FM Real Baseband Delay";
	   case 6: return "FM Difference Based";
	}
	return "";
}

struct RefChain {
	chain_cfg	cfg;
	int32_t		inputRate, fmRate;
//	member order of fm-processor.h:172-185
	Oscillator	localOscillator;
	SinCos		mySinCos;
	DecimatingFIR	fmBand_1;
	DecimatingFIR	fmBand_2;
	fftFilterPeek	fmAudioFilter;
	fftFilterPeek	inputFilter;
	pilotRecovery	pilotRecover;
	PerfectStereoSeparation pPSS;
	fftFilterPeek	rdsBandPassFilter;
	fftFilterHilbert rdsHilbertFilter;
	fm_Demodulator	theDemodulator;      // owned by RadioInterface in the reference (radio.cpp:190)
	squelch		mySquelch;           // fm-processor.cpp:87-88
	DecimatingFIR	rdsDecimator;        // local of run (), fm-processor.cpp:382
	std::vector<float> rdsPhaseBuffer;
	int		rdsPhaseIndex;
	bool		inputFilterOn, fmAudioFilterActive;
	float		Lgain, Rgain;
	int32_t		loFrequency;
	float		pilotDelayPSS;
	DSPCOMPLEX	lastAudioSample;
	float		deemphAlpha, volumeFactor, panorama, leftChannel, rightChannel;
	DSPCOMPLEX	RfDC;
	float		rfDcAlpha;

	RefChain (const chain_cfg &c) :
	   cfg (c), inputRate (c.input_rate), fmRate (c.fm_rate),
	   localOscillator (c.input_rate),
	   mySinCos (c.fm_rate),
//	fm-processor.cpp:68-75 with IRate = inputRate / 6
	   fmBand_1 (4 * c.input_rate / (c.input_rate / 6) + 1,
	             c.fm_rate / 2, c.input_rate, c.input_rate / (c.input_rate / 6)),
	   fmBand_2 ((c.input_rate / 6) / c.fm_rate + 1,
	             c.fm_rate / 2, c.input_rate / 6, (c.input_rate / 6) / c.fm_rate),
	   fmAudioFilter (2 * 4096, 756),
	   inputFilter (2 * 32768, 251),
	   pilotRecover (c.fm_rate,
	                 ((float (PILOT_FREQUENCY)) / c.fm_rate) * (2 * M_PI),
	                 10 * (2 * M_PI) / c.fm_rate, &mySinCos),
	   pPSS (c.fm_rate, 10.0f / c.fm_rate, &mySinCos),
	   rdsBandPassFilter (FFT_SIZE, PILOTFILTER_SIZE),
	   rdsHilbertFilter (FFT_SIZE, PILOTFILTER_SIZE),
	   theDemodulator (c.fm_rate),
	   mySquelch (1, 70000, c.fm_rate / 20, c.fm_rate),
	   rdsDecimator (11, RDS_RATE / 2, c.fm_rate, c.fm_rate / RDS_RATE),
	   rdsPhaseBuffer (RDS_SAMPLE_DELAY, 0.0f) {
	   Lgain = c.lgain; Rgain = c.rgain;            // :110-111 / setAttenuation
	   loFrequency = c.lo_hz;                       // :144 / set_localOscillator
	   RfDC = DSPCOMPLEX (0, 0);                    // :135
	   rfDcAlpha = 1.0f / inputRate;                // :379
	   pilotDelayPSS = 0;                           // :160
	   rdsPhaseIndex = 0;                           // :170
	   lastAudioSample = 0;                         // :173
	   inputFilter. setLowPass (0.95 * fmRate / 2, inputRate);   // :148
	   inputFilterOn = false;
	   if (c.input_filter_hz > 0) {                 // setBandwidth :232-239 then run :397-401
	      inputFilter. setLowPass (c.input_filter_hz / 2, inputRate);
	      inputFilterOn = true;
	   }
	   fmAudioFilterActive = false;
	   if (c.lf_cutoff_hz > 0) {                    // setlfcutoff :762-770 then run :403-408
	      fmAudioFilter. setLowPass (c.lf_cutoff_hz, fmRate);
	      fmAudioFilterActive = true;
	   }
	   rdsBandPassFilter. setBand (RDS_FREQUENCY - RDS_WIDTH / 2,
	                               RDS_FREQUENCY + RDS_WIDTH / 2, fmRate);  // :166-168
	   {  // setDeemphasis :291-297
	      float Tau = 1000000.0 / c.deemph_us;
	      deemphAlpha = 1.0 / (float (fmRate) / Tau + 1.0);
	   }
	   volumeFactor = std::pow (10.0f, c.volume_db / 20.0f);             // :299-301
	   panorama = (float)c.panorama / 100.0f;                             // :277-280
	   leftChannel  = (c.balance > 0 ? (100 - c.balance) / 100.0 : 1.0f); // :282-286
	   rightChannel = (c.balance < 0 ? (100 + c.balance) / 100.0 : 1.0f);
	   theDemodulator. setDecoder (QString (decoderName (c.decoder)));
	   mySquelch. setSquelchLevel (c.squelch_value);                    // :410-413 (squelchValue != oldSquelchValue)
	}

//	fm-processor.cpp:689-759
	void process_signal_with_rds (const float demod,
	                              std::complex<float> *audioOut,
	                              std::complex<float> *rdsValueCmpl,
	                              float *phaseOut, bool *lockedOut) {
	   float currentPilotPhase = pilotRecover. getPilotPhase (5 * demod);
	   const bool pilotLocked = pilotRecover. isLocked ();
	   *phaseOut = currentPilotPhase;
	   *lockedOut = pilotLocked;

	   if (!pilotLocked) {
	      pilotDelayPSS = 0;
	      pPSS. reset ();
	   }

	   if (cfg.fm_mode != MODE_MONO && (pilotLocked || !cfg.auto_mono)) {
	      float phaseforLRDiff =
	                2 * (currentPilotPhase + M_PI_4 + 0) - pilotDelayPSS;
	      if (phaseforLRDiff < - 2 * M_PI)
	         phaseforLRDiff += 4 * M_PI;
	      phaseforLRDiff = fmod (phaseforLRDiff, 2 * M_PI);
	      pilotDelayPSS = cfg.pss_on ?
	                 pPSS. process_sample (demod, phaseforLRDiff) : 0;
	      float LRDiff =
	                2.0 * (cfg.sound_sel == S_LEFTminusRIGHT_Test ?
	                         mySinCos. getSin (phaseforLRDiff) :
	                         mySinCos. getCos (phaseforLRDiff)) * demod;
	      float LRPlus = demod;
	      *audioOut = DSPCOMPLEX (LRPlus, LRDiff);
	   }
	   else {
	      *audioOut = DSPCOMPLEX (demod, 0);
	   }

	   if (cfg.rds_on) {
	      float rdsBaseBp = rdsBandPassFilter. Pass (demod);
	      std::complex<float> rdsBaseHilb = rdsHilbertFilter. Pass (rdsBaseBp);
	      float thePhase = 3 * (rdsPhaseBuffer [rdsPhaseIndex] + 0);
	      rdsPhaseBuffer [rdsPhaseIndex] = currentPilotPhase;
	      rdsPhaseIndex = (rdsPhaseIndex + 1) % RDS_SAMPLE_DELAY;
	      std::complex<float> oscValue =
	                std::complex<float> (cos (thePhase), -sin (thePhase));
	      *rdsValueCmpl = oscValue * rdsBaseHilb;
	   }
	}

//	everything fmProcessor::run does with one fm-rate sample after the discriminator
//	(:515-648); also entered directly by ref_process_demod
	void after_demod (float demod, std::complex<float> v, const chain_taps *t,
	                  int64_t &nfm, int64_t &nrds) {
	   std::complex<float> audio;
	   std::complex<float> rdsDataCplx (0, 0);
	   float phase; bool locked;
	   process_signal_with_rds (demod, &audio, &rdsDataCplx, &phase, &locked);

	   const float sumLR  = real (audio);
	   const float diffLR = imag (audio);
	   const float diffLRWeightend =
	         diffLR * (cfg.fm_mode == MODE_PANO ? panorama : 1.0f);
	   const float left  = sumLR + diffLRWeightend;
	   const float right = sumLR - diffLRWeightend;
	   switch (cfg.sound_sel) {
	      default:
	      case S_STEREO:
	         audio = std::complex<float> (left, right); break;
	      case S_STEREO_SWAPPED:
	         audio = std::complex<float> (right, left); break;
	      case S_LEFT:
	         audio = std::complex<float> (left, left); break;
	      case S_RIGHT:
	         audio = std::complex<float> (right, right); break;
	      case S_LEFTplusRIGHT:
	         audio = std::complex<float> (sumLR, sumLR); break;
	      case S_LEFTminusRIGHT:
	      case S_LEFTminusRIGHT_Test:
	         audio = std::complex<float> (diffLRWeightend, diffLRWeightend);
	         break;
	   }
	   if (t -> fm_z) { t -> fm_z [2 * nfm] = real (v); t -> fm_z [2 * nfm + 1] = imag (v); }
	   if (t -> demod) t -> demod [nfm] = demod;
	   if (t -> pilot_phase) t -> pilot_phase [nfm] = phase;
	   if (t -> locked) t -> locked [nfm] = locked ? 1 : 0;
	   if (t -> pss_delay) t -> pss_delay [nfm] = pilotDelayPSS;
	   if (t -> lr) { t -> lr [2 * nfm] = real (audio); t -> lr [2 * nfm + 1] = imag (audio); }
	   if (t -> rds_cplx) {
	      t -> rds_cplx [2 * nfm] = real (rdsDataCplx);
	      t -> rds_cplx [2 * nfm + 1] = imag (rdsDataCplx);
	   }

	   if (cfg.rds_on) {
	      std::complex<float> rdsSample;
	      if (rdsDecimator. Pass (rdsDataCplx, &rdsSample)) {
	         if (t -> rds24) {
	            t -> rds24 [2 * nrds] = real (rdsSample);
	            t -> rds24 [2 * nrds + 1] = imag (rdsSample);
	         }
	         nrds ++;
	      }
	   }

	   if (fmAudioFilterActive)
	      audio = fmAudioFilter. Pass (audio);

	   audio = lastAudioSample =
	      (audio - lastAudioSample) * deemphAlpha + lastAudioSample;

//	audioGainCorrection :303-306
	   const float gl = volumeFactor * leftChannel * real (audio);
	   const float gr = volumeFactor * rightChannel * imag (audio);
	   if (t -> audio192) { t -> audio192 [2 * nfm] = gl; t -> audio192 [2 * nfm + 1] = gr; }
	}

//	fm-processor.cpp:423-446 (per pulled block) and :461-648 (per sample)
	int64_t process (const float *iq, int64_t n_in, const chain_taps *t,
	                 int64_t *n_rds24) {
	   int64_t nfm = 0, nrds = 0;
	   for (int64_t i = 0; i < n_in; i ++) {
	      DSPCOMPLEX x (iq [2 * i], iq [2 * i + 1]);
	      if (cfg.dc_remove) {
	         RfDC = (x - RfDC) * rfDcAlpha + RfDC;
	         constexpr float DCRlimit = 0.01f;
	         float rfDcReal = real (RfDC);
	         float rfDcImag = imag (RfDC);
	         if (rfDcReal > +DCRlimit) rfDcReal = +DCRlimit;
	         else if (rfDcReal < -DCRlimit) rfDcReal = -DCRlimit;
	         if (rfDcImag > +DCRlimit) rfDcImag = +DCRlimit;
	         else if (rfDcImag < -DCRlimit) rfDcImag = -DCRlimit;
	         x -= std::complex<float> (rfDcReal, rfDcImag);
	      }
	      std::complex<float> v =
	            std::complex<float> (real (x) * Lgain, imag (x) * Rgain);
	      v = v * localOscillator. nextValue (loFrequency);
	      if (inputFilterOn)
	         v = inputFilter. Pass (v);
	      if (inputRate / fmRate > 1) {
	         if (!fmBand_1. Pass (v, &v))
	            continue;
	         if (!fmBand_2. Pass (v, &v))
	            continue;
	      }
	      float demod = theDemodulator. demodulate (v);
	      switch (cfg.squelch_mode) {                                   // :499-510
	         case 1: demod = mySquelch. do_noise_squelch (demod); break;
	         case 2: demod = mySquelch. do_level_squelch (demod, theDemodulator. get_carrier_ampl ()); break;
	         default:;
	      }
	      after_demod (demod, v, t, nfm, nrds);
	      nfm ++;
	   }
	   if (n_rds24) *n_rds24 = nrds;
	   return nfm;
	}
};
}	// namespace

extern "C" {

void	*ref_create (const chain_cfg *cfg) { return new RefChain (*cfg); }
void	ref_destroy (void *h) { delete (RefChain *)h; }
int64_t	ref_process (void *h, const float *iq, int64_t n_in,
	             const chain_taps *taps, int64_t *n_rds24) {
	return ((RefChain *)h) -> process (iq, n_in, taps, n_rds24);
}

int64_t	ref_process_demod (void *h, const float *demod, int64_t n_fm,
	                   const chain_taps *taps, int64_t *n_rds24) {
RefChain *c = (RefChain *)h;
int64_t nfm = 0, nrds = 0;
	for (int64_t i = 0; i < n_fm; i ++) {
	   float d = demod [i];
	   if (c -> cfg.squelch_mode == 1) d = c -> mySquelch. do_noise_squelch (d);   // (the level squelch needs the carrier level)
	   c -> after_demod (d, std::complex<float> (0, 0), taps, nfm, nrds);
	   nfm ++;
	}
	if (n_rds24) *n_rds24 = nrds;
	return nfm;
}

//	The setters the GUI thread calls while run () is going (fm-processor.cpp:240-301, 855-933): the new
//	chain_cfg carries fm_mode, decoder, sound_sel, panorama, balance, volume, de-emphasis, auto_mono, pss_on,
//	IQ gains, LO, squelch; filters and rates are not touched.  actions: 1 = restartPssAnalyzer (:857-860),
//	2 = triggerFrequencyChange (:849-855: the PSS restart is the part inside this chain),
//	4 = setDCRemove (cfg -> dc_remove) (:924-927: also zeroes RfDC).
void	ref_update (void *h, const chain_cfg *n, int32_t actions) {
RefChain *c = (RefChain *)h;
	if (n -> decoder != c -> cfg.decoder)
	   c -> theDemodulator. setDecoder (QString (decoderName (n -> decoder)));
	c -> cfg.decoder = n -> decoder;
	c -> cfg.fm_mode = n -> fm_mode; c -> cfg.sound_sel = n -> sound_sel;
	c -> cfg.auto_mono = n -> auto_mono; c -> cfg.pss_on = n -> pss_on;
	c -> cfg.squelch_mode = n -> squelch_mode;
	if (n -> squelch_value != c -> cfg.squelch_value) c -> mySquelch. setSquelchLevel (n -> squelch_value);
	c -> cfg.squelch_value = n -> squelch_value;
	c -> Lgain = n -> lgain; c -> Rgain = n -> rgain;
	c -> loFrequency = n -> lo_hz;
	{
	   float Tau = 1000000.0 / n -> deemph_us;
	   c -> deemphAlpha = 1.0 / (float (c -> fmRate) / Tau + 1.0);
	}
	c -> volumeFactor = std::pow (10.0f, n -> volume_db / 20.0f);
	c -> panorama = (float)n -> panorama / 100.0f;
	c -> leftChannel  = (n -> balance > 0 ? (100 - n -> balance) / 100.0 : 1.0f);
	c -> rightChannel = (n -> balance < 0 ? (100 + n -> balance) / 100.0 : 1.0f);
	if (actions & 3) { c -> pilotDelayPSS = 0; c -> pPSS. reset (); }
	if (actions & 4) { c -> cfg.dc_remove = n -> dc_remove; c -> RfDC = DSPCOMPLEX (0, 0); }
}
void	ref_get_meta (void *h, chain_meta *m) {
RefChain *c = (RefChain *)h;
	m -> dc_rf_re = real (c -> RfDC);
	m -> dc_rf_im = imag (c -> RfDC);
	m -> dc_if = c -> theDemodulator. get_DcComponent ();
	m -> carrier_ampl = c -> theDemodulator. get_carrier_ampl ();
	m -> pss_phase_shift = c -> pilotDelayPSS;
	m -> pss_mean_error = c -> pPSS. get_mean_error ();
	m -> pss_minimized = c -> pPSS. is_error_minimized ();
	m -> pilot_lock_strength = c -> pilotRecover. getLockedStrength ();
	m -> pilot_locked = c -> pilotRecover. isLocked ();
	m -> squelch_active = c -> mySquelch. getSquelchActive ();
}

int32_t	ref_dump_taps (void *h, int which, float *out, int32_t cap) {
RefChain *c = (RefChain *)h;
const std::complex<float> *src = nullptr;
int32_t n = 0;
	switch (which) {
	   case DUMP_FMBAND1: src = c -> fmBand_1. getKernel (); n = c -> fmBand_1. filterSize; break;
	   case DUMP_FMBAND2: src = c -> fmBand_2. getKernel (); n = c -> fmBand_2. filterSize; break;
	   case DUMP_RDSDECIM: src = c -> rdsDecimator. getKernel (); n = c -> rdsDecimator. filterSize; break;
	   case DUMP_INPUT_FILTER_FREQ: src = c -> inputFilter. freq (); n = c -> inputFilter. size (); break;
	   case DUMP_RDS_BP_FREQ: src = c -> rdsBandPassFilter. freq (); n = c -> rdsBandPassFilter. size (); break;
	   case DUMP_AUDIO_LP_FREQ: src = c -> fmAudioFilter. freq (); n = c -> fmAudioFilter. size (); break;
	   case DUMP_PSS_LP_FREQ: src = c -> pPSS. lpFilter. filterVector; n = c -> pPSS. lpFilter. fftSize; break;
	   case DUMP_SINCOS: src = c -> mySinCos. Table; n = c -> mySinCos. Rate; break;
	   case DUMP_ATAN: {
	      compAtan &a = c -> theDemodulator. myAtan;
	      const float *tabs [8] = { a.ATAN2_TABLE_PPY, a.ATAN2_TABLE_PPX, a.ATAN2_TABLE_PNY,
	                                a.ATAN2_TABLE_PNX, a.ATAN2_TABLE_NPY, a.ATAN2_TABLE_NPX,
	                                a.ATAN2_TABLE_NNY, a.ATAN2_TABLE_NNX };
	      if (cap < 8 * 8193 / 2) return -1;
	      for (int t = 0; t < 8; t ++)
	         memcpy (out + t * 8193, tabs [t], 8193 * sizeof (float));
	      return 8 * 8193 / 2;
	   }
	   case DUMP_CONSTS: {
	      if (cap < 4) return -1;
	      out [0] = c -> theDemodulator. K_FM;   out [1] = c -> deemphAlpha;
	      out [2] = c -> volumeFactor;            out [3] = c -> pilotRecover. omega;
	      out [4] = c -> pilotRecover. gain;      out [5] = c -> pPSS. alpha;
	      out [6] = c -> pPSS. lockAlpha;         out [7] = c -> rfDcAlpha;
	      return 4;
	   }
	   case DUMP_SQUELCH_IIR: {
//	      the two filters mySquelch builds (squelchClass.cpp:12-21), constructed again here because
//	      the members are private: HighPassIIR (20, 70000 - 100, fs, S_CHEBYSHEV), LowPassIIR (20, 70000, ..)
	      if (cap < 41) return -1;
	      HighPassIIR hp (20, 70000 - 100, c -> fmRate, S_CHEBYSHEV);
	      LowPassIIR lp (20, 70000, c -> fmRate, S_CHEBYSHEV);
	      int k = 0;
	      Basic_IIR *f [2] = { &hp, &lp };
	      for (int w = 0; w < 2; w ++) {
	         out [k ++] = f [w] -> gain;
	         for (int i = 0; i < f [w] -> numofQuads; i ++) {
	            out [k ++] = f [w] -> Quads [i]. A1; out [k ++] = f [w] -> Quads [i]. A2;
	            out [k ++] = f [w] -> Quads [i]. B1; out [k ++] = f [w] -> Quads [i]. B2;
	         }
	      }
	      return 41;
	   }
	   default: return -1;
	}
	if (n > cap) n = cap;
	memcpy (out, src, (size_t)n * 2 * sizeof (float));
	return n;
}

//	---- RDS symbol stage, mode RDS_1 (rds-decoder.cpp:36-41, 69-82) -------------------------------
struct RefRds1 {
	Costas		my_costas;
	rdsDecoder_1	decoder;
	RefRds1 (int32_t rate): my_costas (rate, 1.0f / 16.0f, 0.02f / 16.0f, 10.0f), decoder (nullptr, rate) {}
};
void	*ref_rds1_create (int32_t rate) { return new RefRds1 (rate); }
void	ref_rds1_destroy (void *h) { delete (RefRds1 *)h; }
int64_t	ref_rds1_process (void *h, const float *rds24, int64_t n, uint8_t *bits, int64_t cap) {
RefRds1 *c = (RefRds1 *)h;
int64_t nb = 0;
	for (int64_t i = 0; i < n; i ++) {
	   DSPCOMPLEX v (rds24 [2 * i], rds24 [2 * i + 1]);
	   v = c -> my_costas. process_sample (v);
	   uint8_t theBit;
	   if (c -> decoder. doDecode (real (v), &theBit)) {
	      if (nb < cap) bits [nb] = theBit;
	      nb ++;
	   }
	}
	return nb;
}
//	magCplx of rdsDecoder::doDecode in mode RDS_1 (rds-decoder.cpp:75-77): 4 x the Costas output of every
//	24 kHz sample, from a fresh loop — what the LF scope shows as RDS_DEMOD (fm-processor.cpp:571-573)
void	ref_rds1_mag (int32_t rate, const float *rds24, int64_t n, float *out) {
Costas my_costas (rate, 1.0f / 16.0f, 0.02f / 16.0f, 10.0f);
	for (int64_t i = 0; i < n; i ++) {
	   DSPCOMPLEX v (rds24 [2 * i], rds24 [2 * i + 1]);
	   v = my_costas. process_sample (v);
	   const DSPCOMPLEX m = v * 4.0f;
	   out [2 * i] = real (m); out [2 * i + 1] = imag (m);
	}
}
int32_t	ref_rds1_dump (void *h, int which, float *out, int32_t cap) {
RefRds1 *c = (RefRds1 *)h;
	switch (which) {
	   case 0: {
	      const int n = c -> decoder. rdsBufferSize;
	      if (cap < n) return -1;
	      for (int i = 0; i < n; i ++) out [i] = c -> decoder. rdsKernel [i];
	      return n;
	   }
	   case 1: {
	      const int n = c -> decoder. rdsFilter. filterSize;
	      if (cap < n) return -1;
	      for (int i = 0; i < n; i ++) out [i] = real (c -> decoder. rdsFilter. filterKernel [i]);
	      return n;
	   }
	   case 2: {
	      Basic_IIR &f = c -> decoder. sharpFilter;
	      if (cap < 1 + 4 * f. numofQuads) return -1;
	      int k = 0;
	      out [k ++] = f. gain;
	      for (int i = 0; i < f. numofQuads; i ++) {
	         out [k ++] = f. Quads [i]. A1; out [k ++] = f. Quads [i]. A2;
	         out [k ++] = f. Quads [i]. B1; out [k ++] = f. Quads [i]. B2;
	      }
	      return k;
	   }
	}
	return -1;
}

//	---- station scan, fm-processor.cpp:478-495 with getSignal / getNoise (:886-904) restated ----------
int64_t	ref_scan_blocks (const float *fm_z, int64_t n, float *out) {
std::complex<float> scanBuffer [1024];
int64_t nb = 0;
	for (int64_t b = 0; (b + 1) * 1024 <= n; b ++) {
	   for (int i = 0; i < 1024; i ++)
	      scanBuffer [i] = std::complex<float> (fm_z [2 * (b * 1024 + i)], fm_z [2 * (b * 1024 + i) + 1]);
	   Fft_transform (scanBuffer, 1024, false);
	   float signal = 0, Noise = 0;
	   for (int i = 5; i < 25; i ++) signal += abs (scanBuffer [i]);
	   for (int i = 5; i < 25; i ++) signal += abs (scanBuffer [1024 - 1 - i]);
	   signal = signal / 40;
	   for (int i = 5; i < 25; i ++) Noise += abs (scanBuffer [1024 / 2 - 1 - i]);
	   for (int i = 5; i < 25; i ++) Noise += abs (scanBuffer [1024 / 2 + 1 + i]);
	   Noise = Noise / 40;
	   out [2 * nb] = get_db (signal, 256);
	   out [2 * nb + 1] = get_db (Noise, 256);
	   nb ++;
	}
	return nb;
}

//	---- LF scope display spectrum: ls_scope (src/scopes-qwt6/ls-scope.cpp), restated ------------------
//	ls-scope.cpp is a Qwt widget class and cannot be compiled here; the arithmetic of its constructor's
//	window (:50-52), processLFSpectrum (:76-92), mapSpectrum (:130-176) and add_to_average (:178-193) is
//	restated below around the reference's own Fft_transform.  v: the LF scope stream (complex); one
//	displayBuffer (display doubles) is written per completed block of spectrumSize samples after the
//	first (the first block has refresh = true: it only primes the average, :91-92).  Returns the blocks.
int64_t	ref_lf_spectrum (const float *v, int64_t n, int32_t spectrumSize, int32_t displaySize,
	                 int32_t averageCount, int32_t zoomFactor, int32_t showFull, double *display_out) {
std::vector<float> Window (spectrumSize);
std::vector<std::complex<float>> ftBuffer (spectrumSize);
std::vector<double> averageBuffer (displaySize, 0.0), displayBuffer (displaySize, 0.0), Y (displaySize);
	for (int i = 0; i < spectrumSize; i ++)
	   Window [i] = 0.43 - 0.5 * cos ((2.0 * M_PI * i) / spectrumSize)
	                     + 0.08 * cos ((4.0 * M_PI * i) / (spectrumSize - 1));
int64_t nb = 0;
bool refresh = true;
	for (int64_t b = 0; (b + 1) * spectrumSize <= n; b ++) {
	   for (int i = 0; i < spectrumSize; i ++) {
	      std::complex<float> tmp (v [2 * (b * spectrumSize + i)], v [2 * (b * spectrumSize + i) + 1]);
	      if (std::isinf (abs (tmp)) || std::isnan (abs (tmp)))
	         ftBuffer [i] = std::complex<float> (0, 0);
	      else
	         ftBuffer [i] = std::complex<float> (real (tmp) * Window [i], imag (tmp) * Window [i]);   // cmul
	   }
	   Fft_transform (ftBuffer. data (), spectrumSize, false);
	   int16_t factor = spectrumSize / displaySize;
	   factor /= 2;
	   int32_t zoom = zoomFactor;
	   if (factor / zoom >= 1) factor /= zoom;
	   else { zoom = factor; factor = 1; }
	   if (showFull) {
	      for (int32_t i = 0; i < displaySize / 2; i ++) {
	         double f = 0;
	         for (int32_t j = 0; j < factor; j ++) f += abs (ftBuffer [i * factor + j]);
	         Y [displaySize / 2 + i] = f / factor;
	         f = 0;
	         for (int32_t j = 0; j < factor; j ++) f += abs (ftBuffer [spectrumSize - 1 - (i * factor + j)]);
	         Y [displaySize / 2 - 1 - i] = f / factor;
	      }
	   }
	   else {
	      for (int32_t i = 0; i < displaySize; i ++) {
	         double f = 0;
	         for (int32_t j = 0; j < factor; j ++) f += abs (ftBuffer [i * factor + j]);
	         Y [i] = f / factor;
	      }
	   }
	   const double alpha = 1.0 / averageCount, beta = (averageCount - 1.0) / averageCount;
	   if (refresh) {
	      for (int32_t i = 0; i < displaySize; i ++) averageBuffer [i] = Y [i];
	      refresh = false;
	   }
	   else {
	      for (int32_t i = 0; i < displaySize; i ++) averageBuffer [i] = alpha * Y [i] + beta * averageBuffer [i];
	      for (int32_t i = 0; i < displaySize; i ++) displayBuffer [i] = averageBuffer [i];
	   }
	   memcpy (display_out + nb * displaySize, displayBuffer. data (), displaySize * sizeof (double));
	   nb ++;
	}
	return nb;
}
}

// ---- working-rate post-processing: insertTestTone, evaluatePeakLevel ---------------------------
// Members of the Qt class fmProcessor (src/fm/fm-processor.cpp:772-823; state in includes/fm/fm-processor.h:
// 54-75 DelayLine, 233-251), restated statement by statement around DSPCOMPLEX and PI_Constrain from the
// reference's own fm-constants.h.  Fed with working-rate PCM (the GPU's, behind its fade-in): the 192 -> 48 kHz
// step in front of it is libsamplerate in the reference (parity unpinned), so this is where the comparison starts.
namespace {
struct RefPost {
	int32_t workingRate;
	struct TestTone {                              // fm-processor.h:241-249
	   bool     Enabled = false;
	   float    TimePeriod = 2.0f;
	   float    SignalDuration = 0.025f;
	   uint32_t TimePeriodCounter = 0;
	   uint32_t NoSamplRemain = 0;
	   float    CurPhase = 0.0f;
	   float    PhaseIncr = 0.0f;
	} testTone;
	int32_t  peakLevelCurSampleCnt = 0, peakLevelSampleMax;
	DSPFLOAT absPeakLeft = 0, absPeakRight = 0;    // fm-processor.cpp:125-126
	uint32_t DataPtrIdx = 0;                       // DelayLine<DSPCOMPLEX> delayLine {(-40, -40)}
	std::vector<DSPCOMPLEX> DelayBuffer;
	DSPCOMPLEX mDefault = DSPCOMPLEX (-40.0f, -40.0f);
	explicit RefPost (int32_t wr) : workingRate (wr), peakLevelSampleMax (wr / 50) { set_delay_steps (0); }   // :142
	void set_delay_steps (uint32_t iSteps) { DataPtrIdx = 0; DelayBuffer. assign (iSteps + 1, mDefault); }
	// (the reference's resize (n, default) keeps old entries when growing; a fresh line is what
	//  setDispDelay yields on a processor whose line was never longer)
	const DSPCOMPLEX &get_set_value (const DSPCOMPLEX &iVal) {
	   DelayBuffer [DataPtrIdx] = iVal;
	   DataPtrIdx = (DataPtrIdx + 1) % DelayBuffer. size ();
	   return DelayBuffer [DataPtrIdx];
	}
	void insertTestTone (DSPCOMPLEX &ioS) {        // :800-823
	   float toneFreqHz = 1000.0f;
	   float level = 0.9f;
	   if (!testTone. Enabled)
	      return;
	   ioS *= (1.0f - level);
	   if (testTone. NoSamplRemain > 0) {
	      testTone. NoSamplRemain --;
	      testTone. CurPhase += testTone. PhaseIncr;
	      testTone. CurPhase = PI_Constrain (testTone. CurPhase);
	      const float smpl = sin (testTone. CurPhase);
	      ioS += level * DSPCOMPLEX (smpl, smpl);
	   }
	   else
	   if (++testTone. TimePeriodCounter > workingRate * testTone. TimePeriod) {
	      testTone. TimePeriodCounter = 0;
	      testTone. NoSamplRemain = workingRate * testTone. SignalDuration;
	      testTone. CurPhase = 0.0f;
	      testTone. PhaseIncr = 2 * M_PI / workingRate * toneFreqHz;
	   }
	}
	bool evaluatePeakLevel (const DSPCOMPLEX s, float *l, float *r) {     // :772-798; true: showPeakLevel emitted
	   const float absLeft  = std::abs (real (s));
	   const float absRight = std::abs (imag (s));
	   if (absLeft  > absPeakLeft)  absPeakLeft  = absLeft;
	   if (absRight > absPeakRight) absPeakRight = absRight;
	   peakLevelCurSampleCnt ++;
	   if (peakLevelCurSampleCnt > peakLevelSampleMax) {
	      peakLevelCurSampleCnt = 0;
	      float leftDb  = (absPeakLeft  > 0.0f ? 20.0f * std::log10 (absPeakLeft)  : -40.0f);
	      float rightDb = (absPeakRight > 0.0f ? 20.0f * std::log10 (absPeakRight) : -40.0f);
	      DSPCOMPLEX delayed = get_set_value (DSPCOMPLEX (leftDb, rightDb));
	      *l = real (delayed); *r = imag (delayed);
	      absPeakLeft = 0.0f; absPeakRight = 0.0f;
	      return true;
	   }
	   return false;
	}
};
}

extern "C" {
void	*ref_post_create (int32_t working_rate) { return new RefPost (working_rate); }
void	ref_post_destroy (void *h) { delete (RefPost *)h; }
// tone_on: setTestTone; delay_steps >= 0: setDispDelay, < 0: leave the line as it is
void	ref_post_set (void *h, int32_t tone_on, int32_t delay_steps) {
RefPost *p = (RefPost *)h;
	p -> testTone. Enabled = tone_on != 0;
	if (delay_steps >= 0) p -> set_delay_steps ((uint32_t)delay_steps);
}
// pcm: n working-rate (left, right) samples behind the fade-in -> pcm_out: what goes to sendSampletoOutput;
// peaks: the (left dB, right dB) pairs showPeakLevel is emitted with; returns their number
int64_t	ref_post_process (void *h, const float *pcm, int64_t n, float *pcm_out, float *peaks, int64_t cap_pairs) {
RefPost *p = (RefPost *)h;
int64_t ne = 0;
	for (int64_t i = 0; i < n; i ++) {
	   DSPCOMPLEX s (pcm [2 * i], pcm [2 * i + 1]);
	   p -> insertTestTone (s);
	   float l, r;
	   if (p -> evaluatePeakLevel (s, &l, &r) && ne < cap_pairs) { peaks [2 * ne] = l; peaks [2 * ne + 1] = r; ne ++; }
	   pcm_out [2 * i] = real (s); pcm_out [2 * i + 1] = imag (s);
	}
	return ne;
}
}

// ---- RDS symbol stage, mode RDS_2: the reference's own rdsDecoder_2 (src/rds/rds-decoder-2.cpp), fed sample by
// sample as rdsDecoder::doDecode does in case RDS_2 (src/rds/rds-decoder.cpp:84-88)
extern "C" {
void	*ref_rds2_create (int32_t rate) {
rdsDecoder_2 *d = new rdsDecoder_2 (nullptr, rate);
	d -> previousBit = false;          // left uninitialised by the reference's constructor
	return d;
}
void	ref_rds2_destroy (void *h) { delete (rdsDecoder_2 *)h; }
int64_t	ref_rds2_process (void *h, const float *rds24, int64_t n, uint8_t *bits, int64_t cap) {
rdsDecoder_2 *d = (rdsDecoder_2 *)h;
int64_t nb = 0;
	for (int64_t i = 0; i < n; i ++) {
	   std::complex<float> m;
	   uint8_t b;
	   if (d -> doDecode (std::complex<float> (rds24 [2 * i], rds24 [2 * i + 1]), &m, &b) && nb < cap) bits [nb ++] = b;
	}
	return nb;
}
int32_t	ref_rds2_dump (void *h, float *out, int32_t cap) {
rdsDecoder_2 *d = (rdsDecoder_2 *)h;
int32_t n = (int32_t)d -> my_matchedFltKernelVec. size ();
	if (n > cap) n = cap;
	for (int32_t i = 0; i < n; i ++) out [i] = d -> my_matchedFltKernelVec [i];
	return n;
}
}

// ---- RDS symbol stage, mode RDS_3: Costas (the one shared with mode 1) + the reference's rdsDecoder_3, which reads the
// block synchroniser's error count to re-synchronise its bit clock (src/rds/rds-decoder-3.cpp:96-101), so the
// synchroniser and the group are the reference's own too; sequencing of rdsDecoder::doDecode case RDS_3 and
// rdsDecoder::processBit (src/rds/rds-decoder.cpp:90-98, 104-131) without the group decoder (GUI strings).
// The signal bodies moc would generate for rds-blocksynchronizer.h:
void	rdsBlockSynchronizer::setRDSisSynchronized (bool) {}
void	rdsBlockSynchronizer::setbitErrorRate (double) {}
struct RefRds3 {
	RDSGroup		group;
	rdsBlockSynchronizer	sync;
	Costas			my_costas;
	rdsDecoder_3		decoder;
	RefRds3 (int32_t rate): sync (nullptr), my_costas (rate, 1.0f / 16.0f, 0.02f / 16.0f, 10.0f),
	                        decoder (nullptr, rate, &sync, &group, nullptr) {
	   group. clear ();
	   sync. setFecEnabled (true);
	}
//	returns true when a group was completed
	bool processBit (bool bit) {
	   switch (sync. pushBit (bit, &group)) {
	      case rdsBlockSynchronizer::RDS_NO_SYNC:
	      case rdsBlockSynchronizer::RDS_NO_CRC:
	         sync. resync ();
	         return false;
	      case rdsBlockSynchronizer::RDS_COMPLETE_GROUP:
	         return true;                 // (the caller copies the group, then clears it)
	      default:
	         return false;
	   }
	}
};
extern "C" {
void	*ref_rds3_create (int32_t rate) { return new RefRds3 (rate); }
void	ref_rds3_destroy (void *h) { delete (RefRds3 *)h; }
int64_t	ref_rds3_process (void *h, const float *rds24, int64_t n, uint8_t *bits, int64_t cap,
	                  uint16_t *groups, int64_t cap_groups, int64_t *n_groups, int32_t *n_resync) {
RefRds3 *c = (RefRds3 *)h;
int64_t nb = 0, ng = 0;
int32_t nrs = 0;
	for (int64_t i = 0; i < n; i ++) {
	   DSPCOMPLEX v (rds24 [2 * i], rds24 [2 * i + 1]);
	   v = c -> my_costas. process_sample (v);
	   if (c -> decoder. Resync || c -> sync. getNumSyncErrors () > 3) nrs ++;        // the decoder re-synchronises on this sample
	   uint8_t theBit;
	   if (c -> decoder. doDecode (real (v), &theBit)) {
	      if (nb < cap) bits [nb] = theBit;
	      nb ++;
	      if (c -> processBit (theBit)) {
	         if (groups && ng < cap_groups)
	            for (int b = 0; b < 4; b ++) groups [4 * ng + b] = c -> group. getBlock ((RDSGroup::RdsBlock)b);
	         ng ++;
	         c -> group. clear ();
	      }
	   }
	}
	if (n_groups) *n_groups = ng;
	if (n_resync) *n_resync = nrs;
	return nb;
}
int64_t	ref_blocksync_groups (const uint8_t *bits, int64_t n, uint16_t *groups, int64_t cap_groups) {
RDSGroup group;
rdsBlockSynchronizer sync (nullptr);
int64_t ng = 0;
	group. clear ();
	sync. setFecEnabled (true);
	for (int64_t i = 0; i < n; i ++) {
	   switch (sync. pushBit (bits [i] != 0, &group)) {
	      case rdsBlockSynchronizer::RDS_NO_SYNC:
	      case rdsBlockSynchronizer::RDS_NO_CRC:
	         sync. resync (); break;
	      case rdsBlockSynchronizer::RDS_COMPLETE_GROUP:
	         if (ng < cap_groups) for (int b = 0; b < 4; b ++) groups [4 * ng + b] = group. getBlock ((RDSGroup::RdsBlock)b);
	         ng ++;
	         group. clear ();
	         break;
	      default: break;
	   }
	}
	return ng;
}
}

// ---- HF scope display spectrum: hs_scope::addElement + doAverage (src/scopes-qwt6/hs-scope.cpp:102-151, 175-203) ----
// hs_scope is a Qwt widget class; its arithmetic is restated here around the reference's own Fft_transform.
// x: raw IQ; one displayBuffer (displaySize doubles, before Scope::Display's dB scaling) per completed segment.
extern "C" int64_t ref_hf_spectrum (const float *x, int64_t n, int32_t displaySize, int32_t sampleRate, int32_t freq,
                                    double *display_out, int64_t cap_blocks) {
const int32_t segmentSize = sampleRate / freq;                    // :41
const int32_t spectrumSize = 4 * displaySize, spectrumFillpoint = spectrumSize;   // :45-46
int32_t averageCount = 0;                                         // :44 (setAverager is commented out)
std::vector<float> Window (spectrumFillpoint);
std::vector<std::complex<float>> inputBuffer (spectrumFillpoint), ftBuffer (spectrumSize);
std::vector<double> displayBuffer (displaySize, 0.0), averageBuffer (displaySize, 0.0);
	for (int16_t i = 0; i < spectrumFillpoint; i ++)
	   Window [i] = 0.43 - 0.5 * cos ((2.0 * M_PI * i) / spectrumFillpoint)
	                     + 0.08 * cos ((4.0 * M_PI * i) / (spectrumFillpoint - 1));
int32_t fillPointer = 0, sampleCounter = 0;
int64_t nb = 0;
	for (int64_t s = 0; s < n; s ++) {
	   const std::complex<float> v (x [2 * s], x [2 * s + 1]);
	   if (fillPointer < spectrumFillpoint)
	      inputBuffer [fillPointer ++] = v;
	   sampleCounter ++;
	   if (sampleCounter < segmentSize)
	      continue;
	   fillPointer = 0;
	   sampleCounter = 0;
	   for (int i = 0; i < spectrumFillpoint; i ++) {
	      std::complex<float> tmp = inputBuffer [i];
	      if (std::isinf (abs (tmp)) || std::isnan (abs (tmp)))
	         ftBuffer [i] = std::complex<float> (0, 0);
	      else
	         ftBuffer [i] = std::complex<float> (real (tmp) * Window [i], imag (tmp) * Window [i]);   // cmul
	   }
	   for (int i = spectrumFillpoint; i < spectrumSize; i ++)
	      ftBuffer [i] = std::complex<float> (0, 0);
	   Fft_transform (ftBuffer. data (), spectrumSize, false);
	   int ratio = spectrumSize / displaySize;
	   for (int i = 0; i < displaySize / 2; i ++) {
	      float sum = 0;
	      for (int j = 0; j < ratio; j ++)
	         sum += abs (ftBuffer [i * ratio + j]);
	      displayBuffer [displaySize / 2 + i] = sum / ratio;
	      sum = 0;
	      for (int j = 0; j < ratio; j ++)
	         sum += abs (ftBuffer [spectrumSize / 2 + i * ratio + j]);
	      displayBuffer [i] = sum / ratio;
	   }
	   if (averageCount > 0) { /* never: see above */ }
	   else {
	      for (int i = 0; i < displaySize; i ++) {
	         if (displayBuffer [i] != displayBuffer [i])
	            displayBuffer [i] = 0;
	         averageBuffer [i] = ((double)(freq / 2 - 1)) / (freq / 2) * averageBuffer [i]
	                             + 1.0f / (freq / 2) * displayBuffer [i];
	         displayBuffer [i] = averageBuffer [i];
	      }
	   }
	   if (nb < cap_blocks) memcpy (display_out + nb * displaySize, displayBuffer. data (), displaySize * sizeof (double));
	   nb ++;
	}
	return nb;
}
