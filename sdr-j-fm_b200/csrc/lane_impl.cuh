// Launch sequencing of the B200 FM path for one lane of streams (included by sdrjfm_b200.cu).
//
// Per process call and per lane of streams the sequence mirrors the order of
// fmProcessor::run's per-sample loop (src/fm/fm-processor.cpp:461-648):
//   K1t / K1 / K1g  front end: DC block sums + polyphase FIR to the fm rate   (:423-446,:462-475)
//                   (K1t: TMA-fed, 2.304 MS/s complex float; K1g: any decimation / sample format;
//                    resampler mode: K1g /5 + resample_b_kernel)
//   K2  dc_tile + discriminator_kernel   DC subtract, gains, normalise, atan   (:497 -> fm-demodulator.cpp:111-195)
//   Ksc scan_kernel (only while scanning: nothing behind it runs)              (:478-495)
//   K3  pilot_kernel (parallel in time) | sequential_kernel (PLL / AM decoder, squelch)
//                                         AFC, pilot PLL, lock                  (fm-demodulator.cpp:197-241, pilot-recover.cpp, squelchClass.cpp)
//   K4  stereo_kernel                    PSS + 38 kHz demod + L/R selector     (:689-730,:517-549)
//   K5  rds_* on the lane's side stream  band-pass, Hilbert, x3 pilot mix, /8  (:733-758,:551-553)
//       (+ optional symbol stage: Costas, rdsDecoder_1 -> bits, running past the call's join)
//   K6  audio_kernel                     [audio low-pass] de-emphasis, gain, 192->48 kHz, fade-in   (:589-595,:630-642)
// There is NO CPU fallback: without a CUDA device sdrjfm_create fails with
// SDRJFM_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cstddef>
#include <algorithm>
#include <complex>
#include <string>
#include <vector>

#include "../../include/sdrjfm_b200.h"
#include "tables.hpp"
#include "common.cuh"
#include "frontend_fir.cuh"
#include "frontend_poly.cuh"
#include "frontend_tma.cuh"
#include "frontend_tmab.cuh"
#include "frontend_exact.cuh"
#include "resample.cuh"
#include "discriminator.cuh"
#include "sequential.cuh"
#include "pilot.cuh"
#include "stereo.cuh"
#include "rds.cuh"
#include "scan.cuh"
#include "spectrum.cuh"
#include "audio_out.cuh"

// One LANE = the complete launch sequence and state for a group of IQ streams on its own CUDA
// stream.  The C ABI (sdrjfm_b200.cu) splits a handle's streams over a few lanes so that the
// latency-bound per-stream kernels of one lane overlap the kernels of the others.
#pragma once
using namespace sdrjfm;

static thread_local std::string g_create_error;
struct Lane;
static int lane_destroy (Lane *h);
static int lane_restart_pss_analyzer (Lane *h);
static int lane_get_meta (Lane *h, sdrjfm_meta *meta);
static int run_hf_spectrum (Lane *h, const void *d_iq, RawFmt rf, int64_t pitch, int64_t n_in);
static cudaError_t poly_set_attr (int shape);
static RawFmt make_rawfmt (const Lane *h, int32_t fmt, float scale);


// front-end shapes: decimation, outputs per thread, tap groups narrow / with inputFilter
struct FeShape { int D, gpt, ng, ngw; };
static const FeShape kFeShapes [] = { { 12, 4, 4, 25 }, { 30, 2, 2, 11 }, { 48, 1, 2, 7 },
                                      { kRsStageADecim, 4, 10, 0 } };     // last: stage A of the rational resampler
constexpr int kShapeResample = 3;
constexpr int kShapeGeneric = 4;          // any other decimation 6 * D2: reference-order front end only (frontend_exact.cuh)
constexpr int kInputFilterDelay = kInputFftSize - kInputDegree;       // 65285 input samples

static inline FeShape fe_shape (int shape, int decim) {
	if (shape == kShapeGeneric) { const FeShape g = { decim, 1, 2, 0 }; return g; }
	return kFeShapes [shape];
}

// Host image of everything this configuration keeps in __constant__ memory.  The constant banks
// belong to the one loaded module, i.e. they are shared by every handle of the process: each lane
// keeps its image and a content signature, and whoever launches makes sure the device holds ITS
// image (consts_ensure).  Handles with different configurations may therefore coexist; they must
// not be driven from different threads at the same time.
struct ConstImage {
	float  comp [40];
	float  rs_taps [kRsTaps + 3];
	float  wide [kDecim][kFwGroups + 3];
	float  poly [kPolyMaxTaps];
	float2 pss [kPssTaps + 1];
	float  alp [kAlpTaps];
	float2 fx [kFxTaps1 + kFxMaxTaps2];
	uint64_t sig;
};
// signature of the image each DEVICE holds now (the constant banks are per device and context:
// a second handle on another device of the same process needs its own upload)
constexpr int kMaxDevices = 64;
static uint64_t g_const_sig [kMaxDevices] = { 0 };

struct Lane {
	sdrjfm_config cfg;
	Settings      set;
	TableBlob     tables;
	float        *d_tables = nullptr;
	cudaStream_t  stream = nullptr;
	cudaStream_t  stream_rds = nullptr;     // the RDS branch runs beside the stereo decoder (both only read K3's outputs)
	cudaEvent_t   ev_k3 = nullptr, ev_rds = nullptr, ev_k2 = nullptr;
	cudaStream_t  stream_k3 = nullptr;      // the pilot stage of time slice j + 1 beside K4 / K5 / K6 of slice j
	int32_t       slice_fm = 0;             // fm-rate samples per time slice (0: calls are not sliced; SDRJFM_FM_SLICE)
	int           n_sm = 0;
	bool          smem_lut_ok = false;
	SinLut        lut;
	float        *d_sin_quarter = nullptr;

	int64_t cap_in = 0, cap_fm = 0, cap_audio = 0, cap_rds = 0;   // per-stream capacities (pitches)
	// front-end shape: decimation D = 6 * (inputRate / 6 / fmRate) and the kernel instantiation for it
	int           decim = 12, shape = 0;    // shape indexes kFeShapes
	int           hist_len = 0, hist_len_w = 0;   // raw-sample history kept per stream (narrow / input filter on)
	int           fw_delay = 0, fw_shift = 0;     // inputFilter latency 65285 = decim * fw_delay + fw_shift
	int           ngw = 0;                        // tap groups of the wide composite
	bool          use_tma = true;           // SDRJFM_NO_TMA=1: the register-staged K1 instead of K1t
	int           tma_ctas = 0;             // persistent CTAs of K1t per launch (SDRJFM_TMA_CTAS overrides)
	bool          force_generic = false;    // SDRJFM_GENERIC_FE=1: K1g also where the tuned D = 12 kernels apply
	float2 *d_in = nullptr;                 // staging [S][cap_in] (8 bytes per sample: any format fits)
	float2 *d_hist [2] = { nullptr, nullptr }; int hist_sel = 0;
	float2 *d_pend = nullptr; int pend = 0; // leftover raw samples (< decim per stream), in format pend_fmt
	int     pend_fmt = 0;
	// reference-order front end (frontend_exact.cuh; allocated when first used)
	bool    auto_exact = true;              // SDRJFM_NO_AUTO_EXACT=1: only front_end_mode 2 selects it
	float2 *d_xd = nullptr;                 // [S][cap_in] samples behind the per-sample DC remover
	float2 *d_xhist [2] = { nullptr, nullptr }; int xhist_sel = 0;
	bool    fx_hist_valid = false;          // d_xhist continues the stream (false after a composite call)
	bool    seq_dc = false;                 // SDRJFM_SEQ_DC=1: the lane-per-stream DC walker instead of the parallel solver (cross-check)
	bool    hybrid_last = false;            // the last call ran the per-sample DC remover in front of the wide composite
	double2 *d_dcnow = nullptr;             // [S] RfDC advanced through the fm-rate delay line (metadata, inputFilter on)
	int64_t in_total = 0;                   // input samples consumed so far (per stream)
	// rational polyphase resampler (front_end_mode 1): stage-A output and the stage-B state
	bool    resample = false;
	int     rsL = 0, rsM = 0, rsP = 0, rsHB = 0;
	int64_t cap_a = 0, a_total = 0;         // stage-A samples per call (pitch) / produced so far
	float2 *d_A = nullptr, *d_SA = nullptr;
	float2 *d_bha [2] = { nullptr, nullptr }, *d_bhs [2] = { nullptr, nullptr }; int bh_sel = 0;
	float2 *d_U = nullptr, *d_S = nullptr, *d_iqn = nullptr, *d_fmz = nullptr;
	float  *d_res = nullptr, *d_zabs = nullptr, *d_demod = nullptr, *d_phase = nullptr;
	float  *d_pssd = nullptr;
	uint8_t *d_locked = nullptr;
	float2 *d_lr = nullptr, *d_a192 = nullptr, *d_rdsc = nullptr, *d_rds24 = nullptr;
	float2 *d_ahist [2] = { nullptr, nullptr }; int ahist_sel = 0;
	float2 *d_audio = nullptr;              // [S][cap_audio] working-rate stereo
	StreamState *d_state = nullptr;
	// input filter ON (allocated when first switched on): wide front end + fm-rate delay lines
	float2 *d_histw [2] = { nullptr, nullptr }; int histw_sel = 0;
	float2 *d_Uw = nullptr, *d_Sw = nullptr;
	float2 *d_udel [2] = { nullptr, nullptr }, *d_sdel [2] = { nullptr, nullptr }; int del_sel = 0;
	float   wide_sumC = 0, wide_sumCm = 0;
	// local oscillator (table built when lo first becomes non-zero)
	float2 *d_lo_tab = nullptr;
	int64_t lo_phase = 0;                   // Oscillator::LOPhase after the last processed sample
	float   lo_Hre = 1.f, lo_Him = 0.f;     // H (lo) of the active composite taps
	// audio low-pass (allocated when first switched on)
	float2 *d_alp_hist [2] = { nullptr, nullptr }; int alp_sel = 0;
	float2 *d_lrf = nullptr;
	// RDS branch (allocated when RDS is first switched on)
	float   *d_rds_dring = nullptr, *d_rds_pring = nullptr;     // [S][131072] demod / pilot phase by rds index
	float   *d_rds_bp = nullptr, *d_rds_hi = nullptr;           // [S][2][32000] / [S][2][32768]
	float2  *d_rds_R = nullptr, *d_rds_tw = nullptr, *d_rds_tws = nullptr, *d_rds_dtaps = nullptr;
	float2  *d_rds_hist [2] = { nullptr, nullptr }; int rds_hist_sel = 0;
	int64_t  rds_total = 0, rds_last_block = -1;
	// RDS symbol stage (mode RDS_1), optional: Costas loop + rdsDecoder_1 -> bits
	bool     rds_symbols = false;
	RdsSymState *d_rsy_state = nullptr; uint8_t *d_rsy_bits = nullptr; int32_t *d_rsy_nbits = nullptr;
	float   *d_rsy_c = nullptr, *d_rsy_v = nullptr, *d_rsy_w = nullptr;    // [S][cap_rds] Costas / low-pass / matched-filter outputs
	float2  *d_rsy_in = nullptr;            // [S][cap_rds] private copy of the 24 kHz baseband of the call
	Rds2State *d_rs2_state = nullptr; float2 *d_rs2_m = nullptr;     // mode RDS_2: rdsDecoder_2 state, matched-filter output
	Rds3State *d_rs3_state = nullptr; float *d_rs3_sin = nullptr;     // mode RDS_3: rdsDecoder_3 + block synchroniser, SinCos (24000)
	uint16_t *d_rs3_groups = nullptr; int32_t *d_rs3_ngroups = nullptr, *d_rs3_stat = nullptr;
	int32_t  cap_groups = 0; float rs3_omega = 0.f; int8_t rs3_kmap [kRs3Sym] = { 0 };
	int32_t  cap_bits = 0;
	dcplx   *d_tileB = nullptr; DiscrSnap *d_snap = nullptr; int32_t ntiles_cap = 0;   // K2 pre-pass
	float2  *d_pss_ring = nullptr;          // [S][2048] PSS filter input ring (state)
	int32_t *d_iter_stats = nullptr;        // pilot_kernel diagnostics: [S][4]
	SquelchState *d_sq = nullptr;           // [S], allocated when the squelch is first switched on
	// LF scope stream (setlfPlotType): -1 = no stream wanted, else ELfPlot; d_plot holds the float streams
	// the chain does not keep otherwise (diffLR, pre-gain audio) for the last call
	int32_t  lf_plot = -1;
	float   *d_plot = nullptr;
	const float2 *last_rds_ptr = nullptr; int64_t last_rds_pitch = 0;
	// LF scope display spectrum (ls_scope::processLFSpectrum on the selected stream); spec_N = 0: off
	int32_t  spec_N = 0, spec_logN = 0, spec_display = 0, spec_avg_count = 1, spec_zoom = 1;
	bool     spec_refresh = true;           // lfBuffer_newFlag
	float2  *d_spec_in = nullptr;           // [S][cap_fm] the stream of the call as complex samples
	float2  *d_spec_carry [2] = { nullptr, nullptr }; int spec_sel = 0, spec_carry = 0;
	float   *d_spec_win = nullptr;
	double  *d_spec_Y = nullptr, *d_spec_avg = nullptr, *d_spec_disp = nullptr;
	int32_t  cap_specblk = 0, last_nspec = 0;
	cudaEvent_t ev_sym = nullptr;           // Costas output of the call ready (RDS_DEMOD spectrum)
	// HF scope display spectrum (hs_scope::addElement on the raw input); hf_N = 0: off
	int32_t  hf_N = 0, hf_logN = 0, hf_display = 0, hf_seg = 0, hf_half_freq = 1, last_nhf = 0;
	int64_t  hf_total = 0;                  // raw input samples seen since the HF spectrum was switched on
	float2  *d_hf_blk = nullptr; float *d_hf_win = nullptr; double *d_hf_avg = nullptr, *d_hf_disp = nullptr;
	// station scan (startScanning / stopScanning): 1024-sample blocks of fm-rate samples -> (signal, noise) dB
	bool     scanning = false;
	float2  *d_scan_carry [2] = { nullptr, nullptr }; int scan_sel = 0, scan_carry = 0;
	float2  *d_scan_db = nullptr; int32_t cap_scan = 0, last_nscan = 0;
	// airspy native-rate input (kFmtAirspy): per-millisecond linear interpolation tables of the handler
	int32_t  air_blk = 0;                   // native samples per millisecond block (native rate / 1000)
	int16_t *d_air_int = nullptr; float *d_air_frac = nullptr;
	short2  *d_air_pend = nullptr; int air_pend = 0;      // [S][air_blk + 1] native samples carried to the next call
	bool    pilot_wide = false;             // SDRJFM_PILOT_WIDE=1: 1024 threads, 8192-sample windows (experiment: no gain, see lane_create)
	bool    sequential_pll = false;         // SDRJFM_SEQUENTIAL_PLL=1: lane-per-stream K3 (cross-check)
	int64_t fm_total = 0;                   // fm-rate samples produced so far (per stream)
	int32_t fade_cnt = 0, fade_max = 0;     // suppressAudioSampleCnt(Max), fm-processor.cpp:130-131
	// test tone (setTestTone / insertTestTone), peak meter (evaluatePeakLevel, setDispDelay), second converter
	bool     tone_on = false; int32_t tone_arm = 0, tone_burst = 0; int64_t tone_pos = 0;
	float   *d_tone_tab = nullptr;
	int32_t  peak_block = 961, peak_sel = 0, peak_delay = 0;
	int64_t  peak_e_set = 0, peak_e0 = 0, peak_e1 = 0;   // emission index at the last setDispDelay; emissions of the last call
	float2  *d_peak_ring = nullptr;
	int32_t  cvL = 1, cvM = 1;              // audio_rate / working_rate = cvL / cvM
	int64_t  cap_out = 0;                   // per-stream capacity of the audio output at audio_rate
	int64_t  cv_in_total = 0;               // working-rate samples fed to the converter so far
	float   *d_cv_taps = nullptr; float2 *d_cv_hist [2] = { nullptr, nullptr }; int cv_sel = 0;
	int64_t last_nfm = 0, last_naudio = 0, last_nrds = 0;
	int64_t launches = 0;
	ConstImage ci;
	std::string err;
};

#define CK(call)                                                                          \
	do { cudaError_t e_ = (call); if (e_ != cudaSuccess) {                                \
	   char b_ [256]; snprintf (b_, sizeof b_, "%s failed: %s (%s:%d)", #call,           \
	                           cudaGetErrorString (e_), __FILE__, __LINE__);              \
	   h -> err = b_; return SDRJFM_ERR_CUDA; } } while (0)

template <typename T> static cudaError_t dalloc (T **p, size_t n) {
	cudaError_t e = cudaMalloc ((void **)p, n * sizeof (T));
	if (e == cudaSuccess) e = cudaMemset (*p, 0, n * sizeof (T));
	return e;
}

static void default_settings (Settings &s, int32_t fm_rate) {
	memset (&s, 0, sizeof s);
	s.fm_mode = 0;            // FM_Mode::Stereo, fm-processor.cpp:155
	s.decoder = 3;            // MIXED, fm-demodulator.cpp:66
	s.sound_sel = 0;          // S_STEREO, :156
	s.rds_mode = 0;           // RDS_OFF, :133
	s.auto_mono = 1; s.pss_on = 1; s.dc_remove = 1;     // :121,:122,:134
	s.lgain = s.rgain = 1.0f; // :110-111
	s.volume = 0.5f;          // :127
	s.panorama = 1.0f;        // :128
	s.left_ch = s.right_ch = 1.0f;                      // :157-158
	s.squelch_value = 0;      // squelchValue, fm-processor.cpp:194
	s.deemph_us = 50;
	{  // the constructor's own formula (:174) is always overwritten by setDeemphasis through
	   // make_newProcessor (radio.cpp:940, default 50 us radio.cpp:2129); start from the latter
	   float Tau = 1000000.0 / 50;
	   s.deemph_alpha = 1.0 / (float (fm_rate) / Tau + 1.0);
	}
}

// copies the lane's constant image to the device (all launches of the process are drained first:
// kernels of another configuration may still be reading the banks)
static int consts_upload (Lane *h) {
	CK (cudaSetDevice (h -> cfg.device));
	CK (cudaDeviceSynchronize ());
	CK (cudaMemcpyToSymbol (c_comp, h -> ci.comp, sizeof h -> ci.comp));
	CK (cudaMemcpyToSymbol (c_rs_taps, h -> ci.rs_taps, sizeof h -> ci.rs_taps));
	CK (cudaMemcpyToSymbol (c_wide, h -> ci.wide, sizeof h -> ci.wide));
	CK (cudaMemcpyToSymbol (c_poly, h -> ci.poly, sizeof h -> ci.poly));
	CK (cudaMemcpyToSymbol (c_pss_taps, h -> ci.pss, sizeof h -> ci.pss));
	CK (cudaMemcpyToSymbol (c_alp_taps, h -> ci.alp, sizeof h -> ci.alp));
	CK (cudaMemcpyToSymbol (c_fx_taps, h -> ci.fx, sizeof h -> ci.fx));
	g_const_sig [h -> cfg.device] = h -> ci.sig;
	return SDRJFM_OK;
}
// new contents: recompute the signature (FNV-1a over the image) and upload
static int consts_commit (Lane *h) {
uint64_t x = 1469598103934665603ull;
const unsigned char *p = reinterpret_cast<const unsigned char *>(&h -> ci);
	for (size_t i = 0; i < offsetof (ConstImage, sig); i ++) { x ^= p [i]; x *= 1099511628211ull; }
	h -> ci.sig = x ? x : 1;
	if (h -> ci.sig == g_const_sig [h -> cfg.device]) return SDRJFM_OK;
	return consts_upload (h);
}
// before launching: the device must hold this lane's image
static inline int consts_ensure (Lane *h) {
	return h -> ci.sig == g_const_sig [h -> cfg.device] ? SDRJFM_OK : consts_upload (h);
}

// uploads constant-memory taps and derives the launch parameters that depend on the tables
static int upload_tables (Lane *h) {
const TableHeader &th = h -> tables.hdr ();
	if (h -> d_tables) { cudaFree (h -> d_tables); h -> d_tables = nullptr; }
	CK (cudaMalloc ((void **)&h -> d_tables, th.payload_floats * sizeof (float)));
	CK (cudaMemcpy (h -> d_tables, h -> tables.payload (), th.payload_floats * sizeof (float),
	                cudaMemcpyHostToDevice));
float comp [96] = { 0 };
	memcpy (comp, h -> tables.payload () + th.off_comp, th.ncomp * sizeof (float));
const int32_t lo_hz = h -> set.lo_hz;
	h -> lo_Hre = 1.f; h -> lo_Him = 0.f;
	if (lo_hz != 0) {
//	With the oscillator on, the DC folding of tables.cpp does not apply (the subtracted DC is rotated
//	sample by sample): run the plain composite C[d1 j + i] = t2[j] t1[i] and take the DC term out at the
//	fm rate through H (lo) = sum_t C[t] exp (+2 pi i lo t / inputRate)  (discriminator.cuh).
	   const cf32 *k1 = reinterpret_cast<const cf32 *>(h -> tables.payload () + th.off_fmband1);
	   const cf32 *k2 = reinterpret_cast<const cf32 *>(h -> tables.payload () + th.off_fmband2);
	   std::vector<double> cd (th.ncomp, 0.0);
	   for (int j = 0; j < th.ntaps2; j ++)
	      for (int i = 0; i < th.ntaps1; i ++)
	         cd [th.decim1 * j + i] += (double)k2 [j].imag () * (double)k1 [i].imag ();
	   std::complex<double> H (0, 0);
	   for (int t = 0; t < th.ncomp; t ++) {
	      comp [t] = (float)cd [t];
	      H += (double)comp [t] * std::polar (1.0, 2 * M_PI * (double)lo_hz * t / th.input_rate);
	   }
	   h -> lo_Hre = (float)H.real (); h -> lo_Him = (float)H.imag ();
	}
	memcpy (h -> ci.comp, comp, sizeof h -> ci.comp);
	{  // K1x: the two filterKernels exactly as DecimatingFIR builds them (complex, fir-filters.cpp:327-347)
	   memset (h -> ci.fx, 0, sizeof h -> ci.fx);
	   if (th.ntaps1 == kFxTaps1 && th.ntaps2 <= kFxMaxTaps2) {
	      memcpy (h -> ci.fx, h -> tables.payload () + th.off_fmband1, kFxTaps1 * sizeof (float2));
	      memcpy (h -> ci.fx + kFxTaps1, h -> tables.payload () + th.off_fmband2, th.ntaps2 * sizeof (float2));
	   }
	}

//	audio decimator taps (our own design, audio_out.cuh): Blackman-windowed sinc, fc = 20 kHz
	{
	   float t [kRsTaps + 3] = { 0 };
	   double sum = 0;
	   std::vector<double> d (kRsTaps);
	   const double fc = 20000.0 / th.fm_rate;
	   for (int i = 0; i < kRsTaps; i ++) {
	      const int k = i - kRsTaps / 2;
	      const double s = k == 0 ? 2 * fc : sin (2 * M_PI * fc * k) / (M_PI * k);
	      const double w = 0.42 - 0.5 * cos (2 * M_PI * i / (kRsTaps - 1))
	                            + 0.08 * cos (4 * M_PI * i / (kRsTaps - 1));
	      d [i] = s * w; sum += d [i];
	   }
	   for (int i = 0; i < kRsTaps; i ++) t [i] = (float)(d [i] / sum);
	   memcpy (h -> ci.rs_taps, t, sizeof t);
	}

//	K1g taps, narrow: c_poly[p][g] = C'[D g + D - 1 - p]
const int D = h -> decim;
const FeShape fs = fe_shape (h -> shape, h -> decim);
float cpoly [kPolyMaxTaps];
	memset (cpoly, 0, sizeof cpoly);
	if (h -> resample) {      // stage A of the rational resampler: its own 49 taps, no DC folding
	   if (th.rs_L != h -> rsL || th.rs_M != h -> rsM || th.rs_P != h -> rsP || th.rs_ntapsA > D * fs.ng) {
	      h -> err = "table blob carries no (or another) resampler design"; return SDRJFM_ERR_ARG;
	   }
	   memset (comp, 0, sizeof comp);
	   memcpy (comp, h -> tables.payload () + th.off_rsA, th.rs_ntapsA * sizeof (float));
	}
const int ncomp = h -> resample ? th.rs_ntapsA : th.ncomp;
	for (int g = 0; g < fs.ng; g ++)
	   for (int p = 0; p < D; p ++) {
	      const int i = D * g + D - 1 - p;
	      cpoly [p * fs.ng + g] = i < ncomp ? comp [i] : 0.f;
	   }

//	input filter ON: composite of the 251-tap low-pass and the decimator cascade, shifted by
//	fw_shift samples (frontend_fir.cuh), RF DC removal folded in exactly like the narrow taps (tables.cpp)
	if (th.ncomp_wide > 0) {
	   const float *wide = h -> tables.payload () + th.off_comp_wide;
	   const int len = fs.ngw * D;
	   if (th.ncomp_wide + h -> fw_shift > len) { h -> err = "wide composite does not fit its tap groups"; return SDRJFM_ERR_UNSUPPORTED; }
	   std::vector<double> cw (len, 0.0), g (len, 0.0);
	   for (int t = 0; t < th.ncomp_wide; t ++) cw [t + h -> fw_shift] = (double)wide [t];
	   for (int i = 0; i < len; i ++)
	      for (int k = i + 1; k < len; k ++) g [i] += cw [k];
	   const double alpha = lo_hz != 0 ? 0.0 : (double)(1.0f / th.input_rate);
	   float cwide [kDecim][kFwGroups + 3];
	   memset (cwide, 0, sizeof cwide);
	   memset (cpoly, 0, sizeof cpoly);
	   double sC = 0, sCm = 0;
	   for (int i = 0; i < len; i ++) {
	      const float f = (float)(cw [i] + alpha * g [i]);
	      sC += cw [i]; sCm += f;
	      if (D == kDecim) cwide [11 - i % kDecim][i / kDecim] = f;
	      cpoly [(D - 1 - i % D) * fs.ngw + i / D] = f;
	   }
	   h -> wide_sumC = (float)sC; h -> wide_sumCm = (float)sCm;
	   if (lo_hz != 0) {
	      std::complex<double> H (0, 0);
	      for (int t = 0; t < len; t ++)
	         H += (double)(float)cw [t] * std::polar (1.0, 2 * M_PI * (double)lo_hz * t / th.input_rate);
	      h -> lo_Hre = (float)H.real (); h -> lo_Him = (float)H.imag ();
	   }
	   if (D == kDecim) memcpy (h -> ci.wide, cwide, sizeof cwide);
	}
	memcpy (h -> ci.poly, cpoly, sizeof cpoly);

//	PSS low-pass taps: lpFilter (2048, 295).setLowPass (15000, rate), stereo-separation.cpp:31-39
	{
	   std::vector<cf32> lp = design_lowpass (kPssTaps, 15000, th.fm_rate);
	   float2 t [kPssTaps + 1];
	   memset (t, 0, sizeof t);
	   for (int i = 0; i < kPssTaps; i ++) t [i] = make_float2 (lp [i].real (), lp [i].real ());
	   memcpy (h -> ci.pss, t, sizeof t);
	}

//	quarter-wave sine table + exception list (see sequential.cuh)
const cf32 *sc = reinterpret_cast<const cf32 *>(h -> tables.payload () + th.off_sincos);
const int32_t R = th.fm_rate, Q = R / 4;
SinLut &L = h -> lut;
	memset (&L, 0, sizeof L);
	L.rate = R; L.quarter = Q; L.C = R / (2 * M_PI);
	for (int e = 0; e < kMaxSinExc; e ++) L.sin_exc_idx [e] = L.cos_exc_idx [e] = -1;
	h -> smem_lut_ok = (R % 4 == 0);
	if (h -> smem_lut_ok) {
	   std::vector<float> q (Q + 1);
	   for (int i = 0; i <= Q; i ++) q [i] = sc [i].imag ();
	   auto refl = [&](int idx) -> float {
	      if (idx <= Q) return q [idx];
	      if (idx <= 2 * Q) return q [2 * Q - idx];
	      if (idx <= 3 * Q) return -q [idx - 2 * Q];
	      return -q [R - idx];
	   };
	   int ns = 0, nc = 0;
	   for (int i = 0; i < R && h -> smem_lut_ok; i ++) {
	      float a = refl (i), b = sc [i].imag ();
	      if (memcmp (&a, &b, 4) != 0 && !(a == 0.f && b == 0.f)) {
	         if (ns < kMaxSinExc) { L.sin_exc_idx [ns] = i; L.sin_exc_val [ns] = b; ns ++; }
	         else h -> smem_lut_ok = false;
	      }
	      a = refl ((i + Q) % R); b = sc [i].real ();
	      if (memcmp (&a, &b, 4) != 0 && !(a == 0.f && b == 0.f)) {
	         if (nc < kMaxSinExc) { L.cos_exc_idx [nc] = i; L.cos_exc_val [nc] = b; nc ++; }
	         else h -> smem_lut_ok = false;
	      }
	   }
	   L.sin_at_half = sc [R / 2].imag ();
	   for (int e = 0; e < ns; e ++) if (L.sin_exc_idx [e] != R / 2) h -> smem_lut_ok = false;
	   if (h -> smem_lut_ok) {
	      if (h -> d_sin_quarter) cudaFree (h -> d_sin_quarter);
	      CK (cudaMalloc ((void **)&h -> d_sin_quarter, (Q + 1) * sizeof (float)));
	      CK (cudaMemcpy (h -> d_sin_quarter, q.data (), (Q + 1) * sizeof (float),
	                      cudaMemcpyHostToDevice));
	      L.q = h -> d_sin_quarter;
	   }
	}
	if (!h -> smem_lut_ok) {
	   h -> err = "SinCos table is not quarter-wave symmetric for this fm_rate";
	   return SDRJFM_ERR_UNSUPPORTED;
	}
	return consts_commit (h);
}

static int rebuild_tables (Lane *h) {
	h -> tables = build_tables (h -> cfg.input_rate, h -> cfg.fm_rate,
	                            h -> set.input_filter_hz, h -> set.lf_cutoff_hz);
	return upload_tables (h);
}

// allocates and zeroes the RDS state, uploads the band-pass spectrum, twiddles and decimator taps
static int rds_setup (Lane *h) {
const int64_t S = h -> cfg.n_streams;
const TableHeader &th = h -> tables.hdr ();
	if (!h -> d_rds_dring) {
	   CK (dalloc (&h -> d_rds_dring, (size_t)S * kRdsRing));
	   CK (dalloc (&h -> d_rds_pring, (size_t)S * kRdsRing));
	   CK (dalloc (&h -> d_rds_bp, (size_t)S * 2 * kRdsBlock));
	   CK (dalloc (&h -> d_rds_hi, (size_t)S * 2 * kRdsN));
	   CK (dalloc (&h -> d_rds_hist [0], (size_t)S * (kRdsDecTaps - 1)));
	   CK (dalloc (&h -> d_rds_hist [1], (size_t)S * (kRdsDecTaps - 1)));
	   CK (dalloc (&h -> d_rds_R, (size_t)kRdsNh + 1));
	   CK (dalloc (&h -> d_rds_tw, (size_t)kRdsNh));
	   CK (dalloc (&h -> d_rds_tws, (size_t)kRdsNh));
	   CK (dalloc (&h -> d_rds_dtaps, (size_t)kRdsDecTaps));
	   CK (cudaFuncSetAttribute (rds_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRdsFftSmem));
//	rdsBandPassFilter.setBand (RDS_FREQUENCY -+ RDS_WIDTH / 2, fmRate), fm-processor.cpp:166-168; Pass (float)
//	returns 3 * Re (conv): real taps 3 Re k[j].  R[k] = (1/Nh) sum_j 3 r[j] exp (-2 pi i j k / N), k = 0..Nh
	   std::vector<cf32> bp = design_bandpass (kRdsTaps, 57000 - 2400, 57000 + 2400, th.fm_rate);
	   std::vector<std::complex<double>> w (kRdsN);
	   for (int t = 0; t < kRdsN; t ++) w [t] = std::polar (1.0, -2 * M_PI * t / kRdsN);
	   std::vector<float2> R (kRdsNh + 1), tw (kRdsNh);
	   for (int k = 0; k <= kRdsNh; k ++) {
	      std::complex<double> a (0, 0);
	      for (int j = 0; j < kRdsTaps; j ++)
	         a += 3.0 * (double)bp [j].real () * w [(int)(((int64_t)j * k) & (kRdsN - 1))];
	      a /= (double)kRdsNh;
	      R [k] = make_float2 ((float)a.real (), (float)a.imag ());
	   }
	   for (int k = 0; k < kRdsNh; k ++) tw [k] = make_float2 ((float)w [k].real (), (float)w [k].imag ());
	   CK (cudaMemcpy (h -> d_rds_R, R.data (), R.size () * sizeof (float2), cudaMemcpyHostToDevice));
	   CK (cudaMemcpy (h -> d_rds_tw, tw.data (), tw.size () * sizeof (float2), cudaMemcpyHostToDevice));
	   std::vector<float2> tws (kRdsNh, make_float2 (1.f, 0.f));      // tws[half + pos] = exp (-2 pi i pos / (2 half))
	   for (int half = 1; half <= kRdsNh / 2; half <<= 1)
	      for (int pos = 0; pos < half; pos ++) {
	         const std::complex<double> v = w [(int)((int64_t)pos * (kRdsN / 2) / half) & (kRdsN - 1)];   // exp(-2 pi i pos/(2 half))
	         tws [half + pos] = make_float2 ((float)v.real (), (float)v.imag ());
	      }
	   CK (cudaMemcpy (h -> d_rds_tws, tws.data (), tws.size () * sizeof (float2), cudaMemcpyHostToDevice));
	   CK (cudaMemcpy (h -> d_rds_dtaps, h -> tables.payload () + th.off_rdsdecim,
	                   kRdsDecTaps * sizeof (float2), cudaMemcpyHostToDevice));
	}
	else {
	   CK (cudaMemsetAsync (h -> d_rds_dring, 0, (size_t)S * kRdsRing * sizeof (float), h -> stream));
	   CK (cudaMemsetAsync (h -> d_rds_pring, 0, (size_t)S * kRdsRing * sizeof (float), h -> stream));
	   CK (cudaMemsetAsync (h -> d_rds_hist [0], 0, (size_t)S * (kRdsDecTaps - 1) * sizeof (float2), h -> stream));
	   CK (cudaMemsetAsync (h -> d_rds_hist [1], 0, (size_t)S * (kRdsDecTaps - 1) * sizeof (float2), h -> stream));
	}
	h -> rds_total = 0; h -> rds_last_block = -1;
	return SDRJFM_OK;
}

static Lane *lane_create (const sdrjfm_config *cfg, int *status) {
int dummy; if (!status) status = &dummy;
	*status = SDRJFM_ERR_ARG;
	if (!cfg || cfg -> n_streams < 1 || cfg -> max_samples_per_call < 1) {
	   g_create_error = "bad config"; return nullptr;
	}
//	the decimation follows from the rates exactly as in the reference's constructor
//	(fm-processor.cpp:36,68-75): IRate = inputRate / 6, stage 1 /6, stage 2 / (IRate / fmRate)
int decim = 0, shape = -1;
int rsL = 0, rsM = 0, rsP = 0;
	if (cfg -> front_end_mode == 1) {
	   std::vector<float> ha, hb;
	   if (cfg -> fm_rate == 192000 && design_resampler (cfg -> input_rate, cfg -> fm_rate, rsL, rsM, rsP, ha, hb)) {
	      decim = kRsStageADecim; shape = kShapeResample;
	   }
	}
	else if (cfg -> front_end_mode != 0 && cfg -> front_end_mode != 2) { g_create_error = "front_end_mode must be 0, 1 or 2"; return nullptr; }
	else if (cfg -> fm_rate == 192000 && cfg -> input_rate >= 6 * cfg -> fm_rate) {
	   const int32_t irate = cfg -> input_rate / 6;
	   decim = (cfg -> input_rate / irate) * (irate / cfg -> fm_rate);
	   for (int i = 0; i < (int)(sizeof kFeShapes / sizeof kFeShapes [0]); i ++)
	      if (kFeShapes [i].D == decim && i != kShapeResample) shape = i;
//	   every other rate whose constructor arithmetic is well defined (stage 2 = D2 + 1 taps / D2, D2 = 1 .. 10)
	   if (shape < 0 && irate >= cfg -> fm_rate && cfg -> input_rate / irate == 6 && decim >= 6 && decim <= 60) shape = kShapeGeneric;
	}
	if (shape < 0) {
	   g_create_error = "unsupported rates: fm_rate must be 192000 and input_rate between 1152000 and 12670000 (front-end "
	                    "decimation 6 .. 60; 12, 30 and 48 have tuned kernels); the resampler "
	                    "mode needs 5 * 192000 / input_rate = L / M with L <= 16";
	   *status = SDRJFM_ERR_UNSUPPORTED; return nullptr;
	}
int ndev = 0;
	if (cfg -> device < 0 || cfg -> device >= kMaxDevices ||
	    cudaGetDeviceCount (&ndev) != cudaSuccess || ndev <= cfg -> device) {
	   g_create_error = "no CUDA device: this library has no CPU fallback";
	   *status = SDRJFM_ERR_NO_DEVICE; return nullptr;
	}
cudaDeviceProp prop;
	if (cudaSetDevice (cfg -> device) != cudaSuccess ||
	    cudaGetDeviceProperties (&prop, cfg -> device) != cudaSuccess || prop.major < 10) {
	   g_create_error = "device is not sm_100-class (B200)";
	   *status = SDRJFM_ERR_NO_DEVICE; return nullptr;
	}
Lane *h = new Lane ();
	h -> cfg = *cfg;
	h -> decim = decim; h -> shape = shape;
	h -> resample = shape == kShapeResample;
	h -> rsL = rsL; h -> rsM = rsM; h -> rsP = rsP; h -> rsHB = rsP - 1 + kRsHistPad;
	h -> ngw = fe_shape (shape, decim).ngw;
	h -> fw_delay = kInputFilterDelay / decim; h -> fw_shift = kInputFilterDelay % decim;
	{  const FeShape fs = fe_shape (shape, decim);
	   const int rows = fs.D * fs.gpt;
	   h -> hist_len   = ((fs.ng - 1 + fs.gpt - 1) / fs.gpt) * rows;      // Poly<>::HaloIn
	   h -> hist_len_w = std::max (((fs.ngw - 1 + fs.gpt - 1) / fs.gpt) * rows, fs.ngw * fs.D);
	   const char *env = getenv ("SDRJFM_GENERIC_FE"); h -> force_generic = env && env [0] == '1';
	   env = getenv ("SDRJFM_NO_TMA"); h -> use_tma = !(env && env [0] == '1');
	   env = getenv ("SDRJFM_TMA_CTAS"); h -> tma_ctas = env && atoi (env) > 0 ? atoi (env) : 0;
	   env = getenv ("SDRJFM_NO_AUTO_EXACT"); h -> auto_exact = !(env && env [0] == '1');
	   env = getenv ("SDRJFM_SEQ_DC"); h -> seq_dc = env && env [0] == '1';
//	   time slices of 6 pilot windows = 24576 fm samples (0.128 s: 16 PSS blocks, 12 audio tiles); 3 windows when few
//	   streams leave the SMs idle anyway (measured, 32 streams x 0.5 s: 1.60 / 1.48 / 1.61 ms at 2 / 3 / 6 windows)
	   env = getenv ("SDRJFM_FM_SLICE"); h -> slice_fm = env ? atoi (env) : (cfg -> n_streams <= 64 ? 3 : 6) * kPiWin;
	   if (h -> slice_fm % 4096) h -> slice_fm = 0; }
	if (h -> cfg.working_rate <= 0) h -> cfg.working_rate = 48000;
	if (h -> cfg.audio_rate <= 0) h -> cfg.audio_rate = h -> cfg.working_rate;
	h -> n_sm = prop.multiProcessorCount;
	if (h -> tma_ctas == 0) h -> tma_ctas = h -> n_sm;       // one persistent CTA per SM saturates HBM (measured)
	default_settings (h -> set, cfg -> fm_rate);
	h -> fade_max = h -> cfg.working_rate / 2;
	h -> fade_cnt = h -> fade_max;
const int64_t S = cfg -> n_streams;
	h -> cap_in    = ((cfg -> max_samples_per_call + decim + 15) / 16) * 16;
	h -> cap_fm    = ((h -> cap_in / decim + 1 + 15) / 16) * 16;
	if (h -> resample) {
	   h -> cap_a  = h -> cap_fm;
	   h -> cap_fm = ((h -> cap_a * rsL / rsM + 2 + 15) / 16) * 16;
	}
	h -> cap_audio = ((h -> cap_fm / kRsDecim + 1 + 15) / 16) * 16;
	h -> cap_rds   = ((h -> cap_fm / 8 + 1 + 15) / 16) * 16;
	h -> cap_out   = h -> cap_audio;
std::vector<float> cv_taps;
	if (h -> cfg.working_rate * kRsDecim != cfg -> fm_rate) {
	   g_create_error = "working_rate must be fm_rate / 4 (48000)"; *status = SDRJFM_ERR_UNSUPPORTED; delete h; return nullptr;
	}
	if (h -> cfg.audio_rate != h -> cfg.working_rate) {
//	   theConverter (workingRate, audioRate, workingRate / 20), fm-processor.cpp:89-91, 831-836
	   if (!convert_design (h -> cfg.working_rate, h -> cfg.audio_rate, h -> cvL, h -> cvM, cv_taps)) {
	      g_create_error = "unsupported audio_rate (audio_rate / working_rate must reduce to L / M with L <= 640)";
	      *status = SDRJFM_ERR_UNSUPPORTED; delete h; return nullptr;
	   }
	   h -> cap_out = ((h -> cap_audio * h -> cvL / h -> cvM + 2 + 15) / 16) * 16;
	}
	h -> peak_block = h -> cfg.working_rate / 50 + 1;          // peakLevelSampleMax = workingRate / 50, test `> max` (:142, :782)
auto fail = [&](cudaError_t e, const char *what) -> Lane * {
	   g_create_error = std::string (what) + ": " + cudaGetErrorString (e);
	   *status = SDRJFM_ERR_CUDA; lane_destroy (h); return nullptr;
	};
cudaError_t e;
#define AL(p, n) if ((e = dalloc (&h -> p, (size_t)(n))) != cudaSuccess) return fail (e, "cudaMalloc " #p)
	if ((e = cudaStreamCreateWithFlags (&h -> stream, cudaStreamNonBlocking)) != cudaSuccess ||
	    (e = cudaStreamCreateWithFlags (&h -> stream_rds, cudaStreamNonBlocking)) != cudaSuccess ||
	    (e = cudaStreamCreateWithFlags (&h -> stream_k3, cudaStreamNonBlocking)) != cudaSuccess ||
	    (e = cudaEventCreateWithFlags (&h -> ev_k2, cudaEventDisableTiming)) != cudaSuccess ||
	    (e = cudaEventCreateWithFlags (&h -> ev_k3, cudaEventDisableTiming)) != cudaSuccess ||
	    (e = cudaEventCreateWithFlags (&h -> ev_rds, cudaEventDisableTiming)) != cudaSuccess)
	   return fail (e, "cudaStreamCreate");
	AL (d_in, S * h -> cap_in);
	AL (d_hist [0], S * h -> hist_len); AL (d_hist [1], S * h -> hist_len);
	AL (d_pend, S * decim);
	if (h -> resample) {
	   AL (d_A, S * h -> cap_a); AL (d_SA, S * h -> cap_a);
	   for (int i = 0; i < 2; i ++) { AL (d_bha [i], S * h -> rsHB); AL (d_bhs [i], S * h -> rsHB); }
	}
	AL (d_U, S * h -> cap_fm); AL (d_S, S * h -> cap_fm);
	AL (d_iqn, S * h -> cap_fm); AL (d_fmz, S * h -> cap_fm);
	AL (d_res, S * h -> cap_fm); AL (d_zabs, S * h -> cap_fm);
	AL (d_demod, S * h -> cap_fm); AL (d_phase, S * h -> cap_fm); AL (d_pssd, S * h -> cap_fm);
	AL (d_locked, S * h -> cap_fm);
	AL (d_lr, S * h -> cap_fm); AL (d_a192, S * h -> cap_fm);
	AL (d_rdsc, S * h -> cap_fm); AL (d_rds24, S * h -> cap_rds);
	AL (d_ahist [0], S * kRsHist); AL (d_ahist [1], S * kRsHist);
	AL (d_audio, S * h -> cap_audio);
	AL (d_peak_ring, S * kPeakRing);
	if (!cv_taps.empty ()) {
	   AL (d_cv_taps, cv_taps.size ());
	   AL (d_cv_hist [0], S * kCvTapsPerPhase); AL (d_cv_hist [1], S * kCvTapsPerPhase);
	   if ((e = cudaMemcpy (h -> d_cv_taps, cv_taps.data (), cv_taps.size () * sizeof (float), cudaMemcpyHostToDevice)) != cudaSuccess)
	      return fail (e, "converter taps upload");
	}
	{  std::vector<float> tt;
	   tone_design (h -> cfg.working_rate, h -> tone_arm, tt);
	   h -> tone_burst = (int32_t)tt.size ();
	   AL (d_tone_tab, tt.size () + 1);
	   if ((e = cudaMemcpy (h -> d_tone_tab, tt.data (), tt.size () * sizeof (float), cudaMemcpyHostToDevice)) != cudaSuccess)
	      return fail (e, "tone table upload"); }
	AL (d_state, S);
	AL (d_iter_stats, S * 4);
	h -> ntiles_cap = (int32_t)((h -> cap_fm + kDiBlock - 1) / kDiBlock);
	AL (d_tileB, S * h -> ntiles_cap); AL (d_snap, S);
	AL (d_pss_ring, S * kPssRing);
#undef AL
	{  // initial member values of the reference objects
	   std::vector<StreamState> st (S);
	   memset (st.data (), 0, S * sizeof (StreamState));
	   for (auto &s : st) { s.Imin1 = s.Qmin1 = s.Imin2 = s.Qmin2 = 0.01f; }   // fm-demodulator.cpp:79-82
	   if ((e = cudaMemcpy (h -> d_state, st.data (), S * sizeof (StreamState),
	                        cudaMemcpyHostToDevice)) != cudaSuccess) return fail (e, "state upload");
	}
	if ((e = cudaFuncSetAttribute (frontend_fir_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                               kFeSmemBytes)) != cudaSuccess ||
	    (e = cudaFuncSetAttribute (frontend_fir_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                               kFeSmemBytes)) != cudaSuccess ||
	    (e = cudaFuncSetAttribute (frontend_tma_kernel<kDecim, kFeGpt, 37>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                               kFtSmemBytes)) != cudaSuccess ||
	    (e = cudaFuncSetAttribute (frontend_tma_kernel<48, 1, 73>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                               kFtSmemBytes)) != cudaSuccess) return fail (e, "smem attr K1");
	if ((e = poly_set_attr (shape)) != cudaSuccess) return fail (e, "smem attr K1g");
	if ((e = cudaFuncSetAttribute (frontend_exact_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fx<2>::SmemBytes)) != cudaSuccess ||
	    (e = cudaFuncSetAttribute (frontend_exact_kernel<5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fx<5>::SmemBytes)) != cudaSuccess ||
	    (e = cudaFuncSetAttribute (frontend_exact_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fx<8>::SmemBytes)) != cudaSuccess ||
	    (e = cudaFuncSetAttribute (frontend_exact_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fx<2>::SmemBytes)) != cudaSuccess ||
	    (e = cudaFuncSetAttribute (frontend_exact_kernel<5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fx<5>::SmemBytes)) != cudaSuccess ||
	    (e = cudaFuncSetAttribute (frontend_exact_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fx<8>::SmemBytes)) != cudaSuccess)
	   return fail (e, "smem attr K1x");
	{  const int seqsm = (cfg -> fm_rate / 4 + 1) * (int)sizeof (float);
	   const auto A = cudaFuncAttributeMaxDynamicSharedMemorySize;
	   if ((e = cudaFuncSetAttribute (sequential_kernel<0, false>, A, seqsm)) != cudaSuccess ||
	       (e = cudaFuncSetAttribute (sequential_kernel<1, false>, A, seqsm)) != cudaSuccess ||
	       (e = cudaFuncSetAttribute (sequential_kernel<2, false>, A, seqsm)) != cudaSuccess ||
	       (e = cudaFuncSetAttribute (sequential_kernel<0, true>, A, seqsm)) != cudaSuccess ||
	       (e = cudaFuncSetAttribute (sequential_kernel<1, true>, A, seqsm)) != cudaSuccess ||
	       (e = cudaFuncSetAttribute (sequential_kernel<2, true>, A, seqsm)) != cudaSuccess)
	      return fail (e, "smem attr K3"); }
	if ((e = cudaFuncSetAttribute (pilot_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                               (int)sizeof (PilotSmem))) != cudaSuccess ||
	    (e = cudaFuncSetAttribute (pilot_kernel<false, kPiThreadsWide>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                               (int)sizeof (PilotSmemT<kPiThreadsWide>))) != cudaSuccess) return fail (e, "smem attr pilot");
	{ const char *env = getenv ("SDRJFM_PILOT_WIDE");
//	  measured (32 / 64 / 128 streams x 0.5 s): 1.496 / 1.773 / 2.703 ms against 1.499 / 1.778 / 2.566 ms for the
//	  512-thread shape — one pass more per window and slower 32-warp scans eat what the wider pass gains: off by default
	  h -> pilot_wide = env && env [0] == '1'; }
	{ const char *env = getenv ("SDRJFM_SEQUENTIAL_PLL"); h -> sequential_pll = env && env [0] == '1'; }
int rc = rebuild_tables (h);
	if (rc != SDRJFM_OK) { g_create_error = h -> err; *status = rc; lane_destroy (h); return nullptr; }
	*status = SDRJFM_OK;
	return h;
}

static int lane_destroy (Lane *h) {
	if (!h) return SDRJFM_ERR_ARG;
	cudaSetDevice (h -> cfg.device);
	if (h -> stream) cudaStreamSynchronize (h -> stream);
void *ptrs [] = { h -> d_tables, h -> d_sin_quarter, h -> d_in, h -> d_hist [0], h -> d_hist [1],
	              h -> d_pend, h -> d_U, h -> d_S, h -> d_iqn, h -> d_fmz, h -> d_res, h -> d_zabs,
	              h -> d_demod, h -> d_phase, h -> d_pssd, h -> d_locked, h -> d_lr, h -> d_a192,
	              h -> d_rdsc, h -> d_rds24, h -> d_ahist [0], h -> d_ahist [1], h -> d_audio,
	              h -> d_state, h -> d_iter_stats, h -> d_pss_ring, h -> d_tileB, h -> d_snap,
	              h -> d_rds_dring, h -> d_rds_pring, h -> d_rds_bp, h -> d_rds_hi, h -> d_rds_R, h -> d_rds_tw,
	              h -> d_rds_dtaps, h -> d_rds_tws, h -> d_rds_hist [0], h -> d_rds_hist [1],
	              h -> d_histw [0], h -> d_histw [1], h -> d_Uw, h -> d_Sw, h -> d_udel [0], h -> d_udel [1],
	              h -> d_sdel [0], h -> d_sdel [1], h -> d_alp_hist [0], h -> d_alp_hist [1], h -> d_lrf, h -> d_lo_tab,
	              h -> d_A, h -> d_SA, h -> d_bha [0], h -> d_bha [1], h -> d_bhs [0], h -> d_bhs [1], h -> d_sq,
	              h -> d_air_int, h -> d_air_frac, h -> d_air_pend, h -> d_rsy_state, h -> d_rsy_bits, h -> d_rsy_nbits,
	              h -> d_rsy_c, h -> d_rsy_v, h -> d_rsy_w, h -> d_rsy_in,
	              h -> d_scan_carry [0], h -> d_scan_carry [1], h -> d_scan_db, h -> d_plot,
	              h -> d_spec_in, h -> d_spec_carry [0], h -> d_spec_carry [1], h -> d_spec_win,
	              h -> d_spec_Y, h -> d_spec_avg, h -> d_spec_disp, h -> d_xd, h -> d_xhist [0], h -> d_xhist [1], h -> d_dcnow,
	              h -> d_tone_tab, h -> d_peak_ring, h -> d_cv_taps, h -> d_cv_hist [0], h -> d_cv_hist [1],
	              h -> d_rs2_state, h -> d_rs2_m, h -> d_rs3_state, h -> d_rs3_sin, h -> d_rs3_groups, h -> d_rs3_ngroups, h -> d_rs3_stat, h -> d_hf_blk, h -> d_hf_win, h -> d_hf_avg, h -> d_hf_disp };
	for (void *p : ptrs) if (p) cudaFree (p);
	if (h -> stream_rds) { cudaStreamSynchronize (h -> stream_rds); cudaStreamDestroy (h -> stream_rds); }
	if (h -> stream_k3) { cudaStreamSynchronize (h -> stream_k3); cudaStreamDestroy (h -> stream_k3); }
	if (h -> ev_k2) cudaEventDestroy (h -> ev_k2);
	if (h -> ev_k3) cudaEventDestroy (h -> ev_k3);
	if (h -> ev_rds) cudaEventDestroy (h -> ev_rds);
	if (h -> ev_sym) cudaEventDestroy (h -> ev_sym);
	if (h -> stream) cudaStreamDestroy (h -> stream);
	delete h;
	return SDRJFM_OK;
}

static void *lane_cuda_stream (Lane *h) { return h ? (void *)h -> stream : nullptr; }
static int64_t lane_launch_count (const Lane *h) { return h ? h -> launches : 0; }

static int lane_pilot_stats (Lane *h, int32_t *out /* [n_streams][4] */) {
	if (!h || !out) return SDRJFM_ERR_ARG;
	CK (cudaMemcpyAsync (out, h -> d_iter_stats, (size_t)h -> cfg.n_streams * 4 * sizeof (int32_t),
	                     cudaMemcpyDeviceToHost, h -> stream));
	CK (cudaStreamSynchronize (h -> stream));
	return SDRJFM_OK;
}

static int lane_sync (Lane *h) {
	if (!h) return SDRJFM_ERR_ARG;
	CK (cudaStreamSynchronize (h -> stream));
	CK (cudaStreamSynchronize (h -> stream_rds));
	CK (cudaStreamSynchronize (h -> stream_k3));
	return SDRJFM_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*TmaEncodeFn) (CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmaEncodeFn tma_encoder () {
static TmaEncodeFn fn = [] () -> TmaEncodeFn {
	   void *p = nullptr;
	   cudaDriverEntryPointQueryResult q;
	   if (cudaGetDriverEntryPoint ("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
	       q != cudaDriverEntryPointSuccess) return nullptr;
	   return (TmaEncodeFn)p;
	} ();
	return fn;
}

// ---- K1g dispatch over the instantiated shapes ---------------------------------------------
template <int D, int GPT, int NG> static cudaError_t poly_attr_one () {
	return cudaFuncSetAttribute (frontend_poly_kernel<D, GPT, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                             Poly<D, GPT, NG>::SmemBytes + Poly<D, GPT, NG>::RawBytes);
}
static cudaError_t poly_set_attr (int shape) {
cudaError_t e;
	switch (shape) {
	   case kShapeGeneric: e = cudaSuccess; break;
	   case 0:  e = poly_attr_one<12, 4, 4> (); if (e == cudaSuccess) e = poly_attr_one<12, 4, 25> (); break;
	   case 1:  e = poly_attr_one<30, 2, 2> (); if (e == cudaSuccess) e = poly_attr_one<30, 2, 11> (); break;
	   case 2:  e = poly_attr_one<48, 1, 2> (); if (e == cudaSuccess) e = poly_attr_one<48, 1, 7> (); break;
	   default: e = poly_attr_one<kRsStageADecim, 4, 10> (); break;
	}
	return e;
}
template <int D, int GPT, int NG>
static void poly_launch (Lane *h, const void *src, int64_t pitch, RawFmt rf, const float2 *hist, int hist_len,
                         float2 *U, float2 *Sb, int32_t M, const LoParams &lp, int32_t tile0 = 0) {
typedef Poly<D, GPT, NG> P;
dim3 grid ((unsigned)((M + P::TileOut - 1) / P::TileOut - tile0), (unsigned)h -> cfg.n_streams);
	frontend_poly_kernel<D, GPT, NG><<<grid, kFeThreads, P::SmemBytes + (lp.tab ? P::RawBytes : 0), h -> stream>>> (
	      src, pitch, rf, hist, hist_len, U, Sb, h -> resample ? h -> cap_a : h -> cap_fm, M, lp, tile0);
}

// K1t over the whole tiles of a call (complex float, aligned rows, oscillator and inputFilter off);
// returns the number of tiles per stream it produced (0: not applicable), the caller runs the rest
template <int D, int GPT, int NT>
static int32_t tma_launch (Lane *h, const float2 *x, int64_t pitch, int32_t M, const float2 *hist, int hlen,
                           float2 *U, float2 *Sb) {
const int S = h -> cfg.n_streams;
const int32_t tileOut = kFeThreads * GPT;
	if (!h -> use_tma || M < tileOut || ((uintptr_t)x & 15) != 0 || (pitch & 1) != 0 || !tma_encoder ()) return 0;
const int32_t tiles = M / tileOut;
CUtensorMap map;
const cuuint64_t gdim [4] = { 32, 3, (cuuint64_t)tiles * kFtRows, (cuuint64_t)S };
const cuuint64_t gstr [3] = { 128, 384, (cuuint64_t)pitch * sizeof (float2) };
const cuuint32_t box [4] = { 32, 3, kFtBoxRows, 1 };
const cuuint32_t estr [4] = { 1, 1, 1, 1 };
	if (tma_encoder () (&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)x, gdim, gstr, box, estr,
	                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
	                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 0;
const int64_t total = (int64_t)tiles * S;
const unsigned g = (unsigned)std::min<int64_t> (total, (int64_t)h -> tma_ctas);
	frontend_tma_kernel<D, GPT, NT><<<g, kFeThreads, kFtSmemBytes, h -> stream>>> (map, hist, hlen, U, Sb, h -> cap_fm, tiles, S);
	h -> launches ++;
	return tiles;
}

// K1tb over the whole tiles of a call: device sample formats (and complex float at 6 MS/s) through a plain
// 3-D tensor map of 32-bit words; returns the tiles per stream it produced (0: not applicable)
template <int D, int GPT, int NT, int FMT>
static int32_t tmab_launch (Lane *h, const void *x, int64_t pitch, RawFmt rf, int32_t M, const float2 *hist, int hlen,
                            float2 *U, float2 *Sb) {
typedef Fb<D, GPT, NT, FMT> F;
const int S = h -> cfg.n_streams;
const int32_t tileOut = kFeThreads * GPT;
	if (!h -> use_tma || M < tileOut || ((uintptr_t)x & 15) != 0 || ((pitch * F::Bps) & 15) != 0 || !tma_encoder ()) return 0;
FbConv cv = { 0u, 0.f };
	if (FMT == kFmtU8) { cv.magic = 0x47800000u; cv.offset = 65536.0f + 127.0f / 128.0f; }
	else if (FMT == kFmtS8) { cv.magic = 0x47800000u; cv.offset = 65536.0f + 1.0f; }
	else if (FMT == kFmtS16) {
	   int k = 0; while (k < 16 && ldexpf (1.0f, -k) != rf.scale) k ++;
	   if (k < 8 || k >= 16) return 0;                              // denominators 256 .. 32768: the integer fits the mantissa field
	   cv.magic = (uint32_t)(150 - k) << 23; cv.offset = ldexpf (1.0f, 23 - k) + ldexpf (1.0f, 15 - k);
	}
const int32_t tiles = M / tileOut;
static bool attr_done = false;
	if (!attr_done) {
	   if (cudaFuncSetAttribute (frontend_tmab_kernel<D, GPT, NT, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SmemBytes) != cudaSuccess) return 0;
	   attr_done = true;
	}
CUtensorMap map;
const cuuint64_t gdim [3] = { (cuuint64_t)F::RowBytes / 4, (cuuint64_t)tiles * kFtRows, (cuuint64_t)S };
const cuuint64_t gstr [2] = { (cuuint64_t)F::RowBytes, (cuuint64_t)pitch * F::Bps };
const cuuint32_t box [3] = { (cuuint32_t)F::RowBytes / 4, (cuuint32_t)F::BoxRows, 1 };
const cuuint32_t estr [3] = { 1, 1, 1 };
	if (tma_encoder () (&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)x, gdim, gstr, box, estr,
	                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
	                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 0;
const int64_t total = (int64_t)tiles * S;
const int per_sm = F::SmemBytes * 2 <= 200 * 1024 ? 2 : 1;           // small stages: two persistent CTAs per SM
const unsigned g = (unsigned)std::min<int64_t> (total, (int64_t)h -> tma_ctas * per_sm);
	frontend_tmab_kernel<D, GPT, NT, FMT><<<g, kFeThreads, F::SmemBytes, h -> stream>>> (map, cv, hist, hlen, U, Sb,
	      h -> resample ? h -> cap_a : h -> cap_fm, tiles, S);
	h -> launches ++;
	return tiles;
}
// the three byte formats of one front-end shape
template <int D, int GPT, int NT>
static int32_t tmab_launch_fmt (Lane *h, const void *x, int64_t pitch, RawFmt rf, int32_t M, const float2 *hist, int hlen,
                                float2 *U, float2 *Sb) {
	switch (rf.fmt) {
	   case kFmtU8:  return tmab_launch<D, GPT, NT, kFmtU8>  (h, x, pitch, rf, M, hist, hlen, U, Sb);
	   case kFmtS8:  return tmab_launch<D, GPT, NT, kFmtS8>  (h, x, pitch, rf, M, hist, hlen, U, Sb);
	   case kFmtS16: return tmab_launch<D, GPT, NT, kFmtS16> (h, x, pitch, rf, M, hist, hlen, U, Sb);
	   default:      return 0;
	}
}

static int launch_frontend (Lane *h, const void *src, RawFmt rf, int64_t pitch, int32_t M) {
const int S = h -> cfg.n_streams;
	{ const int rc = consts_ensure (h); if (rc != SDRJFM_OK) return rc; }
LoParams lp;
	memset (&lp, 0, sizeof lp);
const bool lo = h -> set.lo_hz != 0;
	if (lo) {
	   lp.tab = h -> d_lo_tab; lp.rate = h -> cfg.input_rate; lp.lo = h -> set.lo_hz;
	   int64_t s128 = (128 * (int64_t)h -> set.lo_hz) % lp.rate; if (s128 < 0) s128 += lp.rate;
	   lp.step128 = (int32_t)s128; lp.phase = h -> lo_phase;
	   lp.lgain = h -> set.lgain; lp.rgain = h -> set.rgain;
	}
const bool wide = h -> set.input_filter_hz > 0;
const float2 *hist = wide ? h -> d_histw [h -> histw_sel] : h -> d_hist [h -> hist_sel];
const int hlen = wide ? h -> hist_len_w : h -> hist_len;
float2 *U = wide ? h -> d_Uw : h -> d_U, *Sb = wide ? h -> d_Sw : h -> d_S;
	if (h -> resample) { U = h -> d_A; Sb = h -> d_SA; }
	if (h -> decim == kDecim && rf.fmt == kFmtCF32 && !h -> force_generic) {
//	   the tuned kernels of the reference's own rate and sample format
	   const float2 *x = (const float2 *)src;
	   dim3 grid ((unsigned)((M + kFeTileOut - 1) / kFeTileOut), (unsigned)S);
	   if (wide) {
	      if (lo) frontend_wide_kernel<true><<<grid, kFeThreads, kFwSmemBytes, h -> stream>>> (
	            x, pitch, hist, hlen, U, Sb, h -> cap_fm, M, lp);
	      else    frontend_wide_kernel<false><<<grid, kFeThreads, kFwSmemBytes, h -> stream>>> (
	            x, pitch, hist, hlen, U, Sb, h -> cap_fm, M, lp);
	   }
	   else if (lo)
	      frontend_fir_kernel<true><<<grid, kFeThreads, kFeSmemBytes, h -> stream>>> (
	            x, pitch, hist, hlen, U, Sb, h -> cap_fm, M, lp, 0);
	   else {
//	      whole 512-output tiles through TMA (K1t), the ragged last tile through the plain kernel
	      const int32_t tiles = tma_launch<kDecim, kFeGpt, 37> (h, x, pitch, M, hist, hlen, U, Sb);
	      if (tiles > 0) {
	         if (M % kFeTileOut)
	            frontend_fir_kernel<false><<<dim3 (1, S), kFeThreads, kFeSmemBytes, h -> stream>>> (
	                  x, pitch, hist, hlen, U, Sb, h -> cap_fm, M, lp, tiles);
	         else h -> launches --;                          // (counted once below)
	      }
	      else frontend_fir_kernel<false><<<grid, kFeThreads, kFeSmemBytes, h -> stream>>> (
	            x, pitch, hist, hlen, U, Sb, h -> cap_fm, M, lp, 0);
	   }
	}
	else if (h -> shape == 2 && !wide && !lo && rf.fmt == kFmtCF32 && !h -> force_generic) {
//	   10 MS/s complex float: the same TMA kernel with one output per row (48 = D samples, 73 taps)
	   const int32_t tiles = tma_launch<48, 1, 73> (h, (const float2 *)src, pitch, M, hist, hlen, U, Sb);
	   if (tiles == 0) poly_launch<48, 1, 2> (h, src, pitch, rf, hist, hlen, U, Sb, M, lp);
	   else if (M % kFeThreads) poly_launch<48, 1, 2> (h, src, pitch, rf, hist, hlen, U, Sb, M, lp, tiles);
	   else h -> launches --;
	}
	else if (!wide && !lo && !h -> force_generic && rf.fmt != kFmtAirspy &&
	         h -> shape != kShapeResample && (rf.fmt != kFmtCF32 || h -> shape == 1)) {
//	   (stage A of the rational resampler stays on K1g: its 49 taps / 5 measured slower through TMA, 0.38 against 0.68)
//	   device sample formats (and complex float at 6 MS/s) through TMA: K1tb over the whole tiles, K1g for the ragged rest
	   int32_t tiles = 0;
	   if (h -> shape == 0) {
	      tiles = tmab_launch_fmt<12, 4, 37> (h, src, pitch, rf, M, hist, hlen, U, Sb);
	      if (tiles == 0) poly_launch<12, 4, 4> (h, src, pitch, rf, hist, hlen, U, Sb, M, lp);
	      else if (M % (kFeThreads * 4)) poly_launch<12, 4, 4> (h, src, pitch, rf, hist, hlen, U, Sb, M, lp, tiles);
	      else h -> launches --;
	   }
	   else if (h -> shape == 1) {
//	      rows of 60 samples: 480 / 240 bytes; the 8-bit formats (120-byte rows) stay on K1g
	      tiles = rf.fmt == kFmtCF32 ? tmab_launch<30, 2, 55, kFmtCF32> (h, src, pitch, rf, M, hist, hlen, U, Sb)
	            : rf.fmt == kFmtS16  ? tmab_launch<30, 2, 55, kFmtS16> (h, src, pitch, rf, M, hist, hlen, U, Sb) : 0;
	      if (tiles == 0) poly_launch<30, 2, 2> (h, src, pitch, rf, hist, hlen, U, Sb, M, lp);
	      else if (M % (kFeThreads * 2)) poly_launch<30, 2, 2> (h, src, pitch, rf, hist, hlen, U, Sb, M, lp, tiles);
	      else h -> launches --;
	   }
	   else {
	      tiles = tmab_launch_fmt<48, 1, 73> (h, src, pitch, rf, M, hist, hlen, U, Sb);
	      if (tiles == 0) poly_launch<48, 1, 2> (h, src, pitch, rf, hist, hlen, U, Sb, M, lp);
	      else if (M % kFeThreads) poly_launch<48, 1, 2> (h, src, pitch, rf, hist, hlen, U, Sb, M, lp, tiles);
	      else h -> launches --;
	   }
	}
	else switch (h -> shape * 2 + (wide ? 1 : 0)) {
	   case 0: poly_launch<12, 4, 4>  (h, src, pitch, rf, hist, hlen, U, Sb, M, lp); break;
	   case 1: poly_launch<12, 4, 25> (h, src, pitch, rf, hist, hlen, U, Sb, M, lp); break;
	   case 2: poly_launch<30, 2, 2>  (h, src, pitch, rf, hist, hlen, U, Sb, M, lp); break;
	   case 3: poly_launch<30, 2, 11> (h, src, pitch, rf, hist, hlen, U, Sb, M, lp); break;
	   case 4: poly_launch<48, 1, 2>  (h, src, pitch, rf, hist, hlen, U, Sb, M, lp); break;
	   case 5: poly_launch<48, 1, 7> (h, src, pitch, rf, hist, hlen, U, Sb, M, lp); break;
	   default: poly_launch<kRsStageADecim, 4, 10> (h, src, pitch, rf, hist, hlen, U, Sb, M, lp); break;
	}
	h -> launches ++;
	CK (cudaGetLastError ());
	return SDRJFM_OK;
}

// the reference-order front end is in force for this call?
static bool exact_wanted (const Lane *h) {
	if (h -> shape == kShapeGeneric) return true;       // rates without a composite kernel
	if (h -> resample || h -> set.input_filter_hz > 0 || h -> shape > 2) return false;   // inputFilter: an FFT filter sits between the oscillator and fmBand_1
	if (h -> cfg.front_end_mode == 2) return true;
	return h -> auto_exact && (h -> set.decoder == 2 || h -> set.decoder == 5 || h -> set.lo_hz != 0);
}

// inputFilter on (the composite has to stand in for the reference's FFT filter) but the decoder or the oscillator
// asks for the reference's own DC arithmetic: the per-sample DC remover of K1x runs in front of the wide composite
static bool hybrid_wanted (const Lane *h) {
	if (h -> resample || h -> set.input_filter_hz <= 0 || h -> shape > 2 || !h -> set.dc_remove) return false;
	if (h -> cfg.front_end_mode == 2) return true;
	return h -> auto_exact && (h -> set.decoder == 2 || h -> set.decoder == 5 || h -> set.lo_hz != 0);
}

// the per-sample RF DC remover: exact parallel-in-time solver (or the sequential walker as a cross-check)
static void launch_dc_exact (Lane *h, const void *src, int64_t pitch, RawFmt rf, int64_t n_proc, int write_state) {
const int S = h -> cfg.n_streams;
const float alpha = 1.0f / (float)h -> cfg.input_rate;                    // rfDcAlpha, fm-processor.cpp:379
	if (h -> seq_dc)
	   fx_dc_kernel<<<(S + 31) / 32, 32, 0, h -> stream>>> (src, pitch, rf, n_proc, S, alpha, h -> d_state, h -> d_xd, h -> cap_in, write_state);
	else if (rf.fmt == kFmtCF32)
	   fx_dc_par_kernel<true><<<S, kFdThreads, 0, h -> stream>>> (src, pitch, rf, n_proc, alpha, h -> d_state, h -> d_xd, h -> cap_in,
	                                                             write_state, getenv ("SDRJFM_DC_STATS") ? h -> d_iter_stats : nullptr);
	else
	   fx_dc_par_kernel<false><<<S, kFdThreads, 0, h -> stream>>> (src, pitch, rf, n_proc, alpha, h -> d_state, h -> d_xd, h -> cap_in,
	                                                              write_state, getenv ("SDRJFM_DC_STATS") ? h -> d_iter_stats : nullptr);
	h -> launches ++;
}

// K1x: per-sample DC removal, then the two decimators in the reference's operation order -> d_U = fm-rate samples
static int launch_frontend_exact (Lane *h, const void *src, RawFmt rf, int64_t pitch, int32_t M, int64_t n_proc, bool dry) {
const int S = h -> cfg.n_streams;
	{ const int rc = consts_ensure (h); if (rc != SDRJFM_OK) return rc; }
	if (!h -> d_xd) {
	   CK (dalloc (&h -> d_xd, (size_t)S * h -> cap_in));
	   CK (dalloc (&h -> d_xhist [0], (size_t)S * kFxHist)); CK (dalloc (&h -> d_xhist [1], (size_t)S * kFxHist));
	}
LoParams lp;
	memset (&lp, 0, sizeof lp);
	lp.lgain = h -> set.lgain; lp.rgain = h -> set.rgain;
	lp.rate = h -> cfg.input_rate; lp.phase = h -> lo_phase;
	if (h -> set.lo_hz != 0) { lp.tab = h -> d_lo_tab; lp.lo = h -> set.lo_hz; }
	if (!h -> fx_hist_valid && !dry) {
	   if (h -> in_total > 0) {
	      fx_seed_hist_kernel<<<S, kFxHist, 0, h -> stream>>> (h -> d_hist [h -> hist_sel], h -> hist_len, lp, h -> d_state,
	                                                         h -> set.dc_remove, h -> d_xhist [h -> xhist_sel]);
	      h -> launches ++;
	   }
	   else CK (cudaMemsetAsync (h -> d_xhist [h -> xhist_sel], 0, (size_t)S * kFxHist * sizeof (float2), h -> stream));
	   h -> fx_hist_valid = true;
	}
const void *fsrc = src; RawFmt frf = rf; int64_t fpitch = pitch;
	if (h -> set.dc_remove) {
	   launch_dc_exact (h, src, pitch, rf, n_proc, dry ? 0 : 1);
	   fsrc = h -> d_xd; fpitch = h -> cap_in;
	   memset (&frf, 0, sizeof frf); frf.fmt = kFmtCF32; frf.scale = 1.f;
	}
const dim3 grid ((unsigned)((M + kFxThreads - 1) / kFxThreads), (unsigned)S);
const float2 *xh = h -> d_xhist [h -> xhist_sel];
const bool plain = frf.fmt == kFmtCF32 && lp.tab == nullptr;
#define FX_LAUNCH(D2) do { if (plain) frontend_exact_kernel<D2, true><<<grid, kFxThreads, Fx<D2>::SmemBytes, h -> stream>>> (fsrc, fpitch, frf, xh, lp, h -> d_U, h -> cap_fm, M); \
	                      else frontend_exact_kernel<D2, false><<<grid, kFxThreads, Fx<D2>::SmemBytes, h -> stream>>> (fsrc, fpitch, frf, xh, lp, h -> d_U, h -> cap_fm, M); } while (0)
	switch (h -> shape == kShapeGeneric ? 0 : h -> decim / 6) {
	   case 2:  FX_LAUNCH (2); break;
	   case 5:  FX_LAUNCH (5); break;
	   case 8:  FX_LAUNCH (8); break;
	   default: frontend_exact_generic_kernel<<<grid, kFxThreads, 0, h -> stream>>> (fsrc, fpitch, frf, h -> decim / 6, xh, lp,
	                                                                             h -> d_U, h -> cap_fm, M); break;
	}
#undef FX_LAUNCH
	h -> launches ++;
	if (!dry) {
	   fx_roll_hist_kernel<<<S, kFxHist, 0, h -> stream>>> (fsrc, fpitch, frf, lp, xh, h -> d_xhist [h -> xhist_sel ^ 1], n_proc);
	   h -> xhist_sel ^= 1; h -> launches ++;
	}
	CK (cudaGetLastError ());
	return SDRJFM_OK;
}

// buffers of the wide (input filter ON) front end; cleared start
static int wide_setup (Lane *h) {
const int64_t S = h -> cfg.n_streams;
const size_t HL = (size_t)h -> hist_len_w, DL = (size_t)h -> fw_delay;
	if (!h -> d_Uw) {
	   CK (dalloc (&h -> d_histw [0], S * HL)); CK (dalloc (&h -> d_histw [1], S * HL));
	   CK (dalloc (&h -> d_Uw, (size_t)S * h -> cap_fm)); CK (dalloc (&h -> d_Sw, (size_t)S * h -> cap_fm));
	   for (int i = 0; i < 2; i ++) {
	      CK (dalloc (&h -> d_udel [i], S * DL)); CK (dalloc (&h -> d_sdel [i], S * DL));
	   }
	   CK (cudaFuncSetAttribute (frontend_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwSmemBytes));
	   CK (cudaFuncSetAttribute (frontend_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwSmemBytes));
	}
	else {
	   for (int i = 0; i < 2; i ++) {
	      CK (cudaMemsetAsync (h -> d_histw [i], 0, S * HL * sizeof (float2), h -> stream));
	      CK (cudaMemsetAsync (h -> d_udel [i], 0, S * DL * sizeof (float2), h -> stream));
	      CK (cudaMemsetAsync (h -> d_sdel [i], 0, S * DL * sizeof (float2), h -> stream));
	   }
	}
	return SDRJFM_OK;
}

static int lane_run_frontend_only (Lane *h, const void *d_iq, int32_t fmt, float scale, int64_t n_in, int64_t in_pitch) {
	if (!h || !d_iq || n_in < h -> decim || in_pitch < n_in) return SDRJFM_ERR_ARG;
	if (n_in > h -> cfg.max_samples_per_call) { h -> err = "n_in exceeds max_samples_per_call"; return SDRJFM_ERR_CAPACITY; }
	CK (cudaSetDevice (h -> cfg.device));
	if (fmt == kFmtAirspy) { h -> err = "front-end-only timing takes the rate-converted formats"; return SDRJFM_ERR_ARG; }
const RawFmt rf = make_rawfmt (h, fmt, scale);
	if (exact_wanted (h)) return launch_frontend_exact (h, d_iq, rf, in_pitch, (int32_t)(n_in / h -> decim), (n_in / h -> decim) * h -> decim, true);
	return launch_frontend (h, d_iq, rf, in_pitch, (int32_t)(n_in / h -> decim));
}

// LF scope display spectrum of the call (spectrum.cuh): the selected ELfPlot stream as complex samples,
// cut into spectrumSize blocks across call boundaries, transformed, mapped and averaged
static int run_lf_spectrum (Lane *h, int32_t M) {
const int S = h -> cfg.n_streams;
const int type = h -> lf_plot;
const bool rds_rate = type >= 8 && h -> set.rds_mode != 0;
const int32_t n = rds_rate ? (int32_t)h -> last_nrds : M;
	h -> last_nspec = 0;
	if (n <= 0) return SDRJFM_OK;
LfGather G = {};
	G.type = type; G.n = n; G.mul = 1.f;
	switch (type) {
	   case 1: G.c_src = h -> d_fmz; G.c_pitch = h -> cap_fm; break;
	   case 2: case 3: G.f_src = h -> d_demod; G.f_pitch = h -> cap_fm; break;
	   case 4: case 5: case 6: case 7: G.f_src = h -> d_plot; G.f_pitch = h -> cap_fm; break;
	   case 8: if (rds_rate) { G.c_src = h -> last_rds_ptr; G.c_pitch = h -> last_rds_pitch; G.mul = 20.f; } break;
	   case 9:
	      if (rds_rate) {
	         if (!h -> rds_symbols) { h -> err = "the RDS_DEMOD scope stream needs the RDS symbol stage (sdrjfm_set_rds_symbol_stage)"; return SDRJFM_ERR_UNSUPPORTED; }
	         CK (cudaStreamWaitEvent (h -> stream, h -> ev_sym, 0));
	         G.f_src = h -> d_rsy_c; G.f_pitch = h -> cap_rds; G.i_src = h -> d_plot; G.i_pitch = h -> cap_fm; G.mul = 4.f;
	      }
	      break;
	   default: break;                                                     // OFF: zeros
	}
	lf_gather_kernel<<<dim3 ((unsigned)((n + 255) / 256), (unsigned)S), 256, 0, h -> stream>>> (G, h -> d_spec_in, h -> cap_fm);
const int N = h -> spec_N;
const int32_t nblk = std::min ((h -> spec_carry + n) / N, h -> cap_specblk);
	if (nblk > 0) {
	   LfSpecParams P;
	   P.N = N; P.logN = h -> spec_logN; P.display = h -> spec_display; P.n_carry = h -> spec_carry;
	   int factor = (N / h -> spec_display) / 2;                           // mapSpectrum, ls-scope.cpp:134-145
	   factor = factor / h -> spec_zoom >= 1 ? factor / h -> spec_zoom : 1;
	   P.factor = factor;
	   P.full = (type == 1 || type == 8 || type == 9) ? 1 : 0;             // showFullSpectrum, fm-processor.cpp:247-263
	   lf_spectrum_kernel<<<dim3 ((unsigned)nblk, (unsigned)S), kSpecThreads, (size_t)N * sizeof (float2), h -> stream>>> (
	         h -> d_spec_in, h -> cap_fm, h -> d_spec_carry [h -> spec_sel], h -> d_spec_win, P, h -> d_spec_Y, h -> cap_specblk);
	   lf_average_kernel<<<dim3 ((unsigned)((h -> spec_display + 127) / 128), (unsigned)S), 128, 0, h -> stream>>> (
	         h -> d_spec_Y, h -> cap_specblk, nblk, h -> spec_display, h -> spec_avg_count, h -> spec_refresh ? 1 : 0,
	         h -> d_spec_avg, h -> d_spec_disp);
	   h -> spec_refresh = false;
	   h -> launches += 2;
	}
	lf_carry_kernel<<<S, 256, 0, h -> stream>>> (h -> d_spec_in, h -> cap_fm, n, N, h -> d_spec_carry [h -> spec_sel],
	                                             h -> spec_carry, h -> d_spec_carry [h -> spec_sel ^ 1]);
	h -> spec_sel ^= 1; h -> spec_carry = (h -> spec_carry + n) % N;
	h -> last_nspec = nblk;
	h -> launches += 2;
	return SDRJFM_OK;
}

// HF scope display spectrum of the call's raw samples (spectrum.cuh): of every segment of hf_seg input samples the
// first hf_N are gathered (across calls) and transformed when the segment is complete, as hs_scope::addElement does
static int run_hf_spectrum (Lane *h, const void *d_iq, RawFmt rf, int64_t pitch, int64_t n_in) {
const int S = h -> cfg.n_streams;
const int64_t T0 = h -> hf_total, T1 = T0 + n_in;
const int64_t seg = h -> hf_seg;
int nb = 0;
	for (int64_t k = T0 / seg; k * seg < T1; k ++) {
	   const int64_t lo = std::max (T0, k * seg), hi = std::min (T1, k * seg + h -> hf_N);
	   if (hi > lo) {
	      hf_gather_kernel<<<dim3 ((unsigned)((hi - lo + 255) / 256), (unsigned)S), 256, 0, h -> stream>>> (
	            d_iq, pitch, rf, lo - T0, (int32_t)(hi - lo), (int32_t)(lo - k * seg), h -> hf_N, h -> d_hf_blk);
	      h -> launches ++;
	   }
	   if (T1 >= (k + 1) * seg) {                                           // sampleCounter reached segmentSize (:109-110)
	      HfSpecParams P = { h -> hf_N, h -> hf_logN, h -> hf_display, h -> hf_half_freq };
	      hf_spectrum_kernel<<<S, kSpecThreads, (size_t)h -> hf_N * sizeof (float2), h -> stream>>> (
	            h -> d_hf_blk, h -> d_hf_win, P, h -> d_hf_avg, h -> d_hf_disp);
	      h -> launches ++; nb ++;
	   }
	}
	h -> hf_total = T1; h -> last_nhf = nb;
	CK (cudaGetLastError ());
	return SDRJFM_OK;
}

// the launch sequence behind both process entry points; `src` is a device pointer holding
// (pending | new) samples contiguously per stream with row pitch `pitch`
static int run_chain (Lane *h, const void *src, RawFmt rf, int64_t pitch, int64_t n_proc,
                      float2 *d_audio_out, int64_t audio_pitch, int64_t *n_audio,
                      float2 *d_rds_out, int64_t rds_pitch, int64_t *n_rds) {
const int S = h -> cfg.n_streams;
const Settings &st = h -> set;
const TableHeader &th = h -> tables.hdr ();
const float *T = h -> d_tables;
const int32_t M1 = (int32_t)(n_proc / h -> decim);      // front-end (stage A) outputs of this call
int32_t M = M1;
	if (h -> resample)      // fm samples that exist once stage-A sample a_total + M1 - 1 does (resample.cuh)
	   M = (int32_t)(((h -> a_total + M1) * h -> rsL + h -> rsM - 1) / h -> rsM - h -> fm_total);
	h -> last_nfm = M; h -> last_naudio = 0; h -> last_nrds = 0; h -> last_nspec = 0;
	if (n_audio) *n_audio = 0;
	if (n_rds) *n_rds = 0;
	if (M1 == 0) return SDRJFM_OK;
int rc;
//	K1 ------------------------------------------------------------------------------------
const bool wide = st.input_filter_hz > 0;
const bool exact = exact_wanted (h);
const bool hybrid = !exact && hybrid_wanted (h);
	if (wide && hybrid != h -> hybrid_last && h -> in_total > 0) {
//	   the wide history holds raw samples after a plain call, DC-free ones after a hybrid call
	   hist_shift_dc_kernel<<<dim3 ((h -> hist_len_w + 127) / 128, S), 128, 0, h -> stream>>> (
	         h -> d_histw [h -> histw_sel], h -> hist_len_w, h -> d_state, hybrid ? -1.0f : 1.0f);
	   h -> launches ++;
	}
	h -> hybrid_last = wide && hybrid;
	if (hybrid) {
	   if (!h -> d_xd) {
	      CK (dalloc (&h -> d_xd, (size_t)S * h -> cap_in));
	      CK (dalloc (&h -> d_xhist [0], (size_t)S * kFxHist)); CK (dalloc (&h -> d_xhist [1], (size_t)S * kFxHist));
	   }
	   launch_dc_exact (h, src, pitch, rf, n_proc, 1);
	   src = h -> d_xd; pitch = h -> cap_in;
	   memset (&rf, 0, sizeof rf); rf.fmt = kFmtCF32; rf.scale = 1.f;
	}
	if (exact) rc = launch_frontend_exact (h, src, rf, pitch, M1, n_proc, false);
	else { rc = launch_frontend (h, src, rf, pitch, M1); h -> fx_hist_valid = false; }
	if (rc != SDRJFM_OK) return rc;
	h -> in_total += n_proc;
	if (wide) {
	   roll_history_raw_kernel<<<dim3 ((h -> hist_len_w + 159) / 160, S), 160, 0, h -> stream>>> (
	         src, pitch, rf, h -> d_histw [h -> histw_sel], h -> d_histw [h -> histw_sel ^ 1], n_proc, h -> hist_len_w);
	   h -> histw_sel ^= 1;
	   const int DL = h -> fw_delay;
	   dim3 g ((unsigned)((std::max (M, DL) + 255) / 256), (unsigned)S);
	   fm_delay_kernel<<<g, 256, 0, h -> stream>>> (h -> d_Uw, h -> cap_fm, M, DL, h -> d_udel [h -> del_sel],
	                                                h -> d_udel [h -> del_sel ^ 1], h -> d_U);
	   fm_delay_kernel<<<g, 256, 0, h -> stream>>> (h -> d_Sw, h -> cap_fm, M, DL, h -> d_sdel [h -> del_sel],
	                                                h -> d_sdel [h -> del_sel ^ 1], h -> d_S);
	   h -> del_sel ^= 1;
	   h -> launches += 3;
	}
	else {
	   roll_history_raw_kernel<<<dim3 ((h -> hist_len + 63) / 64, S), 64, 0, h -> stream>>> (
	         src, pitch, rf, h -> d_hist [h -> hist_sel], h -> d_hist [h -> hist_sel ^ 1], n_proc, h -> hist_len);
	   h -> hist_sel ^= 1; h -> launches ++;
	}
	if (h -> resample) {
//	   stage B: rational L / M from the stage-A rate to the fm rate
	   if (M > 0) {
	      ResampleParams q;
	      q.L = h -> rsL; q.MB = h -> rsM; q.P = h -> rsP; q.HB = h -> rsHB;
	      q.a0 = h -> a_total; q.m0 = h -> fm_total; q.M = M;
//	      group delay of the two symmetric filters in stage-A samples (24 input samples + the prototype's
//	      centre): the DC estimate subtracted from fm sample m is the one at the CENTRE of its window
	      q.dA = (int32_t)lround ((kRsStageATaps - 1) / 2.0 / kRsStageADecim + (h -> rsL * h -> rsP - 1) / 2.0 / h -> rsL);
	      dim3 g ((unsigned)((M + 255) / 256), (unsigned)S);
	      resample_b_kernel<<<g, 256, 0, h -> stream>>> (h -> d_A, h -> d_SA, h -> cap_a, h -> d_bha [h -> bh_sel],
	            h -> d_bhs [h -> bh_sel], T + th.off_rsB, q, h -> d_U, h -> d_S, h -> cap_fm);
	      h -> launches ++;
	   }
	   dim3 gr ((unsigned)((h -> rsHB + 63) / 64), (unsigned)S);
	   roll_history_kernel<<<gr, 64, 0, h -> stream>>> (h -> d_A, h -> cap_a, h -> d_bha [h -> bh_sel],
	                                                   h -> d_bha [h -> bh_sel ^ 1], M1, h -> rsHB);
	   roll_history_kernel<<<gr, 64, 0, h -> stream>>> (h -> d_SA, h -> cap_a, h -> d_bhs [h -> bh_sel],
	                                                   h -> d_bhs [h -> bh_sel ^ 1], M1, h -> rsHB);
	   h -> bh_sel ^= 1; h -> launches += 2;
	   h -> a_total += M1;
	   if (M == 0) { CK (cudaGetLastError ()); return SDRJFM_OK; }
	}
//	K2 ------------------------------------------------------------------------------------
const float *consts = h -> tables.payload () + th.off_comp_consts;
DiscrParams dp;
	dp.sumC = consts [0]; dp.sumCm = consts [1];
	dp.gb0 = consts [5]; dp.gb1 = consts [6]; dp.gb2 = consts [7];
	if (wide) { dp.sumC = h -> wide_sumC; dp.sumCm = h -> wide_sumCm; dp.gb0 = dp.gb1 = dp.gb2 = 0.f; }
	dp.lo_tab = nullptr; dp.lo_rate = h -> cfg.input_rate; dp.lo_hz = st.lo_hz; dp.lo_moff = wide ? -h -> fw_delay : 0;
	dp.decim = h -> decim;
	dp.lo_phase = h -> lo_phase; dp.Hre = h -> lo_Hre; dp.Him = h -> lo_Him;
	if (st.lo_hz != 0) {
	   if (!exact) dp.lo_tab = h -> d_lo_tab;
	   int64_t np = (h -> lo_phase - (int64_t)st.lo_hz * n_proc) % h -> cfg.input_rate;
	   h -> lo_phase = np < 0 ? np + h -> cfg.input_rate : np;       // LOPhase after this call's samples
	}
	dp.Gre = consts [2]; dp.Gim = consts [3];
	dp.alpha = (double)(1.0f / h -> cfg.input_rate);          // rfDcAlpha, fm-processor.cpp:379
	dp.beta = pow (1.0 - dp.alpha, (double)h -> decim);
	if (h -> resample) {
//	   unit-gain real cascade (every polyphase branch normalised), no DC folding; the estimate advances
//	   by inputRate / fmRate input samples per fm sample on average
	   dp.sumC = dp.sumCm = 1.f; dp.gb0 = dp.gb1 = dp.gb2 = 0.f; dp.Gre = 1.f; dp.Gim = 0.f;
	   dp.beta = pow (1.0 - dp.alpha, (double)h -> cfg.input_rate / (double)h -> cfg.fm_rate);
	}
	dp.lgain = st.lgain; dp.rgain = st.rgain;
	dp.dc_remove = st.dc_remove; dp.decoder = st.decoder;
	dp.scan_only = h -> scanning;
	dp.exact = exact;
	if (hybrid) dp.dc_remove = 0;      // the DC went out sample by sample in front of the composite
	if (exact) {      // K1x delivered the reference's fm-rate samples: K2 only normalises and discriminates
	   dp.dc_remove = 0; dp.sumC = dp.sumCm = 0.f; dp.gb0 = dp.gb1 = dp.gb2 = 0.f;
	   dp.lgain = dp.rgain = 1.f; dp.Gre = 1.f; dp.Gim = 0.f;
	}
const int32_t ntiles = (M + kDiBlock - 1) / kDiBlock;
	{
	   dim3 g ((unsigned)ntiles, (unsigned)S);
	   dc_tile_kernel<<<g, kDiThreads, 0, h -> stream>>> (h -> d_S, h -> cap_fm, M, dp, h -> d_state,
	                                                       h -> d_tileB, ntiles, h -> d_snap);
	   const dim3 gd ((unsigned)((ntiles + kDiTilesPerCta - 1) / kDiTilesPerCta), (unsigned)S);
	   discriminator_kernel<<<gd, kDiThreads, 0, h -> stream>>> (
	         h -> d_U, h -> d_S, h -> cap_fm, M, dp, T + th.off_atan, T + th.off_arcsine,
	         h -> d_state, h -> d_tileB, ntiles, h -> d_snap, h -> d_res, h -> d_zabs,
	         (st.decoder == 2 || st.decoder == 1) ? h -> d_iqn : nullptr,
	         (h -> cfg.keep_taps || h -> scanning || h -> lf_plot == 1) ? h -> d_fmz : nullptr);
	   h -> launches += 2;
	}
	if (h -> scanning) {
//	   fm-processor.cpp:478-495: while scanning nothing behind the decimators runs ("continue")
	   const int32_t nblk = (h -> scan_carry + M) / kScanN;
	   h -> last_nscan = std::min (nblk, h -> cap_scan);
	   if (h -> last_nscan > 0)
	      scan_kernel<<<dim3 ((unsigned)h -> last_nscan, (unsigned)S), kScanThreads, 0, h -> stream>>> (
	            h -> d_fmz, h -> cap_fm, h -> d_scan_carry [h -> scan_sel], h -> scan_carry, h -> d_scan_db, h -> cap_scan);
	   scan_carry_kernel<<<S, 256, 0, h -> stream>>> (h -> d_fmz, h -> cap_fm, M, h -> d_scan_carry [h -> scan_sel],
	                                                  h -> scan_carry, h -> d_scan_carry [h -> scan_sel ^ 1]);
	   h -> scan_sel ^= 1; h -> scan_carry = (h -> scan_carry + M) % kScanN;
	   h -> launches += 2;
	   CK (cudaGetLastError ());
	   return SDRJFM_OK;
	}
//	K3 .. K6 run per TIME SLICE of the call: the per-stream recurrences (pilot solver, PSS loop) walk a stream in
//	order, so inside one call K3 of slice j + 1 (on its own CUDA stream) runs beside K4 / K5 / K6 of slice j — the
//	critical path of a call becomes max (K3, K4) instead of their sum.  A slice is a whole number of pilot windows,
//	PSS blocks and audio tiles; stream state is carried exactly as it is between calls, so slicing is invisible in
//	the results.  The per-call side outputs (scope stream, symbol stage) keep the whole call in one piece.
const int32_t Mcall = M;
const bool may_slice = h -> slice_fm > 0 && !h -> rds_symbols && h -> lf_plot < 0 && h -> spec_N == 0 && Mcall >= 2 * h -> slice_fm;
const int32_t nsl = may_slice ? (Mcall + h -> slice_fm - 1) / h -> slice_fm : 1;
	if (nsl > 1) {
	   CK (cudaEventRecord (h -> ev_k2, h -> stream));
	   CK (cudaStreamWaitEvent (h -> stream_k3, h -> ev_k2, 0));
	}
cudaStream_t ks = nsl > 1 ? h -> stream_k3 : h -> stream;
int64_t na_tot = 0, nr_tot = 0, nq_tot = 0;
	for (int32_t sl = 0; sl < nsl; sl ++) {
const int64_t o = may_slice ? (int64_t)sl * h -> slice_fm : 0;
const int32_t M = may_slice ? (int32_t)std::min<int64_t> (h -> slice_fm, Mcall - o) : Mcall;
SeqParams sp;
	sp.K_FM = consts [4];
	sp.omega = (float)(((float)19000 / h -> cfg.fm_rate) * (2 * M_PI));   // OMEGA_PILOT, :34
	sp.gain = (float)(10 * (2 * M_PI) / h -> cfg.fm_rate);                // :79
	sp.lock_half_rate = h -> cfg.fm_rate >> 1;
	sp.decoder = st.decoder;
	{  // pllC ctor, fm-demodulator.cpp:66-72 / pllC.cpp:38-58
	   const float maxdev = 0.95 * (0.5 * h -> cfg.fm_rate);
	   const float fac = 2.0 * M_PI / h -> cfg.fm_rate;
	   const float bw = 0.85 * h -> cfg.fm_rate;
	   sp.pll_beta = exp (-2.0 * M_PI * bw / 2 / h -> cfg.fm_rate);
	   sp.pll_lo = -maxdev * fac; sp.pll_hi = maxdev * fac;
	   sp.pll_reset = 0.0f;
	}
	sp.n_streams = S;
const int seq_blocks = (S + kSeqLanes - 1) / kSeqLanes;
const size_t seq_smem = (h -> lut.quarter + 1) * sizeof (float);
//	the squelch (20th-order IIRs per sample), the PLL and the AM decoder are per-sample float
//	recurrences: lane per stream.  Everything else: the parallel-in-time pilot kernel.
SquelchParams qp;
	memset (&qp, 0, sizeof qp);
	if (st.squelch_mode != 0) {
	   qp.mode = st.squelch_mode;
	   qp.hold = h -> cfg.fm_rate / 20;                              // holdPeriod, fm-processor.cpp:87
	   qp.thr_level = std::pow (10.0f, (st.squelch_value - 80) / 30.0f);   // setSquelchLevel, squelchClass.cpp:34-38
	   qp.thr_noise = 1.0f - st.squelch_value / 100.0f;
	   qp.weight = (float)(h -> cfg.fm_rate / 100);
	   const float *sq = h -> tables.payload () + th.off_squelch;
	   memcpy (qp.hp, sq, sizeof qp.hp);
	   memcpy (qp.lp, sq + 1 + 4 * kSqQuads, sizeof qp.lp);
	}
const int dec = st.decoder == 2 ? 1 : st.decoder == 1 ? 2 : 0;
#define SEQ_LAUNCH(D, Q) sequential_kernel<D, Q><<<seq_blocks, kSeqLanes, seq_smem, ks>>> ( \
	         (h -> d_res + o), (h -> d_zabs + o), (h -> d_iqn + o), h -> cap_fm, M, sp, h -> lut, T + th.off_atan, \
	         h -> d_state, (h -> d_demod + o), (h -> d_phase + o), (h -> d_locked + o), qp, h -> d_sq)
	if (st.squelch_mode != 0) {
	   if (dec == 1) SEQ_LAUNCH (1, true); else if (dec == 2) SEQ_LAUNCH (2, true); else SEQ_LAUNCH (0, true);
	}
	else if (dec == 1) SEQ_LAUNCH (1, false);
	else if (dec == 2) SEQ_LAUNCH (2, false);
	else if (h -> sequential_pll) SEQ_LAUNCH (0, false);
#undef SEQ_LAUNCH
	else {
	   PilotParams pp;
	   pp.K_FM = sp.K_FM; pp.omega = sp.omega; pp.gain = sp.gain;
	   pp.lock_half_rate = sp.lock_half_rate; pp.n_streams = S;
	   if (h -> pilot_wide)
	      pilot_kernel<false, kPiThreadsWide><<<S, kPiThreadsWide, sizeof (PilotSmemT<kPiThreadsWide>), ks>>> (
	            (h -> d_res + o), (h -> d_zabs + o), h -> cap_fm, M, pp, h -> lut, h -> d_state,
	            (h -> d_demod + o), (h -> d_phase + o), (h -> d_locked + o), h -> d_iter_stats, sl > 0 ? 1 : 0);
	   else
	      pilot_kernel<false><<<S, kPiThreads, sizeof (PilotSmem), ks>>> (
	            (h -> d_res + o), (h -> d_zabs + o), h -> cap_fm, M, pp, h -> lut, h -> d_state,
	            (h -> d_demod + o), (h -> d_phase + o), (h -> d_locked + o), h -> d_iter_stats, sl > 0 ? 1 : 0);
	}
	h -> launches ++;
	if (st.rds_mode != 0 || nsl > 1) CK (cudaEventRecord (h -> ev_k3, ks));
	if (nsl > 1) CK (cudaStreamWaitEvent (h -> stream, h -> ev_k3, 0));
//	K4 ------------------------------------------------------------------------------------
	{
	   StereoParams q;
	   q.fm_mode = st.fm_mode; q.auto_mono = st.auto_mono; q.pss_on = st.pss_on;
	   q.sound_sel = st.sound_sel; q.panorama = st.panorama;
	   q.pss_alpha = 10.0f / h -> cfg.fm_rate;                 // fm-processor.cpp:81-82
	   q.pss_lock_alpha = 1.0f / h -> cfg.fm_rate;             // stereo-separation.cpp:32
	   q.rate3 = 3 * h -> cfg.fm_rate;
	   q.write_pss_tap = h -> cfg.keep_taps;
	   stereo_kernel<<<S, kStThreads, 0, h -> stream>>> (
	         (h -> d_demod + o), (h -> d_phase + o), (h -> d_locked + o), h -> cap_fm, M, q,
	         reinterpret_cast<const float2 *>(T + th.off_sincos), h -> d_state, h -> d_pss_ring,
	         (h -> d_lr + o), (h -> d_pssd + o), h -> lf_plot == 4 ? h -> d_plot + o : nullptr);
	   h -> launches ++;
	}
//	K5 ------------------------------------------------------------------------------------
	if (st.rds_mode != 0) {
//	   fork: the RDS branch on its own stream, ordered after K3 (it reads demod and the pilot phase)
	   cudaStream_t rs = h -> stream_rds;
	   CK (cudaStreamWaitEvent (rs, h -> ev_k3, 0));
	   const int64_t n0 = h -> rds_total;
	   for (int32_t m = 0; m < M; ) {
	      const int64_t K = (n0 + m) / kRdsBlock;
	      const int32_t mEnd = (int32_t)std::min<int64_t> (M, (K + 1) * kRdsBlock - n0);
	      dim3 g ((unsigned)((mEnd - m + 255) / 256), (unsigned)S);
	      rds_append_kernel<<<g, 256, 0, rs>>> ((h -> d_demod + o), (h -> d_phase + o), h -> cap_fm, m, mEnd, n0,
	                                                    h -> d_rds_dring, h -> d_rds_pring);
	      h -> launches ++;
	      if (K >= 1 && h -> rds_last_block < K - 1) {
	         rds_block_kernel<<<S, kRdsThreads, kRdsFftSmem, rs>>> (
	               h -> d_rds_dring, K - 1, h -> d_rds_tw, h -> d_rds_tws, h -> d_rds_R, h -> d_rds_bp, h -> d_rds_hi);
	         h -> launches ++;
	         h -> rds_last_block = K - 1;
	      }
	      rds_mix_kernel<<<g, 256, 0, rs>>> (h -> d_rds_pring, h -> d_rds_bp, h -> d_rds_hi,
	                                                 h -> cap_fm, m, mEnd, n0, (h -> d_rdsc + o));
	      h -> launches ++;
	      m = mEnd;
	   }
	   const int32_t nout = (int32_t)((n0 + M) / kRdsDecim - n0 / kRdsDecim);
	   float2 *rout = (d_rds_out ? d_rds_out : h -> d_rds24) + nr_tot;
	   const int64_t rpitch = d_rds_out ? rds_pitch : h -> cap_rds;
	   dim3 g ((unsigned)((std::max (nout, 1) + 127) / 128), (unsigned)S);
	   rds_decim_kernel<<<g, 128, 0, rs>>> ((h -> d_rdsc + o), h -> cap_fm, M, n0, h -> d_rds_dtaps,
	         h -> d_rds_hist [h -> rds_hist_sel], h -> d_rds_hist [h -> rds_hist_sel ^ 1], rout, rpitch, nout,
	         h -> rds_symbols ? h -> d_rsy_in : nullptr, h -> cap_rds);
//	   join point of the 24 kHz baseband; the symbol stage below keeps running on the side stream,
//	   beside the NEXT call's front end and pilot stage (its bits are waited for where they are read)
	   CK (cudaEventRecord (h -> ev_rds, rs));
	   h -> launches ++;
	   h -> rds_hist_sel ^= 1;
	   h -> rds_total += M;
	   nr_tot += nout;
	   h -> last_nrds = nr_tot;
	   h -> last_rds_ptr = rout - (nr_tot - nout); h -> last_rds_pitch = rpitch;
	   if (n_rds) *n_rds = nr_tot;
	   if (h -> rds_symbols && nout > 0 && st.rds_mode == 2) {
//	      symbol stage, mode RDS_2 (rds-decoder.cpp:84-88): rdsDecoder_2 (matched filter, AGC, M&M timing, Costas)
	      Rds2Params p2;
	      design_rds2_matched_filter (24000, p2.taps);
	      p2.agc_rate = 2e-3f; p2.agc_ref = 0.38f;                             // my_AGC (2e-3f, 0.38f, 9.0f), rds-decoder-2.cpp:46
	      p2.sps = 24000 / (float)1187.5f; p2.mm_alpha = 0.01;                  // :52, :59
	      p2.c_alpha = 1.0f; p2.c_beta = 0.02f; p2.freq_limit = 2 * M_PI * 10.0f / 24000.0f;     // my_Costas (rate, 1.0f, 0.02f, 10.0f)
	      const int64_t bp = h -> cap_rds;
	      rds2_match_kernel<<<dim3 ((unsigned)((nout + 127) / 128), (unsigned)S), 128, 0, rs>>> (
	            h -> d_rsy_in, bp, nout, p2, h -> d_rs2_state, h -> d_rs2_m);
	      rds2_seq_kernel<<<(S + kRsyLanes - 1) / kRsyLanes, kRsyLanes, 0, rs>>> (
	            h -> d_rsy_in, h -> d_rs2_m, bp, nout, S, p2, h -> d_rs2_state, h -> d_rsy_bits, h -> cap_bits, h -> d_rsy_nbits);
	      h -> launches += 2;
	   }
	   else if (h -> rds_symbols && nout > 0 && st.rds_mode == 3) {
//	      symbol stage, mode RDS_3 (rds-decoder.cpp:90-98): the Costas loop of mode 1, then rdsDecoder_3 with the block
//	      synchroniser in its loop
	      RdsSymParams sp2;
	      sp2.alpha = 1.0f / 16.0f; sp2.beta = 0.02f / 16.0f;
	      sp2.freq_limit = 2 * M_PI * 10.0f / 24000.0f;
	      const float *tab = h -> tables.payload () + th.off_rds_sym;
	      memcpy (sp2.match, tab, sizeof sp2.match);
	      memcpy (sp2.lp, tab + kRsyMatch, sizeof sp2.lp);
	      memcpy (sp2.bp, tab + kRsyMatch + kRsyLp, sizeof sp2.bp);
	      Rds3Params p3;
	      p3.sin_tab = h -> d_rs3_sin; p3.C = 24000 / (2 * M_PI); p3.rate = 24000; p3.omega = h -> rs3_omega;
	      memcpy (p3.kmap, h -> rs3_kmap, sizeof p3.kmap);
	      const int64_t bp = h -> cap_rds;
	      rds_costas_kernel<<<(S + kRsyLanes - 1) / kRsyLanes, kRsyLanes, 0, rs>>> (
	            h -> d_rsy_in, bp, nout, S, sp2, h -> d_rsy_state, h -> d_rsy_c, bp,
	            h -> lf_plot == 9 ? h -> d_plot : nullptr, h -> cap_fm);
	      if (h -> lf_plot == 9 && h -> spec_N) CK (cudaEventRecord (h -> ev_sym, rs));
	      const dim3 gf ((unsigned)((nout + 127) / 128), (unsigned)S);
	      rds_fir_kernel<kRsyLp, false><<<gf, 128, 0, rs>>> (h -> d_rsy_c, h -> d_rsy_v, bp, nout, sp2, h -> d_rsy_state);
	      rds3_seq_kernel<<<(S + kRsyLanes - 1) / kRsyLanes, kRsyLanes, 0, rs>>> (
	            h -> d_rsy_c, h -> d_rsy_v, bp, nout, S, p3, h -> d_rsy_state, h -> d_rs3_state,
	            h -> d_rsy_bits, h -> cap_bits, h -> d_rsy_nbits, h -> d_rs3_groups, h -> cap_groups, h -> d_rs3_ngroups, h -> d_rs3_stat);
	      rds_sym_roll_kernel<<<S, 64, 0, rs>>> (h -> d_rsy_c, h -> d_rsy_v, bp, nout, h -> d_rsy_state);
	      h -> launches += 4;
	   }
	   else if (h -> rds_symbols && nout > 0) {
//	      symbol stage, mode RDS_1 (rds-decoder.cpp:69-82): Costas (rate, 1/16, 0.02/16, 10 Hz) + decoder 1
	      RdsSymParams sp2;
	      sp2.alpha = 1.0f / 16.0f; sp2.beta = 0.02f / 16.0f;
	      sp2.freq_limit = 2 * M_PI * 10.0f / 24000.0f;
	      const float *tab = h -> tables.payload () + th.off_rds_sym;
	      memcpy (sp2.match, tab, sizeof sp2.match);
	      memcpy (sp2.lp, tab + kRsyMatch, sizeof sp2.lp);
	      memcpy (sp2.bp, tab + kRsyMatch + kRsyLp, sizeof sp2.bp);
	      const int64_t bp = h -> cap_rds;
	      rds_costas_kernel<<<(S + kRsyLanes - 1) / kRsyLanes, kRsyLanes, 0, rs>>> (
	            h -> d_rsy_in, bp, nout, S, sp2, h -> d_rsy_state, h -> d_rsy_c, bp,
	            h -> lf_plot == 9 ? h -> d_plot : nullptr, h -> cap_fm);
	      if (h -> lf_plot == 9 && h -> spec_N) CK (cudaEventRecord (h -> ev_sym, rs));
	      const dim3 gf ((unsigned)((nout + 127) / 128), (unsigned)S);
	      rds_fir_kernel<kRsyLp, false><<<gf, 128, 0, rs>>> (h -> d_rsy_c, h -> d_rsy_v, bp, nout, sp2, h -> d_rsy_state);
	      rds_fir_kernel<kRsyMatch, true><<<gf, 128, 0, rs>>> (h -> d_rsy_v, h -> d_rsy_w, bp, nout, sp2, h -> d_rsy_state);
	      const int spb = kRsyLanes / kRsyQuads;                               // streams per block
	      rds_bits_kernel<<<(S + spb - 1) / spb, kRsyLanes, 0, rs>>> (
	            h -> d_rsy_w, bp, nout, S, sp2, h -> d_rsy_state, h -> d_rsy_bits, h -> cap_bits, h -> d_rsy_nbits);
	      rds_sym_roll_kernel<<<S, 64, 0, rs>>> (h -> d_rsy_c, h -> d_rsy_v, bp, nout, h -> d_rsy_state);
	      h -> launches += 5;
	   }
	}
//	K6 ------------------------------------------------------------------------------------
const float2 *lr_in = (h -> d_lr + o);
	if (st.lf_cutoff_hz > 0) {
	   dim3 g ((unsigned)((M + kAlpTile - 1) / kAlpTile), (unsigned)S);
	   audio_lp_kernel<<<g, kAlpThreads, 0, h -> stream>>> ((h -> d_lr + o), h -> cap_fm, M,
	                                                        h -> d_alp_hist [h -> alp_sel], (h -> d_lrf + o));
	   roll_history_kernel<<<dim3 (kAlpHist / 256, S), 256, 0, h -> stream>>> (
	         (h -> d_lr + o), h -> cap_fm, h -> d_alp_hist [h -> alp_sel], h -> d_alp_hist [h -> alp_sel ^ 1], M, kAlpHist);
	   h -> alp_sel ^= 1; h -> launches += 2;
	   lr_in = (h -> d_lrf + o);
	}
const int64_t q0 = h -> fm_total / kRsDecim;
const int64_t q1 = (h -> fm_total + M) / kRsDecim;
const int32_t nq = (int32_t)(q1 - q0);
const bool convert = h -> cfg.audio_rate != h -> cfg.working_rate;
//	working-rate PCM: straight into the caller's buffer unless the second converter follows
float2 *aout = ((d_audio_out && !convert) ? d_audio_out : h -> d_audio) + nq_tot;
const int64_t apitch = (d_audio_out && !convert) ? audio_pitch : h -> cap_audio;
	{
	   AudioParams ap;
	   ap.alpha = st.deemph_alpha;
	   ap.gl = st.volume * st.left_ch; ap.gr = st.volume * st.right_ch;   // fm-processor.cpp:304-305
	   ap.M = M; ap.g0 = h -> fm_total; ap.q0 = q0; ap.nq = nq;
	   ap.fade_cnt = h -> fade_cnt; ap.fade_max = h -> fade_max;
	   ap.write_tap = h -> cfg.keep_taps;
	   ap.sel = h -> ahist_sel;
	   ap.plot = (h -> lf_plot >= 5 && h -> lf_plot <= 7) ? h -> lf_plot : 0;
	   ap.tone.on = h -> tone_on; ap.tone.arm = h -> tone_arm; ap.tone.burst = h -> tone_burst; ap.tone.pos = h -> tone_pos;
	   dim3 g ((unsigned)((M + kAuTile - 1) / kAuTile), (unsigned)S);
	   audio_kernel<<<g, kAuThreads, 0, h -> stream>>> (
	         lr_in, h -> cap_fm, ap, h -> d_ahist [h -> ahist_sel], h -> d_ahist [h -> ahist_sel ^ 1],
	         h -> d_state, (h -> d_a192 + o), aout, apitch, h -> d_plot ? h -> d_plot + o : nullptr, h -> d_tone_tab);
	   h -> launches ++;
	   h -> ahist_sel ^= 1;
	   h -> fade_cnt = h -> fade_cnt > nq ? h -> fade_cnt - nq : 0;
	   if (h -> tone_on) h -> tone_pos = (h -> tone_pos + nq) % (h -> tone_arm + h -> tone_burst);
	}
//	evaluatePeakLevel (:772-798): one read-out per 961 PCM samples, counted from the start of the processor
	if (sl == 0) h -> peak_e0 = q0 / h -> peak_block;
	h -> peak_e1 = q1 / h -> peak_block;
	if (nq > 0) {
	   const unsigned nb = (unsigned)((q1 - 1) / h -> peak_block - q0 / h -> peak_block + 1);
	   peak_kernel<<<dim3 (nb, (unsigned)S), 32, 0, h -> stream>>> (aout, apitch, q0, nq, h -> peak_block, h -> d_state,
	                                                             h -> peak_sel, h -> d_peak_ring);
	   h -> peak_sel ^= 1; h -> launches ++;
	}
int64_t n_out = nq;
	if (convert) {
//	   sendSampletoOutput (:825-838): the second converter, working rate -> audio rate
	   ConvertParams cp;
	   cp.L = h -> cvL; cp.M = h -> cvM; cp.P = kCvTapsPerPhase;
	   cp.t0 = h -> cv_in_total; cp.nt = nq;
	   const int64_t t1 = cp.t0 + nq;
	   cp.k0 = (cp.t0 * cp.L + cp.M - 1) / cp.M;                        // outputs that existed before this call: ceil (T L / M)
	   cp.nk = (int32_t)((t1 * cp.L + cp.M - 1) / cp.M - cp.k0);
	   float2 *cout = d_audio_out ? d_audio_out + na_tot : nullptr;     // (no caller buffer: nothing to write)
	   const int64_t cpitch = d_audio_out ? audio_pitch : h -> cap_out;
	   if (d_audio_out && cp.nk > 0) {
	      audio_convert_kernel<<<dim3 ((unsigned)((cp.nk + 255) / 256), (unsigned)S), 256, 0, h -> stream>>> (
	            aout, apitch, h -> d_cv_hist [h -> cv_sel], h -> d_cv_taps, cp, cout, cpitch);
	      h -> launches ++;
	   }
	   convert_roll_kernel<<<S, kCvTapsPerPhase, 0, h -> stream>>> (aout, apitch, nq, kCvTapsPerPhase, h -> d_cv_hist [h -> cv_sel],
	                                                                h -> d_cv_hist [h -> cv_sel ^ 1]);
	   h -> cv_sel ^= 1; h -> launches ++;
	   h -> cv_in_total = t1;
	   n_out = cp.nk;
	}
	h -> fm_total += M;
	na_tot += n_out; nq_tot += nq;
	}        // time slices
	h -> last_naudio = na_tot;
	if (n_audio) *n_audio = na_tot;
	if (st.rds_mode != 0) CK (cudaStreamWaitEvent (h -> stream, h -> ev_rds, 0));     // join
	if (h -> spec_N && h -> lf_plot >= 0) {
	   const int rc = run_lf_spectrum (h, Mcall);
	   if (rc != SDRJFM_OK) return rc;
	}
	CK (cudaGetLastError ());
	return SDRJFM_OK;
}

// stage (pending | new) samples when the call is not aligned to the decimation; returns the source to
// read.  Samples are moved as bytes (bps = bytes per IQ sample of the call's format).
static int stage_input (Lane *h, const void *iq, int fmt, int64_t n_in, int64_t in_pitch,
                        cudaMemcpyKind kind, const void **src, int64_t *pitch, int64_t *n_proc) {
const int S = h -> cfg.n_streams;
const int D = h -> decim;
const size_t bps = (size_t)fmt_bytes (fmt);
	if (h -> pend && fmt != h -> pend_fmt) { h -> err = "sample format changed while samples were pending"; return SDRJFM_ERR_ARG; }
const int64_t total = h -> pend + n_in;
	*n_proc = (total / D) * D;
//	zero-copy needs rows the kernel can address as whole samples
	if (kind == cudaMemcpyDeviceToDevice && h -> pend == 0 && *n_proc == n_in && ((uintptr_t)iq % bps) == 0) {
	   *src = iq; *pitch = in_pitch;
	   return SDRJFM_OK;
	}
char *din = (char *)h -> d_in; char *dp = (char *)h -> d_pend;
const size_t rowb = (size_t)h -> cap_in * bps;
	if (h -> pend)
	   CK (cudaMemcpy2DAsync (din, rowb, dp, D * bps, h -> pend * bps, S, cudaMemcpyDeviceToDevice, h -> stream));
	if (n_in)
	   CK (cudaMemcpy2DAsync (din + h -> pend * bps, rowb, iq, in_pitch * bps, n_in * bps, S, kind, h -> stream));
const int newpend = (int)(total - *n_proc);
	if (newpend)
	   CK (cudaMemcpy2DAsync (dp, D * bps, din + *n_proc * bps, rowb, newpend * bps, S,
	                          cudaMemcpyDeviceToDevice, h -> stream));
	h -> pend = newpend; h -> pend_fmt = fmt;
	*src = h -> d_in; *pitch = h -> cap_in;
	return SDRJFM_OK;
}

static RawFmt make_rawfmt (const Lane *h, int32_t fmt, float scale) {
RawFmt rf;
	memset (&rf, 0, sizeof rf);
	rf.fmt = fmt; rf.scale = scale;
	if (fmt == kFmtAirspy) { rf.blk = h -> air_blk; rf.map_int = h -> d_air_int; rf.map_frac = h -> d_air_frac; }
	return rf;
}

// airspy: n_in NATIVE samples arrive; a block of B native samples (+ the first one of the next block)
// gives 2304 samples at 2.304 MS/s (airspy-handler.cpp:283-309).  Stages (carried | new) native
// samples and returns how many OUTPUT samples this call produces (a multiple of 2304, hence of 12).
static int stage_input_airspy (Lane *h, const void *iq, int64_t n_in, int64_t in_pitch,
                               const void **src, int64_t *pitch, int64_t *n_proc) {
const int S = h -> cfg.n_streams;
const int64_t B = h -> air_blk;
	if (B <= 0) { h -> err = "set the native rate first (sdrjfm_set_native_rate)"; return SDRJFM_ERR_ARG; }
	if (h -> pend) { h -> err = "sample format changed while samples were pending"; return SDRJFM_ERR_ARG; }
const int64_t total = h -> air_pend + n_in;
const int64_t blocks = total >= 1 ? (total - 1) / B : 0;
	*n_proc = blocks * kAirspyOut;
const size_t bps = sizeof (short2);
const size_t rowb = (size_t)h -> cap_in * sizeof (float2);             // staging rows are cap_in * 8 bytes
	if ((size_t)total * bps > rowb || *n_proc > h -> cfg.max_samples_per_call) {
	   h -> err = "airspy call too large for max_samples_per_call"; return SDRJFM_ERR_CAPACITY;
	}
char *din = (char *)h -> d_in; char *dp = (char *)h -> d_air_pend;
const size_t prow = (size_t)(B + 1) * bps;
	if (h -> air_pend)
	   CK (cudaMemcpy2DAsync (din, rowb, dp, prow, h -> air_pend * bps, S, cudaMemcpyDeviceToDevice, h -> stream));
	if (n_in)
	   CK (cudaMemcpy2DAsync (din + h -> air_pend * bps, rowb, iq, in_pitch * bps, n_in * bps, S,
	                          cudaMemcpyDeviceToDevice, h -> stream));
const int64_t keep = total - blocks * B;                               // 1 .. B once a sample has arrived
	if (keep)
	   CK (cudaMemcpy2DAsync (dp, prow, din + blocks * B * bps, rowb, keep * bps, S,
	                          cudaMemcpyDeviceToDevice, h -> stream));
	h -> air_pend = (int)keep;
	*src = h -> d_in; *pitch = (int64_t)(rowb / bps);
	return SDRJFM_OK;
}

static int lane_process_device (Lane *h, const void *d_iq, int32_t fmt, float scale, int64_t n_in, int64_t in_pitch,
                           float *d_audio, int64_t audio_pitch, int64_t *n_audio,
                           float *d_rds24, int64_t rds_pitch, int64_t *n_rds) {
	if (!h || n_in < 0 || (n_in > 0 && (!d_iq || in_pitch < n_in))) return SDRJFM_ERR_ARG;
	if (fmt < kFmtCF32 || fmt > kFmtAirspy) return SDRJFM_ERR_ARG;
	if (n_in > h -> cfg.max_samples_per_call) { h -> err = "n_in exceeds max_samples_per_call"; return SDRJFM_ERR_CAPACITY; }
	CK (cudaSetDevice (h -> cfg.device));
//	every argument / format check comes BEFORE anything is staged or any state moves
	if (fmt != kFmtAirspy && h -> air_pend) { h -> err = "sample format changed while samples were pending"; return SDRJFM_ERR_ARG; }
	if (h -> hf_N && n_in > 0 && fmt != kFmtAirspy) {
	   const int rc2 = run_hf_spectrum (h, d_iq, make_rawfmt (h, fmt, scale), in_pitch, n_in);
	   if (rc2 != SDRJFM_OK) return rc2;
	}
	else h -> last_nhf = 0;
const void *src; int64_t pitch, n_proc;
int rc = fmt == kFmtAirspy ? stage_input_airspy (h, d_iq, n_in, in_pitch, &src, &pitch, &n_proc)
                           : stage_input (h, d_iq, fmt, n_in, in_pitch, cudaMemcpyDeviceToDevice, &src, &pitch, &n_proc);
	if (rc != SDRJFM_OK) return rc;
const RawFmt rf = make_rawfmt (h, fmt, scale);
	return run_chain (h, src, rf, pitch, n_proc, (float2 *)d_audio, audio_pitch, n_audio,
	                  (float2 *)d_rds24, rds_pitch, n_rds);
}

static int lane_get_meta (Lane *h, sdrjfm_meta *meta) {
	if (!h || !meta) return SDRJFM_ERR_ARG;
const int S = h -> cfg.n_streams;
std::vector<StreamState> st (S);
	CK (cudaMemcpyAsync (st.data (), h -> d_state, S * sizeof (StreamState), cudaMemcpyDeviceToHost, h -> stream));
//	inputFilter on: the fm-rate DC estimate runs behind the 5440-sample delay line; RfDC as of the last input sample
const bool adv = h -> set.input_filter_hz > 0 && h -> set.dc_remove && !h -> hybrid_last && h -> d_sdel [0] && h -> in_total > 0;
std::vector<double2> dcnow (adv ? S : 0);
	if (adv) {
	   CK (cudaSetDevice (h -> cfg.device));
	   if (!h -> d_dcnow) CK (dalloc (&h -> d_dcnow, (size_t)S));
	   const double alpha = (double)(1.0f / h -> cfg.input_rate);
	   dc_advance_kernel<<<S, 256, 0, h -> stream>>> (h -> d_sdel [h -> del_sel], h -> fw_delay, alpha,
	                                                 pow (1.0 - alpha, (double)h -> decim), h -> d_state, h -> d_dcnow);
	   CK (cudaMemcpyAsync (dcnow.data (), h -> d_dcnow, S * sizeof (double2), cudaMemcpyDeviceToHost, h -> stream));
	}
//	showPeakLevel: the newest read-out of the peak meter behind the display delay line (-40 dB before the first)
std::vector<float2> peaks (S, make_float2 (-40.0f, -40.0f));
	{  const int64_t E = (h -> fm_total / kRsDecim) / h -> peak_block - 1 - h -> peak_delay;
	   if (E >= h -> peak_e_set && E >= 0)
	      CK (cudaMemcpy2DAsync (peaks.data (), sizeof (float2), h -> d_peak_ring + (E & (kPeakRing - 1)), kPeakRing * sizeof (float2),
	                             sizeof (float2), S, cudaMemcpyDeviceToHost, h -> stream)); }
	CK (cudaStreamSynchronize (h -> stream));
	for (int s = 0; s < S; s ++) {
	   sdrjfm_meta &m = meta [s];
	   const StreamState &x = st [s];
	   m.dc_rf_re = (float)(adv ? dcnow [s].x : x.dc_re); m.dc_rf_im = (float)(adv ? dcnow [s].y : x.dc_im);
	   m.dc_rf_db = h -> set.dc_remove ?
	         20 * log10f (hypotf (m.dc_rf_re, m.dc_rf_im) + 1.0f / 32768) : -99.99f;
	   m.dc_if = x.fm_afc;
	   m.carrier_ampl = x.am_carr_ampl;
	   m.pss_phase_shift_deg = (float)(x.pss_delay / M_PI * 180.0f);
	   m.pss_phase_change = x.pss_mean_error * 1000;
	   const bool locked = h -> set.fm_mode != 2 && x.pilot_locked;       // isPilotLocked, :869-879
	   m.pilot_locked = locked;
	   m.pilot_lock_strength = h -> set.fm_mode != 2 ? x.pilot_lock : 0.f;
	   m.pss_state = (h -> set.pss_on && locked) ? (x.pss_minimized ? 2 : 1) : 0;
	   m.peak_left_db = peaks [s].x; m.peak_right_db = peaks [s].y;
	   m.squelch_active = 0;
	}
	if (h -> d_sq) {                                                   // getSquelchState (:217-219)
	   std::vector<SquelchState> sq (S);
	   CK (cudaMemcpyAsync (sq.data (), h -> d_sq, S * sizeof (SquelchState), cudaMemcpyDeviceToHost, h -> stream));
	   CK (cudaStreamSynchronize (h -> stream));
	   for (int s = 0; s < S; s ++) meta [s].squelch_active = sq [s].suppress;
	}
	return SDRJFM_OK;
}

static int64_t lane_read_tap (Lane *h, int which, int32_t stream, void *out, int64_t cap) {
	if (!h || !out || stream < 0 || stream >= h -> cfg.n_streams) return SDRJFM_ERR_ARG;
const void *src = nullptr; size_t esz = 0; int64_t n = h -> last_nfm; int64_t pitch = h -> cap_fm;
	switch (which) {
	   case SDRJFM_TAP_FM_Z:        src = h -> d_fmz;    esz = 8; break;
	   case SDRJFM_TAP_DEMOD:       src = h -> d_demod;  esz = 4; break;
	   case SDRJFM_TAP_PILOT_PHASE: src = h -> d_phase;  esz = 4; break;
	   case SDRJFM_TAP_LOCKED:      src = h -> d_locked; esz = 1; break;
	   case SDRJFM_TAP_PSS_DELAY:   src = h -> d_pssd;   esz = 4; break;
	   case SDRJFM_TAP_LR:          src = h -> d_lr;     esz = 8; break;
	   case SDRJFM_TAP_AUDIO192:    src = h -> d_a192;   esz = 8; break;
	   case SDRJFM_TAP_RDS_CPLX:    src = h -> d_rdsc;   esz = 8; break;
	   case SDRJFM_TAP_RDS24:       src = h -> d_rds24;  esz = 8; n = h -> last_nrds; pitch = h -> cap_rds; break;
	   case SDRJFM_TAP_FE_U:        src = h -> d_U;      esz = 8; break;
	   case SDRJFM_TAP_FE_S:        src = h -> d_S;      esz = 8; break;
	   default: return SDRJFM_ERR_ARG;
	}
	if (n > cap) n = cap;
	if (n <= 0) return 0;
	CK (cudaMemcpyAsync (out, (const char *)src + (size_t)stream * pitch * esz, n * esz,
	                     cudaMemcpyDeviceToHost, h -> stream));
	CK (cudaStreamSynchronize (h -> stream));
	return n;
}

// ---- settings (fm-processor.cpp:232-301, 351-359, 762-770, 840-933) -----------------------
static int lane_set_fm_mode (Lane *h, int32_t m) {
	if (!h || m < 0 || m > 2) return SDRJFM_ERR_ARG;
	h -> set.fm_mode = m; return SDRJFM_OK;
}
static int lane_set_fm_decoder (Lane *h, int32_t d) {
	if (!h || d < 1 || d > 6) return SDRJFM_ERR_ARG;
	h -> set.decoder = d; return SDRJFM_OK;
}
static int lane_set_sound_mode (Lane *h, int32_t s) {
	if (!h || s < 0 || s > 6) return SDRJFM_ERR_ARG;
	h -> set.sound_sel = s; return SDRJFM_OK;
}
static int lane_set_stereo_panorama (Lane *h, int32_t pan) {
	if (!h) return SDRJFM_ERR_ARG;
	h -> set.panorama = (float)pan / 100.0f; return SDRJFM_OK;           // :279
}
static int lane_set_sound_balance (Lane *h, int32_t balance) {
	if (!h) return SDRJFM_ERR_ARG;
	h -> set.left_ch  = (balance > 0 ? (100 - balance) / 100.0 : 1.0f);   // :284-285
	h -> set.right_ch = (balance < 0 ? (100 + balance) / 100.0 : 1.0f);
	return SDRJFM_OK;
}
static int lane_set_deemphasis (Lane *h, int32_t v) {
	if (!h || v < 1) return SDRJFM_ERR_ARG;
//	K6 restarts the one-pole kAuWarm = 384 samples ahead of every tile from a zero state: the state
//	error left is (1 - alpha)^384, below 1e-8 only up to 100 us (the GUI offers 1 ("Off"), 50 and 75)
	if (v > 100) { h -> err = "de-emphasis time constants above 100 us are not supported (tile warm-up of the one-pole)"; return SDRJFM_ERR_UNSUPPORTED; }
float Tau = 1000000.0 / v;                                                // :295-296
	h -> set.deemph_us = v;
	h -> set.deemph_alpha = 1.0 / (float (h -> cfg.fm_rate) / Tau + 1.0);
	return SDRJFM_OK;
}
static int lane_set_volume_db (Lane *h, float db) {
	if (!h) return SDRJFM_ERR_ARG;
	h -> set.volume = std::pow (10.0f, db / 20.0f); return SDRJFM_OK;     // :300
}
static int lane_set_lf_cutoff (Lane *h, int32_t hz) {
	if (!h) return SDRJFM_ERR_ARG;
//	setlfcutoff (:762-770): <= 0 switches fmAudioFilter off; else it is re-designed and starts cleared
	CK (cudaSetDevice (h -> cfg.device));
	const int32_t v = hz > 0 ? hz : 0;
	if (v == h -> set.lf_cutoff_hz) return SDRJFM_OK;
	h -> set.lf_cutoff_hz = v;
	if (v > 0) {
	   const int64_t S = h -> cfg.n_streams;
	   if (!h -> d_lrf) {
	      CK (dalloc (&h -> d_alp_hist [0], (size_t)S * kAlpHist)); CK (dalloc (&h -> d_alp_hist [1], (size_t)S * kAlpHist));
	      CK (dalloc (&h -> d_lrf, (size_t)S * h -> cap_fm));
	   }
	   else for (int i = 0; i < 2; i ++)
	      CK (cudaMemsetAsync (h -> d_alp_hist [i], 0, (size_t)S * kAlpHist * sizeof (float2), h -> stream));
	   std::vector<cf32> lp = design_lowpass (kAlpTaps, v, h -> cfg.fm_rate);
	   for (int i = 0; i < kAlpTaps; i ++) h -> ci.alp [i] = lp [i].real ();
	   int rc = consts_commit (h);
	   if (rc != SDRJFM_OK) return rc;
	}
	return SDRJFM_OK;
}
static int lane_set_bandwidth (Lane *h, int32_t hz) {
	if (!h) return SDRJFM_ERR_ARG;
//	setBandwidth (:232-239): "Off" -> 0; else fmBandwidth in Hz, the low-pass corner is hz / 2 (:398).
//	Switching the filter on (or changing it) starts it from a cleared state.
	CK (cudaSetDevice (h -> cfg.device));
	const int32_t v = hz > 0 ? hz : 0;
	if (v > 0 && h -> resample) { h -> err = "inputFilter is not available in the resampler mode"; return SDRJFM_ERR_UNSUPPORTED; }
	if (v > 0 && h -> shape == kShapeGeneric) { h -> err = "inputFilter is built for the decimations 12, 30 and 48 (2.304 / 2.4, 6, 10 MS/s)"; return SDRJFM_ERR_UNSUPPORTED; }
	if (v == h -> set.input_filter_hz) return SDRJFM_OK;
	h -> set.input_filter_hz = v;
	int rc = rebuild_tables (h);
	if (rc != SDRJFM_OK) return rc;
	if (v > 0) return wide_setup (h);
	return SDRJFM_OK;
}
static int lane_set_attenuation (Lane *h, float l, float r) {
	if (!h) return SDRJFM_ERR_ARG;
	h -> set.lgain = l; h -> set.rgain = r; return SDRJFM_OK;             // :356-357
}
static int lane_set_rds_mode (Lane *h, int32_t m) {
	if (!h || m < 0 || m > 3) return SDRJFM_ERR_ARG;
//	all three RDS demodulator variants (rds-decoder.cpp) consume the same 24 kHz baseband; the
//	selector only matters to the (host-side, untouched) rdsDecoder.  Switching RDS on starts
//	the branch from cleared filters (the reference would resume with stale block contents).
	if (m != 0 && h -> set.rds_mode == 0) {
	   CK (cudaSetDevice (h -> cfg.device));
	   int rc = rds_setup (h);
	   if (rc != SDRJFM_OK) return rc;
	}
	if (h -> lf_plot >= 8) h -> spec_refresh = true;                      // setfmRdsSelector: new_lfSpectrum (:843-846)
	h -> set.rds_mode = m; return SDRJFM_OK;
}
// the symbol stage of rdsDecoder::doDecode for mode RDS_1 on the GPU (optional).
// Switching it on starts the Costas loop and the decoder's filters from their constructor state.
static int rs2_reset (Lane *h) {            // rdsDecoder_2's constructor state (rds-decoder-2.cpp:44-77)
const size_t S = h -> cfg.n_streams;
std::vector<Rds2State> st (S);
	memset (st.data (), 0, S * sizeof (Rds2State));
	for (auto &q : st) { q.gain = 9.0f; q.skip = 3; }
	CK (cudaMemcpy (h -> d_rs2_state, st.data (), S * sizeof (Rds2State), cudaMemcpyHostToDevice));
	return SDRJFM_OK;
}
static int rs3_reset (Lane *h) {
const size_t S = h -> cfg.n_streams;
std::vector<Rds3State> st (S);
	memset (st.data (), 0, S * sizeof (Rds3State));
	for (auto &q : st) q.resync = 1;
	CK (cudaMemcpy (h -> d_rs3_state, st.data (), S * sizeof (Rds3State), cudaMemcpyHostToDevice));
	CK (cudaMemset (h -> d_rs3_ngroups, 0, S * sizeof (int32_t)));
	CK (cudaMemset (h -> d_rs3_stat, 0, S * 4 * sizeof (int32_t)));
	return SDRJFM_OK;
}
static int lane_set_rds_symbol_stage (Lane *h, int32_t on) {
	if (!h) return SDRJFM_ERR_ARG;
	CK (cudaSetDevice (h -> cfg.device));
const size_t S = h -> cfg.n_streams;
	if (on && !h -> d_rs3_state) {
//	   rdsDecoder_3's constructor state (rds-decoder-3.cpp:44-79): everything zero, Resync set; rdsBlockSynchronizer::reset
	   std::vector<float> tab;
	   design_rds3_clock (24000, tab, &h -> rs3_omega, h -> rs3_kmap);
	   CK (dalloc (&h -> d_rs3_sin, tab.size ()));
	   CK (cudaMemcpy (h -> d_rs3_sin, tab.data (), tab.size () * sizeof (float), cudaMemcpyHostToDevice));
	   h -> cap_groups = (int32_t)(h -> cap_rds / 2000 + 4);               // a group is 104 bits = 2101 samples
	   CK (dalloc (&h -> d_rs3_state, S)); CK (dalloc (&h -> d_rs3_groups, S * h -> cap_groups * 4));
	   CK (dalloc (&h -> d_rs3_ngroups, S)); CK (dalloc (&h -> d_rs3_stat, S * 4));
	   const int rc = rs3_reset (h); if (rc != SDRJFM_OK) return rc;
	}
	else if (on && !h -> rds_symbols) {
	   CK (cudaStreamSynchronize (h -> stream)); CK (cudaStreamSynchronize (h -> stream_rds));
	   const int rc = rs3_reset (h); if (rc != SDRJFM_OK) return rc;
	}
	if (on && !h -> d_rs2_state) {
	   CK (dalloc (&h -> d_rs2_state, S)); CK (dalloc (&h -> d_rs2_m, S * h -> cap_rds));
	   const int rc = rs2_reset (h); if (rc != SDRJFM_OK) return rc;
	}
	else if (on && !h -> rds_symbols) {
	   CK (cudaStreamSynchronize (h -> stream)); CK (cudaStreamSynchronize (h -> stream_rds));
	   const int rc = rs2_reset (h); if (rc != SDRJFM_OK) return rc;
	}
	if (on && !h -> d_rsy_state) {
	   h -> cap_bits = (int32_t)(h -> cap_rds / 16 + 16);          // ~20.2 samples per bit
	   CK (dalloc (&h -> d_rsy_state, S));
	   CK (dalloc (&h -> d_rsy_bits, S * h -> cap_bits));
	   CK (dalloc (&h -> d_rsy_nbits, S));
	   CK (dalloc (&h -> d_rsy_c, S * h -> cap_rds)); CK (dalloc (&h -> d_rsy_v, S * h -> cap_rds));
	   CK (dalloc (&h -> d_rsy_w, S * h -> cap_rds));
	   CK (dalloc (&h -> d_rsy_in, S * h -> cap_rds));
	}
	else if (on && !h -> rds_symbols) {
	   CK (cudaStreamSynchronize (h -> stream));
	   CK (cudaStreamSynchronize (h -> stream_rds));
	   CK (cudaMemset (h -> d_rsy_state, 0, S * sizeof (RdsSymState)));
	   CK (cudaMemset (h -> d_rsy_nbits, 0, S * sizeof (int32_t)));
	}
	h -> rds_symbols = on != 0;
	return SDRJFM_OK;
}
// bits decoded for `stream` by the LAST process call; returns their number
static int64_t lane_read_rds_bits (Lane *h, int32_t stream, uint8_t *out, int64_t cap) {
	if (!h || !out || stream < 0 || stream >= h -> cfg.n_streams) return SDRJFM_ERR_ARG;
	if (!h -> rds_symbols || !h -> d_rsy_nbits || h -> last_nrds == 0) return 0;
	CK (cudaSetDevice (h -> cfg.device));
int32_t n = 0;
	CK (cudaStreamSynchronize (h -> stream_rds));                        // the symbol stage runs past the call's join
	CK (cudaMemcpyAsync (&n, h -> d_rsy_nbits + stream, sizeof n, cudaMemcpyDeviceToHost, h -> stream));
	CK (cudaStreamSynchronize (h -> stream));
	if (n > h -> cap_bits) n = h -> cap_bits;
	if (n > cap) n = (int32_t)cap;
	if (n > 0) {
	   CK (cudaMemcpyAsync (out, h -> d_rsy_bits + (size_t)stream * h -> cap_bits, n, cudaMemcpyDeviceToHost, h -> stream));
	   CK (cudaStreamSynchronize (h -> stream));
	}
	return n;
}
// groups completed for `stream` by the LAST process call in mode RDS_3 (blocks A..D each), and the synchroniser's status:
// status [0] synchronised, [1] bit-clock re-synchronisations in that call, [2] sync errors, [3] crc errors
static int64_t lane_read_rds_groups (Lane *h, int32_t stream, uint16_t *out, int64_t cap, int32_t *status) {
	if (!h || stream < 0 || stream >= h -> cfg.n_streams || (cap > 0 && !out)) return SDRJFM_ERR_ARG;
	if (status) memset (status, 0, 4 * sizeof (int32_t));
	if (!h -> rds_symbols || !h -> d_rs3_ngroups || h -> set.rds_mode != 3 || h -> last_nrds == 0) return 0;
	CK (cudaSetDevice (h -> cfg.device));
int32_t n = 0;
	CK (cudaStreamSynchronize (h -> stream_rds));
	CK (cudaMemcpyAsync (&n, h -> d_rs3_ngroups + stream, sizeof n, cudaMemcpyDeviceToHost, h -> stream));
	if (status) CK (cudaMemcpyAsync (status, h -> d_rs3_stat + 4 * stream, 4 * sizeof (int32_t), cudaMemcpyDeviceToHost, h -> stream));
	CK (cudaStreamSynchronize (h -> stream));
	if (n > h -> cap_groups) n = h -> cap_groups;
	if (n > cap) n = (int32_t)cap;
	if (n > 0) {
	   CK (cudaMemcpyAsync (out, h -> d_rs3_groups + (size_t)stream * h -> cap_groups * 4, (size_t)n * 4 * sizeof (uint16_t),
	                        cudaMemcpyDeviceToHost, h -> stream));
	   CK (cudaStreamSynchronize (h -> stream));
	}
	return n;
}
// setlfPlotType (fm-processor.cpp:244-266).  type -1: no scope stream (nothing extra is written)
static int lane_set_lf_plot_type (Lane *h, int32_t type) {
	if (!h || type < -1 || type > 9) return SDRJFM_ERR_ARG;
	CK (cudaSetDevice (h -> cfg.device));
	if (((type >= 4 && type <= 7) || type == 9) && !h -> d_plot) CK (dalloc (&h -> d_plot, (size_t)h -> cfg.n_streams * h -> cap_fm));
	h -> lf_plot = type;
	h -> spec_refresh = true;                                             // lfBuffer_newFlag (:265)
	return SDRJFM_OK;
}
// setlfPlotZoomFactor (fm-processor.cpp:268-271)
static int lane_set_lf_plot_zoom (Lane *h, int32_t zoom) {
	if (!h || zoom < 1) return SDRJFM_ERR_ARG;
	h -> spec_zoom = zoom; h -> spec_refresh = true;
	return SDRJFM_OK;
}
// The LF scope's display spectrum on the GPU: ls_scope (spectrumSize, displaySize, averageCount), radio.cpp:238-249,
// 2009-2014.  spectrum_size 0 switches it off.
static int lane_set_lf_spectrum (Lane *h, int32_t N, int32_t display, int32_t average_count) {
	if (!h) return SDRJFM_ERR_ARG;
	CK (cudaSetDevice (h -> cfg.device));
	if (N == 0) { h -> spec_N = 0; return SDRJFM_OK; }
	if (N < 64 || N > kSpecMaxN || (N & (N - 1)) || display < 2 || (display & (display - 1)) || N < 2 * display || average_count < 1) {
	   h -> err = "LF spectrum: spectrum_size a power of two in 64..4096, display_size a power of two <= spectrum_size / 2, average_count >= 1";
	   return SDRJFM_ERR_ARG;
	}
const size_t S = h -> cfg.n_streams;
	CK (cudaStreamSynchronize (h -> stream));
	for (void *p : { (void *)h -> d_spec_carry [0], (void *)h -> d_spec_carry [1], (void *)h -> d_spec_win,
	                 (void *)h -> d_spec_Y, (void *)h -> d_spec_avg, (void *)h -> d_spec_disp }) if (p) cudaFree (p);
	h -> d_spec_carry [0] = h -> d_spec_carry [1] = nullptr; h -> d_spec_win = nullptr;
	h -> d_spec_Y = h -> d_spec_avg = h -> d_spec_disp = nullptr;
	h -> spec_N = 0;
	if (!h -> d_spec_in) CK (dalloc (&h -> d_spec_in, S * h -> cap_fm));
	if (!h -> ev_sym) CK (cudaEventCreateWithFlags (&h -> ev_sym, cudaEventDisableTiming));
	h -> cap_specblk = (int32_t)(h -> cap_fm / N + 2);
	CK (dalloc (&h -> d_spec_carry [0], S * N)); CK (dalloc (&h -> d_spec_carry [1], S * N));
	CK (dalloc (&h -> d_spec_win, (size_t)N));
	CK (dalloc (&h -> d_spec_Y, S * h -> cap_specblk * display));
	CK (dalloc (&h -> d_spec_avg, S * display)); CK (dalloc (&h -> d_spec_disp, S * display));
std::vector<float> win ((size_t)N);
	for (int i = 0; i < N; i ++)                                          // ls-scope.cpp:50-52, evaluated in double
	   win [i] = 0.43 - 0.5 * cos ((2.0 * M_PI * i) / N) + 0.08 * cos ((4.0 * M_PI * i) / (N - 1));
	CK (cudaMemcpy (h -> d_spec_win, win.data (), (size_t)N * sizeof (float), cudaMemcpyHostToDevice));
	h -> spec_logN = 0; while ((1 << h -> spec_logN) < N) h -> spec_logN ++;
	h -> spec_N = N; h -> spec_display = display; h -> spec_avg_count = average_count;
	h -> spec_sel = 0; h -> spec_carry = 0; h -> spec_refresh = true; h -> last_nspec = 0;
	return SDRJFM_OK;
}
// hs_scope (displaySize, rasterSize, SampleRate = inputRate, freq = repeatRate), radio.cpp:232-237; display_size 0: off
static int lane_set_hf_spectrum (Lane *h, int32_t display, int32_t repeat_rate) {
	if (!h) return SDRJFM_ERR_ARG;
	CK (cudaSetDevice (h -> cfg.device));
	CK (cudaStreamSynchronize (h -> stream));
	for (void *p : { (void *)h -> d_hf_blk, (void *)h -> d_hf_win, (void *)h -> d_hf_avg, (void *)h -> d_hf_disp }) if (p) cudaFree (p);
	h -> d_hf_blk = nullptr; h -> d_hf_win = nullptr; h -> d_hf_avg = h -> d_hf_disp = nullptr; h -> hf_N = 0;
	if (display == 0) return SDRJFM_OK;
	if (display < 16 || display > kSpecMaxN / 4 || (display & (display - 1)) || repeat_rate < 2 ||
	    h -> cfg.input_rate / repeat_rate < 4 * display) {
	   h -> err = "HF spectrum: display_size a power of two in 16..1024, repeat_rate >= 2, inputRate / repeat_rate >= 4 display_size";
	   return SDRJFM_ERR_ARG;
	}
const size_t S = h -> cfg.n_streams;
const int N = 4 * display;                                                  // spectrumSize = 4 * displaySize (:43)
	CK (dalloc (&h -> d_hf_blk, S * N)); CK (dalloc (&h -> d_hf_win, (size_t)N));
	CK (dalloc (&h -> d_hf_avg, S * display)); CK (dalloc (&h -> d_hf_disp, S * display));
std::vector<float> win ((size_t)N);
	for (int i = 0; i < N; i ++)                                                // hs-scope.cpp:66-68, evaluated in double
	   win [i] = 0.43 - 0.5 * cos ((2.0 * M_PI * i) / N) + 0.08 * cos ((4.0 * M_PI * i) / (N - 1));
	CK (cudaMemcpy (h -> d_hf_win, win.data (), (size_t)N * sizeof (float), cudaMemcpyHostToDevice));
	h -> hf_logN = 0; while ((1 << h -> hf_logN) < N) h -> hf_logN ++;
	h -> hf_display = display; h -> hf_seg = h -> cfg.input_rate / repeat_rate; h -> hf_half_freq = repeat_rate / 2;
	h -> hf_total = 0; h -> last_nhf = 0;
	h -> hf_N = N;
	return SDRJFM_OK;
}
static int64_t lane_read_hf_spectrum (Lane *h, int32_t stream, double *out, int64_t cap, int32_t *blocks) {
	if (!h || !out || stream < 0 || stream >= h -> cfg.n_streams) return SDRJFM_ERR_ARG;
	if (!h -> hf_N) { h -> err = "the HF spectrum is off (sdrjfm_set_hf_spectrum)"; return SDRJFM_ERR_ARG; }
	if (cap < h -> hf_display) return SDRJFM_ERR_CAPACITY;
	CK (cudaSetDevice (h -> cfg.device));
	CK (cudaMemcpyAsync (out, h -> d_hf_disp + (size_t)stream * h -> hf_display, h -> hf_display * sizeof (double),
	                     cudaMemcpyDeviceToHost, h -> stream));
	CK (cudaStreamSynchronize (h -> stream));
	if (blocks) *blocks = h -> last_nhf;
	return h -> hf_display;
}
// displayBuffer of one stream (display_size doubles) as the last process call left it
static int64_t lane_read_lf_spectrum (Lane *h, int32_t stream, double *out, int64_t cap, int32_t *blocks) {
	if (!h || !out || stream < 0 || stream >= h -> cfg.n_streams) return SDRJFM_ERR_ARG;
	if (!h -> spec_N) { h -> err = "the LF spectrum is off (sdrjfm_set_lf_spectrum)"; return SDRJFM_ERR_ARG; }
	if (cap < h -> spec_display) return SDRJFM_ERR_CAPACITY;
	CK (cudaSetDevice (h -> cfg.device));
	CK (cudaMemcpyAsync (out, h -> d_spec_disp + (size_t)stream * h -> spec_display, h -> spec_display * sizeof (double),
	                     cudaMemcpyDeviceToHost, h -> stream));
	CK (cudaStreamSynchronize (h -> stream));
	if (blocks) *blocks = h -> last_nspec;
	return h -> spec_display;
}
// What the LAST process call pushed into spectrumBuffer_lf for one stream (fm-processor.cpp:565-627):
// complex samples, at the fm rate for every type except RDS_INPUT / RDS_DEMOD with the RDS branch on
// (24 kHz).  The float streams become (x, 0) exactly as push_back (float) builds them.
static int64_t lane_read_lf_plot (Lane *h, int32_t stream, float *out, int64_t cap) {
	if (!h || !out || stream < 0 || stream >= h -> cfg.n_streams) return SDRJFM_ERR_ARG;
	if (h -> lf_plot < 0) { h -> err = "no LF scope stream selected (sdrjfm_set_lf_plot_type)"; return SDRJFM_ERR_ARG; }
	if (h -> scanning) return 0;                                          // :478-495: "continue" before the demodulator
const int type = h -> lf_plot;
const bool rds_rate = type >= 8 && h -> set.rds_mode != 0;
int64_t n = rds_rate ? h -> last_nrds : h -> last_nfm;
	if (n > cap) n = cap;
	if (n <= 0) return 0;
	CK (cudaSetDevice (h -> cfg.device));
const void *src = nullptr; size_t esz = 4; int64_t pitch = h -> cap_fm; float mul = 1.f;
	switch (type) {
	   case 0: memset (out, 0, (size_t)n * 8); return n;                  // OFF: zeros
	   case 1: src = h -> d_fmz; esz = 8; break;                          // IF_FILTERED: v
	   case 2: case 3: src = h -> d_demod; break;                         // DEMODULATOR, AF_SUM (sumLR = demod, :724-729)
	   case 4: case 5: case 6: case 7: src = h -> d_plot; break;          // AF_DIFF, AF_*_FILTERED
	   case 8:                                                            // RDS_INPUT: 20 rdsSample
	      if (!rds_rate) { memset (out, 0, (size_t)n * 8); return n; }    // RDS off: zeros at the fm rate (:579-586)
	      src = h -> last_rds_ptr; esz = 8; pitch = h -> last_rds_pitch; mul = 20.f; break;
	   case 9:                                                            // RDS_DEMOD: magCplx = 4 * Costas output (rds-decoder.cpp:76-77)
	      if (!rds_rate) { memset (out, 0, (size_t)n * 8); return n; }
	      if (!h -> rds_symbols) { h -> err = "the RDS_DEMOD scope stream needs the RDS symbol stage (sdrjfm_set_rds_symbol_stage)"; return SDRJFM_ERR_UNSUPPORTED; }
	      CK (cudaStreamSynchronize (h -> stream_rds));                   // the symbol stage runs past the call's join
	      {  // real parts from the symbol stage's own buffer, imaginary parts from d_plot
	         std::vector<float> im ((size_t)n);
	         CK (cudaMemcpyAsync (out + n, h -> d_rsy_c + (size_t)stream * h -> cap_rds, (size_t)n * 4, cudaMemcpyDeviceToHost, h -> stream));
	         CK (cudaMemcpyAsync (im.data (), h -> d_plot + (size_t)stream * h -> cap_fm, (size_t)n * 4, cudaMemcpyDeviceToHost, h -> stream));
	         CK (cudaStreamSynchronize (h -> stream));
	         for (int64_t i = 0; i < n; i ++) { const float re = out [n + i]; out [2 * i] = re * 4.0f; out [2 * i + 1] = im [i] * 4.0f; }
	         return n;
	      }
	}
	if (!src) return 0;
float *dst = esz == 8 ? out : out + n;                                    // floats land in the upper half, then spread
	CK (cudaMemcpyAsync (dst, (const char *)src + (size_t)stream * pitch * esz, (size_t)n * esz,
	                     cudaMemcpyDeviceToHost, h -> stream));
	CK (cudaStreamSynchronize (h -> stream));
	if (esz == 4)
	   for (int64_t i = 0; i < n; i ++) { const float v = dst [i]; out [2 * i] = v; out [2 * i + 1] = 0.f; }
	else if (mul != 1.f)
	   for (int64_t i = 0; i < 2 * n; i ++) out [i] *= mul;
	return n;
}
// startScanning / stopScanning (fm-processor.cpp:361-367).  scanPointer is a local of run (): it
// restarts at 0 whenever the run loop does; here it restarts with every startScanning.
static int lane_set_scanning (Lane *h, int32_t on) {
	if (!h) return SDRJFM_ERR_ARG;
	CK (cudaSetDevice (h -> cfg.device));
	if (on && !h -> d_scan_db) {
	   const size_t S = h -> cfg.n_streams;
	   h -> cap_scan = (int32_t)(h -> cap_fm / kScanN + 2);
	   CK (dalloc (&h -> d_scan_carry [0], S * kScanN)); CK (dalloc (&h -> d_scan_carry [1], S * kScanN));
	   CK (dalloc (&h -> d_scan_db, S * h -> cap_scan));
	}
	if (on && !h -> scanning) h -> scan_carry = 0;
	h -> scanning = on != 0; h -> last_nscan = 0;
	return SDRJFM_OK;
}
// (signal dB, noise dB) pairs of the 1024-sample blocks completed by the LAST process call
static int64_t lane_read_scan (Lane *h, int32_t stream, float *out, int64_t cap_pairs) {
	if (!h || !out || stream < 0 || stream >= h -> cfg.n_streams) return SDRJFM_ERR_ARG;
int64_t n = h -> scanning ? h -> last_nscan : 0;
	if (n > cap_pairs) n = cap_pairs;
	if (n <= 0) return 0;
	CK (cudaSetDevice (h -> cfg.device));
	CK (cudaMemcpyAsync (out, h -> d_scan_db + (size_t)stream * h -> cap_scan, n * sizeof (float2), cudaMemcpyDeviceToHost, h -> stream));
	CK (cudaStreamSynchronize (h -> stream));
	return n;
}
static int lane_set_local_oscillator (Lane *h, int32_t hz) {
	if (!h) return SDRJFM_ERR_ARG;
//	set_localOscillator (:865-867).  The oscillator table (inputRate complex entries, oscillator.cpp:26-37)
//	is built the first time lo is non-zero.
	if (hz <= -h -> cfg.input_rate || hz >= h -> cfg.input_rate) return SDRJFM_ERR_ARG;
	if (hz != 0 && h -> resample) { h -> err = "the local oscillator is not available in the resampler mode"; return SDRJFM_ERR_UNSUPPORTED; }
	CK (cudaSetDevice (h -> cfg.device));
	if (hz != 0 && !h -> d_lo_tab) {
	   const int32_t R = h -> cfg.input_rate;
	   std::vector<float2> t (R);
	   for (int32_t i = 0; i < R; i ++)
	      t [i] = make_float2 ((float)cos (2.0 * M_PI * i / R), (float)sin (2.0 * M_PI * i / R));
	   CK (cudaMalloc ((void **)&h -> d_lo_tab, (size_t)R * sizeof (float2)));
	   CK (cudaMemcpy (h -> d_lo_tab, t.data (), (size_t)R * sizeof (float2), cudaMemcpyHostToDevice));
	}
	if (hz == h -> set.lo_hz) return SDRJFM_OK;
	h -> set.lo_hz = hz;
	CK (cudaStreamSynchronize (h -> stream));
	return upload_tables (h);        // taps with / without the DC folding, H (lo)
}
static int lane_set_squelch_mode (Lane *h, int32_t m) {
	if (!h || m < 0 || m > 2) return SDRJFM_ERR_ARG;
//	set_squelchMode (:882-884): ESqMode OFF / NSQ / LSQ.  The squelch object lives as long as the
//	processor: its filters and counters keep their state while the mode is off.
	if (m != 0 && !h -> d_sq) {
	   CK (cudaSetDevice (h -> cfg.device));
	   CK (dalloc (&h -> d_sq, (size_t)h -> cfg.n_streams));
	}
	h -> set.squelch_mode = m; return SDRJFM_OK;
}
static int lane_set_squelch_value (Lane *h, int32_t n) {
	if (!h || n < 0 || n > 100) return SDRJFM_ERR_ARG;
	h -> set.squelch_value = n; return SDRJFM_OK;                       // set_squelchValue (:213-215)
}
// airspy: the handler picks the device rate closest to 2.0 MS/s and converts every millisecond of
// it to 2304 samples by linear interpolation (airspy-handler.cpp:108-128); tables as built there
static int lane_set_native_rate (Lane *h, int32_t hz) {
	if (!h || hz < 1000 * kAirspyOut / 2 || hz > 20000000 || hz % 1000) return SDRJFM_ERR_ARG;
	if (h -> cfg.input_rate != 2304000 || h -> resample) { h -> err = "the airspy conversion delivers 2 304 000 samples/s"; return SDRJFM_ERR_UNSUPPORTED; }
	CK (cudaSetDevice (h -> cfg.device));
	CK (cudaStreamSynchronize (h -> stream));
const int32_t B = hz / 1000;
	if ((int64_t)B + kAirspyOut > h -> cfg.max_samples_per_call) { h -> err = "max_samples_per_call is smaller than one airspy block"; return SDRJFM_ERR_CAPACITY; }
std::vector<int16_t> mi (kAirspyOut); std::vector<float> mf (kAirspyOut);
	for (int i = 0; i < kAirspyOut; i ++) {
	   float inVal = float (hz / 1000);
	   mi [i] = int (floor (i * (inVal / 2304.0)));
	   mf [i] = i * (inVal / 2304.0) - mi [i];
	}
	for (void *p : { (void *)h -> d_air_int, (void *)h -> d_air_frac, (void *)h -> d_air_pend }) if (p) cudaFree (p);
	h -> d_air_int = nullptr; h -> d_air_frac = nullptr; h -> d_air_pend = nullptr;
	CK (cudaMalloc ((void **)&h -> d_air_int, kAirspyOut * sizeof (int16_t)));
	CK (cudaMalloc ((void **)&h -> d_air_frac, kAirspyOut * sizeof (float)));
	CK (dalloc (&h -> d_air_pend, (size_t)h -> cfg.n_streams * (B + 1)));
	CK (cudaMemcpy (h -> d_air_int, mi.data (), kAirspyOut * sizeof (int16_t), cudaMemcpyHostToDevice));
	CK (cudaMemcpy (h -> d_air_frac, mf.data (), kAirspyOut * sizeof (float), cudaMemcpyHostToDevice));
	h -> air_blk = B; h -> air_pend = 0;
	return SDRJFM_OK;
}
// setTestTone (fm-processor.cpp:931-933): the burst state machine keeps its position while switched off
static int lane_set_test_tone (Lane *h, int32_t on) {
	if (!h) return SDRJFM_ERR_ARG;
	h -> tone_on = on != 0; return SDRJFM_OK;
}
// setDispDelay (:935-937): delayLine.set_delay_steps (n) restarts the line filled with (-40, -40)
static int lane_set_disp_delay (Lane *h, int32_t steps) {
	if (!h || steps < 0 || steps > kPeakRing / 2) return SDRJFM_ERR_ARG;
	h -> peak_delay = steps;
	h -> peak_e_set = (h -> fm_total / kRsDecim) / h -> peak_block;
	return SDRJFM_OK;
}
// the (left dB, right dB) pairs showPeakLevel was emitted with during the LAST process call, oldest first
static int64_t lane_read_peak_levels (Lane *h, int32_t stream, float *out, int64_t cap_pairs) {
	if (!h || !out || stream < 0 || stream >= h -> cfg.n_streams) return SDRJFM_ERR_ARG;
int64_t e0 = h -> peak_e0, e1 = h -> peak_e1;
	if (e1 - e0 > kPeakRing - h -> peak_delay) e0 = e1 - (kPeakRing - h -> peak_delay);     // older raw read-outs left the ring
int64_t n = e1 - e0;
	if (n > cap_pairs) { e0 = e1 - cap_pairs; n = cap_pairs; }
	if (n <= 0) return 0;
	CK (cudaSetDevice (h -> cfg.device));
std::vector<float2> ring (kPeakRing);
	CK (cudaMemcpyAsync (ring.data (), h -> d_peak_ring + (size_t)stream * kPeakRing, kPeakRing * sizeof (float2),
	                     cudaMemcpyDeviceToHost, h -> stream));
	CK (cudaStreamSynchronize (h -> stream));
	for (int64_t E = e0; E < e1; E ++) {
	   const int64_t idx = E - h -> peak_delay;
	   const float2 v = (idx >= h -> peak_e_set && idx >= 0) ? ring [idx & (kPeakRing - 1)] : make_float2 (-40.0f, -40.0f);
	   out [2 * (E - e0)] = v.x; out [2 * (E - e0) + 1] = v.y;
	}
	return n;
}
static int lane_set_auto_mono (Lane *h, int32_t on) {
	if (!h) return SDRJFM_ERR_ARG;
	h -> set.auto_mono = on != 0; return SDRJFM_OK;
}
static int lane_set_pss_mode (Lane *h, int32_t on) {
	if (!h) return SDRJFM_ERR_ARG;
	h -> set.pss_on = on != 0; return SDRJFM_OK;
}
static int lane_set_dc_remove (Lane *h, int32_t on) {
	if (!h) return SDRJFM_ERR_ARG;
	h -> set.dc_remove = on != 0;
//	setDCRemove also zeroes RfDC (:917-920): clear the DC fields of every stream
	CK (cudaSetDevice (h -> cfg.device));
	CK (cudaMemset2DAsync (h -> d_state, sizeof (StreamState), 0, 2 * sizeof (double),
	                       h -> cfg.n_streams, h -> stream));
	return SDRJFM_OK;
}
static int lane_trigger_frequency_change (Lane *h) {
	if (!h) return SDRJFM_ERR_ARG;
	h -> fade_cnt = h -> fade_max;                                        // :848
	h -> spec_refresh = true;                                             // new_lfSpectrum (:853)
	return lane_restart_pss_analyzer (h);
}
static int lane_restart_pss_analyzer (Lane *h) {
	if (!h) return SDRJFM_ERR_ARG;
//	pilotDelayPSS = 0; pPSS.reset () (:857-860): zero the six PSS fields of every stream
	CK (cudaSetDevice (h -> cfg.device));
	CK (cudaMemset2DAsync ((char *)h -> d_state + offsetof (StreamState, pss_delay), sizeof (StreamState), 0,
	                       offsetof (StreamState, pss_inp) - offsetof (StreamState, pss_delay),
	                       h -> cfg.n_streams, h -> stream));
	return SDRJFM_OK;
}

static int64_t lane_tables_nbytes (const Lane *h) {
	return h ? (int64_t)h -> tables.bytes.size () : SDRJFM_ERR_ARG;
}
static int lane_tables_export (const Lane *h, void *out, int64_t cap) {
	if (!h || !out || cap < (int64_t)h -> tables.bytes.size ()) return SDRJFM_ERR_ARG;
	memcpy (out, h -> tables.bytes.data (), h -> tables.bytes.size ());
	return SDRJFM_OK;
}
static int lane_tables_import (Lane *h, const void *blob, int64_t nbytes) {
	if (!h || !blob || nbytes < (int64_t)sizeof (TableHeader)) return SDRJFM_ERR_ARG;
const TableHeader *th = (const TableHeader *)blob;
	if (th -> magic != 0x54464A53u || th -> payload_floats < 0 ||
	    nbytes != (int64_t)(sizeof (TableHeader) + th -> payload_floats * sizeof (float)) ||
	    th -> input_rate != h -> cfg.input_rate || th -> fm_rate != h -> cfg.fm_rate)
	   return SDRJFM_ERR_ARG;
//	the blob comes from another rank: nothing in it is trusted.  It must describe THIS handle's
//	configuration (the kernels are chosen from the handle's settings) and every table must lie
//	inside the payload.
const TableBlob &mine = h -> tables;
const TableHeader &my = mine.hdr ();
	if (th -> version != my.version || th -> ncomp != my.ncomp || th -> ncomp > 96 || th -> ncomp < 1 ||
	    th -> ntaps1 != my.ntaps1 || th -> ntaps2 != my.ntaps2 || th -> decim1 != my.decim1 || th -> decim2 != my.decim2 ||
	    th -> ncomp_wide != my.ncomp_wide || th -> payload_floats != my.payload_floats ||
	    th -> input_filter_hz != h -> set.input_filter_hz ||
	    th -> rs_L != my.rs_L || th -> rs_M != my.rs_M || th -> rs_P != my.rs_P || th -> rs_ntapsA != my.rs_ntapsA) {
	   h -> err = "table blob does not match this handle's rates / filter settings"; return SDRJFM_ERR_ARG;
	}
	{  // same designer, same configuration: the layout must be the one this handle computed itself
	   const int64_t *a = &th -> off_fmband1, *b = &my.off_fmband1;
	   const int noff = (int)((&th -> off_comp_wide - &th -> off_fmband1) + 1);
	   for (int i = 0; i < noff; i ++) if (a [i] != b [i] || a [i] < 0 || a [i] > th -> payload_floats) {
	      h -> err = "table blob layout differs from this build's"; return SDRJFM_ERR_ARG;
	   }
	   if (th -> off_rsA != my.off_rsA || th -> off_rsB != my.off_rsB || th -> off_squelch != my.off_squelch ||
	       th -> off_rds_sym != my.off_rds_sym) { h -> err = "table blob layout differs from this build's"; return SDRJFM_ERR_ARG; }
	}
	CK (cudaSetDevice (h -> cfg.device));
	CK (cudaStreamSynchronize (h -> stream));
	h -> tables.bytes.assign ((const unsigned char *)blob, (const unsigned char *)blob + nbytes);
	return upload_tables (h);
}

