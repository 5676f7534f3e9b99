// Host-side design of every tap set and look-up table the FM path needs.
//
// These are one-off, host-evaluated formulas; they follow the reference's float/double
// promotion exactly so that the device tables are bit-identical
// to what the reference's constructors build:
//   windowed-sinc designers   src/various/fir-filters.cpp:41-62 (LowPassFIR),
//                             :197-222 (BandPassFIR), :327-347 (DecimatingFIR)
//   SinCos table              src/various/sincos.cpp:36-45
//   Oscillator table          src/various/oscillator.cpp:26-37
//   compAtan tables           src/various/Xtan2.cpp:12-39
//   FFT twiddles              src/various/fft-complex.cpp:65-71
//   fftFilter spectra         src/various/fft-filters.cpp:71-95 (setBand / setLowPass)
// The result is ONE contiguous blob (header + float payload) so that rank 0 can design it
// once and broadcast it to the other GPUs.
#pragma once
#include <cstdint>
#include <complex>
#include <vector>

namespace sdrjfm {

typedef std::complex<float> cf32;

constexpr int   kAtanSize      = 8192;          // Xtan2.cpp:9
constexpr int   kAtanTables    = 8;
constexpr int   kCompositeTaps = 37;            // 25 + 6*(3-1)
constexpr int   kRdsFftSize    = 32768;         // FFT_SIZE, fm-constants.h:107
constexpr int   kRdsDegree     = 768;           // PILOTFILTER_SIZE, fm-constants.h:106
constexpr int   kPssFftSize    = 2048;          // stereo-separation.cpp:31
constexpr int   kPssDegree     = 295;
constexpr int   kAudioFftSize  = 8192;          // fm-processor.cpp:76
constexpr int   kAudioDegree   = 756;
constexpr int   kInputFftSize  = 65536;         // fm-processor.cpp:77
constexpr int   kInputDegree   = 251;
constexpr int   kRdsDecimTaps  = 11;            // fm-processor.cpp:382
constexpr int   kArcsineSize   = 4 * 8192;      // fm-demodulator.cpp:73

// offsets are in floats from the start of the payload
struct TableHeader {
	uint32_t magic;           // 'SJFT'
	uint32_t version;
	int32_t  input_rate, fm_rate;
	int32_t  decim1, decim2;  // 6 and 2 at 2304000
	int32_t  ntaps1, ntaps2;  // 25 and 3
	int32_t  ncomp;           // composite real taps (ntaps1 + decim1*(ntaps2-1))
	int32_t  input_filter_hz, audio_lp_hz;
	int32_t  reserved0;
	int64_t  payload_floats;
	int64_t  off_fmband1, off_fmband2, off_rdsdecim;     // complex taps
	int64_t  off_comp;          // ncomp real composite taps C'[i] = C[i] + alpha g[i] (DC removal folded in)
	int64_t  off_comp_consts;   // [sumC, sumC', Gre, Gim, K_FM, alpha*gbar0, alpha*gbar1, alpha*gbar2]
	int64_t  off_atan;          // 8 * 8193 floats, order PPY PPX PNY PNX NPY NPX NNY NNX
	int64_t  off_sincos;        // fm_rate complex (cos, sin)
	int64_t  off_arcsine;       // 32769 floats
	int64_t  off_tw2048, off_tw8192, off_tw32768;        // n/2 complex each
	int64_t  off_pss_lp, off_rds_bp, off_audio_lp;       // frequency-domain filter vectors
	int64_t  off_input_taps;    // 251 real time-domain taps (0 if input filter off)
	int64_t  off_comp_wide;     // composite of input filter and the decimator cascade
	int32_t  ncomp_wide;
	int32_t  reserved1;
	// rational polyphase resampler (front_end_mode 1, resample.cuh); rs_L = 0 when the rates admit none
	int32_t  rs_L, rs_M, rs_P, rs_ntapsA;   // stage B: up L, down M, P taps per phase; stage A taps (49)
	int64_t  off_rsA;                        // rs_ntapsA real taps, unit DC gain
	int64_t  off_rsB;                        // [rs_L][rs_P] real taps, every phase unit DC gain
	int64_t  off_squelch;                    // kSquelchFloats floats, see design_squelch_iir
	int64_t  off_rds_sym;                    // kRdsSymFloats floats, see design_rds_symbol_tables
};

struct TableBlob {
	std::vector<unsigned char> bytes;      // TableHeader followed by the float payload
	const TableHeader &hdr () const { return *reinterpret_cast<const TableHeader *>(bytes.data ()); }
	TableHeader &hdr () { return *reinterpret_cast<TableHeader *>(bytes.data ()); }
	const float *payload () const { return reinterpret_cast<const float *>(bytes.data () + sizeof (TableHeader)); }
	float *payload () { return reinterpret_cast<float *>(bytes.data () + sizeof (TableHeader)); }
};

// designers (exposed for the unit tests through the C ABI's table export)
std::vector<float> design_sinc_blackman (int ntaps, float f);              // un-normalised tmp[]
std::vector<cf32>  design_decimating_lowpass (int ntaps, int32_t low, int32_t fs);
std::vector<cf32>  design_lowpass (int ntaps, int32_t fc, int32_t fs);
std::vector<cf32>  design_bandpass (int ntaps, int32_t low, int32_t high, int32_t fs);
void               fft_radix2_reference_order (cf32 *v, int n);            // same op order as fft-complex.cpp:50-102
std::vector<cf32>  fft_twiddles (int n);

// rational resampler design (resample.cuh): false when 5 * fm_rate / input_rate does not reduce to
// L / M with L <= 16 and P <= 128
bool design_resampler (int32_t input_rate, int32_t fm_rate, int &L, int &M, int &P,
                       std::vector<float> &hA, std::vector<float> &hB);

// squelch filters (squelchClass.cpp:12-21): HighPassIIR (20, 70000 - 100, fs, S_CHEBYSHEV) and
// LowPassIIR (20, 70000, fs, S_CHEBYSHEV), src/various/iir-filters.cpp.  out = 82 floats:
// high-pass gain, 10 x (A1 A2 B1 B2), low-pass gain, 10 x (A1 A2 B1 B2).
constexpr int kSquelchQuads = 10;
constexpr int kSquelchFloats = 2 * (1 + 4 * kSquelchQuads);
void design_squelch_iir (int32_t fm_rate, float *out);

// RDS symbol stage at 24 kHz, mode RDS_1 (src/rds/rds-decoder-1.cpp:45-112): out =
// match kernel [43], rdsFilter real taps [21] (LowPassFIR (21, RDS_WIDTH, rate)), sharpFilter
// = BandPassIIR (7, 1187.5 - 6, 1187.5 + 6, rate, S_BUTTERWORTH): gain, 8 x (A1 A2 B1 B2)  -> 97 floats
constexpr int kRdsMatchTaps = 43, kRdsLpTaps = 21, kRdsBpQuads = 8;
constexpr int kRdsSymFloats = kRdsMatchTaps + kRdsLpTaps + 1 + 4 * kRdsBpQuads;
void design_rds_symbol_tables (int32_t rate, float *out);

// RDS symbol stage of mode RDS_2: the 45-tap matched filter ShapingFilter ().root_raised_cosine (1.0, rate,
// 2 * 1187.5, 1.0, 45) of rds-decoder-2.cpp:67-71 (src/various/shaping_filter.cpp:4-56)
void design_rds2_matched_filter (int32_t rate, float *out /* [45] */);
void design_rds3_clock (int32_t rate, std::vector<float> &sin_tab, float *omega, int8_t *kmap /* [21] */);

TableBlob build_tables (int32_t input_rate, int32_t fm_rate, int32_t input_filter_hz,
                        int32_t audio_lp_hz);

}	// namespace sdrjfm
