// K6p — what fmProcessor::run does with every working-rate (48 kHz) PCM sample behind the fade-in
// (src/fm/fm-processor.cpp:636-647):
//   insertTestTone      :800-823   applied inside audio_kernel (audio_out.cuh) from the table built here
//   evaluatePeakLevel   :772-798   peak_kernel: peak of |left|, |right| per 961 samples -> dB
//   sendSampletoOutput  :825-838   audio_convert_kernel: the second newConverter when audioRate != workingRate
//
// The second converter is libsamplerate in the reference (newconverter.cpp:26-80), like the 192 -> 48 kHz step:
// PARITY UNPINNED.  What runs here is a documented rational polyphase windowed-sinc converter
// of our own, checked against a float64 model of the same taps.
#pragma once
#include "common.cuh"

namespace sdrjfm {

// ---- test tone ----------------------------------------------------------------------------------
// insertTestTone is a small state machine per PCM sample: while enabled every sample is scaled by
// (1 - 0.9); TimePeriodCounter counts the samples without tone and, when it exceeds workingRate * 2.0,
// arms NoSamplRemain = workingRate * 0.025 tone samples; CurPhase restarts at 0 with every burst, so the
// burst is ONE fixed sequence of floats: built on the host with the reference's own statement order
// (float accumulate, PI_Constrain in double, sinf) and indexed by the position in the cycle on the device.
struct ToneParams {
	int32_t on;
	int32_t arm;               // samples of a cycle before the burst: (uint32) counter > workingRate * TimePeriod first holds at arm
	int32_t burst;             // NoSamplRemain = workingRate * SignalDuration
	int64_t pos;               // position in the cycle of the first output sample of this call
};

inline void tone_design (int32_t working_rate, int32_t &arm, std::vector<float> &table) {
const float toneFreqHz = 1000.0f, TimePeriod = 2.0f, SignalDuration = 0.025f;      // fm-processor.cpp:801, fm-processor.h:243-244
uint32_t c = 0;
	while (!((float)(++ c) > working_rate * TimePeriod)) ;           // :816-817 (uint32 compared as float)
	arm = (int32_t)c;
const uint32_t burst = working_rate * SignalDuration;                    // :819
	table.resize (burst);
float CurPhase = 0.0f;
const float PhaseIncr = 2 * M_PI / working_rate * toneFreqHz;            // :821
	for (uint32_t k = 0; k < burst; k ++) {
	   CurPhase += PhaseIncr;                                                // :810
	   if (!(0 <= CurPhase && CurPhase < 2 * M_PI)) {                        // PI_Constrain, fm-constants.h:148-158
	      if (CurPhase >= 2 * M_PI) CurPhase = fmod (CurPhase, 2 * M_PI);
	      else if (CurPhase > -2 * M_PI) CurPhase = CurPhase + 2 * M_PI;
	      else CurPhase = 2 * M_PI - fmod (-CurPhase, 2 * M_PI);
	   }
	   table [k] = sinf (CurPhase);                                          // :812 (std::sin (float))
	}
}

__device__ __forceinline__ float2 tone_apply (const ToneParams &T, const float *__restrict__ tab, float2 s, int64_t k) {
const float level = 0.9f;
const float keep = fsub (1.0f, level);
	s.x = fmul (s.x, keep); s.y = fmul (s.y, keep);                          // ioS *= (1.0f - level)
const int64_t pos = (T.pos + k) % (int64_t)(T.arm + T.burst);
	if (pos >= T.arm) {
	   const float t = fmul (level, tab [pos - T.arm]);                       // ioS += level * DSPCOMPLEX (smpl, smpl)
	   s.x = fadd (s.x, t); s.y = fadd (s.y, t);
	}
	return s;
}

// ---- peak level meter -----------------------------------------------------------------------------
// peakLevelSampleMax = workingRate / 50 (:142) and the counter test is `> max` (:782): one read-out per
// kPeakBlock = 961 PCM samples, counted from the start of the processor.  Block E covers output samples
// 961 E .. 961 E + 960.  Raw (left dB, right dB) pairs go into a ring by E; the display delay line
// (delayLine, setDispDelay) is an index shift applied where the pairs are read.
constexpr int kPeakRing = 1024;

// pcm: [S][pitch] working-rate samples of this call (q0 = global index of the first, nq of them)
// carry_in / carry_out: which of StreamState::peak_carry holds the running maxima of the open block
__global__ void __launch_bounds__ (32)
peak_kernel (const float2 *__restrict__ pcm, int64_t pitch, int64_t q0, int32_t nq, int32_t block,
             StreamState *__restrict__ state, int sel, float2 *__restrict__ ring) {
const int stream = blockIdx.y, lane = threadIdx.x;
const int64_t E = q0 / block + blockIdx.x;
const int64_t lo = max (E * block, q0), hi = min ((E + 1) * (int64_t)block, q0 + nq);
const float2 *p = pcm + (int64_t)stream * pitch;
float ml = 0.f, mr = 0.f;
	for (int64_t q = lo + lane; q < hi; q += 32) {
	   const float2 v = p [q - q0];
	   ml = fmaxf (ml, fabsf (v.x)); mr = fmaxf (mr, fabsf (v.y));
	}
#pragma unroll
	for (int k = 16; k >= 1; k >>= 1) {
	   ml = fmaxf (ml, __shfl_xor_sync (0xffffffffu, ml, k));
	   mr = fmaxf (mr, __shfl_xor_sync (0xffffffffu, mr, k));
	}
	if (lane != 0) return;
StreamState &st = state [stream];
	if (E * block < q0) {                      // the block was opened by an earlier call
	   ml = fmaxf (ml, st.peak_carry [sel][0]); mr = fmaxf (mr, st.peak_carry [sel][1]);
	}
	if ((E + 1) * (int64_t)block <= q0 + nq) {  // complete: one showPeakLevel read-out
	   const float l = ml > 0.0f ? fmul (20.0f, log10f (ml)) : -40.0f;
	   const float r = mr > 0.0f ? fmul (20.0f, log10f (mr)) : -40.0f;
	   ring [(int64_t)stream * kPeakRing + (E & (kPeakRing - 1))] = make_float2 (l, r);
	   if ((E + 1) * (int64_t)block == q0 + nq) { st.peak_carry [sel ^ 1][0] = 0.f; st.peak_carry [sel ^ 1][1] = 0.f; }
	}
	else { st.peak_carry [sel ^ 1][0] = ml; st.peak_carry [sel ^ 1][1] = mr; }
}

// ---- second converter: working rate -> audio rate, rational L / M ---------------------------------
// out [k] = sum_{j < P} h [(k M mod L) + j L] in [floor (k M / L) - j]; output k exists once input
// floor (k M / L) does, i.e. ceil (T L / M) outputs after T inputs.  h: Blackman-windowed sinc at the rate
// L x working, cut-off 0.45 min (working, audio), DC gain L (unit gain per phase on average).
constexpr int kCvTapsPerPhase = 32;

inline bool convert_design (int32_t working_rate, int32_t audio_rate, int &L, int &M, std::vector<float> &h) {
int a = audio_rate, b = working_rate;
	while (b) { const int t = a % b; a = b; b = t; }
	L = audio_rate / a; M = working_rate / a;
	if (L < 1 || L > 640 || M < 1 || M > 4096) return false;
const int P = kCvTapsPerPhase, N = L * P;
const double fc = 0.45 * std::min (working_rate, audio_rate) / ((double)L * working_rate);     // cycles per sample at L x working
std::vector<double> d (N);
	for (int i = 0; i < N; i ++) {
	   const double t = i - (N - 1) / 2.0;
	   const double s = t == 0.0 ? 2 * fc : sin (2 * M_PI * fc * t) / (M_PI * t);
	   const double w = 0.42 - 0.5 * cos (2 * M_PI * i / (N - 1)) + 0.08 * cos (4 * M_PI * i / (N - 1));
	   d [i] = s * w;
	}
//	every polyphase branch normalised to unit DC gain: no level ripple between output phases
	h.assign (N, 0.f);
	for (int p = 0; p < L; p ++) {
	   double sum = 0;
	   for (int j = 0; j < P; j ++) sum += d [p + j * L];
	   for (int j = 0; j < P; j ++) h [p + j * L] = (float)(d [p + j * L] / sum);
	}
	return true;
}

struct ConvertParams {
	int32_t L, M, P;
	int64_t k0;                // global index of the first output of this call
	int32_t nk;                // outputs of this call
	int64_t t0;                // global index of the first input sample of this call
	int32_t nt;                // inputs of this call
};

// in: [S][in_pitch] this call's working-rate samples; hist: [S][P] the P samples before them (oldest first);
// out: [S][out_pitch]
__global__ void __launch_bounds__ (256)
audio_convert_kernel (const float2 *__restrict__ in, int64_t in_pitch, const float2 *__restrict__ hist,
                      const float *__restrict__ taps, ConvertParams C, float2 *__restrict__ out, int64_t out_pitch) {
const int stream = blockIdx.y;
const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= C.nk) return;
const int64_t k = C.k0 + i;
const int64_t n = (k * C.M) / C.L;
const int ph = (int)((k * C.M) % C.L);
const float2 *is = in + (int64_t)stream * in_pitch;
const float2 *hs = hist + (int64_t)stream * C.P;
float2 acc = make_float2 (0.f, 0.f);
	for (int j = 0; j < C.P; j ++) {
	   const int64_t m = n - j - C.t0;                       // local input index
	   const float2 v = m >= 0 ? is [m] : (m >= -C.P ? hs [C.P + m] : make_float2 (0.f, 0.f));
	   acc = ffma2 (taps [ph + j * C.L], v, acc);
	}
	out [(int64_t)stream * out_pitch + i] = acc;
}

// the last P inputs of (hist | in) for the next call
__global__ void convert_roll_kernel (const float2 *__restrict__ in, int64_t in_pitch, int32_t nt, int P,
                                     const float2 *__restrict__ hist, float2 *__restrict__ new_hist) {
const int stream = blockIdx.x, i = threadIdx.x;
	if (i >= P) return;
const int m = nt - P + i;
	new_hist [(int64_t)stream * P + i] = m >= 0 ? in [(int64_t)stream * in_pitch + m] : hist [(int64_t)stream * P + P + m];
}

}	// namespace sdrjfm
