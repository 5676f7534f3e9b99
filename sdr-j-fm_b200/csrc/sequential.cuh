// K3 — the per-stream SEQUENTIAL stage: one GPU lane per IQ stream steps through the
// fm-rate samples and runs, bit for bit, the recurrences that cannot be split in time
// (a re-started pilot PLL never re-converges closer than ~2e-4 rad):
//   fm_afc / am_carr_ampl one-poles and the output scaling   fm-demodulator.cpp:130-131,197-198
//   (optionally) the PLL discriminator                        pllC.cpp:67-90
//   the 19 kHz pilot PLL and its lock detector                pilot-recover.cpp:54-83
// Everything that is memoryless has been done by K2; this kernel is latency-bound by the
// pilot recurrence (NCO table look-up -> phase update -> wrap), so the sine table is kept
// in shared memory as ONE quarter wave (48001 floats = 188 KB): the reference table
// (sincos.cpp:41-43) is bit-identical to its quarter-wave reflection except at the three
// zero crossings, which are patched from an exception list (checked on the host when the
// tables are built).
#pragma once
#include "common.cuh"
#include "discriminator.cuh"

namespace sdrjfm {

struct SinLut {
	const float *q;          // quarter wave, Rate/4 + 1 entries (global copy)
	int32_t rate, quarter;
	int32_t sin_exc_idx [kMaxSinExc]; float sin_exc_val [kMaxSinExc];
	int32_t cos_exc_idx [kMaxSinExc]; float cos_exc_val [kMaxSinExc];
	double  C;               // Rate / (2 pi), sincos.cpp:45
	float   sin_at_half;     // table value at idx = Rate / 2: the only sine entry the reflection can miss
};

// Table value sin (2 pi idx / Rate) for 0 <= idx < Rate from the quarter wave `q`.
// Every exception sits where the reflected index is 0 (the zero crossings), so the
// exception list is only consulted there.
// (the fm rate is a compile-time constant of the device code: 192000 is the only rate the
// reference really supports; the host refuses anything else)
constexpr int32_t kFmRate = 192000;
__device__ int32_t phase_index_rare (float phase);

__device__ __forceinline__ float lut_sin_idx (const SinLut &L, const float *q, int32_t idx) {
constexpr int32_t Q = kFmRate / 4, H = 2 * Q;
int32_t k = idx >= H ? idx - H : idx;          // [0, H)
	k = k > Q ? H - k : k;                         // [0, Q]
float v = q [k];
	v = idx >= H ? -v : v;
	if (k == 0 && idx != 0) {
#pragma unroll
	   for (int e = 0; e < kMaxSinExc; e ++)
	      if (idx == L.sin_exc_idx [e]) v = L.sin_exc_val [e];
	}
	return v;
}

__device__ __forceinline__ float lut_cos_idx (const SinLut &L, const float *q, int32_t idx) {
constexpr int32_t Q = kFmRate / 4, H = 2 * Q;
int32_t s = idx + Q;
	if (s >= kFmRate) s -= kFmRate;
int32_t k = s >= H ? s - H : s;
	k = k > Q ? H - k : k;
float v = q [k];
	v = s >= H ? -v : v;
	if (k == 0) {
#pragma unroll
	   for (int e = 0; e < kMaxSinExc; e ++)
	      if (idx == L.cos_exc_idx [e]) v = L.cos_exc_val [e];
	}
	return v;
}

// SinCos::fromPhasetoIndex for Phase >= 0 (sincos.cpp:54-56): int32 (Phase * C) % Rate
__device__ __forceinline__ int32_t phase_index (const SinLut &L, float phase) {
int32_t i = (int32_t)((double)phase * (kFmRate / (2 * M_PI)));
	if (i >= kFmRate) i -= kFmRate;                // phases here are < 4 pi ...
	if (i >= kFmRate || i < 0) i = phase_index_rare (phase);   // ... anything else: exact slow path
	return i;
}

// SinCos::getSin, sincos.cpp:75-79
__device__ __forceinline__ float lut_getSin (const SinLut &L, const float *q, float phase) {
	if (phase < 0.f) return -lut_sin_idx (L, q, phase_index (L, -phase));
	return lut_sin_idx (L, q, phase_index (L, phase));
}

// the phase normalisation shared by SinCos::getCos / getComplex, sincos.cpp:81-91
__device__ __forceinline__ int32_t cos_phase_index (const SinLut &L, float phase) {
	while (phase < 0.f) phase = (float)((double)phase + 2 * M_PI);
	phase = (float)fmod ((double)phase, 2 * M_PI);
	return phase_index (L, phase);
}

// PI_Constrain, includes/fm-constants.h:148-158.  The reference compares the float against
// the DOUBLE 2*M_PI; kTwoPiF is the float just above that double and there is no float in
// between, so the float comparisons below select exactly the same branch.  fmod (v, 2 pi)
// for 2 pi <= v < 4 pi is the exact difference v - 2 pi (Sterbenz), the common wrap.
__device__ __noinline__ float pi_constrain_rare (float val) {
const float kTwoPiF = 6.2831855f;
const double v = (double)val;
	if (val >= kTwoPiF) return (float)fmod (v, 2 * M_PI);
	if (val > -kTwoPiF) return (float)(v + 2 * M_PI);
	return (float)(2 * M_PI - fmod (-v, 2 * M_PI));
}

__device__ __forceinline__ float pi_constrain (float val) {
const float kTwoPiF = 6.2831855f;
	if (val >= 0.f && val < kTwoPiF) return val;
	if (val >= kTwoPiF && val < 12.566370f) return (float)((double)val - 2 * M_PI);
	return pi_constrain_rare (val);
}

// rare paths of SinCos index arithmetic, kept out of line so the hot loop stays small
__device__ __noinline__ int32_t phase_index_rare (float phase) {
int32_t i = (int32_t)((double)phase * (kFmRate / (2 * M_PI)));
	return i % kFmRate;
}

struct SeqParams {
	float   K_FM;             // fm-demodulator.cpp:64
	float   omega, gain;      // pilotRecovery ctor args, fm-processor.cpp:78-80
	int32_t lock_half_rate;   // Rate_in >> 1
	int32_t decoder;
	// pllC constants (fm-demodulator.cpp:67-72 -> pllC.cpp:43-58)
	float   pll_beta, pll_lo, pll_hi, pll_reset;
	int32_t n_streams;
};

constexpr int kSeqLanes = 32;

// ---- squelch (src/various/squelchClass.cpp; in the chain at fm-processor.cpp:499-510) -----------
// Noise squelch: the demodulated sample goes through a 20th-order Chebyshev high-pass and low-pass
// (10 biquads each, Basic_IIR::Pass (float), includes/various/iir-filters.h:90-104), the magnitudes
// are averaged (one-pole, weight fs / 100) and every holdPeriod = fs / 20 samples the two averages
// decide; level squelch: the carrier level decides.  A suppressed sample is multiplied by 0.
// These are per-sample float recurrences with poles next to the unit circle: they are restated
// operation by operation in the lane-per-stream kernel (no re-association).
constexpr int kSqQuads = 10;
struct SquelchParams {
	int32_t mode;                       // 0 OFF, 1 NSQ, 2 LSQ (fm-processor.h:87)
	int32_t hold;                       // holdPeriod
	float   thr_noise, thr_level;       // setSquelchLevel, squelchClass.cpp:34-38
	float   weight;                     // sampleRate / 100
	float   hp [1 + 4 * kSqQuads];      // gain, then (A1 A2 B1 B2) per biquad
	float   lp [1 + 4 * kSqQuads];
};
struct SquelchState {                  // squelch members (squelchClass.h:41-52) + the biquad memories
	float   hp_m1 [kSqQuads], hp_m2 [kSqQuads], lp_m1 [kSqQuads], lp_m2 [kSqQuads];
	float   avg_high, avg_low;
	int32_t count, suppress;
};

template <int NQ>
__device__ __forceinline__ float iir_pass (const float *coef, float *m1, float *m2, float v) {
float o = fmul (v, coef [0]);
#pragma unroll
	for (int i = 0; i < NQ; i ++) {
	   const float A1 = coef [1 + 4 * i], A2 = coef [2 + 4 * i], B1 = coef [3 + 4 * i], B2 = coef [4 + 4 * i];
	   const float w = fsub (fsub (o, fmul (m1 [i], B1)), fmul (m2 [i], B2));
	   o = fadd (fadd (w, fmul (m1 [i], A1)), fmul (m2 [i], A2));
	   m2 [i] = m1 [i];
	   m1 [i] = w;
	}
	return o;
}
__device__ __forceinline__ float decaying_average (float old, float input, float weight) {      // :40-45
	if (weight <= 1) return input;
	return (float)((double)input * (1.0 / (double)weight) + (double)old * (1.0 - (1.0 / (double)weight)));
}
__device__ __forceinline__ float squelch_step (const SquelchParams &Q, SquelchState &s, float sample, float carrier) {
const float hystN = 0.001f;             // SQUELCH_HYSTERESIS_NSQ; _LSQ = 0; LEVELREDUCTIONFACTOR = 0
	if (Q.mode == 1) {
	   const float v1 = fabsf (iir_pass<kSqQuads> (Q.hp, s.hp_m1, s.hp_m2, sample));
	   const float v2 = fabsf (iir_pass<kSqQuads> (Q.lp, s.lp_m1, s.lp_m2, sample));
	   s.avg_high = decaying_average (s.avg_high, v1, Q.weight);
	   s.avg_low  = decaying_average (s.avg_low, v2, Q.weight);
	   if (++ s.count >= Q.hold) {
	      s.count = 0;
	      if (Q.thr_noise < hystN) s.suppress = 1;
	      else if (s.avg_high < fsub (fmul (s.avg_low, Q.thr_noise), hystN)) s.suppress = 0;
	      else if (s.avg_high >= fadd (fmul (s.avg_low, Q.thr_noise), hystN)) s.suppress = 1;
	   }
	}
	else {
	   if (++ s.count >= Q.hold) {
	      s.count = 0;
	      if (carrier < fsub (Q.thr_level, 0.0f)) s.suppress = 1;
	      else if (carrier >= fadd (Q.thr_level, 0.0f)) s.suppress = 0;
	   }
	}
	return s.suppress ? fmul (sample, 0.0f) : sample;
}

struct SeqCarry {      // loop-carried registers of one stream
	float fm_afc, am, phase, oldv, plock, nco, incr;
	int   locked, stable;
};

// DEC: 0 = discriminator output computed by K2 (Mixed, complex/real baseband delay, difference),
//      1 = PLL decoder (pllC on the normalised sample), 2 = AM decoder (pllC on the raw sample for
//      the AFC read-out, audio from the envelope: fm_Demodulator::decodeAM, fm-demodulator.cpp:215-241)
template <int DEC, bool SQ>
__device__ __forceinline__ void seq_step (const SeqParams &P, const SinLut &L, const float *q,
                                          const float *atanPPY, SeqCarry &c, float res, float zAbs,
                                          float2 nqv, const SquelchParams &Q, SquelchState &qs,
                                          float &demod_o, float &phase_o, uint8_t &lock_o) {
const float carrierAlpha = 0.0010f, fmDcAlpha = 0.0001f;          // fm-demodulator.cpp:115-117
const float oneMinusCarrier = fsub (1.0f, carrierAlpha);
const float oneMinusDc = fsub (1.0f, fmDcAlpha);
const float lockAlpha = 1.0f / 3000.0f;                           // pilot-recover.cpp:57
const double oneMinusLock = 1.0 - (double)lockAlpha;
	c.am = fadd (fmul (oneMinusCarrier, c.am), fmul (carrierAlpha, zAbs));
	if (DEC != 0) {
//	pllC::do_pll on the normalised (AM: the raw) sample, pllC.cpp:67-90
	   const int32_t ci = cos_phase_index (L, c.nco);
	   const float2 osc = make_float2 (lut_cos_idx (L, q, ci), lut_sin_idx (L, q, ci));
	   const float2 d = cmul_rn (make_float2 (osc.x, -osc.y), nqv);
	   const float perr = lut_atan2 (atanPPY, d.y, d.x);
	   c.incr = fadd (fmul (fsub (1.0f, P.pll_beta), perr), fmul (P.pll_beta, c.incr));
	   if (c.incr < P.pll_lo || c.incr > P.pll_hi) c.incr = P.pll_reset;
	   c.nco = fadd (c.nco, c.incr);
	   if ((double)c.nco >= 2 * M_PI) c.nco = (float)fmod ((double)c.nco, 2 * M_PI);
	   else while (c.nco < 0.f) c.nco = (float)((double)c.nco + 2 * M_PI);
	   res = c.incr;
	}
	c.fm_afc = fadd (fmul (oneMinusDc, c.fm_afc), fmul (fmDcAlpha, res));
float demod = fdiv (fmul (fmul (20.0f, fsub (res, c.fm_afc)), 1.0f), P.K_FM);
	if (DEC == 2) {
//	envelope minus carrier level, normalised to the carrier level, limited to +-1 (:230-241)
	   const float gainLimit = 0.01f;
	   demod = fdiv (fsub (zAbs, c.am), c.am < gainLimit ? gainLimit : c.am);
	   demod = demod > 1.0f ? 1.0f : (demod < -1.0f ? -1.0f : demod);
	}
	if (SQ) demod = squelch_step (Q, qs, demod, c.am);               // fm-processor.cpp:499-510
//	pilotRecovery::getPilotPhase (5 * demod), pilot-recover.cpp:54-83
const float pilot = fmul (5.0f, demod);
const float osc = lut_getSin (L, q, c.phase);
const float perr = fmul (pilot, osc);
	c.phase = fadd (c.phase, fmul (perr, P.gain));
const float cur = pi_constrain (c.phase);
	c.phase = pi_constrain (fadd (c.phase, P.omega));
const float quad = fdiv (fsub (osc, c.oldv), P.omega);
	c.oldv = osc;
	c.plock = (float)((double)fmul (lockAlpha, fmul (-quad, pilot)) + (double)c.plock * oneMinusLock);
	if (c.plock > 0.07f) {
	   if (c.locked || ++c.stable > P.lock_half_rate) c.locked = 1;
	}
	else { c.locked = 0; c.stable = 0; }
	demod_o = demod; phase_o = cur; lock_o = (uint8_t)c.locked;
}

// res_raw, zabs, iqn : K2 outputs.  demod / pilot_phase / locked : fm-rate outputs.
// One lane per stream; the quarter-wave sine table lives in shared memory.
template <int DEC, bool SQ>
__global__ void __launch_bounds__ (kSeqLanes)
sequential_kernel (const float *__restrict__ res_raw, const float *__restrict__ zabs,
                   const float2 *__restrict__ iqn, int64_t pitch, int32_t M,
                   const SeqParams P, const SinLut L, const float *__restrict__ atanPPY,
                   StreamState *__restrict__ state,
                   float *__restrict__ demod_out, float *__restrict__ phase_out,
                   uint8_t *__restrict__ locked_out,
                   const SquelchParams Q, SquelchState *__restrict__ sqstate) {
extern __shared__ float sq [];
	for (int i = threadIdx.x; i <= L.quarter; i += blockDim.x) sq [i] = L.q [i];
	__syncthreads ();
const int stream = blockIdx.x * blockDim.x + threadIdx.x;
	if (stream >= P.n_streams) return;
StreamState &st = state [stream];
const float *rr = res_raw + (int64_t)stream * pitch;
const float *za = zabs + (int64_t)stream * pitch;
const float2 *nq = iqn + (int64_t)stream * pitch;
float *dm = demod_out + (int64_t)stream * pitch;
float *ph = phase_out + (int64_t)stream * pitch;
uint8_t *lk = locked_out + (int64_t)stream * pitch;

SeqCarry c;
	c.fm_afc = st.fm_afc; c.am = st.am_carr_ampl;
	c.phase = st.pilot_phase; c.oldv = st.pilot_old; c.plock = st.pilot_lock;
	c.locked = st.pilot_locked; c.stable = st.pilot_stable_cnt;
	c.nco = st.pll_nco_phase; c.incr = st.pll_phase_incr;
SquelchState qs;
	if (SQ) qs = sqstate [stream];

int32_t m = 0;
	for (; m + 4 <= M; m += 4) {
	   const float4 r4 = *reinterpret_cast<const float4 *>(rr + m);
	   const float4 z4 = *reinterpret_cast<const float4 *>(za + m);
	   float2 n4 [4];
	   if (DEC != 0) {
#pragma unroll
	      for (int k = 0; k < 4; k ++) n4 [k] = nq [m + k];
	   }
	   const float rv [4] = { r4.x, r4.y, r4.z, r4.w }, zv [4] = { z4.x, z4.y, z4.z, z4.w };
	   float d4 [4], p4 [4]; uint8_t l4 [4];
#pragma unroll
	   for (int k = 0; k < 4; k ++)
	      seq_step<DEC, SQ> (P, L, sq, atanPPY, c, rv [k], zv [k],
	                         DEC != 0 ? n4 [k] : make_float2 (0.f, 0.f), Q, qs, d4 [k], p4 [k], l4 [k]);
	   *reinterpret_cast<float4 *>(dm + m) = make_float4 (d4 [0], d4 [1], d4 [2], d4 [3]);
	   *reinterpret_cast<float4 *>(ph + m) = make_float4 (p4 [0], p4 [1], p4 [2], p4 [3]);
	   *reinterpret_cast<uchar4 *>(lk + m) = make_uchar4 (l4 [0], l4 [1], l4 [2], l4 [3]);
	}
	for (; m < M; m ++) {
	   float d, p; uint8_t l;
	   seq_step<DEC, SQ> (P, L, sq, atanPPY, c, rr [m], za [m],
	                      DEC != 0 ? nq [m] : make_float2 (0.f, 0.f), Q, qs, d, p, l);
	   dm [m] = d; ph [m] = p; lk [m] = l;
	}
	st.fm_afc = c.fm_afc; st.am_carr_ampl = c.am;
	st.pilot_phase = c.phase; st.pilot_old = c.oldv; st.pilot_lock = c.plock;
	st.pilot_locked = c.locked; st.pilot_stable_cnt = c.stable;
	st.pll_nco_phase = c.nco; st.pll_phase_incr = c.incr;
	if (SQ) sqstate [stream] = qs;
}

}	// namespace sdrjfm
