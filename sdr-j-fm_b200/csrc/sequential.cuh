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
	const float *q;          // quarter wave, Rate/4 + 1 entries (shared or global memory)
	int32_t rate, quarter;
	int32_t sin_exc_idx [kMaxSinExc]; float sin_exc_val [kMaxSinExc];
	int32_t cos_exc_idx [kMaxSinExc]; float cos_exc_val [kMaxSinExc];
	double  C;               // Rate / (2 pi), sincos.cpp:45
};

__device__ __forceinline__ float lut_sin_idx (const SinLut &L, int32_t idx) {
#pragma unroll
	for (int e = 0; e < kMaxSinExc; e ++)
	   if (idx == L.sin_exc_idx [e]) return L.sin_exc_val [e];
const int32_t Q = L.quarter;
	if (idx <= Q) return L.q [idx];
	if (idx <= 2 * Q) return L.q [2 * Q - idx];
	if (idx <= 3 * Q) return -L.q [idx - 2 * Q];
	return -L.q [L.rate - idx];
}

__device__ __forceinline__ float lut_cos_idx (const SinLut &L, int32_t idx) {
#pragma unroll
	for (int e = 0; e < kMaxSinExc; e ++)
	   if (idx == L.cos_exc_idx [e]) return L.cos_exc_val [e];
int32_t s = idx + L.quarter;
	if (s >= L.rate) s -= L.rate;
const int32_t Q = L.quarter;
	if (s <= Q) return L.q [s];
	if (s <= 2 * Q) return L.q [2 * Q - s];
	if (s <= 3 * Q) return -L.q [s - 2 * Q];
	return -L.q [L.rate - s];
}

// SinCos::fromPhasetoIndex for Phase >= 0 (sincos.cpp:54-56): int32 (Phase * C) % Rate
__device__ __forceinline__ int32_t phase_index (const SinLut &L, float phase) {
int32_t i = (int32_t)((double)phase * L.C);
	if (i >= L.rate) i %= L.rate;
	return i;
}

// SinCos::getSin, sincos.cpp:75-79
__device__ __forceinline__ float lut_getSin (const SinLut &L, float phase) {
	if (phase < 0.f) return -lut_sin_idx (L, phase_index (L, -phase));
	return lut_sin_idx (L, phase_index (L, phase));
}

// the phase normalisation shared by SinCos::getCos / getComplex, sincos.cpp:81-91
__device__ __forceinline__ int32_t cos_phase_index (const SinLut &L, float phase) {
	while (phase < 0.f) phase = (float)((double)phase + 2 * M_PI);
	phase = (float)fmod ((double)phase, 2 * M_PI);
	return phase_index (L, phase);
}

// PI_Constrain, includes/fm-constants.h:148-158 (all comparisons and fmod in double)
__device__ __forceinline__ float pi_constrain (float val) {
const double v = (double)val;
	if (0 <= v && v < 2 * M_PI) return val;
	if (v >= 2 * M_PI) return (float)fmod (v, 2 * M_PI);
	if (v > -2 * M_PI) return (float)(v + 2 * M_PI);
	return (float)(2 * M_PI - fmod (-v, 2 * M_PI));
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

// res_raw, zabs, iqn : K2 outputs.  demod / pilot_phase / locked : fm-rate outputs.
template <bool SMEM_LUT>
__global__ void __launch_bounds__ (kSeqLanes)
sequential_kernel (const float *__restrict__ res_raw, const float *__restrict__ zabs,
                   const float2 *__restrict__ iqn, int64_t pitch, int32_t M,
                   SeqParams P, SinLut L, const float *__restrict__ atanPPY,
                   StreamState *__restrict__ state,
                   float *__restrict__ demod_out, float *__restrict__ phase_out,
                   uint8_t *__restrict__ locked_out) {
extern __shared__ float sq [];
	if (SMEM_LUT) {
	   for (int i = threadIdx.x; i <= L.quarter; i += blockDim.x) sq [i] = L.q [i];
	   __syncthreads ();
	   L.q = sq;
	}
const int stream = blockIdx.x * blockDim.x + threadIdx.x;
	if (stream >= P.n_streams) return;
StreamState &st = state [stream];
const float *rr = res_raw + (int64_t)stream * pitch;
const float *za = zabs + (int64_t)stream * pitch;
const float2 *nq = iqn ? iqn + (int64_t)stream * pitch : nullptr;
float *dm = demod_out + (int64_t)stream * pitch;
float *ph = phase_out + (int64_t)stream * pitch;
uint8_t *lk = locked_out + (int64_t)stream * pitch;

float fm_afc = st.fm_afc, am = st.am_carr_ampl;
float phase = st.pilot_phase, oldv = st.pilot_old, plock = st.pilot_lock;
int   locked = st.pilot_locked, stable = st.pilot_stable_cnt;
float nco = st.pll_nco_phase, incr = st.pll_phase_incr;

const float carrierAlpha = 0.0010f, fmDcAlpha = 0.0001f;          // fm-demodulator.cpp:115-117
const float oneMinusCarrier = fsub (1.0f, carrierAlpha);
const float oneMinusDc = fsub (1.0f, fmDcAlpha);
const float lockAlpha = 1.0f / 3000.0f;                           // pilot-recover.cpp:57
const double oneMinusLock = 1.0 - (double)lockAlpha;

	for (int32_t m = 0; m < M; m ++) {
	   float res = rr [m];
	   const float zAbs = za [m];
	   am = fadd (fmul (oneMinusCarrier, am), fmul (carrierAlpha, zAbs));
	   if (P.decoder == 2 || P.decoder == 1) {
//	pllC::do_pll on the normalised sample, pllC.cpp:67-90
	      const float2 s = nq [m];
	      const int32_t ci = cos_phase_index (L, nco);
	      const float2 osc = make_float2 (lut_cos_idx (L, ci), lut_sin_idx (L, ci));
	      const float2 d = cmul_rn (make_float2 (osc.x, -osc.y), s);
	      const float perr = lut_atan2 (atanPPY, d.y, d.x);
	      incr = fadd (fmul (fsub (1.0f, P.pll_beta), perr), fmul (P.pll_beta, incr));
	      if (incr < P.pll_lo || incr > P.pll_hi) incr = P.pll_reset;
	      nco = fadd (nco, incr);
	      if ((double)nco >= 2 * M_PI) nco = (float)fmod ((double)nco, 2 * M_PI);
	      else while (nco < 0.f) nco = (float)((double)nco + 2 * M_PI);
	      res = incr;
	   }
	   fm_afc = fadd (fmul (oneMinusDc, fm_afc), fmul (fmDcAlpha, res));
	   const float demod = fdiv (fmul (fmul (20.0f, fsub (res, fm_afc)), 1.0f), P.K_FM);

//	pilotRecovery::getPilotPhase (5 * demod), pilot-recover.cpp:54-83
	   const float pilot = fmul (5.0f, demod);
	   const float osc = lut_getSin (L, phase);
	   const float perr = fmul (pilot, osc);
	   phase = fadd (phase, fmul (perr, P.gain));
	   const float cur = pi_constrain (phase);
	   phase = pi_constrain (fadd (phase, P.omega));
	   const float quad = fdiv (fsub (osc, oldv), P.omega);
	   oldv = osc;
	   plock = (float)((double)fmul (lockAlpha, fmul (-quad, pilot)) +
	                   (double)plock * oneMinusLock);
	   if (plock > 0.07f) {
	      if (locked || ++stable > P.lock_half_rate) locked = 1;
	   }
	   else { locked = 0; stable = 0; }

	   dm [m] = demod;
	   ph [m] = cur;
	   lk [m] = (uint8_t)locked;
	}
	st.fm_afc = fm_afc; st.am_carr_ampl = am;
	st.pilot_phase = phase; st.pilot_old = oldv; st.pilot_lock = plock;
	st.pilot_locked = locked; st.pilot_stable_cnt = stable;
	st.pll_nco_phase = nco; st.pll_phase_incr = incr;
}

// K6a — de-emphasis one-pole and gain, fm-processor.cpp:594-595 and :303-306, one lane per
// stream (two independent float chains, left and right), bit-exact.
struct DeemphParams { float alpha, gl, gr; int32_t n_streams; };

__global__ void deemphasis_kernel (const float2 *__restrict__ lr, int64_t pitch, int32_t M,
                                   DeemphParams P, StreamState *__restrict__ state,
                                   float2 *__restrict__ out) {
const int stream = blockIdx.x * blockDim.x + threadIdx.x;
	if (stream >= P.n_streams) return;
StreamState &st = state [stream];
const float2 *in = lr + (int64_t)stream * pitch;
float2 *o = out + (int64_t)stream * pitch;
float l = st.deemph_l, r = st.deemph_r;
int32_t m = 0;
	for (; m + 4 <= M; m += 4) {
	   float2 v [4];
#pragma unroll
	   for (int k = 0; k < 4; k ++) v [k] = in [m + k];
#pragma unroll
	   for (int k = 0; k < 4; k ++) {
	      l = fadd (fmul (fsub (v [k].x, l), P.alpha), l);
	      r = fadd (fmul (fsub (v [k].y, r), P.alpha), r);
	      o [m + k] = make_float2 (fmul (P.gl, l), fmul (P.gr, r));
	   }
	}
	for (; m < M; m ++) {
	   const float2 v = in [m];
	   l = fadd (fmul (fsub (v.x, l), P.alpha), l);
	   r = fadd (fmul (fsub (v.y, r), P.alpha), r);
	   o [m] = make_float2 (fmul (P.gl, l), fmul (P.gr, r));
	}
	st.deemph_l = l; st.deemph_r = r;
}

}	// namespace sdrjfm
