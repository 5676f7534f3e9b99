// Shared device helpers and the per-stream state layout of the B200 FM path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sdrjfm {

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

constexpr int kDecim      = 12;    // input samples per fm-rate sample (2304000 / 192000)
constexpr int kHist       = 36;    // raw IQ history the 37-tap composite needs (3 groups)
constexpr int kMaxSinExc  = 4;

// The reference build has no FMA contraction (x86-64 baseline), so
// wherever a reference recurrence is restated sample for sample the operations are spelled
// with the round-to-nearest intrinsics, which nvcc never fuses.
__device__ __forceinline__ float fmul (float a, float b) { return __fmul_rn (a, b); }
__device__ __forceinline__ float fadd (float a, float b) { return __fadd_rn (a, b); }
__device__ __forceinline__ float fsub (float a, float b) { return __fsub_rn (a, b); }
__device__ __forceinline__ float fdiv (float a, float b) { return __fdiv_rn (a, b); }

// Blackwell packed FP32: one FFMA2 instruction does acc.{x,y} = c * w.{x,y} + acc.{x,y} (two fma.rn,
// the scalar tap is broadcast).  Real taps times complex samples is the shape of every FIR here.
__device__ __forceinline__ float2 ffma2 (float c, float2 w, float2 acc) {
unsigned long long a, b, d, r;
	asm ("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(c), "f"(c));
	asm ("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(w.x), "f"(w.y));
	asm ("mov.b64 %0, {%1,%2};" : "=l"(d) : "f"(acc.x), "f"(acc.y));
	asm ("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(d));
float2 o;
	asm ("mov.b64 {%0,%1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
	return o;
}

// the same with the tap already duplicated into a register pair (taps stored as float2 {c, c} in
// constant memory: one 64-bit constant load feeds the instruction directly, no packing moves)
__device__ __forceinline__ float2 ffma2p (float2 cc, float2 w, float2 acc) {
unsigned long long a, b, d, r;
	asm ("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(cc.x), "f"(cc.y));
	asm ("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(w.x), "f"(w.y));
	asm ("mov.b64 %0, {%1,%2};" : "=l"(d) : "f"(acc.x), "f"(acc.y));
	asm ("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(d));
float2 o;
	asm ("mov.b64 {%0,%1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
	return o;
}

// std::complex<float> operator* as GCC emits it without -ffast-math: four products, each
// rounded, then one subtraction and one addition.
__device__ __forceinline__ float2 cmul_rn (float2 a, float2 b) {
	return make_float2 (fsub (fmul (a.x, b.x), fmul (a.y, b.y)),
	                    fadd (fmul (a.x, b.y), fmul (a.y, b.x)));
}

// generic inclusive scan of affine maps x -> A x + B over the CTA (thread order); returns
// the value entering this thread given `carry` entering thread 0, and the value leaving the
// last thread through *total.
__device__ __forceinline__ double block_affine_start (double A, double B, double carry,
                                                       double *sA, double *sB, double *total) {
const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
double iA = A, iB = B;
#pragma unroll
	for (int k = 1; k < 32; k <<= 1) {
	   const double yA = __shfl_up_sync (0xffffffffu, iA, k);
	   const double yB = __shfl_up_sync (0xffffffffu, iB, k);
	   if (lane >= k) { iB = iA * yB + iB; iA = iA * yA; }
	}
	if (lane == 31) { sA [warp] = iA; sB [warp] = iB; }
	__syncthreads ();
double w = carry;
	for (int q = 0; q < warp; q ++) w = sA [q] * w + sB [q];
double eA = __shfl_up_sync (0xffffffffu, iA, 1);
double eB = __shfl_up_sync (0xffffffffu, iB, 1);
	if (lane == 0) { eA = 1.0; eB = 0.0; }
const double start = eA * w + eB;
	if (total) {
	   double t = carry;
	   for (int q = 0; q < (int)(blockDim.x >> 5); q ++) t = sA [q] * t + sB [q];
	   *total = t;
	}
	__syncthreads ();
	return start;
}

// State carried from one process call to the next, one record per IQ stream.
// (The reference keeps the same quantities in fmProcessor / fm_Demodulator / pilotRecovery
// / PerfectStereoSeparation member variables.)
struct StreamState {
	// RF DC remover, fm-processor.cpp:425 — tracked in double at the fm rate
	double  dc_re, dc_im;
	float   sprev [4];               // 12-sample block sums S[m-1], S[m-2] (re, im each)
	// fm_Demodulator members, fm-demodulator.cpp:79-86
	float   Imin1, Qmin1, Imin2, Qmin2;
	float   fm_afc, am_carr_ampl;
	// pllC (PLL decoder), pllC.cpp:43-58
	float   pll_nco_phase, pll_phase_incr;
	// pilotRecovery members, pilot-recover.cpp:28-43
	float   pilot_phase, pilot_old, pilot_lock;
	int32_t pilot_locked, pilot_stable_cnt;
	// fmProcessor::pilotDelayPSS and PerfectStereoSeparation members
	float   pss_delay, pss_acc, pss_mean_error;
	int32_t pss_minimized, pss_lock_cnt, pss_unlock_cnt;
	int32_t pss_inp;                 // fftFilter::inp of the PSS low-pass
	// de-emphasis, fm-processor.cpp:594-595
	float   deemph [2][2];           // [buffer][left,right]: read from one, written to the other (K6)
	// RDS: fftFilter::inp of band-pass and Hilbert (equal), rdsPhaseIndex, rdsDecimator counter
	int32_t rds_inp, rds_phase_idx, rds_decim_cnt;
	// audio: samples still to fade in (suppressAudioSampleCnt), peak meter
	int32_t fade_cnt;
	// evaluatePeakLevel (fm-processor.cpp:772-798): running |left|, |right| maxima of the open 961-sample block,
	// double-buffered (the CTA closing a call's last block writes what the next call's first block reads)
	float   peak_carry [2][2];
	int32_t pad [2];
};

// Settings snapshot taken at a process boundary (fm-processor.cpp:397-413 does the same
// at every 16384-sample block).
struct Settings {
	int32_t fm_mode, decoder, sound_sel, rds_mode;
	int32_t auto_mono, pss_on, dc_remove, lo_hz;
	int32_t input_filter_hz, lf_cutoff_hz, squelch_mode, deemph_us;
	int32_t squelch_value, pad1 [3];
	float   lgain, rgain, volume, panorama;
	float   left_ch, right_ch, deemph_alpha, pad0;
};

}	// namespace sdrjfm
