// K6 — audio back end at the fm rate, one launch, parallel over tiles AND streams:
//   de-emphasis one-pole      src/fm/fm-processor.cpp:594-595   (alpha from setDeemphasis :291-297)
//   audioGainCorrection       src/fm/fm-processor.cpp:303-306
//   192 kHz -> working rate   src/fm/fm-processor.cpp:633-634   (newConverter = libsamplerate)
//   fade-in after (re)tune    src/fm/fm-processor.cpp:638-642
//
// De-emphasis is the linear recurrence y[n] = y[n-1] + alpha (x[n] - y[n-1]).  With
// alpha = 0.094 (50 us) a state error decays by (1 - alpha)^384 < 1e-16, so every tile
// simply starts kAuWarm samples early from a zero state and is exact to float rounding;
// the tile that contains the first sample of the call starts from the carried state instead.
// Inside a tile the recurrence is a block scan of affine maps.
//
// The reference does the 192 -> 48 kHz step with libsamplerate (SRC_SINC_MEDIUM_QUALITY,
// src/various/newconverter.cpp:26-80).  That library is neither vendored in the reference
// tree nor installed here, so its coefficient table cannot be restated: PARITY UNPINNED for
// this stage.  What is built instead is a documented polyphase windowed-sinc
// decimator of our own,
//     y[q] = sum_{i<129} h[i] a[4 q + 3 - i],   h = Blackman-windowed sinc, fc = 20 kHz
// validated against a float64 model of the same taps.
#pragma once
#include <vector>
#include <cmath>
#include "common.cuh"
#include "audio_post.cuh"

namespace sdrjfm {

constexpr int kRsTaps  = 129;
constexpr int kRsHist  = 128;
constexpr int kRsDecim = 4;

constexpr int kAuThreads = 256;
constexpr int kAuRun     = 10;                          // consecutive fm samples per thread
constexpr int kAuSpan    = kAuThreads * kAuRun;         // 2560 samples staged per CTA
constexpr int kAuTile    = 2048;                        // fm samples owned by a CTA
constexpr int kAuLead    = kAuSpan - kAuTile;           // 512 = 384 warm-up + 128 filter history
constexpr int kAuWarm    = kAuLead - kRsHist;

__constant__ float c_rs_taps [kRsTaps + 3];

struct AudioParams {
	float   alpha, gl, gr;
	int32_t M;                 // fm samples of this call
	int64_t g0;                // global fm index of local sample 0 (decimation phase)
	int64_t q0;                // global output index of the first output of this call
	int32_t nq;                // outputs of this call
	int32_t fade_cnt, fade_max;
	int32_t write_tap;         // keep the 192 kHz stream (SDRJFM_TAP_AUDIO192)
	int32_t sel;               // which de-emphasis state buffer holds the carried state
	int32_t plot;              // 0, or ELfPlot AF_MONO / AF_LEFT / AF_RIGHT _FILTERED (5 / 6 / 7): scope stream wanted
	ToneParams tone;           // insertTestTone (fm-processor.cpp:800-823), behind the fade-in
};

// lr    : [S][pitch] fm-rate (left, right) of this call
// hist  : [S][128] de-emphasised, gain-corrected samples preceding this call; new_hist: same, rolled
// a192  : [S][pitch] tap (optional);  out: [S][out_pitch] working-rate stereo
// plot  : [S][pitch] the de-emphasised audio BEFORE the gain, as the LF scope sees it (fm-processor.cpp:614-622)
__global__ void __launch_bounds__ (kAuThreads)
audio_kernel (const float2 *__restrict__ lr, int64_t pitch, AudioParams P,
              const float2 *__restrict__ hist, float2 *__restrict__ new_hist,
              StreamState *__restrict__ state, float2 *__restrict__ a192,
              float2 *__restrict__ out, int64_t out_pitch, float *__restrict__ plot,
              const float *__restrict__ tone_tab) {
// phase-major staging: sample with global index g sits at [(g - G0) & 3][(g - G0) >> 2]
__shared__ float2 sA [4][kAuSpan / 4 + 2];
__shared__ float  sWl [kAuThreads / 32], sWr [kAuThreads / 32], sWa [kAuThreads / 32];
const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
const int stream = blockIdx.y;
const int t0 = blockIdx.x * kAuTile;                    // first owned local sample
const int s0 = t0 - kAuLead;                            // first staged local sample
const float2 *in = lr + (int64_t)stream * pitch;
const int64_t G0 = (P.g0 + s0) & ~(int64_t)3;           // staging origin, multiple of 4 (may be < 0)
const int shift = (int)((P.g0 + s0) - G0);

//	-- de-emphasis: thread-local zero-state response, block scan, second pass ---------------
const int n0 = s0 + tid * kAuRun;
float2 x [kAuRun];
const float oma = fsub (1.0f, P.alpha);
float A = 1.0f, Bl = 0.f, Br = 0.f;
#pragma unroll
	for (int j = 0; j < kAuRun; j ++) {
	   const int n = n0 + j;
	   const bool ok = n >= 0 && n < P.M;
	   x [j] = ok ? in [n] : make_float2 (0.f, 0.f);
	   if (ok) {
	      Bl = fadd (fmul (fsub (x [j].x, Bl), P.alpha), Bl);
	      Br = fadd (fmul (fsub (x [j].y, Br), P.alpha), Br);
	      A *= oma;
	   }
	}
float iA = A, iBl = Bl, iBr = Br;                       // inclusive scan of (A, B) over lanes
#pragma unroll
	for (int k = 1; k < 32; k <<= 1) {
	   const float yA = __shfl_up_sync (0xffffffffu, iA, k);
	   const float yl = __shfl_up_sync (0xffffffffu, iBl, k);
	   const float yr = __shfl_up_sync (0xffffffffu, iBr, k);
	   if (lane >= k) { iBl = fmaf (iA, yl, iBl); iBr = fmaf (iA, yr, iBr); iA *= yA; }
	}
	if (lane == 31) { sWa [warp] = iA; sWl [warp] = iBl; sWr [warp] = iBr; }
	__syncthreads ();
//	state entering the staged span: the carried state if the span starts at (or before) the
//	first sample of the call, else zero (the warm-up forgets it)
float yl = 0.f, yr = 0.f;
	if (s0 <= 0) { yl = state [stream].deemph [P.sel][0]; yr = state [stream].deemph [P.sel][1]; }
	for (int q = 0; q < warp; q ++) { yl = fmaf (sWa [q], yl, sWl [q]); yr = fmaf (sWa [q], yr, sWr [q]); }
	{
	   float eA = __shfl_up_sync (0xffffffffu, iA, 1);
	   float el = __shfl_up_sync (0xffffffffu, iBl, 1);
	   float er = __shfl_up_sync (0xffffffffu, iBr, 1);
	   if (lane == 0) { eA = 1.f; el = 0.f; er = 0.f; }
	   yl = fmaf (eA, yl, el); yr = fmaf (eA, yr, er);
	}
float2 *tap = a192 ? a192 + (int64_t)stream * pitch : nullptr;
#pragma unroll
	for (int j = 0; j < kAuRun; j ++) {
	   const int n = n0 + j;
	   float2 v;
	   if (n >= 0 && n < P.M) {
	      yl = fadd (fmul (fsub (x [j].x, yl), P.alpha), yl);       // :594-595
	      yr = fadd (fmul (fsub (x [j].y, yr), P.alpha), yr);
	      v = make_float2 (fmul (P.gl, yl), fmul (P.gr, yr));       // :304-305
	      if (tap && n >= t0 && P.write_tap) tap [n] = v;
	      if (P.plot && n >= t0)
	         plot [(int64_t)stream * pitch + n] = P.plot == 5 ? fadd (yl, yr) : P.plot == 6 ? yl : yr;
	      if (n == P.M - 1) { state [stream].deemph [P.sel ^ 1][0] = yl; state [stream].deemph [P.sel ^ 1][1] = yr; }
	   }
	   else if (n < 0 && n >= -kRsHist) v = hist [(int64_t)stream * kRsHist + kRsHist + n];
	   else v = make_float2 (0.f, 0.f);
	   const int r = n - s0 + shift;
	   sA [r & 3][r >> 2] = v;
	}
	__syncthreads ();

//	-- history for the next call: the last 128 samples of (old history | this call) ----------
	if (t0 + kAuTile >= P.M && t0 < P.M && tid < kRsHist) {        // the CTA owning the last sample
	   const int n = P.M - kRsHist + tid;
	   float2 v;
	   if (n >= s0 && n >= -kRsHist) { const int r = n - s0 + shift; v = sA [r & 3][r >> 2]; }
	   else v = hist [(int64_t)stream * kRsHist + kRsHist + n];    // n < -128 + ... only when M < 128
	   if (n < -kRsHist) v = make_float2 (0.f, 0.f);
	   new_hist [(int64_t)stream * kRsHist + tid] = v;
	}

//	-- 129-tap polyphase decimator: outputs at global fm index g = 3 mod 4 -------------------
	for (int o = tid; o < kAuTile / kRsDecim + 1; o += kAuThreads) {
	   // o-th candidate output of the tile: the first g >= g(t0) with g = 3 mod 4
	   const int64_t gt0 = P.g0 + t0;
	   const int64_t g = ((gt0 + 0) | 3) + 4 * (int64_t)o;          // (gt0 | 3) is the first such g
	   const int n = (int)(g - P.g0);                                // local index of the newest sample
	   if (n >= t0 + kAuTile || n >= P.M) break;
	   const int r = n - s0 + shift;                                 // r & 3 == 3 by construction
	   const int k = r >> 2;
	   float2 acc = make_float2 (0.f, 0.f);
#pragma unroll
	   for (int ph = 0; ph < 4; ph ++) {                             // taps i = 4 j + ph hit phase 3 - ph
	      const float2 *col = sA [3 - ph];
#pragma unroll 8
	      for (int j = 0; 4 * j + ph < kRsTaps; j ++) {
	         const float2 v = col [k - j];
	         const float c = c_rs_taps [4 * j + ph];
	         acc = ffma2 (c, v, acc);
	      }
	   }
	   const int64_t q = g >> 2;                                     // global output index
	   const int32_t cnt = P.fade_cnt - (int32_t)(q - P.q0);         // :638-642
	   if (cnt > 0) {
	      const float f = fdiv (fsub ((float)P.fade_max, (float)cnt), (float)P.fade_max);
	      acc.x = fmul (acc.x, f); acc.y = fmul (acc.y, f);
	   }
	   if (P.tone.on) acc = tone_apply (P.tone, tone_tab, acc, q - P.q0);   // :644
	   out [(int64_t)stream * out_pitch + (q - P.q0)] = acc;
	}
}

// ---- optional audio low-pass -------------------------------------------------------------------
// fmAudioFilter = fftFilter (8192, 756).setLowPass (lowPassFrequency, fmRate), applied to the
// (left, right) pair after the selector and before de-emphasis (fm-processor.cpp:76,403-408,
// 589-591).  The taps are real, so it is the same 756-tap FIR on each channel, delayed by
// NumofSamples = 8192 - 756 = 7436 samples:  y[n] = sum_j h[j] q[n - 7436 - j].
constexpr int kAlpTaps    = 756;
constexpr int kAlpDelay   = 8192 - kAlpTaps;          // 7436
constexpr int kAlpHist    = 8192;                     // >= 7436 + 755 samples of (left,right) history
constexpr int kAlpThreads = 256;
constexpr int kAlpPer     = 4;
constexpr int kAlpTile    = kAlpThreads * kAlpPer;    // 1024 outputs per CTA
constexpr int kAlpSpan    = kAlpTile + kAlpTaps;      // staged inputs (1780 >= 1024 + 755)

__constant__ float c_alp_taps [kAlpTaps];

// q: [S][pitch] this call's (left,right); hist: [S][8192] the samples before it (oldest first)
__global__ void __launch_bounds__ (kAlpThreads)
audio_lp_kernel (const float2 *__restrict__ q, int64_t pitch, int32_t M,
                 const float2 *__restrict__ hist, float2 *__restrict__ out) {
// phase-major staging: staged element e sits at [e & 3][e >> 2] (conflict-free sliding windows)
__shared__ float2 sQ [4][kAlpSpan / 4 + 2];
const int tid = threadIdx.x;
const int stream = blockIdx.y;
const int t0 = blockIdx.x * kAlpTile;
const int lo = t0 - kAlpDelay - (kAlpTaps - 1);       // local index of staged element 0
const float2 *qs = q + (int64_t)stream * pitch;
const float2 *hs = hist + (int64_t)stream * kAlpHist;
	for (int e = tid; e < kAlpSpan; e += kAlpThreads) {
	   const int n = lo + e;
	   float2 v = make_float2 (0.f, 0.f);
	   if (n >= 0) { if (n < M) v = qs [n]; }
	   else if (n >= -kAlpHist) v = hs [kAlpHist + n];
	   sQ [e & 3][e >> 2] = v;
	}
	__syncthreads ();
//	output t0 + 4 tid + k, tap j reads staged element 4 tid + k + 755 - j
float2 acc [kAlpPer];
#pragma unroll
	for (int k = 0; k < kAlpPer; k ++) acc [k] = make_float2 (0.f, 0.f);
float2 w [kAlpPer];
#pragma unroll
	for (int k = 0; k < kAlpPer; k ++) { const int e = 4 * tid + k + kAlpTaps - 1; w [k] = sQ [e & 3][e >> 2]; }
#pragma unroll 4
	for (int j = 0; j < kAlpTaps; j ++) {
	   const float c = c_alp_taps [j];
#pragma unroll
	   for (int k = 0; k < kAlpPer; k ++) {
	      acc [k] = ffma2 (c, w [k], acc [k]);
	   }
#pragma unroll
	   for (int k = kAlpPer - 1; k > 0; k --) w [k] = w [k - 1];
	   const int e = 4 * tid + kAlpTaps - 2 - j;          // element for k = 0 at tap j + 1
	   w [0] = e >= 0 ? sQ [e & 3][e >> 2] : make_float2 (0.f, 0.f);
	}
#pragma unroll
	for (int k = 0; k < kAlpPer; k ++) {
	   const int n = t0 + 4 * tid + k;
	   if (n < M) out [(int64_t)stream * pitch + n] = acc [k];
	}
}

// mono / unlocked path of process_signal_with_rds + the L/R matrix, fm-processor.cpp:728-730
// and :517-549 with diffLR = 0: left = right = sumLR = demod for every selector except
// S_LEFTminusRIGHT(_Test), which yields 0.
__global__ void mono_matrix_kernel (const float *__restrict__ demod, int64_t pitch, int32_t M,
                                    int32_t sound_sel, float2 *__restrict__ lr) {
const int stream = blockIdx.y;
const int m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= M) return;
const int64_t o = (int64_t)stream * pitch + m;
const float d = demod [o];
	lr [o] = (sound_sel == 5 || sound_sel == 6) ? make_float2 (0.f, 0.f) : make_float2 (d, d);
}

}	// namespace sdrjfm
