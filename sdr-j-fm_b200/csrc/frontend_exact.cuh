// K1x — the front end in the REFERENCE'S OWN OPERATION ORDER (front_end_mode 2, and chosen
// automatically for the table-indexed decoders and for lo != 0).
//
// The composite front end (frontend_tma.cuh / frontend_fir.cuh) re-associates the two decimators
// into one real 37-tap FIR and finishes the RF DC removal at the fm rate: fm-rate samples within
// 3e-7 of the reference's.  That is far inside the audio tolerance, but the PLL and real-baseband
// decoders read look-up tables whose index flips on such a difference.  These kernels instead
// restate, operation by operation,
//     RfDC = (x - RfDC) * rfDcAlpha + RfDC; x -= clamp (RfDC, +-0.01)      fm-processor.cpp:423-446
//     v = (re * Lgain, im * Rgain) * Oscillator::nextValue (lo)            fm-processor.cpp:462-466
//     fmBand_1.Pass: tmp += Buffer [newest - i] * filterKernel [i], i = 0 .. 24, every 6th sample
//     fmBand_2.Pass: the same with D2 + 1 taps on every D2-th stage-1 output   fir-filters.cpp:397-424
// with std::complex<float> products in the four-multiply form GCC emits without -ffast-math and no
// FMA contraction, so the fm-rate samples are BIT-IDENTICAL to the
// reference's (tests/test_gpu_parity.py::test_exact_front_end_is_bit_identical).
//
//   fx_dc_kernel      the RF DC one-pole is a float32 recurrence at the INPUT rate: every rounding
//                     feeds the next step, so it is walked sample by sample, one lane per stream
//                     (both components), loads prefetched a batch ahead.  ~14 cycles per sample:
//                     this stage is latency-bound (7 ms per 0.5 s of signal for any number of
//                     streams up to 32 x SMs), which is why the mode is opt-in / per decoder.
//   frontend_exact_kernel   tile of 128 fm-rate outputs per CTA: the DC-free samples are staged in
//                     shared memory (gain and oscillator applied on the way), thread t evaluates
//                     the D2 stage-1 outputs behind fm sample t and the stage-2 sum in order.
#pragma once
#include "common.cuh"
#include "frontend_fir.cuh"
#include "frontend_poly.cuh"

namespace sdrjfm {

constexpr int kFxThreads = 128;                    // fm-rate outputs per CTA
constexpr int kFxHist    = 32;                     // filter-input history kept per stream (>= 25)
constexpr int kFxTaps1   = 25;                     // fmBand_1, fm-processor.cpp:68-71
constexpr int kFxBatch   = 32;                     // samples per prefetch batch of the DC walker

constexpr int kFxMaxTaps2 = 11;                    // fmBand_2 has IRate / fmRate + 1 taps: rates up to 12.6 MS/s
__constant__ float2 c_fx_taps [kFxTaps1 + kFxMaxTaps2];  // filterKernel of fmBand_1 [25], then of fmBand_2 [D2 + 1]

// ---- RF DC removal, sample by sample ---------------------------------------------------------
// x : [S][in_pitch] samples in format rf; xd : [S][out_pitch] float2 = x - clamp (RfDC) per component
__global__ void __launch_bounds__ (32)
fx_dc_kernel (const void *__restrict__ x, int64_t in_pitch, RawFmt rf, int64_t N, int32_t S,
              float alpha, StreamState *__restrict__ state, float2 *__restrict__ xd, int64_t out_pitch,
              int write_state) {
const int stream = blockIdx.x * 32 + threadIdx.x;
	if (stream >= S) return;
StreamState &st = state [stream];
const void *xs = reinterpret_cast<const char *>(x) + (int64_t)stream * in_pitch * fmt_bytes (rf.fmt);
float2 *out = xd + (int64_t)stream * out_pitch;
float rr = (float)st.dc_re, ri = (float)st.dc_im;    // RfDC: float32 in the reference; carried exactly in the double fields
const float lim = 0.01f;                              // DCRlimit, :429
float2 cur [kFxBatch], nxt [kFxBatch];
	if (rf.fmt == kFmtCF32) {
#pragma unroll
	   for (int j = 0; j < kFxBatch; j ++) cur [j] = j < N ? __ldcs (reinterpret_cast<const float2 *>(xs) + j) : make_float2 (0.f, 0.f);
	}
	else {
#pragma unroll
	   for (int j = 0; j < kFxBatch; j ++) cur [j] = j < N ? load_iq_rt (xs, j, rf) : make_float2 (0.f, 0.f);
	}
	for (int64_t n0 = 0; n0 < N; n0 += kFxBatch) {
	   const int64_t n1 = n0 + kFxBatch;
	   if (rf.fmt == kFmtCF32) {
#pragma unroll
	      for (int j = 0; j < kFxBatch; j ++)
	         nxt [j] = n1 + j < N ? __ldcs (reinterpret_cast<const float2 *>(xs) + n1 + j) : make_float2 (0.f, 0.f);
	   }
	   else {
#pragma unroll
	      for (int j = 0; j < kFxBatch; j ++) nxt [j] = n1 + j < N ? load_iq_rt (xs, n1 + j, rf) : make_float2 (0.f, 0.f);
	   }
#pragma unroll
	   for (int j = 0; j < kFxBatch; j ++) {
	      if (n0 + j < N) {
	         const float2 v = cur [j];
	         rr = fadd (fmul (fsub (v.x, rr), alpha), rr);                 // :425
	         ri = fadd (fmul (fsub (v.y, ri), alpha), ri);
	         const float cr = rr > lim ? lim : (rr < -lim ? -lim : rr);     // :430-443
	         const float ci = ri > lim ? lim : (ri < -lim ? -lim : ri);
	         __stcs (out + n0 + j, make_float2 (fsub (v.x, cr), fsub (v.y, ci)));   // :445
	      }
	   }
#pragma unroll
	   for (int j = 0; j < kFxBatch; j ++) cur [j] = nxt [j];
	}
	if (write_state) { st.dc_re = (double)rr; st.dc_im = (double)ri; }
}

// ---- RF DC removal, EXACT and parallel in time -------------------------------------------------
// The recurrence r[n] = fl (fl (fl (x[n] - r[n-1]) alpha) + r[n-1]) is sequential, but it can be SOLVED
// in parallel like the pilot PLL (pilot.cuh): cut a window of 3072 samples into 256 segments of 12, give
// every segment a start value s[t], let 256 threads walk their segments with the exact float32 arithmetic,
// and compare: where e[t-1] (the end of segment t-1) equals s[t] bit for bit for every t, the window IS the
// sequential trajectory (s[0] is the exact carried state, every link an exact reference step).  Otherwise
// the starts are corrected by the prefix sum of the mismatches — on a float grid the segment map is a pure
// translation except where a rounding or a binade changes, so one correction leaves at most a few 1-ulp
// mismatches — and the walk repeats.  The exact prefix grows by at least one segment per pass, so 257 passes
// always suffice; measured: 3-4 passes per window in steady state, a few tens in the first windows of a
// stream where the estimate climbs through the binades (prototype against the sequential loop: bit-identical
// on six signal classes).  One CTA per stream, both components per thread.
constexpr int kFdThreads = 256;
constexpr int kFdSeg     = 12;
constexpr int kFdWin     = kFdThreads * kFdSeg;     // 3072 input samples per window
constexpr int kFdMaxPass = kFdThreads + 1;

template <bool CF32>
__global__ void __launch_bounds__ (kFdThreads, 2)
fx_dc_par_kernel (const void *__restrict__ x, int64_t in_pitch, RawFmt rf, int64_t N, float alpha,
                  StreamState *__restrict__ state, float2 *__restrict__ xd, int64_t out_pitch, int write_state,
                  int32_t *__restrict__ pass_stats) {
__shared__ float2 sx [kFdWin + kFdWin / kFdSeg];     // slot (n) = n + n / 12: thread t's samples at 13 t + j
__shared__ float2 sE [kFdThreads];
__shared__ double sWx [kFdThreads / 32], sWy [kFdThreads / 32];
const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
const int stream = blockIdx.x;
StreamState &st = state [stream];
const void *xs = reinterpret_cast<const char *>(x) + (int64_t)stream * in_pitch * fmt_bytes (rf.fmt);
float2 *out = xd + (int64_t)stream * out_pitch;
float2 carry = make_float2 ((float)st.dc_re, (float)st.dc_im);      // RfDC: exact float32 values in the double fields
const float lim = 0.01f;                                             // DCRlimit, fm-processor.cpp:429
int passes = 0, worst = 0;

float2 nxt [kFdSeg];
#pragma unroll
	for (int k = 0; k < kFdSeg; k ++) {
	   const int64_t n = (int64_t)k * kFdThreads + tid;
	   nxt [k] = n < N ? (CF32 ? __ldcs (reinterpret_cast<const float2 *>(xs) + n) : load_iq_rt (xs, n, rf)) : make_float2 (0.f, 0.f);
	}
	for (int64_t w0 = 0; w0 < N; w0 += kFdWin) {
	   const int valid = (int)min ((int64_t)kFdWin, N - w0);
//	   stage this window (coalesced loads were issued one window ahead), request the next one
#pragma unroll
	   for (int k = 0; k < kFdSeg; k ++) {
	      const int n = k * kFdThreads + tid;
	      sx [n + n / kFdSeg] = nxt [k];
	   }
#pragma unroll
	   for (int k = 0; k < kFdSeg; k ++) {
	      const int64_t n = w0 + kFdWin + (int64_t)k * kFdThreads + tid;
	      nxt [k] = n < N ? (CF32 ? __ldcs (reinterpret_cast<const float2 *>(xs) + n) : load_iq_rt (xs, n, rf)) : make_float2 (0.f, 0.f);
	   }
	   __syncthreads ();
	   float2 v [kFdSeg];
#pragma unroll
	   for (int j = 0; j < kFdSeg; j ++) v [j] = sx [(kFdSeg + 1) * tid + j];
	   const int cnt = max (0, min (kFdSeg, valid - kFdSeg * tid));
	   float2 s = carry;
	   int pass = 0;
	   for (; pass < kFdMaxPass; pass ++) {
	      float rr = s.x, ri = s.y;
#pragma unroll
	      for (int j = 0; j < kFdSeg; j ++) {
	         if (j < cnt) {
	            rr = fadd (fmul (fsub (v [j].x, rr), alpha), rr);            // :425
	            ri = fadd (fmul (fsub (v [j].y, ri), alpha), ri);
	         }
	      }
	      sE [tid] = make_float2 (rr, ri);
	      __syncthreads ();
	      const float2 prev = tid ? sE [tid - 1] : s;
	      double dx = (double)prev.x - (double)s.x, dy = (double)prev.y - (double)s.y;      // mismatch at this segment's start
	      const int bad = (dx != 0.0) || (dy != 0.0);
#pragma unroll
	      for (int k = 1; k < 32; k <<= 1) {                                  // inclusive prefix sums over the warp
	         const double ux = __shfl_up_sync (0xffffffffu, dx, k), uy = __shfl_up_sync (0xffffffffu, dy, k);
	         if (lane >= k) { dx += ux; dy += uy; }
	      }
	      if (lane == 31) { sWx [warp] = dx; sWy [warp] = dy; }
	      if (!__syncthreads_or (bad)) break;
	      for (int q = 0; q < warp; q ++) { dx += sWx [q]; dy += sWy [q]; }
	      s = make_float2 ((float)((double)s.x + dx), (float)((double)s.y + dy));
	   }
	   carry = sE [kFdThreads - 1];
	   passes += pass + 1; worst = max (worst, pass + 1);
//	   the starts are exact now: one more walk subtracts the clamped estimate; results go back through shared
//	   memory for coalesced stores
	   {
	      float rr = s.x, ri = s.y;
#pragma unroll
	      for (int j = 0; j < kFdSeg; j ++) {
	         if (j < cnt) {
	            rr = fadd (fmul (fsub (v [j].x, rr), alpha), rr);
	            ri = fadd (fmul (fsub (v [j].y, ri), alpha), ri);
	            const float cr = rr > lim ? lim : (rr < -lim ? -lim : rr);    // :430-443
	            const float ci = ri > lim ? lim : (ri < -lim ? -lim : ri);
	            sx [(kFdSeg + 1) * tid + j] = make_float2 (fsub (v [j].x, cr), fsub (v [j].y, ci));  // :445
	         }
	      }
	   }
	   __syncthreads ();
#pragma unroll
	   for (int k = 0; k < kFdSeg; k ++) {
	      const int n = k * kFdThreads + tid;
	      if (n < valid) __stcs (out + w0 + n, sx [n + n / kFdSeg]);
	   }
	   __syncthreads ();
	}
	if (tid == 0) {
	   if (write_state) { st.dc_re = (double)carry.x; st.dc_im = (double)carry.y; }
	   if (pass_stats) { pass_stats [2 * stream] = passes; pass_stats [2 * stream + 1] = worst; }
	}
}

// gain and oscillator of one filter-input sample (fm-processor.cpp:462-466); n = its index in the call
__device__ __forceinline__ float2 fx_stage (const LoParams &L, float2 v, int64_t n) {
	v = make_float2 (fmul (v.x, L.lgain), fmul (v.y, L.rgain));
	if (L.tab) v = cmul_rn (v, L.tab [lo_index (L, n)]);     // lo == 0 multiplies by Table [0] = (1, 0): the identity
	return v;
}

template <int D2_>
struct Fx {
	static constexpr int D2   = D2_;                       // stage-2 decimation (IRate / fmRate)
	static constexpr int D    = 6 * D2;                    // input samples per fm-rate sample
	static constexpr int NT2  = D2 + 1;                    // stage-2 taps, fm-processor.cpp:72-75
	static constexpr int Span = D * kFxThreads + kFxTaps1; // staged inputs: D m0 - 25 .. D (m0 + 128) - 1
	static constexpr int Slots = Span + Span / D + 2;      // one pad slot per D samples: odd thread stride
	static constexpr int SmemBytes = Slots * (int)sizeof (float2);
};

// src   : filter-input samples of this call BEFORE gain / oscillator: xd (float2, fmt cf32) when the
//         DC remover is on, else the raw samples in their device format
// xhist : [S][kFxHist] the last filter inputs AFTER gain / oscillator of the previous call
// Z     : [S][out_pitch] fm-rate samples (the reference's v after fmBand_2, :474)
template <int D2, bool PLAIN>       // PLAIN: complex-float source (the DC solver's output), oscillator off
__global__ void __launch_bounds__ (kFxThreads)
frontend_exact_kernel (const void *__restrict__ src, int64_t in_pitch, RawFmt rf,
                       const float2 *__restrict__ xhist, const LoParams lop,
                       float2 *__restrict__ Z, int64_t out_pitch, int32_t M) {
typedef Fx<D2> P;
constexpr int D = P::D;
extern __shared__ float2 sm [];
__shared__ float2 sY [kFxThreads + 1];
const int tid = threadIdx.x;
const int stream = blockIdx.y;
const int64_t m0 = (int64_t)blockIdx.x * kFxThreads;
const int64_t O = D * m0 - kFxTaps1;                   // call-relative index of staged element 0
const int64_t N = (int64_t)M * D;
const void *xs = reinterpret_cast<const char *>(src) + (int64_t)stream * in_pitch * fmt_bytes (rf.fmt);
const float2 *hs = xhist + (int64_t)stream * kFxHist;
	for (int e = tid; e < P::Span; e += kFxThreads) {
	   const int64_t n = O + e;
	   float2 v = make_float2 (0.f, 0.f);
	   if (n < 0) v = hs [kFxHist + n];
	   else if (n < N) {
	      if (PLAIN) {
	         v = __ldcs (reinterpret_cast<const float2 *>(xs) + n);
	         v = make_float2 (fmul (v.x, lop.lgain), fmul (v.y, lop.rgain));
	      }
	      else v = fx_stage (lop, rf.fmt == kFmtCF32 ? __ldcs (reinterpret_cast<const float2 *>(xs) + n)
	                                                 : load_iq_rt (xs, n, rf), n);
	   }
	   sm [e + e / D] = v;
	}
	__syncthreads ();

//	stage 1: the D2 outputs k = D2 (m0 + t) + j; output k reads inputs 6 k + 5 - i = D (m0 + t) + 6 j + 5 - i,
//	i.e. staged element D t + q with q = 6 j + 30 - i
const float2 *mine = sm + (D + 1) * tid;
float2 y1 [D2];
#pragma unroll
	for (int j = 0; j < D2; j ++) {
	   float2 acc = make_float2 (0.f, 0.f);
#pragma unroll
	   for (int i = 0; i < kFxTaps1; i ++) {
	      const int q = 6 * j + 30 - i;
	      const float2 p = cmul_rn (mine [q + q / D], c_fx_taps [i]);
	      acc.x = fadd (acc.x, p.x); acc.y = fadd (acc.y, p.y);
	   }
	   y1 [j] = acc;
	}
	sY [tid + 1] = y1 [D2 - 1];
	if (tid == 0) {             // stage-1 output D2 m0 - 1 (the newest one of the fm sample before the tile)
	   float2 acc = make_float2 (0.f, 0.f);
#pragma unroll
	   for (int i = 0; i < kFxTaps1; i ++) {
	      const int q = 24 - i;
	      const float2 p = cmul_rn (sm [q + q / D], c_fx_taps [i]);
	      acc.x = fadd (acc.x, p.x); acc.y = fadd (acc.y, p.y);
	   }
	   sY [0] = acc;
	}
	__syncthreads ();
//	stage 2: tmp += y1 [D2 m + D2 - 1 - i] * K2 [i], i = 0 .. D2
float2 z = make_float2 (0.f, 0.f);
#pragma unroll
	for (int i = 0; i < P::NT2; i ++) {
	   const float2 a = i < D2 ? y1 [D2 - 1 - i] : sY [tid];
	   const float2 p = cmul_rn (a, c_fx_taps [kFxTaps1 + i]);
	   z.x = fadd (z.x, p.x); z.y = fadd (z.y, p.y);
	}
	if (m0 + tid < M) Z [(int64_t)stream * out_pitch + m0 + tid] = z;
}

// ANY other rate the reference's constructor arithmetic accepts (fm-processor.cpp:36,68-75: stage 1 = 25 taps / 6,
// stage 2 = D2 + 1 taps / D2 with D2 = (inputRate / 6) / fmRate = 1 .. 10; e.g. 2.048 MS/s -> / 6, 4 MS/s -> / 18,
// 8 MS/s -> / 36): the same arithmetic with D2 at run time, one thread per fm-rate output reading its inputs
// straight from global memory.  Not tuned — the three device rates have their own kernels — but exact.
__global__ void __launch_bounds__ (kFxThreads)
frontend_exact_generic_kernel (const void *__restrict__ src, int64_t in_pitch, RawFmt rf, int D2,
                               const float2 *__restrict__ xhist, const LoParams lop,
                               float2 *__restrict__ Z, int64_t out_pitch, int32_t M) {
const int stream = blockIdx.y;
const int64_t m = (int64_t)blockIdx.x * kFxThreads + threadIdx.x;
	if (m >= M) return;
const int D = 6 * D2;
const void *xs = reinterpret_cast<const char *>(src) + (int64_t)stream * in_pitch * fmt_bytes (rf.fmt);
const float2 *hs = xhist + (int64_t)stream * kFxHist;
float2 z = make_float2 (0.f, 0.f);
	for (int i2 = 0; i2 <= D2; i2 ++) {                 // stage 2, newest stage-1 output first
	   const int64_t k6 = 6 * (D2 * m + D2 - 1 - i2) + 5;   // newest input of stage-1 output D2 m + D2 - 1 - i2
	   float2 acc = make_float2 (0.f, 0.f);
	   for (int i = 0; i < kFxTaps1; i ++) {
	      const int64_t n = k6 - i;
	      const float2 v = n < 0 ? hs [kFxHist + n] : fx_stage (lop, load_iq_rt (xs, n, rf), n);
	      const float2 p = cmul_rn (v, c_fx_taps [i]);
	      acc.x = fadd (acc.x, p.x); acc.y = fadd (acc.y, p.y);
	   }
	   const float2 p = cmul_rn (acc, c_fx_taps [kFxTaps1 + i2]);
	   z.x = fadd (z.x, p.x); z.y = fadd (z.y, p.y);
	}
	(void)D;
	Z [(int64_t)stream * out_pitch + m] = z;
}

// history for the next call: the last kFxHist filter inputs AFTER gain / oscillator
__global__ void fx_roll_hist_kernel (const void *__restrict__ src, int64_t in_pitch, RawFmt rf, const LoParams lop,
                                     const float2 *__restrict__ old_hist, float2 *__restrict__ new_hist, int64_t n_proc) {
const int stream = blockIdx.x, i = threadIdx.x;
	if (i >= kFxHist) return;
const int64_t pos = n_proc - kFxHist + i;
const void *xs = reinterpret_cast<const char *>(src) + (int64_t)stream * in_pitch * fmt_bytes (rf.fmt);
	new_hist [(int64_t)stream * kFxHist + i] =
	      pos >= 0 ? fx_stage (lop, load_iq_rt (xs, pos, rf), pos) : old_hist [(int64_t)stream * kFxHist + kFxHist + pos];
}

// entering the mode in mid-stream (decoder or oscillator changed between calls): the filter-input history
// is rebuilt from the raw-sample history the composite front end keeps, with the DC estimate as it stands
// (the reference subtracted the estimate of each sample's own time: a difference of < 1e-8 on 25 samples)
__global__ void fx_seed_hist_kernel (const float2 *__restrict__ raw_hist, int hist_len, const LoParams lop,
                                     const StreamState *__restrict__ state, int dc_remove, float2 *__restrict__ new_hist) {
const int stream = blockIdx.x, i = threadIdx.x;
	if (i >= kFxHist) return;
float2 v = raw_hist [(int64_t)stream * hist_len + hist_len - kFxHist + i];
	if (dc_remove) {
	   const float lim = 0.01f;
	   const float rr = (float)state [stream].dc_re, ri = (float)state [stream].dc_im;
	   v.x = fsub (v.x, rr > lim ? lim : (rr < -lim ? -lim : rr));
	   v.y = fsub (v.y, ri > lim ? lim : (ri < -lim ? -lim : ri));
	}
	new_hist [(int64_t)stream * kFxHist + i] = fx_stage (lop, v, (int64_t)i - kFxHist);
}

// the filter-input history of a composite front end changes meaning when the per-sample DC remover is
// switched in or out in front of it (raw samples <-> DC-free samples): shift it by the clamped estimate
__global__ void hist_shift_dc_kernel (float2 *__restrict__ hist, int hist_len, const StreamState *__restrict__ state, float sign) {
const int stream = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hist_len) return;
const float lim = 0.01f;
const float rr = (float)state [stream].dc_re, ri = (float)state [stream].dc_im;
float2 &v = hist [(int64_t)stream * hist_len + i];
	v.x = fadd (v.x, fmul (sign, rr > lim ? lim : (rr < -lim ? -lim : rr)));
	v.y = fadd (v.y, fmul (sign, ri > lim ? lim : (ri < -lim ? -lim : ri)));
}

// inputFilter on: the fm-rate stage sees the block sums behind an fm-rate delay line of `len` entries, so
// StreamState::dc is the estimate of `len` fm samples ago.  RfDC as of NOW (metadata, DcValRf) = that state
// pushed through the sums still waiting in the line (oldest first).
__global__ void __launch_bounds__ (256)
dc_advance_kernel (const float2 *__restrict__ sdel, int len, double alpha, double beta,
                   const StreamState *__restrict__ state, double2 *__restrict__ out) {
__shared__ double sA [8], sB [8];
const int stream = blockIdx.x, tid = threadIdx.x;
const float2 *s = sdel + (int64_t)stream * len;
const int per = (len + 255) / 256, i0 = tid * per, i1 = min (i0 + per, len);
double A = 1.0, Br = 0.0, Bi = 0.0;
	for (int i = i0; i < i1; i ++) { Br = Br * beta + alpha * (double)s [i].x; Bi = Bi * beta + alpha * (double)s [i].y; A *= beta; }
double tr, ti;
	block_affine_start (A, Br, state [stream].dc_re, sA, sB, &tr);
	block_affine_start (A, Bi, state [stream].dc_im, sA, sB, &ti);
	if (tid == 0) out [stream] = make_double2 (tr, ti);
}

}	// namespace sdrjfm
