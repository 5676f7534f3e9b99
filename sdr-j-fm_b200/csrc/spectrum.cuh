// LF scope display spectrum: what ls_scope::processLFSpectrum
// (src/scopes-qwt6/ls-scope.cpp:76-92) does with every spectrumSize block of the LF scope stream
// (fmProcessor::processLfSpectrum, src/fm/fm-processor.cpp:906-912) — window (:50-52), radix-2 FFT,
// mapSpectrum (:130-176: magnitudes averaged over `factor` neighbouring bins, two- or one-sided),
// add_to_average (:178-193: displayBuffer = running average over averageCount blocks, in double).
//   lf_gather_kernel    the selected ELfPlot stream of this call as complex samples
//   lf_spectrum_kernel  one CTA per (block, stream): window, FFT in shared memory, bin magnitudes
//   lf_average_kernel   one thread per display bin walks the call's blocks in order (the recurrence)
//   lf_carry_kernel     the samples behind the last whole block wait for the next call
#pragma once
#include "common.cuh"
#include "frontend_poly.cuh"

namespace sdrjfm {

constexpr int kSpecThreads = 256;
constexpr int kSpecMaxN    = 4096;           // spectrumSize = 4 * displaySize, displaySize <= 1024 (radio.cpp:238-242)

struct LfGather {
	int32_t type;                             // ELfPlot
	int32_t n;                                // samples of this call (fm rate, or 24 kHz for the RDS types)
	const float2 *c_src; int64_t c_pitch;     // complex source (IF_FILTERED, RDS_INPUT)
	const float  *f_src; int64_t f_pitch;     // real source (demod, diffLR, pre-gain audio, Costas real part)
	const float  *i_src; int64_t i_pitch;     // imaginary parts (RDS_DEMOD)
	float   mul;                              // 20 (RDS_INPUT, :568), 4 (RDS_DEMOD, rds-decoder.cpp:77), else 1
};

__global__ void lf_gather_kernel (const LfGather G, float2 *__restrict__ out, int64_t out_pitch) {
const int stream = blockIdx.y;
const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= G.n) return;
float2 v = make_float2 (0.f, 0.f);
	if (G.c_src) { v = G.c_src [(int64_t)stream * G.c_pitch + i]; v.x = fmul (v.x, G.mul); v.y = fmul (v.y, G.mul); }
	else if (G.f_src) {
	   v.x = fmul (G.f_src [(int64_t)stream * G.f_pitch + i], G.mul);
	   if (G.i_src) v.y = fmul (G.i_src [(int64_t)stream * G.i_pitch + i], G.mul);
	}
	out [(int64_t)stream * out_pitch + i] = v;
}

struct LfSpecParams {
	int32_t N, logN;                          // spectrumSize
	int32_t display, factor, full;            // displaySize, bins per display point after the zoom, showFull
	int32_t n_carry;                          // samples carried in from the previous call
};

// blocks are cut from (carry | z).  Y: [S][cap_blk][display]
__global__ void __launch_bounds__ (kSpecThreads)
lf_spectrum_kernel (const float2 *__restrict__ z, int64_t pitch, const float2 *__restrict__ carry,
                    const float *__restrict__ window, const LfSpecParams P,
                    double *__restrict__ Y, int32_t cap_blk) {
extern __shared__ float2 sp_a [];
const int tid = threadIdx.x, stream = blockIdx.y, blk = blockIdx.x;
const int N = P.N;
const float2 *zs = z + (int64_t)stream * pitch;
const float2 *cs = carry + (int64_t)stream * N;
	for (int i = tid; i < N; i += kSpecThreads) {
	   const int64_t g = (int64_t)blk * N + i - P.n_carry;               // index into z (negative: carry)
	   float2 v = g >= 0 ? zs [g] : cs [P.n_carry + g];
	   const float m = (float)sqrt ((double)v.x * v.x + (double)v.y * v.y);
	   if (isinf (m) || isnan (m)) v = make_float2 (0.f, 0.f);           // :79-81
	   else { const float w = window [i]; v = make_float2 (fmul (v.x, w), fmul (v.y, w)); }
	   sp_a [__brev ((unsigned)i) >> (32 - P.logN)] = v;                 // decimation in time, fft-complex.cpp:73-98
	}
	__syncthreads ();
	for (int half = 1; half < N; half <<= 1) {
	   for (int b = tid; b < N / 2; b += kSpecThreads) {
	      const int pos = b & (half - 1);
	      const int i0 = ((b - pos) << 1) + pos;
	      float sn, cs2;
	      sincospif (-(float)pos / (float)half, &sn, &cs2);
	      const float2 u = sp_a [i0], v = sp_a [i0 + half];
	      const float2 t = make_float2 (v.x * cs2 - v.y * sn, v.x * sn + v.y * cs2);
	      sp_a [i0] = make_float2 (u.x + t.x, u.y + t.y);
	      sp_a [i0 + half] = make_float2 (u.x - t.x, u.y - t.y);
	   }
	   __syncthreads ();
	}
double *y = Y + ((int64_t)stream * cap_blk + blk) * P.display;
const int F = P.factor;
	for (int o = tid; o < P.display; o += kSpecThreads) {
	   double f = 0;
	   if (P.full) {
	      const int hd = P.display / 2;
	      if (o >= hd) {                                                  // 0 Hz .. rate/2 -> mid to end of the display
	         const int i = o - hd;
	         for (int j = 0; j < F; j ++) { const float2 v = sp_a [i * F + j]; f += (double)(float)sqrt ((double)v.x * v.x + (double)v.y * v.y); }
	      }
	      else {                                                          // rate/2 down to 0 Hz -> begin to mid
	         const int i = hd - 1 - o;
	         for (int j = 0; j < F; j ++) { const float2 v = sp_a [N - 1 - (i * F + j)]; f += (double)(float)sqrt ((double)v.x * v.x + (double)v.y * v.y); }
	      }
	   }
	   else
	      for (int j = 0; j < F; j ++) { const float2 v = sp_a [o * F + j]; f += (double)(float)sqrt ((double)v.x * v.x + (double)v.y * v.y); }
	   y [o] = f / F;
	}
}

// avg, disp: [S][display] (state).  refresh: the first block only primes the average (:181-185, :91-92)
__global__ void lf_average_kernel (const double *__restrict__ Y, int32_t cap_blk, int32_t nblk, int32_t display,
                                   int32_t average_count, int32_t refresh,
                                   double *__restrict__ avg, double *__restrict__ disp) {
const int stream = blockIdx.y;
const int o = blockIdx.x * blockDim.x + threadIdx.x;
	if (o >= display) return;
const double alpha = 1.0 / average_count, beta = (average_count - 1.0) / average_count;
double a = avg [(int64_t)stream * display + o];
double d = disp [(int64_t)stream * display + o];
	for (int b = 0; b < nblk; b ++) {
	   const double in = Y [((int64_t)stream * cap_blk + b) * display + o];
	   if (refresh && b == 0) a = in;
	   else { a = __dadd_rn (__dmul_rn (alpha, in), __dmul_rn (beta, a)); d = a; }
	}
	avg [(int64_t)stream * display + o] = a;
	disp [(int64_t)stream * display + o] = d;
}

// new carry = the last (n_carry + n) mod N samples of (carry | z)
__global__ void lf_carry_kernel (const float2 *__restrict__ z, int64_t pitch, int32_t n, int32_t N,
                                 const float2 *__restrict__ carry, int32_t n_carry, float2 *__restrict__ carry_new) {
const int stream = blockIdx.x;
const int total = n_carry + n, keep = total % N;
	for (int i = threadIdx.x; i < keep; i += blockDim.x) {
	   const int64_t g = (int64_t)(total - keep + i) - n_carry;
	   carry_new [(int64_t)stream * N + i] = g >= 0 ? z [(int64_t)stream * pitch + g] : carry [(int64_t)stream * N + n_carry + g];
	}
}

// ---- HF scope display spectrum: hs_scope::addElement (src/scopes-qwt6/hs-scope.cpp:102-151, 175-203) -------------
// The HF scope looks at the RAW input (hfBuffer, fm-processor.cpp:420): of every segment of segmentSize =
// inputRate / repeatRate samples the first spectrumSize = 4 displaySize are windowed (:114-120) and transformed;
// displayBuffer [display/2 + i] = mean of |F [4 i + j]|, displayBuffer [i] = mean of |F [N/2 + 4 i + j]| (:127-136),
// then doAverage (:190-199: averageCount stays 0, so the running average has the weight 1 / (repeatRate / 2)).
//   hf_gather_kernel    the part of the open segment's first N samples that this call delivers -> blk (converted)
//   hf_spectrum_kernel  one CTA per stream: window, FFT in shared memory, map, average (in place)
struct HfSpecParams { int32_t N, logN, display, half_freq; };

__global__ void hf_gather_kernel (const void *__restrict__ src, int64_t pitch, RawFmt rf, int64_t first, int32_t count,
                                  int32_t dst0, int32_t N, float2 *__restrict__ blk) {
const int stream = blockIdx.y;
const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
const void *xs = reinterpret_cast<const char *>(src) + (int64_t)stream * pitch * fmt_bytes (rf.fmt);
	blk [(int64_t)stream * N + dst0 + i] = load_iq_rt (xs, first + i, rf);
}

__global__ void __launch_bounds__ (kSpecThreads)
hf_spectrum_kernel (const float2 *__restrict__ blk, const float *__restrict__ window, const HfSpecParams P,
                    double *__restrict__ avg, double *__restrict__ disp) {
extern __shared__ float2 sp_a [];
const int tid = threadIdx.x, stream = blockIdx.x;
const int N = P.N;
const float2 *zs = blk + (int64_t)stream * N;
	for (int i = tid; i < N; i += kSpecThreads) {
	   float2 v = zs [i];
	   const float m = (float)sqrt ((double)v.x * v.x + (double)v.y * v.y);
	   if (isinf (m) || isnan (m)) v = make_float2 (0.f, 0.f);           // :115-118
	   else { const float w = window [i]; v = make_float2 (fmul (v.x, w), fmul (v.y, w)); }
	   sp_a [__brev ((unsigned)i) >> (32 - P.logN)] = v;
	}
	__syncthreads ();
	for (int half = 1; half < N; half <<= 1) {
	   for (int b = tid; b < N / 2; b += kSpecThreads) {
	      const int pos = b & (half - 1);
	      const int i0 = ((b - pos) << 1) + pos;
	      float sn, cs2;
	      sincospif (-(float)pos / (float)half, &sn, &cs2);
	      const float2 u = sp_a [i0], v = sp_a [i0 + half];
	      const float2 t = make_float2 (v.x * cs2 - v.y * sn, v.x * sn + v.y * cs2);
	      sp_a [i0] = make_float2 (u.x + t.x, u.y + t.y);
	      sp_a [i0 + half] = make_float2 (u.x - t.x, u.y - t.y);
	   }
	   __syncthreads ();
	}
const int ratio = N / P.display;
const double beta = (double)(P.half_freq - 1) / P.half_freq;            // :194
const float alpha = 1.0f / P.half_freq;                                 // :195
	for (int o = tid; o < P.display; o += kSpecThreads) {
	   const int hd = P.display / 2;
	   const int base = o >= hd ? (o - hd) * ratio : N / 2 + o * ratio;
	   float sum = 0.f;
	   for (int j = 0; j < ratio; j ++) { const float2 v = sp_a [base + j]; sum = fadd (sum, (float)sqrt ((double)v.x * v.x + (double)v.y * v.y)); }
	   double d = (double)fdiv (sum, (float)ratio);
	   if (d != d) d = 0;                                                   // :191-192
	   const double a = __dadd_rn (__dmul_rn (beta, avg [(int64_t)stream * P.display + o]), __dmul_rn ((double)alpha, d));
	   avg [(int64_t)stream * P.display + o] = a;
	   disp [(int64_t)stream * P.display + o] = a;
	}
}

}	// namespace sdrjfm
