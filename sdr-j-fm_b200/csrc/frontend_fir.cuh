// K1 — decimating front end: complex IQ at the input rate -> fm-rate samples.
//
// Replaces, per input sample, DecimatingFIR::Pass of fmBand_1 (25 taps, /6) followed by
// fmBand_2 (3 taps, /2) (src/various/fir-filters.cpp:397-424, constructed at
// src/fm/fm-processor.cpp:68-75) — "the real cpu killer" — and prepares the RF DC
// removal of fm-processor.cpp:423-446 so that it can be finished at the fm rate.
//
// Formulation: both reference kernels are real prototypes times a constant
// complex gain, so the cascade is ONE real 37-tap polyphase FIR decimating by 12,
//     U[m] = sum_{i<37} C[i] * x[12 m + 11 - i]          (index contract: bit-exact)
// with the constant gain G applied downstream.  DC removal is linear and its estimate
// moves by < 5e-7 per sample, so the kernel only emits the plain 12-sample sums
//     S[m] = sum_{i<12} x[12 m + i]
// from which the fm-rate stage advances the one-pole DC estimate and subtracts
// clamp (RfDC) * sum (C) (+ first-order trend) from U.
//
// Data movement: one coalesced 8-byte load per input sample (the only HBM read of the IQ
// stream), staged in shared memory in POLYPHASE order — sample n of the tile lives at
// row (n mod 48), column (n div 48) — so that a thread producing four adjacent outputs
// reads every operand with unit lane stride (conflict-free) and every staged value is
// reused from registers for up to 4x3 taps.  Taps sit in constant memory and enter the
// FMAs as immediate constant-bank operands.
#pragma once
#include "common.cuh"

namespace sdrjfm {

constexpr int kFeThreads = 128;
constexpr int kFeGpt     = 4;                       // fm-rate outputs per thread
constexpr int kFeTileOut = kFeThreads * kFeGpt;     // 512 outputs per CTA
constexpr int kFeTileIn  = kFeTileOut * kDecim;     // 6144 input samples per CTA (48 KB)
constexpr int kFeRows    = kDecim * kFeGpt;         // 48 polyphase rows
constexpr int kFePitch   = kFeThreads + 1;          // 129 columns: column 0 is the halo
constexpr int kFeSmemBytes = kFeRows * kFePitch * (int)sizeof (float2);   // 49536

__constant__ float c_comp [40];                     // composite taps C[0..36]

// Local oscillator (fm-processor.cpp:462-466, oscillator.cpp:26-58): before the filters every
// sample is scaled per component (IQ gain) and multiplied by Table[LOPhase], LOPhase stepping
// by -lo per sample modulo inputRate.  tab = the reference's inputRate-entry table.
struct LoParams {
	const float2 *tab;         // nullptr: LO off (lo = 0 multiplies by (1,0): identity)
	int32_t rate;              // inputRate
	int32_t lo;                // loFrequency (Hz = table steps per sample)
	int32_t step128;           // (128 * lo) mod rate, in [0, rate)
	int64_t phase;             // LOPhase after the sample before src[0]
	float   lgain, rgain;
};

__device__ __forceinline__ int32_t lo_index (const LoParams &L, int64_t n) {     // sample src[n]
int64_t t = (L.phase - (int64_t)L.lo * (n + 1)) % L.rate;
	return (int32_t)(t < 0 ? t + L.rate : t);
}
__device__ __forceinline__ float2 lo_apply (const LoParams &L, float2 v, int32_t idx) {
	return cmul_rn (make_float2 (fmul (v.x, L.lgain), fmul (v.y, L.rgain)), L.tab [idx]);
}

// Block sums of the RAW samples when the oscillator is on (the RF DC estimate follows the samples
// before gain and rotation), WITHOUT atomics: lanes hold consecutive samples j, a block of D
// consecutive samples is cut only at multiples of 32, i.e. into at most kRawSlots warp segments.
// Each segment is summed by a shuffle scan in a fixed order and stored in ITS slot; the slots are
// added in a fixed order at the end, so the result never depends on timing.
constexpr int kRawSlots = 3;                        // D <= 64
__device__ __forceinline__ void raw_block_sum (float2 *sPart /* [blocks][kRawSlots] */, int j, float2 v, int D) {
const int lane = threadIdx.x & 31;
const int blk = j / D;
float2 s = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
	   const float ox = __shfl_up_sync (0xffffffffu, s.x, d), oy = __shfl_up_sync (0xffffffffu, s.y, d);
	   const int ob = __shfl_up_sync (0xffffffffu, blk, d);
	   if (lane >= d && ob == blk) { s.x += ox; s.y += oy; }
	}
const int nb = __shfl_down_sync (0xffffffffu, blk, 1);
	if (lane == 31 || nb != blk) {
	   const int jfirst = max (blk * D, j - lane);      // block start or warp start
	   sPart [blk * kRawSlots + ((jfirst >> 5) - ((blk * D) >> 5))] = s;
	}
}
__device__ __forceinline__ float2 raw_block_total (const float2 *sPart, int blk) {
float2 t = sPart [blk * kRawSlots];
#pragma unroll
	for (int q = 1; q < kRawSlots; q ++) { t.x += sPart [blk * kRawSlots + q].x; t.y += sPart [blk * kRawSlots + q].y; }
	return t;
}

// x      : [n_streams][in_pitch] complex, this call's samples (N = 12 * M per stream)
// hist   : [n_streams][hist_len] complex, the raw samples preceding x[.][0] (the last 36 are used)
// U, S   : [n_streams][out_pitch] complex
template <bool LO>
__global__ void __launch_bounds__ (kFeThreads, 4)
frontend_fir_kernel (const float2 *__restrict__ x, int64_t in_pitch,
                     const float2 *__restrict__ hist, int hist_len,
                     float2 *__restrict__ U, float2 *__restrict__ S,
                     int64_t out_pitch, int32_t M, const LoParams lop, int32_t tile0) {
extern __shared__ float2 sm [];
__shared__ float2 sRaw [LO ? kFeTileOut * kRawSlots : 1];
	if (LO) { for (int i = threadIdx.x; i < kFeTileOut * kRawSlots; i += kFeThreads) sRaw [i] = make_float2 (0.f, 0.f); __syncthreads (); }
const int tid    = threadIdx.x;
const int stream = blockIdx.y;
const int tile = blockIdx.x + tile0;                     // tile0: the tiles before it were done by K1t
const int64_t out0 = (int64_t)tile * kFeTileOut;         // first output of the tile
const int64_t in0  = out0 * kDecim;                       // first input of the tile
const int64_t N    = (int64_t)M * kDecim;
const float2 *xs = x + (int64_t)stream * in_pitch;

//	halo: the 36 samples before the tile, polyphase rows 12..47 of column 0
	if (tid < kHist) {
	   float2 v;
	   if (tile == 0) v = hist [(int64_t)stream * hist_len + (hist_len - kHist) + tid];
	   else                 v = xs [in0 - kHist + tid];
	   if (LO) v = lo_apply (lop, v, lo_index (lop, in0 - kHist + tid));
	   sm [(kDecim + tid) * kFePitch] = v;
	}
int32_t loIdx = LO ? lo_index (lop, in0 + tid) : 0;

//	body: 48 coalesced 8-byte loads per thread, issued in batches so that 16 are in flight
#pragma unroll
	for (int b = 0; b < 3; b ++) {
	   float2 v [16];
#pragma unroll
	   for (int k = 0; k < 16; k ++) {
	      const int j = (b * 16 + k) * kFeThreads + tid;
	      const int64_t n = in0 + j;
	      v [k] = (n < N) ? __ldcs (xs + n) : make_float2 (0.f, 0.f);
	      if (LO) {
	         // the RF DC estimate follows the RAW samples: block sums before gain and rotation
	         raw_block_sum (sRaw, j, v [k], kDecim);
	         v [k] = lo_apply (lop, v [k], loIdx);
	         loIdx -= lop.step128; if (loIdx < 0) loIdx += lop.rate;
	      }
	   }
#pragma unroll
	   for (int k = 0; k < 16; k ++) {
	      const int j = (b * 16 + k) * kFeThreads + tid;
	      const int col = j / kFeRows;
	      const int row = j - col * kFeRows;
	      sm [row * kFePitch + col + 1] = v [k];
	   }
	}
	__syncthreads ();

//	thread t -> outputs 4t..4t+3 of the tile.  Column t+1 holds their own 48 samples
//	(row 12 k + p = phase p of output k), column t rows 12..47 hold the three outputs before.
float2 acc [kFeGpt], dcs [kFeGpt];
#pragma unroll
	for (int k = 0; k < kFeGpt; k ++) {
	   acc [k] = make_float2 (0.f, 0.f);
	   dcs [k] = make_float2 (0.f, 0.f);
	}
const float2 *colp = sm + tid;        // previous column
const float2 *colc = sm + tid + 1;    // own column
#pragma unroll
	for (int p = 0; p < kDecim; p ++) {
	   float2 v [6];
	   v [0] = colp [(24 + p) * kFePitch];
	   v [1] = colp [(36 + p) * kFePitch];
	   v [2] = colc [(p) * kFePitch];
	   v [3] = colc [(12 + p) * kFePitch];
	   v [4] = colc [(24 + p) * kFePitch];
	   v [5] = colc [(36 + p) * kFePitch];
	   const float c0 = c_comp [11 - p], c1 = c_comp [23 - p], c2 = c_comp [35 - p];
#pragma unroll
	   for (int k = 0; k < kFeGpt; k ++) {
	      acc [k] = ffma2 (c0, v [k + 2], acc [k]);
	      acc [k] = ffma2 (c1, v [k + 1], acc [k]);
	      acc [k] = ffma2 (c2, v [k], acc [k]);
	      dcs [k].x += v [k + 2].x;
	      dcs [k].y += v [k + 2].y;
	   }
	   if (p == kDecim - 1) {            // tap 36 reaches phase 11 three outputs back
	      const float c3 = c_comp [36];
	      const float2 w = colp [(12 + p) * kFePitch];
	      acc [0] = ffma2 (c3, w, acc [0]);      acc [1] = ffma2 (c3, v [0], acc [1]);
	      acc [2] = ffma2 (c3, v [1], acc [2]);  acc [3] = ffma2 (c3, v [2], acc [3]);
	   }
	}

const int64_t m0 = out0 + (int64_t)tid * kFeGpt;
float2 *Us = U + (int64_t)stream * out_pitch;
float2 *Ss = S + (int64_t)stream * out_pitch;
	if (LO) {
#pragma unroll
	   for (int k = 0; k < kFeGpt; k ++) dcs [k] = raw_block_total (sRaw, tid * kFeGpt + k);
	}
	if (m0 + kFeGpt <= M && (out_pitch & 1) == 0) {
	   float4 *u4 = reinterpret_cast<float4 *>(Us + m0);
	   float4 *s4 = reinterpret_cast<float4 *>(Ss + m0);
	   u4 [0] = make_float4 (acc [0].x, acc [0].y, acc [1].x, acc [1].y);
	   u4 [1] = make_float4 (acc [2].x, acc [2].y, acc [3].x, acc [3].y);
	   s4 [0] = make_float4 (dcs [0].x, dcs [0].y, dcs [1].x, dcs [1].y);
	   s4 [1] = make_float4 (dcs [2].x, dcs [2].y, dcs [3].x, dcs [3].y);
	}
	else {
#pragma unroll
	   for (int k = 0; k < kFeGpt; k ++)
	      if (m0 + k < M) { Us [m0 + k] = acc [k]; Ss [m0 + k] = dcs [k]; }
	}
}

// After the front end has run: roll the raw-sample history forward.  new_hist[i] is the
// sample at position n_proc - 36 + i of the concatenation (old_hist | x[0..n_proc)).
__global__ void roll_history_kernel (const float2 *__restrict__ x, int64_t in_pitch,
                                     const float2 *__restrict__ old_hist,
                                     float2 *__restrict__ new_hist, int64_t n_proc, int hist_len) {
const int stream = blockIdx.y;
const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hist_len) return;
const int64_t pos = n_proc - hist_len + i;
	new_hist [(int64_t)stream * hist_len + i] =
	      pos >= 0 ? x [(int64_t)stream * in_pitch + pos]
	               : old_hist [(int64_t)stream * hist_len + (hist_len + pos)];
}

// ---- input filter ON ------------------------------------------------------------------------
// inputFilter = fftFilter (65536, 251) (fm-processor.cpp:77,397-401,469-470) is a linear
// convolution with 251 real low-pass taps delayed by NumofSamples = 65285 input samples
// (SURVEY.md §8(a) a4).  Cascaded with the decimators it is ONE real FIR of 251 + 37 - 1 = 287
// taps Cw decimating by 12:   z[m] = G sum_t Cw[t] x[12 m + 11 - 65285 - t].
// 65285 = 12 * 5440 + 5, so with taps shifted by 5 (Cws[t'] = Cw[t' - 5], 292 taps)
//     F[m] = sum_{t'<292} Cws[t'] x[12 m + 11 - t'],      z[m] = G F[m - 5440]:
// the same polyphase form as the narrow kernel with 25 tap groups instead of 3, followed by
// a pure fm-rate delay of 5440 samples (fm_delay_kernel).  ~600 FMA per output: this variant
// is FP32-bound, not HBM-bound.
constexpr int kFwGroups  = 25;                      // 12 * 25 = 300 >= 292 taps
constexpr int kFwHist    = 12 * kFwGroups;          // 300 raw samples of history kept per stream
constexpr int kFwHalo    = 6;                       // halo columns: 24 outputs = 288 samples
constexpr int kFwPitch   = kFeThreads + kFwHalo + 1;
constexpr int kFwSmemBytes = kFeRows * kFwPitch * (int)sizeof (float2);
constexpr int kFwDelay   = 5440;                    // fm-rate samples

__constant__ float c_wide [kDecim][kFwGroups + 3];  // c_wide[p][g] = C'ws[12 g + 11 - p]

template <bool LO>
__global__ void __launch_bounds__ (kFeThreads, 3)
frontend_wide_kernel (const float2 *__restrict__ x, int64_t in_pitch,
                      const float2 *__restrict__ hist, int hist_len,
                      float2 *__restrict__ U, float2 *__restrict__ S,
                      int64_t out_pitch, int32_t M, const LoParams lop) {
extern __shared__ float2 sm [];
__shared__ float2 sRaw [LO ? kFeTileOut * kRawSlots : 1];
	if (LO) { for (int i = threadIdx.x; i < kFeTileOut * kRawSlots; i += kFeThreads) sRaw [i] = make_float2 (0.f, 0.f); __syncthreads (); }
const int tid    = threadIdx.x;
const int stream = blockIdx.y;
const int64_t out0 = (int64_t)blockIdx.x * kFeTileOut;
const int64_t in0  = out0 * kDecim;
const int64_t N    = (int64_t)M * kDecim;
const float2 *xs = x + (int64_t)stream * in_pitch;
constexpr int kHaloIn = kFwHalo * kFeRows;           // 288 samples before the tile

//	halo: columns 0..5 (sample in0 - 288 + i sits at row i % 48, column i / 48)
	for (int i = tid; i < kHaloIn; i += kFeThreads) {
	   float2 v;
	   if (blockIdx.x == 0) v = hist [(int64_t)stream * hist_len + (hist_len - kHaloIn) + i];
	   else                 v = xs [in0 - kHaloIn + i];
	   if (LO) v = lo_apply (lop, v, lo_index (lop, in0 - kHaloIn + i));
	   sm [(i % kFeRows) * kFwPitch + i / kFeRows] = v;
	}
int32_t loIdx = LO ? lo_index (lop, in0 + tid) : 0;
#pragma unroll
	for (int b = 0; b < 3; b ++) {
	   float2 v [16];
#pragma unroll
	   for (int k = 0; k < 16; k ++) {
	      const int j = (b * 16 + k) * kFeThreads + tid;
	      const int64_t n = in0 + j;
	      v [k] = (n < N) ? __ldcs (xs + n) : make_float2 (0.f, 0.f);
	      if (LO) {
	         // the RF DC estimate follows the RAW samples: block sums before gain and rotation
	         raw_block_sum (sRaw, j, v [k], kDecim);
	         v [k] = lo_apply (lop, v [k], loIdx);
	         loIdx -= lop.step128; if (loIdx < 0) loIdx += lop.rate;
	      }
	   }
#pragma unroll
	   for (int k = 0; k < 16; k ++) {
	      const int j = (b * 16 + k) * kFeThreads + tid;
	      const int col = j / kFeRows;
	      const int row = j - col * kFeRows;
	      sm [row * kFwPitch + col + kFwHalo] = v [k];
	   }
	}
	__syncthreads ();

//	thread t -> outputs 4t..4t+3; output index 4t + q (q may be negative) sits in column
//	t + kFwHalo + floor (q / 4), rows 12 * (q mod 4) + p
float2 acc [kFeGpt], dcs [kFeGpt];
#pragma unroll
	for (int k = 0; k < kFeGpt; k ++) { acc [k] = make_float2 (0.f, 0.f); dcs [k] = make_float2 (0.f, 0.f); }
const float2 *col0 = sm + tid + kFwHalo;
#pragma unroll 1
	for (int p = 0; p < kDecim; p ++) {
	   float2 w3 = col0 [(36 + p) * kFwPitch];       // q = 3
	   float2 w2 = col0 [(24 + p) * kFwPitch];       // q = 2
	   float2 w1 = col0 [(12 + p) * kFwPitch];       // q = 1
	   float2 w0 = col0 [(p) * kFwPitch];            // q = 0
	   dcs [3].x += w3.x; dcs [3].y += w3.y; dcs [2].x += w2.x; dcs [2].y += w2.y;
	   dcs [1].x += w1.x; dcs [1].y += w1.y; dcs [0].x += w0.x; dcs [0].y += w0.y;
#pragma unroll
	   for (int g = 0; g < kFwGroups; g ++) {
	      const float c = c_wide [p][g];
	      acc [3] = ffma2 (c, w3, acc [3]); acc [2] = ffma2 (c, w2, acc [2]);
	      acc [1] = ffma2 (c, w1, acc [1]); acc [0] = ffma2 (c, w0, acc [0]);
	      w3 = w2; w2 = w1; w1 = w0;
	      if (g + 1 < kFwGroups) {
	         const int q = -1 - g;                        // next older output index relative to 4t
	         const int cq = (q - 3) / 4;                  // floor (q / 4) for negative q
	         const int rq = q - 4 * cq;                   // q mod 4 in 0..3
	         w0 = col0 [cq + (12 * rq + p) * kFwPitch];
	      }
	   }
	}
const int64_t m0 = out0 + (int64_t)tid * kFeGpt;
float2 *Us = U + (int64_t)stream * out_pitch;
float2 *Ss = S + (int64_t)stream * out_pitch;
#pragma unroll
	for (int k = 0; k < kFeGpt; k ++)
	   if (m0 + k < M) { Us [m0 + k] = acc [k]; Ss [m0 + k] = LO ? raw_block_total (sRaw, tid * kFeGpt + k) : dcs [k]; }
}

// fm-rate delay line: out[m] = (hist | in)[m], new_hist = the last D entries of (hist | in)
__global__ void fm_delay_kernel (const float2 *__restrict__ in, int64_t pitch, int32_t M, int D,
                                 const float2 *__restrict__ hist, float2 *__restrict__ new_hist,
                                 float2 *__restrict__ out) {
const int stream = blockIdx.y;
const int i = blockIdx.x * blockDim.x + threadIdx.x;
const float2 *is = in + (int64_t)stream * pitch;
const float2 *hs = hist + (int64_t)stream * D;
	if (i < M) out [(int64_t)stream * pitch + i] = i < D ? hs [i] : is [i - D];
	if (i < D) {
	   const int j = M + i;                              // index into (hist | in)
	   new_hist [(int64_t)stream * D + i] = j < D ? hs [j] : is [j - D];
	}
}

}	// namespace sdrjfm
