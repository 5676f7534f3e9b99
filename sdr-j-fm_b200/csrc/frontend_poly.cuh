// K1g — the decimating front end for ANY integer decimation D and ANY device sample format.
//
// The reference builds its two decimators from the input rate (fm-processor.cpp:36,68-75):
// IRate = inputRate / 6, fmBand_1 = 25 taps /6, fmBand_2 = (IRate / fmRate + 1) taps
// / (IRate / fmRate).  Both kernels are a real prototype times a constant complex gain
// (tables.cpp), so for every rate the cascade is ONE real polyphase FIR decimating by
// D = 6 * (IRate / fmRate):   U[m] = sum_i C[i] x[D m + D - 1 - i]
//     2 304 000, 2 400 000 -> D = 12, 37 taps      (rtlsdr / the reference's own rate)
//     6 000 000            -> D = 30, 55 taps      (airspy)
//    10 000 000            -> D = 48, 73 taps      (sdrplay)
// and with inputFilter on (251 taps, delay 65285 = D q + 5 for all three D) one real FIR of
// 251 + ncomp - 1 (+5) taps followed by an fm-rate delay of q samples.
//
// The samples may arrive in the device's NATIVE format; the conversion the reference's
// device handlers do on the CPU before the ring buffer is fused into the only HBM read:
//     u8   (b - 127) / 128         rtlsdr-handler.cpp:286-293
//     s8   b / 128                 hackrf-handler.cpp:355-368
//     s16  v / denominator         sdrplay-handler.cpp:481-488 (2048 / 8192), sdrplay-handler-v3.cpp
//                                  :254-263 (2048 / 4096), pluto-handler.cpp:574-583, lime, airspy (2048)
// All denominators are powers of two, so the conversion is exact and commutes with nothing
// it should not: the floats entering the FIR are bit-identical to the handler's.
//
// Layout: as in frontend_fir.cuh — the tile is staged in shared memory in polyphase order
// (sample j at row j mod (D GPT), column j div (D GPT)), a thread owns GPT adjacent outputs
// (one column) and slides a GPT-wide register window over the tap groups; taps are
// constant-bank operands c_poly[p][g] = C'[D g + D - 1 - p].
#pragma once
#include "common.cuh"
#include "frontend_fir.cuh"

namespace sdrjfm {

constexpr int kPolyMaxTaps = 352;                   // max D * NG over the instantiated shapes
__constant__ float c_poly [kPolyMaxTaps];

// kFmtAirspy: s16 IQ at the airspy's native rate, brought to 2.304 MS/s by the handler's own
// per-millisecond linear interpolation (airspy-handler.cpp:117-128, 283-309): block b = native samples
// b B .. (b + 1) B (B = native rate / 1000), output j of the block =
// x[b B + mapInt[j] + 1] * mapFrac[j] + x[b B + mapInt[j]] * (1 - mapFrac[j]).  Sample index n of the
// kernel is the OUTPUT index; the tables are the reference's (built on the host with its formulas).
enum { kFmtCF32 = 0, kFmtU8 = 1, kFmtS8 = 2, kFmtS16 = 3, kFmtAirspy = 4 };
constexpr int kAirspyOut = 2304;                    // output samples per millisecond block

struct RawFmt {
	int32_t fmt;               // kFmt*
	float   scale;             // 1 / denominator (u8, s8: 1/128)
	int32_t blk;               // kFmtAirspy: native samples per block
	const int16_t *map_int;    // kFmtAirspy: mapTable_int [2304]
	const float   *map_frac;   // kFmtAirspy: mapTable_float [2304]
};

__host__ __device__ constexpr int fmt_bytes (int fmt) {
	return fmt == kFmtCF32 ? 8 : (fmt == kFmtS16 || fmt == kFmtAirspy) ? 4 : 2;
}

template <int FMT>
__device__ __forceinline__ float2 load_iq (const void *base, int64_t n, const RawFmt &rf) {
const float scale = rf.scale;
	if (FMT == kFmtAirspy) {
	   const int b = (int)n / kAirspyOut, j = (int)n - b * kAirspyOut;
	   const short2 *p = reinterpret_cast<const short2 *>(base) + ((int64_t)b * rf.blk + rf.map_int [j]);
	   const short2 s0 = __ldg (p), s1 = __ldg (p + 1);
	   const float r = rf.map_frac [j], r1 = fsub (1.0f, r);
	   const float2 x0 = make_float2 ((float)s0.x * scale, (float)s0.y * scale);
	   const float2 x1 = make_float2 ((float)s1.x * scale, (float)s1.y * scale);
	   return make_float2 (fadd (fmul (x1.x, r), fmul (x0.x, r1)), fadd (fmul (x1.y, r), fmul (x0.y, r1)));
	}
	if (FMT == kFmtU8) {
	   const uchar2 v = __ldcs (reinterpret_cast<const uchar2 *>(base) + n);
	   return make_float2 ((float)((int)v.x - 127) * scale, (float)((int)v.y - 127) * scale);
	}
	if (FMT == kFmtS8) {
	   const char2 v = __ldcs (reinterpret_cast<const char2 *>(base) + n);
	   return make_float2 ((float)v.x * scale, (float)v.y * scale);
	}
	if (FMT == kFmtS16) {
	   const short2 v = __ldcs (reinterpret_cast<const short2 *>(base) + n);
	   return make_float2 ((float)v.x * scale, (float)v.y * scale);
	}
	return __ldcs (reinterpret_cast<const float2 *>(base) + n);
}

__device__ __forceinline__ float2 load_iq_rt (const void *base, int64_t n, const RawFmt &rf) {
	switch (rf.fmt) {
	   case kFmtU8:  return load_iq<kFmtU8> (base, n, rf);
	   case kFmtS8:  return load_iq<kFmtS8> (base, n, rf);
	   case kFmtS16: return load_iq<kFmtS16> (base, n, rf);
	   case kFmtAirspy: return load_iq<kFmtAirspy> (base, n, rf);
	   default:      return load_iq<kFmtCF32> (base, n, rf);
	}
}

constexpr int poly_batch (int loads) {
	return loads % 16 == 0 ? 16 : loads % 15 == 0 ? 15 : loads % 12 == 0 ? 12 : loads % 10 == 0 ? 10 : 8;
}

template <int D_, int GPT_, int NG_>
struct Poly {
	static constexpr int D = D_, GPT = GPT_, NG = NG_;
	static constexpr int Rows    = D * GPT;                        // polyphase rows = loads per thread
	static constexpr int Halo    = (NG - 1 + GPT - 1) / GPT;       // halo columns (older outputs)
	static constexpr int Pitch   = (kFeThreads + Halo) | 1;        // odd: conflict-free 8-byte row strides
	static constexpr int TileOut = kFeThreads * GPT;
	static constexpr int TileIn  = TileOut * D;
	static constexpr int HaloIn  = Halo * Rows;                    // raw samples before the tile
	static constexpr int SmemBytes = Rows * Pitch * (int)sizeof (float2);
	static constexpr int Batch   = poly_batch (Rows);
	static constexpr int RawBytes = TileOut * kRawSlots * (int)sizeof (float2);   // oscillator on: raw block-sum slots behind the tile
	static constexpr int MinCtas = (226 * 1024) / (SmemBytes + 1024) > 8 ? 8 : (226 * 1024) / (SmemBytes + 1024);
	static_assert (D * NG <= kPolyMaxTaps, "tap table too small");
	static_assert (Rows % Batch == 0, "batching");
};

template <class P, int FMT>
__device__ __forceinline__ void poly_stage (float2 *sm, float2 *sRaw, const void *xs, int64_t in0, int64_t N,
                                            const float2 *hist_s, bool first_tile, const LoParams &lop,
                                            bool lo, const RawFmt &rf, int tid) {
//	halo: sample in0 - HaloIn + i sits at row i % Rows, column i / Rows
	for (int i = tid; i < P::HaloIn; i += kFeThreads) {
	   float2 v;
	   if (first_tile) v = hist_s [i];
	   else            v = load_iq<FMT> (xs, in0 - P::HaloIn + i, rf);
	   if (lo) v = lo_apply (lop, v, lo_index (lop, in0 - P::HaloIn + i));
	   sm [(i % P::Rows) * P::Pitch + i / P::Rows] = v;
	}
int32_t loIdx = lo ? lo_index (lop, in0 + tid) : 0;
#pragma unroll
	for (int b = 0; b < P::Rows / P::Batch; b ++) {
	   float2 v [P::Batch];
#pragma unroll
	   for (int k = 0; k < P::Batch; k ++) {
	      const int j = (b * P::Batch + k) * kFeThreads + tid;
	      const int64_t n = in0 + j;
	      v [k] = (n < N) ? load_iq<FMT> (xs, n, rf) : make_float2 (0.f, 0.f);
	   }
	   if (lo) {
#pragma unroll
	      for (int k = 0; k < P::Batch; k ++) {
	         const int j = (b * P::Batch + k) * kFeThreads + tid;
	         // the RF DC estimate follows the RAW samples: block sums before gain and rotation
	         raw_block_sum (sRaw, j, v [k], P::D);
	         v [k] = lo_apply (lop, v [k], loIdx);
	         loIdx -= lop.step128; if (loIdx < 0) loIdx += lop.rate;
	      }
	   }
#pragma unroll
	   for (int k = 0; k < P::Batch; k ++) {
	      const int j = (b * P::Batch + k) * kFeThreads + tid;
	      const int col = j / P::Rows;
	      const int row = j - col * P::Rows;
	      sm [row * P::Pitch + col + P::Halo] = v [k];
	   }
	}
}

// x      : [n_streams][in_pitch] samples in format rf.fmt, this call's samples (N = D * M per stream)
// hist   : [n_streams][hist_len] float2 (converted, before gain / oscillator), the samples preceding x[.][0]
// U, S   : [n_streams][out_pitch] complex: FIR output and plain D-sample block sums
template <int D, int GPT, int NG>
__global__ void __launch_bounds__ (kFeThreads, Poly<D, GPT, NG>::MinCtas)
frontend_poly_kernel (const void *__restrict__ x, int64_t in_pitch, RawFmt rf,
                      const float2 *__restrict__ hist, int hist_len,
                      float2 *__restrict__ U, float2 *__restrict__ S,
                      int64_t out_pitch, int32_t M, const LoParams lop, int32_t tile0) {
typedef Poly<D, GPT, NG> P;
extern __shared__ float2 sm [];
float2 *sRaw = sm + P::Rows * P::Pitch;              // present only when the oscillator is on (launch adds RawBytes)
const int tid    = threadIdx.x;
const int stream = blockIdx.y;
const bool lo    = lop.tab != nullptr;
	if (lo) { for (int i = tid; i < P::TileOut * kRawSlots; i += kFeThreads) sRaw [i] = make_float2 (0.f, 0.f); __syncthreads (); }
const int64_t out0 = (int64_t)(blockIdx.x + tile0) * P::TileOut;      // tile0: the tiles before it were done by K1t
const int64_t in0  = out0 * D;
const int64_t N    = (int64_t)M * D;
const void *xs = reinterpret_cast<const char *>(x) + (int64_t)stream * in_pitch * fmt_bytes (rf.fmt);
const float2 *hs = hist + (int64_t)stream * hist_len + (hist_len - P::HaloIn);
const bool first = blockIdx.x + tile0 == 0;
	switch (rf.fmt) {
	   case kFmtU8:  poly_stage<P, kFmtU8>  (sm, sRaw, xs, in0, N, hs, first, lop, lo, rf, tid); break;
	   case kFmtS8:  poly_stage<P, kFmtS8>  (sm, sRaw, xs, in0, N, hs, first, lop, lo, rf, tid); break;
	   case kFmtS16: poly_stage<P, kFmtS16> (sm, sRaw, xs, in0, N, hs, first, lop, lo, rf, tid); break;
	   case kFmtAirspy: poly_stage<P, kFmtAirspy> (sm, sRaw, xs, in0, N, hs, first, lop, lo, rf, tid); break;
	   default:      poly_stage<P, kFmtCF32>(sm, sRaw, xs, in0, N, hs, first, lop, lo, rf, tid); break;
	}
	__syncthreads ();

//	thread t -> outputs GPT t .. GPT t + GPT - 1.  Output GPT t + q (q < 0: older) sits in column
//	t + Halo + floor (q / GPT), rows D * (q mod GPT) + p.
float2 acc [GPT], dcs [GPT];
#pragma unroll
	for (int k = 0; k < GPT; k ++) { acc [k] = make_float2 (0.f, 0.f); dcs [k] = make_float2 (0.f, 0.f); }
const float2 *col0 = sm + tid + P::Halo;
#pragma unroll (NG <= 4 ? D : 1)
	for (int p = 0; p < D; p ++) {
	   float2 w [GPT];                          // w[k] = phase p of output GPT t + k - (groups done)
#pragma unroll
	   for (int k = 0; k < GPT; k ++) {
	      w [k] = col0 [(D * k + p) * P::Pitch];
	      dcs [k].x += w [k].x; dcs [k].y += w [k].y;
	   }
#pragma unroll
	   for (int g = 0; g < NG; g ++) {
	      const float c = c_poly [p * NG + g];
#pragma unroll
	      for (int k = 0; k < GPT; k ++) acc [k] = ffma2 (c, w [k], acc [k]);
	      if (g + 1 < NG) {
#pragma unroll
	         for (int k = GPT - 1; k > 0; k --) w [k] = w [k - 1];
	         const int q  = -1 - g;                       // next older output relative to GPT t
	         const int cq = -((g + GPT) / GPT);           // floor (q / GPT)
	         const int rq = q - GPT * cq;                 // q mod GPT
	         w [0] = col0 [cq + (D * rq + p) * P::Pitch];
	      }
	   }
	}

const int64_t m0 = out0 + (int64_t)tid * GPT;
float2 *Us = U + (int64_t)stream * out_pitch;
float2 *Ss = S + (int64_t)stream * out_pitch;
	if (lo) {
#pragma unroll
	   for (int k = 0; k < GPT; k ++) dcs [k] = raw_block_total (sRaw, tid * GPT + k);
	}
	if (GPT % 2 == 0 && m0 + GPT <= M && (out_pitch & 1) == 0) {
#pragma unroll
	   for (int k = 0; k + 1 < GPT; k += 2) {
	      *reinterpret_cast<float4 *>(Us + m0 + k) = make_float4 (acc [k].x, acc [k].y, acc [k + 1].x, acc [k + 1].y);
	      *reinterpret_cast<float4 *>(Ss + m0 + k) = make_float4 (dcs [k].x, dcs [k].y, dcs [k + 1].x, dcs [k + 1].y);
	   }
	}
	else {
#pragma unroll
	   for (int k = 0; k < GPT; k ++)
	      if (m0 + k < M) { Us [m0 + k] = acc [k]; Ss [m0 + k] = dcs [k]; }
	}
}

// After the front end has run: roll the raw-sample history forward, converting to float.
// new_hist[i] is the sample at position n_proc - hist_len + i of (old_hist | x[0..n_proc)).
__global__ void roll_history_raw_kernel (const void *__restrict__ x, int64_t in_pitch, RawFmt rf,
                                         const float2 *__restrict__ old_hist,
                                         float2 *__restrict__ new_hist, int64_t n_proc, int hist_len) {
const int stream = blockIdx.y;
const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hist_len) return;
const int64_t pos = n_proc - hist_len + i;
const void *xs = reinterpret_cast<const char *>(x) + (int64_t)stream * in_pitch * fmt_bytes (rf.fmt);
	new_hist [(int64_t)stream * hist_len + i] =
	      pos >= 0 ? load_iq_rt (xs, pos, rf)
	               : old_hist [(int64_t)stream * hist_len + (hist_len + pos)];
}

}	// namespace sdrjfm
