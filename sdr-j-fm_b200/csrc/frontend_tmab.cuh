// K1tb — the TMA-fed front end for the DEVICE sample formats and the rates whose row is not 384 bytes.
//
// frontend_tma.cuh feeds complex-float rows of 48 samples through a 128-byte-swizzled tensor map.  The
// device handlers' formats (frontend_poly.cuh) are 2 or 4 bytes per IQ sample,
// and 6 MS/s has rows of 60 samples: here the stream is described to TMA as a plain 3-D tensor of
// 32-bit words [row words][rows][streams], no swizzle, a tile = 128 rows + the row before it, and the
// integer -> float conversion of the reference's handlers happens when a thread reads ITS row from shared
// memory with 128-bit loads — no load, convert or store-to-shared instruction is spent per sample on the
// way in (the generic kernel K1g spends ~14, which bounds it at 0.2-0.35 of the HBM peak).
//
// Conversion (bit-identical to the handlers', all divisors are powers of two):
//     u8   (b - 127) / 128     as_float (0x47800000 | b) - (65536 + 127/128)       rtlsdr-handler.cpp:286-293
//     s8   b / 128             as_float (0x47800000 | (b ^ 0x80)) - (65536 + 1)    hackrf-handler.cpp:355-368
//     s16  v / 2^k             as_float ((150 - k) << 23 | (v ^ 0x8000)) - (2^(23-k) + 2^(15-k))   sdrplay / pluto / lime
// (the integer sits in the low mantissa bits of a float whose ulp is the handler's scale, so ONE byte
// permute and ONE exact subtraction per component replace I2F + FMUL.)
//
// Arithmetic, tap tables, accumulation order and outputs are those of frontend_tma_kernel, so a stream
// delivered as bytes gives bit-identical results to the same stream delivered as the floats the handler
// would have made of it (tests/test_gpu_rates_formats.py).
#pragma once
#include "frontend_tma.cuh"

namespace sdrjfm {


template <int D, int GPT, int NT, int FMT>
struct Fb {
	static constexpr int RowSamples = D * GPT;                     // 48 (D = 12, 48), 60 (D = 30) or 40 (D = 5: resampler stage A)
	static constexpr int Prev       = (NT - D + RowSamples - 1) / RowSamples;   // rows of FIR history in front of a thread's own row
	static constexpr int Bps        = FMT == kFmtCF32 ? 8 : (FMT == kFmtS16 ? 4 : 2);
	static constexpr int RowBytes   = RowSamples * Bps;
	static constexpr int PerChunk   = 16 / Bps;                     // samples per 16-byte chunk
	static constexpr int Chunks     = RowBytes / 16;
	static constexpr int BoxRows    = kFtRows + Prev;
	static constexpr int StageBytes = BoxRows * RowBytes;
	static constexpr int StageStride = (StageBytes + 127) / 128 * 128;
	static constexpr int Stages     = StageStride * 4 <= 100 * 1024 ? 4 : (StageStride * 3 <= 100 * 1024 ? 3 : 2);   // small tiles: more in flight
	static constexpr int SmemBytes  = Stages * StageStride + 128;
	static_assert (RowBytes % 16 == 0, "rows are read with 128-bit loads");
};

struct FbConv { uint32_t magic; float offset; };     // see the table above (unused for complex float)

__device__ __forceinline__ void tma_load_3d (uint32_t dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
	asm volatile ("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
	              :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32 (bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// sample `idx` (0 .. PerChunk - 1) of a 16-byte chunk as the float pair the handler would have produced
template <int FMT>
__device__ __forceinline__ float2 fb_sample (const uint4 &c, const int idx, const FbConv &cv) {
	if (FMT == kFmtCF32) {
	   const uint32_t a = idx == 0 ? c.x : c.z, b = idx == 0 ? c.y : c.w;
	   return make_float2 (__uint_as_float (a), __uint_as_float (b));
	}
	if (FMT == kFmtS16) {                               // word idx = (I, Q) as two int16
	   uint32_t w = idx == 0 ? c.x : idx == 1 ? c.y : idx == 2 ? c.z : c.w;
	   w ^= 0x80008000u;
	   const float i = __uint_as_float (__byte_perm (w, cv.magic, 0x7610)) - cv.offset;
	   const float q = __uint_as_float (__byte_perm (w, cv.magic, 0x7632)) - cv.offset;
	   return make_float2 (i, q);
	}
//	8-bit: word idx / 2 holds two samples (I, Q, I, Q)
	uint32_t w = (idx >> 1) == 0 ? c.x : (idx >> 1) == 1 ? c.y : (idx >> 1) == 2 ? c.z : c.w;
	if (FMT == kFmtS8) w ^= 0x80808080u;
	const uint32_t si = (idx & 1) ? 0x7652u : 0x7650u, sq = (idx & 1) ? 0x7653u : 0x7651u;
	const float i = __uint_as_float (__byte_perm (w, cv.magic, si)) - cv.offset;
	const float q = __uint_as_float (__byte_perm (w, cv.magic, sq)) - cv.offset;
	return make_float2 (i, q);
}

// tap i of the composite: c_comp for D = 12, else the K1g layout c_poly[p][g] = C[D g + D - 1 - p] with NG groups
template <int D, int NG>
__device__ __forceinline__ float fb_tap (const int i) {
	return D == kDecim ? c_comp [i] : c_poly [(D - 1 - i % D) * NG + i / D];
}
template <int D, int GPT, int NT>
__device__ __forceinline__ void fb_accum (const int j, float2 v, float2 (&acc) [GPT], float2 (&dcs) [GPT]) {
constexpr int NG = (NT + D - 1) / D;
#pragma unroll
	for (int k = 0; k < GPT; k ++) {
	   const int i = D * k + D - 1 - j;                      // tap index
	   if (i >= 0 && i < NT) acc [k] = ffma2 (fb_tap<D, NG> (i), v, acc [k]);
	}
	if (j >= 0) { dcs [j / D].x += v.x; dcs [j / D].y += v.y; }
}

// the samples of one box row from row-local position J0 on (own row: JOFF = 0; r rows back: JOFF = -r RowSamples)
template <int D, int GPT, int NT, int FMT, int J0, int JOFF>
__device__ __forceinline__ void fb_row (const unsigned char *row, const FbConv &cv, float2 (&acc) [GPT], float2 (&dcs) [GPT]) {
typedef Fb<D, GPT, NT, FMT> F;
#pragma unroll
	for (int q = J0 / F::PerChunk; q < F::Chunks; q ++) {
	   const uint4 c = *reinterpret_cast<const uint4 *>(row + 16 * q);
#pragma unroll
	   for (int i = 0; i < F::PerChunk; i ++) {
	      const int j = q * F::PerChunk + i;
	      if (j >= J0) fb_accum<D, GPT, NT> (j + JOFF, fb_sample<FMT> (c, i, cv), acc, dcs);
	   }
	}
}

// map  : 3-D tensor map over this call's samples as 32-bit words, dims (innermost first) [RowBytes / 4][rows][streams]
// hist : [n_streams][hist_len] float2, the converted samples preceding the call (the last NT - D are used)
template <int D, int GPT, int NT, int FMT>
__global__ void __launch_bounds__ (kFeThreads, 2)
frontend_tmab_kernel (const __grid_constant__ CUtensorMap map, const FbConv cv,
                      const float2 *__restrict__ hist, int hist_len,
                      float2 *__restrict__ U, float2 *__restrict__ S,
                      int64_t out_pitch, int32_t tiles_per_stream, int32_t n_streams) {
typedef Fb<D, GPT, NT, FMT> F;
constexpr int NG = (NT + D - 1) / D;
extern __shared__ unsigned char fb_smem_raw [];
__shared__ __align__ (8) uint64_t sFull [F::Stages];
const int tid = threadIdx.x;
unsigned char *ring = fb_smem_raw + ((128u - (smem_u32 (fb_smem_raw) & 127u)) & 127u);
const int total = tiles_per_stream * n_streams;
const int G = gridDim.x;
	static_assert (F::Prev == 1 || F::Prev == 2, "one or two rows of history");

	if (tid == 0) {
#pragma unroll
	   for (int s = 0; s < F::Stages; s ++) mbar_init (&sFull [s], 1);
	   asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads ();
	if (tid == 0) {
#pragma unroll
	   for (int s = 0; s < F::Stages; s ++) {
	      const int w = blockIdx.x + s * G;
	      if (w < total) {
	         const int stream = w / tiles_per_stream, tile = w - stream * tiles_per_stream;
	         mbar_expect_tx (&sFull [s], F::StageBytes);
	         tma_load_3d (smem_u32 (ring + s * F::StageStride), &map, &sFull [s], 0, tile * kFtRows - F::Prev, stream);
	      }
	   }
	}

int it = 0;
	for (int w = blockIdx.x; w < total; w += G, it ++) {
	   const int s = it % F::Stages;
	   const uint32_t parity = (uint32_t)(it / F::Stages) & 1u;
	   const int stream = w / tiles_per_stream, tile = w - stream * tiles_per_stream;
	   while (!mbar_try_wait (&sFull [s], parity)) { }
	   const unsigned char *stage = ring + s * F::StageStride;

	   float2 acc [GPT], dcs [GPT];
#pragma unroll
	   for (int k = 0; k < GPT; k ++) { acc [k] = make_float2 (0.f, 0.f); dcs [k] = make_float2 (0.f, 0.f); }
//	   box rows tid .. tid + Prev - 1 = the rows before this thread's outputs (their last NT - D samples matter), box
//	   row tid + Prev = its own.  The rows before the first tile of a call lie outside the tensor: TMA fills them
//	   with zero BYTES, which are not zero SAMPLES in the integer formats — the threads whose history reaches
//	   there skip those rows and take the carried (float) history instead (below).
	   if (F::Prev == 2) {
	      if (!(tile == 0 && tid < 2))
	         fb_row<D, GPT, NT, FMT, 2 * F::RowSamples - (NT - D), -2 * F::RowSamples> (stage + tid * F::RowBytes, cv, acc, dcs);
	      if (!(tile == 0 && tid < 1))
	         fb_row<D, GPT, NT, FMT, 0, -F::RowSamples> (stage + (tid + 1) * F::RowBytes, cv, acc, dcs);
	   }
	   else if (!(tile == 0 && tid == 0))
	      fb_row<D, GPT, NT, FMT, F::RowSamples - (NT - D), -F::RowSamples> (stage + tid * F::RowBytes, cv, acc, dcs);
	   fb_row<D, GPT, NT, FMT, 0, 0> (stage + (tid + F::Prev) * F::RowBytes, cv, acc, dcs);
	   asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");      // see frontend_tma.cuh
	   __syncthreads ();

	   if (tid == 0) {
	      const int wn = w + F::Stages * G;
	      if (wn < total) {
	         const int sn = wn / tiles_per_stream, tn = wn - sn * tiles_per_stream;
	         mbar_expect_tx (&sFull [s], F::StageBytes);
	         tma_load_3d (smem_u32 (ring + s * F::StageStride), &map, &sFull [s], 0, tn * kFtRows - F::Prev, sn);
	      }
	   }
	   if (tile == 0 && tid < F::Prev) {
//	      the samples of this thread's window that precede the call: carried history hs [n], n < 0 the index in the call
	      const float2 *hs = hist + (int64_t)stream * hist_len + hist_len;
	      for (int k = 0; k < GPT; k ++)
	         for (int i = 0; i < NT; i ++) {
	            const int n = tid * F::RowSamples + D * k + D - 1 - i;
	            if (n < 0) acc [k] = ffma2 (fb_tap<D, NG> (i), hs [n], acc [k]);
	         }
	   }
	   const int64_t m0 = ((int64_t)tile * kFeThreads + tid) * GPT;
	   float2 *up = U + (int64_t)stream * out_pitch + m0, *sp = S + (int64_t)stream * out_pitch + m0;
	   if (GPT == 4) {
	      float4 *u4 = reinterpret_cast<float4 *>(up), *s4 = reinterpret_cast<float4 *>(sp);
	      u4 [0] = make_float4 (acc [0].x, acc [0].y, acc [GPT > 1 ? 1 : 0].x, acc [GPT > 1 ? 1 : 0].y);
	      u4 [1] = make_float4 (acc [GPT > 2 ? 2 : 0].x, acc [GPT > 2 ? 2 : 0].y, acc [GPT > 3 ? 3 : 0].x, acc [GPT > 3 ? 3 : 0].y);
	      s4 [0] = make_float4 (dcs [0].x, dcs [0].y, dcs [GPT > 1 ? 1 : 0].x, dcs [GPT > 1 ? 1 : 0].y);
	      s4 [1] = make_float4 (dcs [GPT > 2 ? 2 : 0].x, dcs [GPT > 2 ? 2 : 0].y, dcs [GPT > 3 ? 3 : 0].x, dcs [GPT > 3 ? 3 : 0].y);
	   }
	   else if (GPT == 2) {
	      *reinterpret_cast<float4 *>(up) = make_float4 (acc [0].x, acc [0].y, acc [GPT > 1 ? 1 : 0].x, acc [GPT > 1 ? 1 : 0].y);
	      *reinterpret_cast<float4 *>(sp) = make_float4 (dcs [0].x, dcs [0].y, dcs [GPT > 1 ? 1 : 0].x, dcs [GPT > 1 ? 1 : 0].y);
	   }
	   else {
#pragma unroll
	      for (int k = 0; k < GPT; k ++) { up [k] = acc [k]; sp [k] = dcs [k]; }
	   }
	}
}

}	// namespace sdrjfm
