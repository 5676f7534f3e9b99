// K1t — the decimating front end of the reference's own rate and sample format (2.304 MS/s
// complex float, LO off, inputFilter off: the benchmarked configuration) with the IQ stream
// brought into shared memory by TMA.
//
// Same arithmetic as frontend_fir.cuh (ONE real 37-tap FIR decimating by 12 with the RF DC
// removal folded into the taps, plus the plain 12-sample block sums); what changes is how the
// samples travel:
//   * persistent CTAs (two per SM) walk the (stream, tile) work items; a tile is 128 rows of 48
//     samples (= 512 fm-rate outputs) plus the row before it (the 36-sample FIR history);
//   * one elected thread issues ONE cp.async.bulk.tensor per tile (49.5 KB, 4-D tensor map
//     [stream][row][3][32 floats], 128-byte swizzle) into a two-stage ring, completion on an
//     mbarrier: no load instructions, no staging registers, no store-to-shared instructions, and
//     the next tile is in flight while the current one is computed;
//   * the hardware swizzle XORs the 16-byte chunk index with the 128-byte line number, so a
//     thread reading ITS row (line 3 t + c) with 128-bit loads is conflict-free across the
//     quarter-warp although consecutive threads are 384 bytes apart; out-of-range rows (the
//     row before the first tile of a stream) are zero-filled by TMA and the carried history is
//     added by the one thread that needs it.
// Only whole tiles run here; the ragged last tile of a call goes through frontend_fir_kernel.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "frontend_fir.cuh"
#include "frontend_poly.cuh"

namespace sdrjfm {

constexpr int kFtRows        = kFeThreads;                    // rows (threads) per tile
constexpr int kFtRowSamples  = kDecim * kFeGpt;               // 48 samples = 384 bytes = 3 lines of 128 B
constexpr int kFtBoxRows     = kFtRows + 1;                   // + the row before the tile
constexpr int kFtStageBytes  = kFtBoxRows * kFtRowSamples * (int)sizeof (float2);     // 49536
constexpr int kFtStageStride = (kFtStageBytes + 1023) / 1024 * 1024;                  // swizzle-128B wants 1 KB alignment
constexpr int kFtStages      = 2;
constexpr int kFtSmemBytes   = kFtStages * kFtStageStride + 1024;                    // + alignment slack

__device__ __forceinline__ uint32_t smem_u32 (const void *p) { return (uint32_t)__cvta_generic_to_shared (p); }
__device__ __forceinline__ void mbar_init (uint64_t *b, uint32_t count) {
	asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32 (b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx (uint64_t *b, uint32_t bytes) {
	asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32 (b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait (uint64_t *b, uint32_t parity) {
uint32_t ok;
	asm volatile ("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
	              : "=r"(ok) : "r"(smem_u32 (b)), "r"(parity) : "memory");
	return ok != 0;
}
__device__ __forceinline__ void tma_load_4d (uint32_t dst, const CUtensorMap *map, uint64_t *bar,
                                             int c0, int c1, int c2, int c3) {
	asm volatile ("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
	              :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32 (bar)),
	                 "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// The kernel is instantiated for the shapes whose row (D * GPT samples) is 48 samples = 384 bytes:
// (D, GPT, NT) = (12, 4, 37) — 2.304 / 2.4 MS/s — and (48, 1, 73) — 10 MS/s.
// tap i of the composite: c_comp for D = 12, the K1g layout c_poly[p][g] (NG = 2) for D = 48
template <int D>
__device__ __forceinline__ float ft_tap (const int i) {
	return D == kDecim ? c_comp [i] : c_poly [(D - 1 - i % D) * 2 + i / D];
}

// one sample at row-local position j (own row: 0..47, previous row: -26..-1) into the GPT outputs;
// j is a compile-time constant at every call site (fully unrolled loops), so the tap test and the
// constant-bank index fold away
template <int D, int GPT, int NT>
__device__ __forceinline__ void ft_accum (const int j, float2 v, float2 (&acc) [GPT], float2 (&dcs) [GPT]) {
#pragma unroll
	for (int k = 0; k < GPT; k ++) {
	   const int i = D * k + D - 1 - j;                      // tap index
	   if (i >= 0 && i < NT) acc [k] = ffma2 (ft_tap<D> (i), v, acc [k]);
	}
	if (j >= 0) { dcs [j / D].x += v.x; dcs [j / D].y += v.y; }
}

// chunks LC0 .. LC0 + N - 1 of the row whose first 128-byte line is line0; chunk lc holds the
// row-local samples 2 lc and 2 lc + 1 (+ JOFF: -48 for the previous row).  Physical address of chunk
// q of line L under the 128-byte swizzle: L * 128 + ((q ^ (L & 7)) << 4).
template <int D, int GPT, int NT, int LC0, int N, int JOFF>
__device__ __forceinline__ void ft_row (const unsigned char *stage, int line0,
                                        float2 (&acc) [GPT], float2 (&dcs) [GPT]) {
#pragma unroll
	for (int n = 0; n < N; n ++) {
	   const int lc = LC0 + n, c = lc >> 3, q = lc & 7;
	   const int L = line0 + c;
	   const float4 v = *reinterpret_cast<const float4 *>(stage + L * 128 + ((q ^ (L & 7)) << 4));
	   ft_accum<D, GPT, NT> (2 * lc + JOFF, make_float2 (v.x, v.y), acc, dcs);
	   ft_accum<D, GPT, NT> (2 * lc + 1 + JOFF, make_float2 (v.z, v.w), acc, dcs);
	}
}

// map  : 4-D tensor map over this call's samples, dims (innermost first) [32 floats][3][rows][streams]
// hist : [n_streams][hist_len] the raw samples preceding x[.][0] (the last 36 are used)
// U, S : [n_streams][out_pitch]; tiles_per_stream whole tiles (128 GPT outputs each) are produced
template <int D, int GPT, int NT>
__global__ void __launch_bounds__ (kFeThreads, 2)
frontend_tma_kernel (const __grid_constant__ CUtensorMap map,
                     const float2 *__restrict__ hist, int hist_len,
                     float2 *__restrict__ U, float2 *__restrict__ S,
                     int64_t out_pitch, int32_t tiles_per_stream, int32_t n_streams) {
extern __shared__ unsigned char ft_smem_raw [];
__shared__ __align__ (8) uint64_t sFull [kFtStages];
const int tid = threadIdx.x;
unsigned char *ring = ft_smem_raw + ((1024u - (smem_u32 (ft_smem_raw) & 1023u)) & 1023u);
const int total = tiles_per_stream * n_streams;
const int G = gridDim.x;

	if (tid == 0) {
#pragma unroll
	   for (int s = 0; s < kFtStages; s ++) mbar_init (&sFull [s], 1);
	   asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads ();
	if (tid == 0) {
#pragma unroll
	   for (int s = 0; s < kFtStages; s ++) {
	      const int w = blockIdx.x + s * G;
	      if (w < total) {
	         const int stream = w / tiles_per_stream, tile = w - stream * tiles_per_stream;
	         mbar_expect_tx (&sFull [s], kFtStageBytes);
	         tma_load_4d (smem_u32 (ring + s * kFtStageStride), &map, &sFull [s], 0, 0, tile * kFtRows - 1, stream);
	      }
	   }
	}

int it = 0;
	for (int w = blockIdx.x; w < total; w += G, it ++) {
	   const int s = it % kFtStages;
	   const uint32_t parity = (uint32_t)(it / kFtStages) & 1u;
	   const int stream = w / tiles_per_stream, tile = w - stream * tiles_per_stream;
	   while (!mbar_try_wait (&sFull [s], parity)) { }
	   const unsigned char *stage = ring + s * kFtStageStride;

	   static_assert (D * GPT == kFtRowSamples && NT - D <= 26, "row of 48 samples, history within the previous row's last 26");
	   float2 acc [GPT], dcs [GPT];
#pragma unroll
	   for (int k = 0; k < GPT; k ++) { acc [k] = make_float2 (0.f, 0.f); dcs [k] = make_float2 (0.f, 0.f); }
//	   box row tid = the row before this thread's outputs (samples -25..-1 matter: chunks 11..23),
//	   box row tid + 1 = its own 48 samples
	   ft_row<D, GPT, NT, 11, 13, -kFtRowSamples> (stage, tid * 3, acc, dcs);
	   ft_row<D, GPT, NT, 0, 24, 0> (stage, (tid + 1) * 3, acc, dcs);
//	   The next TMA write into this stage is an async-proxy access; this thread's shared loads are
//	   generic-proxy accesses that may still be in flight when it reaches the barrier (their consumers
//	   can be scheduled behind it).  The proxy fence orders them before anything the async proxy does
//	   afterwards — without it a busy shared-memory pipe (another kernel on the SM) lets the refill
//	   overtake the tail of a thread's row.
	   asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
	   __syncthreads ();                                    // every thread has read the stage

	   if (tid == 0) {
	      const int wn = w + kFtStages * G;
	      if (wn < total) {
	         const int sn = wn / tiles_per_stream, tn = wn - sn * tiles_per_stream;
	         mbar_expect_tx (&sFull [s], kFtStageBytes);
	         tma_load_4d (smem_u32 (ring + s * kFtStageStride), &map, &sFull [s], 0, 0, tn * kFtRows - 1, sn);
	      }
	      if (tile == 0) {
//	         the row before the first tile of the call was zero-filled: add the carried history
	         const float2 *hs = hist + (int64_t)stream * hist_len + hist_len;       // hs[j], j = -36..-1
#pragma unroll
	         for (int k = 0; k < GPT; k ++)
	            for (int i = D * k + D; i < NT; i ++)
	               acc [k] = ffma2 (ft_tap<D> (i), hs [D * k + D - 1 - i], acc [k]);
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
	   else {
#pragma unroll
	      for (int k = 0; k < GPT; k ++) { up [k] = acc [k]; sp [k] = dcs [k]; }
	   }
	}
}

}	// namespace sdrjfm
