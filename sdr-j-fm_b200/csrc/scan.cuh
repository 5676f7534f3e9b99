// Station scan (fmProcessor::run, src/fm/fm-processor.cpp:478-495, 886-904):
// while scanning, the fm-rate samples are NOT demodulated; every 1024 of them are Fourier transformed
// and the level around the carrier (bins 5..24 either side of 0) is compared with the level at the
// band edge (bins around 512):  get_db (signal, 256) - get_db (noise, 256) > thresHold  => station.
// One CTA per (1024-sample block, stream): radix-2 FFT in shared memory; the two dB figures go out,
// the threshold comparison (a constructor argument of the reference) is the caller's.
#pragma once
#include "common.cuh"

namespace sdrjfm {

constexpr int kScanN = 1024, kScanThreads = 256;

// blocks are cut from (carry | z): carry = the n_carry fm-rate samples left over by the previous call
__global__ void __launch_bounds__ (kScanThreads)
scan_kernel (const float2 *__restrict__ z, int64_t pitch, const float2 *__restrict__ carry, int32_t n_carry,
             float2 *__restrict__ out_db, int32_t out_pitch) {
__shared__ float2 a [kScanN];
__shared__ float sAbs [80];
const int tid = threadIdx.x, stream = blockIdx.y, blk = blockIdx.x;
const float2 *zs = z + (int64_t)stream * pitch;
const float2 *cs = carry + (int64_t)stream * kScanN;
//	bit-reversed load (decimation in time, like fft-complex.cpp:73-98)
	for (int i = tid; i < kScanN; i += kScanThreads) {
	   const int64_t g = (int64_t)blk * kScanN + i - n_carry;            // index into z (negative: carry)
	   const float2 v = g >= 0 ? zs [g] : cs [n_carry + g];
	   a [__brev ((unsigned)i) >> 22] = v;
	}
	__syncthreads ();
	for (int half = 1; half < kScanN; half <<= 1) {
#pragma unroll
	   for (int q = 0; q < kScanN / 2 / kScanThreads; q ++) {
	      const int b = tid + q * kScanThreads;
	      const int pos = b & (half - 1);
	      const int i0 = ((b - pos) << 1) + pos;
	      float sn, cs2;
	      sincospif (-(float)pos / (float)half, &sn, &cs2);               // exp (-2 pi i pos / (2 half))
	      const float2 u = a [i0], v = a [i0 + half];
	      const float2 t = make_float2 (v.x * cs2 - v.y * sn, v.x * sn + v.y * cs2);
	      a [i0] = make_float2 (u.x + t.x, u.y + t.y);
	      a [i0 + half] = make_float2 (u.x - t.x, u.y - t.y);
	   }
	   __syncthreads ();
	}
//	getSignal / getNoise (:886-904): 40 magnitudes each, summed in the reference's order
	if (tid < 80) {
	   const int k = tid % 20 + 5, which = tid / 20;                      // 0,1: signal; 2,3: noise
	   const int idx = which == 0 ? k : which == 1 ? kScanN - 1 - k : which == 2 ? kScanN / 2 - 1 - k : kScanN / 2 + 1 + k;
	   const float2 v = a [idx];
	   sAbs [tid] = (float)sqrt ((double)v.x * v.x + (double)v.y * v.y);
	}
	__syncthreads ();
	if (tid == 0) {
	   float sig = 0.f, noi = 0.f;
	   for (int i = 0; i < 40; i ++) sig = fadd (sig, sAbs [i]);
	   for (int i = 0; i < 40; i ++) noi = fadd (noi, sAbs [40 + i]);
	   sig = fdiv (sig, 40.f); noi = fdiv (noi, 40.f);
	   // get_db (x, 256) = 20 log10 ((x + 1) / 256), fm-constants.h:144-146
	   out_db [(int64_t)stream * out_pitch + blk] =
	         make_float2 (20.f * log10f (fdiv (fadd (sig, 1.f), 256.f)), 20.f * log10f (fdiv (fadd (noi, 1.f), 256.f)));
	}
}

// new carry = the last (n_carry + M) mod 1024 samples of (carry | z)
__global__ void scan_carry_kernel (const float2 *__restrict__ z, int64_t pitch, int32_t M,
                                   const float2 *__restrict__ carry, int32_t n_carry, float2 *__restrict__ carry_new) {
const int stream = blockIdx.x;
const int total = n_carry + M, keep = total % kScanN;
	for (int i = threadIdx.x; i < keep; i += blockDim.x) {
	   const int64_t g = (int64_t)(total - keep + i) - n_carry;
	   carry_new [(int64_t)stream * kScanN + i] = g >= 0 ? z [(int64_t)stream * pitch + g] : carry [(int64_t)stream * kScanN + n_carry + g];
	}
}

}	// namespace sdrjfm
