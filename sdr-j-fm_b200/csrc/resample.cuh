// K1r — second stage of the rational polyphase resampler (front_end_mode 1).
//
// BASELINE config 4 asks for 2.4 / 6 / 10 MS/s device streams "polyphase-resampled to 192 kHz".
// The reference has no such block (it decimates by an integer and then treats 200 / 200 /
// 208.33 kHz as 192 kHz); this is the new block, defined here:
//
//   stage A  (frontend_poly.cuh, D = 5): 49-tap Blackman-windowed sinc low-pass, cut-off 0.1 fs
//            (= half the stage-A output rate, for every fs), decimate by 5:
//                A[k] = sum_i hA[i] x[5 k + 4 - i]                          fa = fs / 5
//   stage B  (this file): rational L / MB resampler fa -> 192 kHz, L / MB = 5 * 192000 / fs
//            reduced (2/5, 4/25, 12/125 for 2.4, 6, 10 MS/s), prototype = Blackman-windowed sinc,
//            cut-off 96 kHz, P = ceil (8 MB / L) taps per phase, every phase normalised to unit DC gain:
//                y[m] = sum_{j<P} hB[phi_m][j] A[n_m - j],   n_m = floor (m MB / L),  phi_m = (m MB) mod L
//            Index contract: fm sample m is emitted when stage-A sample n_m exists, i.e. when input
//            sample 5 n_m + 4 has arrived; ceil (nA L / MB) fm samples exist after nA stage-A samples.
//
// Stage A is the HBM-bound part (8 B read + 1.6 B written per input sample); stage B runs at
// the fm rate from L2-resident data.  The RF DC remover is handled as in the reference mode:
// plain block sums S travel with the samples (stage B adds up the stage-A sums between
// consecutive outputs) and the one-pole runs at the fm rate (discriminator.cuh).  Both filters are
// symmetric, so filtering the slowly moving DC estimate r equals r at the CENTRE of the window to
// second order: the block sums handed on are those dA stage-A samples back (dA = group delay), which
// by linearity of the one-pole delays the estimate by exactly that much.
#pragma once
#include "common.cuh"

namespace sdrjfm {

constexpr int kRsStageADecim = 5;
constexpr int kRsStageATaps  = 49;
constexpr int kRsMaxL        = 16;
constexpr int kRsMaxP        = 128;
constexpr int kRsHistPad     = 16;        // stage-A samples kept beyond P - 1 (block sums of the previous output)

struct ResampleParams {
	int32_t L, MB, P, HB;      // HB = history length (P - 1 + kRsHistPad)
	int64_t a0;                // absolute index of A[.][0]
	int64_t m0;                // absolute index of the first fm sample of this call
	int32_t M;                 // fm samples to produce
	int32_t dA;                // group delay of the cascade in stage-A samples (block sums are taken there)
};

// A, SA   : [n_streams][a_pitch] stage-A samples / 5-sample block sums of this call
// histA/S : [n_streams][HB] the stage-A samples / sums preceding A[.][0]
// taps    : [L][P]
__global__ void __launch_bounds__ (256)
resample_b_kernel (const float2 *__restrict__ A, const float2 *__restrict__ SA, int64_t a_pitch,
                   const float2 *__restrict__ histA, const float2 *__restrict__ histS,
                   const float *__restrict__ taps, const ResampleParams q,
                   float2 *__restrict__ U, float2 *__restrict__ S, int64_t out_pitch) {
const int stream = blockIdx.y;
const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= q.M) return;
const int64_t m = q.m0 + i;
const int64_t t = m * q.MB;
const int64_t n = t / q.L;
const int phi = (int)(t - n * q.L);
const int64_t nprev = m > 0 ? ((m - 1) * q.MB) / q.L : -1;
const int rel = (int)(n - q.a0);                     // position of n_m in this call's A (may be < 0)
const float2 *As = A + (int64_t)stream * a_pitch;
const float2 *Ss = SA + (int64_t)stream * a_pitch;
const float2 *hA = histA + (int64_t)stream * q.HB + q.HB;
const float2 *hS = histS + (int64_t)stream * q.HB + q.HB;
const float *h = taps + phi * q.P;
float2 acc = make_float2 (0.f, 0.f);
	for (int j = 0; j < q.P; j ++) {
	   const int r = rel - j;
	   const float2 v = r >= 0 ? As [r] : hA [r];
	   acc = ffma2 (h [j], v, acc);
	}
float2 s = make_float2 (0.f, 0.f);
	for (int64_t k = nprev + 1; k <= n; k ++) {
	   const int r = (int)(k - q.dA - q.a0);
	   const float2 v = r >= 0 ? Ss [r] : hS [r];
	   s.x += v.x; s.y += v.y;
	}
	U [(int64_t)stream * out_pitch + i] = acc;
	S [(int64_t)stream * out_pitch + i] = s;
}

}	// namespace sdrjfm
