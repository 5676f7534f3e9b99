// K6b — fm-rate audio to working rate (192 kHz -> 48 kHz) and the start-up fade-in.
//
// The reference does this step with libsamplerate (newConverter, SRC_SINC_MEDIUM_QUALITY,
// src/various/newconverter.cpp:26-80, called at src/fm/fm-processor.cpp:633-634).  That
// library is neither vendored in the reference tree nor installed here, so its
// coefficient table cannot be restated: PARITY UNPINNED for this stage.
// What is built instead is a documented polyphase windowed-sinc decimator of our own:
//     y[q] = sum_{i<129} h[i] a[4 q + 3 - i],   h = Blackman-windowed sinc, fc = 20 kHz
// validated against a float64 model of the same taps.  The fade-in restates
// fm-processor.cpp:638-642 exactly.
#pragma once
#include "common.cuh"

namespace sdrjfm {

constexpr int kRsTaps = 129;
constexpr int kRsHist = 128;
constexpr int kRsDecim = 4;

__constant__ float c_rs_taps [kRsTaps + 3];

// a      : [S][pitch] fm-rate stereo of this call (M samples); hist: [S][128] previous ones
// out    : [S][out_pitch] working-rate stereo
// g0     : global fm index of a[.][0];  q0: global output index of out[.][0];  nq: outputs
__global__ void resample4_kernel (const float2 *__restrict__ a, int64_t pitch,
                                  const float2 *__restrict__ hist,
                                  float2 *__restrict__ out, int64_t out_pitch,
                                  int64_t g0, int64_t q0, int32_t nq,
                                  int32_t fade_cnt, int32_t fade_max) {
const int stream = blockIdx.y;
const int q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= nq) return;
const float2 *as = a + (int64_t)stream * pitch;
const float2 *hs = hist + (int64_t)stream * kRsHist;
const int64_t top = (q0 + q) * kRsDecim + (kRsDecim - 1) - g0;   // local index of newest sample
float2 acc = make_float2 (0.f, 0.f);
	for (int i = 0; i < kRsTaps; i ++) {
	   const int64_t j = top - i;
	   const float2 v = j >= 0 ? as [j] : hs [kRsHist + j];
	   acc.x = fmaf (c_rs_taps [i], v.x, acc.x);
	   acc.y = fmaf (c_rs_taps [i], v.y, acc.y);
	}
//	fade-in: pcmSample *= (max - cnt) / max while cnt > 0; cnt decrements per output sample
const int32_t cnt = fade_cnt - q;
	if (cnt > 0) {
	   const float f = fdiv (fsub ((float)fade_max, (float)cnt), (float)fade_max);
	   acc.x = fmul (acc.x, f); acc.y = fmul (acc.y, f);
	}
	out [(int64_t)stream * out_pitch + q] = acc;
}

// new_hist[i] = sample at local index M - 128 + i of (old_hist | a[0..M))
__global__ void roll_audio_history_kernel (const float2 *__restrict__ a, int64_t pitch,
                                           const float2 *__restrict__ old_hist,
                                           float2 *__restrict__ new_hist, int32_t M) {
const int stream = blockIdx.x;
const int i = threadIdx.x;
	if (i >= kRsHist) return;
const int64_t pos = (int64_t)M - kRsHist + i;
	new_hist [(int64_t)stream * kRsHist + i] =
	      pos >= 0 ? a [(int64_t)stream * pitch + pos]
	               : old_hist [(int64_t)stream * kRsHist + (kRsHist + pos)];
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
