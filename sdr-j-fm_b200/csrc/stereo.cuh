// K4 — stereo decoder at the fm rate: 38 kHz carrier phase from the pilot phase, the
// "Perfect Stereo Separation" (PSS) phase trim, L-R demodulation, L/R matrix and selector.
//   fmProcessor::process_signal_with_rds          src/fm/fm-processor.cpp:689-730
//   PerfectStereoSeparation::process_sample       src/fm/stereo-separation.cpp:60-110
//   matrix + soundSelector                        src/fm/fm-processor.cpp:517-549
//
// The PSS loop low-passes cos/sin (phi38) * demod with fftFilter (2048, 295) and feeds the
// product of the two filtered paths back into phi38.  fftFilter is an overlap-add FFT
// convolution whose output lags its input by exactly NumofSamples = 2048 - 295 = 1753
// samples (src/various/fft-filters.cpp:132-163), i.e.
//     lp[n] = sum_{j<295} h[j] u[n - 1753 - j]            (h real, u the filter input)
// so inside any block of <= 1753 samples every lp value is already determined by earlier
// inputs: the feedback loop is sequential over BLOCKS only.  One CTA per stream walks the
// call in blocks of kStBlock samples; inside a block
//   1. the 295-tap FIR is evaluated directly from a 2048-entry ring of past inputs (shared memory),
//   2. the loop filter  acc += alpha*err  (clamped to +-pi/4), the mean-error one-pole and the
//      "error minimized" hysteresis run as scans; whenever the clamp or a hysteresis
//      transition could fall inside the block, one lane walks the block with the reference's
//      statement order instead (rare: start-up and once per 3 s at most),
//   3. phi38, the new filter inputs, L-R and the matrix are element-wise.
// The filter input only advances on samples that are actually stereo-decoded, as in the
// reference, so blocks are cut where the (locked, mode) category changes.
#pragma once
#include "common.cuh"
#include "sequential.cuh"

namespace sdrjfm {

constexpr int kStThreads = 256;
constexpr int kStPer     = 6;
constexpr int kStBlock   = kStThreads * kStPer;   // 1536 <= 1753
constexpr int kPssDelay  = 1753;                  // fftSize - filterDegree
constexpr int kPssTaps   = 295;
constexpr int kPssRing   = 2048;
constexpr int kStMaxTrans = 32;

__constant__ float2 c_pss_taps [kPssTaps + 1];    // LowPassFIR (295, 15000, fmRate) real taps, each as {c, c}

struct StereoParams {
	int32_t fm_mode, auto_mono, pss_on, sound_sel;
	float   panorama;
	float   pss_alpha;         // 10 / fmRate        fm-processor.cpp:81-82
	float   pss_lock_alpha;    // 1 / fmRate         stereo-separation.cpp:32
	int32_t rate3;             // 3 * fmRate         stereo-separation.cpp:91,97
	int32_t write_pss_tap;
};

struct PssState {              // PerfectStereoSeparation members + fmProcessor::pilotDelayPSS
	float   delay, acc, mean;
	int32_t minimized, lockCnt, unlockCnt;
};

// fmod (v, 2 pi) for a float v, through double as the reference evaluates it.  fmod is exact, and for
// 2 pi <= v < 6 pi the exact result is the plain double difference v - 2 pi n (Sterbenz: operands
// within a factor two), so the common cases cost one DADD.  The comparisons use the floats just above
// the doubles 2 pi and 4 pi (no float lies in between), which select the same branch as the reference.
__device__ __noinline__ float fmod_2pi_rare (float v) { return (float)fmod ((double)v, 2 * M_PI); }
__device__ __forceinline__ float fmod_2pi (float v) {
const float kTwoPiF = 6.2831855f, kFourPiF = 12.566371f, kSixPiF = 18.849556f;
	if (v >= 0.f && v < kTwoPiF) return v;
	if (v >= kTwoPiF && v < kFourPiF) return (float)((double)v - 2 * M_PI);
	if (v >= kFourPiF && v < kSixPiF) return (float)((double)v - 4 * M_PI);
	return fmod_2pi_rare (v);
}

// SinCos::getComplex / getCos index, sincos.cpp:81-91
__device__ __forceinline__ int32_t sincos_index (float phase) {
	while (phase < 0.f) phase = (float)((double)phase + 2 * M_PI);
	phase = fmod_2pi (phase);
int32_t i = (int32_t)((double)phase * (kFmRate / (2 * M_PI)));
	return i >= kFmRate ? i % kFmRate : i;
}

// phaseforLRDiff, fm-processor.cpp:707-714
__device__ __forceinline__ float phase_for_lr (float cur, float pssDelay) {
float ph = (float)(2 * ((double)cur + M_PI_4 + 0) - (double)pssDelay);
	if (ph < -6.283185f) ph = (float)((double)ph + 4 * M_PI);      // (double)ph < -2 pi: -6.283185f is the float just above -2 pi
	return fmod_2pi (ph);
}

// matrix and selector, fm-processor.cpp:517-549
__device__ __forceinline__ float2 lr_matrix (float sumLR, float diffLR, const StereoParams &P) {
const float dw = fmul (diffLR, P.fm_mode == 1 ? P.panorama : 1.0f);
const float left = fadd (sumLR, dw), right = fsub (sumLR, dw);
	switch (P.sound_sel) {
	   default:
	   case 0: return make_float2 (left, right);
	   case 1: return make_float2 (right, left);
	   case 2: return make_float2 (left, left);
	   case 3: return make_float2 (right, right);
	   case 4: return make_float2 (sumLR, sumLR);
	   case 5: case 6: return make_float2 (dw, dw);
	}
}

// one reference statement sequence of PerfectStereoSeparation::process_sample after the
// filter (stereo-separation.cpp:82-109), on the raw error re*im
__device__ __forceinline__ void pss_step (PssState &s, float err, const StereoParams &P) {
	if (!s.minimized) err = fmul (err, 10.0f);
	s.acc = fadd (s.acc, fmul (P.pss_alpha, err));
	s.mean = fadd (fmul (P.pss_lock_alpha, err), fmul (s.mean, fsub (1.0f, P.pss_lock_alpha)));
	if (fabsf (s.mean) < 0.001f) {
	   if (s.minimized || (++ s.lockCnt > P.rate3)) s.minimized = 1;
	   s.unlockCnt = 0;
	}
	else {
	   if (!s.minimized || (++ s.unlockCnt > P.rate3)) s.minimized = 0;
	   s.lockCnt = 0;
	}
	if ((double)s.acc < -M_PI_4) s.acc = (float)(-M_PI_4);
	else if ((double)s.acc > M_PI_4) s.acc = (float)M_PI_4;
}

// demod, phase, locked : K3 outputs.  sincos: the reference SinCos table (cos, sin) x fmRate.
// ring : [S][2048] past PSS filter inputs (state).  lr: [S][pitch] out.  pssd: tap (optional)
// diff_out : [S][pitch] diffLR of every sample (optional: the AF_DIFF scope stream, fm-processor.cpp:611-613)
__global__ void __launch_bounds__ (kStThreads)
stereo_kernel (const float *__restrict__ demod, const float *__restrict__ phase,
               const uint8_t *__restrict__ locked, int64_t pitch, int32_t M,
               const StereoParams P, const float2 *__restrict__ sincos,
               StreamState *__restrict__ state, float2 *__restrict__ ring_g,
               float2 *__restrict__ lr, float *__restrict__ pssd, float *__restrict__ diff_out) {
__shared__ float2 sRing [kPssRing];
__shared__ float  sErr [kStBlock];              // raw re*im per sample of the block
__shared__ float  sDel [kStBlock + 1];          // pilotDelayPSS BEFORE sample m (entry 0 = carry)
__shared__ double sSa [kStThreads / 32], sSb [kStThreads / 32];
__shared__ PssState sS;
__shared__ int sLen, sPos, sNTrans;
__shared__ int sTrans [kStMaxTrans];
const int tid = threadIdx.x;
const int stream = blockIdx.x;
StreamState &st = state [stream];
const float *dm = demod + (int64_t)stream * pitch;
const float *ph = phase + (int64_t)stream * pitch;
const uint8_t *lk = locked + (int64_t)stream * pitch;
float2 *out = lr + (int64_t)stream * pitch;
float *dtap = diff_out ? diff_out + (int64_t)stream * pitch : nullptr;
float *tapd = pssd ? pssd + (int64_t)stream * pitch : nullptr;
float2 *ring = ring_g + (int64_t)stream * kPssRing;

	for (int i = tid; i < kPssRing; i += kStThreads) sRing [i] = ring [i];
	if (tid == 0) {
	   sS.delay = st.pss_delay; sS.acc = st.pss_acc; sS.mean = st.pss_mean_error;
	   sS.minimized = st.pss_minimized; sS.lockCnt = st.pss_lock_cnt; sS.unlockCnt = st.pss_unlock_cnt;
	   sPos = st.pss_inp;                       // ring write position = inputs so far mod 2048
	}
	__syncthreads ();

//	positions where the lock flag changes (rare: typically one per station), found once for the
//	whole call so that the block loop never waits on global memory for them
	if (tid == 0) sNTrans = 0;
	__syncthreads ();
	for (int32_t m0 = 1 + tid * 8; m0 < M; m0 += kStThreads * 8) {       // 8 flags per step: the loads overlap
	   uint8_t f [9];
#pragma unroll
	   for (int k = 0; k < 9; k ++) f [k] = (m0 - 1 + k < M) ? lk [m0 - 1 + k] : 0;
#pragma unroll
	   for (int k = 1; k < 9; k ++)
	      if (m0 - 1 + k < M && (f [k] != 0) != (f [k - 1] != 0)) {
	         const int slot = atomicAdd (&sNTrans, 1);
	         if (slot < kStMaxTrans) sTrans [slot] = m0 - 1 + k;
	      }
	}
	__syncthreads ();
const int nTrans = sNTrans;                       // > kStMaxTrans: fall back to scanning per block
bool curLock = M > 0 ? lk [0] != 0 : false;

	for (int32_t p = 0; p < M; ) {
//	-- category of the block: 0 not stereo-decoded, 1 locked stereo, 2 unlocked stereo -------
	   const bool lock0 = curLock;
	   const int cat = (P.fm_mode != 2 && (lock0 || !P.auto_mono)) ? (lock0 ? 1 : 2) : 0;
	   const int want = min (kStBlock, M - p);
	   int len = want;
	   if (nTrans <= kStMaxTrans) {
	      for (int i = 0; i < nTrans; i ++) { const int t = sTrans [i]; if (t > p && t - p < len) len = t - p; }
	   }
	   else {
	      int firstDiff = want;
	      for (int m = tid; m < want; m += kStThreads)
	         if ((lk [p + m] != 0) != lock0) { firstDiff = m; break; }
	      if (tid == 0) sLen = want;
	      __syncthreads ();
	      if (firstDiff < want) atomicMin (&sLen, firstDiff);
	      __syncthreads ();
	      len = sLen;
	   }
	   if (p + len < M) curLock = lk [p + len] != 0;    // (consumed one block later: latency hidden)
	   const int m0 = tid * kStPer;
	   const int pos = sPos;
//	   this thread's demod / pilot-phase samples: requested now, used in step 3 (behind the FIR)
	   float dv [kStPer], pv [kStPer];
	   if (cat != 0) {
#pragma unroll
	      for (int k = 0; k < kStPer; k ++) {
	         const bool ok = m0 + k < len;
	         dv [k] = ok ? dm [p + m0 + k] : 0.f;
	         pv [k] = ok ? ph [p + m0 + k] : 0.f;
	      }
	   }

	   if (cat == 0) {
//	   mono / not locked with autoMono: audioOut = (demod, 0), :728-730
	      for (int m = tid; m < len; m += kStThreads) {
	         out [p + m] = lr_matrix (dm [p + m], 0.f, P);
	         if (dtap) dtap [p + m] = 0.f;
	         if (tapd && P.write_pss_tap) tapd [p + m] = lock0 ? sS.delay : 0.f;
	      }
	      __syncthreads ();
	      if (tid == 0 && !lock0) {               // pilotDelayPSS = 0; pPSS.reset (), :699-702
	         sS.delay = 0.f; sS.acc = 0.f; sS.mean = 0.f;
	         sS.minimized = 0; sS.lockCnt = 0; sS.unlockCnt = 0;
	      }
	      __syncthreads ();
	      p += len;
	      continue;
	   }

	   const bool pss = P.pss_on != 0;
	   if (pss) {
//	   1. lp[m] = sum_j h[j] u[n - 1753 - j]: ring index of u[n - 1753 - j] for n = block start + m
//	      is pos + m - 1753 - j  (pos = ring slot of sample `block start`)
	      float2 acc [kStPer];
#pragma unroll
	      for (int k = 0; k < kStPer; k ++) acc [k] = make_float2 (0.f, 0.f);
	      if (m0 < len) {
//	      element i of the window = u[block start + m0 + i - 1753]; at tap j output k reads element k - j.
//	      Elements live in register slot (i mod 6), so with the tap loop unrolled by 6 every slot is static.
	         const int base = pos + m0 - kPssDelay + 2 * kPssRing;      // ring slot of element 0
	         float2 v [kStPer];
#pragma unroll
	         for (int k = 0; k < kStPer; k ++) v [k] = sRing [(base + k) & (kPssRing - 1)];
	         int nxt = base - 1;                                          // ring slot of element -j-1
#pragma unroll
	         for (int j0 = 0; j0 + kStPer <= kPssTaps; j0 += kStPer) {
#pragma unroll
	            for (int u = 0; u < kStPer; u ++) {
	               const float2 c = c_pss_taps [j0 + u];
#pragma unroll
	               for (int k = 0; k < kStPer; k ++) {
	                  const float2 w = v [(k - u + kStPer) % kStPer];
	                  acc [k] = ffma2p (c, w, acc [k]);
	               }
	               v [(kStPer - 1 - u) % kStPer] = sRing [nxt & (kPssRing - 1)];
	               nxt --;
	            }
	         }
#pragma unroll
	         for (int u = 0; u < kPssTaps % kStPer; u ++) {               // remaining taps (295 = 49 * 6 + 1)
	            const float2 c = c_pss_taps [(kPssTaps / kStPer) * kStPer + u];
#pragma unroll
	            for (int k = 0; k < kStPer; k ++) {
	               const float2 w = v [(k - u + kStPer) % kStPer];
	               acc [k] = ffma2p (c, w, acc [k]);
	            }
	            v [(kStPer - 1 - u) % kStPer] = sRing [nxt & (kPssRing - 1)];
	            nxt --;
	         }
	      }
#pragma unroll
	      for (int k = 0; k < kStPer; k ++)
	         if (m0 + k < len) sErr [m0 + k] = fmul (acc [k].x, acc [k].y);       // :80
	   }
	   __syncthreads ();

	   if (pss && cat == 1) {
//	   2. loop filter.  Fast path valid iff the hysteresis cannot flip and the clamp is not hit.
	      const PssState s0 = sS;
	      const bool canFlip = s0.minimized ? (s0.unlockCnt + len > P.rate3) : (s0.lockCnt + len > P.rate3);
	      const float scale = s0.minimized ? 1.0f : 10.0f;
	      const double cM = (double)fsub (1.0f, P.pss_lock_alpha);
	      double bAcc = 0.0, bMean = 0.0, aMean = 1.0;
#pragma unroll
	      for (int k = 0; k < kStPer; k ++) {
	         if (m0 + k < len) {
	            const float e = fmul (sErr [m0 + k], scale);
	            bAcc += (double)fmul (P.pss_alpha, e);
	            bMean = bMean * cM + (double)fmul (P.pss_lock_alpha, e);
	            aMean *= cM;
	         }
	      }
	      double accv  = block_affine_start (1.0, bAcc, (double)s0.acc, sSa, sSb, nullptr);
	      double meanv = block_affine_start (aMean, bMean, (double)s0.mean, sSa, sSb, nullptr);
	      bool bad = false;
	      int lastNot = -1, lastIs = -1;          // last sample with |mean| >= 0.001 / < 0.001
	      float accEnd = 0.f, meanEnd = 0.f;
#pragma unroll
	      for (int k = 0; k < kStPer; k ++) {
	         if (m0 + k < len) {
	            const float e = fmul (sErr [m0 + k], scale);
	            sDel [m0 + k] = (m0 + k == 0) ? s0.delay : (float)accv;
	            accv += (double)fmul (P.pss_alpha, e);
	            meanv = meanv * cM + (double)fmul (P.pss_lock_alpha, e);
	            if (fabs (accv) >= M_PI_4) bad = true;
	            if (fabsf ((float)meanv) < 0.001f) lastIs = m0 + k; else lastNot = m0 + k;
	            if (m0 + k == len - 1) { accEnd = (float)accv; meanEnd = (float)meanv; }
	         }
	      }
	      const int slow = __syncthreads_or (bad || canFlip);
	      if (slow) {
	         if (tid == 0) {                      // reference statement order, one lane
	            PssState s = sS;
	            for (int m = 0; m < len; m ++) {
	               sDel [m] = s.delay;
	               pss_step (s, sErr [m], P);
	               s.delay = s.acc;
	            }
	            sS = s;
	         }
	      }
	      else {
//	      counters: trailing run lengths decide lockCnt / unlockCnt at the block end
	         int a = lastNot, b = lastIs;
#pragma unroll
	         for (int k = 16; k >= 1; k >>= 1) {
	            a = max (a, __shfl_xor_sync (0xffffffffu, a, k));
	            b = max (b, __shfl_xor_sync (0xffffffffu, b, k));
	         }
	         __shared__ int sLn [kStThreads / 32], sLi [kStThreads / 32];
	         if ((tid & 31) == 0) { sLn [tid >> 5] = a; sLi [tid >> 5] = b; }
	         __syncthreads ();
	         if (m0 <= len - 1 && len - 1 < m0 + kStPer) {      // owner of the last sample
	            int ln = -1, li = -1;
	            for (int q = 0; q < kStThreads / 32; q ++) { ln = max (ln, sLn [q]); li = max (li, sLi [q]); }
	            PssState s = s0;
	            s.acc = accEnd; s.mean = meanEnd; s.delay = accEnd;
	            if (!s0.minimized) {
	               // lockCnt counts the current run of |mean| < 0.001; unlockCnt is zeroed by any such sample
	               s.lockCnt = (ln < 0) ? s0.lockCnt + len : (len - 1 - ln);
	               if (li >= 0) s.unlockCnt = 0;
	            }
	            else {
	               s.unlockCnt = (li < 0) ? s0.unlockCnt + len : (len - 1 - li);
	               if (ln >= 0) s.lockCnt = 0;
	            }
	            sS = s;
	         }
	      }
	      __syncthreads ();
	   }
	   else if (pss && cat == 2) {
//	   unlocked but stereo-decoded (autoMono off): the loop is reset before every sample (:699-702)
	      for (int m = tid; m < len; m += kStThreads) sDel [m] = 0.f;
	      __syncthreads ();
	      if (tid == 0) {
	         PssState s;
	         s.delay = 0.f; s.acc = 0.f; s.mean = 0.f; s.minimized = 0; s.lockCnt = 0; s.unlockCnt = 0;
	         pss_step (s, sErr [len - 1], P);
	         s.delay = s.acc;
	         sS = s;
	      }
	      __syncthreads ();
	   }
	   else {
	      for (int m = tid; m < len; m += kStThreads) sDel [m] = 0.f;      // pssActive false: delay = 0
	      __syncthreads ();
	      if (tid == 0) {
	         if (cat == 2) { sS.acc = 0.f; sS.mean = 0.f; sS.minimized = 0; sS.lockCnt = 0; sS.unlockCnt = 0; }
	         sS.delay = 0.f;
	      }
	      __syncthreads ();
	   }

//	   3. phi38, new filter inputs, L-R, matrix.  All table look-ups of the thread are issued before
//	      the first one is used.
	   float phv [kStPer]; float2 csv [kStPer];
#pragma unroll
	   for (int k = 0; k < kStPer; k ++) {
	      phv [k] = 0.f; csv [k] = make_float2 (0.f, 0.f);
	      if (m0 + k < len) {
	         phv [k] = phase_for_lr (pv [k], sDel [m0 + k]);
	         csv [k] = sincos [sincos_index (phv [k])];
	      }
	   }
#pragma unroll
	   for (int k = 0; k < kStPer; k ++) {
	      const int m = m0 + k;
	      if (m < len) {
	         const float d = dv [k];
	         const float phLR = phv [k];
	         const float2 cs = csv [k];
	         if (pss) sRing [(pos + m) & (kPssRing - 1)] = make_float2 (fmul (cs.x, d), fmul (cs.y, d));   // :66-67
	         float osc = cs.x;
	         if (P.sound_sel == 6) {               // S_LEFTminusRIGHT_Test: getSin, :721-723
	            osc = phLR < 0.f ? -sincos [((int32_t)((double)(-phLR) * (kFmRate / (2 * M_PI)))) % kFmRate].y
	                             :  sincos [((int32_t)((double)phLR * (kFmRate / (2 * M_PI)))) % kFmRate].y;
	         }
	         // (float)(2.0 * osc * d) evaluated in double: the double product of two floats is exact and
	         // doubling is exact, so this is the single float rounding of osc * d, doubled
	         const float diff = fmul (2.0f, fmul (osc, d));
	         out [p + m] = lr_matrix (d, diff, P);
	         if (dtap) dtap [p + m] = diff;
	         if (tapd && P.write_pss_tap) {
	            // pilotDelayPSS after the sample = the value entering the next one
	            float after;
	            if (!pss) after = 0.f;
	            else if (cat == 2) {
	               PssState s; s.delay = 0.f; s.acc = 0.f; s.mean = 0.f; s.minimized = 0; s.lockCnt = 0; s.unlockCnt = 0;
	               pss_step (s, sErr [m], P); after = s.acc;
	            }
	            else after = 0.f;                 // cat 1: filled below from sDel
	            tapd [p + m] = after;
	         }
	      }
	   }
	   __syncthreads ();
	   if (tapd && P.write_pss_tap && pss && cat == 1) {
	      for (int m = tid; m < len; m += kStThreads)
	         tapd [p + m] = (m + 1 < len) ? sDel [m + 1] : sS.delay;
	   }
	   if (tid == 0 && pss) sPos = (pos + len) & (kPssRing - 1);
	   __syncthreads ();
	   p += len;
	}

	for (int i = tid; i < kPssRing; i += kStThreads) ring [i] = sRing [i];
	if (tid == 0) {
	   st.pss_delay = sS.delay; st.pss_acc = sS.acc; st.pss_mean_error = sS.mean;
	   st.pss_minimized = sS.minimized; st.pss_lock_cnt = sS.lockCnt; st.pss_unlock_cnt = sS.unlockCnt;
	   st.pss_inp = sPos;
	}
}

}	// namespace sdrjfm
