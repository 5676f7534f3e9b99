// K3 — fm-rate stage after the discriminator, PARALLEL IN TIME and still bit-exact:
//   fm_afc one-pole + output scaling            fm-demodulator.cpp:197-198   (linear scan)
//   am_carr_ampl one-pole                       fm-demodulator.cpp:130-131   (linear scan)
//   19 kHz pilot PLL                            pilot-recover.cpp:54-62      (exact, see below)
//   pilot lock detector and its 0.5 s hysteresis pilot-recover.cpp:63-80     (linear scan + run length)
//
// The pilot PLL is a nonlinear float32 recurrence phi[n+1] = F_n (phi[n]); re-starting it
// from an approximate state never re-converges to the reference trajectory (SURVEY.md
// Appendix C), so it cannot be chunked.  It CAN be solved as a fixed-point problem:
//
//   * Every ~10.1 samples the phase passes through [4, 2 pi), where float32 has a uniform
//     grid of 2^-21.  Pick one such sample per cycle as an ANCHOR.  The map from one anchor
//     to the next (a SEGMENT of 10-11 exact float32 steps) is, on that grid, a pure
//     translation except when a sine-table index or a binade crossing changes — so an error
//     of k grid units at an anchor stays k units at the next one instead of being scrambled
//     by rounding, which is what makes the iteration below contract all the way to zero.
//   * Iterate over a window of kPiWin samples: every segment is advanced in parallel (one
//     lane per segment, exact reference arithmetic) from its current anchor value P[c] to
//     G[c]; if G[c] == P[c+1] bit for bit for every c the window IS the sequential
//     trajectory (each link is an exact reference step and the first anchor is the exact
//     carried state).  Otherwise the anchor values are corrected with a Newton step — the
//     affine recurrence d[c+1] = (G[c] - P[c+1]) + a[c] d[c], a[c] = product of
//     1 + gain x cos (phi) over the segment, solved by a scan — and the loop repeats.
//     Measured on the oracle: 3-5 iterations per 2048-sample window, result bit-identical
//     to pilotRecovery::getPilotPhase run sample by sample.
//   * Termination is unconditional: the exact prefix grows by at least one segment per
//     iteration, and after kPiMaxIter iterations one lane simply walks the window.
//
// One CTA per IQ stream; windows are taken in order; the quarter-wave sine table lives in
// shared memory (sequential.cuh).
#pragma once
#include "common.cuh"
#include "sequential.cuh"

namespace sdrjfm {

// 512 threads x 8 samples: a 4096-sample window has about 410 segments, so most lanes walk one
// segment per iteration; at 64 registers two CTAs share an SM (32 warps: the walk is a dependent
// chain of ~9-cycle instructions, only thread-level parallelism hides it).
constexpr int kPiThreads = 512;                      // the throughput shape: 4096-sample windows, two CTAs per SM
constexpr int kPiThreadsWide = 1024;                 // 8192-sample windows (SDRJFM_PILOT_WIDE=1): a pass covers twice the samples,
                                                     // but needs about one pass more per window (up to 9 without a pilot) and
                                                     // its 32-warp scans are slower: measured no faster for 32-128 streams, not used
constexpr int kPiPer     = 8;                        // consecutive samples per thread
constexpr int kPiWin     = kPiThreads * kPiPer;      // 4096 fm samples per window (time slices are multiples of it)
constexpr int kPiMaxIter = 24;

template <int THREADS>
struct PilotSmemT {
	static constexpr int Win = THREADS * kPiPer, MaxSeg = Win * 7 / 32, Warps = THREADS / 32;    // ~0.1 anchors per sample in practice
	float   x [Win];                // 5 * demod
	float   est [Win];              // phase BEFORE the step of sample n (the unknowns)
	double  delta [MaxSeg + 1];     // Newton correction per anchor
	double  resid [MaxSeg + 1];
	float   G [MaxSeg + 1];         // phase after the last step of segment c
	float   der [MaxSeg + 1];
	int16_t anc [MaxSeg + 2];       // first sample of segment c
	double  warpA [Warps], warpB [Warps];
	int     warpI [Warps];
	int     nseg, overflow;
	double  carry [4];
};
typedef PilotSmemT<kPiThreads> PilotSmem;
constexpr size_t kPiLutBytes  = ((size_t)(kFmRate / 4 + 1) * sizeof (float) + 15) / 16 * 16;
constexpr size_t kPiSmemBytes = kPiLutBytes + sizeof (PilotSmem);

// ---- constant-coefficient linear recurrence y[n] = c y[n-1] + b[n] over a window -----------
// Thread t owns kPiPer consecutive samples.  `B` is the value its chunk reaches from a zero
// state; the function returns the state just BEFORE the chunk, given the state `carry`
// before the window.  pw[k] = c^(kPiPer 2^k), k = 0..5  (pw[5] = one warp).
struct LinPow { double pw [6]; };

__device__ __forceinline__ LinPow lin_pow (double c) {
LinPow L;
double p = c;
#pragma unroll
	for (int i = 1; i < kPiPer; i <<= 1) p *= p;       // c^kPiPer (kPiPer is a power of two)
	L.pw [0] = p;
#pragma unroll
	for (int k = 1; k < 6; k ++) L.pw [k] = L.pw [k - 1] * L.pw [k - 1];
	return L;
}

__device__ __forceinline__ double lin_scan_start (double B, double carry, const LinPow &L,
                                                   double *sWarp) {
const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
double inc = B;
#pragma unroll
	for (int k = 0; k < 5; k ++) {
	   const double y = __shfl_up_sync (0xffffffffu, inc, 1 << k);
	   if (lane >= (1 << k)) inc += y * L.pw [k];
	}
	if (lane == 31) sWarp [warp] = inc;
	__syncthreads ();
double w = carry;
	for (int q = 0; q < warp; q ++) w = w * L.pw [5] + sWarp [q];
double dl = 1.0;
#pragma unroll
	for (int k = 0; k < 5; k ++) if (lane & (1 << k)) dl *= L.pw [k];
double excl = __shfl_up_sync (0xffffffffu, inc, 1);
	if (lane == 0) excl = 0.0;
	__syncthreads ();                                   // sWarp may be reused by the caller
	return w * dl + excl;
}

struct PilotParams {
	float   K_FM, omega, gain;
	int32_t lock_half_rate;
	int32_t n_streams;
};

// one exact step of pilotRecovery::getPilotPhase (phase part), pilot-recover.cpp:55-62
__device__ __forceinline__ float pilot_step (const SinLut &L, const float *q, float phi, float pilot,
                                             float gain, float omega, float &osc, float &cur) {
	osc = lut_getSin (L, q, phi);
const float p1 = fadd (phi, fmul (fmul (pilot, osc), gain));
	cur = pi_constrain (p1);
	return pi_constrain (fadd (p1, omega));
}

// The same step for the segment walk: phi is known to lie in [0, 2 pi] (it is a pi_constrain or
// wrap_2pi result), only the next phase is wanted, and the derivative is accumulated to first
// order (sum of x cos phi).  Bit-identical to pilot_step for such phi.
template <bool DER>
__device__ __forceinline__ float pilot_walk_step (const SinLut &L, const float *q, float phi, float pilot,
                                                  float gain, float omega, float &dsum) {
constexpr int32_t Q = kFmRate / 4, H = 2 * Q;
const float kTwoPiF = 6.2831855f;
int32_t i = (int32_t)((double)phi * (kFmRate / (2 * M_PI)));     // SinCos::fromPhasetoIndex, Phase >= 0
	if (i >= kFmRate) i -= kFmRate;
const bool neg = i >= H;
int32_t k = neg ? i - H : i;
	k = k > Q ? H - k : k;
float v = q [k];
	v = neg ? -v : v;
	if (i == H) v = L.sin_at_half;               // the zero crossing at pi: the table holds sin (pi) as computed, not -0
	if (DER) dsum = fmaf (pilot, __cosf (phi), dsum);
float p2 = fadd (fadd (phi, fmul (fmul (pilot, v), gain)), omega);
	if (p2 >= kTwoPiF) p2 = (float)((double)p2 - 2 * M_PI);     // PI_Constrain: fmod (v, 2 pi), v < 4 pi
	else if (p2 < 0.f) p2 = pi_constrain (p2);                  // cannot happen for |pilot| < 1000
	return p2;
}

__device__ __forceinline__ float wrap_2pi (double v) {          // any double -> float in [0, 2 pi]
	v -= 2 * M_PI * floor (v * (1.0 / (2 * M_PI)));
float f = (float)v;
	return f < 0.f ? 0.f : f;
}

template <bool LUT_SMEM, int THREADS = kPiThreads>
__global__ void __launch_bounds__ (THREADS, (LUT_SMEM || THREADS > 512) ? 1 : 2)
pilot_kernel (const float *__restrict__ res_raw, const float *__restrict__ zabs,
              int64_t pitch, int32_t M, const PilotParams P, const SinLut L,
              StreamState *__restrict__ state,
              float *__restrict__ demod_out, float *__restrict__ phase_out,
              uint8_t *__restrict__ locked_out, int32_t *__restrict__ iter_stats, int accumulate) {
extern __shared__ __align__ (16) unsigned char smem_raw [];
// LUT_SMEM: the quarter-wave sine table staged in shared memory (192 KB: only fits with windows of <= 2048 samples, a
// shape no launch uses any more).  Otherwise it is read through L1, which leaves room for two CTAs per SM; folding the
// table to a quarter of its footprint moved the step by < 1 % (profiles/r2_summary.md), so the look-up is not the limiter.
const float *sq = LUT_SMEM ? reinterpret_cast<const float *>(smem_raw) : L.q;
constexpr int kPiThreads = THREADS, kPiWin = THREADS * kPiPer, kPiMaxSeg = PilotSmemT<THREADS>::MaxSeg, kPiWarps = THREADS / 32;
PilotSmemT<THREADS> &S = *reinterpret_cast<PilotSmemT<THREADS> *>(smem_raw + (LUT_SMEM ? kPiLutBytes : 0));
const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
const int stream = blockIdx.x;
	if (LUT_SMEM) { float *w = reinterpret_cast<float *>(smem_raw); for (int i = tid; i <= kFmRate / 4; i += kPiThreads) w [i] = L.q [i]; }
StreamState &st = state [stream];
const float *rr = res_raw + (int64_t)stream * pitch;
const float *za = zabs + (int64_t)stream * pitch;
float *dm = demod_out + (int64_t)stream * pitch;
float *ph = phase_out + (int64_t)stream * pitch;
uint8_t *lk = locked_out + (int64_t)stream * pitch;

const float carrierAlpha = 0.0010f, fmDcAlpha = 0.0001f;          // fm-demodulator.cpp:115-117
const float oneMinusCarrier = fsub (1.0f, carrierAlpha);
const float oneMinusDc = fsub (1.0f, fmDcAlpha);
const float lockAlpha = 1.0f / 3000.0f;                           // pilot-recover.cpp:57
const double oneMinusLock = 1.0 - (double)lockAlpha;
const LinPow powDc = lin_pow ((double)oneMinusDc);
const LinPow powCa = lin_pow ((double)oneMinusCarrier);
const LinPow powLk = lin_pow (oneMinusLock);

	if (tid == 0) {
	   S.carry [0] = (double)st.fm_afc;
	   S.carry [1] = (double)st.am_carr_ampl;
	   S.carry [2] = (double)st.pilot_lock;
	}
float  phi0    = st.pilot_phase;              // exact carried PLL state
float  oscPrev = st.pilot_old;
int    runCarry = st.pilot_locked ? (1 << 29) : st.pilot_stable_cnt;
double winc    = (double)P.omega;
int    itTotal = 0, itMax = 0, nFallback = 0;
	__syncthreads ();

	for (int32_t base = 0; base < M; base += kPiWin) {
	   const int Tw = min (kPiWin, M - base);
	   const int n0 = tid * kPiPer;
//	-- 1. AFC one-pole, output scaling; carrier level --------------------------------------
	   float r [kPiPer], zb [kPiPer];
#pragma unroll
	   for (int j = 0; j < kPiPer; j ++) {
	      const bool ok = n0 + j < Tw;
	      r [j]  = ok ? rr [base + n0 + j] : 0.f;
	      zb [j] = ok ? za [base + n0 + j] : 0.f;
	   }
	   double bDc = 0.0, bCa = 0.0;
#pragma unroll
	   for (int j = 0; j < kPiPer; j ++) {
	      if (n0 + j < Tw) {
	         bDc = bDc * (double)oneMinusDc + (double)fmul (fmDcAlpha, r [j]);
	         bCa = bCa * (double)oneMinusCarrier + (double)fmul (carrierAlpha, zb [j]);
	      }
	      else {       // identity steps on the padding keep the constant-power scan valid
	         bDc = bDc * (double)oneMinusDc;
	         bCa = bCa * (double)oneMinusCarrier;
	      }
	   }
	   double afc = lin_scan_start (bDc, S.carry [0], powDc, S.warpA);
	   double am  = lin_scan_start (bCa, S.carry [1], powCa, S.warpB);
//	   (the carrier level is only read at the end of a call: its per-sample replay is left to the owner of the
//	   window's last sample)
	   const bool ownsLast = n0 <= Tw - 1 && Tw - 1 < n0 + kPiPer;
	   if (ownsLast) {
#pragma unroll
	      for (int j = 0; j < kPiPer; j ++)
	         if (n0 + j < Tw) am = am * (double)oneMinusCarrier + (double)fmul (carrierAlpha, zb [j]);
	   }
#pragma unroll
	   for (int j = 0; j < kPiPer; j ++) {
	      if (n0 + j < Tw) {
	         afc = afc * (double)oneMinusDc + (double)fmul (fmDcAlpha, r [j]);
	         const float demod = fdiv (fmul (fmul (20.0f, fsub (r [j], (float)afc)), 1.0f), P.K_FM);
	         dm [base + n0 + j] = demod;
	         S.x [n0 + j] = fmul (5.0f, demod);
	         if (n0 + j == Tw - 1) { S.carry [0] = afc; S.carry [1] = am; }
	      }
	   }
//	-- 2. pilot PLL: initial guess = carried phase advancing by the last mean increment ----
#pragma unroll
	   for (int j = 0; j < kPiPer; j ++)
	      if (n0 + j < Tw) S.est [n0 + j] = wrap_2pi ((double)phi0 + winc * (n0 + j));
	   if (tid == 0) S.est [0] = phi0;
	   __syncthreads ();

	   int it = 0;
	   bool converged = false;
	   for (; it < kPiMaxIter; it ++) {
//	   anchors: sample 0, and the first sample of every run of est inside [4.6, 5.3).  They only
//	   depend on the estimate to ~0.05 rad (the first guess is good to ~1e-3 in a locked window), so they are
//	   only refreshed when a window has not converged after 6, 12 and 18 passes.
	      const bool refresh = it == 0 || it == 6 || it == 12 || it == 18;
	      if (refresh) {
	         unsigned flags = 0;
	         {
	            bool prevIn = false;
	            if (n0 > 0 && n0 - 1 < Tw) { const float e = S.est [n0 - 1]; prevIn = e >= 4.6f && e < 5.3f; }
#pragma unroll
	            for (int j = 0; j < kPiPer; j ++) {
	               bool in = false;
	               if (n0 + j < Tw) { const float e = S.est [n0 + j]; in = e >= 4.6f && e < 5.3f; }
	               if ((in && !prevIn && n0 + j > 0) || (n0 + j == 0)) flags |= 1u << j;
	               prevIn = in;
	            }
	         }
	         int cnt = __popc (flags);
	         int inc = cnt;
#pragma unroll
	         for (int k = 1; k < 32; k <<= 1) {
	            const int y = __shfl_up_sync (0xffffffffu, inc, k);
	            if (lane >= k) inc += y;
	         }
	         if (lane == 31) S.warpI [warp] = inc;
	         __syncthreads ();
	         int off = inc - cnt;
	         for (int q = 0; q < warp; q ++) off += S.warpI [q];
	         if (tid == kPiThreads - 1) {
	            S.nseg = off + cnt;
	            S.overflow = (off + cnt > kPiMaxSeg);
	         }
#pragma unroll
	         for (int j = 0; j < kPiPer; j ++)
	            if ((flags >> j) & 1u) { if (off < kPiMaxSeg) S.anc [off] = (int16_t)(n0 + j); off ++; }
	         __syncthreads ();
	         if (S.overflow) break;
	      }
	      const int nseg = S.nseg;
//	   advance every segment with the exact reference arithmetic; a segment is consistent when it
//	   ends bit-exactly on the value the next segment starts from
//	   (the derivative of a segment moves by ~1e-6 of itself once the estimate is within a few ulp: it is taken in
//	   the first two passes and after an anchor refresh, and kept afterwards)
	      int bad = 0;
	      const bool wantDer = it < 2 || refresh;
	      for (int c = tid; c < nseg; c += kPiThreads) {
	         const int a0 = S.anc [c];
	         const int a1 = (c + 1 < nseg) ? (int)S.anc [c + 1] : Tw;
	         float p = S.est [a0];
	         float dsum = 0.0f;
	         if (wantDer) {
	            for (int n = a0; n < a1; n ++) {
	               if (n > a0) S.est [n] = p;
	               p = pilot_walk_step<true> (L, sq, p, S.x [n], P.gain, P.omega, dsum);
	            }
	            S.der [c] = fmaf (P.gain, dsum, 1.0f);   // d(end)/d(start) to first order
	         }
	         else {
	            for (int n = a0; n < a1; n ++) {
	               if (n > a0) S.est [n] = p;
	               p = pilot_walk_step<false> (L, sq, p, S.x [n], P.gain, P.omega, dsum);
	            }
	         }
	         S.G [c] = p;
	         double rs = 0.0;
	         if (c + 1 < nseg) {
	            const float nextP = S.est [a1];          // only its owner's walk reads it; nobody writes it here
	            if (nextP != p) { bad = 1; rs = (double)p - (double)nextP; }
	         }
	         S.resid [c] = rs;
	      }
	      const int nbad = __syncthreads_count (bad);
#ifdef SDRJFM_PILOT_TRACE      // per-pass convergence of a few windows of stream 0 (profiles/r1_summary.md)
	      if (blockIdx.x == 0 && tid == 0 && base >= kPiWin * 30 && base < kPiWin * 34) {
	         double mr = 0; int first = -1;
	         for (int c = 0; c < nseg; c ++) { if (S.resid [c] != 0.0 && first < 0) first = c; mr = fmax (mr, fabs (S.resid [c])); }
	         printf ("win %d it %d nseg %d nbad %d first %d maxresid %.3e\n", base / kPiWin, it, nseg, nbad, first, mr);
	      }
#endif
	      if (nbad == 0) { converged = true; break; }
//	   Newton correction of the anchors: delta[c+1] = resid[c] + der[c] delta[c], delta[0] = 0,
//	   as a block-wide scan of affine maps (thread t owns segments per*t .. per*t + per - 1)
	      {
	         const int per = (nseg + kPiThreads - 1) / kPiThreads;
	         const int c0 = tid * per, c1 = min (c0 + per, nseg);
	         double A = 1.0, B = 0.0;
	         for (int c = c0; c < c1; c ++) { B = S.resid [c] + (double)S.der [c] * B; A *= (double)S.der [c]; }
	         double dIn = block_affine_start (A, B, 0.0, S.warpA, S.warpB, nullptr);
	         for (int c = c0; c < c1; c ++) {
	            if (c > 0 && dIn != 0.0) {
	               const int a0 = S.anc [c];
	               S.est [a0] = wrap_2pi ((double)S.est [a0] + dIn);
	            }
	            dIn = S.resid [c] + (double)S.der [c] * dIn;
	         }
	      }
	      __syncthreads ();
	   }
	   if (!converged) {          // guaranteed-exact fall-back: one lane walks the window
	      if (tid == 0) {
	         float p = phi0;
	         for (int n = 0; n < Tw; n ++) {
	            S.est [n] = p;
	            float osc, cur;
	            p = pilot_step (L, sq, p, S.x [n], P.gain, P.omega, osc, cur);
	         }
	         S.G [0] = p; S.nseg = 1;
	      }
	      nFallback ++;
	      __syncthreads ();
	   }
	   itTotal += it; itMax = max (itMax, it);
	   const float phiEnd = S.G [S.nseg - 1];
	   __syncthreads ();
//	-- 3. per-sample outputs, lock detector -------------------------------------------------
	   float oscv [kPiPer];
	   float oprev = 0.f;
	   double bLk = 0.0;
	   int wraps = 0;
#pragma unroll
	   for (int j = 0; j < kPiPer; j ++) {
	      if (n0 + j < Tw) {
	         float cur;
	         const float pn = pilot_step (L, sq, S.est [n0 + j], S.x [n0 + j], P.gain, P.omega, oscv [j], cur);
	         ph [base + n0 + j] = cur;
	         wraps += pn < S.est [n0 + j];
	      }
	      else oscv [j] = 0.f;
	   }
	   {  // oscillator value of the sample before this thread's chunk (pilot_oldValue)
	      float last = oscv [kPiPer - 1];
	      if (n0 + kPiPer - 1 >= Tw) {      // chunk straddles the window end: take the last valid one
#pragma unroll
	         for (int j = 0; j < kPiPer; j ++) if (n0 + j == Tw - 1) last = oscv [j];
	      }
	      S.der [tid] = last;                // neighbours pass through der[] (free by now)
	      __syncthreads ();
	      oprev = tid == 0 ? oscPrev : S.der [tid - 1];
	   }
	   float vq [kPiPer];
#pragma unroll
	   for (int j = 0; j < kPiPer; j ++) {
	      if (n0 + j < Tw) {
	         const float quad = fdiv (fsub (oscv [j], oprev), P.omega);
	         vq [j] = fmul (lockAlpha, fmul (-quad, S.x [n0 + j]));
	         bLk = bLk * oneMinusLock + (double)vq [j];
	         oprev = oscv [j];
	      }
	      else { vq [j] = 0.f; bLk = bLk * oneMinusLock; }
	   }
	   double plk = lin_scan_start (bLk, S.carry [2], powLk, S.warpA);
//	   last index (within the window) at which the lock metric was NOT above threshold
	   int lastFalse = -1;
	   unsigned above = 0;
#pragma unroll
	   for (int j = 0; j < kPiPer; j ++) {
	      if (n0 + j < Tw) {
	         plk = plk * oneMinusLock + (double)vq [j];
	         if ((float)plk > 0.07f) above |= 1u << j; else lastFalse = n0 + j;
	         if (n0 + j == Tw - 1) S.carry [2] = plk;
	      }
	   }
	   int lf = lastFalse;
#pragma unroll
	   for (int k = 1; k < 32; k <<= 1) {
	      const int y = __shfl_up_sync (0xffffffffu, lf, k);
	      if (lane >= k) lf = max (lf, y);
	   }
	   if (lane == 31) S.warpI [warp] = lf;
	   int wsum = wraps;
#pragma unroll
	   for (int k = 16; k >= 1; k >>= 1) wsum += __shfl_xor_sync (0xffffffffu, wsum, k);
	   __syncthreads ();
	   if (lane == 0) S.anc [warp] = (int16_t)wsum;
	   int lfBefore = __shfl_up_sync (0xffffffffu, lf, 1);     // last false before this thread's chunk
	   if (lane == 0) lfBefore = -1;
	   for (int q = 0; q < warp; q ++) lfBefore = max (lfBefore, S.warpI [q]);
	   int runEnd = 0;
#pragma unroll
	   for (int j = 0; j < kPiPer; j ++) {
	      if (n0 + j < Tw) {
	         int run;
	         if ((above >> j) & 1u) {
	            run = lfBefore >= 0 ? (n0 + j - lfBefore) : min (runCarry + n0 + j + 1, 1 << 29);
	         }
	         else { run = 0; lfBefore = n0 + j; }
	         lk [base + n0 + j] = (uint8_t)(run > P.lock_half_rate);
	         if (n0 + j == Tw - 1) runEnd = run;
	      }
	   }
	   __syncthreads ();
	   int totalWraps = 0;
	   for (int q = 0; q < kPiWarps; q ++) totalWraps += S.anc [q];
//	-- 4. carry to the next window ----------------------------------------------------------
	   if (n0 <= Tw - 1 && Tw - 1 < n0 + kPiPer) {      // owner of the last sample publishes
	      S.warpI [0] = runEnd;
	      float lastOsc = 0.f;
#pragma unroll
	      for (int j = 0; j < kPiPer; j ++) if (n0 + j == Tw - 1) lastOsc = oscv [j];
	      S.G [1] = lastOsc;
	   }
	   __syncthreads ();
	   runCarry = S.warpI [0];
	   oscPrev  = S.G [1];
	   winc = ((double)phiEnd - (double)phi0 + 2 * M_PI * totalWraps) / Tw;
	   phi0 = phiEnd;
	   __syncthreads ();
	}
	if (tid == 0) {
	   st.fm_afc = (float)S.carry [0];
	   st.am_carr_ampl = (float)S.carry [1];
	   st.pilot_lock = (float)S.carry [2];
	   st.pilot_phase = phi0; st.pilot_old = oscPrev;
	   st.pilot_locked = runCarry > P.lock_half_rate;
	   st.pilot_stable_cnt = min (runCarry, P.lock_half_rate + 1);
	   if (iter_stats) {          // accumulate: a later time slice of the same call
	      int32_t *q = iter_stats + stream * 4;
	      q [0] = (accumulate ? q [0] : 0) + itTotal;
	      q [1] = max (accumulate ? q [1] : 0, itMax);
	      q [2] = (accumulate ? q [2] : 0) + nFallback;
	      q [3] = (accumulate ? q [3] : 0) + (M + kPiWin - 1) / kPiWin;
	   }
	}
}

}	// namespace sdrjfm
