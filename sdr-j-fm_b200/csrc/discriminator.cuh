// K2 — fm-rate parallel stage: finish DC removal, apply IQ gain and the constant complex
// gain of the decimator cascade, normalise, and run the memoryless part of
// fm_Demodulator::demodulate (src/fm/fm-demodulator.cpp:111-195).
//
// One CTA per IQ stream walks the stream in blocks of 256 threads x 8 samples.  The only
// recurrence here is the RF DC one-pole at the fm rate, r_m = beta r_{m-1} + alpha S_m
// (fm-processor.cpp:425 aggregated over 12 input samples): a block-wide
// scan in double gives every thread its start value, after which the thread steps
// through its 8 samples.  The previous normalised sample (Imin1, Qmin1) crosses threads
// through shared memory and crosses calls through StreamState.
#pragma once
#include "common.cuh"

namespace sdrjfm {

constexpr int kDiThreads = 256;
constexpr int kDiRun     = 8;                 // consecutive fm samples per thread
constexpr int kDiBlock   = kDiThreads * kDiRun;
constexpr int kDiTilesPerCta = 4;             // the 32 KB arctan table is staged once per CTA

struct DiscrParams {
	float  sumC, sumCm;         // sum of the plain composite taps C; of the DC-folded taps C'
	float  gb0, gb1, gb2;       // alpha * block means of the tail sums g (tables.cpp)
	float  Gre, Gim;            // constant complex gain of the cascade
	double alpha, beta;         // alpha = (float)1/inputRate; beta = (1 - alpha)^decim
	float  lgain, rgain;
	int32_t dc_remove, decoder;
	int32_t scan_only;          // station scan: the demodulator is not called, its state must not move
	int32_t exact;              // K1x ran (frontend_exact.cuh): U already is the reference's fm-rate sample, bit for bit
	// local oscillator on: gains and rotation were applied per input sample by K1; the DC value
	// the reference subtracted BEFORE the rotation comes back out as clamp (r) * gains *
	// Table[LOPhase at decim (m + lo_moff) + decim - 1] * H,  H = sum_t C[t] exp (+2 pi i lo t / inputRate)
	const float2 *lo_tab;
	int32_t lo_rate, lo_hz, lo_moff, decim;      // decim: input samples per fm-rate sample
	int64_t lo_phase;
	float  Hre, Him;
};

// compAtan::atan2, src/various/Xtan2.cpp:56-100.  Only the first-octant table PPY is kept
// (in shared memory); the other seven tables are single float operations on PPY
// (Xtan2.cpp:30-37) and are re-derived on the fly with the same roundings.
// (int)(q + 0.5) with the sum formed in double (Xtan2.cpp:74-97).  For 0 <= q <= 8192 the float sum
// truncates to the same integer: q + 0.5 is exact unless it crosses into the next binade, and there
// the rounding (< 2^-11) cannot cross an integer from below.
__device__ __forceinline__ int atan_index (float num_scaled, float den) {
	return (int)fadd (fdiv (num_scaled, den), 0.5f);
}

// The reference picks one of eight tables by quadrant and octant (Xtan2.cpp:74-97); every one of them is the first-
// octant table PPY looked up at (int)(8192 * small / large + 0.5) and then negated and / or shifted by pi/2 or pi with
// ONE float operation (Xtan2.cpp:30-37).  Division, rounding and the final add are sign-symmetric, so the eight
// branches collapse into one division, one look-up and one add, chosen by three comparisons — the same floats,
// without divergence:
//     x > 0, y >= 0 :  x >= y :  t                |  else  H - t = -(t - H)
//     x > 0, y <  0 :  x >= -y: -t                |  else  t - H
//     x < 0, y >= 0 : -x >= y :  S - t = -(t - S) |  else  t + H
//     x < 0, y <  0 :  x <= y :  t - S            |  else -H - t = -(t + H)
__device__ __forceinline__ float lut_atan2 (const float *PPY, float y, float x) {
const float S = (float)M_PI;
const float H = fmul (S, 0.5f);
	if (isinf (x) || isinf (y)) return 0.f;
	if (isnan (x) || isnan (y)) return 0.f;
	if (x == 0.f) {
	   if (y == 0.f) return 0.f;
	   return y > 0.f ? (float)(M_PI / 2) : (float)(-M_PI / 2);
	}
const bool xpos = x > 0.f, ypos = y >= 0.f;
const float ax = fabsf (x), ay = fabsf (y);
const bool swap = ay > ax;                               // the octant: every branch above keeps the ratio <= 1, ties unswapped
const float t = PPY [atan_index (fmul (8192.f, swap ? ax : ay), swap ? ay : ax)];
const float add = xpos ? (swap ? -H : 0.f) : (swap ? H : -S);
const bool neg = xpos ? (ypos == swap) : (ypos != swap);
const float r = add == 0.f ? t : fadd (t, add);
	return neg ? -r : r;
}

struct dcplx { double re, im; };

__device__ __forceinline__ dcplx shfl_up_d (dcplx v, int d) {
dcplx r;
	r.re = __shfl_up_sync (0xffffffffu, v.re, d);
	r.im = __shfl_up_sync (0xffffffffu, v.im, d);
	return r;
}

// What K2 carries from one call to the next, snapshotted by the pre-pass so that the tiles
// of a call (which all read it) never race with the tile that writes the new values.
struct DiscrSnap {
	double dc_re, dc_im;
	float2 sp1, sp2;           // block sums S[-1], S[-2]
	float2 p1, p2;             // normalised samples -1 and -2 (Imin1/Qmin1, Imin2/Qmin2)
};

// pre-pass, grid (tiles, streams): zero-state response of the RF DC one-pole over each FULL tile
// (tileB), and the snapshot of the carried state.
__global__ void __launch_bounds__ (kDiThreads)
dc_tile_kernel (const float2 *__restrict__ Ssum, int64_t pitch, int32_t M, DiscrParams P,
                const StreamState *__restrict__ state, dcplx *__restrict__ tileB, int32_t ntiles,
                DiscrSnap *__restrict__ snap) {
__shared__ dcplx sWarp [kDiThreads / 32];
const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
const int stream = blockIdx.y, tile = blockIdx.x;
	if (tile == 0 && tid == 0) {
	   const StreamState &st = state [stream];
	   DiscrSnap q;
	   q.dc_re = st.dc_re; q.dc_im = st.dc_im;
	   q.sp1 = make_float2 (st.sprev [0], st.sprev [1]); q.sp2 = make_float2 (st.sprev [2], st.sprev [3]);
	   q.p1 = make_float2 (st.Imin1, st.Qmin1); q.p2 = make_float2 (st.Imin2, st.Qmin2);
	   snap [stream] = q;
	}
	if (!P.dc_remove || (int64_t)(tile + 1) * kDiBlock > M) return;      // only full tiles are predecessors
const float2 *Ss = Ssum + (int64_t)stream * pitch + (int64_t)tile * kDiBlock + tid * kDiRun;
double pw [5];
	{  double b8 = P.beta; b8 *= b8; b8 *= b8; b8 *= b8; pw [0] = b8;
#pragma unroll
	   for (int k = 1; k < 5; k ++) pw [k] = pw [k - 1] * pw [k - 1]; }
dcplx a; a.re = 0.0; a.im = 0.0;
#pragma unroll
	for (int j = 0; j < kDiRun; j ++) {
	   const float2 s = Ss [j];
	   a.re = a.re * P.beta + P.alpha * (double)s.x;
	   a.im = a.im * P.beta + P.alpha * (double)s.y;
	}
#pragma unroll
	for (int k = 0; k < 5; k ++) {
	   dcplx y = shfl_up_d (a, 1 << k);
	   if (lane >= (1 << k)) { a.re += y.re * pw [k]; a.im += y.im * pw [k]; }
	}
	if (lane == 31) sWarp [warp] = a;
	__syncthreads ();
	if (tid == 0) {
	   const double pwarp = pw [4] * pw [4];              // beta^(8*32)
	   dcplx w; w.re = 0.0; w.im = 0.0;
	   for (int q = 0; q < kDiThreads / 32; q ++) { w.re = w.re * pwarp + sWarp [q].re; w.im = w.im * pwarp + sWarp [q].im; }
	   tileB [(int64_t)stream * ntiles + tile] = w;
	}
}

// one fm-rate sample: DC subtraction (free-running or clamped), gains / LO correction, constant
// complex gain, magnitude, normalisation (fm-processor.cpp:423-446,462-466; fm-demodulator.cpp:119-126)
__device__ __forceinline__ void disc_sample (const DiscrParams &P, float2 u, float2 s, float2 sm1, float2 sm2,
                                             double rre, double rim, int64_t j,
                                             float2 &z, float &za, float2 &nq) {
float2 c = make_float2 (0.f, 0.f);
	if (P.dc_remove) {
	   const float lim = 0.01f;            // DCRlimit, fm-processor.cpp:429
	   c.x = fminf (fmaxf ((float)rre, -lim), lim);
	   c.y = fminf (fmaxf ((float)rim, -lim), lim);
	}
//	K1 ran the DC-folded taps C' = C + alpha g.  Free-running component: subtract r sum(C').
//	Component on the clamp (or DC removal off): the subtracted value is the constant c, so take
//	alpha sum_i g[i] x[n-i] back out via the 12-sample block sums.
const float wx = P.gb0 * s.x + P.gb1 * sm1.x + P.gb2 * sm2.x;
const float wy = P.gb0 * s.y + P.gb1 * sm1.y + P.gb2 * sm2.y;
const bool freex = P.dc_remove && c.x == (float)rre;
const bool freey = P.dc_remove && c.y == (float)rim;
const float kx = freex ? c.x * P.sumCm : c.x * P.sumC + wx;
const float ky = freey ? c.y * P.sumCm : c.y * P.sumC + wy;
//	IQ gain (fm-processor.cpp:462-464) commutes with the real-tap FIR
float vx = (u.x - kx) * P.lgain;
float vy = (u.y - ky) * P.rgain;
	if (P.lo_tab) {
	   int64_t t = (P.lo_phase - (int64_t)P.lo_hz * ((int64_t)P.decim * (j + (int64_t)P.lo_moff) + P.decim)) % P.lo_rate;
	   if (t < 0) t += P.lo_rate;
	   const float2 o = P.lo_tab [t];
	   const float2 a = make_float2 (c.x * P.lgain, c.y * P.rgain);
	   const float2 b = make_float2 (a.x * o.x - a.y * o.y, a.x * o.y + a.y * o.x);
	   vx = u.x - (b.x * P.Hre - b.y * P.Him);
	   vy = u.y - (b.x * P.Him + b.y * P.Hre);
	}
	z = make_float2 (vx * P.Gre - vy * P.Gim, vx * P.Gim + vy * P.Gre);
//	std::abs (complex<float>) (fm-demodulator.cpp:119).  z itself carries ~1e-6 of re-association
//	error against the reference (composite FIR), so the magnitude is taken in float (1e-7) instead of
//	through a double square root, and the two divisions share one reciprocal: q = a r, corrected
//	once with the exact residual (q + (a - q b) r), which is the correctly rounded quotient.
//	The decoders other than MIXED read look-up tables whose index flips on a 1-ulp change of I or Q
//	(arcsine table, the PLL's sine table): for them the magnitude is the reference's own rounding,
//	hypotf = the double-precision root rounded once to float.
	if (P.decoder != 3 || P.exact) za = (float)sqrt ((double)z.x * (double)z.x + (double)z.y * (double)z.y);
	else za = __fsqrt_rn (fmaf (z.x, z.x, fmul (z.y, z.y)));
	if (za <= 0.001f) nq = make_float2 (0.001f, 0.001f);          // :120-122
	else {
	   const float r = __frcp_rn (za);
	   float qx = fmul (z.x, r), qy = fmul (z.y, r);
	   qx = fmaf (fmaf (-qx, za, z.x), r, qx);
	   qy = fmaf (fmaf (-qy, za, z.y), r, qy);
	   nq = make_float2 (qx, qy);
	}
}

// U, Ssum : front-end outputs; res_raw: discriminator output before AFC (float);
// zabs: |z| ; fmz (optional): the fm-rate complex sample (tap after fmBand_2).
// grid (tiles of kDiBlock samples, streams): every tile is independent given the snapshot and
// the tile aggregates of the DC one-pole.
__global__ void __launch_bounds__ (kDiThreads, 3)
discriminator_kernel (const float2 *__restrict__ U, const float2 *__restrict__ Ssum,
                      int64_t pitch, int32_t M, DiscrParams P,
                      const float *__restrict__ atanPPY, const float *__restrict__ arcsine,
                      StreamState *__restrict__ state, const dcplx *__restrict__ tileB, int32_t ntiles,
                      const DiscrSnap *__restrict__ snap,
                      float *__restrict__ res_raw, float *__restrict__ zabs,
                      float2 *__restrict__ iqn, float2 *__restrict__ fmz) {
__shared__ float  sPPY [8193 + 3];
__shared__ dcplx  sWarp [kDiThreads / 32];
__shared__ float2 sLastIQ [kDiThreads + 1];
__shared__ float2 sLastIQ2 [kDiThreads + 1];
__shared__ dcplx  sCarry;
const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
const int stream = blockIdx.y;
StreamState &st = state [stream];
const float2 *Us = U + (int64_t)stream * pitch;
const float2 *Ss = Ssum + (int64_t)stream * pitch;
const DiscrSnap sn = snap [stream];

	for (int i = tid; i < 8193; i += kDiThreads) sPPY [i] = atanPPY [i];
	for (int tile = blockIdx.x * kDiTilesPerCta; tile < min ((int)(blockIdx.x + 1) * kDiTilesPerCta, ntiles); tile ++) {
const int64_t base = (int64_t)tile * kDiBlock;
//	powers of beta used by the scan: pw[k] = beta^(8 * 2^k)
double pw [6];
	{  double b8 = P.beta; b8 *= b8; b8 *= b8; b8 *= b8;   // beta^8
	   pw [0] = b8;
#pragma unroll
	   for (int k = 1; k < 6; k ++) pw [k] = pw [k - 1] * pw [k - 1];
	}
	if (tid == 0) {
//	   DC estimate entering the tile: carried state pushed through the earlier (full) tiles
	   dcplx r; r.re = sn.dc_re; r.im = sn.dc_im;
	   if (P.dc_remove && tile > 0) {
	      double pt = pw [5]; pt *= pt; pt *= pt; pt *= pt;           // beta^(8*32*8) = beta^2048
	      for (int q = 0; q < tile; q ++) {
	         const dcplx b = tileB [(int64_t)stream * ntiles + q];
	         r.re = r.re * pt + b.re; r.im = r.im * pt + b.im;
	      }
	   }
	   sCarry = r;
//	   the two normalised samples before the tile
	   if (tile == 0) { sLastIQ [0] = sn.p1; sLastIQ2 [0] = sn.p2; }
	   else {
	      const float2 s1 = Ss [base - 1], s2 = Ss [base - 2], s3 = Ss [base - 3];
	      const float2 s4 = base >= 4 ? Ss [base - 4] : sn.sp1;
	      // r after sample base-1 is the carry; one step back for sample base-2
	      const double r2re = P.dc_remove ? (r.re - P.alpha * (double)s1.x) / P.beta : 0.0;
	      const double r2im = P.dc_remove ? (r.im - P.alpha * (double)s1.y) / P.beta : 0.0;
	      float2 z, n1, n2; float za;
	      disc_sample (P, Us [base - 1], s1, s2, s3, r.re, r.im, base - 1, z, za, n1);
	      disc_sample (P, Us [base - 2], s2, s3, s4, r2re, r2im, base - 2, z, za, n2);
	      sLastIQ [0] = n1; sLastIQ2 [0] = n2;
	   }
	}
	__syncthreads ();

	{
	   const int64_t j0 = base + (int64_t)tid * kDiRun;
	   float2 u [kDiRun], s [kDiRun];
#pragma unroll
	   for (int j = 0; j < kDiRun; j ++) {
	      const bool ok = j0 + j < M;
	      u [j] = ok ? Us [j0 + j] : make_float2 (0.f, 0.f);
	      s [j] = ok ? Ss [j0 + j] : make_float2 (0.f, 0.f);
	   }
//	-- DC one-pole: thread aggregate, block scan, per-sample values -----------------
	   dcplx a; a.re = 0.0; a.im = 0.0;
	   if (P.dc_remove) {
#pragma unroll
	      for (int j = 0; j < kDiRun; j ++) {
	         a.re = a.re * P.beta + P.alpha * (double)s [j].x;
	         a.im = a.im * P.beta + P.alpha * (double)s [j].y;
	      }
	   }
	   dcplx inc = a;                                   // inclusive scan over threads
#pragma unroll
	   for (int k = 0; k < 5; k ++) {
	      dcplx y = shfl_up_d (inc, 1 << k);
	      if (lane >= (1 << k)) { inc.re += y.re * pw [k]; inc.im += y.im * pw [k]; }
	   }
	   if (lane == 31) sWarp [warp] = inc;
	   __syncthreads ();
//	start value of this thread = carry decayed to the thread + all earlier threads
	   dcplx r;
	   {
	      dcplx w; w.re = sCarry.re; w.im = sCarry.im;   // value at tile start
	      for (int q = 0; q < warp; q ++) {
	         w.re = w.re * pw [5] + sWarp [q].re;
	         w.im = w.im * pw [5] + sWarp [q].im;
	      }
	      double dl = 1.0;
#pragma unroll
	      for (int k = 0; k < 5; k ++) if (lane & (1 << k)) dl *= pw [k];
	      dcplx excl = shfl_up_d (inc, 1);
	      if (lane == 0) { excl.re = 0.0; excl.im = 0.0; }
	      r.re = w.re * dl + excl.re;
	      r.im = w.im * dl + excl.im;
	   }
	   // block sums of the two fm-rate samples before this thread's run (for the clamped case)
	   float2 sm1, sm2;
	   sm1 = (j0 >= 1 && j0 - 1 < M) ? Ss [j0 - 1] : sn.sp1;
	   sm2 = (j0 >= 2 && j0 - 2 < M) ? Ss [j0 - 2] : (j0 == 1 ? sn.sp1 : sn.sp2);
//	-- per-sample: corrected, gained, normalised sample ------------------------------
	   float2 z [kDiRun], nq [kDiRun];
	   float  za [kDiRun];
#pragma unroll
	   for (int j = 0; j < kDiRun; j ++) {
	      if (P.dc_remove) {
	         r.re = r.re * P.beta + P.alpha * (double)s [j].x;
	         r.im = r.im * P.beta + P.alpha * (double)s [j].y;
	         if (j0 + j == M - 1) { st.dc_re = r.re; st.dc_im = r.im; }    // state handed to the next call
	      }
	      disc_sample (P, u [j], s [j], sm1, sm2, r.re, r.im, j0 + j, z [j], za [j], nq [j]);
	      sm2 = sm1; sm1 = s [j];
	      if (j0 + j == M - 1) {
	         st.sprev [0] = sm1.x; st.sprev [1] = sm1.y; st.sprev [2] = sm2.x; st.sprev [3] = sm2.y;
	      }
	   }
//	hand the last normalised samples to the next thread
	   sLastIQ [tid + 1]  = nq [kDiRun - 1];
	   sLastIQ2 [tid + 1] = nq [kDiRun - 2];
	   __syncthreads ();
	   float2 p1 = sLastIQ [tid];         // Imin1, Qmin1 before this thread's first sample
	   float2 p2 = sLastIQ2 [tid];        // Imin2, Qmin2 (the sample two back)
#pragma unroll
	   for (int j = 0; j < kDiRun; j ++) {
	      const float I = nq [j].x, Q = nq [j].y;
	      float res;
	      switch (P.decoder) {
	         case 4: {   // ComplexBasebandDelay: argX (z * conj (prev)), fm-demodulator.cpp:174-177
	            const float2 m = cmul_rn (make_float2 (I, Q), make_float2 (p1.x, -p1.y));
	            res = lut_atan2 (sPPY, m.y, m.x);
	            break;
	         }
	         case 5: {   // RealBasebandDelay, :179-187
	            float t = (float)((double)fadd (fsub (fmul (p1.x, Q), fmul (p1.y, I)), 1.0f) / 2.0);
	            int index = (int)floorf (fmul (t, 32768.f));
	            if (index < 0) index = 0;
	            if (index >= 32768) index = 32768;
	            res = arcsine [index];
	            break;
	         }
	         case 6: {   // DifferenceBased, :189-194
	            res = fsub (fmul (p1.x, fsub (Q, p2.y)), fmul (p1.y, fsub (I, p2.x)));
	            res = fdiv (res, fmul (fadd (fmul (p1.x, p1.x), fmul (p1.y, p1.y)), sqrtf (2.f)));
	            break;
	         }
	         default:    // MixedDemodulator, :168-171 (also feeds nothing for the PLL decoder)
	            res = lut_atan2 (sPPY, fsub (fmul (Q, p1.x), fmul (I, p1.y)),
	                                   fadd (fmul (I, p1.x), fmul (Q, p1.y)));
	      }
	      p2 = p1;
	      p1 = make_float2 (I, Q);
	      if (j0 + j == M - 1 && !P.scan_only) {
	         st.Imin1 = p1.x; st.Qmin1 = p1.y; st.Imin2 = p2.x; st.Qmin2 = p2.y;
	      }
	      if (j0 + j < M) {
	         const int64_t o = (int64_t)stream * pitch + j0 + j;
	         res_raw [o] = res;
	         zabs [o] = za [j];
	         if (iqn) iqn [o] = P.decoder == 1 ? z [j] : nq [j];       // AM: pllC runs on the raw sample
	         if (fmz) fmz [o] = z [j];
	      }
	   }
	}
	   __syncthreads ();          // the per-tile shared scratch is reused by the next tile
	}
}

}	// namespace sdrjfm
