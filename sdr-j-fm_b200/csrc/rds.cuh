// K5 — RDS branch at the fm rate (fm-processor.cpp:733-758, 551-553):
//   rdsBandPassFilter = fftFilter (32768, 768).setBand (57 kHz -+ 2.4 kHz)      :166-168, Pass(float) = 3 Re{conv}
//   rdsHilbertFilter  = fftFilterHilbert (32768, 768)                           fft-filters.cpp:165-201
//   theta = 3 * pilot phase delayed by RDS_SAMPLE_DELAY = 64000 samples         fm-processor.h:53, .cpp:744-746
//   rdsDataCplx = (cos theta, -sin theta) * hilbert                             :752-754
//   rdsDecimator = DecimatingFIR (11, 12000, 192000, 8) -> 24 kHz               :382, :553
//
// Both filters are overlap-add FFT filters with NumofSamples = 32000:
//   * the band-pass is a true linear convolution, delayed by one block:
//         bp[n] = 3 sum_{j<768} r[j] d[n - 32000 - j],  r = Re (BandPassFIR kernel)
//   * the Hilbert "filter" multiplies the spectrum of each ZERO-PADDED 32000-sample block by
//     the analytic mask (1,2,..,2,1,0,..,0).  That is NOT a convolution: the result depends on
//     where the blocks fall, so the same blocks are transformed here:
//     block k = bp[32000 k .. 32000 (k+1)), emitted one block later; the last 768 samples of a
//     block's 32768-point result are added to the head of the next one (Overloop).
//     The real part of the masked inverse transform is the zero-padded block itself (exactly),
//     the imaginary part is its circular Hilbert transform: only the latter needs FFTs.
// rds_block_kernel produces, per stream and per block k, bp block k (overlap-save fast
// convolution) and its Hilbert transform with four in-place 16384-point complex FFTs in
// shared memory (real-input packing).  rds_mix_kernel then forms rdsDataCplx per sample and
// rds_decim_kernel the 24 kHz output, restating DecimatingFIR::Pass tap by tap.
#pragma once
#include "common.cuh"
#include "sequential.cuh"

namespace sdrjfm {

constexpr int kRdsN       = 32768;              // FFT_SIZE, fm-constants.h:106
constexpr int kRdsNh      = kRdsN / 2;          // complex FFT length after real packing
constexpr int kRdsTaps    = 768;                // PILOTFILTER_SIZE, fm-constants.h:105
constexpr int kRdsBlock   = kRdsN - kRdsTaps;   // 32000 = NumofSamples
constexpr int kRdsDelay   = 2 * kRdsBlock;      // RDS_SAMPLE_DELAY
constexpr int kRdsRing    = 131072;             // per-stream history of demod / pilot phase
constexpr int kRdsThreads = 1024;
constexpr int kRdsFftSmem = kRdsNh * (int)sizeof (float2);      // 131072 B
constexpr int kRdsDecTaps = 11;
constexpr int kRdsDecim   = 8;

__device__ __forceinline__ float2 cmulf (float2 a, float2 b) {
	return make_float2 (a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cconj (float2 a) { return make_float2 (a.x, -a.y); }
__device__ __forceinline__ int brev14 (int k) { return (int)(__brev ((unsigned)k) >> 18); }

// in-place radix-2 FFTs of length kRdsNh in shared memory.  tws = per-stage twiddle table:
// tws[half + pos] = exp (-2 pi i pos / (2 half)), pos < half  (kRdsNh entries, coalesced per stage).
// Every thread first loads the operands and twiddles of all its butterflies of a stage, then
// computes, then stores, so the shared/global load latencies overlap instead of serialising.
constexpr int kRdsBpt = kRdsNh / 2 / kRdsThreads;       // butterflies per thread per stage (8)

// Two radix-2 stages are merged into one pass over shared memory (a radix-4 butterfly written as the
// two radix-2 layers it consists of, outputs left where the radix-2 stages would have put them, so the
// overall permutation is still plain bit reversal): 7 passes and barriers instead of 14.
constexpr int kRdsQpt = kRdsNh / 4 / kRdsThreads;       // quads per thread per pass (4)
static_assert ((kRdsNh & (kRdsNh - 1)) == 0 && (31 - __builtin_clz (kRdsNh)) % 2 == 0, "an even number of radix-2 stages");

// forward: decimation in frequency, natural order in -> bit-reversed order out
__device__ void fft_dif (float2 *a, const float2 *__restrict__ tws) {
	for (int h = kRdsNh / 2; h >= 2; h >>= 2) {           // stages `half = h` and `half = h / 2`
	   const int hh = h >> 1;
	   float2 x0 [kRdsQpt], x1 [kRdsQpt], x2 [kRdsQpt], x3 [kRdsQpt], w [kRdsQpt], w2 [kRdsQpt];
	   int idx [kRdsQpt];
#pragma unroll
	   for (int q = 0; q < kRdsQpt; q ++) {
	      const int b = threadIdx.x + q * kRdsThreads;
	      const int pos = b & (hh - 1);
	      idx [q] = ((b - pos) << 2) + pos;
	      x0 [q] = a [idx [q]]; x1 [q] = a [idx [q] + hh]; x2 [q] = a [idx [q] + h]; x3 [q] = a [idx [q] + h + hh];
	      w [q] = __ldg (tws + h + pos);                 // exp (-2 pi i pos / (2 h))
	      w2 [q] = __ldg (tws + hh + pos);               // exp (-2 pi i pos / h)
	   }
#pragma unroll
	   for (int q = 0; q < kRdsQpt; q ++) {
	      const float2 A = make_float2 (x0 [q].x + x2 [q].x, x0 [q].y + x2 [q].y);
	      const float2 B = make_float2 (x1 [q].x + x3 [q].x, x1 [q].y + x3 [q].y);
	      const float2 C = cmulf (make_float2 (x0 [q].x - x2 [q].x, x0 [q].y - x2 [q].y), w [q]);
	      const float2 T = cmulf (make_float2 (x1 [q].x - x3 [q].x, x1 [q].y - x3 [q].y), w [q]);
	      const float2 D = make_float2 (T.y, -T.x);     // twiddle of pos + h/2 is -i times that of pos
	      a [idx [q]]          = make_float2 (A.x + B.x, A.y + B.y);
	      a [idx [q] + hh]     = cmulf (make_float2 (A.x - B.x, A.y - B.y), w2 [q]);
	      a [idx [q] + h]      = make_float2 (C.x + D.x, C.y + D.y);
	      a [idx [q] + h + hh] = cmulf (make_float2 (C.x - D.x, C.y - D.y), w2 [q]);
	   }
	   __syncthreads ();
	}
}
// inverse: decimation in time with conjugate twiddles, bit-reversed order in -> natural order out
__device__ void ifft_dit (float2 *a, const float2 *__restrict__ tws) {
	for (int h = 1; h <= kRdsNh / 4; h <<= 2) {           // stages `half = h` and `half = 2 h`
	   float2 x0 [kRdsQpt], x1 [kRdsQpt], x2 [kRdsQpt], x3 [kRdsQpt], w [kRdsQpt], w2 [kRdsQpt];
	   int idx [kRdsQpt];
#pragma unroll
	   for (int q = 0; q < kRdsQpt; q ++) {
	      const int b = threadIdx.x + q * kRdsThreads;
	      const int pos = b & (h - 1);
	      idx [q] = ((b - pos) << 2) + pos;
	      x0 [q] = a [idx [q]]; x1 [q] = a [idx [q] + h]; x2 [q] = a [idx [q] + 2 * h]; x3 [q] = a [idx [q] + 3 * h];
	      w [q] = cconj (__ldg (tws + h + pos));
	      w2 [q] = cconj (__ldg (tws + 2 * h + pos));
	   }
#pragma unroll
	   for (int q = 0; q < kRdsQpt; q ++) {
	      const float2 t1 = cmulf (x1 [q], w [q]), t2 = cmulf (x3 [q], w [q]);
	      const float2 y0 = make_float2 (x0 [q].x + t1.x, x0 [q].y + t1.y), y1 = make_float2 (x0 [q].x - t1.x, x0 [q].y - t1.y);
	      const float2 y2 = make_float2 (x2 [q].x + t2.x, x2 [q].y + t2.y), y3 = make_float2 (x2 [q].x - t2.x, x2 [q].y - t2.y);
	      const float2 t = cmulf (y2, w2 [q]), u = cmulf (y3, w2 [q]);
	      const float2 tp = make_float2 (-u.y, u.x);    // conjugate twiddle of pos + h is +i times that of pos
	      a [idx [q]]         = make_float2 (y0.x + t.x, y0.y + t.y);
	      a [idx [q] + h]     = make_float2 (y1.x + tp.x, y1.y + tp.y);
	      a [idx [q] + 2 * h] = make_float2 (y0.x - t.x, y0.y - t.y);
	      a [idx [q] + 3 * h] = make_float2 (y1.x - tp.x, y1.y - tp.y);
	   }
	   __syncthreads ();
	}
}

// Spectrum pass between a forward and an inverse transform of a REAL 32768-point sequence held
// as 16384 packed complex values (bit-reversed order): untangle X[k], X[k+Nh], multiply by
// (A[k], B[k]), re-tangle for the inverse.  HILBERT: A = -i, B = +i (0 at k = 0); else
// A = R[k], B = conj (R[Nh-k]) with R the (scaled) spectrum of the real band-pass taps.
template <bool HILBERT>
__device__ __forceinline__ float2 spectrum_one (float2 Zk, float2 Zc /* conj Z[Nh-k] */, float2 W,
                                                int k, const float2 *__restrict__ R) {
const float2 E = make_float2 (0.5f * (Zk.x + Zc.x), 0.5f * (Zk.y + Zc.y));
const float2 D = make_float2 (0.5f * (Zk.x - Zc.x), 0.5f * (Zk.y - Zc.y));
const float2 O = make_float2 (D.y, -D.x);                   // -i D
const float2 WO = cmulf (W, O);
float2 X0 = make_float2 (E.x + WO.x, E.y + WO.y);           // X[k]
float2 X1 = make_float2 (E.x - WO.x, E.y - WO.y);           // X[k + Nh]
float2 Y0, Y1;
	if (HILBERT) {
	   const float s = (k == 0) ? 0.f : 1.0f / kRdsNh;
	   Y0 = make_float2 (X0.y * s, -X0.x * s);                  // -i X[k]
	   Y1 = make_float2 (-X1.y * s, X1.x * s);                  // +i X[k + Nh]
	}
	else {
	   Y0 = cmulf (X0, R [k]);
	   Y1 = cmulf (X1, cconj (R [kRdsNh - k]));
	}
const float2 Ep = make_float2 (0.5f * (Y0.x + Y1.x), 0.5f * (Y0.y + Y1.y));
const float2 Dp = make_float2 (0.5f * (Y0.x - Y1.x), 0.5f * (Y0.y - Y1.y));
const float2 Op = cmulf (Dp, cconj (W));
	return make_float2 (Ep.x - Op.y, Ep.y + Op.x);              // E' + i O'
}

template <bool HILBERT>
__device__ void spectrum_pass (float2 *a, const float2 *__restrict__ tw, const float2 *__restrict__ R) {
	for (int k = threadIdx.x; k <= kRdsNh / 2; k += kRdsThreads) {
	   const int k2 = (kRdsNh - k) & (kRdsNh - 1);
	   const int p = brev14 (k), p2 = brev14 (k2);
	   const float2 Z1 = a [p], Z2 = a [p2];
	   const float2 n1 = spectrum_one<HILBERT> (Z1, cconj (Z2), tw [k], k, R);
	   if (k2 != k) {
	      const float2 n2 = spectrum_one<HILBERT> (Z2, cconj (Z1), tw [k2], k2, R);
	      a [p2] = n2;
	   }
	   a [p] = n1;
	}
	__syncthreads ();
}

// dring : [S][kRdsRing] demod by rds sample index (ring);  blk: Hilbert block index k >= 0
// bpb   : [S][2][32000] band-pass block (slot k & 1);  hib : [S][2][32768] its Hilbert transform
// R     : kRdsNh + 1 complex, spectrum of 3 r[j] scaled by 1/Nh;  tw: exp (-2 pi i k / N), k < Nh (spectrum
//         pass);  tws: per-stage twiddles of the Nh-point transforms
__global__ void __launch_bounds__ (kRdsThreads, 1)
rds_block_kernel (const float *__restrict__ dring, int64_t blk,
                  const float2 *__restrict__ tw, const float2 *__restrict__ tws, const float2 *__restrict__ R,
                  float *__restrict__ bpb, float *__restrict__ hib) {
extern __shared__ __align__ (16) float2 fa [];
const int tid = threadIdx.x;
const int stream = blockIdx.x;
const float *dr = dring + (int64_t)stream * kRdsRing;
float *bpo = bpb + ((int64_t)stream * 2 + (blk & 1)) * kRdsBlock;
float *hio = hib + ((int64_t)stream * 2 + (blk & 1)) * kRdsN;
const int64_t base = (blk - 1) * (int64_t)kRdsBlock - (kRdsTaps - 1);   // rds index of s[0]
//	overlap-save segment s[i] = d[base + i], i < 32767, packed two per complex
	for (int m = tid; m < kRdsNh; m += kRdsThreads) {
	   const int64_t i0 = base + 2 * m, i1 = i0 + 1;
	   const float v0 = i0 >= 0 ? dr [i0 & (kRdsRing - 1)] : 0.f;
	   const float v1 = (i1 >= 0 && 2 * m + 1 < kRdsN - 1) ? dr [i1 & (kRdsRing - 1)] : 0.f;
	   fa [m] = make_float2 (v0, v1);
	}
	__syncthreads ();
	fft_dif (fa, tws);
	spectrum_pass<false> (fa, tw, R);
	ifft_dit (fa, tws);
//	valid outputs c[767 + i], i < 32000 -> bp block; re-pack zero-padded for the Hilbert pass
float2 z [kRdsNh / kRdsThreads];
#pragma unroll
	for (int q = 0; q < kRdsNh / kRdsThreads; q ++) {
	   const int m = tid + q * kRdsThreads;              // z[m] = (bp[2m], bp[2m+1])
	   float2 v = make_float2 (0.f, 0.f);
	   if (2 * m < kRdsBlock) v = make_float2 (fa [(kRdsTaps - 2) / 2 + m].y, fa [kRdsTaps / 2 + m].x);
	   z [q] = v;
	}
	__syncthreads ();
#pragma unroll
	for (int q = 0; q < kRdsNh / kRdsThreads; q ++) {
	   const int m = tid + q * kRdsThreads;
	   fa [m] = z [q];
	   if (2 * m < kRdsBlock) reinterpret_cast<float2 *>(bpo) [m] = z [q];
	}
	__syncthreads ();
	fft_dif (fa, tws);
	spectrum_pass<true> (fa, tw, R);
	ifft_dit (fa, tws);
	for (int m = tid; m < kRdsNh; m += kRdsThreads) reinterpret_cast<float2 *>(hio) [m] = fa [m];
}

// appends this call's demod and pilot phase to the per-stream rings (rds sample index n0 + m)
__global__ void rds_append_kernel (const float *__restrict__ demod, const float *__restrict__ phase,
                                   int64_t pitch, int32_t m_begin, int32_t m_end, int64_t n0,
                                   float *__restrict__ dring, float *__restrict__ pring) {
const int stream = blockIdx.y;
const int m = m_begin + blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= m_end) return;
const int64_t slot = (n0 + m) & (kRdsRing - 1);
	dring [(int64_t)stream * kRdsRing + slot] = demod [(int64_t)stream * pitch + m];
	pring [(int64_t)stream * kRdsRing + slot] = phase [(int64_t)stream * pitch + m];
}

// rdsDataCplx for local samples [m_begin, m_end), all inside output block K = (n0 + m) / 32000
__global__ void rds_mix_kernel (const float *__restrict__ pring, const float *__restrict__ bpb,
                                const float *__restrict__ hib, int64_t pitch,
                                int32_t m_begin, int32_t m_end, int64_t n0,
                                float2 *__restrict__ rdsc) {
const int stream = blockIdx.y;
const int m = m_begin + blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= m_end) return;
const int64_t n = n0 + m;
const int64_t K = n / kRdsBlock;
const int i = (int)(n - K * kRdsBlock);
float2 hil = make_float2 (0.f, 0.f);
	if (K >= 1) {
	   const int64_t k = K - 1;
	   hil.x = bpb [((int64_t)stream * 2 + (k & 1)) * kRdsBlock + i];
	   hil.y = hib [((int64_t)stream * 2 + (k & 1)) * kRdsN + i];
	   if (i < kRdsTaps && k >= 1)                       // Overloop: tail of the previous block's transform
	      hil.y += hib [((int64_t)stream * 2 + ((k - 1) & 1)) * kRdsN + kRdsBlock + i];
	}
//	thePhase = 3 * (rdsPhaseBuffer [rdsPhaseIndex] + 0): the pilot phase 64000 samples back (zeros before)
const float pold = n >= kRdsDelay ? pring [(int64_t)stream * kRdsRing + ((n - kRdsDelay) & (kRdsRing - 1))] : 0.f;
const float th = fmul (3.0f, fadd (pold, 0.0f));
const float2 osc = make_float2 (cosf (th), -sinf (th));
	rdsc [(int64_t)stream * pitch + m] = cmul_rn (osc, hil);
}

// DecimatingFIR::Pass for rdsDecimator (fir-filters.cpp:397-424): output q <-> input 8 q + 7,
// y = sum_{i<11} x[8q+7-i] * K[i], complex kernel, taps in order i = 0..10, no contraction.
// hist: [S][10] last inputs of the previous calls (complex); c0: inputs consumed so far
__global__ void rds_decim_kernel (const float2 *__restrict__ rdsc, int64_t pitch, int32_t M,
                                  int64_t c0, const float2 *__restrict__ taps,
                                  const float2 *__restrict__ hist, float2 *__restrict__ new_hist,
                                  float2 *__restrict__ out, int64_t out_pitch, int32_t nout) {
const int stream = blockIdx.y;
const int o = blockIdx.x * blockDim.x + threadIdx.x;
const float2 *x = rdsc + (int64_t)stream * pitch;
const float2 *hs = hist + (int64_t)stream * (kRdsDecTaps - 1);
	if (blockIdx.x == 0 && threadIdx.x < kRdsDecTaps - 1) {      // roll the history
	   const int pos = M - (kRdsDecTaps - 1) + threadIdx.x;
	   new_hist [(int64_t)stream * (kRdsDecTaps - 1) + threadIdx.x] =
	         pos >= 0 ? x [pos] : hs [kRdsDecTaps - 1 + pos];
	}
	if (o >= nout) return;
//	first output of this call: smallest global input index g >= c0 with g = 7 mod 8
const int64_t g = ((c0 + 0) | 7) + 8 * (int64_t)o;
const int top = (int)(g - c0);
float2 acc = make_float2 (0.f, 0.f);
#pragma unroll
	for (int i = 0; i < kRdsDecTaps; i ++) {
	   const int j = top - i;
	   const float2 v = j >= 0 ? x [j] : hs [kRdsDecTaps - 1 + j];
	   const float2 t = cmul_rn (v, taps [i]);
	   acc.x = fadd (acc.x, t.x); acc.y = fadd (acc.y, t.y);
	}
	out [(int64_t)stream * out_pitch + o] = acc;
}

// ---- RDS symbol stage at 24 kHz, mode RDS_1 ---------------------------
//   Costas loop            includes/various/costas.h:21-33 (ctor args rds-decoder.cpp:40-41)
//   rdsDecoder_1::doDecode src/rds/rds-decoder-1.cpp:126-143: LowPassFIR (21), matched filter (43 taps),
//                          BandPassIIR (8 biquads) on the squared signal, bit at every top of that sine
// Everything here is a per-sample float recurrence (loop filter, ring-buffer FIRs summed in the
// reference's order, IIR): one lane per stream walks the call's 24 kHz samples.  The bits go to the
// host, where block synchronisation and group decoding stay (rds-blocksynchronizer.cpp, rds-groupdecoder.cpp).
constexpr int kRsyLp = 21, kRsyMatch = 43, kRsyQuads = 8, kRsyLanes = 32;
struct RdsSymState {               // Costas + rdsDecoder_1 members
	float   freq, phase;
	float   lp_buf [kRsyLp], mt_buf [kRsyMatch];
	int32_t lp_ip, mt_ip;
	float   m1 [kRsyQuads], m2 [kRsyQuads];
	float   last_sync_slope, last_sync, last_data;
	int32_t prev_bit;
};
struct RdsSymParams {
	float alpha, beta, freq_limit;  // 1/16, 0.02/16, 2 pi 10 / rate
	float match [kRsyMatch], lp [kRsyLp], bp [1 + 4 * kRsyQuads];
};

__global__ void __launch_bounds__ (kRsyLanes)
rds_symbol_kernel (const float2 *__restrict__ rds24, int64_t pitch, int32_t n, int32_t n_streams,
                   const RdsSymParams P, RdsSymState *__restrict__ state,
                   uint8_t *__restrict__ bits, int32_t cap_bits, int32_t *__restrict__ nbits) {
__shared__ float sLp [kRsyLp][kRsyLanes], sMt [kRsyMatch][kRsyLanes];
const int lane = threadIdx.x;
const int stream = blockIdx.x * kRsyLanes + lane;
	if (stream >= n_streams) return;
RdsSymState st = state [stream];
#pragma unroll
	for (int i = 0; i < kRsyLp; i ++) sLp [i][lane] = st.lp_buf [i];
#pragma unroll
	for (int i = 0; i < kRsyMatch; i ++) sMt [i][lane] = st.mt_buf [i];
int lp_ip = st.lp_ip, mt_ip = st.mt_ip;
float m1 [kRsyQuads], m2 [kRsyQuads];
#pragma unroll
	for (int i = 0; i < kRsyQuads; i ++) { m1 [i] = st.m1 [i]; m2 [i] = st.m2 [i]; }
const float2 *x = rds24 + (int64_t)stream * pitch;
uint8_t *out = bits + (int64_t)stream * cap_bits;
int nb = 0;
	for (int32_t t = 0; t < n; t ++) {
//	   Costas: r = z * exp (-i phase); the loop runs on re * im
	   float sn, cs;
	   sincosf (-st.phase, &sn, &cs);
	   const float2 r = cmul_rn (x [t], make_float2 (cs, sn));
	   const float err = fmul (r.x, r.y);
	   st.freq = fadd (st.freq, fmul (P.beta, err));
	   if (fabsf (st.freq) > P.freq_limit) st.freq = 0.f;
	   st.phase = fadd (st.phase, fadd (st.freq, fmul (P.alpha, err)));
	   st.phase = pi_constrain (st.phase);
//	   rdsFilter.Pass (float): newest first, fir-filters.h:95-108
	   sLp [lp_ip][lane] = r.x;
	   float v = 0.f;
	   { int idx = lp_ip;
#pragma unroll
	     for (int i = 0; i < kRsyLp; i ++) {
	        v = fadd (v, fmul (sLp [idx][lane], P.lp [i]));
	        idx = idx == 0 ? kRsyLp - 1 : idx - 1;
	     } }
	   lp_ip = lp_ip + 1 == kRsyLp ? 0 : lp_ip + 1;
//	   Match, rds-decoder-1.cpp:108-122
	   sMt [mt_ip][lane] = v;
	   float w = 0.f;
	   { int idx = mt_ip;
#pragma unroll
	     for (int i = 0; i < kRsyMatch; i ++) {
	        w = fadd (w, fmul (sMt [idx][lane], P.match [i]));
	        idx = idx == 0 ? kRsyMatch - 1 : idx - 1;
	     } }
	   mt_ip = mt_ip + 1 == kRsyMatch ? 0 : mt_ip + 1;
//	   sharpFilter on the squared signal; a bit at every top of the resulting sine, :130-142
	   const float mag = iir_pass<kRsyQuads> (P.bp, m1, m2, fmul (w, w));
	   const float slope = fsub (mag, st.last_sync);
	   st.last_sync = mag;
	   if (slope < 0.f && st.last_sync_slope >= 0.f) {
	      const int b = st.last_data >= 0.f ? 1 : 0;
	      if (nb < cap_bits) out [nb] = (uint8_t)(b ^ st.prev_bit);
	      nb ++;
	      st.prev_bit = b;
	   }
	   st.last_data = w;
	   st.last_sync_slope = slope;
	}
#pragma unroll
	for (int i = 0; i < kRsyLp; i ++) st.lp_buf [i] = sLp [i][lane];
#pragma unroll
	for (int i = 0; i < kRsyMatch; i ++) st.mt_buf [i] = sMt [i][lane];
	st.lp_ip = lp_ip; st.mt_ip = mt_ip;
#pragma unroll
	for (int i = 0; i < kRsyQuads; i ++) { st.m1 [i] = m1 [i]; st.m2 [i] = m2 [i]; }
	state [stream] = st;
	nbits [stream] = nb;
}

}	// namespace sdrjfm
