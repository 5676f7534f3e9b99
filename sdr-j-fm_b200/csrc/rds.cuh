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
constexpr int kRdsThreads = 512;       // 512 x 64 registers + 128 KB: the SM keeps room for a CTA of another kernel
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
constexpr int kRdsQpt = 4;                              // quads per thread and batch
constexpr int kRdsQbatches = kRdsNh / 4 / kRdsThreads / kRdsQpt;   // batches per pass (disjoint elements: no barrier between them)
static_assert ((kRdsNh & (kRdsNh - 1)) == 0 && (31 - __builtin_clz (kRdsNh)) % 2 == 0, "an even number of radix-2 stages");

// forward: decimation in frequency, natural order in -> bit-reversed order out
__device__ void fft_dif (float2 *a, const float2 *__restrict__ tws) {
	for (int h = kRdsNh / 2; h >= 2; h >>= 2) {           // stages `half = h` and `half = h / 2`
	   const int hh = h >> 1;
	   for (int bt = 0; bt < kRdsQbatches; bt ++) {
	   float2 x0 [kRdsQpt], x1 [kRdsQpt], x2 [kRdsQpt], x3 [kRdsQpt], w [kRdsQpt], w2 [kRdsQpt];
	   int idx [kRdsQpt];
#pragma unroll
	   for (int q = 0; q < kRdsQpt; q ++) {
	      const int b = threadIdx.x + (bt * kRdsQpt + q) * kRdsThreads;
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
	   }
	   __syncthreads ();
	}
}
// inverse: decimation in time with conjugate twiddles, bit-reversed order in -> natural order out
__device__ void ifft_dit (float2 *a, const float2 *__restrict__ tws) {
	for (int h = 1; h <= kRdsNh / 4; h <<= 2) {           // stages `half = h` and `half = 2 h`
	   for (int bt = 0; bt < kRdsQbatches; bt ++) {
	   float2 x0 [kRdsQpt], x1 [kRdsQpt], x2 [kRdsQpt], x3 [kRdsQpt], w [kRdsQpt], w2 [kRdsQpt];
	   int idx [kRdsQpt];
#pragma unroll
	   for (int q = 0; q < kRdsQpt; q ++) {
	      const int b = threadIdx.x + (bt * kRdsQpt + q) * kRdsThreads;
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
__global__ void __launch_bounds__ (kRdsThreads, 2)      // <= 64 registers (shared memory allows one CTA per SM)
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
//	(in two halves of the index range to keep the register count at 64: a half reads entries
//	383 + m, 384 + m and, after a barrier, writes entries m of ITS range only; the second half's
//	sources lie above everything the first half wrote)
constexpr int kZ = kRdsNh / kRdsThreads / 2;
	for (int part = 0; part < 2; part ++) {
	   float2 z [kZ];
#pragma unroll
	   for (int q = 0; q < kZ; q ++) {
	      const int m = tid + (part * kZ + q) * kRdsThreads;              // z[m] = (bp[2m], bp[2m+1])
	      float2 v = make_float2 (0.f, 0.f);
	      if (2 * m < kRdsBlock) v = make_float2 (fa [(kRdsTaps - 2) / 2 + m].y, fa [kRdsTaps / 2 + m].x);
	      z [q] = v;
	   }
	   __syncthreads ();
#pragma unroll
	   for (int q = 0; q < kZ; q ++) {
	      const int m = tid + (part * kZ + q) * kRdsThreads;
	      fa [m] = z [q];
	      if (2 * m < kRdsBlock) reinterpret_cast<float2 *>(bpo) [m] = z [q];
	   }
	   __syncthreads ();
	}
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
                                  float2 *__restrict__ out, int64_t out_pitch, int32_t nout,
                                  float2 *__restrict__ out2, int64_t out2_pitch) {     // out2: private copy for the symbol stage
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
	if (out2) out2 [(int64_t)stream * out2_pitch + o] = acc;
}

// ---- RDS symbol stage at 24 kHz, mode RDS_1 ---------------------------
//   Costas loop            includes/various/costas.h:21-33 (ctor args rds-decoder.cpp:40-41)
//   rdsDecoder_1::doDecode src/rds/rds-decoder-1.cpp:126-143: LowPassFIR (21), matched filter (43 taps),
//                          BandPassIIR (8 biquads) on the squared signal, bit at every top of that sine
// Only the Costas loop (a non-linear feedback loop) and the IIR are recurrences:
//   rds_costas_kernel   one lane per stream walks the call's samples (the loop filter)
//   rds_fir_kernel      the two ring-buffer FIRs, one thread per output, summed in the reference's order
//   rds_bits_kernel     the 8 biquads as a systolic pipeline over 8 lanes (lane q runs biquad q on the
//                       sample lane q-1 finished one step earlier: every sample still sees the
//                       reference's operations in the reference's order), slope detector on the last lane
// The bits go to the host, where block synchronisation and group decoding stay
// (rds-blocksynchronizer.cpp, rds-groupdecoder.cpp).
constexpr int kRsyLp = 21, kRsyMatch = 43, kRsyQuads = 8, kRsyLanes = 32;
struct RdsSymState {               // Costas + rdsDecoder_1 members
	float   freq, phase;
	float   hist_c [kRsyLp - 1];    // the last 20 Costas outputs (oldest first) = rdsFilter's buffer
	float   hist_v [kRsyMatch - 1]; // the last 42 low-pass outputs = rdsBuffer
	float   m1 [kRsyQuads], m2 [kRsyQuads];
	float   last_sync_slope, last_sync, last_data;
	int32_t prev_bit;
};
struct RdsSymParams {
	float alpha, beta, freq_limit;  // 1/16, 0.02/16, 2 pi 10 / rate
	float match [kRsyMatch], lp [kRsyLp], bp [1 + 4 * kRsyQuads];
};

__global__ void __launch_bounds__ (kRsyLanes)
rds_costas_kernel (const float2 *__restrict__ rds24, int64_t pitch, int32_t n, int32_t n_streams,
                   const RdsSymParams P, RdsSymState *__restrict__ state, float *__restrict__ cbuf, int64_t cpitch,
                   float *__restrict__ qbuf, int64_t qpitch) {          // qbuf (optional): the imaginary parts, for the RDS_DEMOD scope stream
const int stream = blockIdx.x * kRsyLanes + threadIdx.x;
	if (stream >= n_streams) return;
float freq = state [stream].freq, phase = state [stream].phase;
const float2 *x = rds24 + (int64_t)stream * pitch;
float *c = cbuf + (int64_t)stream * cpitch;
float *qi = qbuf ? qbuf + (int64_t)stream * qpitch : nullptr;
//	the samples are fetched one chunk ahead, so the loop never waits on global memory
constexpr int CH = 8;
float2 cur [CH], nxt [CH];
#pragma unroll
	for (int k = 0; k < CH; k ++) cur [k] = k < n ? x [k] : make_float2 (0.f, 0.f);
	for (int32_t t0 = 0; t0 < n; t0 += CH) {
#pragma unroll
	   for (int k = 0; k < CH; k ++) nxt [k] = t0 + CH + k < n ? x [t0 + CH + k] : make_float2 (0.f, 0.f);
#pragma unroll
	   for (int k = 0; k < CH; k ++) {
	      if (t0 + k < n) {
//	         r = z * exp (-i phase); the loop runs on re * im  (costas.h:21-33)
	         float sn, cs;
	         sincosf (-phase, &sn, &cs);
	         const float2 r = cmul_rn (cur [k], make_float2 (cs, sn));
	         const float err = fmul (r.x, r.y);
	         freq = fadd (freq, fmul (P.beta, err));
	         if (fabsf (freq) > P.freq_limit) freq = 0.f;
	         phase = pi_constrain (fadd (phase, fadd (freq, fmul (P.alpha, err))));
	         c [t0 + k] = r.x;
	         if (qi) qi [t0 + k] = r.y;
	      }
	   }
#pragma unroll
	   for (int k = 0; k < CH; k ++) cur [k] = nxt [k];
	}
	state [stream].freq = freq; state [stream].phase = phase;
}

// out[t] = sum_{i<NT} in[t - i] * taps[i], i ascending (Basic_FIR::Pass (float), fir-filters.h:95-108 and
// rdsDecoder_1::Match, rds-decoder-1.cpp:108-122); in[t] for t < 0 comes from the carried history
template <int NT, bool MATCH>
__global__ void rds_fir_kernel (const float *__restrict__ in, float *__restrict__ out, int64_t pitch, int32_t n,
                                const RdsSymParams P, const RdsSymState *__restrict__ state) {
const int stream = blockIdx.y;
const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
const float *x = in + (int64_t)stream * pitch;
const float *hist = MATCH ? state [stream].hist_v : state [stream].hist_c;     // NT - 1 entries, oldest first
float acc = 0.f;
#pragma unroll
	for (int i = 0; i < NT; i ++) {
	   const int k = t - i;
	   const float v = k >= 0 ? x [k] : hist [NT - 1 + k];
	   acc = fadd (acc, fmul (v, MATCH ? P.match [i] : P.lp [i]));
	}
	out [(int64_t)stream * pitch + t] = acc;
}

// 8 lanes per stream; w = matched-filter output of this call
__global__ void __launch_bounds__ (kRsyLanes)
rds_bits_kernel (const float *__restrict__ wbuf, int64_t pitch, int32_t n, int32_t n_streams,
                 const RdsSymParams P, RdsSymState *__restrict__ state,
                 uint8_t *__restrict__ bits, int32_t cap_bits, int32_t *__restrict__ nbits) {
const int q = threadIdx.x & (kRsyQuads - 1);
const int stream = blockIdx.x * (kRsyLanes / kRsyQuads) + (threadIdx.x >> 3);
const bool live = stream < n_streams;
const int sidx = live ? stream : 0;
RdsSymState &st = state [sidx];
const float A1 = P.bp [1 + 4 * q], A2 = P.bp [2 + 4 * q], B1 = P.bp [3 + 4 * q], B2 = P.bp [4 + 4 * q];
float m1 = st.m1 [q], m2 = st.m2 [q];
float lastSlope = st.last_sync_slope, lastSync = st.last_sync, lastData = st.last_data;
int prev = st.prev_bit, nb = 0;
const float *w = wbuf + (int64_t)sidx * pitch;
uint8_t *out = bits + (int64_t)sidx * cap_bits;
float o = 0.f;
//	every lane fetches the matched-filter output of ITS sample (t = s - q) one chunk ahead
constexpr int CH = 8;
const int32_t steps = n + kRsyQuads - 1;
float cur [CH], nxt [CH];
#pragma unroll
	for (int k = 0; k < CH; k ++) { const int32_t t = k - q; cur [k] = (t >= 0 && t < n) ? w [t] : 0.f; }
	for (int32_t s0 = 0; s0 < steps; s0 += CH) {
#pragma unroll
	   for (int k = 0; k < CH; k ++) { const int32_t t = s0 + CH + k - q; nxt [k] = (t >= 0 && t < n) ? w [t] : 0.f; }
#pragma unroll
	   for (int k = 0; k < CH; k ++) {
	      const int32_t s = s0 + k;
	      const float fromPrev = __shfl_up_sync (0xffffffffu, o, 1, kRsyQuads);    // what lane q-1 produced last step
	      const int32_t t = s - q;                                                 // the sample this lane works on now
	      const bool act = t >= 0 && t < n;
	      const float wt = cur [k];
	      const float in = q == 0 ? fmul (fmul (wt, wt), P.bp [0]) : fromPrev;     // o = v * gain, iir-filters.h:93
	      if (act) {
	         const float ww = fsub (fsub (in, fmul (m1, B1)), fmul (m2, B2));
	         o = fadd (fadd (ww, fmul (m1, A1)), fmul (m2, A2));
	         m2 = m1; m1 = ww;
	      }
	      if (q == kRsyQuads - 1 && act && live) {
//	         rdsMag = o: a bit at every top of the sine, rds-decoder-1.cpp:130-142
	         const float slope = fsub (o, lastSync);
	         lastSync = o;
	         if (slope < 0.f && lastSlope >= 0.f) {
	            const int b = lastData >= 0.f ? 1 : 0;
	            if (nb < cap_bits) out [nb] = (uint8_t)(b ^ prev);
	            nb ++;
	            prev = b;
	         }
	         lastData = wt;
	         lastSlope = slope;
	      }
	   }
#pragma unroll
	   for (int k = 0; k < CH; k ++) cur [k] = nxt [k];
	}
	if (live) {
	   st.m1 [q] = m1; st.m2 [q] = m2;
	   if (q == kRsyQuads - 1) {
	      st.last_sync_slope = lastSlope; st.last_sync = lastSync; st.last_data = lastData; st.prev_bit = prev;
	      nbits [stream] = nb;
	   }
	}
}

// carry the FIR histories to the next call: the last 20 Costas outputs and the last 42 low-pass outputs
__global__ void rds_sym_roll_kernel (const float *__restrict__ cbuf, const float *__restrict__ vbuf, int64_t pitch,
                                     int32_t n, RdsSymState *__restrict__ state) {
const int stream = blockIdx.x, i = threadIdx.x;
RdsSymState &st = state [stream];
float a = 0.f, b = 0.f;
	if (i < kRsyLp - 1) { const int k = n - (kRsyLp - 1) + i; a = k >= 0 ? cbuf [(int64_t)stream * pitch + k] : st.hist_c [kRsyLp - 1 + k]; }
	if (i < kRsyMatch - 1) { const int k = n - (kRsyMatch - 1) + i; b = k >= 0 ? vbuf [(int64_t)stream * pitch + k] : st.hist_v [kRsyMatch - 1 + k]; }
	__syncthreads ();
	if (i < kRsyLp - 1) st.hist_c [i] = a;
	if (i < kRsyMatch - 1) st.hist_v [i] = b;
}

// ---- symbol stage of mode RDS_2: rdsDecoder_2::doDecode (src/rds/rds-decoder-2.cpp:96-158) ------------------
//   doMatchFiltering   45-tap root-raised-cosine matched filter on the complex 24 kHz baseband (:79-93)
//   AGC                output = input * curGain; curGain += 2e-3 (0.38 - |output|)         (includes/various/agc.h)
//   process_sample     Mueller & Mueller timing recovery on hard decisions: one symbol per ~20.2 samples (:115-158)
//   Costas             (rate, 1.0, 0.02, 10 Hz) on the SYMBOLS (includes/various/costas.h), then the differential bit
// The matched filter is a FIR (thread per output, taps accumulated in the reference's order); everything behind
// it is a per-sample recurrence with data-dependent decimation: one lane per stream, statement by statement.
constexpr int kRs2Taps = 45;
struct Rds2State {                  // rdsDecoder_2 members
	float2  hist [kRs2Taps - 1];     // the last 44 inputs of the matched filter (oldest first)
	float   gain;                    // AGC::curGain
	float2  sb [3];                  // sampleBuffer
	float   mu;                      // mMu
	int32_t skip, count;             // skipNrSamples, sampleCount
	float   freq, phase;             // Costas
	int32_t prev_bit, started;
};
struct Rds2Params {
	float taps [kRs2Taps];
	float agc_rate, agc_ref;         // 2e-3, 0.38 (rds-decoder-2.cpp:46)
	float sps, mm_alpha;             // rate / 1187.5, 0.01
	float c_alpha, c_beta, freq_limit;
};

__global__ void __launch_bounds__ (128)
rds2_match_kernel (const float2 *__restrict__ in, int64_t pitch, int32_t n, const Rds2Params P,
                   const Rds2State *__restrict__ state, float2 *__restrict__ out) {
const int stream = blockIdx.y;
const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
const float2 *x = in + (int64_t)stream * pitch;
const float2 *hs = state [stream].hist;
float2 tmp = make_float2 (0.f, 0.f);
#pragma unroll 5
	for (int i = 0; i < kRs2Taps; i ++) {                       // tmp += buf [newest - i] * kernel [i]
	   const int k = t - i;
	   const float2 v = k >= 0 ? x [k] : hs [kRs2Taps - 1 + k];
	   tmp.x = fadd (tmp.x, fmul (v.x, P.taps [i])); tmp.y = fadd (tmp.y, fmul (v.y, P.taps [i]));
	}
	out [(int64_t)stream * pitch + t] = tmp;
}

__global__ void __launch_bounds__ (kRsyLanes)
rds2_seq_kernel (const float2 *__restrict__ in, const float2 *__restrict__ mf, int64_t pitch, int32_t n, int32_t n_streams,
                 const Rds2Params P, Rds2State *__restrict__ state,
                 uint8_t *__restrict__ bits, int32_t cap_bits, int32_t *__restrict__ nbits) {
const int stream = blockIdx.x * kRsyLanes + threadIdx.x;
	if (stream >= n_streams) return;
Rds2State &st = state [stream];
const float2 *m = mf + (int64_t)stream * pitch;
uint8_t *out = bits + (int64_t)stream * cap_bits;
float gain = st.gain, mu = st.mu, freq = st.freq, phase = st.phase;
float2 s0 = st.sb [0], s1 = st.sb [1], s2 = st.sb [2];
int skip = st.skip, count = st.count, prev = st.prev_bit, nb = 0;
	for (int32_t t = 0; t < n; t ++) {
	   float2 v = m [t];
	   v = make_float2 (fmul (v.x, gain), fmul (v.y, gain));                                         // AGC::process_sample
	   const float mag = (float)sqrt ((double)v.x * (double)v.x + (double)v.y * (double)v.y);        // std::abs (complex<float>)
	   gain = fadd (gain, fmul (P.agc_rate, fsub (P.agc_ref, mag)));
	   s0 = s1; s1 = s2; s2 = v;                                                                       // :117-119
	   if (++ count >= skip) {
	      const float2 r0 = make_float2 (s0.x > 0.0f ? 1.0f : -1.0f, s0.y > 0.0f ? 1.0f : -1.0f);
	      const float2 r1 = make_float2 (s1.x > 0.0f ? 1.0f : -1.0f, s1.y > 0.0f ? 1.0f : -1.0f);
	      const float2 r2 = make_float2 (s2.x > 0.0f ? 1.0f : -1.0f, s2.y > 0.0f ? 1.0f : -1.0f);
	      const float x = fadd (fmul (fsub (r2.x, r0.x), s1.x), fmul (fsub (r2.y, r0.y), s1.y));      // :130-134
	      const float y = fadd (fmul (fsub (s2.x, s0.x), r1.x), fmul (fsub (s2.y, s0.y), r1.y));      // :135-139
	      const float mm_val = fsub (y, x);
	      mu = fadd (mu, fadd (P.sps, fmul (P.mm_alpha, mm_val)));                                     // :143
	      skip = (int32_t)mu;
	      mu = fsub (mu, (float)skip);
	      count = 0;
//	      Costas on the symbol (costas.h:21-33), then the differential bit (:104-108)
	      float sn, cs;
	      sincosf (-phase, &sn, &cs);
	      const float2 r = cmul_rn (s2, make_float2 (cs, sn));
	      const float err = fmul (r.x, r.y);
	      freq = fadd (freq, fmul (P.c_beta, err));
	      if (fabsf (freq) > P.freq_limit) freq = 0.f;
	      phase = pi_constrain (fadd (phase, fadd (freq, fmul (P.c_alpha, err))));
	      const int b = r.x >= 0.f ? 1 : 0;
	      if (nb < cap_bits) out [nb] = (uint8_t)(b ^ prev);
	      nb ++;
	      prev = b;
	   }
	}
	st.gain = gain; st.mu = mu; st.freq = freq; st.phase = phase;
	st.sb [0] = s0; st.sb [1] = s1; st.sb [2] = s2;
	st.skip = skip; st.count = count; st.prev_bit = prev;
	nbits [stream] = nb;
//	matched-filter history for the next call: the last 44 inputs of (hist | in)
const float2 *x = in + (int64_t)stream * pitch;
float2 keep [kRs2Taps - 1];
	for (int i = 0; i < kRs2Taps - 1; i ++) { const int k = n - (kRs2Taps - 1) + i; keep [i] = k >= 0 ? x [k] : st.hist [kRs2Taps - 1 + k]; }
	for (int i = 0; i < kRs2Taps - 1; i ++) st.hist [i] = keep [i];
}

// ---- symbol stage of mode RDS_3: Costas (shared with mode 1) + rdsDecoder_3 (src/rds/rds-decoder-3.cpp:83-118) ----
//   rdsFilter.Pass          the 21-tap low-pass of mode 1, kept in a ring of 21 = ceil (24000 / 1187.5) values (:88-89)
//   synchronizeOnBitClk     folds that ring over half a bit-clock period and puts the clock phase on its rising edge (:121-158);
//                           it runs on the first sample and whenever the block synchroniser has counted more than 3
//                           sync errors (:91-96) — which is why the synchroniser (src/rds/rds-blocksynchronizer.cpp) and
//                           the group it fills (src/rds/rds-group.cpp) live on the device in this mode
//   bit clock               sine of a phase advancing 2 pi 1187.5 / 24000 per sample (the reference's table SinCos (24000));
//                           the Costas output is integrated against it and a bit is taken at every rising edge (:98-115)
// Integer state machine and float recurrence alike: one lane per stream, statement by statement.
constexpr int kRs3Sym = 21;                     // symbolCeiling; symbolFloor = 20
constexpr int kRs3BlockBits = 26, kRs3CrcBits = 10;
struct Rds3Sync {                  // rdsBlockSynchronizer members + RDSGroup::rdsBlocks
	uint32_t bitstream;
	int32_t  synced, block, bits_in_block;
	uint16_t sync_errors, crc_errors, bits_processed, bit_errors;
	uint16_t grp [4];
};
struct Rds3State {                 // rdsDecoder_3 members
	float    integ, clk_phase, prev_clk;
	int32_t  prev_bit, resync;
	Rds3Sync sy;
};
struct Rds3Params {
	const float *sin_tab;           // (float) sin (2 pi i / 24000), i < 24000 (SinCos::SinCos (Rate), sincos.cpp:35-44)
	double   C;                     // Rate / (2 pi)
	int32_t  rate;
	float    omega;                 // omegaRDS
	int8_t   kmap [kRs3Sym];        // correlationVector slot of ring element i (a constant of the clock, :133-147)
};
enum { kRs3Waiting = 0, kRs3Buffering, kRs3NoSync, kRs3NoCrc, kRs3Group };

// SinCos::getSin (sincos.cpp:53-58, 76-80)
__device__ __forceinline__ float rs3_sin (const Rds3Params &P, float phase) {
const bool neg = phase < 0.f;
const float a = neg ? -phase : phase;
const int32_t i = ((int32_t)((double)a * P.C)) % P.rate;
const float v = P.sin_tab [i];
	return neg ? -v : v;
}
// remainder of (bits ^ offset) x^10 modulo the RDS generator polynomial, fed msb first as the reference's shift
// register does (getSyndrome, rds-blocksynchronizer.cpp:109-125): zero for an error-free block
__device__ __forceinline__ uint32_t rs3_syndrome (uint32_t bits, uint32_t offset) {
const uint32_t block = bits ^ offset;
uint32_t reg = 0;
	for (int k = kRs3BlockBits - 1; k >= 0; k --) {
	   const bool top = (reg >> (kRs3CrcBits - 1)) & 1u;
	   reg <<= 1;
	   if (top) reg ^= 0x5B9u;
	   if ((block >> k) & 1u) reg ^= 0x31Bu;
	}
	return reg;
}
__device__ __forceinline__ uint32_t rs3_offset (int block, bool typeB) {
	return block == 0 ? 0xFCu : block == 1 ? 0x198u : block == 2 ? (typeB ? 0x350u : 0x168u) : 0x1B4u;
}
// rdsBlockSynchronizer::pushBit (:215-335) with decodeBlock (:127-163) and doMeggit (:171-191) folded in
__device__ __forceinline__ int rs3_push_bit (Rds3Sync &s, bool b) {
const bool typeB = (s.grp [1] >> 11) & 1u;                          // RDSGroup::isTypeBGroup
	s.bitstream = (s.bitstream << 1) | (b ? 1u : 0u);
	if (s.synced) {
	   s.bits_in_block = (s.bits_in_block + 1) & 0xffff;
	   if (s.bits_in_block < kRs3BlockBits) return kRs3Buffering;
	   s.bits_in_block = 0;
	   const uint32_t syn = rs3_syndrome (s.bitstream, rs3_offset (s.block, typeB));
	   s.bits_processed += 16;
	   if (syn != 0) {
//	      the single-burst correction runs on the bit stream; its verdict is not read back (:140-147)
	      uint32_t y = syn, mask = 1u << (kRs3BlockBits - 1);
	      for (int i = 0; i < 16; i ++) {
	         if (y & 0x200u) {
	            if ((y & 0x1fu) == 0) { s.bitstream ^= mask; s.bit_errors ++; }
	            else y ^= 0x5B9u;
	         }
	         y <<= 1; mask >>= 1;
	      }
	      s.bit_errors += 16;
	   }
	   if (s.bits_processed >= 4000) { s.bit_errors = 0; s.bits_processed = 0; }
	   if (syn != 0) { s.crc_errors ++; return kRs3NoCrc; }
	   s.grp [s.block] = (uint16_t)(s.bitstream >> kRs3CrcBits);
	   const int res = s.block == 3 ? kRs3Group : kRs3Buffering;
	   s.block = (s.block + 1) & 3;
	   return res;
	}
	if (s.block == 0) {                                               // shifting until a valid block A passes
	   if (rs3_syndrome (s.bitstream & 0x3FFFFFFu, rs3_offset (0, typeB)) != 0) return kRs3Waiting;
	   s.grp [0] = (uint16_t)(s.bitstream >> kRs3CrcBits);
	   s.bits_in_block = 0;
	   s.block = 1;
	   return kRs3Buffering;
	}
	if (s.bits_in_block < kRs3BlockBits - 1) { s.bits_in_block ++; return kRs3Buffering; }
	s.bits_in_block = 0;
	if (rs3_syndrome (s.bitstream, rs3_offset (s.block, typeB)) != 0) { s.sync_errors ++; return kRs3NoSync; }
	s.grp [s.block] = (uint16_t)(s.bitstream >> kRs3CrcBits);
	if (s.block < 2) { s.block ++; return kRs3Buffering; }
	s.synced = 1;                                                      // blocks A, B, C in a row: synchronised
	const int res = s.block == 3 ? kRs3Group : kRs3Buffering;
	s.block = (s.block + 1) & 3;
	return res;
}
__device__ __forceinline__ void rs3_resync (Rds3Sync &s) { s.block = 0; s.synced = 0; s.bits_in_block = 0; }

// cbuf: Costas outputs (real part) of this call; vbuf: their low-pass (rds_fir_kernel<kRsyLp>); hist_v: the low-pass outputs
// before this call (RdsSymState::hist_v, newest last).  groups: [S][cap_groups][4]; stat: [S][4] = synchronised, bit-clock
// re-synchronisations in this call, sync errors, crc errors
__global__ void __launch_bounds__ (kRsyLanes)
rds3_seq_kernel (const float *__restrict__ cbuf, const float *__restrict__ vbuf, int64_t pitch, int32_t n, int32_t n_streams,
                 const Rds3Params P, const RdsSymState *__restrict__ sym, Rds3State *__restrict__ state,
                 uint8_t *__restrict__ bits, int32_t cap_bits, int32_t *__restrict__ nbits,
                 uint16_t *__restrict__ groups, int32_t cap_groups, int32_t *__restrict__ ngroups, int32_t *__restrict__ stat) {
const int stream = blockIdx.x * kRsyLanes + threadIdx.x;
	if (stream >= n_streams) return;
Rds3State st = state [stream];
const float *c = cbuf + (int64_t)stream * pitch;
const float *v = vbuf + (int64_t)stream * pitch;
const float *hv = sym [stream].hist_v;                                // kRsyMatch - 1 entries
uint8_t *out = bits + (int64_t)stream * cap_bits;
uint16_t *gout = groups + (int64_t)stream * cap_groups * 4;
int nb = 0, ng = 0, nrs = 0;
	for (int32_t t = 0; t < n; t ++) {
	   if (st.resync || st.sy.sync_errors > 3) {
//	      synchronizeOnBitClk on the last 21 low-pass outputs, oldest first
	      float corr [kRs3Sym];
#pragma unroll
	      for (int i = 0; i < kRs3Sym; i ++) corr [i] = 0.f;
	      for (int i = 0; i < kRs3Sym; i ++) {
	         const int k = t - (kRs3Sym - 1) + i;
	         const float x = k >= 0 ? v [k] : hv [kRsyMatch - 1 + k];
	         corr [P.kmap [i]] = fadd (corr [P.kmap [i]], x);
	      }
	      int iMin = 0;
	      while (iMin < kRs3Sym - 1) { const float q = corr [iMin ++]; if (!(q > 0.f)) break; }
	      while (iMin < kRs3Sym - 1) { const float q = corr [iMin ++]; if (!(q < 0.f)) break; }
	      float ph = (float)fmod ((double)fmul (-P.omega, (float)(iMin - 1)), 2 * M_PI);
	      while (ph < 0.f) ph = (float)((double)ph + 2 * M_PI);
	      st.clk_phase = ph;
	      rs3_resync (st.sy);
	      st.sy.sync_errors = 0;
	      st.resync = 0;
	      nrs ++;
	   }
	   const float clk = rs3_sin (P, st.clk_phase);
	   st.integ = fadd (st.integ, fmul (clk, c [t]));
	   if (st.prev_clk <= 0.f && clk > 0.f) {                          // rising edge: look at the integrator
	      const int theBit = st.integ >= 0.f ? 1 : 0;
	      const int d = theBit ^ st.prev_bit;
	      st.integ = 0.f;
	      st.prev_bit = theBit;
	      if (nb < cap_bits) out [nb] = (uint8_t)d;
	      nb ++;
//	      rdsDecoder::processBit (rds-decoder.cpp:104-131)
	      const int r = rs3_push_bit (st.sy, d != 0);
	      if (r == kRs3NoSync || r == kRs3NoCrc) rs3_resync (st.sy);
	      else if (r == kRs3Group) {
	         if (ng < cap_groups) { for (int b = 0; b < 4; b ++) gout [4 * ng + b] = st.sy.grp [b]; }
	         ng ++;
	         for (int b = 0; b < 4; b ++) st.sy.grp [b] = 0;            // my_rdsGroup.clear ()
	      }
	   }
	   st.prev_clk = clk;
	   st.clk_phase = (float)fmod ((double)fadd (st.clk_phase, P.omega), 2 * M_PI);
	}
	state [stream] = st;
	nbits [stream] = nb;
	ngroups [stream] = ng;
	stat [4 * stream] = st.sy.synced; stat [4 * stream + 1] = nrs;
	stat [4 * stream + 2] = st.sy.sync_errors; stat [4 * stream + 3] = st.sy.crc_errors;
}

}	// namespace sdrjfm
