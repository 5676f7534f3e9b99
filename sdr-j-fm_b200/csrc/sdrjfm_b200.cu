// C ABI (include/sdrjfm_b200.h) of the B200 FM path.
//
// A handle owns n_streams independent IQ streams.  They are split over a few LANES
// (lane_impl.cuh): each lane runs the complete kernel sequence for its group of streams on
// its own CUDA stream, so the latency-bound per-stream kernels (discriminator, pilot PLL,
// PSS, RDS block transforms) of one lane overlap the kernels of the others.  Streams never
// interact, so the split is invisible in the results.
// There is NO CPU fallback: without an sm_100 device sdrjfm_create fails with
// SDRJFM_ERR_NO_DEVICE.
#include "lane_impl.cuh"
#include <cstdlib>
#include <mutex>

// A handle works on up to 15 CUDA streams (per lane: its own, the RDS side stream, the pilot stream of the next
// time slice; the caller's stream; two copy streams).  The driver maps streams onto 8 hardware queues by default, which makes independent
// streams wait for each other; when this library is loaded before the process creates its CUDA context
// (and the host did not choose a value itself) it asks for 32.  Measured: 4.36 -> 4.30 ms per bench step.
// This writes one variable of the host process's environment at load time and never overrides a value
// the host set; SDRJFM_NO_ENV_HINT=1 switches it off (documented in include/sdrjfm_b200.h).
namespace { struct QueueHint { QueueHint () {
	const char *off = getenv ("SDRJFM_NO_ENV_HINT");
	if (!(off && off [0] == '1')) setenv ("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); } } g_queue_hint; }

// The tap sets live in __constant__ banks shared by every handle on a device (lane_impl.cuh), and a
// call's launches must see the image that was ensured at its start: calls into handles on the SAME
// device are serialised on the host side (enqueueing only; the GPU work of different handles still
// overlaps).  Recursive: the host entry points call the device ones.
static std::recursive_mutex g_dev_mutex [kMaxDevices];

struct sdrjfm_handle {
	sdrjfm_config cfg;
	std::vector<Lane *> lanes;
	std::vector<int32_t> first;             // first stream of each lane; first [lanes] = n_streams
	cudaStream_t stream = nullptr;          // the stream callers time / order against
	bool own_stream = false;
	cudaStream_t copy_stream = nullptr;     // H2D of the next time slice while the current one computes
	cudaStream_t out_stream = nullptr;      // D2H of a finished slice's outputs while the next ones compute
	cudaEvent_t  ev_fork = nullptr, ev_h2d [2] = { nullptr, nullptr }, ev_done [2] = { nullptr, nullptr };
	std::vector<cudaEvent_t> ev_join;
	int64_t cap_in = 0, cap_audio = 0, cap_rds = 0;
	float2 *d_in [2] = { nullptr, nullptr }; // host-call staging [S][cap_in] (second one on first pipelined call)
	float2 *d_audio = nullptr, *d_rds24 = nullptr;
	bool poisoned = false;                  // a call failed after state had moved: streams are desynchronised
	std::string err;
};
#define LOCK_DEV(h) std::lock_guard<std::recursive_mutex> lock_ (g_dev_mutex [(h) -> cfg.device])
#define CHECK_POISON(h) do { if ((h) -> poisoned) { (h) -> err = "handle poisoned: an earlier call failed after stream state had moved; destroy and re-create it"; return SDRJFM_ERR_CUDA; } } while (0)

#define HK(call)                                                                          \
	do { cudaError_t e_ = (call); if (e_ != cudaSuccess) {                                \
	   char b_ [256]; snprintf (b_, sizeof b_, "%s failed: %s (%s:%d)", #call,           \
	                           cudaGetErrorString (e_), __FILE__, __LINE__);              \
	   h -> err = b_; return SDRJFM_ERR_CUDA; } } while (0)

static int lane_of (const sdrjfm_handle *h, int32_t stream) {
	for (size_t i = 0; i < h -> lanes.size (); i ++)
	   if (stream >= h -> first [i] && stream < h -> first [i + 1]) return (int)i;
	return -1;
}

// applies one lane-level call to every lane, stopping at (and reporting) the first failure
template <typename F> static int for_lanes (sdrjfm_handle *h, F f) {
	if (!h) return SDRJFM_ERR_ARG;
	LOCK_DEV (h);
	CHECK_POISON (h);
	for (size_t i = 0; i < h -> lanes.size (); i ++) {
	   Lane *l = h -> lanes [i];
	   const int rc = f (l);
	   if (rc != SDRJFM_OK) {
	      h -> err = l -> err;
	      if (i > 0) h -> poisoned = true;    // earlier lanes already took the setting: the lanes disagree now
	      return rc;
	   }
	}
	return SDRJFM_OK;
}

extern "C" int sdrjfm_destroy (sdrjfm_handle *h);

// fork: every lane waits for what is queued on the handle's stream, runs its group, and the
// handle's stream waits for every lane (join).  With one lane the two streams are the same.
template <typename F> static int fork_join (sdrjfm_handle *h, F f) {
const size_t nl = h -> lanes.size ();
	if (nl > 1) HK (cudaEventRecord (h -> ev_fork, h -> stream));
	for (size_t i = 0; i < nl; i ++) {
	   Lane *l = h -> lanes [i];
	   if (nl > 1) HK (cudaStreamWaitEvent (l -> stream, h -> ev_fork, 0));
	   const int rc = f (l, h -> first [i]);
	   if (rc != SDRJFM_OK) {
	      h -> err = l -> err;
//	      argument errors are detected by every lane before it touches anything; any other failure, or
//	      one after an earlier lane ran, leaves the lanes' streams out of step
	      if (i > 0 || (rc != SDRJFM_ERR_ARG && rc != SDRJFM_ERR_CAPACITY && rc != SDRJFM_ERR_UNSUPPORTED)) {
	         cudaDeviceSynchronize ();
	         h -> poisoned = true;
	      }
	      return rc;
	   }
	   if (nl > 1) {
	      HK (cudaEventRecord (h -> ev_join [i], l -> stream));
	      HK (cudaStreamWaitEvent (h -> stream, h -> ev_join [i], 0));
	   }
	}
	return SDRJFM_OK;
}

extern "C" {

const char *sdrjfm_version (void) { return "sdrjfm_b200 0.2 (sm_100a)"; }

const char *sdrjfm_last_error (const sdrjfm_handle *h) {
	return h ? h -> err.c_str () : g_create_error.c_str ();
}

sdrjfm_handle *sdrjfm_create (const sdrjfm_config *cfg, int *status) {
int dummy; if (!status) status = &dummy;
	*status = SDRJFM_ERR_ARG;
	if (!cfg || cfg -> n_streams < 1 || cfg -> max_samples_per_call < 1) {
	   g_create_error = "bad config"; return nullptr;
	}
//	lanes: with K3 of time slice j + 1 running beside K4-K6 of slice j inside every lane (lane_impl.cuh), two lanes
//	are enough to keep a B200 busy from 128 streams on, and one below (measured, 256 streams x 0.5 s: 1 / 2 / 4 lanes
//	4.28 / 4.28 / 4.38 ms; 32 streams: every extra lane only adds launches).  SDRJFM_LANES overrides.
int nl = cfg -> n_streams >= 128 ? 2 : 1;
	{ const char *env = getenv ("SDRJFM_LANES"); if (env && atoi (env) > 0) nl = std::min (atoi (env), cfg -> n_streams); }
	if (cfg -> device < 0 || cfg -> device >= kMaxDevices) { g_create_error = "bad device ordinal"; return nullptr; }
std::lock_guard<std::recursive_mutex> lock_ (g_dev_mutex [cfg -> device]);
sdrjfm_handle *h = new sdrjfm_handle ();
	h -> cfg = *cfg;
	for (int i = 0; i < nl; i ++) {
	   const int32_t lo = (int32_t)((int64_t)cfg -> n_streams * i / nl);
	   const int32_t hi = (int32_t)((int64_t)cfg -> n_streams * (i + 1) / nl);
	   sdrjfm_config c = *cfg;
	   c.n_streams = hi - lo;
	   Lane *l = lane_create (&c, status);
	   if (!l) { sdrjfm_destroy (h); return nullptr; }
	   h -> lanes.push_back (l);
	   h -> first.push_back (lo);
	}
	h -> first.push_back (cfg -> n_streams);
//	K1t is persistent: the lanes share the SMs (one CTA per SM in total keeps HBM saturated and leaves
//	the rest of each SM to the latency-bound kernels of the other lanes)
	if (!getenv ("SDRJFM_TMA_CTAS"))
	   for (Lane *l : h -> lanes) l -> tma_ctas = std::max (1, (l -> n_sm + nl - 1) / nl);
	h -> cfg = h -> lanes [0] -> cfg; h -> cfg.n_streams = cfg -> n_streams;
	h -> cap_in = h -> lanes [0] -> cap_in; h -> cap_audio = h -> lanes [0] -> cap_out;     // at audio_rate
	h -> cap_rds = h -> lanes [0] -> cap_rds;
auto fail = [&](cudaError_t e, const char *what) -> sdrjfm_handle * {
	   g_create_error = std::string (what) + ": " + cudaGetErrorString (e);
	   *status = SDRJFM_ERR_CUDA; sdrjfm_destroy (h); return nullptr;
	};
cudaError_t e;
	if (nl == 1) h -> stream = h -> lanes [0] -> stream;
	else {
	   if ((e = cudaStreamCreateWithFlags (&h -> stream, cudaStreamNonBlocking)) != cudaSuccess) return fail (e, "cudaStreamCreate");
	   h -> own_stream = true;
	   if ((e = cudaEventCreateWithFlags (&h -> ev_fork, cudaEventDisableTiming)) != cudaSuccess) return fail (e, "cudaEventCreate");
	   h -> ev_join.resize (nl, nullptr);
	   for (int i = 0; i < nl; i ++)
	      if ((e = cudaEventCreateWithFlags (&h -> ev_join [i], cudaEventDisableTiming)) != cudaSuccess) return fail (e, "cudaEventCreate");
	}
const size_t S = cfg -> n_streams;
	if ((e = cudaMalloc ((void **)&h -> d_in [0], S * h -> cap_in * sizeof (float2))) != cudaSuccess ||
	    (e = cudaMalloc ((void **)&h -> d_audio, S * h -> cap_audio * sizeof (float2))) != cudaSuccess ||
	    (e = cudaMalloc ((void **)&h -> d_rds24, S * h -> cap_rds * sizeof (float2))) != cudaSuccess)
	   return fail (e, "cudaMalloc staging");
	*status = SDRJFM_OK;
	return h;
}

int sdrjfm_destroy (sdrjfm_handle *h) {
	if (!h) return SDRJFM_ERR_ARG;
	LOCK_DEV (h);
	cudaSetDevice (h -> cfg.device);
	if (h -> stream) cudaStreamSynchronize (h -> stream);
	for (Lane *l : h -> lanes) lane_destroy (l);
	for (float2 *p : { h -> d_in [0], h -> d_in [1], h -> d_audio, h -> d_rds24 }) if (p) cudaFree (p);
	for (cudaEvent_t ev : h -> ev_join) if (ev) cudaEventDestroy (ev);
	for (cudaEvent_t ev : { h -> ev_fork, h -> ev_h2d [0], h -> ev_h2d [1], h -> ev_done [0], h -> ev_done [1] })
	   if (ev) cudaEventDestroy (ev);
	if (h -> copy_stream) cudaStreamDestroy (h -> copy_stream);
	if (h -> out_stream) cudaStreamDestroy (h -> out_stream);
	if (h -> own_stream && h -> stream) cudaStreamDestroy (h -> stream);
	delete h;
	return SDRJFM_OK;
}

void *sdrjfm_cuda_stream (sdrjfm_handle *h) { return h ? (void *)h -> stream : nullptr; }

int64_t sdrjfm_launch_count (const sdrjfm_handle *h) {
int64_t n = 0;
	if (h) for (const Lane *l : h -> lanes) n += l -> launches;
	return n;
}

int sdrjfm_sync (sdrjfm_handle *h) {
	if (!h) return SDRJFM_ERR_ARG;
	LOCK_DEV (h);
	HK (cudaStreamSynchronize (h -> stream));
	return for_lanes (h, [](Lane *l) { return lane_sync (l); });
}

// scale of a device sample format: what the reference's handler divides by (power of two: exact)
static int raw_scale (sdrjfm_handle *h, int32_t format, int32_t denominator, float *scale) {
	switch (format) {
	   case SDRJFM_IQ_CF32: *scale = 1.0f; return SDRJFM_OK;
	   case SDRJFM_IQ_U8: case SDRJFM_IQ_S8: *scale = 1.0f / 128.0f; return SDRJFM_OK;
	   case SDRJFM_IQ_S16: case SDRJFM_IQ_AIRSPY_S16:
	      if (denominator <= 0 || (denominator & (denominator - 1)) != 0) {
	         h -> err = "int16 denominator must be a power of two (2048, 4096, 8192 in the reference's handlers)";
	         return SDRJFM_ERR_ARG;
	      }
	      *scale = 1.0f / (float)denominator; return SDRJFM_OK;
	   default: h -> err = "unknown sample format"; return SDRJFM_ERR_ARG;
	}
}

static int process_dev (sdrjfm_handle *h, const void *d_iq, int32_t fmt, float scale, int64_t n_in, int64_t in_pitch,
                        float *d_audio, int64_t audio_pitch, int64_t *n_audio,
                        float *d_rds24, int64_t rds_pitch, int64_t *n_rds) {
	if (!h || n_in < 0 || (n_in > 0 && (!d_iq || in_pitch < n_in))) return SDRJFM_ERR_ARG;
	if (n_in > h -> cfg.max_samples_per_call) { h -> err = "n_in exceeds max_samples_per_call"; return SDRJFM_ERR_CAPACITY; }
	LOCK_DEV (h);
	CHECK_POISON (h);
	HK (cudaSetDevice (h -> cfg.device));
int64_t na = 0, nr = 0;
const int64_t bps = fmt_bytes (fmt);
const int rc = fork_join (h, [&](Lane *l, int32_t s0) {
	   return lane_process_device (l, (const char *)d_iq + bps * (int64_t)s0 * in_pitch, fmt, scale, n_in, in_pitch,
	                               d_audio ? d_audio + 2 * (int64_t)s0 * audio_pitch : nullptr, audio_pitch, &na,
	                               d_rds24 ? d_rds24 + 2 * (int64_t)s0 * rds_pitch : nullptr, rds_pitch, &nr);
	});
	if (n_audio) *n_audio = na;
	if (n_rds) *n_rds = nr;
	return rc;
}

int sdrjfm_process_device (sdrjfm_handle *h, const float *d_iq, int64_t n_in, int64_t in_pitch,
                           float *d_audio, int64_t audio_pitch, int64_t *n_audio,
                           float *d_rds24, int64_t rds_pitch, int64_t *n_rds) {
	return process_dev (h, d_iq, SDRJFM_IQ_CF32, 1.0f, n_in, in_pitch, d_audio, audio_pitch, n_audio,
	                    d_rds24, rds_pitch, n_rds);
}

int sdrjfm_process_raw_device (sdrjfm_handle *h, const void *d_iq, int32_t format, int32_t denominator,
                               int64_t n_in, int64_t in_pitch,
                               float *d_audio, int64_t audio_pitch, int64_t *n_audio,
                               float *d_rds24, int64_t rds_pitch, int64_t *n_rds) {
	if (!h) return SDRJFM_ERR_ARG;
float scale;
int rc = raw_scale (h, format, denominator, &scale);
	if (rc != SDRJFM_OK) return rc;
	return process_dev (h, d_iq, format, scale, n_in, in_pitch, d_audio, audio_pitch, n_audio,
	                    d_rds24, rds_pitch, n_rds);
}

int sdrjfm_run_frontend_only (sdrjfm_handle *h, const float *d_iq, int64_t n_in, int64_t in_pitch) {
	return sdrjfm_run_frontend_only_raw (h, d_iq, SDRJFM_IQ_CF32, 1, n_in, in_pitch);
}

int sdrjfm_run_frontend_only_raw (sdrjfm_handle *h, const void *d_iq, int32_t format, int32_t denominator,
                                  int64_t n_in, int64_t in_pitch) {
	if (!h || !d_iq) return SDRJFM_ERR_ARG;
float scale;
int rc = raw_scale (h, format, denominator, &scale);
	if (rc != SDRJFM_OK) return rc;
	LOCK_DEV (h);
	CHECK_POISON (h);
	HK (cudaSetDevice (h -> cfg.device));
const int64_t bps = fmt_bytes (format);
	return fork_join (h, [&](Lane *l, int32_t s0) {
	   return lane_run_frontend_only (l, (const char *)d_iq + bps * (int64_t)s0 * in_pitch, format, scale, n_in, in_pitch);
	});
}

static int process_host (sdrjfm_handle *h, const void *iq, int32_t fmt, float scale, int64_t n_in, int64_t in_pitch,
                         float *audio, int64_t audio_pitch, int64_t *n_audio,
                         float *rds24, int64_t rds_pitch, int64_t *n_rds, sdrjfm_meta *meta) {
	if (!h || n_in < 0 || (n_in > 0 && (!iq || in_pitch < n_in))) return SDRJFM_ERR_ARG;
	if (n_in > h -> cfg.max_samples_per_call) { h -> err = "n_in exceeds max_samples_per_call"; return SDRJFM_ERR_CAPACITY; }
	LOCK_DEV (h);
	CHECK_POISON (h);
	HK (cudaSetDevice (h -> cfg.device));
const int S = h -> cfg.n_streams;
const size_t bps = (size_t)fmt_bytes (fmt);
const size_t rowb = (size_t)h -> cap_in * bps;                 // staging row pitch in bytes
int rc;
int64_t na = 0, nr = 0;
//	Large offline calls (no tap read-back wanted): the call is cut into time slices and the
//	host->device copy of slice c+1 runs on a second stream while slice c computes.  The chain is
//	stateful, so this is exactly the sequence of smaller calls the GUI cadence would make.
int64_t kSlices = 16;
	{ const char *env = getenv ("SDRJFM_SLICES"); if (env && atoi (env) > 0) kSlices = atoi (env); }
const int64_t unit = 256 * (int64_t)h -> lanes [0] -> decim;
const int64_t slice = ((n_in / kSlices) / unit) * unit;          // multiple of the decimation
//	The per-call side outputs (RDS bits, scan blocks, scope stream) describe ONE chain call: they would
//	be overwritten slice by slice, so calls that produce them are not cut.
const Lane *l0 = h -> lanes [0];
	if (!h -> cfg.keep_taps && l0 -> lf_plot < 0 && !l0 -> rds_symbols && !l0 -> scanning && l0 -> hf_N == 0 && slice >= (1 << 16)) {
//	   output capacities are checked before the first slice moves any state: a call of n_in samples
//	   yields at most n_in / decim / 4 + 1 audio and n_in / decim / 8 + 1 RDS samples per stream
	   const int64_t max_fm = (n_in + l0 -> pend) / l0 -> decim + 1;
	   if ((audio && audio_pitch < (max_fm / kRsDecim + 1) * l0 -> cvL / l0 -> cvM + 1) || (rds24 && l0 -> set.rds_mode != 0 && rds_pitch < max_fm / 8 + 1)) {
	      h -> err = "audio_pitch / rds_pitch too small for this call"; return SDRJFM_ERR_ARG;
	   }
	   if (!h -> copy_stream) {
	      HK (cudaStreamCreateWithFlags (&h -> copy_stream, cudaStreamNonBlocking));
	      HK (cudaStreamCreateWithFlags (&h -> out_stream, cudaStreamNonBlocking));
	      for (int i = 0; i < 2; i ++) {
	         HK (cudaEventCreateWithFlags (&h -> ev_h2d [i], cudaEventDisableTiming));
	         HK (cudaEventCreateWithFlags (&h -> ev_done [i], cudaEventDisableTiming));
	      }
	      HK (cudaMalloc ((void **)&h -> d_in [1], (size_t)S * h -> cap_in * sizeof (float2)));
	   }
	   int64_t pos = 0; int c = 0;
	   std::vector<int64_t> peak_e0;               // the peak read-outs of the whole call, not of its last slice
	   while (pos < n_in) {
	      const int64_t rest = n_in - pos;
	      const int64_t len = rest < 2 * slice ? rest : slice;     // the last slice takes the ragged tail
	      const int b = c & 1;
	      if (c >= 2) HK (cudaStreamWaitEvent (h -> copy_stream, h -> ev_done [b], 0));
	      HK (cudaMemcpy2DAsync (h -> d_in [b], rowb, (const char *)iq + pos * bps,
	                             in_pitch * bps, len * bps, S, cudaMemcpyHostToDevice, h -> copy_stream));
	      HK (cudaEventRecord (h -> ev_h2d [b], h -> copy_stream));
	      HK (cudaStreamWaitEvent (h -> stream, h -> ev_h2d [b], 0));
	      int64_t a1 = 0, r1 = 0;
	      rc = process_dev (h, h -> d_in [b], fmt, scale, len, h -> cap_in,
	                        (float *)(h -> d_audio + na), h -> cap_audio, &a1,
	                        (float *)(h -> d_rds24 + nr), h -> cap_rds, &r1);
	      if (rc != SDRJFM_OK) {
	         if (c > 0) { cudaDeviceSynchronize (); h -> poisoned = true; }     // earlier slices already advanced the streams
	         return rc;
	      }
	      if (c == 0) for (Lane *l : h -> lanes) peak_e0.push_back (l -> peak_e0);
	      HK (cudaEventRecord (h -> ev_done [b], h -> stream));
//	      this slice's outputs go back on a third stream (PCIe is full duplex) while the next slices run
	      if ((audio && a1 > 0) || (rds24 && r1 > 0)) {
	         HK (cudaStreamWaitEvent (h -> out_stream, h -> ev_done [b], 0));
	         if (audio && a1 > 0)
	            HK (cudaMemcpy2DAsync ((float2 *)audio + na, audio_pitch * sizeof (float2), h -> d_audio + na,
	                                   h -> cap_audio * sizeof (float2), a1 * sizeof (float2), S,
	                                   cudaMemcpyDeviceToHost, h -> out_stream));
	         if (rds24 && r1 > 0)
	            HK (cudaMemcpy2DAsync ((float2 *)rds24 + nr, rds_pitch * sizeof (float2), h -> d_rds24 + nr,
	                                   h -> cap_rds * sizeof (float2), r1 * sizeof (float2), S,
	                                   cudaMemcpyDeviceToHost, h -> out_stream));
	      }
	      na += a1; nr += r1; pos += len; c ++;
	   }
	   HK (cudaStreamSynchronize (h -> out_stream));
	   HK (cudaStreamSynchronize (h -> stream));
	   for (size_t i = 0; i < peak_e0.size (); i ++) h -> lanes [i] -> peak_e0 = peak_e0 [i];
	   if (n_audio) *n_audio = na;
	   if (n_rds) *n_rds = nr;
	   if (meta) return sdrjfm_get_meta (h, meta);
	   return SDRJFM_OK;
	}
	else {
	   if (n_in)
	      HK (cudaMemcpy2DAsync (h -> d_in [0], rowb, iq, in_pitch * bps, n_in * bps, S,
	                             cudaMemcpyHostToDevice, h -> stream));
	   rc = process_dev (h, h -> d_in [0], fmt, scale, n_in, h -> cap_in, (float *)h -> d_audio,
	                     h -> cap_audio, &na, (float *)h -> d_rds24, h -> cap_rds, &nr);
	   if (rc != SDRJFM_OK) return rc;
	}
	if (audio && na > 0) {
	   if (audio_pitch < na) return SDRJFM_ERR_ARG;
	   HK (cudaMemcpy2DAsync (audio, audio_pitch * sizeof (float2), h -> d_audio,
	                          h -> cap_audio * sizeof (float2), na * sizeof (float2), S,
	                          cudaMemcpyDeviceToHost, h -> stream));
	}
	if (rds24 && nr > 0) {
	   if (rds_pitch < nr) return SDRJFM_ERR_ARG;
	   HK (cudaMemcpy2DAsync (rds24, rds_pitch * sizeof (float2), h -> d_rds24,
	                          h -> cap_rds * sizeof (float2), nr * sizeof (float2), S,
	                          cudaMemcpyDeviceToHost, h -> stream));
	}
	HK (cudaStreamSynchronize (h -> stream));
	if (n_audio) *n_audio = na;
	if (n_rds) *n_rds = nr;
	if (meta) return sdrjfm_get_meta (h, meta);
	return SDRJFM_OK;
}

int sdrjfm_process (sdrjfm_handle *h, const float *iq, int64_t n_in, int64_t in_pitch,
                    float *audio, int64_t audio_pitch, int64_t *n_audio,
                    float *rds24, int64_t rds_pitch, int64_t *n_rds, sdrjfm_meta *meta) {
	return process_host (h, iq, SDRJFM_IQ_CF32, 1.0f, n_in, in_pitch, audio, audio_pitch, n_audio,
	                     rds24, rds_pitch, n_rds, meta);
}

int sdrjfm_process_raw (sdrjfm_handle *h, const void *iq, int32_t format, int32_t denominator,
                        int64_t n_in, int64_t in_pitch,
                        float *audio, int64_t audio_pitch, int64_t *n_audio,
                        float *rds24, int64_t rds_pitch, int64_t *n_rds, sdrjfm_meta *meta) {
	if (!h) return SDRJFM_ERR_ARG;
float scale;
int rc = raw_scale (h, format, denominator, &scale);
	if (rc != SDRJFM_OK) return rc;
	return process_host (h, iq, format, scale, n_in, in_pitch, audio, audio_pitch, n_audio,
	                     rds24, rds_pitch, n_rds, meta);
}

int sdrjfm_get_meta (sdrjfm_handle *h, sdrjfm_meta *meta) {
	if (!h || !meta) return SDRJFM_ERR_ARG;
	LOCK_DEV (h);
	HK (cudaStreamSynchronize (h -> stream));
	for (size_t i = 0; i < h -> lanes.size (); i ++) {
	   const int rc = lane_get_meta (h -> lanes [i], meta + h -> first [i]);
	   if (rc != SDRJFM_OK) { h -> err = h -> lanes [i] -> err; return rc; }
	}
	return SDRJFM_OK;
}

int sdrjfm_pilot_stats (sdrjfm_handle *h, int32_t *out) {
	if (!h || !out) return SDRJFM_ERR_ARG;
	LOCK_DEV (h);
	HK (cudaStreamSynchronize (h -> stream));
	for (size_t i = 0; i < h -> lanes.size (); i ++) {
	   const int rc = lane_pilot_stats (h -> lanes [i], out + 4 * h -> first [i]);
	   if (rc != SDRJFM_OK) { h -> err = h -> lanes [i] -> err; return rc; }
	}
	return SDRJFM_OK;
}

int64_t sdrjfm_read_tap (sdrjfm_handle *h, int which, int32_t stream, void *out, int64_t cap) {
	if (!h || !out) return SDRJFM_ERR_ARG;
	LOCK_DEV (h);
const int i = lane_of (h, stream);
	if (i < 0) return SDRJFM_ERR_ARG;
	HK (cudaStreamSynchronize (h -> stream));
const int64_t n = lane_read_tap (h -> lanes [i], which, stream - h -> first [i], out, cap);
	if (n < 0) h -> err = h -> lanes [i] -> err;
	return n;
}

// ---- settings: forwarded to every lane ------------------------------------------------------
#define FWD1(name, T)                                                                     \
int sdrjfm_##name (sdrjfm_handle *h, T v) { return for_lanes (h, [&](Lane *l) { return lane_##name (l, v); }); }
FWD1 (set_fm_mode, int32_t)
FWD1 (set_fm_decoder, int32_t)
FWD1 (set_sound_mode, int32_t)
FWD1 (set_stereo_panorama, int32_t)
FWD1 (set_sound_balance, int32_t)
FWD1 (set_deemphasis, int32_t)
FWD1 (set_volume_db, float)
FWD1 (set_lf_cutoff, int32_t)
FWD1 (set_bandwidth, int32_t)
FWD1 (set_rds_mode, int32_t)
FWD1 (set_local_oscillator, int32_t)
FWD1 (set_squelch_mode, int32_t)
FWD1 (set_squelch_value, int32_t)
FWD1 (set_native_rate, int32_t)
FWD1 (set_rds_symbol_stage, int32_t)
FWD1 (set_scanning, int32_t)
FWD1 (set_lf_plot_type, int32_t)
FWD1 (set_lf_plot_zoom, int32_t)
int sdrjfm_set_lf_spectrum (sdrjfm_handle *h, int32_t spectrum_size, int32_t display_size, int32_t average_count) {
	return for_lanes (h, [&](Lane *l) { return lane_set_lf_spectrum (l, spectrum_size, display_size, average_count); });
}
int64_t sdrjfm_read_lf_spectrum (sdrjfm_handle *h, int32_t stream, double *display, int64_t cap, int32_t *blocks) {
	if (!h || !display) return SDRJFM_ERR_ARG;
	LOCK_DEV (h);
const int i = lane_of (h, stream);
	if (i < 0) return SDRJFM_ERR_ARG;
	HK (cudaStreamSynchronize (h -> stream));
const int64_t n = lane_read_lf_spectrum (h -> lanes [i], stream - h -> first [i], display, cap, blocks);
	if (n < 0) h -> err = h -> lanes [i] -> err;
	return n;
}
int sdrjfm_set_hf_spectrum (sdrjfm_handle *h, int32_t display_size, int32_t repeat_rate) {
	return for_lanes (h, [&](Lane *l) { return lane_set_hf_spectrum (l, display_size, repeat_rate); });
}
int64_t sdrjfm_read_hf_spectrum (sdrjfm_handle *h, int32_t stream, double *display, int64_t cap, int32_t *blocks) {
	if (!h || !display) return SDRJFM_ERR_ARG;
	LOCK_DEV (h);
const int i = lane_of (h, stream);
	if (i < 0) return SDRJFM_ERR_ARG;
	HK (cudaStreamSynchronize (h -> stream));
const int64_t n = lane_read_hf_spectrum (h -> lanes [i], stream - h -> first [i], display, cap, blocks);
	if (n < 0) h -> err = h -> lanes [i] -> err;
	return n;
}
int64_t sdrjfm_read_lf_plot (sdrjfm_handle *h, int32_t stream, float *out, int64_t cap,
                             int32_t *sample_rate, int32_t *show_full) {
	if (!h || !out) return SDRJFM_ERR_ARG;
	LOCK_DEV (h);
const int i = lane_of (h, stream);
	if (i < 0) return SDRJFM_ERR_ARG;
	HK (cudaStreamSynchronize (h -> stream));
Lane *l = h -> lanes [i];
const int64_t n = lane_read_lf_plot (l, stream - h -> first [i], out, cap);
	if (n < 0) { h -> err = l -> err; return n; }
//	spectrumSampleRate / showFullSpectrum as setlfPlotType leaves them (fm-processor.cpp:247-263; RDS_RATE = 24000)
const int t = l -> lf_plot;
	if (sample_rate) *sample_rate = t == 8 ? 24000 : t == 9 ? 24000 / 16 : h -> cfg.fm_rate;
	if (show_full) *show_full = (t == 1 || t == 8 || t == 9) ? 1 : 0;
	return n;
}
int64_t sdrjfm_read_scan (sdrjfm_handle *h, int32_t stream, float *out, int64_t cap_pairs) {
	if (!h || !out) return SDRJFM_ERR_ARG;
	LOCK_DEV (h);
const int i = lane_of (h, stream);
	if (i < 0) return SDRJFM_ERR_ARG;
	HK (cudaStreamSynchronize (h -> stream));
const int64_t n = lane_read_scan (h -> lanes [i], stream - h -> first [i], out, cap_pairs);
	if (n < 0) h -> err = h -> lanes [i] -> err;
	return n;
}
int64_t sdrjfm_read_rds_bits (sdrjfm_handle *h, int32_t stream, uint8_t *out, int64_t cap) {
	if (!h || !out) return SDRJFM_ERR_ARG;
	LOCK_DEV (h);
const int i = lane_of (h, stream);
	if (i < 0) return SDRJFM_ERR_ARG;
	HK (cudaStreamSynchronize (h -> stream));
const int64_t n = lane_read_rds_bits (h -> lanes [i], stream - h -> first [i], out, cap);
	if (n < 0) h -> err = h -> lanes [i] -> err;
	return n;
}
int64_t sdrjfm_read_rds_groups (sdrjfm_handle *h, int32_t stream, uint16_t *blocks, int64_t cap_groups, int32_t *status) {
	if (!h || cap_groups < 0) return SDRJFM_ERR_ARG;
	LOCK_DEV (h);
const int i = lane_of (h, stream);
	if (i < 0) return SDRJFM_ERR_ARG;
	HK (cudaStreamSynchronize (h -> stream));
const int64_t n = lane_read_rds_groups (h -> lanes [i], stream - h -> first [i], blocks, cap_groups, status);
	if (n < 0) h -> err = h -> lanes [i] -> err;
	return n;
}
FWD1 (set_test_tone, int32_t)
FWD1 (set_disp_delay, int32_t)
int64_t sdrjfm_read_peak_levels (sdrjfm_handle *h, int32_t stream, float *out, int64_t cap_pairs) {
	if (!h || !out) return SDRJFM_ERR_ARG;
	LOCK_DEV (h);
const int i = lane_of (h, stream);
	if (i < 0) return SDRJFM_ERR_ARG;
	HK (cudaStreamSynchronize (h -> stream));
const int64_t n = lane_read_peak_levels (h -> lanes [i], stream - h -> first [i], out, cap_pairs);
	if (n < 0) h -> err = h -> lanes [i] -> err;
	return n;
}
FWD1 (set_auto_mono, int32_t)
FWD1 (set_pss_mode, int32_t)
FWD1 (set_dc_remove, int32_t)
int sdrjfm_set_attenuation (sdrjfm_handle *h, float l, float r) {
	return for_lanes (h, [&](Lane *ln) { return lane_set_attenuation (ln, l, r); });
}
int sdrjfm_trigger_frequency_change (sdrjfm_handle *h) { return for_lanes (h, [](Lane *l) { return lane_trigger_frequency_change (l); }); }
int sdrjfm_restart_pss_analyzer (sdrjfm_handle *h) { return for_lanes (h, [](Lane *l) { return lane_restart_pss_analyzer (l); }); }

// ---- shared tables --------------------------------------------------------------------------
int64_t sdrjfm_design_tables (int32_t input_rate, int32_t fm_rate, int32_t input_filter_hz,
                              int32_t audio_lp_hz, void *out, int64_t cap) {
	if (input_rate < 6 * fm_rate || fm_rate <= 0) return SDRJFM_ERR_ARG;
TableBlob b = build_tables (input_rate, fm_rate, input_filter_hz, audio_lp_hz);
	if (out && cap >= (int64_t)b.bytes.size ()) memcpy (out, b.bytes.data (), b.bytes.size ());
	return (int64_t)b.bytes.size ();
}
// host-only designers of the small tables that are not part of the blob (no device needed; used by the CPU tests)
int64_t sdrjfm_design_aux (int32_t which, int32_t a, int32_t b, float *out, int64_t cap) {
	if (!out) return SDRJFM_ERR_ARG;
	if (which == 0) {                                     // RDS_2 matched filter at rate a
	   if (cap < kRs2Taps) return SDRJFM_ERR_CAPACITY;
	   design_rds2_matched_filter (a, out); return kRs2Taps;
	}
	if (which == 1) {                                     // test-tone burst at working rate a; out [0] = samples before the burst
	   std::vector<float> t; int32_t arm = 0;
	   tone_design (a, arm, t);
	   if (cap < (int64_t)t.size () + 1) return SDRJFM_ERR_CAPACITY;
	   out [0] = (float)arm; memcpy (out + 1, t.data (), t.size () * sizeof (float));
	   return (int64_t)t.size () + 1;
	}
	if (which == 2) {                                     // second converter a -> b: out [0] = L, out [1] = M, then L * 32 taps
	   std::vector<float> t; int L = 0, M = 0;
	   if (!convert_design (a, b, L, M, t)) return SDRJFM_ERR_UNSUPPORTED;
	   if (cap < (int64_t)t.size () + 2) return SDRJFM_ERR_CAPACITY;
	   out [0] = (float)L; out [1] = (float)M; memcpy (out + 2, t.data (), t.size () * sizeof (float));
	   return (int64_t)t.size () + 2;
	}
	return SDRJFM_ERR_ARG;
}
int64_t sdrjfm_tables_nbytes (const sdrjfm_handle *h) { return h ? lane_tables_nbytes (h -> lanes [0]) : SDRJFM_ERR_ARG; }
int sdrjfm_tables_export (const sdrjfm_handle *h, void *out, int64_t cap) {
	return h ? lane_tables_export (h -> lanes [0], out, cap) : SDRJFM_ERR_ARG;
}
int sdrjfm_tables_import (sdrjfm_handle *h, const void *blob, int64_t nbytes) {
//	validated by every lane BEFORE anything is replaced (lane_tables_import), so a bad blob leaves the handle as it was
	return for_lanes (h, [&](Lane *l) { return lane_tables_import (l, blob, nbytes); });
}

}	// extern "C"
