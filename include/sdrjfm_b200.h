/*
 * sdrjfm_b200 — C ABI of the B200-native FM demodulation path.
 *
 * This is the drop-in boundary for ONE path of JvanKatwijk/sdr-j-fm: what
 * fmProcessor::run does between deviceHandler::getSamples and audioSink::putSample /
 * rdsDecoder::doDecode.  Plain pointers and sizes only; no C++/Qt/torch types; never
 * throws; every entry point returns an int status (0 = SDRJFM_OK) unless stated.
 * One calling thread per handle (the reference has exactly one: the fmProcessor QThread); different
 * handles may be driven from different threads (calls into one device are serialised inside the
 * library: the tap sets live in per-device constant banks), and a process may hold handles on several
 * devices (sdrjfm_config.device).
 * A call that fails with an argument / capacity / unsupported status leaves the handle as it was; a call
 * that fails after stream state has moved (a CUDA error in mid-flight) poisons the handle: every later
 * call returns SDRJFM_ERR_CUDA until it is destroyed.
 * Environment: when the library is loaded before the process creates its CUDA context it sets
 * CUDA_DEVICE_MAX_CONNECTIONS=32 unless the host set a value itself (a handle works on up to 11
 * streams); SDRJFM_NO_ENV_HINT=1 leaves the environment untouched.
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * reference tree).  INTEGRATION.md shows the reference-side binding.
 *
 * Sample formats (same as the reference):
 *   IQ in      : interleaved float32 (re, im) = std::complex<float>, what
 *                deviceHandler::getSamples delivers (devices/device-handler.h:72-75)
 *   audio out  : interleaved float32 (left, right) = DSPCOMPLEX re=L im=R, what
 *                audioSink::putSample takes (includes/output/audiosink.h:44)
 *   rds out    : interleaved float32 complex at 24 kHz, what rdsDecoder::doDecode takes
 *                (includes/rds/rds-decoder.h:67-69)
 */
#ifndef SDRJFM_B200_H
#define SDRJFM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDRJFM_OK              0
#define SDRJFM_ERR_ARG        -1   /* bad argument                                        */
#define SDRJFM_ERR_CUDA       -2   /* a CUDA call failed; see sdrjfm_last_error           */
#define SDRJFM_ERR_NO_DEVICE  -3   /* no sm_100 device: this library has NO CPU fallback  */
#define SDRJFM_ERR_CAPACITY   -4   /* n_in exceeds max_samples_per_call                   */
#define SDRJFM_ERR_UNSUPPORTED -5  /* setting not implemented on the GPU path             */

typedef struct sdrjfm_handle sdrjfm_handle;

/* Constructor arguments of fmProcessor (src/fm/fm-processor.cpp:48-63) that matter to
 * the arithmetic, plus the batch shape.  Zero-initialise, then fill.                    */
typedef struct sdrjfm_config {
    int32_t input_rate;            /* inputRate, 2304000 (includes/fm-constants.h:35); any rate from
                                      1152000 to 12670000 with the reference's own integer-decimation
                                      arithmetic (fm-processor.cpp:36,68-75): stage 1 /6, stage 2
                                      /(inputRate/6/fmRate), i.e. /6 .. /60.  /12, /30, /48 (2.304 / 2.4,
                                      6, 10 MS/s) have the tuned HBM-bound kernels; every other rate runs
                                      the reference-order front end (front_end_mode 2 semantics)          */
    int32_t fm_rate;               /* fmRate, 192000 (radio.cpp:68)                      */
    int32_t working_rate;          /* workingRate, 48000 (radio.cpp:233)                 */
    int32_t audio_rate;            /* audioRate, 48000 (main.cpp:42); 192000 with -m (main.cpp:57-65), or the ini
                                      file's value (radio.cpp:250-252).  When it differs from working_rate the
                                      second converter runs (theConverter, fm-processor.cpp:89-91, 825-838) and
                                      the audio output of the process calls is at audio_rate               */
    int32_t n_streams;             /* independent IQ streams handled by this handle (>=1)*/
    int32_t device;                /* CUDA device ordinal                                */
    int64_t max_samples_per_call;  /* per stream; sizes the device buffers               */
    int32_t keep_taps;             /* 1: keep fm-rate intermediates for sdrjfm_read_tap  */
    int32_t front_end_mode;        /* 0: the reference's integer decimation (12m+11 index contract), evaluated as
                                         ONE real polyphase FIR with the DC removal finished at the fm rate
                                         (HBM-bound; fm-rate samples within 3e-7 of the reference's).  The
                                         library moves a stream onto mode 2 by itself while the PLL or
                                         real-baseband decoder is selected or the oscillator is non-zero
                                         (their look-up-table indices flip on a 3e-7 difference);
                                      1: rational polyphase resampler to exactly fm_rate (new block,
                                         BASELINE config 4): /5 low-pass, then L/M = 5 fm_rate / input_rate
                                         (2/5, 4/25, 12/125 at 2.4, 6, 10 MS/s); csrc/resample.cuh;
                                      2: the same arithmetic as 0 in the REFERENCE'S OPERATION ORDER: RF DC
                                         one-pole per input sample in float32, IQ gain, oscillator, fmBand_1
                                         and fmBand_2 tap by tap (fm-processor.cpp:423-446, 462-475;
                                         fir-filters.cpp:397-424): fm-rate samples bit-identical to the
                                         reference's.  Latency-bound (the DC recurrence); with inputFilter
                                         on, mode 0 is used; csrc/frontend_exact.cuh                       */
} sdrjfm_config;

/* fmProcessor::SMetaData (includes/fm/fm-processor.h:91-101) per stream, plus the RF DC
 * estimate and carrier level the GUI derives its read-outs from.                        */
typedef struct sdrjfm_meta {
    float   dc_rf_re, dc_rf_im;    /* RfDC (fm-processor.cpp:425)                        */
    float   dc_rf_db;              /* DcValRf (fm-processor.cpp:668-670)                 */
    float   dc_if;                 /* DcValIf = fm_afc (fm-demodulator.cpp:197)          */
    float   carrier_ampl;          /* am_carr_ampl (fm-demodulator.cpp:130)              */
    float   pss_phase_shift_deg;   /* PssPhaseShiftDegree (:673)                         */
    float   pss_phase_change;      /* PssPhaseChange (:675)                              */
    int32_t pss_state;             /* EPssState 0 OFF, 1 ANALYZING, 2 ESTABLISHED (:676) */
    float   pilot_lock_strength;   /* PilotPllLockStrength (:666)                        */
    int32_t pilot_locked;          /* PilotPllLocked                                     */
    float   peak_left_db, peak_right_db; /* the newest showPeakLevel read-out (evaluatePeakLevel, :772-798: peak of
                                            |left|, |right| per 961 PCM samples in dB, behind the display delay
                                            line); -40 before the first.  All read-outs: sdrjfm_read_peak_levels */
    int32_t squelch_active;        /* getSquelchState (:217-219)                          */
} sdrjfm_meta;

/* intermediate taps (fm rate unless noted) readable after a process call when keep_taps */
enum sdrjfm_tap {
    SDRJFM_TAP_FM_Z = 0,       /* complex: after fmBand_2 (fm-processor.cpp:474)         */
    SDRJFM_TAP_DEMOD = 1,      /* float  : fm_Demodulator::demodulate (:497)             */
    SDRJFM_TAP_PILOT_PHASE = 2,/* float  : pilotRecovery::getPilotPhase (:695)           */
    SDRJFM_TAP_LOCKED = 3,     /* uint8  : pilotRecovery::isLocked (:697)                */
    SDRJFM_TAP_PSS_DELAY = 4,  /* float  : pilotDelayPSS (:716)                          */
    SDRJFM_TAP_LR = 5,         /* complex: (left,right) after the selector (:527-549)    */
    SDRJFM_TAP_AUDIO192 = 6,   /* complex: after de-emphasis and gain (:594-595,:630)    */
    SDRJFM_TAP_RDS_CPLX = 7,   /* complex: rdsDataCplx (:754)                            */
    SDRJFM_TAP_RDS24 = 8,      /* complex @24 kHz: rdsSample (:553)                      */
    SDRJFM_TAP_FE_U = 9,       /* complex: raw output of the front-end FIR (before DC subtraction and the constant gain) */
    SDRJFM_TAP_FE_S = 10       /* complex: block sums of the input samples behind each fm-rate sample (DC remover)      */
};

/* --- life cycle: replaces `new fmProcessor (...)` / `delete` (radio.cpp:908-949, 629-644) */
sdrjfm_handle *sdrjfm_create (const sdrjfm_config *cfg, int *status);
int  sdrjfm_destroy (sdrjfm_handle *h);
const char *sdrjfm_last_error (const sdrjfm_handle *h);   /* h may be NULL: create errors */
const char *sdrjfm_version (void);

/* --- data path: replaces the body of fmProcessor::run (src/fm/fm-processor.cpp:387-686).
 * HOST buffers.  iq: n_streams rows of n_in complex samples, row r at iq + 2*r*in_pitch
 * floats (in_pitch in complex samples, >= n_in).  Stateful across calls; any n_in >= 0.
 * audio: n_streams rows, row pitch audio_pitch complex samples, at working_rate;
 * rds24: n_streams rows, row pitch rds_pitch, at 24 kHz.  audio/rds24/meta may be NULL.
 * *n_audio / *n_rds receive the per-stream counts produced by THIS call (equal for all
 * streams: they depend on the sample count only, never on content).                     */
int  sdrjfm_process (sdrjfm_handle *h,
                     const float *iq, int64_t n_in, int64_t in_pitch,
                     float *audio, int64_t audio_pitch, int64_t *n_audio,
                     float *rds24, int64_t rds_pitch, int64_t *n_rds,
                     sdrjfm_meta *meta /* [n_streams] */);

/* Same, with DEVICE pointers (inputs already resident in HBM, outputs left in HBM):
 * the zero-copy form used by batch replay and by bench.py's device-resident timing.
 * Asynchronous on the handle's stream; sdrjfm_sync waits.                               */
int  sdrjfm_process_device (sdrjfm_handle *h,
                            const float *d_iq, int64_t n_in, int64_t in_pitch,
                            float *d_audio, int64_t audio_pitch, int64_t *n_audio,
                            float *d_rds24, int64_t rds_pitch, int64_t *n_rds);

/* Device-native sample formats.  The reference's device handlers convert to complex float on
 * the CPU before the ring buffer; here the conversion is fused into the front-end kernel's only
 * read of the samples, so 2 or 4 bytes per IQ sample cross PCIe / HBM instead of 8.  The floats
 * entering the filters are bit-identical to the handler's (all divisors are powers of two). */
enum sdrjfm_iq_format {
    SDRJFM_IQ_CF32 = 0,   /* interleaved float32: what getSamples delivers                          */
    SDRJFM_IQ_U8   = 1,   /* rtlsdr: (b - 127) / 128        devices/rtlsdr-handler/rtlsdr-handler.cpp:286-293 */
    SDRJFM_IQ_S8   = 2,   /* hackrf: b / 128                devices/hackrf-handler/hackrf-handler.cpp:355-368 */
    SDRJFM_IQ_S16  = 3,   /* v / denominator: sdrplay 2048|8192 (sdrplay-handler.cpp:266-270,481-488),
                             sdrplay v3 2048|4096 (sdrplay-handler-v3.cpp:254-263,293), pluto 2048
                             (pluto-handler.cpp:574-583), lime 2048                                  */
    SDRJFM_IQ_AIRSPY_S16 = 4  /* airspy: int16 / 2048 at the device's NATIVE rate (sdrjfm_set_native_rate),
                             converted to 2 304 000 samples/s by the handler's per-millisecond linear
                             interpolation (devices/airspy/airspy-handler.cpp:117-128, 283-309), fused
                             into the front-end kernel's read.  n_in / in_pitch count native samples.  */
};
/* native rate of the airspy (a multiple of 1000 Hz; the handler picks the one closest to 2.0 MS/s,
 * airspy-handler.cpp:108-115).  The handle's input_rate must be 2304000.                          */
int  sdrjfm_set_native_rate (sdrjfm_handle *h, int32_t hz);
/* sdrjfm_process / sdrjfm_process_device for samples in a device format; in_pitch in IQ samples.
 * denominator: the int16 divisor (power of two); ignored for the 8-bit formats.               */
int  sdrjfm_process_raw (sdrjfm_handle *h, const void *iq, int32_t format, int32_t denominator,
                         int64_t n_in, int64_t in_pitch,
                         float *audio, int64_t audio_pitch, int64_t *n_audio,
                         float *rds24, int64_t rds_pitch, int64_t *n_rds,
                         sdrjfm_meta *meta /* [n_streams] */);
int  sdrjfm_process_raw_device (sdrjfm_handle *h, const void *d_iq, int32_t format, int32_t denominator,
                                int64_t n_in, int64_t in_pitch,
                                float *d_audio, int64_t audio_pitch, int64_t *n_audio,
                                float *d_rds24, int64_t rds_pitch, int64_t *n_rds);
int  sdrjfm_sync (sdrjfm_handle *h);
int  sdrjfm_get_meta (sdrjfm_handle *h, sdrjfm_meta *meta /* [n_streams] */);
/* copies tap `which` of stream `stream` from the LAST process call to host memory;
 * returns the number of entries (complex entries for complex taps) or a negative status */
int64_t sdrjfm_read_tap (sdrjfm_handle *h, int which, int32_t stream, void *out, int64_t cap);
/* the cudaStream_t the handle launches on (for event timing by the caller)              */
void *sdrjfm_cuda_stream (sdrjfm_handle *h);
/* stage timing: runs only the decimating front end (DC/LO/FIR, the roofline kernel) on
 * device input, without touching stream state.  For bench.py / ncu.                     */
int  sdrjfm_run_frontend_only (sdrjfm_handle *h, const float *d_iq, int64_t n_in, int64_t in_pitch);
int  sdrjfm_run_frontend_only_raw (sdrjfm_handle *h, const void *d_iq, int32_t format, int32_t denominator,
                                   int64_t n_in, int64_t in_pitch);
/* number of kernel launches issued by the handle since creation                          */
int64_t sdrjfm_launch_count (const sdrjfm_handle *h);
/* diagnostics of the parallel-in-time pilot PLL solver for the last process call, per stream:
 * [sum of iterations over windows, max iterations of a window, windows that fell back to the
 * one-lane walk, number of windows]                                                      */
int  sdrjfm_pilot_stats (sdrjfm_handle *h, int32_t *out /* [n_streams][4] */);

/* --- settings: one per fmProcessor setter (includes/fm/fm-processor.h:122-157).  Applied
 * at the next process call, to all streams — the reference applies them at the next
 * 16384-sample block (fm-processor.cpp:397-413).                                         */
int  sdrjfm_set_fm_mode (sdrjfm_handle *h, int32_t mode);          /* setfmMode: 0 Stereo 1 StereoPano 2 Mono */
int  sdrjfm_set_fm_decoder (sdrjfm_handle *h, int32_t decoder);    /* setFMdecoder / fm_Demodulator::setDecoder: 1..6 */
int  sdrjfm_set_sound_mode (sdrjfm_handle *h, int32_t selector);   /* setSoundMode: Channels enum */
int  sdrjfm_set_stereo_panorama (sdrjfm_handle *h, int32_t pan);   /* setStereoPanorama 0..200 */
int  sdrjfm_set_sound_balance (sdrjfm_handle *h, int32_t balance); /* setSoundBalance -100..100 */
int  sdrjfm_set_deemphasis (sdrjfm_handle *h, int32_t usec);       /* setDeemphasis, 1 ("Off") .. 100 us (the GUI offers 1, 50, 75);
                                                                       above: SDRJFM_ERR_UNSUPPORTED */
int  sdrjfm_set_volume_db (sdrjfm_handle *h, float db);            /* setVolume */
int  sdrjfm_set_lf_cutoff (sdrjfm_handle *h, int32_t hz);          /* setlfcutoff (<=0: off) */
int  sdrjfm_set_bandwidth (sdrjfm_handle *h, int32_t hz);          /* setBandwidth ("Off" -> 0) */
int  sdrjfm_set_attenuation (sdrjfm_handle *h, float l, float r);  /* setAttenuation */
int  sdrjfm_set_rds_mode (sdrjfm_handle *h, int32_t mode);         /* setfmRdsSelector: 0 off, 1..3 */
int  sdrjfm_set_local_oscillator (sdrjfm_handle *h, int32_t hz);   /* set_localOscillator */
/* RDS symbol stage on the GPU (optional, off by default): what rdsDecoder::doDecode (src/rds/rds-decoder.cpp:63-102)
 * does with every 24 kHz sample before processBit, for the mode sdrjfm_set_rds_mode selects:
 *   RDS_1  the Costas loop (includes/various/costas.h, ctor args rds-decoder.cpp:40-41) and rdsDecoder_1::doDecode
 *          (src/rds/rds-decoder-1.cpp:126-143);
 *   RDS_2  rdsDecoder_2::doDecode (src/rds/rds-decoder-2.cpp:96-158): matched filter, AGC, timing recovery, Costas;
 *   RDS_3  the same Costas loop and rdsDecoder_3::doDecode (src/rds/rds-decoder-3.cpp:83-118).  That decoder reads the
 *          block synchroniser's error count to re-synchronise its bit clock, so in this mode rdsBlockSynchronizer::pushBit
 *          (src/rds/rds-blocksynchronizer.cpp) and rdsDecoder::processBit (rds-decoder.cpp:104-131) run on the device too.
 * The bits of the last process call are read per stream (sdrjfm_read_rds_bits) and go to rdsDecoder::processBit on the
 * host.  In mode RDS_3 sdrjfm_read_rds_groups returns the groups the device-side synchroniser completed in the last call
 * (blocks A, B, C, D per group: what processBit hands to rdsGroupDecoder::decode) and, in status [4], synchronised /
 * bit-clock re-synchronisations in that call / sync errors / crc errors; group decoding (GUI strings) stays on the host. */
/* startScanning / stopScanning (src/fm/fm-processor.cpp:361-367, 478-495): while scanning, process
 * calls produce no audio and no RDS; per completed block of 1024 fm-rate samples the level around the
 * carrier and at the band edge are returned as (get_db (signal, 256), get_db (noise, 256)) pairs
 * (:886-904); the reference's test is  signal - noise > thresHold  (a constructor argument).        */
int  sdrjfm_set_scanning (sdrjfm_handle *h, int32_t on);
int64_t sdrjfm_read_scan (sdrjfm_handle *h, int32_t stream, float *db_pairs, int64_t cap_pairs);
/* setlfPlotType (src/fm/fm-processor.cpp:244-266; enum ELfPlot, includes/fm/fm-processor.h:84-86): selects the
 * stream the processor pushes into spectrumBuffer_lf / lfBuffer for the LF scope (:565-627, :651-658).
 * -1 (default) = no scope stream, nothing extra is kept.  sdrjfm_read_lf_plot returns, for one stream, the
 * complex samples the LAST process call pushed (interleaved re, im; real-valued types as (x, 0)): one per
 * fm-rate sample, or one per 24 kHz sample for RDS_INPUT / RDS_DEMOD while the RDS branch is on.
 * *sample_rate / *show_full = spectrumSampleRate / showFullSpectrum of the selected type.  The adapter cuts
 * the stream into spectrumSize blocks for ls_scope::processLFSpectrum.  RDS_DEMOD needs the
 * RDS symbol stage.  Host calls (sdrjfm_process) of more than ~1 M samples are cut into pipelined time slices
 * (H2D of slice c+1 under the compute of slice c) — except while a scope stream is selected, the RDS symbol
 * stage is on or the station scan runs: the per-call side outputs (sdrjfm_read_lf_plot, _read_rds_bits,
 * _read_scan) always describe the WHOLE last call.                                                          */
enum sdrjfm_lf_plot {
    SDRJFM_LFPLOT_NONE = -1, SDRJFM_LFPLOT_OFF = 0, SDRJFM_LFPLOT_IF_FILTERED = 1, SDRJFM_LFPLOT_DEMODULATOR = 2,
    SDRJFM_LFPLOT_AF_SUM = 3, SDRJFM_LFPLOT_AF_DIFF = 4, SDRJFM_LFPLOT_AF_MONO_FILTERED = 5,
    SDRJFM_LFPLOT_AF_LEFT_FILTERED = 6, SDRJFM_LFPLOT_AF_RIGHT_FILTERED = 7, SDRJFM_LFPLOT_RDS_INPUT = 8,
    SDRJFM_LFPLOT_RDS_DEMOD = 9
};
int  sdrjfm_set_lf_plot_type (sdrjfm_handle *h, int32_t type);
int64_t sdrjfm_read_lf_plot (sdrjfm_handle *h, int32_t stream, float *out_complex, int64_t cap,
                             int32_t *sample_rate, int32_t *show_full);
/* The LF scope's display spectrum on the GPU (optional): what
 * ls_scope::processLFSpectrum (src/scopes-qwt6/ls-scope.cpp:76-92, window :50-52, mapSpectrum :130-176,
 * add_to_average :178-193) computes from every spectrumSize block of the selected scope stream, for every
 * stream of the batch; blocks run across call boundaries.  sdrjfm_set_lf_spectrum (spectrumSize, displaySize,
 * averageCount as radio.cpp:238-249 passes them; spectrum_size 0 = off); sdrjfm_set_lf_plot_zoom =
 * setlfPlotZoomFactor (fm-processor.cpp:268-271); sdrjfm_read_lf_spectrum copies displayBuffer
 * (display_size doubles) of one stream and reports how many blocks the last call completed.  The first
 * block after a (re)start only primes the average (lfBuffer_newFlag), as in the reference.            */
int  sdrjfm_set_lf_spectrum (sdrjfm_handle *h, int32_t spectrum_size, int32_t display_size, int32_t average_count);
int  sdrjfm_set_lf_plot_zoom (sdrjfm_handle *h, int32_t zoom);
int64_t sdrjfm_read_lf_spectrum (sdrjfm_handle *h, int32_t stream, double *display, int64_t cap,
                                 int32_t *blocks_in_last_call);
/* The HF scope's display spectrum on the GPU (optional): what hs_scope::addElement
 * (src/scopes-qwt6/hs-scope.cpp:102-151, doAverage :175-203) computes from the RAW input the processor copies into
 * hfBuffer (fm-processor.cpp:420): of every segment of inputRate / repeat_rate samples the first 4 display_size are
 * windowed and transformed, mapped to display_size points and averaged with the weight 1 / (repeat_rate / 2).
 * sdrjfm_set_hf_spectrum (displaySize, repeatRate as radio.cpp:232-249 passes them; display_size 0 = off);
 * sdrjfm_read_hf_spectrum copies displayBuffer (display_size doubles, before the widget's dB scaling) of one stream
 * and reports how many segments the last call completed.  Not available for the airspy native-rate format.      */
int  sdrjfm_set_hf_spectrum (sdrjfm_handle *h, int32_t display_size, int32_t repeat_rate);
int64_t sdrjfm_read_hf_spectrum (sdrjfm_handle *h, int32_t stream, double *display, int64_t cap,
                                 int32_t *segments_in_last_call);
int  sdrjfm_set_rds_symbol_stage (sdrjfm_handle *h, int32_t on);
int64_t sdrjfm_read_rds_bits (sdrjfm_handle *h, int32_t stream, uint8_t *bits, int64_t cap);
int64_t sdrjfm_read_rds_groups (sdrjfm_handle *h, int32_t stream, uint16_t *blocks, int64_t cap_groups, int32_t *status);
int  sdrjfm_set_squelch_mode (sdrjfm_handle *h, int32_t mode);     /* set_squelchMode: 0 OFF, 1 NSQ (noise), 2 LSQ (level) */
int  sdrjfm_set_squelch_value (sdrjfm_handle *h, int32_t value);   /* set_squelchValue: 0..100 */
int  sdrjfm_set_auto_mono (sdrjfm_handle *h, int32_t on);          /* setAutoMonoMode */
/* setTestTone (fm-processor.cpp:931-933; insertTestTone :800-823): while on, every working-rate PCM sample is
 * scaled by (1 - 0.9) and a 25 ms burst of a 1 kHz tone at level 0.9 is added every 2 s + 1 sample, behind the
 * fade-in and in front of the peak meter, exactly as the reference's state machine does.                   */
int  sdrjfm_set_test_tone (sdrjfm_handle *h, int32_t on);
/* setDispDelay (:935-937): steps of the peak meter's display delay line (0 .. 512); restarts the line.    */
int  sdrjfm_set_disp_delay (sdrjfm_handle *h, int32_t steps);
/* the (left dB, right dB) pairs the reference would have emitted through showPeakLevel during the LAST process
 * call (one per 961 working-rate samples, counted from the start of the processor), oldest first; returns
 * their number.  At most 1024 - delay read-outs of a call are kept.                                        */
int64_t sdrjfm_read_peak_levels (sdrjfm_handle *h, int32_t stream, float *db_pairs, int64_t cap_pairs);
int  sdrjfm_set_pss_mode (sdrjfm_handle *h, int32_t on);           /* setPSSMode */
int  sdrjfm_set_dc_remove (sdrjfm_handle *h, int32_t on);          /* setDCRemove (also zeroes RfDC) */
int  sdrjfm_trigger_frequency_change (sdrjfm_handle *h);           /* triggerFrequencyChange */
int  sdrjfm_restart_pss_analyzer (sdrjfm_handle *h);               /* restartPssAnalyzer */

/* --- shared tables (tap sets and LUTs) as ONE blob, so that rank 0 can design them and
 * broadcast them (ncclBroadcast / torch.distributed.broadcast) to the other GPUs' ranks
 * (SURVEY.md §8(e)).  export: host copy of the blob this handle uses; import: replace the
 * handle's tables by a blob received from rank 0 (host memory).                          */
/* host-only designer (no device needed): writes the blob for the given rates/filters into
 * out (if cap suffices) and returns its size in bytes, or a negative status.             */
int64_t sdrjfm_design_tables (int32_t input_rate, int32_t fm_rate, int32_t input_filter_hz,
                              int32_t audio_lp_hz, void *out, int64_t cap);
/* host-only designers of the small tables outside the blob (no device needed): which = 0: the RDS_2 matched filter
 * (45 taps at rate a; rds-decoder-2.cpp:67-71); 1: the test-tone burst at working rate a (out [0] = samples of a
 * cycle before the burst, then the burst; fm-processor.cpp:800-823); 2: the second converter a -> b (out [0] = L,
 * out [1] = M, then 32 L taps).  Returns the number of floats written or a negative status.                */
int64_t sdrjfm_design_aux (int32_t which, int32_t a, int32_t b, float *out, int64_t cap);
int64_t sdrjfm_tables_nbytes (const sdrjfm_handle *h);
int  sdrjfm_tables_export (const sdrjfm_handle *h, void *out, int64_t cap);
int  sdrjfm_tables_import (sdrjfm_handle *h, const void *blob, int64_t nbytes);

#ifdef __cplusplus
}
#endif
#endif
