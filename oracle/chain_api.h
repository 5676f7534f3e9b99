/*
 * TEST INFRASTRUCTURE — not product code.
 *
 * Common C API of the two CPU checkers under oracle/:
 *   - oracle/_ref/libsdrjfm_ref.so   : the reference's own DSP classes, compiled from
 *                                      /root/reference where they lie, driven by
 *                                      ref_harness.cpp (prefix  ref_)
 *   - oracle/_build/libsdrjfm_oracle.so : our plain C++ restatement fm_oracle.cpp
 *                                      (prefix  orc_)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load these libraries.
 *
 * The chain restated is fmProcessor::run (src/fm/fm-processor.cpp:461-648) together with
 * process_signal_with_rds (:689-759), with every GUI-pushed setting made explicit.
 */
#ifndef SDRJFM_CHAIN_API_H
#define SDRJFM_CHAIN_API_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct chain_cfg {
    int32_t input_rate;      /* 2304000 (fm-constants.h:35)                          */
    int32_t fm_rate;         /* 192000  (radio.cpp:68)                               */
    int32_t fm_mode;         /* 0 Stereo, 1 StereoPano, 2 Mono (fm-processor.h:107)  */
    int32_t decoder;         /* 1 AM 2 PLL 3 MIXED 4 CBB 5 RBB 6 DIFF (fm-demodulator.cpp:27-32) */
    int32_t sound_sel;       /* Channels enum, 0 = S_STEREO (fm-processor.h:112-114) */
    int32_t rds_on;          /* rdsModus != RDS_OFF                                  */
    int32_t auto_mono;       /* fm-processor.cpp:121                                 */
    int32_t pss_on;          /* fm-processor.cpp:122                                 */
    int32_t dc_remove;       /* fm-processor.cpp:134                                 */
    int32_t input_filter_hz; /* 0 = off, else fmBandwidth in Hz (setBandwidth :232)  */
    int32_t lf_cutoff_hz;    /* 0 = off, else setlfcutoff (:762)                     */
    int32_t lo_hz;           /* set_localOscillator (:865)                           */
    float   lgain, rgain;    /* setAttenuation (:351)                                */
    int32_t deemph_us;       /* setDeemphasis (:291), 50                             */
    float   volume_db;       /* setVolume (:299)                                     */
    int32_t panorama;        /* setStereoPanorama (:277), 100                        */
    int32_t balance;         /* setSoundBalance (:282), 0                            */
    int32_t squelch_mode;    /* set_squelchMode (:882): 0 OFF, 1 NSQ, 2 LSQ (fm-processor.h:87); ref_ only */
    int32_t squelch_value;   /* set_squelchValue (:213), 0..100                      */
} chain_cfg;

/* Output taps. Any pointer may be NULL. Capacities are the caller's business:
 * fm-rate taps need n_in/12 + 1 entries, rds24 needs n_in/96 + 1.                  */
typedef struct chain_taps {
    float   *fm_z;        /* complex, after fmBand_2 (fm-processor.cpp:474)          */
    float   *demod;       /* theDemodulator->demodulate (:497)                       */
    float   *pilot_phase; /* currentPilotPhase (:695)                                */
    uint8_t *locked;      /* pilotRecover.isLocked() (:697)                          */
    float   *pss_delay;   /* pilotDelayPSS after the sample (:716)                   */
    float   *lr;          /* complex (left,right) after the selector (:527-549)      */
    float   *audio192;    /* complex after de-emphasis and gain (:594-595,:630)      */
    float   *rds_cplx;    /* complex rdsDataCplx @fm rate (:754)                     */
    float   *rds24;       /* complex rdsSample @24 kHz (:553)                        */
} chain_taps;

typedef struct chain_meta {   /* SMetaData, fm-processor.h:91-101, as of the last sample */
    float dc_rf_re, dc_rf_im; /* RfDC                                                   */
    float dc_if;              /* fm_afc                                                 */
    float carrier_ampl;       /* am_carr_ampl                                           */
    float pss_phase_shift;    /* pilotDelayPSS                                          */
    float pss_mean_error;     /* pPSS.get_mean_error()                                  */
    int32_t pss_minimized;    /* pPSS.is_error_minimized()                              */
    float pilot_lock_strength;/* pilotRecover.getLockedStrength()                       */
    int32_t pilot_locked;
    int32_t squelch_active;   /* mySquelch.getSquelchActive () (ref_ only)                */
} chain_meta;

#define CHAIN_DECL(P)                                                                   \
    void   *P##_create (const chain_cfg *cfg);                                          \
    void    P##_destroy (void *h);                                                      \
    /* returns fm-rate samples produced; *n_rds24 = 24 kHz samples produced */          \
    int64_t P##_process (void *h, const float *iq, int64_t n_in,                        \
                         const chain_taps *taps, int64_t *n_rds24);                     \
    /* same chain entered AFTER the discriminator: feeds n_fm given demod values (e.g. the   \
       GPU's own demod tap) through process_signal_with_rds .. gain, so that the stages      \
       behind the discriminator can be compared on bit-identical inputs */                  \
    int64_t P##_process_demod (void *h, const float *demod, int64_t n_fm,                   \
                               const chain_taps *taps, int64_t *n_rds24);                   \
    void    P##_get_meta (void *h, chain_meta *m);                                      \
    /* table dumps used to pin the restatement bit for bit */                           \
    int32_t P##_dump_taps (void *h, int which, float *out, int32_t cap);

CHAIN_DECL(ref)
CHAIN_DECL(orc)

/* RDS symbol stage at 24 kHz, mode RDS_1 (src/rds/rds-decoder.cpp:69-82): Costas loop
 * (includes/various/costas.h) + rdsDecoder_1::doDecode (src/rds/rds-decoder-1.cpp:126-143) —
 * the reference's own classes (ref_ only).  Feeds n complex samples, writes the decoded bits
 * (0/1, one byte each) and returns their number.  dump: 0 = match kernel (43 floats),
 * 1 = rdsFilter kernel real parts (21 floats), 2 = sharpFilter: gain, 8 x (A1 A2 B1 B2) (33 floats). */
/* station scan (src/fm/fm-processor.cpp:478-495, 886-904): per block of 1024 fm-rate complex samples,
 * the reference's own Fft_transform followed by a restatement of getSignal / getNoise / get_db
 * (private members of the Qt class fmProcessor).  out: nblocks pairs (signal dB, noise dB).          */
int64_t ref_scan_blocks (const float *fm_z, int64_t n, float *out);
/* run-time setters between process calls (ref_ only): see ref_harness.cpp */
void    ref_update (void *h, const chain_cfg *cfg, int32_t actions);
void   *ref_rds1_create (int32_t rate);
void    ref_rds1_destroy (void *h);
int64_t ref_rds1_process (void *h, const float *rds24, int64_t n, uint8_t *bits, int64_t cap);
int32_t ref_rds1_dump (void *h, int which, float *out, int32_t cap);

/* working-rate post-processing (ref_ only): insertTestTone + evaluatePeakLevel with the display delay line
 * (src/fm/fm-processor.cpp:772-823), restated in ref_harness.cpp.  See there.                               */
void   *ref_post_create (int32_t working_rate);
void    ref_post_destroy (void *h);
void    ref_post_set (void *h, int32_t tone_on, int32_t delay_steps);
int64_t ref_post_process (void *h, const float *pcm, int64_t n, float *pcm_out, float *peaks, int64_t cap_pairs);

/* RDS symbol stage of mode RDS_2 (ref_ only): the reference's own rdsDecoder_2, see ref_harness.cpp */
void   *ref_rds2_create (int32_t rate);
void    ref_rds2_destroy (void *h);
int64_t ref_rds2_process (void *h, const float *rds24, int64_t n, uint8_t *bits, int64_t cap);
int32_t ref_rds2_dump (void *h, float *out, int32_t cap);
/* mode RDS_3 (rds-decoder.cpp:90-98, 104-131): Costas + the reference's rdsDecoder_3, rdsBlockSynchronizer and RDSGroup,
 * sequenced as rdsDecoder::doDecode / processBit do.  groups: [cap_groups][4] blocks A..D of every completed group. */
void   *ref_rds3_create (int32_t rate);
void    ref_rds3_destroy (void *h);
int64_t ref_rds3_process (void *h, const float *rds24, int64_t n, uint8_t *bits, int64_t cap,
                          uint16_t *groups, int64_t cap_groups, int64_t *n_groups, int32_t *n_resync);
/* the synchroniser alone, fed with bits (checks the test signal's checkwords): returns completed groups */
int64_t ref_blocksync_groups (const uint8_t *bits, int64_t n, uint16_t *groups, int64_t cap_groups);

/* HF scope display spectrum (hs-scope.cpp:102-151, 175-203) restated around the reference's Fft_transform (ref_ only) */
int64_t ref_hf_spectrum (const float *x, int64_t n, int32_t displaySize, int32_t sampleRate, int32_t freq,
                         double *display_out, int64_t cap_blocks);

/* `which` for *_dump_taps (complex entries unless noted) */
enum {
    DUMP_FMBAND1 = 0,      /* 25 complex                      */
    DUMP_FMBAND2 = 1,      /* 3 complex                       */
    DUMP_RDSDECIM = 2,     /* 11 complex                      */
    DUMP_INPUT_FILTER_FREQ = 3,  /* 65536 complex (frequency domain filterVector) */
    DUMP_RDS_BP_FREQ = 4,  /* 32768 complex                   */
    DUMP_PSS_LP_FREQ = 5,  /* 2048 complex                    */
    DUMP_AUDIO_LP_FREQ = 6,/* 8192 complex                    */
    DUMP_SINCOS = 7,       /* fm_rate complex (cos, sin)      */
    DUMP_ATAN = 8,         /* 8 x 8193 floats (as 32772 complex slots) PPY PPX PNY PNX NPY NPX NNY NNX */
    DUMP_CONSTS = 9,       /* 8 floats: K_FM deemphAlpha volumeFactor omega gain pssAlpha pssLockAlpha rfDcAlpha */
    DUMP_SQUELCH_IIR = 10  /* ref_ only: squelch filters (squelchClass.cpp:12-21) as floats: high-pass gain, 10 x
                              (A1 A2 B1 B2), low-pass gain, 10 x (A1 A2 B1 B2)  = 82 floats (41 complex slots) */
};

#ifdef __cplusplus
}
#endif
#endif
