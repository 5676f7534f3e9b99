/*
 * fm-processor-b200.cpp — replaces src/fm/fm-processor.cpp: the fmProcessor QThread pulls blocks from the
 * device handler exactly as the reference does (fm-processor.cpp:387-421) and hands them to
 * libsdrjfm_b200.so; what comes back goes where the reference sends it: PCM to audioSink::putSample
 * (:825-838), the 24 kHz RDS baseband to rdsDecoder::doDecode (:553-563), the LF scope stream to lfBuffer
 * (:651-658), metadata and peak levels to the GUI signals (:662-684, :793).
 * There is no CPU fallback: without a B200 the constructor throws, like a device handler that cannot open
 * its device (devices/device-exceptions.h).
 */
#include <cstdio>
#include <string>
#include <stdexcept>
#include "fm-processor-b200.h"
#include "device-handler.h"
#include "audiosink.h"
#include "radio.h"

#define RDS_RATE 24000                                        // fm-processor.h:46

// the decoder names of fm_Demodulator::listNameofDecoder (src/fm/fm-demodulator.cpp:36-44) -> decoder codes
static int32_t decoderCode (const QString &s) {
static const char *names [] = { "AM", "FM PLL Decoder", "FM Mixed Demod", "FM Complex Baseband Delay",
                                "FM Real Baseband Delay", "FM Difference Based" };
	for (int i = 0; i < 6; i ++)
	   if (s == QString (names [i])) return i + 1;
	return 3;                                              // no match leaves the constructor's MIXED (:66)
}

fmProcessor::fmProcessor (deviceHandler *theDevice, RadioInterface *RI, audioSink *mySink, fm_Demodulator *,
                          int32_t inputRate, int32_t fmRate, int32_t workingRate, int32_t audioRate,
                          int displaySize, int spectrumSize, int32_t repeatRate, int ptyLocale,
                          RingBuffer<std::complex<float>> *hfBuffer, RingBuffer<std::complex<float>> *lfBuffer,
                          RingBuffer<DSPCOMPLEX> *iqBuffer, int16_t thresHold)
	: myRig (theDevice), theSink (mySink), hfBuffer (hfBuffer), lfBuffer (lfBuffer), iqBuffer (iqBuffer),
	  myRdsDecoder (RI, RDS_RATE), inputRate (inputRate), fmRate (fmRate), workingRate (workingRate), audioRate (audioRate),
	  spectrumSize (spectrumSize), repeatRate (repeatRate), ptyLocale (ptyLocale), thresHold (thresHold) {
	(void)displaySize;
sdrjfm_config c = {};
	c. input_rate = inputRate; c. fm_rate = fmRate; c. working_rate = workingRate; c. audio_rate = audioRate;
	c. n_streams = 1; c. device = 0; c. max_samples_per_call = bufferSize; c. keep_taps = 0; c. front_end_mode = 0;
int st = 0;
	h = sdrjfm_create (&c, &st);
	if (h == nullptr)
	   throw std::runtime_error (std::string ("sdrjfm_b200: ") + sdrjfm_last_error (nullptr));
	sdrjfm_set_lf_plot_type (h, (int32_t)ELfPlot::OFF);
	connect (this, SIGNAL (hfBufferLoaded ()), RI, SLOT (hfBufferLoaded ()));
	connect (this, SIGNAL (lfBufferLoaded (bool, bool, int)), RI, SLOT (lfBufferLoaded (bool, bool, int)));
	connect (this, SIGNAL (iqBufferLoaded ()), RI, SLOT (iqBufferLoaded ()));
	connect (this, SIGNAL (showMetaData (const SMetaData *)), RI, SLOT (showMetaData (const SMetaData *)));
	connect (this, SIGNAL (scanresult ()), RI, SLOT (scanresult ()));
	connect (this, SIGNAL (showPeakLevel (const float, const float)), RI, SLOT (showPeakLevel (const float, const float)));
}

fmProcessor::~fmProcessor () {
	stop ();
	sdrjfm_destroy (h);
}

void	fmProcessor::stop () {
	if (running. load ()) {
	   running. store (false);
	   while (!isFinished ()) usleep (100);
	}
}

// every setter forwards 1:1; the library applies it at the next process call, the reference at the next
// 16384-sample block (fm-processor.cpp:397-413)
void	fmProcessor::setfmMode (FM_Mode m)		{ sdrjfm_set_fm_mode (h, (int32_t)m); }
void	fmProcessor::setFMdecoder (const QString &s)	{ sdrjfm_set_fm_decoder (h, decoderCode (s)); }
void	fmProcessor::setSoundMode (uint8_t s)		{ sdrjfm_set_sound_mode (h, s); }
void	fmProcessor::setStereoPanorama (int16_t p)	{ sdrjfm_set_stereo_panorama (h, p); }
void	fmProcessor::setSoundBalance (int16_t b)	{ sdrjfm_set_sound_balance (h, b); }
void	fmProcessor::setDeemphasis (int16_t v)		{ sdrjfm_set_deemphasis (h, v); }
void	fmProcessor::setVolume (const float dB)		{ lastVolumeDb = dB; sdrjfm_set_volume_db (h, dB); }
void	fmProcessor::setlfcutoff (int32_t hz)		{ sdrjfm_set_lf_cutoff (h, hz); }
void	fmProcessor::startDumping (SNDFILE *f)		{ dumpFile = f; dumping. store (f != nullptr); }
void	fmProcessor::stopDumping ()			{ dumping. store (false); }
void	fmProcessor::setBandwidth (const QString &f) {       // "Off" or e.g. "165kHz" (radio.cpp:2099; :232-239)
	if (f == QString ("Off")) { sdrjfm_set_bandwidth (h, 0); return; }
	sdrjfm_set_bandwidth (h, 1000 * std::stoi (f. toStdString ()));
}
void	fmProcessor::setAttenuation (DSPFLOAT l, DSPFLOAT r) { sdrjfm_set_attenuation (h, l, r); }
void	fmProcessor::setfmRdsSelector (rdsDecoder::ERdsMode m) {
	rdsModus. store (m);
	sdrjfm_set_rds_mode (h, (int32_t)m);
}
void	fmProcessor::triggerFrequencyChange ()		{ sdrjfm_trigger_frequency_change (h); resetRds (); }
void	fmProcessor::restartPssAnalyzer ()		{ sdrjfm_restart_pss_analyzer (h); }
void	fmProcessor::resetRds ()			{ myRdsDecoder. reset (); }
void	fmProcessor::set_localOscillator (int32_t lo)	{ sdrjfm_set_local_oscillator (h, lo); }
void	fmProcessor::set_squelchMode (ESqMode m)	{ sdrjfm_set_squelch_mode (h, (int32_t)m); }
bool	fmProcessor::getSquelchState ()			{ return lastMeta. squelch_active != 0; }
void	fmProcessor::setlfPlotType (ELfPlot m)		{ sdrjfm_set_lf_plot_type (h, (int32_t)m); lfBuffer_newFlag. store (true); }
void	fmProcessor::setlfPlotZoomFactor (int32_t z)	{ zoomFactor = z; sdrjfm_set_lf_plot_zoom (h, z); lfBuffer_newFlag. store (true); }
bool	fmProcessor::isPilotLocked (float &oLockStrength) const {
	oLockStrength = lastMeta. pilot_lock_strength;
	return lastMeta. pilot_locked != 0;
}
void	fmProcessor::setAutoMonoMode (const bool b)	{ sdrjfm_set_auto_mono (h, b); }
void	fmProcessor::setPSSMode (const bool b)		{ sdrjfm_set_pss_mode (h, b); }
void	fmProcessor::setDCRemove (const bool b)		{ sdrjfm_set_dc_remove (h, b); }
void	fmProcessor::new_lfSpectrum ()			{ lfBuffer_newFlag. store (true); }
void	fmProcessor::setTestTone (const bool b)		{ sdrjfm_set_test_tone (h, b); }
void	fmProcessor::setDispDelay (const int n)		{ sdrjfm_set_disp_delay (h, n); }
float	fmProcessor::get_demodDcComponent ()		{ return lastMeta. dc_if; }
void	fmProcessor::startScanning ()			{ scanning. store (true); sdrjfm_set_scanning (h, 1); }
void	fmProcessor::stopScanning ()			{ scanning. store (false); sdrjfm_set_scanning (h, 0); }
void	fmProcessor::set_squelchValue (int16_t n)	{ sdrjfm_set_squelch_value (h, n); }
void	fmProcessor::set_ptyLocale (int l)		{ ptyLocale = l; }

void	fmProcessor::run () {
std::vector<std::complex<float>> in (bufferSize);
const int32_t pcmCap = (int32_t)((int64_t)(bufferSize / 48 + 2) * audioRate / workingRate + 4);
std::vector<std::complex<float>> pcm (pcmCap);                 // audio-rate (left, right)
std::vector<std::complex<float>> rds (bufferSize / 96 + 2);   // 24 kHz RDS baseband
std::vector<std::complex<float>> plot (bufferSize / 12 + 2);  // LF scope stream of this pull
std::vector<float> scan (2 * (bufferSize / 12 / 1024 + 2));
std::vector<float> peaks (2 * 8);
int32_t sinceMeta = 0;
	running. store (true);
	while (running. load ()) {
	   while (running. load () && myRig -> Samples () < bufferSize)          // :388-390
	      msleep (1);
	   if (!running. load ()) break;
	   const int32_t amount = myRig -> getSamples (in. data (), bufferSize, IandQ);   // :416-417
	   hfBuffer -> putDataIntoBuffer (in. data (), amount);                  // :420-421
	   emit hfBufferLoaded ();
	   if (dumping. load ())                                                 // :448-455 (raw IQ dump; the reference dumps
	      sf_writef_float (dumpFile, (const float *)in. data (), amount);    //  behind its DC remover)
	   int64_t na = 0, nr = 0;
	   if (sdrjfm_process (h, (const float *)in. data (), amount, amount, (float *)pcm. data (), pcmCap, &na,
	                       (float *)rds. data (), (int64_t)rds. size (), &nr, &lastMeta) != SDRJFM_OK) {
	      fprintf (stderr, "sdrjfm_b200: %s\n", sdrjfm_last_error (h));
	      continue;
	   }
	   if (scanning. load ()) {                                              // :478-495
	      const int64_t nb = sdrjfm_read_scan (h, 0, scan. data (), (int64_t)scan. size () / 2);
	      for (int64_t i = 0; i < nb; i ++)
	         if (scan [2 * i] - scan [2 * i + 1] > thresHold) { emit scanresult (); break; }
	      continue;
	   }
	   for (int64_t i = 0; i < na; i ++)                                     // sendSampletoOutput, :825-838
	      theSink -> putSample (pcm [i]);
	   const int64_t np = sdrjfm_read_peak_levels (h, 0, peaks. data (), (int64_t)peaks. size () / 2);
	   for (int64_t i = 0; i < np; i ++)                                     // :793
	      emit showPeakLevel (peaks [2 * i], peaks [2 * i + 1]);
	   const rdsDecoder::ERdsMode mode = rdsModus. load ();
	   if (mode != rdsDecoder::ERdsMode::RDS_OFF) {                          // :553-563
	      bool any = false;
	      for (int64_t i = 0; i < nr; i ++) {
	         DSPCOMPLEX magCplx;
	         if (myRdsDecoder. doDecode (rds [i], &magCplx, mode, ptyLocale)) {
	            iqBuffer -> putDataIntoBuffer (&magCplx, 1);
	            any = true;
	         }
	      }
	      if (any) emit iqBufferLoaded ();
	   }
//	   LF scope (:565-627, :650-660): the stream of the selected ELfPlot type, cut into spectrumSize blocks
	   int32_t showFull = 0;
	   const int64_t npl = sdrjfm_read_lf_plot (h, 0, (float *)plot. data (), (int64_t)plot. size (),
	                                            &spectrumSampleRate, &showFull);
	   for (int64_t i = 0; i < npl; i ++) {
	      spectrumBuffer_lf. push_back (plot [i]);
	      if ((int32_t)spectrumBuffer_lf. size () >= spectrumSize) {
	         lfBuffer -> putDataIntoBuffer (spectrumBuffer_lf. data (), spectrumSize);
	         emit lfBufferLoaded (showFull != 0, lfBuffer_newFlag. load (), zoomFactor);
	         lfBuffer_newFlag. store (false);
	         spectrumBuffer_lf. resize (0);
	      }
	   }
	   if ((sinceMeta += (int32_t)((int64_t)amount * fmRate / inputRate)) > (fmRate >> 1)) {   // every 500 ms, :662-684
	      sinceMeta = 0;
	      metaData. PilotPllLocked = lastMeta. pilot_locked != 0;
	      metaData. PilotPllLockStrength = lastMeta. pilot_lock_strength;
	      metaData. DcValRf = lastMeta. dc_rf_db;
	      metaData. DcValIf = lastMeta. dc_if;
	      metaData. PssPhaseShiftDegree = lastMeta. pss_phase_shift_deg;
	      metaData. PssPhaseChange = lastMeta. pss_phase_change;
	      metaData. PssState = (SMetaData::EPssState)lastMeta. pss_state;
	      metaData. GuiPilotStrength = 0;          // only set under USE_EXTRACT_LEVELS, which the reference leaves off (fm-processor.h:51)
	      emit showMetaData (&metaData);
	   }
	}
}
