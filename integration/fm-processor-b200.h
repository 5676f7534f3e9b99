/*
 * fm-processor-b200.h — drop-in replacement for includes/fm/fm-processor.h of JvanKatwijk/sdr-j-fm.
 *
 * Same class name, nested types, constructor, public methods and signals as the reference's
 * fmProcessor (includes/fm/fm-processor.h:78-157, 283-293), so RadioInterface, the scopes, the audio
 * sink and the RDS decoder compile and run untouched; the DSP of fmProcessor::run is replaced by
 * libsdrjfm_b200.so (include/sdrjfm_b200.h).  To use it, put this header's directory ahead of
 * includes/fm on the include path as "fm-processor.h" (or rename), drop src/fm/fm-processor.cpp from
 * the build in favour of fm-processor-b200.cpp, and link -lsdrjfm_b200.
 *
 * tests/test_integration_compiles.py compiles this pair against the reference's own headers
 * (device-handler.h, audiosink.h, rds-decoder.h, ringbuffer.h, fm-constants.h) with the Qt stand-ins of
 * integration/qt_stub (Qt itself is not in the build image).
 */
#ifndef __FM_PROCESSOR_B200__
#define __FM_PROCESSOR_B200__

#include <QThread>
#include <QObject>
#include <QString>
#include <sndfile.h>
#include <atomic>
#include <vector>
#include <complex>
#include "fm-constants.h"
#include "ringbuffer.h"
#include "rds-decoder.h"
#include "sdrjfm_b200.h"

class deviceHandler;
class RadioInterface;
class audioSink;
class fm_Demodulator;

class fmProcessor : public QThread {
Q_OBJECT
public:
	enum class FM_Mode { Stereo, StereoPano, Mono };
	enum class ELfPlot { OFF, IF_FILTERED, DEMODULATOR, AF_SUM,
	                     AF_DIFF, AF_MONO_FILTERED, AF_LEFT_FILTERED,
	                     AF_RIGHT_FILTERED, RDS_INPUT, RDS_DEMOD };
	enum class ESqMode { OFF, NSQ, LSQ };
	enum Channels { S_STEREO, S_STEREO_SWAPPED, S_LEFT, S_RIGHT, S_LEFTplusRIGHT,
	                S_LEFTminusRIGHT, S_LEFTminusRIGHT_Test };
	struct SMetaData {
	   enum class EPssState { OFF, ANALYZING, ESTABLISHED };
	   float DcValRf;
	   float DcValIf;
	   float PssPhaseShiftDegree;
	   float PssPhaseChange;
	   EPssState PssState;
	   float GuiPilotStrength;
	   float PilotPllLockStrength;
	   bool  PilotPllLocked;
	};

	fmProcessor (deviceHandler *, RadioInterface *, audioSink *, fm_Demodulator *,
	             int32_t inputRate, int32_t fmRate, int32_t workingRate, int32_t audioRate,
	             int displaySize, int spectrumSize, int32_t repeatRate, int ptyLocale,
	             RingBuffer<std::complex<float>> *hfBuffer,
	             RingBuffer<std::complex<float>> *lfBuffer,
	             RingBuffer<DSPCOMPLEX> *iqBuffer, int16_t thresHold);
	~fmProcessor ();

	void	stop			();
	void	setfmMode		(FM_Mode);
	void	setFMdecoder		(const QString &);
	void	setSoundMode		(uint8_t);
	void	setStereoPanorama	(int16_t);
	void	setSoundBalance		(int16_t);
	void	setDeemphasis		(int16_t);
	void	setVolume		(const float iVolGainDb);
	void	setlfcutoff		(int32_t);
	void	startDumping		(SNDFILE *);
	void	stopDumping		();
	void	setBandwidth		(const QString &);
	void	setAttenuation		(DSPFLOAT, DSPFLOAT);
	void	setfmRdsSelector	(rdsDecoder::ERdsMode);
	void	triggerFrequencyChange	();
	void	restartPssAnalyzer	();
	void	resetRds		();
	void	set_localOscillator	(int32_t);
	void	set_squelchMode		(ESqMode);
	bool	getSquelchState		();
	void	setlfPlotType		(ELfPlot);
	void	setlfPlotZoomFactor	(int32_t);
	bool	isPilotLocked		(float &oLockStrength) const;
	void	setAutoMonoMode		(const bool);
	void	setPSSMode		(const bool);
	void	setDCRemove		(const bool);
	void	new_lfSpectrum		();
	void	setTestTone		(const bool);
	void	setDispDelay		(const int);
	float	get_demodDcComponent	();
	void	startScanning		();
	void	stopScanning		();
	void	set_squelchValue	(int16_t);
	void	set_ptyLocale		(int);

private:
	void	run			() override;

	static constexpr int32_t bufferSize = 16384;          // the reference's pull size, fm-processor.cpp:374
	sdrjfm_handle	*h = nullptr;
	sdrjfm_meta	lastMeta {};
	SMetaData	metaData {};
	deviceHandler	*myRig;
	audioSink	*theSink;
	RingBuffer<std::complex<float>> *hfBuffer;
	RingBuffer<std::complex<float>> *lfBuffer;
	RingBuffer<DSPCOMPLEX> *iqBuffer;
	rdsDecoder	myRdsDecoder;
	std::atomic<rdsDecoder::ERdsMode> rdsModus { rdsDecoder::ERdsMode::RDS_OFF };
	std::vector<std::complex<float>> spectrumBuffer_lf;
	int32_t		inputRate, fmRate, workingRate, audioRate, spectrumSize, repeatRate;
	int32_t		spectrumSampleRate = 0, zoomFactor = 1;
	int		ptyLocale;
	int16_t		thresHold;
	std::atomic<bool> lfBuffer_newFlag { true };
	std::atomic<bool> running { false };
	std::atomic<bool> scanning { false };
	std::atomic<bool> dumping { false };
	SNDFILE		*dumpFile = nullptr;
	float		lastVolumeDb = 0;

signals:
	void	setPLLisLocked		(bool);
	void	hfBufferLoaded		();
	void	lfBufferLoaded		(bool, bool, int);
	void	iqBufferLoaded		();
	void	showMetaData		(const SMetaData *);
	void	scanresult		();
	void	showPeakLevel		(const float, const float);
};
#endif
