/* COMPILE-CHECK STAND-IN for PortAudio's header (includes/output/audiosink.h holds a stream by pointer) */
#pragma once
typedef void PaStream;
typedef double PaTime;
typedef int PaDeviceIndex;
typedef unsigned long PaSampleFormat;
typedef unsigned long PaStreamCallbackFlags;
typedef struct PaStreamParameters { PaDeviceIndex device; int channelCount; PaSampleFormat sampleFormat;
                                    PaTime suggestedLatency; void *hostApiSpecificStreamInfo; } PaStreamParameters;
typedef struct PaStreamCallbackTimeInfo { PaTime inputBufferAdcTime, currentTime, outputBufferDacTime; } PaStreamCallbackTimeInfo;
