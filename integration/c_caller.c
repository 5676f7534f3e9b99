/* A plain C program on the C ABI (include/sdrjfm_b200.h): no Python, no C++, no torch.
 *   gcc -std=c99 -Iinclude integration/c_caller.c -Lsdr-j-fm_b200 -lsdrjfm_b200 -lm -o c_caller
 * Demodulates 0.25 s of a 1 kHz mono FM tone in two calls and checks the audio: the demodulated tone must
 * sit at 1 kHz with the amplitude the chain's scaling gives (2 pi 75000 / 192000 * 20 / K_FM ~ 1.587).
 * Exit status 0 = ok.  tests/test_integration_compiles.py builds it (CPU) and runs it (GPU). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "sdrjfm_b200.h"

int main (void) {
const int64_t n = 2304000 / 4;
float *iq = (float *)malloc (sizeof (float) * 2 * n);
float *audio = (float *)malloc (sizeof (float) * 2 * (n / 48 + 2));
double phi = 0;
	for (int64_t i = 0; i < n; i ++) {
	   phi += 2 * M_PI * 75000.0 * sin (2 * M_PI * 1000.0 * i / 2304000.0) / 2304000.0;
	   iq [2 * i] = (float)(0.5 * cos (phi)); iq [2 * i + 1] = (float)(0.5 * sin (phi));
	}
sdrjfm_config cfg = { 0 };
	cfg. input_rate = 2304000; cfg. fm_rate = 192000; cfg. working_rate = 48000; cfg. audio_rate = 48000;
	cfg. n_streams = 1; cfg. device = 0; cfg. max_samples_per_call = n; cfg. keep_taps = 0;
int status = 0;
sdrjfm_handle *h = sdrjfm_create (&cfg, &status);
	if (!h) { fprintf (stderr, "sdrjfm_create: status %d: %s\n", status, sdrjfm_last_error (NULL)); return 2; }
	printf ("%s\n", sdrjfm_version ());
	if (sdrjfm_set_fm_mode (h, 2) != SDRJFM_OK || sdrjfm_set_volume_db (h, 0.0f) != SDRJFM_OK) return 3;
int64_t na = 0, total = 0;
sdrjfm_meta meta;
const int64_t half = (n / 2 / 12) * 12;
	if (sdrjfm_process (h, iq, half, half, audio, n / 48 + 2, &na, NULL, 0, NULL, &meta) != SDRJFM_OK) {
	   fprintf (stderr, "sdrjfm_process: %s\n", sdrjfm_last_error (h)); return 4;
	}
	total = na;
	if (sdrjfm_process (h, iq + 2 * half, n - half, n - half, audio + 2 * total, n / 48 + 2 - total, &na, NULL, 0, NULL, &meta) != SDRJFM_OK) {
	   fprintf (stderr, "sdrjfm_process: %s\n", sdrjfm_last_error (h)); return 4;
	}
	total += na;
/*	correlate the second half of the left channel (behind the fade-in ramp's start) with a 1 kHz tone */
double cr = 0, ci = 0, e = 0;
const int64_t a0 = total / 2;
	for (int64_t k = a0; k < total; k ++) {
	   const double v = audio [2 * k] / ((double)(k < 24000 ? k : 24000) / 24000.0);     /* undo the 0.5 s fade-in */
	   cr += v * cos (2 * M_PI * 1000.0 * k / 48000.0); ci += v * sin (2 * M_PI * 1000.0 * k / 48000.0); e += v * v;
	}
const double amp = 2 * sqrt (cr * cr + ci * ci) / (double)(total - a0);
const double rmsv = sqrt (e / (double)(total - a0));
	printf ("audio samples %lld, 1 kHz amplitude %.4f, rms %.4f, dc_rf_db %.2f, launches %lld\n",
	        (long long)total, amp, rmsv, meta. dc_rf_db, (long long)sdrjfm_launch_count (h));
	sdrjfm_destroy (h);
	free (iq); free (audio);
	if (total != n / 48) return 5;
	if (!(amp > 1.2 && amp < 1.7) || fabs (amp / sqrt (2.0) - rmsv) > 0.05 * rmsv) return 6;
	return 0;
}
