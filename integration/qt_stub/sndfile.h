/* COMPILE-CHECK STAND-IN for libsndfile's header (the adapter only forwards the raw-IQ dump file) */
#pragma once
typedef struct SNDFILE_tag SNDFILE;
typedef long long sf_count_t;
extern "C" sf_count_t sf_writef_float (SNDFILE *, const float *, sf_count_t);
