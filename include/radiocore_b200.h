/* radiocore_b200.h -- C ABI of the B200-native FM receive engine.
 *
 * Drop-in boundary for the reference's Tuner -> {FM | MFM | WBFM} hot path
 * (luigifcruz/radio-core @ 209dc88).  The reference has no FFI of its own: its
 * boundary is the Python class surface, whose arithmetic is dispatched through
 * radiocore/_internal/injector.py:16-29 to NumPy/SciPy or CuPy/cuSignal.  This
 * library replaces that dispatch outright; each entry point below names the
 * reference method whose arithmetic it performs.  The Python mirror of the
 * classes (radio-core_b200/radiocore) binds these symbols with ctypes -- see
 * INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success or a negative rc_status; it never
 *     throws; rc_last_error() returns a message for the calling thread.
 *   - "dev" pointers are CUDA device pointers on the handle's device (e.g.
 *     torch.Tensor.data_ptr()); `stream` is a cudaStream_t passed as void*
 *     (torch.cuda.current_stream().cuda_stream); NULL = legacy default stream.
 *     Calls are asynchronous on that stream.
 *   - complex samples are interleaved float32 pairs (complex64); audio is
 *     float32.  Block sizes must factor into 2^a 3^b 5^c; FM/MFM/WBFM and
 *     real-input Decimate additionally need even sizes.
 *   - handles own all scratch, tables and carried filter state; a handle is
 *     not re-entrant.
 */
#ifndef RADIOCORE_B200_H
#define RADIOCORE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    RC_OK = 0,
    RC_ERR_INVALID = -1,      /* bad argument (maps to the reference's ValueError) */
    RC_ERR_UNSUPPORTED = -2,  /* size does not factor into 2,3,5 / odd size */
    RC_ERR_CUDA = -3,         /* CUDA runtime failure, see rc_last_error() */
    RC_ERR_STATE = -4         /* call order (e.g. run before commit/load) */
} rc_status;

typedef enum { RC_MODE_FM = 0, RC_MODE_MFM = 1, RC_MODE_WBFM = 2,
               RC_MODE_NONE = 3 /* engine only: channel without demodulator (IQ via rc_engine_channel_iq) */ } rc_mode;

const char* rc_last_error(void);
int rc_version(void);                       /* 10000*major + 100*minor + patch */
int rc_size_supported(int64_t n);           /* 1 if n factors into 2^a 3^b 5^c */

/* ---- multi-channel engine: Tuner + its registered demodulators -----------
 * Tuner.add_channel / request_bandwidth (tuner.py:77-119,163-174) stay on the
 * host; the engine receives the resulting roll (= int(f_in - f_c), tuner.py:153)
 * and sizes.  rc_engine_load = Tuner.load (tuner.py:137-138); rc_engine_run =
 * for every channel Tuner.run (tuner.py:151-161) followed by
 * channel.demodulator.run (fm.py:46-72 / mfm.py:51-71 / wbfm.py:66-105), i.e.
 * the loop body of examples/multi_fm_server.py:98-106, in batched kernels.   */
typedef struct rc_engine rc_engine;
int rc_engine_create(int device, int64_t n_input, rc_engine** out);
int rc_engine_destroy(rc_engine* e);
int rc_engine_add_channel(rc_engine* e, int64_t roll_bins, int64_t bandwidth, int64_t audio_size,
                          int mode, double deemphasis_tau, int* out_index);
int rc_engine_commit(rc_engine* e);
int rc_engine_audio_floats(rc_engine* e, int64_t* total_floats);
int rc_engine_channel_layout(rc_engine* e, int index, int64_t* offset_floats, int64_t* audio_size,
                             int* audio_channels);
int rc_engine_load(rc_engine* e, const void* iq_dev, void* stream);
int rc_engine_run(rc_engine* e, float* audio_dev, void* stream);
int rc_engine_channel_iq(rc_engine* e, int index, void* iq_out_dev, void* stream);
int rc_engine_spectrum(rc_engine* e, void* spectrum_out_dev, void* stream);
int rc_engine_reset_state(rc_engine* e);
int rc_engine_workspace_bytes(rc_engine* e, int64_t* bytes);

/* ---- Tuner.load sharded over the GPUs of one box (radiocore/tools/sharding.py) -------------
 * The reference computes the whole N-point spectrum in one place (tuner.py:137-138).  With G
 * ranks, rank g transforms its commutator branch x[G m + g] (rc_fft_exec, M = N/G points), the
 * ranks exchange pieces over NVLink, rc_subband_combine does the remaining radix-G step, and a
 * second exchange leaves on every rank the contiguous sub-band [x_lo, x_lo + x_len) (cyclic bin
 * indices of the N-bin spectrum) its own channels read.  rc_engine_set_subband (before commit)
 * tells the engine that it will be handed such a sub-band instead of a block;
 * rc_engine_load_subband = Tuner.load for it: no copy, the engine reads `spectrum_dev` during
 * the following rc_engine_run / rc_engine_channel_iq.  rc_engine_load is refused in this mode. */
int rc_engine_set_subband(rc_engine* e, int64_t x_lo, int64_t x_len);
int rc_engine_load_subband(rc_engine* e, const void* spectrum_dev);
int rc_subband_combine(int device, int n_ranks, int64_t piece_len, int64_t n_input, int64_t k0_base,
                       const void* pieces_dev /* [G][P] complex64 */, void* bins_dev /* [G][P] */, void* stream);

/* The same two steps with the exchanges fused into their stores (peer-mapped destination buffers,
 * e.g. torch.distributed._symmetric_memory): rc_fft_exec_scatter writes output element i to
 * piece_bases[i / piece_len][i % piece_len] (batch 1, piece_len even), rc_subband_combine_scatter
 * writes bin (k1, j) of the combine to seg.dst[j - seg.j_lo] for every segment with seg.k1 == k1 and
 * j in [j_lo, j_hi) (at most 4 segments per k1).  Destinations may be NVLink peers' memory. */
typedef struct { int32_t k1; int32_t reserved; int64_t j_lo, j_hi; void* dst; } rc_scatter_seg;
int rc_subband_combine_scatter(int device, int n_ranks, int64_t piece_len, int64_t n_input, int64_t k0_base,
                               const void* pieces_dev, const rc_scatter_seg* segs, int n_segs, void* stream);

/* ---- persistent complex FFT plan (batched, in != out): the local transform of the sharded load */
typedef struct rc_fft rc_fft;
int rc_fft_create(int device, int64_t n, int batch, rc_fft** out);
int rc_fft_destroy(rc_fft* f);
int rc_fft_exec(rc_fft* f, int sign, const void* in_dev, void* out_dev, void* stream);
int rc_fft_exec_scatter(rc_fft* f, int sign, const void* in_dev, void* const* piece_bases, int n_pieces,
                        int64_t piece_len, void* stream);

/* ---- standalone demodulators: FM.run / MFM.run / WBFM.run on `batch` blocks */
typedef struct rc_demod rc_demod;
int rc_demod_create(int device, int mode, int64_t input_size, int64_t output_size,
                    double deemphasis_tau, int batch, rc_demod** out);
int rc_demod_destroy(rc_demod* d);
int rc_demod_run(rc_demod* d, const void* iq_dev, float* audio_dev, void* stream);
int rc_demod_reset_state(rc_demod* d);

/* ---- Decimate.run (decimate.py:35-50): Fourier resampling, Hamming taper --- */
typedef struct rc_decimate rc_decimate;
int rc_decimate_create(int device, int64_t input_size, int64_t output_size, rc_decimate** out);
int rc_decimate_destroy(rc_decimate* d);
int rc_decimate_run_real(rc_decimate* d, const float* in_dev, float* out_dev, void* stream);
int rc_decimate_run_complex(rc_decimate* d, const void* in_dev, void* out_dev, void* stream);

/* ---- Deemphasis.run (deemphasis.py:51-66): stateful 51-tap FIR ------------ */
typedef struct rc_deemph rc_deemph;
int rc_deemph_create(int device, int64_t size, double tau, rc_deemph** out);
int rc_deemph_destroy(rc_deemph* d);
int rc_deemph_run(rc_deemph* d, const float* in_dev, float* out_dev, void* stream);
int rc_deemph_reset_state(rc_deemph* d);
int rc_deemph_taps(double tau, int64_t size, float* taps51_host, float* zi50_host);

/* ---- Bandpass.run (bandpass.py:59-74): firwin taps + zero-phase filtfilt --- */
typedef struct rc_bandpass rc_bandpass;
int rc_bandpass_create(int device, int64_t size, double start_hz, double stop_hz, int num_taps,
                       const char* window, rc_bandpass** out);
int rc_bandpass_destroy(rc_bandpass* b);
int rc_bandpass_run(rc_bandpass* b, const float* in_dev, float* out_dev, void* stream);
int rc_bandpass_taps(rc_bandpass* b, float* taps_host, int capacity);

/* ---- PLL.step / real / image (pll.py:25-58): Hilbert analytic signal ------- */
typedef struct rc_pll rc_pll;
int rc_pll_create(int device, int64_t size, rc_pll** out);
int rc_pll_destroy(rc_pll* p);
int rc_pll_step(rc_pll* p, const float* in_dev, void* stream);
int rc_pll_eval(rc_pll* p, double mult, int imag, float* out_dev, void* stream);

/* ---- per-kernel timing for bench.py: CUDA events around every launch ------ */
int rc_profile_enable(int on);
int rc_profile_reset(void);
int64_t rc_profile_launches(void);          /* kernels launched since the last reset */
int rc_profile_report(char* json_out, int capacity);   /* returns the size needed */

/* ---- test hook: plain batched complex FFT (sign -1 forward, +1 inverse,
 *      unnormalised), used by the parity tests of the FFT engine itself ---- */
int rc_fft_c2c(int device, int64_t n, int batch, int sign, const void* in_dev, void* out_dev,
               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RADIOCORE_B200_H */
