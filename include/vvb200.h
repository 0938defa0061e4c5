/* vvb200.h — C ABI of the B200-native VietVoice-TTS synthesis engine (libvvb200.so).
 *
 * This is the drop-in boundary for the reference's three ONNX Runtime sessions.  What each entry point
 * replaces (paths relative to /root/reference):
 *
 *   vv_engine_create / vv_engine_load_blob / vv_engine_finalize / vv_engine_destroy
 *       <- onnxruntime.InferenceSession(model_bytes, sess_options, providers)   vietvoicetts/core/model.py:98-102
 *          ModelSessionManager.load_models / cleanup                            vietvoicetts/core/model.py:131-135,216-221
 *   vv_batch_create / vv_batch_destroy
 *       <- the per-chunk loop state of TTSEngine.synthesize                     vietvoicetts/core/tts_engine.py:225-238
 *   vv_preprocess        <- sessions['preprocess'].run  (3 feeds -> 8 outputs)  vietvoicetts/core/tts_engine.py:133-146
 *   vv_sample            <- sessions['transformer'].run x (nfe_step-1)          vietvoicetts/core/tts_engine.py:148-174
 *   vv_decode            <- sessions['decode'].run      (2 feeds -> 1 output)   vietvoicetts/core/tts_engine.py:176-187
 *   vv_synthesize_batch  <- the whole preprocess -> steps -> decode body of the loop at tts_engine.py:225-238,
 *                           for B independent chunks at once, HOST buffers in, HOST int16 PCM out
 *   vv_prompt_put / vv_prompt_drop / vv_prompt_cache_clear / vv_preprocess_prompt
 *       <- ModelSessionManager.select_sample re-reading the prompt WAV from the tar on every call
 *          (vietvoicetts/core/model.py:204-211) and TTSEngine's never-used sample_cache
 *          (vietvoicetts/core/tts_engine.py:30): prompt PCM, its log-mel and ref_signal_len stay in HBM
 *   vv_crossfade_pcm / vv_batch_crossfade / vv_synthesize_joined
 *       <- AudioProcessor.fix_clipped_audio + concatenate_with_crossfade_improved
 *          (vietvoicetts/core/audio_processor.py:47-58,123-193; called at core/tts_engine.py:244-246)
 *   vv_get_tensor / vv_set_noise  <- numpy arrays handed between session.run calls (tts_engine.py:229-235)
 *   vv_last_error        <- the exception text wrapped at tts_engine.py:256-257 / model.py:125-129
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a negative vv_status
 * otherwise, with a thread-local message available from vv_last_error().  All `host` pointers are host memory;
 * `dev` pointers are device memory on the engine's GPU.  There is NO CPU fallback: without a CUDA device every
 * compute entry point fails with VV_ERR_CUDA.
 */
#ifndef VVB200_H
#define VVB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum vv_status {
  VV_OK = 0,
  VV_ERR_ARG = -1,     /* bad argument / shape                         */
  VV_ERR_CUDA = -2,    /* CUDA runtime or driver error, or no device   */
  VV_ERR_FORMAT = -3,  /* weight blob malformed / tensor missing       */
  VV_ERR_STATE = -4    /* call order (e.g. sample before preprocess)   */
};

/* Architecture constants of the three graphs.  Mirrors ArchConfig (vietvoice-tts_b200/arch.py) field for field:
 * 23 x int32 then 10 x float. */
typedef struct vv_arch {
  int32_t dim, depth, heads, head_dim, ff_dim, n_mel, text_dim, conv_pos_k, conv_pos_groups, time_freq_dim,
      rope_heads, vocab, text_layers, text_ff, pos_table_len, n_fft, hop, sample_rate, voc_dim, voc_ff, voc_layers,
      voc_k, nfe;
  float rope_theta, cfg_strength, sway, ln_eps, target_rms, mel_clamp, mel_fmin, mel_fmax, mag_clip, pcm_scale;
} vv_arch;

typedef struct vv_engine vv_engine;
typedef struct vv_batch vv_batch;

const char* vv_last_error(void);
int vv_version(void);
/* number of CUDA devices visible (0 when none); never fails */
int vv_device_count(void);

/* ---- engine lifetime ------------------------------------------------------------------------------------ */
/* `stream`: a cudaStream_t to run on (e.g. the caller's current stream), or NULL for an engine-owned stream. */
int vv_engine_create(const vv_arch* arch, int device, void* stream, vv_engine** out);
/* Load one weight blob (format: vietvoice-tts_b200/artifact.py).  May be called once per graph. */
int vv_engine_load_blob(vv_engine* e, const void* blob, size_t nbytes);
/* Convert weights to device layouts, build per-NFE time/modulation tables.  Call after all blobs are loaded. */
int vv_engine_finalize(vv_engine* e);
void vv_engine_destroy(vv_engine* e);
/* kernels launched by this engine since creation (for bench.py's gpu_launches) */
int64_t vv_engine_launch_count(const vv_engine* e);
void* vv_engine_stream(const vv_engine* e);

/* ---- one batch of B independent chunks ------------------------------------------------------------------ */
/* total_frames[b] = max_duration of chunk b (prompt + target mel frames, tts_engine.py:117-119). */
int vv_batch_create(vv_engine* e, int B, const int64_t* total_frames, vv_batch** out);
void vv_batch_destroy(vv_batch* b);

/* preprocess graph for chunk `idx`.  audio: int16 [n_samples]; text_ids: int32 [n_ids] exactly as
 * TextProcessor.text_to_indices returns them; noise_or_null: fp32 [T, n_mel] to inject y0 (parity), or NULL to
 * draw N(0,1) from Philox(seed, chunk_key).  ref_len_out receives ref_signal_len. */
int vv_preprocess(vv_batch* b, int idx, const int16_t* audio, int64_t n_samples, const int32_t* text_ids,
                  int64_t n_ids, const float* noise_or_null, uint64_t seed, uint64_t chunk_key,
                  int64_t* ref_len_out);

/* ---- device-resident prompts ------------------------------------------------------------------------------ */
/* Every prompt the engine sees (vv_preprocess, vv_synthesize_batch, vv_prompt_put) is keyed by a 128-bit content hash
 * of its PCM and kept in HBM together with its log-mel and ref_signal_len, in an LRU of VVB200_PROMPT_CACHE entries
 * (default 64; 0 disables caching): all chunks of a long text and all requests for one voice share ONE upload and ONE
 * mel computation.  vv_prompt_put makes a prompt resident ahead of time and returns its id (never 0);
 * vv_preprocess_prompt is vv_preprocess for a resident prompt (no audio argument, no hashing); it fails with
 * VV_ERR_STATE if the prompt has been evicted.  vv_prompt_cache_stats: what[0] = prompts resident, what[1] = uploads
 * (misses) so far, what[2] = hits so far. */
int vv_prompt_put(vv_engine* e, const int16_t* audio, int64_t n_samples, uint64_t* prompt_id_out,
                  int64_t* ref_len_out);
int vv_prompt_drop(vv_engine* e, uint64_t prompt_id);
int vv_prompt_cache_clear(vv_engine* e);
int vv_prompt_cache_stats(vv_engine* e, int64_t* what);
int vv_preprocess_prompt(vv_batch* b, int idx, uint64_t prompt_id, const int32_t* text_ids, int64_t n_ids,
                         const float* noise_or_null, uint64_t seed, uint64_t chunk_key, int64_t* ref_len_out);

/* Run `n_steps` Euler steps of the sampler starting at step index `first_step` on the nfe-point time grid
 * (nfe <= 0: the arch default).  n_steps = nfe-1 with first_step = 0 is the whole loop and is replayed from a
 * CUDA graph; n_steps = 1 is one `transformer` session call.  Asynchronous on the engine stream. */
int vv_sample(vv_batch* b, int nfe, int first_step, int n_steps);

/* decode graph for chunk `idx`: writes (T - ref_len - 1) * hop int16 samples; n_out receives the count. */
int vv_decode(vv_batch* b, int idx, int16_t* pcm_out, int64_t capacity, int64_t* n_out);
/* decode all chunks into device memory and copy back in one go; pcm_out[b] sized by vv_batch_pcm_len */
int vv_decode_all(vv_batch* b, int16_t* const* pcm_out, int64_t* n_out);
int64_t vv_batch_pcm_len(const vv_batch* b, int idx);

/* Tensor taps, host fp32.  names: "noise" [T,n_mel], "cat_mel_text" / "cat_mel_text_drop" [T,n_mel+text_dim],
 * "mel" [ref_len,n_mel], "hidden" [2,T,dim] (residual stream after the last executed step, cond then uncond),
 * "voc_head" [T_tgt, n_fft+2].  Returns the element count written (<= capacity) or a negative status. */
int64_t vv_get_tensor(vv_batch* b, int idx, const char* name, float* out, int64_t capacity);
int vv_set_noise(vv_batch* b, int idx, const float* noise);
/* override the conditioning (the `cat_mel_text*` feeds of a single transformer session call) */
int vv_set_cond(vv_batch* b, int idx, const float* cat_mel_text, const float* cat_mel_text_drop);
/* override ref_signal_len (the second feed of a standalone `decode` session call, tts_engine.py:182-185) */
int vv_set_ref_len(vv_batch* b, int idx, int64_t ref_len);
int vv_sync(vv_engine* e);
/* parity aid: input embedding + the first n_layers DiT blocks of `step` (no Euler update), for vv_get_tensor taps
 * ("x0", "hidden", "qkv", "attn", "hb", "h1b", "ffb", "cond_proj", "v") */
int vv_debug_partial_step(vv_batch* b, int nfe, int step, int n_layers);

/* ---- whole path, host buffers in / host PCM out ---------------------------------------------------------- */
typedef struct vv_request {
  const int16_t* audio;     /* prompt PCM, 24 kHz mono                                  */
  int64_t n_samples;
  const int32_t* text_ids;  /* ids of reference_text + chunk (tts_engine.py:121-122)   */
  int64_t n_ids;
  int64_t total_frames;     /* max_duration                                             */
  const float* noise;       /* optional injected y0 [total_frames, n_mel], else NULL    */
  uint64_t chunk_key;       /* keys the Philox stream when noise == NULL                */
  int16_t* pcm_out;         /* capacity pcm_capacity samples                            */
  int64_t pcm_capacity;
  int64_t n_out;            /* written by the call                                      */
  uint64_t prompt_id;       /* id from vv_prompt_put, or 0: the prompt is found by hashing `audio`; with a
                               resident id `audio` may be NULL                          */
} vv_request;
/* Synchronous: returns when every pcm_out is filled.  Thread-safe on one engine (calls are serialised by the engine's
 * own lock, like every other entry point; vv_last_error is per thread).  Batches are cached per tuple of total_frames
 * in an LRU of VVB200_BATCH_CACHE entries (default 6); their buffers come from the device's stream-ordered memory
 * pool and all batches share one instantiated CUDA graph of the sampling loop per nfe, updated in place. */
int vv_synthesize_batch(vv_engine* e, vv_request* reqs, int B, int nfe, uint64_t seed);

/* ---- cross-fade on the device --------------------------------------------------------------------------- */
/* AudioProcessor.concatenate_with_crossfade_improved (vietvoicetts/core/audio_processor.py:123-193) with its per-chunk
 * fix_clipped_audio (:47-58), bit-exact with the reference's numpy arithmetic: float64 clip fix, float32 level ratio
 * from numpy's pairwise mean of squares (clipped to [0.7, 1.5], only when both RMS values exceed 100), float64
 * cos^2 / sin^2 seam, every cast truncating toward zero.  fade_out / fade_in: the n_fade-point float64 tables exactly as
 * the reference computes them — np.cos(np.linspace(0, pi/2, n_fade)) ** 2 and np.sin(...) ** 2 — passed in because
 * libm and numpy differ in the last bit of cos / sin.  Every chunk must hold at least 2 * n_fade samples and n >= 2
 * (otherwise VV_ERR_ARG: use the host path); the joined wave has sum(len) - n_fade * (n - 1) samples.
 *   vv_crossfade_pcm     host chunks in, host wave out (e.g. waves gathered from several ranks)
 *   vv_batch_crossfade   decoded chunks order[0..n) of a batch (order == NULL: 0..n-1), straight from device PCM
 *   vv_synthesize_joined vv_synthesize_batch for the chunks of ONE text + the cross-fade in request order: only the
 *                        joined wave is copied to the host (a single chunk is returned untouched, as upstream) */
int vv_crossfade_pcm(vv_engine* e, const int16_t* const* waves, const int64_t* lens, int n, const double* fade_out,
                     const double* fade_in, int n_fade, int16_t* pcm_out, int64_t capacity, int64_t* n_out);
int vv_batch_crossfade(vv_batch* b, const int32_t* order, int n, const double* fade_out, const double* fade_in,
                       int n_fade, int16_t* pcm_out, int64_t capacity, int64_t* n_out);
int vv_synthesize_joined(vv_engine* e, vv_request* reqs, int B, int nfe, uint64_t seed, const double* fade_out,
                         const double* fade_in, int n_fade, int16_t* pcm_out, int64_t capacity, int64_t* n_out);

/* Same path with every input already resident in HBM (uploaded by a previous vv_preprocess of each chunk): recomputes
 * mel + text conditioning, restores y0, runs the (nfe-1)-step loop and the decode; PCM stays on the device
 * (fetch with vv_decode).  No host<->device copies and no host synchronisation: this is what bench.py's `value` times. */
int vv_run_resident(vv_batch* b, int nfe);
/* One eager DiT step with CUDA events around every launch; ms_out[8] = milliseconds per kernel class:
 * 0 qkv GEMM, 1 out-proj GEMM, 2 ffn-up GEMM, 3 ffn-down GEMM, 4 attention, 5 LN-modulate, 6 conv_pos, 7 rest. */
int vv_profile_step(vv_batch* b, int nfe, int step, float* ms_out);
/* The resident path with CUDA events at the stage boundaries; ms_out[4] = preprocess (mel of the distinct prompts,
 * text ConvNeXt, conditioning projection) | sampling loop | decode (Vocos + iSTFT) | whole call.  Synchronous. */
int vv_profile_stages(vv_batch* b, int nfe, float* ms_out);

/* ---- kernel-level entry points (device pointers; used by the parity tests and micro-benchmarks) ---------- */
typedef struct vv_gemm_epilogue {
  const float* bias;        /* [N] or NULL                           */
  const float* gate;        /* [N] or NULL                           */
  const float* resid;       /* fp32 [M, ld_resid] or NULL            */
  int32_t ld_resid;
  float* out_f32;           /* fp32 [M, ld_f32] or NULL              */
  int32_t ld_f32;
  void* out_bf16;           /* bf16 [M, ld_bf16] or NULL             */
  int32_t ld_bf16;
  const uint8_t* row_mask;  /* [M] or NULL: 0 -> row written as 0    */
  const int32_t* row_pos;   /* [M] RoPE positions (rope_dim > 0)     */
  int32_t rope_dim, rope_off2;
  int32_t act;              /* 0 none, 1 gelu-tanh, 2 gelu-erf, 3 mish */
} vv_gemm_epilogue;
/* C[M,N] = A[M,K] (bf16, ld lda) * B[N,K]^T (bf16, ld ldb), fused epilogue.  K % 64 == 0.  bn in {64,128,256} = 1-CTA tile
   width; bn = 512 = 256x256 tile on a 2-CTA pair (needs N % 256 == 0); any other value lets the engine choose. */
int vv_gemm_bf16(vv_engine* e, const void* A, int lda, const void* B, int ldb, int M, int N, int K,
                 const vv_gemm_epilogue* epi, int bn);
/* grouped conv over rows (implicit GEMM): X bf16 [M, groups*64], Wt bf16 [groups*taps*64, 64] */
int vv_conv_rows_bf16(vv_engine* e, const void* X, int ldx, const void* Wt, int M, int groups, int taps,
                      const vv_gemm_epilogue* epi);
/* attention over packed rows; seq_off / seq_len are HOST arrays of n_seq entries */
int vv_attention_bf16(vv_engine* e, const void* qkv, void* out, int total_rows, const int32_t* seq_off,
                      const int32_t* seq_len, int n_seq, int heads);
int vv_ln_modulate(vv_engine* e, const float* x, int rows, int dim, const float* shift, const float* scale,
                   float eps, void* out_bf16);

#ifdef __cplusplus
}
#endif
#endif /* VVB200_H */
