/*
 * calm_b200.h -- C ABI of libcalm_b200.so, a B200 (sm_100a) backend for calm's
 * per-token forward() path.
 *
 * The first four functions ARE the reference's CUDA backend boundary: the
 * reference driver declares exactly these prototypes (reference src/run.c:22-25)
 * and the reference backend defines them extern "C" (src/infer.cu:69, 73, 743,
 * 761).  Linking the unmodified reference run.c/tensors.c/tokenizer.c/sampler.c
 * against this library instead of infer.cu gives a working `run` binary -- see
 * INTEGRATION.md.  Everything prefixed calm_b200_ is additive.
 *
 * Error behaviour follows the reference (infer.cu:12-20, 755): there are no
 * error codes; a CUDA failure or an unsupported configuration prints one line
 * to stderr and abort()s.  There is NO CPU fallback: without a usable sm_100
 * device prepare_cuda() aborts.
 *
 * Threading: one host thread, one model per process at a time (the reference
 * keeps global statics too, infer.cu:40-48).  calm_b200_release() makes the
 * library reusable for another model in the same process.
 */
#ifndef CALM_B200_H
#define CALM_B200_H

#include <stddef.h>
#include <stdint.h>

#include "calm_model.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ *
 * Reference boundary (names, arguments and protocol of the reference) *
 * ------------------------------------------------------------------ */

/* Copy `size` bytes of a model tensor to the device and return the device
 * pointer; the caller stores it in struct Weights (run.c:552-560).
 * Replaces reference infer.cu:69-71 (upload_cuda -> cuda_devicecopy :50-55). */
void* upload_cuda(void* host, size_t size);

/* One-time set-up after struct Weights holds device pointers: validates the
 * configuration, allocates activations and the KV cache (honouring
 * state.kvbits, run.c:534-540), allocates host-visible logits and stores them
 * in state.logits, builds the per-token launch plan.
 * Replaces reference infer.cu:73-131. */
void prepare_cuda(struct Transformer* transformer);

/* Run one token at position `pos`.  Returns a host-readable float[vocab_size]
 * valid until the next call (the caller may overwrite it, sampler.c:55-66), or
 * NULL without synchronising when flags & FF_UPDATE_KV_ONLY.
 * Replaces reference infer.cu:743-759 (forward_cuda -> forward<> :651-741). */
float* forward_cuda(struct Transformer* transformer, int token, int pos, unsigned flags);

/* Print the per-stage time / bandwidth table gathered so far (only when
 * profiling was enabled with CALM_B200_PERF=1 or CUDA_INJECTION64_PATH, as in
 * the reference).  Replaces reference infer.cu:761-801. */
void perf_cuda(void);

/* ------------------------------------------------------------------ *
 * Additive entry points                                               *
 * ------------------------------------------------------------------ */

#define CALM_B200_ABI_VERSION 1
int calm_b200_abi_version(void);

/* Select the CUDA device used by upload_cuda/prepare_cuda (default: env
 * CALM_B200_DEVICE, else 0 like reference infer.cu:79).  Call before upload. */
void calm_b200_set_device(int device);

/* Free a pointer returned by upload_cuda(). */
void calm_b200_free(void* device_ptr);

/* Free everything prepare_cuda() created for this transformer (the reference
 * never frees, run.c:636) so another model can be prepared in this process. */
void calm_b200_release(struct Transformer* transformer);

/* Tensor parallelism for models that do not fit (or should not sit on) one GPU -- the 70B shape of
 * BASELINE.json configs[4]; the reference has no counterpart (single device, infer.cu:79).  One process per GPU.
 * Rank 0 obtains a 128-byte NCCL id with calm_b200_tp_unique_id() and hands it to the other ranks by any means
 * (file, pipe, torch.distributed broadcast); every rank then calls calm_b200_tp_init(rank, world, id) BEFORE
 * prepare_cuda() and afterwards drives the library exactly as a single-GPU host does: upload the FULL tensors,
 * prepare, forward the same (token, pos) on every rank.  prepare_cuda() keeps this rank's 1/world of the query /
 * kv heads and of the FFN rows (row ranges of wq/wk/wv/w1/w3 in place, packed column ranges of wo/w2); each
 * forward sums the two partial projections per layer over NVLink (inside the CUDA graph) and every rank returns the full, identical logits.  Requires world | n_heads, n_kv_heads and 32*world |
 * hidden_dim; dense models.  libnccl.so.2 is bound with dlopen on first use.
 * calm_b200_tp_mode(): 0 = not tensor-parallel, 2 = the wo / w2 kernels sum their partials themselves through
 * CUDA-IPC-mapped peer memory (one-shot push all-reduce inside k_matres; default), 1 = ncclAllReduce between
 * kernels (peer mapping unavailable, or env CALM_B200_TP_FUSED=0). */
void calm_b200_tp_unique_id(void* out128);
void calm_b200_tp_init(int rank, int world, const void* id128);
int calm_b200_tp_world(void);
int calm_b200_tp_mode(void);


/* The prompt as ONE batched pass (the reference feeds prompt tokens one at a time through forward(.., FF_UPDATE_KV_ONLY),
 * run.c:206-209, README.md:80): n tokens at positions pos0 .. pos0+n-1 go through every layer together; the projections
 * are tensor-core GEMMs (tcgen05.mma, TMEM accumulators, weights dequantised on the fly, activations as an f16 hi + lo
 * pair so inputs keep 22 bits), attention is causal over the block and the cache.  Afterwards the KV cache holds what n
 * serial KV-only calls would have left (within the stated tolerance); no logits, no synchronisation with the host beyond
 * reading `tokens`.  Returns 1 when the batched pass ran, 0 when the model shape is not served by it (MoE, tensor
 * parallelism, dims not multiples of 128, head_dim other than 64 / 128, a block that would roll the cache over) and the
 * tokens were fed one by one instead -- same result either way.  A host integrates it by replacing the prompt loop:
 * forward_prefill_cuda(t, prompt, n - 1, 0); logits = forward_cuda(t, prompt[n - 1], n - 1, 0);  (INTEGRATION.md). */
int forward_prefill_cuda(struct Transformer* transformer, const int* tokens, int n, int pos0);

/* forward + device-side greedy sample.  Returns argmax(logits) with the
 * reference's tie rule (lowest index, sampler.c:34-42).  The logits are still
 * written to state.logits. */
int calm_b200_forward_argmax(struct Transformer* transformer, int token, int pos);

/* Device-resident greedy decode: feeds token0 at pos0, then n_tokens-1 times the
 * argmax of the previous step, without a host round trip between tokens.
 * out_tokens[i] receives the token sampled after step i (n_tokens values).
 * This is what bench.py times as the in-HBM throughput. */
void calm_b200_decode_greedy(struct Transformer* transformer, int token0, int pos0, int n_tokens, int* out_tokens);

/* The reference's sample() (sampler.c:80-90) on the device: temperature / min-p sampling with the reference's
 * xorshift* generator (sampler.c:7-17), no logits transfer and no host scan of the vocabulary per token.
 * temperature == 0 or minp >= 1 is greedy and leaves *rng_state alone, as in the reference.  Survivors are added
 * in index order like the reference does (exactly its additions up to 2048 survivors; chunk by chunk beyond).
 * decode_sample feeds token0 at pos0 and then n_tokens-1 times its own sample; out_tokens[i] = sample after step i.
 * The logits stay in HBM (calm_b200_read_device_logits copies the last step's); state.logits is not updated. */
void calm_b200_decode_sample(struct Transformer* transformer, int token0, int pos0, int n_tokens, float temperature, float minp,
                             unsigned long long* rng_state, int* out_tokens);
int calm_b200_forward_sample(struct Transformer* transformer, int token, int pos, float temperature, float minp, unsigned long long* rng_state);
void calm_b200_read_device_logits(float* out_vocab);

/* The same sampler kernels on logits supplied by the host (any vocabulary size, no model needed): returns the token,
 * advances *rng_state exactly as sample() does; *device_us (optional) = duration of the two sampler kernels. */
int calm_b200_sample_logits(const float* logits_host, int vocab, float temperature, float minp, unsigned long long* rng_state, float* device_us);

/* Device timer on the library's stream (CUDA events): start, ..., stop -> ms. */
void calm_b200_timer_start(void);
float calm_b200_timer_stop(void);

/* The cudaStream_t all kernels of this library are launched on. */
void* calm_b200_stream(void);

/* Number of kernels this library has launched since load (bench "gpu_launches"). */
uint64_t calm_b200_launch_count(void);

/* Read back one KV-cache entry as floats in the reference's logical order
 * (kv_dim values each, reference infer.c:378-381); for parity tests. */
void calm_b200_read_kv(struct Transformer* transformer, int layer, int kv_pos, float* k_out, float* v_out);

/* Fill the KV cache positions [0, n_pos) of every layer with a deterministic
 * pseudo-random pattern (bench: decode at a late position without running the
 * whole prefix). */
void calm_b200_fill_kv(struct Transformer* transformer, int n_pos, uint64_t seed);

/* Per-stage profiling at run time (what CALM_B200_PERF=1 enables from the start).  While on, every token runs the
 * PRODUCTION CUDA graph (same kernels, programmatic dependent launch) with one extra argument per launch: a stamp slot
 * into which the kernel folds min(start) / max(end) of its CTAs, read from %globaltimer -- the reference's coopstage
 * scheme (infer.cu:390-402), so perf_cuda() describes the path that produces the tokens, not an eager variant.
 * set_perf(1) clears the counters.  stage_stats() returns 0 past the last stage; totals are over all launches of that
 * stage since set_perf(1): milliseconds (first CTA past its dependency wait -> last CTA done), algorithmic bytes (the
 * reference's per-stage accounting, infer.cu:683-699) and launch count.  perf_token_ms(): mean span of a token, first
 * stamp to last (consecutive kernels overlap under PDL, so the stage sums can exceed it). */
void calm_b200_set_perf(int on);
int calm_b200_stage_stats(int stage, char* name, int name_cap, double* ms_total, double* bytes_total, long* launches);
double calm_b200_perf_token_ms(void);
/* (debug) 16 raw %globaltimer stamps the attention kernel of the middle layer left during the last profiled token. */
void calm_b200_debug_stamps(unsigned long long* out16);

/* Stand-alone run of the production matvec kernel: y[d] = W[d,n] . x[n] with W
 * in the `dbits` format at device pointer `w_device`; x and y are HOST arrays.
 * Returns the kernel's device time in milliseconds (mean over `iters` launches
 * after `warmup` launches).  Used by the unit parity tests (vs reference
 * infer.c:209-221) and by the per-kernel roofline bench. */
float calm_b200_matvec(int dbits, const void* w_device, const float* x_host, float* y_host, int n, int d, int warmup, int iters);

#ifdef __cplusplus
}
#endif

#endif /* CALM_B200_H */
