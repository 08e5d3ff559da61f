/*
 * calm_model.h -- data contract of the calm backend boundary (C, C++ and CUDA).
 *
 * The four records below are what a calm host program hands to a backend.  A
 * backend that wants to be a drop-in for calm's CUDA path must agree with the
 * reference on them BYTE FOR BYTE, because the reference driver fills them
 * (reference src/run.c:32-117, 520-582) and only then calls the backend.  The
 * field order, types and array bounds therefore follow reference
 * src/model.h:6-89 exactly; the names are kept so that host code written for
 * calm compiles against this header unchanged.  Layout is pinned by the
 * static assertions at the end of the file (x86-64 / LP64; values measured
 * against the reference header with offsetof in this container).
 *
 * Nothing in this file is executable; see include/calm_b200.h for the entry
 * points.
 */
#ifndef CALM_MODEL_H
#define CALM_MODEL_H

#include <stdbool.h>
#include <stddef.h>

/* array bounds of the per-layer pointer tables (reference model.h:6-7) */
#define MAX_LAYERS 128
#define MAX_EXPERTS 64

/* number of leading cache entries that stay pinned once the context rolls
 * over (StreamingLLM attention sinks; reference model.h:9-10) */
#define KV_SINKS 2

/* Hyper-parameters, read from the .calm "__metadata__" block (run.c:32-69). */
struct Config {
	int dim;          /* width of the residual stream */
	int hidden_dim;   /* FFN inner width (per expert) */
	int head_dim;     /* width of one attention head */
	int n_layers;
	int n_heads;      /* query heads */
	int n_kv_heads;   /* key/value heads (GQA when < n_heads) */
	int vocab_size;
	int seq_len;      /* KV-cache capacity in positions */
	float rope_theta;
	int rotary_dim;   /* per head; dims >= rotary_dim are not rotated */
	int n_experts;    /* 0 for dense models */
	int n_experts_ac; /* experts evaluated per token (0 for dense models) */
	float norm_eps;
	bool act_gelu;    /* tanh-GELU gate instead of SiLU */
	bool norm_ln;     /* mean-subtracting LayerNorm (no bias) instead of RMSNorm */
	bool norm_par;    /* FFN re-uses the attention norm output (no 2nd norm) */
	float qkv_clip;   /* clamp q,k,v to +-qkv_clip; FLT_MAX when absent */
};

/*
 * Weight pointers.  `dbits` tells how to read every `void*`:
 *   16 -> IEEE half, 8 -> fp8 e5m2 (the high byte of a half), 4 -> gf4
 * (uint32 words, 8 weights each: bits 0..7 an e5m2 group scale s, then eight
 * 3-bit codes q_k; w_k = (q_k - 4) * s / -4; reference infer.c:28-40).
 * All matrices are row-major (out_features, in_features).  On the CUDA path
 * every pointer is a DEVICE pointer by the time prepare_cuda() runs
 * (run.c:552-576).
 */
struct Weights {
	int dbits;

	void* token_embedding_table;        /* (vocab_size, dim) */
	float* rms_att_weight[MAX_LAYERS];  /* (dim) */
	float* rms_ffn_weight[MAX_LAYERS];  /* (dim); NULL when norm_par */
	void* wq[MAX_LAYERS];               /* (n_heads*head_dim, dim) */
	void* wk[MAX_LAYERS];               /* (n_kv_heads*head_dim, dim) */
	void* wv[MAX_LAYERS];               /* (n_kv_heads*head_dim, dim) */
	void* wo[MAX_LAYERS];               /* (dim, n_heads*head_dim) */
	void* w1[MAX_LAYERS];               /* ([n_experts,] hidden_dim, dim) */
	void* w2[MAX_LAYERS];               /* ([n_experts,] dim, hidden_dim) */
	void* w3[MAX_LAYERS];               /* ([n_experts,] hidden_dim, dim) */
	float* rms_final_weight;            /* (dim) */
	void* wcls;                         /* (vocab_size, dim); == embedding when tied */
	float* bqkv[MAX_LAYERS];            /* ((n_heads+2*n_kv_heads)*head_dim) or NULL */
	void* moegate[MAX_LAYERS];          /* (n_experts, dim) or NULL */
};

/*
 * Per-sequence state.  Everything except `kvbits` (set by the caller before
 * prepare, run.c:534-540) and `logits` (read by the caller) is owned by and
 * opaque to the backend; this backend keeps its own device-side state and
 * only fills the fields the reference CUDA backend fills (infer.cu:99-112).
 */
struct RunState {
	float* x;
	float* xb;
	float* xb2;
	float* hb;
	float* hb2;
	float* he;
	float* q;
	float* k;
	float* v;
	float* att;
	float* exp;
	float* logits;     /* host-readable float[vocab_size] */
	int kvbits;        /* 16: half cache, 8: e5m2 cache */
	void* key_cache;
	void* value_cache;
};

struct Transformer {
	struct Config config;
	struct Weights weights;
	struct RunState state;
	size_t n_params, n_bytes, n_bandwidth;
	float* (*forward)(struct Transformer* transformer, int token, int pos, unsigned flags);
};

enum ForwardFlags {
	FF_UPDATE_KV_ONLY = 1 << 0, /* advance the KV cache, return NULL, no logits */
};

#if defined(__cplusplus)
#define CALM_LAYOUT_ASSERT(c, m) static_assert(c, m)
#else
#define CALM_LAYOUT_ASSERT(c, m) _Static_assert(c, m)
#endif
#if defined(__x86_64__) || defined(__aarch64__)
CALM_LAYOUT_ASSERT(sizeof(struct Config) == 60, "Config layout drifted from reference model.h");
CALM_LAYOUT_ASSERT(offsetof(struct Config, qkv_clip) == 56, "Config.qkv_clip");
CALM_LAYOUT_ASSERT(sizeof(struct Weights) == 11296, "Weights layout drifted from reference model.h");
CALM_LAYOUT_ASSERT(offsetof(struct Weights, wcls) == 9240, "Weights.wcls");
CALM_LAYOUT_ASSERT(offsetof(struct Weights, moegate) == 10272, "Weights.moegate");
CALM_LAYOUT_ASSERT(sizeof(struct RunState) == 120, "RunState layout drifted from reference model.h");
CALM_LAYOUT_ASSERT(offsetof(struct RunState, kvbits) == 96, "RunState.kvbits");
CALM_LAYOUT_ASSERT(sizeof(struct Transformer) == 11512, "Transformer layout drifted from reference model.h");
CALM_LAYOUT_ASSERT(offsetof(struct Transformer, state) == 11360, "Transformer.state");
CALM_LAYOUT_ASSERT(offsetof(struct Transformer, forward) == 11504, "Transformer.forward");
#endif

#endif /* CALM_MODEL_H */
