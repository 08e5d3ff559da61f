/*
 * calm_oracle.c -- CPU restatement of calm's per-token forward() path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file exists so that the CUDA path in
 * calm_b200/csrc can be checked against an independent statement of the
 * reference algorithm.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product library
 * (libcalm_b200.so) never links, loads or calls anything in oracle/.
 *
 * What it restates: reference src/infer.c (the CPU backend, which the
 * reference author uses as the de-facto oracle for the CUDA backend through
 * CALM_CPU=1, run.c:503-506).  Every function cites the lines it follows.
 * It is written as plain scalar C (no intrinsics, no -ffast-math); OpenMP is used
 * only across independent output rows / heads, so results do not depend on
 * compiler flags or thread count.
 *
 * How it is pinned: the reference has no golden vectors or unit tests
 * (SURVEY.md section 4), so parity is pinned by EXECUTING the reference:
 * oracle/Makefile compiles the unmodified reference sources from
 * /root/reference/src into oracle/_ref/, tests/test_oracle.py runs
 * both on the same seeded synthetic models, and tools/make_golden.py stores
 * reference outputs as fixtures under tests/golden/ for boxes that do not
 * have /root/reference.
 *
 * Two builds (see oracle/Makefile):
 *   libcalm_oracle.so      acc_t = float : the reference's arithmetic
 *                          (fp32 multiply-add, 2x8 lane partial sums like the
 *                          AVX2 path of infer.c:44-140).
 *   libcalm_oracle_f64.so  acc_t = double: same algorithm with every
 *                          reduction carried in double -- a referee that says
 *                          which of two fp32 implementations is nearer the
 *                          exact result when they disagree.
 */
#include <assert.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/calm_model.h"

#ifdef ORACLE_F64
typedef double acc_t;
#define ACC_FMA(a, b, c) ((double)(a) * (double)(b) + (c))
#define ACC_SQRT sqrt
#define ACC_EXP exp
#else
typedef float acc_t;
#define ACC_FMA(a, b, c) fmaf((a), (b), (c))
#define ACC_SQRT sqrtf
#define ACC_EXP expf
#endif

typedef _Float16 half_t; /* KV cache element, reference infer.c:18-26 */

/* ---- weight decoders ---------------------------------------------------- */

/* e5m2 byte -> value: the byte is the high byte of an IEEE half
 * (reference infer.c:28-35). */
float oracle_fp8_to_float(uint8_t v) {
	union {
		uint16_t u;
		half_t h;
	} cvt;
	cvt.u = (uint16_t)(v << 8);
	return (float)cvt.h;
}

/* float -> e5m2 -> float: what the reference CUDA backend's fp8 KV cache does to an entry (KVT = __nv_fp8_e5m2 when
 * kvbits == 8, run.c:537-539; stores infer.cu:476-481 through the __nv_fp8_e5m2(float) constructor = round to nearest
 * even, saturate to the largest finite value 57344; loads are exact).  e5m2: 2 mantissa bits, normal exponents
 * 2^-14..2^15, subnormal quantum 2^-16.  Every e5m2 value is a half, so the cache keeps half_t storage. */
float oracle_round_e5m2(float v) {
	if (v != v) return v;
	float a = fabsf(v);
	if (a >= 57344.f) return copysignf(57344.f, v); /* includes +-inf: SATFINITE */
	int e;
	frexpf(a, &e);  /* a = m * 2^e, m in [0.5, 1) -> leading bit at 2^(e-1) */
	int qe = e - 1 - 2; /* quantum: two bits below the leading one */
	if (qe < -16) qe = -16;
	float q = ldexpf(1.f, qe);
	float r = nearbyintf(a / q) * q; /* exact scaling by a power of two; default rounding mode = nearest even */
	if (r > 57344.f) r = 57344.f;
	return copysignf(r, v);
}

/* gf4 word -> weight k (0..7): (q_k - 4) * (s / -4) with s the e5m2 scale in
 * the low byte (reference infer.c:37-40). */
float oracle_gf4_to_float(uint32_t word, int k) {
	float s = oracle_fp8_to_float((uint8_t)(word & 0xff)) / -4.f;
	return (float)((int)((word >> (8 + k * 3)) & 7) - 4) * s;
}

static float weight_at(int dbits, const void* w, size_t idx) {
	switch (dbits) {
	case 16:
		return (float)((const half_t*)w)[idx];
	case 8:
		return oracle_fp8_to_float(((const uint8_t*)w)[idx]);
	default:
		return oracle_gf4_to_float(((const uint32_t*)w)[idx / 8], (int)(idx % 8));
	}
}

/* ---- dot product -------------------------------------------------------- */

/* One output row.  Follows the summation structure of the reference's AVX2
 * kernels (infer.c:44-98 for fp16/fp8, :100-140 for gf4): element j goes to
 * lane j%8 of accumulator (j/8)%2, each step is a multiply-add, and the
 * lanes are folded acc0+acc1 -> 8 -> 4 -> 1 at the end.  For gf4 the group
 * scale multiplies x first and the 3-bit code is looked up as (q-4)/-4
 * (infer.c:110-131).  n must be a multiple of 16 (32 for gf4), as the
 * reference asserts (infer.c:47, 73, 103). */
static acc_t dot_row(int dbits, const void* w, int n, size_t row, const float* x) {
	acc_t lane[16];
	for (int i = 0; i < 16; ++i) lane[i] = 0;

	if (dbits == 4) {
		const uint32_t* r = (const uint32_t*)w + row * (size_t)n / 8;
		for (int j = 0; j < n; j += 32) {
			for (int g = 0; g < 4; ++g) { /* four groups of 8 -> acc0,acc1,acc0,acc1 */
				uint32_t word = r[j / 8 + g];
				float scale = oracle_fp8_to_float((uint8_t)(word & 0xff));
				for (int k = 0; k < 8; ++k) {
					float code = (float)((int)((word >> (8 + k * 3)) & 7) - 4) / -4.f;
#ifdef ORACLE_F64
					lane[(g & 1) * 8 + k] += (double)code * ((double)x[j + g * 8 + k] * (double)scale);
#else
					lane[(g & 1) * 8 + k] = fmaf(code, x[j + g * 8 + k] * scale, lane[(g & 1) * 8 + k]);
#endif
				}
			}
		}
	} else {
		for (int j = 0; j < n; j += 16) {
			for (int k = 0; k < 16; ++k) {
				float wv = weight_at(dbits, w, row * (size_t)n + j + k);
				lane[k] = ACC_FMA(x[j + k], wv, lane[k]);
			}
		}
	}

	acc_t acc8[8], acc4[4];
	for (int i = 0; i < 8; ++i) acc8[i] = lane[i] + lane[8 + i];
	for (int i = 0; i < 4; ++i) acc4[i] = acc8[i] + acc8[4 + i];
	return (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
}

/* xout[d] = W[d,n] . x[n] (+ b)  -- reference infer.c:209-221. */
void oracle_matvec(int dbits, const void* w, const float* x, const float* b, float* xout, int n, int d) {
	assert(dbits == 4 || dbits == 8 || dbits == 16);
	assert(n % (dbits == 4 ? 32 : 16) == 0);
	/* rows are independent, so the result does not depend on the thread count */
#pragma omp parallel for schedule(static)
	for (int i = 0; i < d; ++i) {
		acc_t val = dot_row(dbits, w, n, (size_t)i, x);
		if (b) val += b[i];
		xout[i] = (float)val;
	}
}

/* x = decode(E[token, :])  -- reference infer.c:335-347. */
void oracle_embed(int dbits, const void* table, int token, int dim, float* x) {
	for (int i = 0; i < dim; ++i) x[i] = weight_at(dbits, table, (size_t)token * dim + i);
}

/* ---- normalisation, RoPE, attention, gate ------------------------------- */

/* RMSNorm, or mean-subtracting LayerNorm without bias when ln
 * (reference infer.c:183-207).  o may alias x. */
void oracle_rmsnorm(float* o, const float* x, const float* weight, int size, float eps, int ln) {
	acc_t mean = 0;
	if (ln) {
		for (int j = 0; j < size; ++j) mean += x[j];
		mean /= size;
	}
	acc_t ss = 0;
	for (int j = 0; j < size; ++j) ss += ((acc_t)x[j] - mean) * ((acc_t)x[j] - mean);
	acc_t var = ss / size;
	acc_t scale = (acc_t)1 / ACC_SQRT(var + eps);
	for (int j = 0; j < size; ++j) o[j] = (float)(((acc_t)x[j] - mean) * scale * weight[j]);
}

/* Rotate adjacent pairs (i, i+1); frequency theta^(-j/rotary_dim) with
 * j = i mod head_dim, zero beyond rotary_dim (reference infer.c:223-236).
 * The angle pos*freq is formed in fp32 exactly as the reference does, so both
 * builds rotate by the same angle; only cos/sin are evaluated in acc_t. */
void oracle_rope(float* vec, int d, int head_dim, int pos, float theta, int rotary_dim) {
	for (int i = 0; i < d; i += 2) {
		int j_head = i % head_dim;
		float freq = j_head >= rotary_dim ? 0.f : 1.0f / powf(theta, (float)j_head / (float)rotary_dim);
		float val = pos * freq;
#ifdef ORACLE_F64
		double fcr = cos((double)val), fci = sin((double)val);
#else
		float fcr = cosf(val), fci = sinf(val);
#endif
		acc_t v0 = vec[i], v1 = vec[i + 1];
		vec[i] = (float)(v0 * fcr - v1 * fci);
		vec[i + 1] = (float)(v0 * fci + v1 * fcr);
	}
}

/* One query head against kv_len cached positions: scaled scores, max-shifted
 * softmax, value mix (reference infer.c:238-267).  kh/vh point at this head's
 * first element of position 0; consecutive positions are kv_dim apart. */
void oracle_attn_head(float* xout, float* atth, const float* qh, const half_t* kh, const half_t* vh, int head_dim, int kv_dim, int kv_len) {
	acc_t score_max = -FLT_MAX;
	for (int t = 0; t < kv_len; ++t) {
		acc_t score = 0;
		for (int j = 0; j < head_dim; ++j) score += (acc_t)qh[j] * (acc_t)(float)kh[(size_t)t * kv_dim + j];
		score /= ACC_SQRT((acc_t)head_dim);
		if (score_max < score) score_max = score;
		atth[t] = (float)score;
	}
	acc_t score_sum = 0;
	for (int t = 0; t < kv_len; ++t) {
		atth[t] = (float)ACC_EXP((acc_t)atth[t] - score_max);
		score_sum += atth[t];
	}
	for (int j = 0; j < head_dim; ++j) {
		acc_t res = 0;
		for (int t = 0; t < kv_len; ++t) res += ((acc_t)atth[t] / score_sum) * (acc_t)(float)vh[(size_t)t * kv_dim + j];
		xout[j] = (float)res;
	}
}

static float act_gelu(float x) { /* reference infer.c:269-271 */
	return 0.5f * x * (1.0f + tanhf(0.797885f * (x + 0.044715f * x * x * x)));
}

static float act_silu(float x) { /* reference infer.c:273-275 */
	return x / (1.0f + expf(-x));
}

/* Top-k routing: repeated arg-max with strict '>' (lowest index wins ties),
 * weights = exp(l - max) renormalised over the selected experts
 * (reference infer.c:277-305). */
void oracle_moe_gate(float* moe_weights, int* moe_experts, const float* x, int d, int active) {
	float max_val = -FLT_MAX;
	for (int j = 0; j < d; ++j)
		if (max_val < x[j]) max_val = x[j];

	uint64_t mask = 0;
	float wsum = 0.0f;
	for (int k = 0; k < active; ++k) {
		int best = -1;
		for (int j = 0; j < d; ++j)
			if ((mask & (1ull << j)) == 0 && (best == -1 || x[j] > x[best])) best = j;
		moe_experts[k] = best;
		wsum += expf(x[best] - max_val);
		mask |= 1ull << best;
	}
	for (int k = 0; k < active; ++k) moe_weights[k] = expf(x[moe_experts[k]] - max_val) / wsum;
}

static float clipf(float x, float v) { /* reference infer.c:307-309 */
	return x < -v ? -v : (x > v ? v : x);
}

/* ---- state ---------------------------------------------------------------- */

/* Allocate activations and an fp16 KV cache laid out [layer][pos][kv_dim]
 * (reference infer.c:142-181).  The reference CPU path insists on kvbits == 16 (:160); kvbits == 8 here restates
 * the reference CUDA path's e5m2 cache (see oracle_round_e5m2) so the device's fp8 cache has a CPU checker too. */
void oracle_prepare(struct Transformer* t) {
	struct Config* p = &t->config;
	struct RunState* s = &t->state;
	int q_dim = p->head_dim * p->n_heads;
	int kv_dim = p->head_dim * p->n_kv_heads;
	int nact = p->n_experts_ac ? p->n_experts_ac : 1;

	s->x = calloc(p->dim, sizeof(float));
	s->xb = calloc(p->dim, sizeof(float));
	s->xb2 = calloc(q_dim > p->dim ? q_dim : p->dim, sizeof(float));
	s->hb = calloc(p->hidden_dim > p->dim ? p->hidden_dim : p->dim, sizeof(float));
	s->hb2 = calloc(p->hidden_dim, sizeof(float));
	s->q = calloc(q_dim, sizeof(float));
	s->k = calloc(kv_dim, sizeof(float));
	s->v = calloc(kv_dim, sizeof(float));
	s->att = calloc((size_t)p->n_heads * p->seq_len, sizeof(float));
	s->exp = calloc(p->n_experts + nact * 2, sizeof(float));
	s->logits = calloc(p->vocab_size, sizeof(float));
	assert(s->kvbits == 16 || s->kvbits == 8);
	s->key_cache = calloc((size_t)p->n_layers * p->seq_len * kv_dim, sizeof(half_t));
	s->value_cache = calloc((size_t)p->n_layers * p->seq_len * kv_dim, sizeof(half_t));
	if (!s->x || !s->xb || !s->xb2 || !s->hb || !s->hb2 || !s->q || !s->k || !s->v || !s->att || !s->exp || !s->logits || !s->key_cache || !s->value_cache) {
		fprintf(stderr, "oracle_prepare: out of memory\n");
		abort();
	}
}

void oracle_release(struct Transformer* t) {
	struct RunState* s = &t->state;
	free(s->x), free(s->xb), free(s->xb2), free(s->hb), free(s->hb2), free(s->q), free(s->k), free(s->v);
	free(s->att), free(s->exp), free(s->logits), free(s->key_cache), free(s->value_cache);
	int kvbits = s->kvbits;
	memset(s, 0, sizeof(*s));
	s->kvbits = kvbits;
}

/* ---- the path ------------------------------------------------------------- */

/* One token.  Order of operations is the reference's (infer.c:311-472):
 * embed 335-347 | per layer: norm 352, q/k/v (+bias) 360-362, clip 365-371,
 * RoPE 374-375, cache write 378-381, sink re-rotation 384-394, attention
 * 397-406, wo 410 + residual 413-415, ffn norm 417-420, gate 425-432, per
 * active expert w1/w3 437-438, gate activation 440-450, w2 452, weighted
 * residual 454-456 | final norm 466, classifier 469.
 * Optional taps (may be NULL) expose per-layer intermediates to kernel-level
 * tests: tap_q  [n_layers][q_dim]  rotated queries,
 *        tap_att[n_layers][q_dim]  attention output before wo,
 *        tap_x  [n_layers][dim]    residual stream after the layer. */
float* oracle_forward_taps(struct Transformer* t, int token, int pos, unsigned flags, float* tap_q, float* tap_att, float* tap_x) {
	struct Config* p = &t->config;
	struct Weights* w = &t->weights;
	struct RunState* s = &t->state;
	assert(w->dbits == 4 || w->dbits == 8 || w->dbits == 16);

	float* x = s->x;
	int dim = p->dim, hidden_dim = p->hidden_dim;
	int q_dim = p->head_dim * p->n_heads;
	int kv_dim = p->head_dim * p->n_kv_heads;
	int kv_mul = p->n_heads / p->n_kv_heads;
	int nact = p->n_experts_ac ? p->n_experts_ac : 1;

	/* rolling cache with attention sinks, infer.c:330-332 */
	int kv_sink = pos >= p->seq_len ? KV_SINKS : 0;
	int kv_pos = kv_sink + (pos - kv_sink) % (p->seq_len - kv_sink);
	int kv_len = pos >= p->seq_len ? p->seq_len : pos + 1;

	oracle_embed(w->dbits, w->token_embedding_table, token, dim, x);

	for (int l = 0; l < p->n_layers; ++l) {
		oracle_rmsnorm(s->xb, x, w->rms_att_weight[l], dim, p->norm_eps, p->norm_ln);

		size_t loff = (size_t)l * p->seq_len * kv_dim;
		half_t* kb = (half_t*)s->key_cache + loff;
		half_t* vb = (half_t*)s->value_cache + loff;

		oracle_matvec(w->dbits, w->wq[l], s->xb, w->bqkv[l], s->q, dim, q_dim);
		oracle_matvec(w->dbits, w->wk[l], s->xb, w->bqkv[l] ? w->bqkv[l] + q_dim : NULL, s->k, dim, kv_dim);
		oracle_matvec(w->dbits, w->wv[l], s->xb, w->bqkv[l] ? w->bqkv[l] + q_dim + kv_dim : NULL, s->v, dim, kv_dim);

		for (int i = 0; i < q_dim; ++i) s->q[i] = clipf(s->q[i], p->qkv_clip);
		for (int i = 0; i < kv_dim; ++i) {
			s->k[i] = clipf(s->k[i], p->qkv_clip);
			s->v[i] = clipf(s->v[i], p->qkv_clip);
		}

		oracle_rope(s->q, q_dim, p->head_dim, pos, p->rope_theta, p->rotary_dim);
		oracle_rope(s->k, kv_dim, p->head_dim, pos, p->rope_theta, p->rotary_dim);
		if (tap_q) memcpy(tap_q + (size_t)l * q_dim, s->q, q_dim * sizeof(float));

		const int kv8 = s->kvbits == 8;
		for (int i = 0; i < kv_dim; ++i) {
			kb[(size_t)kv_pos * kv_dim + i] = (half_t)(kv8 ? oracle_round_e5m2(s->k[i]) : s->k[i]);
			vb[(size_t)kv_pos * kv_dim + i] = (half_t)(kv8 ? oracle_round_e5m2(s->v[i]) : s->v[i]);
		}

		for (int r = 0; r < kv_sink; ++r) {
			for (int i = 0; i < kv_dim; ++i) s->k[i] = (float)kb[(size_t)r * kv_dim + i];
			oracle_rope(s->k, kv_dim, p->head_dim, 1, p->rope_theta, p->rotary_dim);
			for (int i = 0; i < kv_dim; ++i) kb[(size_t)r * kv_dim + i] = (half_t)(kv8 ? oracle_round_e5m2(s->k[i]) : s->k[i]);
		}

#pragma omp parallel for schedule(static)
		for (int h = 0; h < p->n_heads; ++h) {
			oracle_attn_head(s->xb2 + h * p->head_dim, s->att + (size_t)h * p->seq_len, s->q + h * p->head_dim,
			                 kb + (h / kv_mul) * p->head_dim, vb + (h / kv_mul) * p->head_dim, p->head_dim, kv_dim, kv_len);
		}
		if (tap_att) memcpy(tap_att + (size_t)l * q_dim, s->xb2, q_dim * sizeof(float));

		oracle_matvec(w->dbits, w->wo[l], s->xb2, NULL, s->hb, q_dim, dim);
		for (int i = 0; i < dim; ++i) x[i] += s->hb[i];

		if (!p->norm_par) oracle_rmsnorm(s->xb, x, w->rms_ffn_weight[l], dim, p->norm_eps, p->norm_ln);

		float* moe_weights = s->exp + p->n_experts;
		int* moe_experts = (int*)moe_weights + nact;
		if (p->n_experts) {
			oracle_matvec(w->dbits, w->moegate[l], s->xb, NULL, s->exp, dim, p->n_experts);
			oracle_moe_gate(moe_weights, moe_experts, s->exp, p->n_experts, p->n_experts_ac);
		} else {
			moe_weights[0] = 1.0f;
			moe_experts[0] = 0;
		}

		for (int e = 0; e < nact; ++e) {
			size_t esize = (size_t)dim * hidden_dim * (size_t)w->dbits / 8;
			const char* w1 = (const char*)w->w1[l] + moe_experts[e] * esize;
			const char* w2 = (const char*)w->w2[l] + moe_experts[e] * esize;
			const char* w3 = (const char*)w->w3[l] + moe_experts[e] * esize;
			oracle_matvec(w->dbits, w1, s->xb, NULL, s->hb, dim, hidden_dim);
			oracle_matvec(w->dbits, w3, s->xb, NULL, s->hb2, dim, hidden_dim);
			for (int i = 0; i < hidden_dim; ++i) s->hb[i] = (p->act_gelu ? act_gelu(s->hb[i]) : act_silu(s->hb[i])) * s->hb2[i];
			oracle_matvec(w->dbits, w2, s->hb, NULL, s->xb2, hidden_dim, dim);
			for (int i = 0; i < dim; ++i) x[i] += s->xb2[i] * moe_weights[e];
		}
		if (tap_x) memcpy(tap_x + (size_t)l * dim, x, dim * sizeof(float));
	}

	if (flags & FF_UPDATE_KV_ONLY) return NULL;

	oracle_rmsnorm(x, x, w->rms_final_weight, dim, p->norm_eps, p->norm_ln);
	oracle_matvec(w->dbits, w->wcls, x, NULL, s->logits, p->dim, p->vocab_size);
	return s->logits;
}

float* oracle_forward(struct Transformer* t, int token, int pos, unsigned flags) {
	return oracle_forward_taps(t, token, pos, flags, NULL, NULL, NULL);
}

/* Read one cache entry back as floats (kv_dim each) -- test helper matching
 * calm_b200_read_kv() on the device side. */
void oracle_read_kv(struct Transformer* t, int layer, int kv_pos, float* k_out, float* v_out) {
	struct Config* p = &t->config;
	int kv_dim = p->head_dim * p->n_kv_heads;
	size_t off = ((size_t)layer * p->seq_len + kv_pos) * kv_dim;
	for (int i = 0; i < kv_dim; ++i) {
		k_out[i] = (float)((half_t*)t->state.key_cache)[off + i];
		v_out[i] = (float)((half_t*)t->state.value_cache)[off + i];
	}
}

/* Greedy pick: first maximum wins (reference sampler.c:34-42). */
int oracle_argmax(const float* logits, int n) {
	int max_i = -1;
	float max_p = -FLT_MAX;
	for (int i = 0; i < n; ++i) {
		if (logits[i] > max_p) {
			max_i = i;
			max_p = logits[i];
		}
	}
	return max_i;
}

/* xorshift* generator of the reference (sampler.c:7-17). */
static unsigned int oracle_random_u32(unsigned long long* state) {
	*state ^= *state >> 12;
	*state ^= *state << 25;
	*state ^= *state >> 27;
	return (unsigned int)((*state * 0x2545F4914F6CDD1Dull) >> 32);
}

float oracle_random_f32(unsigned long long* state) {
	return (float)(oracle_random_u32(state) >> 8) / 16777216.0f;
}

/* sample(): greedy when temperature == 0 or minp >= 1, else min-p truncated sampling (reference
 * sampler.c:44-90).  Survivors (logit >= max + log(minp) * T) get exp((l - max) / T); the running sum is formed in
 * index order twice: once for the total, once to find the first index whose cumulative sum exceeds coin * total;
 * the last survivor is the fallback.  Unlike the reference, `logits` is left untouched. */
int oracle_sample(const float* logits, int n, float temperature, float minp, unsigned long long* rng_state) {
	if (temperature == 0.0f || minp >= 1.0f) return oracle_argmax(logits, n);
	float coin = oracle_random_f32(rng_state);
	float max_logit = -FLT_MAX;
	for (int i = 0; i < n; ++i) max_logit = logits[i] > max_logit ? logits[i] : max_logit;
	float cutoff = max_logit + logf(minp) * temperature;
	float cum = 0.0f;
	int fallback = 0;
	for (int i = 0; i < n; ++i)
		if (logits[i] >= cutoff) {
			cum += expf((logits[i] - max_logit) / temperature);
			fallback = i;
		}
	float r = coin * cum, cdf = 0.0f;
	for (int i = 0; i < n; ++i)
		if (logits[i] >= cutoff) {
			cdf += expf((logits[i] - max_logit) / temperature);
			if (r < cdf) return i;
		}
	return fallback;
}

/* bench.py's CPU arm: copy a row-major matrix so that the OpenMP thread which will later READ a row (static row split,
 * as the reference's matmul does, infer.c:209-221) is the one that first touches its destination pages. */
void oracle_parallel_copy(void* dst, const void* src, size_t rows, size_t row_bytes) {
#pragma omp parallel for schedule(static)
	for (long long r = 0; r < (long long)rows; ++r) memcpy((char*)dst + (size_t)r * row_bytes, (const char*)src + (size_t)r * row_bytes, row_bytes);
}
