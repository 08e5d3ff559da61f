// persist.cuh -- the "persistent" engine: one cooperative kernel per token that runs the stage code of
// stages.cuh back to back, separated by grid barriers, and keeps HBM busy ACROSS the barriers:
//
//   * grid = one CTA per SM, 16 warps; every warp owns a fixed, strided set of row pairs of every matrix;
//   * weights are read with the same 16-byte streaming loads as the staged kernels (no staging of weights
//     in shared memory: a warp's share of a row pair goes HBM -> registers -> FMA);
//   * the 126 MB L2 is the run-ahead buffer: before a warp enters the grid barrier that ends a stage it asks
//     the L2 for ALL the rows it will read in the NEXT stage (cp.async.bulk.prefetch.L2, SASS UBLKPF:
//     fire-and-forget, no registers; weights and old KV entries never depend on this token's activations),
//     so HBM streams in program order straight through the barriers and the activation staging, and the
//     demand loads of the next stage hit L2 (or a line already in flight); demand loads carry an
//     evict-first policy so that consumed lines make room for the prefetched ones;
//   * the activation vector of a stage is staged once per CTA in shared memory (permuted layout of
//     common.cuh, RMSNorm applied on the way), exactly as the staged kernels do;
//   * attention, epilogues and the arithmetic are those of stages.cuh (reference infer.c:311-472).
//
// Dense models with an fp16 or fp8 KV cache at pos < seq_len; everything else stays on the staged engine.
#pragma once

#include "fused.cuh" // grid_arrive / grid_wait / watchdog
#include "stages.cuh"

#define PERSIST_WARPS 16
#define PERSIST_THREADS (PERSIST_WARPS * 32)

struct PersistLayer {
	const void *wq, *wk, *wv, *wo, *w1, *w2, *w3;
	const float *rms_att, *rms_ffn, *bqkv;
};

struct PersistArgs {
	int dim, hidden, q_dim, kv_dim, head_dim, n_heads, n_kv_heads, n_layers, vocab, seq_len, kv_mul;
	float eps, clip;
	int ln, norm_par, gelu;
	float *x, *xb, *q, *att, *hb, *logits;
	float* attn_partial;
	unsigned* attn_counter;
	void *kc, *vc;
	const float* rope_freq;
	const void* embed;
	const void* wcls;
	const float* rms_final;
	const TokenParams* tp;
	unsigned* bar;
	int* err;
	unsigned long long* perf; // optional [4][8] ns per stage, thread 0 of CTA 0: {busy, barrier wait, activation staging, -}
	float* cand_val;
	int* cand_idx;
	int mode;
	int attn_nsplit, attn_hg, attn_qgroups, attn_lpp;
	float inv_sqrt_hd;
	int pf_pairs, pf_bytes; // L2 run-ahead per warp at a stage end: row pairs, bytes per row
};

__constant__ PersistLayer c_persist_layers[MAX_LAYERS];

// ---------------------------------------------------------------- split-phase row-pair matvec

template <int DBITS>
struct RowBatch { // U vectors per lane of two rows: 2 * U * 512 bytes in flight per warp
	static constexpr int U = DBITS == 4 ? 4 : 8;
	uint4 w[U][2];
};

// 16-byte weight load whose L2 line becomes the first candidate for eviction: the line was brought in by a
// prefetch, is used exactly once, and should make room for the lines being prefetched for later stages
__device__ __forceinline__ uint4 ldg_stream_ef(const uint4* p, uint64_t pol) {
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
	return r;
}

template <int DBITS>
__device__ __forceinline__ void batch_issue(RowBatch<DBITS>& b, const uint4* r0, const uint4* r1, int nvec, int v0, uint64_t pol) {
#pragma unroll
	for (int u = 0; u < RowBatch<DBITS>::U; ++u) {
		int v = v0 + 32 * u;
		b.w[u][0] = v < nvec ? ldg_stream_ef(r0 + v, pol) : make_uint4(0, 0, 0, 0);
		b.w[u][1] = v < nvec ? ldg_stream_ef(r1 + v, pol) : make_uint4(0, 0, 0, 0);
	}
}

template <int DBITS>
__device__ __forceinline__ void batch_consume(const RowBatch<DBITS>& b, int nvec, int v0, const float4* __restrict__ xs4, float& acc0, float& acc1) {
	constexpr int Q = WFmt<DBITS>::VW / 4;
	const int lane = threadIdx.x & 31;
#pragma unroll
	for (int u = 0; u < RowBatch<DBITS>::U; ++u) {
		int v = v0 + 32 * u;
		if (v < nvec) {
			float4 xv[Q];
			const float4* xp = xs4 + (size_t)(v >> 5) * Q * 32 + lane;
#pragma unroll
			for (int q = 0; q < Q; ++q) xv[q] = xp[q * 32];
			acc0 = dot_vec<DBITS>(b.w[u][0], xv, acc0);
			acc1 = dot_vec<DBITS>(b.w[u][1], xv, acc1);
		}
	}
}

// one lane: ask the L2 for the head of two rows (pieces of <= 16 KB)
__device__ __forceinline__ void prefetch_pair(const void* r0, const void* r1, int rowbytes) {
	for (int off = 0; off < rowbytes; off += 16384) {
		uint32_t n = (uint32_t)min(16384, rowbytes - off);
		l2_prefetch((const char*)r0 + off, n);
		l2_prefetch((const char*)r1 + off, n);
	}
}

// Two rows against the staged vector.
template <int DBITS>
__device__ __forceinline__ void warp_pair(const uint4* r0, const uint4* r1, int nvec, const float4* __restrict__ xs4, uint64_t pol, float& o0, float& o1) {
	constexpr int U = RowBatch<DBITS>::U;
	const int lane = threadIdx.x & 31;
	float a0 = 0.f, a1 = 0.f;
	for (int v0 = lane; v0 < nvec; v0 += 32 * U) {
		RowBatch<DBITS> b;
		batch_issue<DBITS>(b, r0, r1, nvec, v0, pol);
		batch_consume<DBITS>(b, nvec, v0, xs4, a0, a1);
	}
	o0 = warp_sum(a0), o1 = warp_sum(a1);
}

// attention stage as its own function: its register needs are unrelated to the matvec loops
template <typename KVT, int HH>
__device__ __noinline__ void persist_attention(const PersistArgs& a, const void* kc_l, const void* vc_l, int kv_len, float* scratch, int* flag) {
	AttnArgs aa;
	aa.q = a.q, aa.kc = kc_l, aa.vc = vc_l, aa.partial = a.attn_partial, aa.counter = a.attn_counter, aa.out = a.att, aa.tp = a.tp;
	aa.head_dim = a.head_dim, aa.seq_len = a.seq_len, aa.nsplit = a.attn_nsplit, aa.lpp = a.attn_lpp;
	aa.kv_mul = a.kv_mul, aa.qgroups = a.attn_qgroups, aa.inv_sqrt_hd = a.inv_sqrt_hd;
	const int unit = blockIdx.x / a.attn_nsplit, split = blockIdx.x % a.attn_nsplit;
	if constexpr (HH == 1 || HH == 2 || HH == 4) { // transposing score path (head_dim 128 / 64, even head groups)
		if (a.attn_lpp == 16 && (a.attn_hg % 2 == 0 || a.attn_hg == 1)) {
			attn_item<KVT, HH, (HH == 4 ? 2 : 4), false, 16>(aa, a.attn_hg, 2, unit, split, kv_len, scratch, flag);
			return;
		}
		if (a.attn_lpp == 8 && (a.attn_hg % 2 == 0 || a.attn_hg == 1)) {
			attn_item<KVT, HH, (HH == 4 ? 2 : 4), false, 8>(aa, a.attn_hg, 2, unit, split, kv_len, scratch, flag);
			return;
		}
	}
	attn_item<KVT, HH, 4, false>(aa, a.attn_hg, 2, unit, split, kv_len, scratch, flag);
}

// ---------------------------------------------------------------- the kernel

// HH = query heads per warp in the attention stage (the 16 warps form two sets that split the unit's heads)
template <int DBITS, typename KVT, int HH>
__global__ void __launch_bounds__(PERSIST_THREADS, 1) k_persist(const __grid_constant__ PersistArgs a) {
	extern __shared__ __align__(16) float smem[];
	__shared__ float rope_cos[128], rope_sin[128];
	__shared__ float bval[PERSIST_WARPS];
	__shared__ int bidx[PERSIST_WARPS];
	__shared__ int flag;
	float* red = smem;
	float* xs = smem + 32;
	const float4* xs4 = reinterpret_cast<const float4*>(xs);

	constexpr int VW = WFmt<DBITS>::VW;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int NW = gridDim.x * PERSIST_WARPS;       // warps in the grid
	const int gw = blockIdx.x * PERSIST_WARPS + warp; // this warp's index: it owns row pairs gw, gw + NW, ...
	const TokenParams tp = *a.tp;
	const bool leader = threadIdx.x == 0;
	const bool timer = a.perf && blockIdx.x == 0 && leader;

	const int nv_dim = a.dim / VW, nv_q = a.q_dim / VW, nv_hid = a.hidden / VW;
	const int np_qkv = (a.q_dim + 2 * a.kv_dim) / 2, np_o = a.dim / 2, np_up = a.hidden, np_out = (a.vocab + 1) / 2;
	const size_t kv_layer = (size_t)a.n_kv_heads * a.seq_len * a.head_dim;

	// RoPE angles of this token, once (host-libm frequencies, reference infer.c:225-229)
	for (int i = threadIdx.x; i < a.head_dim / 2; i += blockDim.x) sincosf((float)tp.pos * a.rope_freq[i], &rope_sin[i], &rope_cos[i]);
	// x = decode(E[token]) (reference infer.c:335-347): every CTA writes the whole row (identical values) and
	// therefore only reads back its own writes in stage 1 -- no grid barrier needed here
	for (int i = threadIdx.x; i < a.dim; i += blockDim.x) __stcg(a.x + i, weight_at<DBITS>(a.embed, (size_t)tp.token * a.dim + i));
	__syncthreads();

	unsigned long long t_stage = timer ? globaltimer_ns() : 0ull;
	auto barrier = [&](int st, int code) { // publish this CTA's results, wait for all CTAs
		__syncthreads();
		if (leader) {
			unsigned long long t1 = timer ? globaltimer_ns() : 0ull;
			unsigned old = grid_arrive(a.bar);
			grid_wait(a.bar, old, a.err, code);
			if (timer) {
				unsigned long long t2 = globaltimer_ns();
				a.perf[st] += t1 - t_stage, a.perf[8 + st] += t2 - t1;
				t_stage = t2;
			}
		}
		__syncthreads();
	};

	// row pointers of pair p of the q|k|v stage
	auto qkv_rows = [&](const PersistLayer& L, int p, const uint4*& r0, const uint4*& r1, int& which, int& k) {
		int j = 2 * p;
		const void* w;
		if (j < a.q_dim) {
			w = L.wq, k = j, which = 0;
		} else if (j < a.q_dim + a.kv_dim) {
			w = L.wk, k = j - a.q_dim, which = 1;
		} else {
			w = L.wv, k = j - a.q_dim - a.kv_dim, which = 2;
		}
		r0 = reinterpret_cast<const uint4*>(w) + (size_t)k * nv_dim;
		r1 = r0 + nv_dim;
	};

	// bytes of a row / number of row pairs that a warp asks the L2 for ahead of a stage
	const int rb_dim = min(nv_dim * 16, a.pf_bytes) & ~15, rb_q = min(nv_q * 16, a.pf_bytes) & ~15, rb_hid = min(nv_hid * 16, a.pf_bytes) & ~15;
	const int pf_stop = gw + a.pf_pairs * NW; // first pair index beyond the run-ahead window
	const uint64_t pol = l2_policy_evict_first();
	if (lane == 0)
		for (int p = gw; p < np_qkv && p < pf_stop; p += NW) {
			const uint4 *r0, *r1;
			int which, k;
			qkv_rows(c_persist_layers[0], p, r0, r1, which, k);
			prefetch_pair(r0, r1, rb_dim);
		}

	for (int l = 0; l < a.n_layers; ++l) {
		const PersistLayer& L = c_persist_layers[l];
		KVT* kc_l = reinterpret_cast<KVT*>(a.kc) + l * kv_layer;
		KVT* vc_l = reinterpret_cast<KVT*>(a.vc) + l * kv_layer;

		// ---------------- stage 1: norm -> q,k,v (+bias, clip, RoPE) -> q vector / cache append (infer.c:352-381)
		stage_vector<DBITS>(xs, red, a.x, a.dim, L.rms_att, a.eps, a.ln != 0, (a.norm_par && blockIdx.x == 0) ? a.xb : nullptr);
		if (timer) a.perf[16 + 1] += globaltimer_ns() - t_stage;
		for (int p = gw; p < np_qkv; p += NW) {
			const uint4 *r0, *r1;
			int which, k;
			qkv_rows(L, p, r0, r1, which, k);
			float v0, v1;
			warp_pair<DBITS>(r0, r1, nv_dim, xs4, pol, v0, v1);
			if (lane == 0) {
				int j = 2 * p;
				if (L.bqkv) v0 += L.bqkv[j], v1 += L.bqkv[j + 1];
				v0 = fminf(fmaxf(v0, -a.clip), a.clip);
				v1 = fminf(fmaxf(v1, -a.clip), a.clip);
				if (which < 2) {
					int i = (k % a.head_dim) >> 1;
					float fcr = rope_cos[i], fci = rope_sin[i];
					float t0 = v0 * fcr - v1 * fci, t1 = v0 * fci + v1 * fcr;
					v0 = t0, v1 = t1;
				}
				if (which == 0) {
					__stcg(reinterpret_cast<float2*>(a.q + k), make_float2(v0, v1));
				} else {
					KVT* dst = (which == 1 ? kc_l : vc_l) + ((size_t)(k / a.head_dim) * a.seq_len + tp.kv_pos) * a.head_dim + (k % a.head_dim);
					kv_store(dst, v0);
					kv_store(dst + 1, v1);
				}
			}
		}
		if (lane == 0) {
			// next matrix stage: wo; and this CTA's slice of K/V for the attention stage (positions before this token)
			for (int p = gw; p < np_o && p < pf_stop; p += NW) prefetch_pair(reinterpret_cast<const uint4*>(L.wo) + (size_t)(2 * p) * nv_q, reinterpret_cast<const uint4*>(L.wo) + (size_t)(2 * p + 1) * nv_q, rb_q);
			const int units = a.n_kv_heads * a.attn_qgroups;
			if ((int)blockIdx.x < units * a.attn_nsplit) {
				const int unit = blockIdx.x / a.attn_nsplit, split = blockIdx.x % a.attn_nsplit, kvh = unit / a.attn_qgroups;
				const int chunk = (tp.kv_len + a.attn_nsplit - 1) / a.attn_nsplit;
				const int t0 = split * chunk, t1 = min(tp.kv_len, t0 + chunk);
				const int per = (t1 - t0 + PERSIST_WARPS - 1) / PERSIST_WARPS; // positions per warp
				const int ta = t0 + warp * per, tb = min(t1, ta + per);
				if (tb > ta) {
					const size_t off = ((size_t)kvh * a.seq_len + ta) * a.head_dim;
					uint32_t n = (uint32_t)((size_t)(tb - ta) * a.head_dim * sizeof(KVT)) & ~15u;
					if (n > 16384u) n = 16384u;
					if (n) l2_prefetch(kc_l + off, n), l2_prefetch(vc_l + off, n);
				}
			}
		}
		barrier(1, 301);

		// ---------------- stage 2: attention over the cache (stages.cuh attn_item; scratch = the staging area)
		if ((int)blockIdx.x < a.n_kv_heads * a.attn_qgroups * a.attn_nsplit) persist_attention<KVT, HH>(a, kc_l, vc_l, tp.kv_len, smem, &flag);
		barrier(2, 302);

		// ---------------- stage 3: x += wo . att (infer.c:410-415)
		stage_vector<DBITS>(xs, red, a.att, a.q_dim, nullptr, 0.f, false, nullptr);
		if (timer) a.perf[16 + 3] += globaltimer_ns() - t_stage;
		for (int p = gw; p < np_o; p += NW) {
			const uint4* r0 = reinterpret_cast<const uint4*>(L.wo) + (size_t)(2 * p) * nv_q;
			float v0, v1;
			warp_pair<DBITS>(r0, r0 + nv_q, nv_q, xs4, pol, v0, v1);
			if (lane == 0) {
				float2* dst = reinterpret_cast<float2*>(a.x + 2 * p);
				float2 cur = __ldcg(dst);
				__stcg(dst, make_float2(cur.x + v0, cur.y + v1));
			}
		}
		if (lane == 0) // next: w1 | w3
			for (int p = gw; p < np_up && p < pf_stop; p += NW) prefetch_pair(reinterpret_cast<const uint4*>(L.w1) + (size_t)p * nv_dim, reinterpret_cast<const uint4*>(L.w3) + (size_t)p * nv_dim, rb_dim);
		barrier(3, 303);

		// ---------------- stage 4: norm -> act(w1 . xn) * (w3 . xn) (infer.c:417-450)
		if (a.norm_par)
			stage_vector<DBITS>(xs, red, a.xb, a.dim, nullptr, 0.f, false, nullptr);
		else
			stage_vector<DBITS>(xs, red, a.x, a.dim, L.rms_ffn, a.eps, a.ln != 0, nullptr);
		if (timer) a.perf[16 + 4] += globaltimer_ns() - t_stage;
		for (int p = gw; p < np_up; p += NW) {
			float v1, v3;
			warp_pair<DBITS>(reinterpret_cast<const uint4*>(L.w1) + (size_t)p * nv_dim, reinterpret_cast<const uint4*>(L.w3) + (size_t)p * nv_dim, nv_dim, xs4, pol, v1, v3);
			if (lane == 0) __stcg(a.hb + p, (a.gelu ? act_gelu(v1) : act_silu(v1)) * v3);
		}
		if (lane == 0) // next: w2
			for (int p = gw; p < np_o && p < pf_stop; p += NW) prefetch_pair(reinterpret_cast<const uint4*>(L.w2) + (size_t)(2 * p) * nv_hid, reinterpret_cast<const uint4*>(L.w2) + (size_t)(2 * p + 1) * nv_hid, rb_hid);
		barrier(4, 304);

		// ---------------- stage 5: x += w2 . hb (infer.c:452-456)
		stage_vector<DBITS>(xs, red, a.hb, a.hidden, nullptr, 0.f, false, nullptr);
		if (timer) a.perf[16 + 5] += globaltimer_ns() - t_stage;
		for (int p = gw; p < np_o; p += NW) {
			const uint4* r0 = reinterpret_cast<const uint4*>(L.w2) + (size_t)(2 * p) * nv_hid;
			float v0, v1;
			warp_pair<DBITS>(r0, r0 + nv_hid, nv_hid, xs4, pol, v0, v1);
			if (lane == 0) {
				float2* dst = reinterpret_cast<float2*>(a.x + 2 * p);
				float2 cur = __ldcg(dst);
				__stcg(dst, make_float2(cur.x + v0, cur.y + v1));
			}
		}
		// next: q|k|v of the next layer, or the classifier
		if (lane == 0) {
			if (l + 1 < a.n_layers) {
				for (int p = gw; p < np_qkv && p < pf_stop; p += NW) {
					const uint4 *r0, *r1;
					int which, k;
					qkv_rows(c_persist_layers[l + 1], p, r0, r1, which, k);
					prefetch_pair(r0, r1, rb_dim);
				}
			} else if (a.mode != 0) { // the classifier is larger than L2: only the first pairs of every warp
				for (int p = gw; p < np_out && p < pf_stop; p += NW) {
					int ra = 2 * p, rb = min(2 * p + 1, a.vocab - 1);
					prefetch_pair(reinterpret_cast<const uint4*>(a.wcls) + (size_t)ra * nv_dim, reinterpret_cast<const uint4*>(a.wcls) + (size_t)rb * nv_dim, rb_dim);
				}
			}
		}
		barrier(5, 305);
	}

	if (a.mode == 0) return;

	// ---------------- classifier: logits = wcls . norm(x), greedy candidates (infer.c:466-469, sampler.c:34-42)
	stage_vector<DBITS>(xs, red, a.x, a.dim, a.rms_final, a.eps, a.ln != 0, nullptr);
	float best = -FLT_MAX;
	int besti = 0x7fffffff;
	for (int p = gw; p < np_out; p += NW) {
		int ra = 2 * p, rb = min(2 * p + 1, a.vocab - 1);
		float v0, v1;
		warp_pair<DBITS>(reinterpret_cast<const uint4*>(a.wcls) + (size_t)ra * nv_dim, reinterpret_cast<const uint4*>(a.wcls) + (size_t)rb * nv_dim, nv_dim, xs4, pol, v0, v1);
		if (lane == 0) {
			a.logits[ra] = v0;
			if (v0 > best) best = v0, besti = ra;
			if (ra + 1 < a.vocab) {
				a.logits[ra + 1] = v1;
				if (v1 > best) best = v1, besti = ra + 1;
			}
		}
	}
	if (a.cand_val) {
		if (lane == 0) bval[warp] = best, bidx[warp] = besti;
		__syncthreads();
		if (leader) {
			for (int w = 1; w < PERSIST_WARPS; ++w)
				if (bval[w] > best || (bval[w] == best && bidx[w] < besti)) best = bval[w], besti = bidx[w];
			a.cand_val[blockIdx.x] = best, a.cand_idx[blockIdx.x] = besti;
		}
	}
	if (timer) a.perf[6] += globaltimer_ns() - t_stage;
}
