// stages.cuh -- one kernel per stage of calm's per-token forward() ("staged" engine).
//
// Stage list per layer (reference CPU order, infer.c:349-458; CUDA stages infer.cu:441-620):
//   k_qkv       RMSNorm -> q,k,v matvec (+bias, clip) -> RoPE -> q vector / KV-cache append
//   k_attn      flash-decoding over the KV cache: scores, online softmax, value mix, split over positions
//   k_matres    x += W . v              (attention output projection wo; FFN down projection w2, MoE-weighted)
//   k_ffn_up    RMSNorm -> [router] -> silu/gelu(w1 . xn) * (w3 . xn)
// and once per token: k_embed (embedding row + sink re-rotation) and k_output (final norm + classifier
// [+ greedy argmax]).  Every weight byte is read exactly once per token with 16-byte loads.
#pragma once

#include "common.cuh"

#define CALM_MAX_ACTIVE 8 // experts evaluated per token (Mixtral: 2)

struct MoeSel { // router decision of the current layer, written by k_ffn_up, read by k_matres
	int expert[CALM_MAX_ACTIVE];
	float weight[CALM_MAX_ACTIVE];
};

// ------------------------------------------------------------------------------------------------
// Warp-level matvec core: R rows (any pointers) against the staged activation vector.
// Each lane owns every 32nd 16-byte vector of a row; U vectors per row are requested before the
// first is consumed, so a warp keeps R*U*512 bytes in flight.

template <int DBITS, int R>
__device__ __forceinline__ void warp_dot_rows(const uint4* const (&rp)[R], int nvec, const float4* __restrict__ xs4, float (&out)[R]) {
	constexpr int Q = WFmt<DBITS>::VW / 4;
	constexpr int U = DBITS == 4 ? 2 : 4;
	const int lane = threadIdx.x & 31;

	float acc[R];
#pragma unroll
	for (int r = 0; r < R; ++r) acc[r] = 0.f;

	for (int v0 = lane; v0 < nvec; v0 += 32 * U) {
		uint4 w[U][R];
#pragma unroll
		for (int u = 0; u < U; ++u) {
			int v = v0 + 32 * u;
#pragma unroll
			for (int r = 0; r < R; ++r) w[u][r] = v < nvec ? ldg_stream(rp[r] + v) : make_uint4(0, 0, 0, 0);
		}
#pragma unroll
		for (int u = 0; u < U; ++u) {
			int v = v0 + 32 * u;
			if (v < nvec) {
				float4 xv[Q];
				const float4* xp = xs4 + (size_t)(v >> 5) * Q * 32 + lane;
#pragma unroll
				for (int q = 0; q < Q; ++q) xv[q] = xp[q * 32];
#pragma unroll
				for (int r = 0; r < R; ++r) acc[r] = dot_vec<DBITS>(w[u][r], xv, acc[r]);
			}
		}
	}
#pragma unroll
	for (int r = 0; r < R; ++r) out[r] = warp_sum(acc[r]);
}

// Stage an activation vector into shared memory in the permuted layout of common.cuh, optionally
// applying RMSNorm / LayerNorm-without-bias first (reference infer.c:183-207: mean only when ln,
// variance around the mean, eps inside the sqrt, then * weight).  Ends with a __syncthreads().
template <int DBITS>
__device__ __forceinline__ void stage_vector(float* xs, float* red, const float* __restrict__ x, int n, const float* __restrict__ normw, float eps, bool ln,
                                             float* xb_out) {
	const int tid = threadIdx.x, nthr = blockDim.x;
	float mean = 0.f, scale = 1.f;
	if (normw) {
		if (ln) {
			float s = 0.f;
			for (int j = tid; j < n; j += nthr) s += x[j];
			mean = block_sum(s, red) / n;
		}
		float ss = 0.f;
		for (int j = tid; j < n; j += nthr) {
			float d = x[j] - mean;
			ss = fmaf(d, d, ss);
		}
		ss = block_sum(ss, red);
		scale = 1.0f / sqrtf(ss / n + eps);
	}
	const int total = xs_floats<DBITS>(n);
	for (int j = tid; j < total; j += nthr) {
		float v = 0.f;
		if (j < n) {
			v = x[j];
			if (normw) v = (v - mean) * scale * normw[j];
			if (xb_out) xb_out[j] = v;
		}
		xs[xs_index<DBITS>(j)] = v;
	}
	__syncthreads();
}

__device__ __forceinline__ float act_silu(float x) { // reference infer.c:273-275
	return x / (1.0f + expf(-x));
}
__device__ __forceinline__ float act_gelu(float x) { // reference infer.c:269-271
	return 0.5f * x * (1.0f + tanhf(0.797885f * (x + 0.044715f * x * x * x)));
}

// ------------------------------------------------------------------------------------------------
// k_embed: x = decode(E[token]) (reference infer.c:335-347, infer.cu:142-148), and -- in the extra
// blocks -- the attention-sink re-rotation by one RoPE step for every layer once the cache has rolled
// over (reference infer.c:384-394, infer.cu:150-180).

template <typename KVT>
struct EmbedArgs {
	float* x;
	const void* table;
	const TokenParams* tp;
	int dim;
	int embed_blocks;
	// sinks
	KVT* key_cache;
	const float* rope_freq;
	int n_layers, n_kv_heads, head_dim, seq_len;
};

template <int DBITS, typename KVT>
__global__ void k_embed(const EmbedArgs<KVT> a) {
	if ((int)blockIdx.x < a.embed_blocks) {
		int i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i < a.dim) a.x[i] = weight_at<DBITS>(a.table, (size_t)a.tp->token * a.dim + i);
		return;
	}
	const int kv_sink = a.tp->kv_sink;
	if (kv_sink == 0) return;
	const int kv_dim = a.n_kv_heads * a.head_dim;
	const int pairs = a.n_layers * kv_sink * (kv_dim / 2);
	for (int i = (blockIdx.x - a.embed_blocks) * blockDim.x + threadIdx.x; i < pairs; i += (gridDim.x - a.embed_blocks) * blockDim.x) {
		int k = (i % (kv_dim / 2)) * 2;
		int r = (i / (kv_dim / 2)) % kv_sink;
		int l = i / (kv_dim / 2) / kv_sink;
		int h = k / a.head_dim, d = k % a.head_dim;
		KVT* p = a.key_cache + (((size_t)l * a.n_kv_heads + h) * a.seq_len + r) * a.head_dim + d;
		float fcr, fci;
		sincosf(a.rope_freq[d >> 1], &fci, &fcr); // one position step
		float v0 = kv_load(p), v1 = kv_load(p + 1);
		kv_store(p, v0 * fcr - v1 * fci);
		kv_store(p + 1, v0 * fci + v1 * fcr);
	}
}

// ------------------------------------------------------------------------------------------------
// k_qkv

template <typename KVT>
struct QkvArgs {
	const float* x;
	const float* normw;
	const void* wq;
	const void* wk;
	const void* wv;
	const float* bias;
	float* q_out;
	KVT* kc; // this layer: [n_kv_heads][seq_len][head_dim]
	KVT* vc;
	const float* rope_freq; // [head_dim/2], theta^(-j/rotary_dim) or 0 beyond rotary_dim
	float* xb_out;          // normalised x for the FFN when norm_par, else NULL
	const TokenParams* tp;
	int dim, q_dim, kv_dim, head_dim, seq_len;
	float eps, clip;
	int ln;
};

template <int DBITS, typename KVT>
__global__ void __launch_bounds__(256) k_qkv(const QkvArgs<KVT> a) {
	extern __shared__ __align__(16) float smem[];
	float* red = smem;
	float* xs = smem + 32;
	stage_vector<DBITS>(xs, red, a.x, a.dim, a.normw, a.eps, a.ln != 0, blockIdx.x == 0 ? a.xb_out : nullptr);

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	const int nvec = a.dim / WFmt<DBITS>::VW;
	const size_t rowvecs = (size_t)nvec; // 16-byte vectors per row
	const int npairs = (a.q_dim + 2 * a.kv_dim) / 2;
	const int pos = a.tp->pos, kv_pos = a.tp->kv_pos;

	for (int p = blockIdx.x * nwarps + warp; p < npairs; p += gridDim.x * nwarps) {
		int j = 2 * p; // row in the concatenated [wq; wk; wv]
		const void* w;
		int k;
		if (j < a.q_dim) {
			w = a.wq, k = j;
		} else if (j < a.q_dim + a.kv_dim) {
			w = a.wk, k = j - a.q_dim;
		} else {
			w = a.wv, k = j - a.q_dim - a.kv_dim;
		}
		const uint4* rp[2] = {reinterpret_cast<const uint4*>(w) + (size_t)k * rowvecs, reinterpret_cast<const uint4*>(w) + (size_t)(k + 1) * rowvecs};
		float v[2];
		warp_dot_rows<DBITS, 2>(rp, nvec, reinterpret_cast<const float4*>(xs), v);

		if (lane == 0) {
			float v0 = v[0], v1 = v[1];
			if (a.bias) v0 += a.bias[j], v1 += a.bias[j + 1];
			v0 = fminf(fmaxf(v0, -a.clip), a.clip);
			v1 = fminf(fmaxf(v1, -a.clip), a.clip);
			if (j < a.q_dim + a.kv_dim) { // rotate q and k (reference infer.c:223-236)
				float fcr, fci;
				sincosf((float)pos * a.rope_freq[(j % a.head_dim) >> 1], &fci, &fcr);
				float r0 = v0 * fcr - v1 * fci, r1 = v0 * fci + v1 * fcr;
				v0 = r0, v1 = r1;
			}
			if (j < a.q_dim) {
				a.q_out[k] = v0, a.q_out[k + 1] = v1;
			} else {
				KVT* c = (j < a.q_dim + a.kv_dim) ? a.kc : a.vc;
				int h = k / a.head_dim, d = k % a.head_dim;
				KVT* dst = c + ((size_t)h * a.seq_len + kv_pos) * a.head_dim + d;
				kv_store(dst, v0);
				kv_store(dst + 1, v1);
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// k_attn: one CTA per (kv head, slice of positions).  All KVMUL query heads that share the kv head
// are processed together so K and V are read once.  A group of LPP lanes owns one position at a time
// (8 head dims per lane, one 16-byte load of K and of V); P positions are in flight per group.
// Scores are scaled by 1/sqrt(head_dim) after the dot (reference infer.c:246), softmax is the
// max-shifted one of infer.c:252-258 evaluated online, and the 1/sum is applied once at the end.

struct AttnArgs {
	const float* q;   // [n_heads*head_dim] rotated queries
	const void* kc;   // this layer
	const void* vc;
	float* partial;   // [units][nsplit][KVMUL][head_dim + 2], unit = (kv head, group of KVMUL query heads)
	unsigned* counter; // [units], zero between launches
	float* out;       // [n_heads*head_dim]
	const TokenParams* tp;
	int head_dim, seq_len, nsplit, lpp; // lpp: lanes per position (power of two >= head_dim/8)
	int kv_mul, qgroups;                // query heads per kv head; qgroups = kv_mul / KVMUL
	float inv_sqrt_hd;
};

template <typename KVT, int KVMUL>
__global__ void __launch_bounds__(128) k_attn(const AttnArgs a) {
	constexpr int P = KVMUL > 4 ? 2 : 4;
	extern __shared__ __align__(16) float smem[];
	__shared__ int is_last;

	const int unit = blockIdx.x / a.nsplit, split = blockIdx.x % a.nsplit;
	const int kvh = unit / a.qgroups;
	const int hbase = kvh * a.kv_mul + (unit % a.qgroups) * KVMUL; // first query head of this unit
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	const int hd = a.head_dim, lpp = a.lpp;
	const int G = 32 / lpp;          // position groups per warp
	const int grp = lane / lpp, li = lane % lpp;
	const bool dact = li * 8 < hd;   // lane holds real head dims

	const int kv_len = a.tp->kv_len;
	const int chunk = (kv_len + a.nsplit - 1) / a.nsplit;
	const int t0 = split * chunk, t1 = min(kv_len, t0 + chunk);

	const KVT* kbase = reinterpret_cast<const KVT*>(a.kc) + (size_t)kvh * a.seq_len * hd + li * 8;
	const KVT* vbase = reinterpret_cast<const KVT*>(a.vc) + (size_t)kvh * a.seq_len * hd + li * 8;

	float qr[KVMUL][8], acc[KVMUL][8], m[KVMUL], l[KVMUL];
#pragma unroll
	for (int h = 0; h < KVMUL; ++h) {
		m[h] = -FLT_MAX, l[h] = 0.f;
#pragma unroll
		for (int d = 0; d < 8; ++d) {
			acc[h][d] = 0.f;
			qr[h][d] = dact ? a.q[(size_t)(hbase + h) * hd + li * 8 + d] : 0.f;
		}
	}

	const int stride = nwarps * G;
	// the trip count must be warp-uniform: the score reduction below shuffles across the full warp
	for (int tw = t0 + warp * G; tw < t1; tw += stride * P) {
		const int tb = tw + grp;
		float kf[P][8], vf[P][8];
		bool ok[P];
#pragma unroll
		for (int i = 0; i < P; ++i) {
			int t = tb + i * stride;
			ok[i] = t < t1;
			if (ok[i] && dact) {
				kv_load8(kbase + (size_t)t * hd, kf[i]);
				kv_load8(vbase + (size_t)t * hd, vf[i]);
			} else {
#pragma unroll
				for (int d = 0; d < 8; ++d) kf[i][d] = 0.f, vf[i][d] = 0.f;
			}
		}
#pragma unroll
		for (int h = 0; h < KVMUL; ++h) {
			float s[P], smax = m[h];
#pragma unroll
			for (int i = 0; i < P; ++i) {
				float d = 0.f;
#pragma unroll
				for (int e = 0; e < 8; ++e) d = fmaf(qr[h][e], kf[i][e], d);
				for (int o = 1; o < lpp; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
				s[i] = ok[i] ? d * a.inv_sqrt_hd : -FLT_MAX;
				smax = fmaxf(smax, s[i]);
			}
			float corr = expf(m[h] - smax);
			m[h] = smax;
			float pw[P], ps = 0.f;
#pragma unroll
			for (int i = 0; i < P; ++i) {
				pw[i] = ok[i] ? expf(s[i] - smax) : 0.f;
				ps += pw[i];
			}
			l[h] = fmaf(l[h], corr, ps);
#pragma unroll
			for (int e = 0; e < 8; ++e) {
				float v = acc[h][e] * corr;
#pragma unroll
				for (int i = 0; i < P; ++i) v = fmaf(pw[i], vf[i][e], v);
				acc[h][e] = v;
			}
		}
	}

	// merge the position groups of a warp
	for (int o = lpp; o < 32; o <<= 1) {
#pragma unroll
		for (int h = 0; h < KVMUL; ++h) {
			float mo = __shfl_xor_sync(0xffffffffu, m[h], o), lo = __shfl_xor_sync(0xffffffffu, l[h], o);
			float mn = fmaxf(m[h], mo);
			float ca = expf(m[h] - mn), cb = expf(mo - mn);
			l[h] = l[h] * ca + lo * cb;
#pragma unroll
			for (int e = 0; e < 8; ++e) {
				float ao = __shfl_xor_sync(0xffffffffu, acc[h][e], o);
				acc[h][e] = acc[h][e] * ca + ao * cb;
			}
			m[h] = mn;
		}
	}

	// merge warps through shared memory: rec[warp][h][hd+2]
	const int rec = hd + 2;
	if (grp == 0) {
#pragma unroll
		for (int h = 0; h < KVMUL; ++h) {
			float* r = smem + ((size_t)warp * KVMUL + h) * rec;
			if (dact) {
#pragma unroll
				for (int e = 0; e < 8; ++e) r[li * 8 + e] = acc[h][e];
			}
			if (li == 0) r[hd] = m[h], r[hd + 1] = l[h];
		}
	}
	__syncthreads();
	float* part = a.partial + ((size_t)unit * a.nsplit + split) * KVMUL * rec;
	for (int idx = threadIdx.x; idx < KVMUL * rec; idx += blockDim.x) {
		int h = idx / rec, e = idx % rec;
		float mn = -FLT_MAX;
		for (int w = 0; w < nwarps; ++w) mn = fmaxf(mn, smem[((size_t)w * KVMUL + h) * rec + hd]);
		float v;
		if (e == hd) {
			v = mn;
		} else {
			v = 0.f;
			for (int w = 0; w < nwarps; ++w) {
				const float* r = smem + ((size_t)w * KVMUL + h) * rec;
				v += r[e] * expf(r[hd] - mn); // e == hd+1 merges the sums the same way
			}
		}
		part[idx] = v;
	}

	// the last slice of this kv head to finish folds all slices and writes the normalised output
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned old = atomicAdd(a.counter + unit, 1u);
		is_last = (old == (unsigned)a.nsplit - 1);
	}
	__syncthreads();
	if (!is_last) return;
	__threadfence();
	const float* pk = a.partial + (size_t)unit * a.nsplit * KVMUL * rec;
	for (int idx = threadIdx.x; idx < KVMUL * hd; idx += blockDim.x) {
		int h = idx / hd, e = idx % hd;
		float mn = -FLT_MAX;
		for (int s = 0; s < a.nsplit; ++s) mn = fmaxf(mn, __ldcg(pk + ((size_t)s * KVMUL + h) * rec + hd));
		float num = 0.f, den = 0.f;
		for (int s = 0; s < a.nsplit; ++s) {
			const float* r = pk + ((size_t)s * KVMUL + h) * rec;
			float c = expf(__ldcg(r + hd) - mn);
			num = fmaf(__ldcg(r + e), c, num);
			den = fmaf(__ldcg(r + hd + 1), c, den);
		}
		a.out[(size_t)(hbase + h) * hd + e] = num / den;
	}
	if (threadIdx.x == 0) a.counter[unit] = 0;
}

// ------------------------------------------------------------------------------------------------
// k_matres: y[row] (+)= sum_e weight_e * (W_e[row] . xin_e)   -- wo (+ residual, infer.c:410-415) and
// w2 (* router weight, + residual, infer.c:452-456).  Experts are visited in selection order by the
// same lane, so the sum is deterministic (the reference CUDA path uses atomicAdd, infer.cu:618).

struct MatResArgs {
	const float* xin; // [nact][n]
	const void* w;    // [n_experts?][d][n]
	float* y;         // [d]
	const MoeSel* sel; // NULL: one pass with expert 0, weight 1
	int n, d, nact;
	int accumulate;   // 1: y += ..., 0: y = ...
};

template <int DBITS>
__global__ void __launch_bounds__(256) k_matres(const MatResArgs a) {
	extern __shared__ __align__(16) float smem[];
	float* red = smem;
	float* xs = smem + 32;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	const int nvec = a.n / WFmt<DBITS>::VW;
	const size_t esize = (size_t)a.d * nvec; // 16-byte vectors per expert

	for (int e = 0; e < a.nact; ++e) {
		if (e > 0) __syncthreads();
		stage_vector<DBITS>(xs, red, a.xin + (size_t)e * a.n, a.n, nullptr, 0.f, false, nullptr);
		const int ex = a.sel ? a.sel->expert[e] : 0;
		const float ew = a.sel ? a.sel->weight[e] : 1.f;
		const uint4* wb = reinterpret_cast<const uint4*>(a.w) + (size_t)ex * esize;
		for (int p = blockIdx.x * nwarps + warp; p < a.d / 2; p += gridDim.x * nwarps) {
			const uint4* rp[2] = {wb + (size_t)(2 * p) * nvec, wb + (size_t)(2 * p + 1) * nvec};
			float v[2];
			warp_dot_rows<DBITS, 2>(rp, nvec, reinterpret_cast<const float4*>(xs), v);
			if (lane == 0) {
				float2* dst = reinterpret_cast<float2*>(a.y + 2 * p);
				float2 cur = (a.accumulate || e > 0) ? *dst : make_float2(0.f, 0.f);
				cur.x += v[0] * ew;
				cur.y += v[1] * ew;
				*dst = cur;
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// k_ffn_up: hb[e][i] = act(w1_e[i] . xn) * (w3_e[i] . xn) (reference infer.c:437-450), with the
// router evaluated first by every CTA for MoE models (logits = gate . xn, top-k by repeated arg-max
// with strict '>' so the lowest index wins ties, weights = softmax over the selected; infer.c:277-305).

struct FfnUpArgs {
	const float* x;
	const float* normw; // NULL: stage `x` as is (norm_par: x = saved attention-norm output)
	const void* gate;   // router (n_experts, dim) or NULL
	const void* w1;
	const void* w3;
	float* hb;          // [nact][hidden]
	MoeSel* sel;        // written by CTA 0
	int dim, hidden, n_experts, nact;
	float eps;
	int ln, gelu;
};

template <int DBITS>
__global__ void __launch_bounds__(256) k_ffn_up(const FfnUpArgs a) {
	extern __shared__ __align__(16) float smem[];
	__shared__ float glog[64];
	__shared__ MoeSel ssel;
	float* red = smem;
	float* xs = smem + 32;
	stage_vector<DBITS>(xs, red, a.x, a.dim, a.normw, a.eps, a.ln != 0, nullptr);

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	const int nvec = a.dim / WFmt<DBITS>::VW;
	const float4* xs4 = reinterpret_cast<const float4*>(xs);

	if (a.n_experts) {
		for (int e = warp; e < a.n_experts; e += nwarps) {
			const uint4* rp[1] = {reinterpret_cast<const uint4*>(a.gate) + (size_t)e * nvec};
			float v[1];
			warp_dot_rows<DBITS, 1>(rp, nvec, xs4, v);
			if (lane == 0) glog[e] = v[0];
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			float mx = -FLT_MAX;
			for (int j = 0; j < a.n_experts; ++j) mx = fmaxf(mx, glog[j]);
			unsigned long long mask = 0;
			float wsum = 0.f;
			for (int k = 0; k < a.nact; ++k) {
				int best = -1;
				for (int j = 0; j < a.n_experts; ++j)
					if (!((mask >> j) & 1ull) && (best < 0 || glog[j] > glog[best])) best = j;
				ssel.expert[k] = best;
				ssel.weight[k] = expf(glog[best] - mx);
				wsum += ssel.weight[k];
				mask |= 1ull << best;
			}
			for (int k = 0; k < a.nact; ++k) ssel.weight[k] /= wsum;
			if (blockIdx.x == 0) *a.sel = ssel;
		}
		__syncthreads();
	}

	const size_t esize = (size_t)a.hidden * nvec;
	const int total = a.nact * a.hidden;
	for (int p = blockIdx.x * nwarps + warp; p < total; p += gridDim.x * nwarps) {
		int e = p / a.hidden, i = p % a.hidden;
		size_t off = (a.n_experts ? (size_t)ssel.expert[e] * esize : 0) + (size_t)i * nvec;
		const uint4* rp[2] = {reinterpret_cast<const uint4*>(a.w1) + off, reinterpret_cast<const uint4*>(a.w3) + off};
		float v[2];
		warp_dot_rows<DBITS, 2>(rp, nvec, xs4, v);
		if (lane == 0) a.hb[p] = (a.gelu ? act_gelu(v[0]) : act_silu(v[0])) * v[1];
	}
}

// ------------------------------------------------------------------------------------------------
// k_output: logits = Wcls . norm(x) (reference infer.c:466-469, infer.cu:628-649) written in 128-byte
// pieces (the destination may be pinned host memory), plus per-CTA greedy candidates
// (first maximum wins, reference sampler.c:34-42).

struct OutputArgs {
	const float* x;
	const float* normw;
	const void* wcls;
	float* logits;
	float* cand_val; // [gridDim.x]
	int* cand_idx;
	int dim, vocab;
	float eps;
	int ln;
};

template <int DBITS>
__global__ void __launch_bounds__(256) k_output(const OutputArgs a) {
	constexpr int R = 4;
	extern __shared__ __align__(16) float smem[];
	__shared__ float outbuf[32]; // nwarps * R
	__shared__ float bval[8];
	__shared__ int bidx[8];
	float* red = smem;
	float* xs = smem + 32;
	stage_vector<DBITS>(xs, red, a.x, a.dim, a.normw, a.eps, a.ln != 0, nullptr);

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5; // nwarps == 8
	const int nvec = a.dim / WFmt<DBITS>::VW;
	const int rows_per_iter = nwarps * R;
	float best = -FLT_MAX;
	int besti = 0x7fffffff;

	for (int base = blockIdx.x * rows_per_iter; base < a.vocab; base += gridDim.x * rows_per_iter) {
		int r0 = base + warp * R;
		const uint4* rp[R];
#pragma unroll
		for (int r = 0; r < R; ++r) rp[r] = reinterpret_cast<const uint4*>(a.wcls) + (size_t)min(r0 + r, a.vocab - 1) * nvec;
		float v[R];
		warp_dot_rows<DBITS, R>(rp, nvec, reinterpret_cast<const float4*>(xs), v);
		if (lane == 0) {
#pragma unroll
			for (int r = 0; r < R; ++r) {
				outbuf[warp * R + r] = v[r];
				if (r0 + r < a.vocab && v[r] > best) best = v[r], besti = r0 + r;
			}
		}
		__syncthreads();
		if (threadIdx.x < rows_per_iter && base + (int)threadIdx.x < a.vocab) a.logits[base + threadIdx.x] = outbuf[threadIdx.x];
		__syncthreads();
	}

	if (a.cand_val) {
		if (lane == 0) bval[warp] = best, bidx[warp] = besti;
		__syncthreads();
		if (threadIdx.x == 0) {
			for (int w = 1; w < nwarps; ++w)
				if (bval[w] > best || (bval[w] == best && bidx[w] < besti)) best = bval[w], besti = bidx[w];
			a.cand_val[blockIdx.x] = best;
			a.cand_idx[blockIdx.x] = besti;
		}
	}
}

// Fold the per-CTA candidates, publish the greedy token, and advance the token parameters so the
// next replay of the graph consumes it (device-resident decode loop).
__global__ void k_advance(const float* cand_val, const int* cand_idx, int ncand, TokenParams* tp, int* out_tokens, int* last_token, int advance, int vocab) {
	__shared__ float sv[256];
	__shared__ int si[256];
	float best = -FLT_MAX;
	int besti = 0x7fffffff;
	for (int i = threadIdx.x; i < ncand; i += blockDim.x) {
		float v = cand_val[i];
		int ix = cand_idx[i];
		if (v > best || (v == best && ix < besti)) best = v, besti = ix;
	}
	sv[threadIdx.x] = best, si[threadIdx.x] = besti;
	__syncthreads();
	for (int s = blockDim.x / 2; s > 0; s >>= 1) {
		if ((int)threadIdx.x < s) {
			float v = sv[threadIdx.x + s];
			int ix = si[threadIdx.x + s];
			if (v > sv[threadIdx.x] || (v == sv[threadIdx.x] && ix < si[threadIdx.x])) sv[threadIdx.x] = v, si[threadIdx.x] = ix;
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		int tok = si[0];
		if (tok < 0 || tok >= vocab) tok = 0; // only reachable when every logit is NaN
		*last_token = tok;
		if (advance) {
			out_tokens[tp->step] = tok;
			int pos = tp->pos + 1, seq_len = tp->seq_len;
			int sink = pos >= seq_len ? 2 : 0; // KV_SINKS
			tp->token = tok;
			tp->pos = pos;
			tp->kv_sink = sink;
			tp->kv_pos = sink + (pos - sink) % (seq_len - sink);
			tp->kv_len = pos >= seq_len ? seq_len : pos + 1;
			tp->step = tp->step + 1;
		}
	}
}

__global__ void k_set_params(TokenParams* tp, int token, int pos, int seq_len, int step) {
	int sink = pos >= seq_len ? 2 : 0; // reference infer.c:330-332
	tp->token = token;
	tp->pos = pos;
	tp->kv_sink = sink;
	tp->kv_pos = sink + (pos - sink) % (seq_len - sink);
	tp->kv_len = pos >= seq_len ? seq_len : pos + 1;
	tp->step = step;
	tp->seq_len = seq_len;
}

// plain matvec (unit tests, per-kernel roofline bench): y = W . x
struct MatvecArgs {
	const float* x;
	const void* w;
	float* y;
	int n, d;
};

template <int DBITS>
__global__ void __launch_bounds__(256) k_matvec(const MatvecArgs a) {
	extern __shared__ __align__(16) float smem[];
	float* red = smem;
	float* xs = smem + 32;
	stage_vector<DBITS>(xs, red, a.x, a.n, nullptr, 0.f, false, nullptr);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	const int nvec = a.n / WFmt<DBITS>::VW;
	for (int p = blockIdx.x * nwarps + warp; p < (a.d + 1) / 2; p += gridDim.x * nwarps) {
		int r1 = min(2 * p + 1, a.d - 1);
		const uint4* rp[2] = {reinterpret_cast<const uint4*>(a.w) + (size_t)(2 * p) * nvec, reinterpret_cast<const uint4*>(a.w) + (size_t)r1 * nvec};
		float v[2];
		warp_dot_rows<DBITS, 2>(rp, nvec, reinterpret_cast<const float4*>(xs), v);
		if (lane == 0) {
			a.y[2 * p] = v[0];
			if (2 * p + 1 < a.d) a.y[2 * p + 1] = v[1];
		}
	}
}

// deterministic pseudo-random cache fill (bench: decode at a late position without the prefix)
template <typename KVT>
__global__ void k_fill_kv(KVT* kc, KVT* vc, size_t n_layers_heads, int seq_len, int head_dim, int n_pos, unsigned long long seed) {
	size_t total = n_layers_heads * (size_t)n_pos * head_dim;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		size_t lh = i / ((size_t)n_pos * head_dim), rem = i % ((size_t)n_pos * head_dim);
		size_t idx = lh * seq_len * head_dim + rem;
		unsigned long long z = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
		z ^= z >> 31;
		float a = ((z & 0xffffff) / 16777216.f - 0.5f) * 2.f;         // U(-1,1)
		float b = (((z >> 24) & 0xffffff) / 16777216.f - 0.5f) * 2.f;
		kv_store(kc + idx, a);
		kv_store(vc + idx, b);
	}
}
