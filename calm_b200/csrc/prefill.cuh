// prefill.cuh -- batched prompt pass (reference behaviour replaced: run.c:206-209 feeds the prompt one
// token at a time through forward(), README.md:80).  T prompt tokens go through every layer together, so each weight
// matrix is read once per 128-token tile instead of once per token and the projections become real GEMMs on the
// 5th-generation tensor cores:
//
//   k_pf_gemm   D[128 weight rows][128 tokens] (+)= W_tile . X_tile^T with tcgen05.mma (kind::f16, cta_group::1), fp32
//               accumulators in TMEM.  A = the weight tile, dequantised ON THE FLY by four warps from the model's own
//               format (fp16 / e5m2 / gf4: all exact in f16) into the 128-byte-swizzled K-major shared-memory layout
//               the MMA reads; B = the activations as TWO f16 matrices hi + lo (x = hi + lo to 22 bits, so the result
//               carries fp32-grade inputs: two MMAs per k-step into the same accumulator), brought in by 2-D TMA
//               (cp.async.bulk.tensor, SWIZZLE_128B).  One warp issues TMA, one thread issues MMAs, completion flows
//               through mbarriers (tcgen05.commit), the epilogue reads TMEM with tcgen05.ld and applies what the
//               decode kernels apply: bias / clip / RoPE / KV-cache append (QKV), residual add (wo, w2),
//               act(w1 x) * (w3 x) with both accumulators side by side in TMEM (FFN up).
//   k_pf_attn   causal attention of the block against the cache (fp32 math on the cached fp16 / e5m2 entries, online
//               softmax; reference infer.c:238-267), one warp per query token, K/V tiles staged in shared memory.
//   k_pf_norm   RMSNorm / LayerNorm per token (reference infer.c:183-207) -> hi / lo f16 rows.
//
// Parity contract (tests/test_prefill_gpu.py): after forward_prefill_cuda(tokens, n, pos0) the KV cache and the logits of
// the following token equal those of n serial forward(FF_UPDATE_KV_ONLY) calls within the stated tolerances.
#pragma once

#include <cuda.h>

#include "stages.cuh"

#define PF_BM 128 // weight rows per tile  (TMEM lanes)
#define PF_BN 128 // tokens per tile       (TMEM columns per accumulator)
#define PF_BK 64  // k per stage: 64 f16 = 128 bytes = one swizzle atom
#define PF_THREADS 192
enum { PF_QKV = 0, PF_WO = 1, PF_UP = 2, PF_DOWN = 3 };

// ---------------------------------------------------------------- tcgen05 / TMA wrappers
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) { // one whole warp
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) { // the allocating warp
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, f16 x f16 -> f32, issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "setp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
	    ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
	    : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` when they have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 columns of 32-bit accumulators: thread t of the warp gets row (lane base + t), columns c .. c+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
	uint32_t r[32];
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
	      "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
	      "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
	    : "r"(taddr)
	    : "memory");
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
	for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)), "l"(map),
	             "r"(c0), "r"(c1), "r"(smem_u32(bar))
	             : "memory");
}
// bounded wait (about 2 s): a broken pipeline traps with a message instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, uint32_t parity) {
	unsigned spins = 0;
	while (!mbar_try_wait(bar, parity)) {
		if (++spins > (1u << 26)) {
			printf("calm_b200: k_pf_gemm pipeline stalled (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
			__trap();
		}
	}
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { // generic-proxy writes to shared memory -> visible to the async proxy (tensor core reads)
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// shared-memory matrix descriptor: K-major tile, rows of 128 bytes (64 f16), SWIZZLE_128B, 8-row groups 1024 bytes apart, sm_100 version bit
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* tile) {
	const uint64_t addr = (uint64_t)(smem_u32(tile) & 0x3FFFF) >> 4;
	return addr | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::f16: D = f32, A = B = f16, both K-major, M = 128, N = 128
#define PF_IDESC ((1u << 4) | ((uint32_t)(PF_BN >> 3) << 17) | ((uint32_t)(PF_BM >> 4) << 24))

// ---------------------------------------------------------------- small kernels

// f32 -> (hi, lo) f16 pair with x = hi + lo up to 2^-22 |x|
__device__ __forceinline__ void split_hl(float x, __half& hi, __half& lo) {
	hi = __float2half_rn(x);
	lo = __float2half_rn(x - __half2float(hi));
}

// X[t][:] = decode(E[tokens[t]])  (reference infer.c:335-347)
template <int DBITS>
__global__ void k_pf_embed(float* X, const void* table, const int* tokens, int n, int dim) {
	pdl_enter();
	const int t = blockIdx.x;
	if (t >= n) return;
	const size_t base = (size_t)tokens[t] * dim;
	for (int i = threadIdx.x; i < dim; i += blockDim.x) X[(size_t)t * dim + i] = weight_at<DBITS>(table, base + i);
}

// rope[t][j] = (cos, sin)((pos0 + t) * freq[j])
__global__ void k_pf_rope(float2* rope, const float* freq, int n, int half_hd, int pos0) {
	pdl_enter();
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n * half_hd) return;
	float s, c;
	sincosf((float)(pos0 + i / half_hd) * freq[i % half_hd], &s, &c);
	rope[i] = make_float2(c, s);
}

// one CTA per token: norm (infer.c:183-207) then hi / lo split; rows beyond n (tile padding) are zeroed once at allocation
__global__ void __launch_bounds__(256) k_pf_norm(const float* X, const float* normw, __half* hi, __half* lo, int n, int dim, float eps, int ln) {
	pdl_enter();
	__shared__ float red[32];
	const int t = blockIdx.x;
	if (t >= n) return;
	const float* x = X + (size_t)t * dim;
	float mean = 0.f;
	if (ln) {
		float s = 0.f;
		for (int i = threadIdx.x; i < dim; i += blockDim.x) s += x[i];
		mean = block_sum(s, red) / dim;
	}
	float ss = 0.f;
	for (int i = threadIdx.x; i < dim; i += blockDim.x) {
		const float d = x[i] - mean;
		ss = fmaf(d, d, ss);
	}
	ss = block_sum(ss, red);
	const float scale = 1.0f / sqrtf(ss / dim + eps);
	for (int i = threadIdx.x; i < dim; i += blockDim.x) {
		const float v = (x[i] - mean) * scale * normw[i];
		__half h, l;
		split_hl(v, h, l);
		hi[(size_t)t * dim + i] = h, lo[(size_t)t * dim + i] = l;
	}
}

// ---------------------------------------------------------------- the GEMM

struct PfGemmArgs {
	const void* w[3]; // QKV: wq, wk, wv   UP: w1, w3   WO / DOWN: w
	int K, n_tokens;
	// QKV epilogue
	float* Q;         // [T][q_dim]
	int q_dim, kv_dim, head_dim, seq_len, pos0;
	const float* bias;
	float clip;
	const float2* rope; // [T][head_dim / 2]
	void* kc;           // this layer: [n_kv_heads][seq_len][head_dim]
	void* vc;
	// WO / DOWN epilogue
	float* X; // [T][dim] residual stream
	int dim;
	// UP epilogue
	__half* Hhi;
	__half* Hlo;
	int hidden, gelu;
};

// 64 weights of one row (one k-stage) -> 8 swizzled 16-byte chunks of f16 in the A tile
template <int DBITS>
__device__ __forceinline__ void pf_load_raw(const unsigned char* row, int kb, uint4 (&raw)[DBITS / 2]) {
	const uint4* p = reinterpret_cast<const uint4*>(row + (size_t)kb * (PF_BK * DBITS / 8));
#pragma unroll
	for (int i = 0; i < DBITS / 2; ++i) raw[i] = ldg_stream(p + i);
}
template <int DBITS>
__device__ __forceinline__ void pf_store_a(unsigned char* tile, int r, const uint4 (&raw)[DBITS / 2]) {
	unsigned char* rowp = tile + r * 128;
	auto put = [&](int c, uint4 v) { *reinterpret_cast<uint4*>(rowp + ((c ^ (r & 7)) << 4)) = v; };
	if constexpr (DBITS == 16) {
#pragma unroll
		for (int c = 0; c < 8; ++c) put(c, raw[c]);
	} else if constexpr (DBITS == 8) { // an e5m2 byte is the high byte of the f16 (reference infer.c:28-35)
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
			put(2 * i, make_uint4(__byte_perm(w[0], 0, 0x1404), __byte_perm(w[0], 0, 0x3424), __byte_perm(w[1], 0, 0x1404), __byte_perm(w[1], 0, 0x3424)));
			put(2 * i + 1, make_uint4(__byte_perm(w[2], 0, 0x1404), __byte_perm(w[2], 0, 0x3424), __byte_perm(w[3], 0, 0x1404), __byte_perm(w[3], 0, 0x3424)));
		}
	} else { // gf4: (q - 4) * s / -4 (reference infer.c:37-40); at most 6 significant bits: exact in f16
#pragma unroll
		for (int i = 0; i < 2; ++i) {
			const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const float sf = e5m2_to_float((uint8_t)(w[j] & 0xff)) * -0.25f;
				uint32_t h[4];
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					const float a0 = (float)((int)((w[j] >> (8 + 6 * k)) & 7) - 4) * sf, a1 = (float)((int)((w[j] >> (11 + 6 * k)) & 7) - 4) * sf;
					h[k] = pack_h2(__float2half_rn(a0), __float2half_rn(a1));
				}
				put(4 * i + j, make_uint4(h[0], h[1], h[2], h[3]));
			}
		}
	}
}

template <int MODE>
__host__ __device__ constexpr int pf_stages() {
	return MODE == PF_UP ? 3 : 4;
}
template <int MODE>
__host__ __device__ constexpr size_t pf_stage_bytes() {
	return (size_t)(MODE == PF_UP ? 2 : 1) * PF_BM * 128 + 2 * PF_BN * 128; // A tile(s) + B hi + B lo
}
template <int MODE>
__host__ __device__ constexpr size_t pf_smem_bytes() {
	return pf_stages<MODE>() * pf_stage_bytes<MODE>() + 1024; // + alignment slack
}

template <int MODE, int DBITS, typename KVT>
__global__ void __launch_bounds__(PF_THREADS, 1) k_pf_gemm(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const PfGemmArgs a) {
	constexpr int STAGES = pf_stages<MODE>(), NA = MODE == PF_UP ? 2 : 1;
	constexpr uint32_t TMEM_COLS = NA * PF_BN;
	extern __shared__ unsigned char smem_dyn[];
	__shared__ __align__(8) uint64_t full[STAGES], empty[STAGES], accum;
	__shared__ uint32_t tmem_base_s;
	unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int row0 = blockIdx.x * PF_BM, tok0 = blockIdx.y * PF_BN;
	const int nkb = a.K / PF_BK;
	auto a_tile = [&](int s, int i) { return smem + (size_t)s * pf_stage_bytes<MODE>() + (size_t)i * PF_BM * 128; };
	auto b_tile = [&](int s, int i) { return smem + (size_t)s * pf_stage_bytes<MODE>() + (size_t)NA * PF_BM * 128 + (size_t)i * PF_BN * 128; };

	pdl_launch_next();
	if (tid == 0) {
		for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1 + 4), mbar_init(&empty[s], 1);
		mbar_init(&accum, 1);
		mbar_init_fence();
	}
	if (warp == 1) tmem_alloc(&tmem_base_s, TMEM_COLS);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = tmem_base_s;
	pdl_wait_prev(); // activations (and, for the KV rows, the cache) come from the previous kernels of the layer

	if (warp == 0) {
		// ===== TMA producer: the two activation tiles of every k-stage
		if (lane == 0) {
			for (int kb = 0; kb < nkb; ++kb) {
				const int s = kb % STAGES;
				mbar_wait_guard(&empty[s], ((kb / STAGES) & 1) ^ 1);
				mbar_expect_tx(&full[s], 2 * PF_BN * 128);
				tma_load_2d(b_tile(s, 0), &tm_hi, kb * PF_BK, tok0, &full[s]);
				tma_load_2d(b_tile(s, 1), &tm_lo, kb * PF_BK, tok0, &full[s]);
			}
		}
	} else if (warp == 1) {
		// ===== MMA issuer
		if (lane == 0) {
			for (int kb = 0; kb < nkb; ++kb) {
				const int s = kb % STAGES;
				mbar_wait_guard(&full[s], (kb / STAGES) & 1);
				tc_fence_after();
#pragma unroll
				for (int k = 0; k < PF_BK / 16; ++k) { // 16 f16 = 32 bytes along K inside the swizzle atom: +2 in the 16-byte address field
					const uint64_t db_hi = umma_desc_sw128(b_tile(s, 0)) + 2 * k, db_lo = umma_desc_sw128(b_tile(s, 1)) + 2 * k;
#pragma unroll
					for (int i = 0; i < NA; ++i) {
						const uint64_t da = umma_desc_sw128(a_tile(s, i)) + 2 * k;
						umma_f16(tmem_base + i * PF_BN, da, db_hi, PF_IDESC, (kb | k) ? 1u : 0u);
						umma_f16(tmem_base + i * PF_BN, da, db_lo, PF_IDESC, 1u);
					}
				}
				umma_commit(&empty[s]); // frees the stage when these MMAs have read it
			}
			umma_commit(&accum);
		}
	} else {
		// ===== dequantising producers (one tile row per thread), then the epilogue (one accumulator row per thread)
		const int r = tid - 64; // 0..127
		const unsigned char* rowp[NA];
		{
			const size_t rowbytes = (size_t)a.K * DBITS / 8;
			const int gr = row0 + r;
			if constexpr (MODE == PF_QKV) {
				const void* w = gr < a.q_dim ? a.w[0] : (gr < a.q_dim + a.kv_dim ? a.w[1] : a.w[2]);
				const int k = gr < a.q_dim ? gr : (gr < a.q_dim + a.kv_dim ? gr - a.q_dim : gr - a.q_dim - a.kv_dim);
				rowp[0] = reinterpret_cast<const unsigned char*>(w) + (size_t)k * rowbytes;
			} else if constexpr (MODE == PF_UP) {
				rowp[0] = reinterpret_cast<const unsigned char*>(a.w[0]) + (size_t)gr * rowbytes;
				rowp[1] = reinterpret_cast<const unsigned char*>(a.w[1]) + (size_t)gr * rowbytes;
			} else {
				rowp[0] = reinterpret_cast<const unsigned char*>(a.w[0]) + (size_t)gr * rowbytes;
			}
		}
		uint4 raw[NA][DBITS / 2];
#pragma unroll
		for (int i = 0; i < NA; ++i) pf_load_raw<DBITS>(rowp[i], 0, raw[i]);
		for (int kb = 0; kb < nkb; ++kb) {
			const int s = kb % STAGES;
			mbar_wait_guard(&empty[s], ((kb / STAGES) & 1) ^ 1);
#pragma unroll
			for (int i = 0; i < NA; ++i) pf_store_a<DBITS>(a_tile(s, i), r, raw[i]);
			if (kb + 1 < nkb) { // the next stage's bytes travel while this one is being multiplied
#pragma unroll
				for (int i = 0; i < NA; ++i) pf_load_raw<DBITS>(rowp[i], kb + 1, raw[i]);
			}
			fence_async_smem();
			__syncwarp();
			if (lane == 0) mbar_arrive(&full[s]);
		}

		mbar_wait_guard(&accum, 0);
		tc_fence_after();
		const int q4 = warp & 3; // the TMEM lane quarter this warp may read
		const int lr = q4 * 32 + lane, gr = row0 + lr;
		for (int c0 = 0; c0 < PF_BN; c0 += 32) {
			float v[32];
			tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + c0, v);
			if constexpr (MODE == PF_QKV) {
				const bool is_q = gr < a.q_dim, is_k = !is_q && gr < a.q_dim + a.kv_dim;
				const int k = is_q ? gr : (is_k ? gr - a.q_dim : gr - a.q_dim - a.kv_dim);
				const int d = k % a.head_dim, h = k / a.head_dim;
				const float b = a.bias ? a.bias[gr] : 0.f;
#pragma unroll
				for (int j = 0; j < 32; ++j) {
					const int t = tok0 + c0 + j;
					float x = fminf(fmaxf(v[j] + b, -a.clip), a.clip);
					const float other = __shfl_xor_sync(0xffffffffu, x, 1); // the row pair a rotation mixes sits in adjacent lanes
					if (t < a.n_tokens) {
						if (is_q || is_k) {
							const float2 cs = a.rope[(size_t)t * (a.head_dim / 2) + (d >> 1)];
							x = (lane & 1) ? fmaf(other, cs.y, x * cs.x) : fmaf(-other, cs.y, x * cs.x);
						}
						if (is_q) {
							a.Q[(size_t)t * a.q_dim + k] = x;
						} else {
							KVT* dst = reinterpret_cast<KVT*>(is_k ? a.kc : a.vc) + ((size_t)h * a.seq_len + a.pos0 + t) * a.head_dim + d;
							kv_store(dst, x);
						}
					}
				}
			} else if constexpr (MODE == PF_UP) {
				float v3[32];
				tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + PF_BN + c0, v3);
#pragma unroll
				for (int j = 0; j < 32; ++j) {
					const int t = tok0 + c0 + j;
					if (t < a.n_tokens) {
						const float hv = (a.gelu ? act_gelu(v[j]) : act_silu(v[j])) * v3[j];
						__half hh, hl;
						split_hl(hv, hh, hl);
						a.Hhi[(size_t)t * a.hidden + gr] = hh, a.Hlo[(size_t)t * a.hidden + gr] = hl;
					}
				}
			} else {
#pragma unroll
				for (int j = 0; j < 32; ++j) {
					const int t = tok0 + c0 + j;
					if (t < a.n_tokens) a.X[(size_t)t * a.dim + gr] += v[j];
				}
			}
		}
		tc_fence_before();
	}
	__syncthreads();
	if (warp == 1) {
		tc_fence_after();
		tmem_dealloc(tmem_base, TMEM_COLS);
	}
}

// ---------------------------------------------------------------- causal attention of the block against the cache
#define PFA_WARPS 8
#define PFA_TILE 32 // keys per shared-memory tile: one per lane

struct PfAttnArgs {
	const float* Q; // [T][q_dim]
	const void* kc; // this layer
	const void* vc;
	__half* Ohi;    // [T][q_dim]
	__half* Olo;
	int n_tokens, pos0, seq_len, q_dim, kv_mul;
	float inv_sqrt_hd;
};

// DPL (2 or 4) consecutive cache elements -> floats with a single shared-memory load
template <int DPL>
__device__ __forceinline__ void pf_load_dims(const __half* p, float (&o)[DPL]) {
	if constexpr (DPL == 4) {
		const uint2 r = *reinterpret_cast<const uint2*>(p);
		const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
		o[0] = a.x, o[1] = a.y, o[2] = b.x, o[3] = b.y;
	} else {
		const float2 a = __half22float2(*reinterpret_cast<const __half2*>(p));
		o[0] = a.x, o[1] = a.y;
	}
}
template <int DPL>
__device__ __forceinline__ void pf_load_dims(const uint8_t* p, float (&o)[DPL]) {
	if constexpr (DPL == 4) {
		const uint32_t r = *reinterpret_cast<const uint32_t*>(p);
		const float2 a = e5m2x2_lo(r), b = e5m2x2_hi(r);
		o[0] = a.x, o[1] = a.y, o[2] = b.x, o[3] = b.y;
	} else {
		const uint32_t r = *reinterpret_cast<const unsigned short*>(p);
		const float2 a = e5m2x2_lo(r);
		o[0] = a.x, o[1] = a.y;
	}
}

// CTA = (kv head, 8 consecutive query tokens); warp = one token with all kv_mul query heads of the kv head (KM of them per pass)
template <typename KVT, int HD, int KM>
__global__ void __launch_bounds__(PFA_WARPS * 32) k_pf_attn(const PfAttnArgs a) {
	pdl_enter();
	constexpr int DPL = HD / 32;               // head dims owned by a lane in the value pass
	constexpr int KSTR = HD * sizeof(KVT) + 16; // padded key row: 16-byte reads of one row per lane are conflict-free per quarter warp
	__shared__ __align__(16) unsigned char ks[PFA_TILE * KSTR];
	__shared__ __align__(16) unsigned char vs[PFA_TILE * HD * sizeof(KVT)];
	__shared__ __align__(16) float qs[PFA_WARPS][KM * HD];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int kvh = blockIdx.y, t = blockIdx.x * PFA_WARPS + warp;
	const bool live = t < a.n_tokens;
	const int last = a.pos0 + min(a.n_tokens - 1, (int)blockIdx.x * PFA_WARPS + PFA_WARPS - 1); // last key position any warp of the CTA needs
	const int mypos = a.pos0 + t;
	const KVT* kbase = reinterpret_cast<const KVT*>(a.kc) + (size_t)kvh * a.seq_len * HD;
	const KVT* vbase = reinterpret_cast<const KVT*>(a.vc) + (size_t)kvh * a.seq_len * HD;
	for (int hg = 0; hg < a.kv_mul; hg += KM) { // KM query heads per pass over the keys
		const int h0 = kvh * a.kv_mul + hg;
		for (int i = lane; i < KM * HD; i += 32) qs[warp][i] = live ? a.Q[(size_t)t * a.q_dim + (size_t)h0 * HD + i] : 0.f;
		float m[KM], l[KM], acc[KM][DPL];
#pragma unroll
		for (int h = 0; h < KM; ++h) {
			m[h] = -FLT_MAX, l[h] = 0.f;
#pragma unroll
			for (int d = 0; d < DPL; ++d) acc[h][d] = 0.f;
		}
		for (int k0 = 0; k0 <= last; k0 += PFA_TILE) {
			__syncthreads(); // the previous tile is no longer being read
			for (int i = threadIdx.x; i < PFA_TILE * HD * (int)sizeof(KVT) / 16; i += blockDim.x) {
				const int row = i / (HD * (int)sizeof(KVT) / 16), c = i % (HD * (int)sizeof(KVT) / 16);
				const bool in = k0 + row <= last;
				const uint4 z = make_uint4(0, 0, 0, 0);
				*reinterpret_cast<uint4*>(ks + row * KSTR + c * 16) = in ? __ldcg(reinterpret_cast<const uint4*>(kbase + (size_t)(k0 + row) * HD) + c) : z;
				*reinterpret_cast<uint4*>(vs + row * HD * sizeof(KVT) + c * 16) = in ? __ldcg(reinterpret_cast<const uint4*>(vbase + (size_t)(k0 + row) * HD) + c) : z;
			}
			__syncthreads();
			if (!live || k0 > mypos) continue; // (warp-uniform) causal: nothing of this tile is visible to this token
			const bool vis = k0 + lane <= mypos;
			// scores of key (k0 + lane) against the KM heads
			float sc[KM];
#pragma unroll
			for (int h = 0; h < KM; ++h) sc[h] = 0.f;
			const unsigned char* krow = ks + lane * KSTR;
#pragma unroll 4
			for (int c = 0; c < HD / 8; ++c) {
				float kf[8];
				kv_load8(reinterpret_cast<const KVT*>(krow) + c * 8, kf);
#pragma unroll
				for (int h = 0; h < KM; ++h) {
					const float4 q0 = *reinterpret_cast<const float4*>(&qs[warp][h * HD + c * 8]), q1 = *reinterpret_cast<const float4*>(&qs[warp][h * HD + c * 8 + 4]);
					float s = sc[h];
					s = fmaf(q0.x, kf[0], s), s = fmaf(q0.y, kf[1], s), s = fmaf(q0.z, kf[2], s), s = fmaf(q0.w, kf[3], s);
					s = fmaf(q1.x, kf[4], s), s = fmaf(q1.y, kf[5], s), s = fmaf(q1.z, kf[6], s), s = fmaf(q1.w, kf[7], s);
					sc[h] = s;
				}
			}
			float p[KM];
#pragma unroll
			for (int h = 0; h < KM; ++h) {
				const float s = vis ? sc[h] * a.inv_sqrt_hd : -FLT_MAX;
				const float mn = fmaxf(m[h], warp_max(s));
				const float corr = expf(m[h] - mn);
				p[h] = vis ? expf(s - mn) : 0.f;
				l[h] = fmaf(l[h], corr, warp_sum(p[h]));
				m[h] = mn;
#pragma unroll
				for (int d = 0; d < DPL; ++d) acc[h][d] *= corr;
			}
			// values: lane owns head dims [lane * DPL, lane * DPL + DPL)
			const int nk = min(PFA_TILE, mypos - k0 + 1);
			for (int j = 0; j < nk; ++j) {
				float vf[DPL];
				pf_load_dims<DPL>(reinterpret_cast<const KVT*>(vs) + (size_t)j * HD + lane * DPL, vf); // one 8- / 4- / 2-byte load
#pragma unroll
				for (int h = 0; h < KM; ++h) {
					const float pj = __shfl_sync(0xffffffffu, p[h], j);
#pragma unroll
					for (int d = 0; d < DPL; ++d) acc[h][d] = fmaf(pj, vf[d], acc[h][d]);
				}
			}
		}
		if (live) {
#pragma unroll
			for (int h = 0; h < KM; ++h) {
				const float inv = 1.0f / l[h];
#pragma unroll
				for (int d = 0; d < DPL; ++d) {
					__half hh, hl;
					split_hl(acc[h][d] * inv, hh, hl);
					const size_t o = (size_t)t * a.q_dim + (size_t)(h0 + h) * HD + lane * DPL + d;
					a.Ohi[o] = hh, a.Olo[o] = hl;
				}
			}
		}
		__syncwarp();
	}
}
