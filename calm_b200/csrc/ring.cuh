// ring.cuh -- the two largest matvec stages (FFN up, FFN down; and wo) fed by warp-private bulk-TMA rings.
//
// Why (profiles/README.md, round 1 + 2): the register-fed stage kernels keep their bytes in flight in registers --
// a warp requests one batch (8 x 16 B per lane), waits a DRAM round trip, consumes it, and only then requests the next --
// so the bytes in flight per SM are duty-cycled, bounded by the register file, and nothing can be requested before the
// dependency wait without costing a CTA of occupancy.  Here every warp owns a small ring of shared-memory slots that it
// fills itself with cp.async.bulk (one elected lane, completion on the slot's mbarrier): requests run NS chunks ahead of
// the arithmetic, continuously, across task boundaries, and the first NS chunks are requested BEFORE griddepcontrol.wait
// (weights are immutable), i.e. while the previous kernel drains and the activation vector is staged.  There is no
// producer warp and no CTA-wide hand-off: a warp waits only on its own mbarriers.
//
// Work split: one CTA owns a contiguous range of "tasks"; its warps take tasks from a shared-memory counter, so all warps
// of an SM finish within one chunk of each other whatever the task count (the static row-pair stride of k_ffn_up costs a
// nearly empty fifth round: 14336 pairs over 3552 warps).  A task is a row pair (R = 2 rows share every activation load)
// restricted to a K-slice of S chunks; kernels with few long rows (w2: 2048 pairs of 14336 weights) use S = 1 and
// fold the slices of a pair in slice order (deterministic) through shared memory -- the last slice to finish does it.
//
// Arithmetic is dot_vec<> of common.cuh on the same vectors in the same per-lane order as warp_dot_rows, so results are
// bit-identical to the register-fed kernels for whole rows.  Dense models on one GPU with row bytes a multiple of 512;
// everything else (MoE expert offsets known only after the router, the in-kernel tensor-parallel exchange, odd shapes)
// keeps k_ffn_up / k_matres.
#pragma once

#include "stages.cuh"

#define RING_MAX_WARPS 16 // warps per CTA: 8 (two CTAs per SM) or 16 (one CTA per SM when the activation vector is long: w2)
#define RING_MAX_NS 4
#define RING_MAX_PAIRS 48 // row pairs per CTA when K-slices are folded through shared memory
#define RING_MAX_SLICES 16

struct RingCtl {
	uint64_t bar[RING_MAX_WARPS][RING_MAX_NS];
	int slot_task[RING_MAX_WARPS][RING_MAX_NS];
	int slot_piece[RING_MAX_WARPS][RING_MAX_NS];
	int ctr; // next task of this CTA
};

// shared memory of a ring kernel: [32 floats block-reduce][activation vector][ring: warps x NS x 2 rows x CH bytes]
template <int DBITS>
__host__ __device__ inline size_t ring_smem_bytes(int n, int u, int ns, int warps) {
	size_t xs = ((size_t)(32 + xs_floats<DBITS>(n)) * sizeof(float) + 127) & ~(size_t)127;
	return xs + (size_t)warps * ns * 2 * u * 512;
}

// gf4 consume of a 2-row slot on the tensor cores.  SIMT gf4 costs ~3.3 issue slots per half-byte weight (shift + LOP3 + FFMA and the
// fp32 activation reads) -- issue-bound at about a third of the HBM roofline even when ring-fed.  Here the two rows of the slot are
// columns 0 / 1 of the B operand of mma.m16n8k16 ((q - 4) * scale as exact f16: at most 6 significant bits) and the activation
// vector, split into an f16 hi / lo pair (22 bits after a power-of-two pre-scale), is rows 0 / 1 of A: lane (g, t) decodes code
// pair t of two consecutive words of row (g & 1) per k-step and supplies x pairs when g < 2 -- ~18 instructions per 32 weights.
// D[0][n] + D[1][n] = row n . x.  The slots are the same 2 KB-per-row bulk copies as for the other formats (a variant that
// gave each MMA row its own 256-byte copy starved the ring: profiles/README.md).
__device__ __forceinline__ uint32_t gf4_pair_scaled(uint32_t w, int sh, __half2 sc) {
	const uint32_t x = w >> sh;
	uint32_t h = (x & 7u) | ((x & 0x38u) << 13) | 0x64006400u; // f16 integers 1024 + q
	const __half2 v = __hmul2(__hsub2(*reinterpret_cast<__half2*>(&h), __half2half2(__ushort_as_half(0x6404))), sc);
	return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void mma_16816_rows01(float (&d)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) { // A rows 8..15 are zero
	asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%5}, {%7,%8}, {%0,%1,%2,%3};"
	             : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
	             : "r"(a0), "r"(0u), "r"(a2), "r"(b0), "r"(b1));
}

// The streaming loop of one warp.  rows(task, rp0, rp1, chunk0): byte pointers of the task's two rows and its first chunk
// index inside the row; done(task, v0, v1): the two finished dot products (called by every lane, values valid in all).
// issue_prologue() must be called (warp-uniformly) before consume_all(); in between the caller waits for the previous grid
// and stages the activation vector.
template <int DBITS, int U, int NS, typename RowsFn, typename DoneFn>
struct RingWarp {
	static constexpr int CH = U * 512; // bytes per row and chunk
	RingCtl* ctl;
	unsigned char* ring; // this warp's NS slots
	int warp, lane;
	int t_hi, S; // end of the CTA's task range; chunks per task
	uint64_t policy;
	int issued = 0, consumed = 0, cur_task = 0, cur_piece = 0;
	bool more = true;
	RowsFn rows;
	DoneFn done;

	__device__ __forceinline__ RingWarp(RingCtl* c, unsigned char* ring_base, int t_hi_, int S_, RowsFn r, DoneFn d)
	    : ctl(c), warp(threadIdx.x >> 5), lane(threadIdx.x & 31), t_hi(t_hi_), S(S_), rows(r), done(d) {
		ring = ring_base + (size_t)warp * NS * 2 * CH;
		policy = l2_policy_evict_first();
		cur_piece = S; // forces a task grab
	}

	__device__ __forceinline__ void issue_one() { // warp-uniform control flow
		if (cur_piece == S) {
			int t = 0;
			if (lane == 0) t = atomicAdd(&ctl->ctr, 1);
			t = __shfl_sync(0xffffffffu, t, 0);
			if (t >= t_hi) {
				more = false;
				return;
			}
			cur_task = t, cur_piece = 0;
		}
		const int slot = issued % NS;
		if (lane == 0) {
			const unsigned char *rp0, *rp1;
			int chunk0;
			rows(cur_task, rp0, rp1, chunk0);
			const size_t off = (size_t)(chunk0 + cur_piece) * CH;
			ctl->slot_task[warp][slot] = cur_task, ctl->slot_piece[warp][slot] = cur_piece;
			uint64_t* bar = &ctl->bar[warp][slot];
			mbar_expect_tx(bar, 2 * CH);
			tma_load_1d_hint(ring + (size_t)slot * 2 * CH, rp0 + off, CH, bar, policy);
			tma_load_1d_hint(ring + (size_t)slot * 2 * CH + CH, rp1 + off, CH, bar, policy);
		}
		++issued, ++cur_piece;
	}

	__device__ __forceinline__ void issue_prologue() {
#pragma unroll
		for (int i = 0; i < NS; ++i)
			if (more) issue_one();
		__syncwarp();
	}

	__device__ __forceinline__ void consume_all(const float4* __restrict__ xs4) {
		constexpr int Q = WFmt<DBITS>::VW / 4;
		float acc0 = 0.f, acc1 = 0.f;
		while (consumed < issued) {
			const int slot = consumed % NS;
			mbar_wait(&ctl->bar[warp][slot], (consumed / NS) & 1);
			const int task = ctl->slot_task[warp][slot], piece = ctl->slot_piece[warp][slot];
			int chunk0;
			{
				const unsigned char *rp0, *rp1;
				rows(task, rp0, rp1, chunk0);
			}
			const uint4* s0 = reinterpret_cast<const uint4*>(ring + (size_t)slot * 2 * CH);
			const uint4* s1 = s0 + CH / 16;
			const float4* xp = xs4 + (size_t)((chunk0 + piece) * U) * Q * 32 + lane; // chunk = U groups of 32 vectors
			uint4 w0[U], w1[U];
#pragma unroll
			for (int u = 0; u < U; ++u) w0[u] = s0[32 * u + lane], w1[u] = s1[32 * u + lane];
#pragma unroll
			for (int u = 0; u < U; ++u) {
				float4 xv[Q];
#pragma unroll
				for (int q = 0; q < Q; ++q) xv[q] = xp[(u * Q + q) * 32];
				acc0 = dot_vec<DBITS>(w0[u], xv, acc0);
				acc1 = dot_vec<DBITS>(w1[u], xv, acc1);
			}
			++consumed;
			__syncwarp(); // every lane has read the slot (its values are in registers): the slot may be refilled
			if (more) issue_one();
			if (piece == S - 1) {
				const float v0 = warp_sum(acc0), v1 = warp_sum(acc1);
				done(task, v0, v1);
				acc0 = acc1 = 0.f;
			}
		}
	}

	// gf4 on the tensor cores (see gf4_pair_scaled above).  xh / xl: the activation vector as f16 pairs (u32), hi and lo;
	// out_scale: 2^e of the pre-scale.  The format's -1/4 (infer.c:37-40) is applied to the sums.
	__device__ __forceinline__ void consume_all_g4(const uint32_t* __restrict__ xh, const uint32_t* __restrict__ xl, float out_scale) {
		static_assert(DBITS == 4, "gf4 only");
		const int g = lane >> 2, t = lane & 3;
		const uint32_t* xa = g == 0 ? xh : xl; // A row g: 0 = hi, 1 = lo, the rest zero
		const bool arow = g < 2;
		const int sh = 8 + 6 * t;
		float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
		while (consumed < issued) {
			const int slot = consumed % NS;
			mbar_wait(&ctl->bar[warp][slot], (consumed / NS) & 1);
			const int task = ctl->slot_task[warp][slot], piece = ctl->slot_piece[warp][slot];
			int chunk0;
			{
				const unsigned char *rp0, *rp1;
				rows(task, rp0, rp1, chunk0);
			}
			const uint4* srow = reinterpret_cast<const uint4*>(ring + (size_t)slot * 2 * CH + (size_t)(g & 1) * CH); // B column g: row g & 1 of the slot
			const uint32_t* xp = xa + (size_t)(chunk0 + piece) * CH; // CH bytes = 2 CH weights = CH f16 pairs
#pragma unroll 4
			for (int q = 0; q < CH / 16; ++q) { // 4 words = 32 weights of the row: two k-steps
				const uint4 w = srow[q];
				uint32_t s0 = __byte_perm(w.x, 0, 0x0404), s1 = __byte_perm(w.y, 0, 0x0404), s2 = __byte_perm(w.z, 0, 0x0404), s3 = __byte_perm(w.w, 0, 0x0404);
				const uint32_t b00 = gf4_pair_scaled(w.x, sh, *reinterpret_cast<__half2*>(&s0)), b01 = gf4_pair_scaled(w.y, sh, *reinterpret_cast<__half2*>(&s1));
				const uint32_t b10 = gf4_pair_scaled(w.z, sh, *reinterpret_cast<__half2*>(&s2)), b11 = gf4_pair_scaled(w.w, sh, *reinterpret_cast<__half2*>(&s3));
				const uint32_t a00 = arow ? xp[16 * q + t] : 0u, a02 = arow ? xp[16 * q + 4 + t] : 0u;
				const uint32_t a10 = arow ? xp[16 * q + 8 + t] : 0u, a12 = arow ? xp[16 * q + 12 + t] : 0u;
				mma_16816_rows01(d0, a00, a02, b00, b01);
				mma_16816_rows01(d1, a10, a12, b10, b11);
			}
			++consumed;
			__syncwarp();
			if (more) issue_one();
			if (piece == S - 1) {
				// lane (g, 0) holds D[g][0], D[g][1]: hi (g = 0) and lo (g = 1) parts of the two rows
				const float r0 = d0[0] + d1[0], r1 = d0[1] + d1[1];
				const float k = -0.25f * out_scale;
				const float v0 = (__shfl_sync(0xffffffffu, r0, 0) + __shfl_sync(0xffffffffu, r0, 4)) * k;
				const float v1 = (__shfl_sync(0xffffffffu, r1, 0) + __shfl_sync(0xffffffffu, r1, 4)) * k;
				done(task, v0, v1);
#pragma unroll
				for (int i = 0; i < 4; ++i) d0[i] = d1[i] = 0.f;
			}
		}
	}
};

template <int NS>
__device__ __forceinline__ void ring_init(RingCtl* ctl, int t_lo) {
	if (threadIdx.x == 0) ctl->ctr = t_lo;
	if (threadIdx.x < (blockDim.x >> 5) * NS) mbar_init(&ctl->bar[threadIdx.x / NS][threadIdx.x % NS], 1);
	if (threadIdx.x == 0) mbar_init_fence();
	__syncthreads();
}

// ------------------------------------------------------------------------------------------------
// FFN up: hb[i] = act(w1[i] . xn) * (w3[i] . xn)   (reference infer.c:437-450); task i = rows (w1[i], w3[i]), whole rows.

template <int DBITS, int U, int NS>
__global__ void __launch_bounds__(256, 2) k_ffn_up_ring(const FfnUpArgs a) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ RingCtl ctl;
	float* red = reinterpret_cast<float*>(smem_raw);
	float* xs = red + 32;
	unsigned char* ring = smem_raw + (((size_t)(32 + xs_floats<DBITS>(a.dim)) * sizeof(float) + 127) & ~(size_t)127);
	const size_t rowbytes = (size_t)a.dim * DBITS / 8;
	const int cpt = (int)(rowbytes / (U * 512));
	const int t_lo = (int)(((long long)blockIdx.x * a.hidden) / gridDim.x), t_hi = (int)(((long long)(blockIdx.x + 1) * a.hidden) / gridDim.x);
	pdl_launch_next();
	ring_init<NS>(&ctl, t_lo);
	float post = 1.f;
	const unsigned char* w1 = reinterpret_cast<const unsigned char*>(a.w1);
	const unsigned char* w3 = reinterpret_cast<const unsigned char*>(a.w3);
	auto rows = [&](int t, const unsigned char*& rp0, const unsigned char*& rp1, int& chunk0) {
		rp0 = w1 + (size_t)t * rowbytes, rp1 = w3 + (size_t)t * rowbytes, chunk0 = 0;
	};
	auto done = [&](int t, float v0, float v1) {
		if ((threadIdx.x & 31) == 0) {
			const float u1 = v0 * post, u3 = v1 * post;
			a.hb[t] = (a.gelu ? act_gelu(u1) : act_silu(u1)) * u3;
		}
	};
	RingWarp<DBITS, U, NS, decltype(rows), decltype(done)> rw(&ctl, ring, t_hi, cpt, rows, done);
	rw.issue_prologue(); // the first NS chunks of every warp are in flight before the previous kernel has finished
	pdl_wait_prev();
	stamp_begin(a.stamp);
	if constexpr (DBITS == 4) { // the activation vector as f16 hi / lo pairs for the tensor-core consume (norm applied while staging)
		uint2* H = reinterpret_cast<uint2*>(xs);
		uint2* Lo = H + a.dim / 4;
		const float out_scale = (a.dim / 4 <= (int)blockDim.x * 4) ? stage_vector_h<4>(H, Lo, red, a.x, a.dim, a.normw, a.eps, a.ln != 0)
		                                                           : stage_vector_h_long(H, Lo, red, a.x, a.dim, a.normw, a.eps, a.ln != 0);
		rw.consume_all_g4(reinterpret_cast<const uint32_t*>(H), reinterpret_cast<const uint32_t*>(Lo), out_scale);
	} else {
		post = stage_vector<DBITS>(xs, red, a.x, a.dim, a.normw, a.eps, a.ln != 0, nullptr);
		rw.consume_all(reinterpret_cast<const float4*>(xs));
	}
	stamp_end(a.stamp);
}

// ------------------------------------------------------------------------------------------------
// y[row] += W[row] . xin  (wo, w2; dense).  Task = (row pair p, K-slice s of S chunks), t = p * nsl + s; with nsl > 1 the
// slices of a pair land in part[][] and the last one to arrive adds them up in slice order.

template <int DBITS, int U, int NS>
__global__ void __launch_bounds__(512, 1) k_matres_ring(const MatResArgs a, const int S) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ RingCtl ctl;
	__shared__ float2 part[RING_MAX_PAIRS][RING_MAX_SLICES];
	__shared__ int cnt[RING_MAX_PAIRS];
	float* red = reinterpret_cast<float*>(smem_raw);
	float* xs = red + 32;
	unsigned char* ring = smem_raw + (((size_t)(32 + xs_floats<DBITS>(a.n)) * sizeof(float) + 127) & ~(size_t)127);
	const size_t rowbytes = (size_t)a.n * DBITS / 8;
	const int cpt = (int)(rowbytes / (U * 512)), nsl = cpt / S;
	const int npairs = a.d / 2;
	const int p_lo = (int)(((long long)blockIdx.x * npairs) / gridDim.x), p_hi = (int)(((long long)(blockIdx.x + 1) * npairs) / gridDim.x);
	pdl_launch_next();
	if (threadIdx.x < RING_MAX_PAIRS) cnt[threadIdx.x] = 0;
	ring_init<NS>(&ctl, p_lo * nsl);
	const unsigned char* w = reinterpret_cast<const unsigned char*>(a.w);
	auto rows = [&](int t, const unsigned char*& rp0, const unsigned char*& rp1, int& chunk0) {
		const int p = t / nsl, s = t - p * nsl;
		rp0 = w + (size_t)(2 * p) * rowbytes, rp1 = rp0 + rowbytes, chunk0 = s * S;
	};
	auto done = [&](int t, float v0, float v1) {
		if ((threadIdx.x & 31) != 0) return;
		const int p = t / nsl, s = t - p * nsl;
		if (nsl > 1) {
			const int pl = p - p_lo;
			part[pl][s] = make_float2(v0, v1);
			__threadfence_block();
			if (atomicAdd(&cnt[pl], 1) != nsl - 1) return;
			__threadfence_block();
			v0 = 0.f, v1 = 0.f;
			for (int k = 0; k < nsl; ++k) v0 += part[pl][k].x, v1 += part[pl][k].y; // slice order: deterministic
		}
		float2* dst = reinterpret_cast<float2*>(a.y + 2 * p);
		float2 cur = a.accumulate ? *dst : make_float2(0.f, 0.f);
		cur.x += v0, cur.y += v1;
		*dst = cur;
	};
	RingWarp<DBITS, U, NS, decltype(rows), decltype(done)> rw(&ctl, ring, p_hi * nsl, S, rows, done);
	rw.issue_prologue();
	pdl_wait_prev();
	stamp_begin(a.stamp);
	if constexpr (DBITS == 4) {
		uint2* H = reinterpret_cast<uint2*>(xs);
		uint2* Lo = H + a.n / 4;
		const float out_scale = (a.n / 4 <= (int)blockDim.x * 4) ? stage_vector_h<4>(H, Lo, red, a.xin, a.n, nullptr, 0.f, false)
		                                                         : stage_vector_h_long(H, Lo, red, a.xin, a.n, nullptr, 0.f, false);
		rw.consume_all_g4(reinterpret_cast<const uint32_t*>(H), reinterpret_cast<const uint32_t*>(Lo), out_scale);
	} else {
		stage_vector<DBITS, 16>(xs, red, a.xin, a.n, nullptr, 0.f, false, nullptr);
		rw.consume_all(reinterpret_cast<const float4*>(xs));
	}
	stamp_end(a.stamp);
}
