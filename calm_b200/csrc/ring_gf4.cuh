// ring_gf4.cuh -- gf4 weights: ring-fed matvec stages with the dot products on the tensor cores.
//
// Why: a gf4 weight is half a byte, so the HBM roofline asks an SM for ~47 weights per clock -- and the SIMT decode costs
// ~3.3 issue slots per weight (shift + LOP3 + FFMA and the fp32 activation reads): issue-bound at a quarter of the roofline
// (round 1: 0.26; Mistral-7B gf4 decoded SLOWER than fp8).  Here a warp owns 16 rows; one mma.sync.m16n8k8 takes ONE gf4
// word (8 weights sharing a scale) of each row: every lane extracts just two pairs of 3-bit codes (its own k columns of
// rows g and g + 8) as exact small f16 integers (q - 4), the B operand carries the activation vector as an f16 hi / lo pair in
// columns 0 / 1 (22 bits after a power-of-two pre-scale, as in k_ffn_up_mma), the 8-term sums come back in fp32, and the
// group scale is applied to the SUM (4 FFMA per word and lane) -- 0.2 issue slots per weight.
//   sum_k w_k x_k = sum_groups (s_g / -4) * sum_{k in g} (q_k - 4) x_k          (reference infer.c:37-40, helpers.cuh:101-113)
// Weights arrive through warp-private bulk-TMA rings as in ring.cuh (one 256-byte piece per row and chunk, 16 copies per
// chunk issued by 16 lanes); rows sit 272 bytes apart in the slot so the 8 row groups of a fragment load hit 8 bank groups.
// Tasks (16-row tile x K-slice of S chunks) are handed out through a shared-memory counter; slices are folded in slice
// order (deterministic) by the last one to finish.  Dense models on one GPU, row bytes a multiple of 256.
#pragma once

#include "ring.cuh"

#define G4_CH 256                 // bytes per row and chunk: 64 words = 512 weights
#define G4_STRIDE (G4_CH + 16)    // row pitch inside a slot
#define G4_SLOT (16 * G4_STRIDE)  // 4352 bytes
#define G4_WARPS 8
#define G4_MAX_SL 32

__device__ __forceinline__ void mma_1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
	asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%7,%7,%7};"
	             : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
	             : "r"(a0), "r"(a1), "r"(b0), "f"(0.f));
}
// codes q_{2t}, q_{2t+1} of a gf4 word as the f16 pair (q - 4): 0x6400 | q is the f16 integer 1024 + q
__device__ __forceinline__ uint32_t gf4_pair(uint32_t w, int sh) {
	const uint32_t x = w >> sh;
	uint32_t h = (x & 7u) | ((x & 0x38u) << 13) | 0x64006400u;
	const __half2 r = __hsub2(*reinterpret_cast<__half2*>(&h), __half2half2(__ushort_as_half(0x6404))); // 1028
	return *reinterpret_cast<const uint32_t*>(&r);
}

// shared memory: [32 floats][x hi: n/4 uint2][x lo: n/4 uint2][ring: warps x NS x G4_SLOT][fold: tiles x nsl x 16 floats][fold counters]
__host__ __device__ inline size_t g4_smem_bytes(int n, int ns, int tiles_per_cta, int nsl) {
	size_t x = ((size_t)32 * 4 + (size_t)n * 4 + 127) & ~(size_t)127;
	size_t fold = nsl > 1 ? (size_t)tiles_per_cta * nsl * 16 * 4 + (size_t)tiles_per_cta * 4 : 0;
	return x + (size_t)G4_WARPS * ns * G4_SLOT + fold + 128;
}

struct G4Ctl {
	uint64_t bar[G4_WARPS][RING_MAX_NS];
	int slot_task[G4_WARPS][RING_MAX_NS];
	int slot_piece[G4_WARPS][RING_MAX_NS];
	int ctr;
};

// The streaming loop of one warp.  rowptr(task, r): byte pointer of row r (0..15) of the task's tile at the start of the task's K-slice;
// xword0(task): index of the first gf4 word of the slice inside a row; done(task, v): v = finished dot product of row g (lanes with t == 0:
// v[0] row g, v[1] row g + 8), already multiplied by the activation pre-scale and -1/4.
template <int NS, typename RowFn, typename WordFn, typename DoneFn>
struct G4Warp {
	G4Ctl* ctl;
	unsigned char* ring;
	int warp, lane, t_hi, S;
	uint64_t policy;
	int issued = 0, consumed = 0, cur_task = 0, cur_piece = 0;
	bool more = true;
	RowFn rowptr;
	WordFn xword0;
	DoneFn done;
	__device__ __forceinline__ G4Warp(G4Ctl* c, unsigned char* ring_base, int t_hi_, int S_, RowFn r, WordFn xw, DoneFn d)
	    : ctl(c), warp(threadIdx.x >> 5), lane(threadIdx.x & 31), t_hi(t_hi_), S(S_), rowptr(r), xword0(xw), done(d) {
		ring = ring_base + (size_t)warp * NS * G4_SLOT;
		policy = l2_policy_evict_first();
		cur_piece = S;
	}
	__device__ __forceinline__ void issue_one() {
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
		uint64_t* bar = &ctl->bar[warp][slot];
		if (lane == 0) {
			ctl->slot_task[warp][slot] = cur_task, ctl->slot_piece[warp][slot] = cur_piece;
			mbar_expect_tx(bar, 16 * G4_CH);
		}
		__syncwarp();
		if (lane < 16) tma_load_1d_hint(ring + (size_t)slot * G4_SLOT + lane * G4_STRIDE, rowptr(cur_task, lane) + (size_t)cur_piece * G4_CH, G4_CH, bar, policy);
		++issued, ++cur_piece;
	}
	__device__ __forceinline__ void issue_prologue() {
#pragma unroll
		for (int i = 0; i < NS; ++i)
			if (more) issue_one();
		__syncwarp();
	}
	// xh: the activation vector as f16 pairs, hi then lo ([n/2] u32 each); out_scale = 2^e of the pre-scale
	__device__ __forceinline__ void consume_all(const uint32_t* __restrict__ xh, const uint32_t* __restrict__ xl, float out_scale) {
		const int g = lane >> 2, t = lane & 3;
		const uint32_t* xb = (g == 1) ? xl : xh; // B column g: 0 = hi, 1 = lo, the rest unused (any finite value)
		const int sh = 8 + 6 * t;
		float acc[4] = {0.f, 0.f, 0.f, 0.f}; // (row g, hi) (row g, lo) (row g+8, hi) (row g+8, lo) on lanes with t == 0
		while (consumed < issued) {
			const int slot = consumed % NS;
			mbar_wait(&ctl->bar[warp][slot], (consumed / NS) & 1);
			const int task = ctl->slot_task[warp][slot], piece = ctl->slot_piece[warp][slot];
			const uint32_t* r0 = reinterpret_cast<const uint32_t*>(ring + (size_t)slot * G4_SLOT + g * G4_STRIDE);
			const uint32_t* r1 = reinterpret_cast<const uint32_t*>(ring + (size_t)slot * G4_SLOT + (g + 8) * G4_STRIDE);
			const uint32_t* xp = xb + (size_t)(xword0(task) + piece * (G4_CH / 4)) * 4 + t; // word w covers x[8w .. 8w+7] = u32 pairs 4w .. 4w+3
#pragma unroll 8
			for (int w = 0; w < G4_CH / 4; ++w) {
				const uint32_t w0 = r0[w], w1 = r1[w];
				float d[4];
				mma_1688(d, gf4_pair(w0, sh), gf4_pair(w1, sh), xp[4 * w]);
				const float s0 = e5m2_to_float((uint8_t)(w0 & 0xff)), s1 = e5m2_to_float((uint8_t)(w1 & 0xff));
				acc[0] = fmaf(s0, d[0], acc[0]), acc[1] = fmaf(s0, d[1], acc[1]);
				acc[2] = fmaf(s1, d[2], acc[2]), acc[3] = fmaf(s1, d[3], acc[3]);
			}
			++consumed;
			__syncwarp();
			if (more) issue_one();
			if (piece == S - 1) {
				const float k = -0.25f * out_scale;
				float v[2] = {(acc[0] + acc[1]) * k, (acc[2] + acc[3]) * k};
				done(task, v);
				acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
			}
		}
	}
};

struct G4Fold { // slices of a tile -> rows, in slice order, by the last slice to arrive
	float* part; // [tiles][nsl][16]
	int* cnt;    // [tiles]
	int nsl;
	// lanes with t == 0 call; returns true (and the folded values in v) for the finisher
	__device__ __forceinline__ bool add(int tile_local, int s, int g, float (&v)[2]) {
		if (nsl == 1) return true;
		float* p = part + ((size_t)tile_local * nsl + s) * 16;
		p[g] = v[0], p[g + 8] = v[1];
		__threadfence_block();
		int old = 0;
		if (g == 0) old = atomicAdd(&cnt[tile_local], 1);
		old = __shfl_sync(0x11111111u, old, 0); // the 8 lanes with t == 0
		if (old != nsl - 1) return false;
		__threadfence_block();
		v[0] = v[1] = 0.f;
		for (int k = 0; k < nsl; ++k) v[0] += part[((size_t)tile_local * nsl + k) * 16 + g], v[1] += part[((size_t)tile_local * nsl + k) * 16 + g + 8];
		return true;
	}
};

template <int NS>
__device__ __forceinline__ void g4_init(G4Ctl* ctl, int t_lo, int* cnt, int ntiles_local) {
	if (threadIdx.x == 0) ctl->ctr = t_lo;
	if (threadIdx.x < G4_WARPS * NS) mbar_init(&ctl->bar[threadIdx.x / NS][threadIdx.x % NS], 1);
	for (int i = threadIdx.x; i < ntiles_local; i += blockDim.x) cnt[i] = 0;
	if (threadIdx.x == 0) mbar_init_fence();
	__syncthreads();
}

// FFN up, gf4: tile = rows [8 tile, 8 tile + 8) of w1 (fragment rows 0..7) and of w3 (fragment rows 8..15): lane (g, t = 0) ends up with
// w1[8 tile + g] . x and w3[8 tile + g] . x.  Task = (tile, K-slice of S chunks).
template <int NS>
__global__ void __launch_bounds__(G4_WARPS * 32, 2) k_ffn_up_g4(const FfnUpArgs a, const int S, const int tiles_per_cta) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ G4Ctl ctl;
	float* red = reinterpret_cast<float*>(smem_raw);
	uint2* H = reinterpret_cast<uint2*>(red + 32);
	uint2* Lo = H + a.dim / 4;
	unsigned char* ring = smem_raw + (((size_t)32 * 4 + (size_t)a.dim * 4 + 127) & ~(size_t)127);
	const size_t rowbytes = (size_t)a.dim / 2;
	const int cpt = (int)(rowbytes / G4_CH), nsl = cpt / S, ntiles = a.hidden / 8;
	const int tile_lo = (int)(((long long)blockIdx.x * ntiles) / gridDim.x), tile_hi = (int)(((long long)(blockIdx.x + 1) * ntiles) / gridDim.x);
	float* part = reinterpret_cast<float*>(ring + (size_t)G4_WARPS * NS * G4_SLOT);
	int* cnt = reinterpret_cast<int*>(part + (size_t)(nsl > 1 ? tiles_per_cta * nsl * 16 : 0));
	pdl_launch_next();
	g4_init<NS>(&ctl, tile_lo * nsl, cnt, tile_hi - tile_lo);
	const unsigned char* w1 = reinterpret_cast<const unsigned char*>(a.w1);
	const unsigned char* w3 = reinterpret_cast<const unsigned char*>(a.w3);
	auto rowptr = [&](int task, int r) {
		const int tile = task / nsl, s = task - tile * nsl;
		return (r < 8 ? w1 + (size_t)(tile * 8 + r) * rowbytes : w3 + (size_t)(tile * 8 + r - 8) * rowbytes) + (size_t)s * S * G4_CH;
	};
	auto xword0 = [&](int task) { return (task % nsl) * S * (G4_CH / 4); };
	G4Fold fold = {part, cnt, nsl};
	auto done = [&](int task, float (&v)[2]) {
		const int lane = threadIdx.x & 31;
		if (lane & 3) return;
		const int tile = task / nsl, s = task - tile * nsl, g = lane >> 2;
		if (!fold.add(tile - tile_lo, s, g, v)) return;
		a.hb[tile * 8 + g] = (a.gelu ? act_gelu(v[0]) : act_silu(v[0])) * v[1];
	};
	G4Warp<NS, decltype(rowptr), decltype(xword0), decltype(done)> rw(&ctl, ring, tile_hi * nsl, S, rowptr, xword0, done);
	rw.issue_prologue();
	pdl_wait_prev();
	stamp_begin(a.stamp);
	float out_scale;
	if (a.dim / 4 <= (int)blockDim.x * 4) out_scale = stage_vector_h<4>(H, Lo, red, a.x, a.dim, a.normw, a.eps, a.ln != 0);
	else out_scale = stage_vector_h_long(H, Lo, red, a.x, a.dim, a.normw, a.eps, a.ln != 0);
	rw.consume_all(reinterpret_cast<const uint32_t*>(H), reinterpret_cast<const uint32_t*>(Lo), out_scale);
	stamp_end(a.stamp);
}

// y[row] += W[row] . xin (wo, w2), gf4: tile = 16 consecutive rows.
template <int NS>
__global__ void __launch_bounds__(G4_WARPS * 32, 1) k_matres_g4(const MatResArgs a, const int S, const int tiles_per_cta) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ G4Ctl ctl;
	float* red = reinterpret_cast<float*>(smem_raw);
	uint2* H = reinterpret_cast<uint2*>(red + 32);
	uint2* Lo = H + a.n / 4;
	unsigned char* ring = smem_raw + (((size_t)32 * 4 + (size_t)a.n * 4 + 127) & ~(size_t)127);
	const size_t rowbytes = (size_t)a.n / 2;
	const int cpt = (int)(rowbytes / G4_CH), nsl = cpt / S, ntiles = a.d / 16;
	const int tile_lo = (int)(((long long)blockIdx.x * ntiles) / gridDim.x), tile_hi = (int)(((long long)(blockIdx.x + 1) * ntiles) / gridDim.x);
	float* part = reinterpret_cast<float*>(ring + (size_t)G4_WARPS * NS * G4_SLOT);
	int* cnt = reinterpret_cast<int*>(part + (size_t)(nsl > 1 ? tiles_per_cta * nsl * 16 : 0));
	pdl_launch_next();
	g4_init<NS>(&ctl, tile_lo * nsl, cnt, tile_hi - tile_lo);
	const unsigned char* w = reinterpret_cast<const unsigned char*>(a.w);
	auto rowptr = [&](int task, int r) {
		const int tile = task / nsl, s = task - tile * nsl;
		return w + (size_t)(tile * 16 + r) * rowbytes + (size_t)s * S * G4_CH;
	};
	auto xword0 = [&](int task) { return (task % nsl) * S * (G4_CH / 4); };
	G4Fold fold = {part, cnt, nsl};
	auto done = [&](int task, float (&v)[2]) {
		const int lane = threadIdx.x & 31;
		if (lane & 3) return;
		const int tile = task / nsl, s = task - tile * nsl, g = lane >> 2;
		if (!fold.add(tile - tile_lo, s, g, v)) return;
		float* y = a.y + tile * 16;
		y[g] = (a.accumulate ? y[g] : 0.f) + v[0];
		y[g + 8] = (a.accumulate ? y[g + 8] : 0.f) + v[1];
	};
	G4Warp<NS, decltype(rowptr), decltype(xword0), decltype(done)> rw(&ctl, ring, tile_hi * nsl, S, rowptr, xword0, done);
	rw.issue_prologue();
	pdl_wait_prev();
	stamp_begin(a.stamp);
	float out_scale;
	if (a.n / 4 <= (int)blockDim.x * 4) out_scale = stage_vector_h<4>(H, Lo, red, a.xin, a.n, nullptr, 0.f, false);
	else out_scale = stage_vector_h_long(H, Lo, red, a.xin, a.n, nullptr, 0.f, false);
	rw.consume_all(reinterpret_cast<const uint32_t*>(H), reinterpret_cast<const uint32_t*>(Lo), out_scale);
	stamp_end(a.stamp);
}
