// fused.cuh -- the "fused" engine: ONE persistent kernel runs a whole token.
//
// Why: at batch 1 a Llama-3-8B token is ~165 stage executions of 2-18 us of HBM time each, so launch
// gaps, pipeline fill and wave tails -- not bandwidth -- decide the roofline fraction of a kernel-per-stage
// design.  Here the weights never stop streaming:
//
//   * grid = one CTA per SM (cooperative launch), 1 producer warp + 16 consumer warps;
//   * every CTA owns a fixed, contiguous row range of every matrix, so its share of the model is a
//     STATIC list of contiguous byte ranges ("tiles");
//   * the producer lane walks that list for the whole token and moves tile after tile into a ring of
//     shared-memory slots with 1-D bulk TMA (cp.async.bulk + mbarrier complete_tx).  It never looks at
//     activations, so it runs ahead across stage boundaries: while the consumers sit in a grid barrier
//     the ring (up to ~200 KB per SM, ~30 MB chip-wide) keeps filling with the next stages' weights;
//   * consumers keep their slice of the activation vector IN REGISTERS (a warp group of WG warps owns a
//     row, each thread a fixed set of 16-byte column vectors), read weights from the ring with
//     conflict-free 16-byte shared loads, dequantise in registers, FMA in fp32, reduce by shuffles, and
//     the last warp to finish a tile folds the per-warp partial sums in a fixed order (deterministic)
//     and runs the stage epilogue (bias/clip/RoPE/KV append, residual, SiLU-gate, logits/argmax);
//   * attention streams K/V tiles through the same ring (flash-decoding split over positions; the
//     entry written this step is read through the generic path after the barrier instead).
//
// Stage order and arithmetic are those of stages.cuh (reference infer.c:311-472).  Five grid barriers per
// layer (after QKV, attention, wo, FFN-up, FFN-down); MoE models, fp8 KV caches and the rolled-over
// cache (pos >= seq_len) are served by the staged engine.
#pragma once

#include <cooperative_groups.h>

#include "stages.cuh"

#define FUSED_NCW 16                         // consumer warps (4 per scheduler)
#define FUSED_AW 8                           // attention: warps per half (each half serves half of the query heads)
#define FUSED_THREADS ((FUSED_NCW + 1) * 32) // + 1 producer warp
#define FUSED_MAX_SLOTS 16
#define FUSED_SPIN_LIMIT (1u << 27)          // watchdog for every spin loop (~1 s)

struct FusedLayer {
	const void *wq, *wk, *wv, *wo, *w1, *w2, *w3;
	const float *rms_att, *rms_ffn, *bqkv;
};

struct FusedArgs {
	int dim, hidden, q_dim, kv_dim, head_dim, n_heads, n_kv_heads, n_layers, vocab, seq_len, kv_mul;
	float eps, clip;
	int ln, norm_par, gelu;
	float *x, *xb, *q, *att, *hb, *logits;
	float* attn_partial;
	unsigned* attn_counter;
	__half *kc, *vc;
	const float* rope_freq;
	const void* embed;
	const void* wcls;
	const float* rms_final;
	const TokenParams* tp;
	unsigned* bar; // grid barrier word
	int* err;      // watchdog report: nonzero = which wait timed out
	unsigned long long* perf; // optional [4][8] ns per stage, seen by thread 0 of CTA 0: {busy, grid-barrier wait, activation load, waiting for tiles}
	                          // (cf. reference coopstage, infer.cu:390-402)
	float* cand_val;
	int* cand_idx;
	int mode;      // 0 kv only, >= 1 logits (cand_val != NULL: also greedy candidates)
	int dbg;       // experiments (results are wrong): 1 = consumers skip the math, 2 = producer skips the copies
	int slot_bytes, nslots;
	int window;    // at most this many tiles requested but not yet landed (bounds the L2->SM queue that demand loads wait behind)
	int attn_nsplit, attn_hg, attn_qgroups, attn_lpp, attn_scratch_bytes;
	int xbuf_bytes, nwbuf_bytes; // shared staging: activation vector (stages whose rows are shared by < all warps), norm weights
	float inv_sqrt_hd;
};

__constant__ FusedLayer c_fused_layers[MAX_LAYERS];

// ---------------------------------------------------------------- PTX wrappers

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
	return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
	uint32_t ok;
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
	    "selp.u32 %0, 1, 0, p;\n\t}"
	    : "=r"(ok)
	    : "r"(smem_u32(bar)), "r"(parity)
	    : "memory");
	return ok != 0;
}
// bulk global -> shared copy, completion counted on an mbarrier; weights are read once: evict-first in L2
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
	             "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
	             : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
	uint64_t p;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
	return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
	uint64_t p;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
	return p;
}
__device__ __forceinline__ void consumer_sync() { // named barrier 1: the 8 consumer warps only
	asm volatile("bar.sync 1, %0;" ::"n"(FUSED_NCW * 32) : "memory");
}
__device__ __forceinline__ uint4 lds128(const void* p) {
	uint4 r;
	asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_u32(p)));
	return r;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

__device__ __noinline__ void fused_fail(int* err, int code) {
	atomicCAS(err, 0, code);
	__threadfence_system();
	__trap();
}

#define SPIN_WAIT(cond, err, code)                          \
	do {                                                    \
		unsigned spins_ = 0;                                \
		while (!(cond)) {                                   \
			if (++spins_ > FUSED_SPIN_LIMIT) fused_fail(err, code); \
		}                                                   \
	} while (0)

// Self-resetting grid barrier (counter gains exactly 2^31 per round: CTA 0 adds 2^31 - (G-1), the
// others add 1; a round is over when bit 31 flips).  Split into arrive / wait so the caller can place
// independent work in between.  Called by one thread per CTA.
__device__ __forceinline__ unsigned grid_arrive(unsigned* bar) {
	unsigned nb = blockIdx.x == 0 ? 0x80000000u - (gridDim.x - 1) : 1u;
	unsigned old;
	__threadfence();
	asm volatile("atom.add.release.gpu.u32 %0, [%1], %2;" : "=r"(old) : "l"(bar), "r"(nb) : "memory");
	return old;
}
__device__ __forceinline__ void grid_wait(unsigned* bar, unsigned old, int* err, int code) {
	unsigned cur;
	unsigned spins = 0;
	do {
		asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(cur) : "l"(bar) : "memory");
		if (++spins > FUSED_SPIN_LIMIT) fused_fail(err, code);
	} while (((old ^ cur) & 0x80000000u) == 0);
}

// ---------------------------------------------------------------- shared-memory layout
#define FUSED_NP_MAX 2

struct FusedShared {
	uint64_t full[FUSED_MAX_SLOTS];  // producer -> consumers: tile landed (tx bytes)
	uint64_t empty[FUSED_MAX_SLOTS]; // consumers -> producer: tile consumed (FUSED_NCW arrivals)
	int cnt[FUSED_MAX_SLOTS];        // warps done with the tile in this slot (last one finalises)
	float red[FUSED_MAX_SLOTS][FUSED_NP_MAX * FUSED_NCW]; // per-row, per-warp partial sums: [vr * wg + warp_in_group]
	float rope_cos[128], rope_sin[128];        // cos/sin(pos * freq[i]) of this token (head_dim <= 256)
	float scratch[64];               // block reductions
	float best_val[FUSED_NCW];       // greedy candidates per warp
	int best_idx[FUSED_NCW];
	int flag;
};

// ---------------------------------------------------------------- tile schedule (shared by producer and consumers)

__device__ __forceinline__ void cta_range(int units, int& u0, int& u1) {
	u0 = (int)(((long long)blockIdx.x * units) / gridDim.x);
	u1 = (int)(((long long)(blockIdx.x + 1) * units) / gridDim.x);
}

// How the 8 consumer warps share a row of `nvec` 16-byte vectors: the smallest power-of-two number of
// warps per row (wg) that keeps the vectors per thread (it) within the register budget; the remaining
// factor (ng = 8 / wg warp groups) works on different rows of a tile at the same time.
struct RowMap {
	int wg, ng, it, tg; // warps per group, groups, vectors per thread, threads per group
};
template <int DBITS, int XR>
__device__ __forceinline__ RowMap row_map(int nvec) {
	constexpr int ITMAX = XR / WFmt<DBITS>::VW;
	RowMap m;
	m.wg = 1;
	while ((nvec + m.wg * 32 - 1) / (m.wg * 32) > ITMAX && m.wg < FUSED_NCW) m.wg *= 2;
	m.ng = FUSED_NCW / m.wg;
	m.tg = m.wg * 32;
	m.it = (nvec + m.tg - 1) / m.tg;
	return m;
}

// rows per tile (per segment): as many as fit the slot, at most FUSED_NP rows per warp group and tile,
// at most 32 rows in all, a multiple of `unit`
#define FUSED_NP FUSED_NP_MAX
__device__ __forceinline__ int tile_rows(int slot_bytes, int nseg, int rowbytes, int unit, int ng) {
	int r = slot_bytes / (nseg * rowbytes);
	if (r > 32 / nseg) r = 32 / nseg;
	if (r > FUSED_NP * ng / nseg) r = FUSED_NP * ng / nseg;
	r -= r % unit;
	return r < unit ? unit : r;
}

// A matrix stage streams up to three row ranges (q | k | v) of matrices with the same row length.
struct StageRanges {
	int n;
	int r0[3], r1[3];
};

// Ring bookkeeping common to both sides: slot index and phase of the k-th tile of this CTA.
struct RingPos {
	int slot;
	uint32_t phase;
	int nslots;
	__device__ __forceinline__ void init(int n) { slot = 0, phase = 0, nslots = n; }
	__device__ __forceinline__ void advance() {
		if (++slot == nslots) slot = 0, phase ^= 1;
	}
};

// ---------------------------------------------------------------- producer

struct Producer {
	FusedShared* sh;
	char* ring;
	int slot_bytes;
	RingPos rp;
	RingPos lp;      // `window` tiles behind rp: the oldest tile that may still be in flight
	int ahead;       // tiles issued since lp
	int window;
	uint64_t pol_w, pol_kv;
	int* err;
	int dbg;

	// one tile = one ring slot: `bytes` from src0 (and from src1, placed at dst + off1)
	__device__ __forceinline__ void push(const char* src0, const char* src1, uint32_t bytes, uint32_t off1, uint64_t policy) {
		if (ahead == window) { // keep the request queue short: wait for the oldest outstanding tile to land
			SPIN_WAIT(mbar_try_wait(&sh->full[lp.slot], lp.phase), err, 103);
			lp.advance();
			--ahead;
		}
		++ahead;
		SPIN_WAIT(mbar_try_wait(&sh->empty[rp.slot], rp.phase ^ 1), err, 101);
		if (dbg & 2) {
			mbar_arrive(&sh->full[rp.slot]);
		} else {
			mbar_expect_tx(&sh->full[rp.slot], src1 ? 2 * bytes : bytes);
			char* dst = ring + (size_t)rp.slot * slot_bytes;
			tma_load_1d(dst, src0, bytes, &sh->full[rp.slot], policy);
			if (src1) tma_load_1d(dst + off1, src1, bytes, &sh->full[rp.slot], policy);
		}
		rp.advance();
	}

	// all tiles of rows [r0, r1) of one matrix (or of two matrices read in lock step)
	__device__ __forceinline__ void matrix(const void* w0, const void* w1, int rowbytes, int r0, int r1, int R) {
		for (int a = r0; a < r1; a += R) {
			uint32_t bytes = (uint32_t)min(R, r1 - a) * rowbytes;
			push((const char*)w0 + (size_t)a * rowbytes, w1 ? (const char*)w1 + (size_t)a * rowbytes : nullptr, bytes, bytes, pol_w);
		}
	}
};

// ---------------------------------------------------------------- consumer: matvec over ring tiles

template <int DBITS, int XR>
struct Consumer {
	static constexpr int VW = WFmt<DBITS>::VW;
	static constexpr int ITMAX = XR / VW;
	static constexpr int Q = VW / 4;

	FusedShared* sh;
	const char* ring;
	int slot_bytes;
	int* err;
	int dbg;
	int cw, lane; // consumer warp 0..7
	unsigned long long* tile_wait; // thread 0 of CTA 0 only: accumulates ns spent waiting for tiles
	float* xbuf;                   // shared: one coalesced copy of the activation vector per CTA
	const float* nwbuf;            // shared: norm weights of this stage, prefetched before the previous barrier
	float xr[ITMAX][VW];

	// sum over all 256 consumer threads (every thread gets it)
	__device__ __forceinline__ float block_total(float v) {
		v = warp_sum(v);
		consumer_sync();
		if (lane == 0) sh->scratch[cw] = v;
		consumer_sync();
		float r = lane < FUSED_NCW ? sh->scratch[lane] : 0.f;
		return warp_sum(r);
	}

	// --- activation slice into registers -------------------------------------------------------
	// xin: global vector written by other CTAs in this launch (read through L2).  When several warp groups
	// need the same vector it is fetched ONCE per CTA with coalesced loads into shared memory and
	// distributed from there (all SMs read the same few KB right after a barrier: every redundant or
	// half-used sector costs microseconds).  Optional norm (reference infer.c:183-207) with weights that
	// were prefetched into shared memory before the previous grid barrier.
	__device__ __forceinline__ void load_x(const RowMap& m, const float* xin, int n, bool norm, float eps, bool ln, float* xb_out) {
		const int nvec = n / VW;
		const int tgi = (cw % m.wg) * 32 + lane;
		const int tid = cw * 32 + lane;
		const bool staged = m.ng > 1;
		if (staged) {
			const float4* src = reinterpret_cast<const float4*>(xin);
			float4* dst = reinterpret_cast<float4*>(xbuf);
			for (int i = tid; i < n / 4; i += FUSED_NCW * 32) dst[i] = __ldcg(src + i);
			consumer_sync();
		}
		float ssum = 0.f, ssq = 0.f;
#pragma unroll
		for (int it = 0; it < ITMAX; ++it) {
			int v = tgi + it * m.tg;
			bool ok = it < m.it && v < nvec;
#pragma unroll
			for (int q = 0; q < Q; ++q) {
				float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
				if (ok) f = staged ? *(reinterpret_cast<const float4*>(xbuf + (size_t)v * VW) + q) : __ldcg(reinterpret_cast<const float4*>(xin + (size_t)v * VW) + q);
				xr[it][q * 4 + 0] = f.x, xr[it][q * 4 + 1] = f.y, xr[it][q * 4 + 2] = f.z, xr[it][q * 4 + 3] = f.w;
				ssum += (f.x + f.y) + (f.z + f.w);
			}
		}
		if (!norm) return;
		float mean = 0.f;
		if (ln) mean = block_total(ssum) / (float)(m.ng * n);
#pragma unroll
		for (int it = 0; it < ITMAX; ++it) {
			int v = tgi + it * m.tg;
			bool ok = it < m.it && v < nvec;
#pragma unroll
			for (int e = 0; e < VW; ++e) {
				float d = ok ? xr[it][e] - mean : 0.f;
				ssq = fmaf(d, d, ssq);
			}
		}
		float var = block_total(ssq) / (float)(m.ng * n); // every group holds one full copy of the vector
		float scale = 1.0f / sqrtf(var + eps);
#pragma unroll
		for (int it = 0; it < ITMAX; ++it) {
			int v = tgi + it * m.tg;
			bool ok = it < m.it && v < nvec;
			if (ok) {
#pragma unroll
				for (int q = 0; q < Q; ++q) {
					float4 w = *(reinterpret_cast<const float4*>(nwbuf + (size_t)v * VW) + q);
					float* xp = &xr[it][q * 4];
					xp[0] = (xp[0] - mean) * scale * w.x, xp[1] = (xp[1] - mean) * scale * w.y;
					xp[2] = (xp[2] - mean) * scale * w.z, xp[3] = (xp[3] - mean) * scale * w.w;
					if (xb_out && cw / m.wg == 0) *(reinterpret_cast<float4*>(xb_out + (size_t)v * VW) + q) = make_float4(xp[0], xp[1], xp[2], xp[3]);
				}
			}
		}
	}

	// --- one matrix stage ---------------------------------------------------------------------
	// Consumes the tiles the producer pushed for the row ranges of `rg` (R rows per tile, nseg segments).
	// epi(range, row, v0, v1): called by one lane per row (v1 = second segment's value when nseg == 2); for
	// PAIR stages it is called once per even row with (range, row, value(row), value(row + 1)).
	template <typename Epi>
	__device__ __forceinline__ void matrix(RingPos& rp, const RowMap& m, int nvec, int nseg, bool PAIR, const StageRanges& rg, int R, Epi epi) {
		constexpr int NP = FUSED_NP;
		const int grp = cw / m.wg, wig = cw % m.wg;
		const int tgi = wig * 32 + lane;
		const int rowbytes = nvec * 16;
		for (int ri = 0; ri < rg.n; ++ri) {
			for (int a = rg.r0[ri]; a < rg.r1[ri]; a += R) {
				const int rows = min(R, rg.r1[ri] - a);
				const int vrows = rows * nseg; // <= NP * ng by construction of R
				if (tile_wait) {
					unsigned long long tw = globaltimer_ns();
					SPIN_WAIT(mbar_try_wait(&sh->full[rp.slot], rp.phase), err, 201);
					*tile_wait += globaltimer_ns() - tw;
				} else {
					SPIN_WAIT(mbar_try_wait(&sh->full[rp.slot], rp.phase), err, 201);
				}
				const char* tile = ring + (size_t)rp.slot * slot_bytes;
				if (!(dbg & 1)) {
					// this group's rows of the tile: vr = grp + k * ng.  All loads first, then the math, then one
					// transposing shuffle reduction for the NP row sums (zero weights contribute exactly 0).
					uint4 w[NP][ITMAX];
#pragma unroll
					for (int k = 0; k < NP; ++k) {
						const int vr = grp + k * m.ng;
						const uint4* rowp = reinterpret_cast<const uint4*>(tile + (size_t)vr * rowbytes);
#pragma unroll
						for (int it = 0; it < ITMAX; ++it) {
							int v = tgi + it * m.tg;
							w[k][it] = (vr < vrows && it < m.it && v < nvec) ? lds128(rowp + v) : make_uint4(0, 0, 0, 0);
						}
					}
					// branch-free math (rows / vectors beyond the tile carry zero weights): NP * 2 independent FMA chains
					float acc[NP][2];
#pragma unroll
					for (int k = 0; k < NP; ++k) acc[k][0] = 0.f, acc[k][1] = 0.f;
#pragma unroll
					for (int it = 0; it < ITMAX; ++it) {
						float4 xv[Q];
#pragma unroll
						for (int q = 0; q < Q; ++q) xv[q] = make_float4(xr[it][q * 4], xr[it][q * 4 + 1], xr[it][q * 4 + 2], xr[it][q * 4 + 3]);
#pragma unroll
						for (int k = 0; k < NP; ++k) acc[k][it & 1] = dot_vec<DBITS>(w[k][it], xv, acc[k][it & 1]);
					}
					// two sums -> lanes 0 and 16 (5 shuffles instead of 10)
					{
						const bool hi = lane & 16;
						const float a0 = acc[0][0] + acc[0][1], a1 = acc[1][0] + acc[1][1];
						float c = (hi ? a1 : a0) + __shfl_xor_sync(0xffffffffu, hi ? a0 : a1, 16);
						c += __shfl_xor_sync(0xffffffffu, c, 8);
						c += __shfl_xor_sync(0xffffffffu, c, 4);
						c += __shfl_xor_sync(0xffffffffu, c, 2);
						c += __shfl_xor_sync(0xffffffffu, c, 1);
						const int vr = grp + (lane >> 4) * m.ng;
						if ((lane & 15) == 0 && vr < vrows) sh->red[rp.slot][vr * m.wg + wig] = c;
					}
				}
				// this warp is done reading the tile; find out whether it is the last one
				__syncwarp();
				int last = 0;
				if (lane == 0) {
					__threadfence_block();
					last = atomicAdd(&sh->cnt[rp.slot], 1) == FUSED_NCW - 1;
				}
				last = __shfl_sync(0xffffffffu, last, 0);
				if (last) {
					__threadfence_block();
					if (lane == 0) sh->cnt[rp.slot] = 0;
					if (!(dbg & 1)) {
						// fold the per-warp partials in a fixed order and run the epilogue
						const int ncalls = PAIR ? rows / 2 : rows;
						for (int i = lane; i < ncalls; i += 32) {
							int r = PAIR ? 2 * i : i;
							float v0 = 0.f, v1 = 0.f;
							const int rb = PAIR ? r + 1 : rows + r; // second value: next row, or same row of segment 2
							for (int w_ = 0; w_ < m.wg; ++w_) v0 += sh->red[rp.slot][r * m.wg + w_];
							if (PAIR || nseg == 2)
								for (int w_ = 0; w_ < m.wg; ++w_) v1 += sh->red[rp.slot][rb * m.wg + w_];
							epi(ri, a + r, v0, v1);
						}
					}
					__syncwarp();
				}
				if (lane == 0) mbar_arrive(&sh->empty[rp.slot]);
				rp.advance();
			}
		}
	}
};

// ---------------------------------------------------------------- consumer: attention over ring tiles
// Same arithmetic as k_attn (stages.cuh).  One work item = (unit of HG query heads sharing a kv head,
// slice [t0, t1) of positions).  K and V of 'TP' positions arrive per ring slot (K at the slot base, V at
// TP * kvrow); the entry of the current step (written by stage 1 of this launch, possibly after the tile was
// prefetched) is skipped in the tiles and read through L2 in one extra pass instead.

__device__ __forceinline__ void halves8(const uint4& r, float (&o)[8]) {
	float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
	float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
	float2 c = __half22float2(*reinterpret_cast<const __half2*>(&r.z));
	float2 d = __half22float2(*reinterpret_cast<const __half2*>(&r.w));
	o[0] = a.x, o[1] = a.y, o[2] = b.x, o[3] = b.y, o[4] = c.x, o[5] = c.y, o[6] = d.x, o[7] = d.y;
}

// HH = query heads per warp.  The 16 consumer warps form two halves of 8: both halves walk the same
// positions of a tile (position p belongs to warp p-slot `wq` of each half), half 0 serves the first HH
// heads of the unit, half 1 the rest -- half the registers per thread, K/V read twice from shared memory.
template <int HH>
__device__ __noinline__ RingPos fused_attention(const FusedArgs& a, RingPos rp, FusedShared* sh, const char* ring, float* scratch, const TokenParams& tp,
                                                const __half* kc_l, const __half* vc_l, int unit, int split, int kvh, int t0, int t1, int TP, int warp, int lane) {
	constexpr int P = 2;
	const int HG = a.attn_hg;
	const int hd = a.head_dim, lpp = a.attn_lpp;
	const int G = 32 / lpp;
	const int grp = lane / lpp, li = lane % lpp;
	const bool dact = li * 8 < hd;
	const int kvrow = hd * 2;
	const int half = warp / FUSED_AW, wq = warp % FUSED_AW;
	const int h0 = half * HH;                               // first head (within the unit) of this warp
	const int nh = max(0, min(HH, HG - h0));                // heads this warp really has
	const int hbase = kvh * a.kv_mul + (unit % a.attn_qgroups) * HG;
	const int tid = warp * 32 + lane;

	float qr[HH][8], acc[HH][8], m[HH], l[HH];
#pragma unroll
	for (int h = 0; h < HH; ++h) {
		m[h] = -FLT_MAX, l[h] = 0.f;
		float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
		if (dact && h < nh) {
			const float4* qp = reinterpret_cast<const float4*>(a.q + (size_t)(hbase + h0 + h) * hd + li * 8);
			q0 = __ldcg(qp), q1 = __ldcg(qp + 1);
		}
		qr[h][0] = q0.x, qr[h][1] = q0.y, qr[h][2] = q0.z, qr[h][3] = q0.w, qr[h][4] = q1.x, qr[h][5] = q1.y, qr[h][6] = q1.z, qr[h][7] = q1.w;
#pragma unroll
		for (int d = 0; d < 8; ++d) acc[h][d] = 0.f;
	}

	const bool fresh_here = tp.kv_pos >= t0 && tp.kv_pos < t1;
	const int ntiles = t1 > t0 ? (t1 - t0 + TP - 1) / TP : 0;
	for (int ti = 0; ti < ntiles + (fresh_here ? 1 : 0); ++ti) {
		const bool fresh = ti == ntiles; // extra pass: the entry appended this step, straight from L2
		const int tb = t0 + ti * TP;
		const int np = fresh ? 1 : min(TP, t1 - tb);
		const char *kt = nullptr, *vt = nullptr;
		if (!fresh) {
			SPIN_WAIT(mbar_try_wait(&sh->full[rp.slot], rp.phase), a.err, 202);
			kt = ring + (size_t)rp.slot * a.slot_bytes;
			vt = kt + (size_t)TP * kvrow;
		}
		for (int pb = 0; pb < ((a.dbg & 1) ? 0 : np); pb += FUSED_AW * G * P) { // warp-uniform trip count
			float kf[P][8], vf[P][8];
			bool ok[P];
#pragma unroll
			for (int i = 0; i < P; ++i) {
				int p = pb + (i * FUSED_AW + wq) * G + grp;
				ok[i] = p < np && (fresh || (tb + p) != tp.kv_pos);
				if (ok[i] && dact) {
					if (fresh) {
						size_t off = ((size_t)kvh * a.seq_len + tp.kv_pos) * hd + li * 8;
						halves8(__ldcg(reinterpret_cast<const uint4*>(kc_l + off)), kf[i]);
						halves8(__ldcg(reinterpret_cast<const uint4*>(vc_l + off)), vf[i]);
					} else {
						halves8(lds128(kt + (size_t)p * kvrow + li * 16), kf[i]);
						halves8(lds128(vt + (size_t)p * kvrow + li * 16), vf[i]);
					}
				} else {
#pragma unroll
					for (int d = 0; d < 8; ++d) kf[i][d] = 0.f, vf[i][d] = 0.f;
				}
			}
			// online-softmax update with P (position, K, V) triples held by this lane group
#pragma unroll
			for (int h = 0; h < HH; ++h) {
				float sc[P], smax = m[h];
#pragma unroll
				for (int i = 0; i < P; ++i) {
					float d = 0.f;
#pragma unroll
					for (int e = 0; e < 8; ++e) d = fmaf(qr[h][e], kf[i][e], d);
					for (int o = 1; o < lpp; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
					sc[i] = ok[i] ? d * a.inv_sqrt_hd : -FLT_MAX;
					smax = fmaxf(smax, sc[i]);
				}
				float corr = expf(m[h] - smax);
				m[h] = smax;
				float pw[P], ps = 0.f;
#pragma unroll
				for (int i = 0; i < P; ++i) {
					pw[i] = ok[i] ? expf(sc[i] - smax) : 0.f;
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
		if (!fresh) {
			__syncwarp();
			if (lane == 0) mbar_arrive(&sh->empty[rp.slot]);
			rp.advance();
		}
	}

	// merge lane groups of a warp
	for (int o = lpp; o < 32; o <<= 1) {
#pragma unroll
		for (int h = 0; h < HH; ++h) {
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
	// merge warps: scratch[wq][head][hd + 2] (the two halves write disjoint heads)
	const int rec = hd + 2;
	if (grp == 0) {
#pragma unroll
		for (int h = 0; h < HH; ++h) {
			if (h < nh) {
				float* r = scratch + ((size_t)wq * HG + h0 + h) * rec;
				if (dact) {
#pragma unroll
					for (int e = 0; e < 8; ++e) r[li * 8 + e] = acc[h][e];
				}
				if (li == 0) r[hd] = m[h], r[hd + 1] = l[h];
			}
		}
	}
	consumer_sync();
	float* part = a.attn_partial + ((size_t)unit * a.attn_nsplit + split) * HG * rec;
	for (int idx = tid; idx < HG * rec; idx += FUSED_NCW * 32) {
		int h = idx / rec, e = idx % rec;
		float mn = -FLT_MAX;
		for (int w = 0; w < FUSED_AW; ++w) mn = fmaxf(mn, scratch[((size_t)w * HG + h) * rec + hd]);
		float v;
		if (e == hd) {
			v = mn;
		} else {
			v = 0.f;
			for (int w = 0; w < FUSED_AW; ++w) {
				const float* r = scratch + ((size_t)w * HG + h) * rec;
				v += r[e] * expf(r[hd] - mn);
			}
		}
		__stcg(part + idx, v);
	}
	__threadfence();
	consumer_sync();
	if (tid == 0) {
		unsigned old = atomicAdd(a.attn_counter + unit, 1u);
		sh->flag = (old == (unsigned)a.attn_nsplit - 1);
	}
	consumer_sync();
	if (!sh->flag) return rp;

	// Last slice of this unit to finish: fold all slices.  Two passes so that every global load is
	// independent of the previous one: (1) maxima and sums of all slices -> coefficients in shared memory,
	// (2) weighted sum of the accumulators, four slices in flight per thread.
	__threadfence();
	const int ns = a.attn_nsplit;
	const float* pk = a.attn_partial + (size_t)unit * ns * HG * rec;
	float* coef = scratch;                  // [ns][HG]  l_s, then exp(m_s - M)
	float* msv = scratch + (size_t)ns * HG; // [ns][HG]  m_s
	for (int i = tid; i < ns * HG; i += FUSED_NCW * 32) {
		const float* r = pk + (size_t)i * rec; // i = s * HG + h
		coef[i] = __ldcg(r + hd + 1);
		msv[i] = __ldcg(r + hd);
	}
	consumer_sync();
	if (tid < HG) {
		float M = -FLT_MAX;
		for (int s_ = 0; s_ < ns; ++s_) M = fmaxf(M, msv[s_ * HG + tid]);
		float L = 0.f;
		for (int s_ = 0; s_ < ns; ++s_) {
			float c = expf(msv[s_ * HG + tid] - M);
			L = fmaf(coef[s_ * HG + tid], c, L);
			coef[s_ * HG + tid] = c;
		}
		sh->scratch[tid] = 1.0f / L; // HG <= 8
	}
	consumer_sync();
	for (int idx = tid; idx < HG * hd; idx += FUSED_NCW * 32) {
		int h = idx / hd, e = idx % hd;
		float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
		int s_ = 0;
		for (; s_ + 4 <= ns; s_ += 4) {
			float p0 = __ldcg(pk + ((size_t)(s_ + 0) * HG + h) * rec + e), p1 = __ldcg(pk + ((size_t)(s_ + 1) * HG + h) * rec + e);
			float p2 = __ldcg(pk + ((size_t)(s_ + 2) * HG + h) * rec + e), p3 = __ldcg(pk + ((size_t)(s_ + 3) * HG + h) * rec + e);
			n0 = fmaf(p0, coef[(s_ + 0) * HG + h], n0), n1 = fmaf(p1, coef[(s_ + 1) * HG + h], n1);
			n2 = fmaf(p2, coef[(s_ + 2) * HG + h], n2), n3 = fmaf(p3, coef[(s_ + 3) * HG + h], n3);
		}
		for (; s_ < ns; ++s_) n0 = fmaf(__ldcg(pk + ((size_t)s_ * HG + h) * rec + e), coef[s_ * HG + h], n0);
		__stcg(a.att + (size_t)(hbase + h) * hd + e, ((n0 + n1) + (n2 + n3)) * sh->scratch[h]);
	}
	if (tid == 0) a.attn_counter[unit] = 0;
	return rp;
}

// ---------------------------------------------------------------- stages (one noinline function each:
// bounded register live ranges and code size; the activation slice lives only inside its stage)

struct StageCtx {
	FusedShared* sh;
	const char* ring;
	float* attn_scratch;
	float* xbuf;
	float* nwbuf;
	int warp, lane;
};

// Norm weights of the NEXT normed stage -> shared memory; issued before a grid barrier so the loads are
// off the critical path after it.
__device__ __forceinline__ void prefetch_normw(const StageCtx& cx, const float* w, int n) {
	if (!w) return;
	const float4* src = reinterpret_cast<const float4*>(w);
	float4* dst = reinterpret_cast<float4*>(cx.nwbuf);
	for (int i = cx.warp * 32 + cx.lane; i < n / 4; i += FUSED_NCW * 32) dst[i] = __ldg(src + i);
}

__device__ __forceinline__ unsigned long long stage_begin(const FusedArgs& a) {
	return (a.perf && blockIdx.x == 0 && threadIdx.x == 0) ? globaltimer_ns() : 0ull;
}
__device__ __forceinline__ void stage_end(const FusedArgs& a, int st, unsigned long long t0) { // stage without a barrier
	if (a.perf && blockIdx.x == 0 && threadIdx.x == 0) a.perf[st] += globaltimer_ns() - t0;
}

// all consumer warps finished the stage -> publish; then wait for every CTA
__device__ __forceinline__ void stage_barrier(const FusedArgs& a, int code, int st, unsigned long long t0) {
	consumer_sync();
	if (threadIdx.x == 0) {
		unsigned long long t1 = (a.perf && blockIdx.x == 0) ? globaltimer_ns() : 0ull;
		unsigned old = grid_arrive(a.bar);
		grid_wait(a.bar, old, a.err, code);
		if (a.perf && blockIdx.x == 0) {
			unsigned long long t2 = globaltimer_ns();
			a.perf[st] += t1 - t0;
			a.perf[8 + st] += t2 - t1;
		}
	}
	consumer_sync();
}

// One function serves every matrix stage (the tile loop, the activation load and the epilogues exist
// once in the instruction stream: the per-layer code footprint has to stay inside the instruction cache).
enum StageKind { SK_QKV = 1, SK_WO = 3, SK_UP = 4, SK_DOWN = 5, SK_OUT = 6 };

template <int DBITS, int XR>
__device__ __noinline__ RingPos stage_matrix(const FusedArgs& a, const StageCtx cx, RingPos rp, const int kind, const int l, const TokenParams tp) {
	constexpr int VW = WFmt<DBITS>::VW;
	const unsigned long long t_begin = stage_begin(a);
	const FusedLayer& L = c_fused_layers[l];
	FusedShared* sh = cx.sh;
	Consumer<DBITS, XR> c;
	c.sh = cx.sh, c.ring = cx.ring, c.slot_bytes = a.slot_bytes, c.err = a.err, c.dbg = a.dbg;
	c.cw = cx.warp, c.lane = cx.lane;
	c.xbuf = cx.xbuf, c.nwbuf = cx.nwbuf;
	c.tile_wait = (a.perf && blockIdx.x == 0 && threadIdx.x == 0) ? a.perf + 24 + kind : nullptr;

	// what this stage reads, which rows it owns, how its tiles are cut
	const float* xin = a.x;
	const float* normw = nullptr;
	int n = a.dim, nseg = 1, unit = 1;
	StageRanges rg;
	rg.n = 1;
	switch (kind) {
	case SK_QKV: // norm -> q,k,v (+bias, clip, RoPE) -> q vector / cache append   (reference infer.c:352-381)
		normw = L.rms_att, unit = 2, rg.n = 3;
		cta_range(a.q_dim / 2, rg.r0[0], rg.r1[0]), rg.r0[0] *= 2, rg.r1[0] *= 2;
		cta_range(a.kv_dim / 2, rg.r0[1], rg.r1[1]), rg.r0[1] *= 2, rg.r1[1] *= 2;
		rg.r0[2] = rg.r0[1], rg.r1[2] = rg.r1[1];
		break;
	case SK_WO: // x += wo . att   (reference infer.c:410-415)
		xin = a.att, n = a.q_dim;
		cta_range(a.dim, rg.r0[0], rg.r1[0]);
		break;
	case SK_UP: // norm -> act(w1 . xn) * (w3 . xn)   (reference infer.c:417-450)
		if (a.norm_par)
			xin = a.xb;
		else
			normw = L.rms_ffn;
		nseg = 2;
		cta_range(a.hidden, rg.r0[0], rg.r1[0]);
		break;
	case SK_DOWN: // x += w2 . hb   (reference infer.c:452-456)
		xin = a.hb, n = a.hidden;
		cta_range(a.dim, rg.r0[0], rg.r1[0]);
		break;
	default: // SK_OUT: logits = wcls . norm(x), greedy candidates   (reference infer.c:466-469, sampler.c:34-42)
		normw = a.rms_final;
		cta_range(a.vocab, rg.r0[0], rg.r1[0]);
		break;
	}
	const int nv = n / VW;
	const RowMap m = row_map<DBITS, XR>(nv);
	const int R = tile_rows(a.slot_bytes, nseg, nv * 16, unit, m.ng);

	c.load_x(m, xin, n, normw != nullptr, a.eps, a.ln != 0, (kind == SK_QKV && a.norm_par && blockIdx.x == 0) ? a.xb : nullptr);
	if (c.tile_wait) a.perf[16 + kind] += globaltimer_ns() - t_begin;

	const size_t kv_layer = (size_t)a.n_kv_heads * a.seq_len * a.head_dim;
	float best_v = -FLT_MAX;
	int best_i = 0x7fffffff;
	c.matrix(rp, m, nv, nseg, kind == SK_QKV, rg, R, [&](int which, int r, float v0, float v1) {
		switch (kind) {
		case SK_QKV: {
			int j = which == 0 ? r : (which == 1 ? a.q_dim + r : a.q_dim + a.kv_dim + r);
			if (L.bqkv) v0 += L.bqkv[j], v1 += L.bqkv[j + 1];
			v0 = fminf(fmaxf(v0, -a.clip), a.clip);
			v1 = fminf(fmaxf(v1, -a.clip), a.clip);
			if (which < 2) { // rotate the pair (reference infer.c:223-236); cos/sin of this position are in shared memory
				int i = (r % a.head_dim) >> 1;
				float fcr = sh->rope_cos[i], fci = sh->rope_sin[i];
				float r0 = v0 * fcr - v1 * fci, r1 = v0 * fci + v1 * fcr;
				v0 = r0, v1 = r1;
			}
			if (which == 0) {
				__stcg(reinterpret_cast<float2*>(a.q + r), make_float2(v0, v1));
			} else {
				__half* cbase = (which == 1 ? a.kc : a.vc) + l * kv_layer;
				int h = r / a.head_dim, d = r % a.head_dim;
				*reinterpret_cast<__half2*>(cbase + ((size_t)h * a.seq_len + tp.kv_pos) * a.head_dim + d) = __floats2half2_rn(v0, v1);
			}
			break;
		}
		case SK_WO:
		case SK_DOWN:
			__stcg(a.x + r, __ldcg(a.x + r) + v0);
			break;
		case SK_UP:
			__stcg(a.hb + r, (a.gelu ? act_gelu(v0) : act_silu(v0)) * v1);
			break;
		default:
			a.logits[r] = v0;
			if (v0 > best_v || (v0 == best_v && r < best_i)) best_v = v0, best_i = r; // first maximum wins
			break;
		}
	});

	if (kind != SK_OUT) {
		// norm weights of the next normed stage (this stage's copy was consumed by load_x above)
		if (kind == SK_QKV && !a.norm_par) prefetch_normw(cx, L.rms_ffn, a.dim);
		if (kind == SK_UP) prefetch_normw(cx, l + 1 < a.n_layers ? c_fused_layers[l + 1].rms_att : (a.mode != 0 ? a.rms_final : nullptr), a.dim);
		stage_barrier(a, 300 + kind, kind, t_begin);
		return rp;
	}
	if (a.cand_val) {
		float bv = best_v;
		int bi = best_i;
		for (int o = 16; o > 0; o >>= 1) {
			float ov = __shfl_xor_sync(0xffffffffu, bv, o);
			int oi = __shfl_xor_sync(0xffffffffu, bi, o);
			if (ov > bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
		}
		if (cx.lane == 0) sh->best_val[cx.warp] = bv, sh->best_idx[cx.warp] = bi;
		consumer_sync();
		if (threadIdx.x == 0) {
			bv = -FLT_MAX, bi = 0x7fffffff;
			for (int w = 0; w < FUSED_NCW; ++w)
				if (sh->best_val[w] > bv || (sh->best_val[w] == bv && sh->best_idx[w] < bi)) bv = sh->best_val[w], bi = sh->best_idx[w];
			a.cand_val[blockIdx.x] = bv, a.cand_idx[blockIdx.x] = bi;
		}
	}
	stage_end(a, 6, t_begin);
	return rp;
}

// ---------------------------------------------------------------- the kernel

template <int DBITS, int XR>
__global__ void __launch_bounds__(FUSED_THREADS, 1) k_fused(const __grid_constant__ FusedArgs a) {
	extern __shared__ __align__(128) char smem_raw[];
	FusedShared* sh = reinterpret_cast<FusedShared*>(smem_raw);
	float* attn_scratch = reinterpret_cast<float*>(smem_raw + ((sizeof(FusedShared) + 127) & ~(size_t)127));
	float* xbuf = reinterpret_cast<float*>(reinterpret_cast<char*>(attn_scratch) + a.attn_scratch_bytes);
	float* nwbuf = reinterpret_cast<float*>(reinterpret_cast<char*>(xbuf) + a.xbuf_bytes);
	char* ring = reinterpret_cast<char*>(nwbuf) + a.nwbuf_bytes;

	constexpr int VW = WFmt<DBITS>::VW;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const TokenParams tp = *a.tp;

	if (threadIdx.x == 0) {
		for (int i = 0; i < a.nslots; ++i) {
			mbar_init(&sh->full[i], 1);
			mbar_init(&sh->empty[i], FUSED_NCW);
			sh->cnt[i] = 0;
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	// RoPE angles of this token, once (host-libm frequencies, reference infer.c:225-229)
	for (int i = threadIdx.x; i < a.head_dim / 2; i += blockDim.x) sincosf((float)tp.pos * a.rope_freq[i], &sh->rope_sin[i], &sh->rope_cos[i]);
	__syncthreads();

	// attention work item of this CTA: (unit, split) -> positions [t0, t1) of kv head `kvh`
	const int units = a.n_kv_heads * a.attn_qgroups;
	const int item = blockIdx.x;
	const bool has_item = item < units * a.attn_nsplit;
	const int unit = item / a.attn_nsplit, split = item % a.attn_nsplit;
	const int kvh = unit / a.attn_qgroups;
	const int chunk = (tp.kv_len + a.attn_nsplit - 1) / a.attn_nsplit;
	const int t0 = split * chunk, t1 = min(tp.kv_len, t0 + chunk);
	const int kvrow = a.head_dim * 2;                       // bytes per cached position
	const int TP = a.slot_bytes / (2 * kvrow) < 64 ? a.slot_bytes / (2 * kvrow) : 64; // positions per tile (K + V)
	const size_t kv_layer = (size_t)a.n_kv_heads * a.seq_len * a.head_dim;

	// ================================================================= producer warp
	if (warp == FUSED_NCW) {
		if (lane != 0) return;
		const int nv_dim = a.dim / VW, nv_q = a.q_dim / VW, nv_hid = a.hidden / VW;
		const int rb_dim = nv_dim * 16, rb_q = nv_q * 16, rb_hid = nv_hid * 16; // bytes per row
		const int ng_dim = row_map<DBITS, XR>(nv_dim).ng, ng_q = row_map<DBITS, XR>(nv_q).ng, ng_hid = row_map<DBITS, XR>(nv_hid).ng;
		const int R_qkv = tile_rows(a.slot_bytes, 1, rb_dim, 2, ng_dim);
		const int R_wo = tile_rows(a.slot_bytes, 1, rb_q, 1, ng_q);
		const int R_up = tile_rows(a.slot_bytes, 2, rb_dim, 1, ng_dim);
		const int R_down = tile_rows(a.slot_bytes, 1, rb_hid, 1, ng_hid);
		const int R_out = tile_rows(a.slot_bytes, 1, rb_dim, 1, ng_dim);
		int q0, q1, k0, k1, o0, o1, h0, h1, c0, c1;
		cta_range(a.q_dim / 2, q0, q1), q0 *= 2, q1 *= 2;
		cta_range(a.kv_dim / 2, k0, k1), k0 *= 2, k1 *= 2;
		cta_range(a.dim, o0, o1);
		cta_range(a.hidden, h0, h1);
		cta_range(a.vocab, c0, c1);

		Producer p;
		p.sh = sh, p.ring = ring, p.slot_bytes = a.slot_bytes, p.err = a.err, p.dbg = a.dbg;
		p.rp.init(a.nslots);
		p.lp.init(a.nslots);
		p.ahead = 0, p.window = a.window;
		p.pol_w = l2_policy_evict_first();
		p.pol_kv = l2_policy_evict_last();
		for (int l = 0; l < a.n_layers; ++l) {
			const FusedLayer& L = c_fused_layers[l];
			p.matrix(L.wq, nullptr, rb_dim, q0, q1, R_qkv);
			p.matrix(L.wk, nullptr, rb_dim, k0, k1, R_qkv);
			p.matrix(L.wv, nullptr, rb_dim, k0, k1, R_qkv);
			if (has_item) {
				const char* kb = (const char*)(a.kc + l * kv_layer + (size_t)kvh * a.seq_len * a.head_dim);
				const char* vb = (const char*)(a.vc + l * kv_layer + (size_t)kvh * a.seq_len * a.head_dim);
				for (int t_ = t0; t_ < t1; t_ += TP)
					p.push(kb + (size_t)t_ * kvrow, vb + (size_t)t_ * kvrow, (uint32_t)min(TP, t1 - t_) * kvrow, (uint32_t)TP * kvrow, p.pol_kv);
			}
			p.matrix(L.wo, nullptr, rb_q, o0, o1, R_wo);
			p.matrix(L.w1, L.w3, rb_dim, h0, h1, R_up);
			p.matrix(L.w2, nullptr, rb_hid, o0, o1, R_down);
		}
		if (a.mode != 0) p.matrix(a.wcls, nullptr, rb_dim, c0, c1, R_out);
		return;
	}

	// ================================================================= consumer warps
	StageCtx cx;
	cx.sh = sh, cx.ring = ring, cx.attn_scratch = attn_scratch, cx.xbuf = xbuf, cx.nwbuf = nwbuf, cx.warp = warp, cx.lane = lane;
	prefetch_normw(cx, c_fused_layers[0].rms_att, a.dim);
	RingPos rp;
	rp.init(a.nslots);

	// x = decode(E[token]) (reference infer.c:335-347).  Every CTA writes the whole row (identical values), so
	// it only ever reads back what it wrote itself: no grid barrier is needed before stage 1.
	for (int i = threadIdx.x; i < a.dim; i += FUSED_NCW * 32) a.x[i] = weight_at<DBITS>(a.embed, (size_t)tp.token * a.dim + i);
	__threadfence();
	consumer_sync();

	for (int l = 0; l < a.n_layers; ++l) {
		rp = stage_matrix<DBITS, XR>(a, cx, rp, SK_QKV, l, tp);

		// stage 2: attention over the cache (K/V tiles from the ring)
		const unsigned long long t_attn = stage_begin(a);
		if (has_item) {
			const __half* kc_l = a.kc + l * kv_layer;
			const __half* vc_l = a.vc + l * kv_layer;
			switch ((a.attn_hg + 1) / 2) { // heads per warp
			case 1: rp = fused_attention<1>(a, rp, sh, ring, attn_scratch, tp, kc_l, vc_l, unit, split, kvh, t0, t1, TP, warp, lane); break;
			case 2: rp = fused_attention<2>(a, rp, sh, ring, attn_scratch, tp, kc_l, vc_l, unit, split, kvh, t0, t1, TP, warp, lane); break;
			case 3: rp = fused_attention<3>(a, rp, sh, ring, attn_scratch, tp, kc_l, vc_l, unit, split, kvh, t0, t1, TP, warp, lane); break;
			default: rp = fused_attention<4>(a, rp, sh, ring, attn_scratch, tp, kc_l, vc_l, unit, split, kvh, t0, t1, TP, warp, lane); break;
			}
		}
		stage_barrier(a, 302, 2, t_attn);

		rp = stage_matrix<DBITS, XR>(a, cx, rp, SK_WO, l, tp);
		rp = stage_matrix<DBITS, XR>(a, cx, rp, SK_UP, l, tp);
		rp = stage_matrix<DBITS, XR>(a, cx, rp, SK_DOWN, l, tp);
	}
	if (a.mode != 0) rp = stage_matrix<DBITS, XR>(a, cx, rp, SK_OUT, 0, tp);
}

// ---------------------------------------------------------------- micro-benchmark: cost of one grid barrier
// (same arrive/wait code as the fused kernel, all SMs, 288 threads per CTA, no work in between)
__global__ void __launch_bounds__(FUSED_THREADS, 1) k_barrier_bench(unsigned* bar, int* err, int rounds, unsigned long long* ns_out) {
	const int warp = threadIdx.x >> 5;
	if (warp == FUSED_NCW) return;
	unsigned long long t0 = 0;
	if (blockIdx.x == 0 && threadIdx.x == 0) t0 = globaltimer_ns();
	for (int i = 0; i < rounds; ++i) {
		consumer_sync();
		if (threadIdx.x == 0) {
			unsigned old = grid_arrive(bar);
			grid_wait(bar, old, err, 900);
		}
		consumer_sync();
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) *ns_out = globaltimer_ns() - t0;
}
