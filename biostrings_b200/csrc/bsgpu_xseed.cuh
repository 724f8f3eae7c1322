// bsgpu_xseed.cuh -- exact whole-pattern dictionaries of width w >= 25: the byte-seed scan with
// a SHORT core and EXTENDED seeds, so that the subject is probed every S = w - 9 letters
// (16 for 25-mers) instead of every w - 15 (10).
//
// What bounds the byte-seed scan (bsgpu_qindex.cuh) is the number of random 32-byte sectors it
// gathers from the L2-resident filter: one per probe, L / S probes.  S is limited by the seed
// length k (S = w - k + 1), and a 16-letter seed is what keeps random seed hits rare.  Here the
// two jobs of the seed are separated:
//   core  : the K = 10 letters at [p, p + K) every occurrence covering probe position p shares,
//           whatever its offset t = p - a in 0..S-1; it only ADDRESSES a 32-byte filter block
//           (4^10 cores over ~1 M blocks);
//   groups: an occurrence with t < S/2 also covers the E = 8 letters after the core, one with
//           t >= S/2 the E letters before it.  So every (pattern, t) is filed under an extended
//           seed of K + E = 18 letters -- group 0: [p, p + 18), group 1: [p - 8, p + 10) -- and
//           a probe tests BOTH groups against the one block it gathered: an entry sets one bit
//           in each of the block's 8 words (a blocked Bloom filter of 256-bit blocks, 16 bits
//           and 8 hash bits per entry: false positives ~0.1 % per test, against 0.9 % for 4 bits
//           in a 16-byte half -- every false positive costs two dependent DRAM misses in a
//           confirm warp, which is what the first version of this kernel was waiting for).
// One gather (LDG.256, one sector) per 16 letters; random hits of an 18-letter seed are
// negligible, so the candidates are the rare Bloom false positives and the real seeds.
// Candidates go through the same queue / entry table / byte-for-byte confirmation as the
// byte-seed scan (p == s, the fixed=TRUE table of src/match_pdict_utils.c:33-50), the entry
// table being keyed by the digest of the extended seed.
//
// Replaces walk_tb_subject (src/match_pdict_ACtree2.c:1001-1021) / walk_subject
// (src/match_pdict_Twobit.c:158-175) for Trusted Band = whole pattern, fixedS = TRUE.
#pragma once
#include "bsgpu_qindex.cuh"

#define BSGPU_XSEED_CORE 10
#define BSGPU_XSEED_EXT 8
#define BSGPU_XSEED_MIN_W 25            // S = w - 9 >= 2 * EXT

struct __align__(32) BsgpuFilterBlock {
	uint32_t v[8];
};

struct BsgpuXSeedView {
	const BsgpuFilterBlock *filter; // 32-byte blocks: an entry sets one bit in each of the 8 words
	uint32_t filter_shift;          // block = 20-bit packed core >> filter_shift
	int32_t kbits;                  // filter bits per extended seed (4 or 8, see xseed_set)
	const uint4 *table;             // four {overflow:1 | fp | e} entries per bucket, e = pattern * S + t
	uint32_t nbuckets;
	const uint8_t *tb;              // [M * w (+ 8)] pattern letters
	uint32_t id_bits;
	uint32_t fp_mask;
	int32_t s;                      // probe stride S
	int32_t w;                      // pattern width
};

// 4 letters (one word) -> 8 bits, letter i at bits 2i, two multiplies and a mask: w * 6 puts a
// distinct 2-bit code of a base letter at bits 3..4 of its byte (A 0, C 1, G 3, T 2: the bits of
// 6 = 00110b form a de Bruijn sequence), the second multiply gathers the four codes in the top
// byte (all partial products land on distinct bits).  Other bytes give some 8-bit value.
__host__ __device__ __forceinline__ uint32_t xseed_pack4_top(uint32_t w)
{
	return ((w * 6u) & 0x18181818u) * 0x00208208u;          // the codes are in bits 24..31
}

// A probe from the 32 bytes [p - 8, p + 24) as 8 little-endian words:
//   core  : bytes  8..17  (w2, w3, low half of w4), 2-bit packed: it IS the block address.  A hash
//           of the core would put Poisson(1) of the 4^10 cores in a block and the blocks' fill
//           (16 entries per core at 1M 25-mers) would vary from 0 to several times the mean, which
//           costs a Bloom filter an order of magnitude in false positives (measured 0.75 % against
//           0.14 % per test).
//   group0: bytes  8..25  (core + high half of w4, w5, low half of w6)
//   group1: bytes  0..17  (w0, w1 + core)
// The group digests are 32-bit multilinear sums over the raw letters (one IMAD per word); their
// well-mixed high bits are what the filter and the entry table use.
__host__ __device__ __forceinline__ void xseed_digests(const uint32_t (&w)[8], uint32_t &core,
						       uint32_t &x0, uint32_t &x1)
{
	uint32_t a = w[2] * 0x9E3779B1u + w[3] * 0x85EBCA77u + (w[4] & 0xFFFFu) * 0xC2B2AE3Du;

	x0 = a + (w[4] >> 16) * 0x27D4EB2Fu + w[5] * 0x165667B1u + (w[6] & 0xFFFFu) * 0xD3A2646Du;
	x1 = a + w[0] * 0xFD7046C5u + w[1] * 0xB55A4F09u;
	uint32_t p2 = xseed_pack4_top(w[2]), p3 = xseed_pack4_top(w[3]), p4 = xseed_pack4_top(w[4]);
#ifdef __CUDA_ARCH__
	core = __byte_perm(__byte_perm(p2, p3, 0x0073), p4, 0x0710) & 0x000FFFFFu;
#else
	core = (p2 >> 24) | ((p3 >> 24) << 8) | (((p4 >> 24) & 0xFu) << 16);
#endif
}

// block of a core: any 20 - shift of its bits (small dictionaries use fewer blocks)
__host__ __device__ __forceinline__ uint32_t xseed_block(uint32_t core, uint32_t shift)
{
	return core >> shift;
}

// what a candidate carries of its extended seed: 29 digest bits and the group
__host__ __device__ __forceinline__ uint32_t xseed_payload(uint32_t x, int group)
{
	return ((x >> 3) << 1) | (uint32_t) group;
}

// Filter bits of an entry / a test, KB per extended seed:
//   KB = 4: group g owns words 4g..4g+3 of the block, one bit in each (two 16-byte blocked Bloom
//           filters side by side, ~8 entries each): 10 instructions per test, ~0.4 % false positives
//   KB = 8: one bit in each of the block's 8 words, both groups mixed (~16 entries): twice the
//           instructions, ~0.14 % false positives
// The 5-bit positions come from bit 12 on of x * K (and x * K' for the second four).
#define BSGPU_XSEED_K1 0x85EBCA6Bu
#define BSGPU_XSEED_K2 0xC2B2AE35u

__host__ __forceinline__ void xseed_set(uint32_t *v, uint32_t x, int group, int kb)
{
	uint32_t ba = x * BSGPU_XSEED_K1, bb = x * BSGPU_XSEED_K2;

	for (int k = 0; k < 4; k++) {
		if (kb == 4) {
			v[4 * group + k] |= 1u << ((ba >> (12 + 5 * k)) & 31u);
		} else {
			v[k] |= 1u << ((ba >> (12 + 5 * k)) & 31u);
			v[4 + k] |= 1u << ((bb >> (12 + 5 * k)) & 31u);
		}
	}
}

template <int KB, int GROUP>
__device__ __forceinline__ uint32_t xseed_test(const uint32_t (&b)[8], uint32_t x)
{
	uint32_t ba = x * BSGPU_XSEED_K1;

	if (KB == 4) {
		constexpr int o = 4 * GROUP;
		return __funnelshift_r(b[o + 0], 0, ba >> 12) & __funnelshift_r(b[o + 1], 0, ba >> 17) &
		       __funnelshift_r(b[o + 2], 0, ba >> 22) & __funnelshift_r(b[o + 3], 0, ba >> 27) & 1u;
	}
	uint32_t bb = x * BSGPU_XSEED_K2;
	return __funnelshift_r(b[0], 0, ba >> 12) & __funnelshift_r(b[1], 0, ba >> 17) &
	       __funnelshift_r(b[2], 0, ba >> 22) & __funnelshift_r(b[3], 0, ba >> 27) &
	       __funnelshift_r(b[4], 0, bb >> 12) & __funnelshift_r(b[5], 0, bb >> 17) &
	       __funnelshift_r(b[6], 0, bb >> 22) & __funnelshift_r(b[7], 0, bb >> 27) & 1u;
}

__host__ __device__ __forceinline__ int xseed_group_of(int t, int s)
{
	return t < (s >> 1) ? 0 : 1;
}

// one 32-byte filter block, kept in L2 against the subject streaming through it
__device__ __forceinline__ void xseed_load_block(const BsgpuFilterBlock *p, uint32_t (&b)[8])
{
	asm volatile("ld.global.nc.L2::evict_last.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		     : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7])
		     : "l"(p));
}

// bulk copy whose lines are the first to leave L2 (the subject is read once)
__device__ __forceinline__ void tma_load_1d_stream(void *dst, const void *src, uint32_t bytes, uint64_t *bar,
						   uint64_t policy)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
		     :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
		     :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
	uint64_t p;

	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
	return p;
}

__device__ __forceinline__ void xseed_confirm(const BsgpuSubjectView &S, const BsgpuXSeedView &T,
		const BsgpuReport &R, unsigned long long cand, int64_t own_lo, int64_t own_hi)
{
	const int64_t p = (int64_t) (cand >> 30);
	const uint32_t payload = (uint32_t) cand & 0x3FFFFFFFu;
	const int group = (int) (payload & 1u);
	const uint32_t idmask = (1u << T.id_bits) - 1u;
	uint32_t bucket, fp;

	byteseed_table_slot(payload, T.nbuckets, bucket, fp);
	uint32_t fw = sampled_fp_of(fp, T.id_bits, T.fp_mask);

	for (;;) {
		uint4 bb = __ldg(T.table + bucket);
		uint32_t fits = 0;
		fits |= (((bb.x ^ fw) & T.fp_mask) == 0 && (bb.x & BSGPU_S_EMPTY) != BSGPU_S_EMPTY) ? 1u : 0u;
		fits |= (((bb.y ^ fw) & T.fp_mask) == 0 && (bb.y & BSGPU_S_EMPTY) != BSGPU_S_EMPTY) ? 2u : 0u;
		fits |= (((bb.z ^ fw) & T.fp_mask) == 0 && (bb.z & BSGPU_S_EMPTY) != BSGPU_S_EMPTY) ? 4u : 0u;
		fits |= (((bb.w ^ fw) & T.fp_mask) == 0 && (bb.w & BSGPU_S_EMPTY) != BSGPU_S_EMPTY) ? 8u : 0u;
		while (fits) {
			uint32_t ent = (fits & 1u) ? bb.x : ((fits & 2u) ? bb.y : ((fits & 4u) ? bb.z : bb.w));
			uint32_t e = ent & idmask;

			fits &= fits - 1;
			int id = (int) (e / (uint32_t) T.s);
			int t = (int) (e - (uint32_t) id * (uint32_t) T.s);
			int64_t a = p - t;
			int64_t end = a + T.w - 1;

			// an entry is confirmed by the group it was filed under only: the other
			// group's candidate of the same probe must not report it a second time
			if (xseed_group_of(t, T.s) != group || a < 0 || end < own_lo || end >= own_hi)
				continue;
			if (!byteseed_same_bytes(S.bytes + a, T.tb + (int64_t) id * T.w, T.w))
				continue;
			if (R.mode == BSGPU_R_COUNT) {
				count_add(R.counts + id, 1);
			} else {
				int seg = S.nseg == 1 ? 0 : find_segment(S, p);

				report_match(R, S, id, seg, end - __ldg(S.seg_start + seg) + 1, end, 0);
			}
		}
		if (!(bb.w & BSGPU_S_OVERFLOW))
			break;
		bucket = bucket + 1 == T.nbuckets ? 0 : bucket + 1;
	}
}

// out-of-line copy for the scan kernel (its parameters are __grid_constant__: their address is
// the constant bank, nothing is copied for the call)
__device__ __noinline__ void xseed_confirm_call(const BsgpuSubjectView *S, const BsgpuXSeedView *T,
		const BsgpuReport *R, unsigned long long cand, int64_t own_lo, int64_t own_hi)
{
	xseed_confirm(*S, *T, *R, cand, own_lo, own_hi);
}

#ifndef BSGPU_XSEED_WARPS
#define BSGPU_XSEED_WARPS 8             // all of them scan
#endif
#ifndef BSGPU_XSEED_STAGES
#define BSGPU_XSEED_STAGES 2            // 2 stages 0.0955 ms, 3 stages 0.0983, 6 stages 0.18: shared memory taken from L1 costs the gathers more than a deeper ring gains
#endif
#ifndef BSGPU_XSEED_MIN_BLOCKS
#define BSGPU_XSEED_MIN_BLOCKS 3        // resident blocks per SM the register budget is cut for
#endif

// bytes of one ring stage: the windows [p - 8, p + 24) of np * 32 probes S apart (+ alignment)
__host__ __device__ __forceinline__ uint32_t xseed_stage_bytes(int np, int s, bool aligned)
{
	if (aligned)            // S == 16, p - 8 on a 16-byte boundary: exactly (np * 32 + 1) 16-byte chunks
		return (uint32_t) ((np * 32 * 16 + 16 + 127) & ~127);
	return (uint32_t) ((np * 32 * s + 15 + 36 + 15 + 127) & ~127);
}

// candidates one warp may write: every probe of every tile it scans, both groups
__host__ __device__ __forceinline__ int64_t xseed_warp_capacity(int64_t nprobes, int np, int grid)
{
	int64_t tile = (int64_t) np * 32, ntiles = (nprobes + tile - 1) / tile;
	int64_t nwarps = (int64_t) grid * BSGPU_XSEED_WARPS;
	int64_t per_warp = (ntiles + nwarps - 1) / nwarps;

	return per_warp * tile * 2;
}

// Probe g of the launch sits at concat index first + g * S.  Every warp of a block runs its own
// ring of TMA bulk copies (tiles of NP * 32 probes, STAGES deep): wait for the tile, hash its
// probes, gather their filter blocks (one LDG.256 each), test, and append the probes the filter
// lets through (false positives and real seeds, ~1 % of the probes) to the warp's own slice of
// the candidate buffer, cand[gw * cap .. + n_cand[gw]) for global warp gw -- no atomics, the lanes
// take their places from a ballot.  confirm_xseed_kernel checks them, one per thread.
// Why two kernels: a candidate costs two or three dependent DRAM misses (entry bucket, pattern
// letters, subject letters).  Confirm warps inside the scan kernel (the byte-seed scan's design)
// take them 32 at a time, one round trip after the other: ~800 candidates per block were 80 us
// of latency for a scan that is done in 50 -- and the confirmation code cost the scan its
// registers.  As a kernel of their own all candidates are in flight at once (a few us).
// The gather latency is hidden by the other warps of the SM, not by software pipelining inside
// the warp: keeping the gathers of tile i - 1 in flight across the loop edge made ptxas put them
// on the scoreboard slot the next tile's mbarrier wait also uses, so every iteration waited for
// them before it could look at its letters (ncu: a fifth of all samples on that one instruction);
// cp.async gathers (LDGSTS: no registers, no scoreboard) are four times slower for fully
// divergent addresses.
// ALIGNED (S == 16 and first - 8 on a 16-byte boundary): a probe's 32 bytes are two aligned
// 16-byte shared-memory chunks, consecutive lanes read consecutive chunks (no bank conflicts,
// no realignment shifts).
// FUSED: a warp confirms its own candidates as soon as it holds 32 of them (and the rest when it
// runs out of tiles), one per lane, while the other warps of the SM keep scanning: the DRAM round
// trips of the confirmations hide under the scan and the second launch disappears.
template <int NP, bool ALIGNED, int KB = 4, int VAR = 0, bool FUSED = true, int STAGES = BSGPU_XSEED_STAGES,
	  int MINB = BSGPU_XSEED_MIN_BLOCKS>
__global__ void __launch_bounds__(BSGPU_XSEED_WARPS * 32, MINB)
scan_xseed_kernel(const __grid_constant__ BsgpuSubjectView S, const __grid_constant__ BsgpuXSeedView T,
		  const __grid_constant__ BsgpuReport R, int64_t first, int64_t nprobes,
		  unsigned long long *__restrict__ cand, int64_t cap, unsigned int *__restrict__ n_cand,
		  int64_t own_lo, int64_t own_hi)
{
	constexpr int NW = BSGPU_XSEED_WARPS;
	extern __shared__ __align__(128) uint8_t xseed_smem[];
	__shared__ __align__(8) uint64_t bars[NW][STAGES];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	const int64_t tile = (int64_t) NP * 32;
	const int64_t ntiles = (nprobes + tile - 1) / tile;
	const int64_t gw = (int64_t) blockIdx.x * NW + warp, nwarps = (int64_t) gridDim.x * NW;
	const int s = ALIGNED ? 16 : T.s;
	const uint32_t stage_bytes = xseed_stage_bytes(NP, s, ALIGNED);
	uint8_t *ring = xseed_smem + (size_t) warp * STAGES * stage_bytes;
	uint64_t *bar = bars[warp];
	const uint64_t policy = l2_policy_evict_first();
	unsigned long long *mine = cand + gw * cap;
	unsigned int n_mine = 0;                // candidates of this warp so far (the same in every lane)
	unsigned int n_done = 0;                // ... of which confirmed (FUSED)

	// tile tl -> the letters from 8 before its first probe on (16-byte aligned source)
	auto issue = [&](int64_t tl, int stage) {
		int64_t g_lo = tl * tile, b0 = first + g_lo * s - 8, base16 = b0 & ~(int64_t) 15;
		int np = (int) min(tile, nprobes - g_lo);
		uint32_t nbytes = ALIGNED ? (uint32_t) (np * 16 + 16)
					  : (uint32_t) (((int) (b0 - base16) + (np - 1) * s + 36 + 15) & ~15);

		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		if (VAR & 4)
			tma_load_1d(ring + (size_t) stage * stage_bytes, S.bytes + base16, nbytes, &bar[stage]);
		else
			tma_load_1d_stream(ring + (size_t) stage * stage_bytes, S.bytes + base16, nbytes, &bar[stage], policy);
	};

	if (lane == 0) {
		for (int k = 0; k < STAGES; k++)
			mbar_init(&bar[k], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		for (int k = 0; k < STAGES; k++)
			if (gw + k * nwarps < ntiles)
				issue(gw + k * nwarps, k);
	}
	__syncwarp();

	uint32_t phase = 0;
	int stage = 0;
	for (int64_t tl = gw; tl < ntiles; tl += nwarps) {
		int64_t g_lo = tl * tile, b0 = first + g_lo * s;
		int np = (int) min(tile, nprobes - g_lo);
		const uint8_t *sm = ring + (size_t) stage * stage_bytes;
		uint32_t xc[NP], x0[NP], x1[NP], b[NP][8];
		// probes of this lane in the tile: q * 32 + lane < np
		uint32_t live = (1u << min(NP, max(0, (np - lane + 31) >> 5))) - 1u;

		mbar_wait(&bar[stage], phase);
		if (ALIGNED) {
			const uint4 *ch = reinterpret_cast<const uint4 *>(sm) + lane;
			if (np == NP * 32) {            // a full tile: constant offsets, nothing to clamp
#pragma unroll
				for (int q = 0; q < NP; q++) {
					uint4 lo4 = ch[q * 32], hi4 = ch[q * 32 + 1];
					uint32_t wd[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};

					xseed_digests(wd, xc[q], x0[q], x1[q]);
				}
			} else {
#pragma unroll
				for (int q = 0; q < NP; q++) {
					// past the last probe of a short tile: rehash the last one (masked by live)
					int j = min(q * 32, np - 1 - lane);
					uint4 lo4 = ch[j], hi4 = ch[j + 1];
					uint32_t wd[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};

					xseed_digests(wd, xc[q], x0[q], x1[q]);
				}
			}
		} else {
			const int off = (int) ((b0 - 8) & 15), step = 32 * s;
			int o = off + lane * s, o_max = off + (np - 1) * s;
#pragma unroll
			for (int q = 0; q < NP; q++) {
				int oo = min(o, o_max);
				const uint32_t *a = reinterpret_cast<const uint32_t *>(sm + (oo & ~3));
				uint32_t sh = 8u * (uint32_t) (oo & 3);
				uint32_t raw[9], wd[8];
#pragma unroll
				for (int k = 0; k < 9; k++)
					raw[k] = a[k];
#pragma unroll
				for (int k = 0; k < 8; k++)
					wd[k] = __funnelshift_r(raw[k], raw[k + 1], sh);
				xseed_digests(wd, xc[q], x0[q], x1[q]);
				o += step;
			}
		}
		__syncwarp();           // every lane has read this stage: refill it
		if (lane == 0 && tl + STAGES * nwarps < ntiles)
			issue(tl + STAGES * nwarps, stage);
		if (++stage == STAGES) {
			stage = 0;
			phase ^= 1u;
		}
#pragma unroll
		for (int q = 0; q < NP; q++)
			xseed_load_block(T.filter + xseed_block(xc[q], T.filter_shift), b[q]);
		uint32_t rare0 = 0, rare1 = 0;
#pragma unroll
		for (int q = 0; q < NP; q++) {
			rare0 |= xseed_test<KB, 0>(b[q], x0[q]) << q;
			rare1 |= xseed_test<KB, 1>(b[q], x1[q]) << q;
		}
		rare0 &= live;
		rare1 &= live;
		if (VAR & 16)
			rare0 = rare1 = 0;
		// a round per candidate of the busiest lane (usually one): the lanes that still hold one
		// take consecutive places after the warp's count
		for (;;) {
			const bool have = (rare0 | rare1) != 0u;
			const unsigned int who = __ballot_sync(0xFFFFFFFFu, have);

			if (who == 0u)
				break;
			if (have) {
				int grp = rare0 ? 0 : 1;
				int q = __ffs(grp ? rare1 : rare0) - 1;
				uint32_t x = grp ? x1[0] : x0[0];
#pragma unroll
				for (int k = 1; k < NP; k++)
					x = q == k ? (grp ? x1[k] : x0[k]) : x;
				if (grp)
					rare1 &= rare1 - 1;
				else
					rare0 &= rare0 - 1;
				int64_t p = b0 + (int64_t) (q * 32 + lane) * s;
				mine[n_mine + __popc(who & ((1u << lane) - 1u))] =
					((unsigned long long) p << 30) | xseed_payload(x, grp);
			}
			n_mine += __popc(who);
		}
		if (FUSED && n_mine - n_done >= 32u) {
			__syncwarp();           // the lanes' writes to the slice
			do {
				xseed_confirm_call(&S, &T, &R, mine[n_done + lane], own_lo, own_hi);
				n_done += 32u;
			} while (n_mine - n_done >= 32u);
		}
	}
	if (FUSED) {
		__syncwarp();
		if ((unsigned int) lane < n_mine - n_done)
			xseed_confirm_call(&S, &T, &R, mine[n_done + lane], own_lo, own_hi);
		n_done = n_mine;
	}
	if (lane == 0)
		n_cand[gw] = n_mine - n_done;    // what confirm_xseed_kernel still has to do
}

// One candidate per thread.  blockIdx.y walks the scan blocks, whose 8 warps each left a slice:
// the block's threads take the candidates of the 8 slices as one list.
__global__ void __launch_bounds__(256)
confirm_xseed_kernel(BsgpuSubjectView S, BsgpuXSeedView T, BsgpuReport R,
		     const unsigned long long *__restrict__ cand, int64_t cap,
		     const unsigned int *__restrict__ n_cand, int64_t own_lo, int64_t own_hi)
{
	constexpr int NW = BSGPU_XSEED_WARPS;
	unsigned int upto[NW], n = 0;
#pragma unroll
	for (int k = 0; k < NW; k++) {
		n += __ldg(n_cand + (int64_t) blockIdx.y * NW + k);
		upto[k] = n;
	}
	for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		int k = 0;
		unsigned int before = 0;
#pragma unroll
		for (int j = 0; j < NW - 1; j++)
			if (i >= upto[j]) {
				k = j + 1;
				before = upto[j];
			}
		xseed_confirm(S, T, R, __ldcs(cand + ((int64_t) blockIdx.y * NW + k) * cap + (i - before)), own_lo, own_hi);
	}
}
