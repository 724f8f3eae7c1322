// bsgpu_qindex.cuh -- the two seed-index engines, used where the reference's answer does not
// depend on HOW candidates are found:
//
//   scan_byteseed_kernel  a TB_PDict whose Trusted Band is the whole pattern (exact matching,
//       w >= 17): 16-letter seeds of every pattern at offsets 0..S-1 (S = w - 15) are indexed
//       and the subject is probed only at every S-th position: an occurrence starting at a is
//       met exactly once, at the probe position p = a + ((-a) mod S).  The seeds are hashed from
//       the raw encoded bytes (TMA-staged tiles), filtered by a blocked Bloom filter and
//       confirmed byte for byte by dedicated warps -- no packed planes.
//   qscan2_kernel / qscan_kernel  an MTB_PDict (max.mismatch = k >= 1; qscan2_kernel for one
//       pattern width <= 29: rank directory + patterns inline in the slots, see below; qscan_kernel
//       for variable widths and up to 64 letters): the union over its k+1 Trusted Bands
//       (R/matchPDict.R:335-377, src/MIndex_class.c:230-263) is exactly the set of alignments
//       with min <= #mismatches <= max (pigeonhole), so any complete filter for Hamming
//       distance <= k gives the same result.  One gapped shape of weight q applied at a small
//       set of shifts (tools/find_seed_shapes.py) needs ONE directory probe per subject
//       position and |shifts|*M/4^q candidates, instead of k+1 probes and
//       sum_b M/4^w_b candidates for the reference's short bands.  Every candidate is
//       verified against the whole pattern on the packed planes (XOR + popcount), with
//       letters outside the segment and non-base subject letters counted as mismatches
//       (src/lowlevel_matching.c:79-86), so the result is exact.
#pragma once
#include "bsgpu_kernels.cuh"

#define BSGPU_Q_MAX_RUNS 8
#define BSGPU_Q_MAX_SHIFTS 32


struct BsgpuQIndexView {
	const uint32_t *bitmap;      // [2^key_bits / 32] bit = some entry has this directory key
	const uint2 *dir;            // [2^key_bits / 4] {base offset, 4 x u8 counts}
	const uint32_t *entries;     // fp | id << shift_bits | shift index (see q_ent_*), grouped by seed key
	const ulonglong2 *pat;       // [M] 2-bit packed pattern letters (<= 64)
	const unsigned long long *pat8;      // [M] the low 32 letters only (patterns of <= 32 letters)
	const uint8_t *pat_len;      // [M]
	int32_t q;                   // seed weight (letters)
	int32_t mode;                // 0 = exact, sampled; 1 = Hamming (gapped shape)
	int32_t stride;              // probe every stride-th position (mode 0), else 1
	int32_t min_w;               // window the shape lives in (mode 1)
	int32_t nruns;               // runs of consecutive shape offsets
	int32_t run_src[BSGPU_Q_MAX_RUNS];   // first offset of the run
	int32_t run_len[BSGPU_Q_MAX_RUNS];
	// the same runs as the key extraction wants them: key += ((win >> run_shift) & run_mask) * run_mul
	int32_t run_shift[BSGPU_Q_MAX_RUNS];
	uint32_t run_mask[BSGPU_Q_MAX_RUNS]; // 0 for unused runs
	uint32_t run_mul[BSGPU_Q_MAX_RUNS];  // 1 << (2 * letters of the runs before)
	uint64_t shape1;             // 1 bit per letter
	uint64_t shape2;             // the same, spread to the 2-bit layout (bit 2i)
	uint64_t mask2;              // both bits of every shape letter in the 2-bit layout
	int32_t key_bits;            // directory key = top key_bits of (window & mask2) * C
	int32_t const_len;           // > 0: every pattern has this many letters
	int32_t nshifts;
	int32_t shift[BSGPU_Q_MAX_SHIFTS];   // ascending
	// entry word: [fingerprint: 2 * nfp bits][pattern id: id_bits][shift index: shift_bits]
	int32_t shift_bits, id_bits;
	int32_t nfp;                 // 0..4 pattern letters outside the seed kept in the entry
	uint32_t fp_pos[BSGPU_Q_MAX_SHIFTS]; // per shift index: their pattern positions, one byte each
	// rank directory + inline patterns (qscan2_kernel; constant width <= 29, mode 1): see below
	int32_t v2;                  // 1: cells2 / entries2 are built
	const uint2 *cells2;         // [2^key_bits / 16] {16 x 2-bit slot counts, index of the cell's first slot}
	const unsigned long long *entries2;  // slots grouped by seed key, then the overflow chains
};

__host__ __device__ __forceinline__ int q_ent_shift(const BsgpuQIndexView &Q, uint32_t ent)
{
	return (int) (ent & ((1u << Q.shift_bits) - 1u));
}

__host__ __device__ __forceinline__ int q_ent_id(const BsgpuQIndexView &Q, uint32_t ent)
{
	return (int) ((ent >> Q.shift_bits) & ((1u << Q.id_bits) - 1u));
}

__host__ __device__ __forceinline__ uint32_t q_ent_fp(const BsgpuQIndexView &Q, uint32_t ent)
{
	return Q.nfp ? ent >> (Q.shift_bits + Q.id_bits) : 0u;
}

// Directory key of a window: the letters under the shape, compacted run by run into 2q bits.
// The key is injective on purpose: with a hash of the masked window (a random function onto
// 2^24 keys) popular keys collect both more dictionary entries and more subject windows, and
// the number of candidates per position doubles (measured 0.79 against 0.42 for 7 M entries).
__host__ __device__ __forceinline__ uint32_t q_key(const BsgpuQIndexView &Q, uint64_t win)
{
	uint32_t key = 0;
#pragma unroll
	for (int i = 0; i < BSGPU_Q_MAX_RUNS; i++) {
		if (i >= Q.nruns)
			break;
		key += ((uint32_t) (win >> Q.run_shift[i]) & Q.run_mask[i]) * Q.run_mul[i];    // disjoint fields
	}
	return key;
}

// 128-bit logical right shift by s (< 128) returning the low 64 bits
__device__ __forceinline__ uint64_t shr128_lo(uint64_t lo, uint64_t hi, int s)
{
	if (s == 0)
		return lo;
	if (s < 64)
		return (lo >> s) | (hi << (64 - s));
	return s == 64 ? hi : hi >> (s - 64);
}

// bits [2*from, 2*to) at even positions, as a 128-bit (lo, hi) mask
__device__ __forceinline__ void range_mask2(int from, int to, uint64_t &lo, uint64_t &hi)
{
	const uint64_t E = 0x5555555555555555ull;
	auto below = [&](int n, uint64_t &l, uint64_t &h) {      // letters [0, n)
		l = n >= 32 ? E : (n <= 0 ? 0ull : E & ((1ull << (2 * n)) - 1));
		h = n <= 32 ? 0ull : (n >= 64 ? E : E & ((1ull << (2 * (n - 32))) - 1));
	};
	uint64_t al, ah, bl, bh;
	below(to, al, ah);
	below(from, bl, bh);
	lo = al & ~bl;
	hi = ah & ~bh;
}

template <bool LONG>
__device__ __forceinline__ void q_verify(const BsgpuSubjectView &S, const BsgpuQIndexView &Q,
		const BsgpuReport &R, uint32_t ent, int64_t p, int max_nmis, int min_nmis,
		int64_t own_lo, int64_t own_hi)
{
	const uint64_t E = 0x5555555555555555ull;
	int id = q_ent_id(Q, ent), ti = q_ent_shift(Q, ent);
	int t = Q.shift[ti];
	int64_t a = p - t;
	int len = Q.const_len > 0 ? Q.const_len : (int) __ldg(Q.pat_len + id);
	int64_t own = Q.mode == 0 ? a + len - 1 : a + Q.min_w - 1;

	if (own < own_lo || own >= own_hi)
		return;
	int seg = S.nseg == 1 ? 0 : find_segment(S, p);
	int64_t lo = __ldg(S.seg_start + seg), L = __ldg(S.seg_len + seg);
	int64_t wi = a >> 5;
	int sh = 2 * (int) (a & 31);
	int nbefore = (int) min((int64_t) len, max((int64_t) 0, lo - a));
	int nafter = (int) min((int64_t) len, max((int64_t) 0, a + len - (lo + L)));
	uint64_t d_lo, d_hi = 0;

	if (!LONG) {
		// patterns of <= 32 letters: one 64-bit lane
		uint64_t c0 = __ldg(S.codes + wi), c1 = __ldg(S.codes + wi + 1);
		uint64_t i0 = __ldg(S.inv2 + wi), i1 = __ldg(S.inv2 + wi + 1);
		uint64_t x = sh ? (c0 >> sh) | (c1 << (64 - sh)) : c0;
		uint64_t iv = sh ? (i0 >> sh) | (i1 << (64 - sh)) : i0;
		uint64_t P = __ldg(Q.pat8 + id);

		d_lo = x ^ P;
		d_lo = ((d_lo | (d_lo >> 1)) | iv) & E;
		// letters outside the segment are mismatches
		if (nbefore)
			d_lo |= E & (nbefore >= 32 ? ~0ull : (1ull << (2 * nbefore)) - 1);
		if (nafter)
			d_lo |= E & ~((len - nafter) >= 32 ? ~0ull : (1ull << (2 * (len - nafter))) - 1);
		d_lo &= len >= 32 ? E : E & ((1ull << (2 * len)) - 1);
	} else {
		uint64_t c0 = __ldg(S.codes + wi), c1 = __ldg(S.codes + wi + 1), c2 = __ldg(S.codes + wi + 2);
		uint64_t i0 = __ldg(S.inv2 + wi), i1 = __ldg(S.inv2 + wi + 1), i2 = __ldg(S.inv2 + wi + 2);
		uint64_t x_lo = sh ? (c0 >> sh) | (c1 << (64 - sh)) : c0;
		uint64_t x_hi = sh ? (c1 >> sh) | (c2 << (64 - sh)) : c1;
		uint64_t v_lo = sh ? (i0 >> sh) | (i1 << (64 - sh)) : i0;
		uint64_t v_hi = sh ? (i1 >> sh) | (i2 << (64 - sh)) : i1;
		ulonglong2 P = Q.pat[id];
		uint64_t m_lo, m_hi;

		d_lo = x_lo ^ P.x;
		d_hi = x_hi ^ P.y;
		d_lo = ((d_lo | (d_lo >> 1)) | v_lo) & E;
		d_hi = ((d_hi | (d_hi >> 1)) | v_hi) & E;
		if (nbefore) {
			range_mask2(0, nbefore, m_lo, m_hi);
			d_lo |= m_lo;
			d_hi |= m_hi;
		}
		if (nafter) {
			range_mask2(len - nafter, len, m_lo, m_hi);
			d_lo |= m_lo;
			d_hi |= m_hi;
		}
		range_mask2(0, len, m_lo, m_hi);
		d_lo &= m_lo;
		d_hi &= m_hi;
	}
	int total = __popcll(d_lo) + __popcll(d_hi);
	if (total > max_nmis || total < min_nmis)
		return;
	if (Q.mode == 1) {
		// report from the first shift whose seed is clean, so that every alignment is
		// reported exactly once
		for (int k = 0; k < ti; k++)
			if ((shr128_lo(d_lo, d_hi, 2 * Q.shift[k]) & Q.shape2) == 0)
				return;
	}
	int64_t n = a - lo + len;               // 1-based local end of the match
	report_match(R, S, id, seg, n, a + len - 1, 0);
}

__device__ __forceinline__ uint32_t q_prefix(uint32_t cnts, int i)
{
	// sum of the count bytes below byte i
	uint32_t m = cnts & ((1u << (8 * i)) - 1u);     // i in 0..3
	m = (m & 0x00FF00FFu) + ((m >> 8) & 0x00FF00FFu);
	return (m & 0xFFFFu) + (m >> 16);
}

#define QSCAN_GROUP 8

// positions r (0..31) of a word whose seed letters (shape offsets) are all base letters;
// v = valid bits of 64 letters starting at the word
__device__ __forceinline__ uint32_t q_valid_positions(const BsgpuQIndexView &Q, uint64_t v)
{
	uint64_t m = ~0ull;

#pragma unroll
	for (int i = 0; i < BSGPU_Q_MAX_RUNS; i++)
		if (i < Q.nruns)
			m &= all_set_run(v, Q.run_len[i]) >> Q.run_src[i];
	return (uint32_t) m;
}

// Cheap lower bound of the mismatch count of candidate `ent` at probe position p: letters
// outside the segment are NOT forced to mismatch here, so the bound never exceeds the true
// count and a candidate it rejects is certainly not a match.  ~5 loads + ~30 ALU ops.
template <bool LONG>
__device__ __forceinline__ int q_lower_bound(const BsgpuSubjectView &S, const BsgpuQIndexView &Q,
					     uint32_t ent, int64_t p)
{
	const uint64_t E = 0x5555555555555555ull;
	int id = q_ent_id(Q, ent);
	int64_t a = p - Q.shift[q_ent_shift(Q, ent)];
	int64_t wi = a >> 5;
	int sh = 2 * (int) (a & 31);
	int len = Q.const_len > 0 ? Q.const_len : (int) __ldg(Q.pat_len + id);

	if (!LONG) {
		uint64_t c0 = __ldg(S.codes + wi), c1 = __ldg(S.codes + wi + 1);
		uint64_t i0 = __ldg(S.inv2 + wi), i1 = __ldg(S.inv2 + wi + 1);
		uint64_t x = sh ? (c0 >> sh) | (c1 << (64 - sh)) : c0;
		uint64_t iv = sh ? (i0 >> sh) | (i1 << (64 - sh)) : i0;
		uint64_t P = __ldg(Q.pat8 + id);
		uint64_t d = x ^ P;

		d = ((d | (d >> 1)) | iv) & (len >= 32 ? E : E & ((1ull << (2 * len)) - 1));
		return __popcll(d);
	} else {
		uint64_t c0 = __ldg(S.codes + wi), c1 = __ldg(S.codes + wi + 1), c2 = __ldg(S.codes + wi + 2);
		uint64_t i0 = __ldg(S.inv2 + wi), i1 = __ldg(S.inv2 + wi + 1), i2 = __ldg(S.inv2 + wi + 2);
		uint64_t x_lo = sh ? (c0 >> sh) | (c1 << (64 - sh)) : c0;
		uint64_t x_hi = sh ? (c1 >> sh) | (c2 << (64 - sh)) : c1;
		uint64_t v_lo = sh ? (i0 >> sh) | (i1 << (64 - sh)) : i0;
		uint64_t v_hi = sh ? (i1 >> sh) | (i2 << (64 - sh)) : i1;
		ulonglong2 P = Q.pat[id];
		uint64_t m_lo, m_hi, d_lo = x_lo ^ P.x, d_hi = x_hi ^ P.y;

		range_mask2(0, len, m_lo, m_hi);
		d_lo = ((d_lo | (d_lo >> 1)) | v_lo) & m_lo;
		d_hi = ((d_hi | (d_hi >> 1)) | v_hi) & m_hi;
		return __popcll(d_lo) + __popcll(d_hi);
	}
}

template <bool LONG>
__global__ void __launch_bounds__(256)
qscan_kernel(BsgpuSubjectView S, BsgpuQIndexView Q, BsgpuReport R, int max_nmis, int min_nmis,
	     int64_t probe_lo, int64_t probe_hi, int64_t own_lo, int64_t own_hi)
{
	// Warp-synchronous: all 32 lanes of a warp stay in every loop (lanes without work are
	// predicated off), so the candidate loop below reconverges every iteration.
	const unsigned FULL = 0xFFFFFFFFu;
	__shared__ uint64_t wc_all[8][34], wi_all[8][34];
	__shared__ uint16_t plist_all[8][1024];
	int64_t word_lo = probe_lo >> 5, word_hi = (probe_hi + 31) >> 5;
	int lane = threadIdx.x & 31;
	uint64_t *wc = wc_all[threadIdx.x >> 5], *wi = wi_all[threadIdx.x >> 5];
	uint16_t *plist = plist_all[threadIdx.x >> 5];
	int64_t jb = word_lo + (int64_t) blockIdx.x * blockDim.x + (threadIdx.x & ~31);
	int64_t stride = (int64_t) gridDim.x * blockDim.x;

	for (; jb < word_hi; jb += stride) {
		int64_t j = jb + lane;
		int64_t base = j << 5;
		uint64_t c0 = 0, c1 = 0;
		uint32_t m = 0;

		if (!LONG) {
			// the codes / not-a-base words jb-1 .. jb+32 of this batch go to the warp's shared
			// memory: candidate windows (start within 31 letters before the probe position,
			// <= 32 letters long) are cut from there by whichever lane takes the candidate
			bool ld = j < word_hi + 2;          // the planes have spare words after the end

			__syncwarp();
			c0 = ld ? __ldcs(S.codes + j) : 0ull;
			wc[lane + 1] = c0;
			wi[lane + 1] = ld ? __ldcs(S.inv2 + j) : ~0ull;
			if (lane == 0) {
				wc[0] = jb > 0 ? __ldcs(S.codes + jb - 1) : 0ull;
				wi[0] = jb > 0 ? __ldcs(S.inv2 + jb - 1) : ~0ull;
			}
			if (lane == 31) {
				bool ld2 = j + 1 < word_hi + 2;

				wc[33] = ld2 ? __ldcs(S.codes + j + 1) : 0ull;
				wi[33] = ld2 ? __ldcs(S.inv2 + j + 1) : ~0ull;
			}
		}
		if (j < word_hi) {
			if (LONG)
				c0 = __ldcs(S.codes + j);
			c1 = __ldcs(S.codes + j + 1);
			uint64_t v = (uint64_t) __ldcs(S.valid + j) | ((uint64_t) __ldcs(S.valid + j + 1) << 32);
			m = q_valid_positions(Q, v);
			if (Q.stride > 1) {
				uint32_t pm = 0;
				for (int r = (int) ((Q.stride - base % Q.stride) % Q.stride); r < 32; r += Q.stride)
					pm |= 1u << r;
				m &= pm;
			}
			if (base < probe_lo)
				m &= 0xFFFFFFFFu << (int) (probe_lo - base);
			if (base + 32 > probe_hi)
				m &= (probe_hi - base) <= 0 ? 0u : 0xFFFFFFFFu >> (int) (base + 32 - probe_hi);
		}
		// Phase 1: one bitmap word per position (2 MiB for 24 key bits); about a third of
		// the positions have a non-empty directory slot
		uint32_t have = 0;
#pragma unroll
		for (int g = 0; g < 32; g += QSCAN_GROUP) {
			uint32_t key[QSCAN_GROUP], bw[QSCAN_GROUP];

#pragma unroll
			for (int qq = 0; qq < QSCAN_GROUP; qq++) {
				const int r = g + qq;
				uint64_t win = r == 0 ? c0 : (c0 >> (2 * r)) | (c1 << (64 - 2 * r));

				key[qq] = q_key(Q, win);
			}
#pragma unroll
			for (int qq = 0; qq < QSCAN_GROUP; qq++)
				bw[qq] = ((m >> (g + qq)) & 1u) ? __ldg(Q.bitmap + (key[qq] >> 5)) : 0u;
#pragma unroll
			for (int qq = 0; qq < QSCAN_GROUP; qq++)
				have |= ((bw[qq] >> (key[qq] & 31u)) & 1u) << (g + qq);
		}
		// Phase 2, balanced and software-pipelined.  The non-empty positions of the whole batch
		// (32 words) are listed in shared memory and dealt round-robin to the lanes, so every
		// lane gets the same number of positions whatever their distribution over the words.
		// The three dependent gathers of a candidate (directory cell -> entry -> pattern)
		// belong to three different candidates within one iteration: an iteration finishes the
		// candidate whose pattern was requested last time, requests the pattern of the entry
		// that arrived last time, requests the next entry, and requests the directory cell of
		// the lane's next position when the current list is about to end -- one round trip
		// per iteration instead of three.
		int nitems;
		{
			int mine = __popc(have), before = mine;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				int up = __shfl_up_sync(FULL, before, d);
				if (lane >= d)
					before += up;
			}
			nitems = __shfl_sync(FULL, before, 31);
			before -= mine;
			for (uint32_t h = have; h != 0; h &= h - 1)
				plist[before++] = (uint16_t) (lane * 32 + __ffs(h) - 1);
			__syncwarp();
		}
		const int64_t base0 = jb << 5;
		int it = lane;
		uint32_t cur = 0, end = 0;
		int pos = 0;                    // position within the batch
		bool d_valid = false, e_valid = false, p_valid = false;
		uint2 d_cell = make_uint2(0, 0);
		uint32_t d_key = 0, e_ent = 0, p_ent = 0;
		int d_pos = 0, e_pos = 0, p_pos = 0, p_len = 0;
		uint64_t p_pat = 0, p_x = 0, p_iv = 0;
		for (;;) {
			if (!__any_sync(FULL, cur < end || it < nitems || d_valid || e_valid || p_valid))
				break;
			// 1. the pattern requested in the previous iteration has arrived
			if (p_valid) {
				const uint64_t E = 0x5555555555555555ull;
				uint64_t d = p_x ^ p_pat;

				d = ((d | (d >> 1)) | p_iv) & (p_len >= 32 ? E : E & ((1ull << (2 * p_len)) - 1));
				if (__popcll(d) <= max_nmis)
					q_verify<LONG>(S, Q, R, p_ent, base0 + p_pos, max_nmis, min_nmis, own_lo, own_hi);
				p_valid = false;
			}
			// 2. the entry requested in the previous iteration has arrived
			if (e_valid) {
				e_valid = false;
				if (!LONG) {
					int id = q_ent_id(Q, e_ent), ti = q_ent_shift(Q, e_ent);
					int at0 = e_pos - Q.shift[ti] + 32;     // window start, in letters from word jb-1
					int w = at0 >> 5, sh = 2 * (at0 & 31), lb = 0;
					uint64_t x = sh ? (wc[w] >> sh) | (wc[w + 1] << (64 - sh)) : wc[w];
					uint64_t iv = sh ? (wi[w] >> sh) | (wi[w + 1] << (64 - sh)) : wi[w];

					// the entry carries nfp pattern letters from outside the seed: most random
					// candidates already exceed max_nmis on them and never load the pattern
					if (Q.nfp) {
						uint32_t fp = q_ent_fp(Q, e_ent), where = Q.fp_pos[ti];
#pragma unroll
						for (int f = 0; f < 4; f++)
							if (f < Q.nfp) {
								int at = 2 * (int) ((where >> (8 * f)) & 0xFFu);

								lb += (int) ((((x >> at) ^ (fp >> (2 * f))) & 3u) != 0u ||
									     ((iv >> at) & 1u) != 0u);
							}
					}
					if (lb <= max_nmis) {
						p_pat = __ldg(Q.pat8 + id);
						p_len = Q.const_len > 0 ? Q.const_len : (int) __ldg(Q.pat_len + id);
						p_x = x;
						p_iv = iv;
						p_ent = e_ent;
						p_pos = e_pos;
						p_valid = true;
					}
				} else if (q_lower_bound<LONG>(S, Q, e_ent, base0 + e_pos) <= max_nmis) {
					q_verify<LONG>(S, Q, R, e_ent, base0 + e_pos, max_nmis, min_nmis, own_lo, own_hi);
				}
			}
			// 3. next entry of the current list, or start the list whose cell has arrived
			if (cur == end && d_valid) {
				cur = d_cell.x + q_prefix(d_cell.y, (int) (d_key & 3u));
				end = cur + ((d_cell.y >> (8 * (d_key & 3u))) & 0xFFu);
				pos = d_pos;
				d_valid = false;
			}
			if (cur < end) {
				e_ent = __ldg(Q.entries + cur++);
				e_pos = pos;
				e_valid = true;
			}
			// 4. the current list is about to end: request the cell of the lane's next position
			if (!d_valid && it < nitems && cur + 1 >= end) {
				uint64_t k0, k1;
				int r;

				d_pos = plist[it];
				it += 32;
				r = d_pos & 31;
				if (!LONG) {
					k0 = wc[(d_pos >> 5) + 1];
					k1 = wc[(d_pos >> 5) + 2];
				} else {
					k0 = __ldg(S.codes + jb + (d_pos >> 5));
					k1 = __ldg(S.codes + jb + (d_pos >> 5) + 1);
				}
				d_key = q_key(Q, r == 0 ? k0 : (k0 >> (2 * r)) | (k1 << (64 - 2 * r)));
				d_cell = __ldg(Q.dir + (d_key >> 2));
				d_valid = true;
			}
		}
	}
}

// ------------------------------------------------------------------------------------
// qscan2_kernel: the Hamming filter with ONE gather per subject position and ONE per candidate.
//
// qscan_kernel sits on the L1 gather rate (one tag lookup per divergent lane): a bitmap word per
// position, then directory cell -> entry -> pattern for each of the 0.42 candidates per position,
// ~2.1 random sectors per position.  Here
//   cells2   : 8 bytes per 16 seed keys = 16 two-bit slot counts + the index of the cell's first
//              slot.  The position's one gather answers "any entry?" AND "where": the slots of key
//              b start at  first + sum of the counts below b  (two popcounts).
//   entries2 : 8-byte slots grouped by key.  A slot holds the WHOLE pattern in place (2 bits per
//              letter, <= 29 letters) -- except that the 2q bits under the seed shape, which a
//              candidate shares with the subject window by construction (the key is injective),
//              carry the pattern id instead; bits 58..62 = the shift t the seed was cut at.
//              So a candidate is decided by one 8-byte load, an XOR with the subject window (cut
//              from the warp's shared-memory copy of the batch) under the complement of the
//              shifted shape, and a popcount; the pattern table is read for real matches only.
//              Keys with more than 3 entries (0.1 % at 7 M entries over 2^24 keys) keep two in
//              place and a pointer slot to a chain behind the grouped slots.
// Lane l of a warp owns the positions l, l + 32, ... of a batch of 1024: the slots a load group
// found are handed to the warp's ring in shared memory by ballot (no prefix pass), and the ring is
// drained 32 x QS2_U independent candidate loads at a time, so every lane has the same work
// whatever the distribution of the entries over the positions.
// ------------------------------------------------------------------------------------

#ifndef QS2_CAP
#define QS2_CAP 256             // ring items per warp (power of two)
#endif
#ifndef QS2_U
#define QS2_U 2                 // candidate loads in flight per lane
#endif
#ifndef QS2_WARPS
#define QS2_WARPS 2             // measured (tools/mm2_sweep.sh, 250 Mbp, ms): 2 warps x 16 blocks 1.39, 4 x 8 1.40, 8 x 4 1.45
#endif
#define QS2_PTR (1ull << 62)    // slot is a pointer: low 32 bits = index of the chain
#define QS2_MORE (1ull << 63)   // chain slot: the next slot belongs to the same key
#define QS2_T_SHIFT 58
#define QS2_MAX_LEN 29

__host__ __device__ __forceinline__ unsigned long long q2_deposit(const BsgpuQIndexView &Q, uint32_t id)
{
	unsigned long long out = 0;

	for (int i = 0; i < BSGPU_Q_MAX_RUNS; i++)
		if (i < Q.nruns)
			out |= (unsigned long long) ((id / Q.run_mul[i]) & Q.run_mask[i]) << Q.run_shift[i];
	return out;
}

__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
	uint64_t p;

	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
	return p;
}

// index loads that should stay in L2 against the planes streaming through it
__device__ __forceinline__ uint2 ldg_keep_u2(const uint2 *p, uint64_t policy)
{
	uint2 v;

	asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(policy));
	return v;
}

__device__ __forceinline__ unsigned long long ldg_keep_u64(const unsigned long long *p, uint64_t policy)
{
	unsigned long long v;

	asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(policy));
	return v;
}

// A slot that survived the XOR + popcount: its id sits under the shape.  Out of line (real
// matches only); the kernel's parameters are __grid_constant__, so their address is the constant
// bank and nothing is copied for the call.
__device__ __noinline__ void q2_survivor(const BsgpuSubjectView *S, const BsgpuQIndexView *Q, const BsgpuReport *R,
		unsigned long long e, int64_t p, int max_nmis, int min_nmis, int64_t own_lo, int64_t own_hi)
{
	const int t = (int) (e >> QS2_T_SHIFT) & 31;
	const uint32_t id = q_key(*Q, e >> (2 * t)) & ((1u << Q->id_bits) - 1u);
	int ti = 0;

	while (ti + 1 < Q->nshifts && Q->shift[ti] != t)
		ti++;
	q_verify<false>(*S, *Q, *R, (id << Q->shift_bits) | (uint32_t) ti, p, max_nmis, min_nmis, own_lo, own_hi);
}

// The seed key with the shape known at compile time: run i of the shape (letters [I, I + n)) goes
// to the key bits [at, at + 2n) -- one funnel shift and one AND-OR per run, all immediates.  The
// value equals q_key() of the same window (the host builds the index with that one).
__host__ __device__ constexpr int q2_run_len(unsigned long long shape, int i)
{
	int n = 0;

	while ((shape >> (i + n)) & 1ull)
		n++;
	return n;
}

__host__ __device__ constexpr int q2_letters_below(unsigned long long shape, int i)
{
	int n = 0;

	for (int k = 0; k < i; k++)
		n += (int) ((shape >> k) & 1ull);
	return n;
}

template <unsigned long long SHAPE, int I>
__device__ __forceinline__ uint32_t q2_key_runs(uint32_t lo, uint32_t hi)
{
	if constexpr (I >= 32 || (SHAPE >> I) == 0ull) {
		return 0u;
	} else if constexpr (((SHAPE >> I) & 1ull) != 0ull) {
		constexpr int n = q2_run_len(SHAPE, I), at = 2 * q2_letters_below(SHAPE, I), d = 2 * I - at;
		constexpr uint32_t M = (uint32_t) (((1ull << (2 * n)) - 1ull) << at);
		const uint32_t piece = d == 0 ? lo : (d < 32 ? __funnelshift_r(lo, hi, d) : hi >> (d - 32));

		return (piece & M) | q2_key_runs<SHAPE, I + n>(lo, hi);
	} else {
		return q2_key_runs<SHAPE, I + 1>(lo, hi);
	}
}

template <unsigned long long SHAPE>
__device__ __forceinline__ uint32_t q2_key(const BsgpuQIndexView &Q, uint32_t lo, uint32_t hi)
{
	if constexpr (SHAPE == 0ull)
		return q_key(Q, (uint64_t) lo | ((uint64_t) hi << 32));
	else
		return q2_key_runs<SHAPE, 0>(lo, hi);
}

// 64 bits from bit offset sh (0..63) of the 128-bit little-endian value {c0, c1}
__device__ __forceinline__ void q2_cut(uint64_t c0, uint64_t c1, int sh, uint32_t &lo, uint32_t &hi)
{
	const bool up = sh >= 32;
	const uint32_t a0 = (uint32_t) c0, a1 = (uint32_t) (c0 >> 32), b0 = (uint32_t) c1, b1 = (uint32_t) (c1 >> 32);
	const uint32_t w0 = up ? a1 : a0, w1 = up ? b0 : a1, w2 = up ? b1 : b0;

	lo = __funnelshift_r(w0, w1, sh);       // the shift amount wraps at 32
	hi = __funnelshift_r(w1, w2, sh);
}

#ifndef QS2_MIN_BLOCKS
#define QS2_MIN_BLOCKS 16
#endif
#ifndef QS2_GROUP
#define QS2_GROUP 2             // positions per lane and load group: 2 x 96 slots + a round's rest < QS2_CAP
                                // (group 2 / ring 256 / rounds of 64: 1.40 ms; 4 / 512 / 128: 1.66; 8 / 1024 / 128: 3.1)
#endif
static_assert(QS2_GROUP * 96 + 32 * QS2_U <= QS2_CAP && (QS2_CAP & (QS2_CAP - 1)) == 0 && 32 % QS2_GROUP == 0,
	      "ring too small for a load group on top of an unfinished round");

template <unsigned long long SHAPE>
__global__ void __launch_bounds__(QS2_WARPS * 32, QS2_MIN_BLOCKS)
qscan2_kernel(const __grid_constant__ BsgpuSubjectView S, const __grid_constant__ BsgpuQIndexView Q,
	      const __grid_constant__ BsgpuReport R, int max_nmis, int min_nmis,
	      int64_t probe_lo, int64_t probe_hi, int64_t own_lo, int64_t own_hi)
{
	const unsigned FULL = 0xFFFFFFFFu;
	const uint64_t E = 0x5555555555555555ull;
	__shared__ uint64_t wc_all[QS2_WARPS][34], wi_all[QS2_WARPS][34];
	__shared__ uint32_t mw_all[QS2_WARPS][32];
	__shared__ uint2 ring_all[QS2_WARPS][QS2_CAP];  // {slot index, position in the batch}
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint64_t *wc = wc_all[warp], *wi = wi_all[warp];
	uint32_t *mw = mw_all[warp];
	uint2 *ring = ring_all[warp];
	const int64_t word_lo = probe_lo >> 5, word_hi = (probe_hi + 31) >> 5;
	const int64_t stride = (int64_t) gridDim.x * QS2_WARPS * 32;
	const int len = Q.const_len;
	const uint64_t len2 = (1ull << (2 * len)) - 1, lenE = E & len2;
	const uint64_t mask2 = Q.mask2;
	const uint32_t lt = (1u << lane) - 1u;
	const uint64_t keep = l2_policy_evict_last();
	uint32_t head = 0, tail = 0;            // ring cursors (the same in every lane)

	for (int64_t jb = word_lo + ((int64_t) blockIdx.x * QS2_WARPS + warp) * 32; jb < word_hi; jb += stride) {
		const int64_t j = jb + lane, base = j << 5, base0 = jb << 5;
		const bool ld = j < word_hi + 2;        // the planes have spare words after the end
		uint32_t m = 0;
		uint64_t inv = ld ? __ldcs(S.inv2 + j) : ~0ull;

		__syncwarp();
		wc[lane + 1] = ld ? __ldcs(S.codes + j) : 0ull;
		wi[lane + 1] = inv;
		if (lane == 0) {
			uint64_t i0 = jb > 0 ? __ldcs(S.inv2 + jb - 1) : ~0ull;

			wc[0] = jb > 0 ? __ldcs(S.codes + jb - 1) : 0ull;
			wi[0] = i0;
			inv |= i0;
		}
		if (lane == 31) {
			const bool ld2 = j + 1 < word_hi + 2;
			uint64_t i1 = ld2 ? __ldcs(S.inv2 + j + 1) : ~0ull;

			wc[33] = ld2 ? __ldcs(S.codes + j + 1) : 0ull;
			wi[33] = i1;
			inv |= i1;
		}
		// no N (or separator) anywhere near the batch: the candidates skip the not-a-base plane
		const bool any_inv = __any_sync(FULL, inv != 0ull);
		if (j < word_hi) {
			uint64_t v = (uint64_t) __ldcs(S.valid + j) | ((uint64_t) __ldcs(S.valid + j + 1) << 32);

			m = q_valid_positions(Q, v);
			if (base < probe_lo)
				m &= 0xFFFFFFFFu << (int) (probe_lo - base);
			if (base + 32 > probe_hi)
				m &= (probe_hi - base) <= 0 ? 0u : 0xFFFFFFFFu >> (int) (base + 32 - probe_hi);
		}
		// every position of the batch is a probe (the usual case): the load groups skip the masks
		const bool all_on = __all_sync(FULL, m == 0xFFFFFFFFu);
		mw[lane] = m;
		__syncwarp();
#pragma unroll 1
		for (int g0 = 0; g0 <= 32; g0 += QS2_GROUP) {
			// Drain: 32 x QS2_U slots per round, one load + XOR + popcount each.  Between the load
			// groups only full rounds run (what is left stays in the ring).
			while (tail - head >= (g0 == 32 ? 1u : 32u * QS2_U)) {
				const uint32_t nround = min(tail - head, 32u * QS2_U);
				uint2 item[QS2_U];
				unsigned long long ent[QS2_U];
#pragma unroll
				for (int u = 0; u < QS2_U; u++) {
					const uint32_t k = (uint32_t) (u * 32 + lane);

					item[u] = ring[(head + k) & (QS2_CAP - 1)];
					ent[u] = k < nround ? ldg_keep_u64(Q.entries2 + item[u].x, keep) : QS2_PTR;
				}
				head += nround;
				__syncwarp();           // the ring items just read may be overwritten below
#pragma unroll
				for (int u = 0; u < QS2_U; u++) {
					const unsigned long long e = ent[u];
					const int pos = (int) item[u].y;

					if ((e & QS2_PTR) == 0ull) {
						const int t = (int) (e >> QS2_T_SHIFT) & 31;
						const int at0 = pos - t + 32;    // window start, in letters from word jb-1
						const int w = at0 >> 5, sh = 2 * (at0 & 31);
						uint32_t x_lo, x_hi, i_lo = 0, i_hi = 0;

						q2_cut(wc[w], wc[w + 1], sh, x_lo, x_hi);
						if (any_inv)
							q2_cut(wi[w], wi[w + 1], sh, i_lo, i_hi);
						uint64_t d = (((uint64_t) x_lo | ((uint64_t) x_hi << 32)) ^ e) & len2 & ~(mask2 << (2 * t));

						d = (d | (d >> 1) | ((uint64_t) i_lo | ((uint64_t) i_hi << 32))) & lenE;
						if (__popcll(d) <= max_nmis)
							q2_survivor(&S, &Q, &R, e, base0 + pos, max_nmis, min_nmis, own_lo, own_hi);
					}
					// the rare key with more than 3 entries: the chain behind a pointer slot, or the
					// next chain slot (an idle lane's QS2_PTR is masked by the round's length)
					const bool need = (e & (QS2_PTR | QS2_MORE)) != 0ull && (uint32_t) (u * 32 + lane) < nround;
					const uint32_t who = __ballot_sync(FULL, need);

					if (who != 0u) {
						if (need)
							ring[(tail + __popc(who & lt)) & (QS2_CAP - 1)] =
								make_uint2((e & QS2_PTR) ? (uint32_t) e : item[u].x + 1u, (uint32_t) pos);
						tail += __popc(who);
					}
				}
				__syncwarp();
			}
			if (g0 == 32)
				break;
			// Load group: the positions (g0 + qq) * 32 + lane
			uint32_t key[QS2_GROUP];
			uint2 cell[QS2_GROUP];
			uint64_t c0 = wc[g0 + 1];
#pragma unroll
			for (int qq = 0; qq < QS2_GROUP; qq++) {
				const uint64_t c1 = wc[g0 + qq + 2];
				uint32_t w_lo, w_hi;

				q2_cut(c0, c1, 2 * lane, w_lo, w_hi);
				key[qq] = q2_key<SHAPE>(Q, w_lo, w_hi);
				c0 = c1;
			}
			if (all_on) {
#pragma unroll
				for (int qq = 0; qq < QS2_GROUP; qq++)
					cell[qq] = ldg_keep_u2(Q.cells2 + (key[qq] >> 4), keep);
			} else {
#pragma unroll
				for (int qq = 0; qq < QS2_GROUP; qq++)
					cell[qq] = ((mw[g0 + qq] >> lane) & 1u) ? ldg_keep_u2(Q.cells2 + (key[qq] >> 4), keep)
										: make_uint2(0u, 0u);
			}
#pragma unroll
			for (int qq = 0; qq < QS2_GROUP; qq++) {
				const int b2 = 2 * (int) (key[qq] & 15u);
				const uint32_t c = (cell[qq].x >> b2) & 3u;
				const uint32_t b_lo = __ballot_sync(FULL, (c & 1u) != 0u), b_hi = __ballot_sync(FULL, (c & 2u) != 0u);

				if (c != 0u) {
					const uint32_t below = cell[qq].x & ((1u << b2) - 1u);
					const uint32_t first = cell[qq].y + __popc(below & 0x55555555u) + 2 * __popc(below & 0xAAAAAAAAu);
					const uint32_t at = tail + __popc(b_lo & lt) + 2 * __popc(b_hi & lt);
					const uint32_t p = (uint32_t) ((g0 + qq) * 32 + lane);

					ring[at & (QS2_CAP - 1)] = make_uint2(first, p);
					if (c > 1u)
						ring[(at + 1u) & (QS2_CAP - 1)] = make_uint2(first + 1u, p);
					if (c > 2u)
						ring[(at + 2u) & (QS2_CAP - 1)] = make_uint2(first + 2u, p);
				}
				tail += __popc(b_lo) + 2 * __popc(b_hi);
			}
			__syncwarp();
		}
	}
}

// ------------------------------------------------------------------------------------
// Entry-table helpers of the sampled-seed index (exact dictionaries).
// Table: 16-byte buckets of four 4-byte entries {overflow:1 | fp | e}, e = pattern*S + t in
// the low id_bits, fp = hash bits above them; 0x7FFFFFFF = empty.  The overflow bit (on
// slot 3) says that an entry hashed here was placed further on.  Different patterns may
// share a seed (overlapping probes), so every entry carrying the fingerprint is verified.
// ------------------------------------------------------------------------------------

#define BSGPU_S_EMPTY 0x7FFFFFFFu       // all fp and id bits set, overflow bit clear
#define BSGPU_S_OVERFLOW 0x80000000u

// entry word without the id: fingerprint bits in place
__host__ __device__ __forceinline__ uint32_t sampled_fp_of(uint32_t fp, uint32_t id_bits, uint32_t fp_mask)
{
	return (fp << id_bits) & fp_mask;
}

// ------------------------------------------------------------------------------------
// Exact dictionaries straight from the subject's letters (no packed planes).
// Sampled seeds of 16 letters: every pattern of width w >= 17 is indexed under its
// S = min(w - 15, 32) seeds at offsets 0..S-1 and the subject is probed at the positions
// p = 0 (mod S) only: an occurrence starting at a meets exactly one of them
// (t = (-a) mod S), so the number of index probes -- L2 random-sector gathers -- drops by S.  The seed is hashed as the four raw 32-bit
// words of its 16 encoded bytes (multilinear hash, 4 IMAD.WIDE), so a probe costs 5 shared
// memory word loads, 4 funnel shifts, the hash and one gather of a 16-byte Bloom block -- no
// 2-bit conversion and no validity bookkeeping: a window holding a non-base letter, a
// separator or a guard byte simply hashes to something no dictionary seed has (and the exact
// checks below reject it if the filter lets it through).  HBM traffic per letter: 1 byte.
//   filter : blocked Bloom filter, one 16-byte block per ~8 entries, one bit in each of the
//            4 words (false positives ~0.6 %); it stays in L2.
//   table  : 16-byte buckets of four {overflow:1 | fp | e} entries, e = pattern * S + offset;
//            a filter hit walks it and compares the whole pattern of every entry whose
//            fingerprint fits (p == s, the fixed=TRUE table of
//            src/match_pdict_utils.c:33-50).
// ------------------------------------------------------------------------------------

struct BsgpuByteSeedView {
	const uint4 *filter;         // Bloom blocks
	uint32_t nfilter;
	const uint4 *table;          // four {overflow:1 | fp | e} entries per bucket
	uint32_t nbuckets;
	const uint8_t *tb;           // [M * w (+ 8)] pattern letters
	uint32_t id_bits;
	uint32_t fp_mask;
	int32_t s;                   // sampling stride S
	int32_t w;                   // pattern width
};

#define BSGPU_BYTESEED_LETTERS 16

// 32-bit digest of the 16 seed bytes.  The 64-bit sum of the four 32 x 32 -> 64 products is
// (nearly) injective on the 2^32 possible base-letter seeds; folding its halves keeps the
// well-mixed low word and the smooth high word.
__host__ __device__ __forceinline__ uint32_t byteseed_digest(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
{
	uint64_t h = (uint64_t) w0 * 0x9E3779B1u + (uint64_t) w1 * 0x85EBCA77u +
		     (uint64_t) w2 * 0xC2B2AE3Du + (uint64_t) w3 * 0x27D4EB2Fu;

	return (uint32_t) h ^ (uint32_t) (h >> 32);
}

// filter block of a digest, and the word holding its 4 bit positions (5 bits each from bit 12)
__host__ __device__ __forceinline__ void byteseed_filter_slot(uint32_t x, uint32_t nfilter,
							       uint32_t &block, uint32_t &bits)
{
	block = (uint32_t) (((uint64_t) (x * 0x9E3779B1u) * nfilter) >> 32);
	bits = x * 0x85EBCA6Bu;
}

__host__ __device__ __forceinline__ uint32_t byteseed_bit(uint32_t bits, int k)
{
	return (bits >> (12 + 5 * k)) & 31u;
}

// entry table: x = the top 30 bits of the digest (what a candidate record carries); bucket
// from the top of the fully mixed value, fingerprint from its low bits
__host__ __device__ __forceinline__ void byteseed_table_slot(uint32_t x, uint32_t nbuckets,
							      uint32_t &bucket, uint32_t &fp)
{
	x ^= x >> 16;
	x *= 0x7FEB352Du;
	x ^= x >> 15;
	x *= 0x846CA68Bu;
	x ^= x >> 16;
	bucket = (uint32_t) (((uint64_t) x * nbuckets) >> 32);
	fp = x;
}

// the 16 letters at byte offset p of 'bytes' (global or shared) as four little-endian words
__device__ __forceinline__ void byteseed_window(const uint8_t *bytes, int64_t p,
		uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3)
{
	const uint32_t *a = reinterpret_cast<const uint32_t *>(bytes + (p & ~(int64_t) 3));
	uint32_t sh = 8u * (uint32_t) (p & 3);
	uint32_t x0 = a[0], x1 = a[1], x2 = a[2], x3 = a[3], x4 = a[4];

	w0 = __funnelshift_r(x0, x1, sh);
	w1 = __funnelshift_r(x1, x2, sh);
	w2 = __funnelshift_r(x2, x3, sh);
	w3 = __funnelshift_r(x3, x4, sh);
}

// Are the n bytes at x and y the same?  Word loads from the 4-byte aligned addresses below
// x and y, realigned with funnel shifts; both buffers are readable 7 bytes past the end.
// The first 32 bytes are loaded in one batch (18 independent loads).
__device__ __forceinline__ bool byteseed_same_bytes(const uint8_t *__restrict__ x,
		const uint8_t *__restrict__ y, int n)
{
	const uint32_t *xw = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(x) & ~(uintptr_t) 3);
	const uint32_t *yw = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(y) & ~(uintptr_t) 3);
	uint32_t xs = 8u * (uint32_t) (reinterpret_cast<uintptr_t>(x) & 3);
	uint32_t ys = 8u * (uint32_t) (reinterpret_cast<uintptr_t>(y) & 3);
	const int nw = (n + 3) >> 2;            // words to compare; the last one may be partial
	uint32_t xv[9], yv[9], diff = 0;
#pragma unroll
	for (int k = 0; k < 9; k++) {
		bool need = k <= nw;
		xv[k] = need ? __ldg(xw + k) : 0u;
		yv[k] = need ? __ldg(yw + k) : 0u;
	}
#pragma unroll
	for (int k = 0; k < 8; k++) {
		uint32_t d = __funnelshift_r(xv[k], xv[k + 1], xs) ^ __funnelshift_r(yv[k], yv[k + 1], ys);
		int left = n - 4 * k;

		diff |= left >= 4 ? d : (left > 0 ? d & ((1u << (8 * left)) - 1u) : 0u);
	}
	if (diff != 0 || n <= 32)
		return diff == 0;
	uint32_t x0 = xv[8], y0 = yv[8];
	for (int k = 8; 4 * k < n; k++) {
		uint32_t x1 = __ldg(xw + k + 1), y1 = __ldg(yw + k + 1);
		uint32_t d = __funnelshift_r(x0, x1, xs) ^ __funnelshift_r(y0, y1, ys);
		int left = n - 4 * k;

		diff |= left >= 4 ? d : d & ((1u << (8 * left)) - 1u);
		x0 = x1;
		y0 = y1;
	}
	return diff == 0;
}

// A probe the filter let through, as queued by the scan: concat index in the high 34 bits,
// the top 30 bits of the seed digest below.
__device__ __forceinline__ unsigned long long byteseed_candidate(int64_t p, uint32_t x)
{
	return ((unsigned long long) p << 30) | (x >> 2);
}

// Confirms a candidate: entry table (fingerprint), then the whole pattern
// byte for byte (p == s, fixed=TRUE).
__device__ __forceinline__ void byteseed_confirm(const BsgpuSubjectView &S, const BsgpuByteSeedView &T,
		const BsgpuReport &R, unsigned long long cand, int64_t own_lo, int64_t own_hi)
{
	const int64_t p = (int64_t) (cand >> 30);
	const uint32_t idmask = (1u << T.id_bits) - 1u;
	uint32_t bucket, fp;

	byteseed_table_slot((uint32_t) cand & 0x3FFFFFFFu, T.nbuckets, bucket, fp);
	uint32_t fw = sampled_fp_of(fp, T.id_bits, T.fp_mask);

	for (;;) {
		uint4 bb = __ldg(T.table + bucket);
		// slots whose fingerprint fits (usually none or one)
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
			int64_t a = p - (int64_t) (e - (uint32_t) id * (uint32_t) T.s);
			int64_t end = a + T.w - 1;

			if (a < 0 || end < own_lo || end >= own_hi)
				continue;
			// pattern letters are base letters, so equal bytes <=> p == s on valid letters
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

// out-of-line copy for the scan kernel's "candidate slice full" path
__device__ __noinline__ void byteseed_confirm_now(BsgpuSubjectView S, BsgpuByteSeedView T, BsgpuReport R,
		unsigned long long cand, int64_t own_lo, int64_t own_hi)
{
	byteseed_confirm(S, T, R, cand, own_lo, own_hi);
}

// --- TMA bulk copy + mbarrier (sm_90+ PTX) ---
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile("{\n"
		     ".reg .pred P1;\n"
		     "BSGPU_WAIT:\n"
		     "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
		     "@P1 bra BSGPU_DONE;\n"
		     "bra BSGPU_WAIT;\n"
		     "BSGPU_DONE:\n"
		     "}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// global -> shared bulk copy of 'bytes' (multiple of 16, both sides 16-byte aligned) that
// completes on 'bar'
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
		     :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		     :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Probe g of the launch sits at concat index first + g * S.  A block is SCAN_WARPS scan warps
// and 8 - SCAN_WARPS confirm warps.
// Scan warps: each runs its own pipeline over tiles of NP * 32 consecutive probes: the
// letters under a tile (NP * 32 * S + 20 bytes) are brought into the warp's shared-memory
// ring by one TMA bulk copy, STAGES deep; lane l hashes the NP windows l, l + 32, ... of
// tile i while its NP filter gathers of tile i - 1 are in flight, so the HBM stream, the
// hashing and the L2 gathers overlap and no block-wide barrier sits in the loop.
// Probes the filter lets through (false positives, seed hits, matches: ~1 %) are pushed
// into the block's candidate queue in shared memory.
// Confirm warp: pops 32 candidates at a time and confirms them, one per lane, so that the
// dependent table / pattern loads of the confirmations (two DRAM misses each) overlap the
// scan instead of following it.  Producers never wait: when the queue is nearly full they
// confirm their candidate themselves.
// Dynamic shared memory: 7 warps * STAGES * byteseed_stage_bytes(NP, S).
#define BSGPU_BYTESEED_STAGES 3
#ifndef BSGPU_BYTESEED_SCAN_WARPS
#define BSGPU_BYTESEED_SCAN_WARPS 6
#endif
#define BSGPU_BYTESEED_QUEUE 1024

__host__ __device__ __forceinline__ uint32_t byteseed_stage_bytes(int np, int s)
{
	return (uint32_t) ((np * 32 * s + 16 + 20 + 15 + 127) & ~127);
}

template <int NP>
__global__ void __launch_bounds__(256, NP == 8 ? 2 : 4)
scan_byteseed_kernel(BsgpuSubjectView S, BsgpuByteSeedView T, BsgpuReport R, int64_t first,
		     int64_t nprobes, int64_t own_lo, int64_t own_hi)
{
	constexpr int STAGES = BSGPU_BYTESEED_STAGES, NSCAN = BSGPU_BYTESEED_SCAN_WARPS;
	constexpr unsigned int QCAP = BSGPU_BYTESEED_QUEUE;
	extern __shared__ __align__(128) uint8_t byteseed_smem[];
	__shared__ __align__(8) uint64_t bars[NSCAN][STAGES];
	__shared__ unsigned long long queue[QCAP];      // 0 = slot not written yet
	constexpr int NCONF = 8 - NSCAN;
	// q_heads[j]: first slot of the next batch confirm warp j will take (everything below the
	// smallest of them has been consumed); q_tail: slots reserved; q_done: scan warps finished
	__shared__ unsigned int q_heads[NCONF], q_tail, q_done;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	for (unsigned int i = threadIdx.x; i < QCAP; i += blockDim.x)
		queue[i] = 0ull;
	if (threadIdx.x == 0) {
		for (int j = 0; j < NCONF; j++)
			q_heads[j] = 32u * j;
		q_tail = 0;
		q_done = 0;
	}
	__syncthreads();                // the only block-wide barrier

	if (warp >= NSCAN) {
		// ---- confirm warps: warp j takes the batches j, j + NCONF, ... of 32 queue slots ----
		const int j = warp - NSCAN;
		for (unsigned int k = j;; k += NCONF) {
			const unsigned int lo = 32u * k;
			unsigned int n;

			for (;;) {
				unsigned int done = *(volatile unsigned int *) &q_done;
				unsigned int tail = *(volatile unsigned int *) &q_tail;

				if (tail - lo >= 32u && tail >= lo) {
					n = 32u;
					break;
				}
				if (done == NSCAN) {
					n = tail > lo ? tail - lo : 0u;
					break;
				}
				__nanosleep(256);
			}
			if (n == 0u)
				break;
			unsigned long long c = 0ull;
			if ((unsigned int) lane < n) {
				volatile unsigned long long *slot = &queue[(lo + lane) % QCAP];
				while ((c = *slot) == 0ull)
					;               // reserved, about to be written
				*slot = 0ull;
			}
			__syncwarp();
			if (lane == 0)
				*(volatile unsigned int *) &q_heads[j] = lo + 32u * NCONF;      // my next batch
			if ((unsigned int) lane < n)
				byteseed_confirm(S, T, R, c, own_lo, own_hi);
			__syncwarp();
			if (n < 32u)
				break;
		}
		return;
	}

	// ---- scan warps ----
	const int64_t tile = (int64_t) NP * 32;
	const int64_t ntiles = (nprobes + tile - 1) / tile;
	const int64_t gw = (int64_t) blockIdx.x * NSCAN + warp, nwarps = (int64_t) gridDim.x * NSCAN;
	const uint32_t stage_bytes = byteseed_stage_bytes(NP, T.s);
	uint8_t *ring = byteseed_smem + (size_t) warp * STAGES * stage_bytes;
	uint64_t *bar = bars[warp];

	// tile tl -> (first byte, 16-byte aligned source, bytes to copy)
	auto issue = [&](int64_t tl, int stage) {
		int64_t g_lo = tl * tile, b0 = first + g_lo * T.s, base16 = b0 & ~(int64_t) 15;
		int np = (int) min(tile, nprobes - g_lo);
		uint32_t nbytes = (uint32_t) (((int) (b0 - base16) + (np - 1) * T.s + 20 + 15) & ~15);

		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		tma_load_1d(ring + (size_t) stage * stage_bytes, S.bytes + base16, nbytes, &bar[stage]);
	};
	// filter verdicts of a tile -> candidate queue
	auto emit = [&](const uint4 (&b)[NP], const uint32_t (&xs)[NP], uint32_t live, int64_t b0) {
		uint32_t rare = 0;
#pragma unroll
		for (int q = 0; q < NP; q++) {
			uint32_t bits = xs[q] * 0x85EBCA6Bu;
			uint32_t hit = __funnelshift_r(b[q].x, 0, bits >> 12) & __funnelshift_r(b[q].y, 0, bits >> 17) &
				       __funnelshift_r(b[q].z, 0, bits >> 22) & __funnelshift_r(b[q].w, 0, bits >> 27);

			rare |= (hit & 1u) << q;
		}
		rare &= live;
		while (rare) {
			int q = __ffs(rare) - 1;
			int64_t p = b0 + (int64_t) (q * 32 + lane) * T.s;
			uint32_t x = xs[0];
#pragma unroll
			for (int k = 1; k < NP; k++)
				x = q == k ? xs[k] : x;

			rare &= rare - 1;
			// at most 224 producers sit between this check and their reservation
			unsigned int consumed = *(volatile unsigned int *) &q_heads[0];
#pragma unroll
			for (int j = 1; j < NCONF; j++)
				consumed = min(consumed, *(volatile unsigned int *) &q_heads[j]);
			// slots below consumed - 32 * (NCONF - 1) are free for sure
			if (*(volatile unsigned int *) &q_tail + 32u * NCONF - consumed < QCAP - 256u) {
				unsigned int slot = atomicAdd(&q_tail, 1u);

				*(volatile unsigned long long *) &queue[slot % QCAP] = byteseed_candidate(p, x);
			} else {        // queue nearly full (dense hits): confirm right here
				byteseed_confirm_now(S, T, R, byteseed_candidate(p, x), own_lo, own_hi);
			}
		}
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

	uint4 b[NP];
	uint32_t xs[NP], live = 0, phase = 0;
	int64_t prev_b0 = 0;
	int stage = 0;
	bool have_prev = false;
	for (int64_t tl = gw; tl < ntiles; tl += nwarps) {
		int64_t g_lo = tl * tile, b0 = first + g_lo * T.s;
		int np = (int) min(tile, nprobes - g_lo);
		const uint8_t *sm = ring + (size_t) stage * stage_bytes;
		uint32_t nxs[NP];
		// probes of this lane in the tile: q * 32 + lane < np
		uint32_t nlive = (1u << min(NP, max(0, (np - lane + 31) >> 5))) - 1u;

		mbar_wait(&bar[stage], phase);
		{
			const int off = (int) (b0 & 15), step = 32 * T.s;
			int o = off + lane * T.s, o_max = off + (np - 1) * T.s;
#pragma unroll
			for (int q = 0; q < NP; q++) {
				uint32_t w0, w1, w2, w3;

				// past the last probe of a short tile: rehash the last one (masked by nlive)
				byteseed_window(sm, min(o, o_max), w0, w1, w2, w3);
				nxs[q] = byteseed_digest(w0, w1, w2, w3);
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
		if (have_prev)
			emit(b, xs, live, prev_b0);
#pragma unroll
		for (int q = 0; q < NP; q++) {
			b[q] = __ldg(T.filter + (uint32_t) (((uint64_t) (nxs[q] * 0x9E3779B1u) * T.nfilter) >> 32));
			xs[q] = nxs[q];
		}
		live = nlive;
		prev_b0 = b0;
		have_prev = true;
	}
	if (have_prev)
		emit(b, xs, live, prev_b0);
	__syncwarp();
	__threadfence_block();          // queue writes before the "done" count
	if (lane == 0)
		atomicAdd(&q_done, 1u);
}

// ------------------------------------------------------------------------------------
// Two-kernel form of the q-gram engine (used for the Hamming mode, where a third of the
// positions carry candidates): qprobe_kernel emits one record (entry, position) per
// candidate with warp-aggregated appends, qverify_kernel checks one candidate per thread.
// The fused qscan_kernel spends most of its issue slots in a divergent drain loop (ncu:
// 17 of 32 lanes active, 57 % issue utilisation); split, both kernels run dense.
// ------------------------------------------------------------------------------------

#include <cooperative_groups/scan.h>

__global__ void __launch_bounds__(256)
qprobe_kernel(BsgpuSubjectView S, BsgpuQIndexView Q, int64_t probe_lo, int64_t probe_hi,
	      unsigned long long *__restrict__ records, unsigned long long *__restrict__ n_records,
	      unsigned long long capacity)
{
	int64_t word_lo = probe_lo >> 5, word_hi = (probe_hi + 31) >> 5;
	int64_t j = word_lo + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	int64_t stride = (int64_t) gridDim.x * blockDim.x;

	for (; j < word_hi; j += stride) {
		int64_t base = j << 5;
		uint64_t c0 = __ldcs(S.codes + j), c1 = __ldcs(S.codes + j + 1);
		uint64_t v = (uint64_t) __ldcs(S.valid + j) | ((uint64_t) __ldcs(S.valid + j + 1) << 32);
		uint32_t m = q_valid_positions(Q, v);

		if (base < probe_lo)
			m &= 0xFFFFFFFFu << (int) (probe_lo - base);
		if (base + 32 > probe_hi)
			m &= (probe_hi - base) <= 0 ? 0u : 0xFFFFFFFFu >> (int) (base + 32 - probe_hi);
		if (m == 0)
			continue;
		uint32_t have = 0;
#pragma unroll
		for (int g = 0; g < 32; g += QSCAN_GROUP) {
			uint32_t key[QSCAN_GROUP], bw[QSCAN_GROUP];

			if (((m >> g) & ((1u << QSCAN_GROUP) - 1)) == 0)
				continue;
#pragma unroll
			for (int qq = 0; qq < QSCAN_GROUP; qq++) {
				const int r = g + qq;
				uint64_t win = r == 0 ? c0 : (c0 >> (2 * r)) | (c1 << (64 - 2 * r));

				key[qq] = q_key(Q, win);
			}
#pragma unroll
			for (int qq = 0; qq < QSCAN_GROUP; qq++)
				bw[qq] = ((m >> (g + qq)) & 1u) ? __ldg(Q.bitmap + (key[qq] >> 5)) : 0u;
#pragma unroll
			for (int qq = 0; qq < QSCAN_GROUP; qq++)
				have |= ((bw[qq] >> (key[qq] & 31u)) & 1u) << (g + qq);
		}
		while (have) {
			int r = __ffs(have) - 1;
			have &= have - 1;
			uint64_t win = r == 0 ? c0 : (c0 >> (2 * r)) | (c1 << (64 - 2 * r));
			uint32_t key = q_key(Q, win);
			uint2 G = __ldg(Q.dir + (key >> 2));
			uint32_t cnt = (G.y >> (8 * (key & 3u))) & 0xFFu;
			uint32_t off = G.x + q_prefix(G.y, (int) (key & 3u));
			// one append per warp and drain step
			cg::coalesced_group grp = cg::coalesced_threads();
			unsigned int before = cg::exclusive_scan(grp, cnt, cg::plus<unsigned int>());
			unsigned long long wbase = 0;
			if (grp.thread_rank() == grp.size() - 1)
				wbase = atomicAdd(n_records, (unsigned long long) (before + cnt));
			wbase = grp.shfl(wbase, grp.size() - 1);
			unsigned long long hi = (unsigned long long) (base + r - probe_lo) << 32;
			for (uint32_t e = 0; e < cnt; e++) {
				unsigned long long idx = wbase + before + e;

				if (idx < capacity)
					records[idx] = hi | __ldg(Q.entries + off + e);
			}
		}
	}
}

template <bool LONG>
__global__ void __launch_bounds__(256)
qverify_kernel(BsgpuSubjectView S, BsgpuQIndexView Q, BsgpuReport R, int max_nmis, int min_nmis,
	       int64_t probe_lo, int64_t own_lo, int64_t own_hi,
	       const unsigned long long *__restrict__ records, unsigned long long n)
{
	unsigned long long t = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;

	for (; t < n; t += stride) {
		unsigned long long rec = __ldcs(records + t);

		q_verify<LONG>(S, Q, R, (uint32_t) rec, probe_lo + (int64_t) (rec >> 32), max_nmis, min_nmis,
			       own_lo, own_hi);
	}
}
