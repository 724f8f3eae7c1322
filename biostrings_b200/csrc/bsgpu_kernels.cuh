// bsgpu_kernels.cuh -- sm_100a kernels of the PDict matching path.
//
//   pack_subject_kernel   bytes -> 2-bit code plane + validity planes (HBM streaming)
//   scan_tb_kernel        fixedS=TRUE Trusted-Band scan: one hash probe per window
//                         (replaces walk_subject, src/match_pdict_Twobit.c:158-175 and
//                          walk_tb_subject, src/match_pdict_ACtree2.c:1001-1021)
//   scan_tb_iupac_kernel  fixedS=FALSE windows that hold an IUPAC ambiguity letter
//                         (replaces walk_tb_nonfixed_subject, match_pdict_ACtree2.c:1124-1157)
//   verify_flanks_kernel  head/tail mismatch counting + fan-out to TB duplicates
//                         (replaces _match_pdict_all_flanks, src/match_pdict_utils.c:809-915)
//   match_plain_kernel    non-preprocessed dictionaries: every (segment, shift) x pattern
//                         (replaces the matchPattern loop of match_XStringSet_XString & co,
//                          src/match_pdict.c:174-199, src/match_pattern.c:53-110)
//   finalize_ends_kernel  sorted records -> end coordinates
// The exact whole-pattern scans are in bsgpu_xseed.cuh (w >= 25) and bsgpu_qindex.cuh (17 <= w < 25),
// the q-gram (MTB) scans in bsgpu_qindex.cuh, the band directory (Trusted Bands with heads / tails
// or duplicated bands) in bsgpu_tbdir.cuh.
#pragma once
#include <cooperative_groups.h>
#include "bsgpu_common.cuh"

namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------
// Packing: 32 letters per thread.
// ------------------------------------------------------------------------------------

// For a 32-bit word holding 4 letters: 8 code bits, 4 valid bits, 8 "not a base" bits.
// One PRMT does the per-letter table lookup: the low nibbles of the 4 bytes become the 4
// selector nibbles; selector 1/2/4 pick the table bytes for A/C/G, selector 8 (T) is
// "sign-replicate byte 0" = 0xFF because table byte 0 has its top bit set, selectors 9..15
// (IUPAC codes with bit 3) replicate the clear top bits of bytes 1..7 = 0x00, and every
// code whose low nibble is 0 ('-' '+' '.', the 0x80 separator, 0) reads table byte 0 = 0x80.
// Table value: bit 3 = valid, bits 1..2 = 2-bit code.  (Input domain: DNA codes only.)
__device__ __forceinline__ void pack4(uint32_t x, uint32_t &code8, uint32_t &valid4, uint32_t &inv8)
{
	uint32_t n = x & 0x0F0F0F0Fu;
	n = (n | (n >> 4)) & 0x00FF00FFu;
	n = (n | (n >> 8)) & 0x0000FFFFu;
	//                       byte3 byte2 byte1 byte0          byte7..byte4
	uint32_t v;                                               // A 0x08, C 0x0A, G 0x0C, T 0xFF
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(v) : "r"(0x000A0880u), "r"(0x0000000Cu), "r"(n));
	uint32_t va = (v >> 3) & 0x01010101u;                     // 1 per valid letter

	code8 = (((v >> 1) & 0x03030303u) * 0x01041040u) >> 24;
	valid4 = ((va * 0x00204081u) >> 21) & 0xFu;
	inv8 = ((va ^ 0x01010101u) * 0x01041040u) >> 24;          // bits 0, 2, 4, 6
}

// iupac plane (bit i = letter code in 1..15), only needed for fixedS = FALSE: 4 flag bits
__device__ __forceinline__ uint32_t iupac4_of(uint32_t x)
{
	uint32_t nz = (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
	uint32_t hi = x & 0xF0F0F0F0u;
	uint32_t ge16 = (((hi & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | hi) & 0x80808080u;
	uint32_t iu = nz & ~ge16;

	return (((iu >> 7) * 0x00204081u) >> 21) & 0xFu;
}

// One thread per 16 letters: every load (16 B per lane) and store (4/2/2/4 B per lane) of a
// warp is contiguous.  The 64-bit code / inv2 words and the 32-bit flag words of the planes
// are written as their 32-bit / 16-bit halves.
__global__ void __launch_bounds__(256)
pack_subject_kernel(const uint8_t *__restrict__ bytes, int64_t nwords,
		    uint64_t *__restrict__ codes, uint32_t *__restrict__ valid,
		    uint64_t *__restrict__ inv2)
{
	int64_t nhalf = 2 * nwords;
	int64_t j = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	int64_t stride = (int64_t) gridDim.x * blockDim.x;
	uint32_t *codes32 = reinterpret_cast<uint32_t *>(codes);
	uint32_t *inv32 = reinterpret_cast<uint32_t *>(inv2);
	uint16_t *valid16 = reinterpret_cast<uint16_t *>(valid);

	// 4 independent 16-byte loads in flight per thread
	for (; j < nhalf; j += 4 * stride) {
		uint4 a[4];
#pragma unroll
		for (int u = 0; u < 4; u++)
			if (j + u * stride < nhalf)
				a[u] = __ldcs(reinterpret_cast<const uint4 *>(bytes) + j + u * stride);
#pragma unroll
		for (int u = 0; u < 4; u++) {
			int64_t jj = j + u * stride;

			if (jj >= nhalf)
				break;
			uint32_t w[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
			uint32_t c = 0, n2 = 0, v = 0;
#pragma unroll
			for (int k = 0; k < 4; k++) {
				uint32_t c8, v4, n8;

				pack4(w[k], c8, v4, n8);
				c |= c8 << (8 * k);
				n2 |= n8 << (8 * k);
				v |= v4 << (4 * k);
			}
			codes32[jj] = c;
			inv32[jj] = n2;
			valid16[jj] = (uint16_t) v;
		}
	}
}

// The iupac plane is built on first use (fixedS = FALSE is the rare case).
__global__ void __launch_bounds__(256)
pack_iupac_kernel(const uint8_t *__restrict__ bytes, int64_t nwords, uint32_t *__restrict__ iupac)
{
	int64_t nhalf = 2 * nwords;
	int64_t j = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	int64_t stride = (int64_t) gridDim.x * blockDim.x;
	uint16_t *iupac16 = reinterpret_cast<uint16_t *>(iupac);

	for (; j < nhalf; j += stride) {
		uint4 a = __ldcs(reinterpret_cast<const uint4 *>(bytes) + j);

		iupac16[j] = (uint16_t) (iupac4_of(a.x) | (iupac4_of(a.y) << 4) | (iupac4_of(a.z) << 8) |
					 (iupac4_of(a.w) << 12));
	}
}

// Lays a sequence set out in the concat buffer: one warp per sequence copies its letters
// from the packed upload (src + src_off[i]) to dst + seg_start[i]; the separator bytes were
// memset before.
__global__ void __launch_bounds__(256)
scatter_set_kernel(const uint8_t *__restrict__ src, const int64_t *__restrict__ src_off,
		   const int64_t *__restrict__ seg_start, const int32_t *__restrict__ seg_len,
		   int32_t nseg, uint8_t *__restrict__ dst)
{
	int lane = threadIdx.x & 31;
	int64_t warp = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	int64_t nwarps = ((int64_t) gridDim.x * blockDim.x) >> 5;

	for (int64_t i = warp; i < nseg; i += nwarps) {
		const uint8_t *s = src + src_off[i];
		uint8_t *d = dst + seg_start[i];
		int n = seg_len[i];

		for (int k = lane; k < n; k += 32)
			d[k] = s[k];
	}
}

// ------------------------------------------------------------------------------------
// Reporting a match (shared by the scan and verify kernels).
// ------------------------------------------------------------------------------------

__device__ __forceinline__ int find_segment(const BsgpuSubjectView &S, int64_t e)
{
	// last segment with seg_start <= e
	int lo = 0, hi = S.nseg - 1;

	while (lo < hi) {
		int mid = (lo + hi + 1) >> 1;

		if (__ldg(S.seg_start + mid) <= e)
			lo = mid;
		else
			hi = mid - 1;
	}
	return lo;
}

__device__ __forceinline__ void append_record(const BsgpuReport &R, unsigned long long rec)
{
	cg::coalesced_group g = cg::coalesced_threads();
	unsigned long long base = 0;

	if (g.thread_rank() == 0)
		base = atomicAdd(R.n_records, (unsigned long long) g.size());
	base = g.shfl(base, 0);
	unsigned long long idx = base + g.thread_rank();
	if (idx < R.capacity)
		R.records[idx] = rec;
}

// key: 0-based pattern id; seg: segment of the subject; n: 1-based local TB end;
// e: concat index of the TB end; Tlen: tail width of 'key'.
// counts[i] += v.  System scope: the count vector may live in another GPU's memory (one vector
// for the GPUs of a box, bsgpu_shared_counts_*), where the reduction travels over NVLink and is
// carried out by the owner's L2; for a local vector it is the same RED as before.
__device__ __forceinline__ void count_add(int32_t *p, int32_t v)
{
	atomicAdd_system(p, v);
}

__device__ __forceinline__ void report_match(const BsgpuReport &R, const BsgpuSubjectView &S,
		int key, int seg, int64_t n, int64_t e, int Tlen)
{
	switch (R.mode) {
	case BSGPU_R_COUNT:
		count_add(R.counts + key, 1);
		break;
	case BSGPU_R_HIT_COORD: {
		// end = tb_end + |tail|   (src/match_pdict_utils.c:99-106)
		unsigned long long end = (unsigned long long) (n + Tlen + __ldg(S.seg_coord + seg));
		append_record(R, ((unsigned long long) key << BSGPU_POS_BITS) | end);
		break;
	}
	case BSGPU_R_HIT_TBPOS:
		append_record(R, ((unsigned long long) key << BSGPU_POS_BITS) | (unsigned long long) e);
		break;
	case BSGPU_R_VMAT:
		count_add(R.counts + (int64_t) seg * R.M + key, 1);
		break;
	case BSGPU_R_VC1:
		count_add(R.counts + key, __ldg(R.weight + seg));
		break;
	case BSGPU_R_VC2:
		count_add(R.counts + seg, __ldg(R.weight + key));
		break;
	case BSGPU_R_SEGKEY:
		append_record(R, ((unsigned long long) seg << BSGPU_SEGKEY_BITS) | (unsigned long long) key);
		break;
	}
}

// Mismatches of a 2-bit packed flank (len <= 32 base letters, letter i at bits 2i) against the
// subject letters [a, a + len): XOR of the code planes, the two bits of a letter folded together,
// ORed with the "not a base" plane, popcount.  Letters outside the segment [lo, lo + L) count as
// mismatches (src/lowlevel_matching.c:79-86); so do non-base subject letters, which is the
// fixedS = TRUE table for a base-only pattern (src/lowlevel_matching.c:26-56).
__device__ __forceinline__ int packed_flank_mismatches(const BsgpuSubjectView &S, int64_t a, int len,
		uint64_t pat, int64_t lo, int64_t L)
{
	const uint64_t E = 0x5555555555555555ull;

	if (len == 0)
		return 0;
	int64_t wi = a >> 5;
	int sh = 2 * (int) (a & 31);
	uint64_t c0 = __ldg(S.codes + wi), c1 = __ldg(S.codes + wi + 1);
	uint64_t i0 = __ldg(S.inv2 + wi), i1 = __ldg(S.inv2 + wi + 1);
	uint64_t x = sh ? (c0 >> sh) | (c1 << (64 - sh)) : c0;
	uint64_t iv = sh ? (i0 >> sh) | (i1 << (64 - sh)) : i0;
	uint64_t d = x ^ pat;

	d = ((d | (d >> 1)) | iv) & E;
	int nbefore = (int) min((int64_t) len, max((int64_t) 0, lo - a));
	int nafter = (int) min((int64_t) len, max((int64_t) 0, a + len - (lo + L)));
	if (nbefore)
		d |= E & (nbefore >= 32 ? ~0ull : (1ull << (2 * nbefore)) - 1);
	if (nafter)
		d |= E & ~((len - nafter) >= 32 ? ~0ull : (1ull << (2 * (len - nafter))) - 1);
	d &= len >= 32 ? E : E & ((1ull << (2 * len)) - 1);
	return __popcll(d);
}

// A confirmed Trusted-Band hit of TB id 'tb_id' ending at concat index e.
// direct == 1: no head/tail and singleton TB groups -> report right away;
// direct == 2: every key's head and tail are packed (base-only, <= 32 letters) and fixedS = TRUE:
//              the keys of the TB group {key0} U low2high[key0] (src/match_pdict_utils.c:153-181)
//              are verified right here, head + tail mismatches by XOR + popcount against
//              max / min.mismatch (src/match_pdict_utils.c:183-214 counts the same letters);
// direct == 0: the hit becomes a candidate record (tb_id << 34 | e) for verify_flanks_kernel.
__device__ __forceinline__ void tb_hit(const BsgpuIndexView &I, const BsgpuSubjectView &S,
		const BsgpuReport &R, const BsgpuReport &C, int direct, int tb_id, int64_t e,
		int max_nmis = 0, int min_nmis = 0)
{
	if (direct == 2) {
		int seg = S.nseg == 1 ? 0 : find_segment(S, e);
		int64_t lo = __ldg(S.seg_start + seg), L = __ldg(S.seg_len + seg);
		int g0 = __ldg(I.grp_off + tb_id), g1 = __ldg(I.grp_off + tb_id + 1);

		for (int g = g0; g < g1; g++) {
			int key = g == g0 ? tb_id : __ldg(I.grp_keys + g);      // the leader comes first
			uint32_t fl = __ldg(I.flank_len + key);
			int hl = (int) (fl & 63u), tl = (int) ((fl >> 6) & 63u);
			ulonglong2 f = __ldg(I.flank2 + key);
			int nmis = packed_flank_mismatches(S, e - I.w - hl + 1, hl, f.x, lo, L);

			if (nmis <= max_nmis)
				nmis += packed_flank_mismatches(S, e + 1, tl, f.y, lo, L);
			if (nmis <= max_nmis && nmis >= min_nmis)
				report_match(R, S, key, seg, e - lo + 1, e, tl);
		}
		return;
	}
	if (!direct) {
		append_record(C, ((unsigned long long) tb_id << BSGPU_POS_BITS) | (unsigned long long) e);
		return;
	}
	if (R.mode == BSGPU_R_COUNT) {
		count_add(R.counts + tb_id, 1);
		return;
	}
	int seg = S.nseg == 1 ? 0 : find_segment(S, e);
	int64_t n = e - __ldg(S.seg_start + seg) + 1;
	report_match(R, S, tb_id, seg, n, e, 0);
}

// ------------------------------------------------------------------------------------
// Hash probe.
// ------------------------------------------------------------------------------------

// Full probe: walks the buckets from the home bucket while the overflow flag says that
// a key hashed here was displaced to a later bucket.  Returns the TB id or -1.
__device__ __forceinline__ int probe_from(const BsgpuIndexView &I, uint64_t key, uint32_t bucket,
					  uint32_t fp)
{
	for (;;) {
		uint4 b = __ldg(I.table + bucket);
		uint32_t id0 = b.y & BSGPU_ID_MASK, id1 = b.w & BSGPU_ID_MASK;

		if (b.x == fp && id0 != BSGPU_EMPTY_ID && __ldg(I.keys + id0) == key)
			return (int) id0;
		if (b.z == fp && id1 != BSGPU_EMPTY_ID && __ldg(I.keys + id1) == key)
			return (int) id1;
		if (!(b.w & BSGPU_OVERFLOW_FLAG))
			return -1;
		bucket = (bucket + 1) & I.bucket_mask;
	}
}

__device__ __forceinline__ int probe(const BsgpuIndexView &I, uint64_t key)
{
	uint32_t top, fp;

	bsgpu_hash(key, top, fp);
	return probe_from(I, key, top >> I.bucket_shift, fp);
}

// For TB widths > 32 the hash covers the first 32 letters; the rest is compared here.
// fixedS: s == t; otherwise s < 16 && (s & t) != 0  (TB letters t are base letters).
__device__ __forceinline__ bool tb_rest_matches(const BsgpuIndexView &I, const BsgpuSubjectView &S,
		int tb_id, int64_t e, int fixedS)
{
	const uint8_t *t = I.tb_bytes + (int64_t) tb_id * I.w;
	const uint8_t *s = S.bytes + (e - I.w + 1);

	for (int i = 32; i < I.w; i++) {
		uint32_t sb = s[i], tb = t[i];

		if (fixedS ? sb != tb : (sb >= 16 || (sb & tb) == 0))
			return false;
	}
	return true;
}

// TB leaders sharing a 32-letter seed are chained; with fixedS at most one can match.
__device__ __forceinline__ int resolve_long_tb(const BsgpuIndexView &I, const BsgpuSubjectView &S,
		int id, int64_t e, int fixedS)
{
	for (; id >= 0; id = __ldg(I.seed_next + id))
		if (tb_rest_matches(I, S, id, e, fixedS))
			return id;
	return -1;
}

// mask bit r set iff bits r..r+w-1 of v are all set (w in 1..64)
__device__ __forceinline__ uint64_t all_set_run(uint64_t v, int w)
{
	uint64_t m = v;
	int have = 1;           // m bit r <=> v[r .. r+have-1] all set

	while (have * 2 <= w) {
		m &= m >> have;
		have *= 2;
	}
	if (have < w)
		m &= m >> (w - have);
	return m;
}

// mask bit r set iff the flag bits of the w letters starting at letter 32 * j + r are all
// set; 'first' holds the flag words j and j + 1 (already loaded), further words are read
// 32 letters at a time (w of any size)
__device__ __forceinline__ uint32_t window_mask(const uint32_t *__restrict__ plane, int64_t j,
						uint64_t first, int w)
{
	uint32_t m = (uint32_t) all_set_run(first, min(w, 32));

	for (int k = 32; k < w && m != 0; k += 32) {
		int64_t jj = j + (k >> 5);
		uint64_t v = (uint64_t) __ldg(plane + jj) | ((uint64_t) __ldg(plane + jj + 1) << 32);

		m &= (uint32_t) all_set_run(v, min(w - k, 32));
	}
	return m;
}

// ------------------------------------------------------------------------------------
// Trusted-Band scan, fixedS = TRUE: every window of w base letters is looked up.
// Thread t handles the 32 window starts of code word j; windows may extend into words
// j+1 (and j+2 when w > 32: only the first 32 letters are hashed, the rest compared).
// ------------------------------------------------------------------------------------

#define SCAN_GROUP 8

__global__ void __launch_bounds__(256)
scan_tb_kernel(BsgpuSubjectView S, BsgpuIndexView I, BsgpuReport R, BsgpuReport C,
	       int direct, int64_t start_lo, int64_t start_hi, int max_nmis, int min_nmis)
{
	// window starts c0 in [start_lo, start_hi)
	int64_t word_lo = start_lo >> 5, word_hi = (start_hi + 31) >> 5;
	int64_t j = word_lo + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	int64_t stride = (int64_t) gridDim.x * blockDim.x;
	const int w = I.w, sw = I.seed_w;

	for (; j < word_hi; j += stride) {
		uint64_t c0 = __ldcs(S.codes + j), c1 = __ldcs(S.codes + j + 1);
		uint64_t v = (uint64_t) __ldcs(S.valid + j) | ((uint64_t) __ldcs(S.valid + j + 1) << 32);
		uint32_t m = window_mask(S.valid, j, v, w);

		// clip to [start_lo, start_hi)
		int64_t base = j << 5;
		if (base < start_lo)
			m &= 0xFFFFFFFFu << (int) (start_lo - base);
		if (base + 32 > start_hi)
			m &= (start_hi - base) <= 0 ? 0u : 0xFFFFFFFFu >> (int) (base + 32 - start_hi);
		if (m == 0)
			continue;
		// Phase 1: one 16-byte bucket load per window, SCAN_GROUP loads in flight; a
		// window needs a second look only if a fingerprint matches or the home bucket
		// overflowed (about 1 % of the windows of a half-full table).
		uint32_t rare = 0;
#pragma unroll
		for (int g = 0; g < 32; g += SCAN_GROUP) {
			uint32_t fp[SCAN_GROUP], bucket[SCAN_GROUP];
			uint4 b[SCAN_GROUP];

			if (((m >> g) & ((1u << SCAN_GROUP) - 1)) == 0)
				continue;
#pragma unroll
			for (int q = 0; q < SCAN_GROUP; q++) {
				const int r = g + q;
				uint64_t k = r == 0 ? c0 : (c0 >> (2 * r)) | (c1 << (64 - 2 * r));
				uint32_t top;

				bsgpu_hash(k & I.seed_mask, top, fp[q]);
				bucket[q] = top >> I.bucket_shift;
			}
#pragma unroll
			for (int q = 0; q < SCAN_GROUP; q++) {
				if ((m >> (g + q)) & 1u)
					b[q] = __ldg(I.table + bucket[q]);
				else
					b[q] = make_uint4(~fp[q], BSGPU_EMPTY_ID, ~fp[q], BSGPU_EMPTY_ID);
			}
#pragma unroll
			for (int q = 0; q < SCAN_GROUP; q++) {
				bool look = b[q].x == fp[q] || b[q].z == fp[q] ||
					    (b[q].w & BSGPU_OVERFLOW_FLAG);
				rare |= (look ? 1u : 0u) << (g + q);
			}
		}
		// Phase 2: the few windows that need the full probe, all lanes together.
		while (rare) {
			int r = __ffs(rare) - 1;
			rare &= rare - 1;
			uint64_t k = r == 0 ? c0 : (c0 >> (2 * r)) | (c1 << (64 - 2 * r));
			k &= I.seed_mask;
			int id = probe(I, k);
			int64_t e = base + r + w - 1;

			if (id >= 0 && w > 32)
				id = resolve_long_tb(I, S, id, e, 1);
			if (id >= 0)
				tb_hit(I, S, R, C, direct, id, e, max_nmis, min_nmis);
		}
	}
}

// ------------------------------------------------------------------------------------
// Trusted-Band scan, fixedS = FALSE, windows holding at least one IUPAC ambiguity letter
// (all-base windows are handled by scan_tb_kernel: for them (s & t) != 0 <=> s == t).
// A window matches TB word t iff every letter is < 16 (and not 0) and shares a bit with
// t[i]; all expansions of the ambiguity letters are looked up.  Windows whose number of
// expansions exceeds max_expansions raise *too_many (the reference errors the same way
// when its node set overflows, src/match_pdict_ACtree2.c:1051-1054).
// ------------------------------------------------------------------------------------

#define IUPAC_MAX_AMBIG 12

__global__ void __launch_bounds__(128)
scan_tb_iupac_kernel(BsgpuSubjectView S, BsgpuIndexView I, BsgpuReport R, BsgpuReport C,
		     int direct, int64_t start_lo, int64_t start_hi,
		     unsigned int max_expansions, long long *ovf_pos, unsigned long long *n_ovf,
		     unsigned long long ovf_capacity)
{
	// A warp takes 32 words (1024 window starts) at a time; the windows holding an ambiguity letter
	// come in runs of w (one letter spoils w consecutive windows, all in one or two lanes' words), so
	// they are listed in shared memory first and dealt round-robin to the lanes.
	__shared__ uint16_t list_all[4][1024];
	const unsigned FULL = 0xFFFFFFFFu;
	const int lane = threadIdx.x & 31;
	uint16_t *list = list_all[threadIdx.x >> 5];
	int64_t word_lo = start_lo >> 5, word_hi = (start_hi + 31) >> 5;
	int64_t jb = word_lo + ((int64_t) blockIdx.x * blockDim.x + (threadIdx.x & ~31));
	int64_t stride = (int64_t) gridDim.x * blockDim.x;
	const int w = I.w, sw = I.seed_w;

	for (; jb < word_hi; jb += stride) {
		const int64_t j = jb + lane;
		uint32_t m = 0;

		if (j < word_hi) {
			uint64_t u0 = (uint64_t) __ldg(S.iupac + j) | ((uint64_t) __ldg(S.iupac + j + 1) << 32);
			uint64_t v0 = (uint64_t) __ldg(S.valid + j) | ((uint64_t) __ldg(S.valid + j + 1) << 32);
			uint32_t m_iu = window_mask(S.iupac, j, u0, w), m_va = window_mask(S.valid, j, v0, w);
			int64_t base = j << 5;

			m = m_iu & ~m_va;       // IUPAC-only windows that are not all-base
			if (base < start_lo)
				m &= 0xFFFFFFFFu << (int) (start_lo - base);
			if (base + 32 > start_hi)
				m &= (start_hi - base) <= 0 ? 0u : 0xFFFFFFFFu >> (int) (base + 32 - start_hi);
		}
		if (!__any_sync(FULL, m != 0u))
			continue;
		int mine = __popc(m), before = mine;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			int up = __shfl_up_sync(FULL, before, d);
			if (lane >= d)
				before += up;
		}
		const int nitems = __shfl_sync(FULL, before, 31);
		before -= mine;
		__syncwarp();
		for (uint32_t h = m; h != 0u; h &= h - 1)
			list[before++] = (uint16_t) (lane * 32 + __ffs(h) - 1);
		__syncwarp();
		for (int it = lane; it < nitems; it += 32) {
			int64_t c0 = (jb << 5) + list[it], e = c0 + w - 1;
			const uint8_t *s = S.bytes + c0;
			// ambiguity letters inside the hashed seed (first sw letters)
			uint64_t key0 = 0;
			int apos[IUPAC_MAX_AMBIG], namb = 0;
			uint32_t alet[IUPAC_MAX_AMBIG];
			unsigned long long total = 1;
			bool overflow = false;

			for (int i = 0; i < sw; i++) {
				uint32_t b = s[i];

				if (bsgpu_is_base(b)) {
					key0 |= (uint64_t) bsgpu_base_code(b) << (2 * i);
				} else {
					total *= __popc(b);
					if (namb == IUPAC_MAX_AMBIG || total > max_expansions) {
						overflow = true;
						break;
					}
					apos[namb] = i;
					alet[namb] = b;
					namb++;
				}
			}
			if (overflow) {
				// too many expansions to enumerate: the window goes to the
				// brute-force kernel (every Trusted-Band word against it)
				unsigned long long k = atomicAdd(n_ovf, 1ull);

				if (k < ovf_capacity)
					ovf_pos[k] = c0;
				continue;
			}
			for (unsigned long long combo = 0; combo < total; combo++) {
				uint64_t key = key0;
				unsigned long long rem = combo;

				for (int a = 0; a < namb; a++) {
					uint32_t let = alet[a];
					int nb = __popc(let), d = (int) (rem % nb);

					rem /= nb;
					// d-th set bit of the letter
					for (int q = 0; q < d; q++)
						let &= let - 1;
					uint32_t bit = let & (0u - let);
					key |= (uint64_t) bsgpu_base_code(bit) << (2 * apos[a]);
				}
				int id = probe(I, key);
				if (w > 32) {
					// several TB words may share the seed and match the
					// (ambiguous) rest: report each of them
					for (; id >= 0; id = __ldg(I.seed_next + id))
						if (tb_rest_matches(I, S, id, e, 0))
							tb_hit(I, S, R, C, direct, id, e);
				} else if (id >= 0) {
					tb_hit(I, S, R, C, direct, id, e);
				}
			}
		}
	}
}

// Windows with more IUPAC expansions than scan_tb_iupac_kernel enumerates (runs of N ...):
// every Trusted-Band leader is compared with the window letter by letter,
// (S[i] & t[i]) != 0 for all i -- what the reference's node-set walk amounts to
// (src/match_pdict_ACtree2.c:1028-1061).  blockIdx.y walks the windows, x the leaders.
__global__ void __launch_bounds__(256)
iupac_brute_kernel(BsgpuSubjectView S, BsgpuIndexView I, BsgpuReport R, BsgpuReport C, int direct,
		   const long long *__restrict__ ovf_pos, unsigned long long n_ovf,
		   const int32_t *__restrict__ leaders, int32_t n_leaders)
{
	__shared__ uint8_t win[64];
	const int w = I.w;

	for (unsigned long long q = blockIdx.y; q < n_ovf; q += gridDim.y) {
		int64_t c0 = ovf_pos[q];

		__syncthreads();
		if (threadIdx.x < w)
			win[threadIdx.x] = S.bytes[c0 + threadIdx.x];
		__syncthreads();
		for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_leaders; k += gridDim.x * blockDim.x) {
			int id = __ldg(leaders + k);
			const uint8_t *t = I.tb_bytes + (int64_t) id * w;
			bool ok = true;

			for (int i = 0; i < w && ok; i++)
				ok = (win[i] & t[i]) != 0;
			if (ok)
				tb_hit(I, S, R, C, direct, id, c0 + w - 1);
		}
	}
}

// ------------------------------------------------------------------------------------
// Flank verification.  One thread per candidate (tb_id, e); the thread walks the TB
// duplicate group {key0} U low2high[key0] (src/match_pdict_utils.c:153-181) and counts
// head-then-tail mismatches with the reference's early exit (:183-214); letters outside
// the segment count as mismatches (src/lowlevel_matching.c:79-86).
// ------------------------------------------------------------------------------------

__device__ __forceinline__ bool letters_match(uint32_t p, uint32_t s, int fixedP, int fixedS)
{
	// src/lowlevel_matching.c:26-56, on the raw code bytes
	if (fixedP)
		return fixedS ? p == s : (p & ~s & 0xFFu) == 0;
	return fixedS ? (~p & s & 0xFFu) == 0 : (p & s) != 0;
}

__device__ __forceinline__ int count_mismatches(const uint8_t *__restrict__ P, int Plen,
		const uint8_t *__restrict__ seg, int64_t L, int64_t shift, int max_nmis,
		int fixedP, int fixedS)
{
	// src/lowlevel_matching.c:66-89 (_nmismatch_at_Pshift)
	int nmis = 0;

	for (int i = 0; i < Plen; i++) {
		int64_t jj = shift + i;

		if (jj >= 0 && jj < L && letters_match(P[i], seg[jj], fixedP, fixedS))
			continue;
		if (nmis++ >= max_nmis)
			break;
	}
	return nmis;
}

__global__ void __launch_bounds__(256)
verify_flanks_kernel(BsgpuSubjectView S, BsgpuIndexView I, BsgpuReport R,
		     const unsigned long long *__restrict__ cand, unsigned long long ncand,
		     int max_nmis, int min_nmis, int fixedP, int fixedS)
{
	unsigned long long t = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;

	for (; t < ncand; t += stride) {
		unsigned long long rec = cand[t];
		int key0 = (int) (rec >> BSGPU_POS_BITS);
		int64_t e = (int64_t) (rec & BSGPU_POS_MASK);
		int seg = S.nseg == 1 ? 0 : find_segment(S, e);
		int64_t lo = __ldg(S.seg_start + seg), L = __ldg(S.seg_len + seg);
		int64_t n = e - lo + 1;                 // 1-based TB end inside the segment
		const uint8_t *segp = S.bytes + lo;
		int g0 = __ldg(I.grp_off + key0), g1 = __ldg(I.grp_off + key0 + 1);

		for (int g = g0; g < g1; g++) {
			int key = __ldg(I.grp_keys + g);
			int64_t h0 = I.head ? __ldg(I.head_off + key) : 0;
			int Hlen = I.head ? (int) (__ldg(I.head_off + key + 1) - h0) : 0;
			int64_t t0 = I.tail ? __ldg(I.tail_off + key) : 0;
			int Tlen = I.tail ? (int) (__ldg(I.tail_off + key + 1) - t0) : 0;
			// src/match_pdict_utils.c:199-214: Hshift = tb_end - |H| - w, Tshift = tb_end
			int nmis = count_mismatches(I.head + h0, Hlen, segp, L, n - Hlen - I.w,
						    max_nmis, fixedP, fixedS);
			if (nmis <= max_nmis)
				nmis += count_mismatches(I.tail + t0, Tlen, segp, L, n,
							 max_nmis - nmis, fixedP, fixedS);
			if (nmis <= max_nmis && nmis >= min_nmis)
				report_match(R, S, key, seg, n, e, Tlen);
		}
	}
}

// ------------------------------------------------------------------------------------
// Sorted records -> per-pattern counts and end coordinates.
// ------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------
// Non-preprocessed dictionaries (pattern = plain XStringSet: any letters, any widths;
// match_XStringSet_XString and friends, src/match_pdict.c:174-199,267-293,348-384,434-481).
// The reference runs one matchPattern() per pattern; all its algorithms report the same
// matches as the naive inexact one (src/match_pattern.c:53-76): pattern shift Pshift (0-based
// position of the pattern's first letter in the subject, may be out of limits) matches iff
// min <= nmis <= max, letters outside the subject counting as mismatches, for
//   minP <= Pshift <= L - plen - minP,  minP = plen <= max ? 1 - plen : -max.
// A "slot" is one (segment, Pshift) pair: segment g owns the slots slot_off[g] ..
// slot_off[g+1]-1, Pshift = slot - slot_off[g] - ext (ext = max.mismatch); slots are
// monotone in (segment, Pshift), which is the order the reference reports matches in.
// ------------------------------------------------------------------------------------

struct BsgpuPlainView {
	const uint8_t *pat;          // concatenated pattern letters (+8 readable bytes)
	const int64_t *pat_off;      // [M+1]
	const uint4 *pat_head;       // [M] {offset lo, offset hi, length, first 4 letters}
	int32_t M;
	const int64_t *slot_off;     // [nseg+1]
	int32_t ext;
	int64_t nslots;
	int64_t own_lo, own_hi;      // subject chunks: concat index range of the owned match ends
};

__device__ __forceinline__ int find_slot_segment(const BsgpuPlainView &P, int nseg, int64_t slot)
{
	int lo = 0, hi = nseg - 1;

	while (lo < hi) {
		int mid = (lo + hi + 1) >> 1;

		if (__ldg(P.slot_off + mid) <= slot)
			lo = mid;
		else
			hi = mid - 1;
	}
	return lo;
}

// number of the 4 letter pairs of two words that do not match (src/lowlevel_matching.c:26-56,
// four bytes at a time); bytes whose 'keep' byte is 0 are ignored
__device__ __forceinline__ int mismatches4(uint32_t p, uint32_t s, uint32_t keep, int fixedP, int fixedS)
{
	uint32_t x = fixedP ? (fixedS ? p ^ s : p & ~s) : (fixedS ? ~p & s : p & s);
	// 0x80 in every byte of x that is not 0
	uint32_t nz = (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;

	if (!fixedP && !fixedS)
		nz = ~nz & 0x80808080u;         // (p & s) != 0 is the match
	return __popc(nz & keep);
}

// the 4 bytes at b (any alignment; b - 3 .. b + 7 readable) as a little-endian word
__device__ __forceinline__ uint32_t load_word_unaligned(const uint8_t *b)
{
	const uint32_t *w = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(b) & ~(uintptr_t) 3);

	return __funnelshift_r(__ldg(w), __ldg(w + 1), 8u * (uint32_t) (reinterpret_cast<uintptr_t>(b) & 3));
}

// blockIdx.x strides over the slots, blockIdx.y over the patterns.  A thread keeps the first
// 4 subject letters of its shift in a register; most (pattern, shift) pairs are rejected by
// one broadcast 16-byte load of the pattern's header and one 4-letter compare.  Shifts that
// hang over an end of the segment take the letter-by-letter path.
__global__ void __launch_bounds__(256)
match_plain_kernel(BsgpuSubjectView S, BsgpuPlainView P, BsgpuReport R, int max_nmis, int min_nmis,
		   int fixedP, int fixedS)
{
	int64_t slot = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	int64_t stride = (int64_t) gridDim.x * blockDim.x;

	for (; slot < P.nslots; slot += stride) {
		int seg = S.nseg == 1 ? 0 : find_slot_segment(P, S.nseg, slot);
		int64_t L = __ldg(S.seg_len + seg);
		int64_t Pshift = slot - __ldg(P.slot_off + seg) - P.ext;
		const uint8_t *s = S.bytes + __ldg(S.seg_start + seg);
		// first 4 letters at this shift; only used when the whole pattern lies inside the
		// segment (the separator / guard bytes after the segment make the read safe)
		uint32_t s0 = Pshift >= 0 && Pshift <= L ? load_word_unaligned(s + Pshift) : 0u;

		for (int i = blockIdx.y; i < P.M; i += gridDim.y) {
			uint4 h = __ldg(P.pat_head + i);
			int64_t po = (int64_t) h.x | ((int64_t) h.y << 32);
			int plen = (int) h.z;
			int64_t minP = plen <= max_nmis ? 1 - plen : -max_nmis;
			int nmis;

			if (Pshift < minP || Pshift + plen > L - minP)
				continue;
			int64_t e_idx = (s - S.bytes) + Pshift + plen - 1;      // concat index of the last letter
			if (e_idx < P.own_lo || e_idx >= P.own_hi)
				continue;
			if (Pshift >= 0 && Pshift + plen <= L) {
				uint32_t keep = plen >= 4 ? 0x80808080u : (0x80808080u >> (8 * (4 - plen)));

				nmis = mismatches4(h.w, s0, keep, fixedP, fixedS);
				for (int j = 4; j < plen && nmis <= max_nmis; j += 4) {
					int left = plen - j;

					keep = left >= 4 ? 0x80808080u : (0x80808080u >> (8 * (4 - left)));
					nmis += mismatches4(load_word_unaligned(P.pat + po + j),
							    load_word_unaligned(s + Pshift + j), keep, fixedP, fixedS);
				}
			} else {
				nmis = count_mismatches(P.pat + po, plen, s, L, Pshift, max_nmis, fixedP, fixedS);
			}
			if (nmis > max_nmis || nmis < min_nmis)
				continue;
			if (R.mode == BSGPU_R_COUNT)
				count_add(R.counts + i, 1);
			else            // n = 1-based local end; the record position is the slot
				report_match(R, S, i, seg, Pshift + plen, slot, 0);
		}
	}
}

// mode HIT_COORD: low bits already are the end coordinate.
// mode HIT_TBPOS: low bits = concat index of the TB end; end = local + |tail| + seg_coord.
// plain dictionaries (P.slot_off != NULL): low bits = slot; end = Pshift + plen + seg_coord.
__global__ void __launch_bounds__(256)
finalize_ends_kernel(BsgpuSubjectView S, const int64_t *__restrict__ tail_off, int tbpos_mode,
		     BsgpuPlainView P, const unsigned long long *__restrict__ rec, unsigned long long n,
		     int32_t *__restrict__ ends, int32_t *__restrict__ counts)
{
	unsigned long long t = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;

	for (; t < n; t += stride) {
		unsigned long long r = rec[t];
		int key = (int) (r >> BSGPU_POS_BITS);
		int64_t v = (int64_t) (r & BSGPU_POS_MASK);

		if (P.slot_off != nullptr) {
			int seg = S.nseg == 1 ? 0 : find_slot_segment(P, S.nseg, v);
			int plen = (int) (__ldg(P.pat_off + key + 1) - __ldg(P.pat_off + key));

			v = v - __ldg(P.slot_off + seg) - P.ext + plen + __ldg(S.seg_coord + seg);
		} else if (tbpos_mode) {
			int seg = S.nseg == 1 ? 0 : find_segment(S, v);
			int Tlen = tail_off ? (int) (__ldg(tail_off + key + 1) - __ldg(tail_off + key)) : 0;

			v = v - __ldg(S.seg_start + seg) + 1 + Tlen + __ldg(S.seg_coord + seg);
		}
		ends[t] = (int32_t) v;
		if (counts != nullptr)
			count_add(counts + key, 1);     // (may be the caller's vector on another GPU)
	}
}

// ------------------------------------------------------------------------------------
// Count - scan - write compaction of the hit records into rows (the reference's MatchBuf: one
// int list per pattern, src/match_pdict_utils.c:87-116, src/match_reporting.c:150-201).
//   1. count_rows_kernel   counts[key]++ for every record (key << 34 | position)
//   2. exclusive scan of the counts -> row offsets (the CSR the host receives)
//   3. write_rows_kernel   every record takes a place in its row (per-row cursor)
//   4. order_rows_kernel   rows are tiny: one thread puts a row of <= 32 records in ascending
//      position order (the order the reference appends them in, view by view); the rare
//      longer rows (repeats) are listed and sorted by bitonic_rows_kernel, one block per row
//   5. finalize_ends_kernel turns the ordered records into end coordinates
// No global sort: the work is proportional to the hits and stays inside the rows.
// ------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
count_rows_kernel(const unsigned long long *__restrict__ rec, unsigned long long n, int32_t *__restrict__ counts)
{
	unsigned long long t = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;

	for (; t < n; t += stride)
		atomicAdd(counts + (rec[t] >> BSGPU_POS_BITS), 1);
}

__global__ void __launch_bounds__(256)
write_rows_kernel(const unsigned long long *__restrict__ rec, unsigned long long n,
		  const int64_t *__restrict__ row_off, int32_t *__restrict__ cursor,
		  unsigned long long *__restrict__ rows)
{
	unsigned long long t = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;

	for (; t < n; t += stride) {
		unsigned long long r = rec[t];
		unsigned long long key = r >> BSGPU_POS_BITS;

		rows[row_off[key] + atomicAdd(cursor + key, 1)] = r;
	}
}

#define BSGPU_SMALL_ROW 32

__global__ void __launch_bounds__(256)
order_rows_kernel(unsigned long long *__restrict__ rows, const int64_t *__restrict__ row_off,
		  const int32_t *__restrict__ counts, int32_t M, int32_t *__restrict__ big, int32_t *__restrict__ n_big)
{
	int32_t key = blockIdx.x * blockDim.x + threadIdx.x;
	int32_t stride = gridDim.x * blockDim.x;

	for (; key < M; key += stride) {
		int n = counts[key];

		if (n < 2)
			continue;
		if (n > BSGPU_SMALL_ROW) {
			big[atomicAdd(n_big, 1)] = key;
			continue;
		}
		unsigned long long *row = rows + row_off[key];
		for (int i = 1; i < n; i++) {           // insertion sort: rows arrive almost in order
			unsigned long long v = row[i];
			int j = i - 1;

			while (j >= 0 && row[j] > v) {
				row[j + 1] = row[j];
				j--;
			}
			row[j + 1] = v;
		}
	}
}

// one block per long row: bitonic sort in place (the row is padded to a power of two with
// virtual +infinity elements).  Launched twice: small blocks (cheap barriers, many rows in flight)
// for the rows of n_lo < n <= n_hi records, 1024 threads for the few longer ones.
__global__ void __launch_bounds__(1024)
bitonic_rows_kernel(unsigned long long *__restrict__ rows, const int64_t *__restrict__ row_off,
		    const int32_t *__restrict__ counts, const int32_t *__restrict__ big, const int32_t *__restrict__ n_big,
		    unsigned int n_lo, unsigned int n_hi)
{
	for (int b = blockIdx.x; b < *n_big; b += gridDim.x) {
		int32_t key = big[b];
		unsigned long long *row = rows + row_off[key];
		const unsigned int n = (unsigned int) counts[key];
		unsigned int np2 = 1;

		if (n <= n_lo || n > n_hi)
			continue;               // the other launch's row (uniform over the block)

		while (np2 < n)
			np2 <<= 1;
		// the network with all comparators ascending (first step of a merge pairs i with its
		// mirror i ^ (k - 1)): the virtual tail never has to move
		for (unsigned int k = 2; k <= np2; k <<= 1)
			for (unsigned int j = k >> 1; j > 0; j >>= 1) {
				const unsigned int flip = j == (k >> 1) ? k - 1 : j;

				for (unsigned int i = threadIdx.x; i < np2; i += blockDim.x) {
					unsigned int l = i ^ flip;

					if (l > i && l < n) {
						unsigned long long a = row[i], c = row[l];

						if (a > c) {
							row[i] = c;
							row[l] = a;
						}
					}
				}
				__syncthreads();
			}
	}
}

// flag[t] = 1 iff record t differs from record t-1 (input sorted)
__global__ void __launch_bounds__(256)
count_segkey_kernel(const unsigned long long *__restrict__ rec, unsigned long long n,
		    int64_t *__restrict__ seg_counts)
{
	unsigned long long t = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;

	for (; t < n; t += stride)
		atomicAdd((unsigned long long *) (seg_counts + (rec[t] >> BSGPU_SEGKEY_BITS)), 1ull);
}

// ------------------------------------------------------------------------------------
// Roofline probes.
// ------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
probe_copy_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, int64_t n)
{
	int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	int64_t stride = (int64_t) gridDim.x * blockDim.x;

	for (; i < n; i += stride)
		dst[i] = __ldcs(src + i);
}

// every lane reads one random 16-byte half of a random 32-byte sector, INFLIGHT loads per
// thread issued before any is consumed (the access shape of the index probes)
template <int INFLIGHT>
__global__ void __launch_bounds__(256)
probe_gather_kernel(const uint4 *__restrict__ table, uint32_t mask, int iters,
		    unsigned long long *__restrict__ sink)
{
	uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
	uint32_t acc = 0;

	for (int it = 0; it < iters; it += INFLIGHT) {
		uint4 b[INFLIGHT];
#pragma unroll
		for (int q = 0; q < INFLIGHT; q++) {
			x = x * 1664525u + 1013904223u;
			b[q] = __ldg(table + ((x >> 4) & mask));
		}
#pragma unroll
		for (int q = 0; q < INFLIGHT; q++)
			acc += b[q].x ^ b[q].w;
	}
	if (acc == 0x12345678u)
		atomicAdd(sink, 1ull);
}
