// bsgpu_fasta.cuh -- FASTA text -> encoded letters in the subject layout, on the device
// (subject ingestion).
//
// Replaces the two passes of read_fasta_files() over parse_FASTA_file()
// (src/read_fasta_files.c:240-352,412-451) for one in-memory file:
//   * lines end at LF; a CR right before that LF belongs to the line end
//     (delete_trailing_LF_or_CRLF), any other CR is an ordinary byte;
//   * empty lines and lines starting with ';' are ignored (:283-294);
//   * a line starting with '>' opens a new record, the rest of the line is its description
//     (:297-321);
//   * every other line is sequence data: each byte goes through the integer lookup `lkup`
//     (translate(), :31-50); bytes the lookup does not know are dropped and counted
//     ("ignored %lld invalid one-letter sequence codes", :388-391), the others are appended to
//     the current record;
//   * a non-empty sequence line before the first '>' is an error (:322-327).
// The reference reads lines into a 200001-byte buffer; a longer sequence line is simply read in
// pieces (same result), a longer description or comment line is an error there ("line is too
// long", :286-291) and accepted here.
//
// Parallel formulation.  The text is cut into chunks of BSGPU_FASTA_CHUNK bytes, one thread per
// chunk walks its bytes with the small state machine above.  What a chunk cannot know locally is
// the class of the line it starts in the middle of, so pass 1 (fasta_summary_kernel) counts the
// letters of that first partial line as if it were sequence data, separately from the lines that
// start inside the chunk; the summaries are folded 64 to 1 on the device for both classes a group may
// start in, the host resolves classes and prefix sums over those few thousand group summaries
// (fasta_resolve_groups) and the device expands them back to one state per chunk; pass 2
// (fasta_emit_kernel) walks the chunk again
// and stores every kept letter at  guard + letters before + record index  -- the concat layout
// of bsgpu_subject with one separator after every record -- and the start of every record
// (16 letters per store wherever a 16-byte block of the output belongs to one thread).
// The walk itself (fasta_walk_chunk) is plain C++ shared by both kernels and by the host-side
// check in tests/hostcheck/, which runs it on the CPU against the reference's parser.
#pragma once
#include <stdint.h>
#include <vector>

#ifdef __CUDACC__
#define BSGPU_HD __host__ __device__ __forceinline__
#else
#define BSGPU_HD inline
#endif

#define BSGPU_FASTA_CHUNK 512           // bytes per thread (multiple of 16)
#define BSGPU_FASTA_GUARD 128           // = BSGPU_GUARD: letters in front of the first record

enum { BSGPU_FASTA_SEQ = 0, BSGPU_FASTA_SKIP = 1, BSGPU_FASTA_PASS = 2 };

struct BsgpuFastaSummary {
	int32_t kept_head, invalid_head;        // first partial line, if it is sequence data
	int32_t kept_rest, invalid_rest;        // lines that start in this chunk
	int32_t headers;                        // description lines that start in this chunk
	int32_t class_out;                      // class of the line the chunk ends in; PASS = no line starts here
	int32_t first_header;                   // offset in the chunk of its first description line, or -1
	int32_t first_seqline;                  // offset of the first non-empty sequence line starting here, or -1
};

struct BsgpuFastaResolved {
	int64_t kept_before;                    // letters kept in front of this chunk
	int32_t headers_before;                 // records opened in front of this chunk
	int32_t class_in;                       // class of the line the chunk starts in the middle of
};

// 16 text bytes at offset i (multiple of 16; the buffer is padded to a multiple of 16)
BSGPU_HD void bsgpu_fasta_load16(const uint8_t *text, int64_t i, uint32_t w[4])
{
#ifdef __CUDA_ARCH__
	uint4 v = __ldg(reinterpret_cast<const uint4 *>(text + i));

	w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
#else
	for (int k = 0; k < 4; k++)
		w[k] = (uint32_t) text[i + 4 * k] | ((uint32_t) text[i + 4 * k + 1] << 8) |
		       ((uint32_t) text[i + 4 * k + 2] << 16) | ((uint32_t) text[i + 4 * k + 3] << 24);
#endif
}

// Output staging of the emit pass: a thread's letters go to consecutive positions, so they are
// collected per 16-byte block of the output and a block this thread filled completely is
// written with ONE 16-byte store (byte stores cost a partial-sector write in L2 each); blocks
// shared with a neighbouring thread are written byte by byte, only the bytes that are ours.
struct BsgpuFastaOut {
	uint8_t *out;
	int64_t blk;            // 16-aligned position of the block being filled, -1 = none
	uint64_t lo, hi;        // its bytes 0..7 / 8..15
	uint32_t mask;          // which of them are set
};

BSGPU_HD void fasta_out_flush(BsgpuFastaOut &o)
{
	if (o.blk < 0 || o.mask == 0)
		return;
	if (o.mask == 0xFFFFu) {
#ifdef __CUDA_ARCH__
		*reinterpret_cast<uint4 *>(o.out + o.blk) =
			make_uint4((uint32_t) o.lo, (uint32_t) (o.lo >> 32), (uint32_t) o.hi, (uint32_t) (o.hi >> 32));
#else
		for (int k = 0; k < 16; k++)
			o.out[o.blk + k] = (uint8_t) ((k < 8 ? o.lo >> (8 * k) : o.hi >> (8 * (k - 8))) & 0xFF);
#endif
	} else {
		for (int k = 0; k < 16; k++)
			if (o.mask & (1u << k))
				o.out[o.blk + k] = (uint8_t) ((k < 8 ? o.lo >> (8 * k) : o.hi >> (8 * (k - 8))) & 0xFF);
	}
	o.mask = 0;
}

BSGPU_HD void fasta_out_put(BsgpuFastaOut &o, int64_t pos, uint32_t byte)
{
	int64_t b = pos & ~(int64_t) 15;
	int k = (int) (pos & 15);

	if (b != o.blk) {
		fasta_out_flush(o);
		o.blk = b;
		o.lo = o.hi = 0;
		o.mask = 0;
	}
	if (k < 8)
		o.lo |= (uint64_t) byte << (8 * k);
	else
		o.hi |= (uint64_t) byte << (8 * (k - 8));
	o.mask |= 1u << k;
}

// Walks chunk [a, b) of text[0, n).  EMIT = false: fills *sum.  EMIT = true: `in` holds the
// resolved state in front of the chunk; kept letters go to out[], record starts to seg_start[]
// and the offsets of the description texts (the byte after '>') to desc_off[].
// lkup: 256 entries, < 0 = not a letter of the alphabet.
template <bool EMIT>
BSGPU_HD void fasta_walk_chunk(const uint8_t *text, int64_t n, int64_t a, int64_t b, const int16_t *lkup,
			       BsgpuFastaSummary *sum, BsgpuFastaResolved in, uint8_t *out,
			       int64_t *seg_start, int64_t *desc_off)
{
	bool at_line_start = a == 0 || text[a - 1] == '\n';
	bool in_head = !at_line_start, saw_line_start = false;
	int cls = at_line_start || !EMIT ? BSGPU_FASTA_SEQ : in.class_in;
	int32_t kept_head = 0, invalid_head = 0, kept_rest = 0, invalid_rest = 0, headers = 0;
	int32_t first_header = -1, first_seqline = -1;
	int64_t kept = in.kept_before;
	int64_t hdrs = in.headers_before;
	BsgpuFastaOut o = {out, -1, 0, 0, 0};

	for (int64_t i0 = a; i0 < b; i0 += 16) {
		uint32_t w[4];

		bsgpu_fasta_load16(text, i0, w);
#pragma unroll
		for (int k = 0; k < 16; k++) {
			int64_t i = i0 + k;

			if (i >= b)
				break;
			uint32_t c = (w[k >> 2] >> (8 * (k & 3))) & 0xFFu;
			bool cr_of_crlf = c == '\r' && i + 1 < n && text[i + 1] == '\n';

			if (at_line_start) {
				in_head = false;
				saw_line_start = true;
				if (c == '>') {
					cls = BSGPU_FASTA_SKIP;
					headers++;
					if (first_header < 0)
						first_header = (int32_t) (i - a);
					if (EMIT) {
						// the separator behind the previous record is ours too: writing it
						// keeps the blocks around record boundaries whole
						if (hdrs > 0)
							fasta_out_put(o, BSGPU_FASTA_GUARD + kept + hdrs - 1, 0x80u);
						seg_start[hdrs] = BSGPU_FASTA_GUARD + kept + hdrs;
						desc_off[hdrs] = i + 1;
						hdrs++;
					}
				} else if (c == ';') {
					cls = BSGPU_FASTA_SKIP;
				} else {
					cls = BSGPU_FASTA_SEQ;
					if (c != '\n' && !cr_of_crlf && first_seqline < 0)
						first_seqline = (int32_t) (i - a);
				}
			}
			at_line_start = c == '\n';
			if (c == '\n' || cls != BSGPU_FASTA_SEQ || cr_of_crlf)
				continue;
			int code = lkup[c];
			if (code < 0) {
				if (in_head)
					invalid_head++;
				else
					invalid_rest++;
				continue;
			}
			if (EMIT) {
				if (hdrs > 0)
					fasta_out_put(o, BSGPU_FASTA_GUARD + kept + hdrs - 1, (uint32_t) code);
				kept++;
			} else if (in_head) {
				kept_head++;
			} else {
				kept_rest++;
			}
		}
	}
	if (EMIT)
		fasta_out_flush(o);
	if (!EMIT) {
		sum->kept_head = kept_head;
		sum->invalid_head = invalid_head;
		sum->kept_rest = kept_rest;
		sum->invalid_rest = invalid_rest;
		sum->headers = headers;
		sum->class_out = saw_line_start ? cls : BSGPU_FASTA_PASS;
		sum->first_header = first_header;
		sum->first_seqline = first_seqline;
	}
}

// ---- between the passes: line classes across chunk boundaries, prefix sums ---------------------
// The chunk summaries stay on the device.  BSGPU_FASTA_GROUP consecutive chunks are folded into one
// group summary for BOTH classes the group may start in (fasta_group_kernel, one thread per group),
// the host walks the few thousand group summaries (fasta_resolve_groups), and
// fasta_expand_kernel replays every group from its resolved base to give each chunk its state.

#define BSGPU_FASTA_GROUP 64

struct BsgpuFastaGroup {
	int64_t kept[2], invalid[2];    // indexed by the class the group starts in (SEQ, SKIP)
	int32_t class_out[2];           // class of the line the group ends in, given that class
	int64_t headers;
	int64_t first_header;           // byte offset of the first description line in the group, or -1
	int64_t first_seqline;          // byte offset of the first non-empty sequence line starting in it, or -1
};

BSGPU_HD void fasta_fold_group(const BsgpuFastaSummary *sums, int64_t c0, int64_t c1, BsgpuFastaGroup *g)
{
	for (int h = 0; h < 2; h++) {
		int cls = h;
		int64_t kept = 0, invalid = 0;

		for (int64_t c = c0; c < c1; c++) {
			const BsgpuFastaSummary s = sums[c];

			if (cls == BSGPU_FASTA_SEQ) {   // the first partial line of the chunk is sequence data
				kept += s.kept_head;
				invalid += s.invalid_head;
			}
			kept += s.kept_rest;
			invalid += s.invalid_rest;
			if (s.class_out != BSGPU_FASTA_PASS)
				cls = s.class_out;
		}
		g->kept[h] = kept;
		g->invalid[h] = invalid;
		g->class_out[h] = cls;
	}
	int64_t headers = 0, first_header = -1, first_seqline = -1;

	for (int64_t c = c0; c < c1; c++) {
		const BsgpuFastaSummary s = sums[c];

		headers += s.headers;
		if (first_header < 0 && s.first_header >= 0)
			first_header = c * BSGPU_FASTA_CHUNK + s.first_header;
		if (first_seqline < 0 && s.first_seqline >= 0)
			first_seqline = c * BSGPU_FASTA_CHUNK + s.first_seqline;
	}
	g->headers = headers;
	g->first_header = first_header;
	g->first_seqline = first_seqline;
}

// state of every chunk of one group, replayed from the group's resolved base
BSGPU_HD void fasta_expand_group(const BsgpuFastaSummary *sums, int64_t c0, int64_t c1,
				 BsgpuFastaResolved base, BsgpuFastaResolved *res)
{
	int cls = base.class_in;
	int64_t kept = base.kept_before;
	int64_t headers = base.headers_before;

	for (int64_t c = c0; c < c1; c++) {
		const BsgpuFastaSummary s = sums[c];

		res[c].kept_before = kept;
		res[c].headers_before = (int32_t) headers;
		res[c].class_in = cls;
		if (cls == BSGPU_FASTA_SEQ)
			kept += s.kept_head;
		kept += s.kept_rest;
		headers += s.headers;
		if (s.class_out != BSGPU_FASTA_PASS)
			cls = s.class_out;
	}
}

struct BsgpuFastaTotals {
	int64_t kept = 0, invalid = 0, nrec = 0;
	int64_t first_header = -1;      // byte offset of the first description line
	int64_t first_seqline = -1;     // byte offset of the first non-empty sequence line
};

// Host step: one iteration per group.  bases[g] = state in front of group g.
static inline BsgpuFastaTotals fasta_resolve_groups(const std::vector<BsgpuFastaGroup> &groups,
						    std::vector<BsgpuFastaResolved> &bases)
{
	BsgpuFastaTotals t;
	int cls = BSGPU_FASTA_SEQ;      // the text starts at a line start: the value is never used

	bases.resize(groups.size());
	for (size_t g = 0; g < groups.size(); g++) {
		const BsgpuFastaGroup &G = groups[g];

		bases[g].kept_before = t.kept;
		bases[g].headers_before = (int32_t) t.nrec;
		bases[g].class_in = cls;
		t.kept += G.kept[cls];
		t.invalid += G.invalid[cls];
		t.nrec += G.headers;
		if (t.first_header < 0 && G.first_header >= 0)
			t.first_header = G.first_header;
		if (t.first_seqline < 0 && G.first_seqline >= 0)
			t.first_seqline = G.first_seqline;
		cls = G.class_out[cls];
	}
	return t;
}

#ifdef __CUDACC__
// Letters -> codes through the lookup, 16 per thread (one 16-byte load and store); the smallest
// index of a letter the lookup does not know lands in *first_bad.  src is padded to a multiple
// of 16 bytes, dst + 16 * k is 16-byte aligned (the subject's guard is 128 letters); the last,
// partial group is stored byte by byte so that the guard behind the sequence stays intact.
__global__ void __launch_bounds__(256)
encode_letters_kernel(const uint8_t *__restrict__ src, int64_t n, const int16_t *__restrict__ lkup,
		      uint8_t *__restrict__ dst, unsigned long long *__restrict__ first_bad)
{
	int64_t ngroups = (n + 15) / 16;
	int64_t stride = (int64_t) gridDim.x * blockDim.x;

	for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
		uint32_t w[4], o[4] = {0, 0, 0, 0};
		int64_t i0 = g * 16;

		bsgpu_fasta_load16(src, i0, w);
#pragma unroll
		for (int k = 0; k < 16; k++) {
			if (i0 + k >= n)
				break;
			int code = lkup[(w[k >> 2] >> (8 * (k & 3))) & 0xFFu];

			if (code < 0) {
				atomicMin(first_bad, (unsigned long long) (i0 + k));
				code = 0;
			}
			o[k >> 2] |= (uint32_t) code << (8 * (k & 3));
		}
		if (i0 + 16 <= n) {
			*reinterpret_cast<uint4 *>(dst + i0) = make_uint4(o[0], o[1], o[2], o[3]);
		} else {
			for (int k = 0; i0 + k < n; k++)
				dst[i0 + k] = (uint8_t) ((o[k >> 2] >> (8 * (k & 3))) & 0xFFu);
		}
	}
}

__global__ void __launch_bounds__(128)
fasta_summary_kernel(const uint8_t *__restrict__ text, int64_t n, int64_t nchunks,
		     const int16_t *__restrict__ lkup, BsgpuFastaSummary *__restrict__ sums)
{
	int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;

	if (c >= nchunks)
		return;
	int64_t a = c * BSGPU_FASTA_CHUNK, b = a + BSGPU_FASTA_CHUNK < n ? a + BSGPU_FASTA_CHUNK : n;
	BsgpuFastaResolved none = {0, 0, BSGPU_FASTA_SEQ};

	fasta_walk_chunk<false>(text, n, a, b, lkup, sums + c, none, nullptr, nullptr, nullptr);
}

__global__ void __launch_bounds__(128)
fasta_group_kernel(const BsgpuFastaSummary *__restrict__ sums, int64_t nchunks, int64_t ngroups,
		   BsgpuFastaGroup *__restrict__ groups)
{
	int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;

	if (g >= ngroups)
		return;
	int64_t c0 = g * BSGPU_FASTA_GROUP, c1 = c0 + BSGPU_FASTA_GROUP < nchunks ? c0 + BSGPU_FASTA_GROUP : nchunks;

	fasta_fold_group(sums, c0, c1, groups + g);
}

__global__ void __launch_bounds__(128)
fasta_expand_kernel(const BsgpuFastaSummary *__restrict__ sums, int64_t nchunks, int64_t ngroups,
		    const BsgpuFastaResolved *__restrict__ bases, BsgpuFastaResolved *__restrict__ res)
{
	int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;

	if (g >= ngroups)
		return;
	int64_t c0 = g * BSGPU_FASTA_GROUP, c1 = c0 + BSGPU_FASTA_GROUP < nchunks ? c0 + BSGPU_FASTA_GROUP : nchunks;

	fasta_expand_group(sums, c0, c1, bases[g], res);
}

__global__ void __launch_bounds__(128)
fasta_emit_kernel(const uint8_t *__restrict__ text, int64_t n, int64_t nchunks,
		  const int16_t *__restrict__ lkup, const BsgpuFastaResolved *__restrict__ res,
		  uint8_t *__restrict__ out, int64_t *__restrict__ seg_start, int64_t *__restrict__ desc_off)
{
	int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;

	if (c >= nchunks)
		return;
	int64_t a = c * BSGPU_FASTA_CHUNK, b = a + BSGPU_FASTA_CHUNK < n ? a + BSGPU_FASTA_CHUNK : n;

	fasta_walk_chunk<true>(text, n, a, b, lkup, nullptr, res[c], out, seg_start, desc_off);
}
#endif
