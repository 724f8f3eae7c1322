// bsgpu.cu -- host side of libbsgpu.so: the C ABI declared in include/bsgpu.h.
//
// Plays the role of the reference's C layer under the .Call boundary
// (src/match_pdict.c, src/match_pdict_Twobit.c, src/match_pdict_ACtree2.c,
// src/match_pdict_utils.c, src/match_reporting.c, src/MIndex_class.c) with the per-base
// work done by the kernels in bsgpu_kernels.cuh.  No CPU fallback: every compute entry
// point needs a CUDA device.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>
#include <thrust/iterator/transform_iterator.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/bsgpu.h"
#include "bsgpu_qindex.cuh"
#include "bsgpu_xseed.cuh"
#include "bsgpu_tbdir.cuh"
#include "bsgpu_oligo.cuh"
#include "bsgpu_fasta.cuh"
#include "bsgpu_indels.cuh"

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------

static thread_local char g_err[1024] = "";
static long long g_launches = 0;

static int fail(int code, const char *fmt, ...)
{
	va_list ap;

	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}

#define CUDA_TRY(expr) do { \
	cudaError_t e_ = (expr); \
	if (e_ != cudaSuccess) \
		return fail(e_ == cudaErrorMemoryAllocation ? BSGPU_ENOMEM : BSGPU_ECUDA, \
			    "CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, \
			    __LINE__, cudaGetErrorString(e_)); \
} while (0)

#define TRY(expr) do { int rc_ = (expr); if (rc_ != BSGPU_OK) return rc_; } while (0)

extern "C" const char *bsgpu_last_error(void) { return g_err; }
extern "C" const char *bsgpu_version(void) { return "bsgpu 0.1 (sm_100a)"; }
extern "C" int64_t bsgpu_launch_count(void) { return g_launches; }

static char g_plan[512] = "";
static void set_plan(const char *fmt, ...)
{
	va_list ap;

	va_start(ap, fmt);
	vsnprintf(g_plan, sizeof(g_plan), fmt, ap);
	va_end(ap);
}
extern "C" const char *bsgpu_last_plan(void) { return g_plan; }

// Opt-in timing of the hot kernels with CUDA events on the launching stream
// (bsgpu_kernel_timing / bsgpu_kernel_times; bench.py uses it for the roofline).
struct KernelTime {
	const char *name;
	cudaEvent_t start, stop;
};
static bool g_timing = false;
static std::vector<KernelTime> g_times;

struct KernelTimer {
	cudaStream_t st;
	cudaEvent_t stop = nullptr;

	KernelTimer(const char *name, cudaStream_t st_) : st(st_)
	{
		if (!g_timing || g_times.size() >= 8192)
			return;
		KernelTime kt = {name, nullptr, nullptr};
		if (cudaEventCreate(&kt.start) != cudaSuccess || cudaEventCreate(&kt.stop) != cudaSuccess)
			return;
		cudaEventRecord(kt.start, st);
		stop = kt.stop;
		g_times.push_back(kt);
	}
	~KernelTimer()
	{
		if (stop != nullptr)
			cudaEventRecord(stop, st);
	}
};

static void clear_kernel_times(void)
{
	for (auto &kt : g_times) {
		cudaEventDestroy(kt.start);
		cudaEventDestroy(kt.stop);
	}
	g_times.clear();
}

extern "C" int bsgpu_kernel_timing(int enable)
{
	clear_kernel_times();
	g_timing = enable != 0;
	return BSGPU_OK;
}

extern "C" int bsgpu_kernel_times(const char **names, float *ms, int max_entries)
{
	int n = 0;

	for (auto &kt : g_times) {
		if (n >= max_entries)
			break;
		if (cudaEventSynchronize(kt.stop) != cudaSuccess ||
		    cudaEventElapsedTime(&ms[n], kt.start, kt.stop) != cudaSuccess)
			return fail(BSGPU_ECUDA, "bsgpu_kernel_times(): event query failed"), -1;
		names[n++] = kt.name;
	}
	clear_kernel_times();
	return n;
}

// ---------------------------------------------------------------------------------------
// device selection
// ---------------------------------------------------------------------------------------

static int g_device = -1;
static int g_num_sms = 148;

extern "C" int bsgpu_device_count(void)
{
	int n = 0;

	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

static bool device_state_exists(void);

extern "C" int bsgpu_set_device(int device)
{
	int n = bsgpu_device_count();

	if (n == 0)
		return fail(BSGPU_ENODEVICE, "no CUDA device available: libbsgpu has no CPU fallback");
	if (device < 0 || device >= n)
		return fail(BSGPU_EINVAL, "invalid CUDA device %d (have %d)", device, n);
	if (g_device >= 0 && device != g_device && device_state_exists())
		return fail(BSGPU_EINVAL, "bsgpu_set_device(%d): the library already holds buffers and indexes on device "
			    "%d; select the device before the first call (one process per GPU)", device, g_device);
	CUDA_TRY(cudaSetDevice(device));
	cudaDeviceProp prop;
	CUDA_TRY(cudaGetDeviceProperties(&prop, device));
	g_num_sms = prop.multiProcessorCount;
	g_device = device;
	return BSGPU_OK;
}

static int ensure_device(void)
{
	if (g_device >= 0) {
		CUDA_TRY(cudaSetDevice(g_device));
		return BSGPU_OK;
	}
	const char *env = getenv("BSGPU_DEVICE");
	return bsgpu_set_device(env ? atoi(env) : 0);
}

static int grid_for(long long nthreads, int block, int per_sm)
{
	long long blocks = (nthreads + block - 1) / block;
	long long cap = (long long) g_num_sms * per_sm;

	if (blocks > cap)
		blocks = cap;
	if (blocks < 1)
		blocks = 1;
	return (int) blocks;
}

// Device memory pool: cudaMalloc/cudaFree cost milliseconds for the buffers of a 250 Mbp
// subject, so freed blocks are kept (up to BSGPU_POOL_BYTES, default 8 GiB) and reused
// by later calls of the same process.
struct PoolBlock { void *p; size_t bytes; };
static std::vector<PoolBlock> g_pool;
static size_t g_pool_bytes = 0;

static size_t pool_limit(void)
{
	static size_t lim = 0;

	if (lim == 0) {
		const char *env = getenv("BSGPU_POOL_BYTES");
		lim = env ? (size_t) strtoull(env, nullptr, 10) : ((size_t) 8 << 30);
		if (lim == 0)
			lim = 1;
	}
	return lim;
}

static void pool_flush(void)
{
	for (auto &b : g_pool)
		cudaFree(b.p);
	g_pool.clear();
	g_pool_bytes = 0;
}

static void *pool_alloc(size_t bytes, size_t *got)
{
	int best = -1;

	bytes = (bytes + 511) & ~(size_t) 511;
	for (int i = 0; i < (int) g_pool.size(); i++)
		if (g_pool[i].bytes >= bytes && g_pool[i].bytes <= 2 * bytes + 4096 &&
		    (best < 0 || g_pool[i].bytes < g_pool[best].bytes))
			best = i;
	if (best >= 0) {
		void *p = g_pool[best].p;

		*got = g_pool[best].bytes;
		g_pool_bytes -= g_pool[best].bytes;
		g_pool.erase(g_pool.begin() + best);
		return p;
	}
	void *p = nullptr;
	cudaError_t e = cudaMalloc(&p, bytes);
	if (e != cudaSuccess) {
		cudaGetLastError();
		pool_flush();
		e = cudaMalloc(&p, bytes);
		if (e != cudaSuccess) {
			cudaGetLastError();
			return nullptr;
		}
	}
	*got = bytes;
	return p;
}

static void pool_free(void *p, size_t bytes)
{
	if (p == nullptr)
		return;
	if (bytes > pool_limit() / 2 || g_pool.size() >= 256) {
		cudaFree(p);
		return;
	}
	while (g_pool_bytes + bytes > pool_limit() && !g_pool.empty()) {
		cudaFree(g_pool.front().p);
		g_pool_bytes -= g_pool.front().bytes;
		g_pool.erase(g_pool.begin());
	}
	g_pool.push_back({p, bytes});
	g_pool_bytes += bytes;
}

static long g_live_allocs = 0;           // device blocks currently held by DevBuf objects

template <typename T>
struct DevBuf {
	T *p = nullptr;
	size_t n = 0;           // usable elements
	size_t bytes = 0;       // size of the underlying block

	DevBuf() {}
	DevBuf(const DevBuf &) = delete;
	DevBuf &operator=(const DevBuf &) = delete;
	~DevBuf() { release(); }
	void release()
	{
		if (p != nullptr)
			g_live_allocs--;
		pool_free(p, bytes);
		p = nullptr;
		n = 0;
		bytes = 0;
	}
	void swap(DevBuf &o) { std::swap(p, o.p); std::swap(n, o.n); std::swap(bytes, o.bytes); }
	int alloc(size_t count)
	{
		release();
		if (count == 0)
			count = 1;
		p = (T *) pool_alloc(count * sizeof(T), &bytes);
		if (p == nullptr) {
			bytes = 0;
			return fail(BSGPU_ENOMEM, "cudaMalloc of %zu bytes failed", count * sizeof(T));
		}
		g_live_allocs++;
		n = count;
		return BSGPU_OK;
	}
	int upload(const T *src, size_t count)
	{
		TRY(alloc(count));
		if (count)
			CUDA_TRY(cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice));
		return BSGPU_OK;
	}
};

// ---------------------------------------------------------------------------------------
// PDict3Parts
// ---------------------------------------------------------------------------------------

struct bsgpu_pdict3parts {
	int algo = 0, M = 0, w = 0, seed_w = 0;
	bool has_head = false, has_tail = false;
	int max_H = 0, max_T = 0;
	std::vector<int32_t> high2low;          // NA or 1-based
	std::vector<uint64_t> keys;             // seed key per pattern
	std::vector<uint8_t> tb;                // M*w letters
	std::vector<int32_t> tail_len;          // per pattern
	std::vector<uint8_t> h_head, h_tail;    // host copies of the flank letters
	std::vector<int64_t> h_head_off, h_tail_off;
	std::vector<uint8_t> excluded;          // pp_exclude[i] != NA
	std::vector<uint8_t> is_tb_dup;         // had a non-NA high2low at build time
	mutable std::unique_ptr<struct QIndex> qcache;  // q-gram index built on first use
	mutable std::unique_ptr<struct ByteSeedIndex> bcache;   // byte-seed index (exact, no planes)
	mutable std::unique_ptr<struct XSeedIndex> xcache;      // extended-seed index (exact, w >= 25)
	mutable std::unique_ptr<struct TbDirIndex> tcache;      // band directory (w <= 12, packed flanks)
	// non-preprocessed dictionary (bsgpu_patterns_build): the letters only, no Trusted Band
	bool plain = false;
	std::vector<uint8_t> h_pat;             // host copy of the letters (for the automatic index below)
	std::vector<int64_t> h_pat_off;
	// a large base-only non-preprocessed dictionary matched exactly gets a Trusted-Band index of its
	// own on first use (plain_auto_index): same results as one matchPattern() per pattern
	mutable bsgpu_pdict3parts *auto_idx = nullptr;
	mutable int auto_state = 0;             // 0 = not tried, 1 = built, -1 = not eligible
	~bsgpu_pdict3parts();
	DevBuf<uint8_t> d_pat;
	DevBuf<int64_t> d_pat_off;
	DevBuf<uint4> d_pat_head;
	int32_t max_plen = 0;
	uint32_t nbuckets = 0;
	bool all_singletons = true;
	DevBuf<uint4> d_table;
	DevBuf<uint64_t> d_keys;
	DevBuf<int32_t> d_seed_next, d_leaders;
	int32_t n_leaders = 0;
	DevBuf<int32_t> d_grp_off, d_grp_keys;
	DevBuf<uint8_t> d_tb, d_head, d_tail;
	DevBuf<int64_t> d_head_off, d_tail_off;
	DevBuf<ulonglong2> d_flank2;            // packed head / tail of every key (XOR + popcount check)
	DevBuf<uint16_t> d_flank_len;
	bool flanks_all_packed = false;         // every non-excluded key: base-only flanks of <= 32 letters

	BsgpuIndexView view() const
	{
		BsgpuIndexView I;

		I.table = d_table.p;
		I.keys = d_keys.p;
		I.seed_next = d_seed_next.p;
		int lg = 0;
		while ((1u << lg) < nbuckets)
			lg++;
		I.bucket_shift = 32 - lg;
		I.bucket_mask = nbuckets - 1;
		I.w = w;
		I.seed_w = seed_w;
		I.seed_mask = seed_w >= 32 ? ~0ull : ((1ull << (2 * seed_w)) - 1);
		I.M = M;
		I.grp_off = d_grp_off.p;
		I.grp_keys = d_grp_keys.p;
		I.tb_bytes = d_tb.p;
		I.head = has_head ? d_head.p : nullptr;
		I.head_off = d_head_off.p;
		I.tail = has_tail ? d_tail.p : nullptr;
		I.tail_off = d_tail_off.p;
		I.all_singletons = all_singletons;
		I.flank2 = d_flank2.p;
		I.flank_len = d_flank_len.p;
		return I;
	}
};

// ---------------------------------------------------------------------------------------
// q-gram seed index (bsgpu_qindex.cuh)
// ---------------------------------------------------------------------------------------

struct SeedShape { int m, k, q, span; unsigned long long mask, shifts; };
static const SeedShape g_shapes[] = {
#include "bsgpu_shapes.inc"
};

struct QIndex {
	bool ok = false;        // false: this dictionary is not eligible (cached negative answer)
	int nparts = 0;
	BsgpuQIndexView v = {};
	DevBuf<uint2> d_dir;
	DevBuf<uint32_t> d_entries, d_bitmap;
	DevBuf<uint2> d_cells2;                 // qscan2_kernel: rank directory
	DevBuf<unsigned long long> d_entries2;  // ... and its 8-byte slots
	size_t nent = 0;
	DevBuf<ulonglong2> d_pat;
	DevBuf<unsigned long long> d_pat8;
	DevBuf<uint8_t> d_len;
	int max_len = 0;
	int max_shift = 0;
};

static bool qindex_disabled(void)
{
	const char *env = getenv("BSGPU_NO_QINDEX");
	return env != nullptr && env[0] == '1';
}

// Picks the seed family for a Hamming filter on windows of m letters with <= k mismatches.
static bool pick_shape(int m, int k, SeedShape *out)
{
	if (m >= 12 * (k + 1)) {
		out->m = m; out->k = k; out->q = 12; out->span = 12; out->mask = 0xFFFull;
		out->shifts = 0;
		for (int i = 0; i <= k; i++)
			out->shifts |= 1ull << (12 * i);
		return true;
	}
	for (const SeedShape &sh : g_shapes)
		if (sh.m == m && sh.k == k) {
			*out = sh;
			return true;
		}
	return false;
}

// Builds (once) the q-gram index of the dictionary behind `parts`.
//   nparts == 1: parts[0] must be a whole-pattern Trusted Band (no head, no tail)
//   nparts  > 1: the components of an MTB_PDict (band b = letters [headw_b, headw_b+w_b))
// Returns the cached index; ->ok tells whether the dictionary is eligible.
static int bits_for(unsigned long long v);

static int get_qindex(const bsgpu_pdict3parts *const *parts, int nparts, const QIndex **out)
{
	const bsgpu_pdict3parts *p0 = parts[0];

	*out = nullptr;
	if (p0->qcache && p0->qcache->nparts == nparts) {
		*out = p0->qcache.get();
		return BSGPU_OK;
	}
	std::unique_ptr<QIndex> qi(new QIndex());
	qi->nparts = nparts;
	int M = p0->M;
	auto done = [&]() {
		p0->qcache = std::move(qi);
		*out = p0->qcache.get();
		return BSGPU_OK;
	};
	if (qindex_disabled() || M == 0 || M >= (1 << 27) || p0->plain)
		return done();
	// --- reconstruct the full patterns: head0 + tb0 + tail0 ------------------------------
	std::vector<int32_t> len(M);
	std::vector<const uint8_t *> hp(M, nullptr), tp(M, nullptr);
	std::vector<int32_t> hl(M, 0), tl(M, 0);
	int min_w = 1 << 30, max_w = 0;
	for (int i = 0; i < M; i++) {
		if (p0->has_head) {
			hp[i] = p0->h_head.data() + p0->h_head_off[i];
			hl[i] = (int32_t) (p0->h_head_off[i + 1] - p0->h_head_off[i]);
		}
		if (p0->has_tail) {
			tp[i] = p0->h_tail.data() + p0->h_tail_off[i];
			tl[i] = (int32_t) (p0->h_tail_off[i + 1] - p0->h_tail_off[i]);
		}
		len[i] = hl[i] + p0->w + tl[i];
		if (p0->excluded[i])
			continue;
		min_w = std::min(min_w, len[i]);
		max_w = std::max(max_w, len[i]);
	}
	if (max_w == 0 || max_w > 64)
		return done();
	SeedShape sh;
	if (nparts == 1) {
		// Exact dictionaries: the hash scan (one probe per window) is faster than sampled
		// seeds at 1M patterns on B200 (0.95 vs 1.5 ms per 250 Mbp), so this mode is
		// opt-in for experiments.
		const char *env_x = getenv("BSGPU_Q_EXACT");
		if (env_x == nullptr || env_x[0] != '1')
			return done();
		if (p0->has_head || p0->has_tail || !p0->all_singletons || p0->w < 20)
			return done();
		sh.m = p0->w; sh.k = 0; sh.q = 12; sh.span = 12; sh.mask = 0xFFFull; sh.shifts = 0;
		int stride = std::min(p0->w - 12 + 1, BSGPU_Q_MAX_SHIFTS);
		const char *env_s = getenv("BSGPU_Q_STRIDE");
		if (env_s != nullptr && atoi(env_s) >= 1)
			stride = std::min(stride, atoi(env_s));
		for (int t = 0; t < stride; t++)
			sh.shifts |= 1ull << t;
		qi->v.mode = 0;
		qi->v.stride = stride;
	} else {
		// the parts must be the k+1 bands of one MTB_PDict: band b starts where band b-1
		// ends, heads of band b = everything before it (R/PDict-class.R:476-484)
		int at = 0;
		for (int b = 0; b < nparts; b++) {
			const bsgpu_pdict3parts *pb = parts[b];
			if (pb->M != M || pb->excluded != p0->excluded)
				return done();
			if ((b == 0) != !pb->has_head)
				return done();
			if (b > 0)
				for (int i = 0; i < M; i++)
					if (pb->h_head_off[i + 1] - pb->h_head_off[i] != at)
						return done();
			at += pb->w;
		}
		if (at != min_w)
			return done();
		if (!pick_shape(min_w, nparts - 1, &sh))
			return done();
		qi->v.mode = 1;
		qi->v.stride = 1;
	}
	if (sh.span > 32)
		return done();
	// --- all letters of the non-excluded patterns must be base letters ------------------------
	std::vector<ulonglong2> pat(M, make_ulonglong2(0, 0));
	std::vector<uint8_t> plen(M, 0);
	std::vector<uint8_t> letters(64);
	for (int i = 0; i < M; i++) {
		if (p0->excluded[i])
			continue;
		int n = 0;
		for (int d = 0; d < hl[i]; d++) letters[n++] = hp[i][d];
		for (int d = 0; d < p0->w; d++) letters[n++] = p0->tb[(size_t) i * p0->w + d];
		for (int d = 0; d < tl[i]; d++) letters[n++] = tp[i][d];
		unsigned long long lo = 0, hi = 0;
		for (int d = 0; d < n; d++) {
			if (!bsgpu_is_base(letters[d]))
				return done();
			unsigned long long c = bsgpu_base_code(letters[d]);
			if (d < 32) lo |= c << (2 * d); else hi |= c << (2 * (d - 32));
		}
		pat[i] = make_ulonglong2(lo, hi);
		plen[i] = (uint8_t) n;
	}
	// --- seed family ---------------------------------------------------------------------------
	BsgpuQIndexView &v = qi->v;
	v.q = sh.q;
	v.min_w = nparts == 1 ? p0->w : min_w;
	v.shape1 = sh.mask;
	v.shape2 = 0;
	v.nruns = 0;
	for (int i = 0; i < sh.span; i++)
		if ((sh.mask >> i) & 1) {
			v.shape2 |= 1ull << (2 * i);
			if (i > 0 && ((sh.mask >> (i - 1)) & 1)) {
				v.run_len[v.nruns - 1]++;
			} else {
				if (v.nruns == BSGPU_Q_MAX_RUNS)
					return done();
				v.run_src[v.nruns] = i;
				v.run_len[v.nruns] = 1;
				v.nruns++;
			}
		}
	for (int i = 0, at = 0; i < BSGPU_Q_MAX_RUNS; i++) {
		bool used = i < v.nruns;

		v.run_shift[i] = used ? 2 * v.run_src[i] : 0;
		v.run_mask[i] = used ? (uint32_t) ((1ull << (2 * v.run_len[i])) - 1) : 0u;
		v.run_mul[i] = used ? 1u << at : 0u;
		if (used)
			at += 2 * v.run_len[i];
	}
	v.nshifts = 0;
	for (int t = 0; t < 64; t++)
		if ((sh.shifts >> t) & 1) {
			if (v.nshifts == BSGPU_Q_MAX_SHIFTS || t + sh.span > v.min_w)
				return done();
			v.shift[v.nshifts++] = t;
			qi->max_shift = t;
		}
	// --- entries, counting sort by seed key ---------------------------------------------------
	v.key_bits = 2 * sh.q;          // the key is the q shape letters themselves
	if (v.key_bits > 24 || v.key_bits < 8)
		return done();
	v.mask2 = v.shape2 | (v.shape2 << 1);
	v.const_len = min_w == max_w ? max_w : 0;
	size_t nkeys = (size_t) 1 << v.key_bits;
	std::vector<uint32_t> cnt(nkeys + 1, 0);
	auto seed_key = [&](int i, int t) {
		// the window of 32 letters starting at pattern offset t, in the subject's layout
		unsigned long long win = t == 0 ? pat[i].x
			: (t < 32 ? (pat[i].x >> (2 * t)) | (pat[i].y << (64 - 2 * t)) : pat[i].y >> (2 * (t - 32)));
		return q_key(v, win);
	};
	size_t nent = 0;
	for (int i = 0; i < M; i++) {
		if (p0->excluded[i])
			continue;
		if (nparts == 1 && p0->is_tb_dup[i])
			continue;       // TB duplicates are never reported once dups are blanked
		for (int s2 = 0; s2 < v.nshifts; s2++) {
			cnt[seed_key(i, v.shift[s2])]++;
			nent++;
		}
	}
	for (size_t k2 = 0; k2 < nkeys; k2++)
		if (cnt[k2] > 254)
			return done();  // a very repetitive dictionary: keep the Trusted-Band engine
	v.shift_bits = bits_for((unsigned long long) std::max(v.nshifts - 1, 1));
	v.id_bits = bits_for((unsigned long long) std::max(M - 1, 1));
	if (v.shift_bits + v.id_bits > 32)
		return done();
	// --- rank directory + inline patterns (qscan2_kernel) --------------------------------------
	// constant width <= 29 (the pattern, the shift and two flags share 64 bits), the id fits
	// under the seed shape, the split experiment is off
	if (v.mode == 1 && v.const_len > 0 && v.const_len <= QS2_MAX_LEN && v.id_bits <= v.key_bits &&
	    v.key_bits >= 4 && nent < ((size_t) 1 << 31) && getenv("BSGPU_Q_V1") == nullptr &&
	    getenv("BSGPU_Q_SPLIT") == nullptr) {
		size_t ncells = nkeys / 16, nmain = 0, nover = 0;
		std::vector<uint2> cells(ncells, make_uint2(0u, 0u));
		std::vector<uint32_t> at(nkeys);        // key -> its first slot, then the fill cursor

		for (size_t k2 = 0; k2 < nkeys; k2++) {
			uint32_t slots = std::min<uint32_t>(cnt[k2], 3u);

			if ((k2 & 15) == 0)
				cells[k2 >> 4].y = (uint32_t) nmain;
			cells[k2 >> 4].x |= slots << (2 * (k2 & 15));
			at[k2] = (uint32_t) nmain;
			nmain += slots;
			if (cnt[k2] > 3)
				nover += cnt[k2] - 2;
		}
		std::vector<unsigned long long> slots(nmain + nover + 1, 0ull);
		std::vector<uint8_t> filled(nkeys, 0);  // entries of the key placed so far (counts are <= 254)
		size_t over_at = nmain;

		for (int i = 0; i < M; i++) {
			if (p0->excluded[i])
				continue;
			for (int s2 = 0; s2 < v.nshifts; s2++) {
				int t = v.shift[s2];
				uint32_t k2 = seed_key(i, t), c = cnt[k2], n = filled[k2]++;
				unsigned long long e = (pat[i].x & ~(v.mask2 << (2 * t))) | (q2_deposit(v, (uint32_t) i) << (2 * t)) |
						       ((unsigned long long) t << QS2_T_SHIFT);

				if (c <= 3 || n < 2) {
					slots[at[k2] + n] = e;
					continue;
				}
				if (n == 2) {           // the third slot points at the chain
					slots[at[k2] + 2] = QS2_PTR | (unsigned long long) over_at;
					at[k2] = (uint32_t) over_at - 2;        // chain slot of entry n = at + n
					over_at += c - 2;
				}
				slots[at[k2] + n] = e | (n + 1 < c ? QS2_MORE : 0ull);
			}
		}
		TRY(qi->d_cells2.upload(cells.data(), cells.size()));
		TRY(qi->d_entries2.upload(slots.data(), slots.size()));
		TRY(qi->d_pat.upload(pat.data(), pat.size()));
		{
			std::vector<unsigned long long> p8(M);
			for (int i = 0; i < M; i++)
				p8[i] = pat[i].x;
			TRY(qi->d_pat8.upload(p8.data(), p8.size()));
			v.pat8 = qi->d_pat8.p;
		}
		TRY(qi->d_len.upload(plen.data(), plen.size()));
		v.cells2 = qi->d_cells2.p;
		v.entries2 = qi->d_entries2.p;
		v.pat = qi->d_pat.p;
		v.pat_len = qi->d_len.p;
		v.v2 = 1;
		v.nfp = 0;
		qi->nent = nent;
		qi->max_len = max_w;
		qi->ok = true;
		return done();
	}
	std::vector<uint2> dir(nkeys / 4);
	std::vector<uint32_t> start(nkeys), bitmap(nkeys / 32, 0);
	uint32_t run = 0;
	for (size_t k2 = 0; k2 < nkeys; k2++) {
		start[k2] = run;
		if (cnt[k2])
			bitmap[k2 >> 5] |= 1u << (k2 & 31);
		if ((k2 & 3) == 0) {
			dir[k2 >> 2].x = run;
			dir[k2 >> 2].y = 0;
		}
		dir[k2 >> 2].y |= cnt[k2] << (8 * (k2 & 3));
		run += cnt[k2];
	}
	// entry word = [fingerprint][pattern id][shift index]; the fingerprint is up to 4 pattern
	// letters from outside the seed (the first positions of the window the shape, placed at
	// that shift, does not cover): a cheap pre-check before the pattern is loaded
	v.nfp = std::min(4, (32 - v.shift_bits - v.id_bits) / 2);
	if (getenv("BSGPU_Q_NFP") != nullptr)
		v.nfp = std::min(v.nfp, std::max(0, atoi(getenv("BSGPU_Q_NFP"))));
	for (int s2 = 0; s2 < v.nshifts; s2++) {
		int found = 0;

		v.fp_pos[s2] = 0;
		for (int x = 0; x < std::min(v.min_w, 32) && found < 4; x++) {
			int rel = x - v.shift[s2];
			bool in_seed = rel >= 0 && rel < sh.span && ((sh.mask >> rel) & 1);

			if (!in_seed)
				v.fp_pos[s2] |= (uint32_t) x << (8 * found++);
		}
		v.nfp = std::min(v.nfp, found);
	}
	std::vector<uint32_t> entries(nent ? nent : 1);
	for (int i = 0; i < M; i++) {
		if (p0->excluded[i])
			continue;
		if (nparts == 1 && p0->is_tb_dup[i])
			continue;
		for (int s2 = 0; s2 < v.nshifts; s2++) {
			uint32_t fp = 0;

			for (int f = 0; f < v.nfp; f++)
				fp |= (uint32_t) ((pat[i].x >> (2 * ((v.fp_pos[s2] >> (8 * f)) & 0xFFu))) & 3u) << (2 * f);
			entries[start[seed_key(i, v.shift[s2])]++] =
				(v.nfp ? fp << (v.shift_bits + v.id_bits) : 0u) |
				((uint32_t) i << v.shift_bits) | (uint32_t) s2;
		}
	}
	TRY(qi->d_bitmap.upload(bitmap.data(), bitmap.size()));
	v.bitmap = qi->d_bitmap.p;
	TRY(qi->d_dir.upload(dir.data(), dir.size()));
	TRY(qi->d_entries.upload(entries.data(), entries.size()));
	TRY(qi->d_pat.upload(pat.data(), pat.size()));
	{
		std::vector<unsigned long long> p8(M);
		for (int i = 0; i < M; i++)
			p8[i] = pat[i].x;
		TRY(qi->d_pat8.upload(p8.data(), p8.size()));
		v.pat8 = qi->d_pat8.p;
	}
	TRY(qi->d_len.upload(plen.data(), plen.size()));
	qi->nent = nent;
	v.dir = qi->d_dir.p;
	v.entries = qi->d_entries.p;
	v.pat = qi->d_pat.p;
	v.pat_len = qi->d_len.p;
	qi->max_len = max_w;
	qi->ok = true;
	return done();
}

// Byte-seed index for exact dictionaries (scan_byteseed_kernel): the seeds are 16 encoded
// letters hashed as four raw words, so the scan needs the subject's letters only.
struct ByteSeedIndex {
	bool ok = false;
	BsgpuByteSeedView t = {};
	DevBuf<uint4> d_table, d_filter;
	DevBuf<uint8_t> d_tb;
	size_t nent = 0;
};

static int get_byteseed(const bsgpu_pdict3parts *p3, const ByteSeedIndex **out)
{
	*out = nullptr;
	if (p3->bcache) {
		*out = p3->bcache.get();
		return BSGPU_OK;
	}
	std::unique_ptr<ByteSeedIndex> bi(new ByteSeedIndex());
	auto done = [&]() {
		p3->bcache = std::move(bi);
		*out = p3->bcache.get();
		return BSGPU_OK;
	};
	const int SEED = BSGPU_BYTESEED_LETTERS;
	const char *env = getenv("BSGPU_BYTESEED_STRIDE");
	int M = p3->M, w = p3->w, s2 = std::min(w - SEED + 1, 32);

	if (env != nullptr && atoi(env) >= 1)
		s2 = std::min(s2, atoi(env));
	if (getenv("BSGPU_NO_BYTESEED") != nullptr || p3->has_head || p3->has_tail ||
	    !p3->all_singletons || M == 0)
		return done();
	// the entry id (pattern * S + offset) and >= 6 fingerprint bits share 31 bits
	while (s2 >= 2 && (uint64_t) M * s2 + 2 > (1ull << 25))
		s2--;
	if (s2 < 2)
		return done();
	uint32_t id_bits = 1;
	while (((uint64_t) 1 << id_bits) < (uint64_t) M * s2 + 2)
		id_bits++;
	uint32_t idmask = (1u << id_bits) - 1u, fp_mask = 0x7FFFFFFFu & ~idmask;
	std::vector<uint32_t> digest((size_t) M * s2, 0);
	std::vector<uint8_t> live((size_t) M, 0);
	size_t nent = 0;
	for (int i = 0; i < M; i++) {
		if (p3->excluded[i] || p3->is_tb_dup[i])
			continue;
		live[i] = 1;
		nent += s2;
		const uint8_t *pat = p3->tb.data() + (size_t) i * w;
		for (int t = 0; t < s2; t++) {
			uint32_t wd[4];

			memcpy(wd, pat + t, 16);         // little-endian host, as on the device
			digest[(size_t) i * s2 + t] = byteseed_digest(wd[0], wd[1], wd[2], wd[3]);
		}
	}
	if (nent == 0)
		return done();
	const char *env_l = getenv("BSGPU_SAMPLE_LOAD");
	double load = env_l ? atof(env_l) / 100.0 : 1.5;
	if (load < 0.25) load = 0.25;
	if (load > 3.0) load = 3.0;
	uint64_t nb = std::max<uint64_t>(16, (uint64_t) ((double) nent / load));
	const char *env_f = getenv("BSGPU_BYTESEED_FILTER_LOAD");      // entries per 16-byte Bloom block
	double fload = env_f ? atof(env_f) : 8.0;
	if (fload < 1.0) fload = 1.0;
	uint64_t nf = std::max<uint64_t>(16, (uint64_t) ((double) nent / fload));
	std::vector<uint4> table(nb, make_uint4(BSGPU_S_EMPTY, BSGPU_S_EMPTY, BSGPU_S_EMPTY, BSGPU_S_EMPTY));
	std::vector<uint4> filter(nf, make_uint4(0, 0, 0, 0));
	for (int i = 0; i < M; i++) {
		if (!live[i])
			continue;
		for (int t = 0; t < s2; t++) {
			uint32_t e = (uint32_t) i * s2 + t, b, fp, fb, bits;

			byteseed_table_slot(digest[e] >> 2, (uint32_t) nb, b, fp);
			uint32_t word = sampled_fp_of(fp, id_bits, fp_mask) | e;
			for (;;) {
				uint32_t *slot = &table[b].x;
				int z = 0;
				while (z < 4 && (slot[z] & BSGPU_S_EMPTY) != BSGPU_S_EMPTY)
					z++;
				if (z < 4) {
					slot[z] = (slot[z] & BSGPU_S_OVERFLOW) | word;
					break;
				}
				table[b].w |= BSGPU_S_OVERFLOW;
				b = b + 1 == nb ? 0 : b + 1;
			}
			byteseed_filter_slot(digest[e], (uint32_t) nf, fb, bits);
			filter[fb].x |= 1u << byteseed_bit(bits, 0);
			filter[fb].y |= 1u << byteseed_bit(bits, 1);
			filter[fb].z |= 1u << byteseed_bit(bits, 2);
			filter[fb].w |= 1u << byteseed_bit(bits, 3);
		}
	}
	TRY(bi->d_filter.upload(filter.data(), filter.size()));
	TRY(bi->d_table.upload(table.data(), table.size()));
	{
		std::vector<uint8_t> tbp((size_t) M * w + 8, 0);        // word loads may run 7 bytes over
		memcpy(tbp.data(), p3->tb.data(), (size_t) M * w);
		TRY(bi->d_tb.upload(tbp.data(), tbp.size()));
	}
	bi->nent = nent;
	bi->t.filter = bi->d_filter.p;
	bi->t.nfilter = (uint32_t) nf;
	bi->t.table = bi->d_table.p;
	bi->t.nbuckets = (uint32_t) nb;
	bi->t.tb = bi->d_tb.p;
	bi->t.id_bits = id_bits;
	bi->t.fp_mask = fp_mask;
	bi->t.s = s2;
	bi->t.w = w;
	bi->ok = true;
	return done();
}

// Extended-seed index for exact dictionaries of width >= 25 (scan_xseed_kernel,
// bsgpu_xseed.cuh): every (pattern, offset t) is filed under the filter block of its 10-letter
// core [t, t + 10) and, inside it, under the 18-letter extended seed of its group.
struct XSeedIndex {
	bool ok = false;
	BsgpuXSeedView t = {};
	DevBuf<uint4> d_table;
	DevBuf<BsgpuFilterBlock> d_filter;
	DevBuf<uint8_t> d_tb;
	size_t nent = 0;
	uint64_t nfilter = 0;
};

static int get_xseed(const bsgpu_pdict3parts *p3, const XSeedIndex **out)
{
	*out = nullptr;
	if (p3->xcache) {
		*out = p3->xcache.get();
		return BSGPU_OK;
	}
	std::unique_ptr<XSeedIndex> xi(new XSeedIndex());
	auto done = [&]() {
		p3->xcache = std::move(xi);
		*out = p3->xcache.get();
		return BSGPU_OK;
	};
	const int K = BSGPU_XSEED_CORE, E = BSGPU_XSEED_EXT;
	int M = p3->M, w = p3->w, s2 = std::min(w - K + 1, 32);
	const char *env = getenv("BSGPU_XSEED_STRIDE");

	if (env != nullptr && atoi(env) >= 2 * E)
		s2 = std::min(s2, atoi(env));
	if (getenv("BSGPU_NO_XSEED") != nullptr || getenv("BSGPU_NO_BYTESEED") != nullptr || w < BSGPU_XSEED_MIN_W ||
	    p3->has_head || p3->has_tail || !p3->all_singletons || M == 0)
		return done();
	// the entry id (pattern * S + offset) and >= 6 fingerprint bits share 31 bits
	while (s2 >= 2 * E && (uint64_t) M * s2 + 2 > (1ull << 25))
		s2--;
	if (s2 < 2 * E)
		return done();          // a very large dictionary: the byte-seed engine adapts its stride further
	uint32_t id_bits = 1;
	while (((uint64_t) 1 << id_bits) < (uint64_t) M * s2 + 2)
		id_bits++;
	uint32_t fp_mask = 0x7FFFFFFFu & ~((1u << id_bits) - 1u);
	size_t nent = 0;
	for (int i = 0; i < M; i++)
		if (!p3->excluded[i] && !p3->is_tb_dup[i])
			nent += s2;
	if (nent == 0)
		return done();
	const char *env_l = getenv("BSGPU_SAMPLE_LOAD");
	double load = env_l ? atof(env_l) / 100.0 : 1.5;
	load = std::min(3.0, std::max(0.25, load));
	uint64_t nb = std::max<uint64_t>(16, (uint64_t) ((double) nent / load));
	// filter blocks are addressed by the packed core itself: 4^10 of them (32 MB) when the dictionary
	// brings ~16 entries per block, fewer (leading core bits) for small dictionaries; more than ~24
	// entries per block would drown the filter in false positives: the byte-seed engine takes over
	const char *env_f = getenv("BSGPU_XSEED_FILTER_LOAD");          // target entries per 32-byte block
	double fload = env_f ? atof(env_f) : 8.0;
	fload = std::max(1.0, fload);
	const char *env_k = getenv("BSGPU_XSEED_KBITS");
	int kbits = env_k != nullptr && atoi(env_k) == 8 ? 8 : 4;
	uint32_t fshift = 0;
	while (fshift < 10 && (double) nent / (double) ((1u << 20) >> (fshift + 1)) <= fload)
		fshift++;
	uint64_t nf = (1u << 20) >> fshift;
	if ((double) nent / (double) nf > 24.0)
		return done();
	std::vector<uint4> table(nb, make_uint4(BSGPU_S_EMPTY, BSGPU_S_EMPTY, BSGPU_S_EMPTY, BSGPU_S_EMPTY));
	std::vector<BsgpuFilterBlock> filter(nf);
	memset(filter.data(), 0, nf * sizeof(BsgpuFilterBlock));
	std::vector<uint8_t> buf((size_t) w + 2 * 32, 0);
	for (int i = 0; i < M; i++) {
		if (p3->excluded[i] || p3->is_tb_dup[i])
			continue;
		memcpy(buf.data() + 32, p3->tb.data() + (size_t) i * w, (size_t) w);
		for (int t = 0; t < s2; t++) {
			// the 32 bytes from 8 before the core; letters outside the pattern never reach the
			// digest of the entry's own group (t < S/2: [t, t + 18) <= w; else t >= 8)
			uint32_t wd[8], xc, x0, x1;
			int grp = xseed_group_of(t, s2);

			memcpy(wd, buf.data() + 32 + t - E, 32);        // little-endian host, as on the device
			if (grp == 0) {
				wd[0] = wd[1] = 0;
			} else {
				wd[4] &= 0xFFFFu;
				wd[5] = wd[6] = 0;
			}
			xseed_digests(wd, xc, x0, x1);
			uint32_t x = grp ? x1 : x0, e = (uint32_t) i * s2 + t, b, fp;

			byteseed_table_slot(xseed_payload(x, grp), (uint32_t) nb, b, fp);
			uint32_t word = sampled_fp_of(fp, id_bits, fp_mask) | e;
			for (;;) {
				uint32_t *slot = &table[b].x;
				int z = 0;
				while (z < 4 && (slot[z] & BSGPU_S_EMPTY) != BSGPU_S_EMPTY)
					z++;
				if (z < 4) {
					slot[z] = (slot[z] & BSGPU_S_OVERFLOW) | word;
					break;
				}
				table[b].w |= BSGPU_S_OVERFLOW;
				b = b + 1 == nb ? 0 : b + 1;
			}
			xseed_set(filter[xseed_block(xc, fshift)].v, x, grp, kbits);
		}
	}
	TRY(xi->d_filter.upload(filter.data(), filter.size()));
	TRY(xi->d_table.upload(table.data(), table.size()));
	{
		std::vector<uint8_t> tbp((size_t) M * w + 8, 0);        // word loads may run 7 bytes over
		memcpy(tbp.data(), p3->tb.data(), (size_t) M * w);
		TRY(xi->d_tb.upload(tbp.data(), tbp.size()));
	}
	xi->nent = nent;
	xi->t.filter = xi->d_filter.p;
	xi->t.filter_shift = fshift;
	xi->t.kbits = kbits;
	xi->nfilter = nf;
	xi->t.table = xi->d_table.p;
	xi->t.nbuckets = (uint32_t) nb;
	xi->t.tb = xi->d_tb.p;
	xi->t.id_bits = id_bits;
	xi->t.fp_mask = fp_mask;
	xi->t.s = s2;
	xi->t.w = w;
	xi->ok = true;
	return done();
}

// Band directory for Trusted Bands of <= 32 letters with heads / tails or duplicated bands
// (scan_tbdir_kernel, bsgpu_tbdir.cuh): one slot per key under the first min(w, 12) letters of its
// band, the keys of a band side by side (src/match_pdict_utils.c:153-181).
struct TbDirIndex {
	bool ok = false;
	BsgpuTbDirView v = {};
	DevBuf<uint2> d_cells;
	DevBuf<unsigned long long> d_slots;
	size_t nslots = 0;
};

static int get_tbdir(const bsgpu_pdict3parts *p3, const TbDirIndex **out)
{
	*out = nullptr;
	if (p3->tcache) {
		*out = p3->tcache.get();
		return BSGPU_OK;
	}
	std::unique_ptr<TbDirIndex> ti(new TbDirIndex());
	auto done = [&]() {
		p3->tcache = std::move(ti);
		*out = p3->tcache.get();
		return BSGPU_OK;
	};
	const int M = p3->M, w = p3->w, prefix = std::min(w, TBD_PREFIX);

	if (getenv("BSGPU_NO_TBDIR") != nullptr || getenv("BSGPU_NO_PACKED_FLANKS") != nullptr || w < 2 || w > TBD_MAX_W || M == 0 || M >= (1 << TBD_ID_BITS) ||
	    p3->plain || p3->d_flank_len.n < (size_t) M)
		return done();
	const size_t nkeys = (size_t) 1 << (2 * prefix), ncells = nkeys / 16;
	const uint64_t pmask = ((uint64_t) 1 << (2 * prefix)) - 1;
	std::vector<uint32_t> cnt(nkeys, 0);
	// members of every leader (the groups bsgpu_pdict3parts_blank_dups() leaves are singletons)
	std::vector<int32_t> grp_n(M, 0);
	for (int i = 0; i < M; i++)
		if (p3->high2low[i] != BSGPU_NA)
			grp_n[p3->high2low[i] - 1]++;
	for (int i = 0; i < M; i++)
		if (!p3->excluded[i] && !p3->is_tb_dup[i])
			cnt[p3->keys[i] & pmask] += 1 + (uint32_t) grp_n[i];
	std::vector<uint2> cells(ncells, make_uint2(0u, 0u));
	std::vector<uint32_t> at(nkeys);
	size_t nmain = 0, nover = 0;
	for (size_t k = 0; k < nkeys; k++) {
		uint32_t n3 = std::min<uint32_t>(cnt[k], 3u);

		if ((k & 15) == 0)
			cells[k >> 4].y = (uint32_t) nmain;
		cells[k >> 4].x |= n3 << (2 * (k & 15));
		at[k] = (uint32_t) nmain;
		nmain += n3;
		if (cnt[k] > 3)
			nover += cnt[k] - 2;
	}
	if (nmain + nover >= ((size_t) 1 << 31))
		return done();
	std::vector<unsigned long long> slots(nmain + nover + 1, 0ull);
	std::vector<uint32_t> filled(nkeys, 0);
	size_t over_at = nmain;
	const int nr = std::min(8, w - prefix);
	auto place = [&](int key_id, uint64_t band) {
		int64_t h0 = p3->has_head ? p3->h_head_off[key_id] : 0, t0 = p3->has_tail ? p3->h_tail_off[key_id] : 0;
		int hl = p3->has_head ? (int) (p3->h_head_off[key_id + 1] - h0) : 0;
		int tl = p3->has_tail ? (int) (p3->h_tail_off[key_id + 1] - t0) : 0;
		int nt = 0, nh = 0;
		unsigned long long fp = (band >> (2 * prefix)) & (((unsigned long long) 1 << (2 * nr)) - 1);

		// the leading base letters of the tail, then the trailing base letters of the head
		while (nt < std::min(8 - nr, tl) && bsgpu_is_base(p3->h_tail[(size_t) t0 + nt]))
			nt++;
		while (nh < std::min(8 - nr - nt, hl) && bsgpu_is_base(p3->h_head[(size_t) h0 + hl - 1 - nh]))
			nh++;
		for (int d = 0; d < nt; d++)
			fp |= (unsigned long long) bsgpu_base_code(p3->h_tail[(size_t) t0 + d]) << (2 * (nr + d));
		for (int d = 0; d < nh; d++)    // the last nh letters of the head, in subject order
			fp |= (unsigned long long) bsgpu_base_code(p3->h_head[(size_t) h0 + hl - nh + d]) << (2 * (nr + nt + d));
		unsigned long long e = (unsigned long long) (uint32_t) key_id | (fp << TBD_FP_SHIFT) |
				       ((unsigned long long) nt << TBD_NT_SHIFT) | ((unsigned long long) nh << TBD_NH_SHIFT) |
				       ((unsigned long long) nr << TBD_NR_SHIFT);
		uint64_t cellkey = band & pmask;
		uint32_t c = cnt[cellkey], n = filled[cellkey]++;

		if (c <= 3 || n < 2) {
			slots[at[cellkey] + n] = e;
			return;
		}
		if (n == 2) {           // the third slot points at the chain
			slots[at[cellkey] + 2] = QS2_PTR | (unsigned long long) over_at;
			at[cellkey] = (uint32_t) over_at - 2;   // chain slot of entry n = at + n
			over_at += c - 2;
		}
		slots[at[cellkey] + n] = e | (n + 1 < c ? QS2_MORE : 0ull);
	};
	for (int i = 0; i < M; i++)
		if (!p3->excluded[i] && !p3->is_tb_dup[i])
			place(i, p3->keys[i]);
	for (int i = 0; i < M; i++)
		if (p3->high2low[i] != BSGPU_NA)
			place(i, p3->keys[p3->high2low[i] - 1]);
	TRY(ti->d_cells.upload(cells.data(), cells.size()));
	TRY(ti->d_slots.upload(slots.data(), slots.size()));
	ti->nslots = nmain + nover;
	ti->v.cells = ti->d_cells.p;
	ti->v.slots = ti->d_slots.p;
	ti->v.w = w;
	ti->v.prefix = prefix;
	ti->ok = true;
	return done();
}

static int upload_groups(bsgpu_pdict3parts *p3, bool blank)
{
	int M = p3->M;
	std::vector<int32_t> off(M + 1, 0), members;

	// group of a leader = leader first, then low2high[leader] ascending
	// (src/match_pdict_utils.c:153-181)
	std::vector<int32_t> cnt(M, 0);
	for (int i = 0; i < M; i++)
		if (!blank && p3->high2low[i] != BSGPU_NA)
			cnt[p3->high2low[i] - 1]++;
	for (int i = 0; i < M; i++)
		off[i + 1] = off[i] + 1 + cnt[i];       // every pattern gets a (maybe unused) slot
	members.resize(off[M]);
	std::vector<int32_t> cur(M);
	for (int i = 0; i < M; i++) {
		members[off[i]] = i;
		cur[i] = off[i] + 1;
	}
	bool singles = true;
	for (int i = 0; i < M; i++)
		if (!blank && p3->high2low[i] != BSGPU_NA) {
			int k = p3->high2low[i] - 1;

			members[cur[k]++] = i;
			singles = false;
		}
	p3->all_singletons = singles;
	TRY(p3->d_grp_off.upload(off.data(), off.size()));
	TRY(p3->d_grp_keys.upload(members.data(), members.size()));
	return BSGPU_OK;
}

static int upload_flank(int M, const uint8_t *concat, const int32_t *widths, DevBuf<uint8_t> &d_bytes,
			DevBuf<int64_t> &d_off, bool &present, int &maxw, std::vector<int32_t> *lens,
			std::vector<uint8_t> &h_bytes, std::vector<int64_t> &h_off)
{
	std::vector<int64_t> &off = h_off;

	off.assign(M + 1, 0);

	present = false;
	maxw = 0;
	if (widths != nullptr)
		for (int i = 0; i < M; i++) {
			if (widths[i] < 0)
				return fail(BSGPU_EINVAL, "negative head/tail width for pattern %d", i + 1);
			off[i + 1] = off[i] + widths[i];
			maxw = std::max(maxw, (int) widths[i]);
		}
	if (lens) {
		lens->assign(M, 0);
		if (widths)
			for (int i = 0; i < M; i++)
				(*lens)[i] = widths[i];
	}
	present = maxw > 0;
	if (present)
		h_bytes.assign(concat, concat + off[M]);
	TRY(d_off.upload(off.data(), off.size()));
	TRY(d_bytes.upload(concat, present ? (size_t) off[M] : 0));
	return BSGPU_OK;
}

// Head and tail of every key 2-bit packed for the in-scan XOR + popcount verification
// (tb_hit, direct == 2).  A key qualifies when both flanks are base letters only and at most 32
// letters long; flanks_all_packed tells run_part() that no key needs the byte-wise kernel.
static int upload_packed_flanks(bsgpu_pdict3parts *p3)
{
	int M = p3->M;
	std::vector<ulonglong2> f2((size_t) std::max(M, 1), make_ulonglong2(0, 0));
	std::vector<uint16_t> fl((size_t) std::max(M, 1), 0);
	bool all = true;

	for (int i = 0; i < M; i++) {
		int64_t h0 = p3->has_head ? p3->h_head_off[i] : 0, t0 = p3->has_tail ? p3->h_tail_off[i] : 0;
		int hl = p3->has_head ? (int) (p3->h_head_off[i + 1] - h0) : 0;
		int tl = p3->has_tail ? (int) (p3->h_tail_off[i + 1] - t0) : 0;
		bool ok = hl <= 32 && tl <= 32;
		unsigned long long h = 0, t = 0;

		for (int d = 0; ok && d < hl; d++) {
			uint32_t b = p3->h_head[(size_t) h0 + d];
			ok = bsgpu_is_base(b);
			h |= (unsigned long long) bsgpu_base_code(b) << (2 * d);
		}
		for (int d = 0; ok && d < tl; d++) {
			uint32_t b = p3->h_tail[(size_t) t0 + d];
			ok = bsgpu_is_base(b);
			t |= (unsigned long long) bsgpu_base_code(b) << (2 * d);
		}
		if (ok) {
			f2[i] = make_ulonglong2(h, t);
			fl[i] = (uint16_t) (hl | (tl << 6) | (1 << 15));
		} else if (!(i < (int) p3->excluded.size() && p3->excluded[i])) {
			all = false;
		}
	}
	p3->flanks_all_packed = all && (p3->has_head || p3->has_tail || !p3->all_singletons);
	TRY(p3->d_flank2.upload(f2.data(), f2.size()));
	TRY(p3->d_flank_len.upload(fl.data(), fl.size()));
	return BSGPU_OK;
}

static int bsgpu_pdict3parts_build_impl(int algo, int32_t M,
		const uint8_t *tb_concat, const int32_t *tb_widths,
		const uint8_t *head_concat, const int32_t *head_widths,
		const uint8_t *tail_concat, const int32_t *tail_widths,
		const int32_t *pp_exclude, int32_t *high2low_out,
		bsgpu_pdict3parts **out)
{
	if (out == nullptr || M < 0 || (M > 0 && (tb_concat == nullptr || tb_widths == nullptr)))
		return fail(BSGPU_EINVAL, "bsgpu_pdict3parts_build(): bad arguments");
	*out = nullptr;
	if (algo != BSGPU_ALGO_ACTREE2 && algo != BSGPU_ALGO_TWOBIT)
		return fail(BSGPU_EINVAL, "%d: unsupported Trusted Band type in 'pdict'", algo);
	// --- validation, in the reference's order and with its messages -------------------
	// src/match_pdict_ACtree2.c:801-824,761-763; src/match_pdict_Twobit.c:116-142
	if (M == 0 && algo == BSGPU_ALGO_ACTREE2)
		return fail(BSGPU_EINVAL, "Trusted Band is empty");
	if (M >= 0x3FFFFFFF)
		return fail(BSGPU_EINVAL, "new_ACtree(): tb_length > MAX_P_ID");
	int w = -1;
	{
		const uint8_t *p = tb_concat;

		for (int i = 0; i < M; p += tb_widths[i], i++) {
			if (pp_exclude != nullptr && pp_exclude[i] != BSGPU_NA)
				continue;
			if (algo == BSGPU_ALGO_TWOBIT && tb_widths[i] == 0)
				return fail(BSGPU_EINVAL, "empty trusted region for pattern %d", i + 1);
			if (w == -1) {
				if (tb_widths[i] == 0)
					return fail(BSGPU_EINVAL, "first element in Trusted Band is of length 0");
				w = tb_widths[i];
				if (algo == BSGPU_ALGO_TWOBIT && w > 14)
					return fail(BSGPU_EINVAL, "the width of the Trusted Band must "
						    "be <= 14 when 'type=\"Twobit\"'");
			} else if (tb_widths[i] != w) {
				if (algo == BSGPU_ALGO_TWOBIT)
					return fail(BSGPU_EINVAL, "all the trusted regions must have the same length");
				return fail(BSGPU_EINVAL, "element %d in Trusted Band has a different "
					    "length than first element", i + 1);
			}
			for (int d = 0; d < w; d++)
				if (!bsgpu_is_base(p[d]))
					return fail(BSGPU_EINVAL, algo == BSGPU_ALGO_TWOBIT ?
						    "non-base DNA letter found in Trusted Band for pattern %d" :
						    "non base DNA letter found in Trusted Band for pattern %d",
						    i + 1);
		}
	}
	TRY(ensure_device());

	bsgpu_pdict3parts *p3 = new bsgpu_pdict3parts();
	std::unique_ptr<bsgpu_pdict3parts> guard(p3);
	p3->algo = algo;
	p3->M = M;
	p3->w = w < 0 ? 0 : w;
	p3->seed_w = std::min(p3->w, 32);
	p3->high2low.assign(M, BSGPU_NA);
	p3->keys.assign(M, 0);
	p3->tb.assign((size_t) M * p3->w, 0);

	// --- TB keys, duplicates (high2low), leaders ------------------------------------------
	std::vector<int32_t> leaders;
	std::vector<int32_t> seed_next(M, -1);
	{
		std::unordered_map<uint64_t, int32_t> first32;          // w <= 32: full key
		std::unordered_map<std::string, int32_t> firstlong;     // w > 32: all letters
		std::unordered_map<uint64_t, int32_t> seed_tail;        // w > 32: last leader per seed
		const uint8_t *p = tb_concat;

		first32.reserve((size_t) M * 2);
		for (int i = 0; i < M; p += tb_widths[i], i++) {
			if (pp_exclude != nullptr && pp_exclude[i] != BSGPU_NA)
				continue;
			uint64_t key = 0;
			for (int d = 0; d < p3->seed_w; d++)
				key |= (uint64_t) bsgpu_base_code(p[d]) << (2 * d);
			p3->keys[i] = key;
			memcpy(p3->tb.data() + (size_t) i * w, p, w);
			if (w <= 32) {
				auto it = first32.find(key);
				if (it != first32.end()) {
					p3->high2low[i] = it->second + 1;
					continue;
				}
				first32.emplace(key, i);
			} else {
				std::string s((const char *) p, w);
				auto it = firstlong.find(s);
				if (it != firstlong.end()) {
					p3->high2low[i] = it->second + 1;
					continue;
				}
				firstlong.emplace(std::move(s), i);
				auto st = seed_tail.find(key);
				if (st != seed_tail.end()) {
					seed_next[st->second] = i;      // chain leaders sharing a seed
					st->second = i;
					continue;                       // only the first goes in the table
				}
				seed_tail.emplace(key, i);
			}
			leaders.push_back(i);
		}
	}
	if (high2low_out != nullptr && M > 0)
		memcpy(high2low_out, p3->high2low.data(), (size_t) M * sizeof(int32_t));

	// --- hash table: 16-byte buckets of two {fp,id} slots, linear probing over buckets --------
	uint32_t nb = 16;
	while (nb < 2 * (uint64_t) leaders.size())
		nb <<= 1;
	p3->nbuckets = nb;
	int lg = 0;
	while ((1u << lg) < nb)
		lg++;
	std::vector<uint4> table(nb, make_uint4(0, BSGPU_EMPTY_ID, 0, BSGPU_EMPTY_ID));
	for (int32_t id : leaders) {
		uint32_t top, fp;

		bsgpu_hash(p3->keys[id], top, fp);
		uint32_t b = top >> (32 - lg);
		for (;;) {
			if (table[b].y == BSGPU_EMPTY_ID) {
				table[b].x = fp;
				table[b].y = (uint32_t) id;
				break;
			}
			if ((table[b].w & BSGPU_ID_MASK) == BSGPU_EMPTY_ID) {
				table[b].z = fp;
				table[b].w = (table[b].w & BSGPU_OVERFLOW_FLAG) | (uint32_t) id;
				break;
			}
			table[b].w |= BSGPU_OVERFLOW_FLAG;      // displaced: later buckets must be searched
			b = (b + 1) & (nb - 1);
		}
	}
	TRY(p3->d_table.upload(table.data(), table.size()));
	TRY(p3->d_keys.upload(p3->keys.data(), p3->keys.size()));
	if (w > 32)
		TRY(p3->d_seed_next.upload(seed_next.data(), seed_next.size()));
	TRY(p3->d_tb.upload(p3->tb.data(), p3->tb.size()));
	// every Trusted-Band leader (w > 32: also the chained ones), for the IUPAC brute force
	{
		std::vector<int32_t> all_leaders;
		for (int i = 0; i < M; i++)
			if (!(pp_exclude != nullptr && pp_exclude[i] != BSGPU_NA) && p3->high2low[i] == BSGPU_NA)
				all_leaders.push_back(i);
		p3->n_leaders = (int32_t) all_leaders.size();
		TRY(p3->d_leaders.upload(all_leaders.data(), all_leaders.size()));
	}
	TRY(upload_groups(p3, false));
	TRY(upload_flank(M, head_concat, head_widths, p3->d_head, p3->d_head_off, p3->has_head,
			 p3->max_H, nullptr, p3->h_head, p3->h_head_off));
	TRY(upload_flank(M, tail_concat, tail_widths, p3->d_tail, p3->d_tail_off, p3->has_tail,
			 p3->max_T, &p3->tail_len, p3->h_tail, p3->h_tail_off));
	p3->is_tb_dup.assign(M, 0);
	for (int i = 0; i < M; i++)
		p3->is_tb_dup[i] = p3->high2low[i] != BSGPU_NA;
	p3->excluded.assign(M, 0);
	if (pp_exclude != nullptr)
		for (int i = 0; i < M; i++)
			p3->excluded[i] = pp_exclude[i] != BSGPU_NA;
	TRY(upload_packed_flanks(p3));
	*out = guard.release();
	return BSGPU_OK;
}

extern "C" int bsgpu_pdict3parts_build(int algo, int32_t M,
		const uint8_t *tb_concat, const int32_t *tb_widths,
		const uint8_t *head_concat, const int32_t *head_widths,
		const uint8_t *tail_concat, const int32_t *tail_widths,
		const int32_t *pp_exclude, int32_t *high2low_out,
		bsgpu_pdict3parts **out)
{
	int rc;

	try {
		rc = bsgpu_pdict3parts_build_impl(algo, M, tb_concat, tb_widths, head_concat, head_widths, tail_concat, tail_widths, pp_exclude, high2low_out, out);
	} catch (const std::bad_alloc &) {
		rc = fail(BSGPU_ENOMEM, "bsgpu_pdict3parts_build(): out of host memory");
	} catch (const std::exception &e) {
		rc = fail(BSGPU_ECUDA, "bsgpu_pdict3parts_build(): %s", e.what());
	}
	return rc;
}


// A non-preprocessed dictionary: the handle carries the pattern letters only and is accepted
// by bsgpu_match() / bsgpu_vmatch() like a PDict3Parts (match_XStringSet_XString & co).
static int bsgpu_patterns_build_impl(const uint8_t *concat, const int32_t *widths, int32_t M,
		bsgpu_pdict3parts **out)
{
	if (out == nullptr || M < 0 || (M > 0 && widths == nullptr))
		return fail(BSGPU_EINVAL, "bsgpu_patterns_build(): bad arguments");
	*out = nullptr;
	std::vector<int64_t> off((size_t) M + 1, 0);
	for (int32_t i = 0; i < M; i++) {
		if (widths[i] <= 0)
			return fail(BSGPU_EINVAL, "empty patterns are not supported");
		if (widths[i] > 20000)
			return fail(BSGPU_EINVAL, "patterns with more than 20000 letters are not supported");
		off[i + 1] = off[i] + widths[i];
	}
	int32_t max_plen = 0;
	for (int32_t i = 0; i < M; i++)
		max_plen = std::max(max_plen, widths[i]);
	if (off[M] > 0 && concat == nullptr)
		return fail(BSGPU_EINVAL, "bsgpu_patterns_build(): bad arguments");
	TRY(ensure_device());
	std::unique_ptr<bsgpu_pdict3parts> p3(new bsgpu_pdict3parts());
	p3->plain = true;
	p3->h_pat.assign(concat, concat + off[M]);
	p3->h_pat_off = off;
	p3->max_plen = max_plen;
	p3->M = M;
	p3->w = 0;
	p3->high2low.assign(M, BSGPU_NA);
	TRY(p3->d_pat.alloc((size_t) off[M] + 8));
	if (off[M] > 0)
		CUDA_TRY(cudaMemcpy(p3->d_pat.p, concat, (size_t) off[M], cudaMemcpyHostToDevice));
	TRY(p3->d_pat_off.upload(off.data(), off.size()));
	{
		// {offset, length, first 4 letters} of every pattern: one 16-byte load per pair
		std::vector<uint4> head((size_t) M);
		for (int32_t i = 0; i < M; i++) {
			uint32_t w0 = 0;
			for (int d = 0; d < 4 && d < widths[i]; d++)
				w0 |= (uint32_t) concat[off[i] + d] << (8 * d);
			head[i] = make_uint4((uint32_t) off[i], (uint32_t) ((uint64_t) off[i] >> 32),
					     (uint32_t) widths[i], w0);
		}
		TRY(p3->d_pat_head.upload(head.data(), head.size()));
	}
	*out = p3.release();
	return BSGPU_OK;
}

extern "C" int bsgpu_patterns_build(const uint8_t *concat, const int32_t *widths, int32_t M,
		bsgpu_pdict3parts **out)
{
	int rc;

	try {
		rc = bsgpu_patterns_build_impl(concat, widths, M, out);
	} catch (const std::bad_alloc &) {
		rc = fail(BSGPU_ENOMEM, "bsgpu_patterns_build(): out of host memory");
	} catch (const std::exception &e) {
		rc = fail(BSGPU_ECUDA, "bsgpu_patterns_build(): %s", e.what());
	}
	return rc;
}


extern "C" int bsgpu_pdict3parts_set_flanks(bsgpu_pdict3parts *p3,
		const uint8_t *head_concat, const int32_t *head_widths,
		const uint8_t *tail_concat, const int32_t *tail_widths)
{
	if (p3 == nullptr)
		return fail(BSGPU_EINVAL, "bsgpu_pdict3parts_set_flanks(): NULL handle");
	if (p3->plain)
		return fail(BSGPU_EINVAL, "a non-preprocessed dictionary has no head or tail");
	TRY(ensure_device());
	p3->h_head.clear();
	p3->h_tail.clear();
	TRY(upload_flank(p3->M, head_concat, head_widths, p3->d_head, p3->d_head_off, p3->has_head,
			 p3->max_H, nullptr, p3->h_head, p3->h_head_off));
	TRY(upload_flank(p3->M, tail_concat, tail_widths, p3->d_tail, p3->d_tail_off, p3->has_tail,
			 p3->max_T, &p3->tail_len, p3->h_tail, p3->h_tail_off));
	TRY(upload_packed_flanks(p3));
	p3->qcache.reset();
	p3->bcache.reset();
	p3->xcache.reset();
	p3->tcache.reset();
	return BSGPU_OK;
}

extern "C" int bsgpu_pdict3parts_blank_dups(bsgpu_pdict3parts *p3)
{
	if (p3 == nullptr)
		return fail(BSGPU_EINVAL, "bsgpu_pdict3parts_blank_dups(): NULL handle");
	TRY(ensure_device());
	TRY(upload_groups(p3, true));
	std::fill(p3->high2low.begin(), p3->high2low.end(), BSGPU_NA);
	TRY(upload_packed_flanks(p3));
	p3->qcache.reset();
	p3->bcache.reset();
	p3->xcache.reset();
	p3->tcache.reset();
	return BSGPU_OK;
}

bsgpu_pdict3parts::~bsgpu_pdict3parts() { delete auto_idx; }
extern "C" void bsgpu_pdict3parts_free(bsgpu_pdict3parts *p3) { delete p3; }
extern "C" int32_t bsgpu_pdict3parts_length(const bsgpu_pdict3parts *p3) { return p3 ? p3->M : 0; }
extern "C" int32_t bsgpu_pdict3parts_tb_width(const bsgpu_pdict3parts *p3) { return p3 ? p3->w : 0; }

// ---------------------------------------------------------------------------------------
// Subjects
// ---------------------------------------------------------------------------------------

enum { SUBJ_XSTRING = 0, SUBJ_VIEWS = 1, SUBJ_SET = 2, SUBJ_CHUNK = 3 };

struct bsgpu_subject {
	int kind = SUBJ_XSTRING;
	int32_t nseg = 0;
	int64_t total = 0;              // letters of the concat buffer (multiple of 128)
	int64_t nletters = 0;           // letters scanned
	int64_t scan_lo = 0, scan_hi = 0;       // TB *ends* (concat index) in [scan_lo, scan_hi)
	bool open_left = false, open_right = false;     // chunk edges that are not subject edges
	bool own_from_start = true, own_to_end = true;  // report range reaches the subject's edges
	int64_t halo_left = 0, halo_right = 0;  // letters available around the report range
	std::vector<int64_t> seg_start, seg_coord;
	std::vector<int32_t> seg_len;
	DevBuf<uint8_t> d_bytes;
	// packed planes, built by ensure_packed() on the first scan that needs them (the exact
	// byte-seed scan reads the letters directly)
	mutable DevBuf<uint64_t> d_codes, d_inv2;
	mutable DevBuf<uint32_t> d_valid;
	mutable bool packed = false;
	mutable DevBuf<uint32_t> d_iupac;       // built on first fixedS=FALSE use
	mutable bool iupac_ready = false;
	DevBuf<int64_t> d_seg_start, d_seg_coord;
	DevBuf<int32_t> d_seg_len;

	BsgpuSubjectView view() const
	{
		BsgpuSubjectView S;

		S.bytes = d_bytes.p;
		S.codes = d_codes.p;
		S.valid = d_valid.p;
		S.iupac = d_iupac.p;
		S.inv2 = d_inv2.p;
		S.seg_start = d_seg_start.p;
		S.seg_len = d_seg_len.p;
		S.seg_coord = d_seg_coord.p;
		S.nseg = nseg;
		S.total = total;
		return S;
	}
};

// 2-bit code / validity planes of the subject, derived from its letters on first use.
static int ensure_packed(const bsgpu_subject *s, cudaStream_t st);

// Lays the segments out as [guard][seg0][sep][seg1]...[guard+pad] and uploads them.
// srcs[i] points at host letters of segment i (may alias one big buffer).
// set_concat != NULL: the segments are the consecutive pieces of one packed host buffer
// (a DNAStringSet): it is uploaded as is and laid out by scatter_set_kernel on the device.
// one_src may be a device pointer (bsgpu_subject_element): the copy kind is inferred.
static int subject_finish(bsgpu_subject *s, const uint8_t *const *srcs, const uint8_t *one_src,
			  const uint8_t *set_concat = nullptr)
{
	int64_t pos = BSGPU_GUARD;
	int32_t n = s->nseg;

	s->seg_start.resize(n);
	for (int32_t i = 0; i < n; i++) {
		s->seg_start[i] = pos;
		pos += (int64_t) s->seg_len[i] + 1;     // one separator after every segment
	}
	s->total = ((pos + BSGPU_GUARD + 127) / 128) * 128;
	if (s->total >= (1ll << BSGPU_POS_BITS))
		return fail(BSGPU_EUNSUPPORTED, "subject too long for one device buffer (%lld letters)",
			    (long long) s->total);
	int64_t nwords = s->total / 32;
	TRY(s->d_bytes.alloc((size_t) s->total + 64));
	if (n == 1) {
		// only the guards need the separator byte
		int64_t seg_end = s->seg_start[0] + s->seg_len[0];

		CUDA_TRY(cudaMemsetAsync(s->d_bytes.p, BSGPU_SEP_BYTE, (size_t) s->seg_start[0]));
		CUDA_TRY(cudaMemsetAsync(s->d_bytes.p + seg_end, BSGPU_SEP_BYTE,
					 (size_t) (s->total + 64 - seg_end)));
		if (s->seg_len[0] > 0 && (one_src != nullptr || srcs != nullptr))      // else: filled by the caller
			CUDA_TRY(cudaMemcpyAsync(s->d_bytes.p + s->seg_start[0], one_src ? one_src : srcs[0],
						 (size_t) s->seg_len[0], cudaMemcpyDefault));
	} else {
		CUDA_TRY(cudaMemsetAsync(s->d_bytes.p, BSGPU_SEP_BYTE, (size_t) s->total + 64));
	}
	if (n > 1 && set_concat != nullptr) {
		std::vector<int64_t> src_off((size_t) n);
		int64_t tot = 0;
		for (int32_t i = 0; i < n; i++) {
			src_off[i] = tot;
			tot += s->seg_len[i];
		}
		DevBuf<uint8_t> d_src;
		DevBuf<int64_t> d_src_off;
		TRY(d_src.alloc((size_t) tot));
		if (tot > 0)
			CUDA_TRY(cudaMemcpyAsync(d_src.p, set_concat, (size_t) tot, cudaMemcpyHostToDevice));
		TRY(d_src_off.upload(src_off.data(), (size_t) n));
		TRY(s->d_seg_start.upload(s->seg_start.data(), (size_t) n));
		TRY(s->d_seg_len.upload(s->seg_len.data(), (size_t) n));
		scatter_set_kernel<<<grid_for((long long) n * 32, 256, 8), 256>>>(d_src.p, d_src_off.p,
				s->d_seg_start.p, s->d_seg_len.p, n, s->d_bytes.p);
		g_launches++;
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaDeviceSynchronize());      // d_src goes back to the pool
	} else if (n > 1) {
		// assemble on the host in slabs to bound the staging memory
		const int64_t SLAB = 256ll << 20;
		std::vector<uint8_t> stage;
		int32_t i = 0;

		while (i < n) {
			int64_t a = s->seg_start[i], b = a;
			int32_t k = i;

			while (k < n && (s->seg_start[k] + s->seg_len[k] + 1 - a <= SLAB || k == i)) {
				b = s->seg_start[k] + s->seg_len[k] + 1;
				k++;
			}
			stage.assign((size_t) (b - a), (uint8_t) BSGPU_SEP_BYTE);
			for (int32_t q = i; q < k; q++)
				if (s->seg_len[q] > 0)
					memcpy(stage.data() + (s->seg_start[q] - a), srcs[q], (size_t) s->seg_len[q]);
			CUDA_TRY(cudaMemcpy(s->d_bytes.p + a, stage.data(), (size_t) (b - a),
					    cudaMemcpyHostToDevice));
			i = k;
		}
	}
	s->packed = false;
	s->iupac_ready = false;
	TRY(s->d_seg_start.upload(s->seg_start.data(), (size_t) n));
	TRY(s->d_seg_len.upload(s->seg_len.data(), (size_t) n));
	TRY(s->d_seg_coord.upload(s->seg_coord.data(), (size_t) n));
	CUDA_TRY(cudaDeviceSynchronize());
	if (s->kind != SUBJ_CHUNK) {
		s->scan_lo = BSGPU_GUARD;
		s->scan_hi = pos;
	}
	return BSGPU_OK;
}

extern "C" int bsgpu_subject_xstring(const uint8_t *sq, int64_t length, bsgpu_subject **out)
{
	if (out == nullptr || length < 0 || (length > 0 && sq == nullptr))
		return fail(BSGPU_EINVAL, "bsgpu_subject_xstring(): bad arguments");
	*out = nullptr;
	if (length > INT32_MAX)
		return fail(BSGPU_EINVAL, "subject longer than INT_MAX letters (use chunks or views)");
	TRY(ensure_device());
	std::unique_ptr<bsgpu_subject> s(new bsgpu_subject());
	s->kind = SUBJ_XSTRING;
	s->nseg = 1;
	s->seg_len.assign(1, (int32_t) length);
	s->seg_coord.assign(1, 0);
	s->nletters = length;
	TRY(subject_finish(s.get(), nullptr, sq));
	*out = s.release();
	return BSGPU_OK;
}

extern "C" int bsgpu_subject_views(const uint8_t *sq, int64_t length,
		const int32_t *views_start, const int32_t *views_width, int32_t nviews,
		bsgpu_subject **out)
{
	if (out == nullptr || length < 0 || nviews < 0 || (length > 0 && sq == nullptr) ||
	    (nviews > 0 && (views_start == nullptr || views_width == nullptr)))
		return fail(BSGPU_EINVAL, "bsgpu_subject_views(): bad arguments");
	*out = nullptr;
	std::vector<const uint8_t *> srcs(nviews);
	std::unique_ptr<bsgpu_subject> s(new bsgpu_subject());
	s->kind = SUBJ_VIEWS;
	s->nseg = nviews;
	s->seg_len.resize(nviews);
	s->seg_coord.resize(nviews);
	for (int32_t v = 0; v < nviews; v++) {
		int64_t off = (int64_t) views_start[v] - 1;

		// src/match_pdict.c:252-254
		if (off < 0 || views_width[v] < 0 || off + views_width[v] > length)
			return fail(BSGPU_EINVAL, "'subject' has \"out of limits\" views");
		srcs[v] = sq + off;
		s->seg_len[v] = views_width[v];
		s->seg_coord[v] = off;
		s->nletters += views_width[v];
	}
	TRY(ensure_device());
	TRY(subject_finish(s.get(), srcs.data(), nullptr));
	*out = s.release();
	return BSGPU_OK;
}

extern "C" int bsgpu_subject_set(const uint8_t *concat, const int32_t *widths, int32_t n,
		bsgpu_subject **out)
{
	if (out == nullptr || n < 0 || (n > 0 && widths == nullptr))
		return fail(BSGPU_EINVAL, "bsgpu_subject_set(): bad arguments");
	*out = nullptr;
	std::vector<const uint8_t *> srcs(n);
	std::unique_ptr<bsgpu_subject> s(new bsgpu_subject());
	s->kind = SUBJ_SET;
	s->nseg = n;
	s->seg_len.resize(n);
	s->seg_coord.assign(n, 0);
	int64_t off = 0;
	for (int32_t i = 0; i < n; i++) {
		if (widths[i] < 0)
			return fail(BSGPU_EINVAL, "negative width for subject element %d", i + 1);
		srcs[i] = concat + off;
		s->seg_len[i] = widths[i];
		off += widths[i];
	}
	s->nletters = off;
	TRY(ensure_device());
	TRY(subject_finish(s.get(), srcs.data(), nullptr, concat));
	*out = s.release();
	return BSGPU_OK;
}

extern "C" int bsgpu_subject_chunk(const uint8_t *chunk, int64_t chunk_len, int64_t chunk_start,
		int64_t subject_len, int64_t report_lo, int64_t report_hi, bsgpu_subject **out)
{
	if (out == nullptr || chunk_len < 0 || chunk_start < 0 || chunk_start + chunk_len > subject_len ||
	    (chunk_len > 0 && chunk == nullptr) || chunk_len > INT32_MAX)
		return fail(BSGPU_EINVAL, "bsgpu_subject_chunk(): bad arguments");
	*out = nullptr;
	// clamp the report range (1-based TB ends n, lo < n <= hi) to the chunk
	report_lo = std::max(report_lo, chunk_start);
	report_hi = std::min(report_hi, chunk_start + chunk_len);
	if (report_hi < report_lo)
		report_hi = report_lo;
	TRY(ensure_device());
	std::unique_ptr<bsgpu_subject> s(new bsgpu_subject());
	s->kind = SUBJ_CHUNK;
	s->nseg = 1;
	s->seg_len.assign(1, (int32_t) chunk_len);
	s->seg_coord.assign(1, chunk_start);
	s->nletters = report_hi - report_lo;
	s->open_left = chunk_start > 0;
	s->open_right = chunk_start + chunk_len < subject_len;
	s->halo_left = report_lo - chunk_start;                 // letters before the first reported end
	s->halo_right = chunk_start + chunk_len - report_hi;    // letters after the last reported end
	s->own_from_start = report_lo <= 0;
	s->own_to_end = report_hi >= subject_len;
	TRY(subject_finish(s.get(), nullptr, chunk));
	// TB end n (subject coords) <-> concat index seg_start + (n - chunk_start) - 1
	s->scan_lo = s->seg_start[0] + (report_lo - chunk_start);
	s->scan_hi = s->seg_start[0] + (report_hi - chunk_start);
	*out = s.release();
	return BSGPU_OK;
}

static int ensure_packed(const bsgpu_subject *s, cudaStream_t st)
{
	if (s->packed)
		return BSGPU_OK;
	int64_t nwords = s->total / 32;

	if (s->d_codes.n < (size_t) nwords + 4) {
		TRY(s->d_codes.alloc((size_t) nwords + 4));
		TRY(s->d_valid.alloc((size_t) nwords + 4));
		TRY(s->d_inv2.alloc((size_t) nwords + 4));
		// the pack kernel writes words [0, nwords); only the 4 spare words need a value
		CUDA_TRY(cudaMemsetAsync(s->d_codes.p + nwords, 0, 4 * sizeof(uint64_t), st));
		CUDA_TRY(cudaMemsetAsync(s->d_valid.p + nwords, 0, 4 * sizeof(uint32_t), st));
		CUDA_TRY(cudaMemsetAsync(s->d_inv2.p + nwords, 0xFF, 4 * sizeof(uint64_t), st));
	}
	{
		KernelTimer kt("pack_subject_kernel", st);
		pack_subject_kernel<<<grid_for(2 * nwords, 256, 8), 256, 0, st>>>(s->d_bytes.p, nwords, s->d_codes.p,
										  s->d_valid.p, s->d_inv2.p);
	}
	g_launches++;
	CUDA_TRY(cudaGetLastError());
	s->packed = true;
	return BSGPU_OK;
}

// The caller overwrote the letters in place (same layout): the derived planes are stale.
extern "C" int bsgpu_subject_repack(bsgpu_subject *s, void *stream)
{
	(void) stream;
	if (s == nullptr)
		return fail(BSGPU_EINVAL, "bsgpu_subject_repack(): NULL handle");
	s->packed = false;
	s->iupac_ready = false;
	return BSGPU_OK;
}

extern "C" void bsgpu_subject_free(bsgpu_subject *s) { delete s; }
extern "C" int64_t bsgpu_subject_nletters(const bsgpu_subject *s) { return s ? s->nletters : 0; }

// ---------------------------------------------------------------------------------------
// Subject ingestion: FASTA text -> resident subject (read_fasta_files, src/read_fasta_files.c:412)
// ---------------------------------------------------------------------------------------

extern "C" void bsgpu_fasta_info_free(bsgpu_fasta_info *info)
{
	if (info == nullptr)
		return;
	free(info->widths);
	free(info->desc_offset);
	free(info->desc_length);
	memset(info, 0, sizeof(*info));
}

extern "C" int bsgpu_subject_fasta(const uint8_t *text, int64_t nbytes, const int32_t *lkup, int lkup_length,
		bsgpu_subject **out, bsgpu_fasta_info *info)
{
	if (out == nullptr || info == nullptr || nbytes < 0 || (nbytes > 0 && text == nullptr) ||
	    lkup == nullptr || lkup_length < 0)
		return fail(BSGPU_EINVAL, "bsgpu_subject_fasta(): bad arguments");
	*out = nullptr;
	memset(info, 0, sizeof(*info));
	TRY(ensure_device());
	// translate_byte(): keys >= lkup_length and NA entries are not letters of the alphabet
	int16_t lk16[256];
	for (int k = 0; k < 256; k++)
		lk16[k] = k < lkup_length && lkup[k] != BSGPU_NA ? (int16_t) (lkup[k] & 0xFF) : (int16_t) -1;
	const int64_t nchunks = (nbytes + BSGPU_FASTA_CHUNK - 1) / BSGPU_FASTA_CHUNK;
	const size_t padded = (size_t) ((nbytes + 15) / 16 * 16 + 16);
	DevBuf<uint8_t> d_text;
	DevBuf<int16_t> d_lkup;
	const int64_t ngroups = (nchunks + BSGPU_FASTA_GROUP - 1) / BSGPU_FASTA_GROUP;
	DevBuf<BsgpuFastaSummary> d_sums;
	DevBuf<BsgpuFastaGroup> d_groups;
	DevBuf<BsgpuFastaResolved> d_res, d_bases;
	std::vector<BsgpuFastaGroup> groups((size_t) ngroups);
	std::vector<BsgpuFastaResolved> bases;

	TRY(d_text.alloc(padded));
	{
		size_t tail = std::min<size_t>(padded, 32);             // covers everything past the text

		CUDA_TRY(cudaMemsetAsync(d_text.p + (padded - tail), '\n', tail));
	}
	if (nbytes > 0)
		CUDA_TRY(cudaMemcpyAsync(d_text.p, text, (size_t) nbytes, cudaMemcpyHostToDevice));
	TRY(d_lkup.upload(lk16, 256));
	TRY(d_sums.alloc((size_t) nchunks));
	if (nchunks > 0) {
		KernelTimer kt("fasta_summary_kernel", 0);
		fasta_summary_kernel<<<(unsigned) ((nchunks + 127) / 128), 128>>>(d_text.p, nbytes, nchunks, d_lkup.p, d_sums.p);
		g_launches++;
	}
	CUDA_TRY(cudaGetLastError());
	// between the passes: the chunk summaries stay on the device, only one summary per 64 chunks
	// crosses PCIe for the host's sequential step
	TRY(d_groups.alloc((size_t) ngroups));
	if (ngroups > 0) {
		fasta_group_kernel<<<(unsigned) ((ngroups + 127) / 128), 128>>>(d_sums.p, nchunks, ngroups, d_groups.p);
		g_launches++;
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaMemcpy(groups.data(), d_groups.p, (size_t) ngroups * sizeof(BsgpuFastaGroup),
				    cudaMemcpyDeviceToHost));
	}
	BsgpuFastaTotals t = fasta_resolve_groups(groups, bases);
	if (t.first_seqline >= 0 && (t.first_header < 0 || t.first_seqline < t.first_header)) {
		// src/read_fasta_files.c:322-327; lines are counted like the reference counts them
		long long lineno = 1;
		for (int64_t i = 0; i < t.first_seqline; i++)
			lineno += text[i] == '\n';
		return fail(BSGPU_EINVAL, "\">\" expected at beginning of line %lld", lineno);
	}
	if (t.nrec > INT32_MAX)
		return fail(BSGPU_EUNSUPPORTED, "more than INT_MAX FASTA records");
	const int32_t nrec = (int32_t) t.nrec;
	std::unique_ptr<bsgpu_subject> s(new bsgpu_subject());
	int64_t pos = BSGPU_GUARD + t.kept + nrec;      // one separator after every record

	s->kind = SUBJ_SET;
	s->nseg = nrec;
	s->nletters = t.kept;
	s->total = ((pos + BSGPU_GUARD + 127) / 128) * 128;
	if (s->total >= (1ll << BSGPU_POS_BITS))
		return fail(BSGPU_EUNSUPPORTED, "subject too long for one device buffer (%lld letters)",
			    (long long) s->total);
	TRY(s->d_bytes.alloc((size_t) s->total + 64));
	CUDA_TRY(cudaMemsetAsync(s->d_bytes.p, BSGPU_SEP_BYTE, (size_t) s->total + 64));
	TRY(s->d_seg_start.alloc((size_t) nrec));
	DevBuf<int64_t> d_desc_off;
	TRY(d_desc_off.alloc((size_t) nrec));
	if (nchunks > 0) {
		TRY(d_bases.upload(bases.data(), (size_t) ngroups));
		TRY(d_res.alloc((size_t) nchunks));
		fasta_expand_kernel<<<(unsigned) ((ngroups + 127) / 128), 128>>>(d_sums.p, nchunks, ngroups, d_bases.p, d_res.p);
		g_launches++;
		CUDA_TRY(cudaGetLastError());
		KernelTimer kt("fasta_emit_kernel", 0);
		fasta_emit_kernel<<<(unsigned) ((nchunks + 127) / 128), 128>>>(d_text.p, nbytes, nchunks, d_lkup.p, d_res.p,
				s->d_bytes.p, s->d_seg_start.p, d_desc_off.p);
		g_launches++;
	}
	CUDA_TRY(cudaGetLastError());
	s->seg_start.resize((size_t) nrec);
	s->seg_len.resize((size_t) nrec);
	s->seg_coord.assign((size_t) nrec, 0);
	std::vector<int64_t> desc_off((size_t) nrec);
	if (nrec > 0) {
		CUDA_TRY(cudaMemcpy(s->seg_start.data(), s->d_seg_start.p, (size_t) nrec * sizeof(int64_t),
				    cudaMemcpyDeviceToHost));
		CUDA_TRY(cudaMemcpy(desc_off.data(), d_desc_off.p, (size_t) nrec * sizeof(int64_t),
				    cudaMemcpyDeviceToHost));
	}
	CUDA_TRY(cudaDeviceSynchronize());
	for (int32_t r = 0; r < nrec; r++) {
		int64_t end = r + 1 < nrec ? s->seg_start[(size_t) r + 1] - 1 : pos - 1;
		int64_t len = end - s->seg_start[(size_t) r];

		if (len < 0 || len > INT32_MAX)
			return fail(BSGPU_EUNSUPPORTED, "FASTA record %d has more than INT_MAX letters", r + 1);
		s->seg_len[(size_t) r] = (int32_t) len;
	}
	TRY(s->d_seg_len.upload(s->seg_len.data(), (size_t) nrec));
	TRY(s->d_seg_coord.upload(s->seg_coord.data(), (size_t) nrec));
	s->scan_lo = BSGPU_GUARD;
	s->scan_hi = pos;
	s->packed = false;
	s->iupac_ready = false;
	// what the host layer needs to build names(): where every description starts and ends
	info->nrec = nrec;
	info->ninvalid = t.invalid;
	info->nletters = t.kept;
	info->widths = (int32_t *) malloc((size_t) std::max(nrec, 1) * sizeof(int32_t));
	info->desc_offset = (int64_t *) malloc((size_t) std::max(nrec, 1) * sizeof(int64_t));
	info->desc_length = (int32_t *) malloc((size_t) std::max(nrec, 1) * sizeof(int32_t));
	if (info->widths == nullptr || info->desc_offset == nullptr || info->desc_length == nullptr) {
		bsgpu_fasta_info_free(info);
		return fail(BSGPU_ENOMEM, "out of host memory");
	}
	for (int32_t r = 0; r < nrec; r++) {
		int64_t a = desc_off[(size_t) r];
		const uint8_t *nl = a < nbytes ? (const uint8_t *) memchr(text + a, '\n', (size_t) (nbytes - a)) : nullptr;
		int64_t e = nl != nullptr ? nl - text : nbytes;

		if (nl != nullptr && e > a && text[e - 1] == '\r')      // delete_trailing_LF_or_CRLF
			e--;
		info->widths[r] = s->seg_len[(size_t) r];
		info->desc_offset[r] = a;
		info->desc_length[r] = (int32_t) std::min<int64_t>(e - a, INT32_MAX);
	}
	set_plan("fasta_ingest bytes=%lld chunks=%lld records=%d letters=%lld invalid=%lld", (long long) nbytes,
		 (long long) nchunks, nrec, (long long) t.kept, (long long) t.invalid);
	*out = s.release();
	return BSGPU_OK;
}

// DNAString(<character>) straight onto the device: the letters are uploaded as they are and
// translated by encode_letters_kernel (replaces the copy-with-lookup loop behind
// new_XString_from_CHARACTER, src/XString_class.c:134-159,191-218 + XVector's
// Ocopy_bytes_from_i1i2_with_lkup, with the tables of R/XStringCodec-class.R:114-166).
extern "C" int bsgpu_subject_letters(const uint8_t *letters, int64_t length, const int32_t *lkup, int lkup_length,
		bsgpu_subject **out)
{
	if (out == nullptr || length < 0 || (length > 0 && letters == nullptr) || lkup == nullptr || lkup_length < 0)
		return fail(BSGPU_EINVAL, "bsgpu_subject_letters(): bad arguments");
	*out = nullptr;
	if (length > INT32_MAX)
		return fail(BSGPU_EINVAL, "subject longer than INT_MAX letters (use chunks or views)");
	TRY(ensure_device());
	int16_t lk16[256];
	for (int k = 0; k < 256; k++)
		lk16[k] = k < lkup_length && lkup[k] != BSGPU_NA ? (int16_t) (lkup[k] & 0xFF) : (int16_t) -1;
	std::unique_ptr<bsgpu_subject> s(new bsgpu_subject());
	s->kind = SUBJ_XSTRING;
	s->nseg = 1;
	s->seg_len.assign(1, (int32_t) length);
	s->seg_coord.assign(1, 0);
	s->nletters = length;
	TRY(subject_finish(s.get(), nullptr, nullptr));         // layout and guards; no letters yet
	if (length > 0) {
		const size_t padded = (size_t) ((length + 15) / 16 * 16);
		DevBuf<uint8_t> d_src;
		DevBuf<int16_t> d_lkup;
		DevBuf<unsigned long long> d_bad;
		unsigned long long bad = ~0ull;

		TRY(d_src.alloc(padded));
		CUDA_TRY(cudaMemcpyAsync(d_src.p, letters, (size_t) length, cudaMemcpyHostToDevice));
		TRY(d_lkup.upload(lk16, 256));
		TRY(d_bad.upload(&bad, 1));
		{
			KernelTimer kt("encode_letters_kernel", 0);
			encode_letters_kernel<<<grid_for((long long) (padded / 16), 256, 16), 256>>>(d_src.p, length, d_lkup.p,
					s->d_bytes.p + s->seg_start[0], d_bad.p);
		}
		g_launches++;
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaMemcpy(&bad, d_bad.p, sizeof(bad), cudaMemcpyDeviceToHost));
		if (bad != ~0ull) {
			// XVector's lookup copy stops at the first letter the lookup does not know
			int c = letters[bad];

			return fail(BSGPU_EINVAL, "key %d (char '%c') not in lookup table", c, c);
		}
	}
	*out = s.release();
	return BSGPU_OK;
}

extern "C" int32_t bsgpu_subject_nsequences(const bsgpu_subject *subj) { return subj ? subj->nseg : 0; }

extern "C" int bsgpu_subject_element(const bsgpu_subject *set, int32_t i, bsgpu_subject **out)
{
	if (set == nullptr || out == nullptr)
		return fail(BSGPU_EINVAL, "bsgpu_subject_element(): bad arguments");
	*out = nullptr;
	if (i < 0 || i >= set->nseg)
		return fail(BSGPU_EINVAL, "subscript out of bounds");
	TRY(ensure_device());
	std::unique_ptr<bsgpu_subject> s(new bsgpu_subject());
	s->kind = SUBJ_XSTRING;
	s->nseg = 1;
	s->seg_len.assign(1, set->seg_len[(size_t) i]);
	s->seg_coord.assign(1, 0);
	s->nletters = set->seg_len[(size_t) i];
	TRY(subject_finish(s.get(), nullptr, set->d_bytes.p + set->seg_start[(size_t) i]));
	*out = s.release();
	return BSGPU_OK;
}

extern "C" int bsgpu_subject_download(const bsgpu_subject *subj, uint8_t *concat_out, int32_t *widths_out)
{
	if (subj == nullptr)
		return fail(BSGPU_EINVAL, "bsgpu_subject_download(): NULL subject");
	TRY(ensure_device());
	if (widths_out != nullptr)
		for (int32_t r = 0; r < subj->nseg; r++)
			widths_out[r] = subj->seg_len[(size_t) r];
	if (concat_out == nullptr || subj->nseg == 0)
		return BSGPU_OK;
	// the records lie back to back with one separator in between: one copy, then close the gaps
	int64_t first = subj->seg_start[0];
	int64_t last = subj->seg_start[(size_t) subj->nseg - 1] + subj->seg_len[(size_t) subj->nseg - 1];
	std::vector<uint8_t> stage((size_t) (last - first) + 1);
	CUDA_TRY(cudaMemcpy(stage.data(), subj->d_bytes.p + first, (size_t) (last - first), cudaMemcpyDeviceToHost));
	size_t off = 0;
	for (int32_t r = 0; r < subj->nseg; r++) {
		memcpy(concat_out + off, stage.data() + (subj->seg_start[(size_t) r] - first), (size_t) subj->seg_len[(size_t) r]);
		off += (size_t) subj->seg_len[(size_t) r];
	}
	return BSGPU_OK;
}

// ---------------------------------------------------------------------------------------
// Workspace: candidate / record buffers and counters (per process, grown on demand)
// ---------------------------------------------------------------------------------------

struct Workspace {
	DevBuf<unsigned long long> cand, rec, rec2;
	DevBuf<unsigned int> xseed_counts;      // candidates per scan warp (scan_xseed_kernel)
	DevBuf<int32_t> row_cursor, big_rows;   // count - scan - write of the hit records
	unsigned long long rec_known = 0;       // counters[1] as of the last host read / write (no read-back needed)
	DevBuf<unsigned long long> counters;    // [0] = candidates, [1] = records
	DevBuf<int> flags;
	DevBuf<long long> ovf;                  // windows with too many IUPAC expansions
	DevBuf<int64_t> slot_off;               // plain dictionaries: first slot of every segment
	BsgpuPlainView plain = {};              // view of the last plain scan (for finalize_ends_kernel)
	DevBuf<uint8_t> cub_temp;
	unsigned long long *h_counters = nullptr;       // pinned
	int *h_flags = nullptr;
	int32_t *h_counts = nullptr;                    // pinned, grown on demand
	size_t h_counts_n = 0;
};

static Workspace *g_ws = nullptr;

static bool device_state_exists(void)
{
	return g_ws != nullptr || !g_pool.empty() || g_live_allocs > 0;
}

static int host_counts(Workspace *ws, size_t n, int32_t **out)
{
	if (ws->h_counts_n < n || ws->h_counts == nullptr) {
		if (ws->h_counts != nullptr)
			cudaFreeHost(ws->h_counts);
		ws->h_counts = nullptr;
		ws->h_counts_n = 0;
		size_t want = std::max<size_t>(n, 1024);
		CUDA_TRY(cudaMallocHost((void **) &ws->h_counts, want * sizeof(int32_t)));
		ws->h_counts_n = want;
	}
	*out = ws->h_counts;
	return BSGPU_OK;
}

static int workspace(Workspace **out)
{
	if (g_ws == nullptr) {
		Workspace *w = new Workspace();

		if (w->counters.alloc(4) != BSGPU_OK || w->flags.alloc(4) != BSGPU_OK) {
			delete w;
			return BSGPU_ENOMEM;
		}
		CUDA_TRY(cudaMallocHost((void **) &w->h_counters, 4 * sizeof(unsigned long long)));
		CUDA_TRY(cudaMallocHost((void **) &w->h_flags, 4 * sizeof(int)));
		g_ws = w;
	}
	*out = g_ws;
	return BSGPU_OK;
}

static const size_t CAND_MIN = 1u << 20;
static const size_t CAND_MAX = 1ull << 30;      // records (8 GB)

static int grow(DevBuf<unsigned long long> &buf, size_t need, size_t keep, cudaStream_t st)
{
	if (need <= buf.n)
		return BSGPU_OK;
	size_t cap = std::max(need + need / 8, CAND_MIN);
	DevBuf<unsigned long long> nb;

	TRY(nb.alloc(cap));
	if (keep > 0) {
		CUDA_TRY(cudaMemcpyAsync(nb.p, buf.p, keep * sizeof(unsigned long long),
					 cudaMemcpyDeviceToDevice, st));
		CUDA_TRY(cudaStreamSynchronize(st));
	}
	buf.swap(nb);
	return BSGPU_OK;
}

// ---------------------------------------------------------------------------------------
// The matching driver (match_pdict(), src/match_pdict.c:40-77, for one PDict3Parts)
// ---------------------------------------------------------------------------------------

struct MatchParams {
	int max_nmis, min_nmis, fixedP, fixedS;
};

// blocks per SM the grid-stride kernels of the Trusted-Band engine are capped at
static int tb_grid(void)
{
	const char *env = getenv("BSGPU_TB_GRID");

	return env != nullptr && atoi(env) > 0 ? atoi(env) : 64;
}

static unsigned int iupac_max_expansions(void)
{
	const char *env = getenv("BSGPU_MAX_IUPAC_EXPANSIONS");

	return env ? (unsigned int) strtoul(env, nullptr, 10) : 4096u;
}

static int read_counter(Workspace *ws, int which, unsigned long long *out, cudaStream_t st)
{
	CUDA_TRY(cudaMemcpyAsync(ws->h_counters + which, ws->counters.p + which,
				 sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	*out = ws->h_counters[which];
	if (which == 1)
		ws->rec_known = *out;
	return BSGPU_OK;
}

static int write_counter(Workspace *ws, int which, unsigned long long v, cudaStream_t st)
{
	if (which == 1)
		ws->rec_known = v;
	if (v == 0) {           // stream-ordered, nothing to wait for
		CUDA_TRY(cudaMemsetAsync(ws->counters.p + which, 0, sizeof(unsigned long long), st));
		return BSGPU_OK;
	}
	ws->h_counters[2] = v;
	CUDA_TRY(cudaMemcpyAsync(ws->counters.p + which, ws->h_counters + 2,
				 sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	return BSGPU_OK;
}

static int launch_scans(const bsgpu_pdict3parts *p3, const bsgpu_subject *subj, const MatchParams &mp,
			const BsgpuReport &R, const BsgpuReport &C, int direct,
			int64_t end_lo, int64_t end_hi, Workspace *ws, cudaStream_t st)
{
	// window starts for TB ends in [end_lo, end_hi)
	int64_t start_lo = end_lo - p3->w + 1, start_hi = end_hi - p3->w + 1;

	if (start_lo < 0)
		start_lo = 0;
	if (start_hi <= start_lo)
		return BSGPU_OK;
	int64_t nwords = ((start_hi + 31) >> 5) - (start_lo >> 5);
	BsgpuSubjectView S = subj->view();
	BsgpuIndexView I = p3->view();

	if (direct == 1 && mp.fixedS) {
		// exact dictionary of width >= 25, letters only: short core + extended seeds, one probe
		// every S = w - 9 letters (bsgpu_xseed.cuh)
		const XSeedIndex *xi = nullptr;
		TRY(get_xseed(p3, &xi));
		if (xi != nullptr && xi->ok) {
			int64_t S_ = xi->t.s;
			// S == 16: probes at p = 8 (mod 16), so that a probe's 32 bytes [p - 8, p + 24) are two
			// aligned 16-byte chunks (the concat buffer is 16-byte aligned); any fixed phase works
			bool aligned = S_ == 16 && getenv("BSGPU_XSEED_UNALIGNED") == nullptr;
			int64_t phase = aligned ? 8 : 0;
			int64_t lo_p = std::max<int64_t>(start_lo, 32);         // matches start behind the guard (128)
			int64_t first = lo_p + ((phase - lo_p) % S_ + S_) % S_;
			int64_t probe_hi = start_hi + S_ - 1;           // exclusive
			int64_t nprobes = probe_hi > first ? (probe_hi - first + S_ - 1) / S_ : 0;

			set_plan("byteseed_hash xseed core=%d ext=%d S=%d filter_bytes=%llu buckets=%u table_bytes=%llu "
				 "entries_per_bucket=%.2f%s", BSGPU_XSEED_CORE, BSGPU_XSEED_EXT, xi->t.s,
				 (unsigned long long) xi->nfilter * 32, xi->t.nbuckets,
				 (unsigned long long) xi->t.nbuckets * 16, (double) xi->nent / xi->t.nbuckets,
				 aligned ? " aligned" : "");
			if (nprobes > 0) {
				const char *env_np = getenv("BSGPU_XSEED_NP");
				int np = xi->t.s <= 16 ? 4 : 2;
				if (env_np != nullptr && (atoi(env_np) == 4 || atoi(env_np) == 2))
					np = atoi(env_np);
				size_t smem = (size_t) BSGPU_XSEED_WARPS * BSGPU_XSEED_STAGES * xseed_stage_bytes(np, xi->t.s, aligned);
				int per_sm = 1;
				const char *env_v = getenv("BSGPU_XSEED_VARIANT");     // experiments (tools/bench_exact.py)
				int var = env_v ? atoi(env_v) : 0;
				const bool fused = getenv("BSGPU_XSEED_SPLIT_CONFIRM") == nullptr && var == 0;
#define PREPARE_XSEED(NP, AL, KB, V, FU) do { \
	static bool attr_set = false; \
	if (!attr_set) { \
		CUDA_TRY(cudaFuncSetAttribute(scan_xseed_kernel<NP, AL, KB, V, FU>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
			BSGPU_XSEED_WARPS * BSGPU_XSEED_STAGES * (int) xseed_stage_bytes(NP, AL ? 16 : 32, AL))); \
		attr_set = true; \
	} \
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_xseed_kernel<NP, AL, KB, V, FU>, BSGPU_XSEED_WARPS * 32, smem)); \
} while (0)
#define LAUNCH_XSEED(NP, AL, KB, V, FU) \
	scan_xseed_kernel<NP, AL, KB, V, FU><<<grid, BSGPU_XSEED_WARPS * 32, smem, st>>>(S, xi->t, R, first, nprobes, ws->cand.p, cap, \
								      ws->xseed_counts.p, end_lo, end_hi)
#define XSEED_DISPATCH2(WHAT, FU) do { \
	if (var == 16 && aligned && np == 4 && xi->t.kbits == 4) { WHAT(4, true, 4, 16, false); } \
	else if (xi->t.kbits == 8) { \
		if (aligned) { if (np == 4) WHAT(4, true, 8, 0, FU); else WHAT(2, true, 8, 0, FU); } \
		else { if (np == 4) WHAT(4, false, 8, 0, FU); else WHAT(2, false, 8, 0, FU); } \
	} else { \
		if (aligned) { if (np == 4) WHAT(4, true, 4, 0, FU); else WHAT(2, true, 4, 0, FU); } \
		else { if (np == 4) WHAT(4, false, 4, 0, FU); else WHAT(2, false, 4, 0, FU); } \
	} \
} while (0)
#define XSEED_DISPATCH(WHAT) do { if (fused) XSEED_DISPATCH2(WHAT, true); else XSEED_DISPATCH2(WHAT, false); } while (0)
				XSEED_DISPATCH(PREPARE_XSEED);
				int64_t ntile = (nprobes + (int64_t) np * 32 - 1) / ((int64_t) np * 32);
				int grid = (int) std::max<int64_t>(1, std::min<int64_t>(
					(ntile + BSGPU_XSEED_WARPS - 1) / BSGPU_XSEED_WARPS,
					(int64_t) g_num_sms * std::max(per_sm, 1)));
				// candidate slices: one per scan warp, sized for every probe passing both tests
				int64_t cap = xseed_warp_capacity(nprobes, np, grid);
				size_t nslices = (size_t) grid * BSGPU_XSEED_WARPS;
				if ((size_t) cap * nslices > ws->cand.n)
					TRY(ws->cand.alloc((size_t) cap * nslices));
				if (nslices > ws->xseed_counts.n)
					TRY(ws->xseed_counts.alloc(nslices));
				// (A persisting access-policy window on the filter was tried on top of the per-load
				// evict_last / evict_first hints: no gain for this kernel, and the 64 MB set aside by
				// cudaLimitPersistingL2CacheSize halved the bandwidth of every streaming kernel that ran
				// afterwards in the process -- pack_subject_kernel 0.082 -> 0.135 ms, the copy probe
				// 6.2 -> 2.9 TB/s -- so it is gone.)
				{
					KernelTimer kt("scan_xseed_kernel", st);
					XSEED_DISPATCH(LAUNCH_XSEED);
				}
				g_launches++;
				CUDA_TRY(cudaGetLastError());
				if (!fused) {
					KernelTimer kt("confirm_xseed_kernel", st);
					const char *env_c = getenv("BSGPU_XSEED_CONFIRM_BLOCKS");
					int cx = env_c != nullptr && atoi(env_c) > 0 ? atoi(env_c) : 4;
					confirm_xseed_kernel<<<dim3((unsigned) cx, (unsigned) grid), 256, 0, st>>>(
						S, xi->t, R, ws->cand.p, cap, ws->xseed_counts.p, end_lo, end_hi);
					g_launches++;
					CUDA_TRY(cudaGetLastError());
				}
#undef PREPARE_XSEED
#undef LAUNCH_XSEED
#undef XSEED_DISPATCH
#undef XSEED_DISPATCH2
			}
			return BSGPU_OK;
		}
		// exact dictionary, letters only: one probe every S letters
		const ByteSeedIndex *bi = nullptr;
		TRY(get_byteseed(p3, &bi));
		if (bi != nullptr && bi->ok) {
			int64_t S_ = bi->t.s;
			int64_t first = ((start_lo + S_ - 1) / S_) * S_;
			int64_t probe_hi = start_hi + S_ - 1;           // exclusive
			int64_t nprobes = probe_hi > first ? (probe_hi - first + S_ - 1) / S_ : 0;

			set_plan("byteseed_hash S=%d seed=%d filter_bytes=%llu buckets=%u table_bytes=%llu entries_per_bucket=%.2f",
				 bi->t.s, BSGPU_BYTESEED_LETTERS, (unsigned long long) bi->t.nfilter * 16,
				 bi->t.nbuckets, (unsigned long long) bi->t.nbuckets * 16,
				 (double) bi->nent / bi->t.nbuckets);
			if (nprobes > 0) {
				// a warp's tile = NP * 32 probes, BSGPU_BYTESEED_STAGES of them in its ring
				const char *env_np = getenv("BSGPU_BYTESEED_NP");
				int np = bi->t.s <= 11 ? 8 : (bi->t.s <= 23 ? 4 : 2);
				if (env_np != nullptr && (atoi(env_np) == 4 || atoi(env_np) == 2) && atoi(env_np) < np)
					np = atoi(env_np);
				size_t smem = (size_t) BSGPU_BYTESEED_SCAN_WARPS * BSGPU_BYTESEED_STAGES * byteseed_stage_bytes(np, bi->t.s);
				int per_sm = 1;
#define PREPARE_BYTESEED(NP) do { \
	static bool attr_set = false; \
	if (!attr_set) { \
		CUDA_TRY(cudaFuncSetAttribute(scan_byteseed_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
			BSGPU_BYTESEED_SCAN_WARPS * BSGPU_BYTESEED_STAGES * (int) byteseed_stage_bytes(NP, NP == 8 ? 11 : (NP == 4 ? 23 : 32)))); \
		attr_set = true; \
	} \
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_byteseed_kernel<NP>, 256, smem)); \
} while (0)
				if (np == 8) PREPARE_BYTESEED(8); else if (np == 4) PREPARE_BYTESEED(4); else PREPARE_BYTESEED(2);
#undef PREPARE_BYTESEED
				// one tile per scan warp at least; the grid fills the device once
				int64_t ntile = (nprobes + (int64_t) np * 32 - 1) / ((int64_t) np * 32);
				int grid = (int) std::max<int64_t>(1, std::min<int64_t>(
					(ntile + BSGPU_BYTESEED_SCAN_WARPS - 1) / BSGPU_BYTESEED_SCAN_WARPS,
					(int64_t) g_num_sms * std::max(per_sm, 1)));
#define LAUNCH_BYTESEED(NP) \
	scan_byteseed_kernel<NP><<<grid, 256, smem, st>>>(S, bi->t, R, first, nprobes, end_lo, end_hi)
				{
					KernelTimer kt("scan_byteseed_kernel", st);
					if (np == 8) LAUNCH_BYTESEED(8); else if (np == 4) LAUNCH_BYTESEED(4); else LAUNCH_BYTESEED(2);
				}
#undef LAUNCH_BYTESEED
				g_launches++;
				CUDA_TRY(cudaGetLastError());
			}
			return BSGPU_OK;
		}
	}
	bool base_windows_done = false;
	if (direct != 1) {
		// heads / tails or duplicated bands: the band's first letters address a directory and the
		// scan verifies its own hits (bsgpu_tbdir.cuh); nothing goes through the candidate buffer
		const TbDirIndex *ti = nullptr;
		TRY(get_tbdir(p3, &ti));
		if (ti != nullptr && ti->ok) {
			TRY(ensure_packed(subj, st));
			S = subj->view();
			set_plan("tb_directory w=%d prefix=%d cells=%zu slots=%zu + flank fingerprint in the slot%s", p3->w,
				 ti->v.prefix, ti->d_cells.n, ti->nslots, mp.fixedS ? "" : " + iupac_expansion");
			long long nbatches = (nwords + 31) / 32;
			{
				KernelTimer kt("scan_tbdir_kernel", st);
				if (mp.fixedS)
					scan_tbdir_kernel<false><<<grid_for(nbatches * 32, TBD_WARPS * 32, 1024), TBD_WARPS * 32, 0, st>>>(
						S, I, ti->v, R, start_lo, start_hi, mp.max_nmis, mp.min_nmis, mp.fixedP, mp.fixedS);
				else
					scan_tbdir_kernel<true><<<grid_for(nbatches * 32, TBD_WARPS * 32, 1024), TBD_WARPS * 32, 0, st>>>(
						S, I, ti->v, R, start_lo, start_hi, mp.max_nmis, mp.min_nmis, mp.fixedP, mp.fixedS);
			}
			g_launches++;
			CUDA_TRY(cudaGetLastError());
			if (mp.fixedS)
				return BSGPU_OK;
			base_windows_done = true;       // the windows with ambiguity letters follow below
		}
	}
	TRY(ensure_packed(subj, st));
	S = subj->view();
	if (!base_windows_done)
		set_plan("tb_hash w=%d buckets=%u table_bytes=%llu%s%s", p3->w, p3->nbuckets,
			 (unsigned long long) p3->nbuckets * 16,
			 direct == 1 ? "" : (direct == 2 ? " + packed flank check in the scan" : " + verify_flanks"),
			 mp.fixedS ? "" : " + iupac_expansion");
	if (!base_windows_done) {
		KernelTimer kt("scan_tb_kernel", st);
		scan_tb_kernel<<<grid_for(nwords, 256, tb_grid()), 256, 0, st>>>(S, I, R, C, direct, start_lo, start_hi,
										  mp.max_nmis, mp.min_nmis);
		g_launches++;
		CUDA_TRY(cudaGetLastError());
	}
	if (!mp.fixedS) {
		const size_t OVF_CAP = 1u << 20;
		unsigned long long n_ovf = 0;

		if (ws->ovf.n < OVF_CAP)
			TRY(ws->ovf.alloc(OVF_CAP));
		TRY(write_counter(ws, 2, 0, st));
		scan_tb_iupac_kernel<<<grid_for(nwords, 128, 2 * tb_grid()), 128, 0, st>>>(
			S, I, R, C, direct, start_lo, start_hi, iupac_max_expansions(), ws->ovf.p,
			ws->counters.p + 2, OVF_CAP);
		g_launches++;
		CUDA_TRY(cudaGetLastError());
		TRY(read_counter(ws, 2, &n_ovf, st));
		// Heavily ambiguous windows (runs of N): brute force over the Trusted-Band words, up
		// to a work budget -- the counterpart of the reference's cap on its live node set
		// (src/match_pdict_ACtree2.c:1024,1051-1054), same message.
		const char *env_b = getenv("BSGPU_IUPAC_BRUTE_BUDGET");
		double budget = env_b ? atof(env_b) : 2e10;
		if (n_ovf > OVF_CAP || (double) n_ovf * (double) p3->n_leaders > budget)
			return fail(BSGPU_EINVAL, "too many IUPAC ambiguity letters in 'subject'");
		if (n_ovf > 0 && p3->n_leaders > 0) {
			dim3 grid((unsigned) std::min<long long>((p3->n_leaders + 255) / 256, 64),
				  (unsigned) std::min<unsigned long long>(n_ovf, 65535));

			iupac_brute_kernel<<<grid, 256, 0, st>>>(S, I, R, C, direct, ws->ovf.p, n_ovf,
								  p3->d_leaders.p, p3->n_leaders);
			g_launches++;
			CUDA_TRY(cudaGetLastError());
		}
	}
	return BSGPU_OK;
}

// The automatic index of a non-preprocessed dictionary (the reference runs one
// matchPattern() per pattern, src/match_pdict.c:174-199,267-293,519-550 -- fine for 1000 patterns,
// hopeless for 10^6).  Exact matching of base-only patterns against a literally taken subject is
// what a Trusted-Band dictionary computes: band = the first min(width) letters (at most 32; the
// whole pattern when all widths agree), the rest is the tail, whole-pattern duplicates share a band
// and are reported through its group like any other keys (src/match_pdict_utils.c:153-181) -- so
// the call is handed to the preprocessed engines (extended seeds / band directory) and returns
// the same counts, which-indices and ends.  Not for chunked subjects with open ends (the two paths
// own boundary matches differently), not below BSGPU_PLAIN_INDEX_MIN patterns (default 256: the
// brute force wins), switched off by BSGPU_PLAIN_INDEX=0.
static int plain_auto_index(const bsgpu_pdict3parts *p3, int max_nmis, int min_nmis, int fixedP, int fixedS,
			    bool open_ends, const bsgpu_pdict3parts **out)
{
	*out = nullptr;
	if (!p3->plain || max_nmis != 0 || min_nmis != 0 || !fixedS || open_ends || p3->auto_state < 0)
		return BSGPU_OK;
	(void) fixedP;          // base-only patterns: fixedP does not matter
	if (p3->auto_state > 0) {
		*out = p3->auto_idx;
		return BSGPU_OK;
	}
	p3->auto_state = -1;
	const char *env = getenv("BSGPU_PLAIN_INDEX"), *env_min = getenv("BSGPU_PLAIN_INDEX_MIN");
	const int M = p3->M, min_m = env_min != nullptr ? atoi(env_min) : 256;
	if ((env != nullptr && env[0] == '0') || M < min_m || M >= (1 << 24) || p3->h_pat_off.size() != (size_t) M + 1)
		return BSGPU_OK;
	int64_t min_w = INT64_MAX, max_w = 0;
	for (int i = 0; i < M; i++) {
		int64_t wd = p3->h_pat_off[i + 1] - p3->h_pat_off[i];
		min_w = std::min(min_w, wd);
		max_w = std::max(max_w, wd);
	}
	if (min_w < 8)
		return BSGPU_OK;
	for (uint8_t b : p3->h_pat)
		if (!bsgpu_is_base(b))
			return BSGPU_OK;
	const int tbw = (int) (min_w == max_w ? min_w : std::min<int64_t>(min_w, 32));
	std::vector<uint8_t> tb((size_t) M * tbw), tail;
	std::vector<int32_t> tbwid((size_t) M, tbw), tailwid((size_t) M, 0);
	for (int i = 0; i < M; i++) {
		const uint8_t *src = p3->h_pat.data() + p3->h_pat_off[i];
		int64_t wd = p3->h_pat_off[i + 1] - p3->h_pat_off[i];

		memcpy(tb.data() + (size_t) i * tbw, src, (size_t) tbw);
		tailwid[i] = (int32_t) (wd - tbw);
		tail.insert(tail.end(), src + tbw, src + wd);
	}
	bsgpu_pdict3parts *idx = nullptr;
	const bool has_tail = max_w > tbw;
	int rc = bsgpu_pdict3parts_build_impl(BSGPU_ALGO_ACTREE2, M, tb.data(), tbwid.data(), nullptr, nullptr,
					      has_tail ? tail.data() : nullptr, has_tail ? tailwid.data() : nullptr,
					      nullptr, nullptr, &idx);
	if (rc != BSGPU_OK || idx == nullptr) {
		g_err[0] = 0;           // not an error of the call: the brute force answers it
		return BSGPU_OK;
	}
	p3->auto_idx = idx;
	p3->auto_state = 1;
	*out = idx;
	return BSGPU_OK;
}

// A non-preprocessed dictionary against the subject: every (segment, shift) slot x every
// pattern, brute force as in the reference (which runs one matchPattern per pattern).
static int run_plain(const bsgpu_pdict3parts *p3, const bsgpu_subject *subj, const MatchParams &mp,
		     BsgpuReport R, bool rec_based, Workspace *ws, cudaStream_t st)
{
	if (p3->M == 0)
		return BSGPU_OK;
	int64_t own_lo = INT64_MIN, own_hi = INT64_MAX;
	if (subj->kind == SUBJ_CHUNK) {
		// a match is owned by the chunk that holds its last letter; everything the pattern
		// touches before it must be in the halo (letters missing from a chunk are not
		// mismatches: only the ends of the whole subject are out of limits)
		int64_t need_l = (int64_t) p3->max_plen - 1;

		if (subj->open_left && subj->halo_left < need_l)
			return fail(BSGPU_EINVAL, "subject chunk has an insufficient halo: need %lld "
				    "letters before and %lld after the report range", (long long) need_l, 0ll);
		if (!subj->own_from_start)
			own_lo = subj->scan_lo;
		if (!subj->own_to_end)
			own_hi = subj->scan_hi;
	}
	int32_t n = subj->nseg;
	int ext = mp.max_nmis;
	std::vector<int64_t> slot_off((size_t) n + 1, 0);

	for (int32_t g = 0; g < n; g++)
		slot_off[g + 1] = slot_off[g] + subj->seg_len[g] + 2 * (int64_t) ext + 1;
	if (slot_off[n] >= (1ll << BSGPU_POS_BITS))
		return fail(BSGPU_EUNSUPPORTED, "subject too long for one device buffer");
	TRY(ws->slot_off.upload(slot_off.data(), slot_off.size()));
	BsgpuPlainView P = {};
	P.pat = p3->d_pat.p;
	P.pat_off = p3->d_pat_off.p;
	P.pat_head = p3->d_pat_head.p;
	P.M = p3->M;
	P.slot_off = ws->slot_off.p;
	P.ext = ext;
	P.nslots = slot_off[n];
	P.own_lo = own_lo;
	P.own_hi = own_hi;
	ws->plain = P;
	set_plan("plain_bruteforce patterns=%d slots=%lld", p3->M, (long long) P.nslots);
	// x covers the slots once, y splits the patterns so that a thread amortises its slot set-up
	// over ~64 patterns while the grid still fills the device for short subjects
	long long xblocks = std::max<long long>(1, std::min<long long>((P.nslots + 255) / 256, 1 << 20));
	long long want_y = std::max<long long>(1, (long long) g_num_sms * 16 / xblocks);
	unsigned ysplit = (unsigned) std::max<long long>(want_y, std::min<long long>((p3->M + 63) / 64, 64));
	ysplit = (unsigned) std::min<long long>(ysplit, std::max(p3->M, 1));
	dim3 grid((unsigned) xblocks, ysplit);
	for (;;) {
		unsigned long long rec_before = 0, nrec = 0;

		if (rec_based)
			rec_before = ws->rec_known;
		{
			KernelTimer kt("match_plain_kernel", st);
			match_plain_kernel<<<grid, 256, 0, st>>>(subj->view(), P, R, mp.max_nmis, mp.min_nmis,
								  mp.fixedP, mp.fixedS);
		}
		g_launches++;
		CUDA_TRY(cudaGetLastError());
		if (!rec_based)
			break;
		TRY(read_counter(ws, 1, &nrec, st));
		if (nrec <= R.capacity)
			break;
		TRY(grow(ws->rec, (size_t) nrec, (size_t) rec_before, st));
		R.records = ws->rec.p;
		R.capacity = ws->rec.n;
		TRY(write_counter(ws, 1, rec_before, st));
	}
	return BSGPU_OK;
}

// Runs one part over the subject's scan range, reporting through R.  `rec_based` tells
// whether R appends records (then the record buffer may have to grow and a slab rerun).
// own_shift: for band b of an MTB_PDict on a chunked subject the alignment is owned by the
// end of its first min_w letters, which lies own_shift letters after the band's TB end.
static int run_part(const bsgpu_pdict3parts *p3, const bsgpu_subject *subj, const MatchParams &mp,
		    BsgpuReport R, bool rec_based, Workspace *ws, cudaStream_t st, bool allow_sync,
		    int64_t own_shift = 0)
{
	if (p3->plain)
		return run_plain(p3, subj, mp, R, rec_based, ws, st);
	if (p3->M == 0 || p3->w == 0)
		return BSGPU_OK;
	if (p3->algo == BSGPU_ALGO_TWOBIT && !mp.fixedS)
		return fail(BSGPU_EINVAL, "cannot treat IUPAC extended letters in the subject "
			    "as ambiguities when 'pdict' is a PDict object of the \"Twobit\" type");
	if (subj->kind == SUBJ_CHUNK) {
		// the halo must cover every letter a reported match can touch
		int64_t need_l = (int64_t) p3->w - 1 + p3->max_H + own_shift, need_r = p3->max_T - own_shift;

		if ((subj->open_left && subj->halo_left < need_l) ||
		    (subj->open_right && subj->halo_right < need_r))
			return fail(BSGPU_EINVAL, "subject chunk has an insufficient halo: need %lld "
				    "letters before and %lld after the report range",
				    (long long) need_l, (long long) need_r);
	}
	bool direct = !p3->has_head && !p3->has_tail && p3->all_singletons;
	// every key's flanks are packed and the subject is taken literally: the scan verifies its own
	// hits (XOR + popcount), no candidate buffer and no host synchronisation
	bool inline_flanks = !direct && p3->flanks_all_packed && mp.fixedS && getenv("BSGPU_NO_PACKED_FLANKS") == nullptr;
	if (!direct && mp.fixedS && !inline_flanks) {
		// the band directory verifies any flanks itself (letter by letter where they are not packed)
		const TbDirIndex *ti = nullptr;
		TRY(get_tbdir(p3, &ti));
		inline_flanks = ti != nullptr && ti->ok;
	}
	BsgpuReport C = {};
	int64_t lo = subj->scan_lo, hi = subj->scan_hi;

	if (subj->kind == SUBJ_CHUNK && own_shift != 0) {
		if (!subj->own_from_start)
			lo -= own_shift;
		if (!subj->own_to_end)
			hi -= own_shift;
		if (hi < lo)
			hi = lo;
	}

	if (!mp.fixedS) {
		if (!subj->iupac_ready) {
			int64_t nwords = subj->total / 32;

			TRY(subj->d_iupac.alloc((size_t) nwords + 4));
			CUDA_TRY(cudaMemsetAsync(subj->d_iupac.p + nwords, 0, 4 * sizeof(uint32_t), st));
			pack_iupac_kernel<<<grid_for(2 * nwords, 256, 8), 256, 0, st>>>(subj->d_bytes.p, nwords,
											 subj->d_iupac.p);
			g_launches++;
			CUDA_TRY(cudaGetLastError());
			subj->iupac_ready = true;
		}
	}
	if ((direct || inline_flanks) && !rec_based) {
		// fully asynchronous: counts straight from the scan
		TRY(launch_scans(p3, subj, mp, R, C, direct ? 1 : 2, lo, hi, ws, st));
	} else {
		if (!allow_sync)
			return fail(BSGPU_EUNSUPPORTED, "internal: this configuration needs host synchronisation");
		int64_t a = lo;
		int64_t slab = hi - lo;

		while (a < hi) {
			int64_t b = std::min(hi, a + slab);
			unsigned long long rec_before = 0, ncand = 0, nrec = 0;

			if (rec_based)
				rec_before = ws->rec_known;
			if (direct || inline_flanks) {
				TRY(launch_scans(p3, subj, mp, R, C, direct ? 1 : 2, a, b, ws, st));
			} else {
				TRY(grow(ws->cand, CAND_MIN, 0, st));
				TRY(write_counter(ws, 0, 0, st));
				C.mode = BSGPU_R_HIT_TBPOS;
				C.records = ws->cand.p;
				C.n_records = ws->counters.p + 0;
				C.capacity = ws->cand.n;
				TRY(launch_scans(p3, subj, mp, R, C, 0, a, b, ws, st));
				TRY(read_counter(ws, 0, &ncand, st));
				if (ncand > ws->cand.n) {
					if (ncand <= CAND_MAX) {
						TRY(grow(ws->cand, (size_t) ncand, 0, st));
					} else {
						// too many candidates for one pass: shrink the slab
						slab = std::max<int64_t>(1024, (int64_t) ((double) (b - a) *
							((double) CAND_MAX / (double) ncand) * 0.9));
						TRY(grow(ws->cand, CAND_MAX, 0, st));
					}
					continue;       // rerun this slab
				}
				if (ncand > 0) {
					verify_flanks_kernel<<<grid_for((long long) ncand, 256, tb_grid()), 256, 0, st>>>(
						subj->view(), p3->view(), R, ws->cand.p, ncand,
						mp.max_nmis, mp.min_nmis, mp.fixedP, mp.fixedS);
					g_launches++;
					CUDA_TRY(cudaGetLastError());
				}
			}
			if (rec_based) {
				TRY(read_counter(ws, 1, &nrec, st));
				if (nrec > R.capacity) {
					// grow the record buffer, keep what earlier slabs wrote, redo this slab
					TRY(grow(ws->rec, (size_t) nrec, (size_t) rec_before, st));
					R.records = ws->rec.p;
					R.capacity = ws->rec.n;
					TRY(write_counter(ws, 1, rec_before, st));
					continue;
				}
			}
			a = b;
		}
	}
	return BSGPU_OK;
}

// The q-gram engine over the subject's scan range (see bsgpu_qindex.cuh).
static int run_q(const QIndex *qi, const bsgpu_subject *subj, const MatchParams &mp, BsgpuReport R,
		 bool rec_based, Workspace *ws, cudaStream_t st)
{
	const BsgpuQIndexView &Q = qi->v;
	int64_t own_lo = 0, own_hi = INT64_MAX;

	if (subj->kind == SUBJ_CHUNK) {
		int64_t need_l = (int64_t) Q.min_w - 1, need_r = (int64_t) qi->max_len - Q.min_w;

		if (Q.mode == 0) {
			need_l = qi->max_len - 1;
			need_r = 0;
		}
		if ((subj->open_left && subj->halo_left < need_l) ||
		    (subj->open_right && subj->halo_right < need_r))
			return fail(BSGPU_EINVAL, "subject chunk has an insufficient halo: need %lld "
				    "letters before and %lld after the report range",
				    (long long) need_l, (long long) need_r);
		if (!subj->own_from_start || Q.mode == 0)
			own_lo = subj->scan_lo;
		if (!subj->own_to_end || Q.mode == 0)
			own_hi = subj->scan_hi;
	}
	// every probe position that can carry the seed of an owned alignment
	int64_t probe_lo = std::max<int64_t>(0, subj->scan_lo - 64 - qi->max_shift);
	int64_t probe_hi = std::min<int64_t>(subj->total, subj->scan_hi + 64 + qi->max_shift);
	if (subj->kind != SUBJ_CHUNK) {
		probe_lo = 0;
		probe_hi = subj->total;
	}
	int64_t nwords = ((probe_hi + 31) >> 5) - (probe_lo >> 5);
	if (nwords <= 0)
		return BSGPU_OK;
	// one launch of qscan2_kernel over the probe positions [lo, hi) on stream `sx`
	auto launch_v2 = [&](int64_t lo, int64_t hi, cudaStream_t sx) {
		const char *env_g = getenv("BSGPU_Q_GRID");     // blocks per SM of the grid-stride launch
		int qgrid = env_g != nullptr && atoi(env_g) > 0 ? atoi(env_g) : 1024;
		long long nbatches = ((((hi + 31) >> 5) - (lo >> 5)) + 31) / 32;       // one batch of 32 words per warp
		KernelTimer kt("qscan2_kernel", sx);
#define QS2_LAUNCH(SH) qscan2_kernel<SH><<<grid_for(nbatches * 32, QS2_WARPS * 32, qgrid), QS2_WARPS * 32, 0, sx>>>( \
				subj->view(), Q, R, mp.max_nmis, mp.min_nmis, lo, hi, own_lo, own_hi)
		// the shapes of csrc/bsgpu_shapes.inc for windows of <= 29 letters have their key
		// extraction compiled in; any other shape goes through the generic kernel
		// (BSGPU_Q_GENERIC: tests run the generic kernel on a known shape)
		switch (getenv("BSGPU_Q_GENERIC") != nullptr ? 0ull : (unsigned long long) Q.shape1) {
		case 0x1777ull: QS2_LAUNCH(0x1777ull); break;
		case 0x3777ull: QS2_LAUNCH(0x3777ull); break;
		case 0x7777ull: QS2_LAUNCH(0x7777ull); break;
		case 0x1b6dbull: QS2_LAUNCH(0x1b6dbull); break;
		case 0x36b6bull: QS2_LAUNCH(0x36b6bull); break;
		case 0x7e3full: QS2_LAUNCH(0x7e3full); break;
		case 0x1cb17ull: QS2_LAUNCH(0x1cb17ull); break;
		case 0xe5cbull: QS2_LAUNCH(0xe5cbull); break;
		case 0x5cb97ull: QS2_LAUNCH(0x5cb97ull); break;
		case 0x165965ull: QS2_LAUNCH(0x165965ull); break;
		case 0xfa9full: QS2_LAUNCH(0xfa9full); break;
		case 0xe4ea7ull: QS2_LAUNCH(0xe4ea7ull); break;
		case 0x4c8b97ull: QS2_LAUNCH(0x4c8b97ull); break;
		case 0xFFFull: QS2_LAUNCH(0xFFFull); break;
		default: QS2_LAUNCH(0ull); break;
		}
#undef QS2_LAUNCH
		g_launches++;
	};
	// (Packing the planes in slices on a second stream while the slice before is scanned was tried:
	// pack_subject_kernel takes issue slots and L1 from the gather-bound scan, 4 x 0.375 ms of scans
	// instead of 1.377 ms -- the step went from 1.46 to 1.54 ms.)
	TRY(ensure_packed(subj, st));
	set_plan("qgram_%s q=%d shifts=%d stride=%d window=%d key_bits=%d entries=%zu %s", Q.mode ? "hamming" : "exact",
		 Q.q, Q.nshifts, Q.stride, Q.min_w, Q.key_bits, qi->nent,
		 Q.v2 ? "rank_directory+inline_patterns" : "bitmap+directory+entries");
	// The split form (emit candidates, verify one per thread) measured slower than the
	// warp-synchronous fused kernel on B200 (6.4 vs 3.7 ms per 250 Mbp); kept for experiments.
	const char *env_f = getenv("BSGPU_Q_SPLIT");
	if (Q.mode == 1 && env_f != nullptr && env_f[0] == '1') {
		// two kernels: emit candidates, then verify them one per thread
		int64_t a = probe_lo, slab = std::min<int64_t>(probe_hi - probe_lo, (int64_t) 1 << 31);

		TRY(grow(ws->cand, CAND_MIN, 0, st));
		while (a < probe_hi) {
			int64_t b = std::min(probe_hi, a + slab);
			int64_t nw = ((b + 31) >> 5) - (a >> 5);
			unsigned long long rec_before = 0, nrec = 0, ncand = 0;

			if (rec_based)
				rec_before = ws->rec_known;
			TRY(write_counter(ws, 0, 0, st));
			qprobe_kernel<<<grid_for(nw, 256, 8), 256, 0, st>>>(subj->view(), Q, a, b, ws->cand.p,
					ws->counters.p + 0, ws->cand.n);
			g_launches++;
			CUDA_TRY(cudaGetLastError());
			TRY(read_counter(ws, 0, &ncand, st));
			if (ncand > ws->cand.n) {
				if (ncand <= CAND_MAX) {
					TRY(grow(ws->cand, (size_t) ncand, 0, st));
				} else {
					slab = std::max<int64_t>(4096, (int64_t) ((double) (b - a) *
						((double) CAND_MAX / (double) ncand) * 0.9));
					TRY(grow(ws->cand, CAND_MAX, 0, st));
				}
				continue;
			}
			for (;;) {
				if (ncand > 0) {
					if (qi->max_len > 32)
						qverify_kernel<true><<<grid_for((long long) ncand, 256, 8), 256, 0, st>>>(
							subj->view(), Q, R, mp.max_nmis, mp.min_nmis, a, own_lo, own_hi,
							ws->cand.p, ncand);
					else
						qverify_kernel<false><<<grid_for((long long) ncand, 256, 8), 256, 0, st>>>(
							subj->view(), Q, R, mp.max_nmis, mp.min_nmis, a, own_lo, own_hi,
							ws->cand.p, ncand);
					g_launches++;
					CUDA_TRY(cudaGetLastError());
				}
				if (!rec_based)
					break;
				TRY(read_counter(ws, 1, &nrec, st));
				if (nrec <= R.capacity)
					break;
				TRY(grow(ws->rec, (size_t) nrec, (size_t) rec_before, st));
				R.records = ws->rec.p;
				R.capacity = ws->rec.n;
				TRY(write_counter(ws, 1, rec_before, st));
			}
			a = b;
		}
		return BSGPU_OK;
	}
	for (;;) {
		unsigned long long rec_before = 0, nrec = 0;

		if (rec_based)
			rec_before = ws->rec_known;
		{
			const char *env_g = getenv("BSGPU_Q_GRID");     // blocks per SM of the grid-stride launch
			int qgrid = env_g != nullptr && atoi(env_g) > 0 ? atoi(env_g) : 1024;
			if (Q.v2) {
				launch_v2(probe_lo, probe_hi, st);
				g_launches--;           // (counted below)
			} else if (qi->max_len > 32) {
				KernelTimer kt("qscan_kernel", st);
				qscan_kernel<true><<<grid_for(nwords, 256, qgrid), 256, 0, st>>>(subj->view(), Q, R,
						mp.max_nmis, mp.min_nmis, probe_lo, probe_hi, own_lo, own_hi);
			} else {
				KernelTimer kt("qscan_kernel", st);
				qscan_kernel<false><<<grid_for(nwords, 256, qgrid), 256, 0, st>>>(subj->view(), Q, R,
						mp.max_nmis, mp.min_nmis, probe_lo, probe_hi, own_lo, own_hi);
			}
		}
		g_launches++;
		CUDA_TRY(cudaGetLastError());
		if (!rec_based)
			break;
		TRY(read_counter(ws, 1, &nrec, st));
		if (nrec <= R.capacity)
			break;
		TRY(grow(ws->rec, (size_t) nrec, (size_t) rec_before, st));
		R.records = ws->rec.p;
		R.capacity = ws->rec.n;
		TRY(write_counter(ws, 1, rec_before, st));
	}
	return BSGPU_OK;
}

// Trusted-Band engine over all parts; for MTB bands on a chunk the report range of band b
// is shifted so that an alignment is owned by the end of its first min_w letters.
static int run_parts_tb(const bsgpu_pdict3parts *const *parts, int nparts, const bsgpu_subject *subj,
			const MatchParams &mp, BsgpuReport R, bool rec_based, Workspace *ws,
			cudaStream_t st)
{
	int64_t min_w = 0;

	for (int p = 0; p < nparts; p++)
		min_w += parts[p]->w;
	int64_t at = 0;
	for (int p = 0; p < nparts; p++) {
		at += parts[p]->w;
		if (rec_based) {
			R.records = ws->rec.p;
			R.capacity = ws->rec.n;
		}
		TRY(run_part(parts[p], subj, mp, R, rec_based, ws, st, true, nparts > 1 ? min_w - at : 0));
	}
	return BSGPU_OK;
}

static int check_args(const bsgpu_pdict3parts *const *parts, int nparts, const bsgpu_subject *subj,
		      int max_nmis, int min_nmis)
{
	if (parts == nullptr || nparts < 1 || subj == nullptr)
		return fail(BSGPU_EINVAL, "bsgpu_match(): bad arguments");
	for (int p = 0; p < nparts; p++) {
		if (parts[p] == nullptr)
			return fail(BSGPU_EINVAL, "bsgpu_match(): NULL PDict3Parts handle");
		if (parts[p]->M != parts[0]->M)
			return fail(BSGPU_EINVAL, "cannot combine MIndex objects of different lengths");
	}
	if (max_nmis < 0 || min_nmis < 0)
		return fail(BSGPU_EINVAL, "'max.mismatch' and 'min.mismatch' must be non-negative integers");
	return BSGPU_OK;
}

// sorts ws->rec[0..n) (optionally uniq) -> result in ws->rec; returns new n
static int sort_records(Workspace *ws, unsigned long long n, int end_bit, bool uniq,
			unsigned long long *n_out, cudaStream_t st)
{
	*n_out = n;
	if (n == 0)
		return BSGPU_OK;
	TRY(grow(ws->rec2, (size_t) n, 0, st));
	size_t temp = 0;
	cub::DoubleBuffer<unsigned long long> db(ws->rec.p, ws->rec2.p);
	CUDA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, temp, db, (long long) n, 0, end_bit, st));
	if (temp > ws->cub_temp.n)
		TRY(ws->cub_temp.alloc(temp));
	CUDA_TRY(cub::DeviceRadixSort::SortKeys(ws->cub_temp.p, temp, db, (long long) n, 0, end_bit, st));
	g_launches += 4;
	if (db.Current() != ws->rec.p)
		ws->rec.swap(ws->rec2);
	if (uniq) {
		TRY(grow(ws->rec2, (size_t) n, 0, st));
		temp = 0;
		unsigned long long *d_nsel = ws->counters.p + 3;
		CUDA_TRY(cub::DeviceSelect::Unique(nullptr, temp, ws->rec.p, ws->rec2.p, d_nsel, (long long) n, st));
		if (temp > ws->cub_temp.n)
			TRY(ws->cub_temp.alloc(temp));
		CUDA_TRY(cub::DeviceSelect::Unique(ws->cub_temp.p, temp, ws->rec.p, ws->rec2.p, d_nsel, (long long) n, st));
		g_launches += 2;
		ws->rec.swap(ws->rec2);
		TRY(read_counter(ws, 3, n_out, st));
	}
	return BSGPU_OK;
}

static int bits_for(unsigned long long v)
{
	int b = 1;

	while (b < 64 && (v >> b) != 0)
		b++;
	return b;
}

// The big vectors of a result (counts, CSR offsets) live in pinned host blocks that are
// recycled between calls: the device -> host copy lands in them directly (no staging copy, no
// page faults of a fresh mmap-backed malloc), which is a third of the time of a 250 Mbp call.
struct HostBlock {
	void *p;
	size_t bytes;
	bool pinned;
};
static std::vector<HostBlock> g_free_blocks, g_live_blocks;

static void *result_block_alloc(size_t bytes)
{
	for (size_t i = 0; i < g_free_blocks.size(); i++)
		if (g_free_blocks[i].bytes >= bytes && g_free_blocks[i].bytes <= 2 * bytes + 4096) {
			HostBlock blk = g_free_blocks[i];
			g_free_blocks.erase(g_free_blocks.begin() + i);
			g_live_blocks.push_back(blk);
			return blk.p;
		}
	HostBlock blk = {nullptr, bytes, false};
	if (bytes >= (1u << 16) && bytes <= (256u << 20) &&
	    cudaMallocHost(&blk.p, bytes) == cudaSuccess) {
		blk.pinned = true;
	} else {
		cudaGetLastError();
		blk.p = malloc(bytes);
	}
	if (blk.p != nullptr)
		g_live_blocks.push_back(blk);
	return blk.p;
}

static void result_block_free(void *p)
{
	if (p == nullptr)
		return;
	for (size_t i = 0; i < g_live_blocks.size(); i++)
		if (g_live_blocks[i].p == p) {
			HostBlock blk = g_live_blocks[i];
			g_live_blocks.erase(g_live_blocks.begin() + i);
			if (blk.pinned && g_free_blocks.size() < 8) {
				g_free_blocks.push_back(blk);
				return;
			}
			if (blk.pinned)
				cudaFreeHost(p);
			else
				free(p);
			return;
		}
	free(p);                // not one of ours: plain malloc
}

// offsets[i] = counts[0] + ... + counts[i-1], i = 0..M (64-bit), on the device
struct Int32To64 {
	__host__ __device__ __forceinline__ int64_t operator()(const int32_t &v) const { return (int64_t) v; }
};

extern "C" void bsgpu_result_free(bsgpu_result *res)
{
	if (res == nullptr)
		return;
	result_block_free(res->counts);
	free(res->dcounts);
	free(res->which);
	result_block_free(res->ends_offsets);
	free(res->ends);
	memset(res, 0, sizeof(*res));
}

static void which_from_counts(bsgpu_result *res, const int32_t *counts, int32_t M)
{
	int64_t n = 0;

	for (int32_t i = 0; i < M; i++)
		n += counts[i] != 0;
	res->which = (int32_t *) malloc((size_t) std::max<int64_t>(n, 1) * sizeof(int32_t));
	res->n_which = n;
	n = 0;
	for (int32_t i = 0; i < M; i++)
		if (counts[i] != 0)
			res->which[n++] = i + 1;        // sort(ids)+1, src/match_reporting.c:150-161
}

static int bsgpu_match_impl(const bsgpu_pdict3parts *const *parts, int nparts,
		const bsgpu_subject *subj, int max_nmis, int min_nmis, int fixedP, int fixedS,
		int matches_as, bsgpu_result *res)
{
	TRY(check_args(parts, nparts, subj, max_nmis, min_nmis));
	if (res == nullptr)
		return fail(BSGPU_EINVAL, "bsgpu_match(): NULL result");
	const bsgpu_pdict3parts *auto_part[1];
	if (nparts == 1 && parts[0]->plain) {
		const bsgpu_pdict3parts *ai = nullptr;
		TRY(plain_auto_index(parts[0], max_nmis, min_nmis, fixedP, fixedS,
				     subj->kind == SUBJ_CHUNK && (subj->open_left || subj->open_right), &ai));
		if (ai != nullptr) {
			auto_part[0] = ai;
			parts = auto_part;
		}
	}
	memset(res, 0, sizeof(*res));
	if (subj->kind == SUBJ_SET)
		return fail(BSGPU_EINVAL, "please use vmatchPDict() when 'subject' is an XStringSet "
			    "object (multiple sequence)");
	if (matches_as == BSGPU_MATCHES_AS_NULL)
		return BSGPU_OK;
	if (matches_as != BSGPU_MATCHES_AS_WHICH && matches_as != BSGPU_MATCHES_AS_COUNTS &&
	    matches_as != BSGPU_MATCHES_AS_ENDS)
		return fail(BSGPU_EINVAL, "\"%d\": unknown match storing mode", matches_as);
	TRY(ensure_device());
	Workspace *ws;
	TRY(workspace(&ws));
	cudaStream_t st = 0;
	MatchParams mp = {max_nmis, min_nmis, fixedP != 0, fixedS != 0};
	int32_t M = parts[0]->M;
	res->M = M;
	DevBuf<int32_t> d_counts;               // one spare element: the CSR scan reads M + 1 items
	TRY(d_counts.alloc((size_t) M + 1));
	CUDA_TRY(cudaMemsetAsync(d_counts.p, 0, ((size_t) M + 1) * sizeof(int32_t), st));
	int32_t *counts = nullptr;              // where the device -> host copy of the counts lands
	if (matches_as == BSGPU_MATCHES_AS_COUNTS) {
		res->counts = (int32_t *) result_block_alloc((size_t) std::max(M, 1) * sizeof(int32_t));
		if (res->counts == nullptr)
			return fail(BSGPU_ENOMEM, "out of host memory");
		res->n_counts = M;
		counts = res->counts;           // pinned: no staging copy
	} else {
		TRY(host_counts(ws, (size_t) M, &counts));
	}

	// q-gram engine when the dictionary is eligible (whole-pattern TB or MTB bands, base-only
	// patterns of <= 64 letters, fixedS): same results, far fewer index probes
	const QIndex *qi = nullptr;
	if (mp.fixedS)
		TRY(get_qindex(parts, nparts, &qi));
	bool useq = qi != nullptr && qi->ok;
	// with the q-gram engine every alignment is reported exactly once, so MTB counts need no
	// union -- except on views, where the reference's union also merges overlapping views
	bool need_records = matches_as == BSGPU_MATCHES_AS_ENDS ||
			    (nparts > 1 && !(useq && subj->kind != SUBJ_VIEWS));
	if (!need_records) {
		BsgpuReport R = {};
		R.mode = BSGPU_R_COUNT;
		R.M = M;
		R.counts = d_counts.p;
		if (useq)
			TRY(run_q(qi, subj, mp, R, false, ws, st));
		else
			TRY(run_parts_tb(parts, nparts, subj, mp, R, false, ws, st));
		CUDA_TRY(cudaMemcpyAsync(counts, d_counts.p, (size_t) M * sizeof(int32_t),
					 cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
	} else {
		// records: single part keeps the reference's per-view order (TB-end position);
		// several parts are merged on the end coordinate (sort + uniq,
		// src/MIndex_class.c:245-260)
		bool tbpos = nparts == 1;
		BsgpuReport R = {};
		unsigned long long nrec = 0;

		TRY(grow(ws->rec, CAND_MIN, 0, st));
		TRY(write_counter(ws, 1, 0, st));
		R.mode = tbpos ? BSGPU_R_HIT_TBPOS : BSGPU_R_HIT_COORD;
		R.M = M;
		R.records = ws->rec.p;
		R.n_records = ws->counters.p + 1;
		R.capacity = ws->rec.n;
		if (useq)
			TRY(run_q(qi, subj, mp, R, true, ws, st));
		else
			TRY(run_parts_tb(parts, nparts, subj, mp, R, true, ws, st));
		nrec = ws->rec_known;           // read back by the engine after its last slab
		// Rows without a global sort (count - scan - write, see bsgpu_kernels.cuh) whenever every
		// alignment was reported once; the union of several Trusted-Band components (duplicates
		// to drop) keeps the sort + unique path.
		const bool compact = (nparts == 1 || (useq && subj->kind != SUBJ_VIEWS)) && getenv("BSGPU_ENDS_SORT") == nullptr;
		DevBuf<int32_t> d_ends;
		DevBuf<int64_t> d_off;
		size_t temp = 0;
		thrust::transform_iterator<Int32To64, const int32_t *> in(d_counts.p, Int32To64());
		const unsigned long long *ordered = ws->rec.p;

		TRY(d_off.alloc((size_t) M + 1));
		if (compact && nrec > 0) {
			const int rgrid = grid_for((long long) nrec, 256, 8);

			TRY(grow(ws->rec2, (size_t) nrec, 0, st));
			if (ws->row_cursor.n < (size_t) M + 2)
				TRY(ws->row_cursor.alloc((size_t) M + 2));
			if (ws->big_rows.n < (size_t) M + 1)
				TRY(ws->big_rows.alloc((size_t) M + 1));
			CUDA_TRY(cudaMemsetAsync(ws->row_cursor.p, 0, ((size_t) M + 2) * sizeof(int32_t), st));
			count_rows_kernel<<<rgrid, 256, 0, st>>>(ws->rec.p, nrec, d_counts.p);
			CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, temp, in, d_off.p, M + 1, st));
			if (temp > ws->cub_temp.n)
				TRY(ws->cub_temp.alloc(temp));
			CUDA_TRY(cub::DeviceScan::ExclusiveSum(ws->cub_temp.p, temp, in, d_off.p, M + 1, st));
			write_rows_kernel<<<rgrid, 256, 0, st>>>(ws->rec.p, nrec, d_off.p, ws->row_cursor.p, ws->rec2.p);
			// row_cursor[M + 1] = number of long rows
			order_rows_kernel<<<grid_for(M, 256, 16), 256, 0, st>>>(ws->rec2.p, d_off.p, d_counts.p, M,
					ws->big_rows.p, ws->row_cursor.p + M + 1);
			bitonic_rows_kernel<<<g_num_sms * 16, 128, 0, st>>>(ws->rec2.p, d_off.p, d_counts.p, ws->big_rows.p,
									     ws->row_cursor.p + M + 1, 0u, 8192u);
			bitonic_rows_kernel<<<g_num_sms, 1024, 0, st>>>(ws->rec2.p, d_off.p, d_counts.p, ws->big_rows.p,
									 ws->row_cursor.p + M + 1, 8192u, 0xFFFFFFFFu);
			g_launches += 7;
			CUDA_TRY(cudaGetLastError());
			ordered = ws->rec2.p;
		} else if (!compact) {
			TRY(sort_records(ws, nrec, BSGPU_POS_BITS + bits_for((unsigned long long) std::max(M - 1, 1)),
					 nparts > 1, &nrec, st));
			ordered = ws->rec.p;
		}
		TRY(d_ends.alloc((size_t) nrec));
		if (nrec > 0) {
			finalize_ends_kernel<<<grid_for((long long) nrec, 256, 8), 256, 0, st>>>(
				subj->view(), parts[0]->has_tail ? parts[0]->d_tail_off.p : nullptr,
				tbpos ? 1 : 0, parts[0]->plain ? ws->plain : BsgpuPlainView{}, ordered, nrec,
				d_ends.p, compact ? nullptr : d_counts.p);
			g_launches++;
			CUDA_TRY(cudaGetLastError());
		}
		if (matches_as != BSGPU_MATCHES_AS_ENDS)        // (the CSR below says everything the counts would)
			CUDA_TRY(cudaMemcpyAsync(counts, d_counts.p, (size_t) M * sizeof(int32_t),
						 cudaMemcpyDeviceToHost, st));
		if (matches_as == BSGPU_MATCHES_AS_ENDS) {
			res->ends = (int32_t *) malloc((size_t) std::max<unsigned long long>(nrec, 1) * sizeof(int32_t));
			if (res->ends == nullptr)
				return fail(BSGPU_ENOMEM, "out of host memory");
			if (nrec > 0)
				CUDA_TRY(cudaMemcpyAsync(res->ends, d_ends.p, (size_t) nrec * sizeof(int32_t),
							 cudaMemcpyDeviceToHost, st));
			// CSR offsets = exclusive scan of the counts, on the device, straight into the
			// result's pinned block
			if (!(compact && nrec > 0)) {
				CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, temp, in, d_off.p, M + 1, st));
				if (temp > ws->cub_temp.n)
					TRY(ws->cub_temp.alloc(temp));
				CUDA_TRY(cub::DeviceScan::ExclusiveSum(ws->cub_temp.p, temp, in, d_off.p, M + 1, st));
				g_launches += 2;
			}
			res->n_rows = M;
			res->ends_offsets = (int64_t *) result_block_alloc(((size_t) M + 1) * sizeof(int64_t));
			if (res->ends_offsets == nullptr)
				return fail(BSGPU_ENOMEM, "out of host memory");
			CUDA_TRY(cudaMemcpyAsync(res->ends_offsets, d_off.p, ((size_t) M + 1) * sizeof(int64_t),
						 cudaMemcpyDeviceToHost, st));
			CUDA_TRY(cudaStreamSynchronize(st));
			return BSGPU_OK;
		}
		CUDA_TRY(cudaStreamSynchronize(st));
	}
	if (matches_as != BSGPU_MATCHES_AS_COUNTS)
		which_from_counts(res, counts, M);
	return BSGPU_OK;
}

extern "C" int bsgpu_match(const bsgpu_pdict3parts *const *parts, int nparts,
		const bsgpu_subject *subj, int max_nmis, int min_nmis, int fixedP, int fixedS,
		int matches_as, bsgpu_result *res)
{
	int rc;

	try {
		rc = bsgpu_match_impl(parts, nparts, subj, max_nmis, min_nmis, fixedP, fixedS, matches_as, res);
	} catch (const std::bad_alloc &) {
		rc = fail(BSGPU_ENOMEM, "bsgpu_match(): out of host memory");
	} catch (const std::exception &e) {
		rc = fail(BSGPU_ECUDA, "bsgpu_match(): %s", e.what());
	}
	if (rc != BSGPU_OK && res != nullptr)
		bsgpu_result_free(res);        // nothing of a failed call stays allocated
	return rc;
}


static int bsgpu_count_device_impl(const bsgpu_pdict3parts *const *parts, int nparts,
		const bsgpu_subject *subj, int max_nmis, int min_nmis, int fixedP, int fixedS,
		int32_t *d_counts, void *stream, int *n_launches_out)
{
	TRY(check_args(parts, nparts, subj, max_nmis, min_nmis));
	if (d_counts == nullptr)
		return fail(BSGPU_EINVAL, "bsgpu_count_device(): NULL d_counts");
	const bsgpu_pdict3parts *auto_part[1];
	if (nparts == 1 && parts[0]->plain) {
		const bsgpu_pdict3parts *ai = nullptr;
		TRY(plain_auto_index(parts[0], max_nmis, min_nmis, fixedP, fixedS,
				     subj->kind == SUBJ_CHUNK && (subj->open_left || subj->open_right), &ai));
		if (ai != nullptr) {
			auto_part[0] = ai;
			parts = auto_part;
		}
	}
	TRY(ensure_device());
	Workspace *ws;
	TRY(workspace(&ws));
	cudaStream_t st = (cudaStream_t) stream;
	MatchParams mp = {max_nmis, min_nmis, fixedP != 0, fixedS != 0};
	long long before = g_launches;
	int32_t M = parts[0]->M;

	const QIndex *qi = nullptr;
	if (mp.fixedS)
		TRY(get_qindex(parts, nparts, &qi));
	bool useq = qi != nullptr && qi->ok;
	if (nparts == 1 || (useq && subj->kind != SUBJ_VIEWS)) {
		BsgpuReport R = {};
		R.mode = BSGPU_R_COUNT;
		R.M = M;
		R.counts = d_counts;
		if (useq)
			TRY(run_q(qi, subj, mp, R, false, ws, st));
		else
			TRY(run_parts_tb(parts, nparts, subj, mp, R, false, ws, st));
	} else {
		BsgpuReport R = {};
		unsigned long long nrec = 0;

		TRY(grow(ws->rec, CAND_MIN, 0, st));
		TRY(write_counter(ws, 1, 0, st));
		R.mode = BSGPU_R_HIT_COORD;
		R.M = M;
		R.records = ws->rec.p;
		R.n_records = ws->counters.p + 1;
		R.capacity = ws->rec.n;
		if (useq)
			TRY(run_q(qi, subj, mp, R, true, ws, st));
		else
			TRY(run_parts_tb(parts, nparts, subj, mp, R, true, ws, st));
		nrec = ws->rec_known;           // read back by the engine after its last slab
		TRY(sort_records(ws, nrec, BSGPU_POS_BITS + bits_for((unsigned long long) std::max(M - 1, 1)),
				 true, &nrec, st));
		if (nrec > 0) {
			DevBuf<int32_t> d_ends;
			TRY(d_ends.alloc((size_t) nrec));
			finalize_ends_kernel<<<grid_for((long long) nrec, 256, 8), 256, 0, st>>>(
				subj->view(), nullptr, 0, BsgpuPlainView{}, ws->rec.p, nrec, d_ends.p, d_counts);
			g_launches++;
			CUDA_TRY(cudaGetLastError());
			CUDA_TRY(cudaStreamSynchronize(st));
		}
	}
	if (n_launches_out)
		*n_launches_out = (int) (g_launches - before);
	return BSGPU_OK;
}

extern "C" int bsgpu_count_device(const bsgpu_pdict3parts *const *parts, int nparts,
		const bsgpu_subject *subj, int max_nmis, int min_nmis, int fixedP, int fixedS,
		int32_t *d_counts, void *stream, int *n_launches_out)
{
	int rc;

	try {
		rc = bsgpu_count_device_impl(parts, nparts, subj, max_nmis, min_nmis, fixedP, fixedS, d_counts, stream, n_launches_out);
	} catch (const std::bad_alloc &) {
		rc = fail(BSGPU_ENOMEM, "bsgpu_count_device(): out of host memory");
	} catch (const std::exception &e) {
		rc = fail(BSGPU_ECUDA, "bsgpu_count_device(): %s", e.what());
	}
	return rc;
}


// ---------------------------------------------------------------------------------------
// countPDict on a subject that lives in HOST memory: the upload is cut into slices and slice
// k is scanned while slice k + 1 crosses PCIe, so a call costs the copy plus one slice of
// scanning instead of copy + scan.  The subject lands in a staging buffer the library keeps
// between calls (no allocation, no guard fill when the length repeats); the counts leave the
// device once, after the caller's reduction if there is one (d_counts).
// ---------------------------------------------------------------------------------------

struct StageState {
	bsgpu_subject *subj = nullptr;
	cudaStream_t copy = nullptr;
	std::vector<cudaEvent_t> arrived;       // slice k is on the device
	cudaEvent_t scanned = nullptr;          // the previous call is done with the staging buffer
	bool scanned_pending = false;
	DevBuf<int32_t> d_counts;
};
static StageState g_stage;

static int64_t stage_slice_letters(void)
{
	const char *env = getenv("BSGPU_HOST_SLICE_MB");
	double mb = env != nullptr && atof(env) > 0 ? atof(env) : 32.0;

	return std::max<int64_t>(1 << 16, ((int64_t) (mb * 1048576.0)) & ~(int64_t) 4095);
}

// the range-limited counting step shared by the slices: true when the dictionary's engine can
// count straight into d_counts (no record buffer, no union)
static bool counts_directly(const bsgpu_pdict3parts *const *parts, int nparts, const bsgpu_subject *subj,
			    const QIndex *qi)
{
	bool useq = qi != nullptr && qi->ok;

	return nparts == 1 || (useq && subj->kind != SUBJ_VIEWS);
}

static int bsgpu_count_host_impl(const bsgpu_pdict3parts *const *parts, int nparts,
		const uint8_t *letters, int64_t length, int64_t chunk_start, int64_t subject_len,
		int64_t report_lo, int64_t report_hi,
		int max_nmis, int min_nmis, int fixedP, int fixedS,
		int32_t *counts_out, int32_t *d_counts, void *stream)
{
	if (parts == nullptr || nparts < 1 || parts[0] == nullptr || length < 0 || (length > 0 && letters == nullptr) ||
	    chunk_start < 0 || chunk_start + length > subject_len || length > INT32_MAX)
		return fail(BSGPU_EINVAL, "bsgpu_count_host(): bad arguments");
	if (max_nmis < 0 || min_nmis < 0)
		return fail(BSGPU_EINVAL, "'max.mismatch' and 'min.mismatch' must be non-negative integers");
	for (int p = 0; p < nparts; p++)
		if (parts[p] == nullptr || parts[p]->M != parts[0]->M)
			return fail(BSGPU_EINVAL, "cannot combine MIndex objects of different lengths");
	TRY(ensure_device());
	const bsgpu_pdict3parts *auto_part[1];
	if (nparts == 1 && parts[0]->plain) {
		const bsgpu_pdict3parts *ai = nullptr;
		TRY(plain_auto_index(parts[0], max_nmis, min_nmis, fixedP, fixedS,
				     chunk_start > 0 || chunk_start + length < subject_len, &ai));
		if (ai != nullptr) {
			auto_part[0] = ai;
			parts = auto_part;
		}
	}
	Workspace *ws;
	TRY(workspace(&ws));
	cudaStream_t st = (cudaStream_t) stream;
	MatchParams mp = {max_nmis, min_nmis, fixedP != 0, fixedS != 0};
	const int32_t M = parts[0]->M;
	StageState &G = g_stage;

	if (G.copy == nullptr) {
		CUDA_TRY(cudaStreamCreateWithFlags(&G.copy, cudaStreamNonBlocking));
		CUDA_TRY(cudaEventCreateWithFlags(&G.scanned, cudaEventDisableTiming));
	}
	// ---- the staging subject: rebuilt only when the length changes -------------------------
	report_lo = std::max(report_lo, chunk_start);
	report_hi = std::min(report_hi, chunk_start + length);
	if (report_hi < report_lo)
		report_hi = report_lo;
	if (G.subj == nullptr || G.subj->seg_len[0] != (int32_t) length) {
		if (G.scanned_pending)
			CUDA_TRY(cudaEventSynchronize(G.scanned));
		G.scanned_pending = false;
		delete G.subj;
		G.subj = nullptr;
		std::unique_ptr<bsgpu_subject> s(new bsgpu_subject());
		s->kind = SUBJ_CHUNK;
		s->nseg = 1;
		s->seg_len.assign(1, (int32_t) length);
		s->seg_coord.assign(1, chunk_start);
		TRY(subject_finish(s.get(), nullptr, nullptr));         // guards only; the letters come in slices
		G.subj = s.release();
	}
	bsgpu_subject *s = G.subj;
	if (s->seg_coord[0] != chunk_start) {
		s->seg_coord[0] = chunk_start;
		CUDA_TRY(cudaMemcpyAsync(s->d_seg_coord.p, s->seg_coord.data(), sizeof(int64_t), cudaMemcpyHostToDevice, st));
	}
	const int64_t seg0 = s->seg_start[0];
	const int64_t all_lo = seg0 + (report_lo - chunk_start), all_hi = seg0 + (report_hi - chunk_start);
	const bool from_start = report_lo <= 0, to_end = report_hi >= subject_len;
	const bool open_l = chunk_start > 0, open_r = chunk_start + length < subject_len;
	const int64_t halo_l = report_lo - chunk_start, halo_r = chunk_start + length - report_hi;
	s->nletters = report_hi - report_lo;
	s->packed = false;
	s->iupac_ready = false;

	const QIndex *qi = nullptr;
	if (mp.fixedS)
		TRY(get_qindex(parts, nparts, &qi));
	bool useq = qi != nullptr && qi->ok;
	if (d_counts == nullptr) {
		if (G.d_counts.n < (size_t) M + 1)
			TRY(G.d_counts.alloc((size_t) M + 1));
		d_counts = G.d_counts.p;
		CUDA_TRY(cudaMemsetAsync(d_counts, 0, ((size_t) M + 1) * sizeof(int32_t), st));
	}
	// letters a reported match may touch behind its owner position (tails, the long end of
	// variable-width patterns, the words the packed scans read ahead)
	int64_t margin = 256;
	for (int p = 0; p < nparts; p++)
		margin = std::max<int64_t>(margin, 256 + parts[p]->max_T + parts[p]->max_plen);
	const int64_t SL = std::max(stage_slice_letters(), 8 * margin);
	const int nslices = (int) std::max<int64_t>(1, (length + SL - 1) / SL);
	// fixedS = FALSE builds its IUPAC plane over the whole subject on first use: no slices there
	const bool pipelined = counts_directly(parts, nparts, s, qi) && nslices > 1 && (mp.fixedS || parts[0]->plain) &&
			       getenv("BSGPU_HOST_NO_PIPELINE") == nullptr;

	// the previous call's scans must be done with the buffer before it is overwritten
	if (G.scanned_pending)
		CUDA_TRY(cudaStreamWaitEvent(G.copy, G.scanned, 0));
	while ((int) G.arrived.size() < nslices) {
		cudaEvent_t e;
		CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		G.arrived.push_back(e);
	}
	for (int k = 0; k < nslices; k++) {
		int64_t a = (int64_t) k * SL, b = std::min(length, a + SL);

		CUDA_TRY(cudaMemcpyAsync(s->d_bytes.p + seg0 + a, letters + a, (size_t) (b - a), cudaMemcpyHostToDevice, G.copy));
		CUDA_TRY(cudaEventRecord(G.arrived[k], G.copy));
	}
	int rc = BSGPU_OK;
	if (!pipelined) {
		// whole subject at once (records / union needed, or a short subject)
		CUDA_TRY(cudaStreamWaitEvent(st, G.arrived[nslices - 1], 0));
		s->scan_lo = all_lo;
		s->scan_hi = all_hi;
		s->open_left = open_l;
		s->open_right = open_r;
		s->halo_left = halo_l;
		s->halo_right = halo_r;
		s->own_from_start = from_start;
		s->own_to_end = to_end;
		rc = bsgpu_count_device(parts, nparts, s, max_nmis, min_nmis, fixedP, fixedS, d_counts, st, nullptr);
	} else {
		BsgpuReport R = {};
		R.mode = BSGPU_R_COUNT;
		R.M = M;
		R.counts = d_counts;
		int64_t lo = all_lo, packed_words = 0;
		const int64_t nwords = s->total / 32;
		// the exact engines read the letters themselves, everything else the packed planes
		bool letters_only = parts[0]->plain;
		if (!useq && nparts == 1 && !parts[0]->plain && !parts[0]->has_head && !parts[0]->has_tail &&
		    parts[0]->all_singletons && mp.fixedS) {
			const XSeedIndex *xi = nullptr;
			const ByteSeedIndex *bi = nullptr;
			TRY(get_xseed(parts[0], &xi));
			if (xi == nullptr || !xi->ok)
				TRY(get_byteseed(parts[0], &bi));
			letters_only = (xi != nullptr && xi->ok) || (bi != nullptr && bi->ok);
		}
		const bool need_planes = !letters_only;

		if (need_planes && s->d_codes.n < (size_t) nwords + 4) {
			TRY(s->d_codes.alloc((size_t) nwords + 4));
			TRY(s->d_valid.alloc((size_t) nwords + 4));
			TRY(s->d_inv2.alloc((size_t) nwords + 4));
			CUDA_TRY(cudaMemsetAsync(s->d_codes.p + nwords, 0, 4 * sizeof(uint64_t), st));
			CUDA_TRY(cudaMemsetAsync(s->d_valid.p + nwords, 0, 4 * sizeof(uint32_t), st));
			CUDA_TRY(cudaMemsetAsync(s->d_inv2.p + nwords, 0xFF, 4 * sizeof(uint64_t), st));
		}
		for (int k = 0; k < nslices && rc == BSGPU_OK; k++) {
			const bool last = k == nslices - 1;
			int64_t b = std::min(length, (int64_t) (k + 1) * SL);
			int64_t hi = last ? all_hi : std::min(all_hi, std::max(lo, seg0 + b - margin));

			CUDA_TRY(cudaStreamWaitEvent(st, G.arrived[k], 0));
			if (need_planes) {
				// planes of the words whose letters are all here (the last slice takes the guard words too)
				int64_t w1 = last ? nwords : (seg0 + b) / 32;

				if (w1 > packed_words) {
					KernelTimer kt("pack_subject_kernel", st);
					pack_subject_kernel<<<grid_for(2 * (w1 - packed_words), 256, 8), 256, 0, st>>>(
						s->d_bytes.p + 32 * packed_words, w1 - packed_words, s->d_codes.p + packed_words,
						s->d_valid.p + packed_words, s->d_inv2.p + packed_words);
					g_launches++;
					CUDA_TRY(cudaGetLastError());
					packed_words = w1;
				}
				s->packed = true;
			}
			if (hi <= lo && !last)
				continue;
			s->scan_lo = lo;
			s->scan_hi = hi;
			s->own_from_start = from_start && k == 0;
			s->own_to_end = to_end && last;
			s->open_left = k == 0 ? open_l : true;
			s->open_right = last ? open_r : true;
			s->halo_left = k == 0 ? halo_l : (int64_t) 1 << 40;
			s->halo_right = last ? halo_r : (int64_t) 1 << 40;
			if (useq)
				rc = run_q(qi, s, mp, R, false, ws, st);
			else
				rc = run_parts_tb(parts, nparts, s, mp, R, false, ws, st);
			lo = hi;
		}
		s->packed = false;
	}
	CUDA_TRY(cudaEventRecord(G.scanned, st));
	G.scanned_pending = true;
	if (rc != BSGPU_OK)
		return rc;
	if (counts_out != nullptr) {
		CUDA_TRY(cudaMemcpyAsync(counts_out, d_counts, (size_t) M * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
	}
	return BSGPU_OK;
}

extern "C" int bsgpu_count_host(const bsgpu_pdict3parts *const *parts, int nparts,
		const uint8_t *letters, int64_t length, int64_t chunk_start, int64_t subject_len,
		int64_t report_lo, int64_t report_hi,
		int max_nmis, int min_nmis, int fixedP, int fixedS,
		int32_t *counts_out, int32_t *d_counts, void *stream)
{
	int rc;

	try {
		rc = bsgpu_count_host_impl(parts, nparts, letters, length, chunk_start, subject_len, report_lo, report_hi, max_nmis, min_nmis, fixedP, fixedS, counts_out, d_counts, stream);
	} catch (const std::bad_alloc &) {
		rc = fail(BSGPU_ENOMEM, "bsgpu_count_host(): out of host memory");
	} catch (const std::exception &e) {
		rc = fail(BSGPU_ECUDA, "bsgpu_count_host(): %s", e.what());
	}
	return rc;
}


// ---------------------------------------------------------------------------------------
// vmatch: vcountPDict / vwhichPDict on a set of subjects (src/match_pdict.c:320-346,386-432)
// ---------------------------------------------------------------------------------------

static int bsgpu_vmatch_impl(const bsgpu_pdict3parts *const *parts, int nparts,
		const bsgpu_subject *subj, int max_nmis, int min_nmis, int fixedP, int fixedS,
		int collapse, int weight_is_double, const void *weight, int64_t weight_length,
		int matches_as, bsgpu_result *res)
{
	TRY(check_args(parts, nparts, subj, max_nmis, min_nmis));
	if (res == nullptr)
		return fail(BSGPU_EINVAL, "bsgpu_vmatch(): NULL result");
	const bsgpu_pdict3parts *auto_part[1];
	if (nparts == 1 && parts[0]->plain) {
		const bsgpu_pdict3parts *ai = nullptr;
		TRY(plain_auto_index(parts[0], max_nmis, min_nmis, fixedP, fixedS,
				     subj->kind == SUBJ_CHUNK && (subj->open_left || subj->open_right), &ai));
		if (ai != nullptr) {
			auto_part[0] = ai;
			parts = auto_part;
		}
	}
	memset(res, 0, sizeof(*res));
	if (subj->kind != SUBJ_SET)
		return fail(BSGPU_EINVAL, "please use matchPDict() when 'subject' is an XString or "
			    "MaskedXString object (single sequence)");
	if (matches_as == BSGPU_MATCHES_AS_NULL)
		return fail(BSGPU_EINVAL, "vmatch_PDict3Parts_XStringSet() does not support "
			    "'matches_as=\"MATCHES_AS_NULL\"' yet, sorry");
	if (matches_as != BSGPU_MATCHES_AS_WHICH && matches_as != BSGPU_MATCHES_AS_COUNTS)
		return fail(BSGPU_EINVAL, "vmatchPDict() is not supported yet, sorry");
	if (nparts > 1 && matches_as != BSGPU_MATCHES_AS_WHICH)
		return fail(BSGPU_EINVAL, "MTB_PDict objects are not supported yet, sorry");
	TRY(ensure_device());
	Workspace *ws;
	TRY(workspace(&ws));
	cudaStream_t st = 0;
	MatchParams mp = {max_nmis, min_nmis, fixedP != 0, fixedS != 0};
	int32_t M = parts[0]->M;
	int64_t N = subj->nseg;
	res->M = M;
	const QIndex *qi = nullptr;
	if (mp.fixedS)
		TRY(get_qindex(parts, nparts, &qi));
	bool useq = qi != nullptr && qi->ok;

	bool records = matches_as == BSGPU_MATCHES_AS_WHICH || (collapse != 0 && weight_is_double);
	if (matches_as == BSGPU_MATCHES_AS_COUNTS) {
		if (collapse < 0 || collapse > 2)
			return fail(BSGPU_EINVAL, "'collapse' must be FALSE, 1 or 2");
		if (collapse != 0) {
			int64_t need = collapse == 1 ? N : M;

			if (weight == nullptr || weight_length != need)
				return fail(BSGPU_EINVAL, "'weight' must have length %lld", (long long) need);
		}
	}
	if (!records) {
		int64_t n_out = collapse == 0 ? (int64_t) M * N : (collapse == 1 ? M : N);
		DevBuf<int32_t> d_out, d_w;
		BsgpuReport R = {};

		if (collapse == 0 && n_out > (int64_t) INT32_MAX)
			return fail(BSGPU_EINVAL, "allocMatrix: too many elements specified");
		TRY(d_out.alloc((size_t) n_out));
		CUDA_TRY(cudaMemsetAsync(d_out.p, 0, (size_t) std::max<int64_t>(n_out, 1) * sizeof(int32_t), st));
		R.M = M;
		R.counts = d_out.p;
		if (collapse == 0) {
			R.mode = BSGPU_R_VMAT;
		} else {
			TRY(d_w.upload((const int32_t *) weight, (size_t) weight_length));
			R.mode = collapse == 1 ? BSGPU_R_VC1 : BSGPU_R_VC2;
			R.weight = d_w.p;
		}
		if (useq)
			TRY(run_q(qi, subj, mp, R, false, ws, st));
		else
			TRY(run_part(parts[0], subj, mp, R, false, ws, st, true));
		res->counts = (int32_t *) malloc((size_t) std::max<int64_t>(n_out, 1) * sizeof(int32_t));
		res->n_counts = n_out;
		if (n_out > 0)
			CUDA_TRY(cudaMemcpyAsync(res->counts, d_out.p, (size_t) n_out * sizeof(int32_t),
						 cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		return BSGPU_OK;
	}
	// record-based: (segment << 30 | key), sorted
	BsgpuReport R = {};
	unsigned long long nrec = 0;
	TRY(grow(ws->rec, CAND_MIN, 0, st));
	TRY(write_counter(ws, 1, 0, st));
	R.mode = BSGPU_R_SEGKEY;
	R.M = M;
	R.records = ws->rec.p;
	R.n_records = ws->counters.p + 1;
	R.capacity = ws->rec.n;
	if (useq)
		TRY(run_q(qi, subj, mp, R, true, ws, st));
	else
		TRY(run_parts_tb(parts, nparts, subj, mp, R, true, ws, st));
	nrec = ws->rec_known;
	bool uniq = matches_as == BSGPU_MATCHES_AS_WHICH;
	TRY(sort_records(ws, nrec, BSGPU_SEGKEY_BITS + bits_for((unsigned long long) std::max<int64_t>(N - 1, 1)),
			 uniq, &nrec, st));
	std::vector<unsigned long long> h((size_t) nrec);
	if (nrec > 0)
		CUDA_TRY(cudaMemcpy(h.data(), ws->rec.p, (size_t) nrec * sizeof(unsigned long long),
				    cudaMemcpyDeviceToHost));
	const unsigned long long keymask = (1ull << BSGPU_SEGKEY_BITS) - 1;
	if (matches_as == BSGPU_MATCHES_AS_WHICH) {
		res->n_rows = N;
		res->ends_offsets = (int64_t *) calloc((size_t) N + 1, sizeof(int64_t));
		res->ends = (int32_t *) malloc((size_t) std::max<unsigned long long>(nrec, 1) * sizeof(int32_t));
		for (size_t t = 0; t < h.size(); t++) {
			res->ends_offsets[(h[t] >> BSGPU_SEGKEY_BITS) + 1]++;
			res->ends[t] = (int32_t) (h[t] & keymask) + 1;
		}
		for (int64_t j = 0; j < N; j++)
			res->ends_offsets[j + 1] += res->ends_offsets[j];
		return BSGPU_OK;
	}
	// double weights: fold on the host in the reference's order (subject-major, pattern
	// inner; src/match_pdict.c:409-427) from exact integer counts
	const double *wd = (const double *) weight;
	int64_t n_out = collapse == 1 ? M : N;
	res->dcounts = (double *) calloc((size_t) std::max<int64_t>(n_out, 1), sizeof(double));
	res->n_counts = n_out;
	for (size_t t = 0; t < h.size();) {
		size_t u = t;

		while (u < h.size() && h[u] == h[t])
			u++;
		int64_t j = (int64_t) (h[t] >> BSGPU_SEGKEY_BITS), i = (int64_t) (h[t] & keymask);
		int cnt = (int) (u - t);
		if (collapse == 1)
			res->dcounts[i] += cnt * wd[j];
		else
			res->dcounts[j] += cnt * wd[i];
		t = u;
	}
	return BSGPU_OK;
}

extern "C" int bsgpu_vmatch(const bsgpu_pdict3parts *const *parts, int nparts,
		const bsgpu_subject *subj, int max_nmis, int min_nmis, int fixedP, int fixedS,
		int collapse, int weight_is_double, const void *weight, int64_t weight_length,
		int matches_as, bsgpu_result *res)
{
	int rc;

	try {
		rc = bsgpu_vmatch_impl(parts, nparts, subj, max_nmis, min_nmis, fixedP, fixedS, collapse, weight_is_double, weight, weight_length, matches_as, res);
	} catch (const std::bad_alloc &) {
		rc = fail(BSGPU_ENOMEM, "bsgpu_vmatch(): out of host memory");
	} catch (const std::exception &e) {
		rc = fail(BSGPU_ECUDA, "bsgpu_vmatch(): %s", e.what());
	}
	if (rc != BSGPU_OK && res != nullptr)
		bsgpu_result_free(res);        // nothing of a failed call stays allocated
	return rc;
}


// ---------------------------------------------------------------------------------------
// Host-buffer entry points with the reference's .Call shapes
// ---------------------------------------------------------------------------------------

extern "C" int bsgpu_match_PDict3Parts_XString(const bsgpu_pdict3parts *p3,
		const uint8_t *subject, int64_t subject_length,
		int max_nmis, int min_nmis, int fixedP, int fixedS, int matches_as, bsgpu_result *res)
{
	bsgpu_subject *s = nullptr;

	TRY(bsgpu_subject_xstring(subject, subject_length, &s));
	int rc = bsgpu_match(&p3, 1, s, max_nmis, min_nmis, fixedP, fixedS, matches_as, res);
	bsgpu_subject_free(s);
	return rc;
}

extern "C" int bsgpu_match_PDict3Parts_XStringViews(const bsgpu_pdict3parts *p3,
		const uint8_t *subject, int64_t subject_length,
		const int32_t *views_start, const int32_t *views_width, int32_t nviews,
		int max_nmis, int min_nmis, int fixedP, int fixedS, int matches_as, bsgpu_result *res)
{
	bsgpu_subject *s = nullptr;

	TRY(bsgpu_subject_views(subject, subject_length, views_start, views_width, nviews, &s));
	int rc = bsgpu_match(&p3, 1, s, max_nmis, min_nmis, fixedP, fixedS, matches_as, res);
	bsgpu_subject_free(s);
	return rc;
}

extern "C" int bsgpu_vmatch_PDict3Parts_XStringSet(const bsgpu_pdict3parts *p3,
		const uint8_t *subject_concat, const int32_t *subject_widths, int32_t nsubject,
		int max_nmis, int min_nmis, int fixedP, int fixedS,
		int collapse, int weight_is_double, const void *weight, int64_t weight_length,
		int matches_as, bsgpu_result *res)
{
	bsgpu_subject *s = nullptr;

	TRY(bsgpu_subject_set(subject_concat, subject_widths, nsubject, &s));
	int rc = bsgpu_vmatch(&p3, 1, s, max_nmis, min_nmis, fixedP, fixedS, collapse,
			      weight_is_double, weight, weight_length, matches_as, res);
	bsgpu_subject_free(s);
	return rc;
}

// ByPos_MIndex_combine (src/MIndex_class.c:230-263): per row, sorted distinct union.
extern "C" int bsgpu_mindex_combine(const bsgpu_result *const *comps, int ncomps, bsgpu_result *res)
{
	if (comps == nullptr || ncomps < 1 || res == nullptr)
		return fail(BSGPU_EINVAL, "nothing to combine");
	memset(res, 0, sizeof(*res));
	int64_t rows = comps[0]->n_rows;
	for (int j = 1; j < ncomps; j++)
		if (comps[j]->n_rows != rows)
			return fail(BSGPU_EINVAL, "cannot combine MIndex objects of different lengths");
	std::vector<int32_t> all, buf;
	std::vector<int64_t> off((size_t) rows + 1, 0);
	for (int64_t i = 0; i < rows; i++) {
		buf.clear();
		for (int j = 0; j < ncomps; j++)
			buf.insert(buf.end(), comps[j]->ends + comps[j]->ends_offsets[i],
				   comps[j]->ends + comps[j]->ends_offsets[i + 1]);
		std::sort(buf.begin(), buf.end());
		buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
		all.insert(all.end(), buf.begin(), buf.end());
		off[i + 1] = (int64_t) all.size();
	}
	res->M = comps[0]->M;
	res->n_rows = rows;
	res->ends_offsets = (int64_t *) malloc(((size_t) rows + 1) * sizeof(int64_t));
	memcpy(res->ends_offsets, off.data(), ((size_t) rows + 1) * sizeof(int64_t));
	res->ends = (int32_t *) malloc(std::max<size_t>(all.size(), 1) * sizeof(int32_t));
	memcpy(res->ends, all.data(), all.size() * sizeof(int32_t));
	return BSGPU_OK;
}

// ---------------------------------------------------------------------------------------
// Matching with indels on a non-preprocessed dictionary (countPDict / whichPDict /
// vcountPDict / vwhichPDict with with.indels=TRUE: _match_pattern_indels,
// src/match_pattern_indels.c:58-107).  The per-position core (bsgpu_indels.cuh) is checked
// against the reference on the CPU (tests/test_indels_oracle.py), the whole engine on the device
// (tests/test_gpu_indels.py).
// ---------------------------------------------------------------------------------------

// first[pattern][y] = first pattern position whose letter matches subject byte y, or -1
__global__ void __launch_bounds__(256)
indels_first_kernel(const uint8_t *__restrict__ pat, const int64_t *__restrict__ pat_off, int M,
		    int fixedP, int fixedS, int16_t *__restrict__ first)
{
	int pid = blockIdx.x, y = threadIdx.x;

	if (pid >= M)
		return;
	const uint8_t *P = pat + pat_off[pid];
	int plen = (int) (pat_off[pid + 1] - pat_off[pid]), at = -1;

	for (int n = 0; n < plen; n++)
		if (bsgpu_indels_match(P[n], (uint32_t) y, fixedP, fixedS)) {
			at = n;
			break;
		}
	first[(size_t) pid * 256 + y] = (int16_t) at;
}

// One thread per (letter of the subject, pattern of this grid row): the candidate match that
// starts at the letter, if any, as {pattern << 34 | concat index, width << 8 | nedit}.
__global__ void __launch_bounds__(128)
indels_candidates_kernel(BsgpuSubjectView S, const uint8_t *__restrict__ pat, const int64_t *__restrict__ pat_off,
			 int M, const int16_t *__restrict__ first, int max_nmis, int fixedP, int fixedS,
			 int64_t scan_lo, int64_t scan_hi, unsigned long long *__restrict__ keys,
			 uint32_t *__restrict__ vals, unsigned long long *__restrict__ n_out,
			 unsigned long long capacity)
{
	int64_t stride = (int64_t) gridDim.x * blockDim.x;

	for (int64_t p = scan_lo + (int64_t) blockIdx.x * blockDim.x + threadIdx.x; p < scan_hi; p += stride) {
		int seg = S.nseg == 1 ? 0 : find_segment(S, p);
		int64_t lo = __ldg(S.seg_start + seg), L = __ldg(S.seg_len + seg);

		if (p < lo || p >= lo + L)
			continue;                       // a separator
		for (int pid = blockIdx.y; pid < M; pid += gridDim.y) {
			int64_t o = __ldg(pat_off + pid);
			int plen = (int) (__ldg(pat_off + pid + 1) - o);
			BsgpuIndelCand c;

			if (plen <= max_nmis)
				continue;               // not an indels pattern (src/match_pattern.c:98-99)
			if (!bsgpu_indels_candidate(pat + o, plen, first + (size_t) pid * 256, S.bytes + lo, L, p - lo,
						    max_nmis, fixedP, fixedS, &c))
				continue;
			unsigned long long at = atomicAdd(n_out, 1ull);
			if (at < capacity) {
				keys[at] = ((unsigned long long) pid << BSGPU_POS_BITS) | (unsigned long long) p;
				vals[at] = ((uint32_t) c.width << 8) | (uint32_t) c.nedit;
			}
		}
	}
}

static int bsgpu_count_indels_impl(const bsgpu_pdict3parts *plain, const bsgpu_subject *subj, int max_mismatch,
		int fixedP, int fixedS, int per_sequence, int32_t *counts_out)
{
	if (plain == nullptr || subj == nullptr || counts_out == nullptr)
		return fail(BSGPU_EINVAL, "bsgpu_count_indels(): bad arguments");
	if (!plain->plain)
		return fail(BSGPU_EINVAL, "at the moment, matchPDict() and family only support indels on a "
			    "non-preprocessed pattern dictionary, sorry");
	if (max_mismatch < 1 || max_mismatch > BSGPU_INDELS_MAX_K)
		return fail(BSGPU_EUNSUPPORTED, "with.indels=TRUE needs 1 <= max.mismatch <= %d here", BSGPU_INDELS_MAX_K);
	if (subj->kind == SUBJ_CHUNK)
		return fail(BSGPU_EUNSUPPORTED, "with.indels=TRUE is not supported on subject chunks");
	TRY(ensure_device());
	const int32_t M = plain->M;
	const int64_t N = per_sequence ? subj->nseg : 1;

	if ((int64_t) M * N > INT32_MAX)
		return fail(BSGPU_EINVAL, "allocMatrix: too many elements specified");
	for (int64_t i = 0; i < (int64_t) M * N; i++)
		counts_out[i] = 0;
	if (M == 0 || subj->nseg == 0)
		return BSGPU_OK;
	std::vector<int64_t> pat_off((size_t) M + 1);
	CUDA_TRY(cudaMemcpy(pat_off.data(), plain->d_pat_off.p, ((size_t) M + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
	// Patterns not longer than max.mismatch are counted by the reference's mismatch-only loop
	// whatever the algorithm (src/match_pattern.c:95-96): they go through the plain-dictionary
	// kernel as a dictionary of their own; the indels kernels skip them.
	{
		std::vector<int32_t> short_ids, short_w;
		std::vector<uint8_t> short_letters;
		for (int32_t i = 0; i < M; i++) {
			int64_t plen = pat_off[(size_t) i + 1] - pat_off[(size_t) i];

			if (plen <= max_mismatch) {
				size_t at = short_letters.size();

				short_ids.push_back(i);
				short_w.push_back((int32_t) plen);
				short_letters.resize(at + (size_t) plen);
				CUDA_TRY(cudaMemcpy(short_letters.data() + at, plain->d_pat.p + pat_off[(size_t) i], (size_t) plen,
						    cudaMemcpyDeviceToHost));
			}
		}
		if (!short_ids.empty()) {
			bsgpu_pdict3parts *sp = nullptr;
			bsgpu_result r;
			const bsgpu_pdict3parts *spc;

			TRY(bsgpu_patterns_build(short_letters.data(), short_w.data(), (int32_t) short_ids.size(), &sp));
			spc = sp;
			int rc = per_sequence
				? bsgpu_vmatch(&spc, 1, subj, max_mismatch, 0, fixedP, fixedS, 0, 0, nullptr, 0,
					       BSGPU_MATCHES_AS_COUNTS, &r)
				: bsgpu_match(&spc, 1, subj, max_mismatch, 0, fixedP, fixedS, BSGPU_MATCHES_AS_COUNTS, &r);
			bsgpu_pdict3parts_free(sp);
			TRY(rc);
			const int64_t ns = (int64_t) short_ids.size();
			for (int64_t j = 0; j < N; j++)
				for (int64_t k = 0; k < ns; k++)
					counts_out[j * M + short_ids[(size_t) k]] = r.counts[j * ns + k];
			bsgpu_result_free(&r);
		}
	}
	cudaStream_t st = 0;
	DevBuf<int16_t> d_first;
	DevBuf<unsigned long long> d_keys, d_n;
	DevBuf<uint32_t> d_vals;
	unsigned long long n = 0, zero = 0;
	size_t cap = 1u << 20;

	TRY(d_first.alloc((size_t) M * 256));
	indels_first_kernel<<<M, 256, 0, st>>>(plain->d_pat.p, plain->d_pat_off.p, M, fixedP != 0, fixedS != 0, d_first.p);
	g_launches++;
	CUDA_TRY(cudaGetLastError());
	TRY(d_n.alloc(1));
	long long nletters = subj->scan_hi - subj->scan_lo;
	dim3 grid((unsigned) grid_for(nletters, 128, 8), (unsigned) std::min(M, 64));
	for (;;) {
		TRY(d_keys.alloc(cap));
		TRY(d_vals.alloc(cap));
		CUDA_TRY(cudaMemcpyAsync(d_n.p, &zero, sizeof(zero), cudaMemcpyHostToDevice, st));
		{
			KernelTimer kt("indels_candidates_kernel", st);
			indels_candidates_kernel<<<grid, 128, 0, st>>>(subj->view(), plain->d_pat.p, plain->d_pat_off.p, M,
					d_first.p, max_mismatch, fixedP != 0, fixedS != 0, subj->scan_lo, subj->scan_hi,
					d_keys.p, d_vals.p, d_n.p, (unsigned long long) cap);
		}
		g_launches++;
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaMemcpy(&n, d_n.p, sizeof(n), cudaMemcpyDeviceToHost));
		if (n <= cap)
			break;
		cap = (size_t) n + (size_t) n / 8;      // everything is recomputed into the larger buffers
	}
	set_plan("indels_bruteforce patterns=%d letters=%lld candidates=%llu", M, nletters, n);
	if (n == 0)
		return BSGPU_OK;
	std::vector<unsigned long long> keys((size_t) n);
	std::vector<uint32_t> vals((size_t) n);
	CUDA_TRY(cudaMemcpy(keys.data(), d_keys.p, (size_t) n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	CUDA_TRY(cudaMemcpy(vals.data(), d_vals.p, (size_t) n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	bsgpu_indels_counts_from_candidates(keys.data(), vals.data(), (size_t) n, BSGPU_POS_BITS, subj->seg_start, M,
					    per_sequence != 0, counts_out);
	return BSGPU_OK;
}

extern "C" int bsgpu_count_indels(const bsgpu_pdict3parts *plain, const bsgpu_subject *subj, int max_mismatch,
		int fixedP, int fixedS, int per_sequence, int32_t *counts_out)
{
	int rc;

	try {
		rc = bsgpu_count_indels_impl(plain, subj, max_mismatch, fixedP, fixedS, per_sequence, counts_out);
	} catch (const std::bad_alloc &) {
		rc = fail(BSGPU_ENOMEM, "bsgpu_count_indels(): out of host memory");
	} catch (const std::exception &e) {
		rc = fail(BSGPU_ECUDA, "bsgpu_count_indels(): %s", e.what());
	}
	return rc;
}


// ---------------------------------------------------------------------------------------
// oligonucleotideFrequency (XString_oligo_frequency / XStringSet_oligo_frequency,
// src/letter_frequency.c:634-738)
// ---------------------------------------------------------------------------------------

template <typename CT>
static int launch_oligo(const bsgpu_subject *subj, const BsgpuOligoParams &P, CT *d_out, cudaStream_t st)
{
	size_t shmem = P.slice_bits > 0 ? ((size_t) 1 << P.slice_bits) * sizeof(uint32_t)
					: (P.replicas > 0 ? ((size_t) P.replicas << (2 * P.w)) * sizeof(uint32_t) : 0);

	if (shmem > 48 * 1024)
		CUDA_TRY(cudaFuncSetAttribute(oligo_freq_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
					      (int) shmem));
	{
		KernelTimer kt("oligo_freq_kernel", st);
		if (P.slice_bits > 0) {
			// width 8: 2 slices of 2^15 counters (128 KB, one block of 1024 threads per SM);
			// every slice reads the planes again, which is cheap next to one L2 atomic per window
			dim3 grid((unsigned) grid_for(P.nwords, 1024, 1), 1u << (2 * P.w - P.slice_bits));
			oligo_freq_kernel<CT><<<grid, 1024, shmem, st>>>(subj->view(), P, d_out);
		} else {
			oligo_freq_kernel<CT><<<grid_for(P.nwords, 256, 8), 256, shmem, st>>>(subj->view(), P, d_out);
		}
	}
	g_launches++;
	CUDA_TRY(cudaGetLastError());
	return BSGPU_OK;
}

static int check_oligo_width_step(int width, int step)
{
	// src/utils.c:101-102
	if (width < 1 || width > 15)
		return fail(BSGPU_EINVAL, "_new_TwobitEncodingBuffer(): 'buflength' must be >= 1 and <= 15");
	// R/letterFrequency.R:26-34
	if (step < 1)
		return fail(BSGPU_EINVAL, "'step' must be a positive integer");
	return BSGPU_OK;
}

static int check_oligo_args(const bsgpu_subject *subj, int width, int step)
{
	if (subj == nullptr)
		return fail(BSGPU_EINVAL, "bsgpu_oligo_frequency(): NULL subject");
	return check_oligo_width_step(width, step);
}

static BsgpuOligoParams oligo_params(const bsgpu_subject *subj, int width, int step, int invert,
				     bool per_segment)
{
	BsgpuOligoParams P = {};

	P.w = width;
	P.step = step;
	P.invert = invert != 0;
	P.per_segment = per_segment;
	P.coord_step = subj->kind == SUBJ_CHUNK;
	// one row of at most 4^7 counters: per-block tables in shared memory, replicated while
	// they are small so that the lanes of a warp do not pile up on 4 or 16 counters
	if (!per_segment && width <= 7) {
		int r = width <= 6 ? (4096 >> (2 * width)) : 1;

		P.replicas = std::max(1, std::min(32, r));
	}
	// one row of 4^8 counters: too big for one block's shared memory, small enough to be counted in
	// two slices
	P.fast = !per_segment && step == 1 && getenv("BSGPU_OLIGO_GENERIC") == nullptr;
	// width 8: 0.48 ms per 250 Mbp against 1.64 ms through L2 atomics; width 9 would need 8 slices:
	// 1.87 ms against 1.33 ms, so it stays with the atomics (tools/oligo_w9_probe.py)
	if (P.fast && width == 8 && getenv("BSGPU_OLIGO_NO_SLICES") == nullptr)
		P.slice_bits = 15;      // (the unrolled path only: it counts through oligo_bump)
	P.scan_lo = subj->scan_lo;
	P.scan_hi = subj->scan_hi;
	P.nwords = subj->total / 32;
	return P;
}

extern "C" int bsgpu_oligo_count_device(const bsgpu_subject *subj, int width, int step,
		int invert_twobit_order, unsigned long long *d_counts, void *stream)
{
	TRY(check_oligo_args(subj, width, step));
	if (d_counts == nullptr)
		return fail(BSGPU_EINVAL, "bsgpu_oligo_count_device(): NULL d_counts");
	TRY(ensure_device());
	cudaStream_t st = (cudaStream_t) stream;
	TRY(ensure_packed(subj, st));
	set_plan("oligo_freq w=%d step=%d rows=1 table=%s", width, step, width <= 7 ? "shared" : "global");
	return launch_oligo<unsigned long long>(subj, oligo_params(subj, width, step, invert_twobit_order, false),
						d_counts, st);
}

extern "C" int bsgpu_oligo_frequency(const bsgpu_subject *subj, int width, int step,
		int invert_twobit_order, int collapsed, int as_prob, int32_t *counts_out, double *freq_out)
{
	TRY(check_oligo_args(subj, width, step));
	if ((as_prob ? (void *) freq_out : (void *) counts_out) == nullptr)
		return fail(BSGPU_EINVAL, "bsgpu_oligo_frequency(): NULL output");
	TRY(ensure_device());
	const size_t ncol = (size_t) 1 << (2 * width);
	const size_t nrow = collapsed ? 1 : (size_t) subj->nseg;

	if (nrow * ncol > (size_t) INT32_MAX)
		return fail(BSGPU_EINVAL, "allocMatrix: too many elements specified");
	if (nrow == 0)
		return BSGPU_OK;
	cudaStream_t st = 0;
	TRY(ensure_packed(subj, st));
	BsgpuOligoParams P = oligo_params(subj, width, step, invert_twobit_order, !collapsed);
	set_plan("oligo_freq w=%d step=%d rows=%zu table=%s", width, step, nrow,
		 P.slice_bits ? "shared_slices" : (P.replicas ? "shared" : "global"));
	if (collapsed) {
		// 64-bit counters: a collapsed set may hold more than 2^31 windows of one kind
		DevBuf<unsigned long long> d_out;
		std::vector<unsigned long long> h(ncol);

		TRY(d_out.alloc(ncol));
		CUDA_TRY(cudaMemsetAsync(d_out.p, 0, ncol * sizeof(unsigned long long), st));
		TRY(launch_oligo<unsigned long long>(subj, P, d_out.p, st));
		CUDA_TRY(cudaMemcpyAsync(h.data(), d_out.p, ncol * sizeof(unsigned long long),
					 cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		if (!as_prob) {
			for (size_t c = 0; c < ncol; c++)
				counts_out[c] = (int32_t) (uint32_t) h[c];      // the reference's int wraps too
			return BSGPU_OK;
		}
		// normalize_oligo_freqs(), src/letter_frequency.c:286-301: sum left to right, a row
		// whose sum is 0 is left alone
		double sum = 0.0;
		for (size_t c = 0; c < ncol; c++)
			sum += (freq_out[c] = (double) h[c]);
		if (sum != 0.0)
			for (size_t c = 0; c < ncol; c++)
				freq_out[c] /= sum;
		return BSGPU_OK;
	}
	// one row per sequence: row-major on the device (the counters of a sequence are
	// neighbours), column-major for the caller as allocMatrix() lays it out
	DevBuf<uint32_t> d_out;
	std::vector<uint32_t> h(nrow * ncol);

	TRY(d_out.alloc(nrow * ncol));
	CUDA_TRY(cudaMemsetAsync(d_out.p, 0, nrow * ncol * sizeof(uint32_t), st));
	TRY(launch_oligo<uint32_t>(subj, P, d_out.p, st));
	CUDA_TRY(cudaMemcpyAsync(h.data(), d_out.p, nrow * ncol * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	for (size_t i = 0; i < nrow; i++) {
		const uint32_t *row = h.data() + i * ncol;

		if (!as_prob) {
			for (size_t c = 0; c < ncol; c++)
				counts_out[i + c * nrow] = (int32_t) row[c];
			continue;
		}
		double sum = 0.0;
		for (size_t c = 0; c < ncol; c++)
			sum += (double) row[c];
		for (size_t c = 0; c < ncol; c++)
			freq_out[i + c * nrow] = sum != 0.0 ? (double) row[c] / sum : (double) row[c];
	}
	return BSGPU_OK;
}

// XString_oligo_frequency, src/letter_frequency.c:634
extern "C" int bsgpu_XString_oligo_frequency(const uint8_t *x, int64_t x_length, int width, int step,
		int invert_twobit_order, int as_prob, int32_t *counts_out, double *freq_out)
{
	bsgpu_subject *subj = nullptr;

	TRY(check_oligo_width_step(width, step));
	TRY(bsgpu_subject_xstring(x, x_length, &subj));
	int rc = bsgpu_oligo_frequency(subj, width, step, invert_twobit_order, 1, as_prob, counts_out, freq_out);
	bsgpu_subject_free(subj);
	return rc;
}

// XStringSet_oligo_frequency, src/letter_frequency.c:667
extern "C" int bsgpu_XStringSet_oligo_frequency(const uint8_t *concat, const int32_t *widths, int32_t n,
		int width, int step, int invert_twobit_order, int collapsed, int as_prob,
		int32_t *counts_out, double *freq_out)
{
	bsgpu_subject *subj = nullptr;

	TRY(check_oligo_width_step(width, step));
	TRY(bsgpu_subject_set(concat, widths, n, &subj));
	int rc = bsgpu_oligo_frequency(subj, width, step, invert_twobit_order, collapsed, as_prob,
				       counts_out, freq_out);
	bsgpu_subject_free(subj);
	return rc;
}

// ---------------------------------------------------------------------------------------
// One count vector for the GPUs of a box (include/bsgpu.h): CUDA IPC + peer access; the scan
// kernels count with system-scope reductions (count_add, bsgpu_kernels.cuh), so a vector that
// lives on another GPU is incremented over NVLink by the very launch that finds the match.
// ---------------------------------------------------------------------------------------

static_assert(sizeof(cudaIpcMemHandle_t) == BSGPU_IPC_HANDLE_BYTES, "IPC handle size");

extern "C" int bsgpu_shared_counts_create(int64_t n, int32_t **d_counts_out, unsigned char *handle_out)
{
	if (n <= 0 || d_counts_out == nullptr || handle_out == nullptr)
		return fail(BSGPU_EINVAL, "bsgpu_shared_counts_create(): bad arguments");
	TRY(ensure_device());
	int32_t *p = nullptr;
	cudaIpcMemHandle_t h;

	// a plain cudaMalloc block of its own: IPC handles name whole allocations
	CUDA_TRY(cudaMalloc((void **) &p, (size_t) n * sizeof(int32_t)));
	if (cudaMemset(p, 0, (size_t) n * sizeof(int32_t)) != cudaSuccess || cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
		cudaError_t e = cudaGetLastError();

		cudaFree(p);
		return fail(BSGPU_ECUDA, "bsgpu_shared_counts_create(): %s", cudaGetErrorString(e));
	}
	memcpy(handle_out, &h, sizeof(h));
	*d_counts_out = p;
	return BSGPU_OK;
}

extern "C" int bsgpu_shared_counts_open(const unsigned char *handle, int32_t **d_counts_out)
{
	if (handle == nullptr || d_counts_out == nullptr)
		return fail(BSGPU_EINVAL, "bsgpu_shared_counts_open(): bad arguments");
	TRY(ensure_device());
	cudaIpcMemHandle_t h;
	void *p = nullptr;

	memcpy(&h, handle, sizeof(h));
	CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
	*d_counts_out = (int32_t *) p;
	return BSGPU_OK;
}

extern "C" int bsgpu_shared_counts_close(int32_t *d_counts, int is_owner)
{
	if (d_counts == nullptr)
		return BSGPU_OK;
	if (is_owner)
		CUDA_TRY(cudaFree(d_counts));
	else
		CUDA_TRY(cudaIpcCloseMemHandle(d_counts));
	return BSGPU_OK;
}

// ---------------------------------------------------------------------------------------
// Roofline probes
// ---------------------------------------------------------------------------------------

extern "C" int bsgpu_probe_peaks(double *copy_gbs, double *l2_gather_gsectors, int64_t table_bytes)
{
	TRY(ensure_device());
	cudaEvent_t e0, e1;
	CUDA_TRY(cudaEventCreate(&e0));
	CUDA_TRY(cudaEventCreate(&e1));
	float ms = 0;
	if (copy_gbs) {
		const int64_t n = (1ll << 30) / 16;     // 1 GiB
		DevBuf<uint4> a, b;
		TRY(a.alloc((size_t) n));
		TRY(b.alloc((size_t) n));
		CUDA_TRY(cudaMemset(a.p, 1, (size_t) n * 16));
		double best = 0;
		for (int it = 0; it < 6; it++) {
			CUDA_TRY(cudaEventRecord(e0));
			probe_copy_kernel<<<g_num_sms * 16, 256>>>(a.p, b.p, n);
			g_launches++;
			CUDA_TRY(cudaEventRecord(e1));
			CUDA_TRY(cudaEventSynchronize(e1));
			CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
			best = std::max(best, 2.0 * n * 16 / (ms * 1e-3) / 1e9);
		}
		*copy_gbs = best;
	}
	if (l2_gather_gsectors) {
		int64_t nb = 16;
		while (nb * 2 * 16 <= table_bytes)
			nb *= 2;
		DevBuf<uint4> t;
		DevBuf<unsigned long long> sink;
		TRY(t.alloc((size_t) nb));
		TRY(sink.alloc(1));
		CUDA_TRY(cudaMemset(t.p, 0, (size_t) nb * 16));
		const int iters = 4096, threads = 256;
		double best = 0;
		// best over a few launch shapes (loads in flight per thread x blocks per SM)
		for (int cfg = 0; cfg < 6; cfg++) {
			int blocks = g_num_sms * (cfg % 2 == 0 ? 8 : 4);
			for (int it = 0; it < 4; it++) {
				CUDA_TRY(cudaEventRecord(e0));
				switch (cfg / 2) {
				case 0: probe_gather_kernel<4><<<blocks, threads>>>(t.p, (uint32_t) (nb - 1), iters, sink.p); break;
				case 1: probe_gather_kernel<8><<<blocks, threads>>>(t.p, (uint32_t) (nb - 1), iters, sink.p); break;
				default: probe_gather_kernel<16><<<blocks, threads>>>(t.p, (uint32_t) (nb - 1), iters, sink.p); break;
				}
				g_launches++;
				CUDA_TRY(cudaEventRecord(e1));
				CUDA_TRY(cudaEventSynchronize(e1));
				CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
				best = std::max(best, (double) blocks * threads * iters / (ms * 1e-3) / 1e9);
			}
		}
		*l2_gather_gsectors = best;
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	return BSGPU_OK;
}
