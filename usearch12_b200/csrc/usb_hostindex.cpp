// usb_hostindex.cpp -- see usb_hostindex.h
#include "usb_hostindex.h"
#include <algorithm>
#include <cstring>
#include <thread>

namespace usb {

static inline uint8_t up(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }
static inline uint8_t low(uint8_t c) { return (c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c; }

// fastmask.cpp:88-158.  Runs of one letter and of one letter pair are detected on the
// upper-cased input; a run of period p measured as n >= 5 positions gets letters
// [start+2, end) lower-cased.  The measuring quirks of the reference are kept: the first
// comparison uses an unsigned wrap-around "start" (so nothing is masked before the first
// change), a homopolymer reaching the last letter is closed at the last index (one short), and
// pair runs are never flushed at the end of the sequence.
void fastmask_nt(const uint8_t *in, uint32_t L, uint8_t *out)
{
	std::vector<uint8_t> u(L);
	for (uint32_t i = 0; i < L; ++i)
		u[i] = up(in[i]);
	std::vector<uint8_t> m(u);
	auto soften = [&](uint32_t from, uint32_t to) {
		for (uint32_t j = from; j < to; ++j)
			m[j] = low(m[j]);
	};
	if (L >= 2) {
		uint32_t run_start = UINT32_MAX;
		int prev = -1;
		for (uint32_t i = 0; i < L; ++i) {
			if ((int)u[i] != prev || i + 1 == L) {
				if (i - run_start >= 5)
					soften(run_start + 2, i);
				run_start = i;
			}
			prev = u[i];
		}
		for (uint32_t phase = 0; phase < 2; ++phase) {
			uint32_t pair_start = UINT32_MAX;
			int prev_pair = -1;
			for (uint32_t i = phase; i + 1 < L; i += 2) {
				int pair = (u[i] << 8) | u[i + 1];
				if (pair != prev_pair) {
					if (i - pair_start >= 5)
						soften(pair_start + 2, i);
					pair_start = i;
				}
				prev_pair = pair;
			}
		}
	}
	if (L)
		memcpy(out, m.data(), L);
}

namespace {
struct UdbLetterTables {
	uint8_t nt[256], aa[256];
	UdbLetterTables()
	{
		memset(nt, 0xff, sizeof nt);
		memset(aa, 0xff, sizeof aa);
		nt['A'] = 0; nt['C'] = 1; nt['G'] = 2; nt['T'] = 3; nt['U'] = 3;
		const char *a = "ACDEFGHIKLMNPQRSTVWY"; // alpha.cpp g_CharToLetterAmino
		for (int i = 0; i < 20; ++i)
			aa[(int)a[i]] = (uint8_t)i;
	}
};
}

const uint8_t *udb_letters(uint32_t alpha)
{
	static const UdbLetterTables T; // function-local static: initialised once, thread-safe
	return alpha == 4 ? T.nt : T.aa;
}

uint32_t udb_slots(uint32_t alpha, uint32_t word_length)
{
	uint64_t n = 1;
	for (uint32_t i = 0; i < word_length; ++i)
		n *= alpha;
	return n > 0xffffffffull ? 0u : (uint32_t)n;
}

static unsigned pick_threads(int n_threads, uint32_t items)
{
	unsigned T = n_threads > 0 ? (unsigned)n_threads : std::max(1u, std::thread::hardware_concurrency());
	T = std::min<unsigned>(T, std::max<uint32_t>(1, items / 256));
	return std::max(1u, std::min(T, 64u));
}

template <class F> static void run_threads(unsigned T, F fn)
{
	std::vector<std::thread> th;
	for (unsigned k = 1; k < T; ++k)
		th.emplace_back(fn, k);
	fn(0u);
	for (auto &t : th)
		t.join();
}

void HostSeqs::append(const uint8_t *s, const uint64_t *off, uint32_t count, int dbmask, int n_threads)
{
	const uint32_t n0 = n();
	uint64_t cur = seq_off.back();
	for (uint32_t i = 0; i < count; ++i) {
		const uint64_t L = off[i + 1] - off[i];
		seq_len.push_back((uint32_t)L);
		max_len = std::max(max_len, (uint32_t)L);
		cur += (L + 15) & ~(uint64_t)15;
		seq_off.push_back(cur);
	}
	seqs.resize(cur + 16, 0);
	const unsigned T = pick_threads(n_threads, count);
	run_threads(T, [&](unsigned k) {
		const uint32_t a = (uint32_t)((uint64_t)count * k / T), b = (uint32_t)((uint64_t)count * (k + 1) / T);
		for (uint32_t i = a; i < b; ++i) {
			const uint8_t *src = s + off[i];
			uint8_t *dst = seqs.data() + seq_off[n0 + i];
			const uint32_t L = seq_len[n0 + i];
			if (dbmask)
				fastmask_nt(src, L, dst);
			else if (L)
				memcpy(dst, src, L);
		}
	});
}

namespace {
struct Worker {
	uint32_t t0, t1;
	std::vector<uint32_t> counts; // per word, this worker's targets
	std::vector<uint32_t> stamp;
};
}

template <class F>
static void for_each_unique_word(const HostSeqs &S, uint32_t word_length, uint32_t alpha, uint32_t t, uint32_t stamp_id,
  std::vector<uint32_t> &stamp, F f)
{
	const uint8_t *letters = udb_letters(alpha);
	const uint8_t *s = S.seqs.data() + S.seq_off[t];
	const uint32_t L = S.seq_len[t], w = word_length;
	if (L < w)
		return;
	for (uint32_t p = 0; p + w <= L; ++p) {
		const uint32_t word = udb_word(s + p, w, alpha, letters);
		if (word == UINT32_MAX || stamp[word] == stamp_id)
			continue;
		stamp[word] = stamp_id;
		f(word);
	}
}

void build_csr(const HostSeqs &S, uint32_t first, uint32_t count, uint32_t word_length, uint32_t alpha, int n_threads,
  HostCSR &ix)
{
	ix.base = first;
	ix.count = count;
	ix.slots = udb_slots(alpha, word_length);
	const unsigned T = pick_threads(n_threads, count);
	std::vector<Worker> W(T);
	for (unsigned k = 0; k < T; ++k) {
		W[k].t0 = first + (uint32_t)((uint64_t)count * k / T);
		W[k].t1 = first + (uint32_t)((uint64_t)count * (k + 1) / T);
	}
	// pass 1: count unique words per worker
	run_threads(T, [&](unsigned k) {
		Worker &w = W[k];
		w.counts.assign(ix.slots, 0);
		w.stamp.assign(ix.slots, 0);
		for (uint32_t t = w.t0; t < w.t1; ++t)
			for_each_unique_word(S, word_length, alpha, t, t - first + 1, w.stamp, [&](uint32_t word) { ++w.counts[word]; });
	});
	// row offsets; each worker's counts become its write cursor inside the row
	ix.row_off.assign((size_t)ix.slots + 1, 0);
	ix.row_size.assign(ix.slots, 0);
	uint64_t total = 0;
	ix.n_postings = 0;
	for (uint32_t word = 0; word < ix.slots; ++word) {
		ix.row_off[word] = total;
		for (unsigned k = 0; k < T; ++k) {
			const uint32_t c = W[k].counts[word];
			W[k].counts[word] = (uint32_t)(total - ix.row_off[word]);
			total += c;
		}
		ix.row_size[word] = (uint32_t)(total - ix.row_off[word]);
		ix.n_postings += ix.row_size[word];
		total = (total + 3) & ~(uint64_t)3; // 16-byte aligned rows for vector loads
	}
	ix.row_off[ix.slots] = total;
	ix.postings.assign(total + 4, 0xffffffffu);
	// pass 2: fill (targets ascending within each row because workers own ascending ranges)
	run_threads(T, [&](unsigned k) {
		Worker &w = W[k];
		std::fill(w.stamp.begin(), w.stamp.end(), 0u);
		for (uint32_t t = w.t0; t < w.t1; ++t)
			for_each_unique_word(S, word_length, alpha, t, t - first + 1, w.stamp,
			  [&](uint32_t word) { ix.postings[ix.row_off[word] + w.counts[word]++] = t; });
	});
}

void make_half(const HostCSR &H, uint32_t n_targets, int n_threads, HostHalf &out)
{
	out.dummy0 = ((n_targets + 15) & ~15u) / 4;
	out.row_off.assign((size_t)H.slots + 1, 0);
	out.row_groups.assign(H.slots, 0);
	const unsigned T = pick_threads(n_threads, H.slots);
	// groups per row: the fullest byte class decides
	run_threads(T, [&](unsigned k) {
		const uint32_t a = (uint32_t)((uint64_t)H.slots * k / T), b = (uint32_t)((uint64_t)H.slots * (k + 1) / T);
		for (uint32_t w = a; w < b; ++w) {
			const uint32_t *src = H.postings.data() + H.row_off[w];
			uint32_t c[4] = {0, 0, 0, 0};
			for (uint32_t i = 0; i < H.row_size[w]; ++i)
				++c[src[i] & 3];
			const uint32_t m = std::max(std::max(c[0], c[1]), std::max(c[2], c[3]));
			out.row_groups[w] = (m + 63) / 64;
		}
	});
	uint64_t total = 0;
	for (uint32_t w = 0; w < H.slots; ++w) {
		out.row_off[w] = total;
		total += (uint64_t)out.row_groups[w] * 256;
	}
	out.row_off[H.slots] = total;
	out.postings.assign(total + 8, 0);
	run_threads(T, [&](unsigned k) {
		const uint32_t a = (uint32_t)((uint64_t)H.slots * k / T), b = (uint32_t)((uint64_t)H.slots * (k + 1) / T);
		std::vector<uint32_t> seq[4], bucket;
		for (uint32_t w = a; w < b; ++w) {
			const uint32_t G = out.row_groups[w];
			if (G == 0)
				continue;
			const uint32_t *src = H.postings.data() + H.row_off[w];
			const uint32_t n = H.row_size[w];
			// per byte class: word indexes dealt round-robin over the banks
			for (uint32_t cls = 0; cls < 4; ++cls) {
				uint32_t cnt[32] = {0}, start[33], cur[32], maxc = 0;
				for (uint32_t i = 0; i < n; ++i)
					if ((src[i] & 3) == cls)
						++cnt[(src[i] >> 2) & 31];
				start[0] = 0;
				for (int j = 0; j < 32; ++j) {
					start[j + 1] = start[j] + cnt[j];
					cur[j] = start[j];
					maxc = std::max(maxc, cnt[j]);
				}
				bucket.resize(start[32]);
				for (uint32_t i = 0; i < n; ++i)
					if ((src[i] & 3) == cls)
						bucket[cur[(src[i] >> 2) & 31]++] = src[i] >> 2;
				seq[cls].clear();
				for (uint32_t r = 0; r < maxc; ++r)
					for (int j = 0; j < 32; ++j)
						if (r < cnt[j])
							seq[cls].push_back(bucket[start[j] + r]);
			}
			uint16_t *dst = out.postings.data() + out.row_off[w];
			for (uint32_t g = 0; g < G; ++g)
				for (uint32_t i = 0; i < 8; ++i) {
					const std::vector<uint32_t> &sq = seq[i >> 1];
					for (uint32_t l = 0; l < 32; ++l) {
						const uint32_t q = g * 64 + (i & 1) * 32 + l;
						dst[g * 256 + 8 * l + i] = (uint16_t)(q < sq.size() ? sq[q] : out.dummy0 + l);
					}
				}
		}
	});
}

void merge_csr(const HostCSR &a, const HostCSR &b, HostCSR &out)
{
	out.base = a.base;
	out.count = a.count + b.count;
	out.slots = a.slots;
	out.row_off.assign((size_t)a.slots + 1, 0);
	out.row_size.assign(a.slots, 0);
	uint64_t total = 0;
	for (uint32_t w = 0; w < a.slots; ++w) {
		out.row_off[w] = total;
		out.row_size[w] = a.row_size[w] + b.row_size[w];
		total = (total + out.row_size[w] + 3) & ~(uint64_t)3;
	}
	out.row_off[a.slots] = total;
	out.n_postings = a.n_postings + b.n_postings;
	out.postings.assign(total + 4, 0xffffffffu);
	for (uint32_t w = 0; w < a.slots; ++w) {
		uint32_t *dst = out.postings.data() + out.row_off[w];
		if (a.row_size[w])
			memcpy(dst, a.postings.data() + a.row_off[w], (size_t)a.row_size[w] * 4);
		if (b.row_size[w])
			memcpy(dst + a.row_size[w], b.postings.data() + b.row_off[w], (size_t)b.row_size[w] * 4);
	}
}

} // namespace usb
