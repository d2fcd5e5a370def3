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

// UDB word of w letters starting at s (udbparams.cpp:540-555): upper-case ACGTU only.
static inline uint32_t udb_word(const uint8_t *s, uint32_t w)
{
	uint32_t word = 0;
	for (uint32_t i = 0; i < w; ++i) {
		uint32_t l;
		switch (s[i]) {
		case 'A': l = 0; break;
		case 'C': l = 1; break;
		case 'G': l = 2; break;
		case 'T': case 'U': l = 3; break;
		default: return UINT32_MAX;
		}
		word = (word << 2) | l;
	}
	return word;
}

namespace {
struct Worker {
	uint32_t t0, t1;
	std::vector<uint32_t> counts; // per word, this worker's targets
	std::vector<uint32_t> stamp;
};
}

template <class F>
static void for_each_unique_word(const HostIndex &ix, uint32_t t, std::vector<uint32_t> &stamp, F f)
{
	const uint8_t *s = ix.seqs.data() + ix.seq_off[t];
	uint32_t L = ix.seq_len[t], w = ix.word_length;
	if (L < w)
		return;
	for (uint32_t p = 0; p + w <= L; ++p) {
		uint32_t word = udb_word(s + p, w);
		if (word == UINT32_MAX || stamp[word] == t + 1)
			continue;
		stamp[word] = t + 1;
		f(word);
	}
}

void build_host_index(const uint8_t *seqs, const uint64_t *seq_off, uint32_t n_seq, uint32_t word_length,
  int dbmask, int n_threads, HostIndex &ix)
{
	ix.n_seq = n_seq;
	ix.word_length = word_length;
	ix.slots = 1u << (2 * word_length);
	ix.seq_off.assign((size_t)n_seq + 1, 0);
	ix.seq_len.assign(n_seq, 0);
	ix.max_len = 0;
	uint64_t off = 0;
	for (uint32_t t = 0; t < n_seq; ++t) {
		uint64_t L = seq_off[t + 1] - seq_off[t];
		ix.seq_off[t] = off;
		ix.seq_len[t] = (uint32_t)L;
		ix.max_len = std::max(ix.max_len, (uint32_t)L);
		off += (L + 15) & ~(uint64_t)15;
	}
	ix.seq_off[n_seq] = off;
	ix.seqs.assign(off + 16, 0);

	unsigned T = n_threads > 0 ? (unsigned)n_threads : std::max(1u, std::thread::hardware_concurrency());
	T = std::min<unsigned>(T, std::max<uint32_t>(1, n_seq / 64));
	T = std::max(1u, std::min(T, 64u));
	std::vector<Worker> W(T);
	for (unsigned k = 0; k < T; ++k) {
		W[k].t0 = (uint32_t)((uint64_t)n_seq * k / T);
		W[k].t1 = (uint32_t)((uint64_t)n_seq * (k + 1) / T);
	}
	auto run = [&](auto fn) {
		std::vector<std::thread> th;
		for (unsigned k = 1; k < T; ++k)
			th.emplace_back(fn, k);
		fn(0u);
		for (auto &t : th)
			t.join();
	};
	// pass 1: copy + mask, count unique words per worker
	run([&](unsigned k) {
		Worker &w = W[k];
		w.counts.assign(ix.slots, 0);
		w.stamp.assign(ix.slots, 0);
		for (uint32_t t = w.t0; t < w.t1; ++t) {
			const uint8_t *src = seqs + seq_off[t];
			uint8_t *dst = ix.seqs.data() + ix.seq_off[t];
			if (dbmask)
				fastmask_nt(src, ix.seq_len[t], dst);
			else if (ix.seq_len[t])
				memcpy(dst, src, ix.seq_len[t]);
			for_each_unique_word(ix, t, w.stamp, [&](uint32_t word) { ++w.counts[word]; });
		}
	});
	// row offsets; each worker's counts become its write cursor inside the row
	ix.row_off.assign((size_t)ix.slots + 1, 0);
	ix.row_size.assign(ix.slots, 0);
	uint64_t total = 0;
	ix.n_postings = 0;
	for (uint32_t word = 0; word < ix.slots; ++word) {
		ix.row_off[word] = total;
		for (unsigned k = 0; k < T; ++k) {
			uint32_t c = W[k].counts[word];
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
	run([&](unsigned k) {
		Worker &w = W[k];
		std::fill(w.stamp.begin(), w.stamp.end(), 0u);
		for (uint32_t t = w.t0; t < w.t1; ++t)
			for_each_unique_word(ix, t, w.stamp, [&](uint32_t word) {
				ix.postings[ix.row_off[word] + w.counts[word]++] = t;
			});
	});
}

} // namespace usb
