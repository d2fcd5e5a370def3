// usb_hostindex.h -- host-side construction of the UDB index (CSR postings) and masked SeqDB.
//
// Replaces, for the in-memory layout: MaskDB (makeudb.cpp:11-25) -> FastMaskSeq
// (fastmask.cpp:88-158), UDBParams::SetTargetWords/SetTargetUniqueWords (udbparams.cpp:644-711),
// UDBData::FromSeqDB (udbbuild.cpp:303-398) and, for the growing cluster_fast database,
// UDBData::AddSIToDB_CopyData / AddWord / GrowRow (udbbuild.cpp:74-130,286).
// The reference keeps one heap row per word (m_UDBRows[word], m_Sizes[word]) and grows rows in
// place; here the index is a short list of immutable CSR segments over consecutive target ranges
// (a log-structured merge: appended targets form a new small segment, neighbouring segments of
// similar size are concatenated), so that every posting row fragment stays a contiguous, 16-byte
// aligned HBM range and an append never rewrites more than an amortised O(log N) share.
#pragma once
#include <stdint.h>
#include <vector>
#include "usb_tables.h"

namespace usb {

// Masked target letters; each target padded to a 16-byte boundary.
struct HostSeqs {
	std::vector<uint8_t> seqs;
	std::vector<uint64_t> seq_off{0};   // n+1 padded offsets (multiples of 16)
	std::vector<uint32_t> seq_len;      // true lengths
	uint32_t max_len = 0;
	uint32_t n() const { return (uint32_t)seq_len.size(); }
	// dbmask: 1 = fastnucleo soft-masking, 0 = letters taken verbatim (cluster_fast)
	void append(const uint8_t *s, const uint64_t *off, uint32_t count, int dbmask, int n_threads);
};

// Postings of the targets [base, base+count): row_off[slots+1] (multiples of 4 entries),
// row_size[slots] true lengths, postings = global target indexes ascending per row, every target
// at most once per row; padding entries are 0xffffffff.
struct HostCSR {
	uint32_t base = 0, count = 0, slots = 0;
	std::vector<uint64_t> row_off;
	std::vector<uint32_t> row_size;
	std::vector<uint32_t> postings;
	uint64_t n_postings = 0;
};

// The same index with 2-byte postings in a bank-aware order, for k_rank on indexes of at most
// USB_HALF_MAX_TARGETS targets starting at target 0.  Halves the HBM bytes of the U-sort walk and
// removes most shared-memory bank conflicts of its counter increments, which is what bounds the
// walk on B200 (profiles/README.md).
//   * Targets are cut into blocks of 65 535; a word has one row fragment per block, holding
//     target - 65535 * block (<= 65 534; 0xffff marks padding).  Fragment f = word * n_blocks +
//     block: row_off[f] (in entries, multiples of 8 = 16 bytes), row_size[f].
//   * The order inside a fragment does not matter for U.  It is chosen for the way a warp walks
//     it: lane l loads the 8-entry vector 32 * s + l of slot s and the warp then issues one
//     increment instruction per entry index i = 0..7, i.e. for entries {256 s + 8 l + i}.  The
//     fragment's postings are dealt round-robin over the 32 banks of their 1-byte counters
//     (bank = (target / 4) % 32) and sequence index q = 256 s + 32 i + l goes to entry
//     256 s + 8 l + i, so the 32 increments of one instruction fall into (mostly) different banks.
//     The last m = size % 256 entries use nv = ceil(m / 8) vectors: q -> entry 8 (q % nv) + q / nv.
#define USB_HALF_BLOCK 65535u
#define USB_HALF_MAX_TARGETS (2u * USB_HALF_BLOCK)
struct HostHalf {
	uint32_t n_blocks = 1;
	std::vector<uint64_t> row_off;   // slots * n_blocks + 1
	std::vector<uint32_t> row_size;  // slots * n_blocks
	std::vector<uint16_t> postings;
};
void make_half(const HostCSR &H, uint32_t n_targets, int n_threads, HostHalf &out);

// FastMaskSeq soft-masking (fastmask.cpp:88-158); in == out allowed.
void fastmask_nt(const uint8_t *in, uint32_t L, uint8_t *out);

// UDB alphabet (udbparams.cpp:235-261): alpha = 4 (ACGT, U = T) or 20 (ACDEFGHIKLMNPQRSTVWY);
// a word is `word_length` letters read as a base-alpha number, slots = alpha^word_length.
// Returns the 256-entry character -> letter table (0xff = lower case / not in the alphabet).
const uint8_t *udb_letters(uint32_t alpha);
uint32_t udb_slots(uint32_t alpha, uint32_t word_length);
// word starting at s, or UINT32_MAX when a letter is bad (udbparams.cpp:540-555)
inline uint32_t udb_word(const uint8_t *s, uint32_t w, uint32_t alpha, const uint8_t *letters)
{
	uint32_t word = 0;
	for (uint32_t i = 0; i < w; ++i) {
		const uint32_t l = letters[s[i]];
		if (l == 0xff)
			return UINT32_MAX;
		word = word * alpha + l;
	}
	return word;
}

// n_threads <= 0: hardware concurrency.
void build_csr(const HostSeqs &S, uint32_t first, uint32_t count, uint32_t word_length, uint32_t alpha, int n_threads,
  HostCSR &out);
// a covers [x, y), b covers [y, z): out covers [x, z) with rows = a's row followed by b's row.
void merge_csr(const HostCSR &a, const HostCSR &b, HostCSR &out);

} // namespace usb
